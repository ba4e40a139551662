"""Builds a variant of the library (extra -D flags) next to the product one:  python scratch/build_variant.py NAME -DFOO=1 ...
-> parakeet.cpp_b200/libparakeet_b200_NAME.so (git-ignored; load it with PK_LIB=...)."""
import importlib.util, os, sys
here = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("pk_build", os.path.join(here, "..", "parakeet.cpp_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
name, flags = sys.argv[1], sys.argv[2:]
b.OBJ = os.path.join(b.HERE, "build_" + name)
b.LIB = os.path.join(b.HERE, f"libparakeet_b200_{name}.so")
b.NVCC_FLAGS = b.NVCC_FLAGS + flags
print(b.build(force=False, verbose=False))
