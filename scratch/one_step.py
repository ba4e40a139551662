"""One 64 x 10 s batch through the engine with plain launches (PK_GRAPH=0), for ncu captures of the small kernels:
    ncu --set full ... -k regex:<kernel> python scratch/one_step.py ctc|tdt [steps]"""
import os, sys
sys.path.insert(0, '/root/repo')
os.environ.setdefault('PK_GRAPH', '0')
import numpy as np
import __graft_entry__ as ge
import bench
os.makedirs('/tmp/pk_bench', exist_ok=True)
pkg, synth, cfg, wp = bench.make_checkpoint('/tmp/pk_bench')
eng = pkg.Engine(cfg, wp, 0)
pcms = [synth.make_audio(160000, 1000 + i) for i in range(64)]
buf = np.concatenate(pcms); off = np.arange(65, dtype=np.int64) * 160000
dec = pkg.Decoder.CTC if (len(sys.argv) > 1 and sys.argv[1] == 'ctc') else pkg.Decoder.TDT
eng.stage(buf, off)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    eng.run_staged(dec)
eng.sync()
print('tokens', sum(len(t) for t in eng.fetch(64)))
