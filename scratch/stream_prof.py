"""Per-kernel-class device time of one streaming step (eou-120m, S streams), eager launches with CUDA-event scopes."""
import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, __graft_entry__ as ge
pkg = ge.load_package()
from parakeet_cpp_b200 import synth
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = pkg.make_eou_120m_config(max_batch=max(S, 8))
wp = '/tmp/pk_bench/pkeou120m_seed0.safetensors'
os.makedirs('/tmp/pk_bench', exist_ok=True)
if not os.path.exists(wp):
    synth.save_safetensors(wp, synth.make_weights(cfg, seed=0))
eng = pkg.Engine(cfg, wp, 0)
eng.stream_open(S, 2560)
x = [synth.make_audio(2560 * 40, 1200 + (i % 8)) for i in range(S)]
out = eng._tokens(S)
for k in range(30):
    eng.stream_step([a[k * 2560:(k + 1) * 2560] for a in x], out=out, raw=True)
eng.sync()
eng.profile_begin()
t0 = time.perf_counter()
for k in range(30, 36):
    eng.stream_step([a[k * 2560:(k + 1) * 2560] for a in x], out=out, raw=True)
eng.sync()
wall = (time.perf_counter() - t0) / 6
prof = eng.profile_end()
print('S', S, 'eager wall ms/step', 1e3 * wall)
for k, v in prof.items():
    print(f'   {k:10s} {v[0] / 6:8.3f} ms/step  {v[1] // 6:4d} launches/step')
t0 = time.perf_counter()
for k in range(36, 40):
    eng.stream_step([a[k * 2560:(k + 1) * 2560] for a in x], out=out, raw=True)
eng.sync()
print('graph wall ms/step', 1e3 * (time.perf_counter() - t0) / 4)
a = eng.tdt_phases(); print('tdt steps of last chunk', a[7], 'cycles', a[:7].sum())
