import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, __graft_entry__ as ge
import bench
os.makedirs('/tmp/pk_bench', exist_ok=True)
pkg, synth, cfg, wp = bench.make_checkpoint('/tmp/pk_bench')
eng = pkg.Engine(cfg, wp, 0)
pcms = [synth.make_audio(160000, 1000 + i) for i in range(64)]
buf = torch.from_numpy(np.concatenate(pcms)).pin_memory().numpy()
off = np.arange(65, dtype=np.int64) * 160000
dec = pkg.Decoder.TDT
tok = eng._tokens(64)
for _ in range(4):
    eng.transcribe_packed(buf, off, dec, tok)
import ctypes as C
from parakeet_cpp_b200.engine import _f32p, _i64p
N = 20
ts = np.zeros((N, 4))
for i in range(N):
    eng.flush_l2(); eng.sync()
    t0 = time.perf_counter()
    eng.L.pk_stage_pcm(eng.h, _f32p(buf), _i64p(off), 64)
    t1 = time.perf_counter()
    eng.L.pk_run_staged(eng.h, int(dec))
    t2 = time.perf_counter()
    eng.L.pk_fetch_tokens(eng.h, C.byref(tok[0]))
    t3 = time.perf_counter()
    ts[i] = [t1 - t0, t2 - t1, t3 - t2, t3 - t0]
print('ms: stage(enqueue) %.3f  run(enqueue) %.3f  fetch(wait) %.3f  total %.3f' % tuple(1e3 * ts[5:].mean(0)))
# device-only for comparison
eng.stage(np.concatenate(pcms), off)
for _ in range(3): eng.run_staged(dec)
eng.sync()
t0 = time.perf_counter()
for _ in range(10): eng.run_staged(dec)
eng.sync()
print('resident ms/step (no flush) %.3f' % (1e2 * (time.perf_counter() - t0)))
