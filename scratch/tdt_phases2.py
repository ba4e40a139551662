"""Cycles CTA 0 of the TDT decode spends per phase (P1 LSTM, B1, P2 joint hidden, B2, P3 logits, B3, P4), both batch configs."""
import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, __graft_entry__ as ge
import bench
os.makedirs('/tmp/pk_bench', exist_ok=True)
for name in sys.argv[1:] or ['110m-64x10s', '600m-16x30s']:
    conf = bench.CONFIGS[name]
    pkg, synth, cfg, wp = bench.make_checkpoint('/tmp/pk_bench', conf)
    eng = pkg.Engine(cfg, wp, 0)
    B, n = conf['batch'], conf['clip_samples']
    pcms = [synth.make_audio(n, 1000 + i) for i in range(B)]
    buf = np.concatenate(pcms); off = np.arange(B + 1, dtype=np.int64) * n
    eng.stage(buf, off)
    for _ in range(3): eng.run_staged(pkg.Decoder.TDT)
    eng.sync()
    eng.tdt_passes()
    eng.run_staged(pkg.Decoder.TDT)
    ps = eng.tdt_passes()
    print(name, 'passes', ps[4], 'cycles per pass: stage %.0f  mma %.0f  store+cluster-barrier %.0f  gather+fin %.0f' % tuple(ps[:4] / max(ps[4], 1)))
    a = eng.tdt_phases(); tot = a[:7].sum()
    print(name, 'steps', a[7], 'cycles/step', tot / a[7])
    for nm, v in zip(['P1', 'B1', 'P2', 'B2', 'P3', 'B3', 'P4'], a[:7]): print('   ', nm, f'{v / tot:.1%}', f'{v / a[7]:.0f} cycles/step')
    toks = eng.fetch(B); print('    tokens/utt', np.mean([len(t) for t in toks]))
    eng.close()
