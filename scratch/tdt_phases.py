import sys, os
sys.path.insert(0,'/root/repo')
import numpy as np, __graft_entry__ as ge
import bench
os.makedirs('/tmp/pk_bench', exist_ok=True); pkg, synth, cfg, wp = bench.make_checkpoint('/tmp/pk_bench')
eng = pkg.Engine(cfg, wp, 0)
pcms=[synth.make_audio(160000,1000+i) for i in range(64)]
buf=np.concatenate(pcms); off=np.arange(65,dtype=np.int64)*160000
eng.stage(buf,off)
for _ in range(3): eng.run_staged(pkg.Decoder.TDT)
eng.sync()
a=eng.tdt_phases(); tot=a[:7].sum()
print('steps',a[7],'cycles',tot,'per step',tot/a[7], 'us/step@1.9GHz', tot/a[7]/1900)
for n,v in zip(['P1','B1','P2','B2','P3','B3','P4'],a[:7]): print(n, v, f'{v/tot:.1%}', f'{v/a[7]/1900:.2f} us/step')
toks=eng.fetch(64); print('tokens/utt', np.mean([len(t) for t in toks]))
