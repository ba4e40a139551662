import sys, os, tempfile
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np, __graft_entry__ as ge, oracle as O
pkg=ge.load_package()
from parakeet_cpp_b200 import synth
g=np.load('/root/repo/tests/golden/golden_v1.npz')
ocfg=O.make_110m_config(); W=synth.make_weights(ocfg,seed=0)
td=tempfile.mkdtemp(); wp=td+'/w.safetensors'; synth.save_safetensors(wp,W)
eng=pkg.Engine(pkg.make_110m_config(max_batch=4), wp, 0)
pcm=synth.make_audio(160000,1000)
fo=O.preprocess_audio(pcm); fg=eng.mel([pcm])[0]
err=np.abs(fo-fg).max(1); print('mel per-frame err: max', err.max(), 'argmax', err.argmax(), 'last10', err[-10:].round(5), 'first5', err[:5].round(5))
encs,subs,lays=eng.encode([fo],taps=True)
ge_=g['m110.c0.enc']; print('enc(oracle mel) rel', np.abs(encs[0]-ge_).max()/np.abs(ge_).max(), 'rows', np.abs(encs[0]-ge_).max(1)[110:].round(5))
encs2=eng.encode([fg]); print('enc(gpu mel) rel', np.abs(encs2[0]-ge_).max()/np.abs(ge_).max(), 'rows', np.abs(encs2[0]-ge_).max(1)[100:].round(4))
print('sub rel', np.abs(subs[0]-g['m110.c0.sub']).max()/np.abs(g['m110.c0.sub']).max())
