import sys, os, tempfile
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np, __graft_entry__ as ge, oracle as O
pkg=ge.load_package()
from parakeet_cpp_b200 import synth
ocfg=O.make_110m_config(); W=synth.make_weights(ocfg,seed=0)
td=tempfile.mkdtemp(); wp=td+'/w.safetensors'; synth.save_safetensors(wp,W)
pcm=synth.make_audio(160000,1000); fo=O.preprocess_audio(pcm)
x=O.conv_subsampling(W,fo,ocfg); pos=O.sinusoidal_position_embedding(x.shape[0],x.shape[1])
p='encoder_.layers_.0.'
steps=[]
x1=O.feed_forward(W,p+'ffn1_.',x); steps.append(x1)
x2=O.conformer_attention(W,p+'attn_.',x1,pos,ocfg); steps.append(x2)
x3=O.conformer_conv(W,p+'conv_.',x2,ocfg); steps.append(x3)
x4=O.feed_forward(W,p+'ffn2_.',x3); steps.append(x4)
rel=lambda a,b: float(np.abs(a-b).max()/np.abs(b).max())
for n,want in enumerate(steps,1):
    os.environ['PK_DEBUG_SUBBLOCKS']=str(n)
    eng=pkg.Engine(pkg.make_110m_config(max_batch=2), wp, 0)
    got=eng.encode([fo])[0]; eng.close()
    print('after sub-block',n,'rel',rel(got,want), 'rows err', np.abs(got-want).max(1)[[0,1,62,63,64,65,124,125]].round(5))
os.environ.pop('PK_DEBUG_SUBBLOCKS')
eng=pkg.Engine(pkg.make_110m_config(max_batch=2), wp, 0)
encs,subs,lays=eng.encode([fo],taps=True)
eo,so,lo=O.encoder_forward(W,fo,ocfg,return_layers=True)
print('per-layer rel', [f'{rel(lays[0][i],lo[i]):.1e}' for i in range(17)])
