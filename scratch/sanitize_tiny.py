"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): tiny model, CTC + TDT (cluster kernel, grid barriers,
DSMEM), boosted decode, non-16 kHz input, two streaming steps.  `python scratch/sanitize_tiny.py [tdt|notdt]`."""
import os, sys, tempfile
sys.path.insert(0, '/root/repo')
os.environ.setdefault('PK_GRAPH', '0')
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
from parakeet_cpp_b200 import synth
O = ge.load_oracle()
with_tdt = not (len(sys.argv) > 1 and sys.argv[1] == 'notdt')
cfg, ocfg = pkg.make_tiny_config(), O.make_tiny_config()
td = tempfile.mkdtemp()
wp = os.path.join(td, 'tiny.safetensors')
synth.save_safetensors(wp, synth.make_weights(ocfg, seed=3))
e = pkg.Engine(cfg, wp, 0)
pcms = [synth.make_audio(24000, 11), synth.make_audio(9000, 12), synth.make_audio(400, 13)]
print('ctc', [len(t) for t in e.transcribe_batch(pcms, pkg.Decoder.CTC)])
if with_tdt:
    print('tdt', [len(t) for t in e.transcribe_batch(pcms, pkg.Decoder.TDT)])
e.set_boost([[3, 4], [7]], 4.0)
print('boost ctc', [len(t) for t in e.transcribe_batch(pcms[:2], pkg.Decoder.CTC)])
if with_tdt:
    print('boost tdt', [len(t) for t in e.transcribe_batch(pcms[:2], pkg.Decoder.TDT)])
e.set_boost([], 0.0)
print('rate', [len(t) for t in e.transcribe_batch_rate(pcms[:2], 22050, pkg.Decoder.CTC)])
e.close()
if with_tdt:
    scfg, socfg = pkg.make_tiny_stream_config(), O.make_tiny_stream_config()
    wps = os.path.join(td, 'ts.safetensors')
    synth.save_safetensors(wps, synth.make_weights(socfg, seed=3))
    es = pkg.Engine(scfg, wps, 0)
    es.stream_open(3, 2560)
    x = synth.make_audio(2560 * 4, 77)
    for k in range(4):
        print('stream', [len(t) for t in es.stream_step([x[k * 2560:(k + 1) * 2560], x[k * 2560:k * 2560 + 1000], np.zeros(0, np.float32)])])
    es.close()
print('done')
