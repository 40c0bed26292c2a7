import os, sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, 'closed-loop-seeg-speech-synthesis_b200'); sys.path.insert(0,'oracle')
import test_gpu_nodes as T
from helpers import load, GOLDEN
G = load('train_decode.npz'); blob = open(os.path.join(GOLDEN, 'estimators.pkl'), 'rb').read()
a = T._run_graph((G, blob), True, 64, chunk_size=64)
b = T._run_graph((G, blob), False, 64, chunk_size=64)
bad = [i for i,(p,q) in enumerate(zip(a[3], b[3])) if not np.array_equal(p,q)]
print(len(a[3]), 'bad frames', bad[:40], len(bad))
for i in bad[:5]:
    d = a[3][i].astype(int)-b[3][i].astype(int)
    print(i, len(a[3][i]), len(b[3][i]), np.nonzero(d)[0][:10], d[np.nonzero(d)[0][:10]])
import oracle as O
spec = a[2]
rs = np.random.RandomState(4001)
noise = np.zeros((len(spec), 480))
for k in range(1, len(spec)): noise[k] = rs.rand(480)
gl = O.GriffinLimNode(16, 10, 16000, 40, 8, norm_factor=10)
pcm, _ = gl.synthesize(spec, noise)
fa = np.hstack(a[3]); fb = np.hstack(b[3])
for name, f in (('fused', fa), ('unfused', fb)):
    d = np.abs(f.astype(int) - pcm.astype(int))
    print(name, len(f), len(pcm), 'max', d.max(), 'frames with |d|>1:', sorted(set((np.nonzero(d > 1)[0] // 160).tolist()))[:20])
