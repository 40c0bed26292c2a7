import os, sys, numpy as np
sys.path.insert(0, 'closed-loop-seeg-speech-synthesis_b200')
from sgs import _lib
from sgs.griffinlim import GriffinLimNodeOp
rng = np.random.default_rng(0)
T = 60
spec = rng.normal(-2, 1.5, size=(T, 40))
noise = rng.random((T, 480))
def run(groups):
    op = GriffinLimNodeOp(16, 10, 16000, 40, 8, 7900, 10)
    pos_all = op.positions(T)
    out = []; k = 0; prev = 0
    pcm = np.empty(16*192, np.int16)
    for n in groups:
        n_pcm = _lib.c_int(0)
        fr = np.ascontiguousarray(spec[k:k+n]); nz = np.ascontiguousarray(noise[k:k+n]); ps = np.ascontiguousarray(pos_all[k:k+n])
        _lib.check(_lib.lib().sgs_gl_node_push(op.handle(), _lib.ptr(fr), n, _lib.ptr(ps), int(prev), _lib.ptr(nz), 0, _lib.ptr(pcm), _lib.C.byref(n_pcm), None))
        out.append(pcm[:n_pcm.value].copy()); prev = int(pos_all[k+n-1]); k += n
    return np.hstack(out)
a = run([1]*T)
for groups in ([3]*20, [4]*15, [3,3,3,3,3,3,3,4]*2+[3]*3+[1], [2]*30):
    b = run(groups)
    d = np.nonzero(a.astype(int) != b.astype(int))[0]
    print(groups[:8], len(a), len(b), 'mismatch samples', len(d), d[:5]//160 if len(d) else '')
