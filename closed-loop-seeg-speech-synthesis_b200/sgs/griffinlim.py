"""Griffin-Lim operators on top of libsgs: node semantics (sgs_gl_node_synthesize) and the batch form."""
import numpy as np

from . import _lib
from .design import GriffinLimNodePlan, propagate

LP_CHUNK = 2048


def lowpass_transition(b, a):
    """Zero-input state transition of scipy.signal.lfilter's direct-form-II-transposed recurrence."""
    order = len(a) - 1
    M = np.zeros((order, order))
    for i in range(order):
        M[i, 0] = -a[i + 1]
        if i + 1 < order:
            M[i, i + 1] = 1.0
    return M


class GriffinLimNodeOp:
    """Batched equivalent of GriffinLimSynthesis fed frame by frame (livenodes/GriffinLim.py)."""

    def __init__(self, frame_size_ms=16, frame_shift_ms=10, sample_rate=16000, n_mels=40, iterations=5, cutoff=7900,
                 norm_factor=1.0, use_log_mels=True):
        self.use_log_mels = bool(use_log_mels)
        self.plan = GriffinLimNodePlan(frame_size_ms, frame_shift_ms, sample_rate, n_mels, iterations, 0, cutoff, norm_factor)
        p = self.plan
        self.first_frame = p.block_len - p.context_width - 1
        self.order = len(p.lp_a) - 1
        self.n_mels = n_mels
        self._handle = None

    def handle(self):
        if self._handle is None:
            _lib.ensure_init()
            p = self.plan
            a0 = p.lp_a[0]
            b = np.ascontiguousarray(p.lp_b / a0, dtype=np.float64)
            a = np.ascontiguousarray(p.lp_a / a0, dtype=np.float64)
            phi = np.ascontiguousarray(propagate(lowpass_transition(b, a), LP_CHUNK))
            phi_sub = np.ascontiguousarray(propagate(lowpass_transition(b, a), 64))
            win = np.ascontiguousarray(p.window, dtype=np.float64)
            ola = np.ascontiguousarray(p.ola_window, dtype=np.float64)
            idx = np.ascontiguousarray(p.mel.inv_idx, dtype=np.int32)
            w = np.ascontiguousarray(p.mel.inv_w, dtype=np.float64)
            h = _lib.c_void_p()
            _lib.check(_lib.lib().sgs_gl_node_create(
                _lib.C.byref(h), p.fft_size, p.hop, p.block_len, p.context_width, self.n_mels, _lib.ptr(win), _lib.ptr(ola),
                _lib.ptr(idx), _lib.ptr(w), _lib.ptr(b), _lib.ptr(a), self.order, _lib.ptr(phi), LP_CHUNK, _lib.ptr(phi_sub),
                float(p.norm_factor * 1.01), p.iterations))
            if not self.use_log_mels:
                _lib.check(_lib.lib().sgs_gl_node_set_log_mels(h, 0))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().sgs_gl_node_destroy(self._handle)
        except Exception:
            pass

    def positions(self, n_frames, start_ms=0.0):
        """Write-head position after each frame, with the node's float expression (GriffinLim.py:115-120)."""
        p = self.plan
        key = (int(n_frames), float(start_ms))
        cached = getattr(self, '_pos_cache', None)
        if cached is not None and cached[0] == key:
            return cached[1]
        # ms += frame_shift_ms per frame: np.cumsum adds sequentially in float64, i.e. the same roundings as the node's loop
        ms = np.cumsum(np.concatenate([[float(start_ms)], np.full(n_frames, float(p.frame_shift_ms))]))[1:]
        pos = ((ms / 1000.0) * p.sample_rate).astype(np.int64).astype(np.int32)
        pos.setflags(write=False)                              # shared between calls
        self._pos_cache = (key, pos)
        return pos

    def synthesize(self, logmel, noise=None, seed=0, want_filtered=False, want_blocks=False):
        """logmel (.., T, n_mels) float64 (numpy or torch-CUDA); noise (.., T, 480) or None.
        Returns int16 PCM (.., n_out) [, filtered float64] [, blocks (.., T, 480)]."""
        is_torch = _lib._is_torch(logmel)
        squeeze = logmel.ndim == 2
        if squeeze:
            logmel = logmel[None]
            noise = None if noise is None else noise[None]
        S, T, nm = logmel.shape
        assert nm == self.n_mels
        pos = self.positions(T)
        n_out = int(pos[-1] - pos[self.first_frame - 1]) if T > self.first_frame else 0
        if is_torch:
            import torch
            logmel = logmel.contiguous()
            assert logmel.dtype == torch.float64
            pcm = torch.empty((S, n_out), dtype=torch.int16, device=logmel.device)
            flt = torch.empty((S, n_out), dtype=torch.float64, device=logmel.device) if want_filtered else None
            blk = torch.empty((S, T, 480), dtype=torch.float64, device=logmel.device) if want_blocks else None
            if noise is not None:
                noise = noise.contiguous()
        else:
            logmel = np.ascontiguousarray(logmel, dtype=np.float64)
            pcm = np.empty((S, n_out), dtype=np.int16)
            flt = np.empty((S, n_out), dtype=np.float64) if want_filtered else None
            blk = np.empty((S, T, 480), dtype=np.float64) if want_blocks else None
            if noise is not None:
                noise = np.ascontiguousarray(noise, dtype=np.float64)
        if n_out > 0:
            _lib.check(_lib.lib().sgs_gl_node_synthesize(self.handle(), _lib.ptr(logmel), S, T, _lib.ptr(pos), _lib.ptr(noise),
                                                         int(seed), None, _lib.ptr(pcm), _lib.ptr(flt), _lib.ptr(blk),
                                                         _lib.current_stream(logmel)))
        out = [pcm[0] if squeeze else pcm]
        if want_filtered:
            out.append(flt[0] if squeeze else flt)
        if want_blocks:
            out.append(blk[0] if squeeze else blk)
        return out[0] if len(out) == 1 else tuple(out)


# ---------------------------------------------------------------------------------------------------------
# batch form (local/offline.py:131-192)
# ---------------------------------------------------------------------------------------------------------
_batch_plans = {}


def _batch_plan(win_len, hop, n_mels):
    from .design import MelTables
    key = (win_len, hop, n_mels)
    if key not in _batch_plans:
        _lib.ensure_init()
        mel = MelTables(int(win_len / 2 + 1), n_mels, 16000)
        window = np.ascontiguousarray(np.hanning(win_len + 1)[:-1], dtype=np.float64)     # offline.py:148
        idx = np.ascontiguousarray(mel.inv_idx, dtype=np.int32)
        w = np.ascontiguousarray(mel.inv_w, dtype=np.float64)
        h = _lib.c_void_p()
        _lib.check(_lib.lib().sgs_gl_batch_create(_lib.C.byref(h), win_len, hop, n_mels, _lib.ptr(window), _lib.ptr(idx), _lib.ptr(w)))
        _batch_plans[key] = h
    return _batch_plans[key]


def griffin_lim_batch(logmel, noise, win_length=0.05, hop_size=0.01, num_iterations=8, want_waveform=False):
    """logmel (B, T, n_mels), noise (B, >= 160*(T-1)+800) float64, numpy or torch-CUDA.  Returns int16 (B, 160*T)."""
    win_len = int(win_length * 16000)
    hop = int(win_len / (win_length / hop_size))
    is_torch = _lib._is_torch(logmel)
    B, T, nm = logmel.shape
    n_out = hop * T
    if is_torch:
        import torch
        logmel = logmel.contiguous(); noise = noise.contiguous()
        pcm = torch.empty((B, n_out), dtype=torch.int16, device=logmel.device)
        wave = torch.empty((B, n_out), dtype=torch.float64, device=logmel.device) if want_waveform else None
    else:
        logmel = np.ascontiguousarray(logmel, dtype=np.float64)
        noise = np.ascontiguousarray(noise, dtype=np.float64)
        pcm = np.empty((B, n_out), dtype=np.int16)
        wave = np.empty((B, n_out), dtype=np.float64) if want_waveform else None
    _lib.check(_lib.lib().sgs_gl_batch_synthesize(_batch_plan(win_len, hop, nm), _lib.ptr(logmel), B, T, _lib.ptr(noise),
                                                  int(noise.shape[1]), int(num_iterations), _lib.ptr(pcm), _lib.ptr(wave),
                                                  _lib.current_stream(logmel)))
    return (pcm, wave) if want_waveform else pcm
