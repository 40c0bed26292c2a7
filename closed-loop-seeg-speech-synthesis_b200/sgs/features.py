"""Array-level feature extraction operator on top of libsgs (sgs_feat_extract / sgs_feat_stack).

Inputs may be numpy arrays (host; copied inside the call) or torch CUDA tensors (resident; outputs are
then CUDA tensors too).  Shapes follow the reference: sEEG is (samples x channels), optionally with a
leading session axis.
"""
import os

import numpy as np

from . import _lib
from .design import FeaturePlan, propagate

SM_COUNT = 148
STREAMS_PER_BLOCK = 32          # csrc/feat.cu: one CTA = 4 stage-warps x 32 streams
BLOCKS_PER_SM = 3               # co-resident CTAs needed to cover the hand-off latency (measured optimum)
MIN_CHUNK = 2048                # amortises the window overlap (<= 102 samples) and the carry step
# What a state may still contribute after the zero-state warm-up of a time piece, relative to the state: 2^-50 is below
# fp64's own resolution of the state (2^-52 per operation, summed over the recurrence) and gives a 36 864-sample horizon at
# 2048 Hz; the 2^-70 of round 1 cost 49 152 samples - a quarter more warm-up work (6.1 -> 4.6 ms per step) for digits
# that the recurrence's rounding has already lost.  SGS_FEAT_TOL_LOG2 overrides the exponent.
import os as _os
DEFAULT_TOL = 2.0 ** float(_os.environ.get('SGS_FEAT_TOL_LOG2', '-50'))


class FeatureExtractor:
    """One feature configuration bound to a device plan.  Create lazily (after fork) - see _lib.ensure_init."""

    def __init__(self, sr, window_length=None, window_shift=None, line_noise=50, model_order=4, step_size=5,
                 frame_len_ms=None, frame_shift_ms=None, carry_tol=DEFAULT_TOL):
        self.plan = FeaturePlan(sr, window_length, window_shift, line_noise, model_order, step_size,
                                frame_len_ms, frame_shift_ms)
        self.carry_tol = carry_tol
        self._handle = None
        self._A = None
        self._horizon = None
        self._phi = {}
        self._tables = {}
        self._tail = None
        self._tail_set = False

    # -- device plan --------------------------------------------------------------------------
    def handle(self):
        if self._handle is None:
            _lib.ensure_init()
            p = self.plan
            h = _lib.c_void_p()
            coef = _lib.host(p.coef, np.float64)
            zi = _lib.host(p.zi_unit, np.float64)
            ziw = _lib.host(p.zi_last_warm, np.float64)
            zf = _lib.host(p.zero_fill_response, np.float64)
            _lib.check(_lib.lib().sgs_feat_plan_create(_lib.C.byref(h), p.n_filters, _lib.ptr(coef), _lib.ptr(zi),
                                                       _lib.ptr(ziw), _lib.ptr(zf), p.zero_fill))
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().sgs_feat_plan_destroy(self._handle)
        except Exception:
            pass

    # -- scan planning (host, O(1) in the recording length) -----------------------------------
    def transition(self):
        """48x48 (or 32x32) one-sample state transition matrix of the whole cascade (zero input)."""
        if self._A is None:
            p = self.plan
            ns = 2 * p.n_biquads
            A = np.zeros((ns, ns))
            for j in range(ns):
                s = np.zeros(ns)
                s[j] = 1.0
                v = 0.0
                for i in range(p.n_biquads):
                    b0, b1, b2, a1, a2 = p.coef[i]
                    o = b0 * v + s[2 * i]
                    s[2 * i] = b1 * v - a1 * o + s[2 * i + 1]
                    s[2 * i + 1] = b2 * v - a2 * o
                    v = o
                A[:, j] = s
            self._A = A
        return self._A

    def horizon(self):
        """Smallest multiple of 1024 samples after which the influence of a state is below carry_tol."""
        if self._horizon is None:
            P = propagate(self.transition(), 1024)
            M, w = P.copy(), 1024
            while np.abs(M).max() > self.carry_tol and w < (1 << 22):
                M = M @ P
                w += 1024
            self._horizon = w
        return self._horizon

    def tail(self):
        """Modal tail of the warm-up (sgs/modal.py), or None when switched off (SGS_FEAT_TAIL=0)."""
        if self._tail is None and os.environ.get('SGS_FEAT_TAIL', '1') != '0':
            from .modal import modal_tail
            try:
                self._tail = modal_tail(self.plan.coef, self.carry_tol)
            except ImportError:                   # no mpmath: the device keeps the zero-state warm-up over the whole horizon
                self._tail = None
        return self._tail

    def _ensure_tail(self):
        """Hands the tail tables to the device plan once (only plans that cut time into pieces need them)."""
        if self._tail_set:
            return
        self._tail_set = True
        t = self.tail()
        if t is None or t.n_modes == 0:
            return
        lam = np.ascontiguousarray(np.stack([t.lam.real, t.lam.imag], axis=1))
        shift = np.ascontiguousarray(np.stack([t.warp_shift.real, t.warp_shift.imag], axis=2))
        mode_len = np.ascontiguousarray(t.mode_len, dtype=np.int32)
        kappa = np.ascontiguousarray(np.stack([t.kappa.real, t.kappa.imag], axis=1))
        _lib.check(_lib.lib().sgs_feat_plan_set_tail(self.handle(), t.near_len, t.n_modes, _lib.ptr(mode_len), _lib.ptr(lam),
                                                     _lib.ptr(t.warp_blocks), _lib.ptr(shift), _lib.ptr(t.state_matrix),
                                                     _lib.ptr(kappa)))

    def pieces_with_tail(self, n_samples, n_streams):
        """Whether csrc/api_feat.cu cuts this job into 4 x SM-count balanced pieces on the strength of the modal tail alone."""
        if os.environ.get('SGS_FEAT_PIECES', '1') == '0' or os.environ.get('SGS_FEAT_PIECES_P'):
            return False
        t = self.tail()
        if t is None or t.n_modes == 0 or self.horizon() % 64:
            return False
        groups = -(-n_streams // STREAMS_PER_BLOCK)
        piece_len = -(-(-(-groups * n_samples // (4 * SM_COUNT))) // 64) * 64
        min_len = int(os.environ.get('SGS_FEAT_PIECES_MINLEN', 0)) or 4 * t.near_len
        return min_len <= piece_len <= n_samples and n_samples >= 128

    def phi(self, chunk_len):
        if chunk_len not in self._phi:
            self._phi[chunk_len] = np.ascontiguousarray(propagate(self.transition(), int(chunk_len)))
        return self._phi[chunk_len]

    def scan_plan(self, n_samples, n_streams, chunks=None, horizon=None):
        """(n_chunks, chunk_len, horizon, phi-or-None)."""
        chunks = chunks if chunks is not None else os.environ.get('SGS_FEAT_CHUNKS')
        explicit = chunks
        if chunks is None and horizon is None and self.pieces_with_tail(n_samples, n_streams):
            # balanced pieces (csrc/api_feat.cu takes them for any 2-chunk plan without phi when the pieces are long enough for
            # the modal tail): the same condition as there, so that the (group x chunk) grid never sees this 2-chunk plan
            return 2, -(-(-(-n_samples // 2)) // 64) * 64, self.horizon(), None
        if chunks is None:
            groups = -(-n_streams // STREAMS_PER_BLOCK)
            want = max(1, int(round(SM_COUNT * BLOCKS_PER_SM / groups)))
            chunks = max(1, min(want, n_samples // MIN_CHUNK))
        chunks = int(chunks)
        if chunks <= 1:
            return 1, int(n_samples), 0, None
        chunk_len = -(-n_samples // chunks)
        chunk_len = -(-chunk_len // 64) * 64
        chunks = -(-n_samples // chunk_len)
        if chunks <= 1:
            return 1, int(n_samples), 0, None
        w = int(horizon) if horizon is not None else self.horizon()
        if w >= chunk_len:
            if explicit is None:
                # few streams, long recording (a channel block of a 1 h session): chunks of at least one horizon keep the
                # truncated zero-state pass (work <= 2 T, no carry) where the short chunks above would need the exact carry -
                # a serial walk over hundreds of chunks with one thread per stream (0.5 s for 64 streams x 7.4 M samples)
                groups = -(-n_streams // STREAMS_PER_BLOCK)
                long_len = max(-(-w // 64) * 64 + 64, -(-(-(-n_samples * groups // (2 * SM_COUNT))) // 64) * 64)
                long_chunks = -(-n_samples // long_len)
                if long_chunks >= 2 and groups * long_chunks >= SM_COUNT:
                    return long_chunks, long_len, w, None
            return chunks, chunk_len, chunk_len, (self.phi(chunk_len) if chunks > 2 else None)
        return chunks, chunk_len, w, None

    # -- window tables (cached per recording length) ------------------------------------------
    def windows(self, n_samples, online, chunk_size=32):
        key = (int(n_samples), bool(online), int(chunk_size))
        if key not in self._tables:
            if online:
                usable = (n_samples // chunk_size) * chunk_size
                starts, wl = self.plan.online_window_starts(usable)
            else:
                starts, wl = self.plan.offline_window_starts(n_samples)
            self._tables[key] = (np.ascontiguousarray(starts, dtype=np.int32), int(wl))
            if len(self._tables) > 16:
                self._tables.pop(next(iter(self._tables)))
        return self._tables[key]

    # -- operators ----------------------------------------------------------------------------
    def log_power(self, x, online=False, chunk_size=32, chunks=None, horizon=None):
        """x: (T, C) or (S, T, C), float32/float64, numpy or torch-CUDA.  Returns (.., W, C) float64 log-power."""
        is_torch = _lib._is_torch(x)
        squeeze = x.ndim == 2
        if squeeze:
            x = x[None]
        if is_torch:
            import torch
            if x.dtype not in (torch.float32, torch.float64):
                x = x.to(torch.float32)
            x = x.contiguous()
            is64 = x.dtype == torch.float64
        else:
            if x.dtype not in (np.float32, np.float64):
                x = x.astype(np.float64)
            x = np.ascontiguousarray(x)
            is64 = x.dtype == np.float64
        S, T, Cn = x.shape
        starts, wl = self.windows(T, online, chunk_size)
        nw = len(starts)
        if is_torch:
            import torch
            out = torch.empty((S, nw, Cn), dtype=torch.float64, device=x.device)
        else:
            out = np.empty((S, nw, Cn), dtype=np.float64)
        if nw > 0 and T > 0:
            k, clen, w, phi = self.scan_plan(T, S * Cn, chunks, horizon)
            if k > 1 and phi is None:
                self._ensure_tail()
            _lib.check(_lib.lib().sgs_feat_extract(self.handle(), _lib.ptr(x), int(is64), T, Cn, S, 0, _lib.ptr(starts), nw,
                                                   wl, k, clen, w, _lib.ptr(phi), _lib.ptr(out), _lib.current_stream(x)))
        return out[0] if squeeze else out

    def stack(self, feat, online=False):
        """(.., W, C) -> (.., rows, C*(order+1)) in the reference's c*5+tap column order."""
        is_torch = _lib._is_torch(feat)
        squeeze = feat.ndim == 2
        if squeeze:
            feat = feat[None]
        S, nw, Cn = feat.shape
        order, step = self.plan.model_order, self.plan.step_size
        first = 0 if online else order * step
        rows = max(0, nw - first)
        if is_torch:
            import torch
            out = torch.empty((S, rows, Cn * (order + 1)), dtype=torch.float64, device=feat.device)
            feat = feat.contiguous()
        else:
            out = np.empty((S, rows, Cn * (order + 1)), dtype=np.float64)
            feat = np.ascontiguousarray(feat, dtype=np.float64)
        if rows > 0:
            _lib.ensure_init()
            _lib.check(_lib.lib().sgs_feat_stack(_lib.ptr(feat), S, nw, Cn, rows, first, order, step, _lib.ptr(out),
                                                 _lib.current_stream(feat)))
        return out[0] if squeeze else out
