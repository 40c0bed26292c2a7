"""Import shim that makes the UNMODIFIED reference tree at /root/reference importable in
this container, for golden-vector generation only (oracle/gen_golden.py).

TEST INFRASTRUCTURE - never imported by the product path, by `-m gpu` tests, by
`bench.py` or by `smoke()`: /root/reference does not exist on the GPU box.

What it does:
  * injects stub modules for the absent third-party imports (mne, h5py, matplotlib, pylsl,
    fastdtw, jack, samplerate, pyaudio). `mne.filter.create_filter` is the one stub that
    carries arithmetic: mne 0.19 with iir_params={'order': 8, 'ftype': 'butter'} designs
    scipy.signal.iirfilter(8, [f1, f2]/(sr/2), band, 'butter', output='sos')
    (called from ECogFeatCalc.py:131-133 and local/offline.py:26-28). PARITY UNPINNED for
    that one call: real mne 0.19 is not installable here.
  * restores numpy/scipy aliases removed since the reference's pinned versions
    (np.float, scipy.hanning/blackman/zeros/real).
  * loads livenodes/GriffinLim.py with line 104 (`if self.inputBuffer == []`) patched in
    memory to an isinstance check: the original raises under numpy >= 1.25 once the buffer
    is an ndarray. No reference source is copied into this repo.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import scipy
import scipy.signal

REFERENCE_ROOT = os.environ.get("SGS_REFERENCE_ROOT", "/root/reference")


def _mod(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


def _check_method(method, iir_params, extra_types=()):
    return iir_params, method


def _create_filter(data, sfreq, l_freq, h_freq, filter_length='auto', l_trans_bandwidth='auto',
                   h_trans_bandwidth='auto', method='fir', iir_params=None, phase='zero',
                   fir_window='hamming', fir_design='firwin'):
    nyq = sfreq / 2.0
    if l_freq < h_freq:
        btype, wn = 'bandpass', [l_freq / nyq, h_freq / nyq]
    else:
        btype, wn = 'bandstop', [h_freq / nyq, l_freq / nyq]
    sos = scipy.signal.iirfilter(iir_params['order'], wn, btype=btype, ftype=iir_params['ftype'], output='sos')
    return dict(iir_params, sos=sos)


def install():
    """Make `import train, decode; from livenodes import ...; from local.offline import ...` work."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not present at %s (expected on the GPU box)" % REFERENCE_ROOT)
    np.float = float
    np.object = object
    np.int = int
    scipy.hanning, scipy.blackman = np.hanning, np.blackman
    scipy.zeros, scipy.real = np.zeros, np.real
    flt = _mod('mne.filter', _check_method=_check_method, create_filter=_create_filter)
    _mod('mne', filter=flt, set_log_level=lambda *a, **k: None)
    for n in ('h5py', 'matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.axes_grid1', 'jack',
              'samplerate', 'fastdtw', 'pyaudio'):
        _mod(n)
    sys.modules['mpl_toolkits.axes_grid1'].make_axes_locatable = None
    sys.modules['fastdtw'].fastdtw = None
    sys.modules['jack'].JackError = Exception
    _mod('pylsl', StreamInlet=None, resolve_stream=None, FOREVER=None, StreamInfo=None, StreamOutlet=None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _load_patched_griffinlim()


def _load_patched_griffinlim():
    import livenodes  # namespace package rooted at REFERENCE_ROOT
    path = os.path.join(REFERENCE_ROOT, 'livenodes', 'GriffinLim.py')
    src = open(path).read()
    bad = "if self.inputBuffer == []:"
    assert src.count(bad) == 1, "reference GriffinLim.py changed"
    src = src.replace(bad, "if isinstance(self.inputBuffer, list):")
    spec = importlib.util.spec_from_loader('livenodes.GriffinLim', loader=None, origin=path)
    mod = importlib.util.module_from_spec(spec)
    mod.__file__ = path
    mod.__package__ = 'livenodes'
    sys.modules['livenodes.GriffinLim'] = mod
    exec(compile(src, path, 'exec'), mod.__dict__)
    livenodes.GriffinLim = mod
