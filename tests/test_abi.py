"""The C ABI: every function declared in include/sgs.h is exported by the built library, loads without a GPU,
and the product path fails loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'sgs.h')
LIB = os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200', 'csrc', 'libsgs.so')


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sgs_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_hot_path_entry_points():
    names = declared_functions()
    for must in ('sgs_feat_extract', 'sgs_feat_stack', 'sgs_feat_stream_push', 'sgs_feat_stream_set_cold_start', 'sgs_lda_decode', 'sgs_dequantize',
                 'sgs_gl_node_synthesize', 'sgs_gl_node_push', 'sgs_gl_node_rebase', 'sgs_gl_node_set_log_mels', 'sgs_gl_batch_synthesize', 'sgs_logmel', 'sgs_quantize',
                 'sgs_spearman', 'sgs_lda_stats', 'sgs_last_error', 'sgs_init'):
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        import subprocess
        subprocess.run(['make', '-C', os.path.dirname(LIB), '-j8'], check=True)
    lib = ctypes.CDLL(LIB)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    lib.sgs_abi_version.restype = ctypes.c_int
    assert lib.sgs_abi_version() >= 1


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present; the failure path is exercised on the CPU-only builder")
    import numpy as np
    from sgs import _lib
    from local.offline import herff2016_b
    with pytest.raises(_lib.SgsError, match="no CUDA device|CUDA error"):
        herff2016_b(np.zeros((2048, 4)), 1024)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'closed-loop-seeg-speech-synthesis_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(import|from)\s+oracle\b', src, flags=re.M), os.path.join(base, f)
                assert 'ref_shim' not in src and '/root/reference' not in src
