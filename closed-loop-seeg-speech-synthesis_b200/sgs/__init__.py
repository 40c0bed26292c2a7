"""Host side of libsgs: ctypes binding (sgs._lib), once-per-configuration tables (sgs.design),
and the array-level operators the drop-in modules (livenodes/, local/, train.py, decode.py) call.

There is no CPU fallback anywhere in this package: every operator fails loudly when libsgs.so
or a CUDA device is missing.
"""
