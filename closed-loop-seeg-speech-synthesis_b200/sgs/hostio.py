"""Large host <-> device transfers of PAGEABLE numpy arrays (the arrays the reference's entry points take and return).

cudaMemcpy from pageable memory stages through a small driver buffer on one thread: 6.3 GB/s measured on the B200 box, which
made the upload of a 1 h recording (5.2 GB) the largest stage of train.train.  Here `workers` threads copy row chunks into
cached page-locked slots (numpy releases the GIL for large copies; strided sources such as a channel block eeg[:, c0:c1] are
gathered by the same copy) and queue one cudaMemcpyAsync per chunk on their own streams, so host copies and DMA overlap."""
import threading

import numpy as np

_CHUNK = 32 << 20
_slots = {}
_lock = threading.Lock()


def _get_slots(workers, chunk_bytes):
    import torch
    key = (workers, chunk_bytes, torch.cuda.current_device())
    with _lock:
        if key not in _slots:
            _slots[key] = [[torch.empty(chunk_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)] for _ in range(workers)]
        return _slots[key]


def _torch_dtype(np_dtype):
    import torch
    return torch.from_numpy(np.empty(0, dtype=np_dtype)).dtype


def _run(n_rows, rows_per_chunk, workers, body):
    import torch
    dev = torch.cuda.current_device()
    cur = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in range(workers)]
    for s in streams:
        s.wait_stream(cur)
    errors = []
    n_chunks = -(-n_rows // rows_per_chunk)

    def work(w):
        try:
            torch.cuda.set_device(dev)
            with torch.cuda.stream(streams[w]):
                events = [None, None]
                for k, i in enumerate(range(w, n_chunks, workers)):
                    r0 = i * rows_per_chunk
                    r1 = min(n_rows, r0 + rows_per_chunk)
                    if events[k & 1] is not None:
                        events[k & 1].synchronize()                 # the slot's previous transfer is complete
                    events[k & 1] = body(w, k & 1, r0, r1, streams[w])
                streams[w].synchronize()
        except Exception as e:                                       # surfaced by the caller
            errors.append(e)

    threads = [threading.Thread(target=work, args=(w,)) for w in range(workers)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]


def default_workers():
    """Copy threads per process.  4 threads move 9.8 GB/s of a pageable 1 h recording (the page-locked link takes 55 GB/s);
    12 threads per process measured SLOWER (5.2 GB in 1.14 s instead of 0.53 s on one rank, worse with two ranks on the box)."""
    import os
    env = os.environ.get('SGS_UPLOAD_WORKERS')
    return max(1, int(env)) if env else 4


def upload(a, dtype=None, workers=None, chunk_bytes=_CHUNK):
    """numpy array (any strides) -> torch CUDA tensor of the same shape on the current device, optionally cast to `dtype`."""
    import torch
    workers = workers or default_workers()
    a = np.asarray(a)
    out_dtype = a.dtype if dtype is None else np.dtype(dtype)
    dev = torch.device('cuda', torch.cuda.current_device())
    if a.nbytes < 4 * chunk_bytes or a.ndim == 0:
        return torch.from_numpy(np.ascontiguousarray(a, dtype=out_dtype)).to(dev)
    flat = a.ndim == 1
    if flat:                                                         # chunk a long vector as rows of 64 Ki elements
        width = 1 << 16
        body_rows = len(a) // width
        tail = a[body_rows * width:]
        a2 = a[:body_rows * width].reshape(body_rows, width)
    else:
        a2, tail = a, None
    row_bytes = int(np.prod(a2.shape[1:])) * out_dtype.itemsize
    rows_per_chunk = max(1, chunk_bytes // row_bytes)
    if rows_per_chunk * row_bytes > chunk_bytes:
        return torch.from_numpy(np.ascontiguousarray(a, dtype=out_dtype)).to(dev)
    out = torch.empty(a.shape, dtype=_torch_dtype(out_dtype), device=dev)
    out2 = out[:a2.size].view(a2.shape) if flat else out
    slots = _get_slots(workers, chunk_bytes)

    def body(w, b, r0, r1, stream):
        n = r1 - r0
        slot = slots[w][b]
        view = slot[:n * row_bytes].view(_torch_dtype(out_dtype)).view((n,) + tuple(a2.shape[1:]))
        np.copyto(view.numpy(), a2[r0:r1], casting='unsafe')
        out2[r0:r1].copy_(view, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        return ev

    _run(a2.shape[0], rows_per_chunk, workers, body)
    if flat and len(tail):
        out[a2.size:].copy_(torch.from_numpy(np.ascontiguousarray(tail, dtype=out_dtype)))
    return out


def download(t, workers=None, chunk_bytes=_CHUNK):
    """torch CUDA tensor -> fresh (pageable) numpy array."""
    import torch
    workers = workers or default_workers()
    if not t.is_cuda:
        return t.numpy()
    t = t.contiguous()
    if t.numel() * t.element_size() < 4 * chunk_bytes or t.ndim < 2:
        return t.cpu().numpy()
    row_bytes = int(np.prod(t.shape[1:])) * t.element_size()
    rows_per_chunk = max(1, chunk_bytes // row_bytes)
    if rows_per_chunk * row_bytes > chunk_bytes:
        return t.cpu().numpy()
    out = np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
    slots = _get_slots(workers, chunk_bytes)

    def body(w, b, r0, r1, stream):
        n = r1 - r0
        view = slots[w][b][:n * row_bytes].view(t.dtype).view((n,) + tuple(t.shape[1:]))
        view.copy_(t[r0:r1], non_blocking=True)
        stream.synchronize()
        np.copyto(out[r0:r1], view.numpy())
        return None

    _run(t.shape[0], rows_per_chunk, workers, body)
    return out
