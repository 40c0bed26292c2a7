"""Pinned host-to-device bandwidth of the box, one process per GPU: what bounds bench.py's end-to-end leg.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tools/h2d_probe.py

For every rank: the GPU's PCI address and NUMA node, the CPUs / memory nodes this process may use, then pinned H2D copies
(one cudaMemcpyAsync of 1 GiB per repetition, CUDA-event timed)
  alone        one GPU copying at a time, buffer placed by the default policy
  all_default  all N GPUs copying at once, buffers placed by the default (first-touch) policy
  all_node<k>  all at once, every buffer bound (set_mempolicy MPOL_BIND before the allocation) to NUMA node k, for every
               online node: tells which node is local to which GPU even where sysfs reports numa_node = -1
  all_best     all at once, every buffer on the node that gave its GPU the highest rate in the all_node<k> passes
Rank 0 prints one JSON object.  The aggregate of `all_*` is the ceiling of the 8-GPU end-to-end number."""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

GIB = 1 << 30
MPOL_DEFAULT, MPOL_BIND = 0, 2
SYS_set_mempolicy = 238            # x86_64


def set_mempolicy(node):
    """Bind this thread's future page allocations to one NUMA node (None = default policy).  Returns 0 on success."""
    libc = ctypes.CDLL(None, use_errno=True)
    if node is None or node < 0:
        return libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(SYS_set_mempolicy, MPOL_BIND, ctypes.byref(mask), ctypes.c_ulong(64))


def read(path, default=None):
    try:
        return open(path).read().strip()
    except OSError:
        return default


def gpu_numa_node(index):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    addr = bus.lower()
    if len(addr.split(':')[0]) == 8:
        addr = addr[4:]
    node = read('/sys/bus/pci/devices/%s/numa_node' % addr)
    return addr, (int(node) if node not in (None, '') else None)


def pinned(nbytes, node):
    rc = set_mempolicy(node)
    try:
        buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        buf.fill_(1)                                   # touch every page under the policy in force
    finally:
        set_mempolicy(None)
    return buf, rc


def copy_rate(dst, src, reps=4):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    return reps * src.numel() / (a.elapsed_time(b) * 1e-3) / 1e9


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    addr, node = gpu_numa_node(local)
    status = read('/proc/self/status', '')
    info = {"rank": rank, "pci": addr, "gpu_numa_node": node, "cpus_allowed": len(os.sched_getaffinity(0)),
            "mems_allowed_list": next((l.split(':')[1].strip() for l in status.splitlines() if l.startswith('Mems_allowed_list')), None),
            "numa_nodes_online": read('/sys/devices/system/node/online')}
    dst = torch.empty(GIB, dtype=torch.uint8, device='cuda')

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def gather(v):
        if world == 1:
            return [v]
        out = [None] * world
        dist.all_gather_object(out, v)
        return out

    res = {}
    buf, _ = pinned(GIB, None)
    copy_rate(dst, buf, 1)
    alone = 0.0
    for r in range(world):
        barrier()
        if r == rank:
            alone = copy_rate(dst, buf)
        barrier()
    res['alone'] = gather(round(alone, 2))
    def online_nodes():
        out = []
        for part in (read('/sys/devices/system/node/online') or '0').split(','):
            lo, _, hi = part.partition('-')
            out += list(range(int(lo), int(hi or lo) + 1))
        return out

    def concurrent(name, want):
        nonlocal buf
        if want is not None:
            del buf
            buf, rc = pinned(GIB, want)
            info['set_mempolicy_rc_' + name] = rc
        copy_rate(dst, buf, 1)
        barrier()
        rate = copy_rate(dst, buf)
        barrier()
        per = gather(round(rate, 2))
        res[name] = {"per_gpu_gbs": per, "aggregate_gbs": round(sum(per), 1)}
        return rate

    concurrent('all_default', None)
    by_node = {k: concurrent('all_node%d' % k, k) for k in online_nodes()}
    best = max(by_node, key=by_node.get)
    info['best_node'] = best
    concurrent('all_best', best)
    infos = gather(info)
    if rank == 0:
        print(json.dumps({"gpus": world, "bytes_per_copy": GIB, "ranks": infos, "h2d_gbs": res}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    sys.exit(main())
