"""Minimal HDF5 reader / writer for the flat artefact files of the reference - no h5py needed.

    params.h5   bad_channels, medians_array (40 x 9), estimators (np.void(pickle), an opaque scalar), select   train.py:192-196
    sEEG.hdf    sEEG (samples x channels), sEEG_sr (int32 scalar)                                            decode.py:201-203

Scope: what h5py writes for `create_dataset(name, data=array)` in the root group with default settings - superblock
version 0, version-1 object headers, the root group as symbol table (B-tree v1 + local heap + one or more symbol-table
nodes), little-endian fixed-point / IEEE float / opaque datatypes, contiguous (or compact) layout, no filters - restated from
the HDF5 File Format Specification, version 1.1/2.0 ("Disk Format: Level 0A-2A").  Anything else (chunked or compressed
datasets, nested groups, version-2 object headers, big-endian data) raises Hdf5LiteError instead of guessing.

PARITY UNPINNED: neither libhdf5 nor h5py exists in the build image, so the writer's files have only been read back by
this module's own reader and checked field by field against the specification (tests/test_host_logic.py); decode.py /
train.py use h5py whenever it is importable and come here only when it is not."""
import struct

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16
HEAP_FREE_NULL = 1


class Hdf5LiteError(ValueError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ---------------------------------------------------------------------------------------------------------------------
# datatype messages (message type 0x0003, version 1)
# ---------------------------------------------------------------------------------------------------------------------
def _encode_dtype(dt):
    dt = np.dtype(dt)
    if dt.kind in 'iu':
        bits0 = 0x08 if dt.kind == 'i' else 0x00                         # little-endian, zero padding, signed flag
        return struct.pack('<BBBBI', 0x10 | 0, bits0, 0, 0, dt.itemsize) + struct.pack('<HH', 0, 8 * dt.itemsize)
    if dt.kind == 'f' and dt.itemsize in (4, 8):
        sign, eloc, esize, msize, bias = (31, 23, 8, 23, 127) if dt.itemsize == 4 else (63, 52, 11, 52, 1023)
        head = struct.pack('<BBBBI', 0x10 | 1, 0x20, sign, 0, dt.itemsize)   # little-endian, mantissa normalisation 2 (implied msb)
        return head + struct.pack('<HHBBBBI', 0, 8 * dt.itemsize, eloc, esize, 0, msize, bias)
    if dt.kind == 'V' and dt.fields is None:
        return struct.pack('<BBBBI', 0x10 | 5, 0, 0, 0, dt.itemsize)        # opaque, empty tag
    raise Hdf5LiteError('unsupported dtype %r' % (dt,))


def _decode_dtype(buf):
    cv, b0, b1, b2, size = struct.unpack_from('<BBBBI', buf, 0)
    cls, version = cv & 0x0F, cv >> 4
    if version not in (1, 2, 3):
        raise Hdf5LiteError('datatype message version %d' % version)
    if cls == 0:
        if b0 & 1:
            raise Hdf5LiteError('big-endian integers are not supported')
        return np.dtype('<%s%d' % ('i' if b0 & 0x08 else 'u', size))
    if cls == 1:
        if b0 & 1 or b0 & 0x40:
            raise Hdf5LiteError('big-endian / VAX floats are not supported')
        if size not in (2, 4, 8):
            raise Hdf5LiteError('%d-byte floats are not supported' % size)
        return np.dtype('<f%d' % size)
    if cls == 5:
        return np.dtype('V%d' % size)
    raise Hdf5LiteError('datatype class %d is not supported (integers, floats and opaque only)' % cls)


# ---------------------------------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------------------------------
def _message(mtype, data, flags=0):
    body = data + b'\x00' * (_pad8(len(data)) - len(data))
    return struct.pack('<HHB3x', mtype, len(body), flags) + body


def _object_header(messages):
    body = b''.join(messages)
    return struct.pack('<BxHII4x', 1, len(messages), 1, len(body)) + body


def _dataset_header(arr, data_addr):
    shape = arr.shape
    space = struct.pack('<BBBx4x', 1, len(shape), 0) + b''.join(struct.pack('<Q', d) for d in shape)
    fill = struct.pack('<BBBB', 2, 2, 2, 0)                               # v2: allocate late, write fill if set, undefined value
    nbytes = arr.size * arr.dtype.itemsize
    layout = struct.pack('<BBQQ', 3, 1, data_addr if nbytes else UNDEF, nbytes)
    return _object_header([_message(0x0001, space), _message(0x0003, _encode_dtype(arr.dtype), flags=1),
                           _message(0x0005, fill, flags=1), _message(0x0008, layout)])


def write(path, datasets):
    """datasets: {name: array-like or np.void scalar}; written in the root group, contiguous, little-endian."""
    items = []
    for name, value in datasets.items():
        if isinstance(value, (bytes, bytearray)):
            value = np.void(bytes(value))
        arr = np.asarray(value)
        if arr.dtype.kind == 'V':
            arr = arr.reshape(())
        elif arr.dtype.kind == 'b':
            arr = arr.astype(np.int8)
        elif arr.dtype.kind not in 'iuf':
            raise Hdf5LiteError('dataset %r: unsupported dtype %r' % (name, arr.dtype))
        if arr.dtype.byteorder == '>':
            arr = arr.astype(arr.dtype.newbyteorder('<'))
        items.append((str(name), np.array(arr, order='C', copy=True)))      # (np.ascontiguousarray would turn scalars into 1-d arrays)
    items.sort(key=lambda kv: kv[0].encode())                             # symbol-table nodes hold their entries in name order
    if len(items) > 2 * LEAF_K * 2 * INTERNAL_K:
        raise Hdf5LiteError('too many datasets for a one-level group B-tree')

    # local heap data segment: offset 0 is the empty string (the B-tree's left-most key)
    heap, name_off = bytearray(8), {}
    for name, _ in items:
        name_off[name] = len(heap)
        raw = name.encode() + b'\x00'
        heap += raw + b'\x00' * (_pad8(len(raw)) - len(raw))
    heap = bytes(heap)

    nodes = [items[i:i + 2 * LEAF_K] for i in range(0, len(items), 2 * LEAF_K)] or [[]]
    snod_size = 8 + 2 * LEAF_K * 40
    btree_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    pos = 96
    root_header_addr = pos
    root_header_len = len(_object_header([_message(0x0011, struct.pack('<QQ', 0, 0))]))
    pos += root_header_len
    btree_addr = pos; pos += btree_size
    heap_addr = pos; pos += 32
    heap_data_addr = pos; pos += len(heap)
    snod_addr = []
    for _ in nodes:
        snod_addr.append(pos); pos += snod_size
    header_addr, data_addr = {}, {}
    for name, arr in items:
        header_addr[name] = pos
        pos += len(_dataset_header(arr, 0))
    for name, arr in items:
        pos = _pad8(pos)
        data_addr[name] = pos
        pos += arr.size * arr.dtype.itemsize
    eof = pos

    out = bytearray(eof)
    sb = SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
    sb += struct.pack('<QQII', 0, root_header_addr, 1, 0) + struct.pack('<QQ', btree_addr, heap_addr)
    assert len(sb) == 96
    out[0:96] = sb
    root = _object_header([_message(0x0011, struct.pack('<QQ', btree_addr, heap_addr))])
    out[root_header_addr:root_header_addr + len(root)] = root
    # B-tree node: type 0 (group), level 0; key[i + 1] = heap offset of the largest name in child i, key[0] = ""
    bt = b'TREE' + struct.pack('<BBHQQ', 0, 0, len(nodes) if items else 0, UNDEF, UNDEF) + struct.pack('<Q', 0)
    for node, addr in zip(nodes, snod_addr):
        if node:
            bt += struct.pack('<QQ', addr, name_off[node[-1][0]])
    out[btree_addr:btree_addr + len(bt)] = bt
    out[heap_addr:heap_addr + 32] = b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), HEAP_FREE_NULL, heap_data_addr)
    out[heap_data_addr:heap_data_addr + len(heap)] = heap
    for node, addr in zip(nodes, snod_addr):
        sn = b'SNOD' + struct.pack('<BxH', 1, len(node))
        for name, _ in node:
            sn += struct.pack('<QQII16x', name_off[name], header_addr[name], 0, 0)
        out[addr:addr + len(sn)] = sn
    for name, arr in items:
        h = _dataset_header(arr, data_addr[name])
        out[header_addr[name]:header_addr[name] + len(h)] = h
        raw = arr.tobytes()
        out[data_addr[name]:data_addr[name] + len(raw)] = raw
    with open(path, 'wb') as fh:
        fh.write(bytes(out))


# ---------------------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------------------
class _File:
    def __init__(self, buf):
        self.buf = buf
        self.base = 0
        for off in [0] + [512 << i for i in range(20)]:                  # the superblock may follow a user block
            if buf[off:off + 8] == SIGNATURE:
                self.sb = off
                break
            if off > len(buf):
                raise Hdf5LiteError('not an HDF5 file')
        else:
            raise Hdf5LiteError('not an HDF5 file')
        version = buf[self.sb + 8]
        if version not in (0, 1):
            raise Hdf5LiteError('superblock version %d: only the version-0/1 layout h5py writes by default is supported' % version)
        if buf[self.sb + 13] != 8 or buf[self.sb + 14] != 8:
            raise Hdf5LiteError('only 8-byte offsets and lengths are supported')
        p = self.sb + 24 + (4 if version == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from('<QQQQ', buf, p)
        entry = p + 32
        _, self.root_header, cache, _ = struct.unpack_from('<QQII', buf, entry)
        self.root_btree, self.root_heap = struct.unpack_from('<QQ', buf, entry + 24) if cache == 1 else (None, None)

    def at(self, addr):
        return self.base + addr

    def messages(self, addr):
        """(type, flags, data) of every message of a version-1 object header, continuation blocks included."""
        p = self.at(addr)
        if self.buf[p:p + 4] == b'OHDR':
            raise Hdf5LiteError('version-2 object headers (libver="latest") are not supported')
        version, n_msgs, _, size = struct.unpack_from('<BxHII', self.buf, p)
        if version != 1:
            raise Hdf5LiteError('object header version %d' % version)
        blocks, out = [(p + 16, size)], []
        while blocks and len(out) < n_msgs:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < n_msgs:
                mtype, msize, flags = struct.unpack_from('<HHB', self.buf, q)
                data = self.buf[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x0010:
                    caddr, clen = struct.unpack_from('<QQ', data, 0)
                    blocks.append((self.at(caddr), clen))
                out.append((mtype, flags, data))
        return out

    def heap_name(self, heap_addr, off):
        p = self.at(heap_addr)
        if self.buf[p:p + 4] != b'HEAP':
            raise Hdf5LiteError('bad local heap signature')
        _, _, data_addr = struct.unpack_from('<QQQ', self.buf, p + 8)
        q = self.at(data_addr) + off
        return self.buf[q:self.buf.index(b'\x00', q)].decode()

    def group_entries(self, btree_addr, heap_addr):
        """name -> object header address, walking the group's version-1 B-tree."""
        out = {}
        p = self.at(btree_addr)
        if self.buf[p:p + 4] == b'SNOD':
            n = struct.unpack_from('<H', self.buf, p + 6)[0]
            for i in range(n):
                name_off, header = struct.unpack_from('<QQ', self.buf, p + 8 + 40 * i)
                out[self.heap_name(heap_addr, name_off)] = header
            return out
        if self.buf[p:p + 4] != b'TREE':
            raise Hdf5LiteError('bad B-tree signature')
        ntype, _, used = struct.unpack_from('<BBH', self.buf, p + 4)
        if ntype != 0:
            raise Hdf5LiteError('unexpected B-tree node type %d in a group' % ntype)
        q = p + 24 + 8                                                   # past key[0]
        for _ in range(used):
            child = struct.unpack_from('<Q', self.buf, q)[0]
            out.update(self.group_entries(child, heap_addr))
            q += 16                                                      # child, then the next key
        return out

    def dataset(self, header_addr):
        shape = dtype = layout = None
        for mtype, _, data in self.messages(header_addr):
            if mtype == 0x0001:
                version, rank, flags = struct.unpack_from('<BBB', data, 0)
                off = 8 if version == 1 else 4
                if version not in (1, 2):
                    raise Hdf5LiteError('dataspace message version %d' % version)
                if version == 2 and data[3] == 2:
                    raise Hdf5LiteError('null dataspace')
                shape = struct.unpack_from('<%dQ' % rank, data, off)
            elif mtype == 0x0003:
                dtype = _decode_dtype(data)
            elif mtype == 0x0008:
                layout = data
            elif mtype == 0x000B:
                raise Hdf5LiteError('filtered (compressed) datasets are not supported')
        if shape is None or dtype is None or layout is None:
            raise Hdf5LiteError('object is not a simple dataset')
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        version, cls = layout[0], layout[1]
        if version == 3 and cls == 1:
            addr, size = struct.unpack_from('<QQ', layout, 2)
            raw = b'' if addr == UNDEF or n == 0 else self.buf[self.at(addr):self.at(addr) + n * dtype.itemsize]
        elif version == 3 and cls == 0:
            size = struct.unpack_from('<H', layout, 2)[0]
            raw = layout[4:4 + size]
        elif version in (1, 2) and layout[2] == 1:                       # old layout message: rank, class, 5 reserved, address
            rank = layout[1]
            addr = struct.unpack_from('<Q', layout, 8)[0]
            raw = self.buf[self.at(addr):self.at(addr) + n * dtype.itemsize]
        else:
            raise Hdf5LiteError('only contiguous and compact datasets are supported (layout version %d class %d)' % (version, cls))
        if len(raw) < n * dtype.itemsize:
            raw = bytes(raw) + b'\x00' * (n * dtype.itemsize - len(raw))  # never-written storage reads as the default fill value
        arr = np.frombuffer(bytes(raw), dtype=dtype, count=n).reshape(shape)
        if dtype.kind == 'V' and len(shape) == 0:
            return np.void(arr.tobytes())
        return arr.copy() if len(shape) else arr.reshape(())[()]


def read(path, names=None):
    """{name: value} for the datasets of the root group (all of them, or `names`)."""
    with open(path, 'rb') as fh:
        f = _File(fh.read())
    btree, heap = f.root_btree, f.root_heap
    if btree is None:
        for mtype, _, data in f.messages(f.root_header):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from('<QQ', data, 0)
        if btree is None:
            raise Hdf5LiteError('root group is not a symbol-table group (file written with libver="latest"?)')
    entries = f.group_entries(btree, heap)
    want = list(entries) if names is None else list(names)
    out = {}
    for name in want:
        if name not in entries:
            raise KeyError("Unable to open object (object '%s' doesn't exist)" % name)
        out[name] = f.dataset(entries[name])
    return out
