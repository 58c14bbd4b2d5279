"""Minibatches in the dataset's own byte format (examples/mnist.jl:77-96).

The reference loads the idx-ubyte files on the host and keeps `xtrn = reshape(a ./ 255f0, 784, :)` (Float32,
4 bytes per pixel) and a dense one-hot `ytrn` (10 x N Float32); `minibatch` slices columns out of those.  Fed
to a device that way, a step moves 4 bytes per pixel over PCIe -- 208 MB for a 65536-column batch, which costs
20x the step itself.  `UByteBatch` keeps the dataset's bytes: a batch crosses PCIe as 1 byte per pixel and 1 byte
per label and `ag_ubyte_to_float` / `ag_onehot_labels` rebuild exactly the arrays the reference builds (IEEE
division by 255; exact 0/1), in place, so a captured step graph keeps reading the same device arrays.
"""
import ctypes as C
import gzip
import struct

import numpy as np

from . import darray as D
from . import lazy as _lazy
from ._lib import check, DimensionMismatch


def read_idx_ubyte(path):
    """Parse an idx-ubyte file (optionally .gz) into a uint8 array; examples/mnist.jl:83-86 skips the same
    16-byte (images) / 8-byte (labels) header by offset.  Images come back as (rows*cols, n) column-major."""
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rb") as f:
        raw = f.read()
    zero, dtype, nd = struct.unpack(">HBB", raw[:4])
    if zero != 0 or dtype != 0x08:
        raise ValueError("%s: not an idx-ubyte file" % path)
    dims = struct.unpack(">" + "I" * nd, raw[4:4 + 4 * nd])
    data = np.frombuffer(raw, dtype=np.uint8, offset=4 + 4 * nd)
    if data.size != int(np.prod(dims)):
        raise ValueError("%s: truncated idx file" % path)
    if nd == 1:
        return data.copy()
    return np.asfortranarray(data.reshape(dims[0], -1).T)          # (pixels, n): one image per column


class UByteBatch:
    """Device-side minibatch rebuilt from bytes: `x` (nfeat, ncols) = pixels ./ divisor, `y` (nclass, ncols) one-hot.

    load(pixels, labels) takes host uint8 arrays (ideally pinned, see `pinned_empty`), copies them to the device
    asynchronously and converts in place; the DArrays `x` and `y` keep their addresses from batch to batch."""

    def __init__(self, nfeat, ncols, nclass=10, dtype=np.float32, divisor=255.0, zero_to=10):
        self.nfeat, self.ncols, self.nclass = int(nfeat), int(ncols), int(nclass)
        self.divisor, self.zero_to = float(divisor), int(zero_to)
        self.x = D.DArray.empty((self.nfeat, self.ncols), dtype)
        self.y = D.DArray.empty((self.nclass, self.ncols), dtype)
        # two staging buffer sets (prefetch slots): batch i+1 crosses PCIe into one while batch i is converted from
        # the other, so the copy engine never waits for a conversion
        self._px = [D._Buffer(max(self.nfeat * self.ncols, 16)) for _ in range(2)]
        self._lb = [D._Buffer(max(self.ncols, 16)) for _ in range(2)]
        self._fill, self._cur = 0, 0           # slot the next prefetch fills / slot the next convert reads

    @property
    def h2d_bytes(self):
        return self.nfeat * self.ncols + self.ncols

    def prefetch(self, pixels, labels):
        """Start copying the NEXT batch's bytes on the copy stream (they land in the staging buffers while the step
        on the current batch runs); `convert()` then rebuilds `x` and `y` from them on the main stream."""
        pixels, labels = self._check(pixels, labels)
        k = self._fill
        check(D.lib.ag_prefetch_slot(k))
        check(D.lib.ag_copy_h2d_prefetch(C.c_void_p(self._px[k].ptr), C.c_void_p(pixels.ctypes.data), pixels.size))
        check(D.lib.ag_copy_h2d_prefetch(C.c_void_p(self._lb[k].ptr), C.c_void_p(labels.ctypes.data), labels.size))
        self._fill = k ^ 1

    def convert(self):
        """Rebuild `x`, `y` from the oldest prefetched batch (prefetch and convert alternate: at most two batches are
        in flight)."""
        _lazy.flush()                                   # pending operations must read the previous batch first
        k = self._cur
        check(D.lib.ag_prefetch_slot(k))
        check(D.lib.ag_prefetch_wait())
        self._convert(k)
        check(D.lib.ag_prefetch_release())
        self._cur = k ^ 1
        return self.x, self.y

    def _check(self, pixels, labels):
        pixels, labels = np.asarray(pixels), np.asarray(labels)
        if pixels.dtype != np.uint8 or labels.dtype != np.uint8:
            raise TypeError("UByteBatch.load takes uint8 arrays")
        if pixels.size != self.nfeat * self.ncols or labels.size != self.ncols:
            raise DimensionMismatch("UByteBatch.load: %s pixels / %s labels for a (%d, %d) batch"
                                    % (pixels.shape, labels.shape, self.nfeat, self.ncols))
        if not (pixels.flags.f_contiguous or pixels.flags.c_contiguous and pixels.ndim == 1):
            pixels = np.asfortranarray(pixels)
        return pixels, labels

    def load(self, pixels, labels):
        pixels, labels = self._check(pixels, labels)
        _lazy.flush()                                   # pending operations must read the previous batch first
        check(D.lib.ag_copy_h2d(C.c_void_p(self._px[0].ptr), C.c_void_p(pixels.ctypes.data), pixels.size))
        check(D.lib.ag_copy_h2d(C.c_void_p(self._lb[0].ptr), C.c_void_p(labels.ctypes.data), labels.size))
        self._convert(0)
        return self.x, self.y

    def _convert(self, k):
        check(D.lib.ag_ubyte_to_float(self.x.dt, C.c_void_p(self._px[k].ptr), D.vptr(self.x), self.x.size, self.divisor))
        check(D.lib.ag_onehot_labels(self.y.dt, C.c_void_p(self._lb[k].ptr), self.ncols, self.nclass, self.zero_to,
                                     D.vptr(self.y)))


class StagedInput:
    """Double-buffered host -> device feed for arrays a captured step graph reads at fixed addresses: `prefetch`
    copies the NEXT batch from pinned host memory into staging buffers on the copy stream while the current step
    runs; `commit` moves it into the live arrays (device-to-device, on the main stream)."""

    def __init__(self, *arrays):
        self.arrays = arrays
        self.stage = [[D._Buffer(max(a.nbytes, 16)) for a in arrays] for _ in range(2)]     # two prefetch slots
        self._fill, self._cur = 0, 0

    def prefetch(self, *host_ptrs):
        k = self._fill
        check(D.lib.ag_prefetch_slot(k))
        for a, st, hp in zip(self.arrays, self.stage[k], host_ptrs):
            check(D.lib.ag_copy_h2d_prefetch(C.c_void_p(st.ptr), C.c_void_p(int(hp)), a.nbytes))
        self._fill = k ^ 1

    def commit(self):
        _lazy.flush()
        k = self._cur
        check(D.lib.ag_prefetch_slot(k))
        check(D.lib.ag_prefetch_wait())
        for a, st in zip(self.arrays, self.stage[k]):
            check(D.lib.ag_copy_d2d(D.vptr(a), C.c_void_p(st.ptr), a.nbytes))
        check(D.lib.ag_prefetch_release())
        self._cur = k ^ 1


# ---- character data of examples/charlm.jl:166-203 -------------------------------------------------------------------

def read_chars(source, char_limit=0):
    """`loaddata` (examples/charlm.jl:166-190): the characters of a text (a path, or the text itself) as ids in order
    of first appearance.  Returns (ids, char_to_index, index_to_char); ids are 0-based here (the reference's Int32
    indices minus one), so `index_to_char[ids[k]]` is the k-th character."""
    import os
    text = source
    if isinstance(source, (str, bytes)) and len(source) < 4096 and os.path.exists(source):
        with open(source, "r", encoding="utf-8", errors="replace") as f:
            text = f.read()
    if isinstance(text, bytes):
        text = text.decode("utf-8", errors="replace")
    char_to_index, ids = {}, []
    for c in text:
        ids.append(char_to_index.setdefault(c, len(char_to_index)))
        if char_limit > 0 and len(ids) >= char_limit:
            break
    index_to_char = [None] * len(char_to_index)
    for c, i in char_to_index.items():
        index_to_char[i] = c
    return np.asarray(ids, dtype=np.int64), char_to_index, index_to_char


def charlm_minibatch_ids(ids, batch_size):
    """`minibatch` (examples/charlm.jl:192-203) without the one-hot matrices: batch i holds, for each of the
    batch_size parallel streams j, the character at position i + nbatch*j (0-based), i.e. exactly the column in which
    the reference sets d[char, j] = 1.  Returns an (nbatch, batch_size) int64 array."""
    ids = np.asarray(ids, dtype=np.int64)
    nbatch = ids.size // batch_size
    return np.ascontiguousarray(ids[:nbatch * batch_size].reshape(batch_size, nbatch).T)


class CharBatches:
    """Device form of the charlm data: `ids[t]` (IArray, for the embedding lookup W[:, ids]) and `onehot(t)` (the
    (V, B) matrix the reference builds on the host, rebuilt on the device from one byte per entry) for a window of
    consecutive batches."""

    def __init__(self, batch_ids, vocab, dtype=np.float32):
        if vocab > 255:
            raise DimensionMismatch("CharBatches: vocabulary of %d does not fit one byte per label" % vocab)
        self.batch_ids = np.asarray(batch_ids, dtype=np.int64)
        self.vocab, self.dtype = int(vocab), np.dtype(dtype)
        self.ids = [D.IArray(row) for row in self.batch_ids]
        self._lb = D._Buffer(max(self.batch_ids.size, 16))
        flat = np.ascontiguousarray((self.batch_ids + 1).astype(np.uint8))     # 1-based classes
        check(D.lib.ag_copy_h2d(C.c_void_p(self._lb.ptr), C.c_void_p(flat.ctypes.data), flat.size))
        D._sync_stream()

    def onehot(self, t):
        B = self.batch_ids.shape[1]
        out = D.DArray.empty((self.vocab, B), self.dtype)
        check(D.lib.ag_onehot_labels(out.dt, C.c_void_p(self._lb.ptr + t * B), B, self.vocab, 0, D.vptr(out)))
        return out
