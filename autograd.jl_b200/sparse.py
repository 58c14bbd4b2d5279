"""Sparse gradients on device and the host-side index bookkeeping behind getindex/ungetindex.

Reference: `Sparse{T,N}(container, values, indices)` (src/sparse.jl:15-21) is what `ungetindex`
returns outside of recording (src/getindex.jl:61-79); `addto!(dense, Sparse)` / `full` replay the
(index, value) pairs through the sequential scalar loop `addtoindex!` (src/addto.jl:91-129).

Here the (index, value) pairs of ALL blocks are flattened on the host into one int64 destination
vector (bit-exact with the reference's iteration order: first index fastest, blocks in append
order), the value blocks are concatenated on device, and ONE deterministic sort + segmented reduce
(ag_scatter_add) applies them.  Indices are 0-based here (host language is Python); semantics are
Julia's: one index per dimension = outer-product selection, a single index on an N-d array = linear
column-major index, Bool masks select, repeated indices accumulate.
"""
import numpy as np

from . import darray as D
from ._lib import DimensionMismatch
from .darray import DArray, IArray, isnum


def _prod(s):
    n = 1
    for v in s:
        n *= int(v)
    return n


def to_indices(shape, idx):
    """Normalise an index tuple against `shape` (Base.to_indices as used at src/addto.jl:95).

    Returns (positions, outshape, extents): per indexed dimension an int64 vector of 0-based
    positions, the shape getindex would return, and the extent of each indexed dimension."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    nd = len(shape)
    if len(idx) == 1 and nd != 1:
        extents = [_prod(shape)]        # linear indexing / full Bool mask
    else:
        flat = []
        for i in idx:                   # a CartesianIndex (tuple inside the tuple) splices
            if isinstance(i, tuple):
                flat.extend(i)
            else:
                flat.append(i)
        idx = tuple(flat)
        if len(idx) < nd:
            k = len(idx)                # partial linear indexing: trailing dims fold into the last index
            extents = list(shape[:k - 1]) + [_prod(shape[k - 1:])]
        else:
            extents = list(shape) + [1] * (len(idx) - nd)
    positions, outshape = [], []
    for i, n in zip(idx, extents):
        if isinstance(i, slice):
            p = np.arange(n, dtype=np.int64)[i]
            outshape.append(p.size)
        elif isinstance(i, range):
            p = np.fromiter(i, dtype=np.int64, count=len(i))
            outshape.append(p.size)
        elif isinstance(i, (int, np.integer)) and not isinstance(i, (bool, np.bool_)):
            p = np.array([int(i)], dtype=np.int64)
        else:
            a = np.asarray(i)
            if a.dtype == np.bool_:
                if a.size != n:
                    raise IndexError("BoundsError: Bool mask of length %d on extent %d" % (a.size, n))
                p = np.flatnonzero(a.reshape(-1, order="F")).astype(np.int64)
                outshape.append(p.size)
            else:
                p = a.astype(np.int64).reshape(-1, order="F")
                outshape.extend(a.shape)
        if p.size and (p.min() < 0 or p.max() >= n):
            raise IndexError("BoundsError")
        positions.append(p)
    return positions, tuple(outshape), extents


def flat_dest(shape, idx):
    """0-based column-major destinations of x[idx...] in the reference's iteration order (first index
    fastest): the index bookkeeping of addtoindex! (src/addto.jl:107-129), bit-exact."""
    positions, outshape, extents = to_indices(shape, idx)
    dest = np.zeros(1, dtype=np.int64)
    stride = 1
    for p, n in zip(positions, extents):
        dest = (dest[None, :] + (p * stride)[:, None]).reshape(-1)
        stride *= n
    return dest, outshape


class Sparse:
    """Gradient of getindex: value blocks + the index tuples they belong to, over a container."""

    def __init__(self, container, values, indices):
        self.container, self.values, self.indices = container, list(values), list(indices)

    @property
    def shape(self):
        return self.container.shape

    @property
    def dtype(self):
        return self.container.dtype

    # ---- arithmetic (src/sparse.jl:65-83)
    def __mul__(self, n):
        return Sparse(self.container, [v * n for v in self.values], self.indices)

    __rmul__ = __mul__

    def __truediv__(self, n):
        return Sparse(self.container, [v / n for v in self.values], self.indices)

    def __neg__(self):
        return self * -1

    def __add__(self, o):
        if isinstance(o, Sparse):
            assert self.shape == o.shape
            return Sparse(self.container, self.values + o.values, self.indices + o.indices)
        return self.addto_dense(o, inplace=False)

    __radd__ = __add__

    def __rsub__(self, a):          # a - s
        return (-self).addto_dense(a, inplace=False)

    def __sub__(self, a):           # s - a
        return self.addto_dense(-a, inplace=True)

    # ---- scatter
    def _column_blocks(self):
        """If every index is (Colon, Vector{Int}) on a matrix, return the concatenated column ids."""
        if len(self.shape) != 2:
            return None
        cols = []
        for idx in self.indices:
            if not (isinstance(idx, tuple) and len(idx) == 2 and isinstance(idx[0], slice)
                    and idx[0] == slice(None)):
                return None
            j = idx[1]
            if isinstance(j, IArray):
                cols.append(j)
                continue
            if isinstance(j, (int, np.integer)) and not isinstance(j, (bool, np.bool_)):
                j = [int(j)]
            j = np.asarray(j)
            if j.dtype == np.bool_ or j.ndim != 1:
                return None
            cols.append(j.astype(np.int64))
        return cols

    def addto_dense(self, a, inplace=True):
        """a[idx...] += val for every (idx, val) pair, repeated indices summed
        (addto!(a::AbstractArray, b::Sparse), src/addto.jl:91-98)."""
        if a.shape != self.shape:
            raise DimensionMismatch("Sparse container %s vs accumulator %s" % (self.shape, a.shape))
        if not inplace:
            a = D.copy(a)
        if not self.values:
            return a
        cols = self._column_blocks()
        rows = self.shape[0] if len(self.shape) == 2 else 0
        if cols is not None and all(isinstance(v, DArray) and v.size == rows * c.size
                                    for v, c in zip(self.values, cols)):
            vals = _concat_flat(self.values, self.dtype)
            if all(isinstance(c, IArray) for c in cols):          # indices live on the device: no host round trip
                if len(cols) == 1:
                    ib, n, lo, hi = cols[0], cols[0].size, cols[0].lo, cols[0].hi
                else:
                    ib, n, lo, hi = D.concat_index(cols)
                if n and (lo < 0 or hi >= self.shape[1]):
                    raise IndexError("BoundsError")
                return D.scatter_add_dev_(a, ib, n, lo, hi, vals, block=rows)
            cols = [c.host if isinstance(c, IArray) else c for c in cols]
            dest = np.concatenate(cols) if len(cols) > 1 else cols[0]
            if dest.size and (dest.min() < 0 or dest.max() >= self.shape[1]):
                raise IndexError("BoundsError")
            return D.scatter_add_(a, dest, vals, block=rows)
        dests, vals = [], []
        for idx, v in zip(self.indices, self.values):
            idx = tuple(i.host if isinstance(i, IArray) else i for i in (idx if isinstance(idx, tuple) else (idx,)))
            d, _ = flat_dest(self.shape, idx)
            if isnum(v) or (isinstance(v, DArray) and v.size == 1 and d.size != 1):
                v = D.full_like_shape((d.size,), self.dtype, float(v))
            elif v.size != d.size:
                raise DimensionMismatch("addtoindex!: %d values for %d destinations" % (v.size, d.size))
            dests.append(d)
            vals.append(v)
        dest = np.concatenate(dests) if len(dests) > 1 else dests[0]
        return D.scatter_add_(a, dest, _concat_flat(vals, self.dtype), block=1)

    def full(self):
        """full(b::Sparse) = addto!(zeroslike(b), b)  (src/sparse.jl:57-59)."""
        return self.addto_dense(D.zeros(self.shape, self.dtype), inplace=True)

    def norm(self):
        """norm(::Sparse) ignores repeated indices (src/sparse.jl:53-54)."""
        from .rules import sumabs2
        s = 0.0
        for v in self.values:
            s += float(sumabs2(v)) if isinstance(v, DArray) else float(v) ** 2
        return s ** 0.5


def _concat_flat(blocks, dtype):
    blocks = [D.materialize(b) for b in blocks]
    if len(blocks) == 1:
        return blocks[0]
    tot = sum(b.size for b in blocks)
    out = DArray.empty((tot,), dtype)
    off, segs, isz = 0, [], np.dtype(dtype).itemsize
    for b in blocks:
        segs.append((b.ptr, out.ptr + off * isz, b.size))
        off += b.size
    D.multi_copy_flat(out.dt, segs)
    return out
