"""Jordan-Wigner symmetry sectors as first-class index spaces.

The reference enumerates a sector with itertools (`jw_number_indices`, `jw_sz_indices`,
linalg/sparse_tools.py:250-371), slices the materialised 2^n x 2^n matrix with those lists
(:375-423) and expands sector vectors back with fancy indexing (:507-509).  Here the same
enumeration ORDER is encoded in two small lookup tables

    position(j) = A[hi * K + class(lo)] + B[lo * K + class(hi)],      j = (hi, lo)

so that the device can rank and un-rank basis states without the lists (a 4x4 half-filled
Hubbard sector has 1.66e8 states; its itertools list would take minutes and gigabytes).
The table builders below are pure numpy and are checked against the itertools enumeration
in tests/test_sectors_cpu.py.  Qubit q is index bit n_qubits-1-q; spin-up orbitals are the
even qubits, spin-down the odd ones (sparse_tools.py:26-36 up_index / down_index).
"""
import ctypes
import itertools
from math import comb

import numpy
import scipy.sparse

from openfermion_b200 import _lib
from openfermion_b200._lib import c_void_p, c_int, c_int64


def _popcount(values):
    values = numpy.asarray(values, dtype=numpy.uint64).copy()
    count = numpy.zeros(values.shape, dtype=numpy.int64)
    while values.any():
        count += (values & numpy.uint64(1)).astype(numpy.int64)
        values >>= numpy.uint64(1)
    return count


def _extract_bits(values, mask):
    """Parallel bit extract: the bits of `values` selected by `mask`, packed towards bit 0
    in the same order."""
    values = numpy.asarray(values, dtype=numpy.uint64)
    out = numpy.zeros(values.shape, dtype=numpy.uint64)
    k = 0
    bit = 0
    while mask >> bit:
        if (mask >> bit) & 1:
            out |= ((values >> numpy.uint64(bit)) & numpy.uint64(1)) << numpy.uint64(k)
            k += 1
        bit += 1
    return out


def _descending_rank(width):
    """rank[v] = #{v' > v with popcount(v') == popcount(v)} over width-bit strings: the
    position of v when the strings of one popcount are listed as itertools.combinations of
    the set positions counted from the top bit (lexicographic == descending integer)."""
    size = 1 << width
    values = numpy.arange(size, dtype=numpy.uint64)
    pc = _popcount(values)
    rank = numpy.zeros(size, dtype=numpy.int64)
    for c in range(width + 1):
        members = numpy.nonzero(pc == c)[0]
        rank[members] = numpy.arange(len(members) - 1, -1, -1)
    return rank, pc


def _descending_base(width_hi, width_lo, total):
    """base[h] = #{(h', l') : h' > h, popcount(h') + popcount(l') == total} for width_hi-bit
    prefixes h and width_lo-bit suffixes."""
    size = 1 << width_hi
    pc = _popcount(numpy.arange(size, dtype=numpy.uint64))
    block = numpy.array([comb(width_lo, total - int(c)) if 0 <= total - int(c) <= width_lo else 0
                         for c in range(width_hi + 1)], dtype=numpy.int64)[pc]
    # states with a larger prefix come first
    suffix = numpy.concatenate([numpy.cumsum(block[::-1])[::-1][1:], [0]])
    return suffix.astype(numpy.int64)


def default_lo_bits(n_qubits):
    return n_qubits // 2


def number_sector_dim(n_qubits, n_electrons):
    return comb(n_qubits, n_electrons) if 0 <= n_electrons <= n_qubits else 0


def sz_sector_dim(n_qubits, sz_value, n_electrons=None):
    """len(jw_sz_indices(...)) without enumerating (same argument checks)."""
    sz_integer = _check_sz_arguments(float(sz_value), n_qubits, n_electrons)
    n_sites = n_qubits // 2
    if n_electrons is not None:
        num_up = (n_electrons + sz_integer) // 2
        num_down = n_electrons - num_up
        if not (0 <= num_up <= n_sites and 0 <= num_down <= n_sites):
            return 0
        return comb(n_sites, num_up) * comb(n_sites, num_down)
    shift = abs(sz_integer)
    return sum(comb(n_sites, n) * comb(n_sites, n - shift) for n in range(shift, n_sites + 1))


class SectorTables:
    """Host description of a table-mode sector (what ofb_sector_create_tables takes)."""

    def __init__(self, n_qubits, lo_bits, n_classes, class_mask, table_hi, table_lo, plus, minus,
                 value, dim):
        self.n_qubits = n_qubits
        self.lo_bits = lo_bits
        self.n_classes = n_classes
        self.class_mask = class_mask
        self.table_hi = numpy.ascontiguousarray(table_hi, dtype=numpy.uint64)
        if len(table_lo) and int(numpy.max(table_lo)) >= 1 << 32:
            raise ValueError('sector too large for 32-bit low-half offsets')
        self.table_lo = numpy.ascontiguousarray(table_lo, dtype=numpy.uint32)
        self.plus = numpy.ascontiguousarray(plus, dtype=numpy.uint64)
        self.minus = numpy.ascontiguousarray(minus, dtype=numpy.uint64)
        self.value = numpy.ascontiguousarray(value, dtype=numpy.int32)
        self.dim = int(dim)

    # host restatement of the device lookup (tests, small cases)
    def member(self, j):
        j = numpy.asarray(j, dtype=numpy.uint64)
        ok = numpy.ones(j.shape, dtype=bool)
        for p, m, v in zip(self.plus, self.minus, self.value):
            ok &= (_popcount(j & p) - _popcount(j & m)) == int(v)
        return ok

    def position(self, j):
        j = numpy.asarray(j, dtype=numpy.uint64)
        lo_mask = numpy.uint64((1 << self.lo_bits) - 1)
        hi = (j >> numpy.uint64(self.lo_bits)).astype(numpy.int64)
        lo = (j & lo_mask).astype(numpy.int64)
        if self.n_classes == 1:
            return self.table_hi[hi].astype(numpy.int64) + self.table_lo[lo].astype(numpy.int64)
        cm = numpy.uint64(self.class_mask)
        c_lo = _popcount(j & cm & lo_mask)
        c_hi = _popcount(j & cm & ~lo_mask)
        return (self.table_hi[hi * self.n_classes + c_lo].astype(numpy.int64)
                + self.table_lo[lo * self.n_classes + c_hi].astype(numpy.int64))

    def basis(self):
        """The reference's index list, rebuilt from the tables (small n only)."""
        everything = numpy.arange(1 << self.n_qubits, dtype=numpy.uint64)
        members = everything[self.member(everything)]
        out = numpy.empty(self.dim, dtype=numpy.uint64)
        out[self.position(members)] = members
        return out


def _bit_reverse(values, width):
    values = numpy.asarray(values, dtype=numpy.uint64)
    out = numpy.zeros(values.shape, dtype=numpy.uint64)
    for bit in range(width):
        out |= ((values >> numpy.uint64(bit)) & numpy.uint64(1)) << numpy.uint64(width - 1 - bit)
    return out


def number_sector_tables(n_qubits, n_electrons, lo_bits=None):
    """Tables reproducing jw_number_indices (sparse_tools.py:273-292): occupations are
    itertools.combinations(range(n_qubits), n_electrons) over BIT positions (index = sum 2^k,
    unlike jw_sz_indices), i.e. descending order of the bit-reversed index."""
    if lo_bits is None:
        lo_bits = default_lo_bits(n_qubits)
    hi_bits = n_qubits - lo_bits
    # the reversed index has the reversed low half on top
    base = _descending_base(lo_bits, hi_bits, n_electrons)
    rank, _ = _descending_rank(hi_bits)
    table_lo = base[_bit_reverse(numpy.arange(1 << lo_bits), lo_bits).astype(numpy.int64)]
    table_hi = rank[_bit_reverse(numpy.arange(1 << hi_bits), hi_bits).astype(numpy.int64)]
    full = (1 << n_qubits) - 1
    dim = comb(n_qubits, n_electrons) if 0 <= n_electrons <= n_qubits else 0
    return SectorTables(n_qubits, lo_bits, 1, 0, table_hi, table_lo, [full], [0], [n_electrons], dim)


def _spin_masks(n_qubits):
    up = sum(1 << (n_qubits - 1 - q) for q in range(0, n_qubits, 2))
    down = sum(1 << (n_qubits - 1 - q) for q in range(1, n_qubits, 2))
    return up, down


def _check_sz_arguments(sz_value, n_qubits, n_electrons):
    if n_qubits % 2 != 0:
        raise ValueError('Number of qubits must be even')
    if not (2.0 * sz_value).is_integer():
        raise ValueError('Sz value must be an integer or half-integer')
    sz_integer = int(2.0 * sz_value)
    if n_electrons is not None:
        if (n_electrons + sz_integer) % 2 != 0 or n_electrons < abs(sz_integer):
            raise ValueError('The specified particle number and sz value are incompatible.')
    return sz_integer


def sz_number_sector_tables(n_qubits, sz_value, n_electrons, lo_bits=None):
    """Tables reproducing jw_sz_indices with a fixed particle number (sparse_tools.py:333-346):
    up occupations outermost, down occupations innermost."""
    sz_integer = _check_sz_arguments(float(sz_value), n_qubits, n_electrons)
    num_up = (n_electrons + sz_integer) // 2
    num_down = n_electrons - num_up
    n_sites = n_qubits // 2
    if lo_bits is None:
        lo_bits = default_lo_bits(n_qubits)
    hi_bits = n_qubits - lo_bits
    up, down = _spin_masks(n_qubits)
    lo_mask = (1 << lo_bits) - 1
    n_down_states = comb(n_sites, num_down) if 0 <= num_down <= n_sites else 0
    n_up_states = comb(n_sites, num_up) if 0 <= num_up <= n_sites else 0
    hi_values = numpy.arange(1 << hi_bits, dtype=numpy.uint64)
    lo_values = numpy.arange(1 << lo_bits, dtype=numpy.uint64)
    parts = {}
    for name, mask, total in (('up', up, num_up), ('down', down, num_down)):
        m_hi, m_lo = (mask >> lo_bits), (mask & lo_mask)
        w_hi, w_lo = bin(m_hi).count('1'), bin(m_lo).count('1')
        base = _descending_base(w_hi, w_lo, total)
        rank, _ = _descending_rank(w_lo)
        parts[name] = (base[_extract_bits(hi_values, m_hi).astype(numpy.int64)],
                       rank[_extract_bits(lo_values, m_lo).astype(numpy.int64)])
    table_hi = parts['up'][0] * n_down_states + parts['down'][0]
    table_lo = parts['up'][1] * n_down_states + parts['down'][1]
    return SectorTables(n_qubits, lo_bits, 1, 0, table_hi, table_lo, [up, down], [0, 0],
                        [num_up, num_down], n_up_states * n_down_states)


def sz_sector_tables(n_qubits, sz_value, lo_bits=None):
    """Tables reproducing jw_sz_indices without a particle number (sparse_tools.py:347-371):
    blocks of n = |2 sz| .. n_sites 'more' spins, inside a block the 'more' occupations
    outermost and the n - |2 sz| 'less' occupations innermost."""
    sz_integer = _check_sz_arguments(float(sz_value), n_qubits, None)
    n_sites = n_qubits // 2
    shift = abs(sz_integer)
    if lo_bits is None:
        lo_bits = default_lo_bits(n_qubits)
    hi_bits = n_qubits - lo_bits
    up, down = _spin_masks(n_qubits)
    more, less = (down, up) if sz_integer < 0 else (up, down)
    lo_mask = (1 << lo_bits) - 1
    more_hi, more_lo = more >> lo_bits, more & lo_mask
    less_hi, less_lo = less >> lo_bits, less & lo_mask
    wm_hi, wm_lo = bin(more_hi).count('1'), bin(more_lo).count('1')
    wl_hi, wl_lo = bin(less_hi).count('1'), bin(less_lo).count('1')
    n_classes = max(wm_hi, wm_lo) + 1
    hi_values = numpy.arange(1 << hi_bits, dtype=numpy.uint64)
    lo_values = numpy.arange(1 << lo_bits, dtype=numpy.uint64)
    m_of_hi = _extract_bits(hi_values, more_hi).astype(numpy.int64)
    l_of_hi = _extract_bits(hi_values, less_hi).astype(numpy.int64)
    m_of_lo = _extract_bits(lo_values, more_lo).astype(numpy.int64)
    l_of_lo = _extract_bits(lo_values, less_lo).astype(numpy.int64)
    c_of_hi = _popcount(m_of_hi.astype(numpy.uint64))
    c_of_lo = _popcount(m_of_lo.astype(numpy.uint64))
    rank_more, _ = _descending_rank(wm_lo)
    rank_less, _ = _descending_rank(wl_lo)
    offsets, running = {}, 0
    for n_more in range(shift, n_sites + 1):
        offsets[n_more] = running
        running += comb(n_sites, n_more) * comb(n_sites, n_more - shift)
    dim = running
    table_hi = numpy.zeros((1 << hi_bits, n_classes), dtype=numpy.int64)
    table_lo = numpy.zeros((1 << lo_bits, n_classes), dtype=numpy.int64)
    for n_more in range(shift, n_sites + 1):
        n_less_states = comb(n_sites, n_more - shift)
        base_more = _descending_base(wm_hi, wm_lo, n_more)[m_of_hi]
        base_less = _descending_base(wl_hi, wl_lo, n_more - shift)[l_of_hi]
        for c_lo in range(n_classes):
            rows = numpy.nonzero(c_of_hi + c_lo == n_more)[0]
            table_hi[rows, c_lo] = offsets[n_more] + base_more[rows] * n_less_states + base_less[rows]
        for c_hi in range(n_classes):
            rows = numpy.nonzero(c_of_lo + c_hi == n_more)[0]
            table_lo[rows, c_hi] = rank_more[m_of_lo[rows]] * n_less_states + rank_less[l_of_lo[rows]]
    return SectorTables(n_qubits, lo_bits, n_classes, more, table_hi.reshape(-1), table_lo.reshape(-1),
                        [more], [less], [shift], dim)


# ---- reference enumerations as arrays (host; small sectors and custom spin maps) -----------
def number_indices_list(n_electrons, n_qubits):
    """jw_number_indices (sparse_tools.py:273-292) as a list of Python ints."""
    occupations = itertools.combinations(range(n_qubits), n_electrons)
    return [sum(2**k for k in occupation) for occupation in occupations]


def sz_indices_list(sz_value, n_qubits, n_electrons=None, up_index=None, down_index=None):
    """jw_sz_indices (sparse_tools.py:268-371) for arbitrary spin maps."""
    up_index = up_index or (lambda site: 2 * site)
    down_index = down_index or (lambda site: 2 * site + 1)
    sz_integer = _check_sz_arguments(float(sz_value), n_qubits, n_electrons)
    n_sites = n_qubits // 2
    indices = []
    if n_electrons is not None:
        num_up = (n_electrons + sz_integer) // 2
        num_down = n_electrons - num_up
        blocks = [(up_index, num_up, down_index, num_down)]
    else:
        more_map, less_map = (down_index, up_index) if sz_integer < 0 else (up_index, down_index)
        blocks = [(more_map, n, less_map, n - abs(sz_integer))
                  for n in range(abs(sz_integer), n_sites + 1)]
    for outer_map, outer_count, inner_map, inner_count in blocks:
        inner_values = [sum(2 ** (n_qubits - 1 - inner_map(site)) for site in occupation)
                        for occupation in itertools.combinations(range(n_sites), inner_count)]
        for occupation in itertools.combinations(range(n_sites), outer_count):
            outer_value = sum(2 ** (n_qubits - 1 - outer_map(site)) for site in occupation)
            indices.extend(outer_value + inner for inner in inner_values)
    return indices


# ---- device object -------------------------------------------------------------------------
class Sector:
    """Device-resident sector (an `ofb_sector`): index map + basis list on one GPU."""

    def __init__(self, handle, n_qubits, device, tables=None):
        self.handle = handle
        self.n_qubits = n_qubits
        self.device = device
        self.tables = tables
        dim = c_int64()
        mode = c_int()
        _lib.call('ofb_sector_info', handle, None, ctypes.byref(dim), ctypes.byref(mode))
        self.dim = dim.value
        self.mode = mode.value

    @classmethod
    def from_tables(cls, tables, device=0):
        _lib.require_device()
        handle = c_void_p()
        _lib.call('ofb_sector_create_tables', c_int(tables.n_qubits), c_int(tables.lo_bits),
                  c_int(tables.n_classes), ctypes.c_uint64(tables.class_mask), _lib.ptr(tables.table_hi),
                  _lib.ptr(tables.table_lo), c_int(len(tables.value)), _lib.ptr(tables.plus),
                  _lib.ptr(tables.minus), _lib.ptr(tables.value), c_int64(tables.dim), c_int(device),
                  ctypes.byref(handle))
        return cls(handle, tables.n_qubits, device, tables)

    @classmethod
    def from_list(cls, n_qubits, basis, device=0):
        _lib.require_device()
        basis = numpy.ascontiguousarray(basis, dtype=numpy.uint64)
        handle = c_void_p()
        _lib.call('ofb_sector_create_list', c_int(n_qubits), c_int64(len(basis)), _lib.ptr(basis),
                  c_int(device), ctypes.byref(handle))
        return cls(handle, n_qubits, device)

    @classmethod
    def number(cls, n_qubits, n_electrons, device=0):
        return cls.from_tables(number_sector_tables(n_qubits, n_electrons), device)

    @classmethod
    def sz(cls, n_qubits, sz_value, n_electrons=None, up_index=None, down_index=None, device=0):
        if up_index is not None or down_index is not None:
            return cls.from_list(n_qubits, sz_indices_list(sz_value, n_qubits, n_electrons, up_index,
                                                           down_index), device)
        if n_electrons is None:
            return cls.from_tables(sz_sector_tables(n_qubits, sz_value), device)
        return cls.from_tables(sz_number_sector_tables(n_qubits, sz_value, n_electrons), device)

    def basis(self):
        out = numpy.empty(self.dim, dtype=numpy.uint64)
        _lib.call('ofb_sector_basis', self.handle, _lib.ptr(out))
        return out

    # ---- kernels --------------------------------------------------------------------------
    def _check(self, plan):
        if plan.n_qubits != self.n_qubits:
            raise ValueError('operator and sector act on different numbers of qubits')

    def apply_device(self, plan, d_in, d_out, ncols=1, ld=None, stream=None):
        _lib.call('ofb_sector_apply', plan.handle, self.handle, c_void_p(d_in), c_void_p(d_out),
                  c_int(ncols), c_int64(self.dim if ld is None else ld), c_void_p(stream))

    def apply_host(self, plan, columns):
        self._check(plan)
        cols = numpy.asarray(columns, dtype=numpy.complex128)
        ncols = 1 if cols.ndim == 1 else cols.shape[1]
        cols = numpy.asfortranarray(cols) if cols.ndim == 2 else numpy.ascontiguousarray(cols)
        out = numpy.empty_like(cols, order='F' if cols.ndim == 2 else 'C')
        _lib.call('ofb_sector_apply_host', plan.handle, self.handle, _lib.ptr(cols), _lib.ptr(out),
                  c_int(ncols))
        return out

    def expectation_device(self, plan, d_state, stream=None):
        out = (ctypes.c_double * 2)()
        _lib.call('ofb_sector_expectation', plan.handle, self.handle, c_void_p(d_state), out,
                  c_void_p(stream))
        return complex(out[0], out[1])

    def expectation_host(self, plan, state):
        self._check(plan)
        state = numpy.ascontiguousarray(state, dtype=numpy.complex128).reshape(-1)
        if state.size != self.dim:
            raise ValueError('Dimension mismatch')
        buf = _lib.to_device(state)
        return self.expectation_device(plan, buf.ptr)

    def diagonal_host(self, plan):
        self._check(plan)
        buf = _lib.DeviceBuffer(self.dim * 8)
        _lib.call('ofb_sector_diagonal', plan.handle, self.handle, c_void_p(buf.ptr), c_void_p(None))
        return _lib.to_host(buf, self.dim, numpy.float64)

    def restrict_state_host(self, state):
        """state[indices] (jw_number_restrict_state / jw_sz_restrict_state)."""
        state = numpy.ascontiguousarray(state, dtype=numpy.complex128).reshape(-1)
        if state.size != 1 << self.n_qubits:
            raise ValueError('Dimension mismatch')
        full = _lib.to_device(state)
        out = _lib.DeviceBuffer(self.dim * 16)
        _lib.call('ofb_sector_gather', self.handle, c_void_p(full.ptr), c_void_p(out.ptr), c_void_p(None))
        return _lib.to_host(out, self.dim, numpy.complex128)

    def expand_state_host(self, state):
        """zeros(2^n); expanded[indices] = state (sparse_tools.py:507-509)."""
        state = numpy.ascontiguousarray(state, dtype=numpy.complex128).reshape(-1)
        if state.size != self.dim:
            raise ValueError('Dimension mismatch')
        part = _lib.to_device(state)
        full = _lib.DeviceBuffer((1 << self.n_qubits) * 16)
        _lib.call('ofb_sector_scatter', self.handle, c_void_p(part.ptr), c_void_p(full.ptr), c_void_p(None))
        return _lib.to_host(full, 1 << self.n_qubits, numpy.complex128)

    def csc_host(self, plan, dtype=numpy.complex128):
        """The restricted operator as scipy CSC (dim x dim), rows ascending, exact zeros dropped."""
        self._check(plan)
        nnz = c_int64()
        bits = c_int()
        _lib.call('ofb_sector_csc_count', plan.handle, self.handle, ctypes.byref(nnz), ctypes.byref(bits),
                  c_void_p(None))
        idx_dtype = numpy.int32 if bits.value == 32 else numpy.int64
        indptr = numpy.empty(self.dim + 1, dtype=idx_dtype)
        indices = numpy.empty(nnz.value, dtype=idx_dtype)
        data = numpy.empty(nnz.value, dtype=numpy.complex128)
        _lib.call('ofb_sector_csc_fill_host', plan.handle, self.handle, _lib.ptr(indptr),
                  _lib.ptr(indices), _lib.ptr(data), bits)
        if dtype == numpy.float64:
            data = numpy.ascontiguousarray(data.real)
        mat = scipy.sparse.csc_matrix((data, indices, indptr), shape=(self.dim, self.dim))
        mat.has_sorted_indices = True
        return mat

    def close(self):
        if getattr(self, 'handle', None):
            _lib.call('ofb_sector_destroy', self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
