"""Staged host <-> device copies (csrc/hostcopy.cu: pinned staging buffers + worker threads) that
carry the numpy arrays of the drop-in calls: byte-exact round trips at awkward sizes and thread
counts, and the *_host entry points above the 64 MB switch against the plain copies."""
import ctypes

import numpy
import pytest

from openfermion_b200 import _lib, hamiltonians
from openfermion_b200.linalg import LinearQubitOperator

pytestmark = pytest.mark.gpu
c_int, c_size_t, vp = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p


@pytest.mark.parametrize('nbytes', [1, 4095, (8 << 20) - 1, 8 << 20, (8 << 20) + 1, (40 << 20) + 12345, 200 << 20])
@pytest.mark.parametrize('threads', [1, 3, 8])
def test_round_trip_is_byte_exact(nbytes, threads):
    rng = numpy.random.default_rng(nbytes % 1000 + threads)
    src = rng.integers(0, 256, size=nbytes, dtype=numpy.uint8)
    buf = _lib.DeviceBuffer(nbytes)
    try:
        _lib.call('ofb_copy_h2d_staged', vp(buf.ptr), _lib.ptr(src), c_size_t(nbytes), c_int(0), c_int(threads))
        plain = _lib.to_host(buf, nbytes, numpy.uint8)  # cudaMemcpy sees what the staged copy wrote
        assert numpy.array_equal(plain, src)
        back = numpy.empty(nbytes, dtype=numpy.uint8)  # fresh pages: first touch inside the copy
        _lib.call('ofb_copy_d2h_staged', _lib.ptr(back), vp(buf.ptr), c_size_t(nbytes), c_int(0), c_int(threads))
        assert numpy.array_equal(back, src)
    finally:
        buf.free()


def test_zero_bytes_and_null_arguments():
    _lib.call('ofb_copy_h2d_staged', None, None, c_size_t(0), c_int(0), c_int(0))
    with pytest.raises(_lib.OfbError):
        _lib.call('ofb_copy_d2h_staged', None, None, c_size_t(16), c_int(0), c_int(0))


def test_matvec_above_the_switch_equals_plain_copies(monkeypatch):
    # 23 qubits: 128 MB per vector -> staged H2D and staged slab-wise D2H in ofb_apply_host
    n = 23
    op = hamiltonians.hubbard_jw(2, 3)  # acts on 12 of the 23 qubits
    lin = LinearQubitOperator(op, n)
    rng = numpy.random.default_rng(3)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    got = lin * v
    monkeypatch.setenv('OFB_STAGED_COPY', '0')
    want = LinearQubitOperator(op, n) * v
    assert numpy.array_equal(got, want)
    # two consecutive calls reuse the staging buffers
    assert numpy.array_equal(lin * v, want)


def _shifted(op, shift):
    from openfermion_b200.ops import QubitOperator
    out = QubitOperator()
    out.terms.update({tuple((q + shift, a) for q, a in key): value for key, value in op.terms.items()})
    return out


@pytest.mark.parametrize('case', ['mixed', 'all_groups_inside_a_slab'])
def test_input_slabs_overlapping_the_kernel(case, monkeypatch):
    # 25 qubits: 512 MB per vector, 64 MB slabs -> ofb_apply_host starts the groups that gather
    # inside a slab while the next input slab is still being copied (OFB_H2D_OVERLAP=0: one copy,
    # then the kernels).  Both orders must give the same bytes as the plain-copy path.
    n = 25
    if case == 'mixed':
        op = hamiltonians.random_pauli_sum(n, 40, 11, False)
    else:  # qubits 3..24 = index bits 21..0: no x mask touches the three slab bits
        op = _shifted(hamiltonians.random_pauli_sum(n - 3, 30, 12, True), 3)
    rng = numpy.random.default_rng(4)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    got = LinearQubitOperator(op, n) * v
    monkeypatch.setenv('OFB_H2D_OVERLAP', '0')
    assert numpy.array_equal(LinearQubitOperator(op, n) * v, got)
    monkeypatch.setenv('OFB_STAGED_COPY', '0')
    assert numpy.array_equal(LinearQubitOperator(op, n) * v, got)
    # sampled outputs against the oracle's reference-order term sum
    import pauli_oracle as po
    idx = rng.integers(0, 1 << n, size=64)
    want = po.matvec_sampled(op.terms, n, v, idx)
    assert numpy.max(numpy.abs(got[idx] - want)) <= 1e-12 * numpy.max(numpy.abs(want))
