"""The BASELINE.json configurations at FULL size against the CPU oracle:
sampled outputs (H psi)[j] of the 28-qubit molecular-type Hamiltonian (config 4: T = 88 495) and
of the 32-qubit Hubbard Hamiltonian (config 5, here on one GPU) are evaluated from the
reference-order term dict (oracle/pauli_oracle.py:matvec_sampled) and must agree to 1e-12
(BASELINE.json's fp64 bar).  psi is the counter-based global random state
(ofb_vec_fill_random), whose amplitudes the host regenerates at the indices the oracle asks for
(distributed.global_random_amplitudes), so no 2^n-long vector exists on the host."""
import numpy
import pytest

import pauli_oracle as po
from openfermion_b200 import _lib, hamiltonians
from openfermion_b200.distributed import global_random_amplitudes
from openfermion_b200.linalg import LinearQubitOperator, krylov

pytestmark = pytest.mark.gpu

RTOL = 1e-12  # BASELINE.json: matvecs within 1e-12 relative (fp64)


def _sampled_check(op, n, seed, samples=256):
    dim = 1 << n
    plan = LinearQubitOperator(op, n).plan
    d_in, d_out = _lib.DeviceBuffer(dim * 16), _lib.DeviceBuffer(dim * 16)
    krylov.fill_random(dim, seed, False, d_in.ptr)
    norm = krylov.nrm2(dim, d_in.ptr)
    krylov.scal(dim, 1.0 / norm, d_in.ptr)
    plan.apply_device(d_in.ptr, d_out.ptr)
    rng = numpy.random.default_rng(seed)
    idx = numpy.unique(numpy.concatenate([rng.integers(0, dim, size=samples, dtype=numpy.int64),
                                          [0, dim - 1, dim // 2]]))
    got = numpy.array([krylov.download(1, d_out.ptr + int(j) * 16)[0] for j in idx])
    # the generator on the host reproduces the device state exactly (same integer hash, 2^-53 scaling)
    probe = numpy.array([krylov.download(1, d_in.ptr + int(j) * 16)[0] for j in idx[:8]])
    assert numpy.allclose(probe, global_random_amplitudes(seed, idx[:8]) / norm, rtol=1e-15, atol=0)
    want = po.matvec_sampled(op.terms, n, lambda index: global_random_amplitudes(seed, index) / norm, idx)
    err = numpy.max(numpy.abs(got - want)) / numpy.max(numpy.abs(want))
    assert err <= RTOL, err
    # the fused expectation kernel on the same state equals <psi|H psi> from the apply
    e_apply = krylov.dotc(dim, d_in.ptr, d_out.ptr)
    assert abs(plan.expectation_device(d_in.ptr) - e_apply) <= 1e-12 * max(1.0, abs(e_apply))
    return err


def test_config4_molecular_28_qubits_sampled_outputs_match_oracle():
    op = hamiltonians.molecular_jw(14, seed=1234)
    assert len(op) == 88495
    _sampled_check(op, 28, seed=42)


def test_config5_hubbard_32_qubits_on_one_gpu_sampled_outputs_match_oracle():
    op = hamiltonians.hubbard_jw(4, 4)
    assert len(op) == 177
    _sampled_check(op, 32, seed=42)


def test_config3_hubbard_24_qubits_sampled_outputs_match_oracle():
    _sampled_check(hamiltonians.hubbard_jw(3, 4), 24, seed=7)
