"""Drop-in `LinearQubitOperator` family backed by the sm_100a matrix-free kernel (K1).

Mirrors /root/reference/src/openfermion/linalg/linear_qubit_operator.py: same class and
function names, constructor arguments, attributes and error behaviour; `_matvec` runs
`ofb_apply_host` instead of the per-term numpy loop (linear_qubit_operator.py:107-148).
"""
import os

import numpy
import scipy.sparse.linalg

from openfermion_b200 import sectors
from openfermion_b200.compiler import compile_qubit_operator
from openfermion_b200.ops import count_qubits
from openfermion_b200.plan import PauliPlan


def get_available_cpu_count():
    """Usable CPU cores (reference config.py:29-55), honouring pytest-xdist's worker count."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    workers = os.environ.get('PYTEST_XDIST_WORKER_COUNT')
    if workers:
        cores = max(1, cores // int(workers))
    return cores


class LinearQubitOperatorOptions:
    """Options for ParallelLinearQubitOperator (linear_qubit_operator.py:38-71).

    On the GPU path a term-group "worker" is an x-mask group of one compiled plan, so no
    process pool is ever created: `pool` stays whatever the caller passed (None by default).
    """

    def __init__(self, processes=10, pool=None):
        if processes <= 0:
            raise ValueError('Invalid number of processors specified {} <= 0'.format(processes))
        self.processes = min(processes, get_available_cpu_count())
        self.pool = pool

    def get_processes(self, num):
        return max(min(num, self.processes), 1)

    def get_pool(self, num=None):
        raise NotImplementedError(
            'openfermion_b200 applies all term groups inside one CUDA kernel; there is no '
            'multiprocessing pool to hand out.'
        )


class LinearQubitOperator(scipy.sparse.linalg.LinearOperator):
    """Matrix-free action of a QubitOperator on 2^n complex128 amplitudes.

    Qubit q is index bit n_qubits-1-q (linear_qubit_operator.py:78-84).  The term table
    is compiled into a device plan on first use; constructing the object needs no GPU.
    """

    def __init__(self, qubit_operator, n_qubits=None):
        calculated_n_qubits = count_qubits(qubit_operator)
        if n_qubits is None:
            n_qubits = calculated_n_qubits
        elif n_qubits < calculated_n_qubits:
            raise ValueError(
                'Invalid number of qubits specified '
                '{} < {}.'.format(n_qubits, calculated_n_qubits)
            )
        n_hilbert = 2**n_qubits
        super().__init__(shape=(n_hilbert, n_hilbert), dtype=complex)
        self.qubit_operator = qubit_operator
        self.n_qubits = n_qubits
        self._plan = None

    @property
    def plan(self):
        """The compiled device plan (built on first access; raises without a GPU)."""
        if self._plan is None:
            x, z, c = compile_qubit_operator(self.qubit_operator, self.n_qubits)
            self._plan = PauliPlan(self.n_qubits, x, z, c)
        return self._plan

    def _matvec(self, x):
        arr = numpy.asarray(x)
        vec = numpy.ascontiguousarray(arr.reshape(-1), dtype=numpy.complex128)
        return self.plan.apply_host(vec).reshape(arr.shape)

    def _matmat(self, X):
        X = numpy.asarray(X)
        return self.plan.apply_host(numpy.asfortranarray(X, dtype=numpy.complex128))


class ParallelLinearQubitOperator(scipy.sparse.linalg.LinearOperator):
    """Same attributes as the reference's process-pool variant
    (linear_qubit_operator.py:151-209): `qubit_operator_groups`, `linear_operators`,
    `options`.  The sum over term groups is evaluated by one plan on the GPU."""

    def __init__(self, qubit_operator, n_qubits=None, options=None):
        n_qubits = n_qubits or count_qubits(qubit_operator)
        n_hilbert = 2**n_qubits
        super().__init__(shape=(n_hilbert, n_hilbert), dtype=complex)
        self.qubit_operator = qubit_operator
        self.n_qubits = n_qubits
        self.options = options or LinearQubitOperatorOptions()
        self.qubit_operator_groups = list(
            qubit_operator.get_operator_groups(self.options.processes)
        )
        self.linear_operators = [
            LinearQubitOperator(operator, n_qubits) for operator in self.qubit_operator_groups
        ]
        self._fused = LinearQubitOperator(qubit_operator, n_qubits)

    def _matvec(self, x):
        if not self.linear_operators:
            return numpy.zeros(x.shape)
        return self._fused._matvec(x)

    def _matmat(self, X):
        if not self.linear_operators:
            return numpy.zeros(numpy.asarray(X).shape)
        return self._fused._matmat(X)


class SectorLinearQubitOperator(scipy.sparse.linalg.LinearOperator):
    """Matrix-free action of a QubitOperator inside a Jordan-Wigner symmetry sector.

    As a linear map it equals `jw_number_restrict_operator(get_sparse_operator(H), n_electrons)`
    (sz_value None), `jw_sz_restrict_operator(get_sparse_operator(H), sz_value, n_electrons)`
    otherwise (linalg/sparse_tools.py:375-423), with the same basis order, but neither
    matrix is ever built: the vector has C(n, N) (or C(n/2, N_up) C(n/2, N_down)) amplitudes
    and the sector kernel ranks / un-ranks basis states on the fly (csrc/sector.cu).
    Not part of the reference API; it is what get_ground_state, expectation, Davidson and
    jw_get_ground_state_at_particle_number use when handed a symbolic operator."""

    def __init__(self, qubit_operator, n_qubits=None, n_electrons=None, sz_value=None,
                 up_index=None, down_index=None):
        if isinstance(qubit_operator, LinearQubitOperator):
            base = qubit_operator
            if n_qubits is not None and n_qubits != base.n_qubits:
                raise ValueError('Invalid number of qubits specified.')
        else:
            base = LinearQubitOperator(qubit_operator, n_qubits)
        if n_electrons is None and sz_value is None:
            raise ValueError('Specify n_electrons, sz_value or both.')
        self.linear_qubit_operator = base
        self.qubit_operator = base.qubit_operator
        self.n_qubits = base.n_qubits
        self.n_electrons = n_electrons
        self.sz_value = sz_value
        self._spin_maps = (up_index, down_index)
        self._sector = None
        self._apply_plan = None
        if sz_value is None:
            dim = sectors.number_sector_dim(self.n_qubits, n_electrons)
        elif up_index is None and down_index is None:
            dim = sectors.sz_sector_dim(self.n_qubits, sz_value, n_electrons)
        else:
            dim = len(sectors.sz_indices_list(sz_value, self.n_qubits, n_electrons, up_index,
                                              down_index))
        super().__init__(shape=(dim, dim), dtype=complex)

    @property
    def plan(self):
        return self.linear_qubit_operator.plan

    @property
    def apply_plan(self):
        """The plan the sector matvec / expectation kernels run on: the orbit-reduced term table
        (openfermion_b200/support.py) for the group conditions that sector membership implies;
        the full plan when there are none or with OFB_SUPPORT_HINTS=0."""
        if self._apply_plan is None:
            self._apply_plan = self.plan
            tables = self.sector.tables
            if os.environ.get('OFB_SUPPORT_HINTS', '1') == '1' and tables is not None:
                from openfermion_b200 import support
                x, z, c = compile_qubit_operator(self.qubit_operator, self.n_qubits)
                group_x, mask1, mask2, flags = support.group_conditions(x, z, c)
                flags = support.conditions_implied_by_sector(group_x, mask1, mask2, flags,
                                                             tables.plus, tables.minus)
                if numpy.any(flags):
                    xr, zr, cr = support.reduce_terms(x, z, c, group_x, mask1, mask2, flags)
                    self._apply_plan = PauliPlan(self.n_qubits, xr, zr, cr, hints=False)
        return self._apply_plan

    @property
    def sector(self):
        """The device-resident index map (built on first access; raises without a GPU)."""
        if self._sector is None:
            if self.sz_value is None:
                self._sector = sectors.Sector.number(self.n_qubits, self.n_electrons)
            else:
                self._sector = sectors.Sector.sz(self.n_qubits, self.sz_value, self.n_electrons,
                                                 *self._spin_maps)
        return self._sector

    def indices(self):
        """Basis state at every position: jw_number_indices / jw_sz_indices as an array."""
        return self.sector.basis()

    def diagonal(self):
        return self.sector.diagonal_host(self.plan)

    def _matvec(self, x):
        arr = numpy.asarray(x)
        vec = numpy.ascontiguousarray(arr.reshape(-1), dtype=numpy.complex128)
        return self.sector.apply_host(self.apply_plan, vec).reshape(arr.shape)

    def _matmat(self, X):
        X = numpy.asarray(X)
        return self.sector.apply_host(self.apply_plan, numpy.asfortranarray(X, dtype=numpy.complex128))


def apply_operator(args):
    """Helper kept for API compatibility (linear_qubit_operator.py:212-215)."""
    operator, vec = args
    return operator * vec


def generate_linear_qubit_operator(qubit_operator, n_qubits=None, options=None):
    """LinearQubitOperator when options is None, else the parallel variant
    (linear_qubit_operator.py:218-234)."""
    if options is None:
        return LinearQubitOperator(qubit_operator, n_qubits)
    return ParallelLinearQubitOperator(qubit_operator, n_qubits, options)
