#!/usr/bin/env python
"""Benchmark of the hot path: matrix-free H|psi> (K1) on the BASELINE.json workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one H|psi> over the whole state.  N = 1: synthetic molecular-type Hamiltonian,
28 qubits (BASELINE.json configs[3]); N > 1: Fermi-Hubbard 4x4 (32 qubits, configs[4]) with
the state sharded by its high qubits, launched under torchrun (one rank per GPU).
Prints ONE JSON line (rank 0).  metric = term*amplitude updates/s = T * 2^n * K / time.
`--impl reference` times the CPU restatement of the reference's own path (oracle/, numpy,
term groups in worker processes like ParallelLinearQubitOperator) on a bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'H|psi> term-amplitude updates/s'
UNIT = 'term*amp/s'


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(numpy.median(sm)) if sm else None,
                'sm_max_mhz': max(smax) if smax else None, 'samples': len(sm),
                'reasons': sorted(reasons)}


def workload_single(qubits):
    from openfermion_b200 import hamiltonians
    if qubits % 2:
        raise SystemExit('--qubits must be even (spin orbitals)')
    op = hamiltonians.molecular_jw(qubits // 2, seed=1234)
    name = 'synthetic molecular-type JW Hamiltonian, %d qubits (random_interaction_operator(%d, expand_spin, seed=1234))' % (qubits, qubits // 2)
    return op, name


def reference_sample(op, n_cpu, count, seed=0):
    """A bounded, representative sample of the workload's term table for the CPU arms: `count`
    Pauli strings drawn (seeded) from the terms that act inside the first n_cpu qubits, diagonal
    (x = 0) and flipping (x != 0) strings in the table's own proportion (at least one diagonal
    string, so both code paths of the reference matvec are timed)."""
    rng = numpy.random.default_rng(seed)
    diag, flip = [], []
    n_diag_all = 0
    for key, value in op.terms.items():
        is_diag = all(action == 'Z' for _, action in key)
        n_diag_all += is_diag
        if key and all(q < n_cpu for q, _ in key):
            (diag if is_diag else flip).append((key, value))
    want_diag = min(len(diag), max(1, int(round(count * n_diag_all / max(len(op.terms), 1)))))
    want_flip = min(len(flip), count - want_diag)
    pick = [diag[i] for i in rng.choice(len(diag), want_diag, replace=False)] if want_diag else []
    pick += [flip[i] for i in rng.choice(len(flip), want_flip, replace=False)] if want_flip else []
    return dict(pick), want_diag, want_flip


def cpu_baseline_sample(op, n, n_cpu=22, count=32):
    """Numpy port of the reference matvec (1 thread, like LinearQubitOperator) on a bounded
    sample: 32 representative strings of the same table on a 2^22 state (10-30 s of CPU work;
    the port's cost per term*amplitude does not depend on n beyond cache sizes)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pauli_oracle
    n_cpu = min(n, n_cpu)
    sample, n_diag, n_flip = reference_sample(op, n_cpu, count)
    rng = numpy.random.default_rng(0)
    vec = rng.normal(size=1 << n_cpu) + 1j * rng.normal(size=1 << n_cpu)
    pauli_oracle.matvec(dict(list(sample.items())[:2]), n_cpu, vec)  # warm-up
    t0 = time.perf_counter()
    reps = 0
    while reps < 1 or time.perf_counter() - t0 < 10.0:
        pauli_oracle.matvec(sample, n_cpu, vec)
        reps += 1
    elapsed = time.perf_counter() - t0
    value = reps * len(sample) * float(1 << n_cpu) / elapsed
    return {'value': value, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': '%d strings of the workload table (%d diagonal, %d flipping, seeded draw in the '
                      "table's proportion) on a 2^%d state, %d passes, numpy port of "
                      'LinearQubitOperator._matvec (oracle/pauli_oracle.py), %.1f s'
                      % (len(sample), n_diag, n_flip, n_cpu, reps, elapsed)}


def run_reference(args):
    """Reference arm: the CPU restatement of the reference's own path (LinearQubitOperator: one
    thread; ParallelLinearQubitOperator: term groups in worker processes, the whole vector
    pickled to and from each) on a bounded, representative sample per step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pauli_oracle
    from openfermion_b200 import hamiltonians
    if args.gpus == 1:
        op, name = workload_single(args.qubits)
    else:
        op, name = hamiltonians.hubbard_jw(4, 4), 'Fermi-Hubbard 4x4 spinful JW, 32 qubits'
    n = op.n_qubits
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    # bounded sample: the numpy port costs ~3e-8 s per term*amplitude and process (1.5e7-3e7
    # term*amp/s at n = 20..28), so 32 strings on a 2^22 state keep one
    # step at a few seconds and a --steps 20 run within minutes
    n_cpu = min(n, args.reference_qubits)
    procs = max(1, min(10, cores))  # LinearQubitOperatorOptions(processes=10) capped by the cores
    while (1 << n_cpu) * 16 * 8 * procs > avail and n_cpu > 16:
        n_cpu -= 2
    sample, n_diag, n_flip = reference_sample(op, n_cpu, args.reference_terms)
    rng = numpy.random.default_rng(0)
    vec = rng.normal(size=1 << n_cpu) + 1j * rng.normal(size=1 << n_cpu)

    def one_thread():
        return pauli_oracle.matvec(sample, n_cpu, vec)

    def workers():
        return pauli_oracle.matvec_parallel(sample, n_cpu, vec, procs)

    timings = {}
    for label, fn in (('one thread', one_thread), ('%d worker processes' % procs, workers)):
        t0 = time.perf_counter()
        fn()
        timings[label] = time.perf_counter() - t0
    best = min(timings, key=timings.get)
    step_fn = one_thread if best == 'one thread' else workers
    used = 1 if best == 'one thread' else procs
    for _ in range(max(args.warmup - 1, 0)):
        step_fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_fn()
    elapsed = time.perf_counter() - t0
    value = len(sample) * float(1 << n_cpu) * args.steps / elapsed
    desc = ('%d strings of the workload table per step (%d diagonal, %d flipping: seeded draw in the '
            "table's proportion) on a 2^%d state, numpy port of the reference (oracle/pauli_oracle.py), "
            '%s (faster of one thread %.2f s / %d workers %.2f s per step); rate per term*amplitude '
            'extrapolates linearly to the full table and state'
            % (len(sample), n_diag, n_flip, n_cpu, best, timings['one thread'], procs,
               timings['%d worker processes' % procs]))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * elapsed / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': {'workload': name, 'sample': desc},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': used, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def ncu_fractions(n):
    """Hardware fractions of the dominant kernel from the committed ncu capture of this workload
    (profiles/k_apply_traffic.json, written by scripts/ncu_summary.py; says which commit)."""
    path = os.path.join(ROOT, 'profiles', 'k_apply_traffic.json')
    try:
        rec = json.load(open(path))
        if rec.get('n_qubits') == n:
            return rec
    except Exception:
        pass
    return {}


def oracle_check(op, n, h_in, h_out, samples=256, seed=11):
    """The timed result against the CPU oracle: (H psi)[j] at `samples` seeded output indices,
    evaluated from the reference-order term dict (oracle/pauli_oracle.py:matvec_sampled, not the
    compiled plan).  Returns the max error relative to max |H psi| over the sample."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pauli_oracle
    rng = numpy.random.default_rng(seed)
    idx = numpy.unique(rng.integers(0, 1 << n, size=samples, dtype=numpy.int64))
    want = pauli_oracle.matvec_sampled(op.terms, n, h_in, idx)
    got = h_out[idx]
    return float(numpy.max(numpy.abs(got - want)) / max(float(numpy.max(numpy.abs(want))), 1e-300)), len(idx)


def run_single(args):
    from openfermion_b200 import _lib
    from openfermion_b200.linalg import LinearQubitOperator
    from openfermion_b200.linalg import krylov
    _lib.require_device()
    op, name = workload_single(args.qubits)
    n = op.n_qubits
    dim = 1 << n
    lin = LinearQubitOperator(op, n)
    plan = lin.plan
    T, G = plan.n_terms, plan.n_groups
    vp = ctypes.c_void_p

    d_in = _lib.DeviceBuffer(dim * 16)
    d_out = _lib.DeviceBuffer(dim * 16)
    krylov.fill_random(dim, 42, False, d_in.ptr)
    krylov.scal(dim, 1.0 / krylov.nrm2(dim, d_in.ptr), d_in.ptr)

    ev = [vp() for _ in range(2)]
    for e in ev:
        _lib.call('ofb_event_create', ctypes.byref(e))

    for _ in range(args.warmup):
        plan.apply_device(d_in.ptr, d_out.ptr)
    _lib.call('ofb_device_sync')
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    launches0 = _lib.launch_count()
    _lib.call('ofb_event_record', ev[0], vp(None))
    for _ in range(args.steps):
        plan.apply_device(d_in.ptr, d_out.ptr)
    _lib.call('ofb_event_record', ev[1], vp(None))
    _lib.call('ofb_event_sync', ev[1])
    _lib.call('ofb_device_sync')
    launches = _lib.launch_count() - launches0
    ms = ctypes.c_float()
    _lib.call('ofb_event_elapsed_ms', ev[0], ev[1], ctypes.byref(ms))
    clocks = sampler.stop()
    seconds = ms.value / 1e3
    value = T * float(dim) * args.steps / seconds

    # e2e: the drop-in call a user makes -- LinearQubitOperator.matvec on a numpy array in
    # ordinary (pageable) host memory; H2D, kernel and D2H all inside the timed region
    h_in = krylov.download(dim, d_in.ptr)
    dev_out = krylov.download(dim, d_out.ptr)
    d_in.free()
    d_out.free()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_out = lin.matvec(h_in)  # warm-up: staging buffers, page faults of the first result
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h_out = lin.matvec(h_in)
    e2e_seconds = time.perf_counter() - t0
    e2e_value = T * float(dim) * e2e_steps / e2e_seconds
    assert numpy.array_equal(h_out, dev_out), 'e2e result differs from the device-resident result'
    del dev_out

    # parity of the timed result: sampled outputs against the CPU oracle (reference term order)
    check, n_checked = oracle_check(op, n, h_in, h_out, args.check_samples)
    assert check <= 1e-12, 'H|psi> differs from the oracle: rel err %.3e' % check
    e_fused = plan.expectation_host(h_in)
    e_host = numpy.vdot(h_in, h_out)
    e_check = abs(e_fused - e_host) / max(abs(e_host), 1e-300)
    del h_out

    peak, peak_src = measured_peak()
    alg_bytes = 16.0 * dim * (G + 1)
    per_launch = seconds / args.steps
    achieved = alg_bytes / per_launch / 1e9
    ncu = ncu_fractions(n)
    traffic = ncu.get('dram_bytes_per_launch')
    fracs = {'algorithmic_frac': achieved / peak}
    if traffic:
        fracs['dram_frac'] = traffic / per_launch / 1e9 / peak
    for key in ('l1_frac', 'issue_frac', 'fp64_frac', 'alu_frac', 'lsu_frac'):
        if ncu.get(key) is not None:
            fracs[key] = ncu[key]
    hw = {k: v for k, v in fracs.items() if k != 'algorithmic_frac'}
    bound = max(hw, key=hw.get).replace('_frac', '') if hw else 'hbm'
    cpu = None if args.no_cpu_baseline else cpu_baseline_sample(op, n)
    # same-workload base for the multi-GPU lines (Hubbard 4x4, 32 qubits on ONE GPU)
    aux = None
    if args.aux_hubbard32 and n == 28:
        try:
            del h_in
            plan.close()
            aux = aux_hubbard32(_lib, krylov, ev)
        except Exception as exc:  # e.g. not enough memory on a shared box
            aux = {'error': str(exc)[:200]}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms.value / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': name, 'n_qubits': n, 'terms': T, 'x_groups': G,
                   'applied_terms': getattr(plan, 'n_reduced_terms', T),
                   'support_hints': os.environ.get('OFB_SUPPORT_HINTS', '1'),
                   'support_fraction': plan.support_fraction,
                   'l2_policy': 'inputs exceed L2 (state vector %.1f GB per buffer)' % (dim * 16 / 1e9),
                   'timing': 'CUDA events on the launch stream',
                   'self_check_rel_err': check,
                   'self_check': 'max |(H psi)[j] - oracle| / max |oracle| over %d seeded output indices, '
                                 'oracle = reference-order term sum (oracle/pauli_oracle.py:matvec_sampled); '
                                 'bar 1e-12' % n_checked,
                   'expectation_check_rel_err': e_check},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': dim * 16,
                'd2h_bytes_per_step': dim * 16, 'steps': e2e_steps, 'ms_per_step': 1e3 * e2e_seconds / e2e_steps,
                'path': 'LinearQubitOperator.matvec(numpy array in pageable host memory): staged H2D in 8 slabs '
                        '(slab-local x-groups start behind each slab), remaining groups, staged D2H of output '
                        'slabs behind the kernel (ofb_apply_host, csrc/hostcopy.cu)'},
        'gpu_launches': int(launches),
        'roofline': {'bound': bound, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'kernel': 'k_apply_tile', 'algorithmic_bytes_per_launch': alg_bytes,
                     'frac_of_8TBs_nominal': achieved / 8000.0,
                     'hardware_fractions': hw, 'ncu_capture': ncu.get('source'),
                     'ncu_commit': ncu.get('commit'),
                     'note': 'frac follows SURVEY 8(d) (one HBM stream per x-group); x-groups run in ascending x, '
                             'so the input tile is re-read through L1/L2 and frac exceeds 1.  The hardware '
                             'fractions (DRAM bytes live / peak; L1, issue, FP64 from the ncu capture named in '
                             'ncu_capture) say what binds: `bound` is the largest of them'},
        'cpu_baseline': cpu,
        'aux': aux,
    }
    print(json.dumps(line))


def aux_hubbard32(_lib, krylov, ev):
    """One-GPU numbers of the multi-GPU workload (Hubbard 4x4, 32 qubits): the strong-scaling
    base of the --gpus N lines, and the same Hamiltonian inside its half-filled sector."""
    from openfermion_b200 import hamiltonians
    from openfermion_b200.linalg import LinearQubitOperator, SectorLinearQubitOperator
    vp = ctypes.c_void_p
    hub = LinearQubitOperator(hamiltonians.hubbard_jw(4, 4), 32)
    hdim = 1 << 32
    h_in, h_out = _lib.DeviceBuffer(hdim * 16), _lib.DeviceBuffer(hdim * 16)
    # the same global random state (seed 42, normalised) the --gpus N lines apply H to
    krylov.fill_random(hdim, 42, False, h_in.ptr)
    krylov.scal(hdim, 1.0 / krylov.nrm2(hdim, h_in.ptr), h_in.ptr)
    for _ in range(2):
        hub.plan.apply_device(h_in.ptr, h_out.ptr)
    _lib.call('ofb_event_record', ev[0], vp(None))
    for _ in range(3):
        hub.plan.apply_device(h_in.ptr, h_out.ptr)
    _lib.call('ofb_event_record', ev[1], vp(None))
    _lib.call('ofb_event_sync', ev[1])
    hms = ctypes.c_float()
    _lib.call('ofb_event_elapsed_ms', ev[0], ev[1], ctypes.byref(hms))
    aux = {'workload': 'Fermi-Hubbard 4x4 spinful JW, 32 qubits, ONE GPU (strong-scaling base of the --gpus N lines)',
           'ms_per_step': hms.value / 3, 'value': hub.plan.n_terms * float(hdim) * 3 / (hms.value / 1e3),
           'unit': UNIT, 'terms': hub.plan.n_terms, 'x_groups': hub.plan.n_groups}
    e = krylov.dotc(hdim, h_in.ptr, h_out.ptr)
    aux['expectation_value'] = [e.real, e.imag]  # compare with config.expectation_value of the N > 1 lines
    h_in.free()
    h_out.free()
    sec = SectorLinearQubitOperator(hub, n_electrons=16, sz_value=0.0)
    sdim = sec.shape[0]
    s_in, s_out = _lib.DeviceBuffer(sdim * 16), _lib.DeviceBuffer(sdim * 16)
    krylov.fill_random(sdim, 9, False, s_in.ptr)
    for _ in range(2):
        sec.sector.apply_device(sec.apply_plan, s_in.ptr, s_out.ptr)
    _lib.call('ofb_event_record', ev[0], vp(None))
    for _ in range(3):
        sec.sector.apply_device(sec.apply_plan, s_in.ptr, s_out.ptr)
    _lib.call('ofb_event_record', ev[1], vp(None))
    _lib.call('ofb_event_sync', ev[1])
    _lib.call('ofb_event_elapsed_ms', ev[0], ev[1], ctypes.byref(hms))
    aux['sector'] = {'workload': 'same Hamiltonian inside its (16 electrons, S_z = 0) sector: '
                                 '%d of 2^32 amplitudes, ONE GPU' % sdim,
                     'ms_per_step': hms.value / 3,
                     'sector_term_amp_per_s': hub.plan.n_terms * float(sdim) * 3 / (hms.value / 1e3)}
    s_in.free()
    s_out.free()
    return aux


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--qubits', type=int, default=28, help='N=1 workload size (even)')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-aux', dest='aux_hubbard32', action='store_false',
                    help='skip the one-GPU 32-qubit Hubbard measurement appended to the N=1 line')
    ap.add_argument('--check-samples', type=int, default=256,
                    help='output indices checked against the CPU oracle after the timed loop')
    ap.add_argument('--reference-terms', type=int, default=32,
                    help='strings per step of the --impl reference arm (stratified sample)')
    ap.add_argument('--reference-qubits', type=int, default=22,
                    help='state size of the bounded sample of the --impl reference arm')
    ap.add_argument('--dist-qubits', type=int, default=32, help='N>1 workload size (Hubbard)')
    ap.add_argument('--dist-mode', default='ce', choices=['ce', 'nccl', 'p2p'])
    ap.add_argument('--dist-basis', default='symmetry', choices=['symmetry', 'natural'],
                    help='N>1: shard by conserved parities first (no exchange for them) or by the top qubits')
    ap.add_argument('--dist-order', default='locality', choices=['locality', 'natural'],
                    help='N>1 with --dist-basis symmetry: order of the qubits inside a shard')
    ap.add_argument('--dist-diagnose', action='store_true')
    ap.add_argument('--lanczos-steps', type=int, default=20,
                    help='N>1: Lanczos iterations timed after the H|psi> steps (0 = skip)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    import __graft_entry__
    __graft_entry__.build_library()  # the product arm never builds or loads oracle/ code here
    if args.gpus == 1 and int(os.environ.get('WORLD_SIZE', '1')) == 1:
        return run_single(args)
    from openfermion_b200 import distributed
    return distributed.bench_main(args, METRIC, UNIT, ClockSampler, measured_peak, sharded_output_check)


def sharded_output_check(op, n, natural_indices, amplitude_fn, got):
    """Checker handed to the N > 1 bench (the product package never imports oracle/): sampled
    outputs of the sharded H|psi> against the CPU oracle's reference-order term sum; the input
    amplitudes come from the state's counter-based generator, so no 2^n vector exists on the host."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pauli_oracle
    want = pauli_oracle.matvec_sampled(op.terms, n, amplitude_fn, natural_indices)
    return float(numpy.max(numpy.abs(numpy.asarray(got) - want)) / max(float(numpy.max(numpy.abs(want))), 1e-300))


if __name__ == '__main__':
    main()
