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


def cpu_baseline_sample(op, n, seconds_budget=30.0):
    """Numpy port of the reference matvec (1 thread, like LinearQubitOperator) on a bounded
    sample: the first few terms of the same table on the same-size state."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import pauli_oracle
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    n_cpu = n
    while (1 << n_cpu) * 16 * 8 > avail and n_cpu > 16:
        n_cpu -= 2
    items = [(k, v) for k, v in op.terms.items() if k and all(q < n_cpu for q, _ in k)]
    rng = numpy.random.default_rng(0)
    vec = rng.normal(size=1 << n_cpu) + 1j * rng.normal(size=1 << n_cpu)
    done, elapsed = 0, 0.0
    per_term = None
    while done < len(items):
        take = 1 if per_term is None else max(1, min(8, int((seconds_budget - elapsed) / per_term)))
        if per_term is not None and elapsed + per_term > seconds_budget:
            break
        chunk = dict(items[done:done + take])
        t0 = time.perf_counter()
        pauli_oracle.matvec(chunk, n_cpu, vec)
        dt = time.perf_counter() - t0
        elapsed += dt
        done += len(chunk)
        per_term = elapsed / done
    value = done * float(1 << n_cpu) / elapsed
    return {'value': value, 'unit': UNIT, 'cores': 1, 'kind': 'port',
            'sample': '%d terms of the workload table on a 2^%d state, numpy port of '
                      'LinearQubitOperator._matvec (oracle/pauli_oracle.py), %.1f s' % (done, n_cpu, elapsed)}


def run_reference(args):
    """Reference arm: CPU restatement with term groups in worker processes
    (ParallelLinearQubitOperator semantics), bounded sample per step."""
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
    # bounded sample: the numpy port costs ~3e-8 s per term*amplitude and process, so a 2^26
    # state keeps one step at a few seconds (the per-term*amplitude cost does not depend on n
    # beyond cache sizes: 1.5e7-3e7 at n = 20..28)
    n_cpu = min(n, args.reference_qubits)
    procs = max(1, min(4, cores))
    while (1 << n_cpu) * 16 * 8 * procs > avail and n_cpu > 16:
        n_cpu -= 2
    items = [(k, v) for k, v in op.terms.items() if k and all(q < n_cpu for q, _ in k)][:procs]
    sample = dict(items)
    rng = numpy.random.default_rng(0)
    vec = rng.normal(size=1 << n_cpu) + 1j * rng.normal(size=1 << n_cpu)
    # the reference offers two ways to run this path: LinearQubitOperator (one thread) and
    # ParallelLinearQubitOperator (term groups in worker processes, full vector pickled to and
    # from every worker); time one step of each and keep the faster for the timed steps
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
    desc = ('%d terms of the workload table per step on a 2^%d state, numpy port of the reference '
            '(oracle/pauli_oracle.py), %s (faster of one thread %.2f s / %d workers %.2f s per step)'
            % (len(sample), n_cpu, best, timings['one thread'], procs,
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

    # sanity of the timed result: <v|Hv> through the vector kernels equals the fused K2 value
    e_sep = krylov.dotc(dim, d_in.ptr, d_out.ptr)
    e_fused = plan.expectation_device(d_in.ptr)
    check = abs(e_sep - e_fused) / max(abs(e_fused), 1e-300)

    # e2e: C-ABI call with pinned HOST buffers, H2D + kernel + D2H inside the timed region
    h_in, h_out = vp(), vp()
    _lib.call('ofb_host_alloc', ctypes.byref(h_in), ctypes.c_size_t(dim * 16))
    _lib.call('ofb_host_alloc', ctypes.byref(h_out), ctypes.c_size_t(dim * 16))
    _lib.call('ofb_memcpy_d2h', h_in, vp(d_in.ptr), ctypes.c_size_t(dim * 16), vp(None))
    _lib.call('ofb_device_sync')
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    _lib.call('ofb_apply_host', plan.handle, h_in, h_out, 1)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _lib.call('ofb_apply_host', plan.handle, h_in, h_out, 1)
    e2e_seconds = time.perf_counter() - t0
    e2e_value = T * float(dim) * e2e_steps / e2e_seconds
    host_view = numpy.ctypeslib.as_array(ctypes.cast(h_out, ctypes.POINTER(ctypes.c_double)), (4,))
    dev_head = krylov.download(2, d_out.ptr).view(numpy.float64)
    assert numpy.array_equal(host_view, dev_head), 'e2e result differs from device-resident result'
    _lib.call('ofb_host_free', h_in)
    _lib.call('ofb_host_free', h_out)

    peak, peak_src = measured_peak()
    alg_bytes = 16.0 * dim * (G + 1)
    achieved = alg_bytes / (seconds / args.steps) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'k_apply_traffic.json')
    if os.path.exists(tpath):
        try:
            rec = json.load(open(tpath))
            if rec.get('n_qubits') == n:
                traffic = rec.get('dram_bytes_per_launch')
        except Exception:
            pass
    cpu = None if args.no_cpu_baseline else cpu_baseline_sample(op, n)
    # same-workload base for the multi-GPU lines (Hubbard 4x4, 32 qubits on ONE GPU)
    aux = None
    if args.aux_hubbard32 and n == 28:
        try:
            d_in.free()
            d_out.free()
            plan.close()
            from openfermion_b200 import hamiltonians
            hub = LinearQubitOperator(hamiltonians.hubbard_jw(4, 4), 32)
            hdim = 1 << 32
            h_in, h_out = _lib.DeviceBuffer(hdim * 16), _lib.DeviceBuffer(hdim * 16)
            krylov.fill_random(hdim, 7, False, h_in.ptr)
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
            h_in.free()
            h_out.free()
            # the same Hamiltonian restricted to its half-filled S_z = 0 sector (K7)
            from openfermion_b200.linalg import SectorLinearQubitOperator
            sec = SectorLinearQubitOperator(hub, n_electrons=16, sz_value=0.0)
            sdim = sec.shape[0]
            s_in, s_out = _lib.DeviceBuffer(sdim * 16), _lib.DeviceBuffer(sdim * 16)
            krylov.fill_random(sdim, 9, False, s_in.ptr)
            for _ in range(2):
                sec.sector.apply_device(hub.plan, s_in.ptr, s_out.ptr)
            _lib.call('ofb_event_record', ev[0], vp(None))
            for _ in range(3):
                sec.sector.apply_device(hub.plan, s_in.ptr, s_out.ptr)
            _lib.call('ofb_event_record', ev[1], vp(None))
            _lib.call('ofb_event_sync', ev[1])
            _lib.call('ofb_event_elapsed_ms', ev[0], ev[1], ctypes.byref(hms))
            aux['sector'] = {'workload': 'same Hamiltonian inside its (16 electrons, S_z = 0) sector: '
                                         '%d of 2^32 amplitudes, ONE GPU' % sdim,
                             'ms_per_step': hms.value / 3,
                             'sector_term_amp_per_s': hub.plan.n_terms * float(sdim) * 3 / (hms.value / 1e3)}
            s_in.free()
            s_out.free()
        except Exception as exc:  # e.g. not enough memory on a shared box
            aux = {'error': str(exc)[:200]}
    dram = (traffic / (seconds / args.steps) / 1e9) if traffic else None
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms.value / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': name, 'n_qubits': n, 'terms': T, 'x_groups': G,
                   'support_hints': os.environ.get('OFB_SUPPORT_HINTS'),  # experimental, None = off
                   'l2_policy': 'inputs exceed L2 (state vector %.1f GB per buffer)' % (dim * 16 / 1e9),
                   'timing': 'CUDA events on the launch stream', 'self_check_rel_err': check},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': dim * 16,
                'd2h_bytes_per_step': dim * 16, 'steps': e2e_steps,
                'path': 'ofb_apply_host (C ABI) with pinned host buffers'},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'kernel': 'k_apply', 'algorithmic_bytes_per_launch': alg_bytes,
                     'frac_of_8TBs_nominal': achieved / 8000.0,
                     'dram_achieved': dram, 'dram_frac': (dram / peak) if dram else None,
                     'note': 'x-groups run in ascending x, so the input tile is re-read through L1/L2: '
                             'real DRAM traffic (traffic, ncu) is a fraction of the algorithmic bytes and '
                             'frac exceeds 1; the binding limits are SM issue rate and the L1 data pipe '
                             '(DESIGN.md section 4, profiles/README.md)'},
        'cpu_baseline': cpu,
        'aux': aux,
    }
    print(json.dumps(line))


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
    ap.add_argument('--reference-qubits', type=int, default=26,
                    help='state size of the bounded sample of the --impl reference arm')
    ap.add_argument('--dist-qubits', type=int, default=32, help='N>1 workload size (Hubbard)')
    ap.add_argument('--dist-mode', default='ce', choices=['ce', 'nccl', 'p2p'])
    ap.add_argument('--dist-basis', default='symmetry', choices=['symmetry', 'natural'],
                    help='N>1: shard by conserved parities first (no exchange for them) or by the top qubits')
    ap.add_argument('--dist-order', default='locality', choices=['locality', 'natural'],
                    help='N>1 with --dist-basis symmetry: order of the qubits inside a shard')
    ap.add_argument('--dist-diagnose', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    import __graft_entry__
    __graft_entry__.build()
    if args.gpus == 1 and int(os.environ.get('WORLD_SIZE', '1')) == 1:
        return run_single(args)
    from openfermion_b200 import distributed
    return distributed.bench_main(args, METRIC, UNIT, ClockSampler, measured_peak)


if __name__ == '__main__':
    main()
