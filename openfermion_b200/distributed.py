"""State vectors sharded by their high qubits across GPUs (one process per GPU).

Rank r of W = 2^k owns amplitudes [r * 2^(n-k), (r+1) * 2^(n-k)): the top k index bits,
i.e. OpenFermion qubits 0..k-1 (qubit q is index bit n-1-q,
linalg/linear_qubit_operator.py:126).  For an x-group with mask x = (x_hi, x_lo):

    out_r[j] += S_g((r << m) | j) * in_{r ^ x_hi}[j ^ x_lo]        m = n - k

so a group is local when x_hi == 0 and otherwise needs the shard of the single peer
r ^ x_hi -- a perfect matching of the ranks for every distinct non-zero x_hi ("pattern").
Because the plan orders groups by ascending x, the groups of one pattern are a contiguous
range.  Z bits on the high qubits only enter through the global index in S_g.

Three data paths, same kernel (`ofb_apply_groups`):
  * 'ce'    each rank PULLS the peer shards it needs into staging buffers with
            device-to-device copies on a separate stream (copy engines over NVLink through
            CUDA-IPC-mapped peer pointers: no SM is spent on communication) while the local
            groups run; the kernels of a pattern wait on that copy's event;
  * 'nccl'  pair exchange of whole shards with torch.distributed send/recv into staging
            buffers, posted ahead and overlapped with the local groups;
  * 'p2p'   no exchange at all: the kernel's input pointer is the peer's shard, mapped
            through CUDA IPC, and the gather loads travel over NVLink while the SMs work
            (compute and communication in one kernel).
The reference has no counterpart (its only parallelism is term groups in worker processes
over a replicated vector, linear_qubit_operator.py:176-209).

The host logic (pattern table, peers, staging schedule, reductions) is independent of the
device: tests drive it on CPU tensors over gloo with an injected local-apply function.
"""
import ctypes
import json
import os
import time

import numpy

from openfermion_b200.compiler import compile_qubit_operator
from openfermion_b200.ops import count_qubits
from openfermion_b200.shard_basis import ShardBasis


# ----------------------------------------------------------------------------
# host logic
# ----------------------------------------------------------------------------
class ShardLayout:
    """Which x-group ranges need which peer shard."""

    def __init__(self, n_qubits, world, x_mask):
        if world < 1 or world & (world - 1):
            raise ValueError('world size must be a power of two, got {}'.format(world))
        self.n_qubits = n_qubits
        self.world = world
        self.shard_bits = world.bit_length() - 1
        if self.shard_bits > n_qubits:
            raise ValueError('more ranks than amplitudes')
        self.local_bits = n_qubits - self.shard_bits
        self.local_dim = 1 << self.local_bits
        group_x = numpy.unique(numpy.asarray(x_mask, dtype=numpy.uint64))  # ascending = plan order
        self.group_x = group_x
        hi = (group_x >> numpy.uint64(self.local_bits)).astype(numpy.int64)
        self.patterns = []  # (x_hi, g_begin, g_end), ascending x_hi, pattern 0 (local) first
        start = 0
        for g in range(1, len(hi) + 1):
            if g == len(hi) or hi[g] != hi[start]:
                self.patterns.append((int(hi[start]), start, g))
                start = g
        self.remote_patterns = [p for p in self.patterns if p[0] != 0]
        self.local_pattern = next((p for p in self.patterns if p[0] == 0), None)

    def peer(self, rank, x_hi):
        return rank ^ x_hi

    def index_base(self, rank):
        return rank << self.local_bits

    def exchange_bytes_per_matvec(self):
        """Bytes each rank receives per matvec on the staged path."""
        return len(self.remote_patterns) * self.local_dim * 16


def _splitmix64(z):
    z = z + numpy.uint64(0x9e3779b97f4a7c15)
    z = (z ^ (z >> numpy.uint64(30))) * numpy.uint64(0xbf58476d1ce4e5b9)
    z = (z ^ (z >> numpy.uint64(27))) * numpy.uint64(0x94d049bb133111eb)
    return z ^ (z >> numpy.uint64(31))


def global_random_amplitudes(seed, natural_indices, real_only=False):
    """Host mirror of the device generator (csrc/vec.cu k_fill_random, csrc/relabel.cu): the
    unnormalised amplitudes of the seeded global random state at the given natural indices."""
    g = numpy.asarray(natural_indices, dtype=numpy.uint64)
    with numpy.errstate(over='ignore'):
        s = numpy.uint64(seed)
        a = _splitmix64(s ^ _splitmix64(numpy.uint64(2) * g))
        b = _splitmix64(s ^ _splitmix64(numpy.uint64(2) * g + numpy.uint64(1)))
    scale = 1.0 / 9007199254740992.0
    re = (a >> numpy.uint64(11)).astype(numpy.float64) * scale
    im = numpy.zeros_like(re) if real_only else (b >> numpy.uint64(11)).astype(numpy.float64) * scale
    return re + 1j * im


class TorchComm:
    """All-reduce hook for the Krylov loops over a torch.distributed group."""

    def __init__(self, dist, device, group=None):
        self.dist = dist
        self.group = group
        self.device = device
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def allsum(self, value):
        import torch
        value = complex(value)
        t = torch.tensor([value.real, value.imag], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, group=self.group)
        out = t.cpu()
        return complex(float(out[0]), float(out[1]))

    def barrier(self):
        # an all-reduce orders the streams of all ranks without an extra host round trip
        self.allsum(0.0)


class ShardedPauliOperator:
    """Matrix-free H|psi> on a sharded state.

    local_apply(in_handle, out_handle, g_begin, g_end, accumulate) applies a group range to
    this rank's output shard; the default launches the CUDA kernel.  `vectors` are opaque
    handles created by `new_vector()` (device buffers; CPU arrays in the gloo tests).
    """

    def __init__(self, operator, n_qubits=None, dist=None, group=None, mode='nccl',
                 max_staging=None, local_apply=None, vector_factory=None, tensor_of=None,
                 basis='natural', local_order='natural'):
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist is not None else 0
        self.world = dist.get_world_size(group) if dist is not None else 1
        self.mode = mode
        needed = count_qubits(operator)
        self.n_qubits = needed if n_qubits is None else n_qubits
        if self.n_qubits < needed:
            raise ValueError('Invalid number of qubits specified.')
        x, z, c = compile_qubit_operator(operator, self.n_qubits)
        # basis: 'natural' = shard by the top qubits; 'symmetry' = shard by conserved parities
        # first (shard_basis.py), which removes peer exchanges; or a ShardBasis instance.
        # Shards then hold psi'[c'] = psi[B^-1 c']: see shard_of / natural_indices.
        # local_order='locality' (with basis='symmetry') also permutes the qubits inside a
        # shard so that more x-groups gather from the same cache-resident blocks.
        shard_bits = max(self.world, 1).bit_length() - 1
        if basis == 'natural':
            if local_order != 'natural':
                raise ValueError("local_order='locality' needs basis='symmetry'")
            basis = ShardBasis.natural(self.n_qubits)
        elif basis == 'symmetry':
            basis = ShardBasis.for_sharding(self.n_qubits, min(shard_bits, self.n_qubits), x,
                                            local_order=local_order)
        elif not isinstance(basis, ShardBasis):
            raise ValueError("basis must be 'natural', 'symmetry' or a ShardBasis")
        if basis.n_qubits != self.n_qubits:
            raise ValueError('shard basis is for {} qubits, operator has {}'.format(
                basis.n_qubits, self.n_qubits))
        self.basis = basis
        self._natural_x = x
        x, z, c = basis.transform_terms(x, z, c)
        self.layout = ShardLayout(self.n_qubits, self.world, x)
        self.n_terms = len(c)
        self.n_groups = len(self.layout.group_x)
        self.dim = self.layout.local_dim  # local length: what the Krylov loops see
        self._cuda = local_apply is None
        self._tables = (x, z, c)
        self.plan = None
        if self._cuda:
            from openfermion_b200 import _lib
            from openfermion_b200.plan import PauliPlan
            self._lib = _lib
            device = self._device_index()
            _lib.call('ofb_set_device', ctypes.c_int(device))
            self.plan = PauliPlan(self.n_qubits, x, z, c, device=device)
            self._local_apply = self._cuda_apply
            self._new = self._cuda_vector
            self._tensor_of = self._cuda_tensor
        else:
            self._local_apply = local_apply
            self._new = vector_factory
            self._tensor_of = tensor_of
        n_remote = len(self.layout.remote_patterns)
        want = n_remote if max_staging is None else min(max_staging, n_remote)
        self._staging = [self._new() for _ in range(want)] if mode in ('nccl', 'ce') else []
        self._copy_stream = None
        self._copy_lanes = max(1, int(os.environ.get('OFB_CE_LANES', '1')))
        while self._copy_lanes > 1 and (self.dim * 16) % (self._copy_lanes * 256):
            self._copy_lanes //= 2
        self._events = []
        self._peer_vectors = {}  # p2p: vector id -> list of per-rank device pointers
        self._opened = []

    # -- CUDA plumbing ---------------------------------------------------------------
    def _device_index(self):
        return int(os.environ.get('LOCAL_RANK', self.rank))

    def _cuda_vector(self):
        return self._lib.DeviceBuffer(self.dim * 16)

    def _cuda_tensor(self, vec):
        """float64 torch view of a raw device allocation (for NCCL send/recv)."""
        import torch

        class _Iface:
            pass
        holder = _Iface()
        holder.__cuda_array_interface__ = {'shape': (2 * self.dim,), 'typestr': '<f8',
                                           'data': (vec.ptr, False), 'version': 2}
        return torch.as_tensor(holder, device='cuda:%d' % self._device_index())

    def _stream(self):
        import torch
        return torch.cuda.current_stream().cuda_stream if self.dist is not None else None

    def _cuda_apply(self, vin, vout, g_begin, g_end, accumulate, dot=None):
        """dot = (ofb_kws handle, vector handle): the launch that completes the output shard also
        leaves conj(vector) . out in the workspace's scalar slots 0, 1 (ofb_apply_groups_dot)."""
        src = vin if isinstance(vin, int) else vin.ptr
        if dot is not None and g_end > g_begin:
            ws, vdot = dot
            self._lib.call('ofb_apply_groups_dot', self.plan._apply_handle, ctypes.c_void_p(src),
                           ctypes.c_void_p(vout.ptr), ctypes.c_int(self.layout.local_bits),
                           ctypes.c_uint64(self.layout.index_base(self.rank)), ctypes.c_int64(g_begin),
                           ctypes.c_int64(g_end), ctypes.c_int(1 if accumulate else 0),
                           ctypes.c_void_p(vdot.ptr), ws, ctypes.c_void_p(self._stream()))
            return
        self.plan.apply_groups_device(src, vout.ptr, self.layout.local_bits,
                                      self.layout.index_base(self.rank), g_begin, g_end,
                                      1 if accumulate else 0, self._stream())
        if dot is not None:  # empty group range: the output is zero, and so is the dot product
            ws, vdot = dot
            self._lib.call('ofb_kws_dot', ws, ctypes.c_void_p(vdot.ptr), ctypes.c_void_p(vout.ptr),
                           ctypes.c_void_p(self._stream()))

    def new_vector(self):
        return self._new()

    def fill_global_random(self, vec, seed, real_only=False):
        """This rank's shard of the counter-based random state keyed by the NATURAL index
        (global_random_amplitudes gives the same values on the host): identical for every world
        size, shard basis and local qubit order."""
        cols = numpy.zeros(40, dtype=numpy.uint64)
        cols[:self.n_qubits] = numpy.array(self.basis.cols, dtype=numpy.uint64)
        self._lib.call('ofb_vec_fill_random_mapped', ctypes.c_int64(self.dim), ctypes.c_uint64(seed),
                       ctypes.c_int(1 if real_only else 0), ctypes.c_void_p(vec.ptr),
                       ctypes.c_uint64(self.layout.index_base(self.rank)), ctypes.c_int(self.n_qubits),
                       self._lib.ptr(cols), ctypes.c_void_p(self._stream()))

    def natural_indices(self, rank=None):
        """Natural (reference-order) index of every amplitude of a rank's shard."""
        rank = self.rank if rank is None else rank
        return self.basis.shard_natural_indices(rank, self.layout.local_bits)

    def shard_of(self, full_state, rank=None):
        """This rank's shard of a full-length host vector given in the reference's order."""
        full_state = numpy.asarray(full_state)
        if self.basis.identity:
            rank = self.rank if rank is None else rank
            return full_state[rank * self.dim:(rank + 1) * self.dim].copy()
        return full_state[self.natural_indices(rank).astype(numpy.int64)]

    # -- p2p registration ------------------------------------------------------------------
    def register_peer_vectors(self, vectors):
        """p2p mode: exchange CUDA IPC handles of the given local vectors once, so that later
        matvecs can read any peer's copy of the same vector directly."""
        import torch
        lib = self._lib
        handles = torch.zeros((len(vectors), 64), dtype=torch.uint8)
        for i, vec in enumerate(vectors):
            raw = (ctypes.c_ubyte * 64)()
            lib.call('ofb_ipc_get_handle', ctypes.c_void_p(vec.ptr), raw)
            handles[i] = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
        device = 'cuda:%d' % self._device_index()
        gathered = [torch.zeros_like(handles, device=device) for _ in range(self.world)]
        self.dist.all_gather(gathered, handles.to(device), group=self.group)
        for i, vec in enumerate(vectors):
            ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(vec.ptr)
                    continue
                raw = (ctypes.c_ubyte * 64).from_buffer_copy(bytes(gathered[r][i].cpu().numpy()))
                p = ctypes.c_void_p()
                lib.call('ofb_ipc_open_handle', raw, ctypes.byref(p))
                self._opened.append(p.value)
                ptrs.append(p.value)
            self._peer_vectors[vec.ptr] = ptrs

    def unregister_peer_vectors(self, vectors):
        """Collective: unmap the peers' copies of these vectors and forget them.  MUST run before
        the vectors are freed -- freeing exported memory while an importer still maps it is
        undefined, and a later allocation may reuse the address.  Ends with a barrier, after
        which every rank may free its buffers."""
        import torch
        torch.cuda.synchronize()
        for vec in vectors:
            ptrs = self._peer_vectors.pop(vec.ptr, None)
            if ptrs is None:
                continue
            for r, p in enumerate(ptrs):
                if r != self.rank:
                    self._lib.call('ofb_ipc_close_handle', ctypes.c_void_p(p))
                    if p in self._opened:
                        self._opened.remove(p)
        self.dist.barrier(group=self.group)

    def close(self):
        """Unmap the peers' buffers and free this rank's staging shards (call after the last
        matvec has completed on the device)."""
        for p in self._opened:
            try:
                self._lib.call('ofb_ipc_close_handle', ctypes.c_void_p(p))
            except Exception:
                pass
        self._opened = []
        self._peer_vectors = {}
        for buf in self._staging:
            if hasattr(buf, 'free'):
                buf.free()
        self._staging = []

    # -- the sharded matvec --------------------------------------------------------------------
    def _launch(self, src, vout, g0, g1, accumulate, dot=None):
        if dot is None:
            self._local_apply(src, vout, g0, g1, accumulate)
        else:
            self._local_apply(src, vout, g0, g1, accumulate, dot)

    def _matvec_ce(self, vin, vout, dot=None):
        """Copy-engine pull of the peer shards overlapped with the local groups."""
        lib, layout = self._lib, self.layout
        vp = ctypes.c_void_p
        remote = layout.remote_patterns
        ptrs = self._peer_vectors.get(vin.ptr)
        if ptrs is None:
            raise RuntimeError('ce mode: vector was not registered with register_peer_vectors')
        depth = len(self._staging)
        lanes = self._copy_lanes
        if self._copy_stream is None:
            # a shard copy can be cut into OFB_CE_LANES pieces on as many streams (several copy
            # engines pulling at once).  Measured on 8 B200s: 4 lanes 88.8 ms per step against
            # 72.7 ms with one lane -- the concurrent copies take HBM bandwidth from the
            # kernels they are meant to hide behind -- so the default is one lane.
            self._copy_stream = []
            for _ in range(lanes):
                s = vp()
                lib.call('ofb_stream_create', ctypes.byref(s))
                self._copy_stream.append(s)
            for _ in range((lanes + 1) * max(depth, 1) + 1):
                e = vp()
                lib.call('ofb_event_create', ctypes.byref(e))
                self._events.append(e)
        main = vp(self._stream())
        # copied[slot][lane]: piece `lane` of the copy into staging[slot] landed
        copied = [self._events[k * lanes:(k + 1) * lanes] for k in range(depth)]
        consumed = self._events[depth * lanes:depth * lanes + depth]  # kernel done with staging[slot]
        start = self._events[depth * lanes + depth]
        piece = self.dim * 16 // lanes
        # copies may start once everything already queued on the main stream (the previous
        # step's kernels, which read the staging buffers) is done
        lib.call('ofb_event_record', start, main)
        for stream in self._copy_stream:
            lib.call('ofb_stream_wait_event', stream, start)

        def post(i):
            peer = layout.peer(self.rank, remote[i][0])
            slot = i % depth
            for lane, stream in enumerate(self._copy_stream):
                if i >= depth:
                    lib.call('ofb_stream_wait_event', stream, consumed[slot])
                lib.call('ofb_memcpy_d2d', vp(self._staging[slot].ptr + lane * piece),
                         vp(ptrs[peer] + lane * piece), ctypes.c_size_t(piece), stream)
                lib.call('ofb_event_record', copied[slot][lane], stream)

        for i in range(min(depth, len(remote))):
            post(i)
        first = True
        if layout.local_pattern is not None:
            _, g0, g1 = layout.local_pattern
            self._launch(vin, vout, g0, g1, False, None if remote else dot)
            first = False
        for i, (x_hi, g0, g1) in enumerate(remote):
            for event in copied[i % depth]:
                lib.call('ofb_stream_wait_event', main, event)
            self._launch(self._staging[i % depth], vout, g0, g1, not first,
                         dot if i == len(remote) - 1 else None)
            first = False
            lib.call('ofb_event_record', consumed[i % depth], main)
            if i + depth < len(remote):
                post(i + depth)
        if first:
            self._launch(vin, vout, 0, 0, False, dot)
        return vout

    def matvec(self, vin, vout, dot=None):
        """vout = (H vin) restricted to this rank's shard.  vin/vout: local shard handles.
        In the 'p2p' and 'ce' modes the caller must make sure (barrier) that every peer has
        finished writing its copy of vin before this call and does not overwrite it before
        all ranks have completed the call.  dot = (ofb_kws, vector): the launch that completes
        vout also computes the local part of conj(vector) . vout (CUDA path only)."""
        layout = self.layout
        first = True
        if self.mode == 'ce' and layout.remote_patterns:
            return self._matvec_ce(vin, vout, dot)
        if self.mode == 'p2p' and layout.remote_patterns:
            ptrs = self._peer_vectors.get(vin.ptr)
            if ptrs is None:
                raise RuntimeError('p2p mode: vector was not registered with register_peer_vectors')
            for k, (x_hi, g0, g1) in enumerate(layout.patterns):
                src = vin if x_hi == 0 else ptrs[layout.peer(self.rank, x_hi)]
                self._launch(src, vout, g0, g1, not first, dot if k == len(layout.patterns) - 1 else None)
                first = False
            if first:
                self._launch(vin, vout, 0, 0, False, dot)
            return vout
        remote = layout.remote_patterns
        depth = len(self._staging)
        if remote and depth == 0:
            raise RuntimeError('no staging buffers for the exchange path')
        works = {}

        def post(i):
            x_hi = remote[i][0]
            peer = layout.peer(self.rank, x_hi)
            dist = self.dist
            ops = [dist.P2POp(dist.isend, self._tensor_of(vin), peer, self.group),
                   dist.P2POp(dist.irecv, self._tensor_of(self._staging[i % depth]), peer, self.group)]
            works[i] = dist.batch_isend_irecv(ops)

        for i in range(min(depth, len(remote))):
            post(i)
        if layout.local_pattern is not None:
            _, g0, g1 = layout.local_pattern
            self._launch(vin, vout, g0, g1, False, None if remote else dot)
            first = False
        for i, (x_hi, g0, g1) in enumerate(remote):
            for work in works.pop(i):
                work.wait()
            self._launch(self._staging[i % depth], vout, g0, g1, not first,
                         dot if i == len(remote) - 1 else None)
            first = False
            if i + depth < len(remote):
                post(i + depth)  # enqueued after the kernel that still reads this buffer
        if first:
            self._launch(vin, vout, 0, 0, False, dot)
        return vout


def init_nccl(local_rank):
    """torch.distributed NCCL group whose communication streams have HIGH priority: the
    exchange kernels must get SM slots while the (grid-filling) apply kernel is running,
    otherwise the block scheduler drains the apply grid first and nothing overlaps."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    kwargs = {}
    try:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        kwargs['pg_options'] = opts
    except Exception:
        pass
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank), **kwargs)
    return dist


def lanczos_ground_state(operator, dist, group=None, tol=1e-10, krylov_dim=None, max_restarts=40,
                         seed=1234, stats=None):
    """Distributed device-resident Lanczos: the fused, device-scalar recurrence of csrc/krylov.cu
    with the two reductions of an iteration all-reduced over the ranks (NCCL, stream-ordered).
    Returns (energy, device buffer with the local shard of the unit eigenvector, residual norm)."""
    from openfermion_b200.sharded_lanczos import ShardedLanczos
    solver = ShardedLanczos(operator, dist, group, max_steps=krylov_dim or 400)
    try:
        return solver.solve(tol=tol, max_restarts=max_restarts, seed=seed, stats=stats)
    finally:
        solver.close()


# ----------------------------------------------------------------------------
# bench entry for N > 1 (called by bench.py under torchrun)
# ----------------------------------------------------------------------------
def one_gpu_base(n_qubits):
    """One-GPU measurement of the same Hubbard workload (profiles/sweep_r01.jsonl), so that the
    strong-scaling base of a --gpus N line is explicit (the N = 1 bench line is a different
    workload: BASELINE.json quotes the metric at 28 qubits on 1 GPU and 32 qubits on 2/4/8)."""
    profiles = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles')
    for name in ('one_gpu_hubbard32_r02.json', 'sweep_r01.jsonl'):  # newest record first
        try:
            for raw in open(os.path.join(profiles, name)):
                rec = json.loads(raw)
                if rec.get('n_qubits') == n_qubits and rec.get('workload', '').startswith('hubbard'):
                    return {'ms_per_step': rec['ms_per_matvec'], 'value': rec['term_amp_per_s'],
                            'source': rec.get('source', 'profiles/' + name)}
        except Exception:
            pass
    return None


def bench_main(args, metric, unit, clock_sampler_cls, measured_peak, output_check=None):
    import torch
    import torch.distributed as dist
    from openfermion_b200 import _lib, hamiltonians
    from openfermion_b200.linalg import krylov

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    init_nccl(local_rank)
    _lib.call('ofb_set_device', ctypes.c_int(local_rank))
    n = args.dist_qubits
    if n == 32:
        op_model, name = hamiltonians.hubbard_jw(4, 4), 'Fermi-Hubbard 4x4 spinful JW (t=1, U=4, periodic), 32 qubits'
    else:
        sites = n // 2
        dims = {12: (2, 3), 16: (2, 4), 20: (2, 5), 24: (3, 4), 28: (2, 7), 30: (3, 5)}[sites * 2]
        op_model, name = hamiltonians.hubbard_jw(*dims), 'Fermi-Hubbard %dx%d spinful JW, %d qubits' % (dims + (n,))
    mode = args.dist_mode
    basis_kind = getattr(args, 'dist_basis', 'symmetry')
    local_order = getattr(args, 'dist_order', 'locality') if basis_kind == 'symmetry' else 'natural'
    op = ShardedPauliOperator(op_model, n, dist=dist, mode=mode, basis=basis_kind,
                              local_order=local_order)
    layout = op.layout
    natural_layout = ShardLayout(n, world, op._natural_x)
    dim_local = op.dim
    vin, vout = op.new_vector(), op.new_vector()
    # ONE global random state keyed by the natural index: every world size, shard basis and local
    # qubit order applies H to the same |psi>, so <psi|H|psi> and sampled outputs are comparable
    # across the --gpus N lines (and with the one-GPU aux measurement of the N = 1 line)
    state_seed = 42
    op.fill_global_random(vin, state_seed)
    comm = TorchComm(dist, 'cuda:%d' % local_rank)
    state_norm = krylov.gnorm(dim_local, vin.ptr, comm)
    krylov.scal(dim_local, 1.0 / state_norm, vin.ptr)
    exchanging = mode in ('p2p', 'ce') and bool(layout.remote_patterns)
    if exchanging:
        op.register_peer_vectors([vin])
    T, G = op.n_terms, op.n_groups
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def step():
        if exchanging:
            comm.barrier()  # peers must have finished writing the vector we are about to read
        op.matvec(vin, vout)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = clock_sampler_cls(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    torch.cuda.synchronize()
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms_local = start.elapsed_time(stop)
    t = torch.tensor([ms_local], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # consistency: <v|Hv> must be real for a Hermitian H and identical on every rank
    e = krylov.gdot(dim_local, vin.ptr, vout.ptr, comm)

    diag = None
    if getattr(args, 'dist_diagnose', False) and mode == 'nccl':
        # where does the step time go: exchange alone, kernels alone (staging = stale data)
        def timed(fn, reps=3):
            torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def exchange_only():
            for i, (x_hi, g0, g1) in enumerate(layout.remote_patterns):
                peer = layout.peer(rank, x_hi)
                ops = [dist.P2POp(dist.isend, op._tensor_of(vin), peer),
                       dist.P2POp(dist.irecv, op._tensor_of(op._staging[i % len(op._staging)]), peer)]
                for w in dist.batch_isend_irecv(ops):
                    w.wait()

        def kernels_only():
            first = True
            for x_hi, g0, g1 in layout.patterns:
                src = vin if x_hi == 0 else op._staging[0]
                op._local_apply(src, vout, g0, g1, not first)
                first = False
        diag = {'exchange_only_ms': timed(exchange_only), 'kernels_only_ms': timed(kernels_only)}

    # e2e: host shards in pinned memory -> device, matvec, result shard back to the host
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    vp = ctypes.c_void_p
    h_in, h_out = vp(), vp()
    nbytes = dim_local * 16
    _lib.call('ofb_host_alloc', ctypes.byref(h_in), ctypes.c_size_t(nbytes))
    _lib.call('ofb_host_alloc', ctypes.byref(h_out), ctypes.c_size_t(nbytes))
    stream = vp(torch.cuda.current_stream().cuda_stream)
    _lib.call('ofb_memcpy_d2h', h_in, vp(vin.ptr), ctypes.c_size_t(nbytes), stream)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _lib.call('ofb_memcpy_h2d', vp(vin.ptr), h_in, ctypes.c_size_t(nbytes), stream)
        step()
        _lib.call('ofb_memcpy_d2h', h_out, vp(vout.ptr), ctypes.c_size_t(nbytes), stream)
        torch.cuda.synchronize()
    dist.barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    _lib.call('ofb_host_free', h_in)
    _lib.call('ofb_host_free', h_out)

    # sampled outputs in NATURAL order, gathered on every rank: (H psi)[j] for seeded indices j
    check = None
    n_samples = int(getattr(args, 'check_samples', 256))
    rng = numpy.random.default_rng(11)
    sample_idx = numpy.unique(rng.integers(0, 1 << n, size=n_samples, dtype=numpy.int64)).astype(numpy.uint64)
    where = op.basis.to_basis_index(sample_idx)
    owner = (where >> numpy.uint64(layout.local_bits)).astype(numpy.int64)
    local_pos = (where & numpy.uint64(dim_local - 1)).astype(numpy.int64)
    gathered = torch.zeros((len(sample_idx), 2), dtype=torch.float64, device='cuda')
    host_vals = numpy.zeros((len(sample_idx), 2))
    one = numpy.zeros(1, dtype=numpy.complex128)
    for k in numpy.nonzero(owner == rank)[0]:
        _lib.call('ofb_memcpy_d2h', _lib.ptr(one), vp(vout.ptr + int(local_pos[k]) * 16), ctypes.c_size_t(16), stream)
        torch.cuda.synchronize()
        host_vals[k] = (one[0].real, one[0].imag)
    gathered.copy_(torch.from_numpy(host_vals))
    dist.all_reduce(gathered)
    sampled = gathered.cpu().numpy()
    sampled = sampled[:, 0] + 1j * sampled[:, 1]
    if rank == 0 and output_check is not None:
        # the caller's checker (bench.py: the CPU oracle on the reference-order term table) gets
        # the sampled outputs and a function that evaluates psi at any natural index
        check = output_check(op_model, n, sample_idx,
                             lambda idx: global_random_amplitudes(state_seed, idx) / state_norm, sampled)
    checksum = complex(numpy.sum(sampled))

    # Lanczos: per-iteration time of the device-resident recurrence (config 5), all reductions
    # included, no host synchronisation inside the timed region
    lanczos = None
    lanczos_steps = int(getattr(args, 'lanczos_steps', 20))
    if lanczos_steps > 0:
        from openfermion_b200.sharded_lanczos import ShardedLanczos
        if exchanging:
            op.unregister_peer_vectors([vin])
        vin.free()
        vout.free()
        warm = 3
        solver = ShardedLanczos(op, dist, max_steps=lanczos_steps + warm + 1, vectors=3)
        solver.prepare(seed=state_seed)
        solver.run(warm)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches1 = _lib.launch_count()
        a.record()
        solver.run(lanczos_steps)
        b.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([a.elapsed_time(b) / lanczos_steps], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        _, _, ritz = solver.history()
        per_it = float(t.item())
        lanczos = {'iterations_timed': lanczos_steps, 'warmup_iterations': warm, 'ms_per_iteration': per_it,
                   'overhead_over_bare_apply': per_it / (ms / args.steps) - 1.0,
                   'kernel_launches_per_iteration': (_lib.launch_count() - launches1) / float(lanczos_steps),
                   'lowest_ritz_value_after_%d_iterations' % (lanczos_steps + warm): ritz,
                   'term_amp_per_s_inside_lanczos': T * float(1 << n) / (per_it / 1e3),
                   'what': 'H u (alpha fused into the apply launch) + all-reduce(2) + one fused update pass '
                           '(3 vectors read, 1 written, norm) + all-reduce(1); scalars stay on the device'}
        solver.close()

    if rank == 0:
        dim = float(1 << n)
        seconds = ms / 1e3
        value = T * dim * args.steps / seconds
        peak, peak_src = measured_peak()
        alg_bytes_rank = 16.0 * dim_local * (G + 1)
        achieved = alg_bytes_rank / (seconds / args.steps) / 1e9
        line = {
            'metric': metric, 'value': value, 'unit': unit, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': name, 'n_qubits': n, 'terms': T, 'x_groups': G,
                       'sharding': '%d shard bits over %d ranks: %s' % (
                           layout.shard_bits, world, op.basis.describe(layout.shard_bits)),
                       'local_qubit_order': local_order,
                       'exchange': mode, 'remote_patterns': len(layout.remote_patterns),
                       'remote_patterns_top_qubit_sharding': len(natural_layout.remote_patterns),
                       'exchange_bytes_per_rank_per_step': layout.exchange_bytes_per_matvec() if mode in ('nccl', 'ce') else None,
                       'support_hints': os.environ.get('OFB_SUPPORT_HINTS', '1'),
                       'l2_policy': 'inputs exceed L2 (%.1f GB per shard)' % (dim_local * 16 / 1e9),
                       'timing': 'CUDA events, max over ranks',
                       'state': 'global counter-based random state keyed by the natural index, seed %d, '
                                'normalised: the same |psi> at every N' % state_seed,
                       'expectation_value': [e.real, e.imag],
                       'output_checksum': [checksum.real, checksum.imag],
                       'self_check_rel_err': check,
                       'self_check': 'max |(H psi)[j] - oracle| / max |oracle| over %d seeded natural indices j '
                                     '(oracle: reference-order term sum on the host); bar 1e-12' % len(sample_idx),
                       'lanczos': lanczos,
                       'one_gpu_same_workload': one_gpu_base(n),
                       'diagnose': diag},
            'clocks': clocks,
            'e2e': {'value': T * dim * e2e_steps / e2e_s, 'unit': unit,
                    'h2d_bytes_per_step': nbytes * world, 'd2h_bytes_per_step': nbytes * world,
                    'steps': e2e_steps, 'path': 'pinned host shards -> ofb_apply_groups -> host'},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                         'kernel': 'k_apply (per rank)', 'algorithmic_bytes_per_launch': alg_bytes_rank,
                         'nvlink_bytes_per_rank_per_step': layout.exchange_bytes_per_matvec()},
            'cpu_baseline': None,
        }
        print(json.dumps(line))
    op.close()
    dist.destroy_process_group()
