// Plan construction (host side) and the memory / stream / event plumbing of the C ABI.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "ofb_internal.cuh"

namespace ofb {

static thread_local std::string t_error;
std::atomic<int64_t> g_launches{0};

void set_error(const std::string &msg) { t_error = msg; }
int fail(int code, const std::string &msg) {
    t_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    t_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " at " + file +
              ":" + std::to_string(line);
    cudaGetLastError();  // clear sticky-free errors
    return e == cudaErrorMemoryAllocation ? OFB_ERR_NOMEM : OFB_ERR_CUDA;
}

int ensure_partials(ofb_plan *plan, size_t bytes) {
    if (plan->partials_bytes >= bytes) return OFB_OK;
    if (plan->d_partials) cudaFree(plan->d_partials);
    plan->d_partials = nullptr;
    plan->partials_bytes = 0;
    OFB_CUDA(cudaMalloc(&plan->d_partials, bytes));
    plan->partials_bytes = bytes;
    return OFB_OK;
}

int ensure_stage(ofb_plan *plan, size_t in_bytes, size_t out_bytes) {
    if (plan->stage_in_bytes < in_bytes) {
        if (plan->d_stage_in) cudaFree(plan->d_stage_in);
        plan->d_stage_in = nullptr;
        plan->stage_in_bytes = 0;
        OFB_CUDA(cudaMalloc(&plan->d_stage_in, in_bytes));
        plan->stage_in_bytes = in_bytes;
    }
    if (plan->stage_out_bytes < out_bytes) {
        if (plan->d_stage_out) cudaFree(plan->d_stage_out);
        plan->d_stage_out = nullptr;
        plan->stage_out_bytes = 0;
        OFB_CUDA(cudaMalloc(&plan->d_stage_out, out_bytes));
        plan->stage_out_bytes = out_bytes;
    }
    return OFB_OK;
}

// Multiply (re, im) by (-i)^k (gather form) -- exact.
static inline void mul_neg_i_pow(double &re, double &im, int k) {
    double r = re, i = im;
    switch (k & 3) {
        case 0: break;
        case 1: re = i; im = -r; break;   // * (-i)
        case 2: re = -r; im = -i; break;  // * (-1)
        case 3: re = -i; im = r; break;   // * (+i)
    }
}

template <typename T>
static int upload(T **d_ptr, const std::vector<T> &h) {
    *d_ptr = nullptr;
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    OFB_CUDA(cudaMalloc(d_ptr, bytes));
    if (!h.empty()) OFB_CUDA(cudaMemcpy(*d_ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return OFB_OK;
}

// Sibling-subtree sizes of the binary trie over the ascending group x-masks:
// sib[g][b] = #{h : x_h agrees with x_g above bit b and differs at bit b}.
// For a column c the rows c ^ x_g sort ascending with
//   rank_c(g) = sum over set bits b of (c ^ x_g) of sib[g][b],
// because at the highest differing bit the entry whose xored bit is 0 comes first.
static void build_sibling_table(const std::vector<uint64_t> &gx, int n, std::vector<uint32_t> &sib) {
    size_t G = gx.size();
    sib.assign(G * (size_t)std::max(n, 1), 0);
    for (size_t g = 0; g < G; ++g) {
        for (int b = 0; b < n; ++b) {
            uint64_t q = (gx[g] >> b) ^ 1ull;  // sibling prefix including bit b
            uint64_t lo = q << b;
            uint64_t hi = (b + 1 >= 64) ? ~0ull : ((q + 1) << b);
            auto it_lo = std::lower_bound(gx.begin(), gx.end(), lo);
            auto it_hi = (q + 1 == 0) ? gx.end() : std::lower_bound(gx.begin(), gx.end(), hi);
            sib[g * n + b] = (uint32_t)(it_hi - it_lo);
        }
    }
}

static int finish_plan(ofb_plan *p, const std::vector<uint64_t> &x, const std::vector<uint32_t> &order,
                       std::vector<ScatterRow> &rows_sorted) {
    // groups from the sorted order
    int64_t T = (int64_t)order.size();
    p->h_group_x.clear();
    p->h_group_begin.clear();
    for (int64_t k = 0; k < T; ++k) {
        uint64_t xv = x[order[k]];
        if (k == 0 || xv != p->h_group_x.back()) {
            p->h_group_x.push_back(xv);
            p->h_group_begin.push_back((uint32_t)k);
        }
    }
    p->h_group_begin.push_back((uint32_t)T);
    p->n_groups = (int64_t)p->h_group_x.size();
    p->n_terms = T;
    p->n_diag_terms = (p->n_groups > 0 && p->h_group_x[0] == 0) ? p->h_group_begin[1] : 0;
    std::vector<uint32_t> sib;
    build_sibling_table(p->h_group_x, p->n_qubits, sib);
    int rc;
    if ((rc = upload(&p->d_rows, rows_sorted))) return rc;
    {
        std::vector<ScatterRow> rows_orig(rows_sorted.size());
        for (size_t k = 0; k < order.size(); ++k) rows_orig[order[k]] = rows_sorted[k];
        if ((rc = upload(&p->d_rows_orig, rows_orig))) return rc;
    }
    if ((rc = upload(&p->d_sib, sib))) return rc;
    OFB_CUDA(cudaMallocHost(&p->h_pinned_scalars, 64 * sizeof(double)));
    return OFB_OK;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_version(void) { return 100; }
const char *ofb_last_error(void) { return t_error.c_str(); }
int64_t ofb_launch_count(void) { return g_launches.load(); }

int ofb_device_count(int *count) {
    if (!count) return fail(OFB_ERR_INVALID, "count is NULL");
    *count = 0;
    OFB_CUDA(cudaGetDeviceCount(count));
    return OFB_OK;
}
int ofb_set_device(int device) {
    OFB_CUDA(cudaSetDevice(device));
    return OFB_OK;
}
int ofb_device_info(int device, int *sm_count, size_t *total_bytes, size_t *free_bytes, int *cc_major,
                    int *cc_minor) {
    cudaDeviceProp prop;
    OFB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_bytes || free_bytes) {
        int cur = 0;
        OFB_CUDA(cudaGetDevice(&cur));
        OFB_CUDA(cudaSetDevice(device));
        size_t f = 0, t = 0;
        OFB_CUDA(cudaMemGetInfo(&f, &t));
        OFB_CUDA(cudaSetDevice(cur));
        if (total_bytes) *total_bytes = t;
        if (free_bytes) *free_bytes = f;
    }
    return OFB_OK;
}

int ofb_malloc(void **d_ptr, size_t bytes) {
    if (!d_ptr) return fail(OFB_ERR_INVALID, "d_ptr is NULL");
    *d_ptr = nullptr;
    OFB_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return OFB_OK;
}
int ofb_free(void *d_ptr) {
    if (d_ptr) OFB_CUDA(cudaFree(d_ptr));
    return OFB_OK;
}
int ofb_host_alloc(void **h_ptr, size_t bytes) {
    if (!h_ptr) return fail(OFB_ERR_INVALID, "h_ptr is NULL");
    *h_ptr = nullptr;
    OFB_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 1));
    return OFB_OK;
}
int ofb_host_free(void *h_ptr) {
    if (h_ptr) OFB_CUDA(cudaFreeHost(h_ptr));
    return OFB_OK;
}
int ofb_host_register(void *h_ptr, size_t bytes) {
    OFB_CUDA(cudaHostRegister(h_ptr, bytes, cudaHostRegisterDefault));
    return OFB_OK;
}
int ofb_host_unregister(void *h_ptr) {
    OFB_CUDA(cudaHostUnregister(h_ptr));
    return OFB_OK;
}
int ofb_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream) {
    OFB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return OFB_OK;
}
int ofb_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream) {
    OFB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return OFB_OK;
}
int ofb_memcpy_d2d(void *d_dst, const void *d_src, size_t bytes, void *stream) {
    OFB_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return OFB_OK;
}
int ofb_memset(void *d_ptr, int byte, size_t bytes, void *stream) {
    OFB_CUDA(cudaMemsetAsync(d_ptr, byte, bytes, (cudaStream_t)stream));
    return OFB_OK;
}
int ofb_stream_create(void **stream) {
    cudaStream_t s;
    OFB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return OFB_OK;
}
int ofb_stream_destroy(void *stream) {
    OFB_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return OFB_OK;
}
int ofb_stream_sync(void *stream) {
    OFB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return OFB_OK;
}
int ofb_device_sync(void) {
    OFB_CUDA(cudaDeviceSynchronize());
    return OFB_OK;
}
int ofb_event_create(void **event) {
    cudaEvent_t e;
    OFB_CUDA(cudaEventCreate(&e));
    *event = (void *)e;
    return OFB_OK;
}
int ofb_event_destroy(void *event) {
    OFB_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return OFB_OK;
}
int ofb_event_record(void *event, void *stream) {
    OFB_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return OFB_OK;
}
int ofb_event_sync(void *event) {
    OFB_CUDA(cudaEventSynchronize((cudaEvent_t)event));
    return OFB_OK;
}
int ofb_event_elapsed_ms(void *start, void *stop, float *ms) {
    OFB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return OFB_OK;
}
int ofb_stream_wait_event(void *stream, void *event) {
    OFB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
    return OFB_OK;
}
int ofb_ipc_get_handle(void *d_ptr, void *handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    cudaIpcMemHandle_t h;
    OFB_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle64, &h, 64);
    return OFB_OK;
}
int ofb_ipc_open_handle(const void *handle64, void **d_ptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    OFB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return OFB_OK;
}
int ofb_ipc_close_handle(void *d_ptr) {
    OFB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return OFB_OK;
}
int ofb_enable_peer_access(int peer_device) {
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return OFB_OK;
    }
    OFB_CUDA(e);
    return OFB_OK;
}

int ofb_plan_create(int n_qubits, int64_t n_terms, const uint64_t *x_mask, const uint64_t *z_mask,
                    const double *coeff_ri, int device, ofb_plan **plan) {
    if (!plan) return fail(OFB_ERR_INVALID, "plan is NULL");
    *plan = nullptr;
    ofb_plan *p = nullptr;
    OFB_TRY
    if (n_qubits < 0 || n_qubits > 40) return fail(OFB_ERR_INVALID, "n_qubits out of range [0, 40]");
    if (n_terms < 0 || n_terms >= (1ll << 30)) return fail(OFB_ERR_INVALID, "n_terms out of range");
    if (n_terms > 0 && (!x_mask || !z_mask || !coeff_ri))
        return fail(OFB_ERR_INVALID, "NULL term table");
    uint64_t full = n_qubits >= 64 ? ~0ull : ((1ull << n_qubits) - 1);
    for (int64_t t = 0; t < n_terms; ++t)
        if ((x_mask[t] | z_mask[t]) & ~full)
            return fail(OFB_ERR_INVALID, "term " + std::to_string(t) + " acts outside n_qubits");
    OFB_CUDA(cudaSetDevice(device));
    p = new (std::nothrow) ofb_plan();
    if (!p) return fail(OFB_ERR_NOMEM, "host allocation failed");
    p->n_qubits = n_qubits;
    p->device = device;
    p->projected = false;
    p->z_fits32 = n_qubits <= 32;

    std::vector<uint64_t> x(x_mask, x_mask + n_terms);
    std::vector<uint32_t> order(n_terms);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return x[a] < x[b]; });

    std::vector<ScatterRow> rows(n_terms);
    std::vector<double> gr(n_terms), gi(n_terms);
    for (int64_t k = 0; k < n_terms; ++k) {
        uint32_t t = order[k];
        int y = __builtin_popcountll(x_mask[t] & z_mask[t]);
        double re = coeff_ri[2 * t], im = coeff_ri[2 * t + 1];
        double sre = re, sim = im;
        mul_neg_i_pow(re, im, y);         // gather form: coeff * (-i)^y
        mul_neg_i_pow(sre, sim, (4 - (y & 3)) & 3);  // scatter form: coeff * (+i)^y
        gr[k] = re;
        gi[k] = im;
        rows[k] = ScatterRow{z_mask[t], 0, 0, x_mask[t], sre, sim,
                             sz_flags(coeff_ri[2 * t], coeff_ri[2 * t + 1], z_mask[t], true)};
    }
    int rc = finish_plan(p, x, order, rows);
    if (rc) {
        ofb_plan_destroy(p);
        return rc;
    }
    // gather-form tables: per logical group, terms sorted by class and cut into segments
    std::vector<Segment> segs;
    std::vector<TermZC> zc_re, zc_im;
    zc_re.reserve(n_terms + 1);
    zc_im.reserve(n_terms + 1);
    p->h_group_kind.resize(p->n_groups);
    p->h_seg_begin.assign(p->n_groups + 1, 0);
    std::vector<uint32_t> idx;
    for (int64_t g = 0; g < p->n_groups; ++g) {
        uint32_t b = p->h_group_begin[g], e = p->h_group_begin[g + 1];
        bool all_real = true, all_imag = true;
        for (uint32_t k = b; k < e; ++k) {
            if (gi[k] != 0.0) all_real = false;
            if (gr[k] != 0.0) all_imag = false;
        }
        uint32_t kind = all_real ? KIND_REAL : (all_imag ? KIND_IMAG : KIND_COMPLEX);
        p->h_group_kind[g] = kind;
        {
            // tiled CSC passes (csc_tile.cu): a part that is zero in every term of the group is +0
            // in the sum unless EVERY non-zero term needs the signed-zero replay; only then the
            // group takes the per-row path
            bool slow = kind != KIND_COMPLEX, any = false;
            for (uint32_t k = b; k < e; ++k) {
                if (rows[k].cr == 0.0 && rows[k].ci == 0.0) continue;
                any = true;
                if (!(rows[k].flags & SZ_NEED_REPLAY)) slow = false;
            }
            p->h_group_slow.push_back((slow && any) ? 1 : 0);
        }
        p->h_seg_begin[g] = (uint32_t)segs.size();
        idx.resize(e - b);
        std::iota(idx.begin(), idx.end(), b);
        auto cls_of = [&](uint32_t k) { return (uint32_t)(z_mask[order[k]] >> CLASS_SHIFT) & CLASS_MASK; };
        std::stable_sort(idx.begin(), idx.end(),
                         [&](uint32_t a, uint32_t c2) { return cls_of(a) < cls_of(c2); });
        for (size_t pos = 0; pos < idx.size();) {
            size_t end = std::min(idx.size(), pos + (size_t)SEG_MAX_TERMS);
            // a chunk of a class-sorted list is still class-sorted; a segment also ends before
            // its (SEG_MAX_RUNS + 1)-th class run (only possible with 16 classes)
            {
                uint32_t seen = 0;
                int prev = -1;
                for (size_t q = pos; q < end; ++q) {
                    int cw = (int)cls_of(idx[q]);
                    if (cw != prev) {
                        if (seen == (uint32_t)SEG_MAX_RUNS) {
                            end = q;
                            break;
                        }
                        ++seen;
                        prev = cw;
                    }
                }
            }
            Segment sd{};
            sd.x = p->h_group_x[g];
            sd.x_lo = (uint32_t)p->h_group_x[g];
            sd.rel_begin = (uint32_t)zc_re.size();  // absolute for now; made chunk-relative below
            uint64_t cnts = 0;
            uint32_t ids = 0, runs = 0;
            int last_cls = -1;
            for (size_t q = pos; q < end; ++q) {
                uint32_t k = idx[q];
                int cw = (int)cls_of(k);
                if (cw != last_cls) {
                    ids |= (uint32_t)cw << (4 * runs);
                    ++runs;
                    last_cls = cw;
                }
                cnts += 1ull << (8 * (runs - 1));
                zc_re.push_back(TermZC{z_mask[order[k]], gr[k]});
                zc_im.push_back(TermZC{z_mask[order[k]], gi[k]});
            }
            // bits 0..3 class runs, bit 4 last segment of its group, 8..23 terms, 30..31 kind
            sd.kind_runs = runs | ((end == idx.size() ? 1u : 0u) << 4) | ((uint32_t)(end - pos) << 8) | (kind << 30);
            sd.cnts = cnts;
            sd.ids = ids;
            segs.push_back(sd);
            pos = end;
        }
    }
    p->h_seg_begin[p->n_groups] = (uint32_t)segs.size();
    std::vector<uint32_t> seg_term_begin(segs.size());
    for (size_t sidx = 0; sidx < segs.size(); ++sidx) seg_term_begin[sidx] = segs[sidx].rel_begin;
    // chunks: runs of <= CHUNK_MAX_SEGS segments holding <= CHUNK_MAX_TERMS terms
    std::vector<Chunk> chunks;
    p->h_chunk_of_seg.assign(segs.size(), 0);
    for (size_t sidx = 0; sidx < segs.size(); ++sidx) {
        uint32_t cnt = (segs[sidx].kind_runs >> 8) & 0xffffu;
        if ((segs[sidx].kind_runs >> 30) != KIND_REAL) p->has_imag = true;
        if (chunks.empty() || chunks.back().seg_count >= (uint32_t)CHUNK_MAX_SEGS ||
            chunks.back().term_count + cnt > (uint32_t)CHUNK_MAX_TERMS)
            chunks.push_back(Chunk{(uint32_t)sidx, 0, segs[sidx].rel_begin, 0});
        chunks.back().seg_count += 1;
        chunks.back().term_count += cnt;
        segs[sidx].rel_begin -= chunks.back().term_begin;
        p->h_chunk_of_seg[sidx] = (uint32_t)chunks.size() - 1;
    }
    p->n_chunks = (int64_t)chunks.size();
    p->h_chunks = chunks;
    // one padding entry so that a software-prefetched load past the end stays in bounds
    zc_re.push_back(TermZC{0, 0.0});
    zc_im.push_back(TermZC{0, 0.0});
    std::vector<GroupDesc> groups(p->n_groups);
    for (int64_t g = 0; g < p->n_groups; ++g)
        groups[g] = GroupDesc{p->h_group_x[g], p->h_group_begin[g],
                              p->h_group_begin[g + 1] - p->h_group_begin[g]};
    if ((rc = upload(&p->d_chunks, chunks)) || (rc = upload(&p->d_segs, segs)) ||
        (rc = upload(&p->d_seg_term_begin, seg_term_begin)) ||
        (rc = upload(&p->d_zc, zc_re)) ||
        (rc = upload(&p->d_zc_im, zc_im)) || (rc = upload(&p->d_groups, groups))) {
        ofb_plan_destroy(p);
        return rc;
    }
    *plan = p;
    return OFB_OK;
    OFB_CATCH(ofb_plan_destroy(p))
}

int ofb_plan_create_projected(int n_qubits, int64_t n_rows, const uint64_t *occ_mask,
                              const uint64_t *occ_val, const uint64_t *x_mask,
                              const uint64_t *z_mask, const double *coeff_ri, int device,
                              ofb_plan **plan) {
    if (!plan) return fail(OFB_ERR_INVALID, "plan is NULL");
    *plan = nullptr;
    ofb_plan *p = nullptr;
    OFB_TRY
    if (n_qubits < 0 || n_qubits > 40) return fail(OFB_ERR_INVALID, "n_qubits out of range [0, 40]");
    if (n_rows < 0 || n_rows >= (1ll << 30)) return fail(OFB_ERR_INVALID, "n_rows out of range");
    if (n_rows > 0 && (!occ_mask || !occ_val || !x_mask || !z_mask || !coeff_ri))
        return fail(OFB_ERR_INVALID, "NULL row table");
    uint64_t full = (1ull << n_qubits) - 1;
    for (int64_t t = 0; t < n_rows; ++t)
        if ((x_mask[t] | z_mask[t] | occ_mask[t] | occ_val[t]) & ~full)
            return fail(OFB_ERR_INVALID, "row " + std::to_string(t) + " acts outside n_qubits");
    OFB_CUDA(cudaSetDevice(device));
    p = new (std::nothrow) ofb_plan();
    if (!p) return fail(OFB_ERR_NOMEM, "host allocation failed");
    p->n_qubits = n_qubits;
    p->device = device;
    p->projected = true;
    p->z_fits32 = n_qubits <= 32;
    std::vector<uint64_t> x(x_mask, x_mask + n_rows);
    std::vector<uint32_t> order(n_rows);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return x[a] < x[b]; });
    std::vector<ScatterRow> rows(n_rows);
    for (int64_t k = 0; k < n_rows; ++k) {
        uint32_t t = order[k];
        rows[k] = ScatterRow{z_mask[t], occ_mask[t], occ_val[t], x_mask[t], coeff_ri[2 * t], coeff_ri[2 * t + 1],
                             sz_flags(coeff_ri[2 * t], coeff_ri[2 * t + 1], z_mask[t], false)};
    }
    int rc = finish_plan(p, x, order, rows);
    if (rc) {
        ofb_plan_destroy(p);
        return rc;
    }
    *plan = p;
    return OFB_OK;
    OFB_CATCH(ofb_plan_destroy(p))
}

int ofb_plan_destroy(ofb_plan *p) {
    if (!p) return OFB_OK;
    cudaSetDevice(p->device);
    cudaFree(p->d_groups);
    cudaFree(p->d_conds);
    cudaFree(p->d_segs);
    cudaFree(p->d_seg_term_begin);
    cudaFree(p->d_chunks);
    cudaFree(p->d_zc);
    cudaFree(p->d_zc_im);
    cudaFree(p->d_rows);
    cudaFree(p->d_rows_orig);
    cudaFree(p->d_sort_scratch);
    cudaFree(p->d_sib);
    cudaFree(p->d_seg_group);
    cudaFree(p->d_group_slow);
    cudaFree(p->d_count_slices);
    cudaFree(p->d_fill_slices);
    cudaFree(p->d_colptr);
    cudaFree(p->d_bitmap);
    cudaFree(p->d_partials);
    cudaFree(p->d_stage_in);
    cudaFree(p->d_stage_out);
    if (p->h_pinned_scalars) cudaFreeHost(p->h_pinned_scalars);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    for (cudaEvent_t e : p->slab_done)
        if (e) cudaEventDestroy(e);
    cudaGetLastError();
    delete p;
    return OFB_OK;
}

int ofb_plan_set_support(ofb_plan *p, int64_t n_groups, const uint32_t *mask1, const uint32_t *mask2,
                         const uint8_t *flags) {
    if (!p) return fail(OFB_ERR_INVALID, "plan is NULL");
    if (p->projected) return fail(OFB_ERR_INVALID, "projected plans support only csc_* calls");
    if (n_groups != p->n_groups) return fail(OFB_ERR_INVALID, "n_groups differs from the plan's group count");
    if (n_groups > 0 && (!mask1 || !mask2 || !flags)) return fail(OFB_ERR_INVALID, "NULL argument");
    if (p->n_qubits > 32) return fail(OFB_ERR_INVALID, "support hints need n_qubits <= 32");
    OFB_TRY
    OFB_CUDA(cudaSetDevice(p->device));
    cudaFree(p->d_conds);
    p->d_conds = nullptr;
    p->has_cond = false;
    auto pattern = [](uint32_t m) {  // {r : parity(r & w) = 1} for the class bits of m
        const uint32_t w = (m >> CLASS_SHIFT) & CLASS_MASK;
        uint32_t pat = 0;
        for (uint32_t r = 0; r < (1u << CLASS_BITS); ++r)
            if (__builtin_popcount(r & w) & 1) pat |= 1u << r;
        return pat;
    };
    const uint32_t all = (1u << (1u << CLASS_BITS)) - 1u;  // 16 (or 8) amplitudes per thread
    const size_t n_segs = p->h_seg_begin.empty() ? 0 : p->h_seg_begin[p->n_groups];
    std::vector<SegCond> conds(n_segs);
    bool any = false;
    for (int64_t g = 0; g < n_groups; ++g) {
        SegCond cd{0u, 0u, all | (all << 16), 3u};
        if (flags[g] & 1) {
            cd.m1 = mask1[g];
            cd.pat = (cd.pat & 0xffff0000u) | pattern(mask1[g]);
            cd.v = (cd.v & ~1u) | ((flags[g] >> 1) & 1u);
            any = true;
        }
        if (flags[g] & 4) {
            cd.m2 = mask2[g];
            cd.pat = (cd.pat & 0x0000ffffu) | (pattern(mask2[g]) << 16);
            cd.v = (cd.v & ~2u) | (((flags[g] >> 3) & 1u) << 1);
            any = true;
        }
        for (uint32_t sidx = p->h_seg_begin[g]; sidx < p->h_seg_begin[g + 1]; ++sidx) conds[sidx] = cd;
    }
    if (any) {
        int rc = upload(&p->d_conds, conds);
        if (rc) return rc;
        p->has_cond = true;
    }
    return OFB_OK;
    OFB_CATCH((void)0)
}

int ofb_plan_set_csc_mode(ofb_plan *p, int mode) {
    if (!p) return fail(OFB_ERR_INVALID, "plan is NULL");
    if (mode != 0 && mode != 1) return fail(OFB_ERR_INVALID, "csc mode must be 0 or 1");
    p->csc_mode = mode;
    p->csc_nnz = -1;
    return OFB_OK;
}

int ofb_pauli_entry_zero_sign(int n_qubits, uint64_t x_mask, uint64_t z_mask, double re, double im,
                              uint64_t column, int *has_zero, int *negative) {
    if (!has_zero || !negative || n_qubits < 0 || n_qubits > 40) return fail(OFB_ERR_INVALID, "bad argument");
    const uint64_t flags = sz_flags(re, im, z_mask, true);
    *has_zero = (flags & SZ_HAS_ZERO) ? 1 : 0;
    *negative = 0;
    if (flags & SZ_NEED_REPLAY)  // the kernels' shortcut: everything else ends in +0
        *negative = (int)(sz_replay((uint32_t)(flags & SZ_STATE), x_mask, z_mask, column, n_qubits) & 1u);
    return OFB_OK;
}

int ofb_plan_info(const ofb_plan *p, int *n_qubits, int64_t *n_terms, int64_t *n_groups,
                  int64_t *n_diag_terms, int *device) {
    if (!p) return fail(OFB_ERR_INVALID, "plan is NULL");
    if (n_qubits) *n_qubits = p->n_qubits;
    if (n_terms) *n_terms = p->n_terms;
    if (n_groups) *n_groups = p->n_groups;
    if (n_diag_terms) *n_diag_terms = p->n_diag_terms;
    if (device) *device = p->device;
    return OFB_OK;
}

int ofb_plan_group(const ofb_plan *p, int64_t g, uint64_t *x_mask, int64_t *n_terms) {
    if (!p) return fail(OFB_ERR_INVALID, "plan is NULL");
    if (g < 0 || g >= p->n_groups) return fail(OFB_ERR_INVALID, "group index out of range");
    if (x_mask) *x_mask = p->h_group_x[g];
    if (n_terms) *n_terms = (int64_t)p->h_group_begin[g + 1] - p->h_group_begin[g];
    return OFB_OK;
}

}  // extern "C"
