// Host <-> device copies of the *_host entry points (the drop-in calls hand over ordinary pageable
// numpy memory: LinearQubitOperator._matvec, linalg/linear_qubit_operator.py:74-148, and the CSC
// arrays of qubit_operator_sparse, linalg/sparse_tools.py:136-194).
//
// cudaMemcpy on pageable memory stages through ONE driver-owned pinned buffer filled by the calling
// thread: 11 GB/s host -> device, 4.7 GB/s device -> host into freshly allocated pages (the first
// touch of every destination page happens inside the copy), and the driver serialises such copies
// issued from several threads (measured on the B200 box, profiles/r02_d2h.json).  Here the staging
// is ours: a few worker threads, each with its own stream and two pinned buffers, move disjoint
// chunks -- the CPU side (memcpy between the caller's pages and pinned memory, page faults
// included) runs in parallel and overlaps the DMA of the other buffer.  The pinned buffers are
// allocated once per process (16 MB per worker, at most 16 workers) and reused.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>

#include "ofb_internal.cuh"

namespace ofb {

namespace {

constexpr size_t STAGE_BYTES = (size_t)8 << 20;  // per pinned buffer
constexpr int MAX_WORKERS = 16;

struct Worker {
    void *pin[2] = {nullptr, nullptr};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int device = -1;
};

Worker g_workers[MAX_WORKERS];
std::mutex g_copy_mu;  // one staged copy at a time per process

cudaError_t worker_ready(Worker &w, int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    if (w.device != device) {  // streams and events belong to a device
        if (w.stream) cudaStreamDestroy(w.stream);
        for (int b = 0; b < 2; ++b)
            if (w.ev[b]) cudaEventDestroy(w.ev[b]);
        w.stream = nullptr;
        w.ev[0] = w.ev[1] = nullptr;
        if ((e = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
        for (int b = 0; b < 2; ++b)
            if ((e = cudaEventCreateWithFlags(&w.ev[b], cudaEventDisableTiming)) != cudaSuccess) return e;
        w.device = device;
    }
    for (int b = 0; b < 2; ++b)
        if (!w.pin[b] && (e = cudaHostAlloc(&w.pin[b], STAGE_BYTES, cudaHostAllocPortable)) != cudaSuccess) return e;
    return cudaSuccess;
}

int worker_count(size_t bytes, int threads) {
    if (threads <= 0) {
        const char *env = getenv("OFB_COPY_THREADS");
        threads = env ? atoi(env) : 0;
    }
    // measured on the 16-core B200 box: 4.3 GB host -> device 0.41 s plain, 0.12 s with 8 workers,
    // 0.094 s with 16; 14.9 GB device -> host into fresh pages 3.6 s plain, 0.86 s / 0.52 s
    if (threads <= 0) threads = (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
    threads = std::min(threads, MAX_WORKERS);
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, (bytes + STAGE_BYTES - 1) / STAGE_BYTES));
}

}  // namespace

// Blocking.  `d` device pointer, `h` pageable (or any) host pointer.
int staged_copy(void *d, void *h, size_t bytes, bool to_device, int device, int threads) {
    if (bytes == 0) return OFB_OK;
    if (device < 0) OFB_CUDA(cudaGetDevice(&device));  // the calling thread's current device
    std::lock_guard<std::mutex> lock(g_copy_mu);
    const int T = worker_count(bytes, threads);
    const size_t n_chunks = (bytes + STAGE_BYTES - 1) / STAGE_BYTES;
    cudaError_t status[MAX_WORKERS];
    for (int t = 0; t < MAX_WORKERS; ++t) status[t] = cudaSuccess;
    auto run = [&](int t) {
        Worker &w = g_workers[t];
        cudaError_t e = worker_ready(w, device);
        auto span = [&](size_t c, size_t &off, size_t &len) {
            off = c * STAGE_BYTES;
            len = std::min(STAGE_BYTES, bytes - off);
        };
        size_t off, len;
        if (e == cudaSuccess && to_device) {
            int b = 0;
            bool used[2] = {false, false};
            for (size_t c = (size_t)t; c < n_chunks && e == cudaSuccess; c += (size_t)T, b ^= 1) {
                span(c, off, len);
                if (used[b]) e = cudaEventSynchronize(w.ev[b]);  // the DMA out of this buffer is done
                if (e != cudaSuccess) break;
                memcpy(w.pin[b], (const char *)h + off, len);
                e = cudaMemcpyAsync((char *)d + off, w.pin[b], len, cudaMemcpyHostToDevice, w.stream);
                if (e == cudaSuccess) e = cudaEventRecord(w.ev[b], w.stream);
                used[b] = true;
            }
        } else if (e == cudaSuccess) {
            size_t c = (size_t)t;
            int b = 0;
            if (c < n_chunks) {
                span(c, off, len);
                e = cudaMemcpyAsync(w.pin[0], (const char *)d + off, len, cudaMemcpyDeviceToHost, w.stream);
                if (e == cudaSuccess) e = cudaEventRecord(w.ev[0], w.stream);
            }
            for (; c < n_chunks && e == cudaSuccess; c += (size_t)T, b ^= 1) {
                const size_t next = c + (size_t)T;
                if (next < n_chunks) {  // the other buffer was drained in the previous iteration
                    span(next, off, len);
                    e = cudaMemcpyAsync(w.pin[b ^ 1], (const char *)d + off, len, cudaMemcpyDeviceToHost, w.stream);
                    if (e == cudaSuccess) e = cudaEventRecord(w.ev[b ^ 1], w.stream);
                    if (e != cudaSuccess) break;
                }
                e = cudaEventSynchronize(w.ev[b]);
                if (e != cudaSuccess) break;
                span(c, off, len);
                memcpy((char *)h + off, w.pin[b], len);
            }
        }
        if (w.stream) {
            cudaError_t e2 = cudaStreamSynchronize(w.stream);
            if (e == cudaSuccess) e = e2;
        }
        status[(size_t)t] = e;
    };
    if (T == 1) {
        run(0);
    } else {
        // No exception may leave this function (it is called from extern "C" entry points): a
        // worker whose thread cannot be created runs inline in the calling thread instead.
        std::thread pool[MAX_WORKERS];
        for (int t = 0; t < T; ++t) {
            try {
                pool[t] = std::thread(run, t);
            } catch (...) {
                run(t);
            }
        }
        for (int t = 0; t < T; ++t)
            if (pool[t].joinable()) pool[t].join();
    }
    for (int t = 0; t < T; ++t)
        if (status[t] != cudaSuccess) return cuda_fail(status[t], "staged host/device copy", __FILE__, __LINE__);
    return OFB_OK;
}

}  // namespace ofb

using namespace ofb;

extern "C" {

int ofb_copy_h2d_staged(void *d_dst, const void *h_src, size_t bytes, int device, int threads) {
    if (bytes && (!d_dst || !h_src)) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_TRY
    return staged_copy(d_dst, const_cast<void *>(h_src), bytes, true, device, threads);
    OFB_CATCH((void)0)
}

int ofb_copy_d2h_staged(void *h_dst, const void *d_src, size_t bytes, int device, int threads) {
    if (bytes && (!h_dst || !d_src)) return fail(OFB_ERR_INVALID, "NULL argument");
    OFB_TRY
    return staged_copy(const_cast<void *>(d_src), h_dst, bytes, false, device, threads);
    OFB_CATCH((void)0)
}

}  // extern "C"
