// bfe_host.cu -- host-buffer entry points of the EOF passes (the calls the reference-facing API makes).
//
// The reference's functions take and return HOST arrays (eof.accumulate eof.py:492, eof.accumulated_eval_particles
// eof.py:989).  A monolithic copy-in -> kernels -> copy-out leaves the PCIe link idle in one direction and the GPU
// idle during both copies, so these entry points cut the particle set into chunks and run a three-stage pipeline
//
//     H2D of chunk k+1 (stream s_in)  |  kernels of chunk k (caller's stream)  |  D2H of chunk k-1 (stream s_out)
//
// with double-buffered device staging owned by the handle.  All copies and launches are issued from this one
// C call (a few microseconds each), which keeps the issue rate far above the DMA rate: the same pipeline driven
// from Python spent longer issuing its ~50 small copies than the copies took (profiles/r01_summary.md section 12).
// Pinned host memory gives asynchronous DMA at 40-55 GB/s per direction (profiles/pcie_probe.py); pageable
// memory is accepted and simply copies synchronously.
//
// One upload per snapshot (option "host_reuse", default 1): bfe_eof_accumulate_host lands its chunks in a device buffer
// that is KEPT (process-wide, one per device); a following bfe_eof_force_host on the same host arrays -- same three
// pointers, same length, same content tag -- evaluates from that copy and skips its own 24 B/particle upload, which also
// lets its result copies run at the one-directional PCIe rate (55 instead of 40 GB/s when both directions are busy).
// The tag is a hash of the first and last 64 values and of 1024 evenly spaced values of each array: an in-place change of
// the whole set (a rotation, a drift step) is always seen, an edit of a few isolated particles between the two calls
// may not be -- callers that do that set the option to 0.  bfe_eof_accumulate_host itself ALWAYS uploads.
#include "bfe_internal.h"
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string.h>
#include <thread>
#include <vector>

struct BfeHostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    int64_t cap = 0;                 // particles per staging buffer
    double* din[2] = {nullptr, nullptr};     // [4][cap]
    double* dout[2] = {nullptr, nullptr};    // [6][cap]
    double* coef = nullptr;                  // coefficient scratch (per-chunk block, running total)
    size_t coef_cap = 0;
    cudaEvent_t ev_in[2], ev_cmp[2], ev_out[2], ev_start;
    bool events = false;
    // pinned bounce buffers for PAGEABLE host inputs ([4][cap] each), filled by the copy threads below
    double* hstage[2] = {nullptr, nullptr};
    int64_t hstage_cap = 0;
    cudaEvent_t ev_stage[2];
    bool stage_events = false;
};

// ---- pageable inputs.  A drop-in caller passes plain NumPy arrays; cudaMemcpyAsync from pageable memory is a synchronous,
// single-threaded bounce through the driver's staging (~9 GB/s: 3.5 ms for the 32 MB of 10^6 particles against 0.6 ms from
// pinned memory, profiles/r02_e2e_probe.py).  Here a small pool of copy threads fills a pinned bounce buffer of the NEXT chunk
// (memcpy of the rows in slices, all threads in parallel) while the DMA of the previous chunk runs (option "host_threads").
int g_bfe_host_threads = 4;          // option "host_threads": copy threads for pageable inputs (0: leave it to cudaMemcpyAsync)

class BfeCopyPool {
public:
    struct Job { char* dst; const char* src; size_t bytes; };
    explicit BfeCopyPool(int n) : stop_(false), pending_(0), epoch_(0) {
        for (int i = 0; i < n; ++i) threads_.emplace_back([this, i] { run(i); });
    }
    ~BfeCopyPool() {
        { std::lock_guard<std::mutex> l(mu_); stop_ = true; ++epoch_; }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int size() const { return (int)threads_.size(); }
    // copy all jobs; the calling thread takes its share and returns when every slice is done
    void run_all(const std::vector<Job>& jobs) {
        {
            std::lock_guard<std::mutex> l(mu_);
            jobs_ = &jobs; next_ = 0; pending_ = (int)jobs.size(); ++epoch_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [this] { return pending_ == 0; });
        jobs_ = nullptr;
    }
private:
    void work() {
        for (;;) {
            const Job* j = nullptr;
            {
                std::lock_guard<std::mutex> l(mu_);
                if (!jobs_ || next_ >= jobs_->size()) return;
                j = &(*jobs_)[next_++];
            }
            memcpy(j->dst, j->src, j->bytes);
            std::lock_guard<std::mutex> l(mu_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    void run(int) {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
            }
            work();
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::vector<Job>* jobs_ = nullptr;
    size_t next_ = 0;
    bool stop_;
    int pending_;
    unsigned long long epoch_;
};
static BfeCopyPool* g_copy_pool = nullptr;       // made on first use, lives for the process (threads sleep on a condition variable)
static std::mutex g_copy_pool_mu;

static bool host_is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

int g_bfe_host_chunk = 0;            // option "host_chunk": particles per pipeline chunk (0 = auto)
int g_bfe_host_reused_last = 0;      // option (read-only) "host_reused_last": 1 if the latest bfe_eof_force_host evaluated from the kept copy
int g_bfe_host_reuse = 1;            // option "host_reuse": force_host may evaluate from the copy accumulate_host uploaded

// the particle set the last bfe_eof_accumulate_host call on this device uploaded: rows x, y, z, m of `cap` doubles
struct BfeKeepSet {
    int device = -1;
    double* d = nullptr;
    int64_t cap = 0, n = 0;
    const double* hp[3] = {nullptr, nullptr, nullptr};
    uint64_t tag[3] = {0, 0, 0};
    bool valid = false;
    cudaEvent_t ready = nullptr, released = nullptr;      // upload complete / last reader done
    bool has_reader = false;
};
static BfeKeepSet g_keep[16][2];            // [device][0 = EOF (disc) set, 1 = SL (halo) set]
static std::mutex g_keep_mu;

static uint64_t host_tag(const double* a, int64_t n) {
    uint64_t h = 1469598103934665603ull ^ (uint64_t)n;
    auto mix = [&](int64_t i) { uint64_t v; memcpy(&v, a + i, 8); h = (h ^ v) * 1099511628211ull; h ^= h >> 29; };
    const int64_t edge = std::min<int64_t>(64, n);
    for (int64_t i = 0; i < edge; ++i) mix(i);
    for (int64_t i = std::max<int64_t>(edge, n - 64); i < n; ++i) mix(i);
    if (n > 128) {
        const int64_t step = std::max<int64_t>(1, (n - 128) / 1024);
        for (int64_t i = 64; i < n - 64; i += step) mix(i);
    }
    return h;
}

static BfeKeepSet* keep_slot(int kind) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    BfeKeepSet* k = &g_keep[dev][kind];
    if (k->device != dev) {
        k->device = dev;
        if (cudaEventCreateWithFlags(&k->ready, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&k->released, cudaEventDisableTiming) != cudaSuccess) { k->device = -1; return nullptr; }
    }
    return k;
}

__global__ void bfe_coef_add_kernel(double* __restrict__ a, const double* __restrict__ b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}

void bfe_host_pipe_destroy(void* p_) {
    BfeHostPipe* p = (BfeHostPipe*)p_;
    if (!p) return;
    for (int b = 0; b < 2; ++b) { if (p->din[b]) cudaFree(p->din[b]); if (p->dout[b]) cudaFree(p->dout[b]); }
    if (p->coef) cudaFree(p->coef);
    if (p->events) {
        for (int b = 0; b < 2; ++b) { cudaEventDestroy(p->ev_in[b]); cudaEventDestroy(p->ev_cmp[b]); cudaEventDestroy(p->ev_out[b]); }
        cudaEventDestroy(p->ev_start);
    }
    for (int b = 0; b < 2; ++b) if (p->hstage[b]) cudaFreeHost(p->hstage[b]);
    if (p->stage_events) for (int b = 0; b < 2; ++b) cudaEventDestroy(p->ev_stage[b]);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    delete p;
}

static int64_t pick_chunk(int64_t n) {
    if (g_bfe_host_chunk > 0) return ((int64_t)g_bfe_host_chunk + 15) / 16 * 16;
    // ~250 k particles per chunk, <= 8 chunks (profiles/e2e_probe.py on 10^6 particles: one shot 2.30 ms, 500 k 2.00,
    // 250 k 1.95, 125 k 2.07 -- smaller chunks pay the fixed kernel costs of a chunk more often than they hide copies)
    int64_t k = (n + 125000) / 250000;
    k = std::max<int64_t>(1, std::min<int64_t>(8, k));
    return ((n + k - 1) / k + 15) / 16 * 16;
}

static int pipe_get(void** slot, size_t coef_doubles, int64_t chunk, bool need_out, BfeHostPipe** out) {
    BfeHostPipe* p = (BfeHostPipe*)*slot;
    if (!p) {
        p = new BfeHostPipe();
        *slot = p;
        BFE_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
        BFE_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            BFE_CUDA(cudaEventCreateWithFlags(&p->ev_in[b], cudaEventDisableTiming));
            BFE_CUDA(cudaEventCreateWithFlags(&p->ev_cmp[b], cudaEventDisableTiming));
            BFE_CUDA(cudaEventCreateWithFlags(&p->ev_out[b], cudaEventDisableTiming));
        }
        BFE_CUDA(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
        p->events = true;
    }
    if (coef_doubles > p->coef_cap) {
        if (p->coef) { BFE_CUDA(cudaDeviceSynchronize()); cudaFree(p->coef); p->coef = nullptr; }
        BFE_CUDA(cudaMalloc(&p->coef, coef_doubles * sizeof(double)));
        p->coef_cap = coef_doubles;
    }
    if (chunk > p->cap) {
        BFE_CUDA(cudaDeviceSynchronize());
        for (int b = 0; b < 2; ++b) {
            if (p->din[b]) { cudaFree(p->din[b]); p->din[b] = nullptr; }
            if (p->dout[b]) { cudaFree(p->dout[b]); p->dout[b] = nullptr; }
        }
        p->cap = chunk;
        for (int b = 0; b < 2; ++b) BFE_CUDA(cudaMalloc(&p->din[b], 4 * (size_t)chunk * sizeof(double)));
    }
    if (need_out)
        for (int b = 0; b < 2; ++b)
            if (!p->dout[b]) BFE_CUDA(cudaMalloc(&p->dout[b], 6 * (size_t)p->cap * sizeof(double)));
    *out = p;
    return BFE_OK;
}

// rows[k] + lo, len doubles each -> dev + k * pitch   (one 2-D DMA when the host rows are equally spaced)
static const void* g_no2d_h2d = nullptr;     // rows[0] of the last host set whose 2-D DMA the driver refused
static const void* g_no2d_d2h = nullptr;

static cudaError_t copy_rows_h2d(double* dev, int64_t pitch, const double* const* rows, int nrows, int64_t lo,
                                 int64_t len, cudaStream_t s) {
    bool even = nrows > 1 && g_no2d_h2d != (const void*)rows[0];
    const ptrdiff_t step = nrows > 1 ? (rows[1] - rows[0]) : 0;
    for (int k = 2; k < nrows && even; ++k) even = (rows[k] - rows[k - 1]) == step;
    if (even && step >= len) {
        // one 2-D DMA.  The driver rejects it (invalid argument) when the span between the rows crosses separately pinned
        // allocations with unregistered gaps (small tensors of torch's pinned allocator; found by compute-sanitizer on a
        // 5003-particle set): fall back to one copy per row then.
        cudaError_t e = cudaMemcpy2DAsync(dev, (size_t)pitch * sizeof(double), rows[0] + lo, (size_t)step * sizeof(double),
                                          (size_t)len * sizeof(double), nrows, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) return e;
        cudaGetLastError();
        g_no2d_h2d = (const void*)rows[0];
    }
    for (int k = 0; k < nrows; ++k) {
        cudaError_t e = cudaMemcpyAsync(dev + (size_t)k * pitch, rows[k] + lo, (size_t)len * sizeof(double),
                                        cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static cudaError_t copy_rows_d2h(double* const* rows, int nrows, int64_t lo, const double* dev, int64_t pitch,
                                 int64_t len, cudaStream_t s) {
    bool even = nrows > 1 && g_no2d_d2h != (const void*)rows[0];
    const ptrdiff_t step = nrows > 1 ? (rows[1] - rows[0]) : 0;
    for (int k = 2; k < nrows && even; ++k) even = (rows[k] - rows[k - 1]) == step;
    if (even && step >= len) {
        cudaError_t e = cudaMemcpy2DAsync(rows[0] + lo, (size_t)step * sizeof(double), dev, (size_t)pitch * sizeof(double),
                                          (size_t)len * sizeof(double), nrows, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) return e;
        cudaGetLastError();                      // see copy_rows_h2d
        g_no2d_d2h = (const void*)rows[0];
    }
    for (int k = 0; k < nrows; ++k) {
        cudaError_t e = cudaMemcpyAsync(rows[k] + lo, dev + (size_t)k * pitch, (size_t)len * sizeof(double),
                                        cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Upload `nrows` host rows [lo, lo + len) to dev + k * pitch on stream s.  Pinned rows go as one 2-D DMA (copy_rows_h2d);
// pageable rows are first copied into the pipe's pinned bounce buffer `b` by the copy threads, then sent by DMA.
static int upload_rows(BfeHostPipe* p, int b, bool pageable, double* dev, int64_t pitch, const double* const* rows, int nrows,
                       int64_t lo, int64_t len, cudaStream_t s) {
    if (!pageable || g_bfe_host_threads <= 0) {
        BFE_CUDA(copy_rows_h2d(dev, pitch, rows, nrows, lo, len, s));
        return BFE_OK;
    }
    if (p->hstage_cap < len) {
        BFE_CUDA(cudaDeviceSynchronize());
        for (int q = 0; q < 2; ++q) { if (p->hstage[q]) cudaFreeHost(p->hstage[q]); p->hstage[q] = nullptr; }
        for (int q = 0; q < 2; ++q) BFE_CUDA(cudaHostAlloc((void**)&p->hstage[q], 4 * (size_t)len * sizeof(double), cudaHostAllocDefault));
        p->hstage_cap = len;
    }
    if (!p->stage_events) {
        for (int q = 0; q < 2; ++q) BFE_CUDA(cudaEventCreateWithFlags(&p->ev_stage[q], cudaEventDisableTiming));
        p->stage_events = true;
        for (int q = 0; q < 2; ++q) BFE_CUDA(cudaEventRecord(p->ev_stage[q], s));
    }
    BFE_CUDA(cudaEventSynchronize(p->ev_stage[b]));               // the DMA that last read this bounce buffer is done
    {
        std::lock_guard<std::mutex> l(g_copy_pool_mu);
        if (!g_copy_pool || g_copy_pool->size() != g_bfe_host_threads - 1) {
            delete g_copy_pool;
            g_copy_pool = new BfeCopyPool(std::max(0, g_bfe_host_threads - 1));
        }
        std::vector<BfeCopyPool::Job> jobs;
        const int64_t slice = 32768;                              // 256 kB per job
        for (int k = 0; k < nrows; ++k)
            for (int64_t o = 0; o < len; o += slice) {
                const int64_t m = std::min(slice, len - o);
                jobs.push_back({(char*)(p->hstage[b] + (size_t)k * p->hstage_cap + o), (const char*)(rows[k] + lo + o),
                                (size_t)m * sizeof(double)});
            }
        g_copy_pool->run_all(jobs);
    }
    BFE_CUDA(cudaMemcpy2DAsync(dev, (size_t)pitch * sizeof(double), p->hstage[b], (size_t)p->hstage_cap * sizeof(double),
                               (size_t)len * sizeof(double), nrows, cudaMemcpyHostToDevice, s));
    BFE_CUDA(cudaEventRecord(p->ev_stage[b], s));
    return BFE_OK;
}

// Accumulation on HOST particle arrays (rows x, y, z, m); `run(len, d, pitch, dst)` enqueues the accumulation kernels of
// one chunk whose rows start at d, d + pitch, ... and leaves its ncoef coefficients at dst.  The coefficients stay on the
// device (out_dev) so that a multi-GPU caller can allreduce them before the one small copy out.
template <class Run>
static int accumulate_host_impl(void** pipe_slot, int kind, int64_t n, const double* const rows[4], int ncoef,
                                double* out_dev, cudaStream_t stream, Run run) {
    const int64_t chunk = pick_chunk(n);
    BfeHostPipe* p = nullptr;
    int rc = pipe_get(pipe_slot, (size_t)ncoef, chunk, false, &p);
    if (rc != BFE_OK) return rc;
    BFE_CUDA(cudaEventRecord(p->ev_start, stream));
    BFE_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_start, 0));
    // the upload lands in the kept buffer (full size: no double buffering needed) unless reuse is off or the set is huge
    std::unique_lock<std::mutex> lock(g_keep_mu);
    BfeKeepSet* ks = (g_bfe_host_reuse && n <= ((int64_t)1 << 27)) ? keep_slot(kind) : nullptr;
    if (ks) {
        ks->valid = false;
        if (ks->cap < n) {
            if (ks->d) { BFE_CUDA(cudaDeviceSynchronize()); cudaFree(ks->d); ks->d = nullptr; ks->cap = 0; }
            const int64_t cap = (n + n / 8 + 1023) / 16 * 16;
            if (cudaMalloc(&ks->d, 4 * (size_t)cap * sizeof(double)) != cudaSuccess) { cudaGetLastError(); ks = nullptr; }
            else ks->cap = cap;
        }
    }
    if (ks) {
        if (ks->has_reader) BFE_CUDA(cudaStreamWaitEvent(p->s_in, ks->released, 0));      // the last reader of the old set is done
        for (int q = 0; q < 3; ++q) { ks->hp[q] = rows[q]; ks->tag[q] = host_tag(rows[q], n); }
        ks->n = n;
    }
    const bool pageable = host_is_pageable(rows[0]);
    int k = 0;
    for (int64_t lo = 0; lo < n; lo += chunk, ++k) {
        const int b = k & 1;
        const int64_t len = std::min(chunk, n - lo);
        double* d;
        int64_t pitch;
        if (ks) { d = ks->d + lo; pitch = ks->cap; }
        else {
            if (k >= 2) BFE_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_cmp[b], 0));  // kernels of chunk k-2 have read din[b]
            d = p->din[b]; pitch = p->cap;
        }
        rc = upload_rows(p, b, pageable, d, pitch, rows, 4, lo, len, p->s_in);
        if (rc != BFE_OK) return rc;
        BFE_CUDA(cudaEventRecord(p->ev_in[b], p->s_in));
        BFE_CUDA(cudaStreamWaitEvent(stream, p->ev_in[b], 0));
        rc = run(len, d, pitch, k == 0 ? out_dev : p->coef);
        if (rc != BFE_OK) return rc;
        if (k > 0) {
            bfe_coef_add_kernel<<<(ncoef + 255) / 256, 256, 0, stream>>>(out_dev, p->coef, ncoef);
            BFE_LAUNCH_CHECK("bfe_coef_add_kernel");
        }
        BFE_CUDA(cudaEventRecord(p->ev_cmp[b], stream));
    }
    if (ks) {
        BFE_CUDA(cudaEventRecord(ks->ready, stream));      // after the last chunk's kernels (which waited for its upload)
        ks->has_reader = false;
        ks->valid = true;
    }
    return BFE_OK;
}

// eof.accumulate (eof.py:492) on host arrays: cos_out / sin_out are DEVICE pointers and need not be adjacent
extern "C" int bfe_eof_accumulate_host(bfe_eof* h, int64_t n, const double* hx, const double* hy, const double* hz,
                                       const double* hm, double* cos_out, double* sin_out, void* stream_) {
    if (!h || n < 0 || !cos_out || !sin_out) return BFE_ERR_ARG;
    if (n > 0 && (!hx || !hy || !hz || !hm)) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int ncoef = (h->g.mmax + 1) * h->g.norder;
    if (n == 0) return bfe_eof_accumulate(h, 0, nullptr, nullptr, nullptr, nullptr, cos_out, sin_out, stream_);
    const double* rows[4] = {hx, hy, hz, hm};
    // the pipeline sums chunk blocks of 2 * ncoef adjacent doubles: accumulate into a scratch pair, then copy out
    BfeHostPipe* p = nullptr;
    int rc = pipe_get(&h->host_pipe, 4 * (size_t)ncoef, pick_chunk(n), false, &p);
    if (rc != BFE_OK) return rc;
    double* total = p->coef + 2 * ncoef;
    rc = accumulate_host_impl(&h->host_pipe, 0, n, rows, 2 * ncoef, total, stream,   // (its pipe_get asks for 2 * ncoef <= the 4 * ncoef held)
                              [&](int64_t len, double* d, int64_t pitch, double* dst) {
                                  return bfe_eof_accumulate(h, len, d, d + pitch, d + 2 * pitch, d + 3 * pitch, dst, dst + ncoef, stream_);
                              });
    if (rc != BFE_OK) return rc;
    BFE_CUDA(cudaMemcpyAsync(cos_out, total, ncoef * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    BFE_CUDA(cudaMemcpyAsync(sin_out, total + ncoef, ncoef * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    return BFE_OK;
}

// spheresl.compute_coefficients_solitary (spheresl.py:567) on host arrays: expcoef is a DEVICE pointer
extern "C" int bfe_sl_accumulate_host(bfe_sl* h, int64_t n, const double* hx, const double* hy, const double* hz,
                                      const double* hm, int no_odd, double* expcoef, void* stream_) {
    if (!h || n < 0 || !expcoef) return BFE_ERR_ARG;
    if (n > 0 && (!hx || !hy || !hz || !hm)) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return bfe_sl_accumulate(h, 0, nullptr, nullptr, nullptr, nullptr, no_odd, expcoef, stream_);
    const int ncoef = h->g.nrow * h->g.nmax;
    const double* rows[4] = {hx, hy, hz, hm};
    return accumulate_host_impl(&h->host_pipe, 1, n, rows, ncoef, expcoef, stream,
                                [&](int64_t len, double* d, int64_t pitch, double* dst) {
                                    return bfe_sl_accumulate(h, len, d, d + pitch, d + 2 * pitch, d + 3 * pitch, no_odd, dst, stream_);
                                });
}

// Evaluation on HOST arrays in (x, y, z) and out (six rows); `run(len, d, pitch, o, opitch)` enqueues the evaluation of one
// chunk.  On return all work is enqueued; the six host arrays are complete once `stream` has been synchronised.
template <class Run>
static int force_host_impl(void** pipe_slot, int kind, int64_t n, const double* const rows[3], double* const orow[6],
                           cudaStream_t stream, Run run) {
    const int64_t chunk = pick_chunk(n);
    BfeHostPipe* p = nullptr;
    int rc = pipe_get(pipe_slot, 8, chunk, true, &p);
    if (rc != BFE_OK) return rc;
    BFE_CUDA(cudaEventRecord(p->ev_start, stream));
    BFE_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_start, 0));
    // the particles the accumulation uploaded last, if these are the same host arrays with the same content tag
    std::unique_lock<std::mutex> lock(g_keep_mu);
    BfeKeepSet* ks = g_bfe_host_reuse ? keep_slot(kind) : nullptr;
    if (ks && !(ks->valid && ks->n == n && ks->hp[0] == rows[0] && ks->hp[1] == rows[1] && ks->hp[2] == rows[2] &&
                ks->tag[0] == host_tag(rows[0], n) && ks->tag[1] == host_tag(rows[1], n) && ks->tag[2] == host_tag(rows[2], n)))
        ks = nullptr;
    if (ks) BFE_CUDA(cudaStreamWaitEvent(stream, ks->ready, 0));
    g_bfe_host_reused_last = ks ? 1 : 0;
    const bool pageable = !ks && host_is_pageable(rows[0]);
    int k = 0;
    for (int64_t lo = 0; lo < n; lo += chunk, ++k) {
        const int b = k & 1;
        const int64_t len = std::min(chunk, n - lo);
        double* d;
        int64_t pitch;
        if (ks) { d = ks->d + lo; pitch = ks->cap; }
        else {
            if (k >= 2) BFE_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_cmp[b], 0));  // kernels of chunk k-2 have read din[b]
            rc = upload_rows(p, b, pageable, p->din[b], p->cap, rows, 3, lo, len, p->s_in);
            if (rc != BFE_OK) return rc;
            BFE_CUDA(cudaEventRecord(p->ev_in[b], p->s_in));
            BFE_CUDA(cudaStreamWaitEvent(stream, p->ev_in[b], 0));
            d = p->din[b]; pitch = p->cap;
        }
        if (k >= 2) BFE_CUDA(cudaStreamWaitEvent(stream, p->ev_out[b], 0));       // chunk k-2 has left dout[b]
        double* o = p->dout[b];
        rc = run(len, d, pitch, o, p->cap);
        if (rc != BFE_OK) return rc;
        BFE_CUDA(cudaEventRecord(p->ev_cmp[b], stream));
        BFE_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_cmp[b], 0));
        BFE_CUDA(copy_rows_d2h(orow, 6, lo, o, p->cap, len, p->s_out));
        BFE_CUDA(cudaEventRecord(p->ev_out[b], p->s_out));
    }
    if (ks) { BFE_CUDA(cudaEventRecord(ks->released, stream)); ks->has_reader = true; }
    // the caller's stream completes only after the last copies out
    BFE_CUDA(cudaStreamWaitEvent(stream, p->ev_out[(k - 1) & 1], 0));
    if (k >= 2) BFE_CUDA(cudaStreamWaitEvent(stream, p->ev_out[k & 1], 0));
    return BFE_OK;
}

// eof.accumulated_eval_particles (eof.py:989) on host arrays (uses the held contraction)
extern "C" int bfe_eof_force_host(bfe_eof* h, int64_t n, const double* hx, const double* hy, const double* hz,
                                  double* hp0, double* hp, double* hfr, double* hfp, double* hfz, double* hR,
                                  void* stream_) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (!h->contracted) return BFE_ERR_STATE;
    if (n == 0) return BFE_OK;
    if (!hx || !hy || !hz || !hp0 || !hp || !hfr || !hfp || !hfz || !hR) return BFE_ERR_ARG;
    const double* rows[3] = {hx, hy, hz};
    double* orow[6] = {hp0, hp, hfr, hfp, hfz, hR};
    return force_host_impl(&h->host_pipe, 0, n, rows, orow, (cudaStream_t)stream_,
                           [&](int64_t len, double* d, int64_t pitch, double* o, int64_t op) {
                               return bfe_eof_force_contracted(h, len, d, d + pitch, d + 2 * pitch, o, o + op, o + 2 * op,
                                                               o + 3 * op, o + 4 * op, o + 5 * op, stream_);
                           });
}

// spheresl.all_eval_particles (spheresl.py:1240) potential outputs on host arrays (uses the held contraction)
extern "C" int bfe_sl_force_host(bfe_sl* h, int64_t n, const double* hx, const double* hy, const double* hz,
                                 double* hpot0, double* hpot1, double* hpotr, double* hpott, double* hpotp, double* hrr,
                                 void* stream_) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (!h->contracted) return BFE_ERR_STATE;
    if (n == 0) return BFE_OK;
    if (!hx || !hy || !hz || !hpot0 || !hpot1 || !hpotr || !hpott || !hpotp || !hrr) return BFE_ERR_ARG;
    const double* rows[3] = {hx, hy, hz};
    double* orow[6] = {hpot0, hpot1, hpotr, hpott, hpotp, hrr};
    return force_host_impl(&h->host_pipe, 1, n, rows, orow, (cudaStream_t)stream_,
                           [&](int64_t len, double* d, int64_t pitch, double* o, int64_t op) {
                               return bfe_sl_force_contracted(h, len, d, d + pitch, d + 2 * pitch, o, o + op, o + 2 * op,
                                                              o + 3 * op, o + 4 * op, o + 5 * op, stream_);
                           });
}
