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
#include "bfe_internal.h"
#include <algorithm>

struct BfeHostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    int64_t cap = 0;                 // particles per staging buffer
    double* din[2] = {nullptr, nullptr};     // [4][cap]
    double* dout[2] = {nullptr, nullptr};    // [6][cap]
    double* coef = nullptr;                  // per-chunk coefficient scratch, 2 * (mmax+1) * norder
    cudaEvent_t ev_in[2], ev_cmp[2], ev_out[2], ev_start;
    bool events = false;
};

int g_bfe_host_chunk = 0;            // option "host_chunk": particles per pipeline chunk (0 = auto)

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

static int pipe_get(bfe_eof* h, int64_t chunk, bool need_out, BfeHostPipe** out) {
    BfeHostPipe* p = (BfeHostPipe*)h->host_pipe;
    if (!p) {
        p = new BfeHostPipe();
        h->host_pipe = p;
        BFE_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
        BFE_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            BFE_CUDA(cudaEventCreateWithFlags(&p->ev_in[b], cudaEventDisableTiming));
            BFE_CUDA(cudaEventCreateWithFlags(&p->ev_cmp[b], cudaEventDisableTiming));
            BFE_CUDA(cudaEventCreateWithFlags(&p->ev_out[b], cudaEventDisableTiming));
        }
        BFE_CUDA(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
        p->events = true;
        BFE_CUDA(cudaMalloc(&p->coef, 2 * (size_t)(h->g.mmax + 1) * h->g.norder * sizeof(double)));
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
static cudaError_t copy_rows_h2d(double* dev, int64_t pitch, const double* const* rows, int nrows, int64_t lo,
                                 int64_t len, cudaStream_t s) {
    bool even = nrows > 1;
    const ptrdiff_t step = nrows > 1 ? (rows[1] - rows[0]) : 0;
    for (int k = 2; k < nrows && even; ++k) even = (rows[k] - rows[k - 1]) == step;
    if (even && step >= len)
        return cudaMemcpy2DAsync(dev, (size_t)pitch * sizeof(double), rows[0] + lo, (size_t)step * sizeof(double),
                                 (size_t)len * sizeof(double), nrows, cudaMemcpyHostToDevice, s);
    for (int k = 0; k < nrows; ++k) {
        cudaError_t e = cudaMemcpyAsync(dev + (size_t)k * pitch, rows[k] + lo, (size_t)len * sizeof(double),
                                        cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static cudaError_t copy_rows_d2h(double* const* rows, int nrows, int64_t lo, const double* dev, int64_t pitch,
                                 int64_t len, cudaStream_t s) {
    bool even = nrows > 1;
    const ptrdiff_t step = nrows > 1 ? (rows[1] - rows[0]) : 0;
    for (int k = 2; k < nrows && even; ++k) even = (rows[k] - rows[k - 1]) == step;
    if (even && step >= len)
        return cudaMemcpy2DAsync(rows[0] + lo, (size_t)step * sizeof(double), dev, (size_t)pitch * sizeof(double),
                                 (size_t)len * sizeof(double), nrows, cudaMemcpyDeviceToHost, s);
    for (int k = 0; k < nrows; ++k) {
        cudaError_t e = cudaMemcpyAsync(rows[k] + lo, dev + (size_t)k * pitch, (size_t)len * sizeof(double),
                                        cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// eof.accumulate on HOST particle arrays; the coefficients stay on the device (cos_out / sin_out are DEVICE
// pointers) so that a multi-GPU caller can allreduce them before the one small copy out.
extern "C" int bfe_eof_accumulate_host(bfe_eof* h, int64_t n, const double* hx, const double* hy, const double* hz,
                                       const double* hm, double* cos_out, double* sin_out, void* stream_) {
    if (!h || n < 0 || !cos_out || !sin_out) return BFE_ERR_ARG;
    if (n > 0 && (!hx || !hy || !hz || !hm)) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int ncoef = (h->g.mmax + 1) * h->g.norder;
    if (n == 0) return bfe_eof_accumulate(h, 0, nullptr, nullptr, nullptr, nullptr, cos_out, sin_out, stream_);
    const int64_t chunk = pick_chunk(n);
    BfeHostPipe* p = nullptr;
    int rc = pipe_get(h, chunk, false, &p);
    if (rc != BFE_OK) return rc;
    const double* rows[4] = {hx, hy, hz, hm};
    BFE_CUDA(cudaEventRecord(p->ev_start, stream));
    BFE_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_start, 0));
    int k = 0;
    for (int64_t lo = 0; lo < n; lo += chunk, ++k) {
        const int b = k & 1;
        const int64_t len = std::min(chunk, n - lo);
        if (k >= 2) BFE_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_cmp[b], 0));      // kernels of chunk k-2 have read din[b]
        BFE_CUDA(copy_rows_h2d(p->din[b], p->cap, rows, 4, lo, len, p->s_in));
        BFE_CUDA(cudaEventRecord(p->ev_in[b], p->s_in));
        BFE_CUDA(cudaStreamWaitEvent(stream, p->ev_in[b], 0));
        double* d = p->din[b];
        double* c = k == 0 ? cos_out : p->coef;
        double* s = k == 0 ? sin_out : p->coef + ncoef;
        rc = bfe_eof_accumulate(h, len, d, d + p->cap, d + 2 * p->cap, d + 3 * p->cap, c, s, stream_);
        if (rc != BFE_OK) return rc;
        if (k > 0) {
            bfe_coef_add_kernel<<<(ncoef + 255) / 256, 256, 0, stream>>>(cos_out, p->coef, ncoef);
            bfe_coef_add_kernel<<<(ncoef + 255) / 256, 256, 0, stream>>>(sin_out, p->coef + ncoef, ncoef);
            BFE_LAUNCH_CHECK("bfe_coef_add_kernel");
        }
        BFE_CUDA(cudaEventRecord(p->ev_cmp[b], stream));
    }
    return BFE_OK;
}

// eof.accumulated_eval_particles on HOST arrays in and out (uses the held contraction).  On return all work is
// enqueued; the six host arrays are complete once `stream` has been synchronised.
extern "C" int bfe_eof_force_host(bfe_eof* h, int64_t n, const double* hx, const double* hy, const double* hz,
                                  double* hp0, double* hp, double* hfr, double* hfp, double* hfz, double* hR,
                                  void* stream_) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (!h->contracted) return BFE_ERR_STATE;
    if (n == 0) return BFE_OK;
    if (!hx || !hy || !hz || !hp0 || !hp || !hfr || !hfp || !hfz || !hR) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t chunk = pick_chunk(n);
    BfeHostPipe* p = nullptr;
    int rc = pipe_get(h, chunk, true, &p);
    if (rc != BFE_OK) return rc;
    const double* rows[3] = {hx, hy, hz};
    double* orow[6] = {hp0, hp, hfr, hfp, hfz, hR};
    BFE_CUDA(cudaEventRecord(p->ev_start, stream));
    BFE_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_start, 0));
    int k = 0;
    for (int64_t lo = 0; lo < n; lo += chunk, ++k) {
        const int b = k & 1;
        const int64_t len = std::min(chunk, n - lo);
        if (k >= 2) BFE_CUDA(cudaStreamWaitEvent(p->s_in, p->ev_cmp[b], 0));      // kernels of chunk k-2 have read din[b]
        BFE_CUDA(copy_rows_h2d(p->din[b], p->cap, rows, 3, lo, len, p->s_in));
        BFE_CUDA(cudaEventRecord(p->ev_in[b], p->s_in));
        BFE_CUDA(cudaStreamWaitEvent(stream, p->ev_in[b], 0));
        if (k >= 2) BFE_CUDA(cudaStreamWaitEvent(stream, p->ev_out[b], 0));       // chunk k-2 has left dout[b]
        double* d = p->din[b];
        double* o = p->dout[b];
        rc = bfe_eof_force_contracted(h, len, d, d + p->cap, d + 2 * p->cap, o, o + p->cap, o + 2 * p->cap,
                                      o + 3 * p->cap, o + 4 * p->cap, o + 5 * p->cap, stream_);
        if (rc != BFE_OK) return rc;
        BFE_CUDA(cudaEventRecord(p->ev_cmp[b], stream));
        BFE_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_cmp[b], 0));
        BFE_CUDA(copy_rows_d2h(orow, 6, lo, o, p->cap, len, p->s_out));
        BFE_CUDA(cudaEventRecord(p->ev_out[b], p->s_out));
    }
    // the caller's stream completes only after the last copies out
    BFE_CUDA(cudaStreamWaitEvent(stream, p->ev_out[(k - 1) & 1], 0));
    if (k >= 2) BFE_CUDA(cudaStreamWaitEvent(stream, p->ev_out[k & 1], 0));
    return BFE_OK;
}
