// bfe_ingest.cu -- the step BEFORE the path: pre-accumulation transforms of a
// device-resident snapshot, as Fields.total_coefficients applies them (exptool/basis/potential.py:120-226):
//
//   bfe_bar_fourier      m = 2 Fourier sums of the planar positions -> bar angle  (analysis/pattern.py:155-169)
//   bfe_affine_xy        planar rotation by the bar angle and/or centre shift     (pattern.py:118-139, potential.py:213-219)
//   bfe_inner_com        mass-weighted centre of the `ncenter` innermost particles (potential.py:158-176; the
//                        ranking arrays and the summed arrays are separate arguments because the reference ranks
//                        the DISC particles and applies the indices to the halo arrays, potential.py:190-200):
//                        the reference argsorts all radii; here the ncenter-th smallest radius is found by a
//                        radix select on the bit pattern of r (six 11-bit digit passes: integer shared-memory
//                        histograms, no sort, no floating-point atomics) and the sums run over r below it.
//
// All reductions are per-CTA partials in shared memory followed by a fixed-order reduce in the last CTA.
#include "bfe_device.cuh"
#include <mutex>

namespace {

constexpr int RED_THREADS = 256;
constexpr int SEL_BITS = 11, SEL_BINS = 1 << SEL_BITS, SEL_PASSES = 6;     // 66 >= 64 key bits

struct IngestWs {
    int device = -1;
    double* partial = nullptr;          // [max_ctas][4]
    unsigned int* counter = nullptr;    // last-CTA counter, tie counter
    unsigned int* hist = nullptr;       // [SEL_BINS]
    unsigned long long* sel = nullptr;  // [0] key prefix, [1] remaining rank, [2] threshold key, [3] ties wanted
    int max_ctas = 0;
};
IngestWs g_ws[16];
std::mutex g_ws_mutex;

int ingest_ws(IngestWs** out) {
    int dev = 0;
    BFE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) return BFE_ERR_UNSUPPORTED;
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    IngestWs& w = g_ws[dev];
    if (w.device != dev) {
        int sms = 0;
        BFE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        w.max_ctas = sms * 8;
        BFE_CUDA(cudaMalloc(&w.partial, (size_t)w.max_ctas * 4 * sizeof(double)));
        BFE_CUDA(cudaMalloc(&w.counter, 4 * sizeof(unsigned int)));
        BFE_CUDA(cudaMemset(w.counter, 0, 4 * sizeof(unsigned int)));
        BFE_CUDA(cudaMalloc(&w.hist, SEL_BINS * sizeof(unsigned int)));
        BFE_CUDA(cudaMemset(w.hist, 0, SEL_BINS * sizeof(unsigned int)));
        BFE_CUDA(cudaMalloc(&w.sel, 4 * sizeof(unsigned long long)));
        w.device = dev;
    }
    *out = &w;
    return BFE_OK;
}

int red_grid(int64_t n, int max_ctas) {
    int64_t need = (n + RED_THREADS * 4 - 1) / (RED_THREADS * 4);
    if (need < 1) need = 1;
    return (int)(need < max_ctas ? need : max_ctas);
}

// CTA tree sum of NV values per thread; thread 0 ends with the totals in v[]
template <int NV>
__device__ __forceinline__ void cta_sum(double (&v)[NV], double* s_red /*[NV][RED_THREADS/32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if (lane == 0) s_red[k * (RED_THREADS / 32) + warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0.0;
            for (int w = 0; w < RED_THREADS / 32; ++w) s += s_red[k * (RED_THREADS / 32) + w];
            v[k] = s;
        }
    }
}

// partial[cta][NV] written, last CTA sums the rows in order into out[NV]
template <int NV>
__device__ __forceinline__ void grid_finish(double (&v)[NV], double* __restrict__ partial, unsigned int* __restrict__ counter,
                                            double* __restrict__ out) {
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partial[(size_t)blockIdx.x * 4 + k] = v[k];
        __threadfence();
        s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < NV) {
        __threadfence();
        double s = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(partial + (size_t)b * 4 + threadIdx.x);
        out[threadIdx.x] = s;
        if (threadIdx.x == 0) *counter = 0u;
    }
}

// pattern.py:155-169: sums of cos 2phi and sin 2phi over minr < R < maxr, R = (x^2 + y^2)^0.5.
// cos 2phi = (x^2 - y^2) / R^2, sin 2phi = 2 x y / R^2 (phi = atan2(y, x); R = 0 is outside any window with minr >= 0).
__global__ void __launch_bounds__(RED_THREADS)
bar_fourier_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y, double minr, double maxr,
                   double* __restrict__ partial, unsigned int* __restrict__ counter, double* __restrict__ out) {
    __shared__ double s_red[2 * (RED_THREADS / 32)];
    double v[2] = {0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double px = __ldg(x + i), py = __ldg(y + i);
        const double r2 = px * px + py * py;
        const double r = sqrt(r2);
        if (r > minr && r < maxr) {
            const double inv = 1.0 / r2;
            v[0] += (px * px - py * py) * inv;
            v[1] += 2.0 * px * py * inv;
        }
    }
    cta_sum<2>(v, s_red);
    grid_finish<2>(v, partial, counter, out);
}

// (x, y) -> (x cos a - y sin a - cx, x sin a + y cos a - cy), z -> z - cz  (z arrays optional)
__global__ void __launch_bounds__(256)
affine_kernel(int64_t n, double ca, double sa, double cx, double cy, double cz, const double* __restrict__ x,
              const double* __restrict__ y, const double* __restrict__ z, double* __restrict__ xo,
              double* __restrict__ yo, double* __restrict__ zo) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double px = x[i], py = y[i];
        // pattern.py:118-121 operation order: x cos - y sin, x sin + y cos (no fused multiply-add: the reference
        // rounds each product)
        xo[i] = __dsub_rn(__dsub_rn(__dmul_rn(px, ca), __dmul_rn(py, sa)), cx);
        yo[i] = __dsub_rn(__dadd_rn(__dmul_rn(px, sa), __dmul_rn(py, ca)), cy);
        if (z) zo[i] = z[i] - cz;
    }
}

__device__ __forceinline__ unsigned long long radius_key(double px, double py, double pz) {
    // potential.py:158: (x*x + y*y + z*z)**0.5 ; non-negative doubles order like their bit patterns
    const double r = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz)));
    return (unsigned long long)__double_as_longlong(r);
}

// histogram of digit `pass` (most significant first) among the keys whose higher digits equal sel[0]
__global__ void __launch_bounds__(RED_THREADS)
select_hist_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                   int pass, const unsigned long long* __restrict__ sel, unsigned int* __restrict__ hist) {
    __shared__ unsigned int s_h[SEL_BINS];
    for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x) s_h[b] = 0u;
    __syncthreads();
    // the 64-bit key is read as a 66-bit number (two leading zero bits): pass p covers bits [55 - 11 p, 66 - 11 p)
    const int shift = 55 - SEL_BITS * pass;
    const unsigned long long prefix = sel[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long k = radius_key(__ldg(x + i), __ldg(y + i), __ldg(z + i));
        const unsigned long long hi = (pass == 0) ? 0ull : (k >> (shift + SEL_BITS));
        if (hi == prefix) atomicAdd(&s_h[(unsigned int)((k >> shift) & (SEL_BINS - 1))], 1u);      // integer counters
    }
    __syncthreads();
    for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x) { const unsigned int c = s_h[b]; if (c) atomicAdd(&hist[b], c); }
}

// one CTA: find the digit bin holding the wanted rank, extend the prefix, clear the histogram
__global__ void __launch_bounds__(SEL_BINS / 2)
select_pick_kernel(int pass, unsigned long long* __restrict__ sel, unsigned int* __restrict__ hist) {
    __shared__ unsigned int s_c[SEL_BINS];
    for (int b = threadIdx.x; b < SEL_BINS; b += blockDim.x) { s_c[b] = hist[b]; hist[b] = 0u; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long rank = sel[1];                         // 0-based rank wanted among the prefix-matching keys
        unsigned int d = 0;
        unsigned long long below = 0;
        for (; d < SEL_BINS - 1; ++d) { if (below + s_c[d] > rank) break; below += s_c[d]; }
        sel[0] = (sel[0] << SEL_BITS) | d;                        // after the last pass: the full key
        sel[1] = rank - below;
        if (pass == SEL_PASSES - 1) { sel[2] = sel[0]; sel[3] = rank - below + 1ull; }     // threshold key, ties to take
    }
}

// sums of x m, y m, z m, m over keys < T plus the first `ties` keys == T to claim a slot
__global__ void __launch_bounds__(RED_THREADS)
inner_com_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                 const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ vz,
                 const double* __restrict__ m, const unsigned long long* __restrict__ sel, unsigned int* __restrict__ tie_counter,
                 double* __restrict__ partial, unsigned int* __restrict__ counter, double* __restrict__ out) {
    __shared__ double s_red[4 * (RED_THREADS / 32)];
    const unsigned long long T = sel[2];
    const unsigned int ties = (unsigned int)sel[3];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double px = __ldg(x + i), py = __ldg(y + i), pz = __ldg(z + i);
        const unsigned long long k = radius_key(px, py, pz);
        bool take = k < T;
        if (k == T) take = atomicAdd(tie_counter, 1u) < ties;      // integer slot claim among exact ties
        if (take) {
            const double pm = __ldg(m + i);
            v[0] += __ldg(vx + i) * pm; v[1] += __ldg(vy + i) * pm; v[2] += __ldg(vz + i) * pm; v[3] += pm;
        }
    }
    cta_sum<4>(v, s_red);
    grid_finish<4>(v, partial, counter, out);
}

}  // namespace

extern "C" int bfe_bar_fourier(int64_t n, const double* x, const double* y, double minr, double maxr, double* out2,
                               void* stream_) {
    if (n < 0 || !out2 || (n > 0 && (!x || !y))) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    IngestWs* w = nullptr;
    int rc = ingest_ws(&w);
    if (rc != BFE_OK) return rc;
    bar_fourier_kernel<<<red_grid(n, w->max_ctas), RED_THREADS, 0, stream>>>(n, x, y, minr, maxr, w->partial, w->counter, out2);
    BFE_LAUNCH_CHECK("bar_fourier_kernel");
    return BFE_OK;
}

extern "C" int bfe_affine_xy(int64_t n, double angle, double cx, double cy, double cz, const double* x, const double* y,
                             const double* z, double* xo, double* yo, double* zo, void* stream_) {
    if (n < 0 || (n > 0 && (!x || !y || !xo || !yo)) || ((z != nullptr) != (zo != nullptr))) return BFE_ERR_ARG;
    if (n == 0) return BFE_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    int64_t need = (n + 255) / 256;
    affine_kernel<<<(int)(need < 148 * 16 ? need : 148 * 16), 256, 0, stream>>>(n, cos(angle), sin(angle), cx, cy, cz, x, y, z,
                                                                              xo, yo, zo);
    BFE_LAUNCH_CHECK("affine_kernel");
    return BFE_OK;
}

extern "C" int bfe_inner_com(int64_t n, const double* x, const double* y, const double* z,
                             const double* vx, const double* vy, const double* vz, const double* m,
                             int64_t ncenter, double* out4, void* stream_) {
    if (n <= 0 || !x || !y || !z || !vx || !vy || !vz || !m || !out4 || ncenter < 1) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    IngestWs* w = nullptr;
    int rc = ingest_ws(&w);
    if (rc != BFE_OK) return rc;
    if (ncenter > n) ncenter = n;                                  // rrank.argsort()[0:ncenter] clips
    const unsigned long long init[4] = {0ull, (unsigned long long)(ncenter - 1), 0ull, 0ull};
    BFE_CUDA(cudaMemcpyAsync(w->sel, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    BFE_CUDA(cudaMemsetAsync(w->counter + 1, 0, sizeof(unsigned int), stream));
    const int grid = red_grid(n, w->max_ctas);
    for (int pass = 0; pass < SEL_PASSES; ++pass) {
        select_hist_kernel<<<grid, RED_THREADS, 0, stream>>>(n, x, y, z, pass, w->sel, w->hist);
        select_pick_kernel<<<1, SEL_BINS / 2, 0, stream>>>(pass, w->sel, w->hist);
    }
    BFE_LAUNCH_CHECK("select kernels");
    inner_com_kernel<<<grid, RED_THREADS, 0, stream>>>(n, x, y, z, vx, vy, vz, m, w->sel, w->counter + 1, w->partial,
                                                       w->counter, out4);
    BFE_LAUNCH_CHECK("inner_com_kernel");
    return BFE_OK;
}
