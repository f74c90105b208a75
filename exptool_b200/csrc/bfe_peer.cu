// bfe_peer.cu -- the coefficient sum over the GPUs of one node as ONE kernel over NVLink peer memory.
//
// The only exchange on the path is the sum of the partial coefficient blocks (<= 9 kB; eof.py:1440,
// spheresl.py:471).  Through NCCL it costs a launch plus ~15-25 us of latency per step and its kernel queues behind
// the step's own grids (bench.py at N = 2..8: 0.134 -> 0.151..0.161 ms per step).  Here every rank owns an exchange
// buffer that all other ranks map (CUDA IPC); the kernel below
//   1. PUSHES this rank's block into slot [parity][rank] of every rank's buffer (plain stores over NVLink),
//   2. fences (system scope) and raises flag [parity][rank] = sequence number in every buffer,
//   3. waits until all `world` flags of its OWN buffer carry the sequence number,
//   4. sums the `world` blocks of its own buffer in RANK ORDER: every rank forms the same sum, bit for bit.
// One CTA, no host round trip, no separate communication stream.  Slots are double-buffered by the parity of the
// sequence number: a rank can only push sequence s+2 after it has finished s+1, which needs every peer's push of
// s+1, which the peer issues after it has consumed s (calls on one bfe_peer are stream-ordered on every rank).
// A wait that sees no progress for 20 s raises the error word of EVERY rank's buffer (sticky), returns NaN instead of a
// sum and gives up (no hung GPU, no silently wrong coefficients).
#include "bfe_internal.h"
#include "bfe_device.cuh"
#include <new>
#include <string.h>

#define BFE_PEER_MAXW 16
#define BFE_PEER_HDR (2 * BFE_PEER_MAXW + 2)          // 64-bit words: flags [2][MAXW], error word, spare

struct PeerParams {
    int rank, world;
    long long ncoef_max;
    unsigned long long* buf[BFE_PEER_MAXW];            // every rank's exchange buffer in THIS process's address space
};

struct bfe_peer {
    PeerParams pp;
    unsigned long long seq;
};

__device__ __forceinline__ unsigned long long* peer_flag(unsigned long long* base, int slot, int r) {
    return base + slot * BFE_PEER_MAXW + r;
}
__device__ __forceinline__ double* peer_data(unsigned long long* base, long long ncoef_max, int slot, int r) {
    return reinterpret_cast<double*>(base + BFE_PEER_HDR) + ((long long)slot * BFE_PEER_MAXW + r) * ncoef_max;
}

// The error word of a buffer is STICKY and shared: a rank whose wait times out (or that finds the word already set)
// raises it in EVERY rank's buffer, so the peers stop waiting at once, and from then on every call on this bfe_peer
// returns NaN in `data` -- a sum that could not be formed is never passed on as numbers (round 1 summed whatever the
// slots held).  The host side sees it through bfe_peer_error() or simply through the NaNs.
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(PeerParams pp, double* __restrict__ data, int n, unsigned long long seq) {
    __shared__ int s_fail;
    const int tid = threadIdx.x, slot = (int)(seq & 1ull);
    unsigned long long* my_err = pp.buf[pp.rank] + 2 * BFE_PEER_MAXW;
    if (tid == 0) s_fail = 0;
    bfe_pdl_wait();                                   // the producer of `data` (previous kernel of the stream) is done
    __syncthreads();
    if (tid == 0) {
        unsigned long long e;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(e) : "l"(my_err) : "memory");
        if (e != 0ull) s_fail = 1;                    // an earlier collective on this peer set failed
    }
    __syncthreads();
    if (!s_fail) {
        // 1. push this rank's block to every rank (own buffer included: step 4 then reads one layout)
        for (int p = 0; p < pp.world; ++p) {
            double* dst = peer_data(pp.buf[p], pp.ncoef_max, slot, pp.rank);
            for (int j = tid; j < n; j += blockDim.x) dst[j] = data[j];
        }
        __threadfence_system();
        __syncthreads();
        // 2. one thread per peer raises our flag there (release, system scope)
        if (tid < pp.world) {
            unsigned long long* f = peer_flag(pp.buf[tid], slot, pp.rank);
            asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(f), "l"(seq) : "memory");
        }
        // 3. one thread per peer waits for that peer's flag in OUR buffer (or for the error word)
        if (tid < pp.world) {
            const unsigned long long* f = peer_flag(pp.buf[pp.rank], slot, tid);
            unsigned long long v, e, t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            for (;;) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
                if (v >= seq) break;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(e) : "l"(my_err) : "memory");
                if (e != 0ull) { s_fail = 1; break; }                            // a peer gave up: so do we
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) { s_fail = 1; break; }             // 20 s without the peer: give up
                __nanosleep(200);
            }
        }
    }
    __syncthreads();
    // Dependents are released only now: a pre-launched dependent grid parks its CTAs in griddepcontrol.wait, and
    // parked CTAs of one stream could keep the kernels of another stream -- which a PEER is waiting for -- off the SMs.
    bfe_pdl_trigger();
    if (s_fail) {
        // error word (first sequence that failed) in every rank's buffer, never overwritten; NaN instead of a partial sum
        if (tid < pp.world) atomicCAS_system(pp.buf[tid] + 2 * BFE_PEER_MAXW, 0ull, seq);
        const double qnan = __longlong_as_double(0x7ff8000000000000ll);
        for (int j = tid; j < n; j += blockDim.x) data[j] = qnan;
        return;
    }
    // 4. fixed-order sum of the blocks in our own buffer (written by the peers: read around L1)
    for (int j = tid; j < n; j += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < pp.world; ++r) s += __ldcv(peer_data(pp.buf[pp.rank], pp.ncoef_max, slot, r) + j);
        data[j] = s;
    }
}

static size_t peer_bytes(int64_t ncoef_max) {
    return (size_t)BFE_PEER_HDR * 8 + (size_t)2 * BFE_PEER_MAXW * (size_t)ncoef_max * 8;
}

extern "C" int bfe_peer_buffer_create(int64_t ncoef_max, void** local_ptr, unsigned char* handle64) {
    if (ncoef_max < 1 || !local_ptr || !handle64) return BFE_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    BFE_CUDA(cudaMalloc(&p, peer_bytes(ncoef_max)));
    BFE_CUDA(cudaMemset(p, 0, peer_bytes(ncoef_max)));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); bfe_set_cuda_error(e, "cudaIpcGetMemHandle"); return BFE_ERR_CUDA; }
    memcpy(handle64, &h, 64);
    *local_ptr = p;
    return BFE_OK;
}

extern "C" int bfe_peer_buffer_open(const unsigned char* handle64, void** peer_ptr) {
    if (!handle64 || !peer_ptr) return BFE_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    BFE_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BFE_OK;
}

extern "C" int bfe_peer_buffer_close(void* peer_ptr) {
    if (peer_ptr) BFE_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return BFE_OK;
}

extern "C" int bfe_peer_buffer_destroy(void* local_ptr) {
    if (local_ptr) BFE_CUDA(cudaFree(local_ptr));
    return BFE_OK;
}

extern "C" int bfe_peer_create(int rank, int world, int64_t ncoef_max, void* const* bufs, bfe_peer** out) {
    if (!out || !bufs || world < 1 || world > BFE_PEER_MAXW || rank < 0 || rank >= world || ncoef_max < 1) return BFE_ERR_ARG;
    bfe_peer* p = new (std::nothrow) bfe_peer();
    if (!p) return BFE_ERR_ARG;
    p->pp.rank = rank; p->pp.world = world; p->pp.ncoef_max = ncoef_max;
    for (int r = 0; r < BFE_PEER_MAXW; ++r) p->pp.buf[r] = (r < world) ? (unsigned long long*)bufs[r] : nullptr;
    for (int r = 0; r < world; ++r) if (!bufs[r]) { delete p; return BFE_ERR_ARG; }
    p->seq = 0;
    *out = p;
    return BFE_OK;
}

extern "C" void bfe_peer_destroy(bfe_peer* p) { delete p; }

extern "C" int bfe_peer_allreduce(bfe_peer* p, double* data, int64_t n, void* stream_) {
    if (!p || n < 0) return BFE_ERR_ARG;
    if (n == 0) return BFE_OK;
    if (!data || n > p->pp.ncoef_max) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    ++p->seq;
    const int kt = bfe_kt_begin("peer_allreduce_kernel", stream);
    BFE_CUDA(bfe_launch(peer_allreduce_kernel, dim3(1), dim3(256), 0, stream, nullptr, 0, p->pp, data, (int)n, p->seq));
    bfe_kt_end(kt, stream);
    BFE_LAUNCH_CHECK("peer_allreduce_kernel");
    return BFE_OK;
}

// fault injection (tests): raise this rank's error word as if collective `seq` had timed out
extern "C" int bfe_peer_poison(bfe_peer* p, unsigned long long seq, void* stream_) {
    if (!p || seq == 0) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    BFE_CUDA(cudaMemcpyAsync(p->pp.buf[p->pp.rank] + 2 * BFE_PEER_MAXW, &seq, 8, cudaMemcpyHostToDevice, stream));
    BFE_CUDA(cudaStreamSynchronize(stream));
    return BFE_OK;
}

// first sequence number whose wait timed out (0: none); synchronises the device-to-host copy on `stream`
extern "C" int bfe_peer_error(bfe_peer* p, void* stream_, unsigned long long* first_failed_seq) {
    if (!p || !first_failed_seq) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    BFE_CUDA(cudaMemcpyAsync(first_failed_seq, p->pp.buf[p->pp.rank] + 2 * BFE_PEER_MAXW, 8, cudaMemcpyDeviceToHost, stream));
    BFE_CUDA(cudaStreamSynchronize(stream));
    return BFE_OK;
}
