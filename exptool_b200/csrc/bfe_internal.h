// bfe_internal.h -- handle layouts and launch helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/bfe.h"

// Geometry passed by value to kernels (eof.set_table_params, eof.py:316-347).
struct EofGeom {
    int mmax, norder, numx, numy, cmap;
    int nnode;          // (numx+1)*(numy+1)
    int ny1;            // numy+1 : node = ix*ny1 + iy
    double xmin, dx, ymin, dy, ascale, hscale;
    double inv_dx, inv_dy;
    double inv_ascale, inv_hscale;   // host-rounded reciprocals: r/ascale and z/hscale become multiplies (<= 1 ulp)
};

// halo_methods.init_table:178-220 geometry
struct SlGeom {
    int lmax, nmax, numr, cmap;
    int nrow;           // (lmax+1)^2
    int ln;             // (lmax+1)*nmax
    double scale, xi0, dxi;
    double inv_scale, inv_dxi;       // host-rounded reciprocals (<= 1 ulp from the reference's divisions)
};

struct bfe_eof {
    bfe_eof_params par;
    EofGeom g;
    int device;
    int num_sms;
    // accumulate table: node-major [nnode][nch_pad], channels = cos (m*norder+n) then sin ((m-1)*norder+n, m>=1)
    int nch, nch_pad;
    double* t_acc;
    // reference-layout copies of the six force tables [6][m][n][node] (potC,rfC,zfC,potS,rfS,zfS)
    double* t_force;
    size_t tab_elems;    // (mmax+1)*norder*nnode
    // contracted grids G[node][m][6] (pc,ps,rc,rs,zc,zs)
    double* g_con;
    int gstride;         // 6*(mmax+1)
    int contracted;
    // the same contraction as one contiguous block per CELL: G4[cell][m][corner 00,10,01,11][3 x double2]
    // (built on demand from g_con for the warp-cooperative kernels)
    double* g4;
    int g4_valid;
    float* g4f;          // FP32 copy of G4 for the table_fp32 mode (allocated on first use)
    int g4f_valid;
    // accumulate workspace
    double* partial;     // [max_ctas][nch_pad]
    unsigned int* counter;
    int max_ctas;
    // cell-sorted accumulate workspace (grown on demand)
    int64_t sort_cap;
    void* sort_ws;
    int64_t prepared_n;      // particles currently held cell-sorted in sort_ws (-1: none)
    int prepared_has_mass;
    int owns_tables;         // 0 for a clone (bfe_eof_clone): t_acc / t_force belong to the parent handle
    void* host_pipe;         // staging buffers / streams of the host-array entry points (bfe_host.cu), lazily made
    void* orbit_ws;          // key-sort workspace of the table-coherent field / leapfrog paths (bfe_orbit_sort.cu), grown on demand
    int64_t orbit_cap;
    int64_t orbit_hdr;       // bytes of the workspace header (histogram, starts, block prefixes: depends on option key_subbits)
    void* keycell;           // per-cell key spans {key offset, first interval, span, -} for the SL table of keycell_tag (bfe_orbit_sort.cu)
    int keycell_nkeys, keycell_nkeys2;       // keys with one per interval / one per four intervals
    unsigned long long keycell_tag;
    void* orbit_rec;         // 96-byte orbit records of the key-ordered leapfrog path, grown on demand
    int64_t orbit_rec_cap;
    void* field_pipe;        // aux stream + events of the two-stream point pipeline (bfe_orbit_sort.cu), lazily made
    int table_fp32;          // per-handle table precision of the per-point field kernels: -1 inherit option "table_fp32", 0 FP64, 1 FP32
};

struct bfe_sl {
    bfe_sl_params par;
    SlGeom g;
    int device;
    int num_sms;
    double* e_node;      // node-major [numr][(lmax+1)*nmax], pre-divided by sqrt(ev)
    double* xi;          // [numr]
    double* p0;          // [numr]
    double* d0;          // [numr]
    int have_d0;         // bfe_sl_create was given the model density table
    double* ev;          // eigenvalues [(lmax+1)*nmax] (density rows are ef*sqrt(ev), spheresl.py:148)
    double* ad_con;      // contracted DENSITY rows, same layout as a_con (allocated on first use)
    int dens_contracted;
    double* fac;         // factorial_return [(lmax+1)*(lmax+1)]
    double fac_host[(BFE_MAX_LMAX + 1) * (BFE_MAX_LMAX + 1)];
    double* a_con;       // contracted rows, node-major [numr][kpad] of double2 (cos, sin), kpad = (l,m) pairs
    int kpad;
    int contracted;
    // the same rows as one contiguous block per radial index j: A3[j][(m,l)][rows j-1, j, j+1] (double2)
    double* a3;
    int a3_valid;
    float* a3f;          // FP32 copy of A3 for the table_fp32 mode (allocated on first use)
    int a3f_valid;
    double* a4;          // FP64 per-interval polynomial blocks of the per-lane evaluation (bfe_sl_eval_poly; allocated on first use)
    int a4_valid;
    double* partial;     // [max_ctas][nrow*nmax]
    unsigned int* counter;
    int max_ctas;
    // radial-bin-sorted accumulate workspace (grown on demand)
    int64_t sort_cap;
    void* sort_ws;
    int table_fp32;      // as bfe_eof::table_fp32 (SL-only evaluation calls)
    void* host_pipe;     // staging buffers / streams of the host-array entry points (bfe_host.cu), lazily made
};

extern "C" void bfe_count_launch(int n);
struct SlFacP;                                              // bfe_device.cuh: the factorial factors by value (kernel parameter)

// runtime options (bfe_set_option): 0 = auto, 1 = direct kernels, 2 = cell-sorted kernels
extern int g_bfe_eof_accumulate_mode;
extern int g_bfe_eof_force_mode;
extern int g_bfe_sort_min_particles;
extern int g_bfe_sl_accumulate_mode;
extern int g_bfe_sl_deposit_mode;     // option "sl_deposit_mode": 0/1 shared-memory slab kernel (default), 2 register formulation

bool bfe_sl_sorted_supported(const bfe_sl* h);
int bfe_sl_accumulate_sorted(bfe_sl* h, int64_t n, const double* x, const double* y, const double* z,
                             const double* mass, int no_odd, double* expcoef, cudaStream_t stream);

int bfe_eof_accumulate_sorted(bfe_eof* h, int64_t n, const double* x, const double* y, const double* z,
                              const double* mass, double* cos_out, double* sin_out, cudaStream_t stream);
int bfe_eof_force_sorted(bfe_eof* h, int64_t n, const double* x, const double* y, const double* z,
                         double* p0, double* p, double* fr, double* fp, double* fz, double* R, cudaStream_t stream);
void bfe_set_cuda_error(cudaError_t e, const char* where);
int bfe_eof_ensure_g4(bfe_eof* h, cudaStream_t stream);     // build G4 from g_con if stale
int bfe_sl_ensure_a3(bfe_sl* h, cudaStream_t stream);       // build A3 from a_con if stale
extern int g_bfe_force_mma;                                 // option "force_mma": sorted force eval on DMMA (1, default) / per lane (0)
extern int g_bfe_table_fp32;                                // option "table_fp32": process default for handles that do not set their own
inline bool bfe_use_fp32(const bfe_eof* h) { return (h->table_fp32 >= 0 ? h->table_fp32 : g_bfe_table_fp32) != 0; }
inline bool bfe_use_fp32(const bfe_sl* h) { return (h->table_fp32 >= 0 ? h->table_fp32 : g_bfe_table_fp32) != 0; }
int bfe_eof_ensure_g4f(bfe_eof* h, cudaStream_t stream);
int bfe_sl_ensure_a3f(bfe_sl* h, cudaStream_t stream);
int bfe_tile_colscan(int nbin, int ntile, int* H, int* total, int* bin_start, unsigned int* counter, cudaStream_t stream);   // bfe_sort.cu
int bfe_sl_ensure_a4(bfe_sl* h, cudaStream_t stream);        // build A4 from a_con if stale
extern int g_bfe_blk_eval;                                  // option "blk_eval": per-lane block evaluation with 256-bit loads
extern int g_bfe_staged_eval;                               // option "staged_eval": 1 (default) / 0

// Optional per-kernel CUDA-event timing (option "time_kernels"): bench.py's roofline object reads the live
// duration of every kernel of the step with bfe_kernel_time_ms().  Off by default (no events recorded).
int bfe_kt_begin(const char* name, cudaStream_t stream);     // returns a slot or -1 when disabled
void bfe_kt_end(int slot, cudaStream_t stream);

// Launch helper of the cell-sorted step (bfe_sort.cu, eof_contract_kernel).
//  * option "pdl" (default 1): programmatic dependent launch -- the next kernel of the stream is scheduled while
//    this one drains and runs its prologue up to bfe_pdl_wait() (griddepcontrol.wait), which returns once every
//    earlier kernel of the chain has completed and flushed.  Removes the ~4 us launch gap at each of the seven
//    kernel boundaries of a step (device timeline, profiles/trace_step.py).
//  * option "l2_persist" (default 0/1 see bfe_field.cu): an access-policy window marks a table (t_force, t_acc,
//    g_con) persisting in the L2 set-aside so the 80 MB/step particle stream does not evict it.
void bfe_host_pipe_destroy(void* pipe);
void bfe_field_pipe_destroy(void* pipe);
extern int g_bfe_host_chunk;                               // option "host_chunk"
extern int g_bfe_host_reuse;                               // option "host_reuse"
extern int g_bfe_host_threads;                             // option "host_threads"
extern int g_bfe_host_reused_last;                         // read-only option "host_reused_last"
extern int g_bfe_contract_deep;                            // option "contract_deep": 9 (1) or 6 (0) table loads in flight
extern int g_bfe_pdl;
extern int g_bfe_orbit_resort;                             // option "orbit_resort": steps between re-sorts of a large orbit batch (0: off)
extern int g_bfe_orbit_sort_min;                           // option "orbit_sort_min": smallest batch on the sorted path
extern int g_bfe_field_sort_chunk;                         // option "field_sort_chunk"
extern int g_bfe_field_sort_min;                           // option "field_sort_min"
extern int g_bfe_stage_eval;                               // option "stage_eval"
extern int g_bfe_sort_stable;                              // option "sort_stable"
extern int g_bfe_sl_flush_cost;                            // option "sl_flush_cost"
extern int g_bfe_key_subbits;                              // option "key_subbits"
extern int g_bfe_orbit_key_subbits;                        // option "orbit_key_subbits"
extern int g_bfe_key_mode;                                 // option "key_mode"
extern int g_bfe_keycell_nkeys_last;
extern int g_bfe_field_eval_static;
extern int g_bfe_field_support_slim;
extern int g_bfe_field_gather_stream;
int bfe_field_force_sorted(bfe_eof* he, bfe_sl* hs, int64_t n, const double* x, const double* y, const double* z,
                           double crot, double srot, double* out8, bool cyl, cudaStream_t stream);
int bfe_leapfrog_sorted(bfe_eof* he, bfe_sl* hs, int64_t norbit, int64_t nint, double dt, const double* dt_orbit,
                        double rotfreq, double* state6, int32_t* nsteps_out, cudaStream_t stream);
extern int g_bfe_grid_pct;                                 // option "grid_pct" (default 100)
extern int g_bfe_l2_persist;
extern size_t g_bfe_l2_window_max;                         // cudaDeviceProp::accessPolicyMaxWindowSize
template <typename... KArgs, typename... Args>
inline cudaError_t bfe_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              const void* win_ptr, size_t win_bytes, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    unsigned int na = 0;
    if (g_bfe_pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (g_bfe_l2_persist > 0 && g_bfe_l2_window_max > 0 && win_ptr && win_bytes) {
        at[na].id = cudaLaunchAttributeAccessPolicyWindow;
        at[na].val.accessPolicyWindow.base_ptr = const_cast<void*>(win_ptr);
        at[na].val.accessPolicyWindow.num_bytes = win_bytes < g_bfe_l2_window_max ? win_bytes : g_bfe_l2_window_max;
        at[na].val.accessPolicyWindow.hitRatio = 1.0f;
        at[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        at[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
int bfe_l2_persist_setup(size_t want_bytes);                // sets cudaLimitPersistingL2CacheSize once (device-wide)

#define BFE_CUDA(call)                                                  \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) { bfe_set_cuda_error(_e, #call); return BFE_ERR_CUDA; } \
    } while (0)

#define BFE_LAUNCH_CHECK(where)                                         \
    do {                                                                \
        cudaError_t _e = cudaGetLastError();                            \
        if (_e != cudaSuccess) { bfe_set_cuda_error(_e, where); return BFE_ERR_CUDA; } \
        bfe_count_launch(1);                                            \
    } while (0)
