/*
 * bfe.h -- C ABI of libbfe.so: B200 (sm_100a) kernels for exptool's
 * basis-function-expansion hot path.
 *
 * The reference (michael-petersen/exptool) is pure Python and has no FFI for this
 * path; the only native code, exptool/basis/accumulate_c/
 * accumulate.h:3-13, is a dormant CPython module of five scalar coordinate maps.
 * This header therefore DEFINES the boundary a maintainer binds with ctypes
 * (see INTEGRATION.md); each entry point names the reference function it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all arrays are FP64, particle data is SoA, contiguous;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call
 *     synchronises the device unless stated;
 *   - return value: 0 = ok, negative = error (bfe_error_string());
 *   - the caller owns every buffer; a handle owns only its re-laid-out tables and
 *     workspaces.  A handle must not be used from two streams at once.
 */
#ifndef BFE_H
#define BFE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFE_OK                 0
#define BFE_ERR_ARG          (-1)
#define BFE_ERR_CUDA         (-2)
#define BFE_ERR_UNSUPPORTED  (-3)
#define BFE_ERR_STATE        (-4)

/* limits of the compiled kernels */
#define BFE_MAX_MMAX   16
#define BFE_MAX_LMAX   16

typedef struct bfe_eof_params {
    int32_t mmax, norder, numx, numy, cmap, dens;   /* eof.eof_params, eof.py:98-200           */
    double  xmin, dx, ymin, dy;                     /* eof.set_table_params, eof.py:316-347     */
    double  ascale, hscale;
} bfe_eof_params;

typedef struct bfe_sl_params {
    int32_t lmax, nmax, numr, cmap;                 /* halo_methods.read_cached_table:128-140   */
    double  scale;
} bfe_sl_params;

typedef struct bfe_eof bfe_eof;     /* EOF tables on the device                                */
typedef struct bfe_sl  bfe_sl;      /* SL tables on the device                                 */

const char* bfe_error_string(int code);
const char* bfe_last_cuda_error(void);
int  bfe_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t bfe_launch_count(void);

/* Runtime options (process-wide).  "eof_accumulate_mode", "eof_force_mode", "sl_accumulate_mode":
 * 0 = auto (bin-sorted kernels from "sort_min_particles" particles up, direct kernels below),
 * 1 = direct, 2 = sorted. */
int bfe_set_option(const char* name, int value);
/* current value of an option; INT_MIN if the name is unknown */
int bfe_get_option(const char* name);

/* Table precision of the per-point field kernels (bfe_field_force_*, bfe_leapfrog*, bfe_sl_force_*) PER HANDLE:
 * 1 = the contracted tables are held as float (north_star: "<= 1e-5 where FP32 table interpolation is used"), 0 = FP64,
 * -1 (initial) = follow the process option "table_fp32".  The combined-field calls read the EOF handle's setting. */
int bfe_eof_set_table_fp32(bfe_eof* h, int value);
int bfe_sl_set_table_fp32(bfe_sl* h, int value);

/* Measured FP64 peak of the current device in TFLOP/s: kind 0 = vector DFMA (eight independent chains per thread),
 * kind 1 = FP64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA).  Runs ~10 ms of register-only arithmetic on `stream` and
 * synchronises it.  The FP64 side of bench.py's roofline uses these instead of a nominal figure. */
int bfe_fp64_peak(int kind, double* tflops, void* stream);

/* With option "time_kernels" = 1 the EOF step kernels are bracketed by CUDA events on their stream;
 * bfe_kernel_time_ms(name) synchronises on and returns the duration (ms) of the latest launch of that kernel
 * (e.g. "eof_segsum_kernel"), or a negative value if none was recorded. */
double bfe_kernel_time_ms(const char* name);

/* ---------------------------------------------------------------- EOF (disc) ---------------- */

/* Tables in the reference layout [m][n][ix][iy] (eof.parse_eof, eof.py:224-313), each
 * (mmax+1)*norder*(numx+1)*(numy+1) doubles.  Copies / re-lays them out on the device. */
int bfe_eof_create(const bfe_eof_params* p,
                   const double* potC, const double* rforceC, const double* zforceC,
                   const double* potS, const double* rforceS, const double* zforceS,
                   void* stream, bfe_eof** out);
void bfe_eof_destroy(bfe_eof* h);
/* A second handle on the same device tables with its own contraction, sorted-set workspace and counters: two
 * independent particle sets can then be processed concurrently on two streams.  `src` must outlive the clone. */
int bfe_eof_clone(const bfe_eof* src, void* stream, bfe_eof** out);

/* eof.accumulate (eof.py:492-551) for n particles:
 *   cos_out[m*norder+n] = sum_p -4pi cos(m phi_p) m_p interp_p(potC[m,n]),  sin_out likewise.
 * cos_out/sin_out: (mmax+1)*norder doubles each, overwritten. */
int bfe_eof_accumulate(bfe_eof* h, int64_t n,
                       const double* x, const double* y, const double* z, const double* mass,
                       double* cos_out, double* sin_out, void* stream);

/* Cell-sorted particle set ("prepared" set), held in the handle's workspace until the next
 * bfe_eof_prepare / bfe_eof_accumulate / bfe_eof_force* call on the same handle.  One counting sort by
 * table cell serves both passes over the same particles (accumulate, then force evaluation):
 *   bfe_eof_prepare(h, n, x, y, z, mass)        mass may be NULL if only forces are needed
 *   bfe_eof_accumulate_prepared(h, cos, sin)    == bfe_eof_accumulate on that set
 *   bfe_eof_force_prepared(h, p0, ..., R)       == bfe_eof_force_contracted on that set
 * Outputs are in the caller's particle order.  Supported for mmax <= 6 and (2 mmax+1) norder <= 256. */
int bfe_eof_prepare(bfe_eof* h, int64_t n,
                    const double* x, const double* y, const double* z, const double* mass, void* stream);
int bfe_eof_accumulate_prepared(bfe_eof* h, double* cos_out, double* sin_out, void* stream);
int bfe_eof_force_prepared(bfe_eof* h, double* p0, double* p, double* fr, double* fp, double* fz, double* R,
                           void* stream);

/* Host-array entry points: the same two passes for particle arrays in HOST memory (what the reference's Python
 * functions receive and return: eof.accumulate eof.py:492, eof.accumulated_eval_particles eof.py:989).  The set is
 * cut into chunks and copy-in, kernels and copy-out run as a three-stage pipeline on internal streams (pinned
 * host memory gives asynchronous DMA; pageable memory is accepted and copies synchronously).  One upload per snapshot
 * (option "host_reuse", default 1): bfe_eof_accumulate_host ALWAYS uploads and keeps its device copy; a following
 * bfe_eof_force_host on the same host arrays (same pointers, length and content tag: a hash of the first / last 64 and of
 * 1024 evenly spaced values per array) evaluates from that copy instead of uploading x, y, z again.
 *   bfe_eof_accumulate_host: hx..hm host arrays of n doubles; cos_out / sin_out are DEVICE buffers (so a
 *       multi-GPU caller can allreduce before copying 2 kB out), complete in stream order.
 *   bfe_eof_force_host: hx, hy, hz host inputs; hp0..hR host outputs of n doubles each, complete once `stream`
 *       has been synchronised.  Uses the held contraction (bfe_eof_contract). */
int bfe_eof_accumulate_host(bfe_eof* h, int64_t n,
                            const double* hx, const double* hy, const double* hz, const double* hm,
                            double* cos_out, double* sin_out, void* stream);
int bfe_eof_force_host(bfe_eof* h, int64_t n,
                       const double* hx, const double* hy, const double* hz,
                       double* hp0, double* hp, double* hfr, double* hfp, double* hfz, double* hR, void* stream);

/* Contract the tables with a coefficient set:  G_f[m,trig,node] = sum_{n<nuse} coef[m,n] T_f[m,n,node]
 * for m1 <= m <= min(m2, muse), skipping odd m if no_odd.  Must precede the *_contracted calls,
 * bfe_field_force_cart and bfe_leapfrog.  cosc/sinc: (mmax+1)*norder doubles. */
int bfe_eof_contract(bfe_eof* h, const double* cosc, const double* sinc,
                     int m1, int m2, int nuse, int no_odd, void* stream);

/* eof.accumulated_eval_particles (eof.py:989-1144): p0, p (m>=1 only), fr, fp, fz, R per particle,
 * R = sqrt(x^2+y^2+1e-10).  Uses the contraction held by the handle. */
int bfe_eof_force_contracted(bfe_eof* h, int64_t n,
                             const double* x, const double* y, const double* z,
                             double* p0, double* p, double* fr, double* fp, double* fz, double* R,
                             void* stream);

/* contract + evaluate in one call (the drop-in for eof.compute_forces, eof.py:1263-1317) */
int bfe_eof_force(bfe_eof* h, int64_t n,
                  const double* x, const double* y, const double* z,
                  const double* cosc, const double* sinc, int m1, int m2, int nuse, int no_odd,
                  double* p0, double* p, double* fr, double* fp, double* fz, double* R,
                  void* stream);

/* eof.force_eval (eof.py:756-870) at n cylindrical points (r, z, phi given directly):
 * fr (incl. m=0), fp, fz (incl. m=0), p (incl. m=0), p0.  Uses the held contraction. */
int bfe_eof_force_eval_points(bfe_eof* h, int64_t n,
                              const double* r, const double* z, const double* phi,
                              double* fr, double* fp, double* fz, double* p, double* p0,
                              void* stream);

/* ---------------------------------------------------------------- SL (halo) ----------------- */

/* evtable (lmax+1)*nmax, eftable (lmax+1)*nmax*numr (halo_methods.read_cached_table:96-169);
 * xi, p0, d0: numr each (halo_methods.init_table:178-220). */
int bfe_sl_create(const bfe_sl_params* p, const double* evtable, const double* eftable,
                  const double* xi, const double* p0, const double* d0,
                  void* stream, bfe_sl** out);
void bfe_sl_destroy(bfe_sl* h);

/* spheresl.compute_coefficients_solitary (spheresl.py:567-656):
 * expcoef[(lmax+1)^2 * nmax], rows l^2 (m=0), l^2+2m-1 (cos), l^2+2m (sin); overwritten. */
int bfe_sl_accumulate(bfe_sl* h, int64_t n,
                      const double* x, const double* y, const double* z, const double* mass,
                      int no_odd, double* expcoef, void* stream);

/* A[k,i] = sum_{n<nuse} expcoef[k,n] eftable[l(k),n,i]/sqrt(ev[l,n]) for l1 <= l <= min(l2,lmax)
 * (l=0 always kept), skipping odd l if no_odd.  expcoef: (lmax+1)^2*nmax doubles. */
int bfe_sl_contract(bfe_sl* h, const double* expcoef, int l1, int l2, int nuse,
                    int no_odd, void* stream);

/* spheresl.all_eval_particles (spheresl.py:1240-1362), potential outputs:
 * pot0, pot1, potr, pott, potp, rr per particle; r = sqrt(x^2+y^2+z^2), trig = cos/sin(m phi). */
int bfe_sl_force_contracted(bfe_sl* h, int64_t n,
                            const double* x, const double* y, const double* z,
                            double* pot0, double* pot1, double* potr, double* pott, double* potp,
                            double* rr, void* stream);

/* Host-array entry points of the SL passes, as bfe_eof_*_host: chunked copy-in | kernels | copy-out pipeline, the
 * evaluation of the arrays that were just accumulated runs from the accumulation's upload (option "host_reuse").
 *   bfe_sl_accumulate_host: spheresl.compute_coefficients (spheresl.py:439-475); expcoef is a DEVICE buffer.
 *   bfe_sl_force_host     : spheresl.all_eval_particles / eval_particles (spheresl.py:1240, 502): six HOST outputs
 *                           pot0, pot1, potr, pott, potp, rr, complete once `stream` has been synchronised. */
int bfe_sl_accumulate_host(bfe_sl* h, int64_t n,
                           const double* hx, const double* hy, const double* hz, const double* hm,
                           int no_odd, double* expcoef, void* stream);
int bfe_sl_force_host(bfe_sl* h, int64_t n,
                      const double* hx, const double* hy, const double* hz,
                      double* hpot0, double* hpot1, double* hpotr, double* hpott, double* hpotp, double* hrr,
                      void* stream);

int bfe_sl_force(bfe_sl* h, int64_t n,
                 const double* x, const double* y, const double* z,
                 const double* expcoef, int l1, int l2, int no_odd,
                 double* pot0, double* pot1, double* potr, double* pott, double* potp, double* rr,
                 void* stream);

/* spheresl.force_eval / all_eval (spheresl.py:1107-1234 / 987-1102) at n points (r, costh, phi):
 * potr, pott, potp, pot1, pot0.  trig_index_l=1 reproduces force_eval's cos/sin(l phi)
 * (spheresl.py:1173,1222-1225); 0 gives all_eval's cos/sin(m phi). */
int bfe_sl_force_eval_points(bfe_sl* h, int64_t n,
                             const double* r, const double* costh, const double* phi,
                             int trig_index_l,
                             double* potr, double* pott, double* potp, double* pot1, double* pot0,
                             void* stream);

/* Density outputs den0, den1 (a20; SURVEY.md App. C #8: the two reference functions differ and each is
 * reproduced as written).  bfe_sl_contract_density forms the density rows
 * sum_n expcoef[k,n] eftable[l,n,i] sqrt(ev[l,n]) (get_halo_dens_pot_force, spheresl.py:148) with the same
 * truncation arguments as bfe_sl_contract; the handle must have been created with d0.
 * bfe_sl_density_contracted: spheresl.all_eval_particles (spheresl.py:1271,1323,1351) per particle:
 *   den1 without the monopole, m=0 terms weighted by legs[1][0], densfac = 0.25*pi.
 * bfe_sl_density_eval_points: spheresl.all_eval (spheresl.py:1046,1071,1092) at (r, costh, phi):
 *   den1 includes the monopole, densfac = 0.25/pi. */
int bfe_sl_contract_density(bfe_sl* h, const double* expcoef, int l1, int l2, int nuse,
                            int no_odd, void* stream);
int bfe_sl_density_contracted(bfe_sl* h, int64_t n,
                              const double* x, const double* y, const double* z,
                              double* den0, double* den1, void* stream);
int bfe_sl_density_eval_points(bfe_sl* h, int64_t n,
                               const double* r, const double* costh, const double* phi,
                               double* den0, double* den1, void* stream);

/* ---------------------------------------------------------------- multi-GPU coefficient sum -- */
/* The one exchange step of the path: the sum of the partial coefficient blocks over the GPUs of a node
 * (np.sum(np.array(a_coeffs), axis=0) over Pool workers, eof.py:1440, spheresl.py:471), as ONE kernel over
 * NVLink peer memory: push to every rank's exchange buffer, flag, wait, fixed-order sum (bit-identical on all
 * ranks).  One process per GPU: every rank creates its buffer, the 64-byte IPC handles are exchanged by the
 * host side (torch.distributed), every rank opens the others' and builds a bfe_peer over the `world` pointers.
 * Calls on one bfe_peer must be issued in the same order on every rank (like any collective); n <= ncoef_max. */
typedef struct bfe_peer bfe_peer;
int bfe_peer_buffer_create(int64_t ncoef_max, void** local_ptr, unsigned char* handle64);
int bfe_peer_buffer_open(const unsigned char* handle64, void** peer_ptr);
int bfe_peer_buffer_close(void* peer_ptr);
int bfe_peer_buffer_destroy(void* local_ptr);
int bfe_peer_create(int rank, int world, int64_t ncoef_max, void* const* bufs, bfe_peer** out);
void bfe_peer_destroy(bfe_peer* p);
/* data[0..n) (device, this rank's partial sums) := sum over ranks, in place, on `stream`.  If a rank does not arrive
 * within 20 s the call fills data with NaN on every rank (never a partial sum) and the peer set stays failed. */
int bfe_peer_allreduce(bfe_peer* p, double* data, int64_t n, void* stream);
/* first sequence number that failed on this peer set (0: none); synchronises `stream`. */
int bfe_peer_error(bfe_peer* p, void* stream, unsigned long long* first_failed_seq);
/* fault injection for tests: mark this rank's buffer as failed at collective `seq` (synchronises `stream`). */
int bfe_peer_poison(bfe_peer* p, unsigned long long seq, void* stream);

/* ---------------------------------------------------------------- building blocks ----------- */
/* The per-point pieces the kernels above evaluate inline, callable with the meaning of the reference's own
 * helper functions.  Outputs are point-minor (trailing axis = the n points), like the reference's arrays. */

/* eof.return_bins (eof.py:354-427): X, Y fractional bins, ix, iy truncated and clamped bins (int64).
 * Lower edge clamps X; the upper edge clamps ix only (X extrapolates, eof.py:414-415).  No handle needed. */
int bfe_eof_return_bins(const bfe_eof_params* p, int64_t n, const double* r, const double* z,
                        double* X, double* Y, long long* ix, long long* iy, void* stream);

/* eof.get_pot (eof.py:430-457) on the handle's potC / potS: Vc, Vs of (mmax+1)*norder*n doubles,
 * [m][n][point]; the m=0 plane of Vs is zero (parse_eof leaves potS[0] zero, eof.py:293). */
int bfe_eof_get_pot(bfe_eof* h, int64_t n, const double* r, const double* z, double fac,
                    double* Vc, double* Vs, void* stream);

/* spheresl.get_halo_dens_pot_force (spheresl.py:106-160) / get_halo_pot_matrix (301-335) at n radii:
 * dens, force, pot of (lmax+1)*nmax*n doubles, [l][n][point]; any of the three may be NULL. */
int bfe_sl_radial_matrices(bfe_sl* h, int64_t n, const double* r,
                           double* dens, double* force, double* pot, void* stream);

/* spheresl.legendre_R / dlegendre_R (spheresl.py:664-770) at n arguments: P (and dP unless NULL) of
 * (lmax+1)*(lmax+1)*n doubles, [l][m][point], entries m > l zero, non-finite entries zero. */
int bfe_legendre_tables(int lmax, int64_t n, const double* x, double* P, double* dP, void* stream);

/* ---------------------------------------------------------------- combined field ------------ */

/* Fields.return_forces_cart (potential.py:445-497) at n points in a frame rotated by rotpos:
 * out8 = 8 SoA rows of n: fxdisk, fxhalo, fydisk, fyhalo, fzdisk, fzhalo, diskp, halop+halop0.
 * Both handles must hold a contraction (halofac already folded into expcoef). */
int bfe_field_force_cart(bfe_eof* he, bfe_sl* hs, int64_t n,
                         const double* x, const double* y, const double* z, double rotpos,
                         double* out8, void* stream);

/* Fields.return_forces_cyl (potential.py:389-440), same call shape; out8 rows are
 * diskfr, frhalo, diskfp, -halofp, diskfz, fzhalo, -diskp, halop+halop0 (r2, r3 use +1e-10). */
int bfe_field_force_cyl(bfe_eof* he, bfe_sl* hs, int64_t n,
                        const double* x, const double* y, const double* z, double rotpos,
                        double* out8, void* stream);

/* integrate.leapfrog_integrate (integrate.py:53-190) for norbit independent orbits.
 * state6: 6 SoA rows of norbit (x,y,z,vx,vy,vz): initial state in, state at the last step out.
 * nint steps INCLUDING step 0 (the reference's arrays have nint entries).
 * traj (optional, may be NULL): every traj_stride-th step k = 0, s, 2s, ... is written as
 *   traj[(k/s) * 10 * norbit + q * norbit + orbit], q = x,y,z,vx,vy,vz,pot,fx,fy,fz.
 * apse != 0: count planar apocentres and stop an orbit after ap_max (integrate.py:126,146-153);
 * nsteps_out (optional, int32 per orbit) receives the number of steps taken. */
int bfe_leapfrog(bfe_eof* he, bfe_sl* hs, int64_t norbit, int64_t nint, double dt, double rotfreq,
                 double* state6, double* traj, int64_t traj_stride,
                 int apse, int ap_max, int32_t* nsteps_out, void* stream);

/* Same with one step size per orbit (dt_orbit[norbit], device): the orbit grids of integrate.integrate_grid*
 * (integrate.py:760-922) give every orbit dt = max(compute_timestep(...), dt). */
int bfe_leapfrog_dt(bfe_eof* he, bfe_sl* hs, int64_t norbit, int64_t nint, const double* dt_orbit, double rotfreq,
                    double* state6, double* traj, int64_t traj_stride,
                    int apse, int ap_max, int32_t* nsteps_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Pre-accumulation transforms of a device-resident snapshot (Fields.total_coefficients, potential.py:120-226).
 * bfe_bar_fourier : out2[0] = sum cos 2phi, out2[1] = sum sin 2phi over minr < (x^2+y^2)^0.5 < maxr  (device doubles);
 *                   bar angle = -atan2(out2[1], out2[0]) / 2                     (analysis/pattern.py:104, 155-169)
 * bfe_affine_xy   : xo = x cos a - y sin a - cx, yo = x sin a + y cos a - cy, zo = z - cz (z, zo may both be NULL);
 *                   in place allowed (xo == x ...)                               (pattern.py:118-139, potential.py:213-219)
 * bfe_inner_com   : out4 = sums of vx m, vy m, vz m, m over the indices of the `ncenter` smallest (x^2+y^2+z^2)^0.5
 *                   (potential.py:158-176: rrank.argsort()[0:ncenter]; pass vx = x ... for a set's own centre; the
 *                   reference applies the DISC ranking to the halo arrays, potential.py:190-200); device doubles  */
int bfe_bar_fourier(int64_t n, const double* x, const double* y, double minr, double maxr, double* out2, void* stream);
int bfe_affine_xy(int64_t n, double angle, double cx, double cy, double cz,
                  const double* x, const double* y, const double* z, double* xo, double* yo, double* zo, void* stream);
int bfe_inner_com(int64_t n, const double* x, const double* y, const double* z,
                  const double* vx, const double* vy, const double* vz, const double* m,
                  int64_t ncenter, double* out4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BFE_H */
