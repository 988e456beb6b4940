/*
 * karman_b200.h -- C ABI of libkarman_b200.so (hand-written sm_100a CUDA kernels).
 *
 * The reference (houkensjtu/karman_taichi) has no FFI: its device boundary is
 * Python -> Taichi JIT (every @ti.kernel call, every field[None] access, every
 * reduce() return value).  This header is the boundary a Taichi-free host binds in its
 * place; each entry point cites the reference @ti.kernel / method it replaces
 * (file:line under the reference root).  The Python mirror of the reference classes
 * (karman_taichi_b200/{cgsolver,bicgsolver,simple,karman}.py) calls exactly these
 * symbols through ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - All array arguments are DEVICE pointers unless the name ends in _host.
 *  - A "field" is a (NX, pitch) row-major plane of reals (double for KT_F64, float
 *    for KT_F32); logical shape (NX, NY), NY <= pitch, pitch a multiple of 128 bytes (required: 16-byte vector accesses
 *    never straddle a row) -- kt_pitch_for() returns the recommended multiple of 512 bytes (a warp's 32 x 16 B request then
 *    never straddles a 512-byte boundary: 13 % on the element-wise kernels, profiles/r02_p2p.md).  Pad columns
 *    [NY, pitch) must be zero in every field and stay zero.
 *  - A "coef" array is 5 consecutive planes (ap, aw, ae, an, as), plane stride
 *    NX*pitch: the SoA form of the reference's coef[i,j,0..4] (cgsolver.py:59-60;
 *    aw multiplies x[i-1,j], ae x[i+1,j], an x[i,j+1], as x[i,j-1]).
 *  - Every function returns 0 on success, <0 on CUDA/argument error (message via
 *    kt_last_error()).  Numerical breakdown is NOT an error: it is reported in
 *    *status (KT_STATUS_*), so the host can mirror the reference's print/break/raise.
 *  - Calls are asynchronous on the handle's stream unless they return a value to the
 *    host (kt_dot, kt_*_solve, kt_relax_uv, kt_mdiv_max ...), which synchronise that
 *    stream.  A handle is not thread-safe; one handle per (device, geometry).
 *  - Ownership: the caller owns every field/coef buffer.  The handle owns only its
 *    workspace (block partial sums, device scalars, pinned mirrors, events).
 */
#ifndef KARMAN_B200_H
#define KARMAN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kt_handle_s* kt_handle_t;

enum { KT_F64 = 0, KT_F32 = 1 };

/* solver exit status (bicgsolver.py:96-98,108,118,130; karman_taichi.py:386-389,446-454) */
enum {
  KT_STATUS_CONVERGED = 0,
  KT_STATUS_MAX_ITER = 1,
  KT_STATUS_RHO_ZERO = 2,       /* '>>> BICG failed at first place...' */
  KT_STATUS_ALPHA_DIV_ZERO = 3, /* (r_tld, Ap) == 0 or (p, Ap) == 0: reference raises ZeroDivisionError */
  KT_STATUS_OMEGA_DIV_ZERO = 4, /* (t, t) == 0: reference raises ZeroDivisionError */
  KT_STATUS_NON_FINITE = 5,
  KT_STATUS_OMEGA_ZERO = 6,     /* karman_taichi.py:451-454 */
  KT_STATUS_COMM_TIMEOUT = 7    /* x-slab peer-memory solve: a neighbour's flag never arrived */
};

/* kt_bicgstab_solve flags */
enum {
  KT_FLAG_WARM_START = 1,       /* r0 = b - A x0 (karman_taichi.py:360-371) instead of x0 = 0 */
  KT_FLAG_OMEGA_ZERO_BREAK = 2  /* karman_taichi.py:451-454 */
};

/* ---- library / handle ------------------------------------------------------------- */
int kt_version(void);
const char* kt_last_error(void);
/* recommended row pitch (elements) for a logical row length NY */
int kt_pitch_for(int NY, int dtype);
/* geometry = the (NX, NY) of the fields this handle operates on; stream = cudaStream_t
 * (NULL = legacy default stream).  Replaces ti.init + the SNode allocation of
 * cgsolver.py:17-26 (workspace only -- fields stay with the caller). */
int kt_create(int device, int NX, int NY, int pitch, int dtype, void* stream, kt_handle_t* out);
int kt_destroy(kt_handle_t h);
int kt_set_stream(kt_handle_t h, void* stream);
int kt_sync(kt_handle_t h);
/* Restrict SpMV / vector updates / dots to rows [row_lo, row_hi) of the local plane (x-slab
 * ownership; rows outside are halo rows that are read but never written or summed).
 * Default [0, NX).  No reference counterpart (the reference is single-address-space). */
int kt_set_rows(kt_handle_t h, int row_lo, int row_hi);
/* Resident solvers.  Small grids (~35k f64 unknowns: the solve fits one thread-block cluster) are solved by a single
 * cluster-resident kernel, mid-size grids (up to ~3e5 unknowns) by one persistent cooperative kernel with one CTA per SM,
 * instead of 3-4 launches per iteration.  enable: 1 automatic (default), 0 streaming kernels only, 2 cooperative kernel
 * only, 3 cluster kernel only. */
int kt_set_cluster(kt_handle_t h, int enable);
/* Load every kernel of the library on the handle's device now (CUDA loads device functions lazily, at their first launch, and
 * that load may wait for the device to go idle).  Returns the number of kernels (> 0) or < 0.  No reference counterpart (Taichi
 * JIT-compiles at first call); required before kernels of different streams wait for each other on ONE device -- the in-process
 * multi-rank mode of the x-slab solvers (slab.py, peers='local') -- and harmless otherwise. */
int kt_preload_kernels(kt_handle_t h);
/* number of kernels launched through this handle so far (bench.py: gpu_launches) */
long long kt_launch_count(kt_handle_t h);

/* ---- host <-> device staging (reference: field.from_numpy / to_numpy) ---------------------- */
/* dense (NX, NY) host array -> pitched device field, and back */
int kt_upload_field(kt_handle_t h, void* dst, const void* src_host);
int kt_download_field(kt_handle_t h, void* dst_host, const void* src);
/* reference layout coef[i][j][k] (AoS, k fastest) on host <-> 5 pitched device planes, in row chunks through
 * a caller-owned device staging buffer of staging_bytes (>= one row = NY*5 reals; larger = fewer chunks). */
int kt_upload_coef(kt_handle_t h, void* dst_soa, const void* src_host_aos, void* staging, size_t staging_bytes);
int kt_download_coef(kt_handle_t h, void* dst_host_aos, const void* src_soa, void* staging, size_t staging_bytes);
/* device-side AoS (NX,NY,5) dense <-> SoA pitched conversion */
int kt_coef_aos_to_soa(kt_handle_t h, void* dst_soa, const void* src_aos);
int kt_coef_soa_to_aos(kt_handle_t h, void* dst_aos, const void* src_soa);

/* ---- Krylov building blocks (one per reference kernel; for unit parity) --------------------- */
/* y = A x.  cgsolver.py:56-66 (compute_Ap, compute_Ax), bicgsolver.py:65-76 (compute_As, compute_Ap) */
int kt_spmv(kt_handle_t h, const void* coef, const void* x, void* y);
/* y = A x and *dot_host = (q, y) in one pass (fused compute_Ap + reduce: cgsolver.py:94-95,
 * bicgsolver.py:106-107).  q may equal x. */
int kt_spmv_dot(kt_handle_t h, const void* coef, const void* x, void* y, const void* q, double* dot_host);
/* s = r - alpha*v ; t = A s ; dots_host[0] = (t,s), dots_host[1] = (t,t)
 * (fused update_s + compute_As + copy + 2 reduce: bicgsolver.py:110-117) */
int kt_bicg_s_t(kt_handle_t h, const void* coef, const void* r, const void* v, double alpha,
                void* s, void* t, double* dots_host);
/* sum over the owned rows of p*q.  cgsolver.py:49-54 (reduce).  Deterministic: fixed grid,
 * fixed-order two-stage (warp shuffle -> block -> ordered final) summation. */
int kt_dot(kt_handle_t h, const void* p, const void* q, double* out_host);
int kt_copy(kt_handle_t h, void* dst, const void* src);          /* bicgsolver.py:38-41 */
int kt_fill(kt_handle_t h, void* dst, double value);
/* y = y + a*x      cgsolver.py:68-76 (update_x; update_r with a = -alpha) */
int kt_axpy(kt_handle_t h, void* y, double a, const void* x);
/* out = x + a*y    cgsolver.py:78-81 (update_p; out may alias y) */
int kt_xpay(kt_handle_t h, void* out, const void* x, double a, const void* y);
/* out = x - a*y    bicgsolver.py:60-63 (update_s), :83-86 (update_r) */
int kt_xmay(kt_handle_t h, void* out, const void* x, double a, const void* y);
/* p = r + beta*(p - omega*v)          bicgsolver.py:43-46 (update_p) */
int kt_bicg_update_p(kt_handle_t h, void* p, const void* r, const void* v, double beta, double omega);
/* x = x + (alpha*p + omega*s)         bicgsolver.py:78-81 (update_x) */
int kt_bicg_update_x(kt_handle_t h, void* x, const void* p, const void* s, double alpha, double omega);
/* fused: x += alpha*p + omega*s ; r = s - omega*t ; dots_host = {(r,r), (r,r_tld)}
 * (update_x + update_r + 2 reduce: bicgsolver.py:120-123,95) */
int kt_bicg_xr(kt_handle_t h, void* x, void* r, const void* p, const void* s, const void* t,
               const void* r_tld, double alpha, double omega, double* dots_host);
/* fused: x += alpha*p ; r -= alpha*Ap ; *dot_host = (r,r)   (cgsolver.py:98-101) */
int kt_cg_xr(kt_handle_t h, void* x, void* r, const void* p, const void* Ap, double alpha, double* dot_host);

/* ---- device-resident solvers ---------------------------------------------------------- */
typedef struct {
  void* r;      /* residual                                   (cgsolver.py:18) */
  void* p;      /* search direction                           (cgsolver.py:19) */
  void* Ap;     /* A p  (BiCGSTAB's v)                        (cgsolver.py:20) */
  void* r_tld;  /* BiCGSTAB shadow residual; NULL => alias b when x0 = 0 (bicgsolver.py:9) */
  void* s;      /* BiCGSTAB s                                 (bicgsolver.py:11) */
  void* t;      /* BiCGSTAB t = A s (reference Ashat/t)       (bicgsolver.py:13-14) */
  double* history; /* DEVICE array or NULL: sqrt(rTr) per iteration (cgsolver.py:88,111) */
  long history_cap;
  void* p2;     /* optional second p / Ap planes: when both are given BiCGSTAB fuses the p-update into the v = A p */
  void* Ap2;    /* kernel (ping-pong buffers; 27 N instead of 28 N words per iteration).  NULL = unfused.          */
} kt_krylov_ws;

/* CGSolver.solve (cgsolver.py:83-114): x0 = 0, stop sqrt(rTr) < sqrt(rTr0)*eps, at most
 * max_iter loop bodies.  *iters = loop bodies executed.  *rel_res = sqrt(rTr/rTr0). */
int kt_cg_solve(kt_handle_t h, const void* coef, const void* b, void* x, const kt_krylov_ws* ws,
                double eps, long max_iter, long* iters, double* rel_res, int* status);
/* BICGSolver.solve (bicgsolver.py:88-135) when Minv == NULL and flags == 0: unpreconditioned,
 * x0 = 0, r_tld = b.  With Minv (= 1/M plane) and KT_FLAG_WARM_START|KT_FLAG_OMEGA_ZERO_BREAK:
 * the monolith's Jacobi right-preconditioned kernel (karman_taichi.py:342-456). */
int kt_bicgstab_solve(kt_handle_t h, const void* coef, const void* b, void* x, const void* Minv,
                      const kt_krylov_ws* ws, double eps, long max_iter, int flags,
                      long* iters, double* rel_res, int* status);
/* Fixed-iteration throughput variants used by bench.py: run exactly n_iter iterations with
 * the convergence test disabled (eps = 0); *ms = device time (CUDA events on the handle's stream). */
int kt_bicgstab_bench(kt_handle_t h, const void* coef, const void* b, void* x, const kt_krylov_ws* ws,
                      long n_iter, float* ms);
int kt_cg_bench(kt_handle_t h, const void* coef, const void* b, void* x, const kt_krylov_ws* ws,
                long n_iter, float* ms);
/* Reference-facing end-to-end call with HOST buffers in the reference layout
 * (coef[i][j][k] AoS, dense (NX,NY) b and x): uploads, solves, downloads x.
 * dev_* are caller-owned device buffers: coef_soa (5 planes), staging (staging_bytes, see kt_upload_coef), b, x. */
int kt_bicgstab_solve_host(kt_handle_t h, const void* coef_host_aos, const void* b_host, void* x_host,
                           void* dev_coef_soa, void* dev_staging, size_t staging_bytes, void* dev_b, void* dev_x,
                           const kt_krylov_ws* ws, double eps, long max_iter,
                           long* iters, double* rel_res, int* status);

/* ---- x-slab (multi-GPU) BiCGSTAB phases -------------------------------------------------------
 * No reference counterpart (the reference is single-process).  The plane is this rank's slab: owned rows
 * [row_lo, row_hi) (kt_set_rows) plus one halo row on each interior side that the host fills from the neighbour
 * rank before a phase that reads it (phase 2 reads p, phase 3 reads r and v with halos).  Each phase launches the
 * same kernel the single-GPU solver uses, but only publishes the LOCAL inner products into dots_dev[0..1]; the
 * host all-reduces dots_dev over the ranks and kt_dist_scalar runs the step's scalar algebra (KT_FIN_*) on every
 * rank from the identical reduced sums.  phase: 0 init (a0=b; o0=r, o1=x, o2=r_tld or NULL), 1 p-update (a0=r, a1=v;
 * o0=p), 2 v=Ap (a0=p, a1=r_tld; o0=v), 3 s,t (a0=r, a1=v; o0=s, o1=t), 4 x,r (a0=p, a1=s, a2=t; o0=x, o1=r, o2=r_tld).
 * CG (cgsolver.py:83-114; p needs its halo rows before phase 6): 5 init (a0=b; o0=r, o1=x, o2=p), 6 Ap (a0=p; o0=Ap) with
 * local (p,Ap), 7 x,r (a0=p, a1=Ap; o0=x, o1=r) with local (r,r), 8 p-update (a0=r; o0=p). */
enum { KT_FIN_NONE = 0, KT_FIN_BICG_INIT = 1, KT_FIN_BICG_ALPHA = 2, KT_FIN_BICG_OMEGA = 3, KT_FIN_BICG_END = 4,
       KT_FIN_CG_INIT = 5, KT_FIN_CG_ALPHA = 6, KT_FIN_CG_END = 7 };
int kt_dist_begin(kt_handle_t h, double eps, long max_iter, int flags);
int kt_dist_bicg_phase(kt_handle_t h, int phase, const void* coef, const void* a0, const void* a1, const void* a2,
                       void* o0, void* o1, void* o2, double* dots_dev);
int kt_dist_scalar(kt_handle_t h, int fin, const double* dots_dev, double* history_dev, long history_cap);
int kt_dist_poll(kt_handle_t h, int* done, long* iters, int* status, double* rel_res);

/* Same solve with the halo exchange and the all-reduce done by the library itself over NVLink peer memory (no NCCL
 * in the loop).  Every rank allocates (kt_ipc_alloc, zero-initialised once, never reset) a comm block of
 * kt_p2p_comm_block_bytes() bytes and a halo-receive buffer of kt_p2p_halo_slots() x 2 sides x pitch reals, and passes device
 * pointers, mapped into ITS address space (kt_ipc_open), of every rank's comm block and of its two neighbours'
 * halo-receive buffers: peer_halo_recv[0] = lower neighbour (rank-1) or NULL, [1] = upper neighbour or NULL.  The field
 * planes themselves are ordinary allocations that no other rank maps.  By default the communication is FUSED into the compute
 * kernels: the kernel that produces p / v / r stores its boundary rows straight into the neighbour's receive buffer (peer
 * stores over NVLink), the kernel that consumes them reads the halo rows from there behind a flag wait, and the finalising
 * block of every reducing kernel is the all-reduce (sums as tagged 8-byte words in every rank's comm block, added in rank
 * order).  KT_P2P_FUSED=0 in the environment selects separate one-block comm kernels instead (same results, bit for bit).
 * local_dots: 2 doubles of device scratch.  All ranks must call this with the same eps / max_iter / flags, in lock step;
 * a rank whose peer never signals returns *status = KT_STATUS_COMM_TIMEOUT after 30 s (KT_P2P_TIMEOUT_MS) with the flag and
 * sequence number in kt_last_error.  With Minv, the plane must be valid on the halo rows too (kt_p2p_exchange it after
 * assembling it). */
typedef struct {
  int rank, nranks;
  void* peer_halo_recv[2];
  void* halo_recv;
  void* comm_block[16];
  void* local_dots;
} kt_p2p_comm;
int kt_p2p_comm_block_bytes(void);
/* Buffers that other ranks (processes) map and write into: a dedicated zero-initialised cudaMalloc allocation plus its
 * 64-byte cudaIpcMemHandle_t; kt_ipc_open maps a peer's handle into this rank's device (lazy peer access). */
int kt_ipc_alloc(kt_handle_t h, size_t bytes, void** dev_ptr, unsigned char* handle64);
int kt_ipc_open(kt_handle_t h, const unsigned char* handle64, void** mapped);
int kt_ipc_close(kt_handle_t h, void* mapped);
int kt_ipc_free(kt_handle_t h, void* dev_ptr);
/* enable direct loads/stores from the handle's device to peer_device's memory (idempotent) */
int kt_enable_peer_access(kt_handle_t h, int peer_device);
/* BICGSolver.solve over x-slabs.  Minv / flags as in kt_bicgstab_solve: with Minv and KT_FLAG_WARM_START |
 * KT_FLAG_OMEGA_ZERO_BREAK it is the monolith's Jacobi-preconditioned, warm-started kernel (karman_taichi.py:342-456;
 * ws->r_tld required; the halo rows of x0 are exchanged first). */
int kt_p2p_bicgstab_solve(kt_handle_t h, const void* coef, const void* b, void* x, const void* Minv, const kt_krylov_ws* ws,
                          const kt_p2p_comm* comm, double eps, long max_iter, int flags, int fixed_iters,
                          long* iters, double* rel_res, int* status);
/* The SIMPLE kernels around the solves on x-slabs: exchange the boundary rows of one plane with the two neighbours (the
 * handle's owned rows, kt_set_rows; asynchronous on the handle's stream), and all-reduce two doubles over the ranks in rank
 * order (op 0: sum -- deterministic, the same bits on every rank; op 1: max), returned to the host.  Both use flag slots /
 * sum groups of their own, so they may be freely interleaved with solves.  kt_p2p_halo_slots(): rows x 2 sides the
 * halo-receive buffer must hold (kt_p2p_halo_slots() x 2 x pitch reals). */
int kt_p2p_halo_slots(void);
int kt_p2p_exchange(kt_handle_t h, const kt_p2p_comm* comm, void* plane);
/* the two values: vals_dev2 (device memory) or, when it is NULL, vals_host2 (host memory; they travel as kernel arguments) */
int kt_p2p_allreduce(kt_handle_t h, const kt_p2p_comm* comm, const double* vals_dev2, const double* vals_host2, int op, double* out_host2);
/* CGSolver.solve (cgsolver.py:83-114) over x-slabs, same comm kernels: one halo exchange (p) and two all-reduces per
 * iteration.  ws: r, p, Ap. */
int kt_p2p_cg_solve(kt_handle_t h, const void* coef, const void* b, void* x, const kt_krylov_ws* ws,
                    const kt_p2p_comm* comm, double eps, long max_iter, int fixed_iters,
                    long* iters, double* rel_res, int* status);

/* ---- summation plans (test infrastructure; host-only functions, no device needed) ------------------------------------
 * Every inner product of the library is deterministic: its summation order is a function of the geometry alone.  These
 * two calls publish that function so that a CPU oracle can add in exactly the device's order and must then reproduce the
 * iterates and iteration counts of a solve BIT FOR BIT (oracle/device_order.py, tests/test_gpu_exact.py).  The reference's
 * own `sum +=` reductions (cgsolver.py:49-54) leave the order unspecified.
 * kt_reduce_plan: launch geometry of the streaming kernel that computes the inner product(s) `kind`:
 *   out12 = {family (0 element-wise, 1 stencil), grid, block, V (reals per 16-byte vector),
 *            element-wise: v_lo, v_hi (vector-index range) | stencil: W, TI, nCC, nTiles, row_lo, row_hi, pitch}.
 * kt_resident_plan: which solver kt_bicgstab_solve / kt_cg_solve picks for the geometry and how it splits it:
 *   out8 = {path (0 streaming kernels, 1 cluster / registers, 2 cluster / shared memory, 3 cooperative), CTAs, rows per
 *           CTA, threads per CTA, elements per thread}.  enable = the kt_set_cluster value. */
typedef struct { int NX, NY, pitch, row_lo, row_hi, dtype, num_sms; } kt_geom;
enum { KT_PLAN_DOT = 0,          /* kt_dot */
       KT_PLAN_INIT = 1,         /* (r,r) of init            cgsolver.py:85, bicgsolver.py:90 */
       KT_PLAN_BICG_V = 2,       /* (r_tld, v)               bicgsolver.py:107 */
       KT_PLAN_BICG_ST = 3,      /* (t,s), (t,t)             bicgsolver.py:116-117 */
       KT_PLAN_BICG_XR = 4,      /* (r,r), (r,r_tld) inside kt_bicgstab_solve   bicgsolver.py:123,95 */
       KT_PLAN_BICG_XR_UNIT = 5, /* the same through kt_bicg_xr / the x-slab phases */
       KT_PLAN_RESID = 6,        /* (r,r) of the warm-start residual            karman_taichi.py:376-378 */
       KT_PLAN_CG_AP = 7,        /* (p, Ap)                  cgsolver.py:95 */
       KT_PLAN_CG_XR = 8,        /* (r,r)                    cgsolver.py:101 */
       KT_PLAN_BICG_PV = 9 };    /* (r_tld, v) of the fused p-update + product variant */
int kt_reduce_plan(const kt_geom* g, int kind, int pre, long long* out12);
int kt_resident_plan(const kt_geom* g, int is_cg, int pre, int warm, int enable, long long* out8);
int kt_get_geom(kt_handle_t h, kt_geom* out);

/* ---- per-kernel device timing (bench.py roofline) ----------------------------------------------
 * Between kt_prof_begin and kt_prof_end every launch of the solver kernels is bracketed by a CUDA event pair on
 * the handle's stream; kt_prof_end returns the summed duration and launch count per kernel class
 * (0 bicg p-update, 1 bicg v=Ap+dot, 2 bicg s,t+2 dots, 3 bicg x,r+2 dots, 4 cg Ap+dot, 5 cg x,r+dot, 6 cg p). */
int kt_prof_begin(kt_handle_t h, int max_launches);
int kt_prof_end(kt_handle_t h, double* total_ms, long* count, int n_classes);

/* ---- SIMPLE (cavity) kernels: simple.py / bicgsimp.py -------------------------------------- */
typedef struct {
  int nx, ny;               /* cells */
  double dx, dy, dt, rho, mu;
  double alpha_p, alpha_u, alpha_m;
  double bc_w[2], bc_e[2], bc_n[2], bc_s[2]; /* {normal, tangential}  (simple.py:39) */
  int pin_p;                /* 1: bicgsimp.py:146-151 pin at (1,1); 0: simple.py:165-170 */
  int east_wall_bug;        /* 1: simple.py:215 adds aw instead of ae */
  int pitch_u, pitch_v, pitch_p; /* row pitches of the (nx+3,ny+2), (nx+2,ny+3), (nx+2,ny+2) planes */
} kt_simple_params;

/* compute_coef_u + compute_coef_v (bicgsimp.py:79-112) */
int kt_assemble_uv(kt_handle_t h, const kt_simple_params* P, const void* u, const void* v, const void* p,
                   const void* u0, const void* v0, void* coef_u, void* b_u, void* coef_v, void* b_v);
/* set_bc (bicgsimp.py:153-197); also writes u[1,:], u[nx+1,:], v[:,1], v[:,ny+1] */
int kt_set_bc_cavity(kt_handle_t h, const kt_simple_params* P, void* u, void* v,
                     void* coef_u, void* b_u, void* coef_v, void* b_v);
/* compute_coef_p (bicgsimp.py:124-151): mdiv, b_p, coef_p (+ pin) */
int kt_assemble_p(kt_handle_t h, const kt_simple_params* P, const void* u, const void* v,
                  const void* coef_u, const void* coef_v, void* coef_p, void* b_p, void* mdiv);
/* update_velocity (bicgsimp.py:212-225): *residual_host = sqrt(max|du|^2 + max|dv|^2) */
int kt_relax_uv(kt_handle_t h, const kt_simple_params* P, void* u, void* v, const void* u_mid,
                const void* v_mid, double* residual_host);
/* correct_pressure + correct_velocity (bicgsimp.py:232-244) in one pass.  p == NULL skips the
 * pressure update; u == v == NULL skips the velocity update (the reference's two separate kernels). */
int kt_correct_puv(kt_handle_t h, const kt_simple_params* P, void* p, void* u, void* v, const void* pcor,
                   const void* coef_u, const void* coef_v);
/* compute_mdiv (bicgsimp.py:114-122): *max_host = max |mdiv| */
int kt_mdiv_max(kt_handle_t h, const kt_simple_params* P, const void* u, const void* v, void* mdiv,
                double* max_host);

/* ---- channel + square-cell cylinder: the monolith karman_taichi.py in stencil form -------------------------------
 * Unknowns: u[i,j] i in [1,nx+1], j in [1,ny] of the (nx+3, ny+2) plane; v[i,j] i in [1,nx], j in [1,ny+1] of
 * (nx+2, ny+3); pcor[i,j] i in [1,nx], j in [1,ny] of (nx+2, ny+2).  ct: int32 (nx+2, ny+2) plane with pitch_p,
 * 1 solid / -1 fluid (karman_taichi.py:143-150).  M_* = the reference's Mu/Mv/Mp diagonal planes, Minv_* = 1/M_*
 * (what kt_bicgstab_solve takes). */
typedef struct {
  int nx, ny;
  double dx, dy, dt, rho, mu, velo_rel, p_rel;
  int cyl_cx, cyl_cy, cyl_r2;       /* (i-cx)^2 + (j-cy)^2 < r2 -> solid   (karman_taichi.py:148-149) */
  int pitch_u, pitch_v, pitch_p;
  /* x-slab window (no reference counterpart; all zero = one GPU, the whole grid): the planes passed to a call hold the GLOBAL
   * rows [row0, row0 + rows_*) -- this rank's owned rows plus one halo row on each interior side -- and the call works on the
   * global rows [row_lo, row_hi).  rows_u / rows_p: stored rows of the u-type (nx+3 globally) and of the v- and cell-centred
   * planes (nx+2 globally).  nx, ny stay the GLOBAL cell counts: every boundary test of the reference is on the global index. */
  int row0, row_lo, row_hi, rows_u, rows_p;
} kt_channel_params;

/* init (karman_taichi.py:130-154) */
int kt_channel_init(kt_handle_t h, const kt_channel_params* P, void* p, void* u, void* u0, void* v, void* v0, int* ct);
/* fill_Au + fill_Av (karman_taichi.py:161-294) */
int kt_channel_assemble_uv(kt_handle_t h, const kt_channel_params* P, const int* ct, const void* u, const void* v,
                           const void* p, const void* u0, const void* v0, void* coef_u, void* b_u, void* M_u,
                           void* Minv_u, void* coef_v, void* b_v, void* M_v, void* Minv_v);
/* fill_Ap (karman_taichi.py:296-339) */
int kt_channel_assemble_p(kt_handle_t h, const kt_channel_params* P, const int* ct, const void* u, const void* v,
                          const void* coef_u, const void* coef_v, void* coef_p, void* b_p, void* M_p, void* Minv_p);
/* xu_back + xv_back (karman_taichi.py:458-470): u += velo_rel (x_u - u), v likewise */
int kt_channel_relax(kt_handle_t h, const kt_channel_params* P, void* u, void* v, const void* x_u, const void* x_v);
/* correct_conserv (karman_taichi.py:585-614): scale u[nx+1,:] by sum u[1,:] / sum u[nx+1,:]; udiv (may be NULL);
 * sums_host (may be NULL) receives {inlet sum, outlet sum before scaling, ratio} */
int kt_outlet_scale(kt_handle_t h, const kt_channel_params* P, void* u, const void* v, void* udiv, double* sums_host);
/* correct_conserv on x-slabs: this rank's {inlet, outlet} column sums into sums_dev[0..1] (0 where it does not own the
 * column; add the ranks' pairs in rank order with kt_p2p_allreduce), then the scaling of the outlet column with
 * inlet_sum / outlet_sum on the rank that owns it, then udiv on the window (needs the halo row of u above it). */
int kt_outlet_sums(kt_handle_t h, const kt_channel_params* P, const void* u, double* sums_dev);
int kt_outlet_apply(kt_handle_t h, const kt_channel_params* P, void* u, double inlet_sum, double outlet_sum);
int kt_channel_udiv(kt_handle_t h, const kt_channel_params* P, const void* u, const void* v, void* udiv);
/* puv_correction (karman_taichi.py:506-542): returns max |pcor| (over the window) */
int kt_channel_correct(kt_handle_t h, const kt_channel_params* P, const int* ct, void* u, void* v, void* p,
                       const void* pcor, const void* coef_u, const void* coef_v, double* pcor_max_host);

/* ---- either side of the hot path -------------------------------------------------------------- */
/* The reference's display frame.  scale = 1, monolith = 0: Display.post_process_field + display (display.py:17-49):
 * panels (u, v, p, pcor, mdiv), each scaled by its own min / max (f_max starting at -1e9), img = (nx+2) x 5 (ny+2).
 * scale = 3, monolith = 1: post_process_field + the ti.GUI frame of karman_taichi.py:544-583,662-665: udiv recomputed in
 * the interior (and stored to div, as the reference does), every panel upsampled x3, f_max starting at 0.0, panel order
 * (pcor, udiv, p, u, v), img = 3 (nx+2) x 15 (ny+2).  img: dense row-major device array of reals, img[I][J] with I along x.
 * minmax_host10 (may be NULL): {max u, v, p, pcor, div, min u, v, p, pcor, div}. */
typedef struct {
  int nx, ny, pitch_u, pitch_v, pitch_p;
  int scale, monolith;
  double dx, dy;
} kt_frame_params;
int kt_display_frame(kt_handle_t h, const kt_frame_params* P, const void* u, const void* v, const void* p, const void* pcor,
                     void* div, void* img, double* minmax_host10);
/* One point-Jacobi sweep of the 5-point system (coef, rhs) over rows [i_lo, i_hi) x columns [j_lo, j_hi) of x -- the
 * alternative solver of the reference's validation scripts: jacobian_solve_p (pressure_valid.py:220-232, relax = 1) and
 * jacobian_solve_u (momentum_valid.py:157-170, relax = alpha_u).  x_mid: scratch plane (the reference's pcor_mid / u_mid);
 * *residual_host = max |x - x_mid| before the update.  Every read of the sweep sees the OLD iterate (the reference updates
 * in place in whatever order its threads run). */
int kt_jacobi_sweep(kt_handle_t h, const void* coef, const void* rhs, void* x, void* x_mid, int i_lo, int i_hi, int j_lo,
                    int j_hi, double relax, double* residual_host);

#ifdef __cplusplus
}
#endif
#endif /* KARMAN_B200_H */
