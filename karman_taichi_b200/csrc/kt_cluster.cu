// kt_cluster.cu -- small-grid BiCGSTAB / CG that lives entirely in ONE thread-block cluster (sm_100a).
//
// The reference's own grids are small (cavity 30^2..128^2, Karman 320x64): ~2e4 unknowns.  At that size an iteration
// of the streaming path is four kernel launches of a few microseconds each and the solve is launch-latency bound
// (26-37 us per Krylov iteration measured, profiles/r01_simple_rates.md).  Here the whole solve is one launch: a
// cluster of up to 16 CTAs holds every plane (5 coefficients, b, x, r, p, v, s, t [, 1/M, r~]) in its shared memory
// (16 x 227 KB), each CTA owning a contiguous block of rows.  Same recurrence, same operation order per element, same
// scalar algebra (kt::scalar_algebra) as the streaming kernels.
//
// All traffic between CTAs is PUSHED with st.async (DSMEM store that signals an mbarrier in the destination CTA with its
// byte count) -- there is no cluster barrier and no remote load inside the iteration (measured, profiles/
// r01_simple_rates.md: a cluster barrier that has to release 16 remote stores costs ~1000 cycles, a remote load ~215,
// an mbarrier wake-up ~60):
//   * halo rows: the thread that produces an element of a boundary row of p, v or r also stores it into the
//     neighbour's private halo row; the neighbour waits on that plane's halo mbarrier right before its first use;
//   * inner products: block tree -> lanes 0..C-1 of warp 0 st.async the block's (d0, d1) into slot [rank] of every CTA
//     and signal that CTA's sum mbarrier; every CTA waits for C x 16 bytes and adds the C slots with a fixed tree, so
//     all CTAs, warps and threads hold the identical bits and run the scalar algebra redundantly.
// Every CTA needs every other CTA's contribution to pass a sum, so a sum is also a full synchronisation point; the
// write-after-read safety of the halo rows and slots follows from that (see the comments at the pushes).
//
// BiCGSTAB per iteration (bicgsolver.py:94-135 / karman_taichi.py:383-456):
//   P1 p = r + beta (p - omega v), boundary rows pushed            | wait p halo
//   P2 v = A p^ ; (r~, v), boundary rows of v pushed               | sum -> alpha
//   P3 t = A s^ with s = r - alpha v (halo rows from the r, v halos) ; (t,s), (t,t) | sum -> omega
//   P4 x += alpha p^ + omega s^ ; r = s - omega t, boundary rows pushed ; (r,r), (r,r~) | sum -> test, rho, beta
// CG per iteration (cgsolver.py:92-114): Ap,(p,Ap) | sum | x,r,(r,r) | sum | p = r + beta p, boundary rows pushed.
#include <cooperative_groups.h>
#include "kt_handle.h"

namespace cg = cooperative_groups;

namespace kt {

template <typename T> struct ClArgs {
  const T* coef; size_t plane;     // global SoA pitched planes
  const T* b; T* x; const T* minv; T* r_out;
  int NX, NY, pitch, rows_per, nranks;
  Scalars* sc; double* history; long long hist_cap;
  double eps; long long max_iter; int flags; int warm; int is_cg;
};

constexpr int CL_THREADS = 512;

// -DKT_CLUSTER_PROF (tools/cluster_prof.py builds that variant): thread 0 of CTA 0 accumulates clock64() ticks per
// segment of an iteration into g_clprof; compiled out of the shipped library.
#ifdef KT_CLUSTER_PROF
__device__ long long g_clprof[16];
struct ClProf { long long acc[16]; long long t; };     // per-thread registers; dumped once at the end of the kernel
#define CLP_DECL ClProf clp; for (int k_ = 0; k_ < 16; ++k_) clp.acc[k_] = 0; clp.t = clock64();
#define CLP(k) do { const long long t_ = clock64(); clp.acc[k] += t_ - clp.t; clp.t = t_; } while (0)
#define CLP_ARGS , ClProf& clp
#define CLP_PASS , clp
#define CLP_DUMP do { if (blockIdx.x == 0 && threadIdx.x == 0) for (int k_ = 0; k_ < 16; ++k_) g_clprof[k_] += clp.acc[k_]; } while (0)
#else
#define CLP_DECL
#define CLP(k) do { } while (0)
#define CLP_ARGS
#define CLP_PASS
#define CLP_DUMP do { } while (0)
#endif

enum { HALO_P = 0, HALO_V = 1, HALO_R = 2 };

struct ClShared {                    // per-CTA control block at the start of dynamic shared memory
  double part[3][16][2];             // [dot group][source rank][d0, d1]: 16-byte slots filled by st.async.v2
  unsigned long long mbar_sum[3];    // one mbarrier per dot group: completes when nranks x 16 bytes have landed
  unsigned long long mbar_halo[3];   // p / v / r halo rows: completes when both neighbours' rows have landed
  double red[2][CL_THREADS / 32];
};

// ---- PTX helpers: shared::cluster addresses, mbarrier, st.async --------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned local_addr, unsigned cta_rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
// Bounded wait: a peer that never signals (a bug, not a data condition) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  unsigned ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
    if (ok) break;
    if (clock64() - t0 > (1ll << 32)) __trap();
  }
}
__device__ __forceinline__ void st_async(unsigned remote_addr, double v, unsigned remote_mbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(remote_addr), "l"(__double_as_longlong(v)), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void st_async(unsigned remote_addr, float v, unsigned remote_mbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void st_async2(unsigned remote_addr, double a, double b, unsigned remote_mbar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
               ::"r"(remote_addr), "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "r"(remote_mbar) : "memory");
}

// 16-lane fixed-shape tree: every lane of the warp ends up with v[0] + ... + v[15] (lanes >= 16 mirror lane - 16).
__device__ __forceinline__ double tree16(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fixed-order cluster-wide sum of up to two values.  Warp shuffle tree -> one value per warp in shared memory ->
// warp 0 adds the (16) warp sums with a fixed tree; its lanes 0..C-1 st.async the pair into slot [rank] of CTA `lane`
// (signalling that CTA's mbarrier with 16 bytes) -> every thread waits on its own CTA's mbarrier (C x 16 bytes) ->
// every warp adds the C slots with the same fixed tree.  Identical bits on every CTA, warp and thread.
// `use` counts the completed uses of this group's mbarrier (phase parity = use & 1).
template <int NW = CL_THREADS / 32>
__device__ __forceinline__ void cluster_sum(ClShared* S, int group, unsigned use, double& d0, double& d1, int nranks, int rank CLP_ARGS) {
  static_assert(NW <= 16, "tree16 adds at most 16 warp sums");
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  d0 = warp_sum(d0);
  d1 = warp_sum(d1);
  if (lane == 0) { S->red[0][wid] = d0; S->red[1][wid] = d1; }
  __syncthreads();
  CLP(3);
  const unsigned mbar = smem_u32(&S->mbar_sum[group]);
  if (wid == 0) {
    const double s0 = tree16((lane & 15) < NW ? S->red[0][lane & 15] : 0.0);
    const double s1 = tree16((lane & 15) < NW ? S->red[1][lane & 15] : 0.0);
    if (lane == 0) mbar_arrive_expect(mbar, 16u * (unsigned)nranks);
    if (lane < nranks) st_async2(mapa_u32(smem_u32(&S->part[group][rank][0]), lane), s0, s1, mapa_u32(mbar, lane));
  }
  CLP(4);
  mbar_wait(mbar, use & 1u);
  CLP(5);
  const int r = lane & 15;
  const double2 pr = (r < nranks) ? *reinterpret_cast<const double2*>(&S->part[group][r][0]) : make_double2(0.0, 0.0);
  d0 = tree16(pr.x);
  d1 = tree16(pr.y);
  CLP(6);
}

// ---------------------------------------------------------------------------------------------------------------
// Register-resident variant: NT = 512 threads for up to 2 elements per thread, NT = 256 threads (255 registers each) for up
// to 5 -- together every slice of up to 1280 elements (cavity <= ~140^2); larger slices take the shared-memory kernel.
// ---------------------------------------------------------------------------------------------------------------
// Per-thread element state.  A thread owns the same EPT elements of the CTA's slice (flat index e = threadIdx.x + k * 512)
// for the whole solve and keeps everything only it touches in registers: the five coefficients (REGC; otherwise they stay
// in shared memory), r~, x, r, v, p, s, t and 1/M.  Shared memory holds just the planes a NEIGHBOUR element has to read
// (p, v, r [, 1/M]) plus the private halo rows, so an iteration moves 12 loads + 3 stores per element through the
// shared-memory pipe instead of 35 + 6 -- the SpMV phases of the first versions were bound by exactly that pipe.
enum { EF_VALID = 1, EF_ROW0 = 2, EF_ROWN = 4, EF_J0 = 8, EF_JN = 16 };

template <typename T, bool PRE, int EPT, bool REGC, int NT>
__global__ void __launch_bounds__(NT, 1) k_cluster_solve(const ClArgs<T> A) {
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ClShared* S = reinterpret_cast<ClShared*>(smem_raw);
  const int rank = (int)cl.block_rank();
  const int NY = A.NY;
  const int r0 = min(rank * A.rows_per, A.NX), r1 = min(r0 + A.rows_per, A.NX);
  const int nrows = r1 - r0, n_el = nrows * NY;
  const int cap = A.rows_per * NY;                        // plane slice capacity (same on every CTA)
  T* base = reinterpret_cast<T*>(smem_raw + ((sizeof(ClShared) + 15) / 16) * 16);
  T* ps = base; T* vs = ps + cap; T* rs = vs + cap;
  T* ms = rs + cap;                                       // 1/M (PRE only)
  T* cs = PRE ? ms + cap : ms;                            // five coefficient planes (!REGC only)
  // private halo rows (row r0-1 = "lo", row r1 = "hi") of p, v, r and of the constant 1/M plane
  T* hal = REGC ? cs : cs + 5 * cap;
  T* h_lo[3] = {hal, hal + 2 * NY, hal + 4 * NY};
  T* h_hi[3] = {hal + NY, hal + 3 * NY, hal + 5 * NY};
  T* m_lo = hal + 6 * NY; T* m_hi = hal + 7 * NY;

  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k) { mbar_init(smem_u32(&S->mbar_sum[k]), 1); mbar_init(smem_u32(&S->mbar_halo[k]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- this thread's elements ---------------------------------------------------------------------------------------
  T c[REGC ? EPT : 1][5];
  T rt[EPT], x[EPT], r[EPT], v[EPT], p[EPT], t[EPT], m[EPT];
  unsigned ef[EPT];                 // EF_* flags in the low byte, the column j above it
#define EJ(k) ((int)(ef[k] >> 8))
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int e = threadIdx.x + k * NT;
    ef[k] = 0;
    x[k] = r[k] = v[k] = p[k] = t[k] = rt[k] = T(0); m[k] = T(1);
    if (REGC) { for (int q = 0; q < 5; ++q) c[k][q] = T(0); }
    if (e < n_el) {
      const int li = e / NY, j = e - li * NY;
      const size_t g = (size_t)(r0 + li) * A.pitch + j;
      ef[k] = EF_VALID | (li == 0 ? EF_ROW0 : 0) | (li == nrows - 1 ? EF_ROWN : 0) | (j == 0 ? EF_J0 : 0) | (j == NY - 1 ? EF_JN : 0) |
              ((unsigned)j << 8);
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const T cv = A.coef[q * A.plane + g];
        if (REGC) c[k][q] = cv; else cs[q * cap + e] = cv;
      }
      if (A.warm) { x[k] = A.x[g]; ps[e] = x[k]; }         // the p plane carries x0 for the warm-start residual
      if (PRE) { m[k] = A.minv[g]; ms[e] = m[k]; }
    }
  }
  auto coef = [&](int k, int q, int e) -> T { return REGC ? c[REGC ? k : 0][q] : cs[q * cap + e]; };
  // Every thread keeps its own copy of the Krylov scalars and runs the (cheap) scalar algebra itself on the identical
  // reduced sums: no shared-memory round trip and no block barrier between a reduction and its consumers.
  Scalars sc;
  sc.rho = 0.0; sc.rho_1 = 1.0; sc.alpha = 1.0; sc.beta = 1.0; sc.omega = 1.0;
  sc.rTr0 = sc.rTr = sc.old_rTr = 0.0; sc.dot0 = sc.dot1 = 0.0;
  sc.eps = A.eps; sc.iter = 0; sc.max_iter = A.max_iter; sc.status = 1; sc.done = 0; sc.first = 1; sc.flags = A.flags;
  double* const hist = (rank == 0 && threadIdx.x == 0) ? A.history : nullptr;
#define CL_SCALARS(fin, a0, a1, h_) scalar_algebra(fin, &sc, a0, a1, h_, A.hist_cap)
  cl.sync();      // mbarriers initialised and slices loaded everywhere: the one cluster barrier before the loop

  // neighbours (row r0-1 is the last row of rank-1, row r1 the first row of rank+1: every lower rank is full)
  const bool has_lo = r0 > 0 && nrows > 0, has_hi = r1 < A.NX && nrows > 0;
  const int lo_rank = rank > 0 ? rank - 1 : 0, hi_rank = rank < A.nranks - 1 ? rank + 1 : rank;
  const int lo_off = (A.rows_per - 1) * NY;               // offset of the neighbour's last row
  const unsigned halo_bytes = (unsigned)((has_lo ? 1 : 0) + (has_hi ? 1 : 0)) * (unsigned)NY * (unsigned)sizeof(T);
  unsigned use_sum[3] = {0, 0, 0}, use_halo[3] = {0, 0, 0};
  bool pend[3] = {false, false, false};   // rows of this plane were pushed (by everybody, same control flow) and not yet waited for
  // my first row lands in the lower neighbour's "hi" halo row, my last row in the upper neighbour's "lo" halo row
  // (plane `which` sits 2 * NY elements / one mbarrier further than plane 0)
  const unsigned dst_lo0 = mapa_u32(smem_u32(h_hi[0]), lo_rank), dst_hi0 = mapa_u32(smem_u32(h_lo[0]), hi_rank);
  const unsigned mb_lo0 = mapa_u32(smem_u32(&S->mbar_halo[0]), lo_rank), mb_hi0 = mapa_u32(smem_u32(&S->mbar_halo[0]), hi_rank);
  // push one element of a boundary row to the neighbour that needs it (called by the thread that produced it)
  auto push = [&](int which, unsigned flags, T val) {
    const unsigned off = (unsigned)((which * 2 * NY + (int)(flags >> 8)) * sizeof(T));
    if ((flags & EF_ROW0) && has_lo) st_async(dst_lo0 + off, val, mb_lo0 + 8u * which);
    if ((flags & EF_ROWN) && has_hi) st_async(dst_hi0 + off, val, mb_hi0 + 8u * which);
  };
  // wait until both neighbours' rows of this plane have landed (thread 0 posts the byte count, everybody waits)
  auto halo_wait = [&](int which) {
    pend[which] = false;
    if (halo_bytes == 0) return;
    const unsigned mbar = smem_u32(&S->mbar_halo[which]);
    if (threadIdx.x == 0) mbar_arrive_expect(mbar, halo_bytes);
    mbar_wait(mbar, use_halo[which] & 1u);
    ++use_halo[which];
  };
  // the constant 1/M halo rows and the warm-start x halo are read once, remotely (slices final since cl.sync)
  if (PRE) {
    const T* ms_lo = cl.map_shared_rank(ms, lo_rank); const T* ms_hi = cl.map_shared_rank(ms, hi_rank);
    for (int j = threadIdx.x; j < NY; j += NT) {
      if (has_lo) m_lo[j] = ms_lo[lo_off + j];
      if (has_hi) m_hi[j] = ms_hi[j];
    }
  }
  const T* x0_lo = cl.map_shared_rank(ps, lo_rank) + lo_off; const T* x0_hi = cl.map_shared_rank(ps, hi_rank);

  // ---- stencil product y = A * hat(src) for this thread's element k; src selected by MODE ---------------------------
  //   MODE 0: src = p                  hat = 1/M * p     (neighbours: p plane / pushed p rows)
  //   MODE 1: src = s = r - alpha v    hat = 1/M * s     (neighbours recomputed from the r, v planes / pushed rows)
  //   MODE 2: src = x0 in the p plane (warm-start residual, karman_taichi.py:360-367: no preconditioner; halo read
  //           remotely, once)
  auto nb_in = [&](int mode, int idx, T alpha) -> T {     // neighbour inside this CTA's slice
    T v_ = (mode == 1) ? (rs[idx] - alpha * vs[idx]) : ps[idx];
    if (PRE && mode != 2) v_ = ms[idx] * v_;
    return v_;
  };
  auto nb_halo = [&](int mode, bool lo, int j, T alpha) -> T {
    T v_;
    if (mode == 0) v_ = lo ? h_lo[HALO_P][j] : h_hi[HALO_P][j];
    else if (mode == 1) v_ = lo ? (h_lo[HALO_R][j] - alpha * h_lo[HALO_V][j]) : (h_hi[HALO_R][j] - alpha * h_hi[HALO_V][j]);
    else v_ = lo ? x0_lo[j] : x0_hi[j];
    if (PRE && mode != 2) v_ = (lo ? m_lo[j] : m_hi[j]) * v_;
    return v_;
  };
  auto spmv_elem = [&](int mode, int k, T centre_raw, T alpha) -> T {
    const int e = threadIdx.x + k * NT, j = EJ(k);
    const unsigned f = ef[k];
    const T xc = (PRE && mode != 2) ? m[k] * centre_raw : centre_raw;
    T xw = T(0), xe = T(0), xn = T(0), xso = T(0);
    if (!(f & EF_ROW0)) xw = nb_in(mode, e - NY, alpha);
    else if (has_lo) xw = nb_halo(mode, true, j, alpha);
    if (!(f & EF_ROWN)) xe = nb_in(mode, e + NY, alpha);
    else if (has_hi) xe = nb_halo(mode, false, j, alpha);
    if (!(f & EF_JN)) xn = nb_in(mode, e + 1, alpha);
    if (!(f & EF_J0)) xso = nb_in(mode, e - 1, alpha);
    T acc;
    if (PRE || mode == 2) {   // the monolith's product (karman_taichi.py:362-367,401-406): flat offsets i-ny, i-1, i, i+1, i+ny from 0.0
      acc = T(0) + coef(k, 1, e) * xw;
      acc = acc + coef(k, 4, e) * xso;
      acc = acc + coef(k, 0, e) * xc;
      acc = acc + coef(k, 3, e) * xn;
      acc = acc + coef(k, 2, e) * xe;
    } else {                    // cgsolver.py:59-60, left to right
      acc = coef(k, 0, e) * xc;
      acc = acc + coef(k, 1, e) * xw;
      acc = acc + coef(k, 2, e) * xe;
      acc = acc + coef(k, 3, e) * xn;
      acc = acc + coef(k, 4, e) * xso;
    }
    return acc;
  };
#define CL_EACH(k, e) _Pragma("unroll") for (int k = 0; k < EPT; ++k) if (const int e = threadIdx.x + k * NT; ef[k] & EF_VALID)

  // ---- K0: r = b - A x0 (warm) or r = b ; r~ = r ; (r,r) -----------------------------------------------------------
  CLP_DECL
  double d0 = 0.0, d1 = 0.0;
  CL_EACH(k, e) {
    const int li_ = e / NY;
    T rr = A.b[(size_t)(r0 + li_) * A.pitch + (e - li_ * NY)];
    if (A.warm) rr = rr - spmv_elem(2, k, x[k], T(0));
    r[k] = rr; rt[k] = rr;                                 // r~ = r0 (= b when x0 = 0): bicgsolver.py:28, karman_taichi.py:371
    if (A.is_cg) p[k] = rr;                                // cgsolver.py:90 p = r + 0 * p
    d0 += (double)rr * (double)rr;
  }
  if (A.warm) {
    // everybody must be done reading x0 out of the p planes (locally and remotely) before anything else lands there
    cl.sync();
  }
  CL_EACH(k, e) {
    if (A.is_cg) { ps[e] = p[k]; push(HALO_P, ef[k], p[k]); }
    else { rs[e] = r[k]; push(HALO_R, ef[k], r[k]); }   // r halo: read by P3 of the first iteration
  }
  pend[A.is_cg ? HALO_P : HALO_R] = true;
  // group 2 for this one: the loop's first sum (group 0) may be sent by a fast CTA before a slow one has read the slots
  // of the sum before it, unless another sum lies in between -- group 2 is next used two sums later (P4) or never (CG)
  cluster_sum<NT / 32>(S, 2, use_sum[2]++, d0, d1, A.nranks, rank CLP_PASS);
  CL_SCALARS(A.is_cg ? FIN_CG_INIT : FIN_BICG_INIT, d0, d0, hist);
  if (A.is_cg) { halo_wait(HALO_P); __syncthreads(); }

  // ---- iterations ----------------------------------------------------------------------------------------------
  CLP(15);
  while (!sc.done) {
    if (!A.is_cg) {
      // P1.  The neighbours' p halo rows were last read in their P2 of the previous iteration, which they finished
      // before contributing to that iteration's sums -- and this CTA is past those sums: pushing is safe.
      const bool first = sc.iter == 0;
      const T beta = T(sc.beta), omega_prev = T(sc.omega);
      CL_EACH(k, e) {
        const T pn = first ? r[k] : r[k] + beta * (p[k] - omega_prev * v[k]);
        p[k] = pn; ps[e] = pn;
        push(HALO_P, ef[k], pn);
      }
      pend[HALO_P] = true;
      CLP(0);
      halo_wait(HALO_P);
      __syncthreads();                                     // own p rows visible to the whole CTA
      CLP(1);
      // P2
      d0 = 0.0; d1 = 0.0;
      CL_EACH(k, e) {
        const T y = spmv_elem(0, k, p[k], T(0));
        d0 += (double)rt[k] * (double)y;
        v[k] = y; vs[e] = y;                               // the v plane was last read in P3, two sums ago
        push(HALO_V, ef[k], y);
      }
      pend[HALO_V] = true;
      CLP(2);
      cluster_sum<NT / 32>(S, 0, use_sum[0]++, d0, d1, A.nranks, rank CLP_PASS);
      CL_SCALARS(FIN_BICG_ALPHA, d0, d1, nullptr);
      CLP(7);
      if (sc.done) break;
      // P3 (own v, r planes: visible since the __syncthreads inside the sums; halo rows of v and r: wait for them)
      halo_wait(HALO_V);
      halo_wait(HALO_R);
      const T alpha = T(sc.alpha);
      d0 = 0.0; d1 = 0.0;
      CL_EACH(k, e) {
        const T sraw = r[k] - alpha * v[k];                // bicgsolver.py:63
        const T y = spmv_elem(1, k, sraw, alpha);
        t[k] = y;
        d0 += (double)y * (double)sraw;
        d1 += (double)y * (double)y;
      }
      CLP(8);
      cluster_sum<NT / 32>(S, 1, use_sum[1]++, d0, d1, A.nranks, rank CLP_PASS);
      CL_SCALARS(FIN_BICG_OMEGA, d0, d1, nullptr);
      CLP(7);
      if (sc.done) break;
      // P4 (the neighbours read their r halo in P3, which they finished before contributing to the sum above)
      const T omega = T(sc.omega);
      d0 = 0.0; d1 = 0.0;
      CL_EACH(k, e) {
        const T sk = r[k] - alpha * v[k];                  // s again (same operation, same bits as in P3): one register less
        const T ph = PRE ? m[k] * p[k] : p[k];
        const T sh = PRE ? m[k] * sk : sk;
        x[k] = x[k] + (alpha * ph + omega * sh);
        const T rn = sk - omega * t[k];
        r[k] = rn; rs[e] = rn;
        push(HALO_R, ef[k], rn);
        d0 += (double)rn * (double)rn;
        d1 += (double)rn * (double)rt[k];
      }
      pend[HALO_R] = true;
      CLP(9);
      cluster_sum<NT / 32>(S, 2, use_sum[2]++, d0, d1, A.nranks, rank CLP_PASS);
      CL_SCALARS(FIN_BICG_END, d0, d1, hist);
      CLP(7);
    } else {
      // C1: Ap = A p ; (p, Ap)   (Ap is needed by its owner only: registers)
      d0 = 0.0; d1 = 0.0;
      CL_EACH(k, e) {
        const T y = spmv_elem(0, k, p[k], T(0));
        v[k] = y;
        d0 += (double)p[k] * (double)y;
      }
      cluster_sum<NT / 32>(S, 0, use_sum[0]++, d0, d1, A.nranks, rank CLP_PASS);
      CL_SCALARS(FIN_CG_ALPHA, d0, d1, nullptr);
      if (sc.done) break;
      // C2: x += alpha p ; r -= alpha Ap ; (r,r)
      const T alpha = T(sc.alpha);
      d0 = 0.0; d1 = 0.0;
      CL_EACH(k, e) {
        x[k] = x[k] + alpha * p[k];
        const T rn = r[k] - alpha * v[k];
        r[k] = rn;
        d0 += (double)rn * (double)rn;
      }
      cluster_sum<NT / 32>(S, 1, use_sum[1]++, d0, d1, A.nranks, rank CLP_PASS);
      CL_SCALARS(FIN_CG_END, d0, d1, hist);
      if (sc.done) break;
      // C3: p = r + beta p (the neighbours finished reading their p halo in C1, two sums ago)
      const T beta = T(sc.beta);
      CL_EACH(k, e) {
        const T pn = r[k] + beta * p[k];
        p[k] = pn; ps[e] = pn;
        push(HALO_P, ef[k], pn);
      }
      pend[HALO_P] = true;
      halo_wait(HALO_P);
      __syncthreads();
    }
  }
  // Every CTA leaves the loop at the same point (identical scalars).  Rows pushed by a neighbour may still be in flight
  // towards this CTA (r of the last P4, v of an aborted iteration): absorb them before this CTA exits and frees the
  // shared memory they target; the barrier keeps every slice alive until all CTAs are past their last remote access.
#pragma unroll
  for (int k = 0; k < 3; ++k) if (pend[k]) halo_wait(k);
  cl.sync();
  CL_EACH(k, e) {
    const int li_ = e / NY;
    const size_t g = (size_t)(r0 + li_) * A.pitch + (e - li_ * NY);
    A.x[g] = x[k];
    if (A.r_out) A.r_out[g] = r[k];
  }
  if (rank == 0 && threadIdx.x == 0) {
    Scalars out = sc;
    out.dot0 = out.dot1 = 0.0;          // (not kept live through the loop: nobody reads the raw sums of a cluster solve)
    *A.sc = out;
  }
  CLP_DUMP;
#undef CL_EACH
#undef CL_SCALARS
#undef EJ
}

// ---------------------------------------------------------------------------------------------------------------
// Variant for 3+ elements per thread: every plane in shared memory (12 planes + halo rows), row-group/column thread map.
// At that size the register-resident kernel above spills and is slower (profiles/r01_simple_rates.md).
// ---------------------------------------------------------------------------------------------------------------
template <typename T, bool PRE>
__global__ void __launch_bounds__(CL_THREADS, 1) k_cluster_solve_smem(const ClArgs<T> A) {
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ClShared* S = reinterpret_cast<ClShared*>(smem_raw);
  const int rank = (int)cl.block_rank();
  const int NY = A.NY;
  const int r0 = min(rank * A.rows_per, A.NX), r1 = min(r0 + A.rows_per, A.NX);
  const int nrows = r1 - r0;
  const int cap = A.rows_per * NY;                        // plane slice capacity (same on every CTA)
  T* base = reinterpret_cast<T*>(smem_raw + ((sizeof(ClShared) + 15) / 16) * 16);
  T* c0 = base; T* c1 = c0 + cap; T* c2 = c1 + cap; T* c3 = c2 + cap; T* c4 = c3 + cap;
  T* bs = c4 + cap; T* xs = bs + cap; T* rs = xs + cap; T* ps = rs + cap; T* vs = ps + cap;
  T* ss = vs + cap; T* ts = ss + cap;
  T* ms = ts + cap;                                       // 1/M (PRE only)
  T* qs = PRE ? ms + cap : ms;                            // r~ when it differs from b (warm start)
  const bool own_rt = A.warm != 0;
  const T* rt = own_rt ? qs : bs;
  // private halo rows (row r0-1 = "lo", row r1 = "hi") of p, v, r and of the constant 1/M plane
  T* hal = qs + (own_rt ? cap : 0);
  T* h_lo[3] = {hal, hal + 2 * NY, hal + 4 * NY};
  T* h_hi[3] = {hal + NY, hal + 3 * NY, hal + 5 * NY};
  T* m_lo = hal + 6 * NY; T* m_hi = hal + 7 * NY;

  // thread -> (row group, column) map fixed for the whole solve: no integer division inside the phases.
  const int TC = min(NY, CL_THREADS), TR = CL_THREADS / TC;
  const int tx = threadIdx.x % TC;
  const int ty = (threadIdx.x / TC < TR) ? threadIdx.x / TC : nrows;     // surplus threads own no rows
#define CL_FOR(li, j, e) \
  for (int j = tx; j < NY; j += TC) for (int li = ty, e = ty * NY + j; li < nrows; li += TR, e += TR * NY)

  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; ++k) { mbar_init(smem_u32(&S->mbar_sum[k]), 1); mbar_init(smem_u32(&S->mbar_halo[k]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- load this CTA's rows into shared memory (compact rows of NY) ----------------------------------------------
  CL_FOR(li, j, e) {
    const size_t g = (size_t)(r0 + li) * A.pitch + j;
    c0[e] = A.coef[g]; c1[e] = A.coef[A.plane + g]; c2[e] = A.coef[2 * A.plane + g];
    c3[e] = A.coef[3 * A.plane + g]; c4[e] = A.coef[4 * A.plane + g];
    bs[e] = A.b[g];
    xs[e] = A.warm ? A.x[g] : T(0);
    ps[e] = T(0); vs[e] = T(0); ss[e] = T(0); ts[e] = T(0);
    if (PRE) ms[e] = A.minv[g];
  }
  // Every thread keeps its own copy of the Krylov scalars and runs the (cheap) scalar algebra itself on the identical
  // reduced sums: no shared-memory round trip and no block barrier between a reduction and its consumers.
  Scalars sc;
  sc.rho = 0.0; sc.rho_1 = 1.0; sc.alpha = 1.0; sc.beta = 1.0; sc.omega = 1.0;
  sc.rTr0 = sc.rTr = sc.old_rTr = 0.0; sc.dot0 = sc.dot1 = 0.0;
  sc.eps = A.eps; sc.iter = 0; sc.max_iter = A.max_iter; sc.status = 1; sc.done = 0; sc.first = 1; sc.flags = A.flags;
  double* const hist = (rank == 0 && threadIdx.x == 0) ? A.history : nullptr;
  cl.sync();      // mbarriers initialised and slices loaded everywhere: the one cluster barrier before the loop

  // neighbours (row r0-1 is the last row of rank-1, row r1 the first row of rank+1: every lower rank is full)
  const bool has_lo = r0 > 0 && nrows > 0, has_hi = r1 < A.NX && nrows > 0;
  const int lo_rank = rank > 0 ? rank - 1 : 0, hi_rank = rank < A.nranks - 1 ? rank + 1 : rank;
  const int lo_off = (A.rows_per - 1) * NY;               // offset of the neighbour's last row
  const unsigned halo_bytes = (unsigned)((has_lo ? 1 : 0) + (has_hi ? 1 : 0)) * (unsigned)NY * (unsigned)sizeof(T);
  unsigned use_sum[3] = {0, 0, 0}, use_halo[3] = {0, 0, 0};
  bool pend[3] = {false, false, false};   // rows of this plane were pushed (by everybody, same control flow) and not yet waited for
  // my first row lands in the lower neighbour's "hi" halo row, my last row in the upper neighbour's "lo" halo row
  unsigned dst_lo[3], dst_hi[3], mb_lo[3], mb_hi[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dst_lo[k] = mapa_u32(smem_u32(h_hi[k]), lo_rank); dst_hi[k] = mapa_u32(smem_u32(h_lo[k]), hi_rank);
    mb_lo[k] = mapa_u32(smem_u32(&S->mbar_halo[k]), lo_rank); mb_hi[k] = mapa_u32(smem_u32(&S->mbar_halo[k]), hi_rank);
  }
  // push one element of a boundary row to the neighbour that needs it (called by the thread that produced it)
  auto push = [&](int which, int li, int j, T val) {
    if (li == 0 && has_lo) st_async(dst_lo[which] + (unsigned)(j * sizeof(T)), val, mb_lo[which]);
    if (li == nrows - 1 && has_hi) st_async(dst_hi[which] + (unsigned)(j * sizeof(T)), val, mb_hi[which]);
  };
  // wait until both neighbours' rows of this plane have landed (thread 0 posts the byte count, everybody waits)
  auto halo_wait = [&](int which) {
    pend[which] = false;
    if (halo_bytes == 0) return;
    const unsigned mbar = smem_u32(&S->mbar_halo[which]);
    if (threadIdx.x == 0) mbar_arrive_expect(mbar, halo_bytes);
    mbar_wait(mbar, use_halo[which] & 1u);
    ++use_halo[which];
  };
  // the constant 1/M halo rows and the warm-start x halo are read once, remotely (slices final since cl.sync)
  if (PRE) {
    const T* ms_lo = cl.map_shared_rank(ms, lo_rank); const T* ms_hi = cl.map_shared_rank(ms, hi_rank);
    for (int j = threadIdx.x; j < NY; j += CL_THREADS) {
      if (has_lo) m_lo[j] = ms_lo[lo_off + j];
      if (has_hi) m_hi[j] = ms_hi[j];
    }
  }
  const T* xs_lo = cl.map_shared_rank(xs, lo_rank); const T* xs_hi = cl.map_shared_rank(xs, hi_rank);

  // ---- stencil product y = A * hat(src) at local element e = (li, j); src selected by MODE -----------------------
  //   MODE 0: src = p                  hat = 1/M * p     (halo rows: pushed p rows)
  //   MODE 1: src = s = r - alpha v    hat = 1/M * s     (halo rows recomputed from the pushed r, v rows)
  //   MODE 2: src = x (warm-start residual, karman_taichi.py:360-367: no preconditioner; halo read remotely, once)
  auto val = [&](int mode, const T* a0, const T* a1, const T* mm, int idx, T alpha) -> T {
    T v_ = (mode == 1) ? (a0[idx] - alpha * a1[idx]) : a0[idx];
    if (PRE && mode != 2) v_ = mm[idx] * v_;
    return v_;
  };
  auto spmv_elem = [&](int mode, int li, int j, int e, T alpha, T& centre_raw) -> T {
    const T* a0 = (mode == 0) ? ps : (mode == 1) ? rs : xs;
    const T* a0_lo = (mode == 0) ? h_lo[HALO_P] : (mode == 1) ? h_lo[HALO_R] : xs_lo + lo_off;
    const T* a0_hi = (mode == 0) ? h_hi[HALO_P] : (mode == 1) ? h_hi[HALO_R] : xs_hi;
    centre_raw = (mode == 1) ? (rs[e] - alpha * vs[e]) : a0[e];
    const T xc = (PRE && mode != 2) ? ms[e] * centre_raw : centre_raw;
    T xw = T(0), xe = T(0), xn = T(0), xso = T(0);
    if (li > 0) xw = val(mode, a0, vs, ms, e - NY, alpha);
    else if (has_lo) xw = val(mode, a0_lo, h_lo[HALO_V], m_lo, j, alpha);
    if (li < nrows - 1) xe = val(mode, a0, vs, ms, e + NY, alpha);
    else if (has_hi) xe = val(mode, a0_hi, h_hi[HALO_V], m_hi, j, alpha);
    if (j < NY - 1) xn = val(mode, a0, vs, ms, e + 1, alpha);
    if (j > 0) xso = val(mode, a0, vs, ms, e - 1, alpha);
    T acc;
    if (PRE || mode == 2) {   // the monolith's product (karman_taichi.py:362-367,401-406): flat offsets i-ny, i-1, i, i+1, i+ny from 0.0
      acc = T(0) + c1[e] * xw;
      acc = acc + c4[e] * xso;
      acc = acc + c0[e] * xc;
      acc = acc + c3[e] * xn;
      acc = acc + c2[e] * xe;
    } else {                    // cgsolver.py:59-60, left to right
      acc = c0[e] * xc;
      acc = acc + c1[e] * xw;
      acc = acc + c2[e] * xe;
      acc = acc + c3[e] * xn;
      acc = acc + c4[e] * xso;
    }
    return acc;
  };

  // ---- K0: r = b - A x0 (warm) or r = b ; r~ = r ; (r,r) -----------------------------------------------------------
  CLP_DECL
  double d0 = 0.0, d1 = 0.0;
  CL_FOR(li, j, e) {
    T rr;
    if (A.warm) { T dummy; rr = bs[e] - spmv_elem(2, li, j, e, T(0), dummy); }
    else rr = bs[e];
    rs[e] = rr;
    if (own_rt) qs[e] = rr;
    if (A.is_cg) { ps[e] = rr; push(HALO_P, li, j, rr); }  // cgsolver.py:90 p = r + 0 * p
    else push(HALO_R, li, j, rr);                          // r halo: read by P3 of the first iteration
    d0 += (double)rr * (double)rr;
  }
  pend[A.is_cg ? HALO_P : HALO_R] = true;
  // group 2 for this one: the loop's first sum (group 0) may be sent by a fast CTA before a slow one has read the slots
  // of the sum before it, unless another sum lies in between -- group 2 is next used two sums later (P4) or never (CG)
  cluster_sum(S, 2, use_sum[2]++, d0, d1, A.nranks, rank CLP_PASS);
  scalar_algebra(A.is_cg ? FIN_CG_INIT : FIN_BICG_INIT, &sc, d0, d0, hist, A.hist_cap);
  if (A.is_cg) { halo_wait(HALO_P); __syncthreads(); }

  // ---- iterations ----------------------------------------------------------------------------------------------
  CLP(15);
  while (!sc.done) {
    if (!A.is_cg) {
      // P1.  The neighbours' p halo rows were last read in their P2 of the previous iteration, which they finished
      // before contributing to that iteration's sums -- and this CTA is past those sums: pushing is safe.
      const bool first = sc.iter == 0;
      const T beta = T(sc.beta), omega_prev = T(sc.omega);
      CL_FOR(li, j, e) {
        const T pn = first ? rs[e] : rs[e] + beta * (ps[e] - omega_prev * vs[e]);
        ps[e] = pn;
        push(HALO_P, li, j, pn);
      }
      pend[HALO_P] = true;
      CLP(0);
      halo_wait(HALO_P);
      __syncthreads();                                     // own p rows visible to the whole CTA
      CLP(1);
      // P2
      d0 = 0.0; d1 = 0.0;
      CL_FOR(li, j, e) {
        T raw;
        const T y = spmv_elem(0, li, j, e, T(0), raw);
        d0 += (double)rt[e] * (double)y;
        vs[e] = y;      // v of the previous iteration is needed by nobody any more
        push(HALO_V, li, j, y);                            // neighbours read their v halo in P3 only (two sums ago)
      }
      pend[HALO_V] = true;
      CLP(2);
      cluster_sum(S, 0, use_sum[0]++, d0, d1, A.nranks, rank CLP_PASS);
      scalar_algebra(FIN_BICG_ALPHA, &sc, d0, d1, nullptr, 0);
      CLP(7);
      if (sc.done) break;
      // P3 (own v rows: visible since the __syncthreads inside cluster_sum; halo rows of v and r: wait for them)
      halo_wait(HALO_V);
      halo_wait(HALO_R);
      const T alpha = T(sc.alpha);
      d0 = 0.0; d1 = 0.0;
      CL_FOR(li, j, e) {
        T sraw;
        const T y = spmv_elem(1, li, j, e, alpha, sraw);
        ss[e] = sraw; ts[e] = y;
        d0 += (double)y * (double)sraw;
        d1 += (double)y * (double)y;
      }
      CLP(8);
      cluster_sum(S, 1, use_sum[1]++, d0, d1, A.nranks, rank CLP_PASS);
      scalar_algebra(FIN_BICG_OMEGA, &sc, d0, d1, nullptr, 0);
      CLP(7);
      if (sc.done) break;
      // P4 (the neighbours read their r halo in P3, which they finished before contributing to the sum above)
      const T omega = T(sc.omega);
      d0 = 0.0; d1 = 0.0;
      CL_FOR(li, j, e) {
        const T ph = PRE ? ms[e] * ps[e] : ps[e];
        const T sh = PRE ? ms[e] * ss[e] : ss[e];
        xs[e] = xs[e] + (alpha * ph + omega * sh);
        const T rn = ss[e] - omega * ts[e];
        rs[e] = rn;
        push(HALO_R, li, j, rn);
        d0 += (double)rn * (double)rn;
        d1 += (double)rn * (double)rt[e];
      }
      pend[HALO_R] = true;
      CLP(9);
      cluster_sum(S, 2, use_sum[2]++, d0, d1, A.nranks, rank CLP_PASS);
      scalar_algebra(FIN_BICG_END, &sc, d0, d1, hist, A.hist_cap);
      CLP(7);
    } else {
      // C1: Ap = A p ; (p, Ap)
      d0 = 0.0; d1 = 0.0;
      CL_FOR(li, j, e) {
        T raw;
        const T y = spmv_elem(0, li, j, e, T(0), raw);
        vs[e] = y;
        d0 += (double)raw * (double)y;
      }
      cluster_sum(S, 0, use_sum[0]++, d0, d1, A.nranks, rank CLP_PASS);
      scalar_algebra(FIN_CG_ALPHA, &sc, d0, d1, nullptr, 0);
      if (sc.done) break;
      // C2: x += alpha p ; r -= alpha Ap ; (r,r)
      const T alpha = T(sc.alpha);
      d0 = 0.0; d1 = 0.0;
      CL_FOR(li, j, e) {
        xs[e] = xs[e] + alpha * ps[e];
        const T rn = rs[e] - alpha * vs[e];
        rs[e] = rn;
        d0 += (double)rn * (double)rn;
      }
      cluster_sum(S, 1, use_sum[1]++, d0, d1, A.nranks, rank CLP_PASS);
      scalar_algebra(FIN_CG_END, &sc, d0, d1, hist, A.hist_cap);
      if (sc.done) break;
      // C3: p = r + beta p (the neighbours finished reading their p halo in C1, two sums ago)
      const T beta = T(sc.beta);
      CL_FOR(li, j, e) {
        const T pn = rs[e] + beta * ps[e];
        ps[e] = pn;
        push(HALO_P, li, j, pn);
      }
      pend[HALO_P] = true;
      halo_wait(HALO_P);
      __syncthreads();
    }
  }
  // Every CTA leaves the loop at the same point (identical scalars).  Rows pushed by a neighbour may still be in flight
  // towards this CTA (r of the last P4, v of an aborted iteration): absorb them before this CTA exits and frees the
  // shared memory they target; the barrier keeps every slice alive until all CTAs are past their last remote access.
#pragma unroll
  for (int k = 0; k < 3; ++k) if (pend[k]) halo_wait(k);
  cl.sync();
  CL_FOR(li, j, e) {
    const size_t g = (size_t)(r0 + li) * A.pitch + j;
    A.x[g] = xs[e];
    if (A.r_out) A.r_out[g] = rs[e];
  }
  if (rank == 0 && threadIdx.x == 0) *A.sc = sc;
  CLP_DUMP;
#undef CL_FOR
}

static int cluster_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

constexpr int CL_EPT_REG = 2;       // register-resident kernel, 512 threads: up to this many elements per thread
constexpr int CL_EPT_REG256 = 5;    // register-resident kernel, 256 threads (255 registers each): up to this many (6 spills: slower)

struct ClFit { int C, rows_per, ept, nt; bool reg; size_t smem; };

// Can this geometry be solved inside one cluster?  C = 0: no.
template <typename T>
static ClFit cluster_fit(const kt_handle_s* h, bool pre, bool warm, int limit_c) {
  ClFit f = {0, 0, 0, CL_THREADS, false, 0};
  static int enabled = cluster_env("KT_CLUSTER", 1);
  if (!enabled || !h->allow_cluster || h->row_lo != 0 || h->row_hi != h->NX) return f;
  const size_t ctrl = ((sizeof(ClShared) + 15) / 16) * 16;
  const size_t max_smem = 227 * 1024;
  // 16 CTAs (non-portable size, one GPC) whenever every CTA still gets >= 2 rows; else the portable 8.
  static int max_c = cluster_env("KT_CLUSTER_MAX", 16);
  static int reg_max = cluster_env("KT_CLUSTER_REG", CL_EPT_REG);     // 0: always the shared-memory kernel
  for (int C : {16, 8}) {
    if (C > max_c || C > limit_c || (C == 16 && h->NX < 32)) continue;
    const int rows_per = (h->NX + C - 1) / C;
    const long long cap = (long long)rows_per * h->NY;
    int ept = (int)((cap + CL_THREADS - 1) / CL_THREADS), nt = CL_THREADS;
    bool reg = ept <= (reg_max < CL_EPT_REG ? reg_max : CL_EPT_REG);
    if (!reg && reg_max >= CL_EPT_REG && (cap + 255) / 256 <= CL_EPT_REG256) { reg = true; nt = 256; ept = (int)((cap + 255) / 256); }
    const int planes = reg ? 3 + (pre ? 1 : 0) : 12 + (pre ? 1 : 0) + (warm ? 1 : 0);
    const size_t need = ctrl + ((size_t)planes * cap + 8 * (size_t)h->NY) * sizeof(T);      // + 8 private halo rows
    if (need <= max_smem) { f.C = C; f.rows_per = rows_per; f.ept = ept; f.nt = nt; f.reg = reg; f.smem = need; return f; }
  }
  return f;
}

template <typename K, typename T>
static int cluster_launch_k(kt_handle_s* h, K kern, unsigned long long& configured, const ClArgs<T>& A0, int C, size_t smem, int nt = CL_THREADS) {
  const unsigned long long dev_bit = 1ull << (h->device & 63);   // function attributes are per device
  if (!(configured & dev_bit)) {
    KT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    KT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured |= dev_bit;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C); cfg.blockDim = dim3(nt); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  KT_CUDA(cudaLaunchKernelEx(&cfg, kern, A0));
  h->launches++;
  return 0;
}

template <typename T, bool PRE>
static int cluster_launch(kt_handle_s* h, const ClArgs<T>& A, const ClFit& f) {
  static unsigned long long conf[7] = {0, 0, 0, 0, 0, 0, 0};
  if (f.reg && f.nt == CL_THREADS && f.ept == 1) return cluster_launch_k(h, k_cluster_solve<T, PRE, 1, true, CL_THREADS>, conf[0], A, f.C, f.smem);
  if (f.reg && f.nt == CL_THREADS && f.ept == 2) return cluster_launch_k(h, k_cluster_solve<T, PRE, 2, true, CL_THREADS>, conf[1], A, f.C, f.smem);
  if (f.reg && f.nt == 256) {
    switch (f.ept) {
      case 1: case 2: case 3: return cluster_launch_k(h, k_cluster_solve<T, PRE, 3, true, 256>, conf[3], A, f.C, f.smem, 256);
      case 4: return cluster_launch_k(h, k_cluster_solve<T, PRE, 4, true, 256>, conf[4], A, f.C, f.smem, 256);
      case 5: return cluster_launch_k(h, k_cluster_solve<T, PRE, 5, true, 256>, conf[5], A, f.C, f.smem, 256);
      default: break;
    }
  }
  return cluster_launch_k(h, k_cluster_solve_smem<T, PRE>, conf[2], A, f.C, f.smem);
}

static int g_cluster_sched_limit = 16;   // drops to 8 for the process if a 16-CTA cluster cannot be scheduled here

// What cluster_try_solve would launch for this geometry (kt_resident_plan): 0 = does not fit.
template <typename T>
int cluster_plan(const kt_handle_s* h, bool pre, bool warm, bool is_cg, long long* out) {
  if (is_cg && (pre || warm)) return 0;
  const ClFit f = cluster_fit<T>(h, pre, warm, g_cluster_sched_limit);
  if (f.C == 0) return 0;
  out[0] = f.reg ? 1 : 2; out[1] = f.C; out[2] = f.rows_per; out[3] = f.reg ? f.nt : CL_THREADS; out[4] = f.ept;
  return 1;
}
template int cluster_plan<double>(const kt_handle_s*, bool, bool, bool, long long*);
template int cluster_plan<float>(const kt_handle_s*, bool, bool, bool, long long*);

// Entry used by kt_bicgstab_solve / kt_cg_solve: returns 1 if the solve was done here, 0 if the geometry does not fit
// (caller falls through to the streaming kernels), < 0 on error.
template <typename T>
int cluster_try_solve(kt_handle_s* h, const void* coef, const void* b, void* x, const void* Minv, void* r_out,
                      double* history, long long hist_cap, double eps, long long max_iter, int flags, bool is_cg) {
  const bool pre = Minv != nullptr, warm = (flags & KT_FLAG_WARM_START) != 0;
  if (is_cg && (pre || warm)) return 0;   // CGSolver (cgsolver.py) has neither a preconditioner nor a warm start
  int& sched_limit = g_cluster_sched_limit;
  ClFit f = cluster_fit<T>(h, pre, warm, sched_limit);
  if (f.C == 0) return 0;
  ClArgs<T> A;
  A.coef = (const T*)coef; A.plane = (size_t)h->NX * h->pitch; A.b = (const T*)b; A.x = (T*)x; A.minv = (const T*)Minv;
  A.r_out = (T*)r_out; A.NX = h->NX; A.NY = h->NY; A.pitch = h->pitch; A.rows_per = f.rows_per; A.nranks = f.C;
  A.sc = h->sc; A.history = history; A.hist_cap = hist_cap; A.eps = eps; A.max_iter = max_iter; A.flags = flags;
  A.warm = warm ? 1 : 0; A.is_cg = is_cg ? 1 : 0;
  int rc = pre ? cluster_launch<T, true>(h, A, f) : cluster_launch<T, false>(h, A, f);
  if (rc && f.C == 16) {
    // a 16-CTA cluster could not be scheduled on this device: retry with the portable size
    cudaGetLastError();
    sched_limit = 8;
    f = cluster_fit<T>(h, pre, warm, 8);
    if (f.C == 0) return 0;
    A.rows_per = f.rows_per; A.nranks = f.C;
    rc = pre ? cluster_launch<T, true>(h, A, f) : cluster_launch<T, false>(h, A, f);
  }
  if (rc) {
    // no cluster of a usable size can be scheduled (e.g. MIG slice): fall back to the streaming path
    cudaGetLastError();
    return 0;
  }
  KT_CUDA(cudaMemcpyAsync(h->sc_host, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
  KT_CUDA(cudaStreamSynchronize(h->stream));
  return 1;
}

#ifdef KT_CLUSTER_PROF
extern "C" int kt_debug_cluster_prof(long long* out16, int reset) {
  if (out16) cudaMemcpyFromSymbol(out16, g_clprof, sizeof(long long) * 16);
  if (reset) { long long z[16] = {0}; cudaMemcpyToSymbol(g_clprof, z, sizeof(z)); }
  return 0;
}
#endif

template int cluster_try_solve<double>(kt_handle_s*, const void*, const void*, void*, const void*, void*, double*, long long, double,
                                       long long, int, bool);
template int cluster_try_solve<float>(kt_handle_s*, const void*, const void*, void*, const void*, void*, double*, long long, double,
                                      long long, int, bool);

}  // namespace kt

// every kernel of this translation unit, for kt_preload_kernels (kt_handle.h)
namespace kt {
int preload_kt_cluster() {
  static const void* const fns[] = {
#define KT_PRELOAD_TU_kt_cluster
#include "kt_preload_list.inc"
#undef KT_PRELOAD_TU_kt_cluster
      nullptr};
  return preload_list(fns);
}
}  // namespace kt
