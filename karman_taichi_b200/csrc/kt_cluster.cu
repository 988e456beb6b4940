// kt_cluster.cu -- small-grid BiCGSTAB / CG that lives entirely in ONE thread-block cluster (sm_100a).
//
// The reference's own grids are small (cavity 30^2..128^2, Karman 320x64): ~2e4 unknowns.  At that size an iteration
// of the streaming path is four kernel launches of a few microseconds each and the solve is launch-latency bound
// (26-37 us per Krylov iteration measured, profiles/r01_simple_rates.txt).  Here the whole solve is one launch: a
// cluster of up to 16 CTAs holds every plane (5 coefficients, b, x, r, p, v, s, t [, 1/M, r~]) in its shared memory
// (16 x 227 KB), each CTA owning a contiguous block of rows; halo rows of the stencil operand are read straight out of
// the neighbour CTA's shared memory (DSMEM), inner products are block-reduced, exchanged through DSMEM and summed by
// every CTA in rank order (deterministic, identical on all CTAs), and the phases of an iteration are separated by
// hardware cluster barriers instead of kernel boundaries.  Same recurrence, same operation order per element, same
// scalar algebra (kt::scalar_algebra) as the streaming kernels.
//
// BiCGSTAB per iteration (bicgsolver.py:94-135 / karman_taichi.py:383-456): 4 cluster barriers
//   P1 p = r + beta (p - omega v)                                 | barrier (neighbours read p)
//   P2 v = A p^ ; (r~, v)                                         | barrier -> alpha
//   P3 t = A s^ with s = r - alpha v (halo rows recomputed from the neighbour's r, v) ; (t,s), (t,t) | barrier -> omega
//   P4 x += alpha p^ + omega s^ ; r = s - omega t ; (r,r), (r,r~) | barrier -> test, rho, beta
// CG per iteration (cgsolver.py:92-114): 3 barriers (Ap,(p,Ap) | x,r,(r,r) | p = r + beta p).
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

struct ClShared {           // per-CTA control block at the start of dynamic shared memory
  double part[3][2][16];    // [dot group][value][source rank]: every CTA PUSHES its block sums into all CTAs' slots
  double red[2][CL_THREADS / 32];
};

// 16-lane fixed-shape tree: every lane of the warp ends up with v[0] + ... + v[15] (lanes >= 16 mirror lane - 16).
__device__ __forceinline__ double tree16(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// fixed-order cluster-wide sum of up to two values.  Warp shuffle tree -> one value per warp in shared memory ->
// warp 0 adds the (16) warp sums with a fixed tree and its lanes 0..C-1 store the block sum into slot [rank] of EVERY
// CTA (remote DSMEM stores, fire and forget) -> cluster barrier (release / acquire) -> every warp adds the C slots of
// its own CTA with the same fixed tree.  Identical bits on every CTA, warp and thread; no serial chain, no remote load
// after the barrier.
template <int NDOT>
__device__ __forceinline__ void cluster_sum(cg::cluster_group& cl, ClShared* S, int group, double& d0, double& d1, int nranks,
                                            int rank) {
  static_assert(CL_THREADS / 32 == 16, "tree16 assumes 16 warps per CTA");
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  d0 = warp_sum(d0);
  if (NDOT > 1) d1 = warp_sum(d1);
  if (lane == 0) { S->red[0][wid] = d0; if (NDOT > 1) S->red[1][wid] = d1; }
  __syncthreads();
  if (wid == 0) {
    double s0 = tree16(S->red[0][lane & 15]);
    double s1 = (NDOT > 1) ? tree16(S->red[1][lane & 15]) : 0.0;
    if (lane < nranks) {
      ClShared* R = cl.map_shared_rank(S, lane);
      R->part[group][0][rank] = s0;
      if (NDOT > 1) R->part[group][1][rank] = s1;
    }
  }
  cl.sync();
  const int r = lane & 15;
  d0 = tree16(r < nranks ? S->part[group][0][r] : 0.0);
  d1 = (NDOT > 1) ? tree16(r < nranks ? S->part[group][1][r] : 0.0) : 0.0;
}

template <typename T, bool PRE>
__global__ void __launch_bounds__(CL_THREADS, 1) k_cluster_solve(const ClArgs<T> A) {
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

  // thread -> (row group, column) map fixed for the whole solve: no integer division inside the phases.
  const int TC = min(NY, CL_THREADS), TR = CL_THREADS / TC;
  const int tx = threadIdx.x % TC;
  const int ty = (threadIdx.x / TC < TR) ? threadIdx.x / TC : nrows;     // surplus threads own no rows
#define CL_FOR(li, j, e) \
  for (int j = tx; j < NY; j += TC) for (int li = ty, e = ty * NY + j; li < nrows; li += TR, e += TR * NY)

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
  cl.sync();

  // neighbours' slices (row r0-1 is the last row of rank-1, row r1 the first row of rank+1: every lower rank is full)
  const bool has_lo = r0 > 0 && nrows > 0, has_hi = r1 < A.NX && nrows > 0;
  const int lo_rank = rank > 0 ? rank - 1 : 0, hi_rank = rank < A.nranks - 1 ? rank + 1 : rank;
  const int lo_off = (A.rows_per - 1) * NY;               // offset of the neighbour's last row
  // remote views of the operand planes (DSMEM), resolved once
  const T* ps_lo = cl.map_shared_rank(ps, lo_rank); const T* ps_hi = cl.map_shared_rank(ps, hi_rank);
  const T* rs_lo = cl.map_shared_rank(rs, lo_rank); const T* rs_hi = cl.map_shared_rank(rs, hi_rank);
  const T* vs_lo = cl.map_shared_rank(vs, lo_rank); const T* vs_hi = cl.map_shared_rank(vs, hi_rank);
  const T* xs_lo = cl.map_shared_rank(xs, lo_rank); const T* xs_hi = cl.map_shared_rank(xs, hi_rank);
  const T* ms_lo = cl.map_shared_rank(ms, lo_rank); const T* ms_hi = cl.map_shared_rank(ms, hi_rank);

  // ---- stencil product y = A * hat(src) at local element e = (li, j); src selected by MODE -----------------------
  //   MODE 0: src = p                  hat = 1/M * p
  //   MODE 1: src = s = r - alpha v    hat = 1/M * s     (halo rows recomputed from the neighbour's r, v)
  //   MODE 2: src = x (warm-start residual, karman_taichi.py:360-367: no preconditioner)
  auto val = [&](int mode, const T* a0, const T* a1, const T* mm, int idx, T alpha) -> T {
    T v_ = (mode == 1) ? (a0[idx] - alpha * a1[idx]) : a0[idx];
    if (PRE && mode != 2) v_ = mm[idx] * v_;
    return v_;
  };
  auto spmv_elem = [&](int mode, int li, int j, int e, T alpha, T& centre_raw) -> T {
    const T* a0 = (mode == 0) ? ps : (mode == 1) ? rs : xs;
    const T* a0_lo = (mode == 0) ? ps_lo : (mode == 1) ? rs_lo : xs_lo;
    const T* a0_hi = (mode == 0) ? ps_hi : (mode == 1) ? rs_hi : xs_hi;
    centre_raw = (mode == 1) ? (rs[e] - alpha * vs[e]) : a0[e];
    const T xc = (PRE && mode != 2) ? ms[e] * centre_raw : centre_raw;
    T xw = T(0), xe = T(0), xn = T(0), xso = T(0);
    if (li > 0) xw = val(mode, a0, vs, ms, e - NY, alpha);
    else if (has_lo) xw = val(mode, a0_lo, vs_lo, ms_lo, lo_off + j, alpha);
    if (li < nrows - 1) xe = val(mode, a0, vs, ms, e + NY, alpha);
    else if (has_hi) xe = val(mode, a0_hi, vs_hi, ms_hi, j, alpha);
    if (j < NY - 1) xn = val(mode, a0, vs, ms, e + 1, alpha);
    if (j > 0) xso = val(mode, a0, vs, ms, e - 1, alpha);
    // cgsolver.py:59-60, left to right
    T acc = c0[e] * xc;
    acc = acc + c1[e] * xw;
    acc = acc + c2[e] * xe;
    acc = acc + c3[e] * xn;
    acc = acc + c4[e] * xso;
    return acc;
  };

  // ---- K0: r = b - A x0 (warm) or r = b ; r~ = r ; (r,r) -----------------------------------------------------------
  double d0 = 0.0, d1 = 0.0;
  CL_FOR(li, j, e) {
    T rr;
    if (A.warm) { T dummy; rr = bs[e] - spmv_elem(2, li, j, e, T(0), dummy); }
    else rr = bs[e];
    rs[e] = rr;
    if (own_rt) qs[e] = rr;
    if (A.is_cg) ps[e] = rr;                               // cgsolver.py:90 p = r + 0 * p
    d0 += (double)rr * (double)rr;
  }
  cluster_sum<1>(cl, S, 0, d0, d1, A.nranks, rank);
  scalar_algebra(A.is_cg ? FIN_CG_INIT : FIN_BICG_INIT, &sc, d0, d0, hist, A.hist_cap);

  // ---- iterations ----------------------------------------------------------------------------------------------
  while (!sc.done) {
    if (!A.is_cg) {
      // P1
      const bool first = sc.iter == 0;
      const T beta = T(sc.beta), omega_prev = T(sc.omega);
      CL_FOR(li, j, e) ps[e] = first ? rs[e] : rs[e] + beta * (ps[e] - omega_prev * vs[e]);
      cl.sync();
      // P2
      d0 = 0.0; d1 = 0.0;
      CL_FOR(li, j, e) {
        T raw;
        const T y = spmv_elem(0, li, j, e, T(0), raw);
        d0 += (double)rt[e] * (double)y;
        vs[e] = y;      // v of the previous iteration is needed by nobody any more (P1 is behind the barrier)
      }
      cluster_sum<1>(cl, S, 0, d0, d1, A.nranks, rank);
      scalar_algebra(FIN_BICG_ALPHA, &sc, d0, d1, nullptr, 0);
      if (sc.done) break;
      // P3 (v of the neighbours is final since the barrier inside cluster_sum)
      const T alpha = T(sc.alpha);
      d0 = 0.0; d1 = 0.0;
      CL_FOR(li, j, e) {
        T sraw;
        const T y = spmv_elem(1, li, j, e, alpha, sraw);
        ss[e] = sraw; ts[e] = y;
        d0 += (double)y * (double)sraw;
        d1 += (double)y * (double)y;
      }
      cluster_sum<2>(cl, S, 1, d0, d1, A.nranks, rank);
      scalar_algebra(FIN_BICG_OMEGA, &sc, d0, d1, nullptr, 0);
      if (sc.done) break;
      // P4
      const T omega = T(sc.omega);
      d0 = 0.0; d1 = 0.0;
      CL_FOR(li, j, e) {
        const T ph = PRE ? ms[e] * ps[e] : ps[e];
        const T sh = PRE ? ms[e] * ss[e] : ss[e];
        xs[e] = xs[e] + (alpha * ph + omega * sh);
        const T rn = ss[e] - omega * ts[e];
        rs[e] = rn;
        d0 += (double)rn * (double)rn;
        d1 += (double)rn * (double)rt[e];
      }
      cluster_sum<2>(cl, S, 2, d0, d1, A.nranks, rank);
      scalar_algebra(FIN_BICG_END, &sc, d0, d1, hist, A.hist_cap);
    } else {
      // C1: Ap = A p ; (p, Ap)
      d0 = 0.0; d1 = 0.0;
      CL_FOR(li, j, e) {
        T raw;
        const T y = spmv_elem(0, li, j, e, T(0), raw);
        vs[e] = y;
        d0 += (double)raw * (double)y;
      }
      cluster_sum<1>(cl, S, 0, d0, d1, A.nranks, rank);
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
      cluster_sum<1>(cl, S, 1, d0, d1, A.nranks, rank);
      scalar_algebra(FIN_CG_END, &sc, d0, d1, hist, A.hist_cap);
      if (sc.done) break;
      // C3: p = r + beta p
      const T beta = T(sc.beta);
      CL_FOR(li, j, e) ps[e] = rs[e] + beta * ps[e];
      cl.sync();
    }
  }
  // every CTA leaves the loop at the same iteration (identical scalars); make sure nobody still reads our slices
  cl.sync();
  CL_FOR(li, j, e) {
    const size_t g = (size_t)(r0 + li) * A.pitch + j;
    A.x[g] = xs[e];
    if (A.r_out) A.r_out[g] = rs[e];
  }
  if (rank == 0 && threadIdx.x == 0) *A.sc = sc;
#undef CL_FOR
}

static int cluster_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// Can this geometry be solved inside one cluster?  Returns the cluster size (0 = no) and the shared memory per CTA.
template <typename T, bool PRE>
static int cluster_fit(const kt_handle_s* h, bool warm, size_t* smem_out, int* rows_per_out, int limit_c) {
  static int enabled = cluster_env("KT_CLUSTER", 1);
  if (!enabled || !h->allow_cluster || h->row_lo != 0 || h->row_hi != h->NX) return 0;
  const int planes = 12 + (PRE ? 1 : 0) + (warm ? 1 : 0);
  const size_t ctrl = ((sizeof(ClShared) + 15) / 16) * 16;
  const size_t max_smem = 227 * 1024;
  // 16 CTAs (non-portable size, one GPC) whenever every CTA still gets >= 2 rows: the phases are instruction-issue
  // bound, so twice the SMs is twice the speed (measured 9.5 -> ~5 us / BiCGSTAB iteration at 130^2); else 8.
  static int max_c = cluster_env("KT_CLUSTER_MAX", 16);
  for (int C : {16, 8}) {
    if (C > max_c || C > limit_c || (C == 16 && h->NX < 32)) continue;
    const int rows_per = (h->NX + C - 1) / C;
    const size_t need = ctrl + (size_t)planes * rows_per * h->NY * sizeof(T);
    if (need <= max_smem) { *smem_out = need; *rows_per_out = rows_per; return C; }
  }
  return 0;
}

template <typename T, bool PRE>
static int cluster_launch(kt_handle_s* h, const ClArgs<T>& A0, int C, size_t smem) {
  auto kern = k_cluster_solve<T, PRE>;
  static bool configured = false;
  if (!configured) {
    KT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    KT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  KT_CUDA(cudaLaunchKernelEx(&cfg, kern, A0));
  h->launches++;
  return 0;
}

// Entry used by kt_bicgstab_solve / kt_cg_solve: returns 1 if the solve was done here, 0 if the geometry does not fit
// (caller falls through to the streaming kernels), < 0 on error.
template <typename T>
int cluster_try_solve(kt_handle_s* h, const void* coef, const void* b, void* x, const void* Minv, void* r_out,
                      double* history, long long hist_cap, double eps, long long max_iter, int flags, bool is_cg) {
  const bool pre = Minv != nullptr, warm = (flags & KT_FLAG_WARM_START) != 0;
  static int sched_limit = 16;          // drops to 8 for the process if a 16-CTA cluster cannot be scheduled here
  size_t smem = 0; int rows_per = 0;
  int C = pre ? cluster_fit<T, true>(h, warm, &smem, &rows_per, sched_limit) : cluster_fit<T, false>(h, warm, &smem, &rows_per, sched_limit);
  if (C == 0) return 0;
  ClArgs<T> A;
  A.coef = (const T*)coef; A.plane = (size_t)h->NX * h->pitch; A.b = (const T*)b; A.x = (T*)x; A.minv = (const T*)Minv;
  A.r_out = (T*)r_out; A.NX = h->NX; A.NY = h->NY; A.pitch = h->pitch; A.rows_per = rows_per; A.nranks = C;
  A.sc = h->sc; A.history = history; A.hist_cap = hist_cap; A.eps = eps; A.max_iter = max_iter; A.flags = flags;
  A.warm = warm ? 1 : 0; A.is_cg = is_cg ? 1 : 0;
  int rc = pre ? cluster_launch<T, true>(h, A, C, smem) : cluster_launch<T, false>(h, A, C, smem);
  if (rc && C == 16) {
    // a 16-CTA cluster could not be scheduled on this device: retry with the portable size
    cudaGetLastError();
    sched_limit = 8;
    C = pre ? cluster_fit<T, true>(h, warm, &smem, &rows_per, 8) : cluster_fit<T, false>(h, warm, &smem, &rows_per, 8);
    if (C == 0) return 0;
    A.rows_per = rows_per; A.nranks = C;
    rc = pre ? cluster_launch<T, true>(h, A, C, smem) : cluster_launch<T, false>(h, A, C, smem);
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

template int cluster_try_solve<double>(kt_handle_s*, const void*, const void*, void*, const void*, void*, double*, long long, double,
                                       long long, int, bool);
template int cluster_try_solve<float>(kt_handle_s*, const void*, const void*, void*, const void*, void*, double*, long long, double,
                                      long long, int, bool);

}  // namespace kt
