// kt_coop.cu -- mid-size BiCGSTAB / CG (too big for one cluster, up to ~4.5e5 unknowns) as ONE persistent cooperative
// kernel with one CTA per SM (sm_100a: 148 CTAs x 512 threads).
//
// Between the cluster-resident solver (kt_cluster.cu, <= ~35 k unknowns) and the HBM-bound streaming kernels lies a
// range -- cavity 200^2..550^2, the 640x128 channel karman_taichi.py is committed with -- where the streaming path costs
// ~7.5 us of GPU time per kernel even when replayed from a CUDA graph (4 kernels per BiCGSTAB iteration,
// profiles/r01_simple_rates.md).  Here the whole solve is one launch: every CTA owns a block of rows, every thread owns
// the same <= 6 elements for the whole solve and keeps what only it touches in registers (coefficients when <= 2 elements,
// r~, x, r, v, p, t, 1/M); shared memory holds the p, v, r slices its neighbour elements read; the boundary rows of v and r
// go through the caller's global planes (L2); the inner products go through per-CTA slots that every CTA polls, which is
// at the same time the grid barrier (see grid_sum).  The launch is cooperative only to guarantee co-residency.  Same recurrence, same operation order per element, same scalar algebra
// (kt::scalar_algebra, run redundantly by every thread on the identical sums) as the other two paths.
//
// BiCGSTAB per iteration (bicgsolver.py:94-135 / karman_taichi.py:383-456): 3 grid-wide sums (= barriers)
//   P1 p = r + beta (p - omega v) -- also for the two halo rows, from the neighbours' r rows (global, final since the last
//      barrier) and private copies of their previous p and v rows: same operands, same bits as the owner computes
//   P2 v = A p^ ; boundary rows of v -> global ; (r~, v)                      | barrier -> alpha
//   P3 t = A s^ with s = r - alpha v (halo rows from the neighbours' r, v rows) ; (t,s), (t,t) | barrier -> omega
//   P4 x += alpha p^ + omega s^ ; r = s - omega t ; boundary rows of r -> global ; (r,r), (r,r~) | barrier -> test, rho, beta
// CG per iteration (cgsolver.py:92-114): 2 grid-wide sums.
#include "kt_handle.h"

namespace kt {

template <typename T> struct CoArgs {
  const T* coef; size_t plane;     // global SoA pitched planes
  const T* b; T* x; const T* minv;
  T* gr; T* gv;                    // caller's r and Ap planes: boundary-row exchange; gr receives the final residual
  double* partials;                // [3 groups][G] pairs (d0, d1); all bits set = empty (host memset before the launch)
  int NX, NY, pitch, rows_per, G;
  Scalars* sc; double* history; long long hist_cap;
  double eps; long long max_iter; int flags; int warm; int is_cg;
};

constexpr int CO_THREADS = 512;
constexpr int CO_SLOTS = 8;           // slots per polling lane: up to 256 CTAs
enum { CF_VALID = 1, CF_ROW0 = 2, CF_ROWN = 4, CF_J0 = 8, CF_JN = 16 };

__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }

// xor butterfly: every lane ends up with the same bits (a + b == b + a at every level)
__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, bool PRE, int EPT, bool REGC>
__global__ void __launch_bounds__(CO_THREADS, 1) k_coop_solve(const CoArgs<T> A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[2][CO_THREADS / 32];
  __shared__ double bc[2];
  const int rank = (int)blockIdx.x, G = A.G;
  const int NY = A.NY;
  const int r0 = min(rank * A.rows_per, A.NX), r1 = min(r0 + A.rows_per, A.NX);
  const int nrows = r1 - r0, n_el = nrows * NY;
  const int cap = A.rows_per * NY;                        // plane slice capacity (same on every CTA)
  T* ps = reinterpret_cast<T*>(smem_raw); T* vs = ps + cap; T* rs = vs + cap;
  T* ms = rs + cap;                                       // 1/M (PRE only)
  T* cs = PRE ? ms + cap : ms;                            // five coefficient planes (!REGC only)
  T* hal = REGC ? cs : cs + 5 * cap;                      // private halo rows: row r0-1 = "lo", row r1 = "hi"
  T* p_lo = hal; T* p_hi = hal + NY; T* v_lo = hal + 2 * NY; T* v_hi = hal + 3 * NY;
  T* r_lo = hal + 4 * NY; T* r_hi = hal + 5 * NY; T* m_lo = hal + 6 * NY; T* m_hi = hal + 7 * NY;
  const bool has_lo = r0 > 0 && nrows > 0, has_hi = r1 < A.NX && nrows > 0;
  const size_t g_lo = (size_t)(r0 - 1) * A.pitch, g_hi = (size_t)r1 * A.pitch;      // global offsets of the halo rows

  // ---- this thread's elements ---------------------------------------------------------------------------------------
  T c[REGC ? EPT : 1][5];
  T rt[EPT], x[EPT], r[EPT], v[EPT], p[EPT], t[EPT], m[EPT];
  unsigned ef[EPT];                 // CF_* flags in the low byte, the column j above it
#define CJ(k) ((int)(ef[k] >> 8))
#pragma unroll
  for (int k = 0; k < EPT; ++k) {
    const int e = threadIdx.x + k * CO_THREADS;
    ef[k] = 0;
    x[k] = r[k] = v[k] = p[k] = t[k] = rt[k] = T(0); m[k] = T(1);
    if (REGC) { for (int q = 0; q < 5; ++q) c[k][q] = T(0); }
    if (e < n_el) {
      const int li = e / NY, j = e - li * NY;
      const size_t g = (size_t)(r0 + li) * A.pitch + j;
      ef[k] = CF_VALID | (li == 0 ? CF_ROW0 : 0) | (li == nrows - 1 ? CF_ROWN : 0) | (j == 0 ? CF_J0 : 0) | (j == NY - 1 ? CF_JN : 0) |
              ((unsigned)j << 8);
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const T cv = A.coef[q * A.plane + g];
        if (REGC) c[k][q] = cv; else cs[q * cap + e] = cv;
      }
      if (A.warm) x[k] = A.x[g];
      if (PRE) { m[k] = A.minv[g]; ms[e] = m[k]; }
    }
  }
  if (PRE) {
    for (int j = threadIdx.x; j < NY; j += CO_THREADS) {
      if (has_lo) m_lo[j] = A.minv[g_lo + j];
      if (has_hi) m_hi[j] = A.minv[g_hi + j];
    }
  }
  auto coef = [&](int k, int q, int e) -> T { return REGC ? c[REGC ? k : 0][q] : cs[q * cap + e]; };
  // global index of a boundary-row element of this CTA (only asked for elements with CF_ROW0 / CF_ROWN)
  auto g_of = [&](unsigned f) -> size_t { return (size_t)(r0 + ((f & CF_ROW0) ? 0 : nrows - 1)) * A.pitch + (f >> 8); };
  auto publish = [&](T* plane, unsigned f, T val) {       // boundary rows a neighbour CTA will read after the next barrier
    if (((f & CF_ROW0) && has_lo) || ((f & CF_ROWN) && has_hi)) plane[g_of(f)] = val;
  };
  auto load_halo = [&](T* lo, T* hi, const T* plane) {
    for (int j = threadIdx.x; j < NY; j += CO_THREADS) {
      if (has_lo) lo[j] = ld_cg(plane + g_lo + j);
      if (has_hi) hi[j] = ld_cg(plane + g_hi + j);
    }
  };

  Scalars sc;
  sc.rho = 0.0; sc.rho_1 = 1.0; sc.alpha = 1.0; sc.beta = 1.0; sc.omega = 1.0;
  sc.rTr0 = sc.rTr = sc.old_rTr = 0.0; sc.dot0 = sc.dot1 = 0.0;
  sc.eps = A.eps; sc.iter = 0; sc.max_iter = A.max_iter; sc.status = 1; sc.done = 0; sc.first = 1; sc.flags = A.flags;
  double* const hist = (rank == 0 && threadIdx.x == 0) ? A.history : nullptr;

  // Fixed-order grid-wide sum of two values that is also the grid barrier.  Warp butterfly -> per-warp values in shared
  // memory -> warp 0 adds them; lane 0 fences (the boundary rows every thread of the CTA published are ordered before by
  // the block barrier) and stores the CTA's pair into slot [group][rank] with ONE 16-byte store; the lanes of warp 0 then
  // poll the G slots of the group until none holds the "empty" pattern (all bits set: a NaN no arithmetic produces), adding
  // them in index order (lane l: slots l, l + 32, ...; then the butterfly) and broadcast the result through shared memory:
  // identical bits on every CTA, warp and thread.  One L2 round trip after the last CTA's store, instead of
  // store / fence / barrier arrive / barrier poll / load (measured: ~9 us per phase with cooperative-groups grid.sync).
  // A CTA cannot pass sum N before every CTA has contributed to it, i.e. finished phase N: a sum is a full barrier.
  // Slot reuse: the sums rotate over THREE slot groups and after passing sum N a CTA re-arms its own slot of sum N-1
  // (everybody read it before contributing to N).  That slot is next polled at sum N+2, which nobody reaches without this
  // CTA's contribution to N+1 -- sent, with a fence, after the re-arm.  (Two alternating groups would not do: a fast CTA
  // could poll a slow CTA's slot of sum N+1 before the slow one has re-armed it.)
  int prev_group = -1, next_group = 0;
  auto slot_of = [&](int group, int cta) -> double* { return A.partials + ((size_t)group * G + cta) * 2; };
  // `published`: boundary rows were written to global memory in this phase (the acquire fence is only needed then).
  auto grid_sum = [&](double& d0, double& d1, bool published) {
    const int group = next_group;
    next_group = (next_group == 2) ? 0 : next_group + 1;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    d0 = warp_allsum(d0); d1 = warp_allsum(d1);
    if (lane == 0) { red[0][wid] = d0; red[1][wid] = d1; }
    __syncthreads();
    if (wid == 0) {
      double s0 = (lane < CO_THREADS / 32) ? red[0][lane] : 0.0, s1 = (lane < CO_THREADS / 32) ? red[1][lane] : 0.0;
      s0 = warp_allsum(s0); s1 = warp_allsum(s1);
      if (lane == 0) {
        __threadfence();      // release: this phase's boundary rows AND the re-arm store of the previous sum come first
        asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(slot_of(group, rank)), "d"(s0), "d"(s1) : "memory");
      }
      // every lane owns up to CO_SLOTS slots (l, l + 32, ...): their loads are issued back to back so that one L2 round
      // trip covers all of them, and only the ones still empty are read again
      double u0[CO_SLOTS], u1[CO_SLOTS];
      unsigned pending = 0;
#pragma unroll
      for (int k = 0; k < CO_SLOTS; ++k) { u0[k] = 0.0; u1[k] = 0.0; if (lane + 32 * k < G) pending |= 1u << k; }
      const long long t0 = clock64();
      while (pending) {
#pragma unroll
        for (int k = 0; k < CO_SLOTS; ++k)
          if (pending & (1u << k))
            asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(u0[k]), "=d"(u1[k]) : "l"(slot_of(group, lane + 32 * k)) : "memory");
#pragma unroll
        for (int k = 0; k < CO_SLOTS; ++k)
          if ((pending & (1u << k)) && __double_as_longlong(u0[k]) != -1ll && __double_as_longlong(u1[k]) != -1ll) pending &= ~(1u << k);
        if (pending && clock64() - t0 > (1ll << 32)) __trap();   // a CTA that never arrives is a bug, not a data condition
      }
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int k = 0; k < CO_SLOTS; ++k) { a0 += u0[k]; a1 += u1[k]; }     // index order; absent slots add +0.0
      if (published) __threadfence();                      // acquire: the publishers' boundary rows are visible from here on
      a0 = warp_allsum(a0); a1 = warp_allsum(a1);
      if (lane == 0) {
        bc[0] = a0; bc[1] = a1;
        if (prev_group >= 0) {
          const double empty = __longlong_as_double(-1ll);
          asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(slot_of(prev_group, rank)), "d"(empty), "d"(empty) : "memory");
        }
      }
    }
    __syncthreads();
    d0 = bc[0]; d1 = bc[1];
    prev_group = group;
  };

  // ---- stencil product y = A * hat(src) for this thread's element k -----------------------------------------------
  //   MODE 0: src = p (neighbours: p slice / private p halo rows)       hat = 1/M * p
  //   MODE 1: src = s = r - alpha v (neighbours: r, v slices / halo rows) hat = 1/M * s
  auto nb_in = [&](int mode, int idx, T alpha) -> T {
    T v_ = (mode == 1) ? (rs[idx] - alpha * vs[idx]) : ps[idx];
    if (PRE) v_ = ms[idx] * v_;
    return v_;
  };
  auto nb_halo = [&](int mode, bool lo, int j, T alpha) -> T {
    T v_;
    if (mode == 0) v_ = lo ? p_lo[j] : p_hi[j];
    else v_ = lo ? (r_lo[j] - alpha * v_lo[j]) : (r_hi[j] - alpha * v_hi[j]);
    if (PRE) v_ = (lo ? m_lo[j] : m_hi[j]) * v_;
    return v_;
  };
  auto spmv_elem = [&](int mode, int k, T centre_raw, T alpha) -> T {
    const int e = threadIdx.x + k * CO_THREADS, j = CJ(k);
    const unsigned f = ef[k];
    const T xc = PRE ? m[k] * centre_raw : centre_raw;
    T xw = T(0), xe = T(0), xn = T(0), xso = T(0);
    if (!(f & CF_ROW0)) xw = nb_in(mode, e - NY, alpha);
    else if (has_lo) xw = nb_halo(mode, true, j, alpha);
    if (!(f & CF_ROWN)) xe = nb_in(mode, e + NY, alpha);
    else if (has_hi) xe = nb_halo(mode, false, j, alpha);
    if (!(f & CF_JN)) xn = nb_in(mode, e + 1, alpha);
    if (!(f & CF_J0)) xso = nb_in(mode, e - 1, alpha);
    T acc;
    if (PRE) {   // the monolith's product (karman_taichi.py:362-367,401-406): flat offsets i-ny, i-1, i, i+1, i+ny from 0.0
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
  // warm-start residual (karman_taichi.py:360-367, no preconditioner): x0 straight from the caller's global plane
  auto x0_at = [&](int i, int j) -> T { return (i < 0 || i >= A.NX || j < 0 || j >= NY) ? T(0) : A.x[(size_t)i * A.pitch + j]; };
#define CO_EACH(k, e) _Pragma("unroll") for (int k = 0; k < EPT; ++k) if (const int e = threadIdx.x + k * CO_THREADS; ef[k] & CF_VALID)
#define CO_SCALARS(fin, a0, a1, h_) scalar_algebra(fin, &sc, a0, a1, h_, A.hist_cap)

  // ---- K0: r = b - A x0 (warm) or r = b ; r~ = r ; (r,r) -----------------------------------------------------------
  double d0 = 0.0, d1 = 0.0;
  CO_EACH(k, e) {
    const int li = e / NY, i = r0 + li, j = e - li * NY;
    T rr = A.b[(size_t)i * A.pitch + j];
    if (A.warm) {
      T acc = T(0) + coef(k, 1, e) * x0_at(i - 1, j);      // karman_taichi.py:362-367: the monolith's term order
      acc = acc + coef(k, 4, e) * x0_at(i, j - 1);
      acc = acc + coef(k, 0, e) * x[k];
      acc = acc + coef(k, 3, e) * x0_at(i, j + 1);
      acc = acc + coef(k, 2, e) * x0_at(i + 1, j);
      rr = rr - acc;
    }
    r[k] = rr; rt[k] = rr;                                 // r~ = r0 (= b when x0 = 0): bicgsolver.py:28, karman_taichi.py:371
    rs[e] = rr;
    publish(A.gr, ef[k], rr);
    if (A.is_cg) { p[k] = rr; ps[e] = rr; }                // cgsolver.py:90 p = r + 0 * p
    d0 += (double)rr * (double)rr;
  }
  grid_sum(d0, d1, true);
  CO_SCALARS(A.is_cg ? FIN_CG_INIT : FIN_BICG_INIT, d0, d0, hist);
  if (A.is_cg) {                                           // p = r on the halo rows too
    load_halo(r_lo, r_hi, A.gr);
    for (int j = threadIdx.x; j < NY; j += CO_THREADS) { if (has_lo) p_lo[j] = r_lo[j]; if (has_hi) p_hi[j] = r_hi[j]; }
    __syncthreads();
  }

  // ---- iterations ----------------------------------------------------------------------------------------------
  while (!sc.done) {
    if (!A.is_cg) {
      // P1 (the neighbours' r rows are final since the last grid barrier; their previous v rows were saved in P3)
      const bool first = sc.iter == 0;
      const T beta = T(sc.beta), omega_prev = T(sc.omega);
      for (int j = threadIdx.x; j < NY; j += CO_THREADS) {
        if (has_lo) { const T rr = ld_cg(A.gr + g_lo + j); r_lo[j] = rr; p_lo[j] = first ? rr : rr + beta * (p_lo[j] - omega_prev * v_lo[j]); }
        if (has_hi) { const T rr = ld_cg(A.gr + g_hi + j); r_hi[j] = rr; p_hi[j] = first ? rr : rr + beta * (p_hi[j] - omega_prev * v_hi[j]); }
      }
      CO_EACH(k, e) {
        const T pn = first ? r[k] : r[k] + beta * (p[k] - omega_prev * v[k]);
        p[k] = pn; ps[e] = pn;
      }
      __syncthreads();
      // P2
      d0 = 0.0; d1 = 0.0;
      CO_EACH(k, e) {
        const T y = spmv_elem(0, k, p[k], T(0));
        d0 += (double)rt[k] * (double)y;
        v[k] = y; vs[e] = y;
        publish(A.gv, ef[k], y);                           // neighbours read it in P3, after the barrier below
      }
      grid_sum(d0, d1, true);
      CO_SCALARS(FIN_BICG_ALPHA, d0, d1, nullptr);
      if (sc.done) break;
      // P3
      load_halo(v_lo, v_hi, A.gv);
      __syncthreads();
      const T alpha = T(sc.alpha);
      d0 = 0.0; d1 = 0.0;
      CO_EACH(k, e) {
        const T sraw = r[k] - alpha * v[k];                // bicgsolver.py:63
        const T y = spmv_elem(1, k, sraw, alpha);
        t[k] = y;
        d0 += (double)y * (double)sraw;
        d1 += (double)y * (double)y;
      }
      grid_sum(d0, d1, false);
      CO_SCALARS(FIN_BICG_OMEGA, d0, d1, nullptr);
      if (sc.done) break;
      // P4 (every CTA finished reading the r rows -- P3 -- before the barrier above)
      const T omega = T(sc.omega);
      d0 = 0.0; d1 = 0.0;
      CO_EACH(k, e) {
        const T sk = r[k] - alpha * v[k];                  // s again (same operation, same bits as in P3)
        const T ph = PRE ? m[k] * p[k] : p[k];
        const T sh = PRE ? m[k] * sk : sk;
        x[k] = x[k] + (alpha * ph + omega * sh);
        const T rn = sk - omega * t[k];
        r[k] = rn; rs[e] = rn;
        publish(A.gr, ef[k], rn);
        d0 += (double)rn * (double)rn;
        d1 += (double)rn * (double)rt[k];
      }
      grid_sum(d0, d1, true);
      CO_SCALARS(FIN_BICG_END, d0, d1, hist);
    } else {
      // C1: Ap = A p ; (p, Ap)
      d0 = 0.0; d1 = 0.0;
      CO_EACH(k, e) {
        const T y = spmv_elem(0, k, p[k], T(0));
        v[k] = y;
        d0 += (double)p[k] * (double)y;
      }
      grid_sum(d0, d1, false);
      CO_SCALARS(FIN_CG_ALPHA, d0, d1, nullptr);
      if (sc.done) break;
      // C2: x += alpha p ; r -= alpha Ap ; (r,r)
      const T alpha = T(sc.alpha);
      d0 = 0.0; d1 = 0.0;
      CO_EACH(k, e) {
        x[k] = x[k] + alpha * p[k];
        const T rn = r[k] - alpha * v[k];
        r[k] = rn;
        publish(A.gr, ef[k], rn);
        d0 += (double)rn * (double)rn;
      }
      grid_sum(d0, d1, true);
      CO_SCALARS(FIN_CG_END, d0, d1, hist);
      if (sc.done) break;
      // C3: p = r + beta p, halo rows included (the neighbours' r rows are final since the barrier above)
      const T beta = T(sc.beta);
      for (int j = threadIdx.x; j < NY; j += CO_THREADS) {
        if (has_lo) p_lo[j] = ld_cg(A.gr + g_lo + j) + beta * p_lo[j];
        if (has_hi) p_hi[j] = ld_cg(A.gr + g_hi + j) + beta * p_hi[j];
      }
      CO_EACH(k, e) {
        const T pn = r[k] + beta * p[k];
        p[k] = pn; ps[e] = pn;
      }
      __syncthreads();
    }
  }
  CO_EACH(k, e) {
    const int li = e / NY;
    const size_t g = (size_t)(r0 + li) * A.pitch + (e - li * NY);
    A.x[g] = x[k];
    A.gr[g] = r[k];
  }
  if (rank == 0 && threadIdx.x == 0) {
    Scalars out = sc;
    out.dot0 = out.dot1 = 0.0;
    *A.sc = out;
  }
#undef CO_EACH
#undef CO_SCALARS
#undef CJ
}

static int coop_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

constexpr int CO_EPT_REG = 2;       // coefficients in registers up to this many elements per thread
constexpr int CO_EPT_MAX = 6;

struct CoFit { int G, rows_per, ept; bool regc; size_t smem; };

template <typename T>
static CoFit coop_fit(const kt_handle_s* h, bool pre) {
  CoFit f = {0, 0, 0, false, 0};
  static int enabled = coop_env("KT_COOP", 1);
  if (!enabled || !h->allow_coop || h->row_lo != 0 || h->row_hi != h->NX) return f;
  const int rows_per = (h->NX + h->num_sms - 1) / h->num_sms;
  const long long cap = (long long)rows_per * h->NY;
  const int ept = (int)((cap + CO_THREADS - 1) / CO_THREADS);
  if (ept > CO_EPT_MAX) return f;
  const bool regc = ept <= CO_EPT_REG;
  const int planes = 3 + (pre ? 1 : 0) + (regc ? 0 : 5);
  const size_t need = ((size_t)planes * cap + 8 * (size_t)h->NY) * sizeof(T);
  if (need > 200 * 1024) return f;
  f.G = (h->NX + rows_per - 1) / rows_per; f.rows_per = rows_per; f.ept = ept; f.regc = regc; f.smem = need;
  if (f.G > 32 * CO_SLOTS) f.G = 0;
  return f;
}

template <typename K, typename T>
static int coop_launch_k(kt_handle_s* h, K kern, unsigned long long& configured, CoArgs<T>& A, const CoFit& f) {
  const unsigned long long dev_bit = 1ull << (h->device & 63);   // function attributes are per device
  if (!(configured & dev_bit)) {
    KT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured |= dev_bit;
  }
  KT_CUDA(cudaMemsetAsync(A.partials, 0xFF, sizeof(double) * 6 * (size_t)f.G, h->stream));     // every slot "empty"
  void* args[] = {(void*)&A};
  KT_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(f.G), dim3(CO_THREADS), args, f.smem, h->stream));
  h->launches++;
  return 0;
}

template <typename T, bool PRE>
static int coop_launch(kt_handle_s* h, CoArgs<T>& A, const CoFit& f) {
  static unsigned long long conf[6] = {0, 0, 0, 0, 0, 0};
  switch (f.ept) {
    case 1: return coop_launch_k(h, k_coop_solve<T, PRE, 1, true>, conf[0], A, f);
    case 2: return coop_launch_k(h, k_coop_solve<T, PRE, 2, true>, conf[1], A, f);
    case 3: return coop_launch_k(h, k_coop_solve<T, PRE, 3, false>, conf[2], A, f);
    case 4: return coop_launch_k(h, k_coop_solve<T, PRE, 4, false>, conf[3], A, f);
    case 5: return coop_launch_k(h, k_coop_solve<T, PRE, 5, false>, conf[4], A, f);
    case 6: return coop_launch_k(h, k_coop_solve<T, PRE, 6, false>, conf[5], A, f);
    default: set_error("coop_launch: elements per thread out of range"); return -2;
  }
}

// What coop_try_solve would launch for this geometry (kt_resident_plan): 0 = does not fit.
template <typename T>
int coop_plan(const kt_handle_s* h, bool pre, bool warm, bool is_cg, long long* out) {
  if (is_cg && (pre || warm)) return 0;
  const CoFit f = coop_fit<T>(h, pre);
  if (f.G == 0 || (size_t)6 * f.G > (size_t)KT_MAX_DOTS * h->max_blocks) return 0;
  out[0] = 3; out[1] = f.G; out[2] = f.rows_per; out[3] = CO_THREADS; out[4] = f.ept;
  return 1;
}
template int coop_plan<double>(const kt_handle_s*, bool, bool, bool, long long*);
template int coop_plan<float>(const kt_handle_s*, bool, bool, bool, long long*);

// Entry used by kt_bicgstab_solve / kt_cg_solve after the cluster path declined: returns 1 if the solve was done here,
// 0 if the geometry does not fit or a cooperative grid cannot be launched (caller falls through to the streaming
// kernels), < 0 on error.  r_plane / v_plane: the caller's r and Ap workspace planes.
template <typename T>
int coop_try_solve(kt_handle_s* h, const void* coef, const void* b, void* x, const void* Minv, void* r_plane, void* v_plane,
                   double* history, long long hist_cap, double eps, long long max_iter, int flags, bool is_cg) {
  const bool pre = Minv != nullptr, warm = (flags & KT_FLAG_WARM_START) != 0;
  if (is_cg && (pre || warm)) return 0;
  if (!r_plane || !v_plane) return 0;
  static int coop_ok = -1;
  if (coop_ok < 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, h->device) != cudaSuccess) { cudaGetLastError(); v = 0; }
    coop_ok = v;
  }
  if (!coop_ok) return 0;
  CoFit f = coop_fit<T>(h, pre);
  if (f.G == 0 || (size_t)6 * f.G > (size_t)KT_MAX_DOTS * h->max_blocks) return 0;
  CoArgs<T> A;
  A.coef = (const T*)coef; A.plane = (size_t)h->NX * h->pitch; A.b = (const T*)b; A.x = (T*)x; A.minv = (const T*)Minv;
  A.gr = (T*)r_plane; A.gv = (T*)v_plane; A.partials = h->partials;
  A.NX = h->NX; A.NY = h->NY; A.pitch = h->pitch; A.rows_per = f.rows_per; A.G = f.G;
  A.sc = h->sc; A.history = history; A.hist_cap = hist_cap; A.eps = eps; A.max_iter = max_iter; A.flags = flags;
  A.warm = warm ? 1 : 0; A.is_cg = is_cg ? 1 : 0;
  int rc = pre ? coop_launch<T, true>(h, A, f) : coop_launch<T, false>(h, A, f);
  if (rc) {
    // the grid cannot be co-resident here (e.g. a MIG slice or a busy device): streaming path
    cudaGetLastError();
    return 0;
  }
  KT_CUDA(cudaMemcpyAsync(h->sc_host, h->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, h->stream));
  KT_CUDA(cudaStreamSynchronize(h->stream));
  return 1;
}

template int coop_try_solve<double>(kt_handle_s*, const void*, const void*, void*, const void*, void*, void*, double*, long long, double,
                                    long long, int, bool);
template int coop_try_solve<float>(kt_handle_s*, const void*, const void*, void*, const void*, void*, void*, double*, long long, double,
                                   long long, int, bool);

}  // namespace kt

// every kernel of this translation unit, for kt_preload_kernels (kt_handle.h)
namespace kt {
int preload_kt_coop() {
  static const void* const fns[] = {
#define KT_PRELOAD_TU_kt_coop
#include "kt_preload_list.inc"
#undef KT_PRELOAD_TU_kt_coop
      nullptr};
  return preload_list(fns);
}
}  // namespace kt
