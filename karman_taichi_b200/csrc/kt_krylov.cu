// kt_krylov.cu -- matrix-free 5-point CG / BiCGSTAB for sm_100a.
//
// Replaces the Taichi kernels of cgsolver.py:31-81 and bicgsolver.py:21-86 (one launch per vector
// op, one host sync per inner product) and the serial mega-kernel karman_taichi.py:342-456 with
// three (BiCGSTAB: four) fused HBM-streaming kernels per iteration whose scalars never leave the
// device:
//
//   BiCGSTAB (28 N words / iteration, 31 N with the Jacobi plane):
//     K1 k_bicg_p      p = r + beta (p - omega v)                         4 N   bicgsolver.py:99-103
//     K2 k_stencil     v = A p^, (r~, v)           -> alpha               8 N   :104-108
//     K3 k_stencil     s = r - alpha v, t = A s^, (t,s), (t,t) -> omega   9 N   :110-118
//     K4 k_bicg_xr     x += alpha p^ + omega s^, r = s - omega t,
//                      (r,r), (r,r~)               -> test, rho, beta     7 N   :120-135,95-102
//   CG (16 N words / iteration):
//     C1 k_stencil     Ap = A p, (p, Ap)           -> alpha               7 N   cgsolver.py:94-96
//     C2 k_cg_xr       x += alpha p, r -= alpha Ap, (r,r) -> test, beta   6 N   :98-107
//     C3 k_cg_p        p = r + beta p                                     3 N   :109
//
// The stencil kernel marches each warp down a (TI rows x 32 vectors) tile keeping rows i-1, i, i+1 of
// the operand in registers (one load per element), gets the j+-1 neighbours by warp shuffle (edge lanes
// re-load one scalar, an L1/L2 hit), streams the five coefficient planes with 16-byte no-allocate
// loads, and accumulates the fused inner products in double.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "kt_handle.h"

namespace kt {

// ---------------------------------------------------------------------------------------------
// stencil kernel
// ---------------------------------------------------------------------------------------------
enum { SRC_X = 0, SRC_S = 1, SRC_P = 2 };
enum { EPI_Y = 0, EPI_DOT_Q = 1, EPI_DOT_SRC = 2, EPI_ST = 3, EPI_RESID = 4, EPI_DOT_Q_ST = 5 };

template <typename T> struct StencilArgs {
  const T* c[5];     // ap, aw, ae, an, as planes
  const T* x;        // SRC_X operand            | SRC_S: r                      | SRC_P: r
  const T* v;        //                          | SRC_S: v  (s = r - alpha v)   | SRC_P: p_old
  const T* w;        //                          |                               | SRC_P: v_old  (p = r + beta (p_old - omega v_old))
  const T* minv;     // PRE: Jacobi 1/M plane
  const T* q;        // EPI_DOT_Q / EPI_DOT_Q_ST: external dot vector; EPI_RESID: b
  T* y;              // A*src  (EPI_RESID: r = b - A*src)
  T* s_out;          // EPI_ST: s ; EPI_DOT_Q_ST: p_new ; EPI_RESID: r_tld copy (may be null)
  int NX, pitch, row_lo, row_hi, TI, nCC;
  long long nTiles;
  double alpha_imm;
  int alpha_from_sc;
  int check_done;
  int fin;
  ReduceWs ws;
  // P2P variants (x-slab peer-memory mode, exchange fused into the kernels; unused otherwise).
  //  in : the operand's halo rows row_lo - 1 / row_hi are read from this rank's receive buffer (hx_* for x, hv_* for v) once the
  //       neighbour's flag fx_* / fv_* shows seq_x / seq_v -- the neighbour's producer kernel stored them there;
  //  out: the first / last owned row of y (EPI_RESID: of r) is also stored into the neighbours' receive rows push_lo / push_hi;
  //       ws.hflag_* / ws.hseq announce them.
  const T* hx_lo; const T* hx_hi; const T* hv_lo; const T* hv_hi;
  const unsigned long long* fx_lo; const unsigned long long* fx_hi; const unsigned long long* fv_lo; const unsigned long long* fv_hi;
  unsigned long long seq_x, seq_v;
  T* push_lo; T* push_hi;
  // raise_*: the neighbours' flags for the rows of x that the PREVIOUS kernel of this stream (a p-update, which has no
  // finalising block and no barrier to pay for one) stored into their receive buffers: this kernel's first thread raises them
  // to seq_x before anything else -- stream order puts every store of that kernel before it
  unsigned long long* raise_lo; unsigned long long* raise_hi;
};

// scalars of the on-the-fly operand: SRC_S: a = alpha; SRC_P: a = beta, b = omega, first = (iteration 0: p = r)
template <typename T> struct SrcCoef { T a, b; bool first; };

template <typename T, int SRC, bool PRE, bool P2P = false>
struct SrcRow {
  // vector of the stencil operand at (i, j0..j0+V-1); also returns the un-preconditioned value (x, s or p)
  static __device__ __forceinline__ void load(const StencilArgs<T>& A, int i, int j0, bool act, const SrcCoef<T>& K,
                                              Pack<T>& hat, Pack<T>& raw) {
    constexpr int V = VT<T>::V;
    if (!act || i < 0 || i >= A.NX) { hat = pack_zero<T>(); raw = hat; return; }
    const size_t o = (size_t)i * A.pitch + j0;
    if (P2P && SRC != SRC_P && ((i == A.row_lo - 1 && A.hx_lo) || (i == A.row_hi && A.hx_hi))) {
      // halo row: from the receive buffer the neighbour's producer kernel filled (the caller has waited for its flag)
      const bool lo = i == A.row_lo - 1;
      raw = ld_sys((lo ? A.hx_lo : A.hx_hi) + j0);
      if (SRC == SRC_S) {
        const Pack<T> v = ld_sys((lo ? A.hv_lo : A.hv_hi) + j0);
#pragma unroll
        for (int k = 0; k < V; ++k) raw.a[k] = raw.a[k] - K.a * v.a[k];                 // bicgsolver.py:63
      }
    } else if (SRC == SRC_X) {
      raw = ld_cached(A.x + o);
    } else if (SRC == SRC_S) {
      Pack<T> r = ld_cached(A.x + o), v = ld_cached(A.v + o);
#pragma unroll
      for (int k = 0; k < V; ++k) raw.a[k] = r.a[k] - K.a * v.a[k];                     // bicgsolver.py:63
    } else {
      raw = ld_cached(A.x + o);
      if (!K.first) {
        Pack<T> pp = ld_cached(A.v + o), vv = ld_cached(A.w + o);
#pragma unroll
        for (int k = 0; k < V; ++k) raw.a[k] = raw.a[k] + K.a * (pp.a[k] - K.b * vv.a[k]);   // bicgsolver.py:46
      }
    }
    if (PRE) {
      Pack<T> m = ld_cached(A.minv + o);
#pragma unroll
      for (int k = 0; k < V; ++k) hat.a[k] = m.a[k] * raw.a[k];
    } else {
      hat = raw;
    }
  }
  // scalar of the (preconditioned) operand at (i, j); j may be out of [0, pitch)
  static __device__ __forceinline__ T scalar(const StencilArgs<T>& A, int i, int j, const SrcCoef<T>& K) {
    if (j < 0 || j >= A.pitch || i < 0 || i >= A.NX) return T(0);
    const size_t o = (size_t)i * A.pitch + j;
    T raw;
    if (SRC == SRC_X) raw = A.x[o];
    else if (SRC == SRC_S) raw = A.x[o] - K.a * A.v[o];
    else { raw = A.x[o]; if (!K.first) raw = raw + K.a * (A.v[o] - K.b * A.w[o]); }
    if (PRE) raw = A.minv[o] * raw;
    return raw;
  }
};

#ifndef KT_STENCIL_MINB
#define KT_STENCIL_MINB 4
#endif
#ifndef KT_STENCIL_UNROLL
#define KT_STENCIL_UNROLL 1
#endif
constexpr int kStencilUnroll = KT_STENCIL_UNROLL;   // rows of loads the compiler may batch ahead
constexpr int kEwUnroll = KT_EW_UNROLL;             // strides of loads batched ahead in element-wise kernels

// Column chunks (32 vectors each) a warp processes side by side, and the register budget that goes with it
// (measured, profiles/r01_grid_sweep.md): the plain-operand product (v = A p, Ap = A p) is fastest with two chunks
// per warp at 2 blocks/SM (14 x 512 B in flight per warp); the fused s = r - alpha v / Jacobi variants, which carry
// more operand rows, with one chunk per warp at 4 blocks/SM (<= 64 registers).
#ifdef KT_STENCIL_W
template <int SRC, bool PRE> struct StencilCfg { static constexpr int W = KT_STENCIL_W; static constexpr int MINB = KT_STENCIL_MINB; };
#else
template <int SRC, bool PRE> struct StencilCfg { static constexpr int W = 1; static constexpr int MINB = 4; };
template <> struct StencilCfg<SRC_X, false> { static constexpr int W = 2; static constexpr int MINB = 2; };
template <bool PRE> struct StencilCfg<SRC_P, PRE> { static constexpr int W = 1; static constexpr int MINB = 3; };
#endif

// Waits of a P2P stencil kernel for the neighbour's rows (all lanes poll the same word; every lane runs its own acquire fence)
template <typename T, int SRC>
__device__ __forceinline__ void stencil_halo_wait(const StencilArgs<T>& A, bool lo) {
  const unsigned long long* fx = lo ? A.fx_lo : A.fx_hi;
  const unsigned long long* fv = lo ? A.fv_lo : A.fv_hi;
  if (fx) p2p_wait(fx, A.seq_x, A.ws.sc, A.ws.timeout_ns, 1u << 8 | (lo ? 0u : 1u << 16) | 0x10u);
  if (SRC == SRC_S && fv) p2p_wait(fv, A.seq_v, A.ws.sc, A.ws.timeout_ns, 1u << 8 | (lo ? 0u : 1u << 16) | 0x11u);
}

template <typename T, int SRC, bool PRE, int EPI, bool P2P = false>
__global__ void __launch_bounds__(256, (StencilCfg<SRC, PRE>::MINB)) k_stencil(const StencilArgs<T> A) {
  constexpr int V = VT<T>::V;
  constexpr int W = StencilCfg<SRC, PRE>::W;
  // The Jacobi-preconditioned and the warm-start-residual products are the monolith's (karman_taichi.py:342-456); the
  // class solvers (cgsolver.py / bicgsolver.py) have neither.  The two write the same five products in a different order.
  constexpr bool MONO = PRE || EPI == EPI_RESID;
  pdl_prologue();
  if (A.check_done && A.ws.sc->done) return;
  if (P2P && blockIdx.x == 0 && threadIdx.x == 0 && (A.raise_lo || A.raise_hi)) {
    __threadfence_system();
    if (A.raise_lo) *(volatile unsigned long long*)A.raise_lo = A.seq_x;
    if (A.raise_hi) *(volatile unsigned long long*)A.raise_hi = A.seq_x;
  }
  SrcCoef<T> alpha;   // (named for its SRC_S role)
  alpha.a = T(0); alpha.b = T(0); alpha.first = false;
  if (SRC == SRC_S) alpha.a = T(A.alpha_from_sc ? A.ws.sc->alpha : A.alpha_imm);
  if (SRC == SRC_P) { alpha.a = T(A.ws.sc->beta); alpha.b = T(A.ws.sc->omega); alpha.first = A.ws.sc->iter == 0; }
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long long gwarp = (long long)blockIdx.x * wpb + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * wpb;
  double acc0 = 0.0, acc1 = 0.0;
  bool pushed = false;     // P2P: this thread stored part of a boundary row into a neighbour's receive buffer

  for (long long tile = gwarp; tile < A.nTiles; tile += nwarps) {
    const int rchunk = (int)(tile / A.nCC), cc = (int)(tile - (long long)rchunk * A.nCC);
    const int i0 = A.row_lo + rchunk * A.TI;
    const int i1 = min(i0 + A.TI, A.row_hi);
    int j0[W];
    bool act[W];
    Pack<T> xw[W], xc[W], xe[W];   // (preconditioned) operand rows i-1, i, i+1
    Pack<T> raw_c[W], raw_e[W];    // un-preconditioned operand (s or x) rows i, i+1
    if (P2P && i0 == A.row_lo && A.hx_lo) stencil_halo_wait<T, SRC>(A, true);     // row i0 - 1 comes from the lower neighbour
#pragma unroll
    for (int w = 0; w < W; ++w) {
      j0[w] = ((cc * W + w) * 32 + lane) * V;
      act[w] = j0[w] < A.pitch;
      SrcRow<T, SRC, PRE, P2P>::load(A, i0 - 1, j0[w], act[w], alpha, xw[w], raw_c[w]);
      SrcRow<T, SRC, PRE, P2P>::load(A, i0, j0[w], act[w], alpha, xc[w], raw_c[w]);
    }
#pragma unroll kStencilUnroll
    for (int i = i0; i < i1; ++i) {
      // ---- issue every load of this row step: W x (5 coefficient planes [+ q] + operand row i+1)
      Pack<T> c[W][5], qv[W];
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const size_t o = (size_t)i * A.pitch + j0[w];
        if (act[w]) {
#pragma unroll
          for (int k = 0; k < 5; ++k) c[w][k] = ld_stream(A.c[k] + o);
          if (EPI == EPI_DOT_Q || EPI == EPI_DOT_Q_ST || EPI == EPI_RESID) qv[w] = ld_stream(A.q + o);
        } else {
#pragma unroll
          for (int k = 0; k < 5; ++k) c[w][k] = pack_zero<T>();
          qv[w] = pack_zero<T>();
        }
      }
      if (P2P && i + 1 == A.row_hi && A.hx_hi) stencil_halo_wait<T, SRC>(A, false);  // row i + 1 comes from the upper neighbour
#pragma unroll
      for (int w = 0; w < W; ++w) SrcRow<T, SRC, PRE, P2P>::load(A, i + 1, j0[w], act[w], alpha, xe[w], raw_e[w]);
      // ---- j-1 / j+V neighbours of the centre row: warp shuffles; only the two outer edge lanes reload
      T xs[W], xn[W];
#pragma unroll
      for (int w = 0; w < W; ++w) { xs[w] = shfl_up1(xc[w].a[V - 1]); xn[w] = shfl_down1(xc[w].a[0]); }
#pragma unroll
      for (int w = 1; w < W; ++w) {
        const T from_prev = __shfl_sync(0xffffffffu, xc[w - 1].a[V - 1], 31);
        const T from_next = __shfl_sync(0xffffffffu, xc[w].a[0], 0);
        if (lane == 0) xs[w] = from_prev;
        if (lane == 31) xn[w - 1] = from_next;
      }
      if (lane == 0) xs[0] = SrcRow<T, SRC, PRE, P2P>::scalar(A, i, j0[0] - 1, alpha);
      if (lane == 31) xn[W - 1] = SrcRow<T, SRC, PRE, P2P>::scalar(A, i, j0[W - 1] + V, alpha);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const size_t o = (size_t)i * A.pitch + j0[w];
        Pack<T> y;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const T north = (k < V - 1) ? xc[w].a[k + 1 < V ? k + 1 : k] : xn[w];
          const T south = (k > 0) ? xc[w].a[k > 0 ? k - 1 : 0] : xs[w];
          T acc;
          if (MONO) {
            // karman_taichi.py:362-367,401-406,421-426: Ap[i] = 0.0, then the flat offsets i-ny (x[i-1,j]), i-1 (x[i,j-1]),
            // i, i+1 (x[i,j+1]), i+ny (x[i+1,j]) in that order (the outer two of the seven never hold a non-zero)
            acc = T(0) + c[w][1].a[k] * xw[w].a[k];
            acc = acc + c[w][4].a[k] * south;
            acc = acc + c[w][0].a[k] * xc[w].a[k];
            acc = acc + c[w][3].a[k] * north;
            acc = acc + c[w][2].a[k] * xe[w].a[k];
          } else {
            // cgsolver.py:59-60 left-to-right: ap*x + aw*x[i-1] + ae*x[i+1] + an*x[j+1] + as*x[j-1]
            acc = c[w][0].a[k] * xc[w].a[k];
            acc = acc + c[w][1].a[k] * xw[w].a[k];
            acc = acc + c[w][2].a[k] * xe[w].a[k];
            acc = acc + c[w][3].a[k] * north;
            acc = acc + c[w][4].a[k] * south;
          }
          y.a[k] = acc;
        }
        if (act[w]) {
          if (EPI == EPI_RESID) {
            Pack<T> r;
#pragma unroll
            for (int k = 0; k < V; ++k) { r.a[k] = qv[w].a[k] - y.a[k]; acc0 += (double)r.a[k] * (double)r.a[k]; }
            st_vec(A.y + o, r);
            if (A.s_out) st_vec(A.s_out + o, r);
            if (P2P) {
              if (i == A.row_lo && A.push_lo) { st_vec(A.push_lo + j0[w], r); pushed = true; }
              if (i == A.row_hi - 1 && A.push_hi) { st_vec(A.push_hi + j0[w], r); pushed = true; }
            }
          } else {
            st_vec(A.y + o, y);
            if (P2P) {
              if (i == A.row_lo && A.push_lo) { st_vec(A.push_lo + j0[w], y); pushed = true; }
              if (i == A.row_hi - 1 && A.push_hi) { st_vec(A.push_hi + j0[w], y); pushed = true; }
            }
            if ((EPI == EPI_ST && A.s_out) || EPI == EPI_DOT_Q_ST) st_vec(A.s_out + o, raw_c[w]);
#pragma unroll
            for (int k = 0; k < V; ++k) {
              if (EPI == EPI_DOT_Q || EPI == EPI_DOT_Q_ST) acc0 += (double)qv[w].a[k] * (double)y.a[k];
              if (EPI == EPI_DOT_SRC) acc0 += (double)raw_c[w].a[k] * (double)y.a[k];
              if (EPI == EPI_ST) { acc0 += (double)y.a[k] * (double)raw_c[w].a[k]; acc1 += (double)y.a[k] * (double)y.a[k]; }
            }
          }
        }
        xw[w] = xc[w]; xc[w] = xe[w]; raw_c[w] = raw_e[w];
      }
    }
  }
  if (EPI == EPI_Y) return;
  if (EPI == EPI_ST) block_reduce_finalize<2>(acc0, acc1, A.ws, A.fin, P2P && pushed);
  else block_reduce_finalize<1>(acc0, acc1, A.ws, A.fin, P2P && pushed);
}

// ---------------------------------------------------------------------------------------------
// element-wise kernels (flat over the owned rows; 16-byte vectors; static grid-stride assignment)
// ---------------------------------------------------------------------------------------------
struct EwGeom { long long v_lo, v_hi; };   // vector-index range [row_lo*pitch/V, row_hi*pitch/V)

enum { EW_COPY, EW_FILL, EW_AXPY, EW_XPAY, EW_XMAY, EW_BICG_P, EW_BICG_X };

// Every element-wise kernel walks its statically assigned vectors in batches of U strides: the U x (arrays)
// 16-byte loads of a batch are issued back to back before the first dependent instruction, so each thread
// keeps U x arrays x 16 B in flight (explicit batching -- ptxas will not hoist a possibly-aliasing load above a
// store on its own).  KT_EW_FOR(u, vi) iterates the batch; out-of-range strides are predicated off.
#define KT_EW_OUTER(U)                                                                                     \
  const long long stride = (long long)gridDim.x * blockDim.x;                                              \
  for (long long vb = g.v_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; vb < g.v_hi; vb += stride * (U))
#define KT_EW_FOR(U, u, o)                                                  \
  _Pragma("unroll") for (int u = 0; u < (U); ++u)                           \
    if (const size_t o = (size_t)(vb + u * stride) * V; vb + u * stride < g.v_hi)

template <typename T, int OP>
__global__ void __launch_bounds__(256) k_ew(EwGeom g, T* out, const T* x, const T* y, const T* z, double a_, double b_) {
  constexpr int V = VT<T>::V;
  constexpr int U = kEwUnroll;
  const T a = T(a_), b = T(b_);
  KT_EW_OUTER(U) {
    Pack<T> p0[U], p1[U], p2[U];
    KT_EW_FOR(U, u, o) {
      if (OP == EW_COPY) p0[u] = ld_stream(x + o);
      if (OP == EW_AXPY) { p0[u] = ld_rw(out + o); p1[u] = ld_stream(x + o); }
      if (OP == EW_XPAY || OP == EW_XMAY) { p0[u] = ld_rw(x + o); p1[u] = ld_rw(y + o); }
      if (OP == EW_BICG_P) { p0[u] = ld_stream(x + o); p1[u] = ld_rw(out + o); p2[u] = ld_stream(y + o); }
      if (OP == EW_BICG_X) { p0[u] = ld_rw(out + o); p1[u] = ld_stream(x + o); p2[u] = ld_stream(y + o); }
    }
    KT_EW_FOR(U, u, o) {
      Pack<T> r;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        if (OP == EW_COPY) r.a[k] = p0[u].a[k];
        if (OP == EW_FILL) r.a[k] = a;
        if (OP == EW_AXPY) r.a[k] = p0[u].a[k] + a * p1[u].a[k];                             // cgsolver.py:71,76
        if (OP == EW_XPAY) r.a[k] = p0[u].a[k] + a * p1[u].a[k];                             // cgsolver.py:81
        if (OP == EW_XMAY) r.a[k] = p0[u].a[k] - a * p1[u].a[k];                             // bicgsolver.py:63,86
        if (OP == EW_BICG_P) r.a[k] = p0[u].a[k] + a * (p1[u].a[k] - b * p2[u].a[k]);        // bicgsolver.py:46
        if (OP == EW_BICG_X) r.a[k] = p0[u].a[k] + (a * p1[u].a[k] + b * p2[u].a[k]);        // bicgsolver.py:81
      }
      st_vec(out + o, r);
    }
  }
}

// fill with a non-zero value: logical columns only -- the pad columns [NY, pitch) stay zero (every kernel multiplies and sums
// over whole padded rows)
template <typename T>
__global__ void __launch_bounds__(256) k_fill_cols(T* dst, int pitch, int NY, int row_lo, int row_hi, T value) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= NY) return;
  for (int i = row_lo + (int)blockIdx.y; i < row_hi; i += (int)gridDim.y) dst[(size_t)i * pitch + j] = value;
}

template <typename T>
__global__ void __launch_bounds__(256) k_dot(EwGeom g, const T* p, const T* q, ReduceWs ws) {
  constexpr int V = VT<T>::V;
  constexpr int U = kEwUnroll;
  double acc = 0.0;
  KT_EW_OUTER(U) {
    Pack<T> a[U], b[U];
    KT_EW_FOR(U, u, o) { a[u] = ld_stream(p + o); b[u] = ld_stream(q + o); }
    KT_EW_FOR(U, u, o) {
#pragma unroll
      for (int k = 0; k < V; ++k) acc += (double)a[u].a[k] * (double)b[u].a[k];
    }
  }
  block_reduce_finalize<1>(acc, 0.0, ws, FIN_NONE);
}

// K0: r = b, (r_tld = b), x = 0, [p = r for CG]; dot (r,r).   bicgsolver.py:21-36 / cgsolver.py:31-39,90
// P2P: the boundary rows of r (BiCGSTAB: the operand of the first t = A s) / of p (CG) also go to the neighbours (hp)
template <typename T, bool CG, bool P2P = false>
__global__ void __launch_bounds__(256) k_init(EwGeom g, const T* b, T* r, T* r_tld, T* x, T* p, ReduceWs ws, int fin,
                                              HaloPush<T> hp = HaloPush<T>()) {
  constexpr int V = VT<T>::V;
  constexpr int U = kEwUnroll;
  double acc = 0.0;
  bool pushed = false;
  pdl_prologue();
  KT_EW_OUTER(U) {
    Pack<T> bb[U];
    KT_EW_FOR(U, u, o) bb[u] = ld_stream(b + o);
    KT_EW_FOR(U, u, o) {
      if (P2P) pushed |= ew_push(hp, o, bb[u]);
      st_vec(r + o, bb[u]);
      if (r_tld) st_vec(r_tld + o, bb[u]);
      if (CG) st_vec(p + o, bb[u]);
      st_vec(x + o, pack_zero<T>());
#pragma unroll
      for (int k = 0; k < V; ++k) acc += (double)bb[u].a[k] * (double)bb[u].a[k];
    }
  }
  block_reduce_finalize<1>(acc, 0.0, ws, fin, P2P && pushed);
}

// K1: p = r + beta (p - omega v); first iteration p = r (bicgsolver.py:99-103)
// P2P: the boundary rows of the new p also go into the neighbours' receive buffers; the NEXT kernel of the stream (v = A p^)
// announces them (StencilArgs::raise_*), so this one keeps running without a barrier or a last-block ticket
template <typename T, bool P2P = false>
__global__ void __launch_bounds__(256) k_bicg_p(EwGeom g, T* p, const T* r, const T* v, const Scalars* sc,
                                                HaloPush<T> hp = HaloPush<T>()) {
  constexpr int V = VT<T>::V;
  constexpr int U = kEwUnroll;
  pdl_prologue();
  if (sc->done) return;
  const bool first = sc->iter == 0;
  const T beta = T(sc->beta), omega = T(sc->omega);
  KT_EW_OUTER(U) {
    Pack<T> rr[U], pp[U], vv[U];
    KT_EW_FOR(U, u, o) {
      rr[u] = ld_stream(r + o);
      if (!first) { pp[u] = ld_rw(p + o); vv[u] = ld_stream(v + o); }
    }
    KT_EW_FOR(U, u, o) {
      Pack<T> out = rr[u];
      if (!first) {
#pragma unroll
        for (int k = 0; k < V; ++k) out.a[k] = rr[u].a[k] + beta * (pp[u].a[k] - omega * vv[u].a[k]);
      }
      st_vec(p + o, out);
      if (P2P) ew_push(hp, o, out);
    }
  }
}

// K4: x += alpha p^ + omega s^ ; r = s - omega t ; dots (r,r), (r,r_tld)   (bicgsolver.py:120-123, 95)
template <typename T, bool PRE, bool P2P = false>
__global__ void __launch_bounds__(256) k_bicg_xr(EwGeom g, T* x, T* r, const T* p, const T* s, const T* t,
                                                 const T* r_tld, const T* minv, double alpha_imm, double omega_imm,
                                                 int from_sc, ReduceWs ws, int fin, const T* v_for_s = nullptr,
                                                 HaloPush<T> hp = HaloPush<T>()) {
  constexpr int V = VT<T>::V;
  constexpr int U = (kEwUnroll > 2) ? 2 : kEwUnroll;     // 5-6 operand planes: 2 strides already keep ~190 B in flight
  pdl_prologue();
  if (from_sc && ws.sc->done) return;
  const T alpha = T(from_sc ? ws.sc->alpha : alpha_imm), omega = T(from_sc ? ws.sc->omega : omega_imm);
  double a0 = 0.0, a1 = 0.0;
  bool pushed = false;
  KT_EW_OUTER(U) {
    Pack<T> xx[U], pp[U], ss[U], tt[U], rt[U], mm[U];
    KT_EW_FOR(U, u, o) {
      xx[u] = ld_rw(x + o); pp[u] = ld_stream(p + o); tt[u] = ld_stream(t + o);
      rt[u] = ld_stream(r_tld + o);
      if (PRE) mm[u] = ld_stream(minv + o);
      if (v_for_s) {
        // s was not stored by K3: recompute s = r - alpha v (same expression, same bits) from the old r and v
        Pack<T> ro = ld_rw(r + o), vv = ld_stream(v_for_s + o);
#pragma unroll
        for (int k = 0; k < V; ++k) ss[u].a[k] = ro.a[k] - alpha * vv.a[k];
      } else {
        ss[u] = ld_stream(s + o);
      }
    }
    KT_EW_FOR(U, u, o) {
      Pack<T> rn;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const T ph = PRE ? mm[u].a[k] * pp[u].a[k] : pp[u].a[k];
        const T sh = PRE ? mm[u].a[k] * ss[u].a[k] : ss[u].a[k];
        xx[u].a[k] = xx[u].a[k] + (alpha * ph + omega * sh);
        rn.a[k] = ss[u].a[k] - omega * tt[u].a[k];
        a0 += (double)rn.a[k] * (double)rn.a[k];
        a1 += (double)rn.a[k] * (double)rt[u].a[k];
      }
      st_vec(x + o, xx[u]);
      st_vec(r + o, rn);
      if (P2P) pushed |= ew_push(hp, o, rn);
    }
  }
  block_reduce_finalize<2>(a0, a1, ws, fin, P2P && pushed);
}

// C2: x += alpha p ; r -= alpha Ap ; dot (r,r)   (cgsolver.py:98-101)
template <typename T>
__global__ void __launch_bounds__(256) k_cg_xr(EwGeom g, T* x, T* r, const T* p, const T* Ap, double alpha_imm,
                                               int from_sc, ReduceWs ws, int fin) {
  constexpr int V = VT<T>::V;
  constexpr int U = (kEwUnroll > 2) ? 2 : kEwUnroll;
  pdl_prologue();
  if (from_sc && ws.sc->done) return;
  const T alpha = T(from_sc ? ws.sc->alpha : alpha_imm);
  double a0 = 0.0;
  KT_EW_OUTER(U) {
    Pack<T> xx[U], rr[U], pp[U], aa[U];
    KT_EW_FOR(U, u, o) { xx[u] = ld_rw(x + o); rr[u] = ld_rw(r + o); pp[u] = ld_stream(p + o); aa[u] = ld_stream(Ap + o); }
    KT_EW_FOR(U, u, o) {
#pragma unroll
      for (int k = 0; k < V; ++k) {
        xx[u].a[k] = xx[u].a[k] + alpha * pp[u].a[k];
        rr[u].a[k] = rr[u].a[k] - alpha * aa[u].a[k];
        a0 += (double)rr[u].a[k] * (double)rr[u].a[k];
      }
      st_vec(x + o, xx[u]);
      st_vec(r + o, rr[u]);
    }
  }
  block_reduce_finalize<1>(a0, 0.0, ws, fin);
}

// C3: p = r + beta p   (cgsolver.py:78-81)
template <typename T, bool P2P = false>
__global__ void __launch_bounds__(256) k_cg_p(EwGeom g, T* p, const T* r, const Scalars* sc, HaloPush<T> hp = HaloPush<T>()) {
  constexpr int V = VT<T>::V;
  constexpr int U = kEwUnroll;
  pdl_prologue();
  if (sc->done) return;
  const T beta = T(sc->beta);
  KT_EW_OUTER(U) {
    Pack<T> rr[U], pp[U];
    KT_EW_FOR(U, u, o) { rr[u] = ld_stream(r + o); pp[u] = ld_rw(p + o); }
    KT_EW_FOR(U, u, o) {
#pragma unroll
      for (int k = 0; k < V; ++k) pp[u].a[k] = rr[u].a[k] + beta * pp[u].a[k];
      st_vec(p + o, pp[u]);
      if (P2P) ew_push(hp, o, pp[u]);
    }
  }
}

__global__ void k_sc_reset(Scalars* sc, double eps, long long max_iter, int flags) {
  sc->rho = 0.0; sc->rho_1 = 1.0; sc->alpha = 1.0; sc->beta = 1.0; sc->omega = 1.0;  // bicgsolver.py:32-36
  sc->rTr0 = sc->rTr = sc->old_rTr = 0.0; sc->dot0 = sc->dot1 = 0.0;
  sc->eps = eps; sc->iter = 0; sc->max_iter = max_iter; sc->status = 1; sc->done = 0; sc->first = 1; sc->flags = flags;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Plan {
  int V;
  long long ew_blocks;   // blocks that would give every thread exactly one vector
  EwGeom g;
};

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// Launch of a streaming Krylov kernel (256 threads).  With KT_PDL != 0 (default on) the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization: the kernel may be scheduled while its predecessor in the stream drains
// (every such kernel starts with pdl_prologue(), kt_common.cuh), which hides the launch latency and the block ramp-up between
// the four dependent kernels of an iteration -- what is left of a 100-300 us kernel's fixed cost.
static bool pdl_on(const kt_handle_s* h) {
  static int on = env_int("KT_PDL", 1);
  return on != 0 && !h->capturing;
}

template <typename... KArgs, typename... Args>
static void launch_k(kt_handle_s* h, void (*kernel)(KArgs...), int grid, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
  cudaLaunchAttribute at;
  memset(&at, 0, sizeof(at));
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &at; cfg.numAttrs = pdl_on(h) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Grid sizing.  Work is assigned statically (tile / vector index = f(blockIdx, threadIdx, stride)), so the
// summation order of every inner product depends only on the problem geometry and the grid -- never on which SM
// ran which block.  Measured on B200 (profiles/r01_grid_sweep.md): many small blocks that the hardware scheduler
// hands out as SMs free up beat a grid of exactly-resident persistent blocks by 3-5 % (SMs differ in effective
// bandwidth; a static persistent split leaves the fast ones idle at the end), while too many blocks make the
// last-block sum of the block partials visible.  256-512 KB of traffic per block is the measured sweet spot at
// both 4096^2 and 16384^2, so: blocks = bytes moved / KT_BLOCK_BYTES (512 KB), clamped to [4, 128] blocks per SM.
// Whole waves.  A kernel of a few waves (grid / resident blocks) that ends on a fraction of a wave idles the rest of that wave:
// 3.5 waves run like 4 (the t = A s kernel at 4096^2 sat at 0.89 of the copy peak when its neighbours reached 0.95), 1.76 like
// 2 (config 4 cut 8 ways; measured on one GPU at 2600^2: 0.892 -> 0.940 of the copy peak, at 1300^2 0.907 -> 0.966).  Grids of
// up to 16 waves are therefore rounded to the nearest whole number of waves.  `occ` = resident blocks per SM of the kernel,
// from its register count (cuobjdump --dump-resource-usage: 57..64 registers -> 4 blocks of 256 threads, 65..80 -> 3,
// 106..128 -> 2, 40 -> 6; the stencil kernels' __launch_bounds__ min-blocks).  Forcing the element-wise kernels into these
// classes with __launch_bounds__(256, n) was tried and reverted: it changed their register allocation and cost the p-update
// 7-15 %.  Pure integer arithmetic on the geometry, so kt_reduce_plan (the summation order the oracle follows) sees the same grid.
enum { EWP_DOT = 2, EWP_INIT = 4, EWP_BICG_XR = 8, EWP_BICG_XR_UNIT = 7, EWP_CG_XR = 6 };   // planes the reducing element-wise kernels are sized by
constexpr int OCC_BICG_P = 3, OCC_CG_P = 4;
static int occ_of_planes(int planes_moved) {
  switch (planes_moved) {
    case EWP_DOT: return 4;            // k_dot
    case EWP_INIT: return 6;           // k_init
    case EWP_CG_XR: return 4;          // k_cg_xr
    case EWP_BICG_XR_UNIT: case EWP_BICG_XR: return 3;   // k_bicg_xr
    default: return 0;
  }
}

static int grid_for(const kt_handle_s* h, long long blocks_wanted, int planes_moved, int occ = -1) {
  static long long block_bytes = env_int("KT_BLOCK_BYTES", 1 << 19);
  static int wave_round = env_int("KT_WAVE_ROUND", 1);
  const long long bytes = (long long)planes_moved * (h->row_hi - h->row_lo) * h->pitch * (h->dtype == KT_F64 ? 8 : 4);
  long long g = bytes / block_bytes;
  g = std::max<long long>(g, (long long)h->num_sms * 4);
  g = std::min<long long>(g, (long long)h->num_sms * 128);
  if (occ < 0) occ = occ_of_planes(planes_moved);
  if (wave_round && occ > 0) {
    const long long R = (long long)h->num_sms * occ;
    if (g > R && g < 16 * R) g = (g + R / 2) / R * R;
  }
  g = std::min<long long>(g, h->max_blocks);
  return (int)std::max<long long>(1, std::min<long long>(blocks_wanted, g));
}

struct StPlan { int TI, nCC, grid; long long nTiles; };

// Stencil launch geometry for W column chunks per warp: grid from the traffic, then the tallest row tile (<= 64 rows,
// >= 8) that still leaves every warp of the grid >= 4 tiles to stride over.
static StPlan stencil_plan(const kt_handle_s* h, int V, int W, int planes_moved, int occ) {
  StPlan S;
  const int rows = h->row_hi - h->row_lo;
  S.nCC = (h->pitch / V + 32 * W - 1) / (32 * W);
  const long long max_tiles = (long long)((rows + 7) / 8) * S.nCC;
  S.grid = grid_for(h, (max_tiles + 7) / 8, planes_moved, occ);
  static int ti_env = env_int("KT_TI", 0);
  int TI = 64;
  while (TI > 8 && (long long)((rows + TI - 1) / TI) * S.nCC < 4LL * S.grid * 8) TI >>= 1;
  if (ti_env > 0) TI = ti_env;
  S.TI = TI;
  S.nTiles = (long long)((rows + TI - 1) / TI) * S.nCC;
  S.grid = (int)std::max<long long>(1, std::min<long long>(S.grid, (S.nTiles + 7) / 8));
  return S;
}

static Plan make_plan(const kt_handle_s* h) {
  Plan P;
  P.V = (h->dtype == KT_F64) ? 2 : 4;
  P.g.v_lo = (long long)h->row_lo * h->pitch / P.V;
  P.g.v_hi = (long long)h->row_hi * h->pitch / P.V;
  P.ew_blocks = (P.g.v_hi - P.g.v_lo + 255) / 256;
  return P;
}

static ReduceWs make_ws(kt_handle_s* h, double* history = nullptr, long long cap = 0, double* dots_out = nullptr) {
  ReduceWs ws;
  memset(&ws, 0, sizeof(ws));
  ws.partials = h->partials; ws.ticket = h->ticket; ws.max_blocks = h->max_blocks; ws.sc = h->sc;
  ws.history = history; ws.history_cap = cap; ws.dots_out = dots_out;
  return ws;
}

template <typename T>
static StencilArgs<T> stencil_args(kt_handle_s* h, const Plan& P, const void* coef) {
  StencilArgs<T> A;
  memset(&A, 0, sizeof(A));
  const size_t plane = (size_t)h->NX * h->pitch;
  for (int k = 0; k < 5; ++k) A.c[k] = (const T*)coef + k * plane;
  A.NX = h->NX; A.pitch = h->pitch; A.row_lo = h->row_lo; A.row_hi = h->row_hi;
  A.ws = make_ws(h);
  return A;
}

// planes streamed by a stencil variant: 5 coefficients + operand(s) + outputs (+ q / Minv) -- sizes its grid
constexpr int stencil_planes(int SRC, bool PRE, int EPI) {
  return 5 + (SRC == SRC_S ? 2 : SRC == SRC_P ? 3 : 1) + (PRE ? 1 : 0) + 1 +
         (EPI == EPI_DOT_Q || EPI == EPI_DOT_Q_ST || EPI == EPI_RESID ? 1 : 0) + (EPI == EPI_ST || EPI == EPI_DOT_Q_ST ? 1 : 0);
}

template <typename T, int SRC, bool PRE, int EPI, bool P2P = false>
static int launch_stencil(kt_handle_s* h, const Plan& P, const StencilArgs<T>& A) {
  constexpr int planes = stencil_planes(SRC, PRE, EPI);
  const StPlan S = stencil_plan(h, P.V, StencilCfg<SRC, PRE>::W, planes, StencilCfg<SRC, PRE>::MINB);
  StencilArgs<T> B = A;
  B.TI = S.TI; B.nCC = S.nCC; B.nTiles = S.nTiles;
  launch_k(h, k_stencil<T, SRC, PRE, EPI, P2P>, S.grid, B);
  KT_LAUNCH_CHECK(h);
  return 0;
}

static_assert(sizeof(Scalars) % 8 == 0, "Scalars is published in 8-byte words");
static int read_dots(kt_handle_s* h, double* out, int n) {
  if (int rc = publish(h, h->sc, h->sc_host, (int)(sizeof(Scalars) / 8))) return rc;
  KT_CUDA(cudaStreamSynchronize(h->stream));
  out[0] = h->sc_host->dot0;
  if (n > 1) out[1] = h->sc_host->dot1;
  return 0;
}

// ---- unit entry points --------------------------------------------------------------------------
template <typename T>
static int t_spmv(kt_handle_s* h, const void* coef, const void* x, void* y, const void* q, double* dot_host) {
  Plan P = make_plan(h);
  StencilArgs<T> A = stencil_args<T>(h, P, coef);
  A.x = (const T*)x; A.y = (T*)y; A.q = (const T*)q; A.fin = FIN_NONE;
  if (!q) return launch_stencil<T, SRC_X, false, EPI_Y>(h, P, A);
  int rc = launch_stencil<T, SRC_X, false, EPI_DOT_Q>(h, P, A);
  if (rc) return rc;
  return read_dots(h, dot_host, 1);
}

template <typename T>
static int t_bicg_s_t(kt_handle_s* h, const void* coef, const void* r, const void* v, double alpha, void* s, void* t,
                      double* dots_host) {
  Plan P = make_plan(h);
  StencilArgs<T> A = stencil_args<T>(h, P, coef);
  A.x = (const T*)r; A.v = (const T*)v; A.y = (T*)t; A.s_out = (T*)s; A.alpha_imm = alpha; A.fin = FIN_NONE;
  int rc = launch_stencil<T, SRC_S, false, EPI_ST>(h, P, A);
  if (rc) return rc;
  return read_dots(h, dots_host, 2);
}

template <typename T, int OP>
static int t_ew(kt_handle_s* h, void* out, const void* x, const void* y, const void* z, double a, double b) {
  Plan P = make_plan(h);
  k_ew<T, OP><<<grid_for(h, P.ew_blocks, 3), 256, 0, h->stream>>>(P.g, (T*)out, (const T*)x, (const T*)y, (const T*)z, a, b);
  KT_LAUNCH_CHECK(h);
  return 0;
}

template <typename T>
static int t_dot(kt_handle_s* h, const void* p, const void* q, double* out_host) {
  Plan P = make_plan(h);
  k_dot<T><<<grid_for(h, P.ew_blocks, EWP_DOT), 256, 0, h->stream>>>(P.g, (const T*)p, (const T*)q, make_ws(h));
  KT_LAUNCH_CHECK(h);
  return read_dots(h, out_host, 1);
}

// ---- solvers ------------------------------------------------------------------------------------
static int sc_reset(kt_handle_s* h, double eps, long long max_iter, int flags) {
  k_sc_reset<<<1, 1, 0, h->stream>>>(h->sc, eps, max_iter, flags);
  KT_LAUNCH_CHECK(h);
  return 0;
}

// ---- CUDA-graph replay of iteration chunks ---------------------------------------------------------------------
static bool graphs_wanted(const kt_handle_s* h) {
  static int enabled = env_int("KT_GRAPH", 1);
  static long long max_elems = env_int("KT_GRAPH_MAX_ELEMS", 1 << 22);   // planes up to 32 MB (f64): kernels of <= ~40 us
  return enabled && !h->prof_on && h->graphs_state >= 0 && (long long)(h->row_hi - h->row_lo) * h->pitch <= max_elems;
}

// Returns the cached executable graph of `chunk` iterations for this argument set, capturing it first if needed
// (nullptr: graphs unavailable here -- the caller enqueues directly).
template <typename EnqueueIter>
static kt_handle_s::GraphEntry* graph_for(kt_handle_s* h, const std::vector<const void*>& key, int chunk, EnqueueIter&& enqueue_iter,
                                          int* rc_out) {
  *rc_out = 0;
  for (auto& g : h->graphs)
    if (g.key == key) { g.stamp = ++h->graph_clock; return &g; }
  if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    h->graphs_state = -1;
    return nullptr;
  }
  const long long l0 = h->launches;
  int rc = 0;
  h->capturing = true;
  for (int k = 0; k < chunk && rc == 0; ++k) rc = enqueue_iter();
  h->capturing = false;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  const long long per_chunk = h->launches - l0;
  h->launches = l0;                      // nothing ran yet: launches are counted per replay
  if (rc) { if (graph) cudaGraphDestroy(graph); *rc_out = rc; return nullptr; }
  cudaGraphExec_t exec = nullptr;
  if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  if (e != cudaSuccess) { cudaGetLastError(); h->graphs_state = -1; return nullptr; }
  h->graphs_state = 1;
  if (h->graphs.size() >= 8) {           // small LRU: a SIMPLE loop alternates between three argument sets at most
    size_t victim = 0;
    for (size_t i = 1; i < h->graphs.size(); ++i) if (h->graphs[i].stamp < h->graphs[victim].stamp) victim = i;
    cudaGraphExecDestroy(h->graphs[victim].exec);
    h->graphs.erase(h->graphs.begin() + victim);
  }
  h->graphs.push_back({key, exec, per_chunk, ++h->graph_clock});
  return &h->graphs.back();
}

// Poll loop shared by both solvers: enqueue `chunk` iterations at a time; the device-side `done` flag makes
// every later kernel a no-op, so the host may notice convergence late without changing the result.
// `graph_key` (may be null): every pointer / flag the enqueued launches depend on; when given and the grid is in the
// launch-bound range, full chunks are replayed from a CUDA graph on the handle's private stream.
template <typename EnqueueIter>
static int run_chunks(kt_handle_s* h, long long max_iter, EnqueueIter&& enqueue_iter, const std::vector<const void*>* graph_key = nullptr) {
  const int chunk = 16;                  // even: the ping-pong parity of the fused variant is the same in every chunk
  long long enq = 0;
  int slot = 0;
  bool pending[2] = {false, false};
  cudaStream_t caller_stream = h->stream;
  struct Restore { kt_handle_s* h; cudaStream_t s; ~Restore() { h->stream = s; } } restore{h, caller_stream};
  bool use_graph = graph_key && max_iter >= chunk && graphs_wanted(h);
  if (use_graph) {
    if (!h->gstream) {
      if (cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&h->gev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); h->graphs_state = -1; use_graph = false; }
    }
    if (use_graph) {
      KT_CUDA(cudaEventRecord(h->gev, caller_stream));        // everything enqueued so far (K0, the caller's kernels)
      KT_CUDA(cudaStreamWaitEvent(h->gstream, h->gev, 0));
      h->stream = h->gstream;                                  // launches and polls below go to the private stream
    }
  }
  while (enq < max_iter) {   // iteration max_iter always sets `done`, so nothing beyond it is ever needed
    long long n = std::min<long long>(chunk, max_iter - enq);
    kt_handle_s::GraphEntry* ge = nullptr;
    if (use_graph && n == chunk) {
      int rc = 0;
      ge = graph_for(h, *graph_key, chunk, enqueue_iter, &rc);
      if (rc) return rc;
      if (!ge) use_graph = false;                              // capture refused: direct launches from here on
    }
    if (ge) {
      KT_CUDA(cudaGraphLaunch(ge->exec, h->stream));
      h->launches += ge->launches;
    } else {
      for (long long k = 0; k < n; ++k) { int rc = enqueue_iter(); if (rc) return rc; }
    }
    enq += n;
    if (int rc = publish(h, h->sc, h->sc_host + slot, (int)(sizeof(Scalars) / 8))) return rc;   // (a kernel, not a copy: kt_handle.h)
    KT_CUDA(cudaEventRecord(h->ev[slot], h->stream));
    pending[slot] = true;
    const int other = slot ^ 1;
    if (pending[other]) {
      KT_CUDA(cudaEventSynchronize(h->ev[other]));
      pending[other] = false;
      if (h->sc_host[other].done) break;
    }
    slot = other;
  }
  if (int rc = publish(h, h->sc, h->sc_host, (int)(sizeof(Scalars) / 8))) return rc;
  KT_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

template <typename T>
static int t_bicgstab(kt_handle_s* h, const void* coef, const void* b, void* x, const void* Minv, const kt_krylov_ws* w,
                      double eps, long long max_iter, int flags, long* iters, double* rel_res, int* status,
                      bool fixed_iters) {
  Plan P = make_plan(h);
  const bool pre = Minv != nullptr;
  const bool warm = (flags & KT_FLAG_WARM_START) != 0;
  T* r = (T*)w->r; T* p = (T*)w->p; T* v = (T*)w->Ap; T* s = (T*)w->s; T* t = (T*)w->t;
  T* r_tld = (T*)w->r_tld;
  KT_CHECK_ARG(r && p && v && t, "kt_krylov_ws: r, p, Ap, t are required for BiCGSTAB");
  // ws->s == NULL: K3 does not store s and K4 recomputes it from r and v -- same 28 N words, one write stream less
  // (K3 becomes 7 reads : 1 write like K2) and one plane of HBM saved.
  KT_CHECK_ARG(!(warm && !r_tld), "warm start needs a separate r_tld plane");
  const T* r_tld_c = r_tld ? r_tld : (const T*)b;     // x0 = 0  =>  r_tld == b exactly (bicgsolver.py:25-26)
  int rc = cluster_try_solve<T>(h, coef, b, x, Minv, r, w->history, w->history_cap, fixed_iters ? -1.0 : eps, max_iter, flags, false);
  if (rc < 0) return rc;
  if (rc == 0) {       // mid-size grid: one persistent cooperative kernel (kt_coop.cu)
    rc = coop_try_solve<T>(h, coef, b, x, Minv, r, v, w->history, w->history_cap, fixed_iters ? -1.0 : eps, max_iter, flags, false);
    if (rc < 0) return rc;
  }
  if (rc == 1) {       // solved inside one resident kernel (kt_cluster.cu / kt_coop.cu)
    const Scalars& S = *h->sc_host;
    if (iters) *iters = (long)S.iter;
    if (rel_res) *rel_res = sqrt(S.rTr / S.rTr0);
    if (status) *status = S.status;
    return 0;
  }
  rc = sc_reset(h, fixed_iters ? -1.0 : eps, max_iter, flags);
  if (rc) return rc;
  ReduceWs ws = make_ws(h, w->history, w->history_cap);
  // K0
  if (warm) {
    StencilArgs<T> A = stencil_args<T>(h, P, coef);
    A.ws = ws; A.x = (const T*)x; A.q = (const T*)b; A.y = r; A.s_out = r_tld; A.fin = FIN_BICG_INIT;
    rc = launch_stencil<T, SRC_X, false, EPI_RESID>(h, P, A);
    if (rc) return rc;
  } else {
    k_init<T, false><<<grid_for(h, P.ew_blocks, EWP_INIT), 256, 0, h->stream>>>(P.g, (const T*)b, r, r_tld, (T*)x, nullptr, ws, FIN_BICG_INIT);
    KT_LAUNCH_CHECK(h);
  }
  // Fused K1+K2 (27 N words / iteration instead of 28 N) when the caller provides the second p / v planes: the
  // p-update is recomputed on the fly for the three operand rows of v = A p^ (halo recompute), which needs p_old and
  // v_old to stay intact while p_new / v_new are written -> ping-pong between (p, Ap) and (p2, Ap2).
  T* pb[2] = {p, (T*)w->p2};
  T* vb[2] = {v, (T*)w->Ap2};
  const bool fused = pb[1] != nullptr && vb[1] != nullptr;
  long long it_enq = 0;                       // parity of the iteration being enqueued selects the buffers
  StencilArgs<T> A2 = stencil_args<T>(h, P, coef);
  A2.ws = ws; A2.minv = (const T*)Minv; A2.q = r_tld_c; A2.check_done = 1; A2.fin = FIN_BICG_ALPHA;
  StencilArgs<T> A3 = stencil_args<T>(h, P, coef);
  A3.ws = ws; A3.x = r; A3.minv = (const T*)Minv; A3.y = t; A3.s_out = s; A3.alpha_from_sc = 1;
  A3.check_done = 1; A3.fin = FIN_BICG_OMEGA;
  auto iter = [&]() -> int {
    const int cur = fused ? (int)(it_enq & 1) : 0, prev = fused ? cur ^ 1 : 0;
    ++it_enq;
    T* p_new = pb[cur]; T* v_new = vb[cur];
    int e;
    if (fused) {
      ProfScope ps(h, KT_K_BICG_V);
      A2.x = r; A2.v = pb[prev]; A2.w = vb[prev]; A2.y = v_new; A2.s_out = p_new;
      e = pre ? launch_stencil<T, SRC_P, true, EPI_DOT_Q_ST>(h, P, A2) : launch_stencil<T, SRC_P, false, EPI_DOT_Q_ST>(h, P, A2);
      if (e) return e;
    } else {
      {
        ProfScope ps(h, KT_K_BICG_P);
        launch_k(h, k_bicg_p<T, false>, grid_for(h, P.ew_blocks, 4, OCC_BICG_P), P.g, p_new, r, v_new, h->sc, HaloPush<T>());
      }
      KT_LAUNCH_CHECK(h);
      ProfScope ps(h, KT_K_BICG_V);
      A2.x = p_new; A2.y = v_new;
      e = pre ? launch_stencil<T, SRC_X, true, EPI_DOT_Q>(h, P, A2) : launch_stencil<T, SRC_X, false, EPI_DOT_Q>(h, P, A2);
      if (e) return e;
    }
    {
      ProfScope ps(h, KT_K_BICG_ST);
      A3.v = v_new;
      e = pre ? launch_stencil<T, SRC_S, true, EPI_ST>(h, P, A3) : launch_stencil<T, SRC_S, false, EPI_ST>(h, P, A3);
    }
    if (e) return e;
    ProfScope ps(h, KT_K_BICG_XR);
    const T* v_for_s = s ? nullptr : v_new;
    if (pre) launch_k(h, k_bicg_xr<T, true, false>, grid_for(h, P.ew_blocks, EWP_BICG_XR), P.g, (T*)x, r, p_new, s, t, r_tld_c, (const T*)Minv, 0, 0, 1, ws, FIN_BICG_END, v_for_s, HaloPush<T>());
    else launch_k(h, k_bicg_xr<T, false, false>, grid_for(h, P.ew_blocks, EWP_BICG_XR), P.g, (T*)x, r, p_new, s, t, r_tld_c, nullptr, 0, 0, 1, ws, FIN_BICG_END, v_for_s, HaloPush<T>());
    KT_LAUNCH_CHECK(h);
    return 0;
  };
  const std::vector<const void*> gkey = {(const void*)1, coef, b, x, Minv, r, p, v, s, t, r_tld, w->p2, w->Ap2, w->history,
                                         (const void*)(size_t)w->history_cap, (const void*)(size_t)h->row_lo, (const void*)(size_t)h->row_hi,
                                         (const void*)(size_t)sizeof(T)};
  rc = run_chunks(h, max_iter, iter, &gkey);
  if (rc) return rc;
  const Scalars& S = *h->sc_host;
  if (iters) *iters = (long)S.iter;
  if (rel_res) *rel_res = sqrt(S.rTr / S.rTr0);
  if (status) *status = S.status;
  return 0;
}

template <typename T>
static int t_cg(kt_handle_s* h, const void* coef, const void* b, void* x, const kt_krylov_ws* w, double eps,
                long long max_iter, long* iters, double* rel_res, int* status, bool fixed_iters) {
  Plan P = make_plan(h);
  T* r = (T*)w->r; T* p = (T*)w->p; T* Ap = (T*)w->Ap;
  KT_CHECK_ARG(r && p && Ap, "kt_krylov_ws: r, p, Ap are required for CG");
  int rc = cluster_try_solve<T>(h, coef, b, x, nullptr, r, w->history, w->history_cap, fixed_iters ? -1.0 : eps, max_iter, 0, true);
  if (rc < 0) return rc;
  if (rc == 0) {
    rc = coop_try_solve<T>(h, coef, b, x, nullptr, r, Ap, w->history, w->history_cap, fixed_iters ? -1.0 : eps, max_iter, 0, true);
    if (rc < 0) return rc;
  }
  if (rc == 1) {
    const Scalars& S = *h->sc_host;
    if (iters) *iters = (long)S.iter;
    if (rel_res) *rel_res = sqrt(S.rTr / S.rTr0);
    if (status) *status = S.status;
    return 0;
  }
  rc = sc_reset(h, fixed_iters ? -1.0 : eps, max_iter, 0);
  if (rc) return rc;
  ReduceWs ws = make_ws(h, w->history, w->history_cap);
  k_init<T, true><<<grid_for(h, P.ew_blocks, EWP_INIT), 256, 0, h->stream>>>(P.g, (const T*)b, r, nullptr, (T*)x, p, ws, FIN_CG_INIT);
  KT_LAUNCH_CHECK(h);
  StencilArgs<T> A1 = stencil_args<T>(h, P, coef);
  A1.ws = ws; A1.x = p; A1.y = Ap; A1.check_done = 1; A1.fin = FIN_CG_ALPHA;
  auto iter = [&]() -> int {
    int e;
    {
      ProfScope ps(h, KT_K_CG_AP);
      e = launch_stencil<T, SRC_X, false, EPI_DOT_SRC>(h, P, A1);
    }
    if (e) return e;
    {
      ProfScope ps(h, KT_K_CG_XR);
      launch_k(h, k_cg_xr<T>, grid_for(h, P.ew_blocks, EWP_CG_XR), P.g, (T*)x, r, p, Ap, 0.0, 1, ws, FIN_CG_END);
    }
    KT_LAUNCH_CHECK(h);
    {
      ProfScope ps(h, KT_K_CG_P);
      launch_k(h, k_cg_p<T, false>, grid_for(h, P.ew_blocks, 3, OCC_CG_P), P.g, p, r, h->sc, HaloPush<T>());
    }
    KT_LAUNCH_CHECK(h);
    return 0;
  };
  const std::vector<const void*> gkey = {(const void*)2, coef, b, x, r, p, Ap, w->history, (const void*)(size_t)w->history_cap,
                                         (const void*)(size_t)h->row_lo, (const void*)(size_t)h->row_hi, (const void*)(size_t)sizeof(T)};
  rc = run_chunks(h, max_iter, iter, &gkey);
  if (rc) return rc;
  const Scalars& S = *h->sc_host;
  if (iters) *iters = (long)S.iter;
  if (rel_res) *rel_res = sqrt(S.rTr / S.rTr0);
  if (status) *status = S.status;
  return 0;
}


// ---- x-slab (multi-GPU) phases: the same kernels, but the last block only publishes the LOCAL sums into a
// caller-owned device buffer; the host all-reduces that buffer across ranks (NCCL) and kt_dist_scalar then runs
// the scalar algebra on every rank from the identical reduced values.
__global__ void k_scalar_step(Scalars* sc, int fin, const double* dots, double* history, long long cap) {
  if (sc->done) return;
  scalar_algebra(fin, sc, dots[0], dots[1], history, cap);
}

template <typename T>
static int t_dist_phase(kt_handle_s* h, int phase, const void* coef, const void* a0, const void* a1, const void* a2,
                        void* o0, void* o1, void* o2, double* dots_dev, const void* v_for_s = nullptr) {
  Plan P = make_plan(h);
  ReduceWs ws = make_ws(h, nullptr, 0, dots_dev);
  static const int cls_of_phase[9] = {KT_K_INIT, KT_K_BICG_P, KT_K_BICG_V, KT_K_BICG_ST, KT_K_BICG_XR,
                                      KT_K_INIT, KT_K_CG_AP, KT_K_CG_XR, KT_K_CG_P};
  ProfScope ps(h, (phase >= 0 && phase < 9) ? cls_of_phase[phase] : KT_K_INIT);
  switch (phase) {
    case 0: {   // init: r = b (a0), x = 0 (o1), r (o0), r_tld (o2, may be null); local (r,r)
      k_init<T, false><<<grid_for(h, P.ew_blocks, EWP_INIT), 256, 0, h->stream>>>(P.g, (const T*)a0, (T*)o0, (T*)o2, (T*)o1, nullptr, ws, FIN_NONE);
      break;
    }
    case 1: {   // p (o0) = r (a0) + beta (p - omega v (a1))
      k_bicg_p<T><<<grid_for(h, P.ew_blocks, 4, OCC_BICG_P), 256, 0, h->stream>>>(P.g, (T*)o0, (const T*)a0, (const T*)a1, h->sc);
      break;
    }
    case 2: {   // v (o0) = A p (a0); local (r_tld (a1), v)
      StencilArgs<T> A = stencil_args<T>(h, P, coef);
      A.ws = ws; A.x = (const T*)a0; A.q = (const T*)a1; A.y = (T*)o0; A.check_done = 1; A.fin = FIN_NONE;
      return launch_stencil<T, SRC_X, false, EPI_DOT_Q>(h, P, A);
    }
    case 3: {   // s (o0) = r (a0) - alpha v (a1); t (o1) = A s; local (t,s), (t,t)
      StencilArgs<T> A = stencil_args<T>(h, P, coef);
      A.ws = ws; A.x = (const T*)a0; A.v = (const T*)a1; A.y = (T*)o1; A.s_out = (T*)o0; A.alpha_from_sc = 1; A.check_done = 1;
      A.fin = FIN_NONE;
      return launch_stencil<T, SRC_S, false, EPI_ST>(h, P, A);
    }
    case 4: {   // x (o0) += alpha p (a0) + omega s (a1); r (o1) = s - omega t (a2); local (r,r), (r, r_tld (o2))
      k_bicg_xr<T, false><<<grid_for(h, P.ew_blocks, EWP_BICG_XR_UNIT), 256, 0, h->stream>>>(
          P.g, (T*)o0, (T*)o1, (const T*)a0, (const T*)a1, (const T*)a2, (const T*)o2, nullptr, 0, 0, 1, ws, FIN_NONE, (const T*)v_for_s);
      break;
    }
    // ---- CG (cgsolver.py:83-114) ----
    case 5: {   // init: r (o0) = b (a0), x (o1) = 0, p (o2) = r; local (r,r)
      k_init<T, true><<<grid_for(h, P.ew_blocks, EWP_INIT), 256, 0, h->stream>>>(P.g, (const T*)a0, (T*)o0, nullptr, (T*)o1, (T*)o2, ws, FIN_NONE);
      break;
    }
    case 6: {   // Ap (o0) = A p (a0); local (p, Ap)
      StencilArgs<T> A = stencil_args<T>(h, P, coef);
      A.ws = ws; A.x = (const T*)a0; A.y = (T*)o0; A.check_done = 1; A.fin = FIN_NONE;
      return launch_stencil<T, SRC_X, false, EPI_DOT_SRC>(h, P, A);
    }
    case 7: {   // x (o0) += alpha p (a0); r (o1) -= alpha Ap (a1); local (r,r)
      k_cg_xr<T><<<grid_for(h, P.ew_blocks, EWP_CG_XR), 256, 0, h->stream>>>(P.g, (T*)o0, (T*)o1, (const T*)a0, (const T*)a1, 0.0, 1, ws, FIN_NONE);
      break;
    }
    case 8: {   // p (o0) = r (a0) + beta p
      k_cg_p<T><<<grid_for(h, P.ew_blocks, 3, OCC_CG_P), 256, 0, h->stream>>>(P.g, (T*)o0, (const T*)a0, h->sc);
      break;
    }
    default: set_error("kt_dist_bicg_phase: unknown phase"); return -2;
  }
  KT_LAUNCH_CHECK(h);
  return 0;
}


// ---------------------------------------------------------------------------------------------
// x-slab BiCGSTAB over NVLink peer memory (no NCCL in the loop)
// ---------------------------------------------------------------------------------------------
// Every rank maps its two neighbours' small halo-receive buffers and every rank's comm block (CUDA IPC); the planes
// themselves are ordinary allocations nobody else maps.  The halo exchange and the all-reduce are two tiny kernels that
// live on the same stream as the big ones:
//   k_p2p_halo  : one block per side stores this rank's boundary row straight into the neighbour's receive buffer over
//                 NVLink (16-byte peer stores), fences system-wide, bumps the neighbour's flag, waits for the neighbour's
//                 row to land in its own receive buffer (flag >= seq) and copies it into the halo row of the local plane.
//   k_p2p_reduce: lane r stores this rank's local sums into slot [rank] of rank r's comm block as tagged 8-byte words
//                 (p2p_allreduce_warp, kt_common.cuh); the warp then polls its own block until all nranks slots carry the
//                 tag, adds them in rank order (every rank gets the same bits) and runs the step's scalar algebra --
//                 all-reduce + k_scalar_step in one launch.
// Flags are monotonically increasing sequence numbers (never reset), so a late reader can never see a stale "ready".
// Write-after-read safety: a neighbour can only overwrite a receive row of iteration k after it has seen this rank's
// sums of a later kernel (the sums are published after the halo kernel that copied the row out, in stream order).
// Slot reuse of the sums: passing a sum proves that every rank has WRITTEN its slot, not that every rank has READ all of
// them, so two consecutive sums never share a slot group: BiCGSTAB INIT I, ALPHA 0, OMEGA 1, END 2; CG INIT I, ALPHA 0,
// END 1 (a rank can be at most one sum ahead of the slowest reader of the previous one).  I = 6 + (solve epoch & 1): the
// INIT sum of a solve follows the last sum of the PREVIOUS solve on the same handle, whichever that was (END, or INIT itself
// when the right-hand side was already converged), so it alternates between two groups of its own.
// Solves use sum groups 0-2, 6-7 and halo slots 0-2 (p / v / r); the standalone calls (kt_p2p_exchange, kt_p2p_allreduce: the
// SIMPLE kernels around the solves) rotate over their own sum groups 3-5 and halo slots 3-4, so that two consecutive
// exchanges never meet in a slot a slow neighbour may still be reading.
// how long a wait may last (default 30 s; KT_P2P_TIMEOUT_MS: tests use a short one)
static long long p2p_timeout_ns() {
  static long long ns = (long long)env_int("KT_P2P_TIMEOUT_MS", 30000) * 1000000LL;
  return ns;
}

// halo-receive buffer of a rank: [plane p/v/r][0: row from the lower neighbour, 1: row from the upper neighbour][pitch]
template <typename T>
__global__ void __launch_bounds__(1024) k_p2p_halo(T* plane, int pitch, int row_first, int row_last, T* recv_lo, T* recv_hi, T* recv_mine,
                                                   P2PBlock* mine, P2PBlock* blk_lo, P2PBlock* blk_hi, int which, unsigned long long seq,
                                                   Scalars* sc, int in_solve, long long timeout_ns) {
  constexpr int V = VT<T>::V;
  __shared__ int ok;
  // Once the solve is over (every rank reaches `done` at the same logical step, from identical sums) nobody needs
  // another row: return without signalling or waiting, so ranks whose hosts enqueued a different number of surplus
  // iterations never wait for each other.  (A standalone exchange is unconditional.)
  if (in_solve && sc->done) return;
  const int side = blockIdx.x;                                   // 0: talk to the lower neighbour, 1: to the upper one
  T* peer = side == 0 ? recv_lo : recv_hi;
  if (!peer) return;
  // my first owned row is the lower neighbour's "from upper" row (1); my last owned row the upper neighbour's "from lower" (0)
  const T* src = plane + (size_t)(side == 0 ? row_first : row_last) * pitch;
  T* dst = peer + (size_t)(which * 2 + (side == 0 ? 1 : 0)) * pitch;
  for (int j = threadIdx.x * V; j < pitch; j += blockDim.x * V) st_vec(dst + j, ld_rw(src + j));
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    P2PBlock* pb = side == 0 ? blk_lo : blk_hi;
    volatile unsigned long long* f = &pb->halo_flag[which][side == 0 ? 1 : 0];
    *f = seq;
    ok = p2p_wait(&mine->halo_flag[which][side == 0 ? 0 : 1], seq, sc, timeout_ns, 1u << 8 | (unsigned)which | (unsigned)side << 16) ? 1 : 0;
  }
  __syncthreads();
  if (!ok) return;
  // the neighbour's row has landed in my receive buffer: move it into the plane's halo row (row_first - 1 / row_last + 1)
  const T* got = recv_mine + (size_t)(which * 2 + (side == 0 ? 0 : 1)) * pitch;
  T* halo = plane + (size_t)(side == 0 ? row_first - 1 : row_last + 1) * pitch;
  for (int j = threadIdx.x * V; j < pitch; j += blockDim.x * V) st_vec(halo + j, ld_rw(got + j));
}

struct P2PPeers { P2PBlock* blk[16]; };

__global__ void k_p2p_reduce(P2PPeers peers, int rank, int nranks, int group, int ndot, const double* local, int fin,
                             unsigned long long seq, Scalars* sc, double* history, long long cap, int in_solve, int op_max,
                             long long timeout_ns, double imm0 = 0.0, double imm1 = 0.0) {
  if (in_solve && sc->done) return;
  // local == NULL: the two values arrive as kernel arguments (kt_p2p_allreduce with host values: no host-to-device copy)
  p2p_allreduce_warp(peers.blk, rank, nranks, group, ndot, local ? local[0] : imm0, local ? (ndot > 1 ? local[1] : 0.0) : imm1, fin, seq, sc,
                     history, cap, in_solve, op_max, timeout_ns);
}

// Sequence numbers: (solve epoch << 32) | exchange index within the solve.  The epoch advances once per solve on every
// rank, the index once per enqueued exchange; surplus exchanges enqueued after convergence never touch a flag, so the
// ranks stay consistent even though their hosts notice `done` at different times.  The state belongs to the handle (one
// handle per set of comm blocks); every rank drives its handle through the same sequence of solves.
struct P2PCtx {
  kt_handle_s* h; const kt_p2p_comm* cm; P2PPeers peers; P2PBlock* mine; P2PBlock* blk_lo; P2PBlock* blk_hi; double* dots;
  unsigned long long base;
};

static int p2p_begin(kt_handle_s* h, const kt_p2p_comm* cm, P2PCtx* c, bool new_solve = true) {
  KT_CHECK_ARG(cm->nranks >= 1 && cm->nranks <= 16 && cm->rank >= 0 && cm->rank < cm->nranks, "kt_p2p_comm: bad rank / nranks (max 16)");
  c->h = h; c->cm = cm;
  for (int k = 0; k < 16; ++k) c->peers.blk[k] = k < cm->nranks ? (P2PBlock*)cm->comm_block[k] : nullptr;
  c->mine = c->peers.blk[cm->rank];
  c->blk_lo = cm->rank > 0 ? c->peers.blk[cm->rank - 1] : nullptr;
  c->blk_hi = cm->rank < cm->nranks - 1 ? c->peers.blk[cm->rank + 1] : nullptr;
  c->dots = (double*)cm->local_dots;
  KT_CHECK_ARG(c->dots && c->mine, "kt_p2p_comm: local_dots / comm_block missing");
  KT_CHECK_ARG(cm->nranks == 1 || cm->halo_recv, "kt_p2p_comm: halo_recv missing");
  KT_CHECK_ARG((cm->rank == 0 || cm->peer_halo_recv[0]) && (cm->rank == cm->nranks - 1 || cm->peer_halo_recv[1]),
               "kt_p2p_comm: a neighbour's halo-receive buffer is not mapped");
  if (!new_solve) return 0;
  ++h->p2p_epoch;
  for (int k = 0; k < 3; ++k) h->p2p_halo_k[k] = 0;
  for (int k = 0; k < 8; ++k) h->p2p_dot_k[k] = 0;
  c->base = h->p2p_epoch << 32;
  return 0;
}

// Standalone halo exchange of one plane (the SIMPLE kernels between the solves read u, v, p, p' and ap_u one row beyond
// the slab): same kernel, own rotating slots.
template <typename T>
static int t_p2p_exchange(kt_handle_s* h, const kt_p2p_comm* cm, void* plane) {
  P2PCtx c;
  int rc = p2p_begin(h, cm, &c, false);
  if (rc) return rc;
  if (cm->nranks == 1) return 0;
  const unsigned long long seq = ++h->p2p_xchg_seq;
  const int slot = 3 + (int)(seq & 1ull);
  k_p2p_halo<T><<<2, 1024, 0, h->stream>>>((T*)plane, h->pitch, h->row_lo, h->row_hi - 1, (T*)cm->peer_halo_recv[0], (T*)cm->peer_halo_recv[1],
                                           (T*)cm->halo_recv, c.mine, c.blk_lo, c.blk_hi, slot, seq, h->sc, 0, p2p_timeout_ns());
  KT_LAUNCH_CHECK(h);
  return 0;
}

template <typename T>
static int p2p_halo(P2PCtx& c, int which, T* plane) {
  kt_handle_s* h = c.h;
  if (c.cm->nranks == 1) return 0;
  ++h->p2p_halo_k[which];
  k_p2p_halo<T><<<2, 1024, 0, h->stream>>>(plane, h->pitch, h->row_lo, h->row_hi - 1, (T*)c.cm->peer_halo_recv[0], (T*)c.cm->peer_halo_recv[1],
                                           (T*)c.cm->halo_recv, c.mine, c.blk_lo, c.blk_hi, which, c.base | h->p2p_halo_k[which], h->sc, 1, p2p_timeout_ns());
  KT_LAUNCH_CHECK(h);
  return 0;
}

static int p2p_reduce(P2PCtx& c, int group, int ndot, int fin, double* history, long long cap) {
  kt_handle_s* h = c.h;
  ++h->p2p_dot_k[group];
  k_p2p_reduce<<<1, 32, 0, h->stream>>>(c.peers, c.cm->rank, c.cm->nranks, group, ndot, c.dots, fin, c.base | h->p2p_dot_k[group], h->sc,
                                        history, cap, 1, 0, p2p_timeout_ns());
  KT_LAUNCH_CHECK(h);
  return 0;
}

#define KT_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

// BICGSolver.solve (bicgsolver.py:88-135) over x-slabs; with Minv != NULL / KT_FLAG_WARM_START the monolith's Jacobi
// variant (karman_taichi.py:342-456: warm start needs ws->r_tld and the halo rows of x, exchanged through plane slot 0).
template <typename T>
static int t_p2p_bicgstab_unfused(kt_handle_s* h, const void* coef, const void* b, void* x, const void* Minv, const kt_krylov_ws* w,
                          const kt_p2p_comm* cm, double eps, long long max_iter, int flags, bool fixed_iters, long* iters, double* rel_res,
                          int* status) {
  T* r = (T*)w->r; T* p = (T*)w->p; T* v = (T*)w->Ap; T* s = (T*)w->s; T* t = (T*)w->t; T* r_tld = (T*)w->r_tld;
  KT_CHECK_ARG(r && p && v && t, "kt_krylov_ws: r, p, Ap, t are required (s optional: NULL = recomputed, not stored)");
  const bool warm = (flags & KT_FLAG_WARM_START) != 0, pre = Minv != nullptr;
  KT_CHECK_ARG(!(warm && !r_tld), "warm start needs a separate r_tld plane");
  P2PCtx c;
  KT_TRY(p2p_begin(h, cm, &c));
  KT_TRY(sc_reset(h, fixed_iters ? -1.0 : eps, max_iter, flags & KT_FLAG_OMEGA_ZERO_BREAK));
  const T* r_tld_c = r_tld ? r_tld : (const T*)b;
  Plan P = make_plan(h);
  ReduceWs ws = make_ws(h, nullptr, 0, c.dots);
  if (warm) {
    KT_TRY(p2p_halo<T>(c, 0, (T*)x));                               // halo rows of x0 for r = b - A x0
    StencilArgs<T> A = stencil_args<T>(h, P, coef);
    A.ws = ws; A.x = (const T*)x; A.q = (const T*)b; A.y = r; A.s_out = r_tld; A.fin = FIN_NONE;
    KT_TRY((launch_stencil<T, SRC_X, false, EPI_RESID>(h, P, A)));
  } else {
    KT_TRY(t_dist_phase<T>(h, 0, coef, b, nullptr, nullptr, r, x, r_tld, c.dots));
  }
  KT_TRY(p2p_reduce(c, 6 + (int)(h->p2p_epoch & 1ull), 1, FIN_BICG_INIT, w->history, w->history_cap));
  KT_TRY(p2p_halo<T>(c, 2, r));
  StencilArgs<T> A2 = stencil_args<T>(h, P, coef);
  A2.ws = ws; A2.x = p; A2.minv = (const T*)Minv; A2.q = r_tld_c; A2.y = v; A2.check_done = 1; A2.fin = FIN_NONE;
  StencilArgs<T> A3 = stencil_args<T>(h, P, coef);
  A3.ws = ws; A3.x = r; A3.v = v; A3.minv = (const T*)Minv; A3.y = t; A3.s_out = s; A3.alpha_from_sc = 1; A3.check_done = 1; A3.fin = FIN_NONE;
  auto iter = [&]() -> int {
    {
      ProfScope ps(h, KT_K_BICG_P);
      k_bicg_p<T><<<grid_for(h, P.ew_blocks, 4, OCC_BICG_P), 256, 0, h->stream>>>(P.g, p, r, v, h->sc);
      KT_LAUNCH_CHECK(h);
    }
    KT_TRY(p2p_halo<T>(c, 0, p));
    {
      ProfScope ps(h, KT_K_BICG_V);
      KT_TRY(pre ? (launch_stencil<T, SRC_X, true, EPI_DOT_Q>(h, P, A2)) : (launch_stencil<T, SRC_X, false, EPI_DOT_Q>(h, P, A2)));
    }
    KT_TRY(p2p_reduce(c, 0, 1, FIN_BICG_ALPHA, nullptr, 0));
    KT_TRY(p2p_halo<T>(c, 1, v));
    {
      ProfScope ps(h, KT_K_BICG_ST);
      KT_TRY(pre ? (launch_stencil<T, SRC_S, true, EPI_ST>(h, P, A3)) : (launch_stencil<T, SRC_S, false, EPI_ST>(h, P, A3)));
    }
    KT_TRY(p2p_reduce(c, 1, 2, FIN_BICG_OMEGA, nullptr, 0));
    {
      ProfScope ps(h, KT_K_BICG_XR);
      const T* v_for_s = s ? nullptr : v;
      if (pre) k_bicg_xr<T, true><<<grid_for(h, P.ew_blocks, EWP_BICG_XR_UNIT), 256, 0, h->stream>>>(P.g, (T*)x, r, p, s, t, r_tld_c, (const T*)Minv, 0, 0, 1, ws, FIN_NONE, v_for_s);
      else k_bicg_xr<T, false><<<grid_for(h, P.ew_blocks, EWP_BICG_XR_UNIT), 256, 0, h->stream>>>(P.g, (T*)x, r, p, s, t, r_tld_c, nullptr, 0, 0, 1, ws, FIN_NONE, v_for_s);
      KT_LAUNCH_CHECK(h);
    }
    KT_TRY(p2p_reduce(c, 2, 2, FIN_BICG_END, w->history, w->history_cap));
    KT_TRY(p2p_halo<T>(c, 2, r));
    return 0;
  };
  int rc = run_chunks(h, max_iter, iter);
  if (rc) return rc;
  const Scalars& S = *h->sc_host;
  if (iters) *iters = (long)S.iter;
  if (rel_res) *rel_res = sqrt(S.rTr / S.rTr0);
  if (status) *status = S.status;
  return 0;
}

// CGSolver.solve (cgsolver.py:83-114) over x-slabs with the same two comm kernels: one halo exchange (p) and two
// all-reduces ((p,Ap) and (r,r)) per iteration.
template <typename T>
static int t_p2p_cg_unfused(kt_handle_s* h, const void* coef, const void* b, void* x, const kt_krylov_ws* w, const kt_p2p_comm* cm,
                    double eps, long long max_iter, bool fixed_iters, long* iters, double* rel_res, int* status) {
  T* r = (T*)w->r; T* p = (T*)w->p; T* Ap = (T*)w->Ap;
  KT_CHECK_ARG(r && p && Ap, "kt_krylov_ws: r, p, Ap are required for CG");
  P2PCtx c;
  KT_TRY(p2p_begin(h, cm, &c));
  KT_TRY(sc_reset(h, fixed_iters ? -1.0 : eps, max_iter, 0));
  KT_TRY(t_dist_phase<T>(h, 5, coef, b, nullptr, nullptr, r, x, p, c.dots));
  KT_TRY(p2p_reduce(c, 6 + (int)(h->p2p_epoch & 1ull), 1, FIN_CG_INIT, w->history, w->history_cap));
  KT_TRY(p2p_halo<T>(c, 0, p));
  auto iter = [&]() -> int {
    KT_TRY(t_dist_phase<T>(h, 6, coef, p, nullptr, nullptr, Ap, nullptr, nullptr, c.dots));
    KT_TRY(p2p_reduce(c, 0, 1, FIN_CG_ALPHA, nullptr, 0));
    KT_TRY(t_dist_phase<T>(h, 7, coef, p, Ap, nullptr, x, r, nullptr, c.dots));
    KT_TRY(p2p_reduce(c, 1, 1, FIN_CG_END, w->history, w->history_cap));
    KT_TRY(t_dist_phase<T>(h, 8, coef, r, nullptr, nullptr, p, nullptr, nullptr, c.dots));
    KT_TRY(p2p_halo<T>(c, 0, p));
    return 0;
  };
  int rc = run_chunks(h, max_iter, iter);
  if (rc) return rc;
  const Scalars& S = *h->sc_host;
  if (iters) *iters = (long)S.iter;
  if (rel_res) *rel_res = sqrt(S.rTr / S.rTr0);
  if (status) *status = S.status;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// The same two solvers with the communication FUSED into the compute kernels (default; KT_P2P_FUSED=0 selects the
// separate comm kernels above for comparison).  Per BiCGSTAB iteration the four compute kernels are all that is launched:
//   K1  p-update      stores its first / last owned row of p straight into the neighbours' receive buffers
//   K2  v = A p^      its first thread raises the neighbours' p flags (stream order: every store of K1 is behind it); boundary
//                     tiles wait for the neighbours' p flag and read the halo rows from the receive buffer; stores
//                     the boundary rows of v to the neighbours; its finalising block raises the v flags, then IS the
//                     all-reduce of (r~, v) (p2p_allreduce_warp) and computes alpha
//   K3  t = A s^      waits for the r and v flags, halo rows of s = r - alpha v from the receive buffers; finalising block
//                     all-reduces (t,s), (t,t) -> omega
//   K4  x, r update   stores the boundary rows of the new r to the neighbours; finalising block raises the r flags,
//                     all-reduces (r,r), (r,r~) -> test, rho, beta
// Same grids, same per-thread work, same rank-ordered sums as the unfused path: bit-identical results.  Write-after-read
// safety of a receive row: the neighbour overwrites it one iteration later, after an all-reduce that this rank only joins
// from a kernel that runs (stream order) after the kernel that read the row.
struct P2PRows {      // receive rows and flags of one halo slot, from this rank's point of view
  void* push_lo; void* push_hi;                        // where my first / last owned row goes (neighbours' buffers)
  unsigned long long* raise_lo; unsigned long long* raise_hi;   // the neighbours' flags for them
  const void* in_lo; const void* in_hi;                // where the neighbours' rows land (my buffer)
  const unsigned long long* wait_lo; const unsigned long long* wait_hi;   // my flags for them
};

static P2PRows p2p_rows(const P2PCtx& c, int slot, size_t esize) {
  const kt_handle_s* h = c.h;
  const size_t row = (size_t)h->pitch * esize;
  P2PRows R;
  memset(&R, 0, sizeof(R));
  char* lo = (char*)c.cm->peer_halo_recv[0]; char* hi = (char*)c.cm->peer_halo_recv[1]; char* mine = (char*)c.cm->halo_recv;
  if (lo) { R.push_lo = lo + (size_t)(slot * 2 + 1) * row; R.raise_lo = &c.blk_lo->halo_flag[slot][1];
            R.in_lo = mine + (size_t)(slot * 2 + 0) * row; R.wait_lo = &c.mine->halo_flag[slot][0]; }
  if (hi) { R.push_hi = hi + (size_t)(slot * 2 + 0) * row; R.raise_hi = &c.blk_hi->halo_flag[slot][0];
            R.in_hi = mine + (size_t)(slot * 2 + 1) * row; R.wait_hi = &c.mine->halo_flag[slot][1]; }
  return R;
}

// device copy of the ranks' comm-block pointers (uploaded when they change)
static int p2p_upload_peers(kt_handle_s* h, const P2PCtx& c) {
  if (!h->p2p_peers_dev) KT_CUDA(cudaMalloc(&h->p2p_peers_dev, 16 * sizeof(void*)));
  if (memcmp(h->p2p_peers_host, c.peers.blk, sizeof(h->p2p_peers_host)) != 0) {
    memcpy(h->p2p_peers_host, c.peers.blk, sizeof(h->p2p_peers_host));
    KT_CUDA(cudaMemcpyAsync(h->p2p_peers_dev, h->p2p_peers_host, sizeof(h->p2p_peers_host), cudaMemcpyHostToDevice, h->stream));
  }
  return 0;
}

// ReduceWs of a kernel whose finalising block all-reduces over the ranks (group >= 0) and / or announces pushed rows (R)
static ReduceWs p2p_ws(P2PCtx& c, const kt_krylov_ws* w, int group, const P2PRows* R, int slot) {
  kt_handle_s* h = c.h;
  ReduceWs ws = make_ws(h, w->history, w->history_cap, c.dots);
  ws.timeout_ns = p2p_timeout_ns();
  if (group >= 0 && c.cm->nranks > 1) {
    ws.peers = (P2PBlock* const*)h->p2p_peers_dev; ws.rank = c.cm->rank; ws.nranks = c.cm->nranks; ws.group = group;
    ws.rseq = c.base | ++h->p2p_dot_k[group];
  }
  if (R && c.cm->nranks > 1) { ws.hflag_lo = R->raise_lo; ws.hflag_hi = R->raise_hi; ws.hseq = c.base | ++h->p2p_halo_k[slot]; }
  return ws;
}

template <typename T>
static HaloPush<T> p2p_push(const kt_handle_s* h, const P2PRows& R) {
  HaloPush<T> hp;
  hp.lo = (T*)R.push_lo; hp.hi = (T*)R.push_hi;
  hp.first_off = (long long)h->row_lo * h->pitch; hp.last_off = (long long)(h->row_hi - 1) * h->pitch; hp.pitch = h->pitch;
  return hp;
}

template <typename T>
static int t_p2p_bicgstab_fused(kt_handle_s* h, const void* coef, const void* b, void* x, const void* Minv, const kt_krylov_ws* w,
                                const kt_p2p_comm* cm, double eps, long long max_iter, int flags, bool fixed_iters, long* iters,
                                double* rel_res, int* status) {
  T* r = (T*)w->r; T* p = (T*)w->p; T* v = (T*)w->Ap; T* s = (T*)w->s; T* t = (T*)w->t; T* r_tld = (T*)w->r_tld;
  KT_CHECK_ARG(r && p && v && t, "kt_krylov_ws: r, p, Ap, t are required (s optional: NULL = recomputed, not stored)");
  const bool warm = (flags & KT_FLAG_WARM_START) != 0, pre = Minv != nullptr;
  KT_CHECK_ARG(!(warm && !r_tld), "warm start needs a separate r_tld plane");
  P2PCtx c;
  KT_TRY(p2p_begin(h, cm, &c));
  KT_TRY(p2p_upload_peers(h, c));
  KT_TRY(sc_reset(h, fixed_iters ? -1.0 : eps, max_iter, flags & KT_FLAG_OMEGA_ZERO_BREAK));
  const T* r_tld_c = r_tld ? r_tld : (const T*)b;
  Plan P = make_plan(h);
  const P2PRows Rp = p2p_rows(c, 0, sizeof(T)), Rv = p2p_rows(c, 1, sizeof(T)), Rr = p2p_rows(c, 2, sizeof(T));
  const int g_init = 6 + (int)(h->p2p_epoch & 1ull);
  auto consume = [&](StencilArgs<T>& A, const P2PRows& Rx, unsigned long long seq_x, const P2PRows* Rv_, unsigned long long seq_v) {
    A.hx_lo = (const T*)Rx.in_lo; A.hx_hi = (const T*)Rx.in_hi; A.fx_lo = Rx.wait_lo; A.fx_hi = Rx.wait_hi; A.seq_x = seq_x;
    if (Rv_) { A.hv_lo = (const T*)Rv_->in_lo; A.hv_hi = (const T*)Rv_->in_hi; A.fv_lo = Rv_->wait_lo; A.fv_hi = Rv_->wait_hi; A.seq_v = seq_v; }
  };
  // K0: r (and r~) = b - A x0 or b; (r, r) -> rTr0, rho; the boundary rows of r go to the neighbours (operand of the first K3)
  if (warm) {
    KT_TRY(p2p_halo<T>(c, 0, (T*)x));                               // halo rows of x0 (a plane of the caller's: separate exchange)
    StencilArgs<T> A = stencil_args<T>(h, P, coef);
    A.ws = p2p_ws(c, w, g_init, &Rr, 2);
    A.x = (const T*)x; A.q = (const T*)b; A.y = r; A.s_out = r_tld; A.fin = FIN_BICG_INIT;
    A.push_lo = (T*)Rr.push_lo; A.push_hi = (T*)Rr.push_hi;
    KT_TRY((launch_stencil<T, SRC_X, false, EPI_RESID, true>(h, P, A)));
  } else {
    ProfScope ps(h, KT_K_INIT);
    const ReduceWs ws0 = p2p_ws(c, w, g_init, &Rr, 2);
    k_init<T, false, true><<<grid_for(h, P.ew_blocks, EWP_INIT), 256, 0, h->stream>>>(P.g, (const T*)b, r, r_tld, (T*)x, nullptr, ws0, FIN_BICG_INIT,
                                                                                      p2p_push<T>(h, Rr));
    KT_LAUNCH_CHECK(h);
  }
  StencilArgs<T> A2 = stencil_args<T>(h, P, coef);
  A2.x = p; A2.minv = (const T*)Minv; A2.q = r_tld_c; A2.y = v; A2.check_done = 1; A2.fin = FIN_BICG_ALPHA;
  A2.push_lo = (T*)Rv.push_lo; A2.push_hi = (T*)Rv.push_hi;
  StencilArgs<T> A3 = stencil_args<T>(h, P, coef);
  A3.x = r; A3.v = v; A3.minv = (const T*)Minv; A3.y = t; A3.s_out = s; A3.alpha_from_sc = 1; A3.check_done = 1; A3.fin = FIN_BICG_OMEGA;
  auto iter = [&]() -> int {
    {
      ProfScope ps(h, KT_K_BICG_P);
      ++h->p2p_halo_k[0];                                           // the rows this launch pushes; announced by K2 below
      launch_k(h, k_bicg_p<T, true>, grid_for(h, P.ew_blocks, 4, OCC_BICG_P), P.g, p, r, v, h->sc, p2p_push<T>(h, Rp));
      KT_LAUNCH_CHECK(h);
    }
    {
      ProfScope ps(h, KT_K_BICG_V);
      A2.ws = p2p_ws(c, w, 0, &Rv, 1);
      consume(A2, Rp, c.base | h->p2p_halo_k[0], nullptr, 0);
      A2.raise_lo = Rp.raise_lo; A2.raise_hi = Rp.raise_hi;
      KT_TRY(pre ? (launch_stencil<T, SRC_X, true, EPI_DOT_Q, true>(h, P, A2)) : (launch_stencil<T, SRC_X, false, EPI_DOT_Q, true>(h, P, A2)));
    }
    {
      ProfScope ps(h, KT_K_BICG_ST);
      A3.ws = p2p_ws(c, w, 1, nullptr, 0);
      consume(A3, Rr, c.base | h->p2p_halo_k[2], &Rv, c.base | h->p2p_halo_k[1]);
      KT_TRY(pre ? (launch_stencil<T, SRC_S, true, EPI_ST, true>(h, P, A3)) : (launch_stencil<T, SRC_S, false, EPI_ST, true>(h, P, A3)));
    }
    {
      ProfScope ps(h, KT_K_BICG_XR);
      const T* v_for_s = s ? nullptr : v;
      const ReduceWs ws4 = p2p_ws(c, w, 2, &Rr, 2);
      const HaloPush<T> hp = p2p_push<T>(h, Rr);
      if (pre) launch_k(h, k_bicg_xr<T, true, true>, grid_for(h, P.ew_blocks, EWP_BICG_XR_UNIT), P.g, (T*)x, r, p, s, t, r_tld_c, (const T*)Minv, 0, 0, 1, ws4, FIN_BICG_END, v_for_s, hp);
      else launch_k(h, k_bicg_xr<T, false, true>, grid_for(h, P.ew_blocks, EWP_BICG_XR_UNIT), P.g, (T*)x, r, p, s, t, r_tld_c, nullptr, 0, 0, 1, ws4, FIN_BICG_END, v_for_s, hp);
      KT_LAUNCH_CHECK(h);
    }
    return 0;
  };
  int rc = run_chunks(h, max_iter, iter);
  if (rc) return rc;
  const Scalars& S = *h->sc_host;
  if (iters) *iters = (long)S.iter;
  if (rel_res) *rel_res = sqrt(S.rTr / S.rTr0);
  if (status) *status = S.status;
  return 0;
}

// CG fused the same way: k_init / the p-update push the boundary rows of p, Ap = A p waits for them; the finalising blocks
// of k_init, Ap = A p and the x / r update are the three all-reduces.
template <typename T>
static int t_p2p_cg_fused(kt_handle_s* h, const void* coef, const void* b, void* x, const kt_krylov_ws* w, const kt_p2p_comm* cm,
                          double eps, long long max_iter, bool fixed_iters, long* iters, double* rel_res, int* status) {
  T* r = (T*)w->r; T* p = (T*)w->p; T* Ap = (T*)w->Ap;
  KT_CHECK_ARG(r && p && Ap, "kt_krylov_ws: r, p, Ap are required for CG");
  P2PCtx c;
  KT_TRY(p2p_begin(h, cm, &c));
  KT_TRY(p2p_upload_peers(h, c));
  KT_TRY(sc_reset(h, fixed_iters ? -1.0 : eps, max_iter, 0));
  Plan P = make_plan(h);
  const P2PRows Rp = p2p_rows(c, 0, sizeof(T));
  {
    const ReduceWs ws0 = p2p_ws(c, w, 6 + (int)(h->p2p_epoch & 1ull), &Rp, 0);
    k_init<T, true, true><<<grid_for(h, P.ew_blocks, EWP_INIT), 256, 0, h->stream>>>(P.g, (const T*)b, r, nullptr, (T*)x, p, ws0, FIN_CG_INIT,
                                                                                     p2p_push<T>(h, Rp));
    KT_LAUNCH_CHECK(h);
  }
  StencilArgs<T> A1 = stencil_args<T>(h, P, coef);
  A1.x = p; A1.y = Ap; A1.check_done = 1; A1.fin = FIN_CG_ALPHA;
  auto iter = [&]() -> int {
    {
      ProfScope ps(h, KT_K_CG_AP);
      A1.ws = p2p_ws(c, w, 0, nullptr, 0);
      A1.hx_lo = (const T*)Rp.in_lo; A1.hx_hi = (const T*)Rp.in_hi; A1.fx_lo = Rp.wait_lo; A1.fx_hi = Rp.wait_hi; A1.seq_x = c.base | h->p2p_halo_k[0];
      A1.raise_lo = Rp.raise_lo; A1.raise_hi = Rp.raise_hi;       // rows pushed by k_init / the last p-update (idempotent after k_init)
      KT_TRY((launch_stencil<T, SRC_X, false, EPI_DOT_SRC, true>(h, P, A1)));
    }
    {
      ProfScope ps(h, KT_K_CG_XR);
      const ReduceWs ws2 = p2p_ws(c, w, 1, nullptr, 0);
      launch_k(h, k_cg_xr<T>, grid_for(h, P.ew_blocks, EWP_CG_XR), P.g, (T*)x, r, p, Ap, 0.0, 1, ws2, FIN_CG_END);
      KT_LAUNCH_CHECK(h);
    }
    {
      ProfScope ps(h, KT_K_CG_P);
      ++h->p2p_halo_k[0];                                           // announced by the next Ap = A p
      launch_k(h, k_cg_p<T, true>, grid_for(h, P.ew_blocks, 3, OCC_CG_P), P.g, p, r, h->sc, p2p_push<T>(h, Rp));
      KT_LAUNCH_CHECK(h);
    }
    return 0;
  };
  int rc = run_chunks(h, max_iter, iter);
  if (rc) return rc;
  const Scalars& S = *h->sc_host;
  if (iters) *iters = (long)S.iter;
  if (rel_res) *rel_res = sqrt(S.rTr / S.rTr0);
  if (status) *status = S.status;
  return 0;
}

static bool p2p_fused() {
  static int on = env_int("KT_P2P_FUSED", 1);
  return on != 0;
}

template <typename T>
static int t_p2p_bicgstab(kt_handle_s* h, const void* coef, const void* b, void* x, const void* Minv, const kt_krylov_ws* w,
                          const kt_p2p_comm* cm, double eps, long long max_iter, int flags, bool fixed_iters, long* iters, double* rel_res,
                          int* status) {
  return p2p_fused() ? t_p2p_bicgstab_fused<T>(h, coef, b, x, Minv, w, cm, eps, max_iter, flags, fixed_iters, iters, rel_res, status)
                     : t_p2p_bicgstab_unfused<T>(h, coef, b, x, Minv, w, cm, eps, max_iter, flags, fixed_iters, iters, rel_res, status);
}

template <typename T>
static int t_p2p_cg(kt_handle_s* h, const void* coef, const void* b, void* x, const kt_krylov_ws* w, const kt_p2p_comm* cm, double eps,
                    long long max_iter, bool fixed_iters, long* iters, double* rel_res, int* status) {
  return p2p_fused() ? t_p2p_cg_fused<T>(h, coef, b, x, w, cm, eps, max_iter, fixed_iters, iters, rel_res, status)
                     : t_p2p_cg_unfused<T>(h, coef, b, x, w, cm, eps, max_iter, fixed_iters, iters, rel_res, status);
}
#undef KT_TRY

}  // namespace kt

using namespace kt;

#define KT_DISPATCH(h, call_d, call_f) ((h)->dtype == KT_F64 ? (call_d) : (call_f))

extern "C" {

int kt_spmv(kt_handle_t h, const void* coef, const void* x, void* y) {
  KT_CHECK_ARG(h && coef && x && y, "kt_spmv: null argument");
  KT_DEV_GUARD(h);
  KT_CHECK_ARG(x != y, "kt_spmv: x and y must not alias");
  return KT_DISPATCH(h, t_spmv<double>(h, coef, x, y, nullptr, nullptr), t_spmv<float>(h, coef, x, y, nullptr, nullptr));
}

int kt_spmv_dot(kt_handle_t h, const void* coef, const void* x, void* y, const void* q, double* dot_host) {
  KT_CHECK_ARG(h && coef && x && y && q && dot_host, "kt_spmv_dot: null argument");
  KT_DEV_GUARD(h);
  KT_CHECK_ARG(x != y, "kt_spmv_dot: x and y must not alias");
  return KT_DISPATCH(h, t_spmv<double>(h, coef, x, y, q, dot_host), t_spmv<float>(h, coef, x, y, q, dot_host));
}

int kt_bicg_s_t(kt_handle_t h, const void* coef, const void* r, const void* v, double alpha, void* s, void* t,
                double* dots_host) {
  KT_CHECK_ARG(h && coef && r && v && s && t && dots_host, "kt_bicg_s_t: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, t_bicg_s_t<double>(h, coef, r, v, alpha, s, t, dots_host),
                     t_bicg_s_t<float>(h, coef, r, v, alpha, s, t, dots_host));
}

int kt_dot(kt_handle_t h, const void* p, const void* q, double* out_host) {
  KT_CHECK_ARG(h && p && q && out_host, "kt_dot: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, t_dot<double>(h, p, q, out_host), t_dot<float>(h, p, q, out_host));
}

int kt_copy(kt_handle_t h, void* dst, const void* src) {
  KT_CHECK_ARG(h && dst && src, "kt_copy: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, (t_ew<double, EW_COPY>(h, dst, src, nullptr, nullptr, 0, 0)),
                     (t_ew<float, EW_COPY>(h, dst, src, nullptr, nullptr, 0, 0)));
}
int kt_fill(kt_handle_t h, void* dst, double value) {
  KT_CHECK_ARG(h && dst, "kt_fill: null argument");
  KT_DEV_GUARD(h);
  if (value != 0.0) {
    const dim3 grid((h->NY + 255) / 256, std::max(1, std::min(h->row_hi - h->row_lo, 1024)));
    if (h->dtype == KT_F64) k_fill_cols<double><<<grid, 256, 0, h->stream>>>((double*)dst, h->pitch, h->NY, h->row_lo, h->row_hi, value);
    else k_fill_cols<float><<<grid, 256, 0, h->stream>>>((float*)dst, h->pitch, h->NY, h->row_lo, h->row_hi, (float)value);
    KT_LAUNCH_CHECK(h);
    return 0;
  }
  return KT_DISPATCH(h, (t_ew<double, EW_FILL>(h, dst, nullptr, nullptr, nullptr, value, 0)),
                     (t_ew<float, EW_FILL>(h, dst, nullptr, nullptr, nullptr, value, 0)));
}
int kt_axpy(kt_handle_t h, void* y, double a, const void* x) {
  KT_CHECK_ARG(h && y && x, "kt_axpy: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, (t_ew<double, EW_AXPY>(h, y, x, nullptr, nullptr, a, 0)),
                     (t_ew<float, EW_AXPY>(h, y, x, nullptr, nullptr, a, 0)));
}
int kt_xpay(kt_handle_t h, void* out, const void* x, double a, const void* y) {
  KT_CHECK_ARG(h && out && x && y, "kt_xpay: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, (t_ew<double, EW_XPAY>(h, out, x, y, nullptr, a, 0)),
                     (t_ew<float, EW_XPAY>(h, out, x, y, nullptr, a, 0)));
}
int kt_xmay(kt_handle_t h, void* out, const void* x, double a, const void* y) {
  KT_CHECK_ARG(h && out && x && y, "kt_xmay: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, (t_ew<double, EW_XMAY>(h, out, x, y, nullptr, a, 0)),
                     (t_ew<float, EW_XMAY>(h, out, x, y, nullptr, a, 0)));
}
int kt_bicg_update_p(kt_handle_t h, void* p, const void* r, const void* v, double beta, double omega) {
  KT_CHECK_ARG(h && p && r && v, "kt_bicg_update_p: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, (t_ew<double, EW_BICG_P>(h, p, r, v, nullptr, beta, omega)),
                     (t_ew<float, EW_BICG_P>(h, p, r, v, nullptr, beta, omega)));
}
int kt_bicg_update_x(kt_handle_t h, void* x, const void* p, const void* s, double alpha, double omega) {
  KT_CHECK_ARG(h && x && p && s, "kt_bicg_update_x: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, (t_ew<double, EW_BICG_X>(h, x, p, s, nullptr, alpha, omega)),
                     (t_ew<float, EW_BICG_X>(h, x, p, s, nullptr, alpha, omega)));
}

int kt_bicg_xr(kt_handle_t h, void* x, void* r, const void* p, const void* s, const void* t, const void* r_tld,
               double alpha, double omega, double* dots_host) {
  KT_CHECK_ARG(h && x && r && p && s && t && r_tld && dots_host, "kt_bicg_xr: null argument");
  KT_DEV_GUARD(h);
  Plan P = make_plan(h);
  ReduceWs ws = make_ws(h);
  if (h->dtype == KT_F64)
    k_bicg_xr<double, false><<<grid_for(h, P.ew_blocks, EWP_BICG_XR_UNIT), 256, 0, h->stream>>>(P.g, (double*)x, (double*)r, (const double*)p, (const double*)s,
                                                              (const double*)t, (const double*)r_tld, nullptr, alpha, omega, 0, ws, FIN_NONE);
  else
    k_bicg_xr<float, false><<<grid_for(h, P.ew_blocks, EWP_BICG_XR_UNIT), 256, 0, h->stream>>>(P.g, (float*)x, (float*)r, (const float*)p, (const float*)s,
                                                             (const float*)t, (const float*)r_tld, nullptr, alpha, omega, 0, ws, FIN_NONE);
  KT_LAUNCH_CHECK(h);
  return read_dots(h, dots_host, 2);
}

int kt_cg_xr(kt_handle_t h, void* x, void* r, const void* p, const void* Ap, double alpha, double* dot_host) {
  KT_CHECK_ARG(h && x && r && p && Ap && dot_host, "kt_cg_xr: null argument");
  KT_DEV_GUARD(h);
  Plan P = make_plan(h);
  ReduceWs ws = make_ws(h);
  if (h->dtype == KT_F64)
    k_cg_xr<double><<<grid_for(h, P.ew_blocks, EWP_CG_XR), 256, 0, h->stream>>>(P.g, (double*)x, (double*)r, (const double*)p, (const double*)Ap, alpha, 0, ws, FIN_NONE);
  else
    k_cg_xr<float><<<grid_for(h, P.ew_blocks, EWP_CG_XR), 256, 0, h->stream>>>(P.g, (float*)x, (float*)r, (const float*)p, (const float*)Ap, alpha, 0, ws, FIN_NONE);
  KT_LAUNCH_CHECK(h);
  return read_dots(h, dot_host, 1);
}

int kt_cg_solve(kt_handle_t h, const void* coef, const void* b, void* x, const kt_krylov_ws* ws, double eps, long max_iter,
                long* iters, double* rel_res, int* status) {
  KT_CHECK_ARG(h && coef && b && x && ws, "kt_cg_solve: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, t_cg<double>(h, coef, b, x, ws, eps, max_iter, iters, rel_res, status, false),
                     t_cg<float>(h, coef, b, x, ws, eps, max_iter, iters, rel_res, status, false));
}

int kt_bicgstab_solve(kt_handle_t h, const void* coef, const void* b, void* x, const void* Minv, const kt_krylov_ws* ws,
                      double eps, long max_iter, int flags, long* iters, double* rel_res, int* status) {
  KT_CHECK_ARG(h && coef && b && x && ws, "kt_bicgstab_solve: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, t_bicgstab<double>(h, coef, b, x, Minv, ws, eps, max_iter, flags, iters, rel_res, status, false),
                     t_bicgstab<float>(h, coef, b, x, Minv, ws, eps, max_iter, flags, iters, rel_res, status, false));
}


int kt_dist_begin(kt_handle_t h, double eps, long max_iter, int flags) {
  KT_CHECK_ARG(h, "kt_dist_begin: null handle");
  KT_DEV_GUARD(h);
  return sc_reset(h, eps, max_iter, flags);
}

int kt_dist_bicg_phase(kt_handle_t h, int phase, const void* coef, const void* a0, const void* a1, const void* a2, void* o0,
                       void* o1, void* o2, double* dots_dev) {
  KT_CHECK_ARG(h, "kt_dist_bicg_phase: null handle");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, t_dist_phase<double>(h, phase, coef, a0, a1, a2, o0, o1, o2, dots_dev),
                     t_dist_phase<float>(h, phase, coef, a0, a1, a2, o0, o1, o2, dots_dev));
}

int kt_dist_scalar(kt_handle_t h, int fin, const double* dots_dev, double* history_dev, long history_cap) {
  KT_CHECK_ARG(h && dots_dev, "kt_dist_scalar: null argument");
  KT_DEV_GUARD(h);
  k_scalar_step<<<1, 1, 0, h->stream>>>(h->sc, fin, dots_dev, history_dev, history_cap);
  KT_LAUNCH_CHECK(h);
  return 0;
}

int kt_dist_poll(kt_handle_t h, int* done, long* iters, int* status, double* rel_res) {
  KT_CHECK_ARG(h, "kt_dist_poll: null handle");
  KT_DEV_GUARD(h);
  if (int rc = publish(h, h->sc, h->sc_host, (int)(sizeof(Scalars) / 8))) return rc;
  KT_CUDA(cudaStreamSynchronize(h->stream));
  const Scalars& S = *h->sc_host;
  if (done) *done = S.done;
  if (iters) *iters = (long)S.iter;
  if (status) *status = S.status;
  if (rel_res) *rel_res = sqrt(S.rTr / S.rTr0);
  return 0;
}

int kt_p2p_comm_block_bytes(void) { return (int)sizeof(P2PBlock); }

// a wait gave up: say which flag, what it waited for and what it saw (kt_last_error)
static void p2p_timeout_error(const char* where, const kt_handle_s* h, const kt_p2p_comm* cm) {
  const Scalars& S = *h->sc_host;
  char buf[384];
  const unsigned w = (unsigned)S.wait_what;
  if ((w >> 8 & 0xff) == 1)
    snprintf(buf, sizeof buf, "%s: rank %d of %d gave up waiting for halo slot %u, side %u: expected sequence number %llu (epoch %llu, index %llu), "
             "last saw %llu (epoch %llu, index %llu)", where, cm->rank, cm->nranks, w & 0xff, w >> 16, S.wait_seq, S.wait_seq >> 32,
             S.wait_seq & 0xffffffffull, S.wait_seen, S.wait_seen >> 32, S.wait_seen & 0xffffffffull);
  else   // sums travel as tagged words: tag = 12 bits of the epoch | 20 bits of the index
    snprintf(buf, sizeof buf, "%s: rank %d of %d gave up waiting for sum group %u of peer rank %u: expected epoch %llu, index %llu (tag 0x%llx), "
             "last saw tag 0x%llx (epoch %llu, index %llu)", where, cm->rank, cm->nranks, w & 0xff, w >> 16, S.wait_seq >> 32,
             S.wait_seq & 0xffffffffull, ((S.wait_seq >> 32) & 0xfffull) << 20 | (S.wait_seq & 0xfffffull), S.wait_seen, S.wait_seen >> 20,
             S.wait_seen & 0xfffffull);
  set_error(buf);
}

int kt_p2p_bicgstab_solve(kt_handle_t h, const void* coef, const void* b, void* x, const void* Minv, const kt_krylov_ws* ws,
                          const kt_p2p_comm* comm, double eps, long max_iter, int flags, int fixed_iters, long* iters, double* rel_res,
                          int* status) {
  KT_CHECK_ARG(h && coef && b && x && ws && comm, "kt_p2p_bicgstab_solve: null argument");
  KT_DEV_GUARD(h);
  const int rc = KT_DISPATCH(h, t_p2p_bicgstab<double>(h, coef, b, x, Minv, ws, comm, eps, max_iter, flags, fixed_iters != 0, iters, rel_res, status),
                             t_p2p_bicgstab<float>(h, coef, b, x, Minv, ws, comm, eps, max_iter, flags, fixed_iters != 0, iters, rel_res, status));
  if (rc == 0 && h->sc_host->status == KT_STATUS_COMM_TIMEOUT) p2p_timeout_error("kt_p2p_bicgstab_solve", h, comm);   // still rc 0: reported in *status
  return rc;
}

int kt_p2p_halo_slots(void) { return P2P_SLOTS; }

int kt_p2p_exchange(kt_handle_t h, const kt_p2p_comm* comm, void* plane) {
  KT_CHECK_ARG(h && comm && plane, "kt_p2p_exchange: null argument");
  KT_DEV_GUARD(h);
  return KT_DISPATCH(h, t_p2p_exchange<double>(h, comm, plane), t_p2p_exchange<float>(h, comm, plane));
}

int kt_p2p_allreduce(kt_handle_t h, const kt_p2p_comm* comm, const double* vals_dev2, const double* vals_host2, int op, double* out_host2) {
  KT_CHECK_ARG(h && comm && (vals_dev2 || vals_host2) && out_host2 && (op == 0 || op == 1), "kt_p2p_allreduce: bad argument");
  KT_DEV_GUARD(h);
  P2PCtx c;
  int rc = p2p_begin(h, comm, &c, false);
  if (rc) return rc;
  const unsigned long long seq = ++h->p2p_red_seq;
  const int group = 3 + (int)(seq % 3ull);
  k_p2p_reduce<<<1, 32, 0, h->stream>>>(c.peers, comm->rank, comm->nranks, group, 2, vals_dev2, FIN_NONE, seq, h->sc, nullptr, 0, 0, op, p2p_timeout_ns(),
                                        vals_dev2 ? 0.0 : vals_host2[0], vals_dev2 ? 0.0 : vals_host2[1]);
  KT_LAUNCH_CHECK(h);
  rc = read_dots(h, out_host2, 2);
  if (rc) return rc;
  if (h->sc_host->status == KT_STATUS_COMM_TIMEOUT && h->sc_host->done) { p2p_timeout_error("kt_p2p_allreduce (or an exchange before it)", h, comm); return -5; }
  return 0;
}

int kt_p2p_cg_solve(kt_handle_t h, const void* coef, const void* b, void* x, const kt_krylov_ws* ws, const kt_p2p_comm* comm,
                    double eps, long max_iter, int fixed_iters, long* iters, double* rel_res, int* status) {
  KT_CHECK_ARG(h && coef && b && x && ws && comm, "kt_p2p_cg_solve: null argument");
  KT_DEV_GUARD(h);
  const int rc = KT_DISPATCH(h, t_p2p_cg<double>(h, coef, b, x, ws, comm, eps, max_iter, fixed_iters != 0, iters, rel_res, status),
                             t_p2p_cg<float>(h, coef, b, x, ws, comm, eps, max_iter, fixed_iters != 0, iters, rel_res, status));
  if (rc == 0 && h->sc_host->status == KT_STATUS_COMM_TIMEOUT) p2p_timeout_error("kt_p2p_cg_solve", h, comm);
  return rc;
}

// Launch geometry of the kernel that computes an inner product, for a handle-free geometry: everything the summation
// order depends on (see kt_common.cuh: per-thread sequential -> warp tree -> warps in order -> block partials ->
// finalising block).  Test infrastructure: oracle/device_order.py adds in this order on the CPU.
int kt_reduce_plan(const kt_geom* g, int kind, int pre, long long* out12) {
  KT_CHECK_ARG(g && out12, "kt_reduce_plan: null argument");
  KT_CHECK_ARG(g->NX > 0 && g->pitch >= g->NY && 0 <= g->row_lo && g->row_lo <= g->row_hi && g->row_hi <= g->NX && g->num_sms > 0,
               "kt_reduce_plan: bad geometry");
  kt_handle_s tmp;
  tmp.NX = g->NX; tmp.NY = g->NY; tmp.pitch = g->pitch; tmp.dtype = g->dtype; tmp.row_lo = g->row_lo; tmp.row_hi = g->row_hi;
  tmp.num_sms = g->num_sms; tmp.max_blocks = g->num_sms * 512;
  const Plan P = make_plan(&tmp);
  for (int k = 0; k < 12; ++k) out12[k] = 0;
  int ew_planes = 0, SRC = -1, EPI = -1;
  switch (kind) {
    case KT_PLAN_DOT: ew_planes = EWP_DOT; break;
    case KT_PLAN_INIT: ew_planes = EWP_INIT; break;
    case KT_PLAN_BICG_XR: ew_planes = EWP_BICG_XR; break;
    case KT_PLAN_BICG_XR_UNIT: ew_planes = EWP_BICG_XR_UNIT; break;
    case KT_PLAN_CG_XR: ew_planes = EWP_CG_XR; break;
    case KT_PLAN_BICG_V: SRC = SRC_X; EPI = EPI_DOT_Q; break;
    case KT_PLAN_BICG_ST: SRC = SRC_S; EPI = EPI_ST; break;
    case KT_PLAN_RESID: SRC = SRC_X; EPI = EPI_RESID; pre = 0; break;
    case KT_PLAN_CG_AP: SRC = SRC_X; EPI = EPI_DOT_SRC; pre = 0; break;
    case KT_PLAN_BICG_PV: SRC = SRC_P; EPI = EPI_DOT_Q_ST; break;
    default: set_error("kt_reduce_plan: unknown kind"); return -2;
  }
  out12[2] = 256; out12[3] = P.V;
  if (ew_planes) {
    out12[0] = 0; out12[1] = grid_for(&tmp, P.ew_blocks, ew_planes); out12[4] = P.g.v_lo; out12[5] = P.g.v_hi;
    return 0;
  }
  int W = 1, occ = 0;
  if (SRC == SRC_X) { W = pre ? StencilCfg<SRC_X, true>::W : StencilCfg<SRC_X, false>::W; occ = pre ? StencilCfg<SRC_X, true>::MINB : StencilCfg<SRC_X, false>::MINB; }
  else if (SRC == SRC_S) { W = pre ? StencilCfg<SRC_S, true>::W : StencilCfg<SRC_S, false>::W; occ = pre ? StencilCfg<SRC_S, true>::MINB : StencilCfg<SRC_S, false>::MINB; }
  else { W = pre ? StencilCfg<SRC_P, true>::W : StencilCfg<SRC_P, false>::W; occ = pre ? StencilCfg<SRC_P, true>::MINB : StencilCfg<SRC_P, false>::MINB; }
  const StPlan S = stencil_plan(&tmp, P.V, W, stencil_planes(SRC, pre != 0, EPI), occ);
  out12[0] = 1; out12[1] = S.grid; out12[4] = W; out12[5] = S.TI; out12[6] = S.nCC; out12[7] = S.nTiles;
  out12[8] = g->row_lo; out12[9] = g->row_hi; out12[10] = g->pitch;
  return 0;
}

int kt_resident_plan(const kt_geom* g, int is_cg, int pre, int warm, int enable, long long* out8) {
  KT_CHECK_ARG(g && out8, "kt_resident_plan: null argument");
  kt_handle_s tmp;
  tmp.NX = g->NX; tmp.NY = g->NY; tmp.pitch = g->pitch; tmp.dtype = g->dtype; tmp.row_lo = g->row_lo; tmp.row_hi = g->row_hi;
  tmp.num_sms = g->num_sms; tmp.max_blocks = g->num_sms * 512;
  tmp.allow_cluster = enable == 1 || enable == 3; tmp.allow_coop = enable == 1 || enable == 2;
  for (int k = 0; k < 8; ++k) out8[k] = 0;
  const bool f64 = g->dtype == KT_F64;
  if (f64 ? cluster_plan<double>(&tmp, pre, warm, is_cg, out8) : cluster_plan<float>(&tmp, pre, warm, is_cg, out8)) return 0;
  if (f64 ? coop_plan<double>(&tmp, pre, warm, is_cg, out8) : coop_plan<float>(&tmp, pre, warm, is_cg, out8)) return 0;
  return 0;    // out8[0] == 0: the streaming kernels (kt_reduce_plan)
}

int kt_get_geom(kt_handle_t h, kt_geom* out) {
  KT_CHECK_ARG(h && out, "kt_get_geom: null argument");
  KT_DEV_GUARD(h);
  out->NX = h->NX; out->NY = h->NY; out->pitch = h->pitch; out->row_lo = h->row_lo; out->row_hi = h->row_hi;
  out->dtype = h->dtype; out->num_sms = h->num_sms;
  return 0;
}

int kt_prof_begin(kt_handle_t h, int max_launches) {
  KT_CHECK_ARG(h && max_launches > 0, "kt_prof_begin: bad argument");
  KT_DEV_GUARD(h);
  while ((int)h->prof_ev.size() < 2 * max_launches) {
    cudaEvent_t e;
    KT_CUDA(cudaEventCreate(&e));
    h->prof_ev.push_back(e);
  }
  h->prof_cls.assign(h->prof_ev.size() / 2, -1);
  h->prof_used = 0;
  h->prof_on = true;
  return 0;
}

int kt_prof_end(kt_handle_t h, double* total_ms, long* count, int n_classes) {
  KT_CHECK_ARG(h && total_ms && count && n_classes > 0, "kt_prof_end: bad argument");
  KT_DEV_GUARD(h);
  h->prof_on = false;
  KT_CUDA(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < n_classes; ++c) { total_ms[c] = 0.0; count[c] = 0; }
  for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
    float ms = 0.f;
    KT_CUDA(cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]));
    const int c = h->prof_cls[i / 2];
    if (c >= 0 && c < n_classes) { total_ms[c] += ms; count[c] += 1; }
  }
  h->prof_used = 0;
  return 0;
}

int kt_bicgstab_bench(kt_handle_t h, const void* coef, const void* b, void* x, const kt_krylov_ws* ws, long n_iter, float* ms) {
  KT_CHECK_ARG(h && coef && b && x && ws && ms, "kt_bicgstab_bench: null argument");
  KT_DEV_GUARD(h);
  KT_CUDA(cudaEventRecord(h->t0, h->stream));
  long it = 0; int st = 0; double rr = 0;
  int rc = KT_DISPATCH(h, t_bicgstab<double>(h, coef, b, x, nullptr, ws, 0.0, n_iter, 0, &it, &rr, &st, true),
                       t_bicgstab<float>(h, coef, b, x, nullptr, ws, 0.0, n_iter, 0, &it, &rr, &st, true));
  if (rc) return rc;
  KT_CUDA(cudaEventRecord(h->t1, h->stream));
  KT_CUDA(cudaEventSynchronize(h->t1));
  KT_CUDA(cudaEventElapsedTime(ms, h->t0, h->t1));
  if (it != n_iter) { set_error("kt_bicgstab_bench: solver stopped early (status " + std::to_string(st) + ")"); return -3; }
  return 0;
}

int kt_cg_bench(kt_handle_t h, const void* coef, const void* b, void* x, const kt_krylov_ws* ws, long n_iter, float* ms) {
  KT_CHECK_ARG(h && coef && b && x && ws && ms, "kt_cg_bench: null argument");
  KT_DEV_GUARD(h);
  KT_CUDA(cudaEventRecord(h->t0, h->stream));
  long it = 0; int st = 0; double rr = 0;
  int rc = KT_DISPATCH(h, t_cg<double>(h, coef, b, x, ws, 0.0, n_iter, &it, &rr, &st, true),
                       t_cg<float>(h, coef, b, x, ws, 0.0, n_iter, &it, &rr, &st, true));
  if (rc) return rc;
  KT_CUDA(cudaEventRecord(h->t1, h->stream));
  KT_CUDA(cudaEventSynchronize(h->t1));
  KT_CUDA(cudaEventElapsedTime(ms, h->t0, h->t1));
  if (it != n_iter) { set_error("kt_cg_bench: solver stopped early (status " + std::to_string(st) + ")"); return -3; }
  return 0;
}

}  // extern "C"

// every kernel of this translation unit, for kt_preload_kernels (kt_handle.h)
namespace kt {
int preload_kt_krylov() {
  static const void* const fns[] = {
#define KT_PRELOAD_TU_kt_krylov
#include "kt_preload_list.inc"
#undef KT_PRELOAD_TU_kt_krylov
      nullptr};
  return preload_list(fns);
}
}  // namespace kt
