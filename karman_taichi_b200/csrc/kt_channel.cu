// kt_channel.cu -- channel + square-cell cylinder (the monolith karman_taichi.py) for sm_100a, in stencil form.
//
// The reference assembles n x n pointer-SNode matrices Au/Av/Ap (karman_taichi.py:162-339) whose every entry lies on
// the offsets 0, +-1, +-stride; here each system is the class version's five coefficient planes
// (centre, i-1, i+1, j+1, j-1) plus a Jacobi plane M (and 1/M for the solver), assembled by one thread per unknown that
// walks the reference's first-match if/elif chains.  Reference quirks that change the numbers are kept (SURVEY.md
// A.4): the second u-loop next to the obstacle that neither clears the stale j-1 entry nor refreshes Mu (:221-239),
// outlet pressure rows whose diagonal becomes 1 while the off-diagonals stay (:327-339), cell (nx,1) left un-pinned.
// Every formula keeps the reference's operation order (library built -fmad=false) => bit-identical to the oracle.
#include <algorithm>
#include "kt_handle.h"

namespace kt {

struct CP {   // device copy of kt_channel_params
  int nx, ny;
  double dx, dy, dt, rho, mu, velo_rel, p_rel;
  int cx, cy, r2;
  int pu, pv, pp;
  size_t plu, plv, plp;
  // x-slab window: the planes hold the GLOBAL rows [i0, i0 + local rows); a launch works on the global rows [ia, ib).
  // One GPU: i0 = 0, [ia, ib) = [0, nx + 3).  Every index test below is on the global row; memory is addressed by i - i0.
  int i0, ia, ib;
};

static CP make_cp(const kt_channel_params* P) {
  CP c;
  c.nx = P->nx; c.ny = P->ny; c.dx = P->dx; c.dy = P->dy; c.dt = P->dt; c.rho = P->rho; c.mu = P->mu;
  c.velo_rel = P->velo_rel; c.p_rel = P->p_rel; c.cx = P->cyl_cx; c.cy = P->cyl_cy; c.r2 = P->cyl_r2;
  c.pu = P->pitch_u; c.pv = P->pitch_v; c.pp = P->pitch_p;
  const bool slab = P->row_hi > 0;
  const int rows_u = slab ? P->rows_u : P->nx + 3, rows_p = slab ? P->rows_p : P->nx + 2;
  c.plu = (size_t)rows_u * P->pitch_u; c.plv = (size_t)rows_p * P->pitch_v; c.plp = (size_t)rows_p * P->pitch_p;
  c.i0 = slab ? P->row0 : 0; c.ia = slab ? P->row_lo : 0; c.ib = slab ? P->row_hi : P->nx + 3;
  return c;
}

// rows of a launch: blockIdx.y strides over [ia, ib) (gridDim.y <= 65535 whatever the grid height)
#define KT_CH_ROWS(i) for (int i = S.ia + (int)blockIdx.y; i < S.ib; i += (int)gridDim.y)
#define RU(i) ((size_t)((i) - S.i0) * S.pu)
#define RV(i) ((size_t)((i) - S.i0) * S.pv)
#define RP(i) ((size_t)((i) - S.i0) * S.pp)
#define CU(i, j) u[RU(i) + (j)]
#define CV(i, j) v[RV(i) + (j)]
#define CPp(i, j) p[RP(i) + (j)]
#define CT(i, j) ct[RP(i) + (j)]

// init (karman_taichi.py:130-154): p = 100 - 4000 i / nx, u = v = u0 = v0 = 0, ct ring / fluid / cylinder
template <typename T>
__global__ void __launch_bounds__(256) k_ch_init(CP S, T* p, T* u, T* u0, T* v, T* v0, int* ct) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  KT_CH_ROWS(i) {
    if (i <= S.nx + 2 && j <= S.ny + 1) { CU(i, j) = T(0); u0[RU(i) + j] = T(0); }
    if (i <= S.nx + 1 && j <= S.ny + 2) { CV(i, j) = T(0); v0[RV(i) + j] = T(0); }
    if (i <= S.nx + 1 && j <= S.ny + 1) {
      CPp(i, j) = T(100) - T(4000.0) * T(i) / T(S.nx);
      int c = (i >= 1 && i <= S.nx && j >= 1 && j <= S.ny) ? -1 : 1;
      if (i < S.nx && j < S.ny && ((i - S.cx) * (i - S.cx) + (j - S.cy) * (j - S.cy)) < S.r2) c = 1;
      CT(i, j) = c;
    }
  }
}

// fill_Au (z == 0, :161-248) and fill_Av (z == 1, :250-294)
template <typename T>
__global__ void __launch_bounds__(256) k_ch_assemble_uv(CP S, const int* __restrict__ ct, const T* __restrict__ u,
                                                        const T* __restrict__ v, const T* __restrict__ p,
                                                        const T* __restrict__ u0, const T* __restrict__ v0, T* cu, T* bu,
                                                        T* mu_, T* mui, T* cv, T* bv, T* mv_, T* mvi) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const T dx = T(S.dx), dy = T(S.dy), dt = T(S.dt), rho = T(S.rho), mu = T(S.mu);
  const T half = T(0.5), two = T(2), one = T(1), zero = T(0);
  const T tt = rho * dx * dy / dt;
  KT_CH_ROWS(i) {
  if (blockIdx.z == 0) {
    if (i < 1 || i > S.nx + 1 || j < 1 || j > S.ny) continue;
    const int ctl = CT(i - 1, j), ctr = CT(i, j), ctru = CT(i, j - 1), ctrd = CT(i, j + 1), ctlu = CT(i - 1, j - 1);
    const T aN = -mu * dx / dy - rho * half * (CV(i - 1, j + 1) + CV(i, j + 1)) * dx;
    const T aS = -mu * dx / dy - (-rho * half * (CV(i - 1, j) + CV(i, j)) * dx);
    const T aW = -mu * dy / dx - rho * half * (CU(i, j) + CU(i - 1, j)) * dy;
    const T aE = -mu * dy / dx - (-rho * half * (CU(i, j) + CU(i + 1, j)) * dy);
    const size_t o = RU(i) + j;
    const T bb = (CPp(i - 1, j) - CPp(i, j)) * dy + rho * dx * dy / dt * u0[o];
    const T C_c = -aN - aW - aE + tt + two * mu * dx / dy;
    T C, W = zero, E = zero, N = zero, Sx = zero, B = zero;
    if (ctl == 1) { C = one; E = -one; }
    else if (ctr == 1) { C = one; W = -one; }
    else if (ctr + ctru == 0) { N = aN; W = aW; E = aE; C = C_c; B = bb; }
    else if (ctr + ctrd == 0) { Sx = aS; W = aW; E = aE; C = -aS - aW - aE + tt + two * mu * dx / dy; B = bb; }
    else { Sx = aS; N = aN; W = aW; E = aE; C = -aS - aN - aW - aE + tt; B = bb; }
    const T M = C;                                   // Mu[k,k] = Au[k,k] after the first loop only (:218)
    if (i >= 2 && i <= S.nx && j >= 2 && j <= S.ny - 1) {      // second loop (:221-248)
      const int s2 = ctl + ctr;
      if (s2 == 0 || s2 == 2) { W = E = N = Sx = zero; C = one; B = zero; }
      else if (ctr + ctru == 0 || ctl + ctlu == 0) { N = aN; W = aW; E = aE; C = C_c; B = bb; }   // S keeps loop 1's value
    }
    cu[o] = C; cu[S.plu + o] = W; cu[2 * S.plu + o] = E; cu[3 * S.plu + o] = N; cu[4 * S.plu + o] = Sx;
    bu[o] = B; mu_[o] = M; mui[o] = one / M;
  } else {
    if (i < 1 || i > S.nx || j < 1 || j > S.ny + 1) continue;
    const int ctc = CT(i, j), ctu = CT(i, j - 1), ctw = CT(i - 1, j), cte = CT(i + 1, j);
    const T aS = -mu * dx / dy - (-rho * half * (CV(i, j - 1) + CV(i, j)) * dx);
    const T aN = -mu * dx / dy - (rho * half * (CV(i, j + 1) + CV(i, j)) * dx);
    const T aE = -mu * dy / dx - (-rho * half * (CU(i + 1, j - 1) + CU(i + 1, j)) * dy);
    const T aW = -mu * dy / dx - (rho * half * (CU(i, j) + CU(i, j - 1)) * dy);
    const size_t o = RV(i) + j;
    const T bb = (CPp(i, j) - CPp(i, j - 1)) * dx + rho * dx * dy / dt * v0[o];
    T C, W = zero, E = zero, N = zero, Sx = zero, B;
    const int s = ctc + ctu;
    if (s == 0 || s == 2) { C = one; B = CV(i, j); }
    else if (ctc + ctw == 0) { Sx = aS; N = aN; E = aE; C = -aS - aN - aE + tt + two * mu * dy / dx; B = bb; }
    else if (ctc + cte == 0) { Sx = aS; N = aN; W = aW; C = -aS - aN - aW + tt + two * mu * dy / dx; B = bb; }
    else { Sx = aS; N = aN; W = aW; E = aE; C = -aS - aN - aW - aE + tt; B = bb; }
    cv[o] = C; cv[S.plv + o] = W; cv[2 * S.plv + o] = E; cv[3 * S.plv + o] = N; cv[4 * S.plv + o] = Sx;
    bv[o] = B; mv_[o] = C; mvi[o] = one / C;
  }
  }
}

// fill_Ap (:296-339)
template <typename T>
__global__ void __launch_bounds__(256) k_ch_assemble_p(CP S, const int* __restrict__ ct, const T* __restrict__ u,
                                                       const T* __restrict__ v, const T* __restrict__ cu,
                                                       const T* __restrict__ cv, T* cp, T* bp, T* mp, T* mpi) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const T dx = T(S.dx), dy = T(S.dy), rho = T(S.rho), zero = T(0), one = T(1);
  KT_CH_ROWS(i) {
  if (i < 1 || i > S.nx || j < 1 || j > S.ny) continue;
  T B = rho * (CU(i, j) - CU(i + 1, j)) * dy + rho * (CV(i, j + 1) - CV(i, j)) * dx;
  T Sx = -rho * dx * dx / cv[RV(i) + j];
  T N = -rho * dx * dx / cv[RV(i) + j + 1];
  T W = -rho * dy * dy / cu[RU(i) + j];
  T E = -rho * dy * dy / cu[RU(i + 1) + j];
  const int ctc = CT(i, j);
  if (ctc + CT(i, j - 1) == 0) Sx = zero;
  if (ctc + CT(i, j + 1) == 0) N = zero;
  if (ctc + CT(i - 1, j) == 0) W = zero;
  if (ctc + CT(i + 1, j) == 0) E = zero;
  // writes that fall outside the reference's n x n matrix are dropped and read back as 0
  if (i == 1 && j == 1) Sx = zero;
  if (i == 1) W = zero;
  if (i == S.nx) E = zero;
  T C = -Sx - N - W - E;
  T M = C;
  if (i == 1) { C = one; M = one; B = zero; Sx = N = W = E = zero; }          // inlet rows: identity (:327-333)
  if (i == S.nx && j >= 2) { C = one; M = one; B = zero; }                    // outlet rows n-1 .. n-ny+1: diagonal only
  const size_t o = RP(i) + j;
  cp[o] = C; cp[S.plp + o] = W; cp[2 * S.plp + o] = E; cp[3 * S.plp + o] = N; cp[4 * S.plp + o] = Sx;
  bp[o] = B; mp[o] = M; mpi[o] = one / M;
  }
}

// xu_back / xv_back (:458-470): u += velo_rel (xu - u) on every unknown
template <typename T>
__global__ void __launch_bounds__(256) k_ch_relax(CP S, T* u, T* v, const T* __restrict__ xu, const T* __restrict__ xv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const T rel = T(S.velo_rel);
  KT_CH_ROWS(i) {
    if (i >= 1 && i <= S.nx + 1 && j >= 1 && j <= S.ny) {
      const size_t o = RU(i) + j;
      u[o] = u[o] + rel * (xu[o] - u[o]);
    }
    if (i >= 1 && i <= S.nx && j >= 1 && j <= S.ny + 1) {
      const size_t o = RV(i) + j;
      v[o] = v[o] + rel * (xv[o] - v[o]);
    }
  }
}

// correct_conserv, first part (:586-600): scale the outlet column by sum(u[1,:]) / sum(u[nx+1,:]).  One block; fixed-order
// (deterministic) sums.  On x-slabs the inlet column belongs to the first rank and the outlet column to the last: every rank
// publishes (its inlet sum or 0, its outlet sum or 0), the ranks' pairs are added in rank order (x + 0 == x: the same bits
// as on one GPU) and the last rank scales with the quotient.
template <typename T>
__global__ void __launch_bounds__(256) k_ch_outlet_sums(CP S, const T* u, double* out) {
  __shared__ double sm[2][8];
  const bool has_in = S.ia <= 1 && 1 < S.ib, has_out = S.ia <= S.nx + 1 && S.nx + 1 < S.ib;
  double a = 0.0, b = 0.0;
  for (int j = 1 + threadIdx.x; j <= S.ny; j += blockDim.x) {
    if (has_in) a += (double)CU(1, j);
    if (has_out) b += (double)CU(S.nx + 1, j);
  }
  a = warp_sum(a); b = warp_sum(b);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sm[0][wid] = a; sm[1][wid] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < 8; ++w) { sa += sm[0][w]; sb += sm[1][w]; }
    out[0] = sa; out[1] = sb; out[2] = sa / sb;
  }
}

// sums: device {inlet sum, outlet sum, ratio} (one GPU) -- or the ratio passed by value after the ranks' sums were added
template <typename T>
__global__ void __launch_bounds__(256) k_ch_outlet_apply(CP S, T* u, const double* sums, double ratio_imm) {
  if (!(S.ia <= S.nx + 1 && S.nx + 1 < S.ib)) return;
  const T coef = T(sums ? sums[2] : ratio_imm);
  for (int j = 1 + threadIdx.x; j <= S.ny; j += blockDim.x) CU(S.nx + 1, j) = coef * CU(S.nx + 1, j);
}

// correct_conserv, second part (:606-612): udiv
template <typename T>
__global__ void __launch_bounds__(256) k_ch_udiv(CP S, const T* __restrict__ u, const T* __restrict__ v, T* udiv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  KT_CH_ROWS(i) {
    if (i < 1 || i > S.nx || j < 1 || j > S.ny) continue;
    udiv[RP(i) + j] = (CU(i, j) - CU(i + 1, j)) * T(S.dy) + (CV(i, j + 1) - CV(i, j)) * T(S.dx);
  }
}

// puv_correction (:506-542): u, v corrections on non-interface faces, p += p_rel pcor on non-solid cells, max |pcor|
template <typename T>
__global__ void __launch_bounds__(256) k_ch_correct(CP S, const int* __restrict__ ct, T* u, T* v, T* p,
                                                    const T* __restrict__ pc, const T* __restrict__ cu,
                                                    const T* __restrict__ cv, double* partials, int max_blocks,
                                                    unsigned int* ticket, double* out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  double mx = 0.0;
  KT_CH_ROWS(i) {
    if (i > S.nx + 1) continue;
    if (i >= 1 && j >= 1 && j <= S.ny) {                          // u faces, i in [1, nx+1]
      const int s = CT(i - 1, j) + CT(i, j);
      if (!(s == 0 || s == 2)) {
        const size_t o = RU(i) + j;
        const T ucor = (pc[RP(i - 1) + j] - pc[RP(i) + j]) * T(S.dy) / cu[o];
        u[o] = u[o] + ucor;
      }
    }
    if (i >= 1 && i <= S.nx && j >= 1 && j <= S.ny + 1) {         // v faces, j in [1, ny+1]
      const int s = CT(i, j) + CT(i, j - 1);
      if (!(s == 0 || s == 2)) {
        const size_t o = RV(i) + j;
        const T vcor = (pc[RP(i) + j] - pc[RP(i) + j - 1]) * T(S.dx) / cv[o];
        v[o] = v[o] + vcor;
      }
    }
    if (j <= S.ny + 1) {                                          // all cells of the (nx+2, ny+2) plane
      const size_t o = RP(i) + j;
      const T pcv = pc[o];
      if (CT(i, j) != 1) p[o] = p[o] + T(S.p_rel) * pcv;
      mx = fmax(mx, (double)fabs(pcv));
    }
  }
  block_max_finalize(mx, 0.0, partials, max_blocks, ticket, out);
}

static unsigned ch_rows(const CP& S) { return (unsigned)std::max(1, std::min(S.ib - S.ia, 65535)); }

static int ch_read(kt_handle_s* h, double* out, int n) {
  if (int rc = publish(h, h->red_out, h->red_host, 4)) return rc;
  KT_CUDA(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < n; ++k) out[k] = h->red_host[k];
  return 0;
}

}  // namespace kt

using namespace kt;

#define KT_CH_ARGS_OK(P) KT_CHECK_ARG((P) && (P)->nx >= 4 && (P)->ny >= 4 && (P)->pitch_u >= (P)->ny + 2 && \
                                      (P)->pitch_v >= (P)->ny + 3 && (P)->pitch_p >= (P)->ny + 2, "kt_channel_params: bad geometry")
#define KT_CH_LAUNCH(kernel, grid, block, ...)                                        \
  do {                                                                                \
    if (h->dtype == KT_F64) kernel<double><<<grid, block, 0, h->stream>>>(__VA_ARGS__); \
    else kernel<float><<<grid, block, 0, h->stream>>>(__VA_ARGS__);                     \
    KT_LAUNCH_CHECK(h);                                                               \
  } while (0)

extern "C" {

int kt_channel_init(kt_handle_t h, const kt_channel_params* P, void* p, void* u, void* u0, void* v, void* v0, int* ct) {
  KT_CHECK_ARG(h && p && u && u0 && v && v0 && ct, "kt_channel_init: null argument");
  KT_DEV_GUARD(h);
  KT_CH_ARGS_OK(P);
  CP S = make_cp(P);
  dim3 grid((P->ny + 3 + 255) / 256, ch_rows(S));
  if (h->dtype == KT_F64) k_ch_init<double><<<grid, 256, 0, h->stream>>>(S, (double*)p, (double*)u, (double*)u0, (double*)v, (double*)v0, ct);
  else k_ch_init<float><<<grid, 256, 0, h->stream>>>(S, (float*)p, (float*)u, (float*)u0, (float*)v, (float*)v0, ct);
  KT_LAUNCH_CHECK(h);
  return 0;
}

int kt_channel_assemble_uv(kt_handle_t h, const kt_channel_params* P, const int* ct, const void* u, const void* v, const void* p,
                           const void* u0, const void* v0, void* coef_u, void* b_u, void* M_u, void* Minv_u, void* coef_v,
                           void* b_v, void* M_v, void* Minv_v) {
  KT_CHECK_ARG(h && ct && u && v && p && u0 && v0 && coef_u && b_u && M_u && Minv_u && coef_v && b_v && M_v && Minv_v,
               "kt_channel_assemble_uv: null argument");
  KT_DEV_GUARD(h);
  KT_CH_ARGS_OK(P);
  CP S = make_cp(P);
  dim3 grid((P->ny + 2 + 255) / 256, ch_rows(S), 2);
  if (h->dtype == KT_F64)
    k_ch_assemble_uv<double><<<grid, 256, 0, h->stream>>>(S, ct, (const double*)u, (const double*)v, (const double*)p, (const double*)u0,
        (const double*)v0, (double*)coef_u, (double*)b_u, (double*)M_u, (double*)Minv_u, (double*)coef_v, (double*)b_v, (double*)M_v, (double*)Minv_v);
  else
    k_ch_assemble_uv<float><<<grid, 256, 0, h->stream>>>(S, ct, (const float*)u, (const float*)v, (const float*)p, (const float*)u0,
        (const float*)v0, (float*)coef_u, (float*)b_u, (float*)M_u, (float*)Minv_u, (float*)coef_v, (float*)b_v, (float*)M_v, (float*)Minv_v);
  KT_LAUNCH_CHECK(h);
  return 0;
}

int kt_channel_assemble_p(kt_handle_t h, const kt_channel_params* P, const int* ct, const void* u, const void* v, const void* coef_u,
                          const void* coef_v, void* coef_p, void* b_p, void* M_p, void* Minv_p) {
  KT_CHECK_ARG(h && ct && u && v && coef_u && coef_v && coef_p && b_p && M_p && Minv_p, "kt_channel_assemble_p: null argument");
  KT_DEV_GUARD(h);
  KT_CH_ARGS_OK(P);
  CP S = make_cp(P);
  dim3 grid((P->ny + 2 + 255) / 256, ch_rows(S));
  if (h->dtype == KT_F64)
    k_ch_assemble_p<double><<<grid, 256, 0, h->stream>>>(S, ct, (const double*)u, (const double*)v, (const double*)coef_u, (const double*)coef_v,
        (double*)coef_p, (double*)b_p, (double*)M_p, (double*)Minv_p);
  else
    k_ch_assemble_p<float><<<grid, 256, 0, h->stream>>>(S, ct, (const float*)u, (const float*)v, (const float*)coef_u, (const float*)coef_v,
        (float*)coef_p, (float*)b_p, (float*)M_p, (float*)Minv_p);
  KT_LAUNCH_CHECK(h);
  return 0;
}

int kt_channel_relax(kt_handle_t h, const kt_channel_params* P, void* u, void* v, const void* x_u, const void* x_v) {
  KT_CHECK_ARG(h && u && v && x_u && x_v, "kt_channel_relax: null argument");
  KT_DEV_GUARD(h);
  KT_CH_ARGS_OK(P);
  CP S = make_cp(P);
  dim3 grid((P->ny + 3 + 255) / 256, ch_rows(S));
  if (h->dtype == KT_F64) k_ch_relax<double><<<grid, 256, 0, h->stream>>>(S, (double*)u, (double*)v, (const double*)x_u, (const double*)x_v);
  else k_ch_relax<float><<<grid, 256, 0, h->stream>>>(S, (float*)u, (float*)v, (const float*)x_u, (const float*)x_v);
  KT_LAUNCH_CHECK(h);
  return 0;
}

static int outlet_udiv(kt_handle_s* h, const kt_channel_params* P, const CP& S, const void* u, const void* v, void* udiv) {
  dim3 grid((P->ny + 2 + 255) / 256, ch_rows(S));
  if (h->dtype == KT_F64) k_ch_udiv<double><<<grid, 256, 0, h->stream>>>(S, (const double*)u, (const double*)v, (double*)udiv);
  else k_ch_udiv<float><<<grid, 256, 0, h->stream>>>(S, (const float*)u, (const float*)v, (float*)udiv);
  KT_LAUNCH_CHECK(h);
  return 0;
}

int kt_outlet_scale(kt_handle_t h, const kt_channel_params* P, void* u, const void* v, void* udiv, double* sums_host) {
  KT_CHECK_ARG(h && u && v, "kt_outlet_scale: null argument");
  KT_DEV_GUARD(h);
  KT_CH_ARGS_OK(P);
  CP S = make_cp(P);
  if (h->dtype == KT_F64) {
    k_ch_outlet_sums<double><<<1, 256, 0, h->stream>>>(S, (const double*)u, h->red_out);
    KT_LAUNCH_CHECK(h);
    k_ch_outlet_apply<double><<<1, 256, 0, h->stream>>>(S, (double*)u, h->red_out, 0.0);
  } else {
    k_ch_outlet_sums<float><<<1, 256, 0, h->stream>>>(S, (const float*)u, h->red_out);
    KT_LAUNCH_CHECK(h);
    k_ch_outlet_apply<float><<<1, 256, 0, h->stream>>>(S, (float*)u, h->red_out, 0.0);
  }
  KT_LAUNCH_CHECK(h);
  if (udiv) { int rc = outlet_udiv(h, P, S, u, v, udiv); if (rc) return rc; }
  if (sums_host) return ch_read(h, sums_host, 3);
  return 0;
}

// x-slab form of correct_conserv: (1) this rank's column sums -> sums_dev[0..1] (0 where the column is not owned); the
// caller adds the ranks' pairs in rank order (kt_p2p_allreduce); (2) the last rank scales its outlet column with
// inlet / outlet and every rank recomputes udiv on its rows (needs the halo row of u above).
int kt_outlet_sums(kt_handle_t h, const kt_channel_params* P, const void* u, double* sums_dev) {
  KT_CHECK_ARG(h && u && sums_dev, "kt_outlet_sums: null argument");
  KT_DEV_GUARD(h);
  KT_CH_ARGS_OK(P);
  CP S = make_cp(P);
  if (h->dtype == KT_F64) k_ch_outlet_sums<double><<<1, 256, 0, h->stream>>>(S, (const double*)u, h->red_out);
  else k_ch_outlet_sums<float><<<1, 256, 0, h->stream>>>(S, (const float*)u, h->red_out);
  KT_LAUNCH_CHECK(h);
  if (int rc = publish(h, h->red_out, sums_dev, 2)) return rc;
  return 0;
}

int kt_outlet_apply(kt_handle_t h, const kt_channel_params* P, void* u, double inlet_sum, double outlet_sum) {
  KT_CHECK_ARG(h && u, "kt_outlet_apply: null argument");
  KT_DEV_GUARD(h);
  KT_CH_ARGS_OK(P);
  CP S = make_cp(P);
  const double ratio = inlet_sum / outlet_sum;        // karman_taichi.py:598 (the same IEEE division the device performs)
  if (h->dtype == KT_F64) k_ch_outlet_apply<double><<<1, 256, 0, h->stream>>>(S, (double*)u, nullptr, ratio);
  else k_ch_outlet_apply<float><<<1, 256, 0, h->stream>>>(S, (float*)u, nullptr, ratio);
  KT_LAUNCH_CHECK(h);
  return 0;
}

int kt_channel_udiv(kt_handle_t h, const kt_channel_params* P, const void* u, const void* v, void* udiv) {
  KT_CHECK_ARG(h && u && v && udiv, "kt_channel_udiv: null argument");
  KT_DEV_GUARD(h);
  KT_CH_ARGS_OK(P);
  CP S = make_cp(P);
  return outlet_udiv(h, P, S, u, v, udiv);
}

int kt_channel_correct(kt_handle_t h, const kt_channel_params* P, const int* ct, void* u, void* v, void* p, const void* pcor,
                       const void* coef_u, const void* coef_v, double* pcor_max_host) {
  KT_CHECK_ARG(h && ct && u && v && p && pcor && coef_u && coef_v && pcor_max_host, "kt_channel_correct: null argument");
  KT_DEV_GUARD(h);
  KT_CH_ARGS_OK(P);
  CP S = make_cp(P);
  const int gx = (P->ny + 3 + 255) / 256;
  int gy = h->max_blocks / gx; if (gy > S.ib - S.ia) gy = S.ib - S.ia; if (gy > 65535) gy = 65535; if (gy < 1) gy = 1;
  dim3 grid(gx, gy);
  if (h->dtype == KT_F64)
    k_ch_correct<double><<<grid, 256, 0, h->stream>>>(S, ct, (double*)u, (double*)v, (double*)p, (const double*)pcor, (const double*)coef_u,
        (const double*)coef_v, h->partials, h->max_blocks, h->ticket, h->red_out);
  else
    k_ch_correct<float><<<grid, 256, 0, h->stream>>>(S, ct, (float*)u, (float*)v, (float*)p, (const float*)pcor, (const float*)coef_u,
        (const float*)coef_v, h->partials, h->max_blocks, h->ticket, h->red_out);
  KT_LAUNCH_CHECK(h);
  return ch_read(h, pcor_max_host, 1);
}

}  // extern "C"

// every kernel of this translation unit, for kt_preload_kernels (kt_handle.h)
namespace kt {
int preload_kt_channel() {
  static const void* const fns[] = {
#define KT_PRELOAD_TU_kt_channel
#include "kt_preload_list.inc"
#undef KT_PRELOAD_TU_kt_channel
      nullptr};
  return preload_list(fns);
}
}  // namespace kt
