// kt_simple.cu -- SIMPLE (lid-driven cavity) assembly / boundary / correction kernels for sm_100a.
//
// Replaces the @ti.kernel methods of SIMPLESolver (bicgsimp.py:79-256, simple.py variant flags).
// Every formula keeps the reference's left-to-right operation order and the library is compiled with
// -fmad=false, so coef_*/b_*/u/v/p are bit-identical to the NumPy oracle for identical inputs.
//
// Layout: u (nx+3, ny+2), v (nx+2, ny+3), p-type (nx+2, ny+2) pitched planes; coef_* five SoA planes.
// Threads map to j (contiguous); these kernels run once per SIMPLE outer iteration (<0.5 % of the
// traffic of the Krylov iterations they feed) and are written as straightforward coalesced passes.
#include <algorithm>
#include "kt_handle.h"

namespace kt {

struct SP {   // device copy of kt_simple_params + derived plane strides
  int nx, ny;
  double dx, dy, dt, rho, mu, alpha_p, alpha_u, alpha_m;
  double bcw0, bcw1, bce0, bce1, bcn0, bcn1, bcs0, bcs1;
  int pin_p, east_wall_bug;
  int pu, pv, pp;           // pitches
  size_t plu, plv, plp;     // plane strides of coef_u/v/p
  int ib;                   // first row of this launch (row = ib + blockIdx.y): grids taller than 65535 rows go in chunks
};

static SP make_sp(const kt_simple_params* P) {
  SP s;
  s.nx = P->nx; s.ny = P->ny; s.dx = P->dx; s.dy = P->dy; s.dt = P->dt; s.rho = P->rho; s.mu = P->mu;
  s.alpha_p = P->alpha_p; s.alpha_u = P->alpha_u; s.alpha_m = P->alpha_m;
  s.bcw0 = P->bc_w[0]; s.bcw1 = P->bc_w[1]; s.bce0 = P->bc_e[0]; s.bce1 = P->bc_e[1];
  s.bcn0 = P->bc_n[0]; s.bcn1 = P->bc_n[1]; s.bcs0 = P->bc_s[0]; s.bcs1 = P->bc_s[1];
  s.pin_p = P->pin_p; s.east_wall_bug = P->east_wall_bug;
  s.pu = P->pitch_u; s.pv = P->pitch_v; s.pp = P->pitch_p;
  s.plu = (size_t)(P->nx + 3) * P->pitch_u; s.plv = (size_t)(P->nx + 2) * P->pitch_v; s.plp = (size_t)(P->nx + 2) * P->pitch_p;
  s.ib = 0;
  return s;
}

// one launch per chunk of <= 32768 rows (gridDim.y <= 65535 whatever nx is); S.ib carries the chunk's first row
#define KT_ROW_CHUNKS(S, rows, gx, gz, LAUNCH)                                      \
  for (int _r0 = 0; _r0 < (rows); _r0 += 32768) {                                   \
    (S).ib = _r0;                                                                   \
    const dim3 grid((gx), std::min(32768, (rows) - _r0), (gz));                     \
    LAUNCH;                                                                         \
    KT_LAUNCH_CHECK(h);                                                             \
  }

#define U(i, j) u[(size_t)(i) * S.pu + (j)]
#define Vv(i, j) v[(size_t)(i) * S.pv + (j)]
#define Pp(i, j) p[(size_t)(i) * S.pp + (j)]

// compute_coef_u (blockIdx.z == 0) and compute_coef_v (blockIdx.z == 1): bicgsimp.py:79-112
template <typename T>
__global__ void __launch_bounds__(256) k_assemble_uv(SP S, const T* __restrict__ u, const T* __restrict__ v,
                                                     const T* __restrict__ p, const T* __restrict__ u0,
                                                     const T* __restrict__ v0, T* cu, T* bu, T* cv, T* bv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = S.ib + (int)blockIdx.y;
  const T dx = T(S.dx), dy = T(S.dy), dt = T(S.dt), rho = T(S.rho), mu = T(S.mu);
  const T half = T(0.5);
  if (blockIdx.z == 0) {
    if (i < 2 || i > S.nx || j < 1 || j > S.ny) return;
    const T uc = U(i, j), uw = U(i - 1, j), ue = U(i + 1, j);
    const T vwn = Vv(i - 1, j + 1), vn = Vv(i, j + 1), vws = Vv(i - 1, j), vs = Vv(i, j);
    const T aw = -(mu * dy / dx + half * rho * half * (uc + uw) * dy);
    const T ae = -(mu * dy / dx - half * rho * half * (uc + ue) * dy);
    const T an = -(mu * dx / dy - half * rho * half * (vwn + vn) * dx);
    const T as = -(mu * dx / dy + half * rho * half * (vws + vs) * dx);
    const T ap = -(aw + ae + an + as) + rho * half * (uc + ue) * dy - rho * half * (uc + uw) * dy +
                 rho * half * (vwn + vn) * dx - rho * half * (vws + vs) * dx + rho * dx * dy / dt;
    const size_t o = (size_t)i * S.pu + j;
    cu[o] = ap; cu[S.plu + o] = aw; cu[2 * S.plu + o] = ae; cu[3 * S.plu + o] = an; cu[4 * S.plu + o] = as;
    bu[o] = (Pp(i - 1, j) - Pp(i, j)) * dy + rho * dx * dy / dt * u0[o];
  } else {
    if (i < 1 || i > S.nx || j < 2 || j > S.ny) return;
    const T vc = Vv(i, j), vn = Vv(i, j + 1), vs = Vv(i, j - 1);
    const T uc = U(i, j), us = U(i, j - 1), ues = U(i + 1, j - 1), ue = U(i + 1, j);
    const T aw = -(mu * dy / dx + half * rho * half * (uc + us) * dy);
    const T ae = -(mu * dy / dx - half * rho * half * (ues + ue) * dy);
    const T an = -(mu * dx / dy - half * rho * half * (vn + vc) * dx);
    const T as = -(mu * dx / dy + half * rho * half * (vs + vc) * dx);
    const T ap = -(aw + ae + an + as) + rho * half * (ues + ue) * dy - rho * half * (uc + us) * dy +
                 rho * half * (vn + vc) * dx - rho * half * (vs + vc) * dx + rho * dx * dy / dt;
    const size_t o = (size_t)i * S.pv + j;
    cv[o] = ap; cv[S.plv + o] = aw; cv[2 * S.plv + o] = ae; cv[3 * S.plv + o] = an; cv[4 * S.plv + o] = as;
    bv[o] = (Pp(i, j - 1) - Pp(i, j)) * dx + rho * dx * dy / dt * v0[o];
  }
}

// set_bc (bicgsimp.py:153-197).  The reference runs four 1-D loops one after another; loop 3 touches
// only u-rows i in [2,nx] at j = 1, ny and loop 1 only (2|nx, j): the (element, slot) sets written by
// different loops are disjoint except b_u[2|nx, 1|ny] and coef_u ap/aw/ae/as/an slots, which are
// different slots -- so one thread per boundary element applying its loops' statements in reference
// order reproduces the sequential result.  blockIdx.y selects the loop.
template <typename T>
__global__ void __launch_bounds__(128) k_set_bc(SP S, T* u, T* v, T* cu, T* bu, T* cv, T* bv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nx = S.nx, ny = S.ny;
  if (blockIdx.y == 0) {          // loop 1: j in [1, ny]
    const int j = t + 1;
    if (j > ny) return;
    size_t o = (size_t)2 * S.pu + j;
    bu[o] += -cu[S.plu + o] * T(S.bcw0);
    cu[S.plu + o] = T(0);
    U(1, j) = T(S.bcw0);
    o = (size_t)nx * S.pu + j;
    bu[o] += -cu[2 * S.plu + o] * T(S.bce0);
    cu[2 * S.plu + o] = T(0);
    U(nx + 1, j) = T(S.bce0);
  } else if (blockIdx.y == 1) {   // loop 2: i in [1, nx]
    const int i = t + 1;
    if (i > nx) return;
    size_t o = (size_t)i * S.pv + 2;
    bv[o] += -cv[4 * S.plv + o] * T(S.bcs0);
    cv[4 * S.plv + o] = T(0);
    Vv(i, 1) = T(S.bcs0);
    o = (size_t)i * S.pv + ny;
    bv[o] += -cv[3 * S.plv + o] * T(S.bcn0);
    cv[3 * S.plv + o] = T(0);
    Vv(i, ny + 1) = T(S.bcn0);
  }
}

// loops 3 and 4 must see loop 1/2 results on shared b_u / b_v elements -> second launch.
template <typename T>
__global__ void __launch_bounds__(128) k_set_bc_walls(SP S, T* cu, T* bu, T* cv, T* bv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int nx = S.nx, ny = S.ny;
  const T mu = T(S.mu), dx = T(S.dx), dy = T(S.dy);
  const T two = T(2);
  if (blockIdx.y == 0) {          // loop 3: i in [2, nx]
    const int i = t + 2;
    if (i > nx) return;
    size_t o = (size_t)i * S.pu + 1;
    bu[o] += two * mu * T(S.bcs1) * dx / dy;
    cu[o] += (cu[4 * S.plu + o] + two * mu * dx / dy);
    cu[4 * S.plu + o] = T(0);
    o = (size_t)i * S.pu + ny;
    bu[o] += two * mu * T(S.bcn1) * dx / dy;
    cu[o] += (cu[3 * S.plu + o] + two * mu * dx / dy);
    cu[3 * S.plu + o] = T(0);
  } else {                        // loop 4: j in [2, ny]
    const int j = t + 2;
    if (j > ny) return;
    size_t o = (size_t)1 * S.pv + j;
    bv[o] += two * mu * T(S.bcw1) * dy / dx;
    cv[o] += (cv[S.plv + o] + two * mu * dy / dx);
    cv[S.plv + o] = T(0);
    o = (size_t)nx * S.pv + j;
    bv[o] += two * mu * T(S.bce1) * dy / dx;
    const int k_east = S.east_wall_bug ? 1 : 2;      // simple.py:215 vs bicgsimp.py:196
    cv[o] += (cv[k_east * S.plv + o] + two * mu * dy / dx);
    cv[2 * S.plv + o] = T(0);
  }
}

// compute_coef_p (bicgsimp.py:124-151)
template <typename T>
__global__ void __launch_bounds__(256) k_assemble_p(SP S, const T* __restrict__ u, const T* __restrict__ v,
                                                    const T* __restrict__ cu, const T* __restrict__ cv, T* cp, T* bp,
                                                    T* mdiv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = S.ib + (int)blockIdx.y;
  if (i < 1 || i > S.nx || j < 1 || j > S.ny) return;
  const T dx = T(S.dx), dy = T(S.dy), rho = T(S.rho);
  const T m = rho * (U(i, j) - U(i + 1, j)) * dy + rho * (Vv(i, j) - Vv(i, j + 1)) * dx;
  const size_t o = (size_t)i * S.pp + j;
  mdiv[o] = m;
  // the reference divides by the never-written (zero) ap of the boundary faces and then overwrites the
  // resulting -inf with 0 (:137-144); branch instead of computing the inf.
  T aw = (i == 1) ? T(0) : -rho * dy * dy / cu[(size_t)i * S.pu + j];
  T ae = (i == S.nx) ? T(0) : -rho * dy * dy / cu[(size_t)(i + 1) * S.pu + j];
  T an = (j == S.ny) ? T(0) : -rho * dx * dx / cv[(size_t)i * S.pv + j + 1];
  T as = (j == 1) ? T(0) : -rho * dx * dx / cv[(size_t)i * S.pv + j];
  T ap = -(aw + ae + an + as);
  T b = m;
  if (S.pin_p && i == 1 && j == 1) { aw = ae = an = as = T(0); ap = T(1); b = T(0); }
  cp[o] = ap; cp[S.plp + o] = aw; cp[2 * S.plp + o] = ae; cp[3 * S.plp + o] = an; cp[4 * S.plp + o] = as;
  bp[o] = b;
}

// update_velocity (bicgsimp.py:212-225): grid-stride over rows so the partial count is bounded.
template <typename T>
__global__ void __launch_bounds__(256) k_relax_uv(SP S, T* u, T* v, const T* __restrict__ um, const T* __restrict__ vm,
                                                  double* partials, int max_blocks, unsigned int* ticket, double* out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const T am = T(S.alpha_m), one = T(1);
  double mu_ = 0.0, mv_ = 0.0;
  for (int i = blockIdx.y; i <= S.nx; i += gridDim.y) {
    if (i >= 2 && j >= 1 && j <= S.ny) {
      const size_t o = (size_t)i * S.pu + j;
      const T a = um[o], b = u[o];
      mu_ = fmax(mu_, (double)fabs(a - b));
      u[o] = am * a + (one - am) * b;
    }
    if (i >= 1 && j >= 2 && j <= S.ny) {
      const size_t o = (size_t)i * S.pv + j;
      const T a = vm[o], b = v[o];
      mv_ = fmax(mv_, (double)fabs(a - b));
      v[o] = am * a + (one - am) * b;
    }
  }
  block_max_finalize(mu_, mv_, partials, max_blocks, ticket, out);
}

// correct_pressure + correct_velocity (bicgsimp.py:232-244)
template <typename T>
__global__ void __launch_bounds__(256) k_correct(SP S, T* p, T* u, T* v, const T* __restrict__ pc,
                                                 const T* __restrict__ cu, const T* __restrict__ cv, int do_p, int do_uv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = S.ib + (int)blockIdx.y;
  if (i < 1 || i > S.nx || j < 1 || j > S.ny) return;
  const T dx = T(S.dx), dy = T(S.dy);
  const T pcc = pc[(size_t)i * S.pp + j];
  if (do_p) Pp(i, j) += T(S.alpha_p) * pcc;
  if (!do_uv) return;
  if (i >= 2) {
    const size_t o = (size_t)i * S.pu + j;
    u[o] += T(S.alpha_u) * (pc[(size_t)(i - 1) * S.pp + j] - pcc) * dy / cu[o];
  }
  if (j >= 2) {
    const size_t o = (size_t)i * S.pv + j;
    v[o] += T(S.alpha_u) * (pc[(size_t)i * S.pp + j - 1] - pcc) * dx / cv[o];
  }
}

// compute_mdiv (bicgsimp.py:114-122)
template <typename T>
__global__ void __launch_bounds__(256) k_mdiv(SP S, const T* __restrict__ u, const T* __restrict__ v, T* mdiv,
                                              double* partials, int max_blocks, unsigned int* ticket, double* out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const T dx = T(S.dx), dy = T(S.dy), rho = T(S.rho);
  double mx = 0.0;
  for (int i = 1 + blockIdx.y; i <= S.nx; i += gridDim.y) {
    if (j >= 1 && j <= S.ny) {
      const T m = rho * (U(i, j) - U(i + 1, j)) * dy + rho * (Vv(i, j) - Vv(i, j + 1)) * dx;
      mdiv[(size_t)i * S.pp + j] = m;
      mx = fmax(mx, (double)fabs(m));
    }
  }
  block_max_finalize(mx, 0.0, partials, max_blocks, ticket, out);
}

static int read_red(kt_handle_s* h, double* out, int n) {
  if (int rc = publish(h, h->red_out, h->red_host, 2)) return rc;
  KT_CUDA(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < n; ++k) out[k] = h->red_host[k];
  return 0;
}

// grid.y for the max-reducing kernels: bounded so gridDim.x*gridDim.y <= max_blocks
static int rows_grid(const kt_handle_s* h, int gx, int rows) {
  int gy = h->max_blocks / gx;
  if (gy < 1) gy = 1;
  return gy < rows ? gy : rows;
}

}  // namespace kt

using namespace kt;

#define KT_SIMPLE_ARGS_OK(P) KT_CHECK_ARG((P) && (P)->nx >= 2 && (P)->ny >= 2 && (P)->pitch_u >= (P)->ny + 2 && \
                                          (P)->pitch_v >= (P)->ny + 3 && (P)->pitch_p >= (P)->ny + 2, "kt_simple_params: bad geometry")

extern "C" {

int kt_assemble_uv(kt_handle_t h, const kt_simple_params* P, const void* u, const void* v, const void* p, const void* u0,
                   const void* v0, void* coef_u, void* b_u, void* coef_v, void* b_v) {
  KT_CHECK_ARG(h && u && v && p && u0 && v0 && coef_u && b_u && coef_v && b_v, "kt_assemble_uv: null argument");
  KT_DEV_GUARD(h);
  KT_SIMPLE_ARGS_OK(P);
  SP S = make_sp(P);
  if (h->dtype == KT_F64) {
    KT_ROW_CHUNKS(S, P->nx + 1, (P->ny + 2 + 255) / 256, 2,
                  (k_assemble_uv<double><<<grid, 256, 0, h->stream>>>(S, (const double*)u, (const double*)v, (const double*)p, (const double*)u0,
                                                                     (const double*)v0, (double*)coef_u, (double*)b_u, (double*)coef_v, (double*)b_v)));
  } else {
    KT_ROW_CHUNKS(S, P->nx + 1, (P->ny + 2 + 255) / 256, 2,
                  (k_assemble_uv<float><<<grid, 256, 0, h->stream>>>(S, (const float*)u, (const float*)v, (const float*)p, (const float*)u0,
                                                                    (const float*)v0, (float*)coef_u, (float*)b_u, (float*)coef_v, (float*)b_v)));
  }
  return 0;
}

int kt_set_bc_cavity(kt_handle_t h, const kt_simple_params* P, void* u, void* v, void* coef_u, void* b_u, void* coef_v, void* b_v) {
  KT_CHECK_ARG(h && u && v && coef_u && b_u && coef_v && b_v, "kt_set_bc_cavity: null argument");
  KT_DEV_GUARD(h);
  KT_SIMPLE_ARGS_OK(P);
  SP S = make_sp(P);
  const int n = (P->nx > P->ny ? P->nx : P->ny);
  dim3 grid((n + 127) / 128, 2);
  if (h->dtype == KT_F64) {
    k_set_bc<double><<<grid, 128, 0, h->stream>>>(S, (double*)u, (double*)v, (double*)coef_u, (double*)b_u, (double*)coef_v, (double*)b_v);
    KT_LAUNCH_CHECK(h);
    k_set_bc_walls<double><<<grid, 128, 0, h->stream>>>(S, (double*)coef_u, (double*)b_u, (double*)coef_v, (double*)b_v);
  } else {
    k_set_bc<float><<<grid, 128, 0, h->stream>>>(S, (float*)u, (float*)v, (float*)coef_u, (float*)b_u, (float*)coef_v, (float*)b_v);
    KT_LAUNCH_CHECK(h);
    k_set_bc_walls<float><<<grid, 128, 0, h->stream>>>(S, (float*)coef_u, (float*)b_u, (float*)coef_v, (float*)b_v);
  }
  KT_LAUNCH_CHECK(h);
  return 0;
}

int kt_assemble_p(kt_handle_t h, const kt_simple_params* P, const void* u, const void* v, const void* coef_u, const void* coef_v,
                  void* coef_p, void* b_p, void* mdiv) {
  KT_CHECK_ARG(h && u && v && coef_u && coef_v && coef_p && b_p && mdiv, "kt_assemble_p: null argument");
  KT_DEV_GUARD(h);
  KT_SIMPLE_ARGS_OK(P);
  SP S = make_sp(P);
  if (h->dtype == KT_F64) {
    KT_ROW_CHUNKS(S, P->nx + 1, (P->ny + 2 + 255) / 256, 1,
                  (k_assemble_p<double><<<grid, 256, 0, h->stream>>>(S, (const double*)u, (const double*)v, (const double*)coef_u, (const double*)coef_v,
                                                                    (double*)coef_p, (double*)b_p, (double*)mdiv)));
  } else {
    KT_ROW_CHUNKS(S, P->nx + 1, (P->ny + 2 + 255) / 256, 1,
                  (k_assemble_p<float><<<grid, 256, 0, h->stream>>>(S, (const float*)u, (const float*)v, (const float*)coef_u, (const float*)coef_v,
                                                                   (float*)coef_p, (float*)b_p, (float*)mdiv)));
  }
  return 0;
}

int kt_relax_uv(kt_handle_t h, const kt_simple_params* P, void* u, void* v, const void* u_mid, const void* v_mid, double* residual_host) {
  KT_CHECK_ARG(h && u && v && u_mid && v_mid && residual_host, "kt_relax_uv: null argument");
  KT_DEV_GUARD(h);
  KT_SIMPLE_ARGS_OK(P);
  SP S = make_sp(P);
  const int gx = (P->ny + 2 + 255) / 256;
  dim3 grid(gx, rows_grid(h, gx, P->nx + 1));
  if (h->dtype == KT_F64)
    k_relax_uv<double><<<grid, 256, 0, h->stream>>>(S, (double*)u, (double*)v, (const double*)u_mid, (const double*)v_mid, h->partials,
                                                   h->max_blocks, h->ticket, h->red_out);
  else
    k_relax_uv<float><<<grid, 256, 0, h->stream>>>(S, (float*)u, (float*)v, (const float*)u_mid, (const float*)v_mid, h->partials,
                                                  h->max_blocks, h->ticket, h->red_out);
  KT_LAUNCH_CHECK(h);
  double m[2];
  int rc = read_red(h, m, 2);
  if (rc) return rc;
  *residual_host = sqrt(m[0] * m[0] + m[1] * m[1]);   // bicgsimp.py:225
  return 0;
}

int kt_correct_puv(kt_handle_t h, const kt_simple_params* P, void* p, void* u, void* v, const void* pcor, const void* coef_u,
                   const void* coef_v) {
  KT_CHECK_ARG(h && pcor, "kt_correct_puv: null argument");
  KT_DEV_GUARD(h);
  const int do_p = p != nullptr, do_uv = (u != nullptr) || (v != nullptr);
  KT_CHECK_ARG(do_p || do_uv, "kt_correct_puv: nothing to correct");
  KT_CHECK_ARG(!do_uv || (u && v && coef_u && coef_v), "kt_correct_puv: velocity correction needs u, v, coef_u, coef_v");
  KT_SIMPLE_ARGS_OK(P);
  SP S = make_sp(P);
  if (h->dtype == KT_F64) {
    KT_ROW_CHUNKS(S, P->nx + 1, (P->ny + 2 + 255) / 256, 1,
                  (k_correct<double><<<grid, 256, 0, h->stream>>>(S, (double*)p, (double*)u, (double*)v, (const double*)pcor, (const double*)coef_u,
                                                                 (const double*)coef_v, do_p, do_uv)));
  } else {
    KT_ROW_CHUNKS(S, P->nx + 1, (P->ny + 2 + 255) / 256, 1,
                  (k_correct<float><<<grid, 256, 0, h->stream>>>(S, (float*)p, (float*)u, (float*)v, (const float*)pcor, (const float*)coef_u,
                                                                (const float*)coef_v, do_p, do_uv)));
  }
  return 0;
}

int kt_mdiv_max(kt_handle_t h, const kt_simple_params* P, const void* u, const void* v, void* mdiv, double* max_host) {
  KT_CHECK_ARG(h && u && v && mdiv && max_host, "kt_mdiv_max: null argument");
  KT_DEV_GUARD(h);
  KT_SIMPLE_ARGS_OK(P);
  SP S = make_sp(P);
  const int gx = (P->ny + 2 + 255) / 256;
  dim3 grid(gx, rows_grid(h, gx, P->nx));
  if (h->dtype == KT_F64)
    k_mdiv<double><<<grid, 256, 0, h->stream>>>(S, (const double*)u, (const double*)v, (double*)mdiv, h->partials, h->max_blocks, h->ticket, h->red_out);
  else
    k_mdiv<float><<<grid, 256, 0, h->stream>>>(S, (const float*)u, (const float*)v, (float*)mdiv, h->partials, h->max_blocks, h->ticket, h->red_out);
  KT_LAUNCH_CHECK(h);
  return read_red(h, max_host, 1);
}

}  // extern "C"

// every kernel of this translation unit, for kt_preload_kernels (kt_handle.h)
namespace kt {
int preload_kt_simple() {
  static const void* const fns[] = {
#define KT_PRELOAD_TU_kt_simple
#include "kt_preload_list.inc"
#undef KT_PRELOAD_TU_kt_simple
      nullptr};
  return preload_list(fns);
}
}  // namespace kt
