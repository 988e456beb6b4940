// kt_post.cu -- the rows either side of the hot path: the display frame of the reference and the
// Jacobi sweeps of its validation scripts, as device kernels.
//
//   kt_display_frame  display.py:17-49 (class version) and karman_taichi.py:544-583,662-665 (monolith): cell-centre
//                     interpolation of u and v, the five panels scaled to [0, 1] by their own min / max, concatenated along
//                     j into one dense image, the monolith's x3 nearest-neighbour upsample included.
//   kt_jacobi_sweep   pressure_valid.py:220-239 / momentum_valid.py:157-170: one point-Jacobi sweep of a 5-point system with
//                     the max-norm of the update as residual.
#include <algorithm>
#include "kt_handle.h"

namespace kt {

struct FrameArgs {
  int nx, ny, pu, pv, pp;       // cells; row pitches of the u (nx+3, ny+2), v (nx+2, ny+3) and cell-centred (nx+2, ny+2) planes
  int scale;                    // 1: display.py, 3: karman_taichi.py:561-567
  int monolith;                 // panel order (pcor, udiv, p, u, v) and f_max starting at 0.0 (karman_taichi.py:575,665)
  double dx, dy;
};

// value of panel q (0 u, 1 v, 2 p, 3 pcor, 4 div) at cell (i, j) of the (nx+2, ny+2) canvas
template <typename T>
__device__ __forceinline__ T panel_value(const FrameArgs& A, int q, int i, int j, const T* u, const T* v, const T* p, const T* pcor,
                                         const T* div) {
  switch (q) {
    case 0: return T(0.5) * (u[(size_t)i * A.pu + j] + u[(size_t)(i + 1) * A.pu + j]);      // display.py:31, karman_taichi.py:550
    case 1: return T(0.5) * (v[(size_t)i * A.pv + j] + v[(size_t)i * A.pv + j + 1]);        // display.py:32, karman_taichi.py:551
    case 2: return p[(size_t)i * A.pp + j];
    case 3: return pcor[(size_t)i * A.pp + j];
    default:
      if (A.monolith && i >= 1 && i <= A.nx && j >= 1 && j <= A.ny)                           // karman_taichi.py:555-556
        return (u[(size_t)i * A.pu + j] - u[(size_t)(i + 1) * A.pu + j]) * T(A.dy) + (v[(size_t)i * A.pv + j + 1] - v[(size_t)i * A.pv + j]) * T(A.dx);
      return div[(size_t)i * A.pp + j];
  }
}

// pass 1: min / max of every panel (scale_field, display.py:17-24: f_max starts at -1e9 -- 0.0 in the monolith,
// karman_taichi.py:575 -- and f_min at 1e9; a NaN never wins a comparison).  One block per row chunk; partials -> last block.
template <typename T>
__global__ void __launch_bounds__(256) k_frame_minmax(FrameArgs A, const T* u, const T* v, const T* p, const T* pcor, T* div,
                                                      double* partials, unsigned int* ticket, double* out10) {
  __shared__ double sm[10][8];
  __shared__ bool is_last;
  double mx[5], mn[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) { mx[q] = A.monolith ? 0.0 : -1.0e9; mn[q] = 1.0e9; }
  const int NXc = A.nx + 2, NYc = A.ny + 2;
  for (int i = blockIdx.x; i < NXc; i += gridDim.x)
    for (int j = threadIdx.x; j < NYc; j += blockDim.x) {
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const double f = (double)panel_value<T>(A, q, i, j, u, v, p, pcor, div);
        if (q == 4 && A.monolith && i >= 1 && i <= A.nx && j >= 1 && j <= A.ny) div[(size_t)i * A.pp + j] = T(f);   // the udiv field itself
        mx[q] = fmax(mx[q], f); mn[q] = fmin(mn[q], f);
      }
    }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    for (int o = 16; o > 0; o >>= 1) { mx[q] = fmax(mx[q], __shfl_down_sync(0xffffffffu, mx[q], o)); mn[q] = fmin(mn[q], __shfl_down_sync(0xffffffffu, mn[q], o)); }
    if (lane == 0) { sm[q][wid] = mx[q]; sm[5 + q][wid] = mn[q]; }
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double a = sm[threadIdx.x][0];
    for (int w = 1; w < 8; ++w) a = threadIdx.x < 5 ? fmax(a, sm[threadIdx.x][w]) : fmin(a, sm[threadIdx.x][w]);
    partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < 10) {
    double a = __ldcg(partials + (size_t)threadIdx.x * gridDim.x);
    for (unsigned b = 1; b < gridDim.x; ++b) {
      const double c = __ldcg(partials + (size_t)threadIdx.x * gridDim.x + b);
      a = threadIdx.x < 5 ? fmax(a, c) : fmin(a, c);
    }
    out10[threadIdx.x] = a;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// pass 2: img[I, q' * S * NYc + J] = (f - f_min) / (f_max - f_min + 1e-9) with f the value of cell (I / S, J / S) of the panel shown at
// position q' (display.py:26,45-46; karman_taichi.py:562-567,583,665)
template <typename T>
__global__ void __launch_bounds__(256) k_frame_write(FrameArgs A, const T* u, const T* v, const T* p, const T* pcor, const T* div,
                                                     const double* mm10, T* img) {
  const int S = A.scale, NYc = A.ny + 2;
  const int W = 5 * S * NYc;                      // image row length
  const int I = blockIdx.y;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= W) return;
  const int pos = col / (S * NYc), J = col - pos * (S * NYc);
  const int order_class[5] = {0, 1, 2, 3, 4}, order_mono[5] = {3, 4, 2, 0, 1};
  const int q = A.monolith ? order_mono[pos] : order_class[pos];
  const T f = panel_value<T>(A, q, I / S, J / S, u, v, p, pcor, div);
  const T fmax_ = T(mm10[q]), fmin_ = T(mm10[5 + q]);
  img[(size_t)I * W + col] = (f - fmin_) / (fmax_ - fmin_ + T(1.0e-9));
}

// One point-Jacobi sweep over rows [i_lo, i_hi), columns [j_lo, j_hi):
//   x_mid = (-aw x[i-1,j] - ae x[i+1,j] - an x[i,j+1] - as x[i,j-1] + rhs) / ap       pressure_valid.py:224-228
//   residual = max |x - x_mid|                                                        :229-230
// (the reference then overwrites x in the same loop, i.e. whatever order its threads run in; here the sweep reads the old
// iterate everywhere -- the schedule in which every read precedes every write -- and kt_jacobi_sweep blends afterwards)
template <typename T>
__global__ void __launch_bounds__(256) k_jacobi_mid(const T* __restrict__ coef, size_t plane, const T* __restrict__ rhs,
                                                    const T* __restrict__ x, T* __restrict__ x_mid, int pitch, int i_lo, int i_hi,
                                                    int j_lo, int j_hi, double* partials, int max_blocks, unsigned int* ticket, double* out) {
  const int j = j_lo + blockIdx.x * blockDim.x + threadIdx.x;
  double res = 0.0;
  for (int i = i_lo + blockIdx.y; i < i_hi; i += gridDim.y) {
    if (j < j_hi) {
      const size_t o = (size_t)i * pitch + j;
      T acc = -coef[plane + o] * x[o - pitch];
      acc = acc - coef[2 * plane + o] * x[o + pitch];
      acc = acc - coef[3 * plane + o] * x[o + 1];
      acc = acc - coef[4 * plane + o] * x[o - 1];
      const T mid = (acc + rhs[o]) / coef[o];
      x_mid[o] = mid;
      res = fmax(res, (double)fabs(x[o] - mid));
    }
  }
  block_max_finalize(res, 0.0, partials, max_blocks, ticket, out);
}

template <typename T>
__global__ void __launch_bounds__(256) k_jacobi_blend(T* __restrict__ x, const T* __restrict__ x_mid, int pitch, int i_lo, int i_hi,
                                                      int j_lo, int j_hi, T relax) {
  const int j = j_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= j_hi) return;
  for (int i = i_lo + blockIdx.y; i < i_hi; i += gridDim.y) {
    const size_t o = (size_t)i * pitch + j;
    // momentum_valid.py:168 u = (1 - alpha_u) u + alpha_u u_mid ; pressure_valid.py:231 pcor = pcor_mid (relax = 1: exact copy)
    x[o] = (relax == T(1)) ? x_mid[o] : (T(1) - relax) * x[o] + relax * x_mid[o];
  }
}

}  // namespace kt

using namespace kt;

extern "C" {

int kt_display_frame(kt_handle_t h, const kt_frame_params* P, const void* u, const void* v, const void* p, const void* pcor, void* div,
                     void* img, double* minmax_host10) {
  KT_CHECK_ARG(h && P && u && v && p && pcor && div && img, "kt_display_frame: null argument");
  KT_DEV_GUARD(h);
  KT_CHECK_ARG(P->nx >= 1 && P->ny >= 1 && (P->scale == 1 || P->scale == 3), "kt_display_frame: nx, ny >= 1 and scale 1 or 3");
  FrameArgs A;
  A.nx = P->nx; A.ny = P->ny; A.pu = P->pitch_u; A.pv = P->pitch_v; A.pp = P->pitch_p; A.scale = P->scale; A.monolith = P->monolith;
  A.dx = P->dx; A.dy = P->dy;
  const int g1 = std::min(P->nx + 2, h->max_blocks / 8);
  double* mm = h->partials + 10 * (size_t)g1;            // 10 doubles behind the block partials
  const int W = 5 * A.scale * (A.ny + 2);
  dim3 g2((W + 255) / 256, 1);
  if (h->dtype == KT_F64) k_frame_minmax<double><<<g1, 256, 0, h->stream>>>(A, (const double*)u, (const double*)v, (const double*)p, (const double*)pcor, (double*)div, h->partials, h->ticket, mm);
  else k_frame_minmax<float><<<g1, 256, 0, h->stream>>>(A, (const float*)u, (const float*)v, (const float*)p, (const float*)pcor, (float*)div, h->partials, h->ticket, mm);
  KT_LAUNCH_CHECK(h);
  g2.y = A.scale * (A.nx + 2);
  KT_CHECK_ARG(g2.y <= 65535, "kt_display_frame: canvases taller than 65535 rows are not supported");
  if (h->dtype == KT_F64) k_frame_write<double><<<g2, 256, 0, h->stream>>>(A, (const double*)u, (const double*)v, (const double*)p, (const double*)pcor, (const double*)div, mm, (double*)img);
  else k_frame_write<float><<<g2, 256, 0, h->stream>>>(A, (const float*)u, (const float*)v, (const float*)p, (const float*)pcor, (const float*)div, mm, (float*)img);
  KT_LAUNCH_CHECK(h);
  if (minmax_host10) {
    KT_CUDA(cudaMemcpyAsync(minmax_host10, mm, 10 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    KT_CUDA(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int kt_jacobi_sweep(kt_handle_t h, const void* coef, const void* rhs, void* x, void* x_mid, int i_lo, int i_hi, int j_lo, int j_hi,
                    double relax, double* residual_host) {
  KT_CHECK_ARG(h && coef && rhs && x && x_mid, "kt_jacobi_sweep: null argument");
  KT_DEV_GUARD(h);
  KT_CHECK_ARG(1 <= i_lo && i_lo <= i_hi && i_hi <= h->NX - 1 && 1 <= j_lo && j_lo <= j_hi && j_hi <= h->NY - 1,
               "kt_jacobi_sweep: the swept range must leave a one-cell ring inside the (NX, NY) plane");
  if (i_lo == i_hi || j_lo == j_hi) { if (residual_host) *residual_host = 0.0; return 0; }
  const size_t plane = (size_t)h->NX * h->pitch;
  const int gx = (j_hi - j_lo + 255) / 256;
  int gy = std::min(i_hi - i_lo, std::max(1, std::min(65535, h->max_blocks / gx)));
  dim3 grid(gx, gy);
  if (h->dtype == KT_F64) {
    k_jacobi_mid<double><<<grid, 256, 0, h->stream>>>((const double*)coef, plane, (const double*)rhs, (const double*)x, (double*)x_mid, h->pitch, i_lo, i_hi, j_lo, j_hi, h->partials, h->max_blocks, h->ticket, h->red_out);
    KT_LAUNCH_CHECK(h);
    k_jacobi_blend<double><<<grid, 256, 0, h->stream>>>((double*)x, (const double*)x_mid, h->pitch, i_lo, i_hi, j_lo, j_hi, relax);
  } else {
    k_jacobi_mid<float><<<grid, 256, 0, h->stream>>>((const float*)coef, plane, (const float*)rhs, (const float*)x, (float*)x_mid, h->pitch, i_lo, i_hi, j_lo, j_hi, h->partials, h->max_blocks, h->ticket, h->red_out);
    KT_LAUNCH_CHECK(h);
    k_jacobi_blend<float><<<grid, 256, 0, h->stream>>>((float*)x, (const float*)x_mid, h->pitch, i_lo, i_hi, j_lo, j_hi, (float)relax);
  }
  KT_LAUNCH_CHECK(h);
  if (residual_host) {
    KT_CUDA(cudaMemcpyAsync(h->red_host, h->red_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    KT_CUDA(cudaStreamSynchronize(h->stream));
    *residual_host = h->red_host[0];
  }
  return 0;
}

}  // extern "C"

// every kernel of this translation unit, for kt_preload_kernels (kt_handle.h)
namespace kt {
int preload_kt_post() {
  static const void* const fns[] = {
#define KT_PRELOAD_TU_kt_post
#include "kt_preload_list.inc"
#undef KT_PRELOAD_TU_kt_post
      nullptr};
  return preload_list(fns);
}
}  // namespace kt
