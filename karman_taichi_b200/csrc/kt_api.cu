// kt_api.cu -- handle lifetime, error reporting, host<->device staging and layout conversion.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include "kt_handle.h"

namespace kt {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in `" + what + "` at " + file + ":" + std::to_string(line);
  return -1;
}

// dense AoS (NX, NY, 5) <-> 5 pitched planes.  One thread per (i, j); the AoS side is staged through shared
// memory so both sides are coalesced.
template <typename T, bool TO_SOA>
__global__ void __launch_bounds__(256) k_coef_layout(T* soa, T* aos, int NX, int NY, int pitch, int row0) {
  __shared__ T tile[256 * 5];
  const int i = row0 + blockIdx.y;          // plane row; the AoS buffer starts at row0 (chunked staging)
  const int jb = blockIdx.x * 256;
  const int nj = min(256, NY - jb);
  const size_t plane = (size_t)NX * pitch;
  T* a = aos + ((size_t)blockIdx.y * NY + jb) * 5;
  if (TO_SOA) {
    for (int e = threadIdx.x; e < nj * 5; e += 256) tile[e] = a[e];
    __syncthreads();
    if ((int)threadIdx.x < nj)
      for (int k = 0; k < 5; ++k) soa[k * plane + (size_t)i * pitch + jb + threadIdx.x] = tile[threadIdx.x * 5 + k];
  } else {
    if ((int)threadIdx.x < nj)
      for (int k = 0; k < 5; ++k) tile[threadIdx.x * 5 + k] = soa[k * plane + (size_t)i * pitch + jb + threadIdx.x];
    __syncthreads();
    for (int e = threadIdx.x; e < nj * 5; e += 256) a[e] = tile[e];
  }
}

template <typename T, bool TO_SOA>
static int coef_layout(kt_handle_s* h, void* soa, void* aos, int row0 = 0, int nrows = -1) {
  if (nrows < 0) nrows = h->NX;
  for (int r = 0; r < nrows; r += 32768) {            // gridDim.y <= 65535
    const int nr = std::min(32768, nrows - r);
    dim3 grid((h->NY + 255) / 256, nr);
    k_coef_layout<T, TO_SOA><<<grid, 256, 0, h->stream>>>((T*)soa, (T*)aos + (size_t)r * h->NY * 5, h->NX, h->NY, h->pitch, row0 + r);
    KT_LAUNCH_CHECK(h);
  }
  return 0;
}

static size_t esize(const kt_handle_s* h) { return h->dtype == KT_F64 ? 8 : 4; }

// kt::publish (kt_handle.h): one warp moves n 8-byte words; dst may be pinned host memory (zero-copy stores)
__global__ void k_publish(const unsigned long long* src, unsigned long long* dst, int n) {
  for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
}

int publish(kt_handle_s* h, const void* src_dev, void* dst, int n_words) {
  k_publish<<<1, 32, 0, h->stream>>>((const unsigned long long*)src_dev, (unsigned long long*)dst, n_words);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_publish launch", __FILE__, __LINE__);
  return 0;
}

}  // namespace kt

using namespace kt;

extern "C" {

int kt_version(void) { return 100; }

const char* kt_last_error(void) { return g_err.c_str(); }

// Rows are padded to a multiple of 512 bytes: a warp moves 32 x 16 B = 512 B per array and request, and on B200 a warp-wide
// request that straddles a 512 B boundary costs DRAM efficiency (measured, profiles/r02_p2p.md: the x/r kernel at a 128 B
// phase runs 13 % slower, at 256 B 4 % slower).  With this pitch every row -- hence the first owned row of every x-slab --
// starts on such a boundary.  (KT_PITCH_BYTES: measurement aid; kt_create accepts any multiple of 128 bytes.)
int kt_pitch_for(int NY, int dtype) {
  static int bytes = [] { const char* v = getenv("KT_PITCH_BYTES"); int b = (v && *v) ? atoi(v) : 512; return (b >= 128 && b % 128 == 0) ? b : 512; }();
  const int q = bytes / ((dtype == KT_F64) ? 8 : 4);
  return (NY + q - 1) / q * q;
}

int kt_create(int device, int NX, int NY, int pitch, int dtype, void* stream, kt_handle_t* out) {
  KT_CHECK_ARG(out, "kt_create: out is null");
  KT_CHECK_ARG(dtype == KT_F64 || dtype == KT_F32, "kt_create: dtype must be KT_F64 or KT_F32");
  KT_CHECK_ARG(NX > 0 && NY > 0, "kt_create: NX, NY must be positive");
  KT_CHECK_ARG(pitch >= NY && pitch % ((dtype == KT_F64) ? 16 : 32) == 0, "kt_create: pitch must be >= NY and a multiple of 128 bytes");
  int ndev = 0;
  KT_CUDA(cudaGetDeviceCount(&ndev));
  KT_CHECK_ARG(device >= 0 && device < ndev, "kt_create: no such CUDA device");
  kt::DeviceGuard guard(device);          // allocate on `device`, leave the caller's current device untouched
  kt_handle_s* h = new kt_handle_s();
  h->device = device; h->NX = NX; h->NY = NY; h->pitch = pitch; h->dtype = dtype;
  h->row_lo = 0; h->row_hi = NX;
  h->stream = (cudaStream_t)stream;
  cudaDeviceProp prop;
  KT_CUDA(cudaGetDeviceProperties(&prop, device));
  h->num_sms = prop.multiProcessorCount;
  h->max_blocks = h->num_sms * 512;  // upper bound of any grid that produces block partials (1.2 MB of partials)
  KT_CUDA(cudaMalloc(&h->partials, sizeof(double) * KT_MAX_DOTS * h->max_blocks));
  KT_CUDA(cudaMalloc(&h->ticket, sizeof(unsigned int)));
  KT_CUDA(cudaMemset(h->ticket, 0, sizeof(unsigned int)));
  KT_CUDA(cudaMalloc(&h->sc, sizeof(Scalars)));
  KT_CUDA(cudaMemset(h->sc, 0, sizeof(Scalars)));
  KT_CUDA(cudaMallocHost(&h->sc_host, 2 * sizeof(Scalars)));
  KT_CUDA(cudaMalloc(&h->red_out, 4 * sizeof(double)));
  KT_CUDA(cudaMallocHost(&h->red_host, 4 * sizeof(double)));
  for (int k = 0; k < 2; ++k) KT_CUDA(cudaEventCreateWithFlags(&h->ev[k], cudaEventDisableTiming));
  KT_CUDA(cudaEventCreate(&h->t0));
  KT_CUDA(cudaEventCreate(&h->t1));
  *out = h;
  return 0;
}

int kt_destroy(kt_handle_t h) {
  if (!h) return 0;
  kt::DeviceGuard guard(h->device);
  cudaFree(h->partials); cudaFree(h->ticket); cudaFree(h->sc); cudaFreeHost(h->sc_host);
  cudaFree(h->red_out); cudaFreeHost(h->red_host);
  if (h->p2p_peers_dev) cudaFree(h->p2p_peers_dev);
  for (int k = 0; k < 2; ++k) if (h->ev[k]) cudaEventDestroy(h->ev[k]);
  if (h->t0) cudaEventDestroy(h->t0);
  if (h->t1) cudaEventDestroy(h->t1);
  for (auto& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (h->gstream) cudaStreamDestroy(h->gstream);
  if (h->gev) cudaEventDestroy(h->gev);
  delete h;
  return 0;
}

int kt_set_stream(kt_handle_t h, void* stream) {
  KT_CHECK_ARG(h, "kt_set_stream: null handle");
  KT_DEV_GUARD(h);
  h->stream = (cudaStream_t)stream;
  return 0;
}

int kt_sync(kt_handle_t h) {
  KT_CHECK_ARG(h, "kt_sync: null handle");
  KT_DEV_GUARD(h);
  KT_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int kt_set_rows(kt_handle_t h, int row_lo, int row_hi) {
  KT_CHECK_ARG(h, "kt_set_rows: null handle");
  KT_DEV_GUARD(h);
  KT_CHECK_ARG(0 <= row_lo && row_lo <= row_hi && row_hi <= h->NX, "kt_set_rows: need 0 <= row_lo <= row_hi <= NX");
  h->row_lo = row_lo; h->row_hi = row_hi;
  return 0;
}

long long kt_launch_count(kt_handle_t h) { return h ? h->launches : 0; }

int kt_enable_peer_access(kt_handle_t h, int peer_device) {
  KT_CHECK_ARG(h, "kt_enable_peer_access: null handle");
  KT_DEV_GUARD(h);
  if (peer_device == h->device) return 0;
  int can = 0;
  KT_CUDA(cudaDeviceCanAccessPeer(&can, h->device, peer_device));
  if (!can) { set_error("kt_enable_peer_access: device " + std::to_string(h->device) + " cannot access device " + std::to_string(peer_device)); return -4; }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
  return 0;
}

// ---- CUDA-IPC helpers for the peer-memory x-slab solver: buffers other ranks (processes) map and write into ----------
int kt_ipc_alloc(kt_handle_t h, size_t bytes, void** dev_ptr, unsigned char* handle64) {
  KT_CHECK_ARG(h && dev_ptr && handle64 && bytes > 0, "kt_ipc_alloc: bad argument");
  KT_DEV_GUARD(h);
  void* p = nullptr;
  KT_CUDA(cudaMalloc(&p, bytes));                 // a dedicated allocation: its base address IS the pointer
  KT_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t mh;
  KT_CUDA(cudaIpcGetMemHandle(&mh, p));
  static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &mh, 64);
  KT_CUDA(cudaDeviceSynchronize());
  *dev_ptr = p;
  return 0;
}

int kt_ipc_open(kt_handle_t h, const unsigned char* handle64, void** mapped) {
  KT_CHECK_ARG(h && handle64 && mapped, "kt_ipc_open: bad argument");
  KT_DEV_GUARD(h);
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle64, 64);
  KT_CUDA(cudaIpcOpenMemHandle(mapped, mh, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int kt_ipc_close(kt_handle_t h, void* mapped) {
  KT_CHECK_ARG(h, "kt_ipc_close: null handle");
  KT_DEV_GUARD(h);
  if (mapped) cudaIpcCloseMemHandle(mapped);
  return 0;
}

int kt_ipc_free(kt_handle_t h, void* dev_ptr) {
  KT_CHECK_ARG(h, "kt_ipc_free: null handle");
  KT_DEV_GUARD(h);
  if (dev_ptr) cudaFree(dev_ptr);
  return 0;
}

int kt_preload_kernels(kt_handle_t h) {
  KT_CHECK_ARG(h, "kt_preload_kernels: null handle");
  KT_DEV_GUARD(h);
  int (*const tus[])() = {kt::preload_kt_api, kt::preload_kt_krylov, kt::preload_kt_simple, kt::preload_kt_channel, kt::preload_kt_cluster,
                          kt::preload_kt_coop, kt::preload_kt_post};
  int total = 0;
  for (auto f : tus) {
    const int n = f();
    if (n < 0) { set_error("kt_preload_kernels: a kernel of this library could not be loaded on this device"); return -4; }
    total += n;
  }
  return total;
}

int kt_set_cluster(kt_handle_t h, int enable) {
  KT_CHECK_ARG(h, "kt_set_cluster: null handle");
  KT_DEV_GUARD(h);
  h->allow_cluster = enable == 1 || enable == 3;
  h->allow_coop = enable == 1 || enable == 2;
  return 0;
}

int kt_upload_field(kt_handle_t h, void* dst, const void* src_host) {
  KT_CHECK_ARG(h && dst && src_host, "kt_upload_field: null argument");
  KT_DEV_GUARD(h);
  const size_t es = esize(h);
  KT_CUDA(cudaMemcpy2DAsync(dst, h->pitch * es, src_host, h->NY * es, h->NY * es, h->NX, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

int kt_download_field(kt_handle_t h, void* dst_host, const void* src) {
  KT_CHECK_ARG(h && dst_host && src, "kt_download_field: null argument");
  KT_DEV_GUARD(h);
  const size_t es = esize(h);
  KT_CUDA(cudaMemcpy2DAsync(dst_host, h->NY * es, src, h->pitch * es, h->NY * es, h->NX, cudaMemcpyDeviceToHost, h->stream));
  KT_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int kt_coef_aos_to_soa(kt_handle_t h, void* dst_soa, const void* src_aos) {
  KT_CHECK_ARG(h && dst_soa && src_aos, "kt_coef_aos_to_soa: null argument");
  KT_DEV_GUARD(h);
  return h->dtype == KT_F64 ? coef_layout<double, true>(h, dst_soa, (void*)src_aos) : coef_layout<float, true>(h, dst_soa, (void*)src_aos);
}

int kt_coef_soa_to_aos(kt_handle_t h, void* dst_aos, const void* src_soa) {
  KT_CHECK_ARG(h && dst_aos && src_soa, "kt_coef_soa_to_aos: null argument");
  KT_DEV_GUARD(h);
  return h->dtype == KT_F64 ? coef_layout<double, false>(h, (void*)src_soa, dst_aos) : coef_layout<float, false>(h, (void*)src_soa, dst_aos);
}

// Host AoS <-> device SoA in row chunks through a caller-owned staging buffer of any size >= one row
// (NY*5 reals): H2D copy of a chunk, then the layout kernel for those rows, on the handle's stream.
int kt_upload_coef(kt_handle_t h, void* dst_soa, const void* src_host_aos, void* staging, size_t staging_bytes) {
  KT_CHECK_ARG(h && dst_soa && src_host_aos && staging, "kt_upload_coef: null argument");
  KT_DEV_GUARD(h);
  const size_t row_bytes = (size_t)h->NY * 5 * esize(h);
  const int chunk = (int)std::min<size_t>(staging_bytes / row_bytes, (size_t)h->NX);
  KT_CHECK_ARG(chunk >= 1, "kt_upload_coef: staging buffer smaller than one row of coefficients");
  for (int r = 0; r < h->NX; r += chunk) {
    const int nr = std::min(chunk, h->NX - r);
    KT_CUDA(cudaMemcpyAsync(staging, (const char*)src_host_aos + (size_t)r * row_bytes, (size_t)nr * row_bytes,
                            cudaMemcpyHostToDevice, h->stream));
    int rc = h->dtype == KT_F64 ? coef_layout<double, true>(h, dst_soa, staging, r, nr) : coef_layout<float, true>(h, dst_soa, staging, r, nr);
    if (rc) return rc;
  }
  return 0;
}

int kt_download_coef(kt_handle_t h, void* dst_host_aos, const void* src_soa, void* staging, size_t staging_bytes) {
  KT_CHECK_ARG(h && dst_host_aos && src_soa && staging, "kt_download_coef: null argument");
  KT_DEV_GUARD(h);
  const size_t row_bytes = (size_t)h->NY * 5 * esize(h);
  const int chunk = (int)std::min<size_t>(staging_bytes / row_bytes, (size_t)h->NX);
  KT_CHECK_ARG(chunk >= 1, "kt_download_coef: staging buffer smaller than one row of coefficients");
  for (int r = 0; r < h->NX; r += chunk) {
    const int nr = std::min(chunk, h->NX - r);
    int rc = h->dtype == KT_F64 ? coef_layout<double, false>(h, (void*)src_soa, staging, r, nr)
                                : coef_layout<float, false>(h, (void*)src_soa, staging, r, nr);
    if (rc) return rc;
    KT_CUDA(cudaMemcpyAsync((char*)dst_host_aos + (size_t)r * row_bytes, staging, (size_t)nr * row_bytes,
                            cudaMemcpyDeviceToHost, h->stream));
  }
  KT_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int kt_bicgstab_solve_host(kt_handle_t h, const void* coef_host_aos, const void* b_host, void* x_host, void* dev_coef_soa,
                           void* dev_staging, size_t staging_bytes, void* dev_b, void* dev_x, const kt_krylov_ws* ws,
                           double eps, long max_iter, long* iters, double* rel_res, int* status) {
  KT_CHECK_ARG(h && coef_host_aos && b_host && x_host && dev_coef_soa && dev_staging && dev_b && dev_x && ws,
               "kt_bicgstab_solve_host: null argument");
  KT_DEV_GUARD(h);
  int rc = kt_upload_coef(h, dev_coef_soa, coef_host_aos, dev_staging, staging_bytes);
  if (rc) return rc;
  rc = kt_upload_field(h, dev_b, b_host);
  if (rc) return rc;
  rc = kt_bicgstab_solve(h, dev_coef_soa, dev_b, dev_x, nullptr, ws, eps, max_iter, 0, iters, rel_res, status);
  if (rc) return rc;
  return kt_download_field(h, x_host, dev_x);
}

}  // extern "C"

// every kernel of this translation unit, for kt_preload_kernels (kt_handle.h)
namespace kt {
int preload_kt_api() {
  static const void* const fns[] = {
#define KT_PRELOAD_TU_kt_api
#include "kt_preload_list.inc"
#undef KT_PRELOAD_TU_kt_api
      nullptr};
  return preload_list(fns);
}
}  // namespace kt
