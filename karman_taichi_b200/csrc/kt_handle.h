// kt_handle.h -- host-side handle shared by the .cu translation units of libkarman_b200.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "kt_common.cuh"
#include "../../include/karman_b200.h"

struct kt_handle_s {
  int device = 0;
  int NX = 0, NY = 0, pitch = 0, dtype = KT_F64;
  int row_lo = 0, row_hi = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  // workspace (owned)
  double* partials = nullptr;        // [KT_MAX_DOTS][max_blocks]
  unsigned int* ticket = nullptr;
  int max_blocks = 0;
  kt::Scalars* sc = nullptr;         // device
  kt::Scalars* sc_host = nullptr;    // pinned, 2 slots (double-buffered polling)
  double* red_out = nullptr;         // device scratch for max-reductions [4]
  double* red_host = nullptr;        // pinned [4]
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  long long launches = 0;            // kernels launched through this handle (bench: gpu_launches)
  bool allow_cluster = true;         // small grids: whole solve inside one thread-block cluster (kt_cluster.cu)
  bool allow_coop = true;            // mid-size grids: whole solve in one persistent cooperative kernel (kt_coop.cu)
  // mid-size grids (too big for a cluster, kernels of a few microseconds): the 16-iteration chunks of the streaming
  // solvers are captured once into a CUDA graph per argument set and replayed, so the solve is not bound by the host's
  // launch rate.  Replays run on a private stream (the caller's may be the legacy default stream, which cannot be captured).
  struct GraphEntry { std::vector<const void*> key; cudaGraphExec_t exec; long long launches; unsigned long long stamp; };
  std::vector<GraphEntry> graphs;
  cudaStream_t gstream = nullptr;
  cudaEvent_t gev = nullptr;
  bool capturing = false;            // inside cudaStreamBeginCapture / EndCapture (programmatic launches are left out of graphs)
  int graphs_state = 0;              // 0 untested, 1 usable, -1 unavailable (capture refused)
  unsigned long long graph_clock = 0;
  // x-slab peer-memory solves (kt_p2p_*): solve epoch and exchange counters behind the flag sequence numbers
  unsigned long long p2p_epoch = 0, p2p_halo_k[3] = {0, 0, 0}, p2p_dot_k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  void* p2p_peers_dev = nullptr;     // device array [16] of the ranks' comm blocks (fused all-reduce in the finalising blocks)
  void* p2p_peers_host[16] = {};     // what was last uploaded into it
  unsigned long long p2p_xchg_seq = 0, p2p_red_seq = 0;      // standalone exchanges / all-reduces (kt_p2p_exchange, kt_p2p_allreduce)
  // optional per-kernel-class device timing (bench.py roofline): events bracket every launch of a class
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;  // pairs (begin, end)
  std::vector<int> prof_cls;         // class id per pair
  size_t prof_used = 0;
};

// kernel classes for kt_prof_*
enum { KT_K_BICG_P = 0, KT_K_BICG_V = 1, KT_K_BICG_ST = 2, KT_K_BICG_XR = 3, KT_K_CG_AP = 4, KT_K_CG_XR = 5, KT_K_CG_P = 6,
       KT_K_INIT = 7, KT_K_NCLASS = 8 };

namespace kt {
struct ProfScope {   // records an event pair around a launch when profiling is on
  kt_handle_s* h; int idx;
  ProfScope(kt_handle_s* h_, int cls) : h(h_), idx(-1) {
    if (h->prof_on && h->prof_used + 2 <= h->prof_ev.size()) {
      idx = (int)h->prof_used; h->prof_used += 2; h->prof_cls[idx / 2] = cls;
      cudaEventRecord(h->prof_ev[idx], h->stream);
    }
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(h->prof_ev[idx + 1], h->stream); }
};
}  // namespace kt


namespace kt {
// kt_cluster.cu: whole solve inside one thread-block cluster when every plane fits its shared memory.
// Returns 1 = solved (h->sc_host holds the final scalars), 0 = does not fit / not schedulable, < 0 = error.
template <typename T>
int cluster_try_solve(kt_handle_s* h, const void* coef, const void* b, void* x, const void* Minv, void* r_out,
                      double* history, long long hist_cap, double eps, long long max_iter, int flags, bool is_cg);
// kt_coop.cu: whole solve in one persistent cooperative kernel (one CTA per SM) for grids between the cluster limit and
// ~3e5 unknowns.  Same return convention; r_plane / v_plane are the caller's r and Ap workspace planes.
template <typename T>
int coop_try_solve(kt_handle_s* h, const void* coef, const void* b, void* x, const void* Minv, void* r_plane, void* v_plane,
                   double* history, long long hist_cap, double eps, long long max_iter, int flags, bool is_cg);
// kt_resident_plan: what the two resident solvers would launch for a geometry (1 = fits; out[0..4] = path, CTAs, rows per
// CTA, threads per CTA, elements per thread)
template <typename T> int cluster_plan(const kt_handle_s* h, bool pre, bool warm, bool is_cg, long long* out);
template <typename T> int coop_plan(const kt_handle_s* h, bool pre, bool warm, bool is_cg, long long* out);
void set_error(const std::string& s);
// n 8-byte words from device memory to pinned host memory (or device memory) by a one-warp KERNEL on the handle's stream,
// instead of cudaMemcpyAsync: copies of all streams queue up in the copy engines in submission order, so a copy that waits
// for a kernel which in turn waits for ANOTHER stream's kernel (the x-slab comm kernels with all ranks on one device) can
// block that other stream's copies for good.  Kernels of different streams have no such order.
int publish(kt_handle_s* h, const void* src_dev, void* dst, int n_words);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
}  // namespace kt

namespace kt {
// Every extern "C" entry point runs on the handle's device and leaves the caller's current device as it found it (a
// single process may drive several GPUs: Field(device='cuda:1')).
struct DeviceGuard {
  int prev = -1; bool switched = false;
  explicit DeviceGuard(int want) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != want) switched = cudaSetDevice(want) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
}  // namespace kt
#define KT_DEV_GUARD(h) kt::DeviceGuard _kt_dev_guard((h)->device)

#define KT_CUDA(expr)                                                   \
  do {                                                                  \
    cudaError_t _e = (expr);                                            \
    if (_e != cudaSuccess) return kt::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define KT_CHECK_ARG(cond, msg)                      \
  do {                                               \
    if (!(cond)) { kt::set_error(std::string("argument error: ") + msg); return -2; } \
  } while (0)

#define KT_LAUNCH_CHECK(h)                           \
  do {                                               \
    (h)->launches++;                                 \
    cudaError_t _e = cudaGetLastError();             \
    if (_e != cudaSuccess) return kt::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

// ---- kernel preloading (kt_preload_kernels) ---------------------------------------------------------------------------------
// CUDA loads device functions lazily, at their first launch, and that load may have to wait until the device is idle.  Kernels of
// this library WAIT FOR EACH OTHER across streams (the x-slab peer-memory exchange): when all ranks share one device (in-process
// multi-rank mode) a rank spinning for a neighbour whose next kernel is still being loaded would dead-lock.  Every translation
// unit lists its kernels (kt_preload_list.inc, generated from the compiled objects) and loads them up front.
#define KT_K(...) (const void*)&__VA_ARGS__,
namespace kt {
inline int preload_list(const void* const* fns) {
  int n = 0;
  for (; fns[n]; ++n) {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, fns[n]) != cudaSuccess) { cudaGetLastError(); return -1 - n; }
  }
  return n;
}
int preload_kt_api(); int preload_kt_krylov(); int preload_kt_simple(); int preload_kt_channel(); int preload_kt_cluster();
int preload_kt_coop(); int preload_kt_post();
}  // namespace kt
