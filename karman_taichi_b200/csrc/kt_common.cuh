// kt_common.cuh -- shared device helpers for libkarman_b200 (sm_100a).
//
// Design notes (DESIGN.md has the full story):
//  * Every field is a pitched (NX, pitch) plane, pitch a multiple of 128 B, so each thread can move
//    one 16-byte vector (double2 / float4) per array per row with fully coalesced, aligned accesses.
//  * Streaming operands (coefficients, vectors read once) use ld.global.nc.L1::no_allocate; the
//    stencil operand uses plain cached loads so row i+-1 / edge re-reads hit L1/L2.
//  * Inner products are deterministic: per-thread sequential accumulation over a statically assigned
//    tile list -> warp shuffle tree -> fixed-order sum over warps in shared memory -> one partial per
//    block -> the LAST block to finish (threadfence + ticket) sums the partials in index order and runs
//    the scalar algebra of the Krylov step.  No floating-point atomics anywhere.
//  * The library is compiled with -fmad=false: element-wise results are bit-identical to the NumPy
//    oracle (which cannot contract to FMA); the kernels are HBM-bound so this costs nothing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KT_WARP 32
#define KT_MAX_DOTS 2

namespace kt {

template <typename T> struct VT;
template <> struct VT<double> { static constexpr int V = 2; };
template <> struct VT<float>  { static constexpr int V = 4; };

// One 16-byte vector of reals.
template <typename T> struct __align__(16) Pack { T a[VT<T>::V]; };

template <typename T> __device__ __forceinline__ Pack<T> pack_zero() {
  Pack<T> p;
#pragma unroll
  for (int k = 0; k < VT<T>::V; ++k) p.a[k] = T(0);
  return p;
}

// ---- 16-byte global loads / stores ---------------------------------------------------------
// Memory-level parallelism is explicit: every kernel issues all the loads of a batch (U strides of an
// element-wise kernel, W column chunks x one row of the stencil kernel) back to back before the first
// dependent instruction.  The asms are volatile (KT_LD_VOLATILE=0 makes them plain asm for experiments), so
// the issue order is exactly the program order.
//   ld_stream : read-only operand, streamed once      -> ld.global.nc.L1::no_allocate
//   ld_rw     : operand this kernel also writes        -> ld.global.L1::no_allocate (no .nc)
//   ld_cached : stencil operand, re-read by neighbours -> ld.global (L1 allocating)
#ifndef KT_LD_VOLATILE
#define KT_LD_VOLATILE 1
#endif
#if KT_LD_VOLATILE
#define KT_ASM asm volatile
#else
#define KT_ASM asm
#endif
#ifndef KT_EW_UNROLL
#define KT_EW_UNROLL 4
#endif
__device__ __forceinline__ Pack<double> ld_stream(const double* p) {
  Pack<double> r;
  KT_ASM("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.a[0]), "=d"(r.a[1]) : "l"(p));
  return r;
}
__device__ __forceinline__ Pack<float> ld_stream(const float* p) {
  Pack<float> r;
  KT_ASM("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(r.a[0]), "=f"(r.a[1]), "=f"(r.a[2]), "=f"(r.a[3]) : "l"(p));
  return r;
}
__device__ __forceinline__ Pack<double> ld_rw(const double* p) {
  Pack<double> r;
  KT_ASM("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.a[0]), "=d"(r.a[1]) : "l"(p));
  return r;
}
__device__ __forceinline__ Pack<float> ld_rw(const float* p) {
  Pack<float> r;
  KT_ASM("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(r.a[0]), "=f"(r.a[1]), "=f"(r.a[2]), "=f"(r.a[3]) : "l"(p));
  return r;
}
__device__ __forceinline__ Pack<double> ld_cached(const double* p) {
  Pack<double> r;
  KT_ASM("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.a[0]), "=d"(r.a[1]) : "l"(p));
  return r;
}
__device__ __forceinline__ Pack<float> ld_cached(const float* p) {
  Pack<float> r;
  KT_ASM("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(r.a[0]), "=f"(r.a[1]), "=f"(r.a[2]), "=f"(r.a[3]) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_vec(double* p, const Pack<double>& v) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.a[0]), "d"(v.a[1]) : "memory");
}
__device__ __forceinline__ void st_vec(float* p, const Pack<float>& v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.a[0]), "f"(v.a[1]), "f"(v.a[2]), "f"(v.a[3])
               : "memory");
}

// ---- programmatic dependent launch ------------------------------------------------------------------------------------
// First statements of every streaming Krylov kernel.  When the host launches the kernel with
// cudaLaunchAttributeProgrammaticStreamSerialization (kt_krylov.cu: launch_k), its blocks may become resident while the
// previous kernel of the stream is still draining its last wave / running its finalising block; they then sit in
// griddepcontrol.wait until that kernel has completed and its writes are visible, so nothing of the previous kernel
// (vectors, scalars, tickets, flags) is touched early.  Launched the ordinary way both instructions do nothing.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- warp shuffles for double/float ------------------------------------------------------
template <typename T> __device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <typename T> __device__ __forceinline__ T shfl_down1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

// Device-resident Krylov scalars.  Lives in the handle's workspace; mirrored to pinned host memory
// when the host polls.  All in double regardless of the field dtype.
struct Scalars {
  double rho, rho_1, alpha, beta, omega;  // bicgsolver.py:15-18 / cgsolver.py:22-23
  double rTr0, rTr, old_rTr;              // initial / current / previous residual norms squared
  double dot0, dot1;                      // last finalised raw sums (unit entry points, multi-GPU)
  double eps;
  long long iter;                         // completed loop bodies
  long long max_iter;
  int status;
  int done;                               // every kernel of the solve returns at once when set
  int first;                              // 1 until the first p-update has run (p = r: bicgsolver.py:99-100)
  int flags;
  // x-slab peer-memory waits that gave up (KT_STATUS_COMM_TIMEOUT): which flag (kind << 8 | slot / group, << 16 | peer lane),
  // the sequence number it waited for and the value it last saw -- goes into kt_last_error
  unsigned long long wait_what, wait_seq, wait_seen;
};

// Kinds of scalar algebra run by the finalising block.
enum Fin {
  FIN_NONE = 0,     // just publish dot0/dot1
  FIN_BICG_INIT,    // after K0: rTr0, rho
  FIN_BICG_ALPHA,   // after v = A p_hat: alpha = rho / (r_tld, v)
  FIN_BICG_OMEGA,   // after t = A s_hat: omega = (t,s)/(t,t)
  FIN_BICG_END,     // after x/r update: convergence, rho, beta
  FIN_CG_INIT,
  FIN_CG_ALPHA,
  FIN_CG_END
};

// ---- x-slab peer-memory communication (kt_krylov.cu: "x-slab BiCGSTAB over NVLink peer memory") ------------------------
// The comm block every rank allocates (zero-initialised) and every other rank maps: the partial sums of the inner products,
// one slot per rank and slot group, and the "row has landed" flags of the halo-receive buffer.  Flags carry monotonically
// increasing sequence numbers, never reset.
constexpr int P2P_GROUPS = 8, P2P_SLOTS = 5;
struct P2PBlock {
  // all-reduce slots, [group][sending rank]: two doubles as four 8-byte words, each = 32 data bits | 32-bit tag << 32.  An
  // 8-byte store is atomic, so a word whose tag matches carries valid data: no fence on either side, no separate flag
  // (the low-latency protocol of the collective libraries, for 16 bytes per rank)
  unsigned long long ll[P2P_GROUPS][16][4];
  unsigned long long halo_flag[P2P_SLOTS][2]; // [slot][0: from lower neighbour, 1: from upper neighbour]
};

__device__ __forceinline__ bool p2p_wait(const volatile unsigned long long* flag, unsigned long long seq, Scalars* sc, long long timeout_ns,
                                         unsigned what) {
  unsigned long long t0 = 0, seen;
  while ((seen = *flag) < seq) {
    __nanosleep(32);
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) t0 = now;
    if ((long long)(now - t0) > timeout_ns) {  // a peer died or the ranks disagree on the schedule -- fail, never hang
      sc->wait_what = what; sc->wait_seq = seq; sc->wait_seen = seen;
      *(volatile int*)&sc->status = 7; *(volatile int*)&sc->done = 1;
      return false;
    }
  }
  __threadfence_system();
  return true;
}

// 16-byte load of a row another GPU wrote into this rank's receive buffer during this kernel: never through L1 / the
// read-only path (the line may have been fetched before the neighbour's store landed)
__device__ __forceinline__ Pack<double> ld_sys(const double* p) {
  Pack<double> r;
  asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.a[0]), "=d"(r.a[1]) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ Pack<float> ld_sys(const float* p) {
  Pack<float> r;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.a[0]), "=f"(r.a[1]), "=f"(r.a[2]), "=f"(r.a[3]) : "l"(p) : "memory");
  return r;
}

struct ReduceWs {
  double* partials;        // [KT_MAX_DOTS][max_blocks]
  unsigned int* ticket;    // one counter, reset by the finalising block
  int max_blocks;
  Scalars* sc;
  double* history;         // device, may be null
  long long history_cap;
  double* dots_out;        // device, may be null: raw finalised sums also land here (x-slab mode: the
                           // buffer that is all-reduced across ranks before kt_dist_scalar runs the algebra)
  // x-slab peer-memory mode, communication fused into the compute kernels (all null / 0 otherwise):
  //  * the kernel stored its boundary rows straight into the neighbours' receive buffers; once EVERY block has done so the
  //    finalising block raises the neighbours' flags hflag_lo / hflag_hi to hseq;
  //  * the finalising block is also the all-reduce: it stores this rank's sums into slot [rank] of every rank's comm block,
  //    waits for all ranks' slots in its own block, adds them in rank order and runs the scalar algebra on the global sums.
  unsigned long long* hflag_lo;
  unsigned long long* hflag_hi;
  unsigned long long hseq;
  P2PBlock* const* peers;  // device array [nranks]: every rank's comm block as mapped by this rank
  int rank, nranks, group;
  unsigned long long rseq;
  long long timeout_ns;
};

// neighbours' "row has landed" flags; called by ONE thread after every block's row stores are ordered before it
__device__ __forceinline__ void p2p_raise_halo_flags(const ReduceWs& ws) {
  if (ws.hflag_lo) *(volatile unsigned long long*)ws.hflag_lo = ws.hseq;
  if (ws.hflag_hi) *(volatile unsigned long long*)ws.hflag_hi = ws.hseq;
}

__device__ __forceinline__ void scalar_algebra(int fin, Scalars* sc, double d0, double d1, double* history,
                                               long long history_cap) {
  sc->dot0 = d0;
  sc->dot1 = d1;
  switch (fin) {
    case FIN_BICG_INIT: {
      // bicgsolver.py:90 initial_rTr; :95 rho = (r, r_tld), and r_tld == r at this point in both the
      // class version (r = r_tld = b) and the monolith (r_tld = r = b - A x0), so rho == rTr0 exactly.
      sc->rTr0 = d0;
      sc->rTr = d0;
      sc->rho = d0;
      if (history && history_cap > 0) history[0] = sqrt(d0);
      if (sc->max_iter <= 0) { sc->status = 1; sc->done = 1; }
      else if (d0 == 0.0) { sc->status = 2; sc->done = 1; }  // :96-98
      break;
    }
    case FIN_BICG_ALPHA: {
      if (d0 == 0.0) { sc->status = 3; sc->done = 1; }        // :108 ZeroDivisionError
      else sc->alpha = sc->rho / d0;
      break;
    }
    case FIN_BICG_OMEGA: {
      if (d1 == 0.0 && !(sc->flags & 2)) { sc->status = 4; sc->done = 1; }  // :118 ZeroDivisionError
      else sc->omega = d0 / d1;   // monolith (flag 2): Taichi f64 division, may yield nan/inf
      break;
    }
    case FIN_BICG_END: {
      // d0 = (r,r), d1 = (r, r_tld)
      sc->rTr = d0;
      sc->iter += 1;
      if (history && sc->iter < history_cap) history[sc->iter] = sqrt(d0);
      // The three divisions and the square root of this step are independent: issued together they overlap instead of
      // queueing behind the branches below (the resident solvers run this on their critical path).  Same operations, same
      // operands, same results as the reference's order.
      const double ratio = d0 / sc->rTr0;
      const double q_rho = d1 / sc->rho;                        // rho_new / rho_1 with rho_1 := the current rho
      const double q_ao = sc->alpha / sc->omega;
      if (sqrt(ratio) < sc->eps) { sc->status = 0; sc->done = 1; break; }   // :130
      if ((sc->flags & 2) && sc->omega == 0.0) { sc->status = 6; sc->done = 1; break; }  // karman_taichi.py:451
      if (!isfinite(d0)) { sc->status = 5; sc->done = 1; break; }
      if (sc->iter >= sc->max_iter) { sc->status = 1; sc->done = 1; break; }
      sc->rho_1 = sc->rho;                                      // :135
      sc->rho = d1;                                             // :95 of the next body
      if (d1 == 0.0) { sc->status = 2; sc->done = 1; break; }   // :96-98
      if (!(sc->flags & 2) && sc->omega == 0.0) { sc->status = 4; sc->done = 1; break; }  // :102 alpha/omega raises
      sc->beta = q_rho * q_ao;                                  // :102 (rho / rho_1) * (alpha / omega)
      break;
    }
    case FIN_CG_INIT: {
      sc->rTr0 = d0; sc->rTr = d0; sc->old_rTr = d0;            // cgsolver.py:85,89
      if (history && history_cap > 0) history[0] = sqrt(d0);
      if (sc->max_iter <= 0) { sc->status = 1; sc->done = 1; }
      break;
    }
    case FIN_CG_ALPHA: {
      if (d0 == 0.0) { sc->status = 3; sc->done = 1; }          // :96 ZeroDivisionError
      else sc->alpha = sc->old_rTr / d0;
      break;
    }
    case FIN_CG_END: {
      sc->rTr = d0;
      sc->iter += 1;
      const double beta_new = d0 / sc->old_rTr;                 // :107 (issued before the test: overlaps the square roots)
      if (sqrt(d0) < sqrt(sc->rTr0) * sc->eps) { sc->status = 0; sc->done = 1; break; }  // :102
      if (!isfinite(d0)) { sc->status = 5; sc->done = 1; break; }
      sc->beta = beta_new;
      sc->old_rTr = d0;                                         // :110
      if (history && sc->iter < history_cap) history[sc->iter] = sqrt(d0);  // :111
      if (sc->iter >= sc->max_iter) { sc->status = 1; sc->done = 1; }
      break;
    }
    default: break;
  }
}

// All-reduce of (up to) two sums over the ranks + the scalar algebra, by ONE WARP (all 32 lanes call it): lane r stores this
// rank's sums into slot [rank] of rank r's comm block as four tagged 8-byte words, then polls slot [r] of its OWN block until
// all four tags match; the values are added in rank order (the same bits on every rank) and lane 0 runs the algebra.
// Tag = 12 bits of the solve epoch | 20 bits of the exchange index: it only has to differ from what the slot held before,
// which is always the previous use of the same group.  Slot groups: see kt_krylov.cu.
__device__ __forceinline__ void p2p_allreduce_warp(P2PBlock* const* peers, int rank, int nranks, int group, int ndot, double l0, double l1,
                                                   int fin, unsigned long long seq, Scalars* sc, double* history, long long cap,
                                                   int in_solve, int op_max, long long timeout_ns) {
  const int lane = threadIdx.x & 31;
  const unsigned long long tag = (((seq >> 32) & 0xfffull) << 20 | (seq & 0xfffffull)) << 32;
  double a0 = 0.0, a1 = 0.0;
  if (lane < nranks) {
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(l0);
    const unsigned long long b1 = (unsigned long long)__double_as_longlong(ndot > 1 ? l1 : 0.0);
    unsigned long long* dst = peers[lane]->ll[group][rank];
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"((b0 & 0xffffffffull) | tag), "l"((b0 >> 32) | tag) : "memory");
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst + 2), "l"((b1 & 0xffffffffull) | tag), "l"((b1 >> 32) | tag) : "memory");
    const unsigned long long* src = peers[rank]->ll[group][lane];
    unsigned long long w0, w1, w2, w3, t0 = 0;
    for (;;) {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(src + 2) : "memory");
      const unsigned long long m = 0xffffffff00000000ull;
      if ((w0 & m) == tag && (w1 & m) == tag && (w2 & m) == tag && (w3 & m) == tag) break;
      __nanosleep(20);
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if ((long long)(now - t0) > timeout_ns) {   // a peer died or the ranks disagree on the schedule -- fail, never hang
        sc->wait_what = 2u << 8 | (unsigned)group | (unsigned)lane << 16; sc->wait_seq = seq; sc->wait_seen = w0 >> 32;
        *(volatile int*)&sc->status = 7; *(volatile int*)&sc->done = 1;
        break;
      }
    }
    a0 = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    a1 = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
  }
  double d0 = 0.0, d1 = 0.0;
  for (int r = 0; r < nranks; ++r) {            // rank order: the same bits on every rank
    const double x0 = __shfl_sync(0xffffffffu, a0, r), x1 = __shfl_sync(0xffffffffu, a1, r);
    if (op_max) { d0 = fmax(d0, x0); d1 = fmax(d1, x1); } else { d0 += x0; d1 += x1; }
  }
  if (lane == 0 && (!in_solve || !*(volatile int*)&sc->done)) scalar_algebra(fin, sc, d0, d1, history, cap);
}

// Block-level deterministic reduction of up to two running sums + last-block finalisation.
// Must be called by every thread of every block of the grid (blockDim.x multiple of 32, <= 1024).
// `pushed`: this thread stored part of a boundary row into a neighbour GPU's receive buffer (P2P kernels only).
template <int NDOT>
__device__ __forceinline__ void block_reduce_finalize(double a0, double a1, const ReduceWs& ws, int fin, bool pushed = false) {
  __shared__ double sm[KT_MAX_DOTS][32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  a0 = warp_sum(a0);
  if (NDOT > 1) a1 = warp_sum(a1);
  if (lane == 0) { sm[0][wid] = a0; if (NDOT > 1) sm[1][wid] = a1; }
  // only the few blocks that really stored peer rows pay for a system-scope fence (it costs microseconds under full memory load)
  int block_pushed = 0;
  if (ws.hflag_lo || ws.hflag_hi) block_pushed = __syncthreads_or(pushed ? 1 : 0); else __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int w = 0; w < nw; ++w) { s0 += sm[0][w]; if (NDOT > 1) s1 += sm[1][w]; }
    ws.partials[blockIdx.x] = s0;
    if (NDOT > 1) ws.partials[ws.max_blocks + blockIdx.x] = s1;
    // (rows this block stored into a neighbour GPU's receive buffer are ordered before the ticket at system scope)
    if (block_pushed) __threadfence_system(); else __threadfence();
    unsigned int t = atomicAdd(ws.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  if (ws.hflag_lo || ws.hflag_hi) {
    // every block's boundary rows have landed: tell the neighbours first, the sums below are not on their critical path
    if (threadIdx.x == 0) { __threadfence_system(); p2p_raise_halo_flags(ws); }
  } else {
    __threadfence();
  }
  // fixed-order final sum: thread k owns partials k, k+B, k+2B ...; then the same tree as above.
  // (eight independent L2 loads in flight per thread, then added in index order: the sum order is fixed, the
  // latency of the partial reads is paid once per eight instead of once per partial)
  double f0 = 0.0, f1 = 0.0;
  const int nb = (int)gridDim.x;
  for (int base = threadIdx.x; base < nb; base += 8 * blockDim.x) {
    double v0[8], v1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int b = base + k * blockDim.x;
      v0[k] = (b < nb) ? __ldcg(ws.partials + b) : 0.0;
      v1[k] = (NDOT > 1 && b < nb) ? __ldcg(ws.partials + ws.max_blocks + b) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { f0 += v0[k]; if (NDOT > 1) f1 += v1[k]; }
  }
  f0 = warp_sum(f0);
  if (NDOT > 1) f1 = warp_sum(f1);
  __syncthreads();
  if (lane == 0) { sm[0][wid] = f0; if (NDOT > 1) sm[1][wid] = f1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0.0, s1 = 0.0;
    for (int w = 0; w < nw; ++w) { s0 += sm[0][w]; if (NDOT > 1) s1 += sm[1][w]; }
    *ws.ticket = 0u;
    if (ws.dots_out) { ws.dots_out[0] = s0; if (NDOT > 1) ws.dots_out[1] = s1; }
    if (ws.nranks > 1) { sm[0][0] = s0; sm[1][0] = s1; }
    else scalar_algebra(fin, ws.sc, s0, s1, ws.history, ws.history_cap);
  }
  if (ws.nranks > 1) {     // fused all-reduce over the ranks: warp 0 of the finalising block
    __syncthreads();
    if (wid == 0) p2p_allreduce_warp(ws.peers, ws.rank, ws.nranks, ws.group, NDOT, sm[0][0], NDOT > 1 ? sm[1][0] : 0.0, fin, ws.rseq, ws.sc,
                                     ws.history, ws.history_cap, 1, 0, ws.timeout_ns);
  }
}

// The neighbours' receive rows for the plane an element-wise kernel writes (flat vector index -> row test by offset).
template <typename T> struct HaloPush {
  T* lo; T* hi;                     // null: no neighbour on that side / not pushing
  long long first_off, last_off;    // element offsets of this rank's first / last owned row in the plane
  int pitch;
};
template <typename T> __device__ __forceinline__ bool ew_push(const HaloPush<T>& hp, size_t o, const Pack<T>& v) {
  bool pushed = false;
  if (hp.lo) { const long long d = (long long)o - hp.first_off; if (d >= 0 && d < hp.pitch) { st_vec(hp.lo + d, v); pushed = true; } }
  if (hp.hi) { const long long d = (long long)o - hp.last_off; if (d >= 0 && d < hp.pitch) { st_vec(hp.hi + d, v); pushed = true; } }
  return pushed;
}

// block max -> partials[block] -> last block takes the max (order-independent, exact)
__device__ __forceinline__ void block_max_finalize(double m0, double m1, double* partials, int max_blocks,
                                                   unsigned int* ticket, double* out) {
  __shared__ double sm[2][32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  m0 = warp_max(m0); m1 = warp_max(m1);
  if (lane == 0) { sm[0][wid] = m0; sm[1][wid] = m1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < nw; ++w) { a = fmax(a, sm[0][w]); b = fmax(b, sm[1][w]); }
    partials[bid] = a; partials[max_blocks + bid] = b;
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == (unsigned)nblocks - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double a = 0.0, b = 0.0;
  for (int k = threadIdx.x; k < nblocks; k += blockDim.x) { a = fmax(a, __ldcg(partials + k)); b = fmax(b, __ldcg(partials + max_blocks + k)); }
  a = warp_max(a); b = warp_max(b);
  __syncthreads();
  if (lane == 0) { sm[0][wid] = a; sm[1][wid] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = 0.0; b = 0.0;
    for (int w = 0; w < nw; ++w) { a = fmax(a, sm[0][w]); b = fmax(b, sm[1][w]); }
    out[0] = a; out[1] = b;
    *ticket = 0u;
  }
}


}  // namespace kt
