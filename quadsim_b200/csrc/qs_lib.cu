// qs_lib.cu -- CUDA kernels (sm_100a) and the C ABI of libquadsim_b200.so (include/quadsim_b200.h).
//
// One thread per environment.  Persistent per-env state lives in HBM as structure-of-arrays of 16-byte rows
// (float4 for fp32, double2 for fp64) so every load/store instruction of a warp covers one contiguous 512-byte
// span; the whole step (target controller -> 2x Drone.step -> dock ports -> state2rel -> flags/shaping/reward ->
// auto-reset with Philox) runs in registers between one set of loads and one set of stores.
//
// There is NO CPU fallback: qs_create fails with QS_ERR_NO_DEVICE when no CUDA device is usable.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/quadsim_b200.h"
#include "qs_core.cuh"
#include "qs_env_kernels.cuh"

namespace {

constexpr int64_t kHostChunk = 1 << 17;  // envs per pipelined chunk in qs_step_host
constexpr int kHostStreams = 4;

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define QS_CUDA(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail(QS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

}  // namespace
// shared with qs_policy.cu: set the thread-local error string and return `code`
int qs_fail_(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
namespace {

// NVTX range around every entry point that enqueues work ("NVTX ranges around step"); a no-op costing a
// few ns unless a profiler is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

using StepArgs = EnvStepArgs;  // qs_env_kernels.cuh

template <typename T, int KIND, typename P>
__global__ void __launch_bounds__(kBlock, min_blocks<T>()) dock_step_kernel(const __grid_constant__ StepArgs a, const __grid_constant__ P prm) {
  pdl_prologue();
  constexpr int NW = dock_words<T>();
  const int64_t i = a.first + (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const bool active = i < a.first + a.n;
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  T w[NW], act[4];
  load_words<T, NW>(a.state, i, w);
  load_action<T>(a.actions, i, act);
  const StepTail tail = dock_step_loaded<T, KIND, dock_spec<T, P>()>(a, prm, i, w, act);
  warp_stats_update(a.warp_stats, i, mask, tail.done, tail.docking, tail.nonfinite, tail.ret, tail.len);
}

// =========================================================================================================
// Persistent, bulk-async (TMA 1-D) pipelined variant of the docking step: the default for fp64, an A/B option for fp32.
//   * grid = SMs x resident CTAs; each CTA walks 128-env tiles handed out by a global atomic counter;
//   * one elected thread issues two cp.async.bulk (SASS UBLKCP) copies -- the tile's 20 KB state block (contiguous in the
//     tile-major layout) and its 2 KB of actions -- into shared memory, completion counted on an mbarrier;
//   * the tile for iteration j+1 is requested as soon as every thread has moved tile j from shared memory to registers,
//     so its HBM latency is covered by the ~1k-instruction compute phase of tile j: registers are the second buffer and
//     ONE shared stage per CTA suffices (fp32: 22 KB x 6 CTAs/SM; fp64: 44 KB x 3 CTAs/SM = 132 KB in flight per SM).
// Results go straight from registers to HBM (coalesced 16-byte stores), as in the direct kernel.
// =========================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
}

template <typename T> constexpr int pipe_rows() { return (dock_words<T>() + 4) / Row<T>::W; }  // state rows + action rows (16 B each)
template <typename T> constexpr int pipe_smem_bytes() { return pipe_rows<T>() * kBlock * 16 + 64; }

// (A warp-granular version -- every warp its own stage, mbarrier and work counter, no CTA barrier, a unit = 20 copies of
// 512 B -- was measured slower: 158 vs 134 us fp64, 83 vs 73 us fp32.  One 40 KB copy per tile is what the copy engine likes.)
template <typename T, int KIND, typename P>
__global__ void __launch_bounds__(kBlock, min_blocks<T>()) dock_step_pipe_kernel(const __grid_constant__ StepArgs a, const __grid_constant__ P prm, unsigned int* sched) {
  using V = typename Row<T>::V;
  constexpr int NW = dock_words<T>(), W = Row<T>::W, NR = NW / W, AR = 4 / W;  // AR: 16-byte rows per action
  pdl_prologue();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  V* buf = reinterpret_cast<V*>(smem_raw);                                   // [NR + AR][kBlock]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (NR + AR) * kBlock * 16);
  volatile int* s_tile = reinterpret_cast<volatile int*>(full + 1);         // [2]: tile index for iteration parity
  const int tid = threadIdx.x;
  const int tiles = (int)((a.n + kBlock - 1) / kBlock);
  const V* state = reinterpret_cast<const V*>(a.state);
  const V* actions = reinterpret_cast<const V*>(a.actions);

  auto request = [&](int tile) {  // elected thread: one mbarrier phase = two bulk copies
    const int64_t env0 = a.first + (int64_t)tile * kBlock;
    const int64_t left = a.first + a.n - env0;
    const uint32_t cnt = (uint32_t)(left < kBlock ? left : kBlock);
    mbar_expect_tx(full, NR * kBlock * 16 + cnt * AR * 16);
    bulk_g2s(buf, state + env0 * NR, NR * kBlock * 16, full);  // tile-major layout: the tile's state is one contiguous span
    bulk_g2s(buf + NR * kBlock, actions + env0 * AR, cnt * AR * 16, full);
  };

  int next = -1;  // thread 0 only: tile whose data is in flight / landed
  if (tid == 0) {
    mbar_init(full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int t0 = (int)atomicAdd(sched, 1u);
    s_tile[0] = t0 < tiles ? t0 : -1;
    if (t0 < tiles) {
      request(t0);
      next = (int)atomicAdd(sched, 1u);
    }
  }
  __syncthreads();

  for (int it = 0;; it++) {
    const int tile = s_tile[it & 1];
    if (tile < 0) break;
    mbar_wait(full, it & 1);
    const int64_t i = a.first + (int64_t)tile * kBlock + tid;
    const bool active = i < a.first + a.n;
    T w[NW];
#pragma unroll
    for (int r = 0; r < NR; r++) unpack(buf[r * kBlock + tid], &w[r * W]);
    T act[4];
#pragma unroll
    for (int r = 0; r < AR; r++) unpack(buf[NR * kBlock + tid * AR + r], &act[r * W]);
    if (tid == 0) s_tile[(it + 1) & 1] = next < tiles ? next : -1;
    __syncthreads();  // every thread has drained the stage (and sees the next tile index)
    if (tid == 0 && next >= 0 && next < tiles) {
      request(next);
      next = (int)atomicAdd(sched, 1u);  // latency hidden behind this tile's compute
    }
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    if (active) {
      const StepTail tail = dock_step_loaded<T, KIND, dock_spec<T, P>()>(a, prm, i, w, act);
      warp_stats_update(a.warp_stats, i, mask, tail.done, tail.docking, tail.nonfinite, tail.ret, tail.len);
    }
  }
  // self-resetting scheduler: the last CTA out clears the counters for the next launch on this stream
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(sched + 1, 1u) == gridDim.x - 1) { sched[0] = 0; sched[1] = 0; }
  }
}

// =========================================================================================================
// hovering-v0 step: gym_docking/envs/hovering_env.py:47-78 + auto-reset to the per-env initial state (:80-82)
// =========================================================================================================
template <typename T, typename P>
__global__ void __launch_bounds__(kBlock) hover_step_kernel(const __grid_constant__ StepArgs a, const __grid_constant__ P prm) {
  pdl_prologue();
  __shared__ T stage[kBlock * 13];
  const int64_t tile0 = a.first + (int64_t)blockIdx.x * kBlock;
  const int64_t i = tile0 + threadIdx.x;
  const bool active = i < a.first + a.n;
  const int64_t left = a.first + a.n - tile0;
  const int cnt = (int)(left < kBlock ? left : kBlock);
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  T w[kHoverWords];
  T obs[13];
  HoverEnv<T> e;
  StepOut<T> o;
  o.done = o.docking = o.nonfinite = false;
  double stat_ret = 0.0;
  int stat_len = 0;
  if (active) {
    load_words<T, kHoverWords>(a.state, i, w);
    T act[4];
    load_action<T>(a.actions, i, act);
    words_to_hover<T>(w, e);
    hover_env_step<T>(prm, e, act, obs, o);
    reinterpret_cast<T*>(a.reward)[i] = o.reward;
    a.done[i] = o.done ? 1 : 0;
    if (a.info) a.info[i] = (uint8_t)o.info();
    stat_ret = (double)e.ep_return;
    stat_len = e.tt;
  }
  if (a.term_obs) store_obs13_tile<T>(stage, a.term_obs, tile0, cnt, obs, active);
  if (active && o.done) {
    if (a.term_chaser) store_terminal_state<T>(a.term_chaser, i, e.d, 0.0, 0.0, 0.0);
    if (a.auto_reset) {
      e.episode += 1;
      hover_env_seat<T>(prm, e, a.rng, i, obs);
    }
  }
  store_obs13_tile<T>(stage, a.obs, tile0, cnt, obs, active);
  if (active) {
    hover_to_words<T>(e, w);
    store_words<T, kHoverWords>(a.state, i, w);
    warp_stats_update(a.warp_stats, i, mask, o.done, o.docking, o.nonfinite, stat_ret, stat_len);
  }
}

// =========================================================================================================
// T consecutive steps in ONE launch (qs_step_n): open-loop stepping with pre-generated actions[T,N,4] (BASELINE config 2,
// random-action rollouts, replaying an action tape).  Each thread loads its env once, keeps the whole state in registers
// across the T steps (no state traffic between steps, no launch per step) and stores it once; per step it reads 4 action
// words and writes the step's observation / reward / done.  Same arithmetic (qs_core.cuh) as the single-step kernels.
// =========================================================================================================
template <typename T, int KIND, typename P>
__global__ void __launch_bounds__(kBlock, min_blocks<T>()) multi_step_kernel(const __grid_constant__ StepArgs a, const __grid_constant__ P prm, int64_t n_steps) {
  constexpr bool HOVER = KIND == QS_HOVER;
  constexpr int NW = HOVER ? kHoverWords : dock_words<T>();
  constexpr int OD = HOVER ? 13 : 12;
  __shared__ T stage[HOVER ? kBlock * 13 : 1];
  const int64_t tile0 = (int64_t)blockIdx.x * kBlock;
  const int64_t i = tile0 + threadIdx.x;
  const bool active = i < a.n;
  const int64_t left = a.n - tile0;
  const int cnt = (int)(left < kBlock ? left : kBlock);
  const unsigned mask = __ballot_sync(0xffffffffu, active);
  T w[NW];
  DockEnv<T> de;
  HoverEnv<T> he;
  if (active) {
    load_words<T, NW>(a.state, i, w);
    if constexpr (HOVER) words_to_hover<T>(w, he); else words_to_dock<T, NW>(w, de);
  }
  for (int64_t t = 0; t < n_steps; t++) {
    const int64_t row0 = t * a.n;  // outputs and actions are [T, N, ...]
    T obs[OD];
    StepOut<T> o;
    o.done = o.docking = o.nonfinite = false;
    double stat_ret = 0.0;
    int stat_len = 0;
    if (active) {
      T act[4];
      load_action<T>(a.actions, row0 + i, act);
      if constexpr (HOVER) {
        hover_env_step<T>(prm, he, act, obs, o);
        stat_ret = (double)he.ep_return; stat_len = he.tt;
      } else {
        dock_env_step<T, KIND>(prm, de, act, obs, o);
        stat_ret = (double)de.last_shaping; stat_len = de.tt;
      }
      reinterpret_cast<T*>(a.reward)[row0 + i] = o.reward;
      a.done[row0 + i] = o.done ? 1 : 0;
      if (o.done && a.auto_reset) {
        if constexpr (HOVER) { he.episode += 1; hover_env_seat<T>(prm, he, a.rng, i, obs); }
        else { de.episode += 1; dock_env_seat<T>(prm, de, a.rng, i, obs); }
      }
    }
    if constexpr (HOVER) store_obs13_tile<T>(stage, a.obs, row0 + tile0, cnt, obs, active);
    else if (active) store_row<T, 12>(a.obs, row0 + i, obs);
    if (active) warp_stats_update(a.warp_stats, i, mask, o.done, o.docking, o.nonfinite, stat_ret, stat_len);
  }
  if (active) {
    if constexpr (HOVER) hover_to_words<T>(he, w); else dock_to_words<T, NW>(de, w);
    store_words<T, NW>(a.state, i, w);
  }
}

// =========================================================================================================
// construct / reset kernels (VecEnv.reset(): docking_env.py:233-244, hovering_env.py:80-82)
//   mode 0: construct -- episode 0, des_q = identity (docking_env.py:60-63), u = 0
//   mode 1: reset()   -- envs with t != 0 start their next episode; des_q survives (Q4)
// =========================================================================================================
template <typename T, int KIND, typename P>
__global__ void __launch_bounds__(kBlock) seat_kernel(const __grid_constant__ StepArgs a, const __grid_constant__ P prm, int mode) {
  const int64_t i = a.first + (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= a.first + a.n) return;
  if (KIND == QS_HOVER) {
    T w[kHoverWords];
    HoverEnv<T> e;
    e.episode = 0;
    if (mode == 1) {
      load_words<T, kHoverWords>(a.state, i, w);
      words_to_hover<T>(w, e);
      e.episode += e.tt != 0 ? 1 : 0;
    }
    T obs[13];
    hover_env_seat<T>(prm, e, a.rng, i, obs);
    hover_to_words<T>(e, w);
    store_words<T, kHoverWords>(a.state, i, w);
    if (a.obs) store_obs13<T>(a.obs, i, obs);
  } else {
    constexpr int NW = dock_words<T>();
    T w[NW];
    DockEnv<T> e;
    e.episode = 0;
    e.dq = Quat<T>{T(1), T(0), T(0), T(0)};
    if (mode == 1) {
      load_words<T, NW>(a.state, i, w);
      words_to_dock<T, NW>(w, e);
      e.episode += e.tt != 0 ? 1 : 0;
    }
    T obs[12];
    dock_env_seat<T>(prm, e, a.rng, i, obs);
    dock_to_words<T, NW>(e, w);
    store_words<T, NW>(a.state, i, w);
    if (a.obs) store_row<T, 12>(a.obs, i, obs);
  }
}

// =========================================================================================================
// state export / import (array-of-struct views <-> row layout)
// =========================================================================================================
struct StateArrays { void *chaser, *target, *uc, *ut, *desq, *last_shaping, *ep_return; int32_t *t, *episode; };

template <typename T, bool HOVER, bool IMPORT>
__global__ void __launch_bounds__(kBlock) state_io_kernel(void* state, int64_t stride, int64_t n, StateArrays s, double ox, double oy, double oz) {
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  constexpr int NW = HOVER ? kHoverWords : dock_words<T>();
  T w[NW];
  load_words<T, NW>(state, i, w);
  // word index of state[k] (k = 0..12) relative to a drone's base word; positions (k < 3) are stored relative to the handle's
  // origin (fp32 docking: the target's start position; otherwise 0)
  const int sidx[13] = {0, 1, 2, 4, 5, 6, 8, 9, 10, 11, 3, 7, -1};
  const double org[3] = {ox, oy, oz};
  double* ch = reinterpret_cast<double*>(s.chaser);
  double* tg = reinterpret_cast<double*>(s.target);
  double* uc = reinterpret_cast<double*>(s.uc);
  double* ut = reinterpret_cast<double*>(s.ut);
  double* dq = reinterpret_cast<double*>(s.desq);
  double* ls = reinterpret_cast<double*>(s.last_shaping);
  double* er = reinterpret_cast<double*>(s.ep_return);
  constexpr int WZ_C = HOVER ? 16 : 36, WZ_T = 37;
#pragma unroll
  for (int k = 0; k < 13; k++) {
    const double o = k < 3 ? org[k] : 0.0;
    const int wc = k == 12 ? WZ_C : sidx[k];
    if (ch) { if (IMPORT) w[wc] = (T)(ch[i * 13 + k] - o); else ch[i * 13 + k] = (double)w[wc] + o; }
    if (!HOVER && tg) {
      const int wt = k == 12 ? WZ_T : 16 + sidx[k];
      if (IMPORT) w[wt] = (T)(tg[i * 13 + k] - o); else tg[i * 13 + k] = (double)w[wt] + o;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (uc) { if (IMPORT) w[12 + k] = (T)uc[i * 4 + k]; else uc[i * 4 + k] = (double)w[12 + k]; }
    if (!HOVER && ut) { if (IMPORT) w[28 + k] = (T)ut[i * 4 + k]; else ut[i * 4 + k] = (double)w[28 + k]; }
    if (!HOVER && dq) { if (IMPORT) w[32 + k] = (T)dq[i * 4 + k]; else dq[i * 4 + k] = (double)w[32 + k]; }
  }
  if (!HOVER && ls) { if (IMPORT) w[38] = (T)ls[i]; else ls[i] = (double)w[38]; }
  int32_t tt, ep;
  if (HOVER) {
    if (er) { if (IMPORT) w[17] = (T)er[i]; else er[i] = (double)w[17]; }
    tt = word_to_int(w[18]); ep = word_to_int(w[19]);
  } else {
    if (er && !IMPORT) er[i] = (double)w[38];  // docking: the running return IS last_shaping; nothing to import
    unpack_counters(w[39], tt, ep);
  }
  if (s.t) { if (IMPORT) tt = s.t[i]; else s.t[i] = tt; }
  if (s.episode) { if (IMPORT) ep = s.episode[i]; else s.episode[i] = ep; }
  if (IMPORT) {
    if (HOVER) { w[18] = int_to_word<T>(tt); w[19] = int_to_word<T>(ep); }
    else w[39] = pack_counters<T>(tt, ep);
  }
  if (IMPORT) store_words<T, NW>(state, i, w);
}

// fp32 docking handles pack t (10 bits) and episode (22 bits) into one state word: reject what would not round-trip
__global__ void __launch_bounds__(256) counter_range_kernel(int64_t n, const int32_t* t, const int32_t* ep, int* flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool bad = false;
  if (t) bad |= t[i] < 0 || t[i] > 1023;
  if (ep) bad |= ep[i] < 0 || ep[i] >= (1 << 22);
  if (bad) *flag = 1;
}

// statistics: fixed-order two-phase reduction of the per-warp slots (then cleared): kStatBlocks partial sums, then one
// block adds the partials in index order -- bit-reproducible whatever the launch geometry of the step kernels was
constexpr int kStatBlocks = 64;
__global__ void __launch_bounds__(256) stats_partial_kernel(double* warp_stats, int64_t n_slots, double* partial) {
  __shared__ double sm[5][256];
  double acc[5] = {0, 0, 0, 0, 0};
  for (int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x; s < n_slots; s += (int64_t)kStatBlocks * 256) {
    double* p = warp_stats + s * kStatWords;
#pragma unroll
    for (int k = 0; k < 5; k++) { acc[k] += p[k]; p[k] = 0.0; }
  }
#pragma unroll
  for (int k = 0; k < 5; k++) sm[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
#pragma unroll
      for (int k = 0; k < 5; k++) sm[k][threadIdx.x] += sm[k][threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x < 5) partial[blockIdx.x * 8 + threadIdx.x] = sm[threadIdx.x][0];
}
__global__ void stats_final_kernel(const double* partial, double* out) {
  if (threadIdx.x < 5) {
    double a = 0.0;
    for (int b = 0; b < kStatBlocks; b++) a += partial[b * 8 + threadIdx.x];
    out[threadIdx.x] = a;
  }
}

// =========================================================================================================
// leaf-function kernels (array-of-struct in/out; same device functions as the fused kernels)
// =========================================================================================================
template <typename T> __device__ __forceinline__ Drone<T> load_state13(const T* s, const T* u) {
  Drone<T> d = {s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], s[8], s[9], s[10], s[11], s[12], T(0), T(0), T(0), T(0)};
  if (u) { d.u0 = u[0]; d.u1 = u[1]; d.u2 = u[2]; d.u3 = u[3]; }
  return d;
}

template <typename T>
__global__ void fn_kernel(int fn, int64_t n, const T* in0, const T* in1, const T* in2, T* out0, T* out1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  switch (fn) {
    case QS_FN_QUAT2EULER: {
      Euler<T> e = quat2euler(in0[4 * i], in0[4 * i + 1], in0[4 * i + 2], in0[4 * i + 3]);
      out0[3 * i] = e.phi; out0[3 * i + 1] = e.theta; out0[3 * i + 2] = e.psi;
    } break;
    case QS_FN_EULER2QUAT: {
      Quat<T> q = euler2quat(in0[3 * i], in0[3 * i + 1], in0[3 * i + 2]);
      out0[4 * i] = q.w; out0[4 * i + 1] = q.x; out0[4 * i + 2] = q.y; out0[4 * i + 3] = q.z;
    } break;
    case QS_FN_QUAT2ROT: {
      Rot<T> R = quat2rot(in0[4 * i], in0[4 * i + 1], in0[4 * i + 2], in0[4 * i + 3]);
      T* o = out0 + 9 * i;
      o[0] = T(1); o[1] = R.r01; o[2] = R.r02; o[3] = R.r10; o[4] = T(1); o[5] = R.r12; o[6] = R.r20; o[7] = R.r21; o[8] = T(1);
    } break;
    case QS_FN_ROT2EULER: {
      const T* R = in0 + 9 * i;
      Euler<T> e = rot2euler(R[5], R[3], R[4], R[2], R[8]);
      out0[3 * i] = e.phi; out0[3 * i + 1] = e.theta; out0[3 * i + 2] = e.psi;
    } break;
    case QS_FN_U_LIMIT: {
      u_limit(in0[4 * i], in0[4 * i + 1], in0[4 * i + 2], in0[4 * i + 3], out0[4 * i], out0[4 * i + 1], out0[4 * i + 2], out0[4 * i + 3]);
    } break;
    case QS_FN_DRONE_STEP: {
      Drone<T> d = load_state13(in0 + 13 * i, in1 + 4 * i);
      drone_step(d, in2[4 * i], in2[4 * i + 1], in2[4 * i + 2], in2[4 * i + 3]);
      drone_to_obs13(d, out0 + 13 * i);
      out1[4 * i] = d.u0; out1[4 * i + 1] = d.u1; out1[4 * i + 2] = d.u2; out1[4 * i + 3] = d.u3;
    } break;
    case QS_FN_PID:
    case QS_FN_VEL_CTRL: {
      const T* des = in0 + 13 * i;
      Drone<T> now = load_state13(in1 + 13 * i, (const T*)nullptr);
      Quat<T> dq = {des[6], des[7], des[8], des[9]};
      T F, M0, M1, M2;
      if (fn == QS_FN_PID) pid(des[0], des[1], des[2], des[3], des[4], des[5], dq, now, F, M0, M1, M2);
      else vel_controller(des[3], des[4], des[5], dq, now, F, M0, M1, M2);
      out0[4 * i] = F; out0[4 * i + 1] = M0; out0[4 * i + 2] = M1; out0[4 * i + 3] = M2;
      out1[4 * i] = dq.w; out1[4 * i + 1] = dq.x; out1[4 * i + 2] = dq.y; out1[4 * i + 3] = dq.z;
    } break;
    case QS_FN_STATE2REL: {
      Drone<T> c = load_state13(in0 + 13 * i, (const T*)nullptr), t = load_state13(in1 + 13 * i, (const T*)nullptr);
      T rel[12];
      state2rel(c, t, rel);
#pragma unroll
      for (int k = 0; k < 12; k++) out0[12 * i + k] = rel[k];
    } break;
    case QS_FN_MATH: {
      const T y = in0[2 * i], x = in0[2 * i + 1];
      T sn, cs;
      m_sincos(y, &sn, &cs);
      T* o = out0 + 6 * i;
      o[0] = m_atan2(y, x); o[1] = sn; o[2] = cs; o[3] = m_rsqrt(m_abs(x)); o[4] = m_sqrt(m_abs(x)); o[5] = m_asin(y);
    } break;
    default: break;
  }
}

// Runner._run's GAE loop, rl_baselines/ppo2/ppo2.py:511-521.  numpy's promotion rules make the reference evaluate it as:
// gamma*nextvalues in float32 (python scalar x float32 array), everything after that in float64 (nextnonterminal =
// 1.0 - bool array is float64), advs stored as float32, returns = advs + values in float32.  Mirrored exactly.
__global__ void __launch_bounds__(256) gae_kernel(int64_t T, int64_t n, const float* __restrict__ rew, const float* __restrict__ val,
                                                  const uint8_t* __restrict__ dones, const float* __restrict__ last_val,
                                                  const uint8_t* __restrict__ last_done, double gamma, double lam,
                                                  float* __restrict__ adv, float* __restrict__ ret) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double nextnonterminal = 1.0 - (double)last_done[i];
  float nextvalues = last_val[i];
  double lastgaelam = 0.0;
  const double gl = gamma * lam;
  const float gf = (float)gamma;
  // The loads of a step do not depend on the recurrence: eight steps' worth are requested up front, then the recurrence (a dozen
  // dependent fp64 operations per step) runs over the registers.  What the plain loop waited for were its three DRAM round trips
  // per step.
  int64_t t = T;
  while (t > 0) {
    const int m = t >= 8 ? 8 : (int)t;  // steps t-1, t-2, .., t-m
    float vv[8], rr[8];
    uint8_t dd[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k < m) {
        const int64_t o = (t - 1 - k) * n + i;
        vv[k] = val[o]; rr[k] = rew[o]; dd[k] = dones[o];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k < m) {
        const int64_t o = (t - 1 - k) * n + i;
        const float v = vv[k];
        const double delta = __dadd_rn(__dadd_rn((double)rr[k], __dmul_rn((double)__fmul_rn(gf, nextvalues), nextnonterminal)), -(double)v);
        lastgaelam = __dadd_rn(delta, __dmul_rn(__dmul_rn(gl, nextnonterminal), lastgaelam));
        const float a = (float)lastgaelam;
        adv[o] = a;
        ret[o] = __fadd_rn(a, v);
        nextnonterminal = 1.0 - (double)dd[k];
        nextvalues = v;
      }
    }
    t -= m;
  }
}

// PID expert for the chaser (qs::pid_expert_*; run_expert_policy.py:44-69, run_expert_record.py:40-139): one thread per env,
// reads the env state rows, keeps its own [N,8] memory (des vel xyz, pad, des q wxyz).  In alias mode the first step of an
// episode re-seats the chaser from that memory (the scripts' state_des IS env.chaser_ini_state) and rewrites the reset
// observation the caller is about to record.
template <typename T, typename P>
__global__ void __launch_bounds__(kBlock) expert_kernel(void* state, int64_t n, T* __restrict__ mem, int alias, T* __restrict__ actions,
                                                        void* obs, const __grid_constant__ P prm) {
  constexpr int NW = dock_words<T>();
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  T w[NW];
  load_words<T, NW>(state, i, w);
  DockEnv<T> e;
  words_to_dock<T, NW>(w, e);
  ExpertMem<T> m = {mem[8 * i], mem[8 * i + 1], mem[8 * i + 2], Quat<T>{mem[8 * i + 4], mem[8 * i + 5], mem[8 * i + 6], mem[8 * i + 7]}};
  const bool first = e.tt == 0;
  const bool reseated = pid_expert_reseat(e.c, m, first, alias != 0);
  if (reseated && obs) {  // env.reset()'s observation of the re-seated chaser (docking_env.py:240-244)
    T rel[12];
    state2rel(prm, e.c, e.t, rel);
    store_row<T, 12>(obs, i, rel);
  }
  T act[4];
  pid_expert_cmd(prm, e.c, e.t, m, first, alias != 0, act);
  if (reseated) {
    dock_to_words<T, NW>(e, w);
    store_words<T, NW>(state, i, w);
  }
  mem[8 * i] = m.dvx; mem[8 * i + 1] = m.dvy; mem[8 * i + 2] = m.dvz; mem[8 * i + 3] = T(0);
  mem[8 * i + 4] = m.dq.w; mem[8 * i + 5] = m.dq.x; mem[8 * i + 6] = m.dq.y; mem[8 * i + 7] = m.dq.z;
#pragma unroll
  for (int k = 0; k < 4; k++) actions[4 * i + k] = act[k];
}

// A few envs' drone states for an observer (the Unity bridge, server/pub_server.py:25-36): gather k env ids into
// [k,13] float64 world-coordinate rows (chaser / single drone, and target for docking).  t_out: env.t (the sim step).
template <typename T, bool HOVER>
__global__ void peek_state_kernel(const void* state, int64_t n, const int64_t* ids, int64_t k, double* chaser, double* target, int32_t* t_out,
                                  double ox, double oy, double oz) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const int64_t i = ids[j];
  if (i < 0 || i >= n) return;
  constexpr int NW = HOVER ? kHoverWords : dock_words<T>();
  T w[NW];
  load_words<T, NW>(state, i, w);
  const int sidx[13] = {0, 1, 2, 4, 5, 6, 8, 9, 10, 11, 3, 7, -1};
  const double org[3] = {ox, oy, oz};
#pragma unroll
  for (int c = 0; c < 13; c++) {
    const double o = c < 3 ? org[c] : 0.0;
    chaser[j * 13 + c] = (double)w[c == 12 ? (HOVER ? 16 : 36) : sidx[c]] + o;
    if (!HOVER && target) target[j * 13 + c] = (double)w[c == 12 ? 37 : 16 + sidx[c]] + o;
  }
  if (t_out) {
    int32_t tt, ep;
    if (HOVER) tt = word_to_int(w[18]); else unpack_counters(w[NW - 1], tt, ep);
    t_out[j] = tt;
  }
}

__global__ void philox_kernel(int64_t n, const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  U4 c = {ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]};
  U4 r = philox4x32_10(c, key[2 * i], key[2 * i + 1]);
  out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

}  // namespace

// =========================================================================================================
// host side
// =========================================================================================================
struct qs_env {
  qs_config_t cfg;
  int obs_dim;
  int n_rows;
  int64_t stride;
  size_t esize;
  void* state = nullptr;
  double* warp_stats = nullptr;
  int64_t n_slots = 0;
  double* stats_dev = nullptr;
  double* stats_host = nullptr;  // pinned
  float jitter_scale;
  int use_pipe = 0;               // persistent bulk-async docking kernel (default for fp64)
  int pipe_grid = 0;
  unsigned int* sched = nullptr;  // [kHostStreams + 1][2] tile counters (one pair per stream slot)
  int* import_flag = nullptr;     // device word set by the import validation kernel (fp32 counter ranges)
  // model parameters: custom = handle created with use_params (kernels read prm_f / prm_d from the constant bank); origin =
  // what the state rows' positions are relative to (fp32 docking: the target's start position; otherwise 0)
  bool custom = false;
  qs_env_params_t params;
  RuntimeParams<float> prm_f;
  RuntimeParams<double> prm_d;
  double origin[3] = {0, 0, 0};
  // qs_step_host pipeline
  cudaEvent_t ev_fork = nullptr;  // recorded on the caller's stream, waited on by every internal stream
  cudaStream_t hstreams[kHostStreams] = {};
  void *h_act = nullptr, *h_obs = nullptr, *h_rew = nullptr;
  uint8_t* h_done = nullptr;
};

namespace {

StepArgs make_args(qs_env* e, int64_t first, int64_t n) {
  StepArgs a;
  memset(&a, 0, sizeof a);
  a.state = e->state; a.stride = e->stride; a.first = first; a.n = n;
  a.warp_stats = e->warp_stats;
  a.rng = RngCfg{e->cfg.seed, e->cfg.env_id0, e->cfg.randomize, e->jitter_scale};
  a.auto_reset = e->cfg.auto_reset;
  return a;
}

// ---- dispatch on (dtype, env kind, compiled-in vs runtime parameters): f(TypeTag<T>, IntTag<KIND>, prm) with prm the
// kernel-parameter object of the model constants
template <typename T> struct TypeTag { using type = T; };
template <int V> struct IntTag { static constexpr int value = V; };
template <typename T> const RuntimeParams<T>& runtime_prm(const qs_env* e);
template <> const RuntimeParams<float>& runtime_prm<float>(const qs_env* e) { return e->prm_f; }
template <> const RuntimeParams<double>& runtime_prm<double>(const qs_env* e) { return e->prm_d; }

template <typename T, typename F> int dispatch_kind(const qs_env* e, F&& f) {
  if (e->custom) {
    switch (e->cfg.env_kind) {
      case QS_HOVER: return f(TypeTag<T>{}, IntTag<QS_HOVER>{}, runtime_prm<T>(e));
      case QS_DOCK_V0: return f(TypeTag<T>{}, IntTag<QS_DOCK_V0>{}, runtime_prm<T>(e));
      default: return f(TypeTag<T>{}, IntTag<QS_DOCK_V2>{}, runtime_prm<T>(e));
    }
  }
  switch (e->cfg.env_kind) {
    case QS_HOVER: return f(TypeTag<T>{}, IntTag<QS_HOVER>{}, DefaultParams<T, QS_HOVER>{});
    case QS_DOCK_V0: return f(TypeTag<T>{}, IntTag<QS_DOCK_V0>{}, DefaultParams<T, QS_DOCK_V0>{});
    default: return f(TypeTag<T>{}, IntTag<QS_DOCK_V2>{}, DefaultParams<T, QS_DOCK_V2>{});
  }
}
template <typename F> int dispatch_env(const qs_env* e, F&& f) {
  return e->cfg.dtype == QS_F32 ? dispatch_kind<float>(e, f) : dispatch_kind<double>(e, f);
}

bool use_pdl() {  // QUADSIM_B200_PDL=0 launches without programmatic dependent launch (A/B)
  static const bool pdl = [] { const char* v = getenv("QUADSIM_B200_PDL"); return !(v && v[0] == '0'); }();
  return pdl;
}

// slot >= 0: the scheduler-counter pair of the persistent pipelined kernel to use (launches that may run concurrently need
// different slots); slot < 0: always the direct kernel (qs_step_range: callers may step slices on any number of streams).
int launch_step(qs_env* e, const StepArgs& a, cudaStream_t st, int slot = 0) {
  if (a.n <= 0) return QS_OK;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kBlock); cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = use_pdl() ? 1 : 0;
  if (e->use_pipe && slot >= 0) {
    const int tiles = (int)((a.n + kBlock - 1) / kBlock);
    unsigned int* sched = e->sched + 2 * slot;
    cfg.gridDim = dim3((unsigned)(tiles < e->pipe_grid ? tiles : e->pipe_grid));
    return dispatch_env(e, [&](auto tt, auto kind, auto prm) -> int {
      using T = typename decltype(tt)::type;
      constexpr int KIND = decltype(kind)::value;
      if constexpr (KIND != QS_HOVER) {
        cfg.dynamicSmemBytes = pipe_smem_bytes<T>();
        QS_CUDA(cudaLaunchKernelEx(&cfg, dock_step_pipe_kernel<T, KIND, decltype(prm)>, a, prm, sched));
      }
      return QS_OK;
    });
  }
  cfg.gridDim = dim3((unsigned)((a.n + kBlock - 1) / kBlock));
  return dispatch_env(e, [&](auto tt, auto kind, auto prm) -> int {
    using T = typename decltype(tt)::type;
    using P = decltype(prm);
    constexpr int KIND = decltype(kind)::value;
    if constexpr (KIND == QS_HOVER) QS_CUDA(cudaLaunchKernelEx(&cfg, hover_step_kernel<T, P>, a, prm));
    else QS_CUDA(cudaLaunchKernelEx(&cfg, dock_step_kernel<T, KIND, P>, a, prm));
    return QS_OK;
  });
}

int launch_seat(qs_env* e, void* obs, int mode, cudaStream_t st) {
  StepArgs a = make_args(e, 0, e->cfg.n_envs);
  a.obs = obs;
  const unsigned grid = (unsigned)((a.n + kBlock - 1) / kBlock);
  return dispatch_env(e, [&](auto tt, auto kind, auto prm) -> int {
    using T = typename decltype(tt)::type;
    seat_kernel<T, decltype(kind)::value, decltype(prm)><<<grid, kBlock, 0, st>>>(a, prm, mode);
    QS_CUDA(cudaGetLastError());
    return QS_OK;
  });
}

template <typename T, bool IMPORT> int launch_state_io(qs_env* e, const qs_state_arrays_t* s, cudaStream_t st) {
  StateArrays sa = {s->chaser, s->target, s->u_chaser, s->u_target, s->des_q, s->last_shaping, s->ep_return, s->t, s->episode};
  const unsigned grid = (unsigned)((e->cfg.n_envs + kBlock - 1) / kBlock);
  const double* o = e->origin;
  if (e->cfg.env_kind == QS_HOVER) state_io_kernel<T, true, IMPORT><<<grid, kBlock, 0, st>>>(e->state, e->stride, e->cfg.n_envs, sa, o[0], o[1], o[2]);
  else state_io_kernel<T, false, IMPORT><<<grid, kBlock, 0, st>>>(e->state, e->stride, e->cfg.n_envs, sa, o[0], o[1], o[2]);
  QS_CUDA(cudaGetLastError());
  return QS_OK;
}

}  // namespace

// internal (not part of the C ABI; used by qs_rollout in qs_policy.cu): the step arguments of the whole handle for the fused
// policy-forward + env-step kernel -- float32 docking handles with the compiled-in constants only
int qs_env_fused_args_(qs_env_t* e, void* out, size_t out_size, int* kind) {
  if (!e || !out || out_size != sizeof(StepArgs) || !kind) return QS_ERR_INVALID;
  if (e->cfg.dtype != QS_F32 || e->cfg.env_kind == QS_HOVER || e->custom) return QS_ERR_INVALID;
  const StepArgs a = make_args(e, 0, e->cfg.n_envs);
  memcpy(out, &a, sizeof a);
  *kind = e->cfg.env_kind;
  return QS_OK;
}

extern "C" {

int qs_abi_version(void) { return QS_ABI_VERSION; }
const char* qs_last_error(void) { return g_err; }

int qs_default_params(int env_kind, qs_env_params_t* p) {
  if (!p) return fail(QS_ERR_INVALID, "qs_default_params: null argument");
  if (env_kind < QS_HOVER || env_kind > QS_DOCK_V2) return fail(QS_ERR_INVALID, "qs_default_params: unknown env_kind %d", env_kind);
  memset(p, 0, sizeof *p);
  p->dt = 0.02; p->gravity = 9.81; p->mass = 0.18; p->arm_length = 0.086;                          // quadrotor.py:12-21
  p->inertia[0] = 0.00025; p->inertia[1] = 0.000232; p->inertia[2] = 0.0003738;                    // :17-19
  p->motor_lambda = 1.5e-9 / 6.11e-8;                                                               // :43-45
  p->att_limit_deg[0] = 85; p->att_limit_deg[1] = 85; p->att_limit_deg[2] = 175;                   // :146-168
  p->kp_att[0] = -10; p->kp_att[1] = -10; p->kp_att[2] = -9.5;                                      // PIDController.py:10-12
  p->kd_att[0] = 5.1; p->kd_att[1] = 5.1; p->kd_att[2] = 4.0;                                       // :14-16
  p->kp_pos[0] = -1.0; p->kp_pos[1] = -1.0; p->kp_pos[2] = 50;                                      // :19-21
  p->kd_pos[0] = -1.65; p->kd_pos[1] = -1.65; p->kd_pos[2] = 8;                                     // :23-25
  p->kp_vel[0] = -0.7; p->kp_vel[1] = -0.7; p->kp_vel[2] = 1;                                       // :27-33
  const bool hover = env_kind == QS_HOVER;
  p->chaser_ini_pos[0] = hover ? 0 : 8; p->chaser_ini_pos[1] = hover ? 0 : -50; p->chaser_ini_pos[2] = 5;  // docking_env.py:34 / hovering_env.py:23
  p->target_ini_pos[0] = 10; p->target_ini_pos[1] = -50; p->target_ini_pos[2] = 5;                 // docking_env.py:47
  p->target_des_pos[0] = 10; p->target_des_pos[1] = -50; p->target_des_pos[2] = 5;                 // docking_env.py:59
  p->target_des_vel[0] = env_kind == QS_DOCK_V2 ? 0.2 : 0.0;                                        // moving_docking_env.py:62-66
  p->dock_port = 0.1;                                                                               // docking_env.py:38,51
  p->dock_pos_tol = 0.1; p->dock_vel_tol = 0.1; p->dock_angle_deg = 10;                             // docking_env.py:130-134
  p->range_limit = hover ? 100 : (env_kind == QS_DOCK_V0 ? 3 : 10);                                 // docking_env.py:141, moving_docking_env.py:148, hovering_env.py:68
  p->min_altitude = 0.1;                                                                            // docking_env.py:142
  p->max_steps = 600;                                                                               // docking_env.py:152
  return QS_OK;
}

int qs_create(qs_env_t** out, const qs_config_t* cfg) {
  if (!out || !cfg) return fail(QS_ERR_INVALID, "qs_create: null argument");
  *out = nullptr;
  if (cfg->env_kind < QS_HOVER || cfg->env_kind > QS_DOCK_V2) return fail(QS_ERR_INVALID, "qs_create: unknown env_kind %d", cfg->env_kind);
  if (cfg->dtype != QS_F32 && cfg->dtype != QS_F64) return fail(QS_ERR_INVALID, "qs_create: unknown dtype %d", cfg->dtype);
  if (cfg->n_envs <= 0) return fail(QS_ERR_INVALID, "qs_create: n_envs must be positive (got %lld)", (long long)cfg->n_envs);
  if (cfg->randomize < QS_RAND_NONE || cfg->randomize > QS_RAND_ONCE) return fail(QS_ERR_INVALID, "qs_create: unknown randomize mode %d", cfg->randomize);
  if (cfg->use_params) {
    const qs_env_params_t& q = cfg->params;
    const bool pos = q.dt > 0 && q.gravity > 0 && q.mass > 0 && q.arm_length > 0 && q.inertia[0] > 0 && q.inertia[1] > 0 && q.inertia[2] > 0 &&
                     q.motor_lambda > 0 && q.range_limit > 0 && q.dock_port >= 0 && q.dock_pos_tol >= 0 && q.dock_vel_tol >= 0 && q.dock_angle_deg >= 0;
    if (!pos) return fail(QS_ERR_INVALID, "qs_create: params: dt, gravity, mass, arm_length, inertia, motor_lambda and range_limit must be positive, tolerances non-negative");
    // the clamp pre-test is written for the reference's geometry: roll / pitch limits below 90 degrees, yaw limit above
    if (!(q.att_limit_deg[0] > 0 && q.att_limit_deg[0] < 90 && q.att_limit_deg[1] > 0 && q.att_limit_deg[1] < 90 && q.att_limit_deg[2] > 90 && q.att_limit_deg[2] <= 180))
      return fail(QS_ERR_INVALID, "qs_create: params.att_limit_deg must satisfy 0 < roll, pitch < 90 and 90 < yaw <= 180 (reference: 85, 85, 175)");
    if (q.max_steps < 1 || (cfg->dtype == QS_F32 && cfg->env_kind != QS_HOVER && q.max_steps > 1023))
      return fail(QS_ERR_INVALID, "qs_create: params.max_steps must be >= 1 (and <= 1023 for float32 docking handles: t is a 10-bit field)");
  }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(QS_ERR_NO_DEVICE, "qs_create: no CUDA device (%s); quadsim_b200 has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(QS_ERR_INVALID, "qs_create: device %d out of range [0,%d)", cfg->device, ndev);
  DeviceGuard g(cfg->device);
  if (!g.ok) return fail(QS_ERR_CUDA, "qs_create: cannot select device %d", cfg->device);
  qs_env* e = new (std::nothrow) qs_env();
  if (!e) return fail(QS_ERR_ALLOC, "qs_create: out of host memory");
  e->cfg = *cfg;
  e->custom = cfg->use_params != 0;
  if (e->custom) e->params = cfg->params; else qs_default_params(cfg->env_kind, &e->params);
  make_runtime_params<float>(e->params, cfg->env_kind, e->prm_f);
  make_runtime_params<double>(e->params, cfg->env_kind, e->prm_d);
  if (cfg->env_kind != QS_HOVER) {
    if (cfg->dtype == QS_F32) { e->origin[0] = (double)e->prm_f.ox; e->origin[1] = (double)e->prm_f.oy; e->origin[2] = (double)e->prm_f.oz; }
#ifdef QS_LITERAL32
    e->origin[0] = e->origin[1] = e->origin[2] = 0.0;  // the literal fp32 build keeps world coordinates
#endif
  }
  if (e->cfg.half_width <= 0) e->cfg.half_width = cfg->env_kind == QS_HOVER ? 1.0 : 0.5;
  e->jitter_scale = (float)(2.0 * e->cfg.half_width);
  e->obs_dim = cfg->env_kind == QS_HOVER ? 13 : 12;
  e->esize = cfg->dtype == QS_F32 ? 4 : 8;
  const int words = cfg->env_kind == QS_HOVER ? kHoverWords : 40;
  e->n_rows = words / (cfg->dtype == QS_F32 ? 4 : 2);
  e->stride = (cfg->n_envs + kBlock - 1) / kBlock * kBlock;
  e->n_slots = e->stride / 32;
  cudaError_t err;
  if ((err = cudaMalloc(&e->state, (size_t)e->n_rows * e->stride * 16)) != cudaSuccess ||
      (err = cudaMalloc(&e->warp_stats, (size_t)e->n_slots * kStatWords * sizeof(double))) != cudaSuccess ||
      (err = cudaMalloc(&e->stats_dev, (8 + 8 * kStatBlocks) * sizeof(double))) != cudaSuccess ||
      (err = cudaMallocHost(&e->stats_host, 8 * sizeof(double))) != cudaSuccess ||
      (err = cudaMemset(e->state, 0, (size_t)e->n_rows * e->stride * 16)) != cudaSuccess ||
      (err = cudaMemset(e->warp_stats, 0, (size_t)e->n_slots * kStatWords * sizeof(double))) != cudaSuccess) {
    int rc = fail(QS_ERR_ALLOC, "qs_create: device allocation failed: %s", cudaGetErrorString(err));
    qs_destroy(e);
    return rc;
  }
  // fp64 docking runs 12 warps per SM (168 registers): too few to hide HBM latency with plain loads, so the persistent
  // bulk-async kernel is its default (138.8 vs 152.1 us per 2^20-env step); fp32 (24 warps per SM) is faster direct (63 vs 73).
  // QUADSIM_B200_KERNEL=pipe|direct overrides.
  e->use_pipe = e->cfg.dtype == QS_F64;
  if (const char* k = getenv("QUADSIM_B200_KERNEL")) e->use_pipe = strcmp(k, "pipe") == 0 ? 1 : (strcmp(k, "direct") == 0 ? 0 : e->use_pipe);
  if (e->cfg.env_kind == QS_HOVER) e->use_pipe = 0;
  // QUADSIM_B200_CONSTS=bank: run even default handles on the RuntimeParams kernels (constants as constant-bank operands
  // instead of immediates -- an A/B switch: 64-bit immediates cost the fp64 kernel an extra instruction each)
  if (const char* c = getenv("QUADSIM_B200_CONSTS")) { if (strcmp(c, "bank") == 0) e->custom = true; }
  if (e->use_pipe) {
    cudaDeviceProp prop;
    int per_sm = 0;
    cudaError_t perr = cudaGetDeviceProperties(&prop, cfg->device);
    if (perr == cudaSuccess) {
      dispatch_env(e, [&](auto tt, auto kind, auto prm) -> int {
        using T = typename decltype(tt)::type;
        constexpr int KIND = decltype(kind)::value;
        if constexpr (KIND != QS_HOVER) {
          auto kfn = dock_step_pipe_kernel<T, KIND, decltype(prm)>;
          perr = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, pipe_smem_bytes<T>());
          if (perr == cudaSuccess) perr = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kBlock, pipe_smem_bytes<T>());
        }
        return QS_OK;
      });
    }
    if (perr != cudaSuccess || per_sm < 1 || cudaMalloc(&e->sched, (kHostStreams + 1) * 2 * sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(e->sched, 0, (kHostStreams + 1) * 2 * sizeof(unsigned int)) != cudaSuccess) {
      int rc = fail(QS_ERR_CUDA, "qs_create: pipelined kernel setup failed: %s", cudaGetErrorString(cudaGetLastError()));
      qs_destroy(e);
      return rc;
    }
    if (const char* g = getenv("QUADSIM_B200_PIPE_CTAS")) per_sm = atoi(g) < per_sm && atoi(g) > 0 ? atoi(g) : per_sm;
    e->pipe_grid = per_sm * prop.multiProcessorCount;
  }
  int rc = launch_seat(e, nullptr, 0, 0);
  if (rc != QS_OK) { qs_destroy(e); return rc; }
  if ((err = cudaStreamSynchronize(0)) != cudaSuccess) {
    rc = fail(QS_ERR_CUDA, "qs_create: init kernel failed: %s", cudaGetErrorString(err));
    qs_destroy(e);
    return rc;
  }
  *out = e;
  return QS_OK;
}

int qs_destroy(qs_env_t* e) {
  if (!e) return QS_OK;
  DeviceGuard g(e->cfg.device);
  for (auto& s : e->hstreams) if (s) cudaStreamDestroy(s);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  cudaFree(e->import_flag);
  cudaFree(e->h_act); cudaFree(e->h_obs); cudaFree(e->h_rew); cudaFree(e->h_done);
  cudaFree(e->state); cudaFree(e->warp_stats); cudaFree(e->stats_dev); cudaFree(e->sched);
  if (e->stats_host) cudaFreeHost(e->stats_host);
  delete e;
  return QS_OK;
}

int qs_obs_dim(const qs_env_t* e) { return e ? e->obs_dim : fail(QS_ERR_INVALID, "null handle"); }
int64_t qs_num_envs(const qs_env_t* e) { return e ? e->cfg.n_envs : (int64_t)fail(QS_ERR_INVALID, "null handle"); }

int qs_env_dtype(const qs_env_t* e) { return e ? e->cfg.dtype : fail(QS_ERR_INVALID, "null handle"); }
int qs_env_device(const qs_env_t* e) { return e ? e->cfg.device : fail(QS_ERR_INVALID, "null handle"); }
uint64_t qs_env_id0(const qs_env_t* e) { return e ? e->cfg.env_id0 : 0; }

int qs_seed(qs_env_t* e, uint64_t seed) {
  if (!e) return fail(QS_ERR_INVALID, "null handle");
  e->cfg.seed = seed;
  return QS_OK;
}

int qs_reset(qs_env_t* e, void* obs_out, void* stream) {
  if (!e || !obs_out) return fail(QS_ERR_INVALID, "qs_reset: null argument");
  NvtxRange nv("qs_reset");
  DeviceGuard g(e->cfg.device);
  return launch_seat(e, obs_out, 1, (cudaStream_t)stream);
}

int qs_step_ex(qs_env_t* e, const void* actions, void* obs_out, void* reward_out, uint8_t* done_out, const qs_step_extras_t* x,
               void* stream) {
  if (!e || !actions || !obs_out || !reward_out || !done_out) return fail(QS_ERR_INVALID, "qs_step: null argument");
  NvtxRange nv("qs_step");
  DeviceGuard g(e->cfg.device);
  StepArgs a = make_args(e, 0, e->cfg.n_envs);
  a.actions = actions; a.obs = obs_out; a.reward = reward_out; a.done = done_out;
  if (x) { a.info = x->info; a.term_obs = x->terminal_obs; a.term_chaser = x->terminal_chaser; a.term_target = x->terminal_target; }
  return launch_step(e, a, (cudaStream_t)stream);
}

int qs_step_range(qs_env_t* e, int64_t first, int64_t n, const void* actions, void* obs_out, void* reward_out, uint8_t* done_out,
                  void* stream) {
  if (!e || !actions || !obs_out || !reward_out || !done_out) return fail(QS_ERR_INVALID, "qs_step_range: null argument");
  if (first < 0 || n < 0 || first + n > e->cfg.n_envs || first % kBlock != 0)
    return fail(QS_ERR_INVALID, "qs_step_range: [first, first+n) must lie inside the handle's envs and start on a multiple of %d", kBlock);
  NvtxRange nv("qs_step_range");
  DeviceGuard g(e->cfg.device);
  StepArgs a = make_args(e, first, n);
  a.actions = actions; a.obs = obs_out; a.reward = reward_out; a.done = done_out;
  return launch_step(e, a, (cudaStream_t)stream, -1);
}

int qs_step_n(qs_env_t* e, int64_t n_steps, const void* actions, void* obs_out, void* reward_out, uint8_t* done_out, void* stream) {
  if (!e || !actions || !obs_out || !reward_out || !done_out) return fail(QS_ERR_INVALID, "qs_step_n: null argument");
  if (n_steps < 0) return fail(QS_ERR_INVALID, "qs_step_n: n_steps < 0");
  if (n_steps == 0) return QS_OK;
  NvtxRange nv("qs_step_n");
  DeviceGuard g(e->cfg.device);
  StepArgs a = make_args(e, 0, e->cfg.n_envs);
  a.actions = actions; a.obs = obs_out; a.reward = reward_out; a.done = done_out;
  const unsigned grid = (unsigned)((a.n + kBlock - 1) / kBlock);
  cudaStream_t st = (cudaStream_t)stream;
  return dispatch_env(e, [&](auto tt, auto kind, auto prm) -> int {
    using T = typename decltype(tt)::type;
    multi_step_kernel<T, decltype(kind)::value, decltype(prm)><<<grid, kBlock, 0, st>>>(a, prm, n_steps);
    QS_CUDA(cudaGetLastError());
    return QS_OK;
  });
}

int qs_step(qs_env_t* e, const void* actions, void* obs_out, void* reward_out, uint8_t* done_out, void* stream) {
  return qs_step_ex(e, actions, obs_out, reward_out, done_out, nullptr, stream);
}

int qs_host_alloc(void** out, uint64_t bytes) {
  if (!out) return fail(QS_ERR_INVALID, "qs_host_alloc: null argument");
  QS_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
  return QS_OK;
}
int qs_host_free(void* p) {
  if (p) QS_CUDA(cudaFreeHost(p));
  return QS_OK;
}

int qs_step_host(qs_env_t* e, const void* actions_host, void* obs_host, void* reward_host, uint8_t* done_host, void* stream) {
  if (!e || !actions_host || !obs_host || !reward_host || !done_host) return fail(QS_ERR_INVALID, "qs_step_host: null argument");
  NvtxRange nv("qs_step_host");
  DeviceGuard g(e->cfg.device);
  const int64_t N = e->cfg.n_envs;
  const size_t es = e->esize;
  if (!e->h_act) {
    for (auto& s : e->hstreams) QS_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    QS_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    QS_CUDA(cudaMalloc(&e->h_act, (size_t)N * 4 * es));
    QS_CUDA(cudaMalloc(&e->h_obs, (size_t)N * e->obs_dim * es));
    QS_CUDA(cudaMalloc(&e->h_rew, (size_t)N * es));
    QS_CUDA(cudaMalloc((void**)&e->h_done, (size_t)N));
  }
  // fork: whatever the caller enqueued on ITS stream before this call (reset, import, a device-side step of the same
  // handle) is ordered before the chunks on the internal non-blocking streams -- no host-side wait, no assumption that the
  // caller works on the legacy default stream
  QS_CUDA(cudaEventRecord(e->ev_fork, (cudaStream_t)stream));
  for (auto& s : e->hstreams) QS_CUDA(cudaStreamWaitEvent(s, e->ev_fork, 0));
  int k = 0;
  // reward (4-8 B/env) and done (1 B/env) are too small to be worth a DMA descriptor per chunk: when the caller's buffers are
  // pinned (mapped into the device's address space under UVA) the kernel writes them straight across PCIe -- posted,
  // coalesced 128-byte / 32-byte stores -- and only obs (48 B/env) goes through the copy engine.  QUADSIM_B200_HOST_DIRECT=0
  // forces the copies (A/B); pageable buffers always take them.  Measured: 8.58e8 -> 9.03e8 env-steps/s (1 GPU, 2^20 envs).
  // The same trick for obs LOSES 3x (3.0e8): its 16-byte stores at a 48-byte stride become tiny PCIe writes.
  void *rew_dev = e->h_rew, *done_dev = e->h_done;
  bool direct = false;
  {
    static const bool allow = [] { const char* v = getenv("QUADSIM_B200_HOST_DIRECT"); return !(v && v[0] == '0'); }();
    cudaPointerAttributes ar, ad;
    if (allow && cudaPointerGetAttributes(&ar, reward_host) == cudaSuccess && cudaPointerGetAttributes(&ad, done_host) == cudaSuccess &&
        ar.type == cudaMemoryTypeHost && ad.type == cudaMemoryTypeHost && ar.devicePointer && ad.devicePointer) {
      rew_dev = ar.devicePointer; done_dev = ad.devicePointer; direct = true;
    }
    cudaGetLastError();  // a failed query of a pageable pointer is not an error of this call
  }
  // chunk size: QUADSIM_B200_HOST_CHUNK (envs, a multiple of 128) overrides the default for experiments
  static const int64_t chunk = [] {
    const char* v = getenv("QUADSIM_B200_HOST_CHUNK");
    const int64_t c = v ? atoll(v) : 0;
    return c >= kBlock && c % kBlock == 0 ? c : kHostChunk;
  }();
  // Default schedule: graduated chunks (32 Ki, 64 Ki, 128 Ki, then 256 Ki envs each) -- the first result copy can start after
  // a 32 Ki-env upload + kernel instead of a 128 Ki one, and the long tail runs on few large copies (fewer gaps between DMA
  // descriptors).  A fixed size (QUADSIM_B200_HOST_CHUNK) keeps the old uniform schedule for A/B.
  static const bool fixed = getenv("QUADSIM_B200_HOST_CHUNK") != nullptr;
  static const int64_t c0 = [] { const char* v = getenv("QUADSIM_B200_HOST_CHUNK0"); const int64_t c = v ? atoll(v) : 0; return c >= kBlock && c % kBlock == 0 ? c : (int64_t)32768; }();
  static const int64_t cap = [] { const char* v = getenv("QUADSIM_B200_HOST_CHUNK_CAP"); const int64_t c = v ? atoll(v) : 0; return c >= kBlock && c % kBlock == 0 ? c : (int64_t)262144; }();
  int64_t cur = fixed ? chunk : c0;
  for (int64_t first = 0; first < N; k++) {
    const int64_t n = (N - first) < cur ? (N - first) : cur;
    cudaStream_t st = e->hstreams[k % kHostStreams];
    QS_CUDA(cudaMemcpyAsync((char*)e->h_act + first * 4 * es, (const char*)actions_host + first * 4 * es, (size_t)n * 4 * es, cudaMemcpyHostToDevice, st));
    StepArgs a = make_args(e, first, n);
    a.actions = e->h_act; a.obs = e->h_obs; a.reward = rew_dev; a.done = (uint8_t*)done_dev;
    int rc = launch_step(e, a, st, 1 + k % kHostStreams);
    if (rc != QS_OK) return rc;
    QS_CUDA(cudaMemcpyAsync((char*)obs_host + first * e->obs_dim * es, (char*)e->h_obs + first * e->obs_dim * es, (size_t)n * e->obs_dim * es, cudaMemcpyDeviceToHost, st));
    if (!direct) {
      QS_CUDA(cudaMemcpyAsync((char*)reward_host + first * es, (char*)e->h_rew + first * es, (size_t)n * es, cudaMemcpyDeviceToHost, st));
      QS_CUDA(cudaMemcpyAsync(done_host + first, e->h_done + first, (size_t)n, cudaMemcpyDeviceToHost, st));
    }
    first += n;
    if (!fixed && cur < cap) cur *= 2;
  }
  // join: the call is blocking by contract (results are in host memory on return), which also orders everything the
  // caller enqueues afterwards, on any stream, behind the chunks
  for (auto& s : e->hstreams) QS_CUDA(cudaStreamSynchronize(s));
  return QS_OK;
}

int qs_export_state(qs_env_t* e, const qs_state_arrays_t* dst, void* stream) {
  if (!e || !dst) return fail(QS_ERR_INVALID, "qs_export_state: null argument");
  DeviceGuard g(e->cfg.device);
  return e->cfg.dtype == QS_F32 ? launch_state_io<float, false>(e, dst, (cudaStream_t)stream) : launch_state_io<double, false>(e, dst, (cudaStream_t)stream);
}
int qs_import_state(qs_env_t* e, const qs_state_arrays_t* src, void* stream) {
  if (!e || !src) return fail(QS_ERR_INVALID, "qs_import_state: null argument");
  DeviceGuard g(e->cfg.device);
  if (e->cfg.dtype == QS_F32 && e->cfg.env_kind != QS_HOVER && (src->t || src->episode)) {
    // the packed counters word (see the header): validate first, so a rejected import leaves the state untouched
    cudaStream_t st = (cudaStream_t)stream;
    int bad = 0;
    if (!e->import_flag) QS_CUDA(cudaMalloc((void**)&e->import_flag, sizeof(int)));
    QS_CUDA(cudaMemsetAsync(e->import_flag, 0, sizeof(int), st));
    counter_range_kernel<<<(unsigned)((e->cfg.n_envs + 255) / 256), 256, 0, st>>>(e->cfg.n_envs, src->t, src->episode, e->import_flag);
    QS_CUDA(cudaGetLastError());
    QS_CUDA(cudaMemcpyAsync(&bad, e->import_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    QS_CUDA(cudaStreamSynchronize(st));
    if (bad) return fail(QS_ERR_INVALID, "qs_import_state: a float32 docking handle keeps t in 10 bits and episode in 22 bits; got t > 1023 or episode outside [0, 2^22)");
  }
  return e->cfg.dtype == QS_F32 ? launch_state_io<float, true>(e, src, (cudaStream_t)stream) : launch_state_io<double, true>(e, src, (cudaStream_t)stream);
}

int qs_get_stats(qs_env_t* e, double stats[5], void* stream) {
  if (!e || !stats) return fail(QS_ERR_INVALID, "qs_get_stats: null argument");
  DeviceGuard g(e->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  stats_partial_kernel<<<kStatBlocks, 256, 0, st>>>(e->warp_stats, e->n_slots, e->stats_dev + 8);
  stats_final_kernel<<<1, 32, 0, st>>>(e->stats_dev + 8, e->stats_dev);
  QS_CUDA(cudaGetLastError());
  QS_CUDA(cudaMemcpyAsync(e->stats_host, e->stats_dev, 5 * sizeof(double), cudaMemcpyDeviceToHost, st));
  QS_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 5; k++) stats[k] = e->stats_host[k];
  return QS_OK;
}

int qs_fn_apply(int fn, int dtype, int64_t n, const void* in0, const void* in1, const void* in2, void* out0, void* out1, void* stream) {
  if (n <= 0) return QS_OK;
  if (fn < QS_FN_QUAT2EULER || fn > QS_FN_MATH) return fail(QS_ERR_INVALID, "qs_fn_apply: unknown fn %d", fn);
  if (!in0 || !out0) return fail(QS_ERR_INVALID, "qs_fn_apply: null in0/out0");
  const bool two_in = fn == QS_FN_DRONE_STEP || fn == QS_FN_PID || fn == QS_FN_VEL_CTRL || fn == QS_FN_STATE2REL || fn == QS_FN_PHILOX;
  const bool two_out = fn == QS_FN_DRONE_STEP || fn == QS_FN_PID || fn == QS_FN_VEL_CTRL;
  if ((two_in && !in1) || (two_out && !out1) || (fn == QS_FN_DRONE_STEP && !in2)) return fail(QS_ERR_INVALID, "qs_fn_apply: missing operand for fn %d", fn);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (fn == QS_FN_PHILOX) philox_kernel<<<grid, 128, 0, st>>>(n, (const uint32_t*)in0, (const uint32_t*)in1, (uint32_t*)out0);
  else if (dtype == QS_F32) fn_kernel<float><<<grid, 128, 0, st>>>(fn, n, (const float*)in0, (const float*)in1, (const float*)in2, (float*)out0, (float*)out1);
  else if (dtype == QS_F64) fn_kernel<double><<<grid, 128, 0, st>>>(fn, n, (const double*)in0, (const double*)in1, (const double*)in2, (double*)out0, (double*)out1);
  else return fail(QS_ERR_INVALID, "qs_fn_apply: unknown dtype %d", dtype);
  QS_CUDA(cudaGetLastError());
  return QS_OK;
}

int qs_gae(int64_t T, int64_t n, const float* rewards, const float* values, const uint8_t* dones, const float* last_values,
           const uint8_t* last_dones, double gamma, double lam, float* advs_out, float* returns_out, void* stream) {
  if (T <= 0 || n <= 0) return fail(QS_ERR_INVALID, "qs_gae: n_steps and n_envs must be positive");
  if (!rewards || !values || !dones || !last_values || !last_dones || !advs_out || !returns_out) return fail(QS_ERR_INVALID, "qs_gae: null argument");
  gae_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, n, rewards, values, dones, last_values, last_dones, gamma, lam, advs_out, returns_out);
  QS_CUDA(cudaGetLastError());
  return QS_OK;
}

int qs_expert_action(qs_env_t* e, void* expert_mem, int alias_ini_state, void* actions_out, void* obs_inout, void* stream) {
  if (!e || !expert_mem || !actions_out) return fail(QS_ERR_INVALID, "qs_expert_action: null argument");
  if (e->cfg.env_kind == QS_HOVER) return fail(QS_ERR_INVALID, "qs_expert_action: the PID expert flies the chaser of a docking environment");
  DeviceGuard g(e->cfg.device);
  const int64_t n = e->cfg.n_envs;
  const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
  return dispatch_env(e, [&](auto tt, auto kind, auto prm) -> int {
    using T = typename decltype(tt)::type;
    if constexpr (decltype(kind)::value != QS_HOVER) {
      expert_kernel<T, decltype(prm)><<<grid, kBlock, 0, (cudaStream_t)stream>>>(e->state, n, (T*)expert_mem, alias_ini_state, (T*)actions_out, obs_inout, prm);
      QS_CUDA(cudaGetLastError());
    }
    return QS_OK;
  });
}

int qs_peek_state(qs_env_t* e, const int64_t* env_ids, int64_t k, double* chaser_out, double* target_out, int32_t* t_out, void* stream) {
  if (!e || !env_ids || !chaser_out) return fail(QS_ERR_INVALID, "qs_peek_state: null argument");
  if (k <= 0) return QS_OK;
  DeviceGuard g(e->cfg.device);
  const unsigned grid = (unsigned)((k + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  const bool hover = e->cfg.env_kind == QS_HOVER;
  const double* o = e->origin;
  if (e->cfg.dtype == QS_F32) {
    if (hover) peek_state_kernel<float, true><<<grid, 128, 0, st>>>(e->state, e->cfg.n_envs, env_ids, k, chaser_out, target_out, t_out, o[0], o[1], o[2]);
    else peek_state_kernel<float, false><<<grid, 128, 0, st>>>(e->state, e->cfg.n_envs, env_ids, k, chaser_out, target_out, t_out, o[0], o[1], o[2]);
  } else {
    if (hover) peek_state_kernel<double, true><<<grid, 128, 0, st>>>(e->state, e->cfg.n_envs, env_ids, k, chaser_out, target_out, t_out, o[0], o[1], o[2]);
    else peek_state_kernel<double, false><<<grid, 128, 0, st>>>(e->state, e->cfg.n_envs, env_ids, k, chaser_out, target_out, t_out, o[0], o[1], o[2]);
  }
  QS_CUDA(cudaGetLastError());
  return QS_OK;
}

}  // extern "C"
