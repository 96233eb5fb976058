// qs_env_kernels.cuh -- device code shared by every kernel that steps an environment: the HBM row layout, the word <->
// struct mapping, the per-warp statistics and the "one env whose state words are in registers" step (dock_step_loaded).
// Included by qs_lib.cu (the env step kernels) and by qs_policy.cu (the fused policy-forward + env-step kernel of the
// rollout); everything sits in an anonymous namespace, i.e. each translation unit compiles its own copy.
#pragma once
#include <cuda_runtime.h>

#include "qs_core.cuh"

namespace {

using namespace qs;

constexpr int kBlock = 128;        // threads per CTA (4 warps); grid = ceil(n / 128)
#ifndef QS_MINB32
#define QS_MINB32 1
#endif
// resident CTAs per SM requested from ptxas for the fp32 step kernels (register cap = 65536 / (128 * minb))
#ifndef QS_MINB64
#define QS_MINB64 1
#endif
template <typename T> constexpr int min_blocks() { return is_f32<T>::value ? QS_MINB32 : QS_MINB64; }
constexpr int kStatWords = 8;      // doubles per warp statistics slot (5 used), 64-byte stride

// ---------------------------------------------------------------------------------------------------------
// HBM layout.  A "row" is 16 bytes per env: Row<float>::V = float4 (4 words), Row<double>::V = double2 (2 words).
// Rows are stored TILE-MAJOR: the NR rows of one 128-env tile are adjacent, row r of env i sits at
// rows[((i / 128) * NR + r) * 128 + i % 128].  A warp-level access is one contiguous 512 B span (every sector fully
// used) and a CTA's whole tile state is ONE contiguous NR x 2 KB block, so a step streams a single read stream and a
// single write stream through the state instead of NR interleaved ones (DRAM page locality).  Word order (docking, exactly 40 words = 10 fp32 rows / 20 fp64 rows):
//    0- 3  chaser p.xyz, w.x      4- 7  chaser v.xyz, w.y     8-11  chaser q      12-15  chaser u_prev
//   16-19  target p.xyz, w.x     20-23  target v.xyz, w.y    24-27  target q      28-31  target u_prev
//   32-35  des_q                 36-39  chaser w.z, target w.z, last_shaping, counters
// counters = t | episode << 10 in a 32-bit word (fp32: t < 1024 -- episodes end at t = 600 --, episode mod 2^22) or
// t | episode << 32 in a 64-bit word (fp64).  The episode return is not stored: it equals last_shaping (telescoping).
// Hover (20 words): 0-15 as the chaser above; 16-19 = w.z, ep_return, t (int), episode (uint).
// ---------------------------------------------------------------------------------------------------------
template <typename T> struct Row;
template <> struct Row<float> { using V = float4; static constexpr int W = 4; };
template <> struct Row<double> { using V = double2; static constexpr int W = 2; };

template <typename T> constexpr int dock_words() { return 40; }
constexpr int kHoverWords = 20;

__device__ __forceinline__ void unpack(const float4& v, float* w) { w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
__device__ __forceinline__ void unpack(const double2& v, double* w) { w[0] = v.x; w[1] = v.y; }
__device__ __forceinline__ float4 pack(const float* w) { return make_float4(w[0], w[1], w[2], w[3]); }
__device__ __forceinline__ double2 pack(const double* w) { return make_double2(w[0], w[1]); }

template <typename T, int NW> __device__ __forceinline__ void load_words(const void* base, int64_t i, T (&w)[NW]) {
  using V = typename Row<T>::V;
  constexpr int W = Row<T>::W, NR = NW / W;
  const V* rows = reinterpret_cast<const V*>(base) + (i / kBlock) * (NR * kBlock) + (i % kBlock);
#pragma unroll
  for (int r = 0; r < NR; r++) unpack(rows[r * kBlock], &w[r * W]);
}
template <typename T, int NW> __device__ __forceinline__ void store_words(void* base, int64_t i, const T (&w)[NW]) {
  using V = typename Row<T>::V;
  constexpr int W = Row<T>::W, NR = NW / W;
  V* rows = reinterpret_cast<V*>(base) + (i / kBlock) * (NR * kBlock) + (i % kBlock);
#pragma unroll
  for (int r = 0; r < NR; r++) rows[r * kBlock] = pack(&w[r * W]);
}

__device__ __forceinline__ int32_t word_to_int(float x) { return __float_as_int(x); }
__device__ __forceinline__ int32_t word_to_int(double x) { return (int32_t)__double_as_longlong(x); }
template <typename T> __device__ __forceinline__ T int_to_word(int32_t v);
template <> __device__ __forceinline__ float int_to_word<float>(int32_t v) { return __int_as_float(v); }
template <> __device__ __forceinline__ double int_to_word<double>(int32_t v) { return __longlong_as_double((long long)v); }
// docking counters word
__device__ __forceinline__ void unpack_counters(float x, int32_t& t, int32_t& ep) {
  const uint32_t u = __float_as_uint(x);
  t = (int32_t)(u & 1023u); ep = (int32_t)(u >> 10);
}
__device__ __forceinline__ void unpack_counters(double x, int32_t& t, int32_t& ep) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  t = (int32_t)(u & 0xffffffffull); ep = (int32_t)(u >> 32);
}
__device__ __forceinline__ float pack_counters_f(int32_t t, int32_t ep) {
  return __uint_as_float((uint32_t)(t > 1023 ? 1023 : t) | ((uint32_t)ep << 10));
}
__device__ __forceinline__ double pack_counters_d(int32_t t, int32_t ep) {
  return __longlong_as_double((long long)((unsigned long long)(uint32_t)t | ((unsigned long long)(uint32_t)ep << 32)));
}
template <typename T> __device__ __forceinline__ T pack_counters(int32_t t, int32_t ep);
template <> __device__ __forceinline__ float pack_counters<float>(int32_t t, int32_t ep) { return pack_counters_f(t, ep); }
template <> __device__ __forceinline__ double pack_counters<double>(int32_t t, int32_t ep) { return pack_counters_d(t, ep); }

// row-major [N, C] outputs written as 16-byte vectors (C*sizeof(T) is a multiple of 16 for C in {4, 12})
template <typename T, int C> __device__ __forceinline__ void store_row(void* base, int64_t i, const T* v) {
  using V = typename Row<T>::V;
  constexpr int W = Row<T>::W;
  static_assert(C % W == 0, "row width must be a whole number of 16-byte vectors");
  V* p = reinterpret_cast<V*>(reinterpret_cast<T*>(base) + i * C);
#pragma unroll
  for (int k = 0; k < C / W; k++) p[k] = pack(&v[k * W]);
}
template <typename T> __device__ __forceinline__ void load_action(const void* base, int64_t i, T a[4]) {
  using V = typename Row<T>::V;
  constexpr int W = Row<T>::W;
  const V* p = reinterpret_cast<const V*>(reinterpret_cast<const T*>(base) + i * 4);
#pragma unroll
  for (int k = 0; k < 4 / W; k++) unpack(__ldg(p + k), &a[k * W]);
}

struct EnvStepArgs {
  void* state;
  int64_t stride;   // envs per row (n_envs rounded up)
  int64_t first, n; // this launch covers envs [first, first + n)
  const void* actions;
  void* obs;
  void* reward;
  uint8_t* done;
  uint8_t* info;      // may be null
  void* term_obs;     // may be null
  double *term_chaser, *term_target;  // may be null: [N,13] float64 world coordinates, rows of envs that finish in this step
  double* warp_stats; // [ceil(N/32)][kStatWords]
  RngCfg rng;
  int auto_reset;
};

// Per-warp statistics: deterministic (fixed lane order, one writer per slot and launch).  Slot = global warp index of the env range, so
// concurrent chunk launches (qs_step_host) never share a slot.
__device__ __forceinline__ void warp_stats_update(double* warp_stats, int64_t i, unsigned mask, bool done, bool docking,
                                                  bool nonfinite, double ep_ret, int ep_len) {
  unsigned bd = __ballot_sync(mask, done), bk = __ballot_sync(mask, docking), bn = __ballot_sync(mask, nonfinite);
  if ((bd | bk | bn) == 0) return;
  double sret = 0.0, slen = 0.0;
  for (unsigned m = bd; m; m &= m - 1) {
    int l = __ffs(m) - 1;
    sret += __shfl_sync(mask, ep_ret, l);
    slen += (double)__shfl_sync(mask, ep_len, l);
  }
  if ((threadIdx.x & 31) == (__ffs(mask) - 1)) {
    double* s = warp_stats + (i >> 5) * kStatWords;
    // RED (result unused): a slot has one writer per launch, so the sums are still order-deterministic, and the warp does
    // not wait a DRAM round trip for the old value at its very end (a read-modify-write here was ~10 % of the fp64 step's
    // stall samples)
    if (bd) { atomicAdd(s + 0, sret); atomicAdd(s + 1, slen); atomicAdd(s + 2, (double)__popc(bd)); }
    if (bk) atomicAdd(s + 3, (double)__popc(bk));
    if (bn) atomicAdd(s + 4, (double)__popc(bn));
  }
}

// Programmatic dependent launch (sm_90+): step k+1 is the same grid over the same state, so consecutive launches are fully
// dependent -- but the ~3 us between the last CTA of one step retiring and the first CTA of the next doing work (launch
// latency, CTA ramp) need not be serial.  Every step kernel tells the scheduler at its first instruction that dependents
// may be staged (they are only launched once ALL of this grid's CTAs have started, so the tail of a step shares the SMs with
// the head of the next), and waits for the previous grid's completion and memory flush before touching global memory.
// Both are no-ops for launches without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// word <-> struct mapping of the row layout above
template <typename T, int NW> __device__ __forceinline__ void words_to_dock(const T (&w)[NW], DockEnv<T>& e) {
  e.c = Drone<T>{w[0], w[1], w[2], w[4], w[5], w[6], w[8], w[9], w[10], w[11], w[3], w[7], w[36], w[12], w[13], w[14], w[15]};
  e.t = Drone<T>{w[16], w[17], w[18], w[20], w[21], w[22], w[24], w[25], w[26], w[27], w[19], w[23], w[37], w[28], w[29], w[30], w[31]};
  e.dq = Quat<T>{w[32], w[33], w[34], w[35]};
  e.last_shaping = w[38];
  unpack_counters(w[39], e.tt, e.episode);
}
template <typename T, int NW> __device__ __forceinline__ void dock_to_words(const DockEnv<T>& e, T (&w)[NW]) {
  const Drone<T>&c = e.c, &t = e.t;
  w[0] = c.px; w[1] = c.py; w[2] = c.pz; w[3] = c.wx; w[4] = c.vx; w[5] = c.vy; w[6] = c.vz; w[7] = c.wy;
  w[8] = c.q0; w[9] = c.q1; w[10] = c.q2; w[11] = c.q3; w[12] = c.u0; w[13] = c.u1; w[14] = c.u2; w[15] = c.u3;
  w[16] = t.px; w[17] = t.py; w[18] = t.pz; w[19] = t.wx; w[20] = t.vx; w[21] = t.vy; w[22] = t.vz; w[23] = t.wy;
  w[24] = t.q0; w[25] = t.q1; w[26] = t.q2; w[27] = t.q3; w[28] = t.u0; w[29] = t.u1; w[30] = t.u2; w[31] = t.u3;
  w[32] = e.dq.w; w[33] = e.dq.x; w[34] = e.dq.y; w[35] = e.dq.z;
  w[36] = c.wz; w[37] = t.wz; w[38] = e.last_shaping; w[39] = pack_counters<T>(e.tt, e.episode);
}
template <typename T> __device__ __forceinline__ void words_to_hover(const T (&w)[kHoverWords], HoverEnv<T>& e) {
  e.d = Drone<T>{w[0], w[1], w[2], w[4], w[5], w[6], w[8], w[9], w[10], w[11], w[3], w[7], w[16], w[12], w[13], w[14], w[15]};
  e.ep_return = w[17]; e.tt = word_to_int(w[18]); e.episode = word_to_int(w[19]);
}
template <typename T> __device__ __forceinline__ void hover_to_words(const HoverEnv<T>& e, T (&w)[kHoverWords]) {
  const Drone<T>& d = e.d;
  w[0] = d.px; w[1] = d.py; w[2] = d.pz; w[3] = d.wx; w[4] = d.vx; w[5] = d.vy; w[6] = d.vz; w[7] = d.wy;
  w[8] = d.q0; w[9] = d.q1; w[10] = d.q2; w[11] = d.q3; w[12] = d.u0; w[13] = d.u1; w[14] = d.u2; w[15] = d.u3;
  w[16] = d.wz; w[17] = e.ep_return; w[18] = int_to_word<T>(e.tt); w[19] = int_to_word<T>(e.episode);
}
template <typename T> __device__ __forceinline__ void store_obs13(void* base, int64_t i, const T* v) {
  T* o = reinterpret_cast<T*>(base) + i * 13;
#pragma unroll
  for (int k = 0; k < 13; k++) o[k] = v[k];
}
// [N,13] rows are 52 (fp32) / 104 (fp64) bytes: per-thread scalar stores would touch 13 sectors per instruction.  Each WARP
// stages its 32 rows in its own slice of shared memory (row stride 13 words: odd, so conflict-free) and writes the warp's
// contiguous 32*13 words with 13 fully coalesced stores; only __syncwarp is needed (no CTA-wide barrier on the step path).
// Must be called by every thread of the warp (tile0 = first env of the CTA's tile, cnt = valid envs in it).
template <typename T> __device__ __forceinline__ void store_obs13_tile(T* stage, void* base, int64_t tile0, int cnt, const T* v, bool active) {
  const int lane = threadIdx.x & 31, w0 = threadIdx.x & ~31;  // first row of this warp within the tile
  T* ws = stage + w0 * 13;
  __syncwarp();  // the previous use of this warp's slice is finished
  if (active) {
#pragma unroll
    for (int k = 0; k < 13; k++) ws[lane * 13 + k] = v[k];
  }
  __syncwarp();
  const int rows = cnt - w0 < 32 ? cnt - w0 : 32;  // valid rows of this warp (<= 0: none)
  T* o = reinterpret_cast<T*>(base) + (tile0 + w0) * 13;
#pragma unroll
  for (int k = 0; k < 13; k++) {
    const int idx = k * 32 + lane;
    if (idx < rows * 13) o[idx] = ws[idx];
  }
}

// info['chaser'] / info['target'] of an episode's last step (docking_env.py:226-229): the 13-state in float64 world
// coordinates, written before the auto-reset re-seats the env.  Cold: ~1.4 % of env-steps under random actions.
template <typename T> __device__ __forceinline__ void store_terminal_state(double* out, int64_t i, const Drone<T>& d, double ox, double oy, double oz) {
  double* o = out + i * 13;
  o[0] = (double)d.px + ox; o[1] = (double)d.py + oy; o[2] = (double)d.pz + oz;
  o[3] = (double)d.vx; o[4] = (double)d.vy; o[5] = (double)d.vz;
  o[6] = (double)d.q0; o[7] = (double)d.q1; o[8] = (double)d.q2; o[9] = (double)d.q3;
  o[10] = (double)d.wx; o[11] = (double)d.wy; o[12] = (double)d.wz;
}

// =========================================================================================================
// Fused docking step: gym_docking/envs/docking_env.py:104-231 (KIND = QS_DOCK_V0) and
// moving_docking_env.py:111-192 (KIND = QS_DOCK_V2) + the stable-baselines worker's auto-reset.
// The arithmetic is qs::dock_env_step / dock_env_seat (qs_core.cuh); this kernel is load -> step -> store.
// =========================================================================================================
// What a thread contributes to the per-warp statistics after its step.
struct StepTail { double ret; int len; bool done, docking, nonfinite; };

#ifdef QS_NO_SPEC64
constexpr bool kSpec64 = false;
#else
constexpr bool kSpec64 = kFast64;  // -DQS_NO_SPEC64: the guarded fp64 arithmetic on the main path (A/B)
#endif

template <int KIND, typename P> __device__ __noinline__ StepTail dock_step_replay(const EnvStepArgs& a, const P& prm, int64_t i);
// kernel-parameter type of the model constants: DefaultParams (empty: every constant an immediate) or RuntimeParams (the
// handle's qs_env_params_t, read from the constant bank)
template <typename T, int KIND, bool CUSTOM> struct PrmSel { using type = DefaultParams<T, KIND>; };
template <typename T, int KIND> struct PrmSel<T, KIND, true> { using type = RuntimeParams<T>; };

// step -> outputs -> auto-reset -> store for one env whose state words and action are in registers.  SPEC (fp64 only):
// the speculative straight-line step (qs::sdouble); an env that fails its check is replayed out of line from its state
// in HBM, which nothing has overwritten yet.
template <typename T, int KIND, bool SPEC, typename P>
__device__ __forceinline__ StepTail dock_step_loaded(const EnvStepArgs& a, const P& prm, int64_t i, T (&w)[dock_words<T>()], const T (&act)[4]) {
  constexpr int NW = dock_words<T>();
  DockEnv<T> e;
  words_to_dock<T, NW>(w, e);
  T rel[12];
  StepOut<T> o;
  if constexpr (SPEC) {
    if (!dock_env_step_spec<KIND>(prm, e, act, rel, o)) return dock_step_replay<KIND, P>(a, prm, i);
  } else {
    dock_env_step<T, KIND>(prm, e, act, rel, o);
  }
  reinterpret_cast<T*>(a.reward)[i] = o.reward;
  a.done[i] = o.done ? 1 : 0;
  if (a.info) a.info[i] = (uint8_t)o.info();
  if (a.term_obs) store_row<T, 12>(a.term_obs, i, rel);
  const StepTail tail = {(double)e.last_shaping, e.tt, o.done, o.docking, o.nonfinite};  // ret == sum of the episode's rewards (telescoping)
  if (o.done) {
    if (a.term_chaser) store_terminal_state<T>(a.term_chaser, i, e.c, (double)prm.ox, (double)prm.oy, (double)prm.oz);
    if (a.term_target) store_terminal_state<T>(a.term_target, i, e.t, (double)prm.ox, (double)prm.oy, (double)prm.oz);
    if (a.auto_reset) {
      e.episode += 1;
      dock_env_seat<T>(prm, e, a.rng, i, rel);
    }
  }
  store_row<T, 12>(a.obs, i, rel);
  dock_to_words<T, NW>(e, w);
  store_words<T, NW>(a.state, i, w);
  return tail;
}
template <int KIND, typename P> __device__ __noinline__ StepTail dock_step_replay(const EnvStepArgs& a, const P& prm, int64_t i) {
  constexpr int NW = dock_words<double>();
  double w[NW], act[4];
  load_words<double, NW>(a.state, i, w);
  load_action<double>(a.actions, i, act);
  return dock_step_loaded<double, KIND, false, P>(a, prm, i, w, act);
}
template <typename T, typename P> constexpr bool dock_spec() { return !is_f32<T>::value && kSpec64; }

}  // namespace
