// qs_policy.cu -- the reference's PPO2 MlpPolicy on 5th-gen tensor cores (tcgen05 + TMEM): fused forward pass for rollouts,
// fused forward + loss + backward for the update, and the device-resident rollout loop around them.
//
// Policy: net_arch=[128, dict(pi=[128], vf=[128])], relu, diagonal Gaussian (run_moving_docking_ppo2.py:88-91;
// weights as in the shipped archives, e.g. trained_model/best_model_v0.zip):
//     h  = relu(obs[12] W1^T + b1)                          12 -> 128   (shared_fc0)
//     hp = relu(h Wp^T + bp) ; hv = relu(h Wv^T + bv)       128 -> 128  (pi_fc0, vf_fc0)
//     mean = hp W3p^T + b3p (4) ; value = hv . w3v + b3v    128 -> 4, 1 (pi, vf)
//     action = mean + exp(logstd) * N(0,1) ; neglogp        (model.step, rl_baselines/ppo2/ppo2.py:472)
// This is the CALLER of the env hot path: with the env step at ~66 us per 2^20 envs, the unfused torch
// forward (1.5-4 ms) dominated rollout-in-the-loop by 20-60x.  It is the only dense contraction in the system, so all three
// layers go on the tensor cores; one 128-env tile = one M=128 UMMA tile:
//   layer 1:  D1[128x128] = A1[128x32]  . B1[128x32]^T    obs split hi+lo in bf16 (x = hi + lo, |lo| <= 2^-9 |x|) against W1
//             duplicated along K, so inputs keep ~16 mantissa bits (a plain bf16 cast would quantise a 1.8 m range to 8 mm);
//             two spare K slots carry 1.0 against (b1_hi, b1_lo): the bias add rides the MMA.  A1, B1 in shared memory.
//   layer 2:  D2[128x256] = A2[128x144] . [Wp;Wv | b_hi b_lo]^T   one N=256 accumulator, 9 tcgen05.mma (K=16 each); A2 in
//             shared memory (canonical no-swizzle K-major UMMA layout: 8x8 core matrices, 8 rows x 16 bytes contiguous).
//   layer 3:  D3[128x16]  = A3[128x256] . B3[16x256]^T    rows 0-3 = W3p over the pi half, row 4 = w3v over the vf half.
//             A3 never touches shared memory: relu(D2) is packed to bf16 and written back INTO TENSOR MEMORY over the
//             accumulator columns it came from (tcgen05.st), and the MMA takes its A operand from TMEM (the .ts form).
// The CUDA cores only move accumulators: tcgen05.ld -> cvt.rn.relu.bf16x2 (relu and the bf16 pack are ONE instruction per
// two elements) -> st.shared / tcgen05.st, then Philox/Box-Muller sampling and the stores.  Shared-memory traffic per tile
// is 176 KB (operand reads by the tensor cores + the A1/A2 stores); with A3 in shared memory as well (304 KB) the
// shared-memory pipe was the busiest unit (59 % at 127 us per 2^20 envs, profiles/r1_policy_v2_ncu_full.txt).
//
// Two kernels share these operand images and helpers:
//   policy_kernel<TRAIN>   one persistent 256-thread CTA per SM = two independent 128-thread GROUPS, each with its own 36 KB
//       A-operand buffer, 256 TMEM columns (D1, D2, A3, D3 alias: each is dead when the next is written), a named barrier
//       and an mbarrier; the weight images (88 KB, bulk-copied once per CTA with cp.async.bulk) are shared.  While one group
//       waits on its MMAs the other runs its epilogue.  TRAIN = true is the PPO2 minibatch gradient (forward + loss +
//       backward through the activations, see the kernel); TRAIN = false is the earlier rollout forward, kept for A/B.
//   policy_fwd_ws_kernel   the rollout forward that qs_policy_step launches: the same two groups plus one MMA-issuer warp per
//       group and mbarrier hand-offs (see the kernel).
// file map: operand layout + PTX wrappers | pack_kernel (weights -> operand images) | policy_kernel<TRAIN> |
//           policy_fwd_ws_kernel | C ABI (qs_policy_*, qs_rollout)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/quadsim_b200.h"
#include "qs_core.cuh"
#include "qs_env_kernels.cuh"

extern int qs_fail_(int code, const char* fmt, ...);  // qs_lib.cu: sets the thread-local error string
extern int qs_env_fused_args_(qs_env_t* e, void* out, size_t out_size, int* kind);  // qs_lib.cu: EnvStepArgs of a whole fp32 docking handle

namespace {

constexpr int kTileRows = 128;                       // envs per tile == UMMA M == TMEM lanes
constexpr int kGroupThreads = kTileRows, kGroups = 2, kThreads = kGroupThreads * kGroups;
constexpr int kObs = 12, kHid = 128, kAct = 4;
constexpr int kK1 = 32;    // layer-1 K: [hi(12) 1 1 0 0 | lo(12) 0 0 0 0]
constexpr int kK2 = 144;   // layer-2 K: h(128) 1 1 0...
constexpr int kK3 = 256;   // layer-3 K: [hp(128) | hv(128)]
constexpr int kN3 = 16;    // layer-3 N (UMMA minimum for M=128): mean(4), value, zero padding

// ---- shared-memory map (bytes); [W1 | W2 | W3 | tail] is also the packed global image ----
constexpr int kOffW1 = 0;                                  // B1 [128 x 32] bf16
constexpr int kOffW2 = kOffW1 + kHid * kK1 * 2;            // B2 [256 x 144] bf16
constexpr int kOffW3 = kOffW2 + 2 * kHid * kK2 * 2;        // B3 [16 x 256] bf16
constexpr int kOffTail = kOffW3 + kN3 * kK3 * 2;           // fp32: b3p(4), b3v, pad(3), logstd(4), pad(4)
constexpr int kTailFloats = 16;
constexpr int kPackedBytes = kOffTail + kTailFloats * 4;
constexpr int kOffA = (kPackedBytes + 127) / 128 * 128;    // per group: A2 [128 x 144] bf16 (36 KB); A1 / dOut alias its head
constexpr int kABytes = kTileRows * kK2 * 2;
constexpr int kOffBar = kOffA + kGroups * kABytes;         // weight mbarrier, one mbarrier per group, TMEM base slot
constexpr int kSmemBytes = kOffBar + 64;
// ---- tensor-memory map, columns relative to the group's 256-column window ----
constexpr uint32_t kTmemCols = 512, kTmemColsPerGroup = 256;
constexpr uint32_t kColD3 = 128;                           // 16 columns just behind the packed A3 (consumed D2 columns)
constexpr uint32_t kColD5 = 128;                           // backward: dA3 halves, then dA2 [128 x 128], behind the packed bf16 operand
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(kPackedBytes % 16 == 0, "bulk copies move 16-byte units");

// canonical K-major, no-swizzle layout: element (row, k) of an operand with R rows
__host__ __device__ constexpr int canon_off(int row, int k, int rows) {
  return ((k >> 3) * (rows >> 3) + (row >> 3)) * 128 + (row & 7) * 16 + (k & 7) * 2;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(kGroupThreads) : "memory"); }
// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout_type = SWIZZLE_NONE [61,64).  No-swizzle canonical layouts (cute mma_traits_sm100.hpp, T = 8 bf16 = 16 bytes):
//   K-major : ((8,m),(T,2)) : ((1T,SBO),(1,LBO))      8 rows x 16 B core matrix; SBO = next 8 rows, LBO = next 8 k
//   MN-major: ((T,1,m),(8,k)) : ((1,T,SBO),(1T,LBO))  the same 128 B core matrix read transposed: 8 MN contiguous x 8 k;
//                                                     SBO = next 8 MN, LBO = next 8 k
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), A K-major,
// B K-major or MN-major (bit 16)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool b_mn_major = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]^T
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T : A is 128 lanes x K/2 columns, column j of lane i = (A[i][2j] : low half, A[i][2j+1] : high half)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns: thread l of the warp gets lane (warp%4)*32 + l.  No wait: pair with tmem_wait_ld().
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 16 consecutive 32-bit columns from registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),
               "r"(r[14]), "r"(r[15])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// {bf16(max(hi,0)) : bf16(max(lo,0))} from two fp32 accumulator words.  volatile: must stay behind tcgen05.wait::ld.
__device__ __forceinline__ uint32_t relu_pack_bf16(uint32_t lo_f32_bits, uint32_t hi_f32_bits) {
  uint32_t d;
  asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_f32_bits)), "f"(__uint_as_float(lo_f32_bits)));
  return d;
}
// relu backward: {bf16(hi) : bf16(lo)} of two fp32 gradient words, each half zeroed where the forward activation's half
// (a packed, non-negative bf16x2 word) is zero
__device__ __forceinline__ uint32_t relu_grad_pack_bf16(uint32_t lo_f32_bits, uint32_t hi_f32_bits, uint32_t act_bf16x2) {
  uint32_t d;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi_f32_bits)), "f"(__uint_as_float(lo_f32_bits)));
  return d & __vcmpne2(act_bf16x2, 0u);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}

// ---- one-time weight packing: torch layout [out, in] fp32 -> canonical bf16 operand images + fp32 tail ----
struct RawWeights { const float *w1, *b1, *wp, *bp, *wv, *bv, *w3p, *b3p, *w3v, *b3v, *logstd; };

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

// fp32 image of the same weights for the accurate forward (policy_fwd_f32_kernel), in floats: K-major so that a warp reads
// consecutive output features -- W1T [12][128] | b1 [128] | W2T [128][256] (columns 0-127 pi, 128-255 vf) | b2 [256] |
// W3Q [256][4] (feature n < 128: W3p[0..3][n]; n >= 128: (w3v[n-128], 0, 0, 0)) | tail [16] as in the bf16 image
constexpr int kF32W1T = 0, kF32B1 = kF32W1T + kObs * kHid, kF32W2T = kF32B1 + kHid, kF32B2 = kF32W2T + kHid * 2 * kHid;
constexpr int kF32W3Q = kF32B2 + 2 * kHid, kF32Tail = kF32W3Q + 2 * kHid * 4, kF32Floats = kF32Tail + kTailFloats;
static_assert(kF32Floats % 4 == 0, "the image is copied in 16-byte units");

__global__ void pack_kernel(RawWeights w, unsigned char* packed, float* img32) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < kF32Floats; i += nth) {
    float v = 0.f;
    if (i < kF32B1) { const int k = i / kHid, n = i % kHid; v = w.w1[n * kObs + k]; }
    else if (i < kF32W2T) v = w.b1[i - kF32B1];
    else if (i < kF32B2) { const int k = (i - kF32W2T) / (2 * kHid), n = (i - kF32W2T) % (2 * kHid); v = n < kHid ? w.wp[n * kHid + k] : w.wv[(n - kHid) * kHid + k]; }
    else if (i < kF32W3Q) { const int n = i - kF32B2; v = n < kHid ? w.bp[n] : w.bv[n - kHid]; }
    else if (i < kF32Tail) { const int n = (i - kF32W3Q) / 4, j = (i - kF32W3Q) % 4; v = n < kHid ? w.w3p[j * kHid + n] : (j == 0 ? w.w3v[n - kHid] : 0.f); }
    else { const int t = i - kF32Tail; v = t < 4 ? w.b3p[t] : (t == 4 ? w.b3v[0] : (t >= 8 && t < 12 ? w.logstd[t - 8] : 0.f)); }
    img32[i] = v;
  }
  __nv_bfloat16* W1 = reinterpret_cast<__nv_bfloat16*>(packed + kOffW1);
  __nv_bfloat16* W2 = reinterpret_cast<__nv_bfloat16*>(packed + kOffW2);
  __nv_bfloat16* W3 = reinterpret_cast<__nv_bfloat16*>(packed + kOffW3);
  float* tail = reinterpret_cast<float*>(packed + kOffTail);
  for (int i = tid; i < kHid * kK1; i += nth) {  // B1[n][k]: the hi half and the lo half of the input see the same weights
    const int n = i / kK1, k = i % kK1, kk = k & 15;
    float v = 0.f;
    if (kk < kObs) v = w.w1[n * kObs + kk];
    else if (k == 12) v = w.b1[n];                            // against A1[:,12] = 1
    else if (k == 13) v = w.b1[n] - bf16_round(w.b1[n]);      // against A1[:,13] = 1: bias keeps ~16 mantissa bits
    W1[canon_off(n, k, kHid) / 2] = __float2bfloat16(v);
  }
  for (int i = tid; i < 2 * kHid * kK2; i += nth) {  // B2 = [Wp ; Wv], 256 rows, K = 128 weights + (bias_hi, bias_lo) + zero padding
    const int n = i / kK2, k = i % kK2;
    const float b = n < kHid ? w.bp[n] : w.bv[n - kHid];
    float v = 0.f;
    if (k < kHid) v = n < kHid ? w.wp[n * kHid + k] : w.wv[(n - kHid) * kHid + k];
    else if (k == kHid) v = b;
    else if (k == kHid + 1) v = b - bf16_round(b);
    W2[canon_off(n, k, 2 * kHid) / 2] = __float2bfloat16(v);
  }
  for (int i = tid; i < kN3 * kK3; i += nth) {  // B3: rows 0-3 = W3p over k < 128, row 4 = w3v over k >= 128
    const int n = i / kK3, k = i % kK3;
    float v = 0.f;
    if (n < kAct && k < kHid) v = w.w3p[n * kHid + k];
    else if (n == kAct && k >= kHid) v = w.w3v[k - kHid];
    W3[canon_off(n, k, kN3) / 2] = __float2bfloat16(v);
  }
  for (int i = tid; i < kTailFloats; i += nth) {
    float v = 0.f;
    if (i < 4) v = w.b3p[i];
    else if (i == 4) v = w.b3v[0];
    else if (i >= 8 && i < 12) v = w.logstd[i - 8];
    tail[i] = v;
  }
}

struct StepArgs {
  const unsigned char* packed;
  const float* obs;  // [N,12] (TRAIN: the whole rollout, rows picked through idx)
  int64_t n;         // rows this launch covers (TRAIN: the minibatch size)
  uint64_t seed, env_id0;
  uint32_t step_idx;
  int deterministic;
  float *actions, *clipped, *values, *neglogp;  // all but values may be null
};
// TRAIN only: PPO2's clipped-surrogate loss and its backward pass through the activations (qs_policy_ppo_grad)
struct TrainArgs {
  const float *actions, *returns, *values, *neglogp;  // rollout, flat; rows picked through idx
  const int64_t* idx;                                  // [n] or null (rows 0..n-1)
  const float* adv_stats;                              // device [2]: mean and 1/(std + 1e-8) of (returns - values) over the minibatch
  float cliprange, cliprange_vf, vf_coef, inv_batch;
  // bf16 operand IMAGES of the weight-gradient contraction (dw_kernel below): per 128-sample tile and operand one block in
  // the canonical no-swizzle MN-major UMMA layout -- the 16-byte chunk (8 features b*8..b*8+7) of sample row r sits at
  // tile * F * 256 + (b * 16 + (r >> 3)) * 128 + (r & 7) * 16 -- F = 32, 144, 256, 256, 128, 16 features
  uint4 *a1, *h, *hpv, *dz2, *dz1, *dout;
  float4* per_sample;                                  // [n,16] fp32 (optional: NULL skips the store)
  float* col_sums;                                     // [16] fp32: column sums of per_sample over the minibatch (atomics)
  const float4* packed;                                // optional: 128-byte rows (obs | action | return, value, neglogp), see pack_rollout_kernel
};
// where the 16-byte chunk `b` of tile row `r` of an F-feature operand image lives (in uint4 units)
__device__ __forceinline__ int64_t img_chunk(int64_t tile, int F, int r, int b) { return tile * (F * 16) + (b * 16 + (r >> 3)) * 8 + (r & 7); }

struct ObsRow { float4 a, b, c; };
__device__ __forceinline__ ObsRow load_obs(const float* obs, int64_t i, int64_t n, const int64_t* idx) {
  ObsRow o = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  if (i < n) {
    const int64_t src = idx ? __ldg(idx + i) : i;
    const float4* p = reinterpret_cast<const float4*>(obs + src * kObs);
    o.a = __ldg(p); o.b = __ldg(p + 1); o.c = __ldg(p + 2);
  }
  return o;
}

// the same through a row index that is already in a register (policy_kernel<TRAIN> keeps the minibatch indices one tile ahead)
__device__ __forceinline__ ObsRow load_obs_at(const float* obs, int64_t src, bool valid) {
  ObsRow o = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  if (valid) {
    const float4* p = reinterpret_cast<const float4*>(obs + src * kObs);
    o.a = __ldg(p); o.b = __ldg(p + 1); o.c = __ldg(p + 2);
  }
  return o;
}

// One kernel, two instantiations.  TRAIN = false: model.step (rollout).  TRAIN = true: the same forward pass over a minibatch
// gathered through idx, then PPO2's loss (rl_baselines/ppo2/ppo2.py:138-189) per sample and the backward pass through the
// activations, on the tensor cores as well:
//   dOut[128x16]  = (d loss / d mean(4), d loss / d value, 0...)                       per sample, fp32 math, bf16 operand
//   dA3 [128x256] = dOut . W3              B operand = the forward W3 image read MN-major (no transposed copy)
//   dZ2           = dA3 * relu'(hp|hv)     mask bits kept in registers from the forward epilogue; packed in place in TMEM
//   dA2 [128x128] = dZ2 . [Wp;Wv]          A from TMEM (.ts), B = the forward W2 image read MN-major
//   dZ1           = dA2 * relu'(h)
// The operands of the weight-gradient GEMMs (A1, h, hp|hv, dZ2, dZ1 as bf16 rows, plus 16 fp32 per sample: dOut, d logstd
// and the logged loss terms) go to HBM; those GEMMs reduce over the whole minibatch (K = batch) and are left to the caller.
// EXPERIMENT ONLY (-DQS_WS_TRACE=1): SM-clock stamps of the first kTrTiles tiles of CTA 0 / group 0 -- epilogue thread 0 in
// slots [0,16), the issuer's lane 0 in [16,24) -- read back with qs_debug_ws_trace (tools/ws_trace.py draws the timeline)
#ifndef QS_WS_TRACE
#define QS_WS_TRACE 0
#endif
#if QS_WS_TRACE
constexpr int kTrTiles = 28, kTrSlots = 24;
__device__ long long g_ws_trace[kTrTiles * kTrSlots];
#define QS_TR(slot) do { if (trace_on && it < (uint32_t)kTrTiles) g_ws_trace[it * kTrSlots + (slot)] = clock64(); } while (0)
#define QS_TRT(slot) do { if (TRAIN && trace_on && it < (uint32_t)kTrTiles) g_ws_trace[it * kTrSlots + (slot)] = clock64(); } while (0)
#else
#define QS_TR(slot) do { } while (0)
#define QS_TRT(slot) do { } while (0)
#endif

template <bool TRAIN>
__global__ void __launch_bounds__(kThreads, TRAIN ? 1 : 2) policy_kernel(const StepArgs a, const TrainArgs t) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = tid / kGroupThreads, row = tid % kGroupThreads;  // group; accumulator row == TMEM lane == env within the tile
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_w = sbase + kOffBar, bar_mma = sbase + kOffBar + 8 + 8 * g;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 32);
  const float* tail = reinterpret_cast<const float*>(smem + kOffTail);
  const int tiles = (int)((a.n + kTileRows - 1) / kTileRows);
  const int tile0 = blockIdx.x * kGroups + g, tile_stride = gridDim.x * kGroups;
  const int64_t* idx = TRAIN ? t.idx : nullptr;

  // Rows are gathered through idx (TRAIN): index -> row is a dependent pair of DRAM round trips (~3.5 k clocks each under this
  // kernel's store load; the trace showed 7.2 k clocks per tile exposed right here).  So the indices run TWO tiles ahead of their
  // use and the rows and records they select one tile ahead: nothing a tile needs is requested inside that tile.
  const int64_t i_first = (int64_t)tile0 * kTileRows + row, i_step = (int64_t)tile_stride * kTileRows;
  auto row_index = [&](int64_t i) -> int64_t { return i < a.n ? (idx ? __ldg(idx + i) : i) : 0; };
  struct Rec { float4 act; float ret, oldv, oldnlp; };  // TRAIN: a sample's rollout record (ppo2.py:244-262 feeds)
  const float4* packed = TRAIN ? t.packed : nullptr;    // one 128-byte line per sample instead of five arrays (qs_policy_pack_rollout)
  auto load_rec = [&](int64_t src, bool valid) {
    Rec r = {make_float4(0.f, 0.f, 0.f, 0.f), 0.f, 0.f, 0.f};
    if (TRAIN && valid) {
      if (packed) {
        r.act = __ldg(packed + src * 8 + 3);
        const float4 q = __ldg(packed + src * 8 + 4);
        r.ret = q.x; r.oldv = q.y; r.oldnlp = q.z;
      } else {
        r.act = __ldg(reinterpret_cast<const float4*>(t.actions) + src);
        r.ret = __ldg(t.returns + src); r.oldv = __ldg(t.values + src); r.oldnlp = __ldg(t.neglogp + src);
      }
    }
    return r;
  };
  auto load_row = [&](int64_t src, bool valid) {
    if (packed && valid) {
      ObsRow r;
      r.a = __ldg(packed + src * 8); r.b = __ldg(packed + src * 8 + 1); r.c = __ldg(packed + src * 8 + 2);
      return r;
    }
    return load_obs_at(a.obs, src, valid);
  };
  const int64_t src_first = tile0 < tiles ? row_index(i_first) : 0;
  int64_t src_next = tile0 < tiles ? row_index(i_first + i_step) : 0;
  ObsRow o = load_row(src_first, tile0 < tiles && i_first < a.n);  // in flight during the set-up
  Rec rec_next = load_rec(src_first, tile0 < tiles && i_first < a.n);        // rows AND records run one tile ahead of their use

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(sbase + kOffBar + 8, 1);
    mbar_init(sbase + kOffBar + 16, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_w, kPackedBytes);  // weights: [W1 | W2 | W3 | tail] contiguous in smem and in the packed image
    for (int off = 0; off < kPackedBytes; off += 16384) {
      const int bytes = kPackedBytes - off < 16384 ? kPackedBytes - off : 16384;
      bulk_g2s(sbase + kOffW1 + off, a.packed + off, bytes, bar_w);
    }
  }
  if (warp == 0) {  // all 512 TMEM columns: one CTA per SM, 256 per group
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot + g * kTmemColsPerGroup;          // this group's accumulator columns (lane field 0)
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's 32 TMEM lanes
  unsigned char* a_row = smem + kOffA + g * kABytes + row * 16;      // canon_off(row, 0, 128) == row * 16
  const uint32_t a_base = sbase + kOffA + g * kABytes;
  mbar_wait(bar_w, 0);
  // (a start skew between the groups, which takes 20 % off the rollout forward -- QS_POLICY_WS_SKEW -- measures no different
  // here: 14.26 vs 14.28-14.32 ms per epoch with 4 k / 12 k / 24 k clocks; this kernel waits on its image stores, not the pipe)
  constexpr uint32_t idesc1 = umma_idesc(128, kHid), idesc2 = umma_idesc(128, 2 * kHid), idesc3 = umma_idesc(128, kN3);
  constexpr uint32_t idesc4 = umma_idesc(128, kHid, true), idesc5 = umma_idesc(128, kHid, true);  // B read MN-major
  uint32_t phase = 0;
  float col_acc[13] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // TRAIN: this thread's share of col_sums
#if QS_WS_TRACE
  const bool trace_on = blockIdx.x == 0 && tid == 0;
#endif
  uint32_t it = 0;
  for (int tile = tile0; tile < tiles; tile += tile_stride, it++) {
    QS_TRT(0);
    const int64_t i = (int64_t)tile * kTileRows + row;
    const bool active = i < a.n;
    // ---- A1: the observation row split hi/lo, K = 32 = four 16-byte core-matrix rows ----
    {
      const float x[12] = {o.a.x, o.a.y, o.a.z, o.a.w, o.b.x, o.b.y, o.b.z, o.b.w, o.c.x, o.c.y, o.c.z, o.c.w};
      uint32_t hi[6], lo[6];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const float h0 = bf16_round(x[2 * k]), h1 = bf16_round(x[2 * k + 1]);
        hi[k] = pack_bf16(h0, h1);
        lo[k] = pack_bf16(x[2 * k] - h0, x[2 * k + 1] - h1);
      }
      const uint4 q0 = make_uint4(hi[0], hi[1], hi[2], hi[3]), q1 = make_uint4(hi[4], hi[5], 0x3F803F80u, 0u);  // k = 12, 13: 1.0 against (b1_hi, b1_lo)
      const uint4 q2 = make_uint4(lo[0], lo[1], lo[2], lo[3]), q3 = make_uint4(lo[4], lo[5], 0u, 0u);
      *reinterpret_cast<uint4*>(a_row + 0 * 2048) = q0;
      *reinterpret_cast<uint4*>(a_row + 1 * 2048) = q1;
      *reinterpret_cast<uint4*>(a_row + 2 * 2048) = q2;
      *reinterpret_cast<uint4*>(a_row + 3 * 2048) = q3;
      if (TRAIN) {  // rows past the minibatch (last tile) are stored as zeros: they must not contribute to the contraction
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        t.a1[img_chunk(tile, 32, row, 0)] = active ? q0 : z4; t.a1[img_chunk(tile, 32, row, 1)] = active ? q1 : z4;
        t.a1[img_chunk(tile, 32, row, 2)] = active ? q2 : z4; t.a1[img_chunk(tile, 32, row, 3)] = active ? q3 : z4;
      }
    }
    proxy_fence();       // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    tc_fence_before();   // orders the previous tile's tcgen05.ld before the barrier (TMEM columns are reused)
    QS_TRT(1);
    group_sync(g);
    QS_TRT(2);
    if (row == 0) {
      tc_fence_after();
      // operand rows = 128 -> SBO = 128 B, LBO = 16 row-groups * 128 B = 2048 B; one K=16 step = two k-blocks = 4096 B
#pragma unroll
      for (int ks = 0; ks < kK1 / 16; ks++)
        umma_ss(tmem, umma_desc(a_base + ks * 4096, 2048, 128), umma_desc(sbase + kOffW1 + ks * 4096, 2048, 128), idesc1, ks);
      umma_commit(bar_mma);
    }
    // this sample's record was requested a tile ago; the next tile's row and record go out now, under this tile's work
    const float4 act4 = rec_next.act;
    const float ret = rec_next.ret, oldv = rec_next.oldv, oldnlp = rec_next.oldnlp;
    o = load_row(src_next, i + i_step < a.n);
    rec_next = load_rec(src_next, i + i_step < a.n);
    src_next = row_index(i + 2 * i_step);  // consumed a whole tile from now
    QS_TRT(3);
    mbar_wait(bar_mma, phase);
    phase ^= 1;
    tc_fence_after();
    QS_TRT(4);
    // ---- epilogue 1: h = relu(D1) -> A2[:, 0:128]; A2[:, 128:144] = (1, 1, 0, ...) carries layer 2's bias ----
#pragma unroll 1
    for (int c0 = 0; c0 < kHid; c0 += 64) {
      uint32_t r[64];
      tmem_ld32(trow + c0, r);
      tmem_ld32(trow + c0 + 32, r + 32);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; j++) {  // one 8x8 core-matrix row (8 k = 16 bytes) per store; k-blocks are 2048 B apart for 128 rows
        uint4 w;
        w.x = relu_pack_bf16(r[8 * j + 0], r[8 * j + 1]);
        w.y = relu_pack_bf16(r[8 * j + 2], r[8 * j + 3]);
        w.z = relu_pack_bf16(r[8 * j + 4], r[8 * j + 5]);
        w.w = relu_pack_bf16(r[8 * j + 6], r[8 * j + 7]);
        *reinterpret_cast<uint4*>(a_row + ((c0 >> 3) + j) * 2048) = w;
        if (TRAIN) t.h[img_chunk(tile, 144, row, (c0 >> 3) + j)] = active ? w : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    *reinterpret_cast<uint4*>(a_row + 16 * 2048) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(a_row + 17 * 2048) = make_uint4(0u, 0u, 0u, 0u);
    if (TRAIN) {  // the bias block goes out with h: dz2^T . [h | 1 1 0...] carries d bias in column 128
      t.h[img_chunk(tile, 144, row, 16)] = active ? make_uint4(0x3F803F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
      t.h[img_chunk(tile, 144, row, 17)] = make_uint4(0u, 0u, 0u, 0u);
    }
    proxy_fence();
    tc_fence_before();
    QS_TRT(5);
    group_sync(g);
    QS_TRT(6);
    if (row == 0) {
      tc_fence_after();
      // B rows = 256 -> LBO = 32 row-groups * 128 B = 4096 B, one K step = 8192 B
#pragma unroll
      for (int ks = 0; ks < kK2 / 16; ks++)
        umma_ss(tmem, umma_desc(a_base + ks * 4096, 2048, 128), umma_desc(sbase + kOffW2 + ks * 8192, 4096, 128), idesc2, ks);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, phase);
    phase ^= 1;
    tc_fence_after();
    QS_TRT(7);
    // ---- epilogue 2: relu(D2) -> bf16 pairs written back over the columns they came from: A3 = columns [0,128).  Chunk c's
    //      packed columns land in columns this thread has already loaded, so nothing is overwritten before it is read. ----
#pragma unroll 1
    for (int c0 = 0; c0 < 2 * kHid; c0 += 64) {
      uint32_t r[64], w[32];
      tmem_ld32(trow + c0, r);
      tmem_ld32(trow + c0 + 32, r + 32);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; j++) w[j] = relu_pack_bf16(r[2 * j], r[2 * j + 1]);
      tmem_st16(trow + (c0 >> 1), w);
      tmem_st16(trow + (c0 >> 1) + 16, w + 16);
      if (TRAIN) {
#pragma unroll
        for (int j = 0; j < 8; j++)
          t.hpv[img_chunk(tile, 256, row, (c0 >> 3) + j)] = active ? make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
    tmem_wait_st();
    tc_fence_before();
    QS_TRT(8);
    group_sync(g);
    QS_TRT(9);
    if (row == 0) {
      tc_fence_after();
      // A from tensor memory: one K=16 step = 8 columns.  B rows = 16 -> LBO = 2 row-groups * 128 B = 256 B, one K step = 512 B
#pragma unroll
      for (int ks = 0; ks < kK3 / 16; ks++) umma_ts(tmem + kColD3, tmem + 8 * ks, umma_desc(sbase + kOffW3 + ks * 512, 256, 128), idesc3, ks);
      umma_commit(bar_mma);
    }
    // ---- Gaussian noise while layer 3 runs (DiagGaussianProbabilityDistribution.sample) ----
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
    if (!TRAIN && !a.deterministic) {
      qs::U4 c = {(uint32_t)(a.env_id0 + (uint64_t)i), (uint32_t)((a.env_id0 + (uint64_t)i) >> 32), a.step_idx, 2u};
      qs::U4 x = qs::philox4x32_10(c, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
      const float s = 5.9604644775390625e-08f;  // 2^-24; u in (0,1)
      const float u0 = ((float)(x.x >> 8) + 0.5f) * s, u1 = ((float)(x.y >> 8) + 0.5f) * s;
      const float u2 = ((float)(x.z >> 8) + 0.5f) * s, u3 = ((float)(x.w >> 8) + 0.5f) * s;
      const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
      float s0, c0_, s1, c1_;
      __sincosf(6.283185307179586f * u1, &s0, &c0_);
      __sincosf(6.283185307179586f * u3, &s1, &c1_);
      n0 = r0 * c0_; n1 = r0 * s0; n2 = r1 * c1_; n3 = r1 * s1;  // Box-Muller
    }
    const float ls0 = tail[8], ls1 = tail[9], ls2 = tail[10], ls3 = tail[11];
    mbar_wait(bar_mma, phase);
    phase ^= 1;
    tc_fence_after();
    QS_TRT(10);
    uint32_t d3[8];
    tmem_ld8(trow + kColD3, d3);
    tmem_wait_ld();
    const float mu0 = __uint_as_float(d3[0]) + tail[0], mu1 = __uint_as_float(d3[1]) + tail[1];
    const float mu2 = __uint_as_float(d3[2]) + tail[2], mu3 = __uint_as_float(d3[3]) + tail[3];
    const float val = __uint_as_float(d3[4]) + tail[4];
    if (!TRAIN) {
      // ---- sample, log-prob, store ----
      if (active) {
        const float a0 = fmaf(__expf(ls0), n0, mu0), a1 = fmaf(__expf(ls1), n1, mu1), a2 = fmaf(__expf(ls2), n2, mu2), a3 = fmaf(__expf(ls3), n3, mu3);
        // neglogp = 0.5 * sum(((a - mean)/std)^2) + 0.5*log(2 pi)*4 + sum(logstd); (a - mean)/std == the noise itself
        const float nlp = 0.5f * (n0 * n0 + n1 * n1 + n2 * n2 + n3 * n3) + 3.6757541328186907f + (ls0 + ls1 + ls2 + ls3);
        if (a.actions) reinterpret_cast<float4*>(a.actions)[i] = make_float4(a0, a1, a2, a3);
        if (a.clipped)
          reinterpret_cast<float4*>(a.clipped)[i] = make_float4(fminf(fmaxf(a0, -1.f), 1.f), fminf(fmaxf(a1, -1.f), 1.f), fminf(fmaxf(a2, -1.f), 1.f), fminf(fmaxf(a3, -1.f), 1.f));
        a.values[i] = val;
        if (a.neglogp) a.neglogp[i] = nlp;
      }
    } else {
      // ---- PPO2 loss of this sample and its gradient w.r.t. the network outputs (ppo2.py:138-189; `train_model.proba
      //      _distribution.neglogp`, the clipped value loss and the clipped surrogate).  Everything is scaled by 1/batch:
      //      the losses are means over the minibatch. ----
      float dm0 = 0.f, dm1 = 0.f, dm2 = 0.f, dm3 = 0.f, dval = 0.f, dl0 = 0.f, dl1 = 0.f, dl2 = 0.f, dl3 = 0.f;
      float pg = 0.f, vf = 0.f, kl = 0.f, cf = 0.f;
      if (active) {
        const float is0 = __expf(-ls0), is1 = __expf(-ls1), is2 = __expf(-ls2), is3 = __expf(-ls3);
        const float z0 = (act4.x - mu0) * is0, z1 = (act4.y - mu1) * is1, z2 = (act4.z - mu2) * is2, z3 = (act4.w - mu3) * is3;
        const float nlp = 0.5f * (z0 * z0 + z1 * z1 + z2 * z2 + z3 * z3) + 3.6757541328186907f + (ls0 + ls1 + ls2 + ls3);
        const float adv = ((ret - oldv) - __ldg(t.adv_stats)) * __ldg(t.adv_stats + 1);  // ppo2.py:265-266
        const float ratio = __expf(oldnlp - nlp);
        const float rc = fminf(fmaxf(ratio, 1.f - t.cliprange), 1.f + t.cliprange);
        const float p1 = -adv * ratio, p2 = -adv * rc;
        pg = fmaxf(p1, p2);
        // d max(p1, p2) / d ratio: p1 != p2 only where the clip is active (d p2 = 0); equal values share the same slope
        const float g_ratio = p1 >= p2 ? -adv : 0.f;
        const float g_nlp = g_ratio * (-ratio) * t.inv_batch;
        dm0 = g_nlp * (-z0 * is0); dm1 = g_nlp * (-z1 * is1); dm2 = g_nlp * (-z2 * is2); dm3 = g_nlp * (-z3 * is3);
        dl0 = g_nlp * (1.f - z0 * z0); dl1 = g_nlp * (1.f - z1 * z1); dl2 = g_nlp * (1.f - z2 * z2); dl3 = g_nlp * (1.f - z3 * z3);
        float e1 = val - ret, e2 = e1, inside = 1.f;
        if (t.cliprange_vf >= 0.f) {
          const float dv = val - oldv;
          e2 = (oldv + fminf(fmaxf(dv, -t.cliprange_vf), t.cliprange_vf)) - ret;
          inside = fabsf(dv) <= t.cliprange_vf ? 1.f : 0.f;
        }
        const float v1 = e1 * e1, v2 = e2 * e2;
        vf = 0.5f * fmaxf(v1, v2);
        dval = t.vf_coef * t.inv_batch * (v1 >= v2 ? e1 : e2 * inside);
        kl = 0.5f * (nlp - oldnlp) * (nlp - oldnlp);
        cf = fabsf(ratio - 1.f) > t.cliprange ? 1.f : 0.f;
        if (t.per_sample) {
          float4* ps = t.per_sample + i * 4;
          ps[0] = make_float4(dm0, dm1, dm2, dm3); ps[1] = make_float4(dval, 0.f, 0.f, 0.f);
          ps[2] = make_float4(dl0, dl1, dl2, dl3); ps[3] = make_float4(pg, vf, kl, cf);
        }
      }
      {  // column sums over the minibatch (bias gradients of the heads, d logstd, the logged loss terms): summed per thread over
         // its tiles here, reduced over the warp and added to global memory ONCE per warp after the tile loop (an atomic per
         // tile and warp put 680 k same-address atomics into every minibatch: +60 % kernel time)
        const float cs[13] = {dm0, dm1, dm2, dm3, dval, dl0, dl1, dl2, dl3, pg, vf, kl, cf};
#pragma unroll
        for (int k = 0; k < 13; k++) col_acc[k] += cs[k];
      }
      {  // dOut as a bf16 operand image row (16 features: d mean 0-3, d value, zeros) for d [W3p; w3v] = hpv^T . dOut
        t.dout[img_chunk(tile, 16, row, 0)] = make_uint4(pack_bf16(dm0, dm1), pack_bf16(dm2, dm3), pack_bf16(dval, 0.f), 0u);
        t.dout[img_chunk(tile, 16, row, 1)] = make_uint4(0u, 0u, 0u, 0u);
      }
      // dOut as the A operand of the first backward MMAs: [128 x 16] bf16 in k-blocks 16-17 of the A buffer (the bias block,
      // rewritten every tile); h stays intact in k-blocks 0-15: its zeros are relu' for the last step
      *reinterpret_cast<uint4*>(a_row + 16 * 2048) = make_uint4(pack_bf16(dm0, dm1), pack_bf16(dm2, dm3), pack_bf16(dval, 0.f), 0u);
      *reinterpret_cast<uint4*>(a_row + 17 * 2048) = make_uint4(0u, 0u, 0u, 0u);
      proxy_fence();
      QS_TRT(11);
      // dZ2 = (dOut . W3) * relu'(hp|hv), one 128-feature half at a time: the half's dA3 lands in columns [128,256), its
      // forward activations (bf16 pairs) are still in TMEM at [64 half, 64 half + 64), and dZ2 is packed over them in place
#pragma unroll 1
      for (int half = 0; half < 2; half++) {
        tc_fence_before();
        group_sync(g);
        if (row == 0) {
          tc_fence_after();
          // B = the W3 image (16 rows x 256 k, K-major) read MN-major: MN = feature (SBO = next 8 features = 256 B),
          // K = output index (LBO = next 8 outputs = 128 B); one K=16 step, N = 128 features of this half
          umma_ss(tmem + kColD5, umma_desc(a_base + 16 * 2048, 2048, 128), umma_desc(sbase + kOffW3 + half * 4096, 128, 256), idesc4, 0);
          umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase);
        phase ^= 1;
        tc_fence_after();
        QS_TRT(12 + 2 * half);
#pragma unroll 1
        for (int c0 = 0; c0 < kHid; c0 += 32) {
          uint32_t r[32], act[16], w[16];
          tmem_ld32(trow + kColD5 + c0, r);
          tmem_ld16(trow + 64 * half + (c0 >> 1), act);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; j++) w[j] = relu_grad_pack_bf16(r[2 * j], r[2 * j + 1], act[j]);
          tmem_st16(trow + 64 * half + (c0 >> 1), w);
#pragma unroll
          for (int j = 0; j < 4; j++)  // rows past the minibatch carry dOut = 0, hence zeros here
            t.dz2[img_chunk(tile, 256, row, 16 * half + (c0 >> 3) + j)] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
        }
        tmem_wait_st();
        QS_TRT(13 + 2 * half);
      }
      tc_fence_before();
      group_sync(g);
      QS_TRT(16);
      if (row == 0) {
        tc_fence_after();
        // dA2[128 x 128] = dZ2[128 x 256] . [Wp;Wv]: A from TMEM (8 columns per K=16 step); B = the W2 image (256 rows x 144 k)
        // read MN-major: MN = input feature k (SBO = next 8 k = 4096 B), K = output row n (LBO = next 8 rows = 128 B;
        // one K=16 step = 256 B)
#pragma unroll
        for (int ks = 0; ks < 2 * kHid / 16; ks++) umma_ts(tmem + kColD5, tmem + 8 * ks, umma_desc(sbase + kOffW2 + ks * 256, 128, 4096), idesc5, ks);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      QS_TRT(17);
      // ---- dZ1 = dA2 * relu'(h) -> bf16 rows in HBM; h is read back from this thread's row of the A buffer ----
#pragma unroll 1
      for (int c0 = 0; c0 < kHid; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + kColD5 + c0, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const uint4 hw = *reinterpret_cast<const uint4*>(a_row + ((c0 >> 3) + j) * 2048);
          uint4 w;
          w.x = relu_grad_pack_bf16(r[8 * j + 0], r[8 * j + 1], hw.x);
          w.y = relu_grad_pack_bf16(r[8 * j + 2], r[8 * j + 3], hw.y);
          w.z = relu_grad_pack_bf16(r[8 * j + 4], r[8 * j + 5], hw.z);
          w.w = relu_grad_pack_bf16(r[8 * j + 6], r[8 * j + 7], hw.w);
          t.dz1[img_chunk(tile, 128, row, (c0 >> 3) + j)] = w;
        }
      }
    }
  }
  if (TRAIN) {  // this thread's column sums -> warp tree -> one atomic per warp and column
#pragma unroll
    for (int k = 0; k < 13; k++) {
      float v = col_acc[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if ((tid & 31) == 0) atomicAdd(t.col_sums + (k < 5 ? k : k + 3), v);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(kTmemCols) : "memory");
}


// =========================================================================================================
// Weight-gradient contraction of the PPO2 update (the "dW GEMMs"): for a minibatch of B samples
//     d [Wp;Wv | bias]  g2[256 x 144] = dZ2^T . [h | 1 1 0...]        d W1 (hi | bias | lo)  g1[128 x 32] = dZ1^T . A1
//     d [W3p ; w3v]     g3[256 x 16]  = [hp|hv]^T . dOut
// All three reduce over the SAMPLE index (K = B ~ 1.7 M): split-K over one persistent CTA per SM, each walking a contiguous
// range of 128-sample tiles; fp32 accumulators live in tensor memory for the CTA's whole range (352 of 512 columns) and are
// added to the global results once at the end (red.global.add.f32).  Both operands of every product are stored
// sample-major, i.e. MN-major for the tensor core: policy_kernel<true> writes them as ready-made canonical no-swizzle UMMA
// images per tile, so a tile's operand is ONE contiguous block that an elected thread pulls in with cp.async.bulk (SASS
// UBLKCP) and the MMAs read through fixed descriptors (instruction-descriptor bits 15/16: A and B MN-major).
// The kernel is HBM-bound by construction: 208 KB of operands per tile against 40 tcgen05.mma (~2.2 k tensor-pipe clocks).
// Two shared-memory stages of ~100 KB each hold half a tile's operands (stage 0: dZ2 + h -> g2; stage 1: hpv + dOut + dZ1 + A1 ->
// g3, g1): warp 0 produces, warp 1 issues, tcgen05.commit hands a stage back to the producer.
// Replaces three library GEMMs (cuBLASLt nvjet_tss_* + splitKreduce in profiles/r1_launches_bench_final.txt).
// =========================================================================================================
constexpr int kDwImgDz2 = 256 * 256, kDwImgH = kK2 * 256, kDwImgHpv = 256 * 256, kDwImgDout = 16 * 256, kDwImgDz1 = 128 * 256, kDwImgA1 = kK1 * 256;
constexpr int kDwStage0 = kDwImgDz2 + kDwImgH;                           // 102400 B
constexpr int kDwStage1 = kDwImgHpv + kDwImgDout + kDwImgDz1 + kDwImgA1; // 110592 B
constexpr int kDwOff1 = kDwStage0;
constexpr int kDwOffBar = kDwOff1 + kDwStage1;
constexpr int kDwSmemBytes = kDwOffBar + 64;
static_assert(kDwSmemBytes <= 227 * 1024, "shared memory budget of dw_kernel");
constexpr uint32_t kDwColG2 = 0, kDwColG1 = 288, kDwColG3 = 320;  // g2: two M=128 halves x 144 columns; g1: 32; g3: two halves x 16

struct DwArgs {
  const unsigned char *a1, *h, *hpv, *dz2, *dz1, *dout;  // operand images (see TrainArgs)
  int tiles;
  float *g1, *g2, *g3;  // [128,32], [256,144], [256,16] fp32, zeroed by the caller
};

__host__ __device__ constexpr uint32_t umma_idesc_mn(int M, int N) {  // D = f32, A = B = bf16, both MN-major
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) dw_kernel(const DwArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + kDwOffBar, bar_empty = sbase + kDwOffBar + 16, bar_done = sbase + kDwOffBar + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kDwOffBar + 40);
  // this CTA's contiguous range of tiles
  const int lo = (int)((int64_t)a.tiles * blockIdx.x / gridDim.x), hi = (int)((int64_t)a.tiles * (blockIdx.x + 1) / gridDim.x);
  if (lo >= hi) return;  // (the launch never uses more CTAs than tiles)
  if (tid == 0) {
    for (int s = 0; s < 2; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- producer: two bulk-copy stages per tile
    auto pull = [&](uint32_t dst, const unsigned char* src, int bytes, uint32_t bar) {
      for (int off = 0; off < bytes; off += 16384) bulk_g2s(dst + off, src + off, bytes - off < 16384 ? bytes - off : 16384, bar);
    };
    for (int j = lo, n = 0; j < hi; j++, n++) {
      const uint32_t par = (n & 1) ^ 1;  // a fresh barrier passes a wait on parity 1: both stages start out empty
      mbar_wait(bar_empty, par);
      mbar_expect_tx(bar_full, kDwStage0);
      pull(sbase, a.dz2 + (int64_t)j * kDwImgDz2, kDwImgDz2, bar_full);
      pull(sbase + kDwImgDz2, a.h + (int64_t)j * kDwImgH, kDwImgH, bar_full);
      mbar_wait(bar_empty + 8, par);
      mbar_expect_tx(bar_full + 8, kDwStage1);
      uint32_t d = sbase + kDwOff1;
      pull(d, a.hpv + (int64_t)j * kDwImgHpv, kDwImgHpv, bar_full + 8); d += kDwImgHpv;
      pull(d, a.dout + (int64_t)j * kDwImgDout, kDwImgDout, bar_full + 8); d += kDwImgDout;
      pull(d, a.dz1 + (int64_t)j * kDwImgDz1, kDwImgDz1, bar_full + 8); d += kDwImgDz1;
      pull(d, a.a1 + (int64_t)j * kDwImgA1, kDwImgA1, bar_full + 8);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------------------------------------------------------- MMA issuer
    // MN-major images: 8 features x 8 samples per 128-byte core matrix; next 8 samples (LBO) = 128 B, next 8 features (SBO) =
    // 16 core matrices = 2048 B.  One K = 16 step = 16 samples = 256 B further into the image; an M = 128 half = 16 * 2048 B.
    constexpr uint32_t id_g2 = umma_idesc_mn(128, kK2), id_g1 = umma_idesc_mn(128, kK1), id_g3 = umma_idesc_mn(128, 16);
    for (int j = lo, n = 0; j < hi; j++, n++) {
      const uint32_t par = n & 1, acc0 = n > 0;
      mbar_wait(bar_full, par);
      tc_fence_after();
      {
        const uint32_t dz2 = sbase, hh = sbase + kDwImgDz2;
#pragma unroll
        for (int half = 0; half < 2; half++)
#pragma unroll
          for (int ks = 0; ks < 8; ks++)
            umma_ss(tmem + kDwColG2 + half * kK2, umma_desc(dz2 + half * 16 * 2048 + ks * 256, 128, 2048), umma_desc(hh + ks * 256, 128, 2048), id_g2, acc0 | (ks > 0));
      }
      umma_commit(bar_empty);  // stage 0 may be refilled once these MMAs have read it
      mbar_wait(bar_full + 8, par);
      tc_fence_after();
      {
        const uint32_t hpv = sbase + kDwOff1, dout = hpv + kDwImgHpv, dz1 = dout + kDwImgDout, a1 = dz1 + kDwImgDz1;
#pragma unroll
        for (int half = 0; half < 2; half++)
#pragma unroll
          for (int ks = 0; ks < 8; ks++)
            umma_ss(tmem + kDwColG3 + half * 16, umma_desc(hpv + half * 16 * 2048 + ks * 256, 128, 2048), umma_desc(dout + ks * 256, 128, 2048), id_g3, acc0 | (ks > 0));
#pragma unroll
        for (int ks = 0; ks < 8; ks++)
          umma_ss(tmem + kDwColG1, umma_desc(dz1 + ks * 256, 128, 2048), umma_desc(a1 + ks * 256, 128, 2048), id_g1, acc0 | (ks > 0));
      }
      umma_commit(bar_empty + 8);
    }
    umma_commit(bar_done);  // arrives when every MMA issued above has completed
  }
  __syncwarp();
  // ---------------------------------------------------------------- read-out: TMEM lane = feature row of the result
  mbar_wait(bar_done, 0);
  tc_fence_after();
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const int m = warp * 32 + lane;
  auto flush = [&](uint32_t col, int ncols, float* dst, int ld) {  // dst[(m) * ld + c] += D[m][col + c]
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(trow + col + c0, r);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 16; c++) atomicAdd(dst + (int64_t)m * ld + c0 + c, __uint_as_float(r[c]));
    }
  };
  flush(kDwColG2, kK2, a.g2, kK2);
  flush(kDwColG2 + kK2, kK2, a.g2 + 128 * kK2, kK2);
  flush(kDwColG1, kK1, a.g1, kK1);
  flush(kDwColG3, 16, a.g3, 16);
  flush(kDwColG3 + 16, 16, a.g3 + 128 * 16, 16);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(kTmemCols) : "memory");
}

// =========================================================================================================
// The optimiser step of the PPO2 update in ONE launch: tf.clip_by_global_norm + tf.train.AdamOptimizer(epsilon=1e-5)
// (rl_baselines/ppo2/ppo2.py:200-207) over the policy's ~35 k parameters.  One CTA: phase 1 = global gradient norm
// (fixed-order tree: bit-reproducible), phase 2 = clip and Adam in TensorFlow's formulation
//     lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m += (1 - b1)(g - m);  v += (1 - b2)(g^2 - v);  p -= lr_t * m / (sqrt(v) + eps)
// =========================================================================================================
constexpr int kAdamMaxTensors = 16;
struct AdamArgs {
  float* param[kAdamMaxTensors];
  const float* grad[kAdamMaxTensors];
  float *m[kAdamMaxTensors], *v[kAdamMaxTensors];
  int n[kAdamMaxTensors];
  int count;
  float lr_t, lr, beta1, beta2, eps, max_norm;
  float* norm_out;      // [1], may be null
  long long* step_dev;  // may be null: the step count lives on the device (graph replay), lr_t is derived in the kernel
};
constexpr int kAdamCtas = 16;
__global__ void __launch_bounds__(1024, 1) adam_kernel(const __grid_constant__ AdamArgs a) {
  // Every CTA computes the global norm itself (35 k gradient words, L2-resident: cheaper than a grid-wide reduction) and then
  // updates its own share of the parameters.  Tensors are walked four elements per thread and iteration so that the loads
  // of an iteration are independent (the single-CTA, one-element version was a 50 us chain of dependent global loads).
  __shared__ float red[1024];
  __shared__ float s_scale, s_lr_t;
  float ss = 0.f;
  for (int k = 0; k < a.count; k++) {
    const float* g = a.grad[k];
    const int n = a.n[k];
    for (int i = threadIdx.x * 4; i < n; i += 4096) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = i + j < n ? g[i + j] : 0.f;
#pragma unroll
      for (int j = 0; j < 4; j++) ss = fmaf(v[j], v[j], ss);
    }
  }
  red[threadIdx.x] = ss;
  __syncthreads();
  for (int off = 512; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = sqrtf(red[0]);
    // tf.clip_by_global_norm: g * clip / max(norm, clip); max_norm <= 0: no clipping
    s_scale = a.max_norm > 0.f ? a.max_norm / fmaxf(norm, a.max_norm) : 1.f;
    if (a.norm_out && blockIdx.x == 0) *a.norm_out = norm;
    s_lr_t = a.lr_t;
    if (a.step_dev) {  // every CTA derives lr_t from the counter as it was at launch; CTA 0 advances it (nobody reads it later in this launch)
      const long long t = *a.step_dev + 1;
      s_lr_t = (float)((double)a.lr * sqrt(1.0 - pow((double)a.beta2, (double)t)) / (1.0 - pow((double)a.beta1, (double)t)));
    }
  }
  __syncthreads();
  const float scale = s_scale, b1 = a.beta1, b2 = a.beta2, lr_t = s_lr_t;
  for (int k = 0; k < a.count; k++) {
    const int n = a.n[k];
    for (int i = blockIdx.x * 1024 + threadIdx.x; i < n; i += gridDim.x * 1024) {
      const float g = a.grad[k][i] * scale;
      const float m = a.m[k][i] + (1.f - b1) * (g - a.m[k][i]);
      const float v = a.v[k][i] + (1.f - b2) * (g * g - a.v[k][i]);
      a.m[k][i] = m; a.v[k][i] = v;
      a.param[k][i] -= lr_t * m / (sqrtf(v) + a.eps);
    }
  }
}
// the device step counter is advanced by its own tiny launch right behind adam_kernel (stream-ordered: no CTA of the
// optimiser launch can still be reading it)
__global__ void adam_advance_kernel(long long* step_dev) { *step_dev += 1; }

// ---- advantage statistics of a minibatch (ppo2.py:265-266: advs = returns - values; (advs - advs.mean()) / (advs.std() + 1e-8)) ----
// mean and 1 / (population std + 1e-8) of (returns - values)[idx] in ONE launch: per-block partial sums in double, and the last
// block to finish (ticket counter) adds the partials in block order -- bit-reproducible -- writes the two floats the gradient
// kernel reads, and re-arms the counter.  Replaces a gather, a Welford reduction and six elementwise launches of the caller.
constexpr int kAdvBlocks = 1184;  // 8 per SM: the gathers are latency-bound, so many loads in flight
__global__ void __launch_bounds__(256) pack_rollout_kernel(const float* __restrict__ obs, const float* __restrict__ act, const float* __restrict__ ret,
                                                           const float* __restrict__ val, const float* __restrict__ nlp, int64_t rows, float4* __restrict__ out) {
  // one 128-byte row per sample: [obs 12 | action 4 | return, value, neglogp, 0 | 0 x 12]; 8 consecutive threads write one row
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x, r = e >> 3;
  const int c = (int)(e & 7);
  if (r >= rows) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < 3) v = __ldg(reinterpret_cast<const float4*>(obs + r * kObs) + c);
  else if (c == 3) v = __ldg(reinterpret_cast<const float4*>(act) + r);
  else if (c == 4) v = make_float4(__ldg(ret + r), __ldg(val + r), __ldg(nlp + r), 0.f);
  out[e] = v;
}

__global__ void __launch_bounds__(256) adv_stats_kernel(const float* __restrict__ ret, const float* __restrict__ val, const int64_t* __restrict__ idx,
                                                        int64_t n, double* partial, unsigned int* ticket, float* stats, const float4* __restrict__ packed) {
  double s = 0.0, q = 0.0;
  const int64_t stride = (int64_t)gridDim.x * 256;
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {  // four independent gathers in flight per thread
    int64_t j[4];
    float a[4];
#pragma unroll
    for (int u = 0; u < 4; u++) j[u] = idx ? idx[i + u * stride] : i + u * stride;
#pragma unroll
    for (int u = 0; u < 4; u++) {  // float32, as the reference forms it
      if (packed) { const float4 q = __ldg(packed + j[u] * 8 + 4); a[u] = q.x - q.y; }
      else a[u] = ret[j[u]] - val[j[u]];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) { s += (double)a[u]; q += (double)a[u] * (double)a[u]; }
  }
  for (; i < n; i += stride) {
    const int64_t j = idx ? idx[i] : i;
    float a;
    if (packed) { const float4 q = __ldg(packed + j * 8 + 4); a = q.x - q.y; }
    else a = ret[j] - val[j];
    s += (double)a;
    q += (double)a * (double)a;
  }
  __shared__ double sh[2][8];
  __shared__ bool last;
  auto block_sum = [&](double& x, double& y) {  // fixed order: lanes by xor-shuffle, then the 8 warps in turn
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { x += __shfl_xor_sync(0xffffffffu, x, o); y += __shfl_xor_sync(0xffffffffu, y, o); }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = x; sh[1][threadIdx.x >> 5] = y; }
    __syncthreads();
    x = y = 0.0;
    for (int w = 0; w < 8; w++) { x += sh[0][w]; y += sh[1][w]; }
  };
  block_sum(s, q);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = s;
    partial[2 * blockIdx.x + 1] = q;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();  // the last block to finish adds the partials, every thread a fixed subset, and publishes the statistics
  double S = 0.0, Q = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += 256) { S += ((volatile double*)partial)[2 * b]; Q += ((volatile double*)partial)[2 * b + 1]; }
  block_sum(S, Q);
  if (threadIdx.x == 0) {
    const double mean = S / (double)n;
    double var = Q / (double)n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[0] = (float)mean;
    stats[1] = 1.0f / ((float)sqrt(var) + 1e-8f);
    *ticket = 0;
  }
}

// =========================================================================================================
// Minibatch shuffle of the PPO2 update (np.random.shuffle(inds), rl_baselines/ppo2/ppo2.py:350-352) without a sort: a keyed
// bijection of [0, n) -- a 4-round Feistel network on the next power of four above n, cycle-walked back into range -- so
// index i of the permutation is computed independently by one thread (torch.randperm = a radix sort of n random keys).
// =========================================================================================================
__device__ __forceinline__ uint32_t feistel_round(uint32_t x, uint32_t k) {
  x = (x ^ k) * 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13; x *= 0xC2B2AE3Du; x ^= x >> 16;
  return x;
}
__global__ void __launch_bounds__(256) permutation_kernel(int64_t n, int half_bits, uint64_t seed, int64_t* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t mask = (1u << half_bits) - 1u;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint64_t x = (uint64_t)i;
  do {  // cycle walking: the domain is at most 4x the range, so this loops 1-4 times on average
    uint32_t l = (uint32_t)(x >> half_bits) & mask, r = (uint32_t)x & mask;
#pragma unroll
    for (int rnd = 0; rnd < 4; rnd++) {
      const uint32_t f = feistel_round(r, (rnd & 1 ? k1 : k0) + 0x632BE5ABu * (uint32_t)(rnd + 1)) & mask;
      const uint32_t nl = r;
      r = l ^ f; l = nl;
    }
    x = ((uint64_t)l << half_bits) | r;
  } while (x >= (uint64_t)n);
  out[i] = (int64_t)x;
}

// =========================================================================================================
// Rollout forward, warp-specialised (the kernel qs_policy_step launches).  Same math, same operand images and the same TMEM
// tricks as policy_kernel<false>; what changes is WHO waits for WHAT.  With two 128-env tiles in flight per SM the kernel
// is a dependency chain (MMA -> read-out -> pack -> MMA ...), not a bandwidth problem (tensor pipe 38 % busy, TMEM read-out
// measured at 907 B/clk/SM vs ~50 used: profiles/r1_tmem_read_bandwidth.json), so the chain is shortened:
//   * a dedicated MMA-issuer warp per group (warps 8, 9): the 128 epilogue threads never stall behind the issue loop; they
//     signal "operand ready" with mbarrier arrivals and move on, the tensor pipe signals back with tcgen05.commit;
//   * layer 2 is issued as two N=128 halves with separate commits: the pi half is read out and packed while the vf half
//     is still in the tensor pipe, and layer 3's first 8 K-steps (pi) run while the vf half is being packed;
//   * the NEXT tile's A1 is built and its layer 1 issued before the current tile's layer-3 result is consumed (D1 lives in
//     the vf half's columns, free by then), so sampling and the output stores overlap the next tile's first MMA.
//   * optionally (QS_POLICY_WS_SIDES=2) two epilogue warps per TMEM lane quadrant, each reading out / packing half of every
//     128-column phase, side 0 sampling the action and side 1 storing the value -- measured SLOWER (96 vs 91 us): what is
//     left of the chain is barrier / commit / fence round trips, not the epilogues' instruction count.
// TMEM columns of a group's 256-column window (two sides): D1 [128,256) | D2 pi [0,128): side s reads [64s, 64s+64) and
// packs A3 pi into its own first 32 columns [64s, 64s+32); D3 [32,48) | D2 vf [128,256) likewise -> A3 vf [128+64s, +32).
// Every barrier completes exactly once per tile, so the wait parity is the tile counter's low bit.
// =========================================================================================================
#ifndef QS_EXP_D3_SPLIT
#define QS_EXP_D3_SPLIT 0
#endif
#ifndef QS_EXP_L3_STEPS
#define QS_EXP_L3_STEPS 8      // EXPERIMENT ONLY (results are wrong below 8): layer-3 K-steps issued per half
#endif
#ifndef QS_POLICY_WS_SIDES
#define QS_POLICY_WS_SIDES 1   // epilogue warps per TMEM lane quadrant and group: 1 or 2 (each side takes half the columns);
                               // measured 90.8 us (1) vs 96.4 us (2) per 2^20 envs: the chain is sync latency, not ALU work
#endif
constexpr int kWsSides = QS_POLICY_WS_SIDES;
#ifndef QS_POLICY_WS_SKEW
#define QS_POLICY_WS_SKEW 2000 // SM clocks by which group g delays its first tile (x g).  The groups start together and contend
                               // symmetrically for the in-order tensor pipe, which PRESERVES their relative phase: started in phase
                               // they stay in phase -- both wait for the pipe at the same time, then both run their epilogues while
                               // it idles (88 us per 2^20 envs).  Started half a tile apart, one group's epilogue runs under the
                               // other's layer 2: 77.5 us (700 .. 2800 clocks measure alike), 69.4 us together with kEarlyL1.
#endif
#ifndef QS_POLICY_WS_EARLY_L1
#define QS_POLICY_WS_EARLY_L1 1
#endif
#ifndef QS_POLICY_WS_GROUPS
#define QS_POLICY_WS_GROUPS 2  // tiles in flight per SM.  2: 256 TMEM columns per group, both layer-2 halves in the tensor pipe
                               // at once.  3: 144 columns per group -- every accumulator of a tile lives in [0,128) (+ D3 in
                               // [128,144)), so the vf half of layer 2 can only be issued once the pi half has been packed
#endif
constexpr int kWsGroups = QS_POLICY_WS_GROUPS;
static_assert(kWsGroups == 2 || (kWsGroups == 3 && kWsSides == 1), "two groups, or three with one epilogue warp per quadrant");
constexpr int kWsEpi = kTileRows * kWsSides;               // epilogue threads per group
constexpr int kWsThreads = kWsGroups * kWsEpi + 32 * kWsGroups; // + one issuer warp per group
constexpr uint32_t kWsWindow = kWsGroups == 2 ? 256 : 144; // TMEM columns per group
constexpr int kWsCols = kHid / kWsSides;                   // accumulator columns one thread reads out per 128-column phase
#ifndef QS_POLICY_WS_CHUNK
#define QS_POLICY_WS_CHUNK (QS_POLICY_WS_SIDES == 1 ? 64 : 32)
#endif
constexpr int kWsChunk = QS_POLICY_WS_CHUNK;          // ... per tcgen05.wait::ld (register budget: 576 threads x <= 112)
static_assert(kWsSides == 1 || kWsSides == 2, "one or two epilogue warps per quadrant");
constexpr int kWsOffBar = kOffA + kWsGroups * kABytes;
constexpr int kWsBarBytes = 8 + kWsGroups * 64 + 8;        // weights | per group: full_a1, full_a2, full_a3p, full_a3v, done_1, done_p, done_v, done_3 | TMEM slot
constexpr int kWsSmemBytes = kWsOffBar + (kWsBarBytes + 63) / 64 * 64;
static_assert(kWsSmemBytes <= 227 * 1024, "shared memory budget");
enum { kFullA1 = 0, kFullA2, kFullA3p, kFullA3v, kDone1, kDoneP, kDoneV, kDone3 };
// ---- fused policy-forward + env-step kernel (policy_env_step_kernel): the same CTA plus four ENV WARPS per group ----
// The epilogue threads hand the clipped actions of a tile to the group's env warps through shared memory (two 2 KB slots per
// group, full / empty mbarriers); the env warps hold the tile's state rows in registers by then (loads issued before the
// wait) and run the very device code of dock_step_kernel (dock_step_loaded, qs_env_kernels.cuh) on them.  One launch per
// rollout step does model.step AND env.step; the env warps use the issue slots the policy's dependency chain leaves idle
// (issue-active 28 % for the policy kernel alone), instead of a second kernel that can only share SMs by accident.
#ifndef QS_FUSED_CHUNK
#define QS_FUSED_CHUNK 32  // accumulator columns per read-out chunk of the fused kernel's policy half: at 64 the 96-register cap of the 576-thread CTA spills (2.49 vs 2.18 ms per rollout)
#endif
#ifndef QS_FUSED_ENV_SETS
#define QS_FUSED_ENV_SETS 1   // env warp-sets (4 warps each) per group: set s serves action slot s, i.e. every kFuSets-th tile of the
                              // group.  Two sets (plain coalesced state loads: their stages would not fit) measure SLOWER in every
                              // form tried -- 2.25 vs 2.18 ms per 16-step rollout at the uniform 72 registers of an 832-thread CTA,
                              // 3.06 ms with the registers split by role through setmaxnreg: the step is not bound by how many env
                              // tiles are in flight.
#endif
constexpr int kFuSets = QS_FUSED_ENV_SETS;
static_assert(kFuSets == 1 || kFuSets == 2, "one or two env warp-sets per group (two action slots)");
constexpr int kFuEnvThreads = kWsGroups * kFuSets * kTileRows; // one env thread per accumulator row and set
constexpr int kFuThreads = kWsThreads + kFuEnvThreads;        // 576
constexpr int kFuOffAct = kWsSmemBytes;                       // [group][slot][128] float4
constexpr int kFuOffBar = kFuOffAct + kWsGroups * 2 * kTileRows * 16;  // [group][slot]{full, empty}, then [group][set] state-stage barrier
constexpr int kFuStateBytes = (dock_words<float>() / 4) * kTileRows * 16;  // one tile's state rows: 20 KB, contiguous in HBM (tile-major layout)
constexpr int kFuOffState = (kFuOffBar + kWsGroups * 2 * 16 + kWsGroups * 2 * 8 + 127) / 128 * 128;  // [group][set] stage
constexpr bool kFuStage = kFuSets == 1;  // one set per group: bulk-copy state stage; two sets: plain coalesced loads (four stages do not fit)
constexpr int kFuSmemBytes = kFuOffState + (kFuStage ? kWsGroups * kFuSets * kFuStateBytes : 0);
constexpr bool kFuFits = kFuSmemBytes <= 227 * 1024;  // three groups' state stages do not fit: the fused step kernel is then unavailable
static_assert(kFuFits || kWsGroups == 3, "shared memory budget");

// programmatic dependent launch: pdl_prologue() of qs_env_kernels.cuh (stage the next kernel early, wait for the previous one)
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

template <bool FUSED, int KIND, int CHUNK = kWsChunk>
__device__ __forceinline__ void policy_fwd_ws_body(const StepArgs& a, const EnvStepArgs& ea) {
  pdl_prologue();
  extern __shared__ __align__(1024) unsigned char smem[];
  // FUSED: the env warps come FIRST (QS_FUSED_ENV_FIRST, default): the warp scheduler favours the higher warp slots, and it is
  // the policy's dependency chain (epilogue -> issuer -> tensor pipe -> epilogue) that sets the step time -- the env warps are
  // there to fill the issue slots that chain leaves idle, not to compete for them
#ifndef QS_FUSED_ENV_FIRST
#define QS_FUSED_ENV_FIRST 1
#endif
  constexpr int kEnvFirst = FUSED && QS_FUSED_ENV_FIRST ? kFuEnvThreads : 0;  // threads in front of the policy's
  const int lane = threadIdx.x & 31, tid_hw = threadIdx.x;
  const int tid = tid_hw - kEnvFirst, warp = tid >> 5;           // policy-relative ids (negative: env threads placed first)
  constexpr int kEpiWarps = kWsGroups * kWsEpi / 32;
  const bool env_warp = FUSED && (kEnvFirst ? tid < 0 : warp >= kEpiWarps + kWsGroups);
  const int env_tid = kEnvFirst ? tid_hw : tid - kWsThreads;     // 0 .. kFuEnvThreads-1 for env threads
  const bool issuer = !env_warp && warp >= kEpiWarps;
  const int g = env_warp ? env_tid / (kFuSets * kTileRows) : (issuer ? warp - kEpiWarps : tid / kWsEpi);  // group
  const int env_set = env_warp ? (env_tid / kTileRows) % kFuSets : 0;
  const int side = issuer || env_warp ? 0 : (tid % kWsEpi) / kTileRows; // which half of the columns this epilogue thread handles
  const int row = tid & 127;                               // epilogue threads: accumulator row == TMEM lane == env within the tile
  const uint32_t sbase = smem_u32(smem);
#if QS_WS_TRACE
  const bool trace_on = blockIdx.x == 0 && g == 0 && lane == 0 && !env_warp && (issuer || (tid % kWsEpi) == 0);
#endif
  const uint32_t bar_w = sbase + kWsOffBar, bars = sbase + kWsOffBar + 8 + 64 * g;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kWsOffBar + 8 + kWsGroups * 64);
  const float* tail = reinterpret_cast<const float*>(smem + kOffTail);
  const int tiles = (int)((a.n + kTileRows - 1) / kTileRows);
  const int tile0 = blockIdx.x * kWsGroups + g, tile_stride = gridDim.x * kWsGroups;

  // FUSED: action hand-off slots of this group and their barriers
  const uint32_t act_bar = sbase + kFuOffBar + g * 32;  // + slot * 16: full, + 8: empty
  float4* act_buf = reinterpret_cast<float4*>(smem + kFuOffAct) + g * 2 * kTileRows;

  ObsRow o = {};
  if (!issuer && !env_warp) o = load_obs(a.obs, (int64_t)tile0 * kTileRows + row, tile0 < tiles ? a.n : 0, nullptr);  // in flight during the set-up

  if (tid_hw == 0) {
    mbar_init(bar_w, 1);
    for (int gg = 0; gg < kWsGroups; gg++)
      for (int j = 0; j < 8; j++) mbar_init(sbase + kWsOffBar + 8 + 64 * gg + 8 * j, j < kDone1 ? kWsEpi : 1);
    if (FUSED) {
      for (int j = 0; j < kWsGroups * 4; j++) mbar_init(sbase + kFuOffBar + 8 * j, kTileRows);
      for (int j = 0; j < kWsGroups * 2; j++) mbar_init(sbase + kFuOffBar + kWsGroups * 32 + 8 * j, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_w, kPackedBytes);
    for (int off = 0; off < kPackedBytes; off += 16384) {
      const int bytes = kPackedBytes - off < 16384 ? kPackedBytes - off : 16384;
      bulk_g2s(sbase + kOffW1 + off, a.packed + off, bytes, bar_w);
    }
  }
  if (tid_hw < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot + g * kWsWindow;
  const uint32_t a_base = sbase + kOffA + g * kABytes;
  constexpr uint32_t idesc_n128 = umma_idesc(128, kHid), idesc3 = umma_idesc(128, kN3);
  // two groups: D1 and the vf half at [128,256), D3 right behind side 0's packed A3 pi; three groups: everything at [0,128)
  constexpr uint32_t kColD1 = kWsGroups == 2 ? 128 : 0, kColD2v = kWsGroups == 2 ? 128 : 0;
  // kEarlyL1 (two groups, one side): the NEXT tile's layer 1 is issued BEFORE this tile's layer-3 vf steps, so that it is not
  // queued behind them (the issuer blocks in every tcgen05.mma while the other group's layer 2 holds the pipe; measured
  // with QS_WS_TRACE: the next tile's A1 waited ~850 clk for the issuer).  D1 must then avoid A3 vf [128,192) and D3, which
  // layer 3 vf still uses: it is split into two N=64 halves, [0,64) (A3 pi, dead once layer 3 pi is in the pipe ahead of it)
  // and [192,256) (the read-out part of D2 vf); D3 moves to [112,128), the tail of the consumed D2 pi columns.
  constexpr bool kEarlyL1 = kWsGroups == 2 && kWsSides == 1 && QS_POLICY_WS_EARLY_L1;
  static_assert(!kEarlyL1 || CHUNK <= 64, "D1 is read in two 64-column pieces");
  constexpr uint32_t kColD3ws = kEarlyL1 ? 112 : (kWsGroups == 2 ? kWsCols / 2 : 128);
  auto d1_col = [](int c) -> uint32_t { return kEarlyL1 ? (uint32_t)(c < 64 ? c : 128 + c) : kColD1 + (uint32_t)c; };
  constexpr uint32_t idesc_n64 = umma_idesc(128, 64);

  if (env_warp) {
    // ---------------------------------------------------------------- env warps (FUSED): thread = env, as in dock_step_kernel
    if constexpr (FUSED) {
      constexpr int NW = dock_words<float>(), NR = NW / 4;
      const int er = env_tid % kTileRows;
      // The tile's state rows arrive by bulk copy (cp.async.bulk, one elected thread, completion on an mbarrier) into this
      // set's 20 KB stage; the request for the set's NEXT tile goes out as soon as all 128 threads have moved the current one
      // to registers -- its HBM latency hides under this tile's arithmetic (registers are the second buffer, as in
      // dock_step_pipe_kernel).  With plain loads the env warps, not the policy chain, set the step time: 155 us vs ~90.
      const float4* state_rows = reinterpret_cast<const float4*>(ea.state);
      const uint32_t st_bar = sbase + kFuOffBar + kWsGroups * 32 + (g * 2 + env_set) * 8;
      const float4* st_buf = reinterpret_cast<const float4*>(smem + kFuOffState + (g * kFuSets + env_set) * kFuStateBytes);
      const uint32_t st_dst = sbase + kFuOffState + (g * kFuSets + env_set) * kFuStateBytes;
      auto request = [&](uint32_t it_next) {
        const int64_t env0 = ea.first + ((int64_t)tile0 + (int64_t)it_next * tile_stride) * kTileRows;
        mbar_expect_tx(st_bar, kFuStateBytes);
        for (int off = 0; off < kFuStateBytes; off += 10240)
          bulk_g2s(st_dst + off, reinterpret_cast<const unsigned char*>(state_rows + (env0 / kTileRows) * (NR * kTileRows)) + off, 10240, st_bar);
      };
      if (kFuStage && er == 0 && tile0 + (int64_t)env_set * tile_stride < tiles) request(env_set);
      uint32_t n_it = 0;
      for (uint32_t it = env_set; tile0 + (int64_t)it * tile_stride < tiles; it += kFuSets, n_it++) {  // this set's share of the group's tiles
        const int tile = tile0 + (int)it * tile_stride;
        const int64_t i = ea.first + (int64_t)tile * kTileRows + er;
        const bool active = i < ea.first + ea.n;
        float w[NW];
        if constexpr (kFuStage) {
          mbar_wait(st_bar, n_it & 1);
#pragma unroll
          for (int r = 0; r < NR; r++) unpack(st_buf[r * kTileRows + er], &w[r * 4]);
          asm volatile("bar.sync %0, %1;" ::"r"(1 + g * kFuSets + env_set), "n"(kTileRows) : "memory");  // the stage is drained
          if (er == 0 && tile0 + (int64_t)(it + kFuSets) * tile_stride < tiles) request(it + kFuSets);
        } else {
          if (active) load_words<float, NW>(ea.state, i, w);
        }
        const uint32_t slot = it & 1, par = (it >> 1) & 1;
        mbar_wait(act_bar + slot * 16, par);
        const float4 c = act_buf[slot * kTileRows + er];
        mbar_arrive(act_bar + slot * 16 + 8);  // the slot may be rewritten
        const unsigned mask = __ballot_sync(0xffffffffu, active);
        if (active) {
          const float act[4] = {c.x, c.y, c.z, c.w};
          const StepTail tl = dock_step_loaded<float, KIND, false>(ea, DefaultParams<float, KIND>{}, i, w, act);
          warp_stats_update(ea.warp_stats, i, mask, tl.done, tl.docking, tl.nonfinite, tl.ret, tl.len);
        }
      }
    }
  } else if (issuer) {
    // ---------------------------------------------------------------- MMA issuer: one elected lane, the warp waits together
    mbar_wait(bar_w, 0);
    uint32_t it = 0;
    auto issue_l1 = [&](uint32_t parity) {  // layer 1 of the tile whose A1 completes FullA1's phase `parity`
      mbar_wait(bars + 8 * kFullA1, parity);
      tc_fence_after();
      QS_TR(16);
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < kK1 / 16; ks++) {
          if (kEarlyL1) {
#pragma unroll
            for (int h = 0; h < 2; h++)  // W1 rows 64.. start 8 row-groups = 1024 B into each K chunk of the image
              umma_ss(tmem + d1_col(64 * h), umma_desc(a_base + ks * 4096, 2048, 128), umma_desc(sbase + kOffW1 + h * 1024 + ks * 4096, 2048, 128), idesc_n64, ks);
          } else {
            umma_ss(tmem + kColD1, umma_desc(a_base + ks * 4096, 2048, 128), umma_desc(sbase + kOffW1 + ks * 4096, 2048, 128), idesc_n128, ks);
          }
        }
        umma_commit(bars + 8 * kDone1);
      }
      __syncwarp();
      QS_TR(17);
    };
    if (kEarlyL1 && tile0 < tiles) issue_l1(0);
    for (int tile = tile0; tile < tiles; tile += tile_stride, it++) {
      const uint32_t ph = it & 1;
      if (!kEarlyL1) issue_l1(ph);
      mbar_wait(bars + 8 * kFullA2, ph);
      tc_fence_after();
      QS_TR(18);
      // layer 2 as two N=128 halves of the 256-row W2 image (rows 128.. start 16 row-groups = 2048 B in; LBO stays the image's
      // 4096 B), each with its own commit; layer 3's 8 K-steps of a half follow as soon as that half has been packed
      auto issue_l2 = [&](int half) {
#pragma unroll
        for (int ks = 0; ks < kK2 / 16; ks++)
          umma_ss(tmem + half * kColD2v, umma_desc(a_base + ks * 4096, 2048, 128), umma_desc(sbase + kOffW2 + half * 2048 + ks * 8192, 4096, 128), idesc_n128, ks);
        umma_commit(bars + 8 * (half == 0 ? kDoneP : kDoneV));
      };
      if (lane == 0) {
        issue_l2(0);
        if (kWsGroups == 2) issue_l2(1);  // the vf half has its own columns: both halves go into the pipe back to back
      }
      __syncwarp();
      QS_TR(19);
#pragma unroll
      for (int half = 0; half < 2; half++) {
        if (kEarlyL1 && half == 1 && tile + tile_stride < tiles) issue_l1(ph ^ 1);  // the next tile's layer 1 goes ahead of our vf steps
        mbar_wait(bars + 8 * (half == 0 ? kFullA3p : kFullA3v), ph);
        tc_fence_after();
        QS_TR(20 + 2 * half);
        if (lane == 0) {
  #pragma unroll
          for (int k8 = 0; k8 < QS_EXP_L3_STEPS; k8++) {  // K-step k8 of this half: features 16 k8 .., packed by side (16 k8) / kWsCols
            const int ks = 8 * half + k8, sd = (16 * k8) / kWsCols, kk = (16 * k8) % kWsCols;
#if QS_EXP_D3_SPLIT
            // EXPERIMENT: even / odd K-steps accumulate into two D3 tiles ([112,128) and [96,112)), summed by the epilogue
            umma_ts(tmem + kColD3ws - 16 * (ks & 1), tmem + half * kColD2v + sd * kWsCols + kk / 2, umma_desc(sbase + kOffW3 + ks * 512, 256, 128), idesc3, ks >> 1);
#else
            umma_ts(tmem + kColD3ws, tmem + half * kColD2v + sd * kWsCols + kk / 2, umma_desc(sbase + kOffW3 + ks * 512, 256, 128), idesc3, ks);
#endif
          }
          if (half == 1) umma_commit(bars + 8 * kDone3);
          // three groups: the vf half reuses the pi half's columns -- in order behind the layer-3 steps that read A3 pi
          if (kWsGroups == 3 && half == 0) issue_l2(1);
          }
        __syncwarp();
        QS_TR(21 + 2 * half);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue threads
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    unsigned char* a_row = smem + kOffA + g * kABytes + row * 16;
    auto build_a1 = [&](const ObsRow& ob) {
      const float x[12] = {ob.a.x, ob.a.y, ob.a.z, ob.a.w, ob.b.x, ob.b.y, ob.b.z, ob.b.w, ob.c.x, ob.c.y, ob.c.z, ob.c.w};
      uint32_t hi[6], lo[6];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const float h0 = bf16_round(x[2 * k]), h1 = bf16_round(x[2 * k + 1]);
        hi[k] = pack_bf16(h0, h1);
        lo[k] = pack_bf16(x[2 * k] - h0, x[2 * k + 1] - h1);
      }
      if (kWsSides == 1 || side == 0) {
        *reinterpret_cast<uint4*>(a_row + 0 * 2048) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(a_row + 1 * 2048) = make_uint4(hi[4], hi[5], 0x3F803F80u, 0u);  // k = 12, 13: 1.0 against (b1_hi, b1_lo)
      }
      if (kWsSides == 1 || side == 1) {
        *reinterpret_cast<uint4*>(a_row + 2 * 2048) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(a_row + 3 * 2048) = make_uint4(lo[4], lo[5], 0u, 0u);
      }
      proxy_fence();
      mbar_arrive(bars + 8 * kFullA1);
    };
    mbar_wait(bar_w, 0);  // the fp32 tail (biases of the heads, logstd) is read from shared memory below
    if (QS_POLICY_WS_SKEW > 0 && g > 0 && tile0 + tile_stride < tiles) {  // start the groups out of phase (see QS_POLICY_WS_SKEW)
      const long long t_go = clock64() + (long long)QS_POLICY_WS_SKEW * g;
      while (clock64() < t_go) { }
    }
    if (tile0 < tiles) build_a1(o);
    // The tail of a tile (layer-3 read-out, sampling, stores) is deferred into the next iteration: D3 is read right before
    // the next tile's layer-2 operand is handed over (its pi half overwrites D3's columns), and the arithmetic and the
    // stores then run while that layer 2 is in the tensor pipe.
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;  // this tile's noise, carried to the deferred tail
    uint32_t d3[8];
    int64_t i_prev = -1;
    auto read_d3 = [&](uint32_t ph_prev) {
      mbar_wait(bars + 8 * kDone3, ph_prev);
      tc_fence_after();
      tmem_ld8(trow + kColD3ws, d3);
#if QS_EXP_D3_SPLIT
      uint32_t d3b[8];
      tmem_ld8(trow + kColD3ws - 16, d3b);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; j++) d3[j] = __float_as_uint(__uint_as_float(d3[j]) + __uint_as_float(d3b[j]));
#endif
      tmem_wait_ld();
      tc_fence_before();  // D3 is read before the next tile's pi half (ordered by our full_a2 arrival) overwrites it
    };
    auto finish = [&](int64_t i, uint32_t it_of_i, float z0, float z1, float z2, float z3) {
      const bool live = i < a.n;
      const float ls0 = tail[8], ls1 = tail[9], ls2 = tail[10], ls3 = tail[11];
      const float mu0 = __uint_as_float(d3[0]) + tail[0], mu1 = __uint_as_float(d3[1]) + tail[1];
      const float mu2 = __uint_as_float(d3[2]) + tail[2], mu3 = __uint_as_float(d3[3]) + tail[3];
      const float val = __uint_as_float(d3[4]) + tail[4];
      const float a0 = fmaf(__expf(ls0), z0, mu0), a1 = fmaf(__expf(ls1), z1, mu1), a2 = fmaf(__expf(ls2), z2, mu2), a3 = fmaf(__expf(ls3), z3, mu3);
      const float4 clipped = make_float4(fminf(fmaxf(a0, -1.f), 1.f), fminf(fmaxf(a1, -1.f), 1.f), fminf(fmaxf(a2, -1.f), 1.f), fminf(fmaxf(a3, -1.f), 1.f));
      if (FUSED && (kWsSides == 1 || side == 0)) {  // hand the tile's clipped actions to the env warps (every row arrives: the barrier counts 128)
        const uint32_t slot = it_of_i & 1, par = ((it_of_i >> 1) & 1) ^ 1;  // a fresh barrier passes a wait on parity 1: both slots start out empty
        mbar_wait(act_bar + slot * 16 + 8, par);
        act_buf[slot * kTileRows + row] = clipped;
        mbar_arrive(act_bar + slot * 16);
      }
      if (!live) return;
      if (kWsSides == 1 || side == 0) {
        // neglogp = 0.5 * sum(((a - mean)/std)^2) + 0.5*log(2 pi)*4 + sum(logstd); (a - mean)/std == the noise itself
        const float nlp = 0.5f * (z0 * z0 + z1 * z1 + z2 * z2 + z3 * z3) + 3.6757541328186907f + (ls0 + ls1 + ls2 + ls3);
        if (a.actions) reinterpret_cast<float4*>(a.actions)[i] = make_float4(a0, a1, a2, a3);
        if (a.clipped) reinterpret_cast<float4*>(a.clipped)[i] = clipped;
        if (a.neglogp) a.neglogp[i] = nlp;
      }
      if (kWsSides == 1 || side == 1) a.values[i] = val;
    };
    uint32_t it_prev = 0;
    uint32_t it = 0;
    for (int tile = tile0; tile < tiles; tile += tile_stride, it++) {
      const uint32_t ph = it & 1;
      const int64_t i = (int64_t)tile * kTileRows + row;
      const bool more = tile + tile_stride < tiles;
      QS_TR(0);
      o = load_obs(a.obs, i + (int64_t)tile_stride * kTileRows, a.n, nullptr);  // next tile's row, in flight under this tile's work
      // ---- h = relu(D1) -> A2[:, 0:128]; A2[:, 128:144] = (1, 1, 0, ...) carries layer 2's bias ----
      mbar_wait(bars + 8 * kDone1, ph);
      tc_fence_after();
      QS_TR(1);
#pragma unroll 1
      for (int c0 = kWsCols * side; c0 < kWsCols * (side + 1); c0 += CHUNK) {
        uint32_t r[CHUNK];
#pragma unroll
        for (int q = 0; q < CHUNK; q += 32) tmem_ld32(trow + d1_col(c0) + q, r + q);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < CHUNK / 8; j++) {
          uint4 w;
          w.x = relu_pack_bf16(r[8 * j + 0], r[8 * j + 1]);
          w.y = relu_pack_bf16(r[8 * j + 2], r[8 * j + 3]);
          w.z = relu_pack_bf16(r[8 * j + 4], r[8 * j + 5]);
          w.w = relu_pack_bf16(r[8 * j + 6], r[8 * j + 7]);
          *reinterpret_cast<uint4*>(a_row + ((c0 >> 3) + j) * 2048) = w;
        }
      }
      if (kWsSides == 1 || side == 0) *reinterpret_cast<uint4*>(a_row + 16 * 2048) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
      if (kWsSides == 1 || side == 1) *reinterpret_cast<uint4*>(a_row + 17 * 2048) = make_uint4(0u, 0u, 0u, 0u);
      QS_TR(2);
      proxy_fence();
      QS_TR(3);
      const float p0 = n0, p1 = n1, p2 = n2, p3 = n3;
      if (i_prev >= 0) read_d3(ph ^ 1);  // the previous tile's layer-3 result, before this tile's pi half lands on it
      QS_TR(4);
      tc_fence_before();  // D1 is read; the vf half of layer 2 will overwrite its columns
      mbar_arrive(bars + 8 * kFullA2);
      QS_TR(5);
      if (i_prev >= 0) finish(i_prev, it_prev, p0, p1, p2, p3);  // under this tile's layer 2
      QS_TR(6);
      // ---- relu(D2 half) -> bf16 pairs packed in place: A3 pi = [0,64), A3 vf = [128,192) ----
#pragma unroll 1
      for (int half = 0; half < 2; half++) {
        mbar_wait(bars + 8 * (half == 0 ? kDoneP : kDoneV), ph);
        tc_fence_after();
        QS_TR(7 + 3 * half);
        const uint32_t base = trow + half * kColD2v + kWsCols * side;  // this side's columns; its packed A3 lands at their start
#pragma unroll 1
        for (int c0 = 0; c0 < kWsCols; c0 += CHUNK) {
          uint32_t r[CHUNK], w[CHUNK / 2];
#pragma unroll
          for (int q = 0; q < CHUNK; q += 32) tmem_ld32(base + c0 + q, r + q);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < CHUNK / 2; j++) w[j] = relu_pack_bf16(r[2 * j], r[2 * j + 1]);
#pragma unroll
          for (int q = 0; q < CHUNK / 2; q += 16) tmem_st16(base + (c0 >> 1) + q, w + q);
        }
        QS_TR(8 + 3 * half);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bars + 8 * (half == 0 ? kFullA3p : kFullA3v));
        QS_TR(9 + 3 * half);
      }
      // ---- the next tile's A1 (the A buffer is free: both layer-2 halves have completed) and its layer 1, under our tail ----
      if (more) build_a1(o);
      QS_TR(13);
      // ---- Gaussian noise for this tile while its layer 3 and the next tile's layer 1 run (DiagGaussianProbability
      //      Distribution.sample) ----
      n0 = n1 = n2 = n3 = 0.f;
      if ((kWsSides == 1 || side == 0) && !a.deterministic) {
        qs::U4 c = {(uint32_t)(a.env_id0 + (uint64_t)i), (uint32_t)((a.env_id0 + (uint64_t)i) >> 32), a.step_idx, 2u};
        qs::U4 x = qs::philox4x32_10(c, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        const float s = 5.9604644775390625e-08f;  // 2^-24; u in (0,1)
        const float u0 = ((float)(x.x >> 8) + 0.5f) * s, u1 = ((float)(x.y >> 8) + 0.5f) * s;
        const float u2 = ((float)(x.z >> 8) + 0.5f) * s, u3 = ((float)(x.w >> 8) + 0.5f) * s;
        const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
        float s0, c0_, s1, c1_;
        __sincosf(6.283185307179586f * u1, &s0, &c0_);
        __sincosf(6.283185307179586f * u3, &s1, &c1_);
        n0 = r0 * c0_; n1 = r0 * s0; n2 = r1 * c1_; n3 = r1 * s1;  // Box-Muller
      }
      QS_TR(14);
      i_prev = i;
      it_prev = it;
    }
    if (i_prev >= 0) {  // the last tile's tail
      read_d3((it - 1) & 1);
      finish(i_prev, it_prev, n0, n1, n2, n3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid_hw < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(kTmemCols) : "memory");
}
__global__ void __launch_bounds__(kWsThreads, 1) policy_fwd_ws_kernel(const StepArgs a) { policy_fwd_ws_body<false, kDockV2>(a, EnvStepArgs{}); }
// The same kernel reading its accumulators out in chunks of 32 columns: 96 registers instead of 128, 73 instead of 68 us per
// 2^20 envs on its own -- but in a sliced rollout three env-step CTAs fit next to it on an SM instead of two, and the pair
// overlaps better: 2.06 vs 2.12 ms per 16-step rollout.  qs_rollout launches this one when it runs more than one slice.
__global__ void __launch_bounds__(kWsThreads, 1) policy_fwd_ws_kernel_lean(const StepArgs a) { policy_fwd_ws_body<false, kDockV2, 32>(a, EnvStepArgs{}); }
// model.step + env.step of one rollout step in ONE launch (see the kFu* constants above): rl_baselines/ppo2/ppo2.py:472-482
template <int KIND>
__global__ void __launch_bounds__(kFuThreads, 1) policy_env_step_kernel(const StepArgs a, const __grid_constant__ EnvStepArgs ea) {
  policy_fwd_ws_body<true, KIND, QS_FUSED_CHUNK>(a, ea);
}
// Register-capped builds of the same kernel for the sliced rollout: a policy CTA at 128 registers x 320 threads leaves an SM
// room for ONE 128-thread env CTA (80 registers) next to it, at 80 registers for three -- and it is the env kernel's occupancy
// that decides how much of its HBM time hides under the other slice's policy kernel.  The cap costs ~100 bytes of spills.
template <int REGS> __global__ void __maxnreg__(REGS) policy_fwd_ws_kernel_capped(const StepArgs a) { policy_fwd_ws_body<false, kDockV2>(a, EnvStepArgs{}); }

// =========================================================================================================
// The accurate forward (qs_policy_set_precision(QS_POLICY_FP32)): the same model.step in plain fp32 on the CUDA cores --
// fp32 weights, fp32 activations, fp32 accumulation, i.e. what the reference's TensorFlow policy computes
// (run_moving_docking_ppo2.py:88-91), to within summation order.  For evaluation runs and for checking the bf16 tensor-core
// path; ~25x slower than it: 1.76 ms per 2^20 envs = 58 % of the fp32 FMA peak (36.7 G FMA per 2^20 envs on the FMA pipe instead of the tensor pipe).
// One 256-thread CTA per SM walks 64-env tiles; all weights (140 KB) stay in shared memory.  Layer 2 is register-tiled:
// warp w owns rows 8w..8w+7, lane l the pi columns 4l..4l+3 and the vf columns 128+4l..: 64 accumulators per thread, per k
// two broadcast float4 of activations and two conflict-free float4 of weights feed 64 FMAs.  Layer 3 is folded into the
// same threads (each multiplies its 8 relu'd columns into 4 + 1 partial sums per row) and reduced over the warp.
// Sampling, clipping and neglogp follow policy_fwd_ws_kernel instruction for instruction: the same (seed, env, step) draws
// the same noise in either precision.
// =========================================================================================================
constexpr int kF32Rows = 64, kF32Threads = 256;
constexpr int kF32OffObs = kF32Floats;                          // obsT [12][64]
constexpr int kF32OffH = kF32OffObs + kObs * kF32Rows;          // hT [128][64]
constexpr int kF32OffOut = kF32OffH + kHid * kF32Rows;          // out [64][8]: mean(4), value
constexpr int kF32SmemBytes = (kF32OffOut + kF32Rows * 8) * 4;
static_assert(kF32SmemBytes <= 227 * 1024 && kF32OffObs % 4 == 0 && kF32OffH % 4 == 0, "shared memory budget / alignment");

__global__ void __launch_bounds__(kF32Threads, 1) policy_fwd_f32_kernel(const StepArgs a, const float* __restrict__ img) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kF32Floats / 4; i += kF32Threads) reinterpret_cast<float4*>(sm)[i] = reinterpret_cast<const float4*>(img)[i];
  const float *W1T = sm + kF32W1T, *B1 = sm + kF32B1, *W2T = sm + kF32W2T, *B2 = sm + kF32B2, *W3Q = sm + kF32W3Q, *tail = sm + kF32Tail;
  float *obsT = sm + kF32OffObs, *hT = sm + kF32OffH, *out = sm + kF32OffOut;
  const int64_t tiles = (a.n + kF32Rows - 1) / kF32Rows;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t i0 = tile * kF32Rows;
    __syncthreads();  // the previous tile's readers are done (and the weights are in place)
    for (int e = tid; e < kF32Rows * kObs; e += kF32Threads) {  // coalesced read of the tile's [64][12] block, transposed into obsT
      const int r = e / kObs, k = e % kObs;
      obsT[k * kF32Rows + r] = i0 + r < a.n ? a.obs[(i0 + r) * kObs + k] : 0.f;
    }
    __syncthreads();
    {  // layer 1: thread = (row r, 32 output features)
      const int r = tid % kF32Rows, n0 = (tid / kF32Rows) * 32;
      float x[kObs];
#pragma unroll
      for (int k = 0; k < kObs; k++) x[k] = obsT[k * kF32Rows + r];
#pragma unroll 4
      for (int n = n0; n < n0 + 32; n += 4) {
        float4 acc = *reinterpret_cast<const float4*>(B1 + n);
#pragma unroll
        for (int k = 0; k < kObs; k++) {
          const float4 w = *reinterpret_cast<const float4*>(W1T + k * kHid + n);
          acc.x = fmaf(x[k], w.x, acc.x); acc.y = fmaf(x[k], w.y, acc.y); acc.z = fmaf(x[k], w.z, acc.z); acc.w = fmaf(x[k], w.w, acc.w);
        }
        hT[(n + 0) * kF32Rows + r] = fmaxf(acc.x, 0.f); hT[(n + 1) * kF32Rows + r] = fmaxf(acc.y, 0.f);
        hT[(n + 2) * kF32Rows + r] = fmaxf(acc.z, 0.f); hT[(n + 3) * kF32Rows + r] = fmaxf(acc.w, 0.f);
      }
    }
    __syncthreads();
    {  // layers 2 and 3
      const int r0 = warp * 8, cp = lane * 4, cv = kHid + lane * 4;
      float acc[8][8];
      {
        const float4 bp = *reinterpret_cast<const float4*>(B2 + cp), bv = *reinterpret_cast<const float4*>(B2 + cv);
#pragma unroll
        for (int r = 0; r < 8; r++) { acc[r][0] = bp.x; acc[r][1] = bp.y; acc[r][2] = bp.z; acc[r][3] = bp.w; acc[r][4] = bv.x; acc[r][5] = bv.y; acc[r][6] = bv.z; acc[r][7] = bv.w; }
      }
#pragma unroll 2
      for (int k = 0; k < kHid; k++) {
        const float4 ha = *reinterpret_cast<const float4*>(hT + k * kF32Rows + r0), hb = *reinterpret_cast<const float4*>(hT + k * kF32Rows + r0 + 4);
        const float4 wp = *reinterpret_cast<const float4*>(W2T + k * 2 * kHid + cp), wv = *reinterpret_cast<const float4*>(W2T + k * 2 * kHid + cv);
        const float h[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
        const float w[8] = {wp.x, wp.y, wp.z, wp.w, wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int c = 0; c < 8; c++) acc[r][c] = fmaf(h[r], w[c], acc[r][c]);
      }
      float4 q[4];
      float wvv[4];
#pragma unroll
      for (int c = 0; c < 4; c++) { q[c] = *reinterpret_cast<const float4*>(W3Q + (cp + c) * 4); wvv[c] = W3Q[(cv + c) * 4]; }
#pragma unroll
      for (int r = 0; r < 8; r++) {
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, v = 0.f;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const float hp = fmaxf(acc[r][c], 0.f), hv = fmaxf(acc[r][4 + c], 0.f);
          m0 = fmaf(hp, q[c].x, m0); m1 = fmaf(hp, q[c].y, m1); m2 = fmaf(hp, q[c].z, m2); m3 = fmaf(hp, q[c].w, m3);
          v = fmaf(hv, wvv[c], v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          m0 += __shfl_xor_sync(0xffffffffu, m0, o); m1 += __shfl_xor_sync(0xffffffffu, m1, o); m2 += __shfl_xor_sync(0xffffffffu, m2, o);
          m3 += __shfl_xor_sync(0xffffffffu, m3, o); v += __shfl_xor_sync(0xffffffffu, v, o);
        }
        if (lane == 0) {
          float* o8 = out + (r0 + r) * 8;
          o8[0] = m0; o8[1] = m1; o8[2] = m2; o8[3] = m3; o8[4] = v;
        }
      }
    }
    __syncthreads();
    if (tid < kF32Rows && i0 + tid < a.n) {  // sampling and stores: one thread per env, as the epilogue of policy_fwd_ws_kernel
      const int64_t i = i0 + tid;
      const float* o8 = out + tid * 8;
      float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
      if (!a.deterministic) {
        qs::U4 c = {(uint32_t)(a.env_id0 + (uint64_t)i), (uint32_t)((a.env_id0 + (uint64_t)i) >> 32), a.step_idx, 2u};
        qs::U4 x = qs::philox4x32_10(c, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        const float s = 5.9604644775390625e-08f;  // 2^-24; u in (0,1)
        const float u0 = ((float)(x.x >> 8) + 0.5f) * s, u1 = ((float)(x.y >> 8) + 0.5f) * s;
        const float u2 = ((float)(x.z >> 8) + 0.5f) * s, u3 = ((float)(x.w >> 8) + 0.5f) * s;
        const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
        float s0, c0_, s1, c1_;
        __sincosf(6.283185307179586f * u1, &s0, &c0_);
        __sincosf(6.283185307179586f * u3, &s1, &c1_);
        n0 = r0 * c0_; n1 = r0 * s0; n2 = r1 * c1_; n3 = r1 * s1;  // Box-Muller
      }
      const float ls0 = tail[8], ls1 = tail[9], ls2 = tail[10], ls3 = tail[11];
      const float mu0 = o8[0] + tail[0], mu1 = o8[1] + tail[1], mu2 = o8[2] + tail[2], mu3 = o8[3] + tail[3], val = o8[4] + tail[4];
      const float a0 = fmaf(__expf(ls0), n0, mu0), a1 = fmaf(__expf(ls1), n1, mu1), a2 = fmaf(__expf(ls2), n2, mu2), a3 = fmaf(__expf(ls3), n3, mu3);
      const float nlp = 0.5f * (n0 * n0 + n1 * n1 + n2 * n2 + n3 * n3) + 3.6757541328186907f + (ls0 + ls1 + ls2 + ls3);
      if (a.actions) reinterpret_cast<float4*>(a.actions)[i] = make_float4(a0, a1, a2, a3);
      if (a.clipped) reinterpret_cast<float4*>(a.clipped)[i] = make_float4(fminf(fmaxf(a0, -1.f), 1.f), fminf(fmaxf(a1, -1.f), 1.f), fminf(fmaxf(a2, -1.f), 1.f), fminf(fmaxf(a3, -1.f), 1.f));
      if (a.neglogp) a.neglogp[i] = nlp;
      a.values[i] = val;
    }
  }
}

}  // namespace

struct NvtxRange {  // see qs_lib.cu
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

constexpr int kMaxSlices = 8;
constexpr int kDefaultPolicyRegs = 128;  // per-thread register build of the rollout forward kernel (see policy_fwd_ws_kernel_capped)
struct qs_policy {
  int device = 0;
  unsigned char* packed = nullptr;
  float* img32 = nullptr;       // fp32 image of the weights (policy_fwd_f32_kernel)
  double* adv_ws = nullptr;     // adv_stats_kernel: [2 * kAdvBlocks] partial sums | ticket (4 B) | pad | stats (2 floats)
  int precision = QS_POLICY_BF16;
  int sms = 0;
  int grid = 0;
  bool have_weights = false;
  // qs_rollout with n_slices > 1: one stream per env slice, forked from / joined into the caller's stream
  cudaStream_t slice_stream[kMaxSlices] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxSlices] = {};
};

#define QSP_CUDA(call)                                                                                          \
  do {                                                                                                          \
    cudaError_t e_ = (call);                                                                                    \
    if (e_ != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#if QS_WS_TRACE
extern "C" int qs_debug_ws_trace(long long* out, int max_words) {
  const int n = kTrTiles * kTrSlots < max_words ? kTrTiles * kTrSlots : max_words;
  return cudaMemcpyFromSymbol(out, g_ws_trace, n * sizeof(long long)) == cudaSuccess ? n : -1;
}
#endif
extern "C" {

int qs_policy_create(qs_policy_t** out, int device) {
  if (!out) return qs_fail_(QS_ERR_INVALID, "qs_policy_create: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return qs_fail_(QS_ERR_NO_DEVICE, "qs_policy_create: no CUDA device; there is no CPU fallback");
  if (device < 0 || device >= ndev) return qs_fail_(QS_ERR_INVALID, "qs_policy_create: device %d out of range", device);
  int prev = 0;
  QSP_CUDA(cudaGetDevice(&prev));
  QSP_CUDA(cudaSetDevice(device));
  qs_policy* p = new (std::nothrow) qs_policy();
  if (!p) { cudaSetDevice(prev); return qs_fail_(QS_ERR_ALLOC, "qs_policy_create: out of host memory"); }
  p->device = device;
  cudaDeviceProp prop;
  int per_sm = 0;
  cudaError_t err;
  if ((err = cudaGetDeviceProperties(&prop, device)) != cudaSuccess || (err = cudaMalloc(&p->packed, kPackedBytes)) != cudaSuccess ||
      (err = cudaMalloc((void**)&p->img32, kF32Floats * sizeof(float))) != cudaSuccess ||
      (err = cudaMalloc((void**)&p->adv_ws, (2 * kAdvBlocks + 2) * sizeof(double))) != cudaSuccess ||
      (err = cudaMemset(p->adv_ws, 0, (2 * kAdvBlocks + 2) * sizeof(double))) != cudaSuccess ||
      (err = cudaFuncSetAttribute(policy_fwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF32SmemBytes)) != cudaSuccess ||
      (err = cudaFuncSetAttribute(policy_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)) != cudaSuccess ||
      (err = cudaFuncSetAttribute(policy_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)) != cudaSuccess ||
      (err = cudaFuncSetAttribute(dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemBytes)) != cudaSuccess ||
      (kFuFits && (err = cudaFuncSetAttribute(policy_env_step_kernel<kDockV0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemBytes)) != cudaSuccess) ||
      (kFuFits && (err = cudaFuncSetAttribute(policy_env_step_kernel<kDockV2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemBytes)) != cudaSuccess) ||
      (err = cudaFuncSetAttribute(policy_fwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemBytes)) != cudaSuccess ||
      (err = cudaFuncSetAttribute(policy_fwd_ws_kernel_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemBytes)) != cudaSuccess ||
      (err = cudaFuncSetAttribute(policy_fwd_ws_kernel_capped<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemBytes)) != cudaSuccess ||
      (err = cudaFuncSetAttribute(policy_fwd_ws_kernel_capped<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemBytes)) != cudaSuccess ||
      (err = cudaFuncSetAttribute(policy_fwd_ws_kernel_capped<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmemBytes)) != cudaSuccess ||
      (err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, policy_kernel<true>, kThreads, kSmemBytes)) != cudaSuccess) {
    int rc = qs_fail_(QS_ERR_CUDA, "qs_policy_create: %s", cudaGetErrorString(err));
    cudaFree(p->packed);
    cudaFree(p->img32);
    cudaFree(p->adv_ws);
    delete p;
    cudaSetDevice(prev);
    return rc;
  }
  if (per_sm < 1) {
    int rc = qs_fail_(QS_ERR_CUDA, "qs_policy_create: the policy kernel does not fit an SM (%d B shared memory)", kSmemBytes);
    cudaFree(p->packed);
    cudaFree(p->img32);
    cudaFree(p->adv_ws);
    delete p;
    cudaSetDevice(prev);
    return rc;
  }
  p->grid = p->sms = prop.multiProcessorCount;  // one persistent CTA per SM: it owns all 512 TMEM columns
  // QUADSIM_B200_POLICY_CTAS=<n> runs the forward kernel on fewer SMs (A/B: a sliced rollout leaves the rest to the env kernel)
  if (const char* v = getenv("QUADSIM_B200_POLICY_CTAS")) { const int c = atoi(v); if (c > 0 && c < p->grid) p->grid = c; }
  cudaSetDevice(prev);
  *out = p;
  return QS_OK;
}

int qs_policy_destroy(qs_policy_t* p) {
  if (!p) return QS_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);
  cudaFree(p->packed);
  cudaFree(p->img32);
  cudaFree(p->adv_ws);
  for (int k = 0; k < kMaxSlices; k++) {
    if (p->slice_stream[k]) cudaStreamDestroy(p->slice_stream[k]);
    if (p->ev_join[k]) cudaEventDestroy(p->ev_join[k]);
  }
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  cudaSetDevice(prev);
  delete p;
  return QS_OK;
}

int qs_policy_set_precision(qs_policy_t* p, int precision) {
  if (!p || (precision != QS_POLICY_BF16 && precision != QS_POLICY_FP32)) return qs_fail_(QS_ERR_INVALID, "qs_policy_set_precision: expected QS_POLICY_BF16 or QS_POLICY_FP32");
  p->precision = precision;
  return QS_OK;
}

int qs_policy_set_weights(qs_policy_t* p, const qs_policy_weights_t* w, void* stream) {
  if (!p || !w) return qs_fail_(QS_ERR_INVALID, "qs_policy_set_weights: null argument");
  const float* ptrs[11] = {w->shared_w, w->shared_b, w->pi_fc_w, w->pi_fc_b, w->vf_fc_w, w->vf_fc_b, w->pi_w, w->pi_b, w->vf_w, w->vf_b, w->logstd};
  for (const float* q : ptrs)
    if (!q) return qs_fail_(QS_ERR_INVALID, "qs_policy_set_weights: null weight pointer");
  int prev = 0;
  QSP_CUDA(cudaGetDevice(&prev));
  QSP_CUDA(cudaSetDevice(p->device));
  RawWeights r = {w->shared_w, w->shared_b, w->pi_fc_w, w->pi_fc_b, w->vf_fc_w, w->vf_fc_b, w->pi_w, w->pi_b, w->vf_w, w->vf_b, w->logstd};
  pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(r, p->packed, p->img32);
  cudaError_t e = cudaGetLastError();
  cudaSetDevice(prev);
  if (e != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "qs_policy_set_weights: %s", cudaGetErrorString(e));
  p->have_weights = true;
  return QS_OK;
}

}  // extern "C"

// lean: the 96-register build of the forward kernel (policy_fwd_ws_kernel_lean), for launches that share the SMs with env-step kernels
static int policy_step_impl(qs_policy_t* p, int64_t n, const float* obs, uint64_t seed, uint64_t env_id0, uint32_t step_idx, int deterministic,
                            float* actions_out, float* clipped_out, float* values_out, float* neglogp_out, void* stream, bool lean) {
  if (!p || !obs || !values_out) return qs_fail_(QS_ERR_INVALID, "qs_policy_step: null argument");
  if (!p->have_weights) return qs_fail_(QS_ERR_INVALID, "qs_policy_step: qs_policy_set_weights has not been called");
  if (n <= 0) return QS_OK;
  NvtxRange nv("qs_policy_step");
  int prev = 0;
  QSP_CUDA(cudaGetDevice(&prev));
  QSP_CUDA(cudaSetDevice(p->device));
  StepArgs a = {p->packed, obs, n, seed, env_id0, step_idx, deterministic, actions_out, clipped_out, values_out, neglogp_out};
  const int ctas = (int)(((n + kTileRows - 1) / kTileRows + kGroups - 1) / kGroups);
  // QUADSIM_B200_POLICY=v3 selects the previous (leader-thread-in-group) forward kernel for A/B
  static const bool use_v3 = [] { const char* v = getenv("QUADSIM_B200_POLICY"); return v && strcmp(v, "v3") == 0; }();
  if (p->precision == QS_POLICY_FP32) {
    const int64_t t64 = (n + kF32Rows - 1) / kF32Rows;
    policy_fwd_f32_kernel<<<(unsigned)(t64 < p->sms ? t64 : p->sms), kF32Threads, kF32SmemBytes, (cudaStream_t)stream>>>(a, p->img32);
  } else if (use_v3) policy_kernel<false><<<ctas < p->grid ? ctas : p->grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(a, TrainArgs{});
  else {
    static const bool pdl = [] { const char* v = getenv("QUADSIM_B200_PDL"); return !(v && v[0] == '0'); }();
    // Launch priority: in a sliced rollout the policy kernel of one slice and the env kernel of the other become runnable at
    // the same moment.  The env kernel's thousands of small CTAs would otherwise take the registers an SM frees before the one
    // big policy CTA (164 KB of shared memory, all of TMEM) can be placed, and the two kernels end up running one after the
    // other; with the policy kernel at the highest priority its CTA goes first and the env CTAs fill what is left.
    // QUADSIM_B200_POLICY_PRIO=0 launches at the default priority (A/B).
    static const int prio = [] {
      const char* v = getenv("QUADSIM_B200_POLICY_PRIO");
      int least = 0, greatest = 0;
      if ((v && v[0] == '0') || cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return 0;
      return greatest;
    }();
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      na++;
    }
    if (prio != 0) {
      attr[na].id = cudaLaunchAttributePriority;
      attr[na].val.priority = prio;
      na++;
    }
    cudaLaunchConfig_t cfg = {};
    const int ws_ctas = (int)(((n + kTileRows - 1) / kTileRows + kWsGroups - 1) / kWsGroups);
    cfg.gridDim = dim3(ws_ctas < p->grid ? ws_ctas : p->grid); cfg.blockDim = dim3(kWsThreads); cfg.dynamicSmemBytes = kWsSmemBytes;
    cfg.stream = (cudaStream_t)stream; cfg.attrs = attr; cfg.numAttrs = na;
    // QUADSIM_B200_POLICY_REGS = 128 (uncapped build) | 96 | 80 | 72: see policy_fwd_ws_kernel_capped
    static const int regs = [] { const char* v = getenv("QUADSIM_B200_POLICY_REGS"); return v ? atoi(v) : kDefaultPolicyRegs; }();
    static const int lean_env = [] { const char* v = getenv("QUADSIM_B200_POLICY_LEAN"); return v ? (v[0] == '0' ? 0 : 1) : -1; }();  // A/B: force either build
    const bool use_lean = lean_env < 0 ? lean : lean_env == 1;
    cudaError_t le = regs == 96   ? cudaLaunchKernelEx(&cfg, policy_fwd_ws_kernel_capped<96>, a)
                     : regs == 80 ? cudaLaunchKernelEx(&cfg, policy_fwd_ws_kernel_capped<80>, a)
                     : regs == 72 ? cudaLaunchKernelEx(&cfg, policy_fwd_ws_kernel_capped<72>, a)
                     : use_lean   ? cudaLaunchKernelEx(&cfg, policy_fwd_ws_kernel_lean, a)
                                  : cudaLaunchKernelEx(&cfg, policy_fwd_ws_kernel, a);
    if (le != cudaSuccess) { cudaSetDevice(prev); return qs_fail_(QS_ERR_CUDA, "qs_policy_step: %s", cudaGetErrorString(le)); }
  }
  cudaError_t e = cudaGetLastError();
  cudaSetDevice(prev);
  if (e != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "qs_policy_step: %s", cudaGetErrorString(e));
  return QS_OK;
}

extern "C" {

int qs_policy_step(qs_policy_t* p, int64_t n, const float* obs, uint64_t seed, uint64_t env_id0, uint32_t step_idx, int deterministic,
                   float* actions_out, float* clipped_out, float* values_out, float* neglogp_out, void* stream) {
  return policy_step_impl(p, n, obs, seed, env_id0, step_idx, deterministic, actions_out, clipped_out, values_out, neglogp_out, stream, false);
}

int qs_policy_ppo_grad(qs_policy_t* p, const qs_ppo_grad_t* g, void* stream) {
  if (!p || !g) return qs_fail_(QS_ERR_INVALID, "qs_policy_ppo_grad: null argument");
  if (!p->have_weights) return qs_fail_(QS_ERR_INVALID, "qs_policy_ppo_grad: qs_policy_set_weights has not been called");
  if (!g->obs || !g->actions || !g->returns || !g->values || !g->neglogp || !g->a1 || !g->h || !g->hpv || !g->dz2 || !g->dz1 ||
      !g->dout || !g->col_sums || !g->g1 || !g->g2 || !g->g3)
    return qs_fail_(QS_ERR_INVALID, "qs_policy_ppo_grad: null buffer");
  if (g->batch <= 0) return qs_fail_(QS_ERR_INVALID, "qs_policy_ppo_grad: batch must be positive");
  NvtxRange nv("qs_policy_ppo_grad");
  int prev = 0;
  QSP_CUDA(cudaGetDevice(&prev));
  QSP_CUDA(cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = (int)((g->batch + kTileRows - 1) / kTileRows);
  cudaError_t e = cudaSuccess;
  constexpr size_t n1 = (size_t)kHid * kK1, n2 = (size_t)2 * kHid * kK2, n3 = (size_t)2 * kHid * 16;
  if (g->g2 == g->g1 + n1 && g->g3 == g->g2 + n2 && g->col_sums == g->g3 + n3) {  // one allocation (the caller's all-reduce bucket): one memset
    e = cudaMemsetAsync(g->g1, 0, (n1 + n2 + n3 + 16) * sizeof(float), st);
  } else {
    e = cudaMemsetAsync(g->col_sums, 0, 16 * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->g1, 0, n1 * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->g2, 0, n2 * sizeof(float), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->g3, 0, n3 * sizeof(float), st);
  }
  const float* adv_stats = g->adv_stats;
  if (e == cudaSuccess && !adv_stats) {  // the minibatch's advantage statistics, computed here (one launch)
    float* stats = reinterpret_cast<float*>(p->adv_ws + 2 * kAdvBlocks + 1);
    adv_stats_kernel<<<kAdvBlocks, 256, 0, st>>>(g->returns, g->values, g->idx, g->batch, p->adv_ws, reinterpret_cast<unsigned int*>(p->adv_ws + 2 * kAdvBlocks), stats,
                                                 (const float4*)g->packed_rows);
    e = cudaGetLastError();
    adv_stats = stats;
  }
  if (e == cudaSuccess) {
    StepArgs a = {p->packed, g->obs, g->batch, 0, 0, 0, 1, nullptr, nullptr, nullptr, nullptr};
    TrainArgs t = {g->actions, g->returns, g->values, g->neglogp, g->idx, adv_stats, g->cliprange, g->cliprange_vf, g->vf_coef,
                   (float)(1.0 / (double)g->batch), (uint4*)g->a1, (uint4*)g->h, (uint4*)g->hpv, (uint4*)g->dz2, (uint4*)g->dz1, (uint4*)g->dout,
                   (float4*)g->per_sample, g->col_sums, (const float4*)g->packed_rows};
    const int ctas = (tiles + kGroups - 1) / kGroups;
    policy_kernel<true><<<ctas < p->grid ? ctas : p->grid, kThreads, kSmemBytes, st>>>(a, t);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    DwArgs d = {(const unsigned char*)g->a1, (const unsigned char*)g->h, (const unsigned char*)g->hpv, (const unsigned char*)g->dz2,
                (const unsigned char*)g->dz1, (const unsigned char*)g->dout, tiles, g->g1, g->g2, g->g3};
    dw_kernel<<<tiles < p->grid ? tiles : p->grid, 128, kDwSmemBytes, st>>>(d);
    e = cudaGetLastError();
  }
  cudaSetDevice(prev);
  if (e != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "qs_policy_ppo_grad: %s", cudaGetErrorString(e));
  return QS_OK;
}

int qs_policy_pack_rollout(const float* obs, const float* actions, const float* returns, const float* values, const float* neglogp, int64_t rows,
                           void* packed_out, void* stream) {
  if (!obs || !actions || !returns || !values || !neglogp || !packed_out) return qs_fail_(QS_ERR_INVALID, "qs_policy_pack_rollout: null argument");
  if (rows <= 0) return qs_fail_(QS_ERR_INVALID, "qs_policy_pack_rollout: rows must be positive");
  if (((uintptr_t)packed_out & 15) || ((uintptr_t)obs & 15) || ((uintptr_t)actions & 15)) return qs_fail_(QS_ERR_INVALID, "qs_policy_pack_rollout: obs, actions and packed_out must be 16-byte aligned");
  const int64_t elems = rows * 8;
  pack_rollout_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, (cudaStream_t)stream>>>(obs, actions, returns, values, neglogp, rows, (float4*)packed_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "qs_policy_pack_rollout: %s", cudaGetErrorString(e));
  return QS_OK;
}

int qs_adam_step(const qs_adam_tensor_t* tensors, int n_tensors, float lr, float beta1, float beta2, float eps, int64_t step, int64_t* step_counter,
                 float max_grad_norm, float* grad_norm_out, void* stream) {
  if (!tensors || n_tensors < 1 || n_tensors > kAdamMaxTensors) return qs_fail_(QS_ERR_INVALID, "qs_adam_step: 1..%d tensors", kAdamMaxTensors);
  if (step < 1 && !step_counter) return qs_fail_(QS_ERR_INVALID, "qs_adam_step: step counts from 1 (or pass a device step counter)");
  if (step < 1) step = 1;
  AdamArgs a;
  memset(&a, 0, sizeof a);
  for (int k = 0; k < n_tensors; k++) {
    const qs_adam_tensor_t& t = tensors[k];
    if (!t.param || !t.grad || !t.m || !t.v || t.n < 0 || t.n > (int64_t)1 << 30) return qs_fail_(QS_ERR_INVALID, "qs_adam_step: bad tensor %d", k);
    a.param[k] = t.param; a.grad[k] = t.grad; a.m[k] = t.m; a.v[k] = t.v; a.n[k] = (int)t.n;
  }
  a.count = n_tensors;
  // tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), evaluated in double on the host
  a.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, (double)step)) / (1.0 - pow((double)beta1, (double)step)));
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_grad_norm; a.norm_out = grad_norm_out;
  a.step_dev = reinterpret_cast<long long*>(step_counter);
  adam_kernel<<<kAdamCtas, 1024, 0, (cudaStream_t)stream>>>(a);
  if (a.step_dev) adam_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(a.step_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "qs_adam_step: %s", cudaGetErrorString(e));
  return QS_OK;
}

int qs_permutation(int64_t n, uint64_t seed, int64_t* out, void* stream) {
  if (n <= 0 || !out) return qs_fail_(QS_ERR_INVALID, "qs_permutation: n must be positive and out non-null");
  if (n > (int64_t)1 << 40) return qs_fail_(QS_ERR_INVALID, "qs_permutation: n too large");
  int bits = 2;
  while (((int64_t)1 << bits) < n) bits += 2;  // an even number of bits: two equal Feistel halves
  permutation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, bits / 2, seed, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "qs_permutation: %s", cudaGetErrorString(e));
  return QS_OK;
}

int qs_rollout(qs_env_t* env, qs_policy_t* pol, int64_t n_steps, const qs_rollout_t* b, uint64_t seed, uint32_t step_idx0, int n_slices,
               void* stream) {
  if (!env || !pol || !b || !b->obs || !b->actions || !b->values || !b->neglogp || !b->rewards || !b->dones || !b->clipped)
    return qs_fail_(QS_ERR_INVALID, "qs_rollout: null argument");
  if (n_steps < 0 || n_slices < 0 || n_slices > kMaxSlices) return qs_fail_(QS_ERR_INVALID, "qs_rollout: n_steps < 0 or n_slices not in [0, %d]", kMaxSlices);
  if (qs_env_dtype(env) != QS_F32 || qs_obs_dim(env) != kObs)
    return qs_fail_(QS_ERR_INVALID, "qs_rollout: needs a float32 docking environment (12 observations, what the policy was built for)");
  if (qs_env_device(env) != pol->device)
    return qs_fail_(QS_ERR_INVALID, "qs_rollout: the environment lives on CUDA device %d, the policy on device %d", qs_env_device(env), pol->device);
  NvtxRange nv("qs_rollout");
  const int64_t n = qs_num_envs(env);
  const uint64_t id0 = qs_env_id0(env);
  if (n_slices == 0) {
    // ---- fused: ONE launch per step does model.step and env.step (policy_env_step_kernel); then the bootstrap value pass
    EnvStepArgs ea;
    int kind = 0;
    if (!kFuFits) return qs_fail_(QS_ERR_INVALID, "qs_rollout: this build has no fused step kernel (n_slices = 0)");
    if (qs_env_fused_args_(env, &ea, sizeof ea, &kind) != QS_OK)
      return qs_fail_(QS_ERR_INVALID, "qs_rollout: the fused step kernel (n_slices = 0) needs a float32 docking handle with the default model parameters");
    if (!pol->have_weights) return qs_fail_(QS_ERR_INVALID, "qs_rollout: qs_policy_set_weights has not been called");
    if (pol->precision != QS_POLICY_BF16) return qs_fail_(QS_ERR_INVALID, "qs_rollout: the fused step kernel (n_slices = 0) runs the tensor-core forward only (QS_POLICY_BF16)");
    int prev = 0;
    QSP_CUDA(cudaGetDevice(&prev));
    QSP_CUDA(cudaSetDevice(pol->device));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool pdl = [] { const char* v = getenv("QUADSIM_B200_PDL"); return !(v && v[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    const int ws_ctas = (int)(((n + kTileRows - 1) / kTileRows + kWsGroups - 1) / kWsGroups);
    cfg.gridDim = dim3(ws_ctas < pol->grid ? ws_ctas : pol->grid); cfg.blockDim = dim3(kFuThreads); cfg.dynamicSmemBytes = kFuSmemBytes;
    cfg.stream = (cudaStream_t)stream; cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t le = cudaSuccess;
    for (int64_t t = 0; t < n_steps && le == cudaSuccess; t++) {
      StepArgs a = {pol->packed, b->obs + t * n * kObs, n, seed, id0, step_idx0 + (uint32_t)t, 0, b->actions + t * n * kAct, nullptr,
                    b->values + t * n, b->neglogp + t * n};
      ea.actions = nullptr; ea.obs = b->obs + (t + 1) * n * kObs; ea.reward = b->rewards + t * n; ea.done = b->dones + (t + 1) * n;
      le = kind == QS_DOCK_V0 ? cudaLaunchKernelEx(&cfg, policy_env_step_kernel<kDockV0>, a, ea) : cudaLaunchKernelEx(&cfg, policy_env_step_kernel<kDockV2>, a, ea);
    }
    cudaSetDevice(prev);
    if (le != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "qs_rollout: %s", cudaGetErrorString(le));
    return qs_policy_step(pol, n, b->obs + n_steps * n * kObs, seed, id0, 0, 1, nullptr, nullptr, b->values + n_steps * n, nullptr, stream);
  }
  // slice boundaries on whole 128-env tiles (the state layout's unit)
  int64_t first[kMaxSlices + 1];
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  if (n_slices > tiles) n_slices = (int)tiles;
  for (int k = 0; k <= n_slices; k++) first[k] = k == n_slices ? n : (tiles * k / n_slices) * kTileRows;
  cudaStream_t st[kMaxSlices];
  for (int k = 0; k < kMaxSlices; k++) st[k] = (cudaStream_t)stream;
  if (n_slices > 1) {
    int prev = 0;
    QSP_CUDA(cudaGetDevice(&prev));
    QSP_CUDA(cudaSetDevice(pol->device));
    cudaError_t e = cudaSuccess;
    if (!pol->ev_fork) {
      for (int k = 0; k < kMaxSlices && e == cudaSuccess; k++) {
        e = cudaStreamCreateWithFlags(&pol->slice_stream[k], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pol->ev_join[k], cudaEventDisableTiming);
      }
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pol->ev_fork, cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventRecord(pol->ev_fork, (cudaStream_t)stream);
    for (int k = 0; k < n_slices && e == cudaSuccess; k++) {
      st[k] = pol->slice_stream[k];
      e = cudaStreamWaitEvent(st[k], pol->ev_fork, 0);
    }
    cudaSetDevice(prev);
    if (e != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "qs_rollout: %s", cudaGetErrorString(e));
  }
  int rc = QS_OK;
  for (int64_t t = 0; t <= n_steps && rc == QS_OK; t++) {
    for (int k = 0; k < n_slices && rc == QS_OK; k++) {
      const int64_t f = first[k], m = first[k + 1] - first[k];
      if (t == n_steps) {
        // last_values = model.value(obs) (ppo2.py:510): deterministic pass, only the value head is stored
        rc = policy_step_impl(pol, m, b->obs + (t * n + f) * kObs, seed, id0 + f, 0, 1, nullptr, nullptr, b->values + t * n + f, nullptr, st[k], false);
        continue;
      }
      // model.step(obs) (ppo2.py:472) writes straight into the rollout's rows; np.clip to the Box (ppo2.py:478-481) -> clipped
      rc = policy_step_impl(pol, m, b->obs + (t * n + f) * kObs, seed, id0 + f, step_idx0 + (uint32_t)t, 0, b->actions + (t * n + f) * kAct,
                            b->clipped + f * kAct, b->values + t * n + f, b->neglogp + t * n + f, st[k], n_slices > 1);
      // env.step(clipped_actions) (ppo2.py:482): the next observations / this step's rewards / the next row of dones
      if (rc == QS_OK)
        rc = qs_step_range(env, f, m, b->clipped, b->obs + (t + 1) * n * kObs, b->rewards + t * n, b->dones + (t + 1) * n, st[k]);
    }
  }
  if (n_slices > 1) {  // join even after an error so the caller's stream stays ordered behind whatever was enqueued
    for (int k = 0; k < n_slices; k++)
      if (cudaEventRecord(pol->ev_join[k], st[k]) != cudaSuccess || cudaStreamWaitEvent((cudaStream_t)stream, pol->ev_join[k], 0) != cudaSuccess)
        if (rc == QS_OK) rc = qs_fail_(QS_ERR_CUDA, "qs_rollout: stream join failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  return rc;
}

}  // extern "C"
