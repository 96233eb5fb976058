// qs_policy.cu -- fused forward pass of the reference's PPO2 MlpPolicy on 5th-gen tensor cores (tcgen05 + TMEM).
//
// Policy: net_arch=[128, dict(pi=[128], vf=[128])], relu, diagonal Gaussian (run_moving_docking_ppo2.py:88-91;
// weights as in the shipped archives, e.g. trained_model/best_model_v0.zip):
//     h  = relu(obs[12] W1^T + b1)                          12 -> 128   (shared_fc0)
//     hp = relu(h Wp^T + bp) ; hv = relu(h Wv^T + bv)       128 -> 128  (pi_fc0, vf_fc0)
//     mean = hp W3p^T + b3p (4) ; value = hv . w3v + b3v    128 -> 4, 1 (pi, vf)
//     action = mean + exp(logstd) * N(0,1) ; neglogp        (model.step, rl_baselines/ppo2/ppo2.py:472)
// This is the CALLER of the env hot path: with the env step at ~66 us per 2^20 envs, the unfused torch
// forward (1.5-4 ms) dominated rollout-in-the-loop by 20-60x.  It is the only dense contraction in the system, so it goes
// on the tensor cores: one CTA = 128 envs = one M=128 UMMA tile.
//   layer 1:  D1[128x128] = A1[128x32] . B1[128x32]^T   obs split hi+lo in bf16 (x = hi + lo, |lo| <= 2^-9 |x|) against W1
//             duplicated along K, so inputs keep ~16 mantissa bits (a plain bf16 cast would quantise a 1.8 m range to 8 mm)
//   layer 2:  D2[128x256] = A2[128x128] . [Wp;Wv]^T      one N=256 accumulator, 8 tcgen05.mma (K=16 each)
//   layer 3 + sampling: in registers from the TMEM read-out (tcgen05.ld), 5 dot products per row on the CUDA cores.
// Operands sit in shared memory in the canonical no-swizzle K-major UMMA layout (8x8 core matrices: 8 rows x 16 bytes
// contiguous; SBO = next 8 rows, LBO = next 8 k), weights pre-packed once on the device by pack_kernel and bulk-copied
// (cp.async.bulk) per CTA.  D1 and D2 share TMEM columns (D1 is dead once the layer-1 epilogue has written A2), so a CTA
// needs 256 of the SM's 512 TMEM columns and 108 KB of shared memory: two CTAs per SM overlap one CTA's MMAs with the
// other's epilogue.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/quadsim_b200.h"
#include "qs_core.cuh"

extern int qs_fail_(int code, const char* fmt, ...);  // qs_lib.cu: sets the thread-local error string

namespace {

constexpr int kThreads = 128;
constexpr int kObs = 12, kHid = 128, kAct = 4;
constexpr int kK1 = 32;  // layer-1 K: [hi(12)+pad(4) | lo(12)+pad(4)]

// ---- shared-memory map (bytes) ----
constexpr int kOffA = 0;                          // A2 [128 x 128] bf16 (32 KB); A1 [128 x 32] aliases its first 8 KB
constexpr int kOffW1 = 32768;                     // B1 [128 x 32] bf16 (8 KB)
constexpr int kOffW2 = kOffW1 + 8192;             // B2 [256 x 128] bf16 (64 KB)
constexpr int kOffF = kOffW2 + 65536;             // fp32 tail
constexpr int kTailFloats = 128 + 256 + 512 + 128 + 8 + 4;  // b1, b2=[bp;bv], W3p[k][4], w3v[k], b3(4)+b3v+pad(3), logstd(4)
constexpr int kOffBar = kOffF + kTailFloats * 4;  // 2 mbarriers + weight barrier + tmem ptr
constexpr int kSmemBytes = kOffBar + 64;
constexpr int kPackedBytes = 8192 + 65536 + kTailFloats * 4;  // global image of [W1 | W2 | tail]
constexpr uint32_t kTmemCols = 256;

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
// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout_type = SWIZZLE_NONE [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), K-major both
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns: thread l of the warp gets lane (warp%4)*32 + l
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- one-time weight packing: torch layout [out, in] fp32 -> canonical bf16 operand images + fp32 tail ----
struct RawWeights { const float *w1, *b1, *wp, *bp, *wv, *bv, *w3p, *b3p, *w3v, *b3v, *logstd; };

__global__ void pack_kernel(RawWeights w, unsigned char* packed) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  __nv_bfloat16* W1 = reinterpret_cast<__nv_bfloat16*>(packed);
  __nv_bfloat16* W2 = reinterpret_cast<__nv_bfloat16*>(packed + 8192);
  float* tail = reinterpret_cast<float*>(packed + 8192 + 65536);
  for (int i = tid; i < kHid * kK1; i += nth) {  // B1[n][k] = W1[n][k % 16] for k % 16 < 12 (hi half and lo half see the same weights)
    const int n = i / kK1, k = i % kK1, kk = k & 15;
    W1[canon_off(n, k, kHid) / 2] = __float2bfloat16(kk < kObs ? w.w1[n * kObs + kk] : 0.f);
  }
  for (int i = tid; i < 2 * kHid * kHid; i += nth) {  // B2 = [Wp ; Wv], 256 rows
    const int n = i / kHid, k = i % kHid;
    const float v = n < kHid ? w.wp[n * kHid + k] : w.wv[(n - kHid) * kHid + k];
    W2[canon_off(n, k, 2 * kHid) / 2] = __float2bfloat16(v);
  }
  for (int i = tid; i < kTailFloats; i += nth) {
    float v = 0.f;
    if (i < 128) v = w.b1[i];
    else if (i < 256) v = w.bp[i - 128];
    else if (i < 384) v = w.bv[i - 256];
    else if (i < 896) { const int k = (i - 384) >> 2, a = (i - 384) & 3; v = w.w3p[a * kHid + k]; }  // W3p[k][a]
    else if (i < 1024) v = w.w3v[i - 896];
    else if (i < 1028) v = w.b3p[i - 1024];
    else if (i == 1028) v = w.b3v[0];
    else if (i >= 1032 && i < 1036) v = w.logstd[i - 1032];
    tail[i] = v;
  }
}

struct StepArgs {
  const unsigned char* packed;
  const float* obs;  // [N,12]
  int64_t n;
  uint64_t seed, env_id0;
  uint32_t step_idx;
  int deterministic;
  float *actions, *clipped, *values, *neglogp;  // clipped may be null
};

__global__ void __launch_bounds__(kThreads, 2) policy_step_kernel(const StepArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_w = sbase + kOffBar, bar_mma = sbase + kOffBar + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 32);
  const float* tail = reinterpret_cast<const float*>(smem + kOffF);
  const int tiles = (int)((a.n + kThreads - 1) / kThreads);

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_w, kPackedBytes);  // weights: W1 | W2 | tail are contiguous in smem and in the packed image
    for (int off = 0; off < kPackedBytes; off += 16384) {
      const int bytes = kPackedBytes - off < 16384 ? kPackedBytes - off : 16384;
      bulk_g2s(sbase + kOffW1 + off, a.packed + off, bytes, bar_w);
    }
  }
  if (warp == 0) {  // TMEM: 256 columns for this CTA (two CTAs fit an SM's 512)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes
  mbar_wait(bar_w, 0);

  constexpr uint32_t idesc1 = umma_idesc(128, 128), idesc2 = umma_idesc(128, 256);
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t i = (int64_t)tile * kThreads + tid;
    const bool active = i < a.n;
    // ---- A1: this thread's observation row, split hi/lo, into the canonical layout (row = tid, K = 32) ----
    {
      float o[16];
#pragma unroll
      for (int k = 0; k < 16; k++) o[k] = 0.f;
      if (active) {
        const float4* p = reinterpret_cast<const float4*>(a.obs + i * kObs);
        const float4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
        o[0] = v0.x; o[1] = v0.y; o[2] = v0.z; o[3] = v0.w; o[4] = v1.x; o[5] = v1.y; o[6] = v1.z; o[7] = v1.w;
        o[8] = v2.x; o[9] = v2.y; o[10] = v2.z; o[11] = v2.w;
      }
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float h0 = __bfloat162float(__float2bfloat16(o[2 * k])), h1 = __bfloat162float(__float2bfloat16(o[2 * k + 1]));
        hi[k] = pack_bf16(h0, h1);
        lo[k] = pack_bf16(o[2 * k] - h0, o[2 * k + 1] - h1);
      }
      unsigned char* A1 = smem + kOffA;
      *reinterpret_cast<uint4*>(A1 + canon_off(tid, 0, 128)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(A1 + canon_off(tid, 8, 128)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      *reinterpret_cast<uint4*>(A1 + canon_off(tid, 16, 128)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(A1 + canon_off(tid, 24, 128)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    }
    proxy_fence();       // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    tc_fence_before();   // orders the previous tile's tcgen05.ld before the barrier (TMEM columns are reused)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // layer 1: K = 32 -> two K=16 steps; operand rows = 128 -> SBO = 128 B, LBO = 16 row-groups * 128 B = 2048 B
#pragma unroll
      for (int ks = 0; ks < 2; ks++)
        umma_bf16(tmem, umma_desc(sbase + kOffA + ks * 4096, 2048, 128), umma_desc(sbase + kOffW1 + ks * 4096, 2048, 128), idesc1, ks);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 1: h = relu(D1 + b1) -> bf16 -> A2 (row = tid), 32 columns at a time ----
#pragma unroll 1
    for (int c0 = 0; c0 < kHid; c0 += 32) {
      float v[32];
      tmem_ld32(trow + c0, v);
#pragma unroll
      for (int g = 0; g < 4; g++) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = c0 + g * 8 + j * 2;
          w[j] = pack_bf16(fmaxf(v[g * 8 + j * 2] + tail[c], 0.f), fmaxf(v[g * 8 + j * 2 + 1] + tail[c + 1], 0.f));
        }
        *reinterpret_cast<uint4*>(smem + kOffA + canon_off(tid, c0 + g * 8, 128)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // layer 2: D2[128 x 256] = A2[128 x 128] . [Wp;Wv]^T ; B rows = 256 -> LBO = 32 row-groups * 128 B = 4096 B
#pragma unroll
      for (int ks = 0; ks < 8; ks++)
        umma_bf16(tmem, umma_desc(sbase + kOffA + ks * 4096, 2048, 128), umma_desc(sbase + kOffW2 + ks * 8192, 4096, 128), idesc2, ks);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 2: hp/hv = relu(D2 + b2); layer 3 on the CUDA cores: mean[4] and value ----
    float m0 = tail[1024], m1 = tail[1025], m2 = tail[1026], m3 = tail[1027], val = tail[1028];
#pragma unroll 1
    for (int c0 = 0; c0 < 2 * kHid; c0 += 32) {
      float v[32];
      tmem_ld32(trow + c0, v);
      if (c0 < kHid) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float h = fmaxf(v[j] + tail[128 + c0 + j], 0.f);
          const float4 w3 = *reinterpret_cast<const float4*>(tail + 384 + (c0 + j) * 4);
          m0 = fmaf(h, w3.x, m0); m1 = fmaf(h, w3.y, m1); m2 = fmaf(h, w3.z, m2); m3 = fmaf(h, w3.w, m3);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) val = fmaf(fmaxf(v[j] + tail[128 + c0 + j], 0.f), tail[896 + c0 - kHid + j], val);
      }
    }
    // ---- sample, log-prob, store (DiagGaussianProbabilityDistribution: sample / neglogp) ----
    if (active) {
      const float ls0 = tail[1032], ls1 = tail[1033], ls2 = tail[1034], ls3 = tail[1035];
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
      const float a0 = fmaf(__expf(ls0), n0, m0), a1 = fmaf(__expf(ls1), n1, m1), a2 = fmaf(__expf(ls2), n2, m2), a3 = fmaf(__expf(ls3), n3, m3);
      // neglogp = 0.5 * sum(((a - mean)/std)^2) + 0.5*log(2 pi)*4 + sum(logstd); (a - mean)/std == the noise itself
      const float nlp = 0.5f * (n0 * n0 + n1 * n1 + n2 * n2 + n3 * n3) + 3.6757541328186907f + (ls0 + ls1 + ls2 + ls3);
      reinterpret_cast<float4*>(a.actions)[i] = make_float4(a0, a1, a2, a3);
      if (a.clipped)
        reinterpret_cast<float4*>(a.clipped)[i] = make_float4(fminf(fmaxf(a0, -1.f), 1.f), fminf(fmaxf(a1, -1.f), 1.f), fminf(fmaxf(a2, -1.f), 1.f), fminf(fmaxf(a3, -1.f), 1.f));
      a.values[i] = val;
      a.neglogp[i] = nlp;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace

struct qs_policy {
  int device = 0;
  unsigned char* packed = nullptr;
  int grid = 0;
  bool have_weights = false;
};

#define QSP_CUDA(call)                                                                                          \
  do {                                                                                                          \
    cudaError_t e_ = (call);                                                                                    \
    if (e_ != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

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
      (err = cudaFuncSetAttribute(policy_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)) != cudaSuccess ||
      (err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, policy_step_kernel, kThreads, kSmemBytes)) != cudaSuccess) {
    int rc = qs_fail_(QS_ERR_CUDA, "qs_policy_create: %s", cudaGetErrorString(err));
    cudaFree(p->packed);
    delete p;
    cudaSetDevice(prev);
    return rc;
  }
  if (per_sm > 2) per_sm = 2;  // TMEM: 2 x 256 columns
  if (per_sm < 1) per_sm = 1;
  p->grid = per_sm * prop.multiProcessorCount;
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
  cudaSetDevice(prev);
  delete p;
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
  pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(r, p->packed);
  cudaError_t e = cudaGetLastError();
  cudaSetDevice(prev);
  if (e != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "qs_policy_set_weights: %s", cudaGetErrorString(e));
  p->have_weights = true;
  return QS_OK;
}

int qs_policy_step(qs_policy_t* p, int64_t n, const float* obs, uint64_t seed, uint64_t env_id0, uint32_t step_idx, int deterministic,
                   float* actions_out, float* clipped_out, float* values_out, float* neglogp_out, void* stream) {
  if (!p || !obs || !actions_out || !values_out || !neglogp_out) return qs_fail_(QS_ERR_INVALID, "qs_policy_step: null argument");
  if (!p->have_weights) return qs_fail_(QS_ERR_INVALID, "qs_policy_step: qs_policy_set_weights has not been called");
  if (n <= 0) return QS_OK;
  int prev = 0;
  QSP_CUDA(cudaGetDevice(&prev));
  QSP_CUDA(cudaSetDevice(p->device));
  StepArgs a = {p->packed, obs, n, seed, env_id0, step_idx, deterministic, actions_out, clipped_out, values_out, neglogp_out};
  const int tiles = (int)((n + kThreads - 1) / kThreads);
  policy_step_kernel<<<tiles < p->grid ? tiles : p->grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  cudaSetDevice(prev);
  if (e != cudaSuccess) return qs_fail_(QS_ERR_CUDA, "qs_policy_step: %s", cudaGetErrorString(e));
  return QS_OK;
}

}  // extern "C"
