// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 operands, M = 128, K = 16, cta_group::1) as a function of N and of
// where A comes from (shared memory or tensor memory), issued back to back by one thread per SM -- the per-instruction cost
// the policy kernels' layer shapes pay.  Operands are whatever is in shared memory: only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/_build/umma_rate tools/microbench/umma_rate.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <bool A_TMEM>
__global__ void umma_kernel(int n, int reps, unsigned long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) unsigned long long bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;  // finite bf16 values
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem), tmem = slot, bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    const uint32_t idesc = umma_idesc(128, n);
    const uint64_t adesc = umma_desc(sbase, 2048, 128), bdesc = umma_desc(sbase + 8192, (uint32_t)(n / 8) * 128, 128);
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      if (A_TMEM)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem), "r"(tmem + 256), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_a), "r"(0u) : "memory");
    cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount, reps = 4000;
  unsigned long long* cyc;
  cudaMallocManaged(&cyc, sms * sizeof *cyc);
  cudaFuncSetAttribute(umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  cudaFuncSetAttribute(umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  printf("{\"gpu\": \"%s\", \"what\": \"cycles per tcgen05.mma kind::f16 bf16 M=128 K=16, back to back, one issuing thread per SM (max over SMs)\", \"results\": [", prop.name);
  bool first = true;
  for (int a_tmem = 0; a_tmem < 2; a_tmem++)
    for (int n : {16, 32, 64, 128, 256}) {
      for (int pass = 0; pass < 2; pass++) {
        if (a_tmem) umma_kernel<true><<<sms, 128, 48 * 1024>>>(n, reps, cyc); else umma_kernel<false><<<sms, 128, 48 * 1024>>>(n, reps, cyc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      }
      unsigned long long mx = 0;
      for (int i = 0; i < sms; i++) mx = cyc[i] > mx ? cyc[i] : mx;
      printf("%s{\"a\": \"%s\", \"n\": %d, \"cycles_per_mma\": %.1f, \"macs_per_clk_per_sm\": %.0f}", first ? "" : ", ", a_tmem ? "tmem" : "smem", n, (double)mx / reps,
             128.0 * n * 16 * reps / (double)mx);
      first = false;
    }
  printf("]}\n");
  return 0;
}
