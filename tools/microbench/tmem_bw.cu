// Micro-benchmark: tensor-memory read-out rate (tcgen05.ld) per SM on this GPU -- the denominator of the policy kernel's
// roofline.  One CTA per SM allocates all 512 TMEM columns; W warps (4, 8, 16: each warp may only
// touch the 32 lanes of its quadrant, warp % 4) issue tcgen05.ld.32x32b.x32 over the columns R times, with one
// tcgen05.wait::ld per 1, 2 or 4 loads in flight.  Prints bytes / clock / SM from clock64() around the loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/_build/tmem_bw tools/microbench/tmem_bw.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int INFLIGHT>
__global__ void tmem_read_kernel(int reps, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t trow = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
#pragma unroll 1
    for (int c = 0; c < 512; c += 32 * INFLIGHT) {
      uint32_t v[32 * INFLIGHT];
#pragma unroll
      for (int k = 0; k < INFLIGHT; k++) tmem_ld32(trow + c + 32 * k, v + 32 * k);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 32 * INFLIGHT; k += 8) acc ^= v[k];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount, reps = 2000;
  unsigned long long* cyc;
  uint32_t* sink;
  cudaMallocManaged(&cyc, sms * sizeof *cyc);
  cudaMalloc(&sink, 4);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"what\": \"tcgen05.ld.32x32b.x32 read-out of 512 columns x 128 lanes per SM, bytes per clock per SM\", \"results\": [", prop.name, sms);
  bool first = true;
  for (int warps : {4, 8, 16}) {
    for (int inflight : {1, 2, 4}) {
      for (int pass = 0; pass < 2; pass++) {
        if (inflight == 1) tmem_read_kernel<1><<<sms, warps * 32>>>(reps, cyc, sink);
        else if (inflight == 2) tmem_read_kernel<2><<<sms, warps * 32>>>(reps, cyc, sink);
        else tmem_read_kernel<4><<<sms, warps * 32>>>(reps, cyc, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      }
      unsigned long long mx = 0;
      for (int i = 0; i < sms; i++) mx = cyc[i] > mx ? cyc[i] : mx;
      // every warp reads its quadrant's 32 lanes x 512 columns x 4 B per rep
      const double bytes = (double)warps * 32 * 512 * 4 * reps;
      printf("%s{\"warps\": %d, \"loads_in_flight\": %d, \"bytes_per_clk_per_sm\": %.1f}", first ? "" : ", ", warps, inflight, bytes / (double)mx);
      first = false;
    }
  }
  printf("]}\n");
  return 0;
}
