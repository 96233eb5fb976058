// Micro-benchmark: what HBM bandwidth does the ACCESS PATTERN of the fused docking step reach with no arithmetic at all?
// MEASURED_PEAKS.json's hbm_gbs is a 4 GB torch copy; the step kernel's launch moves 389 B per env (fp32: 10 tile-major state
// rows in and out, one action row in, 12 observation words + reward + done out) over a 408 MB working set that is re-read by
// the next launch, part of which the 126 MB L2 still holds.  This kernel does exactly those loads and stores (a trivial
// dependency between them so nothing is optimised away), launched back to back like the step loop; its number is the
// ceiling the step kernel's 405 B/env-step "algorithmic" figure should be read against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/_build/step_pattern_bw tools/microbench/step_pattern_bw.cu
//   ./step_pattern_bw [log2_envs=20] [steps=300]   -> one JSON line
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

template <typename V, int NR, int AR, int OW>  // V: 16-byte row type; NR state rows; AR action rows; OW observation words of 4/8 B
__global__ void __launch_bounds__(128) pattern_kernel(V* __restrict__ state, const V* __restrict__ actions, V* __restrict__ obs,
                                                      decltype(V().x)* __restrict__ reward, uint8_t* __restrict__ done, int64_t n) {
  // programmatic dependent launch, exactly as the step kernels use it (no-ops without the launch attribute)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  V* rows = state + (i / 128) * (NR * 128) + (i % 128);
  V w[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) w[r] = rows[r * 128];
  V a[AR];
#pragma unroll
  for (int r = 0; r < AR; r++) a[r] = actions[i * AR + r];
  decltype(V().x) acc = a[0].x;
#pragma unroll
  for (int r = 0; r < NR; r++) { acc += w[r].y; w[r].x += a[AR - 1].y; }
  constexpr int OR = OW * sizeof(acc) / 16;  // observation rows of 16 B
#pragma unroll
  for (int r = 0; r < OR; r++) { V o = w[r]; o.y = acc; obs[i * OR + r] = o; }
  reward[i] = acc;
  done[i] = acc > 1e30f;
#pragma unroll
  for (int r = 0; r < NR; r++) rows[r * 128] = w[r];
}

template <typename V, int NR, int AR, int OW> double run(int64_t n, int steps, bool pdl, const char* name, double algo_bytes, double layout_bytes) {
  using S = decltype(V().x);
  V *state, *obs, *actions[8];
  S* reward;
  uint8_t* done;
  const int64_t tiles = (n + 127) / 128;
  cudaMalloc(&state, tiles * NR * 128 * sizeof(V));
  cudaMemset(state, 0, tiles * NR * 128 * sizeof(V));
  for (auto& a : actions) { cudaMalloc(&a, n * AR * sizeof(V)); cudaMemset(a, 0, n * AR * sizeof(V)); }
  cudaMalloc(&obs, n * OW * sizeof(S));
  cudaMalloc(&reward, n * sizeof(S));
  cudaMalloc(&done, n);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    for (int k = 0; k < steps; k++) {
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)tiles); cfg.blockDim = dim3(128); cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
      cudaLaunchKernelEx(&cfg, pattern_kernel<V, NR, AR, OW>, state, (const V*)actions[k % 8], obs, reward, done, n);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double us = best * 1e3 / steps;
  printf("{\"pattern\": \"%s\", \"pdl\": %s, \"envs\": %lld, \"us_per_launch\": %.2f, \"layout_GBps\": %.0f, \"algorithmic_GBps\": %.0f, \"env_steps_per_s_ceiling\": %.4g}\n", name,
         pdl ? "true" : "false", (long long)n, us, layout_bytes * n / us * 1e-3, algo_bytes * n / us * 1e-3, n / us * 1e6);
  return us;
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 20, steps = argc > 2 ? atoi(argv[2]) : 300;
  const int64_t n = 1ll << lg;
  for (int pdl = 0; pdl < 2; pdl++) {
  run<float4, 10, 1, 12>(n, steps, pdl, "docking fp32 (10 rows in/out, action 16 B, obs 48 B, reward 4 B, done 1 B)", 405, 389);
  run<double2, 20, 2, 12>(n, steps, pdl, "docking fp64 (20 rows in/out, action 32 B, obs 96 B, reward 8 B, done 1 B)", 793, 777);
  }
  return cudaDeviceSynchronize() != cudaSuccess;
}
