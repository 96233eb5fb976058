// qs_torch_ext.cpp -- the PyTorch extension face of libquadsim_b200.so (BASELINE.json north_star: "a batched VecEnv whose
// Python host code calls the CUDA library through a thin C-ABI via a PyTorch extension").
//
// Nothing here computes: each function checks the tensors it is given (device, dtype, shape, contiguity), takes the CURRENT
// CUDA stream of the tensors' device from PyTorch, and forwards plain pointers to the C ABI (include/quadsim_b200.h).  What it
// buys over the ctypes binding (quadsim_b200/_capi.py, kept for everything that is not on the per-step path) is the per-call
// cost: one pybind11 call with tensor arguments instead of seven ctypes conversions plus a python-level stream lookup --
// which is what bounds small batches (BASELINE config 2: 4096 envs) and the 35 us hovering kernel.
// The reference's per-step entry point this serves: VecEnv.step, rl_baselines/ppo2/ppo2.py:482.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "../../include/quadsim_b200.h"

namespace {

inline void check(const at::Tensor& t, const char* name, at::ScalarType dt, int64_t rows, int64_t cols, int device) {
  TORCH_CHECK(t.is_cuda() && t.get_device() == device, name, " must live on cuda:", device);
  TORCH_CHECK(t.scalar_type() == dt, name, " has dtype ", t.scalar_type(), ", the handle needs ", dt);
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
  TORCH_CHECK(cols < 0 ? (t.dim() == 1 && t.size(0) == rows) : (t.dim() == 2 && t.size(0) == rows && t.size(1) == cols), name, " has the wrong shape");
}

struct Handle {
  qs_env_t* env;
  int64_t n;
  int obs_dim, device;
  at::ScalarType dt;
};

// One cached view of a handle per Python-side env: sizes / dtype / device are queried once, not per step.
int64_t bind(int64_t handle) {
  auto* env = reinterpret_cast<qs_env_t*>(handle);
  auto* h = new Handle{env, qs_num_envs(env), qs_obs_dim(env), qs_env_device(env), qs_env_dtype(env) == QS_F32 ? at::kFloat : at::kDouble};
  return reinterpret_cast<int64_t>(h);
}
void unbind(int64_t bound) { delete reinterpret_cast<Handle*>(bound); }

// VecEnv.step on device tensors.  extras: address of a qs_step_extras_t the caller keeps alive (0 = none).  Returns qs_status.
int64_t step(int64_t bound, const at::Tensor& actions, const at::Tensor& obs, const at::Tensor& reward, const at::Tensor& done, int64_t extras) {
  const Handle* h = reinterpret_cast<const Handle*>(bound);
  check(actions, "actions", h->dt, h->n, 4, h->device);
  check(obs, "obs", h->dt, h->n, h->obs_dim, h->device);
  check(reward, "reward", h->dt, h->n, -1, h->device);
  TORCH_CHECK(done.is_cuda() && done.get_device() == h->device && done.is_contiguous() && done.numel() == h->n && done.element_size() == 1,
              "done must be a contiguous [N] uint8 / bool tensor on the handle's device");
  return qs_step_ex(h->env, actions.data_ptr(), obs.data_ptr(), reward.data_ptr(), static_cast<uint8_t*>(done.data_ptr()),
                    reinterpret_cast<const qs_step_extras_t*>(extras), at::cuda::getCurrentCUDAStream(h->device).stream());
}

int64_t reset(int64_t bound, const at::Tensor& obs) {
  const Handle* h = reinterpret_cast<const Handle*>(bound);
  check(obs, "obs", h->dt, h->n, h->obs_dim, h->device);
  return qs_reset(h->env, obs.data_ptr(), at::cuda::getCurrentCUDAStream(h->device).stream());
}

// T steps in one launch (qs_step_n): actions [T,N,4] -> obs [T,N,obs_dim], reward [T,N], done [T,N]
int64_t step_n(int64_t bound, const at::Tensor& actions, const at::Tensor& obs, const at::Tensor& reward, const at::Tensor& done) {
  const Handle* h = reinterpret_cast<const Handle*>(bound);
  TORCH_CHECK(actions.is_cuda() && actions.get_device() == h->device && actions.scalar_type() == h->dt && actions.is_contiguous() && actions.dim() == 3 &&
                  actions.size(1) == h->n && actions.size(2) == 4, "actions must be a contiguous [T, N, 4] tensor of the handle's dtype on its device");
  const int64_t T = actions.size(0);
  TORCH_CHECK(obs.is_cuda() && obs.scalar_type() == h->dt && obs.is_contiguous() && obs.numel() == T * h->n * h->obs_dim, "obs must be [T, N, obs_dim]");
  TORCH_CHECK(reward.is_cuda() && reward.scalar_type() == h->dt && reward.is_contiguous() && reward.numel() == T * h->n, "reward must be [T, N]");
  TORCH_CHECK(done.is_cuda() && done.is_contiguous() && done.numel() == T * h->n && done.element_size() == 1, "done must be [T, N] uint8 / bool");
  return qs_step_n(h->env, T, actions.data_ptr(), obs.data_ptr(), reward.data_ptr(), static_cast<uint8_t*>(done.data_ptr()),
                   at::cuda::getCurrentCUDAStream(h->device).stream());
}

std::string last_error() { return qs_last_error(); }

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "PyTorch-extension face of libquadsim_b200.so: tensors in, current CUDA stream, plain pointers to the C ABI";
  m.def("bind", &bind);
  m.def("unbind", &unbind);
  m.def("step", &step);
  m.def("reset", &reset);
  m.def("step_n", &step_n);
  m.def("last_error", &last_error);
  m.def("abi_version", [] { return qs_abi_version(); });
}
