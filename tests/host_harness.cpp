// TEST-ONLY: compiles quadsim_b200/csrc/qs_core.cuh (the exact arithmetic the CUDA kernels run) as plain C++ so the
// CPU test-suite can check the fp32 "fast" path and the fp64 literal path against the reference's golden tapes
// without a GPU.  Built into tests/_build/ by tests/host_harness.py.  NEVER loaded by the quadsim_b200 package: the
// product path is the CUDA library only.
#include <cstdint>
#include <cstring>

#include "../quadsim_b200/csrc/qs_core.cuh"

using namespace qs;

namespace {
template <typename T> struct Vec {
  int kind; int64_t n; RngCfg rng; int auto_reset;
  DockEnv<T>* dock; HoverEnv<T>* hover;
  int spec = 1;         // fp64 docking: 0 = guarded arithmetic only (what the replay runs), for A/B
  int64_t replays = 0;  // fp64 docking: steps whose speculative evaluation was redone with the guarded arithmetic
  bool custom = false;  // runtime model parameters (what a handle created with use_params = 1 runs)
  RuntimeParams<T> prm;
};
// f(provider): the compiled-in constants, or the runtime parameter block -- the same switch the CUDA launch code makes
template <typename T, typename F> void with_params(const Vec<T>* v, F&& f) {
  if (v->custom) f(v->prm);
  else if (v->kind == kHover) f(DefaultParams<T, kHover>{});
  else if (v->kind == kDockV0) f(DefaultParams<T, kDockV0>{});
  else f(DefaultParams<T, kDockV2>{});
}

template <typename T> void* vec_create(int kind, int64_t n, uint64_t seed, uint64_t env_id0, int randomize, double half_width, int auto_reset,
                                       const qs_env_params_t* params) {
  auto* v = new Vec<T>();
  v->kind = kind; v->n = n; v->auto_reset = auto_reset;
  v->custom = params != nullptr;
  if (params) make_runtime_params<T>(*params, kind, v->prm);
  if (half_width <= 0) half_width = kind == kHover ? 1.0 : 0.5;
  v->rng = RngCfg{seed, env_id0, randomize, (float)(2.0 * half_width)};
  v->dock = kind == kHover ? nullptr : new DockEnv<T>[n];
  v->hover = kind == kHover ? new HoverEnv<T>[n] : nullptr;
  with_params(v, [&](auto prm) {
    T obs[13];
    for (int64_t i = 0; i < n; i++) {
      if (kind == kHover) { v->hover[i].episode = 0; hover_env_seat<T>(prm, v->hover[i], v->rng, i, obs); }
      else { v->dock[i].episode = 0; v->dock[i].dq = Quat<T>{T(1), T(0), T(0), T(0)}; dock_env_seat<T>(prm, v->dock[i], v->rng, i, obs); }
    }
  });
  return v;
}
template <typename T> void vec_reset(void* h, T* obs) {
  auto* v = (Vec<T>*)h;
  with_params(v, [&](auto prm) {
    for (int64_t i = 0; i < v->n; i++) {
      if (v->kind == kHover) { auto& e = v->hover[i]; e.episode += e.tt != 0; hover_env_seat<T>(prm, e, v->rng, i, obs + 13 * i); }
      else { auto& e = v->dock[i]; e.episode += e.tt != 0; dock_env_seat<T>(prm, e, v->rng, i, obs + 12 * i); }
    }
  });
}
// What dock_step_kernel does per env (qs_lib.cu): fp32 -> the guarded step; fp64 -> the speculative straight-line step
// (qs::sdouble) and, when its check fails, a replay of the guarded step from the original state.
template <typename P> inline void dock_step_like_kernel(const P& prm, int kind, DockEnv<float>& e, const float* act, float* rel, StepOut<float>& o, int64_t&, int) {
  if (kind == kDockV0) dock_env_step<float, kDockV0>(prm, e, act, rel, o);
  else dock_env_step<float, kDockV2>(prm, e, act, rel, o);
}
template <typename P> inline void dock_step_like_kernel(const P& prm, int kind, DockEnv<double>& e, const double* act, double* rel, StepOut<double>& o, int64_t& replays, int spec) {
  if (kFast64 && spec) {
    const DockEnv<double> e0 = e;
    const bool ok = kind == kDockV0 ? dock_env_step_spec<kDockV0>(prm, e, act, rel, o) : dock_env_step_spec<kDockV2>(prm, e, act, rel, o);
    if (ok) return;
    e = e0;
    replays++;
  }
  if (kind == kDockV0) dock_env_step<double, kDockV0>(prm, e, act, rel, o);
  else dock_env_step<double, kDockV2>(prm, e, act, rel, o);
}
template <typename T> void vec_step(void* h, const T* act, T* obs, T* rew, uint8_t* done, uint8_t* info, T* term) {
  auto* v = (Vec<T>*)h;
  with_params(v, [&](auto prm) {
  for (int64_t i = 0; i < v->n; i++) {
    StepOut<T> o;
    if (v->kind == kHover) {
      auto& e = v->hover[i];
      T ob[13];
      hover_env_step<T>(prm, e, act + 4 * i, ob, o);
      if (term) memcpy(term + 13 * i, ob, sizeof ob);
      if (o.done && v->auto_reset) { e.episode += 1; hover_env_seat<T>(prm, e, v->rng, i, ob); }
      memcpy(obs + 13 * i, ob, sizeof ob);
    } else {
      auto& e = v->dock[i];
      T rel[12];
      dock_step_like_kernel(prm, v->kind, e, act + 4 * i, rel, o, v->replays, v->spec);
      if (term) memcpy(term + 12 * i, rel, sizeof rel);
      if (o.done && v->auto_reset) { e.episode += 1; dock_env_seat<T>(prm, e, v->rng, i, rel); }
      memcpy(obs + 12 * i, rel, sizeof rel);
    }
    rew[i] = o.reward; done[i] = o.done; if (info) info[i] = (uint8_t)o.info();
  }
  });
}
// state views are float64 in world coordinates (like qs_export_state / qs_import_state); org = Origin<T> for docking
template <typename T> void put13(const Drone<T>& d, double* s, double* u, const double org[3]) {
  T o[13];
  drone_to_obs13(d, o);
  for (int k = 0; k < 13; k++) s[k] = (double)o[k] + (k < 3 ? org[k] : 0.0);
  u[0] = d.u0; u[1] = d.u1; u[2] = d.u2; u[3] = d.u3;
}
template <typename T> void get13(Drone<T>& d, const double* s, const double* u, const double org[3]) {
  d = Drone<T>{T(s[0] - org[0]), T(s[1] - org[1]), T(s[2] - org[2]), T(s[3]), T(s[4]), T(s[5]), T(s[6]), T(s[7]), T(s[8]), T(s[9]),
               T(s[10]), T(s[11]), T(s[12]), T(u[0]), T(u[1]), T(u[2]), T(u[3])};
}
// export/import: chaser[13] u_c[4] target[13] u_t[4] des_q[4] last_shaping ep_return t episode = 42 doubles per env
template <typename T> void vec_io(void* h, double* buf, int import) {
  auto* v = (Vec<T>*)h;
  const double zero[3] = {0, 0, 0};
  double org[3] = {(double)Origin<T>::x, (double)Origin<T>::y, (double)Origin<T>::z};
  if (v->custom) { org[0] = (double)v->prm.ox; org[1] = (double)v->prm.oy; org[2] = (double)v->prm.oz; }
  for (int64_t i = 0; i < v->n; i++) {
    double* b = buf + 42 * i;
    if (v->kind == kHover) {
      auto& e = v->hover[i];
      if (import) { get13(e.d, b, b + 13, zero); e.ep_return = T(b[39]); e.tt = (int)b[40]; e.episode = (int)b[41]; }
      else { memset(b, 0, 42 * sizeof(double)); put13(e.d, b, b + 13, zero); b[39] = e.ep_return; b[40] = e.tt; b[41] = e.episode; }
    } else {
      auto& e = v->dock[i];
      if (import) {
        get13(e.c, b, b + 13, org); get13(e.t, b + 17, b + 30, org); e.dq = Quat<T>{T(b[34]), T(b[35]), T(b[36]), T(b[37])};
        e.last_shaping = T(b[38]); e.tt = (int)b[40]; e.episode = (int)b[41];
      } else {
        put13(e.c, b, b + 13, org); put13(e.t, b + 17, b + 30, org); b[34] = e.dq.w; b[35] = e.dq.x; b[36] = e.dq.y; b[37] = e.dq.z;
        b[38] = e.last_shaping; b[39] = e.last_shaping; b[40] = e.tt; b[41] = e.episode;
      }
    }
  }
}
// PID expert (qs::pid_expert_*) over the dock envs; mem [n,8] and obs [n,12] in T, exactly what expert_kernel does
template <typename T> void vec_expert(void* h, T* mem, int alias, T* act, T* obs) {
  auto* v = (Vec<T>*)h;
  with_params(v, [&](auto prm) {
  for (int64_t i = 0; i < v->n; i++) {
    auto& e = v->dock[i];
    ExpertMem<T> m = {mem[8 * i], mem[8 * i + 1], mem[8 * i + 2], Quat<T>{mem[8 * i + 4], mem[8 * i + 5], mem[8 * i + 6], mem[8 * i + 7]}};
    const bool first = e.tt == 0;
    if (pid_expert_reseat(e.c, m, first, alias != 0) && obs) state2rel(prm, e.c, e.t, obs + 12 * i);
    pid_expert_cmd(prm, e.c, e.t, m, first, alias != 0, act + 4 * i);
    mem[8 * i] = m.dvx; mem[8 * i + 1] = m.dvy; mem[8 * i + 2] = m.dvz; mem[8 * i + 3] = T(0);
    mem[8 * i + 4] = m.dq.w; mem[8 * i + 5] = m.dq.x; mem[8 * i + 6] = m.dq.y; mem[8 * i + 7] = m.dq.z;
  }
  });
}
template <typename T> void vec_destroy(void* h) { auto* v = (Vec<T>*)h; delete[] v->dock; delete[] v->hover; delete v; }
}  // namespace

extern "C" {
// params: NULL = the compiled-in constants (DefaultParams); else a qs_env_params_t -> RuntimeParams (make_runtime_params)
void* hh_create(int f64, int kind, int64_t n, uint64_t seed, uint64_t env_id0, int randomize, double hw, int auto_reset, const void* params) {
  const qs_env_params_t* p = (const qs_env_params_t*)params;
  return f64 ? vec_create<double>(kind, n, seed, env_id0, randomize, hw, auto_reset, p) : vec_create<float>(kind, n, seed, env_id0, randomize, hw, auto_reset, p);
}
void hh_destroy(int f64, void* h) { f64 ? vec_destroy<double>(h) : vec_destroy<float>(h); }
void hh_reset(int f64, void* h, void* obs) { f64 ? vec_reset<double>(h, (double*)obs) : vec_reset<float>(h, (float*)obs); }
void hh_step(int f64, void* h, const void* act, void* obs, void* rew, uint8_t* done, uint8_t* info, void* term) {
  if (f64) vec_step<double>(h, (const double*)act, (double*)obs, (double*)rew, done, info, (double*)term);
  else vec_step<float>(h, (const float*)act, (float*)obs, (float*)rew, done, info, (float*)term);
}
void hh_expert(int f64, void* h, void* mem, int alias, void* act, void* obs) {
  if (f64) vec_expert<double>(h, (double*)mem, alias, (double*)act, (double*)obs);
  else vec_expert<float>(h, (float*)mem, alias, (float*)act, (float*)obs);
}
void hh_set_spec(int f64, void* h, int on) { if (f64) ((Vec<double>*)h)->spec = on; else ((Vec<float>*)h)->spec = on; }
int64_t hh_replays(int f64, void* h) { return f64 ? ((Vec<double>*)h)->replays : ((Vec<float>*)h)->replays; }
void hh_io(int f64, void* h, double* buf, int import) { f64 ? vec_io<double>(h, buf, import) : vec_io<float>(h, buf, import); }
// the fp64 scalar math wrappers (qs::f64m polynomials on the host): out[n,6] = atan2(y,x), sin(y), cos(y), rsqrt|x|, sqrt|x|, asin(y)
void hh_math(int64_t n, const double* y, const double* x, double* out) {
  for (int64_t i = 0; i < n; i++) {
    double s, c;
    qs::m_sincos(y[i], &s, &c);
    double* o = out + 6 * i;
    o[0] = qs::m_atan2(y[i], x[i]); o[1] = s; o[2] = c; o[3] = qs::m_rsqrt(fabs(x[i])); o[4] = qs::m_sqrt(fabs(x[i])); o[5] = qs::m_asin(y[i]);
  }
}
}
