/* quadsim_b200.h -- C ABI of the B200-native batched QuadSim environment library (libquadsim_b200.so).
 *
 * The reference (sd196821/QuadSim) is pure Python and has NO FFI/plugin boundary; its boundary for this hot
 * path is the gym / stable-baselines VecEnv protocol.  Each entry point below names the
 * reference interface it stands in for.  All `file:line` citations are into the reference tree.
 *
 * Conventions
 *   - plain C types only; `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *   - every call returns QS_OK (0) or a negative qs_status and never throws; qs_last_error() (thread-local)
 *     gives the message;
 *   - "device pointer" arguments are caller-owned, contiguous, in the handle's dtype (float for QS_F32, double
 *     for QS_F64); all work is enqueued on the caller's stream, with no hidden synchronisation on the step path
 *     (qs_step_host is the documented exception: it is a blocking call on host buffers, see there);
 *   - a handle is bound to one CUDA device, owns all persistent per-env state, and is not thread-safe.
 */
#ifndef QUADSIM_B200_H
#define QUADSIM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QS_ABI_VERSION 2   /* v2: qs_env_params_t in qs_config_t, qs_step_extras_t, qs_step_host takes a stream */

typedef enum {
  QS_OK = 0,
  QS_ERR_INVALID = -1,    /* bad argument / configuration */
  QS_ERR_CUDA = -2,       /* a CUDA runtime call failed (message has the CUDA error string) */
  QS_ERR_NO_DEVICE = -3,  /* no usable CUDA device: the library has no CPU fallback */
  QS_ERR_ALLOC = -4
} qs_status;

/* gym ids registered by gym-docking/gym_docking/__init__.py:3-21 */
typedef enum {
  QS_HOVER = 0,   /* 'hovering-v0'  HoveringEnv        gym_docking/envs/hovering_env.py:10-95 */
  QS_DOCK_V0 = 1, /* 'docking-v0'   DockingEnv         gym_docking/envs/docking_env.py:12-254 */
  QS_DOCK_V2 = 2  /* 'docking-v2'   MovingDockingEnv   gym_docking/envs/moving_docking_env.py:12-219 */
} qs_env_kind;

typedef enum { QS_F32 = 0, QS_F64 = 1 } qs_dtype;

/* Initial-state randomisation (Philox4x32-10, key = seed, counter = (env_id lo, env_id hi, episode, stream)).
 * The reference has the docking jitter commented out (docking_env.py:34) and samples the hover / docking-v1 init
 * once at construction (hovering_env.py:23-24, imitating_docking_env.py:34). */
typedef enum {
  QS_RAND_NONE = 0,    /* reference docking-v0/v2 behaviour: fixed initial state */
  QS_RAND_EPISODE = 1, /* new jitter every episode (north-star auto-reset) */
  QS_RAND_ONCE = 2     /* one jitter per env for its lifetime (hovering-v0 / docking-v1 behaviour) */
} qs_randomize;

/* bits of info_out (per env, per step); 'flag_docking' and 'done_overlimit' are the reference's info keys
 * (docking_env.py:226-229) */
enum { QS_INFO_DONE = 1, QS_INFO_OVERLIMIT = 2, QS_INFO_DOCKING = 4, QS_INFO_OVERTIME = 8 };

typedef struct qs_env qs_env_t;

/* Model parameters.  The reference has no configuration system: every constant is a literal in its sources
 * (dynamics/quadrotor.py:10-63, controller/PIDController.py:10-44, the envs' __init__).  qs_default_params() fills this
 * block with exactly those literals; a handle created with use_params = 0 runs the kernels that have them compiled in as
 * immediates, a handle with use_params = 1 runs the same kernels reading this block from the constant bank (it travels
 * as a __grid_constant__ kernel parameter: per handle, no global state). */
typedef struct {
  double dt, gravity, mass, arm_length;   /* quadrotor.py:12-21   0.02, 9.81, 0.18, 0.086 */
  double inertia[3];                      /* quadrotor.py:23      diag(2.5e-4, 2.32e-4, 3.738e-4) */
  double motor_lambda;                    /* quadrotor.py:43-45   km / kf = 1.5e-9 / 6.11e-8 */
  double att_limit_deg[3];                /* quadrotor.py:35-39, :146-168   roll, pitch, yaw clamp: 85, 85, 175 */
  double kp_att[3], kd_att[3];            /* PIDController.py:10-21   (-10,-10,-9.5), (5.1,5.1,4) */
  double kp_pos[3], kd_pos[3];            /* PIDController.py:23-33   (-1,-1,50), (-1.65,-1.65,8) */
  double kp_vel[3];                       /* PIDController.py:35-44   (-0.7,-0.7,1); the kd_v* terms are identically 0 */
  double chaser_ini_pos[3];               /* docking_env.py:34    (8,-50,5);  hover: hovering_env.py:23 (0,0,5) */
  double target_ini_pos[3];               /* docking_env.py:47    (10,-50,5) */
  double target_des_pos[3];               /* docking_env.py:59    (10,-50,5)  docking-v0 set-point */
  double target_des_vel[3];               /* moving_docking_env.py:62-66  (0.2,0,0)  docking-v2 */
  double dock_port;                       /* docking_env.py:38,51  chaser port +0.1 m, target port -0.1 m along body x */
  double dock_pos_tol, dock_vel_tol, dock_angle_deg; /* docking_env.py:130-134  0.1 m, 0.1 m/s, 10 deg */
  double range_limit;                     /* docking_env.py:141 (3) / moving_docking_env.py:148 (10); hover: hovering_env.py:68 (100) */
  double min_altitude;                    /* docking_env.py:142   0.1 */
  int32_t max_steps;                      /* docking_env.py:152   600 */
  int32_t reserved;
} qs_env_params_t;

typedef struct {
  int32_t env_kind;   /* qs_env_kind */
  int32_t dtype;      /* qs_dtype */
  int64_t n_envs;     /* envs owned by this handle (one handle per GPU / rank) */
  uint64_t seed;      /* Philox key */
  uint64_t env_id0;   /* global id of this handle's env 0 (rank * n_envs): trajectories do not depend on the GPU count */
  int32_t randomize;  /* qs_randomize */
  int32_t device;     /* CUDA device ordinal */
  double half_width;  /* jitter half-width; <= 0 selects the default (0.5 docking, 1.0 hover position) */
  int32_t auto_reset; /* 1: stable-baselines VecEnv semantics (done envs re-seat inside the step);
                         0: plain gym.Env semantics (state stays terminal until qs_reset) */
  int32_t use_params; /* 0: the reference's literals (compiled in); 1: `params` below */
  qs_env_params_t params;
} qs_config_t;

/* Device pointers to array-of-struct views of the per-env state: ALWAYS float64 (t/episode: int32), whatever the handle's
 * dtype, so that export -> import is bit-transparent (the fp32 layout keeps positions relative to a fixed origin; the
 * conversion to world coordinates is exact in float64).  NULL members are skipped.  Used for parity injection and for
 * env checkpoint/restore.
 * Fields mirror the reference objects: env.chaser.state / env.target.state (quadrotor.py:26), Drone.u (:33),
 * env.target_state_des[6:10] (docking_env.py:60-64), env.last_shaping (:27), env.t (:22). */
typedef struct {
  void* chaser;        /* [N,13]  (hover: the single drone) */
  void* target;        /* [N,13]  docking only */
  void* u_chaser;      /* [N,4]   previous limited command (one-step latency) */
  void* u_target;      /* [N,4]   docking only */
  void* des_q;         /* [N,4]   docking only */
  void* last_shaping;  /* [N]     docking only */
  void* ep_return;     /* [N] */
  int32_t* t;          /* [N]     steps taken in the current episode */
  int32_t* episode;    /* [N]     episode index (Philox counter word 2) */
} qs_state_arrays_t;

int qs_abi_version(void);
const char* qs_last_error(void);
/* The reference's literals for `env_kind` (see qs_env_params_t). */
int qs_default_params(int env_kind, qs_env_params_t* out);

/* gym.make(id) / SubprocVecEnv([...]*n)  (run_moving_docking_ppo2.py:65-67): build n_envs environments. */
int qs_create(qs_env_t** out, const qs_config_t* cfg);
int qs_destroy(qs_env_t* env);
/* observation_space.shape[0]: 12 for docking (docking_env.py:85-95), 13 for hover (hovering_env.py:36-40). */
int qs_obs_dim(const qs_env_t* env);
int64_t qs_num_envs(const qs_env_t* env);
int qs_env_dtype(const qs_env_t* env);       /* qs_dtype of the handle */
int qs_env_device(const qs_env_t* env);      /* CUDA device ordinal the handle is bound to */
uint64_t qs_env_id0(const qs_env_t* env);    /* global id of env 0 (qs_config_t.env_id0) */
/* env.seed(seed)  (docking_env.py:252-254; the reference never reads it -- here it re-keys Philox). */
int qs_seed(qs_env_t* env, uint64_t seed);

/* VecEnv.reset() -> obs[N,obs_dim]  (docking_env.py:233-244, hovering_env.py:80-82).  Every env that has stepped in
 * its current episode starts its next episode; envs that have not stepped stay (idempotent). */
int qs_reset(qs_env_t* env, void* obs_out, void* stream);

/* VecEnv.step(actions[N,4]) -> obs[N,obs_dim], reward[N], done[N]  (rl_baselines/ppo2/ppo2.py:482;
 * docking_env.py:104-231, moving_docking_env.py:111-192, hovering_env.py:47-78) with the stable-baselines worker's
 * auto-reset: for done envs obs is the reset observation while reward/done are the terminal step's. */
int qs_step(qs_env_t* env, const void* actions, void* obs_out, void* reward_out, uint8_t* done_out, void* stream);
/* Optional per-step outputs (any member may be NULL).  info: [N] bitmask (QS_INFO_*).  terminal_obs: [N,obs_dim], the
 * pre-reset observation (SB's info['terminal_observation']).  terminal_chaser / terminal_target: [N,13] float64 world
 * coordinates -- the reference's info['chaser'] / info['target'] (docking_env.py:226-229) of an episode's LAST step, which
 * the in-kernel auto-reset would otherwise overwrite; only the rows of envs that finished in this step are written. */
typedef struct {
  uint8_t* info;
  void* terminal_obs;
  double* terminal_chaser;
  double* terminal_target;
} qs_step_extras_t;
/* As qs_step, plus the per-step info the reference returns.  extras may be NULL. */
int qs_step_ex(qs_env_t* env, const void* actions, void* obs_out, void* reward_out, uint8_t* done_out,
               const qs_step_extras_t* extras, void* stream);
/* qs_step for the env slice [first, first + n) only (first: a multiple of 128).  Pointers are the FULL arrays' bases, as in
 * qs_step.  Slices of one handle may be stepped concurrently on different streams (they share nothing); used by qs_rollout
 * to overlap the HBM-bound env kernel of one slice with the tensor-core policy kernel of another. */
int qs_step_range(qs_env_t* env, int64_t first, int64_t n, const void* actions, void* obs_out, void* reward_out,
                  uint8_t* done_out, void* stream);
/* n_steps consecutive qs_step calls in ONE launch, for open-loop stepping (BASELINE config 2: random actions; replaying an
 * action tape): actions[T,N,4] in, obs_out[T,N,obs_dim], reward_out[T,N], done_out[T,N] out (time-major, the handle's
 * dtype).  Each env's state stays in registers across the T steps -- no state traffic and no launch per step -- so small
 * batches are bound by the dependent-instruction chain of one step instead of launch latency.  Same arithmetic as T qs_step calls, auto-reset
 * included (a separate compilation of it: results agree to the last bit or two, within the same parity bars). */
int qs_step_n(qs_env_t* env, int64_t n_steps, const void* actions, void* obs_out, void* reward_out, uint8_t* done_out,
              void* stream);
/* Same step for HOST buffers (what a numpy VecEnv caller holds): copies actions in, steps, copies results out, chunked
 * over internal streams so H2D, kernel and D2H overlap.  The internal streams are forked from `stream` (an event recorded
 * there; work the caller enqueued on it before -- reset, import, a device-side step -- is seen) and the call BLOCKS until
 * the results are in host memory.  Buffers should come from qs_host_alloc (pinned). */
int qs_step_host(qs_env_t* env, const void* actions_host, void* obs_host, void* reward_host, uint8_t* done_host, void* stream);
int qs_host_alloc(void** out, uint64_t bytes);
int qs_host_free(void* p);

/* env.chaser.state / env.target.state / ... read and write access (parity injection, checkpoint/restore).
 * Counter ranges.  QS_F64 handles keep t and episode as 32-bit words.  QS_F32 docking handles pack both into ONE state
 * word: t in 10 bits (episodes end at t = max_steps = 600; with auto_reset = 0 a caller that keeps stepping a finished
 * env sees t saturate at 1023, done stays set) and episode in 22 bits (the Philox counter word: after 2^22 episodes of one
 * env its start-position sequence repeats).  qs_import_state on such a handle returns QS_ERR_INVALID, leaving the state
 * untouched, if any t > 1023 or episode outside [0, 2^22) -- it synchronises `stream` to find out.  Hover handles and
 * QS_F64 handles accept the full int32 range. */
int qs_export_state(qs_env_t* env, const qs_state_arrays_t* dst, void* stream);
int qs_import_state(qs_env_t* env, const qs_state_arrays_t* src, void* stream);

/* The states an observer needs from a few envs (the reference streams env.chaser.state / env.target.state to its Unity
 * client every step: server/pub_server.py:25-36, gym_docking/envs/video_docking_env.py): env_ids[k] (device, int64; ids
 * outside the handle are skipped) -> chaser_out[k,13], target_out[k,13] (docking; may be NULL) float64 world coordinates,
 * t_out[k] = env.t (may be NULL).  O(k) work: the batch is not exported. */
int qs_peek_state(qs_env_t* env, const int64_t* env_ids, int64_t k, double* chaser_out, double* target_out, int32_t* t_out,
                  void* stream);

/* Episode statistics accumulated on the device since the previous call:
 * [sum of episode returns, sum of episode lengths, episodes finished, docked env-steps, env-steps with a non-finite
 * observation or reward].  Synchronises `stream`.  The reference logs these through SB's Monitor/ep_info
 * (rl_baselines/ppo2/ppo2.py:494-497). */
int qs_get_stats(qs_env_t* env, double stats[5], void* stream);

/* Leaf functions of the reference on device arrays of n rows (dtype as given), for function-level parity tests and
 * for callers that use utils/transform.py, Drone or controller directly.  Unused pointers are NULL. */
typedef enum {
  QS_FN_QUAT2EULER = 0, /* utils/transform.py:94-120    in0 q[n,4]                     -> out0 euler[n,3] */
  QS_FN_EULER2QUAT = 1, /* utils/transform.py:123-136   in0 euler[n,3]                 -> out0 q[n,4] */
  QS_FN_QUAT2ROT = 2,   /* utils/transform.py:4-20      in0 q[n,4]                     -> out0 R[n,9] */
  QS_FN_ROT2EULER = 3,  /* utils/transform.py:23-46     in0 R[n,9]                     -> out0 euler[n,3] */
  QS_FN_U_LIMIT = 4,    /* quadrotor.py:171-187         in0 u[n,4]                     -> out0 u[n,4] */
  QS_FN_DRONE_STEP = 5, /* quadrotor.py:126-144         in0 state[n,13] in1 u_prev[n,4] in2 u_cmd[n,4] -> out0 state, out1 u_prev */
  QS_FN_PID = 6,        /* PIDController.py:179-185     in0 des[n,13] in1 now[n,13]    -> out0 u[n,4], out1 des_q[n,4] */
  QS_FN_VEL_CTRL = 7,   /* PIDController.py:106-141     in0 des[n,13] in1 now[n,13]    -> out0 u[n,4], out1 des_q[n,4] */
  QS_FN_STATE2REL = 8,  /* docking_env.py:257-295       in0 chaser[n,13] in1 target[n,13] -> out0 rel[n,12] */
  QS_FN_PHILOX = 9,     /* in0 uint32 ctr[n,4], in1 uint32 key[n,2] -> out0 uint32[n,4] (dtype ignored) */
  QS_FN_MATH = 10       /* the kernels' scalar math (numpy's arctan2/sin/cos/sqrt in the reference): in0 (y,x)[n,2] ->
                           out0 [n,6] = atan2(y,x), sin(y), cos(y), 1/sqrt(|x|), sqrt(|x|), asin(y) */
} qs_fn;
int qs_fn_apply(int fn, int dtype, int64_t n, const void* in0, const void* in1, const void* in2, void* out0,
                void* out1, void* stream);

/* Generalised advantage estimation over a device-resident rollout, the backward loop of Runner._run
 * (rl_baselines/ppo2/ppo2.py:511-521): for step = T-1..0
 *   nextnonterminal = 1 - (step == T-1 ? last_dones : dones[step+1]);  nextvalues = (step == T-1 ? last_values : values[step+1])
 *   delta = rewards[step] + gamma * nextvalues * nextnonterminal - values[step]
 *   advs[step] = lastgaelam = delta + gamma * lam * nextnonterminal * lastgaelam ;  returns = advs + values
 * All arrays are [T, N] time-major float32 device pointers (dones: uint8, dones[t] = the done flag BEFORE step t, as the
 * reference stores it); last_values/last_dones are [N].  gamma/lam are doubles like the reference's python floats.
 * One thread per env, coalesced along N; numpy's float32/float64 promotion in that loop is mirrored exactly. */
int qs_gae(int64_t n_steps, int64_t n_envs, const float* rewards, const float* values, const uint8_t* dones,
           const float* last_values, const uint8_t* last_dones, double gamma, double lam, float* advs_out,
           float* returns_out, void* stream);

/* ---- PID expert for the chaser (imitation-learning data) ----------------------------------------------------
 * The controller loop of run_expert_policy.py:44-69 and run_expert_record.py:40-139 (generate_PID_expert_traj) for every
 * env of a docking handle:  des_vel = 0.35 * (target_pos + [-0.2,0,0] - chaser_pos);  u = controller.vel_controller(
 * state_des, chaser_state, .) (PIDController.py:106-141);  action = (inv(rotor2control) @ u - action_mean) / action_std.
 * expert_mem: [N,8] in the handle's dtype, caller-owned, initialised to (0,0,0,0, 1,0,0,0) per env -- the slice of the
 * scripts' state_des that persists (des vel [3:6], pad, des quaternion [6:10]).
 * alias_ini_state != 0 reproduces the scripts literally: state_des IS env.chaser_ini_state, which env.reset() aliases as the
 * chaser's state (quadrotor.py:75), so every episode after the first starts from the mutated initial state and its first
 * command is plain hover thrust; on those first steps the chaser rows of the env state are re-seated and, if obs_inout
 * ([N,12], the observations returned by the last step/reset) is given, the reset observation is rewritten to match.
 * actions_out [N,4] is what env.step expects (already inside the Box for every state the expert reaches). */
int qs_expert_action(qs_env_t* env, void* expert_mem, int alias_ini_state, void* actions_out, void* obs_inout, void* stream);

/* ---- fused policy forward (tcgen05 / TMEM), the caller side of BASELINE config 5 ------------------------------------
 * model.step(obs) of the reference's PPO2 MlpPolicy (rl_baselines/ppo2/ppo2.py:472; architecture and hyper-parameters
 * run_moving_docking_ppo2.py:88-91: 12 -> 128 relu -> {128 relu -> 4 Gaussian mean, 128 relu -> 1 value}, state-independent
 * logstd).  Weights are float32 device pointers in torch.nn.Linear layout [out, in] (== the SB archive's [in, out] transposed). */
typedef struct qs_policy qs_policy_t;
typedef struct {
  const float *shared_w, *shared_b;  /* [128,12], [128]   model/shared_fc0 */
  const float *pi_fc_w, *pi_fc_b;    /* [128,128], [128]  model/pi_fc0 */
  const float *vf_fc_w, *vf_fc_b;    /* [128,128], [128]  model/vf_fc0 */
  const float *pi_w, *pi_b;          /* [4,128], [4]      model/pi */
  const float *vf_w, *vf_b;          /* [1,128], [1]      model/vf */
  const float *logstd;               /* [4]               model/pi/logstd */
} qs_policy_weights_t;
int qs_policy_create(qs_policy_t** out, int device);
int qs_policy_destroy(qs_policy_t* p);
/* Packs the weights into bf16 tensor-core operand images on the device; call again after every optimiser step. */
int qs_policy_set_weights(qs_policy_t* p, const qs_policy_weights_t* w, void* stream);
/* Arithmetic of the rollout forward (qs_policy_step, and qs_rollout with n_slices >= 1).  QS_POLICY_BF16 (default): bf16
 * weights and hidden activations, fp32 accumulation, on the tensor cores (inputs and hidden biases keep ~16 mantissa bits):
 * within 5e-2 of the fp32 policy on O(1) outputs, 69 us per 2^20 envs.  QS_POLICY_FP32: plain fp32 on the CUDA cores -- what
 * the reference's TensorFlow policy computes (run_moving_docking_ppo2.py:88-91) to within summation order (1e-4), ~25x
 * slower (1.76 ms per 2^20 envs, 58 % of the fp32 FMA peak); for evaluation and for checking the fast path.  The same (seed, env id, step) draws the same action noise in
 * either mode.  qs_policy_ppo_grad always runs the tensor-core formulation. */
#define QS_POLICY_BF16 0
#define QS_POLICY_FP32 1
int qs_policy_set_precision(qs_policy_t* p, int precision);
/* actions = mean + exp(logstd) * N(0,1) (noise: Philox4x32-10 keyed by seed, counter (env_id0 + i, step_idx, stream 2),
 * Box-Muller), values, neglogp for obs[N,12]; clipped_out (nullable) = actions clamped to the action Box [-1,1]
 * (ppo2.py:478-481).  deterministic != 0 returns the mean (model.predict(deterministic=True)). */
/* actions_out, clipped_out and neglogp_out may each be NULL (values only: model.value). */
int qs_policy_step(qs_policy_t* p, int64_t n, const float* obs, uint64_t seed, uint64_t env_id0, uint32_t step_idx,
                   int deterministic, float* actions_out, float* clipped_out, float* values_out, float* neglogp_out,
                   void* stream);


/* ---- PPO2 minibatch gradient: forward + loss + backward through the activations in one kernel, weight gradients in a second
 * The loss graph of the reference's PPO2 (rl_baselines/ppo2/ppo2.py:138-189: clipped surrogate on normalised advantages
 * (:265-266), clipped value loss, entropy bonus) for the minibatch rows idx[0..batch) of a flat rollout.  Kernel 1: per
 * 128-sample tile the policy forward, the per-sample loss and d loss / d (mean, value, logstd) in fp32, and the gradient pushed
 * back through the two hidden layers on the tensor cores; it leaves the bf16 operands of the weight-gradient contraction in
 * the caller's workspace as ready-made tensor-core operand images (opaque; sizes below) and adds the column sums of the
 * per-sample block to col_sums.  Kernel 2 (split-K over the minibatch, accumulators in tensor memory) reduces them to
 *     g1[128,32]  = dz1^T . a1        columns 0-11 + 16-27 = d W1 (the two halves of the hi/lo input split), column 12 = d b1
 *     g2[256,144] = dz2^T . [h|1 1 0] rows 0-127 = d Wp, rows 128-255 = d Wv, column 128 = d bias
 *     g3[256,16]  = [hp|hv]^T . dOut  rows 0-127, columns 0-3 = d W3p^T; rows 128-255, column 4 = d w3v
 * per_sample[i] (optional) = (dOut: d mean 0-3, d value, 0,0,0 | d logstd 0-3 | pg_loss_i, vf_loss_i, approxkl_i, clipfrac_i);
 * col_sums[16] = its column sums: d (b3p, b3v) = [0..4], d logstd = [8..11], the logged loss terms = [12..15].  The gradients
 * already carry the 1/batch of the mean and vf_coef; the entropy term (-ent_coef per logstd) is the caller's.
 * Workspace bytes for a minibatch of `batch` samples, T = ceil(batch / 128) tiles: a1 T*8192, h T*36864, hpv T*65536,
 * dz2 T*65536, dz1 T*32768, dout T*4096; each 16-byte aligned (any CUDA allocation is). */
typedef struct {
  const float *obs, *actions, *returns, *values, *neglogp; /* the rollout, flat: [M,12], [M,4], [M], [M], [M] */
  const int64_t* idx;      /* [batch] row numbers, or NULL for rows 0..batch-1 */
  int64_t batch;
  const float* adv_stats;  /* device [2]: mean and 1/(std + 1e-8) of (returns - values) over the minibatch (ppo2.py:265-266);
                            * NULL: computed here, from returns - values over idx (population std, sums in double, one launch) */
  float cliprange, cliprange_vf, vf_coef; /* cliprange_vf < 0: unclipped value loss */
  void *a1, *h, *hpv, *dz2, *dz1, *dout;  /* workspace (operand images, see above) */
  float* per_sample;                      /* [batch,16] out, or NULL */
  float* col_sums;                        /* [16] out */
  float *g1, *g2, *g3;                    /* [128,32], [256,144], [256,16] out */
  const void* packed_rows;                /* optional: the same rollout as ONE 128-byte record per row (qs_policy_pack_rollout);
                                           * the kernels then gather a sample with one cache line instead of touching five arrays
                                           * (1.1 GB -> 0.2 GB of DRAM reads per 1.7 M-sample minibatch).  NULL: the five arrays. */
} qs_ppo_grad_t;
int qs_policy_ppo_grad(qs_policy_t* p, const qs_ppo_grad_t* g, void* stream);
/* Interleaves a flat rollout into 128-byte rows for qs_ppo_grad_t.packed_rows: floats [0,12) obs, [12,16) action, 16 return,
 * 17 value, 18 neglogp, rest zero.  packed_out: rows * 128 bytes, 16-byte aligned.  Once per update (ppo2.py:340-362 runs
 * noptepochs * nminibatches minibatch steps over the same rollout); the arrays stay the source of truth. */
int qs_policy_pack_rollout(const float* obs, const float* actions, const float* returns, const float* values,
                           const float* neglogp, int64_t rows, void* packed_out, void* stream);

/* The optimiser step of the update in one launch: tf.clip_by_global_norm(grads, max_grad_norm) followed by
 * tf.train.AdamOptimizer(learning_rate, epsilon=1e-5).apply_gradients (rl_baselines/ppo2/ppo2.py:200-207), TensorFlow's
 * formulation: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t); m += (1 - b1)(g - m); v += (1 - b2)(g^2 - v); p -= lr_t m / (sqrt(v) + eps).
 * Up to 16 float32 device tensors; `step` counts from 1 -- or, with a device-resident int64 `step_counter` (nullable), the
 * kernel increments that counter and derives lr_t from it (what a captured CUDA graph needs: nothing step-dependent is baked
 * into the launch); max_grad_norm <= 0 disables clipping; grad_norm_out (device [1], nullable) receives the unclipped norm. */
typedef struct {
  float* param;
  const float* grad;
  float *m, *v;
  int64_t n;
} qs_adam_tensor_t;
int qs_adam_step(const qs_adam_tensor_t* tensors, int n_tensors, float lr, float beta1, float beta2, float eps, int64_t step,
                 int64_t* step_counter, float max_grad_norm, float* grad_norm_out, void* stream);

/* The minibatch shuffle of the update (np.random.shuffle(inds), rl_baselines/ppo2/ppo2.py:350-352): out[0..n) = a permutation
 * of 0..n-1 determined by `seed` (a keyed Feistel bijection with cycle walking: no sort, every entry independent). */
int qs_permutation(int64_t n, uint64_t seed, int64_t* out, void* stream);

/* ---- device-resident rollout: the step loop of Runner._run (rl_baselines/ppo2/ppo2.py:471-510) ------------------------
 * for t in 0..T-1:  actions, values, neglogp = model.step(obs[t]);  obs[t+1], rewards[t], dones[t+1] = env.step(clip(actions))
 * then values[T] = model.value(obs[T]).  Every kernel reads and writes the rollout storage in place (no staging copies,
 * no host round trip: two launches per step).  The caller seeds obs[0] / dones[0] (env.reset() the first time, the previous
 * rollout's row T afterwards) and runs qs_gae on the result.  float32 docking environments only. */
typedef struct {
  float* obs;       /* [T+1, N, 12]  row 0 in, rows 1..T out */
  float* actions;   /* [T, N, 4]     unclipped samples (what the PPO loss consumes) */
  float* values;    /* [T+1, N]      row T = bootstrap value of obs[T] */
  float* neglogp;   /* [T, N] */
  float* rewards;   /* [T, N] */
  uint8_t* dones;   /* [T+1, N]      row 0 in; row t+1 = done flag returned by step t (== "done before step t+1") */
  float* clipped;   /* [N, 4]        scratch: the clipped actions handed to the env */
} qs_rollout_t;
/* Noise for step t uses Philox counter word step_idx0 + t (see qs_policy_step).  n_slices = 0: the FUSED step kernel -- one
 * launch per step does model.step and env.step: the policy half of every CTA hands a tile's clipped actions to env warps of
 * the same CTA through shared memory, and those run the env step kernel's own device code while the tensor pipe works on the
 * next tile (float32 docking handles with the default model parameters; `clipped` is not used).  n_slices = 1: two launches
 * per step, everything in order on `stream`.  n_slices = 2..8: the envs are cut into that many slices (whole 128-env tiles) that run their step loops on
 * internal streams (forked from and joined back into `stream`), so one slice's env kernel (HBM-bound) overlaps another's
 * policy kernel (tensor cores); results are bit-identical to n_slices = 1. */
int qs_rollout(qs_env_t* env, qs_policy_t* policy, int64_t n_steps, const qs_rollout_t* buffers, uint64_t seed,
               uint32_t step_idx0, int n_slices, void* stream);

#ifdef __cplusplus
}
#endif
#endif
