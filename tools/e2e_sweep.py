import os, sys, time, json
sys.path.insert(0, os.getcwd())
import torch, quadsim_b200 as qs
n = 1 << 20
env = qs.make_vec("docking-v2", n, seed=0, randomize="episode", want_info=False)
h_acts, h_obs, h_rew, h_done = env.pinned_host_buffers(n_action_buffers=4)
import numpy as np
for a in h_acts: a[...] = np.random.uniform(-1, 1, a.shape).astype(np.float32)
env.reset()
for k in range(5): env.step_host(h_acts[k % 4], out=(h_obs, h_rew, h_done))
t0 = time.perf_counter()
for k in range(60): env.step_host(h_acts[k % 4], out=(h_obs, h_rew, h_done))
el = time.perf_counter() - t0
print(os.environ.get("QUADSIM_B200_HOST_CHUNK", "default"), "%.3g env-steps/s" % (n * 60 / el), "%.3f ms/step" % (el / 60 * 1e3))
