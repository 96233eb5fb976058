#!/usr/bin/env python
"""Timeline of the rollout forward kernel (policy_fwd_ws_kernel) from SM-clock stamps: needs a library built with
-DQS_WS_TRACE=1 (QUADSIM_B200_LIB=variants/ws_trace.so).  Prints, for CTA 0 / group 0, the mean clocks each phase of a tile
takes on the epilogue thread and on the MMA-issuer lane, and the per-tile period.

    QUADSIM_B200_LIB=variants/ws_trace.so python tools/ws_trace.py [--n 1048576]
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from quadsim_b200 import _capi, ppo
from quadsim_b200.fused_policy import FusedPolicy

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
args = ap.parse_args()
lib = _capi.load()
pol = ppo.MlpPolicy().cuda()
fused = FusedPolicy(pol)
obs = torch.randn(args.n, 12, device="cuda")
for _ in range(3):
    fused.step(obs)
torch.cuda.synchronize()
TILES, SLOTS = 28, 24
buf = np.zeros(TILES * SLOTS, np.int64)
lib.qs_debug_ws_trace.restype = C.c_int
lib.qs_debug_ws_trace.argtypes = [C.c_void_p, C.c_int]
got = lib.qs_debug_ws_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
assert got == buf.size, got
t = buf.reshape(TILES, SLOTS).astype(np.float64)
tiles = min(TILES, (args.n // 128) // (148 * 2))
t = t[2:tiles - 1]  # steady state
EPI = ["loop top", "Done1 seen (layer 1 complete)", "h read out, packed, stored to A2", "proxy fence", "previous D3 read", "FullA2 signalled",
       "previous tile finished (sampling maths, stores)", "DoneP seen (layer 2 pi half)", "pi half packed into TMEM", "FullA3p signalled",
       "DoneV seen (layer 2 vf half)", "vf half packed into TMEM", "FullA3v signalled", "next A1 built + FullA1", "noise drawn (loop bottom)"]
ISS = {16: "FullA1 seen", 17: "layer 1 issued", 18: "FullA2 seen", 19: "layer 2 issued", 20: "FullA3p seen", 21: "layer 3 pi issued",
       22: "FullA3v seen", 23: "layer 3 vf issued"}
out = {"period_clk": float(np.mean(np.diff(buf.reshape(TILES, SLOTS)[2:tiles - 1, 0]))), "epilogue": [], "issuer": [], "cross": {}}
for k in range(1, 15):
    out["epilogue"].append((EPI[k], round(float(np.mean(t[:, k] - t[:, k - 1])), 1)))
for k in range(16, 24):
    prev = t[:, k - 1] if k > 16 else None
    out["issuer"].append((ISS[k], None if prev is None else round(float(np.mean(t[:, k] - prev)), 1)))
# hand-off latencies between the two roles (same SM clock)
X = out["cross"]
X["FullA1 arrive (prev loop 13) -> issuer sees it (16)"] = round(float(np.mean(t[1:, 16] - t[:-1, 13])), 1)
X["layer 1 issued (17) -> epilogue sees Done1 (1)"] = round(float(np.mean(t[:, 1] - t[:, 17])), 1)
X["FullA2 arrive (5) -> issuer sees it (18)"] = round(float(np.mean(t[:, 18] - t[:, 5])), 1)
X["layer 2 issued (19) -> epilogue sees DoneP (7)"] = round(float(np.mean(t[:, 7] - t[:, 19])), 1)
X["layer 2 issued (19) -> epilogue sees DoneV (10)"] = round(float(np.mean(t[:, 10] - t[:, 19])), 1)
X["FullA3p arrive (9) -> issuer sees it (20)"] = round(float(np.mean(t[:, 20] - t[:, 9])), 1)
X["FullA3v arrive (12) -> issuer sees it (22)"] = round(float(np.mean(t[:, 22] - t[:, 12])), 1)
X["layer 3 vf issued (23) -> next loop's D3 read done (4)"] = round(float(np.mean(t[1:, 4] - t[:-1, 23])), 1)
print(json.dumps(out, indent=1))
