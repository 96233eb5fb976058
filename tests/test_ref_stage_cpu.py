"""oracle/_ref (the UNMODIFIED reference files bench.py times on the GPU box, staged by oracle/make_ref.py) and the
`gym_docking` drop-in package that gives the reference's own import paths / env ids a CUDA-backed meaning."""
import json
import os
import subprocess
import sys

import pytest

from qs_testutil import ROOT

sys.path.insert(0, ROOT)
from oracle import make_ref  # noqa: E402


@pytest.mark.skipif(not os.path.isdir("/root/reference/dynamics"), reason="the reference tree exists in the build container only")
def test_staged_reference_is_byte_identical_to_the_source_tree():
    man = make_ref.stage(verbose=False)
    assert man is not None and make_ref.staged()
    assert sorted(man["files"]) == sorted(make_ref.FILES) and len(make_ref.FILES) == 6  # exactly the hot-path files
    for rel, digest in man["files"].items():
        assert make_ref.sha256(os.path.join("/root/reference", rel)) == digest == make_ref.sha256(os.path.join(make_ref.REF_DIR, rel)), rel
    # nothing of the staged tree is tracked by git (the reference's sources never enter this repository's history)
    tracked = subprocess.check_output(["git", "ls-files", "oracle/_ref"], cwd=ROOT).decode().strip()
    assert tracked == ""
    ign = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in ign
    gpi = os.path.join(ROOT, ".gpurunignore")
    assert not os.path.exists(gpi) or "oracle/_ref" not in open(gpi).read()  # it must travel to the GPU box


@pytest.mark.skipif(not make_ref.staged(), reason="oracle/_ref not staged")
def test_reference_worker_and_config1_run_from_the_staged_tree():
    """One SubprocVecEnv-style worker (env.step + reset-on-done) and BASELINE config 1 (run_sim_PID.py:43-54) on the staged,
    unmodified reference, in fresh interpreters exactly as bench.py's cpu_baseline leg starts them."""
    from oracle import ref_bench
    r = ref_bench.run_workers("docking-v2", 0.5, 2)
    assert r["procs"] == 2 and r["failed"] == 0 and r["steps"] > 50 and 50 < r["per_core"] < 5000
    c1 = ref_bench.run_config1()
    assert c1["steps"] == 2000 and abs(c1["final_pos"][0] + 0.2) < 1e-6 and abs(c1["final_pos"][1] - 0.2) < 1e-6 and abs(c1["final_pos"][2] - 0.2) < 1e-6


def test_gym_docking_drop_in_package_resolves_the_reference_names():
    """`import gym_docking`, `gym_docking.envs.<Class>`, `from gym_docking.envs.docking_env import DockingEnv, state2rel` -- the
    reference's import paths (gym-docking/gym_docking/__init__.py:3-21, envs/__init__.py) -- resolve to the CUDA-backed facades.
    Run in a fresh interpreter: the oracle's import shim puts the reference's own gym_docking first in this process."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import gym_docking, gym_docking.envs as E\n"
            "from gym_docking.envs.docking_env import DockingEnv, state2rel\n"
            "from gym_docking.envs.moving_docking_env import MovingDockingEnv\n"
            "from gym_docking.envs.imitating_docking_env import ImitatingDockingEnv\n"
            "from gym_docking.envs.hovering_env import HoveringEnv\n"
            "import quadsim_b200.envs as Q\n"
            "assert E.DockingEnv is DockingEnv is Q.DockingEnv and E.MovingDockingEnv is MovingDockingEnv is Q.MovingDockingEnv\n"
            "assert E.ImitatingDockingEnv is ImitatingDockingEnv and E.HoveringEnv is HoveringEnv\n"
            "assert sorted(gym_docking.ENV_IDS) == ['docking-v0', 'docking-v1', 'docking-v2', 'hovering-v0']\n"
            "assert MovingDockingEnv.env_id == 'docking-v2' and callable(state2rel) and callable(gym_docking.make)\n"
            "print('ok')\n") % ROOT
    out = subprocess.check_output([sys.executable, "-c", code], cwd="/tmp").decode()
    assert out.strip().endswith("ok")
