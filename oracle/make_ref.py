"""TEST INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference sources of the hot path under oracle/_ref/ so that the
reference itself can be timed on the GPU box's host cores (bench.py's cpu_baseline leg, kind "reference").

    python oracle/make_ref.py            # copies from /root/reference (or $QUADSIM_REFERENCE_ROOT); no-op when absent

The reference is pure Python; /root/reference exists only in the build container.  oracle/_ref/ is git-ignored (the sources
never enter this repository's history) but NOT gpurun-ignored, so the staged files travel to the GPU box the same way the
built .so files do.  Exactly the files SURVEY.md section 8(a) cites are staged, byte for byte, in the reference's own
directory layout.  The reference's package __init__ files are deliberately NOT staged (controller/__init__.py imports the
joystick thread, gym_docking/envs/__init__.py the screen-grab env); EMPTY ones are generated in their place so that the
staged packages are regular packages and win over this repository's own `gym_docking` drop-in wherever both are importable.
A manifest with each file's sha256 is written next to them and checked against the source tree by the CPU tests.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SRC_ROOT = os.environ.get("QUADSIM_REFERENCE_SRC", "/root/reference")

# the functions on the hot path live in these six files
FILES = [
    "dynamics/quadrotor.py",
    "controller/PIDController.py",
    "utils/transform.py",
    "gym-docking/gym_docking/envs/docking_env.py",
    "gym-docking/gym_docking/envs/moving_docking_env.py",
    "gym-docking/gym_docking/envs/hovering_env.py",
]


def sha256(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def staged():
    """True when every file of the manifest is present under oracle/_ref/."""
    man = os.path.join(REF_DIR, "MANIFEST.json")
    if not os.path.exists(man):
        return False
    try:
        files = json.load(open(man))["files"]
    except Exception:
        return False
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in files) and set(files) == set(FILES)


def stage(src_root=SRC_ROOT, verbose=True):
    """Copy the hot-path files; returns the manifest dict, or None when the reference tree is not here."""
    if not os.path.isdir(os.path.join(src_root, "dynamics")):
        if verbose:
            print("make_ref: no reference tree at %s (fine on the GPU box: oracle/_ref/ arrives pre-staged)" % src_root)
        return None
    manifest = {"source": src_root, "files": {}}
    for rel in FILES:
        dst = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, rel), dst)
        manifest["files"][rel] = sha256(dst)
    for pkg in ("dynamics", "controller", "utils", "gym-docking/gym_docking", "gym-docking/gym_docking/envs"):
        open(os.path.join(REF_DIR, pkg, "__init__.py"), "w").close()  # generated, empty: not a copy of the reference's
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print("make_ref: staged %d reference files under %s" % (len(FILES), REF_DIR))
    return manifest


if __name__ == "__main__":
    sys.exit(0 if (stage() is not None or staged()) else 0)
