"""Build libquadsim_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m quadsim_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libquadsim_b200.so")
SOURCES = [os.path.join(CSRC, "qs_lib.cu"), os.path.join(CSRC, "qs_policy.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "qs_core.cuh"), os.path.join(CSRC, "qs_env_kernels.cuh"), os.path.join(ROOT, "include", "quadsim_b200.h"), os.path.abspath(__file__)]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr", "-Xptxas", "-v",
    # fp32 only (fp64 is unaffected): flush denormals, 2-ulp division and sqrt -- measured +13 % on the fp32 step kernel,
    # identical episode statistics; sinf/cosf/atan2f/asinf stay the accurate library versions (no --use_fast_math)
    "-ftz=true", "-prec-div=false", "-prec-sqrt=false",
    # 6 resident CTAs/SM for the fp32 step kernels (<= 80 registers): measured best of {4,5,6,8}
    "-DQS_MINB32=6",
    # fp64 step kernels: 3 CTAs/SM (168 registers, no spills) measured best of {3,4,5,6} with the default (shortcut) fp64
    # arithmetic: 187.7 / 196.4 / 228.7 / 254.6 us per step; the literal build (-DQS_LITERAL64) preferred 4: 293 us
    "-DQS_MINB64=3",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: quadsim_b200 needs the CUDA toolkit to build its library")


def up_to_date():
    return os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile the CUDA library if it is missing or older than its sources; returns its path."""
    if out is None and not force and up_to_date():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    target = out or LIB_PATH
    # g++ from the system: the image exports CC/CXX pointing at a toolchain nvcc does not need
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-I", os.path.join(ROOT, "include"), "-o", target] + SOURCES
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(LIB_DIR, "build.log" if out is None else os.path.basename(out) + ".log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (exit %d); see %s" % (res.returncode, log))
    return target


# ---------------------------------------------------------------------------------------------------------------------------
# The PyTorch-extension face (csrc/qs_torch_ext.cpp): host-only C++ (pybind11 + ATen), links libquadsim_b200.so by rpath.
EXT_NAME = "_qs_torch_ext"
EXT_SRC = os.path.join(CSRC, "qs_torch_ext.cpp")


def ext_path():
    import sysconfig
    return os.path.join(LIB_DIR, EXT_NAME + sysconfig.get_config_var("EXT_SUFFIX"))


def ext_up_to_date():
    p = ext_path()
    deps = [EXT_SRC, os.path.join(ROOT, "include", "quadsim_b200.h")]
    return os.path.exists(p) and all(os.path.getmtime(p) >= os.path.getmtime(d) for d in deps)


def build_ext(force=False, verbose=False):
    """g++ -> quadsim_b200/_lib/_qs_torch_ext.<abi>.so (in-tree, travels to the GPU box like the CUDA library)."""
    if not force and ext_up_to_date():
        return ext_path()
    import sysconfig
    import torch
    from torch.utils import cpp_extension as E
    build(force=False)  # the extension links the CUDA library
    cuda_home = os.environ.get("CUDA_HOME") or "/usr/local/cuda"
    inc = E.include_paths() + [sysconfig.get_paths()["include"], os.path.join(cuda_home, "include"), os.path.join(ROOT, "include")]
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-DTORCH_EXTENSION_NAME=" + EXT_NAME,
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    for i in inc:
        cmd += ["-isystem", i]
    cmd += [EXT_SRC, "-o", ext_path(), "-L" + LIB_DIR, "-lquadsim_b200", "-Wl,-rpath,$ORIGIN"]
    for lp in E.library_paths():
        cmd += ["-L" + lp, "-Wl,-rpath," + lp]
    cmd += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(os.path.join(LIB_DIR, "build_ext.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building the PyTorch extension failed (exit %d); see %s" % (res.returncode, os.path.join(LIB_DIR, "build_ext.log")))
    return ext_path()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_ext(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
