"""Builds libplssvm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("PLSSVM_B200_LIB") or os.path.join(HERE, "libplssvm_b200.so")  # (the override is for A/B measurements of two builds)
SOURCES = ["backend.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))) + [os.path.join("..", "..", "include", "plssvm_b200.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """PLSSVM_B200_EXPERIMENTAL=1 in the environment adds -DPLSSVM_B200_EXPERIMENTAL: the measured-but-not-faster tile-kernel variants
    (impl 4 / 5 / 8 / 9) are compiled in as well.  The default library does not contain them."""
    if not force and not needs_build():
        return LIB
    experimental = ["-DPLSSVM_B200_EXPERIMENTAL"] if os.environ.get("PLSSVM_B200_EXPERIMENTAL", "0") not in ("", "0") else []
    experimental += os.environ.get("PLSSVM_B200_DEFINES", "").split()  # extra -D flags (A/B builds)
    cmd = [_nvcc(), *NVCC_FLAGS, *experimental, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
    env = dict(os.environ)
    env.pop("CXX", None)  # the image's CXX points at a compiler wrapper without OpenMP specs; nvcc should use the distro g++
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libplssvm_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
