"""In-tree build of libshc_b200.so (hand-written CUDA for sm_100a + the C-ABI of include/shc_b200.h)."""
from __future__ import annotations

import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libshc_b200.so")
SOURCES = ["shc_engine.cu"]
HEADERS = ["shc_math.cuh", "shc_consts.h", "shc_layout.h", "shc_cycle.cuh", "shc_host.cuh", "shc_pack.cuh", "shc_msgs.cuh", "shc_sequence.cuh",
           "shc_startup.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    inc = os.path.join(os.path.dirname(PKG_DIR), "include")
    deps += [os.path.join(inc, f) for f in ("shc_b200.h", "shc_config.h", "shc_state.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile with nvcc (cross-compiles without a GPU).  Returns the library path."""
    if force or _stale():
        nvcc = os.environ.get("NVCC", "nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + \
              [os.path.join(CSRC, s) for s in SOURCES]
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
