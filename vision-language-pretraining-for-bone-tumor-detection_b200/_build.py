"""In-tree build of the CUDA extension (``csrc/libvlpclip.so``) for sm_100a.

``nvcc`` cross-compiles without a GPU; the resulting shared object travels with the tree
(git-ignored, not gpurun-ignored).  No JIT cache, no torch cpp_extension: the library is a plain
C-ABI ``.so`` (``include/vlpclip.h``) loaded with ctypes.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libvlpclip.so")
# VLP_B200_LIB=/path/to/other.so loads that library instead (same C ABI) and never rebuilds it: lets the
# whole GPU test-suite and bench.py run against an experimental build (e.g. one of tools/diag_build.sh)
# before it is promoted.  Unset (the default) = the shipped library built from csrc/*.cu.
_LIB_OVERRIDE = os.environ.get("VLP_B200_LIB") or None
if _LIB_OVERRIDE:
    LIB_PATH = os.path.abspath(_LIB_OVERRIDE)
SOURCES = ["lse_fwd.cu", "grad_bwd.cu", "prologue.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC",
    "-diag-suppress", "177",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 extension cannot be built (no fallback path)")


def sources() -> list:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if _LIB_OVERRIDE:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"VLP_B200_LIB={LIB_PATH} does not exist")
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "vlpclip.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every kernel for sm_100a into one shared library; returns its path."""
    if _LIB_OVERRIDE:          # a library chosen with VLP_B200_LIB is used as it is, never rebuilt
        needs_build()
        return LIB_PATH
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
