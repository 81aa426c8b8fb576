"""Build the CUDA shared library in-tree (``mocodad_b200/libmocodad_b200.so``) for sm_100a.

nvcc cross-compiles without a GPU, so this also runs on the CPU-only build machine
(``__graft_entry__.build()``).  The library links the CUDA runtime statically and has no
torch / Python dependency: the C ABI in ``include/mocodad_b200.h`` is all it exports.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_PATH = os.path.join(PKG, "libmocodad_b200.so")
STAMP = LIB_PATH + ".stamp"
SOURCES = ("mcd_api.cu",)
HEADERS = ("mcd_kernels.cuh", "mcd_block_tc.cuh", "mcd_block_cf.cuh", "mcd_edge_blocks.cuh", "mcd_latent.cuh",
           os.path.join(ROOT, "include", "mocodad_b200.h"))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the mocodad_b200 CUDA library cannot be built (no fallback exists)")


def _digest() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        p = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_current() -> bool:
    try:
        return os.path.exists(LIB_PATH) and open(STAMP).read().strip() == _digest()
    except OSError:
        return False


TRACE_LIB_PATH = os.path.join(PKG, "libmocodad_b200_trace.so")


def build_trace_variant() -> str:
    """Debug build with the device-timeline instrumentation of the tensor-core block kernel compiled in
    (``-DMCD_TC_TRACE=1``); used by tools/trace_block.py through ``MOCODAD_B200_LIB``.  Never loaded by default."""
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DMCD_TC_TRACE=1", "-o", TRACE_LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return TRACE_LIB_PATH


def build_variant(tag: str, defines) -> str:
    """Experiment build ``libmocodad_b200_<tag>.so`` with extra ``-D`` flags (A/B runs on one GPU box through
    ``MOCODAD_B200_LIB``).  Never loaded by default."""
    out = os.path.join(PKG, f"libmocodad_b200_{tag}.so")
    cmd = [_nvcc()] + NVCC_FLAGS + ["-Xptxas", "-v"] + [f"-D{d}" for d in defines] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(out + ".ptxas.log", "w") as fh:
        fh.write(res.stderr)
    return out


def build_extension(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if sources changed (or ``force``).  Returns the .so path."""
    if not force and is_current():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    import sys
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    elif "--trace" in sys.argv:
        print(build_trace_variant())
    else:
        print(build_extension(force="--force" in sys.argv, verbose="-v" in sys.argv))
