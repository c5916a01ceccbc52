"""Builds libnalgebra_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m nalgebra_b200.build [--force]

No torch involvement: the product is a plain C-ABI shared library (include/nalgebra_b200.h).
nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# A build with extra flags (NAB_EXTRA_NVCC_FLAGS="-DNAB_GETF2_PROF -DNAB_DEBUG_HOOKS": per-phase cycle counters and the
# debug exports the tools/ scripts use) goes to its own object directory and library; select it with
# NAB_LIB=nalgebra_b200/libnalgebra_b200_dbg.so.  The product library never carries those symbols.
_DBG = bool(os.environ.get("NAB_EXTRA_NVCC_FLAGS", "").strip())
OBJ_DIR = os.path.join(HERE, "build_dbg" if _DBG else "build")
LIB_PATH = os.path.join(HERE, "libnalgebra_b200_dbg.so" if _DBG else "libnalgebra_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
] + os.environ.get("NAB_EXTRA_NVCC_FLAGS", "").split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "nalgebra_b200.h"))
    return hs


def _compile(src: str, force: bool) -> str:
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    newest_dep = max(os.path.getmtime(p) for p in [src] + _headers())
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest_dep:
        return obj
    cmd = [NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed on {src}")
    return obj


def build(force: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), srcs))
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(o) > os.path.getmtime(LIB_PATH) for o in objs):
        cmd = [NVCC, "-shared", "-cudart", "static", "-o", LIB_PATH] + objs + ["-lpthread", "-ldl", "-lrt"]
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
