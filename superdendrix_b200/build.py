"""Builds superdendrix_b200/libsdx.so (the C-ABI library of include/sdx.h) in-tree.

    python -m superdendrix_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; one object per translation unit
(compiled in parallel), then one shared library with the CUDA runtime linked
statically so the .so travels to the GPU box as is.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libsdx.so")
UNITS = ["sdx_api", "sdx_score", "sdx_curveball", "sdx_lanes", "sdx_scan", "sdx_stats"]
HOST_UNITS = ["sdx_io"]          # plain C++ (the host-side format path), compiled by the host compiler through nvcc
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wno-format-truncation", "-Xptxas", "-v"]
if os.environ.get("SDX_BUILD_DEFINES"):          # debugging only, e.g. SDX_BUILD_DEFINES="-DSDX_SM_DEBUG"
    NVCC_FLAGS += os.environ["SDX_BUILD_DEFINES"].split()


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsdx.so cannot be built (there is no CPU fallback)")


def _sources_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "sdx.h"), __file__]
    return max(os.path.getmtime(p) for p in paths)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _sources_mtime():
        return SO
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)

    def compile_unit(u):
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, u + ".cu"), "-o", os.path.join(OBJ, u + ".o")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(OBJ, u + ".ptxas.log"), "w") as fh:
            fh.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s.cu:\n%s" % (u, r.stdout + r.stderr))
        return r.stderr

    def compile_host_unit(u):
        cmd = [nvcc, "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-pthread", "-c", os.path.join(CSRC, u + ".cpp"), "-o", os.path.join(OBJ, u + ".o")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host compile failed on %s.cpp:\n%s" % (u, r.stdout + r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=len(UNITS) + len(HOST_UNITS)) as ex:
        host_logs = [ex.submit(compile_host_unit, u) for u in HOST_UNITS]
        logs = list(ex.map(compile_unit, UNITS))
        logs += [h.result() for h in host_logs]
    if verbose:
        sys.stderr.write("".join(logs))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", SO,
           *[os.path.join(OBJ, u + ".o") for u in UNITS + HOST_UNITS], "-Xcompiler", "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link of libsdx.so failed:\n" + r.stdout + r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
