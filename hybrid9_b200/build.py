"""Build recipe for libh9gpu.so (sm_100a only) -- explicit nvcc, in-tree output.

Three translation units with different floating-point contracts:
  h9_kernels_exact.cu  -fmad=false -prec-div=true   (H9_MATH_EXACT: the reference's op order)
  h9_kernels_fast.cu   explicit FMAs + MUFU math, -fmad=false so that all launch
                       variants of one source give the same bits (H9_MATH_FAST)
  h9_pack.cu, h9_api.cu                             (ingest, budget, C-ABI)
cudart is linked statically; the library has no torch / python dependency.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libh9gpu.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
UNITS = {
    "h9_kernels_exact.cu": ["-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"],
    "h9_kernels_fast.cu": ["-fmad=false"],  # FMAs are explicit: every launch variant gives the same bits
    "h9_kernels_pair.cu": ["-fmad=false"],  # two lanes per cell for small shards (H9_MATH_FAST)
    "h9_pack.cu": [],
    "h9_api.cu": [],
}
HEADERS = ["h9_physics.h", "h9_exact_tables.h", "h9_physics_fast.cuh", "h9_physics_fast_tp.cuh", "h9_physics_pair.cuh", "h9_device.h", "h9_kernels.cuh"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [
        os.path.join(HERE, "..", "include", "h9gpu.h"), os.path.abspath(__file__)]
    jobs = []
    objs = []
    for src, flags in UNITS.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append(([_nvcc()] + ARCH + COMMON + flags + ["-c", s, "-o", o], src))

    def run(job):
        cmd, name = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(objdir, name + ".ptxas.log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{r.stderr[-4000:]}")
        if verbose:
            sys.stderr.write(r.stderr)
        return name

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(OUT, objs):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
