"""Build libh9gpu variants with extra -D flags for A/B timing: variants/libh9gpu_<name>.so
(select with H9GPU_LIB).  usage: build_variant.py name [-DFLAG ...]"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybrid9_b200 import build as b

def main(name, flags):
    out = os.path.join(ROOT, "variants"); obj = os.path.join(out, name); os.makedirs(obj, exist_ok=True)
    objs = []
    def one(item):
        src, fl = item
        o = os.path.join(obj, src.replace(".cu", ".o"))
        r = subprocess.run([b._nvcc()] + b.ARCH + b.COMMON + fl + flags + ["-c", os.path.join(b.CSRC, src), "-o", o], capture_output=True, text=True)
        if r.returncode: raise RuntimeError(r.stderr[-3000:])
        if src == "h9_kernels_fast.cu":
            lines = r.stderr.split("\n")
            for i, l in enumerate(lines):
                if "days_kernel_fastILi64ELi8" in l and "Function properties" in l:
                    print(name, lines[i + 1].strip(), lines[i + 2].strip())
        return o
    with ThreadPoolExecutor(4) as ex: objs = list(ex.map(one, b.UNITS.items()))
    lib = os.path.join(out, f"libh9gpu_{name}.so")
    subprocess.run([b._nvcc()] + b.ARCH + ["-shared", "-o", lib] + objs + ["-cudart", "static"], check=True)
    print(lib)
main(sys.argv[1], sys.argv[2:])
