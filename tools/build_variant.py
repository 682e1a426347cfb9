#!/usr/bin/env python3
"""Development A/B: builds yulio_raytracer_b200/lib/variants/libyrt_<name>.so with extra nvcc flags (e.g. -DYRT_SHADE_MINBLOCKS=8).
    python tools/build_variant.py <name> [nvcc flags...]"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
name, extra = sys.argv[1], sys.argv[2:]
obj = os.path.join(g.REPO, "build", "variants", name); out = os.path.join(g.LIBDIR, "variants")
os.makedirs(obj, exist_ok=True); os.makedirs(out, exist_ok=True)
def cc(src):
    o = os.path.join(obj, src + ".o")
    r = subprocess.run(["nvcc"] + g.NVCC_FLAGS + extra + ["-c", os.path.join(g.CSRC, src), "-o", o], capture_output=True, text=True)
    if r.returncode: sys.stderr.write(r.stderr); raise SystemExit(1)
    if src == "kernels.cu":
        for l in r.stderr.splitlines():
            if "k_shade" in l and "Function properties" in l: print(l.strip())
            if "Used" in l and prev and "k_shade" in prev: print(l.strip())
            if "spill" in l and prev2 and "k_shade" in prev2: print(l.strip())
            prev2 = globals().get("prev"); globals()["prev2"] = prev2; globals()["prev"] = l
    return o
prev = prev2 = None
with ThreadPoolExecutor(6) as ex: objs = list(ex.map(cc, g.CU_SOURCES))
lib = os.path.join(out, f"libyrt_{name}.so")
subprocess.check_call(["nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lz", "-ldl"])
print("built", lib)
