"""Build experiment variants of the library (defines on the nvcc command line) into pose_refine_b200/variants/.
    python scripts/build_variants.py name1:DEF=1,DEF2=3 name2:...   (timed on the GPU by scripts/gpu_variants.sh)"""
import os, sys, subprocess, re
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib
import pose_refine_b200
bm = importlib.import_module('pose_refine_b200.build')
bm = sys.modules['pose_refine_b200.build']
VDIR = os.environ.get("VDIR", os.path.join(ROOT, "pose_refine_b200", "variants"))
os.makedirs(VDIR, exist_ok=True)
for f in ([] if os.environ.get("KEEP") else os.listdir(VDIR)):
    if f.endswith(".so"): os.remove(os.path.join(VDIR, f))
def one(spec):
    name, _, defs = spec.partition(":")
    defines = [d for d in defs.split(",") if d]
    out = os.path.join(VDIR, f"lib_{name}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + bm.NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-ccbin", "/usr/bin/g++", "-o", out] + [os.path.join(bm.CSRC, s) for s in bm.SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode: return f"{name}: FAILED\n{r.stderr[-2000:]}"
    m = re.search(r"icp_persistent_kernelINS_11PackedScene.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", r.stderr, re.S)
    return f"{name}: regs={m.group(3)} spill={m.group(2)} stack={m.group(1)}" if m else f"{name}: built"
with ThreadPoolExecutor(4) as ex:
    for line in ex.map(one, sys.argv[1:]): print(line)
