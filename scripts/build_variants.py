"""Build experiment variants of the library (defines on the nvcc command line for icp.cu / raster.cu) into
pose_refine_b200/variants/ (git-ignored; travels to the GPU box).
    python scripts/build_variants.py name1:DEF=1,DEF2=3 name2:...   (timed on the GPU by scripts/gpu_variants.sh)
Every variant is compiled with -DPR_DEBUG, which enables the PR_HYP_CLUSTER environment knob of icp.cu."""
import os, sys, re, importlib
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
importlib.import_module("pose_refine_b200.build")
bm = sys.modules["pose_refine_b200.build"]
VDIR = os.environ.get("VDIR", os.path.join(ROOT, "pose_refine_b200", "variants"))
os.makedirs(VDIR, exist_ok=True)
for f in ([] if os.environ.get("KEEP") else os.listdir(VDIR)):
    if f.endswith(".so") or f.endswith(".stamp"): os.remove(os.path.join(VDIR, f))
ONLY = os.environ.get("ONLY", "icp.cu").split(",")
def one(spec):
    name, _, defs = spec.partition(":")
    defines = ["PR_DEBUG"] + [d for d in defs.split(",") if d]
    out = os.path.join(VDIR, f"lib_{name}.so")
    try:
        bm.build(out=out, defines=defines, only=ONLY)
    except Exception as e:
        return f"{name}: FAILED {e}"
    log = bm.LAST_LOG
    res = []
    for kern in ("icp_hyp_kernelINS_11PackedScene", "icp_hyp_kernelINS_13PackedNnScene", "raster_tile_kernel"):
        m = re.search(kern + r".*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", log, re.S)
        if m: res.append(f"{kern[:22]}: regs={m.group(3)} spill={m.group(2)}")
    return f"{name}: " + "; ".join(res)
for spec in sys.argv[1:]:        # builds share the object cache; compile one after the other (each build is parallel inside)
    print(one(spec), flush=True)
