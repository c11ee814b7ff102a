"""Build libpose_refine_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pose_refine_b200.build [--force]

Every source is compiled to its own object under pose_refine_b200/build/ (re-used while the source, the
headers and the flags are unchanged) and the objects are linked into the .so next to this file, so that it
travels with the repo snapshot to the GPU box.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpose_refine_b200.so")
HEADER = os.path.join(HERE, "..", "include", "pose_refine_b200.h")
SOURCES = ["raster.cu", "cloud.cu", "scene.cu", "icp.cu", "refiner.cu", "multi_gpu.cu", "host_io.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xptxas=-v", "-Xcompiler", "-fPIC,-O3", "-ccbin", "/usr/bin/g++",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-ccbin", "/usr/bin/g++", "-ldl"]


def _nvcc():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _headers_digest():
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + [HEADER, __file__]:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        if path.endswith((".cuh", ".h", ".hpp", ".py")):
            with open(path, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def _object(src, defines, digest, log):
    """Compile one source (cached by content + headers + defines); returns the object path."""
    with open(os.path.join(CSRC, src), "rb") as fh:
        key = hashlib.sha1(fh.read() + digest.encode() + " ".join(defines).encode()).hexdigest()[:16]
    obj = os.path.join(OBJ, f"{os.path.splitext(src)[0]}.{key}.o")
    if not os.path.exists(obj):
        cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        with open(obj + ".log", "w") as fh:
            fh.write(res.stderr)
    with open(obj + ".log") as fh:
        log.append(fh.read())
    return obj


def build(force=False, verbose=False, out=None, defines=(), only=None):
    """out / defines: experiment builds (e.g. defines=["PR_HYP_ILP=8"]) written where `out` says; `only` limits the
    defines to the named sources (the others come from the default build's cache).  Returns the library path
    (and keeps the ptxas -v output of the last build in build.LAST_LOG)."""
    global LAST_LOG
    os.makedirs(OBJ, exist_ok=True)
    digest = _headers_digest()
    log = []
    defines = list(defines)

    def one(src):
        d = defines if (only is None or src in only) else []
        return _object(src, d, digest, log)

    with ThreadPoolExecutor(4) as ex:
        objs = list(ex.map(one, SOURCES))
    out = out or LIB
    stamp = out + ".stamp"
    want = " ".join(objs)
    have = open(stamp).read() if os.path.exists(stamp) and os.path.exists(out) else ""
    if force or have != want:
        res = subprocess.run([_nvcc()] + LINK_FLAGS + ["-o", out] + objs, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed for libpose_refine_b200.so")
        with open(stamp, "w") as fh:
            fh.write(want)
    LAST_LOG = "\n".join(log)
    if verbose:
        print(LAST_LOG)
    return out


LAST_LOG = ""

if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
