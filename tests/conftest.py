import json
import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


@pytest.fixture(scope="session")
def golden():
    arrays = dict(np.load(os.path.join(GOLDEN, "fixture.npz")))
    with open(os.path.join(GOLDEN, "fixture.json")) as f:
        scal = json.load(f)
    return arrays, scal


@pytest.fixture(scope="session")
def mesh():
    from pose_refine_b200 import workloads as wl
    return wl.load_mesh_npz(os.path.join(GOLDEN, "obj_06_mesh.npz"))


def _checker(kind):
    from oracle import binding
    if not binding.available(kind):
        if kind == "port":
            import subprocess
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
        else:
            import subprocess
            if os.path.isdir("/root/reference/cuda_icp"):      # build the checker where the reference is mounted
                subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, capture_output=True)
            if not binding.available(kind):
                pytest.skip("oracle/_ref/libpose_refine_ref.so not built (needs /root/reference)")
    chk = binding.load(kind)
    chk.set_threads(1)
    return chk


@pytest.fixture(scope="session")
def port():
    """our CPU restatement (oracle/oracle.cpp)"""
    return _checker("port")


@pytest.fixture(scope="session")
def ref():
    """the reference's own CPU sources (oracle/_ref), when present"""
    return _checker("reference")


@pytest.fixture(scope="session")
def fixture_scene(port, mesh, golden):
    """The test.cpp scenario on the CPU oracle: depth of both fixture poses, model cloud, both scenes."""
    arrays, _ = golden
    K, proj, poses = arrays["K"], arrays["proj"], arrays["poses"]
    depth = port.render(mesh, poses, 640, 480, proj)
    cloud = port.depth2cloud(depth[0], K)
    return {"K": K, "proj": proj, "poses": poses, "depth": depth, "cloud": cloud, "scene_depth": depth[1]}
