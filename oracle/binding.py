"""ctypes doorway onto the CPU checkers -- TEST INFRASTRUCTURE ONLY.

`load("port")`      -> oracle/liboracle.so                 (our restatement, oracle/oracle.cpp)
`load("reference")` -> oracle/_ref/libpose_refine_ref.so   (the reference's own CPU sources,
                       compiled verbatim by oracle/Makefile; see oracle/ref_glue.cpp)

Both expose the same entry points (prefix orc_ / ref_), wrapped here as one class so a test can
run the same check against either.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package
(pose_refine_b200/) must never do so.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {
    "port": os.path.join(_HERE, "liboracle.so"),
    "reference": os.path.join(_HERE, "_ref", "libpose_refine_ref.so"),
}
_PREFIX = {"port": "orc_", "reference": "ref_"}

NODE_DTYPE = np.dtype([  # Node_kdtree, cuda_icp/scene/pcd_scene/pcd_scene.h:5-25 (52 bytes)
    ("parent", "<i4"), ("child1", "<i4"), ("child2", "<i4"), ("split_v", "<f4"),
    ("bbox", "<f4", (6,)), ("split_dim", "<i4"), ("left", "<i4"), ("right", "<i4")])
assert NODE_DTYPE.itemsize == 52


def available(kind):
    return os.path.exists(_PATHS[kind])


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Scene:
    def __init__(self, lib, handle, kind, W=None, H=None):
        self.lib, self.handle, self.kind, self.W, self.H = lib, handle, kind, W, H

    def sizes(self):
        n, m = C.c_long(), C.c_long()
        self.lib._fn("scene_sizes")(self.handle, C.byref(n), C.byref(m))
        return n.value, m.value

    def arrays(self):
        """(pcd [n,3], normal [n,3], nodes structured[n_nodes] or None)."""
        n, m = self.sizes()
        pcd = np.zeros((n, 3), np.float32)
        nrm = np.zeros((n, 3), np.float32)
        nodes = np.zeros(m, NODE_DTYPE) if m else None
        self.lib._fn("scene_get")(self.handle, pcd.ctypes.data, nrm.ctypes.data,
                                  nodes.ctypes.data if m else None)
        return pcd, nrm, nodes

    def close(self):
        if self.handle:
            self.lib._fn("scene_destroy")(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CpuChecker:
    def __init__(self, kind):
        if not available(kind):
            raise FileNotFoundError(f"{_PATHS[kind]} is not built (run `make -C oracle`)")
        self.kind = kind
        self.dll = C.CDLL(_PATHS[kind])
        self.p = _PREFIX[kind]
        vp, sz, f, i, u32, lng = C.c_void_p, C.c_size_t, C.c_float, C.c_int, C.c_uint32, C.c_long
        sig = {
            "set_threads": (None, [i]),
            "max_threads": (i, []),
            "compute_proj": (None, [vp, i, i, f, f, vp]),
            "render": (None, [vp, sz, vp, sz, sz, sz, vp, vp, vp]),
            "depth2cloud": (lng, [vp, i, u32, u32, vp, u32, u32, u32, vp, lng]),
            "get_normal": (None, [vp, i, i, i, vp, vp]),
            "scene_projective_create": (vp, [vp, i, vp, sz, sz, f]),
            "scene_nn_create": (vp, [vp, i, vp, sz, sz]),
            "scene_destroy": (None, [vp]),
            "scene_sizes": (None, [vp, C.POINTER(lng), C.POINTER(lng)]),
            "scene_get": (None, [vp, vp, vp, vp]),
            "query": (None, [vp, vp, sz, vp, vp, vp]),
            "pcd2ab": (None, [vp, vp, sz, vp]),
            "solve_666": (None, [vp, vp, vp]),
            "icp": (None if kind == "reference" else i, [vp, vp, sz, f, f, i, vp]),
            "pipeline": (C.c_double, [vp, vp, sz, vp, sz, sz, sz, vp, vp, f, f, i, i, vp, vp]),
        }
        if kind == "reference":
            sig["load_model"] = (lng, [C.c_char_p, vp, lng])
        else:
            sig["raw2depth_mask"] = (None, [vp, sz, vp, vp])
            sig["scene_nn_from_arrays"] = (vp, [vp, vp, lng, vp, lng, f])
            sig["query_nn_stats"] = (None, [vp, vp, sz, vp, C.POINTER(lng), C.POINTER(lng)])
        self._fns = {}
        for name, (res, args) in sig.items():
            fn = getattr(self.dll, self.p + name)
            fn.restype, fn.argtypes = res, args
            self._fns[name] = fn

    def _fn(self, name):
        return self._fns[name]

    # ---- threads ----
    def set_threads(self, n):
        self._fn("set_threads")(int(n))

    def max_threads(self):
        return self._fn("max_threads")()

    # ---- renderer ----
    def load_model(self, path):
        """reference only: cuda_renderer::Model(path).tris as [T,9] float32."""
        T = self._fn("load_model")(path.encode(), None, 0)
        tris = np.zeros((T, 9), np.float32)
        self._fn("load_model")(path.encode(), tris.ctypes.data, T)
        return tris

    def compute_proj(self, K, W, H, near=10.0, far=10000.0):
        K = _f32(K).reshape(9)
        out = np.zeros(16, np.float32)
        self._fn("compute_proj")(K.ctypes.data, W, H, near, far, out.ctypes.data)
        return out.reshape(4, 4)

    def render(self, tris, poses, W, H, proj, roi=(0, 0, 0, 0)):
        tris, poses, proj = _f32(tris).reshape(-1, 9), _f32(poses).reshape(-1, 16), _f32(proj).reshape(16)
        roi = np.asarray(roi, np.int32)
        rw, rh = (int(roi[2]), int(roi[3])) if roi[2] > 0 and roi[3] > 0 else (W, H)
        out = np.zeros((len(poses), rh, rw), np.int32)
        self._fn("render")(tris.ctypes.data, len(tris), poses.ctypes.data, len(poses), W, H,
                           proj.ctypes.data, roi.ctypes.data, out.ctypes.data)
        return out

    def raw2depth_mask(self, raw):
        raw = np.ascontiguousarray(raw, np.int32)
        d = np.zeros(raw.shape, np.uint16)
        m = np.zeros(raw.shape, np.uint8)
        self._fn("raw2depth_mask")(raw.ctypes.data, raw.size, d.ctypes.data, m.ctypes.data)
        return d, m

    # ---- clouds / scenes ----
    @staticmethod
    def _depth(depth):
        depth = np.ascontiguousarray(depth)
        assert depth.dtype in (np.int32, np.uint16), depth.dtype
        return depth, int(depth.dtype == np.int32)

    def depth2cloud(self, depth, K, stride=1, tl_x=0, tl_y=0):
        depth, is_i32 = self._depth(depth)
        H, W = depth.shape
        K = _f32(K).reshape(9)
        out = np.zeros((W * H, 3), np.float32)
        n = self._fn("depth2cloud")(depth.ctypes.data, is_i32, W, H, K.ctypes.data, stride, tl_x, tl_y,
                                    out.ctypes.data, W * H)
        if n < 0:
            raise ValueError("depth2cloud: unsupported arguments")
        return out[:n].copy()

    def get_normal(self, depth, K):
        depth, is_i32 = self._depth(depth)
        H, W = depth.shape
        K = _f32(K).reshape(9)
        out = np.zeros((H, W, 3), np.float32)
        self._fn("get_normal")(depth.ctypes.data, is_i32, W, H, K.ctypes.data, out.ctypes.data)
        return out

    def scene_projective(self, depth, K, max_dist=0.1):
        depth, is_i32 = self._depth(depth)
        H, W = depth.shape
        K = _f32(K).reshape(9)
        h = self._fn("scene_projective_create")(depth.ctypes.data, is_i32, K.ctypes.data, W, H, max_dist)
        return Scene(self, h, 0, W, H)

    def scene_nn(self, depth, K):
        depth, is_i32 = self._depth(depth)
        H, W = depth.shape
        K = _f32(K).reshape(9)
        h = self._fn("scene_nn_create")(depth.ctypes.data, is_i32, K.ctypes.data, W, H)
        return Scene(self, h, 1, W, H)

    def scene_nn_from_arrays(self, pcd, nrm, nodes, max_dist=0.1):
        pcd, nrm = _f32(pcd), _f32(nrm)
        nodes = np.ascontiguousarray(nodes)
        h = self._fn("scene_nn_from_arrays")(pcd.ctypes.data, nrm.ctypes.data, len(pcd), nodes.ctypes.data,
                                             len(nodes), max_dist)
        return Scene(self, h, 1)

    # ---- ICP pieces ----
    def query(self, scene, pts):
        pts = _f32(pts).reshape(-1, 3)
        n = len(pts)
        dst, nrm, valid = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.uint8)
        self._fn("query")(scene.handle, pts.ctypes.data, n, dst.ctypes.data, nrm.ctypes.data, valid.ctypes.data)
        return dst, nrm, valid.astype(bool)

    def query_nn_stats(self, scene, pts):
        pts = _f32(pts).reshape(-1, 3)
        idx = np.zeros(len(pts), np.int32)
        v, t = C.c_long(), C.c_long()
        self._fn("query_nn_stats")(scene.handle, pts.ctypes.data, len(pts), idx.ctypes.data, C.byref(v), C.byref(t))
        return idx, v.value, t.value

    def pcd2ab(self, scene, pts):
        pts = _f32(pts).reshape(-1, 3)
        out = np.zeros(29, np.float32)
        self._fn("pcd2ab")(scene.handle, pts.ctypes.data, len(pts), out.ctypes.data)
        return out

    def solve_666(self, A, b):
        A, b = _f32(A).reshape(36), _f32(b).reshape(6)
        T = np.zeros(16, np.float32)
        self._fn("solve_666")(A.ctypes.data, b.ctypes.data, T.ctypes.data)
        return T.reshape(4, 4)

    def icp(self, scene, pts, rel_fit=1e-5, rel_rmse=1e-5, max_iter=30):
        """Returns dict(T[4,4], rmse, fitness, pts (transformed copy), last_pass or None)."""
        pts = _f32(pts).reshape(-1, 3).copy()
        out = np.zeros(18, np.float32)
        it = self._fn("icp")(scene.handle, pts.ctypes.data, len(pts), rel_fit, rel_rmse, max_iter, out.ctypes.data)
        return {"T": out[:16].reshape(4, 4).copy(), "rmse": float(out[16]), "fitness": float(out[17]),
                "pts": pts, "last_pass": it if self.kind == "port" else None, "raw": out}

    def pipeline(self, scene, tris, poses, W, H, proj, K, rel_fit=0.0, rel_rmse=0.0, max_iter=30, schedule=1):
        """render -> depth2cloud -> ICP for every pose on the CPU; returns (seconds, results[P,18], n_pts[P])."""
        tris, poses = _f32(tris).reshape(-1, 9), _f32(poses).reshape(-1, 16)
        proj, K = _f32(proj).reshape(16), _f32(K).reshape(9)
        P = len(poses)
        res = np.zeros((P, 18), np.float32)
        npts = np.zeros(P, np.int64)
        sec = self._fn("pipeline")(scene.handle, tris.ctypes.data, len(tris), poses.ctypes.data, P, W, H,
                                   proj.ctypes.data, K.ctypes.data, rel_fit, rel_rmse, max_iter, schedule,
                                   res.ctypes.data, npts.ctypes.data)
        return sec, res, npts


_cache = {}


def load(kind="port"):
    if kind not in _cache:
        _cache[kind] = CpuChecker(kind)
    return _cache[kind]
