"""ctypes binding of libpose_refine_b200.so -- one prototype per symbol of include/pose_refine_b200.h."""
import ctypes as C
import os

from .build import LIB


class Criteria(C.Structure):  # pr_icp_criteria
    _fields_ = [("relative_fitness", C.c_float), ("relative_rmse", C.c_float), ("max_iteration", C.c_int)]


class Roi(C.Structure):  # pr_roi
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("width", C.c_int), ("height", C.c_int)]


class SceneProjective(C.Structure):  # pr_scene_projective
    _fields_ = [("width", C.c_uint64), ("height", C.c_uint64), ("max_dist_diff", C.c_float), ("K", C.c_float * 9),
                ("pcd_dev", C.c_void_p), ("normal_dev", C.c_void_p)]


class SceneNN(C.Structure):  # pr_scene_nn
    _fields_ = [("max_dist_diff", C.c_float), ("pcd_dev", C.c_void_p), ("normal_dev", C.c_void_p),
                ("nodes_dev", C.c_void_p), ("n_points", C.c_uint64), ("n_nodes", C.c_uint64)]


class MeshClusters(C.Structure):  # pr_mesh_clusters
    _fields_ = [("n_clusters", C.c_size_t), ("vert_off_dev", C.c_void_p), ("verts_dev", C.c_void_p)]


_vp, _sz, _i, _u32, _f = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_float
_fp = C.POINTER(C.c_float)

# every exported symbol of include/pose_refine_b200.h: name -> (restype, argtypes)
PROTOTYPES = {
    "pr_version": (_i, []),
    "pr_error_string": (C.c_char_p, [_i]),
    "pr_device_check": (_i, []),
    "pr_device_malloc": (_i, [C.POINTER(_vp), _sz]),
    "pr_device_free": (_i, [_vp]),
    "pr_memcpy_h2d": (_i, [_vp, _vp, _sz, _vp]),
    "pr_memcpy_d2h": (_i, [_vp, _vp, _sz, _vp]),
    "pr_stream_synchronize": (_i, [_vp]),
    "pr_load_ply": (_i, [C.c_char_p, _vp, _sz, C.POINTER(_sz)]),
    "pr_compute_proj": (_i, [_vp, _i, _i, _f, _f, _vp]),
    "pr_render_workspace_bytes": (_sz, [_sz, _sz, _sz, _sz]),
    "pr_render_batch": (_i, [_vp, _sz, _vp, _i, _sz, _sz, _sz, _vp, Roi, _vp, _vp, _sz, _vp]),
    "pr_render_outputs_batch": (_i, [_vp, _sz, _vp, _i, _sz, _sz, _sz, _vp, Roi, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pr_mesh_index": (_i, [_vp, _sz, _vp, _vp, C.POINTER(_sz)]),
    "pr_render_indexed_workspace_bytes": (_sz, [_sz, _sz, _sz, _sz, _sz]),
    "pr_render_indexed_batch": (_i, [_vp, _sz, _vp, _sz, _vp, _i, _sz, _sz, _sz, _vp, Roi, _vp, _vp, _sz, _vp]),
    "pr_render_cloud_workspace_bytes": (_sz, [_sz, _sz, _sz, _sz, _sz]),
    "pr_mesh_cluster": (_i, [_vp, _sz, _vp, _sz, _vp, _vp, C.POINTER(_sz)]),
    "pr_render_cloud_batch": (_i, [_vp, _sz, _vp, _sz, _vp, _i, _sz, _sz, _sz, _vp, _vp, _vp, _vp, _sz, _u32, _vp, _vp, _vp,
                                   _vp, _vp, _sz, _vp]),
    "pr_raw2depth_mask": (_i, [_vp, _sz, _vp, _vp, _vp]),
    "pr_depth2cloud_workspace_bytes": (_sz, [_sz, _u32, _u32]),
    "pr_depth2cloud_count": (_i, [_vp, _i, _sz, _u32, _u32, _u32, _u32, _sz, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pr_depth2cloud_fill": (_i, [_vp, _i, _sz, _u32, _u32, _vp, _u32, _u32, _u32, _vp, _vp, _sz, _vp, _sz, _vp]),
    "pr_scene_projective_init": (_i, [_vp, _i, _u32, _u32, _vp, _vp, _vp, _vp]),
    "pr_scene_nn_build_host": (_i, [_vp, _i, _u32, _u32, _vp, _i, _vp, _vp, _sz, _vp, _sz, C.POINTER(_sz), C.POINTER(_sz)]),
    "pr_scene_nn_build_workspace_bytes": (_sz, [_u32, _u32]),
    "pr_scene_nn_build": (_i, [_vp, _i, _u32, _u32, _vp, _i, _vp, _vp, _sz, _vp, _sz, C.POINTER(_sz), C.POINTER(_sz), _vp, _sz, _vp]),
    "pr_icp_workspace_bytes": (_sz, [_sz, _sz, _sz]),
    "pr_icp_nn_workspace_bytes": (_sz, [_sz, _sz, _sz, _sz]),
    "pr_icp_projective_batch": (_i, [_vp, _vp, _vp, _sz, _sz, C.POINTER(SceneProjective), Criteria, _vp, _i, _vp, _sz, _vp]),
    "pr_icp_nn_batch": (_i, [_vp, _vp, _vp, _sz, _sz, C.POINTER(SceneNN), Criteria, _vp, _i, _vp, _sz, _vp]),
    "pr_solve_666": (_i, [_vp, _vp, _vp]),
    "pr_pcd2ab_projective": (_i, [_vp, _sz, C.POINTER(SceneProjective), _vp, _vp]),
    "pr_pcd2ab_nn": (_i, [_vp, _sz, C.POINTER(SceneNN), _vp, _vp]),
    "pr_refiner_create": (_i, [C.POINTER(_vp), _vp, _sz, _u32, _u32, _vp, _sz, _sz]),
    "pr_refiner_destroy": (None, [_vp]),
    "pr_refiner_set_scene_projective": (_i, [_vp, _vp, _i, _f]),
    "pr_refiner_set_scene_nn": (_i, [_vp, _vp, _i]),
    "pr_refiner_run": (_i, [_vp, _vp, _sz, Criteria, _vp, _vp]),
    "pr_refiner_run_device": (_i, [_vp, _vp, _sz, Criteria, _vp, _vp]),
    "pr_refiner_buffers": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "pr_refiner_overflow_flag": (_i, [_vp, C.POINTER(_vp)]),
    "pr_refiner_scene_buffers": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "pr_refiner_stage_ms": (_i, [_vp, C.POINTER(_f), C.POINTER(_f), C.POINTER(_u32)]),
    "pr_launch_count": (C.c_uint64, []),
    "pr_debug_div_check": (_i, [C.c_uint64, _u32, _vp, _vp]),
    "pr_scene_projective_packed_bytes": (_sz, [_u32, _u32]),
    "pr_scene_projective_pack": (_i, [C.POINTER(SceneProjective), _vp, _vp]),
    "pr_icp_projective_batch_packed": (_i, [_vp, _vp, _vp, _sz, _sz, C.POINTER(SceneProjective), _vp, Criteria, _vp, _i, _vp, _sz, _vp]),
    "pr_pass_sums_projective": (_i, [_vp, _vp, _vp, _sz, _sz, C.POINTER(SceneProjective), _vp, _vp, _sz, _vp]),
    "pr_pass_sums_nn": (_i, [_vp, _vp, _vp, _sz, _sz, C.POINTER(SceneNN), _vp, _vp, _sz, _vp]),
    "pr_correspondences_projective": (_i, [_vp, _sz, C.POINTER(SceneProjective), _vp, _vp, _sz, _vp]),
    "pr_correspondences_nn": (_i, [_vp, _sz, C.POINTER(SceneNN), _vp, _vp, _sz, _vp]),
    "pr_solve_666_device": (_i, [_vp, _sz, _i, _vp, _vp]),
    "pr_nn_walk_stats": (_i, [_vp, _sz, C.POINTER(SceneNN), _vp, _vp, _sz, _vp]),
    "pr_refiner_set_scene_projective_device": (_i, [_vp, _vp, _i, _f, _vp]),
    "pr_refiner_set_scene_nn_device": (_i, [_vp, _vp, _i, _vp]),
    "pr_device_count": (_i, [C.POINTER(_i)]),
    "pr_set_device": (_i, [_i]),
    "pr_shard_plan": (_i, [_sz, _i, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "pr_nccl_unique_id": (_i, [_vp]),
    "pr_comm_create": (_i, [C.POINTER(_vp), _vp, _i, _i]),
    "pr_comm_adopt": (_i, [C.POINTER(_vp), _vp, _i, _i]),
    "pr_comm_destroy": (None, [_vp]),
    "pr_broadcast_scene": (_i, [_vp, _vp, _sz, _i, _vp]),
    "pr_gather_results": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "pr_gather_wait": (_i, [_vp, _vp, _i]),
}

_dll = None
_path = LIB


def use_library(path):
    """Experiment scripts: bind an alternative build of the same library (scripts/build_variants.py) before first use."""
    global _path
    if _dll is not None:
        raise RuntimeError("the library is already loaded")
    _path = path


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _dll
    if _dll is None:
        path = _path
        if not os.path.exists(path):
            raise ImportError(
                f"{LIB} is missing: build it with `python -m pose_refine_b200.build` "
                "(nvcc, sm_100a). pose_refine_b200 has no CPU fallback.")
        dll = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(dll, name)  # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        _dll = dll
    return _dll


class PoseRefineError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        super().__init__(f"{where}: {lib().pr_error_string(status).decode()} (status {status})")


def check(status, where):
    if status != 0:
        raise PoseRefineError(status, where)
