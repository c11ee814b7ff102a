"""Python host mirror of the reference's interface for the hot path.

Names, argument meaning and error behaviour follow cuda_renderer/renderer.h and cuda_icp/icp.h
(+ scene/*.h); every function here only marshals arguments into the C ABI of
libpose_refine_b200.so (include/pose_refine_b200.h).  torch supplies device memory and the
current CUDA stream -- nothing is computed in Python and there is no CPU fallback.

    reference (C++)                                   here
    ------------------------------------------------  -------------------------------------------
    Model(path).tris                 renderer.cpp:11   load_ply(path)
    compute_proj                     renderer.cpp:161  compute_proj
    render_cuda_keep_in_gpu          renderer.cu:269   render_cuda_keep_in_gpu
    render_cuda                      renderer.cu:189   render_cuda
    raw2depth_uint16/mask/..._cuda   renderer.cu:354   raw2depth_uint16_cuda / raw2mask_uint8_cuda / raw2depth_mask_cuda
    depth2cloud_cuda                 icp.cu:256        depth2cloud_cuda (+ depth2cloud_batch)
    Scene_projective::init_..._cuda  depth_scene.cu:3  SceneProjective.init_cuda
    Scene_nn::init_Scene_nn_cuda     pcd_scene.cu:3    SceneNN.init_cuda
    ICP_Point2Plane_cuda             icp.cu:156        ICP_Point2Plane_cuda (+ icp_batch)
    eigen_slover_666                 icp.cpp:29        eigen_solver_666
    PoseRenderer / test.cpp:143-172  pose_renderer.h   PoseRefiner
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

NODE_DTYPE = np.dtype([  # Node_kdtree, pcd_scene.h:5-25
    ("parent", "<i4"), ("child1", "<i4"), ("child2", "<i4"), ("split_v", "<f4"),
    ("bbox", "<f4", (6,)), ("split_dim", "<i4"), ("left", "<i4"), ("right", "<i4")])


def _require_device():
    if not torch.cuda.is_available():
        raise RuntimeError("pose_refine_b200 needs an sm_100 CUDA device (there is no CPU fallback)")
    check(lib().pr_device_check(), "pr_device_check")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _dev(t, dtype=None):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.ascontiguousarray(t))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if not t.is_cuda:
        t = t.cuda()
    return t.contiguous()


class _CudaView:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _view(ptr, shape, typestr):
    """torch view (no copy) of device memory owned by the library."""
    return torch.as_tensor(_CudaView(ptr, shape, typestr), device="cuda")


# ---------------------------------------------------------------------------------------------
@dataclass
class ICPConvergenceCriteria:  # icp.h:38-50
    relative_fitness: float = 1e-5
    relative_rmse: float = 1e-5
    max_iteration: int = 30

    def c(self):
        return _lib.Criteria(self.relative_fitness, self.relative_rmse, self.max_iteration)


@dataclass
class RegistrationResult:  # icp.h:26-36
    transformation_: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    inlier_rmse_: float = 0.0
    fitness_: float = 0.0


def load_ply(path):
    """Model(path).tris (renderer.cpp:11-58): [T, 9] float32, face order, mm."""
    n = C.c_size_t()
    check(lib().pr_load_ply(path.encode(), None, 0, C.byref(n)), "pr_load_ply")
    tris = np.zeros((n.value, 9), np.float32)
    check(lib().pr_load_ply(path.encode(), tris.ctypes.data, n.value, C.byref(n)), "pr_load_ply")
    return tris


def compute_proj(K, width, height, near=10.0, far=10000.0):
    K = _f32c(K).reshape(9)
    out = np.zeros(16, np.float32)
    check(lib().pr_compute_proj(K.ctypes.data, width, height, near, far, out.ctypes.data), "pr_compute_proj")
    return out.reshape(4, 4)


def eigen_solver_666(A, b):
    """eigen_slover_666 (icp.cpp:29-45), host."""
    A, b = _f32c(A).reshape(36), _f32c(b).reshape(6)
    T = np.zeros(16, np.float32)
    check(lib().pr_solve_666(A.ctypes.data, b.ctypes.data, T.ctypes.data), "pr_solve_666")
    return T.reshape(4, 4)


# ---------------------------------------------------------------------------------------------
# renderer
def render_cuda_keep_in_gpu(tris, poses, width, height, proj_mat, roi=(0, 0, 0, 0), use_tiles=True):
    """-> int32 cuda tensor [P, H', W'] (renderer.cu:269-336). tris: [T,9] (host or cuda)."""
    _require_device()
    tris = _dev(tris, torch.float32).reshape(-1, 9)
    proj = _f32c(proj_mat).reshape(16)
    roi_c = _lib.Roi(*[int(v) for v in roi])
    rw, rh = (roi_c.width, roi_c.height) if roi_c.width > 0 and roi_c.height > 0 else (width, height)
    on_dev = isinstance(poses, torch.Tensor) and poses.is_cuda
    if on_dev:
        poses_t = poses.to(torch.float32).contiguous().reshape(-1, 16)
        n_poses, poses_ptr = poses_t.shape[0], poses_t.data_ptr()
    else:
        poses_h = _f32c(poses.cpu().numpy() if isinstance(poses, torch.Tensor) else poses).reshape(-1, 16)
        n_poses, poses_ptr = poses_h.shape[0], poses_h.ctypes.data
    out = torch.empty((n_poses, rh, rw), dtype=torch.int32, device="cuda")
    if use_tiles:
        ws_bytes = lib().pr_render_workspace_bytes(n_poses, tris.shape[0], width, height)
    else:
        ws_bytes = max(256, n_poses * 64 + 256)   # poses only -> global-atomic path
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    check(lib().pr_render_batch(tris.data_ptr(), tris.shape[0], poses_ptr, int(on_dev), n_poses, width, height,
                                proj.ctypes.data, roi_c, out.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
          "pr_render_batch")
    if not on_dev:
        torch.cuda.current_stream().synchronize()   # host poses were read asynchronously
    return out


def render_depth_mask_cuda(tris, poses, width, height, proj_mat, want_depth=True, want_mask=True, want_raw=False, roi=(0, 0, 0, 0)):
    """The rasteriser with PoseRenderer's outputs folded into its write-out (pr_render_outputs_batch):
    -> (uint16 depth [P,H',W'] or None, uint8 mask [P,H',W'] or None, int32 raw [P,H',W'] or None), cuda tensors."""
    _require_device()
    tris = _dev(tris, torch.float32).reshape(-1, 9)
    proj = _f32c(proj_mat).reshape(16)
    roi_c = _lib.Roi(*[int(v) for v in roi])
    rw, rh = (roi_c.width, roi_c.height) if roi_c.width > 0 and roi_c.height > 0 else (width, height)
    poses_t = _dev(poses, torch.float32).reshape(-1, 16)
    P = poses_t.shape[0]
    d16 = torch.empty((P, rh, rw), dtype=torch.uint16, device="cuda") if want_depth else None
    m8 = torch.empty((P, rh, rw), dtype=torch.uint8, device="cuda") if want_mask else None
    raw = torch.empty((P, rh, rw), dtype=torch.int32, device="cuda") if want_raw else None
    ws_bytes = lib().pr_render_workspace_bytes(P, tris.shape[0], width, height)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    check(lib().pr_render_outputs_batch(tris.data_ptr(), tris.shape[0], poses_t.data_ptr(), 1, P, width, height, proj.ctypes.data, roi_c,
                                        raw.data_ptr() if want_raw else None, d16.data_ptr() if want_depth else None,
                                        m8.data_ptr() if want_mask else None, ws.data_ptr(), ws_bytes, _stream()), "pr_render_outputs_batch")
    return d16, m8, raw


class PoseRenderer:
    """PoseRenderer (pose_renderer.h:9-32, pose_renderer.cpp): model uploaded once, batches of poses rendered to uint16
    depth / uint8 mask; down_sample renders at width/ds x height/ds with the FULL-resolution projection
    (pose_renderer.cpp:25-36)."""

    def __init__(self, tris_or_path):
        _require_device()
        tris = load_ply(tris_or_path) if isinstance(tris_or_path, str) else _f32c(tris_or_path).reshape(-1, 9)
        self.tris = torch.as_tensor(tris).cuda()

    def set_K_width_height(self, K, width, height):
        self.K, self.width, self.height = _f32c(K).reshape(3, 3), int(width), int(height)
        self.proj_mat = compute_proj(self.K, self.width, self.height)

    def _render(self, init_poses, down_sample, want_depth, want_mask):
        w, h = int(self.width / down_sample), int(self.height / down_sample)
        return render_depth_mask_cuda(self.tris, init_poses, w, h, self.proj_mat, want_depth, want_mask)

    def render_depth(self, init_poses, down_sample=1):
        return self._render(init_poses, down_sample, True, False)[0]

    def render_mask(self, init_poses, down_sample=1):
        return self._render(init_poses, down_sample, False, True)[1]

    def render_depth_mask(self, init_poses, down_sample=1):
        d, m, _ = self._render(init_poses, down_sample, True, True)
        return d, m


def mesh_index(tris):
    """Deduplicate a triangle soup [T,9] -> (vertices [V,3] float32, faces [T,3] int32) (pr_mesh_index, host)."""
    tris = _f32c(tris).reshape(-1, 9)
    n = C.c_size_t()
    verts = np.zeros((tris.shape[0] * 3, 3), np.float32)
    faces = np.zeros((tris.shape[0], 3), np.int32)
    check(lib().pr_mesh_index(tris.ctypes.data, tris.shape[0], verts.ctypes.data, faces.ctypes.data, C.byref(n)), "pr_mesh_index")
    return verts[: n.value].copy(), faces


def render_indexed_keep_in_gpu(verts, faces, poses, width, height, proj_mat, roi=(0, 0, 0, 0)):
    """Indexed-mesh variant of render_cuda_keep_in_gpu: same output, vertices projected once per pose."""
    _require_device()
    verts = _dev(verts, torch.float32).reshape(-1, 3)
    faces = _dev(faces, torch.int32).reshape(-1, 3)
    proj = _f32c(proj_mat).reshape(16)
    roi_c = _lib.Roi(*[int(v) for v in roi])
    rw, rh = (roi_c.width, roi_c.height) if roi_c.width > 0 and roi_c.height > 0 else (width, height)
    poses_t = _dev(poses, torch.float32).reshape(-1, 16)
    n_poses = poses_t.shape[0]
    out = torch.empty((n_poses, rh, rw), dtype=torch.int32, device="cuda")
    ws_bytes = lib().pr_render_indexed_workspace_bytes(n_poses, verts.shape[0], faces.shape[0], width, height)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    check(lib().pr_render_indexed_batch(verts.data_ptr(), verts.shape[0], faces.data_ptr(), faces.shape[0], poses_t.data_ptr(), 1,
                                        n_poses, width, height, proj.ctypes.data, roi_c, out.data_ptr(), ws.data_ptr(), ws_bytes,
                                        _stream()), "pr_render_indexed_batch")
    return out


def mesh_cluster(verts, faces):
    """pr_mesh_cluster (host): Morton-ordered faces + per-cluster unique vertex lists.
    -> (faces [T,3] int32 reordered, vert_off [C+1] int32, cluster_verts [n] int32)"""
    verts = _f32c(verts).reshape(-1, 3)
    faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3).copy()
    T = faces.shape[0]
    off = np.zeros((T + 63) // 64 + 1, np.int32)
    cv = np.zeros(3 * T, np.int32)
    n = C.c_size_t()
    check(lib().pr_mesh_cluster(verts.ctypes.data, verts.shape[0], faces.ctypes.data, T, off.ctypes.data, cv.ctypes.data, C.byref(n)),
          "pr_mesh_cluster")
    off = off[: n.value + 1].copy()
    return faces, off, cv[: off[-1]].copy()


def render_clustered_keep_in_gpu(verts, faces, poses, width, height, proj_mat, clusters, out=None, ws=None):
    """Depth only through the cluster-binned rasteriser (pr_render_cloud_batch without clouds): same output as
    render_indexed_keep_in_gpu.  verts / faces / poses may be device tensors; out / ws can be reused between calls."""
    _require_device()
    verts = _dev(verts, torch.float32).reshape(-1, 3)
    faces = _dev(faces, torch.int32).reshape(-1, 3)
    proj = _f32c(proj_mat).reshape(16)
    poses_t = _dev(poses, torch.float32).reshape(-1, 16)
    n_poses = poses_t.shape[0]
    off_d, cv_d = _dev(clusters[0], torch.int32), _dev(clusters[1], torch.int32)
    cl = _lib.MeshClusters(off_d.shape[0] - 1, off_d.data_ptr(), cv_d.data_ptr())
    if out is None:
        out = torch.empty((n_poses, height, width), dtype=torch.int32, device="cuda")
    ws_bytes = lib().pr_render_cloud_workspace_bytes(n_poses, verts.shape[0], faces.shape[0], width, height)
    if ws is None or ws.numel() < ws_bytes:
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    check(lib().pr_render_cloud_batch(verts.data_ptr(), verts.shape[0], faces.data_ptr(), faces.shape[0], poses_t.data_ptr(), 1,
                                      n_poses, width, height, proj.ctypes.data, None, out.data_ptr(), None, 0, 4, None, None, None,
                                      C.cast(C.pointer(cl), C.c_void_p), ws.data_ptr(), ws.numel(), _stream()), "pr_render_cloud_batch")
    return out


def render_cloud_batch(verts, faces, poses, width, height, proj_mat, K, capacity_points=None, align_points=4, clusters=None):
    """Fused render_cuda_keep_in_gpu + depth2cloud_cuda per pose (pr_render_cloud_batch).
    clusters: (vert_off, cluster_verts) from mesh_cluster -- faces must then be mesh_cluster's reordered faces.
    -> (depth [P,H,W] int32, pts [cap,3] float32, offsets [P+1] int32, counts [P] int32), all on the device.
    Cloud i = pts[offsets[i] : offsets[i] + counts[i]], the points of depth2cloud in screen-tile order."""
    _require_device()
    verts = _dev(verts, torch.float32).reshape(-1, 3)
    faces = _dev(faces, torch.int32).reshape(-1, 3)
    proj = _f32c(proj_mat).reshape(16)
    Kc = _f32c(K).reshape(9)
    poses_t = _dev(poses, torch.float32).reshape(-1, 16)
    n_poses = poses_t.shape[0]
    depth = torch.empty((n_poses, height, width), dtype=torch.int32, device="cuda")
    cap = int(capacity_points) if capacity_points is not None else n_poses * width * height
    pts = torch.empty((cap + 8, 3), dtype=torch.float32, device="cuda")
    counts = torch.empty(n_poses, dtype=torch.int32, device="cuda")
    offsets = torch.empty(n_poses + 1, dtype=torch.int32, device="cuda")
    overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws_bytes = lib().pr_render_cloud_workspace_bytes(n_poses, verts.shape[0], faces.shape[0], width, height)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    cl_ptr = None
    if clusters is not None:
        off_d, cv_d = _dev(clusters[0], torch.int32), _dev(clusters[1], torch.int32)
        cl = _lib.MeshClusters(off_d.shape[0] - 1, off_d.data_ptr(), cv_d.data_ptr())
        cl_ptr = C.cast(C.pointer(cl), C.c_void_p)
    check(lib().pr_render_cloud_batch(verts.data_ptr(), verts.shape[0], faces.data_ptr(), faces.shape[0], poses_t.data_ptr(), 1,
                                      n_poses, width, height, proj.ctypes.data, Kc.ctypes.data, depth.data_ptr(), pts.data_ptr(),
                                      cap, align_points, counts.data_ptr(), offsets.data_ptr(), overflow.data_ptr(), cl_ptr,
                                      ws.data_ptr(), ws_bytes, _stream()), "pr_render_cloud_batch")
    return depth, pts, offsets, counts


def render_cuda(tris, poses, width, height, proj_mat, roi=(0, 0, 0, 0)):
    """-> host int32 array (renderer.cu:189-267)."""
    return render_cuda_keep_in_gpu(tris, poses, width, height, proj_mat, roi).cpu().numpy()


def raw2depth_mask_cuda(raw):
    """(uint16 depth, uint8 mask) cuda tensors shaped like raw (renderer.cu:402-439)."""
    _require_device()
    raw = _dev(raw, torch.int32)
    d = torch.empty(raw.shape, dtype=torch.uint16, device="cuda")
    m = torch.empty(raw.shape, dtype=torch.uint8, device="cuda")
    check(lib().pr_raw2depth_mask(raw.data_ptr(), raw.numel(), d.data_ptr(), m.data_ptr(), _stream()), "pr_raw2depth_mask")
    return d, m


def raw2depth_uint16_cuda(raw):
    _require_device()
    raw = _dev(raw, torch.int32)
    d = torch.empty(raw.shape, dtype=torch.uint16, device="cuda")
    check(lib().pr_raw2depth_mask(raw.data_ptr(), raw.numel(), d.data_ptr(), None, _stream()), "pr_raw2depth_mask")
    return d


def raw2mask_uint8_cuda(raw):
    _require_device()
    raw = _dev(raw, torch.int32)
    m = torch.empty(raw.shape, dtype=torch.uint8, device="cuda")
    check(lib().pr_raw2depth_mask(raw.data_ptr(), raw.numel(), None, m.data_ptr(), _stream()), "pr_raw2depth_mask")
    return m


# ---------------------------------------------------------------------------------------------
# clouds
def depth2cloud_batch(depth, K, stride=1, tl_x=0, tl_y=0, align_points=4):
    """depth: cuda tensor [P,H,W] int32 or uint16 -> (pts [cap,3] f32, offsets [P+1] i32, counts [P] i32)."""
    _require_device()
    assert depth.is_cuda and depth.dim() == 3 and depth.dtype in (torch.int32, torch.uint16)
    depth = depth.contiguous()
    P, H, W = depth.shape
    K = _f32c(K).reshape(9)
    is_i32 = int(depth.dtype == torch.int32)
    ws_bytes = lib().pr_depth2cloud_workspace_bytes(P, W, H)
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device="cuda")
    counts = torch.empty(max(P, 1), dtype=torch.int32, device="cuda")
    offsets = torch.empty(P + 1, dtype=torch.int32, device="cuda")
    check(lib().pr_depth2cloud_count(depth.data_ptr(), is_i32, P, W, H, stride, align_points, 0, counts.data_ptr(),
                                     offsets.data_ptr(), None, ws.data_ptr(), ws_bytes, _stream()), "pr_depth2cloud_count")
    total = int(offsets[P].item()) if P else 0          # same D2H read upstream does (icp.cu:272-274)
    pts = torch.zeros((max(total, 4), 3), dtype=torch.float32, device="cuda")
    check(lib().pr_depth2cloud_fill(depth.data_ptr(), is_i32, P, W, H, K.ctypes.data, stride, tl_x, tl_y,
                                    offsets.data_ptr(), pts.data_ptr(), pts.shape[0], ws.data_ptr(), ws_bytes, _stream()),
          "pr_depth2cloud_fill")
    return pts, offsets, counts[:P]


def depth2cloud_cuda(depth, width, height, K, stride=1, tl_x=0, tl_y=0):
    """depth2cloud_cuda<T> (icp.cu:256-286) for one device image -> cuda float32 [N,3]."""
    depth = depth.reshape(1, height, width)
    pts, offsets, counts = depth2cloud_batch(depth, K, stride, tl_x, tl_y, align_points=1)
    return pts[: int(counts[0].item())]


# ---------------------------------------------------------------------------------------------
# scenes
class SceneProjective:
    """Scene_projective (depth_scene.h:7-48); buffers live on the device, owned by this object."""

    def __init__(self):
        self.width, self.height, self.max_dist_diff = 640, 480, 0.1
        self.K = np.eye(3, dtype=np.float32)
        self.pcd = self.normal = None

    def init_cuda(self, scene_depth, K, width=640, height=480, max_dist_diff=0.1):
        """init_Scene_projective_cuda (depth_scene.cu:3-20). scene_depth: [H,W] uint16 / int32, host or cuda."""
        _require_device()
        d = scene_depth if isinstance(scene_depth, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(scene_depth))
        assert d.dtype in (torch.int32, torch.uint16), "CV_16U or CV_32S (depth_scene.cpp:11-12)"
        assert tuple(d.shape) == (height, width)
        d = d.cuda().contiguous()
        self.width, self.height, self.max_dist_diff = width, height, max_dist_diff
        self.K = _f32c(K).reshape(3, 3)
        self.pcd = torch.empty((height * width, 3), dtype=torch.float32, device="cuda")
        self.normal = torch.empty((height * width, 3), dtype=torch.float32, device="cuda")
        check(lib().pr_scene_projective_init(d.data_ptr(), int(d.dtype == torch.int32), width, height,
                                             self.K.ctypes.data, self.pcd.data_ptr(), self.normal.data_ptr(), _stream()),
              "pr_scene_projective_init")
        return self

    def c(self):
        s = _lib.SceneProjective()
        s.width, s.height, s.max_dist_diff = self.width, self.height, self.max_dist_diff
        s.K[:] = [float(v) for v in self.K.reshape(9)]
        s.pcd_dev, s.normal_dev = self.pcd.data_ptr(), self.normal.data_ptr()
        return s


class SceneNN:
    """Scene_nn + KDTree_cuda (pcd_scene.h:37-136); buffers on the device, owned by this object."""

    def __init__(self):
        self.max_dist_diff = 0.1
        self.pcd = self.normal = self.nodes = None
        self.nodes_host = None

    def init_cuda(self, scene_depth, K, max_leaf=10):
        """init_Scene_nn_cuda (pcd_scene.cu:3-20) with the kd-tree built on the DEVICE (pr_scene_nn_build): the same
        arrays the host build produces, bit for bit."""
        _require_device()
        d = scene_depth if isinstance(scene_depth, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(scene_depth))
        assert d.dtype in (torch.int32, torch.uint16), "CV_16U or CV_32S (pcd_scene.cpp:6-7)"
        d = d.cuda().contiguous()
        H, W = d.shape
        K = _f32c(K).reshape(9)
        cap = H * W
        self.pcd = torch.zeros((cap, 3), dtype=torch.float32, device="cuda")
        self.normal = torch.zeros((cap, 3), dtype=torch.float32, device="cuda")
        nodes = torch.zeros((2 * cap + 1) * 52, dtype=torch.uint8, device="cuda")
        ws_bytes = lib().pr_scene_nn_build_workspace_bytes(W, H)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        n_pts, n_nodes = C.c_size_t(), C.c_size_t()
        check(lib().pr_scene_nn_build(d.data_ptr(), int(d.dtype == torch.int32), W, H, K.ctypes.data, max_leaf, self.pcd.data_ptr(),
                                      self.normal.data_ptr(), cap, nodes.data_ptr(), 2 * cap + 1, C.byref(n_pts), C.byref(n_nodes),
                                      ws.data_ptr(), ws_bytes, _stream()), "pr_scene_nn_build")
        self.pcd, self.normal = self.pcd[: n_pts.value].contiguous(), self.normal[: n_pts.value].contiguous()
        self.nodes = nodes[: max(n_nodes.value, 1) * 52].contiguous()
        self.nodes_host = np.ascontiguousarray(self.nodes[: n_nodes.value * 52].cpu().numpy()).view(NODE_DTYPE)
        return self

    def init_host_build(self, scene_depth, K, max_leaf=10):
        """init_Scene_nn_cuda (pcd_scene.cu:3-20) as upstream does it: kd-tree built on the host, uploaded."""
        _require_device()
        d = np.ascontiguousarray(scene_depth.cpu().numpy() if isinstance(scene_depth, torch.Tensor) else scene_depth)
        assert d.dtype in (np.int32, np.uint16), "CV_16U or CV_32S (pcd_scene.cpp:6-7)"
        H, W = d.shape
        K = _f32c(K).reshape(9)
        cap = H * W
        pcd = np.zeros((cap, 3), np.float32)
        nrm = np.zeros((cap, 3), np.float32)
        nodes = np.zeros(2 * cap + 1, NODE_DTYPE)
        n_pts, n_nodes = C.c_size_t(), C.c_size_t()
        check(lib().pr_scene_nn_build_host(d.ctypes.data, int(d.dtype == np.int32), W, H, K.ctypes.data, max_leaf,
                                           pcd.ctypes.data, nrm.ctypes.data, cap, nodes.ctypes.data, len(nodes),
                                           C.byref(n_pts), C.byref(n_nodes)), "pr_scene_nn_build_host")
        return self.from_arrays(pcd[: n_pts.value], nrm[: n_pts.value], nodes[: n_nodes.value])

    def from_arrays(self, pcd, normal, nodes, max_dist_diff=0.1):
        _require_device()
        self.max_dist_diff = max_dist_diff
        self.nodes_host = np.ascontiguousarray(nodes)
        self.pcd = torch.as_tensor(_f32c(pcd).reshape(-1, 3)).cuda()
        self.normal = torch.as_tensor(_f32c(normal).reshape(-1, 3)).cuda()
        self.nodes = torch.as_tensor(self.nodes_host.view(np.uint8).reshape(-1)).cuda() if len(nodes) else torch.zeros(52, dtype=torch.uint8, device="cuda")
        return self

    def c(self):
        s = _lib.SceneNN()
        s.max_dist_diff = self.max_dist_diff
        s.pcd_dev, s.normal_dev, s.nodes_dev = self.pcd.data_ptr(), self.normal.data_ptr(), self.nodes.data_ptr()
        s.n_points, s.n_nodes = self.pcd.shape[0], len(self.nodes_host)
        return s


# ---------------------------------------------------------------------------------------------
# ICP
def _icp_workspace(P, cap, scene):
    if isinstance(scene, SceneProjective):
        ws_bytes = lib().pr_icp_workspace_bytes(P, cap, scene.width * scene.height)
    else:
        ws_bytes = lib().pr_icp_nn_workspace_bytes(P, cap, scene.pcd.shape[0], len(scene.nodes_host))
    return torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device="cuda"), ws_bytes


def icp_batch(pts, offsets, counts, scene, criteria=None, update_points=False, reference_arithmetic=False):
    """Batched ICP_Point2Plane_cuda: pts cuda [cap,3]; offsets/counts cuda int32 -> cuda float32 [P,18]
    (row-major 4x4, inlier_rmse_, fitness_ per hypothesis).  reference_arithmetic: the cross-check driver
    (PR_ICP_REFERENCE_ARITHMETIC: one launch per pass, every operation in the reference's order)."""
    _require_device()
    criteria = criteria or ICPConvergenceCriteria()
    assert pts.is_cuda and pts.dtype == torch.float32 and pts.is_contiguous()
    offsets, counts = _dev(offsets, torch.int32), _dev(counts, torch.int32)
    P = counts.shape[0]
    cap = pts.shape[0]
    res = torch.empty((max(P, 1), 18), dtype=torch.float32, device="cuda")
    ws, ws_bytes = _icp_workspace(P, cap, scene)
    flags = (1 if update_points else 0) | (2 if reference_arithmetic else 0)
    sc = scene.c()
    if isinstance(scene, SceneProjective):
        rc = lib().pr_icp_projective_batch(pts.data_ptr(), offsets.data_ptr(), counts.data_ptr(), P, cap, C.byref(sc),
                                           criteria.c(), res.data_ptr(), flags, ws.data_ptr(), ws_bytes, _stream())
    else:
        rc = lib().pr_icp_nn_batch(pts.data_ptr(), offsets.data_ptr(), counts.data_ptr(), P, cap, C.byref(sc),
                                   criteria.c(), res.data_ptr(), flags, ws.data_ptr(), ws_bytes, _stream())
    check(rc, "pr_icp_batch")
    return res[:P]


def ICP_Point2Plane_cuda(model_pcd, scene, criteria=None):
    """icp.cu:156-217 for one cloud: model_pcd (cuda [N,3] float32) is transformed IN PLACE, like upstream."""
    assert model_pcd.is_cuda and model_pcd.dtype == torch.float32 and model_pcd.is_contiguous()
    n = model_pcd.shape[0]
    offsets = torch.tensor([0], dtype=torch.int32, device="cuda")
    counts = torch.tensor([n], dtype=torch.int32, device="cuda")
    res = icp_batch(model_pcd, offsets, counts, scene, criteria, update_points=True)[0].cpu().numpy()
    return RegistrationResult(res[:16].reshape(4, 4).copy(), float(res[16]), float(res[17]))


def pass_sums(pts, offsets, counts, scene):
    """One evaluation pass of the SHIPPED ICP kernel (identity transform) over a ragged batch -> [P,29] float32 (host):
    the 29 sums of thrust__pcd2Ab per hypothesis (pr_pass_sums_*)."""
    _require_device()
    assert pts.is_cuda and pts.dtype == torch.float32 and pts.is_contiguous()
    offsets, counts = _dev(offsets, torch.int32), _dev(counts, torch.int32)
    P, cap = counts.shape[0], pts.shape[0]
    out = torch.zeros((max(P, 1), 32), dtype=torch.float32, device="cuda")
    ws, ws_bytes = _icp_workspace(P, cap, scene)
    sc = scene.c()
    fn = lib().pr_pass_sums_projective if isinstance(scene, SceneProjective) else lib().pr_pass_sums_nn
    check(fn(pts.data_ptr(), offsets.data_ptr(), counts.data_ptr(), P, cap, C.byref(sc), out.data_ptr(), ws.data_ptr(), ws_bytes,
             _stream()), "pr_pass_sums")
    return out[:P, :29].cpu().numpy()


def correspondences(pts, scene):
    """For every point of a cloud, the scene index the shipped kernel's search picks (-1: none) -> int32 [n] (host).
    Projective: pixel u + v*W after the depth gate; nn: index into the leaf-ordered scene points."""
    _require_device()
    pts = _dev(pts, torch.float32).reshape(-1, 3)
    n = pts.shape[0]
    idx = torch.empty(max(n, 1), dtype=torch.int32, device="cuda")
    ws, ws_bytes = _icp_workspace(1, n, scene)
    sc = scene.c()
    fn = lib().pr_correspondences_projective if isinstance(scene, SceneProjective) else lib().pr_correspondences_nn
    check(fn(pts.data_ptr(), n, C.byref(sc), idx.data_ptr(), ws.data_ptr(), ws_bytes, _stream()), "pr_correspondences")
    return idx[:n].cpu().numpy()


def solve_666_device(S29, fast=True):
    """n normal-equation systems (the 29 sums each, thrust__pcd2Ab's order) solved on the device, one thread each
    -> [n,4,4] float32.  fast: the solver the ICP kernel runs between passes; else Eigen's pivoted LDL^T restated."""
    _require_device()
    S = _dev(_f32c(S29).reshape(-1, 29), torch.float32)
    n = S.shape[0]
    E = torch.empty((max(n, 1), 16), dtype=torch.float32, device="cuda")
    check(lib().pr_solve_666_device(S.data_ptr(), n, int(bool(fast)), E.data_ptr(), _stream()), "pr_solve_666_device")
    return E[:n].cpu().numpy().reshape(-1, 4, 4)


def pcd2ab(pts, scene):
    """One transform_reduce of thrust__pcd2Ab (icp.cu:170-172) over a cloud -> 29 floats (host); reference-arithmetic kernel."""
    _require_device()
    pts = _dev(pts, torch.float32).reshape(-1, 3)
    out = torch.zeros(32, dtype=torch.float32, device="cuda")
    sc = scene.c()
    fn = lib().pr_pcd2ab_projective if isinstance(scene, SceneProjective) else lib().pr_pcd2ab_nn
    check(fn(pts.data_ptr(), pts.shape[0], C.byref(sc), out.data_ptr(), _stream()), "pr_pcd2ab")
    return out[:29].cpu().numpy()


# ---------------------------------------------------------------------------------------------
class PoseRefiner:
    """render -> depth2cloud -> ICP for a batch of pose hypotheses behind one call (pr_refiner)."""

    def __init__(self, tris, width, height, K, max_hyp, capacity_points=0):
        _require_device()
        tris = _f32c(tris).reshape(-1, 9)
        self.K = _f32c(K).reshape(9)
        self.width, self.height, self.max_hyp = width, height, max_hyp
        self._h = C.c_void_p()
        check(lib().pr_refiner_create(C.byref(self._h), tris.ctypes.data, tris.shape[0], width, height,
                                      self.K.ctypes.data, max_hyp, capacity_points), "pr_refiner_create")

    def close(self):
        if self._h:
            lib().pr_refiner_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_scene_projective(self, depth, max_dist_diff=0.1):
        d = np.ascontiguousarray(depth)
        assert d.dtype in (np.int32, np.uint16) and d.shape == (self.height, self.width)
        check(lib().pr_refiner_set_scene_projective(self._h, d.ctypes.data, int(d.dtype == np.int32), max_dist_diff),
              "pr_refiner_set_scene_projective")

    def set_scene_nn(self, depth):
        d = np.ascontiguousarray(depth)
        assert d.dtype in (np.int32, np.uint16) and d.shape == (self.height, self.width)
        check(lib().pr_refiner_set_scene_nn(self._h, d.ctypes.data, int(d.dtype == np.int32)), "pr_refiner_set_scene_nn")

    def set_scene_projective_device(self, depth_dev, max_dist_diff=0.1):
        """Scene from a cuda tensor [H,W] int32 / uint16 (no host copy; asynchronous on the current stream)."""
        assert depth_dev.is_cuda and depth_dev.dtype in (torch.int32, torch.uint16) and tuple(depth_dev.shape) == (self.height, self.width)
        depth_dev = depth_dev.contiguous()
        check(lib().pr_refiner_set_scene_projective_device(self._h, depth_dev.data_ptr(), int(depth_dev.dtype == torch.int32),
                                                           max_dist_diff, _stream()), "pr_refiner_set_scene_projective_device")

    def set_scene_nn_device(self, depth_dev):
        assert depth_dev.is_cuda and depth_dev.dtype in (torch.int32, torch.uint16) and tuple(depth_dev.shape) == (self.height, self.width)
        depth_dev = depth_dev.contiguous()
        check(lib().pr_refiner_set_scene_nn_device(self._h, depth_dev.data_ptr(), int(depth_dev.dtype == torch.int32), _stream()),
              "pr_refiner_set_scene_nn_device")

    def run(self, poses_host, criteria=None, results_host=None):
        """poses_host: [P,4,4] float32 (numpy or pinned torch) -> results [P,18] host. H2D + D2H inside."""
        criteria = criteria or ICPConvergenceCriteria()
        if isinstance(poses_host, torch.Tensor):
            assert not poses_host.is_cuda and poses_host.dtype == torch.float32 and poses_host.is_contiguous()
            P, pptr = poses_host.shape[0], poses_host.data_ptr()
        else:
            poses_host = _f32c(poses_host).reshape(-1, 16)
            P, pptr = poses_host.shape[0], poses_host.ctypes.data
        if results_host is None:
            results_host = np.zeros((P, 18), np.float32)
        rptr = results_host.data_ptr() if isinstance(results_host, torch.Tensor) else results_host.ctypes.data
        check(lib().pr_refiner_run(self._h, pptr, P, criteria.c(), rptr, _stream()), "pr_refiner_run")
        return results_host

    def run_device(self, poses_dev, criteria=None, results_dev=None):
        """Everything resident: poses cuda [P,16] -> results cuda [P,18]; asynchronous."""
        criteria = criteria or ICPConvergenceCriteria()
        poses_dev = poses_dev.contiguous()
        P = poses_dev.shape[0]
        if results_dev is None:
            results_dev = torch.empty((P, 18), dtype=torch.float32, device="cuda")
        check(lib().pr_refiner_run_device(self._h, poses_dev.data_ptr(), P, criteria.c(), results_dev.data_ptr(), _stream()),
              "pr_refiner_run_device")
        return results_dev

    def overflowed(self):
        """True when a cloud of the last run did not fit capacity_points (pr_refiner_overflow_flag; synchronises)."""
        p = C.c_void_p()
        check(lib().pr_refiner_overflow_flag(self._h, C.byref(p)), "pr_refiner_overflow_flag")
        torch.cuda.synchronize()
        return bool(torch.as_tensor(_CudaView(p.value, (1,), "<i4"), device="cuda").item())

    def stage_ms(self):
        """(mean render->cloud ms, mean ICP ms, number of runs) of the runs since the last call, from the CUDA events the
        refiner records around its two stages."""
        a, b, n = C.c_float(), C.c_float(), C.c_uint32()
        check(lib().pr_refiner_stage_ms(self._h, C.byref(a), C.byref(b), C.byref(n)), "pr_refiner_stage_ms")
        return a.value, b.value, n.value

    def scene_buffers(self):
        """Views of the prepared projective scene: (pcd [W*H,3], normal [W*H,3]) float32 on the device."""
        p, n = C.c_void_p(), C.c_void_p()
        check(lib().pr_refiner_scene_buffers(self._h, C.byref(p), C.byref(n), None), "pr_refiner_scene_buffers")
        return _view(p.value, (self.width * self.height, 3), "<f4"), _view(n.value, (self.width * self.height, 3), "<f4")

    def results_device(self, n_hyp):
        """View of the device copy of the results the last run() (host variant) produced: [n_hyp, 18] float32."""
        r = C.c_void_p()
        check(lib().pr_refiner_scene_buffers(self._h, None, None, C.byref(r)), "pr_refiner_scene_buffers")
        return _view(r.value, (n_hyp, 18), "<f4")

    def buffers(self, n_hyp):
        """Views of the refiner's own device buffers after a run of n_hyp hypotheses:
        (depth [P,H,W] int32, pts [total,3] float32, offsets [P+1] int32, counts [P] int32)."""
        d, p, o, c = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().pr_refiner_buffers(self._h, C.byref(d), C.byref(p), C.byref(o), C.byref(c)), "pr_refiner_buffers")
        depth = _view(d.value, (n_hyp, self.height, self.width), "<i4")
        offsets = _view(o.value, (n_hyp + 1,), "<i4")
        counts = _view(c.value, (n_hyp,), "<i4")
        total = int(offsets[n_hyp].item())
        pts = _view(p.value, (max(total, 1), 3), "<f4")
        return depth, pts, offsets, counts
