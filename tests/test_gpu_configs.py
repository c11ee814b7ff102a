"""The other BASELINE.json configurations at sizes the CPU oracle finishes in seconds:
C3 (kd-tree scene of ~100k points), C4 (1280x720 scene and clouds of ~100k points), C5 (49,920-triangle sphere,
uniformly random poses).  Same bars as test_gpu_parity.py."""
import numpy as np
import pytest

from pose_refine_b200 import workloads as wl
from test_gpu_parity import assert_result_close, REL_TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available()
    from pose_refine_b200 import api as a
    return a


def test_c5_sphere_render_bit_exact(api, port):
    import torch
    tris = wl.uv_sphere()
    assert tris.shape == (49920, 9)
    K = wl.LINEMOD_K
    proj = api.compute_proj(K, 640, 480)
    poses = wl.shoemake_poses(12, seed=99)
    want = port.render(tris, poses, 640, 480, proj)
    got = api.render_cuda_keep_in_gpu(tris, poses, 640, 480, proj)
    assert np.array_equal(got.cpu().numpy(), want)
    assert (want > 0).sum() > 12 * 5000
    # a larger batch: both raster paths agree image for image, and the batch is deterministic
    many = wl.shoemake_poses(256, seed=7)
    a = api.render_cuda_keep_in_gpu(tris, many, 640, 480, proj, use_tiles=True)
    b = api.render_cuda_keep_in_gpu(tris, many, 640, 480, proj, use_tiles=False)
    assert torch.equal(a, b)


def test_c4_1280x720_projective(api, port, mesh):
    import torch
    K = wl.k_1280x720()
    W, H = 1280, 720
    proj = api.compute_proj(K, W, H)
    assert np.array_equal(proj, port.compute_proj(K, W, H))
    _, scene_pose = wl.fixture_poses()
    scene_depth = port.render(mesh, scene_pose[None], W, H, proj)[0]
    assert np.array_equal(api.render_cuda(mesh, scene_pose[None], W, H, proj)[0], scene_depth)
    poses = wl.hypotheses(3, seed=4321, scene_pose=scene_pose, max_angle_deg=4.0, max_shift_mm=8.0)
    depth = api.render_cuda_keep_in_gpu(mesh, poses, W, H, proj)
    want_depth = port.render(mesh, poses, W, H, proj)
    assert np.array_equal(depth.cpu().numpy(), want_depth)
    pts, offsets, counts = api.depth2cloud_batch(depth, K)
    assert int(counts.min()) > 80000          # ~4x the 640x480 clouds
    scene = api.SceneProjective().init_cuda(scene_depth, K, W, H)
    ps = port.scene_projective(scene_depth, K)
    wp, wn, _ = ps.arrays()
    assert np.array_equal(scene.pcd.cpu().numpy(), wp) and np.array_equal(scene.normal.cpu().numpy(), wn)
    res = api.icp_batch(pts, offsets, counts, scene, api.ICPConvergenceCriteria(0.0, 0.0, 30)).cpu().numpy()
    port.set_threads(8)
    for i in range(3):
        cloud = port.depth2cloud(want_depth[i], K)
        runs = [port.icp(ps, cloud, 0.0, 0.0, 30)["raw"]]
        port.set_threads(3); runs.append(port.icp(ps, cloud, 0.0, 0.0, 30)["raw"]); port.set_threads(8)
        spread = np.abs(runs[0][:16] - runs[1][:16]).max()
        err = np.abs(res[i, :16] - runs[0][:16]).max()
        assert err <= max(REL_TOL, 3 * spread), (i, err, spread)
    port.set_threads(1)


def test_c3_kdtree_scene_100k(api, port, mesh, fixture_scene, golden):
    import torch
    arrays, _ = golden
    K = arrays["K"]
    scene_depth = wl.plane_scene_depth(fixture_scene["scene_depth"], target_valid=100000)
    assert 95000 < (scene_depth > 0).sum() < 105000
    scene = api.SceneNN().init_cuda(scene_depth, K)
    pn = port.scene_nn(scene_depth, K)
    wp, wn, wnodes = pn.arrays()
    assert np.array_equal(scene.pcd.cpu().numpy(), wp) and scene.nodes_host.tobytes() == wnodes.tobytes()
    poses = wl.hypotheses(2, seed=1234)
    depth = api.render_cuda_keep_in_gpu(mesh, poses, 640, 480, arrays["proj"])
    pts, offsets, counts = api.depth2cloud_batch(depth, K)
    res = api.icp_batch(pts, offsets, counts, scene, api.ICPConvergenceCriteria(0.0, 0.0, 30)).cpu().numpy()
    port.set_threads(8)
    h_pts, h_off, h_cnt = pts.cpu().numpy(), offsets.cpu().numpy(), counts.cpu().numpy()
    for i in range(2):
        want = port.icp(pn, h_pts[h_off[i]: h_off[i] + h_cnt[i]], 0.0, 0.0, 30)["raw"]
        assert_result_close(res[i], want, f"C3 hyp {i}")
    port.set_threads(1)
    # one pass of sums through both kernels: the reference-arithmetic one (its stackless walk) and the shipped one
    # (packed tree, child-box pruning) count the same inliers as the oracle on this 100k-point tree
    ref = port.pcd2ab(pn, h_pts[: h_cnt[0]]).astype(np.float64)
    got = api.pcd2ab(h_pts[: h_cnt[0]], scene).astype(np.float64)
    assert got[28] == ref[28]
    shipped = api.pass_sums(pts, offsets, counts, scene).astype(np.float64)
    assert shipped[0, 28] == ref[28]
    from test_gpu_parity import _terms_f64
    q, n, valid = port.query(pn, h_pts[: h_cnt[0]])
    terms = _terms_f64(h_pts[: h_cnt[0]], q, n, valid)
    truth, mag = terms.sum(0), np.abs(terms).sum(0)
    assert np.all(np.abs(shipped[0] - truth) <= 2e-6 * mag + 1e-30)
