"""The CPU restatement against the reference's own CPU sources compiled verbatim (oracle/_ref),
function by function on fresh inputs (beyond the committed golden vectors).  Skipped where
oracle/_ref is absent."""
import numpy as np

from pose_refine_b200 import workloads as wl


def test_render_random_poses(port, ref, mesh, golden):
    arrays, _ = golden
    poses = wl.hypotheses(6, seed=77, max_angle_deg=40.0, max_shift_mm=60.0)
    assert np.array_equal(port.render(mesh, poses, 640, 480, arrays["proj"]), ref.render(mesh, poses, 640, 480, arrays["proj"]))
    roi = (100, 60, 333, 217)
    assert np.array_equal(port.render(mesh, poses, 640, 480, arrays["proj"], roi), ref.render(mesh, poses, 640, 480, arrays["proj"], roi))


def test_render_sphere_and_behind_camera(port, ref, golden):
    arrays, _ = golden
    tris = wl.uv_sphere(50.0, 40, 37)
    poses = wl.shoemake_poses(3, seed=5)
    poses[2, 2, 3] = 30.0   # camera inside the sphere: triangles behind / across the camera plane
    a, b = port.render(tris, poses, 640, 480, arrays["proj"]), ref.render(tris, poses, 640, 480, arrays["proj"])
    assert np.array_equal(a, b)
    assert (a[0] > 0).sum() > 1000


def test_scene_and_tree_on_composited_scene(port, ref, fixture_scene, golden):
    arrays, _ = golden
    K = arrays["K"]
    sd = wl.plane_scene_depth(fixture_scene["scene_depth"], target_valid=30000)
    for dt in (np.int32, np.uint16):
        d = sd.astype(dt)
        assert np.array_equal(port.get_normal(d, K), ref.get_normal(d, K))
        a, b = port.scene_projective(d, K).arrays(), ref.scene_projective(d, K).arrays()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        a, b = port.scene_nn(d, K).arrays(), ref.scene_nn(d, K).arrays()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2].tobytes() == b[2].tobytes()


def test_queries_and_sums(port, ref, fixture_scene, golden):
    arrays, _ = golden
    K, cloud = arrays["K"], fixture_scene["cloud"]
    rng = np.random.RandomState(3)
    pts = (cloud[::7] + rng.normal(scale=0.004, size=cloud[::7].shape)).astype(np.float32)
    pts[:5] = [[0, 0, 0], [1, 1, -1], [np.nan, 0, 1], [1e30, 0, 1e-30], [-0.2, -0.2, 0.3]]
    for mk in ("scene_projective", "scene_nn"):
        sa, sb = getattr(port, mk)(fixture_scene["scene_depth"], K), getattr(ref, mk)(fixture_scene["scene_depth"], K)
        qa, qb = port.query(sa, pts), ref.query(sb, pts)
        assert np.array_equal(qa[2], qb[2])
        assert np.array_equal(qa[0][qa[2]], qb[0][qb[2]]) and np.array_equal(qa[1][qa[2]], qb[1][qb[2]])
        good = pts[5:]
        assert np.array_equal(port.pcd2ab(sa, good), ref.pcd2ab(sb, good))


def test_solver_random(port, ref):
    rng = np.random.RandomState(11)
    for _ in range(200):
        J = rng.normal(size=(12, 6)).astype(np.float32) * rng.uniform(0.01, 3.0, 6).astype(np.float32)
        A = (J.T @ J).astype(np.float32)
        A = ((A + A.T) / 2).astype(np.float32)
        b = rng.normal(size=6).astype(np.float32)
        Ta, Tb = port.solve_666(A, b), ref.solve_666(A, b)
        assert np.allclose(Ta, Tb, rtol=0, atol=2e-7), (Ta, Tb)


def test_icp_random_hypotheses(port, ref, mesh, fixture_scene, golden):
    arrays, _ = golden
    K = arrays["K"]
    sa, sb = port.scene_projective(fixture_scene["scene_depth"], K), ref.scene_projective(fixture_scene["scene_depth"], K)
    poses = wl.hypotheses(3, seed=99)
    depth = port.render(mesh, poses, 640, 480, arrays["proj"])
    for d in depth:
        c = port.depth2cloud(d, K)
        for crit in ((0.0, 0.0, 30), (1e-5, 1e-5, 30)):
            ra, rb = port.icp(sa, c, *crit), ref.icp(sb, c, *crit)
            assert np.array_equal(ra["raw"], rb["raw"])
            assert np.array_equal(ra["pts"], rb["pts"])   # the in-place transformed cloud too


def test_pipeline_entry_points_agree(port, ref, mesh, fixture_scene, golden):
    arrays, _ = golden
    K = arrays["K"]
    sa, sb = port.scene_projective(fixture_scene["scene_depth"], K), ref.scene_projective(fixture_scene["scene_depth"], K)
    poses = wl.hypotheses(2, seed=5)
    _, ra, na = port.pipeline(sa, mesh, poses, 640, 480, arrays["proj"], K, 0.0, 0.0, 5, schedule=1)
    _, rb, nb = ref.pipeline(sb, mesh, poses, 640, 480, arrays["proj"], K, 0.0, 0.0, 5, schedule=1)
    assert np.array_equal(na, nb) and np.array_equal(ra, rb)
