"""The CPU restatement (oracle/oracle.cpp) against the golden vectors that scripts/make_golden.py
recorded from the reference's own CPU build.  Integer outputs and single-thread float outputs must
match bit for bit (compared through CRC32 of the raw bytes, or exactly)."""
import numpy as np

from conftest import crc
from pose_refine_b200 import workloads as wl


def test_mesh_and_proj(port, mesh, golden):
    arrays, scal = golden
    assert len(mesh) == scal["n_tris"] == 31468
    assert crc(mesh) == scal["tris_crc"]
    assert np.array_equal(port.compute_proj(arrays["K"], 640, 480), arrays["proj"])
    assert np.array_equal(port.compute_proj(arrays["K_small"], 161, 121), arrays["proj_small"])
    p1, p2 = wl.fixture_poses()
    assert np.array_equal(np.stack([p1, p2]), arrays["poses"])
    assert np.array_equal(wl.hypotheses(8, seed=1234), arrays["hyp8"])


def test_render_fixture(fixture_scene, golden):
    _, scal = golden
    for d, g in zip(fixture_scene["depth"], scal["render"]):
        assert int((d > 0).sum()) == g["valid"]
        assert int(d.sum()) == g["sum"]
        assert crc(d) == g["crc"]
        assert int(d[240, 320]) == g["center"]
    # SURVEY.md App. C known answers
    assert scal["render"][0]["valid"] == 26210 and scal["render"][1]["valid"] == 21960
    assert scal["render"][0]["crc"] == 0x6ff28ca9 and scal["render"][1]["crc"] == 0x99d44aec


def test_render_roi_small_near_batch(port, mesh, golden):
    arrays, scal = golden
    d = port.render(mesh, arrays["poses"], 640, 480, arrays["proj"], wl.ROI_FIXTURE)
    assert [crc(x) for x in d] == [g["crc"] for g in scal["render_roi"]]
    d = port.render(mesh, arrays["poses"], 161, 121, arrays["proj_small"])
    assert [crc(x) for x in d] == [g["crc"] for g in scal["render_small"]]
    d = port.render(mesh, arrays["poses"], 161, 121, arrays["proj_small"], (33, 17, 71, 53))
    assert [crc(x) for x in d] == [g["crc"] for g in scal["render_small_roi"]]
    d = port.render(mesh, arrays["pose_near"][None], 640, 480, arrays["proj"])
    assert crc(d[0]) == scal["render_near"]["crc"] and int(d.min()) == scal["render_near"]["min"]
    d = port.render(mesh, arrays["hyp8"], 640, 480, arrays["proj"])
    assert [crc(x) for x in d] == [g["crc"] for g in scal["render_hyp8"]]


def test_depth2cloud(port, fixture_scene, mesh, golden):
    arrays, scal = golden
    c = fixture_scene["cloud"]
    assert len(c) == scal["cloud"]["n"] == 26210 and crc(c) == scal["cloud"]["crc"]
    droi = port.render(mesh, arrays["poses"], 640, 480, arrays["proj"], wl.ROI_FIXTURE)
    c = port.depth2cloud(droi[0], arrays["K"], 1, wl.ROI_FIXTURE[0], wl.ROI_FIXTURE[1])
    assert len(c) == scal["cloud_roi_tl"]["n"] and crc(c) == scal["cloud_roi_tl"]["crc"]
    c = port.depth2cloud(fixture_scene["depth"][1].astype(np.uint16), arrays["K"])
    assert len(c) == scal["cloud_u16"]["n"] and crc(c) == scal["cloud_u16"]["crc"]


def test_scenes(port, fixture_scene, golden):
    arrays, scal = golden
    sd, K = fixture_scene["scene_depth"], arrays["K"]
    n = port.get_normal(sd, K)
    assert crc(n) == scal["normals"]["crc"] and int((np.abs(n).sum(-1) > 0).sum()) == scal["normals"]["nonzero"] == 21958
    sp = port.scene_projective(sd, K)
    pcd, nrm, _ = sp.arrays()
    assert crc(pcd) == scal["scene_projective"]["pcd_crc"] and crc(nrm) == scal["scene_projective"]["normal_crc"]
    sn = port.scene_nn(sd, K)
    pcd, nrm, nodes = sn.arrays()
    g = scal["scene_nn"]
    assert (len(pcd), len(nodes)) == (g["n_pts"], g["n_nodes"]) == (21960, 6287)
    assert crc(pcd) == g["pcd_crc"] and crc(nrm) == g["normal_crc"] and crc(nodes) == g["nodes_crc"]
    assert int((nodes["child1"] < 0).sum()) == g["n_leaves"] == 3144


def test_pcd2ab_query_solver(port, fixture_scene, golden):
    arrays, scal = golden
    sd, K, cloud = fixture_scene["scene_depth"], arrays["K"], fixture_scene["cloud"]
    sp, sn = port.scene_projective(sd, K), port.scene_nn(sd, K)
    assert np.array_equal(port.pcd2ab(sp, cloud), arrays["pcd2ab_projective"])
    assert np.array_equal(port.pcd2ab(sn, cloud), arrays["pcd2ab_nn"])
    assert int(port.query(sp, cloud)[2].sum()) == scal["query_projective_valid"]
    assert int(port.query(sn, cloud)[2].sum()) == scal["query_nn_valid"]
    for A, b, T in zip(arrays["solve_A"], arrays["solve_b"], arrays["solve_T"]):
        assert np.array_equal(port.solve_666(A, b), T)


def test_icp(port, fixture_scene, golden):
    arrays, _ = golden
    sd, K, cloud = fixture_scene["scene_depth"], arrays["K"], fixture_scene["cloud"]
    sp, sn = port.scene_projective(sd, K), port.scene_nn(sd, K)
    r = port.icp(sp, cloud, 0.0, 0.0, 30)
    assert np.array_equal(r["raw"], arrays["icp_projective_fixed30"]) and r["last_pass"] == 30
    r = port.icp(sp, cloud, 0.0, 0.0, 3)
    assert np.array_equal(r["raw"], arrays["icp_projective_fixed3"]) and r["last_pass"] == 3
    r = port.icp(sp, cloud)
    assert np.array_equal(r["raw"], arrays["icp_projective_default"]) and r["last_pass"] == 9
    r = port.icp(sn, cloud, 0.0, 0.0, 30)
    assert np.array_equal(r["raw"], arrays["icp_nn_fixed30"])
    r = port.icp(sn, cloud)
    assert np.array_equal(r["raw"], arrays["icp_nn_default"]) and r["last_pass"] == 6


def test_icp_hyp8(port, mesh, golden):
    arrays, _ = golden
    K = arrays["K"]
    d = port.render(mesh, arrays["hyp8"], 640, 480, arrays["proj"])
    scene_depth = port.render(mesh, arrays["poses"][1:2], 640, 480, arrays["proj"])[0]
    sp = port.scene_projective(scene_depth, K)
    for i in range(8):
        c = port.depth2cloud(d[i], K)
        assert len(c) == arrays["hyp8_counts"][i]
        assert np.array_equal(port.icp(sp, c, 0.0, 0.0, 30)["raw"], arrays["icp_hyp8_projective_fixed30"][i])


def test_icp_edge_cases(port, fixture_scene, golden):
    arrays, _ = golden
    sp = port.scene_projective(fixture_scene["scene_depth"], arrays["K"])
    # empty cloud and a cloud that never overlaps the scene: count == 0 on the first pass -> identity, 0, 0
    for pts in (np.zeros((0, 3), np.float32), np.tile(np.array([[5.0, 5.0, 1.0]], np.float32), (16, 1))):
        r = port.icp(sp, pts)
        assert np.array_equal(r["T"], np.eye(4, dtype=np.float32)) and r["rmse"] == 0 and r["fitness"] == 0 and r["last_pass"] == 0
    # max_iteration = 0: one evaluation pass, no update
    r = port.icp(sp, fixture_scene["cloud"], 0.0, 0.0, 0)
    assert np.array_equal(r["T"], np.eye(4, dtype=np.float32)) and 0.3 < r["fitness"] < 0.6
    assert port.depth2cloud(np.zeros((480, 640), np.int32), arrays["K"]).shape == (0, 3)
