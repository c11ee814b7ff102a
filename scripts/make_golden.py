"""Generate tests/golden/ from the reference's own CPU code (oracle/_ref, built by oracle/Makefile).

Runs only where /root/reference is mounted (the build container).  Everything written here is an
OUTPUT of the reference (or a re-encoding of its test asset), never its source:

  tests/golden/obj_06_mesh.npz   the reference's test mesh test/obj_06.ply re-encoded as float32
                                  vertices + int32 faces (the GPU box has no /root/reference)
  tests/golden/fixture.npz       arrays: poses, projection, ICP results, 29-sums, solver vectors ...
  tests/golden/fixture.json      scalars: counts, sums, CRC32s of whole images / clouds / trees

    python scripts/make_golden.py
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding  # noqa: E402
from pose_refine_b200 import workloads as wl  # noqa: E402

REF_PLY = "/root/reference/test/obj_06.ply"
OUT = os.path.join(ROOT, "tests", "golden")


def crc(a):
    return int(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = binding.load("reference")
    ref.set_threads(1)  # sequential float sums: bit-reproducible

    # ---- mesh ---------------------------------------------------------------------------------
    tris = ref.load_model(REF_PLY)
    faces = []
    with open(REF_PLY) as f:
        n_vert = n_face = 0
        for line in f:
            if line.startswith("element vertex"):
                n_vert = int(line.split()[2])
            if line.startswith("element face"):
                n_face = int(line.split()[2])
            if line.startswith("end_header"):
                break
        for _ in range(n_vert):
            f.readline()
        for _ in range(n_face):
            p = f.readline().split()
            assert p[0] == "3"
            faces.append([int(p[1]), int(p[2]), int(p[3])])
    faces = np.asarray(faces, np.int32)
    verts = np.zeros((n_vert, 3), np.float32)
    verts[faces.reshape(-1)] = tris.reshape(-1, 3)
    assert np.array_equal(verts[faces.reshape(-1)].reshape(-1, 9), tris)
    np.savez_compressed(os.path.join(OUT, "obj_06_mesh.npz"), vertices=verts, faces=faces)

    W, H = 640, 480
    K = wl.LINEMOD_K
    proj = ref.compute_proj(K, W, H)
    pose1, pose2 = wl.fixture_poses()
    poses = np.stack([pose1, pose2])
    arrays = {"K": K, "proj": proj, "poses": poses}
    scal = {"n_tris": int(len(tris)), "tris_crc": crc(tris)}

    # ---- renderer -------------------------------------------------------------------------------
    depth = ref.render(tris, poses, W, H, proj)
    scal["render"] = [{"valid": int((d > 0).sum()), "sum": int(d.sum()), "min": int(d[d > 0].min()), "max": int(d.max()),
                       "crc": crc(d), "center": int(d[240, 320])} for d in depth]
    droi = ref.render(tris, poses, W, H, proj, wl.ROI_FIXTURE)
    scal["render_roi"] = [{"valid": int((d > 0).sum()), "crc": crc(d)} for d in droi]
    hyp = wl.hypotheses(8, seed=1234)
    arrays["hyp8"] = hyp
    dh = ref.render(tris, hyp, W, H, proj)
    scal["render_hyp8"] = [{"valid": int((d > 0).sum()), "crc": crc(d)} for d in dh]
    # a pose that pushes part of the object behind / onto the camera plane (no near clipping upstream)
    near = pose1.copy(); near[2, 3] = 40.0
    arrays["pose_near"] = near
    dn = ref.render(tris, near[None], W, H, proj)
    scal["render_near"] = {"valid": int((dn[0] != 0).sum()), "crc": crc(dn[0]), "min": int(dn.min())}
    # small odd-sized image + ROI
    Ks = K.copy(); Ks[0, 2] = 80.3; Ks[1, 2] = 60.1; Ks[0, 0] = 143.1; Ks[1, 1] = 143.4
    arrays["K_small"] = Ks
    projs = ref.compute_proj(Ks, 161, 121)
    arrays["proj_small"] = projs
    ds = ref.render(tris, poses, 161, 121, projs)
    scal["render_small"] = [{"valid": int((d > 0).sum()), "crc": crc(d)} for d in ds]
    dsr = ref.render(tris, poses, 161, 121, projs, (33, 17, 71, 53))
    scal["render_small_roi"] = [{"valid": int((d > 0).sum()), "crc": crc(d)} for d in dsr]

    # ---- depth2cloud ----------------------------------------------------------------------------
    cloud = ref.depth2cloud(depth[0], K)
    scal["cloud"] = {"n": int(len(cloud)), "crc": crc(cloud)}
    cloud_tl = ref.depth2cloud(droi[0], K, 1, wl.ROI_FIXTURE[0], wl.ROI_FIXTURE[1])
    scal["cloud_roi_tl"] = {"n": int(len(cloud_tl)), "crc": crc(cloud_tl)}
    cu16 = ref.depth2cloud(depth[1].astype(np.uint16), K)
    scal["cloud_u16"] = {"n": int(len(cu16)), "crc": crc(cu16)}

    # ---- scenes ---------------------------------------------------------------------------------
    scene_depth = depth[1]
    nrm = ref.get_normal(scene_depth, K)
    scal["normals"] = {"crc": crc(nrm), "nonzero": int((np.abs(nrm).sum(-1) > 0).sum())}
    sp = ref.scene_projective(scene_depth, K)
    pcd_o, nrm_o, _ = sp.arrays()
    scal["scene_projective"] = {"pcd_crc": crc(pcd_o), "normal_crc": crc(nrm_o)}
    sn = ref.scene_nn(scene_depth, K)
    pcd_t, nrm_t, nodes = sn.arrays()
    scal["scene_nn"] = {"n_pts": int(len(pcd_t)), "n_nodes": int(len(nodes)), "pcd_crc": crc(pcd_t), "normal_crc": crc(nrm_t),
                        "nodes_crc": crc(nodes), "n_leaves": int((nodes["child1"] < 0).sum())}

    # ---- ICP pieces -------------------------------------------------------------------------------
    arrays["pcd2ab_projective"] = ref.pcd2ab(sp, cloud)
    arrays["pcd2ab_nn"] = ref.pcd2ab(sn, cloud)
    _, _, v = ref.query(sp, cloud)
    scal["query_projective_valid"] = int(v.sum())
    _, _, v = ref.query(sn, cloud)
    scal["query_nn_valid"] = int(v.sum())

    rng = np.random.RandomState(7)
    As, bs, Ts = [], [], []
    for i in range(6):
        J = rng.normal(size=(40, 6)).astype(np.float32) * np.array([0.2, 0.2, 0.2, 1, 1, 1], np.float32)
        r = (rng.normal(size=40) * 0.01).astype(np.float32)
        A = (J.T @ J).astype(np.float32); A = ((A + A.T) / 2).astype(np.float32)
        b = (J.T @ r).astype(np.float32)
        As.append(A); bs.append(b); Ts.append(ref.solve_666(A, b))
    # the real system of the first fixture pass
    S = arrays["pcd2ab_projective"]
    A0 = np.zeros((6, 6), np.float32); k = 0
    for y in range(6):
        for x in range(y, 6):
            A0[y, x] = A0[x, y] = S[k]; k += 1
    As.append(A0); bs.append(S[21:27].copy()); Ts.append(ref.solve_666(A0, S[21:27]))
    arrays["solve_A"], arrays["solve_b"], arrays["solve_T"] = np.stack(As), np.stack(bs), np.stack(Ts)

    def icp_pack(r):
        return r["raw"].copy()

    arrays["icp_projective_fixed30"] = icp_pack(ref.icp(sp, cloud, 0.0, 0.0, 30))
    arrays["icp_projective_default"] = icp_pack(ref.icp(sp, cloud))
    arrays["icp_projective_fixed3"] = icp_pack(ref.icp(sp, cloud, 0.0, 0.0, 3))
    arrays["icp_nn_fixed30"] = icp_pack(ref.icp(sn, cloud, 0.0, 0.0, 30))
    arrays["icp_nn_default"] = icp_pack(ref.icp(sn, cloud))
    # the 8-hypothesis batch: per-hypothesis cloud sizes and projective results at (0,0,30)
    res8, n8 = [], []
    for d in dh:
        c = ref.depth2cloud(d, K)
        n8.append(len(c))
        res8.append(icp_pack(ref.icp(sp, c, 0.0, 0.0, 30)))
    arrays["icp_hyp8_projective_fixed30"] = np.stack(res8)
    arrays["hyp8_counts"] = np.asarray(n8, np.int32)

    np.savez_compressed(os.path.join(OUT, "fixture.npz"), **arrays)
    with open(os.path.join(OUT, "fixture.json"), "w") as f:
        json.dump(scal, f, indent=1, sort_keys=True)
    print(json.dumps(scal, indent=1)[:1500])
    for k in ("icp_projective_fixed30", "icp_nn_fixed30", "icp_projective_default", "icp_nn_default"):
        print(k, arrays[k])


if __name__ == "__main__":
    main()
