"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d): fixtures, hypothesis batches, scenes.

Pure numpy; shared by tests/ and bench.py.  Constants are the reference's own test fixtures
(test.cpp:22-46, cuda_renderer/test.cpp:51-62, helper.h:187-209).
"""
import numpy as np

# K "from hinter dataset" (test.cpp:26, helper.h:41)
LINEMOD_K = np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]], np.float32)
# test.cpp:29-32
R_REN = np.array([[0.34768538, 0.93761126, 0.00000000],
                  [0.70540612, -0.26157897, -0.65877056],
                  [-0.61767070, 0.22904489, -0.75234390]], np.float32)
T_REN = np.array([0.0, 0.0, 300.0], np.float32)
T_REN2 = np.array([20.0, 20.0, 320.0], np.float32)
ROI_FIXTURE = (160, 80, 320, 240)  # cuda_renderer/test.cpp:122


def euler_zyx(theta):
    """helper::eulerAnglesToRotationMatrix (helper.h:187-209): R = Rz * Ry * Rx, float32."""
    t = np.asarray(theta, np.float32)
    c, s = np.cos(t), np.sin(t)
    Rx = np.array([[1, 0, 0], [0, c[0], -s[0]], [0, s[0], c[0]]], np.float32)
    Ry = np.array([[c[1], 0, s[1]], [0, 1, 0], [-s[1], 0, c[1]]], np.float32)
    Rz = np.array([[c[2], -s[2], 0], [s[2], c[2], 0], [0, 0, 1]], np.float32)
    return (Rz @ Ry @ Rx).astype(np.float32)


def pose44(R, t):
    m = np.eye(4, dtype=np.float32)
    m[:3, :3] = R
    m[:3, 3] = t
    return m


def fixture_poses():
    """(model pose, scene pose) of test.cpp:29-46: the scene is the model rotated by 10/180*3.14 rad
    about all three axes and moved to (20, 20, 320) mm."""
    a = np.float32(10.0) / np.float32(180.0) * np.float32(3.14)
    R2 = (euler_zyx([a, a, a]) @ R_REN).astype(np.float32)
    return pose44(R_REN, T_REN), pose44(R2, T_REN2)


def hypotheses(n, seed=1234, scene_pose=None, max_angle_deg=10.0, max_shift_mm=20.0):
    """R_i = Rzyx(ax,ay,az) * R_scene, a ~ U(-10,10) deg; t_i = t_scene + U(-20,20) mm.
    MT19937 (numpy RandomState) stands in for std::mt19937 -- only determinism matters here."""
    if scene_pose is None:
        scene_pose = fixture_poses()[1]
    rng = np.random.RandomState(seed)
    out = np.zeros((n, 4, 4), np.float32)
    for i in range(n):
        ang = np.deg2rad(rng.uniform(-max_angle_deg, max_angle_deg, 3)).astype(np.float32)
        shift = rng.uniform(-max_shift_mm, max_shift_mm, 3).astype(np.float32)
        out[i] = pose44((euler_zyx(ang) @ scene_pose[:3, :3]).astype(np.float32), scene_pose[:3, 3] + shift)
    return out


def k_1280x720():
    """C4: LINEMOD K zoomed 2x (-> 1280x960) then centre-cropped to 720 rows."""
    K = LINEMOD_K.copy()
    K[0, 0] *= 2; K[1, 1] *= 2; K[0, 2] *= 2
    K[1, 2] = 2 * LINEMOD_K[1, 2] - 120
    return K


def plane_scene_depth(obj_depth, target_valid=100000, tol=0.01):
    """C3: object depth composited (per-pixel min of non-zero) over a tilted ground plane
    z(u,v) = 450 + 0.25 (v-240) + 0.05 (u-320) mm restricted to the rows that give about
    target_valid valid pixels.  obj_depth: [H,W] int32."""
    H, W = obj_depth.shape
    v, u = np.mgrid[0:H, 0:W]
    plane = np.round(450 + 0.25 * (v - 240) + 0.05 * (u - 320)).astype(np.int32)
    best = None
    for rows in range(1, H + 1):
        r0 = (H - rows) // 2
        d = obj_depth.copy()
        band = np.zeros_like(d)
        band[r0:r0 + rows] = plane[r0:r0 + rows]
        both = (d > 0) & (band > 0)
        d = np.where(both, np.minimum(d, band), np.maximum(d, band))
        n = int((d > 0).sum())
        best = d
        if n >= target_valid * (1 - tol):
            break
    return best.astype(np.int32)


def uv_sphere(radius=50.0, segments=160, rings=157):
    """C5 mesh: UV sphere, two triangles per quad, poles as fans -> 2*segments*(rings-1) triangles
    (49,920 for 160 x 157).  Returns [T,9] float32."""
    tris = []
    th = np.linspace(0, np.pi, rings + 1)
    ph = np.linspace(0, 2 * np.pi, segments, endpoint=False)

    def pt(i, j):
        j = j % segments
        return (radius * np.sin(th[i]) * np.cos(ph[j]), radius * np.sin(th[i]) * np.sin(ph[j]), radius * np.cos(th[i]))

    for j in range(segments):
        tris.append(pt(0, 0) + pt(1, j) + pt(1, j + 1))
        tris.append(pt(rings, 0) + pt(rings - 1, j + 1) + pt(rings - 1, j))
    for i in range(1, rings - 1):
        for j in range(segments):
            tris.append(pt(i, j) + pt(i + 1, j) + pt(i + 1, j + 1))
            tris.append(pt(i, j) + pt(i + 1, j + 1) + pt(i, j + 1))
    return np.asarray(tris, np.float32)


def shoemake_poses(n, seed=99):
    """C5 poses: uniform rotations (Shoemake), t = (U(-40,40), U(-30,30), U(250,400)) mm."""
    rng = np.random.RandomState(seed)
    out = np.zeros((n, 4, 4), np.float32)
    for i in range(n):
        u1, u2, u3 = rng.uniform(0, 1, 3)
        q = np.array([np.sqrt(1 - u1) * np.sin(2 * np.pi * u2), np.sqrt(1 - u1) * np.cos(2 * np.pi * u2),
                      np.sqrt(u1) * np.sin(2 * np.pi * u3), np.sqrt(u1) * np.cos(2 * np.pi * u3)])
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float32)
        t = np.array([rng.uniform(-40, 40), rng.uniform(-30, 30), rng.uniform(250, 400)], np.float32)
        out[i] = pose44(R, t)
    return out


def load_mesh_npz(path):
    """tests/golden/obj_06_mesh.npz (float32 vertices + int32 faces) -> [T,9] float32 triangles."""
    z = np.load(path)
    return z["vertices"][z["faces"].reshape(-1)].reshape(-1, 9).astype(np.float32)
