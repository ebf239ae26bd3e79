"""Synthetic YCB-shaped inputs for the volumetric-pose hot path (no datasets offline).

Shapes and conventions follow the reference's data pipeline:
  * 21 YCB classes, voxel pitch = bbox diagonal / 32 (datasets/ycb_video/models.py:113-115;
    table ros/src/morefusion_ros/include/morefusion_ros/utils/data.h:12-32)
  * origin = median(points) - pitch * 15.5 (model.py:197-205, rgbd_pose_estimation/base.py:153-156)
  * 1000 points per object, voxel-frame points = (cam - origin) / pitch (model.py:236)
  * grid_nontarget_empty: free space + other objects as seen from the camera (train.py:50-54)
Objects are boxes / cylinders / spheres whose bbox diagonal is 32 * pitch(class).
"""

import numpy as np

from .contrib.singleview_3d.models.model import YCB_VOXEL_PITCH_32

F32 = np.float32


def init_weights(n_fg_class=21, seed=0, with_occupancy=True):
    """Seeded LeCun-normal weights (chainer's default initialiser) and small random biases for
    the 3-D section of the pose model, keyed by the reference's link names (model.py:62-91):
    {'<link>/W': ndarray, '<link>/b': ndarray}.  No trained weights exist offline."""
    rs = np.random.RandomState(seed)
    w = {}

    def conv(name, cout, cin, *k):
        fan_in = cin * int(np.prod(k)) if k else cin
        shape = (cout, cin) + tuple(k)
        w[name + "/W"] = (rs.normal(0, 1.0 / np.sqrt(fan_in), shape)).astype(np.float32)
        w[name + "/b"] = rs.uniform(-0.05, 0.05, cout).astype(np.float32)

    conv("conv1_rgb", 64, 32, 1)
    conv("conv1_pcd", 8, 3, 1)
    conv("conv2_rgb", 128, 64, 1)
    conv("conv2_pcd", 16, 8, 1)
    if with_occupancy:
        conv("conv1_occ", 8, 1, 3, 3, 3)
        conv("conv2_occ", 16, 8, 3, 3, 3)
    cin3 = 144 + (16 if with_occupancy else 0)
    conv("conv3", 256, cin3, 4, 4, 4)
    conv("conv4", 512, 256, 4, 4, 4)
    for head, cout in (("rot", n_fg_class * 4), ("trans", n_fg_class * 3), ("conf", n_fg_class)):
        conv(f"conv1_{head}", 640, 984, 1)
        conv(f"conv2_{head}", 256, 640, 1)
        conv(f"conv3_{head}", 128, 256, 1)
        conv(f"conv4_{head}", cout, 128, 1)
    return w


def _rot(rs):
    q = rs.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _primitive(class_id, kinds=("box", "cylinder", "sphere")):
    """kind, half-extents (metres) with bbox diagonal = 32 * pitch."""
    diag = 32.0 * YCB_VOXEL_PITCH_32[class_id]
    kind = kinds[class_id % len(kinds)]
    if kind == "box":
        ratio = np.array([1.0, 0.7, 0.45])
    elif kind == "cylinder":
        ratio = np.array([0.6, 0.6, 1.0])
    else:
        ratio = np.array([1.0, 1.0, 1.0])
    half = ratio / np.linalg.norm(2 * ratio) * diag
    return kind, half


def sdf_primitive(kind, half, p):
    """Signed distance, POSITIVE INSIDE (trimesh convention, datasets/ycb_video/models.py:77)."""
    if kind == "sphere":
        return half[0] - np.linalg.norm(p, axis=-1)
    if kind == "box":
        q = np.abs(p) - half
        outside = np.linalg.norm(np.maximum(q, 0), axis=-1)
        inside = np.minimum(q.max(axis=-1), 0)
        return -(outside + inside)
    # cylinder along z
    dr = np.linalg.norm(p[..., :2], axis=-1) - half[0]
    dz = np.abs(p[..., 2]) - half[2]
    outside = np.linalg.norm(np.maximum(np.stack([dr, dz], -1), 0), axis=-1)
    inside = np.minimum(np.maximum(dr, dz), 0)
    return -(outside + inside)


def surface_points(kind, half, n, rs):
    """Uniform-ish points on the primitive's surface (object frame)."""
    if kind == "sphere":
        d = rs.normal(size=(n, 3))
        return d / np.linalg.norm(d, axis=1, keepdims=True) * half[0]
    if kind == "box":
        p = rs.uniform(-1, 1, (n, 3)) * half
        ax = rs.randint(0, 3, n)
        p[np.arange(n), ax] = np.sign(rs.uniform(-1, 1, n)) * half[ax]
        return p
    th = rs.uniform(0, 2 * np.pi, n)
    z = rs.uniform(-1, 1, n) * half[2]
    p = np.stack([np.cos(th) * half[0], np.sin(th) * half[0], z], 1)
    cap = rs.uniform(size=n) < 0.25
    r = np.sqrt(rs.uniform(size=n)) * half[0]
    p[cap] = np.stack([np.cos(th) * r, np.sin(th) * r, np.sign(rs.uniform(-1, 1, n)) * half[2]], 1)[cap]
    return p


class SyntheticYCBModels:
    """Stand-in for morefusion.datasets.YCBVideoModels (a multi-GB download, unavailable
    offline) with the three methods the pose model touches (model.py:199, :353, :416):
    YCB-shaped primitives whose bounding-box diagonal is 32 x the class voxel pitch."""

    # datasets/ycb_video/class_names.py:32-46: bowl, wood_block, large_clamp, extra_large_clamp,
    # foam_brick
    class_ids_symmetric = np.array([13, 16, 19, 20, 21], dtype=np.int32)

    def __init__(self, n_pcd=2000, kinds=("box", "cylinder", "sphere")):
        self._n_pcd = n_pcd
        self._kinds = kinds
        self._pcd = {}

    def get_voxel_pitch(self, dimension, class_id):
        return 32.0 * YCB_VOXEL_PITCH_32[int(class_id)] / dimension

    def get_pcd(self, class_id):
        class_id = int(class_id)
        if class_id not in self._pcd:
            kind, half = _primitive(class_id, self._kinds)
            rs = np.random.RandomState(class_id)
            self._pcd[class_id] = surface_points(kind, half, self._n_pcd, rs).astype(F32)
        return self._pcd[class_id]

    def get_sdf(self, class_id):
        return sdf_lattice(int(class_id), self._kinds)


def sdf_lattice(class_id, kinds=("box", "cylinder", "sphere")):
    """Stand-in for YCBVideoModels.get_sdf (models.py:66-79): interior lattice points at the
    class pitch with their signed distance (positive inside)."""
    kind, half = _primitive(class_id, kinds)
    pitch = YCB_VOXEL_PITCH_32[class_id]
    ax = [np.arange(-h, h + 1e-9, pitch) for h in half]
    g = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    d = sdf_primitive(kind, half, g)
    keep = d >= -0.5 * pitch
    return g[keep].astype(F32), d[keep].astype(F32)


def make_cnn_batch(B=8, P=1000, seed=0, D=32):
    """One batch of objects for the 3D-CNN hot path (BASELINE config 2 shape)."""
    rs = np.random.RandomState(seed)
    class_id = ((np.arange(B) + seed) % 21 + 1).astype(np.int32)
    values = rs.normal(0, 1, (B, 32, P)).astype(F32)      # stands in for PSPNet features
    points = np.zeros((B, 3, P), F32)
    pitch = np.zeros(B, F32)
    origin = np.zeros((B, 3), F32)
    gne = np.zeros((B, D, D, D), bool)
    for i in range(B):
        kind, half = _primitive(int(class_id[i]))
        pitch[i] = YCB_VOXEL_PITCH_32[int(class_id[i])]
        R = _rot(rs)
        t = np.array([rs.uniform(-0.2, 0.2), rs.uniform(-0.15, 0.15), rs.uniform(0.5, 0.8)])
        sp = surface_points(kind, half, 6 * P, rs)
        cam = sp @ R.T + t
        # visible side only: surface normal ~ (p - centre) facing the camera at the origin
        vis = ((cam - t) * (-cam)).sum(1) > 0
        cam = cam[vis]
        cam = cam + rs.normal(0, 0.003, cam.shape) * (cam / np.linalg.norm(cam, axis=1, keepdims=True))
        keep = rs.permutation(cam.shape[0])[:P]
        if keep.shape[0] < P:
            keep = np.r_[keep, rs.randint(0, cam.shape[0], P - keep.shape[0])]
        cam = cam[keep].astype(F32)
        origin[i] = (np.median(cam, axis=0) - pitch[i] * (D / 2.0 - 0.5)).astype(F32)
        points[i] = ((cam - origin[i]) / pitch[i]).T
        # free space in front of the surface along the viewing rays + a random occluder slab
        ijk = np.stack(np.meshgrid(*(np.arange(D),) * 3, indexing="ij"), -1).astype(F32)
        centres = ijk * pitch[i] + origin[i]
        obj = (centres - t) @ R
        inside = sdf_primitive(kind, half, obj) > 0
        front = np.linalg.norm(centres, axis=-1) < np.linalg.norm(t) - 0.3 * half.max()
        slab = np.abs(centres[..., 0] - (t[0] + half.max() * 1.2)) < 2 * pitch[i]
        gne[i] = (front | slab) & ~inside
    return dict(class_id=class_id, values=values, points=points, pitch=pitch, origin=origin,
                grid_nontarget_empty=gne)


CAMERA_K = np.array([[619.44, 0, 326.82], [0, 619.32, 239.52], [0, 0, 1.0]])   # pose_refinement/data/camera_info.yaml


def make_rgbd_batch(B=8, H=256, W=256, seed=0, D=32):
    """Synthetic stand-in for one batch of the RGB-D pose-estimation datasets
    (datasets/rgbd_pose_estimation/base.py:125-156): per object a centred H x W crop with
    ``rgb`` uint8, ``pcd`` float32 camera-frame points (NaN where there is no depth: background,
    5 % dropout), the ground-truth pose, class id, pitch, origin and a non-target/empty grid --
    the keyword arguments of Model.__call__ (model.py:277-288)."""
    rs = np.random.RandomState(seed)
    class_id = ((np.arange(B) * 3 + seed) % 21 + 1).astype(np.int32)
    rgb = rs.randint(0, 255, (B, H, W, 3)).astype(np.uint8)
    pcd = np.full((B, H, W, 3), np.nan, F32)
    q_true = np.zeros((B, 4), F32)
    t_true = np.zeros((B, 3), F32)
    pitch = np.zeros(B, F32)
    origin = np.zeros((B, 3), F32)
    gne = np.zeros((B, D, D, D), bool)
    f = CAMERA_K[0, 0] * 2.2           # crop magnification: the object fills most of the crop
    for i in range(B):
        kind, half = _primitive(int(class_id[i]))
        pitch[i] = YCB_VOXEL_PITCH_32[int(class_id[i])]
        R = _rot(rs)
        t = np.array([rs.uniform(-0.02, 0.02), rs.uniform(-0.02, 0.02), rs.uniform(0.55, 0.75)])
        sp = surface_points(kind, half, 60000, rs)
        cam = sp @ R.T + t
        cam = cam[((cam - t) * (-cam)).sum(1) > 0]                      # camera-facing side
        cam += rs.normal(0, 0.003, cam.shape) * (cam / np.linalg.norm(cam, axis=1, keepdims=True))
        u = np.round(f * cam[:, 0] / cam[:, 2] + W / 2 - f * t[0] / t[2]).astype(int)
        v = np.round(f * cam[:, 1] / cam[:, 2] + H / 2 - f * t[1] / t[2]).astype(int)
        ok = (u >= 0) & (u < W) & (v >= 0) & (v < H)
        order = np.argsort(-cam[ok, 2])                                 # nearest point wins
        uu, vv, cc = u[ok][order], v[ok][order], cam[ok][order]
        pcd[i, vv, uu] = cc.astype(F32)
        drop = rs.rand(H, W) < 0.05
        pcd[i][drop] = np.nan
        rgb[i][~np.isnan(pcd[i, :, :, 2])] //= 2
        # quaternion (w, x, y, z) of R
        tr = np.trace(R)
        w_ = np.sqrt(max(1 + tr, 1e-12)) / 2
        q = np.array([w_, (R[2, 1] - R[1, 2]) / (4 * w_), (R[0, 2] - R[2, 0]) / (4 * w_),
                      (R[1, 0] - R[0, 1]) / (4 * w_)])
        q_true[i] = (q / np.linalg.norm(q)).astype(F32)
        t_true[i] = t.astype(F32)
        valid = pcd[i][~np.isnan(pcd[i]).any(axis=2)]
        origin[i] = (np.median(valid, axis=0) - pitch[i] * (D / 2.0 - 0.5)).astype(F32)
        ijk = np.stack(np.meshgrid(*(np.arange(D),) * 3, indexing="ij"), -1).astype(F32)
        centres = ijk * pitch[i] + origin[i]
        inside = sdf_primitive(kind, half, (centres - t) @ R) > 0
        front = np.linalg.norm(centres, axis=-1) < np.linalg.norm(t) - 0.3 * half.max()
        gne[i] = front & ~inside
    return dict(class_id=class_id, rgb=rgb, pcd=pcd, quaternion_true=q_true, translation_true=t_true,
                pitch=pitch, origin=origin, grid_nontarget_empty=gne)


def make_icc_scene(N=8, seed=0, D=32, t_noise=0.01, rot_noise_deg=10.0,
                   kinds=("box", "cylinder", "sphere")):
    """Synthetic stand-in for examples/ycb_video/pose_refinement/data (BASELINE config 4):
    N YCB-shaped primitives resting in contact in a bin, SDF lattice points per object
    (models.get_sdf stand-in), ground-truth poses, perturbed initial poses
    (t ~ N(0, 1 cm), rotation ~ U(0, 10 deg)), per-object 32^3 target / non-target+empty grids."""
    rs = np.random.RandomState(seed)
    class_id = ((np.arange(N) * 5 + seed) % 21 + 1).astype(np.int32)
    prims = [_primitive(int(c), kinds) for c in class_id]
    # place objects on a jittered lattice so that neighbours touch
    T_true = []
    cols = int(np.ceil(np.sqrt(N)))
    for i, (kind, half) in enumerate(prims):
        R = _rot(rs)
        r = float(np.linalg.norm(half))
        cx = (i % cols - (cols - 1) / 2.0) * 0.085 + rs.uniform(-0.01, 0.01)
        cy = (i // cols - (cols - 1) / 2.0) * 0.085 + rs.uniform(-0.01, 0.01)
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = [cx, cy, 0.7 + rs.uniform(-0.02, 0.02) + 0.0 * r]
        T_true.append(T)
    T_true = np.stack(T_true)
    points, sdf, pitch, origin = [], [], [], []
    for i, c in enumerate(class_id):
        p, s = sdf_lattice(int(c), kinds)
        points.append(p)
        sdf.append(s)
        pitch.append(YCB_VOXEL_PITCH_32[int(c)])
        origin.append(T_true[i, :3, 3] - pitch[-1] * (D / 2.0 - 0.5))
    pitch = np.array(pitch, F32)
    origin = np.array(origin, F32)
    gt_grid = np.zeros((N, D, D, D), F32)
    gne = np.zeros((N, D, D, D), F32)
    ijk = np.stack(np.meshgrid(*(np.arange(D),) * 3, indexing="ij"), -1).astype(np.float64)
    for i in range(N):
        centres = ijk * pitch[i] + origin[i]
        d_all = []
        for j, (kind, half) in enumerate(prims):
            obj = (centres - T_true[j, :3, 3]) @ T_true[j, :3, :3]
            d_all.append(sdf_primitive(kind, half, obj))
        d_all = np.stack(d_all)
        d_self = d_all[i]
        # visible (camera-facing) surface shell of object i
        normal_out = centres - T_true[i, :3, 3]
        facing = (normal_out * (-centres)).sum(-1) > 0
        gt_grid[i] = ((np.abs(d_self) < 0.75 * pitch[i]) & facing)
        others = np.delete(d_all, i, 0).max(0) > 0 if N > 1 else np.zeros_like(d_self, bool)
        free = (d_all.max(0) < -1.5 * pitch[i]) & (centres[..., 2] < T_true[i, 2, 3])
        gne[i] = (others | free) & ~(d_self > 0)
    # perturbed initial poses
    T_init = []
    for i in range(N):
        ax = rs.normal(size=3)
        ax /= np.linalg.norm(ax)
        ang = np.deg2rad(rs.uniform(0, rot_noise_deg))
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        dR = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        T = T_true[i].copy()
        T[:3, :3] = T[:3, :3] @ dR
        T[:3, 3] += rs.normal(0, t_noise, 3)
        T_init.append(T)
    return dict(class_id=class_id, points=points, sdf=sdf, pitch=pitch, origin=origin,
                grid_target=gt_grid, grid_nontarget_empty=gne,
                transform_init=np.stack(T_init).astype(F32), transform_true=T_true.astype(F32),
                primitives=prims)


def make_depth_frame(seed=0, H=480, W=640, n_objects=8):
    """A 640x480 synthetic RGB-D frame for the occupancy-map producer (SURVEY.md 8f-3, cfg5
    geometry): `n_objects` YCB-sized boxes at 0.55-0.65 m in front of a tilted table at 0.9-1.1 m,
    camera intrinsics of examples/ycb_video/pose_refinement/data/camera_info.yaml scale (fx ~ 600),
    3 mm depth noise and 5 % dropout (datasets/rgbd_pose_estimation/reindexed.py:70-75).
    Returns (pcd [H,W,3] f32 with NaN rows, instance label [H,W] int32 (0 = background),
    class_ids [n_objects], pitches {instance_id: voxel pitch}); instance ids are 1..n_objects."""
    rs = np.random.RandomState(seed)
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    fx = 600.0
    z = 0.9 + 0.2 * (v / H)
    label = np.zeros((H, W), np.int32)
    class_ids = [1 + (seed + 3 * i) % 21 for i in range(n_objects)]
    pitches = {0: 0.01}
    for ins in range(1, n_objects + 1):
        cu = int(W * (ins - 0.5) / n_objects) + rs.randint(-8, 9)
        cv = H // 2 - 60 + 30 * (ins % 4)
        m = (np.abs(u - cu) < 28) & (np.abs(v - cv) < 36)
        z = np.where(m, 0.55 + 0.03 * (ins % 3) + 0.01 * ((u - cu) / 28.0) ** 2, z)
        label[m] = ins
        pitches[ins] = float(YCB_VOXEL_PITCH_32[class_ids[ins - 1]])
    z = z + 0.003 * rs.randn(H, W)
    pcd = np.stack([(u - W / 2) * z / fx, (v - H / 2) * z / fx, z], -1).astype(F32)
    pcd[rs.rand(H, W) < 0.05] = np.nan
    return pcd, label, class_ids, pitches
