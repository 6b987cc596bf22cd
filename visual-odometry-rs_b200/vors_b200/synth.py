"""Seeded synthetic RGB-D scenes in TUM conventions.

Camera 0 sits at the identity; the world is one textured plane n.X = d seen through a pinhole
camera with the TUM fr1 intrinsics (reference: src/dataset/tum_rgbd.rs:31-35, scaled with the image
width).  The texture is a band-limited sum of sinusoids in plane coordinates plus Gaussian pixel
noise, quantised to u8; depth is the ray-cast Z in metres x 5000 rounded to u16 (0 = unknown,
src/dataset/tum_rgbd.rs:15).  Camera k's pose is camera-to-world, which is what the reference's
Tracker reports (`current_frame_pose`, inverse_compositional.rs:206-208).

Two back-ends with the same maths: numpy (deterministic, used for parity tests and golden vectors)
and torch (used by bench.py to synthesise hundreds of frames on the GPU quickly).
"""
from __future__ import annotations

import dataclasses

import numpy as np

FR1 = dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, skew=0.0)
DEPTH_SCALE = 5000.0


@dataclasses.dataclass
class Scene:
    rows: int
    cols: int
    fx: float
    fy: float
    cx: float
    cy: float
    normal: np.ndarray  # unit plane normal (world = camera-0 frame)
    dist: float         # n.X = dist
    freqs: np.ndarray   # (K, 2) cycles per metre along the plane basis
    phases: np.ndarray  # (K,)
    amps: np.ndarray    # (K,)
    noise_sigma: float
    seed: int


def make_scene(seed: int = 0, rows: int = 480, cols: int = 640, n_waves: int = 24, noise_sigma: float = 2.0,
               dist: float = 2.0) -> Scene:
    rng = np.random.default_rng(seed)
    s = cols / 640.0
    n = np.array([0.1, -0.15, 1.0])
    n /= np.linalg.norm(n)
    # log-uniform wavelengths from ~1.2 m down to ~2 cm: texture at every pyramid level
    f = np.exp(rng.uniform(np.log(0.8), np.log(50.0), n_waves))
    th = rng.uniform(0, 2 * np.pi, n_waves)
    freqs = np.stack([f * np.cos(th), f * np.sin(th)], 1)
    phases = rng.uniform(0, 2 * np.pi, n_waves)
    amps = 1.0 / np.sqrt(f)
    amps *= 45.0 / np.sqrt(0.5 * np.sum(amps ** 2))  # texture std ~45 grey levels
    return Scene(rows, cols, FR1["fx"] * s, FR1["fy"] * s, (FR1["cx"] + 0.5) * s - 0.5, (FR1["cy"] + 0.5) * s - 0.5,
                 n, dist, freqs, phases, amps, noise_sigma, seed)


def _plane_basis(n):
    e1 = np.cross(n, [0.0, 1.0, 0.0])
    e1 /= np.linalg.norm(e1)
    e2 = np.cross(n, e1)
    return e1, e2


def quat_to_rot(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def se3_exp(xi):
    """Float64 exponential map, twist (v, w) -> (t, q xyzw); same formula as src/math/se3.rs:65-95."""
    xi = np.asarray(xi, np.float64)
    v, w = xi[:3], xi[3:]
    th2 = float(w @ w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th2 < 1e-12:
        re, im, c1, c2 = 1 - th2 / 8, 0.5 - th2 / 48, 0.5 - th2 / 24, 1 / 6 - th2 / 120
    else:
        th = np.sqrt(th2)
        re, im = np.cos(th / 2), np.sin(th / 2) / th
        c1, c2 = (1 - np.cos(th)) / th2, (th - np.sin(th)) / (th * th2)
    V = np.eye(3) + c1 * W + c2 * (W @ W)
    q = np.concatenate([im * w, [re]])
    q /= np.linalg.norm(q)
    return V @ v, q


def pose_mul(a, b):
    """Compose two (t, q) poses: a * b."""
    ta, qa = a
    tb, qb = b
    Ra = quat_to_rot(qa)
    x1, y1, z1, w1 = qa
    x2, y2, z2, w2 = qb
    q = np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                  w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
    return ta + Ra @ tb, q / np.linalg.norm(q)


def random_twist(rng, max_v=0.05, max_w=0.03):
    v = rng.normal(size=3)
    v *= rng.uniform(0.3, 1.0) * max_v / np.linalg.norm(v)
    w = rng.normal(size=3)
    w *= rng.uniform(0.3, 1.0) * max_w / np.linalg.norm(w)
    return np.concatenate([v, w])


def render(scene: Scene, pose=None, frame_index: int = 0, holes: int = 0):
    """Render (gray u8 [rows, cols], depth u16 [rows, cols]) seen from camera-to-world `pose` = (t, q xyzw)."""
    if pose is None:
        t, R = np.zeros(3), np.eye(3)
    else:
        t, R = np.asarray(pose[0], np.float64), quat_to_rot(pose[1])
    ys, xs = np.mgrid[0:scene.rows, 0:scene.cols].astype(np.float64)
    rc = np.stack([(xs - scene.cx) / scene.fx, (ys - scene.cy) / scene.fy, np.ones_like(xs)], -1)  # z = 1 rays
    rw = rc @ R.T
    denom = rw @ scene.normal
    lam = (scene.dist - scene.normal @ t) / np.where(np.abs(denom) < 1e-9, 1e-9, denom)
    Xw = t + lam[..., None] * rw
    e1, e2 = _plane_basis(scene.normal)
    p, q = Xw @ e1, Xw @ e2
    tex = np.full(p.shape, 128.0)
    for (fx_, fy_), ph, a in zip(scene.freqs, scene.phases, scene.amps):
        tex += a * np.sin(2 * np.pi * (fx_ * p + fy_ * q) + ph)
    rng = np.random.default_rng([scene.seed, 1, frame_index])
    tex += rng.normal(0.0, scene.noise_sigma, tex.shape)
    gray = np.clip(np.rint(tex), 0, 255).astype(np.uint8)
    z = lam  # camera-frame depth because rays have z = 1
    valid = (z > 0.1) & (z < 13.0)
    depth = np.where(valid, np.rint(z * DEPTH_SCALE), 0).astype(np.uint16)
    if holes:
        for _ in range(holes):
            r0 = int(rng.integers(0, scene.rows - 8))
            c0 = int(rng.integers(0, scene.cols - 8))
            depth[r0:r0 + int(rng.integers(4, scene.rows // 6 + 5)), c0:c0 + int(rng.integers(4, scene.cols // 6 + 5))] = 0
    return gray, depth


def make_pair(seed: int = 1000, rows: int = 480, cols: int = 640, max_v=0.05, max_w=0.03, holes: int = 0):
    """Config-1 style pair: (scene, (gray0, depth0), (gray1, depth1), pose1=(t, q))."""
    scene = make_scene(seed, rows, cols)
    rng = np.random.default_rng(seed + 2)
    pose1 = se3_exp(random_twist(rng, max_v, max_w))
    f0 = render(scene, None, 0, holes)
    f1 = render(scene, pose1, 1, holes)
    return scene, f0, f1, pose1


def trajectory(seed: int, n_frames: int, step_v=0.004, step_w=0.003):
    """Smooth camera-to-world trajectory: pose_0 = identity, pose_k = pose_{k-1} * exp(xi_k) with a slowly
    drifting body-frame twist (handheld-like: ~4 mm and ~0.17 deg per frame)."""
    rng = np.random.default_rng(seed + 3)
    xi = random_twist(rng, step_v, step_w)
    poses = [(np.zeros(3), np.array([0.0, 0.0, 0.0, 1.0]))]
    for _ in range(1, n_frames):
        xi = 0.9 * xi + 0.1 * random_twist(rng, step_v, step_w)
        poses.append(pose_mul(poses[-1], se3_exp(xi)))
    return poses


def make_sequence(seed: int, n_frames: int, rows: int = 480, cols: int = 640, **kw):
    scene = make_scene(seed, rows, cols)
    poses = trajectory(seed, n_frames, **kw)
    frames = [render(scene, p, k) for k, p in enumerate(poses)]
    return scene, frames, poses


def scene_config_kwargs(scene: Scene):
    return dict(fx=scene.fx, fy=scene.fy, cx=scene.cx, cy=scene.cy, skew=0.0, depth_scale=DEPTH_SCALE)


# ------------------------------------------------------------------------------------------
# torch back-end (bench only): renders a batch of frames on the GPU.

def render_batch_torch(scene: Scene, poses, device, frame_seed: int = 0, chunk: int = 16):
    """Same maths as render(), vectorised over frames on `device`.  poses: list of (t, q).
    Returns (gray u8 [B, rows, cols], depth int32 [B, rows, cols]) torch tensors (depth fits u16)."""
    import torch

    dt = torch.float64
    ys, xs = torch.meshgrid(torch.arange(scene.rows, dtype=dt, device=device),
                            torch.arange(scene.cols, dtype=dt, device=device), indexing="ij")
    rc = torch.stack([(xs - scene.cx) / scene.fx, (ys - scene.cy) / scene.fy, torch.ones_like(xs)], -1)  # r,c,3
    n = torch.tensor(scene.normal, dtype=dt, device=device)
    e1, e2 = _plane_basis(scene.normal)
    e1 = torch.tensor(e1, dtype=dt, device=device)
    e2 = torch.tensor(e2, dtype=dt, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(scene.seed) * 1000003 + frame_seed)
    grays, depths = [], []
    for b0 in range(0, len(poses), chunk):
        sub = poses[b0:b0 + chunk]
        t = torch.tensor(np.stack([p[0] for p in sub]), dtype=dt, device=device)  # b,3
        R = torch.tensor(np.stack([quat_to_rot(p[1]) for p in sub]), dtype=dt, device=device)  # b,3,3
        rw = torch.einsum("rck,bjk->brcj", rc, R)  # world-frame rays
        denom = rw @ n
        denom = torch.where(denom.abs() < 1e-9, torch.full_like(denom, 1e-9), denom)
        lam = (scene.dist - (t @ n))[:, None, None] / denom
        Xw = t[:, None, None, :] + lam[..., None] * rw
        p, q = Xw @ e1, Xw @ e2
        tex = torch.full_like(p, 128.0)
        for (fx_, fy_), ph, a in zip(scene.freqs, scene.phases, scene.amps):
            tex += float(a) * torch.sin(2 * np.pi * (float(fx_) * p + float(fy_) * q) + float(ph))
        tex += torch.randn(tex.shape, generator=gen, device=device, dtype=dt) * scene.noise_sigma
        grays.append(torch.clamp(torch.round(tex), 0, 255).to(torch.uint8))
        valid = (lam > 0.1) & (lam < 13.0)
        depths.append(torch.where(valid, torch.round(lam * DEPTH_SCALE), torch.zeros_like(lam)).to(torch.int32))
    return torch.cat(grays), torch.cat(depths)
