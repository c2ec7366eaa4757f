"""Seeded synthetic stereo sequences: one slanted textured plane, ray-cast per pixel.

The same bytes are fed to the CUDA path and to the CPU oracle. Pure numpy (no OpenCV), deterministic.

Conventions follow the reference: pose = (t, r) with camera centre t and camera-to-world rotation R(r)
(src/include/pose_manager.hpp:9-20); the "right" image is rendered from centre t + R·(-b, 0, 0), b =
baseline/fx, which gives right(x + d, y) = left(x, y) with d = baseline/Z — the direction in which the
library searches (src/lib/depth_calculator.cpp:211-215).
"""
import numpy as np


def _rodrigues(r):
    r = np.asarray(r, np.float64)
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * K


def _blur(a, sigma):
    rad = int(3 * sigma + 0.5)
    x = np.arange(-rad, rad + 1)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    for ax in (0, 1):
        p = np.pad(a, [(rad, rad) if i == ax else (0, 0) for i in range(2)], mode="reflect")
        out = np.zeros_like(a)
        for i, w in enumerate(k):
            sl = [slice(None)] * 2
            sl[ax] = slice(i, i + a.shape[ax])
            out += w * p[tuple(sl)]
        a = out
    return a


def make_texture(h, w, seed):
    """Corner-rich, unsaturated texture: blurred noise + random grey rectangles."""
    rng = np.random.default_rng(seed)
    noise = _blur(rng.uniform(0, 255, (h, w)).astype(np.float32), 3.0)
    noise = (noise - noise.min()) / (noise.max() - noise.min())
    blocks = np.full((h, w), 0.5, np.float32)
    for _ in range(max(200, (h * w) // 6000)):
        bw, bh = rng.integers(12, 110, 2)
        x0, y0 = rng.integers(0, w - 12), rng.integers(0, h - 12)
        blocks[y0:y0 + bh, x0:x0 + bw] = rng.uniform(0.05, 0.95)
    blocks = _blur(blocks, 0.8)
    tex = 0.45 * noise + 0.55 * blocks
    tex = (tex - tex.min()) / (tex.max() - tex.min())
    return (10 + 235 * tex).astype(np.float32)


class SyntheticStereo:
    def __init__(self, width=752, height=480, fx=435.2047, fy=435.2047, cx=367.4517, cy=252.2009, baseline=47.9064,
                 seed=1234, max_step_m=0.03, max_step_deg=0.3):
        self.w, self.h = width, height
        self.fx, self.fy, self.cx, self.cy, self.baseline = fx, fy, cx, cy, baseline
        self.seed = seed
        self.tex = make_texture(2 * height, 2 * width, seed)
        rng = np.random.default_rng(seed + 1)
        # sum of 3 sinusoids per axis; amplitudes sized so per-frame steps stay under the limits
        self.freq = rng.uniform(0.02, 0.09, (6, 3))
        self.phase = rng.uniform(0, 2 * np.pi, (6, 3))
        lim = np.array([max_step_m] * 3 + [np.deg2rad(max_step_deg)] * 3)
        self.amp = (lim[:, None] / 3.0) / self.freq * rng.uniform(0.3, 0.6, (6, 3))
        self.plane_n = np.array([-0.5, -0.25, 1.0])
        self.plane_d = 4.0
        u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
        self.rays_cam = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1)

    def pose(self, k):
        """Ground-truth pose 6-vector (x, y, z, rx, ry, rz) of frame k; frame 0 is the identity."""
        p = (self.amp * (np.sin(self.freq * k + self.phase) - np.sin(self.phase))).sum(1)
        return p.astype(np.float64)

    def _render(self, t, R):
        d = self.rays_cam @ R.T
        lam = (self.plane_d - self.plane_n @ t) / (d @ self.plane_n)
        P = t + lam[..., None] * d
        u0 = self.fx * P[..., 0] / P[..., 2] + self.cx + self.w / 2.0
        v0 = self.fy * P[..., 1] / P[..., 2] + self.cy + self.h / 2.0
        u0 = np.clip(u0, 0, 2 * self.w - 1.001)
        v0 = np.clip(v0, 0, 2 * self.h - 1.001)
        x0, y0 = np.floor(u0).astype(np.int64), np.floor(v0).astype(np.int64)
        a, b = (u0 - x0).astype(np.float32), (v0 - y0).astype(np.float32)
        T = self.tex.reshape(-1)              # one flat index per pixel instead of four 2-D fancy indexings (same values, half the time)
        W2 = self.tex.shape[1]
        idx = y0 * W2 + x0
        a1, b1 = 1 - a, 1 - b
        val = a1 * b1 * T[idx] + a * b1 * T[idx + 1] + a1 * b * T[idx + W2] + a * b * T[idx + W2 + 1]
        return np.clip(np.rint(val), 0, 255).astype(np.uint8)

    def render(self, k):
        p = self.pose(k)
        t, R = p[:3], _rodrigues(p[3:])
        left = self._render(t, R)
        b = self.baseline / self.fx
        right = self._render(t + R @ np.array([-b, 0.0, 0.0]), R)
        return left, right

    def depth_at(self, k, u, v):
        """Ground-truth camera-frame depth of pixel (u, v) in frame k."""
        p = self.pose(k)
        t, R = p[:3], _rodrigues(p[3:])
        d = R @ np.array([(u - self.cx) / self.fx, (v - self.cy) / self.fy, 1.0])
        return float((self.plane_d - self.plane_n @ t) / (d @ self.plane_n))


# ---- the BASELINE.json configs (C3 = EuRoC-shaped, C4 = high-density stress) ----
CONFIGS = {
    "C3": dict(width=752, height=480, fx=435.2047, fy=435.2047, cx=367.4517, cy=252.2009, baseline=47.9064,
               grid_width=30, grid_height=24, search_x=50, search_y=6, max_pyramid_levels=4,
               min_pyramid_level_pose_estimation=2, seed=1234, frames=200),
    "C4": dict(width=1280, height=720, fx=800.0, fy=800.0, cx=640.0, cy=360.0, baseline=48.0,
               grid_width=16, grid_height=14, search_x=50, search_y=6, max_pyramid_levels=5,
               min_pyramid_level_pose_estimation=2, seed=4321, frames=50),
    # small case for fast tests
    "S": dict(width=320, height=240, fx=220.0, fy=220.0, cx=160.0, cy=120.0, baseline=22.0,
              grid_width=32, grid_height=24, search_x=30, search_y=4, max_pyramid_levels=4,
              min_pyramid_level_pose_estimation=2, seed=77, frames=30),
    # the small case with four times the camera motion: keyframes #2 and #3 appear at frames 68 and 89, so 100 frames cover
    # find_bad_keypoints / merge_keypoints with old keypoints and frames whose keypoints come from three origin keyframes
    "SF": dict(width=320, height=240, fx=220.0, fy=220.0, cx=160.0, cy=120.0, baseline=22.0,
               grid_width=32, grid_height=24, search_x=30, search_y=4, max_pyramid_levels=4,
               min_pyramid_level_pose_estimation=2, seed=77, frames=100, max_step_m=0.12, max_step_deg=1.2),
}


def settings_dict(cfg):
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    return dict(baseline=c["baseline"], fx=c["fx"], fy=c["fy"], cx=c["cx"], cy=c["cy"], k1=0.0, k2=0.0, k3=0.0, p1=0.0,
                p2=0.0, grid_height=c["grid_height"], grid_width=c["grid_width"], search_x=c["search_x"],
                search_y=c["search_y"], window_size_pose_estimator=4, window_size_opt_flow=31,
                window_size_depth_calculator=31, max_pyramid_levels=c["max_pyramid_levels"],
                min_pyramid_level_pose_estimation=c["min_pyramid_level_pose_estimation"])


def make_sequence(cfg, seed=None):
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    return SyntheticStereo(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["baseline"],
                           c["seed"] if seed is None else seed, c.get("max_step_m", 0.03), c.get("max_step_deg", 0.3))
