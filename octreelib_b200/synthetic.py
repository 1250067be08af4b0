"""
Seeded synthetic workloads for tests and bench.py.

Every coordinate is rounded to float32 and widened back to float64, so the reference (float64
numpy) and this build see identical, exactly representable values.  numpy is used for the small
parity cases; `lidar_street_torch` / `indoor_torch` generate the 1e7..1e8-point benchmark inputs
directly on the GPU (same closed forms, torch RNG).
"""
from __future__ import annotations

import numpy as np

__all__ = ["lidar64_scan", "lidar_street_scan", "indoor_scene", "lidar_street_torch", "indoor_torch"]

_N_BEAMS = 64
_N_AZIMUTH = 1875
_SENSOR_HEIGHT = 1.73
_WALL_Y = 10.0
_WALL_X = 60.0


def _ray_dirs():
    elev = np.deg2rad(np.linspace(-24.8, 2.0, _N_BEAMS))
    azim = np.linspace(0.0, 2.0 * np.pi, _N_AZIMUTH, endpoint=False)
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    d = np.stack([ce * np.cos(azim)[None, :], ce * np.sin(azim)[None, :], np.broadcast_to(se, (_N_BEAMS, _N_AZIMUTH))],
                 axis=-1)
    return d.reshape(-1, 3)


def _cast(origin, dirs, end_walls: bool, max_range: float):
    """Nearest positive hit among z=0, y=+-10 and (optionally) x=+-60."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.full(len(dirs), np.inf)
        for axis, planes in ((2, (0.0,)), (1, (-_WALL_Y, _WALL_Y)), (0, (-_WALL_X, _WALL_X) if end_walls else ())):
            for c in planes:
                ti = (c - origin[axis]) / dirs[:, axis]
                ti = np.where(ti > 1e-6, ti, np.inf)
                t = np.minimum(t, ti)
    t = np.where(t <= max_range, t, np.inf)
    return t


def lidar64_scan(pose: int, seed: int = 0, n_points=None) -> np.ndarray:
    """G-lidar64: one 64-beam x 1875-azimuth scan from (pose*1.0, 0, 1.73) inside a closed street box."""
    rng = np.random.default_rng([seed, pose])
    origin = np.array([pose * 1.0, 0.0, _SENSOR_HEIGHT])
    dirs = _ray_dirs()
    t = _cast(origin, dirs, end_walls=True, max_range=np.inf)
    t = t + rng.normal(0.0, 0.02, size=t.shape)
    pts = origin[None, :] + dirs * t[:, None]
    pts = pts[np.isfinite(pts).all(axis=1)]
    if n_points is not None:
        pts = pts[:n_points]
    return pts.astype(np.float32).astype(np.float64)


def lidar_street_scan(pose: int, seed: int = 0) -> np.ndarray:
    """C4 'infinite street' variant: no end walls, returns beyond 120 m dropped."""
    rng = np.random.default_rng([seed, pose, 4])
    origin = np.array([pose * 1.0, 0.0, _SENSOR_HEIGHT])
    dirs = _ray_dirs()
    t = _cast(origin, dirs, end_walls=False, max_range=120.0)
    t = t + rng.normal(0.0, 0.02, size=t.shape)
    pts = origin[None, :] + dirs * t[:, None]
    pts = pts[np.isfinite(pts).all(axis=1)]
    return pts.astype(np.float32).astype(np.float64)


def indoor_scene(n: int, seed: int = 0) -> np.ndarray:
    """G-indoor: 40 x 40 x 3 m, 4 x 4 rooms; floor / ceiling / x-wall / y-wall with prob 1/4 each."""
    rng = np.random.default_rng([seed, 7])
    kind = rng.integers(0, 4, size=n)
    u = rng.random((n, 3)) * np.array([40.0, 40.0, 3.0])
    noise = rng.normal(0.0, 0.005, size=n)
    wall = rng.integers(0, 5, size=n) * 10.0
    x, y, z = u[:, 0].copy(), u[:, 1].copy(), u[:, 2].copy()
    z = np.where(kind == 0, noise, z)
    z = np.where(kind == 1, 3.0 + noise, z)
    x = np.where(kind == 2, wall + noise, x)
    y = np.where(kind == 3, wall + noise, y)
    pts = np.stack([x, y, z], axis=1)
    return pts.astype(np.float32).astype(np.float64)


# ------------------------------------------------------------------------------------------------
# torch (GPU) versions for the large benchmark inputs
# ------------------------------------------------------------------------------------------------
def lidar_street_torch(pose_first: int, n_poses: int, seed: int, device):
    """`n_poses` street scans as one (n, 3) float64 CUDA tensor plus per-pose counts (python list)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed * 1000003 + pose_first)
    dirs = torch.from_numpy(_ray_dirs()).to(device)  # (R, 3) f64
    poses = torch.arange(pose_first, pose_first + n_poses, device=device, dtype=torch.float64)
    origin = torch.zeros((n_poses, 1, 3), device=device, dtype=torch.float64)
    origin[:, 0, 0] = poses
    origin[:, 0, 2] = _SENSOR_HEIGHT
    d = dirs[None, :, :]
    inf = torch.tensor(float("inf"), device=device, dtype=torch.float64)
    t = torch.full((n_poses, dirs.shape[0]), float("inf"), device=device, dtype=torch.float64)
    for axis, planes in ((2, (0.0,)), (1, (-_WALL_Y, _WALL_Y))):
        for c in planes:
            ti = (c - origin[:, :, axis]) / d[:, :, axis]
            ti = torch.where(ti > 1e-6, ti, inf)
            t = torch.minimum(t, ti)
    t = torch.where(t <= 120.0, t, inf)
    t = t + torch.randn(t.shape, generator=gen, device=device, dtype=torch.float64) * 0.02
    pts = origin + d * t[:, :, None]
    ok = torch.isfinite(pts).all(dim=2)
    counts = ok.sum(dim=1).tolist()
    pts = pts[ok].to(torch.float32).to(torch.float64)
    return pts, counts


def indoor_torch(n: int, seed: int, device):
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed * 7919 + 7)
    kind = torch.randint(0, 4, (n,), generator=gen, device=device)
    u = torch.rand((n, 3), generator=gen, device=device, dtype=torch.float64)
    u = u * torch.tensor([40.0, 40.0, 3.0], device=device, dtype=torch.float64)
    noise = torch.randn((n,), generator=gen, device=device, dtype=torch.float64) * 0.005
    wall = torch.randint(0, 5, (n,), generator=gen, device=device).to(torch.float64) * 10.0
    x, y, z = u[:, 0], u[:, 1], u[:, 2]
    z = torch.where(kind == 0, noise, z)
    z = torch.where(kind == 1, 3.0 + noise, z)
    x = torch.where(kind == 2, wall + noise, x)
    y = torch.where(kind == 3, wall + noise, y)
    return torch.stack([x, y, z], dim=1).to(torch.float32).to(torch.float64)
