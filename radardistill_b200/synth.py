"""Synthetic nuScenes-shaped point clouds for the pillar encoder (numpy only, seeded).

These stand in for the collated ``points`` / ``radar_points`` tensors the reference
hands to the VFE (``pcdet/datasets/dataset_distill.py:237-244`` prepends the batch
index column; ``pcdet/models/__init__.py:36`` moves them to the GPU):

* LiDAR rows are ``[b, x, y, z, intensity, dt]``   -> ``(N, 6)`` float32
* radar rows are ``[b, x, y, z, rcs, vx, vy]``      -> ``(N, 7)`` float32

Rows are shuffled (``data_processor.py:99-114``) and xy is clipped to the *inclusive*
range ``lo <= v <= hi`` (``common_utils.py:85-88``), so points exactly on the upper
bound reach the encoder and must be dropped there.
"""
from __future__ import annotations

import numpy as np

# radar_distill_train.yaml:6 and :63
PC_RANGE = np.array([-54.0, -54.0, -5.0, 54.0, 54.0, 3.0], dtype=np.float32)
VOXEL_SIZE = [0.075, 0.075, 0.2]
STRESS_VOXEL_SIZE = [0.05, 0.05, 0.2]


def grid_size_of(pc_range=PC_RANGE, voxel_size=VOXEL_SIZE) -> np.ndarray:
    """grid = round((hi - lo) / voxel)  (pcdet/datasets/processor/data_processor.py:116-124)."""
    pc_range = np.asarray(pc_range, dtype=np.float64)
    g = (pc_range[3:6] - pc_range[0:3]) / np.asarray(voxel_size, dtype=np.float64)
    return np.round(g).astype(np.int64)


def _inside(xy, pc_range):
    return ((xy[:, 0] >= pc_range[0]) & (xy[:, 0] <= pc_range[3]) &
            (xy[:, 1] >= pc_range[1]) & (xy[:, 1] <= pc_range[4]))


def radar_frame(seed: int = 0, n_points: int = 3000, pc_range=PC_RANGE) -> np.ndarray:
    """One radar frame, ``(n, 6)`` float32 = x, y, z, rcs, vx, vy."""
    rng = np.random.default_rng(1000 + seed)
    r = rng.gamma(2.0, 15.0, n_points)
    th = rng.uniform(0.0, 2.0 * np.pi, n_points)
    pts = np.stack([r * np.cos(th), r * np.sin(th),
                    rng.normal(0.5, 0.3, n_points),
                    rng.uniform(-5.0, 40.0, n_points),
                    rng.normal(0.0, 3.0, n_points),
                    rng.normal(0.0, 3.0, n_points)], axis=1).astype(np.float32)
    pts = pts[_inside(pts, pc_range)]
    return pts[rng.permutation(len(pts))]


def lidar_frame(seed: int = 0, sweeps: int = 10, beams: int = 32, azimuths: int = 1090,
                pc_range=PC_RANGE) -> np.ndarray:
    """One 10-sweep LiDAR frame, ``(n, 5)`` float32 = x, y, z, intensity, dt (config 2).

    32 beams x 1090 azimuths per sweep, elevations -30.67..+10.67 deg, sensor 1.84 m
    above a flat ground; the range of a ray is min(ground hit, Gamma(2, 12 m) obstacle)
    with 1 % noise; sweep k is shifted +0.5*k m in x and stamped dt = 0.05*k.
    """
    rng = np.random.default_rng(2000 + seed)
    elev = np.deg2rad(np.linspace(-30.67, 10.67, beams))
    az = np.linspace(0.0, 2.0 * np.pi, azimuths, endpoint=False)
    out = []
    for k in range(sweeps):
        e, a = np.meshgrid(elev, az + rng.uniform(0, 2 * np.pi / azimuths), indexing="ij")
        e, a = e.ravel(), a.ravel()
        with np.errstate(divide="ignore"):
            ground = np.where(e < 0, 1.84 / np.maximum(np.sin(-e), 1e-6), np.inf)
        obstacle = rng.gamma(2.0, 12.0, e.shape)
        rr = np.minimum(ground, obstacle) * (1.0 + 0.01 * rng.standard_normal(e.shape))
        x = rr * np.cos(e) * np.cos(a) + 0.5 * k
        y = rr * np.cos(e) * np.sin(a)
        z = rr * np.sin(e)  # sensor frame: ground at about -1.84
        inten = rng.uniform(0.0, 255.0, e.shape)
        dt = np.full(e.shape, 0.05 * k)
        out.append(np.stack([x, y, z, inten, dt], axis=1))
    pts = np.concatenate(out, 0).astype(np.float32)
    ego = (np.abs(pts[:, 0]) < 1.0) & (np.abs(pts[:, 1]) < 1.0)  # nuscenes_dataset_distill.py:87-89
    pts = pts[~ego & _inside(pts, pc_range) & np.isfinite(pts).all(1)]
    return pts[rng.permutation(len(pts))]


def stress_frame(seed: int = 0, n_points: int = 1_000_000, pc_range=PC_RANGE) -> np.ndarray:
    """Stress frame (config 5): 50 % uniform, 30 % in 64 Gaussian blobs (8 of them 2 cm wide), 20 % exact duplicates."""
    rng = np.random.default_rng(3000 + seed)
    n_u = n_points // 2
    n_b = (n_points * 3) // 10
    n_d = n_points - n_u - n_b
    lo, hi = pc_range[:3].astype(np.float64), pc_range[3:].astype(np.float64)
    uni = rng.uniform(lo, hi, (n_u, 3))
    centres = rng.uniform(lo[:2] * 0.9, hi[:2] * 0.9, (64, 2))
    which = rng.integers(0, 64, n_b)
    sigma = np.where(which < 8, 0.02, 0.3)[:, None]  # 8 very tight blobs: >= 1000 points in the hottest pillars
    blob_xy = centres[which] + sigma * rng.standard_normal((n_b, 2))
    blob = np.concatenate([blob_xy, rng.normal(0.0, 0.5, (n_b, 1))], 1)
    xyz = np.concatenate([uni, blob], 0)
    feats = np.stack([rng.uniform(0, 255, len(xyz)), rng.uniform(0, 0.5, len(xyz))], 1)
    base = np.concatenate([xyz, feats], 1).astype(np.float32)
    dup = base[rng.integers(0, len(base), n_d)]
    pts = np.concatenate([base, dup], 0)
    pts = pts[_inside(pts, pc_range)]
    return pts[rng.permutation(len(pts))]


def collate(frames) -> np.ndarray:
    """Prepend the float batch-index column and concatenate (dataset_distill.py:237-244)."""
    rows = []
    for b, f in enumerate(frames):
        rows.append(np.concatenate([np.full((len(f), 1), b, np.float32), f.astype(np.float32)], 1))
    if not rows:
        return np.zeros((0, 1), np.float32)
    return np.ascontiguousarray(np.concatenate(rows, 0))


def collate_frames(frames):
    """The frames back to back WITHOUT a batch column, plus int32 frame offsets (batch + 1 entries): the input of the
    device-side input prep (``rdp_index_fwd_frames``) -- no padding pass on the host, 4 bytes per point less to upload."""
    offs = np.zeros(len(frames) + 1, np.int32)
    for b, f in enumerate(frames):
        offs[b + 1] = offs[b] + len(f)
    pts = np.ascontiguousarray(np.concatenate([f.astype(np.float32) for f in frames], 0)) if frames else np.zeros((0, 1), np.float32)
    return pts, offs


def lidar_batch(batch: int, seed0: int = 0, **kw) -> np.ndarray:
    return collate([lidar_frame(seed0 + b, **kw) for b in range(batch)])


def radar_batch(batch: int, seed0: int = 0, **kw) -> np.ndarray:
    return collate([radar_frame(seed0 + b, **kw) for b in range(batch)])


def stress_batch(batch: int, seed0: int = 0, **kw) -> np.ndarray:
    return collate([stress_frame(seed0 + b, **kw) for b in range(batch)])


def edge_case_points(kind: str, pc_range=PC_RANGE, voxel=VOXEL_SIZE) -> np.ndarray:
    """Small adversarial LiDAR-shaped ``(N, 6)`` batches for the cases of SURVEY A.5."""
    rng = np.random.default_rng(77)
    f32 = np.float32
    base = lidar_frame(5, sweeps=1, beams=8, azimuths=64)
    if kind == "boundary":
        extra = np.array([
            [54.0, 0.0, 0.0, 1, 0], [0.0, 54.0, 0.0, 1, 0], [-54.0, -54.0, 0.0, 1, 0],
            [53.999996, 1.0, 0.0, 1, 0], [1.0, 53.999996, 0.0, 1, 0],
            [np.nextafter(f32(54.0), f32(0.0)), 2.0, 0.0, 1, 0],
            [-54.000004, 0.0, 0.0, 1, 0], [0.0, -54.000004, 0.0, 1, 0],
            [53.92, 53.92, 0.0, 1, 0], [-53.9999, -53.9999, 0.0, 1, 0],
            [0.075, 0.075, 0.0, 1, 0], [0.07499999, 0.15, 0.0, 1, 0], [0.0, 0.0, 100.0, 1, 0],
        ], dtype=f32)
        frames = [np.concatenate([base[:200], extra]), np.concatenate([extra, base[200:300]])]
    elif kind == "nonfinite":
        extra = np.array([
            [np.nan, 0.0, 0.0, 1, 0], [0.0, np.nan, 0.0, 1, 0], [np.inf, 0.0, 0.0, 1, 0],
            [0.0, -np.inf, 0.0, 1, 0], [1e30, 1.0, 0.0, 1, 0], [-1e30, 1.0, 0.0, 1, 0],
        ], dtype=f32)
        mix = np.concatenate([base[:150], extra, base[150:260]])
        frames = [mix[rng.permutation(len(mix))], base[260:400]]
    elif kind == "empty_frame":
        frames = [base[:120], base[:0], base[120:260], base[:0]]
    elif kind == "duplicates":
        few = base[:40]
        frames = [np.concatenate([few, few, few[::-1], few[:7]]), np.concatenate([few[:5]] * 9)]
    elif kind == "dense_cell":
        cell = np.array([[10.01, -3.02, 0.0, 5, 0.1]], f32) + \
            (rng.uniform(0, 0.05, (700, 5)) * np.array([1, 1, 3, 50, 1])).astype(f32)
        frames = [np.concatenate([base[:100], cell]), cell[:333]]
    elif kind == "all_outside":
        frames = [np.array([[54.0, 0, 0, 1, 0], [0, 54.0, 0, 1, 0], [np.nan, 0, 0, 1, 0]], f32)]
    elif kind == "empty":
        frames = [base[:0]]
    elif kind == "single":
        frames = [base[:1]]
    else:
        raise KeyError(kind)
    return collate(frames) if kind != "empty" else np.zeros((0, 6), np.float32)


EDGE_KINDS = ["boundary", "nonfinite", "empty_frame", "duplicates", "dense_cell",
              "all_outside", "empty", "single"]
