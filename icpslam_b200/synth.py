"""Seeded synthetic clouds for the ICP hot path.

The reference ships no bags, .pcd files or fixtures, so every workload named in
BASELINE.json is generated here with ``numpy.random.default_rng(seed)``.  All clouds are returned
as C-contiguous ``float32[N, 4]`` arrays {x, y, z, 1} — the layout of ``pcl::PointXYZ`` that the
reference keeps in ``prev_cloud_`` / ``curr_cloud_`` (reference include/icpslam/icp_odometer.h:107-108).

Nothing in this module touches the GPU; it is shared by tests/, bench.py and the smoke check.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

__all__ = [
    "World",
    "make_world",
    "trajectory",
    "hdl64_sweep",
    "sweep_sequence",
    "local_map",
    "planar_room",
    "planar_scan",
    "integer_cloud",
    "random_rigid",
    "rot_xyz",
    "as_xyzw",
]


def as_xyzw(xyz: np.ndarray) -> np.ndarray:
    """float32[N,4] with w = 1 from an [N,3] array."""
    xyz = np.asarray(xyz, dtype=np.float32)
    out = np.ones((xyz.shape[0], 4), dtype=np.float32)
    out[:, :3] = xyz[:, :3]
    return np.ascontiguousarray(out)


def rot_xyz(roll: float, pitch: float, yaw: float) -> np.ndarray:
    """R = Rz(yaw) Ry(pitch) Rx(roll) (the Euler convention of PCL's GICP)."""
    cr, sr = math.cos(roll), math.sin(roll)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cy, sy = math.cos(yaw), math.sin(yaw)
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return rz @ ry @ rx


def random_rigid(rng: np.random.Generator, max_trans: float, max_rot: float) -> np.ndarray:
    """4x4 float64 rigid transform with |t|_inf <= max_trans and Euler angles <= max_rot (rad)."""
    T = np.eye(4)
    ang = rng.uniform(-max_rot, max_rot, 3)
    T[:3, :3] = rot_xyz(*ang)
    T[:3, 3] = rng.uniform(-max_trans, max_trans, 3)
    return T


# ------------------------------------------------------------------------------------------------
# 3-D world + HDL-64 ray casting
# ------------------------------------------------------------------------------------------------
@dataclass
class World:
    half_x: float
    half_y: float
    height: float
    boxes_min: np.ndarray  # [B,3]
    boxes_max: np.ndarray  # [B,3]
    cyl_xy: np.ndarray  # [C,2]
    cyl_r: np.ndarray  # [C]
    cyl_h: np.ndarray  # [C]


def make_world(seed: int, scale: float = 1.0, n_boxes: int = 40, n_cyl: int = 20) -> World:
    """Closed scene so that every ray returns: ground z=0, enclosing box 120x60x12 m (x scale),
    ``n_boxes`` axis-aligned boxes (1-6 m) and ``n_cyl`` vertical cylinders (r 0.2-0.5 m)."""
    rng = np.random.default_rng(seed)
    hx, hy, hz = 60.0 * scale, 30.0 * scale, 12.0
    nb = int(n_boxes * scale * scale)
    nc = int(n_cyl * scale * scale)
    size = rng.uniform(1.0, 6.0, (nb, 3))
    size[:, 2] = rng.uniform(1.0, 6.0, nb)
    cx = rng.uniform(-hx + 4, hx - 4, nb)
    cy = rng.uniform(-hy + 4, hy - 4, nb)
    # keep a corridor |y| < 2.5 m free for the vehicle
    cy = np.where(np.abs(cy) < 2.5 + size[:, 1] / 2, np.sign(cy + 1e-9) * (2.5 + size[:, 1] / 2 + 0.5), cy)
    bmin = np.stack([cx - size[:, 0] / 2, cy - size[:, 1] / 2, np.zeros(nb)], 1)
    bmax = np.stack([cx + size[:, 0] / 2, cy + size[:, 1] / 2, size[:, 2]], 1)
    cyl_xy = np.stack([rng.uniform(-hx + 2, hx - 2, nc), rng.uniform(-hy + 2, hy - 2, nc)], 1)
    cyl_xy[:, 1] = np.where(np.abs(cyl_xy[:, 1]) < 3.0, np.sign(cyl_xy[:, 1] + 1e-9) * 3.5, cyl_xy[:, 1])
    return World(hx, hy, hz, bmin, bmax, cyl_xy, rng.uniform(0.2, 0.5, nc), rng.uniform(2.0, 8.0, nc))


def trajectory(seed: int, n: int, start_x: float = -40.0) -> list[np.ndarray]:
    """n sensor poses (4x4, world <- sensor).  Per step: forward U(0.2,0.6) m, lateral N(0,0.02),
    yaw N(0,1.5 deg), roll/pitch N(0,0.3 deg), z N(0,0.01); sensor 1.73 m above ground."""
    rng = np.random.default_rng(seed)
    x, y, yaw = start_x, 0.0, 0.0
    poses = []
    for _ in range(n):
        roll, pitch = rng.normal(0, math.radians(0.3), 2)
        z = 1.73 + rng.normal(0, 0.01)
        T = np.eye(4)
        T[:3, :3] = rot_xyz(roll, pitch, yaw)
        T[:3, 3] = (x, y, z)
        poses.append(T)
        fwd = rng.uniform(0.2, 0.6)
        lat = rng.normal(0, 0.02)
        yaw_step = rng.normal(0, math.radians(1.5))
        # steer gently back towards the corridor axis so long runs stay inside the scene
        yaw_step -= 0.02 * yaw + 0.002 * y
        x += fwd * math.cos(yaw) - lat * math.sin(yaw)
        y += fwd * math.sin(yaw) + lat * math.cos(yaw)
        yaw += yaw_step
    return poses


def _raycast(world: World, origin: np.ndarray, dirs: np.ndarray) -> np.ndarray:
    """Range along each unit ray (world frame) to the first surface."""
    n = dirs.shape[0]
    o = origin.astype(np.float64)
    d = dirs.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        # enclosing box, seen from the inside: first exit
        lo = np.array([-world.half_x, -world.half_y, 0.0])
        hi = np.array([world.half_x, world.half_y, world.height])
        t1 = (lo - o) * inv
        t2 = (hi - o) * inv
        t_exit = np.nanmin(np.maximum(t1, t2), axis=1)
        best = t_exit
        # axis-aligned boxes: slab test
        if len(world.boxes_min):
            ta = (world.boxes_min[None] - o) * inv[:, None, :]  # [n,B,3]
            tb = (world.boxes_max[None] - o) * inv[:, None, :]
            tn = np.nanmax(np.minimum(ta, tb), axis=2)
            tf = np.nanmin(np.maximum(ta, tb), axis=2)
            hit = (tn < tf) & (tn > 1e-3)
            tbox = np.where(hit, tn, np.inf).min(axis=1)
            best = np.minimum(best, tbox)
        # vertical cylinders (side surface only)
        if len(world.cyl_r):
            ox = o[0] - world.cyl_xy[:, 0][None]  # [1,C]
            oy = o[1] - world.cyl_xy[:, 1][None]
            dx, dy = d[:, 0:1], d[:, 1:2]
            a = dx * dx + dy * dy
            b = 2 * (ox * dx + oy * dy)
            c = ox * ox + oy * oy - world.cyl_r[None] ** 2
            disc = b * b - 4 * a * c
            sq = np.sqrt(np.where(disc > 0, disc, np.nan))
            tc = (-b - sq) / (2 * a)
            zc = o[2] + tc * d[:, 2:3]
            ok = (disc > 0) & (tc > 1e-3) & (zc >= 0) & (zc <= world.cyl_h[None])
            tcy = np.where(ok, tc, np.inf).min(axis=1)
            best = np.minimum(best, tcy)
    assert best.shape == (n,)
    return best


def _hdl64_dirs(n_beams: int, n_az: int) -> np.ndarray:
    elev = np.radians(np.linspace(2.0, -24.8, n_beams))
    az = np.arange(n_az) * (2 * math.pi / n_az)
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    d = np.stack([ce * np.cos(az)[None], ce * np.sin(az)[None], np.broadcast_to(se, (n_beams, n_az))], -1)
    return d.reshape(-1, 3)


def hdl64_sweep(world: World, pose: np.ndarray, rng: np.random.Generator, n_beams: int = 64,
                n_az: int = 1024, noise: float = 0.02) -> np.ndarray:
    """One Velodyne HDL-64-like sweep in the SENSOR frame: n_beams x n_az rays (65 536 by default),
    beam-major order, elevation +2.0 .. -24.8 deg, range noise N(0, noise)."""
    dirs_s = _hdl64_dirs(n_beams, n_az)
    R, t = pose[:3, :3], pose[:3, 3]
    dirs_w = dirs_s @ R.T
    rng_m = _raycast(world, t, dirs_w)
    rng_m = rng_m + rng.normal(0.0, noise, rng_m.shape)
    return as_xyzw(dirs_s * rng_m[:, None])


def sweep_sequence(config: int, n: int, n_beams: int = 64, n_az: int = 1024, world_scale: float = 1.0):
    """(world, poses, [n sweeps]) with seed = 1000*config + scan index."""
    world = make_world(1000 * config, world_scale)
    poses = trajectory(1000 * config + 999, n)
    sweeps = [hdl64_sweep(world, poses[i], np.random.default_rng(1000 * config + i), n_beams, n_az)
              for i in range(n)]
    return world, poses, sweeps


def voxel_dedup_first(xyz: np.ndarray, voxel: float) -> np.ndarray:
    """At most one point per ``voxel`` cube, first-come wins, insertion order kept — the net effect
    of OctreeMapper::addPointsToMap (reference src/icpslam/octree_mapper.cpp:63-71)."""
    key = np.floor(xyz[:, :3].astype(np.float64) / voxel).astype(np.int64)
    key -= key.min(axis=0)
    dims = key.max(axis=0) + 1
    lin = (key[:, 2] * dims[1] + key[:, 1]) * dims[0] + key[:, 0]
    _, first = np.unique(lin, return_index=True)
    first.sort()
    return xyz[first]


def local_map(config: int, n_points: int = 500_000, voxel: float = 0.2, n_sweeps: int = 25,
              world_scale: float = 1.0, max_sweeps: int = 400):
    """Accumulated local map + the next sweep (BASELINE.json configs[1]).

    Returns (map_xyzw[n_points,4] in the frame of the LAST map pose, query sweep in its own sensor
    frame, T_true 4x4 mapping query-frame points into the map frame).  The map is the union of
    sweeps along the trajectory, one point per ``voxel`` cube (first-come), truncated to exactly
    n_points; more sweeps (denser elevation pattern) are added until that many voxels are occupied."""
    m = build_local_map(config, n_points, voxel, n_sweeps, world_scale)
    q_pose = m["poses"][m["k"]]
    query = hdl64_sweep(m["world"], q_pose, np.random.default_rng(1000 * config + 500 + m["k"]))
    return m["map"], query, m["ref_inv"] @ q_pose


def build_local_map(config: int, n_points: int = 500_000, voxel: float = 0.2, n_sweeps: int = 25,
                    world_scale: float = 1.0, max_batches: int = 200) -> dict:
    """The map half of ``local_map`` plus what is needed to cast further sweeps against it:
    dict(map, world, poses, k = index of the first pose after the map, ref_inv = map <- world).

    Union of ``n_sweeps`` HDL-64 sweeps along the trajectory, one point per ``voxel`` cube
    (first-come wins, insertion order kept — OctreeMapper::addPointsToMap's net effect), then PADDED to
    exactly ``n_points`` with returns of extra rays cast in random directions from the same poses
    (the accumulated map of a longer drive), deduplicated the same way, and truncated."""
    world = make_world(1000 * config, world_scale)
    poses = trajectory(1000 * config + 999, n_sweeps + 1200)
    lo = np.array([-world.half_x - 1.0, -world.half_y - 1.0, -1.0])
    dims = np.ceil((np.array([world.half_x, world.half_y, world.height]) + 1.0 - lo) / voxel).astype(np.int64) + 1
    seen = np.empty(0, dtype=np.int64)  # sorted voxel keys already in the map
    chunks = []
    count = 0

    def insert(pw):
        nonlocal seen, count
        key3 = np.floor((pw - lo) / voxel).astype(np.int64)
        key = (key3[:, 2] * dims[1] + key3[:, 1]) * dims[0] + key3[:, 0]
        _, first = np.unique(key, return_index=True)
        first.sort()
        fresh = first[~np.isin(key[first], seen, assume_unique=True)]
        chunks.append(pw[fresh])
        seen = np.union1d(seen, key[fresh])
        count += len(fresh)

    for k in range(n_sweeps):
        s = hdl64_sweep(world, poses[k], np.random.default_rng(1000 * config + k))
        insert(s[:, :3].astype(np.float64) @ poses[k][:3, :3].T + poses[k][:3, 3])
    rng = np.random.default_rng(1000 * config + 777)
    batches = 0
    while count < n_points:
        if batches >= max_batches:
            raise RuntimeError(f"build_local_map: only {count} occupied voxels after {batches} padding batches")
        insert(_sample_surfaces(world, rng, 400_000))
        batches += 1
    total = np.concatenate(chunks)[:n_points]
    k = n_sweeps
    ref = poses[k - 1]  # map frame = frame of the last pose that contributed
    ref_inv = np.linalg.inv(ref)
    map_local = total @ ref_inv[:3, :3].T + ref_inv[:3, 3]
    return dict(map=as_xyzw(map_local), world=world, poses=poses, k=k, ref_inv=ref_inv)


def _sample_surfaces(world: World, rng: np.random.Generator, n: int, noise: float = 0.01) -> np.ndarray:
    """n points drawn area-uniformly from the scene's surfaces (ground, ceiling, walls, box tops and
    sides, cylinder sides) with isotropic noise: what a long drive would have accumulated."""
    hx, hy, hz = world.half_x, world.half_y, world.height
    bs = world.boxes_max - world.boxes_min  # [B,3]
    areas = [4 * hx * hy, 4 * hx * hy, 4 * hy * hz, 4 * hx * hz]
    box_top = bs[:, 0] * bs[:, 1]
    box_sx = 2 * bs[:, 1] * bs[:, 2]  # the two faces normal to x
    box_sy = 2 * bs[:, 0] * bs[:, 2]
    cyl = 2 * np.pi * world.cyl_r * world.cyl_h
    w = np.concatenate([areas, box_top, box_sx, box_sy, cyl])
    kind = rng.choice(len(w), size=n, p=w / w.sum())
    u, v, side = rng.uniform(size=n), rng.uniform(size=n), rng.integers(0, 2, n)
    p = np.zeros((n, 3))
    nb, nc = len(bs), len(cyl)
    m = kind == 0  # ground
    p[m] = np.stack([(2 * u[m] - 1) * hx, (2 * v[m] - 1) * hy, np.zeros(m.sum())], 1)
    m = kind == 1  # ceiling
    p[m] = np.stack([(2 * u[m] - 1) * hx, (2 * v[m] - 1) * hy, np.full(m.sum(), hz)], 1)
    m = kind == 2  # walls x = +-hx
    p[m] = np.stack([(2 * side[m] - 1) * hx, (2 * u[m] - 1) * hy, v[m] * hz], 1)
    m = kind == 3  # walls y = +-hy
    p[m] = np.stack([(2 * u[m] - 1) * hx, (2 * side[m] - 1) * hy, v[m] * hz], 1)
    k0 = 4
    m = (kind >= k0) & (kind < k0 + nb)
    b = kind[m] - k0
    p[m] = np.stack([world.boxes_min[b, 0] + u[m] * bs[b, 0], world.boxes_min[b, 1] + v[m] * bs[b, 1],
                     world.boxes_max[b, 2]], 1)
    k0 += nb
    m = (kind >= k0) & (kind < k0 + nb)
    b = kind[m] - k0
    p[m] = np.stack([np.where(side[m] == 0, world.boxes_min[b, 0], world.boxes_max[b, 0]),
                     world.boxes_min[b, 1] + u[m] * bs[b, 1], v[m] * bs[b, 2]], 1)
    k0 += nb
    m = (kind >= k0) & (kind < k0 + nb)
    b = kind[m] - k0
    p[m] = np.stack([world.boxes_min[b, 0] + u[m] * bs[b, 0],
                     np.where(side[m] == 0, world.boxes_min[b, 1], world.boxes_max[b, 1]), v[m] * bs[b, 2]], 1)
    k0 += nb
    m = (kind >= k0) & (kind < k0 + nc)
    c = kind[m] - k0
    ang = 2 * np.pi * u[m]
    p[m] = np.stack([world.cyl_xy[c, 0] + world.cyl_r[c] * np.cos(ang), world.cyl_xy[c, 1] + world.cyl_r[c] * np.sin(ang),
                     v[m] * world.cyl_h[c]], 1)
    return p + rng.normal(0.0, noise, p.shape)


def map_queries(m: dict, config: int, first: int, count: int, trans_sigma: float = 0.08,
                rot_sigma_deg: float = 0.4, variant: int = 0):
    """``count`` sweeps taken after the map of ``build_local_map``, each already moved into the map
    frame by its true pose composed with a small odometry error — what
    OctreeMapper::refineTransformAndGrowMap hands to ICP after `cloud_in_map = raw_pose (x) cloud`
    (reference src/icpslam/octree_mapper.cpp:136).  Returns (list of float32[N,4], list of the 4x4
    corrections T_fix with  T_fix * query ~ map).  ``variant`` re-draws the noise of the same poses."""
    out, fixes = [], []
    for j in range(first, first + count):
        idx = m["k"] + j
        pose = m["poses"][idx]
        # variant > 0: the same pose seen again with other range noise and another odometry error
        rng = np.random.default_rng(1000 * config + 500 + idx + 1_000_000 * variant)
        sweep = hdl64_sweep(m["world"], pose, rng)
        err = np.eye(4)
        err[:3, :3] = rot_xyz(*np.radians(rng.normal(0, rot_sigma_deg, 3)))
        err[:3, 3] = np.clip(rng.normal(0, trans_sigma, 3), -0.25, 0.25)
        T = err @ (m["ref_inv"] @ pose)
        q = sweep[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
        out.append(as_xyzw(q))
        fixes.append(np.linalg.inv(err))
    return out, fixes


def _pitch(a: float) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = rot_xyz(0.0, a, 0.0)
    return T


# ------------------------------------------------------------------------------------------------
# 2-D planar lidar (BASELINE.json configs[2])
# ------------------------------------------------------------------------------------------------
@dataclass
class Room:
    segs: np.ndarray  # [S,4] x0,y0,x1,y1
    pillars: np.ndarray  # [P,3] x,y,r


def planar_room(seed: int) -> Room:
    rng = np.random.default_rng(seed)
    hx, hy = 10.0, 6.0
    corners = [(-hx, -hy), (hx, -hy), (hx, hy * 0.4), (hx * 0.6, hy), (-hx, hy)]
    segs = np.array([corners[i] + corners[(i + 1) % len(corners)] for i in range(len(corners))], dtype=np.float64)
    pillars = np.stack([rng.uniform(-8, 8, 6), rng.uniform(-4.5, 4.5, 6), rng.uniform(0.15, 0.4, 6)], 1)
    pillars[:, 1] = np.where(np.abs(pillars[:, 1]) < 1.0, np.sign(pillars[:, 1] + 1e-9) * 1.5, pillars[:, 1])
    return Room(segs, pillars)


def planar_scan(room: Room, x: float, y: float, yaw: float, rng: np.random.Generator, n_beams: int = 1080,
                fov_deg: float = 270.0, noise: float = 0.01) -> np.ndarray:
    """1080-beam 270 deg planar scan (0.25 deg step), z = 0, in the sensor frame."""
    ang = np.radians(np.arange(n_beams) * (fov_deg / n_beams) - fov_deg / 2)
    d = np.stack([np.cos(ang + yaw), np.sin(ang + yaw)], 1)
    o = np.array([x, y])
    best = np.full(n_beams, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        for x0, y0, x1, y1 in room.segs:
            e = np.array([x1 - x0, y1 - y0])
            den = d[:, 0] * e[1] - d[:, 1] * e[0]
            w = np.array([x0, y0]) - o
            t = (w[0] * e[1] - w[1] * e[0]) / den
            u = (w[0] * d[:, 1] - w[1] * d[:, 0]) / den
            ok = (t > 1e-3) & (u >= 0) & (u <= 1)
            best = np.where(ok & (t < best), t, best)
        for px, py, pr in room.pillars:
            oc = o - np.array([px, py])
            b = 2 * (d @ oc)
            c = oc @ oc - pr * pr
            disc = b * b - 4 * c
            t = (-b - np.sqrt(np.where(disc > 0, disc, np.nan))) / 2
            ok = (disc > 0) & (t > 1e-3)
            best = np.where(ok & (t < best), t, best)
    best = np.where(np.isfinite(best), best, 30.0) + rng.normal(0, noise, n_beams)
    loc = np.stack([np.cos(ang), np.sin(ang), np.zeros(n_beams)], 1) * best[:, None]
    return as_xyzw(loc)


def planar_stream(config: int, n: int):
    """n consecutive planar scans, motion 2.5 cm / 0.5 deg per scan (40 Hz stream)."""
    room = planar_room(1000 * config)
    rng = np.random.default_rng(1000 * config + 999)
    x, y, yaw = -5.0, 0.0, 0.0
    scans, poses = [], []
    for i in range(n):
        scans.append(planar_scan(room, x, y, yaw, np.random.default_rng(1000 * config + i)))
        poses.append((x, y, yaw))
        x += 0.025 * math.cos(yaw)
        y += 0.025 * math.sin(yaw)
        yaw += math.radians(0.5) * (1 if (i // 40) % 2 == 0 else -1) + rng.normal(0, 1e-4)
    return room, poses, scans


# ------------------------------------------------------------------------------------------------
# Integer fixtures for bit-exact NN parity: coordinates in Z ∩ [-1024, 1024] make
# every float32 subtraction, square and 3-term sum exact, whatever the operation order or fusion.
# ------------------------------------------------------------------------------------------------
def integer_cloud(seed: int, n: int, lo: int = -1024, hi: int = 1024, unique: bool = True) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if not unique:
        return as_xyzw(rng.integers(lo, hi + 1, (n, 3)))
    if n > (hi - lo + 1) ** 3:
        raise ValueError("integer_cloud: more unique points requested than lattice sites")
    got = np.empty((0, 3), dtype=np.int64)
    while len(got) < n:
        cand = rng.integers(lo, hi + 1, (int((n - len(got)) * 1.2) + 16, 3))
        allp = np.concatenate([got, cand])
        _, first = np.unique(allp, axis=0, return_index=True)
        first.sort()
        got = allp[first]
    return as_xyzw(got[:n])
