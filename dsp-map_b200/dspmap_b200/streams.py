"""Deterministic synthetic depth-cloud + pose streams.

One stream = F frames of (cloud in the SENSOR frame, sensor position, sensor attitude quaternion, time stamp),
exactly the arguments DSPMap::update() takes (include/dsp_dynamic.h:181-184). The scene is a ground plane, a back
wall, side walls and K axis-aligned boxes, some of which translate in x/y so that the map sees dynamic clusters,
voxel-boundary crossings and non-trivial future status. Everything is float32 and a pure function of
(config, seed, frames, points), so the reference arm, the oracle and the GPU arm consume identical bytes.
"""
import numpy as np


def _quat(roll, pitch, yaw):
    cr, sr = np.cos(roll / 2), np.sin(roll / 2)
    cp, sp = np.cos(pitch / 2), np.sin(pitch / 2)
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    q = np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                  cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy])
    return q / np.linalg.norm(q)


def _rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_stream(cfg, seed=1, frames=100, points=None, dt=0.1, speed=0.2, n_boxes=6, dynamic=True):
    """cfg: a dict from configs.CONFIGS. Returns dict(points[F,M,3] f32, n[F] i32, pos[F,3] f32, quat[F,4] f32 (w,x,y,z),
    t[F] f64)."""
    M = int(points if points is not None else cfg["points"])
    rng = np.random.default_rng(seed)
    half = 0.5 * cfg["res"] * np.array([cfg["nx"], cfg["ny"], cfg["nz"]], dtype=np.float64)
    hfov = np.deg2rad(cfg["half_fov_h"]) * 0.98
    vfov = np.deg2rad(cfg["half_fov_v"]) * 0.98
    sensor_h = min(1.0, 0.7 * half[2])  # sensor height above the ground plane
    # boxes: centre xy (world), half sizes, height, velocity
    bx = rng.uniform(0.3 * half[0], 1.3 * half[0], n_boxes)
    by = rng.uniform(-0.6 * half[1], 0.6 * half[1], n_boxes)
    bs = rng.uniform(0.15, 0.35, (n_boxes, 2))
    bh = rng.uniform(0.6, 1.4, n_boxes)
    bv = np.zeros((n_boxes, 2))
    if dynamic:
        moving = np.arange(n_boxes) % 2 == 0
        # moving obstacles are thin (pedestrian-like) so their clusters stay under the reference's
        # DYNAMIC_CLUSTER_MAX_POINT_NUM = 200 points (dsp_dynamic.h:52) at these cloud densities
        shrink = np.sqrt(10000.0 / max(M, 2500)) * (half[0] / 4.95)
        bs[moving] = rng.uniform(0.07, 0.12, (int(moving.sum()), 2)) * shrink
        bh[moving] = rng.uniform(0.5, 0.9, int(moving.sum()))
        ang = rng.uniform(0, 2 * np.pi, n_boxes)
        spd = rng.uniform(0.5, 1.5, n_boxes)
        bv[moving, 0] = (spd * np.cos(ang))[moving]
        bv[moving, 1] = (spd * np.sin(ang))[moving]
    wall_x = 1.6 * half[0]
    out_p = np.zeros((frames, M, 3), np.float32)
    pos = np.zeros((frames, 3), np.float32)
    quat = np.zeros((frames, 4), np.float32)
    ts = np.zeros(frames, np.float64)
    for f in range(frames):
        t = f * dt
        s = np.array([speed * t, 0.05 * np.sin(0.5 * t), sensor_h + 0.02 * np.sin(0.3 * t)])
        q = _quat(0.02 * np.sin(0.7 * t), 0.03 * np.sin(0.4 * t), 0.15 * np.sin(0.2 * t))
        R = _rot(q)
        pts = np.zeros((0, 3))
        rounds, skip = 0, np.zeros(n_boxes, bool)
        while pts.shape[0] < M:
            rounds += 1
            if rounds == 20:   # the sensor has flown into an obstacle (every return is nearer than 0.2 m): see through it
                cxy = np.stack([bx + bv[:, 0] * t, by + bv[:, 1] * t], 1)
                skip = np.all(np.abs(cxy - s[:2]) < bs + 0.5, axis=1)
                pts = np.zeros((0, 3))
            if rounds > 200:
                raise RuntimeError("make_stream: frame %d has no valid returns (%d of %d points)" % (f, pts.shape[0], M))
            n = 2 * M
            az = rng.uniform(-hfov, hfov, n)
            ev = rng.uniform(-vfov, vfov, n)
            d = np.stack([np.ones(n), np.tan(az), np.tan(ev)], 1)
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            dw = d @ R.T  # world direction
            rng_hit = np.full(n, np.inf)
            # ground z = 0
            with np.errstate(divide="ignore", invalid="ignore"):
                tg = (0.0 - s[2]) / dw[:, 2]
                tg[(dw[:, 2] >= 0) | (tg <= 0)] = np.inf
                rng_hit = np.minimum(rng_hit, tg)
                tw = (wall_x - s[0]) / dw[:, 0]
                tw[(dw[:, 0] <= 0) | (tw <= 0)] = np.inf
                rng_hit = np.minimum(rng_hit, tw)
                for side in (-1.0, 1.0):  # corridor side walls, fixed in the world
                    ty = (side * 0.85 * half[1] - s[1]) / dw[:, 1]
                    ty[(ty <= 0)] = np.inf
                    rng_hit = np.minimum(rng_hit, ty)
                for k in range(n_boxes):
                    if skip[k]:
                        continue
                    c = np.array([bx[k] + bv[k, 0] * t, by[k] + bv[k, 1] * t])
                    lo = np.array([c[0] - bs[k, 0], c[1] - bs[k, 1], 0.0])
                    hi = np.array([c[0] + bs[k, 0], c[1] + bs[k, 1], bh[k]])
                    t1 = (lo - s) / dw
                    t2 = (hi - s) / dw
                    tn = np.max(np.minimum(t1, t2), axis=1)
                    tf = np.min(np.maximum(t1, t2), axis=1)
                    hit = (tn <= tf) & (tn > 0.05)
                    rng_hit = np.where(hit, np.minimum(rng_hit, tn), rng_hit)
            ok = np.isfinite(rng_hit)
            rel = dw * rng_hit[:, None]  # sensor-centred, world-aligned
            ok &= np.all(np.abs(rel) < 0.985 * half, axis=1) & (rng_hit > 0.2)
            pts = np.concatenate([pts, (d * rng_hit[:, None])[ok]], 0)
        out_p[f] = pts[:M].astype(np.float32)
        pos[f] = s.astype(np.float32)
        quat[f] = q.astype(np.float32)
        ts[f] = t
    return dict(points=out_p, n=np.full(frames, M, np.int32), pos=pos, quat=quat, t=ts)


def make_depth_cloud(width=640, height=480, seed=1, hfov_deg=90.0, max_range=9.0, invalid=0.08, stride=3):
    """A raw depth-camera cloud as the application receives it (map_sim_example.cpp:305-309): one point per pixel in the
    CAMERA frame (x right, y down, z forward), raster order, invalid pixels NaN.  Scene: floor, back wall, a few boxes.
    Returns (height*width, stride) float32; columns past 2 are padding, like the 16-byte points of a PointCloud2."""
    rng = np.random.default_rng(seed)
    f = 0.5 * width / np.tan(np.radians(hfov_deg) / 2)
    u, v = np.meshgrid(np.arange(width, dtype=np.float64) - width / 2 + 0.5, np.arange(height, dtype=np.float64) - height / 2 + 0.5)
    dx, dy = u / f, v / f                                  # ray direction (dx, dy, 1)
    depth = np.full(u.shape, max_range * 1.2)
    with np.errstate(divide="ignore", invalid="ignore"):
        floor = np.where(dy > 1e-6, 1.0 / dy, np.inf)       # floor 1 m below the camera (y down)
    depth = np.minimum(depth, floor)
    depth = np.minimum(depth, 6.0 + 0.3 * np.sin(3 * dx))  # gently curved back wall
    for _ in range(5):                                     # boxes: axis-aligned slabs in front of the wall
        cx, cy, cz = rng.uniform(-2.5, 2.5), rng.uniform(-0.8, 0.9), rng.uniform(1.0, 5.0)
        hx, hy = rng.uniform(0.15, 0.6), rng.uniform(0.2, 0.9)
        hit = (np.abs(dx * cz - cx) < hx) & (np.abs(dy * cz - cy) < hy)
        depth = np.where(hit, np.minimum(depth, cz), depth)
    depth = depth + rng.normal(0, 0.004, depth.shape) * depth
    bad = (depth > max_range) | (rng.random(depth.shape) < invalid)
    pts = np.zeros((height * width, stride), np.float32)
    pts[:, 0] = (dx * depth).ravel()
    pts[:, 1] = (dy * depth).ravel()
    pts[:, 2] = depth.ravel()
    pts[bad.ravel(), :3] = np.nan
    return pts


def write_stream(path, stream, frames=None):
    """Binary stream file read by dsp-map_b200/tools/dspmap_replay.cpp and tests/dropin_main.cpp."""
    import struct
    F = len(stream["t"]) if frames is None else frames
    with open(path, "wb") as f:
        f.write(struct.pack("i", F))
        for k in range(F):
            n = int(stream["n"][k])
            f.write(struct.pack("i", n))
            f.write(stream["pos"][k].astype(np.float32).tobytes() + stream["quat"][k].astype(np.float32).tobytes())
            f.write(struct.pack("d", float(stream["t"][k])))
            f.write(stream["points"][k][:n].astype(np.float32).tobytes())
