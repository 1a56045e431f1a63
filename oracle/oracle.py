"""ctypes wrapper + build recipe for oracle/dsp_oracle.cpp (the CPU restatement).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle.so")
SRC = os.path.join(HERE, "dsp_oracle.cpp")


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", SO])
    return SO


class Config(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("resolution", C.c_float),
                ("angle_resolution", C.c_int32), ("half_fov_h", C.c_int32), ("half_fov_v", C.c_int32),
                ("max_ppv", C.c_int32), ("safe_ppv", C.c_int32), ("safe_pyramid", C.c_int32),
                ("neighbor_n", C.c_int32), ("model", C.c_int32), ("prediction_times", C.c_int32),
                ("future_time", C.c_float * 8), ("occlusion_margin", C.c_float), ("init_particle_num", C.c_int32),
                ("init_weight", C.c_float), ("table_seed", C.c_uint64), ("uniform_seed", C.c_uint64),
                ("gaussian_table_size", C.c_int32), ("obs_max_per_pyramid", C.c_int32), ("pi_is_double", C.c_int32)]


def make_config(cfg, seed=1, init_particles=0, init_weight=0.01, safe_ppv=0, safe_pyramid=0, table_size=10000000):
    c = Config()
    c.nx, c.ny, c.nz = cfg["nx"], cfg["ny"], cfg["nz"]
    c.resolution = cfg["res"]
    c.angle_resolution = cfg["angle_res"]
    c.half_fov_h, c.half_fov_v = cfg["half_fov_h"], cfg["half_fov_v"]
    c.max_ppv = cfg["max_ppv"]
    c.safe_ppv, c.safe_pyramid = safe_ppv, safe_pyramid
    c.neighbor_n = cfg["neighbor_n"]
    c.model = 1 if cfg["model"] == "static" else 0
    c.prediction_times = len(cfg["future_times"])
    for i, t in enumerate(cfg["future_times"]):
        c.future_time[i] = t
    c.occlusion_margin = 0.3 if cfg["header"] == "dsp_dynamic.h" else cfg["res"]
    c.init_particle_num, c.init_weight = init_particles, init_weight
    c.table_seed = seed
    c.uniform_seed = seed
    c.gaussian_table_size = table_size
    c.obs_max_per_pyramid = 100
    c.pi_is_double = 0 if cfg["header"] == "dsp_dynamic.h" else 1
    return c


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class OracleMap:
    def __init__(self, cfg, seed=1, init_particles=0, init_weight=0.01, p_std=0.05, v_std=0.05, ob_std=0.1,
                 newborn_weight=1e-4, newborn_num=20, apply_setters=True, **kw):
        L = self.lib = C.CDLL(build())
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(Config)]
        vp = C.c_void_p
        L.oracle_update.argtypes = ([vp, C.c_int, C.c_int, C.POINTER(C.c_float)] + [C.c_float] * 3 + [C.c_double] +
                                    [C.c_float] * 4 + [C.POINTER(C.c_float), C.c_int])
        L.oracle_get_occupancy.argtypes = [vp, C.c_float, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
        L.oracle_set_prediction_variance.argtypes = [vp, C.c_float, C.c_float]
        L.oracle_set_observation_stddev.argtypes = [vp, C.c_float]
        L.oracle_set_newborn_weight.argtypes = [vp, C.c_float]
        L.oracle_set_newborn_number.argtypes = [vp, C.c_int]
        L.oracle_voxel_index.argtypes = [vp, C.c_float, C.c_float, C.c_float]
        L.oracle_set_cursors.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64]
        L.oracle_set_stage_limit.argtypes = [vp, C.c_int]
        L.oracle_set_last_pose.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_double]
        for f in ("oracle_destroy", "oracle_dims", "oracle_clear_prediction", "oracle_dump_particles",
                  "oracle_load_particles", "oracle_dump_voxel_objects", "oracle_dump_observations",
                  "oracle_dump_pyramid_lists", "oracle_dump_neighbors", "oracle_cursors", "oracle_counters",
                  "oracle_gaussian_tables", "oracle_pdf_table", "oracle_plane_normals", "oracle_voxel_center"):
            getattr(L, f).argtypes = None
        self.cfg = make_config(cfg, seed, init_particles, init_weight, **kw)
        self.h = vp(L.oracle_create(C.byref(self.cfg)))
        d = np.zeros(16, np.int32)
        L.oracle_dims(self.h, _ip(d))
        (self.V, self.S, self.P, self.L, self.T, self.Nh, self.Nv, self.NBW, self.max_ppv, self.nx, self.ny, self.nz,
         self.obs_max, self.static) = [int(x) for x in d[:14]]
        if apply_setters:
            L.oracle_set_prediction_variance(self.h, p_std, v_std)
            L.oracle_set_observation_stddev(self.h, ob_std)
            L.oracle_set_newborn_number(self.h, newborn_num)
            L.oracle_set_newborn_weight(self.h, newborn_weight)

    def __del__(self):
        try:
            self.lib.oracle_destroy(self.h)
        except Exception:
            pass

    def update(self, pts, pos, t, quat, tagged=None, stride=3):
        pts = np.ascontiguousarray(pts, np.float32)
        n = pts.size // stride
        if tagged is not None:
            tagged = np.ascontiguousarray(tagged, np.float32)
        return self.lib.oracle_update(self.h, n, stride, _fp(pts), float(pos[0]), float(pos[1]), float(pos[2]), float(t),
                                      float(quat[0]), float(quat[1]), float(quat[2]), float(quat[3]), _fp(tagged),
                                      0 if tagged is None else tagged.shape[0])

    def occupancy(self, threshold=0.7, with_future=True):
        xyz = np.zeros((self.V, 3), np.float32)
        fut = np.zeros((self.V, self.T), np.float32) if with_future else None
        n = self.lib.oracle_get_occupancy(self.h, threshold, _fp(xyz), self.V, _fp(fut))
        return xyz[:n].copy(), fut

    def clear_prediction(self):
        self.lib.oracle_clear_prediction(self.h)

    def particles(self):
        n = self.lib.oracle_dump_particles(self.h, None, None, 0)
        ids = np.zeros((n, 2), np.int32)
        vals = np.zeros((n, 8), np.float32)
        if n:
            self.lib.oracle_dump_particles(self.h, _ip(ids), _fp(vals), n)
        return ids, vals

    def load_particles(self, ids, vals):
        ids = np.ascontiguousarray(ids, np.int32)
        vals = np.ascontiguousarray(vals, np.float32)
        self.lib.oracle_load_particles(self.h, _ip(ids), _fp(vals), ids.shape[0])

    def voxel_objects(self):
        out = np.zeros((self.V, 4 + self.T), np.float32)
        self.lib.oracle_dump_voxel_objects(self.h, _fp(out))
        return out

    def observations(self):
        cnt = np.zeros(self.P, np.int32)
        mx = np.zeros(self.P, np.float32)
        pts = np.zeros((self.P, self.obs_max, 5), np.float32)
        self.lib.oracle_dump_observations(self.h, _ip(cnt), _fp(mx), _fp(pts))
        return cnt, mx, pts

    def pyramid_lists(self):
        off = np.zeros(self.P + 1, np.int32)
        n = self.lib.oracle_dump_pyramid_lists(self.h, _ip(off), None, 0)
        ent = np.zeros((n, 2), np.int32)
        if n:
            self.lib.oracle_dump_pyramid_lists(self.h, _ip(off), _ip(ent), n)
        return off, ent

    def neighbors(self):
        out = np.zeros((self.P, self.NBW), np.int32)
        self.lib.oracle_dump_neighbors(self.h, _ip(out))
        return out

    def cursors(self):
        c = np.zeros(4, np.int64)
        self.lib.oracle_cursors(self.h, c.ctypes.data_as(C.POINTER(C.c_int64)))
        return c

    def set_cursors(self, p, v, u):
        self.lib.oracle_set_cursors(self.h, int(p), int(v), int(u))

    def set_stage_limit(self, k):
        self.lib.oracle_set_stage_limit(self.h, k)

    def set_last_pose(self, pos, t):
        self.lib.oracle_set_last_pose(self.h, float(pos[0]), float(pos[1]), float(pos[2]), float(t))

    def counters(self):
        c = np.zeros(16, np.int64)
        self.lib.oracle_counters(self.h, c.ctypes.data_as(C.POINTER(C.c_int64)))
        names = ["n_in", "n_left_map", "n_voxel_full", "n_pyramid_full", "n_moved", "n_fov", "n_candidates", "n_born",
                 "n_low_weight", "n_pre", "n_old", "n_out", "n_valid_points"]
        return dict(zip(names, [int(x) for x in c]))

    def gaussian_tables(self, n=None):
        n = n or self.cfg.gaussian_table_size
        p = np.zeros(n, np.float32)
        v = np.zeros(n, np.float32)
        self.lib.oracle_gaussian_tables(self.h, _fp(p), _fp(v), n)
        return p, v

    def pdf_table(self):
        out = np.zeros(20000, np.float32)
        self.lib.oracle_pdf_table(self.h, _fp(out))
        return out

    def plane_normals(self):
        h = np.zeros((self.Nh + 1, 3), np.float32)
        v = np.zeros((self.Nv + 1, 3), np.float32)
        self.lib.oracle_plane_normals(self.h, _fp(h), _fp(v))
        return h, v

    def voxel_index(self, x, y, z):
        return self.lib.oracle_voxel_index(self.h, float(x), float(y), float(z))

    def voxel_center(self, idx):
        out = np.zeros(3, np.float32)
        self.lib.oracle_voxel_center(self.h, int(idx), _fp(out))
        return out
