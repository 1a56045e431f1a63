"""dspmap_b200 — Python host-side mirror of the DSP map's public interface over the C-ABI (include/dspmap_b200.h).

`DSPMap` keeps the reference's method names and argument meaning (g-ch/DSP-map include/dsp_dynamic.h:142-446):
update / getOccupancyMap / getOccupancyMapWithFutureStatus / clearOccupancyMapPrediction / the six setters /
getKMClusterResult / getVoxelPositionFromIndexPublic / getPointVoxelsIndexPublic.  All work happens in
dsp-map_b200/lib/libdspmap_b200.so (hand-written sm_100a kernels); there is no CPU path: importing works without a
GPU (so the symbol table can be checked), creating a map without one raises.
"""
import ctypes as C
import os

import numpy as np

from .configs import CONFIGS, derive  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libdspmap_b200.so")

OK, REJECTED = 1, 0
MAX_T = 8

COUNTER_NAMES = ["n_in", "n_left_map", "n_voxel_full", "n_pyramid_full", "n_moved", "n_fov", "n_candidates", "n_born",
                 "n_low_weight", "n_pre", "n_old", "n_out", "n_valid_points", "overflow", "launches_frame",
                 "launches_total"]


class Config(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("resolution", C.c_float),
                ("angle_resolution", C.c_int32), ("half_fov_h", C.c_int32), ("half_fov_v", C.c_int32),
                ("max_particles_per_voxel", C.c_int32), ("safe_particles_per_voxel", C.c_int32),
                ("safe_particles_per_pyramid", C.c_int32), ("pyramid_neighbor_n", C.c_int32), ("model", C.c_int32),
                ("prediction_times", C.c_int32), ("prediction_future_time", C.c_float * MAX_T),
                ("occlusion_margin", C.c_float), ("init_particle_num", C.c_int32), ("init_weight", C.c_float),
                ("table_seed", C.c_uint64), ("uniform_seed", C.c_uint64), ("gaussian_table_size", C.c_int32),
                ("max_observations_per_pyramid", C.c_int32), ("device", C.c_int32), ("max_points", C.c_int32),
                ("shard_z_begin", C.c_int32), ("shard_z_end", C.c_int32), ("pi_is_double", C.c_int32)]


class DSPMapError(RuntimeError):
    pass


_lib = None


def load_library():
    """Loads the C-ABI library. Fails loudly if it has not been built (python dsp-map_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("DSPMAP_B200_LIB") or LIB_PATH   # (the override serves A/B runs of compile-time variants: tests/ab_variants.sh)
    if not os.path.exists(path):
        raise DSPMapError("%s is missing: build it with `python dsp-map_b200/build.py` (there is no CPU fallback)" % path)
    L = C.CDLL(path)
    vp, f, i, fp, ip = C.c_void_p, C.c_float, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32)
    sig = {
        "dspmap_default_config": (None, [C.POINTER(Config)]),
        "dspmap_create": (i, [C.POINTER(Config), C.POINTER(vp)]),
        "dspmap_destroy": (None, [vp]),
        "dspmap_last_error": (C.c_char_p, []),
        "dspmap_update": (i, [vp, i, i, fp, f, f, f, C.c_double, f, f, f, f]),
        "dspmap_update_tagged": (i, [vp, i, i, fp, f, f, f, C.c_double, f, f, f, f, fp, i]),
        "dspmap_update_device": (i, [vp, i, vp, f, f, f, C.c_double, f, f, f, f, vp, i]),
        "dspmap_set_prediction_variance": (i, [vp, f, f]),
        "dspmap_set_observation_stddev": (i, [vp, f]),
        "dspmap_set_newborn_weight": (i, [vp, f]),
        "dspmap_set_newborn_number": (i, [vp, i]),
        "dspmap_set_particle_record_flag": (i, [vp, i, f, C.c_char_p]),
        "dspmap_set_voxel_filter_resolution": (i, [vp, f]),
        "dspmap_get_occupancy": (i, [vp, f, fp, i, ip, fp]),
        "dspmap_get_occupancy_device": (i, [vp, f, vp, i, vp, vp]),
        "dspmap_get_occupancy_async": (i, [vp, f, i, ip]),
        "dspmap_wait_occupancy": (i, [vp, i, C.POINTER(fp), ip, C.POINTER(fp)]),
        "dspmap_clear_prediction": (i, [vp]),
        "dspmap_last_reader_bytes": (C.c_longlong, [vp]),
        "dspmap_last_update_bytes": (None, [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
        "dspmap_estimator_stats": (i, [vp, ip]),
        "dspmap_timeline": (i, [vp, fp]),
        "dspmap_pin_host_buffer": (i, [vp, vp, C.c_size_t]),
        "dspmap_get_tagged_cloud": (i, [vp, fp, i]),
        "dspmap_voxel_center": (None, [vp, i, fp]),
        "dspmap_voxel_index": (i, [vp, f, f, f, ip]),
        "dspmap_uniform": (f, [vp, f, f]),
        "dspmap_dims": (None, [vp, ip]),
        "dspmap_dump_particles": (i, [vp, ip, fp, i]),
        "dspmap_load_particles": (i, [vp, ip, fp, i]),
        "dspmap_dump_voxel_objects": (i, [vp, fp]),
        "dspmap_dump_observations": (i, [vp, ip, fp, fp]),
        "dspmap_dump_pyramid_lists": (i, [vp, ip, ip, i]),
        "dspmap_dump_plane_normals": (i, [vp, fp, fp]),
        "dspmap_cursors": (i, [vp, C.POINTER(C.c_int64)]),
        "dspmap_set_cursors": (i, [vp, C.c_int64, C.c_int64, C.c_int64]),
        "dspmap_counters": (i, [vp, C.POINTER(C.c_int64)]),
        "dspmap_set_stage_limit": (i, [vp, i]),
        "dspmap_set_last_pose": (i, [vp, f, f, f, C.c_double]),
        "dspmap_set_stream": (i, [vp, vp]),
        "dspmap_synchronize": (i, [vp]),
        "dspmap_profile_enable": (i, [vp, i]),
        "dspmap_profile_read": (i, [vp, C.POINTER(C.c_char_p), fp, ip, i]),
        "dspmap_profile_read_kernels": (i, [vp, C.POINTER(C.c_char_p), fp, ip, i]),
        "dspmap_shard_config": (i, [vp, i, i, vp, vp, i, vp, vp, i, vp, vp]),
        "dspmap_shard_gather_records": (i, [vp, i]),
        "dspmap_shard_phase": (i, [vp, i, i, vp, f, f, f, C.c_double, f, f, f, f, vp, i]),
        "dspmap_shard_unique_id": (i, [vp]),
        "dspmap_shard_init": (i, [vp, i, i, vp, i, i]),
        "dspmap_shard_init_local": (i, [C.POINTER(vp), i, i, i]),
        "dspmap_shard_update": (i, [vp, i, vp, f, f, f, C.c_double, f, f, f, f, vp, i]),
        "dspmap_shard_update_local": (i, [C.POINTER(vp), i, i, vp, f, f, f, C.c_double, f, f, f, f, vp, i]),
        "dspmap_shard_get_occupancy": (i, [vp, f, vp, i, vp, vp]),
        "dspmap_shard_get_occupancy_local": (i, [C.POINTER(vp), i, f, C.POINTER(vp), i, C.POINTER(vp), C.POINTER(vp)]),
        "dspmap_shard_info": (i, [vp, ip]),
        "dspmap_estimator_create": (vp, [C.POINTER(Config), f]),
        "dspmap_estimator_destroy": (None, [vp]),
        "dspmap_estimator_estimate": (i, [vp, i, fp, f, f, f, f, f, f, f, f, fp, i]),
        "dspmap_estimator_set_threaded": (i, [vp, i]),
        "dspmap_euclidean_clusters": (i, [fp, i, f, i, i, i, ip]),
        "dspmap_prefilter_create": (i, [i, i, i, C.c_longlong, C.POINTER(vp)]),
        "dspmap_prefilter_destroy": (None, [vp]),
        "dspmap_prefilter_set_stream": (i, [vp, vp]),
        "dspmap_prefilter_last_error": (C.c_char_p, []),
        "dspmap_prefilter_launches": (C.c_longlong, [vp]),
        "dspmap_prefilter_run": (i, [vp, i, i, fp, f, fp, fp, fp, i, ip]),
        "dspmap_prefilter_run_device": (i, [vp, i, i, vp, f, fp, fp, vp, i, vp]),
        "dspmap_update_raw": (i, [vp, vp, i, i, fp, f, fp, fp, f, f, f, C.c_double, f, f, f, f, ip]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = the library does not export what include/dspmap_b200.h declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "dspmap_default_config", "dspmap_create", "dspmap_destroy", "dspmap_last_error", "dspmap_update",
    "dspmap_update_tagged", "dspmap_update_device", "dspmap_set_prediction_variance", "dspmap_set_observation_stddev",
    "dspmap_set_newborn_weight", "dspmap_set_newborn_number", "dspmap_set_particle_record_flag",
    "dspmap_set_voxel_filter_resolution", "dspmap_get_occupancy", "dspmap_get_occupancy_device",
    "dspmap_get_occupancy_async", "dspmap_wait_occupancy",
    "dspmap_clear_prediction", "dspmap_last_reader_bytes", "dspmap_last_update_bytes", "dspmap_estimator_stats", "dspmap_timeline", "dspmap_pin_host_buffer", "dspmap_get_tagged_cloud", "dspmap_voxel_center", "dspmap_voxel_index", "dspmap_uniform",
    "dspmap_dims", "dspmap_dump_particles", "dspmap_load_particles", "dspmap_dump_voxel_objects",
    "dspmap_dump_observations", "dspmap_dump_pyramid_lists", "dspmap_dump_plane_normals", "dspmap_cursors", "dspmap_set_cursors", "dspmap_counters",
    "dspmap_set_stage_limit", "dspmap_set_last_pose", "dspmap_set_stream", "dspmap_synchronize", "dspmap_profile_enable", "dspmap_profile_read", "dspmap_profile_read_kernels",
    "dspmap_estimator_create", "dspmap_estimator_destroy", "dspmap_estimator_estimate", "dspmap_estimator_set_threaded",
    "dspmap_euclidean_clusters",
    "dspmap_prefilter_create", "dspmap_prefilter_destroy", "dspmap_prefilter_set_stream", "dspmap_prefilter_last_error",
    "dspmap_prefilter_launches", "dspmap_prefilter_run", "dspmap_prefilter_run_device", "dspmap_update_raw",
    "dspmap_shard_config", "dspmap_shard_gather_records", "dspmap_shard_phase", "dspmap_shard_unique_id", "dspmap_shard_init",
    "dspmap_shard_init_local", "dspmap_shard_update", "dspmap_shard_update_local", "dspmap_shard_get_occupancy",
    "dspmap_shard_get_occupancy_local", "dspmap_shard_info",
]


def make_config(cfg, seed=1, init_particles=0, init_weight=0.01, device=0, max_points=0, safe_ppv=0, safe_pyramid=0,
                table_size=10000000):
    """cfg: a dict from configs.CONFIGS (the reference's compile-time parameters)."""
    c = Config()
    c.nx, c.ny, c.nz = cfg["nx"], cfg["ny"], cfg["nz"]
    c.resolution = cfg["res"]
    c.angle_resolution = cfg["angle_res"]
    c.half_fov_h, c.half_fov_v = cfg["half_fov_h"], cfg["half_fov_v"]
    c.max_particles_per_voxel = cfg["max_ppv"]
    c.safe_particles_per_voxel = safe_ppv
    c.safe_particles_per_pyramid = safe_pyramid
    c.pyramid_neighbor_n = cfg["neighbor_n"]
    c.model = 1 if cfg["model"] == "static" else 0
    c.prediction_times = len(cfg["future_times"])
    for k, t in enumerate(cfg["future_times"]):
        c.prediction_future_time[k] = t
    c.occlusion_margin = 0.3 if cfg["header"] == "dsp_dynamic.h" else cfg["res"]
    c.init_particle_num, c.init_weight = init_particles, init_weight
    c.table_seed = seed
    c.uniform_seed = seed
    c.gaussian_table_size = table_size
    c.max_observations_per_pyramid = 100
    c.device = device
    c.max_points = max_points
    c.pi_is_double = 0 if cfg["header"] == "dsp_dynamic.h" else 1   # dyn:543 with glibc's float M_PIf32; mn:78 / st:74 double
    return c


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


class DSPMap:
    """The reference's `class DSPMap` (dsp_dynamic.h:142) over the B200 library."""

    def __init__(self, cfg, seed=1, init_particle_num=0, init_weight=0.01, device=0, **kw):
        self.lib = load_library()
        self.cfg_dict = cfg
        self.config = make_config(cfg, seed, init_particle_num, init_weight, device, **kw)
        h = C.c_void_p()
        rc = self.lib.dspmap_create(C.byref(self.config), C.byref(h))
        if rc != OK:
            raise DSPMapError("dspmap_create failed (%d): %s" % (rc, self.lib.dspmap_last_error().decode()))
        self.h = h
        d = np.zeros(16, np.int32)
        self.lib.dspmap_dims(self.h, _ip(d))
        (self.V, self.S, self.P, self.L, self.T, self.Nh, self.Nv, self.NBW, self.max_ppv, self.nx, self.ny, self.nz,
         self.obs_max, self.static) = [int(x) for x in d[:14]]

    def close(self):
        if getattr(self, "h", None):
            if getattr(self, "_pinned", None) is not None:
                self.lib.dspmap_pin_host_buffer(self.h, None, 0)
                self._pinned = None
            self.lib.dspmap_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise DSPMapError("dspmap call failed (%d): %s" % (rc, self.lib.dspmap_last_error().decode()))
        return rc

    # ---- the reference's public methods -----------------------------------------------------------------------
    def update(self, point_cloud_num, size_of_one_point, point_cloud, sensor_px, sensor_py, sensor_pz,
               time_stamp_second, qw, qx, qy, qz, tagged=None):
        """DSPMap::update (dsp_dynamic.h:181). Returns 1 / 0 like the reference. `tagged` (n x 7, world frame)
        overrides the built-in velocity estimation with an explicit newborn input."""
        pts = np.ascontiguousarray(point_cloud, np.float32)
        if tagged is None:
            return self._check(self.lib.dspmap_update(self.h, point_cloud_num, size_of_one_point, _fp(pts), sensor_px,
                                                      sensor_py, sensor_pz, time_stamp_second, qw, qx, qy, qz))
        tg = np.ascontiguousarray(tagged, np.float32)
        return self._check(self.lib.dspmap_update_tagged(self.h, point_cloud_num, size_of_one_point, _fp(pts), sensor_px,
                                                         sensor_py, sensor_pz, time_stamp_second, qw, qx, qy, qz,
                                                         _fp(tg), tg.shape[0]))

    def setPredictionVariance(self, p_stddev, v_stddev):
        self._check(self.lib.dspmap_set_prediction_variance(self.h, p_stddev, v_stddev))

    def setObservationStdDev(self, ob_stddev):
        self._check(self.lib.dspmap_set_observation_stddev(self.h, ob_stddev))

    def setNewBornParticleWeight(self, weight):
        self._check(self.lib.dspmap_set_newborn_weight(self.h, weight))

    def setNewBornParticleNumberofEachPoint(self, num):
        self._check(self.lib.dspmap_set_newborn_number(self.h, num))

    def setParticleRecordFlag(self, record_particle_flag, record_csv_time=1.0, folder="."):
        sep = "/" if self.cfg_dict.get("header", "dsp_dynamic.h") == "dsp_dynamic.h" else ""   # dyn:333 vs mn:335 / st:330
        self._check(self.lib.dspmap_set_particle_record_flag(self.h, record_particle_flag, record_csv_time, (folder + sep).encode()))

    def setOriginalVoxelFilterResolution(self, res):
        self._check(self.lib.dspmap_set_voxel_filter_resolution(self.h, res))

    def getOccupancyMap(self, threshold=0.7):
        """Returns (obstacles_num, cloud[n,3]); zeroes the future columns like the reference (dsp_dynamic.h:385-402)."""
        xyz = getattr(self, "_xyz_buf", None)
        if xyz is None:
            xyz = self._xyz_buf = np.zeros((self.V, 3), np.float32)
        n = C.c_int32(0)
        self._check(self.lib.dspmap_get_occupancy(self.h, threshold, _fp(xyz), self.V, C.byref(n), None))
        return n.value, xyz[:n.value].copy()

    def getOccupancyMapWithFutureStatus(self, threshold=0.7, future_status=None):
        """Returns (obstacles_num, cloud[n,3], future_status[V,T]) (dsp_dynamic.h:405-426)."""
        xyz = getattr(self, "_xyz_buf", None)   # reused: the library fills the first n rows, which are copied out below
        if xyz is None:
            xyz = self._xyz_buf = np.zeros((self.V, 3), np.float32)
        if future_status is None:
            future_status = np.zeros((self.V, self.T), np.float32)
        n = C.c_int32(0)
        self._check(self.lib.dspmap_get_occupancy(self.h, threshold, _fp(xyz), self.V, C.byref(n), _fp(future_status)))
        return n.value, xyz[:n.value].copy(), future_status

    def pin_host_buffer(self, arr):
        """Page-locks a numpy array used as future_status so the reader DMAs straight into it; keep `arr` alive."""
        self._pinned = arr
        return self.lib.dspmap_pin_host_buffer(self.h, C.c_void_p(arr.ctypes.data if arr is not None else 0), arr.nbytes if arr is not None else 0)

    def clearOccupancyMapPrediction(self):
        self._check(self.lib.dspmap_clear_prediction(self.h))

    def getKMClusterResult(self):
        n = self.lib.dspmap_get_tagged_cloud(self.h, None, 0)
        out = np.zeros((n, 7), np.float32)
        if n:
            self.lib.dspmap_get_tagged_cloud(self.h, _fp(out), n)
        return out

    def getVoxelPositionFromIndexPublic(self, index):
        out = np.zeros(3, np.float32)
        self.lib.dspmap_voxel_center(self.h, int(index), _fp(out))
        return out

    def getPointVoxelsIndexPublic(self, px, py, pz):
        idx = C.c_int32(-1)
        ok = self.lib.dspmap_voxel_index(self.h, px, py, pz, C.byref(idx))
        return ok, idx.value

    # ---- state access (tests, checkpointing) ------------------------------------------------------------------
    def particles(self):
        n = self._check(self.lib.dspmap_dump_particles(self.h, None, None, 0))
        ids = np.zeros((n, 2), np.int32)
        vals = np.zeros((n, 8), np.float32)
        if n:
            self._check(self.lib.dspmap_dump_particles(self.h, _ip(ids), _fp(vals), n))
        return ids, vals

    def load_particles(self, ids, vals):
        ids = np.ascontiguousarray(ids, np.int32)
        vals = np.ascontiguousarray(vals, np.float32)
        self._check(self.lib.dspmap_load_particles(self.h, _ip(ids), _fp(vals), ids.shape[0]))

    def voxel_objects(self):
        out = np.zeros((self.V, 4 + self.T), np.float32)
        self._check(self.lib.dspmap_dump_voxel_objects(self.h, _fp(out)))
        return out

    def observations(self):
        cnt = np.zeros(self.P, np.int32)
        mx = np.zeros(self.P, np.float32)
        pts = np.zeros((self.P, self.obs_max, 5), np.float32)
        self._check(self.lib.dspmap_dump_observations(self.h, _ip(cnt), _fp(mx), _fp(pts)))
        return cnt, mx, pts

    def pyramid_lists(self):
        off = np.zeros(self.P + 1, np.int32)
        n = self._check(self.lib.dspmap_dump_pyramid_lists(self.h, _ip(off), None, 0))
        ent = np.zeros((n, 2), np.int32)
        if n:
            self._check(self.lib.dspmap_dump_pyramid_lists(self.h, _ip(off), _ip(ent), n))
        return off, ent

    def plane_normals(self):
        h = np.zeros((self.Nh + 1, 3), np.float32)
        v = np.zeros((self.Nv + 1, 3), np.float32)
        self._check(self.lib.dspmap_dump_plane_normals(self.h, _fp(h), _fp(v)))
        return h, v

    def cursors(self):
        c = np.zeros(4, np.int64)
        self._check(self.lib.dspmap_cursors(self.h, c.ctypes.data_as(C.POINTER(C.c_int64))))
        return c

    def set_cursors(self, p, v, u):
        self._check(self.lib.dspmap_set_cursors(self.h, int(p), int(v), int(u)))

    def counters(self):
        c = np.zeros(16, np.int64)
        self._check(self.lib.dspmap_counters(self.h, c.ctypes.data_as(C.POINTER(C.c_int64))))
        return dict(zip(COUNTER_NAMES, [int(x) for x in c]))

    def last_future_bytes(self):
        """Device-to-host bytes of the last blocking reader call (dspmap_last_reader_bytes)."""
        return int(self.lib.dspmap_last_reader_bytes(self.h))

    def last_update_bytes(self):
        """(host-to-device, device-to-host) bytes of the last update() call (dspmap_last_update_bytes)."""
        a, b = C.c_longlong(0), C.c_longlong(0)
        self.lib.dspmap_last_update_bytes(self.h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def estimator_stats(self):
        """(on_device, dict of the device front end's counters for the last frame) — dspmap_estimator_stats."""
        d = np.zeros(8, np.int32)
        on = int(self.lib.dspmap_estimator_stats(self.h, _ip(d)))
        names = ("in_view", "clusters", "ground_points", "dynamic_clusters", "dynamic_points", "static_cluster_points", "cells", "tagged")
        return on, {k: int(v) for k, v in zip(names, d)}

    def timeline(self):
        """ms from the start of the last frame to its milestones (dspmap_timeline; maps created with DSPMAP_TIMELINE=1)."""
        d = np.zeros(8, np.float32)
        if self.lib.dspmap_timeline(self.h, _fp(d)) != 1:
            return None
        names = ("features_on_host", "arrived", "obs_binned", "newborn_placed", "weights_start", "norm_done", "frame_end")
        return {k: round(float(v), 4) for k, v in zip(names, d)}

    def fast_paths(self):
        """(fast_res, fast_sigma): whether the exhaustively verified exact fast divisions are enabled (sigma: after the next update)."""
        d = np.zeros(16, np.int32)
        self.lib.dspmap_dims(self.h, _ip(d))
        return int(d[14]), int(d[15])

    def set_stage_limit(self, k):
        self._check(self.lib.dspmap_set_stage_limit(self.h, k))

    def set_last_pose(self, pos, t):
        self._check(self.lib.dspmap_set_last_pose(self.h, float(pos[0]), float(pos[1]), float(pos[2]), float(t)))

    def set_stream(self, cuda_stream):
        self._check(self.lib.dspmap_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self.lib.dspmap_synchronize(self.h))

    def profile_enable(self, on=True):
        self._check(self.lib.dspmap_profile_enable(self.h, 1 if on else 0))

    def profile_read(self):
        names = (C.c_char_p * 32)()
        ms = np.zeros(32, np.float32)
        ln = np.zeros(32, np.int32)
        n = self.lib.dspmap_profile_read(self.h, names, _fp(ms), _ip(ln), 32)
        return {names[k].decode(): (float(ms[k]), int(ln[k])) for k in range(n)}

    def profile_read_kernels(self):
        """{kernel name: (summed ms, launches)} measured with CUDA events on the launching stream while profiling is on."""
        names = (C.c_char_p * 96)()
        ms = np.zeros(96, np.float32)
        ln = np.zeros(96, np.int32)
        n = self.lib.dspmap_profile_read_kernels(self.h, names, _fp(ms), _ip(ln), 96)
        return {names[k].decode(): (float(ms[k]), int(ln[k])) for k in range(n)}

    # device-resident entry points (bench): pointers are raw device addresses (e.g. torch tensor .data_ptr())
    def update_device(self, n, d_pts, pos, t, quat, d_tagged, n_tagged):
        return self._check(self.lib.dspmap_update_device(self.h, n, C.c_void_p(d_pts), float(pos[0]), float(pos[1]),
                                                         float(pos[2]), float(t), float(quat[0]), float(quat[1]),
                                                         float(quat[2]), float(quat[3]), C.c_void_p(d_tagged), n_tagged))

    def shard_config(self, rank, nranks, xsend, xrecv, cap_x, gsend, grecv, cap_g, czinv, shared):
        return self._check(self.lib.dspmap_shard_config(self.h, rank, nranks, C.c_void_p(xsend), C.c_void_p(xrecv), cap_x,
                                                        C.c_void_p(gsend), C.c_void_p(grecv), cap_g, C.c_void_p(czinv),
                                                        C.c_void_p(shared)))

    def shard_gather_records(self, records):
        return self._check(self.lib.dspmap_shard_gather_records(self.h, records))

    def shard_phase(self, phase, n, d_pts, pos, t, quat, d_tagged, n_tagged):
        return self._check(self.lib.dspmap_shard_phase(self.h, phase, n, C.c_void_p(d_pts), float(pos[0]), float(pos[1]),
                                                       float(pos[2]), float(t), float(quat[0]), float(quat[1]), float(quat[2]),
                                                       float(quat[3]), C.c_void_p(d_tagged), n_tagged))

    # ---- one map over several GPUs, orchestrated by the library (include/dspmap_b200.h: dspmap_shard_init / _update) ------
    def shard_init(self, rank, nranks, nccl_id, cap_x=0, cap_g=0):
        """nccl_id: the 128 bytes of shard_unique_id() from rank 0."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(nccl_id))
        return self._check(self.lib.dspmap_shard_init(self.h, rank, nranks, C.cast(buf, C.c_void_p), cap_x, cap_g))

    def shard_update(self, n, d_pts, pos, t, quat, d_tagged, n_tagged):
        return self._check(self.lib.dspmap_shard_update(self.h, n, C.c_void_p(d_pts), float(pos[0]), float(pos[1]), float(pos[2]),
                                                        float(t), float(quat[0]), float(quat[1]), float(quat[2]), float(quat[3]),
                                                        C.c_void_p(d_tagged), n_tagged))

    def shard_get_occupancy(self, threshold, d_xyz, cap, d_count, d_future):
        return self._check(self.lib.dspmap_shard_get_occupancy(self.h, threshold, C.c_void_p(d_xyz), cap, C.c_void_p(d_count),
                                                               C.c_void_p(d_future)))

    def shard_info(self):
        o = np.zeros(4, np.int32)
        self._check(self.lib.dspmap_shard_info(self.h, _ip(o)))
        return dict(cap_x=int(o[0]), cap_g=int(o[1]), gather_records=int(o[2]), frames=int(o[3]))

    def get_occupancy_device(self, threshold, d_xyz, cap, d_count, d_future):
        return self._check(self.lib.dspmap_get_occupancy_device(self.h, threshold, C.c_void_p(d_xyz), cap,
                                                                C.c_void_p(d_count), C.c_void_p(d_future)))


    def get_occupancy_async(self, threshold, with_future=True):
        """Pipelined reader: enqueues the reader kernels + the copies on a second stream; returns a ticket."""
        t = C.c_int32(-1)
        self._check(self.lib.dspmap_get_occupancy_async(self.h, threshold, 1 if with_future else 0, C.byref(t)))
        return t.value

    def wait_occupancy(self, ticket):
        """(count, (count, 3) voxel centres, (V, T) future status or None): zero-copy views of the library's page-locked
        slot, valid until the second get_occupancy_async call after the one that returned `ticket`."""
        xyz, fut, n = C.POINTER(C.c_float)(), C.POINTER(C.c_float)(), C.c_int32(0)
        self._check(self.lib.dspmap_wait_occupancy(self.h, ticket, C.byref(xyz), C.byref(n), C.byref(fut)))
        a = np.ctypeslib.as_array(xyz, shape=(max(n.value, 1), 3))[:n.value]
        f_ = np.ctypeslib.as_array(fut, shape=(self.V, self.T)) if fut else None
        return n.value, a, f_


class VelocityEstimator:
    """Host-only velocity estimation (the reference's side thread, dsp_dynamic.h:1377-1544); needs no GPU."""

    def __init__(self, cfg, seed=1, filter_res=0.1):
        self.lib = load_library()
        self.config = make_config(cfg, seed)
        self.h = C.c_void_p(self.lib.dspmap_estimator_create(C.byref(self.config), filter_res))
        self.last_t = None
        self.cap = 1 << 16

    def set_threaded(self, on=True):
        """Runs estimate() on the library's persistent helper thread (same results; see include/dspmap_b200.h)."""
        return self.lib.dspmap_estimator_set_threaded(self.h, 1 if on else 0)

    def estimate(self, pts, pos, t, quat):
        """Returns the tagged cloud (n x 7) of this frame, or None when no point is in view."""
        pts = np.ascontiguousarray(pts, np.float32)
        dt = np.float32(0.0) if self.last_t is None else np.float32(float(t) - float(self.last_t))
        self.last_t = t
        out = np.zeros((max(len(pts), 1), 7), np.float32)
        n = self.lib.dspmap_estimator_estimate(self.h, len(pts), _fp(pts), float(pos[0]), float(pos[1]), float(pos[2]),
                                               float(dt), float(quat[0]), float(quat[1]), float(quat[2]), float(quat[3]),
                                               _fp(out), out.shape[0])
        return None if n < 0 else out[:n].copy()

    def __del__(self):
        try:
            self.lib.dspmap_estimator_destroy(self.h)
        except Exception:
            pass


class Prefilter:
    """The application's preprocessing (map_sim_example.cpp:305-336: VoxelGrid, axis swap, crop, cut) on the GPU."""

    def __init__(self, max_raw_points=640 * 480, max_stride=4, max_out_points=5000, max_leaves=0, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.dspmap_prefilter_create(device, max_raw_points * max_stride, max_out_points, max_leaves, C.byref(h))
        if rc != OK:
            raise DSPMapError("dspmap_prefilter_create failed (%d): %s" % (rc, self.lib.dspmap_prefilter_last_error().decode()))
        self.h, self.cap = h, max_out_points

    def close(self):
        if getattr(self, "h", None):
            self.lib.dspmap_prefilter_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise DSPMapError("prefilter call failed (%d): %s" % (rc, self.lib.dspmap_prefilter_last_error().decode()))
        return rc

    def set_stream(self, cuda_stream):
        self._check(self.lib.dspmap_prefilter_set_stream(self.h, C.c_void_p(cuda_stream)))

    def launches(self):
        return int(self.lib.dspmap_prefilter_launches(self.h))

    def run(self, pts, leaf, range_min, range_max, cap=None):
        """pts (n, stride) float32 host array, camera frame -> (m, 3) float32 cloud for DSPMap.update."""
        pts = np.ascontiguousarray(pts, np.float32)
        pts = pts.reshape(-1, 3) if pts.ndim == 1 else pts
        cap = self.cap if cap is None else min(cap, self.cap)
        lo, hi = np.ascontiguousarray(range_min, np.float32), np.ascontiguousarray(range_max, np.float32)
        out, n = np.zeros((cap, 3), np.float32), C.c_int32(0)
        self._check(self.lib.dspmap_prefilter_run(self.h, pts.shape[0], pts.shape[1], _fp(pts), float(leaf), _fp(lo), _fp(hi),
                                                  _fp(out), cap, C.byref(n)))
        return out[:n.value]

    def run_device(self, n, stride, d_pts, leaf, range_min, range_max, d_out, cap, d_n_out):
        lo, hi = np.ascontiguousarray(range_min, np.float32), np.ascontiguousarray(range_max, np.float32)
        return self._check(self.lib.dspmap_prefilter_run_device(self.h, n, stride, C.c_void_p(d_pts), float(leaf), _fp(lo), _fp(hi),
                                                                C.c_void_p(d_out), cap, C.c_void_p(d_n_out)))

    def update_raw(self, m, pts, leaf, range_min, range_max, pos, t, quat):
        """dspmap_update_raw: preprocessing + DSPMap.update in one call. Returns (update's return code, filtered count)."""
        pts = np.ascontiguousarray(pts, np.float32)
        lo, hi = np.ascontiguousarray(range_min, np.float32), np.ascontiguousarray(range_max, np.float32)
        nf = C.c_int32(0)
        rc = self.lib.dspmap_update_raw(m.h, self.h, pts.shape[0], pts.shape[1], _fp(pts), float(leaf), _fp(lo), _fp(hi),
                                        float(pos[0]), float(pos[1]), float(pos[2]), float(t), float(quat[0]), float(quat[1]),
                                        float(quat[2]), float(quat[3]), C.byref(nf))
        if rc < 0:
            raise DSPMapError("dspmap_update_raw failed (%d): %s / %s" % (rc, self.lib.dspmap_prefilter_last_error().decode(),
                                                                         self.lib.dspmap_last_error().decode()))
        return rc, nf.value


def euclidean_clusters(xyz, tolerance, min_size=1, max_size=1 << 30, path=0):
    """The estimator's clustering on its own: (cluster count, int32 label per point, -1 = not in a kept cluster)."""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    labels = np.full(len(xyz), -1, np.int32)
    n = load_library().dspmap_euclidean_clusters(_fp(xyz), len(xyz), float(tolerance), int(min_size), int(max_size), int(path), _ip(labels))
    if n < 0:
        raise DSPMapError("dspmap_euclidean_clusters: path %d cannot take this cloud (%d)" % (path, n))
    return n, labels


def bytes_per_update(counters, V, T, M):
    """Algorithmic HBM bytes of one update:
    64 N_in + 36 N_fov + 32 N_born + 32 N_pre + 32 N_out + 4 T N_old + V (20 + 12 T) + 20 M."""
    c = counters
    return (64 * c["n_in"] + 36 * c["n_fov"] + 32 * c["n_born"] + 32 * c["n_pre"] + 32 * c["n_out"] +
            4 * T * c["n_old"] + V * (20 + 12 * T) + 20 * M)


def shard_unique_id():
    """ncclGetUniqueId through the library (rank 0 calls it and hands the 128 bytes to every rank)."""
    lib = load_library()
    buf = (C.c_char * 128)()
    rc = lib.dspmap_shard_unique_id(C.cast(buf, C.c_void_p))
    if rc != OK:
        raise DSPMapError("dspmap_shard_unique_id failed (%d): %s" % (rc, lib.dspmap_last_error().decode()))
    return bytes(buf)


class LocalShardedMap:
    """All shards of one map in ONE process on ONE device, driven by the library's own C++ orchestrator
    (dspmap_shard_init_local / _update_local / _get_occupancy_local): the code path of the NCCL build with the collectives
    done as device-to-device copies.  For tests on a single-GPU box."""

    def __init__(self, cfg, nranks, seed=1, device=0, max_points=0, setters=None, cap_x=0, cap_g=0, **kw):
        import torch
        self.torch = torch
        self.cfg, self.nranks = cfg, nranks
        self.maps = [DSPMap(cfg, seed=seed, device=device, max_points=max_points or 65536, **kw) for _ in range(nranks)]
        self.lib = self.maps[0].lib
        for m in self.maps:
            if setters:
                setters(m)
        self.hs = (C.c_void_p * nranks)(*[m.h for m in self.maps])
        rc = self.lib.dspmap_shard_init_local(self.hs, nranks, cap_x, cap_g)
        if rc != OK:
            raise DSPMapError("dspmap_shard_init_local failed (%d): %s" % (rc, self.lib.dspmap_last_error().decode()))
        self.dev = torch.device("cuda", device)
        m = self.maps[0]
        self.V, self.T = m.V, m.T
        self.d_xyz = [torch.zeros((m.V, 3), dtype=torch.float32, device=self.dev) for _ in range(nranks)]
        self.d_cnt = [torch.zeros(1, dtype=torch.int32, device=self.dev) for _ in range(nranks)]
        self.d_fut = [torch.zeros((m.V, max(m.T, 1)), dtype=torch.float32, device=self.dev) for _ in range(nranks)]
        zpr = (cfg["nz"] + nranks - 1) // nranks
        self.v_lo = [min(cfg["nz"], r * zpr) * cfg["nx"] * cfg["ny"] for r in range(nranks)]
        self.v_hi = [min(cfg["nz"], (r + 1) * zpr) * cfg["nx"] * cfg["ny"] for r in range(nranks)]

    def update(self, pts, pos, t, quat, tagged):
        torch = self.torch
        d_pts = torch.from_numpy(np.ascontiguousarray(pts, np.float32)).to(self.dev)
        tg = np.ascontiguousarray(tagged, np.float32).reshape(-1, 7)
        d_tag = torch.from_numpy(tg if len(tg) else np.zeros((1, 7), np.float32)).to(self.dev)
        torch.cuda.synchronize()
        rc = self.lib.dspmap_shard_update_local(self.hs, self.nranks, len(pts), C.c_void_p(d_pts.data_ptr()), float(pos[0]), float(pos[1]),
                                                float(pos[2]), float(t), float(quat[0]), float(quat[1]), float(quat[2]), float(quat[3]),
                                                C.c_void_p(d_tag.data_ptr()), len(tg))
        if rc < 0:
            raise DSPMapError("dspmap_shard_update_local failed (%d): %s" % (rc, self.lib.dspmap_last_error().decode()))
        for m in self.maps:
            m.synchronize()
        return rc

    def occupancy(self, threshold):
        """Every rank's copy of the whole map's results (they must all be equal): list of (n, xyz, future)."""
        n = self.nranks
        xs = (C.c_void_p * n)(*[x.data_ptr() for x in self.d_xyz])
        cs = (C.c_void_p * n)(*[x.data_ptr() for x in self.d_cnt])
        fs = (C.c_void_p * n)(*[x.data_ptr() for x in self.d_fut])
        rc = self.lib.dspmap_shard_get_occupancy_local(self.hs, n, threshold, xs, self.V, cs, fs)
        if rc < 0:
            raise DSPMapError("dspmap_shard_get_occupancy_local failed (%d): %s" % (rc, self.lib.dspmap_last_error().decode()))
        for m in self.maps:
            m.synchronize()
        out = []
        for r in range(n):
            k = int(self.d_cnt[r].item())
            out.append((k, self.d_xyz[r][:k].cpu().numpy(), self.d_fut[r].cpu().numpy()))
        return out

    def particles(self):
        parts = [m.particles() for m in self.maps]
        ids = np.concatenate([p[0] for p in parts])
        vals = np.concatenate([p[1] for p in parts])
        order = np.argsort(ids[:, 0].astype(np.int64) * 128 + ids[:, 1], kind="stable")
        return ids[order], vals[order]

    def voxel_objects(self):
        out = None
        for r, m in enumerate(self.maps):
            vo = m.voxel_objects()
            if out is None:
                out = np.zeros_like(vo)
            out[self.v_lo[r]:self.v_hi[r], :4] = vo[self.v_lo[r]:self.v_hi[r], :4]
            out[:, 4:] += vo[:, 4:]
        return out

    def close(self):
        for m in self.maps:
            m.close()
