"""ctypes loader for oracle/_ref/libdspref_<cfg>.so (the unmodified reference header compiled against the stand-in
dependency headers, see oracle/build_ref.py).  TEST INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by bench.py's reference / cpu_baseline legs, never by the product package.

The reference keeps all map state in file-static arrays and function-static variables, so one loaded library is one
map for the life of the process; RefMap() therefore loads a private temporary copy of the .so per instance.
"""
import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available(cfg_name):
    return os.path.exists(os.path.join(REF_DIR, "libdspref_%s.so" % cfg_name))


def fast_variant(cfg_name):
    """Name of the timing-only build with the reference's own compiler flags (oracle/build_ref.py: FAST_FLAGS), or None
    when it has not been built or this host lacks AVX2 / FMA."""
    if not available(cfg_name + "_fast"):
        return None
    try:
        flags = next(line for line in open("/proc/cpuinfo") if line.startswith("flags")).split()
    except (OSError, StopIteration):
        return None
    return cfg_name + "_fast" if "avx2" in flags and "fma" in flags else None


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class RefMap:
    def __init__(self, cfg_name, seed=1, init_particles=0, init_weight=0.01, p_std=0.05, v_std=0.05, ob_std=0.1,
                 newborn_weight=1e-4, newborn_num=20, filter_res=0.1, apply_setters=True):
        src = os.path.join(REF_DIR, "libdspref_%s.so" % cfg_name)
        if not os.path.exists(src):
            raise FileNotFoundError(src + " (run python oracle/build_ref.py where /root/reference exists)")
        fd, self._tmp = tempfile.mkstemp(prefix="dspref_%s_" % cfg_name, suffix=".so")
        os.close(fd)
        shutil.copyfile(src, self._tmp)
        L = self.lib = C.CDLL(self._tmp)
        os.unlink(self._tmp)
        L.ref_create.argtypes = [C.c_uint64, C.c_int, C.c_float]
        L.ref_update.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)] + [C.c_float] * 3 + [C.c_double] + [C.c_float] * 4
        L.ref_timed_frame.argtypes = ([C.c_int, C.c_int, C.POINTER(C.c_float)] + [C.c_float] * 3 + [C.c_double] +
                                      [C.c_float] * 4 + [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_int)])
        L.ref_timed_frame.restype = C.c_double
        L.ref_get_occupancy.argtypes = [C.c_float, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
        L.ref_resolution.restype = C.c_float
        L.ref_uniform.restype = C.c_float
        L.ref_uniform.argtypes = [C.c_float, C.c_float]
        L.ref_voxel_index.argtypes = [C.c_float] * 3
        for f, t in (("ref_set_prediction_variance", [C.c_float, C.c_float]), ("ref_set_observation_stddev", [C.c_float]),
                     ("ref_set_newborn_weight", [C.c_float]), ("ref_set_newborn_number", [C.c_int]),
                     ("ref_set_voxel_filter_resolution", [C.c_float]),
                     ("ref_set_particle_record_flag", [C.c_int, C.c_float, C.c_char_p])):
            getattr(L, f).argtypes = t
        d = np.zeros(16, np.int32)
        L.ref_dims(_ip(d))
        (self.V, self.S, self.P, self.L, self.T, self.Nh, self.Nv, self.NBW, self.max_ppv, self.nx, self.ny, self.nz,
         self.obs_max, self.static) = [int(x) for x in d[:14]]
        self.res = float(L.ref_resolution())
        assert L.ref_create(seed, init_particles, init_weight) == 1
        if apply_setters:  # src/map_sim_example.cpp:522-526
            L.ref_set_prediction_variance(p_std, v_std)
            L.ref_set_observation_stddev(ob_std)
            L.ref_set_newborn_number(newborn_num)
            L.ref_set_newborn_weight(newborn_weight)
            L.ref_set_voxel_filter_resolution(filter_res)

    def update(self, pts, pos, t, quat, stride=3):
        pts = np.ascontiguousarray(pts, np.float32)
        n = pts.size // stride
        return self.lib.ref_update(n, stride, _fp(pts), float(pos[0]), float(pos[1]), float(pos[2]), float(t),
                                   float(quat[0]), float(quat[1]), float(quat[2]), float(quat[3]))

    def timed_frame(self, pts, pos, t, quat, threshold, future):
        pts = np.ascontiguousarray(pts, np.float32)
        n_occ = C.c_int(0)
        s = self.lib.ref_timed_frame(pts.size // 3, 3, _fp(pts), float(pos[0]), float(pos[1]), float(pos[2]), float(t),
                                     float(quat[0]), float(quat[1]), float(quat[2]), float(quat[3]), threshold,
                                     _fp(future), C.byref(n_occ))
        return s, n_occ.value

    def occupancy(self, threshold=0.7, with_future=True):
        xyz = np.zeros((self.V, 3), np.float32)
        fut = np.zeros((self.V, self.T), np.float32) if with_future else None
        n = self.lib.ref_get_occupancy(threshold, _fp(xyz), self.V, _fp(fut) if with_future else None)
        return xyz[:n].copy(), fut

    def clear_prediction(self):
        self.lib.ref_clear_prediction()

    def set_particle_record_flag(self, flag, record_time=1.0, folder="."):
        """setParticleRecordFlag (dsp_dynamic.h:375) with the header's global particle_save_folder set to `folder`."""
        self.lib.ref_set_particle_record_flag(int(flag), float(record_time), folder.encode())

    def tagged_cloud(self):
        n = self.lib.ref_tagged_cloud(None, 0)
        out = np.zeros((n, 7), np.float32)
        if n:
            self.lib.ref_tagged_cloud(_fp(out), n)
        return out

    def particles(self):
        n = self.lib.ref_dump_particles(None, None, 0)
        ids = np.zeros((n, 2), np.int32)
        vals = np.zeros((n, 8), np.float32)
        if n:
            self.lib.ref_dump_particles(_ip(ids), _fp(vals), n)
        return ids, vals

    def voxel_objects(self):
        out = np.zeros((self.V, 4 + self.T), np.float32)
        self.lib.ref_dump_voxel_objects(_fp(out))
        return out

    def observations(self):
        cnt = np.zeros(self.P, np.int32)
        mx = np.zeros(self.P, np.float32)
        pts = np.zeros((self.P, self.obs_max, 5), np.float32)
        self.lib.ref_dump_observations(_ip(cnt), _fp(mx), _fp(pts))
        return cnt, mx, pts

    def pyramid_lists(self):
        off = np.zeros(self.P + 1, np.int32)
        n = self.lib.ref_dump_pyramid_lists(_ip(off), None, 0)
        ent = np.zeros((n, 2), np.int32)
        if n:
            self.lib.ref_dump_pyramid_lists(_ip(off), _ip(ent), n)
        return off, ent

    def neighbors(self):
        out = np.zeros((self.P, self.NBW), np.int32)
        self.lib.ref_dump_neighbors(_ip(out))
        return out

    def cursors(self):
        c = np.zeros(4, np.int64)
        self.lib.ref_cursors(c.ctypes.data_as(C.POINTER(C.c_int64)))
        return c

    def gaussian_tables(self, n=10000000):
        p = np.zeros(n, np.float32)
        v = np.zeros(n, np.float32)
        self.lib.ref_gaussian_tables(_fp(p), _fp(v), n)
        return p, v

    def pdf_table(self):
        out = np.zeros(20000, np.float32)
        self.lib.ref_pdf_table(_fp(out))
        return out

    def plane_normals(self):
        h = np.zeros((self.Nh + 1, 3), np.float32)
        v = np.zeros((self.Nv + 1, 3), np.float32)
        self.lib.ref_plane_normals(_fp(h), _fp(v))
        return h, v

    def voxel_index(self, x, y, z):
        return self.lib.ref_voxel_index(float(x), float(y), float(z))

    def voxel_center(self, idx):
        out = np.zeros(3, np.float32)
        self.lib.ref_voxel_center(int(idx), _fp(out))
        return out
