"""Map configurations: the five BASELINE.json workloads plus small edge-case maps used by the parity tests.

Field meaning follows the reference's compile-time parameters (include/dsp_dynamic.h:38-66):
nx/ny/nz = MAP_{LENGTH,WIDTH,HEIGHT}_VOXEL_NUM, res = VOXEL_RESOLUTION, angle_res = ANGLE_RESOLUTION,
max_ppv = MAX_PARTICLE_NUM_VOXEL, half_fov_h/v (degrees), future_times = prediction_future_time,
neighbor_n = PYRAMID_NEIGHBOR_N (1 for dsp_dynamic.h's fixed 3x3 block, dsp_dynamic.h:1135-1136),
model = "dynamic" | "static" (dsp_static.h), header = the reference header the config is a variant of.
Derived sizes (dsp_dynamic.h:58-66, dsp_static.h:63) are computed by derive().
"""

DEFAULT_FUTURE = [0.05, 0.2, 0.5, 1.0, 1.5, 2.0]

CONFIGS = {
    # BASELINE.json configs[0..4]
    "cfg1": dict(header="dsp_static.h", model="static", nx=20, ny=20, nz=10, res=0.3, angle_res=3, max_ppv=8,
                 half_fov_h=42, half_fov_v=27, future_times=[0.05], neighbor_n=1, points=1000),
    "cfg2": dict(header="dsp_dynamic.h", model="dynamic", nx=66, ny=66, nz=40, res=0.15, angle_res=3, max_ppv=24,
                 half_fov_h=45, half_fov_v=30, future_times=DEFAULT_FUTURE, neighbor_n=1, points=10000),
    "cfg3": dict(header="dsp_dynamic_multiple_neighbors.h", model="dynamic", nx=66, ny=66, nz=40, res=0.15,
                 angle_res=1, max_ppv=24, half_fov_h=45, half_fov_v=30, future_times=DEFAULT_FUTURE, neighbor_n=2,
                 points=20000),
    "cfg4": dict(header="dsp_dynamic.h", model="dynamic", nx=66, ny=66, nz=40, res=0.15, angle_res=3, max_ppv=24,
                 half_fov_h=45, half_fov_v=30, future_times=[0.5, 1.0, 1.5, 2.0, 2.5, 3.0], neighbor_n=1,
                 points=10000),
    "cfg5": dict(header="dsp_dynamic.h", model="dynamic", nx=132, ny=132, nz=80, res=0.15, angle_res=3, max_ppv=36,
                 half_fov_h=45, half_fov_v=30, future_times=DEFAULT_FUTURE, neighbor_n=1, points=30000),
    # small maps for fast parity tests and capacity-overflow edge cases
    "tiny_dyn": dict(header="dsp_dynamic.h", model="dynamic", nx=16, ny=16, nz=10, res=0.25, angle_res=3, max_ppv=6,
                     half_fov_h=42, half_fov_v=24, future_times=DEFAULT_FUTURE, neighbor_n=1, points=400),
    "tiny_mn": dict(header="dsp_dynamic_multiple_neighbors.h", model="dynamic", nx=16, ny=16, nz=10, res=0.25,
                    angle_res=1, max_ppv=6, half_fov_h=42, half_fov_v=27, future_times=DEFAULT_FUTURE,
                    neighbor_n=2, points=400),
    "tiny_static": dict(header="dsp_static.h", model="static", nx=16, ny=16, nz=10, res=0.25, angle_res=3,
                        max_ppv=4, half_fov_h=42, half_fov_v=27, future_times=[0.05], neighbor_n=1, points=400),
    # the reference tree exactly as shipped (dsp_dynamic.h:38-50)
    "ref_default": dict(header="dsp_dynamic.h", model="dynamic", nx=66, ny=66, nz=40, res=0.15, angle_res=3,
                        max_ppv=9, half_fov_h=42, half_fov_v=24, future_times=DEFAULT_FUTURE, neighbor_n=1,
                        points=3000),
}


def derive(cfg):
    """Derived sizes, same integer arithmetic as dsp_dynamic.h:58-66 (static: dsp_static.h:63)."""
    d = dict(cfg)
    d["V"] = cfg["nx"] * cfg["ny"] * cfg["nz"]
    d["Nh"] = cfg["half_fov_h"] * 2 // cfg["angle_res"]
    d["Nv"] = cfg["half_fov_v"] * 2 // cfg["angle_res"]
    d["P"] = d["Nh"] * d["Nv"]
    pyramid_num = 360 * 180 // cfg["angle_res"] // cfg["angle_res"]
    safe_particle_num = int(d["V"] * cfg["max_ppv"] + 1e5)
    d["S"] = cfg["max_ppv"] * (5 if cfg["model"] == "static" else 2)
    d["L"] = safe_particle_num // pyramid_num * 2
    d["T"] = len(cfg["future_times"])
    d["NB"] = (2 * cfg["neighbor_n"] + 1) ** 2
    # occlusion margin: 0.3 m in dsp_dynamic.h:70,761; voxel_resolution in the mn / static headers (mn:761)
    d["occlusion_margin"] = 0.3 if cfg["header"] == "dsp_dynamic.h" else cfg["res"]
    return d
