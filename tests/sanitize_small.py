"""A few frames of the small configurations through every kernel family, for compute-sanitizer (GPU box):
    compute-sanitizer --tool memcheck  --error-exitcode 1 python tests/sanitize_small.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tests/sanitize_small.py
Covers the default frame, the device-side estimation front end (DSPMAP_EST_GPU=1), the exact overflow replay (tiny_mn: lists of
2 entries overflow every frame), the static model, the sparse and dense host readers, the pipelined reader and a 3-shard map."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dsp-map_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dspmap_b200 as dm
from common import gpu_map, gpu_update, make_stream


def run(name, frames, env=None, pinned=True):
    for k, v in (env or {}).items():
        os.environ[k] = v
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=4, frames=frames)
    g = gpu_map(name, seed=7, max_points=cfg["points"])
    for k in (env or {}):
        os.environ.pop(k)
    fut = np.zeros((g.V, g.T), np.float32)
    if pinned:
        g.pin_host_buffer(fut)
    born = 0
    for f in range(frames):
        assert gpu_update(g, st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]) == 1
        n, xyz, _ = g.getOccupancyMapWithFutureStatus(0.2, fut)
        born += g.counters()["n_born"]
    tk = g.get_occupancy_async(0.2, True)
    g.wait_occupancy(tk)
    g.close()
    print("%-12s %s frames %d: born %d, occupied %d" % (name, env or {}, frames, born, n), flush=True)


run("tiny_dyn", 5)
run("tiny_dyn", 5, {"DSPMAP_EST_GPU": "1"})
run("tiny_mn", 4)
run("tiny_static", 3, {"DSPMAP_EST_GPU": "1"}, pinned=False)
# three shards of one map through the library's orchestrator (collectives as device copies)
from common import SET
cfg = dm.CONFIGS["tiny_dyn"]
st = make_stream(cfg, seed=4, frames=4)
est = dm.VelocityEstimator(cfg, seed=7, filter_res=SET["filter_res"])


def setters(g):
    g.setPredictionVariance(SET["p_std"], SET["v_std"])
    g.setObservationStdDev(SET["ob_std"])
    g.setNewBornParticleNumberofEachPoint(SET["newborn_num"])
    g.setNewBornParticleWeight(SET["newborn_weight"])


lc = dm.LocalShardedMap(cfg, 3, seed=7, max_points=cfg["points"], setters=setters)
last = np.zeros((0, 7), np.float32)
for f in range(4):
    tg = est.estimate(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
    last = tg if tg is not None else last
    lc.update(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f], last)
    res = lc.occupancy(0.2)
print("3 shards: occupied %d" % res[0][0], flush=True)
for m in lc.maps:
    m.close()
print("ok")
