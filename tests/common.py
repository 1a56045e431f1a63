"""Shared fixtures of the test-suite."""
import os

import numpy as np

import dspmap_b200 as dm
from dspmap_b200.streams import make_stream  # noqa: F401

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = [("tiny_dyn", 6, 0), ("tiny_static", 5, 0), ("tiny_dyn", 4, 1500)]
SET = dict(p_std=0.05, v_std=0.05, ob_std=0.1, newborn_num=20, newborn_weight=1e-4, filter_res=0.1)  # ex:522-526


def load_golden(name, frames, initp):
    return np.load(os.path.join(GOLDEN_DIR, "%s_%d_%d.npz" % (name, frames, initp)))


def gpu_map(cfg_name, seed=7, init_particles=0, **kw):
    g = dm.DSPMap(dm.CONFIGS[cfg_name], seed=seed, init_particle_num=init_particles, **kw)
    g.setPredictionVariance(SET["p_std"], SET["v_std"])
    g.setObservationStdDev(SET["ob_std"])
    g.setNewBornParticleNumberofEachPoint(SET["newborn_num"])
    g.setNewBornParticleWeight(SET["newborn_weight"])
    g.setOriginalVoxelFilterResolution(SET["filter_res"])
    return g


def gpu_update(g, pts, pos, t, q, tagged=None, stride=3):
    n = np.asarray(pts).size // stride
    return g.update(n, stride, pts, float(pos[0]), float(pos[1]), float(pos[2]), float(t), float(q[0]), float(q[1]), float(q[2]),
                    float(q[3]), tagged=tagged)


def check_against_golden(m, G, f, is_gpu):
    """m: OracleMap or DSPMap after frame f. Bit-exact for everything except the (atomically accumulated) GPU future grid."""
    from parity import same
    ids, vals = m.particles()
    assert same(ids, G["ids_%d" % f]), "frame %d: particle (voxel, slot) ids" % f
    assert same(vals, G["vals_%d" % f]), "frame %d: particle fields" % f
    off, ent = m.pyramid_lists()
    assert same(off, G["pyr_off_%d" % f]) and same(ent, G["pyr_ent_%d" % f]), "frame %d: pyramid lists" % f
    cnt, mx, pts = m.observations()
    assert same(cnt, G["obs_cnt_%d" % f]) and same(mx, G["obs_max_%d" % f]), "frame %d: observation bins" % f
    msk = np.arange(pts.shape[1])[None, :] < cnt[:, None]
    assert same(pts[msk], G["obs_pts_%d" % f]), "frame %d: binned points / C_z" % f
    vo = m.voxel_objects()
    ref = np.zeros_like(vo)
    ref[G["vox_idx_%d" % f]] = G["vox_val_%d" % f]
    assert same(vo[:, :4], ref[:, :4]), "frame %d: occupancy / mean velocity" % f
    if is_gpu:
        assert np.array_equal(vo[:, 4:] != 0, ref[:, 4:] != 0) and np.allclose(vo[:, 4:], ref[:, 4:], rtol=2e-6, atol=0), "frame %d: future grid" % f
    else:
        assert same(vo[:, 4:], ref[:, 4:]), "frame %d: future grid" % f
    assert np.array_equal(m.cursors()[:3], G["cursors_%d" % f]), "frame %d: cursors" % f
