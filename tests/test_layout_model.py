"""CPU model of the column-major pair-buffer layout (DSPMAP_G_COL): the addresses k_pair_eval_col
writes are exactly the addresses k_cz_chain_col and k_weight_col read, every bulk copy is 16-byte aligned, and a
pyramid's block is filled without gaps or overlaps.  It restates the kernels' index arithmetic (dspmap_frame.cuh); the
kernels themselves are checked on the GPU by tests/ab_toggles.py."""
import numpy as np


def test_column_major_pair_buffer_addressing_is_consistent():
    rng = np.random.default_rng(0)
    Nh, Nv, OBS = 6, 5, 100
    P = Nh * Nv
    plen = rng.integers(0, 90, P)
    plen[rng.random(P) < 0.2] = 0
    obs = rng.integers(0, 40, P)
    obs[rng.random(P) < 0.3] = 0
    nbr = []
    for p in range(P):
        h, v = divmod(p, Nv)
        nbr.append([(h + i) * Nv + v + j for i in (-1, 0, 1) for j in (-1, 0, 1) if 0 <= h + i < Nh and 0 <= v + j < Nv])
    # k_pair_prep(col = 1)
    cum, totlen, pairs = {}, np.zeros(P, int), np.zeros(P, int)
    for i in range(P):
        c = 0
        for ns, b in enumerate(nbr[i]):
            cum[(i, ns)] = c
            c += plen[b]
        totlen[i] = c
        npi = min(obs[i], OBS - 1)
        pairs[i] = (npi + 1) * ((c + 3) & ~3) if npi > 0 else 0
    rowbase = np.concatenate([[0], np.cumsum(pairs)])
    assert (rowbase % 4 == 0).all()
    chunks = (plen + 31) >> 5
    G = {}
    # k_pair_eval_col: item = (chunk of pyramid a, neighbour slot ns), lane = particle, one store per point
    for a in range(P):
        for ch in range(chunks[a]):
            k0 = ch << 5
            nrows = min(32, plen[a] - k0)
            for i in nbr[a]:
                npi = min(obs[i], OBS - 1)
                if npi == 0:
                    continue
                tl = (totlen[i] + 3) & ~3
                j = cum[(i, nbr[i].index(a))] + k0
                for lane in range(nrows):
                    gb = rowbase[i] + j + lane
                    for z in range(npi + 1):
                        assert gb + z * tl not in G
                        G[gb + z * tl] = (i, j + lane, z)
    # k_cz_chain_col: stages of RT rows of every column, aligned bulk copies
    for i in range(P):
        npi = min(obs[i], OBS - 1)
        if npi == 0:
            continue
        rows, ncol = totlen[i], npi + 1
        tl = (rows + 3) & ~3
        RT = min(512, (4096 // ncol - 4) & ~7)
        assert RT >= 8 and ((RT + 4) // 4) % 2 == 1 and ncol * (RT + 4) <= 4096
        for j0 in range(0, rows, RT):
            cur = min(RT, rows - j0)
            assert j0 + ((cur + 3) & ~3) <= tl
            for c in range(ncol):
                src = rowbase[i] + c * tl + j0
                assert src % 4 == 0
                for jj in range(cur):
                    assert G[src + jj] == (i, j0 + jj, c)
        assert rowbase[i] + ncol * tl == rowbase[i + 1]
    # k_weight_col: lane r of the chunk's warp reads point z of neighbour b at column z, row jb + r
    for a in range(P):
        for ch in range(chunks[a]):
            k0 = ch << 5
            nrows = min(32, plen[a] - k0)
            for b in nbr[a]:
                npb = min(obs[b], OBS - 1)
                if npb == 0:
                    continue
                jb = cum[(b, nbr[b].index(a))] + k0
                tl = (totlen[b] + 3) & ~3
                for f in range(32 * npb):
                    r, z = f & 31, f >> 5
                    if r < nrows:
                        assert G[rowbase[b] + jb + z * tl + r] == (b, jb + r, z)
