"""Host-side restatement of the dense non-interacting-set decomposition of the word-lattice Metropolis kernel
(brawl_b200/csrc/word_metropolis.cuh: BrwDenseGeom, brw_make_dense_step and the lane mapping).  CPU only: checks
the invariant the GPU kernel's exactness rests on -- all 2M sites touched in one step are distinct lattice sites
inside the active region and NO two of them are neighbours in shells 1-4 of the bcc Hamiltonian
(reference tables: src/bw_hamiltonian.f90:162-169, 225-230, 286-297, 361-384)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bcc_offsets(n_shells=4):
    src = open(os.path.join(ROOT, "brawl_b200", "csrc", "shell_tables.inc")).read()
    body = re.search(r"brw_bcc_off\[168\]\[3\] = \{(.*?)\};", src, re.S).group(1)
    offs = [tuple(int(v) for v in t.split(",")) for t in re.findall(r"\{(-?\d+,-?\d+,-?\d+)\}", body)]
    return offs[: (8, 14, 26, 50)[n_shells - 1]]


def dense_step_sites(q1, q2, rot1, rot2, si, swap, BX=64, BY=64, BZ=32, M=4, group=None):
    """group=None: one warp group over all NP planes (SPLIT=false).  group=0/1: the rows of warp group A (planes
    [0, NPA)) / B (planes [NPA+1, NP)) of the SPLIT kernels; rot1/rot2 are then rotations within the group."""
    NI = (BX - 2 * M) // 4
    HALF = ((BY - 3 * M) // 2) & ~3
    NJ, NP = HALF // 4, (BZ - 2 * M) // 4
    NROWS, YU = NJ * NP, M + HALF + M
    plane0 = 0
    if group is not None:
        NPA = (NP - 1) // 2
        NPB = NP - 1 - NPA
        NROWS = NJ * (NPA if group == 0 else NPB)
        plane0 = 0 if group == 0 else NPA + 1

    def cls(q):
        return [(q & 1) + 2 * ((q >> 1) & 1), (q & 1) + 2 * ((q >> 2) & 1), (q & 1) + 2 * ((q >> 3) & 1)]
    o = [cls(q1), cls(q2)]
    base = {}
    for h in range(2):
        ylo = M if h == 0 else YU
        for sub in range(2):
            base[(h, sub)] = (M + ((o[h][0] + 2 * sub - M) & 3), ylo + ((o[h][1] + 2 * sub - ylo) & 3),
                              M + ((o[h][2] + 2 * sub - M) & 3))
    s1, s2 = [], []
    for warp in range(min(32, NROWS)):
        for lane in range(2 * NI):
            sub = 1 if lane >= NI else 0
            li = lane - sub * NI
            r1, r2 = (warp + rot1) % NROWS, (warp + rot2) % NROWS
            X, Y, Z = base[(0, sub)]
            s1.append((X + 4 * li, Y + 4 * (r1 % NJ), Z + 4 * (r1 // NJ + plane0)))
            X, Y, Z = base[(1, sub ^ swap)]
            s2.append((X + 4 * ((li + si) % NI), Y + 4 * (r2 % NJ), Z + 4 * (r2 // NJ + plane0)))
    return s1, s2, (NI, NROWS)


def test_dense_step_sites_are_pairwise_non_interacting():
    offs = bcc_offsets(4)
    rng = np.random.default_rng(0)
    BX, BY, BZ, M = 64, 64, 32, 4
    cases = [(q1, q2, 0, 0, 0, 0) for q1 in range(16) for q2 in range(16)]
    cases += [tuple(int(v) for v in (rng.integers(16), rng.integers(16), rng.integers(36), rng.integers(36),
                                     rng.integers(14), rng.integers(2))) for _ in range(150)]
    for c in cases:
        s1, s2, (NI, NROWS) = dense_step_sites(*c)
        assert len(s1) == len(s2) == 896
        sites = s1 + s2
        S = set(sites)
        assert len(S) == len(sites)                                  # all distinct
        for (x, y, z) in sites:
            assert (x - z) % 2 == 0 and (y - z) % 2 == 0             # bcc sites
            assert M <= x < BX - M and M <= y < BY - M and M <= z < BZ - M   # every neighbour is inside the box
            for d in offs:
                assert (x + d[0], y + d[1], z + d[2]) not in S       # no two touched sites interact


def test_split_groups_never_touch_each_others_zone():
    """SPLIT kernels (two warp groups per CTA with independent barriers, word_metropolis.cuh): 64x64x32 box, planes
    {0,1} -> group A (12 warps), plane 2 = frozen gap, planes {3,4,5} -> group B (18 warps).  For ANY pair of steps
    (the groups are not synchronised) the sites one group touches, and every neighbour it reads, are disjoint from the
    sites the other group may write; within a group the dense-set invariant holds as before.  The layer pitch of 28
    with a box depth of 32 tiles a 256-plane lattice with shared frozen margins: 9 * 28 + 4 = 256."""
    offs = bcc_offsets(4)
    rng = np.random.default_rng(7)
    M, BZ = 4, 32
    zoneA = range(M, M + 8)                       # active z of group A: planes 0, 1
    zoneB = range(M + 12, BZ - M)                 # planes 3, 4, 5
    for _ in range(60):
        ca = tuple(int(v) for v in (rng.integers(16), rng.integers(16), rng.integers(12), rng.integers(12), rng.integers(14), rng.integers(2)))
        cb = tuple(int(v) for v in (rng.integers(16), rng.integers(16), rng.integers(18), rng.integers(18), rng.integers(14), rng.integers(2)))
        a1, a2, _ = dense_step_sites(*ca, group=0)
        b1, b2, _ = dense_step_sites(*cb, group=1)
        assert len(a1) == 12 * 28 and len(b1) == 18 * 28
        for sites, zone, other in ((a1 + a2, zoneA, zoneB), (b1 + b2, zoneB, zoneA)):
            S = set(sites)
            assert len(S) == len(sites)
            for (x, y, z) in sites:
                assert z in zone and M <= x < 64 - M and M <= y < 64 - M
                for d in offs:
                    assert (x + d[0], y + d[1], z + d[2]) not in S
                    assert z + d[2] not in other                      # reads stay out of the other group's zone
    assert 9 * 28 + (32 - 28) == 256 and (32 - 28) >= 3             # shared margin planes >= interaction reach


def test_dense_lanes_hit_distinct_banks():
    """Row pitch 32 words, plane pitch 1024: the 28 active lanes of a warp (14 A sites 2 words apart + 14 B sites, odd
    word offset) address 28 different shared-memory banks for site 1 and for site 2."""
    PXP, PLP = 32, 1024
    rng = np.random.default_rng(1)
    for _ in range(100):
        c = tuple(int(v) for v in (rng.integers(16), rng.integers(16), rng.integers(36), rng.integers(36), rng.integers(14),
                                   rng.integers(2)))
        s1, s2, _ = dense_step_sites(*c)
        for sites in (s1, s2):
            for w in range(32):
                banks = [((z * PLP + (y >> 1) * PXP + (x >> 1)) & 31) for (x, y, z) in sites[28 * w: 28 * w + 28]]
                assert len(set(banks)) == 28


def test_pair_word_gather_plan_counts_every_neighbour_once():
    """tools/gen_pair_gather.py (the generator of brawl_b200/csrc/pair_gather.inc): emulate the pair-word lattice
    W[c] = nibbles(c) | nibbles(c+1) << 16, the 30 loads, the nibble accumulators and the nibble->byte expansion on a
    random configuration and compare with a direct count of every shell's neighbours per species."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_pair_gather", os.path.join(ROOT, "tools", "gen_pair_gather.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    rng = np.random.default_rng(5)
    PX, PY, PZ = 32, 32, 12
    spec_c = rng.integers(0, 5, size=(PZ, PY, PX))                       # species 0..4 on the compact box
    nib = np.where(spec_c < 4, 1 << (4 * np.minimum(spec_c, 3)), 0).astype(np.uint64)
    W = nib.copy()
    W[:, :, :-1] |= nib[:, :, 1:] << np.uint64(16)

    def nib2byte(x):
        x = int(x) & 0xFFFF
        return sum(((x >> (4 * k)) & 15) << (8 * k) for k in range(4))
    offs = gen.bcc_offsets(4)
    for par in (0, 1):
        names, groups = gen.build(4, par)
        assert sum(len(o) for _, _, o in names) == 30
        for _ in range(40):
            zc, yc, xc = (int(v) for v in (rng.integers(4, PZ - 4), rng.integers(4, PY - 4), rng.integers(4, PX - 4)))
            # accumulators whose high lane is not a neighbour load the low 16 bits only (LDS.U16)
            acc = {nm: sum(int(W[zc + dz, yc + dyc, xc + dxc]) & (0xFFFF if role[1] is None else 0xFFFFFFFF)
                           for dz, dyc, dxc in o) & 0xFFFFFFFF for nm, role, o in names}
            C = [sum(nib2byte(sum((acc[nm] >> (16 * lane)) for nm, lane in g)) for g in gs) for gs in groups]
            want = [[0] * 4 for _ in range(4)]
            for (dx, dy, dz), n in offs:
                s = spec_c[zc + dz, yc + (par + dy) // 2, xc + (par + dx) // 2]
                if s < 4:
                    want[n][s] += 1
            got = [[(C[n] >> (8 * s)) & 255 for s in range(4)] for n in range(4)]
            assert got == want, (par, got, want)
