#!/usr/bin/env python3
"""Generate brawl_b200/csrc/pair_gather.inc: the neighbour-count gathers of the pair-word Metropolis kernel
(word_metropolis.cuh, PAIRW = true).

Shared-memory word of compact site c:  W[c] = n(c) | n(c+1) << 16,  n = one-hot nibble of the species
(1 << 4*species for species 0..3, 0 for species 4).  One LDS.32 therefore returns TWO x-adjacent sites for any c (loads whose high lane is not a neighbour are LDS.U16),
and the 50 neighbours of a bcc site (shells 1-4, reference tables src/bw_hamiltonian.f90:162-169, 225-230, 286-297,
361-384 via shell_tables.inc) are covered by 30 loads.  Each load is added to an accumulator chosen by the pair of
shells its two 16-bit lanes belong to ("X" = a lane that is not a neighbour; its sums are never used).  An
accumulator takes at most 15 loads, so no 4-bit field overflows.  At the end the lanes of one shell are added in
groups of at most 15 counts, expanded from nibbles to bytes and summed:  C[shell] = 4 byte fields = neighbours of
species 0..3 in that shell (species 4 is inferred, as in the 32-bit-word kernel).

Usage: python tools/gen_pair_gather.py   (writes the .inc; deterministic)"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bcc_offsets(n_shells):
    src = open(os.path.join(ROOT, "brawl_b200", "csrc", "shell_tables.inc")).read()
    body = re.search(r"brw_bcc_off\[168\]\[3\] = \{(.*?)\};", src, re.S).group(1)
    offs = [tuple(int(v) for v in t.split(",")) for t in re.findall(r"\{(-?\d+,-?\d+,-?\d+)\}", body)]
    counts = [8, 6, 12, 24, 8, 6, 24, 24, 24, 32]
    out, k = [], 0
    for n in range(n_shells):
        for _ in range(counts[n]):
            out.append((offs[k], n))
            k += 1
    return out


def fdiv2(v):
    return v // 2          # Python floor division == floor for negatives


def plan(n_shells, par):
    """-> loads [(dz, dyc, dxc, role_lo, role_hi)], role = shell index or None"""
    rows = {}
    for (dx, dy, dz), n in bcc_offsets(n_shells):
        dxc, dyc = fdiv2(par + dx), fdiv2(par + dy)
        rows.setdefault((dz, dyc), {})
        assert dxc not in rows[(dz, dyc)], "two neighbours on one compact site"
        rows[(dz, dyc)][dxc] = n
    loads = []
    for (dz, dyc) in sorted(rows):
        d = dict(rows[(dz, dyc)])
        for x in sorted(d):
            if x not in d:
                continue
            lo = d.pop(x)
            hi = d.pop(x + 1, None)
            if (dz, dyc) == (0, 0) and x + 1 == 0:
                hi = None                                  # the hi lane is the site itself
            loads.append((dz, dyc, x, lo, hi))
    return loads


def build(n_shells, par):
    """-> (accumulators [(name, (role_lo, role_hi), [(dz, dyc, dxc)])], groups[shell] = [[(name, lane)]])"""
    loads = plan(n_shells, par)
    accs = {}
    for dz, dyc, dxc, lo, hi in loads:
        accs.setdefault((lo, hi), []).append((dz, dyc, dxc))
    final = []                                             # an accumulator takes at most 15 loads
    for role, offs in sorted(accs.items(), key=lambda t: (t[0][0], -1 if t[0][1] is None else t[0][1])):
        for i in range(0, len(offs), 15):
            final.append((role, offs[i:i + 15]))
    names = [("a%d_%s%s" % (idx, role[0] + 1, "x" if role[1] is None else role[1] + 1), role, offs)
             for idx, (role, offs) in enumerate(final)]
    groups = []
    for n in range(n_shells):
        terms = []
        for nm, role, offs in names:
            if role[0] == n:
                terms.append((len(offs), nm, 0))
            if role[1] == n:
                terms.append((len(offs), nm, 1))
        terms.sort(key=lambda t: -t[0])
        gs = []                                            # first-fit decreasing into groups of <= 15 counts
        for cnt, nm, lane in terms:
            for g in gs:
                if g[0] + cnt <= 15:
                    g[0] += cnt
                    g[1].append((nm, lane))
                    break
            else:
                gs.append([cnt, [(nm, lane)]])
        assert sum(g[0] for g in gs) == [8, 6, 12, 24, 8, 6][n], (n, gs)
        groups.append([g[1] for g in gs])
    return names, groups


def emit(n_shells, par):
    names, groups = build(n_shells, par)
    lines = []
    for nm, role, offs in names:
        # hi lane unused ("x"): 16-bit load of the low half only -- the high half may be another thread's trial site of
        # the same step (offset = (2,2,2) mod 4), which that thread is free to rewrite
        fmt = "brw_lo16(wc + (%d) * PLP + (%d) * PXP + (%d))" if role[1] is None else "wc[(%d) * PLP + (%d) * PXP + (%d)]"
        terms = [fmt % o for o in offs]
        lines.append("    const uint32_t %s = %s;" % (nm, " + ".join(terms)))
    for n, gs in enumerate(groups):
        ex = []
        for g in gs:
            ex.append("brw_nib2byte(%s)" % " + ".join(nm if lane == 0 else "(%s >> 16)" % nm for nm, lane in g))
        lines.append("    C[%d] = %s;" % (n, " + ".join(ex)))
    return sum(len(o) for _, _, o in names), lines


def main():
    out = ["/* GENERATED by tools/gen_pair_gather.py -- neighbour-count gathers of the pair-word kernel (bcc).",
           " * wc points at the word of the centre site; W[c] = nibbles(site c) | nibbles(site c+1) << 16.  Data-derived",
           " * from shell_tables.inc (reference neighbour tables, src/bw_hamiltonian.f90:134-877). */",
           "template <int NSH, int PXP, int PLP, int PAR> struct BrwPairGatherBcc;"]
    for nsh in (4,):
        for par in (0, 1):
            n, lines = emit(nsh, par)
            out.append("template <int PXP, int PLP> struct BrwPairGatherBcc<%d, PXP, PLP, %d> {   // %d loads" % (nsh, par, n))
            out.append("  static constexpr int n_loads = %d;" % n)
            out.append("  static __device__ __forceinline__ void run(const uint32_t *wc, uint32_t (&C)[%d]) {" % nsh)
            out += lines
            out.append("  }")
            out.append("};")
    path = os.path.join(ROOT, "brawl_b200", "csrc", "pair_gather.inc")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
