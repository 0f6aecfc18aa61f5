"""Occupancy model of the L2 front table (no GPU needed): which share of the HOT k-mers ends up in a way of its home
set (served by the hot pass), displaced into the neighbouring set (served by the parked pass at L2 speed), or not at
all (big table), for a given geometry.  First come, never evicted, like the kernels (mcx_front_add_slow /
mcx_front2_add_slow, mccortex_b200/csrc/mcx_table.cuh).  Keys arrive in random order; `junk` cold keys per hot key
(error k-mers that are seen once) arrive interleaved at the rate they have in the first pass over the genome.

usage: python scripts/front_sim.py [hot_keys] [junk_per_hot]
"""
import sys
import numpy as np

H = int(sys.argv[1]) if len(sys.argv) > 1 else 4_600_000
JUNK = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02   # cold keys per hot key while the hot set is being claimed
rng = np.random.default_rng(1)


def simulate(set_bits, ways, displace=True):
    nsets = 1 << set_bits
    n = int(H * (1 + JUNK))
    hot = np.zeros(n, dtype=bool); hot[rng.permutation(n)[:H]] = True
    home = rng.integers(0, nsets, size=n)
    fill = np.zeros(nsets, dtype=np.int32)
    where = np.zeros(n, dtype=np.int8)   # 1 home, 2 displaced, 0 none
    # sequential claim in arrival order (vectorised in rounds: keys of one round go to distinct sets)
    order = np.arange(n)
    pending = order
    while len(pending):
        s = home[pending]
        _, first = np.unique(s, return_index=True)          # one key per set per round keeps arrival order within a set
        first.sort()
        take = pending[first]
        st = home[take]
        ok = fill[st] < ways
        fill[st[ok]] += 1
        where[take[ok]] = 1
        rest = take[~ok]
        if displace and len(rest):
            # neighbouring set (set ^ 1); several keys of this round may aim at the same neighbour: serialise those
            nb = home[rest] ^ 1
            o = np.argsort(nb, kind="stable")
            nb_s, rest_s = nb[o], rest[o]
            rank = np.arange(len(nb_s)) - np.searchsorted(nb_s, nb_s, side="left")
            ok2 = fill[nb_s] + rank < ways
            np.add.at(fill, nb_s[ok2], 1)
            where[rest_s[ok2]] = 2
        mask = np.ones(len(pending), dtype=bool); mask[first] = False
        pending = pending[mask]
    h = where[hot]
    return (h == 1).mean(), (h == 2).mean(), (h == 0).mean(), fill.sum() / (nsets * ways)


print("hot keys %d, cold keys per hot key during warm-up %.3f" % (H, JUNK))
print("%-44s %8s %10s %8s %8s" % ("geometry", "home", "displaced", "none", "fill"))
for label, sb, ways in (("k<=31 shipped: 2^21 sets x 4 ways x 8 B   (64 MB)", 21, 4),
                        ("k>31  first:   2^21 sets x 2 ways x 16 B  (64 MB)", 21, 2),
                        ("k>31  shipped: 2^22 sets x 2 ways x 16 B (128 MB)", 22, 2),
                        ("k>31  next?:   2^19 sets x 8 ways x 16 B  (64 MB)", 19, 8),
                        ("k>31  next?:   2^20 sets x 8 ways x 16 B (128 MB)", 20, 8),
                        ("k>31  next?:   2^20 sets x 4 ways x 16 B  (64 MB)", 20, 4)):
    a, b, c, f = simulate(sb, ways)
    print("%-44s %7.1f%% %9.1f%% %7.1f%% %7.1f%%" % (label, 100 * a, 100 * b, 100 * c, 100 * f))
