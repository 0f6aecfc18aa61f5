"""GPU counterpart of tests/test_oracle.py::test_device_math_fuzz: the CUDA path through the C ABI against the oracle on
seeded adversarial reads (random odd k in 3..63, homopolymer cut-offs 2..k, quality cut-offs with many values at the
threshold, runs of one base around the cut-off length, N runs, reads around k and across chunk boundaries, two colours).
Last in collection order on purpose: it is the newest GPU test of the round."""
import random

import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import mccortex_b200 as M
    assert M.device_count() > 0, "GPU tests need a CUDA device"
    return M


def _reads(rng, k, hp):
    reads = []
    for _ in range(rng.randint(20, 120)):
        n = rng.choice([rng.randint(1, k + 3), rng.randint(k, 3 * k + 5), rng.randint(100, 400), rng.randint(2000, 5000)])
        s = []
        while len(s) < n:
            r = rng.random()
            if r < 0.08:
                s += [rng.choice("ACGT")] * rng.choice([max(1, hp - 1), hp or 3, (hp or 3) + 1, rng.randint(1, 2 * k)])
            elif r < 0.11:
                s += ["N"] * rng.randint(1, 3)
            else:
                s += [rng.choice("ACGTacgt") for _ in range(rng.randint(1, 2 * k))]
        reads.append("".join(s[:n]))
    return reads


@pytest.mark.parametrize("block", range(6))
def test_gpu_fuzz_matches_oracle(M, oracle, block):
    rng = random.Random(900 + block)
    for case in range(6):
        k = rng.choice(range(3, 64, 2))
        hp = rng.choice([0, 0, 2, 3, rng.randint(2, k), k])
        cut = rng.choice([0, 0, rng.randint(36, 73)])
        ncols = rng.choice([1, 1, 2])
        og = oracle.Graph(k, ncols, 1 << 21)
        g = M.Graph(k, ncols, 1 << 21)
        ost = oracle.Stats()
        for col in range(ncols):
            reads = _reads(rng, k, hp)
            quals = None
            if cut:
                quals = ["".join(chr(rng.choice([cut - 1, cut, cut, cut + 1, rng.randint(35, 74)])) for _ in r) for r in reads]
                for i in range(0, len(reads), 5):
                    quals[i] = quals[i][:rng.randint(0, len(quals[i]))]
            for i, r in enumerate(reads):
                og.add_read(r, qual=(quals[i].encode("latin1") or None) if cut else None, colour=col, fq_cutoff=cut, hp_cutoff=hp, stats=ost)
            if cut:
                g.add_reads(reads, colour=col, hp_cutoff=hp, quals=quals, fq_cutoff=cut)
            elif rng.random() < 0.5:
                g.add_reads(reads, colour=col, hp_cutoff=hp)
            else:
                g.add_lines("".join(r + "\n" for r in reads).encode(), colour=col, hp_cutoff=hp)
        st = g.sync()
        got, n, _ = g.export_records()
        want = og.dump_sorted()[len(og.header()):]
        og.close()
        g.close()
        assert got == want, (block, case, k, hp, cut, ncols)
        assert n == ost.num_kmers_novel and st.num_kmers_loaded == ost.num_kmers_loaded, (block, case, k, hp, cut)
        assert st.contigs_parsed == ost.contigs_parsed and st.total_bases_loaded == ost.total_bases_loaded, (block, case, k, hp, cut)
