import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (test infrastructure)."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def emul():
    """CPU executable that runs the MCX_HD device math (tests/emul)."""
    src = os.path.join(ROOT, "tests", "emul", "emul_frontend.cpp")
    exe = os.path.join(ROOT, "tests", "emul", "emul_frontend")
    deps = [src] + [os.path.join(ROOT, "mccortex_b200", "csrc", f) for f in ("mcx_device.cuh", "mcx_chunk.cuh")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, src])
    return exe


COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def rand_reads(rng, n, L, G, perr=0.01, pN=0.002, lower=0.1):
    """n reads of length L (int or (lo,hi)) from a random genome of G bases, both strands,
    substitutions, a few non-ACGT bytes and some lower-case reads."""
    genome = "".join(rng.choice("ACGT") for _ in range(G))
    reads = []
    for _ in range(n):
        ln = L if isinstance(L, int) else rng.randint(*L)
        s = rng.randrange(0, G - ln + 1)
        r = list(genome[s:s + ln])
        if rng.random() < 0.5:
            r = [COMP[c] for c in reversed(r)]
        for j in range(ln):
            x = rng.random()
            if x < perr:
                r[j] = rng.choice("ACGT")
            elif x < perr + pN:
                r[j] = rng.choice("NnRY-.")
        r = "".join(r)
        if rng.random() < lower:
            r = r.lower()
        reads.append(r)
    return reads


EDGE_READS = ["", "A", "ACGT" * 50, "T" * 200, "acgtn" * 40, "N" * 100, "AC" * 100, "G" * 31, "C" * 63]


@pytest.fixture(scope="session")
def reads_small():
    rng = random.Random(20261017)
    return rand_reads(rng, 400, (1, 400), 6000, perr=0.01, pN=0.01) + EDGE_READS


def oracle_records(O, reads, k, ncols=1, colour=0, hp_cutoff=0, capacity=1 << 22):
    """(records bytes, Stats) from the oracle for a list of reads in one colour."""
    g = O.Graph(k, ncols, capacity)
    st = O.Stats()
    for r in reads:
        g.add_read(r, colour=colour, hp_cutoff=hp_cutoff, stats=st)
    full = g.dump_sorted()
    hdr = len(g.header())
    g.close()
    return full[hdr:], st
