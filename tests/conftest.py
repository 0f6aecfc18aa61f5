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
    deps = [src] + [os.path.join(ROOT, "mccortex_b200", "csrc", f) for f in ("mcx_device.cuh", "mcx_chunk.cuh", "mcx_pcr.cuh", "mcx_lane.cuh")]
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


def pcr_reads(rng, n, k, paired=0.5, G=4000, sites=60, lens=(5, 160)):
    """reads for --remove-pcr tests: few start sites (so many reads share a start k-mer), both strands, some
    mutated / lower-case / too short.  Returns a list of units: (seq,) or (seq1, seq2)."""
    genome = "".join(rng.choice("ACGT") for _ in range(G))
    tr = str.maketrans("ACGTacgt", "TGCAtgca")

    def one():
        st = rng.randrange(sites) * (G // (sites + 4))
        r = genome[st:st + rng.randint(*lens)]
        if rng.random() < 0.5:
            r = r[::-1].translate(tr)
        r = "".join((rng.choice("ACGTN") if rng.random() < 0.01 else c) for c in r)
        return r.lower() if rng.random() < 0.1 else r
    return [(one(), one()) if rng.random() < paired else (one(),) for _ in range(n)]


@pytest.fixture(scope="session")
def ingest_dump():
    """the host driver's --remove-pcr ingest linked against stubs of the library (tests/emul/ingest_dump.c)"""
    src = os.path.join(ROOT, "tests", "emul", "ingest_dump.c")
    exe = os.path.join(ROOT, "tests", "emul", "ingest_dump")
    host = os.path.join(ROOT, "mccortex_b200", "host")
    deps = [src, os.path.join(host, "seq_ingest.c"), os.path.join(host, "seq_ingest_par.c"), os.path.join(host, "util.c"),
            os.path.join(host, "mcx_host.h"), os.path.join(ROOT, "include", "mcx_gpu.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-D_GNU_SOURCE", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                               os.path.join(host, "seq_ingest.c"), os.path.join(host, "seq_ingest_par.c"), os.path.join(host, "util.c"),
                               "-lz", "-lm", "-lpthread"])
    return exe


def build_hostcheck():
    """the host driver's C sources linked against tests/emul/abi_shim.c (the C ABI implemented by the oracle) instead
    of libmcxgpu.so: test infrastructure for the CPU-only run, never part of the product"""
    import glob
    if os.environ.get("MCX_HOSTCHECK_EXE"):   # e.g. the same sources built with -fsanitize=address,undefined / thread
        return os.environ["MCX_HOSTCHECK_EXE"]
    exe = os.path.join(ROOT, "tests", "emul", "hostcheck")
    host = sorted(glob.glob(os.path.join(ROOT, "mccortex_b200", "host", "*.c")))
    srcs = host + [os.path.join(ROOT, "tests", "emul", "abi_shim.c"), os.path.join(ROOT, "oracle", "mcx_oracle.c")]
    deps = srcs + [os.path.join(ROOT, "mccortex_b200", "host", "mcx_host.h"), os.path.join(ROOT, "include", "mcx_gpu.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["gcc", "-O2", "-std=c99", "-D_GNU_SOURCE", "-w", "-I", os.path.join(ROOT, "include"), "-o", exe] + srcs +
                              ["-lz", "-lpthread", "-lm"])
    return exe


@pytest.fixture(scope="session")
def hostcheck():
    return build_hostcheck()


def write_pcr_files(rng, d, n=400, k=21):
    """se.fa, p1.fq / p2.fq (p2 one read longer), il.fq (pairs by name, /1 /2 names, singles) for --remove-pcr tests"""
    def q(L, lowp=0.05):
        return "".join(chr(33 + (rng.randint(0, 12) if rng.random() < lowp else rng.randint(15, 40))) for _ in range(L))
    units = pcr_reads(rng, 3 * n, k, paired=1.0, G=3000, sites=60)
    reads = [m for u in units for m in u]
    it = iter(reads)
    with open(os.path.join(d, "se.fa"), "w") as f:
        for i in range(n):
            f.write(">r%d\n%s\n" % (i, next(it)))
    with open(os.path.join(d, "p1.fq"), "w") as f1, open(os.path.join(d, "p2.fq"), "w") as f2:
        for i in range(n):
            a, b = next(it), next(it)
            f1.write("@p%d/1\n%s\n+\n%s\n" % (i, a, q(len(a))))
            f2.write("@p%d/2\n%s\n+\n%s\n" % (i, b, q(len(b))))
        f2.write("@extra/2\nACGTACGTACGTACGTACGTACGTACGTAAAC\n+\n%s\n" % q(32))
    with open(os.path.join(d, "il.fq"), "w") as f:
        for i in range(n):
            a = next(it)
            x = rng.random()
            if x < 0.5:
                b = next(it)
                f.write("@i%d/1 x\n%s\n+\n%s\n@i%d/2 y\n%s\n+\n%s\n" % (i, a, q(len(a)), i, b, q(len(b))))
            elif x < 0.7:
                b = next(it)
                f.write("@i%d extra\n%s\n+\n%s\n@i%d\n%s\n+\n%s\n" % (i, a, q(len(a)), i, b, q(len(b))))
            else:
                f.write("@s%d\n%s\n+\n%s\n" % (i, a, q(len(a))))
