"""GPU parity tests of the drop-in command line: `mccortex-b200 build -S ...` must write the
same bytes as the compiled reference did for the committed golden cases (tests/golden/), and
as oracle/_ref does when it is present, on fresh seeded inputs."""
import hashlib
import json
import os
import random
import subprocess

import pytest

from conftest import ROOT, rand_reads, EDGE_READS

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
CASES = json.load(open(os.path.join(GOLD, "cases.json")))


def _driver():
    import mccortex_b200 as M
    assert M.device_count() > 0
    assert os.path.exists(M.driver_path())
    return M.driver_path()


def _run(args, cwd=None, check=True, env=None):
    r = subprocess.run([_driver(), "build"] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       env=dict(os.environ, **env) if env else None)
    if check:
        assert r.returncode == 0, r.stderr.decode()[-3000:]
    return r


@pytest.mark.parametrize("case", [c["name"] for c in CASES])
def test_cli_matches_golden_ctx(case, tmp_path):
    c = next(x for x in CASES if x["name"] == case)
    out = str(tmp_path / "out.ctx")
    args = ["-q", "-f", "-m", "1G", "-n", "1M", "-k", str(c["k"]), "-S"]
    args += [a if not a.endswith((".fa", ".fq")) else os.path.join(GOLD, a) for a in c["ref_args"]] + [out]
    _run(args)
    got = open(out, "rb").read()
    ref = open(os.path.join(GOLD, c["ctx"]), "rb").read()
    assert hashlib.md5(ref).hexdigest() == c["md5"]
    assert got == ref


GRAPH_CASES = json.load(open(os.path.join(GOLD, "graph_cases.json")))


@pytest.mark.parametrize("case", [c["name"] for c in GRAPH_CASES])
def test_cli_graph_loading_matches_golden_ctx(case, tmp_path):
    """build --graph [into:]in.ctx[:from]: graph files are merged on the GPU
    (mcx_graph_load_records), header metadata on the host; bytes == the compiled reference's"""
    c = next(x for x in GRAPH_CASES if x["name"] == case)
    out = str(tmp_path / "out.ctx")
    args = ["-q", "-f", "-m", "1G", "-n", "1M", "-k", str(c["k"]), "-S"]
    args += [a.replace("@/", GOLD + "/") for a in c["ref_args"]] + [out]
    _run(args)
    got = open(out, "rb").read()
    ref = open(os.path.join(GOLD, c["ctx"]), "rb").read()
    assert hashlib.md5(ref).hexdigest() == c["md5"]
    assert got == ref


@pytest.mark.parametrize("case", [c["name"] for c in GRAPH_CASES if c["name"].startswith("pcr_")])
def test_cli_remove_pcr_small_batches(case, tmp_path):
    """build --remove-pcr with the host cutting the input into many tiny batches: the read-start
    marks carry over from batch to batch, pairs stay whole, a read waiting for its mate moves to the next batch"""
    c = next(x for x in GRAPH_CASES if x["name"] == case)
    out = str(tmp_path / "out.ctx")
    args = ["-q", "-f", "-m", "1G", "-n", "1M", "-k", str(c["k"]), "-S"]
    args += [a.replace("@/", GOLD + "/") for a in c["ref_args"]] + [out]
    _run(args, env={"MCX_BATCH_BYTES": "3000"})
    assert open(out, "rb").read() == open(os.path.join(GOLD, c["ctx"]), "rb").read()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mccortex31")), reason="oracle/_ref not built")
def test_cli_remove_pcr_matches_reference_binary(tmp_path, oracle):
    """fresh seeded inputs, the compiled reference run with one worker thread (its deterministic mode)"""
    from conftest import write_pcr_files
    write_pcr_files(random.Random(2024), str(tmp_path), n=600, k=27)
    d = str(tmp_path)
    # (one input task per colour: with two tasks in a colour the reference interleaves their reads as they arrive)
    for k, args in ((27, ["-p", "-s", "a", "-2", d + "/p1.fq:" + d + "/p2.fq", "-s", "b", "-1", d + "/se.fa"]),
                    (41, ["-p", "-H", "6", "-M", "RF", "-s", "a", "-i", d + "/il.fq", "-s", "b", "-Q", "15", "-2", d + "/p1.fq:" + d + "/p2.fq"])):
        ref = oracle.ref_build(k, args, str(tmp_path / "ref.ctx"), threads=1)
        out = str(tmp_path / "out.ctx")
        _run(["-q", "-f", "-m", "1G", "-n", "1M", "-k", str(k), "-S"] + args + [out])
        assert open(out, "rb").read() == ref, args


def test_cli_many_tasks_header_quirks(tmp_path, oracle):
    """13 tasks in 2 colours: batches of 10 tasks, stats credited to the first task's colour
    (quirk Q1) and the lossy mean_read_length round trips (Q4).  Checked against the oracle,
    which is pinned to the reference on exactly this shape (tests/test_oracle.py)."""
    rng = random.Random(13)
    files = []
    for i in range(13):
        p = tmp_path / ("t%d.fa" % i)
        p.write_text("".join(">r\n%s\n" % r for r in rand_reads(rng, 20, (10, 120), 3000)))
        files.append(str(p))
    args = ["-q", "-f", "-m", "1G", "-n", "1M", "-k", "15", "-S", "-s", "c0"]
    for f in files[:12]:
        args += ["-1", f]
    args += ["-s", "c1", "-1", files[12], str(tmp_path / "out.ctx")]
    _run(args)
    want, _ = oracle.build_ctx(15, [("c0", files[:12]), ("c1", files[12:])])
    assert open(tmp_path / "out.ctx", "rb").read() == want


def test_cli_gz_plain_fastq_inputs(tmp_path, oracle):
    import gzip
    rng = random.Random(21)
    reads = rand_reads(rng, 300, (20, 250), 5000) + EDGE_READS
    fa_gz = tmp_path / "a.fa.gz"
    with gzip.open(fa_gz, "wt") as f:
        for i, r in enumerate(reads):
            f.write(">r%d some description\n" % i)
            for j in range(0, len(r), 70):
                f.write(r[j:j + 70] + "\r\n")
    plain = tmp_path / "b.txt"
    plain.write_text("\n".join(r for r in reads if r) + "\n")
    fq = tmp_path / "c.fq"
    fq.write_text("".join("@r%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)) for i, r in enumerate(reads) if r))
    out = tmp_path / "out.ctx"
    _run(["-q", "-f", "-m", "1G", "-n", "1M", "-k", "31", "-S", "-s", "gz", "-1", str(fa_gz), "-s", "plain", "-1",
          str(plain), "-s", "fq", "-1", str(fq), str(out)])
    want, _ = oracle.build_ctx(31, [("gz", [str(fa_gz)]), ("plain", [str(plain)]), ("fq", [str(fq)])])
    assert open(out, "rb").read() == want


def test_cli_refuses_to_overwrite_and_bad_args(tmp_path):
    fa = os.path.join(GOLD, "a.fa")
    out = tmp_path / "o.ctx"
    base = ["-q", "-m", "1G", "-n", "1M", "-k", "31", "-s", "x", "-1", fa, str(out)]
    _run(base)
    r = _run(base, check=False)
    assert r.returncode == 1 and b"File already exists" in r.stderr   # file_util.c:164-174
    _run(["-f"] + base)
    assert _run(["-q", "-k", "31", "-1", fa, str(out)], check=False).returncode == 1        # sample first
    assert _run(["-q", "-k", "30", "-s", "x", "-1", fa, str(out)], check=False).returncode == 1  # even k
    assert _run(["-q", "-f", "-k", "31", "-s", "x", "-1", fa, "-H", "3", str(out)], check=False).returncode == 1  # trailing pref
    r = _run(["-q", "-f", "-m", "1G", "-n", "1024", "-k", "31", "-s", "x", "-1", fa, str(out)], check=False)
    assert r.returncode == 1 and b"Hash table is full" in r.stderr    # hash_table.c:119-123


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mccortex31")), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("k", [31, 63])
def test_cli_matches_reference_binary_and_check(tmp_path, oracle, k):
    """fresh input: bytes equal to the reference binary's `build -S`; the reference's own
    `check` (edge reciprocity, db_graph_healthcheck) accepts our file"""
    rng = random.Random(1000 + k)
    reads = rand_reads(rng, 2000, 150, 30000, perr=0.005)
    fa = tmp_path / "r.fa"
    fa.write_text("".join(">r\n%s\n" % r for r in reads))
    mine = tmp_path / "mine.ctx"
    _run(["-q", "-f", "-m", "1G", "-n", "4M", "-k", str(k), "-S", "-s", "s", "-1", str(fa), str(mine)])
    ref = oracle.ref_build(k, ["-s", "s", "-1", str(fa)], str(tmp_path / "ref.ctx"), threads=4, nkmers="4M")
    assert open(mine, "rb").read() == ref
    r = oracle.ref_run(k, ["check", "-q", str(mine)], check=False)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    # unsorted output + reference `sort` == sorted output (tests/sort/Makefile:25-45)
    uns = tmp_path / "uns.ctx"
    _run(["-q", "-f", "-m", "1G", "-n", "4M", "-k", str(k), "-s", "s", "-1", str(fa), str(uns)])
    oracle.ref_run(k, ["sort", "-q", str(uns)])
    assert open(uns, "rb").read() == ref


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mccortex31")), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("k", [27, 41])
def test_cli_graph_loading_matches_reference_binary(tmp_path, oracle, k):
    """build --graph on fresh inputs: our own unsorted 2-colour output and a reference-built graph go
    back in through colour filters; bytes equal to the reference binary given the same arguments,
    and the reference's `check` accepts the result"""
    rng = random.Random(2000 + k)
    fas = []
    for i in range(3):
        p = tmp_path / ("r%d.fa" % i)
        p.write_text("".join(">r%d\n%s\n" % (j, r) for j, r in enumerate(rand_reads(rng, 400, (20, 220), 9000, perr=0.004))))
        fas.append(str(p))
    g2 = str(tmp_path / "g2.ctx")   # written by us, unsorted, 2 colours
    _run(["-q", "-f", "-m", "1G", "-n", "2M", "-k", str(k), "-s", "a", "-1", fas[0], "-s", "b", "-1", fas[1], g2])
    g1 = str(tmp_path / "g1.ctx")   # written by the reference
    oracle.ref_build(k, ["-s", "c", "-1", fas[2]], g1, threads=2, nkmers="2M")
    for n, args in enumerate((["-g", g2, "-g", g1, "-s", "n", "-1", fas[0]],
                              ["-g", "1:" + g2 + ":0", "-s", "n", "-1", fas[1], "-1", fas[2]],
                              ["-g", g2 + ":1-0", "-g", "0,0:" + g2, "-s", "n", "-1", fas[2]])):
        mine = str(tmp_path / ("mine%d.ctx" % n))
        _run(["-q", "-f", "-m", "1G", "-n", "2M", "-k", str(k), "-S"] + args + [mine])
        ref = oracle.ref_build(k, args, str(tmp_path / "ref.ctx"), threads=3, nkmers="2M")
        assert open(mine, "rb").read() == ref, args
        r = oracle.ref_run(k, ["check", "-q", mine], check=False)
        assert r.returncode == 0, r.stderr.decode()[-2000:]


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mccortex31")), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("k", [25, 45])
def test_cli_intersect_matches_reference_binary(tmp_path, oracle, k):
    """build --intersect on fresh inputs (reads that only partly overlap the intersection graph, so that
    contigs contain runs of k-mers that are not found): bytes equal to the reference binary's"""
    rng = random.Random(3000 + k)
    fas = []
    for i in range(3):
        p = tmp_path / ("r%d.fa" % i)
        p.write_text("".join(">r%d\n%s\n" % (j, r) for j, r in enumerate(rand_reads(rng, 500, (20, 260), 6000, perr=0.01))))
        fas.append(str(p))
    # the three files sample the same seeded genome? no: rand_reads draws a new genome per call -- share one
    genome_reads = rand_reads(random.Random(5), 1500, (20, 260), 6000, perr=0.01)
    for i in range(3):
        with open(fas[i], "w") as f:
            f.write("".join(">r%d\n%s\n" % (j, r) for j, r in enumerate(genome_reads[i * 500:(i + 1) * 500])))
    isec = str(tmp_path / "isec.ctx")
    oracle.ref_build(k, ["-s", "i", "-1", fas[0]], isec, threads=2, nkmers="2M")
    other = str(tmp_path / "other.ctx")
    oracle.ref_build(k, ["-s", "o", "-1", fas[1], "-s", "p", "-1", fas[2]], other, threads=2, nkmers="2M")
    for n, args in enumerate((["-I", isec, "-s", "n", "-1", fas[1], "-1", fas[2]],
                              ["-I", isec, "-g", other, "-s", "n", "-H", "6", "-1", fas[2]],
                              ["-I", other + ":1", "-I", isec, "-g", other + ":0", "-s", "a", "-1", fas[0], "-s", "b", "-1", fas[1]])):
        mine = str(tmp_path / ("mine%d.ctx" % n))
        _run(["-q", "-f", "-m", "1G", "-n", "2M", "-k", str(k), "-S"] + args + [mine])
        ref = oracle.ref_build(k, args, str(tmp_path / "ref.ctx"), threads=3, nkmers="2M")
        assert open(mine, "rb").read() == ref, args


def _run_cmd(cmd, args, check=True):
    r = subprocess.run([_driver(), cmd] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    if check:
        assert r.returncode == 0, r.stderr.decode()[-3000:]
    return r


@pytest.mark.parametrize("k", [31, 63])
def test_cli_sort_command(tmp_path, oracle, k):
    """`mccortex-b200 sort` (src/commands/ctx_sort.c): in place and with -o; sort(unsorted build) == build -S
    (the reference asserts the same, tests/sort/Makefile:25-45), and == the reference's own `sort`"""
    import shutil
    rng = random.Random(4000 + k)
    fa = tmp_path / "r.fa"
    fa.write_text("".join(">r\n%s\n" % r for r in rand_reads(rng, 1500, (30, 200), 20000, perr=0.005)))
    fb = tmp_path / "s.fa"
    fb.write_text("".join(">r\n%s\n" % r for r in rand_reads(rng, 500, (30, 200), 20000, perr=0.005)))
    base = ["-q", "-f", "-m", "1G", "-n", "4M", "-k", str(k), "-s", "a", "-1", str(fa), "-s", "b", "-1", str(fb)]
    uns, srt = tmp_path / "uns.ctx", tmp_path / "sorted.ctx"
    _run(base + [str(uns)])
    _run(base[:2] + ["-S"] + base[2:] + [str(srt)])
    want = open(srt, "rb").read()
    assert open(uns, "rb").read() != want
    out = tmp_path / "out.ctx"
    _run_cmd("sort", ["-q", "-o", str(out), str(uns)])
    assert open(out, "rb").read() == want
    r = _run_cmd("sort", ["-q", "-o", str(out), str(uns)], check=False)
    assert r.returncode == 1 and b"File already exists" in r.stderr
    inplace = tmp_path / "inplace.ctx"
    shutil.copy(uns, inplace)
    _run_cmd("sort", ["-q", str(inplace)])
    assert open(inplace, "rb").read() == want
    assert _run_cmd("sort", ["-q", str(inplace) + ":0"], check=False).returncode == 1   # no filters
    if oracle.ref_binary(k) is not None:
        refcopy = tmp_path / "ref.ctx"
        shutil.copy(uns, refcopy)
        oracle.ref_run(k, ["sort", "-q", str(refcopy)])
        assert open(refcopy, "rb").read() == want


def test_cli_parallel_ingest_and_background_device_start(tmp_path, oracle):
    """uncompressed FASTA / plain files cut at record starts and parsed by several host threads
    (seq_ingest_par.c; tiny segments so that a small file is cut hundreds of times), the device table created on
    a second thread meanwhile: same bytes as the oracle's build, and as the sequential reader's"""
    rng = random.Random(77)
    reads = rand_reads(rng, 3000, (20, 300), 40000, perr=0.004) + EDGE_READS
    fa = tmp_path / "a.fa"
    with open(fa, "w") as f:
        for i, r in enumerate(reads):
            f.write(">r%d >x @y\n" % i)
            for j in range(0, len(r), 60):
                f.write(r[j:j + 60] + ("\r\n" if i % 7 == 0 else "\n"))
    plain = tmp_path / "b.txt"
    plain.write_text("\n".join(r for r in reads if r) + "\n")
    # FASTQ with a quality cut-off: strict records go to the worker threads, the multi-line record near the end
    # hands the rest of the file back to the sequential reader
    fq = tmp_path / "c.fq"
    with open(fq, "w") as f:
        for i, r in enumerate(x for x in reads if x):
            q = "".join(rng.choice("!#+5@IIIIII") for _ in r)
            if i == 2500 and len(r) > 10:
                f.write("@r%d\n%s\n%s\n+\n%s\n%s\n" % (i, r[:7], r[7:], q[:7], q[7:]))
            else:
                f.write("@r%d\n%s\n+\n%s\n" % (i, r, q))
    want, _ = oracle.build_ctx(31, [("fa", [str(fa)]), ("plain", [str(plain)]), ("fq", [dict(path=str(fq), fq_cutoff=12)])])
    args = ["-q", "-f", "-m", "1G", "-n", "2M", "-k", "31", "-S", "-s", "fa", "-1", str(fa), "-s", "plain", "-1", str(plain),
            "-s", "fq", "-Q", "12", "-1", str(fq)]
    for env in ({"MCX_PARSE_THREADS": "4", "MCX_PARSE_SEG_BYTES": "5000"}, {"MCX_PARSE_THREADS": "1"}):
        out = tmp_path / "out.ctx"
        _run(args + [str(out)], env=env)
        assert open(out, "rb").read() == want, env


def test_cli_files_of_one_colour_loaded_concurrently(tmp_path, oracle):
    """consecutive --seq files of one colour are read by one thread each (gz, FASTQ with a cut-off, FASTA, plain mixed);
    the table updates commute and the header credits the batch (quirk Q1), so the bytes are those of the oracle's
    build -- and of the same command with MCX_FILE_THREADS=1 (one file after the other)"""
    import gzip
    rng = random.Random(31)
    genome_reads = rand_reads(rng, 2400, (20, 250), 20000, perr=0.004)
    files = []
    for i in range(6):
        part = genome_reads[i * 400:(i + 1) * 400]
        if i % 3 == 0:
            p = tmp_path / ("f%d.fa.gz" % i)
            with gzip.open(p, "wt") as f:
                f.write("".join(">r%d\n%s\n" % (j, r) for j, r in enumerate(part)))
        elif i % 3 == 1:
            p = tmp_path / ("f%d.fq" % i)
            p.write_text("".join("@r%d\n%s\n+\n%s\n" % (j, r, "".join(rng.choice("#5III") for _ in r)) for j, r in enumerate(part)))
        else:
            p = tmp_path / ("f%d.txt" % i)
            p.write_text("\n".join(part) + "\n")
        files.append(str(p))
    tasks0 = [dict(path=f, fq_cutoff=10) for f in files[:4]]
    tasks1 = [dict(path=f, fq_cutoff=10) for f in files[4:]]
    want, _ = oracle.build_ctx(25, [("a", tasks0), ("b", tasks1)])
    args = ["-q", "-f", "-m", "1G", "-n", "2M", "-k", "25", "-S", "-Q", "10", "-s", "a"]
    for f in files[:4]:
        args += ["-1", f]
    args += ["-s", "b"]
    for f in files[4:]:
        args += ["-1", f]
    for env in ({}, {"MCX_FILE_THREADS": "1"}):
        out = tmp_path / "out.ctx"
        _run(args + [str(out)], env=env)
        assert open(out, "rb").read() == want, env

