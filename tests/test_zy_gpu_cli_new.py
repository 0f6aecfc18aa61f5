"""GPU command-line tests added at the end of round 1, whose CUDA side has not yet run on a B200 (the host side of each
runs on the CPU stand-in via tests/test_host_cli.py): the pipelined record writer, `join`, `build -D a,b,..`.  Kept in a
module that is collected after the bulk of the GPU suite."""
import hashlib
import json
import os
import random
import subprocess

import pytest

import test_gpu_cli as G
from conftest import ROOT, rand_reads, EDGE_READS
from test_gpu_cli import GOLD

pytestmark = pytest.mark.gpu


def _run(args, cwd=None, check=True, env=None):
    return G._run(args, cwd=cwd, check=check, env=env)


def _run_cmd(cmd, args, check=True):
    return G._run_cmd(cmd, args, check=check)


def test_cli_pipelined_writer(tmp_path):
    """outputs >= 256 MB are written by four pwrite threads from pinned chunks (write_records_parallel, ctx_build.c);
    forced here on a small graph with 100-record chunks: same file as the plain fwrite loop, header included"""
    rng = random.Random(91)
    fa = tmp_path / "r.fa"
    fa.write_text("".join(">r\n%s\n" % r for r in rand_reads(rng, 1500, 150, 30000, perr=0.005)))
    outs = []
    for env in ({"MCX_OUT_PIPE_MIN": "1", "MCX_OUT_CHUNK_RECS": "100"}, {}):
        for k, sort in ((31, True), (63, False)):
            out = tmp_path / "o.ctx"
            _run(["-q", "-f", "-m", "1G", "-n", "4M", "-k", str(k), "-s", "s", "-1", str(fa)] + (["-S"] if sort else []) + [str(out)], env=env)
            outs.append(open(out, "rb").read())
    assert outs[0] == outs[2] and len(outs[0]) > 100000

    def recs(ctx, rb):   # unsorted dumps: same header, same records in some order
        h = ctx.index(b"CORTEX", 6) + 6
        assert (len(ctx) - h) % rb == 0
        return ctx[:h], sorted(ctx[i:i + rb] for i in range(h, len(ctx), rb))
    assert recs(outs[1], 21) == recs(outs[3], 21)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mccortex31")), reason="oracle/_ref not shipped")
@pytest.mark.parametrize("k", [21, 39])
def test_cli_join_matches_reference_binary(tmp_path, oracle, k):
    """`mccortex-b200 join` (src/commands/ctx_join.c without --intersect): colours side by side, on top of each other,
    picked by filter, with offsets; the merged header; -S output byte-identical to the reference's join, unsorted
    multi-file output identical after the reference's `sort`, single-file stream filter identical as it is"""
    r = oracle.ref_run(k, ["join"], check=False)
    if b"usage" not in r.stderr + r.stdout or b"unknown command" in r.stderr:
        pytest.skip("oracle/_ref was built without join")
    rng = random.Random(5000 + k)
    reads = rand_reads(rng, 1800, (20, 220), 12000, perr=0.004)
    fas = []
    for i in range(3):
        p = tmp_path / ("r%d.fa" % i)
        p.write_text("".join(">r%d\n%s\n" % (j, x) for j, x in enumerate(reads[i * 600:(i + 1) * 600])))
        fas.append(str(p))
    a, b = str(tmp_path / "a.ctx"), str(tmp_path / "b.ctx")
    oracle.ref_build(k, ["-s", "sa0", "-1", fas[0], "-s", "sa1", "-1", fas[1]], a, nkmers="2M", sort=False)   # 2 colours, unsorted
    oracle.ref_build(k, ["-s", "sb", "-1", fas[2], "-1", fas[0]], b, nkmers="2M")                             # 1 colour
    cases = [[a, b], [b + ":0", "0:" + a + ":1"], ["1:" + b, a + ":1,0"], [a + ":0-1", "0:" + b, "3:" + a + ":0"], [a]]
    for n, files in enumerate(cases):
        mine, ref = str(tmp_path / ("mine%d.ctx" % n)), str(tmp_path / ("ref%d.ctx" % n))
        _run_cmd("join", ["-q", "-f", "-m", "1G", "-n", "2M", "-S", "-o", mine] + files)
        oracle.ref_run(k, ["join", "-q", "-f", "-m", "1G", "-n", "2M", "-S", "-o", ref] + files)
        assert open(mine, "rb").read() == open(ref, "rb").read(), files
        assert oracle.ref_run(k, ["check", "-q", mine], check=False).returncode == 0
    # unsorted: several files -> same after `sort`; one file through a filter -> the stream keeps the input order
    mine, ref = str(tmp_path / "mu.ctx"), str(tmp_path / "ru.ctx")
    _run_cmd("join", ["-q", "-f", "-m", "1G", "-n", "2M", "-o", mine, a, b])
    oracle.ref_run(k, ["join", "-q", "-f", "-m", "1G", "-n", "2M", "-o", ref, a, b])
    oracle.ref_run(k, ["sort", "-q", mine]); oracle.ref_run(k, ["sort", "-q", ref])
    assert open(mine, "rb").read() == open(ref, "rb").read()
    for flt in (a + ":1", a + ":1,0", "2:" + a + ":0"):
        _run_cmd("join", ["-q", "-f", "-o", mine, flt])
        oracle.ref_run(k, ["join", "-q", "-f", "-o", ref, flt])
        assert open(mine, "rb").read() == open(ref, "rb").read(), flt
    # errors: no output, existing output, mixed kmer sizes, --intersect
    assert _run_cmd("join", ["-q", a], check=False).returncode == 1
    r = _run_cmd("join", ["-q", "-o", mine, a], check=False)
    assert r.returncode == 1 and b"File already exists" in r.stderr
    other = str(tmp_path / "k.ctx")
    oracle.ref_build(k + 2, ["-s", "x", "-1", fas[0]], other, nkmers="2M")
    assert _run_cmd("join", ["-q", "-f", "-o", mine, a, other], check=False).returncode == 1
    assert _run_cmd("join", ["-q", "-f", "-o", mine, "-i", b, a], check=False).returncode == 1


def test_cli_replicas_on_several_devices(tmp_path, oracle):
    """build -D a,b,c: one replica of the graph per listed device (here three on device 0), batches of reads dealt round
    robin, replicas folded into the first before the dump (coverage adds, edges OR): same bytes as the oracle's build and as
    the one-device run, also with --graph (loaded into the first replica only) and --intersect (looked up in every replica)"""
    rng = random.Random(123)
    reads = rand_reads(rng, 3000, (20, 250), 15000, perr=0.004)
    fas = []
    for i in range(3):
        p = tmp_path / ("r%d.fa" % i)
        p.write_text("".join(">r%d\n%s\n" % (j, x) for j, x in enumerate(reads[i * 1000:(i + 1) * 1000])))
        fas.append(str(p))
    env = {"MCX_BATCH_BYTES": "20000"}    # many batches, so that every replica gets reads of every file
    base = ["-q", "-f", "-m", "1G", "-n", "2M", "-k", "27", "-S"]
    args = ["-s", "a", "-1", fas[0], "-1", fas[1], "-s", "b", "-H", "5", "-1", fas[2]]
    want, _ = oracle.build_ctx(27, [("a", fas[:2]), ("b", [dict(path=fas[2], hp_cutoff=5)])])
    one, many = str(tmp_path / "one.ctx"), str(tmp_path / "many.ctx")
    _run(base + args + [one], env=env)
    _run(base + ["-D", "0,0,0"] + args + [many], env=env)
    assert open(one, "rb").read() == want and open(many, "rb").read() == want
    # --graph + --intersect
    g0, isec = str(tmp_path / "g0.ctx"), str(tmp_path / "isec.ctx")
    _run(base + ["-s", "g", "-1", fas[0], g0], env=env)
    _run(base + ["-s", "i", "-1", fas[1], isec], env=env)
    args2 = ["-I", isec, "-g", g0, "-s", "n", "-1", fas[2], "-1", fas[0]]
    _run(base + args2 + [one], env=env)
    _run(base + ["-D", "0,0"] + args2 + [many], env=env)
    assert open(one, "rb").read() == open(many, "rb").read() and len(open(one, "rb").read()) > 1000
    r = _run(base + ["-D", "0,0", "-p", "-s", "x", "-1", fas[0], many], check=False)
    assert r.returncode == 1 and b"--remove-pcr" in r.stderr


JOIN_CASES = json.load(open(os.path.join(GOLD, "join_cases.json")))


@pytest.mark.parametrize("case", [c["name"] for c in JOIN_CASES])
def test_cli_join_matches_golden_ctx(case, tmp_path):
    """`mccortex-b200 join` on committed graph files: bytes == what the compiled reference's join wrote
    (tests/golden/make_golden_join.py) -- sorted merges and the unsorted single-file stream filter"""
    c = next(x for x in JOIN_CASES if x["name"] == case)
    out = str(tmp_path / "out.ctx")
    _run_cmd("join", ["-q", "-f", "-m", "1G", "-n", "1M"] + (["-S"] if c["sort"] else []) + ["-o", out] +
             [a.replace("@/", GOLD + "/") for a in c["args"]])
    ref = open(os.path.join(GOLD, c["ctx"]), "rb").read()
    assert hashlib.md5(ref).hexdigest() == c["md5"]
    assert open(out, "rb").read() == ref


def _devices_for_shard():
    """two device ids for `-D a,b --shard`: the CPU stand-in takes any; the real driver needs two GPUs"""
    if "hostcheck" in os.path.basename(G._driver()):
        return "0,1"
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("--shard needs two GPUs")
    return "0,1"


@pytest.mark.parametrize("k", [31, 63])
def test_cli_shard_mode(tmp_path, oracle, k):
    """`build -D 0,1 --shard`: ONE graph hash-partitioned over the devices (mcx_shardset_*), dump = merge of the shards'
    sorted runs: the bytes of the single-device build (and of the reference binary); several colours, several files, a batch
    size that cuts the input into many pieces; refusals"""
    devs = _devices_for_shard()
    rng = random.Random(300 + k)
    fas = []
    for i in range(3):
        fa = tmp_path / ("s%d.fa" % i)
        fa.write_text("".join(">r\n%s\n" % r for r in rand_reads(rng, 2500, (60, 200), 40000, perr=0.004) + EDGE_READS))
        fas.append(str(fa))
    args = ["-q", "-f", "-m", "1G", "-n", "4M", "-k", str(k), "-S", "-s", "a", "-1", fas[0], "-1", fas[1], "-s", "b", "-1", fas[2]]
    one, two = str(tmp_path / "one.ctx"), str(tmp_path / "two.ctx")
    _run(args + [one])
    _run(args[:2] + ["-D", devs, "--shard"] + args[2:] + [two], env={"MCX_BATCH_BYTES": "200000"})
    assert open(one, "rb").read() == open(two, "rb").read()
    if oracle.ref_binary(k):
        ref = str(tmp_path / "ref.ctx")
        oracle.ref_run(k, ["build", "-q", "-f", "-m", "1G", "-n", "4M", "-k", str(k), "-S", "-s", "a", "-1", fas[0], "-1", fas[1], "-s", "b", "-1", fas[2], ref])
        assert open(ref, "rb").read() == open(two, "rb").read()
    # unsorted dump: the same records in shard order; `sort` brings it to the same file
    uns = str(tmp_path / "uns.ctx")
    _run(["-q", "-f", "-D", devs, "--shard", "-m", "1G", "-n", "4M", "-k", str(k), "-s", "a", "-1", fas[0], "-1", fas[1], "-s", "b", "-1", fas[2], uns])
    _run_cmd("sort", ["-q", uns])
    assert open(uns, "rb").read() == open(one, "rb").read()
    # what the production pipeline passes (scripts/make-pipeline.pl:342-346): FASTQ with --fq-cutoff, plus a homopolymer cut-off
    fq = tmp_path / "q.fq"
    reads = rand_reads(rng, 3000, (40, 220), 30000, perr=0.004)
    with open(fq, "w") as f:
        for i, r in enumerate(reads):
            if r:
                f.write("@r%d\n%s\n+\n%s\n" % (i, r, "".join(chr(33 + rng.choice((40,) * 40 + (2, 9, 10, 11, 30))) for _ in r)))
    qargs = ["-q", "-f", "-m", "1G", "-n", "4M", "-k", str(k), "-S", "-s", "q", "-Q", "10", "-H", "5", "-1", str(fq)]
    _run(qargs + [one])
    _run(qargs[:2] + ["-D", devs, "--shard"] + qargs[2:] + [two], env={"MCX_BATCH_BYTES": "150000"})
    assert open(one, "rb").read() == open(two, "rb").read() and os.path.getsize(one) > 10000
    # refusals: one device, graph files
    assert _run(["-q", "-f", "--shard", "-m", "1G", "-n", "1M", "-k", str(k), "-s", "a", "-1", fas[0], two], check=False).returncode == 1
    assert _run(["-q", "-f", "-D", devs, "--shard", "-m", "1G", "-n", "1M", "-k", str(k), "-s", "a", "-1", fas[0], "-g", one, two], check=False).returncode == 1
