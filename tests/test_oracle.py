"""CPU tests: the oracle against golden vectors / the compiled reference, and the
device math (run on the CPU through tests/emul) against the oracle."""
import hashlib
import os
import random
import subprocess

import pytest

from conftest import ROOT, rand_reads, oracle_records, EDGE_READS

GOLD = os.path.join(ROOT, "tests", "golden")


# ---- known-answer vectors taken from the compiled reference -------------
def test_lookup3_known_answers(oracle):
    # `mccortex31 hashtest -t 1 -k 31 -F N` prints the XOR of bklk3_hashlittle({i}, 0), i < N
    L = oracle.lib()
    assert L.orc_hashtest_xor(1) == 1489077439
    assert L.orc_hashtest_xor(1000) == 2609835747
    assert L.orc_hashtest_xor(1000000) == 2564219928


def test_tiny_two_colour_graph_md5(oracle, tmp_path):
    g1 = tmp_path / "g1.fa"
    g2 = tmp_path / "g2.fa"
    g1.write_text(">r1\nACGTACGTTAGCNNACGTTAGCATCGATCGGATCGAT\n>r2\nacgtacgttagc\n>r3\nAC\n"
                  ">r4\nTTTTTTTTTTTTTTT\nGGGGGGGGGCA\n")
    g2.write_text(">s1\nATCGATCCGATCGATGCTAACGT\n")
    out, _ = oracle.build_ctx(11, [("one", [str(g1)]), ("two", [str(g2)])])
    assert len(out) == 634
    assert hashlib.md5(out).hexdigest() == "3e484f5ef0bdd644ffd233cde667c426"


def test_seq_err_bytes(oracle):
    # [probed] header seq_err bytes for a colour with sequence
    g = oracle.Graph(31, 1, 1024)
    g.set_name(0, "samp")
    st = g.add_read("ACGT" * 20)
    g.update_ginfo(0, st)
    h = g.header()
    assert len(h) == 89
    assert bytes.fromhex("00d8a3703d0ad7a3f83f000000000000") in h


def _golden_cases():
    import json
    with open(os.path.join(GOLD, "cases.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", [c["name"] for c in __import__("json").load(open(os.path.join(GOLD, "cases.json")))])
def test_oracle_vs_golden_ctx(oracle, case):
    """tests/golden/*.ctx were written by the compiled reference (make_golden.py)."""
    c = next(x for x in _golden_cases() if x["name"] == case)
    samples = [(s["name"], [dict(path=os.path.join(GOLD, t["file"]), fq_cutoff=t.get("fq_cutoff", 0),
                                 fq_offset=t.get("fq_offset", 0), hp_cutoff=t.get("hp_cutoff", 0))
                            for t in s["tasks"]]) for s in c["samples"]]
    out, _ = oracle.build_ctx(c["k"], samples)
    with open(os.path.join(GOLD, c["ctx"]), "rb") as f:
        ref = f.read()
    assert hashlib.md5(ref).hexdigest() == c["md5"]
    assert out == ref


GRAPH_CASES = __import__("json").load(open(os.path.join(GOLD, "graph_cases.json")))


@pytest.mark.parametrize("case", [c["name"] for c in GRAPH_CASES])
def test_oracle_graph_loading_vs_golden_ctx(oracle, case):
    """build --graph: the restatement of graph_load / file_filter / graph_info_merge against what the
    compiled reference wrote (tests/golden/graph_cases.json, make_golden.py)."""
    c = next(x for x in GRAPH_CASES if x["name"] == case)
    out = oracle.build_ctx_args(c["k"], [a.replace("@/", GOLD + "/") for a in c["ref_args"]])
    with open(os.path.join(GOLD, c["ctx"]), "rb") as f:
        ref = f.read()
    assert hashlib.md5(ref).hexdigest() == c["md5"]
    assert out == ref


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mccortex31")),
                    reason="oracle/_ref not built")
def test_oracle_graph_loading_vs_reference_binary(oracle, tmp_path):
    """fresh seeded inputs: two graphs built by the reference, then merged with reads through filters"""
    rng = random.Random(99)
    fas = []
    for i in range(3):
        p = tmp_path / ("r%d.fa" % i)
        p.write_text("".join(">r%d\n%s\n" % (j, r) for j, r in enumerate(rand_reads(rng, 120, (20, 220), 5000))))
        fas.append(str(p))
    k = 27
    g2 = str(tmp_path / "g2.ctx")
    oracle.ref_build(k, ["-s", "a", "-1", fas[0], "-s", "b", "-1", fas[1]], g2, threads=2)
    g1 = str(tmp_path / "g1.ctx")
    oracle.ref_build(k, ["-s", "c", "-1", fas[2]], g1, threads=2, sort=False)
    for args in (["-g", g2, "-g", g1, "-s", "n", "-1", fas[0]],
                 ["-g", "1:" + g2 + ":0", "-s", "n", "-1", fas[1], "-1", fas[2]],
                 ["-g", g2 + ":1-0", "-g", "0,0:" + g2, "-s", "n", "-1", fas[2]]):
        ref = oracle.ref_build(k, args, str(tmp_path / "ref.ctx"), threads=3)
        assert oracle.build_ctx_args(k, args) == ref, args


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mccortex31")),
                    reason="oracle/_ref not built")
@pytest.mark.parametrize("k,seed", [(31, 1), (63, 2), (33, 3), (21, 4), (5, 5)])
def test_oracle_vs_reference_binary(oracle, tmp_path, k, seed):
    rng = random.Random(seed)
    reads = rand_reads(rng, 300, (10, 260), 8000)
    fa = tmp_path / "r.fa"
    fa.write_text("".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(reads)))
    mine, _ = oracle.build_ctx(k, [("s", [str(fa)])])
    ref = oracle.ref_build(k, ["-s", "s", "-1", str(fa)], str(tmp_path / "ref.ctx"), threads=3)
    assert mine == ref


# ---- table sizing (row I) ----------------------------------------------------------------
def test_hash_table_cap(oracle):
    import ctypes as C
    nb, bs = C.c_uint64(), C.c_uint8()
    # [probed] `-n 64M` -> 2^21 buckets x 32
    cap = oracle.lib().orc_hash_table_cap(64 << 20, C.byref(nb), C.byref(bs))
    assert (nb.value, bs.value, cap) == (1 << 21, 32, 64 << 20)
    cap = oracle.lib().orc_hash_table_cap(1, C.byref(nb), C.byref(bs))
    assert (nb.value, bs.value, cap) == (1024, 1, 1024)


# ---- device math on the CPU --------------------------------------------------------------
def _emul(emul, tmp_path, reads, k, hp=0, piece=0):
    p = tmp_path / "lines.txt"
    p.write_text("".join(r + "\n" for r in reads))
    r = subprocess.run([emul, str(p), str(k), str(hp), str(piece)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       check=True)
    cnt = dict(x.split("=") for x in r.stderr.decode().split())
    return r.stdout, {a: int(b) for a, b in cnt.items()}


@pytest.mark.parametrize("k", [3, 11, 31, 33, 47, 63])
def test_device_math_matches_oracle(oracle, emul, tmp_path, reads_small, k):
    recs, st = oracle_records(oracle, reads_small, k)
    got, cnt = _emul(emul, tmp_path, reads_small, k)
    assert got == recs
    assert cnt == dict(kmers=st.num_kmers_loaded, novel=st.num_kmers_novel, contigs=st.contigs_parsed,
                       reads=len(reads_small))


def _emul_lane(emul, tmp_path, reads, k, piece=0):
    p = tmp_path / "lines.txt"
    p.write_text("".join(r + "\n" for r in reads))
    r = subprocess.run([emul, "--lane", str(p), str(k), str(piece)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    cnt = dict(x.split("=") for x in r.stderr.decode().split())
    return r.stdout, {a: int(b) for a, b in cnt.items()}


@pytest.mark.parametrize("k", [3, 5, 11, 15, 17, 21, 29, 31])
def test_lane_front_end_matches_oracle(oracle, emul, tmp_path, reads_small, k):
    """the warp-autonomous front end (mcx_lane.cuh: 16 windows per lane, pieces of 16 bytes) against the oracle"""
    recs, st = oracle_records(oracle, reads_small, k)
    got, cnt = _emul_lane(emul, tmp_path, reads_small, k)
    assert got == recs
    assert cnt == dict(kmers=st.num_kmers_loaded, novel=st.num_kmers_novel, contigs=st.contigs_parsed,
                       reads=len(reads_small))


@pytest.mark.parametrize("piece", [16, 160, 512, 2048, 4112])
def test_lane_front_end_staging_cuts(oracle, emul, tmp_path, reads_small, piece):
    for k in (31, 19):
        recs, st = oracle_records(oracle, reads_small, k)
        got, cnt = _emul_lane(emul, tmp_path, reads_small, k, piece=piece)
        assert got == recs
        assert cnt["kmers"] == st.num_kmers_loaded and cnt["reads"] == len(reads_small) and cnt["contigs"] == st.contigs_parsed


def test_lane_front_end_odd_reads(oracle, emul, tmp_path):
    """reads shorter than k, empty reads, one very long read, reads that end exactly on piece / tile boundaries"""
    import random
    rng = random.Random(77)
    long_read = "".join(rng.choice("ACGT") for _ in range(20000))
    reads = ["", "A", "ACGT" * 7 + "AC", "ACGT" * 8, long_read, "", "N" * 40, "acgtn" * 30, long_read[100:100 + 511], long_read[7:7 + 15],
             long_read[50:50 + 527], "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACG" + "-" + "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT"]
    for k in (31, 15, 3):
        recs, st = oracle_records(oracle, reads, k)
        for piece in (0, 48):
            got, cnt = _emul_lane(emul, tmp_path, reads, k, piece=piece)
            assert got == recs
            assert cnt == dict(kmers=st.num_kmers_loaded, novel=st.num_kmers_novel, contigs=st.contigs_parsed, reads=len(reads))


@pytest.mark.parametrize("k,hp", [(11, 2), (21, 5), (31, 4), (31, 31), (63, 6)])
def test_device_math_homopolymer_cutoff(oracle, emul, tmp_path, reads_small, k, hp):
    recs, st = oracle_records(oracle, reads_small, k, hp_cutoff=hp)
    got, cnt = _emul(emul, tmp_path, reads_small, k, hp=hp)
    assert got == recs
    assert cnt["contigs"] == st.contigs_parsed and cnt["kmers"] == st.num_kmers_loaded


@pytest.mark.parametrize("piece", [16, 160, 2048, 4112])
def test_device_math_staging_cuts(oracle, emul, tmp_path, reads_small, piece):
    """host batches are cut into pieces with 16 B look-back / 80 B look-ahead (mcx_abi.cu)"""
    for k, hp in ((31, 0), (63, 4)):
        recs, st = oracle_records(oracle, reads_small, k, hp_cutoff=hp)
        got, cnt = _emul(emul, tmp_path, reads_small, k, hp=hp, piece=piece)
        assert got == recs
        assert cnt["kmers"] == st.num_kmers_loaded and cnt["reads"] == len(reads_small)


def _emul_q(emul, tmp_path, reads, quals, k, qcut, hp=0):
    p = tmp_path / "lines.txt"
    q = tmp_path / "qual.txt"
    p.write_text("".join(r + "\n" for r in reads))
    with open(q, "wb") as f:
        for r, s in zip(reads, quals):
            b = s.encode("latin1")[:len(r)]
            f.write(b + b"\x7f" * (len(r) - len(b)) + b"!")
    r = subprocess.run([emul, str(p), str(k), str(hp), "0", str(q), str(qcut)], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, check=True)
    cnt = dict(x.split("=") for x in r.stderr.decode().split())
    return r.stdout, {a: int(b) for a, b in cnt.items()}


@pytest.mark.parametrize("k,hp,cut,eqp", [(21, 0, 43, 0.05), (31, 0, 50, 0.3), (31, 4, 43, 0.9), (63, 5, 60, 0.5)])
def test_device_math_quality_cutoff(oracle, emul, tmp_path, k, hp, cut, eqp):
    """start (qual > cutoff) / extend (qual >= cutoff) asymmetry via the carry chain + chunk summaries"""
    rng = random.Random(k * 100 + cut)
    reads = rand_reads(rng, 300, (1, 500), 6000, perr=0.01, pN=0.005)
    reads += rand_reads(rng, 2, 9000, 12000, perr=0, pN=0.0003, lower=0)   # carries across chunks
    quals = ["".join(chr(cut) if rng.random() < eqp else chr(rng.randint(35, 74)) for _ in r) for r in reads]
    for i in range(0, len(reads), 7):
        quals[i] = quals[i][:len(quals[i]) // 2]
    for i in range(3, len(reads), 11):
        quals[i] = ""
    g = oracle.Graph(k, 1, 1 << 21)
    st = oracle.Stats()
    for r, q in zip(reads, quals):
        g.add_read(r, qual=q.encode("latin1") if q else None, fq_cutoff=cut, hp_cutoff=hp, stats=st)
    recs = g.dump_sorted()[len(g.header()):]
    got, cnt = _emul_q(emul, tmp_path, reads, quals, k, cut, hp)
    assert got == recs
    assert cnt["kmers"] == st.num_kmers_loaded and cnt["contigs"] == st.contigs_parsed


# ---- build --remove-pcr (row N3) -------------------------------------------------------------
def _pcr_lines(units, quals, matedir):
    """LINES + parallel quality bytes + mate bytes as the host driver hands them to mcx_graph_add_reads_pcr"""
    lines, qlines, mates = [], [], bytearray()
    for u, q in zip(units, quals):
        for m, (r, qs) in enumerate(zip(u, q)):
            lines.append(r)
            qb = qs[:len(r)]
            qlines.append(qb + "\x7f" * (len(r) - len(qb)))
            kind = 0 if len(u) == 1 else 1 + m
            flip = (matedir & 2) if m == 0 else (matedir & 1)
            mates.append(kind | (4 if flip else 0))
    return lines, qlines, bytes(mates)


@pytest.mark.parametrize("k,hp,cut,matedir,batch", [(21, 0, 0, 1, 0), (31, 0, 0, 3, 37), (21, 4, 45, 1, 0), (33, 0, 50, 2, 64),
                                                  (63, 5, 0, 0, 101), (11, 0, 40, 1, 2)])
def test_device_math_remove_pcr(oracle, emul, tmp_path, k, hp, cut, matedir, batch):
    """orient / mark (atomicMin of read ordinals) / mask as in mcx_pcr.cu against the read-by-read bit test
    of seq_reads_are_novel restated in the oracle"""
    from conftest import pcr_reads
    rng = random.Random(1000 * k + cut + matedir)
    units = pcr_reads(rng, 500, k)
    quals = [tuple("".join(chr(cut) if rng.random() < 0.03 else chr(rng.randint(max(cut - 3, 35), 74)) for _ in r) if cut else ""
                   for r in u) for u in units]
    g = oracle.Graph(k, 1, 1 << 20)
    st = oracle.Stats()
    for u, q in zip(units, quals):
        g.add_reads_pcr(u[0], q[0] or None, u[1] if len(u) > 1 else None, (q[1] or None) if len(u) > 1 else None,
                        fq_cutoff=cut, hp_cutoff=hp, matedir=matedir, stats=st)
    recs = g.dump_sorted()[len(g.header()):]
    lines, qlines, mates = _pcr_lines(units, quals, matedir)
    (tmp_path / "l.txt").write_text("".join(r + "\n" for r in lines))
    (tmp_path / "q.txt").write_bytes("".join(q + "!" for q in qlines).encode("latin1"))
    (tmp_path / "m.bin").write_bytes(mates)
    r = subprocess.run([emul, "--pcr", str(tmp_path / "l.txt"), str(k), str(hp), str(tmp_path / "q.txt") if cut else "-",
                        str(cut), str(tmp_path / "m.bin"), str(batch)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    cnt = {a: int(b) for a, b in (x.split("=") for x in r.stderr.decode().split())}
    assert r.stdout == recs
    assert cnt["dupse"] == st.num_dup_se_reads and cnt["duppe"] == st.num_dup_pe_pairs
    assert st.num_dup_se_reads + st.num_dup_pe_pairs > 50
    assert cnt["kmers"] == st.num_kmers_loaded and cnt["contigs"] == st.contigs_parsed


@pytest.mark.parametrize("mode,k,cut,hp,matedir,batch_bytes", [("se", 21, 0, 0, 1, 0), ("se", 31, 0, 0, 3, 3000), ("pe", 21, 10, 4, 1, 0),
                                                            ("pe", 33, 0, 0, 2, 5000), ("il", 21, 0, 0, 1, 2000), ("il", 21, 12, 0, 3, 0)])
def test_host_ingest_remove_pcr(oracle, emul, ingest_dump, tmp_path, mode, k, cut, hp, matedir, batch_bytes):
    """host driver's pairing (file pairs, interleaved names), mate bytes and batch cuts + the device math
    == the oracle's restatement of seq_parse_*_sf + build_graph_from_reads_mt with --remove-pcr"""
    from conftest import write_pcr_files
    write_pcr_files(random.Random(k + cut), str(tmp_path))
    f1 = str(tmp_path / {"se": "se.fa", "pe": "p1.fq", "il": "il.fq"}[mode])
    f2 = str(tmp_path / "p2.fq") if mode == "pe" else None
    g = oracle.Graph(k, 1, 1 << 20)
    st = g.load_pcr(f1, f2, mode == "il", fq_cutoff=cut, hp_cutoff=hp, matedir=matedir)
    recs = g.dump_sorted()[len(g.header()):]
    env = dict(os.environ)
    if batch_bytes:
        env["MCX_BATCH_BYTES"] = str(batch_bytes)
    pre = str(tmp_path / "d")
    out = subprocess.run([ingest_dump, pre, mode, str(cut), "0", str(hp), str(matedir), f1] + ([f2] if f2 else []),
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True, env=env).stdout.decode()
    info = dict(x.split("=") for x in out.split())
    assert int(info["nbatches"]) > (3 if batch_bytes else 0)
    assert (int(info["se"]), int(info["pe"])) == (st.num_se_reads, st.num_pe_reads)
    r = subprocess.run([emul, "--pcr", pre + ".lines", str(k), str(hp), pre + ".qual" if cut else "-", info["fq_cutoff"],
                        pre + ".mate", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    cnt = {a: int(b) for a, b in (x.split("=") for x in r.stderr.decode().split())}
    assert r.stdout == recs
    assert (cnt["dupse"], cnt["duppe"]) == (st.num_dup_se_reads, st.num_dup_pe_pairs)
    assert st.num_dup_se_reads + st.num_dup_pe_pairs > 20


# ---- multi-threaded ingest (seq_ingest_par.c) == the sequential reader ----------------------------------------
def _tricky_fasta(rng, nrec):
    out = []
    for i in range(nrec):
        hdr = ">" + "".join(rng.choice("ACGTxyz >@|_") for _ in range(rng.randint(0, 30)))
        eol = rng.choice(["\n", "\n", "\n", "\r\n"])
        out.append(hdr + eol)
        nlines = rng.choice([0, 1, 1, 1, 2, 5])
        for _ in range(nlines):
            line = "".join(rng.choice("ACGTacgtNn") for _ in range(rng.randint(1, 90)))
            if rng.random() < 0.03:
                line = line[:len(line) // 2] + "\r" + line[len(line) // 2:]      # a CR inside a line stays in the read
            out.append(line + rng.choice(["\n", "\n", "\r\n", "\r\r\n"]))
            if rng.random() < 0.05:
                out.append(rng.choice(["\n", "\r\n", "\r"]))                      # empty lines / a bare CR at a line start
    return "".join(out)


@pytest.mark.parametrize("kind,seg,tail", [("fasta", 700, "\n"), ("fasta", 5000, ""), ("fasta", 64, ">last"), ("fasta", 300, ">"),
                                           ("plain", 500, "\n"), ("plain", 90, "ACGT")])
def test_parallel_ingest_matches_sequential_reader(ingest_dump, tmp_path, kind, seg, tail):
    """the LINES bytes handed to mcx_graph_add_reads are identical whether the file goes through the sequential
    reader (MCX_PARSE_THREADS=1) or is cut at record starts and parsed by 4 threads (tiny segments: hundreds of cuts);
    multi-line records, CR LF, empty lines, '>' and '@' inside headers, leading white-space lines, odd file ends"""
    rng = random.Random(seg)
    if kind == "fasta":
        txt = "\n\n" + _tricky_fasta(rng, 600) + tail
    else:
        lines = []
        for _ in range(3000):
            r = rng.random()
            if r < 0.05:
                lines.append(rng.choice(["", " ignored ACGT", "\tACGT", "\r"]))
            else:
                lines.append("".join(rng.choice("ACGTacgtN") for _ in range(rng.randint(1, 160))) + rng.choice(["", "", "\r"]))
        txt = "\n".join(lines) + "\n" + tail
    path = tmp_path / ("in." + kind)
    path.write_bytes(txt.encode())
    outs = []
    for threads in (1, 4):
        env = dict(os.environ, MCX_PARSE_THREADS=str(threads), MCX_PARSE_SEG_BYTES=str(seg), MCX_BATCH_BYTES="4000")
        pre = str(tmp_path / ("d%d" % threads))
        info = subprocess.run([ingest_dump, pre, "file", "0", "0", "0", "1", str(path)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, check=True, env=env).stdout.decode()
        info = dict(x.split("=") for x in info.split())
        outs.append((open(pre + ".lines", "rb").read(), int(info["nreads"]), int(info["nbatches"])))
    assert outs[0][0] == outs[1][0]
    assert outs[0][1] == outs[1][1] > 500
    assert outs[1][2] > 10     # the threaded run really was cut into many segments


def _fastq(rng, nrec, deviations):
    """4-line FASTQ with quality lines that often start with '@' or '+', CR LF here and there; `deviations` places
    records the strict parser must hand back to the sequential reader (multi-line, junk, short / missing quality)"""
    out = []
    for i in range(nrec):
        n = rng.randint(1, 120)
        seq = "".join(rng.choice("ACGTacgtN") for _ in range(n))
        qual = "".join(rng.choice("@+IIIIFFF#5<") for _ in range(n + rng.choice([0, 0, 0, 3])))
        eol = "\r\n" if rng.random() < 0.1 else "\n"
        kind = deviations.get(i)
        if kind == "multi" and n > 4:
            h = n // 2
            out.append("@r%d\n%s\n%s\n+r%d\n%s\n%s\n" % (i, seq[:h], seq[h:], i, qual[:h], qual[h:]))
        elif kind == "junk":
            out.append("@r%d\n%s\n+\n%s\nthis line is skipped\n\n" % (i, seq, qual))
        elif kind == "short":
            out.append("@r%d\n%s\n+\n%s\n%s\n" % (i, seq, qual[:n // 2], qual))   # a second quality line is read
        elif kind == "empty":
            out.append("@r%d\n\n%s\n+\n%s\n" % (i, seq, qual))
        else:
            out.append("@r%d some text%s%s%s+%s%s%s" % (i, eol, seq, eol, eol, qual, eol))
    return "".join(out)


@pytest.mark.parametrize("seg,cut,devs,tail", [(600, 0, {}, ""), (600, 10, {}, ""), (250, 12, {700: "multi"}, ""), (1000, 0, {300: "junk", 900: "multi"}, ""),
                                               (400, 10, {1500: "short"}, ""), (700, 0, {1100: "empty"}, ""), (500, 10, {}, "@last\nACGTACGT\n+\nIIIIIIII"),
                                               (500, 0, {}, "@last\nACGTACGT\n+"), (300, 10, {2: "multi"}, "")])
def test_parallel_fastq_ingest_matches_sequential_reader(ingest_dump, tmp_path, seg, cut, devs, tail):
    """FASTQ through the multi-threaded ingest: strict 4-line records are parsed by the workers, the first record that is
    not strict sends the rest of the file to the sequential reader -- sequence bytes, quality bytes, the cut-off
    (ASCII offset auto-detected) and the read count are those of the sequential reader alone"""
    rng = random.Random(seg + cut)
    path = tmp_path / "in.fq"
    path.write_bytes((_fastq(rng, 2000, devs) + tail).encode())
    outs = []
    for threads in (1, 4):
        env = dict(os.environ, MCX_PARSE_THREADS=str(threads), MCX_PARSE_SEG_BYTES=str(seg), MCX_BATCH_BYTES="3000")
        pre = str(tmp_path / ("d%d" % threads))
        r = subprocess.run([ingest_dump, pre, "file", str(cut), "0", "0", "1", str(path)], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, env=env)
        info = dict(x.split("=") for x in r.stdout.decode().split())
        outs.append((open(pre + ".lines", "rb").read(), open(pre + ".qual", "rb").read(), info["fq_cutoff"], info["nreads"],
                     b"Input error" in r.stderr))
    assert outs[0] == outs[1]
    assert int(outs[0][3]) >= 2000


def test_backward_colour_ranges_are_the_reference_s(oracle, tmp_path):
    """a descending range a-b in a colour filter runs down to 0 in the reference (src/basic/range.c:67-68) and only the
    first range_get_num() entries are used: 'g.ctx:2-1,0,0-2' selects colours 2,1,0,0,0,1.  Oracle (CtxFile) == the compiled
    reference's `join` on exactly that"""
    if oracle.ref_binary(15) is None:
        pytest.skip("oracle/_ref not built")
    rng = random.Random(3)
    fas = []
    for i in range(3):
        p = tmp_path / ("r%d.fa" % i)
        p.write_text("".join(">r\n%s\n" % "".join(rng.choice("ACGT") for _ in range(60)) for _ in range(30)))
        fas.append(str(p))
    g3 = str(tmp_path / "g3.ctx")
    oracle.ref_build(15, ["-s", "a", "-1", fas[0], "-s", "b", "-1", fas[1], "-s", "c", "-1", fas[2]], g3, nkmers="100K")
    f = oracle.CtxFile(g3 + ":2-1,0,0-2")
    assert [a for a, _ in f.filter] == [2, 1, 0, 0, 0, 1] and [b for _, b in f.filter] == [0, 1, 2, 3, 4, 5]
    out = str(tmp_path / "j.ctx")
    oracle.ref_run(15, ["join", "-q", "-f", "-o", out, g3 + ":2-1,0,0-2"])
    j = oracle.CtxFile(out)
    assert j.ncols == 6
    src = oracle.CtxFile(g3)
    assert [g["total"] for g in j.ginfo] == [src.ginfo[c]["total"] for c in (2, 1, 0, 0, 0, 1)]


@pytest.mark.parametrize("block", range(6))
def test_device_math_fuzz(oracle, emul, tmp_path, block):
    """the kernels' host/device math (tests/emul) against the oracle on seeded adversarial reads: random odd k in 3..63,
    homopolymer cut-offs from 2 to k, quality cut-offs with many values exactly at the threshold, runs of one base around
    the cut-off length, N runs, reads shorter than / equal to / just over k, reads long enough to cross chunk (2048) and
    staging-piece boundaries; records and counters must be identical"""
    rng = random.Random(500 + block)
    for case in range(8):
        k = rng.choice(range(3, 64, 2))
        hp = rng.choice([0, 0, 2, 3, rng.randint(2, k), k])
        cut = rng.choice([0, 0, rng.randint(36, 73)])
        piece = rng.choice([0, 0, 16, 160, 2048, 4112])
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
        if cut:
            quals = ["".join(chr(rng.choice([cut - 1, cut, cut, cut + 1, rng.randint(35, 74)])) for _ in r) for r in reads]
            for i in range(0, len(reads), 5):
                quals[i] = quals[i][:rng.randint(0, len(quals[i]))]
        g = oracle.Graph(k, 1, 1 << 21)
        st = oracle.Stats()
        for i, r in enumerate(reads):
            g.add_read(r, qual=(quals[i].encode("latin1") or None) if cut else None, fq_cutoff=cut, hp_cutoff=hp, stats=st)
        recs = g.dump_sorted()[len(g.header()):]
        g.close()
        if cut:
            got, cnt = _emul_q(emul, tmp_path, reads, quals, k, cut, hp)
        else:
            got, cnt = _emul(emul, tmp_path, reads, k, hp=hp, piece=piece)
        assert got == recs, (block, case, k, hp, cut, piece)
        assert cnt["kmers"] == st.num_kmers_loaded and cnt["contigs"] == st.contigs_parsed, (block, case, k, hp, cut)


def test_front_table_hashes_are_bijections(emul):
    """(set, tag) of the L2 front tables identifies the key exactly, k <= 31 and k <= 63 (mcx_fhash / mcx_fhash2 and
    their inverses, mcx_device.cuh), and k-mer-like keys spread over the sets like random ones."""
    out = subprocess.check_output([emul, "--fhash", "2000000", "11"]).decode().split()
    vals = dict(x.split("=") for x in out)
    assert vals["bad"] == "0"
    mean = float(vals["mean"])
    assert int(vals["max_load1"]) < mean + 6 * mean ** 0.5 and int(vals["max_load2"]) < mean + 6 * mean ** 0.5


def test_warp_contig_chain_and_chunk_summary(emul):
    """Quality modes: the warp-wide in_contig chain (carry functions composed by a prefix scan) and the closed-form
    chunk summary -- the lane-local code the device runs (mcx_chain_lane_*, mcx_summary_lane_*, mcx_chunk.cuh), walked
    lane by lane -- equal the serial mcx_contig_chain (seq_contig_start2 / seq_contig_end2 with a quality cut-off,
    src/basic/seq_reader.c:61-172) on random masks of every density, for both carry-ins."""
    out = subprocess.check_output([emul, "--chain", "30000", "17"]).decode().split()
    vals = dict(x.split("=") for x in out)
    assert vals["bad"] == "0" and int(vals["cases"]) == 60000
    assert int(vals["carry_dependent"]) > 100 and int(vals["last_window_in_contig"]) > 10000

