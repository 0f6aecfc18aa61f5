"""GPU parity tests: the CUDA path, called through the C ABI (libmcxgpu.so), against the
oracle on the same seeded inputs -- bit-exact records and counters."""
import random

import pytest

from conftest import rand_reads, oracle_records, EDGE_READS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import mccortex_b200 as M
    assert M.device_count() > 0, "GPU tests need a CUDA device"
    return M


def _check_stats(st, ost, nreads):
    assert st.num_kmers_loaded == ost.num_kmers_loaded
    assert st.num_kmers_novel == ost.num_kmers_novel
    assert st.contigs_parsed == ost.contigs_parsed
    assert st.total_bases_loaded == ost.total_bases_loaded
    assert st.total_bases_read == ost.total_bases_read
    assert st.num_se_reads == nreads


@pytest.mark.parametrize("k", [3, 11, 21, 31, 33, 47, 63])
def test_lines_batch_matches_oracle(M, oracle, reads_small, k):
    recs, ost = oracle_records(oracle, reads_small, k)
    g = M.Graph(k, 1, 1 << 20)
    g.add_lines("".join(r + "\n" for r in reads_small).encode())
    st = g.sync()
    got, n, rb = g.export_records(sorted=True)
    assert rb == 8 * g.W + 5
    assert n == ost.num_kmers_novel
    assert got == recs
    _check_stats(st, ost, len(reads_small))
    assert g.stats()[0] == n
    g.close()


@pytest.mark.parametrize("k", [31, 63])
def test_offsets_batch_matches_oracle(M, oracle, reads_small, k):
    recs, ost = oracle_records(oracle, reads_small, k)
    g = M.Graph(k, 1, 1 << 20)
    g.prepare_host()   # mcx_graph_prepare_host: the staging ring up front (optional)
    g.add_reads(reads_small)
    st = g.sync()
    got, n, _ = g.export_records()
    assert got == recs
    _check_stats(st, ost, len(reads_small))
    g.close()


@pytest.mark.parametrize("k,hp", [(11, 2), (21, 5), (31, 4), (31, 31), (63, 6)])
def test_homopolymer_cutoff(M, oracle, reads_small, k, hp):
    recs, ost = oracle_records(oracle, reads_small, k, hp_cutoff=hp)
    g = M.Graph(k, 1, 1 << 20)
    g.add_lines("".join(r + "\n" for r in reads_small).encode(), hp_cutoff=hp)
    st = g.sync()
    got, _, _ = g.export_records()
    assert got == recs
    _check_stats(st, ost, len(reads_small))
    g.close()


@pytest.mark.parametrize("k,ncols", [(31, 2), (31, 4), (63, 3), (21, 7)])
def test_colours(M, oracle, k, ncols):
    """per-colour covg/edges arrays (reference: col_covgs / col_edges, db_graph.h:39-40)"""
    rng = random.Random(100 + ncols)
    og = oracle.Graph(k, ncols, 1 << 20)
    g = M.Graph(k, ncols, 1 << 20)
    for c in range(ncols):
        reads = rand_reads(rng, 120, (40, 200), 3000)  # same genome size, different genomes: some shared k-mers unlikely
        if c:
            reads += first[:40]  # make colours share k-mers
        else:
            first = reads
        for r in reads:
            og.add_read(r, colour=c)
        g.add_lines("".join(r + "\n" for r in reads).encode(), colour=c)
    g.sync()
    full = og.dump_sorted()
    recs = full[len(og.header()):]
    got, n, rb = g.export_records()
    assert rb == 8 * g.W + 5 * ncols
    assert got == recs
    g.close()


def test_batches_accumulate_and_unsorted_export(M, oracle, reads_small):
    """several add_reads calls == one; unsorted export is a permutation of the sorted one"""
    k = 31
    recs, ost = oracle_records(oracle, reads_small + reads_small[:100], k)
    g = M.Graph(k, 1, 1 << 20)
    g.add_lines("".join(r + "\n" for r in reads_small[:150]).encode())
    g.add_reads(reads_small[150:])
    for r in reads_small[:100]:
        if len(r) >= k and all(c in "ACGTacgt" for c in r):
            g.add_str(r)  # build_graph_from_str_mt
        else:
            g.add_reads([r])
    g.sync()
    got, n, rb = g.export_records(sorted=True)
    assert got == recs
    uns, n2, _ = g.export_records(sorted=False)
    assert n2 == n
    assert sorted(uns[i:i + rb] for i in range(0, len(uns), rb)) == sorted(got[i:i + rb] for i in range(0, len(got), rb))
    g.close()


def test_exactly_once_novelty_under_concurrency(M, oracle):
    """reference hash_table_tests.c:97-129: the same keys inserted from many threads must be
    novel exactly once.  Here: the same 200k-window batch submitted 6 times on concurrent streams."""
    rng = random.Random(5)
    reads = rand_reads(rng, 1500, 150, 50000, perr=0.0, pN=0.0, lower=0.0)
    blob = "".join(r + "\n" for r in reads).encode()
    recs, ost = oracle_records(oracle, reads, 31)
    g = M.Graph(31, 1, 1 << 19)
    for _ in range(6):
        g.add_lines(blob)
    st = g.sync()
    assert st.num_kmers_novel == ost.num_kmers_novel
    assert st.num_kmers_loaded == 6 * ost.num_kmers_loaded
    got, n, rb = g.export_records()
    # same keys and edges, 6x the coverage
    assert n == ost.num_kmers_novel
    for i in range(0, len(recs), rb * 997):
        assert got[i:i + 8] == recs[i:i + 8]
        assert int.from_bytes(got[i + 8:i + 12], "little") == 6 * int.from_bytes(recs[i + 8:i + 12], "little")
        assert got[i + 12] == recs[i + 12]
    g.close()


def test_table_full_is_reported(M):
    rng = random.Random(9)
    reads = rand_reads(rng, 200, 150, 20000, perr=0.0, pN=0.0, lower=0.0)
    g = M.Graph(31, 1, 1024)
    g.add_lines("".join(r + "\n" for r in reads).encode())
    with pytest.raises(M.McxError) as e:
        g.sync()
    assert e.value.status == "MCX_ERR_TABLE_FULL"
    g.close()


def test_large_host_batch_crosses_staging_pieces(M, oracle):
    """> 32 Mi positions so the host path cuts the batch (look-back/look-ahead logic on real HW)"""
    rng = random.Random(77)
    base = rand_reads(rng, 3000, 150, 200000, perr=0.002, pN=0.001, lower=0.05)
    blob = "".join(r + "\n" for r in base).encode()
    reps = (40 << 20) // len(blob) + 1
    big = blob * reps
    recs, ost = oracle_records(oracle, base, 31)
    g = M.Graph(31, 1, 1 << 20)
    g.add_lines(big)
    st = g.sync()
    assert st.num_kmers_loaded == reps * ost.num_kmers_loaded
    assert st.num_kmers_novel == ost.num_kmers_novel
    assert st.num_se_reads == reps * len(base)
    got, n, rb = g.export_records()
    assert n == ost.num_kmers_novel
    assert got[:8] == recs[:8] and got[-rb:-rb + 8] == recs[-rb:-rb + 8]
    for i in range(0, len(recs), rb * 1009):
        assert int.from_bytes(got[i + 8:i + 12], "little") == reps * int.from_bytes(recs[i + 8:i + 12], "little")
        assert got[i + 12] == recs[i + 12]
    g.close()


def _to_dev(torch, blob, dev):
    seq = torch.zeros(len(blob) + 64, dtype=torch.uint8, device=dev)
    seq[:len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    return seq


def test_tuple_path_matches_fused_path(M, oracle, reads_small):
    """kernel B (reads -> binned tuples, count 1 each) + kernel C (insert) == kernel A, and bins respect ownership"""
    import torch
    k, nparts = 31, 4
    blob = "".join(r + "\n" for r in reads_small).encode()
    recs, ost = oracle_records(oracle, reads_small, k)
    dev = torch.device("cuda:0")
    seq = _to_dev(torch, blob, dev)
    cap = ost.num_kmers_loaded + 16
    keys = torch.zeros(nparts * cap, dtype=torch.int64, device=dev)
    meta = torch.zeros(nparts * cap, dtype=torch.int32, device=dev)
    counts = torch.zeros(nparts, dtype=torch.int64, device=dev)
    g = M.Graph(k, 1, 1 << 20)
    g.kmer_tuples(seq.data_ptr(), len(blob), nparts, cap, keys.data_ptr(), meta.data_ptr(), counts.data_ptr())
    st = g.sync()
    assert st.num_kmers_loaded == ost.num_kmers_loaded and st.contigs_parsed == ost.contigs_parsed
    cnt = counts.cpu().tolist()
    assert sum(cnt) == ost.num_kmers_loaded
    for d in range(nparts):
        kd = keys[d * cap: d * cap + cnt[d]].cpu().tolist()
        md = meta[d * cap: d * cap + cnt[d]].cpu().tolist()
        assert all((m >> 8) == 1 for m in md)
        for x in kd[:50]:
            assert M.key_owner([x & 0xFFFFFFFFFFFFFFFF], k, nparts) == d
        g.insert_tuples(keys[d * cap:].data_ptr(), meta[d * cap:].data_ptr(), cnt[d])
    st2 = g.sync()
    assert st2.num_kmers_novel == ost.num_kmers_novel
    got, _, _ = g.export_records()
    assert got == recs
    g.close()


@pytest.mark.parametrize("k,nparts", [(31, 2), (31, 3), (63, 2), (21, 4)])
def test_sharded_build_on_one_device(M, oracle, k, nparts):
    """the multi-GPU algorithm with all shards on cuda:0: each shard runs the sharded kernel on its
    slice of the reads (local front-table aggregation + owner routing), bins are handed over by
    pointer, front tables are flushed, and the UNION of the shards' sorted records must equal the
    oracle's sorted records; every shard holds only keys it owns"""
    import torch
    rng = random.Random(31 * nparts + k)
    # high multiplicity (hot k-mers absorbed by the front tables) + errors (cold path)
    base = rand_reads(rng, 1500, 150, 8000, perr=0.004, pN=0.001, lower=0.02) + ["A" * 150] * 300
    rng.shuffle(base)
    recs, ost = oracle_records(oracle, base, k, capacity=1 << 22)
    dev = torch.device("cuda:0")
    W = (k + 31) // 32
    shards = [M.Graph(k, 1, 1 << 20) for _ in range(nparts)]
    cap = ost.num_kmers_loaded + 1024
    bins = []
    for p in range(nparts):
        mine = base[p::nparts]
        blob = "".join(r + "\n" for r in mine).encode()
        seq = _to_dev(torch, blob, dev)
        half = (len(mine) // 2)
        cut = len("".join(r + "\n" for r in mine[:half]))
        keys = torch.zeros(nparts * cap * W, dtype=torch.int64, device=dev)
        meta = torch.zeros(nparts * cap, dtype=torch.int32, device=dev)
        counts = torch.zeros(nparts, dtype=torch.int64, device=dev)
        # two batches per shard (cut at a read boundary; the device pointer must stay 16-byte aligned)
        cut -= cut % 16
        while blob[cut - 1:cut] != b"\n":
            cut -= 16
            assert cut > 0
        for lo, hi in ((0, cut), (cut, len(blob))):
            shards[p].add_reads_sharded(seq.data_ptr() + lo, hi - lo, nparts, p, cap, keys.data_ptr(), meta.data_ptr(),
                                        counts.data_ptr())
            torch.cuda.synchronize()
            cnt = counts.cpu().tolist()
            assert cnt[p] == 0
            for d in range(nparts):
                if cnt[d]:
                    shards[d].insert_tuples(keys[d * cap * W:].data_ptr(), meta[d * cap:].data_ptr(), cnt[d])
            torch.cuda.synchronize()
        bins.append((keys, meta, counts, seq))
    for p in range(nparts):
        keys, meta, counts, _ = bins[p]
        shards[p].flush_sharded(nparts, p, cap, keys.data_ptr(), meta.data_ptr(), counts.data_ptr())
        torch.cuda.synchronize()
        cnt = counts.cpu().tolist()
        for d in range(nparts):
            if cnt[d]:
                shards[d].insert_tuples(keys[d * cap * W:].data_ptr(), meta[d * cap:].data_ptr(), cnt[d])
        torch.cuda.synchronize()
    rb = 8 * W + 5
    allrecs, loaded, novel, runs = [], 0, 0, []
    for p in range(nparts):
        st = shards[p].sync()
        loaded += st.num_kmers_loaded
        novel += st.num_kmers_novel
        got, n, _ = shards[p].export_records()
        runs.append(bytes(got))
        for i in range(0, len(got), rb):
            key = [int.from_bytes(got[i + 8 * w:i + 8 * w + 8], "little") for w in range(W)]
            assert M.key_owner(key, k, nparts) == p
            allrecs.append(got[i:i + rb])
        shards[p].close()
    assert loaded == ost.num_kmers_loaded and novel == ost.num_kmers_novel

    def sort_key(r):
        return tuple(int.from_bytes(r[8 * w:8 * w + 8], "little") for w in range(W))
    assert b"".join(sorted(allrecs, key=sort_key)) == recs
    # the file a sharded build writes: streaming P-way merge of the shards' sorted exports (multi.merge_sorted_runs)
    from mccortex_b200.multi import merge_sorted_runs
    assert b"".join(x.tobytes() for x in merge_sorted_runs(runs, rb, W, chunk_recs=777)) == recs


@pytest.mark.parametrize("k,hp,cut,nparts", [(31, 0, 43, 2), (21, 4, 45, 3), (63, 6, 40, 2)])
def test_sharded_build_with_quality_and_homopolymer_cutoffs(M, oracle, k, hp, cut, nparts):
    """the production pipeline's flags (--fq-cutoff, also --cut-hp) through the sharded kernels: every shard filters its
    own reads with the quality carry chain (summary pass + sharded insert pass), tuples travel, the union of the shards
    equals the oracle's graph of the same reads and qualities"""
    import torch
    rng = random.Random(17 * nparts + k + cut)
    reads = [r for r in rand_reads(rng, 1200, (30, 260), 9000, perr=0.004, pN=0.002, lower=0.05) if r] + ["A" * 150] * 100
    rng.shuffle(reads)
    quals = ["".join(chr(rng.choice((cut - 1, cut, cut + 1, cut + 20, cut + 20, cut + 20, cut + 20, cut + 20))) for _ in r) for r in reads]
    og = oracle.Graph(k, 1, 1 << 22)
    ost = oracle.Stats()
    for r, q in zip(reads, quals):
        og.add_read(r, qual=q, fq_cutoff=cut, hp_cutoff=hp, stats=ost)
    want = og.dump_sorted()[len(og.header()):]
    og.close()
    dev = torch.device("cuda:0")
    W = (k + 31) // 32
    shards = [M.Graph(k, 1, 1 << 20) for _ in range(nparts)]
    cap = ost.num_kmers_loaded + 1024
    keep = []
    for p in range(nparts):
        seq = _to_dev(torch, "".join(r + "\n" for r in reads[p::nparts]).encode(), dev)
        qual = _to_dev(torch, "".join(q + "\n" for q in quals[p::nparts]).encode(), dev)
        nb = sum(len(r) + 1 for r in reads[p::nparts])
        keys = torch.zeros(nparts * cap * W, dtype=torch.int64, device=dev)
        meta = torch.zeros(nparts * cap, dtype=torch.int32, device=dev)
        counts = torch.zeros(nparts, dtype=torch.int64, device=dev)
        shards[p].add_reads_sharded(seq.data_ptr(), nb, nparts, p, cap, keys.data_ptr(), meta.data_ptr(), counts.data_ptr(),
                                    hp_cutoff=hp, qual_dev_addr=qual.data_ptr(), fq_cutoff=cut)
        torch.cuda.synchronize()
        cnt = counts.cpu().tolist()
        for d in range(nparts):
            if cnt[d]:
                shards[d].insert_tuples(keys[d * cap * W:].data_ptr(), meta[d * cap:].data_ptr(), cnt[d])
        torch.cuda.synchronize()
        shards[p].flush_sharded(nparts, p, cap, keys.data_ptr(), meta.data_ptr(), counts.data_ptr())
        torch.cuda.synchronize()
        cnt = counts.cpu().tolist()
        for d in range(nparts):
            if cnt[d]:
                shards[d].insert_tuples(keys[d * cap * W:].data_ptr(), meta[d * cap:].data_ptr(), cnt[d])
        torch.cuda.synchronize()
        keep.append((seq, qual, keys, meta, counts))
    runs, loaded = [], 0
    for p in range(nparts):
        st = shards[p].sync()
        loaded += st.num_kmers_loaded
        got, n, _ = shards[p].export_records()
        runs.append(bytes(got))
        shards[p].close()
    assert loaded == ost.num_kmers_loaded
    from mccortex_b200.multi import merge_sorted_runs
    assert b"".join(x.tobytes() for x in merge_sorted_runs(runs, 8 * W + 5, W, chunk_recs=500)) == want


@pytest.mark.parametrize("k,nparts", [(31, 2), (31, 4), (63, 3)])
def test_routed_build_on_one_device(M, oracle, k, nparts):
    """the fused compute+exchange path (RoutedBuilder: the sharded kernel appends tuples straight into
    the OWNER's receive ring, the insert kernel reads its count from device memory) with all shards on
    cuda:0 -- rings are handed over as plain pointers instead of CUDA IPC mappings, the counter
    all_to_all is a transpose.  Union of the shards' records == oracle, every shard holds only its keys."""
    import torch
    from mccortex_b200.multi import RoutedBuilder
    rng = random.Random(77 * nparts + k)
    base = rand_reads(rng, 1800, 150, 9000, perr=0.004, pN=0.001, lower=0.02) + ["C" * 150] * 200
    rng.shuffle(base)
    recs, ost = oracle_records(oracle, base, k, capacity=1 << 22)
    dev = torch.device("cuda:0")
    W = (k + 31) // 32
    cap = ost.num_kmers_loaded + 1024
    sbs = [RoutedBuilder(M, None, p, nparts, dev, k, 1 << 20, cap) for p in range(nparts)]
    for sb in sbs:
        sb.connect_local(sbs)

    def exchange():
        torch.cuda.synchronize()
        j = sbs[0].batch % RoutedBuilder.NRING
        for d in range(nparts):
            for s_ in range(nparts):
                sbs[d].rcounts[j][s_] = sbs[s_].counts[j][d]
        torch.cuda.synchronize()

    seqs = []
    for p in range(nparts):
        mine = base[p::nparts]
        third = len(mine) // 3
        parts = [mine[:third], mine[third:2 * third], mine[2 * third:]]   # three batches: ring reuse
        blobs = ["".join(r + "\n" for r in part).encode() for part in parts]
        seqs.append([(_to_dev(torch, b_, dev), len(b_)) for b_ in blobs])
    for b in range(3):
        for p in range(nparts):
            t, n = seqs[p][b]
            sbs[p].produce(t.data_ptr(), n)
        exchange()
        for p in range(nparts):
            sbs[p].consume()
    for p in range(nparts):
        sbs[p].produce_flush()
    exchange()
    for p in range(nparts):
        sbs[p].consume()
    torch.cuda.synchronize()
    rb = 8 * W + 5
    allrecs, loaded, novel = [], 0, 0
    for p in range(nparts):
        st = sbs[p].g.sync()
        loaded += st.num_kmers_loaded
        novel += st.num_kmers_novel
        got, n, _ = sbs[p].g.export_records()
        for i in range(0, len(got), rb):
            key = [int.from_bytes(got[i + 8 * w:i + 8 * w + 8], "little") for w in range(W)]
            assert M.key_owner(key, k, nparts) == p
            allrecs.append(got[i:i + rb])
        sbs[p].close()
    assert loaded == ost.num_kmers_loaded and novel == ost.num_kmers_novel

    def sort_key(r):
        return tuple(int.from_bytes(r[8 * w:8 * w + 8], "little") for w in range(W))
    assert b"".join(sorted(allrecs, key=sort_key)) == recs


def _rand_quals(rng, reads, cut, eqp, lo=35, hi=74):
    quals = ["".join(chr(cut) if rng.random() < eqp else chr(rng.randint(lo, hi)) for _ in r) for r in reads]
    for i in range(0, len(reads), 7):
        quals[i] = quals[i][:len(quals[i]) // 2]      # short quality string
    for i in range(3, len(reads), 11):
        quals[i] = ""                                # no quality
    return quals


@pytest.mark.parametrize("k,hp,cut,eqp", [(21, 0, 43, 0.05), (31, 0, 50, 0.3), (31, 4, 43, 0.9), (63, 5, 60, 0.5), (11, 0, 43, 0.0)])
def test_quality_cutoff_matches_oracle(M, oracle, k, hp, cut, eqp):
    """--fq-cutoff: a contig starts where all k quals are > cutoff and extends while >= cutoff
    (seq_reader.c:84,149); many bases exactly at the cut-off stress the asymmetry"""
    rng = random.Random(k * 100 + cut)
    reads = rand_reads(rng, 400, (1, 500), 6000, perr=0.01, pN=0.005)
    quals = _rand_quals(rng, reads, cut, eqp)
    og = oracle.Graph(k, 1, 1 << 21)
    ost = oracle.Stats()
    for r, q in zip(reads, quals):
        og.add_read(r, qual=q.encode("latin1") if q else None, fq_cutoff=cut, fq_offset=0, hp_cutoff=hp, stats=ost)
    recs = og.dump_sorted()[len(og.header()):]
    for layout in ("offsets", "lines"):
        g = M.Graph(k, 1, 1 << 20)
        if layout == "offsets":
            g.add_reads(reads, hp_cutoff=hp, quals=quals, fq_cutoff=cut)
        else:
            blob = "".join(r + "\n" for r in reads).encode()
            qb = b"".join((q.encode("latin1")[:len(r)] + b"\x7f" * (len(r) - min(len(r), len(q))) + b"!") for r, q in zip(reads, quals))
            g.add_lines(blob, hp_cutoff=hp, qual=qb, fq_cutoff=cut)
        st = g.sync()
        got, _, _ = g.export_records()
        assert got == recs, layout
        assert st.num_kmers_loaded == ost.num_kmers_loaded and st.contigs_parsed == ost.contigs_parsed
        assert st.total_bases_loaded == ost.total_bases_loaded
        g.close()


def test_quality_carry_across_chunks(M, oracle):
    """long reads whose bases sit exactly AT the cut-off: in_contig is carried over many 2 KB chunks"""
    rng = random.Random(99)
    long_reads = rand_reads(rng, 4, 12000, 20000, perr=0, pN=0.0003, lower=0)
    for mode in range(3):
        quals = []
        for r in long_reads:
            q = [chr(50)] * len(r)
            for _ in range(3 if mode else 0):
                s = rng.randrange(0, len(r) - 80)
                for j in range(s, s + 70):
                    q[j] = chr(60)
            if mode == 2:
                for _ in range(5):
                    q[rng.randrange(len(r))] = chr(40)
            quals.append("".join(q))
        og = oracle.Graph(31, 1, 1 << 20)
        ost = oracle.Stats()
        for r, q in zip(long_reads, quals):
            og.add_read(r, qual=q.encode("latin1"), fq_cutoff=50, stats=ost)
        recs = og.dump_sorted()[len(og.header()):]
        g = M.Graph(31, 1, 1 << 20)
        g.add_reads(long_reads, quals=quals, fq_cutoff=50)
        st = g.sync()
        got, _, _ = g.export_records()
        assert got == recs
        assert st.num_kmers_loaded == ost.num_kmers_loaded and st.contigs_parsed == ost.contigs_parsed
        g.close()


def test_high_multiplicity_kmers_drain_the_front_table(M, oracle):
    """k-mers seen far more often than the front table's count field holds (poly-A, a short
    tandem repeat): counts are drained to the big table on the fly and nothing is lost"""
    rng = random.Random(4)
    reads = ["A" * 150] * 40000 + ["ACGTTGCA" * 20] * 2000 + rand_reads(rng, 500, 150, 5000, perr=0.0, pN=0.0, lower=0.0)
    rng.shuffle(reads)
    recs, ost = oracle_records(oracle, reads, 31)
    g = M.Graph(31, 1, 1 << 18)
    blob = "".join(r + "\n" for r in reads).encode()
    for _ in range(3):
        g.add_lines(blob)
    st = g.sync()
    assert st.num_kmers_loaded == 3 * ost.num_kmers_loaded and st.num_kmers_novel == ost.num_kmers_novel
    got, n, rb = g.export_records()
    assert n == ost.num_kmers_novel
    for i in range(0, len(recs), rb):
        assert got[i:i + 8] == recs[i:i + 8]
        assert int.from_bytes(got[i + 8:i + 12], "little") == 3 * int.from_bytes(recs[i + 8:i + 12], "little")
        assert got[i + 12] == recs[i + 12]
    assert max(int.from_bytes(got[i + 8:i + 12], "little") for i in range(0, len(got), rb)) >= 3 * 40000 * 120
    g.close()


@pytest.mark.parametrize("k", [21, 47])
def test_load_records_and_intersect_match_oracle(M, oracle, tmp_path, k):
    """graph_load() through the C ABI (mcx_graph_load_records): a 3-colour graph file goes into a 2-colour
    graph through the colour filter 2->1, 0->1 on top of reads; then the same file as an INTERSECTION
    graph: must-exist graph load + must-exist reads + finish.  Both against the oracle's restatement
    (pinned to the compiled reference in tests/test_oracle.py)."""
    rng = random.Random(500 + k)
    shared = rand_reads(rng, 900, (20, 240), 7000, perr=0.006)
    parts = [shared[0:300], shared[300:600], shared[600:900]]
    src = oracle.Graph(k, 3, 1 << 20)
    for c, reads in enumerate(parts):
        for r in reads:
            src.add_read(r, colour=c)
    path = tmp_path / "three.ctx"
    path.write_bytes(src.dump_sorted())
    src.close()
    extra = shared[250:700]

    # ---- build --graph 1,1:three.ctx:2,0 on top of reads in colour 0
    og = oracle.Graph(k, 2, 1 << 20)
    for r in extra:
        og.add_read(r, colour=0)
    ctx, oloaded, onovel = og.load_ctx("1,1:%s:2,0" % path)
    want = og.dump_sorted()[len(og.header()):]
    og.close()
    g = M.Graph(k, 2, 1 << 20)
    g.add_lines("".join(r + "\n" for r in extra).encode(), colour=0)
    g.sync()
    loaded, novel = g.load_records(ctx.records, ctx.ncols, [f for f, _ in ctx.filter], [t for _, t in ctx.filter])
    assert (loaded, novel) == (oloaded, onovel)
    g.sync()
    got, n, _ = g.export_records()
    assert got == want
    g.close()

    # ---- build --intersect three.ctx:1 --graph three.ctx:0,2 + reads
    og = oracle.Graph(k, 3, 1 << 20)
    og.set_intersect(True)
    ictx, _, _ = og.load_ctx("%s:1" % path, 0, isec=True)
    gctx, gl, _ = og.load_ctx("%s:0,2" % path, 0, must_exist=True, mask_isec=True)
    ost = oracle.Stats()
    for r in extra:
        og.add_read(r, colour=2, stats=ost)
    og.finish_intersect()
    want = og.dump_sorted()[len(og.header()):]
    og.close()
    g = M.Graph(k, 3, 1 << 20, flags=M.MCX_GRAPH_INTERSECT)
    g.load_records(ictx.records, ictx.ncols, [f for f, _ in ictx.filter], [0] * len(ictx.filter), M.MCX_LOAD_INTO_ISEC)
    l2, n2 = g.load_records(gctx.records, gctx.ncols, [f for f, _ in gctx.filter], [t for _, t in gctx.filter],
                            M.MCX_LOAD_MUST_EXIST | M.MCX_LOAD_MASK_ISEC)
    assert (l2, n2) == (gl, 0)
    g.sync()
    g.add_lines("".join(r + "\n" for r in extra).encode(), colour=2, must_exist=True)
    st = g.sync()
    assert st.num_kmers_loaded == ost.num_kmers_loaded and st.num_kmers_novel == 0
    assert st.contigs_parsed == ost.contigs_parsed and st.total_bases_loaded == ost.total_bases_loaded
    kept = g.finish_intersect()
    got, n, _ = g.export_records()
    assert n == kept and got == want
    g.close()


# ---- build --remove-pcr (row N3) -------------------------------------------------------------
@pytest.mark.parametrize("k,hp,cut,matedir,nbatch", [(21, 0, 0, 1, 1), (31, 0, 0, 3, 4), (21, 4, 45, 1, 3), (33, 0, 50, 2, 2),
                                                   (63, 5, 0, 0, 5), (11, 0, 40, 1, 7)])
def test_remove_pcr_matches_oracle(M, oracle, k, hp, cut, matedir, nbatch):
    """mcx_graph_add_reads_pcr (orient / mark / mask kernels + the normal build) against seq_reads_are_novel restated
    read by read in the oracle (itself byte-identical to the reference run with one thread)"""
    from conftest import pcr_reads
    rng = random.Random(1000 * k + cut + matedir)
    units = pcr_reads(rng, 3000, k, G=20000, sites=300)
    quals = [tuple("".join(chr(cut) if rng.random() < 0.03 else chr(rng.randint(max(cut - 3, 35), 74)) for _ in r) if cut else ""
                   for r in u) for u in units]
    og = oracle.Graph(k, 1, 1 << 20)
    ost = oracle.Stats()
    for u, q in zip(units, quals):
        og.add_reads_pcr(u[0], q[0] or None, u[1] if len(u) > 1 else None, (q[1] or None) if len(u) > 1 else None,
                         fq_cutoff=cut, hp_cutoff=hp, matedir=matedir, stats=ost)
    recs = og.dump_sorted()[len(og.header()):]
    g = M.Graph(k, 1, 1 << 20, flags=M.MCX_GRAPH_READSTRT)
    step = (len(units) + nbatch - 1) // nbatch
    for i in range(0, len(units), step):
        g.add_units_pcr(units[i:i + step], quals[i:i + step], fq_cutoff=cut, hp_cutoff=hp, matedir=matedir)
    st = g.sync()
    got, n, _ = g.export_records()
    assert got == recs
    assert (st.num_dup_se_reads, st.num_dup_pe_pairs) == (ost.num_dup_se_reads, ost.num_dup_pe_pairs)
    assert ost.num_dup_se_reads + ost.num_dup_pe_pairs > 300
    assert st.num_kmers_loaded == ost.num_kmers_loaded and st.contigs_parsed == ost.contigs_parsed
    # (the reference's num_kmers_novel also counts one per mate WITHOUT a k-mer, build_graph.c:75-76 -- a log-only
    # statistic; the library reports the slots claimed)
    assert st.num_kmers_novel == n
    assert st.total_bases_read == ost.total_bases_read and st.total_bases_loaded == ost.total_bases_loaded
    g.close()


def test_remove_pcr_reset_between_colours(M, oracle):
    """the read-start marks are wiped when the colour changes (ctx_build.c:392-395)"""
    from conftest import pcr_reads
    rng = random.Random(77)
    units = pcr_reads(rng, 1500, 31, paired=0.3)
    og = oracle.Graph(31, 2, 1 << 20)
    g = M.Graph(31, 2, 1 << 20, flags=M.MCX_GRAPH_READSTRT)
    for col in (0, 1):
        if col:
            og.wipe_readstrt()
            g.pcr_reset()
        for u in units:
            og.add_reads_pcr(u[0], None, u[1] if len(u) > 1 else None, None, colour=col)
        g.add_units_pcr(units, colour=col)
    g.sync()
    got, _, _ = g.export_records()
    assert got == og.dump_sorted()[len(og.header()):]
    g.close()


# ---- round-1 advisor findings ------------------------------------------------------------------------------------
def test_wide_records_export(M, oracle):
    """hundreds of colours (a `join` of many samples): the export formats records of any width (the first version
    needed 128 x record bytes of shared memory and failed above ~360 colours)"""
    rng = random.Random(41)
    ncols, k = 600, 31
    base = rand_reads(rng, 60, 150, 2000)
    og = oracle.Graph(k, ncols, 1 << 16)
    g = M.Graph(k, ncols, 1 << 16)
    for c in (0, 1, 255, 256, 361, 598, 599):
        reads = base[:20] + rand_reads(rng, 20, 150, 2000)
        for r in reads:
            og.add_read(r, colour=c)
        g.add_lines("".join(r + "\n" for r in reads).encode(), colour=c)
    g.sync()
    want = og.dump_sorted()[len(og.header()):]
    og.close()
    for srt in (True, False):
        got, n, rb = g.export_records(sorted=srt)
        assert rb == 8 + 5 * ncols and n * rb == len(want)
        if srt:
            assert got == want
        else:
            assert sorted(got[i:i + rb] for i in range(0, len(got), rb)) == sorted(want[i:i + rb] for i in range(0, len(want), rb))
    g.close()


def test_loaded_coverage_saturates(M):
    """graph_load adds FILE coverages (db_node.c:139-144 saturates at UINT32_MAX): 0xFFFFFFFF on top of 5 must stay
    0xFFFFFFFF, in either order, and 0xF0000000 + 0x20000000 must saturate too"""
    import struct
    k = 31
    key = 0x0123456789ABCDE  # any canonical-looking 62-bit value: the table does not care
    def rec(covg, edges=0x11, kk=key):
        return struct.pack("<QIB", kk, covg, edges)
    for first, second, want in ((5, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 5, 0xFFFFFFFF), (0xF0000000, 0x20000000, 0xFFFFFFFF),
                                (0x7FFFFFFF, 0x7FFFFFFF, 0xFFFFFFFE), (7, 9, 16)):
        g = M.Graph(k, 1, 1 << 10)
        g.load_records(rec(first) + rec(3, 0x02, key + 1), 1, [0], [0])
        g.load_records(rec(second), 1, [0], [0])
        got, n, rb = g.export_records()
        assert n == 2 and rb == 13
        assert struct.unpack("<QIB", got[:13]) == (key, want, 0x11)
        assert struct.unpack("<QIB", got[13:]) == (key + 1, 3, 0x02)
        g.close()


def test_undersized_table_fails_fast(M):
    """an -n far too small for the input must report "Hash table is full" quickly (the reference dies at once,
    hash_table.c:119-123,280), not scan the whole table for every missing k-mer"""
    import time
    rng = random.Random(5)
    reads = rand_reads(rng, 60000, 150, 5_000_000, perr=0.0, pN=0.0, lower=0.0)   # ~4.5 M distinct k-mers
    blob = "".join(r + "\n" for r in reads).encode()
    for k in (31, 63):
        g = M.Graph(k, 1, 1 << 18)
        t0 = time.time()
        g.add_lines(blob)
        with pytest.raises(M.McxError) as e:
            g.sync()
        assert e.value.status == "MCX_ERR_TABLE_FULL"
        assert time.time() - t0 < 20.0
        g.close()


def test_sort_records_compares_whole_words(M):
    """`sort` orders records by their whole 64-bit key words whatever k says (ctx_sort.c:117-155): keys with bits
    above 2k (a damaged header k) must come out where the reference's compare puts them"""
    import struct
    rng = random.Random(17)
    for k, W in ((11, 1), (31, 1), (41, 2)):
        recs = []
        for _ in range(3000):
            words = [rng.getrandbits(64) for _ in range(W)]
            recs.append(b"".join(struct.pack("<Q", w) for w in words) + struct.pack("<IB", rng.getrandbits(32), rng.getrandbits(8)))
        blob = b"".join(recs)
        got = M.sort_records(k, 1, blob)
        rb = 8 * W + 5
        want = sorted(recs, key=lambda r: struct.unpack("<%dQ" % W, r[:8 * W]))
        assert [got[i:i + rb] for i in range(0, len(got), rb)] == want
