"""GPU tests at BASELINE.json's full single-GPU sizes (configs 2, 3 and 4), where the oracle would take
minutes to hours: the CUDA path is checked through size-independent properties of the domain,

  * every occurrence is counted:        sum of all coverages == num_kmers_loaded == reads x (L - k + 1)
  * the export is a sorted SET:         keys strictly increasing, record count == num_kmers_novel
  * keys are canonical k-mers:          key < 4^k and key <= revcomp(key)            (sampled)
  * edges are reciprocal:               a k-mer with an outgoing edge to base b has the neighbour
                                        (k-1 suffix + b) in the graph                (sampled)
  * linearity / idempotence:            loading the same reads again doubles every coverage and leaves
                                        keys and edges unchanged (checksums)
  * colours are independent:            colour c of the 4-colour graph == the 1-colour graph of sample c
                                        (checksum of checksums)
  * a prefix of the workload, small enough for the oracle, is bit-exact (records and counters)

Synthetic reads are the bench workload (mccortex_b200/tools/mcx_synth.c)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, oracle_records

pytestmark = pytest.mark.gpu

GENOME, READ_LEN, P_ERR = 4_600_000, 150, 0.001
M64 = (1 << 64) - 1


@pytest.fixture(scope="module")
def M():
    import mccortex_b200 as M
    assert M.device_count() > 0, "GPU tests need a CUDA device"
    return M


@pytest.fixture(scope="module")
def synth():
    L = C.CDLL(os.path.join(ROOT, "mccortex_b200", "lib", "libmcxsynth.so"))
    L.mcx_synth_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    L.mcx_synth_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64,
                                  C.c_double, C.c_int, C.c_uint64]
    genome = C.create_string_buffer(GENOME)
    L.mcx_synth_genome(genome, GENOME, 0)
    return L, genome


def _device_reads(M, synth, first, nreads, seed_xor=0):
    """reads [first, first + nreads) of the bench workload as a device tensor in LINES layout"""
    import torch
    SL, genome = synth
    nbytes = nreads * (READ_LEN + 1)
    host = M.host_alloc(nbytes + 4096)
    SL.mcx_synth_reads(host, first, nreads, READ_LEN, genome, GENOME, P_ERR, 0, seed_xor)
    d = torch.empty(nbytes + 4096, dtype=torch.uint8, device="cuda:0")
    d[:nbytes].copy_(torch.frombuffer((C.c_uint8 * nbytes).from_address(host), dtype=torch.uint8))
    torch.cuda.synchronize()
    return d, nbytes, host


def _records(M, g, W, ncols, chunk_recs=8_000_000):
    """yield the sorted export as numpy structured arrays, chunk by chunk"""
    n, rb = C.c_uint64(), C.c_uint32()
    M.binding._ck(M.lib().mcx_graph_export_begin(g.h, 1, C.byref(n), C.byref(rb)), "export_begin")
    dt = np.dtype([("key", "<u8", (W,)), ("covg", "<u4", (ncols,)), ("edges", "u1", (ncols,))])
    assert dt.itemsize == rb.value == 8 * W + 5 * ncols
    try:
        buf = np.empty(chunk_recs, dtype=dt)
        at = 0
        while at < n.value:
            m = min(chunk_recs, n.value - at)
            M.binding._ck(M.lib().mcx_graph_export_read(g.h, at, m, buf.ctypes.data_as(C.c_void_p)), "export_read")
            yield at, buf[:m], n.value
            at += m
    finally:
        M.lib().mcx_graph_export_end(g.h)


def _revcomp(key_hi, key_lo, k):
    """reverse complement of k-mers held as (hi, lo) uint64 arrays (hi = 0 for k <= 32), numpy"""
    def rc64(x):
        x = ~x
        x = ((x >> np.uint64(2)) & np.uint64(0x3333333333333333)) | ((x & np.uint64(0x3333333333333333)) << np.uint64(2))
        x = ((x >> np.uint64(4)) & np.uint64(0x0F0F0F0F0F0F0F0F)) | ((x & np.uint64(0x0F0F0F0F0F0F0F0F)) << np.uint64(4))
        return x.byteswap()
    if k <= 32:
        return np.zeros_like(key_lo), rc64(key_lo) >> np.uint64(64 - 2 * k)
    s = np.uint64(128 - 2 * k)
    hi, lo = rc64(key_lo), rc64(key_hi)
    return hi >> s, (hi << (np.uint64(64) - s)) | (lo >> s)


def _scan(M, g, k, ncols, sample_every=997):
    """one pass over the sorted export: sortedness, per-colour coverage sums, checksums, canonical
    sample; returns a dict and a sample of (key words, edges) for the reciprocity check"""
    W = (k + 31) // 32
    tot = np.zeros(ncols, dtype=np.uint64)
    nz = np.zeros(ncols, dtype=np.uint64)
    key_ck = np.uint64(0)
    col_ck = [np.uint64(0)] * ncols
    prev = None
    nrec = 0
    sample = []
    mul = np.uint64(0x9E3779B97F4A7C15)
    for at, r, n in _records(M, g, W, ncols):
        key = r["key"]
        hi = key[:, 0] if W == 2 else np.zeros(len(r), dtype=np.uint64)
        lo = key[:, W - 1]
        # strictly increasing (hi, lo), also across chunk borders
        if W == 1:
            assert np.all(lo[1:] > lo[:-1])
        else:
            assert np.all((hi[1:] > hi[:-1]) | ((hi[1:] == hi[:-1]) & (lo[1:] > lo[:-1])))
        if prev is not None:
            assert (prev[0], prev[1]) < (int(hi[0]), int(lo[0]))
        prev = (int(hi[-1]), int(lo[-1]))
        assert int(hi.max()) < (1 << max(0, 2 * k - 64)) or W == 1
        if W == 1:
            assert int(lo.max()) < (1 << (2 * k))
        tot += r["covg"].sum(axis=0, dtype=np.uint64)
        nz += (r["covg"] > 0).sum(axis=0).astype(np.uint64)
        with np.errstate(over="ignore"):
            h = (lo * mul) ^ (hi * np.uint64(0xC2B2AE3D27D4EB4F))
            key_ck ^= np.bitwise_xor.reduce(h ^ (r["edges"].astype(np.uint64).sum(axis=1) << np.uint64(7)))
            for c in range(ncols):
                present = r["covg"][:, c] > 0
                hc = (h[present] * np.uint64(31) + r["covg"][present, c].astype(np.uint64)) * mul + r["edges"][present, c].astype(np.uint64)
                col_ck[c] = col_ck[c] + hc.sum(dtype=np.uint64)
        s = slice(at % sample_every and sample_every - at % sample_every, None, sample_every)
        rh, rl = _revcomp(hi[s], lo[s], k)
        assert np.all((hi[s] < rh) | ((hi[s] == rh) & (lo[s] < rl))), "non-canonical key in the export"
        sample.append((hi[s].copy(), lo[s].copy(), r["edges"][s].copy()))
        nrec = n
    return {"nrec": nrec, "covg_sum": tot, "present": nz, "key_ck": int(key_ck), "col_ck": [int(x) for x in col_ck]}, sample


def _check_reciprocity(M, g, k, ncols, sample, max_keys=20000):
    """for sampled (key, edges): every outgoing edge (union over colours) leads to a k-mer of the graph.
    The neighbours are looked up by inserting them into a scratch 1-colour graph?  No: membership is
    tested against the sorted export itself with searchsorted (k <= 31 only, where a key is one word)."""
    if k > 31:
        return
    W = 1
    keys = np.concatenate([lo for _, lo, _ in sample])[:max_keys]
    edges = np.concatenate([np.bitwise_or.reduce(e, axis=1) for _, _, e in sample])[:max_keys]
    mask = np.uint64((1 << (2 * k)) - 1)
    want = []
    for b in range(4):
        fw = ((keys << np.uint64(2)) | np.uint64(b)) & mask                      # next k-mer on the forward strand
        _, rfw = _revcomp(np.zeros_like(fw), fw, k)
        want.append(np.minimum(fw, rfw)[(edges >> b) & 1 == 1])
        _, rk = _revcomp(np.zeros_like(keys), keys, k)                            # forward edges of the reverse strand
        rv = ((rk << np.uint64(2)) | np.uint64(b)) & mask
        _, rrv = _revcomp(np.zeros_like(rv), rv, k)
        want.append(np.minimum(rv, rrv)[(edges >> (4 + b)) & 1 == 1])
    want = np.unique(np.concatenate(want))
    found = np.zeros(len(want), dtype=bool)
    for at, r, n in _records(M, g, W, ncols):
        lo = r["key"][:, 0]
        i = np.searchsorted(lo, want)
        i[i >= len(lo)] = len(lo) - 1
        found |= lo[i] == want
    assert found.all(), "%d of %d edge targets are missing from the graph" % ((~found).sum(), len(want))


def _capacity(nreads_total, k):
    return int((GENOME + nreads_total * READ_LEN * P_ERR * k * 1.05) / 0.75)


@pytest.mark.parametrize("k", [31, 63])
def test_config2_config3_full_size(M, synth, oracle, k):
    """configs[1] (k=31) and configs[2] (k=63): 50 M x 150 bp reads, one colour, one GPU"""
    R = 50_000_000
    nk = READ_LEN - k + 1
    d, nbytes, host = _device_reads(M, synth, 0, R)
    g = M.Graph(k, 1, _capacity(R, k))
    g.add_reads_raw(d.data_ptr(), nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE)
    st = g.sync()
    assert st.num_kmers_loaded == R * nk and st.num_se_reads == R and st.contigs_parsed == R
    assert st.total_bases_loaded == R * READ_LEN
    one, sample = _scan(M, g, k, 1)
    assert one["nrec"] == st.num_kmers_novel == g.stats()[0]
    assert int(one["covg_sum"][0]) == R * nk
    assert one["nrec"] >= GENOME - k + 1 - 64          # (almost) every genomic k-mer at ~1600x coverage
    _check_reciprocity(M, g, k, 1, sample)
    # the same reads again: coverage doubles, keys and edges do not move; host-buffer path this time
    g.add_reads_raw(host, nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_HOST)
    st2 = g.sync()
    assert st2.num_kmers_loaded == R * nk and st2.num_kmers_novel == 0
    assert g.stats()[0] == one["nrec"]
    if k == 31:   # (k=63 has 3.5x the records: one scan is enough there)
        two, _ = _scan(M, g, k, 1)
        assert two["nrec"] == one["nrec"] and two["key_ck"] == one["key_ck"]
        assert int(two["covg_sum"][0]) == 2 * R * nk
    g.close()
    # a prefix the oracle can do: bit-exact
    n_small = 40_000
    SL, genome = synth
    buf = C.create_string_buffer(n_small * (READ_LEN + 1))
    SL.mcx_synth_reads(buf, 0, n_small, READ_LEN, genome, GENOME, P_ERR, 0, 0)
    reads = buf.raw.decode().split("\n")[:n_small]
    recs, ost = oracle_records(oracle, reads, k, capacity=1 << 24)
    gs = M.Graph(k, 1, 1 << 23)
    gs.add_reads_raw(d.data_ptr(), n_small * (READ_LEN + 1), M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE)
    sts = gs.sync()
    got, n, _ = gs.export_records()
    assert got == recs and sts.num_kmers_novel == ost.num_kmers_novel and sts.num_kmers_loaded == ost.num_kmers_loaded
    gs.close()
    M.host_free(host)


def test_config4_four_colours(M, synth):
    """configs[3]: 4 samples x 25 M x 150 bp reads, k=31, per-colour coverage and edges, one GPU.
    Samples are disjoint read-index ranges of the workload (the per-sample SNP seed of SURVEY 8d is
    the generator's seed_xor for the errors)."""
    k, R, C4 = 31, 25_000_000, 4
    nk = READ_LEN - k + 1
    g = M.Graph(k, C4, _capacity(C4 * R, k))
    keep = []
    for c in range(C4):
        d, nbytes, host = _device_reads(M, synth, c * R, R)
        g.add_reads_raw(d.data_ptr(), nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE, colour=c)
        st = g.sync()
        assert st.num_kmers_loaded == R * nk
        M.host_free(host)
        keep.append((d, nbytes) if c == 2 else None)
    four, sample = _scan(M, g, k, C4)
    assert [int(x) for x in four["covg_sum"]] == [R * nk] * C4
    assert four["nrec"] == g.stats()[0]
    _check_reciprocity(M, g, k, C4, sample)
    g.close()
    # colours are independent: colour 2 of the joint graph == the graph of sample 2 alone
    c = 2
    d, nbytes = keep[c]
    g1 = M.Graph(k, 1, _capacity(R, k))
    g1.add_reads_raw(d.data_ptr(), nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE)
    st = g1.sync()
    one, _ = _scan(M, g1, k, 1)
    assert one["nrec"] == int(four["present"][c]) == st.num_kmers_novel
    assert one["col_ck"][0] == four["col_ck"][c]
    g1.close()


# ---- byte parity with the reference at the stated sizes (`cmp` on configs 1-4) -------------------------
# tests/golden/fullsize_md5.json holds md5 / size of `oracle/_ref/mccortex31|63 build -S` outputs on the stated inputs,
# made ONCE on a GPU box by scripts/gpu_fullsize_golden.sh (which also `cmp`s them with this driver's files there; the
# reference needs 3-7 minutes of 16 cores per config).  Here the driver's `build -S` output must have the same md5.
def _golden_fullsize():
    import json
    p = os.path.join(ROOT, "tests", "golden", "fullsize_md5.json")
    return json.load(open(p)) if os.path.exists(p) else {}


@pytest.mark.parametrize("name", ["config1a_1Mbase_record_80col", "config1b_6536_reads", "config2_50M_reads_k31",
                                  "config3_50M_reads_k63", "config4_4x25M_reads_4_colours"])
def test_cli_output_md5_equals_reference_at_full_size(name, tmp_path):
    import hashlib
    import shutil
    import subprocess
    gold = _golden_fullsize().get(name)
    if not gold or gold.get("cmp") != "identical":
        pytest.skip("no reference md5 recorded for %s" % name)
    import mccortex_b200 as M
    synth_bin = os.path.join(ROOT, "mccortex_b200", "bin", "mcx-synth")
    need = {"config2_50M_reads_k31": 10e9, "config3_50M_reads_k63": 14e9, "config4_4x25M_reads_4_colours": 20e9}.get(name, 1e8)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need * 1.2 else str(tmp_path)
    d = os.path.join(base, "mcx_full_%d" % os.getpid())
    os.makedirs(d, exist_ok=True)
    try:
        def synth(path, G, first, n):
            with open(path, "wb") as f:
                subprocess.run([synth_bin, str(G), str(first), str(n), "150", "0.001", "1"], stdout=f, check=True)
        if name.startswith("config1a"):
            SL = C.CDLL(os.path.join(ROOT, "mccortex_b200", "lib", "libmcxsynth.so"))
            SL.mcx_synth_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
            g = C.create_string_buffer(1000000); SL.mcx_synth_genome(g, 1000000, 0x1A)
            s = "".join("ACGT"[b] for b in g.raw[:1000000])
            open(os.path.join(d, "c1a.fa"), "w").write(">one\n" + "\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n")
            inputs = ["--sample", "s", "--seq", os.path.join(d, "c1a.fa")]
        elif name.startswith("config1b"):
            synth(os.path.join(d, "c1b.fa"), 100000, 0, 6536)
            inputs = ["--sample", "s", "--seq", os.path.join(d, "c1b.fa")]
        elif name.startswith("config4"):
            inputs = []
            for c in range(4):
                synth(os.path.join(d, "s%d.fa" % c), GENOME, c * 25_000_000, 25_000_000)
                inputs += ["--sample", "s%d" % c, "--seq", os.path.join(d, "s%d.fa" % c)]
        else:
            synth(os.path.join(d, "r.fa"), GENOME, 0, 50_000_000)
            inputs = ["--sample", "s", "--seq", os.path.join(d, "r.fa")]
        out = os.path.join(d, "out.ctx")
        r = subprocess.run([M.driver_path(), "build", "-f", "-q", "-m", "100G", "-n", str(gold["nslots"]), "-k", str(gold["k"]), "-S"] + inputs + [out],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr[-2000:]
        assert os.path.getsize(out) == gold["bytes"]
        h = hashlib.md5()
        with open(out, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        assert h.hexdigest() == gold["md5"], "the .ctx differs from the reference's (%s)" % gold["args"]
    finally:
        shutil.rmtree(d, ignore_errors=True)
