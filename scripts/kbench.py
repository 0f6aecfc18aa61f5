"""kernel-only A/B of the build kernels on the bench workload (configs[1] shape), device-resident reads, CUDA events.
usage: python scripts/kbench.py [reads] [variant ...]      variant = name:ENV=VAL,ENV=VAL
Every variant must produce the same sorted records (md5 printed) -- parity across kernels at scale."""
import ctypes as C, hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mccortex_b200 as M
import bench as B

R = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
variants = sys.argv[2:] or ["fused:MCX_KERNEL=fused", "warp:MCX_KERNEL=warp", "warp2x20:MCX_KERNEL=warp,MCX_CLASSES=2,MCX_FRONT_BITS=20",
                            "warp2x21:MCX_KERNEL=warp,MCX_CLASSES=2,MCX_FRONT_BITS=21", "warp4x19:MCX_KERNEL=warp,MCX_CLASSES=4,MCX_FRONT_BITS=19"]
K = int(os.environ.get("KBENCH_K", 31))
MD5 = os.environ.get("KBENCH_MD5", "1") != "0"
QUAL = float(os.environ.get("KBENCH_QUAL", "0"))   # fraction of bases at or below the --fq-cutoff 10 threshold; 0 = no quality
dev = torch.device("cuda:0")
SL = B.synth_lib()
genome = C.create_string_buffer(B.GENOME); SL.mcx_synth_genome(genome, B.GENOME, 0)
stride = B.READ_LEN + 1; nbytes = R * stride
host = M.host_alloc(nbytes + 4096)
SL.mcx_synth_reads(host, 0, R, B.READ_LEN, genome, B.GENOME, B.P_ERR, 0, 0)
dseq = torch.empty(nbytes + 4096, dtype=torch.uint8, device=dev)
dseq[:nbytes].copy_(torch.frombuffer((C.c_uint8 * nbytes).from_address(host), dtype=torch.uint8))
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
qual = None
if QUAL > 0:
    torch.manual_seed(12345)
    qual = torch.full((nbytes + 4096,), 70, dtype=torch.uint8, device=dev)
    qual[:nbytes][torch.rand(nbytes, device=dev) < QUAL] = 40
    torch.cuda.synchronize()
qstep = (1 << 30) // (16 * stride) * (16 * stride)
occ = R * (B.READ_LEN - K + 1)
cap = int((B.GENOME + R * B.READ_LEN * B.P_ERR * K * 1.05) / 0.75)
ref = None
for v in variants:
    name, _, envs = v.partition(":")
    keys = []
    for kv in filter(None, envs.split(",")):
        a, _, b = kv.partition("="); os.environ[a] = b; keys.append(a)
    g = M.Graph(K, 1, cap); g.set_stream(stream.cuda_stream)
    times = []
    for it in range(4):
        g.clear(); torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        if qual is None:
            g.add_reads_raw(dseq.data_ptr(), nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE)
        else:
            for lo in range(0, nbytes, qstep):
                g.add_reads_raw(dseq.data_ptr() + lo, min(qstep, nbytes - lo), M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE,
                                qual_addr=qual.data_ptr() + lo, fq_cutoff=43)
        e1.record(stream)
        g.flush()
        e2.record(stream); torch.cuda.synchronize()
        times.append((e0.elapsed_time(e2), e0.elapsed_time(e1)))
    st = g.sync()
    assert qual is not None or st.num_kmers_loaded == occ, (name, st.num_kmers_loaded, occ)
    if qual is not None: occ = st.num_kmers_loaded
    best = min(times[1:])
    line = "%-10s %d reads k=%d: %.2f ms (kernel %.2f) %.2f G k-mers/s frac %.3f  reads=%d contigs=%d distinct=%d" % (
        name, R, K, best[0], best[1], occ / best[0] / 1e6, occ * (19.25 if K <= 31 else 27.7) / best[0] / 1e6 / 6537.0,
        st.num_se_reads, st.contigs_parsed, g.stats()[0])
    if MD5:
        recs, n, rb = g.export_records(sorted=True)
        h = hashlib.md5(recs).hexdigest()
        line += " records=%d md5=%s" % (n, h)
        if ref is None: ref = h
        line += " PARITY-OK" if h == ref else " PARITY-MISMATCH"
    print(line, flush=True)
    g.close()
    for a in keys: os.environ.pop(a, None)
