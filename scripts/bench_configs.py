"""kernel-only throughput of the other single-GPU configs of BASELINE.json (not bench lines: parity cases whose speed is
worth knowing): configs[2] k=63 and configs[3] four colours, device-resident reads, CUDA events.
usage: python scripts/bench_configs.py [reads]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mccortex_b200 as M
import bench as B
R = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
dev = torch.device("cuda:0")
SL = B.synth_lib()
genome = C.create_string_buffer(B.GENOME); SL.mcx_synth_genome(genome, B.GENOME, 0)
stride = B.READ_LEN + 1; nbytes = R * stride
host = M.host_alloc(nbytes + 4096)
SL.mcx_synth_reads(host, 0, R, B.READ_LEN, genome, B.GENOME, B.P_ERR, 0, 0)
dseq = torch.empty(nbytes + 4096, dtype=torch.uint8, device=dev)
dseq[:nbytes].copy_(torch.frombuffer((C.c_uint8 * nbytes).from_address(host), dtype=torch.uint8))
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
def run(k, ncols, label):
    occ = R * (B.READ_LEN - k + 1)
    cap = int((B.GENOME + R * B.READ_LEN * B.P_ERR * k * 1.05) / 0.75)
    g = M.Graph(k, ncols, cap); g.set_stream(stream.cuda_stream)
    per = nbytes // ncols // stride * stride
    best = 1e9
    for it in range(3):
        g.clear(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for c in range(ncols):
            g.add_reads_raw(dseq.data_ptr() + c * per, per if c + 1 < ncols else nbytes - c * per, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE, colour=c)
        g.flush()
        e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = g.sync()
    assert st.num_kmers_loaded == occ, (st.num_kmers_loaded, occ)
    print("%-34s %d reads: %.1f ms, %.2f G k-mers/s (distinct %d)" % (label, R, best, occ / best / 1e6, g.stats()[0]), flush=True)
    g.close()
run(31, 1, "configs[1] k=31, 1 colour")
run(63, 1, "configs[2] k=63, 1 colour")
run(31, 4, "configs[3] k=31, 4 colours")
run(21, 1, "k=21, 1 colour")
