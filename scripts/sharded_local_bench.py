"""the sharded build kernel on ONE GPU with local bins (no NVLink): what does routing cost by itself?"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mccortex_b200 as M
import bench as B
R = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
SL = B.synth_lib()
genome = C.create_string_buffer(B.GENOME); SL.mcx_synth_genome(genome, B.GENOME, 0)
stride = B.READ_LEN + 1; nbytes = R * stride
host = M.host_alloc(nbytes + 4096)
SL.mcx_synth_reads(host, 0, R, B.READ_LEN, genome, B.GENOME, B.P_ERR, 0, 0)
dseq = torch.empty(nbytes + 4096, dtype=torch.uint8, device=dev)
dseq[:nbytes].copy_(torch.frombuffer((C.c_uint8 * nbytes).from_address(host), dtype=torch.uint8))
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
occ = R * (B.READ_LEN - B.K + 1)
cap = int((B.GENOME + R * B.READ_LEN * B.P_ERR * B.K * 1.05) / 0.75)
g = M.Graph(B.K, 1, cap); g.set_stream(stream.cuda_stream)
cap_part = occ // 8
keys = torch.empty(P * cap_part, dtype=torch.int64, device=dev)
meta = torch.empty(P * cap_part, dtype=torch.int32, device=dev)
counts = torch.zeros(P, dtype=torch.int64, device=dev)
for label in ("fused", "sharded(local bins, %d parts)" % P):
    best = 1e9
    for it in range(3):
        g.clear(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if label == "fused":
            g.add_reads_raw(dseq.data_ptr(), nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE)
        else:
            g.add_reads_sharded(dseq.data_ptr(), nbytes, P, 0, cap_part, keys.data_ptr(), meta.data_ptr(), counts.data_ptr())
        e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    extra = ""
    if label != "fused":
        extra = " tuples %s" % counts.tolist()
        g.flush_sharded(P, 0, cap_part, keys.data_ptr(), meta.data_ptr(), counts.data_ptr())
    print("%-34s %d reads: %.2f ms%s" % (label, R, best, extra), flush=True)
    g.sync()
