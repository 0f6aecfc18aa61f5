"""one-GPU timing of the sharded kernel (nparts=2, bins local) next to the fused kernel on the same reads"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mccortex_b200 as M
import bench as B
R = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nparts = int(os.environ.get("NPARTS", 2))
dev = torch.device("cuda:0")
SL = B.synth_lib()
genome = C.create_string_buffer(B.GENOME); SL.mcx_synth_genome(genome, B.GENOME, 0)
stride = B.READ_LEN + 1; nbytes = R * stride
host = M.host_alloc(nbytes + 4096)
SL.mcx_synth_reads(host, 0, R, B.READ_LEN, genome, B.GENOME, B.P_ERR, 0, 0)
dseq = torch.empty(nbytes + 4096, dtype=torch.uint8, device=dev)
dseq[:nbytes].copy_(torch.frombuffer((C.c_uint8 * nbytes).from_address(host), dtype=torch.uint8))
occ = R * (B.READ_LEN - B.K + 1)
cap_table = int((B.GENOME + R * B.READ_LEN * B.P_ERR * B.K * 1.05) / 0.75)
g = M.Graph(B.K, 1, cap_table)
cap = int(occ * 0.2)
keys = torch.empty(nparts * cap, dtype=torch.int64, device=dev)
meta = torch.empty(nparts * cap, dtype=torch.int32, device=dev)
counts = torch.zeros(nparts, dtype=torch.int64, device=dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); g.set_stream(stream.cuda_stream)
def t(fn, n=3):
    best = 1e9
    for _ in range(n):
        g.clear(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
fused = t(lambda: g.add_reads_raw(dseq.data_ptr(), nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE))
g.flush(); g.sync()
shard = t(lambda: g.add_reads_sharded(dseq.data_ptr(), nbytes, nparts, 0, cap, keys.data_ptr(), meta.data_ptr(), counts.data_ptr()))
print("reads %d: fused %.2f ms (%.1f G/s), sharded(nparts=%d) %.2f ms (%.1f G/s), tuples out %s" % (
    R, fused, occ / fused / 1e6, nparts, shard, occ / shard / 1e6, counts.tolist()))
