"""the insert kernel (kernel C, mcx_insert_tuples_kernel) alone on a COLD table: distinct random keys into a table >> L2.
This is the north star's "insert kernel at >= 50 % of the HBM roofline" measured literally.
usage: python scripts/insert_bench.py [log2_slots=30] [Mtuples=256]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mccortex_b200 as M
LS = int(sys.argv[1]) if len(sys.argv) > 1 else 30
N = (int(sys.argv[2]) if len(sys.argv) > 2 else 256) * 1_000_000
K = int(os.environ.get("KBENCH_K", 31))
dev = torch.device("cuda:0")
torch.cuda.init(); torch.zeros(1, device=dev)
if os.environ.get("L2FETCH"):
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    cur = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(cur), 5)
    r = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["L2FETCH"])))
    new = ctypes.c_size_t(); rt.cudaDeviceGetLimit(ctypes.byref(new), 5)
    print("cudaLimitMaxL2FetchGranularity: was %d, set rc=%d, now %d" % (cur.value, r, new.value), flush=True)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
g = M.Graph(K, 1, 1 << LS); g.set_stream(stream.cuda_stream)
W = 1 if K <= 31 else 2
gen = torch.Generator(device=dev); gen.manual_seed(1)
keys = torch.randint(0, 1 << 62, (N * W,), dtype=torch.int64, device=dev, generator=gen)
if W == 2: keys[0::2] &= (1 << (2 * (K - 32))) - 1
meta = torch.full((N,), (1 << 8) | 0x21, dtype=torch.int32, device=dev)
slot_bytes = 16 if W == 1 else 32
alg = 8 * W + 8 + 2 + (8 * W + 4)   # key compare + covg RMW + edge RMW + the tuple read
def run(label):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    g.insert_tuples(keys.data_ptr(), meta.data_ptr(), N)
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-28s k=%d table 2^%d slots (%.1f GB), %d M tuples: %.2f ms  %.2f G inserts/s  logical %.0f GB/s = %.3f of 6537 (alg %d B/insert); "
          "64 B/insert physical floor -> %.0f GB/s" % (label, K, LS, (slot_bytes << LS) / 1e9, N // 10**6, ms, N / ms / 1e6, N * alg / ms / 1e6,
                                                       N * alg / ms / 1e6 / 6537.0, alg, N * 64 / ms / 1e6), flush=True)
run("novel keys (cold, load 0->%.2f)" % (N / (1 << LS)))
st = g.sync(); assert st.num_kmers_novel >= N * 0.999, st.as_dict()
run("same keys again (all found)")
run("same keys third time")
st = g.sync(); assert st.num_kmers_novel == 0
g.close()
