"""debug: warp kernel vs fused kernel on the same small inputs, with forced tiny grids (long tile runs per warp)"""
import ctypes as C, hashlib, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import mccortex_b200 as M
    import bench as B
    R = int(sys.argv[2])
    SL = B.synth_lib()
    G = 200000
    genome = C.create_string_buffer(G); SL.mcx_synth_genome(genome, G, 0)
    stride = 151; nbytes = R * stride
    host = M.host_alloc(nbytes + 4096)
    SL.mcx_synth_reads(host, 0, R, 150, genome, G, 0.001, 0, 0)
    dseq = torch.empty(nbytes + 4096, dtype=torch.uint8, device="cuda:0")
    dseq[:nbytes].copy_(torch.frombuffer((C.c_uint8 * nbytes).from_address(host), dtype=torch.uint8))
    torch.cuda.synchronize()
    g = M.Graph(31, 1, 1 << 24)
    g.add_reads_raw(dseq.data_ptr(), nbytes, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE)
    st = g.sync()
    recs, n, rb = g.export_records(sorted=True)
    print("kmers=%d novel=%d contigs=%d reads=%d md5=%s" % (st.num_kmers_loaded, st.num_kmers_novel, st.contigs_parsed, st.num_se_reads, hashlib.md5(recs).hexdigest()))
    sys.exit(0)
for R in (200000, 2000000):
    for name, env in (("fused", {"MCX_KERNEL": "fused"}), ("warp", {}), ("warp grid148", {"MCX_W_GRID": "148"}), ("warp grid296", {"MCX_W_GRID": "296"}),
                      ("warp nofront", {"MCX_FRONT_BITS": "0"}), ("warp 2 classes", {"MCX_CLASSES": "2", "MCX_FRONT_BITS": "18"})):
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, __file__, "child", str(R)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        out = r.stdout.strip().splitlines()
        for l in out[:-1][:30]: print("    " + l)
        print("%-8d %-22s %s" % (R, name, out[-1] if out else "rc=%d" % r.returncode), flush=True)
