"""worker of tests/test_gpu_multi.py: one process per GPU (torchrun), the routed sharded build over real CUDA IPC rings
compared with a single-GPU build AND with the oracle (test infrastructure)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import mccortex_b200 as M
    from mccortex_b200.multi import routed_parity_check
    import bench as B
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    SL = B.synth_lib()
    k, G, per_rank, batches = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    genome = C.create_string_buffer(G)
    SL.mcx_synth_genome(genome, G, 0)

    def check(records):
        # third opinion: the oracle on the same reads (rank 0 only)
        from oracle import oracle as O
        stride = B.READ_LEN + 1
        buf = C.create_string_buffer(per_rank * stride)
        og = O.Graph(k, 1, 1 << 24)
        for r in range(world):
            SL.mcx_synth_reads(buf, r * per_rank, per_rank, B.READ_LEN, genome, G, 0.002, 0, 0)
            for line in buf.raw.split(b"\n")[:per_rank]:
                og.add_read(line.decode())
        want = og.dump_sorted()[len(og.header()):]
        og.close()
        return want == records

    out = routed_parity_check(M, dist, rank, world, dev, stream, k, per_rank, lambda r: r * per_rank,
                              (SL, genome, G, B.READ_LEN, 0.002), batches=batches, check=check)
    if rank == 0:
        print("PARITY-OK %s" % out, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
