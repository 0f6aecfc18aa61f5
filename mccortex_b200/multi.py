"""Multi-GPU plumbing for the build path: one process per GPU, one table shard per process.

Ownership = top bits of the Lookup3 hash.  Per batch the sharded kernel
(`mcx_graph_add_reads_sharded`) counts hot k-mers in the LOCAL L2-resident front table whoever
owns them, inserts what it owns into its big table, and bins the rest as (key, count<<8|edges)
tuples; the bins are exchanged (one all-to-all step: a count matrix, then point-to-point sends of
the ragged bins, NCCL over NVLink on the GPUs / gloo in the CPU tests); kernel C
(`mcx_graph_insert_tuples`) inserts what arrived.  At the end of a step the front tables are
flushed the same way, one aggregated tuple per k-mer.  No remote atomics, no other collective
on the data path.

`exchange_bins` is backend-agnostic so the host-side logic is covered by world_size-2 gloo tests
on CPU tensors (tests/test_multi_cpu.py).
"""
import ctypes as C
import json
import os
import time


# ---------------------------------------------------------------------------------------------
# One sorted .ctx from P shards (output).  Every shard exports its records in ascending
# key order (mcx_graph_export_begin(sorted=1)); ownership is by hash, so each shard's keys cover the
# whole key range and the file order -- ascending (b[0], b[1]), what HASH_ITERATE_SORTED gives the
# reference (src/graph/hash_table.c:362-374) -- needs a P-way merge of disjoint sorted runs.  The merge
# streams: per round it takes a chunk from every run, everything up to the smallest "last key of a
# chunk" is final, is merged (sort of the concatenation) and emitted.  Memory = P chunks.
def _rec_keys(recs, W):
    """recs: uint8 array [n, rec_bytes] -> (w0, w1 or None) as uint64 arrays"""
    import numpy as np
    k = np.ascontiguousarray(recs[:, :8 * W]).view("<u8").reshape(len(recs), W)
    return k[:, 0], (k[:, 1] if W == 2 else None)


def merge_sorted_runs(runs, rec_bytes, W, chunk_recs=1 << 20):
    """runs: list of objects supporting len() in bytes via .nbytes / len and slicing to bytes-like (bytes,
    bytearray, numpy uint8 arrays, numpy.memmap): each a whole number of records sorted by key, keys
    disjoint between runs.  Yields uint8 arrays [m, rec_bytes] whose concatenation is the merged run."""
    import numpy as np
    arrs = []
    for r in runs:
        a = np.frombuffer(r, dtype=np.uint8) if isinstance(r, (bytes, bytearray, memoryview)) else np.asarray(r, dtype=np.uint8).reshape(-1)
        if a.size % rec_bytes:
            raise ValueError("run is not a whole number of records")
        arrs.append(a.reshape(-1, rec_bytes))
    pos = [0] * len(arrs)
    while True:
        live = [i for i, a in enumerate(arrs) if pos[i] < len(a)]
        if not live:
            return
        if len(live) == 1:
            i = live[0]
            while pos[i] < len(arrs[i]):
                yield np.array(arrs[i][pos[i]:pos[i] + chunk_recs])
                pos[i] += chunk_recs
            return
        chunks = {i: np.array(arrs[i][pos[i]:pos[i] + chunk_recs]) for i in live}
        keys = {i: _rec_keys(c, W) for i, c in chunks.items()}
        # everything <= the smallest last key is final: later records of any run are larger
        bound = min((int(keys[i][0][-1]), int(keys[i][1][-1]) if W == 2 else 0) for i in live)
        take, tk0, tk1 = [], [], []
        for i in live:
            w0, w1 = keys[i]
            if W == 1:
                n = int(np.searchsorted(w0, np.uint64(bound[0]), side="right"))
            else:
                n = int(np.count_nonzero((w0 < np.uint64(bound[0])) | ((w0 == np.uint64(bound[0])) & (w1 <= np.uint64(bound[1])))))
            if n:
                take.append(chunks[i][:n]); tk0.append(w0[:n])
                if W == 2:
                    tk1.append(w1[:n])
                pos[i] += n
        cat = np.concatenate(take)
        k0 = np.concatenate(tk0)
        order = np.argsort(k0, kind="stable") if W == 1 else np.lexsort((np.concatenate(tk1), k0))
        yield cat[order]


def write_ctx_from_shards(out_fh, header_bytes, shard_paths, rec_bytes, W, chunk_recs=1 << 20):
    """rank 0 of a sharded build: header + the merge of the shards' sorted record files (each rank has written
    its `export_records(sorted=True)` to shard_paths[rank]).  Returns the number of records written."""
    import numpy as np
    out_fh.write(header_bytes)
    runs = [np.memmap(p, dtype=np.uint8, mode="r") if os.path.getsize(p) else np.zeros(0, dtype=np.uint8) for p in shard_paths]
    n = 0
    for piece in merge_sorted_runs(runs, rec_bytes, W, chunk_recs):
        out_fh.write(piece.tobytes())
        n += len(piece)
    return n


FRONT_SPAN = 0xE0000000  # positions between two flushes of a sharded graph's front table (library limit: 0xF0000000)


def exchange_counts(dist, counts):
    """counts[d] = tuples this rank holds for rank d  ->  recv[s] = tuples rank s holds for us"""
    import torch
    recv = torch.empty_like(counts)
    dist.all_to_all_single(recv, counts)
    return recv


def exchange_bins(dist, rank, world, keys, masks, counts_host, recv_counts_host, cap, W, recv_keys, recv_masks):
    """Ragged all-to-all of the bins.

    keys: [world*cap*W] int64, masks: [world*cap] uint8 (bin d starts at d*cap); recv_* likewise
    (bin s = tuples received from rank s).  counts_host / recv_counts_host: python ints.
    The local bin is copied, the others go through grouped isend/irecv (one NCCL group)."""
    ops = []
    for peer in range(world):
        n_out, n_in = counts_host[peer], recv_counts_host[peer]
        if n_in > cap:
            raise RuntimeError("receive bin overflow: %d > %d" % (n_in, cap))
        if peer == rank:
            recv_keys[peer * cap * W: (peer * cap + n_in) * W].copy_(keys[peer * cap * W: (peer * cap + n_out) * W])
            recv_masks[peer * cap: peer * cap + n_in].copy_(masks[peer * cap: peer * cap + n_out])
            continue
        if n_out:
            ops.append(dist.P2POp(dist.isend, keys[peer * cap * W: (peer * cap + n_out) * W], peer))
            ops.append(dist.P2POp(dist.isend, masks[peer * cap: peer * cap + n_out], peer))
        if n_in:
            ops.append(dist.P2POp(dist.irecv, recv_keys[peer * cap * W: (peer * cap + n_in) * W], peer))
            ops.append(dist.P2POp(dist.irecv, recv_masks[peer * cap: peer * cap + n_in], peer))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()


class ShardedBuilder:
    """One rank's side of the sharded build (device tensors owned by torch = plumbing only)."""

    def __init__(self, M, dist, rank, world, device, k, capacity_per_shard, cap_per_part):
        import torch
        self.M, self.dist, self.rank, self.world, self.dev = M, dist, rank, world, device
        self.k, self.W, self.cap = k, (k + 31) // 32, cap_per_part
        self.g = M.Graph(k, 1, capacity_per_shard, device=device.index)
        n = world * cap_per_part
        self.keys = torch.empty(n * self.W, dtype=torch.int64, device=device)
        self.masks = torch.empty(n, dtype=torch.int32, device=device)   # meta = count << 8 | edge mask
        self.rkeys = torch.empty(n * self.W, dtype=torch.int64, device=device)
        self.rmasks = torch.empty(n, dtype=torch.int32, device=device)
        self.counts = torch.zeros(world, dtype=torch.int64, device=device)
        self.launches = 0
        self.pending = 0

    def set_stream(self, stream):
        self.g.set_stream(stream.cuda_stream)

    def _exchange_and_insert(self):
        g, W, cap, world = self.g, self.W, self.cap, self.world
        recv = exchange_counts(self.dist, self.counts)
        ch, rh = self.counts.tolist(), recv.tolist()
        if max(ch) > cap:
            raise RuntimeError("send bin overflow: %d > %d" % (max(ch), cap))
        exchange_bins(self.dist, self.rank, world, self.keys, self.masks, ch, rh, cap, W, self.rkeys, self.rmasks)
        for s in range(world):
            if rh[s]:
                g.insert_tuples(self.rkeys[s * cap * W:].data_ptr(), self.rmasks[s * cap:].data_ptr(), rh[s])
                self.launches += 1
        return sum(ch)

    def add_batch(self, seq_addr, nbytes, aggregate=True):
        """one batch of local reads; aggregate=False is the unaggregated baseline (kernel B: every
        occurrence leaves as a tuple)"""
        g = self.g
        if aggregate:
            if self.pending + nbytes >= FRONT_SPAN:   # 32-bit front-table counters (see RoutedBuilder.add_batch)
                self.flush()
            self.pending += nbytes
            g.add_reads_sharded(seq_addr, nbytes, self.world, self.rank, self.cap, self.keys.data_ptr(),
                                self.masks.data_ptr(), self.counts.data_ptr())
        else:
            g.kmer_tuples(seq_addr, nbytes, self.world, self.cap, self.keys.data_ptr(), self.masks.data_ptr(),
                          self.counts.data_ptr())
        self.launches += 1
        return self._exchange_and_insert()

    def flush(self):
        """end of a step: forward the aggregated front-table records of keys owned elsewhere"""
        self.pending = 0
        self.g.flush_sharded(self.world, self.rank, self.cap, self.keys.data_ptr(), self.masks.data_ptr(),
                             self.counts.data_ptr())
        self.launches += 1
        return self._exchange_and_insert()


class RoutedBuilder:
    """One rank's side of the sharded build with the exchange fused into the kernel.

    Every rank owns two receive rings (double buffer) of `world` regions x `cap` tuples, allocated
    by the library and mapped into every peer process through CUDA IPC.  The sharded kernel of
    rank s appends the tuples it produces for shard d straight into region s of d's ring -- its
    stores travel over NVLink while the kernel is still hashing -- so the only thing left to
    exchange per batch is the `world` tuple counters (one tiny all_to_all, stream ordered: it also
    tells the receiver that every sender's kernel, and therefore its stores, has completed).  The
    receiver's insert kernels read their tuple count from device memory, so the host never waits
    for the GPU inside a step.

    Ring reuse: rank s writes ring j of rank d again in batch b+2; its kernel b+2 is ordered behind
    the counter all_to_all of batch b+1, which completes only after d has entered it, i.e. after d's
    inserts of batch b (stream order on d).  Two rings suffice.

    peers: None = IPC (one process per GPU); tests pass rings of sibling builders on the same
    device instead (see connect_local)."""

    NRING = 2

    def __init__(self, M, dist, rank, world, device, k, capacity_per_shard, cap_per_part, ncols=1):
        import torch
        self.M, self.dist, self.rank, self.world, self.dev = M, dist, rank, world, device
        self.k, self.W, self.cap = k, (k + 31) // 32, cap_per_part
        self.g = M.Graph(k, ncols, capacity_per_shard, device=device.index)
        self.kbytes = world * cap_per_part * self.W * 8
        self.mbytes = world * cap_per_part * 4
        self.ring_k = [M.device_alloc(device.index, self.kbytes) for _ in range(self.NRING)]
        self.ring_m = [M.device_alloc(device.index, self.mbytes) for _ in range(self.NRING)]
        self.counts = [torch.zeros(world, dtype=torch.int64, device=device) for _ in range(self.NRING)]
        self.rcounts = [torch.zeros(world, dtype=torch.int64, device=device) for _ in range(self.NRING)]
        self.peer_k = self.peer_m = None   # [ring][dest] addresses of MY region on dest
        self.opened = []
        self.batch = 0
        self.launches = 0
        self.prof = None
        self.sent = []
        self.pending = 0
        self.insert_stream = None
        self.sent_total = torch.zeros((), dtype=torch.int64, device=device)

    def set_stream(self, stream):
        self.g.set_stream(stream.cuda_stream)

    def connect_ipc(self):
        """exchange the rings' IPC handles (all_gather) and map every peer's rings"""
        import torch
        M, dist, world, rank = self.M, self.dist, self.world, self.rank
        mine = b"".join(M.ipc_export(a) for a in self.ring_k + self.ring_m)
        t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.dev)
        allh = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allh, t)
        bases = []
        for d in range(world):
            if d == rank:
                bases.append((list(self.ring_k), list(self.ring_m)))
                continue
            raw = bytes(allh[d].cpu().numpy().tobytes())
            hs = [raw[i * 64:(i + 1) * 64] for i in range(2 * self.NRING)]
            ptrs = [M.ipc_open(self.dev.index, h) for h in hs]
            self.opened += ptrs
            bases.append((ptrs[:self.NRING], ptrs[self.NRING:]))
        self._set_peers(bases)

    def connect_local(self, builders):
        """all shards in one process (tests): peers' rings are plain device pointers"""
        self._set_peers([(list(b.ring_k), list(b.ring_m)) for b in builders])

    def _set_peers(self, bases):
        r, cap, W = self.rank, self.cap, self.W
        self.peer_k = [[bases[d][0][j] + r * cap * W * 8 for d in range(self.world)] for j in range(self.NRING)]
        self.peer_m = [[bases[d][1][j] + r * cap * 4 for d in range(self.world)] for j in range(self.NRING)]

    # the three stages of one batch; the one-process test drives them shard by shard
    def produce(self, seq_addr, nbytes, colour=0, hp_cutoff=0):
        j = self.batch % self.NRING
        self.g.add_reads_routed(seq_addr, nbytes, self.world, self.rank, self.cap, self.peer_k[j], self.peer_m[j],
                                self.counts[j].data_ptr(), hp_cutoff=hp_cutoff, colour=colour)
        self.launches += 1

    def produce_flush(self):
        j = self.batch % self.NRING
        self.g.flush_routed(self.world, self.rank, self.cap, self.peer_k[j], self.peer_m[j], self.counts[j].data_ptr())
        self.launches += 1

    def exchange_counts(self):
        j = self.batch % self.NRING
        self.sent_total += self.counts[j].sum()   # device-side: tuples this rank shipped (for the roofline accounting)
        if self.insert_stream is not None and self.batch >= 1:
            # completing this exchange tells the peers that the ring of the PREVIOUS batch is free again
            import torch
            torch.cuda.current_stream().wait_event(self.insert_done[(self.batch - 1) % self.NRING])
        self.dist.all_to_all_single(self.rcounts[j], self.counts[j])
        if self.prof is not None:
            self.sent.append(self.counts[j].tolist())

    def consume(self, colour=0):
        """insert what arrived for this batch.  With an insert stream (use_insert_stream) the kernels
        run beside the NEXT batch's sharded kernel: they are ordered behind this batch's counter
        exchange, and the next exchange -- the event that lets peers overwrite this ring two batches
        later -- waits for them."""
        j = self.batch % self.NRING
        cap, W = self.cap, self.W
        ist = self.insert_stream
        if ist is not None:
            import torch
            ist.wait_stream(torch.cuda.current_stream())
        for s in range(self.world):
            if s == self.rank:
                continue
            args = (self.ring_k[j] + s * cap * W * 8, self.ring_m[j] + s * cap * 4, self.rcounts[j].data_ptr() + 8 * s, cap)
            if ist is not None:
                self.g.insert_tuples_on(ist.cuda_stream, *args, colour=colour)
            else:
                self.g.insert_tuples_n(*args, colour=colour)
            self.launches += 1
        if ist is not None:
            self.insert_done[j].record(ist)
        self.batch += 1

    def use_insert_stream(self, stream):
        import torch
        self.insert_stream = stream
        self.insert_done = [torch.cuda.Event() for _ in range(self.NRING)]

    def join_inserts(self):
        """make the current stream wait for every insert kernel queued so far"""
        if self.insert_stream is not None:
            import torch
            torch.cuda.current_stream().wait_stream(self.insert_stream)

    def _mark(self, name):
        # optional per-stage device timing (MCX_MULTI_PROFILE=1): CUDA events on the current stream
        if self.prof is not None:
            import torch
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.prof.append((name, e))

    def add_batch(self, seq_addr, nbytes, colour=0):
        # the front table counts in 32 bits: it must be flushed (a collective step: every rank does it
        # at the same batch, so batches are assumed to be the same size on every rank) before 2^32
        # positions have gone through it
        if self.pending + nbytes >= FRONT_SPAN:
            self.flush()
        self.pending += nbytes
        self._mark("start")
        self.produce(seq_addr, nbytes, colour)
        self._mark("produce")
        self.exchange_counts()
        self._mark("counts")
        self.consume(colour)
        self._mark("consume")

    def flush(self):
        self.pending = 0
        self._mark("start")
        self.produce_flush()
        self._mark("flush_produce")
        self.exchange_counts()
        self._mark("counts")
        self.consume()
        self.join_inserts()
        self._mark("flush_consume")

    def profile_summary(self):
        tot = {}
        for (n0, e0), (n1, e1) in zip(self.prof, self.prof[1:]):
            if n1 != "start":
                tot[n1] = tot.get(n1, 0.0) + e0.elapsed_time(e1)
        return tot

    def close(self):
        M = self.M
        self.g.close()
        for p in self.opened:
            M.ipc_close(self.dev.index, p)
        self.opened = []
        for a in self.ring_k + self.ring_m:
            M.device_free(self.dev.index, a)
        self.ring_k = self.ring_m = []


def routed_parity_check(M, dist, rank, world, dev, stream, k, reads_per_rank, first_read_of, synth, batches=5, check=None):
    """Hardware parity of the sharded build: the routed build -- IPC rings over NVLink, two-ring
    reuse, insert stream, front-table flush -- of reads_per_rank reads on every rank, in `batches` batches so that every
    ring is reused, must give byte for byte the sorted records of ONE single-GPU fused build of the same reads.
    Every shard is exported sorted, gathered on rank 0 and merged (merge_sorted_runs); rank 0 builds the union of the
    reads on its own GPU with the single-GPU kernel and compares md5s.  `check(records_bytes)` (tests: the oracle) may
    add a third opinion.  Returns a dict (rank 0) / None; raises on every rank if the records differ."""
    import hashlib
    import numpy as np
    import torch
    SL, genome, G, read_len, p_err = synth
    stride = read_len + 1
    nb = reads_per_rank * stride
    host = M.host_alloc(nb + 4096)
    SL.mcx_synth_reads(host, first_read_of(rank), reads_per_rank, read_len, genome, G, p_err, 0, 0)
    dseq = torch.empty(nb + 4096, dtype=torch.uint8, device=dev)
    dseq[:nb].copy_(torch.frombuffer((C.c_uint8 * nb).from_address(host), dtype=torch.uint8))
    torch.cuda.synchronize()
    M.host_free(host)
    distinct_est = int(G + world * reads_per_rank * read_len * p_err * k * 1.05) if G < 1e9 else int(world * reads_per_rank * (read_len - k + 1) * 1.05)
    cap_shard = int(distinct_est / world / 0.75 * 1.1) + 4096
    per_batch = ((reads_per_rank + batches - 1) // batches + 15) // 16 * 16   # device batches must start 16-byte aligned
    cap_part = int(per_batch * (read_len - k + 1) * (0.5 if G < 1e9 else 1.3) / max(1, world - 1)) + (4 << 20)
    sb = RoutedBuilder(M, dist, rank, world, dev, k, cap_shard, cap_part)
    sb.connect_ipc()
    sb.set_stream(stream)
    sb.use_insert_stream(torch.cuda.Stream(device=dev))
    for b in range(batches):   # (every rank runs every batch: the counter exchange is a collective)
        lo = min(b * per_batch, reads_per_rank)
        n = min(per_batch, reads_per_rank - lo)
        sb.add_batch(dseq.data_ptr() + lo * stride, n * stride)
    sb.flush()
    st = sb.g.sync()
    recs, nrec, rb = sb.g.export_records(sorted=True)
    sb.close()
    # gather the shards' sorted runs on rank 0
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = len(recs)
    dist.all_reduce(sizes)
    sizes = sizes.tolist()
    mx = max(max(sizes), 1)
    mine = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if recs:
        mine[:len(recs)] = torch.frombuffer(bytearray(recs), dtype=torch.uint8).to(dev)
    gathered = [torch.empty(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    tot = torch.tensor([st.num_kmers_loaded], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    verdict = torch.zeros(1, dtype=torch.int64, device=dev)
    out = None
    if rank == 0:
        W = (k + 31) // 32
        runs = [gathered[r][:sizes[r]].cpu().numpy() for r in range(world)]
        h = hashlib.md5(); merged_n = 0; pieces = []
        for piece in merge_sorted_runs(runs, rb, W):
            b = piece.tobytes(); h.update(b); merged_n += len(piece)
            if check is not None:
                pieces.append(b)
        # the same reads, one GPU, single-GPU kernel
        allb = world * nb
        hall = M.host_alloc(allb + 4096)
        for r in range(world):
            SL.mcx_synth_reads(hall + r * nb, first_read_of(r), reads_per_rank, read_len, genome, G, p_err, 0, 0)
        dall = torch.empty(allb + 4096, dtype=torch.uint8, device=dev)
        dall[:allb].copy_(torch.frombuffer((C.c_uint8 * allb).from_address(hall), dtype=torch.uint8))
        torch.cuda.synchronize()
        M.host_free(hall)
        g1 = M.Graph(k, 1, int(distinct_est / 0.75 * 1.1) + 4096, device=dev.index)
        g1.set_stream(stream.cuda_stream)
        g1.add_reads_raw(dall.data_ptr(), allb, M.MCX_LAYOUT_LINES, M.MCX_MEM_DEVICE)
        st1 = g1.sync()
        recs1, n1, _ = g1.export_records(sorted=True)
        g1.close()
        md5_1 = hashlib.md5(recs1).hexdigest()
        ok = (md5_1 == h.hexdigest()) and n1 == merged_n and st1.num_kmers_loaded == int(tot[0])
        if ok and check is not None:
            ok = bool(check(b"".join(pieces)))
        out = {"ok": ok, "reads": world * reads_per_rank, "batches_per_rank": batches, "records": merged_n,
               "md5_sharded_merged": h.hexdigest(), "md5_single_gpu": md5_1, "kmers_loaded": int(tot[0]),
               "what": "routed %d-shard build (IPC rings over NVLink) vs one single-GPU fused build of the same reads: sorted .ctx records" % world}
        verdict[0] = 1 if ok else 2
    dist.broadcast(verdict, src=0)
    if int(verdict[0]) != 1:
        raise RuntimeError("sharded build parity FAILED on %d GPUs: %s" % (world, out))
    return out


def bind_to_gpu_numa_node(local):
    """run this rank on the cores of the NUMA node its GPU hangs off, so that the pinned read buffer it is about to
    allocate and fill (first touch) is local to the GPU's PCIe root: eight ranks copying 7.5 GB each from one host share
    the inter-socket links otherwise.  Returns the node number or None (information missing: nothing changed)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        if hasattr(pr, "pci_bus_id") and hasattr(pr, "pci_device_id"):
            bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        else:
            import subprocess
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                                 stdout=subprocess.PIPE, text=True, timeout=10).stdout.strip()
            bus = out[-12:] if out else None     # 0000:19:00.0
        if not bus:
            return None
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus.lower()).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def bench_multi(args, rank, world, local, dist):
    """bench.py body for N > 1 (weak scaling: args.reads reads per GPU, disjoint read index ranges
    of the same genome)."""
    import torch
    import mccortex_b200 as M
    import bench as B

    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    SL = B.synth_lib()
    R, stride = args.reads, B.READ_LEN + 1
    genome = C.create_string_buffer(B.GENOME)
    SL.mcx_synth_genome(genome, B.GENOME, 0)
    batch_reads = min(R, int(os.environ.get("MCX_MULTI_BATCH_READS", 4_000_000)))
    nb = (R + batch_reads - 1) // batch_reads
    nbytes = R * stride
    host = M.host_alloc(nbytes + 4096)
    SL.mcx_synth_reads(host, rank * R, R, B.READ_LEN, genome, B.GENOME, B.P_ERR, 0, 0)
    dseq = torch.empty(nbytes + 4096, dtype=torch.uint8, device=dev)
    dseq[:nbytes].copy_(torch.frombuffer((C.c_uint8 * nbytes).from_address(host), dtype=torch.uint8))
    torch.cuda.synchronize()
    hview = torch.frombuffer((C.c_uint8 * nbytes).from_address(host), dtype=torch.uint8)  # pinned (cudaHostAlloc)

    occ_per_rank = R * (B.READ_LEN - B.K + 1)
    distinct_est = int(B.GENOME + world * R * B.READ_LEN * B.P_ERR * B.K * 1.05)
    cap_shard = int(distinct_est / world / 0.75 * 1.05)
    # bins: the unaggregated baseline ships every occurrence; the aggregated path ships what the
    # front table cannot absorb (sized generously: a quarter of the occurrences) and, at the
    # flush, at most one tuple per front-table slot (8.4 M)
    per_peer = batch_reads * (B.READ_LEN - B.K + 1) / world
    aggregate_env = os.environ.get("MCX_MULTI_AGGREGATE", "1") != "0"
    cap_part = int(max(per_peer * (float(os.environ.get("MCX_MULTI_BIN_FRAC", 0.3)) if aggregate_env else 1.25), (10 << 20) / max(1, world - 1) * 1.3)) + 4096
    aggregate = aggregate_env
    # default: the exchange is fused into the kernel (stores into the owners' rings over NVLink);
    # MCX_MULTI_EXCHANGE=nccl keeps the bins local and moves them with NCCL send/recv (baseline)
    routed = aggregate and os.environ.get("MCX_MULTI_EXCHANGE", "peer") != "nccl"
    if routed:
        sb = RoutedBuilder(M, dist, rank, world, dev, B.K, cap_shard, cap_part)
        sb.connect_ipc()
    else:
        sb = ShardedBuilder(M, dist, rank, world, dev, B.K, cap_shard, cap_part)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sb.set_stream(stream)
    if routed and os.environ.get("MCX_MULTI_OVERLAP", "1") != "0":
        sb.use_insert_stream(torch.cuda.Stream(device=dev))

    def add(addr, nbytes_):
        if routed:
            sb.add_batch(addr, nbytes_)
        else:
            sb.add_batch(addr, nbytes_, aggregate=aggregate)

    def step():
        sb.g.clear()
        for b in range(nb):
            lo = b * batch_reads
            n = min(batch_reads, R - lo)
            add(dseq.data_ptr() + lo * stride, n * stride)
        if aggregate:
            sb.flush()

    # e2e: the same step fed from the PINNED HOST buffer: H2D of every batch inside the timed region
    # (double-buffered on a copy stream so batch b+1 uploads while batch b is processed), and the
    # counters are read back to the host at the end of every step
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty(batch_reads * stride + 4096, dtype=torch.uint8, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def upload(b):
        lo = b * batch_reads
        n = min(batch_reads, R - lo)
        s = b % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            stage[s][:n * stride].copy_(hview[lo * stride:(lo + n) * stride], non_blocking=True)
            ready[s].record(copy_stream)
        return n

    def step_host():
        sb.g.clear()
        n = upload(0)
        for b in range(nb):
            s = b % 2
            stream.wait_event(ready[s])
            n_next = upload(b + 1) if b + 1 < nb else 0
            add(stage[s].data_ptr(), n * stride)
            freed[s].record(stream)
            n = n_next
        if aggregate:
            sb.flush()
        st_ = sb.g.sync()
        # the job's result: this shard's records, sorted, in host memory (the sorted runs the P-way merge writes out)
        nrec, rb = C.c_uint64(), C.c_uint32()
        M.binding._ck(M.lib().mcx_graph_export_begin(sb.g.h, 1, C.byref(nrec), C.byref(rb)), "export_begin")
        nb_out = int(nrec.value) * int(rb.value)
        if nb_out > hout_bytes[0]:
            if hout[0]:
                M.host_free(hout[0])
            hout_bytes[0] = nb_out + nb_out // 8 + 4096
            hout[0] = M.host_alloc(hout_bytes[0])
        M.binding._ck(M.lib().mcx_graph_export_read(sb.g.h, 0, nrec.value, C.c_void_p(hout[0])), "export_read")
        M.lib().mcx_graph_export_end(sb.g.h)
        d2h[0] = nb_out
        return st_

    hout, hout_bytes, d2h = [0], [0], [0]
    for _ in range(args.warmup):
        step()
    st = sb.g.sync()
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sb.launches = 0
    if routed:
        sb.sent_total.zero_()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    sent_timed = int(sb.sent_total) if routed else 0   # (before the optional profiling step adds its own)
    if routed and os.environ.get("MCX_MULTI_PROFILE"):
        import sys
        sb.prof = []
        step()
        torch.cuda.synchronize()
        print("rank %d stage ms over one step: %s; tuples sent per batch (cap %d): %s" % (
            rank, json.dumps(sb.profile_summary()), cap_part, sb.sent), file=sys.stderr, flush=True)
        sb.prof = None
        dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    st = sb.g.sync()
    tot = torch.tensor([st.num_kmers_loaded, st.num_kmers_novel, sb.launches, sent_timed], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    # hardware parity of the routed build on a bounded sample of the same workload (1M reads per rank), every run
    parity = None
    if routed and os.environ.get("MCX_MULTI_PARITY", "1") != "0":
        parity = routed_parity_check(M, dist, rank, world, dev, stream, B.K, min(R, 1_000_000), lambda r: r * R,
                                     (SL, genome, B.GENOME, B.READ_LEN, B.P_ERR))
        if rank == 0:
            import sys
            print("parity: ok (%d GPUs, %d reads, %d records, md5 %s == single-GPU build)" % (
                world, parity["reads"], parity["records"], parity["md5_sharded_merged"]), file=sys.stderr, flush=True)

    e_steps = max(1, min(args.steps, 3))
    step_host()
    dist.barrier()
    torch.cuda.synchronize()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record(stream)
    t0 = time.perf_counter()
    for _ in range(e_steps):
        st_h = step_host()
    ee1.record(stream)
    torch.cuda.synchronize()
    e_wall = time.perf_counter() - t0
    dist.barrier()
    e_ms = torch.tensor([max(ee0.elapsed_time(ee1), e_wall * 1e3)], device=dev)
    dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e_tot = torch.tensor([st_h.num_kmers_loaded], dtype=torch.int64, device=dev)
    dist.all_reduce(e_tot)
    d2h_tot = torch.tensor([d2h[0]], dtype=torch.int64, device=dev)
    dist.all_reduce(d2h_tot)
    assert int(e_tot[0]) == occ_per_rank * world
    e2e_val = occ_per_rank * world * e_steps / (float(e_ms[0]) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    assert int(tot[0]) == occ_per_rank * world, (int(tot[0]), occ_per_rank * world)
    ms_total = float(ms[0])
    value = occ_per_rank * world * args.steps / (ms_total * 1e-3)
    peak, peak_src = B.measured_peaks()
    if rank == 0:
        # algorithmic bytes: every occurrence 19.25 B; only what became a tuple pays the tuple's write, send
        # and re-read (+24 B) -- on this workload the local front tables absorb almost everything
        tuples_per_step = int(tot[3]) / max(1, args.steps)
        ach = (value * B.B_ALG + tuples_per_step / (ms_total / args.steps * 1e-3) * (B.B_ALG_MULTI - B.B_ALG)) / 1e9 / world
        line = {
            "metric": "kmers_per_sec_build_k31", "value": value, "unit": "k-mers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": B.workload_config(args, world),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": None, "peak_source": peak_src,
                         "kernel": "mcx_build_sharded_kernel + mcx_insert_tuples_kernel (per GPU)",
                         "note": "per GPU: (occurrences x 19.25 B + tuples x 24 B) / time; tuples = what the local front tables did not absorb",
                         "alg_bytes_per_kmer": B.B_ALG, "alg_bytes_per_tuple": B.B_ALG_MULTI - B.B_ALG,
                         "tuples_per_step": tuples_per_step, "nvlink_bytes_per_step": tuples_per_step * 12 * (world - 1) / world},
            "cpu_baseline": None,
            "e2e": {"value": e2e_val, "unit": "k-mers/s", "h2d_bytes_per_step": nbytes * world,
                    "d2h_bytes_per_step": int(d2h_tot[0]) + (64 + 72) * world, "steps": e_steps, "ms_per_step": float(e_ms[0]) / e_steps,
                    "what": "per rank: pinned host reads -> H2D -> sharded build -> this shard's sorted records back in host memory"},
            "gpu_launches": int(tot[2]),
            "clocks": clocks,
            "parity": parity,
            "extra": {"distinct_kmers_total": int(tot[1]), "shard_slots": cap_shard, "batches_per_step": nb, "numa_node_rank0": numa,
                      "exchange": ("fused into the kernel: peer stores over NVLink (CUDA IPC rings), counters all_to_all"
                                   if routed else "NCCL send/recv of local bins") +
                                  (", front-table aggregated" if aggregate else ", every occurrence (baseline)")},
        }
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    if routed:
        sb.close()
    else:
        sb.g.close()
    M.host_free(host)
    dist.barrier()
    dist.destroy_process_group()
