"""CPU (gloo, world_size 2) test of the multi-GPU host plumbing in mccortex_b200/multi.py:
the ragged bin exchange delivers every tuple to the rank that owns it (owner = the same
function the kernels use, exported as mcx_key_owner), nothing is lost or duplicated."""
import os
import random
import socket
import sys

import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import mccortex_b200 as M
    from mccortex_b200.multi import exchange_counts, exchange_bins
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for k, W in ((31, 1), (63, 2)):
            rng = random.Random(1234 + rank + k)
            cap = 700
            tuples = []
            for _ in range(rng.randint(300, 900)):
                key = [rng.getrandbits(62 if k == 31 else 62), rng.getrandbits(64)][:W]
                tuples.append((key, rng.getrandbits(8)))
            keys = torch.zeros(world * cap * W, dtype=torch.int64)
            masks = torch.zeros(world * cap, dtype=torch.uint8)
            counts = [0] * world
            for key, m in tuples:
                d = M.key_owner(key, k, world)
                at = d * cap + counts[d]
                for w in range(W):
                    v = key[w]
                    keys[at * W + w] = v - (1 << 64) if v >= (1 << 63) else v
                masks[at] = m
                counts[d] += 1
            ct = torch.tensor(counts, dtype=torch.int64)
            recv = exchange_counts(dist, ct)
            rkeys = torch.zeros_like(keys)
            rmasks = torch.zeros_like(masks)
            exchange_bins(dist, rank, world, keys, masks, counts, recv.tolist(), cap, W, rkeys, rmasks)
            got = []
            for s in range(world):
                for i in range(int(recv[s])):
                    at = s * cap + i
                    key = [int(rkeys[at * W + w]) & ((1 << 64) - 1) for w in range(W)]
                    assert M.key_owner(key, k, world) == rank
                    got.append((tuple(key), int(rmasks[at])))
            # gather everything everywhere and compare multisets
            mine_all = [None] * world
            dist.all_gather_object(mine_all, [(tuple(key), m) for key, m in tuples])
            want = sorted(t for lst in mine_all for t in lst if M.key_owner(list(t[0]), k, world) == rank)
            assert sorted(got) == want, "rank %d k %d: exchange lost or duplicated tuples" % (rank, k)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, "FAIL: %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_bin_exchange_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def _routed_worker(rank, world, port, q):
    """RoutedBuilder host logic without a GPU: ring addressing (every sender owns region `rank` of the
    destination's ring) and the counter all_to_all, with the library calls replaced by recorders."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from mccortex_b200.multi import RoutedBuilder
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        class FakeGraph:
            def __init__(self, *a, **kw): pass
            def add_reads_routed(self, seq, nbytes, nparts, my_part, cap, kptrs, mptrs, counts, hp_cutoff=0, colour=0):
                calls.append(("produce", list(kptrs), list(mptrs)))
            def flush_routed(self, nparts, my_part, cap, kptrs, mptrs, counts):
                calls.append(("flush", list(kptrs), list(mptrs)))
            def insert_tuples_n(self, keys, meta, n_dev, n_max, colour=0):
                calls.append(("insert", keys, meta, n_dev, n_max))
            def close(self): pass

        class FakeM:
            Graph = FakeGraph
            _next = [1 << 40]
            @staticmethod
            def device_alloc(dev, nbytes):
                a = FakeM._next[0] + (rank << 36)
                FakeM._next[0] += (nbytes + 255) // 256 * 256
                return a
            @staticmethod
            def device_free(dev, a): pass

        class Dev:
            index = 0
        k, W, cap = 63, 2, 1000
        sb = RoutedBuilder(FakeM, dist, rank, world, torch.device("cpu"), k, 1 << 20, cap)
        # exchange ring bases the way connect_ipc does, but as plain integers
        mine = torch.tensor(sb.ring_k + sb.ring_m, dtype=torch.int64)
        allb = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine)
        bases = [(b[:2].tolist(), b[2:].tolist()) for b in allb]
        sb._set_peers(bases)
        for j in range(2):
            for d in range(world):
                assert sb.peer_k[j][d] == bases[d][0][j] + rank * cap * W * 8
                assert sb.peer_m[j][d] == bases[d][1][j] + rank * cap * 4
        for b in range(3):
            j = b % 2
            sb.counts[j][:] = torch.tensor([100 * rank + d + b for d in range(world)])
            sb.produce(0, 0)
            sb.exchange_counts()
            assert sb.rcounts[j].tolist() == [100 * s + rank + b for s in range(world)]
            sb.consume()
            ins = [c for c in calls if c[0] == "insert"][-(world - 1):]
            srcs = [s for s in range(world) if s != rank]
            for c, s in zip(ins, srcs):
                assert c[1] == sb.ring_k[j] + s * cap * W * 8 and c[2] == sb.ring_m[j] + s * cap * 4
                assert c[3] == sb.rcounts[j].data_ptr() + 8 * s and c[4] == cap
        assert sb.batch == 3
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %r %s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


def test_routed_builder_host_logic_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_routed_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


# ---- one sorted .ctx from P shards: the streaming P-way merge of mccortex_b200/multi.py -------------------------
import random as _random

import pytest as _pytest


@_pytest.mark.parametrize("W,ncols,nshards,n,chunk", [(1, 1, 2, 5000, 64), (1, 4, 8, 20000, 1000), (2, 1, 3, 7000, 50), (2, 2, 5, 3000, 1 << 20),
                                                      (1, 1, 4, 0, 16), (2, 1, 1, 1000, 7)])
def test_merge_of_sorted_shard_runs(tmp_path, W, ncols, nshards, n, chunk):
    """records with distinct random keys are dealt to shards by a hash of the key (as the build does), every shard sorts its
    own: the merge must reproduce the globally sorted file, through memory and through shard files; empty shards, one
    shard, chunks far smaller than the runs"""
    import numpy as np
    from mccortex_b200.multi import merge_sorted_runs, write_ctx_from_shards
    rng = _random.Random(W * 1000 + nshards)
    rb = 8 * W + 5 * ncols
    keys = set()
    while len(keys) < n:
        # few distinct high words so that the second word decides often (W = 2)
        keys.add((rng.getrandbits(62) if W == 1 else rng.randrange(40), rng.getrandbits(64) if W == 2 else 0))
    recs = []
    for k0, k1 in keys:
        r = k0.to_bytes(8, "little") + (k1.to_bytes(8, "little") if W == 2 else b"") + rng.randbytes(5 * ncols)
        recs.append(((k0, k1), r))
    want = b"".join(r for _, r in sorted(recs))
    shards = [[] for _ in range(nshards)]
    for key, r in recs:
        shards[hash(key) % nshards].append((key, r))
    if nshards > 2:
        shards[1] = []   # an empty shard
        want = b"".join(r for _, r in sorted(x for s in shards for x in s))
    runs = [b"".join(r for _, r in sorted(s)) for s in shards]
    got = b"".join(p.tobytes() for p in merge_sorted_runs(runs, rb, W, chunk))
    assert got == want
    paths = []
    for i, run in enumerate(runs):
        p = tmp_path / ("shard%d.bin" % i)
        p.write_bytes(run)
        paths.append(str(p))
    out = tmp_path / "out.ctx"
    with open(out, "wb") as fh:
        nrec = write_ctx_from_shards(fh, b"HEADER", paths, rb, W, chunk)
    assert nrec == len(want) // rb and out.read_bytes() == b"HEADER" + want
    with _pytest.raises(ValueError):
        list(merge_sorted_runs([b"x" * (rb + 1)], rb, W))
