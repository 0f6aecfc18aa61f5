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
