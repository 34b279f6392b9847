"""The multi-GPU exchange logic on CPU: world_size 2, gloo backend.

modimizer_b200.dist.exchange() (the variable all-to-all used by ShardedModset)
is backend-agnostic; here each rank selects the modimizers of ITS chunk of the
reads with the host build of the kernel arithmetic (test infrastructure),
buckets them by owner exactly as partition.cu does (mg_owner), exchanges them
over gloo and counts what it owns.  The union of the two shards must be the
oracle's single modset, bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import harness as H
    import hostemul as he
    from modimizer_b200.dist import exchange
    k, d, seed = 19, 31, 17
    orc = H.oracle()
    f1 = orc.hasher(k, d, seed)["factor1"]
    sp = he.read_spec(12345, 50000, 3, 1000)
    nreads = 400
    per = nreads // world
    data = he.reads(sp, rank * per, per)
    offs = np.arange(per + 1, dtype=np.uint64) * np.uint64(1000)
    km, _, _ = he.select(k, d, f1, data, offs)
    owner = np.array([he.lib().hm_owner(int(x), world) for x in km], np.int64)
    order = np.argsort(owner, kind="stable")
    counts = np.bincount(owner, minlength=world)
    send = torch.from_numpy(km[order].view(np.int64).copy())
    recv, rc = exchange(send, counts.tolist())
    mine = recv.numpy().view(np.uint64)
    assert all(he.lib().hm_owner(int(x), world) == rank for x in mine[:200])
    vals, cnts = np.unique(mine, return_counts=True)
    np.savez(os.path.join(tmpdir, "shard%d.npz" % rank), v=vals, c=cnts, sent=len(km))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_count_equals_single(tmp_path, orc):
    world = 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    import hostemul as he
    sp = he.read_spec(12345, 50000, 3, 1000)
    data = he.reads(sp, 0, 400)
    offs = np.arange(401, dtype=np.uint64) * np.uint64(1000)
    ms = orc.modset_new(20, 19, 31, 17)
    tot = orc.modset_add(ms, data, offs)
    ov, od, _ = orc.modset_sorted(ms)
    shards = [np.load(os.path.join(str(tmp_path), "shard%d.npz" % r)) for r in range(world)]
    assert sum(int(s["sent"]) for s in shards) == tot
    v = np.concatenate([s["v"] for s in shards]); c = np.concatenate([s["c"] for s in shards])
    o = np.argsort(v)
    assert np.array_equal(v[o], ov) and np.array_equal(np.minimum(c[o], 65535).astype(np.uint16), od)
    assert len(set(shards[0]["v"].tolist()) & set(shards[1]["v"].tolist())) == 0      # disjoint ownership
    orc._modset_free(ms)


def test_exchange_single_rank():
    """world_size 1 degenerates to a copy (no collective needed)"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(31500 + os.getpid() % 2000)
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        from modimizer_b200.dist import exchange
        send = torch.arange(10, dtype=torch.int64)
        recv, rc = exchange(send, [7])
        assert rc == [7] and recv.tolist() == list(range(7))
    finally:
        dist.destroy_process_group()
