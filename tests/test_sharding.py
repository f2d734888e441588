"""Multi-GPU host logic on CPU: LPT sharding and the final ragged gather over gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_lpt_balances_fisher_like_lengths(pkg):
    sh = __import__("importlib").import_module(PKG + ".sharding")
    rng = np.random.RandomState(0)
    frames = rng.randint(56, 401, size=10000).tolist()
    for world in (2, 4, 8):
        owned = sh.shard_utterances(frames, world)
        assert sorted(i for o in owned for i in o) == list(range(10000))
        assert sh.imbalance(frames, owned) < 1.001
    # degenerate: fewer utterances than ranks
    owned = sh.shard_utterances([100, 50], 4)
    assert [len(o) for o in owned] == [1, 1, 0, 0]


def test_length_buckets(pkg):
    sh = __import__("importlib").import_module(PKG + ".sharding")
    frames = [400, 56, 300, 57, 60, 399]
    b = sh.length_buckets(frames, range(6), max_frames=500)
    assert sorted(i for x in b for i in x) == list(range(6))
    assert all(sum(frames[i] for i in x) <= 500 or len(x) == 1 for x in b)
    assert [frames[i] for x in b for i in x] == sorted(frames)


def _worker(rank, world, port, q):
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module(PKG + ".sharding")
    frames = [12, 7, 30, 5, 9, 21, 6]
    owned = sh.shard_utterances(frames, world, n_iter=4)
    mine = owned[rank]
    waves = [torch.full(((frames[i] - 1) * 3,), float(i)) + torch.arange((frames[i] - 1) * 3) * 1e-3 for i in mine]
    out = sh.gather_waveforms(mine, waves, len(frames), dst=0)
    if rank == 0:
        ok = all(out[i].shape == ((frames[i] - 1) * 3,) and float(out[i][0]) == float(i) for i in range(len(frames)))
        q.put(ok)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_waveforms_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True
