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
    # balanced: the number of batches is fixed first, then they are cut evenly (no small remainder batch)
    rng = np.random.RandomState(1)
    frames = rng.randint(56, 401, size=3000).tolist()
    total = sum(frames)
    b = sh.length_buckets(frames, range(3000), max_frames=300000, balanced=True)
    sizes = [sum(frames[i] for i in x) for x in b]
    assert len(b) == -(-total // 300000) and max(sizes) <= 300000 and min(sizes) > 0.9 * max(sizes)
    assert sorted(i for x in b for i in x) == list(range(3000))


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
    # the flat form (one tensor + lengths, what synthesize_flat returns) gives the same result
    out_flat = sh.gather_waveforms(mine, torch.cat(waves) if waves else torch.empty(0), len(frames), dst=0,
                                   local_lengths=[w.numel() for w in waves])
    if rank == 0:
        assert all(torch.equal(a, b) for a, b in zip(out, out_flat)) and len(out_flat.to_list()) == len(frames)
    # layout known to every rank (deterministic sharding): no metadata exchange; asynchronous start, wait() later
    layout = [(owned[r], [(frames[i] - 1) * 3 for i in owned[r]]) for r in range(world)]
    h = sh.gather_waveforms(mine, torch.cat(waves), len(frames), dst=0, local_lengths=layout[rank][1], layout=layout, async_op=True)
    res = h.wait()
    if rank == 0:
        assert all(torch.equal(res[i], out[i]) for i in range(len(frames)))
    else:
        assert res is None
    # a rank with nothing to send, and a non-zero destination
    mine2 = list(range(len(frames))) if rank == 0 else []
    waves2 = [torch.full((frames[i],), float(i)) for i in mine2]
    st = {}
    out2 = sh.gather_waveforms(mine2, waves2, len(frames), dst=1, stats=st)
    if rank == 1:
        assert all(out2[i].numel() == frames[i] and float(out2[i][-1]) == float(i) for i in range(len(frames)))
        assert st["bytes_to_dst"] == 4 * sum(frames)
    else:
        assert out2 is None and st["bytes_sent"] == 4 * sum(frames)
    dist.barrier()
    dist.destroy_process_group()


def _nccl_worker(rank, world, port, q):
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pkg = importlib.import_module(PKG)
    sh = importlib.import_module(PKG + ".sharding")
    from conftest import seeded_phase, synth_logmel
    voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=8).cuda()
    frames = [40, 9, 77, 56, 120, 33, 64]
    owned = sh.shard_utterances(frames, world, n_iter=8)
    mine = owned[rank]
    feats = [synth_logmel(frames[i], 700 + i).cuda() for i in mine]
    phases = [seeded_phase(800 + i, frames[i]) for i in mine]
    waves = voc.synthesize_batch(feats, init_phase=phases, n_iter=8)
    out = sh.gather_waveforms(mine, waves, len(frames), dst=0)
    if rank == 0:
        # every utterance, wherever it was synthesised, must equal the single-GPU result bitwise (small calls run the
        # frame-parallel kernel, whose results do not depend on the batch)
        ok = True
        for i, T in enumerate(frames):
            alone = voc.synthesize_batch([synth_logmel(T, 700 + i).cuda()], init_phase=[seeded_phase(800 + i, T)], n_iter=8)[0]
            ok = ok and out[i].is_cuda and out[i].shape == ((T - 1) * 300,) and bool(torch.equal(out[i], alone))
        q.put(ok)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_synthesis_and_gather_nccl_world2(built_lib):
    """Two GPUs: LPT sharding -> per-rank ragged synthesis -> point-to-point gather over NCCL (needs >= 2 devices;
    run with `gpurun --gpus 2`)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get() is True


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
