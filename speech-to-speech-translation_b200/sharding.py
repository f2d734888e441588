"""Utterance sharding across GPUs: one process per GPU, no collective on the hot path.

The reference shards synthesis with ``--num-shards N --shard-id i`` (independent processes,
examples/s2s_trans/generate_waveform.py:166-167) after sorting by source length
(fairseq/data/audio/speech_to_text_dataset.py:358-366).  Here utterances are dealt by work
(frames x iterations) with longest-processing-time-first, bucketed by length inside a rank, and the
only exchange is the final ragged gather of waveforms (``gather_waveforms``), which works on any
torch.distributed backend (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def lpt_assign(costs: Sequence[float], n_ranks: int) -> List[List[int]]:
    """Longest-processing-time-first: returns, per rank, the indices it owns (deterministic)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * n_ranks
    owned = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: (loads[k], k))
        owned[r].append(i)
        loads[r] += costs[i]
    return owned


def shard_utterances(n_frames: Sequence[int], world_size: int, n_iter: int = 64) -> List[List[int]]:
    """Assign utterances to ranks by Griffin-Lim work T_i * (n_iter + 1)."""
    return lpt_assign([t * (n_iter + 1) for t in n_frames], world_size)


def length_buckets(n_frames: Sequence[int], indices: Sequence[int], max_frames: int) -> List[List[int]]:
    """Sort ``indices`` by length and cut into batches of at most ``max_frames`` frames (>= 1 utterance)."""
    order = sorted(indices, key=lambda i: (n_frames[i], i))
    batches, cur, cur_frames = [], [], 0
    for i in order:
        if cur and cur_frames + n_frames[i] > max_frames:
            batches.append(cur)
            cur, cur_frames = [], 0
        cur.append(i)
        cur_frames += n_frames[i]
    if cur:
        batches.append(cur)
    return batches


def gather_waveforms(local_ids: Sequence[int], local_waves: Sequence[torch.Tensor], n_total: int, dst: int = 0,
                     group=None):
    """Final ragged gather: rank ``dst`` receives every utterance's waveform in global order.

    Lengths travel with an all_gather; the samples with one padded all_gather (a single collective,
    96 KB per audio-second -- off the critical path).  Returns a list of n_total tensors on ``dst``,
    ``None`` elsewhere.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    device = local_waves[0].device if len(local_waves) else torch.device(
        "cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    meta = torch.tensor([len(local_ids), int(sum(w.numel() for w in local_waves))], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    counts = [int(m[0]) for m in metas]
    sizes = [int(m[1]) for m in metas]
    max_count, max_size = max(counts + [1]), max(sizes + [1])
    idx = torch.full((2, max_count), -1, dtype=torch.int64, device=device)
    if len(local_ids):
        idx[0, : len(local_ids)] = torch.as_tensor(list(local_ids), dtype=torch.int64)
        idx[1, : len(local_ids)] = torch.as_tensor([w.numel() for w in local_waves], dtype=torch.int64)
    payload = torch.zeros(max_size, dtype=torch.float32, device=device)
    if len(local_waves):
        flat = torch.cat([w.reshape(-1).float() for w in local_waves])
        payload[: flat.numel()] = flat
    all_idx = [torch.zeros_like(idx) for _ in range(world)]
    all_payload = [torch.zeros_like(payload) for _ in range(world)]
    dist.all_gather(all_idx, idx, group=group)
    dist.all_gather(all_payload, payload, group=group)
    if rank != dst:
        return None
    out = [None] * n_total
    for r in range(world):
        ids = all_idx[r][0, : counts[r]].tolist()
        lens = all_idx[r][1, : counts[r]].tolist()
        off = 0
        for i, n in zip(ids, lens):
            out[i] = all_payload[r][off: off + n]
            off += n
    return out


def imbalance(n_frames: Sequence[int], owned: List[List[int]]) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    loads = np.array([sum(n_frames[i] for i in o) for o in owned], dtype=np.float64)
    return float(loads.max() / max(loads.mean(), 1e-30))
