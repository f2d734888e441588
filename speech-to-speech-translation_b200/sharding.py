"""Utterance sharding across GPUs: one process per GPU, no collective on the hot path.

The reference shards synthesis with ``--num-shards N --shard-id i`` (independent processes,
examples/s2s_trans/generate_waveform.py:166-167) after sorting by source length
(fairseq/data/audio/speech_to_text_dataset.py:358-366).  Here utterances are dealt by work
(frames x iterations) with longest-processing-time-first, bucketed by length inside a rank, and the
only exchange is the final ragged gather of waveforms to ONE rank (``gather_waveforms``: exact-size point-to-point
messages in one group, no padded all-gather), which works on any torch.distributed backend (NCCL on GPUs, gloo in the
CPU tests).
"""
from typing import List, Optional, Sequence

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


def length_buckets(n_frames: Sequence[int], indices: Sequence[int], max_frames: int, balanced: bool = False) -> List[List[int]]:
    """Sort ``indices`` by length and cut into batches of at most ``max_frames`` frames (>= 1 utterance).

    balanced: first fix the number of batches (ceil(total / max_frames)), then cut them to (nearly) equal frame
    counts, so that no small remainder batch is left to under-fill the GPU."""
    order = sorted(indices, key=lambda i: (n_frames[i], i))
    limit = max_frames
    if balanced and order:
        total = sum(n_frames[i] for i in order)
        n_batches = max(1, -(-total // max_frames))
        limit = min(max_frames, -(-total // n_batches) + max(n_frames[i] for i in order))
    batches, cur, cur_frames = [], [], 0
    for i in order:
        if cur and cur_frames + n_frames[i] > limit:
            batches.append(cur)
            cur, cur_frames = [], 0
        cur.append(i)
        cur_frames += n_frames[i]
    if cur:
        batches.append(cur)
    return batches


class GatheredWaveforms:
    """What ``gather_waveforms`` hands to the destination rank: the payload of every rank in one receive buffer per
    rank plus the (utterance id, length) tables.  Indexing by global utterance id returns a view into the buffer;
    views are created on demand (materialising ten thousand tensor views is host work that has nothing to do with
    the transfer).  ``wait()`` makes the current stream wait for the transfer when it was started with ``async_op``."""

    def __init__(self, n_total, payloads, ids, lens, reqs=()):
        self.n_total, self._payloads, self._reqs = n_total, payloads, list(reqs)
        self._where = {}
        for r, (ii, ll) in enumerate(zip(ids, lens)):
            off = 0
            for i, n in zip(ii, ll):
                self._where[int(i)] = (r, off, int(n))
                off += int(n)

    def wait(self):
        for req in self._reqs:
            req.wait()
        self._reqs = []
        return self

    def __len__(self):
        return self.n_total

    def __getitem__(self, i):
        self.wait()
        r, off, n = self._where[int(i)]
        return self._payloads[r][off: off + n]

    def __iter__(self):
        return (self[i] for i in range(self.n_total))

    def to_list(self):
        return [self[i] if i in self._where else None for i in range(self.n_total)]


class _Pending:
    """Handle of an asynchronous gather on a rank that only sends."""

    def __init__(self, reqs, keep):
        self._reqs, self._keep = list(reqs), keep

    def wait(self):
        for req in self._reqs:
            req.wait()
        self._reqs, self._keep = [], None
        return None


def gather_waveforms(local_ids: Sequence[int], local_waves, n_total: int, dst: int = 0,
                     group=None, stats: Optional[dict] = None, local_lengths: Optional[Sequence[int]] = None,
                     layout=None, async_op: bool = False):
    """Final ragged gather: rank ``dst`` receives every utterance's waveform, addressable by global utterance id.

    The only exchange of the multi-GPU path (the reference's shards each write their own files,
    generate_waveform.py:166-167).  The samples go point to point to ``dst`` ONLY -- every rank sends one exact-size
    payload (no padding), ``dst`` receives them into one buffer per rank, all in one batched send/recv group (NCCL:
    ncclGroupStart/End; gloo in the CPU tests).

    * ``local_waves``: a list of per-utterance tensors, or -- cheaper for thousands of utterances -- ONE flat tensor with
      the utterances back to back and ``local_lengths`` giving their sizes (what ``synthesize_flat`` returns).
    * ``layout``: optional ``[(ids_r, lengths_r) for every rank r]`` when all ranks already know who owns what (a
      deterministic sharding of a known list): the (count, size) all_gather and the id messages are skipped, and nothing
      in the call synchronises the host.  Without it one small all_gather of sizes and one int64 message per rank travel.
    * ``async_op``: start the transfer and return at once; call ``.wait()`` on the result before using it (``dst``) or
      before freeing / overwriting the sent tensor (other ranks).  This is how the transfer of one bucket overlaps the
      synthesis of the next.

    Returns a ``GatheredWaveforms`` (indexable by utterance id, ``to_list()``) on ``dst``; ``None`` (or a pending
    handle with ``async_op``) elsewhere.  ``stats`` (optional dict) receives the bytes that cross the fabric."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nccl = dist.get_backend(group) == "nccl"
    flat_in = isinstance(local_waves, torch.Tensor)
    if flat_in:
        assert local_lengths is not None and int(sum(local_lengths)) == local_waves.numel()
        lens_local = [int(n) for n in local_lengths]
        device = local_waves.device
    else:
        lens_local = [int(w.numel()) for w in local_waves]
        device = local_waves[0].device if len(local_waves) else (
            torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu"))
    n_local = len(local_ids)
    size_local = int(sum(lens_local))
    if flat_in:
        flat = local_waves.reshape(-1).float()
    else:
        flat = (torch.cat([w.reshape(-1).float() for w in local_waves]) if n_local
                else torch.empty(0, dtype=torch.float32, device=device))
    ops, recv_idx, recv_payload = [], {}, {}
    if layout is not None:
        assert len(layout) == world and list(layout[rank][0]) == list(local_ids)
        counts = [len(ii) for ii, _ in layout]
        sizes = [int(sum(ll)) for _, ll in layout]
        idx = None
    else:
        meta = torch.tensor([n_local, size_local], dtype=torch.int64, device=device)
        metas = [torch.zeros_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta, group=group)
        counts = [int(m[0]) for m in metas]
        sizes = [int(m[1]) for m in metas]
        # (id, length) pairs are a second, tiny int64 message of the same group
        idx = torch.empty(2, n_local, dtype=torch.int64, device=device)
        if n_local:
            idx.copy_(torch.tensor([list(local_ids), lens_local], dtype=torch.int64))
    if rank == dst:
        for r in range(world):
            if r == dst:
                continue
            recv_payload[r] = torch.empty(sizes[r], dtype=torch.float32, device=device)
            if layout is None:
                recv_idx[r] = torch.empty(2, counts[r], dtype=torch.int64, device=device)
                if counts[r]:
                    ops.append(dist.P2POp(dist.irecv, recv_idx[r], _global_rank(r, group), group))
            if sizes[r]:
                ops.append(dist.P2POp(dist.irecv, recv_payload[r], _global_rank(r, group), group))
    else:
        if layout is None and n_local:
            ops.append(dist.P2POp(dist.isend, idx, _global_rank(dst, group), group))
        if size_local:
            ops.append(dist.P2POp(dist.isend, flat, _global_rank(dst, group), group))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    if stats is not None:
        stats["bytes_to_dst"] = 4 * sum(sizes[r] for r in range(world) if r != dst)
        stats["bytes_sent"] = 0 if rank == dst else 4 * size_local
    if rank != dst:
        pending = _Pending(reqs, (flat, idx))
        return pending if async_op else pending.wait()
    if layout is None:
        for req in reqs:  # the id tables are needed on the host now
            req.wait()
        reqs = []
    recv_payload[dst] = flat
    ids, lens = [], []
    for r in range(world):
        if r == dst:
            ids.append(list(local_ids))
            lens.append(lens_local)
        elif layout is not None:
            ids.append(list(layout[r][0]))
            lens.append(list(layout[r][1]))
        else:
            ii, ll = recv_idx[r].tolist()  # one download per sending rank
            ids.append(ii)
            lens.append(ll)
    out = GatheredWaveforms(n_total, [recv_payload[r] for r in range(world)], ids, lens, reqs)
    return out if async_op else out.wait()


def _global_rank(group_rank: int, group) -> int:
    return group_rank if group is None else dist.get_global_rank(group, group_rank)


def imbalance(n_frames: Sequence[int], owned: List[List[int]]) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    loads = np.array([sum(n_frames[i] for i in o) for o in owned], dtype=np.float64)
    return float(loads.max() / max(loads.mean(), 1e-30))
