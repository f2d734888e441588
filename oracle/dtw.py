"""CPU restatement of the DTW / distance part of the MCD metric (numpy float32, plain loops: small cases only).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

* ``batch_dynamic_time_warping``  examples/s2s_trans/tasks/s2s_translation.py:414-464
* ``compute_rms_dist``            examples/s2s_trans/tasks/s2s_translation.py:467-475
"""
import numpy as np


def batch_dynamic_time_warping(distance, shapes=None):
    """distance [B, M, N] float32 -> (cumdist float32, backptr int32, pathmap int32).  Pointers: 0 = left, 1 = up-left,
    2 = up; ties go to the first of (left, up-left, up), like torch's min over the stacked candidates on the CPU."""
    distance = np.asarray(distance, np.float32)
    B, M, N = distance.shape
    cum = np.zeros_like(distance)
    bp = np.full(distance.shape, -1, np.int32)
    path = np.zeros(distance.shape, np.int32)
    for b in range(B):
        d = distance[b]
        # first row / column: torch.cumsum on the CPU sums sequentially in a double accumulator and rounds every
        # prefix to float32
        cum[b, 0, :] = np.cumsum(d[0, :].astype(np.float64)).astype(np.float32)
        cum[b, :, 0] = np.cumsum(d[:, 0].astype(np.float64)).astype(np.float32)
        bp[b, 0, :] = 0
        bp[b, :, 0] = 2
        for i in range(1, M):
            for j in range(1, N):
                cands = (cum[b, i, j - 1], cum[b, i - 1, j - 1], cum[b, i - 1, j])
                k = int(np.argmin(cands))   # first minimum
                bp[b, i, j] = k
                cum[b, i, j] = np.float32(cands[k] + d[i, j])
        i = M - 1 if shapes is None else int(shapes[b][0]) - 1
        j = N - 1 if shapes is None else int(shapes[b][1]) - 1
        n = 1
        path[b, i, j] = 1
        while (i != 0 or j != 0) and n < 10000:
            di, dj = {0: (0, -1), 1: (-1, -1), 2: (-1, 0)}[int(bp[b, i, j])]
            i, j = i + di, j + dj
            path[b, i, j] = 1
            n += 1
    return cum, bp, path


def compute_rms_dist(x1, x2):
    x1, x2 = np.asarray(x1, np.float64), np.asarray(x2, np.float64)
    d2 = ((x1[:, None, :] - x2[None, :, :]) ** 2).sum(-1)
    return np.sqrt(d2 / x1.shape[1]).astype(np.float32)


def batch_dynamic_time_warping_torch(distance, shapes=None):
    """The reference's FORMULATION of the same recurrence (s2s_translation.py:414-464) in torch: one round of gathers /
    min / scatter per anti-diagonal and a host back trace with one .item() per step -- kept for timing the reference's
    approach on the GPU next to the fused kernel (tests/measure/adjacent_rows.py); results equal the loops above."""
    import torch
    bsz, m, n = distance.shape
    cum = torch.zeros_like(distance)
    bp = torch.full(distance.shape, -1, dtype=torch.int32, device=distance.device)
    cum[:, 0, :] = distance[:, 0, :].cumsum(-1)
    cum[:, :, 0] = distance[:, :, 0].cumsum(-1)
    bp[:, 0, :] = 0
    bp[:, :, 0] = 2
    for off in range(2, m + n - 1):
        j = torch.arange(max(1, off - m + 1), min(n, off), device=distance.device)
        i = off - j
        cand = torch.stack([cum[:, i, j - 1], cum[:, i - 1, j - 1], cum[:, i - 1, j]], dim=2)
        v, b = cand.min(dim=-1)
        bp[:, i, j] = b.int()
        cum[:, i, j] = v + distance[:, i, j]
    path = torch.zeros_like(bp)
    step = {0: (0, -1), 1: (-1, -1), 2: (-1, 0)}
    for b in range(bsz):
        i = m - 1 if shapes is None else int(shapes[b][0]) - 1
        j = n - 1 if shapes is None else int(shapes[b][1]) - 1
        cells = [(i, j)]
        while (i != 0 or j != 0) and len(cells) < 10000:
            di, dj = step[bp[b, i, j].item()]
            i, j = i + di, j + dj
            cells.append((i, j))
        idx = torch.tensor(cells, device=distance.device)
        path[b, idx[:, 0], idx[:, 1]] = 1
    return cum, bp, path
