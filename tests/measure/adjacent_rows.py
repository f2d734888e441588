"""Measurements for the rows next to the hot path: utterance CMVN, SpecAugment fill, batched DTW.
Algorithmic bytes: utterance CMVN 8 B per element (read + write; the kernel reads twice, the second time from L2),
DTW 16 B per cell (distance in; cumulative distance, back pointer, path out).   python tests/measure/adjacent_rows.py"""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from oracle import dtw as odtw
pkg = importlib.import_module(bench.PKG)
mcd = importlib.import_module(bench.PKG + ".mcd")
dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6536.7


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
rng = np.random.RandomState(0)
frames = [int(t) for t in rng.randint(800, 2000, 2000)]
x = torch.randn(sum(frames), 80, device=dev)
uc = pkg.feature_transforms.get_audio_feature_transform("utterance_cmvn")()
ms = timeit(lambda: uc.apply_cuda(x, frames))
out["utterance_cmvn"] = {"utterances": len(frames), "rows": sum(frames), "ms": ms, "GBps_algorithmic": x.numel() * 8 / ms / 1e6,
                         "frac_of_hbm": x.numel() * 8 / ms / 1e6 / peak}
sa = pkg.feature_transforms.get_audio_feature_transform("specaugment").from_config_dict(
    {"freq_mask_N": 2, "freq_mask_F": 27, "time_mask_N": 2, "time_mask_T": 100, "time_mask_p": 1.0})
np.random.seed(0)
sa.apply_cuda(x, frames); torch.cuda.synchronize()  # first call: allocations, pinned staging
t0 = time.perf_counter(); y = sa.apply_cuda(x, frames); torch.cuda.synchronize()
out["specaugment_batch_2000_utts_ms_wall"] = 1e3 * (time.perf_counter() - t0)  # host mask drawing + clone + mean + fill
B, M, N = 64, 480, 470
d = torch.rand(B, M, N, device=dev)
ms = timeit(lambda: mcd.batch_dynamic_time_warping(d), n=5)
out["dtw"] = {"pairs": B, "M": M, "N": N, "ms": ms, "cells_per_s": B * M * N / ms * 1e3, "GBps_algorithmic": B * M * N * 16 / ms / 1e6}
t0 = time.perf_counter(); ref = odtw.batch_dynamic_time_warping_torch(d[:8]); torch.cuda.synchronize()
dt = time.perf_counter() - t0
mine = mcd.batch_dynamic_time_warping(d[:8].contiguous())
out["dtw_reference_formulation_torch_gpu"] = {"pairs": 8, "ms": 1e3 * dt, "cells_per_s": 8 * M * N / dt,
                                              "path_equal": bool(torch.equal(ref[2], mine[2]))}
print(json.dumps(out, indent=1))
