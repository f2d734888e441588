"""Front-end throughput on one B200 (BASELINE config 3 shape, scaled by --utts): fbank80 (+fused CMVN),
logmelspec80 (+fused CMVN), stand-alone CMVN.  Prints one JSON object; algorithmic bytes per frame from
SURVEY 8(d): fbank 960 B, logmel 1520 B, CMVN 640 B."""
import argparse, importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
ap = argparse.ArgumentParser(); ap.add_argument("--utts", type=int, default=2000); ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
hbm, _ = bench.peaks()
rng = np.random.RandomState(0)

def timeit(fn, n):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

out = {}
mean = torch.randn(80, device=dev) - 4; std = torch.rand(80, device=dev) * 1.5 + 0.5
for name, sr in (("fbank80_16k", 16000), ("fbank80_8k", 8000)):
    lens = (rng.uniform(8, 20, a.utts) * sr).astype(np.int64)
    waves = [(torch.randn(int(n), device=dev) * 3000) for n in lens]
    plan = importlib.import_module(bench.PKG + ".plans").get_fbank_plan(dev, sr, 80)
    frames = [1 + (int(n) - plan.win) // plan.shift for n in lens]
    flat = torch.cat(waves).contiguous()
    fo = torch.from_numpy(np.concatenate([[0], np.cumsum(frames)]).astype(np.int32)).to(dev)
    wo = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
    total = int(sum(frames)); o = torch.empty(total, 80, device=dev)
    lib = pkg._lib.load()
    def run():
        pkg._lib.check(lib.s2st_fbank(plan.handle, a.utts, total, pkg._lib.ptr(wo), pkg._lib.ptr(fo), pkg._lib.ptr(flat),
                                      pkg._lib.ptr(mean), pkg._lib.ptr(std), None, pkg._lib.ptr(o), pkg._lib.stream_ptr(dev)), "fbank")
    ms = timeit(run, a.steps)
    audio = float(lens.sum()) / sr
    out[name] = {"audio_s_per_s": audio / (ms * 1e-3), "ms": ms, "frames": total,
                 "GBps_algorithmic": total * 960 / (ms * 1e-3) / 1e9, "frac_of_hbm": total * 960 / (ms * 1e-3) / 1e9 / hbm}
    del waves, flat, o
# logmelspec80 @ 24 kHz
sr = 24000
lens = (rng.uniform(8, 20, a.utts) * sr).astype(np.int64)
flat = torch.rand(int(lens.sum()), device=dev) * 0.2 - 0.1
frames = [1 + int(n) // 300 for n in lens]; total = int(sum(frames))
plans = importlib.import_module(bench.PKG + ".plans")
plan = plans.get_stft_plan(dev, 2048, 1200, 300, 80, torch.hann_window(1200), mel=pkg.get_mel_filters(24000, 2048, 80, 20, 8000))
fo = torch.from_numpy(np.concatenate([[0], np.cumsum(frames)]).astype(np.int32)).to(dev)
wo = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
o = torch.empty(total, 80, device=dev)
lib = pkg._lib.load()
def run():
    pkg._lib.check(lib.s2st_logmel(plan.handle, a.utts, total, pkg._lib.ptr(wo), pkg._lib.ptr(fo), pkg._lib.ptr(flat), 1e-5,
                                   pkg._lib.ptr(mean), pkg._lib.ptr(std), None, pkg._lib.ptr(o), pkg._lib.stream_ptr(dev)), "logmel")
ms = timeit(run, a.steps)
out["logmelspec80_24k"] = {"audio_s_per_s": float(lens.sum()) / sr / (ms * 1e-3), "ms": ms, "frames": total,
                           "GBps_algorithmic": total * 1520 / (ms * 1e-3) / 1e9, "frac_of_hbm": total * 1520 / (ms * 1e-3) / 1e9 / hbm}
# stand-alone CMVN on the features just produced
cm = importlib.import_module(bench.PKG + ".feature_transforms.global_cmvn")
o2 = torch.empty_like(o)
ms = timeit(lambda: cm.cmvn_apply_cuda(o, mean, std, out=o2), a.steps)
out["global_cmvn"] = {"ms": ms, "frames": total, "GBps_algorithmic": total * 640 / (ms * 1e-3) / 1e9,
                      "frac_of_hbm": total * 640 / (ms * 1e-3) / 1e9 / hbm}
# CPU reference for fbank (torchaudio, what the reference calls) on a few utterances
import torchaudio.compliance.kaldi as K
w = (torch.randn(16000 * 14) * 3000)[None]
t0 = time.perf_counter()
for _ in range(5): K.fbank(w, num_mel_bins=80, sample_frequency=16000)
out["cpu_torchaudio_fbank_16k_audio_s_per_s"] = 5 * 14 / (time.perf_counter() - t0)
out["utts"] = a.utts; out["hbm_peak_GBps"] = hbm
print(json.dumps(out))
