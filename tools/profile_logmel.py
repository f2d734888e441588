"""Small driver for ncu captures of the logmelspec80 kernel: 400 utterances of 8-20 s at 24 kHz."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
utts = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dev = torch.device("cuda", 0)
rng = np.random.RandomState(0)
sr = 24000
lens = (rng.uniform(8, 20, utts) * sr).astype(np.int64)
flat = torch.rand(int(lens.sum()), device=dev) * 0.2 - 0.1
frames = [1 + int(n) // 300 for n in lens]; total = int(sum(frames))
plans = importlib.import_module(bench.PKG + ".plans")
plan = plans.get_stft_plan(dev, 2048, 1200, 300, 80, torch.hann_window(1200), mel=pkg.get_mel_filters(24000, 2048, 80, 20, 8000))
fo = torch.from_numpy(np.concatenate([[0], np.cumsum(frames)]).astype(np.int32)).to(dev)
wo = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
o = torch.empty(total, 80, device=dev)
mean = torch.randn(80, device=dev) - 4; std = torch.rand(80, device=dev) * 1.5 + 0.5
lib = pkg._lib.load()
for _ in range(3):
    pkg._lib.check(lib.s2st_logmel(plan.handle, utts, total, pkg._lib.ptr(wo), pkg._lib.ptr(fo), pkg._lib.ptr(flat), 1e-5,
                                   pkg._lib.ptr(mean), pkg._lib.ptr(std), None, pkg._lib.ptr(o), pkg._lib.stream_ptr(dev)), "logmel")
torch.cuda.synchronize()
print("ok frames", total, float(o.mean()))
