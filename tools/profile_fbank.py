"""Small driver for ncu captures of the fbank kernel: 400 utterances of 8-20 s at the given rate."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
sr = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
utts = int(sys.argv[2]) if len(sys.argv) > 2 else 400
dev = torch.device("cuda", 0)
rng = np.random.RandomState(0)
lens = (rng.uniform(8, 20, utts) * sr).astype(np.int64)
flat = torch.randn(int(lens.sum()), device=dev) * 3000
plan = importlib.import_module(bench.PKG + ".plans").get_fbank_plan(dev, sr, 80)
frames = [1 + (int(n) - plan.win) // plan.shift for n in lens]
fo = torch.from_numpy(np.concatenate([[0], np.cumsum(frames)]).astype(np.int32)).to(dev)
wo = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
total = int(sum(frames)); o = torch.empty(total, 80, device=dev)
mean = torch.randn(80, device=dev) - 4; std = torch.rand(80, device=dev) * 1.5 + 0.5
lib = pkg._lib.load()
for _ in range(3):
    pkg._lib.check(lib.s2st_fbank(plan.handle, utts, total, pkg._lib.ptr(wo), pkg._lib.ptr(fo), pkg._lib.ptr(flat),
                                  pkg._lib.ptr(mean), pkg._lib.ptr(std), None, pkg._lib.ptr(o), pkg._lib.stream_ptr(dev)), "fbank")
torch.cuda.synchronize()
print("ok frames", total, float(o.mean()))
