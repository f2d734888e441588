"""Small driver for ncu captures of the two tensor-core projections on the config-2 batch (58 043 frames)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
dev = torch.device("cuda", 0)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=1).cuda()
frames, logmel, _ = bench.config2_batch(0)
total = int(sum(frames))
lm = torch.from_numpy(logmel).cuda()
plan = voc._plan(dev)
lib, ptr, sptr = pkg._lib.load(), pkg._lib.ptr, pkg._lib.stream_ptr
mag = torch.empty(total, 1025, device=dev)
plans = importlib.import_module(bench.PKG + ".plans")
mplan = plans.get_stft_plan(dev, 2048, 2048, 512, 80, torch.ones(2048), mel=pkg.get_mel_filters(24000, 2048, 80, 20.0, 8000.0))
out = torch.empty(total, 80, device=dev)
for _ in range(3):
    pkg._lib.check(lib.s2st_inverse_mel(plan.handle, total, ptr(lm), 1, ptr(mag), sptr(dev)), "inverse_mel")
    pkg._lib.check(lib.s2st_mel_project(mplan.handle, total, ptr(mag), ptr(out), sptr(dev)), "mel_project")
torch.cuda.synchronize()
print("ok", float(mag.mean()), float(out.mean()))
