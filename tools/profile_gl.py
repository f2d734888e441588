"""Small driver for ncu captures of the Griffin-Lim pass kernel: config-2-shaped batch, few iterations."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 4
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=n_iter).cuda()
frames = bench.batch_frames(0)
total = sum(frames)
x = torch.from_numpy(np.concatenate([bench.synth_logmel_np(T, 1234 + i) for i, T in enumerate(frames)])).cuda()
ph = (torch.rand(total, 1025, device="cuda") * 2 - 1) * np.pi
for _ in range(2):
    y = voc.synthesize_flat(x, frames, ph)
torch.cuda.synchronize()
print("ok", y.shape, float(y.abs().mean()))
