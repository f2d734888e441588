"""One small synthesis call (T frames, 8 iterations) for an ncu launch list: python tools/small_probe.py T team(0/1)"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
T, team = int(sys.argv[1]), int(sys.argv[2])
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=8).cuda()
plan = voc._plan(torch.device("cuda", 0))
plan.set_option(pkg._lib.OPT_GL_TEAM, team)
x = torch.from_numpy(bench.synth_logmel_np(T, 1)).cuda()
ph = ((torch.rand(T, 1025, device="cuda") * 2 - 1) * np.pi).contiguous()
for _ in range(3):
    y = voc.synthesize_flat(x, [T], ph)
torch.cuda.synchronize()
plan.set_pass_timing(True)
y = voc.synthesize_flat(x, [T], ph)
print("pass times (events, us):", " ".join(f"{1e3 * t:.1f}" for t in plan.pass_times_ms()))
