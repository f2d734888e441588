"""Single-utterance pass time as a function of the pinned strip length: slope = time of one frame for a warp
that has the SM (almost) to itself, intercept = launch + strip prologue / epilogue."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=32).cuda()
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = torch.from_numpy(bench.synth_logmel_np(T, 1)).cuda()
plan = voc._plan(x.device)
plan.set_pass_timing(True)
for S in (4, 8, 16, 32, 64):
    plan.set_strip_frames(S)
    for _ in range(3): voc.synthesize_flat(x, [T], None, seed=1, n_iter=32)
    torch.cuda.synchronize()
    t = np.asarray(plan.pass_times_ms())[1:]
    print(f"T={T} S={S:3d} strips={-(-T // S):4d}  pass median {np.median(t) * 1e3:.1f} us  min {t.min() * 1e3:.1f} us")
