"""Time the Griffin-Lim iteration kernel on the config-2 batch: median per-pass ms over a 16-iteration call,
optionally for several pinned strip lengths.   python tools/time_pass.py [S ...]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=16).cuda()
frames = bench.batch_frames(0)
total = sum(frames)
x = torch.from_numpy(np.concatenate([bench.synth_logmel_np(T, 1234 + i) for i, T in enumerate(frames)])).cuda()
ph = (torch.rand(total, 1025, device="cuda") * 2 - 1) * np.pi
plan = voc._plan(x.device)
plan.set_pass_timing(True)
for S in [int(a) for a in sys.argv[1:]] or [0]:
    plan.set_strip_frames(S)
    for _ in range(3):
        y = voc.synthesize_flat(x, frames, ph, n_iter=16)
    torch.cuda.synchronize()
    t = np.asarray(plan.pass_times_ms())
    algo = 6500.0 * total
    if os.environ.get("ALL_PASSES"):
        print(" ".join(f"{v:.4f}" for v in t))
    print(f"S={S:3d}  first {t[0]:.4f} ms  iteration median {np.median(t[1:]):.4f} ms  min {t[1:].min():.4f}  "
          f"-> {algo / np.median(t[1:]) / 1e6:.0f} GB/s algorithmic, {total / np.median(t[1:]) * 1e-3:.1f} M frame-iter/s")
