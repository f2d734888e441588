"""Latency of small synthesis calls (config 1: one 500-frame utterance; a 100-frame one; one 60 s utterance; 8, 32 and 128
utterances of 230 frames), device-resident: the strip kernels (one launch per iteration) against the frame-parallel
single-launch kernel (gl_frames.cuh, the default for calls of up to a few thousand frames)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
plan = voc._plan(torch.device("cuda", 0))
cases = [("T=100", [100]), ("T=500", [500]), ("T=1000", [1000]), ("T=2300", [2300]), ("T=4800", [4800]), ("8x230", [230] * 8),
         ("16x230", [230] * 16), ("32x230", [230] * 32), ("40x230", [230] * 40), ("64x230", [230] * 64), ("80x230", [230] * 80), ("100x230", [230] * 100), ("128x230", [230] * 128)]
if len(sys.argv) > 1:
    cases = [c for c in cases if c[0] in sys.argv[1:]]
for name, frames in cases:
    total = sum(frames)
    x = torch.from_numpy(np.concatenate([bench.synth_logmel_np(T, 1 + i) for i, T in enumerate(frames)])).cuda()
    ph = ((torch.rand(total, 1025, device="cuda") * 2 - 1) * np.pi).contiguous()
    res, outs = [], []
    for mode in (0, 2):  # 0: strip kernels, one launch per iteration; 2: frame-parallel kernel
        plan.set_option(pkg._lib.OPT_GL_FRAMES, 1 << 20 if mode == 2 else 0)
        try:
            for _ in range(3): y = voc.synthesize_flat(x, frames, ph)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): y = voc.synthesize_flat(x, frames, ph)
            e1.record(); torch.cuda.synchronize()
            res.append((e0.elapsed_time(e1) / 10, plan.gl_launch_count(64)))
            outs.append(y)
        except Exception as e:  # noqa
            res.append((float("nan"), -1)); outs.append(None); print("  mode", mode, "failed:", e)
    audio = (total - len(frames)) * 300 / 24000
    rel = float((outs[1] - outs[0]).norm() / outs[0].norm()) if outs[1] is not None and outs[0] is not None else float("nan")
    print(f"{name:8s} strips {res[0][0]:.3f} ms ({audio / res[0][0] * 1e3:.0f} audio-s/s) | "
          f"frame-parallel {res[1][0]:.3f} ms ({audio / res[1][0] * 1e3:.0f}, {res[1][1]} launches) | "
          f"rel-L2 frame-parallel vs strips {rel:.2e}", flush=True)
plan.set_option(pkg._lib.OPT_GL_FRAMES, 16 * 148 * 4)
