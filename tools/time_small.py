"""Latency of small synthesis calls (config 1: one 500-frame utterance; a 100-frame one; one 60 s utterance; 8, 32 and 128
utterances of 230 frames), device-resident: one warp per strip, the persistent single launch, and the team mode."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
plan = voc._plan(torch.device("cuda", 0))
cases = [("T=100", [100]), ("T=500", [500]), ("T=4800", [4800]), ("8x230", [230] * 8), ("32x230", [230] * 32), ("128x230", [230] * 128)]
for name, frames in cases:
    total = sum(frames)
    x = torch.from_numpy(np.concatenate([bench.synth_logmel_np(T, 1 + i) for i, T in enumerate(frames)])).cuda()
    ph = ((torch.rand(total, 1025, device="cuda") * 2 - 1) * np.pi).contiguous()
    res, outs = [], []
    for mode in (0, 1, 2):  # 0: one warp per strip, one launch per iteration; 1: + persistent launch; 2: team mode (default)
        plan.set_option(pkg._lib.OPT_GL_TEAM, 1 if mode == 2 else 0)
        plan.set_option(pkg._lib.OPT_GL_PERSISTENT, 1 if mode == 1 else 0)
        for _ in range(3): y = voc.synthesize_flat(x, frames, ph)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): y = voc.synthesize_flat(x, frames, ph)
        e1.record(); torch.cuda.synchronize()
        res.append((e0.elapsed_time(e1) / 10, plan.gl_launch_count(64)))
        outs.append(y)
    audio = (total - len(frames)) * 300 / 24000
    print(f"{name:8s} one warp per strip {res[0][0]:.3f} ms ({audio / res[0][0] * 1e3:.0f} audio-s/s) | persistent launch {res[1][0]:.3f} ms "
          f"({audio / res[1][0] * 1e3:.0f}, {res[1][1]} launches) | team of 4 warps per strip (default for small calls) {res[2][0]:.3f} ms "
          f"({audio / res[2][0] * 1e3:.0f}) | bitwise equal {bool(torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]))}")
plan.set_option(pkg._lib.OPT_GL_PERSISTENT, 0)
plan.set_option(pkg._lib.OPT_GL_TEAM, 1)
