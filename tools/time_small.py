"""Latency of small synthesis calls (config 1: one 500-frame utterance; one 60 s utterance), device-resident."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
for T in (100, 500, 4800):
    x = torch.from_numpy(bench.synth_logmel_np(T, 1)).cuda()
    for _ in range(3): voc.synthesize_flat(x, [T], None, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): voc.synthesize_flat(x, [T], None, seed=1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"T={T:5d}  {ms:.3f} ms per 64-iteration call  {(T - 1) * 300 / 24000 / (ms * 1e-3):.0f} audio-s/s  {ms / 65 * 1e3:.1f} us per pass")
