import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
frames = bench.batch_frames(0); total = sum(frames)
x = torch.from_numpy(np.concatenate([bench.synth_logmel_np(T, 1234 + i) for i, T in enumerate(frames)])).cuda()
rng = np.random.RandomState(0)
ph_np = torch.from_numpy(np.angle(np.exp(2j * np.pi * rng.rand(total, 1025))).astype(np.float32)).cuda()
ph_u = (torch.rand(total, 1025, device="cuda") * 2 - 1) * np.pi
for name, ph in (("np.angle phase", ph_np), ("uniform phase", ph_u), ("device-drawn", None)):
    for n in (5, 20):
        for _ in range(3): voc.synthesize_flat(x, frames, ph)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(n): voc.synthesize_flat(x, frames, ph)
        e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
        print(f"{os.environ.get('S2ST_GL_KERNEL','classic'):8s} {name:16s} n={n:2d}  gpu ms/step {e0.elapsed_time(e1)/n:7.3f}  host enqueue ms/step {1e3*(t1-t0)/n:7.3f}")
