"""Where does a synthesis step spend its time?  Times sub-sequences with CUDA events."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
frames = bench.batch_frames(0); total = sum(frames)
x = torch.from_numpy(np.concatenate([bench.synth_logmel_np(T, 1234 + i) for i, T in enumerate(frames)])).cuda()
ph = (torch.rand(total, 1025, device="cuda") * 2 - 1) * np.pi
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for n_iter in (0, 1, 2, 8, 64):
    print("n_iter", n_iter, "ms/step", round(timeit(lambda: voc.synthesize_flat(x, frames, ph, n_iter=n_iter)), 3))
plan = voc._plan(x.device)
plan.set_pass_timing(True)
voc.synthesize_flat(x, frames, ph, n_iter=8)
print("pass times ms:", np.round(plan.pass_times_ms(), 3))
import time
t0 = time.perf_counter(); voc.synthesize_flat(x, frames, ph, n_iter=64); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue ms", round(1e3 * (t1 - t0), 3), "until done ms", round(1e3 * (t2 - t0), 3))
