import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
pkg = importlib.import_module(bench.PKG)
plans = importlib.import_module(bench.PKG + ".plans")
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
frames = bench.batch_frames(0); total = sum(frames)
x = torch.from_numpy(np.concatenate([bench.synth_logmel_np(T, 1234 + i) for i, T in enumerate(frames)])).cuda()
dev = x.device
for _ in range(3): voc.synthesize_flat(x, frames, None)
torch.cuda.synchronize()
plan = voc._plan(dev); lib = pkg._lib.load()
fo = np.zeros(len(frames) + 1, np.int32); fo[1:] = np.cumsum(frames)
n_samples = (total - len(frames)) * 300
def T(): return time.perf_counter()
for rep in range(4):
    t0 = T(); fo_d = plans.upload_small(fo, dev)
    t1 = T(); wave = torch.empty(n_samples, dtype=torch.float32, device=dev)
    t2 = T(); ws = plan.workspace(len(frames), total)
    t3 = T()
    with torch.cuda.device(dev):
        rc = lib.s2st_gl_synthesize(plan.handle, len(frames), total, pkg._lib.ptr(fo_d), fo.ctypes.data, pkg._lib.ptr(x), None, None, 1234, 64,
                                    pkg._lib.ptr(wave), pkg._lib.ptr(ws), ws.numel(), pkg._lib.stream_ptr(dev))
    t4 = T()
    print(f"rep {rep}: upload {1e3*(t1-t0):.3f}  empty {1e3*(t2-t1):.3f}  workspace {1e3*(t3-t2):.3f}  gl_synthesize {1e3*(t4-t3):.3f} ms")
torch.cuda.synchronize()
