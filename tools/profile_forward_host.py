"""Host-side profile (cProfile) of GriffinLimVocoder.forward on one 500-frame utterance: where the ~0.5 ms of host time
per call goes (the GPU work is ~0.45 ms and overlaps with it)."""
import cProfile, importlib, os, pstats, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
x = torch.from_numpy(bench.synth_logmel_np(500, 1)).cuda()
np.random.seed(0)
for _ in range(20): voc(x)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(300): voc(x)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).strip_dirs().sort_stats("tottime").print_stats(32)
