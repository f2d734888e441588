import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
for T in [int(a) for a in sys.argv[1:]] or [100, 500, 2300]:
    x = torch.from_numpy(bench.synth_logmel_np(T, 1)).cuda()
    ph = ((torch.rand(T, 1025, device="cuda") * 2 - 1) * np.pi).contiguous()
    print("T", T, flush=True)
    y = voc.synthesize_flat(x, [T], ph); torch.cuda.synchronize()
