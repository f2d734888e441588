"""Phase breakdown of one warp of the frame-parallel kernel: run with a -DS2ST_FRAMES_PROF build of the library
(python speech-to-speech-translation_b200/build.py -DS2ST_FRAMES_PROF -o$PWD/build_variants/frames_prof.so;
S2ST_B200_LIB=$PWD/build_variants/frames_prof.so python tools/frames_prof.py 500).  Output -> profiles/r02_small_calls.txt."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
for T in [int(a) for a in sys.argv[1:]] or [100, 500, 2300]:
    x = torch.from_numpy(bench.synth_logmel_np(T, 1)).cuda()
    ph = ((torch.rand(T, 1025, device="cuda") * 2 - 1) * np.pi).contiguous()
    print("T", T, flush=True)
    y = voc.synthesize_flat(x, [T], ph); torch.cuda.synchronize()
