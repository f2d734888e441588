import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import synth_logmel, seeded_phase
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
plan = voc._plan(torch.device("cuda", 0))
for T in (12, 40):
    x = synth_logmel(T, 7).cuda(); ph = seeded_phase(3, T)
    plan.set_strip_frames(T)
    base = voc.synthesize_batch([x], init_phase=[ph], n_iter=1)[0]
    plan.set_strip_frames(0)
    outs = {}
    for w in (1, 2, 4, 8, 16):
        os.environ["S2ST_DBG_FRAMES_WARPS"] = str(w)
        outs[w] = voc.synthesize_batch([x], init_phase=[ph], n_iter=1)[0]
    print("T", T, {w: (int((outs[w] != outs[1]).sum()), int((outs[w] != base).sum())) for w in outs})
