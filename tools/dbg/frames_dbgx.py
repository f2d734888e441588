import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import synth_logmel, seeded_phase
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
plan = voc._plan(torch.device("cuda", 0))
win = torch.hann_window(1200).numpy().astype(np.float32)
for T in (5, 12, 40):
    x = synth_logmel(T, 7).cuda(); ph = seeded_phase(3, T)
    plan.set_strip_frames(T)
    w0 = voc.synthesize_batch([x], init_phase=[ph], n_iter=0)[0].cpu().numpy()
    plan.set_strip_frames(0)
    d = voc.synthesize_batch([x], init_phase=[ph], n_iter=1)[0].cpu().numpy()
    L = (T - 1) * 300
    bad = tot = 0
    for f in range(0, T, 4):
        for m in range(16, 1200):   # skip the first 16 (overwritten by the previous dumped frame's zero tail)
            j = f * 300 - 600 + m
            if 0 <= j < L - 1:
                tot += 1
                e = np.float32(w0[j] * win[m])
                if e != d[j]:
                    bad += 1
                    if bad < 5: print("  T", T, "f", f, "m", m, "j", j, e, d[j])
    print("T", T, "checked", tot, "mismatch", bad)
