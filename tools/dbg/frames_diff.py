import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import synth_logmel, seeded_phase
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
plan = voc._plan(torch.device("cuda", 0))
for T in (5, 12, 40):
    for n_iter in (1,):
        x = synth_logmel(T, 7).cuda(); ph = seeded_phase(3, T)
        outs = {}
        for name, strip, team in (("S=T team", T, 1), ("S=T noteam", T, 0), ("S=T+9 noteam", T + 9, 0), ("S=4 noteam", 4, 0)):
            plan.set_strip_frames(strip); plan.set_option(pkg._lib.OPT_GL_TEAM, team)
            outs[name] = voc.synthesize_batch([x], init_phase=[ph], n_iter=n_iter)[0]
        plan.set_strip_frames(0); plan.set_option(pkg._lib.OPT_GL_TEAM, 1)
        outs["frames"] = voc.synthesize_batch([x], init_phase=[ph], n_iter=n_iter)[0]
        ks = list(outs)
        for i in range(len(ks)):
            for j in range(i + 1, len(ks)):
                print(T, n_iter, ks[i], "vs", ks[j], "n_diff", int((outs[ks[i]] != outs[ks[j]]).sum()))
