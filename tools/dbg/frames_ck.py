import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import synth_logmel, seeded_phase
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
plan = voc._plan(torch.device("cuda", 0))
T = 12
x = synth_logmel(T, 7).cuda(); ph = seeded_phase(3, T)
plan.set_strip_frames(T); plan.set_option(pkg._lib.OPT_GL_TEAM, 0)
print("strips", flush=True)
voc.synthesize_batch([x], init_phase=[ph], n_iter=1); torch.cuda.synchronize()
plan.set_strip_frames(0)
print("frames", flush=True)
voc.synthesize_batch([x], init_phase=[ph], n_iter=1); torch.cuda.synchronize()
os.environ["S2ST_QUIET"] = "1"
for team in (0, 1):
    plan.set_strip_frames(T); plan.set_option(pkg._lib.OPT_GL_TEAM, team)
    a = voc.synthesize_batch([x], init_phase=[ph], n_iter=1)[0]
    plan.set_strip_frames(0)
    b = voc.synthesize_batch([x], init_phase=[ph], n_iter=1)[0]
    d = (a != b).nonzero().flatten().cpu().numpy()
    print("team", team, "n_diff", d.size, d[:10], float((a - b).abs().max()))
    i = int(d[0]) if d.size else 0
    print("  sample", i, a[i].item(), b[i].item(), "rel", abs(a[i].item() - b[i].item()) / abs(a[i].item()))
