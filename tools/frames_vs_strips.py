"""Frame-parallel kernel against the strip kernels with one strip per utterance: rel-L2 and the largest difference per hop
(the two are separately compiled and round differently in the last bit: see DESIGN.md 4.1c)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import synth_logmel, seeded_phase
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
plan = voc._plan(torch.device("cuda", 0))
for T in (5, 6, 12, 56):
    for n_iter in (1, 2):
        x = synth_logmel(T, 1900, "iid").cuda(); ph = seeded_phase(1950, T)
        plan.set_strip_frames(T)
        a = voc.synthesize_batch([x], init_phase=[ph], n_iter=n_iter)[0].cpu().numpy()
        plan.set_strip_frames(0)
        b = voc.synthesize_batch([x], init_phase=[ph], n_iter=n_iter)[0].cpu().numpy()
        d = np.abs(a - b)
        hops = d.reshape(-1, 300).max(axis=1)
        print(T, n_iter, "rel", np.linalg.norm(a - b) / np.linalg.norm(a), "max per hop:", " ".join(f"{h:.1e}" for h in hops[:8]), "...", " ".join(f"{h:.1e}" for h in hops[-6:]))
