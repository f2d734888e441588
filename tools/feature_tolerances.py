"""Achieved feature tolerances per golden fixture (rel-L2 over the matrix, max |diff|, max |diff| / max(|ref|, 1)):
CUDA front-end vs the fixtures produced by the unmodified reference (tests/golden)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import load_golden
pkg = importlib.import_module(bench.PKG)
def rep(name, got, ref):
    d = np.abs(got.astype(np.float64) - ref)
    print(f"{name:28s} shape {str(ref.shape):12s} rel-L2 {np.linalg.norm(d) / np.linalg.norm(ref):.2e}  max|d| {d.max():.2e}  "
          f"max|d|/max(|ref|,1) {(d / np.maximum(np.abs(ref), 1)).max():.2e}")
l = load_golden("logmel.npz")
for i in range(3):
    rep(f"logmel80 2048/300/1200 #{i}", pkg.logmel_batch([torch.from_numpy(l["wave%d" % i])], f_min=20.0)[0].cpu().numpy(), l["feat%d" % i])
g = load_golden("logmel_default.npz")
for i in range(3):
    rep(f"logmel80 1024/256/1024 #{i}", pkg.extract_logmel_spectrogram(torch.from_numpy(g["wave%d" % i])[None], 22050), g["feat%d" % i])
f = load_golden("fbank.npz")
for i in range(4):
    sr = int(f["sr%d" % i])
    rep(f"fbank80 {sr} Hz #{i}", pkg.fbank_batch([torch.from_numpy(f["wave%d" % i])], sr)[0].cpu().numpy(), f["feat%d" % i])
