import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import synth_logmel, seeded_phase
from oracle import griffin_lim as ogl
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).cuda()
plan = voc._plan(torch.device("cuda", 0))

def one_iter64(w0, mag):   # float64 all the way (the "truth" of one iteration from w0)
    half = 1024
    xp = np.pad(w0.astype(np.float64), (half, half), mode="reflect")
    T = 1 + (xp.shape[0] - 2048) // 300
    w = ogl.padded_window(2048, 1200).astype(np.float64)
    idx = np.arange(2048)[None, :] + 300 * np.arange(T)[:, None]
    spec = np.fft.rfft(xp[idx] * w[None], axis=1)
    ph = spec / np.maximum(np.abs(spec), 1e-300)
    frames = np.fft.irfft(mag.T.astype(np.float64) * ph, n=2048, axis=1) * w[None]
    n = 2048 + 300 * (T - 1)
    y = np.zeros(n); wss = np.zeros(n)
    for t in range(T):
        y[t * 300: t * 300 + 2048] += frames[t]; wss[t * 300: t * 300 + 2048] += w ** 2
    nz = wss > 1e-30
    y[nz] /= wss[nz]
    return y[half:-half]

for T in (12, 40, 200):
    x = synth_logmel(T, 7).cuda(); ph = seeded_phase(3, T)
    mag = voc.inv_mel_transform(x.exp().t()).cpu().numpy()   # [1025, T]
    plan.set_strip_frames(T)
    w0 = voc.synthesize_batch([x], init_phase=[ph], n_iter=0)[0].cpu().numpy()
    s1 = voc.synthesize_batch([x], init_phase=[ph], n_iter=1)[0].cpu().numpy()
    plan.set_strip_frames(0)
    f1 = voc.synthesize_batch([x], init_phase=[ph], n_iter=1)[0].cpu().numpy()
    t1 = one_iter64(w0, mag)
    print("T", T, "strip err", np.abs(s1 - t1).max(), np.sqrt(np.mean((s1 - t1) ** 2)), "frames err", np.abs(f1 - t1).max(),
          np.sqrt(np.mean((f1 - t1) ** 2)), "strip vs frames", np.abs(s1 - f1).max(), "scale", np.abs(t1).max())
