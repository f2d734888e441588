"""The "existing kernels" bar of SURVEY section 8d: the reference's dense-convolution Griffin-Lim (oracle/conv_formulation.py,
bit-identical to the reference on CPU) run on the GPU through cuDNN, one utterance per call like the reference's
generator loop, with TF32 off (parity grade) and on (what a user gets by default), next to this library on the same
inputs.   python tests/measure/conv_formulation_on_gpu.py"""
import importlib, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from oracle import conv_formulation as ocf
from oracle import griffin_lim as ogl
pkg = importlib.import_module(bench.PKG)
dev = torch.device("cuda", 0)
frames = bench.batch_frames(0)
sample = [frames[int(i)] for i in np.linspace(0, len(frames) - 1, 16)] + [500]
basis = ogl.pinv_mel_basis(bench.SR, bench.N_FFT, bench.N_MELS, bench.F_MIN, bench.F_MAX)
inputs = []
for i, T in enumerate(sample):
    np.random.seed(1000 + i)
    inputs.append((bench.synth_logmel_np(T, 1000 + i), ogl.random_phase((1025, T))))
audio = sum((T - 1) * 300 for T in sample) / 24000.0
out = {"sample": f"{len(sample)} utterances, {sum(sample)} frames, {audio:.1f} audio-s, 64 iterations, one utterance per call"}
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64).to(dev)
ours = []
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ours = voc.synthesize_batch([torch.from_numpy(x).to(dev) for x, _ in inputs], init_phase=[p for _, p in inputs], n_iter=64)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["this_library_one_batch"] = {"audio_s_per_s": audio / dt, "ms": 1e3 * dt}
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    gl = ocf.ConvGriffinLim(n_iter=64, device=dev)
    ys = []
    for rep in range(2):
        ys = []
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.no_grad():
            for x, ph in inputs:
                ys.append(ocf.vocoder_forward(x, ph, 64, basis, gl=gl))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    err = max(ogl.rel_l2(y, o.cpu().numpy()) for y, o in zip(ys, ours))
    out["cudnn_conv_tf32_" + ("on" if tf32 else "off")] = {"audio_s_per_s": audio / dt, "ms": 1e3 * dt,
                                                          "max_rel_l2_vs_this_library": float(err)}
print(json.dumps(out, indent=1))
