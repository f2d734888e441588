"""A small pass over the front-end / adjacent kernels for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool racecheck python tools/sanitize_frontend.py"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import synth_audio, synth_logmel
pkg = importlib.import_module(bench.PKG)
mcd = importlib.import_module(bench.PKG + ".mcd")
feat = importlib.import_module(bench.PKG + ".features")
dev = torch.device("cuda", 0)
mean, std = torch.randn(80) - 4, torch.rand(80) + 0.5
for sr, lens in ((16000, (4001, 16000, 401)), (8000, (2001, 9000)), (22050, (5000,))):   # fast modes 0 / 1, generic kernel
    w = [torch.from_numpy(synth_audio(n, sr, 3 + i) * 20000) for i, n in enumerate(lens)]
    st = torch.zeros(2, 80, dtype=torch.float64, device=dev)
    pkg.fbank_batch(w, sr, cmvn_mean=mean, cmvn_std=std, stats=st)
for n_fft, win, hop in ((2048, 1200, 300), (1024, 1024, 256), (512, 400, 128)):             # register-resident / generic log-mel
    w = [torch.from_numpy(synth_audio(n, 24000, 7 + i)) for i, n in enumerate((6000, 12345))]
    st = torch.zeros(2, 80, dtype=torch.float64, device=dev)
    feat.logmel_batch(w, 24000, win, hop, n_fft, cmvn_mean=mean, cmvn_std=std, stats=st)
spec = pkg.TTSSpectrogram(2048, 1200, 300, return_phase=True).cuda()
mag, ph = spec(torch.from_numpy(synth_audio(9000, 24000, 1))[None].cuda())
mel = pkg.TTSMelScale(80, 24000, 20, 8000, 1025).cuda()(mag)                                  # tcgen05 mel projection
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=1).cuda()
voc.inv_mel_transform(synth_logmel(300, 5).cuda().exp().t())                                   # tcgen05 inverse mel
ft = pkg.feature_transforms
x = np.random.RandomState(0).randn(333, 80).astype(np.float32)
ft.get_audio_feature_transform("utterance_cmvn")()(x)
ft.get_audio_feature_transform("global_cmvn")  # registry lookup only (needs a stats file)
ft.get_audio_feature_transform("specaugment").from_config_dict({"freq_mask_N": 2, "freq_mask_F": 27, "time_mask_N": 2, "time_mask_T": 50, "time_mask_p": 1.0})(x)
d = torch.rand(3, 60, 50, device=dev)
mcd.batch_dynamic_time_warping(d, torch.tensor([[60, 50], [33, 50], [60, 7]]))
mcd.compute_rms_dist(torch.rand(13, 40, device=dev), torch.rand(17, 40, device=dev))
torch.cuda.synchronize()
print("sanitize_frontend ok")
