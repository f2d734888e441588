"""A small pass over the round-2 kernels for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import synth_logmel, seeded_phase, synth_audio
pkg = importlib.import_module(bench.PKG)
voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=3).cuda()
# frame-parallel kernel: edge shapes, a ragged batch, T < 4 through the initial inverse, forward() with the device RNG
for frames in ([5], [12], [7, 40, 9]):
    feats = [synth_logmel(T, 1 + i).cuda() for i, T in enumerate(frames)]
    y = voc.synthesize_batch(feats, init_phase=[seeded_phase(9 + i, T) for i, T in enumerate(frames)], n_iter=3)
voc.synthesize_batch([synth_logmel(T, 5).cuda() for T in (2, 3, 4)], init_phase=[seeded_phase(3, T) for T in (2, 3, 4)], n_iter=0)
np.random.seed(1)
voc(synth_logmel(20, 2).cuda())
# strip kernels (pinned strip length), snake order with more strips than warps is exercised by the benchmarks
plan = voc._plan(torch.device("cuda", 0))
plan.set_strip_frames(4)
voc.synthesize_batch([synth_logmel(33, 3).cuda()], init_phase=[seeded_phase(4, 33)], n_iter=2)
plan.set_strip_frames(0)
# time warp, PCM16 in / out
ft = pkg.feature_transforms
tr = ft.get_audio_feature_transform("specaugment").from_config_dict({"time_warp_W": 5, "freq_mask_N": 1, "freq_mask_F": 10})
tr(np.random.RandomState(0).randn(60, 80).astype(np.float32))
pcm = torch.from_numpy((synth_audio(4001, 16000, 5) * 20000).astype(np.int16))
w = pkg.pcm16_to_waves(pcm.pin_memory(), normalization=False)
pkg.fbank_batch([w], 16000)
pkg.waves_to_pcm16(w / 32768.0)
torch.cuda.synchronize()
print("sanitize_small ok")
