"""GPU parity tests for the feature front-end: logmelspec80, Kaldi fbank80, global CMVN.
Feature tolerance (BASELINE.json: 1e-5 relative) is applied as relative L2 over each feature matrix;
see DESIGN.md for why an elementwise bound on log values is limited by the reference's own fp32 noise."""
import numpy as np
import pytest
import torch

from conftest import load_golden, synth_audio
from oracle import frontend as ofe
from oracle import griffin_lim as ogl

pytestmark = pytest.mark.gpu


def test_logmel_matches_reference_and_oracle(pkg, built_lib):
    l = load_golden("logmel.npz")
    waves = [l["wave%d" % i] for i in range(3)]
    feats = pkg.logmel_batch([torch.from_numpy(w) for w in waves], f_min=20.0)
    for w, f, i in zip(waves, feats, range(3)):
        f = f.cpu().numpy()
        ref = l["feat%d" % i]
        assert f.shape == ref.shape
        assert ogl.rel_l2(f, ref) < 1e-5 and np.abs(f - ref).max() < 1e-4
        assert ogl.rel_l2(f, ofe.logmel_spectrogram(w)) < 1e-5
    # same-signature single-utterance extractor
    f1 = pkg.extract_logmel_spectrogram(torch.from_numpy(waves[0])[None], 24000, win_length=1200, hop_length=300,
                                        n_fft=2048, f_min=20.0, f_max=8000.0)
    assert isinstance(f1, np.ndarray) and np.array_equal(f1, feats[0].cpu().numpy())
    f2 = pkg.extract_logmel_spectrogram(torch.from_numpy(waves[0])[None], 24000, win_length=1200, hop_length=300,
                                        n_fft=2048, f_min=20.0, f_max=8000.0, target_length=50)
    assert f2.shape == (50, 80) and np.all(f2[41:] == 0)


def test_logmel_fused_cmvn_and_silence(pkg, built_lib):
    rng = np.random.RandomState(1)
    mean, std = rng.randn(80).astype(np.float32), rng.uniform(0.5, 2, 80).astype(np.float32)
    w = synth_audio(9000, 24000, 3)
    plain = pkg.logmel_batch([torch.from_numpy(w)], f_min=20.0)[0]
    fused = pkg.logmel_batch([torch.from_numpy(w)], f_min=20.0, cmvn_mean=mean, cmvn_std=std)[0]
    ref = (plain.cpu().numpy() - mean) / std
    assert np.abs(fused.cpu().numpy() - ref).max() < 1e-5
    silent = pkg.logmel_batch([torch.zeros(3000)], f_min=20.0)[0]
    assert torch.allclose(silent, torch.full_like(silent, float(np.log(1e-5))))
    with pytest.raises(RuntimeError, match="Padding size"):
        pkg.logmel_batch([torch.zeros(1024)])


def test_tts_modules_compose_like_reference(pkg, built_lib):
    """TTSSpectrogram -> TTSMelScale -> clamp/log, the reference's extract_logmel_spectrogram body."""
    w = torch.from_numpy(synth_audio(7000, 24000, 8)).cuda()[None]
    spec = pkg.TTSSpectrogram(2048, 1200, 300)(w)
    mel = pkg.TTSMelScale(80, 24000, 20, 8000, 1025).cuda()(spec)
    got = torch.clamp(mel, min=1e-5).log().squeeze().t().cpu().numpy()
    assert ogl.rel_l2(got, ofe.logmel_spectrogram(w[0].cpu().numpy())) < 1e-5


def test_fbank_matches_reference_oracle_and_torchaudio(pkg, built_lib):
    fb = load_golden("fbank.npz")
    for i in range(4):
        w, sr, ref = fb["wave%d" % i], int(fb["sr%d" % i]), fb["feat%d" % i]
        f = pkg.fbank_batch([torch.from_numpy(w)], sr)[0].cpu().numpy()
        assert f.shape == ref.shape
        assert ogl.rel_l2(f, ref) < 1e-5 and np.abs(f - ref).max() < 2e-3
        assert ogl.rel_l2(f, ofe.kaldi_fbank(w, sr)) < 1e-5
    import torchaudio.compliance.kaldi as K  # third-party library the reference calls (audio_utils.py:141-147)
    w = (synth_audio(20000, 16000, 21) * 2 ** 15).astype(np.float32)
    f = pkg.fbank_batch([torch.from_numpy(w)], 16000)[0].cpu().numpy()
    ta = K.fbank(torch.from_numpy(w)[None], num_mel_bins=80, sample_frequency=16000).numpy()
    assert ogl.rel_l2(f, ta) < 1e-5


def test_fbank_ragged_batch_edges(pkg, built_lib):
    rng = np.random.RandomState(3)
    lens = [400, 399, 16000, 560, 5003, 401]
    waves = [(rng.randn(n) * 3000).astype(np.float32) for n in lens]
    outs = pkg.fbank_batch([torch.from_numpy(w) for w in waves], 16000)
    assert [o.shape[0] for o in outs] == [1, 0, 98, 2, 29, 1]
    for w, o in zip(waves, outs):
        ref = ofe.kaldi_fbank(w, 16000)
        assert o.shape == ref.shape
        if ref.shape[0]:
            assert ogl.rel_l2(o.cpu().numpy(), ref) < 1e-5
    # same-signature helpers
    f = pkg.extract_fbank_features(torch.from_numpy(waves[2] / 2 ** 15)[None], 16000)
    assert ogl.rel_l2(f, ofe.kaldi_fbank(waves[2], 16000)) < 1e-5
    audio_utils = __import__("importlib").import_module(pkg.__name__ + ".audio_utils")
    f2 = audio_utils._get_torchaudio_fbank(waves[2][None], 16000, 80)
    assert f2.shape == (98, 80) and f2.dtype == np.float32
    for sr in (8000, 24000):
        w = (synth_audio(sr, sr, 5) * 2 ** 15).astype(np.float32)
        assert ogl.rel_l2(pkg.fbank_batch([torch.from_numpy(w)], sr)[0].cpu().numpy(), ofe.kaldi_fbank(w, sr)) < 1e-5


def test_cmvn_bit_exact(pkg, built_lib, tmp_path):
    c = load_golden("cmvn.npz")
    p = tmp_path / "stats.npz"
    np.savez(p, mean=c["mean"], std=c["std"])
    for name in ("global_cmvn", "src_global_cmvn", "tgt_global_cmvn"):
        t = pkg.get_audio_feature_transform(name).from_config_dict({"stats_npz_path": str(p)})
        y = t(c["x"])
        assert isinstance(y, np.ndarray) and y.dtype == np.float32
        assert np.array_equal(y, c[name])  # IEEE subtract + divide: bit-exact with numpy
        yd = t.apply_cuda(torch.from_numpy(c["x"]).cuda())
        assert np.array_equal(yd.cpu().numpy(), c[name])
    den = pkg.gcmvn_denormalize(torch.from_numpy(c["global_cmvn"])[None].cuda(), c["mean"], c["std"])
    assert np.array_equal(den[0].cpu().numpy(), c["denorm"])
    # odd column count -> scalar kernel path
    x = np.random.RandomState(0).randn(7, 13).astype(np.float32)
    m, s = np.arange(13, dtype=np.float32), np.linspace(0.5, 2, 13).astype(np.float32)
    cm = __import__("importlib").import_module(pkg.__name__ + ".feature_transforms.global_cmvn")
    y = cm.cmvn_apply_cuda(torch.from_numpy(x).cuda(), torch.from_numpy(m).cuda(), torch.from_numpy(s).cuda())
    assert np.array_equal(y.cpu().numpy(), (x - m) / s)


def test_cmvn_stats(pkg, built_lib):
    rng = np.random.RandomState(2)
    feats = [rng.randn(n, 80).astype(np.float32) * 2 - 3 for n in (100, 1, 333)]
    st = pkg.global_cmvn_stats([torch.from_numpy(f).cuda() for f in feats])
    ref = ofe.global_cmvn_stats(feats)
    assert np.allclose(st["mean"], ref["mean"], atol=1e-5) and np.allclose(st["std"], ref["std"], atol=1e-5)


def test_mel_projection_tensor_core_vs_simt_vs_float64(pkg, built_lib):
    """TTSMelScale.forward: the tcgen05 3 x TF32 kernel against the FP32 SIMT CSR kernel and a float64 matmul, incl. a
    ragged last 128-frame tile, a single frame, and a bank that uses all 1025 bins (17 K chunks)."""
    import importlib
    plans = importlib.import_module(pkg.__name__ + ".plans")
    rng = np.random.RandomState(4)
    for n_mels, sr, f_max in ((80, 24000, 8000.0), (128, 24000, 12000.0), (40, 24000, 3000.0)):
        mel = pkg.TTSMelScale(n_mels, sr, 20.0, f_max, 1025).cuda()
        plan = plans.get_stft_plan("cuda", 2048, 2048, 512, n_mels, torch.ones(2048), mel=mel.basis)
        for T in (1, 127, 300):
            spec = torch.from_numpy(np.abs(rng.randn(1025, T)).astype(np.float32) * np.exp(rng.randn(1025, 1) * 2).astype(np.float32))
            want = mel.basis.double().cpu() @ spec.double()
            outs = {}
            for mode in (0, 1):
                plan.set_option(pkg._lib.OPT_MEL_PROJECT, mode)
                outs[mode] = mel(spec.cuda()).cpu().double()
            plan.set_option(pkg._lib.OPT_MEL_PROJECT, 0)
            for mode in (0, 1):
                assert outs[mode].shape == (n_mels, T)
                assert float((outs[mode] - want).norm() / want.norm()) < 2e-6, (n_mels, T, mode)
            assert float((outs[0] - outs[1]).abs().max() / want.abs().max()) < 2e-6


def test_default_arguments_and_other_fft_sizes(pkg, built_lib):
    """extract_logmel_spectrogram called with the reference's OWN defaults (win 1024 / hop 256 / n_fft 1024, f_min 0;
    examples/speech_synthesis/data_utils.py:46-52) and TTSSpectrogram / TTSMelScale at n_fft 512 / 4096: the generic
    kernels, against fixtures from the reference modules and the oracle."""
    g = load_golden("logmel_default.npz")
    for i in range(3):
        w = torch.from_numpy(g["wave%d" % i])[None]
        f = pkg.extract_logmel_spectrogram(w, 22050)  # every geometry argument at its default
        assert isinstance(f, np.ndarray) and f.shape == g["feat%d" % i].shape
        assert ogl.rel_l2(f, g["feat%d" % i]) < 1e-5
        assert np.abs(f - g["feat%d" % i]).max() < 1e-4
    st = pkg.TTSSpectrogram(512, 400, 160, window_fn=torch.hamming_window, return_phase=True)
    mag, ph = st(torch.from_numpy(g["stft512_in"]).cuda()[None])
    mag, ph = mag[0].cpu().numpy(), ph[0].cpu().numpy()
    assert mag.shape == g["stft512_mag"].shape == (257, 26)
    assert ogl.rel_l2(mag, g["stft512_mag"]) < 2e-6
    d = np.angle(np.exp(1j * (ph.astype(np.float64) - g["stft512_phase"])))
    assert np.sqrt((g["stft512_mag"] * d ** 2).sum() / g["stft512_mag"].sum()) < 1e-4
    # a large transform, ragged batch, fused CMVN and statistics through the generic path
    waves = [torch.from_numpy(synth_audio(n, 48000, 90 + i)) for i, n in enumerate((30000, 5000, 9001))]
    mean, std = np.linspace(-5, -3, 64).astype(np.float32), np.linspace(0.5, 2, 64).astype(np.float32)
    stats = torch.zeros(2, 64, dtype=torch.float64, device="cuda")
    outs = pkg.logmel_batch(waves, 48000, 2400, 600, 4096, torch.hann_window, 64, 50.0, 12000.0, stats=stats)
    fused = pkg.logmel_batch(waves, 48000, 2400, 600, 4096, torch.hann_window, 64, 50.0, 12000.0, cmvn_mean=mean, cmvn_std=std)
    for w, o, fo in zip(waves, outs, fused):
        ref = ofe.logmel_spectrogram(w.numpy(), 48000, 2400, 600, 4096, 64, 50.0, 12000.0)
        assert o.shape == ref.shape and ogl.rel_l2(o.cpu().numpy(), ref) < 1e-5
        assert np.abs(fo.cpu().numpy() - (o.cpu().numpy() - mean) / std).max() < 2e-5
    allf = torch.cat(outs).double()
    assert torch.allclose(stats[0], allf.sum(0), rtol=1e-6, atol=1e-4)
    # mel projection module at another size
    spec = torch.rand(2, 257, 9)
    mel = pkg.TTSMelScale(40, 16000, 20.0, 7600.0, 257)
    want = torch.matmul(mel.basis, spec)
    assert torch.allclose(mel(spec.cuda()).cpu(), want, rtol=1e-5, atol=1e-6)
    # what stays 2048-only says so, clearly, before any kernel runs
    with pytest.raises(ValueError, match="n_fft"):
        pkg.GriffinLim(1024, 1024, 256, 4)
    with pytest.raises(ValueError, match="n_fft"):
        pkg.GriffinLimVocoder(22050, 1024, 256, 1024, 80, 0, 8000, torch.hann_window)
    with pytest.raises(ValueError, match="power of two"):
        pkg.extract_logmel_spectrogram(torch.zeros(1, 4000), 22050, n_fft=1200, win_length=1200, hop_length=300)


def test_config3_shape_properties(pkg, built_lib):
    """BASELINE config 3 shape (fbank80 + global CMVN over many 8-20 s utterances at 16 kHz; 600 of them here, 8.4 M
    frames' worth of code paths) through size-independent properties: an utterance's features do not depend on the
    batch (bitwise), the fused CMVN equals features-then-CMVN to rounding, the fused statistics equal the sums of the
    returned features, runs are deterministic, and sampled utterances match the oracle."""
    import importlib
    rng = np.random.RandomState(0)
    n_utts, sr = 600, 16000
    lens = (rng.uniform(8, 20, n_utts) * sr).astype(np.int64)
    g = torch.Generator(device="cuda").manual_seed(3)
    flat = (torch.randn(int(lens.sum()), device="cuda", generator=g) * 0.1).clamp_(-1, 1) * 2 ** 15
    waves = list(torch.split(flat, lens.tolist()))
    mean = torch.from_numpy((rng.randn(80) - 4).astype(np.float32))
    std = torch.from_numpy(rng.uniform(0.5, 2, 80).astype(np.float32))
    stats = torch.zeros(2, 80, dtype=torch.float64, device="cuda")
    plain = pkg.fbank_batch(waves, sr, stats=stats)
    again = pkg.fbank_batch(waves, sr)
    fused = pkg.fbank_batch(waves, sr, cmvn_mean=mean, cmvn_std=std)
    assert [p.shape[0] for p in plain] == [1 + (int(n) - 400) // 160 for n in lens]
    assert all(torch.equal(a, b) for a, b in zip(plain, again))  # deterministic
    allf = torch.cat(plain)
    assert torch.isfinite(allf).all()
    assert torch.allclose(stats[0], allf.double().sum(0), rtol=1e-7, atol=1e-3)
    assert torch.allclose(stats[1], (allf.double() ** 2).sum(0), rtol=1e-7, atol=1e-2)
    ref_fused = (allf - mean.cuda()) / std.cuda()
    assert float((torch.cat(fused) - ref_fused).abs().max()) < 2e-5
    for i in (0, 17, 311, n_utts - 1):
        alone = pkg.fbank_batch([waves[i]], sr)[0]
        assert torch.equal(alone, plain[i])  # batch invariant, bitwise
        if i in (0, 311):
            ref = ofe.kaldi_fbank(waves[i].cpu().numpy(), sr)
            assert ogl.rel_l2(plain[i].cpu().numpy(), ref) < 1e-5
    cm = importlib.import_module(pkg.__name__ + ".feature_transforms.global_cmvn")
    sep = cm.cmvn_apply_cuda(allf, mean.cuda(), std.cuda())
    assert torch.equal(sep, ref_fused)  # the stand-alone kernel is the IEEE subtract / divide


def test_get_global_cmvn_drop_in_matches_reference_fixture(pkg, built_lib, tmp_path):
    """get_global_cmvn(feature_root, output_path=None) -- the reference's signature -- against the fixture generated by
    the reference's own function: bit-exact when the files are visited in the order the reference saw them (per-file
    float32 sums on the GPU, s2st_utterance_sums, added in that order); through the public entry (Path.glob order of
    THIS file system, which may differ) equal to float32 rounding of the running sums."""
    import importlib
    from test_oracle_golden import gcmvn_fixture_arrays
    feats = importlib.import_module(pkg.__name__ + ".features")
    g, arrays = gcmvn_fixture_arrays()
    for name, a in arrays.items():
        np.save(tmp_path / (name + ".npy"), a)
    st = feats._global_cmvn_from_paths([tmp_path / (str(n) + ".npy") for n in g["glob_order"]], batch_bytes=700000)
    assert st["mean"].dtype == np.float32 and st["std"].dtype == np.float32
    assert np.array_equal(st["mean"], g["mean"]) and np.array_equal(st["std"], g["std"])
    pub = pkg.get_global_cmvn(tmp_path)
    assert np.allclose(pub["mean"], g["mean"], rtol=2e-6, atol=1e-6) and np.allclose(pub["std"], g["std"], rtol=2e-6, atol=1e-6)
    out = tmp_path / "stats.npz"
    assert pkg.get_global_cmvn(tmp_path, out) is None
    saved = np.load(out)
    assert np.array_equal(saved["mean"], pub["mean"]) and np.array_equal(saved["std"], pub["std"])
    # the registry transform consumes exactly this file (global_cmvn.py:18-21)
    t = pkg.GlobalCMVN.from_config_dict({"stats_npz_path": str(out)})
    x = arrays["utt_01"]
    assert np.array_equal(t(x), np.divide(np.subtract(x, pub["mean"]), pub["std"]))


def test_fused_statistics_in_extraction_kernels(pkg, built_lib):
    """Sum / sum of squares accumulated inside the fbank / log-mel kernels (fast and generic variants) equal the sums
    over the features they return (float64), and give the statistics of a separate pass over the features."""
    import importlib
    plans = importlib.import_module(pkg.__name__ + ".plans")
    rng = np.random.RandomState(3)
    for sr, lens in ((16000, [401, 7777, 12345, 3001]), (8000, [281, 8000, 199, 2763]), (22050, [9000, 4000])):
        waves = [torch.from_numpy((rng.randn(n) * 2000).astype(np.float32)) for n in lens]
        for generic in (0, 1):
            plans.get_fbank_plan("cuda", sr, 80).set_option(pkg._lib.OPT_FRONTEND_GENERIC, generic)
            stats = torch.zeros(2, 80, dtype=torch.float64, device="cuda")
            mean, std = rng.randn(80).astype(np.float32), rng.uniform(0.5, 2, 80).astype(np.float32)
            pkg.fbank_batch(waves, sr, stats=stats, cmvn_mean=mean, cmvn_std=std)  # statistics are taken BEFORE the CMVN
            feats = torch.cat(pkg.fbank_batch(waves, sr)).double()
            assert torch.allclose(stats[0], feats.sum(0), rtol=1e-6, atol=1e-4), (sr, generic)
            assert torch.allclose(stats[1], (feats ** 2).sum(0), rtol=1e-6, atol=1e-3), (sr, generic)
        plans.get_fbank_plan("cuda", sr, 80).set_option(pkg._lib.OPT_FRONTEND_GENERIC, 0)
    waves = [torch.from_numpy(synth_audio(n, 24000, 70 + i)) for i, n in enumerate((24000, 7001, 1201, 40000))]
    stats = torch.zeros(2, 80, dtype=torch.float64, device="cuda")
    feats = pkg.logmel_batch(waves, f_min=20.0, stats=stats)
    allf = torch.cat(feats).double()
    assert torch.allclose(stats[0], allf.sum(0), rtol=1e-6, atol=1e-4) and torch.allclose(stats[1], (allf ** 2).sum(0), rtol=1e-6, atol=1e-3)
    st = pkg.global_cmvn_from_sums(stats, allf.shape[0])
    ref = ofe.global_cmvn_stats([f.cpu().numpy() for f in feats])
    assert np.allclose(st["mean"], ref["mean"], atol=1e-5) and np.allclose(st["std"], ref["std"], atol=1e-5)


def test_fast_kernels_ragged_chunks_match_oracle_and_generic(pkg, built_lib, monkeypatch):
    """The chunked kernels (k_fbank_fast: 16 frames per half-warp, two frames per transform at 8 kHz; k_logmel_fast:
    8 frames per warp) on batches whose utterances end inside chunks, have odd sample offsets (unaligned loads), odd
    frame counts and single frames -- against the oracle, and against the one-warp-per-frame kernels they replace."""
    import importlib
    plans = importlib.import_module(pkg.__name__ + ".plans")
    rng = np.random.RandomState(11)
    for sr, lens in ((8000, [200, 281, 8000, 199, 1000, 2763, 4001, 360]), (16000, [401, 7777, 400, 12345, 3001, 561])):
        waves = [(rng.randn(n) * 2000).astype(np.float32) for n in lens]
        mean, std = rng.randn(80).astype(np.float32), rng.uniform(0.5, 2, 80).astype(np.float32)
        outs = pkg.fbank_batch([torch.from_numpy(w) for w in waves], sr)
        fused = pkg.fbank_batch([torch.from_numpy(w) for w in waves], sr, cmvn_mean=mean, cmvn_std=std)
        fplan = plans.get_fbank_plan("cuda", sr, 80)
        fplan.set_option(pkg._lib.OPT_FRONTEND_GENERIC, 1)
        generic = pkg.fbank_batch([torch.from_numpy(w) for w in waves], sr)
        fplan.set_option(pkg._lib.OPT_FRONTEND_GENERIC, 0)
        for w, o, fo, g in zip(waves, outs, fused, generic):
            ref = ofe.kaldi_fbank(w, sr)
            assert o.shape == ref.shape == g.shape
            if ref.shape[0]:
                assert ogl.rel_l2(o.cpu().numpy(), ref) < 1e-5
                assert ogl.rel_l2(o.cpu().numpy(), g.cpu().numpy()) < 1e-5
                assert np.abs(fo.cpu().numpy() - (o.cpu().numpy() - mean) / std).max() < 2e-5
    lens = [1025, 3000, 1500, 24000, 2047, 7001, 1201]
    waves = [synth_audio(n, 24000, 30 + i) for i, n in enumerate(lens)]
    outs = pkg.logmel_batch([torch.from_numpy(w) for w in waves], f_min=20.0)
    lplan = plans.get_stft_plan("cuda", 2048, 1200, 300, 80, torch.hann_window(1200),
                                mel=pkg.get_mel_filters(24000, 2048, 80, 20.0, 8000.0))
    lplan.set_option(pkg._lib.OPT_FRONTEND_GENERIC, 1)
    generic = pkg.logmel_batch([torch.from_numpy(w) for w in waves], f_min=20.0)
    lplan.set_option(pkg._lib.OPT_FRONTEND_GENERIC, 0)
    for w, o, g in zip(waves, outs, generic):
        ref = ofe.logmel_spectrogram(w)
        assert o.shape == ref.shape
        assert ogl.rel_l2(o.cpu().numpy(), ref) < 1e-5
        assert ogl.rel_l2(o.cpu().numpy(), g.cpu().numpy()) < 1e-5


def test_utterance_cmvn_bit_exact_and_ragged_batch(pkg, built_lib):
    """utterance_cmvn: drop-in __call__ bit-identical to the reference's golden outputs (all four flag
    combinations, T = 1 ... 1998), and the ragged device batch equals the per-utterance calls."""
    ft = pkg.feature_transforms
    t = load_golden("transforms.npz")
    xs = [t[f"ucmvn_{n}_x"] for n in "abcd"]
    for nm in (0, 1):
        for nv in (0, 1):
            tr = ft.get_audio_feature_transform("utterance_cmvn").from_config_dict(
                {"norm_means": bool(nm), "norm_vars": bool(nv)})
            for n, x in zip("abcd", xs):
                y = tr(x)
                assert y.dtype == np.float32 and np.array_equal(y, t[f"ucmvn_{n}_m{nm}v{nv}"]), (n, nm, nv)
                assert np.array_equal(y, ofe.utterance_cmvn(x, bool(nm), bool(nv)))
            flat = torch.from_numpy(np.concatenate(xs)).cuda()
            yb = tr.apply_cuda(flat, [x.shape[0] for x in xs]).cpu().numpy()
            assert np.array_equal(yb, np.concatenate([t[f"ucmvn_{n}_m{nm}v{nv}"] for n in "abcd"]))
    empty = ft.get_audio_feature_transform("utterance_cmvn")()(np.zeros((0, 80), np.float32))
    assert empty.shape == (0, 80)


def test_specaugment_matches_reference(pkg, built_lib):
    """specaugment: masks are drawn on the host with the reference's RNG call sequence, the fill (and the local-mean
    mask value) run on the GPU.  Explicit mask values are bit-exact; the local mean is a float64-accumulated mean
    rounded to float32, which can differ from numpy's float32 pairwise mean in the last place (tolerance 1e-6)."""
    from test_oracle_golden import SPEC_CFGS
    ft = pkg.feature_transforms
    t = load_golden("transforms.npz")
    keys = {"freq_mask_n": "freq_mask_N", "freq_mask_f": "freq_mask_F", "time_mask_n": "time_mask_N",
            "time_mask_t": "time_mask_T", "time_mask_p": "time_mask_p", "mask_value": "mask_value"}
    for cname, cfg in SPEC_CFGS.items():
        tr = ft.get_audio_feature_transform("specaugment").from_config_dict({keys[k]: v for k, v in cfg.items()})
        xs, ys = [], []
        for name, T in (("s", 9), ("m", 250), ("l", 1203)):
            x, ref = t[f"spec_{cname}_{name}_x"], t[f"spec_{cname}_{name}_y"]
            np.random.seed(1000 + T)
            y = tr(x)
            assert y.shape == ref.shape and y.dtype == np.float32
            masked = ref != x
            assert np.array_equal(y[~masked], x[~masked])           # untouched cells are copied bit for bit
            if cfg["mask_value"] is None:
                assert np.allclose(y[masked], ref[masked], rtol=1e-6, atol=0)
            else:
                assert np.array_equal(y, ref)
            xs.append(x)
            ys.append(y)
        # ragged device batch: utterances draw their masks in order, one fill launch
        np.random.seed(77)
        singles = [tr(x) for x in xs]
        np.random.seed(77)
        yb = tr.apply_cuda(torch.from_numpy(np.concatenate(xs)).cuda(), [x.shape[0] for x in xs]).cpu().numpy()
        assert np.array_equal(yb, np.concatenate(singles))
    # fewer feature columns than freq_mask_F: the reference returns its input untouched
    tr = ft.get_audio_feature_transform("specaugment").from_config_dict({"freq_mask_N": 1, "freq_mask_F": 100})
    x = np.ones((5, 80), np.float32)
    assert tr(x) is x


def test_specaugment_time_warp_bit_exact(pkg, built_lib):
    """SpecAugment with time_warp_W > 0 (specaugment.py:96-110, cv2.resize INTER_LINEAR) against fixtures produced by
    the reference class with OpenCV 4.13: bit-exact for both arithmetics in the field (IPP on = the pip wheel's
    default, IPP off = OpenCV's own code), including a spectrogram too short to warp, then a ragged batch."""
    from test_oracle_golden import WARP_CASES as _WARP_CASES
    ft = pkg.feature_transforms
    g = load_golden("specaug_warp.npz")
    for cname, (cfg, lengths) in _WARP_CASES.items():
        for arith, tag in (("ipp", "y"), ("opencv", "y_noipp")):
            tr = ft.get_audio_feature_transform("specaugment").from_config_dict(cfg)
            tr.resize_arithmetic = arith
            for T in lengths:
                np.random.seed(2000 + T)
                y = tr(g[f"{cname}_{T}_x"])
                ref = g[f"{cname}_{T}_{tag}"]
                if cfg.get("mask_value", None) is None and "freq_mask_N" in cfg:   # local-mean mask value: 1e-6 relative
                    assert np.allclose(y, ref, rtol=2e-6, atol=0), (cname, T, arith)
                else:
                    assert np.array_equal(y, ref), (cname, T, arith, np.abs(y - ref).max())
    cfg, lengths = _WARP_CASES["w40"]
    tr = ft.get_audio_feature_transform("specaugment").from_config_dict(cfg)
    xs = [g[f"w40_{T}_x"] for T in lengths]
    np.random.seed(5)
    singles = [tr(x) for x in xs]
    np.random.seed(5)
    yb = tr.apply_cuda(torch.from_numpy(np.concatenate(xs)).cuda(), list(lengths)).cpu().numpy()
    assert np.array_equal(yb, np.concatenate(singles))


def test_dtw_bit_exact_and_mcd_metric(pkg, built_lib):
    rng = np.random.RandomState(17)
    """The MCD validation metric (examples/s2s_trans/tasks/s2s_translation.py:414-552): DTW recurrence, back pointers
    and path bit-identical to the reference's functions (ragged shapes, exact ties, full-size default), distance matrix
    within 1e-5, and the end-to-end distortion of two waveform pairs (torchaudio MFCC, as in the reference)."""
    import importlib
    mcd = importlib.import_module(pkg.__name__ + ".mcd")
    from oracle import dtw as odtw
    d = load_golden("dtw.npz")
    dist = torch.from_numpy(d["ragged_dist"]).cuda()
    cum, bp, path = mcd.batch_dynamic_time_warping(dist, torch.from_numpy(d["ragged_shapes"]))
    assert cum.is_cuda and bp.dtype == torch.int32 and path.dtype == torch.int32
    assert torch.equal(cum.cpu(), torch.from_numpy(d["ragged_cum"]))
    assert torch.equal(bp.cpu(), torch.from_numpy(d["ragged_bp"]))
    assert torch.equal(path.cpu(), torch.from_numpy(d["ragged_path"]))
    cum, bp, path = mcd.batch_dynamic_time_warping(torch.from_numpy(d["full_dist"]))   # CPU tensor in -> CPU out
    assert not cum.is_cuda
    assert np.array_equal(cum.numpy(), d["full_cum"]) and np.array_equal(bp.numpy(), d["full_bp"])
    assert np.array_equal(path.numpy(), d["full_path"])
    # a larger case against the oracle (one CTA walks > 1 block-width of cells per diagonal)
    rng = np.random.RandomState(0)
    big = rng.rand(3, 300, 280).astype(np.float32)
    oc, ob, op = odtw.batch_dynamic_time_warping(big[:1, :120, :100].copy())
    c, b, p = mcd.batch_dynamic_time_warping(torch.from_numpy(big[:1, :120, :100].copy()).cuda())
    assert np.array_equal(c.cpu().numpy(), oc) and np.array_equal(b.cpu().numpy(), ob) and np.array_equal(p.cpu().numpy(), op)
    c, b, p = mcd.batch_dynamic_time_warping(torch.from_numpy(big).cuda())
    assert torch.isfinite(c).all() and int(p[0].sum()) >= 300 and (p.sum(dim=(1, 2)) <= 579).all()
    # distance
    r = mcd.compute_rms_dist(torch.from_numpy(d["x1"]).cuda(), torch.from_numpy(d["x2"]).cuda()).cpu().numpy()
    assert ogl.rel_l2(r, d["rms"]) < 1e-5 and ogl.rel_l2(r, odtw.compute_rms_dist(d["x1"], d["x2"])) < 1e-6
    # end to end
    ya = [torch.from_numpy(d["mcd_ya0"]).cuda(), torch.from_numpy(d["mcd_ya1"]).cuda()]
    yb = [torch.from_numpy(d["mcd_yb0"]).cuda(), torch.from_numpy(d["mcd_yb1"]).cuda()]
    for nt in ("path", "len1", None):
        rets = mcd.batch_mel_cepstral_distortion(ya, yb, 24000, normalize_type=nt)
        got = np.asarray([float(r[0]) for r in rets])
        assert np.allclose(got, d["mcd_" + str(nt)], rtol=2e-3), (nt, got, d["mcd_" + str(nt)])
    with pytest.raises(ValueError, match="not supported"):
        mcd.batch_mel_cepstral_distortion(ya, yb, 24000, normalize_type="bogus")
    # the batched distance builder (one launch, padded on the device) == the per-pair matrices, bitwise; the returned
    # structure is the reference's; a foreign dist_fn takes the per-pair route and gives the same numbers
    feats = lambda y: y.reshape(-1, 13)[: y.numel() // 13]
    za = [torch.from_numpy(rng.randn(13 * n).astype(np.float32)).cuda() for n in (37, 5, 64)]
    zb = [torch.from_numpy(rng.randn(13 * n).astype(np.float32)).cuda() for n in (29, 9, 64)]
    rets = mcd.batch_compute_distortion(za, zb, 0, feats, mcd.compute_rms_dist, "len2")
    rets2 = mcd.batch_compute_distortion(za, zb, 0, feats, lambda a, b: mcd.compute_rms_dist(a, b), "len2")
    for (dist, (x1, x2, dm, cum, bp, pm)), (dist2, other), a, b in zip(rets, rets2, za, zb):
        m, n = a.numel() // 13, b.numel() // 13
        assert dm.shape == (64, 64) and cum.shape == bp.shape == pm.shape == (m, n)
        assert torch.equal(dm[:m, :n], mcd.compute_rms_dist(feats(a), feats(b))) and float(dm[m:].abs().sum() + dm[:, n:].abs().sum()) == 0
        assert float(dist) == float(dist2) == float(cum[-1, -1] / n)  # float32 division on the device, like the reference's tensor / int
        assert int(pm.sum()) >= max(m, n) and int(pm[0, 0]) == 1 and int(pm[-1, -1]) == 1


def test_composite_chain_on_a_ragged_device_batch(pkg, built_lib, tmp_path):
    """CompositeAudioFeatureTransform.apply_cuda == the numpy chain applied utterance by utterance."""
    ft = pkg.feature_transforms
    rng = np.random.RandomState(5)
    np.savez(tmp_path / "stats.npz", mean=rng.randn(80).astype(np.float32), std=rng.uniform(0.5, 2, 80).astype(np.float32))
    cfg = {"transforms": ["utterance_cmvn", "global_cmvn", "specaugment"],
           "global_cmvn": {"stats_npz_path": str(tmp_path / "stats.npz")},
           "specaugment": {"freq_mask_N": 2, "freq_mask_F": 20, "time_mask_N": 1, "time_mask_T": 30, "time_mask_p": 0.5,
                           "mask_value": 0.0}}
    chain = ft.CompositeAudioFeatureTransform.from_config_dict(cfg)
    xs = [(rng.randn(T, 80) * 2 - 3).astype(np.float32) for T in (40, 7, 513)]
    np.random.seed(9)
    singles = [chain(x) for x in xs]
    np.random.seed(9)
    yb = chain.apply_cuda(torch.from_numpy(np.concatenate(xs)).cuda(), [x.shape[0] for x in xs]).cpu().numpy()
    assert np.array_equal(yb, np.concatenate(singles))
    # the post-collate entry for DataLoader pipelines (numpy list in, device tensors out)
    np.random.seed(9)
    outs = chain.apply_cuda_from_host(xs)
    assert all(o.is_cuda and np.array_equal(o.cpu().numpy(), s) for o, s in zip(outs, singles))


def _forked_child(q):
    import importlib
    try:
        pkg = importlib.import_module("speech-to-speech-translation_b200")
        pkg.GlobalCMVN  # noqa: B018
        from numpy import zeros
        importlib.import_module("speech-to-speech-translation_b200.plans").require_cuda()
        q.put("no error")
    except RuntimeError as e:
        q.put(str(e))


def test_forked_worker_fails_with_guidance(pkg, built_lib):
    """The reference's threading model is forked DataLoader workers calling the numpy transforms.  A forked child cannot
    share the parent's CUDA context: the product says what to do instead of crashing inside CUDA."""
    import multiprocessing as mp
    torch.zeros(1).cuda()  # CUDA is initialised in this (parent) process
    ctx = mp.get_context("fork")
    q = ctx.SimpleQueue()
    p = ctx.Process(target=_forked_child, args=(q,))
    p.start()
    p.join(60)
    msg = q.get()
    assert "apply_cuda_from_host" in msg and "spawn" in msg, msg
