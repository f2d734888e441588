"""The oracle (oracle/, numpy restatement) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, seeded_phase
from oracle import frontend as fe
from oracle import griffin_lim as gl
from oracle import mel as omel

CFG = dict(n_fft=2048, win_length=1200, hop_length=300)


def test_slaney_mel_matches_reference(golden_basis):
    mel, _ = golden_basis
    mine = omel.slaney_mel_filters(24000, 2048, 80, 20, 8000)
    assert (mine != 0).sum() == (mel != 0).sum() == 1334
    assert np.abs(mine - mel).max() <= 2.4e-7


def test_pinv_basis_matches_reference(golden_basis):
    _, pinv = golden_basis
    mine = gl.pinv_mel_basis(24000, 2048, 80, 20, 8000)
    assert np.all(mine[683:] == 0) and np.all(pinv[683:] == 0)  # f_max = 8 kHz -> bins >= 683 are exactly zero
    assert gl.rel_l2(mine, pinv) < 5e-5


def test_window_sum_square_matches_reference():
    w = load_golden("wss.npz")
    for k in w.files:
        mine = gl.window_sum_square(int(k[1:]), 300, 1200, 2048)
        assert mine.shape == w[k].shape
        assert np.abs(mine - w[k]).max() < 1e-6


def test_stft_istft_match_reference():
    g = load_golden("gl_small.npz")
    mag, ph = gl.stft(g["stft_in"], **CFG)
    assert mag.shape == g["stft_mag"].shape
    assert gl.rel_l2(mag, g["stft_mag"]) < 2e-6
    d = np.angle(np.exp(1j * (ph.astype(np.float64) - g["stft_phase"])))
    assert np.sqrt((g["stft_mag"] * d ** 2).sum() / g["stft_mag"].sum()) < 1e-5
    back = gl.istft(g["stft_mag"], g["stft_phase"], **CFG)
    assert back.shape == g["istft_out"].shape
    assert gl.rel_l2(back, g["istft_out"]) < 5e-6


@pytest.mark.parametrize("case", ["c0", "c1", "c2", "c3"])
def test_griffin_lim_matches_reference(case, golden_basis):
    """Waveform rel-L2 <= 1e-3 and spectral convergence within 1e-4 (BASELINE.json tolerances)."""
    _, pinv = golden_basis
    g = load_golden("gl_small.npz")
    x, n_iter, seed = g[case + "_logmel"], int(g[case + "_n_iter"]), int(g[case + "_seed"])
    T = x.shape[0]
    phase = seeded_phase(seed, T)
    for basis in (pinv, gl.pinv_mel_basis(24000, 2048, 80, 20, 8000)):
        y = gl.vocoder_forward(x, phase, n_iter, basis=basis)
        assert y.shape == g[case + "_wave"].shape == ((T - 1) * 300,)
        assert gl.rel_l2(y, g[case + "_wave"]) < 1e-3
        sc = gl.spectral_convergence(y, gl.inverse_mel(x, basis), **CFG)
        assert abs(sc - float(g[case + "_sc"])) < 1e-4


def test_batched_forward_phase_layout(golden_basis):
    """[B,T,80] input draws ONE [B,1025,T] phase tensor from the global RNG (vocoder.py:103)."""
    _, pinv = golden_basis
    g = load_golden("gl_batched.npz")
    x = g["logmel"]
    np.random.seed(int(g["seed"]))
    ph = np.angle(np.exp(2j * np.pi * np.random.rand(2, 1025, x.shape[1]))).astype(np.float32)
    for b in range(2):
        y = gl.vocoder_forward(x[b], ph[b], int(g["n_iter"]), basis=pinv)
        assert gl.rel_l2(y, g["wave"][b]) < 1e-3


def test_short_input_raises_like_reference():
    with pytest.raises(RuntimeError):
        gl.stft(np.zeros(900, np.float32), **CFG)  # T=4 -> L=900 <= 1024


def test_logmel_matches_reference():
    l = load_golden("logmel.npz")
    for i in range(3):
        f = fe.logmel_spectrogram(l["wave%d" % i])
        r = l["feat%d" % i]
        assert f.shape == r.shape == (1 + len(l["wave%d" % i]) // 300, 80)
        assert gl.rel_l2(f, r) < 1e-5
        assert np.abs(f - r).max() < 1e-4  # the reference's own fp32 conv noise is ~1.5e-5 abs here


def test_fbank_matches_reference():
    """torchaudio's own fp32-vs-fp64 self-noise on these inputs is up to 5.5e-4 abs, so the
    1e-5 criterion is applied as relative L2 over the feature matrix, plus a loose elementwise bound."""
    fb = load_golden("fbank.npz")
    for i in range(4):
        f = fe.kaldi_fbank(fb["wave%d" % i], int(fb["sr%d" % i]))
        r = fb["feat%d" % i]
        assert f.shape == r.shape
        assert gl.rel_l2(f, r) < 1e-5
        assert np.abs(f - r).max() < 2e-3
    assert fe.kaldi_fbank(np.zeros(399, np.float32), 16000).shape == (0, 80)  # shorter than one window


def test_cmvn_matches_reference_bit_exact():
    c = load_golden("cmvn.npz")
    y = fe.global_cmvn(c["x"], c["mean"], c["std"])
    for name in ("global_cmvn", "src_global_cmvn", "tgt_global_cmvn"):
        assert y.dtype == c[name].dtype and np.array_equal(y, c[name])
    assert np.array_equal(fe.gcmvn_denormalize(c["global_cmvn"], c["mean"], c["std"]), c["denorm"])


def test_cmvn_stats_roundtrip():
    rng = np.random.RandomState(0)
    feats = [rng.randn(n, 80).astype(np.float32) * 3 + 1 for n in (50, 70, 31)]
    st = fe.global_cmvn_stats(feats)
    allf = np.concatenate(feats)
    assert np.allclose(st["mean"], allf.mean(0), atol=1e-4)
    assert np.allclose(st["std"], allf.std(0), atol=1e-3)


def test_generic_geometry_oracle_matches_reference():
    """The reference's own default log-mel geometry (n_fft 1024 / hop 256 / win 1024, f_min 0, 22.05 kHz) and an
    n_fft 512 hamming STFT with phase: oracle vs the fixtures produced by the reference modules."""
    g = load_golden("logmel_default.npz")
    for i in range(3):
        f = fe.logmel_spectrogram(g["wave%d" % i], 22050, 1024, 256, 1024, 80, 0.0, 8000.0)
        assert f.shape == g["feat%d" % i].shape
        assert gl.rel_l2(f, g["feat%d" % i]) < 1e-5
    n = np.arange(400, dtype=np.float64)
    ham = (0.54 - 0.46 * np.cos(2 * np.pi * n / 400)).astype(np.float32)  # torch.hamming_window(400), periodic
    mag, ph = gl.stft(g["stft512_in"], 512, 400, 160, window=ham)
    assert gl.rel_l2(mag, g["stft512_mag"]) < 2e-6
    d = np.angle(np.exp(1j * (ph.astype(np.float64) - g["stft512_phase"])))
    assert np.sqrt((g["stft512_mag"] * d ** 2).sum() / g["stft512_mag"].sum()) < 1e-4


def gcmvn_fixture_arrays():
    """The feature files of tests/golden/gcmvn_stats.npz regenerated from their seeds (make_golden.gcmvn_files)."""
    g = load_golden("gcmvn_stats.npz")
    arrays = {}
    for k, (name, T, seed) in enumerate(zip(g["names"], g["frames"], g["seeds"])):
        rng = np.random.RandomState(int(seed))
        x = (rng.randn(int(T), 80) * rng.uniform(0.2, 3.0, 80) + rng.uniform(-8, 2, 80)).astype(np.float32)
        arrays[str(name)] = x[None] if k == 3 else x
    return g, arrays


def test_get_global_cmvn_matches_reference_bit_exact():
    """Oracle restatement of get_global_cmvn vs the fixture produced by the reference's own function, with the files in
    the order the reference's Path.glob returned them (its float32 running sums depend on that order)."""
    g, arrays = gcmvn_fixture_arrays()
    st = fe.get_global_cmvn([arrays[str(n)] for n in g["glob_order"]])
    assert st["mean"].dtype == np.float32
    assert np.array_equal(st["mean"], g["mean"]) and np.array_equal(st["std"], g["std"])


def test_utterance_cmvn_matches_reference_bit_exact():
    t = load_golden("transforms.npz")
    for name in "abcd":
        x = t[f"ucmvn_{name}_x"]
        for nm in (0, 1):
            for nv in (0, 1):
                y = fe.utterance_cmvn(x, bool(nm), bool(nv))
                assert y.dtype == np.float32 and np.array_equal(y, t[f"ucmvn_{name}_m{nm}v{nv}"])


SPEC_CFGS = {"ld": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0, mask_value=None),
             "zero": dict(freq_mask_n=1, freq_mask_f=10, time_mask_n=3, time_mask_t=40, time_mask_p=0.2, mask_value=0.0),
             "fonly": dict(freq_mask_n=3, freq_mask_f=15, mask_value=None)}


def test_specaugment_matches_reference_bit_exact():
    t = load_golden("transforms.npz")
    for cname, cfg in SPEC_CFGS.items():
        for name, T in (("s", 9), ("m", 250), ("l", 1203)):
            x = t[f"spec_{cname}_{name}_x"]
            np.random.seed(1000 + T)
            y = fe.specaugment(x, **cfg)
            assert np.array_equal(y, t[f"spec_{cname}_{name}_y"])
            assert (y != x).any() or cname == "zero" or T < 20


WARP_CASES = {"w5": ({"time_warp_W": 5, "freq_mask_N": 1, "freq_mask_F": 10, "time_mask_N": 1, "time_mask_T": 20, "time_mask_p": 0.3},
                      (40, 333)),
               "w40": ({"time_warp_W": 40, "freq_mask_N": 2, "freq_mask_F": 27, "time_mask_N": 2, "time_mask_T": 100,
                        "time_mask_p": 1.0, "mask_value": 0.0}, (81, 250, 1203, 60)),
               "wonly": ({"time_warp_W": 8}, (17, 100))}


def test_specaugment_time_warp_oracle_bit_exact():
    """The restated cv2.resize (oracle/frontend.py:resize_rows_linear) inside SpecAugment's time warp
    (specaugment.py:96-110) against the reference class's outputs, for both arithmetics (IPP on / off)."""
    g = load_golden("specaug_warp.npz")
    keys = {"time_warp_W": "time_warp_w", "freq_mask_N": "freq_mask_n", "freq_mask_F": "freq_mask_f", "time_mask_N": "time_mask_n",
            "time_mask_T": "time_mask_t", "time_mask_p": "time_mask_p", "mask_value": "mask_value"}
    for cname, (cfg, lengths) in WARP_CASES.items():
        kw = {keys[k]: v for k, v in cfg.items()}
        kw.setdefault("mask_value", None)
        for T in lengths:
            for tag, ipp in (("y", True), ("y_noipp", False)):
                np.random.seed(2000 + T)
                y = fe.specaugment(g[f"{cname}_{T}_x"], ipp=ipp, **kw)
                assert np.array_equal(y, g[f"{cname}_{T}_{tag}"]), (cname, T, tag)
    assert np.abs(g["w40_1203_y"] - g["w40_1203_y_noipp"]).max() > 1e-5   # the two arithmetics really differ


def test_mt19937_restatement_is_numpy():
    """oracle/mt19937.py (numpy's legacy global generator as one untempered sequence, the form the device kernel uses)
    against numpy itself: uniforms and the state left behind, from fresh seeds, mid-block, odd word positions, draws
    ending on block boundaries."""
    from oracle import mt19937 as mt
    for seed, pre, shape in [(0, 0, (3, 5)), (1, 7, (1025, 50)), (2, 623, (2, 3, 4)), (3, 1247, (1,)), (4, 100, (312,)),
                             (5, 0, (311,)), (6, 0, (312,)), (7, 5, (1025, 64)), (8, 0, (0,))]:
        np.random.seed(seed)
        if pre:
            np.random.rand(pre)
            if pre % 2:
                np.random.randint(0, 2 ** 31)
        st = np.random.get_state()
        u, st2 = mt.rand(st, shape)
        v = np.random.rand(*shape)
        st3 = np.random.get_state()
        assert np.array_equal(u, v) and np.array_equal(st2[1], st3[1]) and st2[2] == st3[2], (seed, pre, shape)
    # the block form (every word of a block from the previous block alone): numpy regenerates a whole block at once
    np.random.seed(11)
    key = np.random.get_state()[1].copy()
    for _ in range(3):
        np.random.bytes(4 * 624)  # consumes exactly one block of 32-bit outputs
        nxt = np.random.get_state()[1].copy()
        assert np.array_equal(mt.next_block(key), nxt)
        key = nxt


def test_conv_formulation_reproduces_reference_bit_for_bit(golden_basis):
    """oracle/conv_formulation.py is the reference's own dense-convolution formulation: same torch ops in the same
    order, so the golden waveforms come back exactly; and it pins the FFT oracle from a second, independent side."""
    from oracle import conv_formulation as cf
    _, pinv = golden_basis
    g = load_golden("gl_small.npz")
    eng = cf.ConvGriffinLim()
    for case in ("c1", "c2"):
        x, n_iter, seed = g[case + "_logmel"], int(g[case + "_n_iter"]), int(g[case + "_seed"])
        y = cf.vocoder_forward(x, seeded_phase(seed, x.shape[0]), n_iter, pinv, gl=eng)
        assert np.array_equal(y, g[case + "_wave"])
        assert gl.rel_l2(gl.vocoder_forward(x, seeded_phase(seed, x.shape[0]), n_iter, basis=pinv), y) < 1e-4
    w = load_golden("wss.npz")
    for k in w.files:
        assert np.array_equal(eng.window_sum_square(int(k[1:])).numpy(), w[k])


def test_dtw_and_rms_dist_match_reference():
    """oracle/dtw.py against the reference's own batch_dynamic_time_warping / compute_rms_dist (golden from the
    unmodified functions, incl. a case full of exact ties and ragged shapes)."""
    from oracle import dtw as odtw
    d = load_golden("dtw.npz")
    cum, bp, path = odtw.batch_dynamic_time_warping(d["ragged_dist"], d["ragged_shapes"])
    for b, (m, n) in enumerate(d["ragged_shapes"]):
        assert np.array_equal(cum[b, :m, :n], d["ragged_cum"][b, :m, :n])
        assert np.array_equal(bp[b, :m, :n], d["ragged_bp"][b, :m, :n])
    assert np.array_equal(path, d["ragged_path"])
    cum, bp, path = odtw.batch_dynamic_time_warping(d["full_dist"])
    assert np.array_equal(cum, d["full_cum"]) and np.array_equal(bp, d["full_bp"]) and np.array_equal(path, d["full_path"])
    assert gl.rel_l2(odtw.compute_rms_dist(d["x1"], d["x2"]), d["rms"]) < 1e-6
