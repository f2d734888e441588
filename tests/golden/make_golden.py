#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference implementation from /root/reference (build container only; the GPU
box has no /root/reference and never runs this script).

    python tests/golden/make_golden.py

The reference modules are loaded by file path with stub parent packages
(SURVEY.md appendix B): ``import fairseq`` itself fails here (omegaconf, hydra
missing) but audio_utils.py / vocoder.py / feature_transforms only need numpy
and torch.  ``librosa`` is not installed; the stand-in below implements
librosa.filters.mel's Slaney recipe (the only librosa call on the path).  The
stand-in is deliberately an independent implementation from oracle/mel.py: it
is cross-checked against torchaudio.functional.melscale_fbanks here.

Outputs (float32 unless noted), all produced by reference code:
  gl_small.npz     logmel [T,80], init phase [1025,T], waveform after n_iter
                   iterations, for a few (T, n_iter, kind) cases; spectral
                   convergence of the reference output
  basis.npz        mel filterbank [80,1025] (sparse form) and its torch.pinverse
                   [1025,80] (rows >= 683 are exactly zero; stored truncated)
  logmel.npz       waveform -> extract_logmel_spectrogram features
  fbank.npz        int16-scaled waveform -> _get_torchaudio_fbank features, 16 kHz and 8 kHz
  cmvn.npz         features, stats, GlobalCMVN output, gcmvn_denormalize output
  wss.npz          GriffinLim.get_window_sum_square for a few frame counts
  logmel_default.npz  log-mel with the reference's default geometry (n_fft 1024) and an n_fft 512 STFT with phase
  specaug_warp.npz  SpecAugmentTransform with time_warp_W > 0 (OpenCV) on seeded inputs
  gcmvn_stats.npz  get_global_cmvn (examples/speech_synthesis/data_utils.py:190-220) run on a directory of .npy
                   feature files: the files' seeds / shapes, the order Path.glob returned them in, mean and std

    python tests/golden/make_golden.py --only gcmvn     regenerates just that fixture
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _slaney_mel_standin(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel restated via torchaudio's slaney filterbank (independent of oracle/)."""
    import torchaudio.functional as taf
    fmax = sr / 2.0 if fmax is None else fmax
    fb = taf.melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax), n_mels, sr,
                             norm="slaney", mel_scale="slaney")
    return fb.T.contiguous().numpy().astype(np.float32)


def load_reference():
    for name in ["fairseq", "fairseq.data", "fairseq.data.audio", "fairseq.models",
                 "fairseq.models.text_to_speech"]:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    librosa = types.ModuleType("librosa")
    librosa.filters = types.ModuleType("librosa.filters")
    librosa.filters.mel = _slaney_mel_standin
    sys.modules["librosa"] = librosa
    sys.modules["librosa.filters"] = librosa.filters

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    au = load("fairseq.data.audio.audio_utils", "fairseq/data/audio/audio_utils.py")
    stub = types.ModuleType("fairseq.data.audio.speech_to_text_dataset")
    stub.S2TDataConfig = type("S2TDataConfig", (), {})
    sys.modules[stub.__name__] = stub
    load("fairseq.models.text_to_speech.hifigan", "fairseq/models/text_to_speech/hifigan.py")
    voc = load("fairseq.models.text_to_speech.vocoder", "fairseq/models/text_to_speech/vocoder.py")
    ft_dir = os.path.join(REF, "fairseq/data/audio/feature_transforms")
    spec = importlib.util.spec_from_file_location(
        "fairseq.data.audio.feature_transforms", os.path.join(ft_dir, "__init__.py"),
        submodule_search_locations=[ft_dir])
    ft = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = ft
    try:
        spec.loader.exec_module(ft)
    except ImportError as e:  # specaugment may need cv2; the cmvn modules are what we need
        print("note: feature_transforms auto-import stopped at:", e)
    return au, voc, ft


def synth_logmel(T, seed, kind):
    g = torch.Generator().manual_seed(seed)
    if kind == "smooth":
        x = 0.1 * torch.cumsum(torch.randn(T, 80, generator=g), dim=0)
        x = x + torch.linspace(0, -4, 80)[None, :] - 2.0
        x = x.clamp(float(np.log(1e-5)), 2.0)
    else:
        x = torch.randn(T, 80, generator=g) - 3.0
    return x.float()


def synth_audio(n, sr, seed):
    rng = np.random.RandomState(seed)
    t = np.arange(n) / sr
    x = 0.1 * rng.randn(n)
    for _ in range(3):
        x += rng.uniform(0.05, 0.3) * np.sin(2 * np.pi * rng.uniform(80, 0.45 * sr) * t + rng.uniform(0, 6.28))
    return np.clip(x, -1, 1).astype(np.float32)


def gcmvn_files(root):
    """The feature directory of the get_global_cmvn fixture: (name, T, seed) -> [T, 80] float32 .npy files (one of them
    saved as [1, T, 80]: the reference squeezes).  Shared with tests/test_frontend_gpu.py through the fixture's
    names / frames / seeds arrays."""
    specs = [("utt_%02d" % i, T, 500 + i) for i, T in enumerate((1998, 7, 333, 812, 64, 1203, 2, 450))]
    for k, (name, T, seed) in enumerate(specs):
        rng = np.random.RandomState(seed)
        x = (rng.randn(T, 80) * rng.uniform(0.2, 3.0, 80) + rng.uniform(-8, 2, 80)).astype(np.float32)
        np.save(os.path.join(root, name + ".npy"), x[None] if k == 3 else x)
    return specs


def make_gcmvn():
    """get_global_cmvn, the reference's own function (compiled by name from its file: the module imports the whole
    fairseq example stack), on a temporary directory.  Path.glob order is directory order, i.e. arbitrary: the fixture
    records the order the reference saw, because its float32 running sums depend on it in the last bits."""
    import ast
    import tempfile
    from pathlib import Path
    from typing import Optional
    from tqdm import tqdm
    src = open(os.path.join(REF, "examples/speech_synthesis/data_utils.py")).read()
    mod = ast.Module([n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_global_cmvn"], [])
    ns = {"np": np, "Path": Path, "Optional": Optional, "tqdm": tqdm}
    exec(compile(mod, "data_utils.py", "exec"), ns)
    with tempfile.TemporaryDirectory() as d:
        specs = gcmvn_files(d)
        order = [p.stem for p in Path(d).glob("*.npy")]
        st = ns["get_global_cmvn"](Path(d))
        out = os.path.join(d, "stats.npz")
        ns["get_global_cmvn"](Path(d), Path(out))
        saved = np.load(out)
        assert np.array_equal(saved["mean"], st["mean"]) and np.array_equal(saved["std"], st["std"])
    assert st["mean"].dtype == np.float32 and st["std"].dtype == np.float32
    np.savez_compressed(os.path.join(HERE, "gcmvn_stats.npz"), names=np.array([n for n, _, _ in specs]),
                        frames=np.array([t for _, t, _ in specs]), seeds=np.array([s for _, _, s in specs]),
                        glob_order=np.array(order), mean=st["mean"], std=st["std"])
    print("gcmvn ok: order", order, "mean[:3]", st["mean"][:3], "std[:3]", st["std"][:3])


def make_logmel_default():
    """extract_logmel_spectrogram with the reference's OWN default geometry (win 1024, hop 256, n_fft 1024, f_min 0,
    f_max 8000; examples/speech_synthesis/data_utils.py:46-52) at 22.05 kHz, and an n_fft = 512 STFT with phase: the
    sizes the generic (non-2048) kernels serve.  Same modules as the logmel.npz section."""
    au, _, _ = load_reference()
    out = {}
    spec_t = au.TTSSpectrogram(n_fft=1024, win_length=1024, hop_length=256, window_fn=torch.hann_window)
    mel_t = au.TTSMelScale(n_mels=80, sample_rate=22050, f_min=0.0, f_max=8000, n_stft=513)
    for i, n in enumerate((11025, 5000, 700)):
        w = synth_audio(n, 22050, 140 + i)
        with torch.no_grad():
            f = torch.clamp(mel_t(spec_t(torch.from_numpy(w)[None])), min=1e-5).log().squeeze(0).t().numpy()
        out["wave%d" % i] = w
        out["feat%d" % i] = f
        print("logmel default geometry", n, f.shape)
    st = au.TTSSpectrogram(n_fft=512, win_length=400, hop_length=160, window_fn=torch.hamming_window, return_phase=True)
    w = synth_audio(4000, 16000, 150)
    with torch.no_grad():
        mg, ph = st(torch.from_numpy(w)[None])
    out.update(stft512_in=w, stft512_mag=mg[0].numpy(), stft512_phase=ph[0].numpy())
    np.savez_compressed(os.path.join(HERE, "logmel_default.npz"), **out)


def make_specaug_warp():
    """SpecAugmentTransform with time warping (specaugment.py:96-110; needs OpenCV, present in the build container):
    the reference class on seeded inputs, several (T, W) incl. a spectrogram too short to be warped."""
    import cv2
    _, _, ft = load_reference()
    reg = ft.AUDIO_FEATURE_TRANSFORM_REGISTRY
    out = {}
    rng = np.random.RandomState(21)
    cases = {"w5": ({"time_warp_W": 5, "freq_mask_N": 1, "freq_mask_F": 10, "time_mask_N": 1, "time_mask_T": 20, "time_mask_p": 0.3}, (40, 333)),
             "w40": ({"time_warp_W": 40, "freq_mask_N": 2, "freq_mask_F": 27, "time_mask_N": 2, "time_mask_T": 100, "time_mask_p": 1.0,
                      "mask_value": 0.0}, (81, 250, 1203, 60)),
             "wonly": ({"time_warp_W": 8}, (17, 100))}
    for cname, (cfg, lengths) in cases.items():
        t = reg["specaugment"].from_config_dict(cfg)
        for T in lengths:
            x = (rng.randn(T, 80) * 2 - 4).astype(np.float32)
            out[f"{cname}_{T}_x"] = x
            for tag, use_ipp in (("y", True), ("y_noipp", False)):  # the pip wheel's default (IPP) / OpenCV's own code
                cv2.ipp.setUseIPP(use_ipp)
                np.random.seed(2000 + T)
                out[f"{cname}_{T}_{tag}"] = t(x)
    cv2.ipp.setUseIPP(True)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "specaug_warp.npz"), **out)
    print("specaug warp ok", len(out))


def main():
    if "--only" in sys.argv:
        {"gcmvn": make_gcmvn, "logmel_default": make_logmel_default,
         "specaug_warp": make_specaug_warp}[sys.argv[sys.argv.index("--only") + 1]]()
        return
    torch.set_num_threads(os.cpu_count())
    au, voc_mod, ft = load_reference()
    voc = voc_mod.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window,
                                    spec_bwd_max_iter=64)
    # ---- basis -------------------------------------------------------------------
    mel = au.get_mel_filters(24000, 2048, 80, 20, 8000).numpy()
    pinv = voc.inv_mel_transform.basis.numpy()
    nz_rows = np.nonzero(np.abs(pinv).sum(axis=1))[0]
    last = int(nz_rows.max())
    assert np.all(pinv[last + 1:] == 0)
    r, c = np.nonzero(mel)
    np.savez_compressed(os.path.join(HERE, "basis.npz"), mel_rows=r.astype(np.int16), mel_cols=c.astype(np.int16),
                        mel_vals=mel[r, c], pinv_head=pinv[: last + 1], last_nonzero_row=last)
    print("basis: mel nnz", len(r), "pinv last nonzero row", last)

    # ---- Griffin-Lim ---------------------------------------------------------------
    out = {}
    cases = [("c0", 60, 64, "smooth", 11), ("c1", 37, 8, "iid", 12), ("c2", 5, 4, "smooth", 13),
             ("c3", 96, 64, "iid", 14)]
    for name, T, n_iter, kind, seed in cases:
        x = synth_logmel(T, 1234 + seed, kind)
        np.random.seed(seed)
        phase = np.angle(np.exp(2j * np.pi * np.random.rand(1025, T))).astype(np.float32)
        voc.gl_transform.n_iter = n_iter
        np.random.seed(seed)  # forward() draws the same phase from the global RNG
        with torch.no_grad():
            y = voc(x).numpy()
            mag = voc.inv_mel_transform(x.exp().transpose(-1, -2))
            m2, _ = voc.gl_transform.transform(torch.from_numpy(y)[None])
            sc = float(torch.norm(m2[0] - mag) / torch.norm(mag))
        out[name + "_logmel"] = x.numpy()
        out[name + "_phase"] = phase.astype(np.float16)  # regenerated exactly from the seed in tests
        out[name + "_seed"] = seed
        out[name + "_n_iter"] = n_iter
        out[name + "_wave"] = y
        out[name + "_sc"] = sc
        print("gl", name, T, n_iter, kind, "L", y.shape, "sc", sc)
    # one STFT / ISTFT pair on a fixed signal, to pin the transforms separately
    w = synth_audio(6000, 24000, 5)
    with torch.no_grad():
        mg, ph = voc.gl_transform.transform(torch.from_numpy(w)[None])
        back = voc.gl_transform.inverse(mg, ph).squeeze().numpy()
    out["stft_in"] = w
    out["stft_mag"] = mg[0].numpy()
    out["stft_phase"] = ph[0].numpy()
    out["istft_out"] = back
    for k in list(out):
        if k.endswith("_phase") and k != "stft_phase":
            del out[k]  # the seed reproduces it bit-exactly; keep the fixture small
    np.savez_compressed(os.path.join(HERE, "gl_small.npz"), **out)

    # batched forward [B,T,80] draws one [B,1025,T] phase tensor
    xb = torch.stack([synth_logmel(24, 77, "smooth"), synth_logmel(24, 78, "iid")])
    voc.gl_transform.n_iter = 4
    np.random.seed(21)
    with torch.no_grad():
        yb = voc(xb).numpy()
    np.savez_compressed(os.path.join(HERE, "gl_batched.npz"), logmel=xb.numpy(), seed=21, n_iter=4, wave=yb)

    # ---- window sum square -------------------------------------------------------------
    wss = {"T%d" % t: voc_mod.GriffinLim.get_window_sum_square(t, 300, 1200, 2048).numpy() for t in (1, 5, 9, 40)}
    np.savez_compressed(os.path.join(HERE, "wss.npz"), **wss)

    # ---- log-mel front-end (speech_synthesis/data_utils.py:46-76 restated with the reference modules) ----
    spec_t = au.TTSSpectrogram(n_fft=2048, win_length=1200, hop_length=300, window_fn=torch.hann_window)
    mel_t = au.TTSMelScale(n_mels=80, sample_rate=24000, f_min=20, f_max=8000, n_stft=1025)
    lm = {}
    for i, n in enumerate((12000, 7777, 2400)):
        w = synth_audio(n, 24000, 40 + i)
        with torch.no_grad():
            f = torch.clamp(mel_t(spec_t(torch.from_numpy(w)[None])), min=1e-5).log().squeeze().t().numpy()
        lm["wave%d" % i] = w
        lm["feat%d" % i] = f
        print("logmel", n, f.shape)
    np.savez_compressed(os.path.join(HERE, "logmel.npz"), **lm)

    # ---- fbank ------------------------------------------------------------------------------
    fb = {}
    for i, (sr, n) in enumerate(((16000, 16000), (16000, 5003), (8000, 6000), (16000, 400))):
        w = (synth_audio(n, sr, 60 + i) * (2 ** 15)).astype(np.float32)
        f = au._get_torchaudio_fbank(w[None, :], sr, 80)
        fb["wave%d" % i] = w
        fb["sr%d" % i] = sr
        fb["feat%d" % i] = f
        print("fbank", sr, n, f.shape)
    np.savez_compressed(os.path.join(HERE, "fbank.npz"), **fb)

    # ---- CMVN ---------------------------------------------------------------------------------
    rng = np.random.RandomState(7)
    mean = rng.randn(80).astype(np.float32) - 4.0
    std = rng.uniform(0.5, 2.0, 80).astype(np.float32)
    stats_path = "/tmp/_golden_stats.npz"
    np.savez(stats_path, mean=mean, std=std)
    x = rng.randn(50, 80).astype(np.float32) * 2 - 4
    reg = ft.AUDIO_FEATURE_TRANSFORM_REGISTRY
    outs = {n: reg[n].from_config_dict({"stats_npz_path": stats_path})(x)
            for n in ("global_cmvn", "src_global_cmvn", "tgt_global_cmvn")}
    xt = torch.from_numpy(outs["global_cmvn"])[None]
    den = (xt * torch.from_numpy(std).view(1, 1, -1).expand_as(xt) + torch.from_numpy(mean).view(1, 1, -1).expand_as(xt))
    np.savez_compressed(os.path.join(HERE, "cmvn.npz"), x=x, mean=mean, std=std, denorm=den[0].numpy(), **outs)
    print("cmvn ok", {k: v.dtype for k, v in outs.items()})

    # ---- per-utterance transforms of the registry: utterance_cmvn, specaugment ----------------------
    tr = {}
    rng = np.random.RandomState(11)
    for name, T in (("a", 1), ("b", 7), ("c", 333), ("d", 1998)):
        xu = (rng.randn(T, 80) * rng.uniform(0.2, 3.0, 80) + rng.uniform(-8, 2, 80)).astype(np.float32)
        tr[f"ucmvn_{name}_x"] = xu
        for nm in (0, 1):
            for nv in (0, 1):
                t = reg["utterance_cmvn"].from_config_dict({"norm_means": bool(nm), "norm_vars": bool(nv)})
                tr[f"ucmvn_{name}_m{nm}v{nv}"] = t(xu)
    cfgs = {"ld": {"freq_mask_N": 2, "freq_mask_F": 27, "time_mask_N": 2, "time_mask_T": 100, "time_mask_p": 1.0},
            "zero": {"freq_mask_N": 1, "freq_mask_F": 10, "time_mask_N": 3, "time_mask_T": 40, "time_mask_p": 0.2,
                     "mask_value": 0.0},
            "fonly": {"freq_mask_N": 3, "freq_mask_F": 15}}
    for cname, cfg in cfgs.items():
        t = reg["specaugment"].from_config_dict(cfg)
        for name, T in (("s", 9), ("m", 250), ("l", 1203)):
            xs = (rng.randn(T, 80) * 2 - 4).astype(np.float32)
            np.random.seed(1000 + T)
            tr[f"spec_{cname}_{name}_x"] = xs
            tr[f"spec_{cname}_{name}_y"] = t(xs)
    np.savez_compressed(os.path.join(HERE, "transforms.npz"), **tr)
    print("transforms ok", len(tr))

    # ---- MCD metric: DTW + distance, the reference's own functions ----------------------------------
    # examples/s2s_trans/tasks/s2s_translation.py cannot be imported (fairseq task machinery), but the metric is a set
    # of pure functions: their unmodified source is compiled from the file.
    import ast
    import torch.nn.functional as F
    src = open(os.path.join(REF, "examples/s2s_trans/tasks/s2s_translation.py")).read()
    want = {"antidiag_indices", "batch_dynamic_time_warping", "compute_l2_dist", "compute_rms_dist", "get_divisor",
            "batch_compute_distortion", "batch_mel_cepstral_distortion"}
    mod = ast.Module([n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in want], [])
    ns = {"torch": torch, "np": np, "F": F}
    exec(compile(mod, "s2s_translation.py", "exec"), ns)
    rng = np.random.RandomState(3)
    dt = {}
    shapes = [(5, 7), (12, 9), (1, 6), (30, 31)]
    dm = np.zeros((len(shapes), 30, 31), np.float32)
    for b, (m, n) in enumerate(shapes):
        dm[b, :m, :n] = rng.rand(m, n).astype(np.float32) * 3
    dm[1, :12, :9] = rng.randint(0, 3, (12, 9))          # many exact ties
    c, bp, pm = ns["batch_dynamic_time_warping"](torch.from_numpy(dm), torch.LongTensor(shapes))
    dt.update(ragged_dist=dm, ragged_shapes=np.asarray(shapes, np.int64), ragged_cum=c.numpy(), ragged_bp=bp.numpy(),
              ragged_path=pm.numpy())
    dfull = (rng.rand(2, 40, 23).astype(np.float32) ** 2)
    c, bp, pm = ns["batch_dynamic_time_warping"](torch.from_numpy(dfull))
    dt.update(full_dist=dfull, full_cum=c.numpy(), full_bp=bp.numpy(), full_path=pm.numpy())
    x1, x2 = rng.randn(37, 13).astype(np.float32) * 4, rng.randn(29, 13).astype(np.float32) * 4
    dt.update(x1=x1, x2=x2, rms=ns["compute_rms_dist"](torch.from_numpy(x1), torch.from_numpy(x2)).numpy())
    # end to end: MCD of two waveform pairs at 24 kHz (torchaudio MFCC, CPU)
    ya = [torch.from_numpy(synth_audio(9000, 24000, 31)), torch.from_numpy(synth_audio(14000, 24000, 32))]
    yb = [torch.from_numpy(synth_audio(11000, 24000, 33)), torch.from_numpy(synth_audio(12500, 24000, 34))]
    for nt in ("path", "len1", None):
        rets = ns["batch_mel_cepstral_distortion"](ya, yb, 24000, normalize_type=nt)
        dt["mcd_" + str(nt)] = np.asarray([float(r[0]) for r in rets], np.float64)
    dt.update(mcd_ya0=ya[0].numpy(), mcd_ya1=ya[1].numpy(), mcd_yb0=yb[0].numpy(), mcd_yb1=yb[1].numpy())
    make_gcmvn()
    make_logmel_default()
    make_specaug_warp()
    np.savez_compressed(os.path.join(HERE, "dtw.npz"), **dt)
    print("dtw ok", {k: v.shape for k, v in dt.items() if k.startswith("mcd_") and v.ndim == 1 and v.size == 2})


if __name__ == "__main__":
    main()
