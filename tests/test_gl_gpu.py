"""GPU parity tests for the Griffin-Lim path: CUDA kernels (through the C ABI) vs the oracle and the
reference's golden vectors.  Tolerances are BASELINE.json's: waveform rel-L2 <= 1e-3 after the
iterations, spectral convergence within 1e-4."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden, seeded_phase, synth_audio, synth_logmel
from oracle import griffin_lim as ogl

pytestmark = pytest.mark.gpu
CFG = dict(n_fft=2048, win_length=1200, hop_length=300)


@pytest.fixture(scope="module")
def voc(pkg, built_lib):
    v = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=64)
    return v.cuda()


@pytest.fixture(scope="module")
def basis(voc):
    return voc.inv_mel_transform.basis.cpu().numpy()


def test_native_library_is_loaded(pkg, built_lib):
    assert isinstance(built_lib, ctypes.CDLL)
    assert "libs2st_b200.so" in open("/proc/self/maps").read()


def test_rfft_irfft_2048(pkg, built_lib):
    from importlib import import_module
    plans = import_module(pkg.__name__ + ".plans")
    plan = plans.get_stft_plan("cuda", 2048, 1200, 300, 1, torch.hann_window(1200))
    rng = np.random.RandomState(0)
    n = 37
    x = rng.randn(n, 2048).astype(np.float32)
    x[3] = 0
    x[4, :] = 1.0
    xd = torch.from_numpy(x).cuda()
    out = torch.empty(n, 1025, 2, device="cuda")
    pkg._lib.check(built_lib.s2st_rfft2048(plan.handle, n, pkg._lib.ptr(xd), pkg._lib.ptr(out), pkg._lib.stream_ptr(xd.device)), "rfft")
    ref = np.fft.rfft(x.astype(np.float64), axis=1)
    got = out[..., 0].cpu().numpy() + 1j * out[..., 1].cpu().numpy()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-6
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-6
    # inverse of a Hermitian half-spectrum, with junk in the imaginary parts of DC / Nyquist (ignored)
    spec = np.stack([ref.real, ref.imag], -1).astype(np.float32)
    spec[:, 0, 1] = 5.0
    sd = torch.from_numpy(spec).cuda()
    back = torch.empty(n, 2048, device="cuda")
    pkg._lib.check(built_lib.s2st_irfft2048(plan.handle, n, pkg._lib.ptr(sd), pkg._lib.ptr(back), pkg._lib.stream_ptr(sd.device)), "irfft")
    assert np.abs(back.cpu().numpy() - x).max() < 5e-6


def test_stft_matches_oracle_and_reference(pkg, voc):
    g = load_golden("gl_small.npz")
    w = torch.from_numpy(g["stft_in"]).cuda()[None]
    mag, ph = voc.gl_transform.transform(w)
    assert mag.shape == (1, 1025, 21)
    mag, ph = mag[0].cpu().numpy(), ph[0].cpu().numpy()
    for ref_mag, ref_ph in ((g["stft_mag"], g["stft_phase"]), ogl.stft(g["stft_in"], **CFG)):
        assert ogl.rel_l2(mag, ref_mag) < 2e-6
        d = np.angle(np.exp(1j * (ph.astype(np.float64) - ref_ph)))
        assert np.sqrt((ref_mag * d ** 2).sum() / ref_mag.sum()) < 1e-5


def test_istft_matches_oracle_and_reference(pkg, voc):
    g = load_golden("gl_small.npz")
    mag = torch.from_numpy(g["stft_mag"]).cuda()[None]
    ph = torch.from_numpy(g["stft_phase"]).cuda()[None]
    y = voc.gl_transform.inverse(mag, ph)
    assert y.shape == (1, 1, 6000)
    y = y[0, 0].cpu().numpy()
    assert ogl.rel_l2(y, g["istft_out"]) < 5e-6
    assert ogl.rel_l2(y, ogl.istft(g["stft_mag"], g["stft_phase"], **CFG)) < 5e-6


@pytest.mark.parametrize("T", [1, 2, 5, 8, 9, 16, 17, 23])
def test_inverse_only_ragged_tile_edges(pkg, voc, T):
    """n_iter = 0 (initial inverse only) across tile boundaries (8 frames per tile), incl. T < 5."""
    rng = np.random.RandomState(T)
    mag = np.abs(rng.randn(1025, T)).astype(np.float32)
    ph = rng.uniform(-np.pi, np.pi, (1025, T)).astype(np.float32)
    y = voc.gl_transform.inverse(torch.from_numpy(mag).cuda()[None], torch.from_numpy(ph).cuda()[None])[0, 0].cpu().numpy()
    ref = ogl.istft(mag, ph, **CFG)
    assert y.shape == ref.shape == ((T - 1) * 300,)
    if T > 1:
        assert ogl.rel_l2(y, ref) < 5e-6


def test_inverse_mel_matches_oracle(pkg, voc, basis):
    x = synth_logmel(50, 5).cuda()
    spec = voc.inv_mel_transform(x.exp().t())  # reference surface: [n_mels, T] linear mel in
    ref = ogl.inverse_mel(x.cpu().numpy(), basis)
    assert spec.shape == (1025, 50)
    s = spec.cpu().numpy()
    assert np.all(s >= 0) and np.all(s[683:] == 0)
    assert ogl.rel_l2(s, ref) < 2e-6


def test_inverse_mel_tensor_core_vs_simt_vs_oracle(pkg, voc, basis, monkeypatch):
    """tcgen05 3xTF32 path against the FP32 SIMT kernel and the oracle, incl. a ragged last 128-frame tile."""
    import os
    plan = voc._plan(torch.device("cuda", 0))
    for n in (1, 127, 128, 300):
        x = torch.cat([synth_logmel(n, 50 + n), ]).cuda()
        outs = {}
        for mode in ("simt", "tc"):
            plan.set_option(pkg._lib.OPT_INVERSE_MEL, 1 if mode == "simt" else 0)
            out = torch.full((n, 1025), -1.0, device="cuda")
            rc = pkg._lib.load().s2st_inverse_mel(plan.handle, n, pkg._lib.ptr(x), 1, pkg._lib.ptr(out),
                                                  pkg._lib.stream_ptr(x.device))
            pkg._lib.check(rc, "s2st_inverse_mel")
            torch.cuda.synchronize()
            outs[mode] = out.cpu().numpy()
        plan.set_option(pkg._lib.OPT_INVERSE_MEL, 0)
        ref = ogl.inverse_mel(x.cpu().numpy(), basis).T
        for mode in ("simt", "tc"):
            assert np.all(outs[mode] >= 0) and np.all(outs[mode][:, 683:] == 0)
            assert ogl.rel_l2(outs[mode], ref) < 2e-6, mode
        assert ogl.rel_l2(outs["tc"], outs["simt"]) < 2e-6


@pytest.mark.parametrize("case", ["c0", "c1", "c2", "c3"])
def test_forward_matches_reference_golden(pkg, voc, basis, case):
    """Drop-in forward(): consumes numpy's global RNG like the reference, so seeding reproduces its output."""
    g = load_golden("gl_small.npz")
    x, n_iter, seed = g[case + "_logmel"], int(g[case + "_n_iter"]), int(g[case + "_seed"])
    voc.gl_transform.n_iter = n_iter
    np.random.seed(seed)
    y = voc(torch.from_numpy(x).cuda())
    voc.gl_transform.n_iter = 64
    assert y.is_cuda and y.dtype == torch.float32 and y.shape == ((x.shape[0] - 1) * 300,)
    y = y.cpu().numpy()
    assert ogl.rel_l2(y, g[case + "_wave"]) < 1e-3
    sc = ogl.spectral_convergence(y, ogl.inverse_mel(x, basis), **CFG)
    assert abs(sc - float(g[case + "_sc"])) < 1e-4
    # and against the oracle on the same inputs
    ref = ogl.vocoder_forward(x, seeded_phase(seed, x.shape[0]), n_iter, basis=basis)
    assert ogl.rel_l2(y, ref) < 1e-3


def test_batched_dense_forward_matches_reference(pkg, voc):
    g = load_golden("gl_batched.npz")
    voc.gl_transform.n_iter = int(g["n_iter"])
    np.random.seed(int(g["seed"]))
    y = voc(torch.from_numpy(g["logmel"]).cuda())
    voc.gl_transform.n_iter = 64
    assert y.shape == g["wave"].shape
    for b in range(2):
        assert ogl.rel_l2(y[b].cpu().numpy(), g["wave"][b]) < 1e-3


def test_batch_invariance_with_pinned_and_automatic_strips(pkg, voc, basis):
    """With a pinned strip length an utterance's waveform is bitwise independent of the rest of the batch;
    with the automatic choice (which depends on the batch) it moves by rounding noise only."""
    plan = voc._plan(torch.device("cuda", 0))
    frames = [300, 301, 64, 400] + [200] * 60
    feats = [synth_logmel(T, 500 + i) for i, T in enumerate(frames)]
    phases = [seeded_phase(600 + i, T) for i, T in enumerate(frames)]
    cu = [f.cuda() for f in feats]
    try:
        plan.set_strip_frames(12)
        full = voc.synthesize_batch(cu, init_phase=phases, n_iter=16)
        for i in (0, 1, 3):
            alone = voc.synthesize_batch([cu[i]], init_phase=[phases[i]], n_iter=16)[0]
            assert torch.equal(alone, full[i])
    finally:
        plan.set_strip_frames(0)
    auto_full = voc.synthesize_batch(cu, init_phase=phases, n_iter=16)
    auto_alone = voc.synthesize_batch([cu[3]], init_phase=[phases[3]], n_iter=16)[0]
    assert ogl.rel_l2(auto_alone.cpu().numpy(), auto_full[3].cpu().numpy()) < 1e-4
    assert ogl.rel_l2(auto_full[3].cpu().numpy(), full[3].cpu().numpy()) < 1e-4
    ref = ogl.vocoder_forward(feats[3].numpy(), phases[3], 16, basis=basis)
    assert ogl.rel_l2(auto_full[3].cpu().numpy(), ref) < 1e-3


def test_ragged_batch_equals_per_utterance_and_oracle(pkg, voc, basis):
    """The data-parallel entry: ragged lengths incl. strip-boundary cases; these small batches all get the same
    (minimum) strip length, so they are also bitwise equal to one-at-a-time runs."""
    frames = [5, 8, 9, 31, 56, 64, 65, 100]
    feats = [synth_logmel(T, 100 + i, "smooth" if i % 2 else "iid") for i, T in enumerate(frames)]
    phases = [seeded_phase(200 + i, T) for i, T in enumerate(frames)]
    outs = voc.synthesize_batch([f.cuda() for f in feats], init_phase=phases, n_iter=16)
    assert [o.numel() for o in outs] == [(T - 1) * 300 for T in frames]
    for i in (0, 2, 5, 7):
        single = voc.synthesize_batch([feats[i].cuda()], init_phase=[phases[i]], n_iter=16)[0]
        assert torch.equal(single, outs[i])  # batch composition must not change results
        ref = ogl.vocoder_forward(feats[i].numpy(), phases[i], 16, basis=basis)
        assert ogl.rel_l2(outs[i].cpu().numpy(), ref) < 1e-3
    # deterministic run to run
    again = voc.synthesize_batch([f.cuda() for f in feats], init_phase=phases, n_iter=16)
    assert all(torch.equal(a, b) for a, b in zip(outs, again))


def test_frame_parallel_path_equals_one_strip_per_utterance(pkg, voc, basis):
    """Calls that fit the resident warps run k_gl_frames (gl_frames.cuh): all iterations in one cooperative launch, a warp
    per frame, the overlap-add gathered from the neighbours' raw frames in frame order -- the same additions in the same
    order as the strip kernel's ring when ONE strip covers the utterance.  The two kernels are compiled separately, and
    ptxas (12.9) fuses mul.rn.f32x2 + add.rn.f32x2 pairs into FFMA2 as it sees fit (even with --fmad=false), so the
    transforms round differently in the last bit: equal to rounding noise (measured up to 5e-6 after 6 iterations, which amplify it),
    not bitwise.  Covered: every edge shape (T = 5 ... 12: both reflect paddings and clipped window sums interact), ragged
    batches, several frames per warp (> 2368 frames), the initial inverse alone, the device-drawn phase; the path is
    deterministic and independent of the batch (bitwise); and it is within tolerance of the oracle."""
    plan = voc._plan(torch.device("cuda", 0))
    cases = [[T] for T in (5, 6, 7, 8, 9, 10, 11, 12, 37, 500)] + [[5, 9, 31, 64, 65, 100, 257, 400], [56] * 30, [230] * 12, [1500, 1300]]
    worst = 0.0
    try:
        for ci, frames in enumerate(cases):
            feats = [synth_logmel(T, 1900 + i, "smooth" if i % 2 else "iid").cuda() for i, T in enumerate(frames)]
            phases = [seeded_phase(1950 + i, T) for i, T in enumerate(frames)]
            for n_iter in (0, 1, 6):
                plan.set_strip_frames(max(frames))
                base = voc.synthesize_batch(feats, init_phase=phases, n_iter=n_iter)
                assert plan.gl_launch_count(n_iter) == 2 + n_iter + 1
                plan.set_strip_frames(0)
                fr = voc.synthesize_batch(feats, init_phase=phases, n_iter=n_iter)
                assert plan.gl_launch_count(n_iter) == 3  # inverse_mel, build_frames, the frame kernel
                for a, b in zip(base, fr):
                    err = ogl.rel_l2(b.cpu().numpy(), a.cpu().numpy())
                    worst = max(worst, err)
                    assert err < (3e-6 if n_iter <= 1 else 2e-5), (frames, n_iter, err)
                again = voc.synthesize_batch(feats, init_phase=phases, n_iter=n_iter)
                assert all(torch.equal(a, b) for a, b in zip(fr, again))  # deterministic
            if ci == 10:  # batch independence: an utterance alone equals the same utterance inside the batch, bitwise
                alone = voc.synthesize_batch([feats[7]], init_phase=[phases[7]], n_iter=6)[0]
                assert torch.equal(alone, fr[7])
        # utterances of fewer than 4 frames (reachable with the initial inverse only: the reference's reflect padding
        # needs T >= 5): both edges interact, the cold per-sample path of the kernel
        for frames in ([2], [3], [4], [2, 3, 4, 9, 1, 5]):
            feats = [synth_logmel(T, 2100 + i).cuda() for i, T in enumerate(frames)]
            phases = [seeded_phase(2150 + i, T) for i, T in enumerate(frames)]
            plan.set_strip_frames(max(frames))
            base = voc.synthesize_batch(feats, init_phase=phases, n_iter=0)
            plan.set_strip_frames(0)
            fr = voc.synthesize_batch(feats, init_phase=phases, n_iter=0)
            assert plan.gl_launch_count(0) == 3
            for T, a, b, x, ph in zip(frames, base, fr, feats, phases):
                assert a.shape == b.shape == ((T - 1) * 300,)
                if T > 1:
                    assert ogl.rel_l2(b.cpu().numpy(), a.cpu().numpy()) < 1e-6, (frames, T)
                    assert ogl.rel_l2(b.cpu().numpy(), ogl.vocoder_forward(x.cpu().numpy(), ph, 0, basis=basis)) < 1e-5
        x = synth_logmel(300, 1).cuda()
        plan.set_strip_frames(300)
        a = voc.synthesize_flat(x, [300], None, n_iter=3, seed=11)
        plan.set_strip_frames(0)
        assert ogl.rel_l2(voc.synthesize_flat(x, [300], None, n_iter=3, seed=11).cpu().numpy(), a.cpu().numpy()) < 5e-6
        # the option switches the path off / bounds it
        plan.set_option(pkg._lib.OPT_GL_FRAMES, 0)
        voc.synthesize_flat(x, [300], None, n_iter=3, seed=11)
        assert plan.gl_launch_count(3) == 2 + 4
        plan.set_option(pkg._lib.OPT_GL_FRAMES, 299)
        voc.synthesize_flat(x, [300], None, n_iter=3, seed=11)
        assert plan.gl_launch_count(3) == 2 + 4
    finally:
        plan.set_strip_frames(0)
        plan.set_option(pkg._lib.OPT_GL_FRAMES, 16 * 148 * 4)
    print("frame-parallel vs one strip per utterance: worst rel-L2", worst)
    xs, ph = synth_logmel(37, 905, "smooth"), seeded_phase(955, 37)
    y = voc.synthesize_batch([xs.cuda()], init_phase=[ph], n_iter=16)[0].cpu().numpy()
    assert ogl.rel_l2(y, ogl.vocoder_forward(xs.numpy(), ph, 16, basis=basis)) < 1e-3


def test_config1_500_frames_64_iters(pkg, voc, basis):
    """BASELINE config 1: one 500-frame utterance, 64 iterations, seeded phase."""
    x = synth_logmel(500, 1234)
    phase = seeded_phase(0, 500)
    y = voc.synthesize_batch([x.cuda()], init_phase=[phase], n_iter=64)[0].cpu().numpy()
    ref = ogl.vocoder_forward(x.numpy(), phase, 64, basis=basis)
    assert y.shape == ref.shape == (149700,)
    assert ogl.rel_l2(y, ref) < 1e-3
    mag = ogl.inverse_mel(x.numpy(), basis)
    assert abs(ogl.spectral_convergence(y, mag, **CFG) - ogl.spectral_convergence(ref, mag, **CFG)) < 1e-4


def test_full_size_properties_long_form(pkg, voc, basis):
    """Config 5 shape (4800 frames): size-independent properties instead of a slow oracle run --
    finite, right length, spectral convergence improves with iterations, first frames match the oracle
    run on a prefix-independent quantity (the initial inverse, which is local)."""
    T = 4800
    x = synth_logmel(T, 99)
    phase = seeded_phase(7, T)
    xc = x.cuda()
    mag = ogl.inverse_mel(x.numpy(), basis)
    sc = []
    for n_iter in (0, 4, 32):
        y = voc.synthesize_batch([xc], init_phase=[phase], n_iter=n_iter)[0]
        assert y.shape == ((T - 1) * 300,) and torch.isfinite(y).all()
        m = voc.gl_transform.transform(y[None])[0][0].cpu().numpy()
        sc.append(np.linalg.norm(m - mag) / np.linalg.norm(mag))
        if n_iter == 0:
            ref0 = ogl.istft(mag[:, :40], phase[:, :40], **CFG)
            # samples covered only by frames < 37 are identical to the 40-frame problem's
            assert ogl.rel_l2(y[:9000].cpu().numpy(), ref0[:9000]) < 5e-6
    assert sc[0] > sc[1] > sc[2]


@pytest.mark.parametrize("n_iter", [64, 256])
def test_config5_long_form_vs_oracle(pkg, voc, basis, n_iter):
    """BASELINE config 5: one 60 s utterance (4800 frames, ~185 strips) at 64 and at 256 iterations against the
    oracle on the same seeded inputs -- the case that stresses overlap-add / window-sum normalisation and where seam
    rounding could accumulate.  Tolerances are BASELINE.json's (waveform rel-L2 <= 1e-3, |dSC| <= 1e-4)."""
    T = 4800
    x = synth_logmel(T, 99)
    phase = seeded_phase(7, T)
    y = voc.synthesize_batch([x.cuda()], init_phase=[phase], n_iter=n_iter)[0].cpu().numpy()
    ref = ogl.vocoder_forward(x.numpy(), phase, n_iter, basis=basis)
    assert y.shape == ref.shape == ((T - 1) * 300,)
    err = ogl.rel_l2(y, ref)
    mag = ogl.inverse_mel(x.numpy(), basis)
    d_sc = abs(ogl.spectral_convergence(y, mag, **CFG) - ogl.spectral_convergence(ref, mag, **CFG))
    print(f"config5 n_iter={n_iter}: rel-L2 {err:.3e}, |dSC| {d_sc:.3e}")
    assert err < 1e-3 and d_sc < 1e-4


def _config2_picks(frames):
    """(first frame, T) of the shortest, the median and the longest utterance of the length-sorted batch."""
    fo = np.concatenate([[0], np.cumsum(frames)])
    return [(i, int(fo[i]), frames[i]) for i in (0, len(frames) // 2, len(frames) - 1)]


def test_config2_full_batch_vs_oracle(pkg, voc, basis):
    """BASELINE config 2 at full shape: the real 256-utterance bench batch (bench.config2_batch), automatic strip
    length, 64 iterations; the shortest, the median and the LONGEST utterance against the oracle -- through the
    device-resident entry and through synthesize_host with the phase supplied from the host (the e2e path)."""
    import bench
    frames, logmel, phase = bench.config2_batch(0)
    total = int(sum(frames))
    assert len(frames) == 256 and frames[0] >= 56 and frames[-1] <= 400
    wave = voc.synthesize_flat(torch.from_numpy(logmel).cuda(), frames, torch.from_numpy(phase).cuda(), n_iter=64)
    wave_h = torch.zeros((total - len(frames)) * 300).pin_memory()
    voc.synthesize_host(torch.from_numpy(logmel).pin_memory(), frames, wave_h, phase_host=torch.from_numpy(phase).pin_memory(),
                        n_iter=64).synchronize()
    assert torch.equal(wave.cpu(), wave_h)  # same kernels, same strip length: bitwise
    wave = wave.cpu().numpy()
    for i, f0, T in _config2_picks(frames):
        w0 = (f0 - i) * 300
        y = wave[w0: w0 + (T - 1) * 300]
        x, ph = logmel[f0: f0 + T], np.ascontiguousarray(phase[f0: f0 + T].T)
        ref = ogl.vocoder_forward(x, ph, 64, basis=basis)
        mag = ogl.inverse_mel(x, basis)
        err = ogl.rel_l2(y, ref)
        d_sc = abs(ogl.spectral_convergence(y, mag, **CFG) - ogl.spectral_convergence(ref, mag, **CFG))
        print(f"config2 utterance {i} (T={T}): rel-L2 {err:.3e}, |dSC| {d_sc:.3e}")
        assert err < 1e-3 and d_sc < 1e-4, (i, T, err, d_sc)


def test_config4_bucket_vs_oracle(pkg, voc, basis):
    """BASELINE config 4: the 10k-utterance list is sharded and cut into length buckets (bench.run_gl_sharded); the
    buckets that hold the shortest and the longest utterance of rank 0's shard at world size 8, synthesised exactly as
    the bench does, against the oracle.  Also: an utterance's waveform does not depend on the sharding (world 8 vs 2)
    beyond rounding noise (the automatic strip length follows the bucket)."""
    import importlib

    import bench
    sh = importlib.import_module(pkg.__name__ + ".sharding")
    frames_all = bench.sharded_frames(0)
    owned8 = sh.shard_utterances(frames_all, 8, 64)
    b8 = sh.length_buckets(frames_all, owned8[0], bench.BUCKET_FRAMES, balanced=True)

    def run_bucket(b):
        fr = [frames_all[i] for i in b]
        xs, ps = zip(*(bench.sharded_utterance_inputs(i, frames_all[i]) for i in b))
        w = voc.synthesize_flat(torch.from_numpy(np.concatenate(xs)).cuda(), fr, torch.from_numpy(np.concatenate(ps)).cuda(), n_iter=64)
        return dict(zip(b, torch.split(w, [(T - 1) * 300 for T in fr])))

    out8 = run_bucket(b8[0])
    if len(b8) > 1:
        out8.update(run_bucket(b8[-1]))
    ids = sorted(out8, key=lambda i: (frames_all[i], i))
    for i in (ids[0], ids[-1]):
        x, p = bench.sharded_utterance_inputs(i, frames_all[i])
        ref = ogl.vocoder_forward(x, np.ascontiguousarray(p.T), 64, basis=basis)
        err = ogl.rel_l2(out8[i].cpu().numpy(), ref)
        print(f"config4 utterance {i} (T={frames_all[i]}): rel-L2 {err:.3e}")
        assert err < 1e-3
    # the same utterance inside a different bucket (a 2-rank sharding puts it among other neighbours)
    i = ids[-1]
    owned2 = sh.shard_utterances(frames_all, 2, 64)
    r = 0 if i in owned2[0] else 1
    b2 = [b for b in sh.length_buckets(frames_all, owned2[r], bench.BUCKET_FRAMES, balanced=True) if i in b][0]
    out2 = run_bucket(b2)
    assert ogl.rel_l2(out2[i].cpu().numpy(), out8[i].cpu().numpy()) < 1e-4


def test_device_drawn_initial_phase(pkg, voc, basis):
    """phase_fm=None: U[-pi, pi) drawn on the device.  Deterministic per seed, different across seeds, and
    statistically equivalent to the host draw: same spectral convergence after the iterations (it is a
    different random start, so waveforms are not compared sample by sample)."""
    frames = [60, 75, 133]
    feats = [synth_logmel(T, 300 + i) for i, T in enumerate(frames)]
    flat = torch.cat(feats).cuda()
    a = voc.synthesize_flat(flat, frames, None, n_iter=32, seed=7)
    b = voc.synthesize_flat(flat, frames, None, n_iter=32, seed=7)
    c = voc.synthesize_flat(flat, frames, None, n_iter=32, seed=8)
    assert torch.equal(a, b) and not torch.equal(a, c) and torch.isfinite(a).all()
    host = voc.synthesize_batch([f.cuda() for f in feats], init_phase=[seeded_phase(40 + i, T) for i, T in enumerate(frames)], n_iter=32)
    off = 0
    for f, T, h in zip(feats, frames, host):
        L = (T - 1) * 300
        mag = ogl.inverse_mel(f.numpy(), basis)
        sc_dev = ogl.spectral_convergence(a[off: off + L].cpu().numpy(), mag, **CFG)
        sc_host = ogl.spectral_convergence(h.cpu().numpy(), mag, **CFG)
        assert abs(sc_dev - sc_host) < 0.25 * sc_host, (sc_dev, sc_host)  # two random starts: same ballpark
        off += L
    # the n_iter = 0 output has the statistics of a uniform phase: compare its energy with the host draw's
    e_dev = float(voc.synthesize_flat(flat, frames, None, n_iter=0, seed=3).pow(2).mean())
    ph = torch.from_numpy(np.concatenate([seeded_phase(60 + i, T).T for i, T in enumerate(frames)])).cuda()
    e_host = float(voc.synthesize_flat(flat, frames, ph.contiguous(), n_iter=0).pow(2).mean())
    assert abs(e_dev - e_host) < 0.1 * e_host


def test_initial_phase_on_device_equals_host_draw(pkg, voc):
    """forward() continues numpy's GLOBAL generator on the device (s2st_phase_from_mt19937: the MT19937 state goes down,
    the advanced state comes back): bitwise the float32 phases of vocoder.py:103, frame-major, and numpy's generator is
    left exactly where the reference's np.random.rand(*shape) would leave it -- for starts at a fresh seed (position
    624), in the middle of a block, at an odd word position, and for draws ending exactly on a block boundary.  The
    host-RNG variant (s2st_phase_from_uniform), kept as the fallback, is checked the same way."""
    import importlib
    vm = importlib.import_module(pkg.__name__ + ".vocoder")
    dev0 = torch.device("cuda", 0)
    for shape, pre in (((1025, 37), 0), ((2, 1025, 50), 11), ((3, 33, 5), 0), ((1025, 500), 3), ((312,  1), 0), ((1, 311, 1), 0),
                       ((5, 7), 1247), ((1025, 8), 623)):
        def prepare():
            np.random.seed(5)
            if pre:
                np.random.rand(pre)
                if pre % 2:
                    np.random.randint(0, 2 ** 31)  # an odd number of 32-bit words consumed
        prepare()
        host = vm.draw_initial_phase(shape)
        state_host = np.random.get_state()
        after_host = np.random.rand(3)
        F, T = shape[-2], shape[-1]
        want = host.reshape(-1, F, T).transpose(0, 2, 1).reshape(-1, F)
        prepare()
        dev, finish = vm.draw_initial_phase_device(shape, dev0)
        finish()
        state_dev = np.random.get_state()
        assert state_dev[2] == state_host[2] and np.array_equal(state_dev[1], state_host[1]), shape
        assert np.array_equal(np.random.rand(3), after_host)  # the same amount of the global stream was consumed
        dev = dev.cpu().numpy()
        assert dev.shape == want.shape and np.array_equal(dev, want), shape
        prepare()
        dev2 = vm._draw_initial_phase_host_rng(shape, dev0).cpu().numpy()
        assert np.array_equal(np.random.rand(3), after_host) and np.array_equal(dev2, want)
    # a second draw before the first one's finish(): the pending state is handed back first (and finish() is idempotent)
    np.random.seed(21)
    h1, h2 = vm.draw_initial_phase((33, 7)), vm.draw_initial_phase((33, 9))
    tail = np.random.rand()
    np.random.seed(21)
    d1, f1 = vm.draw_initial_phase_device((33, 7), dev0)
    d2, f2 = vm.draw_initial_phase_device((33, 9), dev0)
    f1(); f2(); f1()
    assert np.random.rand() == tail
    assert np.array_equal(d1.cpu().numpy(), h1.T) and np.array_equal(d2.cpu().numpy(), h2.T)
    # forward() itself: seeding numpy reproduces the call, and two calls in a row continue the stream
    x = synth_logmel(30, 3).cuda()
    voc.gl_transform.n_iter = 2
    try:
        np.random.seed(9)
        y1, y2 = voc(x), voc(x)
        tail = np.random.rand()
        np.random.seed(9)
        ph1 = vm.draw_initial_phase((1025, 30))
        ph2 = vm.draw_initial_phase((1025, 30))
        assert np.random.rand() == tail
        r1 = voc.synthesize_batch([x], init_phase=[ph1], n_iter=2)[0]
        r2 = voc.synthesize_batch([x], init_phase=[ph2], n_iter=2)[0]
        assert torch.equal(y1, r1) and torch.equal(y2, r2) and not torch.equal(y1, y2)
    finally:
        voc.gl_transform.n_iter = 64


def test_half_precision_io(pkg, voc):
    x = synth_logmel(20, 3).cuda()
    voc.gl_transform.n_iter = 2
    np.random.seed(1)
    y32 = voc(x)
    np.random.seed(1)
    y16 = voc(x.half())
    voc.gl_transform.n_iter = 64
    assert y16.dtype == torch.float16 and y16.shape == y32.shape
    assert torch.isfinite(y16).all()


def test_cpu_tensor_in_gives_cpu_tensor_out(pkg, voc):
    x = synth_logmel(12, 4)
    voc.gl_transform.n_iter = 1
    np.random.seed(2)
    y = voc(x)
    voc.gl_transform.n_iter = 64
    assert y.device.type == "cpu" and y.shape == (3300,)


def test_short_utterance_raises(pkg, voc):
    with pytest.raises(RuntimeError, match="Padding size"):
        voc(synth_logmel(4, 1).cuda())
    with pytest.raises(AssertionError):
        voc(torch.zeros(10, 79).cuda())


def test_griffin_lim_module_forward_on_magnitudes(pkg, voc, basis):
    """GriffinLim.forward on a full 1025-bin magnitude (no zero tail): exercises kb = 1025 incl. Nyquist."""
    rng = np.random.RandomState(5)
    T = 12
    mag = np.abs(rng.randn(1025, T)).astype(np.float32) + 0.1
    gl = voc.gl_transform
    gl.n_iter = 3
    np.random.seed(9)
    y = gl(torch.from_numpy(mag).cuda()).cpu().numpy()
    gl.n_iter = 64
    ref = ogl.griffin_lim(mag, seeded_phase(9, T), 3, **CFG)
    assert ogl.rel_l2(y, ref) < 1e-4


def test_host_pipeline_matches_flat_and_overlaps_buffers(pkg, voc):
    """synthesize_host (pinned in, pinned out, download on a copy stream) == synthesize_flat, also when calls are
    issued back to back into alternating host buffers."""
    frames = [40, 57, 33]
    x = torch.from_numpy(np.concatenate([synth_logmel(T, 50 + i) for i, T in enumerate(frames)]))
    ph = torch.from_numpy(np.concatenate([seeded_phase(60 + i, T).T for i, T in enumerate(frames)]).astype(np.float32))
    want = voc.synthesize_flat(x.cuda(), frames, ph.cuda().contiguous(), n_iter=6).cpu()
    xh, phh = x.pin_memory(), ph.contiguous().pin_memory()
    outs = [torch.zeros(want.numel() + 7).pin_memory() for _ in range(2)]
    events = [voc.synthesize_host(xh, frames, outs[i % 2], phase_host=phh, n_iter=6) for i in range(4)]
    for e in events:
        e.synchronize()
    for o in outs:
        assert torch.equal(o[: want.numel()], want)
        assert float(o[want.numel():].abs().max()) == 0.0
