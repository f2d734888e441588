"""On-disk formats either side of the path: npy-in-stored-ZIP manifests, the batched feature reader,
16-bit PCM WAV writing / reading.  Host logic runs on CPU; the PCM conversion and the batch writer need the GPU."""
import io
import wave
import zipfile

import numpy as np
import pytest
import torch

from conftest import synth_audio, synth_logmel


def _make_archive(tmp_path):
    rng = np.random.RandomState(5)
    feats = {"utt_%02d" % i: (rng.randn(T, 80) * 2 - 3).astype(np.float32) for i, T in enumerate((17, 1, 230, 56))}
    root = tmp_path / "feats"
    root.mkdir()
    for k, v in feats.items():
        np.save(root / (k + ".npy"), v[None] if k == "utt_02" else v)  # one file saved as [1, T, 80]
    return root, feats


def test_stored_zip_manifest_and_reader(pkg, tmp_path):
    root, feats = _make_archive(tmp_path)
    zp = tmp_path / "feats.zip"
    pkg.create_zip(root, zp)
    with zipfile.ZipFile(zp) as z:
        assert all(i.compress_type == zipfile.ZIP_STORED for i in z.infolist()) and len(z.infolist()) == 4
    paths, lengths = pkg.get_zip_manifest(zp)
    assert set(paths) == set(feats)
    assert lengths == {k: (1 if k == "utt_02" else v.shape[0]) for k, v in feats.items()}  # shape[0] like the reference
    for k, v in feats.items():
        f, ptr = pkg.parse_path(paths[k])
        assert f == zp.as_posix() and len(ptr) == 2
        data = pkg.read_from_stored_zip(f, ptr[0], ptr[1])
        assert pkg.is_npy_data(data) and not pkg.is_sf_audio_data(data)
        assert np.array_equal(np.load(io.BytesIO(data)).reshape(v.shape), v)
    assert pkg.parse_path(str(root / "utt_00.npy")) == (str(root / "utt_00.npy"), [])
    with pytest.raises(FileNotFoundError):
        pkg.parse_path(str(tmp_path / "missing.zip") + ":10:20")
    with pytest.raises(AssertionError):
        pkg.parse_path(zp.as_posix() + ":10")
    # batched reader: zip slices and plain files mixed, one buffer, frame-major
    order = ["utt_03", "utt_00", "utt_02", "utt_01"]
    batch, frames = pkg.load_feature_batch([paths[k] if k != "utt_00" else str(root / "utt_00.npy") for k in order], pin_memory=False)
    assert frames == [feats[k].shape[0] for k in order] and batch.shape == (sum(frames), 80) and batch.dtype == torch.float32
    assert np.array_equal(batch.numpy(), np.concatenate([feats[k] for k in order]))
    f64 = tmp_path / "f64.npy"
    np.save(f64, feats["utt_00"].astype(np.float64))
    b2, fr2 = pkg.load_feature_batch([str(f64)], pin_memory=False)
    assert np.array_equal(b2.numpy(), feats["utt_00"])


def test_wav16_reader_matches_stdlib(pkg, tmp_path):
    x = (synth_audio(3000, 24000, 1) * 32767).round().astype("<i2")
    p = tmp_path / "a.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(24000); w.writeframes(x.tobytes())
    y, sr = pkg.read_wav16(p)
    assert sr == 24000 and y.shape == (1, 3000) and np.array_equal(y[0], x.astype(np.float32) / 32768.0)
    au = __import__("importlib").import_module(pkg.__name__ + ".audio_utils")
    w2, sr2 = au.get_waveform(str(p), normalization=False, always_2d=False)  # soundfile is absent: the PCM16 parser
    assert sr2 == 24000 and np.array_equal(w2, x.astype(np.float32))
    with pytest.raises(NotImplementedError, match="resampling"):
        au.get_waveform(str(p), output_sample_rate=16000)
    assert pkg.is_sf_audio_data(p.read_bytes()[:8])


@pytest.mark.gpu
def test_write_wav_batch_from_synthesis(pkg, built_lib, tmp_path):
    """Synthesised batch -> PCM16 on the device -> one WAV per utterance; files parse with Python's wave module and
    hold lrint(x * 32767) saturated, like libsndfile's float -> short conversion."""
    x = torch.cat([torch.from_numpy(synth_audio(n, 24000, 7 + i)) for i, n in enumerate((4000, 1, 12345))])
    x[5], x[6], x[7] = 1.5, -2.0, float("nan")
    pcm = pkg.waves_to_pcm16(x.cuda()).cpu().numpy()
    want = np.clip(np.rint(np.nan_to_num(x.numpy().astype(np.float64) * np.float64(np.float32(32767.0)), nan=0.0)), -32768, 32767)
    want = np.clip(np.rint((x.numpy() * np.float32(32767.0)).astype(np.float64)), -32768, 32767)
    want[7] = 0
    assert pcm.dtype == np.int16 and np.array_equal(pcm, want.astype(np.int16))
    voc = pkg.GriffinLimVocoder(24000, 1200, 300, 2048, 80, 20, 8000, torch.hann_window, spec_bwd_max_iter=4).cuda()
    frames = [20, 33]
    flat = torch.cat([synth_logmel(T, 40 + i) for i, T in enumerate(frames)]).cuda()
    wavef = voc.synthesize_flat(flat, frames, None, seed=3)
    lens = [(T - 1) * 300 for T in frames]
    files = pkg.write_wav_batch(tmp_path / "wav_24000hz_griffin_lim", ["a", "b"], wavef, lens, 24000, output_sample_rate=24000)
    off = 0
    for f, n in zip(files, lens):
        with wave.open(str(f), "rb") as w:
            assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 24000, n)
            got = np.frombuffer(w.readframes(n), dtype="<i2")
        ref = np.clip(np.rint((wavef[off: off + n].cpu().numpy() * np.float32(32767.0)).astype(np.float64)), -32768, 32767)
        assert np.array_equal(got, ref.astype(np.int16))
        y, sr = pkg.read_wav16(f)
        assert sr == 24000 and np.array_equal(y[0], got.astype(np.float32) / 32768.0)
        off += n
    with pytest.raises(NotImplementedError):
        pkg.write_wav_batch(tmp_path / "x", ["a", "b"], wavef, lens, 24000, output_sample_rate=16000)


@pytest.mark.gpu
def test_pcm16_upload_path_equals_float_path(pkg, built_lib):
    """The front-end's input side: 16-bit PCM uploaded as 2 bytes per sample and converted on the device
    (s2st_pcm16_to_wave) is exactly what soundfile's float32 read gives (value / 32768, or the int16 value for the
    Kaldi-style fbank input, audio_utils.py:65-109), for aligned / unaligned and odd lengths, from pinned host memory
    and from a device tensor; and fbank80 of the converted samples equals fbank80 of the float samples bitwise."""
    rng = np.random.RandomState(3)
    for n in (1, 7, 8, 4001, 160000):
        pcm = rng.randint(-32768, 32768, n).astype(np.int16)
        pcm[:3] = (-32768, 32767, 0)[: min(3, n)]
        for norm in (True, False):
            want = pcm.astype(np.float32) / np.float32(32768.0) if norm else pcm.astype(np.float32)
            a = pkg.pcm16_to_waves(torch.from_numpy(pcm).pin_memory(), normalization=norm).cpu().numpy()
            b = pkg.pcm16_to_waves(torch.from_numpy(pcm).cuda()[1:] if n > 1 else torch.from_numpy(pcm).cuda(), normalization=norm).cpu().numpy()
            assert a.dtype == np.float32 and np.array_equal(a, want)
            assert np.array_equal(b, want[1:] if n > 1 else want)  # a 2-byte aligned (not 16-byte aligned) view
    pcm = (synth_audio(32000, 16000, 5) * 20000).astype(np.int16)
    wave_f = torch.from_numpy(pcm.astype(np.float32)).cuda()
    wave_p = pkg.pcm16_to_waves(torch.from_numpy(pcm).pin_memory(), normalization=False)
    fa = pkg.fbank_batch([wave_f], 16000)[0]
    fb = pkg.fbank_batch([wave_p], 16000)[0]
    assert torch.equal(fa, fb)
