"""On-disk formats either side of the hot path, host side.

* feature input: ``.npy`` files, or ``.npy`` members of an UNCOMPRESSED (stored) ZIP addressed as
  ``"<zip path>:<byte offset>:<byte length>"`` -- the manifest format of the reference
  (``fairseq/data/audio/audio_utils.py:171-215``: ``is_npy_data``, ``is_sf_audio_data``, ``mmap_read``,
  ``read_from_stored_zip``, ``parse_path``; ``examples/speech_to_text/data_utils.py:101-132``: ``create_zip``,
  ``get_zip_manifest``).  Same names and behaviour here, plus ``load_feature_batch``: the data-parallel reader that
  puts a whole batch of utterances into ONE pinned ``[sum T, n_feat]`` buffer (the layout ``synthesize_host`` and the
  CMVN kernels take) by parsing the npy headers itself and copying the payloads straight out of the memory-mapped
  archive, instead of one ``np.load(io.BytesIO(...))`` per utterance.
* waveform output: what ``examples/s2s_trans/generate_waveform.py:115-124`` does with ``sf.write`` (float waveform ->
  16-bit PCM WAV).  ``write_wav_batch`` converts the whole concatenated batch to PCM16 on the GPU
  (``s2st_wave_to_pcm16``), downloads 2 bytes per sample and writes one RIFF file per utterance; soundfile is not
  needed.  ``read_wav16`` parses such files back (tests, and the fbank path when soundfile is absent).

Resampling (``--output-sample-rate`` other than the vocoder's, sox ``rate`` in the reference) is third-party
arithmetic that is not reproduced: asking for it raises ``NotImplementedError``.
"""
import io
import mmap
import struct
import zipfile
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .plans import require_cuda

FEATURE_OR_SF_AUDIO_FILE_EXTENSIONS = {".npy", ".wav", ".flac", ".ogg"}


_NPY_MAGIC = b"\x93N"                       # "\x93NUMPY": bytes 147, 78
_AUDIO_MAGICS = (b"RIF", b"fLa", b"Ogg")     # RIFF/WAVE, FLAC, Ogg containers


def is_npy_data(data: bytes) -> bool:
    return bytes(data[:2]) == _NPY_MAGIC


def is_sf_audio_data(data: bytes) -> bool:
    return bytes(data[:3]) in _AUDIO_MAGICS


def mmap_read(path: str, offset: int, length: int) -> bytes:
    """``length`` bytes at ``offset`` of a file through a read-only mapping (no read of the rest of the archive)."""
    with open(path, "rb") as f, mmap.mmap(f.fileno(), length=0, access=mmap.ACCESS_READ) as m:
        return m[offset: offset + length]


def read_from_stored_zip(zip_path: str, offset: int, length: int) -> bytes:
    return mmap_read(zip_path, offset, length)


def parse_path(path: str) -> Tuple[str, List[int]]:
    """A manifest entry is a plain ``.npy/.wav/.flac/.ogg`` file or a slice of a stored ZIP written as
    ``"<zip path>:<byte offset>:<byte length>"``.  Returns (file path, [] or [offset, length]); a sliced entry whose
    archive does not exist raises ``FileNotFoundError``, a malformed slice ``AssertionError`` -- as in the reference."""
    if Path(path).suffix in FEATURE_OR_SF_AUDIO_FILE_EXTENSIONS:
        return path, []
    file_path, *fields = path.split(":")
    if not Path(file_path).is_file():
        raise FileNotFoundError(f"File not found: {file_path}")
    assert len(fields) in {0, 2}, f"Invalid path: {path}"
    return file_path, [int(v) for v in fields]


def create_zip(data_root: Path, zip_path: Path):
    """All ``*.npy`` files of ``data_root`` into one stored (uncompressed) archive."""
    with zipfile.ZipFile(zip_path, "w", zipfile.ZIP_STORED) as f:
        for path in list(Path(data_root).glob("*.npy")):
            f.write(path, arcname=path.name)


def get_zip_manifest(zip_path: Path, zip_root: Optional[Path] = None, is_audio: bool = False):
    """utterance id -> ``"<zip path>:<offset>:<size>"`` and id -> length (frames of a feature matrix, samples of a
    16-bit WAV) for every member of a stored ZIP.  The archive is mapped once; npy lengths come from the header alone."""
    _zip_path = Path.joinpath(zip_root or Path(""), zip_path)
    with zipfile.ZipFile(_zip_path, mode="r") as f:
        info = f.infolist()
    paths, lengths = {}, {}
    with open(_zip_path, "rb") as fh, mmap.mmap(fh.fileno(), length=0, access=mmap.ACCESS_READ) as mm:
        for i in info:
            utt_id = Path(i.filename).stem
            offset, file_size = i.header_offset + 30 + len(i.filename), i.file_size
            paths[utt_id] = f"{zip_path.as_posix()}:{offset}:{file_size}"
            head = mm[offset: offset + min(file_size, 4096)]
            assert len(head) > 1
            if is_audio:
                assert is_sf_audio_data(head), i
                lengths[utt_id] = _wav_info(mm, offset, file_size)[2]
            else:
                assert is_npy_data(head), i
                lengths[utt_id] = _npy_header(head)[0][0]
    return paths, lengths


def _npy_header(head: bytes):
    """(shape, dtype, fortran_order, data offset) of an npy blob from its first bytes (format 1.0 - 3.0)."""
    f = io.BytesIO(head)
    version = np.lib.format.read_magic(f)
    if version == (1, 0):
        shape, fortran, dtype = np.lib.format.read_array_header_1_0(f)
    else:
        shape, fortran, dtype = np.lib.format.read_array_header_2_0(f)
    return shape, dtype, fortran, f.tell()


def load_feature_batch(paths: Sequence[str], pin_memory: bool = True):
    """Feature matrices named by manifest paths -> (features [sum T, n_feat] float32 in one (pinned) buffer, frames).

    Every archive is memory-mapped once per call; each payload is copied exactly once, from the mapping into its rows
    of the batch buffer (float32 C-order payloads; anything else goes through ``np.load`` and a cast)."""
    maps: Dict[str, Tuple[object, mmap.mmap]] = {}
    try:
        items = []
        for p in paths:
            _path, ptr = parse_path(p)
            if _path not in maps:
                fh = open(_path, "rb")
                maps[_path] = (fh, mmap.mmap(fh.fileno(), length=0, access=mmap.ACCESS_READ))
            mm = maps[_path][1]
            off, size = (ptr[0], ptr[1]) if ptr else (0, mm.size())
            head = mm[off: off + min(size, 4096)]
            if not is_npy_data(head):
                raise ValueError(f'"{p}" is not npy data')
            shape, dtype, fortran, data_off = _npy_header(head)
            items.append((mm, off, size, shape, dtype, fortran, data_off))
        squeezed = [tuple(d for d in it[3] if d != 1) if len(it[3]) > 2 else it[3] for it in items]
        if any(len(s) != 2 for s in squeezed):
            raise ValueError("expected [T, n_feat] feature matrices")
        n_feat = squeezed[0][1]
        if any(s[1] != n_feat for s in squeezed):
            raise ValueError("feature files disagree on the number of columns")
        frames = [int(s[0]) for s in squeezed]
        out = torch.empty(sum(frames), n_feat, dtype=torch.float32, pin_memory=pin_memory and torch.cuda.is_available())
        dst = out.numpy()
        row = 0
        for (mm, off, size, shape, dtype, fortran, data_off), T in zip(items, frames):
            if dtype == np.float32 and not fortran:
                src = np.frombuffer(mm, dtype=np.float32, count=T * n_feat, offset=off + data_off)
                dst[row: row + T] = src.reshape(T, n_feat)
                del src  # the view pins the mapping: release it before the archive is closed
            else:
                dst[row: row + T] = np.load(io.BytesIO(mm[off: off + size])).reshape(T, n_feat).astype(np.float32)
            row += T
        return out, frames
    finally:
        for fh, mm in maps.values():
            mm.close()
            fh.close()


# ---- 16-bit PCM WAV ----------------------------------------------------------------------------------------------
def _wav_header(n_samples: int, sample_rate: int, channels: int = 1) -> bytes:
    data_bytes = n_samples * channels * 2
    return (b"RIFF" + struct.pack("<I", 36 + data_bytes) + b"WAVE" + b"fmt " +
            struct.pack("<IHHIIHH", 16, 1, channels, sample_rate, sample_rate * channels * 2, channels * 2, 16) +
            b"data" + struct.pack("<I", data_bytes))


def _wav_info(buf, offset: int = 0, size: Optional[int] = None):
    """(channels, sample_rate, frames, data offset) of a PCM16 RIFF blob inside ``buf``."""
    end = offset + (size if size is not None else len(buf) - offset)
    if buf[offset: offset + 4] != b"RIFF" or buf[offset + 8: offset + 12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    pos, fmt = offset + 12, None
    while pos + 8 <= end:
        cid, csize = buf[pos: pos + 4], struct.unpack("<I", buf[pos + 4: pos + 8])[0]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", buf[pos + 8: pos + 24])
        elif cid == b"data":
            if fmt is None or fmt[0] != 1 or fmt[5] != 16:
                raise ValueError("only 16-bit PCM WAV is supported without soundfile")
            return fmt[1], fmt[2], min(csize, end - pos - 8) // (2 * fmt[1]), pos + 8
        pos += 8 + csize + (csize & 1)
    raise ValueError("no data chunk")


def read_wav16(path_or_bytes, normalization: bool = True) -> Tuple[np.ndarray, int]:
    """16-bit PCM WAV -> (waveform [channels, n] float32, sample rate); values / 32768 like soundfile's float32 read
    when ``normalization`` (audio_utils.py:65-109), int16-scaled otherwise."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else Path(path_or_bytes).read_bytes()
    ch, sr, frames, off = _wav_info(data)
    x = np.frombuffer(data, dtype="<i2", count=frames * ch, offset=off).reshape(frames, ch).T.astype(np.float32)
    return (x / 32768.0 if normalization else x), sr


def waves_to_pcm16(wave_flat: torch.Tensor) -> torch.Tensor:
    """Concatenated float32 CUDA waveforms -> int16 PCM on the device (``s2st_wave_to_pcm16``)."""
    dev = require_cuda(wave_flat.device)
    w = wave_flat.detach().to(dev, torch.float32).contiguous()
    out = torch.empty(w.numel(), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().s2st_wave_to_pcm16(w.numel(), _lib.ptr(w), _lib.ptr(out), _lib.stream_ptr(dev))
    _lib.check(rc, "s2st_wave_to_pcm16")
    return out


def pcm16_to_waves(pcm: torch.Tensor, normalization: bool = True, device=None) -> torch.Tensor:
    """int16 PCM (a CUDA tensor, or a -- preferably pinned -- host tensor, which is uploaded as 2 bytes per sample) ->
    float32 CUDA waveform (``s2st_pcm16_to_wave``): value / 32768 like ``get_waveform(normalization=True)``
    (audio_utils.py:65-109), the int16 value itself otherwise (the input ``get_fbank`` feeds Kaldi's fbank)."""
    assert pcm.dtype == torch.int16
    dev = require_cuda(pcm.device if pcm.is_cuda else device)
    p = pcm.to(dev, non_blocking=True).contiguous()
    out = torch.empty(p.numel(), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().s2st_pcm16_to_wave(p.numel(), _lib.ptr(p), 1.0 / 32768.0 if normalization else 1.0, _lib.ptr(out),
                                            _lib.stream_ptr(dev))
    _lib.check(rc, "s2st_pcm16_to_wave")
    return out.view(pcm.shape)


def write_wav_batch(out_dir: Path, sample_ids: Sequence[str], wave_flat: torch.Tensor, lengths: Sequence[int],
                    sample_rate: int, output_sample_rate: Optional[int] = None, ext: str = "wav") -> List[Path]:
    """``dump_result``'s waveform branch (generate_waveform.py:115-124) for a synthesised batch: utterance i is
    ``lengths[i]`` samples of the concatenated CUDA tensor ``wave_flat``; one 16-bit PCM file per utterance in
    ``out_dir``.  The float -> PCM conversion runs on the GPU and the batch comes down in one 2-byte-per-sample copy."""
    if output_sample_rate is not None and output_sample_rate != sample_rate:
        raise NotImplementedError("resampling (sox 'rate' in the reference) is not provided: write at the vocoder's "
                                  f"sample rate {sample_rate} or resample afterwards")
    if ext != "wav":
        raise NotImplementedError("only 16-bit PCM WAV is written (FLAC needs libsndfile)")
    assert len(sample_ids) == len(lengths) and int(sum(lengths)) == wave_flat.numel()
    pcm = waves_to_pcm16(wave_flat)
    host = torch.empty(pcm.numel(), dtype=torch.int16, pin_memory=True)
    host.copy_(pcm, non_blocking=True)
    torch.cuda.current_stream(pcm.device).synchronize()
    raw = host.numpy()
    out_dir = Path(out_dir)
    out_dir.mkdir(exist_ok=True, parents=True)
    written, off = [], 0
    for sid, n in zip(sample_ids, lengths):
        path = out_dir / f"{sid}.{ext}"
        with open(path, "wb") as f:
            f.write(_wav_header(int(n), int(sample_rate)))
            f.write(raw[off: off + n].astype("<i2", copy=False).tobytes())
        written.append(path)
        off += n
    return written
