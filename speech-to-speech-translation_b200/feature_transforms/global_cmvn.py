"""Global CMVN transforms on the GPU: drop-in for
``fairseq/data/audio/feature_transforms/global_cmvn.py`` (``global_cmvn`` :8-29,
``src_global_cmvn`` :30-51, ``tgt_global_cmvn`` :52-74).

``__call__`` keeps the reference contract (numpy ``[T, n_feat]`` in, numpy out, ``(x - mean) / std``
with true division) but computes in the CUDA library (``s2st_cmvn_apply``); ``apply_cuda`` is the
batched form for features that are already device-resident (post-collate), which is where it pays.
"""
import numpy as np
import torch

from .. import _lib
from ..plans import require_cuda
from . import AudioFeatureTransform, register_audio_feature_transform


def cmvn_apply_cuda(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, out=None) -> torch.Tensor:
    """(x - mean) / std for x [..., n_feat] float32 on CUDA."""
    assert x.is_cuda and x.dtype == torch.float32
    x = x.contiguous()
    out = torch.empty_like(x) if out is None else out
    n_cols = x.shape[-1]
    with torch.cuda.device(x.device):
        rc = _lib.load().s2st_cmvn_apply(x.numel() // n_cols, n_cols, _lib.ptr(x), _lib.ptr(mean), _lib.ptr(std),
                                         _lib.ptr(out), _lib.stream_ptr(x.device))
    _lib.check(rc, "s2st_cmvn_apply")
    return out


def cmvn_denormalize_cuda(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, out=None) -> torch.Tensor:
    """x * std + mean (``gcmvn_denormalize``, fairseq/speech_generator_for_s2st.py:21-29)."""
    assert x.is_cuda and x.dtype == torch.float32
    x = x.contiguous()
    out = torch.empty_like(x) if out is None else out
    n_cols = x.shape[-1]
    with torch.cuda.device(x.device):
        rc = _lib.load().s2st_cmvn_denormalize(x.numel() // n_cols, n_cols, _lib.ptr(x), _lib.ptr(mean), _lib.ptr(std),
                                               _lib.ptr(out), _lib.stream_ptr(x.device))
    _lib.check(rc, "s2st_cmvn_denormalize")
    return out


class _GlobalCMVNBase(AudioFeatureTransform):
    """Global CMVN (cepstral mean and variance normalization). The global mean and variance need to be
    pre-computed and stored in NumPy format (.npz)."""

    @classmethod
    def from_config_dict(cls, config=None):
        _config = {} if config is None else config
        return cls(_config.get("stats_npz_path"))

    def __init__(self, stats_npz_path):
        self.stats_npz_path = stats_npz_path
        stats = np.load(stats_npz_path)
        self.mean, self.std = stats["mean"], stats["std"]
        self._dev = {}

    def __repr__(self):
        return self.__class__.__name__ + f'(stats_npz_path="{self.stats_npz_path}")'

    def _stats(self, device):
        key = (device.type, device.index)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(np.ascontiguousarray(self.mean, np.float32)).to(device),
                              torch.from_numpy(np.ascontiguousarray(self.std, np.float32)).to(device))
        return self._dev[key]

    def apply_cuda(self, x: torch.Tensor, frames=None) -> torch.Tensor:
        """x [..., n_feat] float32 on CUDA; ``frames`` (the ragged layout) is irrelevant to a global transform."""
        mean, std = self._stats(x.device)
        return cmvn_apply_cuda(x, mean, std)

    def __call__(self, x):
        dev = require_cuda(None)
        xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(dev)
        return self.apply_cuda(xd).cpu().numpy()

    def __getstate__(self):  # device caches do not survive pickling into DataLoader workers
        d = dict(self.__dict__)
        d["_dev"] = {}
        return d


@register_audio_feature_transform("global_cmvn")
class GlobalCMVN(_GlobalCMVNBase):
    pass


@register_audio_feature_transform("src_global_cmvn")
class SRCGlobalCMVN(_GlobalCMVNBase):
    pass


@register_audio_feature_transform("tgt_global_cmvn")
class TGTGlobalCMVN(_GlobalCMVNBase):
    pass
