"""Utterance-level CMVN on the GPU: drop-in for ``fairseq/data/audio/feature_transforms/utterance_cmvn.py``
(``utterance_cmvn`` :8-40).

``__call__`` keeps the reference contract (numpy ``[T, n_feat]`` in, numpy out) and is bit-identical to it: the
CUDA kernel accumulates the column sums in float32 in row order, which is what numpy's axis-0 reduction does.
``apply_cuda`` is the batched form for a ragged, device-resident batch (post-collate), one launch for all utterances.
"""
from typing import Sequence

import numpy as np
import torch

from .. import _lib
from ..plans import require_cuda, upload_small
from . import AudioFeatureTransform, register_audio_feature_transform


def utterance_cmvn_cuda(x: torch.Tensor, frames: Sequence[int], norm_means: bool = True, norm_vars: bool = True,
                        out=None) -> torch.Tensor:
    """x [sum T_i, n_feat] float32 on CUDA, utterance i owning the next T_i rows -> same shape."""
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
    x = x.contiguous()
    assert sum(frames) == x.shape[0], "frames must add up to the number of rows"
    out = torch.empty_like(x) if out is None else out
    fo = upload_small(np.concatenate([[0], np.cumsum(frames)]).astype(np.int32), x.device)
    stats = torch.empty(len(frames), 2, x.shape[1], dtype=torch.float32, device=x.device)  # (mean, std) per utterance
    with torch.cuda.device(x.device):
        rc = _lib.load().s2st_utterance_cmvn(len(frames), x.shape[0], _lib.ptr(fo), x.shape[1], _lib.ptr(x),
                                             int(bool(norm_means)), int(bool(norm_vars)), _lib.ptr(out), _lib.ptr(stats),
                                             _lib.stream_ptr(x.device))
    _lib.check(rc, "s2st_utterance_cmvn")
    return out


@register_audio_feature_transform("utterance_cmvn")
class UtteranceCMVN(AudioFeatureTransform):
    """Utterance-level CMVN (cepstral mean and variance normalization)"""

    @classmethod
    def from_config_dict(cls, config=None):
        _config = {} if config is None else config
        return UtteranceCMVN(
            _config.get("norm_means", True),
            _config.get("norm_vars", True),
        )

    def __init__(self, norm_means=True, norm_vars=True):
        self.norm_means, self.norm_vars = norm_means, norm_vars

    def __repr__(self):
        return self.__class__.__name__ + f"(norm_means={self.norm_means}, norm_vars={self.norm_vars})"

    def apply_cuda(self, x: torch.Tensor, frames: Sequence[int]) -> torch.Tensor:
        return utterance_cmvn_cuda(x, frames, self.norm_means, self.norm_vars)

    def __call__(self, x):
        dev = require_cuda(None)
        x = np.asarray(x)
        if x.shape[0] == 0 or not (self.norm_means or self.norm_vars):
            return x
        xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(dev)
        return self.apply_cuda(xd, [x.shape[0]]).cpu().numpy()
