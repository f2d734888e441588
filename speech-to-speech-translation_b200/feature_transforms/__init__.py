"""Audio feature-transform registry: drop-in for ``fairseq/data/audio/feature_transforms/__init__.py``.

Same decorator, lookup and composite API (``register_audio_feature_transform`` :18-36,
``get_audio_feature_transform`` :39-40, auto-import of sibling modules :43-52,
``CompositeAudioFeatureTransform`` incl. the fork's ``from_config_dict_for_src`` /
``from_config_dict_for_tgt`` :55-106) and the same ``ValueError`` behaviour for duplicate names,
duplicate class names and non-subclasses.  The registered CMVN transforms run on the GPU.
"""
import importlib
import pkgutil
from abc import ABC, abstractmethod
from typing import Dict, Optional


class AudioFeatureTransform(ABC):
    @classmethod
    @abstractmethod
    def from_config_dict(cls, config: Optional[Dict] = None):
        pass


AUDIO_FEATURE_TRANSFORM_REGISTRY = {}
AUDIO_FEATURE_TRANSFORM_CLASS_NAMES = set()


def _registration_error(name, cls):
    """The three ways a registration can be refused (feature_transforms/__init__.py:20-31), same messages."""
    if name in AUDIO_FEATURE_TRANSFORM_REGISTRY:
        return f"Cannot register duplicate transform ({name})"
    if not (isinstance(cls, type) and issubclass(cls, AudioFeatureTransform)):
        return f"Transform ({name}: {cls.__name__}) must extend AudioFeatureTransform"
    if cls.__name__ in AUDIO_FEATURE_TRANSFORM_CLASS_NAMES:
        return f"Cannot register audio feature transform with duplicate class name ({cls.__name__})"
    return None


def register_audio_feature_transform(name):
    def decorator(cls):
        problem = _registration_error(name, cls)
        if problem is not None:
            raise ValueError(problem)
        AUDIO_FEATURE_TRANSFORM_CLASS_NAMES.add(cls.__name__)
        AUDIO_FEATURE_TRANSFORM_REGISTRY[name] = cls
        return cls

    return decorator


def get_audio_feature_transform(name):
    return AUDIO_FEATURE_TRANSFORM_REGISTRY[name]


class CompositeAudioFeatureTransform(AudioFeatureTransform):
    @classmethod
    def _build(cls, config, key):
        names = ({} if config is None else config).get(key)
        if names is None:
            return None
        cfg = {} if config is None else config
        return cls([get_audio_feature_transform(n).from_config_dict(cfg.get(n)) for n in names])

    @classmethod
    def from_config_dict(cls, config=None):
        return cls._build(config, "transforms")

    @classmethod
    def from_config_dict_for_src(cls, config=None):
        return cls._build(config, "src_transforms")

    @classmethod
    def from_config_dict_for_tgt(cls, config=None):
        return cls._build(config, "tgt_transforms")

    def __init__(self, transforms):
        self.transforms = [t for t in transforms if t is not None]

    def __call__(self, x):
        for t in self.transforms:
            x = t(x)
        return x

    def apply_cuda(self, x, frames):
        """The chain on a ragged, device-resident batch (x [sum T_i, n_feat] float32 CUDA, utterance i owning the next
        frames[i] rows): one or two launches per transform for the whole batch instead of a numpy round trip per
        utterance.  Every registered transform of this package implements ``apply_cuda(x, frames)``."""
        for t in self.transforms:
            x = t.apply_cuda(x, frames)
        return x

    def apply_cuda_from_host(self, feats, device=None):
        """The post-collate entry for DataLoader pipelines: ``feats`` is the list of per-utterance ``[T_i, n_feat]`` numpy
        arrays a transform-free dataset produced in its (forked) workers.  They are concatenated into one pinned buffer,
        uploaded once, run through the whole chain on the device (``apply_cuda``) and returned as a list of CUDA tensors
        (views of one buffer) -- the numpy ``__call__`` per utterance inside the workers is not needed at all."""
        import numpy as np
        import torch

        from ..plans import require_cuda
        dev = require_cuda(device)
        frames = [int(f.shape[0]) for f in feats]
        n_feat = int(feats[0].shape[1])
        staged = torch.empty(sum(frames), n_feat, dtype=torch.float32, pin_memory=True)
        np.concatenate([np.asarray(f, np.float32) for f in feats], out=staged.numpy())
        x = self.apply_cuda(staged.to(dev, non_blocking=True), frames)
        return list(torch.split(x, frames))

    def __repr__(self):
        lines = [self.__class__.__name__ + "("] + [f"    {t!r}" for t in self.transforms] + [")"]
        return "\n".join(lines)


# import every public sibling module so its transforms register themselves
for _m in pkgutil.iter_modules(__path__):
    if not _m.name.startswith("_"):
        importlib.import_module(f"{__name__}.{_m.name}")
