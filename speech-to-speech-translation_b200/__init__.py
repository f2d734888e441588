"""B200-native (sm_100a) waveform-synthesis and speech-feature front-end.

Drop-in for the Griffin-Lim vocoder and the fbank80 / logmelspec80 / global-CMVN front-end of
fengpeng-yue/speech-to-speech-translation (a fairseq fork), plus the rows next to that path: the other registry
transforms (utterance CMVN, SpecAugment masking) and the DTW of the MCD validation metric (``mcd``).  Host code is Python / PyTorch (device
memory, streams, torch.distributed); the arithmetic is a hand-written CUDA library behind a C ABI
(``include/s2st_b200.h``, loaded with ctypes from ``libs2st_b200.so`` in this directory).  There is
no CPU fallback: without the built library every entry point raises.

The directory name is not a Python identifier; import it as ``import s2st_b200`` (alias module at
the repository root) or with ``importlib.import_module("speech-to-speech-translation_b200")``.
"""
from . import _lib  # noqa: F401
from . import mcd  # noqa: F401
from .audio_utils import (TTSMelScale, TTSSpectrogram, fbank_batch, get_fbank, get_fourier_basis,  # noqa: F401
                          get_mel_filters, get_window)
from .feature_transforms import (AudioFeatureTransform, CompositeAudioFeatureTransform,  # noqa: F401
                                 get_audio_feature_transform, register_audio_feature_transform)
from .feature_transforms.global_cmvn import GlobalCMVN, SRCGlobalCMVN, TGTGlobalCMVN  # noqa: F401
from .feature_transforms.specaugment import SpecAugmentTransform  # noqa: F401
from .feature_transforms.utterance_cmvn import UtteranceCMVN  # noqa: F401
from .features import (extract_fbank_features, extract_logmel_spectrogram, gcmvn_denormalize,  # noqa: F401
                       get_global_cmvn, global_cmvn_from_sums, global_cmvn_stats, logmel_batch)
from .io_utils import (create_zip, get_zip_manifest, is_npy_data, is_sf_audio_data, load_feature_batch,  # noqa: F401
                       mmap_read, parse_path, pcm16_to_waves, read_from_stored_zip, read_wav16, waves_to_pcm16,
                       write_wav_batch)
from .vocoder import GriffinLim, GriffinLimVocoder, PseudoInverseMelScale, get_vocoder  # noqa: F401

__version__ = "0.1.0"
