"""ctypes binding of libs2st_b200.so (the C ABI declared in include/s2st_b200.h).

There is no CPU or eager-PyTorch fallback anywhere in this package: if the CUDA
library has not been built (``python -c "import __graft_entry__ as g; g.build()"``
or ``python speech-to-speech-translation_b200/build.py``) every product entry
point raises ``RuntimeError``.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libs2st_b200.so"
# S2ST_B200_LIB points the loader at another build of the same library (kernel A/B experiments in tools/)
LIB_PATH = os.environ.get("S2ST_B200_LIB") or os.path.join(_HERE, LIB_NAME)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "s2st_b200.h")

S2ST_OK = 0
ABI_VERSION = 2
# s2st_plan_set_option keys (include/s2st_b200.h)
OPT_GL_PDL, OPT_INVERSE_MEL, OPT_FRONTEND_GENERIC, OPT_MEL_PROJECT, OPT_GL_FRAMES = 2, 4, 5, 6, 7

c_f32p = ctypes.c_void_p  # device pointers are passed as integers (tensor.data_ptr())

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "s2st_abi_version": (ctypes.c_int, []),
    "s2st_last_error": (ctypes.c_char_p, []),
    "s2st_plan_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "s2st_plan_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "s2st_plan_active_bins": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]),
    "s2st_gl_workspace_bytes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                                ctypes.POINTER(ctypes.c_size_t)]),
    "s2st_gl_synthesize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "s2st_plan_set_strip_frames": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "s2st_plan_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "s2st_fbank_plan_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "s2st_phase_from_uniform": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p]),
    "s2st_phase_from_mt19937": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_plan_set_pass_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "s2st_plan_get_pass_times": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                                 ctypes.POINTER(ctypes.c_int)]),
    "s2st_gl_launch_count": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(ctypes.c_int)]),
    "s2st_inverse_mel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_mel_project": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    "s2st_stft": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_istft": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                   ctypes.c_void_p]),
    "s2st_rfft2048": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p]),
    "s2st_irfft2048": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "s2st_window_sum_square": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_logmel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_fbank_plan_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int]),
    "s2st_fbank_plan_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "s2st_fbank_frame_params": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int),
                                                ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "s2st_fbank": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p]),
    "s2st_cmvn_apply": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_cmvn_denormalize": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_cmvn_accumulate": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "s2st_utterance_cmvn": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_utterance_sums": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_utterance_sum": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_fill_rects": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p]),
    "s2st_dtw": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_time_warp": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_wave_to_pcm16": (ctypes.c_int, [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_pcm16_to_wave": (ctypes.c_int, [ctypes.c_int64, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_rms_dist_batch": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "s2st_rms_dist": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p]),
}

_lock = threading.Lock()
_lib = None


class S2STLibraryError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raise if the CUDA library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise S2STLibraryError(
                f"{LIB_PATH} not found: the sm_100a CUDA library has not been built "
                "(run __graft_entry__.build()). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.s2st_abi_version() != ABI_VERSION:
            raise S2STLibraryError("libs2st_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != S2ST_OK:
        msg = load().s2st_last_error()
        msg = msg.decode("utf-8", "replace") if msg else ""
        err = {1: ValueError, 2: RuntimeError, 3: RuntimeError}.get(rc, RuntimeError)
        raise err(f"{what} failed (status {rc}): {msg}")


def ptr(t):
    """Device / host pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "s2st_b200 needs contiguous tensors"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
