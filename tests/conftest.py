import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "speech-to-speech-translation_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def built_lib(pkg):
    """The C-ABI library, built in-tree (nvcc cross-compiles without a GPU)."""
    import __graft_entry__
    __graft_entry__.build()
    return pkg._lib.load()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def golden_basis():
    b = load_golden("basis.npz")
    mel = np.zeros((80, 1025), np.float32)
    mel[b["mel_rows"], b["mel_cols"]] = b["mel_vals"]
    pinv = np.zeros((1025, 80), np.float32)
    pinv[: b["pinv_head"].shape[0]] = b["pinv_head"]
    return mel, pinv


def synth_logmel(T, seed, kind="smooth"):
    """SURVEY.md 8(d) synthetic log-mel (same generator as tests/golden/make_golden.py)."""
    import torch
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


def seeded_phase(seed, T, F=1025):
    np.random.seed(seed)
    return np.angle(np.exp(2j * np.pi * np.random.rand(F, T))).astype(np.float32)
