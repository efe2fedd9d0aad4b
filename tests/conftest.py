"""Shared fixtures.  GPU tests are marked `@pytest.mark.gpu` and call through the C-ABI;
everything else runs on CPU (the oracle against the golden vectors, host logic, symbol export)."""

import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def has_gpu() -> bool:
    try:
        from pocket_tts_mlx_b200 import _native
        return _native.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def bundle():
    """Synthetic checkpoint + voices + tokenizer + YAML (seed 0), written once per session."""
    from pocket_tts_mlx_b200.synthetic import default_bundle_dir, write_synthetic_bundle
    return write_synthetic_bundle(default_bundle_dir(), seed=0)


@pytest.fixture(scope="session")
def cfg(bundle):
    from pocket_tts_mlx_b200.config import load_config
    return load_config(bundle)


@pytest.fixture(scope="session")
def weights(cfg):
    from pocket_tts_mlx_b200.safetensors_io import read_safetensors
    return read_safetensors(cfg.weights_path)


@pytest.fixture(scope="session")
def voices(bundle):
    from pocket_tts_mlx_b200.safetensors_io import read_safetensors

    def get(name):
        return read_safetensors(Path(bundle).parent / "embeddings" / f"{name}.safetensors")["audio_prompt"]
    return get


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def snr_db(test, ref):
    test = np.asarray(test, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = ((test - ref) ** 2).sum()
    return float(10 * np.log10((ref ** 2).sum() / max(err, 1e-300)))
