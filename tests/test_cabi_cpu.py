"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/ptts.h declares; without a GPU it refuses to run (no CPU fallback); host-side mirrors behave."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import REPO, has_gpu


@pytest.fixture(scope="session")
def built_lib():
    from pocket_tts_mlx_b200.build_native import build
    return build()


def test_library_exports_every_declared_symbol(built_lib):
    header = (REPO / "include" / "ptts.h").read_text()
    declared = sorted(set(re.findall(r"\b(ptts_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(str(built_lib))
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    from pocket_tts_mlx_b200 import _native
    assert sorted(_native.EXPORTED) == declared
    assert _native.lib().ptts_abi_version() == 1


def test_config_struct_layout_matches_header():
    """Field order/types of the ctypes mirror follow the header's ptts_config."""
    from pocket_tts_mlx_b200 import _native
    header = (REPO / "include" / "ptts.h").read_text()
    body = header[header.index("typedef struct {") + len("typedef struct {"):header.index("} ptts_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        m = re.match(r"\s*(int32_t|int64_t|float)\s+(.*)", decl.strip(), flags=re.S)
        if not m:
            continue
        for part in m.group(2).split(","):
            names.append(re.sub(r"\[.*\]", "", part).strip())
    assert names == [f[0] for f in _native.Config._fields_]


def test_no_cpu_fallback(built_lib, bundle):
    if has_gpu():
        pytest.skip("GPU present")
    from pocket_tts_mlx_b200 import TTSModel, _native
    assert _native.device_count() == 0
    with pytest.raises(_native.PttsError, match="no CUDA device"):
        TTSModel.load_model(str(bundle))


def test_product_never_imports_oracle():
    for p in (REPO / "pocket_tts_mlx_b200").rglob("*.py"):
        src = p.read_text()
        assert "oracle" not in src.replace("# oracle", ""), p


def test_load_model_errors(tmp_path, bundle):
    from pydantic import ValidationError
    from pocket_tts_mlx_b200 import TTSModel
    import yaml
    with pytest.raises(FileNotFoundError):
        TTSModel.load_model(str(tmp_path / "missing.yaml"))
    doc = yaml.safe_load(Path(bundle).read_text())
    doc["unknown_key"] = 1
    bad = tmp_path / "bad.yaml"
    bad.write_text(yaml.safe_dump(doc))
    with pytest.raises(ValidationError):
        TTSModel.load_model(str(bad))
    doc.pop("unknown_key")
    doc["flow_lm"]["weights_path"] = "x.safetensors"
    half = tmp_path / "half.yaml"
    half.write_text(yaml.safe_dump(doc))
    with pytest.raises(ValueError, match="mimi.weights_path"):
        TTSModel.load_model(str(half))


def test_postprocess_matches_oracle():
    from oracle.ptts_oracle import postprocess_audio_start as ref
    from pocket_tts_mlx_b200.tts_model import postprocess_audio_start as mine
    rng = np.random.Generator(np.random.PCG64(0))
    a = rng.standard_normal(5000).astype(np.float32)
    for trim, fade in [(0, 0), (20, 15), (0, 30), (1000, 0), (10, 1000), (0, 1)]:
        assert np.array_equal(mine(a, 24000, trim, fade), ref(a, 24000, trim, fade))


def test_safetensors_roundtrip(tmp_path):
    from pocket_tts_mlx_b200.safetensors_io import (bf16_bits_to_f32, f32_to_bf16_bits, read_safetensors,
                                                    write_safetensors)
    rng = np.random.Generator(np.random.PCG64(1))
    t = {"a.weight": rng.standard_normal((3, 5)).astype(np.float32), "b": np.arange(7, dtype=np.int64)}
    write_safetensors(tmp_path / "x.safetensors", t)
    back = read_safetensors(tmp_path / "x.safetensors")
    assert np.array_equal(back["a.weight"], t["a.weight"]) and np.array_equal(back["b"], t["b"])
    write_safetensors(tmp_path / "y.safetensors", {"a": t["a.weight"]}, bf16=True)
    y = read_safetensors(tmp_path / "y.safetensors")["a"]
    assert y.dtype == np.float32
    assert np.array_equal(y, bf16_bits_to_f32(f32_to_bf16_bits(t["a.weight"])).reshape(3, 5))
    assert np.abs(y - t["a.weight"]).max() <= np.abs(t["a.weight"]).max() * 2 ** -8
    # safetensors' own reader agrees with ours
    from safetensors.numpy import load_file
    assert np.array_equal(load_file(str(tmp_path / "x.safetensors"))["a.weight"], t["a.weight"])


def test_cli_flags_and_exit_code(tmp_path, bundle, built_lib):
    """Same flags as the reference CLI; any failure (here: no GPU / bad config) maps to exit code 1."""
    from pocket_tts_mlx_b200.main import main
    rc = main(["hi", "--config", str(tmp_path / "nope.yaml"), "-o", str(tmp_path / "o.wav"), "--warmup-frames", "2",
               "--trim-start-ms", "5", "--fade-in-ms", "5", "--frames-after-eos", "3", "--max-tokens", "40",
               "--voice", "alba"])
    assert rc == 1
