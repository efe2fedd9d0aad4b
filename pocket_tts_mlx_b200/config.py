"""YAML model-variant schema (strict: unknown keys are rejected).

Accepts the same files as the reference's `pocket_tts_mlx/utils/config.py:9-128`
(`Config(flow_lm=..., mimi=..., weights_path=..., weights_path_without_voice_cloning=...)`),
raises `FileNotFoundError` for a missing file and pydantic `ValidationError` for unknown keys.
"""

from __future__ import annotations

from pathlib import Path
from typing import Optional, Tuple

import yaml
from pydantic import BaseModel, ConfigDict


class _Strict(BaseModel):
    model_config = ConfigDict(extra="forbid")


class FlowConfig(_Strict):
    dim: int
    depth: int


class FlowLMTransformerConfig(_Strict):
    hidden_scale: int
    max_period: int
    d_model: int
    num_heads: int
    num_layers: int


class LookupTable(_Strict):
    dim: int
    n_bins: int
    tokenizer: str
    tokenizer_path: str


class FlowLMConfig(_Strict):
    dtype: str
    flow: FlowConfig
    transformer: FlowLMTransformerConfig
    lookup_table: LookupTable
    weights_path: Optional[str] = None


class SEANetConfig(_Strict):
    dimension: int
    channels: int
    n_filters: int
    n_residual_layers: int
    ratios: list[int]
    kernel_size: int
    residual_kernel_size: int
    last_kernel_size: int
    dilation_base: int
    pad_mode: str
    compress: int


class MimiTransformerConfig(_Strict):
    d_model: int
    input_dimension: int
    output_dimensions: Tuple[int, ...]
    num_heads: int
    num_layers: int
    layer_scale: float
    context: int
    max_period: float = 10000.0
    dim_feedforward: int


class QuantizerConfig(_Strict):
    dimension: int
    output_dimension: int


class MimiConfig(_Strict):
    dtype: str
    sample_rate: int
    channels: int
    frame_rate: float
    seanet: SEANetConfig
    transformer: MimiTransformerConfig
    quantizer: QuantizerConfig
    weights_path: Optional[str] = None


class Config(_Strict):
    flow_lm: FlowLMConfig
    mimi: MimiConfig
    weights_path: Optional[str] = None
    weights_path_without_voice_cloning: Optional[str] = None


def load_config(yaml_path) -> Config:
    yaml_path = Path(yaml_path)
    if not yaml_path.exists():
        raise FileNotFoundError(f"Config file not found: {yaml_path}")
    with open(yaml_path, "r") as f:
        return Config(**yaml.safe_load(f))
