"""Seeded synthetic assets: checkpoint, voice prompts, SentencePiece model, local YAML variant.

There is no network in the build/bench image, so the real kyutai checkpoint, voices and tokenizer
are not available.  Everything here is written in the *checkpoint's* (PyTorch) layout with the key
names the reference's loader walks (`pocket_tts_mlx/models/tts_model.py:153-194`; SURVEY.md
Appendix B), so the same files load in the reference and in this build.  The reference's default
initialisation is not used because it is degenerate (zero transposed-conv weights,
`pocket_tts_mlx/modules/conv.py:54-55`).

All draws come from `numpy.random.Generator(PCG64(seed))` in a fixed key order, so a seed names a
checkpoint on every machine with the same NumPy.
"""

from __future__ import annotations

import os
import tempfile
from pathlib import Path
from typing import Dict, Optional

import numpy as np
import yaml

from .config import Config, load_config
from .safetensors_io import write_safetensors

VOICE_NAMES = ["alba", "marius", "javert", "jean", "fantine", "cosette", "eponine", "azelma"]
_PKG_DIR = Path(__file__).resolve().parent
DEFAULT_TOKENIZER = _PKG_DIR / "assets" / "synthetic_tokenizer.model"


def _uniform(rng, shape, fan_in, gain=1.0):
    a = gain * np.sqrt(3.0 / fan_in)
    return rng.uniform(-a, a, size=shape).astype(np.float32)


def synthetic_state_dict(cfg: Config, seed: int = 0) -> Dict[str, np.ndarray]:
    """Variance-preserving random weights for every decode-path parameter (PyTorch layout)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: Dict[str, np.ndarray] = {}
    fl = cfg.flow_lm
    d = fl.transformer.d_model
    ldim = cfg.mimi.quantizer.dimension
    ff = d * fl.transformer.hidden_scale
    fd = fl.flow.dim

    def ln(prefix, n):
        sd[prefix + ".weight"] = rng.uniform(0.8, 1.2, size=n).astype(np.float32)
        sd[prefix + ".bias"] = (0.05 * rng.standard_normal(n)).astype(np.float32)

    def lin(prefix, n_out, n_in, bias=True, gain=1.0):
        sd[prefix + ".weight"] = _uniform(rng, (n_out, n_in), n_in, gain)
        if bias:
            sd[prefix + ".bias"] = _uniform(rng, (n_out,), n_in)

    # ---- FlowLM --------------------------------------------------------------------------
    sd["flow_lm.conditioner.embed.weight"] = rng.standard_normal(
        (fl.lookup_table.n_bins + 1, fl.lookup_table.dim)).astype(np.float32)
    sd["flow_lm.input_linear.weight"] = _uniform(rng, (d, ldim), ldim)
    sd["flow_lm.emb_std"] = rng.uniform(0.5, 1.5, size=ldim).astype(np.float32)
    sd["flow_lm.emb_mean"] = (0.1 * rng.standard_normal(ldim)).astype(np.float32)
    sd["flow_lm.bos_emb"] = rng.standard_normal(ldim).astype(np.float32)
    sd["flow_lm.speaker_proj_weight"] = _uniform(rng, (d, cfg.mimi.seanet.dimension), cfg.mimi.seanet.dimension)
    for i in range(fl.transformer.num_layers):
        p = f"flow_lm.transformer.layers.{i}"
        lin(p + ".self_attn.in_proj", 3 * d, d, bias=False)
        lin(p + ".self_attn.out_proj", d, d, bias=False, gain=0.7)
        ln(p + ".norm1", d)
        ln(p + ".norm2", d)
        lin(p + ".linear1", ff, d, bias=False)
        lin(p + ".linear2", d, ff, bias=False, gain=0.7)
    ln("flow_lm.out_norm", d)
    lin("flow_lm.out_eos", 1, d)
    for j in range(2):
        p = f"flow_lm.flow_net.time_embed.{j}.mlp"
        lin(p + ".0", fd, 256)
        lin(p + ".2", fd, fd)
        sd[p + ".3.alpha"] = rng.uniform(0.8, 1.2, size=fd).astype(np.float32)
    lin("flow_lm.flow_net.cond_embed", fd, d)
    lin("flow_lm.flow_net.input_proj", fd, ldim)
    for i in range(fl.flow.depth):
        p = f"flow_lm.flow_net.res_blocks.{i}"
        ln(p + ".in_ln", fd)
        lin(p + ".mlp.0", fd, fd)
        lin(p + ".mlp.2", fd, fd)
        lin(p + ".adaLN_modulation.1", 3 * fd, fd, gain=0.5)
    lin("flow_lm.flow_net.final_layer.linear", ldim, fd)
    lin("flow_lm.flow_net.final_layer.adaLN_modulation.1", 2 * fd, fd, gain=0.5)

    # ---- Mimi decode side ------------------------------------------------------------------
    mm = cfg.mimi
    dm = mm.transformer.d_model
    sd["mimi.quantizer.output_proj.weight"] = _uniform(
        rng, (mm.quantizer.output_dimension, mm.quantizer.dimension, 1), mm.quantizer.dimension)
    hop = int(np.prod(mm.seanet.ratios))
    up = int(round(mm.sample_rate / hop / mm.frame_rate))
    sd["mimi.upsample.convtr.convtr.weight"] = _uniform(rng, (mm.seanet.dimension, 1, 2 * up), 2)
    for i in range(mm.transformer.num_layers):
        p = f"mimi.decoder_transformer.transformer.layers.{i}"
        lin(p + ".self_attn.in_proj", 3 * dm, dm, bias=False)
        lin(p + ".self_attn.out_proj", dm, dm, bias=False)
        ln(p + ".norm1", dm)
        ln(p + ".norm2", dm)
        lin(p + ".linear1", mm.transformer.dim_feedforward, dm, bias=False)
        lin(p + ".linear2", dm, mm.transformer.dim_feedforward, bias=False)
        sd[p + ".layer_scale_1.scale"] = rng.uniform(0.05, 0.5, size=dm).astype(np.float32)
        sd[p + ".layer_scale_2.scale"] = rng.uniform(0.05, 0.5, size=dm).astype(np.float32)

    sn = mm.seanet
    assert sn.n_residual_layers == 1, "only the b6369a24 decoder topology is generated"

    def conv(prefix, c_out, c_in, k):
        sd[prefix + ".weight"] = _uniform(rng, (c_out, c_in, k), c_in * k)
        sd[prefix + ".bias"] = _uniform(rng, (c_out,), c_in * k)

    mult = 2 ** len(sn.ratios)
    idx = 0
    conv(f"mimi.decoder.model.{idx}.conv", mult * sn.n_filters, sn.dimension, sn.kernel_size)
    idx += 1
    for r in sn.ratios:
        c_in = mult * sn.n_filters
        c_out = c_in // 2
        idx += 1  # ELU
        sd[f"mimi.decoder.model.{idx}.convtr.weight"] = _uniform(rng, (c_in, c_out, 2 * r), 2 * c_in)
        sd[f"mimi.decoder.model.{idx}.convtr.bias"] = _uniform(rng, (c_out,), 2 * c_in)
        idx += 1
        hidden = c_out // sn.compress
        conv(f"mimi.decoder.model.{idx}.block.1.conv", hidden, c_out, sn.residual_kernel_size)
        conv(f"mimi.decoder.model.{idx}.block.3.conv", c_out, hidden, 1)
        idx += 1
        mult //= 2
    idx += 1  # ELU
    conv(f"mimi.decoder.model.{idx}.conv", sn.channels, sn.n_filters, sn.last_kernel_size)

    # ---- Mimi encode side (voice cloning; drawn after everything else so that the decode-path weights of a seed
    #      are the same with and without it).  Module tree of modules/seanet.py:45-108, models/mimi.py:28-52.
    mult = 1
    idx = 0
    conv(f"mimi.encoder.model.{idx}.conv", mult * sn.n_filters, sn.channels, sn.kernel_size)
    idx += 1
    for r in reversed(sn.ratios):
        c = mult * sn.n_filters
        hidden = c // sn.compress
        conv(f"mimi.encoder.model.{idx}.block.1.conv", hidden, c, sn.residual_kernel_size)
        conv(f"mimi.encoder.model.{idx}.block.3.conv", c, hidden, 1)
        idx += 2  # resblock, ELU
        conv(f"mimi.encoder.model.{idx}.conv", 2 * c, c, 2 * r)
        idx += 1
        mult *= 2
    idx += 1  # ELU
    conv(f"mimi.encoder.model.{idx}.conv", sn.dimension, mult * sn.n_filters, sn.last_kernel_size)
    for i in range(mm.transformer.num_layers):
        p = f"mimi.encoder_transformer.transformer.layers.{i}"
        lin(p + ".self_attn.in_proj", 3 * dm, dm, bias=False)
        lin(p + ".self_attn.out_proj", dm, dm, bias=False)
        ln(p + ".norm1", dm)
        ln(p + ".norm2", dm)
        lin(p + ".linear1", mm.transformer.dim_feedforward, dm, bias=False)
        lin(p + ".linear2", dm, mm.transformer.dim_feedforward, bias=False)
        sd[p + ".layer_scale_1.scale"] = rng.uniform(0.05, 0.5, size=dm).astype(np.float32)
        sd[p + ".layer_scale_2.scale"] = rng.uniform(0.05, 0.5, size=dm).astype(np.float32)
    sd["mimi.downsample.conv.conv.weight"] = _uniform(rng, (sn.dimension, sn.dimension, 2 * up), 2 * up * sn.dimension)
    return sd


def synthetic_voice(seed: int = 1, frames: int = 125, dim: int = 1024) -> np.ndarray:
    """A stand-in for `embeddings/<voice>.safetensors:audio_prompt` -> [1, frames, dim] float32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.standard_normal((1, frames, dim)).astype(np.float32)


def synthetic_token_ids(seed: int, batch: int, n_tok: int, n_bins: int = 4000) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, n_bins, size=(batch, n_tok), dtype=np.int64).astype(np.int32)


def _corpus_lines():
    """Prose-like lines from the Python standard library's docstrings (offline corpus)."""
    import re
    import sysconfig

    root = Path(sysconfig.get_paths()["stdlib"])
    lines = []
    for f in sorted(root.glob("*.py")):
        try:
            text = f.read_text(errors="ignore")
        except OSError:
            continue
        for ln in text.splitlines():
            s = ln.strip().lstrip("#").strip().strip('"').strip("'").strip()
            if len(s) > 30 and re.fullmatch(r"[A-Za-z0-9 ,.;:!?'()\-]+", s) and s.count(" ") >= 4:
                lines.append(s)
        if len(lines) > 20000:
            break
    lines += ["Hello from MLX!", "Hello world.", "What is this? It works! Yes... maybe."]
    return lines


def train_synthetic_tokenizer(path, vocab_size: int = 4000) -> Path:
    """Train a unigram SentencePiece model with exactly `vocab_size` pieces (reference asserts it,
    `pocket_tts_mlx/conditioners/text.py:21`)."""
    import sentencepiece as spm

    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        corpus = Path(td) / "corpus.txt"
        corpus.write_text("\n".join(_corpus_lines()))
        prefix = str(Path(td) / "sp")
        spm.SentencePieceTrainer.train(
            input=str(corpus), model_prefix=prefix, vocab_size=vocab_size, model_type="unigram",
            hard_vocab_limit=True, character_coverage=1.0, num_threads=1,
            minloglevel=2,
        )
        path.write_bytes(Path(prefix + ".model").read_bytes())
    return path


def write_synthetic_bundle(out_dir, base_variant: Optional[str] = None, seed: int = 0,
                           voice_frames: int = 125, bf16: bool = False) -> Path:
    """Write checkpoint + 8 voices + tokenizer + a YAML variant that points at them; return the YAML."""
    out_dir = Path(out_dir)
    (out_dir / "embeddings").mkdir(parents=True, exist_ok=True)
    base = Path(base_variant) if base_variant else _PKG_DIR / "config" / "b6369a24.yaml"
    cfg = load_config(base)
    ckpt = out_dir / f"tts_synthetic_enc_seed{seed}.safetensors"
    if not ckpt.exists():
        tmp = ckpt.with_suffix(f".tmp{os.getpid()}")
        write_safetensors(tmp, synthetic_state_dict(cfg, seed), bf16=bf16)
        os.replace(tmp, ckpt)            # atomic: concurrent ranks never see a partial file
    def atomic(path: Path, write) -> None:
        """Write through a per-process temporary file + rename: concurrent ranks (torchrun) never read a partial file."""
        tmp = path.with_name(f".{path.name}.tmp{os.getpid()}")
        write(tmp)
        os.replace(tmp, path)

    for i, name in enumerate(VOICE_NAMES):
        vp = out_dir / "embeddings" / f"{name}.safetensors"
        if not vp.exists():
            voice = synthetic_voice(1000 + i, voice_frames, cfg.flow_lm.transformer.d_model)
            atomic(vp, lambda t, voice=voice: write_safetensors(t, {"audio_prompt": voice}))
    tok = out_dir / "tokenizer.model"
    if not tok.exists():
        if DEFAULT_TOKENIZER.exists():
            atomic(tok, lambda t: t.write_bytes(DEFAULT_TOKENIZER.read_bytes()))
        else:
            atomic(tok, lambda t: train_synthetic_tokenizer(t, cfg.flow_lm.lookup_table.n_bins))
    doc = yaml.safe_load(base.read_text())
    doc["weights_path"] = str(ckpt)
    doc["weights_path_without_voice_cloning"] = str(ckpt)
    doc["flow_lm"]["lookup_table"]["tokenizer_path"] = str(tok)
    yml = out_dir / "synthetic.yaml"
    text = yaml.safe_dump(doc, sort_keys=False)
    if not (yml.exists() and yml.read_text() == text):
        atomic(yml, lambda t: t.write_text(text))
    return yml


def default_bundle_dir() -> Path:
    return Path(os.environ.get("POCKET_TTS_SYNTH_DIR", Path(tempfile.gettempdir()) / "pocket_tts_b200_synth"))
