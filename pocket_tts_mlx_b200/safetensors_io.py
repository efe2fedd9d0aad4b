"""Torch-free safetensors reader/writer (host side of the weight-loading boundary).

Replaces the reader the reference ships in `pocket_tts_mlx/utils/weight_conversion.py:38-69`
(8-byte little-endian header length, JSON header, raw little-endian tensor bytes; BF16 widened to
float32 by a 16-bit left shift).  The writer is only used for synthetic checkpoints and tests.
"""

from __future__ import annotations

import json
import mmap
import struct
from pathlib import Path
from typing import Dict, Mapping

import numpy as np

_ST_TO_NP = {
    "F64": "<f8", "F32": "<f4", "F16": "<f2",
    "I64": "<i8", "I32": "<i4", "I16": "<i2", "I8": "i1",
    "U64": "<u8", "U32": "<u4", "U16": "<u2", "U8": "u1", "BOOL": "?",
}
_NP_TO_ST = {np.dtype(v).str: k for k, v in _ST_TO_NP.items()}


def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    """uint16 bfloat16 payload -> float32 (exact)."""
    return (bits.astype(np.uint32) << np.uint32(16)).view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 payload with round-to-nearest-even (what `__float2bfloat16_rn` does)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    out = (rounded >> np.uint32(16)).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        out = np.where(nan, np.uint16(0x7FC0), out)
    return out


def read_safetensors(path) -> Dict[str, np.ndarray]:
    """Return {name: ndarray}; BF16 tensors come back as float32, everything else as stored."""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"safetensors file not found: {path}")
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen).decode("utf-8"))
        base = 8 + hlen
        buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            for name, meta in header.items():
                if name == "__metadata__":
                    continue
                lo, hi = meta["data_offsets"]
                shape = tuple(meta["shape"])
                kind = meta["dtype"]
                raw = buf[base + lo: base + hi]
                if kind == "BF16":
                    arr = bf16_bits_to_f32(np.frombuffer(raw, dtype="<u2")).reshape(shape)
                elif kind in _ST_TO_NP:
                    arr = np.frombuffer(raw, dtype=_ST_TO_NP[kind]).reshape(shape).copy()
                else:
                    raise ValueError(f"Unsupported safetensors dtype: {kind}")
                out[name] = arr
        finally:
            buf.close()
    return out


def write_safetensors(path, tensors: Mapping[str, np.ndarray], bf16: bool = False) -> None:
    """Write {name: ndarray}.  With bf16=True float32 tensors are stored as BF16."""
    header = {}
    blobs = []
    off = 0
    for name, arr in tensors.items():
        arr = np.ascontiguousarray(arr)
        if bf16 and arr.dtype == np.float32:
            payload = f32_to_bf16_bits(arr).tobytes()
            kind = "BF16"
        else:
            key = arr.dtype.newbyteorder("<").str if arr.dtype.byteorder == ">" else arr.dtype.str
            if key not in _NP_TO_ST:
                raise ValueError(f"cannot store dtype {arr.dtype}")
            payload = arr.tobytes()
            kind = _NP_TO_ST[key]
        header[name] = {"dtype": kind, "shape": list(arr.shape), "data_offsets": [off, off + len(payload)]}
        blobs.append(payload)
        off += len(payload)
    hjson = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hjson += b" " * ((8 - len(hjson) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hjson)))
        f.write(hjson)
        for b in blobs:
            f.write(b)
