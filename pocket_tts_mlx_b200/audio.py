"""Host-side audio helpers of the voice-cloning branch (NumPy / SciPy / stdlib, no torch).

Mirrors the reference's `data/audio.py:18-45` (`audio_read`: 16-bit PCM WAV through the stdlib `wave` module, other
formats through `soundfile` when it is installed) and `data/audio_utils.py:15-41` (`convert_audio`: channel mix-down
by averaging, polyphase resampling with `scipy.signal.resample_poly`)."""

from __future__ import annotations

import math
import wave
from pathlib import Path
from typing import Tuple, Union

import numpy as np


def audio_read(filepath: Union[str, Path]) -> Tuple[np.ndarray, int]:
    """-> (float32 [1, T] in [-1, 1], sample rate)."""
    filepath = Path(filepath)
    if filepath.suffix.lower() == ".wav":
        with wave.open(str(filepath), "rb") as f:
            rate, n_ch = f.getframerate(), f.getnchannels()
            if f.getsampwidth() != 2:
                raise ValueError("only 16-bit PCM WAV files are supported")
            samples = np.frombuffer(f.readframes(-1), dtype=np.int16).astype(np.float32) / 32768.0
        if n_ch > 1:
            samples = samples.reshape(-1, n_ch).mean(axis=1)
        return samples[None, :], rate
    try:
        import soundfile as sf
    except ImportError as e:
        raise ImportError("soundfile is required to read non-WAV audio files") from e
    data, rate = sf.read(str(filepath), dtype="float32")
    wav = data[None, :] if data.ndim == 1 else data.mean(axis=1)[None, :]
    return wav, rate


def convert_audio(wav, from_rate, to_rate, to_channels: int) -> np.ndarray:
    """Channel conversion (mean / tile) and polyphase resampling -> float32 [C, T]."""
    w = np.asarray(wav)
    w = w[None, :] if w.ndim == 1 else w
    if w.shape[0] != to_channels:
        if to_channels == 1:
            w = w.mean(axis=0, keepdims=True)
        elif w.shape[0] == 1:
            w = np.tile(w, (to_channels, 1))
        else:
            raise ValueError(f"Cannot convert from {w.shape[0]} channels to {to_channels} channels")
    fr, tr = int(round(from_rate)), int(round(to_rate))
    if fr != tr:
        from scipy.signal import resample_poly
        g = math.gcd(fr, tr)
        w = resample_poly(w, tr // g, fr // g, axis=-1)
    return w.astype(np.float32)


def write_wav(path: Union[str, Path], audio: np.ndarray, sample_rate: int) -> None:
    """float [-1, 1] -> 16-bit PCM mono WAV (what the reference CLI writes)."""
    pcm = np.clip(np.asarray(audio, dtype=np.float32).reshape(-1), -1.0, 1.0)
    with wave.open(str(path), "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes((pcm * 32767.0).astype(np.int16).tobytes())
