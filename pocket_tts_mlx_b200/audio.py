"""Host-side audio helpers of the voice-cloning branch (NumPy / SciPy / stdlib, no torch).

Mirrors the reference's `data/audio.py:18-45` (`audio_read`: 16-bit PCM WAV through the stdlib `wave` module, other
formats through `soundfile` when it is installed) and `data/audio_utils.py:15-41` (`convert_audio`: channel mix-down
by averaging, polyphase resampling with `scipy.signal.resample_poly`)."""

from __future__ import annotations

import math
import os
import sys
import wave
from contextlib import nullcontext
from pathlib import Path
from typing import Any, Iterable, Optional, Tuple, Union

import numpy as np

# seconds of audio held back before the first bytes are written (same environment knob as the reference)
FIRST_CHUNK_LENGTH_SECONDS = float(os.environ.get("FIRST_CHUNK_LENGTH_SECONDS", "0"))


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


# ---- streaming output (reference data/audio.py:48-130) ---------------------------------------------------
def to_pcm16(chunk: Any) -> np.ndarray:
    """Float samples -> int16 the way the reference does it (`clip(x, -1, 1) * 32767` truncated, data/audio.py:70).
    int16 input passes through untouched: with `pcm16=True` the GPU already produced exactly these values in the
    kernels that make the final samples, so no host conversion is left."""
    a = np.asarray(chunk)
    if a.dtype == np.int16:
        return a.reshape(-1)
    return (np.clip(a.reshape(-1), -1, 1) * 32767).astype(np.int16)


class StreamingWAVWriter:
    """WAV stream whose length is not known when the header goes out (data/audio.py:48-100): the header announces
    10^9 frames, PCM is appended with `writeframesraw`, the first bytes can be held back until
    FIRST_CHUNK_LENGTH_SECONDS of audio exist, `finalize` appends 0.2 s of silence and closes without patching the
    header (the sink may be a pipe)."""

    PLACEHOLDER_FRAMES = 1_000_000_000
    TRAILING_SILENCE_SECONDS = 0.2

    def __init__(self, output_stream, sample_rate: int):
        self.output_stream = output_stream
        self.sample_rate = int(sample_rate)
        self.wave_writer: Optional[wave.Wave_write] = None
        self.first_chunk_buffer: Optional[list] = []

    def write_header(self, sample_rate: Optional[int] = None):
        if sample_rate is not None:
            self.sample_rate = int(sample_rate)
        w = wave.open(self.output_stream, "wb")
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(self.sample_rate)
        w.setnframes(self.PLACEHOLDER_FRAMES)
        self.wave_writer = w

    def write_pcm_data(self, audio_chunk: Any):
        data = to_pcm16(audio_chunk).astype("<i2", copy=False).tobytes()
        if self.first_chunk_buffer is not None:
            self.first_chunk_buffer.append(data)
            held = sum(len(c) for c in self.first_chunk_buffer)
            if held < int(self.sample_rate * FIRST_CHUNK_LENGTH_SECONDS) * 2:
                return
            self._flush()
            return
        self.wave_writer.writeframesraw(data)

    def _flush(self):
        if self.first_chunk_buffer is not None:
            self.wave_writer.writeframesraw(b"".join(self.first_chunk_buffer))
            self.first_chunk_buffer = None

    def finalize(self):
        self._flush()
        self.wave_writer.writeframesraw(bytes(int(self.sample_rate * self.TRAILING_SILENCE_SECONDS) * 2))
        self.wave_writer._patchheader = lambda: None      # keep the placeholder length: the sink may not be seekable
        self.wave_writer.close()


def is_file_like(obj) -> bool:
    return all(hasattr(obj, a) for a in ("write", "close"))


def stream_audio_chunks(path, audio_chunks: Iterable[Any], sample_rate: int) -> None:
    """Drain a chunk generator into a WAV stream: a path, an open binary handle, "-" for stdout, or None to only run
    the generator (data/audio.py:108-130)."""
    if path == "-":
        f = sys.stdout.buffer
    elif path is None:
        f = nullcontext()
    elif is_file_like(path):
        f = path
    else:
        f = open(path, "wb")
    with f:
        writer = None
        if path is not None:
            writer = StreamingWAVWriter(f, sample_rate)
            writer.write_header(sample_rate)
        for chunk in audio_chunks:
            if writer is not None:
                writer.write_pcm_data(chunk)
        if writer is not None:
            writer.finalize()
