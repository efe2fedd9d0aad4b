"""Command line: same flags and defaults as the reference CLI (`pocket_tts_mlx/main.py:16-85`)."""

import argparse
import logging
import sys
import wave
from pathlib import Path

import numpy as np

from . import TTSModel

logger = logging.getLogger(__name__)


def write_wav(path: Path, audio: np.ndarray, sample_rate: int) -> None:
    """16-bit PCM mono WAV via the stdlib (what soundfile's default WAV subtype produces for float input)."""
    pcm = (np.clip(np.asarray(audio, dtype=np.float32), -1.0, 1.0) * 32767.0).astype("<i2")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sample_rate)
        w.writeframes(pcm.tobytes())


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description="Generate speech from text using pocket-tts on a B200 GPU")
    p.add_argument("text", help="Text to convert to speech")
    p.add_argument("--voice", "-v", default="marius", help="Voice name (default: marius)")
    p.add_argument("--output", "-o", default="output.wav", help="Output WAV file")
    p.add_argument("--max-tokens", type=int, default=500, help="Max tokens per chunk")
    p.add_argument("--frames-after-eos", type=int, default=7, help="Frames after EOS")
    p.add_argument("--trim-start-ms", type=int, default=0, help="Trim this many milliseconds from start of generated audio")
    p.add_argument("--fade-in-ms", type=int, default=0, help="Apply linear fade-in over this many milliseconds")
    p.add_argument("--warmup-frames", type=int, default=1,
                   help="Number of initial Mimi frames to decode and discard for cleaner onset")
    p.add_argument("--verbose", "-V", action="store_true", help="Verbose logging")
    p.add_argument("--config", default=None, help="(extension) model variant name or .yaml path")
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"], help="(extension) storage precision")
    p.add_argument("--stream", action="store_true",
                   help="(extension) write the WAV while generating (16-bit PCM converted on the GPU, placeholder-length "
                        "header, 0.2 s trailing silence); implied by --output -  (stdout)")
    args = p.parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO, format="%(message)s")
    try:
        logger.info("Loading model...")
        model = TTSModel.load_model(**({"config": args.config} if args.config else {}), precision=args.precision)
        logger.info("Loading voice: %s", args.voice)
        state = model.get_state_for_audio_prompt(args.voice)
        logger.info("Generating audio...")
        if args.stream or args.output == "-":
            # streaming sink of the reference's data/audio.py:108-130; trim / fade need the whole waveform and are skipped
            from .audio import stream_audio_chunks
            if args.output != "-":
                Path(args.output).parent.mkdir(parents=True, exist_ok=True)
            chunks = model.generate_audio_stream(model_state=state, text_to_generate=args.text, max_tokens=args.max_tokens,
                                                 frames_after_eos=args.frames_after_eos, warmup_frames=args.warmup_frames,
                                                 pcm16=True)
            stream_audio_chunks(args.output, chunks, model.config.mimi.sample_rate)
            return 0
        audio = model.generate_audio(model_state=state, text_to_generate=args.text, max_tokens=args.max_tokens,
                                     frames_after_eos=args.frames_after_eos, trim_start_ms=args.trim_start_ms,
                                     fade_in_ms=args.fade_in_ms, warmup_frames=args.warmup_frames)
        out = Path(args.output)
        out.parent.mkdir(parents=True, exist_ok=True)
        write_wav(out, np.asarray(audio), model.config.mimi.sample_rate)
        logger.info("Wrote %s (%.2fs)", out, audio.shape[-1] / model.config.mimi.sample_rate)
        return 0
    except Exception as exc:  # same contract as the reference: any failure -> exit code 1
        logger.error("Error: %s", exc)
        if args.verbose:
            import traceback
            traceback.print_exc()
        return 1


if __name__ == "__main__":
    sys.exit(main())
