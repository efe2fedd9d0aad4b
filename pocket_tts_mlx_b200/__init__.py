"""pocket-tts streaming generation on NVIDIA B200 (sm_100a); same public surface as pocket_tts_mlx."""

from .tts_model import TTSModel

__all__ = ["TTSModel"]
__version__ = "0.1.0"
