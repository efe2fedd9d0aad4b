"""Host-side text handling: SentencePiece tokenisation, prompt normalisation, sentence packing.

Behavioural mirror of the reference's host logic (`pocket_tts_mlx/conditioners/text.py:15-29`,
`pocket_tts_mlx/models/tts_model.py:521-593`) including its quirks: the
8-space short-text pad is applied by `prepare_text_prompt` but stripped again before tokenising, and
the per-chunk prepared text is discarded -- the ids that reach the model are `encode(chunk)`.
"""

from __future__ import annotations

from typing import List, Tuple

import numpy as np

_SENTENCE_ENDERS = ".!...?"


class SentencePieceTokenizer:
    """Thin wrapper; asserts the vocabulary size like the reference (`text.py:21`)."""

    def __init__(self, n_bins: int, model_path):
        import sentencepiece

        self.sp = sentencepiece.SentencePieceProcessor(str(model_path))
        if self.sp.vocab_size() != n_bins:
            raise AssertionError(
                f"sentencepiece tokenizer has vocab_size={self.sp.vocab_size()} but n_bins={n_bins} was specified")

    def encode(self, text: str) -> np.ndarray:
        return np.asarray(self.sp.encode(text, out_type=int), dtype=np.int32)

    def decode(self, ids) -> str:
        return self.sp.decode([int(i) for i in ids])

    __call__ = encode


def prepare_text_prompt(text: str) -> Tuple[str, int]:
    """Normalise a prompt and guess how many frames to keep after EOS (3 if <=4 words else 1)."""
    text = text.strip()
    if not text:
        raise ValueError("Text prompt cannot be empty")
    text = text.replace("\n", " ").replace("\r", " ").replace("  ", " ")
    guess = 3 if len(text.split()) <= 4 else 1
    if not text[0].isupper():
        text = text[0].upper() + text[1:]
    if text[-1].isalnum():
        text += "."
    if len(text.split()) < 5:
        text = " " * 8 + text
    return text, guess


def split_into_best_sentences(tokenizer: SentencePieceTokenizer, text: str, max_tokens: int) -> List[str]:
    """Cut at sentence-final tokens, then greedily pack sentences into chunks of <= max_tokens."""
    text, _ = prepare_text_prompt(text)
    ids = tokenizer.encode(text.strip()).tolist()
    enders = set(tokenizer.encode(_SENTENCE_ENDERS).tolist()[1:])

    starts = [0]
    in_ender_run = False
    for pos, tok in enumerate(ids):
        if tok in enders:
            in_ender_run = True
        else:
            if in_ender_run:
                starts.append(pos)
            in_ender_run = False
    starts.append(len(ids))

    sentences = [(hi - lo, tokenizer.decode(ids[lo:hi])) for lo, hi in zip(starts[:-1], starts[1:])]

    chunks: List[str] = []
    cur, cur_n = "", 0
    for n, sent in sentences:
        if cur == "":
            cur, cur_n = sent, n
        elif cur_n + n > max_tokens:
            chunks.append(cur.strip())
            cur, cur_n = sent, n
        else:
            cur, cur_n = cur + " " + sent, cur_n + n
    if cur != "":
        chunks.append(cur.strip())
    return chunks
