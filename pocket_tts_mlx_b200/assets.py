"""Resolve weight / tokenizer / voice locations without network access.

The reference downloads `hf://repo/file@rev` and `http(s)://` assets on demand
(`pocket_tts_mlx/utils/utils.py:56-84`).  This image has no network, so only two branches exist here:
a local path is returned as is (the reference's last branch, `utils.py:84`), and an `hf://` URL is
looked up in `$POCKET_TTS_ASSETS` (flat directory holding the file under its repo-relative name) and
in the local Hugging Face cache.  Anything else raises `FileNotFoundError` with the place to put it.
"""

from __future__ import annotations

import os
from pathlib import Path


def _hf_cache_dirs():
    roots = []
    if os.environ.get("HF_HUB_CACHE"):
        roots.append(Path(os.environ["HF_HUB_CACHE"]))
    if os.environ.get("HF_HOME"):
        roots.append(Path(os.environ["HF_HOME"]) / "hub")
    roots.append(Path.home() / ".cache" / "huggingface" / "hub")
    return roots


def resolve_asset(location: str) -> Path:
    location = str(location)
    if location.startswith(("http://", "https://")):
        raise FileNotFoundError(
            f"{location}: remote download is not available in this build; fetch the file and pass its local path")
    if not location.startswith("hf://"):
        return Path(location)
    rest = location[len("hf://"):]
    parts = rest.split("/")
    repo_id, filename = "/".join(parts[:2]), "/".join(parts[2:])
    revision = None
    if "@" in filename:
        filename, revision = filename.split("@")
    env = os.environ.get("POCKET_TTS_ASSETS")
    if env:
        for cand in (Path(env) / filename, Path(env) / Path(filename).name):
            if cand.exists():
                return cand
    for root in _hf_cache_dirs():
        repo_dir = root / ("models--" + repo_id.replace("/", "--")) / "snapshots"
        if not repo_dir.is_dir():
            continue
        snaps = [repo_dir / revision] if revision else sorted(repo_dir.iterdir())
        for snap in snaps:
            cand = snap / filename
            if cand.exists():
                return cand
    raise FileNotFoundError(
        f"{location} is not available locally (no network here). Put '{filename}' under $POCKET_TTS_ASSETS "
        f"or into the Hugging Face cache for {repo_id}.")
