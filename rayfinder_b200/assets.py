"""Locating the baked scenes.  `assets/<Name>.pt` is produced by `__graft_entry__.build()` where the
reference's GLB assets are mounted; `assets/<Name>.pt.xz` is the same file xz-compressed (what travels to
the GPU box — a third of the size)."""
from __future__ import annotations

import lzma
from pathlib import Path

from .api import PtFormat

ASSETS = Path(__file__).resolve().parent.parent / "assets"


def scene_path(name: str) -> Path | None:
    for candidate in (ASSETS / f"{name}.pt", ASSETS / f"{name}.pt.xz"):
        if candidate.exists():
            return candidate
    return None


def load_scene(name: str) -> PtFormat:
    path = scene_path(name)
    if path is None:
        raise FileNotFoundError(f"Failed to open file: {ASSETS / (name + '.pt')}")
    if path.suffix == ".xz":
        return PtFormat.loads(lzma.decompress(path.read_bytes()))
    return PtFormat.load(path)


def compress_scene(name: str, preset: int = 2) -> Path:
    src, dst = ASSETS / f"{name}.pt", ASSETS / f"{name}.pt.xz"
    if not dst.exists() or dst.stat().st_mtime < src.stat().st_mtime:
        dst.write_bytes(lzma.compress(src.read_bytes(), preset=preset))
    return dst
