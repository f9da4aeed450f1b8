"""Sprite packer for the Flappy Bird hot path.

The reference loads six PNGs by relative path at import time
(game/flappy_bird_utils.py:19-32, :56-73) and derives hitmasks from their
alpha channel (game/flappy_bird_utils.py:103-124).  The device library takes
the same sprites as ONE packed blob (``fb_assets_load``); every derived table
(hitmask bit rows, pipe/bird observation masks) is built inside the library.

Blob layout (little endian), all arrays indexed [x][y][rgba] like a pygame
surface (``get_at((x, y))``):

    0   char[4]  "FBPK"
    4   u32      version (1)
    8   u32[6]   bird_w, bird_h, pipe_w, pipe_h, base_w, base_h
    32  u8       bird  [3][bird_w][bird_h][4]   up, mid, down flap
        u8       pipe  [pipe_w][pipe_h][4]      lower pipe; upper = rot180
        u8       base  [base_w][base_h][4]
"""
from __future__ import annotations

import os
import struct

import numpy as np

MAGIC = b"FBPK"
VERSION = 1

# file names used by the reference, game/flappy_bird_utils.py:19-32
PLAYER_FILES = ("redbird-upflap.png", "redbird-midflap.png", "redbird-downflap.png")
PIPE_FILE = "pipe-green.png"
BASE_FILE = "base.png"
BACKGROUND_FILE = "background-black.png"

_DEFAULT_BLOB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "sprites.fbpk")


def _load_rgba_xy(path: str) -> np.ndarray:
    """PNG -> u8[x][y][4] (pygame surface indexing)."""
    from PIL import Image

    im = np.array(Image.open(path).convert("RGBA"), dtype=np.uint8)  # [y][x][4]
    return np.ascontiguousarray(im.transpose(1, 0, 2))


def pack_sprites(sprites_dir: str) -> bytes:
    """Pack the six sprites the hot path uses from ``<assets>/sprites``."""
    birds = [_load_rgba_xy(os.path.join(sprites_dir, f)) for f in PLAYER_FILES]
    pipe = _load_rgba_xy(os.path.join(sprites_dir, PIPE_FILE))
    base = _load_rgba_xy(os.path.join(sprites_dir, BASE_FILE))
    bg = _load_rgba_xy(os.path.join(sprites_dir, BACKGROUND_FILE))
    if bg.shape[:2] != (288, 512):
        raise ValueError("background must be 288x512 (wrapped_flappy_bird.py:16-17)")
    if bg[..., :3].any():
        raise ValueError("only the black background of the reference is supported")
    if len({b.shape for b in birds}) != 1:
        raise ValueError("player sprites must share one size")
    for a in birds + [pipe, base]:
        al = a[..., 3]
        if not np.isin(al, (0, 255)).all():
            raise ValueError("sprites must have binary alpha (blit == masked copy)")
    bw, bh = birds[0].shape[:2]
    pw, ph = pipe.shape[:2]
    sw, sh = base.shape[:2]
    head = MAGIC + struct.pack("<I6I", VERSION, bw, bh, pw, ph, sw, sh)
    body = b"".join(b.tobytes() for b in birds) + pipe.tobytes() + base.tobytes()
    return head + body


def unpack_sprites(blob: bytes) -> dict:
    """Inverse of pack_sprites: dict of u8[x][y][4] arrays (for tests/oracle)."""
    if blob[:4] != MAGIC:
        raise ValueError("not an FBPK blob")
    ver, bw, bh, pw, ph, sw, sh = struct.unpack_from("<I6I", blob, 4)
    if ver != VERSION:
        raise ValueError(f"unsupported FBPK version {ver}")
    off = 32
    n = 3 * bw * bh * 4
    bird = np.frombuffer(blob, np.uint8, n, off).reshape(3, bw, bh, 4); off += n
    n = pw * ph * 4
    pipe = np.frombuffer(blob, np.uint8, n, off).reshape(pw, ph, 4); off += n
    n = sw * sh * 4
    base = np.frombuffer(blob, np.uint8, n, off).reshape(sw, sh, 4); off += n
    if off != len(blob):
        raise ValueError("FBPK blob has trailing or missing bytes")
    return {"bird": bird, "pipe": pipe, "base": base}


def load_blob(assets_dir: str | None = None) -> bytes:
    """Blob for ``assets_dir`` (a reference-style ``assets`` directory), or the
    packed default shipped with the package when ``assets_dir`` is None."""
    if assets_dir is not None:
        d = assets_dir
        if os.path.isdir(os.path.join(d, "sprites")):
            d = os.path.join(d, "sprites")
        return pack_sprites(d)
    with open(_DEFAULT_BLOB, "rb") as f:
        return f.read()
