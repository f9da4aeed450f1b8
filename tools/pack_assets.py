#!/usr/bin/env python
"""Regenerate dqnflappybird_b200/assets/sprites.fbpk from a reference checkout.

    python tools/pack_assets.py [/root/reference/assets/sprites]

The blob is derived data (six PNGs -> one packed RGBA table); it is committed
so GPU boxes, which have no /root/reference, can run the tests and the bench.
"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dqnflappybird_b200.assets import pack_sprites, _DEFAULT_BLOB  # noqa: E402


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/assets/sprites"
    blob = pack_sprites(src)
    os.makedirs(os.path.dirname(_DEFAULT_BLOB), exist_ok=True)
    with open(_DEFAULT_BLOB, "wb") as f:
        f.write(blob)
    print(f"wrote {_DEFAULT_BLOB}: {len(blob)} bytes sha256={hashlib.sha256(blob).hexdigest()[:16]}")


if __name__ == "__main__":
    main()
