"""Minimal pygame stand-in -- TEST INFRASTRUCTURE ONLY.

Just enough of the pygame API for the reference's game/wrapped_flappy_bird.py
and game/flappy_bird_utils.py to be imported and executed VERBATIM (pygame and
SDL are not installable in this image).  Used by tests/golden/make_golden.py
in the build container (where /root/reference exists) to produce the golden
trajectories the oracle and the CUDA path are pinned against.

Restated primitive semantics (the only part that is not reference code):
  * Surface = u8[x][y][4] RGBA; blit copies pixels whose source alpha != 0
    (every sprite of the reference has binary alpha), clipped to the target,
    destination coordinates truncated toward zero like SDL_Rect ([assumed]);
  * transform.rotate(s, 180) = flip both axes;
  * Rect = int-truncated (x, y, w, h) with pygame's clip();
  * surfarray.array3d = u8[x][y][3] copy.
"""
import numpy as np

from . import display, event, image, surfarray, time, transform, locals  # noqa: F401
from ._surface import Surface  # noqa: F401


def init():
    return (6, 0)


class Rect:
    def __init__(self, x, y, w, h):
        self.x, self.y, self.width, self.height = int(x), int(y), int(w), int(h)

    def clip(self, o):
        x0, y0 = max(self.x, o.x), max(self.y, o.y)
        x1 = min(self.x + self.width, o.x + o.width)
        y1 = min(self.y + self.height, o.y + o.height)
        if x1 > x0 and y1 > y0:
            return Rect(x0, y0, x1 - x0, y1 - y0)
        return Rect(self.x, self.y, 0, 0)
