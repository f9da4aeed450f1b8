import numpy as np
from ._surface import Surface

_screen = None


def set_mode(size, flags=0, depth=0):
    global _screen
    px = np.zeros((size[0], size[1], 4), np.uint8)
    px[..., 3] = 255
    _screen = Surface(px)
    return _screen


def set_caption(*a, **k):
    pass


def get_surface():
    return _screen


def update(*a, **k):
    pass
