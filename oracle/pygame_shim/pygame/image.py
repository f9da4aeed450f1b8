import numpy as np
from ._surface import Surface


def load(path):
    from PIL import Image
    im = np.array(Image.open(path).convert("RGBA"), dtype=np.uint8)   # [y][x][4]
    return Surface(im.transpose(1, 0, 2))
