import numpy as np


def array3d(surface):
    return np.array(surface.px[..., :3], dtype=np.uint8, copy=True)
