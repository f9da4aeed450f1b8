from ._surface import Surface


def rotate(surface, angle):
    if angle % 360 != 180:
        raise NotImplementedError("shim only rotates by 180 degrees")
    return Surface(surface.px[::-1, ::-1])
