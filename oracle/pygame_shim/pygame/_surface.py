import numpy as np


class Surface:
    def __init__(self, rgba_xy):
        self.px = np.ascontiguousarray(rgba_xy, dtype=np.uint8)   # [x][y][4]

    def get_width(self):
        return self.px.shape[0]

    def get_height(self):
        return self.px.shape[1]

    def get_at(self, pos):
        x, y = pos
        return tuple(int(v) for v in self.px[x, y])

    def convert(self):
        return self

    def convert_alpha(self):
        return self

    def blit(self, src, dest):
        dx, dy = int(dest[0]), int(dest[1])
        W, H = self.px.shape[:2]
        w, h = src.px.shape[:2]
        x0, y0 = max(dx, 0), max(dy, 0)
        x1, y1 = min(dx + w, W), min(dy + h, H)
        if x1 <= x0 or y1 <= y0:
            return
        s = src.px[x0 - dx:x1 - dx, y0 - dy:y1 - dy]
        d = self.px[x0:x1, y0:y1]
        m = s[..., 3] != 0
        d[m] = s[m]
