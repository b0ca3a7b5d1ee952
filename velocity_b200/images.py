"""boundingRect / insidebbox with the semantics of utils/images.py:9-27 (no cv2 needed:
cv2.boundingRect of float32 points is x0 = floor(min x), w = floor(max x) - x0 + 1)."""
import numpy as np


def boundingRect(x, imshape, border=(0, 0)):
    x = np.asarray(x)
    fx, fy = np.floor(x[:, 0]), np.floor(x[:, 1])
    x0, y0 = int(fx.min()), int(fy.min())
    width, height = int(fx.max()) - x0 + 1, int(fy.max()) - y0 + 1
    xa, ya = x0 - border[0], y0 - border[1]
    xb, yb = x0 + width + border[0], y0 + height + border[1]
    return max(xa, 1), min(xb, imshape[1]), max(ya, 1), min(yb, imshape[0])


def insidebbox(x, box):
    x0, x1, y0, y1 = box
    return (x[:, 0] > x0) & (x[:, 0] < x1) & (x[:, 1] > y0) & (x[:, 1] < y1)
