"""ORACLE — test infrastructure only; never imported by the product path.

CPU restatement (numpy, the reference's own expressions) of the image formats either side of the path:
  * preprocess   /root/reference/scripts/inference.py:111-116  (after cv2.resize: astype(float32) / 127.5 - 1.0, HWC -> CHW)
  * postprocess  /root/reference/scripts/inference.py:121-127  (before cv2.resize: CHW -> HWC, (y + 1.0) * 127.5, clip, uint8)
Batched (the reference handles one image; the arithmetic is per element).

Parity status: PINNED — tests/golden/image_io_kat.npz was produced by calling the unmodified reference functions
(tests/golden/make_golden_image_io.py) and tests/test_oracle.py replays it against this file.
"""
import numpy as np


def preprocess_u8(images_hwc: np.ndarray) -> np.ndarray:
    """uint8 [N,H,W,3] -> float32 [N,3,H,W]."""
    x = images_hwc.astype(np.float32) / 127.5 - 1.0
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


def postprocess_u8(images_nchw: np.ndarray) -> np.ndarray:
    """float32 [N,3,H,W] -> uint8 [N,H,W,3]."""
    y = images_nchw.transpose(0, 2, 3, 1)
    y = (y + 1.0) * 127.5
    return np.ascontiguousarray(np.clip(y, 0, 255).astype(np.uint8))


def _resize_taps(dst: int, src: int, is_x: bool):
    """OpenCV resize.cpp, INTER_LINEAR tables: fx = (float)((d + 0.5) * scale - 0.5), s = floor(fx), fx -= s;
    x taps outside the image are folded into the border pixel (weight reset), y taps keep their weights (rows are
    clamped later); coefficients = saturate_cast<short>(c * 2048) (round half to even)."""
    scale = 1.0 / (np.float64(dst) / np.float64(src))
    ofs = np.zeros(dst, np.int64)
    w = np.zeros((dst, 2), np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if is_x:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= src - 1:
                f, s = np.float32(0), src - 1
        ofs[d] = s
        w[d, 0] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(2048))))
        w[d, 1] = int(np.rint(np.float32(f * np.float32(2048))))
    return ofs, w


def resize_bilinear_u8(images_hwc: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h)) for uint8 [N,H,W,3] (scripts/inference.py:109,130; OpenCV 4.x HResizeLinear /
    VResizeLinear fixed-point path; identical to the area-averaging path OpenCV takes for an exact 2x shrink)."""
    n, sh, sw, _ = images_hwc.shape
    xo, xa = _resize_taps(dst_w, sw, True)
    yo, yb = _resize_taps(dst_h, sh, False)
    S = images_hwc.astype(np.int64)
    x1 = np.minimum(xo + 1, sw - 1)
    H = S[:, :, xo, :] * xa[:, 0][None, None, :, None] + S[:, :, x1, :] * xa[:, 1][None, None, :, None]
    y0, y1 = np.clip(yo, 0, sh - 1), np.clip(yo + 1, 0, sh - 1)
    b0, b1 = yb[:, 0][None, :, None, None], yb[:, 1][None, :, None, None]
    out = (((b0 * (H[:, y0] >> 4)) >> 16) + ((b1 * (H[:, y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
