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
