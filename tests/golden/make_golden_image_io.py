#!/usr/bin/env python
"""Golden vectors for the image formats either side of the path, FROM THE UNMODIFIED REFERENCE
(/root/reference/scripts/inference.py: preprocess_image, postprocess_image).  Build container only.

The reference functions resize with cv2; the fixtures use target size == image size, where cv2.resize is the identity,
so that they pin exactly the arithmetic this repo implements (normalisation, layout, clip, truncation).

    python tests/golden/make_golden_image_io.py
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "ref_shim"), "/root/reference"]

spec = importlib.util.spec_from_file_location("ref_inference", "/root/reference/scripts/inference.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

from oracle import image_io_oracle  # noqa: E402

import cv2  # noqa: E402

rng = np.random.default_rng(7)
S = 32
rgb = rng.integers(0, 256, size=(2, S, S, 3), dtype=np.uint8)
rgb[0, 0, :8, 0] = [0, 1, 2, 127, 128, 254, 255, 63]     # the interesting byte values
pre = []
with tempfile.TemporaryDirectory() as d:
    for i in range(rgb.shape[0]):
        path = os.path.join(d, f"im{i}.png")
        cv2.imwrite(path, cv2.cvtColor(rgb[i], cv2.COLOR_RGB2BGR))       # lossless; the reference converts BGR -> RGB back
        x, size = ref.preprocess_image(path, S)
        assert size == (S, S) and x.shape == (1, 3, S, S) and x.dtype == np.float32
        pre.append(x[0])
pre = np.stack(pre)
assert np.array_equal(image_io_oracle.preprocess_u8(rgb), pre), "oracle preprocess differs from the reference"

# model outputs: inside and outside [-1, 1], exact .5 boundaries of the uint8 grid, tiny negatives
y = rng.uniform(-1.3, 1.3, size=(2, 3, S, S)).astype(np.float32)
y[0, 0, 0, :6] = [-1.0, 1.0, 0.0, -1.0000001, 0.99999994, 2.0 / 255 - 1]
post = np.stack([ref.postprocess_image(y[i:i + 1], (S, S)) for i in range(y.shape[0])])
assert post.dtype == np.uint8 and post.shape == (2, S, S, 3)
assert np.array_equal(image_io_oracle.postprocess_u8(y), post), "oracle postprocess differs from the reference"

# cv2.resize exactly as the reference calls it (inference.py:109 and :130): default interpolation, dsize = (width, height)
resize = {}
for k, (sh, sw, dh, dw) in enumerate([(24, 32, 64, 64), (50, 70, 32, 32), (64, 64, 32, 32), (33, 17, 40, 56), (7, 5, 64, 48),
                                      (48, 48, 48, 48), (96, 64, 32, 32)]):
    src = rng.integers(0, 256, size=(sh, sw, 3), dtype=np.uint8)
    dst = cv2.resize(src, (dw, dh))
    assert np.array_equal(image_io_oracle.resize_bilinear_u8(src[None], dh, dw)[0], dst), ("oracle resize differs from cv2", sh, sw, dh, dw)
    resize[f"rs{k}_src"], resize[f"rs{k}_dst"] = src, dst
# a photo-sized check that is not stored
big = rng.integers(0, 256, size=(1080, 1920, 3), dtype=np.uint8)
assert np.array_equal(image_io_oracle.resize_bilinear_u8(big[None], 256, 256)[0], cv2.resize(big, (256, 256)))
small = rng.integers(0, 256, size=(256, 256, 3), dtype=np.uint8)
assert np.array_equal(image_io_oracle.resize_bilinear_u8(small[None], 1080, 1920)[0], cv2.resize(small, (1920, 1080)))

np.savez_compressed(os.path.join(HERE, "image_io_kat.npz"), rgb=rgb, pre=pre, y=y, post=post, cv2_version=cv2.__version__, **resize)
print("image_io_kat.npz written:", rgb.shape, pre.shape, y.shape, post.shape)
