"""Seeded synthetic inputs for parity tests and benches.

page(seed): H x W x 3 u8 "scanned page": light noisy background with ~48 dark
text-line blobs carrying a vertical-stroke texture; one in four lines is exactly
axis-aligned with integer corners (exercises the reference's crop fast path,
oar-ocr-core/src/utils/transform.rs:150-152), the rest are skewed by up to 4
degrees (homography + bicubic path).  crop(seed): 48 x 320 x 3 stroke texture
for the recognizer-only configuration.
"""
from __future__ import annotations

import numpy as np


def _stroke_row(rng, n, lo=20, hi=90):
    """1-D glyph-like pattern of length n: alternating dark / less-dark strokes."""
    out = np.empty(n, np.uint8)
    x = 0
    dark = True
    while x < n:
        wdt = int(rng.integers(2, 6))
        v = int(rng.integers(lo, lo + 25)) if dark else int(rng.integers(hi - 30, hi + 1))
        out[x:x + wdt] = v
        x += wdt
        dark = not dark
    return out


def page(seed: int, size: int = 960, rows_pitch: int = 50, margin: int = 40) -> np.ndarray:
    rng = np.random.default_rng(1000 + seed)
    img = rng.integers(225, 246, size=(size, size, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    y = margin
    while y + 36 <= size - margin:
        x = margin + int(rng.integers(0, 30))
        while True:
            wdt = int(rng.integers(100, 381))
            if x + wdt > size - margin:
                break
            hgt = int(rng.integers(22, 35))
            aligned = rng.random() < 0.25
            cx, cy = x + wdt / 2.0, y + 17.0
            if aligned:
                ang = 0.0
                cx, cy = float(int(cx)) + (0.5 if wdt % 2 else 0.0), float(int(cy)) + (0.5 if hgt % 2 else 0.0)
            else:
                amax = min(np.deg2rad(4.0), np.arcsin(min(1.0, 7.0 / wdt)))
                ang = float(rng.uniform(-amax, amax))
            ca, sa = np.cos(ang), np.sin(ang)
            x0, x1 = max(0, int(cx - wdt / 2 - 12)), min(size, int(cx + wdt / 2 + 12))
            y0, y1 = max(0, int(cy - hgt / 2 - 12)), min(size, int(cy + hgt / 2 + 12))
            dx = xx[y0:y1, x0:x1] + 0.5 - cx
            dy = yy[y0:y1, x0:x1] + 0.5 - cy
            u = dx * ca + dy * sa
            v = -dx * sa + dy * ca
            inside = (np.abs(u) <= wdt / 2.0) & (np.abs(v) <= hgt / 2.0)
            pat = _stroke_row(rng, wdt + 2)
            ui = np.clip((u + wdt / 2.0).astype(np.int32), 0, wdt + 1)
            vals = pat[ui]
            band = (np.abs(v) > hgt * 0.38)
            vals = np.where(band, np.minimum(vals + 20, 110), vals).astype(np.uint8)
            sub = img[y0:y1, x0:x1]
            for c in range(3):
                sub[..., c] = np.where(inside, vals, sub[..., c])
            x += wdt + 24 + int(rng.integers(0, 20))
        y += rows_pitch
    return img


def crop(seed: int, h: int = 48, w: int = 320) -> np.ndarray:
    rng = np.random.default_rng(2000 + seed)
    img = rng.integers(225, 246, size=(h, w, 3), dtype=np.uint8)
    pat = _stroke_row(rng, w)
    top, bot = int(rng.integers(4, 10)), h - int(rng.integers(4, 10))
    for c in range(3):
        img[top:bot, :, c] = pat[None, :]
    return img
