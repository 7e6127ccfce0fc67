"""Seeded synthetic CIF I420 clips (the reference's data/*.yuv are absent from the mount).

One generator per BASELINE.json config family:

* ``akiyo``      static smooth background + fixed texture + a drifting head-and-shoulders ellipse + N(0,1) noise
* ``intra``      the same family with a stronger texture (config 1, all-intra)
* ``highmotion`` panned textured field + 4 independently moving 48x48 objects + N(0,2) noise (config 2/3)
* ``flat``       constant regions + one moving square: forces zero-SAD early breaks and the carried
                 spiral-search state of the reference's motionEstimation (ENC:2094-2148)

All clips keep well below 8 bit/pixel so the reference's unchecked W*H*nframes bitstream buffer
(ENC:4874-4875) does not overflow.  Frames are planar I420: Y (w*h) then Cb, Cr (w/2*h/2 each).
"""
from __future__ import annotations

import numpy as np

CIF_W, CIF_H = 352, 288


def frame_bytes(w: int = CIF_W, h: int = CIF_H) -> int:
    return w * h * 3 // 2


def _pack(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    n = y.shape[0]
    return np.concatenate([y.reshape(n, -1), cb.reshape(n, -1), cr.reshape(n, -1)], axis=1).astype(np.uint8)


def _smooth_chroma(n: int, w: int, h: int, phase: float) -> tuple[np.ndarray, np.ndarray]:
    yy, xx = np.mgrid[0:h // 2, 0:w // 2].astype(np.float64)
    cb = 128 + 30 * np.sin(xx / 23.0 + phase) + 10 * np.cos(yy / 17.0)
    cr = 128 + 25 * np.cos(xx / 19.0) + 15 * np.sin(yy / 13.0 + phase)
    return (np.broadcast_to(cb, (n, h // 2, w // 2)).copy(), np.broadcast_to(cr, (n, h // 2, w // 2)).copy())


def akiyo(nframes: int, seed: int = 20261017, w: int = CIF_W, h: int = CIF_H, texture: int = 24) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    bg = 128 + 50 * np.sin(xx / 37.0) + 40 * np.cos(yy / 29.0) + rng.integers(0, texture, size=(h, w))
    y = np.empty((nframes, h, w), np.float64)
    cb, cr = _smooth_chroma(nframes, w, h, 0.3)
    for n in range(nframes):
        cx = w / 2 + 20 * np.sin(n / 15.0)
        cy = h / 2 + 10 * np.cos(n / 21.0)
        head = ((xx - cx) / 60.0) ** 2 + ((yy - cy) / 80.0) ** 2 <= 1.0
        f = bg.copy()
        f[head] = 90 + 0.25 * (xx[head] - cx) + 0.15 * (yy[head] - cy) + 8 * np.sin((xx[head] + yy[head]) / 5.0)
        f += rng.normal(0.0, 1.0, size=(h, w))
        y[n] = f
        hc = head[::2, ::2]
        cb[n][hc] += 12
        cr[n][hc] -= 9
    return _pack(np.clip(np.rint(y), 0, 255), np.clip(np.rint(cb), 0, 255), np.clip(np.rint(cr), 0, 255))


def intra(nframes: int, seed: int = 1234, w: int = CIF_W, h: int = CIF_H) -> np.ndarray:
    return akiyo(nframes, seed=seed, w=w, h=h, texture=40)


def highmotion(nframes: int, seed: int = 4242, w: int = CIF_W, h: int = CIF_H) -> np.ndarray:
    rng = np.random.default_rng(seed)
    big = rng.integers(0, 64, size=(h + 64, w + 64)).astype(np.float64)
    # low-pass the field a little so that it is textured but compressible
    big = (big + np.roll(big, 1, 0) + np.roll(big, 1, 1) + np.roll(big, (1, 1), (0, 1))) / 4.0
    yy, xx = np.mgrid[0:h + 64, 0:w + 64].astype(np.float64)
    big += 96 + 40 * np.sin(xx / 31.0) + 30 * np.cos(yy / 27.0)
    objs = [dict(x=float(rng.integers(0, w - 48)), y=float(rng.integers(0, h - 48)),
                 vx=float(rng.integers(3, 10)) * rng.choice([-1, 1]), vy=float(rng.integers(3, 10)) * rng.choice([-1, 1]),
                 tex=rng.integers(0, 256, size=(48, 48)).astype(np.float64) * 0.25 + 40 * k) for k in range(4)]
    y = np.empty((nframes, h, w), np.float64)
    cb, cr = _smooth_chroma(nframes, w, h, 1.1)
    for n in range(nframes):
        ox, oy = (2 * n) % 32, n % 32
        f = big[oy:oy + h, ox:ox + w].copy()
        for o in objs:
            x0 = int(o["x"]) % (w - 48)
            y0 = int(o["y"]) % (h - 48)
            f[y0:y0 + 48, x0:x0 + 48] = o["tex"]
            cb[n, y0 // 2:y0 // 2 + 24, x0 // 2:x0 // 2 + 24] = 100 + 10 * (objs.index(o))
            cr[n, y0 // 2:y0 // 2 + 24, x0 // 2:x0 // 2 + 24] = 160 - 12 * (objs.index(o))
            o["x"] += o["vx"]
            o["y"] += o["vy"]
        f += rng.normal(0.0, 2.0, size=(h, w))
        y[n] = f
    return _pack(np.clip(np.rint(y), 0, 255), np.clip(np.rint(cb), 0, 255), np.clip(np.rint(cr), 0, 255))


def flat(nframes: int, seed: int = 7, w: int = CIF_W, h: int = CIF_H) -> np.ndarray:
    """Constant regions + one moving square, no noise: many exact (zero-SAD) matches."""
    rng = np.random.default_rng(seed)
    y = np.empty((nframes, h, w), np.uint8)
    base = np.full((h, w), 60, np.uint8)
    base[:, w // 3:] = 120
    base[h // 2:, : w // 2] = 200
    base[h // 4: h // 4 + 40, w // 2: w // 2 + 100] = rng.integers(0, 256, size=(40, 100))
    sq = rng.integers(0, 256, size=(32, 32)).astype(np.uint8)
    cb = np.full((nframes, h // 2, w // 2), 110, np.uint8)
    cr = np.full((nframes, h // 2, w // 2), 140, np.uint8)
    for n in range(nframes):
        f = base.copy()
        x0 = (5 * n) % (w - 32)
        y0 = (3 * n) % (h - 32)
        f[y0:y0 + 32, x0:x0 + 32] = sq
        y[n] = f
        cb[n, y0 // 2:y0 // 2 + 16, x0 // 2:x0 // 2 + 16] = 90
        cr[n, y0 // 2:y0 // 2 + 16, x0 // 2:x0 // 2 + 16] = 170
    return _pack(y, cb, cr)


GENERATORS = {"akiyo": akiyo, "intra": intra, "highmotion": highmotion, "flat": flat}


def make_clip(kind: str, nframes: int, seed: int | None = None, w: int = CIF_W, h: int = CIF_H) -> np.ndarray:
    """uint8 array [nframes][w*h*3/2] of planar I420 frames."""
    gen = GENERATORS[kind]
    return gen(nframes, w=w, h=h) if seed is None else gen(nframes, seed=seed, w=w, h=h)
