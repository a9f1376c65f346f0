"""Procedural sprite models in the on-disk SPR format Mod_LoadSpriteModel reads
(ref_soft/r_model.c:1671-1741, structs in ref_soft/spritegn.h:70-110): a header, then per frame a
type word followed by either one dspriteframe_t + pixels or a group (count, intervals, frames).
Texel 255 is transparent (d_sprite.c:166)."""
from __future__ import annotations

import struct

import numpy as np

SPR_VP_PARALLEL_UPRIGHT, SPR_FACING_UPRIGHT, SPR_VP_PARALLEL, SPR_ORIENTED, SPR_VP_PARALLEL_ORIENTED = range(5)


def _frame_pixels(w: int, h: int, seed: int) -> bytes:
    yy, xx = np.mgrid[0:h, 0:w]
    cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
    r = np.sqrt(((xx - cx) / (w / 2.0)) ** 2 + ((yy - cy) / (h / 2.0)) ** 2)
    pix = (32 + ((xx * 7 + yy * 13 + seed * 29) % 160)).astype(np.uint8)
    pix[(xx + yy + seed) % 11 == 0] = 251            # a few full-bright texels
    pix[r > 1.0] = 255                                # transparent outside the disc
    pix[(r < 0.25) & ((xx + seed) % 3 == 0)] = 255    # and some holes inside
    return pix.tobytes()


def _frame(w: int, h: int, seed: int, origin=None) -> bytes:
    ox, oy = origin if origin is not None else (-(w // 2), h // 2)
    return struct.pack("<iiii", ox, oy, w, h) + _frame_pixels(w, h, seed)


def sprite(sprtype: int = SPR_VP_PARALLEL, width: int = 32, height: int = 32, nframes: int = 2, group: bool = False,
           beamlength: float = 0.0, synctype: int = 0, seed: int = 1) -> bytes:
    """nframes single frames; with group=True frame 0 is a three-picture group (intervals 0.1, 0.2, 0.35)."""
    frames = b""
    for i in range(nframes):
        if group and i == 0:
            frames += struct.pack("<i", 1) + struct.pack("<i", 3) + struct.pack("<fff", 0.1, 0.2, 0.35)
            for j in range(3):
                frames += _frame(width - 4 * j, height - 2 * j, seed + 10 + j)
        else:
            frames += struct.pack("<i", 0) + _frame(width, height, seed + i)
    radius = float(np.sqrt((width / 2.0) ** 2 + (height / 2.0) ** 2))
    hdr = b"IDSP" + struct.pack("<iifiiifi", 1, sprtype, radius, width, height, nframes, beamlength, synctype)
    return hdr + frames
