"""Procedural palette / colormap / miptex generators in the reference's on-disk formats.

No game data is available, so every workload is generated.
Formats follow /root/reference:
  gfx/palette.lmp   768 B RGB                                  common/host.c:827
  gfx/colormap.lmp  64 rows x 256 B + fullbright count byte     common/host.c:830, win32/vid_win.c:2151
  miptex_t          name[16], w, h, offsets[4], w*h*85/64 B     common/bspfile.h:95-100
"""
from __future__ import annotations

import struct
import numpy as np

SEED = 0x51AB
NUM_FULLBRIGHT = 32


def make_palette() -> np.ndarray:
    """16 ramps x 16 shades; the last 32 entries are the 'fullbright' colours."""
    pal = np.zeros((256, 3), dtype=np.uint8)
    hues = [
        (1.0, 1.0, 1.0), (0.9, 0.7, 0.5), (0.6, 0.6, 0.9), (0.5, 0.8, 0.4),
        (0.9, 0.3, 0.2), (0.8, 0.8, 0.3), (0.9, 0.6, 0.6), (0.7, 0.5, 0.3),
        (0.7, 0.4, 0.8), (0.4, 0.8, 0.8), (0.9, 0.8, 0.6), (0.3, 0.5, 0.9),
        (0.6, 0.9, 0.6), (0.5, 0.5, 0.5), (1.0, 0.9, 0.3), (1.0, 0.5, 0.1),
    ]
    for r in range(16):
        for s in range(16):
            k = (s + 1) / 16.0
            pal[r * 16 + s] = [int(255 * hues[r][c] * k) for c in range(3)]
    pal[0] = (0, 0, 0)
    return pal


def _nearest(pal: np.ndarray, rgb: np.ndarray) -> np.ndarray:
    """Index of the palette entry closest (squared RGB distance) to each rgb row."""
    d = ((pal[None, :, :].astype(np.int32) - rgb[:, None, :].astype(np.int32)) ** 2).sum(axis=2)
    return d.argmin(axis=1).astype(np.uint8)


def make_colormap(pal: np.ndarray, fullbrights: int = NUM_FULLBRIGHT) -> bytes:
    """64 light rows x 256 colours, row 0 brightest ... row 63 darkest, as sampled by
    `colormap[(light & 0xFF00) + texel]` (ref_soft/r_surf.c:217).  Shaded entries map
    to the nearest palette colour of rgb*(63-row)/32 (over-bright clamps), the last
    `fullbrights` entries map to themselves.  One trailing byte = fullbright count
    (win32/vid_win.c:2151 reads `host_colormap[16384]`)."""
    nshade = 256 - fullbrights
    out = np.zeros((64, 256), dtype=np.uint8)
    base = pal[:nshade].astype(np.int32)
    for row in range(64):
        lvl = 63 - row
        rgb = np.minimum((base * lvl) >> 5, 255)
        out[row, :nshade] = _nearest(pal, rgb)
        out[row, nshade:] = np.arange(nshade, 256, dtype=np.uint8)
    return out.tobytes() + bytes([fullbrights])


def _lcg_noise(seed: int, w: int, h: int) -> np.ndarray:
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    return rng.randint(0, 1 << 16, size=(h, w)).astype(np.int32)


def make_texture_pixels(seed: int, w: int, h: int, style: str = "brick") -> np.ndarray:
    """w x h palette indices in 0..223 with visible structure so renders can be eyeballed."""
    n = _lcg_noise(seed, w, h)
    yy, xx = np.mgrid[0:h, 0:w]
    ramp = (seed * 7) % 13 + 1  # palette ramp 1..13
    if style == "brick":
        row = yy // 16
        bx = (xx + (row % 2) * 16) % 32
        mortar = (yy % 16 == 0) | (bx == 0)
        shade = 8 + (n % 6)
        shade = np.where(mortar, 3 + (n % 3), shade)
    elif style == "tile":
        edge = (xx % 32 < 2) | (yy % 32 < 2)
        shade = 9 + ((xx // 32 + yy // 32) % 2) * 3 + (n % 3)
        shade = np.where(edge, 4, shade)
    elif style == "noise":
        shade = 4 + (n % 11)
    elif style == "water":
        shade = 6 + ((xx // 8 + yy // 8 + (n % 3)) % 8)
    elif style == "stripes":
        shade = 5 + ((xx // 4) % 2) * 6 + (n % 3)
    else:
        raise ValueError(style)
    px = ramp * 16 + np.clip(shade, 0, 15)
    # sprinkle a few fullbright texels so the colormap's identity tail is exercised
    fb = (n % 97) == 0
    px = np.where(fb, 224 + (n % 32), px)
    return px.astype(np.uint8)


def make_miptex(name: str, pixels: np.ndarray) -> bytes:
    """miptex_t + 4 mip levels (nearest-sample decimation), common/bspfile.h:95-100."""
    h, w = pixels.shape
    assert w % 16 == 0 and h % 16 == 0
    mips = [np.ascontiguousarray(pixels[:: 1 << m, :: 1 << m]) for m in range(4)]
    ofs = []
    o = 40
    for m in mips:
        ofs.append(o)
        o += m.size
    nm = name.encode("ascii")[:15]
    hdr = nm + b"\0" * (16 - len(nm)) + struct.pack("<II4I", w, h, *ofs)
    return hdr + b"".join(m.tobytes() for m in mips)


def make_sky_pixels(seed: int) -> np.ndarray:
    """256x128 sky: left half = front (cloud) layer with index 0 as transparent,
    right half = back layer (ref_soft/r_sky.c:53-63 R_InitSky)."""
    n = _lcg_noise(seed, 256, 128)
    yy, xx = np.mgrid[0:128, 0:256]
    back = 3 * 16 + 4 + ((xx // 16 + yy // 16) % 6) + (n % 2)
    cloud = ((xx // 12 + yy // 10 + (n % 2)) % 4 == 0)
    front = np.where(cloud, 16 * 1 + 10 + (n % 4), 0)
    px = np.where(xx < 128, front, back)
    return px.astype(np.uint8)


def encode_pcx(pixels: np.ndarray, palette: np.ndarray) -> bytes:
    """8-bit run-length PCX, the format LoadPCX accepts (ref_soft/pcx.c:33-95: manufacturer 0x0a,
    version 5, encoding 1, 8 bits per pixel) and WritePCXfile emits (pcx.c:103-160)."""
    import struct
    h, w = pixels.shape
    hdr = struct.pack("<BBBBHHHHHH48sBBHH58s", 0x0A, 5, 1, 8, 0, 0, w - 1, h - 1, w, h, bytes(48), 0, 1, w, 2, bytes(58))
    out = bytearray(hdr)
    for y in range(h):
        row = pixels[y]
        x = 0
        while x < w:
            v = int(row[x])
            run = 1
            while x + run < w and run < 63 and int(row[x + run]) == v:
                run += 1
            if run > 1 or (v & 0xC0) == 0xC0:
                out.append(0xC0 | run)
            out.append(v)
            x += run
    out.append(0x0C)
    out += palette.astype(np.uint8).tobytes()
    return bytes(out)


def skybox_pcx(name: str, seed: int = 0x51AB) -> dict:
    """Six 256x256 pictures for `loadsky <name>` (env/<name>{rt,bk,lf,ft,up,dn}.pcx, r_sky.c:138-187):
    a different gradient + noise pattern per side so that a swapped or mirrored face is visible."""
    pal = make_palette()
    out = {}
    for k, suf in enumerate(["rt", "bk", "lf", "ft", "up", "dn"]):
        yy, xx = np.mgrid[0:256, 0:256]
        noise = ((xx * 1103515245 + yy * 12345 + seed * (k + 1)) >> 7) & 7
        band = ((xx * (k + 1) + yy * (6 - k)) >> 4) & 15
        pix = (16 * (2 * k + 1) + ((band + noise) & 15)).astype(np.uint8)
        pix[:8, :] = 250 - k                      # a bright stripe along the top edge
        pix[:, :4] = 200 + k
        out[suf] = encode_pcx(pix, pal)
    return out


# ---------------------------------------------------------------------------------------------------
# 2D assets: gfx.wad (WAD2, common/wad.c:66-110, wad.h:30-75) and qpic .lmp files (Draw_CachePic, r_draw.c:72)

TYP_QPIC, TYP_LUMPY = 66, 64


def make_qpic(w: int, h: int, seed: int, transparent: bool = False) -> bytes:
    """qpic_t: int width, height, then w*h palette indices; index 255 (TRANSPARENT_COLOR) only when asked for."""
    pix = (_lcg_noise(seed, w, h).astype(np.uint32) * 7 + np.add.outer(np.arange(h), np.arange(w)).astype(np.uint32)) % 224
    pix = pix.astype(np.uint8)
    if transparent:
        hole = (_lcg_noise(seed ^ 0x5A5A, w, h) & 3) == 0
        pix[hole] = 255
    return struct.pack("<ii", w, h) + pix.tobytes()


def make_conchars(seed: int = 0x51AB) -> bytes:
    """128x128 character sheet, 0 = transparent (Draw_Character, r_draw.c:152-205)."""
    pix = _lcg_noise(seed, 128, 128)
    glyph = (_lcg_noise(seed ^ 0x77, 128, 128) & 1).astype(bool)
    out = np.where(glyph, (pix % 200 + 1).astype(np.uint8), 0).astype(np.uint8)
    return out.tobytes()


def make_wad2(lumps) -> bytes:
    """lumps: list of (name, type, bytes).  Header, lump data, then the info table."""
    blob = bytearray(b"WAD2" + struct.pack("<ii", len(lumps), 0))
    info = []
    for name, typ, data in lumps:
        while len(blob) & 3:
            blob.append(0)
        info.append((len(blob), len(data), typ, name))
        blob += data
    while len(blob) & 3:
        blob.append(0)
    ofs = len(blob)
    for pos, size, typ, name in info:
        blob += struct.pack("<iiibbbb16s", pos, size, size, typ, 0, 0, 0, name.upper().encode())
    struct.pack_into("<i", blob, 8, ofs)
    return bytes(blob)


def gfx_wad() -> bytes:
    return make_wad2([
        ("conchars", TYP_LUMPY, make_conchars()),
        ("disc", TYP_QPIC, make_qpic(24, 24, 11, transparent=True)),
        ("backtile", TYP_QPIC, make_qpic(64, 64, 12)),
        ("sbar", TYP_QPIC, make_qpic(320, 24, 13)),
        ("num_3", TYP_QPIC, make_qpic(24, 24, 14, transparent=True)),
        ("face1", TYP_QPIC, make_qpic(21, 19, 15, transparent=True)),      # width not a multiple of 8: the "general" loops
        ("ibar", TYP_QPIC, make_qpic(320, 24, 16)),
    ])
