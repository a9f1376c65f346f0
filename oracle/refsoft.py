"""TEST INFRASTRUCTURE ONLY: ctypes driver for oracle/_ref/librefsoft.so (the unmodified
reference ref_soft C path compiled by oracle/Makefile).  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import this.

The reference keeps all state in globals, so one process = one renderer instance at one
resolution; use `run_isolated()` / a subprocess for every (asset dir, resolution) pair."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "librefsoft.so")
LIB_BIG = os.path.join(HERE, "_ref", "librefsoft_big.so")   # MAXWIDTH 2047 / MAXHEIGHT 1200 / MAXSPANS 6000 (`make -C oracle big`)

sys.path.insert(0, os.path.dirname(HERE))
from tochris_b200.refdef import refdef_t  # noqa: E402


class orc_edge_t(C.Structure):
    _fields_ = [("u", C.c_int), ("u_step", C.c_int), ("v", C.c_int), ("v2", C.c_int),
                ("surfs", C.c_ushort * 2), ("nearzi", C.c_float), ("owner", C.c_int)]


class orc_surf_t(C.Structure):
    _fields_ = [("key", C.c_int), ("flags", C.c_int), ("insubmodel", C.c_int), ("face", C.c_int),
                ("nearzi", C.c_float), ("d_ziorigin", C.c_float), ("d_zistepu", C.c_float),
                ("d_zistepv", C.c_float), ("nspans", C.c_int), ("entity", C.c_int)]


class orc_span_t(C.Structure):
    _fields_ = [("u", C.c_int), ("v", C.c_int), ("count", C.c_int), ("surf", C.c_int)]


def available(big: bool = False) -> bool:
    return os.path.exists(LIB_BIG if big else LIB)


class orc_cmd2d_t(C.Structure):
    _fields_ = [("op", C.c_int), ("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int), ("arg", C.c_int),
                ("pic", C.c_char * 40), ("fromwad", C.c_int), ("table", C.c_void_p)]


class RefSoft:
    def __init__(self, basedir: str, width: int, height: int, heap_mb: int = 64, verbose: int = 0):
        big = width > 1280 or height > 1024       # the stock build's static arrays end there (r_shared.h:40-41)
        if big and not available(True):
            raise RuntimeError("oracle/_ref/librefsoft_big.so missing: run `make -C oracle big` where /root/reference exists")
        if not available():
            raise RuntimeError("oracle/_ref/librefsoft.so missing: run `make -C oracle` where /root/reference exists")
        self.lib = C.CDLL(LIB_BIG if big else LIB)
        L = self.lib
        L.orc_last_error.restype = C.c_char_p
        L.orc_model.restype = C.c_void_p
        L.orc_load_world.restype = C.c_void_p
        L.orc_end_frame.argtypes = [C.c_void_p]
        L.orc_draw2d.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_set_cvar.argtypes = [C.c_char_p, C.c_float]
        L.orc_set_cvar_string.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_render.argtypes = [C.POINTER(refdef_t), C.c_void_p, C.c_void_p]
        L.orc_render_batch.argtypes = [C.POINTER(refdef_t), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        self.width, self.height = width, height
        if L.orc_init(basedir.encode(), width, height, heap_mb, verbose) != 0:
            raise RuntimeError("orc_init: " + L.orc_last_error().decode())

    def err(self) -> str:
        return self.lib.orc_last_error().decode()

    def load_world(self, name: str) -> int:
        p = self.lib.orc_load_world(name.encode())
        if not p:
            raise RuntimeError("orc_load_world: " + self.err())
        return p

    def model(self, name: str) -> int:
        p = self.lib.orc_model(name.encode())
        if not p:
            raise RuntimeError("orc_model: " + self.err())
        return p

    def set_cvar(self, name: str, value):
        if isinstance(value, str):
            if self.lib.orc_set_cvar_string(name.encode(), value.encode()) != 0:
                raise KeyError(name)
        elif self.lib.orc_set_cvar(name.encode(), value) != 0:
            raise KeyError(name)

    def struct_sizes(self):
        buf = (C.c_int * 16)()
        n = self.lib.orc_struct_sizes(buf, 16)
        return list(buf[:n])

    def clear(self, fbval: int = 0, zval: int = 0):
        self.lib.orc_clear_buffers(fbval, zval)

    def flush_caches(self):
        self.lib.orc_flush_caches()

    def render(self, rd: refdef_t, want_z: bool = True):
        fb = np.empty((self.height, self.width), dtype=np.uint8)
        z = np.empty((self.height, self.width), dtype=np.int16) if want_z else None
        rc = self.lib.orc_render(C.byref(rd), fb.ctypes.data, z.ctypes.data if want_z else None)
        if rc != 0:
            raise RuntimeError("orc_render: " + self.err())
        return fb, z

    def draw2d(self, cmds):
        """cmds: list of dicts (op, x, y, w, h, arg, pic, fromwad, table) -> the reference's Draw_* calls on vid.buffer
        as the last render() left it; returns the frame buffer."""
        arr = (orc_cmd2d_t * max(1, len(cmds)))()
        keep = []
        for i, c in enumerate(cmds):
            a = arr[i]
            a.op, a.x, a.y, a.w, a.h, a.arg = c["op"], c.get("x", 0), c.get("y", 0), c.get("w", 0), c.get("h", 0), c.get("arg", 0)
            a.pic = c.get("pic", "").encode()
            a.fromwad = int(c.get("fromwad", 0))
            if c.get("table") is not None:
                t = (C.c_ubyte * 256)(*c["table"])
                keep.append(t)
                a.table = C.cast(t, C.c_void_p)
        fb = np.empty((self.height, self.width), dtype=np.uint8)
        if self.lib.orc_draw2d(arr, len(cmds), fb.ctypes.data) != 0:
            raise RuntimeError("orc_draw2d: " + self.err())
        return fb

    def screenshot(self):
        if self.lib.orc_screenshot() != 0:
            raise RuntimeError("orc_screenshot: " + self.err())

    def point_contents(self, org) -> int:
        a = (C.c_float * 3)(*org)
        return self.lib.orc_point_contents(a)

    def end_frame(self):
        """R_EndFrame after render() of the same view: the palette R_UpdatePalette uploaded (768 bytes), or None
        when the reference skipped the upload (no colour shifts and gamma unchanged)."""
        pal = np.empty(768, dtype=np.uint8)
        rc = self.lib.orc_end_frame(pal.ctypes.data)
        if rc < 0:
            raise RuntimeError("orc_end_frame: " + self.err())
        return pal if rc == 1 else None

    def render_batch(self, rds, want_fb: bool = True, want_z: bool = True):
        n = len(rds)
        fb = np.empty((n, self.height, self.width), dtype=np.uint8) if want_fb else None
        z = np.empty((n, self.height, self.width), dtype=np.int16) if want_z else None
        secs = C.c_double(0)
        rc = self.lib.orc_render_batch(rds, n, fb.ctypes.data if want_fb else None,
                                       z.ctypes.data if want_z else None, C.byref(secs))
        if rc != 0:
            raise RuntimeError("orc_render_batch: " + self.err())
        return fb, z, secs.value

    # ---- stage dumps (edges before the scan, surfaces and spans as drawn) ----
    def dump_enable(self, on: bool = True, maxedges: int = 20000, maxsurfs: int = 8000, maxspans: int = 400000):
        self.lib.orc_dump_enable(int(on), maxedges, maxsurfs, maxspans)

    def dump_get(self):
        pe, ps, pp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        ne, ns, np_ = C.c_int(), C.c_int(), C.c_int()
        self.lib.orc_dump_get(C.byref(pe), C.byref(ne), C.byref(ps), C.byref(ns), C.byref(pp), C.byref(np_))

        def arr(ptr, n, ty):
            if n == 0:
                return np.zeros(0, dtype=np.dtype(ty))
            a = (ty * n).from_address(ptr.value)
            return np.frombuffer(a, dtype=np.dtype(ty)).copy()
        return arr(pe, ne.value, orc_edge_t), arr(ps, ns.value, orc_surf_t), arr(pp, np_.value, orc_span_t)

    def counters(self):
        a = (C.c_int * 8).in_dll(self.lib, "orc_counters")
        return dict(outofsurfaces=a[0], outofedges=a[1], nsurfs=a[2], nedges=a[3], c_surf=a[4])
