"""ctypes binding of libref_cuda.so (include/ref_cuda.h).

The library is the product; this module only marshals arguments.  There is no CPU
rendering path: if the CUDA library is missing or no device is present, RefCuda() raises.
`libpath` exists so the test-suite can point the same binding at the host-simulation
build of the C-ABI (tests/hostsim), which is test infrastructure, not a fallback."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .refdef import refdef_t

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("RC_LIBPATH") or os.path.join(HERE, "libref_cuda.so")   # RC_LIBPATH: kernel-variant experiments

RC_STAT = {1: "EDGE_OVERFLOW", 2: "SURF_OVERFLOW", 4: "SPAN_POOL", 8: "AET_OVERFLOW", 16: "STACK_OVERFLOW",
           32: "CACHE_POOL", 64: "STALE_CLIPVERT", 128: "FACELIST", 512: "ALIAS_RANGE", 1024: "BMODEL_OVERFLOW"}


class rc_config_t(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("device", C.c_int), ("basedir", C.c_char_p),
                ("r_clearcolor", C.c_float), ("r_fullbright", C.c_float), ("r_ambient", C.c_float),
                ("r_drawentities", C.c_float), ("r_drawviewmodel", C.c_float), ("r_waterwarp", C.c_float),
                ("r_fastsky", C.c_float), ("r_skycolor", C.c_float), ("r_lerpmodels", C.c_float),
                ("d_mipcap", C.c_float), ("d_mipscale", C.c_float), ("r_maxsurfs", C.c_float),
                ("r_maxedges", C.c_float),
                ("surfcache_bytes", C.c_size_t), ("max_views_in_flight", C.c_int), ("host_threads", C.c_int),
                ("r_draworder", C.c_float)]


class rc_stats_t(C.Structure):
    _fields_ = [("kernel_launches", C.c_int),
                ("edges", C.c_longlong), ("surfs", C.c_longlong), ("spans", C.c_longlong), ("blocks", C.c_longlong),
                ("cache_texels", C.c_longlong), ("cache_jobs", C.c_longlong), ("faces", C.c_longlong),
                ("ms_walk", C.c_float), ("ms_faces", C.c_float), ("ms_scan", C.c_float), ("ms_surfprep", C.c_float),
                ("ms_cache", C.c_float), ("ms_draw", C.c_float), ("ms_total", C.c_float),
                ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("ms_spanseg", C.c_float), ("ms_cells", C.c_float),
                ("fullcells", C.c_longlong), ("slowcells", C.c_longlong),
                ("ms_alias", C.c_float), ("ms_particles", C.c_float), ("ms_zresolve", C.c_float), ("pad_", C.c_float),
                ("alias_range_events", C.c_longlong), ("alias_fragments", C.c_longlong),
                ("alias_fragments_front", C.c_longlong), ("alias_spans", C.c_longlong)]


class rc_timedemo_t(C.Structure):    # include/ref_cuda.h rc_timedemo_t
    _fields_ = [("frames", C.c_int), ("seconds", C.c_double), ("fps", C.c_double), ("status", C.c_int), ("text", C.c_char * 64)]


class rc_modelcache_t(C.Structure):   # include/ref_cuda.h rc_modelcache_t
    _fields_ = [("loaded", C.c_int), ("resident", C.c_int), ("resident_bytes", C.c_size_t), ("budget_bytes", C.c_size_t),
                ("uploads", C.c_int), ("evictions", C.c_int), ("rebuilds", C.c_int)]


class rc_edge_t(C.Structure):
    _fields_ = [("u", C.c_int), ("u_step", C.c_int), ("v", C.c_short), ("v2", C.c_short),
                ("surfs", C.c_ushort * 2), ("rank", C.c_uint), ("pad", C.c_int * 3)]


class rc_surf_t(C.Structure):       # rc_types.h rc_surf_t, 128 B
    _fields_ = [("key", C.c_int), ("flags", C.c_int), ("face", C.c_int), ("insubmodel", C.c_int),
                ("nearzi", C.c_float), ("d_ziorigin", C.c_float), ("d_zistepu", C.c_float), ("d_zistepv", C.c_float),
                ("miplevel", C.c_int), ("hasspans", C.c_int), ("entity", C.c_int), ("pad0", C.c_int),
                ("d_sdivzstepv", C.c_float), ("d_tdivzstepv", C.c_float), ("d_sdivzorigin", C.c_float), ("d_tdivzorigin", C.c_float),
                ("cacheofs", C.c_int), ("cachewidth", C.c_int), ("srckind", C.c_int), ("drawflags", C.c_int),
                ("d_sdivzstepu", C.c_float), ("d_tdivzstepu", C.c_float), ("d_zistepu_draw", C.c_float), ("izistep", C.c_int),
                ("sadjust", C.c_int), ("tadjust", C.c_int), ("bbextents", C.c_int), ("bbextentt", C.c_int),
                ("pad2", C.c_int * 4)]


class rc_draw2d_t(C.Structure):    # include/ref_cuda.h rc_draw2d_t
    _fields_ = [("op", C.c_int), ("x", C.c_int), ("y", C.c_int), ("w", C.c_int), ("h", C.c_int), ("arg", C.c_int),
                ("pic", C.c_int), ("table", C.c_int)]


class rc_span_t(C.Structure):       # rc_types.h rc_span_t, 32 B
    _fields_ = [("u", C.c_short), ("count", C.c_short), ("pixofs", C.c_int), ("segbase", C.c_int), ("gsurf", C.c_int),
                ("iziorg", C.c_int), ("izistep", C.c_int), ("cacheofs", C.c_int), ("cwk", C.c_uint)]


assert C.sizeof(rc_surf_t) == 128 and C.sizeof(rc_span_t) == 32 and C.sizeof(rc_edge_t) == 32


def status_names(status: int):
    return [n for b, n in RC_STAT.items() if status & b]


GROUPFN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int)


class RefCuda:
    """One renderer instance per process (the C library keeps one global renderer, as the
    reference does)."""

    def __init__(self, basedir: str, width: int, height: int, device: int = 0, libpath: str | None = None, **cvars):
        path = libpath or LIBPATH
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        self.lib = C.CDLL(path)
        L = self.lib
        L.RC_LastError.restype = C.c_char_p
        L.RC_RenderViews.argtypes = [C.POINTER(refdef_t), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.RC_SetVidBuffers.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.RC_SetVidBuffers.restype = None
        L.R_RenderView.argtypes = [C.POINTER(refdef_t)]
        L.R_RenderView.restype = None
        L.R_ScreenShot_f.restype = None
        L.RC_TimeRefresh.argtypes = [C.POINTER(refdef_t), C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.RC_SetLoadLits.argtypes = [C.c_int]
        L.RC_SetLoadLits.restype = None
        L.RC_DrawPicFromWad.argtypes = [C.c_char_p]
        L.RC_DrawCachePic.argtypes = [C.c_char_p]
        L.RC_DrawTranslation.argtypes = [C.c_void_p]
        L.RC_SetConsoleVersion.argtypes = [C.c_char_p]
        L.RC_SetConsoleVersion.restype = None
        L.RC_DrawOverlay.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.RC_DrawOverlayDevice.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.RC_SetGamma.argtypes = [C.c_float]
        L.RC_SetGamma.restype = None
        L.RC_ViewPalette.argtypes = [C.POINTER(refdef_t), C.c_void_p]
        L.RC_PresentViews.argtypes = [C.POINTER(refdef_t), C.c_int, C.c_void_p, C.c_void_p]
        L.RC_PresentViewsDevice.argtypes = [C.POINTER(refdef_t), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.RC_RenderViewsAsync.argtypes = [C.POINTER(refdef_t), C.c_int, C.c_void_p, C.c_void_p]
        L.RC_WaitViews.argtypes = [C.POINTER(C.c_int)]
        L.RC_RenderViewsDevice.argtypes = [C.POINTER(refdef_t), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.RC_DebugRead.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.RC_SetSkybox.argtypes = [C.c_char_p]
        L.RC_DeviceAlloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
        L.RC_DeviceFree.argtypes = [C.c_void_p]
        L.RC_IpcExport.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte)]
        L.RC_IpcImport.argtypes = [C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]
        L.RC_IpcClose.argtypes = [C.c_void_p]
        L.RC_SetMirrorOutput.argtypes = [C.c_void_p, C.c_void_p]
        L.RC_SetMirrorGroups.argtypes = [C.c_int]
        L.RC_SetGroupCallback.argtypes = [GROUPFN, C.c_void_p, C.c_void_p, C.c_int]
        L.RC_SetGroupCallback.restype = None
        L.RC_SceneNew.restype = C.c_void_p
        L.RC_SceneFree.argtypes = [C.c_void_p]
        L.RC_SceneFree.restype = None
        L.RC_SceneClear.argtypes = [C.c_void_p]
        L.RC_SceneClear.restype = None
        L.RC_SceneReset.argtypes = [C.c_void_p]
        L.RC_SceneReset.restype = None
        L.RC_SceneAddEntity.argtypes = [C.c_void_p, C.c_void_p]
        L.RC_SceneAddParticle.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int]
        L.RC_SceneAddDlight.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_float]
        L.RC_SceneAddCshift.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.RC_SceneEndFrame.argtypes = [C.c_void_p, C.POINTER(refdef_t), C.POINTER(C.c_int)]
        L.RC_SceneFrames.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.RC_SceneFrames.restype = C.POINTER(refdef_t)
        L.RC_CalcFov.argtypes = [C.c_float, C.c_float, C.c_float]
        L.RC_CalcFov.restype = C.c_float
        L.RC_TimeDemo.argtypes = [C.POINTER(refdef_t), C.c_int, C.c_int, C.c_void_p, C.POINTER(rc_timedemo_t)]
        L.RC_SetModelCacheBytes.argtypes = [C.c_size_t]
        L.RC_SetModelCacheBytes.restype = None
        L.RC_ModelCacheInfo.argtypes = [C.POINTER(rc_modelcache_t)]
        L.RC_ModelCacheInfo.restype = None
        L.Mod_TouchModel.argtypes = [C.c_char_p]
        L.Mod_TouchModel.restype = None
        L.Mod_ForName.restype = C.c_void_p
        L.Mod_ForName.argtypes = [C.c_char_p, C.c_int]
        cfg = rc_config_t()
        L.RC_DefaultConfig(C.byref(cfg))
        cfg.width, cfg.height, cfg.device = width, height, device
        self._basedir = basedir.encode()
        cfg.basedir = self._basedir
        loadlits = cvars.pop("r_loadlits", 0)            # not a field of rc_config_t: its own call (RC_SetLoadLits)
        gamma = cvars.pop("gamma", None)
        for k, v in cvars.items():
            setattr(cfg, k, v)
        self.width, self.height = width, height
        rc = L.RC_Init(C.byref(cfg))
        if rc != 0:
            raise RuntimeError(f"RC_Init failed ({rc}): {L.RC_LastError().decode()}")
        L.RC_SetLoadLits(int(bool(loadlits)))
        if gamma is not None:
            L.RC_SetGamma(float(gamma))

    def err(self) -> str:
        return self.lib.RC_LastError().decode()

    def load_world(self, name: str):
        rc = self.lib.RC_LoadWorld(name.encode())
        if rc != 0:
            raise RuntimeError(f"RC_LoadWorld({name}) failed ({rc}): {self.err()}")

    def model(self, name: str) -> int:
        p = self.lib.Mod_ForName(name.encode(), 0)
        if not p:
            raise RuntimeError(f"Mod_ForName({name}): {self.err()}")
        return p

    def render(self, rds, n: int | None = None, want_z: bool = True, fb=None, z=None):
        """rds: ctypes array of refdef_t.  Returns (fb[n,h,w] uint8, z[n,h,w] int16 | None, status)."""
        n = len(rds) if n is None else n
        if fb is None:
            fb = np.empty((n, self.height, self.width), dtype=np.uint8)
        if want_z and z is None:
            z = np.empty((n, self.height, self.width), dtype=np.int16)
        st = C.c_int(0)
        rc = self.lib.RC_RenderViews(rds, n, fb.ctypes.data, z.ctypes.data if want_z else None, C.byref(st))
        if rc != 0:
            raise RuntimeError(f"RC_RenderViews failed ({rc}): {self.err()}")
        return fb, (z if want_z else None), st.value

    # ---- 2D overlay (r_draw.c) ----
    def set_console_version(self, text: str):
        self.lib.RC_SetConsoleVersion(text.encode())

    def draw_init(self):
        if self.lib.RC_DrawInit() != 0:
            raise RuntimeError(f"RC_DrawInit failed: {self.err()}")

    def pic(self, name: str, fromwad: bool) -> int:
        h = (self.lib.RC_DrawPicFromWad if fromwad else self.lib.RC_DrawCachePic)(name.encode())
        if h <= 0:
            raise RuntimeError(f"picture {name}: {self.err()}")
        return h

    def translation(self, table) -> int:
        t = (C.c_ubyte * 256)(*table)
        h = self.lib.RC_DrawTranslation(t)
        if h <= 0:
            raise RuntimeError(f"RC_DrawTranslation: {self.err()}")
        return h

    def pack_overlays(self, overlays):
        """overlays: per view, a list of dicts like oracle.refsoft.RefSoft.draw2d takes -> (rc_draw2d_t[], first[])."""
        total = sum(len(o) for o in overlays)
        cmds = (rc_draw2d_t * max(1, total))()
        first = (C.c_int * (len(overlays) + 1))()
        k = 0
        for i, o in enumerate(overlays):
            first[i] = k
            for c in o:
                a = cmds[k]
                a.op, a.x, a.y, a.w, a.h, a.arg = c["op"], c.get("x", 0), c.get("y", 0), c.get("w", 0), c.get("h", 0), c.get("arg", 0)
                if c.get("pic"):
                    a.pic = self.pic(c["pic"], bool(c.get("fromwad", 0)))
                if c.get("table") is not None:
                    a.table = self.translation(c["table"])
                k += 1
        first[len(overlays)] = k
        return cmds, first

    def draw_overlay(self, overlays, fb: np.ndarray) -> int:
        """RC_DrawOverlay: applies each view's command list to its plane of `fb` (host, in place); returns the C return code."""
        cmds, first = self.pack_overlays(overlays)
        assert fb.flags["C_CONTIGUOUS"] and fb.dtype == np.uint8
        return self.lib.RC_DrawOverlay(cmds, first, len(overlays), fb.ctypes.data)

    def timerefresh(self, rd, want_frames: bool = True):
        """RC_TimeRefresh: the reference's `timerefresh` sweep (128 yaw steps of `rd`) as one batch -> (frames or None, seconds, status)."""
        fb = np.empty((128, self.height, self.width), dtype=np.uint8) if want_frames else None
        secs, st = C.c_double(0), C.c_int(0)
        rc = self.lib.RC_TimeRefresh(C.byref(rd), fb.ctypes.data if want_frames else None, C.byref(secs), C.byref(st))
        if rc != 0:
            raise RuntimeError(f"RC_TimeRefresh failed ({rc}): {self.err()}")
        return fb, secs.value, st.value

    # ---- the client's scene lists and timedemo (rc_scene.c; cl_view.c:67-148,889-943, cl_demo.c:324-365) ----
    def calc_fov(self, fov_x: float, width: float, height: float) -> float:
        return float(self.lib.RC_CalcFov(fov_x, width, height))

    def record_scene(self, rds, n: int | None = None, add=None):
        """Feeds every refdef_t of `rds` to the scene recorder the way the client would -- V_ClearScene, then one V_Add*
        call per entity, particle, dlight and colour shift, then the frame's refdef_t -- and returns (scene handle,
        refdef_t pointer, number of frames, [per frame: how many Add calls the full lists dropped])."""
        L = self.lib
        n = len(rds) if n is None else n
        s = L.RC_SceneNew()
        addv = (C.c_int * 4)(*add) if add is not None else None
        dropped = []
        for i in range(n):
            rd = rds[i]
            L.RC_SceneClear(s)
            d = 0
            for j in range(rd.num_entities):
                d += 1 - L.RC_SceneAddEntity(s, C.byref(rd.entities[j]))
            for j in range(rd.num_particles):
                d += 1 - L.RC_SceneAddParticle(s, rd.particles[j].org, rd.particles[j].color)
            for j in range(rd.num_dlights):
                d += 1 - L.RC_SceneAddDlight(s, rd.dlights[j].origin, rd.dlights[j].radius)
            for j in range(rd.num_cshifts):
                c = rd.cshifts[j]
                d += 1 - L.RC_SceneAddCshift(s, c.destcolor[0], c.destcolor[1], c.destcolor[2], c.percent)
            dropped.append(d)
            if L.RC_SceneEndFrame(s, C.byref(rd), addv) != i:
                raise RuntimeError("RC_SceneEndFrame failed")
        cnt = C.c_int(0)
        frames = L.RC_SceneFrames(s, C.byref(cnt))
        return s, frames, cnt.value, dropped

    def scene_free(self, s):
        self.lib.RC_SceneFree(s)

    def timedemo(self, frames, n: int, batch: int = 64, fb=None) -> dict:
        """RC_TimeDemo over a refdef_t array; fb: optional (n, h, w) uint8 array that receives every frame."""
        out = rc_timedemo_t()
        rc = self.lib.RC_TimeDemo(frames, n, batch, fb.ctypes.data if fb is not None else None, C.byref(out))
        if rc != 0:
            raise RuntimeError(f"RC_TimeDemo failed ({rc}): {self.err()}")
        return dict(frames=out.frames, seconds=out.seconds, fps=out.fps, status=out.status, text=out.text.decode())

    def point_contents(self, org) -> int:
        a = (C.c_float * 3)(*org)
        return self.lib.RC_PointContents(a)

    def view_flags_for_origin(self, org) -> int:
        a = (C.c_float * 3)(*org)
        return self.lib.RC_ViewFlagsForOrigin(a)

    def set_gamma(self, gamma: float):
        self.lib.RC_SetGamma(gamma)

    def view_palette(self, rd) -> np.ndarray:
        """RC_ViewPalette: the 768-byte palette R_UpdatePalette builds for one view."""
        pal = np.empty(768, dtype=np.uint8)
        if self.lib.RC_ViewPalette(C.byref(rd), pal.ctypes.data) != 0:
            raise RuntimeError(f"RC_ViewPalette failed: {self.err()}")
        return pal

    def present(self, rds, fb: np.ndarray) -> np.ndarray:
        """RC_PresentViews: 8-bit frame buffers (host) -> 32-bit display pixels (host), each view through its palette."""
        n = fb.shape[0]
        out = np.empty(fb.shape, dtype=np.uint32)
        fbc = np.ascontiguousarray(fb)
        if self.lib.RC_PresentViews(rds, n, fbc.ctypes.data, out.ctypes.data) != 0:
            raise RuntimeError(f"RC_PresentViews failed: {self.err()}")
        return out

    def present_device(self, rds, n: int, fb_ptr: int, rgbx_ptr: int, stream: int = 0):
        if self.lib.RC_PresentViewsDevice(rds, n, fb_ptr, rgbx_ptr, stream or None) != 0:      # rds None: previous palettes
            raise RuntimeError(f"RC_PresentViewsDevice failed: {self.err()}")

    def render_async(self, rds, n: int, fb, z=None) -> None:
        """RC_RenderViewsAsync: enqueue a batch into caller-owned (pinned) numpy buffers; at most three outstanding (RC_ASYNC_CALLS)."""
        rc = self.lib.RC_RenderViewsAsync(rds, n, fb.ctypes.data, z.ctypes.data if z is not None else None)
        if rc != 0:
            raise RuntimeError(f"RC_RenderViewsAsync failed ({rc}): {self.err()}")

    def wait(self) -> int:
        """RC_WaitViews: block until the oldest outstanding batch has arrived; returns its status bits."""
        st = C.c_int(0)
        rc = self.lib.RC_WaitViews(C.byref(st))
        if rc != 0:
            raise RuntimeError(f"RC_WaitViews failed ({rc}): {self.err()}")
        return st.value

    def render_device(self, rds, n: int, fb_ptr: int, z_ptr: int = 0, stream: int = 0) -> int:
        st = C.c_int(0)
        rc = self.lib.RC_RenderViewsDevice(rds, n, fb_ptr, z_ptr or None, stream or None, C.byref(st))
        if rc != 0:
            raise RuntimeError(f"RC_RenderViewsDevice failed ({rc}): {self.err()}")
        return st.value

    def set_skybox(self, name: str | None):
        """`loadsky <name>` / r_skyname (r_sky.c:138-194); None or '' switches the sky box off."""
        rc = self.lib.RC_SetSkybox((name or "").encode())
        if rc != 0:
            raise RuntimeError(f"RC_SetSkybox({name}) failed ({rc}): {self.err()}")

    def set_group_callback(self, fn, side_stream: int = 0, ngroups: int = 4):
        """fn(first_view, num_views) is called after each group of views of a render_device() batch has been
        enqueued; side_stream (raw cudaStream_t) already waits for the group.  fn None switches it off."""
        if fn is None:
            self._group_cb = None
            self.lib.RC_SetGroupCallback(GROUPFN(), None, None, 0)     # a NULL function pointer
            return
        self._group_cb = GROUPFN(lambda user, first, num: fn(first, num))
        self.lib.RC_SetGroupCallback(self._group_cb, None, C.c_void_p(side_stream), ngroups)

    # ---- multi-GPU output placement (RC_DeviceAlloc / RC_Ipc*) ----
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        if self.lib.RC_DeviceAlloc(C.c_size_t(nbytes), C.byref(p)) != 0:
            raise RuntimeError("RC_DeviceAlloc failed")
        return p.value

    def device_free(self, ptr: int):
        self.lib.RC_DeviceFree(C.c_void_p(ptr))

    def ipc_close(self, ptr: int):
        self.lib.RC_IpcClose(C.c_void_p(ptr))

    def ipc_export(self, ptr: int) -> bytes:
        h = (C.c_ubyte * 64)()
        if self.lib.RC_IpcExport(C.c_void_p(ptr), h) != 0:
            raise RuntimeError("RC_IpcExport failed")
        return bytes(h)

    def ipc_import(self, handle: bytes) -> int:
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        if self.lib.RC_IpcImport(h, C.byref(p)) != 0:
            raise RuntimeError("RC_IpcImport failed")
        return p.value

    def set_mirror_output(self, fb_ptr: int = 0, z_ptr: int = 0):
        """RC_SetMirrorOutput: finished view groups of render_device() are also copied to fb_ptr (peer or host)."""
        self.lib.RC_SetMirrorOutput(C.c_void_p(fb_ptr or None), C.c_void_p(z_ptr or None))

    def set_mirror_groups(self, ngroups: int):
        self.lib.RC_SetMirrorGroups(ngroups)

    def set_mirror_deferred(self, on: bool):
        """RC_SetMirrorDeferred: the mirror copy of a batch overlaps the next batch; mirror_fence() before reading the mirror."""
        self.lib.RC_SetMirrorDeferred(int(on))

    def mirror_fence(self, stream: int = 0):
        if self.lib.RC_MirrorFence(C.c_void_p(stream or None)) != 0:
            raise RuntimeError(f"RC_MirrorFence failed: {self.err()}")

    def flush_caches(self):
        self.lib.RC_FlushCaches()

    def set_model_cache_bytes(self, nbytes: int):
        """device budget of the alias models (0: everything loaded stays resident); least recently used leave first"""
        self.lib.RC_SetModelCacheBytes(nbytes)

    def touch_model(self, name: str):
        self.lib.Mod_TouchModel(name.encode())

    def model_cache_info(self) -> dict:
        s = rc_modelcache_t()
        self.lib.RC_ModelCacheInfo(C.byref(s))
        return {k: getattr(s, k) for k, _ in rc_modelcache_t._fields_}

    def set_profiling(self, on: bool):
        self.lib.RC_SetProfiling(int(on))

    def stats(self) -> dict:
        s = rc_stats_t()
        self.lib.RC_GetStats(C.byref(s))
        return {k: getattr(s, k) for k, _ in rc_stats_t._fields_}

    def debug_read(self, kind: int, view: int = 0, cap: int = 1 << 20):
        ty = [rc_edge_t, rc_surf_t, rc_span_t][kind]
        buf = (ty * cap)()
        n = self.lib.RC_DebugRead(kind, view, buf, cap)
        if n < 0:
            raise RuntimeError("RC_DebugRead failed")
        a = np.frombuffer(buf, dtype=np.dtype(ty))[:n].copy()
        if kind == 2:
            # scanline and view are packed / implied in the record: unpack them for the stage-dump tests
            out = np.zeros(n, dtype=a.dtype.descr + [("v", "<i4"), ("view", "<i4")])
            for name in a.dtype.names:
                out[name] = a[name]
            out["v"] = a["cwk"] >> 20
            out["view"] = a["pixofs"] // (self.width * self.height)
            return out
        return a

    def shutdown(self):
        self.lib.RC_Shutdown()
