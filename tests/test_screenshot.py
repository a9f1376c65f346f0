"""render.h:30 R_ScreenShot_f (ref_soft/r_misc.c:117, WritePCXfile pcx.c:103): the engine-path frame written as quakeNN.pcx,
byte for byte the file the reference writes for the same view."""
import ctypes as C
import glob
import multiprocessing as mp
import os

import numpy as np
import pytest

import rcutil


def _oracle_shot(q, assets_dir, spec):
    from oracle.refsoft import RefSoft
    from tochris_b200.scenes.views import build_refdefs
    r = RefSoft(assets_dir, 320, 200, heap_mb=96)
    r.load_world("maps/small.bsp")
    cm = C.addressof(C.c_char.in_dll(r.lib, "vid")) + 8
    rds = build_refdefs([spec], 320, 200, model_resolver=r.model, colormap_ptr=cm)
    r.render(rds[0], False)
    r.screenshot()
    q.put("ok")


def test_screenshot_file_matches_reference(assets_dir, sim_lib, oracle_available):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views
    gamedir = os.path.join(assets_dir, "id1")
    for f in glob.glob(os.path.join(gamedir, "quake*.pcx")):
        os.remove(f)
    spec = rcutil.scene_views("small", 3, time=0.4, time_step=0.31)[2]
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=_oracle_shot, args=(q, assets_dir, spec))
    p.start()
    assert q.get(timeout=120) == "ok"
    p.join(timeout=30)
    want = open(os.path.join(gamedir, "quake00.pcx"), "rb").read()
    r = RefCuda(assets_dir, 320, 200, libpath=sim_lib)
    r.load_world("maps/small.bsp")
    rds = views.build_refdefs([spec], 320, 200, model_resolver=r.model)
    vidbuf = np.zeros((200, 320), dtype=np.uint8)
    zbuf = np.zeros((200, 320), dtype=np.int16)
    r.lib.RC_SetVidBuffers(vidbuf.ctypes.data, 320, zbuf.ctypes.data)
    r.lib.R_RenderView(C.byref(rds[0]))
    r.lib.R_ScreenShot_f()                                   # quake00.pcx exists: this one becomes quake01.pcx
    got = open(os.path.join(gamedir, "quake01.pcx"), "rb").read()
    r.shutdown()
    for f in glob.glob(os.path.join(gamedir, "quake*.pcx")):
        os.remove(f)
    assert got == want
