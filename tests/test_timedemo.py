"""The step in front of the path: the client's scene lists and `timedemo`, in batched form
(tochris_b200/csrc/rc_scene.c).  Frames are fed to the recorder the way the reference's client feeds its globals --
V_ClearScene, one V_Add* call per entity / particle / dlight / colour shift (client/cl_view.c:85-148), the frame's
refdef_t (V_RenderView, cl_view.c:915-940) -- and rendered by RC_TimeDemo (cl_demo.c:324-365: the first frame does not
count); every frame must equal the reference's render of the same scene, including what the reference's full lists drop."""
import copy
import math
import re

import numpy as np
import pytest

import rcutil
from tochris_b200 import api
from tochris_b200.scenes import views


def _demo_specs(n):
    """a camera path through the `full` map with entities, sprites, particles, dlights and colour shifts; three frames
    overfill a list the way a busy demo does (more than MAX_DLIGHTS / MAX_PARTICLES / MAX_ENTITIES: client/ref.h:31-34)"""
    specs = rcutil.scene_views("full", n, time=0.5, time_step=0.1, entities=6, sprites=3, particles=80, dlights=3, underwater_every=7)
    for i, sp in enumerate(specs):
        sp.cshifts = [(255, 80, 0, 40 + i)] if i % 3 == 0 else []
    one = rcutil.scene_views("full", 3, time=0.5, time_step=0.1, entities=260, particles=2100, dlights=40)
    specs[2].dlights = one[0].dlights
    specs[4].particles = one[1].particles
    specs[5].entities = one[2].entities
    specs[6].cshifts = [(i * 9, 255 - i * 9, 128, 10 + i) for i in range(20)]
    return specs


def _what_the_reference_keeps(specs):
    out = copy.deepcopy(specs)
    for sp in out:
        sp.entities = sp.entities[:256]
        sp.particles = sp.particles[:2048]
        sp.dlights = sp.dlights[:32]
        sp.cshifts = sp.cshifts[:16]
    return out


def _check(lib, assets_dir, n, batch):
    from oracle import run_oracle
    w, h = 320, 240
    specs = _demo_specs(n)
    r = rcutil.render_rc(lib, assets_dir, "maps/full.bsp", w, h, specs, cvars={"_timedemo_batch": batch})
    kept = _what_the_reference_keeps(specs)
    o = run_oracle.render(assets_dir, "maps/full.bsp", w, h, kept)
    assert r["status"] & ~rcutil.RC_STAT_ALIAS_RANGE == 0, api.status_names(r["status"])
    diff = [int((o["fb"][i] != r["fb"][i]).sum()) for i in range(n)]
    assert sum(diff) == 0, f"frames that differ from the reference: {[(i, d) for i, d in enumerate(diff) if d]}"
    # what the full lists dropped: 8 dlights, 52 particles, 4 entities, 4 colour shifts
    assert r["dropped"][2] == 8 and r["dropped"][4] == 52 and r["dropped"][5] == 4 and r["dropped"][6] == 4, r["dropped"]
    assert sum(r["dropped"]) == 68
    for td in (r["timedemo"], r["timedemo_device"]):
        assert td["frames"] == n - 1                      # "the first frame didn't count" (cl_demo.c:333)
        assert td["seconds"] > 0 and abs(td["fps"] - td["frames"] / td["seconds"]) < 1e-6 * td["fps"]
        assert re.fullmatch(r"%d frames +[0-9.]+ seconds +[0-9.]+ fps\n" % (n - 1), td["text"]), td["text"]
    return r


def test_scene_recorder_and_timedemo_match_the_reference_sim(assets_dir, sim_lib, oracle_available):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    _check(sim_lib, assets_dir, 10, 4)


@pytest.mark.gpu
def test_scene_recorder_and_timedemo_match_the_reference_gpu(assets_dir, oracle_available):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    r = _check(api.LIBPATH, assets_dir, 40, 8)
    print("timedemo:", r["timedemo"]["text"].strip(), "| frames kept on the device:", r["timedemo_device"]["text"].strip())


def test_calc_fov_is_the_clients(sim_lib, assets_dir):
    """CalcFov, cl_view.c:629-639; a bad fov_x is Sys_Error there and 0 here"""
    r = api.RefCuda(assets_dir, 320, 200, libpath=sim_lib)
    for fov_x, w, h in [(90.0, 320.0, 200.0), (90.0, 640.0, 480.0), (110.0, 1280.0, 720.0), (10.0, 160.0, 120.0), (170.0, 320.0, 240.0)]:
        x = np.float32(w) / np.float32(math.tan(np.float32(fov_x) * (math.pi / 360.0)))
        want = np.float32(math.atan(np.float32(np.float32(h) / x)) * (360.0 / math.pi))
        assert abs(r.calc_fov(fov_x, w, h) - float(want)) <= 4e-6 * float(want)
        assert abs(r.calc_fov(fov_x, w, h) - views.calc_fov_y(fov_x, int(w), int(h))) < 1e-4
    assert r.calc_fov(0.5, 320, 200) == 0.0 and r.calc_fov(180.0, 320, 200) == 0.0
    r.shutdown()
