"""Host logic + kernel arithmetic on the CPU: the C-ABI host files linked with a loop over the
kernel bodies (tests/hostsim, test infrastructure) against the oracle and the golden vectors.
Sizes are kept small so the whole CPU suite runs in a few minutes."""
import os

import numpy as np
import pytest

import rcutil
from tochris_b200 import api

CASES = [
    ("box", 320, 200, 24, 1.0, 0.0, {}),
    ("small", 320, 200, 8, 0.4, 0.31, {}),
    ("multi", 640, 480, 6, 1.0, 0.0, {}),
    ("full", 320, 240, 24, 0.37, 0.173, {}),
    ("multi", 160, 120, 48, 0.0, 0.01, {}),
    ("multi", 320, 240, 12, 0.2, 0.05, dict(dlights=4)),
    ("full", 320, 200, 8, 0.9, 0.11, dict(dlights=2, underwater_every=2)),
    ("multi", 320, 240, 10, 0.3, 0.21, dict(entities=8, particles=120)),
    ("full", 640, 400, 4, 1.3, 0.4, dict(entities=20, particles=300, dlights=2, underwater_every=3)),
    ("full", 320, 240, 24, 0.37, 0.173, dict(doors=2, dlights=2)),
    ("full", 400, 300, 6, 2.0, 0.3, dict(doors=2, entities=6, particles=50)),
    ("full", 320, 240, 24, 0.37, 0.173, dict(skybox="testsky")),                     # loadsky: six box faces instead of the scrolling sky
    ("full", 400, 300, 8, 1.1, 0.3, dict(skybox="testsky", doors=2, entities=6, dlights=2)),
    ("multi", 320, 240, 12, 0.3, 0.21, dict(sprites=10)),                             # sprite entities: five orientation types, groups
    ("full", 400, 300, 8, 0.9, 0.17, dict(sprites=8, entities=6, particles=60, doors=2)),
    # BASELINE config C5 substitute: above 1280x1024 the reference needs MAXWIDTH / MAXHEIGHT / MAXSPANS raised
    ("full", 1920, 1080, 2, 1.7, 0.4, dict(dlights=8, underwater_every=2, big=True)),
]


@pytest.mark.parametrize("scene,w,h,n,t0,dt,extra", CASES)
def test_sim_matches_oracle(assets_dir, sim_lib, oracle_available, scene, w, h, n, t0, dt, extra):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle
    kw = dict(extra)
    skybox = kw.pop("skybox", None)
    cvars = {"r_skyname": skybox} if skybox else {}
    if kw.pop("big", False):
        cvars.update(r_maxsurfs=4000.0, r_maxedges=12000.0)     # same pools on both sides (r_main.c:265-300)
    if scene != "box":
        kw.update(time=t0, time_step=dt)
    specs = rcutil.scene_views(scene, n, **kw)
    o = run_oracle.render(assets_dir, f"maps/{scene}.bsp", w, h, specs, cvars=cvars)
    r = rcutil.render_rc(sim_lib, assets_dir, f"maps/{scene}.bsp", w, h, specs, cvars=cvars)
    rcutil.assert_parity(o, r, specs, w, h, allow_alias_range=False)
    return
    if not extra.get("underwater_every"):
        assert (o["z"] == r["z"]).all()
    else:
        # under water only the 320x200 warp rectangle of the z-buffer is written (r_misc.c:448-454)
        for i, sp in enumerate(specs):
            hh, ww = (min(h, 200), min(w, 320)) if sp.flags & 2 else (h, w)
            assert (o["z"][i, :hh, :ww] == r["z"][i, :hh, :ww]).all()


def test_sim_matches_golden(assets_dir, sim_lib):
    g = np.load(os.path.join(rcutil.ROOT, "tests", "golden", "smoke_small_320x200.npz"))
    specs = rcutil.scene_views("small", int(g["n"]), time=float(g["t0"]), time_step=float(g["dt"]))
    r = rcutil.render_rc(sim_lib, assets_dir, "maps/small.bsp", 320, 200, specs)
    assert rcutil.sha(r["fb"]) == str(g["fb_sha"]) and rcutil.sha(r["z"]) == str(g["z_sha"])


def check_stage_dumps(lib, assets_dir, scene="multi", w=320, h=240, n=4, **kw):
    """Per-stage differential test: the emitted edges (u, u_step, v, v2, and which surfaces they lead / trail
    by BSP key), the surfaces (key, flags, 1/z gradients) and the spans (u, v, count), not only pixels, against the dumps
    the oracle shim takes inside the reference (oracle/headless.c: R_ScanEdges / D_DrawSurfaces interposed)."""
    from oracle import run_oracle
    specs = rcutil.scene_views(scene, n, time=1.0, **kw)
    o = run_oracle.render(assets_dir, f"maps/{scene}.bsp", w, h, specs, want_dump=True)
    r = rcutil.render_rc(lib, assets_dir, f"maps/{scene}.bsp", w, h, specs, want_debug=True)
    assert r["status"] == 0
    for vi in range(n):
        oe, os_, osp, cnt = o["dumps"][vi]
        assert cnt["outofedges"] == 0 and cnt["outofsurfaces"] == 0
        re = r["edges"][vi]
        assert len(re) == len(oe)
        a = sorted((int(e["u"]), int(e["u_step"]), int(e["v"]), int(e["v2"])) for e in oe)
        b = sorted((int(e["u"]), int(e["u_step"]), int(e["v"]), int(e["v2"])) for e in re)
        assert a == b
        # the surfaces that got spans: same faces, flags and 1/z planes (d_ziorigin / d_zistepu / d_zistepv, r_rast.c:660-667)
        rs = r["surfs"][vi]
        used = set(int(sp["surf"]) for sp in osp)
        osd = sorted((int(os_[i]["face"]), int(os_[i]["flags"]), float(os_[i]["d_ziorigin"]), float(os_[i]["d_zistepu"]),
                      float(os_[i]["d_zistepv"])) for i in used)
        rsd = sorted((int(s["face"]), int(s["flags"]), float(s["d_ziorigin"]), float(s["d_zistepu"]), float(s["d_zistepv"]))
                     for s in rs[1:] if s["hasspans"] and s["flags"] >= 0)
        assert osd == rsd, (vi, len(osd), len(rsd), osd[:3], rsd[:3])
        rsp = r["spans"][(r["spans"]["view"] == vi) & (r["spans"]["count"] > 0)]
        a = sorted((int(s["u"]), int(s["v"]), int(s["count"])) for s in osp)
        b = sorted((int(s["u"]), int(s["v"]), int(s["count"])) for s in rsp)
        assert a == b


def test_stage_dumps_match(assets_dir, sim_lib, oracle_available):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    check_stage_dumps(sim_lib, assets_dir)


RECTS = [(8, 4, 304, 190), (13, 7, 291, 173), (0, 0, 317, 199), (5, 0, 40, 33), (100, 50, 9, 9)]


def _rect_specs(n):
    specs = rcutil.scene_views("full", n, time=0.8, time_step=0.19)
    for i, sp in enumerate(specs):
        sp.rect = RECTS[i % len(RECTS)]
    return specs


def check_rect_parity(o, r, specs):
    for i, sp in enumerate(specs):
        x, y, w, h = sp.rect
        assert np.array_equal(o["fb"][i, y:y + h, x:x + w], r["fb"][i, y:y + h, x:x + w]), (i, sp.rect)
        assert np.array_equal(o["z"][i, y:y + h, x:x + w], r["z"][i, y:y + h, x:x + w]), (i, sp.rect)


def test_sub_rectangles_match_oracle(assets_dir, sim_lib, oracle_available):
    """Views that use a sub-rectangle of the screen (r_refdef.vrect, r_main.c:312-417): span ends and view-rectangle
    edges that are not multiples of 8 -- the cells the draw stage treats as mixed -- compared inside the rectangle."""
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle
    specs = _rect_specs(10)
    o = run_oracle.render(assets_dir, "maps/full.bsp", 320, 200, specs)
    r = rcutil.render_rc(sim_lib, assets_dir, "maps/full.bsp", 320, 200, specs)
    assert r["status"] == 0
    check_rect_parity(o, r, specs)


def test_timerefresh_matches_oracle_sweep(assets_dir, sim_lib, oracle_available):
    """RC_TimeRefresh = R_TimeRefresh_f (r_misc.c:29-64): 128 yaw steps of one refdef, as one batch; BASELINE config C1."""
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views
    specs = views.timerefresh_views((0, 0, 0), 128)
    o = run_oracle.render(assets_dir, "maps/box.bsp", 320, 200, specs)
    r = RefCuda(assets_dir, 320, 200, libpath=sim_lib)
    r.load_world("maps/box.bsp")
    rds = views.build_refdefs(specs[:1], 320, 200, model_resolver=r.model)
    fb, secs, st = r.timerefresh(rds[0])
    assert st == 0 and secs > 0
    assert np.array_equal(fb, o["fb"])
    r.shutdown()


def test_async_call_bookkeeping(assets_dir, sim_lib):
    """RC_RenderViewsAsync / RC_WaitViews host logic (include/ref_cuda.h): RC_ASYNC_CALLS = 3 calls may be outstanding,
    a fourth is refused without side effects, RC_WaitViews answers in submission order and fails when nothing is
    outstanding; the bytes equal the blocking call's."""
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views
    r = RefCuda(assets_dir, 160, 120, libpath=sim_lib)
    r.load_world("maps/small.bsp")
    batches, want = [], []
    for b in range(5):
        specs = rcutil.scene_views("small", 2, time=0.3 + b, time_step=0.11)
        rds = views.build_refdefs(specs, 160, 120, model_resolver=r.model)
        batches.append(rds)
        fb, z, st = r.render(rds)
        assert st == 0
        want.append((fb.copy(), z.copy()))
    bufs = [(np.zeros((2, 120, 160), np.uint8), np.zeros((2, 120, 160), np.int16)) for _ in range(3)]
    got = []
    with pytest.raises(RuntimeError):
        r.wait()
    for k, rds in enumerate(batches):
        if k >= 3:
            assert r.wait() == 0
            got.append(tuple(a.copy() for a in bufs[k % 3]))
        r.render_async(rds, 2, *bufs[k % 3])
        if k == 2:
            with pytest.raises(RuntimeError):
                r.render_async(rds, 2, *bufs[0])
    for k in range(2, 5):
        assert r.wait() == 0
        got.append(tuple(a.copy() for a in bufs[k % 3]))
    with pytest.raises(RuntimeError):
        r.wait()
    assert len(got) == 5
    for (wf, wz), (gf, gz) in zip(want, got):
        assert np.array_equal(wf, gf) and np.array_equal(wz, gz)
    r.shutdown()
