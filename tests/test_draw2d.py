"""SURVEY.md 8(f) rank 2, second half: the 2D overlay (ref_soft/r_draw.c) -- characters, pictures (opaque,
transparent, translated, widths that are / are not multiples of 8), the stretched console background with its version
stamp, tile clear, fill and the menu fade -- applied in order on top of rendered views and compared byte for byte with
the reference's own Draw_* functions run on its frame buffer (oracle/_ref)."""
import numpy as np
import pytest

import rcutil

CHAR, PIC, TRANSPIC, TRANSLATE, CONBACK, TILECLEAR, FILL, FADE = range(1, 9)
VERSION = "(Linux TOChriSQuake 1.30) 1.67"       # r_draw.c:424-426 in a __linux__ build of the reference (quakedef.h:25-27)


def overlay(i, w, h):
    """A HUD-like command list for view i; every kind of command, overlapping on purpose (order matters)."""
    tab = [(7 * k + i) & 255 for k in range(256)]
    o = []
    if i % 3 == 0:
        o.append(dict(op=CONBACK, arg=h // 2 + 3 * i))
    if i % 4 == 1:
        o.append(dict(op=TILECLEAR, x=0, y=h - 48, w=w, h=48))
        o.append(dict(op=TILECLEAR, x=5 + i, y=3, w=100, h=70))
    o.append(dict(op=PIC, x=0, y=h - 24, pic="sbar", fromwad=1))
    o.append(dict(op=PIC, x=w - 24 - i, y=4, pic="disc", fromwad=1))            # alpha picture through Draw_Pic
    o.append(dict(op=TRANSPIC, x=24 + 3 * i, y=h - 24, pic="num_3", fromwad=1))
    o.append(dict(op=TRANSPIC, x=100 + i, y=h - 23, pic="face1", fromwad=1))    # width 21: the general loop
    o.append(dict(op=TRANSLATE, x=60 + i, y=30, pic="gfx/menuplyr.lmp", fromwad=0, table=tab))
    o.append(dict(op=FILL, x=10, y=10 + i, w=37, h=5, arg=(31 * i + 4) & 255))
    for k, ch in enumerate("Batch %d!" % i):
        o.append(dict(op=CHAR, x=8 + 8 * k, y=40 + i, arg=ord(ch) + (128 if k & 1 else 0)))
    o.append(dict(op=CHAR, x=16, y=-3, arg=65))           # clipped at the top
    o.append(dict(op=CHAR, x=16, y=-8, arg=66))           # off screen: ignored
    o.append(dict(op=CHAR, x=w - 7, y=50, arg=67))        # too far right: ignored
    o.append(dict(op=CHAR, x=40, y=h - 7, arg=68))        # too low: ignored
    if i % 5 == 2:
        o.append(dict(op=FADE))
        o.append(dict(op=PIC, x=w // 2 - 160, y=h // 2 - 12, pic="ibar", fromwad=1))
    return o


def _run(lib, assets_dir, w, h, n):
    from oracle import run_oracle
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views
    specs = rcutil.scene_views("small", n, time=0.4, time_step=0.31)
    ovs = [overlay(i, w, h) for i in range(n)]
    o = run_oracle.render(assets_dir, "maps/small.bsp", w, h, specs, overlays=ovs)
    r = RefCuda(assets_dir, w, h, libpath=lib)
    r.load_world("maps/small.bsp")
    r.set_console_version(VERSION)
    rds = views.build_refdefs(specs, w, h, model_resolver=r.model)
    fb, _, st = r.render(rds)
    assert st == 0
    fb = np.ascontiguousarray(fb)
    assert r.draw_overlay(ovs, fb) == 0, r.err()
    for i in range(n):
        assert np.array_equal(fb[i], o["fb"][i]), f"view {i}: {int((fb[i] != o['fb'][i]).sum())} pixels differ"
    # what the reference Sys_Errors on is refused before anything is drawn
    bad = [[dict(op=PIC, x=w - 100, y=0, pic="sbar", fromwad=1)]]
    keep = fb[:1].copy()
    assert r.draw_overlay(bad, keep) != 0 and "bad coordinates" in r.err()
    assert np.array_equal(keep, fb[:1])
    r.shutdown()


@pytest.mark.parametrize("w,h", [(320, 200), (400, 300)])
def test_overlay_matches_reference_sim(assets_dir, sim_lib, oracle_available, w, h):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    _run(sim_lib, assets_dir, w, h, 7)


@pytest.mark.gpu
def test_overlay_matches_reference_gpu(assets_dir, oracle_available):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from tochris_b200 import api
    _run(api.LIBPATH, assets_dir, 640, 480, 16)
