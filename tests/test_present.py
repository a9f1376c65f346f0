"""the end of a frame.  R_UpdatePalette's palette (colour shifts in 8.8 fixed
point + the gamma table, ref_soft/r_main.c:794-875) and the 8-bit -> 32-bit expansion through it, against the
reference run here (oracle/_ref: R_RenderView then R_EndFrame, palette captured in VID_ShiftPalette)."""
import numpy as np
import pytest

import rcutil

CSHIFTS = [
    [(255, 0, 0, 64)],                                    # damage flash
    [(130, 80, 50, 128), (0, 255, 0, 32)],                # water tint + a second shift on top
    [(0, 0, 0, 256)],                                     # full black-out
    [(255, 255, 255, 1), (10, 20, 30, 255), (200, 100, 0, 77), (1, 2, 3, 4)],   # NUM_CSHIFTS = 4
]


def _specs(n):
    specs = rcutil.scene_views("small", n, time=0.4, time_step=0.31)
    for i, sp in enumerate(specs):
        sp.cshifts = list(CSHIFTS[i % len(CSHIFTS)])
    return specs


def _expand(fb, pals):
    out = np.empty(fb.shape, dtype=np.uint32)
    for i, pal in enumerate(pals):
        p = pal.reshape(256, 3).astype(np.uint32)
        lut = p[:, 0] | (p[:, 1] << 8) | (p[:, 2] << 16) | np.uint32(0xFF000000)
        out[i] = lut[fb[i]]
    return out


@pytest.mark.parametrize("gamma", [1.0, 0.7, 1.35])
def test_view_palette_and_present_match_reference(assets_dir, sim_lib, oracle_available, gamma):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views
    specs = _specs(8)
    o = run_oracle.render(assets_dir, "maps/small.bsp", 320, 200, specs, cvars={"gamma": gamma}, want_palette=True)
    assert all(p is not None for p in o["palettes"])           # every view has colour shifts: the reference uploads
    r = RefCuda(assets_dir, 320, 200, libpath=sim_lib)
    r.load_world("maps/small.bsp")
    r.set_gamma(gamma)
    rds = views.build_refdefs(specs, 320, 200, model_resolver=r.model)
    for i in range(len(specs)):
        assert np.array_equal(r.view_palette(rds[i]), o["palettes"][i]), f"palette of view {i}"
    rgbx = r.present(rds, o["fb"])
    assert np.array_equal(rgbx, _expand(o["fb"], o["palettes"]))
    r.shutdown()


def test_palette_without_cshifts_is_base_palette_through_gamma(assets_dir, sim_lib):
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views, assets
    specs = rcutil.scene_views("small", 1, time=0.4)
    r = RefCuda(assets_dir, 320, 200, libpath=sim_lib)
    r.load_world("maps/small.bsp")
    rds = views.build_refdefs(specs, 320, 200, model_resolver=r.model)
    assert np.array_equal(r.view_palette(rds[0]), assets.make_palette().reshape(-1))
    r.shutdown()


@pytest.mark.gpu
def test_present_on_device_matches_reference(assets_dir, oracle_available):
    """RC_PresentViews through the CUDA kernel (k_present), odd plane size included."""
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views
    specs = _specs(12)
    r = RefCuda(assets_dir, 320, 200)
    r.load_world("maps/small.bsp")
    r.set_gamma(0.8)
    rds = views.build_refdefs(specs, 320, 200, model_resolver=r.model)
    fb, _, st = r.render(rds)
    assert st == 0
    pals = [r.view_palette(rds[i]) for i in range(len(specs))]
    if oracle_available:
        from oracle import run_oracle
        o = run_oracle.render(assets_dir, "maps/small.bsp", 320, 200, specs, cvars={"gamma": 0.8}, want_palette=True)
        assert np.array_equal(fb, o["fb"])
        for a, b in zip(pals, o["palettes"]):
            assert np.array_equal(a, b)
    assert np.array_equal(r.present(rds, fb), _expand(fb, pals))
    r.shutdown()


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(322, 201), (70, 51), (324, 202)])
def test_present_with_plane_sizes_that_are_not_multiples_of_four(assets_dir, w, h):
    """k_present's vector path needs 4 / 16-byte aligned planes; width * height that is not a multiple of 4 puts every
    view after the first off that alignment: the kernel must take its one-pixel path (misaligned-address fault)."""
    from tochris_b200.api import RefCuda
    from tochris_b200.scenes import views
    specs = _specs(5)
    r = RefCuda(assets_dir, w, h)
    r.load_world("maps/small.bsp")
    rds = views.build_refdefs(specs, w, h, model_resolver=r.model)
    fb, _, st = r.render(rds)
    assert st == 0
    pals = [r.view_palette(rds[i]) for i in range(len(specs))]
    assert np.array_equal(r.present(rds, fb), _expand(fb, pals))
    fb2, _, st = r.render(rds)                      # the context is still healthy
    assert st == 0 and np.array_equal(fb, fb2)
    r.shutdown()
