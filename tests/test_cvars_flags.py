"""The pixel-affecting configuration surface against the oracle, bit for bit (tests/cvar_cases.py): r_fastsky / r_skycolor,
r_fullbright, r_ambient, d_mipcap, d_mipscale, r_drawviewmodel, r_drawentities, r_waterwarp, r_clearcolor, r_lerpmodels,
RF_WEAPONMODEL, RF_DEPTHSCALE, RF_VIEWERMODEL, RF_MINLIGHT, RF_FULLBRIGHT, RF_CULLVIS and RDF_NOWORLDMODEL.
The host-simulator leg runs here on the CPU; the GPU leg is the parity test proper (libref_cuda.so through the C-ABI)."""
import numpy as np
import pytest

import cvar_cases
import rcutil
from tochris_b200 import api

IDS = [c[0] for c in cvar_cases.CVAR_CASES]


def _run(case, lib, assets_dir, sim):
    from oracle import run_oracle
    scene, w, h, specs, cvars, okw = cvar_cases.build(case, sim)
    o = run_oracle.render(assets_dir, f"maps/{scene}.bsp", w, h, specs, cvars=cvars, **okw)
    r = rcutil.render_rc(lib, assets_dir, f"maps/{scene}.bsp", w, h, specs, cvars=cvars)
    if not sim:
        assert r["stats"]["kernel_launches"] > 0
    if okw.get("clear_each"):
        # RDF_NOWORLDMODEL: nothing but the entities is drawn, the z-buffer starts at 0xFFFF (d_modech.c:103-106)
        assert r["status"] & ~rcutil.RC_STAT_ALIAS_RANGE == 0, api.status_names(r["status"])
        for i, sp in enumerate(specs):
            assert np.array_equal(o["fb"][i], r["fb"][i]), (i, int((o["fb"][i] != r["fb"][i]).sum()))
            assert np.array_equal(o["z"][i], r["z"][i]), (i, int((o["z"][i] != r["z"][i]).sum()))
        assert any(r["fb"][i].any() for i in range(len(specs))), "no entity pixel was drawn"
    else:
        rcutil.assert_parity(o, r, specs, w, h, allow_alias_range=False)
    return o, r, specs


@pytest.mark.parametrize("case", cvar_cases.CVAR_CASES, ids=IDS)
def test_cvars_flags_sim(assets_dir, sim_lib, oracle_available, case):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    _run(case, sim_lib, assets_dir, True)


@pytest.mark.gpu
@pytest.mark.parametrize("case", cvar_cases.CVAR_CASES, ids=IDS)
def test_cvars_flags_gpu(assets_dir, oracle_available, case):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    _run(case, api.LIBPATH, assets_dir, False)


def test_cases_change_pixels(assets_dir, oracle_available):
    """every cvar case must differ from the default rendering of the same views (a case that changes nothing tests nothing)"""
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle
    for case in cvar_cases.CVAR_CASES:
        scene, w, h, specs, cvars, okw = cvar_cases.build(case, True)
        if not cvars or case[0] in ("r_clearcolor_default", "RDF_NOWORLDMODEL_particles", "RF_MINLIGHT_FULLBRIGHT", "r_ambient_negative"):  # negative r_ambient is clamped to 0 (r_misc.c:417)
            continue
        a = run_oracle.render(assets_dir, f"maps/{scene}.bsp", w, h, specs, cvars=cvars, **okw)
        b = run_oracle.render(assets_dir, f"maps/{scene}.bsp", w, h, specs, **okw)
        assert (a["fb"] != b["fb"]).any(), case[0]
