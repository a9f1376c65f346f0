"""SURVEY.md 8(f) rank 3, the .lit half of the asset path: Mod_LoadLighting with r_loadlits set (ref_soft/r_model.c:528-566)
replaces the BSP's light samples by max(r, g, b) of "<map>.lit".

The reference cannot run this path itself: the map file is loaded with COM_LoadTempFile (r_model.c:292, one temp
buffer at a time: common.c:1216 / Hunk_TempAlloc) and the .lit is loaded with COM_LoadTempFile again (r_model.c:538),
which frees the map, so every lump read after the lighting lump comes from recycled memory (the oracle build
segfaults right after "maps/x.lit loaded").  The
parity anchor is therefore the reference rendering a BSP whose lighting LUMP already holds max(r, g, b) -- the values
r_model.c:551-556 would have produced -- against ref_cuda rendering the original BSP + .lit with r_loadlits on."""
import os
import struct

import numpy as np
import pytest

import rcutil

LUMP_LIGHTING = 8      # common/bspfile.h


def write_lit(assets_dir, mapname, version=1, magic=b"QLIT", seed=7):
    """<map>.lit with random RGB samples and <map>_lit.bsp = the same map with max(r, g, b) in its lighting lump."""
    mdir = os.path.join(assets_dir, "id1", "maps")
    data = bytearray(open(os.path.join(mdir, mapname + ".bsp"), "rb").read())
    ofs, length = struct.unpack_from("<ii", data, 4 + 8 * LUMP_LIGHTING)
    rgb = np.random.default_rng(seed).integers(0, 256, size=(length, 3), dtype=np.uint8)
    with open(os.path.join(mdir, mapname + ".lit"), "wb") as f:
        f.write(magic + struct.pack("<i", version) + rgb.tobytes())
    data[ofs:ofs + length] = rgb.max(axis=1).tobytes()
    with open(os.path.join(mdir, mapname + "_lit.bsp"), "wb") as f:
        f.write(bytes(data))
    return length


@pytest.mark.parametrize("version,magic,used", [(1, b"QLIT", True), (2, b"QLIT", False), (1, b"XLIT", False)])
def test_lit_lighting(assets_dir, sim_lib, oracle_available, version, magic, used):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    _lit_case(assets_dir, sim_lib, version, magic, used)


@pytest.mark.gpu
@pytest.mark.parametrize("version,magic,used", [(1, b"QLIT", True), (2, b"QLIT", False)])
def test_lit_lighting_gpu(assets_dir, oracle_available, version, magic, used):
    """the same through libref_cuda.so: the loader is host code, the samples end up in the device's lightdata"""
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from tochris_b200 import api
    _lit_case(assets_dir, api.LIBPATH, version, magic, used)


def _lit_case(assets_dir, sim_lib, version, magic, used):
    from oracle import run_oracle
    assert write_lit(assets_dir, "small", version, magic) > 0
    mdir = os.path.join(assets_dir, "id1", "maps")
    try:
        specs = rcutil.scene_views("small", 6, time=0.4, time_step=0.31)
        plain = run_oracle.render(assets_dir, "maps/small.bsp", 320, 200, specs)
        want = run_oracle.render(assets_dir, "maps/small_lit.bsp", 320, 200, specs) if used else plain   # a bad .lit: the lump stays
        assert (want["fb"] != plain["fb"]).any() == used
        r = rcutil.render_rc(sim_lib, assets_dir, "maps/small.bsp", 320, 200, specs, cvars={"r_loadlits": 1.0})
        rcutil.assert_parity(want, r, specs, 320, 200)
        off = rcutil.render_rc(sim_lib, assets_dir, "maps/small.bsp", 320, 200, specs)                 # r_loadlits 0 ignores the file
        rcutil.assert_parity(plain, off, specs, 320, 200)
    finally:
        os.remove(os.path.join(mdir, "small.lit"))
        os.remove(os.path.join(mdir, "small_lit.bsp"))
