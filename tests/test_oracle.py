"""The oracle (unmodified reference compiled into oracle/_ref) against the committed golden
vectors, and the struct layouts the C-ABI mirrors."""
import ctypes as C
import os

import numpy as np
import pytest

import rcutil
from tochris_b200 import refdef


def test_struct_sizes_match_reference(oracle_available, assets_dir):
    """client/ref.h structs as the reference compiler lays them out (136/80/16/16/68 B)."""
    assert C.sizeof(refdef.refdef_t) == 136
    assert C.sizeof(refdef.entity_t) == 80
    assert C.sizeof(refdef.particle_t) == 16
    assert C.sizeof(refdef.dlight_t) == 16
    assert C.sizeof(refdef.lightstyle_t) == 68
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle  # noqa: F401  (sizes are checked inside a fresh process)
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_sizes, args=(q, assets_dir))
    p.start(); sizes = q.get(timeout=120); p.join()
    assert sizes[:5] == [136, 80, 16, 16, 68]
    assert sizes[5:9] == [56, 88, 24, 32]          # edge_t, surf_t, espan_t, finalvert_t


def _sizes(q, base):
    from oracle.refsoft import RefSoft
    q.put(RefSoft(base, 320, 200).struct_sizes())


@pytest.mark.parametrize("name", ["smoke_small_320x200", "full_320x240"])
def test_oracle_reproduces_golden(oracle_available, assets_dir, name):
    if not oracle_available:
        pytest.skip("oracle/_ref not built")
    from oracle import run_oracle
    g = np.load(os.path.join(rcutil.ROOT, "tests", "golden", name + ".npz"))
    scene = str(g["scene"]); w, h, n = int(g["w"]), int(g["h"]), int(g["n"])
    bsp = np.frombuffer(open(os.path.join(assets_dir, "id1", "maps", scene + ".bsp"), "rb").read(), dtype=np.uint8)
    assert rcutil.sha(bsp) == str(g["bsp_sha"]), "the procedural BSP generator changed: regenerate the goldens"
    specs = rcutil.scene_views(scene, n, time=float(g["t0"]), time_step=float(g["dt"]))
    o = run_oracle.render(assets_dir, f"maps/{scene}.bsp", w, h, specs)
    assert rcutil.sha(o["fb"]) == str(g["fb_sha"])
    assert rcutil.sha(o["z"]) == str(g["z_sha"])
    k = g["fb"].shape[0]
    assert (o["fb"][:k] == g["fb"]).all() and (o["z"][:k] == g["z"]).all()
