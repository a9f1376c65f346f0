"""Parity cases for the pixel-affecting configuration surface: the cvars of ref_soft/r_main.c:108-134 and
d_init.c:26-28, the entity flags of common/quakedef.h:228-234 and RDF_NOWORLDMODEL.  Shared by the host-simulator test
(-m "not gpu") and the GPU test (-m gpu); both compare with the oracle bit for bit."""
import math

import rcutil
from tochris_b200.refdef import (RF_VIEWERMODEL, RF_WEAPONMODEL, RF_DEPTHSCALE, RF_MINLIGHT, RF_FULLBRIGHT, RF_CULLVIS,
                                 RDF_NOWORLDMODEL)
from tochris_b200.scenes import views


def _weapon(specs, every=1, flags=RF_WEAPONMODEL):
    """a view model in front of the camera, the way V_CalcRefdef places cl.viewent (client/cl_view.c:760-800): at the view
    origin, facing along the view angles -- first in the entity list; RF_WEAPONMODEL (r_alias.c:746) changes its projection"""
    for i, sp in enumerate(specs):
        if i % every:
            continue
        yaw, pitch = math.radians(sp.angles[1]), math.radians(sp.angles[0])
        fwd = (math.cos(yaw) * math.cos(pitch), math.sin(yaw) * math.cos(pitch), -math.sin(pitch))
        org = tuple(sp.origin[k] + fwd[k] * 22.0 - (6.0 if k == 2 else 0.0) for k in range(3))
        e = views.EntitySpec("progs/blob.mdl", org, (-sp.angles[0], sp.angles[1], 0.0), frame=i % 2, oldframe=(i + 1) % 2,
                             framelerp=0.25 * (i % 4), flags=flags)
        sp.entities = [e] + list(sp.entities)


def _flag_entities(specs, flag, every=2):
    for sp in specs:
        for j, e in enumerate(sp.entities):
            if j % every == 0:
                e.flags |= flag


def _outside(specs, info):
    """cameras outside the sealed map looking back at it: the view leaf is the solid leaf 0, Mod_LeafPVS answers
    'everything' (r_model.c:189), the rooms are seen from behind (back faces culled) and whatever no surface covers is the
    background surface in r_clearcolor (d_edge.c:236-245)"""
    xs = [s[0] for s in info["spawn"]]; ys = [s[1] for s in info["spawn"]]
    cx, cy = 0.5 * (min(xs) + max(xs)), 0.5 * (min(ys) + max(ys))
    rad = 0.5 * max(max(xs) - min(xs), max(ys) - min(ys)) + info.get("room", 400) * 1.5
    for i, sp in enumerate(specs):
        a = 2.0 * math.pi * views.halton(i + 1, 2)
        ox, oy, oz = cx + rad * math.cos(a), cy + rad * math.sin(a), 40.0 + 300.0 * views.halton(i + 1, 3)
        sp.origin = (ox, oy, oz)
        yaw = math.degrees(math.atan2(cy - oy, cx - ox)) + (views.halton(i + 1, 5) - 0.5) * 50.0
        sp.angles = ((views.halton(i + 1, 7)) * 40.0 - 5.0, yaw, 0.0)


def _noworld(specs):
    """the player-setup menu's view (client/menu.c:789-823): no world, one or more alias entities in front of a camera at
    the origin"""
    for i, sp in enumerate(specs):
        sp.flags |= RDF_NOWORLDMODEL
        sp.origin = (0.0, 0.0, 0.0)
        sp.angles = (0.0, 0.0, 0.0)
        sp.fov_x = 40.0 if i % 2 == 0 else 90.0
        sp.dlights = []
        ents = []
        for j in range(1 + i % 3):
            ents.append(views.EntitySpec("progs/ball.mdl" if (i + j) % 2 else "progs/blob.mdl",
                                         (70.0 + 25.0 * j, -14.0 + 22.0 * j, -8.0 + 3.0 * j), (0.0, 17.0 * i + 40.0 * j, 0.0),
                                         frame=j % 2, oldframe=(j + 1) % 2, framelerp=0.5 * (i % 2),
                                         flags=(RF_FULLBRIGHT if j == 1 else 0)))
        sp.entities = ents
        if i % 4 == 3:
            sp.rect = (16, 8, 96, 96)            # the menu draws into a small rectangle


# id, scene, w, h, n (gpu), n (host simulator), t0, dt, scene_views extras, cvars, mutation
CVAR_CASES = [
    ("r_fastsky", "full", 320, 240, 48, 10, 0.37, 0.173, {}, dict(r_fastsky=1.0, r_skycolor=37.0), None),
    ("r_fullbright", "multi", 320, 240, 32, 6, 0.2, 0.05, dict(dlights=2, entities=4), dict(r_fullbright=1.0), None),
    ("r_ambient", "full", 320, 240, 32, 8, 0.3, 0.11, dict(dlights=2, entities=6, doors=2), dict(r_ambient=40.0), None),
    ("r_ambient_negative", "multi", 320, 240, 16, 4, 0.3, 0.11, dict(entities=3), dict(r_ambient=-25.0), None),
    ("d_mipcap1", "multi", 640, 480, 24, 3, 0.1, 0.07, {}, dict(d_mipcap=1.0), None),
    ("d_mipcap2", "full", 320, 240, 32, 6, 0.1, 0.07, dict(doors=2), dict(d_mipcap=2.0), None),
    ("d_mipcap3", "multi", 320, 240, 32, 6, 0.1, 0.07, {}, dict(d_mipcap=3.0), None),
    ("d_mipscale_half", "multi", 640, 480, 24, 3, 0.5, 0.07, {}, dict(d_mipscale=0.5), None),
    ("d_mipscale_two", "full", 640, 480, 24, 3, 0.5, 0.07, dict(doors=2), dict(d_mipscale=2.0), None),
    ("d_mipcap_mipscale", "multi", 400, 300, 24, 4, 0.5, 0.07, {}, dict(d_mipcap=1.0, d_mipscale=3.0), None),
    ("r_drawviewmodel0", "multi", 320, 240, 24, 6, 0.4, 0.09, dict(entities=3), dict(r_drawviewmodel=0.0), "weapon"),
    ("RF_WEAPONMODEL", "multi", 320, 240, 32, 8, 0.4, 0.09, dict(entities=3, particles=40), {}, "weapon"),
    ("RF_WEAPONMODEL_wide", "full", 640, 360, 16, 4, 0.4, 0.09, dict(underwater_every=3), {}, "weapon"),
    ("RF_WEAPONMODEL_DEPTHSCALE", "multi", 320, 240, 24, 6, 0.6, 0.09, dict(entities=2), {}, "weapon_depthscale"),
    ("RF_DEPTHSCALE", "multi", 320, 240, 32, 8, 0.6, 0.09, dict(entities=8), {}, "depthscale"),
    ("RF_VIEWERMODEL", "multi", 320, 240, 16, 4, 0.6, 0.09, dict(entities=4), {}, "viewermodel"),
    ("RF_MINLIGHT_FULLBRIGHT", "multi", 320, 240, 16, 4, 0.6, 0.09, dict(entities=6), dict(r_ambient=0.0), "minlight"),
    ("RF_CULLVIS", "full", 320, 240, 48, 10, 0.8, 0.13, dict(entities=12, sprites=6), {}, "cullvis"),
    ("r_drawentities0", "full", 320, 240, 24, 6, 0.8, 0.13, dict(entities=6, sprites=4, doors=2, particles=60), dict(r_drawentities=0.0), None),
    ("r_waterwarp0", "full", 320, 240, 24, 6, 0.8, 0.13, dict(underwater_every=2, entities=3, particles=30), dict(r_waterwarp=0.0), None),
    ("r_waterwarp0_big", "full", 640, 480, 8, 2, 0.8, 0.13, dict(underwater_every=1), dict(r_waterwarp=0.0), None),
    ("r_clearcolor", "multi", 320, 240, 32, 8, 0.0, 0.0, {}, dict(r_clearcolor=251.0), "outside"),
    ("r_clearcolor_default", "full", 320, 240, 16, 4, 0.0, 0.0, dict(entities=4), {}, "outside"),
    ("r_lerpmodels0", "multi", 320, 240, 32, 8, 0.6, 0.09, dict(entities=8), dict(r_lerpmodels=0.0), None),
    ("r_draworder", "multi", 320, 240, 32, 8, 0.3, 0.11, {}, dict(r_draworder=1.0), None),        # the farthest surface wins (R_GenerateSpansBackward)
    ("r_draworder_full", "full", 400, 300, 24, 6, 0.37, 0.173, dict(doors=2, entities=4, particles=40, dlights=1), dict(r_draworder=1.0), None),
    ("RDF_NOWORLDMODEL", "multi", 320, 200, 24, 8, 0.6, 0.09, {}, {}, "noworld"),
    ("RDF_NOWORLDMODEL_particles", "multi", 320, 200, 8, 4, 0.6, 0.09, dict(particles=50), dict(r_fullbright=0.0), "noworld"),
]


def build(case, sim: bool):
    cid, scene, w, h, n_gpu, n_sim, t0, dt, extra, cvars, mut = case
    n = n_sim if sim else n_gpu
    specs = rcutil.scene_views(scene, n, time=t0, time_step=dt, **extra)
    okw = {}
    if mut == "weapon":
        _weapon(specs)
    elif mut == "weapon_depthscale":
        _weapon(specs, flags=RF_WEAPONMODEL | RF_DEPTHSCALE)
    elif mut == "depthscale":
        _flag_entities(specs, RF_DEPTHSCALE)
    elif mut == "viewermodel":
        _flag_entities(specs, RF_VIEWERMODEL)
    elif mut == "minlight":
        _flag_entities(specs, RF_MINLIGHT, 2)
        _flag_entities(specs, RF_FULLBRIGHT, 3)
    elif mut == "cullvis":
        _flag_entities(specs, RF_CULLVIS, 1)
    elif mut == "outside":
        _outside(specs, rcutil.scene_info(scene))
    elif mut == "noworld":
        _noworld(specs)
        okw["clear_each"] = True
    return scene, w, h, specs, dict(cvars), okw
