"""Generates the committed golden vectors from the oracle (the unmodified reference compiled
into oracle/_ref).  Run where /root/reference exists:  python tests/golden/make_golden.py
Each .npz holds the case description, sha256 of the full colour / z batches and the first
few frames verbatim."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rcutil  # noqa: E402
from oracle import run_oracle  # noqa: E402

CASES = [
    ("smoke_small_320x200", "small", 320, 200, 8, 0.4, 0.31, 2),
    ("box_c1_320x200", "box", 320, 200, 128, 1.0, 0.0, 2),
    ("full_320x240", "full", 320, 240, 32, 0.37, 0.173, 2),
    ("multi_c2_640x480", "multi", 640, 480, 64, 1.0, 0.0, 0),
    ("multi_c4_160x120", "multi", 160, 120, 256, 0.0, 0.01, 4),
    # bench.py extras: the first views of its 320x240 line, and BASELINE config C3 at full size (hashes only)
    ("multi_320x240", "multi", 320, 240, 64, 0.0, 0.01, 0),
    ("full_c3_1280x720", "full", 1280, 720, 32, 2.5, 0.5, 0, dict(entities=256, particles=8192)),
]


def main():
    base = tempfile.mkdtemp(prefix="rc_golden_")
    rcutil.make_assets(base)
    only = set(sys.argv[1:])
    for case in CASES:
        name, scene, w, h, n, t0, dt, keep = case[:8]
        extra = case[8] if len(case) > 8 else {}
        if only and name not in only:
            continue
        specs = rcutil.scene_views(scene, n, time=t0, time_step=dt, **extra) if scene != "box" else rcutil.scene_views(scene, n)
        o = run_oracle.render(base, f"maps/{scene}.bsp", w, h, specs)
        d = dict(scene=scene, w=w, h=h, n=n, t0=t0, dt=dt, fb_sha=rcutil.sha(o["fb"]), z_sha=rcutil.sha(o["z"]),
                 bsp_sha=rcutil.sha(np.frombuffer(open(os.path.join(base, "id1", "maps", scene + ".bsp"), "rb").read(), dtype=np.uint8)))
        d.update({k: int(v) for k, v in extra.items()})
        if keep:
            d["fb"] = o["fb"][:keep]; d["z"] = o["z"][:keep]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, d["fb_sha"][:16], d["z_sha"][:16])


if __name__ == "__main__":
    main()
