"""The step in front of the path, end to end on the GPU (GPU box only):  python tools/replay_session.py [frames] [w h]
A scripted play session (tochris_b200/scenes/session.py) is turned into frames by the producers of rc_client.c and played
by RC_TimeDemo -- through the pipelined host path, and with the frames kept on the device -- printing timedemo's line."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rcutil  # noqa: E402
from tochris_b200 import api, client  # noqa: E402
from tochris_b200.scenes import session  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (640, 480)
base = tempfile.mkdtemp(prefix="rc_replay_")
rcutil.make_assets(base, names=["full"])
r = api.RefCuda(base, w, h)
r.load_world("maps/full.bsp")
L = client.bind(r.lib)
sc, frames, cnt, _ = session.produce(r, L, rcutil.scene_views("full", 64, time=0.0), nframes=n, width=w, height=h)
fb = np.zeros((cnt, h, w), np.uint8)
for label, out in (("host frames", fb), ("device frames", None)):
    r.timedemo(frames, cnt, 64, out)                     # warm-up: surface cache, queues
    td = r.timedemo(frames, cnt, 64, out)
    print(f"{w}x{h} x {cnt} produced frames, {label}: {td['text'].strip()}  status {api.status_names(td['status'])}")
ents = [frames[i].num_entities for i in range(cnt)]
parts = [frames[i].num_particles for i in range(cnt)]
print(f"per frame: entities {min(ents)}-{max(ents)}, particles {min(parts)}-{max(parts)}")
r.scene_free(sc)
r.shutdown()
