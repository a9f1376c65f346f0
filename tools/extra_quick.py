"""One `extra` sub-record of bench.py by itself (GPU box only):  python tools/extra_quick.py c3|c4|v320x240
(select a library variant with RC_LIBPATH; the record is verified against the golden frames like in the bench line)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "c3"
args = argparse.Namespace(c4_views=65536, v320_views=1024, gather="peer-copy")
r = bench.run_extra(kind, args, 0, 1, 0)
keep = {k: r.get(k) for k in ("value", "unit", "ms_per_step", "verified", "status", "kernel_ms_per_step", "entity_stage")}
print(os.environ.get("RC_LIBPATH", "base"), json.dumps(keep))
