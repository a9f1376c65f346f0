"""Extracts the alias-model shading table the renderer must share with the reference.

ref_soft/anorm_dots.h holds r_avertexnormal_dots[16][256] (r_alias.c:29-31): shade
factors per (quantised yaw, vertex-normal index), printed with two decimals by id's
tooling.  They cannot be regenerated bit-exactly from a formula, and pixel parity needs
the identical numbers, so the VALUES (not the source text) are stored as a 16 KiB binary
float32 table, tochris_b200/data/anorm_dots.f32.  Run where /root/reference exists."""
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

txt = open(os.path.join(REF, "ref_soft", "anorm_dots.h")).read()
txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
txt = re.sub(r"//.*", "", txt)
vals = np.array([float(x) for x in re.findall(r"-?\d+\.\d+", txt)], dtype=np.float32)
assert vals.size == 16 * 256, vals.size
out = os.path.join(ROOT, "tochris_b200", "data", "anorm_dots.f32")
vals.tofile(out)
print("wrote", out, vals.size, "floats; checksum", float(vals.astype(np.float64).sum()))

# client/anorms.h: the 162 vertex normals CL_EntityParticles (cl_effects.c:88-138) places its particles on.  Numbers
# printed with six decimals by id's tooling (the vertices of a subdivided icosahedron, not reproducible bit for bit from
# a formula): the VALUES go into tochris_b200/data/anorms.f32 (162 x 3 float32).
txt = open(os.path.join(REF, "client", "anorms.h")).read()
txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
txt = re.sub(r"//.*", "", txt)
vals = np.array([float(x) for x in re.findall(r"-?\d+\.\d+", txt)], dtype=np.float32)
assert vals.size == 162 * 3, vals.size
out = os.path.join(ROOT, "tochris_b200", "data", "anorms.f32")
vals.tofile(out)
print("wrote", out, vals.size, "floats; checksum", float(vals.astype(np.float64).sum()))
