"""Procedural BSP29 compiler for axis-aligned voxel worlds.

Emits the lumps ref_soft's loader reads (ref_soft/r_model.c:1084-1096, layouts in
common/bspfile.h:56-225): PLANES, TEXTURES, VERTEXES, VISIBILITY, NODES, TEXINFO, FACES,
LIGHTING, LEAFS, MARKSURFACES, EDGES, SURFEDGES, MODELS.  ENTITIES/CLIPNODES are not read
by the renderer and are left empty.

World model: a rectilinear grid (xs, ys, zs) of cells, each SOLID / EMPTY / WATER.  The
tree is a kd-style BSP over grid planes whose leaves are content-uniform boxes; every
solid/open (and empty/water) boundary rectangle becomes faces stored on the node whose
plane separates the two cells, exactly as qbsp stores faces on nodes.  Faces are merged
greedily and subdivided to <= 240 units so extents stay <= 256 (r_model.c:791).
Brush submodels (doors) are axis-aligned boxes appended as models 1..n.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from . import assets

SOLID, EMPTY, WATER = 0, 1, 2
CONTENTS = {SOLID: -2, EMPTY: -1, WATER: -3}
TEX_SPECIAL = 1
SUBDIVIDE = 240

# qbsp TextureAxisFromPlane for axial planes: (s axis, t axis) per normal axis
_TEXAXES = {
    0: ((0.0, 1.0, 0.0), (0.0, 0.0, -1.0)),
    1: ((1.0, 0.0, 0.0), (0.0, 0.0, -1.0)),
    2: ((1.0, 0.0, 0.0), (0.0, -1.0, 0.0)),
}
_OTHER = {0: (1, 2), 1: (0, 2), 2: (0, 1)}


@dataclass
class Face:
    axis: int
    coord: float
    side: int               # 0: normal = +axis, 1: normal = -axis (SURF_PLANEBACK)
    rect: Tuple[float, float, float, float]  # b0, b1, c0, c1 in world units (other two axes)
    tex: str
    texinfo: int = -1
    styles: Tuple[int, int, int, int] = (0, 255, 255, 255)
    verts: Optional[List[Tuple[float, float, float]]] = None  # explicit polygon (submodels)


@dataclass
class _Node:
    plane: int
    children: List[int]
    mins: Tuple[int, int, int]
    maxs: Tuple[int, int, int]
    firstface: int = 0
    numfaces: int = 0


@dataclass
class _Leaf:
    contents: int
    mins: Tuple[int, int, int]
    maxs: Tuple[int, int, int]
    box: Tuple[int, int, int, int, int, int] = (0, 0, 0, 0, 0, 0)  # index-space box
    marks: List[int] = field(default_factory=list)
    vis: Optional[np.ndarray] = None


@dataclass
class BrushModel:
    """Axis-aligned box brush entity (door/plat), in model space around `origin`."""
    mins: Tuple[float, float, float]
    maxs: Tuple[float, float, float]
    tex: str = "door"


class BspBuilder:
    def __init__(self, xs, ys, zs, content: np.ndarray, seed: int = assets.SEED,
                 tex_for_face: Optional[Callable[[Face], str]] = None,
                 multi_style_every: int = 7, texinfo_variants: bool = True):
        self.g = [np.asarray(xs, dtype=np.int64), np.asarray(ys, dtype=np.int64), np.asarray(zs, dtype=np.int64)]
        self.content = content
        assert content.shape == tuple(len(a) - 1 for a in self.g)
        # seal the world
        assert (content[0] == SOLID).all() and (content[-1] == SOLID).all()
        assert (content[:, 0] == SOLID).all() and (content[:, -1] == SOLID).all()
        assert (content[:, :, 0] == SOLID).all() and (content[:, :, -1] == SOLID).all()
        self.rng = np.random.RandomState(seed)
        self.tex_for_face = tex_for_face or self._default_tex
        self.multi_style_every = multi_style_every
        self.texinfo_variants = texinfo_variants
        self.planes: List[Tuple[int, float]] = []
        self.plane_idx: Dict[Tuple[int, float], int] = {}
        self.nodes: List[_Node] = []
        self.leaves: List[_Leaf] = [_Leaf(CONTENTS[SOLID], (0, 0, 0), (0, 0, 0))]
        self.faces: List[Face] = []
        self.face_plane: List[int] = []
        self.textures: Dict[str, np.ndarray] = {}
        self.submodels: List[BrushModel] = []
        self.sub_records: List[dict] = []

    # ---------------------------------------------------------------- textures
    def add_texture(self, name: str, pixels: np.ndarray):
        self.textures[name] = pixels

    @staticmethod
    def _default_tex(f: Face) -> str:
        if f.axis == 2:
            return "floor" if f.side == 0 else "ceil"
        return "wall_a" if f.axis == 0 else "wall_b"

    # ------------------------------------------------------------------- tree
    def _plane(self, axis: int, coord: float) -> int:
        k = (axis, float(coord))
        if k not in self.plane_idx:
            self.plane_idx[k] = len(self.planes)
            self.planes.append(k)
        return self.plane_idx[k]

    def _bbox(self, box):
        x0, x1, y0, y1, z0, z1 = box
        return ((int(self.g[0][x0]), int(self.g[1][y0]), int(self.g[2][z0])),
                (int(self.g[0][x1]), int(self.g[1][y1]), int(self.g[2][z1])))

    def _faces_on_plane(self, axis: int, k: int, box) -> List[Face]:
        """Boundary rectangles on grid plane `k` of `axis`, limited to `box`."""
        lo = [box[0], box[2], box[4]]
        hi = [box[1], box[3], box[5]]
        b, c = _OTHER[axis]
        sl_front = [slice(lo[i], hi[i]) for i in range(3)]
        sl_back = list(sl_front)
        sl_front[axis] = k
        sl_back[axis] = k - 1
        front = self.content[tuple(sl_front)]
        back = self.content[tuple(sl_back)]
        out: List[Face] = []
        coord = float(self.g[axis][k])

        def emit(mask: np.ndarray, side: int, kind: str):
            mask = mask.copy()
            nb, nc = mask.shape
            for i in range(nb):
                for j in range(nc):
                    if not mask[i, j]:
                        continue
                    # greedy rectangle: extend in c then in b
                    j1 = j
                    while j1 + 1 < nc and mask[i, j1 + 1]:
                        j1 += 1
                    i1 = i
                    while i1 + 1 < nb and mask[i1 + 1, j:j1 + 1].all():
                        i1 += 1
                    mask[i:i1 + 1, j:j1 + 1] = False
                    b0 = float(self.g[b][lo[b] + i]); b1 = float(self.g[b][lo[b] + i1 + 1])
                    c0 = float(self.g[c][lo[c] + j]); c1 = float(self.g[c][lo[c] + j1 + 1])
                    special = kind != "wall"
                    bs = [b0, b1] if special else _cuts(b0, b1)
                    cs = [c0, c1] if special else _cuts(c0, c1)
                    for bi in range(len(bs) - 1):
                        for ci in range(len(cs) - 1):
                            f = Face(axis, coord, side, (bs[bi], bs[bi + 1], cs[ci], cs[ci + 1]), "")
                            f.tex = "*water" if kind == "water" else self.tex_for_face(f)
                            out.append(f)

        # solid boundary: face looks into the open cell
        emit((front != SOLID) & (back == SOLID), 0, "wall")
        emit((front == SOLID) & (back != SOLID), 1, "wall")
        # liquid surface: one face per side
        liquid = ((front == WATER) & (back == EMPTY)) | ((front == EMPTY) & (back == WATER))
        emit(liquid, 0, "water")
        emit(liquid, 1, "water")
        return out

    def _build(self, box) -> int:
        """Returns child reference: >=0 node index, <0 -(leaf+1)."""
        x0, x1, y0, y1, z0, z1 = box
        sub = self.content[x0:x1, y0:y1, z0:z1]
        first = sub.flat[0]
        if (sub == first).all():
            if first == SOLID:
                return -1
            mins, maxs = self._bbox(box)
            self.leaves.append(_Leaf(CONTENTS[int(first)], mins, maxs, box))
            return -(len(self.leaves) - 1) - 1
        # choose the split: plane carrying the most boundary area, most central on ties
        best = None
        lo = [x0, y0, z0]; hi = [x1, y1, z1]
        for axis in range(3):
            for k in range(lo[axis] + 1, hi[axis]):
                sf = [slice(lo[i], hi[i]) for i in range(3)]; sb = list(sf)
                sf[axis] = k; sb[axis] = k - 1
                diff = int((self.content[tuple(sf)] != self.content[tuple(sb)]).sum())
                if diff == 0:
                    continue
                centred = -abs((k - lo[axis]) - (hi[axis] - k))
                score = (diff, centred, -axis)
                if best is None or score > best[0]:
                    best = (score, axis, k)
        assert best is not None
        _, axis, k = best
        node_index = len(self.nodes)
        mins, maxs = self._bbox(box)
        node = _Node(self._plane(axis, float(self.g[axis][k])), [0, 0], mins, maxs)
        self.nodes.append(node)
        fs = self._faces_on_plane(axis, k, box)
        node.firstface = len(self.faces)
        node.numfaces = len(fs)
        for f in fs:
            self.faces.append(f)
            self.face_plane.append(node.plane)
        front = list(box); back = list(box)
        front[axis * 2] = k      # higher coordinates = positive side = children[0]
        back[axis * 2 + 1] = k
        node.children[0] = self._build(tuple(front))
        node.children[1] = self._build(tuple(back))
        return node_index

    # ------------------------------------------------------------ marksurfaces
    def _mark(self):
        for li, leaf in enumerate(self.leaves):
            if li == 0:
                continue
            lo = [leaf.box[0], leaf.box[2], leaf.box[4]]
            hi = [leaf.box[1], leaf.box[3], leaf.box[5]]
            for fi, f in enumerate(self.faces[:self.num_world_faces]):
                a = f.axis
                b, c = _OTHER[a]
                lo_c = float(self.g[a][lo[a]]); hi_c = float(self.g[a][hi[a]])
                if f.side == 0 and f.coord != lo_c:
                    continue
                if f.side == 1 and f.coord != hi_c:
                    continue
                b0, b1, c0, c1 = f.rect
                if b1 <= self.g[b][lo[b]] or b0 >= self.g[b][hi[b]]:
                    continue
                if c1 <= self.g[c][lo[c]] or c0 >= self.g[c][hi[c]]:
                    continue
                leaf.marks.append(fi)

    # -------------------------------------------------------------------- PVS
    def _pvs(self):
        """Conservative PVS: 2D line of sight between leaf footprints through columns that
        are open at any height (superset of true 3D visibility)."""
        open2d = (self.content != SOLID).any(axis=2)
        xs = self.g[0].astype(np.float64); ys = self.g[1].astype(np.float64)
        n = len(self.leaves) - 1
        pts = []
        for leaf in self.leaves[1:]:
            x0, x1 = xs[leaf.box[0]], xs[leaf.box[1]]
            y0, y1 = ys[leaf.box[2]], ys[leaf.box[3]]
            e = 1.0
            pts.append([((x0 + x1) / 2, (y0 + y1) / 2), (x0 + e, y0 + e), (x1 - e, y0 + e),
                        (x0 + e, y1 - e), (x1 - e, y1 - e)])
        P = np.array(pts)  # n,5,2
        vis = np.eye(n, dtype=bool)
        ii, jj = np.triu_indices(n, 1)
        chunk = 20000
        for s in range(0, len(ii), chunk):
            a = ii[s:s + chunk]; b = jj[s:s + chunk]
            seen = np.zeros(len(a), dtype=bool)
            for pa in range(5):
                for pb in range(5):
                    todo = ~seen
                    if not todo.any():
                        break
                    p = P[a[todo], pa]; q = P[b[todo], pb]
                    seen[np.flatnonzero(todo)] = _segments_clear(p, q, xs, ys, open2d)
            vis[a[seen], b[seen]] = True
            vis[b[seen], a[seen]] = True
        for i, leaf in enumerate(self.leaves[1:]):
            leaf.vis = vis[i]

    # ----------------------------------------------------------------- submodels
    def add_brush_model(self, bm: BrushModel):
        self.submodels.append(bm)

    def _emit_submodel(self, bm: BrushModel):
        """A box brush: 6 outward faces on a chain of 6 nodes; inside is solid."""
        firstface = len(self.faces)
        head = len(self.nodes)
        mn, mx = bm.mins, bm.maxs
        specs = []
        for axis in range(3):
            specs.append((axis, mx[axis], 0))   # +axis face at max
            specs.append((axis, mn[axis], 1))   # -axis face at min
        empty_leaf = len(self.leaves)
        self.leaves.append(_Leaf(CONTENTS[EMPTY], tuple(int(v) for v in mn), tuple(int(v) for v in mx)))
        for n_i, (axis, coord, side) in enumerate(specs):
            b, c = _OTHER[axis]
            f = Face(axis, float(coord), side, (mn[b], mx[b], mn[c], mx[c]), bm.tex)
            pl = self._plane(axis, float(coord))
            node = _Node(pl, [0, 0], tuple(int(v) - 1 for v in mn), tuple(int(v) + 1 for v in mx),
                         firstface=len(self.faces), numfaces=1)
            self.faces.append(f)
            self.face_plane.append(pl)
            nxt = head + n_i + 1 if n_i < 5 else -1      # solid inside at the end of the chain
            outside = -(empty_leaf) - 1
            # side 0 (face at max, normal +): front of plane is outside
            node.children = [outside, nxt] if side == 0 else [nxt, outside]
            self.nodes.append(node)
        self.sub_records.append(dict(mins=mn, maxs=mx, headnode=head, firstface=firstface, numfaces=6))

    # ------------------------------------------------------------------ output
    def build(self) -> bytes:
        nx, ny, nz = self.content.shape
        root = self._build((0, nx, 0, ny, 0, nz))
        assert root == 0
        self.num_world_faces = len(self.faces)
        self.num_world_leaves = len(self.leaves) - 1
        self._mark()
        self._pvs()
        for bm in self.submodels:
            self._emit_submodel(bm)
        return self._serialise()

    def _texinfo_for(self, f: Face, table: Dict, tex_index: Dict[str, int]) -> int:
        s_ax, t_ax = _TEXAXES[f.axis]
        special = f.tex.startswith("*") or f.tex.startswith("sky")
        scale, so, to = 1.0, 0.0, 0.0
        if self.texinfo_variants and not special:
            h = (len(table) * 2654435761 + int(f.coord) * 40503 + f.axis) & 0xFFFF
            if h % 5 == 1:
                scale = 0.5      # vec length 0.5 -> mipadjust 2 (r_model.c:751-758)
            elif h % 5 == 2:
                so, to = 5.0, 37.0
            elif h % 11 == 3:
                scale = 2.0
        key = (f.axis, f.tex, scale, so, to)
        if key not in table:
            vecs = [s_ax[0] * scale, s_ax[1] * scale, s_ax[2] * scale, so,
                    t_ax[0] * scale, t_ax[1] * scale, t_ax[2] * scale, to]
            table[key] = (len(table), vecs, tex_index[f.tex], TEX_SPECIAL if special else 0)
        return table[key][0]

    def _face_verts(self, f: Face):
        if f.verts is not None:
            return f.verts
        b, c = _OTHER[f.axis]
        b0, b1, c0, c1 = f.rect
        ring = [(b0, c0), (b1, c0), (b1, c1), (b0, c1)]
        pts = []
        for (pb, pc) in ring:
            p = [0.0, 0.0, 0.0]
            p[f.axis] = f.coord; p[b] = pb; p[c] = pc
            pts.append(tuple(p))
        # ring is counter-clockwise about (e_b x e_c).  Faces must be CLOCKWISE seen from
        # the front (r_rast.c:323-346: edges running down the screen are trailing edges).
        ebxec = {0: 1.0, 1: -1.0, 2: 1.0}[f.axis]
        normal_sign = 1.0 if f.side == 0 else -1.0
        if ebxec * normal_sign > 0:
            pts.reverse()
        return pts

    def _serialise(self) -> bytes:
        # textures
        names = []
        for f in self.faces:
            if f.tex not in names:
                names.append(f.tex)
        for extra in self.textures:
            if extra not in names:
                names.append(extra)
        tex_index = {n: i for i, n in enumerate(names)}
        miptex = []
        for i, n in enumerate(names):
            if n in self.textures:
                px = self.textures[n]
            elif n.startswith("sky"):
                px = assets.make_sky_pixels(assets.SEED + i)
            elif n.startswith("*"):
                px = assets.make_texture_pixels(assets.SEED + i, 64, 64, "water")
            else:
                style = ["brick", "tile", "noise", "stripes"][i % 4]
                size = 128 if i % 3 == 0 else 64
                px = assets.make_texture_pixels(assets.SEED + i, size, size, style)
            miptex.append(assets.make_miptex(n, px))
        ofs = 4 + 4 * len(miptex)
        dataofs = []
        for m in miptex:
            dataofs.append(ofs)
            ofs += len(m)
        lump_tex = struct.pack("<i", len(miptex)) + struct.pack("<%di" % len(miptex), *dataofs) + b"".join(miptex)

        # texinfo, vertices, edges, faces, lighting
        ti_table: Dict = {}
        verts: List[Tuple[float, float, float]] = []
        vert_idx: Dict[Tuple[float, float, float], int] = {}
        edges: List[Tuple[int, int]] = [(0, 0)]
        edge_idx: Dict[Tuple[int, int], int] = {}
        edge_used_rev: Dict[int, bool] = {}
        surfedges: List[int] = []
        dfaces = []
        light = bytearray()
        for fi, f in enumerate(self.faces):
            f.texinfo = self._texinfo_for(f, ti_table, tex_index)
            pts = self._face_verts(f)
            vi = []
            for p in pts:
                if p not in vert_idx:
                    vert_idx[p] = len(verts)
                    verts.append(p)
                vi.append(vert_idx[p])
            firstedge = len(surfedges)
            for k in range(len(vi)):
                a, b_ = vi[k], vi[(k + 1) % len(vi)]
                if (b_, a) in edge_idx and not edge_used_rev.get(edge_idx[(b_, a)], False):
                    e = edge_idx[(b_, a)]
                    edge_used_rev[e] = True
                    surfedges.append(-e)
                else:
                    e = len(edges)
                    edges.append((a, b_))
                    if (a, b_) not in edge_idx:
                        edge_idx[(a, b_)] = e
                    surfedges.append(e)
            special = f.tex.startswith("*") or f.tex.startswith("sky")
            if special:
                styles = (255, 255, 255, 255); lightofs = -1
            else:
                vecs = ti_table_entry(ti_table, f.texinfo)[1]
                smin, smax, tmin, tmax = _extents(pts, vecs)
                w = (smax - smin) // 16 + 1
                h = (tmax - tmin) // 16 + 1
                nst = 1
                styles = [0, 255, 255, 255]
                if self.multi_style_every and fi % self.multi_style_every == 3:
                    nst = 2 + (fi // self.multi_style_every) % 3
                    pool = [1, 2, 3, 32]
                    for s in range(1, nst):
                        styles[s] = pool[(fi + s) % len(pool)]
                lightofs = len(light)
                # smooth-ish lightmaps (random base + gradient) so interpolation is visible
                base = self.rng.randint(40, 200)
                for s in range(nst):
                    gx = self.rng.randint(-12, 13); gy = self.rng.randint(-12, 13)
                    yy, xx = np.mgrid[0:h, 0:w]
                    lm = base // (1 if s == 0 else 2) + gx * xx + gy * yy + self.rng.randint(-10, 11, size=(h, w))
                    light += np.clip(lm, 0, 255).astype(np.uint8).tobytes()
                styles = tuple(styles)
            f.styles = styles
            dfaces.append(struct.pack("<hhihh4Bi", self.face_plane[fi], f.side, firstedge, len(vi),
                                      f.texinfo, *styles, lightofs))

        lump_planes = b"".join(
            struct.pack("<4fi", *[1.0 if i == ax else 0.0 for i in range(3)], c, ax) for ax, c in self.planes)
        lump_verts = b"".join(struct.pack("<3f", *v) for v in verts)
        lump_edges = b"".join(struct.pack("<HH", a, b_) for a, b_ in edges)
        lump_surfedges = struct.pack("<%di" % len(surfedges), *surfedges)
        tis = sorted(ti_table.values(), key=lambda t: t[0])
        lump_texinfo = b"".join(struct.pack("<8fii", *t[1], t[2], t[3]) for t in tis)
        lump_faces = b"".join(dfaces)

        # visibility + leaves + marksurfaces
        marks: List[int] = []
        vis = bytearray()
        dleafs = []
        nvis = self.num_world_leaves
        for li, leaf in enumerate(self.leaves):
            first = len(marks)
            marks.extend(leaf.marks)
            visofs = -1
            if leaf.vis is not None:
                visofs = len(vis)
                vis += _rle(np.packbits(leaf.vis[:nvis], bitorder="little").tobytes())
            dleafs.append(struct.pack("<ii6hHH4B", leaf.contents, visofs, *leaf.mins, *leaf.maxs,
                                      first, len(leaf.marks), 0, 0, 0, 0))
        lump_leafs = b"".join(dleafs)
        lump_marks = struct.pack("<%dh" % len(marks), *marks) if marks else b""
        lump_nodes = b"".join(
            struct.pack("<ihh6hHH", n.plane, n.children[0], n.children[1], *n.mins, *n.maxs, n.firstface, n.numfaces)
            for n in self.nodes)

        wmins = [float(self.g[i][0]) for i in range(3)]
        wmaxs = [float(self.g[i][-1]) for i in range(3)]
        models = [struct.pack("<9f4i3i", *wmins, *wmaxs, 0, 0, 0, 0, 0, 0, 0, nvis, 0, self.num_world_faces)]
        for r in self.sub_records:
            models.append(struct.pack("<9f4i3i", *r["mins"], *r["maxs"], 0, 0, 0, r["headnode"], 0, 0, 0,
                                      0, r["firstface"], r["numfaces"]))
        lump_models = b"".join(models)

        lumps = [b"", lump_planes, lump_tex, lump_verts, bytes(vis), lump_nodes, lump_texinfo, lump_faces,
                 bytes(light), b"", lump_leafs, lump_marks, lump_edges, lump_surfedges, lump_models]
        out = bytearray()
        hdr_size = 4 + 8 * 15
        pos = hdr_size
        dirs = []
        for l in lumps:
            dirs.append((pos, len(l)))
            pos += (len(l) + 3) & ~3
        out += struct.pack("<i", 29)
        for o, n in dirs:
            out += struct.pack("<ii", o, n)
        for l in lumps:
            out += l + b"\0" * (((len(l) + 3) & ~3) - len(l))
        self.stats = dict(faces=len(self.faces), world_faces=self.num_world_faces, nodes=len(self.nodes),
                          leaves=len(self.leaves), verts=len(verts), edges=len(edges), planes=len(self.planes),
                          textures=len(names), lightbytes=len(light), visbytes=len(vis))
        return bytes(out)


def ti_table_entry(table: Dict, index: int):
    for v in table.values():
        if v[0] == index:
            return v
    raise KeyError(index)


def _extents(pts, vecs):
    """texturemins / extents as CalcSurfaceExtents computes them (r_model.c:770-792)."""
    s = [np.float32(p[0]) * np.float32(vecs[0]) + np.float32(p[1]) * np.float32(vecs[1]) +
         np.float32(p[2]) * np.float32(vecs[2]) + np.float32(vecs[3]) for p in pts]
    t = [np.float32(p[0]) * np.float32(vecs[4]) + np.float32(p[1]) * np.float32(vecs[5]) +
         np.float32(p[2]) * np.float32(vecs[6]) + np.float32(vecs[7]) for p in pts]
    smin = int(np.floor(min(s) / 16)) * 16; smax = int(np.ceil(max(s) / 16)) * 16
    tmin = int(np.floor(min(t) / 16)) * 16; tmax = int(np.ceil(max(t) / 16)) * 16
    assert smax - smin <= 256 and tmax - tmin <= 256, (smin, smax, tmin, tmax)
    return smin, smax, tmin, tmax


def _cuts(a: float, b: float) -> List[float]:
    """Subdivide [a,b] into pieces <= SUBDIVIDE/2 .. SUBDIVIDE (texture scale up to 2 keeps
    extents <= 256 with SUBDIVIDE/2 pieces, so use 112 to be safe for scaled texinfos)."""
    step = 112.0
    out = [a]
    while b - out[-1] > step:
        out.append(out[-1] + step)
    out.append(b)
    return out


def _rle(row: bytes) -> bytes:
    """Quake PVS run-length encoding of zero bytes (r_model.c:146-161 decodes it)."""
    out = bytearray()
    i = 0
    while i < len(row):
        if row[i]:
            out.append(row[i]); i += 1
            continue
        j = i
        while j < len(row) and row[j] == 0 and j - i < 255:
            j += 1
        out += bytes([0, j - i])
        i = j
    return bytes(out)


def _segments_clear(p: np.ndarray, q: np.ndarray, xs: np.ndarray, ys: np.ndarray, open2d: np.ndarray) -> np.ndarray:
    """True where the 2D segment p->q only crosses open columns.  Samples the midpoint of
    every interval between consecutive grid-line crossings (exact for a rectilinear grid)."""
    d = q - p
    with np.errstate(divide="ignore", invalid="ignore"):
        tx = (xs[None, :] - p[:, 0:1]) / d[:, 0:1]
        ty = (ys[None, :] - p[:, 1:2]) / d[:, 1:2]
    t = np.concatenate([tx, ty, np.zeros((len(p), 1)), np.ones((len(p), 1))], axis=1)
    t = np.where(np.isfinite(t), t, 1.0)
    t = np.clip(t, 0.0, 1.0)
    t.sort(axis=1)
    mid = (t[:, 1:] + t[:, :-1]) * 0.5
    mx = p[:, 0:1] + d[:, 0:1] * mid
    my = p[:, 1:2] + d[:, 1:2] * mid
    ix = np.clip(np.searchsorted(xs, mx, side="right") - 1, 0, open2d.shape[0] - 1)
    iy = np.clip(np.searchsorted(ys, my, side="right") - 1, 0, open2d.shape[1] - 1)
    return open2d[ix, iy].all(axis=1)
