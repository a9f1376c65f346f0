/*
 * rc_host_model.c -- BSP29 loader producing flat, pointer-free arrays for upload to HBM.
 *
 * Behaviour follows Mod_LoadBrushModel and its helpers (ref_soft/r_model.c:380-1127):
 * same lump order, same derived fields (texturemins/extents, mipadjust, sky/turb flags,
 * texture animation chains, parents), same rejection of malformed data -- but errors are
 * returned, not Sys_Error'd, and nothing is allocated from a Hunk.
 * Added at load time because the GPU pipeline wants them precomputed:
 *   nodevis    per-leaf bitset of the nodes/leaves R_MarkLeaves (r_main.c:420-447) would
 *              stamp with r_visframecount for that view leaf
 *   separtner  for every surfedge, the other world face referencing the same medge_t
 *              (resolves the cachededgeoffset edge cache of r_rast.c:583-603 without state)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rc_host.h"

/* on-disk structs, common/bspfile.h:56-225 */
typedef struct { int fileofs, filelen; } lump_t;
enum { LUMP_ENTITIES, LUMP_PLANES, LUMP_TEXTURES, LUMP_VERTEXES, LUMP_VISIBILITY, LUMP_NODES, LUMP_TEXINFO,
       LUMP_FACES, LUMP_LIGHTING, LUMP_CLIPNODES, LUMP_LEAFS, LUMP_MARKSURFACES, LUMP_EDGES, LUMP_SURFEDGES,
       LUMP_MODELS, HEADER_LUMPS };
typedef struct { int version; lump_t lumps[HEADER_LUMPS]; } dheader_t;
typedef struct { float normal[3]; float dist; int type; } dplane_t;
typedef struct { int planenum; short children[2]; short mins[3]; short maxs[3]; unsigned short firstface, numfaces; } dnode_t;
typedef struct { float vecs[2][4]; int miptex; int flags; } dtexinfo_t;
typedef struct { short planenum; short side; int firstedge; short numedges; short texinfo; unsigned char styles[4]; int lightofs; } dface_t;
typedef struct { int contents; int visofs; short mins[3]; short maxs[3]; unsigned short firstmarksurface, nummarksurfaces; unsigned char ambient_level[4]; } dleaf_t;
typedef struct { float mins[3], maxs[3], origin[3]; int headnode[4]; int visleafs; int firstface, numfaces; } dmodel_t;
typedef struct { char name[16]; unsigned width, height; unsigned offsets[4]; } miptex_t;
#define TEX_SPECIAL 1

#define FAIL(...) do { snprintf (err, errlen, __VA_ARGS__); goto fail; } while (0)

static void *own (rc_hostworld_t *hw, size_t bytes)
{
	void *p = calloc (1, bytes ? bytes : 1);
	if (p && hw->numowned < 32)
		hw->owned[hw->numowned++] = p;
	return p;
}

void RC_FreeHostWorld (rc_hostworld_t *hw)
{
	int i;
	if (!hw) return;
	for (i = 0; i < hw->numowned; i++)
		free (hw->owned[i]);
	free (hw);
}

static float vector_length (const float *v)
{
	/* common/mathlib.c VectorLength: float sum, double sqrt, float result */
	int i;
	float length = 0;
	for (i = 0; i < 3; i++)
		length += v[i]*v[i];
	length = sqrt (length);
	return length;
}

/* Mod_LoadLighting with r_loadlits set (r_model.c:528-566): "<map>.lit", "QLIT" version 1, RGB triples; the software
 * renderer keeps max (r, g, b) per sample.  Anything else leaves the BSP's own lighting lump in place, as the
 * reference falls back to it.  Returns 1 when the .lit was used. */
int RC_ApplyLit (rc_hostworld_t *hw, const uint8_t *lit, size_t litlen)
{
	uint8_t *out = (uint8_t *)hw->w.lightdata;
	size_t i, s;
	int ver;
	if (!hw->w.haslight || !lit || litlen < 8) return 0;		/* r_model.c:516-520: no lump, no light data at all */
	if (lit[0] != 'Q' || lit[1] != 'L' || lit[2] != 'I' || lit[3] != 'T') return 0;
	memcpy (&ver, lit + 4, 4);
	if (ver != 1) return 0;
	s = (litlen - 8) / 3;
	if (s > hw->lightbytes) s = hw->lightbytes;			/* samples past the lump are never indexed (lightofs is checked against it) */
	for (i = 0; i < s; i++)
	{
		const uint8_t *t = lit + 8 + 3 * i;
		uint8_t p = t[0] > t[1] ? t[0] : t[1];
		out[i] = p > t[2] ? p : t[2];
	}
	return 1;
}

rc_hostworld_t *RC_LoadBrushModel (const void *data, size_t len, const uint8_t *colormap, char *err, size_t errlen)
{
	const unsigned char *base = (const unsigned char *)data;
	const dheader_t *h = (const dheader_t *)data;
	rc_hostworld_t *hw = NULL;
	rc_world_t *w;
	int i, j, k, count;
	char (*texnames)[16] = NULL;
	int *medge_refs = NULL;

	if (len < sizeof(dheader_t)) { snprintf (err, errlen, "bsp too short"); return NULL; }
	if (h->version != 29) { snprintf (err, errlen, "wrong BSP version %d (should be 29)", h->version); return NULL; }
	for (i = 0; i < HEADER_LUMPS; i++)
		if (h->lumps[i].fileofs < 0 || h->lumps[i].filelen < 0 ||
		    (size_t)h->lumps[i].fileofs + (size_t)h->lumps[i].filelen > len)
		{ snprintf (err, errlen, "lump %d out of file", i); return NULL; }

	hw = (rc_hostworld_t *)calloc (1, sizeof(*hw));
	w = &hw->w;
#define LUMP(T, idx, ptr, n) \
	const T *ptr = (const T *)(base + h->lumps[idx].fileofs); \
	if (h->lumps[idx].filelen % sizeof(T)) FAIL ("funny lump size (lump %d)", idx); \
	int n = h->lumps[idx].filelen / sizeof(T);

	/* vertexes, edges, surfedges (r_model.c:597-690, :994-1013) */
	{
		LUMP (float, LUMP_VERTEXES, in, n)
		float *out;
		if (n % 3) FAIL ("funny vertex lump");
		w->numvertexes = n / 3;
		out = (float *)own (hw, sizeof(float) * (n + 24));
		memcpy (out, in, sizeof(float) * n);
		w->vertexes = out;
	}
	{
		LUMP (unsigned short, LUMP_EDGES, in, n)
		uint16_t *out;
		if (n % 2) FAIL ("funny edge lump");
		w->numedges = n / 2;
		out = (uint16_t *)own (hw, sizeof(uint16_t) * (n + 26));
		memcpy (out, in, sizeof(uint16_t) * n);
		w->edges = out;
	}
	{
		LUMP (int, LUMP_SURFEDGES, in, n)
		int *out = (int *)own (hw, sizeof(int) * (n + 24));
		memcpy (out, in, sizeof(int) * n);
		w->numsurfedges = n;
		w->surfedges = out;
		for (i = 0; i < n; i++)
			if (abs (out[i]) >= w->numedges) FAIL ("surfedge %d out of range", i);
	}

	/* textures (r_model.c:360-500) */
	{
		const lump_t *l = &h->lumps[LUMP_TEXTURES];
		rc_texture_t *tx;
		uint8_t *texels;
		size_t total = 0, ofs = 0;
		int nummiptex = 0;
		const int *dataofs = NULL;
		if (l->filelen)
		{
			nummiptex = *(const int *)(base + l->fileofs);
			dataofs = (const int *)(base + l->fileofs + 4);
			if (nummiptex < 0 || (size_t)nummiptex * 4 + 4 > (size_t)l->filelen) FAIL ("bad texture lump");
		}
		w->numtextures = nummiptex;
		tx = (rc_texture_t *)own (hw, sizeof(rc_texture_t) * (nummiptex + 1));
		texnames = (char (*)[16])calloc (nummiptex + 1, 16);
		for (i = 0; i < nummiptex; i++)
		{
			const miptex_t *mt;
			if (dataofs[i] == -1) continue;
			if (dataofs[i] < 0 || (size_t)dataofs[i] + sizeof(miptex_t) > (size_t)l->filelen) FAIL ("miptex %d out of lump", i);
			mt = (const miptex_t *)(base + l->fileofs + dataofs[i]);
			if ((mt->width & 15) || (mt->height & 15)) FAIL ("Texture %.16s is not 16 aligned", mt->name);
			total += ((size_t)mt->width * mt->height / 64 * 85 + 15) & ~(size_t)15;	/* 16-byte aligned for vector loads */
		}
		texels = (uint8_t *)own (hw, total + 16);
		for (i = 0; i < nummiptex; i++)
		{
			const miptex_t *mt;
			size_t pixels;
			tx[i].anim_next = -1;
			tx[i].alternate_anims = -1;
			if (dataofs[i] == -1) { tx[i].width = 0; continue; }
			mt = (const miptex_t *)(base + l->fileofs + dataofs[i]);
			pixels = (size_t)mt->width * mt->height / 64 * 85;
			if ((size_t)dataofs[i] + sizeof(miptex_t) + pixels > (size_t)l->filelen) FAIL ("miptex %d pixels out of lump", i);
			memcpy (texnames[i], mt->name, 16);
			tx[i].width = mt->width;
			tx[i].height = mt->height;
			for (j = 0; j < 4; j++)
			{
				/* offsets are relative to the miptex_t header; the pixels follow it */
				if (mt->offsets[j] < sizeof(miptex_t) || mt->offsets[j] - sizeof(miptex_t) >= pixels) FAIL ("miptex %d bad mip offset", i);
				tx[i].offsets[j] = (int)(ofs + mt->offsets[j] - sizeof(miptex_t));
			}
			memcpy (texels + ofs, (const unsigned char *)(mt + 1), pixels);
			ofs += (pixels + 15) & ~(size_t)15;
		}
		w->skytexofs = -1;
		for (i = 0; i < nummiptex; i++)
			if (tx[i].width && !strncmp (texnames[i], "sky", 3))
			{
				/* R_InitSky (r_sky.c:53-61) memcpy's 256*128 bytes: every "sky*" texture overwrites skydata */
				if (tx[i].width != 256 || tx[i].height != 128) FAIL ("sky texture %.16s is not 256x128", texnames[i]);
				w->skytexofs = tx[i].offsets[0];
			}
		hw->texelbytes = total;
		w->texels = texels;
		w->textures = tx;

		/* sequence the animations (r_model.c:415-500): '+0'..'+9' and '+a'..'+j' */
		for (i = 0; i < nummiptex; i++)
		{
			int anims[10], altanims[10], max, altmax, num;
			if (!tx[i].width || texnames[i][0] != '+') continue;
			if (tx[i].anim_next >= 0) continue;
			for (j = 0; j < 10; j++) anims[j] = altanims[j] = -1;
			max = texnames[i][1];
			altmax = 0;
			if (max >= 'a' && max <= 'z') max -= 'a' - 'A';
			if (max >= '0' && max <= '9') { max -= '0'; altmax = 0; anims[max] = i; max++; }
			else if (max >= 'A' && max <= 'J') { altmax = max - 'A'; max = 0; altanims[altmax] = i; altmax++; }
			else FAIL ("Bad animating texture %.16s", texnames[i]);
			for (j = i + 1; j < nummiptex; j++)
			{
				if (!tx[j].width || texnames[j][0] != '+') continue;
				if (strncmp (texnames[j] + 2, texnames[i] + 2, 14)) continue;
				num = texnames[j][1];
				if (num >= 'a' && num <= 'z') num -= 'a' - 'A';
				if (num >= '0' && num <= '9') { num -= '0'; anims[num] = j; if (num + 1 > max) max = num + 1; }
				else if (num >= 'A' && num <= 'J') { num -= 'A'; altanims[num] = j; if (num + 1 > altmax) altmax = num + 1; }
				else FAIL ("Bad animating texture %.16s", texnames[i]);
			}
			for (j = 0; j < max; j++)
			{
				int t2 = anims[j];
				if (t2 < 0) FAIL ("Missing frame %i of %.16s", j, texnames[i]);
				tx[t2].anim_total = max * 2;
				tx[t2].anim_min = j * 2;
				tx[t2].anim_max = (j + 1) * 2;
				tx[t2].anim_next = anims[(j + 1) % max];
				if (altmax) tx[t2].alternate_anims = altanims[0];
			}
			for (j = 0; j < altmax; j++)
			{
				int t2 = altanims[j];
				if (t2 < 0) FAIL ("Missing frame %i of %.16s", j, texnames[i]);
				tx[t2].anim_total = altmax * 2;
				tx[t2].anim_min = j * 2;
				tx[t2].anim_max = (j + 1) * 2;
				tx[t2].anim_next = altanims[(j + 1) % altmax];
				if (max) tx[t2].alternate_anims = anims[0];
			}
		}
	}

	/* lighting (r_model.c:514-526; r_loadlits 0) */
	{
		const lump_t *l = &h->lumps[LUMP_LIGHTING];
		uint8_t *out = (uint8_t *)own (hw, (size_t)l->filelen + 16);
		memcpy (out, base + l->fileofs, l->filelen);
		w->lightdata = out;
		w->haslight = l->filelen != 0;
		hw->lightbytes = l->filelen;
	}

	/* planes (r_model.c:1023-1055) */
	{
		LUMP (dplane_t, LUMP_PLANES, in, n)
		rc_plane_t *out = (rc_plane_t *)own (hw, sizeof(rc_plane_t) * (n + 6));
		for (i = 0; i < n; i++)
		{
			for (j = 0; j < 3; j++) out[i].normal[j] = in[i].normal[j];
			out[i].dist = in[i].dist;
			out[i].type = in[i].type;
		}
		w->numplanes = n;
		w->planes = out;
	}

	/* texinfo (r_model.c:692-760) */
	{
		LUMP (dtexinfo_t, LUMP_TEXINFO, in, n)
		rc_texinfo_t *out = (rc_texinfo_t *)own (hw, sizeof(rc_texinfo_t) * (n + 6));
		for (i = 0; i < n; i++)
		{
			float len1, len2;
			memcpy (out[i].vecs, in[i].vecs, sizeof(out[i].vecs));
			len1 = vector_length (out[i].vecs[0]);
			len2 = vector_length (out[i].vecs[1]);
			len1 = (len1 + len2)/2;
			if (len1 < 0.32) out[i].mipadjust = 4;
			else if (len1 < 0.49) out[i].mipadjust = 3;
			else if (len1 < 0.99) out[i].mipadjust = 2;
			else out[i].mipadjust = 1;
			out[i].flags = in[i].flags;
			if (!w->numtextures) FAIL ("map without textures is not supported");
			if (in[i].miptex < 0 || in[i].miptex >= w->numtextures) FAIL ("miptex >= loadmodel->numtextures");
			out[i].texture = in[i].miptex;
			if (!w->textures[in[i].miptex].width) FAIL ("texinfo %d references a missing texture", i);
		}
		w->numtexinfo = n;
		w->texinfo = out;
	}

	/* faces (r_model.c:762-870) */
	{
		LUMP (dface_t, LUMP_FACES, in, n)
		rc_face_t *out = (rc_face_t *)own (hw, sizeof(rc_face_t) * (n + 6));
		for (i = 0; i < n; i++)
		{
			rc_face_t *f = &out[i];
			const rc_texinfo_t *tex;
			float mins[2], maxs[2], val;
			int bmins[2], bmaxs[2], e;
			const float *vtx;
			const char *tname;

			f->firstedge = in[i].firstedge;
			f->numedges = in[i].numedges;
			f->flags = 0;
			if (in[i].side) f->flags |= RC_SURF_PLANEBACK;
			f->plane = in[i].planenum;
			f->texinfo = in[i].texinfo;
			f->node = -1;
			if (f->plane < 0 || f->plane >= w->numplanes) FAIL ("face %d bad plane", i);
			if (f->texinfo < 0 || f->texinfo >= w->numtexinfo) FAIL ("face %d bad texinfo", i);
			if (f->firstedge < 0 || f->numedges < 0 || f->firstedge + f->numedges > w->numsurfedges) FAIL ("face %d bad edges", i);
			tex = &w->texinfo[f->texinfo];

			/* CalcSurfaceExtents, r_model.c:770-806 */
			mins[0] = mins[1] = 999999;
			maxs[0] = maxs[1] = -99999;
			for (j = 0; j < f->numedges; j++)
			{
				e = w->surfedges[f->firstedge + j];
				if (e >= 0) vtx = &w->vertexes[3 * w->edges[2 * e + 0]];
				else vtx = &w->vertexes[3 * w->edges[2 * (-e) + 1]];
				for (k = 0; k < 2; k++)
				{
					val = vtx[0] * tex->vecs[k][0] + vtx[1] * tex->vecs[k][1] + vtx[2] * tex->vecs[k][2] + tex->vecs[k][3];
					if (val < mins[k]) mins[k] = val;
					if (val > maxs[k]) maxs[k] = val;
				}
			}
			for (k = 0; k < 2; k++)
			{
				bmins[k] = floor (mins[k]/16);
				bmaxs[k] = ceil (maxs[k]/16);
				f->texturemins[k] = bmins[k] * 16;
				f->extents[k] = (bmaxs[k] - bmins[k]) * 16;
				if (!(tex->flags & TEX_SPECIAL) && f->extents[k] > 256) FAIL ("Bad surface extents");
			}

			for (k = 0; k < RC_MAXLIGHTMAPS; k++) f->styles[k] = in[i].styles[k];
			f->lightofs = in[i].lightofs;
			if (f->lightofs != -1 && (f->lightofs < 0 || (size_t)f->lightofs >= hw->lightbytes + 1)) FAIL ("face %d bad lightofs", i);

			tname = texnames[tex->texture];
			if (!strncmp (tname, "sky", 3))
			{
				f->flags |= (RC_SURF_DRAWSKY | RC_SURF_DRAWTILED);
				for (k = 0; k < 2; k++) f->extents[k] = 256;
				continue;
			}
			if (tname[0] == '*')
			{
				f->flags |= (RC_SURF_DRAWTURB | RC_SURF_DRAWTILED);
				for (k = 0; k < 2; k++) { f->extents[k] = 16384; f->texturemins[k] = -8192; }
				continue;
			}
		}
		w->numfaces = n;
		w->faces = out;
	}

	/* marksurfaces (r_model.c:965-987) */
	{
		LUMP (short, LUMP_MARKSURFACES, in, n)
		int *out = (int *)own (hw, sizeof(int) * (n + 1));
		for (i = 0; i < n; i++)
		{
			if (in[i] >= w->numfaces || in[i] < 0) FAIL ("Mod_ParseMarksurfaces: bad surface number");
			out[i] = in[i];
		}
		w->nummarksurfaces = n;
		w->marksurfaces = out;
		{
			int *ml = (int *)own (hw, sizeof(int) * (n + 1));
			for (i = 0; i <= n; i++) ml[i] = -1;
			w->markleaf = ml;
			w->marksplit = 1;
		}
	}

	/* leafs, nodes (r_model.c:872-963) */
	{
		LUMP (dleaf_t, LUMP_LEAFS, lin, nl)
		LUMP (dnode_t, LUMP_NODES, nin, nn)
		LUMP (dmodel_t, LUMP_MODELS, min_, nm)
		rc_leaf_t *leafs = (rc_leaf_t *)own (hw, sizeof(rc_leaf_t) * (nl + 1));
		rc_node_t *nodes = (rc_node_t *)own (hw, sizeof(rc_node_t) * (nn + 1));
		const unsigned char *visdata = base + h->lumps[LUMP_VISIBILITY].fileofs;
		int vislen = h->lumps[LUMP_VISIBILITY].filelen;
		uint32_t *nodevis;
		int rowwords, visleafs;
		int *stack;

		if (nl < 1 || nn < 1 || nm < 1) FAIL ("bsp needs at least one leaf, node and model");
		for (i = 0; i < nl; i++)
		{
			for (j = 0; j < 3; j++) { leafs[i].minmaxs[j] = lin[i].mins[j]; leafs[i].minmaxs[3 + j] = lin[i].maxs[j]; }
			leafs[i].contents = lin[i].contents;
			leafs[i].firstmark = lin[i].firstmarksurface;
			leafs[i].nummarks = lin[i].nummarksurfaces;
			leafs[i].parent = -1;
			if (leafs[i].firstmark + leafs[i].nummarks > w->nummarksurfaces) FAIL ("leaf %d bad marksurfaces", i);
			for (j = 0; j < leafs[i].nummarks; j++)
			{
				int *ml = (int *)w->markleaf;
				if (ml[leafs[i].firstmark + j] != -1) w->marksplit = 0;
				ml[leafs[i].firstmark + j] = i;
			}
		}
		for (i = 0; i < nn; i++)
		{
			for (j = 0; j < 3; j++) { nodes[i].minmaxs[j] = nin[i].mins[j]; nodes[i].minmaxs[3 + j] = nin[i].maxs[j]; }
			nodes[i].plane = nin[i].planenum;
			if (nodes[i].plane < 0 || nodes[i].plane >= w->numplanes) FAIL ("node %d bad plane", i);
			nodes[i].firstsurface = nin[i].firstface;
			nodes[i].numsurfaces = nin[i].numfaces;
			if (nodes[i].firstsurface + nodes[i].numsurfaces > w->numfaces) FAIL ("node %d bad faces", i);
			for (j = 0; j < nodes[i].numsurfaces; j++)
			{
				((rc_face_t *)w->faces)[nodes[i].firstsurface + j].node = i;
				((rc_face_t *)w->faces)[nodes[i].firstsurface + j].nodeslot = j;
			}
			nodes[i].parent = -1;
			for (j = 0; j < 2; j++)
			{
				int p = nin[i].children[j];
				if (p >= 0) { if (p >= nn) FAIL ("node %d bad child", i); nodes[i].children[j] = p; }
				else { if (-1 - p >= nl) FAIL ("node %d bad leaf child", i); nodes[i].children[j] = p; }
			}
		}
		/* Mod_SetParent from every model's head node (world sets all reachable parents) */
		stack = (int *)malloc (sizeof(int) * (nn + 1));
		for (k = 0; k < nm; k++)
		{
			int sp = 0, hn = min_[k].headnode[0];
			if (hn < 0 || hn >= nn) continue;
			stack[sp++] = hn;
			while (sp)
			{
				int n = stack[--sp];
				for (j = 0; j < 2; j++)
				{
					int c = nodes[n].children[j];
					if (c >= 0) { if (nodes[c].parent == -1 && c != hn) { nodes[c].parent = n; stack[sp++] = c; } }
					else if (leafs[-1 - c].parent == -1) leafs[-1 - c].parent = n;
				}
			}
		}
		free (stack);
		w->numleafs = nl;
		w->numnodes = nn;
		w->leafs = leafs;
		w->nodes = nodes;

		/* submodels (r_model.c:624-655, :1098-1126) */
		hw->numsubmodels = nm;
		hw->submodels = (rc_submodel_t *)own (hw, sizeof(rc_submodel_t) * nm);
		for (i = 0; i < nm; i++)
		{
			rc_submodel_t *sm = &hw->submodels[i];
			float corner[3];
			for (j = 0; j < 3; j++)
			{	/* spread the mins / maxs by a pixel */
				sm->mins[j] = min_[i].mins[j] - 1;
				sm->maxs[j] = min_[i].maxs[j] + 1;
			}
			/* RadiusFromBounds, common/mathlib.c */
			for (j = 0; j < 3; j++)
				corner[j] = fabs (sm->mins[j]) > fabs (sm->maxs[j]) ? fabs (sm->mins[j]) : fabs (sm->maxs[j]);
			sm->radius = vector_length (corner);
			sm->headnode = min_[i].headnode[0];
			sm->visleafs = min_[i].visleafs;
			sm->firstface = min_[i].firstface;
			sm->numfaces = min_[i].numfaces;
			if (sm->firstface < 0 || sm->numfaces < 0 || sm->firstface + sm->numfaces > w->numfaces) FAIL ("model %d bad faces", i);
		}
		visleafs = hw->submodels[0].visleafs;
		if (visleafs < 0 || visleafs + 1 > nl) FAIL ("bad visleafs");
		w->visleafs = visleafs;
		w->numworldfaces = hw->submodels[0].numfaces;

		/* nodevis: what R_MarkLeaves (r_main.c:420-447) marks for each possible view leaf,
		 * PVS rows decoded as Mod_DecompressVis does (r_model.c:126-164) */
		rowwords = (nn + nl + 31) / 32;
		w->visrowwords = rowwords;
		nodevis = (uint32_t *)own (hw, sizeof(uint32_t) * (size_t)rowwords * nl);
		{
			int row = (visleafs + 7) >> 3;
			unsigned char *dec = (unsigned char *)malloc (row + 8);
			for (i = 0; i < nl; i++)
			{
				uint32_t *bits = nodevis + (size_t)i * rowwords;
				int visofs = lin[i].visofs;
				if (i == 0 || visofs == -1)
					memset (dec, 0xff, row);	/* mod_novis / no vis info */
				else
				{
					const unsigned char *in = visdata + visofs;
					const unsigned char *end = visdata + vislen;
					int o = 0;
					if (visofs < 0 || visofs >= vislen) { free (dec); FAIL ("leaf %d bad visofs", i); }
					do
					{
						if (in >= end) { free (dec); FAIL ("leaf %d: PVS row runs off the lump", i); }
						if (*in) { dec[o++] = *in++; continue; }
						if (in + 1 >= end) { free (dec); FAIL ("leaf %d: PVS row runs off the lump", i); }
						k = in[1];
						in += 2;
						while (k && o < row) { dec[o++] = 0; k--; }
					} while (o < row);
				}
				for (j = 0; j < visleafs; j++)
				{
					if (dec[j >> 3] & (1 << (j & 7)))
					{
						int idx = nn + (j + 1);
						int par = leafs[j + 1].parent;
						bits[idx >> 5] |= 1u << (idx & 31);
						while (par >= 0)
						{
							if (bits[par >> 5] & (1u << (par & 31))) break;
							bits[par >> 5] |= 1u << (par & 31);
							par = nodes[par].parent;
						}
					}
				}
			}
			free (dec);
		}
		w->nodevis = nodevis;
	}

	/* nodeaux / leafdfs: static numbering of the world tree for the parallel walk */
	{
		rc_nodeaux_t *aux = (rc_nodeaux_t *)own (hw, sizeof(rc_nodeaux_t) * (w->numnodes + 1));
		int *leafdfs = (int *)own (hw, sizeof(int) * (w->numleafs + 1));
		int *stk = (int *)malloc (sizeof(int) * 2 * (w->numnodes + 2));
		int *order = (int *)malloc (sizeof(int) * (w->numnodes + 2));
		int *dfs_out = (int *)malloc (sizeof(int) * (w->numnodes + 2));
		int sp = 0, counter = 0, norder = 0;
		for (i = 0; i < w->numleafs; i++) leafdfs[i] = -1;
		/* iterative pre-order from every model's head node: model 0 is the world tree the walk
		 * replays, the others are brush-entity trees (only R_MarkLights descends them) */
		for (k = 0; k < hw->numsubmodels; k++)
		{
			int hn = hw->submodels[k].headnode;
			if (hn < 0 || hn >= w->numnodes)
				continue;
			sp = 0; norder = 0;
			if (k > 0) counter = 0;
			stk[sp++] = hn;
			while (sp)
			{
				int n = stk[--sp];
				if (n < 0)
				{
					if (k == 0 && -1 - n != 0) leafdfs[-1 - n] = counter;	/* leaf 0 (solid) is shared: never numbered */
					counter++;
					continue;
				}
				aux[n].dfs_in = counter++;
				order[norder++] = n;
				stk[sp++] = w->nodes[n].children[1];
				stk[sp++] = w->nodes[n].children[0];
			}
			/* dfs_mid and slot counts bottom-up (reverse pre-order) */
			for (i = norder - 1; i >= 0; i--)
			{
				int n = order[i], out = aux[n].dfs_in + 1;
				for (j = 0; j < 2; j++)
				{
					int c = w->nodes[n].children[j];
					if (c < 0) { aux[n].slots[j] = 1; out += 1; }
					else { aux[n].slots[j] = aux[c].slots[0] + aux[c].slots[1] + w->nodes[c].numsurfaces + 1; out += dfs_out[c] - aux[c].dfs_in; }
					if (j == 0) aux[n].dfs_mid = out;
				}
				dfs_out[n] = out;
			}
		}
		free (stk); free (order); free (dfs_out);
		w->nodeaux = aux;
		w->leafdfs = leafdfs;
	}

	/* separtner: world faces sharing a medge_t */
	{
		rc_separtner_t *sp = (rc_separtner_t *)own (hw, sizeof(rc_separtner_t) * (w->numsurfedges + 1));
		int wf0 = hw->submodels[0].firstface, wfn = hw->submodels[0].numfaces;
		int *se_face = (int *)malloc (sizeof(int) * (w->numsurfedges + 1));
		medge_refs = (int *)malloc (sizeof(int) * 3 * (w->numedges + 1));
		for (i = 0; i < w->numsurfedges; i++) se_face[i] = -1;
		for (i = wf0; i < wf0 + wfn; i++)
			for (j = 0; j < w->faces[i].numedges; j++)
				se_face[w->faces[i].firstedge + j] = i;
		for (i = 0; i < w->numedges; i++) { medge_refs[3 * i] = 0; medge_refs[3 * i + 1] = medge_refs[3 * i + 2] = -1; }
		for (i = 0; i < w->numsurfedges; i++) sp[i].partner_face = sp[i].partner_slot = -1;
		for (i = wf0; i < wf0 + wfn; i++)
			for (j = 0; j < w->faces[i].numedges; j++)
			{
				int se = w->faces[i].firstedge + j;
				int m = abs (w->surfedges[se]);
				int c = medge_refs[3 * m]++;
				if (c < 2) medge_refs[3 * m + 1 + c] = se;
			}
		for (i = wf0; i < wf0 + wfn; i++)
			for (j = 0; j < w->faces[i].numedges; j++)
			{
				int se = w->faces[i].firstedge + j;
				int m = abs (w->surfedges[se]);
				int other;
				if (medge_refs[3 * m] != 2) continue;	/* 1: unshared; >2: treated as unshared (never produced by qbsp) */
				other = medge_refs[3 * m + 1] == se ? medge_refs[3 * m + 2] : medge_refs[3 * m + 1];
				k = se_face[other];
				if (k == i || k < 0) continue;		/* both uses inside one face: leave unshared */
				sp[se].partner_face = k;
				sp[se].partner_slot = other - w->faces[k].firstedge;
			}
		w->separtner = sp;
		{
			/* flattened per-surfedge records for the face kernel */
			rc_fedge_t *fe = (rc_fedge_t *)own (hw, sizeof(rc_fedge_t) * (w->numsurfedges + 24 + 1));
			for (i = 0; i < w->numsurfedges; i++)
			{
				int l = w->surfedges[i], m = abs (l);
				const float *a = &w->vertexes[3 * w->edges[2 * m + (l > 0 ? 0 : 1)]];
				const float *b = &w->vertexes[3 * w->edges[2 * m + (l > 0 ? 1 : 0)]];
				memcpy (fe[i].v0, a, sizeof(float) * 3);
				memcpy (fe[i].v1, b, sizeof(float) * 3);
				fe[i].partner_face = sp[i].partner_face;
				fe[i].partner_rev = 0;
				if (sp[i].partner_face >= 0)
				{
					int pl = w->surfedges[w->faces[sp[i].partner_face].firstedge + sp[i].partner_slot];
					fe[i].partner_rev = ((pl > 0) != (l > 0));
				}
			}
			w->fedges = fe;
		}
		free (se_face);
		free (medge_refs);
		medge_refs = NULL;
	}

	/* R_InitSkyBox (r_rast.c:74-157): six faces of a box around the viewer, appended after every other
	 * face.  What depends on the view origin (vertex positions, plane distances, texture offsets) is
	 * applied per view on the device; here: topology, axis, sign, texture axes.  v0/v1 of the box's
	 * rc_fedge_t hold the vertex SIGNS (box_verts), plane.dist holds +-128. */
	{
		static const int skybox_planes[12] = {2,-128, 0,-128, 2,128, 1,128, 0,128, 1,-128};
		static const int box_surfedges[24] = { 1,2,3,4,  -1,5,6,7,  8,9,-6,10,  -2,-7,-9,11,  12,-3,-11,-8,  -12,-10,-5,-4};
		static const int box_edges[24] = { 1,2, 2,3, 3,4, 4,1, 1,5, 5,6, 6,2, 7,8, 8,6, 5,7, 8,3, 7,4};
		static const int box_faces[6] = {0,0,2,2,2,0};
		static const float box_vecs[6][2][3] = {
			{ {0,-1,0}, {-1,0,0} }, { {0,1,0}, {0,0,-1} }, { {0,-1,0}, {1,0,0} },
			{ {1,0,0}, {0,0,-1} }, { {0,-1,0}, {0,0,-1} }, { {-1,0,0}, {0,0,-1} } };
		static const float box_verts[8][3] = { {-1,-1,-1}, {-1,1,-1}, {1,1,-1}, {1,-1,-1}, {-1,-1,1}, {-1,1,1}, {1,-1,1}, {1,1,1} };
		rc_face_t *faces = (rc_face_t *)w->faces;
		rc_plane_t *planes = (rc_plane_t *)w->planes;
		rc_texinfo_t *ti = (rc_texinfo_t *)w->texinfo;
		rc_fedge_t *fe = (rc_fedge_t *)w->fedges;
		w->skyfirst = w->numfaces;
		w->drawskybox = 0;
		for (i = 0; i < 6; i++)
		{
			rc_face_t *f = &faces[w->numfaces + i];
			memset (f, 0, sizeof(*f));
			memset (&planes[w->numplanes + i], 0, sizeof(rc_plane_t));
			planes[w->numplanes + i].normal[skybox_planes[i*2]] = 1;
			planes[w->numplanes + i].dist = skybox_planes[i*2+1];
			planes[w->numplanes + i].type = skybox_planes[i*2];		/* ours: the axis (the reference leaves 0, never read) */
			memset (&ti[w->numtexinfo + i], 0, sizeof(rc_texinfo_t));
			for (j = 0; j < 3; j++) { ti[w->numtexinfo + i].vecs[0][j] = box_vecs[i][0][j]; ti[w->numtexinfo + i].vecs[1][j] = box_vecs[i][1][j]; }
			ti[w->numtexinfo + i].mipadjust = 1; ti[w->numtexinfo + i].texture = 0;
			f->plane = w->numplanes + i;
			f->numedges = 4;
			f->flags = box_faces[i] | RC_SURF_DRAWSKYBOX;
			f->firstedge = w->numsurfedges + i*4;
			f->texinfo = w->numtexinfo + i;
			f->texturemins[0] = -128; f->texturemins[1] = -128;
			f->extents[0] = 256; f->extents[1] = 256;
			f->styles[0] = f->styles[1] = f->styles[2] = f->styles[3] = 255;
			f->lightofs = -1; f->node = -1;
		}
		for (i = 0; i < 24; i++)
		{
			int l = box_surfedges[i], m = abs (l);
			int a = box_edges[(m-1)*2 + (l > 0 ? 0 : 1)] - 1, b = box_edges[(m-1)*2 + (l > 0 ? 1 : 0)] - 1;
			rc_fedge_t *e = &fe[w->numsurfedges + i];
			for (j = 0; j < 3; j++) { e->v0[j] = box_verts[a][j]; e->v1[j] = box_verts[b][j]; }
			e->partner_face = -1; e->partner_rev = 0;
			for (k = 0; k < 24; k++)
				if (k != i && abs (box_surfedges[k]) == m)
				{
					e->partner_face = w->numfaces + k / 4;
					e->partner_rev = ((box_surfedges[k] > 0) != (l > 0));
				}
		}
		w->numfaces += 6;
	}

	w->colormap = colormap;
	free (texnames);
	return hw;
fail:
	free (texnames);
	free (medge_refs);
	RC_FreeHostWorld (hw);
	return NULL;
}
