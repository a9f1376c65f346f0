/*
 * rc_host_alias.c -- host side of alias (MDL) model rendering.
 *
 *  RC_LoadAliasModel   MDL6 -> flat arrays            Mod_LoadAliasModel, ref_soft/r_model.c:1349-1572
 *  RC_AliasSetup       per (view, entity) constants   R_DrawAliasModel and its helpers,
 *                      ref_soft/r_alias.c:79-788: R_AliasCheckBBox, R_AliasSetupSkin,
 *                      R_AliasSetUpTransform, R_AliasSetupLighting (R_LightPoint,
 *                      ref_soft/r_light.c:156-265), R_AliasSetupFrame
 * These run on the host because they need libm (AngleVectors: sin/cos, VectorLength: sqrt)
 * and are per entity, not per vertex or pixel.  Vertex transform, clipping and rasterisation
 * are CUDA kernels (rc_pipeline.h).  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rc_host.h"

#define ALIAS_VERSION 6
#define MAX_LBM_HEIGHT 480          /* d_iface.h:25 */
#define MAXALIASVERTS 2024          /* r_model.h:267 */
#define ALIAS_Z_CLIP_PLANE 5        /* r_local.h:222 */
#define ALIAS_LEFT_CLIP 1
#define ALIAS_TOP_CLIP 2
#define ALIAS_RIGHT_CLIP 4
#define ALIAS_BOTTOM_CLIP 8
#define ALIAS_Z_CLIP 0x10
#define ALIAS_XY_CLIP_MASK 0xF
#define LIGHT_MIN 5                 /* r_alias.c:24 */

#define FAIL(...) do { snprintf (err, errlen, __VA_ARGS__); goto fail; } while (0)

typedef struct { int ident, version; float scale[3], scale_origin[3], boundingradius, eyeposition[3];
                 int numskins, skinwidth, skinheight, numverts, numtris, numframes, synctype, flags; float size; } mdl_t;

static int rc_alias_build_image (rc_hostalias_t *a, const unsigned char *base, const char *modelname);

void RC_FreeHostAlias (rc_hostalias_t *a)
{
	if (!a) return;
	free (a->stverts); free (a->tris); free (a->poseverts); free (a->posebbox); free (a->skinpixels);
	free (a->frames); free (a->skins); free (a->intervals); free (a->skinpicofs);
	free (a);
}

rc_hostalias_t *RC_LoadAliasModel (const void *data, size_t len, const char *modelname, char *err, size_t errlen)
{
	const unsigned char *base = (const unsigned char *)data, *p, *end = base + len;
	const mdl_t *in = (const mdl_t *)data;
	rc_hostalias_t *a = NULL;
	int i, j, k, nposes = 0, nskins = 0, nintervals = 0, skinsize, pass;

	if (len < sizeof(mdl_t)) { snprintf (err, errlen, "mdl too short"); return NULL; }
	if (in->version != ALIAS_VERSION) { snprintf (err, errlen, "wrong version number (%i should be %i)", in->version, ALIAS_VERSION); return NULL; }
	a = (rc_hostalias_t *)calloc (1, sizeof(*a));
	a->numskins = in->numskins; a->skinwidth = in->skinwidth; a->skinheight = in->skinheight;
	a->numverts = in->numverts; a->numtris = in->numtris; a->numframes = in->numframes;
	a->synctype = in->synctype; a->flags = in->flags;
	for (i = 0; i < 3; i++) { a->scale[i] = in->scale[i]; a->scale_origin[i] = in->scale_origin[i]; }
	if (a->skinheight > MAX_LBM_HEIGHT || a->skinheight <= 0 || a->skinwidth <= 0) FAIL ("model has a skin taller than %d", MAX_LBM_HEIGHT);
	if (a->numverts <= 0) FAIL ("model has no vertices");
	if (a->numverts > MAXALIASVERTS) FAIL ("model has too many vertices");
	if (a->numtris <= 0) FAIL ("model has no triangles");
	if (a->skinwidth & 0x03) FAIL ("Mod_LoadAliasModel: skinwidth not multiple of 4");
	if (a->numskins < 1) FAIL ("Mod_LoadAliasModel: Invalid # of skins: %d", a->numskins);
	if (a->numframes < 1) FAIL ("Mod_LoadAliasModel: Invalid # of frames: %d", a->numframes);
	skinsize = a->skinwidth * a->skinheight;

	/* two passes over the variable-length part: count, then fill */
	for (pass = 0; pass < 2; pass++)
	{
		int pose = 0, skin = 0, iv = 0;
		p = base + sizeof(mdl_t);
		for (i = 0; i < a->numskins; i++)
		{
			int type;
			if (p + 4 > end) FAIL ("mdl truncated (skins)");
			type = *(const int *)p; p += 4;
			if (type == 0)
			{
				if (p + skinsize > end) FAIL ("mdl truncated (skin)");
				if (pass) { a->skins[i].type = 0; a->skins[i].first = skin; a->skins[i].count = 1; a->skins[i].intervals = -1;
					    memcpy (a->skinpixels + (size_t)skin * skinsize, p, skinsize); }
				p += skinsize; skin++;
			}
			else
			{
				int n;
				if (p + 4 > end) FAIL ("mdl truncated (skin group)");
				n = *(const int *)p; p += 4;
				if (n < 1 || p + 4 * (size_t)n > end) FAIL ("bad skin group");
				if (pass) { a->skins[i].type = 1; a->skins[i].first = skin; a->skins[i].count = n; a->skins[i].intervals = iv; }
				for (j = 0; j < n; j++)
				{
					float f = *(const float *)(p + 4 * j);
					if (f <= 0) FAIL ("Mod_LoadAliasSkinGroup: interval<=0");
					if (pass) a->intervals[iv + j] = f;
				}
				iv += n; p += 4 * (size_t)n;
				if (p + (size_t)n * skinsize > end) FAIL ("mdl truncated (skin group pixels)");
				if (pass) memcpy (a->skinpixels + (size_t)skin * skinsize, p, (size_t)n * skinsize);
				p += (size_t)n * skinsize; skin += n;
			}
		}
		/* stverts, triangles (r_model.c:1468-1503) */
		if (p + 12 * (size_t)a->numverts + 16 * (size_t)a->numtris > end) FAIL ("mdl truncated (verts/tris)");
		if (pass)
		{
			const int *sv = (const int *)p, *tr = (const int *)(p + 12 * (size_t)a->numverts);
			for (i = 0; i < a->numverts; i++)
			{
				a->stverts[4 * i + 0] = sv[3 * i + 0];
				a->stverts[4 * i + 1] = sv[3 * i + 1] << 16;	/* put s and t in 16.16 format */
				a->stverts[4 * i + 2] = sv[3 * i + 2] << 16;
				a->stverts[4 * i + 3] = 0;
			}
			for (i = 0; i < a->numtris; i++)
			{
				for (j = 0; j < 4; j++) a->tris[4 * i + j] = tr[4 * i + j];
				for (j = 1; j < 4; j++)
					if (tr[4 * i + j] < 0 || tr[4 * i + j] >= a->numverts) FAIL ("triangle %d: bad vertex index", i);
			}
		}
		p += 12 * (size_t)a->numverts + 16 * (size_t)a->numtris;
		/* frames (r_model.c:1505-1545) */
		for (i = 0; i < a->numframes; i++)
		{
			int type;
			if (p + 4 > end) FAIL ("mdl truncated (frames)");
			type = *(const int *)p; p += 4;
			if (type == 0)
			{
				if (p + 24 + 4 * (size_t)a->numverts > end) FAIL ("mdl truncated (frame)");
				if (pass)
				{
					a->frames[i].type = 0; a->frames[i].first = pose; a->frames[i].count = 1; a->frames[i].intervals = -1;
					memcpy (a->frames[i].bboxmin, p, 4); memcpy (a->frames[i].bboxmax, p + 4, 4);
					memcpy (a->posebbox + 8 * (size_t)pose, p, 8);
					memcpy (a->poseverts + 4 * (size_t)pose * a->numverts, p + 24, 4 * (size_t)a->numverts);
				}
				p += 24 + 4 * (size_t)a->numverts; pose++;
			}
			else
			{
				int n;
				if (p + 12 > end) FAIL ("mdl truncated (frame group)");
				n = *(const int *)p;
				if (n < 1) FAIL ("bad frame group");
				if (pass)
				{
					a->frames[i].type = 1; a->frames[i].first = pose; a->frames[i].count = n; a->frames[i].intervals = iv;
					memcpy (a->frames[i].bboxmin, p + 4, 4); memcpy (a->frames[i].bboxmax, p + 8, 4);
				}
				p += 12;
				if (p + 4 * (size_t)n > end) FAIL ("mdl truncated (frame group intervals)");
				for (j = 0; j < n; j++)
				{
					float f = *(const float *)(p + 4 * j);
					if (f <= 0.0) FAIL ("Mod_LoadAliasGroup: interval<=0");
					if (pass) a->intervals[iv + j] = f;
				}
				iv += n; p += 4 * (size_t)n;
				for (j = 0; j < n; j++)
				{
					if (p + 24 + 4 * (size_t)a->numverts > end) FAIL ("mdl truncated (group frame)");
					if (pass)
					{
						memcpy (a->posebbox + 8 * (size_t)pose, p, 8);
						memcpy (a->poseverts + 4 * (size_t)pose * a->numverts, p + 24, 4 * (size_t)a->numverts);
					}
					p += 24 + 4 * (size_t)a->numverts; pose++;
				}
			}
		}
		if (!pass)
		{
			nposes = pose; nskins = skin; nintervals = iv;
			a->numposes = nposes; a->numskinpics = nskins;
			a->stverts = (int *)calloc (4 * (size_t)a->numverts, sizeof(int));
			a->tris = (int *)calloc (4 * (size_t)a->numtris, sizeof(int));
			a->poseverts = (unsigned char *)calloc (4 * (size_t)nposes * a->numverts + 4, 1);
			a->posebbox = (unsigned char *)calloc (8 * (size_t)nposes + 8, 1);
			a->skinpixels = (unsigned char *)calloc ((size_t)nskins * skinsize + 16, 1);
			a->frames = (rc_aliasframe_t *)calloc (a->numframes, sizeof(rc_aliasframe_t));
			a->skins = (rc_aliasframe_t *)calloc (a->numskins, sizeof(rc_aliasframe_t));
			a->intervals = (float *)calloc (nintervals + 1, sizeof(float));
		}
	}
	/* model bounds (r_model.c:1549-1553) from the extremes over all frames */
	{
		int mn[3] = {255, 255, 255}, mx[3] = {-255, -255, -255};
		for (i = 0; i < a->numframes; i++)
			for (k = 0; k < 3; k++)
			{
				if (a->frames[i].bboxmin[k] < mn[k]) mn[k] = a->frames[i].bboxmin[k];
				if (a->frames[i].bboxmax[k] > mx[k]) mx[k] = a->frames[i].bboxmax[k];
			}
		for (i = 0; i < nposes; i++)
			for (k = 0; k < 3; k++)
			{
				if (a->posebbox[8 * i + k] < mn[k]) mn[k] = a->posebbox[8 * i + k];
				if (a->posebbox[8 * i + 4 + k] > mx[k]) mx[k] = a->posebbox[8 * i + 4 + k];
			}
		for (k = 0; k < 3; k++)
		{
			a->mins[k] = mn[k] * a->scale[k] + a->scale_origin[k];
			a->maxs[k] = mx[k] * a->scale[k] + a->scale_origin[k];
		}
	}
	a->device_index = -1;
	if (!rc_alias_build_image (a, base, modelname ? modelname : "?"))
		FAIL ("out of memory (model image)");
	return a;
fail:
	RC_FreeHostAlias (a);
	return NULL;
}

/* ------------------------------------------------------------------ light sampling ---- */
/* ---- the reference's in-memory image of a loaded model -------------------------------------------
 * D_PolysetDrawSpans8 indexes the skin with whatever the edge stepping produced: clipped triangles step up to a few
 * texel rows off the picture (d_polyse.c:385-447), and the reference then reads its OWN model memory around the skin --
 * the hunk header in front of it, the skin descriptors, the triangle and st-vertex arrays before that; the next
 * allocation's header and the frame vertices behind it.  Those bytes decide pixels, so the device keeps, per model, the
 * byte image of the block Mod_LoadAliasModel assembles (r_model.c:1349-1572): every Hunk_AllocName (common/zone.c:404:
 * a 16-byte header {0x1df001ed, size, name[8]} + the size rounded up to 16, zero filled) in the reference's order,
 * from `pheader` on, with the LP64 struct layouts of r_model.h:214-264 / modelgen.h:40-100.  Skin reads inside the image
 * are then exact; only reads that leave the model's block altogether are still reported (RC_STAT_ALIAS_RANGE). */
typedef struct { unsigned char *p; size_t len, cap; char name[8]; } rc_img_t;

static size_t img_alloc (rc_img_t *im, size_t size)
{
	/* returns the offset of the block's payload relative to pheader (the first payload is offset 0) */
	const size_t total = 16 + ((size + 15) & ~(size_t)15), at = im->len;
	if (im->len + total > im->cap)
	{
		size_t ncap = (im->len + total) * 2 + 4096;
		unsigned char *np = (unsigned char *)realloc (im->p, ncap);
		if (!np) return (size_t)-1;
		im->p = np; im->cap = ncap;
	}
	memset (im->p + at, 0, total);
	{
		const int sentinal = 0x1df001ed, sz = (int)total;
		memcpy (im->p + at, &sentinal, 4); memcpy (im->p + at + 4, &sz, 4); memcpy (im->p + at + 8, im->name, 8);
	}
	im->len += total;
	return at + 16 - 16;	/* the image starts at pheader = 16 bytes into the first block: offsets are (at + 16) - 16 */
}
#define IMG(im, ofs) ((im)->p + 16 + (ofs))		/* payload byte `ofs` (relative to pheader) */
static void img_i32 (rc_img_t *im, size_t ofs, int v) { memcpy (IMG (im, ofs), &v, 4); }
static void img_f32 (rc_img_t *im, size_t ofs, float v) { memcpy (IMG (im, ofs), &v, 4); }

static int rc_alias_build_image (rc_hostalias_t *a, const unsigned char *base, const char *modelname)
{
	const mdl_t *in = (const mdl_t *)base;
	const unsigned char *p = base + sizeof(mdl_t);
	const int skinsize = a->skinwidth * a->skinheight;
	const size_t FRAMEDESC = 32, SKINDESC = 24, GROUPFRAMEDESC = 16;	/* maliasframedesc_t, maliasskindesc_t (a pointer inside), maliasgroupframedesc_t */
	rc_img_t im;
	size_t hdr, pmodel, stv, tri, sd;
	int i, j, pic = 0;

	memset (&im, 0, sizeof(im));
	{
		/* loadname = COM_FileBase (mod->name), common/common.c:583-603; strncpy (h->name, name, 8) */
		const char *s = modelname + strlen (modelname) - 1, *s2;
		char basename[64];
		while (s != modelname && *s != '.') s--;
		for (s2 = s; s2 > modelname && *s2 != '/'; s2--) ;
		if (*s2 != '/') s2 = modelname - 1;
		if (s - s2 < 2) strcpy (basename, "?model?");
		else { size_t n = (size_t)(s - 1 - s2); if (n > 63) n = 63; memcpy (basename, s2 + 1, n); basename[n] = 0; }
		strncpy (im.name, basename, 8);
	}
	a->skinpicofs = (int *)calloc ((size_t)a->numskinpics + 1, sizeof(int));
	if (!a->skinpicofs) return 0;
	/* the header block: aliashdr_t + frames[numframes] + mdl_t + stverts + triangles */
	hdr = img_alloc (&im, 16 + FRAMEDESC * a->numframes + sizeof(mdl_t) + 12 * (size_t)a->numverts + 16 * (size_t)a->numtris);
	if (hdr == (size_t)-1) return 0;
	pmodel = 16 + FRAMEDESC * a->numframes; stv = pmodel + sizeof(mdl_t); tri = stv + 12 * (size_t)a->numverts;
	img_i32 (&im, 0, (int)pmodel); img_i32 (&im, 4, (int)stv); img_i32 (&im, 12, (int)tri);
	{
		/* the mdl_t fields the loader sets (r_model.c:1392-1426); ident, version, synctype and flags stay 0 */
		mdl_t m;
		memset (&m, 0, sizeof(m));
		m.boundingradius = in->boundingradius; m.numskins = in->numskins; m.skinwidth = in->skinwidth; m.skinheight = in->skinheight;
		m.numverts = in->numverts; m.numtris = in->numtris; m.numframes = in->numframes;
		m.size = in->size * (1.0 / 11.0);		/* ALIAS_BASE_SIZE_RATIO, modelgen.h:34 */
		for (i = 0; i < 3; i++) { m.scale[i] = in->scale[i]; m.scale_origin[i] = in->scale_origin[i]; m.eyeposition[i] = in->eyeposition[i]; }
		memcpy (IMG (&im, pmodel), &m, sizeof(m));
	}
	/* skins: the descriptor array, then per skin its pixels (a group: group struct, intervals, pictures) */
	sd = img_alloc (&im, SKINDESC * a->numskins);
	if (sd == (size_t)-1) return 0;
	img_i32 (&im, 8, (int)sd);
	for (i = 0; i < a->numskins; i++)
	{
		const int type = *(const int *)p; p += 4;
		img_i32 (&im, sd + SKINDESC * i, type);
		if (type == 0)
		{
			const size_t o = img_alloc (&im, skinsize);
			if (o == (size_t)-1) return 0;
			memcpy (IMG (&im, o), p, skinsize); p += skinsize;
			img_i32 (&im, sd + SKINDESC * i + 16, (int)o);
			a->skinpicofs[pic++] = (int)o;
		}
		else
		{
			const int n = *(const int *)p;
			size_t g, iv;
			p += 4;
			g = img_alloc (&im, 8 + SKINDESC * n);		/* maliasskingroup_t + (n - 1) skindescs */
			iv = g == (size_t)-1 ? g : img_alloc (&im, 4 * (size_t)n);
			if (iv == (size_t)-1) return 0;
			img_i32 (&im, sd + SKINDESC * i + 16, (int)g);
			img_i32 (&im, g, n); img_i32 (&im, g + 4, (int)iv);
			for (j = 0; j < n; j++) img_f32 (&im, iv + 4 * j, *(const float *)(p + 4 * j));
			p += 4 * (size_t)n;
			for (j = 0; j < n; j++)
			{
				const size_t o = img_alloc (&im, skinsize);
				if (o == (size_t)-1) return 0;
				memcpy (IMG (&im, o), p, skinsize); p += skinsize;
				img_i32 (&im, g + 8 + SKINDESC * j + 16, (int)o);
				a->skinpicofs[pic++] = (int)o;
			}
		}
	}
	/* st vertices (s and t in 16.16) and triangles */
	for (i = 0; i < a->numverts; i++)
	{
		const int *sv = (const int *)p + 3 * i;
		img_i32 (&im, stv + 12 * i, sv[0]); img_i32 (&im, stv + 12 * i + 4, sv[1] << 16); img_i32 (&im, stv + 12 * i + 8, sv[2] << 16);
	}
	p += 12 * (size_t)a->numverts;
	memcpy (IMG (&im, tri), p, 16 * (size_t)a->numtris);
	p += 16 * (size_t)a->numtris;
	/* frames: descriptor in the header block; vertices (a group: group struct, intervals, vertices per pose) in their own blocks */
	for (i = 0; i < a->numframes; i++)
	{
		const size_t fd = 16 + FRAMEDESC * i;
		const int type = *(const int *)p; p += 4;
		img_i32 (&im, fd, type);
		if (type == 0)
		{
			size_t o;
			strcpy ((char *)IMG (&im, fd + 16), (const char *)(p + 8));	/* strcpy (name, pdaliasframe->name) */
			memcpy (IMG (&im, fd + 4), p, 3); memcpy (IMG (&im, fd + 8), p + 4, 3);	/* bboxmin.v, bboxmax.v (not lightnormalindex) */
			o = img_alloc (&im, 4 * (size_t)a->numverts);
			if (o == (size_t)-1) return 0;
			memcpy (IMG (&im, o), p + 24, 4 * (size_t)a->numverts);
			img_i32 (&im, fd + 12, (int)o);
			p += 24 + 4 * (size_t)a->numverts;
		}
		else
		{
			const int n = *(const int *)p;
			size_t g, iv;
			g = img_alloc (&im, 8 + GROUPFRAMEDESC * n);	/* maliasgroup_t + (n - 1) frames */
			if (g == (size_t)-1) return 0;
			img_i32 (&im, g, n);
			memcpy (IMG (&im, fd + 4), p + 4, 3); memcpy (IMG (&im, fd + 8), p + 8, 3);
			img_i32 (&im, fd + 12, (int)g);
			p += 12;
			iv = img_alloc (&im, 4 * (size_t)n);
			if (iv == (size_t)-1) return 0;
			img_i32 (&im, g + 4, (int)iv);
			for (j = 0; j < n; j++) img_f32 (&im, iv + 4 * j, *(const float *)(p + 4 * j));
			p += 4 * (size_t)n;
			for (j = 0; j < n; j++)
			{
				const size_t gf = g + 8 + GROUPFRAMEDESC * j;
				size_t o;
				img_i32 (&im, gf + 12, n);
				strcpy ((char *)IMG (&im, fd + 16), (const char *)(p + 8));	/* every pose's name lands in the frame's name field */
				memcpy (IMG (&im, gf), p, 3); memcpy (IMG (&im, gf + 4), p + 4, 3);
				o = img_alloc (&im, 4 * (size_t)a->numverts);
				if (o == (size_t)-1) return 0;
				memcpy (IMG (&im, o), p + 24, 4 * (size_t)a->numverts);
				img_i32 (&im, gf + 8, (int)o);
				p += 24 + 4 * (size_t)a->numverts;
			}
		}
	}
	/* the image handed to the device starts at pheader and ends with the last block (the cache copy takes 16 bytes more,
	 * whatever the hunk held behind the model: not part of it) */
	free (a->skinpixels);
	a->skinpixels = (unsigned char *)malloc (im.len - 16 + 16);
	if (!a->skinpixels) { free (im.p); return 0; }
	memcpy (a->skinpixels, im.p + 16, im.len - 16);
	memset (a->skinpixels + im.len - 16, 0, 16);
	a->skinimagelen = (int)(im.len - 16);
	free (im.p);
	return 1;
}

/* a + (b - a) * frac / 16 with the reference's arithmetic shift */
static int lerp16 (int a, int b, int frac) { return (((b - a) * frac) >> 4) + a; }

/* RecursiveLightPoint, r_light.c:156-245 */
static int light_point_r (const rc_world_t *w, const int *lsv, int node, const float *start, const float *end)
{
	const rc_node_t *nd;
	const rc_plane_t *plane;
	float front, back, frac, mid[3];
	int side, r, i;

	if (node < 0)
		return -1;		/* didn't hit anything */
	nd = &w->nodes[node];
	plane = &w->planes[nd->plane];
	front = (plane->type < 3 ? start[plane->type] : start[0]*plane->normal[0] + start[1]*plane->normal[1] + start[2]*plane->normal[2]) - plane->dist;
	back = (plane->type < 3 ? end[plane->type] : end[0]*plane->normal[0] + end[1]*plane->normal[1] + end[2]*plane->normal[2]) - plane->dist;
	side = front < 0;
	if ((back < 0) == side)
		return light_point_r (w, lsv, nd->children[side], start, end);
	frac = front / (front-back);
	/* LerpVectors (start, frac, end, mid): v1 + frac*(v2 - v1), common/mathlib.c */
	mid[0] = start[0] + frac*(end[0] - start[0]);
	mid[1] = start[1] + frac*(end[1] - start[1]);
	mid[2] = start[2] + frac*(end[2] - start[2]);
	r = light_point_r (w, lsv, nd->children[side], start, mid);
	if (r >= 0)
		return r;
	if ((back < 0) == side)
		return -1;
	for (i = 0; i < nd->numsurfaces; i++)
	{
		const rc_face_t *surf = &w->faces[nd->firstsurface + i];
		const rc_texinfo_t *tex;
		int s, t, ds, dt;
		if (surf->flags & RC_SURF_DRAWTILED)
			continue;	/* no lightmaps */
		tex = &w->texinfo[surf->texinfo];
		s = (mid[0]*tex->vecs[0][0] + mid[1]*tex->vecs[0][1] + mid[2]*tex->vecs[0][2]) + tex->vecs[0][3];
		if (s < surf->texturemins[0])
			continue;
		ds = s - surf->texturemins[0];
		if (ds > surf->extents[0])
			continue;
		t = (mid[0]*tex->vecs[1][0] + mid[1]*tex->vecs[1][1] + mid[2]*tex->vecs[1][2]) + tex->vecs[1][3];
		if (t < surf->texturemins[1])
			continue;
		dt = t - surf->texturemins[1];
		if (dt > surf->extents[1])
			continue;
		r = 0;
		if (surf->lightofs != -1)
		{
			/* the four luxels around (ds, dt), each the sum over the face's light styles (accumulated the way the
			 * reference does: an int that takes a float product per style, truncating every time), then a bilinear
			 * blend in 4-bit fixed point: along s on both rows, then along t (r_light.c:213-236) */
			const int luxw = (surf->extents[0] >> 4) + 1, luxh = (surf->extents[1] >> 4) + 1;
			const int corner[4] = { 0, 1, luxw, luxw + 1 };
			const unsigned char *lm = w->lightdata + surf->lightofs + (dt >> 4) * luxw + (ds >> 4);
			int acc[4] = { 0, 0, 0, 0 }, m, c, top, bottom;
			for (m = 0; m < RC_MAXLIGHTMAPS && surf->styles[m] != 255; m++, lm += luxw * luxh)
			{
				const int style = surf->styles[m];
				const float scale = (style < RC_MAX_LIGHTSTYLES ? lsv[style] : 0) * (1.0 / 256.0);
				for (c = 0; c < 4; c++)
					acc[c] += (float)lm[corner[c]] * scale;
			}
			top = lerp16 (acc[0], acc[1], ds & 15);
			bottom = lerp16 (acc[2], acc[3], ds & 15);
			r = lerp16 (top, bottom, dt & 15);
		}
		return r;
	}
	return light_point_r (w, lsv, nd->children[!side], mid, end);
}

int RC_LightPoint (const rc_hostworld_t *hw, const int *lightstylevalue, const float *p)
{
	float end[3];
	int r;
	if (!hw || !hw->w.haslight)
		return 255;
	end[0] = p[0]; end[1] = p[1]; end[2] = p[2] - 2048;
	r = light_point_r (&hw->w, lightstylevalue, hw->submodels[0].headnode, p, end);
	if (r == -1)
		r = 0;
	return r;
}

/* ------------------------------------------------------------------ per-entity setup --- */
static void alias_transform (const rc_hostalias_t *a, const entity_t *e, const rc_view_t *v, const float *modelorg,
			     int trivial_accept, float out[3][4])
{
	/* R_AliasSetUpTransform, r_alias.c:367-414 */
	int i;
	float angles[3], fwd[3], right[3], up[3], rot[3][4], vm[3][4];
	angles[2] = e->angles[2];
	angles[0] = -e->angles[0];
	angles[1] = e->angles[1];
	RC_AngleVectors (angles, fwd, right, up);
	for (i = 0; i < 3; i++)
	{
		rot[i][0] = fwd[i] * a->scale[0];
		rot[i][1] = -right[i] * a->scale[1];
		rot[i][2] = up[i] * a->scale[2];
		rot[i][3] = fwd[i] * a->scale_origin[0] - right[i] * a->scale_origin[1] + up[i] * a->scale_origin[2] - modelorg[i];
	}
	for (i = 0; i < 3; i++) { vm[0][i] = v->vright[i]; vm[1][i] = -v->vup[i]; vm[2][i] = v->vpn[i]; }
	vm[0][3] = vm[1][3] = vm[2][3] = 0;
	/* R_ConcatTransforms (viewmatrix, rotationmatrix, aliastransform), common/mathlib.c:389-403 */
	for (i = 0; i < 3; i++)
	{
		out[i][0] = vm[i][0] * rot[0][0] + vm[i][1] * rot[1][0] + vm[i][2] * rot[2][0];
		out[i][1] = vm[i][0] * rot[0][1] + vm[i][1] * rot[1][1] + vm[i][2] * rot[2][1];
		out[i][2] = vm[i][0] * rot[0][2] + vm[i][1] * rot[1][2] + vm[i][2] * rot[2][2];
		out[i][3] = vm[i][0] * rot[0][3] + vm[i][1] * rot[1][3] + vm[i][2] * rot[2][3] + vm[i][3];
	}
	if (trivial_accept)
	{
		for (i = 0; i < 4; i++)
		{
			out[0][i] *= v->aliasxscale * (1.0 / ((float)0x8000 * 0x10000));
			out[1][i] *= v->aliasyscale * (1.0 / ((float)0x8000 * 0x10000));
			out[2][i] *= 1.0 / ((float)0x8000 * 0x10000);
		}
	}
}

/* pose for a frame at a time: R_AliasFrameVerts / R_AliasFrameBoundingBox, r_alias.c:79-121,663-699.
 * Quirk kept: for a frame GROUP the reference reads numframes through frames[frame] with the
 * MODEL frame index (r_alias.c:110,678) -- identical for every member, so the value is the group size. */
static int alias_pose (const rc_hostalias_t *a, int frame, float time, const unsigned char **mins, const unsigned char **maxs)
{
	const rc_aliasframe_t *f = &a->frames[frame];
	int i = 0;
	if (f->type == 0)
	{
		if (mins) { *mins = f->bboxmin; *maxs = f->bboxmax; }
		return f->first;
	}
	{
		const float *pintervals = a->intervals + f->intervals;
		int numposes = f->count;
		float fullinterval = pintervals[numposes-1];
		float targettime = time - ((int)(time / fullinterval)) * fullinterval;
		for (i = 0; i < (numposes-1); i++)
			if (pintervals[i] > targettime)
				break;
	}
	if (mins) { *mins = a->posebbox + 8 * (size_t)(f->first + i); *maxs = *mins + 4; }
	return f->first + i;
}

static float fminf_ (float a, float b) { return a < b ? a : b; }
static float fmaxf_ (float a, float b) { return a > b ? a : b; }

/* R_AliasCheckBBox, r_alias.c:135-296: -1 rejected, 1 trivially accepted, 0 needs clipping */
static int alias_check_bbox (const rc_hostalias_t *a, const entity_t *e, const rc_view_t *v, const rc_refdef_t *rd,
			     const rc_cvars_t *cv, float xf[3][4])
{
	static const int aedges[12][2] = { {0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,5},{1,4},{2,7},{3,6} };
	int i, flags, frame, oldframe, numv, vflags[16];
	float zi, basepts[8][3], v0, v1, frac, viewaux[16][3];
	int zclipped, zfullyclipped;
	unsigned anyclip, allclip;
	const unsigned char *mins, *maxs, *oldmins, *oldmaxs;

	frame = e->frame; oldframe = e->oldframe;
	if ((frame >= a->numframes) || (frame < 0)) frame = 0;
	if ((oldframe >= a->numframes) || (oldframe < 0)) oldframe = 0;
	if (cv->r_lerpmodels)
	{
		alias_pose (a, frame, rd->time + e->syncbase, &mins, &maxs);
		alias_pose (a, oldframe, rd->oldtime + e->syncbase, &oldmins, &oldmaxs);
		basepts[0][0] = basepts[1][0] = basepts[2][0] = basepts[3][0] = (float)(mins[0] < oldmins[0] ? mins[0] : oldmins[0]);
		basepts[4][0] = basepts[5][0] = basepts[6][0] = basepts[7][0] = (float)(maxs[0] > oldmaxs[0] ? maxs[0] : oldmaxs[0]);
		basepts[0][1] = basepts[3][1] = basepts[5][1] = basepts[6][1] = (float)(mins[1] < oldmins[1] ? mins[1] : oldmins[1]);
		basepts[1][1] = basepts[2][1] = basepts[4][1] = basepts[7][1] = (float)(maxs[1] > oldmaxs[1] ? maxs[1] : oldmaxs[1]);
		basepts[0][2] = basepts[1][2] = basepts[4][2] = basepts[5][2] = (float)(mins[2] < oldmins[2] ? mins[2] : oldmins[2]);
		basepts[2][2] = basepts[3][2] = basepts[6][2] = basepts[7][2] = (float)(maxs[2] > oldmaxs[2] ? maxs[2] : oldmaxs[2]);
	}
	else
	{
		alias_pose (a, frame, rd->time + e->syncbase, &mins, &maxs);
		basepts[0][0] = basepts[1][0] = basepts[2][0] = basepts[3][0] = (float)mins[0];
		basepts[4][0] = basepts[5][0] = basepts[6][0] = basepts[7][0] = (float)maxs[0];
		basepts[0][1] = basepts[3][1] = basepts[5][1] = basepts[6][1] = (float)mins[1];
		basepts[1][1] = basepts[2][1] = basepts[4][1] = basepts[7][1] = (float)maxs[1];
		basepts[0][2] = basepts[1][2] = basepts[4][2] = basepts[5][2] = (float)mins[2];
		basepts[2][2] = basepts[3][2] = basepts[6][2] = basepts[7][2] = (float)maxs[2];
	}
	(void)fminf_; (void)fmaxf_;
	zclipped = 0;
	zfullyclipped = 1;
	for (i = 0; i < 8; i++)
	{
		const float *in = basepts[i];
		viewaux[i][0] = (in[0]*xf[0][0] + in[1]*xf[0][1] + in[2]*xf[0][2]) + xf[0][3];
		viewaux[i][1] = (in[0]*xf[1][0] + in[1]*xf[1][1] + in[2]*xf[1][2]) + xf[1][3];
		viewaux[i][2] = (in[0]*xf[2][0] + in[1]*xf[2][1] + in[2]*xf[2][2]) + xf[2][3];
		if (viewaux[i][2] < ALIAS_Z_CLIP_PLANE)
		{
			vflags[i] = ALIAS_Z_CLIP;
			zclipped = 1;
		}
		else
		{
			vflags[i] = 0;
			zfullyclipped = 0;
		}
	}
	if (zfullyclipped)
		return -1;
	numv = 8;
	if (zclipped)
	{
		for (i = 0; i < 12; i++)
		{
			int i0 = aedges[i][0], i1 = aedges[i][1];
			if (vflags[i0] ^ vflags[i1])
			{
				frac = (ALIAS_Z_CLIP_PLANE - viewaux[i0][2]) / (viewaux[i1][2] - viewaux[i0][2]);
				viewaux[numv][0] = viewaux[i0][0] + (viewaux[i1][0] - viewaux[i0][0]) * frac;
				viewaux[numv][1] = viewaux[i0][1] + (viewaux[i1][1] - viewaux[i0][1]) * frac;
				viewaux[numv][2] = ALIAS_Z_CLIP_PLANE;
				vflags[numv] = 0;
				numv++;
			}
		}
	}
	anyclip = 0;
	allclip = ALIAS_XY_CLIP_MASK;
	for (i = 0; i < numv; i++)
	{
		if (vflags[i] & ALIAS_Z_CLIP)
			continue;
		zi = 1.0 / viewaux[i][2];
		v0 = (viewaux[i][0] * v->xscale * zi) + v->xcenter;
		v1 = (viewaux[i][1] * v->yscale * zi) + v->ycenter;
		flags = 0;
		if (v0 < (float)v->vrect_x) flags |= ALIAS_LEFT_CLIP;
		if (v1 < (float)v->vrect_y) flags |= ALIAS_TOP_CLIP;
		if (v0 > (float)v->vrectright) flags |= ALIAS_RIGHT_CLIP;
		if (v1 > (float)v->vrectbottom) flags |= ALIAS_BOTTOM_CLIP;
		anyclip |= flags;
		allclip &= flags;
	}
	if (allclip)
		return -1;
	return !anyclip & !zclipped;
}

/* R_VisCullEntity, r_main.c:457-483, against the view leaf's precomputed visible set */
static int vis_cull_r (const rc_world_t *w, const uint32_t *vis, int node, const float *mins, const float *maxs)
{
	for (;;)
	{
		const rc_node_t *nd;
		const rc_plane_t *p;
		float dist1, dist2;
		int sides, i;
		if (node < 0)
		{
			int li = -1 - node;
			if (w->leafs[li].contents == RC_CONTENTS_SOLID) return 0;
			return (vis[(w->numnodes + li) >> 5] >> ((w->numnodes + li) & 31)) & 1;
		}
		if (!((vis[node >> 5] >> (node & 31)) & 1)) return 0;
		nd = &w->nodes[node];
		p = &w->planes[nd->plane];
		/* BOX_ON_PLANE_SIDE, common/mathlib.h + BoxOnPlaneSide (mathlib.c:178-231) */
		if (p->type < 3)
			sides = (p->dist <= mins[p->type]) ? 1 : ((p->dist >= maxs[p->type]) ? 2 : 3);
		else
		{
			dist1 = dist2 = 0;
			for (i = 0; i < 3; i++)
			{
				if (p->normal[i] < 0) { dist1 += p->normal[i]*mins[i]; dist2 += p->normal[i]*maxs[i]; }
				else { dist1 += p->normal[i]*maxs[i]; dist2 += p->normal[i]*mins[i]; }
			}
			sides = 0;
			if (dist1 >= p->dist) sides = 1;
			if (dist2 < p->dist) sides |= 2;
		}
		if (sides == 3)
		{
			if (vis_cull_r (w, vis, nd->children[0], mins, maxs)) return 1;
			node = nd->children[1];
		}
		else if (sides == 1) node = nd->children[0];
		else node = nd->children[1];
	}
}

int RC_VisCullEntity (const rc_hostworld_t *hw, const float *vieworg, const float *mins, const float *maxs)
{
	const rc_world_t *w = &hw->w;
	int n = hw->submodels[0].headnode;
	while (n >= 0)
	{
		const rc_node_t *nd = &w->nodes[n];
		const rc_plane_t *pl = &w->planes[nd->plane];
		float d = (vieworg[0]*pl->normal[0] + vieworg[1]*pl->normal[1] + vieworg[2]*pl->normal[2]) - pl->dist;
		n = (d > 0) ? nd->children[0] : nd->children[1];
	}
	return vis_cull_r (w, w->nodevis + (size_t)(-1 - n) * w->visrowwords, hw->submodels[0].headnode, mins, maxs);
}

/* Returns 1 when the entity is drawn and *out is filled, 0 when it is rejected. */
int RC_AliasSetup (const rc_hostalias_t *a, const rc_hostworld_t *hw, const entity_t *e, const rc_refdef_t *rd,
		   const rc_view_t *v, const rc_cvars_t *cv, rc_aliasinst_t *out)
{
	float modelorg[3], xf[3][4];
	int trivial_accept, i, r, lnum, skinnum, frame, oldframe;
	int ambient;
	float shade, lerp;

	for (i = 0; i < 3; i++)
		modelorg[i] = v->origin[i] - e->origin[i];	/* r_main.c:524 */
	/* R_DrawAliasModel, r_alias.c:734-788 */
	if (e->flags & RF_WEAPONMODEL)
		trivial_accept = 0;
	else
	{
		alias_transform (a, e, v, modelorg, 0, xf);
		trivial_accept = alias_check_bbox (a, e, v, rd, cv, xf);
		if (trivial_accept == -1)
			return 0;
	}
	/* R_AliasSetupSkin, r_alias.c:539-590 */
	skinnum = e->skinnum;
	if ((skinnum >= a->numskins) || (skinnum < 0)) skinnum = 0;
	{
		const rc_aliasframe_t *sk = &a->skins[skinnum];
		int pic = sk->first;
		if (sk->type == 1)
		{
			const float *iv = a->intervals + sk->intervals;
			int numskins = sk->count;
			float fullskininterval = iv[numskins-1];
			float skintime = (a->synctype == 1 /* ST_RAND */) ? rd->time + e->syncbase : rd->time;
			float skintargettime = skintime - ((int)(skintime / fullskininterval)) * fullskininterval;
			for (i = 0; i < (numskins-1); i++)
				if (iv[i] > skintargettime)
					break;
			pic = sk->first + i;
		}
		out->skin = a->skinpicofs[pic];	/* offset of the picture inside the model's image (the device adds the image's base) */
	}
	alias_transform (a, e, v, modelorg, trivial_accept, xf);
	memcpy (out->xf, xf, sizeof(xf));

	/* R_AliasSetupLighting, r_alias.c:597-656 */
	if (e->flags & RF_FULLBRIGHT)
		r = 255;
	else
		r = RC_LightPoint (hw, v->lightstylevalue, e->origin);
	if (r < v->ambientlight)
		r = v->ambientlight;
	ambient = r; shade = r;
	if ((e->flags & RF_MINLIGHT) && ambient < 24)
	{
		ambient = 24; shade = 24;
	}
	for (lnum = 0; lnum < rd->num_dlights; lnum++)
	{
		float dist[3], add;
		dist[0] = e->origin[0] - rd->dlights[lnum].origin[0];
		dist[1] = e->origin[1] - rd->dlights[lnum].origin[1];
		dist[2] = e->origin[2] - rd->dlights[lnum].origin[2];
		add = rd->dlights[lnum].radius - sqrt (dist[0]*dist[0] + dist[1]*dist[1] + dist[2]*dist[2]);
		if (add > 0)
		{
			ambient += add;
			shade += add;
		}
	}
	if (ambient > 128)
		ambient = 128;
	if (ambient + shade > 192)
		shade = 192 - ambient;
	if (ambient < LIGHT_MIN)
		ambient = LIGHT_MIN;
	ambient = (255 - ambient) << 6;
	if (ambient < LIGHT_MIN)
		ambient = LIGHT_MIN;
	if (shade < 0)
		shade = 0;
	shade *= ((100.0 / 255.0) * 64);
	out->ambientlight = ambient;
	out->shadelight = shade;
	out->shadequant = ((int)(e->angles[1] * (16 / 360.0))) & 15;

	/* R_AliasSetupFrame, r_alias.c:704-727 */
	frame = e->frame; oldframe = e->oldframe;
	if ((frame >= a->numframes) || (frame < 0)) frame = 0;
	if ((oldframe >= a->numframes) || (oldframe < 0)) oldframe = 0;
	out->pose_new = alias_pose (a, frame, rd->time + e->syncbase, NULL, NULL);
	out->pose_old = alias_pose (a, oldframe, rd->oldtime + e->syncbase, NULL, NULL);
	if (!cv->r_lerpmodels)
		lerp = 1.0f;
	else
	{
		/* bound (0, e->framelerp, 1) = max (0, min (framelerp, 1)) */
		lerp = e->framelerp < 1 ? e->framelerp : 1;
		lerp = 0 > lerp ? 0 : lerp;
	}
	out->framelerp = lerp;
	out->ziscale = (e->flags & RF_DEPTHSCALE) ? (float)0x8000 * (float)0x10000 * 3.0 : (float)0x8000 * (float)0x10000;
	out->trivial_accept = trivial_accept;
	return 1;
}
