/*
 * rc_pipeline.h -- the per-thread bodies of the ref_cuda kernels.
 *
 * Every function here is the work of ONE CUDA thread (one view, one face, one scanline,
 * one cache texel, one 8-pixel block ...).  rc_gpu.cu wraps them in __global__ kernels;
 * tests/hostsim compiles the same bodies with g++ to step through the pipeline on a CPU
 * while developing (test infrastructure only -- the product never runs them on the host).
 *
 * Arithmetic rules: float expressions are written with the
 * reference's operand types and order; the file is compiled with --fmad=false /
 * -ffp-contract=off; float->int goes through rc_f2i/rc_d2i, which reproduce x86
 * cvttss2si/cvttsd2si (0x80000000 when out of range) instead of CUDA's saturation.
 */
#ifndef RC_PIPELINE_H
#define RC_PIPELINE_H

#include "rc_types.h"
#include <math.h>

#ifdef __CUDACC__
#define RC_HD __device__ __forceinline__
#define RC_HD2 __device__ __forceinline__
#define RC_DEBUG_PRINT(...) do { } while (0)
#define RC_LDG(p) __ldg (p)
#define RC_LDCG(p) __ldcg (p)   /* L2 read: values other threads of the CTA updated with atomics */
#define RC_ATOMIC_ADD_I(p, v)   atomicAdd ((p), (v))
#define RC_ATOMIC_ADD_ULL(p, v) atomicAdd ((p), (v))
#define RC_ATOMIC_OR_I(p, v)    atomicOr ((p), (v))
#define RC_ATOMIC_MIN_I(p, v)   atomicMin ((p), (v))
/* the resolve planes only ever grow: a fragment that does not beat the value read from L2 cannot beat the final one either,
 * so most of an overdrawn view's fragments cost one read instead of one atomic */
#define RC_ATOMIC_MAX_ULL(p, v) do { unsigned long long *p_ = (p); const unsigned long long v_ = (v); if (v_ > __ldcg (p_)) atomicMax (p_, v_); } while (0)
#define RC_ATOMIC_MAX_ULL_BLIND(p, v) atomicMax ((p), (v))
#define RC_ATOMIC_CAS_ULL(p, c, v) atomicCAS ((p), (c), (v))
#else
#define RC_HD static inline
#define RC_HD2 inline
#include <stdio.h>
#include <stdlib.h>
#define RC_DEBUG_PRINT(...) do { if (getenv ("RC_DEBUG")) fprintf (stderr, __VA_ARGS__); } while (0)
#define RC_LDG(p) (*(p))
#define RC_LDCG(p) (*(p))
struct rc_u2 { uint32_t x, y; };
struct rc_u4 { uint32_t x, y, z, w; };
#define uint2 rc_u2
#define uint4 rc_u4
static inline int rc_host_add_i (int *p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned long long rc_host_add_ull (unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline int rc_host_or_i (int *p, int v) { int o = *p; *p = o | v; return o; }
static inline unsigned long long rc_host_cas_ull (unsigned long long *p, unsigned long long c, unsigned long long v) { unsigned long long o = *p; if (o == c) *p = v; return o; }
#define RC_ATOMIC_ADD_I(p, v)   rc_host_add_i ((p), (v))
#define RC_ATOMIC_ADD_ULL(p, v) rc_host_add_ull ((p), (v))
#define RC_ATOMIC_OR_I(p, v)    rc_host_or_i ((p), (v))
#define RC_ATOMIC_MIN_I(p, v)   do { if ((v) < *(p)) *(p) = (v); } while (0)
#define RC_ATOMIC_MAX_ULL(p, v) do { if ((v) > *(p)) *(p) = (v); } while (0)
#define RC_ATOMIC_MAX_ULL_BLIND(p, v) RC_ATOMIC_MAX_ULL (p, v)
#define RC_ATOMIC_CAS_ULL(p, c, v) rc_host_cas_ull ((p), (c), (v))
#endif

/* a place where the reference itself reads or loops out of range (RC_STAT_ALIAS_RANGE): the status bit, and a count of
 * the events in status word 1 so that tests can pin how often it happens */
#define RC_ALIAS_RANGE_HIT(f) do { RC_ATOMIC_OR_I ((f)->status, RC_STAT_ALIAS_RANGE); RC_ATOMIC_ADD_I ((f)->status + 1, 1); } while (0)

#define RC_FACEPOS_NONE    0xFFFFu   /* face not marked by any visited leaf */
#define RC_FACEPOS_MARKED  0xFFFEu   /* marked (surf->visframe == r_framecount), not rendered */
#define RC_MAX_FACE_EDGES  30        /* per face; slot 31 is the synthetic left-clip edge */
#define RC_WALK_STACK      512
#ifndef RC_AET_MAX
#define RC_AET_MAX         256       /* active edges per scanline the scan kernel holds (shared memory) */
#endif
#define RC_STACK_MAX       96        /* active surface stack depth held by the scan thread */
#define RC_CACHE_EMPTY     0xFFFFFFFFFFFFFFFFull
#define RC_WARP_W          320       /* WARP_WIDTH / WARP_HEIGHT, d_iface.h:22-23 */
#define RC_WARP_H          200
#define RC_SEG_BND         9         /* boundaries 8s .. 8s+8 of a span are evaluated by segment s (rc_span_seg) */

/* per-batch buffers (device pointers inside kernels, host pointers in the simulator) */
typedef struct {
    int nviews, width, height;
    int viewbase;                 /* batch index of views[0] (entities / particles carry batch view indices) */
    const rc_view_t *views;
    const rc_dlight_t *dlights;   /* all views' dynamic lights, indexed from rc_view_t.dlight_ofs */
    /* BSP walk */
    rc_facerec_t *facelist; int facecap; int *nfaces;
    uint16_t *facepos;            /* [nviews][numfaces]: list position of a rendered face, else RC_FACEPOS_NONE */
    int *facemark;                /* [nviews][numfaces]: smallest key of a visited leaf that marks the face */
    rc_nodevisit_t *nodevisit;    /* [nviews][numnodes] */
    int *viewleaf;                /* [nviews] */
    /* face stage */
    rc_edge_t *edges; int edgecap; int *nedges;
    rc_surf_t *surfs; int surfcap; /* [nviews][surfcap]; 0 dummy, 1 background, 2+p face-list slot p */
    int maxedges, maxsurfs;       /* the reference's r_maxedges / r_maxsurfs (overflow reporting) */
    /* scan */
    rc_span_t *spans; int spancap;
    rc_cell_t *cellspan;          /* [nviews][height][(width+7)/8]: the span that covers ALL eight pixels of a screen cell, si -1: none */
    int *segspan; int segcap;     /* [segcap]: owning span of every 64-pixel span segment (8 subdivisions) */
    rc_bnd_t *bnd;                /* [segcap * 8 + 8]: boundary b of a span is record segbase * 8 + b (rc_types.h rc_bnd_t) */
#ifdef __CUDACC__
    int2 *slowcells;              /* GPU only: {span, first pixel} of the full cells k_draw queues for the general code; count in alloc[3] */
#else
    void *slowcells;
#endif
    int slowmode;                 /* GPU only: 0 k_draw queues the cells it leaves to the general code (k_slowcells draws them),
                                   * 1 it skips them and k_ends finds them itself (grouped drawing: no per-group launches) */
    unsigned long long *alloc;    /* [0]: lo32 spans used, hi32 segments used  [1]: cache pool cursor  [2]: jobs */
    /* surface cache */
    unsigned long long *cachekeys; int *cachevals; int cachemask;
    uint8_t *cachepool; long long cachecap;
    rc_cachejob_t *jobs; int jobcap;
    /* outputs */
    uint8_t *fb; int16_t *zb;
    int drawview0, drawview1;     /* views [drawview0, drawview1) are drawn by this launch of the draw kernel */
    uint8_t *warpbuf;             /* [nwarp][320*200]: r_warpbuffer of the views that render under water */
    /* alias models and particles */
    const rc_aliasinst_t *alias; int nalias;
    const rc_aliasmodel_t *amodels;
    const int *stverts;           /* [..][4] */
    const int *atris;             /* [..][4] */
    const uint8_t *poseverts;     /* trivertx_t */
    const uint8_t *skinpixels;
    const float *anormdots;       /* r_avertexnormal_dots[16][256] */
    const uint8_t *colormaps;     /* [ncolormaps][16384]; 0 = vid.colormap */
    rc_finalvert_t *finalverts;
    const rc_particle_t *particles; int nparticles;
    const rc_spriteinst_t *sprites; int nsprites;   /* drawn sprite entities */
    const uint8_t *spritepixels;  /* frames of every loaded sprite model */
    /* brush entities */
    const rc_bent_t *bents; int nbents;
    int bentbase;                 /* first brush entity handled by this launch */
    int *leafkey;                 /* [nviews][numleafs]: BSP key of every leaf the walk reached (mleaf_t.key) */
    int *nbsurfs;                 /* [nviews]: surf_t slots handed to brush-entity fragments */
    int bsurfcap;                 /* per view; their slots start at facecap + 2 */
    unsigned long long *zres;     /* [nzres][height*width]: max over fragments of (z, draw order, colour) */
    struct rc_aspan_s *aspans; long long aspancap;   /* GPU: queued spans of alias triangles (k_alias_tris -> k_alias_spans); alloc[5] = triangles << RC_AQ_SHIFT | spans */
    struct rc_aspantri_s *aspantris; int aspantricap;   /* the constants their triangles share */
    int countfrags;               /* GPU, profiling mode: the alias span kernel counts its fragments (alloc[9..11]) */
    int *cacheunits; int cacheunitcap;   /* GPU: {job, first row} of the row bands the cache build works on; count in alloc[8] */
    int *status;
} rc_frame_t;

/* ------------------------------------------------------------------ scalar helpers ---- */
RC_HD int rc_f2i (float x)
{
	/* cvttss2si: truncation, "integer indefinite" (0x80000000) when out of range or NaN;
	 * |x| < 2^31 is false for NaN, and x == -2^31 maps to INT_MIN either way */
	return (fabsf (x) < 2147483648.0f) ? (int)x : (int)0x80000000;
}
RC_HD int rc_d2i (double x)
{
	/* cvttsd2si */
	if (!(x > -2147483649.0 && x < 2147483648.0))
		return (int)0x80000000;
	return (int)x;
}
RC_HD int rc_sar (int x, int s) { return x >> s; }   /* arithmetic on every target we build for */
RC_HD int rc_addw (int a, int b) { return (int)((unsigned)a + (unsigned)b); }   /* wrapping add */
RC_HD int rc_mulw (int a, int b) { return (int)((unsigned)a * (unsigned)b); }

RC_HD float rc_dot3 (const float *x, const float *y) { return x[0]*y[0] + x[1]*y[1] + x[2]*y[2]; }

RC_HD void rc_transform_vector (const rc_view_t *v, const float *in, float *out)
{	/* TransformVector, r_misc.c:308-313 */
	out[0] = rc_dot3 (in, v->vright);
	out[1] = rc_dot3 (in, v->vup);
	out[2] = rc_dot3 (in, v->vpn);
}

/* ------------------------------------------------------------------ stage 1: walk ----- */
RC_HD int rc_visbit (const uint32_t *row, int idx) { return (row[idx >> 5] >> (idx & 31)) & 1; }

/* bbox vs frustum, r_bsp.c:453-491.  Returns 0 when rejected, else 1 with *pclip updated. */
RC_HD int rc_cull_box (const rc_view_t *v, const int16_t *minmaxs, int *pclip)
{
	int clipflags = *pclip;
	if (clipflags)
	{
		for (int i = 0; i < 4; i++)
		{
			if (!(clipflags & (1 << i)))
				continue;
			const int *pindex = v->frustum_idx[i];
			float rejectpt[3], acceptpt[3];
			rejectpt[0] = (float)minmaxs[pindex[0]];
			rejectpt[1] = (float)minmaxs[pindex[1]];
			rejectpt[2] = (float)minmaxs[pindex[2]];
			double d = rc_dot3 (rejectpt, v->clipnormal[i]) - v->clipdist[i];
			if (d <= 0)
				return 0;
			acceptpt[0] = (float)minmaxs[pindex[3+0]];
			acceptpt[1] = (float)minmaxs[pindex[3+1]];
			acceptpt[2] = (float)minmaxs[pindex[3+2]];
			d = rc_dot3 (acceptpt, v->clipnormal[i]) - v->clipdist[i];
			if (d >= 0)
				clipflags &= ~(1 << i);
		}
	}
	*pclip = clipflags;
	return 1;
}

RC_HD float rc_plane_diff (const float *p, const rc_plane_t *pl)
{	/* PlaneDiff, common/mathlib.h:46 */
	return (pl->type < 3 ? p[pl->type] : rc_dot3 (p, pl->normal)) - pl->dist;
}

/* One thread per view: the per-view resets R_BeginEdgeFrame does (r_edge.c:81-112).  The view
 * leaf (Mod_PointInLeaf, r_model.c:83-105) comes from the host with the other view constants:
 * a serial pointer chase has no business on the device. */
RC_HD void rc_view_begin (const rc_world_t *w, const rc_frame_t *f, int vi)
{
	const rc_view_t *v = &f->views[vi];
	rc_surf_t *bg = f->surfs + (size_t)vi * f->surfcap + 1;
	/* surfaces[1]: the background */
	bg->key = 0x7FFFFFFF; bg->flags = RC_SURF_DRAWBACKGROUND; bg->face = -1; bg->insubmodel = 0;
	bg->hasspans = 0; bg->entity = -1; bg->nearzi = 0;
	bg->d_ziorigin = 0; bg->d_zistepu = 0; bg->d_zistepv = 0;
	f->nedges[vi] = 0;
	f->nfaces[vi] = 0;
	if (f->nbsurfs)
		f->nbsurfs[vi] = 0;
	f->viewleaf[vi] = v->viewleaf >= 0 && v->viewleaf < w->numleafs ? v->viewleaf : 0;
}

/* The BSP walk, R_RecursiveWorldNode (r_bsp.c:436-557), in two parallel steps.
 *
 * rc_node_test, one thread per (view, node): everything the recursion computes AT a node that does
 * not depend on the path taken to it: for each of the four frustum planes whether the node's box is
 * rejected (bits 0-3) or fully accepted (bits 4-7) -- the reference only evaluates planes still set
 * in the inherited clipflags, which the walk below re-applies --, the side of the node's plane the
 * view is on (bit 8) and the three-way front/back/on-plane state (bits 9-10).  Bit 15: the node is in
 * the PVS-marked set of the view leaf (precomputed R_MarkLeaves row).
 *
 * rc_visit_walk, one thread per (view, node-or-leaf x): decides whether the recursion reaches x by
 * replaying the root->x descent over those masks: every ancestor must survive the reject test with
 * the clipflags it inherits.  The BSP sort key is accumulated on the way down from static subtree
 * slot counts: front subtree first, then the node's faces, then the back subtree -- the same order
 * in which the reference increments r_currentkey, so keys compare identically.
 *   visited node -> f->nodevisit[x] (side, clipflags, key of its first face)
 *   visited leaf -> atomicMin of the leaf's key into f->facemark[] for its marksurfaces
 *                   (surf->visframe = r_framecount happens when the leaf is reached, so a face
 *                   is drawn only if some leaf with a SMALLER key marked it). */
RC_HD uint16_t rc_node_test (const rc_world_t *w, const rc_frame_t *f, int vi, int n)
{
	const rc_view_t *v = &f->views[vi];
	if (v->rdflags & 1 /* RDF_NOWORLDMODEL */)
		return 0;
	const uint32_t *vis = w->nodevis + (size_t)f->viewleaf[vi] * w->visrowwords;
	if (!rc_visbit (vis, n))
		return 0;
	const rc_node_t *nd = &w->nodes[n];
	int m = 0x8000;
	for (int i = 0; i < 4; i++)
	{
		int one = 1 << i;
		if (!rc_cull_box (v, nd->minmaxs, &one))
			m |= 1 << i;		/* rejectpt behind plane i */
		else if (!one)
			m |= 16 << i;		/* acceptpt in front: the whole box is inside plane i */
	}
	double dot = rc_plane_diff (v->origin, &w->planes[nd->plane]);
	if (dot < 0) m |= 0x100;
	m |= (dot > RC_BACKFACE_EPSILON ? 1 : (dot < -RC_BACKFACE_EPSILON ? 2 : 3)) << 9;
	return (uint16_t)m;
}

RC_HD void rc_visit_walk (const rc_world_t *w, const rc_frame_t *f, int vi, int x, const uint16_t *masks)
{
	const rc_view_t *v = &f->views[vi];
	if (v->rdflags & 1 /* RDF_NOWORLDMODEL */)
		return;
	const int is_leaf = x >= w->numnodes;
	const int L = x - w->numnodes;
	int tx;
	if (is_leaf)
	{
		if (L == 0 || w->leafs[L].contents == RC_CONTENTS_SOLID)
			return;
		tx = w->leafdfs[L];
		if (tx < 0)
			return;
		const uint32_t *vis = w->nodevis + (size_t)f->viewleaf[vi] * w->visrowwords;
		if (!rc_visbit (vis, x))
			return;
	}
	else
	{
		if (!(masks[x] & 0x8000))
			return;
		tx = w->nodeaux[x].dfs_in;
	}
	int cur = 0, clipflags = 15, key = 0;
	for (;;)
	{
		const int m = masks[cur];
		if (m & clipflags & 15)
			return;				/* an ancestor's box is outside a plane still being tested */
		clipflags &= ~((m >> 4) & 15);
		const int side = (m >> 8) & 1;
		const rc_node_t *nd = &w->nodes[cur];
		const rc_nodeaux_t aux = w->nodeaux[cur];
		if (!is_leaf && cur == x)
		{
			rc_nodevisit_t *nv = &f->nodevisit[(size_t)vi * w->numnodes + x];
			nv->clipflags = (int16_t)clipflags;
			nv->facekey = key + aux.slots[side];
			nv->state = (int16_t)((m >> 9) & 3);
			return;
		}
		const int c = (tx < aux.dfs_mid) ? 0 : 1;
		if (c != side)
			key += aux.slots[side] + nd->numsurfaces + 1;
		const int child = nd->children[c];
		if (child < 0)
		{
			const rc_leaf_t *lf = &w->leafs[L];
			if (!rc_cull_box (v, lf->minmaxs, &clipflags))
				return;
			f->leafkey[(size_t)vi * w->numleafs + L] = key;	/* mleaf_t.key; rc_mark_face stamps the leaf's faces */
			if (!w->marksplit)
			{
				int *facemark = f->facemark + (size_t)vi * w->numfaces;
				for (int q = 0; q < lf->nummarks; q++)
					RC_ATOMIC_MIN_I (&facemark[w->marksurfaces[lf->firstmark + q]], key);
			}
			return;
		}
		cur = child;
	}
}

/* One thread per (view, marksurfaces entry): the marking loop of a visited leaf (r_bsp.c:500-512),
 * taken out of the leaf's thread: a room-sized leaf marks a hundred faces. */
RC_HD void rc_mark_face (const rc_world_t *w, const rc_frame_t *f, int vi, int m)
{
	const int leaf = w->markleaf[m];
	if (leaf < 0)
		return;
	const int key = RC_LDCG (&f->leafkey[(size_t)vi * w->numleafs + leaf]);
	if (key != 0x7FFFFFFF)
		RC_ATOMIC_MIN_I (&f->facemark[(size_t)vi * w->numfaces + w->marksurfaces[m]], key);
}

/* One thread per (view, world face): the face loop of R_RecursiveWorldNode (r_bsp.c:520-549):
 * right side of its node's plane, marked by an earlier leaf -> R_RenderFace is called on it. */
RC_HD void rc_list_face (const rc_world_t *w, const rc_frame_t *f, int vi, int fi, int *nfaces, int *skykey)
{
	const rc_face_t *fa = &w->faces[fi];
	if (fa->node < 0)
		return;
	const rc_nodevisit_t nv = f->nodevisit[(size_t)vi * w->numnodes + fa->node];
	if (nv.state != 1 && nv.state != 2)
		return;
	const int want = nv.state == 2 ? RC_SURF_PLANEBACK : 0;
	if ((fa->flags & RC_SURF_PLANEBACK) != want)
		return;
	const int key = nv.facekey + fa->nodeslot;
	if (!(RC_LDCG (&f->facemark[(size_t)vi * w->numfaces + fi]) < key))
		return;
	if ((fa->flags & RC_SURF_DRAWSKY) && w->drawskybox)
		RC_ATOMIC_MIN_I (skykey, key);		/* the first sky face in BSP order triggers R_EmitSkyBox (r_rast.c:530-533) */
	int p = RC_ATOMIC_ADD_I (nfaces, 1);	/* the view's counter: in shared memory on the GPU (same-address L2 atomics serialise) */
	if (p >= f->facecap || p >= (int)RC_FACEPOS_MARKED - 1)
	{
		RC_ATOMIC_OR_I (f->status, RC_STAT_FACELIST);
		return;
	}
	rc_facerec_t *rec = f->facelist + (size_t)vi * f->facecap + p;
	rec->face = fi; rec->clipflags = nv.clipflags; rec->key = key; rec->pad = 0;
	f->facepos[(size_t)vi * w->numfaces + fi] = (uint16_t)p;
}

/* R_EmitSkyBox (r_rast.c:165-205): when a sky face was reached and a sky box is loaded, the six box
 * faces go through R_RenderFace with all four clip planes and the keys 0x7ffffff0.. (behind every
 * world face, in front of the background).  The reference numbers only the faces that emit; fixed
 * keys 0x7ffffff0 + i compare the same.  rec.pad = BSP key of the triggering sky face = the box's
 * place in the emission order (edge rank).  One thread per box face, i = 0..5. */
RC_HD void rc_list_skybox (const rc_world_t *w, const rc_frame_t *f, int vi, int i, int nlisted, int skykey)
{
	const int p = nlisted + i;
	if (p >= f->facecap)
	{
		RC_ATOMIC_OR_I (f->status, RC_STAT_FACELIST);
		return;
	}
	rc_facerec_t *rec = f->facelist + (size_t)vi * f->facecap + p;
	rec->face = w->skyfirst + i; rec->clipflags = 15; rec->key = RC_SKYBOX_KEY + i; rec->pad = skykey;
	f->facepos[(size_t)vi * w->numfaces + w->skyfirst + i] = (uint16_t)p;
}

/* ------------------------------------------------------------- stage 2: face -> edges -- */
typedef struct {
	float u, v, lzi;
	int   ceilv;
} rc_evert_t;

/* transform and project one end point, r_rast.c:236-262 */
RC_HD const rc_xform_t *rc_view_xform (const rc_view_t *vw) { return (const rc_xform_t *)(const void *)vw->vpn; }

RC_HD void rc_project (const rc_view_t *vw, const rc_xform_t *xf, const float *world, rc_evert_t *o)
{
	float local[3], t[3];
	local[0] = world[0] - xf->origin[0];
	local[1] = world[1] - xf->origin[1];
	local[2] = world[2] - xf->origin[2];
	t[0] = rc_dot3 (local, xf->vright);
	t[1] = rc_dot3 (local, xf->vup);
	t[2] = rc_dot3 (local, xf->vpn);
	if (t[2] < RC_NEAR_CLIP)
		t[2] = RC_NEAR_CLIP;
	float lzi = 1.0 / t[2];
	float scale = vw->xscale * lzi;
	float u = (vw->xcenter + scale*t[0]);
	if (u < vw->fvrectx_adj) u = vw->fvrectx_adj;
	if (u > vw->fvrectright_adj) u = vw->fvrectright_adj;
	scale = vw->yscale * lzi;
	float vv = (vw->ycenter - scale*t[1]);
	if (vv < vw->fvrecty_adj) vv = vw->fvrecty_adj;
	if (vv > vw->fvrectbottom_adj) vv = vw->fvrectbottom_adj;
	o->u = u; o->v = vv; o->lzi = lzi;
	o->ceilv = rc_d2i (ceil ((double)vv));
}

#define RC_CLS_PARTIAL   0   /* cacheoffset = 0x7FFFFFFF: clipped, never cached            */
#define RC_CLS_FULLCLIP  1   /* FULLY_CLIPPED_CACHED: second user skips the edge           */
#define RC_CLS_CACHED    2   /* emitted unclipped: second user shares the edge_t           */

typedef struct {
	int   cls;
	int   emitted;        /* R_EmitEdge reached (r_emitted = 1), even for horizontal: no */
	int   made_edge;      /* an edge_t was produced */
	int   trailing;       /* side == 0 */
	int   u, u_step, v, v2;
	float nearzi;         /* edge->nearzi (max of the two 1/z), valid when R_EmitEdge ran */
	int   reached_emit;
	int   leftclipped, rightclipped;
	int   have_leftenter, have_leftexit, have_rightenter, have_rightexit;
	float leftenter[3], leftexit[3], rightenter[3], rightexit[3];
} rc_clipres_t;

/* R_ClipEdge (r_rast.c:391-485, tail recursion unrolled) followed by R_EmitEdge
 * (r_rast.c:213-383) for one polygon edge p0->p1 against the planes in `mask`
 * (bit i = view_clipplanes[i], tested in ascending i).  Pure: returns what the reference
 * would have written to its globals. */
RC_HD void rc_clip_emit (const rc_view_t *vw, const rc_xform_t *xf, const float *p0in, const float *p1in, int mask,
			 int nearzionly, rc_clipres_t *r)
{
	float p0[3], p1[3];
	int clipped = 0;
	p0[0] = p0in[0]; p0[1] = p0in[1]; p0[2] = p0in[2];
	p1[0] = p1in[0]; p1[1] = p1in[1]; p1[2] = p1in[2];
	r->cls = RC_CLS_CACHED;
	r->emitted = 0; r->made_edge = 0; r->reached_emit = 0; r->nearzi = 0;
	r->leftclipped = r->rightclipped = 0;
	r->have_leftenter = r->have_leftexit = r->have_rightenter = r->have_rightexit = 0;

	for (int i = 0; i < 4; i++)
	{
		if (!(mask & (1 << i)))
			continue;
		const float *nrm = xf->clipnormal[i];
		float d0 = rc_dot3 (p0, nrm) - xf->clipdist[i];
		float d1 = rc_dot3 (p1, nrm) - xf->clipdist[i];
		if (d0 >= 0)
		{
			if (d1 >= 0)
				continue;
			/* only point 1 is clipped */
			clipped = 1;
			float fr = d0 / (d0 - d1);
			float c[3];
			c[0] = p0[0] + fr * (p1[0] - p0[0]);
			c[1] = p0[1] + fr * (p1[1] - p0[1]);
			c[2] = p0[2] + fr * (p1[2] - p0[2]);
			if (i == 0) { r->leftclipped = 1; r->have_leftexit = 1; r->leftexit[0] = c[0]; r->leftexit[1] = c[1]; r->leftexit[2] = c[2]; }
			else if (i == 1) { r->rightclipped = 1; r->have_rightexit = 1; r->rightexit[0] = c[0]; r->rightexit[1] = c[1]; r->rightexit[2] = c[2]; }
			p1[0] = c[0]; p1[1] = c[1]; p1[2] = c[2];
		}
		else
		{
			if (d1 < 0)
			{
				/* both points are clipped; cached as fully clipped unless the left plane cut it */
				r->cls = (!r->leftclipped) ? RC_CLS_FULLCLIP : RC_CLS_PARTIAL;
				return;
			}
			clipped = 1;
			float fr = d0 / (d0 - d1);
			float c[3];
			c[0] = p0[0] + fr * (p1[0] - p0[0]);
			c[1] = p0[1] + fr * (p1[1] - p0[1]);
			c[2] = p0[2] + fr * (p1[2] - p0[2]);
			if (i == 0) { r->leftclipped = 1; r->have_leftenter = 1; r->leftenter[0] = c[0]; r->leftenter[1] = c[1]; r->leftenter[2] = c[2]; }
			else if (i == 1) { r->rightclipped = 1; r->have_rightenter = 1; r->rightenter[0] = c[0]; r->rightenter[1] = c[1]; r->rightenter[2] = c[2]; }
			p0[0] = c[0]; p0[1] = c[1]; p0[2] = c[2];
		}
	}
	if (clipped)
		r->cls = RC_CLS_PARTIAL;

	/* R_EmitEdge */
	rc_evert_t e0, e1;
	rc_project (vw, xf, p0, &e0);
	rc_project (vw, xf, p1, &e1);
	float lzi0 = e0.lzi;
	if (e1.lzi > lzi0)
		lzi0 = e1.lzi;
	r->nearzi = lzi0;
	r->reached_emit = 1;
	if (nearzionly)
		return;
	r->emitted = 1;
	if (e0.ceilv == e1.ceilv)
	{
		/* horizontal edge: unclipped ones are cached as fully clipped */
		if (r->cls != RC_CLS_PARTIAL)
			r->cls = RC_CLS_FULLCLIP;
		return;
	}
	int side = e0.ceilv > e1.ceilv;
	float u, u_step;
	r->made_edge = 1;
	if (side == 0)
	{
		r->trailing = 1;
		r->v = e0.ceilv;
		r->v2 = e1.ceilv - 1;
		u_step = ((e1.u - e0.u) / (e1.v - e0.v));
		u = e0.u + ((float)r->v - e0.v) * u_step;
	}
	else
	{
		r->trailing = 0;
		r->v2 = e0.ceilv - 1;
		r->v = e1.ceilv;
		u_step = ((e0.u - e1.u) / (e0.v - e1.v));
		u = e1.u + ((float)r->v - e1.v) * u_step;
	}
	r->u_step = rc_f2i (u_step*0x100000);
	r->u = rc_f2i (u*0x100000 + 0xFFFFF);
	if (r->u < vw->vrect_x_adj_shift20)
		r->u = vw->vrect_x_adj_shift20;
	if (r->u > vw->vrectright_adj_shift20)
		r->u = vw->vrectright_adj_shift20;
}

/* R_RenderFace (r_rast.c:518-716) in two steps so that a face's edges can be clipped and projected
 * by different threads:
 *   rc_face_edge    one polygon edge: the edge-cache decision (partner face's use of the same medge),
 *                   R_ClipEdge + R_EmitEdge; everything it would have written to the reference's
 *                   globals goes into an rc_edgeres_t
 *   rc_face_fold    folds the edge results in winding order (later edges overwrite r_leftenter & co.
 *                   exactly as the globals would be), clips the left-clip edge, runs the right-clip
 *                   nearzi pass and counts the edge records the face emits
 *   rc_face_emit    writes them at the slot the caller reserved, and the surf_t */
#define RC_ER_SKIP      1	/* partner cached it as fully clipped: nothing happens */
#define RC_ER_CACHED    2	/* partner's edge_t carries this surface (R_EmitCachedEdge) */
#define RC_ER_EMITTED   4
#define RC_ER_MADE      8	/* an edge_t comes out */
#define RC_ER_LEFTCLIP  16
#define RC_ER_RIGHTCLIP 32
#define RC_ER_LE        64	/* pl is r_leftenter */
#define RC_ER_LX        128	/* pl is r_leftexit */
#define RC_ER_RE        256	/* pr is r_rightenter */
#define RC_ER_RX        512	/* pr is r_rightexit */
#define RC_ER_REACHED   1024	/* R_EmitEdge ran: nearzi valid */
typedef struct {
	int   flags;
	float nearzi;
	int   u, u_step;
	int   vv;		/* v | v2 << 16 */
	int   surfs;		/* surfs[0] | surfs[1] << 16 */
	float pl[3], pr[3];	/* crossing point with the left / right clip plane */
} rc_edgeres_t;			/* 48 B */

RC_HD void rc_face_edge (const rc_world_t *w, const rc_frame_t *f, int vi, int p, int i, rc_edgeres_t *out)
{
	const rc_view_t *vw = &f->views[vi];
	const rc_facerec_t *list = f->facelist + (size_t)vi * f->facecap;
	const uint16_t *facepos = f->facepos + (size_t)vi * w->numfaces;
	const rc_facerec_t rec = list[p];
	const rc_face_t *fa = &w->faces[rec.face];
	const int surfidx = p + 2;
	rc_clipres_t r;
	rc_fedge_t fe;
#ifdef __CUDACC__
	{
		const uint4 *q4 = reinterpret_cast<const uint4 *>(&w->fedges[fa->firstedge + i]);
		const uint4 x = __ldg (q4), y = __ldg (q4 + 1);
		fe.v0[0] = __uint_as_float (x.x); fe.v0[1] = __uint_as_float (x.y); fe.v0[2] = __uint_as_float (x.z); fe.partner_face = (int)x.w;
		fe.v1[0] = __uint_as_float (y.x); fe.v1[1] = __uint_as_float (y.y); fe.v1[2] = __uint_as_float (y.z); fe.partner_rev = (int)y.w;
	}
#else
	fe = w->fedges[fa->firstedge + i];
#endif
	out->flags = 0; out->nearzi = 0;
	if ((fa->flags & RC_SURF_DRAWSKY) && w->drawskybox)
	{
		out->flags = RC_ER_SKIP;		/* r_rast.c:530-533: the sky box is drawn instead */
		return;
	}
	if (rec.face >= w->skyfirst)
	{
		/* the box follows the viewer: r_skyverts = r_origin + box_verts * 128 (r_rast.c:177-180) */
		for (int j = 0; j < 3; j++)
		{
			fe.v0[j] = vw->origin[j] + fe.v0[j]*128;
			fe.v1[j] = vw->origin[j] + fe.v1[j]*128;
		}
	}
	const float *va = fe.v0, *vb = fe.v1;
	int q = RC_FACEPOS_NONE;
	if (fe.partner_face >= 0)
		q = facepos[fe.partner_face];
	if (w->drawskybox && q < (int)RC_FACEPOS_MARKED && (w->faces[fe.partner_face].flags & RC_SURF_DRAWSKY))
		q = RC_FACEPOS_NONE;			/* a sky face leaves the edge cache alone when the box is drawn in its place */

	if (q < (int)RC_FACEPOS_MARKED && list[q].key < rec.key)
	{
		/* the partner face ran first: redo its R_ClipEdge (same medge, its direction, its
		 * clipflags) to learn what it left in medge_t.cachededgeoffset (r_rast.c:583-603) */
		const float *pa = fe.partner_rev ? vb : va, *pb = fe.partner_rev ? va : vb;
		rc_clip_emit (vw, rc_view_xform (vw), pa, pb, list[q].clipflags, 0, &r);
		if (r.cls == RC_CLS_FULLCLIP)
		{
			out->flags = RC_ER_SKIP;
			return;
		}
		if (r.cls == RC_CLS_CACHED)
		{
			/* R_EmitCachedEdge: the partner's edge_t carries this surface */
			out->flags = RC_ER_CACHED;
			out->nearzi = r.nearzi;
			return;
		}
	}

	rc_clip_emit (vw, rc_view_xform (vw), va, vb, rec.clipflags, 0, &r);
	int fl = 0;
	if (r.have_leftenter) { fl |= RC_ER_LE; out->pl[0] = r.leftenter[0]; out->pl[1] = r.leftenter[1]; out->pl[2] = r.leftenter[2]; }
	if (r.have_leftexit) { fl |= RC_ER_LX; out->pl[0] = r.leftexit[0]; out->pl[1] = r.leftexit[1]; out->pl[2] = r.leftexit[2]; }
	if (r.have_rightenter) { fl |= RC_ER_RE; out->pr[0] = r.rightenter[0]; out->pr[1] = r.rightenter[1]; out->pr[2] = r.rightenter[2]; }
	if (r.have_rightexit) { fl |= RC_ER_RX; out->pr[0] = r.rightexit[0]; out->pr[1] = r.rightexit[1]; out->pr[2] = r.rightexit[2]; }
	if (r.leftclipped) fl |= RC_ER_LEFTCLIP;
	if (r.rightclipped) fl |= RC_ER_RIGHTCLIP;
	if (r.reached_emit) { fl |= RC_ER_REACHED; out->nearzi = r.nearzi; }
	if (r.emitted) fl |= RC_ER_EMITTED;
	if (r.made_edge)
	{
		int s0 = r.trailing ? surfidx : 0, s1 = r.trailing ? 0 : surfidx;
		fl |= RC_ER_MADE;
		if (r.cls == RC_CLS_CACHED && q < (int)RC_FACEPOS_MARKED && list[q].key > rec.key)
		{
			/* the partner will find this edge cached and add itself (R_EmitCachedEdge) */
			if (r.trailing) s1 = q + 2;
			else s0 = q + 2;
		}
		out->u = r.u; out->u_step = r.u_step;
		out->vv = (int)(((uint32_t)(uint16_t)(int16_t)r.v) | ((uint32_t)(uint16_t)(int16_t)r.v2 << 16));
		out->surfs = (int)((uint32_t)s0 | ((uint32_t)s1 << 16));
		if (r.trailing) fl |= (int)0x80000000;
	}
	out->flags = fl;
}

typedef struct {		/* what rc_face_fold hands to rc_face_emit */
	int   neb;		/* edge_t records the face emits */
	int   emitted;
	int   haveleftedge, lu, lu_step, lv, lv2, ltrailing;	/* the left-clip edge (r_rast.c:661-666) */
	float nearzi;
} rc_facefold_t;

RC_HD void rc_face_fold (const rc_world_t *w, const rc_frame_t *f, int vi, int p, const rc_edgeres_t *res, rc_facefold_t *st)
{
	const rc_view_t *vw = &f->views[vi];
	const rc_facerec_t *list = f->facelist + (size_t)vi * f->facecap;
	const rc_facerec_t rec = list[p];
	const rc_face_t *fa = &w->faces[rec.face];
	rc_surf_t *surf = f->surfs + (size_t)vi * f->surfcap + (p + 2);
	const int clipflags = rec.clipflags;
	const int ne = fa->numedges;
	int emitted = 0, makeleft = 0, makeright = 0, neb = 0;
	float nearzi = 0;
	float leftenter[3] = {0,0,0}, leftexit[3] = {0,0,0}, rightenter[3] = {0,0,0}, rightexit[3] = {0,0,0};
	int have_le = 0, have_lx = 0, have_re = 0, have_rx = 0;
	rc_clipres_t r;

	for (int i = 0; i < ne; i++)
	{
		const rc_edgeres_t *e = &res[i];
		const int fl = e->flags;
		if (fl & RC_ER_SKIP)
			continue;
		if (fl & RC_ER_CACHED)
		{
			if (e->nearzi > nearzi) nearzi = e->nearzi;
			emitted = 1;
			continue;
		}
		if (fl & RC_ER_LE) { have_le = 1; leftenter[0] = e->pl[0]; leftenter[1] = e->pl[1]; leftenter[2] = e->pl[2]; }
		if (fl & RC_ER_LX) { have_lx = 1; leftexit[0] = e->pl[0]; leftexit[1] = e->pl[1]; leftexit[2] = e->pl[2]; }
		if (fl & RC_ER_RE) { have_re = 1; rightenter[0] = e->pr[0]; rightenter[1] = e->pr[1]; rightenter[2] = e->pr[2]; }
		if (fl & RC_ER_RX) { have_rx = 1; rightexit[0] = e->pr[0]; rightexit[1] = e->pr[1]; rightexit[2] = e->pr[2]; }
		if (fl & RC_ER_LEFTCLIP) makeleft = 1;
		if (fl & RC_ER_RIGHTCLIP) makeright = 1;
		if ((fl & RC_ER_REACHED) && e->nearzi > nearzi) nearzi = e->nearzi;
		if (fl & RC_ER_EMITTED) emitted = 1;
		if (fl & RC_ER_MADE) neb++;
	}

	st->haveleftedge = 0;
	if (makeleft)
	{
		/* r_rast.c:661-666: R_ClipEdge (&r_leftexit, &r_leftenter, pclip->next) */
		int first = 0;
		while (!(clipflags & (1 << first))) first++;
		int mask = clipflags & ~((2 << first) - 1);
		if (!have_le || !have_lx)
			RC_ATOMIC_OR_I (f->status, RC_STAT_STALE_CLIPVERT);
		rc_clip_emit (vw, rc_view_xform (vw), leftexit, leftenter, mask, 0, &r);
		if (r.have_rightenter) { have_re = 1; rightenter[0] = r.rightenter[0]; rightenter[1] = r.rightenter[1]; rightenter[2] = r.rightenter[2]; }
		if (r.have_rightexit) { have_rx = 1; rightexit[0] = r.rightexit[0]; rightexit[1] = r.rightexit[1]; rightexit[2] = r.rightexit[2]; }
		if (r.reached_emit && r.nearzi > nearzi)
			nearzi = r.nearzi;
		if (r.emitted)
			emitted = 1;
		if (r.made_edge)
		{
			st->haveleftedge = 1;
			st->lu = r.u; st->lu_step = r.u_step; st->lv = r.v; st->lv2 = r.v2; st->ltrailing = r.trailing;
			neb++;
		}
	}
	surf->flags = -1;		/* not posted unless an edge comes out */
	surf->hasspans = 0;
	/* r_rightenter / r_rightexit are globals in the reference: a face whose own edges wrote
	 * only one of them (the other crossing edge was skipped as FULLY_CLIPPED_CACHED) reads the
	 * value an EARLIER face left behind.  Publish what this face wrote; rc_face_fixup resolves
	 * such reads once every face of the view has run. */
	surf->d_sdivzstepu = rightenter[0]; surf->d_tdivzstepu = rightenter[1]; surf->d_sdivzstepv = rightenter[2];
	surf->d_tdivzstepv = rightexit[0]; surf->d_sdivzorigin = rightexit[1]; surf->d_tdivzorigin = rightexit[2];
	surf->miplevel = have_re | (have_rx << 1);
	if (makeright)
	{
		/* r_rast.c:669-675: only the effect on r_nearzi is wanted; planes after the right one */
		if (have_re && have_rx)
		{
			rc_clip_emit (vw, rc_view_xform (vw), rightexit, rightenter, clipflags & 0xC, 1, &r);
			if (r.reached_emit && r.nearzi > nearzi)
				nearzi = r.nearzi;
		}
		else
			surf->miplevel |= 4;
	}
	st->neb = neb; st->emitted = emitted; st->nearzi = nearzi;
}

/* `base`: the face's first slot in the view's edge array (the caller reserved st->neb records) */
RC_HD void rc_face_emit (const rc_world_t *w, const rc_frame_t *f, int vi, int p, const rc_edgeres_t *res, const rc_facefold_t *st, int base)
{
	const rc_view_t *vw = &f->views[vi];
	const rc_facerec_t *list = f->facelist + (size_t)vi * f->facecap;
	const rc_facerec_t rec = list[p];
	const rc_face_t *fa = &w->faces[rec.face];
	rc_surf_t *surf = f->surfs + (size_t)vi * f->surfcap + (p + 2);
	const int surfidx = p + 2;
	const int ne = fa->numedges;
	/* place in the reference's emission order (newedges[] tie order): a world face's BSP key; the six
	 * box faces are emitted where the first sky face was met, five slots each */
	const int isbox = rec.face >= w->skyfirst;
	const uint32_t orderkey = (uint32_t)(isbox ? rec.pad : rec.key);
	const int slot0 = isbox ? (rec.face - w->skyfirst) * 5 : 0;
	if (st->neb)
	{
		if (base + st->neb > f->edgecap)
			RC_ATOMIC_OR_I (f->status, RC_STAT_EDGE_OVERFLOW);
		else
		{
			rc_edge_t *dst = f->edges + (size_t)vi * f->edgecap + base;
			for (int i = 0; i < ne; i++)
			{
				const rc_edgeres_t *e = &res[i];
				if ((e->flags & (RC_ER_SKIP | RC_ER_CACHED)) || !(e->flags & RC_ER_MADE))
					continue;
				const uint32_t trailing = (uint32_t)e->flags >> 31;
				dst->u = e->u; dst->u_step = e->u_step;
				dst->v = (int16_t)(e->vv & 0xFFFF); dst->v2 = (int16_t)((uint32_t)e->vv >> 16);
				dst->surfs[0] = (uint16_t)(e->surfs & 0xFFFF); dst->surfs[1] = (uint16_t)((uint32_t)e->surfs >> 16);
				dst->rank = (trailing << 31) | (orderkey << 5) | (uint32_t)(slot0 + i);
				dst->pad[0] = dst->pad[1] = dst->pad[2] = 0;
				dst++;
			}
			if (st->haveleftedge)
			{
				dst->u = st->lu; dst->u_step = st->lu_step;
				dst->v = (int16_t)st->lv; dst->v2 = (int16_t)st->lv2;
				dst->surfs[0] = st->ltrailing ? surfidx : 0;
				dst->surfs[1] = st->ltrailing ? 0 : surfidx;
				dst->rank = ((uint32_t)st->ltrailing << 31) | (orderkey << 5) | (isbox ? (uint32_t)(slot0 + 4) : 31u);
				dst->pad[0] = dst->pad[1] = dst->pad[2] = 0;
			}
		}
	}
	if (!st->emitted)
		return;

	/* r_rast.c:693-714 */
	const rc_plane_t *pplane = &w->planes[fa->plane];
	float p_normal[3];
	rc_transform_vector (vw, pplane->normal, p_normal);
	float planedist = pplane->dist;
	if (isbox)
		planedist = vw->origin[pplane->type] + pplane->dist;	/* r_rast.c:183-187: r_origin[axis] +- 128 */
	float distinv = 1.0 / (planedist - rc_dot3 (vw->origin, pplane->normal));
	/* r_draworder (R_BeginEdgeFrame, r_edge.c:94-105 / R_LeadingEdgeBackwards :290-366): the background gets the SMALLEST
	 * key and the surface with the LARGEST key is on top of the stack.  Negated keys under the same rule -- smallest on
	 * top, background 0x7FFFFFFF -- give the same order, so the scan stage is shared */
	surf->key = (vw->rdflags & RC_VIEW_DRAWORDER) ? -rec.key : rec.key;
	surf->flags = fa->flags;
	surf->face = rec.face;
	surf->insubmodel = 0;
	surf->entity = -1;
	surf->nearzi = st->nearzi;
	surf->d_zistepu = p_normal[0] * vw->xscaleinv * distinv;
	surf->d_zistepv = -p_normal[1] * vw->yscaleinv * distinv;
	surf->d_ziorigin = p_normal[2] * distinv - vw->xcenter * surf->d_zistepu - vw->ycenter * surf->d_zistepv;
}

/* One thread per (view, face-list slot): both steps serially (host simulation). */
RC_HD void rc_render_face (const rc_world_t *w, const rc_frame_t *f, int vi, int p, int *nedges)
{
	const rc_facerec_t *list = f->facelist + (size_t)vi * f->facecap;
	const rc_face_t *fa = &w->faces[list[p].face];
	rc_surf_t *surf = f->surfs + (size_t)vi * f->surfcap + (p + 2);
	rc_edgeres_t res[RC_MAX_FACE_EDGES];
	if (fa->numedges > RC_MAX_FACE_EDGES)
	{
		surf->flags = -1;
		surf->hasspans = 0;
		RC_ATOMIC_OR_I (f->status, RC_STAT_EDGE_OVERFLOW);
		return;
	}
	for (int i = 0; i < fa->numedges; i++)
		rc_face_edge (w, f, vi, p, i, &res[i]);
	rc_facefold_t st;
	rc_face_fold (w, f, vi, p, res, &st);
	rc_face_emit (w, f, vi, p, res, &st, st.neb ? RC_ATOMIC_ADD_I (nedges, st.neb) : 0);
}

/* second half of the fixup: best_re / best_rx = list slots of the latest earlier writers (-1: none) */
RC_HD void rc_face_fixup_finish (const rc_world_t *w, const rc_frame_t *f, int vi, int p, int best_re, int best_rx)
{
	rc_surf_t *surfs = f->surfs + (size_t)vi * f->surfcap;
	rc_surf_t *surf = &surfs[p + 2];
	const rc_view_t *vw = &f->views[vi];
	const rc_facerec_t *list = f->facelist + (size_t)vi * f->facecap;
	float re[3], rx[3];
	int have_re = surf->miplevel & 1, have_rx = (surf->miplevel >> 1) & 1;
	re[0] = surf->d_sdivzstepu; re[1] = surf->d_tdivzstepu; re[2] = surf->d_sdivzstepv;
	rx[0] = surf->d_tdivzstepv; rx[1] = surf->d_sdivzorigin; rx[2] = surf->d_tdivzorigin;
	if (!have_re && best_re >= 0)
	{ const rc_surf_t *o = &surfs[best_re + 2]; have_re = 1; re[0] = o->d_sdivzstepu; re[1] = o->d_tdivzstepu; re[2] = o->d_sdivzstepv; }
	if (!have_rx && best_rx >= 0)
	{ const rc_surf_t *o = &surfs[best_rx + 2]; have_rx = 1; rx[0] = o->d_tdivzstepv; rx[1] = o->d_sdivzorigin; rx[2] = o->d_tdivzorigin; }
	if (!have_re || !have_rx)
	{
		/* nothing wrote it this frame: the reference would use last frame's leftovers */
		RC_ATOMIC_OR_I (f->status, RC_STAT_STALE_CLIPVERT);
		return;
	}
	if (surf->flags == -1)
		return;
	rc_clipres_t r;
	rc_clip_emit (vw, rc_view_xform (vw), rx, re, list[p].clipflags & 0xC, 1, &r);
	if (r.reached_emit && r.nearzi > surf->nearzi)
		surf->nearzi = r.nearzi;
}

RC_HD void rc_face_fixup (const rc_world_t *w, const rc_frame_t *f, int vi, int p)
{
	const rc_surf_t *surfs = f->surfs + (size_t)vi * f->surfcap;
	const int m0 = surfs[p + 2].miplevel;
	if (!(m0 & 4))
		return;
	const rc_facerec_t *list = f->facelist + (size_t)vi * f->facecap;
	/* the face list is unordered: take, per value, the writer with the largest key below ours */
	const int mykey = list[p].key, nf = f->nfaces[vi] < f->facecap ? f->nfaces[vi] : f->facecap;
	int best_re = -1, best_rx = -1, kre = -1, krx = -1;
	for (int q = 0; q < nf; q++)
	{
		const int k = list[q].key;
		if (k >= mykey)
			continue;
		const int m = surfs[q + 2].miplevel;
		if ((m & 1) && k > kre) { kre = k; best_re = q; }
		if ((m & 2) && k > krx) { krx = k; best_rx = q; }
	}
	rc_face_fixup_finish (w, f, vi, p, best_re, best_rx);
}

/* ------------------------------------------------------------- stage 2b: brush entities -- */
#define RC_BMODEL_VERTS 96      /* reference: MAX_BMODEL_VERTS 500 / MAX_BMODEL_EDGES 1000 per face (r_bsp.c:41-42) */
#define RC_BMODEL_EDGES 192
#define RC_BMODEL_SEQ_BASE (1 << 25)

typedef struct {
	const rc_world_t *w;
	const rc_frame_t *f;
	const rc_view_t *vw;
	const rc_bent_t *be;
	int vi;			/* chunk-local view index */
	int bi;			/* index of the brush entity in f->bents */
	int face;		/* world face index of psurf */
	int facei;		/* index of the face inside the submodel */
	int frag;		/* R_RenderBmodelFace / R_RenderFace calls made for this face so far */
	float bverts[RC_BMODEL_VERTS][3];
	int numbverts;
} rc_bctx_t;

RC_HD const float *rc_bvert (const rc_bctx_t *c, int idx)
{
	return idx < 0 ? &c->w->vertexes[3 * (-1 - idx)] : c->bverts[idx];
}

/* R_RenderBmodelFace (r_rast.c:723-826) and, for entities that lie in a single leaf,
 * R_RenderFace with insubmodel set (r_rast.c:518-716, no edge cache): pushes `n` polygon edges
 * through the entity's clip planes and posts one surf_t whose key is the leaf's key. */
RC_HD void rc_bmodel_emit (rc_bctx_t *c, const int *ev0, const int *ev1, int n, int clipflags, int key)
{
	const rc_frame_t *f = c->f;
	const rc_view_t *vw = c->vw;
	const rc_xform_t *xf = &c->be->x;
	const rc_face_t *fa = &c->w->faces[c->face];
	const int frag = c->frag++;
	if (frag > 255 || n > RC_MAX_FACE_EDGES)
	{
		RC_ATOMIC_OR_I (f->status, RC_STAT_BMODEL_OVERFLOW);
		return;
	}
	int slot = RC_ATOMIC_ADD_I (&f->nbsurfs[c->vi], 1);
	if (slot >= f->bsurfcap)
	{
		RC_ATOMIC_OR_I (f->status, RC_STAT_SURF_OVERFLOW);
		return;
	}
	const int surfidx = f->facecap + 2 + slot;
	rc_surf_t *surf = f->surfs + (size_t)c->vi * f->surfcap + surfidx;
	surf->flags = -1;
	surf->hasspans = 0;
	surf->miplevel = 0;
	const uint32_t seq = (uint32_t)RC_BMODEL_SEQ_BASE + (((uint32_t)c->be->order & 127) << 18) + (((uint32_t)c->facei & 1023) << 8) + (uint32_t)frag;
	if (c->be->order > 127 || c->facei > 1023)
		RC_ATOMIC_OR_I (f->status, RC_STAT_BMODEL_OVERFLOW);
	rc_edge_t eb[RC_MAX_FACE_EDGES + 1];
	int neb = 0, emitted = 0, makeleft = 0, makeright = 0;
	float nearzi = 0;
	float leftenter[3] = {0,0,0}, leftexit[3] = {0,0,0}, rightenter[3] = {0,0,0}, rightexit[3] = {0,0,0};
	int have_le = 0, have_lx = 0, have_re = 0, have_rx = 0;
	rc_clipres_t r;
	for (int i = 0; i < n; i++)
	{
		rc_clip_emit (vw, xf, rc_bvert (c, ev0[i]), rc_bvert (c, ev1[i]), clipflags, 0, &r);
		if (r.have_leftenter) { have_le = 1; leftenter[0] = r.leftenter[0]; leftenter[1] = r.leftenter[1]; leftenter[2] = r.leftenter[2]; }
		if (r.have_leftexit) { have_lx = 1; leftexit[0] = r.leftexit[0]; leftexit[1] = r.leftexit[1]; leftexit[2] = r.leftexit[2]; }
		if (r.have_rightenter) { have_re = 1; rightenter[0] = r.rightenter[0]; rightenter[1] = r.rightenter[1]; rightenter[2] = r.rightenter[2]; }
		if (r.have_rightexit) { have_rx = 1; rightexit[0] = r.rightexit[0]; rightexit[1] = r.rightexit[1]; rightexit[2] = r.rightexit[2]; }
		if (r.leftclipped) makeleft = 1;
		if (r.rightclipped) makeright = 1;
		if (r.reached_emit && r.nearzi > nearzi) nearzi = r.nearzi;
		if (r.emitted) emitted = 1;
		if (r.made_edge)
		{
			rc_edge_t *e = &eb[neb++];
			e->u = r.u; e->u_step = r.u_step; e->v = (int16_t)r.v; e->v2 = (int16_t)r.v2;
			e->surfs[0] = r.trailing ? surfidx : 0;
			e->surfs[1] = r.trailing ? 0 : surfidx;
			e->rank = ((uint32_t)r.trailing << 31) | (seq << 5) | (uint32_t)i;
			e->pad[0] = e->pad[1] = e->pad[2] = 0;
		}
	}
	if (makeleft)
	{
		int first = 0;
		while (!(clipflags & (1 << first))) first++;
		if (!have_le || !have_lx)
			RC_ATOMIC_OR_I (f->status, RC_STAT_STALE_CLIPVERT);
		rc_clip_emit (vw, xf, leftexit, leftenter, clipflags & ~((2 << first) - 1), 0, &r);
		if (r.have_rightenter) { have_re = 1; rightenter[0] = r.rightenter[0]; rightenter[1] = r.rightenter[1]; rightenter[2] = r.rightenter[2]; }
		if (r.have_rightexit) { have_rx = 1; rightexit[0] = r.rightexit[0]; rightexit[1] = r.rightexit[1]; rightexit[2] = r.rightexit[2]; }
		if (r.reached_emit && r.nearzi > nearzi) nearzi = r.nearzi;
		if (r.emitted) emitted = 1;
		if (r.made_edge)
		{
			rc_edge_t *e = &eb[neb++];
			e->u = r.u; e->u_step = r.u_step; e->v = (int16_t)r.v; e->v2 = (int16_t)r.v2;
			e->surfs[0] = r.trailing ? surfidx : 0;
			e->surfs[1] = r.trailing ? 0 : surfidx;
			e->rank = ((uint32_t)r.trailing << 31) | (seq << 5) | 31u;
			e->pad[0] = e->pad[1] = e->pad[2] = 0;
		}
	}
	if (makeright)
	{
		if (!have_re || !have_rx)
			RC_ATOMIC_OR_I (f->status, RC_STAT_STALE_CLIPVERT);
		else
		{
			rc_clip_emit (vw, xf, rightexit, rightenter, clipflags & 0xC, 1, &r);
			if (r.reached_emit && r.nearzi > nearzi) nearzi = r.nearzi;
		}
	}
	if (neb)
	{
		int base = RC_ATOMIC_ADD_I (&f->nedges[c->vi], neb);
		if (base + neb > f->edgecap)
			RC_ATOMIC_OR_I (f->status, RC_STAT_EDGE_OVERFLOW);
		else
		{
			rc_edge_t *dst = f->edges + (size_t)c->vi * f->edgecap + base;
			for (int i = 0; i < neb; i++)
				dst[i] = eb[i];
		}
	}
	if (!emitted)
		return;
	const rc_plane_t *pplane = &c->w->planes[fa->plane];
	float p_normal[3];
	p_normal[0] = rc_dot3 (pplane->normal, xf->vright);
	p_normal[1] = rc_dot3 (pplane->normal, xf->vup);
	p_normal[2] = rc_dot3 (pplane->normal, xf->vpn);
	float distinv = 1.0 / (pplane->dist - rc_dot3 (xf->origin, pplane->normal));
	surf->key = (vw->rdflags & RC_VIEW_DRAWORDER) ? -key : key;
	surf->flags = fa->flags;
	surf->face = c->face;
	surf->insubmodel = 1;
	surf->entity = c->bi;
	surf->nearzi = nearzi;
	surf->d_zistepu = p_normal[0] * vw->xscaleinv * distinv;
	surf->d_zistepv = -p_normal[1] * vw->yscaleinv * distinv;
	surf->d_ziorigin = p_normal[2] * distinv - vw->xcenter * surf->d_zistepu - vw->ycenter * surf->d_zistepv;
}

/* One thread per (brush entity, face): the face loops of R_DrawSubmodelPolygons (r_bsp.c:401-427)
 * and R_DrawSolidClippedSubmodelPolygons (r_bsp.c:319-394) with R_RecursiveClipBPoly
 * (r_bsp.c:161-311): the polygon is split by the world BSP planes below the entity's top node so
 * that every fragment inherits the sort key of the world leaf it lies in. */
RC_HD void rc_bmodel_face (const rc_world_t *w, const rc_frame_t *f, int bi, int facei)
{
	const rc_bent_t *be = &f->bents[bi];
	if (facei >= be->numfaces)
		return;
	const int vi = be->view - f->viewbase;
	const rc_view_t *vw = &f->views[vi];
	const int face = be->firstface + facei;
	const rc_face_t *psurf = &w->faces[face];
	const rc_plane_t *pplane = &w->planes[psurf->plane];
	const float dot = rc_plane_diff (be->x.origin, pplane);
	if (!(((psurf->flags & RC_SURF_PLANEBACK) && (dot < -RC_BACKFACE_EPSILON)) ||
	      (!(psurf->flags & RC_SURF_PLANEBACK) && (dot > RC_BACKFACE_EPSILON))))
		return;
	if (psurf->numedges > RC_MAX_FACE_EDGES || psurf->numedges < 1)
	{
		RC_ATOMIC_OR_I (f->status, RC_STAT_BMODEL_OVERFLOW);
		return;
	}
	rc_bctx_t c;
	c.w = w; c.f = f; c.vw = vw; c.be = be; c.vi = vi; c.bi = bi; c.face = face; c.facei = facei; c.frag = 0; c.numbverts = 0;
	const int *leafkey = f->leafkey + (size_t)vi * w->numleafs;
	int ev0[RC_MAX_FACE_EDGES + 2], ev1[RC_MAX_FACE_EDGES + 2];

	if (be->topnode < 0)
	{
		/* falls entirely in one leaf: all edges go in, 1/z sorting handles the order */
		for (int j = 0; j < psurf->numedges; j++)
		{
			const int lindex = w->surfedges[psurf->firstedge + j];
			const int m = lindex > 0 ? lindex : -lindex;
			ev0[j] = -1 - (int)w->edges[2 * m + (lindex > 0 ? 0 : 1)];
			ev1[j] = -1 - (int)w->edges[2 * m + (lindex > 0 ? 1 : 0)];
		}
		rc_bmodel_emit (&c, ev0, ev1, psurf->numedges, be->clipflags, leafkey[-1 - be->topnode]);
		return;
	}

	/* bedge_t lists as index-linked arrays (r_local.h:31-35) */
	int e_v0[RC_BMODEL_EDGES], e_v1[RC_BMODEL_EDGES], e_next[RC_BMODEL_EDGES];
	int numbedges = psurf->numedges;
	for (int j = 0; j < psurf->numedges; j++)
	{
		const int lindex = w->surfedges[psurf->firstedge + j];
		const int m = lindex > 0 ? lindex : -lindex;
		e_v0[j] = -1 - (int)w->edges[2 * m + (lindex > 0 ? 0 : 1)];
		e_v1[j] = -1 - (int)w->edges[2 * m + (lindex > 0 ? 1 : 0)];
		e_next[j] = j + 1;
	}
	e_next[psurf->numedges - 1] = -1;
	const uint32_t *vis = w->nodevis + (size_t)f->viewleaf[vi] * w->visrowwords;
	int st_head[64], st_node[64], sp = 0;
	int pfrontenter = 0, pfrontexit = 0;		/* file statics in the reference */
	st_head[sp] = 0; st_node[sp] = be->topnode; sp++;
	while (sp)
	{
		sp--;
		int pedges = st_head[sp];
		if (st_node[sp] < 0)
		{
			/* reached a non-solid leaf: r_currentbkey = leaf->key; R_RenderBmodelFace */
			int n = 0;
			for (int e = pedges; e >= 0 && n < RC_MAX_FACE_EDGES + 1; e = e_next[e], n++)
			{
				ev0[n] = e_v0[e]; ev1[n] = e_v1[e];
			}
			rc_bmodel_emit (&c, ev0, ev1, n, be->clipflags, leafkey[-1 - st_node[sp]]);
			continue;
		}
		const rc_node_t *pnode = &w->nodes[st_node[sp]];
		int psideedges[2] = { -1, -1 };
		int makeclippededge = 0;
		/* transform the BSP plane into model space */
		const rc_plane_t *splitplane = &w->planes[pnode->plane];
		float tdist = splitplane->dist - rc_dot3 (be->origin, splitplane->normal);
		float tnormal[3];
		tnormal[0] = rc_dot3 (be->rot[0], splitplane->normal);
		tnormal[1] = rc_dot3 (be->rot[1], splitplane->normal);
		tnormal[2] = rc_dot3 (be->rot[2], splitplane->normal);
		int overflow = 0;
		for (int pnextedge; pedges >= 0; pedges = pnextedge)
		{
			pnextedge = e_next[pedges];
			const int plastvert = e_v0[pedges], pvert = e_v1[pedges];
			const float *lp = rc_bvert (&c, plastvert), *pp = rc_bvert (&c, pvert);
			const float lastdist = rc_dot3 (lp, tnormal) - tdist;
			const int lastside = (lastdist <= 0);
			const float dist = rc_dot3 (pp, tnormal) - tdist;
			const int side = (dist <= 0);
			if (side != lastside)
			{
				if (c.numbverts >= RC_BMODEL_VERTS || numbedges >= RC_BMODEL_EDGES - 1)
				{
					overflow = 1;
					break;
				}
				const float frac = lastdist / (lastdist - dist);
				const int ptvert = c.numbverts++;
				c.bverts[ptvert][0] = lp[0] + frac * (pp[0] - lp[0]);
				c.bverts[ptvert][1] = lp[1] + frac * (pp[1] - lp[1]);
				c.bverts[ptvert][2] = lp[2] + frac * (pp[2] - lp[2]);
				int pt = numbedges;
				e_next[pt] = psideedges[lastside]; psideedges[lastside] = pt; e_v0[pt] = plastvert; e_v1[pt] = ptvert;
				pt = numbedges + 1;
				e_next[pt] = psideedges[side]; psideedges[side] = pt; e_v0[pt] = ptvert; e_v1[pt] = pvert;
				numbedges += 2;
				if (side == 0) pfrontenter = ptvert; else pfrontexit = ptvert;
				makeclippededge = 1;
			}
			else
			{
				e_next[pedges] = psideedges[side];
				psideedges[side] = pedges;
			}
		}
		if (overflow)
		{
			RC_ATOMIC_OR_I (f->status, RC_STAT_BMODEL_OVERFLOW);
			break;
		}
		if (makeclippededge)
		{
			if (numbedges >= RC_BMODEL_EDGES - 2)
			{
				RC_ATOMIC_OR_I (f->status, RC_STAT_BMODEL_OVERFLOW);
				break;
			}
			int pt = numbedges;
			e_next[pt] = psideedges[0]; psideedges[0] = pt; e_v0[pt] = pfrontexit; e_v1[pt] = pfrontenter;
			pt = numbedges + 1;
			e_next[pt] = psideedges[1]; psideedges[1] = pt; e_v0[pt] = pfrontenter; e_v1[pt] = pfrontexit;
			numbedges += 2;
		}
		/* the reference finishes child 0 (leaf: draw it, node: recurse) before child 1: push 1 then 0 */
		for (int i = 1; i >= 0; i--)
		{
			if (psideedges[i] < 0)
				continue;
			const int pn = pnode->children[i];
			const int idx = pn >= 0 ? pn : w->numnodes + (-1 - pn);
			if (!rc_visbit (vis, idx))
				continue;	/* this branch isn't in the PVS */
			if (pn < 0 && w->leafs[-1 - pn].contents == RC_CONTENTS_SOLID)
				continue;
			if (sp >= 64)
			{
				RC_ATOMIC_OR_I (f->status, RC_STAT_BMODEL_OVERFLOW);
				continue;
			}
			st_head[sp] = psideedges[i]; st_node[sp] = pn; sp++;
		}
	}
}

/* ------------------------------------------------------------------ stage 3: scan ------ */
typedef struct { unsigned long long hi, lo; uint32_t surfs; int iu; } rc_aet_t;

typedef struct {
	int surf, key, state, last_u, insub;
} rc_stk_t;

/* Sort key of one active edge on scanline `line` (DESIGN.md "AET order"): the order the
 * reference's linked list has when R_GenerateSpans walks it, in closed form.
 *   hi = u (12.20, biased) : then for edges already active on line-1, ~u_step (the edge that was
 *        further left on the previous line, i.e. the larger step, stays in front:
 *        R_StepActiveU is a stable insertion sort, r_edge.c:188-249); new edges (c = 0) precede
 *        old ones with the same u (R_InsertNewEdges inserts before, r_edge.c:127-163)
 *   lo = later start line first, then newedges[] order: leaders newest-emitted first, trailers
 *        after leaders and oldest first (u_check = u+1 for trailers, r_rast.c:355-379). */
RC_HD void rc_aet_make (const rc_edge_t *ed, int line, rc_aet_t *a)
{
	int isold = line > ed->v;
	int u = rc_addw (ed->u, rc_mulw (line - ed->v, ed->u_step));
	uint32_t c = 0;
	if (isold)
	{
		c = ~((uint32_t)ed->u_step ^ 0x80000000u);
		if (c == 0) c = 1;
	}
	uint32_t em = ed->rank & 0x7FFFFFFFu;
	uint32_t trailing = ed->rank >> 31;
	uint32_t fl = trailing ? em : (~em & 0x7FFFFFFFu);
	a->hi = ((unsigned long long)((uint32_t)u ^ 0x80000000u) << 32) | c;
	a->lo = ((unsigned long long)(uint16_t)(~ed->v) << 32) | ((unsigned long long)trailing << 31) | fl;
	a->surfs = (uint32_t)ed->surfs[0] | ((uint32_t)ed->surfs[1] << 16);
	a->iu = u >> 20;
}

/* accessors so the span walk can run over an array of rc_aet_t (serial scan, host sim) or over
 * the split shared-memory arrays of the warp-cooperative kernel */
struct rc_aet_array {
	const rc_aet_t *a;
	RC_HD2 int iu (int k) const { return a[k].iu; }
	RC_HD2 uint32_t surfs (int k) const { return a[k].surfs; }
	RC_HD2 int u (int k) const { return (int)((uint32_t)(a[k].hi >> 32) ^ 0x80000000u); }
};

/* spans of one scanline before they are published (left to right: the walk emits them in
 * increasing u): surface slot, first column, end column */
typedef struct {
	uint16_t *surf;
	int16_t  *u0, *u1;
	int       cap;
} rc_linespans_t;

#define RC_EMIT_SPAN(sidx, ustart, uend)                                                      \
	do {                                                                                  \
		if (nsp < out->cap) {                                                         \
			out->surf[nsp] = (uint16_t)(sidx);                                    \
			out->u0[nsp] = (int16_t)(ustart); out->u1[nsp] = (int16_t)(uend);     \
		}                                                                             \
		nsp++;                                                                        \
	} while (0)

/* R_GenerateSpans (r_edge.c:536-563) with R_TrailingEdge (:370-407), R_LeadingEdge (:411-527)
 * and R_CleanupSpan (:256-285) over the sorted active edges of one scanline.  The active
 * surface stack is a small array ordered top..background instead of a linked list through
 * surf_t; spanstate lives in the stack entry (a surface is in the stack iff spanstate >= 1).
 * Returns the number of spans emitted into `out` (more than out->cap: overflow). */
template <class AET>
RC_HD int rc_span_walk (const rc_frame_t *f, const rc_view_t *vw, const rc_surf_t *surfs, int line,
			const AET aet, int n, const rc_linespans_t *out)
{
	int nsp = 0;
	rc_stk_t stk[RC_STACK_MAX];
	int depth = 1;
	int nneg = 0;
	rc_stk_t neg[8];
	const int edge_head_u = vw->vrect_x;			/* edge_head.u >> 20 */
	const int edge_tail_u = vw->vrectright;		/* ((vrectright<<20)+0xFFFFF) >> 20 */
	stk[0].surf = 1; stk[0].key = 0x7FFFFFFF; stk[0].state = 1; stk[0].last_u = edge_head_u; stk[0].insub = 0;
	const float fv = (float)line;
	/* R_LeadingEdgeBackwards has no 1/z test between two brush-model surfaces of one leaf: the new one goes in front
	 * ("they'll never be farthest anyway", r_edge.c:312-343) */
	const int backward = (vw->rdflags & RC_VIEW_DRAWORDER) != 0;

	for (int k = 0; k < n; k++)
	{
		const int iu = aet.iu (k);
		const uint32_t ss = aet.surfs (k);
		const int s0 = ss & 0xFFFF, s1 = ss >> 16;
		if (s0)
		{
			/* R_TrailingEdge */
			int pos = -1;
			for (int i = 0; i < depth; i++)
				if (stk[i].surf == s0) { pos = i; break; }
			if (pos >= 0)
			{
				if (--stk[pos].state == 0)
				{
					if (pos == 0)
					{
						if (iu > stk[0].last_u)
							RC_EMIT_SPAN (s0, stk[0].last_u, iu);
						stk[1].last_u = iu;
					}
					for (int i = pos; i < depth - 1; i++)
						stk[i] = stk[i+1];
					depth--;
				}
			}
			else
			{
				/* spanstate goes negative: trailing edge seen before the leading one */
				int ni = -1;
				for (int i = 0; i < nneg; i++)
					if (neg[i].surf == s0) { ni = i; break; }
				if (ni < 0)
				{
					if (nneg < 8) { ni = nneg++; neg[ni].surf = s0; neg[ni].state = 0; }
					else RC_ATOMIC_OR_I (f->status, RC_STAT_STACK_OVERFLOW);
				}
				if (ni >= 0)
					neg[ni].state--;
			}
			if (!s1)
				continue;
		}
		if (s1)
		{
			/* R_LeadingEdge */
			int pos = -1, ni = -1, state;
			for (int i = 0; i < depth; i++)
				if (stk[i].surf == s1) { pos = i; break; }
			if (pos >= 0)
			{
				stk[pos].state++;	/* already in a span: ++spanstate != 1 */
				continue;
			}
			for (int i = 0; i < nneg; i++)
				if (neg[i].surf == s1) { ni = i; break; }
			state = (ni >= 0) ? neg[ni].state : 0;
			state++;
			if (ni >= 0)
			{
				neg[ni].state = state;
				if (state == 0) { neg[ni] = neg[nneg-1]; nneg--; }
			}
			if (state != 1)
				continue;

			const rc_surf_t *sf = &surfs[s1];
			const int skey = sf->key, sinsub = sf->insubmodel;
			int at;			/* insert before stk[at] */
			int newtop = 0;
			if (skey < stk[0].key)
				newtop = 1;
			else if (sinsub && skey == stk[0].key)
			{
				/* two bmodels in the same leaf: sort on 1/z (r_edge.c:436-458) */
				const rc_surf_t *s2 = &surfs[stk[0].surf];
				double fu = (float)(aet.u (k) - 0xFFFFF) * (1.0 / 0x100000);
				double newzi = sf->d_ziorigin + fv*sf->d_zistepv + fu*sf->d_zistepu;
				double newzibottom = newzi * 0.99;
				double testzi = s2->d_ziorigin + fv*s2->d_zistepv + fu*s2->d_zistepu;
				if (backward || newzibottom >= testzi)
					newtop = 1;
				else
				{
					double newzitop = newzi * 1.01;
					if (newzitop >= testzi && sf->d_zistepu >= s2->d_zistepu)
						newtop = 1;
				}
			}
			if (newtop)
			{
				if (iu > stk[0].last_u)
					RC_EMIT_SPAN (stk[0].surf, stk[0].last_u, iu);
				at = 0;
			}
			else
			{
				at = 0;
				for (;;)
				{
					do { at++; } while (skey > stk[at].key);
					if (skey == stk[at].key)
					{
						if (!sinsub)
							continue;
						const rc_surf_t *s2 = &surfs[stk[at].surf];
						double fu = (float)(aet.u (k) - 0xFFFFF) * (1.0 / 0x100000);
						double newzi = sf->d_ziorigin + fv*sf->d_zistepv + fu*sf->d_zistepu;
						double newzibottom = newzi * 0.99;
						double testzi = s2->d_ziorigin + fv*s2->d_zistepv + fu*s2->d_zistepu;
						if (backward || newzibottom >= testzi)
							break;
						double newzitop = newzi * 1.01;
						if (newzitop >= testzi && sf->d_zistepu >= s2->d_zistepu)
							break;
						continue;
					}
					break;
				}
			}
			if (depth >= RC_STACK_MAX)
			{
				RC_ATOMIC_OR_I (f->status, RC_STAT_STACK_OVERFLOW);
				continue;
			}
			for (int i = depth; i > at; i--)
				stk[i] = stk[i-1];
			depth++;
			stk[at].surf = s1; stk[at].key = skey; stk[at].state = 1; stk[at].insub = sinsub;
			stk[at].last_u = newtop ? iu : 0;
		}
	}
	/* R_CleanupSpan */
	if (edge_tail_u > stk[0].last_u)
		RC_EMIT_SPAN (stk[0].surf, stk[0].last_u, edge_tail_u);
	if (nneg)
		RC_ATOMIC_OR_I (f->status, RC_STAT_STACK_OVERFLOW);	/* spanstate would leak into the next line */
	return nsp;
}

/* ---- publishing a scanline's spans -----------------------------------------------------------
 * Work list of the draw stage (DESIGN.md "draw"): world spans tile the view rectangle.  The draw
 * kernel runs one thread per SCREEN-ALIGNED 8-pixel cell that a single span covers completely (one
 * aligned 8-byte colour store and one aligned 16-byte z store, no masks); the cells that hold a span
 * boundary are drawn whole by one owner from the shares of their spans (rc_mixed_owner).
 * cellspan[(view, line, cell)] = {span, subdivision record of the cell's first pixel * 8 + its
 * position inside that subdivision}; si = -1 where no span covers the whole cell. */
/* segments of a span: units of 8 boundary records (and one thread of the segment kernel); nsub + 2 records are needed
 * (B[0] .. B[nsub], the last subdivision's step) */
RC_HD int rc_span_seg_count (int count) { return (((count + 7) >> 3) + 2 + 7) >> 3; }
/* the span field of a cell record: the span, and where the cell (first pixel at span pixel `rel`) meets the last subdivision */
RC_HD int rc_cell_si (int si, int rel, int count)
{
	const int k = rel >> 3, last = ((count + 7) >> 3) - 1;
	return (int)((unsigned)si | (k == last ? (unsigned)RC_CELL_L0 : 0u) | (k + 1 == last ? RC_CELL_L1 : 0u));
}

RC_HD void rc_write_span (const rc_frame_t *f, const rc_view_t *vw, int vi, int line, int si, int surf, int u0, int count, int segbase)
{
	rc_span_t *s = &f->spans[si];
	s->u = (int16_t)u0; s->count = (int16_t)count;
	s->pixofs = (vi * f->height + line) * f->width + u0;
	s->segbase = segbase; s->iziorg = 0; s->izistep = 0; s->cacheofs = 0;
	s->cwk = ((uint32_t)line << 20) | (vw->warp_index >= 0 ? RC_SPAN_WARPBIT : 0u);
	s->gsurf = vi * f->surfcap + surf;
}

RC_HD void rc_line_clear (const rc_frame_t *f, int vi, int line)
{
	const int cw = (f->width + 7) >> 3;
	rc_cell_t *cells = f->cellspan + ((size_t)vi * f->height + line) * cw;
	for (int c = 0; c < cw; c++)
	{
		cells[c].si = -1; cells[c].sub = 0;
	}
}

/* serial version (host simulation; the GPU scan kernel does the same with a warp) */
RC_HD void rc_line_publish (const rc_frame_t *f, const rc_view_t *vw, rc_surf_t *surfs, int vi, int line, const rc_linespans_t *ls, int nsp)
{
	int nseg = 0;
	for (int i = 0; i < nsp; i++)
		nseg += rc_span_seg_count (ls->u1[i] - ls->u0[i]);
	unsigned long long old = RC_ATOMIC_ADD_ULL (&f->alloc[0], (unsigned long long)nsp | ((unsigned long long)nseg << 32));
	int spanbase = (int)(old & 0xFFFFFFFFu), segbase = (int)(old >> 32);
	if (nsp > ls->cap || spanbase + nsp > f->spancap || segbase + nseg > f->segcap)
	{
		RC_ATOMIC_OR_I (f->status, RC_STAT_SPAN_POOL);
		rc_line_clear (f, vi, line);
		return;
	}
	const int cw = (f->width + 7) >> 3;
	rc_cell_t *cells = f->cellspan + ((size_t)vi * f->height + line) * cw;
	rc_line_clear (f, vi, line);
	for (int i = 0; i < nsp; i++)
	{
		const int u0 = ls->u0[i], cnt = ls->u1[i] - u0;
		rc_write_span (f, vw, vi, line, spanbase + i, ls->surf[i], u0, cnt, segbase);
		for (int c = (u0 + 7) >> 3; (c << 3) + 8 <= u0 + cnt; c++)
		{
			cells[c].si = rc_cell_si (spanbase + i, (c << 3) - u0, cnt);
			cells[c].sub = (uint32_t)(segbase * 8) * 8u + (uint32_t)((c << 3) - u0);
		}
		for (int q = rc_span_seg_count (cnt); q > 0; q--)
			f->segspan[segbase++] = spanbase + i;
		surfs[ls->surf[i]].hasspans = 1;
	}
}

/* One thread per (view, scanline), fully serial: the reference implementation of the scan
 * stage (host simulation; the GPU runs the warp-per-line k_scan built from the same parts). */
RC_HD void rc_scan_line (const rc_world_t *w, const rc_frame_t *f, int vi, int line)
{
	const rc_view_t *vw = &f->views[vi];
	if (vw->rdflags & 1 /* RDF_NOWORLDMODEL */)
	{
		/* R_EdgeDrawing returns at once (r_main.c:889): no span, not even the background, is drawn; D_ViewChanged
		 * has set the WHOLE z-buffer to 0xFFFF for the entities to test against (d_modech.c:103-106) */
		rc_line_clear (f, vi, line);
		if (f->zb)
			for (int x = 0; x < f->width; x++)
				f->zb[((size_t)vi * f->height + line) * f->width + x] = (int16_t)-1;
		return;
	}
	if (line < vw->vrect_y || line >= vw->vrectbottom)
	{
		rc_line_clear (f, vi, line);
		return;
	}
	const rc_edge_t *edges = f->edges + (size_t)vi * f->edgecap;
	rc_surf_t *surfs = f->surfs + (size_t)vi * f->surfcap;
	int E = f->nedges[vi];
	if (E > f->edgecap) E = f->edgecap;
	rc_aet_t aet[RC_AET_MAX];
	int n = 0;
	for (int e = 0; e < E; e++)
	{
		if (edges[e].v > line || edges[e].v2 < line)
			continue;
		if (n >= RC_AET_MAX)
		{
			RC_ATOMIC_OR_I (f->status, RC_STAT_AET_OVERFLOW);
			break;
		}
		rc_aet_t a;
		rc_aet_make (&edges[e], line, &a);
		int j = n++;
		while (j > 0 && (aet[j-1].hi > a.hi || (aet[j-1].hi == a.hi && aet[j-1].lo > a.lo)))
		{
			aet[j] = aet[j-1];
			j--;
		}
		aet[j] = a;
	}
	uint16_t lsurf[RC_AET_MAX + 2];
	int16_t lu0[RC_AET_MAX + 2], lu1[RC_AET_MAX + 2];
	rc_linespans_t ls; ls.surf = lsurf; ls.u0 = lu0; ls.u1 = lu1; ls.cap = RC_AET_MAX + 2;
	rc_aet_array acc; acc.a = aet;
	int nsp = rc_span_walk (f, vw, surfs, line, acc, n, &ls);
	rc_line_publish (f, vw, surfs, vi, line, &ls, nsp);
}

/* surf_t slots of a view in use: 1 (background), 2 .. nfaces+1 (world faces), then the brush-entity
 * fragments at facecap+2 ..; maps a dense index to the slot, -1 past the end */
RC_HD int rc_surf_slot (const rc_frame_t *f, int vi, int idx)
{
	const int nf = f->nfaces[vi] < f->facecap ? f->nfaces[vi] : f->facecap;
	if (idx < nf + 1)
		return idx + 1;
	idx -= nf + 1;
	if (f->nbsurfs)
	{
		const int nb = f->nbsurfs[vi] < f->bsurfcap ? f->nbsurfs[vi] : f->bsurfcap;
		if (idx < nb)
			return f->facecap + 2 + idx;
	}
	return -1;
}

/* ------------------------------------------------------------- stage 4: surface prep --- */
RC_HD int rc_lightadj_code (int value)
{
	/* compact code for an 8.8 light scale ((c-'a')*22, or 256 for an unset style,
	 * r_light.c:40-48); 254 = not representable -> the block is built privately */
	if (value == 256) return 255;
	if (value % 22 == 0) { int c = value / 22 + 128; if (c >= 0 && c <= 253) return c; }
	return 254;
}

RC_HD uint32_t rc_hash64 (unsigned long long k)
{
	k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
	return (uint32_t)k;
}

/* D_CalcGradients, d_edge.c:96-157 */
RC_HD void rc_calc_gradients (const rc_view_t *vw, const rc_xform_t *xf, const rc_face_t *fa, const rc_texinfo_t *ti, int miplevel, rc_surf_t *s)
{
	float mipscale = 1.0 / (float)(1 << miplevel);
	float p_saxis[3], p_taxis[3], p_temp1[3];
	float t;
	p_saxis[0] = rc_dot3 (ti->vecs[0], xf->vright); p_saxis[1] = rc_dot3 (ti->vecs[0], xf->vup); p_saxis[2] = rc_dot3 (ti->vecs[0], xf->vpn);
	p_taxis[0] = rc_dot3 (ti->vecs[1], xf->vright); p_taxis[1] = rc_dot3 (ti->vecs[1], xf->vup); p_taxis[2] = rc_dot3 (ti->vecs[1], xf->vpn);
	t = vw->xscaleinv * mipscale;
	s->d_sdivzstepu = p_saxis[0] * t;
	s->d_tdivzstepu = p_taxis[0] * t;
	t = vw->yscaleinv * mipscale;
	s->d_sdivzstepv = -p_saxis[1] * t;
	s->d_tdivzstepv = -p_taxis[1] * t;
	s->d_sdivzorigin = p_saxis[2] * mipscale - vw->xcenter * s->d_sdivzstepu - vw->ycenter * s->d_sdivzstepv;
	s->d_tdivzorigin = p_taxis[2] * mipscale - vw->xcenter * s->d_tdivzstepu - vw->ycenter * s->d_tdivzstepv;
	p_temp1[0] = xf->tmodelorg[0] * mipscale;
	p_temp1[1] = xf->tmodelorg[1] * mipscale;
	p_temp1[2] = xf->tmodelorg[2] * mipscale;
	t = 0x10000*mipscale;
	s->sadjust = rc_f2i ((float)(rc_d2i (rc_dot3 (p_temp1, p_saxis) * 0x10000 + 0.5) -
			((fa->texturemins[0] << 16) >> miplevel)) + ti->vecs[0][3]*t);
	s->tadjust = rc_f2i ((float)(rc_d2i (rc_dot3 (p_temp1, p_taxis) * 0x10000 + 0.5) -
			((fa->texturemins[1] << 16) >> miplevel)) + ti->vecs[1][3]*t);
	s->bbextents = ((fa->extents[0] << 16) >> miplevel) - 1;
	s->bbextentt = ((fa->extents[1] << 16) >> miplevel) - 1;
}

#define RC_SRC_POOL    0
#define RC_SRC_TEXELS  1
#define RC_SRC_SOLID   2    /* background / fast sky: constant colour, z at "infinity" */
#define RC_SRC_SKY     3    /* scrolling two-layer sky (d_sky.c) */

/* R_PushDlights / R_MarkLights (r_light.c:66-151) for one face: the bit of every dynamic light
 * whose BSP descent reaches the face's node with |dist| <= radius.  The reference descends
 * from the root choosing children by plane distance; a node is marked iff no ancestor's
 * test sends the descent away from it and its own plane is within the radius. */
RC_HD int rc_dlight_bits (const rc_world_t *w, const rc_frame_t *f, int headnode, int dlbase, int num, int node)
{
	int bits = 0;
	const int tx = w->nodeaux[node].dfs_in;
	for (int l = 0; l < num; l++)
	{
		const rc_dlight_t *dl = &f->dlights[dlbase + l];
		int cur = headnode;
		for (;;)
		{
			const rc_node_t *nd = &w->nodes[cur];
			float dist = rc_plane_diff (dl->origin, &w->planes[nd->plane]);
			if (cur == node)
			{
				if (!(dist > dl->radius) && !(dist < -dl->radius))
					bits |= 1 << l;
				break;
			}
			const int c = (tx < w->nodeaux[cur].dfs_mid) ? 0 : 1;
			if (dist > dl->radius) { if (c != 0) break; }
			else if (dist < -dl->radius) { if (c != 1) break; }
			cur = nd->children[c];
			if (cur < 0)
				break;
		}
	}
	return bits;
}

#define RC_CACHE_UNIT_TEXELS 1024
RC_HD int rc_cache_band_rows (int bw) { const int r = RC_CACHE_UNIT_TEXELS / (bw > 0 ? bw : 1); return r < 4 ? 4 : r; }

/* One thread per (view, surface): the per-surface part of D_DrawSurfaces (d_edge.c:161-340)
 * for surfaces that own spans: mip selection, D_CacheSurface lookup (d_surf.c:196) against
 * the shared GPU surface cache, gradients.  surf->pad = texel source. */
RC_HD void rc_surf_prep (const rc_world_t *w, const rc_frame_t *f, int vi, int si)
{
	const rc_view_t *vw = &f->views[vi];
	rc_surf_t *s = f->surfs + (size_t)vi * f->surfcap + si;
	if (!s->hasspans || s->flags == -1)
		return;
	s->cacheofs = -1;
	s->miplevel = -1;		/* reused as the cache-table slot to resolve, -1: none */
	s->drawflags = s->flags;
	s->d_zistepu_draw = s->d_zistepu;
	s->izistep = rc_f2i (s->d_zistepu * 0x8000 * 0x10000);
	if (s->flags & RC_SURF_DRAWSKYBOX)
	{
		/* d_edge.c:211-235: one of the six 256x256 pictures (at the start of the cache pool), mip 0,
		 * texture axes of the box face with offsets that follow the viewer (r_rast.c:190-194); the
		 * z-buffer gets the background value (rc_span_seg) */
		const rc_face_t *bf = &w->faces[s->face];
		rc_texinfo_t ti = w->texinfo[bf->texinfo];
		ti.vecs[0][3] = -rc_dot3 (vw->origin, ti.vecs[0]);
		ti.vecs[1][3] = -rc_dot3 (vw->origin, ti.vecs[1]);
		s->pad = RC_SRC_POOL;
		s->cacheofs = (s->face - w->skyfirst) * (256 * 256);
		s->cachewidth = 256;
		s->izistep = 0;
		rc_calc_gradients (vw, rc_view_xform (vw), bf, &ti, 0, s);
		return;
	}
	if ((s->flags & RC_SURF_DRAWSKY) && !vw->fastsky)
	{
		s->pad = RC_SRC_SKY;		/* d_edge.c:186-193; z uses the surface's own 1/z plane */
		return;
	}
	if (s->flags & (RC_SURF_DRAWBACKGROUND | RC_SURF_DRAWSKY))
	{
		s->pad = RC_SRC_SOLID;
		s->cachewidth = (s->flags & RC_SURF_DRAWBACKGROUND) ? vw->clearcolor : vw->skycolor;
		s->d_ziorigin = -0.9; s->d_zistepu = 0; s->d_zistepv = 0;
		s->d_zistepu_draw = 0;
		s->izistep = 0;
		return;
	}
	const rc_face_t *fa = &w->faces[s->face];
	const rc_texinfo_t *ti = &w->texinfo[fa->texinfo];
	/* brush-entity surfaces are textured in the entity's rotated frame (d_edge.c:247-264,290-306) */
	const rc_bent_t *be = s->insubmodel ? &f->bents[s->entity] : (const rc_bent_t *)0;
	const rc_xform_t *xf = be ? &be->x : rc_view_xform (vw);
	if (s->flags & RC_SURF_DRAWTURB)
	{
		s->pad = RC_SRC_TEXELS;
		s->cacheofs = w->textures[ti->texture].offsets[0];
		s->cachewidth = 64;
		rc_calc_gradients (vw, xf, fa, ti, 0, s);
		return;
	}
	/* D_MipLevelForScale, d_edge.c:44-63 */
	float scale = s->nearzi * vw->scale_for_mip * ti->mipadjust;
	int mip;
	if (scale >= vw->d_scalemip[0]) mip = 0;
	else if (scale >= vw->d_scalemip[1]) mip = 1;
	else if (scale >= vw->d_scalemip[2]) mip = 2;
	else mip = 3;
	if (mip < vw->d_minmip) mip = vw->d_minmip;

	/* R_TextureAnimation, r_surf.c:60-87 (world entity: frame 0) */
	int tex = ti->texture;
	if (be && be->frame && w->textures[tex].alternate_anims >= 0)
		tex = w->textures[tex].alternate_anims;
	const int basetex = tex;
	if (w->textures[tex].anim_total)
	{
		int reletive = vw->animtime % w->textures[tex].anim_total;
		int count = 0;
		while (w->textures[tex].anim_min > reletive || w->textures[tex].anim_max <= reletive)
		{
			tex = w->textures[tex].anim_next;
			if (tex < 0 || ++count > 100) { tex = basetex; break; }
		}
	}
	int la[4], codes[4], cacheable = 1;
	for (int k = 0; k < 4; k++)
	{
		int st = fa->styles[k];
		la[k] = st < RC_MAX_LIGHTSTYLES ? vw->lightstylevalue[st] : 0;	/* d_lightstylevalue[256]: 64.. never set */
		codes[k] = rc_lightadj_code (la[k]);
		if (codes[k] == 254) cacheable = 0;
	}
	int dlbits = 0;
	const int dlbase = be ? be->dlight_ofs : vw->dlight_ofs;
	if (vw->num_dlights && fa->node >= 0 && dlbase >= 0)
		dlbits = rc_dlight_bits (w, f, be ? be->headnode : 0, dlbase, vw->num_dlights, fa->node);
	if (dlbits)
		cacheable = 0;		/* cache->dlight: rebuilt every frame (d_surf.c:214-228), private to the view */
	int bw = fa->extents[0] >> mip, bh = fa->extents[1] >> mip;
	int bytes = (bw * bh + 15) & ~15;
	int slot = -1, creator = 0, ofs = -1;
	if (cacheable)
	{
		unsigned long long key = ((unsigned long long)s->face << 44) | ((unsigned long long)mip << 42) |
			((unsigned long long)tex << 32) | ((unsigned long long)codes[0] << 24) |
			((unsigned long long)codes[1] << 16) | ((unsigned long long)codes[2] << 8) | (unsigned long long)codes[3];
		uint32_t h = rc_hash64 (key) & (uint32_t)f->cachemask;
		for (int probe = 0; probe <= f->cachemask; probe++)
		{
			unsigned long long prev = RC_ATOMIC_CAS_ULL (&f->cachekeys[h], RC_CACHE_EMPTY, key);
			if (prev == RC_CACHE_EMPTY) { slot = (int)h; creator = 1; break; }
			if (prev == key) { slot = (int)h; break; }
			h = (h + 1) & (uint32_t)f->cachemask;
		}
		if (slot < 0) { RC_ATOMIC_OR_I (f->status, RC_STAT_CACHE_POOL); cacheable = 0; }
	}
	if (!cacheable || creator)
	{
		unsigned long long o = RC_ATOMIC_ADD_ULL (&f->alloc[1], (unsigned long long)bytes);
		int j = (int)RC_ATOMIC_ADD_ULL (&f->alloc[2], 1ull);
		if ((long long)(o + bytes) > f->cachecap || j >= f->jobcap)
		{
			RC_ATOMIC_OR_I (f->status, RC_STAT_CACHE_POOL);
			ofs = 0;
		}
		else
		{
			rc_cachejob_t *job = &f->jobs[j];
			ofs = (int)o;
			job->face = s->face; job->mip = mip; job->texture = tex;
			job->lightadj[0] = la[0]; job->lightadj[1] = la[1]; job->lightadj[2] = la[2]; job->lightadj[3] = la[3];
			job->ambient = vw->ambientlight; job->dstofs = ofs; job->width = bw; job->height = bh;
			job->view = dlbits ? vi : -1; job->pad = vw->fullbright; job->dlightbits = dlbits;
			job->pad2[0] = dlbase; job->pad2[1] = vw->num_dlights;
			{
				const rc_texture_t *mt = &w->textures[tex];
				job->lightofs = fa->lightofs;
				for (int k = 0; k < 4; k++) job->styles[k] = fa->styles[k];
				job->smax = (fa->extents[0] >> 4) + 1; job->tmax = (fa->extents[1] >> 4) + 1;
				job->tw = mt->width >> mip; job->th = mt->height >> mip;
				job->soffset = ((fa->texturemins[0] >> mip) + (job->tw << 16)) % job->tw;
				job->toffset = ((fa->texturemins[1] >> mip) + (job->th << 16)) % job->th;
				job->srcofs = (int)mt->offsets[mip];
				job->bandrows = rc_cache_band_rows (bw);
				job->items = (bw + 15) >> 4;
				job->qG = 32 / job->items; job->rG = 32 - job->qG * job->items;
				job->dsx = (job->rG << 4) % job->tw; job->wsx = (job->items << 4) % job->tw; job->dsy = job->qG % job->th;
			}
			if (f->cacheunits)
			{
				/* the job as bands of rows of about RC_CACHE_UNIT_TEXELS texels: the build kernel's work units */
				const int rb = rc_cache_band_rows (bw), nb = (bh + rb - 1) / rb;
				const unsigned long long ub = RC_ATOMIC_ADD_ULL (&f->alloc[8], (unsigned long long)nb);
				for (int b = 0; b < nb; b++)
					if (ub + b < (unsigned long long)f->cacheunitcap)
					{
						f->cacheunits[2 * (ub + b)] = j; f->cacheunits[2 * (ub + b) + 1] = b * rb;
					}
			}
		}
		if (creator)
			f->cachevals[slot] = ofs;
	}
	s->pad = RC_SRC_POOL;
	s->cacheofs = ofs;
	s->miplevel = (cacheable && !creator) ? slot : -1;
	s->cachewidth = bw;
	rc_calc_gradients (vw, xf, fa, ti, mip, s);
}

RC_HD void rc_surf_resolve (const rc_frame_t *f, int vi, int si)
{
	rc_surf_t *s = f->surfs + (size_t)vi * f->surfcap + si;
	if (!s->hasspans || s->flags == -1)
		return;
	if (s->pad == RC_SRC_POOL && s->miplevel >= 0)
	{
		s->cacheofs = f->cachevals[s->miplevel];
		if (s->cacheofs == 0)
			RC_ATOMIC_OR_I (f->status, RC_STAT_CACHE_POOL);
	}
}

/* ------------------------------------------------------------- stage 5: cache build ---- */
/* one lightmap sample of R_BuildLightmap (r_light.c:372-425) */
RC_HD unsigned rc_cache_luxel (const rc_world_t *w, const rc_frame_t *f, const rc_cachejob_t *job, int i)
{
	const int smax = job->smax, tmax = job->tmax;
	int size = smax * tmax;
	if (job->pad /* r_fullbright */ || !w->haslight)
		return 0;
	unsigned bl = (unsigned)(job->ambient << 8);
	if (job->lightofs != -1)
	{
		const uint8_t *lightmap = w->lightdata + job->lightofs;
		for (int maps = 0; maps < RC_MAXLIGHTMAPS && job->styles[maps] != 255; maps++)
		{
			unsigned scale = (unsigned)job->lightadj[maps];
			bl += lightmap[i] * scale;
			lightmap += size;
		}
	}
	if (job->dlightbits)
	{
		/* R_AddDynamicLights, r_light.c:274-365 (world entity: no rotation, origin 0) */
		/* job->pad2[0]: first dlight (already relative to the entity and rotated into its frame for brush
		 * entities: r_light.c:300-311), job->pad2[1]: how many */
		const rc_face_t *fa = &w->faces[job->face];
		const rc_texinfo_t *tex = &w->texinfo[fa->texinfo];
		const rc_plane_t *plane = &w->planes[fa->plane];
		const int ls = i % smax, lt = i / smax;
		for (int lnum = 0; lnum < job->pad2[1]; lnum++)
		{
			if (!(job->dlightbits & (1 << lnum)))
				continue;
			const rc_dlight_t *l = &f->dlights[job->pad2[0] + lnum];
			float dlorigin[3], impact[3], local[2];
			dlorigin[0] = l->origin[0] - 0.0f; dlorigin[1] = l->origin[1] - 0.0f; dlorigin[2] = l->origin[2] - 0.0f;
			float rad = l->radius;
			float dist = rc_plane_diff (dlorigin, plane);
			rad -= fabs ((double)dist);
			if (rad < 0)
				continue;
			if (plane->type < 3)
			{
				impact[0] = dlorigin[0]; impact[1] = dlorigin[1]; impact[2] = dlorigin[2];
				impact[plane->type] -= dist;
			}
			else
			{
				for (int q = 0; q < 3; q++)
					impact[q] = dlorigin[q] - plane->normal[q]*dist;
			}
			local[0] = rc_dot3 (impact, tex->vecs[0]) + tex->vecs[0][3];
			local[1] = rc_dot3 (impact, tex->vecs[1]) + tex->vecs[1][3];
			local[0] -= fa->texturemins[0];
			local[1] -= fa->texturemins[1];
			int td = rc_f2i (local[1] - lt*16);
			if (td < 0) td = -td;
			int sd = rc_f2i (local[0] - ls*16);
			if (sd < 0) sd = -sd;
			if (sd > td) dist = sd + (td>>1);
			else dist = td + (sd>>1);
			if (dist < rad)
				bl = (unsigned)(long long)((float)bl + (rad - dist)*256);
		}
	}
	int t = (255*256 - (int)bl) >> (8 - 6);
	if (t < (1 << 6))
		t = (1 << 6);
	return (unsigned)t;
}

/* one texel of R_DrawSurface + R_DrawSurfaceBlock8_mipN (r_surf.c:95-375), in closed form:
 * the block drawers step light with wrapping 32-bit adds, so light at (row i, column b) of
 * a block is lightright + i*rightstep + (bs-1-b)*((lightleft + i*leftstep - that) >> shift). */
RC_HD uint8_t rc_cache_texel (const rc_world_t *w, const rc_cachejob_t *job, const unsigned *bl, int x, int y)
{
	const int mip = job->mip;
	const int shift = 4 - mip, bs = 16 >> mip;
	const int lw = job->smax;
	const int u = x >> shift, b = x & (bs - 1), vb = y >> shift, i = y & (bs - 1);
	unsigned l00 = bl[vb * lw + u], l01 = bl[vb * lw + u + 1];
	unsigned l10 = bl[(vb + 1) * lw + u], l11 = bl[(vb + 1) * lw + u + 1];
	int lightleftstep = (int)((l10 - l00) >> shift);		/* unsigned >> : r_surf.c:202 (H4) */
	int lightrightstep = (int)((l11 - l01) >> shift);
	int lightleft = rc_addw ((int)l00, rc_mulw (i, lightleftstep));
	int lightright = rc_addw ((int)l01, rc_mulw (i, lightrightstep));
	int lighttemp = (int)((unsigned)lightleft - (unsigned)lightright);
	int lightstep = lighttemp >> shift;
	int light = rc_addw (lightright, rc_mulw (bs - 1 - b, lightstep));
	/* source texel with the reference's tiling (texture sizes and texturemins are multiples
	 * of the block size, so its "wrap to 0" equals a modulo) */
	const int tw = job->tw, th = job->th;
	int sx = (job->soffset + u * bs) % tw + b;
	int sy = (job->toffset + y) % th;
	uint8_t pix = w->texels[job->srcofs + sy * tw + sx];
	return w->colormap[(light & 0xFF00) + pix];
}


/* D_Sky_uv_To_st, d_sky.c:34-52 */
RC_HD void rc_sky_uv_to_st (const rc_view_t *vw, int u, int v, int *ps, int *pt)
{
	float wu = (u - vw->xcenter) / vw->xscale;
	float wv = (vw->ycenter - v) / vw->yscale;
	float end[3];
	end[0] = vw->vpn[0] + wu*vw->vright[0] + wv*vw->vup[0];
	end[1] = vw->vpn[1] + wu*vw->vright[1] + wv*vw->vup[1];
	end[2] = vw->vpn[2] + wu*vw->vright[2] + wv*vw->vup[2];
	end[2] *= 3;
	float length = end[0]*end[0] + end[1]*end[1] + end[2]*end[2];
	if (length)
	{
		length = sqrt ((double)length);
		float ilength = 1/length;
		end[0] *= ilength;
		end[1] *= ilength;
		end[2] *= ilength;
	}
	*ps = rc_f2i ((vw->skyshift + 6*(128/2-1)*end[0]) * 0x10000);
	*pt = rc_f2i ((vw->skyshift + 6*(128/2-1)*end[1]) * 0x10000);
}

/* one texel of `newsky` as R_MakeSky composes it (r_sky.c:68-105), without materialising it */
RC_HD uint8_t rc_sky_texel (const rc_world_t *w, const rc_view_t *vw, int s, int t)
{
	const uint8_t *skydata = w->texels + w->skytexofs;
	int xshift2 = rc_f2i (vw->skyshift), yshift2 = xshift2;
	int xshift1 = xshift2 >> 1, yshift1 = yshift2 >> 1;
	int y = (t >> 16) & 127, x = (s >> 16) & 127;
	uint8_t b1 = skydata[((y + yshift1) & 127) * 256 + ((x + xshift1) & 127)];
	if (b1)
		return b1;
	return skydata[((y + yshift2) & 127) * 256 + 128 + ((x + xshift2) & 127)];
}

/* ------------------------------------------------------------------ stage 6: draw ------ */
#ifdef __CUDACC__
RC_HD rc_span_t rc_load_span (const rc_span_t *p)
{
	const uint4 a = __ldg (reinterpret_cast<const uint4 *>(p)), b = __ldg (reinterpret_cast<const uint4 *>(p) + 1);
	rc_span_t s;
	s.u = (int16_t)(a.x & 0xFFFF); s.count = (int16_t)(a.x >> 16);
	s.pixofs = (int)a.y; s.segbase = (int)a.z; s.gsurf = (int)a.w;
	s.iziorg = (int)b.x; s.izistep = (int)b.y; s.cacheofs = (int)b.z; s.cwk = b.w;
	return s;
}
RC_HD rc_bnd_t rc_load_bnd (const rc_bnd_t *p)
{
	const uint2 a = __ldg (reinterpret_cast<const uint2 *>(p));
	rc_bnd_t r;
	r.s = (int)a.x; r.t = (int)a.y;
	return r;
}
RC_HD uint32_t rc_zpair (uint32_t a, uint32_t b) { return __byte_perm (a, b, 0x5410); }
#else
RC_HD rc_span_t rc_load_span (const rc_span_t *p) { return *p; }
RC_HD rc_bnd_t rc_load_bnd (const rc_bnd_t *p) { return *p; }
RC_HD uint32_t rc_zpair (uint32_t a, uint32_t b) { return (a & 0xFFFFu) | (b << 16); }
#endif

/* subdivision k of a span from its boundary records (rp = the span's record k; last: k is the span's last subdivision) */
RC_HD rc_sub_t rc_sub_of (const rc_bnd_t *rp, int last)
{
	const rc_bnd_t b0 = rc_load_bnd (rp), b1 = rc_load_bnd (rp + 1);
	rc_sub_t r;
	r.s = b0.s; r.t = b0.t;
	if (last)
	{
		const rc_bnd_t st = rc_load_bnd (rp + 2);
		r.sstep = st.s; r.tstep = st.t;
	}
	else
	{
		r.sstep = (b1.s - b0.s) >> 3; r.tstep = (b1.t - b0.t) >> 3;
	}
	return r;
}

/* a / d truncated toward zero as C's `/` does, for d in 0 .. 7 (0 and negative: 0 -- the reference does not divide when the
 * last subdivision has one pixel) and any 32-bit a, without a division: |a| * ceil (2^34 / d) >> 34 is exact while
 * |a| * (d - 1) < 2^34, i.e. for every |a| <= 2^31 */
RC_HD int rc_div_small (int a, int d)
{
	if (d <= 1)
		return d == 1 ? a : 0;
	const unsigned long long m = d == 2 ? 0x200000000ull : d == 3 ? 0x155555556ull : d == 4 ? 0x100000000ull :
				     d == 5 ? 0xCCCCCCCDull : d == 6 ? 0xAAAAAAABull : 0x92492493ull;
	const unsigned ua = a < 0 ? 0u - (unsigned)a : (unsigned)a;
	const unsigned q = (unsigned)(((unsigned long long)ua * m) >> 34);
	return a < 0 ? -(int)q : (int)q;
}

/* One thread per span SEGMENT (8 subdivisions = 64 pixels), before drawing: the float side of
 * D_DrawSpans8 (d_scan.c:278-365) and the set-up of D_DrawZSpans (d_scan.c:399-413), once per
 * subdivision instead of once per pixel block.  With B[b] = the 16.16 (s, t) at subdivision boundary b,
 *        b = 0              span start, clamped to [0, bbextent]
 *        0 < b < nsub       after b steps of 8 pixels, clamped to [8, bbextent]   (`snext` of subdivision b-1)
 *        b = nsub           the span's last pixel: nsub-1 steps of 8, then left-1 single steps
 * subdivision k is {B[k], step}: step = (B[k+1] - B[k]) >> 3 (d_scan.c:334-335), and for the last subdivision
 * (B[nsub] - B[nsub-1]) / (left - 1) truncated (d_scan.c:364-365).  What is stored per span is B[0] .. B[nsub] and
 * that one divided step (rc_bnd_t, 8 B each: 1 B per pixel where records of {B[k], step} were 2): the draw kernels
 * take the other steps from two neighbouring records and never divide.  Segment 0 also completes the span record with
 * what the draw kernel needs of the surface: 1/z start and step, texel block and pitch, kind.
 * The accumulators of D_DrawSpans8 are sequentially rounded float sums, so boundary b costs b
 * adds: the segment replays the cheap adds from the span start and does the divide / convert /
 * clamp only for its own boundaries. */
typedef struct {		/* what the segment kernel reads of a surface: five 16-byte groups of rc_surf_t */
	float nearzi, d_ziorigin, d_zistepu, d_zistepv;
	float d_sdivzstepv, d_tdivzstepv, d_sdivzorigin, d_tdivzorigin;
	int   cacheofs, cachewidth, srckind, drawflags;
	float d_sdivzstepu, d_tdivzstepu, d_zistepu_draw; int izistep;
	int   sadjust, tadjust, bbextents, bbextentt;
} rc_segsurf_t;

RC_HD rc_segsurf_t rc_load_segsurf (const rc_surf_t *s)
{
	rc_segsurf_t d;
#ifdef __CUDACC__
	const uint4 *q = reinterpret_cast<const uint4 *>(s);
	const uint4 g1 = q[1], g3 = q[3], g4 = q[4], g5 = q[5], g6 = q[6];
	d.nearzi = __uint_as_float (g1.x); d.d_ziorigin = __uint_as_float (g1.y); d.d_zistepu = __uint_as_float (g1.z); d.d_zistepv = __uint_as_float (g1.w);
	d.d_sdivzstepv = __uint_as_float (g3.x); d.d_tdivzstepv = __uint_as_float (g3.y); d.d_sdivzorigin = __uint_as_float (g3.z); d.d_tdivzorigin = __uint_as_float (g3.w);
	d.cacheofs = (int)g4.x; d.cachewidth = (int)g4.y; d.srckind = (int)g4.z; d.drawflags = (int)g4.w;
	d.d_sdivzstepu = __uint_as_float (g5.x); d.d_tdivzstepu = __uint_as_float (g5.y); d.d_zistepu_draw = __uint_as_float (g5.z); d.izistep = (int)g5.w;
	d.sadjust = (int)g6.x; d.tadjust = (int)g6.y; d.bbextents = (int)g6.z; d.bbextentt = (int)g6.w;
#else
	d.nearzi = s->nearzi; d.d_ziorigin = s->d_ziorigin; d.d_zistepu = s->d_zistepu; d.d_zistepv = s->d_zistepv;
	d.d_sdivzstepv = s->d_sdivzstepv; d.d_tdivzstepv = s->d_tdivzstepv; d.d_sdivzorigin = s->d_sdivzorigin; d.d_tdivzorigin = s->d_tdivzorigin;
	d.cacheofs = s->cacheofs; d.cachewidth = s->cachewidth; d.srckind = s->pad; d.drawflags = s->drawflags;
	d.d_sdivzstepu = s->d_sdivzstepu; d.d_tdivzstepu = s->d_tdivzstepu; d.d_zistepu_draw = s->d_zistepu_draw; d.izistep = s->izistep;
	d.sadjust = s->sadjust; d.tadjust = s->tadjust; d.bbextents = s->bbextents; d.bbextentt = s->bbextentt;
#endif
	return d;
}

/* `stage` (GPU): the thread's eight records go to this (shared-memory) slot as four 16-byte pairs, pair i at index
 * i ^ swz, and the caller copies them out when the function returns 1 (0: nothing to write -- not a textured span, or a
 * segment past the last subdivision) */
RC_HD int rc_span_seg (const rc_world_t *w, const rc_frame_t *f, int seg, rc_bnd_t *stage = 0, int swz = 0)
{
	const int si = f->segspan[seg];
	rc_span_t *spp = &f->spans[si];
	const rc_span_t sp = rc_load_span (spp);
	const int count = sp.count;
	const rc_surf_t *sfp = f->surfs + sp.gsurf;
	const rc_segsurf_t s = rc_load_segsurf (sfp);
	const int sidx = seg - sp.segbase;
	const float du = (float)sp.u, dv = (float)(int)(sp.cwk >> 20);
	const float zistepu = s.d_zistepu, zistepv = s.d_zistepv, ziorigin = s.d_ziorigin;
	const int texels = s.srckind == RC_SRC_POOL || s.srckind == RC_SRC_TEXELS;
	if (sidx == 0)
	{
		double zid = ziorigin + dv*zistepv + du*zistepu;
		if (s.drawflags & RC_SURF_DRAWSKYBOX)
		{
			/* d_edge.c:228-234: d_ziorigin = -0.9, no steps: "effectively at infinity" */
			const float bgzi = -0.9, zero = 0;
			zid = bgzi + dv*zero + du*zero;
		}
		/* the surface's cache block: creators of a block stored its pool offset in the surface-prep
		 * stage, every other user of the same (face, mip, texture, light) key looks it up now */
		int cacheofs = s.cacheofs;
		if (s.srckind == RC_SRC_POOL && sfp->miplevel >= 0)
		{
			cacheofs = f->cachevals[sfp->miplevel];
			/* 0 is never a valid block (the pool starts after the sky-box pictures): the block's creator -- in this
			 * batch or an earlier one -- ran out of pool.  Every batch that meets such a key reports it, so the host
			 * flushes the cache (D_FlushCaches) and the batch is drawn again */
			if (cacheofs == 0)
				RC_ATOMIC_OR_I (f->status, RC_STAT_CACHE_POOL);
		}
		uint32_t kind = RC_KIND_POOL;
		if (s.srckind == RC_SRC_TEXELS) kind = (s.drawflags & RC_SURF_DRAWTURB) ? RC_KIND_TURB : RC_KIND_TEXELS;
		else if (s.srckind == RC_SRC_SKY) kind = RC_KIND_SKY;
		else if (s.srckind == RC_SRC_SOLID) kind = RC_KIND_SOLID;
		const uint32_t cwk = (sp.cwk & 0xFFF80000u) | (kind << 16) | ((uint32_t)s.cachewidth & 0xFFFFu);
		const int iziorg = rc_addw (rc_d2i (zid * 0x8000 * 0x10000), -rc_mulw (sp.pixofs, s.izistep));
#ifdef __CUDACC__
		*(reinterpret_cast<uint4 *>(spp) + 1) = make_uint4 ((uint32_t)iziorg, (uint32_t)s.izistep, (uint32_t)cacheofs, cwk);
#else
		spp->iziorg = iziorg; spp->izistep = s.izistep; spp->cacheofs = cacheofs; spp->cwk = cwk;
#endif
	}
	if (!texels)
		return 0;
	if ((sidx << 3) >= ((count + 7) >> 3))
		return 0;		/* the segment only holds records that follow the last subdivision: its neighbour writes them */
	const float sstepu = s.d_sdivzstepu, tstepu = s.d_tdivzstepu;
	const float sdivz8stepu = sstepu * 8, tdivz8stepu = tstepu * 8, zi8stepu = zistepu * 8;
	float sdivz = s.d_sdivzorigin + dv*s.d_sdivzstepv + du*sstepu;
	float tdivz = s.d_tdivzorigin + dv*s.d_tdivzstepv + du*tstepu;
	float zi = ziorigin + dv*zistepv + du*zistepu;
	float psdivz = sdivz, ptdivz = tdivz, pzi = zi;		/* the accumulators one step back */
	for (int q = sidx << 3; q > 0; q--)
	{
		psdivz = sdivz; ptdivz = tdivz; pzi = zi;
		sdivz += sdivz8stepu;
		tdivz += tdivz8stepu;
		zi += zi8stepu;
	}
	const int nsub = (count + 7) >> 3;
	const float m1 = (float)(count - ((nsub - 1) << 3) - 1);	/* spancountminus1 of the last subdivision */
	const int sadjust = s.sadjust, tadjust = s.tadjust, bbextents = s.bbextents, bbextentt = s.bbextentt;
	/* first the (cheap, serial) accumulator values of the nine boundaries, then nine independent
	 * divide / convert / clamp chains the scheduler can overlap */
	float as[RC_SEG_BND], at[RC_SEG_BND], az[RC_SEG_BND];
	#pragma unroll
	for (int i = 0; i < RC_SEG_BND; i++)
	{
		const int b = (sidx << 3) + i;
		if (b == nsub)
		{
			as[i] = psdivz + sstepu * m1; at[i] = ptdivz + tstepu * m1; az[i] = pzi + zistepu * m1;
		}
		else
		{
			as[i] = sdivz; at[i] = tdivz; az[i] = zi;
		}
		psdivz = sdivz; ptdivz = tdivz; pzi = zi;
		sdivz += sdivz8stepu;
		tdivz += tdivz8stepu;
		zi += zi8stepu;
	}
	int st[RC_SEG_BND * 2];
	#pragma unroll
	for (int i = 0; i < RC_SEG_BND; i++)
	{
		const int b = (sidx << 3) + i;
		const int lo = b ? 8 : 0;
		const float z = (float)0x10000 / az[i];
		int ss = rc_addw (rc_f2i (as[i] * z), sadjust);
		int tt = rc_addw (rc_f2i (at[i] * z), tadjust);
		if (ss > bbextents) ss = bbextents; else if (ss < lo) ss = lo;
		if (tt > bbextentt) tt = bbextentt; else if (tt < lo) tt = lo;
		st[2 * i] = ss; st[2 * i + 1] = tt;
	}
	/* this segment's records: boundaries 8 sidx .. 8 sidx + 7 that are subdivision starts (b < nsub); the segment that
	 * holds the LAST subdivision also writes what follows it -- B[nsub] (the last pixel's coordinates) and that
	 * subdivision's per-pixel step, the one truncating division of a span (d_scan.c:364-365) -- into its own block where
	 * they fit (slots il + 1, il + 2) and into the next segment's otherwise (that segment, all of whose boundaries are
	 * >= nsub, returned above and writes nothing).  A segment that writes at all writes its eight slots. */
	rc_bnd_t *out = f->bnd + (size_t)seg * 8;
	{
		const int il = nsub - 1 - (sidx << 3);			/* the last subdivision's place in this segment */
		if (il >= 0 && il < 8)
		{
			const int lastn = count - ((nsub - 1) << 3);		/* its pixels */
			int ds_ = 0, dt_ = 0, es_ = 0, et_ = 0;
			#pragma unroll
			for (int i = 0; i < 8; i++)
				if (i == il) { es_ = st[2 * i + 2]; et_ = st[2 * i + 3]; ds_ = es_ - st[2 * i]; dt_ = et_ - st[2 * i + 1]; }
			const int sstep = rc_div_small (ds_, lastn - 1), tstep = rc_div_small (dt_, lastn - 1);
			#pragma unroll
			for (int i = 1; i < 8; i++)
			{
				if (i == il + 1) { st[2 * i] = es_; st[2 * i + 1] = et_; }
				if (i == il + 2) { st[2 * i] = sstep; st[2 * i + 1] = tstep; }
			}
			if (il >= 6)
			{
				/* (record 8 is the next block's slot 0) */
#ifdef __CUDACC__
				if (il == 7) reinterpret_cast<uint2 *>(out)[8] = make_uint2 ((uint32_t)es_, (uint32_t)et_);
				reinterpret_cast<uint2 *>(out)[il + 2] = make_uint2 ((uint32_t)sstep, (uint32_t)tstep);
#else
				if (il == 7) { out[8].s = es_; out[8].t = et_; }
				out[il + 2].s = sstep; out[il + 2].t = tstep;
#endif
			}
		}
	}
#ifdef __CUDACC__
	if (stage)
	{
		#pragma unroll
		for (int i = 0; i < 8; i += 2)
			reinterpret_cast<uint4 *>(stage)[(i >> 1) ^ swz] = make_uint4 ((uint32_t)st[2*i], (uint32_t)st[2*i+1], (uint32_t)st[2*i+2], (uint32_t)st[2*i+3]);
		return 1;
	}
	#pragma unroll
	for (int i = 0; i < 8; i += 2)
		reinterpret_cast<uint4 *>(out)[i >> 1] = make_uint4 ((uint32_t)st[2*i], (uint32_t)st[2*i+1], (uint32_t)st[2*i+2], (uint32_t)st[2*i+3]);
#else
	for (int i = 0; i < 8; i++)
	{
		out[i].s = st[2*i]; out[i].t = st[2*i+1];
	}
#endif
	return 1;
}

/* The part of span `sp` that lies inside one 8-pixel window, general form (D_DrawSpans8 d_scan.c:254-386 /
 * Turbulent8 :121-251 / D_DrawSkyScans8 d_sky.c:58 / D_DrawSolidSurface d_edge.c:72-90, fused with
 * D_DrawZSpans d_scan.c:388-439).  Window pixel j is span pixel rel + j; rel may be negative (the span
 * starts inside the window).  Colour and z (as the 16-bit value the reference stores) of the window
 * pixels that belong to the span are written to px[] / zs[]; the others are left alone.  Returns the
 * mask of the pixels written.
 * The reference's 8-pixel subdivision is span relative, so the window touches subdivision
 * k = rel >> 3 from pixel a = rel & 7 on, and subdivision k+1 when it runs past it.  Everything is
 * computed for 8 pixels and predicated on the mask: no per-count code paths. */
RC_HD unsigned rc_cell_part (const rc_world_t *w, const rc_frame_t *f, const rc_span_t sp, int rel, uint32_t *px, uint32_t *zs)
{
	const int lo = rel < 0 ? -rel : 0;
	int hi = sp.count - rel;
	if (hi > 8) hi = 8;
	if (hi <= lo)
		return 0u;
	const unsigned vmask = (0xFFu >> (8 - hi)) & (0xFFu << lo);		/* bit j: window pixel j belongs to the span */
	const int kind = (int)((sp.cwk >> 16) & 7u);
	const int cachewidth = (int)(sp.cwk & 0xFFFFu);
	const int spv = (int)(sp.cwk >> 20);
	const int pix = sp.pixofs + rel;

	/* ---- z (D_DrawZSpans) ---- */
	{
		const int izistep = sp.izistep;
		int zz[8];
		zz[0] = rc_addw (sp.iziorg, rc_mulw (pix, izistep));
		#pragma unroll
		for (int j = 1; j < 8; j++)
			zz[j] = rc_addw (zz[j-1], izistep);
		const int prev0 = rc_addw (zz[0], -izistep);
		uint32_t zt[8];
		#pragma unroll
		for (int j = 0; j < 8; j++)
			zt[j] = ((uint32_t)zz[j] >> 16);
		if ((prev0 | zz[0] | zz[1] | zz[2] | zz[3] | zz[4] | zz[5] | zz[6] | zz[7]) < 0)
		{
			/* a negative 1/z in reach: the second short of each 32-bit store of the reference gets the
			 * sign extension of the first OR-ed in (d_scan.c:421-427: ltemp = izi >> 16 on a signed izi) */
			const int view = sp.gsurf / f->surfcap;
			const int a0 = (sp.pixofs - view * f->width * f->height) & 1;	/* (zwidth*v + u) & 1: the short stored before the pairs */
			const int doublecount = (sp.count - a0) >> 1;
			int prev = prev0;
			#pragma unroll
			for (int j = 0; j < 8; j++)
			{
				const int r = rel + j - a0;
				if (prev < 0 && r >= 0 && (r & 1) && (r >> 1) < doublecount)
					zt[j] = 0xFFFFu;
				prev = zz[j];
			}
		}
		#pragma unroll
		for (int j = 0; j < 8; j++)
			if (vmask & (1u << j))
				zs[j] = zt[j];
	}

	/* ---- colour ---- */
	if (kind == RC_KIND_SOLID)
	{
		#pragma unroll
		for (int j = 0; j < 8; j++)
			if (vmask & (1u << j))
				px[j] = (uint32_t)(cachewidth & 0xFF);
	}
	else if (kind == RC_KIND_SKY)
	{
		/* D_DrawSkyScans8, d_sky.c:58-131: 32-pixel subdivision, end points in closed form; a window
		 * touches at most two of a span's 32-pixel runs */
		const rc_view_t *vw = &f->views[sp.gsurf / f->surfcap];
		int cur = -1, ss = 0, tt = 0, sstep = 0, tstep = 0;
		unsigned long long packed = 0ull;		/* a rolled loop must not index px[] (registers) */
		#pragma unroll 1
		for (int j = lo; j < hi; j++)
		{
			const int pos = rel + j;
			const int b32 = pos >> 5;			/* which 32-pixel run of the span */
			if (b32 != cur)
			{
				const int u0 = sp.u + (b32 << 5);
				const int remaining = sp.count - (b32 << 5);	/* pixels from the start of this run */
				const int spancount = remaining >= 32 ? 32 : remaining;
				int snext = 0, tnext = 0;
				cur = b32; sstep = 0; tstep = 0;
				rc_sky_uv_to_st (vw, u0, spv, &ss, &tt);
				if (remaining - spancount)
				{
					rc_sky_uv_to_st (vw, u0 + spancount, spv, &snext, &tnext);
					sstep = (snext - ss) >> 5;
					tstep = (tnext - tt) >> 5;
				}
				else
				{
					const int spancountminus1 = spancount - 1;
					if (spancountminus1 > 0)
					{
						rc_sky_uv_to_st (vw, u0 + spancountminus1, spv, &snext, &tnext);
						sstep = (snext - ss) / spancountminus1;
						tstep = (tnext - tt) / spancountminus1;
					}
				}
				const int within = pos & 31;
				ss = rc_addw (ss, rc_mulw (within, sstep));
				tt = rc_addw (tt, rc_mulw (within, tstep));
			}
			packed |= (unsigned long long)rc_sky_texel (w, vw, ss, tt) << (8 * j);
			ss = rc_addw (ss, sstep);
			tt = rc_addw (tt, tstep);
		}
		#pragma unroll
		for (int j = 0; j < 8; j++)
			if (vmask & (1u << j))
				px[j] = (uint32_t)(packed >> (8 * j)) & 0xFFu;
	}
	else
	{
		const int turb = kind == RC_KIND_TURB;
		const uint8_t *pbase = (kind == RC_KIND_POOL ? f->cachepool : w->texels) + sp.cacheofs;
		/* a span that starts inside the window: subdivision 0 from window pixel -rel on (a < 0, no crossing) */
		const int k = rel < 0 ? 0 : rel >> 3, a = rel < 0 ? rel : rel & 7;
		const rc_bnd_t *rp = f->bnd + (size_t)sp.segbase * 8 + k;
		const int lastsub = ((sp.count + 7) >> 3) - 1;
		const rc_sub_t r0 = rc_sub_of (rp, k == lastsub);
		rc_sub_t r1 = r0;
		if (a + hi > 8)		/* the window runs on into subdivision k+1 */
			r1 = rc_sub_of (rp + 1, k + 1 == lastsub);
		int sstep = r0.sstep, tstep = r0.tstep;
		int ss = rc_addw (r0.s, rc_mulw (a, sstep)), tt = rc_addw (r0.t, rc_mulw (a, tstep));
		const int *turbtab = w->sintable;
		if (turb)
			turbtab += f->views[sp.gsurf / f->surfcap].turbphase;
		#pragma unroll
		for (int j = 0; j < 8; j++)
		{
			if (a + j == 8)
			{
				ss = r1.s; tt = r1.t; sstep = r1.sstep; tstep = r1.tstep;
			}
			if (vmask & (1u << j))
			{
				if (!turb)
					px[j] = RC_LDG (&pbase[(ss >> 16) + (tt >> 16) * cachewidth]);
				else
				{
					const int sm = ss & ((RC_CYCLE << 16) - 1), tm = tt & ((RC_CYCLE << 16) - 1);
					const int sturb = ((sm + RC_LDG (&turbtab[(tm >> 16) & (RC_CYCLE - 1)])) >> 16) & 63;
					const int tturb = ((tm + RC_LDG (&turbtab[(sm >> 16) & (RC_CYCLE - 1)])) >> 16) & 63;
					px[j] = RC_LDG (&pbase[(tturb << 6) + sturb]);
				}
			}
			ss = rc_addw (ss, sstep);
			tt = rc_addw (tt, tstep);
		}
	}
	return vmask;
}

/* stores the window pixels of `vmask`: one aligned 8-byte colour store and one aligned 16-byte z store
 * when the window is complete and aligned, single pixels otherwise */
RC_HD void rc_store_part (uint8_t *pdest, int16_t *pz, unsigned vmask, const uint32_t *px, const uint32_t *zs)
{
	if (pz)
	{
#ifdef __CUDACC__
		if (vmask == 0xFFu && ((size_t)pz & 15) == 0)
		{
			uint4 q;
			q.x = rc_zpair (zs[0], zs[1]); q.y = rc_zpair (zs[2], zs[3]); q.z = rc_zpair (zs[4], zs[5]); q.w = rc_zpair (zs[6], zs[7]);
			*reinterpret_cast<uint4 *>(pz) = q;
		}
		else
#endif
		{
			#pragma unroll
			for (int j = 0; j < 8; j++)
				if (vmask & (1u << j))
					pz[j] = (int16_t)zs[j];
		}
	}
#ifdef __CUDACC__
	if (vmask == 0xFFu && ((size_t)pdest & 7) == 0)
	{
		uint2 q;
		q.x = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
		q.y = px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24);
		*reinterpret_cast<uint2 *>(pdest) = q;
	}
	else
#endif
	{
		#pragma unroll
		for (int j = 0; j < 8; j++)
			if (vmask & (1u << j))
				pdest[j] = (uint8_t)px[j];
	}
}

/* colour destination of span pixel `rel`: the frame buffer, or the view's 320x200 warp buffer when it
 * renders under water (r_dowarp; the kernel variant without such views drops the test) */
RC_HD uint8_t *rc_span_dest (const rc_frame_t *f, const rc_span_t sp, int rel, const bool warpviews)
{
	if (warpviews && (sp.cwk & RC_SPAN_WARPBIT))
	{
		const rc_view_t *vw = &f->views[sp.gsurf / f->surfcap];
		return f->warpbuf + (size_t)vw->warp_index * (RC_WARP_W * RC_WARP_H) + RC_WARP_W * (int)(sp.cwk >> 20) + sp.u + rel;
	}
	return f->fb + sp.pixofs + rel;
}

/* one span's share of a window, drawn on its own */
RC_HD void rc_draw_pixels (const rc_world_t *w, const rc_frame_t *f, const rc_span_t sp, int rel, const bool warpviews)
{
	uint32_t px[8], zs[8];
	#pragma unroll
	for (int j = 0; j < 8; j++) { px[j] = 0; zs[j] = 0; }
	const unsigned vmask = rc_cell_part (w, f, sp, rel, px, zs);
	if (vmask)
		rc_store_part (rc_span_dest (f, sp, rel, warpviews), f->zb ? f->zb + sp.pixofs + rel : (int16_t *)0, vmask, px, zs);
}

/* The two kinds of draw work item.
 * Full cells: screen cell `c` of row `row` = view * height + line, covered by one span (the GPU draws
 * the common case -- cache block or texel blob, no negative 1/z, frame-buffer target -- with its own
 * straight-line code, k_draw, and comes here for the rest). */
RC_HD void rc_draw_cell (const rc_world_t *w, const rc_frame_t *f, int row, int c)
{
	const int cw = (f->width + 7) >> 3;
	const int cs = f->cellspan[(size_t)row * cw + c].si;
	if (cs == -1)
		return;
	const int si = RC_CELL_SPAN (cs);
	const rc_span_t sp = rc_load_span (&f->spans[si]);
	rc_draw_pixels (w, f, sp, (c << 3) - sp.u, true);
}

/* Mixed cells: the screen cells that hold a span boundary (or an end of the view rectangle).  They are
 * drawn WHOLE by one owner -- complete cells leave as one aligned colour store and one aligned z store
 * instead of single pixels from two threads.  Two candidate items per span:
 *   which = 0: the cell around the span's first pixel, owned by the first span that starts inside it
 *              (the share of the span before it, which covers the cell's first pixel, comes first);
 *   which = 1: the cell around the span's end, when the span covers that cell's first pixel and no
 *              later span of the row starts inside it (the right end of the view rectangle).
 * Returns 1 and the first span to visit, the cell's column and the row's first pixel index. */
RC_HD int rc_mixed_owner (const rc_frame_t *f, int si, int which, int nspans, int *first, int *cellx, int *rowstart)
{
	const rc_span_t *sp = &f->spans[si];
	const int u = sp->u, count = sp->count, e = u + count;
	if (count <= 0)
		return 0;
	*rowstart = sp->pixofs - u;
	*first = si;
	if (!which)
	{
		if (!(u & 7))
			return 0;
		*cellx = u & ~7;
		if (si > 0)
		{
			const rc_span_t *pv = sp - 1;
			if (pv->pixofs - pv->u == *rowstart && pv->count > 0)
			{
				if (pv->u > *cellx)
					return 0;		/* an earlier span starts inside the same cell */
				*first = si - 1;
			}
		}
		return 1;
	}
	if (!(e & 7) || u > (e & ~7))
		return 0;
	*cellx = e & ~7;
	if (si + 1 < nspans)
	{
		const rc_span_t *nx = sp + 1;
		if (nx->pixofs - nx->u == *rowstart && nx->count > 0 && nx->u < *cellx + 8)
			return 0;			/* the next span starts inside that cell and owns it */
	}
	return 1;
}

RC_HD void rc_draw_mixed_cell (const rc_world_t *w, const rc_frame_t *f, int first, int nspans, int cellx, int rowstart, const bool warpviews)
{
	uint32_t px[8], zs[8];
	#pragma unroll
	for (int j = 0; j < 8; j++) { px[j] = 0; zs[j] = 0; }
	unsigned mask = 0u;
	for (int s = first; s < nspans; s++)
	{
		const rc_span_t sp = rc_load_span (&f->spans[s]);
		if (sp.pixofs - sp.u != rowstart || sp.u >= cellx + 8)
			break;
		if (warpviews && (sp.cwk & RC_SPAN_WARPBIT))
			rc_draw_pixels (w, f, sp, cellx - sp.u, warpviews);	/* warp buffer rows are not cell aligned: span by span */
		else
			mask |= rc_cell_part (w, f, sp, cellx - sp.u, px, zs);
	}
	if (mask)
		rc_store_part (f->fb + rowstart + cellx, f->zb ? f->zb + rowstart + cellx : (int16_t *)0, mask, px, zs);
}

RC_HD void rc_draw_mixed (const rc_world_t *w, const rc_frame_t *f, int si, int which, int nspans, const bool warpviews = true)
{
	int first, cellx, rowstart;
	if (rc_mixed_owner (f, si, which, nspans, &first, &cellx, &rowstart))
		rc_draw_mixed_cell (w, f, first, nspans, cellx, rowstart, warpviews);
}


/* One thread per output pixel of an underwater view: D_WarpScreen (d_scan.c:43-89).
 * (u, v) run over the r_refdef rectangle; like the reference's 4-way unrolled loop the
 * columns are processed up to the next multiple of 4. */
RC_HD void rc_warp_pixel (const rc_world_t *w, const rc_frame_t *f, int vi, int u, int v)
{
	const rc_view_t *vw = &f->views[vi];
	if (vw->warp_index < 0 || v >= vw->rd_h || u >= ((vw->rd_w + 3) & ~3))
		return;
	const int ww = vw->vrect_w, hh = vw->vrect_h;
	const float wratio = ww / (float)vw->rd_w;
	const float hratio = hh / (float)vw->rd_h;
	const int *turb = w->intsintable + vw->turbphase;
	const int vv = v + turb[u & (RC_CYCLE - 1)];		/* row[turb[u&127]] = rowptr[v + ...] */
	const int uu = turb[v & (RC_CYCLE - 1)] + u;		/* col[u] = column[turb[v&127] + u] */
	const int srow = vw->vrect_y + rc_f2i ((float)vv * hratio * hh / (hh + 3 * 2));
	const int scol = vw->vrect_x + rc_f2i ((float)uu * wratio * ww / (ww + 3 * 2));
	const uint8_t *src = f->warpbuf + (size_t)vw->warp_index * (RC_WARP_W * RC_WARP_H);
	const size_t viewpix = (size_t)f->width * f->height;
	const size_t o = (size_t)(vw->rd_y + v) * f->width + vw->rd_x + u;
	if (o < viewpix)
		f->fb[viewpix * vi + o] = src[srow * RC_WARP_W + scol];
}


/* ============================================================ alias models & particles ==== */
#define RC_ALIAS_LEFT_CLIP   0x0001   /* r_shared.h:186-194 */
#define RC_ALIAS_TOP_CLIP    0x0002
#define RC_ALIAS_RIGHT_CLIP  0x0004
#define RC_ALIAS_BOTTOM_CLIP 0x0008
#define RC_ALIAS_Z_CLIP      0x0010
#define RC_ALIAS_ONSEAM      0x0020
#define RC_ALIAS_XY_CLIP_MASK 0x000F
#define RC_ALIAS_Z_CLIP_PLANE 5

/* One thread per (alias instance, vertex): R_AliasTransformAndProjectFinalVerts
 * (r_alias.c:453-490) for trivially accepted models, else R_AliasTransformFinalVert +
 * R_AliasProjectFinalVert + clip flags (r_alias.c:304-338,423-511). */
RC_HD void rc_alias_vertex (const rc_frame_t *f, int ii, int vi_)
{
	const rc_aliasinst_t *in = &f->alias[ii];
	const rc_aliasmodel_t *m = &f->amodels[in->model];
	if (vi_ >= m->numverts)
		return;
	const rc_view_t *vw = &f->views[in->view - f->viewbase];
	const uint8_t *pv1 = f->poseverts + 4 * ((size_t)m->poses_ofs + (size_t)in->pose_old * m->numverts + vi_);
	const uint8_t *pv2 = f->poseverts + 4 * ((size_t)m->poses_ofs + (size_t)in->pose_new * m->numverts + vi_);
	const int *st = f->stverts + 4 * ((size_t)m->stverts_ofs + vi_);
	rc_finalvert_t *fv = &f->finalverts[in->vertbase + vi_];
	const float lerp = in->framelerp;
	float v[3];
	v[0] = pv1[0] + lerp * ((float)pv2[0] - (float)pv1[0]);
	v[1] = pv1[1] + lerp * ((float)pv2[1] - (float)pv1[1]);
	v[2] = pv1[2] + lerp * ((float)pv2[2] - (float)pv1[2]);
	const float *dots = f->anormdots + 256 * in->shadequant;
	const float temp = dots[pv1[3]] + lerp * (dots[pv2[3]] - dots[pv1[3]]);
	int light = rc_f2i (in->ambientlight - temp * in->shadelight);
	if (light < 0)
		light = 0;
	fv->v[2] = st[1];
	fv->v[3] = st[2];
	fv->v[4] = light;
	fv->flags = st[0];
	if (in->trivial_accept)
	{
		const float zi = 1.0 / (rc_dot3 (v, in->xf[2]) + in->xf[2][3]);
		fv->v[5] = rc_f2i (zi);
		fv->v[0] = rc_f2i (((rc_dot3 (v, in->xf[0]) + in->xf[0][3]) * zi) + vw->aliasxcenter);
		fv->v[1] = rc_f2i (((rc_dot3 (v, in->xf[1]) + in->xf[1][3]) * zi) + vw->aliasycenter);
		return;
	}
	fv->av[0] = rc_dot3 (v, in->xf[0]) + in->xf[0][3];
	fv->av[1] = rc_dot3 (v, in->xf[1]) + in->xf[1][3];
	fv->av[2] = rc_dot3 (v, in->xf[2]) + in->xf[2][3];
	if (fv->av[2] < RC_ALIAS_Z_CLIP_PLANE)
	{
		fv->flags |= RC_ALIAS_Z_CLIP;
		fv->v[0] = fv->v[1] = fv->v[5] = 0;
	}
	else
	{
		const float zi = 1.0 / fv->av[2];
		fv->v[5] = rc_f2i (zi * in->ziscale);
		fv->v[0] = rc_f2i ((fv->av[0] * vw->aliasxscale * zi) + vw->aliasxcenter);
		fv->v[1] = rc_f2i ((fv->av[1] * vw->aliasyscale * zi) + vw->aliasycenter);
		if (fv->v[0] < vw->vrect_x) fv->flags |= RC_ALIAS_LEFT_CLIP;
		if (fv->v[1] < vw->vrect_y) fv->flags |= RC_ALIAS_TOP_CLIP;
		if (fv->v[0] > vw->vrectright) fv->flags |= RC_ALIAS_RIGHT_CLIP;
		if (fv->v[1] > vw->vrectbottom) fv->flags |= RC_ALIAS_BOTTOM_CLIP;
	}
}

typedef struct {
	const rc_frame_t *f;
	const rc_view_t *vw;
	const uint8_t *skin;
	const uint8_t *colormap;
	unsigned long long *zplane;
	const int16_t *zworld; int zwidth;	/* the view's z-buffer as the world pass left it */
	int skinwidth, skinlo, skinhi, seamfixupX16;	/* texel indices skinlo .. skinhi-1 (relative to the picture) lie inside the model's image */
	unsigned long long orderbits;	/* (order << 8), OR-ed under the z and over the colour */
	int skinofs, cmapidx, viewidx, zresidx;	/* the same as indices, for the queued form */
	/* GPU: 0 = the thread draws its spans itself; 1 = they go to the span queue, `qrows` slots from `qspan0` and triangle
	 * record `qtri` reserved by the caller; 2 = queue, the triangle reserves for itself (clipped triangles) */
	int qmode, qtri, qrows; long long qspan0;
} rc_polyctx_t;

RC_HD int rc_min3i (int a, int b, int c) { int m = a < b ? a : b; return m < c ? m : c; }
RC_HD int rc_max3i (int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }

/* Span queue of the alias rasteriser (GPU).  A triangle is walked edge by edge by ONE thread (the DDA of
 * D_PolysetScanLeftEdge is serial and cheap), but its pixels -- two dependent loads and one atomic each -- are not drawn
 * there: every scanline's span goes to a queue with the stepping state at its first pixel, and k_alias_spans gives each
 * queued span a thread of its own.  (One thread per triangle doing everything took 9 ms at C3, one warp per large triangle
 * sharing each span's pixels 8.5 ms: 2.5 G warp instructions at 24 % issue rate, the rows of a triangle being walked by
 * 32 lanes that had 10 - 20 pixels to share.) */
#define RC_AQ_SHIFT 36		/* alloc[5]: spans reserved in the low 36 bits, triangle records above (28 bits) */
typedef struct rc_aspantri_s {	/* 64 B: what the spans of one rasterised (sub)triangle share */
	int zistepx, lstepx, ststepxwhole, stfracstep;	/* r_zistepx, r_lstepx, a_ststepxwhole, a_sstepxfrac | a_tstepxfrac << 16 */
	unsigned long long orderbits;			/* draw order, triangle, fan index: under the z and over the colour */
	int skinofs, skinwidth, skinlo, skinhi;
	int cmapidx, viewidx, zresidx, zwidth;
	int pad[2];
} rc_aspantri_t;
typedef struct rc_aspan_s {	/* 32 B */
	int xy;			/* x & 0xFFFF | y << 16 */
	int count;		/* <= 0: reserved but not used */
	int ptex, stfrac;	/* l_ptex, l_sfrac | l_tfrac << 16 at the first pixel */
	int light, zi;
	int tri, pad;
} rc_aspan_t;

/* D_PolysetDrawSpans8 (d_polyse.c:385-447), one span: every pixel the reference would z-test becomes one atomicMax of
 * (z, draw order, colour) into the view's resolve plane */
/* returns the number of fragments in front of the world's z-buffer */
RC_HD int rc_polyset_span (const rc_frame_t *f, const uint8_t *skin, const uint8_t *colormap, unsigned long long *zplane,
	const int16_t *zworld, int zwidth, int skinwidth, int skinlo, int skinhi, unsigned long long orderbits,
	int r_zistepx, int r_lstepx, int a_ststepxwhole, int a_sstepxfrac, int a_tstepxfrac,
	int x, const int y, int lcount, int lptex, int lsfrac, int ltfrac, int llight, int lzi)
{
	const int W = f->width, H = f->height;
	int front = 0;
	const bool yin = (unsigned)y < (unsigned)H;
	const int16_t *zwrow = zworld + (size_t)(yin ? y : 0) * zwidth;
	unsigned long long *zprow = zplane + (size_t)(yin ? y : 0) * W;
	/* four pixels per round, so that their four chains of dependent reads -- world z, skin, colormap, the plane's current
	 * value -- wait side by side (one pixel per round, the thread sat out every chain in turn: 2.4 ms at C3).  The stepping
	 * between the pixels is the reference's, in its order. */
	while (lcount > 0)
	{
		const int nb = lcount < 4 ? lcount : 4;
		int px[4], pzi[4], pti[4], plight[4];
		bool in[4];
		#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			px[j] = x; pzi[j] = lzi >> 16; pti[j] = lptex; plight[j] = llight;
			in[j] = j < nb && yin && (unsigned)x < (unsigned)W;
			if (j < nb)
			{
				if (!in[j])
					RC_ALIAS_RANGE_HIT (f);
				x++;
				lzi = rc_addw (lzi, r_zistepx);
				llight = rc_addw (llight, r_lstepx);
				lptex += a_ststepxwhole;
				lsfrac += a_sstepxfrac;
				lptex += lsfrac >> 16;
				lsfrac &= 0xFFFF;
				ltfrac += a_tstepxfrac;
				if (ltfrac & 0x10000)
				{
					lptex += skinwidth;
					ltfrac &= 0xFFFF;
				}
			}
		}
		lcount -= nb;
		/* d_polyse.c:422 tests the fragment against the z-buffer before it reads the skin; what the world pass left
		 * there is a lower bound of what the reference finds: a fragment behind it never lands */
		int zw[4];
		#pragma unroll
		for (int j = 0; j < 4; j++)
			zw[j] = in[j] ? (int)zwrow[px[j]] : 0x7FFFFFFF;
		bool pass[4];
		int sk[4];
		#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			pass[j] = in[j] && pzi[j] >= zw[j];
			int ti = pti[j];
			if (ti < skinlo || ti >= skinhi)
			{
				if (pass[j])
				{
					/* outside the model's whole memory block: the reference reads a neighbour on its heap */
					RC_ALIAS_RANGE_HIT (f);
					RC_DEBUG_PRINT ("skin index %d outside the model image %d .. %d (x %d y %d)\n", ti, skinlo, skinhi, px[j], y);
				}
				ti = 0;
			}
			sk[j] = pass[j] ? (int)skin[ti] : 0;
		}
		unsigned long long key[4], cur[4];
		#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			int ci = sk[j] + (plight[j] & 0xFF00);
			if (ci >= 64 * 256)
			{
				if (pass[j])
				{
					RC_ALIAS_RANGE_HIT (f);
					RC_DEBUG_PRINT ("colormap index %d (light %d)\n", ci, plight[j]);
				}
				ci &= 0x3FFF;
			}
			key[j] = pass[j] ? (((unsigned long long)(unsigned)(pzi[j] + 32768) << 48) | orderbits | colormap[ci]) : 0ull;
			/* the planes only grow: a fragment that does not beat the value read now cannot beat the final one */
			cur[j] = pass[j] ? RC_LDCG (&zprow[px[j]]) : ~0ull;
		}
		#pragma unroll
		for (int j = 0; j < 4; j++)
		{
			front += pass[j];
			if (key[j] > cur[j])
				RC_ATOMIC_MAX_ULL_BLIND (&zprow[px[j]], key[j]);
		}
	}
	return front;
}

/* one queued span (k_alias_spans) */
RC_HD int rc_alias_queued_span (const rc_frame_t *f, const rc_aspan_t &sp, const rc_aspantri_t &t)
{
	const size_t viewpix = (size_t)f->width * f->height;
	return rc_polyset_span (f, f->skinpixels + t.skinofs, f->colormaps + (size_t)t.cmapidx * 16384, f->zres + (size_t)t.zresidx * viewpix,
		f->zb + viewpix * t.viewidx, t.zwidth, t.skinwidth, t.skinlo, t.skinhi, t.orderbits,
		t.zistepx, t.lstepx, t.ststepxwhole, t.stfracstep & 0xFFFF, (int)((unsigned)t.stfracstep >> 16),
		(int)(int16_t)(sp.xy & 0xFFFF), sp.xy >> 16, sp.count, sp.ptex, sp.stfrac & 0xFFFF, (int)((unsigned)sp.stfrac >> 16), sp.light, sp.zi);
}

typedef struct { int mode, tri, rows, n, reserved; long long span0; } rc_aqueue_t;

/* D_PolysetSetUpForLineScan, d_polyse.c:279-310.  adivtab (ref_soft/adivtab.h) is the table of
 * floor quotients / remainders for numerators and denominators in [-15,16] ({0,0} for a zero
 * denominator); FloorDivMod (common/mathlib.c:416-455) handles the rest. */
RC_HD void rc_line_scan_setup (int su, int sv, int eu, int ev, int *ubasestep, int *errorterm, int *adjup, int *adjdown)
{
	const int tm = eu - su, tn = ev - sv;
	*errorterm = -1;
	if (((tm <= 16) && (tm >= -15)) && ((tn <= 16) && (tn >= -15)))
	{
		int q = 0, r = 0;
		if (tn != 0)
		{
			q = tm / tn; r = tm - q * tn;
			if (r != 0 && ((r < 0) != (tn < 0))) { q--; r += tn; }	/* floor division */
		}
		*ubasestep = q; *adjup = r; *adjdown = tn;
	}
	else
	{
		const double numer = (double)tm, denom = (double)tn;
		int q, r;
		double x;
		if (numer >= 0.0)
		{
			x = floor (numer / denom);
			q = rc_d2i (x);
			r = rc_d2i (floor (numer - (x * denom)));
		}
		else
		{
			x = floor (-numer / denom);
			q = -rc_d2i (x);
			r = rc_d2i (floor (-numer - (x * denom)));
			if (r != 0)
			{
				q--;
				r = rc_d2i (denom) - r;
			}
		}
		*ubasestep = q; *adjup = r; *adjdown = rc_d2i (denom);
	}
}

/* the twelve left/right edge orders of d_polyse.c:50-63: {numleft, l0, l1, l2, numright, r0, r1, r2} */
RC_HD void rc_edge_table (int idx, int *t)
{
	const unsigned tab[12] = {
		/* each entry: nibbles numleft,l0,l1,l2,numright,r0,r1,r2 (3 = unused) */
		0x10232012u, 0x21021123u, 0x10231123u, 0x11032120u, 0x20211013u, 0x12131203u,
		0x12132201u, 0x22101203u, 0x11031123u, 0x12131013u, 0x11031203u, 0x10231013u };
	const unsigned e = tab[idx];
	for (int i = 0; i < 8; i++)
		t[i] = (e >> (28 - 4 * i)) & 15;
}

/* One triangle of D_DrawNonSubdiv (d_polyse.c:139-203): backface test, seam fix-up,
 * D_PolysetSetEdgeTable (:713-767), D_RasterizeAliasPolySmooth (:454-706) with
 * D_PolysetCalcGradients (:320-375), D_PolysetScanLeftEdge (:210-271) and
 * D_PolysetDrawSpans8 (:385-447).  The reference scans the whole left side into span
 * packages and then walks the right side; the two sides are independent, so here they are
 * stepped together scanline by scanline.  Each pixel that the reference would z-test becomes
 * one atomicMax of (z, draw order, colour) into the view's resolve plane. */
RC_HD void rc_polyset_triangle_body (const rc_polyctx_t *c, const int *fv0, int fl0, const int *fv1, int fl1,
				const int *fv2, int fl2, int facesfront, int sub, rc_aqueue_t *q)
{
	const rc_frame_t *f = c->f;
	const int d_xdenom = (fv0[1]-fv1[1]) * (fv0[0]-fv2[0]) - (fv0[0]-fv1[0])*(fv0[1]-fv2[1]);
	if (d_xdenom >= 0)
		return;
	int P[3][6];
	for (int i = 0; i < 6; i++) { P[0][i] = fv0[i]; P[1][i] = fv1[i]; P[2][i] = fv2[i]; }
	if (!facesfront)
	{
		if (fl0 & RC_ALIAS_ONSEAM) P[0][2] += c->seamfixupX16;
		if (fl1 & RC_ALIAS_ONSEAM) P[1][2] += c->seamfixupX16;
		if (fl2 & RC_ALIAS_ONSEAM) P[2][2] += c->seamfixupX16;
	}
	/* D_PolysetSetEdgeTable */
	int eti = 0, done = 0;
	if (P[0][1] >= P[1][1])
	{
		if (P[0][1] == P[1][1]) { eti = (P[0][1] < P[2][1]) ? 2 : 5; done = 1; }
		else eti = 1;
	}
	if (!done)
	{
		if (P[0][1] == P[2][1]) { eti = eti ? 8 : 9; done = 1; }
		else if (P[1][1] == P[2][1]) { eti = eti ? 10 : 11; done = 1; }
	}
	if (!done)
	{
		if (P[0][1] > P[2][1]) eti += 2;
		if (P[1][1] > P[2][1]) eti += 4;
	}
	int et[8];
	rc_edge_table (eti, et);
	const int numleft = et[0], numright = et[4];
	const int *plefttop = P[et[1]], *pleftbottom = P[et[2]];
	const int *prighttop = P[et[5]], *prightbottom = P[et[6]];
	const int initialleftheight = pleftbottom[1] - plefttop[1];
	const int initialrightheight = prightbottom[1] - prighttop[1];
	const int skinwidth = c->skinwidth;

	/* D_PolysetCalcGradients */
	int r_lstepx, r_lstepy, r_sstepx, r_sstepy, r_tstepx, r_tstepy, r_zistepx, r_zistepy;
	{
		const float p00_minus_p20 = P[0][0] - P[2][0];
		const float p01_minus_p21 = P[0][1] - P[2][1];
		const float p10_minus_p20 = P[1][0] - P[2][0];
		const float p11_minus_p21 = P[1][1] - P[2][1];
		const float xstepdenominv = 1.0 / (float)d_xdenom;
		const float ystepdenominv = -xstepdenominv;
		float t0, t1;
		t0 = P[0][4] - P[2][4];
		t1 = P[1][4] - P[2][4];
		r_lstepx = rc_d2i (ceil ((double)((t1 * p01_minus_p21 - t0 * p11_minus_p21) * xstepdenominv)));
		r_lstepy = rc_d2i (ceil ((double)((t1 * p00_minus_p20 - t0 * p10_minus_p20) * ystepdenominv)));
		t0 = P[0][2] - P[2][2];
		t1 = P[1][2] - P[2][2];
		r_sstepx = rc_f2i ((t1 * p01_minus_p21 - t0 * p11_minus_p21) * xstepdenominv);
		r_sstepy = rc_f2i ((t1 * p00_minus_p20 - t0* p10_minus_p20) * ystepdenominv);
		t0 = P[0][3] - P[2][3];
		t1 = P[1][3] - P[2][3];
		r_tstepx = rc_f2i ((t1 * p01_minus_p21 - t0 * p11_minus_p21) * xstepdenominv);
		r_tstepy = rc_f2i ((t1 * p00_minus_p20 - t0 * p10_minus_p20) * ystepdenominv);
		t0 = P[0][5] - P[2][5];
		t1 = P[1][5] - P[2][5];
		r_zistepx = rc_f2i ((t1 * p01_minus_p21 - t0 * p11_minus_p21) * xstepdenominv);
		r_zistepy = rc_f2i ((t1 * p00_minus_p20 - t0 * p10_minus_p20) * ystepdenominv);
	}
	const int a_sstepxfrac = r_sstepx & 0xFFFF;
	const int a_tstepxfrac = r_tstepx & 0xFFFF;
	const int a_ststepxwhole = rc_addw (rc_mulw (skinwidth, r_tstepx >> 16), r_sstepx >> 16);

	/* ---- left edge state (d_pdest / d_pz kept as x, y) ---- */
	int lx, ly, l_aspancount, l_ptex, l_sfrac, l_tfrac, l_light, l_zi;
	int l_ubasestep = 0, l_errorterm = 0, l_adjup = 0, l_adjdown = 0, l_countextrastep = 0;
	int l_ptexbase = 0, l_ptexextra = 0, l_sfracbase = 0, l_sfracextra = 0, l_tfracbase = 0, l_tfracextra = 0;
	int l_lightbase = 0, l_lightextra = 0, l_zibase = 0, l_ziextra = 0;
	int l_remaining, l_segment = 1, l_stepping;
#define RC_LEFT_SETUP(top, bottom, height)                                                             \
	do {                                                                                           \
		l_remaining = (height);                                                                \
		l_stepping = (height) != 1;                                                            \
		if (l_stepping) {                                                                      \
			rc_line_scan_setup ((top)[0], (top)[1], (bottom)[0], (bottom)[1], &l_ubasestep, &l_errorterm, &l_adjup, &l_adjdown); \
			const int working_lstepx = (l_ubasestep < 0) ? r_lstepx - 1 : r_lstepx;        \
			l_countextrastep = l_ubasestep + 1;                                            \
			l_ptexbase = rc_addw ((rc_addw (r_sstepy, rc_mulw (r_sstepx, l_ubasestep))) >> 16, \
				rc_mulw ((rc_addw (r_tstepy, rc_mulw (r_tstepx, l_ubasestep))) >> 16, skinwidth)); \
			l_sfracbase = rc_addw (r_sstepy, rc_mulw (r_sstepx, l_ubasestep)) & 0xFFFF;    \
			l_tfracbase = rc_addw (r_tstepy, rc_mulw (r_tstepx, l_ubasestep)) & 0xFFFF;    \
			l_lightbase = rc_addw (r_lstepy, rc_mulw (working_lstepx, l_ubasestep));       \
			l_zibase = rc_addw (r_zistepy, rc_mulw (r_zistepx, l_ubasestep));              \
			l_ptexextra = rc_addw ((rc_addw (r_sstepy, rc_mulw (r_sstepx, l_countextrastep))) >> 16, \
				rc_mulw ((rc_addw (r_tstepy, rc_mulw (r_tstepx, l_countextrastep))) >> 16, skinwidth)); \
			l_sfracextra = rc_addw (r_sstepy, rc_mulw (r_sstepx, l_countextrastep)) & 0xFFFF; \
			l_tfracextra = rc_addw (r_tstepy, rc_mulw (r_tstepx, l_countextrastep)) & 0xFFFF; \
			l_lightextra = rc_addw (l_lightbase, working_lstepx);                          \
			l_ziextra = rc_addw (l_zibase, r_zistepx);                                     \
		}                                                                                      \
	} while (0)

	lx = plefttop[0]; ly = plefttop[1];
	l_aspancount = plefttop[0] - prighttop[0];
	l_ptex = (plefttop[2] >> 16) + (plefttop[3] >> 16) * skinwidth;
	l_sfrac = plefttop[2] & 0xFFFF;
	l_tfrac = plefttop[3] & 0xFFFF;
	l_light = plefttop[4];
	l_zi = plefttop[5];
	RC_LEFT_SETUP (plefttop, pleftbottom, initialleftheight);

	/* ---- right edge state ---- */
	int r_ubasestep, r_errorterm, r_adjup, r_adjdown, r_countextrastep, r_aspancount = 0;
	int r_remaining = initialrightheight, r_segment = 1;
	rc_line_scan_setup (prighttop[0], prighttop[1], prightbottom[0], prightbottom[1], &r_ubasestep, &r_errorterm, &r_adjup, &r_adjdown);
	r_countextrastep = r_ubasestep + 1;
	if (initialleftheight <= 0 || initialrightheight <= 0)
	{
		RC_ALIAS_RANGE_HIT (f);
		return;
	}
	const unsigned long long orderbits = c->orderbits | ((unsigned long long)sub << 8);
#ifdef __CUDACC__
	if (q->mode == 2)
	{
		/* one slot per scanline between the highest and the lowest vertex, one triangle record */
		const int h = rc_max3i (P[0][1], P[1][1], P[2][1]) - rc_min3i (P[0][1], P[1][1], P[2][1]);
		const unsigned long long a = RC_ATOMIC_ADD_ULL (&f->alloc[5], (1ull << RC_AQ_SHIFT) | (unsigned long long)h);
		q->span0 = (long long)(a & ((1ull << RC_AQ_SHIFT) - 1)); q->tri = (int)(a >> RC_AQ_SHIFT); q->rows = h; q->reserved = 1;
	}
	if (q->mode)
	{
		if (q->span0 + q->rows > f->aspancap || q->tri >= f->aspantricap)
			q->mode = 0;		/* queue full: this triangle is drawn here (the host grows the queue for the next batch) */
		else
		{
			rc_aspantri_t *t = &f->aspantris[q->tri];
			uint4 *t4 = reinterpret_cast<uint4 *>(t);
			t4[0] = make_uint4 ((unsigned)r_zistepx, (unsigned)r_lstepx, (unsigned)a_ststepxwhole, (unsigned)a_sstepxfrac | ((unsigned)a_tstepxfrac << 16));
			t4[1] = make_uint4 ((unsigned)orderbits, (unsigned)(orderbits >> 32), (unsigned)c->skinofs, (unsigned)skinwidth);
			t4[2] = make_uint4 ((unsigned)c->skinlo, (unsigned)c->skinhi, (unsigned)c->cmapidx, (unsigned)c->viewidx);
			t4[3] = make_uint4 ((unsigned)c->zresidx, (unsigned)c->zwidth, 0u, 0u);
		}
	}
#endif
	int guard = 0;
	for (;;)
	{
		if (r_remaining == 0)
		{
			if (r_segment == 1 && numright == 2)
			{
				/* d_polyse.c:683-704 */
				const int *oldtop = prighttop, *oldbottom = prightbottom;
				r_aspancount = oldbottom[0] - oldtop[0];
				prighttop = oldbottom;
				prightbottom = P[et[7]];
				r_remaining = prightbottom[1] - prighttop[1];
				rc_line_scan_setup (prighttop[0], prighttop[1], prightbottom[0], prightbottom[1], &r_ubasestep, &r_errorterm, &r_adjup, &r_adjdown);
				r_countextrastep = r_ubasestep + 1;
				r_segment = 2;
				if (r_remaining <= 0)
					break;
			}
			else
				break;
		}
		if (l_remaining == 0)
		{
			if (l_segment == 1 && numleft == 2)
			{
				/* d_polyse.c:576-664: bottom part of the left edge; sfrac/tfrac restart at 0 (H4) */
				const int *top = pleftbottom, *bottom = P[et[3]];
				const int *rt0 = P[et[5]];
				lx = top[0]; ly = top[1];
				l_aspancount = top[0] - rt0[0];
				l_ptex = (top[2] >> 16) + (top[3] >> 16) * skinwidth;
				l_sfrac = 0;
				l_tfrac = 0;
				l_light = top[4];
				l_zi = top[5];
				RC_LEFT_SETUP (top, bottom, bottom[1] - top[1]);
				l_segment = 2;
				if (l_remaining <= 0)
				{
					RC_ALIAS_RANGE_HIT (f);
					break;
				}
			}
			else
			{
				/* the reference would read span packages left over from an earlier triangle */
				RC_ALIAS_RANGE_HIT (f);
				break;
			}
		}
		if (++guard > 4096)
			break;
		/* D_PolysetDrawSpans8 for this scanline */
		const int lcount = r_aspancount - l_aspancount;
		r_errorterm += r_adjup;
		if (r_errorterm >= 0)
		{
			r_aspancount += r_countextrastep;
			r_errorterm -= r_adjdown;
		}
		else
			r_aspancount += r_ubasestep;
		r_remaining--;
		if (lcount > 0)
		{
#ifdef __CUDACC__
			if (q->mode && q->n < q->rows)
			{
				uint4 *s4 = reinterpret_cast<uint4 *>(&f->aspans[q->span0 + q->n]);
				s4[0] = make_uint4 (((unsigned)lx & 0xFFFFu) | ((unsigned)ly << 16), (unsigned)lcount, (unsigned)l_ptex, (unsigned)l_sfrac | ((unsigned)l_tfrac << 16));
				s4[1] = make_uint4 ((unsigned)l_light, (unsigned)l_zi, (unsigned)q->tri, 0u);
				q->n++;
			}
			else
#endif
			rc_polyset_span (f, c->skin, c->colormap, c->zplane, c->zworld, c->zwidth, skinwidth, c->skinlo, c->skinhi, orderbits,
				r_zistepx, r_lstepx, a_ststepxwhole, a_sstepxfrac, a_tstepxfrac, lx, ly, lcount, l_ptex, l_sfrac, l_tfrac, l_light, l_zi);
		}
		else if (lcount < 0)
			RC_ALIAS_RANGE_HIT (f);
		/* D_PolysetScanLeftEdge step */
		l_remaining--;
		if (l_stepping)
		{
			l_errorterm += l_adjup;
			if (l_errorterm >= 0)
			{
				lx += l_ubasestep + 1; ly += 1;		/* d_pdestextrastep = screenwidth + ubasestep + 1 */
				l_aspancount += l_countextrastep;
				l_ptex += l_ptexextra;
				l_sfrac += l_sfracextra;
				l_ptex += l_sfrac >> 16;
				l_sfrac &= 0xFFFF;
				l_tfrac += l_tfracextra;
				if (l_tfrac & 0x10000)
				{
					l_ptex += skinwidth;
					l_tfrac &= 0xFFFF;
				}
				l_light = rc_addw (l_light, l_lightextra);
				l_zi = rc_addw (l_zi, l_ziextra);
				l_errorterm -= l_adjdown;
			}
			else
			{
				lx += l_ubasestep; ly += 1;
				l_aspancount += l_ubasestep;
				l_ptex += l_ptexbase;
				l_sfrac += l_sfracbase;
				l_ptex += l_sfrac >> 16;
				l_sfrac &= 0xFFFF;
				l_tfrac += l_tfracbase;
				if (l_tfrac & 0x10000)
				{
					l_ptex += skinwidth;
					l_tfrac &= 0xFFFF;
				}
				l_light = rc_addw (l_light, l_lightbase);
				l_zi = rc_addw (l_zi, l_zibase);
			}
		}
	}
#undef RC_LEFT_SETUP
}

RC_HD void rc_polyset_triangle (const rc_polyctx_t *c, const int *fv0, int fl0, const int *fv1, int fl1,
				const int *fv2, int fl2, int facesfront, int sub)
{
	rc_aqueue_t q;
	q.mode = c->qmode; q.tri = c->qtri; q.rows = c->qrows; q.span0 = c->qspan0; q.n = 0; q.reserved = c->qmode == 1;
	rc_polyset_triangle_body (c, fv0, fl0, fv1, fl1, fv2, fl2, facesfront, sub, &q);
#ifdef __CUDACC__
	if (q.reserved)
	{
		/* reserved slots the triangle did not use (or could not: queue full) */
		const rc_frame_t *f = c->f;
		long long lim = q.span0 + q.rows;
		if (lim > f->aspancap) lim = f->aspancap;
		for (long long i = q.span0 + (q.mode ? q.n : 0); i < lim; i++)
			f->aspans[i].count = 0;
	}
#endif
}

/* R_Alias_clip_{left,right,top,bottom} (r_aclip.c:96-185): edge 0 = left, 1 = right, 2 = top, 3 = bottom */
RC_HD void rc_alias_clip_xy (const rc_view_t *vw, int which, const rc_finalvert_t *pfv0, const rc_finalvert_t *pfv1, rc_finalvert_t *out)
{
	const int comp = which < 2 ? 0 : 1;
	const int bound = which == 0 ? vw->vrect_x : which == 1 ? vw->vrectright : which == 2 ? vw->vrect_y : vw->vrectbottom;
	float scale;
	if (pfv0->v[1] >= pfv1->v[1])
	{
		scale = (float)(bound - pfv0->v[comp]) / (pfv1->v[comp] - pfv0->v[comp]);
		for (int i = 0; i < 6; i++)
			out->v[i] = rc_d2i (pfv0->v[i] + (pfv1->v[i] - pfv0->v[i])*scale + 0.5);
	}
	else
	{
		scale = (float)(bound - pfv1->v[comp]) / (pfv0->v[comp] - pfv1->v[comp]);
		for (int i = 0; i < 6; i++)
			out->v[i] = rc_d2i (pfv1->v[i] + (pfv0->v[i] - pfv1->v[i])*scale + 0.5);
	}
}

/* R_Alias_clip_z (r_aclip.c:46-91): pfv0 / pfv1 carry their auxvert in .av */
RC_HD void rc_alias_clip_z (const rc_view_t *vw, float ziscale, const rc_finalvert_t *pfv0, const rc_finalvert_t *pfv1, rc_finalvert_t *out)
{
	float scale, avout[3];
	const rc_finalvert_t *a = pfv0, *b = pfv1;
	if (!(pfv0->v[1] >= pfv1->v[1])) { a = pfv1; b = pfv0; }
	scale = (RC_ALIAS_Z_CLIP_PLANE - a->av[2]) / (b->av[2] - a->av[2]);
	avout[0] = a->av[0] + (b->av[0] - a->av[0]) * scale;
	avout[1] = a->av[1] + (b->av[1] - a->av[1]) * scale;
	avout[2] = RC_ALIAS_Z_CLIP_PLANE;
	out->v[2] = rc_f2i (a->v[2] + (b->v[2] - a->v[2]) * scale);
	out->v[3] = rc_f2i (a->v[3] + (b->v[3] - a->v[3]) * scale);
	out->v[4] = rc_f2i (a->v[4] + (b->v[4] - a->v[4]) * scale);
	/* R_AliasProjectFinalVert */
	const float zi = 1.0 / avout[2];
	out->v[5] = rc_f2i (zi * ziscale);
	out->v[0] = rc_f2i ((avout[0] * vw->aliasxscale * zi) + vw->aliasxcenter);
	out->v[1] = rc_f2i ((avout[1] * vw->aliasyscale * zi) + vw->aliasycenter);
	out->av[0] = avout[0]; out->av[1] = avout[1]; out->av[2] = avout[2];
}

/* R_AliasClip, r_aclip.c:190-228.  which: -1 = z plane, 0..3 screen edges */
RC_HD int rc_alias_clip (const rc_view_t *vw, float ziscale, const rc_finalvert_t *in, rc_finalvert_t *out, int flag, int count, int which)
{
	int j = count - 1, k = 0;
	for (int i = 0; i < count; j = i, i++)
	{
		const int oldflags = in[j].flags & flag;
		const int flags = in[i].flags & flag;
		if (flags && oldflags)
			continue;
		if (oldflags ^ flags)
		{
			if (which < 0) rc_alias_clip_z (vw, ziscale, &in[j], &in[i], &out[k]);
			else rc_alias_clip_xy (vw, which, &in[j], &in[i], &out[k]);
			out[k].flags = 0;
			if (out[k].v[0] < vw->vrect_x) out[k].flags |= RC_ALIAS_LEFT_CLIP;
			if (out[k].v[1] < vw->vrect_y) out[k].flags |= RC_ALIAS_TOP_CLIP;
			if (out[k].v[0] > vw->vrectright) out[k].flags |= RC_ALIAS_RIGHT_CLIP;
			if (out[k].v[1] > vw->vrectbottom) out[k].flags |= RC_ALIAS_BOTTOM_CLIP;
			k++;
		}
		if (!flags)
		{
			out[k] = in[i];
			k++;
		}
	}
	return k;
}

/* One thread per (alias instance, triangle): the triangle loops of R_AliasPreparePoints
 * (r_alias.c:340-362) / D_DrawNonSubdiv, including R_AliasClipTriangle (r_aclip.c:235-349). */
/* Triangles whose screen box holds more pixels than this go through the span queue (GPU); smaller ones are drawn by the
 * thread that walks them */
#define RC_QUEUE_MIN_PIXELS 16

/* GPU, before the triangle is walked: the number of span slots an unclipped, front-facing triangle wants from the queue
 * (0: culled, clipped -- those reserve for themselves --, or small enough to be drawn directly) */
RC_HD int rc_alias_triangle_rows (const rc_frame_t *f, int ii, int ti)
{
	const rc_aliasinst_t *in = &f->alias[ii];
	const rc_aliasmodel_t *m = &f->amodels[in->model];
	if (ti >= m->numtris)
		return 0;
	const int *tri = f->atris + 4 * ((size_t)m->tris_ofs + ti);
	const rc_finalvert_t *pfv = f->finalverts + in->vertbase;
	const rc_finalvert_t *a = &pfv[tri[1]], *b = &pfv[tri[2]], *d = &pfv[tri[3]];
	if (!in->trivial_accept && ((a->flags | b->flags | d->flags) & (RC_ALIAS_XY_CLIP_MASK | RC_ALIAS_Z_CLIP)))
		return 0;
	if ((a->v[1] - b->v[1]) * (a->v[0] - d->v[0]) - (a->v[0] - b->v[0]) * (a->v[1] - d->v[1]) >= 0)
		return 0;		/* d_xdenom: faces away */
	const int x0 = rc_min3i (a->v[0], b->v[0], d->v[0]), x1 = rc_max3i (a->v[0], b->v[0], d->v[0]);
	const int y0 = rc_min3i (a->v[1], b->v[1], d->v[1]), y1 = rc_max3i (a->v[1], b->v[1], d->v[1]);
	if ((long long)(x1 - x0 + 1) * (y1 - y0 + 1) <= RC_QUEUE_MIN_PIXELS)
		return 0;
	return y1 - y0;
}

/* qrows > 0 (GPU): the caller reserved that many span slots from qspan0 and the triangle record qtri; qrows = 0 with
 * queue = 1: clipped triangles reserve for themselves, others are drawn directly; queue = 0 (the host simulator, and the
 * reference's own order of work): everything is drawn by this thread */
RC_HD void rc_alias_triangle (const rc_frame_t *f, int ii, int ti, int queue = 0, int qrows = 0, long long qspan0 = 0, int qtri = 0)
{
	const rc_aliasinst_t *in = &f->alias[ii];
	const rc_aliasmodel_t *m = &f->amodels[in->model];
	if (ti >= m->numtris)
		return;
	const rc_view_t *vw = &f->views[in->view - f->viewbase];
	const int *tri = f->atris + 4 * ((size_t)m->tris_ofs + ti);
	const rc_finalvert_t *pfv = f->finalverts + in->vertbase;
	rc_polyctx_t c;
	c.f = f; c.vw = vw;
	c.qmode = qrows > 0 ? 1 : 0; c.qrows = qrows; c.qspan0 = qspan0; c.qtri = qtri;
	c.skinofs = m->skins_ofs + in->skin; c.cmapidx = in->colormap; c.viewidx = in->view - f->viewbase; c.zresidx = vw->zres_index;
	c.skin = f->skinpixels + m->skins_ofs + in->skin;
	c.colormap = f->colormaps + (size_t)in->colormap * 16384;
	c.zplane = f->zres + (size_t)vw->zres_index * f->width * f->height;
	c.zworld = f->zb + (size_t)f->width * f->height * (in->view - f->viewbase); c.zwidth = vw->zwidth;
	c.skinwidth = m->skinwidth; c.skinlo = -in->skin; c.skinhi = m->skins_len - in->skin;
	c.seamfixupX16 = (m->skinwidth >> 1) << 16;
	c.orderbits = (((unsigned long long)in->order << 24) | ((unsigned long long)ti << 8)) << 8;
	const rc_finalvert_t *a = &pfv[tri[1]], *b = &pfv[tri[2]], *d = &pfv[tri[3]];
	const int facesfront = tri[0];
	if (in->trivial_accept)
	{
		rc_polyset_triangle (&c, a->v, a->flags, b->v, b->flags, d->v, d->flags, facesfront, 0);
		return;
	}
	if (a->flags & b->flags & d->flags & (RC_ALIAS_XY_CLIP_MASK | RC_ALIAS_Z_CLIP))
		return;		/* completely clipped */
	if (!((a->flags | b->flags | d->flags) & (RC_ALIAS_XY_CLIP_MASK | RC_ALIAS_Z_CLIP)))
	{
		rc_polyset_triangle (&c, a->v, a->flags, b->v, b->flags, d->v, d->flags, facesfront, 0);
		return;
	}
	/* R_AliasClipTriangle */
	rc_finalvert_t fv[2][8];
	fv[0][0] = *a; fv[0][1] = *b; fv[0][2] = *d;
	if (!facesfront)
		for (int i = 0; i < 3; i++)
			if (fv[0][i].flags & RC_ALIAS_ONSEAM)
				fv[0][i].v[2] += c.seamfixupX16;
	unsigned clipflags = fv[0][0].flags | fv[0][1].flags | fv[0][2].flags;
	int k, pingpong;
	if (clipflags & RC_ALIAS_Z_CLIP)
	{
		k = rc_alias_clip (vw, in->ziscale, fv[0], fv[1], RC_ALIAS_Z_CLIP, 3, -1);
		if (k == 0)
			return;
		pingpong = 1;
		clipflags = fv[1][0].flags | fv[1][1].flags | fv[1][2].flags;	/* first three only (H4) */
	}
	else
	{
		pingpong = 0;
		k = 3;
	}
	const int order[4] = { RC_ALIAS_LEFT_CLIP, RC_ALIAS_RIGHT_CLIP, RC_ALIAS_BOTTOM_CLIP, RC_ALIAS_TOP_CLIP };
	const int which[4] = { 0, 1, 3, 2 };
	for (int s = 0; s < 4; s++)
	{
		if (clipflags & order[s])
		{
			k = rc_alias_clip (vw, in->ziscale, fv[pingpong], fv[pingpong ^ 1], order[s], k, which[s]);
			if (k == 0)
				return;
			pingpong ^= 1;
		}
	}
	for (int i = 0; i < k; i++)
	{
		rc_finalvert_t *p = &fv[pingpong][i];
		if (p->v[0] < vw->vrect_x) p->v[0] = vw->vrect_x;
		else if (p->v[0] > vw->vrectright) p->v[0] = vw->vrectright;
		if (p->v[1] < vw->vrect_y) p->v[1] = vw->vrect_y;
		else if (p->v[1] > vw->vrectbottom) p->v[1] = vw->vrectbottom;
		p->flags = 0;
	}
	c.qmode = queue ? 2 : 0;
	for (int i = 1; i < k - 1; i++)
		rc_polyset_triangle (&c, fv[pingpong][0].v, 0, fv[pingpong][i].v, 0, fv[pingpong][i+1].v, 0, facesfront, i);
}

/* One thread per particle: D_DrawParticle (d_part.c:31-181) with r_pright/r_pup/r_ppn from
 * R_DrawParticles (r_main.c:781-783). */
RC_HD void rc_particle (const rc_frame_t *f, int pi)
{
	const rc_particle_t *p = &f->particles[pi];
	const rc_view_t *vw = &f->views[p->view - f->viewbase];
	float local[3], pright[3], pup[3], t[3];
	for (int i = 0; i < 3; i++)
	{
		local[i] = p->org[i] - vw->origin[i];
		pright[i] = vw->vright[i] * vw->xscaleshrink;
		pup[i] = vw->vup[i] * vw->yscaleshrink;
	}
	t[2] = rc_dot3 (local, vw->vpn);
	if (t[2] < 8.0)
		return;
	t[0] = rc_dot3 (local, pright);
	t[1] = rc_dot3 (local, pup);
	const float zi = 1.0 / t[2];
	const int u = rc_f2i (vw->xcenter + zi * t[0]);
	const int v = rc_f2i (vw->ycenter - zi * t[1]);
	if ((v > vw->d_vrectbottom_particle) || (u > vw->d_vrectright_particle) || (v < vw->d_vrecty) || (u < vw->d_vrectx))
		return;
	const int izi = rc_f2i (zi * 0x8000);
	int pix = izi >> vw->d_pix_shift;
	if (pix < vw->d_pix_min) pix = vw->d_pix_min;
	else if (pix > vw->d_pix_max) pix = vw->d_pix_max;
	const int rows = pix << vw->d_y_aspect_shift;
	unsigned long long *zplane = f->zres + (size_t)vw->zres_index * f->width * f->height;
	const unsigned long long key = ((unsigned long long)(unsigned)(izi + 32768) << 48) |
		((((unsigned long long)1 << 39) | (unsigned long long)p->order) << 8) | (unsigned long long)(p->color & 0xFF);
	if (izi > 32767 || izi < -32768)
		RC_ALIAS_RANGE_HIT (f);
	for (int r = 0; r < rows; r++)
		for (int i = 0; i < pix; i++)
		{
			const int x = u + i, y = v + r;
			/* (no look at the z-buffer or the plane first, as the alias spans do: a particle is a handful of pixels, and
			 * the reads in front of each atomic only made this loop wait -- 0.30 -> 0.54 ms at C3) */
			if ((unsigned)x < (unsigned)f->width && (unsigned)y < (unsigned)f->height)
				RC_ATOMIC_MAX_ULL_BLIND (&zplane[(size_t)y * f->width + x], key);
		}
}

/* ---- sprites (ref_soft/r_sprite.c, d_sprite.c) ---------------------------------------------- */
#define RC_SPRITE_VERTS 12	/* MAXWORKINGVERTS: a quad clipped by four planes has at most 8 */

/* R_ClipSpriteFace, r_sprite.c:57-123: clips the winding in `in` (nump x {x,y,z,s,t}) to one plane */
RC_HD int rc_sprite_clip (const float *normal, float clipdist, float (*in)[5], float (*outv)[5], int nump)
{
	float dists[RC_SPRITE_VERTS + 1];
	int outcount = 0;
	for (int i = 0; i < nump; i++)
		dists[i] = rc_dot3 (in[i], normal) - clipdist;
	dists[nump] = dists[0];
	for (int q = 0; q < 5; q++) in[nump][q] = in[0][q];
	for (int i = 0; i < nump; i++)
	{
		if (dists[i] >= 0)
		{
			for (int q = 0; q < 5; q++) outv[outcount][q] = in[i][q];
			outcount++;
		}
		if (dists[i] == 0 || dists[i+1] == 0)
			continue;
		if ((dists[i] > 0) == (dists[i+1] > 0))
			continue;
		const float frac = dists[i] / (dists[i] - dists[i+1]);
		for (int q = 0; q < 5; q++)
			outv[outcount][q] = in[i][q] + frac*(in[i+1][q] - in[i][q]);
		outcount++;
	}
	return outcount;
}

/* One thread per (drawn sprite, span index k): R_SetupAndDrawSprite (r_sprite.c:131-226) and
 * D_DrawSprite (d_sprite.c:388-441) up to the span list are recomputed by every thread of the sprite
 * (a few hundred operations), then the thread draws the k-th span of the list with D_SpriteDrawSpans
 * (d_sprite.c:36-196): perspective texturing with the 8-pixel subdivision, texel 255 transparent,
 * z test `*pz <= izi >> 16` -- here one atomicMax of (z, draw order, colour) per opaque pixel into
 * the view's resolve plane, like the alias rasteriser. */
RC_HD void rc_sprite_span (const rc_frame_t *f, int si, int k)
{
	const rc_spriteinst_t *sp = &f->sprites[si];
	const rc_view_t *vw = &f->views[sp->view - f->viewbase];
	float cv[2][RC_SPRITE_VERTS + 1][5];
	float pu[RC_SPRITE_VERTS + 1], pv[RC_SPRITE_VERTS + 1];

	/* backface cull */
	if (rc_dot3 (sp->vpn, sp->modelorg) >= 0)
		return;
	/* the poster in world space */
	float right[3], up[3], left[3], down[3];
	for (int i = 0; i < 3; i++)
	{
		right[i] = sp->vright[i] * sp->right; up[i] = sp->vup[i] * sp->up;
		left[i] = sp->vright[i] * sp->left; down[i] = sp->vup[i] * sp->down;
	}
	const float sprite_width = (float)sp->width, sprite_height = (float)sp->height;
	for (int i = 0; i < 3; i++)
	{
		cv[0][0][i] = sp->entorigin[i] + up[i] + left[i];
		cv[0][1][i] = sp->entorigin[i] + up[i] + right[i];
		cv[0][2][i] = sp->entorigin[i] + down[i] + right[i];
		cv[0][3][i] = sp->entorigin[i] + down[i] + left[i];
	}
	cv[0][0][3] = 0; cv[0][0][4] = 0;
	cv[0][1][3] = sprite_width; cv[0][1][4] = 0;
	cv[0][2][3] = sprite_width; cv[0][2][4] = sprite_height;
	cv[0][3][3] = 0; cv[0][3][4] = sprite_height;
	/* clip to the frustum in world space */
	int nump = 4, cur = 0;
	for (int i = 0; i < 4; i++)
	{
		nump = rc_sprite_clip (vw->clipnormal[i], vw->clipdist[i], cv[cur], cv[cur ^ 1], nump);
		cur ^= 1;
		if (nump < 3)
			return;
		if (nump >= RC_SPRITE_VERTS)
			return;		/* the reference: Sys_Error ("too many points") */
	}
	/* transform into view space and project */
	for (int i = 0; i < nump; i++)
	{
		float local[3], t[3];
		local[0] = cv[cur][i][0] - vw->origin[0]; local[1] = cv[cur][i][1] - vw->origin[1]; local[2] = cv[cur][i][2] - vw->origin[2];
		rc_transform_vector (vw, local, t);
		if (t[2] < RC_NEAR_CLIP)
			t[2] = RC_NEAR_CLIP;
		const float zi = 1.0 / t[2];
		float scale = vw->xscale * zi;
		pu[i] = (vw->xcenter + scale * t[0]);
		scale = vw->yscale * zi;
		pv[i] = (vw->ycenter - scale * t[1]);
	}
	/* D_DrawSprite: top and bottom vertices */
	float ymin = 999999.9, ymax = -999999.9;
	int minindex = 0, maxindex = 0;
	for (int i = 0; i < nump; i++)
	{
		if (pv[i] < ymin) { ymin = pv[i]; minindex = i; }
		if (pv[i] > ymax) { ymax = pv[i]; maxindex = i; }
	}
	ymin = ceil ((double)ymin);
	ymax = ceil ((double)ymax);
	if (ymin >= ymax)
		return;
	pu[nump] = pu[0]; pv[nump] = pv[0];

	/* D_SpriteScanLeftEdge, d_sprite.c:204-258: the k-th entry of the span list */
	int span_u = 0, span_v = 0, haveleft = 0;
	{
		int i = minindex, lmaxindex = maxindex, n = 0;
		if (i == 0) i = nump;
		if (lmaxindex == 0) lmaxindex = nump;
		float vtop = ceil ((double)pv[i]);
		do
		{
			const float vbottom = ceil ((double)pv[i - 1]);
			if (vtop < vbottom)
			{
				const float du = pu[i - 1] - pu[i], dv = pv[i - 1] - pv[i];
				const float slope = du / dv;
				const int u_step = rc_f2i (slope * 0x10000);
				const int u = rc_addw (rc_f2i ((pu[i] + (slope * (vtop - pv[i]))) * 0x10000), (0x10000 - 1));
				const int itop = rc_f2i (vtop), ibottom = rc_f2i (vbottom);
				if (!haveleft && k < n + (ibottom - itop))
				{
					const int j = k - n;
					span_u = rc_addw (u, rc_mulw (j, u_step)) >> 16;
					span_v = itop + j;
					haveleft = 1;
				}
				n += ibottom - itop;
			}
			vtop = vbottom;
			i--;
			if (i == 0) i = nump;
		} while (i != lmaxindex);
	}
	/* D_SpriteScanRightEdge, d_sprite.c:266-338 */
	int span_count = 0, haveright = 0;
	{
		int i = minindex, n = 0;
		float vvert = pv[i];
		if (vvert < vw->fvrecty_adj) vvert = vw->fvrecty_adj;
		if (vvert > vw->fvrectbottom_adj) vvert = vw->fvrectbottom_adj;
		float vtop = ceil ((double)vvert);
		do
		{
			float vnext = pv[i + 1];
			if (vnext < vw->fvrecty_adj) vnext = vw->fvrecty_adj;
			if (vnext > vw->fvrectbottom_adj) vnext = vw->fvrectbottom_adj;
			const float vbottom = ceil ((double)vnext);
			if (vtop < vbottom)
			{
				float uvert = pu[i], unext = pu[i + 1];
				if (uvert < vw->fvrectx_adj) uvert = vw->fvrectx_adj;
				if (uvert > vw->fvrectright_adj) uvert = vw->fvrectright_adj;
				if (unext < vw->fvrectx_adj) unext = vw->fvrectx_adj;
				if (unext > vw->fvrectright_adj) unext = vw->fvrectright_adj;
				const float du = unext - uvert, dv = vnext - vvert;
				const float slope = du / dv;
				const int u_step = rc_f2i (slope * 0x10000);
				const int u = rc_addw (rc_f2i ((uvert + (slope * (vtop - vvert))) * 0x10000), (0x10000 - 1));
				const int itop = rc_f2i (vtop), ibottom = rc_f2i (vbottom);
				if (!haveright && k < n + (ibottom - itop))
				{
					span_count = (rc_addw (u, rc_mulw (k - n, u_step)) >> 16) - span_u;
					haveright = 1;
				}
				n += ibottom - itop;
			}
			vtop = vbottom;
			vvert = vnext;
			i++;
			if (i == nump) i = 0;
		} while (i != maxindex);
	}
	if (!haveright)
		return;				/* past DS_SPAN_LIST_END */
	if (!haveleft)
	{
		/* the reference would read a span the left scan never wrote (stack garbage) */
		RC_ALIAS_RANGE_HIT (f);
		return;
	}
	if (span_count <= 0)
		return;

	/* D_SpriteCalculateGradients, d_sprite.c:346-381 */
	float p_normal[3], p_saxis[3], p_taxis[3], p_temp1[3];
	rc_transform_vector (vw, sp->vpn, p_normal);
	rc_transform_vector (vw, sp->vright, p_saxis);
	rc_transform_vector (vw, sp->vup, p_taxis);
	p_taxis[0] = -p_taxis[0]; p_taxis[1] = -p_taxis[1]; p_taxis[2] = -p_taxis[2];
	const float distinv = 1.0 / (-rc_dot3 (sp->modelorg, sp->vpn));
	const float d_sdivzstepu = p_saxis[0] * vw->xscaleinv, d_tdivzstepu = p_taxis[0] * vw->xscaleinv;
	const float d_sdivzstepv = -p_saxis[1] * vw->yscaleinv, d_tdivzstepv = -p_taxis[1] * vw->yscaleinv;
	const float d_zistepu = p_normal[0] * vw->xscaleinv * distinv, d_zistepv = -p_normal[1] * vw->yscaleinv * distinv;
	const float d_sdivzorigin = p_saxis[2] - vw->xcenter * d_sdivzstepu - vw->ycenter * d_sdivzstepv;
	const float d_tdivzorigin = p_taxis[2] - vw->xcenter * d_tdivzstepu - vw->ycenter * d_tdivzstepv;
	const float d_ziorigin = p_normal[2] * distinv - vw->xcenter * d_zistepu - vw->ycenter * d_zistepv;
	rc_transform_vector (vw, sp->modelorg, p_temp1);
	const int cachewidth = sp->width, sheight = sp->height;
	const int sadjust = rc_d2i (rc_dot3 (p_temp1, p_saxis) * 0x10000 + 0.5) - (-(cachewidth >> 1) << 16);
	const int tadjust = rc_d2i (rc_dot3 (p_temp1, p_taxis) * 0x10000 + 0.5) - (-(sheight >> 1) << 16);
	const int bbextents = (cachewidth << 16) - 1, bbextentt = (sheight << 16) - 1;

	/* D_SpriteDrawSpans, d_sprite.c:36-196, one span */
	if ((unsigned)span_v >= (unsigned)f->height)
		return;
	const uint8_t *pbase = f->spritepixels + sp->pixofs;
	unsigned long long *zrow = f->zres + (size_t)vw->zres_index * f->width * f->height + (size_t)span_v * f->width;
	const unsigned long long orderbits = ((unsigned long long)sp->order << 24) << 8;
	const float sdivz8stepu = d_sdivzstepu * 8, tdivz8stepu = d_tdivzstepu * 8, zi8stepu = d_zistepu * 8;
	const int izistep = rc_f2i (d_zistepu * 0x8000 * 0x10000);
	int count = span_count, x = span_u;
	const float du = (float)span_u, dv = (float)span_v;
	float sdivz = d_sdivzorigin + dv*d_sdivzstepv + du*d_sdivzstepu;
	float tdivz = d_tdivzorigin + dv*d_tdivzstepv + du*d_tdivzstepu;
	float zi = d_ziorigin + dv*d_zistepv + du*d_zistepu;
	float z = (float)0x10000 / zi;
	int izi = rc_f2i (zi * 0x8000 * 0x10000);
	int s = rc_addw (rc_f2i (sdivz * z), sadjust);
	if (s > bbextents) s = bbextents; else if (s < 0) s = 0;
	int t = rc_addw (rc_f2i (tdivz * z), tadjust);
	if (t > bbextentt) t = bbextentt; else if (t < 0) t = 0;
	int sstep = 0, tstep = 0;
	do
	{
		int spancount = count >= 8 ? 8 : count, snext, tnext;
		count -= spancount;
		if (count)
		{
			sdivz += sdivz8stepu; tdivz += tdivz8stepu; zi += zi8stepu;
			z = (float)0x10000 / zi;
			snext = rc_addw (rc_f2i (sdivz * z), sadjust);
			if (snext > bbextents) snext = bbextents; else if (snext < 8) snext = 8;
			tnext = rc_addw (rc_f2i (tdivz * z), tadjust);
			if (tnext > bbextentt) tnext = bbextentt; else if (tnext < 8) tnext = 8;
			sstep = (snext - s) >> 3;
			tstep = (tnext - t) >> 3;
		}
		else
		{
			const float spancountminus1 = (float)(spancount - 1);
			sdivz += d_sdivzstepu * spancountminus1; tdivz += d_tdivzstepu * spancountminus1; zi += d_zistepu * spancountminus1;
			z = (float)0x10000 / zi;
			snext = rc_addw (rc_f2i (sdivz * z), sadjust);
			if (snext > bbextents) snext = bbextents; else if (snext < 8) snext = 8;
			tnext = rc_addw (rc_f2i (tdivz * z), tadjust);
			if (tnext > bbextentt) tnext = bbextentt; else if (tnext < 8) tnext = 8;
			if (spancount > 1)
			{
				sstep = (snext - s) / (spancount - 1);
				tstep = (tnext - t) / (spancount - 1);
			}
		}
		do
		{
			const uint8_t btemp = pbase[(s >> 16) + (t >> 16) * cachewidth];
			if (btemp != 255 && (unsigned)x < (unsigned)f->width)
			{
				const int z16 = izi >> 16;
				const unsigned long long key = ((unsigned long long)(unsigned)(z16 + 32768) << 48) | orderbits | btemp;
				RC_ATOMIC_MAX_ULL (&zrow[x], key);
			}
			izi = rc_addw (izi, izistep);
			x++;
			s = rc_addw (s, sstep);
			t = rc_addw (t, tstep);
		} while (--spancount > 0);
		s = snext;
		t = tnext;
	} while (count > 0);
}

/* One thread per pixel of a view that drew alias models or particles: applies the winning
 * fragment if it passes the reference's test against the world z (d_polyse.c:422, d_part.c:82). */
RC_HD void rc_zresolve_pixel (const rc_frame_t *f, int vi, int x, int y)
{
	const rc_view_t *vw = &f->views[vi];
	if (vw->zres_index < 0 || x >= f->width || y >= f->height)
		return;
	const size_t viewpix = (size_t)f->width * f->height;
	const unsigned long long key = f->zres[(size_t)vw->zres_index * viewpix + (size_t)y * f->width + x];
	if (!key)
		return;
	const int zf = (int)(key >> 48) - 32768;
	int16_t *pz = f->zb + viewpix * vi + (size_t)vw->zwidth * y + x;
	if (zf >= *pz)
	{
		*pz = (int16_t)zf;
		uint8_t *base = vw->warp_index < 0 ? f->fb + viewpix * vi : f->warpbuf + (size_t)vw->warp_index * (RC_WARP_W * RC_WARP_H);
		base[(size_t)vw->screenwidth * y + x] = (uint8_t)(key & 0xFF);
	}
}

#endif /* RC_PIPELINE_H */
