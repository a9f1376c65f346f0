#define _GNU_SOURCE
/*
 * rc_api.c -- the C-ABI of libref_cuda.so (include/ref_cuda.h): the reference's
 * client/render.h symbol set, the batched RC_* extension and the refexport_t table.
 * Host-side only; all rendering is done by the CUDA pipeline behind rc_gpu_*.
 * There is no CPU rendering path: without a CUDA device RC_Init fails.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <stdarg.h>
#include <math.h>
#include <time.h>
#include <dlfcn.h>
#include "rc_host.h"
#include "rc_draw2d.h"

#define RC_MAX_MODELS 512
#define RC_MAX_PICS 256

static struct {
	int          inited;
	rc_config_t  cfg;
	char         basedir[512];
	rc_vidinfo_t vid;
	rc_cvars_t   cvars;
	rc_gpu_t    *gpu;
	uint8_t      palette[768];
	uint8_t      colormap[256 * 64 + 4];
	struct model_s *models[RC_MAX_MODELS];
	int          nummodels;
	struct model_s *world;
	char         err[1024];
	int          loadlits;		/* r_loadlits (r_main.c:130) */
	int          host_threads;	/* threads that set up the views of a large batch */
	int         *entofs; int entofscap;	/* per view: scratch slices of the entity set-up */
	/* 2D overlay (r_draw.c): conchars + pictures + translation tables in one blob, mirrored on the device */
	uint8_t     *d2blob; size_t d2len, d2cap;
	rc_pic_t     d2pics[RC_MAX_PICS]; char d2names[RC_MAX_PICS][64]; int d2npics;
	struct mpic_s *d2mpics[RC_MAX_PICS];	/* mpic_t views of the pictures for the render.h functions */
	int          d2dirty, d2ready;
	char         d2version[100];
	uint8_t     *wad; size_t wadlen;
	rc_draw2d_t *d2frame; int d2framen, d2framecap;	/* commands recorded by the render.h Draw_* functions since R_BeginFrame */
	float        gamma; uint8_t gammatable[256]; int gamma_valid;	/* v_gamma + R_BuildGammaTable */
	uint32_t    *palscratch; int palscratchcap;
	/* single-view path targets (RC_SetVidBuffers) */
	uint8_t     *vidbuffer; int vidrowbytes; int16_t *zbuffer;
	rc_view_t   *viewscratch; int viewscratchcap;
	rc_dlight_t *dlscratch; int dlscratchcap;
	rc_aliasinst_t *aliasscratch; int aliasscratchcap;
	rc_particle_t *partscratch; int partscratchcap;
	rc_bent_t   *bentscratch; int bentscratchcap;
	uint8_t     *cmapscratch; int cmapscratchcap;
	float        anorm_dots[16 * 256];
	int          alias_dirty;		/* alias models loaded since the last device upload */
	int          ent_parts;			/* entity-heavy device batches are drawn in this many parts (RC_ENT_PARTS, experiment) */
	/* device residency of the alias models (the reference's Cache_Alloc lifetimes, r_model.c:95-130 / zone.c) */
	size_t       modelcache_bytes;		/* budget, 0: every loaded model stays resident */
	long long    model_clock;
	rc_modelcache_t modelcache;
	int          sprite_dirty;		/* ditto sprite models */
	rc_spriteinst_t *spritescratch; int spritescratchcap;
	rc_batch_t   batch;
} rc;

static int rc_fail (int code, const char *fmt, ...)
{
	va_list ap;
	va_start (ap, fmt);
	vsnprintf (rc.err, sizeof(rc.err), fmt, ap);
	va_end (ap);
	return code;
}

const char *RC_LastError (void) { return rc.err; }

void RC_DefaultConfig (rc_config_t *c)
{
	memset (c, 0, sizeof(*c));
	c->width = 320; c->height = 200; c->device = 0; c->basedir = ".";
	/* reference defaults: r_main.c:108-134, d_init.c:26-28, R_Init :223-224 */
	c->r_clearcolor = 2; c->r_fullbright = 0; c->r_ambient = 0; c->r_drawentities = 1;
	c->r_drawviewmodel = 1; c->r_waterwarp = 1; c->r_fastsky = 0; c->r_skycolor = 4;
	c->r_lerpmodels = 1; c->d_mipcap = 0; c->d_mipscale = 1;
	c->r_maxsurfs = 1000; c->r_maxedges = 2000;
}

void RC_GetConfig (rc_config_t *c)
{
	if (!c) return;
	memset (c, 0, sizeof(*c));
	if (!rc.inited) return;
	*c = rc.cfg;
	c->basedir = rc.basedir;
	c->width = rc.vid.width; c->height = rc.vid.height;
	c->host_threads = rc.host_threads;
}

static void *rc_read_file (const char *name, size_t *len)
{
	char path[1024];
	FILE *f;
	long n;
	void *buf;
	snprintf (path, sizeof(path), "%s/id1/%s", rc.basedir, name);
	f = fopen (path, "rb");
	if (!f) return NULL;
	fseek (f, 0, SEEK_END); n = ftell (f); fseek (f, 0, SEEK_SET);
	buf = malloc (n > 0 ? n : 1);
	if (buf && fread (buf, 1, n, f) != (size_t)n) { free (buf); buf = NULL; }
	fclose (f);
	if (len) *len = n;
	return buf;
}

int RC_Init (const rc_config_t *cfg)
{
	size_t n;
	void *p;
	if (rc.inited) return rc_fail (RC_ERR_ARG, "RC_Init called twice");
	if (!cfg || cfg->width <= 0 || cfg->height <= 0 || cfg->width > RC_MAXWIDTH || cfg->height > RC_MAXHEIGHT)
		return rc_fail (RC_ERR_ARG, "bad config (size limit %dx%d)", RC_MAXWIDTH, RC_MAXHEIGHT);
	rc.cfg = *cfg;
	snprintf (rc.basedir, sizeof(rc.basedir), "%s", cfg->basedir ? cfg->basedir : ".");
	rc.cfg.basedir = rc.basedir;
	/* common/host.c:827-832 */
	p = rc_read_file ("gfx/palette.lmp", &n);
	if (!p || n < 768) { free (p); return rc_fail (RC_ERR_FILE, "Couldn't load gfx/palette.lmp"); }
	memcpy (rc.palette, p, 768); free (p);
	p = rc_read_file ("gfx/colormap.lmp", &n);
	if (!p || n < 256 * 64) { free (p); return rc_fail (RC_ERR_FILE, "Couldn't load gfx/colormap.lmp"); }
	memcpy (rc.colormap, p, 256 * 64); free (p);
	/* VID_Init equivalent: win32/vid_win.c:1351 */
	rc.vid.width = cfg->width; rc.vid.height = cfg->height; rc.vid.rowbytes = cfg->width;
	rc.vid.aspect = ((float)cfg->height / (float)cfg->width) * (320.0 / 240.0);
	rc.cvars.r_clearcolor = cfg->r_clearcolor; rc.cvars.r_fullbright = cfg->r_fullbright;
	rc.cvars.r_ambient = cfg->r_ambient; rc.cvars.r_drawentities = cfg->r_drawentities;
	rc.cvars.r_drawviewmodel = cfg->r_drawviewmodel; rc.cvars.r_waterwarp = cfg->r_waterwarp;
	rc.cvars.r_fastsky = cfg->r_fastsky; rc.cvars.r_skycolor = cfg->r_skycolor;
	rc.cvars.r_lerpmodels = cfg->r_lerpmodels; rc.cvars.d_mipcap = cfg->d_mipcap;
	rc.cvars.d_mipscale = cfg->d_mipscale;
	rc.cvars.r_draworder = cfg->r_draworder;
	rc.cvars.r_maxsurfs = cfg->r_maxsurfs < 1000 ? 1000 : cfg->r_maxsurfs;	/* MINSURFACES, r_main.c:267 */
	rc.cvars.r_maxedges = cfg->r_maxedges < 2000 ? 2000 : cfg->r_maxedges;	/* MINEDGES, r_main.c:293 */
	{
		/* r_avertexnormal_dots (r_alias.c:29-31): data/anorm_dots.f32 next to the library */
		Dl_info di;
		char path[1024];
		FILE *f = NULL;
		if (dladdr ((void *)RC_Init, &di) && di.dli_fname)
		{
			char *slash;
			snprintf (path, sizeof(path) - 32, "%s", di.dli_fname);
			slash = strrchr (path, '/');
			if (slash) *slash = 0; else strcpy (path, ".");
			strcat (path, "/data/anorm_dots.f32");
			f = fopen (path, "rb");
			if (!f)
			{
				/* the host-sim build lives in tests/hostsim/_build */
				snprintf (path, sizeof(path) - 64, "%s", di.dli_fname);
				slash = strrchr (path, '/');
				if (slash) *slash = 0;
				strcat (path, "/../../../tochris_b200/data/anorm_dots.f32");
				f = fopen (path, "rb");
			}
		}
		if (!f || fread (rc.anorm_dots, sizeof(float), 16 * 256, f) != 16 * 256)
		{
			if (f) fclose (f);
			return rc_fail (RC_ERR_FILE, "Couldn't load data/anorm_dots.f32 (alias shading table)");
		}
		fclose (f);
	}
	{
		long ncpu = sysconf (_SC_NPROCESSORS_ONLN);
		{ const char *e = getenv ("RC_ENT_PARTS"); rc.ent_parts = e ? atoi (e) : 0; }
		rc.host_threads = cfg->host_threads > 0 ? cfg->host_threads : (int)(ncpu > 8 ? 8 : (ncpu < 1 ? 1 : ncpu));
		if (rc.host_threads > 64) rc.host_threads = 64;
	}
	rc.gpu = rc_gpu_create (cfg->device, cfg->width, cfg->height, cfg->surfcache_bytes, cfg->max_views_in_flight,
				rc.err, sizeof(rc.err));
	if (!rc.gpu) return RC_ERR_CUDA;
	rc.inited = 1;
	return RC_OK;
}

void RC_Shutdown (void)
{
	Mod_ClearAll ();
	if (rc.gpu) rc_gpu_destroy (rc.gpu);
	free (rc.viewscratch);
	free (rc.dlscratch); free (rc.aliasscratch); free (rc.partscratch); free (rc.cmapscratch); free (rc.bentscratch); free (rc.spritescratch); free (rc.entofs);
	memset (&rc, 0, sizeof(rc));
}

/* ---- models ---------------------------------------------------------------------------- */
void Mod_Init (void) {}

void Mod_ClearAll (void)
{
	int i;
	for (i = 0; i < rc.nummodels; i++)
	{
		struct model_s *m = rc.models[i];
		if (m->type == rc_mod_brush && m->submodel == 0 && m->world)
			RC_FreeHostWorld (m->world);
		RC_FreeHostAlias (m->alias);
		RC_FreeHostSprite (m->sprite);
		free (m);
	}
	rc.nummodels = 0;
	rc.world = NULL;
	rc.alias_dirty = 1;
	rc.sprite_dirty = 1;
	if (rc.gpu) rc_gpu_flush_caches (rc.gpu);	/* D_FlushCaches, r_model.c:178 */
}

static struct model_s *rc_find_model (const char *name)
{
	int i;
	for (i = 0; i < rc.nummodels; i++)
		if (!strcmp (rc.models[i]->name, name))
			return rc.models[i];
	return NULL;
}

static struct model_s *rc_new_model (const char *name)
{
	struct model_s *m;
	if (rc.nummodels >= RC_MAX_MODELS) return NULL;
	m = (struct model_s *)calloc (1, sizeof(*m));
	snprintf (m->name, sizeof(m->name), "%s", name);
	rc.models[rc.nummodels++] = m;
	return m;
}

static struct model_s *rc_load_model (const char *name)
{
	struct model_s *m = rc_find_model (name);
	size_t len;
	void *buf;
	int i;
	if (m) return m;
	buf = rc_read_file (name, &len);
	if (!buf) { rc_fail (RC_ERR_FILE, "Mod_NumForName: %s not found", name); return NULL; }
	if (len >= 4 && *(int *)buf == 29)
	{
		rc_hostworld_t *hw = RC_LoadBrushModel (buf, len, rc.colormap, rc.err, sizeof(rc.err));
		free (buf);
		if (!hw) return NULL;
		if (rc.loadlits)
		{
			/* r_model.c:533-536: the model's name with the extension replaced by ".lit" */
			char litname[600];
			char *dot;
			size_t litlen;
			void *lit;
			snprintf (litname, sizeof(litname) - 5, "%s", name);
			dot = strrchr (litname, '.');
			if (dot && !strchr (dot, '/')) *dot = 0;
			strcat (litname, ".lit");
			lit = rc_read_file (litname, &litlen);
			if (lit) RC_ApplyLit (hw, (const uint8_t *)lit, litlen);
			free (lit);
		}
		/* r_model.c:1098-1126: model 0 under its file name, submodels as "*1", "*2", ... */
		for (i = 0; i < hw->numsubmodels; i++)
		{
			char sub[16];
			struct model_s *sm;
			const rc_submodel_t *bm = &hw->submodels[i];
			if (i == 0) sm = rc_new_model (name);
			else { snprintf (sub, sizeof(sub), "*%i", i); sm = rc_find_model (sub); if (!sm) sm = rc_new_model (sub); }
			if (!sm) { rc_fail (RC_ERR_NOMEM, "too many models"); return NULL; }
			sm->type = rc_mod_brush; sm->world = hw; sm->submodel = i; sm->numframes = 2; sm->flags = 0;
			memcpy (sm->mins, bm->mins, sizeof(sm->mins)); memcpy (sm->maxs, bm->maxs, sizeof(sm->maxs));
			sm->radius = bm->radius;
			if (i == 0) m = sm;
		}
		return m;
	}
	if (len >= 4 && !memcmp (buf, "IDPO", 4))
	{
		rc_hostalias_t *a = RC_LoadAliasModel (buf, len, name, rc.err, sizeof(rc.err));
		free (buf);
		if (!a) return NULL;
		m = rc_new_model (name);
		if (!m) { RC_FreeHostAlias (a); rc_fail (RC_ERR_NOMEM, "too many models"); return NULL; }
		m->type = rc_mod_alias; m->alias = a; m->flags = a->flags; m->numframes = a->numframes; m->synctype = a->synctype;
		memcpy (m->mins, a->mins, sizeof(m->mins)); memcpy (m->maxs, a->maxs, sizeof(m->maxs));
		rc.alias_dirty = 1;
		return m;
	}
	if (len >= 4 && !memcmp (buf, "IDSP", 4))
	{
		/* Mod_LoadSpriteModel, r_model.c:1671-1741 */
		rc_hostsprite_t *sp = RC_LoadSpriteModel (buf, len, rc.err, sizeof(rc.err));
		free (buf);
		if (!sp) return NULL;
		m = rc_new_model (name);
		if (!m) { RC_FreeHostSprite (sp); rc_fail (RC_ERR_NOMEM, "too many models"); return NULL; }
		m->type = rc_mod_sprite; m->sprite = sp; m->flags = 0; m->numframes = sp->numframes; m->synctype = sp->synctype;
		m->mins[0] = m->mins[1] = -sp->maxwidth/2; m->maxs[0] = m->maxs[1] = sp->maxwidth/2;
		m->mins[2] = -sp->maxheight/2; m->maxs[2] = sp->maxheight/2;
		rc.sprite_dirty = 1;
		return m;
	}
	free (buf);
	rc_fail (RC_ERR_FORMAT, "Mod_NumForName: %s has an unsupported format", name);
	return NULL;
}

struct model_s *Mod_ForName (char *name, qboolean crash)
{
	struct model_s *m = rc_load_model (name);
	if (!m && crash)
	{
		/* the reference calls Sys_Error here (r_model.c:296); a library must not exit() */
		fprintf (stderr, "ref_cuda: %s\n", rc.err);
	}
	return m;
}

/* Mod_TouchModel, r_model.c:247-258: a loaded alias model becomes the most recently used one (Cache_Check), so that it is
 * the last to leave the device when the model cache is short of room; never loads anything */
void Mod_TouchModel (char *name)
{
	struct model_s *m = name ? rc_find_model (name) : NULL;
	if (m && m->type == rc_mod_alias && m->alias)
		m->alias->last_used = ++rc.model_clock;
}

static size_t rc_alias_device_bytes (const rc_hostalias_t *a)
{
	return 16 * (size_t)a->numverts + 16 * (size_t)a->numtris + 4 * (size_t)a->numposes * a->numverts + (size_t)a->skinimagelen;
}

/* The device copy of the alias models has the lifetimes of the reference's model cache: the reference keeps alias models
 * in the Cache_Alloc heap, which throws out the least recently used ones when it is short of room (zone.c), reloads a
 * model when something asks for it again (Mod_Extradata -> Mod_LoadModel, r_model.c:95-130) and lets Mod_TouchModel
 * refresh a model's place in that order.  Here the budget is RC_SetModelCacheBytes: a batch's models are always
 * resident while it is drawn; when one of them is not, the resident set is rebuilt -- the batch's models first, then
 * the others from the most to the least recently used while they fit -- and uploaded from the host copy. */
static int rc_alias_make_resident (const refdef_t *views, int n)
{
	int i, j, need = rc.alias_dirty, nloaded = 0, nres = 0;
	rc_hostalias_t *loaded[RC_MAX_MODELS], *res[RC_MAX_MODELS];
	size_t bytes = 0;
	const long long now = ++rc.model_clock;
	for (i = 0; i < n; i++)
		for (j = 0; j < views[i].num_entities; j++)
		{
			const struct model_s *mod = views[i].entities[j].model;
			if (mod && mod->type == rc_mod_alias && mod->alias && mod->alias->last_used != now)
			{
				mod->alias->last_used = now;
				if (mod->alias->device_index < 0) need = 1;
			}
		}
	if (!need)
		return RC_OK;
	for (i = 0; i < rc.nummodels; i++)
		if (rc.models[i]->type == rc_mod_alias && rc.models[i]->alias)
			loaded[nloaded++] = rc.models[i]->alias;
	/* most recently used first (insertion sort: a few dozen models; stable, so equal stamps keep load order) */
	for (i = 1; i < nloaded; i++)
	{
		rc_hostalias_t *a = loaded[i];
		for (j = i; j > 0 && loaded[j - 1]->last_used < a->last_used; j--) loaded[j] = loaded[j - 1];
		loaded[j] = a;
	}
	for (i = 0; i < nloaded; i++)
	{
		rc_hostalias_t *a = loaded[i];
		const size_t b = rc_alias_device_bytes (a);
		const int wanted = a->last_used == now;
		if (wanted || !rc.modelcache_bytes || bytes + b <= rc.modelcache_bytes)
		{
			if (a->device_index < 0) rc.modelcache.uploads++;
			res[nres++] = a;
			bytes += b;
		}
		else
		{
			if (a->device_index >= 0) rc.modelcache.evictions++;
			a->device_index = -1;
		}
	}
	if (rc_gpu_upload_alias (rc.gpu, res, nres, rc.anorm_dots, rc.err, sizeof(rc.err)) != RC_OK)
		return RC_ERR_CUDA;
	rc.modelcache.rebuilds++;
	rc.modelcache.resident = nres; rc.modelcache.loaded = nloaded;
	rc.modelcache.resident_bytes = bytes; rc.modelcache.budget_bytes = rc.modelcache_bytes;
	rc.alias_dirty = 0;
	return RC_OK;
}

void RC_SetModelCacheBytes (size_t bytes)
{
	if (bytes != rc.modelcache_bytes) rc.alias_dirty = 1;	/* the next batch with entities re-packs the resident set */
	rc.modelcache_bytes = bytes;
}

void RC_ModelCacheInfo (rc_modelcache_t *out)
{
	if (out) { *out = rc.modelcache; out->budget_bytes = rc.modelcache_bytes; }
}
int  Mod_Flags (struct model_s *model) { return model ? model->flags : 0; }
void R_TranslatePlayerSkin (int playernum, struct model_s *model, int skinnum, int top, int bottom)
{ (void)playernum; (void)model; (void)skinnum; (void)top; (void)bottom; }

void R_Init (void) {}

void R_NewMap (struct model_s *worldmodel)
{
	if (!worldmodel || worldmodel->type != rc_mod_brush || !rc.gpu) return;
	rc.world = worldmodel;
	if (rc_gpu_upload_world (rc.gpu, worldmodel->world, (int)rc.cvars.r_maxedges, (int)rc.cvars.r_maxsurfs,
				 rc.err, sizeof(rc.err)) != RC_OK)
		rc.world = NULL;
}

int RC_LoadWorld (const char *name)
{
	struct model_s *m;
	if (!rc.inited) return rc_fail (RC_ERR_ARG, "RC_Init first");
	Mod_ClearAll ();
	m = rc_load_model (name);
	if (!m) return RC_ERR_FILE;
	if (m->type != rc_mod_brush) return rc_fail (RC_ERR_FORMAT, "%s is not a brush model", name);
	R_NewMap (m);
	return rc.world ? RC_OK : RC_ERR_CUDA;
}

/* ---- rendering --------------------------------------------------------------------------- */
static int rc_prepare_views (const refdef_t *views, int n)
{
	int i, j, ndl, nwarp = 0, nbent_total = 0, nbents = 0, maxbfaces = 0;
	if (!rc.inited) return rc_fail (RC_ERR_ARG, "RC_Init first");
	if (!views || n <= 0) return rc_fail (RC_ERR_ARG, "no views");
	if (n > rc.viewscratchcap)
	{
		free (rc.viewscratch);
		rc.viewscratch = (rc_view_t *)malloc (sizeof(rc_view_t) * n);
		rc.viewscratchcap = rc.viewscratch ? n : 0;
		if (!rc.viewscratch) return rc_fail (RC_ERR_NOMEM, "out of memory");
	}
	ndl = 0;
	for (i = 0; i < n; i++)
	{
		int nd = views[i].num_dlights > 0 ? views[i].num_dlights : 0, nb = 0;
		if (nd > MAX_DLIGHTS) return rc_fail (RC_ERR_ARG, "view %d: more than %d dlights", i, MAX_DLIGHTS);
		for (j = 0; j < views[i].num_entities; j++)
			if (views[i].entities[j].model && views[i].entities[j].model->type == rc_mod_brush) nb++;
		ndl += nd * (1 + nb);
		nbent_total += nb;
	}
	if (nbent_total > rc.bentscratchcap)
	{
		free (rc.bentscratch);
		rc.bentscratch = (rc_bent_t *)malloc (sizeof(rc_bent_t) * nbent_total);
		rc.bentscratchcap = rc.bentscratch ? nbent_total : 0;
		if (!rc.bentscratch) return rc_fail (RC_ERR_NOMEM, "out of memory");
	}
	if (ndl > rc.dlscratchcap)
	{
		free (rc.dlscratch);
		rc.dlscratch = (rc_dlight_t *)malloc (sizeof(rc_dlight_t) * ndl);
		rc.dlscratchcap = rc.dlscratch ? ndl : 0;
		if (!rc.dlscratch) return rc_fail (RC_ERR_NOMEM, "out of memory");
	}
	/* per-view constants (R_SetupFrame / R_ViewChanged: libm, the view leaf's BSP descent): views are independent, so
	 * large batches (BASELINE config C4: 65,536 views per call) are set up by all the host threads -- same arithmetic per
	 * view, any order */
	{
		int bad = -1, badcode = RC_OK;
		#pragma omp parallel for schedule(static) num_threads(rc.host_threads) if (n >= 512 && rc.host_threads > 1)
		for (i = 0; i < n; i++)
		{
			rc_view_t *v = &rc.viewscratch[i];
			int code = RC_SetupView (&views[i], &rc.vid, &rc.cvars, v);
			if (code != RC_OK)
			{
				#pragma omp critical
				{ if (bad < 0 || i < bad) { bad = i; badcode = code; } }
				continue;
			}
			v->viewleaf = (rc.world && !(views[i].flags & RDF_NOWORLDMODEL)) ? RC_PointInLeaf (rc.world->world, views[i].vieworg) : -1;
		}
		if (bad >= 0) return rc_fail (badcode, "view %d: bad refdef rectangle", bad);
	}
	ndl = 0;
	for (i = 0; i < n; i++)
	{
		rc_view_t *v = &rc.viewscratch[i];
		if (!rc.world && !(views[i].flags & RDF_NOWORLDMODEL))
			return rc_fail (RC_ERR_NOWORLD, "R_RenderView: NULL worldmodel");	/* r_main.c:1003 */
		if (views[i].num_dlights > MAX_DLIGHTS)
			return rc_fail (RC_ERR_ARG, "view %d: more than %d dlights", i, MAX_DLIGHTS);
		if (v->dowarp)
			v->warp_index = nwarp++;
		v->num_dlights = views[i].num_dlights > 0 ? views[i].num_dlights : 0;
		v->dlight_ofs = ndl;
		for (j = 0; j < v->num_dlights; j++, ndl++)
		{
			memcpy (rc.dlscratch[ndl].origin, views[i].dlights[j].origin, sizeof(float) * 3);
			rc.dlscratch[ndl].radius = views[i].dlights[j].radius;
		}
		/* brush entities: R_DrawBEntitiesOnList, r_main.c:640-768 */
		if (rc.cvars.r_drawentities && rc.world && !(views[i].flags & RDF_NOWORLDMODEL))
		{
			int border = 0, k;
			for (j = 0; j < views[i].num_entities; j++)
			{
				const entity_t *e = &views[i].entities[j];
				rc_bent_t *be;
				if (!e->model || e->model->type != rc_mod_brush)
					continue;
				if (e->model->world != rc.world->world)
					return rc_fail (RC_ERR_ARG, "view %d: brush entity %d does not belong to the loaded world", i, j);
				be = &rc.bentscratch[nbents];
				if (!RC_BmodelSetup (rc.world->world, e->model->submodel, e, &views[i], v, rc.vid.aspect, be))
					continue;
				be->view = i;
				be->order = border++;
				if (v->num_dlights)
				{
					/* r_main.c:708-735: the lights are moved into the entity's frame while it is marked and lit */
					int rotated = (e->angles[0] || e->angles[1] || e->angles[2]);
					float axis[3][3];
					if (rotated)
						RC_AngleVectors (e->angles, axis[0], axis[1], axis[2]);
					be->dlight_ofs = ndl;
					for (k = 0; k < v->num_dlights; k++, ndl++)
					{
						const dlight_t *l = &views[i].dlights[k];
						float temp[3];
						temp[0] = l->origin[0] - e->origin[0]; temp[1] = l->origin[1] - e->origin[1]; temp[2] = l->origin[2] - e->origin[2];
						if (rotated)
						{
							rc.dlscratch[ndl].origin[0] = temp[0]*axis[0][0] + temp[1]*axis[0][1] + temp[2]*axis[0][2];
							rc.dlscratch[ndl].origin[1] = -(temp[0]*axis[1][0] + temp[1]*axis[1][1] + temp[2]*axis[1][2]);
							rc.dlscratch[ndl].origin[2] = temp[0]*axis[2][0] + temp[1]*axis[2][1] + temp[2]*axis[2][2];
						}
						else
							memcpy (rc.dlscratch[ndl].origin, temp, sizeof(temp));
						rc.dlscratch[ndl].radius = l->radius;
					}
				}
				if (be->numfaces > maxbfaces) maxbfaces = be->numfaces;
				nbents++;
			}
		}
	}
	/* ---- entities and particles (R_DrawEntitiesOnList r_main.c:489-533, R_DrawParticles :773) ---- */
	{
		int nent = 0, npart = 0, nalias = 0, nverts = 0, ncmap = 1, nzres = 0, maxcmaps = 1, nsprites = 0;
		const byte *cmaps[64];
		for (i = 0; i < n; i++)
		{
			nent += views[i].num_entities > 0 ? views[i].num_entities : 0;
			npart += views[i].num_particles > 0 ? views[i].num_particles : 0;
		}
		if (rc.sprite_dirty && nent)
		{
			rc_hostsprite_t *smodels[RC_MAX_MODELS];
			int nsm = 0;
			for (i = 0; i < rc.nummodels; i++)
				if (rc.models[i]->type == rc_mod_sprite)
					smodels[nsm++] = rc.models[i]->sprite;
			if (rc_gpu_upload_sprites (rc.gpu, smodels, nsm, rc.err, sizeof(rc.err)) != RC_OK)
				return RC_ERR_CUDA;
			rc.sprite_dirty = 0;
		}
		if (nent > rc.spritescratchcap)
		{
			free (rc.spritescratch);
			rc.spritescratch = (rc_spriteinst_t *)malloc (sizeof(rc_spriteinst_t) * nent);
			rc.spritescratchcap = rc.spritescratch ? nent : 0;
			if (!rc.spritescratch) return rc_fail (RC_ERR_NOMEM, "out of memory");
		}
		if (nent && rc_alias_make_resident (views, n) != RC_OK)
			return RC_ERR_CUDA;
		if (nent > rc.aliasscratchcap)
		{
			free (rc.aliasscratch);
			rc.aliasscratch = (rc_aliasinst_t *)malloc (sizeof(rc_aliasinst_t) * nent);
			rc.aliasscratchcap = rc.aliasscratch ? nent : 0;
			if (!rc.aliasscratch) return rc_fail (RC_ERR_NOMEM, "out of memory");
		}
		if (npart > rc.partscratchcap)
		{
			free (rc.partscratch);
			rc.partscratch = (rc_particle_t *)malloc (sizeof(rc_particle_t) * npart);
			rc.partscratchcap = rc.partscratch ? npart : 0;
			if (!rc.partscratch) return rc_fail (RC_ERR_NOMEM, "out of memory");
		}
		(void)maxcmaps;
		cmaps[0] = NULL;
		if (n + 1 > rc.entofscap)
		{
			free (rc.entofs);
			rc.entofs = (int *)malloc (sizeof(int) * 4 * (size_t)(n + 1));	/* per view: first entity slot, first particle, alias count, sprite count */
			rc.entofscap = rc.entofs ? n + 1 : 0;
			if (!rc.entofs) return rc_fail (RC_ERR_NOMEM, "out of memory");
		}
		{
			int eo = 0, po = 0;
			for (i = 0; i < n; i++)
			{
				rc.entofs[4 * i] = eo; rc.entofs[4 * i + 1] = po;
				eo += views[i].num_entities > 0 ? views[i].num_entities : 0;
				po += views[i].num_particles > 0 ? views[i].num_particles : 0;
			}
		}
		/* R_DrawAliasModel / R_DrawSprite set-up (bounding box against the frustum, R_LightPoint's BSP descent, the
		 * transform) is per entity and reads only the loaded models and the view: the views of a batch are set up by all
		 * the host threads into their own slices of the scratch arrays (C3: 8,192 entities per batch were 13 ms on one
		 * thread -- longer than the GPU's step), then one serial pass packs the slices and hands out what is global to the
		 * batch (vertex pool offsets, the colormap table, the resolve planes) */
		{
			int toomany = -1;
			#pragma omp parallel for schedule(dynamic, 1) num_threads(rc.host_threads) private(j) if (nent >= 256 && rc.host_threads > 1)
			for (i = 0; i < n; i++)
			{
				rc_view_t *v = &rc.viewscratch[i];
				rc_aliasinst_t *abase = rc.aliasscratch + rc.entofs[4 * i];
				rc_spriteinst_t *sbase = rc.spritescratch + rc.entofs[4 * i];
				rc_particle_t *pbase = rc.partscratch + rc.entofs[4 * i + 1];
				int order = 0, na = 0, ns = 0;
				if (rc.cvars.r_drawentities)
					for (j = 0; j < views[i].num_entities; j++)
					{
						const entity_t *e = &views[i].entities[j];
						const struct model_s *mod = e->model;
						rc_aliasinst_t *inst;
						int c;
						if (!mod) continue;
						if ((e->flags & RF_WEAPONMODEL) && !rc.cvars.r_drawviewmodel) continue;
						if (mod->type != rc_mod_alias && mod->type != rc_mod_sprite)
							continue;		/* brush entities are handled with the world (R_DrawBEntitiesOnList) */
						if (e->flags & RF_CULLVIS)
						{
							float mins[3], maxs[3];
							for (c = 0; c < 3; c++) { mins[c] = e->origin[c] + mod->mins[c]; maxs[c] = e->origin[c] + mod->maxs[c]; }
							if (!rc.world || !RC_VisCullEntity (rc.world->world, views[i].vieworg, mins, maxs))
								continue;
						}
						if (order >= 32766)
						{
							#pragma omp critical
							{ if (toomany < 0 || i < toomany) toomany = i; }
							break;
						}
						if (mod->type == rc_mod_sprite)
						{
							/* R_DrawSprite, r_sprite.c:288-403 */
							rc_spriteinst_t *si = &sbase[ns];
							if (!RC_SpriteSetup (mod->sprite, e, &views[i], v, mod->synctype, si))
								continue;
							si->view = i;
							si->order = ++order;
							si->pixofs += mod->sprite->device_ofs;
							ns++;
							continue;
						}
						inst = &abase[na];
						if (!RC_AliasSetup (mod->alias, rc.world ? rc.world->world : NULL, e, &views[i], v, &rc.cvars, inst))
							continue;
						inst->view = i;
						inst->model = mod->alias->device_index;
						inst->order = ++order;
						inst->vertbase = mod->alias->numverts;	/* the serial pass turns it into the pool offset */
						inst->tribase = 0;
						/* entity colormap (r_alias.c:770): NULL or the library's own -> vid.colormap; others get their table index
						 * in the serial pass (here: the entity's index + 1) */
						inst->colormap = (e->colormap && e->colormap != rc.colormap) ? j + 1 : 0;
						na++;
					}
				for (j = 0; j < views[i].num_particles; j++)
				{
					rc_particle_t *p = &pbase[j];
					memcpy (p->org, views[i].particles[j].org, sizeof(float) * 3);
					p->color = views[i].particles[j].color;
					p->view = i; p->order = j; p->pad[0] = p->pad[1] = 0;
				}
				rc.entofs[4 * i + 2] = na; rc.entofs[4 * i + 3] = ns;
			}
			if (toomany >= 0) return rc_fail (RC_ERR_ARG, "view %d: too many entities", toomany);
		}
		npart = 0;
		for (i = 0; i < n; i++)
		{
			rc_view_t *v = &rc.viewscratch[i];
			const int na = rc.entofs[4 * i + 2], ns = rc.entofs[4 * i + 3];
			const int np = views[i].num_particles > 0 ? views[i].num_particles : 0;
			int k;
			v->zres_index = -1;
			if (ns && nsprites != rc.entofs[4 * i])
				memmove (&rc.spritescratch[nsprites], &rc.spritescratch[rc.entofs[4 * i]], sizeof(rc_spriteinst_t) * ns);
			nsprites += ns;
			if (na && nalias != rc.entofs[4 * i])
				memmove (&rc.aliasscratch[nalias], &rc.aliasscratch[rc.entofs[4 * i]], sizeof(rc_aliasinst_t) * na);
			for (k = 0; k < na; k++)
			{
				rc_aliasinst_t *inst = &rc.aliasscratch[nalias + k];
				const int numverts = inst->vertbase;
				inst->vertbase = nverts;
				nverts += numverts;
				if (inst->colormap)
				{
					const byte *cm = views[i].entities[inst->colormap - 1].colormap;
					int c;
					for (c = 1; c < ncmap; c++)
						if (cmaps[c] == cm) break;
					if (c == ncmap)
					{
						if (ncmap >= 64) return rc_fail (RC_ERR_ARG, "more than 63 distinct entity colormaps in a batch");
						cmaps[ncmap++] = cm;
					}
					inst->colormap = c;
				}
			}
			nalias += na;
			npart += np;		/* the particle slices are dense already (every particle of a view is kept) */
			if (na || ns || np)
				v->zres_index = nzres++;
		}
		if (ncmap * 16384 > rc.cmapscratchcap)
		{
			free (rc.cmapscratch);
			rc.cmapscratch = (uint8_t *)malloc ((size_t)ncmap * 16384);
			rc.cmapscratchcap = rc.cmapscratch ? ncmap * 16384 : 0;
			if (!rc.cmapscratch) return rc_fail (RC_ERR_NOMEM, "out of memory");
		}
		memcpy (rc.cmapscratch, rc.colormap, 16384);
		for (i = 1; i < ncmap; i++)
			memcpy (rc.cmapscratch + (size_t)i * 16384, cmaps[i], 16384);
		rc.batch.alias = rc.aliasscratch; rc.batch.nalias = nalias; rc.batch.nfinalverts = nverts; rc.batch.ntriwork = 0;
		rc.batch.particles = rc.partscratch; rc.batch.nparticles = npart;
		rc.batch.sprites = rc.spritescratch; rc.batch.nsprites = nsprites;
		rc.batch.colormaps = rc.cmapscratch; rc.batch.ncolormaps = ncmap;
		rc.batch.nzres = nzres;
	}
	rc.batch.views = rc.viewscratch; rc.batch.n = n;
	rc.batch.dlights = rc.dlscratch; rc.batch.ndlights = ndl;
	rc.batch.nwarp = nwarp;
	rc.batch.bents = rc.bentscratch; rc.batch.nbents = nbents; rc.batch.maxbfaces = maxbfaces;
	return RC_OK;
}

int RC_RenderViews (const refdef_t *views, int n, uint8_t *fb_out, int16_t *z_out, int *status)
{
	int r = rc_prepare_views (views, n);
	if (r != RC_OK) return r;
	if (!fb_out) return rc_fail (RC_ERR_ARG, "fb_out is NULL");
	r = rc_gpu_render_host (rc.gpu, &rc.batch, fb_out, z_out, status, rc.err, sizeof(rc.err));
	return r;
}

int RC_RenderViewsAsync (const refdef_t *views, int n, uint8_t *fb_out, int16_t *z_out)
{
	int r = rc_prepare_views (views, n);
	if (r != RC_OK) return r;
	if (!fb_out) return rc_fail (RC_ERR_ARG, "fb_out is NULL");
	r = rc_gpu_render_host_async (rc.gpu, &rc.batch, fb_out, z_out, rc.err, sizeof(rc.err));
	return r;
}

int RC_WaitViews (int *status)
{
	if (!rc.gpu) return rc_fail (RC_ERR_ARG, "RC_Init first");
	return rc_gpu_wait_host (rc.gpu, status, rc.err, sizeof(rc.err));
}

/* a large batch is drawn in about eight parts of at least RC_PART_MIN views (below) */
#define RC_PART_MIN 1024

int RC_RenderViewsDevice (const refdef_t *views, int n, void *fb_dev, void *z_dev, void *stream, int *status)
{
	int r;
	int entparts = 0;
	if (rc.inited && views && fb_dev && n >= 8 && n < 2 * RC_PART_MIN && rc.ent_parts > 1)
	{
		/* a batch whose host set-up is entity work (R_AliasCheckBBox, R_LightPoint, the transform ... per entity;
		 * BASELINE config C3: 8,192 entities and 262,144 particles per 32-view call) is drawn in parts as well */
		long long nent = 0;
		int i;
		for (i = 0; i < n; i++) nent += views[i].num_entities + (views[i].num_particles >> 5);
		if (nent >= 4096) entparts = rc.ent_parts;
	}
	if (rc.inited && views && fb_dev && (n >= 2 * RC_PART_MIN || entparts) && rc_gpu_can_split (rc.gpu))
	{
		const int part = entparts ? (n + entparts - 1) / entparts : (((n + 7) / 8 + 255) / 256 * 256 > RC_PART_MIN ? ((n + 7) / 8 + 255) / 256 * 256 : RC_PART_MIN);
		/* a large batch (BASELINE config C4: 65,536 views per call) in parts: the set-up of part k+1 -- 5 ms of host
		 * work and 47 MB of view records for the whole of C4 -- runs while the device draws part k */
		const size_t px = (size_t)rc.vid.width * rc.vid.height;
		int first, st = 0;
		double t_prep = 0, t_enq = 0, t_fin = 0;
		struct timespec ta, tb;
		r = rc_gpu_render_begin (rc.gpu, fb_dev, stream);
		for (first = 0; r == RC_OK && first < n; first += part)
		{
			const int cnt = n - first < part ? n - first : part;
			clock_gettime (CLOCK_MONOTONIC, &ta);
			r = rc_prepare_views (views + first, cnt);
			clock_gettime (CLOCK_MONOTONIC, &tb);
			t_prep += (tb.tv_sec - ta.tv_sec) * 1e3 + (tb.tv_nsec - ta.tv_nsec) * 1e-6;
			if (r == RC_OK)
				r = rc_gpu_render_part (rc.gpu, &rc.batch, (uint8_t *)fb_dev + px * (size_t)first,
							z_dev ? (void *)((int16_t *)z_dev + px * (size_t)first) : NULL, stream, first == 0, rc.err, sizeof(rc.err));
			clock_gettime (CLOCK_MONOTONIC, &ta);
			t_enq += (ta.tv_sec - tb.tv_sec) * 1e3 + (ta.tv_nsec - tb.tv_nsec) * 1e-6;
		}
		clock_gettime (CLOCK_MONOTONIC, &ta);
		if (r == RC_OK)
			r = rc_gpu_render_finish (rc.gpu, fb_dev, z_dev, n, stream, &st, rc.err, sizeof(rc.err));
		clock_gettime (CLOCK_MONOTONIC, &tb);
		t_fin = (tb.tv_sec - ta.tv_sec) * 1e3 + (tb.tv_nsec - ta.tv_nsec) * 1e-6;
		if (getenv ("RC_TIMING"))
			fprintf (stderr, "RC_RenderViewsDevice in parts of %d: host set-up %.2f ms, enqueue (with uploads) %.2f ms, wait at the end %.2f ms\n", part, t_prep, t_enq, t_fin);
		if (r != RC_OK)
		{
			int dummy;
			rc_gpu_render_finish (rc.gpu, fb_dev, z_dev, 0, stream, &dummy, rc.err, sizeof(rc.err));	/* drain what was enqueued */
			return r;
		}
		if (!(st & (RC_STAT_SPAN_POOL | RC_STAT_CACHE_POOL | RC_STAT_AET_OVERFLOW)))
		{
			if (status) *status = st;
			return RC_OK;
		}
		/* a pool ran out in one of the parts: the whole batch again in one piece, which grows the pools and retries */
	}
	r = rc_prepare_views (views, n);
	if (r != RC_OK) return r;
	if (!fb_dev) return rc_fail (RC_ERR_ARG, "fb_dev is NULL");
	r = rc_gpu_render (rc.gpu, &rc.batch, fb_dev, z_dev, stream, status, rc.err, sizeof(rc.err));
	return r;
}

void RC_SetGroupCallback (rc_group_fn fn, void *user, void *side_stream, int ngroups)
{
	if (rc.gpu) rc_gpu_set_group_callback (rc.gpu, fn, user, side_stream, ngroups);
}

int  RC_DeviceAlloc (size_t bytes, void **ptr) { return rc.gpu ? rc_gpu_device_alloc (rc.gpu, bytes, ptr) : RC_ERR_ARG; }
void RC_DeviceFree (void *ptr) { if (rc.gpu) rc_gpu_device_free (rc.gpu, ptr); }
int  RC_IpcExport (void *ptr, unsigned char handle[64]) { return rc.gpu ? rc_gpu_ipc_export (rc.gpu, ptr, handle) : RC_ERR_ARG; }
int  RC_IpcImport (const unsigned char handle[64], void **ptr) { return rc.gpu ? rc_gpu_ipc_import (rc.gpu, handle, ptr) : RC_ERR_ARG; }
void RC_IpcClose (void *ptr) { if (rc.gpu) rc_gpu_ipc_close (rc.gpu, ptr); }
void RC_SetMirrorOutput (void *fb, void *z) { if (rc.gpu) rc_gpu_set_mirror (rc.gpu, fb, z); }
void RC_SetMirrorGroups (int ngroups) { if (rc.gpu) rc_gpu_set_mirror_groups (rc.gpu, ngroups); }
void RC_SetMirrorDeferred (int on) { if (rc.gpu) rc_gpu_set_mirror_deferred (rc.gpu, on); }
int  RC_MirrorFence (void *stream) { return rc.gpu ? rc_gpu_mirror_fence (rc.gpu, stream, rc.err, sizeof(rc.err)) : RC_ERR_ARG; }

int RC_SetSkybox (const char *name)
{
	/* R_LoadSkybox, r_sky.c:138-187 */
	static const char *suf[6] = {"rt", "bk", "lf", "ft", "up", "dn"};
	static const int r_skysideimage[6] = {5, 2, 4, 1, 0, 3};
	uint8_t *pix;
	int i, r;
	if (!rc.inited) return rc_fail (RC_ERR_ARG, "RC_Init first");
	if (!name || !name[0])
		return rc_gpu_set_skybox (rc.gpu, NULL, rc.err, sizeof(rc.err));
	pix = (uint8_t *)malloc (RC_SKYBOX_BYTES);
	if (!pix) return rc_fail (RC_ERR_NOMEM, "out of memory");
	for (i = 0; i < 6; i++)
	{
		char path[256];
		size_t len = 0;
		int w = 0, h = 0;
		void *buf;
		unsigned char *pic;
		snprintf (path, sizeof(path), "env/%s%s.pcx", name, suf[r_skysideimage[i]]);
		buf = rc_read_file (path, &len);
		pic = buf ? RC_DecodePCX (buf, len, &w, &h) : NULL;
		free (buf);
		if (!pic || w != 256 || h != 256)
		{
			r = pic ? rc_fail (RC_ERR_FORMAT, "%s must be 256x256", path) : rc_fail (RC_ERR_FILE, "Couldn't load %s", path);
			free (pic); free (pix);
			rc_gpu_set_skybox (rc.gpu, NULL, rc.err + strlen (rc.err), 0);
			return r;
		}
		memcpy (pix + (size_t)i * 65536, pic, 65536);
		free (pic);
	}
	r = rc_gpu_set_skybox (rc.gpu, pix, rc.err, sizeof(rc.err));
	free (pix);
	return r;
}

void RC_FlushCaches (void) { if (rc.gpu) rc_gpu_flush_caches (rc.gpu); }
void RC_GetStats (rc_stats_t *out) { if (rc.gpu) rc_gpu_get_stats (rc.gpu, out); else memset (out, 0, sizeof(*out)); }
void RC_SetProfiling (int on) { if (rc.gpu) rc_gpu_set_profiling (rc.gpu, on); }
/* test hook: the memory image of a loaded alias model (the reference's cache block from `pheader` on, see
 * rc_host_alias.c) -> its length; copies min (length, cap) bytes */
int RC_DebugAliasImage (struct model_s *model, void *dst, int cap)
{
	if (!model || model->type != rc_mod_alias || !model->alias) return RC_ERR_ARG;
	if (dst && cap > 0)
		memcpy (dst, model->alias->skinpixels, (size_t)(cap < model->alias->skinimagelen ? cap : model->alias->skinimagelen));
	return model->alias->skinimagelen;
}
int  RC_DebugRead (int kind, int view, void *dst, int cap) { return rc.gpu ? rc_gpu_debug_read (rc.gpu, kind, view, dst, cap) : RC_ERR_ARG; }

/* R_BuildGammaTable, ref_soft/r_main.c:794-817 (libm pow on the host, as the reference) */
static void rc_build_gamma (void)
{
	int i, inf;
	float g = rc.gamma_valid ? rc.gamma : 1.0f;
	if (g == 1.0)
	{
		for (i = 0; i < 256; i++)
			rc.gammatable[i] = (uint8_t)i;
	}
	else
		for (i = 0; i < 256; i++)
		{
			inf = (int)(255.0 * pow (((float)i + 0.5) / 255.5, g) + 0.5);
			if (inf < 0) inf = 0;
			if (inf > 255) inf = 255;
			rc.gammatable[i] = (uint8_t)inf;
		}
	rc.gamma_valid = 2;
}

void RC_SetLoadLits (int on) { rc.loadlits = on; }

/* R_TimeRefresh_f (the `timerefresh` console command, ref_soft/r_misc.c:29-64): 128 views of the given refdef with
 * viewangles[1] = i / 128 * 360 -- here one batch instead of 128 frames.  fb_out (may be NULL) receives the 128 frames;
 * *seconds the wall time of the batch, host copies included when fb_out is given.  Returns RC_OK or RC_ERR_*. */
int RC_TimeRefresh (const refdef_t *base, uint8_t *fb_out, double *seconds, int *status)
{
	static refdef_t rds[128];
	struct timespec t0, t1;
	int i, r;
	if (!rc.inited || !base) return rc_fail (RC_ERR_ARG, "RC_TimeRefresh: bad argument");
	for (i = 0; i < 128; i++)
	{
		rds[i] = *base;
		rds[i].viewangles[1] = i / 128.0 * 360.0;
	}
	clock_gettime (CLOCK_MONOTONIC, &t0);
	if (fb_out)
		r = RC_RenderViews (rds, 128, fb_out, NULL, status);
	else
	{
		size_t px = (size_t)rc.vid.width * rc.vid.height;
		void *fb = NULL;
		if (RC_DeviceAlloc (px * 128, &fb) != RC_OK) return rc_fail (RC_ERR_NOMEM, "RC_TimeRefresh: device memory");
		r = RC_RenderViewsDevice (rds, 128, fb, NULL, NULL, status);
		RC_DeviceFree (fb);
	}
	clock_gettime (CLOCK_MONOTONIC, &t1);
	if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
	return r;
}

/* What the client asks before it builds a refdef_t (cl_view.c:786-791,865-870: contents <= CONTENTS_WATER sets
 * RDF_UNDERWATER): the contents of the world leaf that holds the point, Mod_PointInLeaf (r_model.c:83-105) ->contents.
 * Returns the CONTENTS_* value (negative), or 0 without a world. */
int RC_PointContents (const float *org)
{
	if (!rc.world || !rc.world->world || !org) return 0;
	return rc.world->world->w.leafs[RC_PointInLeaf (rc.world->world, org)].contents;
}

int RC_ViewFlagsForOrigin (const float *org)
{
	return RC_PointContents (org) <= -3 /* CONTENTS_WATER, bspfile.h */ ? RDF_UNDERWATER : 0;
}

void RC_SetGamma (float gamma)
{
	rc.gamma = gamma; rc.gamma_valid = 1;
}

/* R_UpdatePalette, ref_soft/r_main.c:839-875: colour shifts blended in 8.8 fixed point one after the other,
 * then the gamma table.  (r can leave 0..255 only with percent > 256 or destcolor outside 0..255; the
 * reference indexes gammatable with it regardless, here it is clamped to the table.) */
int RC_ViewPalette (const refdef_t *view, unsigned char *pal)
{
	int i, j, c;
	if (!view || !pal) return rc_fail (RC_ERR_ARG, "RC_ViewPalette: NULL argument");
	if (rc.gamma_valid != 2) rc_build_gamma ();
	for (i = 0; i < 256; i++)
	{
		int rgb[3];
		rgb[0] = rc.palette[i * 3]; rgb[1] = rc.palette[i * 3 + 1]; rgb[2] = rc.palette[i * 3 + 2];
		for (j = 0; j < view->num_cshifts; j++)
			for (c = 0; c < 3; c++)
				rgb[c] += (view->cshifts[j].percent * (view->cshifts[j].destcolor[c] - rgb[c])) >> 8;
		for (c = 0; c < 3; c++)
			pal[i * 3 + c] = rc.gammatable[rgb[c] < 0 ? 0 : (rgb[c] > 255 ? 255 : rgb[c])];
	}
	return RC_OK;
}

static int rc_view_palettes32 (const refdef_t *views, int n)
{
	int i, k;
	unsigned char pal[768];
	if (n * 256 > rc.palscratchcap)
	{
		free (rc.palscratch);
		rc.palscratch = (uint32_t *)malloc (sizeof(uint32_t) * 256 * (size_t)n);
		rc.palscratchcap = rc.palscratch ? n * 256 : 0;
		if (!rc.palscratch) return rc_fail (RC_ERR_NOMEM, "out of memory");
	}
	for (i = 0; i < n; i++)
	{
		RC_ViewPalette (&views[i], pal);
		for (k = 0; k < 256; k++)
			rc.palscratch[i * 256 + k] = (uint32_t)pal[k * 3] | ((uint32_t)pal[k * 3 + 1] << 8) | ((uint32_t)pal[k * 3 + 2] << 16) | 0xFF000000u;
	}
	return RC_OK;
}

int RC_PresentViewsDevice (const refdef_t *views, int n, const void *fb_dev, void *rgbx_dev, void *stream)
{
	int r;
	if (!rc.gpu) return rc_fail (RC_ERR_ARG, "RC_Init first");
	if (n <= 0 || !fb_dev || !rgbx_dev) return rc_fail (RC_ERR_ARG, "RC_PresentViewsDevice: bad argument");
	if (!views)		/* the palettes of the previous call again (same n): nothing is uploaded */
		return rc_gpu_present (rc.gpu, NULL, n, fb_dev, rgbx_dev, stream, 0, rc.err, sizeof(rc.err));
	if ((r = rc_view_palettes32 (views, n)) != RC_OK) return r;
	return rc_gpu_present (rc.gpu, rc.palscratch, n, fb_dev, rgbx_dev, stream, 0, rc.err, sizeof(rc.err));
}

int RC_PresentViews (const refdef_t *views, int n, const uint8_t *fb_in, uint32_t *rgbx_out)
{
	int r;
	if (!rc.gpu) return rc_fail (RC_ERR_ARG, "RC_Init first");
	if (!views || n <= 0 || !fb_in || !rgbx_out) return rc_fail (RC_ERR_ARG, "RC_PresentViews: bad argument");
	if ((r = rc_view_palettes32 (views, n)) != RC_OK) return r;
	return rc_gpu_present (rc.gpu, rc.palscratch, n, fb_in, rgbx_out, NULL, 1, rc.err, sizeof(rc.err));
}

/* ------------------------------------------------------------------ 2D overlay (r_draw.c) ---- */
typedef struct { char identification[4]; int numlumps; int infotableofs; } rc_wadinfo_t;			/* wad.h:47-52 */
typedef struct { int filepos, disksize, size; char type, compression, pad1, pad2; char name[16]; } rc_lumpinfo_t;	/* wad.h:54-63 */
struct mpic_s { int width; short height; unsigned char alpha, pad; unsigned char data[4]; };		/* r_shared.h:78-85 */

static int rc_d2_append (const void *data, size_t len, int align)
{
	size_t ofs = (rc.d2len + (size_t)align - 1) & ~((size_t)align - 1);
	if (ofs + len > rc.d2cap)
	{
		size_t ncap = (ofs + len) * 2 + 65536;
		uint8_t *nb = (uint8_t *)realloc (rc.d2blob, ncap);
		if (!nb) return -1;
		memset (nb + rc.d2cap, 0, ncap - rc.d2cap);
		rc.d2blob = nb; rc.d2cap = ncap;
	}
	if (data) memcpy (rc.d2blob + ofs, data, len);
	rc.d2len = ofs + len;
	rc.d2dirty = 1;
	return (int)ofs;
}

/* W_GetLumpName (wad.c:124-140) with W_CleanupName's case folding */
static const uint8_t *rc_wad_lump (const char *name, int *size)
{
	const rc_wadinfo_t *h = (const rc_wadinfo_t *)rc.wad;
	const rc_lumpinfo_t *l;
	int i, k;
	if (!rc.wad) return NULL;
	l = (const rc_lumpinfo_t *)(rc.wad + h->infotableofs);
	for (i = 0; i < h->numlumps; i++, l++)
	{
		for (k = 0; k < 16; k++)
		{
			int a = (unsigned char)l->name[k], b = (unsigned char)name[k];
			if (a >= 'A' && a <= 'Z') a += 'a' - 'A';
			if (b >= 'A' && b <= 'Z') b += 'a' - 'A';
			if (a != b) break;
			if (!a) { k = 16; break; }
		}
		if (k == 16)
		{
			if ((size_t)l->filepos + (size_t)l->size > rc.wadlen) return NULL;
			if (size) *size = l->size;
			return rc.wad + l->filepos;
		}
	}
	return NULL;
}

/* registers a qpic_t (int width, height, data) under `name`; returns the handle */
static int rc_d2_add_qpic (const char *name, const uint8_t *qpic, size_t len, int slot)
{
	int w, h, ofs, i;
	if (len < 8) return rc_fail (RC_ERR_FORMAT, "%s: not a qpic", name);
	memcpy (&w, qpic, 4); memcpy (&h, qpic + 4, 4);
	if (w <= 0 || h <= 0 || w > 4096 || h > 4096 || (size_t)w * h + 8 > len) return rc_fail (RC_ERR_FORMAT, "%s: bad qpic size %dx%d", name, w, h);
	if (slot < 0)
	{
		if (rc.d2npics >= RC_MAX_PICS) return rc_fail (RC_ERR_NOMEM, "too many pictures");
		slot = rc.d2npics++;
	}
	ofs = rc_d2_append (qpic + 8, (size_t)w * h, 16);
	if (ofs < 0) return rc_fail (RC_ERR_NOMEM, "out of memory");
	rc.d2pics[slot].ofs = ofs; rc.d2pics[slot].width = w; rc.d2pics[slot].height = h;
	rc.d2pics[slot].alpha = memchr (qpic + 8, 255, (size_t)w * h) != NULL;		/* r_draw.c:62 */
	snprintf (rc.d2names[slot], sizeof(rc.d2names[slot]), "%s", name);
	free (rc.d2mpics[slot]);
	rc.d2mpics[slot] = (struct mpic_s *)malloc (sizeof(struct mpic_s) + (size_t)w * h);
	if (rc.d2mpics[slot])
	{
		rc.d2mpics[slot]->width = w; rc.d2mpics[slot]->height = (short)h; rc.d2mpics[slot]->alpha = (unsigned char)rc.d2pics[slot].alpha; rc.d2mpics[slot]->pad = 0;
		memcpy (rc.d2mpics[slot]->data, qpic + 8, (size_t)w * h);
	}
	(void)i;
	return slot;
}

/* Draw_CharToConback, r_draw.c:377-400: the version string stamped into the console picture */
static void rc_d2_stamp_conback (void)
{
	const rc_pic_t *pc = &rc.d2pics[2];
	const char *ver = rc.d2version;
	size_t n = strlen (ver), x;
	uint8_t *dest0;
	if (pc->width != 320 || pc->height != 200 || !n || 11 + 8 * n > 320) return;
	dest0 = rc.d2blob + pc->ofs + 320 * 186 + 320 - 11 - 8 * n;
	for (x = 0; x < n; x++)
	{
		int num = (unsigned char)ver[x], drawline, xx;
		const uint8_t *source = rc.d2blob + ((num >> 4) << 10) + ((num & 15) << 3);
		uint8_t *dest = dest0 + (x << 3);
		for (drawline = 0; drawline < 8; drawline++, source += 128, dest += 320)
			for (xx = 0; xx < 8; xx++)
				if (source[xx])
					dest[xx] = (uint8_t)(0x60 + source[xx]);
	}
	rc.d2dirty = 1;
}

void RC_SetConsoleVersion (const char *text)
{
	snprintf (rc.d2version, sizeof(rc.d2version), "%s", text ? text : "");
}

int RC_DrawInit (void)
{
	const uint8_t *l;
	int size = 0, r;
	size_t len;
	uint8_t *cb;
	if (!rc.inited) return rc_fail (RC_ERR_ARG, "RC_Init first");
	if (rc.d2ready) return RC_OK;
	free (rc.wad);
	rc.wad = (uint8_t *)rc_read_file ("gfx.wad", &rc.wadlen);		/* host.c:815 W_LoadWadFile ("gfx.wad") */
	if (!rc.wad || rc.wadlen < sizeof(rc_wadinfo_t) || memcmp (rc.wad, "WAD2", 4))
		return rc_fail (RC_ERR_FILE, "W_LoadWadFile: couldn't load gfx.wad");
	{
		const rc_wadinfo_t *h = (const rc_wadinfo_t *)rc.wad;
		if (h->numlumps < 0 || h->infotableofs < 0 || (size_t)h->infotableofs + (size_t)h->numlumps * sizeof(rc_lumpinfo_t) > rc.wadlen)
			return rc_fail (RC_ERR_FORMAT, "gfx.wad: bad directory");
	}
	rc.d2len = 0;
	l = rc_wad_lump ("conchars", &size);
	if (!l || size < 128 * 128) return rc_fail (RC_ERR_FILE, "W_GetLumpinfo: conchars not found");
	if (rc_d2_append (l, 128 * 128, 16) != 0) return rc_fail (RC_ERR_NOMEM, "out of memory");
	rc.d2npics = 3;			/* 0 unused, 1 backtile, 2 conback */
	memset (&rc.d2pics[0], 0, sizeof(rc.d2pics[0]));
	l = rc_wad_lump ("backtile", &size);
	if (!l) return rc_fail (RC_ERR_FILE, "W_GetLumpinfo: backtile not found");
	if ((r = rc_d2_add_qpic ("backtile", l, (size_t)size, 1)) < 0) return r;
	if (!rc_wad_lump ("disc", &size)) return rc_fail (RC_ERR_FILE, "W_GetLumpinfo: disc not found");
	/* Draw_ConsoleBackground loads gfx/conback.lmp on first use (r_draw.c:415); absent: a 320x200 picture of colour 0 */
	cb = (uint8_t *)rc_read_file ("gfx/conback.lmp", &len);
	if (cb)
		r = rc_d2_add_qpic ("gfx/conback.lmp", cb, len, 2);
	else
	{
		uint8_t *blank = (uint8_t *)calloc (1, 8 + 320 * 200);
		int w = 320, h = 200;
		if (!blank) return rc_fail (RC_ERR_NOMEM, "out of memory");
		memcpy (blank, &w, 4); memcpy (blank + 4, &h, 4);
		r = rc_d2_add_qpic ("gfx/conback.lmp", blank, 8 + 320 * 200, 2);
		free (blank);
	}
	free (cb);
	if (r < 0) return r;
	rc_d2_stamp_conback ();
	rc.d2ready = 1;
	return RC_OK;
}

static int rc_d2_find (const char *name)
{
	int i;
	for (i = 3; i < rc.d2npics; i++)
		if (!strcmp (rc.d2names[i], name)) return i;
	return -1;
}

int RC_DrawPicFromWad (const char *name)
{
	const uint8_t *l;
	int size = 0, i, r;
	char key[64];
	if ((r = RC_DrawInit ()) != RC_OK) return r;
	snprintf (key, sizeof(key), "wad:%s", name);
	if ((i = rc_d2_find (key)) >= 0) return i;
	l = rc_wad_lump (name, &size);
	if (!l) return rc_fail (RC_ERR_FILE, "W_GetLumpinfo: %s not found", name);
	return rc_d2_add_qpic (key, l, (size_t)size, -1);
}

int RC_DrawCachePic (const char *path)
{
	int i, r;
	size_t len;
	uint8_t *buf;
	if ((r = RC_DrawInit ()) != RC_OK) return r;
	if (!strcmp (path, "gfx/conback.lmp")) return 2;
	if ((i = rc_d2_find (path)) >= 0) return i;
	buf = (uint8_t *)rc_read_file (path, &len);
	if (!buf) return rc_fail (RC_ERR_FILE, "Draw_CachePic: failed to load %s", path);
	r = rc_d2_add_qpic (path, buf, len, -1);
	free (buf);
	return r;
}

int RC_DrawTranslation (const unsigned char *table256)
{
	int r, ofs;
	if ((r = RC_DrawInit ()) != RC_OK) return r;
	if (!table256) return rc_fail (RC_ERR_ARG, "RC_DrawTranslation: NULL table");
	ofs = rc_d2_append (table256, 256, 16);
	if (ofs < 0) return rc_fail (RC_ERR_NOMEM, "out of memory");
	return ofs;			/* the handle is the table's place in the blob */
}

/* the checks the reference makes before it touches vid.buffer: Sys_Error there, RC_ERR_ARG here; commands that
 * the reference silently ignores (characters off screen) stay in the list and draw nothing */
static int rc_d2_validate (const rc_draw2d_t *c, int idx)
{
	const int W = rc.vid.width, H = rc.vid.height;
	switch (c->op)
	{
	case RC_D2_CHAR: return RC_OK;
	case RC_D2_PIC: case RC_D2_TRANSPIC: case RC_D2_TRANSPIC_TRANSLATE:
		if (c->pic < 1 || c->pic >= rc.d2npics) return rc_fail (RC_ERR_ARG, "draw command %d: bad picture handle %d", idx, c->pic);
		if (c->x < 0 || c->y < 0 || c->x + rc.d2pics[c->pic].width > W || c->y + rc.d2pics[c->pic].height > H)
			return rc_fail (RC_ERR_ARG, "draw command %d: Draw_Pic: bad coordinates", idx);		/* r_draw.c:237-243,267-271 */
		if (c->op == RC_D2_TRANSPIC_TRANSLATE && (c->table < 128 * 128 || (size_t)c->table + 256 > rc.d2len))
			return rc_fail (RC_ERR_ARG, "draw command %d: bad translation handle", idx);
		return RC_OK;
	case RC_D2_CONBACK:
		if (c->arg < 0 || c->arg > H) return rc_fail (RC_ERR_ARG, "draw command %d: Draw_ConsoleBackground: %d lines", idx, c->arg);
		return RC_OK;
	case RC_D2_TILECLEAR: case RC_D2_FILL:
		/* the reference writes whatever rectangle it is given; outside the screen that is a stray write */
		if (c->w < 0 || c->h < 0 || c->x < 0 || c->y < 0 || c->x + c->w > W || c->y + c->h > H)
			return rc_fail (RC_ERR_ARG, "draw command %d: rectangle outside the screen", idx);
		return RC_OK;
	case RC_D2_FADESCREEN: return RC_OK;
	}
	return rc_fail (RC_ERR_ARG, "draw command %d: unknown op %d", idx, c->op);
}

static int rc_d2_run (const rc_draw2d_t *cmds, const int *first, int n, void *fb, void *stream, int host)
{
	int i, r, total;
	if ((r = RC_DrawInit ()) != RC_OK) return r;
	if (!cmds || !first || n <= 0 || !fb) return rc_fail (RC_ERR_ARG, "RC_DrawOverlay: bad argument");
	total = first[n];
	for (i = 0; i < n; i++)
		if (first[i] < 0 || first[i] > first[i + 1]) return rc_fail (RC_ERR_ARG, "RC_DrawOverlay: `first` must be non-decreasing");
	for (i = first[0]; i < total; i++)
		if ((r = rc_d2_validate (&cmds[i], i)) != RC_OK) return r;
	if (rc.d2dirty)
	{
		if ((r = rc_gpu_upload_draw2d (rc.gpu, rc.d2blob, rc.d2len, rc.d2pics, rc.d2npics, rc.err, sizeof(rc.err))) != RC_OK) return r;
		rc.d2dirty = 0;
	}
	return rc_gpu_draw2d (rc.gpu, cmds, total, first, n, fb, stream, host, rc.err, sizeof(rc.err));
}

int RC_DrawOverlayDevice (const rc_draw2d_t *cmds, const int *first, int n, void *fb_dev, void *stream)
{
	return rc_d2_run (cmds, first, n, fb_dev, stream, 0);
}

int RC_DrawOverlay (const rc_draw2d_t *cmds, const int *first, int n, uint8_t *fb_inout)
{
	return rc_d2_run (cmds, first, n, fb_inout, NULL, 1);
}

/* ---- the render.h names (client/render.h:42-53): one recorded command each, applied by R_EndFrame ---- */
static void rc_d2_record (int op, int x, int y, int w, int h, int arg, int pic, int table)
{
	rc_draw2d_t *c;
	if (rc.d2framen >= rc.d2framecap)
	{
		int ncap = rc.d2framecap ? rc.d2framecap * 2 : 256;
		rc_draw2d_t *nb = (rc_draw2d_t *)realloc (rc.d2frame, sizeof(rc_draw2d_t) * (size_t)ncap);
		if (!nb) return;
		rc.d2frame = nb; rc.d2framecap = ncap;
	}
	c = &rc.d2frame[rc.d2framen++];
	c->op = op; c->x = x; c->y = y; c->w = w; c->h = h; c->arg = arg; c->pic = pic; c->table = table;
}

static int rc_d2_handle_of (struct mpic_s *pic)
{
	int i;
	for (i = 1; i < rc.d2npics; i++)
		if (rc.d2mpics[i] == pic) return i;
	return 0;
}

void Draw_Init (void) { if (RC_DrawInit () != RC_OK) fprintf (stderr, "ref_cuda: %s\n", rc.err); }
struct mpic_s *Draw_PicFromWad (char *name) { int h = RC_DrawPicFromWad (name); return h > 0 ? rc.d2mpics[h] : NULL; }
struct mpic_s *Draw_CachePic (char *path) { int h = RC_DrawCachePic (path); return h > 0 ? rc.d2mpics[h] : NULL; }
void Draw_Character (int x, int y, int num) { rc_d2_record (RC_D2_CHAR, x, y, 0, 0, num, 0, 0); }
void Draw_String (int x, int y, char *str) { for (; *str; str++, x += 8) Draw_Character (x, y, *str); }		/* r_draw.c:211-219 */
void Draw_Pic (int x, int y, struct mpic_s *pic) { rc_d2_record (RC_D2_PIC, x, y, 0, 0, 0, rc_d2_handle_of (pic), 0); }
void Draw_TransPic (int x, int y, struct mpic_s *pic) { rc_d2_record (RC_D2_TRANSPIC, x, y, 0, 0, 0, rc_d2_handle_of (pic), 0); }
void Draw_TransPicTranslate (int x, int y, struct mpic_s *pic, unsigned char *translation)
{
	int t = RC_DrawTranslation (translation);
	if (t > 0) rc_d2_record (RC_D2_TRANSPIC_TRANSLATE, x, y, 0, 0, 0, rc_d2_handle_of (pic), t);
}
void Draw_ConsoleBackground (int lines) { rc_d2_record (RC_D2_CONBACK, 0, 0, 0, 0, lines, 0, 0); }
void Draw_TileClear (int x, int y, int w, int h) { rc_d2_record (RC_D2_TILECLEAR, x, y, w, h, 0, 0, 0); }
void Draw_Fill (int x, int y, int w, int h, int c) { rc_d2_record (RC_D2_FILL, x, y, w, h, c, 0, 0); }
void Draw_FadeScreen (void) { rc_d2_record (RC_D2_FADESCREEN, 0, 0, 0, 0, 0, 0, 0); }

void RC_SetVidBuffers (uint8_t *buffer, int rowbytes, int16_t *zbuffer)
{
	rc.vidbuffer = buffer; rc.vidrowbytes = rowbytes; rc.zbuffer = zbuffer;
}

/* the render.h single-view path: one view through the batched pipeline, copied into the
 * engine's vid.buffer / d_pzbuffer */
void R_BeginFrame (void) { rc.d2framen = 0; }

/* r_main.c:961-980.  The frame's 2D commands are applied to the engine's vid.buffer on the device (upload, k_draw2d,
 * download); the palette is the caller's business (RC_ViewPalette -> VID_ShiftPalette). */
void R_EndFrame (void)
{
	if (rc.d2framen && rc.vidbuffer)
	{
		int first[2];
		first[0] = 0; first[1] = rc.d2framen;
		if (rc.vidrowbytes == rc.vid.width)
		{
			if (RC_DrawOverlay (rc.d2frame, first, 1, rc.vidbuffer) != RC_OK)
				fprintf (stderr, "ref_cuda: %s\n", rc.err);
		}
		else
		{
			/* vid.rowbytes "may be > width if displayed in a window" (client/vid.h:38): the rows are packed for
			 * the device and unpacked again */
			const size_t w = (size_t)rc.vid.width;
			uint8_t *tmp = (uint8_t *)malloc (w * rc.vid.height);
			int y;
			if (tmp)
			{
				for (y = 0; y < rc.vid.height; y++)
					memcpy (tmp + y * w, rc.vidbuffer + (size_t)y * rc.vidrowbytes, w);
				if (RC_DrawOverlay (rc.d2frame, first, 1, tmp) != RC_OK)
					fprintf (stderr, "ref_cuda: %s\n", rc.err);
				else
					for (y = 0; y < rc.vid.height; y++)
						memcpy (rc.vidbuffer + (size_t)y * rc.vidrowbytes, tmp + y * w, w);
				free (tmp);
			}
		}
	}
	rc.d2framen = 0;
}

/* What a frame touches in the engine's buffers (the rest keeps what earlier frames and the 2D code left there: sbar and
 * console redraw their areas only when they change):
 *   colour  the refdef rectangle (under water D_WarpScreen writes the same rectangle, d_scan.c:43-103)
 *   z       the refdef rectangle; under water the warp rectangle at the origin, <= 320x200 (r_misc.c:448-454);
 *           with RDF_NOWORLDMODEL the whole buffer (0xFFFF, d_modech.c:103-106) and of the colour plane only the
 *           entities' pixels -- the ones whose z is no longer 0xFFFF */
void R_RenderView (refdef_t *refdef)
{
	size_t px = (size_t)rc.vid.width * rc.vid.height;
	const int W = rc.vid.width;
	uint8_t *fb;
	int16_t *zb;
	int x, y, status = 0;
	if (!rc.vidbuffer) { rc_fail (RC_ERR_ARG, "RC_SetVidBuffers first"); return; }
	fb = (uint8_t *)malloc (px);
	zb = (int16_t *)malloc (px * 2);
	if (fb && zb && RC_RenderViews (refdef, 1, fb, zb, &status) == RC_OK)
	{
		const int x0 = refdef->x, y0 = refdef->y, x1 = refdef->x + refdef->width, y1 = refdef->y + refdef->height;
		if (refdef->flags & RDF_NOWORLDMODEL)
		{
			for (y = y0; y < y1; y++)
				for (x = x0; x < x1; x++)
					if (zb[(size_t)y * W + x] != (int16_t)-1)
						rc.vidbuffer[(size_t)y * rc.vidrowbytes + x] = fb[(size_t)y * W + x];
			if (rc.zbuffer) memcpy (rc.zbuffer, zb, px * 2);
		}
		else
		{
			int zx0 = x0, zy0 = y0, zx1 = x1, zy1 = y1;
			for (y = y0; y < y1; y++)
				memcpy (rc.vidbuffer + (size_t)y * rc.vidrowbytes + x0, fb + (size_t)y * W + x0, (size_t)(x1 - x0));
			if (rc.cvars.r_waterwarp && (refdef->flags & RDF_UNDERWATER))
			{
				zx0 = 0; zy0 = 0;
				zx1 = refdef->width < RC_WARP_WIDTH ? refdef->width : RC_WARP_WIDTH;
				zy1 = refdef->height < RC_WARP_HEIGHT ? refdef->height : RC_WARP_HEIGHT;
			}
			if (rc.zbuffer)
				for (y = zy0; y < zy1; y++)
					memcpy (rc.zbuffer + (size_t)y * W + zx0, zb + (size_t)y * W + zx0, (size_t)(zx1 - zx0) * 2);
		}
	}
	else
		fprintf (stderr, "ref_cuda: R_RenderView: %s\n", fb && zb ? rc.err : "out of memory");
	free (fb); free (zb);
}

/* WritePCXfile, ref_soft/pcx.c:103-160: 128-byte header, every byte with its two top bits set escaped by 0xc1 (no run
 * lengths), 0x0c and the 768-byte palette.  File output of a finished frame: host code. */
int RC_WritePCX (const char *path, const uint8_t *data, int width, int height, int rowbytes, const uint8_t *palette)
{
	uint8_t *buf, *pack;
	FILE *fp;
	int i, j;
	if (!path || !data || width <= 0 || height <= 0) return rc_fail (RC_ERR_ARG, "RC_WritePCX: bad argument");
	buf = (uint8_t *)calloc (1, (size_t)width * height * 2 + 1000);
	if (!buf) return rc_fail (RC_ERR_NOMEM, "SCR_ScreenShot_f: not enough memory");
	buf[0] = 0x0a; buf[1] = 5; buf[2] = 1; buf[3] = 8;			/* manufacturer, version, encoding, bits_per_pixel */
	buf[8] = (uint8_t)((width - 1) & 255); buf[9] = (uint8_t)(((width - 1) >> 8) & 255);		/* xmax */
	buf[10] = (uint8_t)((height - 1) & 255); buf[11] = (uint8_t)(((height - 1) >> 8) & 255);	/* ymax */
	buf[12] = (uint8_t)(width & 255); buf[13] = (uint8_t)((width >> 8) & 255);			/* hres */
	buf[14] = (uint8_t)(height & 255); buf[15] = (uint8_t)((height >> 8) & 255);			/* vres */
	buf[65] = 1;								/* color_planes (palette[48] at 16, reserved at 64) */
	buf[66] = (uint8_t)(width & 255); buf[67] = (uint8_t)((width >> 8) & 255);			/* bytes_per_line */
	buf[68] = 2;								/* palette_type */
	pack = buf + 128;
	for (i = 0; i < height; i++, data += rowbytes)
		for (j = 0; j < width; j++)
		{
			if ((data[j] & 0xc0) == 0xc0)
				*pack++ = 0xc1;
			*pack++ = data[j];
		}
	*pack++ = 0x0c;
	memcpy (pack, palette ? palette : rc.palette, 768);
	pack += 768;
	fp = fopen (path, "wb");
	if (!fp) { free (buf); return rc_fail (RC_ERR_FILE, "COM_WriteFile: couldn't open %s", path); }
	fwrite (buf, 1, (size_t)(pack - buf), fp);
	fclose (fp);
	free (buf);
	return RC_OK;
}

/* R_ScreenShot_f, ref_soft/r_misc.c:117-149: the first free <gamedir>/quakeNN.pcx, vid.buffer with the base palette */
void R_ScreenShot_f (void)
{
	char name[700];
	int i;
	FILE *fp;
	if (!rc.vidbuffer) { rc_fail (RC_ERR_ARG, "RC_SetVidBuffers first"); return; }
	for (i = 0; i <= 99; i++)
	{
		snprintf (name, sizeof(name), "%s/id1/quake%d%d.pcx", rc.basedir, i / 10, i % 10);
		if ((fp = fopen (name, "rb")) == NULL)
			break;
		fclose (fp);
	}
	if (i == 100) { fprintf (stderr, "SCR_ScreenShot_f: Couldn't create a PCX file\n"); return; }
	if (RC_WritePCX (name, rc.vidbuffer, rc.vid.width, rc.vid.height, rc.vidrowbytes, rc.palette) != RC_OK)
		fprintf (stderr, "ref_cuda: %s\n", rc.err);
}

static struct model_s *rc_register_model (char *name) { return Mod_ForName (name, 0); }

refexport_t GetRefAPI (int api_version)
{
	refexport_t re;
	re.api_version = api_version;
	re.Init = R_Init;
	re.BeginRegistration = R_NewMap;
	re.RegisterModel = rc_register_model;
	re.BeginFrame = R_BeginFrame;
	re.RenderFrame = R_RenderView;
	re.EndFrame = R_EndFrame;
	re.RenderFrames = RC_RenderViews;
	return re;
}
