/*
 * tests/hostsim/rc_sim_backend.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Implements the rc_gpu_* backend interface (tochris_b200/csrc/rc_host.h) by running the
 * per-thread kernel bodies of rc_pipeline.h in plain loops on the CPU.  It exists so that
 * the host logic (loaders, view setup, C-ABI) and the kernel arithmetic can be checked
 * against the oracle in the GPU-less container (`pytest -m "not gpu"`).  It is built into
 * tests/hostsim/_build/libref_cuda_sim.so, never into libref_cuda.so: the product has no
 * CPU path and fails at RC_Init without a CUDA device.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../tochris_b200/csrc/rc_host.h"
#include "../../tochris_b200/csrc/rc_pipeline.h"
#include "../../tochris_b200/csrc/rc_draw2d.h"

struct rc_gpu_s {
	int width, height;
	rc_world_t w;
	bool haveworld;
	int skybox = 0;
	int maxedges, maxsurfs;
	std::vector<int> sintable, intsintable;
	std::vector<unsigned long long> cachekeys;
	std::vector<int> cachevals;
	std::vector<uint8_t> cachepool;
	unsigned long long alloc[4];
	rc_stats_t stats;
	// last batch (debug reads)
	std::vector<rc_edge_t> edges; std::vector<int> nedges; int edgecap;
	std::vector<rc_surf_t> surfs; int surfcap; std::vector<int> nfaces;
	std::vector<rc_span_t> spans; int nspans;
	// alias model blobs
	std::vector<rc_aliasmodel_t> amodels;
	std::vector<int> stverts, atris;
	std::vector<uint8_t> poseverts, skinpixels;
	std::vector<float> anormdots;
	std::vector<uint8_t> spritepixels;
	int maxverts, maxtris;
};

extern "C" {

rc_gpu_t *rc_gpu_create (int device, int width, int height, size_t cache_bytes, int max_views, char *err, size_t errlen)
{
	(void)device; (void)max_views; (void)err; (void)errlen;
	rc_gpu_t *g = new rc_gpu_s ();
	g->width = width; g->height = height; g->haveworld = false;
	g->sintable.resize (RC_SIN_BUFFER_SIZE); g->intsintable.resize (RC_SIN_BUFFER_SIZE);
	RC_BuildTurbTables (g->sintable.data (), g->intsintable.data (), RC_SIN_BUFFER_SIZE);
	g->cachekeys.assign (1 << 16, RC_CACHE_EMPTY);
	g->cachevals.assign (1 << 16, 0);
	g->cachepool.resize (cache_bytes ? cache_bytes : (size_t)64 << 20);
	memset (g->alloc, 0, sizeof(g->alloc));
	memset (&g->stats, 0, sizeof(g->stats));
	return g;
}

void rc_gpu_destroy (rc_gpu_t *g) { delete g; }

int rc_gpu_upload_world (rc_gpu_t *g, const rc_hostworld_t *hw, int maxedges, int maxsurfs, char *err, size_t errlen)
{
	(void)err; (void)errlen;
	g->w = hw->w;
	g->w.sintable = g->sintable.data ();
	g->w.intsintable = g->intsintable.data ();
	g->w.drawskybox = g->skybox;
	g->haveworld = true;
	g->maxedges = maxedges; g->maxsurfs = maxsurfs;
	rc_gpu_flush_caches (g);
	return RC_OK;
}

int rc_gpu_upload_alias (rc_gpu_t *g, rc_hostalias_t **models, int n, const float *anorm_dots, char *err, size_t errlen)
{
	(void)err; (void)errlen;
	g->amodels.clear (); g->stverts.clear (); g->atris.clear (); g->poseverts.clear (); g->skinpixels.clear ();
	g->anormdots.assign (anorm_dots, anorm_dots + 16 * 256);
	g->maxverts = g->maxtris = 0;
	for (int i = 0; i < n; i++)
	{
		rc_hostalias_t *a = models[i];
		rc_aliasmodel_t m;
		m.numverts = a->numverts; m.numtris = a->numtris; m.skinwidth = a->skinwidth; m.skinheight = a->skinheight;
		m.stverts_ofs = (int)g->stverts.size () / 4; m.tris_ofs = (int)g->atris.size () / 4;
		m.poses_ofs = (int)g->poseverts.size () / 4; m.skins_ofs = (int)g->skinpixels.size ();
		g->stverts.insert (g->stverts.end (), a->stverts, a->stverts + 4 * a->numverts);
		g->atris.insert (g->atris.end (), a->tris, a->tris + 4 * a->numtris);
		g->poseverts.insert (g->poseverts.end (), a->poseverts, a->poseverts + 4 * (size_t)a->numposes * a->numverts);
		m.skins_len = a->skinimagelen; m.pad[0] = m.pad[1] = m.pad[2] = 0;
		g->skinpixels.insert (g->skinpixels.end (), a->skinpixels, a->skinpixels + (size_t)a->skinimagelen);
		g->amodels.push_back (m);
		a->device_index = i;
		if (a->numverts > g->maxverts) g->maxverts = a->numverts;
		if (a->numtris > g->maxtris) g->maxtris = a->numtris;
	}
	return RC_OK;
}

int rc_gpu_upload_sprites (rc_gpu_t *g, rc_hostsprite_t **models, int n, char *err, size_t errlen)
{
	(void)err; (void)errlen;
	g->spritepixels.clear ();
	for (int i = 0; i < n; i++)
	{
		models[i]->device_ofs = (int)g->spritepixels.size ();
		g->spritepixels.insert (g->spritepixels.end (), models[i]->pixels, models[i]->pixels + models[i]->pixelbytes);
	}
	return RC_OK;
}

void rc_gpu_flush_caches (rc_gpu_t *g)
{
	std::fill (g->cachekeys.begin (), g->cachekeys.end (), RC_CACHE_EMPTY);
	g->alloc[1] = RC_SKYBOX_BYTES;		/* the pool starts after the sky-box pictures */
}

/* large batches in parts: a device matter (the host's set-up overlapping the device's work); the simulator never splits */
int rc_gpu_can_split (const rc_gpu_t *g) { (void)g; return 0; }
int rc_gpu_render_begin (rc_gpu_t *g, void *fb_dev, void *stream) { (void)g; (void)fb_dev; (void)stream; return RC_OK; }
int rc_gpu_render_part (rc_gpu_t *g, const rc_batch_t *part, void *fb_part, void *z_part, void *stream, int first_part, char *err, size_t errlen)
{ (void)g; (void)part; (void)fb_part; (void)z_part; (void)stream; (void)first_part; snprintf (err, errlen, "the simulator does not split batches"); return RC_ERR_ARG; }
int rc_gpu_render_finish (rc_gpu_t *g, void *fb_dev, void *z_dev, int n, void *stream, int *status, char *err, size_t errlen)
{ (void)g; (void)fb_dev; (void)z_dev; (void)n; (void)stream; (void)status; (void)err; (void)errlen; return RC_ERR_ARG; }
int rc_gpu_device_alloc (rc_gpu_t *g, size_t bytes, void **ptr) { (void)g; *ptr = malloc (bytes); return *ptr ? RC_OK : RC_ERR_NOMEM; }
void rc_gpu_device_free (rc_gpu_t *g, void *ptr) { (void)g; free (ptr); }
int rc_gpu_ipc_export (rc_gpu_t *g, void *ptr, unsigned char *handle) { (void)g; (void)ptr; (void)handle; return RC_ERR_CUDA; }
int rc_gpu_ipc_import (rc_gpu_t *g, const unsigned char *handle, void **ptr) { (void)g; (void)handle; (void)ptr; return RC_ERR_CUDA; }
void rc_gpu_ipc_close (rc_gpu_t *g, void *ptr) { (void)g; (void)ptr; }
void rc_gpu_set_mirror_groups (rc_gpu_t *g, int ngroups) { (void)g; (void)ngroups; }
void rc_gpu_set_mirror_deferred (rc_gpu_t *g, int on) { (void)g; (void)on; }
int rc_gpu_mirror_fence (rc_gpu_t *g, void *stream, char *err, size_t errlen) { (void)g; (void)stream; (void)err; (void)errlen; return RC_OK; }
void rc_gpu_set_mirror (rc_gpu_t *g, void *fb, void *z) { (void)g; (void)fb; (void)z; }

void rc_gpu_set_group_callback (rc_gpu_t *g, rc_group_fn fn, void *user, void *side_stream, int ngroups)
{
	(void)g; (void)fn; (void)user; (void)side_stream; (void)ngroups;	/* the simulator draws a batch in one piece */
}

int rc_gpu_set_skybox (rc_gpu_t *g, const uint8_t *pixels, char *err, size_t errlen)
{
	(void)err; (void)errlen;
	if (pixels) memcpy (g->cachepool.data (), pixels, RC_SKYBOX_BYTES);
	g->skybox = pixels != NULL;
	g->w.drawskybox = g->skybox;
	return RC_OK;
}

int rc_gpu_render (rc_gpu_t *g, const rc_batch_t *batch, void *fb_dev, void *z_dev, void *stream,
		   int *status, char *err, size_t errlen)
{
	(void)stream;
	const rc_view_t *views = batch->views;
	const int n = batch->n;
	const rc_world_t *w = &g->w;
	int stw[16] = {0};
	int &st = stw[0];
	rc_frame_t f;
	memset (&f, 0, sizeof(f));
	if (!g->haveworld) { snprintf (err, errlen, "no world"); return RC_ERR_NOWORLD; }
	std::vector<int16_t> zscratch;		/* a caller that does not want z (z_dev NULL, as the device path allows) */
	if (!z_dev) { zscratch.assign ((size_t)n * g->width * g->height, 0); z_dev = zscratch.data (); }
	f.nviews = n; f.width = g->width; f.height = g->height; f.views = views; f.dlights = batch->dlights; f.viewbase = 0;
	f.facecap = (w->numworldfaces < 65000 ? w->numworldfaces : 65000) + 6;
	std::vector<rc_facerec_t> facelist ((size_t)n * f.facecap);
	std::vector<int> nfaces (n, 0);
	std::vector<uint16_t> facepos ((size_t)n * w->numfaces, 0xFFFF);
	f.facelist = facelist.data (); f.nfaces = nfaces.data (); f.facepos = facepos.data ();
	f.edgecap = g->maxedges; f.maxedges = g->maxedges; f.maxsurfs = g->maxsurfs;
	g->edges.assign ((size_t)n * f.edgecap, rc_edge_t ());
	g->nedges.assign (n, 0);
	f.edges = g->edges.data (); f.nedges = g->nedges.data ();
	f.bsurfcap = batch->nbents ? 512 : 0;
	f.surfcap = f.facecap + 2 + f.bsurfcap;
	std::vector<int> leafkey ((size_t)n * w->numleafs, 0x7FFFFFFF), nbsurfs (n, 0);
	f.leafkey = leafkey.data (); f.nbsurfs = nbsurfs.data (); f.bents = batch->bents; f.nbents = batch->nbents;
	g->surfs.assign ((size_t)n * f.surfcap, rc_surf_t ());
	f.surfs = g->surfs.data ();
	f.spancap = n * g->height * 96;
	const int cw = (g->width + 7) / 8;
	f.segcap = f.spancap + n * g->height * (cw / 8 + 1);
	g->spans.assign (f.spancap, rc_span_t ());
	rc_cell_t nocell; nocell.si = -1; nocell.sub = 0;
	std::vector<rc_cell_t> cellspan ((size_t)n * g->height * cw, nocell);
	f.spans = g->spans.data (); f.cellspan = cellspan.data ();
	std::vector<int> segspan ((size_t)f.segcap, -1);
	std::vector<rc_bnd_t> bnd ((size_t)f.segcap * 8 + 16, rc_bnd_t ());
	f.segspan = segspan.data (); f.bnd = bnd.data ();
	g->alloc[0] = 0; g->alloc[2] = 0;
	f.alloc = g->alloc;
	f.cachekeys = g->cachekeys.data (); f.cachevals = g->cachevals.data (); f.cachemask = (int)g->cachekeys.size () - 1;
	f.cachepool = g->cachepool.data (); f.cachecap = (long long)g->cachepool.size ();
	f.jobcap = n * f.surfcap;
	std::vector<rc_cachejob_t> jobs (f.jobcap);
	f.jobs = jobs.data ();
	f.fb = (uint8_t *)fb_dev; f.zb = (int16_t *)z_dev;
	std::vector<uint8_t> warpbuf ((size_t)(batch->nwarp + 1) * RC_WARP_W * RC_WARP_H, 0);
	f.warpbuf = warpbuf.data ();
	f.status = stw;

	std::vector<int> facemark ((size_t)n * w->numfaces, 0x7F7F7F7F), viewleaf (n, 0);
	std::vector<rc_nodevisit_t> nodevisit ((size_t)n * w->numnodes, rc_nodevisit_t ());
	f.facemark = facemark.data (); f.viewleaf = viewleaf.data (); f.nodevisit = nodevisit.data ();
	for (int vi = 0; vi < n; vi++) rc_view_begin (w, &f, vi);
	{
		std::vector<uint16_t> masks (w->numnodes + 1);
		for (int vi = 0; vi < n; vi++)
		{
			for (int i = 0; i < w->numnodes; i++) masks[i] = rc_node_test (w, &f, vi, i);
			for (int x = 0; x < w->numnodes + w->numleafs; x++) rc_visit_walk (w, &f, vi, x, masks.data ());
			if (w->marksplit)
				for (int m = 0; m < w->nummarksurfaces; m++) rc_mark_face (w, &f, vi, m);
		}
	}
	for (int vi = 0; vi < n; vi++)
	{
		int skykey = 0x7FFFFFFF;
		for (int fi = 0; fi < w->numworldfaces; fi++) rc_list_face (w, &f, vi, fi, &f.nfaces[vi], &skykey);
		if (w->drawskybox && skykey != 0x7FFFFFFF)
		{
			const int listed = nfaces[vi] < f.facecap ? nfaces[vi] : f.facecap;
			for (int i = 0; i < 6; i++) rc_list_skybox (w, &f, vi, i, listed, skykey);
			nfaces[vi] += 6;
		}
	}
	for (int vi = 0; vi < n; vi++)
		for (int p = 0; p < nfaces[vi]; p++) rc_render_face (w, &f, vi, p, &f.nedges[vi]);
	for (int vi = 0; vi < n; vi++)
		for (int p = 0; p < nfaces[vi]; p++) rc_face_fixup (w, &f, vi, p);
	for (int bi = 0; bi < batch->nbents; bi++)
		for (int fi = 0; fi < batch->bents[bi].numfaces; fi++) rc_bmodel_face (w, &f, bi, fi);
	for (int vi = 0; vi < n; vi++)
		for (int line = 0; line < g->height; line++) rc_scan_line (w, &f, vi, line);
	for (int vi = 0; vi < n; vi++)
		for (int i = 0, si; (si = rc_surf_slot (&f, vi, i)) >= 0; i++) rc_surf_prep (w, &f, vi, si);
	for (int vi = 0; vi < n; vi++)
		for (int i = 0, si; (si = rc_surf_slot (&f, vi, i)) >= 0; i++) rc_surf_resolve (&f, vi, si);
	int njobs = (int)g->alloc[2];
	if (njobs > f.jobcap) njobs = f.jobcap;
	long long texels = 0;
	for (int j = 0; j < njobs; j++)
	{
		const rc_cachejob_t *job = &jobs[j];
		const rc_face_t *fa = &w->faces[job->face];
		unsigned bl[18 * 18];
		int size = ((fa->extents[0] >> 4) + 1) * ((fa->extents[1] >> 4) + 1);
		for (int i = 0; i < size; i++) bl[i] = rc_cache_luxel (w, &f, job, i);
		for (int y = 0; y < job->height; y++)
			for (int x = 0; x < job->width; x++)
				g->cachepool[job->dstofs + y * job->width + x] = rc_cache_texel (w, job, bl, x, y);
		texels += (long long)job->width * job->height;
	}
	int nspans = (int)(g->alloc[0] & 0xFFFFFFFFu);
	if (nspans > f.spancap) nspans = f.spancap;
	int nsegs = (int)(g->alloc[0] >> 32);
	if (nsegs > f.segcap) nsegs = f.segcap;
	for (int t = 0; t < nsegs; t++) rc_span_seg (w, &f, t);
	for (int row = 0; row < n * g->height; row++)
		for (int c = 0; c < cw; c++) rc_draw_cell (w, &f, row, c);
	for (int t = 0; t < nspans; t++) { rc_draw_mixed (w, &f, t, 0, nspans); rc_draw_mixed (w, &f, t, 1, nspans); }

	/* alias models, particles, z resolve (R_DrawEntitiesOnList, R_DrawParticles) */
	std::vector<rc_finalvert_t> finalverts ((size_t)batch->nfinalverts + 1);
	std::vector<unsigned long long> zres ((size_t)(batch->nzres ? batch->nzres : 1) * g->width * g->height, 0ull);
	f.alias = batch->alias; f.nalias = batch->nalias; f.amodels = g->amodels.data ();
	f.stverts = g->stverts.data (); f.atris = g->atris.data (); f.poseverts = g->poseverts.data ();
	f.skinpixels = g->skinpixels.data (); f.anormdots = g->anormdots.data (); f.colormaps = batch->colormaps;
	f.finalverts = finalverts.data (); f.particles = batch->particles; f.nparticles = batch->nparticles;
	f.zres = zres.data ();
	for (int ii = 0; ii < batch->nalias; ii++)
		for (int v = 0; v < g->amodels[batch->alias[ii].model].numverts; v++) rc_alias_vertex (&f, ii, v);
	for (int ii = 0; ii < batch->nalias; ii++)
		for (int t = 0; t < g->amodels[batch->alias[ii].model].numtris; t++) rc_alias_triangle (&f, ii, t);
	f.sprites = batch->sprites; f.nsprites = batch->nsprites; f.spritepixels = g->spritepixels.data ();
	for (int si = 0; si < batch->nsprites; si++)
		for (int k = 0; k <= g->height; k++) rc_sprite_span (&f, si, k);
	for (int pi = 0; pi < batch->nparticles; pi++) rc_particle (&f, pi);
	if (batch->nzres)
		for (int vi = 0; vi < n; vi++)
			if (views[vi].zres_index >= 0)
				for (int y = 0; y < g->height; y++)
					for (int x = 0; x < g->width; x++) rc_zresolve_pixel (&f, vi, x, y);
	for (int vi = 0; vi < n; vi++)
		if (views[vi].warp_index >= 0)
			for (int v = 0; v < views[vi].rd_h; v++)
				for (int u = 0; u < ((views[vi].rd_w + 3) & ~3); u++) rc_warp_pixel (w, &f, vi, u, v);
	g->nspans = (int)(g->alloc[0] & 0xFFFFFFFFu);
	if (g->nspans > f.spancap) g->nspans = f.spancap;
	g->edgecap = f.edgecap; g->surfcap = f.surfcap; g->nfaces = nfaces;
	memset (&g->stats, 0, sizeof(g->stats));
	g->stats.cache_jobs = njobs; g->stats.cache_texels = texels; g->stats.blocks = (long long)n * g->height * cw + nspans; g->stats.spans = g->nspans;
	for (int vi = 0; vi < n; vi++) { g->stats.edges += g->nedges[vi]; g->stats.faces += nfaces[vi]; }
	g->stats.alias_range_events = stw[1];
	if (status) *status = st;
	return RC_OK;
}

int rc_gpu_render_host (rc_gpu_t *g, const rc_batch_t *batch, uint8_t *fb_out, int16_t *z_out,
			int *status, char *err, size_t errlen)
{
	const int n = batch->n;
	size_t px = (size_t)g->width * g->height;
	memset (fb_out, 0, px * n);
	std::vector<int16_t> ztmp;
	if (!z_out) { ztmp.assign (px * n, 0); z_out = ztmp.data (); }
	else memset (z_out, 0, px * n * 2);
	return rc_gpu_render (g, batch, fb_out, z_out, NULL, status, err, errlen);
}

/* the simulator has nothing to overlap: the "asynchronous" call renders at once and the wait hands back its status */
static int sim_async_status[RC_ASYNC_DEPTH], sim_async_head, sim_async_n;
int rc_gpu_render_host_async (rc_gpu_t *g, const rc_batch_t *batch, uint8_t *fb_out, int16_t *z_out, char *err, size_t errlen)
{
	if (sim_async_n >= RC_ASYNC_DEPTH) { snprintf (err, errlen, "%d asynchronous calls already in flight: RC_WaitViews first", RC_ASYNC_DEPTH); return RC_ERR_ARG; }
	int st = 0;
	int r = rc_gpu_render_host (g, batch, fb_out, z_out, &st, err, errlen);
	if (r == RC_OK) { sim_async_status[(sim_async_head + sim_async_n) % RC_ASYNC_DEPTH] = st; sim_async_n++; }
	return r;
}
int rc_gpu_wait_host (rc_gpu_t *g, int *status, char *err, size_t errlen)
{
	if (sim_async_n <= 0) { snprintf (err, errlen, "no asynchronous call in flight"); return RC_ERR_ARG; }
	if (status) *status = sim_async_status[sim_async_head];
	sim_async_head = (sim_async_head + 1) % RC_ASYNC_DEPTH; sim_async_n--;
	return RC_OK;
}

int rc_gpu_present (rc_gpu_t *g, const uint32_t *pal32, int n, const void *fb, void *rgbx, void *stream, int host, char *err, size_t errlen)
{
	/* test infrastructure: the simulator only has host memory */
	if (!pal32) { snprintf (err, errlen, "the simulator keeps no palettes between calls"); return RC_ERR_ARG; }
	const size_t px = (size_t)g->width * g->height;
	const uint8_t *in = (const uint8_t *)fb; uint32_t *out = (uint32_t *)rgbx;
	for (int v = 0; v < n; v++)
		for (size_t p = 0; p < px; p++)
			out[v * px + p] = pal32[v * 256 + in[v * px + p]];
	return RC_OK;
}

/* test infrastructure: the 2D overlay as a host loop over the same per-pixel body (rc_draw2d.h) */
static std::vector<uint8_t> sim_d2blob;
static std::vector<rc_pic_t> sim_d2pics;
int rc_gpu_upload_draw2d (rc_gpu_t *g, const uint8_t *blob, size_t len, const void *pics, int npics, char *err, size_t errlen)
{
	sim_d2blob.assign (blob, blob + len);
	sim_d2pics.assign ((const rc_pic_t *)pics, (const rc_pic_t *)pics + npics);
	return RC_OK;
}
int rc_gpu_draw2d (rc_gpu_t *g, const rc_draw2d_t *cmds, int ncmds, const int *first, int n, void *fb, void *stream, int host, char *err, size_t errlen)
{
	rc_d2ctx_t x;
	x.blob = sim_d2blob.data (); x.pics = sim_d2pics.data (); x.width = g->width; x.height = g->height;
	for (int v = 0; v < n; v++)
		for (int ci = first[v]; ci < first[v + 1]; ci++)
		{
			const int cnt = rc_d2_count (&x, &cmds[ci]);
			for (int p = 0; p < cnt; p++)
				rc_d2_pixel (&x, &cmds[ci], p, (uint8_t *)fb + (size_t)v * g->width * g->height);
		}
	return RC_OK;
}

void rc_gpu_get_stats (rc_gpu_t *g, rc_stats_t *out) { *out = g->stats; }
void rc_gpu_set_profiling (rc_gpu_t *g, int on) { (void)g; (void)on; }

int rc_gpu_debug_read (rc_gpu_t *g, int kind, int view, void *dst, int cap)
{
	if (kind == 0)
	{
		int n = g->nedges[view] < cap ? g->nedges[view] : cap;
		memcpy (dst, &g->edges[(size_t)view * g->edgecap], sizeof(rc_edge_t) * n);
		return n;
	}
	if (kind == 1)
	{
		int n = g->nfaces[view] + 2 < cap ? g->nfaces[view] + 2 : cap;
		memcpy (dst, &g->surfs[(size_t)view * g->surfcap], sizeof(rc_surf_t) * n);
		return n;
	}
	if (kind == 2)
	{
		int n = g->nspans < cap ? g->nspans : cap;
		memcpy (dst, g->spans.data (), sizeof(rc_span_t) * n);
		return n;
	}
	return RC_ERR_ARG;
}

} // extern "C"
