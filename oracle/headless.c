/*
 * oracle/headless.c -- TEST INFRASTRUCTURE ONLY (never linked into ref_cuda).
 *
 * Headless platform layer + C harness that lets the UNMODIFIED reference
 * renderer (/root/reference/ref_soft/\*.c, C path, id386 undefined) run as a
 * shared library, oracle/_ref/librefsoft.so.  It plays the part that
 * win32/vid_win.c + win32/sys_win.c + common/host.c play in the real engine:
 *   - owns `vid`, `host_basepal`, `host_colormap`, `d_pzbuffer`'s storage and
 *     the surface cache memory  (pattern: win32/vid_win.c:2150-2153, :1674)
 *   - provides the Sys_/Con_/Host_ imports the renderer links against
 *   - does the Host_Init subset the renderer needs (common/host.c:787-870)
 * plus stage-dump hooks: D_DrawSurfaces / R_ScanEdges are renamed to *_real
 * at compile time (see Makefile) so this file can snapshot edges, surfaces
 * and spans before handing over to the reference's own functions.
 *
 * Deliberate differences from win32/vid_win.c, both documented in DESIGN.md:
 *   - no remap of colormap entries equal to 0 (vid_win.c:2155-2181 is a GDI
 *     palette quirk, not renderer behaviour)
 *   - vid.aspect = ((float)h/(float)w) * (320.0/240.0)   (vid_win.c:1351)
 */
#include "quakedef.h"
#include "r_local.h"
#include "d_local.h"
#include <stdarg.h>
#include <setjmp.h>
#include <time.h>
#include <errno.h>

/* ---- globals the renderer imports ------------------------------------ */
viddef_t	vid;
byte		*host_basepal;
byte		*host_colormap;
qboolean	host_initialized;
quakeparms_t	host_parms;
sizebuf_t	net_message;
vec3_t		r_pright, r_pup, r_ppn;
unsigned short	d_8to16table[256];
unsigned	d_8to24table[256];
qboolean	isDedicated;

static jmp_buf	orc_abort;
static int	orc_abort_armed;
static char	orc_errmsg[1024];
static int	orc_verbose;

const char *orc_last_error (void) { return orc_errmsg; }

void Sys_Error (char *error, ...)
{
	va_list ap;
	va_start (ap, error);
	vsnprintf (orc_errmsg, sizeof(orc_errmsg), error, ap);
	va_end (ap);
	fprintf (stderr, "oracle Sys_Error: %s\n", orc_errmsg);
	if (orc_abort_armed)
		longjmp (orc_abort, 1);
	abort ();
}

void Host_EndGame (char *message, ...)
{
	va_list ap;
	va_start (ap, message);
	vsnprintf (orc_errmsg, sizeof(orc_errmsg), message, ap);
	va_end (ap);
	fprintf (stderr, "oracle Host_EndGame: %s\n", orc_errmsg);
	if (orc_abort_armed)
		longjmp (orc_abort, 1);
	abort ();
}

void Con_Printf (char *fmt, ...)
{
	va_list ap;
	if (!orc_verbose) return;
	va_start (ap, fmt); vfprintf (stderr, fmt, ap); va_end (ap);
}
void Con_DPrintf (char *fmt, ...)
{
	va_list ap;
	if (orc_verbose < 2) return;
	va_start (ap, fmt); vfprintf (stderr, fmt, ap); va_end (ap);
}
void Con_SafePrintf (char *fmt, ...) { (void)fmt; }
void Sys_Printf (char *fmt, ...) { (void)fmt; }
void SV_BroadcastPrintf (char *fmt, ...) { (void)fmt; }
void S_ExtraUpdate (void) {}
/* the palette R_UpdatePalette hands to the video driver (r_main.c:874), kept for orc_end_frame */
static int orc_have_draw;
unsigned char orc_palette[768];
int orc_palette_uploads;
void VID_ShiftPalette (unsigned char *p) { memcpy (orc_palette, p, 768); orc_palette_uploads++; }
void VID_SetPalette (unsigned char *p) { (void)p; }
void VID_Update (vrect_t *r) { (void)r; }
void Sys_mkdir (char *path) { (void)path; }
void Sys_Quit (void) { exit (0); }
/* Com_ClientMaxclients/Com_ServerState come from common/common.c (state 0:
 * no local server => r_misc.c:420 forces r_draworder to 0 every frame; orc_set_cvar ("r_draworder") switches the state). */

double Sys_FloatTime (void)
{
	struct timespec ts;
	clock_gettime (CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

/* ---- file IO (own implementation; same contract as common/sys.h) ------ */
#define ORC_MAXFILES 16
static FILE *orc_files[ORC_MAXFILES];

static int orc_slot (void)
{
	int i;
	for (i = 1; i < ORC_MAXFILES; i++)
		if (!orc_files[i]) return i;
	Sys_Error ("oracle: out of file handles");
	return -1;
}
int Sys_FileOpenRead (char *path, int *hndl)
{
	long len;
	int s = orc_slot ();
	FILE *f = fopen (path, "rb");
	if (!f) { *hndl = -1; return -1; }
	fseek (f, 0, SEEK_END); len = ftell (f); fseek (f, 0, SEEK_SET);
	orc_files[s] = f; *hndl = s;
	return (int)len;
}
int Sys_FileOpenWrite (char *path)
{
	int s = orc_slot ();
	FILE *f = fopen (path, "wb");
	if (!f) Sys_Error ("oracle: cannot write %s: %s", path, strerror (errno));
	orc_files[s] = f;
	return s;
}
void Sys_FileClose (int h) { if (orc_files[h]) fclose (orc_files[h]); orc_files[h] = NULL; }
void Sys_FileSeek (int h, int pos) { fseek (orc_files[h], pos, SEEK_SET); }
int Sys_FileRead (int h, void *dst, int n) { return (int)fread (dst, 1, n, orc_files[h]); }
int Sys_FileWrite (int h, void *src, int n) { return (int)fwrite (src, 1, n, orc_files[h]); }
int Sys_FileTime (char *path)
{
	FILE *f = fopen (path, "rb");
	if (!f) return -1;
	fclose (f);
	return 1;
}

/* ---- init --------------------------------------------------------------- */
static void	*orc_surfcache;
static short	*orc_zbuf;
static byte	*orc_fb;
static char	orc_basedir[512];
static char	*orc_argv[4];
static model_t	*orc_world;

extern void R_InitTextures (void);

/* width/height may exceed the reference's compile-time MAXWIDTH/MAXHEIGHT only
 * in a build whose Makefile raised them (constants-only patch, see Makefile). */
int orc_init (const char *basedir, int width, int height, int heap_mb, int verbose)
{
	int		cachesize;
	static int	done;

	orc_verbose = verbose;
	if (done) { strcpy (orc_errmsg, "orc_init called twice"); return -1; }
	if (width > MAXWIDTH || height > MAXHEIGHT)
	{
		snprintf (orc_errmsg, sizeof(orc_errmsg), "%dx%d exceeds MAXWIDTH/MAXHEIGHT %dx%d",
			width, height, MAXWIDTH, MAXHEIGHT);
		return -1;
	}
	orc_abort_armed = 1;
	if (setjmp (orc_abort)) { orc_abort_armed = 0; return -1; }

	snprintf (orc_basedir, sizeof(orc_basedir), "%s", basedir);
	orc_argv[0] = "oracle";
	memset (&host_parms, 0, sizeof(host_parms));
	host_parms.basedir = orc_basedir;
	host_parms.argc = 1;
	host_parms.argv = orc_argv;
	host_parms.memsize = heap_mb * 1024 * 1024;
	host_parms.membase = malloc (host_parms.memsize);
	COM_InitArgv (host_parms.argc, host_parms.argv);

	/* common/host.c:806-841 subset, same order */
	Memory_Init (host_parms.membase, host_parms.memsize);
	Cbuf_Init ();
	Cmd_Init ();
	COM_Init (orc_basedir);
	Mod_Init ();
	host_basepal = (byte *)COM_LoadHunkFile ("gfx/palette.lmp");
	if (!host_basepal) Sys_Error ("Couldn't load gfx/palette.lmp");
	host_colormap = (byte *)COM_LoadHunkFile ("gfx/colormap.lmp");
	if (!host_colormap) Sys_Error ("Couldn't load gfx/colormap.lmp");

	/* VID_Init equivalent (win32/vid_win.c:2150-2153, :1351, :237-291, :1674) */
	memset (&vid, 0, sizeof(vid));
	vid.width = width;
	vid.height = height;
	vid.rowbytes = width;
	vid.aspect = ((float)height / (float)width) * (320.0 / 240.0);
	vid.numpages = 1;
	memcpy (vid.colormap, host_colormap, 256 * VID_GRADES);
	vid.fullbright = 256 - host_colormap[16384];
	orc_fb = calloc (1, (size_t)width * height + 64);
	vid.buffer = orc_fb;
	orc_zbuf = calloc (sizeof(short), (size_t)width * height + 64);
	d_pzbuffer = orc_zbuf;
	cachesize = D_SurfaceCacheForRes (width, height);
	orc_surfcache = calloc (1, cachesize);
	D_InitCaches (orc_surfcache, cachesize);

	/* host.c:815,839: W_LoadWadFile ("gfx.wad") + Draw_Init, when the 2D assets were generated */
	{
		char path[600];
		FILE *fw;
		snprintf (path, sizeof(path), "%s/id1/gfx.wad", orc_basedir);
		if ((fw = fopen (path, "rb")) != NULL)
		{
			fclose (fw);
			W_LoadWadFile ("gfx.wad");
			Draw_Init ();
			orc_have_draw = 1;
		}
	}

	R_Init ();

	Hunk_AllocName (0, "-HOST_HUNKLEVEL-");
	host_initialized = true;
	done = 1;
	orc_abort_armed = 0;
	return 0;
}

int orc_set_cvar (const char *name, float value)
{
	cvar_t *v = Cvar_FindVar ((char *)name);
	if (!v) return -1;
	Cvar_SetValue (v, value);		/* common/cvar.h:74 takes the cvar_t */
	/* r_draworder is only honoured while a local server runs (r_misc.c:420 resets it every frame otherwise): the shim
	 * says one does, as the engine would after `map` */
	if (!strcmp (name, "r_draworder"))
		Com_SetServerState (value ? 1 : 0);
	return 0;
}

/* string cvars (r_skyname: its change callback loads the sky box, r_sky.c:189-194) */
int orc_set_cvar_string (const char *name, const char *value)
{
	cvar_t *v = Cvar_FindVar ((char *)name);
	if (!v) return -1;
	Cvar_Set (v, (char *)value);
	return 0;
}

void *orc_model (const char *name)
{
	model_t *m;
	orc_abort_armed = 1;
	if (setjmp (orc_abort)) { orc_abort_armed = 0; return NULL; }
	m = Mod_ForName ((char *)name, true);
	orc_abort_armed = 0;
	return m;
}

/* Mod_ForName + R_NewMap, as client/cl_parse.c:271,294 do on a map change. */
void *orc_load_world (const char *name)
{
	orc_abort_armed = 1;
	if (setjmp (orc_abort)) { orc_abort_armed = 0; return NULL; }
	orc_world = Mod_ForName ((char *)name, true);
	R_NewMap (orc_world);
	orc_abort_armed = 0;
	return orc_world;
}

int orc_struct_sizes (int *out, int n)
{
	int v[] = { sizeof(refdef_t), sizeof(entity_t), sizeof(particle_t), sizeof(dlight_t),
		    sizeof(lightstyle_t), sizeof(edge_t), sizeof(surf_t), sizeof(espan_t),
		    sizeof(finalvert_t), MAXWIDTH, MAXHEIGHT, MAXSPANS };
	int i, c = sizeof(v)/sizeof(v[0]);
	for (i = 0; i < c && i < n; i++) out[i] = v[i];
	return c;
}

/* ---- stage dumps -------------------------------------------------------- */
typedef struct { int u, u_step, v, v2; unsigned short surfs[2]; float nearzi; int owner; } orc_edge_t;
typedef struct { int key, flags, insubmodel, face; float nearzi, d_ziorigin, d_zistepu, d_zistepv; int nspans, entity; } orc_surf_t;
typedef struct { int u, v, count, surf; } orc_span_t;

static orc_edge_t *dump_edges; static int dump_edges_cap, dump_edges_n;
static orc_surf_t *dump_surfs; static int dump_surfs_cap, dump_surfs_n;
static orc_span_t *dump_spans; static int dump_spans_cap, dump_spans_n;
static int dump_enabled;
int orc_counters[8];	/* 0: r_outofsurfaces 1: r_outofedges 2: nsurfs 3: nedges 4: c_surf builds */

void orc_dump_enable (int on, int maxedges, int maxsurfs, int maxspans)
{
	dump_enabled = on;
	if (!on) return;
	if (maxedges > dump_edges_cap) { dump_edges = realloc (dump_edges, sizeof(orc_edge_t)*maxedges); dump_edges_cap = maxedges; }
	if (maxsurfs > dump_surfs_cap) { dump_surfs = realloc (dump_surfs, sizeof(orc_surf_t)*maxsurfs); dump_surfs_cap = maxsurfs; }
	if (maxspans > dump_spans_cap) { dump_spans = realloc (dump_spans, sizeof(orc_span_t)*maxspans); dump_spans_cap = maxspans; }
}
int orc_dump_get (void **edges, int *nedges, void **surfs, int *nsurfs, void **spans, int *nspans)
{
	*edges = dump_edges; *nedges = dump_edges_n;
	*surfs = dump_surfs; *nsurfs = dump_surfs_n;
	*spans = dump_spans; *nspans = dump_spans_n;
	return 0;
}

extern void R_ScanEdges_real (void);
extern void D_DrawSurfaces_real (void);

/* r_edge.c is compiled with -DR_ScanEdges=R_ScanEdges_real; callers
 * (r_main.c:950) reach this hook first. */
void R_ScanEdges (void)
{
	if (dump_enabled)
	{
		int v, n = 0;
		edge_t *e;
		/* walk newedges[] in list order so the dump preserves the reference's
		 * per-scanline insertion order (r_rast.c:363-379) */
		for (v = r_oldrefdef.vrect.y; v < r_oldrefdef.vrectbottom; v++)
			for (e = newedges[v]; e; e = e->next)
			{
				if (n >= dump_edges_cap) break;
				dump_edges[n].u = e->u; dump_edges[n].u_step = e->u_step;
				dump_edges[n].v = v; dump_edges[n].v2 = -1;
				dump_edges[n].surfs[0] = e->surfs[0]; dump_edges[n].surfs[1] = e->surfs[1];
				dump_edges[n].nearzi = e->nearzi;
				dump_edges[n].owner = (int)(e - r_edges);	/* emission index */
				n++;
			}
		dump_edges_n = n;
		/* v2 from removeedges[] */
		for (v = r_oldrefdef.vrect.y; v < r_oldrefdef.vrectbottom; v++)
			for (e = removeedges[v]; e; e = e->nextremove)
			{
				int i, idx = (int)(e - r_edges);
				for (i = 0; i < n; i++)
					if (dump_edges[i].owner == idx) { dump_edges[i].v2 = v; break; }
			}
		dump_surfs_n = 0;
		dump_spans_n = 0;
	}
	orc_counters[0] = r_outofsurfaces;
	orc_counters[1] = r_outofedges;
	orc_counters[2] = (int)(surface_p - surfaces);
	orc_counters[3] = (int)(edge_p - r_edges);
	R_ScanEdges_real ();
}

/* d_edge.c is compiled with -DD_DrawSurfaces=D_DrawSurfaces_real. Called from
 * r_edge.c:675 (mid-frame span-pool flush) and :705 (end of frame); spans are
 * appended across flushes. */
void D_DrawSurfaces (void)
{
	if (dump_enabled)
	{
		surf_t *s;
		espan_t *sp;
		int i, ns = (int)(surface_p - surfaces);
		if (ns > dump_surfs_cap) ns = dump_surfs_cap;
		for (i = 1; i < ns; i++)
		{
			s = &surfaces[i];
			dump_surfs[i].key = s->key; dump_surfs[i].flags = s->flags;
			dump_surfs[i].insubmodel = s->insubmodel;
			dump_surfs[i].face = (i == 1 || !s->data) ? -1 :
				(int)((msurface_t *)s->data - ((s->insubmodel && s->entity) ? s->entity->model->surfaces : r_worldmodel->surfaces));
			dump_surfs[i].nearzi = s->nearzi;
			dump_surfs[i].d_ziorigin = s->d_ziorigin;
			dump_surfs[i].d_zistepu = s->d_zistepu;
			dump_surfs[i].d_zistepv = s->d_zistepv;
			dump_surfs[i].entity = (s->entity && r_refdef.entities && s->entity >= r_refdef.entities &&
				s->entity < r_refdef.entities + r_refdef.num_entities) ? (int)(s->entity - r_refdef.entities) : -1;
			for (sp = s->spans; sp; sp = sp->pnext)
			{
				if (dump_spans_n >= dump_spans_cap) break;
				dump_spans[dump_spans_n].u = sp->u; dump_spans[dump_spans_n].v = sp->v;
				dump_spans[dump_spans_n].count = sp->count; dump_spans[dump_spans_n].surf = i;
				dump_spans_n++;
			}
		}
		dump_surfs_n = ns;
	}
	D_DrawSurfaces_real ();
}

/* ---- rendering ---------------------------------------------------------- */
extern int c_surf;

/* One R_RenderView between R_BeginFrame/R_EndFrame-equivalent points
 * (client/cl_screen.c:1768-1807 minus the 2D overlay). fb_out: w*h bytes,
 * z_out: w*h shorts, either may be NULL. */
int orc_render (refdef_t *rd, byte *fb_out, short *z_out)
{
	orc_abort_armed = 1;
	if (setjmp (orc_abort)) { orc_abort_armed = 0; return -1; }
	c_surf = 0;
	R_RenderView (rd);
	orc_counters[4] = c_surf;
	if (fb_out) memcpy (fb_out, vid.buffer, (size_t)vid.width * vid.height);
	if (z_out) memcpy (z_out, d_pzbuffer, (size_t)vid.width * vid.height * sizeof(short));
	orc_abort_armed = 0;
	return 0;
}

/* Renders n views back to back (pattern of R_TimeRefresh_f, r_misc.c:42-58);
 * returns wall seconds through *seconds. Outputs optional, packed per view. */
int orc_render_batch (refdef_t *rds, int n, byte *fb_out, short *z_out, double *seconds)
{
	int i;
	double t0;
	size_t px = (size_t)vid.width * vid.height;
	orc_abort_armed = 1;
	if (setjmp (orc_abort)) { orc_abort_armed = 0; return -1; }
	t0 = Sys_FloatTime ();
	for (i = 0; i < n; i++)
	{
		R_RenderView (&rds[i]);
		if (fb_out) memcpy (fb_out + px * i, vid.buffer, px);
		if (z_out) memcpy (z_out + px * i, d_pzbuffer, px * sizeof(short));
	}
	if (seconds) *seconds = Sys_FloatTime () - t0;
	orc_abort_armed = 0;
	return 0;
}

/* 2D overlay: one reference Draw_* call per command on vid.buffer as the last orc_render left it (our own dispatcher,
 * the drawing is the reference's r_draw.c).  fb_out (may be NULL) receives vid.buffer afterwards. */
typedef struct { int op, x, y, w, h, arg; char pic[40]; int fromwad; const unsigned char *table; } orc_cmd2d_t;
int orc_draw2d (const orc_cmd2d_t *c, int n, byte *fb_out)
{
	int i;
	if (!orc_have_draw) { strcpy (orc_errmsg, "gfx.wad was not present at orc_init"); return -1; }
	orc_abort_armed = 1;
	if (setjmp (orc_abort)) { orc_abort_armed = 0; return -1; }
	for (i = 0; i < n; i++, c++)
	{
		mpic_t *pic = NULL;
		if (c->op >= 2 && c->op <= 4)
			pic = c->fromwad ? Draw_PicFromWad ((char *)c->pic) : Draw_CachePic ((char *)c->pic);
		switch (c->op)
		{
		case 1: Draw_Character (c->x, c->y, c->arg); break;
		case 2: Draw_Pic (c->x, c->y, pic); break;
		case 3: Draw_TransPic (c->x, c->y, pic); break;
		case 4: Draw_TransPicTranslate (c->x, c->y, pic, (byte *)c->table); break;
		case 5: Draw_ConsoleBackground (c->arg); break;
		case 6: Draw_TileClear (c->x, c->y, c->w, c->h); break;
		case 7: Draw_Fill (c->x, c->y, c->w, c->h, c->arg); break;
		case 8: Draw_FadeScreen (); break;
		}
	}
	if (fb_out) memcpy (fb_out, vid.buffer, (size_t)vid.width * vid.height);
	orc_abort_armed = 0;
	return 0;
}

/* R_EndFrame (r_main.c:961-980) after an orc_render of the same view (R_RenderView leaves it in r_refdef):
 * returns 1 and the palette the reference uploaded, 0 when it skipped the upload (no colour shifts, gamma unchanged). */
int orc_end_frame (unsigned char *pal_out)
{
	int before = orc_palette_uploads;
	orc_abort_armed = 1;
	if (setjmp (orc_abort)) { orc_abort_armed = 0; return -1; }
	R_EndFrame ();
	orc_abort_armed = 0;
	if (orc_palette_uploads == before)
		return 0;
	if (pal_out) memcpy (pal_out, orc_palette, 768);
	return 1;
}

/* the `screenshot` command (R_ScreenShot_f, r_misc.c:117-149) on vid.buffer as the last orc_render left it */
int orc_screenshot (void)
{
	extern void R_ScreenShot_f (void);
	orc_abort_armed = 1;
	if (setjmp (orc_abort)) { orc_abort_armed = 0; return -1; }
	R_ScreenShot_f ();
	orc_abort_armed = 0;
	return 0;
}

/* Mod_PointInLeaf (r_model.c:83-105) of the loaded world: the leaf's contents */
int orc_point_contents (const float *org)
{
	extern model_t *r_worldmodel;
	vec3_t p;
	if (!r_worldmodel) return 0;
	p[0] = org[0]; p[1] = org[1]; p[2] = org[2];
	return Mod_PointInLeaf (p, r_worldmodel)->contents;
}

void orc_clear_buffers (int fbval, int zval)
{
	size_t i, px = (size_t)vid.width * vid.height;
	memset (vid.buffer, fbval, px);
	for (i = 0; i < px; i++) d_pzbuffer[i] = (short)zval;
}

void orc_flush_caches (void) { D_FlushCaches (); }
