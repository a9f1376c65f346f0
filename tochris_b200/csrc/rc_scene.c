/*
 * rc_scene.c -- the step in front of the path: the client's scene lists, the assembly of
 * refdef_t, and the timedemo loop, in batched form.
 *
 * The reference's client fills four global arrays per frame -- r_entities / r_particles / r_dlights / r_cshifts,
 * client/cl_view.c:67-78 -- through V_ClearScene / V_AddEntity / V_AddParticle / V_AddDlight / V_AddCshift
 * (cl_view.c:85-148: a full list drops what is added), and V_RenderView (cl_view.c:889-943) points cl.refdef at them and
 * calls R_RenderView at once, so one frame's lists are dead as soon as the next frame starts.  A batched renderer wants
 * MANY frames alive: rc_scene_t keeps the same four lists with the same limits for the frame being built and, at
 * RC_SceneEndFrame, copies them into an arena, so that a whole demo (or a roll-out) is one refdef_t array for
 * RC_RenderViews* / RC_TimeDemo.  `timedemo` (cl_demo.c:324-365) plays a demo as fast as it renders and prints
 * "<frames> frames <seconds> seconds <fps> fps", not counting the first frame; RC_TimeDemo does that over such an array,
 * batch by batch.
 *
 * Host code only; nothing here touches the GPU except through the public RC_* calls.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/ref_cuda.h"

#define RC_MAX_ENTITIES  256	/* client/ref.h:32-34 */
#define RC_MAX_PARTICLES 2048
#define RC_MAX_CSHIFTS   16

typedef struct {		/* one finished frame: offsets into the arena (the arena moves when it grows) */
	refdef_t rd;
	size_t ents, parts, dls, css, styles;
} rc_sceneframe_t;

struct rc_scene_s {
	/* the frame being built: cl_view.c:67-78 */
	int      r_numentities;  entity_t   r_entities[RC_MAX_ENTITIES];
	int      r_numparticles; particle_t r_particles[RC_MAX_PARTICLES];
	int      r_numdlights;   dlight_t   r_dlights[MAX_DLIGHTS];
	int      r_numcshifts;   cshift_t   r_cshifts[RC_MAX_CSHIFTS];
	/* finished frames */
	rc_sceneframe_t *frames; int nframes, framecap;
	unsigned char *arena; size_t arenalen, arenacap;
	size_t laststyles; int havestyles;
	refdef_t *out; int outcap;	/* the array RC_SceneFrames hands out */
};

rc_scene_t *RC_SceneNew (void)
{
	return (rc_scene_t *)calloc (1, sizeof(rc_scene_t));
}

void RC_SceneFree (rc_scene_t *s)
{
	if (!s) return;
	free (s->frames); free (s->arena); free (s->out);
	free (s);
}

/* V_ClearScene, cl_view.c:85-91 */
void RC_SceneClear (rc_scene_t *s)
{
	s->r_numcshifts = 0;
	s->r_numdlights = 0;
	s->r_numentities = 0;
	s->r_numparticles = 0;
}

/* forgets the finished frames as well (a new recording) */
void RC_SceneReset (rc_scene_t *s)
{
	RC_SceneClear (s);
	s->nframes = 0; s->arenalen = 0; s->havestyles = 0;
}

/* V_AddEntity, cl_view.c:98-104; 0 when the list is full (the reference drops the entity without a word) */
int RC_SceneAddEntity (rc_scene_t *s, const entity_t *ent)
{
	if (s->r_numentities >= RC_MAX_ENTITIES)
		return 0;
	s->r_entities[s->r_numentities] = *ent;
	s->r_numentities++;
	return 1;
}

/* V_AddParticle, cl_view.c:111-118 */
int RC_SceneAddParticle (rc_scene_t *s, const float *org, int color)
{
	if (s->r_numparticles >= RC_MAX_PARTICLES)
		return 0;
	s->r_particles[s->r_numparticles].color = color;
	memcpy (s->r_particles[s->r_numparticles].org, org, sizeof(vec3_t));
	s->r_numparticles++;
	return 1;
}

/* V_AddDlight, cl_view.c:125-132 */
int RC_SceneAddDlight (rc_scene_t *s, const float *org, float radius)
{
	if (s->r_numdlights >= MAX_DLIGHTS)
		return 0;
	s->r_dlights[s->r_numdlights].radius = radius;
	memcpy (s->r_dlights[s->r_numdlights].origin, org, sizeof(vec3_t));
	s->r_numdlights++;
	return 1;
}

/* V_AddCshift, cl_view.c:139-148 */
int RC_SceneAddCshift (rc_scene_t *s, int r, int g, int b, int percent)
{
	if (s->r_numcshifts >= RC_MAX_CSHIFTS)
		return 0;
	s->r_cshifts[s->r_numcshifts].destcolor[0] = r;
	s->r_cshifts[s->r_numcshifts].destcolor[1] = g;
	s->r_cshifts[s->r_numcshifts].destcolor[2] = b;
	s->r_cshifts[s->r_numcshifts].percent = percent;
	s->r_numcshifts++;
	return 1;
}

static int arena_put (rc_scene_t *s, const void *src, size_t len, size_t *ofs)
{
	const size_t at = (s->arenalen + 15) & ~(size_t)15;
	if (at + len > s->arenacap)
	{
		size_t cap = s->arenacap ? s->arenacap * 2 : (size_t)1 << 20;
		unsigned char *p;
		while (cap < at + len) cap *= 2;
		p = (unsigned char *)realloc (s->arena, cap);
		if (!p) return 0;
		s->arena = p; s->arenacap = cap;
	}
	if (len) memcpy (s->arena + at, src, len);
	s->arenalen = at + len;
	*ofs = at;
	return 1;
}

/* The second half of V_RenderView (cl_view.c:915-940): the frame's refdef_t points at the lists; `view` supplies what
 * V_CalcRefdef computed (rectangle, fov, origin, angles, flags) and cl.time / cl.oldtime / cl_lightstyle.  add[4] are the
 * cl_add_entities / cl_add_particles / cl_add_lights / cl_add_cshifts cvars (NULL: all 1): a 0 empties that list for the
 * renderer.  Returns the frame's index, or RC_ERR_NOMEM. */
int RC_SceneEndFrame (rc_scene_t *s, const refdef_t *view, const int *add)
{
	rc_sceneframe_t *f;
	if (!s || !view) return RC_ERR_ARG;
	if (s->nframes == s->framecap)
	{
		int cap = s->framecap ? s->framecap * 2 : 256;
		rc_sceneframe_t *p = (rc_sceneframe_t *)realloc (s->frames, sizeof(rc_sceneframe_t) * (size_t)cap);
		if (!p) return RC_ERR_NOMEM;
		s->frames = p; s->framecap = cap;
	}
	f = &s->frames[s->nframes];
	f->rd = *view;
	f->rd.num_entities = s->r_numentities;
	f->rd.num_particles = s->r_numparticles;
	f->rd.num_dlights = s->r_numdlights;
	f->rd.num_cshifts = s->r_numcshifts;
	if (add)
	{
		if (!add[0]) f->rd.num_entities = 0;
		if (!add[1]) f->rd.num_particles = 0;
		if (!add[2]) f->rd.num_dlights = 0;
		if (!add[3]) f->rd.num_cshifts = 0;
	}
	if (!arena_put (s, s->r_entities, sizeof(entity_t) * (size_t)f->rd.num_entities, &f->ents) ||
	    !arena_put (s, s->r_particles, sizeof(particle_t) * (size_t)f->rd.num_particles, &f->parts) ||
	    !arena_put (s, s->r_dlights, sizeof(dlight_t) * (size_t)f->rd.num_dlights, &f->dls) ||
	    !arena_put (s, s->r_cshifts, sizeof(cshift_t) * (size_t)f->rd.num_cshifts, &f->css))
		return RC_ERR_NOMEM;
	/* cl_lightstyle is one global table the client rewrites in place: a frame keeps the table as it was when the frame
	 * ended (shared with the previous frame while nothing changed) */
	f->styles = (size_t)-1;
	if (view->lightstyles)
	{
		if (s->havestyles && !memcmp (s->arena + s->laststyles, view->lightstyles, sizeof(lightstyle_t) * MAX_LIGHTSTYLES))
			f->styles = s->laststyles;
		else
		{
			if (!arena_put (s, view->lightstyles, sizeof(lightstyle_t) * MAX_LIGHTSTYLES, &f->styles))
				return RC_ERR_NOMEM;
			s->laststyles = f->styles; s->havestyles = 1;
		}
	}
	return s->nframes++;
}

/* the finished frames as one refdef_t array (valid until the next RC_SceneEndFrame / RC_SceneReset / RC_SceneFree) */
const refdef_t *RC_SceneFrames (rc_scene_t *s, int *n)
{
	int i;
	if (n) *n = s ? s->nframes : 0;
	if (!s || !s->nframes) return NULL;
	if (s->nframes > s->outcap)
	{
		free (s->out);
		s->out = (refdef_t *)malloc (sizeof(refdef_t) * (size_t)s->nframes);
		s->outcap = s->out ? s->nframes : 0;
		if (!s->out) { if (n) *n = 0; return NULL; }
	}
	for (i = 0; i < s->nframes; i++)
	{
		const rc_sceneframe_t *f = &s->frames[i];
		refdef_t *rd = &s->out[i];
		*rd = f->rd;
		rd->entities = (entity_t *)(s->arena + f->ents);
		rd->particles = (particle_t *)(s->arena + f->parts);
		rd->dlights = (dlight_t *)(s->arena + f->dls);
		rd->cshifts = (cshift_t *)(s->arena + f->css);
		rd->lightstyles = f->styles == (size_t)-1 ? NULL : (lightstyle_t *)(s->arena + f->styles);
	}
	return s->out;
}

/* CalcFov, cl_view.c:629-639 (the vertical field of view of a width x height rectangle).  The reference calls Sys_Error
 * on a bad fov_x; a library returns 0. */
float RC_CalcFov (float fov_x, float width, float height)
{
	float x;
	if (fov_x < 1 || fov_x > 179)
		return 0;
	x = width / tan (fov_x * (M_PI / 360.0));
	return atan (height / x) * (360.0 / M_PI);
}

/* timedemo, cl_demo.c:324-365, over a recorded frame array: frame 0 is rendered first and does not count ("the first
 * frame didn't count": it pays for what loading left undone -- here the surface cache), then the clock runs over frames
 * 1 .. n-1, rendered `batch` views per call.  fb_out (n * height * width bytes, page-locked recommended) receives every
 * frame through the pipelined host path (RC_RenderViewsAsync, three calls outstanding); fb_out NULL keeps the frames on
 * the device, as a demo whose frames nobody looks at does.  out->text is the line CL_FinishTimeDemo prints. */
int RC_TimeDemo (const refdef_t *frames, int n, int batch, uint8_t *fb_out, rc_timedemo_t *out)
{
	struct timespec t0, t1;
	int r, st = 0, status = 0, first, outstanding = 0;
	double secs;
	size_t px;
	void *fbdev = NULL;
	rc_config_t cfg;
	if (!frames || n < 1 || !out) return RC_ERR_ARG;
	if (batch < 1) batch = 64;
	RC_GetConfig (&cfg);
	px = (size_t)cfg.width * cfg.height;
	if (!px) return RC_ERR_ARG;
	memset (out, 0, sizeof(*out));
	if (!fb_out && RC_DeviceAlloc (px * (size_t)batch, &fbdev) != RC_OK)
		return RC_ERR_NOMEM;
	r = fb_out ? RC_RenderViews (frames, 1, fb_out, NULL, &st) : RC_RenderViewsDevice (frames, 1, fbdev, NULL, NULL, &st);
	status |= st;
	clock_gettime (CLOCK_MONOTONIC, &t0);
	for (first = 1; r == RC_OK && first < n; first += batch)
	{
		const int cnt = n - first < batch ? n - first : batch;
		if (fb_out)
		{
			if (outstanding == RC_ASYNC_CALLS)
			{
				r = RC_WaitViews (&st); status |= st; outstanding--;
				if (r != RC_OK) break;
			}
			r = RC_RenderViewsAsync (frames + first, cnt, fb_out + px * (size_t)first, NULL);
			if (r == RC_OK) outstanding++;
		}
		else
		{
			r = RC_RenderViewsDevice (frames + first, cnt, fbdev, NULL, NULL, &st);
			status |= st;
		}
	}
	while (outstanding > 0)
	{
		const int r2 = RC_WaitViews (&st);
		status |= st; outstanding--;
		if (r == RC_OK) r = r2;
	}
	clock_gettime (CLOCK_MONOTONIC, &t1);
	if (fbdev) RC_DeviceFree (fbdev);
	if (r == RC_OK && fb_out && (status & (RC_STAT_SPAN_POOL | RC_STAT_CACHE_POOL | RC_STAT_AET_OVERFLOW)))
	{
		/* a pool ran out in a pipelined call (include/ref_cuda.h: such a batch is rendered again by the blocking call,
		 * which grows the pools): the whole run again, blocking, untimed parts included in the clock */
		status = 0;
		for (first = 1; r == RC_OK && first < n; first += batch)
		{
			const int cnt = n - first < batch ? n - first : batch;
			r = RC_RenderViews (frames + first, cnt, fb_out + px * (size_t)first, NULL, &st);
			status |= st;
		}
		clock_gettime (CLOCK_MONOTONIC, &t1);
	}
	secs = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
	out->frames = n - 1;			/* the first frame didn't count */
	out->seconds = secs;
	if (!secs) secs = 1;
	out->fps = out->frames / secs;
	out->status = status;
	snprintf (out->text, sizeof(out->text), "%i frames %5.1f seconds %5.1f fps\n", out->frames, (float)out->seconds ? (float)out->seconds : 1.0f, (float)out->fps);
	return r;
}
