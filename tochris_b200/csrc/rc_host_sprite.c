/*
 * rc_host_sprite.c -- sprite models on the host: the SPR loader (Mod_LoadSpriteModel,
 * ref_soft/r_model.c:1581-1741; format in ref_soft/spritegn.h) and the per (view, entity) part of
 * R_DrawSprite (ref_soft/r_sprite.c:288-403) that runs before any vertex is touched: frame
 * selection (R_GetSpriteframe :236-281), the sprite's axes for its orientation type (libm sin/cos,
 * sqrt) and R_RotateSprite (:35-46).  Clipping, projection, edge scanning and the span drawer run
 * on the device (rc_sprite_span in rc_pipeline.h).  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "rc_host.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define SPR_VP_PARALLEL_UPRIGHT  0
#define SPR_FACING_UPRIGHT       1
#define SPR_VP_PARALLEL          2
#define SPR_ORIENTED             3
#define SPR_VP_PARALLEL_ORIENTED 4

#define FAIL(...) do { snprintf (err, errlen, __VA_ARGS__); RC_FreeHostSprite (s); return NULL; } while (0)

void RC_FreeHostSprite (rc_hostsprite_t *s)
{
	if (!s) return;
	free (s->frames); free (s->subframes); free (s->intervals); free (s->pixels);
	free (s);
}

static int rd_i (const unsigned char **p) { int v; memcpy (&v, *p, 4); *p += 4; return v; }
static float rd_f (const unsigned char **p) { float v; memcpy (&v, *p, 4); *p += 4; return v; }

rc_hostsprite_t *RC_LoadSpriteModel (const void *data, size_t len, char *err, size_t errlen)
{
	const unsigned char *p = (const unsigned char *)data, *end = p + len;
	rc_hostsprite_t *s = (rc_hostsprite_t *)calloc (1, sizeof(*s));
	int i, j, version, nsub = 0, subcap = 0, nint = 0, intcap = 0;
	size_t npix = 0, pixcap = 0;
	if (!s) { snprintf (err, errlen, "out of memory"); return NULL; }
	if (len < 36) FAIL ("sprite file too short");
	p += 4;						/* ident IDSP */
	version = rd_i (&p);
	if (version != 1) FAIL ("sprite has wrong version number (%i should be 1)", version);
	s->type = rd_i (&p);
	(void)rd_f (&p);				/* boundingradius */
	s->maxwidth = rd_i (&p);
	s->maxheight = rd_i (&p);
	s->numframes = rd_i (&p);
	s->beamlength = rd_f (&p);
	s->synctype = rd_i (&p);
	if (s->numframes < 1) FAIL ("Mod_LoadSpriteModel: Invalid # of frames: %d", s->numframes);
	s->frames = (rc_spriteframe_t *)calloc (s->numframes, sizeof(rc_spriteframe_t));
	if (!s->frames) FAIL ("out of memory");
	for (i = 0; i < s->numframes; i++)
	{
		int count = 1;
		if (p + 4 > end) FAIL ("sprite truncated");
		s->frames[i].type = rd_i (&p);
		s->frames[i].first = nsub;
		s->frames[i].intervals = -1;
		if (s->frames[i].type != 0)
		{
			/* Mod_LoadSpriteGroup, r_model.c:1619-1664 */
			if (p + 4 > end) FAIL ("sprite truncated");
			count = rd_i (&p);
			if (count < 1 || p + 4 * (size_t)count > end) FAIL ("bad sprite group");
			if (nint + count > intcap)
			{
				intcap = (nint + count) * 2;
				s->intervals = (float *)realloc (s->intervals, sizeof(float) * intcap);
				if (!s->intervals) FAIL ("out of memory");
			}
			s->frames[i].intervals = nint;
			for (j = 0; j < count; j++)
			{
				s->intervals[nint] = rd_f (&p);
				if (s->intervals[nint] <= 0.0) FAIL ("Mod_LoadSpriteGroup: interval<=0");
				nint++;
			}
		}
		s->frames[i].count = count;
		for (j = 0; j < count; j++)
		{
			/* Mod_LoadSpriteFrame, r_model.c:1581-1612 */
			rc_spritesub_t *sf;
			int origin[2], w, h;
			if (p + 16 > end) FAIL ("sprite truncated");
			origin[0] = rd_i (&p); origin[1] = rd_i (&p); w = rd_i (&p); h = rd_i (&p);
			if (w < 1 || h < 1 || w > 4096 || h > 4096 || p + (size_t)w * h > end) FAIL ("bad sprite frame");
			if (nsub >= subcap)
			{
				subcap = subcap ? subcap * 2 : 8;
				s->subframes = (rc_spritesub_t *)realloc (s->subframes, sizeof(rc_spritesub_t) * subcap);
				if (!s->subframes) FAIL ("out of memory");
			}
			sf = &s->subframes[nsub++];
			sf->width = w; sf->height = h;
			sf->up = origin[1]; sf->down = origin[1] - h; sf->left = origin[0]; sf->right = w + origin[0];
			sf->pixofs = (int)npix;
			if (npix + (size_t)w * h > pixcap)
			{
				pixcap = (npix + (size_t)w * h) * 2;
				s->pixels = (unsigned char *)realloc (s->pixels, pixcap);
				if (!s->pixels) FAIL ("out of memory");
			}
			memcpy (s->pixels + npix, p, (size_t)w * h);
			npix += (size_t)w * h;
			p += (size_t)w * h;
		}
	}
	s->numsubframes = nsub;
	s->pixelbytes = npix;
	s->device_ofs = -1;
	return s;
}

static float vec_normalize (float *v)
{	/* VectorNormalize, common/mathlib.c: double sqrt of a float sum */
	float length = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
	length = sqrt (length);
	if (length)
	{
		float ilength = 1/length;
		v[0] *= ilength; v[1] *= ilength; v[2] *= ilength;
	}
	return length;
}

/* R_DrawSprite up to R_SetupAndDrawSprite; returns 0 when the reference draws nothing */
int RC_SpriteSetup (const rc_hostsprite_t *s, const entity_t *e, const rc_refdef_t *rd, const rc_view_t *v,
		    int synctype, rc_spriteinst_t *out)
{
	const rc_spritesub_t *sf;
	float vpn[3], vright[3], vup[3], tvec[3], entorigin[3], modelorg[3], dot, angle, sr, cr;
	int i, frame = e->frame;

	/* r_main.c:516-518 */
	for (i = 0; i < 3; i++) { entorigin[i] = e->origin[i]; modelorg[i] = v->origin[i] - entorigin[i]; }

	/* R_GetSpriteframe */
	if (frame >= s->numframes || frame < 0)
		frame = 0;
	if (s->frames[frame].type == 0)
		sf = &s->subframes[s->frames[frame].first];
	else
	{
		const float *pintervals = s->intervals + s->frames[frame].intervals;
		int numframes = s->frames[frame].count;
		float fullinterval = pintervals[numframes-1], time, targettime;
		if (synctype == 1 /* ST_RAND */)
			time = rd->time + e->syncbase;
		else
			time = rd->time;
		targettime = time - ((int)(time / fullinterval)) * fullinterval;
		for (i = 0; i < (numframes-1); i++)
			if (pintervals[i] > targettime)
				break;
		sf = &s->subframes[s->frames[frame].first + i];
	}

	if (s->type == SPR_FACING_UPRIGHT)
	{
		tvec[0] = -modelorg[0]; tvec[1] = -modelorg[1]; tvec[2] = -modelorg[2];
		vec_normalize (tvec);
		dot = tvec[2];
		if ((dot > 0.999848) || (dot < -0.999848))
			return 0;
		vup[0] = 0; vup[1] = 0; vup[2] = 1;
		vright[0] = tvec[1]; vright[1] = -tvec[0]; vright[2] = 0;
		vec_normalize (vright);
		vpn[0] = -vright[1]; vpn[1] = vright[0]; vpn[2] = 0;
	}
	else if (s->type == SPR_VP_PARALLEL)
	{
		for (i = 0; i < 3; i++) { vup[i] = v->vup[i]; vright[i] = v->vright[i]; vpn[i] = v->vpn[i]; }
	}
	else if (s->type == SPR_VP_PARALLEL_UPRIGHT)
	{
		dot = v->vpn[2];
		if ((dot > 0.999848) || (dot < -0.999848))
			return 0;
		vup[0] = 0; vup[1] = 0; vup[2] = 1;
		vright[0] = v->vpn[1]; vright[1] = -v->vpn[0]; vright[2] = 0;
		vec_normalize (vright);
		vpn[0] = -vright[1]; vpn[1] = vright[0]; vpn[2] = 0;
	}
	else if (s->type == SPR_ORIENTED)
	{
		RC_AngleVectors (e->angles, vpn, vright, vup);
	}
	else if (s->type == SPR_VP_PARALLEL_ORIENTED)
	{
		angle = e->angles[2] * (M_PI*2 / 360);
		sr = sin (angle);
		cr = cos (angle);
		for (i = 0; i < 3; i++)
		{
			vpn[i] = v->vpn[i];
			vright[i] = v->vright[i] * cr + v->vup[i] * sr;
			vup[i] = v->vright[i] * -sr + v->vup[i] * cr;
		}
	}
	else
		return 0;		/* the reference: Sys_Error ("R_DrawSprite: Bad sprite type") */

	/* R_RotateSprite */
	if (s->beamlength != 0.0)
	{
		float vec[3];
		for (i = 0; i < 3; i++) vec[i] = vpn[i] * -s->beamlength;
		for (i = 0; i < 3; i++) { entorigin[i] = entorigin[i] + vec[i]; modelorg[i] = modelorg[i] - vec[i]; }
	}

	memset (out, 0, sizeof(*out));
	out->pixofs = sf->pixofs; out->width = sf->width; out->height = sf->height;
	out->up = sf->up; out->down = sf->down; out->left = sf->left; out->right = sf->right;
	for (i = 0; i < 3; i++)
	{
		out->vpn[i] = vpn[i]; out->vright[i] = vright[i]; out->vup[i] = vup[i];
		out->entorigin[i] = entorigin[i]; out->modelorg[i] = modelorg[i];
	}
	return 1;
}
