/*
 * rc_client.c -- the producers that fill the scene lists: the client's particle system and
 * dynamic lights (client/cl_effects.c) and the per-entity step of CL_AddEntities (client/cl_main.c:320-470: origin /
 * angle / frame interpolation, bobbing items, the effect and trail switches).  What the reference keeps in globals
 * (`particles[]`, `cl_dlights[]`, `cl.time`) lives in an rc_fx_t, so that many independent clients -- a roll-out per
 * environment, a demo per stream -- can each own one and feed one batched renderer through rc_scene_t.
 *
 * Data layout.  The reference threads its particles on two linked lists (active: newest first; free: LIFO).  Here the
 * live particles are DENSE arrays, oldest first, so the per-frame pass is a stable compaction plus a straight loop over
 * contiguous floats; list order is index order reversed.  The one thing that depends on WHICH slot of `particles[]` a
 * spawner gets is the `ramp` a spawner does not initialise (CL_EntityParticles, cl_effects.c:127-137: the next pass
 * reads it), so the free list survives as a stack of slot numbers with the ramp every slot last held: pushed in the
 * reference's kill order, popped by fx_new.
 *
 * Random numbers: the spawners call the C library's rand () in exactly the reference's order (the reference uses
 * nothing else), so a client that seeds srand () as the engine's process does gets the engine's particles.
 *
 * Arithmetic is written to evaluate as the reference's expressions do (float / double promotions as C gives them
 * there, no contraction: the library is built with -ffp-contract=off).  Host code only.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ref_cuda.h"

void RC_AngleVectors (const float *angles, float *forward, float *right, float *up);	/* rc_host_view.c */

#define FX_MIN_PARTICLES 512	/* ABSOLUTE_MIN_PARTICLES, cl_effects.c:23 */
#define FX_MAX_PARTICLES 2048	/* MAX_PARTICLES, client/ref.h:33 */
#define FX_NORMALS 162
#define FX_MAX_BEAMS 24		/* MAX_BEAMS, client.h:101 */
#define FX_MAX_TEMP_ENTITIES 128	/* client.h:286 */

enum { PK_STATIC, PK_GRAV, PK_SLOWGRAV, PK_FIRE, PK_EXPLODE, PK_EXPLODE2, PK_BLOB, PK_BLOB2, PK_RAILTRAIL };	/* ptype_t */

/* colour ramps of the fire / explosion kinds (cl_effects.c:31-33; the third one is zero padded there too) */
static const int fx_ramps[3][8] = {
	{ 0x6f, 0x6d, 0x6b, 0x69, 0x67, 0x65, 0x63, 0x61 },
	{ 0x6f, 0x6e, 0x6d, 0x6c, 0x6b, 0x6a, 0x68, 0x66 },
	{ 0x6d, 0x6b, 6, 5, 4, 3, 0, 0 },
};

struct rc_fx_s {
	int    cap, n;			/* cl_numparticles; live particles (dead ones leave in RC_FxAddParticles) */
	float *org, *vel;		/* [cap][3] */
	float *ramp, *die;
	int   *color, *kind;
	int   *slot;			/* which element of the reference's particles[] a live particle would be */
	float *slotramp;		/* the ramp that element last held */
	int   *freestack; int nfree;	/* the free list, its head last */
	rc_fxlight_t lights[MAX_DLIGHTS];
	float  avel[FX_NORMALS][3]; 	/* avelocities, cl_effects.c:92 */
	float  normals[FX_NORMALS][3]; int have_normals;
	int    tracercount;
	cshift_t cshifts[4], prev_cshifts[4];	/* cl.cshifts / cl.prev_cshifts: contents, damage, bonus, powerup (client.h:48-56) */
	cshift_t cshift_empty;			/* the `v_cshift` command's colour for air (cl_view.c:321, 401-407) */
	float  dmg_time, dmg_roll, dmg_pitch;	/* v_dmg_time / _roll / _pitch: the view kick of the last damage (cl_view.c:64) */
	float  oldz;				/* V_CalcRefdef's static: the eye height the stair smoothing has reached */
	struct { int entity; struct model_s *model; float endtime; float start[3], end[3]; } beams[FX_MAX_BEAMS];	/* cl_beams, cl_tent.c:26 */
};

static int fx_load_normals (rc_fx_t *x)
{
	Dl_info di;
	char path[1024];
	const char *tails[2] = { "/data/anorms.f32", "/../../../tochris_b200/data/anorms.f32" };	/* (the host-sim build lives in tests/hostsim/_build) */
	if (x->have_normals)
		return 1;
	if (!dladdr ((void *)RC_FxNew, &di) || !di.dli_fname)
		return 0;
	for (int k = 0; k < 2; k++)
	{
		char *slash;
		FILE *f;
		snprintf (path, sizeof(path) - 64, "%s", di.dli_fname);
		slash = strrchr (path, '/');
		if (slash) *slash = 0; else strcpy (path, ".");
		strcat (path, tails[k]);
		f = fopen (path, "rb");
		if (!f)
			continue;
		x->have_normals = fread (x->normals, sizeof(float), FX_NORMALS * 3, f) == FX_NORMALS * 3;
		fclose (f);
		if (x->have_normals)
			return 1;
	}
	return 0;
}

/* CL_InitParticles (-particles N, floor 512) + CL_ClearParticles + cleared lights */
rc_fx_t *RC_FxNew (int max_particles)
{
	rc_fx_t *x = (rc_fx_t *)calloc (1, sizeof(*x));
	if (!x)
		return NULL;
	x->cap = max_particles <= 0 ? FX_MAX_PARTICLES : max_particles;
	if (x->cap < FX_MIN_PARTICLES) x->cap = FX_MIN_PARTICLES;
	if (x->cap > FX_MAX_PARTICLES) x->cap = FX_MAX_PARTICLES;	/* particles[MAX_PARTICLES] */
	x->org = (float *)calloc ((size_t)x->cap * 3, sizeof(float)); x->vel = (float *)calloc ((size_t)x->cap * 3, sizeof(float));
	x->ramp = (float *)calloc (x->cap, sizeof(float)); x->die = (float *)calloc (x->cap, sizeof(float));
	x->color = (int *)calloc (x->cap, sizeof(int)); x->kind = (int *)calloc (x->cap, sizeof(int));
	x->slot = (int *)calloc (x->cap, sizeof(int)); x->slotramp = (float *)calloc (x->cap, sizeof(float));
	x->freestack = (int *)calloc (x->cap, sizeof(int));
	if (!x->org || !x->vel || !x->ramp || !x->die || !x->color || !x->kind || !x->slot || !x->slotramp || !x->freestack)
	{
		RC_FxFree (x);
		return NULL;
	}
	RC_FxClear (x);
	return x;
}

void RC_FxFree (rc_fx_t *x)
{
	if (!x) return;
	free (x->org); free (x->vel); free (x->ramp); free (x->die); free (x->color); free (x->kind); free (x->slot); free (x->slotramp); free (x->freestack);
	free (x);
}

/* CL_ClearParticles (cl_effects.c:147-157: every slot back on the free list in index order; the slots keep what they
 * held) and the light array of a new map (cl_main.c CL_ClearState) */
void RC_FxClear (rc_fx_t *x)
{
	for (int i = 0; i < x->n; i++)
		x->slotramp[x->slot[i]] = x->ramp[i];
	for (int i = 0; i < x->cap; i++)
		x->freestack[i] = x->cap - 1 - i;		/* element 0 at the head */
	x->nfree = x->cap;
	x->n = 0;
	memset (x->lights, 0, sizeof(x->lights));
	memset (x->beams, 0, sizeof(x->beams));			/* CL_ClearState, cl_main.c:81-82 */
	memset (x->cshifts, 0, sizeof(x->cshifts)); memset (x->prev_cshifts, 0, sizeof(x->prev_cshifts));
	x->cshift_empty.destcolor[0] = 130; x->cshift_empty.destcolor[1] = 80; x->cshift_empty.destcolor[2] = 50; x->cshift_empty.percent = 0;
}

int RC_FxParticleCount (const rc_fx_t *x) { return x->n; }

/* new_particle, cl_effects.c:69-81: -1 when the free list is empty */
static int fx_new (rc_fx_t *x)
{
	if (x->n >= x->cap)
		return -1;
	const int i = x->n++;
	x->slot[i] = x->freestack[--x->nfree];
	x->ramp[i] = x->slotramp[x->slot[i]];
	return i;
}

/* VectorNormalize, mathlib.c:324 */
static float fx_normalize (float *v)
{
	float length = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
	if (length)
	{
		length = sqrt (length);
		const float ilength = 1/length;
		v[0] *= ilength; v[1] *= ilength; v[2] *= ilength;
	}
	return length;
}

static void fx_jitter32 (rc_fx_t *x, int i, const float *org)
{
	/* the explosion family's cloud: position +-16, velocity +-256, one pair of draws per axis */
	for (int j = 0; j < 3; j++)
	{
		x->org[3*i+j] = org[j] + ((rand ()%32)-16);
		x->vel[3*i+j] = (rand ()%512)-256;
	}
}

/* the trail family (cl_effects.c:548-777): a particle every three units from start towards end */
static int fx_trail (rc_fx_t *x, int effect, const float *start_in, const float *end, int color, double time)
{
	float start[3] = { start_in[0], start_in[1], start_in[2] };
	float vec[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
	float len = fx_normalize (vec);
	const float dec = 3;
	int made = 0;
	vec[0] = vec[0]*dec; vec[1] = vec[1]*dec; vec[2] = vec[2]*dec;
	while (len > 0)
	{
		len -= dec;
		const int i = fx_new (x);
		if (i < 0)
			return made;
		made++;
		float *o = &x->org[3*i], *v = &x->vel[3*i];
		v[0] = v[1] = v[2] = 0;
		switch (effect)
		{
		case RC_FX_ROCKET_TRAIL:
		case RC_FX_GRENADE_TRAIL:
			x->die[i] = time + 2;
			x->ramp[i] = (rand ()&3) + (effect == RC_FX_GRENADE_TRAIL ? 2 : 0);
			x->color[i] = fx_ramps[2][(int)x->ramp[i]];
			x->kind[i] = PK_FIRE;
			for (int j = 0; j < 3; j++) o[j] = start[j] + ((rand ()%6)-3);
			break;
		case RC_FX_BLOOD_TRAIL:
		case RC_FX_SLIGHT_BLOOD_TRAIL:
			x->die[i] = time + 2;
			x->kind[i] = PK_GRAV;
			x->color[i] = 67 + (rand ()&3);
			for (int j = 0; j < 3; j++) o[j] = start[j] + ((rand ()%6)-3);
			if (effect == RC_FX_SLIGHT_BLOOD_TRAIL) len -= 3;
			break;
		case RC_FX_TRACER_TRAIL:
			x->die[i] = time + 0.5;
			x->kind[i] = PK_STATIC;
			x->color[i] = color + ((x->tracercount&4)<<1);
			x->tracercount++;
			o[0] = start[0]; o[1] = start[1]; o[2] = start[2];
			if (x->tracercount & 1) { v[0] = 30*vec[1]; v[1] = 30*-vec[0]; }
			else { v[0] = 30*-vec[1]; v[1] = 30*vec[0]; }
			break;
		default: /* RC_FX_VOOR_TRAIL */
			x->die[i] = time + 0.3;
			x->color[i] = 9*16 + 8 + (rand ()&3);
			x->kind[i] = PK_STATIC;
			for (int j = 0; j < 3; j++) o[j] = start[j] + ((rand ()&15)-8);
			break;
		}
		start[0] += vec[0]; start[1] += vec[1]; start[2] += vec[2];
	}
	return made;
}

/* CL_RailTrail, cl_effects.c:446-508: a straight core and a helix around it, 1.5 units a pair */
static int fx_rail (rc_fx_t *x, const float *start_in, const float *end, double time)
{
	float start[3] = { start_in[0], start_in[1], start_in[2] };
	float vec[3] = { end[0] - start[0], end[1] - start[1], end[2] - start[2] };
	float ang[2], right[3], roll = 0.0f;
	int made = 0;
	/* Vector2Angles, mathlib.c:505 (pitch, yaw) */
	if (!vec[0] && !vec[1])
	{
		ang[1] = 0;
		ang[0] = (vec[2] > 0) ? 90 : 270;
	}
	else
	{
		float y = atan2 (vec[1], vec[0]) * 180 / M_PI;
		if (y < 0) y += 360;
		float p = atan2 (vec[2], sqrt (vec[0]*vec[0]+vec[1]*vec[1])) * 180 / M_PI;
		if (p < 0) p += 360;
		ang[0] = p; ang[1] = y;
	}
	ang[0] *= (M_PI * 2 / 360);
	ang[1] *= (M_PI * 2 / 360);
	const float sp = sin (ang[0]), cp = cos (ang[0]), sy = sin (ang[1]), cy = cos (ang[1]);
	float len = fx_normalize (vec);
	vec[0] = vec[0]*1.5; vec[1] = vec[1]*1.5; vec[2] = vec[2]*1.5;
	right[0] = sy; right[1] = cy; right[2] = 0;
	while (len > 0)
	{
		int i = fx_new (x);
		if (i < 0) return made;
		made++;
		memcpy (&x->org[3*i], start, sizeof(start));
		x->vel[3*i] = x->vel[3*i+1] = x->vel[3*i+2] = 0;
		x->color[i] = (rand () & 3) + 225;
		x->kind[i] = PK_RAILTRAIL;
		x->die[i] = time + 1.0;
		i = fx_new (x);
		if (i < 0) return made;
		made++;
		for (int j = 0; j < 3; j++)
		{
			x->org[3*i+j] = start[j] + 4*right[j];
			x->vel[3*i+j] = right[j]*8;
		}
		x->color[i] = (rand () & 7) + 206;
		x->kind[i] = PK_RAILTRAIL;
		x->die[i] = time + 1.0;
		roll += 7.5 * (M_PI / 180.0);
		if (roll > 2*M_PI)
			roll -= 2*M_PI;
		const float sr = sin (roll), cr = cos (roll);
		right[0] = (cr * sy - sr * sp * cy);
		right[1] = -(sr * sp * sy + cr * cy);
		right[2] = -sr * cp;
		len -= 1.5;
		start[0] += vec[0]; start[1] += vec[1]; start[2] += vec[2];
	}
	return made;
}

/* the two splashes that fan particles out from a grid (CL_LavaSplash cl_effects.c:411-444, CL_TeleportSplash :517-546) */
static int fx_splash (rc_fx_t *x, int lava, const float *org, double time)
{
	const int step = lava ? 1 : 4, k0 = lava ? 0 : -24, k1 = lava ? 1 : 32;
	int made = 0;
	for (int i = -16; i < 16; i += step)
		for (int j = -16; j < 16; j += step)
			for (int k = k0; k < k1; k += step)
			{
				float dir[3], vel;
				const int p = fx_new (x);
				if (p < 0) return made;
				made++;
				x->kind[p] = PK_SLOWGRAV;
				if (lava)
				{
					x->die[p] = time + 2 + (rand ()&31) * 0.02;
					x->color[p] = 224 + (rand ()&7);
					dir[0] = j*8 + (rand ()&7);
					dir[1] = i*8 + (rand ()&7);
					dir[2] = 256;
					x->org[3*p+0] = org[0] + dir[0];
					x->org[3*p+1] = org[1] + dir[1];
					x->org[3*p+2] = org[2] + (rand ()&63);
				}
				else
				{
					x->die[p] = time + 0.2 + (rand ()&7) * 0.02;
					x->color[p] = 7 + (rand ()&7);
					dir[0] = j*8; dir[1] = i*8; dir[2] = k*8;
					x->org[3*p+0] = org[0] + i + (rand ()&3);
					x->org[3*p+1] = org[1] + j + (rand ()&3);
					x->org[3*p+2] = org[2] + k + (rand ()&3);
				}
				fx_normalize (dir);
				vel = 50 + (rand ()&63);
				x->vel[3*p+0] = dir[0]*vel; x->vel[3*p+1] = dir[1]*vel; x->vel[3*p+2] = dir[2]*vel;
			}
	return made;
}

/* CL_EntityParticles, cl_effects.c:98-138: 162 sparks on a sphere of radius 64 around the entity, each wobbling on its
 * own three angular velocities (drawn once per client) */
static int fx_entity_particles (rc_fx_t *x, const float *origin, double time)
{
	const float beamlength = 16;
	int made = 0;
	if (!fx_load_normals (x))
		return -1;
	if (!x->avel[0][0])
	{
		float *flat = &x->avel[0][0];
		for (int i = 0; i < FX_NORMALS * 3; i++)
			flat[i] = (rand ()&255) * 0.01;
	}
	for (int i = 0; i < FX_NORMALS; i++)
	{
		float angle, sp, sy, cp, cy, forward[3];
		angle = time * x->avel[i][0];
		sy = sin (angle); cy = cos (angle);
		angle = time * x->avel[i][1];
		sp = sin (angle); cp = cos (angle);
		/* (the roll terms of the reference's body feed nothing) */
		forward[0] = cp*cy; forward[1] = cp*sy; forward[2] = -sp;
		const int p = fx_new (x);
		if (p < 0) return made;
		made++;
		x->die[p] = time + 0.01;
		x->color[p] = 0x6f;
		x->kind[p] = PK_EXPLODE;		/* ramp: whatever the slot held (see the head of the file) */
		x->vel[3*p] = x->vel[3*p+1] = x->vel[3*p+2] = 0;
		for (int j = 0; j < 3; j++)
			x->org[3*p+j] = origin[j] + x->normals[i][j]*64 + forward[j]*beamlength;
	}
	return made;
}

/* One call for every spawner of cl_effects.c; `a`, `b`, `i0`, `i1` as the table in include/ref_cuda.h says.  Returns the
 * number of particles made (the reference stops silently when its free list is empty), -1 for a bad argument. */
int RC_FxEffect (rc_fx_t *x, int effect, const float *a, const float *b, int i0, int i1, double time)
{
	int made = 0;
	if (!x || !a)
		return -1;
	switch (effect)
	{
	case RC_FX_ENTITY_PARTICLES:
		return fx_entity_particles (x, a, time);
	case RC_FX_POINT:		/* a line of CL_ReadPointFile_f (cl_effects.c:159-204): i0 = the running count c */
	{
		const int p = fx_new (x);
		if (p < 0) return 0;
		x->die[p] = 99999; x->color[p] = (-i0)&15; x->kind[p] = PK_STATIC;
		x->vel[3*p] = x->vel[3*p+1] = x->vel[3*p+2] = 0;
		memcpy (&x->org[3*p], a, 3 * sizeof(float));
		return 1;
	}
	case RC_FX_EXPLOSION:		/* CL_ParticleExplosion :233-266 */
	case RC_FX_ROCKET_SPLASH:	/* CL_RocketSplash :372-405: the same cloud (direction and colour are not used there) */
		for (int i = 0; i < 1024; i++, made++)
		{
			const int p = fx_new (x);
			if (p < 0) return made;
			x->die[p] = time + 5;
			x->color[p] = fx_ramps[0][0];
			x->ramp[p] = rand ()&3;
			x->kind[p] = (i & 1) ? PK_EXPLODE : PK_EXPLODE2;
			fx_jitter32 (x, p, a);
		}
		return made;
	case RC_FX_EXPLOSION2:		/* CL_ParticleExplosion2 :274-296: i0 = colorStart, i1 = colorLength */
		if (i1 <= 0) return -1;
		for (int i = 0; i < 512; i++, made++)
		{
			const int p = fx_new (x);
			if (p < 0) return made;
			x->die[p] = time + 0.3;
			x->color[p] = i0 + (i % i1);
			x->kind[p] = PK_BLOB;
			fx_jitter32 (x, p, a);
		}
		return made;
	case RC_FX_BLOB_EXPLOSION:	/* CL_BlobExplosion :304-340 */
		for (int i = 0; i < 1024; i++, made++)
		{
			const int p = fx_new (x);
			if (p < 0) return made;
			x->die[p] = time + 1 + (rand ()&8)*0.05;
			if (i & 1) { x->kind[p] = PK_BLOB; x->color[p] = 66 + rand ()%6; }
			else { x->kind[p] = PK_BLOB2; x->color[p] = 150 + rand ()%6; }
			fx_jitter32 (x, p, a);
		}
		return made;
	case RC_FX_RUN_EFFECT:		/* CL_RunParticleEffect :348-369: b = dir, i0 = color, i1 = count */
		if (!b) return -1;
		for (int i = 0; i < i1; i++, made++)
		{
			const int p = fx_new (x);
			if (p < 0) return made;
			x->die[p] = time + 0.1*(rand ()%5);
			x->color[p] = (i0&~7) + (rand ()&7);
			x->kind[p] = PK_SLOWGRAV;
			for (int j = 0; j < 3; j++)
			{
				x->org[3*p+j] = a[j] + ((rand ()&15)-8);
				x->vel[3*p+j] = b[j]*15;
			}
		}
		return made;
	case RC_FX_LAVA_SPLASH:
		return fx_splash (x, 1, a, time);
	case RC_FX_TELEPORT_SPLASH:
		return fx_splash (x, 0, a, time);
	case RC_FX_RAIL_TRAIL:
		return b ? fx_rail (x, a, b, time) : -1;
	case RC_FX_ROCKET_TRAIL: case RC_FX_GRENADE_TRAIL: case RC_FX_BLOOD_TRAIL: case RC_FX_SLIGHT_BLOOD_TRAIL:
	case RC_FX_TRACER_TRAIL: case RC_FX_VOOR_TRAIL:
		return b ? fx_trail (x, effect, a, b, i0, time) : -1;
	}
	return -1;
}

/* CL_AddParticles, cl_effects.c:784-904: the dead leave (their slots go back on the free list in list order: newest
 * first), the living are handed to the scene -- newest first, as the reference walks its list -- and then moved on by
 * the frame's time.  Returns how many were handed over. */
int RC_FxAddParticles (rc_fx_t *x, rc_scene_t *scene, double time, double oldtime)
{
	if (!x || !x->n)
		return 0;
	const float frametime = time - oldtime;
	const float time3 = frametime * 15, time2 = frametime * 10, time1 = frametime * 5;
	const float grav = frametime * 800 * 0.05;
	const float dvel = 4*frametime;
	for (int i = x->n - 1; i >= 0; i--)
		if (x->die[i] < time)
		{
			x->slotramp[x->slot[i]] = x->ramp[i];
			x->freestack[x->nfree++] = x->slot[i];
		}
	int m = 0;
	for (int i = 0; i < x->n; i++)
		if (!(x->die[i] < time))
		{
			if (m != i)
			{
				memcpy (&x->org[3*m], &x->org[3*i], 3 * sizeof(float)); memcpy (&x->vel[3*m], &x->vel[3*i], 3 * sizeof(float));
				x->ramp[m] = x->ramp[i]; x->die[m] = x->die[i]; x->color[m] = x->color[i]; x->kind[m] = x->kind[i]; x->slot[m] = x->slot[i];
			}
			m++;
		}
	x->n = m;
	for (int i = m - 1; i >= 0; i--)
	{
		float *o = &x->org[3*i], *v = &x->vel[3*i];
		if (scene)
			RC_SceneAddParticle (scene, o, x->color[i]);
		o[0] += v[0]*frametime;
		o[1] += v[1]*frametime;
		o[2] += v[2]*frametime;
		switch (x->kind[i])
		{
		case PK_FIRE:
		case PK_EXPLODE:
		case PK_EXPLODE2:
		{
			const int k = x->kind[i];
			const float limit = k == PK_FIRE ? 6 : 8;
			x->ramp[i] += k == PK_FIRE ? time1 : (k == PK_EXPLODE ? time2 : time3);
			if (x->ramp[i] >= limit)
				x->die[i] = -1;
			else
			{
				/* (a clock running backwards indexes in front of the reference's tables: undefined there, slot 0 here) */
				const int r = (int)x->ramp[i];
				x->color[i] = fx_ramps[k == PK_FIRE ? 2 : (k == PK_EXPLODE ? 0 : 1)][r < 0 ? 0 : r];
			}
			if (k == PK_FIRE)
				v[2] += grav;
			else
			{
				if (k == PK_EXPLODE) for (int j = 0; j < 3; j++) v[j] += v[j]*dvel;
				else for (int j = 0; j < 3; j++) v[j] -= v[j]*frametime;
				v[2] -= grav;
			}
			break;
		}
		case PK_BLOB:
			for (int j = 0; j < 3; j++) v[j] += v[j]*dvel;
			v[2] -= grav;
			break;
		case PK_BLOB2:
			for (int j = 0; j < 2; j++) v[j] -= v[j]*dvel;
			v[2] -= grav;
			break;
		case PK_GRAV:
		case PK_SLOWGRAV:
			v[2] -= grav;
			break;
		default:	/* static, rail trail */
			break;
		}
	}
	return m;
}

/* debug / test access: copies out up to `max` live particles in LIST order (newest first) */
int RC_FxParticles (const rc_fx_t *x, int max, float *org3, float *vel3, float *ramp, float *die, int *color, int *kind)
{
	int k = 0;
	for (int i = x->n - 1; i >= 0 && k < max; i--, k++)
	{
		if (org3) memcpy (&org3[3*k], &x->org[3*i], 3 * sizeof(float));
		if (vel3) memcpy (&vel3[3*k], &x->vel[3*i], 3 * sizeof(float));
		if (ramp) ramp[k] = x->ramp[i];
		if (die) die[k] = x->die[i];
		if (color) color[k] = x->color[i];
		if (kind) kind[k] = x->kind[i];
	}
	return k;
}

/* CL_AllocDlight, cl_effects.c:915-951: the light with this key, else the first dead one, else number 0; zeroed */
rc_fxlight_t *RC_FxAllocDlight (rc_fx_t *x, int key, double time)
{
	rc_fxlight_t *dl = NULL;
	if (key)
		for (int i = 0; i < MAX_DLIGHTS && !dl; i++)
			if (x->lights[i].key == key)
				dl = &x->lights[i];
	for (int i = 0; i < MAX_DLIGHTS && !dl; i++)
		if (x->lights[i].die < time)
			dl = &x->lights[i];
	if (!dl)
		dl = &x->lights[0];
	memset (dl, 0, sizeof(*dl));
	dl->key = key;
	return dl;
}

rc_fxlight_t *RC_FxLights (rc_fx_t *x) { return x->lights; }

/* CL_AddDlights, cl_effects.c:959-968: live lights go to the scene and then decay.  Returns how many were handed over. */
int RC_FxAddDlights (rc_fx_t *x, rc_scene_t *scene, double time, double oldtime)
{
	const float dt = time - oldtime;
	int n = 0;
	for (int i = 0; i < MAX_DLIGHTS; i++)
	{
		rc_fxlight_t *dl = &x->lights[i];
		if (dl->die < time || !dl->radius)
			continue;
		if (scene)
			RC_SceneAddDlight (scene, dl->origin, dl->radius);
		n++;
		dl->radius -= dt*dl->decay;
		if (dl->radius < 0)
			dl->radius = 0;
	}
	return n;
}

/* ---- CL_AddEntities, cl_main.c:320-470 -------------------------------------------------------------------------------
 * The reference walks the entities of the last server packet; for each one it interpolates the state, applies the model
 * flags (rotating / bobbing items, trails) and the effect bits (spark field, muzzle flash and the two glow lights) and
 * adds an entity_t to the scene.  Here the caller passes what the reference reads from its centity_t / model / globals
 * in rc_cent_t and rc_clframe_t; everything that function writes back (lerp_origin) is written back. */
#define BOB_SCALE 180
static float fx_itembob[BOB_SCALE];	/* CL_BuildBobTable, cl_main.c:307-313 */
static int fx_itembob_built;

static void fx_lerp (const float *v1, float frac, const float *v2, float *v)	/* LerpVectors, mathlib.c:478 */
{
	v[0] = v1[0] + frac * (v2[0] - v1[0]);
	v[1] = v1[1] + frac * (v2[1] - v1[1]);
	v[2] = v1[2] + frac * (v2[2] - v1[2]);
}

static void fx_lerp_angles (const float *v1, float frac, const float *v2, float *v)	/* LerpAngles, mathlib.c:485: the short way round */
{
	for (int i = 0; i < 3; i++)
	{
		float d = v2[i] - v1[i];
		if (d > 180) d -= 360;
		else if (d < -180) d += 360;
		v[i] = v1[i] + frac * d;
	}
}

static float fx_anglemod (float a)	/* mathlib.c:152 */
{
	a = (360.0/65536) * ((int)(a*(65536/360.0)) & 65535);
	return a;
}

int RC_ClientAddEntities (rc_fx_t *x, rc_scene_t *scene, rc_cent_t *cents, int n, const rc_clframe_t *fr)
{
	int added = 0;
	if (!x || !cents || !fr)
		return -1;
	if (!fx_itembob_built)
	{
		for (int i = 0; i < BOB_SCALE; i++)
			fx_itembob[i] = sin (i / 90.0f * M_PI) * 5 + 5;
		fx_itembob_built = 1;
	}
	const double time = fr->time;
	for (int e = 0; e < n; e++)
	{
		rc_cent_t *c = &cents[e];
		entity_t ent;
		float f, oldorg[3];
		if (!c->model)
			continue;		/* empty slot / model not precached */
		memset (&ent, 0, sizeof(ent));
		ent.model = c->model;
		if (c->deltalerp > 0)
		{
			f = (time - c->startlerp) / c->deltalerp;
			f = f < 1 ? f : 1;		/* bound (0, f, 1) = max (0, min (f, 1)) */
			f = 0 > f ? 0 : f;
		}
		else
			f = 1;
		fx_lerp (c->prev_origin, f, c->cur_origin, ent.origin);
		fx_lerp_angles (c->prev_angles, f, c->cur_angles, ent.angles);
		memcpy (oldorg, c->lerp_origin, sizeof(oldorg));
		memcpy (c->lerp_origin, ent.origin, sizeof(oldorg));
		ent.colormap = c->colormap;
		ent.number = c->number;
		ent.syncbase = c->syncbase;
		ent.oldframe = c->prev_frame;
		ent.frame = c->cur_frame;
		ent.framelerp = (time - c->frametime) * 10;
		ent.framelerp = ent.framelerp < 1 ? ent.framelerp : 1;
		ent.framelerp = 0 > ent.framelerp ? 0 : ent.framelerp;
		ent.skinnum = c->skin;
		if (c->number > 0 && c->number <= fr->maxclients)
			ent.flags |= RF_MINLIGHT;		/* players never go totally black */
		if (c->modelflags & RC_MF_ROTATE)
		{
			const float bobjrotate = fx_anglemod (100*(time + c->number*0.1));
			ent.angles[1] = bobjrotate;
			if (fr->bobbingitems)
				ent.origin[2] += fx_itembob[(int)bobjrotate % (BOB_SCALE-1)];
		}
		if (c->effects)
		{
			if (c->effects & RC_EF_BRIGHTFIELD)
				fx_entity_particles (x, ent.origin, time);
			if (c->effects & RC_EF_MUZZLEFLASH)
			{
				float fv[3], rt[3], up[3];
				rc_fxlight_t *dl = RC_FxAllocDlight (x, c->number, time);
				memcpy (dl->origin, ent.origin, sizeof(oldorg));
				dl->origin[2] += 16;
				RC_AngleVectors (ent.angles, fv, rt, up);
				dl->origin[0] = dl->origin[0] + 18*fv[0]; dl->origin[1] = dl->origin[1] + 18*fv[1]; dl->origin[2] = dl->origin[2] + 18*fv[2];
				dl->radius = 200 + (rand ()&31);
				dl->die = time + 0.1;
			}
			if (c->effects & RC_EF_BRIGHTLIGHT)
			{
				rc_fxlight_t *dl = RC_FxAllocDlight (x, c->number, time);
				memcpy (dl->origin, ent.origin, sizeof(oldorg));
				dl->origin[2] += 16;
				dl->radius = 400 + (rand ()&31);
				dl->die = time + 0.001;
			}
			if (c->effects & RC_EF_DIMLIGHT)
			{
				rc_fxlight_t *dl = RC_FxAllocDlight (x, c->number, time);
				memcpy (dl->origin, ent.origin, sizeof(oldorg));
				dl->radius = 200 + (rand ()&31);
				dl->die = time + 0.001;
			}
		}
		if (c->modelflags & ~RC_MF_ROTATE)
		{
			const int mf = c->modelflags;
			if (mf & RC_MF_GIB) fx_trail (x, RC_FX_BLOOD_TRAIL, oldorg, ent.origin, 0, time);
			else if (mf & RC_MF_ZOMGIB) fx_trail (x, RC_FX_SLIGHT_BLOOD_TRAIL, oldorg, ent.origin, 0, time);
			else if (mf & RC_MF_TRACER) fx_trail (x, RC_FX_TRACER_TRAIL, oldorg, ent.origin, 52, time);
			else if (mf & RC_MF_TRACER2) fx_trail (x, RC_FX_TRACER_TRAIL, oldorg, ent.origin, 230, time);
			else if (mf & RC_MF_ROCKET)
			{
				fx_trail (x, RC_FX_ROCKET_TRAIL, oldorg, ent.origin, 0, time);
				rc_fxlight_t *dl = RC_FxAllocDlight (x, c->number, time);
				memcpy (dl->origin, ent.origin, sizeof(oldorg));
				dl->radius = 200;
				dl->die = time + 0.01;
			}
			else if (mf & RC_MF_GRENADE) fx_trail (x, RC_FX_GRENADE_TRAIL, oldorg, ent.origin, 0, time);
			else if (mf & RC_MF_TRACER3) fx_trail (x, RC_FX_VOOR_TRAIL, oldorg, ent.origin, 0, time);
		}
		if (c->number == fr->viewentity)
		{
			if (!fr->chase_active)
				continue;		/* the player's own model is only drawn from the chase camera */
			ent.flags |= RF_VIEWERMODEL;
			ent.angles[1] = fr->viewangles[1];
			ent.angles[0] = -fr->viewangles[0] * 0.3;
		}
		if (scene)
			RC_SceneAddEntity (scene, &ent);
		added++;
	}
	return added;
}

/* ---- colour shifts: what CL_AddCshifts hands to the scene (cl_view.c:321-550) ------------------------------------------
 * Four shifts per client -- the medium the eye is in, damage, item pick-up, power-up -- are blended into the palette by
 * R_UpdatePalette (a20).  The reference sends them to the scene only in frames in which one of them CHANGED since the
 * last call (the palette of the previous frame stays otherwise), and lets damage and bonus fade by the frame time. */
enum { CS_CONTENTS, CS_DAMAGE, CS_BONUS, CS_POWERUP, CS_COUNT };

void RC_FxSetContentsColor (rc_fx_t *x, int contents)	/* V_SetContentsColor, cl_view.c:432-455 */
{
	static const cshift_t lava = { { 255, 80, 0 }, 150 }, slime = { { 0, 25, 5 }, 150 }, water = { { 130, 80, 50 }, 128 };
	switch (contents)
	{
	case -1: case -2: x->cshifts[CS_CONTENTS] = x->cshift_empty; break;	/* CONTENTS_EMPTY, CONTENTS_SOLID */
	case -5: x->cshifts[CS_CONTENTS] = lava; break;
	case -4: x->cshifts[CS_CONTENTS] = slime; break;
	default: x->cshifts[CS_CONTENTS] = water; break;
	}
}

void RC_FxSetEmptyColor (rc_fx_t *x, int r, int g, int b, int percent)	/* the `v_cshift` command, cl_view.c:401-407 */
{
	x->cshift_empty.destcolor[0] = r; x->cshift_empty.destcolor[1] = g; x->cshift_empty.destcolor[2] = b; x->cshift_empty.percent = percent;
}

void RC_FxDamage (rc_fx_t *x, int armor, int blood)	/* the colour half of CL_ParseDamage, cl_view.c:331-376 */
{
	float count = (blood + armor) * 0.5;
	cshift_t *cs = &x->cshifts[CS_DAMAGE];
	if (count < 10)
		count = 10;
	if (armor > blood) { cs->destcolor[0] = 200; cs->destcolor[1] = 100; cs->destcolor[2] = 100; }
	else if (armor) { cs->destcolor[0] = 220; cs->destcolor[1] = 50; cs->destcolor[2] = 50; }
	else { cs->destcolor[0] = 255; cs->destcolor[1] = 0; cs->destcolor[2] = 0; }
	cs->percent += 3*count;
	if (cs->percent < 0) cs->percent = 0;
	if (cs->percent > 150) cs->percent = 150;
}

void RC_FxBonusFlash (rc_fx_t *x)	/* V_BonusFlash_f, cl_view.c:417-423 */
{
	cshift_t *cs = &x->cshifts[CS_BONUS];
	cs->destcolor[0] = 215; cs->destcolor[1] = 186; cs->destcolor[2] = 69; cs->percent = 50;
}

void RC_FxClearCshifts (rc_fx_t *x)	/* CL_ClearCshifts, cl_view.c:500-506 */
{
	for (int i = 0; i < CS_COUNT; i++)
		x->cshifts[i].percent = 0;
}

/* CL_AddCshifts, cl_view.c:513-550 (with V_CalcPowerupCshift, :460-498: `items` = cl.items).  Returns the number of shifts
 * handed to the scene: 0 in a frame in which nothing changed. */
int RC_FxAddCshifts (rc_fx_t *x, rc_scene_t *scene, int items, double host_frametime)
{
	static const struct { int bit; cshift_t cs; } powerups[4] = {		/* IT_QUAD, IT_SUIT, IT_INVISIBILITY, IT_INVULNERABILITY (quakedef.h:154-157) */
		{ 4194304, { { 0, 0, 255 }, 30 } }, { 2097152, { { 0, 255, 0 }, 20 } }, { 524288, { { 100, 100, 100 }, 100 } }, { 1048576, { { 255, 255, 100 }, 30 } } };
	int changed = 0, n = 0, k;
	for (k = 0; k < 4; k++)
		if (items & powerups[k].bit)
		{
			x->cshifts[CS_POWERUP] = powerups[k].cs;
			break;
		}
	if (k == 4)
		x->cshifts[CS_POWERUP].percent = 0;		/* (the colour stays) */
	for (int i = 0; i < CS_COUNT; i++)
	{
		if (x->cshifts[i].percent != x->prev_cshifts[i].percent) changed = 1;
		for (int j = 0; j < 3; j++)
			if (x->cshifts[i].destcolor[j] != x->prev_cshifts[i].destcolor[j]) changed = 1;
		x->prev_cshifts[i] = x->cshifts[i];
	}
	x->cshifts[CS_DAMAGE].percent -= host_frametime*150;
	if (x->cshifts[CS_DAMAGE].percent <= 0)
		x->cshifts[CS_DAMAGE].percent = 0;
	x->cshifts[CS_BONUS].percent -= host_frametime*100;
	if (x->cshifts[CS_BONUS].percent <= 0)
		x->cshifts[CS_BONUS].percent = 0;
	if (!changed)
		return 0;
	for (int i = 0; i < CS_COUNT; i++)
		if (x->cshifts[i].percent >= 0)
		{
			if (scene)
				RC_SceneAddCshift (scene, x->cshifts[i].destcolor[0], x->cshifts[i].destcolor[1], x->cshifts[i].destcolor[2], x->cshifts[i].percent);
			n++;
		}
	return n;
}

const cshift_t *RC_FxCshifts (const rc_fx_t *x) { return x->cshifts; }

/* ---- beams: the lightning bolts of cl_tent.c, a chain of model segments every 30 units -------------------------------- */
/* CL_ParseBeam, cl_tent.c:60-105: the beam of this entity is replaced, else the first free one is taken; 0: "beam list
 * overflow" */
int RC_FxBeam (rc_fx_t *x, int entity, struct model_s *model, const float *start, const float *end, double time)
{
	int i;
	for (i = 0; i < FX_MAX_BEAMS; i++)
		if (x->beams[i].entity == entity)
			break;
	if (i == FX_MAX_BEAMS)
	{
		for (i = 0; i < FX_MAX_BEAMS; i++)
			if (!x->beams[i].model || x->beams[i].endtime < time)
				break;
		if (i == FX_MAX_BEAMS)
			return 0;
		x->beams[i].entity = entity;
	}
	x->beams[i].model = model;
	x->beams[i].endtime = time + 0.2;
	memcpy (x->beams[i].start, start, sizeof(x->beams[i].start));
	memcpy (x->beams[i].end, end, sizeof(x->beams[i].end));
	return 1;
}

/* CL_AddTempEntities, cl_tent.c:303-356: every live beam becomes a chain of entities from its start (the viewer's
 * drawn position for the viewer's own beam) towards its end, rolled at random, full bright for the three bolt models; at
 * most 128 a frame.  The reference's segment loop reuses the beam loop's counter (it leaves it at 3), so a drawn beam in
 * slot k makes the loop look at slots k+1 .. k+20 only -- reproduced, short of reading past the array as the reference
 * does for k > 3.  Returns the entities added. */
int RC_ClientAddTempEntities (rc_fx_t *x, rc_scene_t *scene, double time, int viewentity, const float *view_lerp_origin,
			      byte *colormap, struct model_s *const *bolts)
{
	int i, b, added = 0;
	for (i = 0, b = 0; i < FX_MAX_BEAMS && b < FX_MAX_BEAMS; i++, b++)
	{
		float dist[3], org[3], ang[3], d;
		if (!x->beams[b].model || x->beams[b].endtime < time)
			continue;
		if (x->beams[b].entity == viewentity && view_lerp_origin)
			memcpy (x->beams[b].start, view_lerp_origin, sizeof(org));
		for (int j = 0; j < 3; j++)
			dist[j] = x->beams[b].end[j] - x->beams[b].start[j];
		/* Vector2Angles, mathlib.c:505 */
		if (!dist[0] && !dist[1])
		{
			ang[1] = 0;
			ang[0] = (dist[2] > 0) ? 90 : 270;
		}
		else
		{
			float y = atan2 (dist[1], dist[0]) * 180 / M_PI;
			if (y < 0) y += 360;
			float p = atan2 (dist[2], sqrt (dist[0]*dist[0]+dist[1]*dist[1])) * 180 / M_PI;
			if (p < 0) p += 360;
			ang[0] = p; ang[1] = y;
		}
		memcpy (org, x->beams[b].start, sizeof(org));
		d = fx_normalize (dist);
		while (d > 0)
		{
			entity_t ent;
			if (added == FX_MAX_TEMP_ENTITIES)
				return added;		/* CL_NewTempEntity: none left */
			memset (&ent, 0, sizeof(ent));
			memcpy (ent.origin, org, sizeof(org));
			ent.model = x->beams[b].model;
			ent.angles[0] = ang[0];
			ent.angles[1] = ang[1];
			ent.angles[2] = rand ()%360;
			ent.colormap = colormap;
			if (bolts && (ent.model == bolts[0] || ent.model == bolts[1] || ent.model == bolts[2]))
				ent.flags = RF_FULLBRIGHT|RF_NOSHADOW;
			if (scene)
				RC_SceneAddEntity (scene, &ent);
			added++;
			for (i = 0; i < 3; i++)
				org[i] += dist[i]*30;
			d -= 30;
		}
	}
	return added;
}

/* ---- the view itself: V_CalcRefdef / V_CalcIntermissionRefdef (cl_view.c:159-210, 646-886) ---------------------------------
 * From the viewer's interpolated state to the frame's refdef_t: eye position (view height, bob, 1/32 off any node plane,
 * bounded to the player's hull, stair steps smoothed), view angles (movement roll, damage kick, idle sway, punch angle,
 * the dead player's 80 degree roll), rectangle and field of view, RDF_UNDERWATER and the medium's colour shift from the
 * contents at the eye, and the gun model as an entity.  Not here: CL_DriftPitch (input state; a demo does not drift:
 * pass the angles as they are) and the chase camera (it traces against the collision hulls: pass chase_active = 0). */
static float fx_calc_roll (const rc_viewstate_t *in)		/* V_CalcRoll, cl_view.c:159-180 */
{
	float forward[3], right[3], up[3], sign, side, value;
	RC_AngleVectors (in->viewangles, forward, right, up);
	side = in->velocity[0]*right[0] + in->velocity[1]*right[1] + in->velocity[2]*right[2];
	sign = side < 0 ? -1 : 1;
	side = fabs (side);
	value = in->cl_rollangle;
	if (side < in->cl_rollspeed)
		side = side * value / in->cl_rollspeed;
	else
		side = value;
	return side*sign;
}

static float fx_calc_bob (const rc_viewstate_t *in)		/* V_CalcBob, cl_view.c:189-210 */
{
	float bob, cycle;
	cycle = in->time - (int)(in->time/in->cl_bobcycle)*in->cl_bobcycle;
	cycle /= in->cl_bobcycle;
	if (cycle < in->cl_bobup)
		cycle = M_PI * cycle / in->cl_bobup;
	else
		cycle = M_PI + M_PI*(cycle-in->cl_bobup)/(1.0 - in->cl_bobup);
	bob = sqrt (in->velocity[0]*in->velocity[0] + in->velocity[1]*in->velocity[1]) * in->cl_bob;
	bob = bob * (0.3 + 0.7*sin (cycle));
	bob = bob < 4 ? bob : 4;		/* bound (-7, bob, 4) */
	return -7 > bob ? -7 : bob;
}

/* the view-kick half of CL_ParseDamage (cl_view.c:378-391); RC_FxDamage is the colour half */
void RC_FxDamageKick (rc_fx_t *x, int armor, int blood, const float *from_in, const float *viewer_lerp_origin, const float *viewangles,
		      float v_kickroll, float v_kickpitch, float v_kicktime)
{
	float count = (blood + armor) * 0.5, from[3], forward[3], right[3], up[3], side;
	if (count < 10)
		count = 10;
	for (int i = 0; i < 3; i++) from[i] = from_in[i] - viewer_lerp_origin[i];
	fx_normalize (from);
	RC_AngleVectors (viewangles, forward, right, up);
	side = from[0]*right[0] + from[1]*right[1] + from[2]*right[2];
	x->dmg_roll = count*side*v_kickroll;
	side = from[0]*forward[0] + from[1]*forward[1] + from[2]*forward[2];
	x->dmg_pitch = count*side*v_kickpitch;
	x->dmg_time = v_kicktime;
}

int RC_ClientCalcRefdef (rc_fx_t *x, rc_scene_t *scene, const rc_viewstate_t *in, refdef_t *rd)
{
	if (!x || !in || !rd)
		return -1;
	if (in->fov_x < 1 || in->fov_x > 179)
		return -1;			/* CalcFov: Sys_Error ("Bad fov") there */
	rd->x = in->vrect[0]; rd->y = in->vrect[1]; rd->width = in->vrect[2]; rd->height = in->vrect[3];
	rd->fov_x = in->fov_x;
	rd->fov_y = RC_CalcFov (rd->fov_x, rd->width, rd->height);
	if (in->intermission)
	{
		/* V_CalcIntermissionRefdef, cl_view.c:758-795: the player entity's own origin and angles, always swaying */
		for (int i = 0; i < 3; i++) { rd->vieworg[i] = in->current_origin[i]; rd->viewangles[i] = in->current_angles[i]; }
		rd->viewangles[2] += 1 * sin (in->time*in->v_iroll_cycle) * in->v_iroll_level;
		rd->viewangles[0] += 1 * sin (in->time*in->v_ipitch_cycle) * in->v_ipitch_level;
		rd->viewangles[1] += 1 * sin (in->time*in->v_iyaw_cycle) * in->v_iyaw_level;
		rd->vieworg[0] += 1.0/32; rd->vieworg[1] += 1.0/32; rd->vieworg[2] += 1.0/32;
		return 0;
	}
	const float bob = fx_calc_bob (in);
	rd->vieworg[0] = in->lerp_origin[0]; rd->vieworg[1] = in->lerp_origin[1];
	rd->vieworg[2] = in->lerp_origin[2] + in->viewheight + bob;
	rd->vieworg[0] += 1.0/32; rd->vieworg[1] += 1.0/32; rd->vieworg[2] += 1.0/32;
	for (int i = 0; i < 3; i++) rd->viewangles[i] = in->viewangles[i];
	/* V_CalcViewRoll, :734-751 */
	rd->viewangles[2] += fx_calc_roll (in);
	if (x->dmg_time > 0)
	{
		rd->viewangles[2] += x->dmg_time/in->v_kicktime*x->dmg_roll;
		rd->viewangles[0] += x->dmg_time/in->v_kicktime*x->dmg_pitch;
		x->dmg_time -= in->host_frametime;
	}
	if (in->health <= 0)
		rd->viewangles[2] = 80;
	/* V_AddIdle (v_idlescale), :674-679 */
	rd->viewangles[2] += in->v_idlescale * sin (in->time*in->v_iroll_cycle) * in->v_iroll_level;
	rd->viewangles[0] += in->v_idlescale * sin (in->time*in->v_ipitch_cycle) * in->v_ipitch_level;
	rd->viewangles[1] += in->v_idlescale * sin (in->time*in->v_iyaw_cycle) * in->v_iyaw_level;
	/* V_BoundOffsets, :646-666: never outside the player's clipping hull */
	{
		static const float lo[3] = { 14, 14, 22 }, hi[3] = { 14, 14, 30 };
		for (int i = 0; i < 3; i++)
		{
			if (rd->vieworg[i] < in->lerp_origin[i] - lo[i]) rd->vieworg[i] = in->lerp_origin[i] - lo[i];
			else if (rd->vieworg[i] > in->lerp_origin[i] + hi[i]) rd->vieworg[i] = in->lerp_origin[i] + hi[i];
		}
	}
	for (int i = 0; i < 3; i++) rd->viewangles[i] = rd->viewangles[i] + in->punchangle[i];
	/* stair steps, :839-858: the eye follows a step up at 80 units a second, at most 12 behind */
	if (in->onground && in->lerp_origin[2] - x->oldz > 0)
	{
		float steptime = in->time - in->oldtime;
		if (steptime < 0)
			steptime = 0;
		x->oldz += steptime * 80;
		if (x->oldz > in->lerp_origin[2])
			x->oldz = in->lerp_origin[2];
		if (in->lerp_origin[2] - x->oldz > 12)
			x->oldz = in->lerp_origin[2] - 12;
		rd->vieworg[2] += x->oldz - in->lerp_origin[2];
	}
	else
		x->oldz = in->lerp_origin[2];
	/* V_AddGun, :687-726 */
	if (!in->chase_active && !(in->items & 524288 /* IT_INVISIBILITY */) && in->health > 0 && in->gun_model)
	{
		entity_t gun;
		float forward[3], right[3], up[3];
		memset (&gun, 0, sizeof(gun));
		gun.model = in->gun_model;
		gun.frame = in->gun_frame;
		gun.oldframe = in->gun_oldframe;
		gun.flags = RF_WEAPONMODEL|RF_DEPTHSCALE|RF_MINLIGHT|RF_NOSHADOW;
		gun.framelerp = (in->time - in->gun_frametime) * 10;
		gun.framelerp = gun.framelerp < 1 ? gun.framelerp : 1;
		gun.framelerp = 0 > gun.framelerp ? 0 : gun.framelerp;
		gun.colormap = in->colormap;
		for (int i = 0; i < 3; i++) gun.origin[i] = rd->vieworg[i];
		gun.angles[0] = -rd->viewangles[0]; gun.angles[1] = rd->viewangles[1]; gun.angles[2] = in->viewangles[2];
		gun.angles[2] -= in->v_idlescale * sin (in->time*in->v_iroll_cycle) * in->v_iroll_level;
		gun.angles[0] -= in->v_idlescale * sin (in->time*in->v_ipitch_cycle) * in->v_ipitch_level;
		gun.angles[1] -= in->v_idlescale * sin (in->time*in->v_iyaw_cycle) * in->v_iyaw_level;
		RC_AngleVectors (gun.angles, forward, right, up);
		for (int i = 0; i < 3; i++) gun.origin[i] = gun.origin[i] + (bob * 0.4)*forward[i];
		if (scene)
			RC_SceneAddEntity (scene, &gun);
		return 1;
	}
	return 0;
}

/* what both view functions do once the eye is placed (cl_view.c:783-794, 865-873): RDF_UNDERWATER and the colour of the
 * medium from the contents at the eye (RC_PointContents (rd->vieworg)) */
void RC_ClientSetViewContents (rc_fx_t *x, refdef_t *rd, int contents)
{
	rd->flags = contents <= -3 /* CONTENTS_WATER */ ? RDF_UNDERWATER : 0;
	RC_FxSetContentsColor (x, contents);
}
