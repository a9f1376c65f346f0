/* TEST INFRASTRUCTURE ONLY.
 * Platform shim around the UNMODIFIED client/cl_effects.c + cl_main.c + cl_tent.c + cl_view.c + common/mathlib.c of the reference, compiled
 * from where they lie into oracle/_ref/libclfx.so (oracle/Makefile, target `clfx`): the particle system, the dynamic
 * lights and CL_AddEntities, i.e. what tochris_b200/csrc/rc_client.c restates.  Nothing of the product links or loads
 * this; tests/test_client_fx.py drives both sides with the same calls and the same srand () seed.
 * The scene lists are replaced by recorders (V_AddEntity / V_AddParticle / V_AddDlight append to arrays the test reads). */
#include "client.h"

/* ---- data the two client files expect from the rest of the engine ---- */
extern cvar_t chase_active;	/* (cl_view.c) */
cvar_t      cl_forwardspeed = { "cl_forwardspeed", "200" }, scr_fov = { "fov", "90" };
qboolean    con_forcedup;
vrect_t     scr_vrect;
double      host_frametime, realtime;
float       scr_centertime_off;
viddef_t    vid;
char      **com_argv;
centity_t   cl_entities[MAX_EDICTS];
int         cl_parse_entities[MAX_PARSE_ENTITIES];
entity_t    cl_static_entities[MAX_STATIC_ENTITIES];
lightstyle_t cl_lightstyle[MAX_LIGHTSTYLES];

static int clfx_maxclients = 1;
int  Com_ClientMaxclients (void) { return clfx_maxclients; }
server_state_t Com_ServerState (void) { return 0; }
int  Com_ServerMaxclients (void) { return 1; }
cl_state_t Com_ClientState (void) { return 0; }
int  COM_CheckParm (char *parm) { (void)parm; return 0; }
void Cmd_AddCommand (char *name, xcommand_t fn) { (void)name; (void)fn; }
void Con_Printf (char *fmt, ...) { (void)fmt; }
void Con_DPrintf (char *fmt, ...) { (void)fmt; }
int  Mod_Flags (struct model_s *m) { return *(int *)m; }		/* the tests' "models" are ints holding the flags */
void Sys_Error (char *error, ...) { fprintf (stderr, "libclfx: Sys_Error: %s\n", error); abort (); }
/* the server message the parsers read (CL_ParseDamage, CL_ParseBeam): numbers queued by the test */
static float clfx_msg[64]; static int clfx_msgn, clfx_msgpos;
void  clfx_msg_set (const float *v, int n) { memcpy (clfx_msg, v, sizeof(float) * (n < 64 ? n : 64)); clfx_msgn = n; clfx_msgpos = 0; }
static float clfx_msg_next (void) { return clfx_msgpos < clfx_msgn ? clfx_msg[clfx_msgpos++] : 0; }
int   MSG_ReadByte (void) { return (int)clfx_msg_next (); }
int   MSG_ReadShort (void) { return (int)clfx_msg_next (); }
float MSG_ReadCoord (void) { return clfx_msg_next (); }
void clfx_unexpected (const char *name) { fprintf (stderr, "libclfx: %s called (not part of the path under test)\n", name); abort (); }

/* ---- recorders in place of client/cl_view.c's lists ---- */
#define REC_MAX 8192
static particle_t rec_particles[REC_MAX]; static int rec_nparticles;
static dlight_t   rec_dlights[REC_MAX];   static int rec_ndlights;
static entity_t   rec_entities[REC_MAX];  static int rec_nentities;
void V_AddParticle (vec3_t org, int color)
{
	if (rec_nparticles < REC_MAX) { VectorCopy (org, rec_particles[rec_nparticles].org); rec_particles[rec_nparticles].color = color; rec_nparticles++; }
}
void V_AddDlight (vec3_t org, float radius)
{
	if (rec_ndlights < REC_MAX) { VectorCopy (org, rec_dlights[rec_ndlights].origin); rec_dlights[rec_ndlights].radius = radius; rec_ndlights++; }
}
void V_AddEntity (entity_t *ent)
{
	if (rec_nentities < REC_MAX) rec_entities[rec_nentities++] = *ent;
}
extern int r_numcshifts; extern cshift_t r_cshifts[];	/* cl_view.c's own list (its V_AddCshift is left in place) */
void clfx_rec_clear (void) { rec_nparticles = rec_ndlights = rec_nentities = 0; r_numcshifts = 0; }
int  clfx_rec_cshifts (cshift_t *out, int max)
{
	int i;
	for (i = 0; i < r_numcshifts && i < max; i++) out[i] = r_cshifts[i];
	return r_numcshifts;
}
void clfx_cshifts (cshift_t *out) { memcpy (out, cl.cshifts, sizeof(cshift_t) * NUM_CSHIFTS); }
extern struct model_s *cl_bolt1_mod, *cl_bolt2_mod, *cl_bolt3_mod;
void clfx_set_bolts (struct model_s *a, struct model_s *b, struct model_s *c) { cl_bolt1_mod = a; cl_bolt2_mod = b; cl_bolt3_mod = c; }
void clfx_set_frame (double time, double oldtime, double frametime, int items, int viewentity, const float *view_lerp_origin)
{
	cl.time = time; cl.oldtime = oldtime; host_frametime = frametime; cl.items = items; cl.viewentity = viewentity;
	if (view_lerp_origin) VectorCopy (view_lerp_origin, cl_entities[viewentity].lerp_origin);
}
int  clfx_rec_particles (float *org3, int *color, int max)
{
	int i;
	for (i = 0; i < rec_nparticles && i < max; i++) { memcpy (org3 + 3*i, rec_particles[i].org, 12); color[i] = rec_particles[i].color; }
	return rec_nparticles;
}
int  clfx_rec_dlights (float *org3, float *radius, int max)
{
	int i;
	for (i = 0; i < rec_ndlights && i < max; i++) { memcpy (org3 + 3*i, rec_dlights[i].origin, 12); radius[i] = rec_dlights[i].radius; }
	return rec_ndlights;
}
int  clfx_rec_entities (entity_t *out, int max)
{
	int i;
	for (i = 0; i < rec_nentities && i < max; i++) out[i] = rec_entities[i];
	return rec_nentities;
}

/* ---- state ---- */
extern cparticle_t *active_particles, *free_particles;
extern cparticle_t particles[MAX_PARTICLES];
extern int cl_numparticles;
extern vec3_t avelocities[];
extern void CL_InitParticles (void);
extern void CL_ClearParticles (void);
extern void CL_BuildBobTable (void);
extern void CL_AddEntities (void);

void clfx_init (unsigned seed, int numparticles)
{
	srand (seed);
	memset (&cl, 0, sizeof(cl));
	memset (&cls, 0, sizeof(cls));
	memset (cl_dlights, 0, sizeof(cl_dlights));
	memset (cl_beams, 0, sizeof(cl_beams));
	memset (particles, 0, sizeof(particles));
	memset (avelocities, 0, 162 * sizeof(vec3_t));
	CL_InitParticles ();
	if (numparticles > 0) cl_numparticles = numparticles < 512 ? 512 : numparticles;	/* what -particles N does, cl_effects.c:56-62 */
	CL_ClearParticles ();
	CL_BuildBobTable ();
	clfx_rec_clear ();
}
void clfx_set_time (double time, double oldtime) { cl.time = time; cl.oldtime = oldtime; }

/* the active list, head first */
int clfx_particles (float *org3, float *vel3, float *ramp, float *die, int *color, int *type, int max)
{
	int n = 0;
	cparticle_t *p;
	for (p = active_particles; p; p = p->next, n++)
		if (n < max)
		{
			memcpy (org3 + 3*n, p->org, 12); memcpy (vel3 + 3*n, p->vel, 12);
			ramp[n] = p->ramp; die[n] = p->die; color[n] = p->color; type[n] = p->type;
		}
	return n;
}
void clfx_dlights (cdlight_t *out) { memcpy (out, cl_dlights, sizeof(cl_dlights)); }

/* CL_AddEntities over the test's entity records (the layout of rc_cent_t / rc_clframe_t in include/ref_cuda.h) */
typedef struct {
	int number; struct model_s *model; int modelflags; byte *colormap;
	vec3_t prev_origin, prev_angles; int prev_frame;
	vec3_t cur_origin, cur_angles; int cur_frame, effects, skin;
	float startlerp, deltalerp, frametime, syncbase;
	vec3_t lerp_origin;
} clfx_cent_t;
typedef struct { double time; int maxclients, viewentity, chase_active, bobbingitems; vec3_t viewangles; } clfx_frame_t;

int clfx_add_entities (clfx_cent_t *in, int n, const clfx_frame_t *fr)
{
	int i;
	cl.time = fr->time;
	cl.frame.time = fr->time + 0.05; cl.oldframe.time = fr->time - 0.05;	/* CL_LerpPoint leaves cl.time alone */
	cl.frame.num_entities = n; cl.frame.parse_entities = 0;
	cl.viewentity = fr->viewentity;
	VectorCopy (fr->viewangles, cl.viewangles);
	clfx_maxclients = fr->maxclients;
	chase_active.value = fr->chase_active;
	cl_bobbingitems.value = fr->bobbingitems;
	for (i = 0; i < n; i++)
	{
		centity_t *c = &cl_entities[in[i].number];
		cl_parse_entities[i] = in[i].number;
		memset (c, 0, sizeof(*c));
		c->startlerp = in[i].startlerp; c->deltalerp = in[i].deltalerp; c->frametime = in[i].frametime; c->syncbase = in[i].syncbase;
		VectorCopy (in[i].prev_origin, c->prev.origin); VectorCopy (in[i].prev_angles, c->prev.angles); c->prev.frame = in[i].prev_frame;
		VectorCopy (in[i].cur_origin, c->current.origin); VectorCopy (in[i].cur_angles, c->current.angles); c->current.frame = in[i].cur_frame;
		c->current.effects = in[i].effects; c->current.skin = in[i].skin; c->current.colormap = 0;
		c->current.modelindex = in[i].model ? i + 1 : 0;
		cl.model_precache[i + 1] = in[i].model;
		VectorCopy (in[i].lerp_origin, c->lerp_origin);
	}
	CL_AddEntities ();
	for (i = 0; i < n; i++)
		VectorCopy (cl_entities[in[i].number].lerp_origin, in[i].lerp_origin);
	return rec_nentities;
}
void *clfx_vid_colormap (void) { return vid.colormap; }

/* ---- V_CalcRefdef / V_CalcIntermissionRefdef over the test's view state (the layout of rc_viewstate_t) ---- */
typedef struct {
	double time, oldtime, host_frametime;
	vec3_t lerp_origin, current_origin, current_angles, velocity, viewangles, punchangle;
	float viewheight;
	int onground, health, items, intermission, chase_active;
	int vrect[4];
	float fov_x;
	float cl_rollangle, cl_rollspeed, cl_bob, cl_bobcycle, cl_bobup, v_kicktime;
	float v_idlescale, v_iyaw_cycle, v_iroll_cycle, v_ipitch_cycle, v_iyaw_level, v_iroll_level, v_ipitch_level;
	struct model_s *gun_model; int gun_frame, gun_oldframe; float gun_frametime;
	byte *colormap;
} clfx_viewstate_t;
extern cvar_t cl_rollangle, cl_rollspeed, cl_bob, cl_bobcycle, cl_bobup, v_kicktime, v_kickroll, v_kickpitch, v_idlescale,
	v_iyaw_cycle, v_iroll_cycle, v_ipitch_cycle, v_iyaw_level, v_iroll_level, v_ipitch_level;
extern void V_CalcRefdef (void);
extern void V_CalcIntermissionRefdef (void);
extern int r_numentities; extern entity_t r_entities[];
static int clfx_contents = -1;
int  CM_PointContents (vec3_t p) { (void)p; return clfx_contents; }
void clfx_set_contents (int c) { clfx_contents = c; }
void clfx_set_kick (float roll, float pitch) { v_kickroll.value = roll; v_kickpitch.value = pitch; }
void clfx_set_viewer (int viewentity, const float *lerp_origin, const float *viewangles, float kicktime)
{
	cl.viewentity = viewentity; v_kicktime.value = kicktime;
	VectorCopy (lerp_origin, cl_entities[viewentity].lerp_origin);
	VectorCopy (viewangles, cl.viewangles);
}
/* returns the entities the view code added (the gun), copied to *gun */
int clfx_calc_refdef (const clfx_viewstate_t *in, refdef_t *out, entity_t *gun)
{
	centity_t *cent;
	cls.demoplayback = true;		/* no pitch drift: input state */
	cl.viewentity = 1; cent = &cl_entities[1];
	cl.time = in->time; cl.oldtime = in->oldtime; host_frametime = in->host_frametime;
	VectorCopy (in->lerp_origin, cent->lerp_origin);
	VectorCopy (in->current_origin, cent->current.origin); VectorCopy (in->current_angles, cent->current.angles);
	VectorCopy (in->velocity, cl.velocity); VectorCopy (in->viewangles, cl.viewangles); VectorCopy (in->punchangle, cl.punchangle);
	cl.viewheight = in->viewheight; cl.onground = in->onground; cl.stats[STAT_HEALTH] = in->health; cl.items = in->items;
	chase_active.value = in->chase_active;
	scr_vrect.x = in->vrect[0]; scr_vrect.y = in->vrect[1]; scr_vrect.width = in->vrect[2]; scr_vrect.height = in->vrect[3];
	scr_fov.value = in->fov_x;
	cl_rollangle.value = in->cl_rollangle; cl_rollspeed.value = in->cl_rollspeed; cl_bob.value = in->cl_bob;
	cl_bobcycle.value = in->cl_bobcycle; cl_bobup.value = in->cl_bobup; v_kicktime.value = in->v_kicktime;
	v_idlescale.value = in->v_idlescale; v_iyaw_cycle.value = in->v_iyaw_cycle; v_iroll_cycle.value = in->v_iroll_cycle;
	v_ipitch_cycle.value = in->v_ipitch_cycle; v_iyaw_level.value = in->v_iyaw_level; v_iroll_level.value = in->v_iroll_level;
	v_ipitch_level.value = in->v_ipitch_level;
	memset (&cl.viewent, 0, sizeof(cl.viewent));
	cl.viewent.current.modelindex = in->gun_model ? 7 : 0; cl.model_precache[7] = in->gun_model;
	cl.viewent.current.frame = in->gun_frame; cl.viewent.prev.frame = in->gun_oldframe; cl.viewent.frametime = in->gun_frametime;
	r_numentities = 0;
	if (in->intermission) V_CalcIntermissionRefdef (); else V_CalcRefdef ();
	*out = cl.refdef;
	if (r_numentities) *gun = r_entities[0];
	return r_numentities;
}
