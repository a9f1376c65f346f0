/*
 * ref_cuda.h -- C-ABI of libref_cuda.so, the B200 drop-in for the ref_soft renderer of
 * viciious/tochris (ToChriS Quake 1.67).
 *
 * The reference has no plugin struct: its renderer boundary is the set of C symbols in
 * client/render.h:25-53 plus the data structs of client/ref.h:37-108 and the `vid`
 * global of client/vid.h:34-44.  This header declares
 *   (1) the same data structs, byte-compatible on LP64 (refdef_t 136 B, entity_t 80 B ...),
 *   (2) the render.h entry points with identical names and signatures, so engine objects
 *       link against libref_cuda.so in place of ref_soft/ *.o,
 *   (3) a batched extension (RC_*) that renders many independent views per call and
 *       reports errors through return codes instead of Sys_Error/exit,
 *   (4) a refexport_t-shaped table (GetRefAPI) with BASELINE.json's north_star naming.
 * Only plain pointers and sizes cross this boundary.
 */
#ifndef REF_CUDA_H
#define REF_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ client/ref.h ---- */
#ifndef RC_NO_REF_TYPES
typedef float vec_t;
typedef vec_t vec3_t[3];
typedef unsigned char byte;
typedef int qboolean;

#define MAX_LIGHTSTYLES 64          /* common/quakedef.h:105 */
#define MAX_STYLESTRING 64          /* common/quakedef.h:111 */
#define MAX_DLIGHTS     32          /* client/ref.h:31 */

#define RF_VIEWERMODEL  1           /* common/quakedef.h:228-234 */
#define RF_WEAPONMODEL  2
#define RF_DEPTHSCALE   4
#define RF_MINLIGHT     8
#define RF_FULLBRIGHT   16
#define RF_NOSHADOW     32
#define RF_CULLVIS      64
#define RDF_NOWORLDMODEL 1          /* common/quakedef.h:238-239 */
#define RDF_UNDERWATER   2

typedef struct { vec3_t org; int color; } particle_t;                 /* ref.h:37-41 */
typedef struct { vec3_t origin; float radius; } dlight_t;             /* ref.h:43-47 */
typedef struct { int destcolor[3]; int percent; } cshift_t;           /* ref.h:49-53 */
typedef struct { int length; char map[MAX_STYLESTRING]; } lightstyle_t; /* ref.h:55-59 */

struct model_s;
typedef struct entity_s {                                             /* ref.h:61-80 */
    float           framelerp;
    struct model_s *model;
    int             number;
    int             flags;
    int             skinnum;
    byte           *colormap;
    vec3_t          origin;
    vec3_t          angles;
    int             frame;
    int             oldframe;
    float           syncbase;
} entity_t;

typedef struct {                                                      /* ref.h:82-107 */
    int           x, y, width, height;
    float         time;
    float         oldtime;
    int           flags;
    float         fov_x, fov_y;
    vec3_t        vieworg;
    vec3_t        viewangles;
    lightstyle_t *lightstyles;       /* [MAX_LIGHTSTYLES] */
    int           num_entities;
    entity_t     *entities;
    int           num_dlights;
    dlight_t     *dlights;
    int           num_particles;
    particle_t   *particles;
    int           num_cshifts;
    cshift_t     *cshifts;
} refdef_t;
#endif /* RC_NO_REF_TYPES */

/* ------------------------------------------------------- client/render.h:25-53 set ---- */
struct model_s *Mod_ForName (char *name, qboolean crash);      /* render.h:25  r_model.c:339 */
void  Mod_TouchModel (char *name);                             /* render.h:26 */
void  R_Init (void);                                           /* render.h:28  r_main.c:185 */
void  R_RenderView (refdef_t *refdef);                         /* render.h:29  r_main.c:988 */
void  Mod_Init (void);                                         /* render.h:32 */
void  Mod_ClearAll (void);                                     /* render.h:33 */
int   Mod_Flags (struct model_s *model);                       /* render.h:34 */
void  R_BeginFrame (void);                                     /* render.h:36  r_main.c:951 */
void  R_EndFrame (void);                                       /* render.h:37  r_main.c:961 */
void  R_NewMap (struct model_s *worldmodel);                   /* render.h:39  r_main.c:258 */
void  R_ScreenShot_f (void);                                   /* render.h:30  r_misc.c:117: <basedir>/id1/quakeNN.pcx of vid.buffer */
void  R_TranslatePlayerSkin (int playernum, struct model_s *model, int skinnum, int top, int bottom); /* render.h:40 (empty in ref_soft, r_misc.c:372) */
/* 2D (render.h:42-53, ref_soft/r_draw.c): recorded per frame, applied to vid.buffer by R_EndFrame; mpic_t as r_shared.h:78-85 */
struct mpic_s;
void  Draw_Init (void);
void  Draw_Character (int x, int y, int num);
void  Draw_Pic (int x, int y, struct mpic_s *pic);
void  Draw_TransPic (int x, int y, struct mpic_s *pic);
void  Draw_TransPicTranslate (int x, int y, struct mpic_s *pic, unsigned char *translation);
void  Draw_ConsoleBackground (int lines);
void  Draw_TileClear (int x, int y, int w, int h);
void  Draw_Fill (int x, int y, int w, int h, int c);
void  Draw_FadeScreen (void);
void  Draw_String (int x, int y, char *str);
struct mpic_s *Draw_PicFromWad (char *name);
struct mpic_s *Draw_CachePic (char *path);

/* ---------------------------------------------------------------- batched extension ---- */
typedef struct {
    int   width, height;      /* vid.width / vid.height: every view of a batch shares them */
    int   device;             /* CUDA device ordinal (one process per GPU) */
    const char *basedir;      /* files are read from <basedir>/id1/<name> like COM_LoadFile */
    /* cvars that change pixels (r_main.c:108-134, d_init.c:26-28); RC_DefaultConfig fills
     * the reference defaults */
    float r_clearcolor, r_fullbright, r_ambient, r_drawentities, r_drawviewmodel, r_waterwarp,
          r_fastsky, r_skycolor, r_lerpmodels, d_mipcap, d_mipscale, r_maxsurfs, r_maxedges;
    size_t surfcache_bytes;   /* device surface-cache pool (0: default) */
    int   max_views_in_flight;/* views processed per pipeline pass (0: derived from memory) */
    int   host_threads;       /* host threads that set up the views of a large batch (0: min (8, online CPUs); one process
                               * per GPU shares the host: give each rank its share) */
    float r_draworder;        /* r_main.c:108: non-zero draws the FARTHEST surface of every pixel (R_GenerateSpansBackward,
                               * r_edge.c:573; the reference honours it only while a local server runs, r_misc.c:420) */
} rc_config_t;

void RC_DefaultConfig (rc_config_t *cfg);
/* the configuration of the running renderer (zeroes before RC_Init) */
void RC_GetConfig (rc_config_t *cfg);

/* return codes and RC_STAT_* status bits (rc_types.h holds the same values for the library's own files) */
#ifndef RC_OK
#define RC_OK                    0
#define RC_ERR_CUDA              (-1)
#define RC_ERR_ARG               (-2)
#define RC_ERR_NOWORLD           (-3)
#define RC_ERR_FILE              (-4)
#define RC_ERR_FORMAT            (-5)
#define RC_ERR_NOMEM             (-6)
#define RC_STAT_EDGE_OVERFLOW    1    /* a view produced more edges than r_maxedges: the reference drops faces (r_rast.c:544) */
#define RC_STAT_SURF_OVERFLOW    2    /* ditto r_maxsurfs (r_rast.c:537) */
#define RC_STAT_SPAN_POOL        4    /* internal span pool was too small (the blocking calls retry by themselves) */
#define RC_STAT_AET_OVERFLOW     8    /* more active edges on one scanline than the scan kernel holds (ditto) */
#define RC_STAT_STACK_OVERFLOW   16   /* active surface stack deeper than the scan kernel holds */
#define RC_STAT_CACHE_POOL       32   /* surface cache pool too small (ditto) */
#define RC_STAT_STALE_CLIPVERT   64   /* the reference would read r_rightenter / r_rightexit left over from the previous frame */
#define RC_STAT_FACELIST         128  /* face list capacity exceeded */
#define RC_STAT_ALIAS_RANGE      512  /* the entity rasterisers met a case where the reference itself reads or loops out of range */
#define RC_STAT_BMODEL_OVERFLOW  1024 /* brush entity clipping ran out of vertices / edges (MAX_BMODEL_VERTS / EDGES) */
#endif

/* Creates the renderer (CUDA context, tables, palette/colormap from <basedir>/id1/gfx).
 * Replaces VID_Init + R_Init (common/host.c:837-841).  Returns RC_OK or a negative RC_ERR_*. */
int  RC_Init (const rc_config_t *cfg);
void RC_Shutdown (void);
const char *RC_LastError (void);

/* Mod_ForName + R_NewMap in one call, with an error code instead of Sys_Error
 * (client/cl_parse.c:271,294). */
int  RC_LoadWorld (const char *name);
/* the r_loadlits cvar (r_main.c:130, Mod_LoadLighting r_model.c:528-566): brush models loaded afterwards take their
 * light samples from "<map>.lit" (QLIT version 1, max of r, g, b) when that file exists */
void RC_SetLoadLits (int on);
/* the step before the path: the contents of the world leaf holding a point (Mod_PointInLeaf,
 * r_model.c:83) and the refdef_t.flags the client derives from it (RDF_UNDERWATER when contents <= CONTENTS_WATER,
 * cl_view.c:786-791) */
/* the `timerefresh` console command (R_TimeRefresh_f, r_misc.c:29-64), the reference's own renderer benchmark: the
 * 128-step yaw sweep of `base` as ONE batch; fb_out NULL keeps the frames on the device */
int  RC_TimeRefresh (const refdef_t *base, uint8_t *fb_out, double *seconds, int *status);
int  RC_PointContents (const float *org);
int  RC_ViewFlagsForOrigin (const float *org);

/* ---- device lifetimes of the alias models (ref_soft/r_model.c:95-130, 247-258, common/zone.c) ----
 * The reference keeps alias models in the Cache_Alloc heap: the least recently used ones are thrown out when it is short
 * of room, reloaded when asked for again, and Mod_TouchModel refreshes a model's place in that order.  The device copy
 * follows the same rules under RC_SetModelCacheBytes (0, the default: everything loaded stays resident): the models a
 * batch draws are resident while it is drawn (even above the budget); when one of them is not, the resident set is
 * rebuilt -- the batch's models, then the rest from the most to the least recently used while they fit -- from the
 * host copy.  Pixels never depend on it. */
typedef struct {
    int    loaded, resident;          /* alias models loaded on the host / resident on the device */
    size_t resident_bytes, budget_bytes;
    int    uploads, evictions;        /* models that came onto / left the device since RC_Init */
    int    rebuilds;                  /* times the resident set was re-packed */
} rc_modelcache_t;
void RC_SetModelCacheBytes (size_t bytes);
void RC_ModelCacheInfo (rc_modelcache_t *out);

/* ---- the client's scene lists and `timedemo` in batched form (rc_scene.c) ----
 * rc_scene_t keeps the four per-frame lists of client/cl_view.c:67-78 with the reference's limits (256 entities, 2048
 * particles, 32 dlights, 16 colour shifts; what is added to a full list is dropped: the Add calls return 0) and, unlike
 * the reference's globals, keeps every finished frame, so that a demo or a roll-out becomes one refdef_t array:
 *   RC_SceneClear = V_ClearScene, RC_SceneAdd* = V_AddEntity / V_AddParticle / V_AddDlight / V_AddCshift (cl_view.c:85-148),
 *   RC_SceneEndFrame = the assembly in V_RenderView (cl_view.c:915-940; add[4] = cl_add_entities / _particles / _lights /
 *   _cshifts, NULL = all on), RC_SceneFrames = the frames so far, RC_CalcFov = CalcFov (cl_view.c:629; 0 for a bad fov_x),
 *   RC_TimeDemo = timedemo (cl_demo.c:324-365): frame 0 does not count, the rest is rendered `batch` views per call;
 *   fb_out NULL keeps the frames on the device, else n * height * width bytes through the pipelined host path. */
typedef struct rc_scene_s rc_scene_t;
typedef struct {
    int    frames;            /* n - 1: "the first frame didn't count" */
    double seconds, fps;
    int    status;            /* RC_STAT_* bits of all batches */
    char   text[64];          /* "%i frames %5.1f seconds %5.1f fps\n", CL_FinishTimeDemo's line */
} rc_timedemo_t;
rc_scene_t *RC_SceneNew (void);
void  RC_SceneFree (rc_scene_t *s);
void  RC_SceneClear (rc_scene_t *s);
void  RC_SceneReset (rc_scene_t *s);
int   RC_SceneAddEntity (rc_scene_t *s, const entity_t *ent);
int   RC_SceneAddParticle (rc_scene_t *s, const float *org, int color);
int   RC_SceneAddDlight (rc_scene_t *s, const float *org, float radius);
int   RC_SceneAddCshift (rc_scene_t *s, int r, int g, int b, int percent);
int   RC_SceneEndFrame (rc_scene_t *s, const refdef_t *view, const int *add);
const refdef_t *RC_SceneFrames (rc_scene_t *s, int *n);
float RC_CalcFov (float fov_x, float width, float height);
int   RC_TimeDemo (const refdef_t *frames, int n, int batch, uint8_t *fb_out, rc_timedemo_t *out);

/* ---- what FILLS those lists: the client's particle system, dynamic lights and entity interpolation (rc_client.c) ----
 * rc_fx_t owns what client/cl_effects.c keeps in globals (particles[], the active / free lists, cl_dlights[], the spark
 * field's avelocities, the tracer counter); one per client, any number of clients.
 *   RC_FxNew (n)           CL_InitParticles + CL_ClearParticles: n = the `-particles` argument (0: MAX_PARTICLES; floor 512)
 *   RC_FxClear             CL_ClearParticles (cl_effects.c:147) and the light array zeroed as at a map change
 *   RC_FxEffect            every spawner of cl_effects.c behind one call; rand () is drawn in the reference's order.
 *                          Returns the particles made (the reference stops silently at an empty free list), -1: bad argument
 *        effect                     a            b      i0          i1           reference
 *        RC_FX_ENTITY_PARTICLES     ent origin   -      -           -            CL_EntityParticles    :98
 *        RC_FX_POINT                org          -      count c     -            a line of CL_ReadPointFile_f :159
 *        RC_FX_EXPLOSION            org          -      -           -            CL_ParticleExplosion  :233
 *        RC_FX_EXPLOSION2           org          -      colorStart  colorLength  CL_ParticleExplosion2 :274
 *        RC_FX_BLOB_EXPLOSION       org          -      -           -            CL_BlobExplosion      :304
 *        RC_FX_RUN_EFFECT           org          dir    color       count        CL_RunParticleEffect  :348
 *        RC_FX_ROCKET_SPLASH        org          -      -           -            CL_RocketSplash       :372
 *        RC_FX_LAVA_SPLASH          org          -      -           -            CL_LavaSplash         :411
 *        RC_FX_TELEPORT_SPLASH      org          -      -           -            CL_TeleportSplash     :517
 *        RC_FX_RAIL_TRAIL           start        end    -           -            CL_RailTrail          :446
 *        RC_FX_ROCKET_TRAIL .. RC_FX_VOOR_TRAIL  start  end  (tracer: color)     CL_RocketTrail :548 .. CL_VoorTrail :747
 *   RC_FxAddParticles      CL_AddParticles (:784): dead particles leave, the living go to the scene (newest first, as the
 *                          reference's list runs) and are moved on by time - oldtime; returns how many went to the scene
 *   RC_FxAllocDlight       CL_AllocDlight (:915): same key, else first dead, else light 0; zeroed, the caller fills it
 *   RC_FxAddDlights        CL_AddDlights (:959): live lights to the scene, then radius -= dt * decay
 *   RC_ClientAddEntities   the loop body of CL_AddEntities (cl_main.c:320-470) over rc_cent_t[]: interpolation of origin,
 *                          angles (the short way round) and frame, RF_MINLIGHT for players, rotating / bobbing items,
 *                          EF_* effects (spark field, muzzle flash, glow lights), MF_* trails, the viewer's own model only
 *                          from the chase camera (RF_VIEWERMODEL); writes lerp_origin back.  Returns the entities added. */
typedef struct rc_fx_s rc_fx_t;
typedef struct { vec3_t origin; float radius; float die; float decay; int key; } rc_fxlight_t;   /* cdlight_t, client.h:91-98 */
enum { RC_FX_ENTITY_PARTICLES, RC_FX_POINT, RC_FX_EXPLOSION, RC_FX_EXPLOSION2, RC_FX_BLOB_EXPLOSION, RC_FX_RUN_EFFECT,
       RC_FX_ROCKET_SPLASH, RC_FX_LAVA_SPLASH, RC_FX_TELEPORT_SPLASH, RC_FX_RAIL_TRAIL, RC_FX_ROCKET_TRAIL,
       RC_FX_GRENADE_TRAIL, RC_FX_BLOOD_TRAIL, RC_FX_SLIGHT_BLOOD_TRAIL, RC_FX_TRACER_TRAIL, RC_FX_VOOR_TRAIL };
#define RC_EF_BRIGHTFIELD 1          /* common/quakedef.h:210-213 */
#define RC_EF_MUZZLEFLASH 2
#define RC_EF_BRIGHTLIGHT 4
#define RC_EF_DIMLIGHT    8
#define RC_MF_ROCKET   1             /* model flags (Mod_Flags), common/quakedef.h:217-224 */
#define RC_MF_GRENADE  2
#define RC_MF_GIB      4
#define RC_MF_ROTATE   8
#define RC_MF_TRACER   16
#define RC_MF_ZOMGIB   32
#define RC_MF_TRACER2  64
#define RC_MF_TRACER3  128
typedef struct {                     /* what CL_AddEntities reads of a centity_t (client.h:62-78), its model and colormap */
    int    number;                   /* index into cl_entities[] */
    struct model_s *model;           /* cl.model_precache[current.modelindex]; NULL: skipped */
    int    modelflags;               /* Mod_Flags (model) */
    byte  *colormap;                 /* vid.colormap, or the player's translation table */
    vec3_t prev_origin, prev_angles; int prev_frame;                     /* centity_t.prev */
    vec3_t cur_origin, cur_angles;   int cur_frame, effects, skin;       /* centity_t.current */
    float  startlerp, deltalerp, frametime, syncbase;
    vec3_t lerp_origin;              /* in: where the entity was drawn last frame (the trail's start); out: this frame's */
} rc_cent_t;
typedef struct {                     /* the globals CL_AddEntities reads */
    double time;                     /* cl.time */
    int    maxclients;               /* Com_ClientMaxclients () */
    int    viewentity;               /* cl.viewentity */
    int    chase_active, bobbingitems;   /* the cvars, as booleans */
    vec3_t viewangles;               /* cl.viewangles */
} rc_clframe_t;
rc_fx_t *RC_FxNew (int max_particles);
void  RC_FxFree (rc_fx_t *fx);
void  RC_FxClear (rc_fx_t *fx);
int   RC_FxEffect (rc_fx_t *fx, int effect, const float *a, const float *b, int i0, int i1, double time);
int   RC_FxAddParticles (rc_fx_t *fx, rc_scene_t *scene, double time, double oldtime);
int   RC_FxParticleCount (const rc_fx_t *fx);
int   RC_FxParticles (const rc_fx_t *fx, int max, float *org3, float *vel3, float *ramp, float *die, int *color, int *kind);
rc_fxlight_t *RC_FxAllocDlight (rc_fx_t *fx, int key, double time);
rc_fxlight_t *RC_FxLights (rc_fx_t *fx);                /* the MAX_DLIGHTS lights */
int   RC_FxAddDlights (rc_fx_t *fx, rc_scene_t *scene, double time, double oldtime);
int   RC_ClientAddEntities (rc_fx_t *fx, rc_scene_t *scene, rc_cent_t *cents, int n, const rc_clframe_t *frame);
/* The other two producers V_RenderView calls (cl_view.c:903-914).  Colour shifts (cl_view.c:321-550): four per client --
 * medium, damage, bonus, power-up -- sent to the scene only in frames in which one of them changed since the last call,
 * damage and bonus fading by the frame time.  Beams (cl_tent.c:60-105, 303-356): up to 24 bolts, each drawn as a chain of
 * model segments every 30 units, at most 128 segments a frame. */
void  RC_FxSetContentsColor (rc_fx_t *fx, int contents);             /* V_SetContentsColor (RC_PointContents' value) */
void  RC_FxSetEmptyColor (rc_fx_t *fx, int r, int g, int b, int percent);   /* the v_cshift command */
void  RC_FxDamage (rc_fx_t *fx, int armor, int blood);               /* the colour half of CL_ParseDamage */
void  RC_FxBonusFlash (rc_fx_t *fx);                                 /* V_BonusFlash_f */
void  RC_FxClearCshifts (rc_fx_t *fx);                               /* CL_ClearCshifts */
int   RC_FxAddCshifts (rc_fx_t *fx, rc_scene_t *scene, int items, double host_frametime);   /* CL_AddCshifts + V_CalcPowerupCshift */
const cshift_t *RC_FxCshifts (const rc_fx_t *fx);                    /* the four shifts (test / HUD access) */
int   RC_FxBeam (rc_fx_t *fx, int entity, struct model_s *model, const float *start, const float *end, double time);   /* CL_ParseBeam */
int   RC_ClientAddTempEntities (rc_fx_t *fx, rc_scene_t *scene, double time, int viewentity, const float *view_lerp_origin,
                                byte *colormap, struct model_s *const *bolts /* cl_bolt1..3_mod or NULL */);
/* The view itself: V_CalcRefdef / V_CalcIntermissionRefdef (cl_view.c:758-886) -- eye position (view height, bob, hull
 * bound, smoothed stair steps), view angles (movement roll, damage kick, idle sway, punch angle), rectangle and fov into
 * *rd, and the gun (V_AddGun) into the scene; returns the entities added (0 / 1), -1 for a bad argument (fov outside
 * 1..179: Sys_Error there).  The caller then samples the world where the eye is: RC_ClientSetViewContents (fx, rd,
 * RC_PointContents (rd->vieworg)) sets RDF_UNDERWATER and the medium's colour shift.  CL_DriftPitch (input state) and the
 * chase camera (collision traces) are the engine's. */
typedef struct {
    double time, oldtime, host_frametime;                 /* cl.time, cl.oldtime, host_frametime */
    vec3_t lerp_origin;                                   /* cl_entities[cl.viewentity].lerp_origin */
    vec3_t current_origin, current_angles;                /* ... .current (intermission only) */
    vec3_t velocity, viewangles, punchangle;              /* cl.velocity, cl.viewangles, cl.punchangle */
    float  viewheight;                                    /* cl.viewheight */
    int    onground, health, items, intermission, chase_active;   /* cl.onground, cl.stats[STAT_HEALTH], cl.items, ... */
    int    vrect[4];                                      /* scr_vrect: x, y, width, height */
    float  fov_x;                                         /* scr_fov */
    float  cl_rollangle, cl_rollspeed, cl_bob, cl_bobcycle, cl_bobup, v_kicktime;    /* the cvars (cl_view.c:33-57) */
    float  v_idlescale, v_iyaw_cycle, v_iroll_cycle, v_ipitch_cycle, v_iyaw_level, v_iroll_level, v_ipitch_level;
    struct model_s *gun_model; int gun_frame, gun_oldframe; float gun_frametime;    /* cl.viewent */
    byte  *colormap;                                      /* vid.colormap */
} rc_viewstate_t;
void  RC_FxDamageKick (rc_fx_t *fx, int armor, int blood, const float *from, const float *viewer_lerp_origin, const float *viewangles,
                       float v_kickroll, float v_kickpitch, float v_kicktime);      /* the kick half of CL_ParseDamage */
int   RC_ClientCalcRefdef (rc_fx_t *fx, rc_scene_t *scene, const rc_viewstate_t *in, refdef_t *rd);
void  RC_ClientSetViewContents (rc_fx_t *fx, refdef_t *rd, int contents);

/* Renders `n` independent views (R_RenderView semantics each, r_main.c:988-1062) into
 * packed HOST buffers: fb_out n*height*width bytes (8-bit palettised), z_out (may be NULL)
 * n*height*width int16.  Host<->device copies happen inside the call.
 * Returns RC_OK or RC_ERR_*; *status (may be NULL) receives RC_STAT_* bits. */
int  RC_RenderViews (const refdef_t *views, int n, uint8_t *fb_out, int16_t *z_out, int *status);

/* Same, outputs stay on the device (fb_dev / z_dev are CUDA device pointers owned by the
 * caller, e.g. torch tensors; z_dev may be 0).  Work is enqueued on `stream`
 * (a cudaStream_t, 0 = legacy default) and is complete when the stream reaches this point. */
int  RC_RenderViewsDevice (const refdef_t *views, int n, void *fb_dev, void *z_dev, void *stream, int *status);

/* ------------------------------------------------------------------ R_EndFrame on batches ----
 * The reference ends a frame by rebuilding the palette from the view's colour shifts and the gamma cvar and
 * handing it to the video driver (R_UpdatePalette -> VID_ShiftPalette, ref_soft/r_main.c:839-875,794-832); the
 * driver's palette then turns the 8-bit frame buffer into display colours.  For batches that is:
 *   RC_SetGamma            the `gamma` cvar (v_gamma, r_main.c:131), default 1
 *   RC_ViewPalette         the 768-byte palette R_UpdatePalette builds for one view (a pure function of the
 *                          view: where the reference skips the upload -- no colour shifts, gamma unchanged --
 *                          this is the base palette through the gamma table)
 *   RC_PresentViewsDevice  frame buffers of n views (device, as RC_RenderViewsDevice left them) -> 32-bit
 *                          pixels R | G << 8 | B << 16 | 255 << 24 (d_8to24table's layout, ref_gl/gl_vidnt.c),
 *                          each view through its own palette, on `stream`; views == NULL: the palettes of
 *                          the previous call again (nothing is uploaded)
 *   RC_PresentViews        the same from / to host memory (test and tooling path) */
/* WritePCXfile (ref_soft/pcx.c:103): the screenshot format; palette NULL = the base palette */
int  RC_WritePCX (const char *path, const uint8_t *data, int width, int height, int rowbytes, const uint8_t *palette768);
void RC_SetGamma (float gamma);
int  RC_ViewPalette (const refdef_t *view, unsigned char *pal768_out);
int  RC_PresentViewsDevice (const refdef_t *views, int n, const void *fb_dev, void *rgbx_dev, void *stream);
int  RC_PresentViews (const refdef_t *views, int n, const uint8_t *fb_in, uint32_t *rgbx_out);

/* ----------------------------------------------------------------------- 2D overlay ----
 * ref_soft/r_draw.c.  The render.h functions (client/render.h:42-53) are exported with the reference's signatures
 * for an engine build: they record one command each and R_EndFrame applies the frame's list to the engine's
 * vid.buffer on the device.  Batches pass command lists directly: */
enum { RC_D2_CHAR = 1, RC_D2_PIC, RC_D2_TRANSPIC, RC_D2_TRANSPIC_TRANSLATE, RC_D2_CONBACK, RC_D2_TILECLEAR, RC_D2_FILL, RC_D2_FADESCREEN };
typedef struct {
    int op;                   /* RC_D2_* */
    int x, y, w, h;           /* Draw_Character / Draw_Pic*: x, y; Draw_TileClear / Draw_Fill: x, y, w, h */
    int arg;                  /* character number, console lines, fill colour */
    int pic;                  /* handle from RC_DrawPicFromWad / RC_DrawCachePic */
    int table;                /* handle from RC_DrawTranslation (Draw_TransPicTranslate) */
} rc_draw2d_t;
int  RC_DrawInit (void);                                     /* Draw_Init, r_draw.c:122: conchars, disc, backtile of gfx.wad */
int  RC_DrawPicFromWad (const char *name);                   /* Draw_PicFromWad, r_draw.c:54 -> handle > 0, or RC_ERR_* */
int  RC_DrawCachePic (const char *path);                     /* Draw_CachePic, r_draw.c:72 (qpic .lmp under <basedir>/id1) */
int  RC_DrawTranslation (const unsigned char *table256);     /* a 256-byte translation table -> handle */
void RC_SetConsoleVersion (const char *text);                /* the string Draw_ConsoleBackground stamps into conback (r_draw.c:416-433) */
/* applies cmds[first[i] .. first[i+1]) in order to view i's 8-bit plane; a command the reference would Sys_Error on
 * (picture outside the screen) fails the call with RC_ERR_ARG before anything is drawn */
int  RC_DrawOverlayDevice (const rc_draw2d_t *cmds, const int *first, int n, void *fb_dev, void *stream);
int  RC_DrawOverlay (const rc_draw2d_t *cmds, const int *first, int n, uint8_t *fb_inout);

/* Pipelined form of RC_RenderViews for streams of batches (timedemo-style replay, RL roll-outs): the call
 * enqueues the batch and returns; RC_WaitViews blocks until the OLDEST outstanding call has delivered its
 * frame buffers (page-locked host memory recommended) and returns its status.  At most RC_ASYNC_CALLS (three) calls may
 * be outstanding: the device->host copy of one batch (PCIe: the longest single item) runs while the next batch is
 * rendered and a third waits behind it, so the copy engine never idles between batches.  The refdef_t array and what it points to may be reused as soon as the call returns;
 * fb_out / z_out belong to the renderer until the matching RC_WaitViews.  A status with RC_STAT_SPAN_POOL or
 * RC_STAT_CACHE_POOL or RC_STAT_AET_OVERFLOW set means the batch must be rendered again with RC_RenderViews (which
 * grows the pools / switches to the large scan kernel). */
#define RC_ASYNC_CALLS 3
int  RC_RenderViewsAsync (const refdef_t *views, int n, uint8_t *fb_out, int16_t *z_out);
int  RC_WaitViews (int *status);

/* Overlapping a consumer with the draw stage.  With a callback set, RC_RenderViewsDevice draws the
 * views of a batch in `ngroups` groups (batches with alias models, particles or under-water views are
 * still drawn in one piece); after a group's draw kernel is enqueued, `side_stream` (a cudaStream_t)
 * is made to wait for it and fn (user, first_view, num_views) is called, so the caller can enqueue
 * work that reads fb_dev[first_view .. first_view + num_views) on side_stream -- bench.py sends the
 * finished group to rank 0 over NCCL while the next group is drawn.  fn NULL switches it off. */
typedef void (*rc_group_fn) (void *user, int first_view, int num_views);
void RC_SetGroupCallback (rc_group_fn fn, void *user, void *side_stream, int ngroups);

/* Multi-GPU output placement (one process per GPU).  The only cross-GPU step of the path is
 * bringing the finished framebuffers to one rank.  Instead of a collective after the fact, a rank can
 * render straight into the root's memory: the root allocates the gather buffer with RC_DeviceAlloc and
 * exports it (CUDA IPC, 64-byte handle), every other rank imports it and passes its slice as fb_dev to
 * RC_RenderViewsDevice -- the draw kernel's aligned 8-byte stores then travel over NVLink as they are
 * produced --, or, usually faster, keeps rendering into local memory and names its slice as a MIRROR:
 * RC_SetMirrorOutput (fb, z) makes every following RC_RenderViewsDevice copy each finished group of views
 * to fb (any unified-addressing pointer: peer device memory or pinned host memory; z may be NULL) on a
 * second stream while the next group is drawn; the call's stream completes after the copies.
 * RC_SetMirrorOutput (NULL, NULL) switches it off.  Returns RC_OK or RC_ERR_CUDA. */
int  RC_DeviceAlloc (size_t bytes, void **ptr);
void RC_DeviceFree (void *ptr);
int  RC_IpcExport (void *ptr, unsigned char handle[64]);
int  RC_IpcImport (const unsigned char handle[64], void **ptr);
void RC_IpcClose (void *ptr);
void RC_SetMirrorOutput (void *fb, void *z);
/* groups of views a batch is drawn in when a mirror is set (1..8; default 2).  More groups start the copies earlier --
 * worth it when the destination's ingress is the bound (many ranks mirroring into one root) -- at the cost of one
 * more draw launch per group. */
void RC_SetMirrorGroups (int ngroups);
/* Deferred mirror: for a STREAM of batches (one RC_RenderViewsDevice per step) the copy to the mirror need not finish
 * inside the call.  With RC_SetMirrorDeferred (1) a call enqueues its kernels on `stream`, then the copy of the whole
 * batch on the renderer's copy stream, and returns; `stream` does not wait for the copy, so it runs beside the next
 * batch's kernels (the root's NVLink ingress -- (N-1) x the bytes of one rank per step -- is then busy during the whole
 * step instead of after it).  Ordering: a call that draws into a frame buffer whose copy is still in flight waits for
 * that copy first (alternate two frame buffers to keep the overlap); RC_MirrorFence (stream) makes `stream` wait for
 * every copy enqueued so far -- call it before the mirror's contents are read, e.g. at the end of a timed region. */
void RC_SetMirrorDeferred (int on);
int  RC_MirrorFence (void *stream);

/* The `loadsky <name>` command / r_skyname cvar (ref_soft/r_main.c:169-177, r_sky.c:138-194):
 * loads <basedir>/id1/env/<name>{rt,bk,lf,ft,up,dn}.pcx (six 256x256 8-bit PCX pictures); from then on
 * sky surfaces make the renderer draw the six faces of a box around the viewer instead of the
 * scrolling sky (R_EmitSkyBox, r_rast.c:165-205).  name NULL or "" switches it off.  Returns RC_OK,
 * or RC_ERR_FILE / RC_ERR_FORMAT and leaves the sky box off (the reference prints and does the same). */
int  RC_SetSkybox (const char *name);

/* Drops every cached surface block (D_FlushCaches, d_surf.c:101). */
void RC_FlushCaches (void);

/* Counters of the last RC_RenderViews* call: kernels launched, edges, surfs, spans,
 * 8-pixel blocks, surface-cache texels built, per-stage device milliseconds when
 * profiling was enabled with RC_SetProfiling(1). */
typedef struct {
    int    kernel_launches;
    long long edges, surfs, spans, blocks, cache_texels, cache_jobs, faces;
    float  ms_walk, ms_faces, ms_scan, ms_surfprep, ms_cache, ms_draw, ms_total;
    float  ms_h2d, ms_d2h;
    float  ms_spanseg;        /* part of ms_draw spent in the span-segment kernel */
    float  ms_cells;          /* part of ms_draw spent in k_draw proper (full cells, straight-line code) */
    long long fullcells;      /* screen cells covered by one span (k_draw items) */
    long long slowcells;      /* of those, cells k_draw queued for the general code (turbulent, sky, solid, warp views, negative 1/z) */
    float  ms_alias, ms_particles, ms_zresolve;   /* entity stage (profiling mode): alias models + sprites, particles, the ordered z resolve */
    float  pad_;
    long long alias_range_events; /* how often the entity rasterisers met a case the reference itself handles out of range
                               * (status bit 512, RC_STAT_ALIAS_RANGE); blocking calls only */
    long long alias_fragments;    /* profiling mode: pixels of alias-model spans the rasteriser looked at (d_polyse.c:422's z test) */
    long long alias_fragments_front; /* of those, the ones in front of the world's z-buffer (skin and colormap read, resolve plane updated) */
    long long alias_spans;        /* spans that went through the span queue */
} rc_stats_t;
void RC_GetStats (rc_stats_t *out);
void RC_SetProfiling (int on);

/* Stage readback for tests: copies the last batch's per-view records to the host.
 * kind: 0 edges (rc_edge_t), 1 surfs (rc_surf_t), 2 spans (rc_span_t, whole batch; view ignored).
 * Returns the number of records written (<= cap) or a negative error. */
int  RC_DebugRead (int kind, int view, void *dst, int cap);

/* Test hook: the byte image ref_cuda keeps of a loaded alias model -- the block Mod_LoadAliasModel assembles
 * (ref_soft/r_model.c:1349-1572) from `pheader` on, which the polygon rasteriser reads around the skin when clipped
 * triangles step off the picture.  Returns its length and copies min (length, cap) bytes. */
int  RC_DebugAliasImage (struct model_s *model, void *dst, int cap);

/* north_star naming: a refexport_t-style table over the same functions. */
typedef struct {
    int   api_version;
    void (*Init) (void);
    void (*BeginRegistration) (struct model_s *world);   /* = R_NewMap      */
    struct model_s *(*RegisterModel) (char *name);        /* = Mod_ForName   */
    void (*BeginFrame) (void);                            /* = R_BeginFrame  */
    void (*RenderFrame) (refdef_t *fd);                   /* = R_RenderView  */
    void (*EndFrame) (void);                              /* = R_EndFrame    */
    int  (*RenderFrames) (const refdef_t *fds, int n, uint8_t *fb, int16_t *z, int *status); /* = RC_RenderViews */
} refexport_t;
refexport_t GetRefAPI (int api_version);

/* The single-view render.h path draws into the engine's `vid.buffer` / `d_pzbuffer`
 * (client/vid.h:34-44, d_iface.h).  A host engine registers them here once
 * (the reference's vid driver does the same job in VID_Init, win32/vid_win.c:237-291). */
void RC_SetVidBuffers (uint8_t *buffer, int rowbytes, int16_t *zbuffer);

#ifdef __cplusplus
}
#endif
#endif
