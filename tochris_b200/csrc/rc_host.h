/*
 * rc_host.h -- host-side (plain C) declarations of ref_cuda: model loading, per-view
 * setup and the GPU pipeline entry points implemented in rc_gpu.cu.
 */
#ifndef RC_HOST_H
#define RC_HOST_H

#include "../../include/ref_cuda.h"
#include "rc_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RC_WARP_WIDTH   320    /* ref_soft/d_iface.h:22-23 */
#define RC_WARP_HEIGHT  200
#define RC_RDF_NOWORLDMODEL RDF_NOWORLDMODEL
#define RC_RDF_UNDERWATER   RDF_UNDERWATER

typedef refdef_t     rc_refdef_t;
typedef lightstyle_t rc_lightstyle_t;

typedef struct { int width, height, rowbytes; float aspect; } rc_vidinfo_t;

typedef struct {
    float r_clearcolor, r_fullbright, r_ambient, r_drawentities, r_drawviewmodel, r_waterwarp,
          r_fastsky, r_skycolor, r_lerpmodels, d_mipcap, d_mipscale, r_maxsurfs, r_maxedges, r_draworder;
} rc_cvars_t;

/* ---- models ------------------------------------------------------------------------- */
typedef enum { rc_mod_brush, rc_mod_sprite, rc_mod_alias } rc_modtype_t;

typedef struct {               /* mmodel_t, r_model.h:44-50 (+ mins/maxs spread by 1, r_model.c:642) */
    float mins[3], maxs[3];
    float radius;
    int   headnode;
    int   visleafs;
    int   firstface, numfaces;
} rc_submodel_t;

typedef struct rc_hostworld_s {
    rc_world_t w;              /* HOST pointers */
    int        numsubmodels;
    rc_submodel_t *submodels;
    size_t     texelbytes, lightbytes;
    int        sintablesize;
    void      *owned[32];      /* malloc'd blocks to free */
    int        numowned;
} rc_hostworld_t;

typedef struct {               /* maliasframedesc_t / maliasskindesc_t + group info, r_model.h:214-252 */
    int type;                  /* 0 single, 1 group */
    int first, count;          /* poses (or skin pictures) */
    int intervals;             /* index into intervals[] for groups, else -1 */
    unsigned char bboxmin[4], bboxmax[4];
} rc_aliasframe_t;

typedef struct rc_hostalias_s {  /* aliashdr_t + mdl_t, flattened */
    int   numskins, skinwidth, skinheight, numverts, numtris, numframes, synctype, flags;
    float scale[3], scale_origin[3];
    float mins[3], maxs[3];
    int   numposes, numskinpics;
    int  *stverts;             /* [numverts][4]: onseam, s<<16, t<<16, 0 */
    int  *tris;                /* [numtris][4]: facesfront, v0, v1, v2 */
    unsigned char *poseverts;  /* [numposes][numverts] trivertx_t */
    unsigned char *posebbox;   /* [numposes][8] */
    unsigned char *skinpixels; /* the reference's memory image of the loaded model, from `pheader` on (rc_alias_build_image) */
    int   skinimagelen;        /* its length */
    int  *skinpicofs;          /* [numskinpics]: offset of every skin picture inside the image */
    rc_aliasframe_t *frames, *skins;
    float *intervals;
    int   device_index;        /* slot in the device model table, -1 while not resident on the device */
    long long last_used;       /* the library's model clock when a batch last drew it / Mod_TouchModel named it (LRU order) */
} rc_hostalias_t;

typedef struct { int type; int first, count; int intervals; } rc_spriteframe_t;   /* mspriteframedesc_t: single (0) or group */
typedef struct { int width, height; float up, down, left, right; int pixofs; } rc_spritesub_t;   /* mspriteframe_t */
typedef struct rc_hostsprite_s {   /* msprite_t (r_model.h:194-203), flattened */
    int   type, maxwidth, maxheight, numframes, synctype;
    float beamlength;
    rc_spriteframe_t *frames;
    rc_spritesub_t   *subframes; int numsubframes;
    float *intervals;
    unsigned char *pixels; size_t pixelbytes;
    int   device_ofs;          /* offset of pixels[] in the device sprite blob, -1 until uploaded */
} rc_hostsprite_t;

struct model_s {
    char          name[64];
    rc_modtype_t  type;
    int           flags;
    int           numframes;
    float         mins[3], maxs[3], radius;
    /* brush */
    rc_hostworld_t *world;     /* shared by the world model and its '*n' submodels */
    int           submodel;    /* index into world->submodels */
    /* alias: filled by rc_host_alias.c */
    rc_hostalias_t *alias;
    rc_hostsprite_t *sprite;
    int           synctype;
};

int  RC_SetupView (const rc_refdef_t *rd, const rc_vidinfo_t *vid, const rc_cvars_t *cv, rc_view_t *out);
void RC_AngleVectors (const float *angles, float *forward, float *right, float *up);
void RC_ScreenEdges (const rc_view_t *v, float pixelAspect, float fov_x, float screenedge[4][3]);
void RC_TransformFrustum (const float screenedge[4][3], const float *vright, const float *vup,
                          const float *vpn, const float *modelorg, float clipnormal[4][3], float *clipdist);

#define RC_SIN_BUFFER_SIZE (RC_MAXWIDTH + RC_CYCLE + 1)   /* r_shared.h:44 with the raised maxima */
void RC_BuildTurbTables (int *sintable, int *intsintable, int n);

/* BSP29 -> flat arrays.  Mirrors Mod_LoadBrushModel (ref_soft/r_model.c:1062-1127).
 * Returns NULL and fills err on malformed data (the reference calls Sys_Error). */
rc_hostworld_t *RC_LoadBrushModel (const void *data, size_t len, const uint8_t *colormap, char *err, size_t errlen);
int  RC_ApplyLit (rc_hostworld_t *hw, const uint8_t *lit, size_t litlen);
void RC_FreeHostWorld (rc_hostworld_t *hw);

rc_hostalias_t *RC_LoadAliasModel (const void *data, size_t len, const char *modelname, char *err, size_t errlen);
void RC_FreeHostAlias (rc_hostalias_t *a);
int  RC_LightPoint (const rc_hostworld_t *hw, const int *lightstylevalue, const float *p);
int  RC_VisCullEntity (const rc_hostworld_t *hw, const float *vieworg, const float *mins, const float *maxs);
int  RC_AliasSetup (const rc_hostalias_t *a, const rc_hostworld_t *hw, const entity_t *e, const rc_refdef_t *rd,
                    const rc_view_t *v, const rc_cvars_t *cv, rc_aliasinst_t *out);

/* R_DrawBEntitiesOnList per-entity setup (r_main.c:640-768); returns 1 when the entity is drawn */
int  RC_BmodelSetup (const rc_hostworld_t *hw, int submodel, const entity_t *e, const rc_refdef_t *rd, const rc_view_t *v,
                     float pixelAspect, rc_bent_t *out);
int  RC_PointInLeaf (const rc_hostworld_t *hw, const float *p);
rc_hostsprite_t *RC_LoadSpriteModel (const void *data, size_t len, char *err, size_t errlen);
void RC_FreeHostSprite (rc_hostsprite_t *s);
int  RC_SpriteSetup (const rc_hostsprite_t *s, const entity_t *e, const rc_refdef_t *rd, const rc_view_t *v, int synctype, rc_spriteinst_t *out);
unsigned char *RC_DecodePCX (const void *data, size_t len, int *width, int *height);

/* ---- GPU side (rc_gpu.cu) ----------------------------------------------------------- */
typedef struct rc_gpu_s rc_gpu_t;

/* everything one RC_RenderViews call hands to the device, already flattened by rc_api.c */
typedef struct {
    const rc_view_t   *views;    int n;
    const rc_dlight_t *dlights;  int ndlights;
    int nwarp;                   /* views with r_dowarp (each needs a 320x200 warp buffer) */
    const rc_aliasinst_t *alias; int nalias;      /* drawn alias entities, in (view, entity) order */
    int nfinalverts, ntriwork;   /* totals over the instances */
    const rc_particle_t *particles; int nparticles;
    const uint8_t *colormaps;    int ncolormaps;  /* entity colormaps (16 KB each); 0 = vid.colormap */
    int nzres;                   /* views that need an entity z-resolve plane */
    const rc_bent_t *bents; int nbents;           /* drawn brush entities, ordered by (view, entity) */
    const rc_spriteinst_t *sprites; int nsprites; /* drawn sprite entities, ordered by (view, entity) */
    int maxbfaces;               /* largest face count among them */
} rc_batch_t;

rc_gpu_t *rc_gpu_create (int device, int width, int height, size_t cache_bytes, int max_views, char *err, size_t errlen);
void rc_gpu_destroy (rc_gpu_t *g);
/* (re)uploads the given alias models (the resident set of rc_api.c's model cache); models[i]->device_index = i */
int  rc_gpu_upload_alias (rc_gpu_t *g, rc_hostalias_t **models, int n, const float *anorm_dots, char *err, size_t errlen);
/* (re)uploads the pixels of every loaded sprite model; models[i]->device_ofs is set */
int  rc_gpu_upload_sprites (rc_gpu_t *g, rc_hostsprite_t **models, int n, char *err, size_t errlen);
int  rc_gpu_upload_world (rc_gpu_t *g, const rc_hostworld_t *hw, int maxedges, int maxsurfs, char *err, size_t errlen);
/* renders n prepared views; fb_dev/z_dev are device pointers (z may be NULL) */
int  rc_gpu_render (rc_gpu_t *g, const rc_batch_t *batch, void *fb_dev, void *z_dev, void *stream,
                    int *status, char *err, size_t errlen);
int  rc_gpu_can_split (const rc_gpu_t *g);
int  rc_gpu_render_begin (rc_gpu_t *g, void *fb_dev, void *stream);
int  rc_gpu_render_part (rc_gpu_t *g, const rc_batch_t *part, void *fb_part, void *z_part, void *stream, int first_part, char *err, size_t errlen);
int  rc_gpu_render_finish (rc_gpu_t *g, void *fb_dev, void *z_dev, int n, void *stream, int *status, char *err, size_t errlen);
int  rc_gpu_render_host (rc_gpu_t *g, const rc_batch_t *batch, uint8_t *fb_out, int16_t *z_out,
                         int *status, char *err, size_t errlen);
#define RC_ASYNC_DEPTH 3	/* RC_RenderViewsAsync calls that may be outstanding: one being copied out, one drawn, one queued behind it */
int  rc_gpu_render_host_async (rc_gpu_t *g, const rc_batch_t *batch, uint8_t *fb_out, int16_t *z_out, char *err, size_t errlen);
int  rc_gpu_wait_host (rc_gpu_t *g, int *status, char *err, size_t errlen);
void rc_gpu_flush_caches (rc_gpu_t *g);
int  rc_gpu_device_alloc (rc_gpu_t *g, size_t bytes, void **ptr);
void rc_gpu_device_free (rc_gpu_t *g, void *ptr);
int  rc_gpu_ipc_export (rc_gpu_t *g, void *ptr, unsigned char *handle);
int  rc_gpu_ipc_import (rc_gpu_t *g, const unsigned char *handle, void **ptr);
void rc_gpu_ipc_close (rc_gpu_t *g, void *ptr);
void rc_gpu_set_mirror (rc_gpu_t *g, void *fb, void *z);
void rc_gpu_set_mirror_groups (rc_gpu_t *g, int ngroups);
void rc_gpu_set_mirror_deferred (rc_gpu_t *g, int on);
int  rc_gpu_mirror_fence (rc_gpu_t *g, void *stream, char *err, size_t errlen);
void rc_gpu_set_group_callback (rc_gpu_t *g, rc_group_fn fn, void *user, void *side_stream, int ngroups);
/* pixels: six 256x256 pictures in r_skysideimage order, or NULL to switch the sky box off */
int  rc_gpu_set_skybox (rc_gpu_t *g, const uint8_t *pixels, char *err, size_t errlen);
/* pal32: n x 256 palette entries (R | G << 8 | B << 16 | 255 << 24); fb / rgbx device pointers, or host pointers with host != 0 */
int  rc_gpu_present (rc_gpu_t *g, const uint32_t *pal32, int n, const void *fb, void *rgbx, void *stream, int host, char *err, size_t errlen);
/* 2D overlay: the blob (conchars, pictures, translation tables) and picture table; then command lists per view */
struct rc_pic_s;
int  rc_gpu_upload_draw2d (rc_gpu_t *g, const uint8_t *blob, size_t len, const void *pics, int npics, char *err, size_t errlen);
int  rc_gpu_draw2d (rc_gpu_t *g, const rc_draw2d_t *cmds, int ncmds, const int *first, int n, void *fb, void *stream, int host, char *err, size_t errlen);
void rc_gpu_get_stats (rc_gpu_t *g, rc_stats_t *out);
void rc_gpu_set_profiling (rc_gpu_t *g, int on);
int  rc_gpu_debug_read (rc_gpu_t *g, int kind, int view, void *dst, int cap);

#ifdef __cplusplus
}
#endif
#endif
