/*
 * rc_types.h -- plain-old-data records shared by the C host side and the CUDA kernels of
 * ref_cuda.  Everything here is laid out for HBM: flat arrays indexed by int, no pointers
 * inside records, 16/32-byte records so a lane fetches a record with one or two vector loads.
 *
 * Each record names the reference structure it replaces (paths relative to /root/reference).
 */
#ifndef RC_TYPES_H
#define RC_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- limits (ref_soft/r_shared.h:40-41,110-114; raised where the reference needs a
 *      recompile) ------------------------------------------------- */
#define RC_MAXWIDTH        2047
#define RC_MAXHEIGHT       1200
#define RC_MAX_LIGHTSTYLES 64      /* common/quakedef.h:105 */
#define RC_MAXLIGHTMAPS    4
#define RC_MIPLEVELS       4
#define RC_NEAR_CLIP       0.01    /* r_local.h:205 */
#define RC_BACKFACE_EPSILON 0.01   /* r_local.h:68 */
#define RC_CYCLE           128     /* d_iface.h turbulence period */
#define RC_TURB_SPEED      20      /* r_local.h:242 */

/* surface flags, ref_soft/r_model.h:72-78 */
#define RC_SURF_PLANEBACK      2
#define RC_SURF_DRAWSKY        4
#define RC_SURF_DRAWSPRITE     8
#define RC_SURF_DRAWTURB       0x10
#define RC_SURF_DRAWTILED      0x20
#define RC_SURF_DRAWBACKGROUND 0x40
#define RC_SURF_DRAWSKYBOX     0x80

#define RC_CONTENTS_SOLID (-2)
#define RC_SKYBOX_BYTES   (6 * 256 * 256)   /* r_skypixels, r_sky.c:135: reserved at the start of the surface-cache pool */
#define RC_SKYBOX_KEY     0x7ffffff0        /* r_rast.c:199 */

/* ---- world model (replaces model_t's brush arrays, ref_soft/r_model.h:290-340) ------ */
typedef struct {            /* cplane_t, common/mathlib.h */
    float normal[3];
    float dist;
    int   type;
    int   pad[3];
} rc_plane_t;               /* 32 B */

typedef struct {            /* mnode_t, r_model.h:118-134.  child >= 0: node, < 0: leaf -1-child */
    int      plane;
    int      children[2];
    int16_t  minmaxs[6];
    uint16_t firstsurface, numsurfaces;
    int      parent;        /* node index or -1 */
} rc_node_t;                /* 32 B */

typedef struct {            /* mleaf_t, r_model.h:138-153 */
    int     contents;
    int16_t minmaxs[6];
    int     firstmark, nummarks;
    int     parent;
    int     pad;
} rc_leaf_t;                /* 32 B */

typedef struct {            /* msurface_t, r_model.h:98-116 (static part) */
    int      plane;
    int      flags;
    int      firstedge, numedges;
    int16_t  texturemins[2];
    int16_t  extents[2];
    int      texinfo;
    uint8_t  styles[4];
    int      lightofs;      /* offset into lightdata or -1 */
    int      node;          /* owning node (load-time derived) */
    int      nodeslot;      /* index among the node's faces: face - nodes[node].firstsurface */
    int      pad;
} rc_face_t;                /* 48 B */

typedef struct {            /* mtexinfo_t, r_model.h:90-96 */
    float vecs[2][4];
    float mipadjust;
    int   texture;
    int   flags;
    int   pad;
} rc_texinfo_t;             /* 48 B */

typedef struct {            /* texture_t, r_model.h:63-72; offsets index the shared texel blob */
    int width, height;
    int offsets[4];
    int anim_total, anim_min, anim_max;
    int anim_next;          /* texture index or -1 */
    int alternate_anims;    /* texture index or -1 */
    int pad;
} rc_texture_t;             /* 48 B */

/* static per-node data for the parallel BSP walk: a pre-order interval that tells which child
 * holds a given node/leaf, and the number of sort-key slots in each child's subtree */
typedef struct {
    int dfs_in, dfs_mid;    /* pre-order number; numbers < dfs_mid lie under children[0], the rest under children[1] */
    int slots[2];           /* key slots under children[0] / children[1] */
} rc_nodeaux_t;             /* 16 B */

/* per-view result of visiting a node (r_bsp.c:436-557), consumed by the face-list kernel */
typedef struct {
    int16_t  state;         /* 0: not visited, 1: view in front (dot > eps), 2: behind (dot < -eps), 3: on plane */
    int16_t  clipflags;
    int      facekey;       /* sort key of the node's first face */
} rc_nodevisit_t;           /* 8 B */

/* one entry per surfedge: who else uses the same medge_t (r_rast.c:583-603 edge cache) */
typedef struct {
    int partner_face;       /* -1: medge used by this face only */
    int partner_slot;       /* index of the medge inside the partner's surfedge run */
} rc_separtner_t;

/* one per surfedge, flattened at load time so the face kernel reads an edge with two vector loads
 * instead of chasing surfedges -> medges -> vertexes (three dependent levels, twice when the
 * partner's use of the edge has to be re-clipped) */
typedef struct {
    float v0[3];            /* first vertex in this face's winding direction (r_rast.c:560-575) */
    int   partner_face;     /* as rc_separtner_t, -1: none */
    float v1[3];
    int   partner_rev;      /* 1: the partner walks the medge in the opposite direction */
} rc_fedge_t;               /* 32 B */

typedef struct {            /* device-resident world; all pointers are device pointers */
    int numplanes, numnodes, numleafs /* incl. leaf 0 */, numfaces, numsurfedges, numedges,
        numvertexes, numtexinfo, numtextures, nummarksurfaces, visleafs, haslight;
    int numworldfaces;      /* faces of model 0 (without skybox faces) */
    int skytexofs;          /* texel offset of the 256x128 sky texture R_InitSky copied last (r_sky.c:53), -1: none */
    int marksplit;          /* 1: markleaf[] is valid (no two leaves share a marksurfaces entry) */
    int skyfirst;           /* first of the six sky-box faces R_InitSkyBox appends (r_rast.c:108-157); they are the last six of faces[] */
    int drawskybox;         /* r_drawskybox: a sky box is loaded (r_sky.c:189-194); its six 256x256 pictures sit at the start of the cache pool */
    int visrowwords;        /* words per row of nodevis */
    const rc_plane_t     *planes;
    const rc_node_t      *nodes;
    const rc_leaf_t      *leafs;
    const rc_face_t      *faces;
    const int            *surfedges;
    const rc_separtner_t *separtner;
    const rc_fedge_t     *fedges;      /* [numsurfedges] */
    const uint16_t       *edges;       /* 2 x uint16 per medge_t */
    const float          *vertexes;    /* 3 floats per mvertex_t */
    const rc_texinfo_t   *texinfo;
    const rc_texture_t   *textures;
    const uint8_t        *texels;      /* all miptex pixel data */
    const uint8_t        *lightdata;
    const int            *marksurfaces;
    const int            *markleaf;    /* [nummarksurfaces]: the leaf that owns each entry, -1: none */
    const uint32_t       *nodevis;     /* [numleafs][visrowwords]: bit i<numnodes: node i, else leaf */
    const uint8_t        *colormap;    /* vid.colormap, 64*256 */
    const int            *sintable;    /* r_rast.c:45, SIN_BUFFER_SIZE ints */
    const int            *intsintable;
    const rc_nodeaux_t   *nodeaux;     /* [numnodes] */
    const int            *leafdfs;     /* [numleafs] pre-order number of each leaf (-1: not in the world tree) */
} rc_world_t;

/* ---- per-view constants (host computed; replaces the globals written by R_SetupFrame,
 *      R_ViewChanged, D_ViewChanged, D_SetupFrame, R_TransformFrustum,
 *      R_SetUpFrustumIndexes, R_AnimateLight; see rc_host_view.c) --------------------- */
typedef struct {
    int   vrect_x, vrect_y, vrect_w, vrect_h, vrectright, vrectbottom;
    float fvrectx_adj, fvrecty_adj, fvrectright_adj, fvrectbottom_adj;
    int   vrect_x_adj_shift20, vrectright_adj_shift20;
    float xcenter, ycenter, xscale, yscale, xscaleinv, yscaleinv, xscaleshrink, yscaleshrink;
    float vpn[3], vright[3], vup[3], origin[3];
    float tmodelorg[3];            /* TransformVector(modelorg), d_edge.c:173 */
    float clipnormal[4][3];
    float clipdist[4];
    int   frustum_idx[4][6];
    float scale_for_mip;
    float d_scalemip[3];
    int   d_minmip;
    int   ambientlight;
    int   animtime;                /* (int)(time*10)   r_surf.c:72 */
    int   turbphase;               /* (int)(time*SPEED)&(CYCLE-1)  d_scan.c:128 */
    int   rdflags;          /* refdef_t.flags (RDF_*) | RC_VIEW_DRAWORDER */
    int   dowarp;
    int   clearcolor;              /* (int)r_clearcolor.value & 0xFF */
    int   fullbright;              /* r_fullbright.value != 0 */
    int   screenwidth;             /* row pitch of the colour target in bytes */
    int   zwidth;                  /* d_zwidth (= vid.width) */
    int   d_pix_min, d_pix_max, d_pix_shift, d_y_aspect_shift;
    int   d_vrectx, d_vrecty, d_vrectright_particle, d_vrectbottom_particle;
    float aliasxscale, aliasyscale, aliasxcenter, aliasycenter;
    int   lightstylevalue[RC_MAX_LIGHTSTYLES];
    int   skycolor;                /* (int)r_skycolor.value & 0xFF */
    int   fastsky;                 /* r_fastsky.value != 0 */
    float skyshift;                /* R_SetSkyFrame, r_sky.c:112-130 */
    int   num_dlights;             /* refdef.num_dlights (<= 32) */
    int   dlight_ofs;              /* first entry of this view in the batch's dlight array */
    int   warp_index;              /* r_dowarp: which 320x200 warp buffer this view renders into, else -1 */
    int   rd_x, rd_y, rd_w, rd_h;  /* r_refdef rectangle (D_WarpScreen target) */
    int   zres_index;              /* index of this view's entity z-resolve plane, -1: no alias models / particles */
    int   viewleaf;                /* Mod_PointInLeaf (r_model.c:83) of the view origin, done on the host; -1: no world */
} rc_view_t;

/* dlight_t (client/ref.h:43-47) */
typedef struct { float origin[3]; float radius; } rc_dlight_t;

/* ---- per-view transient records ------------------------------------------------------ */
typedef struct {            /* one rendered face: output of the BSP walk (r_bsp.c:436) */
    int face;
    int clipflags;
    int key;                /* order-isomorphic to r_currentkey (r_rast.c:699) */
    int pad;
} rc_facerec_t;             /* 16 B */

/* edge_t replacement (r_shared.h:199-208): 32 B, no pointers.  `rank` reproduces the
 * newedges[] insertion order (r_rast.c:355-379): bit31 = emitted as trailing edge,
 * bits 5..30 = emission order of the face (list position), bits 0..4 = slot in face. */
typedef struct {
    int      u, u_step;     /* 12.20 */
    int16_t  v, v2;         /* first / last scanline */
    uint16_t surfs[2];
    uint32_t rank;
    int      pad[3];
} rc_edge_t;

typedef struct {            /* surf_t replacement (r_shared.h:125-147) + drawing state; 128 B */
    /* -- scan / prep fields (first 64 B) -- */
    int   key;
    int   flags;
    int   face;             /* world face index, -1 for the background */
    int   insubmodel;
    float nearzi;
    float d_ziorigin, d_zistepu, d_zistepv;
    int   miplevel;         /* scratch: cache-table slot to resolve / right-clip flags of the face stage */
    int   hasspans;
    int   entity;
    int   pad0;
    float d_sdivzstepv, d_tdivzstepv, d_sdivzorigin, d_tdivzorigin;
    /* -- draw fields: three 16-byte groups the draw kernel fetches with vector loads,
     *    filled by the surface-prep kernel (d_edge.c:96-135,307) -- */
    int   cacheofs;         /* byte offset of the texel block (cache pool or texel blob) */
    int   cachewidth;       /* row pitch of the block; constant colour for RC_SRC_SOLID */
    int   pad;              /* texel source, RC_SRC_* */
    int   drawflags;        /* copy of flags */
    float d_sdivzstepu, d_tdivzstepu, d_zistepu_draw;
    int   izistep;          /* (int)(d_zistepu * 0x8000 * 0x10000), d_scan.c:399 */
    int   sadjust, tadjust, bbextents, bbextentt;
    int   pad2[4];
} rc_surf_t;                /* 128 B */

typedef struct {            /* espan_t replacement (r_shared.h:117-121) + everything the draw kernel needs of the span's
                             * surface; 32 B, two vector loads.  Written by the scan stage (u, count, pixofs, segbase, gsurf,
                             * v / warp bits of cwk) and completed by the segment stage (izi0, izistep, cacheofs, cwk) */
    int16_t  u, count;
    int      pixofs;        /* (view * height + v) * width + u: index of the first pixel in the batch's colour / z planes */
    int      segbase;       /* first 64-pixel segment record of the span (8 subdivision records each, see rc_span_seg) */
    int      gsurf;         /* view * surfcap + surf_t slot */
    /* -- second 16 bytes: all the draw kernel's full-cell path reads -- */
    int      iziorg;        /* D_DrawZSpans fixed-point 1/z (d_scan.c:411-413) extrapolated to pixel index 0 of the planes:
                             * izi(pixel p) = iziorg + p * izistep in wrapping 32-bit arithmetic, as the reference's adds */
    int      izistep;       /* (int)(d_zistepu * 0x8000 * 0x10000), d_scan.c:399 */
    int      cacheofs;      /* byte offset of the surface's texel block (cache pool or texel blob) */
    uint32_t cwk;           /* bits 0-15 row pitch of the block (constant colour for RC_KIND_SOLID), 16-18 RC_KIND_*,
                             * 19 the view renders into its 320x200 warp buffer (r_dowarp), 20-31 scanline v */
} rc_span_t;
#define RC_KIND_POOL   0    /* surface-cache block, D_DrawSpans8 */
#define RC_KIND_TEXELS 1    /* texel blob (sky-box pictures), D_DrawSpans8 */
#define RC_KIND_TURB   2    /* texel blob, Turbulent8 */
#define RC_KIND_SKY    3    /* D_DrawSkyScans8 */
#define RC_KIND_SOLID  4    /* D_DrawSolidSurface */
#define RC_SPAN_WARPBIT 0x80000u

typedef struct {            /* one 8-pixel subdivision of a span (d_scan.c:278-365): 16.16 texture coordinates of its
                             * first pixel and the per-pixel step the reference uses inside it (in registers; what is
                             * stored are the boundary records below) */
    int s, t, sstep, tstep;
} rc_sub_t;

typedef struct {            /* 8 B: the 16.16 (s, t) at one subdivision boundary of a span.  A span with nsub subdivisions
                             * owns records [segbase * 8 .. + nsub + 1]: B[0] .. B[nsub] (B[nsub] = the last pixel's
                             * coordinates), then one record holding the per-pixel step of the LAST subdivision (the one
                             * step that is a truncating division, d_scan.c:364-365; every other step is
                             * (B[k+1] - B[k]) >> 3, :334-335, which the draw kernels take from two neighbouring records) */
    int s, t;
} rc_bnd_t;

/* cell records flag where a cell touches the span's last subdivision, so that the draw kernel knows which of its two
 * steps to read instead of derive */
#define RC_VIEW_DRAWORDER 0x100        /* rc_view_t.rdflags: r_draworder is set (the farthest surface wins: R_GenerateSpansBackward) */
#define RC_CELL_L0     0x40000000          /* the subdivision of the cell's first pixel is the span's last */
#define RC_CELL_L1     0x80000000u         /* the following one is */
#define RC_CELL_SPAN(x) ((int)((x) & 0x3FFFFFFF))

typedef struct {            /* draw work item: a screen-aligned 8-pixel cell covered by one span */
    int      si;            /* span (| RC_CELL_L0 | RC_CELL_L1), -1: none (the cell is drawn by span heads / tails, or lies outside the view rectangle) */
    uint32_t sub;           /* (segbase * 8 + k) << 3 | a: the subdivision record of the cell's first pixel and its position in it */
} rc_cell_t;

/* surface-cache build job (d_surf.c:196 D_CacheSurface miss) */
typedef struct {
    int face, mip, texture;
    int lightadj[4];
    int ambient;
    int dstofs;             /* into the cache pool */
    int width, height;      /* block size in texels */
    int view;               /* for dynamic lights (-1: none) */
    int pad;                /* bit 0: r_fullbright */
    int dlightbits;
    int pad2[3];            /* [0] first dlight, [1] how many */
    /* what the build needs of the face and the texture, looked up once by the prep kernel (the build kernel's units
     * are short chains of dependent loads: every level taken out of them counts) */
    int lightofs;           /* face lightofs (-1: none) */
    uint8_t styles[4];
    int smax, tmax;         /* luxels per row / rows: (extents >> 4) + 1 */
    int tw, th;             /* texture size at this mip */
    int soffset, toffset;   /* ((texturemins >> mip) + (tw << 16)) % tw: r_surf.c:130-140 */
    int srcofs;             /* first texel of the mip in the texel blob */
    int bandrows;           /* rows per work unit of the build kernel (rc_cache_band_rows) */
    int items;              /* 16-texel items per row: (width + 15) >> 4 */
    int qG, rG;             /* 32 / items, 32 % items: how a warp's lanes step through the items */
    int dsx, wsx, dsy;      /* (16 rG) % tw, (16 items) % tw, qG % th: the source position's steps */
} rc_cachejob_t;            /* 128 B */

/* ---- alias models (MDL): ref_soft/r_alias.c, r_aclip.c, d_polyse.c ------------------------ */
typedef struct {            /* one loaded model's slices of the shared device blobs */
    int numverts, numtris, skinwidth, skinheight;
    int stverts_ofs;        /* into stverts[] (int4: onseam, s<<16, t<<16, 0) */
    int tris_ofs;           /* into tris[] (int4: facesfront, v0, v1, v2) */
    int poses_ofs;          /* into poseverts[] (trivertx_t, 4 B each); pose p at poses_ofs + p*numverts */
    int skins_ofs;          /* into skinpixels[]: start of the model's memory image (the reference's cache block from `pheader` on,
                             * rc_host_alias.c rc_alias_build_image); rc_aliasinst_t.skin is the picture's offset inside it */
    int skins_len;          /* length of that image: skin reads that stay inside it are the reference's own out-of-picture reads */
    int pad[3];
} rc_aliasmodel_t;          /* 48 B */

typedef struct {            /* one drawn alias entity of one view (host-prepared: R_DrawAliasModel, r_alias.c:734-788) */
    int   view;
    int   model;            /* index into the device model table */
    int   trivial_accept;   /* R_AliasCheckBBox result (r_alias.c:135-296) */
    int   order;            /* draw order of the entity inside its view (1-based) */
    float xf[3][4];         /* aliastransform, pre-scaled when trivially accepted (r_alias.c:367-414) */
    int   ambientlight;     /* r_ambientlight after R_AliasSetupLighting */
    float shadelight;
    int   shadequant;       /* row of r_avertexnormal_dots */
    float framelerp;
    int   pose_old, pose_new;
    int   skin;             /* offset of the skin picture inside the model's image (rc_aliasmodel_t.skins_ofs) */
    float ziscale;
    int   colormap;         /* index of the entity colormap in the batch's colormap table (0 = vid.colormap) */
    int   vertbase;         /* first finalvert of this instance in the batch pool */
    int   tribase;          /* first work item of this instance in the triangle work list */
    int   pad;
} rc_aliasinst_t;           /* 112 B */

typedef struct {            /* finalvert_t, d_iface.h:37-41 (+ the auxvert_t view-space point) */
    int   v[6];             /* u, v, s, t, l, 1/z */
    int   flags;
    float av[3];            /* auxvert_t.fv, r_local.h:37-39 (clipped path only) */
    int   pad[2];
} rc_finalvert_t;           /* 48 B */

typedef struct {            /* one drawn sprite entity of one view (host-prepared: R_DrawSprite, r_sprite.c:288-403) */
    int   view;
    int   order;            /* draw order of the entity inside its view (1-based, shared with alias entities) */
    int   pixofs;           /* the chosen frame's pixels in the device sprite blob */
    int   width, height;    /* mspriteframe_t (r_model.h:172-179) */
    float up, down, left, right;
    float vpn[3], vright[3], vup[3];   /* r_spritedesc axes for the sprite's orientation type */
    float entorigin[3];     /* r_entorigin after R_RotateSprite */
    float modelorg[3];      /* modelorg after R_RotateSprite */
    int   pad[4];
} rc_spriteinst_t;          /* 112 B */

typedef struct {            /* particle_t (client/ref.h:37-41) + owning view */
    float org[3];
    int   color;
    int   view;
    int   order;            /* index within the view */
    int   pad[2];
} rc_particle_t;            /* 32 B */

/* the part of rc_view_t that R_RotateBmodel (r_bsp.c:78-151) replaces while a brush entity is
 * drawn; laid out exactly like rc_view_t from `vpn` to `clipdist` so a view can be used as one */
typedef struct {
    float vpn[3], vright[3], vup[3], origin[3];   /* origin = modelorg */
    float tmodelorg[3];
    float clipnormal[4][3];
    float clipdist[4];
} rc_xform_t;               /* 31 floats */

typedef struct {            /* one drawn brush entity of one view (R_DrawBEntitiesOnList, r_main.c:640-768) */
    int   view, order;      /* order: index among the view's brush entities */
    int   firstface, numfaces;
    int   headnode;         /* of the submodel (R_MarkLights) */
    int   topnode;          /* R_FindTopNode: >= 0 node (clip to the world BSP), < 0 leaf -1-topnode */
    int   clipflags;
    int   frame;            /* entity->frame: alternate texture animation (r_surf.c:65-69) */
    int   dlight_ofs;       /* this entity's model-space copies of the view's dlights, -1: none */
    int   pad[2];
    float origin[3];        /* r_entorigin */
    float rot[3][3];        /* entity_rotation */
    rc_xform_t x;           /* rotated basis, modelorg, frustum; tmodelorg as D_DrawSurfaces computes it */
} rc_bent_t;                /* 220 B */

/* error / status bits returned through the C-ABI */
#ifndef RC_OK
#define RC_OK                    0
#define RC_ERR_CUDA              (-1)
#define RC_ERR_ARG               (-2)
#define RC_ERR_NOWORLD           (-3)
#define RC_ERR_FILE              (-4)
#define RC_ERR_FORMAT            (-5)
#define RC_ERR_NOMEM             (-6)
#define RC_STAT_EDGE_OVERFLOW    1   /* a view produced more edges than r_maxedges: the reference drops faces (r_rast.c:544) */
#define RC_STAT_SURF_OVERFLOW    2   /* ditto r_maxsurfs (r_rast.c:537) */
#define RC_STAT_SPAN_POOL        4   /* internal span pool was too small (retried by the host) */
#define RC_STAT_AET_OVERFLOW     8   /* more active edges on one scanline than the scan kernel holds */
#define RC_STAT_STACK_OVERFLOW   16  /* active surface stack deeper than the scan kernel holds */
#define RC_STAT_CACHE_POOL       32  /* surface cache pool too small (retried by the host) */
#define RC_STAT_STALE_CLIPVERT   64  /* reference would read r_rightenter/r_rightexit left over from an earlier face */
#define RC_STAT_FACELIST         128 /* face list capacity exceeded */
/* bit 256 is not used (it flagged entity kinds that were skipped before sprites and brush entities were built) */
#define RC_STAT_ALIAS_RANGE       512  /* the entity rasterisers hit a case where the reference itself reads or loops out of range (clipped
                                        * triangles stepping off the skin, d_polyse.c:385-447; span packages left by an earlier triangle;
                                        * a particle 1/z outside 16 bits): those pixels are undefined in the reference.  rc_stats_t.alias_range_events counts them */
#define RC_STAT_BMODEL_OVERFLOW   1024 /* brush entity clipping ran out of vertices/edges (reference: MAX_BMODEL_VERTS/EDGES) */
#endif

#ifdef __cplusplus
}
#endif
#endif
