/*
 * rc_host_view.c -- per-view constants, computed on the host in C with libm so that the
 * double sin/cos/tan/sqrt results are the very bits the reference gets (SURVEY.md H2:
 * libm != CUDA libdevice).  One rc_view_t replaces the globals written by
 *   R_SetupFrame           ref_soft/r_misc.c:381-478
 *   AngleVectors           common/mathlib.c:234-271
 *   R_AnimateLight         ref_soft/r_light.c:33-49
 *   R_ViewChanged          ref_soft/r_main.c:312-409
 *   D_ViewChanged          ref_soft/d_modech.c:57-109
 *   R_TransformFrustum     ref_soft/r_misc.c:279-299
 *   R_SetUpFrustumIndexes  ref_soft/r_misc.c:339-366
 *   D_SetupFrame           ref_soft/d_init.c:56-82
 * Compile with -ffp-contract=off, no -ffast-math (same as the oracle build).
 */
#include <math.h>
#include <string.h>
#include "rc_host.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static void rc_angle_vectors (const float *angles, float *forward, float *right, float *up)
{
	float angle;
	float sr, sp, sy, cr, cp, cy;

	angle = angles[0] * (M_PI * 2.0 / 360.0);
	sp = sin (angle);
	cp = cos (angle);
	angle = angles[1] * (M_PI * 2.0 / 360.0);
	sy = sin (angle);
	cy = cos (angle);
	angle = angles[2] * (M_PI * 2.0 / 360.0);
	sr = sin (angle);
	cr = cos (angle);

	forward[0] = cp * cy;
	forward[1] = cp * sy;
	forward[2] = -sp;
	right[0] = (-1 * sr * sp * cy + -1 * cr * -sy);
	right[1] = (-1 * sr * sp * sy + -1 * cr * cy);
	right[2] = -1 * sr * cp;
	up[0] = (cr * sp * cy + -sr * -sy);
	up[1] = (cr * sp * sy + -sr * cy);
	up[2] = cr * cp;
}

void RC_AngleVectors (const float *angles, float *forward, float *right, float *up)
{
	rc_angle_vectors (angles, forward, right, up);
}

static void rc_normalize (float *v)
{
	float length, ilength;
	length = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
	if (length)
	{
		length = sqrt (length);
		ilength = 1/length;
		v[0] *= ilength;
		v[1] *= ilength;
		v[2] *= ilength;
	}
}

/* R_TransformFrustum for an arbitrary basis / model-space origin (bmodels re-run it,
 * r_bsp.c:150 and r_main.c:760). */
void RC_TransformFrustum (const float screenedge[4][3], const float *vright, const float *vup,
			  const float *vpn, const float *modelorg, float clipnormal[4][3], float *clipdist)
{
	int i;
	float v[3], v2[3];
	for (i = 0; i < 4; i++)
	{
		v[0] = screenedge[i][2];
		v[1] = -screenedge[i][0];
		v[2] = screenedge[i][1];
		v2[0] = v[1]*vright[0] + v[2]*vup[0] + v[0]*vpn[0];
		v2[1] = v[1]*vright[1] + v[2]*vup[1] + v[0]*vpn[1];
		v2[2] = v[1]*vright[2] + v[2]*vup[2] + v[0]*vpn[2];
		clipnormal[i][0] = v2[0]; clipnormal[i][1] = v2[1]; clipnormal[i][2] = v2[2];
		clipdist[i] = modelorg[0]*v2[0] + modelorg[1]*v2[1] + modelorg[2]*v2[2];
	}
}

void RC_ScreenEdges (const rc_view_t *v, float pixelAspect, float fov_x, float screenedge[4][3])
{
	/* r_main.c:317,346-402 */
	float hfov = 2.0 * tan (fov_x/360*M_PI);
	float xOrigin = 1.0 / 2.0, yOrigin = 1.0 / 2.0;
	float screenAspect = v->vrect_w*pixelAspect / v->vrect_h;
	float vfov = hfov / screenAspect;
	int i;
	screenedge[0][0] = -1.0 / (xOrigin*hfov); screenedge[0][1] = 0; screenedge[0][2] = 1;
	screenedge[1][0] = 1.0 / ((1.0-xOrigin)*hfov); screenedge[1][1] = 0; screenedge[1][2] = 1;
	screenedge[2][0] = 0; screenedge[2][1] = -1.0 / (yOrigin*vfov); screenedge[2][2] = 1;
	screenedge[3][0] = 0; screenedge[3][1] = 1.0 / ((1.0-yOrigin)*vfov); screenedge[3][2] = 1;
	for (i = 0; i < 4; i++)
		rc_normalize (screenedge[i]);
}

int RC_SetupView (const rc_refdef_t *rd, const rc_vidinfo_t *vid, const rc_cvars_t *cv, rc_view_t *out)
{
	int i, j;
	float hfov, pixelAspect;
	float screenedge[4][3];
	float basemip[3] = {1.0, 0.5*0.8, 0.25*0.8};
	rc_view_t *v = out;

	memset (v, 0, sizeof(*v));
	v->rdflags = rd->flags & 0xFF;
	if (cv->r_draworder)
		v->rdflags |= RC_VIEW_DRAWORDER;		/* r_edge.c:94-105 */
	v->ambientlight = (int)cv->r_ambient;			/* r_misc.c:415 */
	if (v->ambientlight < 0)
		v->ambientlight = 0;
	v->fullbright = cv->r_fullbright != 0;
	v->clearcolor = (int)cv->r_clearcolor & 0xFF;
	v->skycolor = (int)cv->r_skycolor & 0xFF;
	v->fastsky = cv->r_fastsky != 0;

	for (i = 0; i < 3; i++)
		v->origin[i] = rd->vieworg[i];
	rc_angle_vectors (rd->viewangles, v->vpn, v->vright, v->vup);

	/* R_AnimateLight, r_light.c:40-48 */
	if (!(rd->flags & RC_RDF_NOWORLDMODEL) && rd->lightstyles)
	{
		int t = (int)(rd->time*10);
		for (j = 0; j < RC_MAX_LIGHTSTYLES; j++)
		{
			const rc_lightstyle_t *ls = &rd->lightstyles[j];
			if (!ls->length)
			{
				v->lightstylevalue[j] = 256;
				continue;
			}
			v->lightstylevalue[j] = (ls->map[t % ls->length] - 'a') * 22;
		}
	}
	{
		/* R_SetSkyFrame, r_sky.c:112-130: iskyspeed 8, iskyspeed2 2 -> gcd 2, s1 4, s2 1 */
		float skyspeed = 8, temp, skytime;
		temp = 128 * 4 * 1;
		skytime = rd->time - ((int)(rd->time / temp) * temp);
		v->skyshift = skytime * skyspeed;
	}
	v->animtime = (int)(rd->time*10);
	v->turbphase = (int)(rd->time*RC_TURB_SPEED) & (RC_CYCLE-1);

	v->dowarp = (cv->r_waterwarp != 0) && (rd->flags & RC_RDF_UNDERWATER);
	v->warp_index = -1;
	v->rd_x = rd->x; v->rd_y = rd->y; v->rd_w = rd->width; v->rd_h = rd->height;
	if (rd->x < 0 || rd->y < 0 || rd->width <= 0 || rd->height <= 0 ||
	    rd->x + rd->width > vid->width || rd->y + rd->height > vid->height)
		return RC_ERR_ARG;
	if (v->dowarp)
	{	/* r_misc.c:448-454 */
		v->vrect_x = 0;
		v->vrect_y = 0;
		v->vrect_w = (rd->width < RC_WARP_WIDTH) ? rd->width : RC_WARP_WIDTH;
		v->vrect_h = (rd->height < RC_WARP_HEIGHT) ? rd->height : RC_WARP_HEIGHT;
	}
	else
	{
		v->vrect_x = rd->x;
		v->vrect_y = rd->y;
		v->vrect_w = rd->width;
		v->vrect_h = rd->height;
	}
	if (v->vrect_w <= 0 || v->vrect_h <= 0 || v->vrect_x < 0 || v->vrect_y < 0 ||
	    (!v->dowarp && (v->vrect_x + v->vrect_w > vid->width || v->vrect_y + v->vrect_h > vid->height)) ||
	    vid->width > RC_MAXWIDTH || vid->height > RC_MAXHEIGHT)
		return RC_ERR_ARG;

	/* R_ViewChanged, r_main.c:316-374 */
	pixelAspect = vid->aspect;
	hfov = 2.0 * tan (rd->fov_x/360*M_PI);
	v->fvrectx_adj = (float)v->vrect_x - 0.5;
	v->vrect_x_adj_shift20 = (v->vrect_x<<20) + (1<<19) - 1;
	v->fvrecty_adj = (float)v->vrect_y - 0.5;
	v->vrectright = v->vrect_x + v->vrect_w;
	v->vrectright_adj_shift20 = (v->vrectright<<20) + (1<<19) - 1;
	v->fvrectright_adj = (float)v->vrectright - 0.5;
	v->vrectbottom = v->vrect_y + v->vrect_h;
	v->fvrectbottom_adj = (float)v->vrectbottom - 0.5;

	v->xcenter = ((float)v->vrect_w * (1.0 / 2.0)) + v->vrect_x - 0.5;
	v->aliasxcenter = v->xcenter;
	v->ycenter = ((float)v->vrect_h * (1.0 / 2.0)) + v->vrect_y - 0.5;
	v->aliasycenter = v->ycenter;
	v->xscale = v->vrect_w / hfov;
	v->aliasxscale = v->xscale;
	v->xscaleinv = 1.0 / v->xscale;
	v->yscale = v->xscale * pixelAspect;
	v->aliasyscale = v->yscale;
	v->yscaleinv = 1.0 / v->yscale;
	v->xscaleshrink = (v->vrect_w-6)/hfov;
	v->yscaleshrink = v->xscaleshrink*pixelAspect;

	RC_ScreenEdges (v, pixelAspect, rd->fov_x, screenedge);

	/* D_ViewChanged, d_modech.c:61-91 */
	v->scale_for_mip = v->xscale;
	if (v->yscale > v->xscale)
		v->scale_for_mip = v->yscale;
	v->zwidth = vid->width;
	v->d_pix_min = v->vrect_w / 320;
	if (v->d_pix_min < 1)
		v->d_pix_min = 1;
	v->d_pix_max = (int)((float)v->vrect_w / (320.0 / 4.0) + 0.5);
	v->d_pix_shift = 8 - (int)((float)v->vrect_w / 320.0 + 0.5);
	if (v->d_pix_max < 1)
		v->d_pix_max = 1;
	if (pixelAspect > 1.4)
		v->d_y_aspect_shift = 1;
	else
		v->d_y_aspect_shift = 0;
	v->d_vrectx = v->vrect_x;
	v->d_vrecty = v->vrect_y;
	v->d_vrectright_particle = v->vrectright - v->d_pix_max;
	v->d_vrectbottom_particle = v->vrectbottom - (v->d_pix_max << v->d_y_aspect_shift);

	/* R_TransformFrustum + R_SetUpFrustumIndexes, r_misc.c:279-366 */
	RC_TransformFrustum (screenedge, v->vright, v->vup, v->vpn, v->origin, v->clipnormal, v->clipdist);
	for (i = 0; i < 4; i++)
		for (j = 0; j < 3; j++)
		{
			if (v->clipnormal[i][j] < 0)
			{
				v->frustum_idx[i][j] = j;
				v->frustum_idx[i][j+3] = j+3;
			}
			else
			{
				v->frustum_idx[i][j] = j+3;
				v->frustum_idx[i][j+3] = j;
			}
		}

	/* TransformVector (modelorg, transformed_modelorg), d_edge.c:173 */
	v->tmodelorg[0] = v->origin[0]*v->vright[0] + v->origin[1]*v->vright[1] + v->origin[2]*v->vright[2];
	v->tmodelorg[1] = v->origin[0]*v->vup[0] + v->origin[1]*v->vup[1] + v->origin[2]*v->vup[2];
	v->tmodelorg[2] = v->origin[0]*v->vpn[0] + v->origin[1]*v->vpn[1] + v->origin[2]*v->vpn[2];

	/* D_SetupFrame, d_init.c:60-80 */
	v->screenwidth = v->dowarp ? RC_WARP_WIDTH : vid->rowbytes;
	v->d_minmip = (int)cv->d_mipcap;
	if (v->d_minmip > 3)
		v->d_minmip = 3;
	else if (v->d_minmip < 0)
		v->d_minmip = 0;
	for (i = 0; i < 3; i++)
		v->d_scalemip[i] = basemip[i] * cv->d_mipscale;
	return RC_OK;
}

/* R_InitTurb, ref_soft/r_main.c:1069-1078 (note the literal 3.14159).  n = SIN_BUFFER_SIZE. */
void RC_BuildTurbTables (int *sintable, int *intsintable, int n)
{
	int i;
	for (i = 0; i < n; i++)
	{
		sintable[i] = (8*0x10000) + sin (i*3.14159*2/RC_CYCLE)*(8*0x10000);
		intsintable[i] = 3 + sin (i*3.14159*2/RC_CYCLE)*3;
	}
}
