/*
 * rc_host_bmodel.c -- per (view, brush entity) setup on the host: the part of
 * R_DrawBEntitiesOnList (ref_soft/r_main.c:640-768) that runs before any face is touched:
 * R_BmodelCheckBBox (:539-595), R_FindTopNode (:602-633), R_RotateBmodel (r_bsp.c:78-151,
 * libm sin/cos) and the model-space copy of the view frustum.  Faces are clipped and emitted
 * on the device (rc_bmodel_face in rc_pipeline.h).  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <string.h>
#include "rc_host.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define BMODEL_FULLY_CLIPPED 0x10

int RC_PointInLeaf (const rc_hostworld_t *hw, const float *p)
{
	const rc_world_t *w = &hw->w;
	int n = hw->submodels[0].headnode;
	while (n >= 0)
	{
		const rc_node_t *nd = &w->nodes[n];
		const rc_plane_t *pl = &w->planes[nd->plane];
		float d = (p[0]*pl->normal[0] + p[1]*pl->normal[1] + p[2]*pl->normal[2]) - pl->dist;
		n = (d > 0) ? nd->children[0] : nd->children[1];
	}
	return -1 - n;
}

static void concat_rotations (float in1[3][3], float in2[3][3], float out[3][3])
{
	int i, j;
	for (i = 0; i < 3; i++)
		for (j = 0; j < 3; j++)
			out[i][j] = in1[i][0] * in2[0][j] + in1[i][1] * in2[1][j] + in1[i][2] * in2[2][j];
}

static void entity_rotate (float rot[3][3], float *vec)
{
	float t[3];
	t[0] = vec[0]; t[1] = vec[1]; t[2] = vec[2];
	vec[0] = rot[0][0]*t[0] + rot[0][1]*t[1] + rot[0][2]*t[2];
	vec[1] = rot[1][0]*t[0] + rot[1][1]*t[1] + rot[1][2]*t[2];
	vec[2] = rot[2][0]*t[0] + rot[2][1]*t[1] + rot[2][2]*t[2];
}

int RC_BmodelSetup (const rc_hostworld_t *hw, int submodel, const entity_t *e, const rc_refdef_t *rd, const rc_view_t *v,
		    float pixelAspect, rc_bent_t *out)
{
	const rc_world_t *w = &hw->w;
	const rc_submodel_t *sm = &hw->submodels[submodel];
	float minmaxs[6], angle, s, c, temp1[3][3], temp2[3][3], temp3[3][3], screenedge[4][3];
	int i, j, rotated, clipflags = 0, node;
	const uint32_t *vis;
	double d;

	rotated = (e->angles[0] || e->angles[1] || e->angles[2]);
	for (j = 0; j < 3; j++)
	{
		if (rotated) { minmaxs[j] = e->origin[j] - sm->radius; minmaxs[3+j] = e->origin[j] + sm->radius; }
		else { minmaxs[j] = e->origin[j] + sm->mins[j]; minmaxs[3+j] = e->origin[j] + sm->maxs[j]; }
	}
	/* R_BmodelCheckBBox */
	if (rotated)
	{
		float wr = hw->submodels[0].radius;		/* r_worldmodel->radius */
		for (i = 0; i < 4; i++)
		{
			d = e->origin[0]*v->clipnormal[i][0] + e->origin[1]*v->clipnormal[i][1] + e->origin[2]*v->clipnormal[i][2];
			d -= v->clipdist[i];
			if (d <= -wr)
				return 0;
			if (d <= wr)
				clipflags |= (1<<i);
		}
	}
	else
	{
		for (i = 0; i < 4; i++)
		{
			const int *pindex = v->frustum_idx[i];
			float rejectpt[3], acceptpt[3];
			rejectpt[0] = minmaxs[pindex[0]]; rejectpt[1] = minmaxs[pindex[1]]; rejectpt[2] = minmaxs[pindex[2]];
			d = rejectpt[0]*v->clipnormal[i][0] + rejectpt[1]*v->clipnormal[i][1] + rejectpt[2]*v->clipnormal[i][2];
			d -= v->clipdist[i];
			if (d <= 0)
				return 0;
			acceptpt[0] = minmaxs[pindex[3+0]]; acceptpt[1] = minmaxs[pindex[3+1]]; acceptpt[2] = minmaxs[pindex[3+2]];
			d = acceptpt[0]*v->clipnormal[i][0] + acceptpt[1]*v->clipnormal[i][1] + acceptpt[2]*v->clipnormal[i][2];
			d -= v->clipdist[i];
			if (d <= 0)
				clipflags |= (1<<i);
		}
	}
	/* R_FindTopNode against the marks R_MarkLeaves made for this view's leaf */
	vis = w->nodevis + (size_t)RC_PointInLeaf (hw, rd->vieworg) * w->visrowwords;
	node = hw->submodels[0].headnode;
	for (;;)
	{
		const rc_node_t *nd;
		const rc_plane_t *p;
		int sides, idx = node >= 0 ? node : w->numnodes + (-1 - node);
		if (!((vis[idx >> 5] >> (idx & 31)) & 1))
			return 0;		/* not visible at all */
		if (node < 0)
		{
			if (w->leafs[-1 - node].contents != RC_CONTENTS_SOLID)
				break;		/* a non-solid leaf: visible and not BSP clipped */
			return 0;
		}
		nd = &w->nodes[node];
		p = &w->planes[nd->plane];
		if (p->type < 3)
			sides = (p->dist <= minmaxs[p->type]) ? 1 : ((p->dist >= minmaxs[3 + p->type]) ? 2 : 3);
		else
		{
			float dist1 = 0, dist2 = 0;
			for (i = 0; i < 3; i++)
			{
				if (p->normal[i] < 0) { dist1 += p->normal[i]*minmaxs[i]; dist2 += p->normal[i]*minmaxs[3+i]; }
				else { dist1 += p->normal[i]*minmaxs[3+i]; dist2 += p->normal[i]*minmaxs[i]; }
			}
			sides = 0;
			if (dist1 >= p->dist) sides = 1;
			if (dist2 < p->dist) sides |= 2;
		}
		if (sides == 3)
			break;			/* this is the splitter */
		node = (sides & 1) ? nd->children[0] : nd->children[1];
	}
	memset (out, 0, sizeof(*out));
	out->topnode = node;
	out->clipflags = clipflags;
	out->firstface = sm->firstface; out->numfaces = sm->numfaces; out->headnode = sm->headnode;
	out->frame = e->frame;
	out->dlight_ofs = -1;
	for (i = 0; i < 3; i++)
	{
		out->origin[i] = e->origin[i];
		out->x.origin[i] = v->origin[i] - e->origin[i];		/* modelorg */
		out->x.vpn[i] = v->vpn[i]; out->x.vright[i] = v->vright[i]; out->x.vup[i] = v->vup[i];
	}
	/* transformed_modelorg as D_DrawSurfaces derives it (d_edge.c:254-256,296-298): local origin in the
	 * UNROTATED view basis */
	out->x.tmodelorg[0] = out->x.origin[0]*v->vright[0] + out->x.origin[1]*v->vright[1] + out->x.origin[2]*v->vright[2];
	out->x.tmodelorg[1] = out->x.origin[0]*v->vup[0] + out->x.origin[1]*v->vup[1] + out->x.origin[2]*v->vup[2];
	out->x.tmodelorg[2] = out->x.origin[0]*v->vpn[0] + out->x.origin[1]*v->vpn[1] + out->x.origin[2]*v->vpn[2];

	/* R_RotateBmodel, r_bsp.c:78-151 */
	angle = e->angles[1];
	angle = angle * M_PI*2 / 360;
	s = sin (angle); c = cos (angle);
	temp1[0][0] = c; temp1[0][1] = s; temp1[0][2] = 0;
	temp1[1][0] = -s; temp1[1][1] = c; temp1[1][2] = 0;
	temp1[2][0] = 0; temp1[2][1] = 0; temp1[2][2] = 1;
	angle = e->angles[0];
	angle = angle * M_PI*2 / 360;
	s = sin (angle); c = cos (angle);
	temp2[0][0] = c; temp2[0][1] = 0; temp2[0][2] = -s;
	temp2[1][0] = 0; temp2[1][1] = 1; temp2[1][2] = 0;
	temp2[2][0] = s; temp2[2][1] = 0; temp2[2][2] = c;
	concat_rotations (temp2, temp1, temp3);
	angle = e->angles[2];
	angle = angle * M_PI*2 / 360;
	s = sin (angle); c = cos (angle);
	temp1[0][0] = 1; temp1[0][1] = 0; temp1[0][2] = 0;
	temp1[1][0] = 0; temp1[1][1] = c; temp1[1][2] = s;
	temp1[2][0] = 0; temp1[2][1] = -s; temp1[2][2] = c;
	concat_rotations (temp1, temp3, out->rot);
	entity_rotate (out->rot, out->x.origin);
	entity_rotate (out->rot, out->x.vpn);
	entity_rotate (out->rot, out->x.vright);
	entity_rotate (out->rot, out->x.vup);
	RC_ScreenEdges (v, pixelAspect, rd->fov_x, screenedge);
	RC_TransformFrustum (screenedge, out->x.vright, out->x.vup, out->x.vpn, out->x.origin, out->x.clipnormal, out->x.clipdist);
	return 1;
}
