/*
 * tests/linkhost/engine_frames.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The ENGINE side of the renderer boundary, written against the reference's own headers (client.h, the header the engine's own client files include, pulls in
 * client/ref.h, client/render.h and client/vid.h from /root/reference by -I; nothing is copied) and calling only what
 * the engine calls: the map-change sequence of client/cl_parse.c:271,294 (Mod_ForName, R_NewMap), the frame sequence of
 * client/cl_screen.c:1768-1807 (R_BeginFrame, V_RenderView -> R_RenderView, the 2D Draw_* calls, R_EndFrame) and the
 * yaw sweep of R_TimeRefresh_f (ref_soft/r_misc.c:32-64).
 *
 * The SAME object file is linked twice: against oracle/_ref/librefsoft.so (the unmodified reference renderer) and
 * against tochris_b200/libref_cuda.so (the product) -- the reference selects its renderer at link time
 * (tochris.dsp:223-236), so this is the drop-in.  tests/test_linkhost.py compares the bytes both leave in vid.buffer
 * and d_pzbuffer.  Only the platform layer differs (platform_refsoft.c / platform_refcuda.c = the job of VID_Init).
 */
#include "client.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern short *d_pzbuffer;			/* ref_soft/d_iface.h: the z-buffer the engine's vid driver owns */
extern int platform_init (const char *basedir, int width, int height);
extern void platform_shutdown (void);
extern int platform_use_refapi (void);		/* product only: go through GetRefAPI (1) instead of the bare symbols */
extern void platform_frame_calls (void (**begin) (void), void (**render) (refdef_t *), void (**end) (void),
				   void (**newmap) (struct model_s *), struct model_s *(**regmodel) (char *));

static struct model_s *direct_register (char *name) { return Mod_ForName (name, true); }

int main (int argc, char **argv)
{
	static lightstyle_t styles[MAX_LIGHTSTYLES];
	static entity_t ents[3];
	static particle_t parts[64];
	static dlight_t dl[1];
	refdef_t rd;
	struct model_s *world, *ball, *blob;
	struct mpic_s *pic;
	void (*begin) (void) = R_BeginFrame;
	void (*render) (refdef_t *) = R_RenderView;
	void (*end) (void) = R_EndFrame;
	void (*newmap) (struct model_s *) = R_NewMap;
	struct model_s *(*regmodel) (char *) = direct_register;
	FILE *out;
	int i, j, width, height, frames;
	float org[3];

	if (argc < 10)
	{
		fprintf (stderr, "usage: %s basedir map width height frames x y z outfile\n", argv[0]);
		return 2;
	}
	width = atoi (argv[3]); height = atoi (argv[4]); frames = atoi (argv[5]);
	org[0] = (float)atof (argv[6]); org[1] = (float)atof (argv[7]); org[2] = (float)atof (argv[8]);
	if (platform_init (argv[1], width, height))
		return 1;
	if (platform_use_refapi ())
		platform_frame_calls (&begin, &render, &end, &newmap, &regmodel);

	/* map change: cl_parse.c:271,294 */
	world = regmodel (argv[2]);
	if (!world) { fprintf (stderr, "no world\n"); return 1; }
	newmap (world);
	ball = regmodel ("progs/ball.mdl");
	blob = regmodel ("progs/blob.mdl");
	pic = Draw_PicFromWad ("sbar");
	if (!pic) pic = Draw_CachePic ("gfx/menuplyr.lmp");

	memset (styles, 0, sizeof(styles));
	for (i = 0; i < MAX_LIGHTSTYLES; i++) { styles[i].length = 1; styles[i].map[0] = 'm'; }
	strcpy (styles[1].map, "mmnmmommommnonmmonqnmmo"); styles[1].length = (int)strlen (styles[1].map);

	out = fopen (argv[9], "wb");
	if (!out) { perror (argv[9]); return 1; }
	for (i = 0; i < frames; i++)
	{
		memset (&rd, 0, sizeof(rd));
		/* the status-bar layout of the engine: the view leaves the bottom rows to the 2D code on odd frames */
		rd.x = 0; rd.y = 0; rd.width = width; rd.height = (i & 1) ? height - 24 : height;
		rd.time = 0.5f + 0.1f * i; rd.oldtime = rd.time;
		rd.fov_x = 90;
		{
			/* CalcFov, client/cl_screen.c */
			float x = rd.width / tan (rd.fov_x / 360 * M_PI);
			rd.fov_y = atan (rd.height / x) * 360 / M_PI;
		}
		VectorCopy (org, rd.vieworg);
		rd.viewangles[0] = (i % 5) * 3.0f - 6.0f;
		rd.viewangles[1] = i / (float)frames * 360.0f;		/* r_misc.c:46 */
		rd.viewangles[2] = 0;
		rd.lightstyles = styles;
		memset (ents, 0, sizeof(ents));
		for (j = 0; j < 3; j++)
		{
			float a = (rd.viewangles[1] + (j - 1) * 25.0f) * (M_PI / 180.0);
			ents[j].model = (j == 1) ? blob : ball;
			ents[j].origin[0] = org[0] + cos (a) * (70.0f + 30.0f * j);
			ents[j].origin[1] = org[1] + sin (a) * (70.0f + 30.0f * j);
			ents[j].origin[2] = org[2] - 12.0f + 9.0f * j;
			ents[j].angles[1] = 30.0f * i + 50.0f * j;
			ents[j].frame = (i + j) % 2; ents[j].oldframe = (i + j + 1) % 2; ents[j].framelerp = 0.25f * (i % 4);
			ents[j].number = j + 1;
			ents[j].colormap = vid.colormap;
		}
		rd.num_entities = 3; rd.entities = ents;
		for (j = 0; j < 64; j++)
		{
			float a = (rd.viewangles[1] + (j - 32) * 1.1f) * (M_PI / 180.0);
			parts[j].org[0] = org[0] + cos (a) * (40.0f + 2.0f * j);
			parts[j].org[1] = org[1] + sin (a) * (40.0f + 2.0f * j);
			parts[j].org[2] = org[2] + (j % 9) - 4.0f;
			parts[j].color = 100 + j;
		}
		rd.num_particles = 64; rd.particles = parts;
		dl[0].origin[0] = org[0] + 40; dl[0].origin[1] = org[1]; dl[0].origin[2] = org[2]; dl[0].radius = 250;
		rd.num_dlights = (i % 3 == 0); rd.dlights = dl;

		begin ();
		render (&rd);
		/* 2D on top, as SCR_UpdateScreen does after V_RenderView (cl_screen.c:1780-1800) */
		if (i & 1)
		{
			Draw_TileClear (0, height - 24, width, 24);
			if (pic) Draw_Pic (0, height - 24, pic);
		}
		Draw_Fill (8, 8, 40, 10, 15 + i);
		Draw_String (12, 9, "fps");
		Draw_Character (width - 16, 8, 48 + (i % 10));
		if (i == frames - 1)
			Draw_FadeScreen ();
		end ();

		for (j = 0; j < height; j++)
			fwrite (vid.buffer + (size_t)j * vid.rowbytes, 1, width, out);
		/* the z-buffer inside the view rectangle (outside it the reference keeps older frames' values) */
		for (j = 0; j < rd.height; j++)
			fwrite (d_pzbuffer + (size_t)j * width, sizeof(short), width, out);
	}
	fclose (out);
	platform_shutdown ();
	return 0;
}
