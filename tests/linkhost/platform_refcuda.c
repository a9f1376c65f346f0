/* TEST INFRASTRUCTURE ONLY: platform layer of the link test for libref_cuda.so -- exactly the VID_Init of
 * the engine owns `vid` and d_pzbuffer and registers them once.  include/ref_cuda.h is
 * included AFTER the reference's headers with RC_NO_REF_TYPES, so every render.h prototype it repeats is checked
 * against the reference's declaration by the compiler. */
#include "client.h"
#define RC_NO_REF_TYPES
#include "ref_cuda.h"

viddef_t vid;
short *d_pzbuffer;
static refexport_t re;
static int use_refapi;

static struct model_s *reg (char *name) { return re.RegisterModel (name); }

int platform_init (const char *basedir, int width, int height)
{
	rc_config_t cfg;
	RC_DefaultConfig (&cfg);
	cfg.width = width; cfg.height = height; cfg.basedir = basedir; cfg.device = 0;
	if (RC_Init (&cfg) != 0)
	{
		fprintf (stderr, "ref_cuda: %s\n", RC_LastError ());
		return 1;
	}
	memset (&vid, 0, sizeof(vid));
	vid.width = width; vid.height = height; vid.rowbytes = width + 32;	/* a window pitch: rowbytes > width (vid.h:38) */
	vid.aspect = ((float)height / (float)width) * (320.0 / 240.0);
	vid.numpages = 1;
	{
		/* host.c:830 loads gfx/colormap.lmp, VID_Init copies it into vid.colormap (win32/vid_win.c:2150-2153) */
		char path[600];
		FILE *f;
		snprintf (path, sizeof(path), "%s/id1/gfx/colormap.lmp", basedir);
		f = fopen (path, "rb");
		if (!f || fread (vid.colormap, 1, sizeof(vid.colormap), f) != sizeof(vid.colormap)) { perror (path); return 1; }
		fclose (f);
	}
	vid.buffer = calloc (1, (size_t)vid.rowbytes * height);
	d_pzbuffer = calloc (sizeof(short), (size_t)width * height);
	RC_SetVidBuffers (vid.buffer, vid.rowbytes, d_pzbuffer);
	Draw_Init ();
	use_refapi = getenv ("LINKHOST_REFAPI") != NULL;
	if (use_refapi)
	{
		re = GetRefAPI (1);
		re.Init ();
	}
	else
		R_Init ();
	return 0;
}
void platform_shutdown (void) { RC_Shutdown (); }
int platform_use_refapi (void) { return use_refapi; }
void platform_frame_calls (void (**begin) (void), void (**render) (refdef_t *), void (**end) (void),
			   void (**newmap) (struct model_s *), struct model_s *(**regmodel) (char *))
{
	*begin = re.BeginFrame; *render = re.RenderFrame; *end = re.EndFrame; *newmap = re.BeginRegistration; *regmodel = reg;
}
