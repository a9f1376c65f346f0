/* TEST INFRASTRUCTURE ONLY: platform layer of the link test for the reference renderer -- oracle/headless.c's orc_init
 * is the VID_Init / Host_Init subset (it owns `vid` and d_pzbuffer). */
#include "client.h"
extern int orc_init (const char *basedir, int width, int height, int heap_mb, int verbose);
extern const char *orc_last_error (void);
int platform_init (const char *basedir, int width, int height)
{
	if (orc_init (basedir, width, height, 64, 0) != 0)
	{
		fprintf (stderr, "orc_init: %s\n", orc_last_error ());
		return 1;
	}
	return 0;
}
void platform_shutdown (void) {}
int platform_use_refapi (void) { return 0; }
void platform_frame_calls (void (**begin) (void), void (**render) (refdef_t *), void (**end) (void),
			   void (**newmap) (struct model_s *), struct model_s *(**regmodel) (char *))
{ (void)begin; (void)render; (void)end; (void)newmap; (void)regmodel; }
