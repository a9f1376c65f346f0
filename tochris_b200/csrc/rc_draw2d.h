/* 2D overlay (ref_soft/r_draw.c): one command = one Draw_* call; the body below is the work of one pixel of a
 * command, shared by the CUDA kernel (k_draw2d, one CTA per view, commands in order) and the host simulator
 * (tests only).  Commands are validated on the host (rc_host_draw.c) the way the reference Sys_Errors. */
#ifndef RC_DRAW2D_H
#define RC_DRAW2D_H
#include "rc_types.h"
#include "../../include/ref_cuda.h"

#ifndef RC_HD
#ifdef __CUDACC__
#define RC_HD __host__ __device__ __forceinline__
#else
#define RC_HD static inline
#endif
#endif

typedef struct {
	int ofs;			/* first pixel in the blob */
	int width, height;
	int alpha;			/* the picture holds TRANSPARENT_COLOR (mpic_t.alpha, r_draw.c:62) */
} rc_pic_t;

typedef struct {
	const uint8_t *blob;		/* conchars (128x128) at 0, then pictures and 256-byte translation tables */
	const rc_pic_t *pics;		/* [0] unused, [1] backtile, [2] conback (320x200, version string drawn in), then registered pictures */
	int width, height;		/* vid.width / vid.height (rowbytes = width) */
} rc_d2ctx_t;

#define RC_D2_TRANSPARENT 255		/* TRANSPARENT_COLOR, d_iface.h */

/* pixels a command works on */
RC_HD int rc_d2_count (const rc_d2ctx_t *x, const rc_draw2d_t *c)
{
	switch (c->op)
	{
	case RC_D2_CHAR: return 64;
	case RC_D2_PIC: case RC_D2_TRANSPIC: case RC_D2_TRANSPIC_TRANSLATE:
		return x->pics[c->pic].width * x->pics[c->pic].height;
	case RC_D2_CONBACK: return c->arg * x->width;
	case RC_D2_TILECLEAR: case RC_D2_FILL: return c->w * c->h;
	case RC_D2_FADESCREEN: return x->width * x->height;
	}
	return 0;
}

/* work item p of command c on the view's 8-bit plane */
RC_HD void rc_d2_pixel (const rc_d2ctx_t *x, const rc_draw2d_t *c, int p, uint8_t *fb)
{
	const int W = x->width;
	switch (c->op)
	{
	case RC_D2_CHAR:
	{
		/* Draw_Character, r_draw.c:152-205: 8x8 cell of the 128x128 sheet, 0 transparent, clipped at the top only */
		const int num = c->arg & 255, i = p & 7, j = p >> 3;
		if (c->y <= -8 || c->y > x->height - 8 || c->x < 0 || c->x > W - 8)
			return;
		if (c->y + j < 0)
			return;
		const uint8_t s = x->blob[((num >> 4) << 10) + ((num & 15) << 3) + j * 128 + i];
		if (s)
			fb[(c->y + j) * W + c->x + i] = s;
		return;
	}
	case RC_D2_PIC: case RC_D2_TRANSPIC: case RC_D2_TRANSPIC_TRANSLATE:
	{
		/* Draw_Pic / Draw_TransPic / Draw_TransPicTranslate, r_draw.c:226-375 */
		const rc_pic_t *pc = &x->pics[c->pic];
		const int u = p % pc->width, v = p / pc->width;
		uint8_t t = x->blob[pc->ofs + p];
		if ((c->op != RC_D2_PIC || pc->alpha) && t == RC_D2_TRANSPARENT)
			return;
		if (c->op == RC_D2_TRANSPIC_TRANSLATE)
			t = x->blob[c->table + t];
		fb[(c->y + v) * W + c->x + u] = t;
		return;
	}
	case RC_D2_CONBACK:
	{
		/* Draw_ConsoleBackground, r_draw.c:440-464: the lower `lines` rows of the 320x200 picture, stretched */
		const int xx = p % W, y = p / W;
		const rc_pic_t *pc = &x->pics[2];
		const int v = (x->height - c->arg + y) * 200 / x->height;
		const int fstep = 320 * 0x10000 / W;
		fb[y * W + xx] = x->blob[pc->ofs + v * 320 + ((xx * fstep) >> 16)];
		return;
	}
	case RC_D2_TILECLEAR:
	{
		/* Draw_TileClear, r_draw.c:520-575: the back tile repeated from the screen origin */
		const rc_pic_t *pc = &x->pics[1];
		const int xx = c->x + p % c->w, y = c->y + p / c->w;
		fb[y * W + xx] = x->blob[pc->ofs + (y % pc->height) * pc->width + (xx % pc->width)];
		return;
	}
	case RC_D2_FILL:
		fb[(c->y + p / c->w) * W + c->x + p % c->w] = (uint8_t)c->arg;		/* Draw_Fill, r_draw.c:585-595 */
		return;
	case RC_D2_FADESCREEN:
	{
		/* Draw_FadeScreen, r_draw.c:604-630 */
		const int xx = p % W, y = p / W;
		if ((xx & 3) != ((y & 1) << 1))
			fb[y * W + xx] = 0;
		return;
	}
	}
}
#endif
