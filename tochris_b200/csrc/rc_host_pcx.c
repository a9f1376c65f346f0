/*
 * rc_host_pcx.c -- 8-bit PCX decoding for the sky-box pictures (ref_soft/pcx.c:33-115 LoadPCX:
 * same header checks, same run-length scheme).  Host side, plain C.
 */
#include <stdlib.h>
#include <string.h>
#include "rc_host.h"

/* returns malloc'd width*height palette indices, or NULL ("Bad pcx file") */
unsigned char *RC_DecodePCX (const void *data, size_t len, int *width, int *height)
{
	const unsigned char *p = (const unsigned char *)data, *end = p + len;
	int xmin, ymin, xmax, ymax, x, y, w, h;
	unsigned char *out, *row;
	if (len < 128)
		return NULL;
	xmin = p[4] | (p[5] << 8); ymin = p[6] | (p[7] << 8);
	xmax = p[8] | (p[9] << 8); ymax = p[10] | (p[11] << 8);
	(void)xmin; (void)ymin;
	/* manufacturer 0x0a, version 5, RLE, 8 bits per pixel, at most 640x480 */
	if (p[0] != 0x0a || p[1] != 5 || p[2] != 1 || p[3] != 8 || xmax >= 640 || ymax >= 480)
		return NULL;
	w = xmax + 1; h = ymax + 1;
	out = (unsigned char *)malloc ((size_t)w * h);
	if (!out)
		return NULL;
	p += 128;
	for (y = 0, row = out; y <= ymax; y++, row += w)
	{
		for (x = 0; x <= xmax; )
		{
			int dataByte, runLength;
			if (p >= end) { free (out); return NULL; }
			dataByte = *p++;
			if ((dataByte & 0xC0) == 0xC0)
			{
				runLength = dataByte & 0x3F;
				if (p >= end) { free (out); return NULL; }
				dataByte = *p++;
			}
			else
				runLength = 1;
			while (runLength-- > 0)
			{
				if (x <= xmax)			/* the reference writes past the row end into the next one; same bytes */
					row[x] = (unsigned char)dataByte;
				else if (row + x < out + (size_t)w * h)
					row[x] = (unsigned char)dataByte;
				x++;
			}
		}
	}
	if (width) *width = w;
	if (height) *height = h;
	return out;
}
