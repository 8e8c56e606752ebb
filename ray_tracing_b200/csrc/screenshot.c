/*
 * screenshot.c -- the pixel path of the reference's screenshot()
 * (src/main.c:637-681): frame[i].{x,y,z} * 255 narrowed to uint8_t
 * (main.c:666-670), rows flipped so the file is top row first
 * (stbi_flip_vertically_on_write(1), main.c:672), written as an 8-bit RGB PNG.
 *
 * The reference encodes with stb_image_write (deflate-compressed); this writer
 * emits a valid PNG with stored (uncompressed) deflate blocks -- same pixels,
 * no third-party code.  ".ppm" paths get a binary PPM instead.
 */
#include "rt_cuda.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint32_t crc_table[256];
static int crc_ready = 0;

static void crc_init(void)
{
	for (uint32_t n = 0; n < 256; n++) {
		uint32_t c = n;
		for (int k = 0; k < 8; k++) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
		crc_table[n] = c;
	}
	crc_ready = 1;
}

static uint32_t crc_update(uint32_t c, const uint8_t *p, size_t n)
{
	for (size_t i = 0; i < n; i++) c = crc_table[(c ^ p[i]) & 0xff] ^ (c >> 8);
	return c;
}

static void put32(uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

static int chunk(FILE *f, const char *type, const uint8_t *data, size_t len)
{
	uint8_t hdr[8];
	put32(hdr, (uint32_t) len);
	memcpy(hdr + 4, type, 4);
	uint32_t c = crc_update(0xffffffffu, hdr + 4, 4);
	if (len) c = crc_update(c, data, len);
	uint8_t tail[4];
	put32(tail, c ^ 0xffffffffu);
	return fwrite(hdr, 1, 8, f) == 8 && (len == 0 || fwrite(data, 1, len, f) == len) && fwrite(tail, 1, 4, f) == 4;
}

/* rgb: h rows of w*3 bytes, top row first */
static int write_png(FILE *f, const uint8_t *rgb, int w, int h)
{
	if (!crc_ready) crc_init();
	static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
	if (fwrite(sig, 1, 8, f) != 8) return 0;
	uint8_t ihdr[13];
	put32(ihdr, (uint32_t) w);
	put32(ihdr + 4, (uint32_t) h);
	ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;     /* 8-bit RGB, no interlace */
	if (!chunk(f, "IHDR", ihdr, 13)) return 0;

	/* zlib stream: header, stored blocks of <= 65535 bytes over (filter byte + row) * h, adler32 */
	size_t row = (size_t) w * 3 + 1, raw = row * (size_t) h;
	size_t nblocks = (raw + 65534) / 65535;
	size_t zlen = 2 + raw + 5 * (nblocks ? nblocks : 1) + 4;
	uint8_t *z = (uint8_t *) malloc(zlen);
	if (!z) return 0;
	size_t o = 0;
	z[o++] = 0x78; z[o++] = 0x01;
	uint32_t a = 1, b = 0;
	size_t done = 0;
	uint8_t *line = (uint8_t *) malloc(row);
	if (!line) { free(z); return 0; }
	size_t in_block = 0, block_left = 0;
	for (int y = 0; y < h; y++) {
		line[0] = 0;                                    /* filter: none */
		memcpy(line + 1, rgb + (size_t) y * w * 3, (size_t) w * 3);
		for (size_t i = 0; i < row; i++) {
			if (block_left == 0) {
				size_t remaining = raw - done;
				block_left = remaining < 65535 ? remaining : 65535;
				z[o++] = remaining <= 65535 ? 1 : 0;    /* BFINAL */
				z[o++] = block_left & 0xff; z[o++] = block_left >> 8;
				z[o++] = ~block_left & 0xff; z[o++] = (~block_left >> 8) & 0xff;
				in_block = 0;
			}
			z[o++] = line[i];
			a = (a + line[i]) % 65521u;
			b = (b + a) % 65521u;
			done++; block_left--; in_block++;
		}
	}
	(void) in_block;
	if (raw == 0) { z[o++] = 1; z[o++] = 0; z[o++] = 0; z[o++] = 0xff; z[o++] = 0xff; }
	put32(z + o, (b << 16) | a);
	o += 4;
	int ok = chunk(f, "IDAT", z, o) && chunk(f, "IEND", NULL, 0);
	free(line);
	free(z);
	return ok;
}

int rt_save_screenshot(const char *path, const float *frame_rgb, int w, int h)
{
	if (!path || !frame_rgb || w <= 0 || h <= 0) return RT_ERR_ARG;
	size_t n = (size_t) w * h;
	uint8_t *q = (uint8_t *) malloc(n * 3), *flipped = (uint8_t *) malloc(n * 3);
	if (!q || !flipped) { free(q); free(flipped); return RT_ERR_NOMEM; }
	rt_quantize_frame(frame_rgb, n, q);                 /* main.c:666-670 */
	for (int y = 0; y < h; y++)                         /* main.c:672: flip vertically on write */
		memcpy(flipped + (size_t) y * w * 3, q + (size_t) (h - 1 - y) * w * 3, (size_t) w * 3);
	FILE *f = fopen(path, "wb");
	int ok = 0;
	if (f) {
		size_t len = strlen(path);
		if (len > 4 && strcmp(path + len - 4, ".ppm") == 0) {
			fprintf(f, "P6\n%d %d\n255\n", w, h);
			ok = fwrite(flipped, 1, n * 3, f) == n * 3;
		} else
			ok = write_png(f, flipped, w, h);
		ok = (fclose(f) == 0) && ok;
	}
	free(q);
	free(flipped);
	return ok ? RT_OK : RT_ERR_IO;
}
