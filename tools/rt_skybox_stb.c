/*
 * rt_skybox_stb.c -- the caller's half of load_cubemap()
 * (src/gpu_and_windowing.c:19-40, face files of src/main.c:500-507) for hosts
 * that are not C programs: decodes the six skybox JPEGs with the SAME decoder
 * the reference uses (stb_image v2.29, vendored by the reference under 3p/ and
 * compiled from there, never copied) so that bench.py and the headless tools
 * feed rt_cuda_upload_skybox() byte-identical texels.  libjpeg/PIL decode the
 * same files with different LSBs.  Built as tools/librt_skybox_stb.so when the
 * reference's 3p/ directory is present at build time.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define STB_IMAGE_IMPLEMENTATION
#define STBI_ONLY_JPEG
#include <stb/stb_image.h>

/* CubeFace order (gpu_and_windowing.h:4-11) */
static const char *const face_names[6] = {"front.jpg", "back.jpg", "left.jpg", "right.jpg", "top.jpg", "bottom.jpg"};

/* Decode the six faces found in `dir`.  On success returns 0, stores the face
 * size and channel count (all faces must agree) and one malloc'ed block of
 * 6*w*h*chan bytes (faces back to back, rows top first) in *out; the caller
 * releases it with rt_skybox_free().  Returns -1 and prints the reference's
 * message when a file cannot be decoded. */
int rt_skybox_load_dir(const char *dir, unsigned char **out, int *w, int *h, int *chan)
{
	unsigned char *all = NULL;
	int fw = 0, fh = 0, fc = 0;
	for (int i = 0; i < 6; i++) {
		char path[4096];
		int cw, ch, cc;
		snprintf(path, sizeof(path), "%s/%s", dir, face_names[i]);
		unsigned char *p = stbi_load(path, &cw, &ch, &cc, 0);
		if (!p) {
			fprintf(stderr, "Couldn't load image '%s'\n", path);
			free(all);
			return -1;
		}
		if (i == 0) {
			fw = cw; fh = ch; fc = cc;
			all = (unsigned char *) malloc((size_t) 6 * fw * fh * fc);
		}
		if (!all || cw != fw || ch != fh || cc != fc) {
			fprintf(stderr, "skybox faces differ in size (%s)\n", path);
			stbi_image_free(p);
			free(all);
			return -1;
		}
		memcpy(all + (size_t) i * fw * fh * fc, p, (size_t) fw * fh * fc);
		stbi_image_free(p);
	}
	*out = all; *w = fw; *h = fh; *chan = fc;
	return 0;
}

void rt_skybox_free(unsigned char *p) { free(p); }
