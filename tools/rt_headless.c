/*
 * rt_headless.c -- headless stand-in for the reference's main() (src/main.c:484-681)
 * on top of the C ABI (include/rt_cuda.h): same command line, no window.
 *
 *   rt_headless --scene scene_0.txt --threads 8 --init-scale 8 \
 *               [--width 1280 --height 960] [--frames 32] [--gpus 1] [--fast]
 *               [--skybox DIR] [--keys WWAD...] [--dump out.ppm] [--dump-f32 out.raw]
 *
 * --scene / --threads / --init-scale are the reference's flags (main.c:585-634;
 * --threads only selects the column layout to reproduce, it is clamped to 32
 * like MAX_COLUMNS).  Each frame is one update_frame(): one pass at the current
 * scale, accumulated and resolved on the device, scale halving after every
 * pass (main.c:402-403).  --keys replays W/A/S/D presses (main.c:536-558), one
 * character per frame (anything else = no event): each press moves the camera by
 * 0.5 and invalidates the accumulation.  --dump writes what
 * screenshot() would (main.c:637-681): (uint8_t)(x*255), flipped vertically,
 * as a PNG (binary PPM if the name ends in .ppm).
 *
 * The skybox JPEGs are decoded by stb_image when this file is compiled with
 * -DRT_HAVE_STB -I<dir containing stb/stb_image.h> (the reference vendors it
 * under 3p/); otherwise a procedural cubemap is used.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "rt_cuda.h"

#ifdef RT_HAVE_STB
#define STB_IMAGE_IMPLEMENTATION
#define STBI_ONLY_JPEG
#include <stb/stb_image.h>
#endif

static double now_s(void)
{
	struct timespec ts;
	timespec_get(&ts, TIME_UTC);
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static int load_skybox(const char *dir, RtCubemap *c)
{
#ifdef RT_HAVE_STB
	/* face order of main.c:500-507 */
	static const char *names[6] = {"front.jpg", "back.jpg", "left.jpg", "right.jpg", "top.jpg", "bottom.jpg"};
	for (int i = 0; i < 6; i++) {
		char path[4096];
		snprintf(path, sizeof(path), "%s/%s", dir, names[i]);
		c->data[i] = stbi_load(path, &c->w, &c->h, &c->chan, 0);
		if (!c->data[i]) {
			fprintf(stderr, "Couldn't load image '%s'\n", path);
			return -1;
		}
	}
	return 0;
#else
	(void) dir; (void) c;
	return -1;
#endif
}

static void procedural_skybox(RtCubemap *c, int size)
{
	c->w = c->h = size;
	c->chan = 3;
	for (int f = 0; f < 6; f++) {
		uint8_t *p = (uint8_t *) malloc((size_t) size * size * 3);
		for (int y = 0; y < size; y++)
			for (int x = 0; x < size; x++) {
				uint8_t *t = p + 3 * ((size_t) y * size + x);
				t[0] = (uint8_t) (90 + 100 * x / size + 8 * f);
				t[1] = (uint8_t) (120 + 90 * y / size);
				t[2] = (uint8_t) (170 + 60 * (x + y) / (2 * size));
			}
		c->data[f] = p;
	}
}

int main(int argc, char **argv)
{
	const char *scene_file = NULL, *sky_dir = "assets/skybox", *dump = NULL, *dump_f32 = NULL, *keys = "";
	int num_columns = -1, init_scale = 8, w = 2 * 640, h = 2 * 480, frames = 16, gpus = 1, fast = 0;
	for (int i = 1; i < argc; i++) {
		const char *a = argv[i];
		const char *v = i + 1 < argc ? argv[i + 1] : NULL;
#define ARG(name) (!strcmp(a, name) && v && (i++, 1))
		if (ARG("--scene")) scene_file = v;
		else if (ARG("--threads")) num_columns = atoi(v);
		else if (ARG("--init-scale")) init_scale = atoi(v);
		else if (ARG("--width")) w = atoi(v);
		else if (ARG("--height")) h = atoi(v);
		else if (ARG("--frames")) frames = atoi(v);
		else if (ARG("--gpus")) gpus = atoi(v);
		else if (ARG("--skybox")) sky_dir = v;
		else if (ARG("--keys")) keys = v;
		else if (ARG("--dump")) dump = v;
		else if (ARG("--dump-f32")) dump_f32 = v;
		else if (!strcmp(a, "--fast")) fast = 1;
		else fprintf(stderr, "Warning: Ignoring option %s\n", a);
	}
	if (!scene_file) { fprintf(stderr, "Error: No scene specified (you should use --scene <filename>)\n"); return -1; }
	if (num_columns < 0) { fprintf(stderr, "Error: Missing --threads <N> option\n"); return -1; }
	if (num_columns == 0) { fprintf(stderr, "Error: Invalid count for --threads\n"); return -1; }
	if (num_columns > 32) num_columns = 32;                       /* MAX_COLUMNS, main.c:632-633 */
	if (init_scale != 1 && init_scale != 2 && init_scale != 4 && init_scale != 8 && init_scale != 16) {
		fprintf(stderr, "Error: Invalid value for --init-scale. It must be a power of 2 between 1 and 16 (included)\n");
		return -1;
	}

	static RtScene scene;
	if (!rt_parse_scene_file(scene_file, &scene)) { fprintf(stderr, "Couldn't parse scene\n"); return -1; }
	fprintf(stderr, "Scene parsed (%d objects)\n", scene.num_objects);

	RtCubemap sky;
	memset(&sky, 0, sizeof(sky));
	if (load_skybox(sky_dir, &sky) != 0) {
		fprintf(stderr, "Skybox JPEGs not loaded (%s); using a procedural cubemap\n", sky_dir);
		procedural_skybox(&sky, 512);
	}
	fprintf(stderr, "Cubemap loaded (%dx%d)\n", sky.w, sky.h);

	if (rt_cuda_init(gpus) != RT_OK || rt_cuda_upload_scene(&scene) != RT_OK || rt_cuda_upload_skybox(&sky) != RT_OK) {
		fprintf(stderr, "Error: %s\n", rt_cuda_last_error());
		return -1;
	}

	RtVector3 *frame = (RtVector3 *) malloc(sizeof(RtVector3) * (size_t) w * h);
	if (!frame) { printf("OUT OF MEMORY\n"); return -1; }

	rt_camera_reset();
	int scale = init_scale;
	uint64_t pass = 0, rays = 0;
	size_t nkeys = strlen(keys);
	double render_ms = 0, t0 = now_s();
	for (int f = 0; f < frames; f++) {
		if ((size_t) f < nkeys) {                                 /* main.c:526-569 */
			int moved = 1;
			switch (keys[f]) {
			case 'W': case 'w': rt_move_camera(RT_UP, 0.5f); break;
			case 'A': case 'a': rt_move_camera(RT_LEFT, 0.5f); break;
			case 'S': case 's': rt_move_camera(RT_DOWN, 0.5f); break;
			case 'D': case 'd': rt_move_camera(RT_RIGHT, 0.5f); break;
			default: moved = 0; break;                            /* any other character: no event before this frame */
			}
			if (moved) {
				rt_cuda_accum_reset();                            /* invalidate_accumulation() */
				scale = init_scale;
			}
		}
		RtCamera cam = rt_camera_snapshot();
		RtRenderOpts o;
		rt_render_opts_default(&o);
		o.scale = scale;
		o.num_columns = num_columns;
		o.pass_index = pass++;
		o.accumulate = 1;
		o.variant = fast ? RT_VARIANT_FAST : RT_VARIANT_EXACT;
		RtRenderStats st;
		if (render_frame_cuda_ex(&cam, frame, w, h, &o, &st) != RT_OK) {
			fprintf(stderr, "Error: %s\n", rt_cuda_last_error());
			return -1;
		}
		rays += st.rays;
		render_ms += st.render_ms;
		if (scale > 1) scale >>= 1;                               /* main.c:402-403 */
	}
	double wall = now_s() - t0;
	printf("{\"frames\": %d, \"width\": %d, \"height\": %d, \"gpus\": %d, \"rays\": %llu, \"device_ms_per_frame\": %.4f, "
	       "\"wall_ms_per_frame\": %.4f, \"device_mrays_per_s\": %.1f, \"accum_count\": %.4f}\n",
	       frames, w, h, rt_cuda_num_gpus(), (unsigned long long) rays, render_ms / frames, 1e3 * wall / frames,
	       render_ms > 0 ? (double) rays / render_ms / 1e3 : 0.0, rt_cuda_accum_count());

	if (dump_f32) {
		FILE *fp = fopen(dump_f32, "wb");
		if (!fp || fwrite(frame, sizeof(RtVector3), (size_t) w * h, fp) != (size_t) w * h)
			fprintf(stderr, "Could not write %s\n", dump_f32);
		if (fp) fclose(fp);
	}
	if (dump) {
		/* screenshot(): quantise, flip vertically, PNG (or PPM by extension) */
		if (rt_save_screenshot(dump, (const float *) frame, w, h) != RT_OK)
			fprintf(stderr, "Could not take screenshot (write error)\n");
		else
			fprintf(stderr, "Took screenshot! (%s)\n", dump);
	}
	rt_cuda_shutdown();
	free(frame);
	return 0;
}
