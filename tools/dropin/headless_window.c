/*
 * headless_window.c -- TEST SCAFFOLDING for the drop-in proof (oracle/Makefile
 * `dropin`): stand-ins for the window half of src/gpu_and_windowing.c
 * (:231-397) so that the reference's own main() (src/main.c:484-581) runs on a
 * box without GLFW or a display.  The reference's file is compiled with these
 * seven names renamed away (-D), so load_cubemap / sample_cubemap still come
 * from it.
 *
 *   RT_DROPIN_FRAMES   frames to present before the window "closes"  (default 7)
 *   RT_DROPIN_KEYS     one character per frame: W A S D = key press before that
 *                      frame (main.c:536-558), anything else = no event
 *   RT_DROPIN_SIZE     WxH instead of the 1280x960 main() asks for
 *   RT_DROPIN_DUMP     file receiving the last presented frame (raw f32 RGB,
 *                      bottom row first)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gpu_and_windowing.h"

static int win_w, win_h, frames_shown, frames_wanted = 7, key_sent = -1, close_sent;
static const char *keys = "";
static Vector3 *last_frame;
static int last_w, last_h;

void startup_window_and_opengl_context_or_exit(int window_w, int window_h, const char *title)
{
	(void) title;
	win_w = window_w;
	win_h = window_h;
	const char *s = getenv("RT_DROPIN_SIZE");
	if (s && sscanf(s, "%dx%d", &win_w, &win_h) != 2) { fprintf(stderr, "bad RT_DROPIN_SIZE\n"); exit(-1); }
	if (getenv("RT_DROPIN_FRAMES")) frames_wanted = atoi(getenv("RT_DROPIN_FRAMES"));
	if (getenv("RT_DROPIN_KEYS")) keys = getenv("RT_DROPIN_KEYS");
}

int pop_event(double *mouse_x, double *mouse_y)
{
	*mouse_x = *mouse_y = 0;
	if (frames_shown >= frames_wanted) {
		/* main() drains the queue until EVENT_EMPTY before it looks at its exit flag (main.c:522-528) */
		if (close_sent) return EVENT_EMPTY;
		close_sent = 1;
		return EVENT_CLOSE;
	}
	if (key_sent < frames_shown) {           /* at most one key per frame, before it is rendered */
		key_sent = frames_shown;
		if ((size_t) frames_shown < strlen(keys))
			switch (keys[frames_shown]) {
			case 'W': case 'w': return EVENT_PRESS_W;
			case 'A': case 'a': return EVENT_PRESS_A;
			case 'S': case 's': return EVENT_PRESS_S;
			case 'D': case 'd': return EVENT_PRESS_D;
			default: break;
			}
	}
	return EVENT_EMPTY;
}

int get_screen_w(void) { return win_w; }
int get_screen_h(void) { return win_h; }

void move_frame_to_the_gpu(int w, int h, Vector3 *data)
{
	last_frame = data;
	last_w = w;
	last_h = h;
	/* the frame buffer is freed and reallocated by main.c on resize only, but copy out now:
	 * main() invalidates (zeroes) it once more before it leaves */
	const char *dump = getenv("RT_DROPIN_DUMP");
	if (dump && frames_shown == frames_wanted - 1) {
		FILE *f = fopen(dump, "wb");
		if (!f || fwrite(data, sizeof(Vector3), (size_t) w * h, f) != (size_t) w * h) fprintf(stderr, "could not write %s\n", dump);
		if (f) fclose(f);
	}
}

void draw_frame(void) { frames_shown++; }

void cleanup_window_and_opengl_context(void)
{
	fprintf(stderr, "headless window: %d frames of %dx%d presented\n", frames_shown, last_w, last_h);
}
