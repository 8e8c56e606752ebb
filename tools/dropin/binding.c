/*
 * binding.c -- what a maintainer of cozis/ray_tracing adds to src/main.c to run
 * the render path on a B200: the four functions below replace
 * the worker pool -- start_workers / stop_workers (main.c:687-718),
 * invalidate_accumulation (main.c:115-124) and update_frame (main.c:450-482).
 * Everything else of the reference (argument parser, scene parser, skybox
 * loader, camera, event loop, presenter) is used as it is.
 *
 * oracle/Makefile `dropin` links this file with the UNMODIFIED reference
 * objects (the four originals demoted to weak symbols, nothing edited) and
 * libraytrace_b200.so; tests/test_dropin.py runs the result.
 */
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>

#include "scene.h"
#include "camera.h"
#include "gpu_and_windowing.h"
#define RT_CUDA_REFERENCE_TYPES      /* reuse the reference's Vector3/Object/Scene/Cubemap */
#include "rt_cuda.h"

#ifndef RT_DROPIN_BUDGET_MS
#define RT_DROPIN_BUDGET_MS 8.0      /* keep adding passes for this long per displayed frame */
#endif

/* globals of src/main.c:50-72 */
extern int num_columns, init_scale;
extern Scene scene;
extern Cubemap skybox;
extern Vector3 *frame;
extern int frame_w, frame_h;

/* main.c:416-449, unchanged */
void realloc_frame_buffer(void);
bool frame_buffer_size_doesnt_match_window(void);

/* the pose getter camera.c lacks (tools/dropin/camera_snapshot.c) */
RtCamera camera_snapshot(void);

static void die(void)
{
	fprintf(stderr, "%s\n", rt_cuda_last_error());
	abort();
}

void start_workers(void)               /* replaces main.c:695-706 */
{
	if (rt_cuda_init(1) != RT_OK ||                     /* or rt_cuda_init(8): row blocks over 8 GPUs */
	    rt_cuda_upload_scene(&scene) != RT_OK ||
	    rt_cuda_upload_skybox(&skybox) != RT_OK ||
	    rt_cuda_set_progressive(init_scale, num_columns) != RT_OK)     /* --init-scale, --threads */
		die();
}

void invalidate_accumulation(void)     /* replaces main.c:115-124 */
{
	if (rt_cuda_invalidate_accumulation() != RT_OK)     /* accum = 0, generation++, scale back to init_scale */
		die();
}

void update_frame(void)                /* replaces main.c:450-482 */
{
	if (frame_buffer_size_doesnt_match_window()) {
		realloc_frame_buffer();
		invalidate_accumulation();
	}
	RtCamera cam = camera_snapshot();
	/* one pass at the workers' current scale (16, 8, 4, 2, 1, 1, ... main.c:402-403), plus as many
	 * more as fit in the budget; accum += pass/scale^2, frame = accum/count (main.c:394-396, 476) */
	if (rt_cuda_update_frame(&cam, frame, frame_w, frame_h, RT_DROPIN_BUDGET_MS, NULL, NULL) != RT_OK)
		die();
	move_frame_to_the_gpu(frame_w, frame_h, frame);     /* unchanged: glTexImage2D(GL_RGB, GL_FLOAT) */
}

void stop_workers(void) { rt_cuda_shutdown(); }
