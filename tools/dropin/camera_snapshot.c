/*
 * camera_snapshot.c -- the one addition to src/camera.c that INTEGRATION.md asks
 * for: a getter for the pose the file keeps in statics (camera.c:23-35), next to
 * get_camera_pos() (camera.c:37-40).  The reference file is compiled in place
 * through the #include; nothing in it is edited.
 */
#include "camera.c"

#include "scene.h"
#include "gpu_and_windowing.h"
#define RT_CUDA_REFERENCE_TYPES
#include "rt_cuda.h"

RtCamera camera_snapshot(void)
{
	RtCamera c;
	c.pos = camera_pos;
	c.front = camera_front;
	c.up = camera_up;
	c.fov = fov;
	return c;
}
