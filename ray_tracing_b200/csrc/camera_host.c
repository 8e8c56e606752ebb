/*
 * camera_host.c -- process-global camera pose with the reference's mutators
 * (reference: src/camera.c:23-93) plus the per-frame constants of
 * ray_through_screen_at() (camera.c:99-118) that the render kernels take by
 * value, and the host-side quantisation rule of screenshot() (main.c:666-670).
 *
 * Built with gcc -std=c11 -O2 -ffp-contract=off: the same binary32/binary64
 * expression shapes as the reference, evaluated by the same libm on the same
 * box (tan, sin, cos are glibc's; SURVEY.md section 8(c) "Third-party arithmetic").
 */
#include "rt_cuda.h"
#include "rt_host.h"

#include <math.h>

/* ---- tiny float3 helpers with the operand order of vector.c ---- */
static RtVector3 v3(float x, float y, float z) { RtVector3 r = {x, y, z}; return r; }
static RtVector3 mix2(RtVector3 u, float a, RtVector3 v, float b)            /* vector.c:145-152 */
{
	return v3(u.x * a + v.x * b, u.y * a + v.y * b, u.z * a + v.z * b);
}
static RtVector3 times(RtVector3 v, float f) { return v3(v.x * f, v.y * f, v.z * f); }
static RtVector3 crossp(RtVector3 u, RtVector3 v)                             /* vector.c:163-170 */
{
	return v3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
static RtVector3 unit(RtVector3 v)                                            /* vector.c:113-135 */
{
	float n = (float) sqrt((double) (v.x * v.x + v.y * v.y + v.z * v.z));
	if ((double) n < 0.00001 && (double) n > -0.00001) return v;
	return v3(v.x / n, v.y / n, v.z / n);
}

/* ---- pose state, camera.c:23-35 ---- */
static int   latch_cursor = 1;
static float yaw_deg, pitch_deg, cursor_x, cursor_y;
static RtCamera pose;
static int   pose_ready = 0;

void rt_camera_reset(void)
{
	latch_cursor = 1;
	yaw_deg = -90.0f;
	pitch_deg = 0.0f;
	cursor_x = 800.0f / 2.0;
	cursor_y = 600.0f / 2.0;
	pose.pos = v3(5, 5, 5);
	pose.front = v3(-1, -1, -1);
	pose.up = v3(0, 1, 0);
	pose.fov = 30.0f;
	pose_ready = 1;
}

static void ensure_pose(void) { if (!pose_ready) rt_camera_reset(); }

RtVector3 rt_get_camera_pos(void) { ensure_pose(); return pose.pos; }
RtCamera  rt_camera_snapshot(void) { ensure_pose(); return pose; }

static float radians(float deg) { return 3.14159265358979323846 * deg / 180; }   /* vector.c:94-97, in double */

/* camera.c:42-78: the first event only latches the cursor, yet still recomputes
 * `front` from yaw/pitch (so it snaps to yaw -90, pitch 0). */
void rt_rotate_camera(double mouse_x, double mouse_y)
{
	ensure_pose();
	float x = mouse_x, y = mouse_y;
	if (latch_cursor) {
		cursor_x = x;
		cursor_y = y;
		latch_cursor = 0;
	}
	float dx = x - cursor_x;
	float dy = cursor_y - y;
	cursor_x = x;
	cursor_y = y;

	const float sensitivity = 0.1f;
	dx *= sensitivity;
	dy *= sensitivity;
	yaw_deg += dx;
	pitch_deg += dy;
	if (pitch_deg > 89.0f) pitch_deg = 89.0f;
	if (pitch_deg < -89.0f) pitch_deg = -89.0f;

	float yr = radians(yaw_deg), pr = radians(pitch_deg);
	RtVector3 f;
	f.x = cos(yr) * cos(pr);        /* double products narrowed on assignment */
	f.y = sin(pr);
	f.z = sin(yr) * cos(pr);
	pose.front = unit(f);
}

/* camera.c:80-88: UP/DOWN travel along the (unnormalised) front vector */
void rt_move_camera(RtDirection dir, float speed)
{
	ensure_pose();
	switch (dir) {
	case RT_UP:    pose.pos = mix2(pose.pos, 1, pose.front, +speed); break;
	case RT_DOWN:  pose.pos = mix2(pose.pos, 1, pose.front, -speed); break;
	case RT_LEFT:  pose.pos = mix2(pose.pos, 1, unit(crossp(pose.front, pose.up)), -speed); break;
	case RT_RIGHT: pose.pos = mix2(pose.pos, 1, unit(crossp(pose.front, pose.up)), +speed); break;
	}
}

/* camera.c:99-118: everything in ray_through_screen_at() that does not depend
 * on (px,py).  The kernels finish with the last combine4 (camera.c:121). */
void rt_host_camera_frame(const RtCamera *cam, float aspect, RtCameraFrame *out)
{
	RtVector3 w = unit(times(cam->front, -1));
	RtVector3 u = unit(crossp(cam->up, w));
	RtVector3 v = crossp(w, u);
	float screen_h = 2 * tan(cam->fov / 2);     /* fov/2 in float, tan in double, narrowed */
	float screen_w = aspect * screen_h;
	RtVector3 horiz = times(u, screen_w);
	RtVector3 vert = times(v, screen_h);
	/* combine4(pos, horiz, vert, w, 1, -0.5, -0.5, -1): ((p*1 + h*-.5) + v*-.5) + w*-1 */
	RtVector3 llc;
	llc.x = cam->pos.x * 1 + horiz.x * -0.5f + vert.x * -0.5f + w.x * -1;
	llc.y = cam->pos.y * 1 + horiz.y * -0.5f + vert.y * -0.5f + w.y * -1;
	llc.z = cam->pos.z * 1 + horiz.z * -0.5f + vert.z * -0.5f + w.z * -1;
	out->origin = cam->pos;
	out->llc = llc;
	out->horiz = horiz;
	out->vert = vert;
}

/* main.c:666-670: implicit float -> uint8_t conversion of x*255 (truncation) */
void rt_quantize_frame(const float *rgb, size_t num_pixels, uint8_t *out)
{
	for (size_t i = 0; i < 3 * num_pixels; i++)
		out[i] = (uint8_t) (rgb[i] * 255);
}
