/*
 * rt_host.h -- internal declarations shared by the host C files and the CUDA
 * translation units (not part of the public ABI; see include/rt_cuda.h).
 */
#ifndef RT_HOST_H
#define RT_HOST_H

#include "rt_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Per-frame constants of ray_through_screen_at() (camera.c:99-118). */
typedef struct {
	RtVector3 origin;   /* camera_pos */
	RtVector3 llc;      /* lower_left_corner */
	RtVector3 horiz;    /* u * screen_w */
	RtVector3 vert;     /* v * screen_h */
} RtCameraFrame;

void rt_host_camera_frame(const RtCamera *cam, float aspect, RtCameraFrame *out);

/* Host mirror of float4. */
typedef struct { float x, y, z, w; } RtF4;

#define RT_MAT_STRIDE 4   /* float4 records per material */

/*
 * Device layout of a scene (structure of arrays, all arrays `n` long):
 *   geomA[i]  sphere: center.xyz, radius*radius     cube: origin.xyz, 0
 *   geomB[i]  sphere: 0,0,0, type bits              cube: (origin*1 + size*1).xyz, type bits
 *             (.w holds the object type as an int bit pattern)
 *   mat[4i+0] f0.xyz, roughness          f0 = f0_d*(1-metallic) + albedo*metallic  (main.c:219-221)
 *   mat[4i+1] (1 - f0).xyz, metal flag   flag = 1.0f iff (double)metallic > 0.001   (main.c:241)
 *   mat[4i+2] (emission_color*emission_power).xyz, emission_power                   (main.c:203,232)
 *   mat[4i+3] (albedo*(1-metallic)).xyz, 0                                          (main.c:248)
 * Everything precomputed here is evaluated with the reference's own float
 * expressions (this file's .c is built with -ffp-contract=off), so the device
 * sees bit-identical operands.
 */
typedef struct {
	int       n;
	int       light_index;   /* first object with emission_power > 0 (main.c:140-146), or -1 */
	RtVector3 light_pos;     /* origin_of(light) (scene.c:10-15) */
	RtF4     *geomA;
	RtF4     *geomB;
	RtF4     *mat;
	/* bounds of all primitives (used for LBVH padding) */
	RtVector3 bounds_lo, bounds_hi;
	int       num_spheres, num_cubes;
	int       div_safe;      /* every cube coordinate is zero or has magnitude in [2^-37, 2^59] */
	/* the ONE object whose emission_color*emission_power is not (+-0, +-0, +-0), or -1 when there
	 * are none or several.  A light sample only asks what its nearest hit EMITS (main.c:199-204:
	 * sampled += emission_color * emission_power of hit2.object), and adding a zero vector changes
	 * nothing, so with a single emitter the question is "is the emitter the nearest hit?" */
	int       only_emitter;
} RtPackedScene;

int  rt_host_pack_scene(const RtObject *objects, int n, RtPackedScene *out);
void rt_host_free_packed(RtPackedScene *p);

/* (float)i / 255 for i in 0..255, the texel -> float rule of sample_cubemap
 * (gpu_and_windowing.c:106-110). */
void rt_host_byte_lut(float lut[256]);

uint64_t rt_host_splitmix64(uint64_t z);

#ifdef __cplusplus
}
#endif
#endif
