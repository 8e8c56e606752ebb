/*
 * scene_pack.c -- AoS `Object` records -> the device's structure-of-arrays
 * (layout documented in rt_host.h).  Per-material and per-primitive constants
 * that the reference recomputes for every ray are evaluated once here with the
 * reference's own binary32/binary64 expressions; built with -ffp-contract=off.
 */
#include "rt_host.h"

#include <float.h>
#include <stdlib.h>
#include <string.h>

static float int_as_float(int32_t i) { float f; memcpy(&f, &i, 4); return f; }

uint64_t rt_host_splitmix64(uint64_t z)
{
	z += 0x9e3779b97f4a7c15ull;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}

uint64_t rt_pixel_key(float px, float py, uint64_t pass_index)
{
	uint32_t bx, by;
	memcpy(&bx, &px, 4);
	memcpy(&by, &py, 4);
	return rt_host_splitmix64((((uint64_t) bx << 32) | by) ^ rt_host_splitmix64(pass_index));
}

void rt_host_byte_lut(float lut[256])
{
	for (int i = 0; i < 256; i++)
		lut[i] = (float) (uint8_t) i / 255;     /* gpu_and_windowing.c:107-109 */
}

/* guard of the hoisted slab-test division (rt_device.cuh: ray_div) */
static int coord_safe(float x)
{
	float a = x < 0 ? -x : x;
	uint32_t bits;
	memcpy(&bits, &x, 4);
	if (x == 0) return bits == 0;          /* +0 only: see div_hoisted */
	return a >= 0x1p-37f && a <= 0x1p59f;
}

static void grow(RtVector3 *lo, RtVector3 *hi, float x, float y, float z)
{
	if (x < lo->x) lo->x = x;
	if (y < lo->y) lo->y = y;
	if (z < lo->z) lo->z = z;
	if (x > hi->x) hi->x = x;
	if (y > hi->y) hi->y = y;
	if (z > hi->z) hi->z = z;
}

int rt_host_pack_scene(const RtObject *objects, int n, RtPackedScene *out)
{
	memset(out, 0, sizeof(*out));
	out->n = n;
	out->light_index = -1;
	out->div_safe = 1;
	out->only_emitter = -1;
	int emitters = 0;
	size_t cnt = n > 0 ? (size_t) n : 1;
	out->geomA = (RtF4 *) calloc(cnt, sizeof(RtF4));
	out->geomB = (RtF4 *) calloc(cnt, sizeof(RtF4));
	out->mat = (RtF4 *) calloc(cnt * RT_MAT_STRIDE, sizeof(RtF4));
	if (!out->geomA || !out->geomB || !out->mat) {
		rt_host_free_packed(out);
		return RT_ERR_NOMEM;
	}
	out->bounds_lo.x = out->bounds_lo.y = out->bounds_lo.z = FLT_MAX;
	out->bounds_hi.x = out->bounds_hi.y = out->bounds_hi.z = -FLT_MAX;

	for (int i = 0; i < n; i++) {
		const RtObject *o = &objects[i];
		const RtMaterial *m = &o->material;
		RtF4 *A = &out->geomA[i], *B = &out->geomB[i];

		if (o->type == RT_OBJECT_SPHERE) {
			float r = o->sphere.radius;
			A->x = o->sphere.center.x; A->y = o->sphere.center.y; A->z = o->sphere.center.z;
			A->w = r * r;                                  /* scene.c:112 */
			B->x = B->y = B->z = 0;
			B->w = int_as_float(RT_OBJECT_SPHERE);
			float ar = r < 0 ? -r : r;
			grow(&out->bounds_lo, &out->bounds_hi, A->x - ar, A->y - ar, A->z - ar);
			grow(&out->bounds_lo, &out->bounds_hi, A->x + ar, A->y + ar, A->z + ar);
			out->num_spheres++;
		} else {
			A->x = o->cube.origin.x; A->y = o->cube.origin.y; A->z = o->cube.origin.z;
			A->w = 0;
			B->x = o->cube.origin.x * 1 + o->cube.size.x * 1;   /* scene.c:27 */
			B->y = o->cube.origin.y * 1 + o->cube.size.y * 1;
			B->z = o->cube.origin.z * 1 + o->cube.size.z * 1;
			/* any type other than sphere falls in the cube arm only if it IS a
			 * cube (scene.c:138-153); unknown types never intersect */
			B->w = int_as_float(o->type == RT_OBJECT_CUBE ? RT_OBJECT_CUBE : 2);
			grow(&out->bounds_lo, &out->bounds_hi, A->x, A->y, A->z);
			grow(&out->bounds_lo, &out->bounds_hi, B->x, B->y, B->z);
			if (!(coord_safe(A->x) && coord_safe(A->y) && coord_safe(A->z) &&
			      coord_safe(B->x) && coord_safe(B->y) && coord_safe(B->z)))
				out->div_safe = 0;
			out->num_cubes++;
		}

		if (out->light_index < 0 && m->emission_power > 0) {   /* main.c:140-146 */
			out->light_index = i;
			if (o->type == RT_OBJECT_SPHERE)                   /* scene.c:10-15 */
				out->light_pos = o->sphere.center;
			else {
				out->light_pos.x = o->cube.origin.x * 1 + o->cube.size.x * 0.5f;
				out->light_pos.y = o->cube.origin.y * 1 + o->cube.size.y * 0.5f;
				out->light_pos.z = o->cube.origin.z * 1 + o->cube.size.z * 0.5f;
			}
		}

		RtF4 *M = &out->mat[(size_t) i * RT_MAT_STRIDE];
		float f0d = 0.16 * m->reflectance * m->reflectance;    /* main.c:219, in double, narrowed */
		float om = 1 - m->metallic;                            /* main.c:221 */
		float f0x = f0d * om + m->albedo.x * m->metallic;
		float f0y = f0d * om + m->albedo.y * m->metallic;
		float f0z = f0d * om + m->albedo.z * m->metallic;
		M[0].x = f0x; M[0].y = f0y; M[0].z = f0z; M[0].w = m->roughness;
		M[1].x = 1.0f * 1 + f0x * -1;                          /* main.c:128 combine(1, f0, 1, -1) */
		M[1].y = 1.0f * 1 + f0y * -1;
		M[1].z = 1.0f * 1 + f0z * -1;
		M[1].w = (m->metallic > 0.001) ? 1.0f : 0.0f;          /* main.c:241, compared in double */
		M[2].x = m->emission_color.x * m->emission_power;      /* main.c:203, 232 */
		M[2].y = m->emission_color.y * m->emission_power;
		M[2].z = m->emission_color.z * m->emission_power;
		M[2].w = m->emission_power;
		if (!(M[2].x == 0 && M[2].y == 0 && M[2].z == 0)) {    /* NaN counts as emitting */
			emitters++;
			out->only_emitter = i;
		}
		M[3].x = m->albedo.x * om;                             /* main.c:248 */
		M[3].y = m->albedo.y * om;
		M[3].z = m->albedo.z * om;
		M[3].w = 0;
	}
	if (emitters != 1) out->only_emitter = -1;
	return RT_OK;
}

void rt_host_free_packed(RtPackedScene *p)
{
	free(p->geomA);
	free(p->geomB);
	free(p->mat);
	p->geomA = p->geomB = p->mat = NULL;
}
