/*
 * rt_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of cozis/ray_tracing's
 * per-pixel render path, written from the behaviour documented in SURVEY.md
 * section 8(a) and checked bit-for-bit against the unmodified reference
 * (oracle/_ref/libref_pixel.so).  Build: oracle/Makefile `port`
 * (-O2 -ffp-contract=off: the reference is built for baseline x86-64 with
 * -std=c11, i.e. no FMA contraction, and so is this).
 *
 * Every float expression below keeps the reference's operand order and its
 * C promotions (where the reference computes in double, so does this).
 */
#include "rt_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

/* ---------------------------------------------------------------- vector.c */

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }

/* vector.c:145-152  combine(u,v,a,b) = u*a + v*b, two products then one sum */
static inline v3 lin2(v3 u, float a, v3 v, float b)
{
	return V(u.x * a + v.x * b, u.y * a + v.y * b, u.z * a + v.z * b);
}

/* vector.c:154-161  combine4: ((u*a + v*b) + g*c) + t*d */
static inline v3 lin4(v3 u, float a, v3 v, float b, v3 g, float c, v3 t, float d)
{
	return V(u.x * a + v.x * b + g.x * c + t.x * d,
	         u.y * a + v.y * b + g.y * c + t.y * d,
	         u.z * a + v.z * b + g.z * c + t.z * d);
}

static inline v3 scale3(v3 v, float f) { return V(v.x * f, v.y * f, v.z * f); }   /* vector.c:137-143 */
static inline v3 had3(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); } /* vector.c:366-373 */
static inline float dot3(v3 u, v3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; } /* vector.c:361-364 */

/* vector.c:163-170 */
static inline v3 cross3(v3 u, v3 v)
{
	return V(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}

/* vector.c:113-135: norm via double sqrt of the float sum of squares (== sqrtf),
 * guard compared in double against 1e-5, three float divisions. */
static inline v3 unit3(v3 v)
{
	float n = (float) sqrt((double) (v.x * v.x + v.y * v.y + v.z * v.z));
	if ((double) n < 0.00001 && (double) n > -0.00001)
		return v;
	return V(v.x / n, v.y / n, v.z / n);
}

static inline float clampf(float x, float lo, float hi)   /* vector.c:52-58 */
{
	if (x < lo) return lo;
	if (x > hi) return hi;
	return x;
}

static inline int near_zero(float f)                      /* vector.c:79-82, double compares */
{
	return (double) f < 0.0001 && (double) f > -0.0001;
}

/* ----------------------------------------------------------------- utils.c */

uint64_t rto_wyhash64(uint64_t *state)                    /* utils.c:62-70 */
{
	*state += 0x60bee2bee120fc15ull;
	__uint128_t t = (__uint128_t) *state * 0xa3b195354a39b70dull;
	uint64_t m1 = (uint64_t) (t >> 64) ^ (uint64_t) t;
	t = (__uint128_t) m1 * 0x1b03738712fad5c9ull;
	return (uint64_t) (t >> 64) ^ (uint64_t) t;
}

float rto_random_float(uint64_t *state)                   /* utils.c:72-75 */
{
	/* (float)u64 / UINT64_MAX: the integer constant converts to float 2^64 */
	return (float) rto_wyhash64(state) / (float) UINT64_MAX;
}

static v3 rand_dir(uint64_t *state)                       /* vector.c:99-111: x, y, z order */
{
	float x = rto_random_float(state) * 2 - 1;
	float y = rto_random_float(state) * 2 - 1;
	float z = rto_random_float(state) * 2 - 1;
	return unit3(V(x, y, z));
}

void rto_random_direction(uint64_t *state, float out[3])
{
	v3 d = rand_dir(state);
	out[0] = d.x; out[1] = d.y; out[2] = d.z;
}

static uint64_t splitmix64(uint64_t z)
{
	z += 0x9e3779b97f4a7c15ull;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}

/* Per-pixel stream key: the reference's stream is per worker thread and thus
 * partition dependent (SURVEY.md section 7 "RNG stream semantics"); the parity
 * contract re-keys the same generator at the top of pixel().  Identical to
 * oracle/ref_driver.c:pixel_key and ray_tracing_b200/csrc/rt_device.cuh. */
uint64_t rto_pixel_key(float px, float py, uint64_t pass)
{
	uint32_t bx, by;
	memcpy(&bx, &px, 4);
	memcpy(&by, &py, 4);
	return splitmix64((((uint64_t) bx << 32) | by) ^ splitmix64(pass));
}

/* ---------------------------------------------------------------- camera.c */

typedef struct { v3 origin, llc, horiz, vert; } CamFrame;

/* camera.c:99-118: everything that does not depend on (px,py) */
static CamFrame cam_frame(const RtoCamera *c, float aspect)
{
	v3 pos = V(c->pos[0], c->pos[1], c->pos[2]);
	v3 front = V(c->front[0], c->front[1], c->front[2]);
	v3 up = V(c->up[0], c->up[1], c->up[2]);
	v3 w = unit3(scale3(front, -1));
	v3 u = unit3(cross3(up, w));
	v3 v = cross3(w, u);
	float screen_h = (float) (2 * tan((double) (c->fov / 2)));   /* fov used as radians, camera.c:107 */
	float screen_w = aspect * screen_h;
	CamFrame f;
	f.origin = pos;
	f.horiz = scale3(u, screen_w);
	f.vert = scale3(v, screen_h);
	f.llc = lin4(pos, 1, f.horiz, -0.5f, f.vert, -0.5f, w, -1);
	return f;
}

static inline v3 cam_dir(const CamFrame *f, float px, float py)   /* camera.c:121 */
{
	return lin4(f->llc, 1, f->horiz, px, f->vert, py, f->origin, -1);
}

void rto_camera_ray(const RtoCamera *cam, float px, float py, float aspect, float out[6])
{
	CamFrame f = cam_frame(cam, aspect);
	v3 d = cam_dir(&f, px, py);
	out[0] = f.origin.x; out[1] = f.origin.y; out[2] = f.origin.z;
	out[3] = d.x; out[4] = d.y; out[5] = d.z;
}

/* ----------------------------------------------------------------- scene.c */

typedef struct { float t; v3 point, normal; int obj; } Hit;

/* scene.c:17-77.  Returns 1 and the entry distance + axis if the slabs overlap.
 * Real divisions; comparisons written so NaNs fall the same way. */
static int box_entry(v3 o, v3 d, const float *g, float *t_out, int *axis_out)
{
	float ax = g[0], ay = g[1], az = g[2];
	float bx = g[0] * 1 + g[3] * 1, by = g[1] * 1 + g[4] * 1, bz = g[2] * 1 + g[5] * 1;
	float lo, hi, l2, h2;
	int axis = 0;

	if (d.x >= 0) { lo = (ax - o.x) / d.x; hi = (bx - o.x) / d.x; }
	else          { hi = (ax - o.x) / d.x; lo = (bx - o.x) / d.x; }
	if (d.y >= 0) { l2 = (ay - o.y) / d.y; h2 = (by - o.y) / d.y; }
	else          { h2 = (ay - o.y) / d.y; l2 = (by - o.y) / d.y; }
	if (lo > h2 || l2 > hi) return 0;
	if (l2 > lo) { lo = l2; axis = 1; }
	if (h2 < hi) hi = h2;
	if (d.z >= 0) { l2 = (az - o.z) / d.z; h2 = (bz - o.z) / d.z; }
	else          { h2 = (az - o.z) / d.z; l2 = (bz - o.z) / d.z; }
	if (lo > h2 || l2 > hi) return 0;
	if (l2 > lo) { lo = l2; axis = 2; }
	*t_out = lo;
	*axis_out = axis;
	return 1;
}

/* scene.c:79-134.  Roots in double exactly as C promotes them. */
static int sphere_entry(v3 o, v3 d, const float *g, float *t_out)
{
	v3 c = V(g[0], g[1], g[2]);
	float r = g[3];
	v3 oc = lin2(c, 1, o, -1);
	float a = dot3(d, d);
	float b = -2 * dot3(oc, d);
	float cc = dot3(oc, oc) - r * r;
	float discr = b * b - 4 * a * cc;
	if (!(discr > 0)) return 0;
	float s0 = (float) (((double) (-b) + sqrt((double) discr)) / (double) (2 * a));
	float s1 = (float) (((double) (-b) - sqrt((double) discr)) / (double) (2 * a));
	if (s0 > s1) { float t = s0; s0 = s1; s1 = t; }
	if (s0 < 0) {
		s0 = s1;
		if (s0 < 0) return 0;
	}
	*t_out = s0;
	return 1;
}

/* scene.c:156-190: linear scan, strict '<' keeps the lowest index on ties */
static Hit nearest_hit(const RtoObject *objs, int n, v3 o, v3 dir)
{
	v3 d = unit3(dir);
	float best = FLT_MAX;
	int best_i = -1, best_axis = 0;
	for (int i = 0; i < n; i++) {
		float t;
		int axis = 0;
		if (objs[i].type == 1) {
			if (!sphere_entry(o, d, objs[i].geom, &t)) continue;
		} else if (objs[i].type == 0) {
			if (!box_entry(o, d, objs[i].geom, &t, &axis)) continue;
		} else
			continue;
		if (t >= 0 && t < best) { best = t; best_i = i; best_axis = axis; }
	}
	Hit h;
	h.obj = best_i;
	if (best_i < 0) {
		h.t = -1; h.point = V(0, 0, 0); h.normal = V(0, 0, 0);
		return h;
	}
	h.t = best;
	h.point = lin2(o, 1, d, best);
	if (objs[best_i].type == 1) {
		/* scene.c:146-147: normal of the nearest sphere (pure function of the
		 * ray and t, so evaluating it only for the winner changes nothing) */
		const float *g = objs[best_i].geom;
		h.normal = unit3(lin2(lin2(o, 1, d, best), 1, V(g[0], g[1], g[2]), -1));
	} else {
		/* scene.c:70-74 */
		float dc = best_axis == 0 ? d.x : best_axis == 1 ? d.y : d.z;
		float s = dc > 0 ? -1.0f : 1.0f;
		h.normal = V(best_axis == 0 ? s : 0, best_axis == 1 ? s : 0, best_axis == 2 ? s : 0);
	}
	return h;
}

void rto_trace_many(const RtoObject *objects, int n, const float *rays, int nrays,
                    float *out7, int32_t *obj)
{
	for (int i = 0; i < nrays; i++) {
		const float *r = rays + 6 * (size_t) i;
		Hit h = nearest_hit(objects, n, V(r[0], r[1], r[2]), V(r[3], r[4], r[5]));
		float *o = out7 + 7 * (size_t) i;
		o[0] = h.t;
		o[1] = h.point.x;  o[2] = h.point.y;  o[3] = h.point.z;
		o[4] = h.normal.x; o[5] = h.normal.y; o[6] = h.normal.z;
		obj[i] = h.obj;
	}
}

/* scene.c:10-15 */
static v3 object_origin(const RtoObject *o)
{
	if (o->type == 1) return V(o->geom[0], o->geom[1], o->geom[2]);
	return lin2(V(o->geom[0], o->geom[1], o->geom[2]), 1, V(o->geom[3], o->geom[4], o->geom[5]), 0.5f);
}

/* --------------------------------------------- gpu_and_windowing.c:42-112 */

static v3 sky_lookup(const RtoSky *s, v3 dir)
{
	float ax = dir.x < 0 ? -dir.x : dir.x;
	float ay = dir.y < 0 ? -dir.y : dir.y;
	float az = dir.z < 0 ? -dir.z : dir.z;
	int face;
	float u, v;
	if (ax > ay && ax > az) {
		if (dir.x > 0) { face = 3; u = -dir.z / (ax + 0.0f); v = -dir.y / (ax + 0.0f); }
		else           { face = 2; u =  dir.z / (ax + 0.0f); v = -dir.y / (ax + 0.0f); }
	} else if (ay > ax && ay > az) {
		if (dir.y > 0) { face = 4; u = dir.x / (ay + 0.0f); v =  dir.z / (ay + 0.0f); }
		else           { face = 5; u = dir.x / (ay + 0.0f); v = -dir.z / (ay + 0.0f); }
	} else {
		if (dir.z > 0) { face = 0; u =  dir.x / (az + 0.0f); v = -dir.y / (az + 0.0f); }
		else           { face = 1; u = -dir.x / (az + 0.0f); v = -dir.y / (az + 0.0f); }
	}
	u = clampf(u, -1, 1);
	v = clampf(v, -1, 1);
	u = 0.5f * (u + 1.0f);
	v = 0.5f * (v + 1.0f);
	int x = (int) (u * (float) (s->w - 1));
	int y = (int) (v * (float) (s->h - 1));
	const uint8_t *p = s->face[face] + ((size_t) y * s->w + x) * s->chan;
	return V((float) p[0] / 255, (float) p[1] / 255, (float) p[2] / 255);
}

void rto_sample_cubemap_many(const RtoSky *sky, const float *dirs, int n, float *out)
{
	for (int i = 0; i < n; i++) {
		v3 c = sky_lookup(sky, V(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
		out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
	}
}

/* ------------------------------------------------------ main.c:126-272 */

static v3 radiance(const RtoWorld *w, const CamFrame *cf, float px, float py,
                   uint64_t *rng, uint64_t *rays)
{
	const RtoObject *objs = w->objects;
	int n = w->num_objects;
	v3 o = cf->origin;
	v3 d = cam_dir(cf, px, py);          /* NOT normalised (camera.c:121) */

	int light = -1;                      /* main.c:140-146 */
	for (int i = 0; i < n; i++)
		if (objs[i].emission_power > 0) { light = i; break; }

	v3 contrib = V(1, 1, 1), result = V(0, 0, 0);

	for (int bounce = 0; bounce < 10; bounce++) {
		Hit hit = nearest_hit(objs, n, o, d);
		++*rays;
		if (hit.obj < 0) {               /* main.c:162-173 */
			v3 sky = sky_lookup(&w->sky, unit3(d));
			result = lin2(result, 1, had3(sky, contrib), 1);
			break;
		}

		v3 sampled = V(0, 0, 0);         /* main.c:180-210 */
		if (light >= 0) {
			v3 to_light = lin2(object_origin(&objs[light]), 1, hit.point, -1);
			int got = 0;
			for (int k = 0; k < 3; k++) {
				v3 rd = rand_dir(rng);
				if (dot3(rd, hit.normal) <= 0) continue;
				v3 sd = unit3(lin2(rd, 0.5f, to_light, 1));
				v3 so = lin2(hit.point, 1, sd, 0.001f);
				Hit h2 = nearest_hit(objs, n, so, sd);
				++*rays;
				if (h2.obj >= 0) {
					const RtoObject *m = &objs[h2.obj];
					sampled = lin2(sampled, 1,
					               V(m->emission_color[0], m->emission_color[1], m->emission_color[2]),
					               m->emission_power);
				}
				got++;
			}
			if (got > 0) sampled = scale3(sampled, 1.0f / got);
		}

		const RtoObject *m = &objs[hit.obj];
		v3 albedo = V(m->albedo[0], m->albedo[1], m->albedo[2]);
		v3 emis = V(m->emission_color[0], m->emission_color[1], m->emission_color[2]);

		v3 view = scale3(d, -1);         /* main.c:214-216 */
		float NoV = clampf(dot3(hit.normal, view), 0, 1);

		float f0d = (float) (0.16 * (double) m->reflectance * (double) m->reflectance);   /* main.c:219 */
		v3 f0 = lin2(V(f0d, f0d, f0d), 1 - m->metallic, albedo, m->metallic);
		float pw = (float) pow(1.0 - (double) NoV, 5.0);                                   /* main.c:128 */
		v3 F = lin2(f0, 1, lin2(V(1, 1, 1), 1, f0, -1), pw);

		v3 rd = rand_dir(rng);           /* main.c:226-228 */
		if (dot3(rd, hit.normal) < 0) rd = scale3(rd, -1);

		result = lin2(result, 1, had3(scale3(emis, m->emission_power), contrib), 1);      /* main.c:232 */

		v3 out;
		/* main.c:241: short-circuit -- the float is drawn only for non-metals */
		if ((double) m->metallic > 0.001 || rto_random_float(rng) <= (F.x + F.y + F.z) / 3) {
			v3 nn = scale3(hit.normal, -1);
			float f = -2 * dot3(nn, d);  /* vector.c:107-111 reflect(dir, normal) */
			v3 refl = lin2(d, 1, nn, f);
			out = unit3(lin2(rd, m->roughness, refl, 1));
		} else {
			out = rd;
			contrib = had3(contrib, scale3(albedo, 1 - m->metallic));
		}
		o = lin2(hit.point, 1, out, 0.001f);

		if (!(near_zero(sampled.x) && near_zero(sampled.y) && near_zero(sampled.z))) {    /* main.c:257-261 */
			float wgt = 0.05f;
			result = lin2(result, 1, had3(sampled, contrib), wgt);
			contrib = scale3(contrib, 1 - wgt);
		}
		d = out;
	}
	return V(clampf(result.x, 0, 1), clampf(result.y, 0, 1), clampf(result.z, 0, 1));
}

void rto_pixel(const RtoWorld *w, float px, float py, float aspect, uint64_t rng_state,
               float out[3], uint64_t *rays)
{
	CamFrame cf = cam_frame(&w->camera, aspect);
	uint64_t r = 0, st = rng_state;
	v3 c = radiance(w, &cf, px, py, &st, &r);
	out[0] = c.x; out[1] = c.y; out[2] = c.z;
	if (rays) *rays = r;
}

/* ------------------------------------------------------ main.c:274-322 */

typedef struct {
	const RtoWorld *w;
	float *out;
	int W, H, scale, ncols, lrow0, lrow1;
	uint64_t pass, rays;
} Band;

static void *band_main(void *arg)
{
	Band *b = (Band *) arg;
	int W = b->W, H = b->H, s = b->scale;
	float aspect = (float) W / H;                   /* main.c:281 */
	CamFrame cf = cam_frame(&b->w->camera, aspect);
	int lw = W / s, lh = H / s;                     /* main.c:284-285 */
	int colw = W / b->ncols;
	uint64_t rays = 0;
	for (int j = b->lrow0; j < b->lrow1 && j < lh; j++) {
		float v = (float) j / (lh - 1);             /* main.c:294,296 */
		v = 1 - v;
		for (int c = 0; c < b->ncols; c++) {
			int colx = colw * c;
			int lcx = colx / s;                     /* main.c:287 */
			/* the reference also traces one clipped pixel per row and column
			 * (main.c:286,302-303); it writes nothing, so it is skipped here */
			for (int i = 0; i * s < colw; i++) {
				float u = (float) (lcx + i) / (lw - 1);   /* main.c:293,295 */
				u = 1 - u;
				uint64_t rng = rto_pixel_key(u, v, b->pass);
				v3 col = radiance(b->w, &cf, u, v, &rng, &rays);
				int tw = s;
				if (tw > colw - i * s) tw = colw - i * s;
				for (int g = 0; g < s; g++)
					for (int t = 0; t < tw; t++) {
						size_t p = (size_t) (j * s + g) * W + (size_t) (colx + i * s + t);
						b->out[3 * p] = col.x; b->out[3 * p + 1] = col.y; b->out[3 * p + 2] = col.z;
					}
			}
		}
	}
	b->rays = rays;
	return NULL;
}

uint64_t rto_render(const RtoWorld *w, float *out, int W, int H, int scale, int num_columns,
                    uint64_t pass, int row0, int row1, int nthreads)
{
	if (num_columns < 1) num_columns = 1;
	if (nthreads < 1) nthreads = 1;
	if (row0 < 0) row0 = 0;
	if (row1 > H) row1 = H;
	if (row1 <= row0) return 0;
	/* rows are produced in units of low-res rows; the band must be scale aligned */
	int l0 = row0 / scale, l1 = (row1 + scale - 1) / scale;
	memset(out + 3 * (size_t) row0 * W, 0, sizeof(float) * 3 * (size_t) (row1 - row0) * W);
	if (nthreads > l1 - l0) nthreads = l1 - l0 > 0 ? l1 - l0 : 1;

	Band *bands = calloc((size_t) nthreads, sizeof(Band));
	pthread_t *th = calloc((size_t) nthreads, sizeof(pthread_t));
	/* interleave low-res rows in chunks so sky-heavy rows spread over threads */
	int per = (l1 - l0 + nthreads - 1) / nthreads;
	for (int t = 0; t < nthreads; t++) {
		bands[t].w = w; bands[t].out = out; bands[t].W = W; bands[t].H = H;
		bands[t].scale = scale; bands[t].ncols = num_columns; bands[t].pass = pass;
		bands[t].lrow0 = l0 + t * per;
		bands[t].lrow1 = l0 + (t + 1) * per < l1 ? l0 + (t + 1) * per : l1;
		pthread_create(&th[t], NULL, band_main, &bands[t]);
	}
	uint64_t rays = 0;
	for (int t = 0; t < nthreads; t++) {
		pthread_join(th[t], NULL);
		rays += bands[t].rays;
	}
	free(bands); free(th);
	return rays;
}

/* ------------------------------------------------ main.c:387-396, 467-477 */

void rto_accumulate(float *accum, const float *data, size_t n, int scale)
{
	float wgt = 1.0f / (scale * scale);
	for (size_t i = 0; i < n; i++)
		accum[i] = accum[i] * 1 + data[i] * wgt;
}

void rto_resolve(float *frame, const float *accum, size_t n, float count)
{
	float inv = 1.0f / count;
	for (size_t i = 0; i < n; i++)
		frame[i] = accum[i] * inv;
}
