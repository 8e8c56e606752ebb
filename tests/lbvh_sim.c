/*
 * lbvh_sim.c -- CPU proof harness for the LBVH rule (ray_tracing_b200/csrc/
 * rt_lbvh_rule.h).  TEST INFRASTRUCTURE (built and run by
 * tests/test_lbvh_rule_cpu.py), not product: it links the oracle
 * (oracle/librt_oracle.so) for the ground truth.
 *
 * It builds the same tree the device builds (Morton keys in binary32, Karras
 * 2012 hierarchy, tight boxes + per-node weight), walks it with the same
 * arithmetic as rt_device.cuh (nearer child first, per-node widening from the
 * ray origin's distance, cull at best + slack, ties to the lower index) and
 * compares every ray with the reference's O(N) scan as restated by the oracle
 * (rto_trace_many: scene.c:156-190).  Rays are the ones a frame actually
 * traces: primary rays of a camera, then for each hit one light-sample ray
 * (main.c:197-198) and one bounce ray (main.c:226-250), for a few generations.
 *
 *   lbvh_sim objects.bin W H generations [cam px py pz fx fy fz] [--dynamic-pad]
 *
 * objects.bin = N records of 68 bytes (RtoObject).  Prints one JSON line:
 * rays, mismatches, internal-node visits and primitive tests per ray, deepest
 * stack.  Default = the product's rule (rt_lbvh_rule.h: every box padded for
 * the bounds' diagonal).  --dynamic-pad = the measured-and-rejected alternative:
 * tight boxes, widened per visited node by the fuzz the actual distance from
 * the ray origin allows (see rt_lbvh_rule.h for the numbers).
 * Environment: SIM_TOPOLOGY=sah|karras (the product's two builders), SIM_FMA / SIM_PACK (the device's
 * slab arithmetic and 16-bit boxes), SIM_ANYHIT (light samples as the device walks them), SIM_AXIS,
 * SIM_RAYS, SIM_DMAX, SIM_SLACK; experiments: SIM_LEAF (multi-primitive leaves), SIM_WIDE (what a
 * 4-wide collapse of the tree would visit and test: DESIGN.md 8).
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../oracle/rt_oracle.h"
#include "../ray_tracing_b200/csrc/rt_lbvh_rule.h"

typedef struct { float x, y, z, w; } f4;

/* ---- 16-bit fixed point, rounded outwards plus one quantum (rt_lbvh.cu: pack_lo / pack_hi).
 * The value kept in the tree is the binary32 number Q = 2^23 + q that the walk's PRMT makes. ---- */
static float clampq(float t) { return fminf(fmaxf(t, 0.0f), 65535.0f); }
static float pack_lo(float v, float c, float s) { return 8388608.0f + clampq(floorf((v - c) * s) - 1.0f + 32768.0f); }
static float pack_hi(float v, float c, float s) { return 8388608.0f + clampq(ceilf((v - c) * s) + 1.0f + 32768.0f); }

/* ---- the rejected per-node rule (experiment only) ---- */
#define DYN_KE       (32.0f * 0x1p-24f)         /* K eps = 2^-19 */
#define DYN_SQRT_KE  0x1.6a09e8p-10f            /* sqrt(K eps), rounded up */
#define DYN_E        0x1p-18f

static float dyn_sphere_weight(float r2)        /* K eps / 2r: sqrt(r^2 + x) - r <= x / 2r */
{
	float r = sqrtf(r2 > 0.0f ? r2 : 0.0f);
	return (DYN_KE * 0.5f) / r * 1.000001f;
}

/* fx, fy, fz: per-axis distance from the ray origin to the farthest face of the
 * node's boxes; w: largest weight below the node; emag = DYN_E * mag */
static float dyn_pad(float fx, float fy, float fz, float w, float emag, float *slack)
{
	float d2 = fmaf(fx, fx, fmaf(fy, fy, fz * fz));
	float d1 = fx + fy + fz;
	float base = fmaf(DYN_E, d1, emag);
	float p = fminf(d2 * w, DYN_SQRT_KE * d1);  /* fminf drops a NaN (0 * inf) */
	*slack = base + base;
	return p + base;
}

typedef struct {
	int n;
	f4 *nodes;          /* 4 per internal node, device layout (rt_params.h) */
	int *prim;          /* Morton slot -> primitive */
	f4 *leaf_lo, *leaf_hi;
	float *leaf_w;
	float emag;
	int depth;          /* deepest leaf, levels below the root */
	int packed;         /* SIM_PACK: boxes rounded outwards to 16-bit fixed point in the frame below (rt_params.h) */
	float cx, cy, cz, scale, inv_scale;
	int global_pad;     /* the product's static rule */
	float t_slack;
} Tree;

static unsigned spread10(unsigned v)
{
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}

static int cmp_u64(const void *a, const void *b)
{
	uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b;
	return x < y ? -1 : x > y;
}

static int delta(const uint64_t *k, int n, int i, int j)
{
	if (j < 0 || j >= n) return -1;
	return __builtin_clzll(k[i] ^ k[j]);
}

static float fmin2(float a, float b) { return fminf(a, b); }
static float fmax2(float a, float b) { return fmaxf(a, b); }


/* SIM_TOPOLOGY=sah: the product's host-built topology instead of the Karras hierarchy */
#include "../ray_tracing_b200/csrc/bvh_sah.c"

static void build(Tree *T, const RtoObject *obj, int n, int global_pad)
{
	memset(T, 0, sizeof(*T));
	T->n = n;
	T->global_pad = global_pad;
	f4 *A = malloc(sizeof(f4) * n), *B = malloc(sizeof(f4) * n);
	float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
	for (int i = 0; i < n; i++) {
		const float *g = obj[i].geom;
		if (obj[i].type == 1) {
			A[i] = (f4){g[0], g[1], g[2], g[3] * g[3]};
			B[i] = (f4){0, 0, 0, 1};
			float r = fabsf(g[3]);
			for (int k = 0; k < 3; k++) { lo[k] = fmin2(lo[k], g[k] - r); hi[k] = fmax2(hi[k], g[k] + r); }
		} else {
			A[i] = (f4){g[0], g[1], g[2], 0};
			B[i] = (f4){g[0] * 1 + g[3] * 1, g[1] * 1 + g[4] * 1, g[2] * 1 + g[5] * 1, 0};
			const float *a = &A[i].x, *b = &B[i].x;
			for (int k = 0; k < 3; k++) { lo[k] = fmin2(lo[k], fmin2(a[k], b[k])); hi[k] = fmax2(hi[k], fmax2(a[k], b[k])); }
		}
	}
	float ext[3], inv[3];
	for (int k = 0; k < 3; k++) { ext[k] = hi[k] - lo[k]; inv[k] = ext[k] > 0 ? 1.0f / ext[k] : 0.0f; }
	uint64_t *keys = malloc(sizeof(uint64_t) * n);
	for (int i = 0; i < n; i++) {
		float c[3];
		if (B[i].w == 1) { c[0] = A[i].x; c[1] = A[i].y; c[2] = A[i].z; }
		else { c[0] = 0.5f * (A[i].x + B[i].x); c[1] = 0.5f * (A[i].y + B[i].y); c[2] = 0.5f * (A[i].z + B[i].z); }
		unsigned q[3];
		for (int k = 0; k < 3; k++) {
			float x = fmin2(fmax2((c[k] - lo[k]) * inv[k], 0.0f), 1.0f);
			unsigned v = (unsigned) (x * 1024.0f);
			q[k] = v > 1023u ? 1023u : v;
		}
		unsigned code = (spread10(q[0]) << 2) | (spread10(q[1]) << 1) | spread10(q[2]);
		keys[i] = ((uint64_t) code << 32) | (unsigned) i;
	}
	qsort(keys, n, sizeof(uint64_t), cmp_u64);
	T->prim = malloc(sizeof(int) * n);
	for (int i = 0; i < n; i++) T->prim[i] = (int) (keys[i] & 0xffffffffu);

	int *children = malloc(sizeof(int) * 2 * (n > 1 ? n - 1 : 1));
	int *parent = malloc(sizeof(int) * (2 * n));
	int *first = malloc(sizeof(int) * (n > 1 ? n - 1 : 1)), *count = malloc(sizeof(int) * (n > 1 ? n - 1 : 1));
	int leaf_max = getenv("SIM_LEAF") ? atoi(getenv("SIM_LEAF")) : 1;      /* experiment: collapse small subtrees into one leaf */
	for (int i = 0; i < n - 1; i++) {
		int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
		int dmin = delta(keys, n, i, i - d);
		int lmax = 2;
		while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
		int l = 0;
		for (int t = lmax / 2; t >= 1; t /= 2)
			if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
		int j = i + l * d;
		int dnode = delta(keys, n, i, j);
		int s = 0;
		for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
			if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
			if (t == 1) break;
		}
		int gamma = i + s * d + (d < 0 ? d : 0);
		int lo_i = i < j ? i : j, hi_i = i < j ? j : i;
		int left = (lo_i == gamma) ? ~gamma : gamma;
		int right = (hi_i == gamma + 1) ? ~(gamma + 1) : gamma + 1;
		children[2 * i] = left;
		children[2 * i + 1] = right;
		first[i] = lo_i;
		count[i] = hi_i - lo_i + 1;
		parent[left >= 0 ? left : (n - 1) + ~left] = i;
		parent[right >= 0 ? right : (n - 1) + ~right] = i;
	}
	parent[0] = -1;
	if (getenv("SIM_TOPOLOGY") && !strcmp(getenv("SIM_TOPOLOGY"), "sah") && n >= 2) {
		double mag0 = 0;
		for (int k = 0; k < 3; k++) mag0 = fmax(mag0, fmax(fabs(lo[k]), fabs(hi[k])));
		RtLbvhPads p0 = rt_lbvh_pads(mag0, rt_lbvh_default_dmax(ext[0], ext[1], ext[2]), RT_LBVH_FUZZ_K, RT_LBVH_SLACK);
		int sah_depth;
		RtF4 *Bt = malloc(sizeof(RtF4) * n);               /* the device layout keeps the type as int bits in .w */
		for (int i = 0; i < n; i++) {
			int ty = B[i].w == 1 ? RT_OBJECT_SPHERE : RT_OBJECT_CUBE;
			Bt[i] = (RtF4){B[i].x, B[i].y, B[i].z, 0};
			memcpy(&Bt[i].w, &ty, 4);
		}
		if (rt_host_bvh_sah((const RtF4 *) A, Bt, n, p0.fuzz_r2, p0.cube_pad, p0.extra, T->prim, children, parent, &sah_depth)) exit(2);
		free(Bt);
		for (int i = 0; i < n - 1; i++) count[i] = n;      /* no SIM_LEAF collapsing on this topology */
	}

	for (int sl = 0; sl < n && n >= 2; sl++) {
		int d = 0;
		for (int cur = parent[(n - 1) + sl]; cur >= 0; cur = parent[cur]) d++;
		if (d > T->depth) T->depth = d;
	}

	/* leaf boxes */
	double mag = 0;
	for (int k = 0; k < 3; k++) mag = fmax(mag, fmax(fabs(lo[k]), fabs(hi[k])));
	T->emag = DYN_E * (float) mag;
	double dx = ext[0], dy = ext[1], dz = ext[2];
	float d_max = rt_lbvh_default_dmax(dx, dy, dz);
	if (getenv("SIM_DMAX")) d_max = (float) atof(getenv("SIM_DMAX"));      /* what rt_api.cu does for a far camera */
	RtLbvhPads pads = rt_lbvh_pads(mag, d_max, RT_LBVH_FUZZ_K, getenv("SIM_SLACK") ? atof(getenv("SIM_SLACK")) : RT_LBVH_SLACK);
	double fuzz_r2 = pads.fuzz_r2;
	float cube_pad = pads.cube_pad, extra = pads.extra;
	T->t_slack = pads.t_slack;
	T->leaf_lo = malloc(sizeof(f4) * n);
	T->leaf_hi = malloc(sizeof(f4) * n);
	T->leaf_w = malloc(sizeof(float) * n);
	for (int s = 0; s < n; s++) {
		int p = T->prim[s];
		f4 a = A[p], b = B[p];
		if (b.w == 1) {
			float rp;
			if (global_pad) {
				rp = (float) sqrt((double) fmax2(a.w, 0.0f) + fuzz_r2);
				rp = rp * 1.000001f + extra;
			} else
				rp = sqrtf(fmax2(a.w, 0.0f)) * 1.000001f;
			T->leaf_lo[s] = (f4){a.x - rp, a.y - rp, a.z - rp, 0};
			T->leaf_hi[s] = (f4){a.x + rp, a.y + rp, a.z + rp, 0};
			T->leaf_w[s] = dyn_sphere_weight(a.w);
		} else {
			float pad = global_pad ? cube_pad + extra : 0.0f;
			T->leaf_lo[s] = (f4){fmin2(a.x, b.x) - pad, fmin2(a.y, b.y) - pad, fmin2(a.z, b.z) - pad, 0};
			T->leaf_hi[s] = (f4){fmax2(a.x, b.x) + pad, fmax2(a.y, b.y) + pad, fmax2(a.z, b.z) + pad, 0};
			T->leaf_w[s] = 0.0f;
		}
	}
	/* bottom-up refit: process internal nodes in an order where children come first */
	T->nodes = calloc(4 * (size_t) (n > 1 ? n - 1 : 1), sizeof(f4));
	if (n >= 2) {
		f4 *blo = malloc(sizeof(f4) * (n - 1)), *bhi = malloc(sizeof(f4) * (n - 1));
		float *bw = malloc(sizeof(float) * (n - 1));
		int *visit = calloc(n - 1, sizeof(int));
		for (int s = 0; s < n; s++) {
			int cur = parent[(n - 1) + s];
			while (cur >= 0) {
				if (visit[cur]++ == 0) break;
				int cl = children[2 * cur], cr = children[2 * cur + 1];
				f4 llo = cl < 0 ? T->leaf_lo[~cl] : blo[cl], lhi = cl < 0 ? T->leaf_hi[~cl] : bhi[cl];
				f4 rlo = cr < 0 ? T->leaf_lo[~cr] : blo[cr], rhi = cr < 0 ? T->leaf_hi[~cr] : bhi[cr];
				float wl = cl < 0 ? T->leaf_w[~cl] : bw[cl], wr = cr < 0 ? T->leaf_w[~cr] : bw[cr];
				f4 *nd = T->nodes + 4 * (size_t) cur;
				/* a subtree of at most leaf_max primitives becomes one leaf (a run of
				 * Morton slots): code ~(first | (count - 1) << 27) */
				int32_t icl = cl >= 0 && count[cl] <= leaf_max ? ~(first[cl] | (count[cl] - 1) << 27) : cl;
				int32_t icr = cr >= 0 && count[cr] <= leaf_max ? ~(first[cr] | (count[cr] - 1) << 27) : cr;
				nd[0] = llo; memcpy(&nd[0].w, &icl, 4);
				nd[1] = lhi;
				nd[2] = rlo; memcpy(&nd[2].w, &icr, 4);
				nd[3] = rhi;
				nd[1].w = fmax2(wl, wr);             /* the node's weight */
				blo[cur] = (f4){fmin2(llo.x, rlo.x), fmin2(llo.y, rlo.y), fmin2(llo.z, rlo.z), 0};
				bhi[cur] = (f4){fmax2(lhi.x, rhi.x), fmax2(lhi.y, rhi.y), fmax2(lhi.z, rhi.z), 0};
				bw[cur] = fmax2(wl, wr);
				cur = parent[cur];
			}
		}
		free(blo); free(bhi); free(bw); free(visit);
	}
	if (getenv("SIM_PACK") && global_pad && n >= 2) {
		/* rt_lbvh.cu: rt_lbvh_refit(): the frame, then every box rounded outwards to 16-bit fixed point */
		double pad = sqrt(fuzz_r2) + extra + cube_pad;
		double hx = 0.5 * ((double) hi[0] - lo[0]) + pad, hy = 0.5 * ((double) hi[1] - lo[1]) + pad, hz = 0.5 * ((double) hi[2] - lo[2]) + pad;
		double half = fmax(fmax(hx, hy), fmax(hz, 1e-30));
		int e = (int) floor(log2(32000.0 / half));
		if (e > 100) e = 100;
		if (e < -100) e = -100;
		T->packed = 1;
		T->scale = (float) ldexp(1.0, e);
		T->inv_scale = (float) ldexp(1.0, -e);
		T->cx = (float) (0.5 * ((double) hi[0] + lo[0]));
		T->cy = (float) (0.5 * ((double) hi[1] + lo[1]));
		T->cz = (float) (0.5 * ((double) hi[2] + lo[2]));
		for (int i = 0; i < n - 1; i++)
			for (int k = 0; k < 4; k += 2) {
				f4 *l = &T->nodes[4 * (size_t) i + k], *h = l + 1;
				l->x = pack_lo(l->x, T->cx, T->scale); l->y = pack_lo(l->y, T->cy, T->scale); l->z = pack_lo(l->z, T->cz, T->scale);
				h->x = pack_hi(h->x, T->cx, T->scale); h->y = pack_hi(h->y, T->cy, T->scale); h->z = pack_hi(h->z, T->cz, T->scale);
			}
	}
	free(A); free(B); free(keys); free(children); free(parent); free(first); free(count);
}

typedef struct { float t; int obj; } Best;

/* rt_device.cuh: node_overlap with widened boxes */
static int g_fma;       /* SIM_FMA: slab distances as fma(plane, inv, -(o*inv)) */

static int g_wide;      /* SIM_WIDE */
static long g_wide_boxes;
static int g_packed;    /* the tree holds Q = 2^23 + q; o[] holds oi = fma(K, inv, o' * inv) (rt_device.cuh: walk_ray) */

static int overlap(f4 lo, f4 hi, const float o[3], const float inv[3], float pad, float tmax, float *tn)
{
	if (g_packed) {
		/* rt_device.cuh: node_span(): near / far plane picked by the sign bit of inv */
		float nx = fmaf(signbit(inv[0]) ? hi.x : lo.x, inv[0], -o[0]), fx = fmaf(signbit(inv[0]) ? lo.x : hi.x, inv[0], -o[0]);
		float ny = fmaf(signbit(inv[1]) ? hi.y : lo.y, inv[1], -o[1]), fy = fmaf(signbit(inv[1]) ? lo.y : hi.y, inv[1], -o[1]);
		float nz = fmaf(signbit(inv[2]) ? hi.z : lo.z, inv[2], -o[2]), fz = fmaf(signbit(inv[2]) ? lo.z : hi.z, inv[2], -o[2]);
		*tn = fmax2(fmax2(nx, ny), fmax2(nz, 0.0f));
		float tf = fmin2(fmin2(fx, fy), fmin2(fz, tmax));
		return *tn <= tf;
	}
	if (g_fma) {
		float ox = o[0] * inv[0], oy = o[1] * inv[1], oz = o[2] * inv[2];
		float tx1 = fmaf(lo.x, inv[0], -ox), tx2 = fmaf(hi.x, inv[0], -ox);
		float ty1 = fmaf(lo.y, inv[1], -oy), ty2 = fmaf(hi.y, inv[1], -oy);
		float tz1 = fmaf(lo.z, inv[2], -oz), tz2 = fmaf(hi.z, inv[2], -oz);
		*tn = fmax2(fmax2(fmin2(tx1, tx2), fmin2(ty1, ty2)), fmax2(fmin2(tz1, tz2), 0.0f));
		float tf = fmin2(fmin2(fmax2(tx1, tx2), fmax2(ty1, ty2)), fmin2(fmax2(tz1, tz2), tmax));
		return *tn <= tf;
	}
	float tx1 = ((lo.x - o[0]) - pad) * inv[0], tx2 = ((hi.x - o[0]) + pad) * inv[0];
	float ty1 = ((lo.y - o[1]) - pad) * inv[1], ty2 = ((hi.y - o[1]) + pad) * inv[1];
	float tz1 = ((lo.z - o[2]) - pad) * inv[2], tz2 = ((hi.z - o[2]) + pad) * inv[2];
	*tn = fmax2(fmax2(fmin2(tx1, tx2), fmin2(ty1, ty2)), fmax2(fmin2(tz1, tz2), 0.0f));
	float tf = fmin2(fmin2(fmax2(tx1, tx2), fmax2(ty1, ty2)), fmin2(fmax2(tz1, tz2), tmax));
	return *tn <= tf;
}

static float far3(float a, float b, float c, float d)
{
	return fmax2(fmax2(fabsf(a), fabsf(b)), fmax2(fabsf(c), fabsf(d)));
}

/* anyhit_light >= 0: shadow-ray mode (rt_device.cuh: the only object whose emission is not zero is
 * `anyhit_light`): the light is tested first, then the walk only has to find ANY other object
 * in front of it (or tying with a lower index) and stops there.  out->obj = the light if it is
 * the nearest hit, otherwise -2 (occluded) or -1 (the light is not hit: nothing to add). */
static void walk(const Tree *T, const RtoObject *obj, const float ray[6], Best *out, long *nodes, long *tests, int *deepest, int anyhit_light)
{
	/* the primitive test is the oracle's own: one-object scan (which normalises
	 * the direction exactly as trace_ray does, scene.c:158) */
	float o[3] = {ray[0], ray[1], ray[2]};
	float out7[7];
	int32_t hit;
	/* normalised direction for the slabs */
	float n = sqrtf(ray[3] * ray[3] + ray[4] * ray[4] + ray[5] * ray[5]);
	float d[3] = {ray[3], ray[4], ray[5]};
	if (!(n <= 0x1.4f8b58p-17f && n >= -0x1.4f8b58p-17f)) { d[0] /= n; d[1] /= n; d[2] /= n; }
	float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
	if (g_fma)      /* rt_device.cuh: walk_inverse() keeps the reciprocals finite for the fma form */
		for (int k = 0; k < 3; k++) inv[k] = copysignf(fminf(fabsf(inv[k]) * (T->packed ? T->inv_scale : 1.0f), 0x1p100f), inv[k]);
	float ow[3] = {o[0], o[1], o[2]};        /* world origin for nothing but clarity: primitives use `ray` */
	(void) ow;
	if (T->packed) {
		/* rt_device.cuh: walk_ray(): o[] becomes oi */
		const float K = 8421376.0f;
		o[0] = fmaf(K, inv[0], (o[0] - T->cx) * T->scale * inv[0]);
		o[1] = fmaf(K, inv[1], (o[1] - T->cy) * T->scale * inv[1]);
		o[2] = fmaf(K, inv[2], (o[2] - T->cz) * T->scale * inv[2]);
	}
	Best best = {FLT_MAX, -1};
	if (anyhit_light >= 0) {
		(*tests)++;
		rto_trace_many(&obj[anyhit_light], 1, ray, 1, out7, &hit);
		if (hit < 0) { *out = best; return; }
		best.t = out7[0];
		best.obj = anyhit_light;
	}
	int stack[128], sp = 0;
	float stack_t[128];
	int pop_cull = getenv("SIM_POP_CULL") != NULL;
	int node = T->n == 1 ? ~0 : 0;
	if (T->n <= 0) { *out = best; return; }
	for (;;) {
		if (node < 0) {
			int code = ~node, slot0 = code & ((1 << 27) - 1), cnt = (code >> 27) + 1;
			for (int k = 0; k < cnt; k++) {
				int p = T->prim[slot0 + k];
				if (p == anyhit_light) continue;
				(*tests)++;
				rto_trace_many(&obj[p], 1, ray, 1, out7, &hit);
				if (hit >= 0) {
					float t = out7[0];
					if (t < best.t || (t == best.t && p < best.obj)) {
						if (anyhit_light >= 0) { out->t = t; out->obj = -2; return; }
						best.t = t; best.obj = p;
					}
				}
			}
		} else if (g_wide) {
			/* SIM_WIDE: what a 4-wide collapse of this tree would do (experiment for DESIGN.md 8): a visit
			 * tests the boxes of the node's grandchildren (a child that is a leaf stands for itself), hit
			 * ones are entered near-first.  nodes = wide-node visits, g_wide_boxes = boxes tested. */
			(*nodes)++;
			const f4 *nb = T->nodes + 4 * (size_t) node;
			f4 clo[4], chi[4];
			int32_t cid[4];
			int nc = 0;
			for (int side = 0; side < 2; side++) {
				int32_t c;
				memcpy(&c, &nb[2 * side].w, 4);
				if (c < 0) { clo[nc] = nb[2 * side]; chi[nc] = nb[2 * side + 1]; cid[nc++] = c; }
				else {
					const f4 *cb = T->nodes + 4 * (size_t) c;
					for (int s2 = 0; s2 < 2; s2++) { clo[nc] = cb[2 * s2]; chi[nc] = cb[2 * s2 + 1]; memcpy(&cid[nc], &cb[2 * s2].w, 4); nc++; }
				}
			}
			float lim = best.t < FLT_MAX ? best.t + T->t_slack : FLT_MAX;
			float tn[4];
			int hit_i[4], nh = 0;
			for (int i = 0; i < nc; i++) {
				__atomic_add_fetch(&g_wide_boxes, 1, __ATOMIC_RELAXED);
				if (overlap(clo[i], chi[i], o, inv, 0, lim, &tn[i])) hit_i[nh++] = i;
			}
			for (int i = 1; i < nh; i++)           /* near first */
				for (int j = i; j > 0 && tn[hit_i[j]] < tn[hit_i[j - 1]]; j--) { int t = hit_i[j]; hit_i[j] = hit_i[j - 1]; hit_i[j - 1] = t; }
			for (int i = nh - 1; i >= 1; i--) { stack_t[sp] = tn[hit_i[i]]; stack[sp++] = cid[hit_i[i]]; }
			if (sp > *deepest) *deepest = sp;
			if (nh) { node = cid[hit_i[0]]; continue; }
		} else {
			(*nodes)++;
			const f4 *nb = T->nodes + 4 * (size_t) node;
			f4 llo = nb[0], lhi = nb[1], rlo = nb[2], rhi = nb[3];
			float pad = 0, slack = T->t_slack;
			if (!T->global_pad) {
				float fx = far3(llo.x - o[0], lhi.x - o[0], rlo.x - o[0], rhi.x - o[0]);
				float fy = far3(llo.y - o[1], lhi.y - o[1], rlo.y - o[1], rhi.y - o[1]);
				float fz = far3(llo.z - o[2], lhi.z - o[2], rlo.z - o[2], rhi.z - o[2]);
				pad = dyn_pad(fx, fy, fz, lhi.w, T->emag, &slack);
			}
			float lim = best.t < FLT_MAX ? best.t + slack : FLT_MAX;
			float tl, tr;
			int hl = overlap(llo, lhi, o, inv, pad, lim, &tl);
			int hr = overlap(rlo, rhi, o, inv, pad, lim, &tr);
			int32_t cl, cr;
			memcpy(&cl, &llo.w, 4);
			memcpy(&cr, &rlo.w, 4);
			if (hl && hr) {
				int lf = tl <= tr;
				stack_t[sp] = lf ? tr : tl;
				stack[sp++] = lf ? cr : cl;
				if (sp > *deepest) *deepest = sp;
				node = lf ? cl : cr;
				continue;
			}
			if (hl) { node = cl; continue; }
			if (hr) { node = cr; continue; }
		}
		for (;;) {
			if (sp == 0) { node = 0x7fffffff; break; }
			node = stack[--sp];
			if (!pop_cull || !(stack_t[sp] > best.t + T->t_slack)) break;
		}
		if (node == 0x7fffffff) break;
	}
	*out = best;
}

int main(int argc, char **argv)
{
	if (argc < 5) { fprintf(stderr, "usage: lbvh_sim objects.bin W H generations [px py pz fx fy fz] [--dynamic-pad]\n"); return 2; }
	FILE *f = fopen(argv[1], "rb");
	if (!f) { perror(argv[1]); return 2; }
	fseek(f, 0, SEEK_END);
	long bytes = ftell(f);
	fseek(f, 0, SEEK_SET);
	int n = (int) (bytes / sizeof(RtoObject));
	RtoObject *obj = malloc(bytes);
	if (fread(obj, sizeof(RtoObject), n, f) != (size_t) n) return 2;
	fclose(f);
	int W = atoi(argv[2]), H = atoi(argv[3]), gens = atoi(argv[4]);
	RtoCamera cam = {{5, 5, 5}, {-1, -1, -1}, {0, 1, 0}, 30.0f};
	int global_pad = 1, argi = 5;
	if (argc >= 11 && argv[5][0] != '-') {
		for (int k = 0; k < 3; k++) { cam.pos[k] = (float) atof(argv[5 + k]); cam.front[k] = (float) atof(argv[8 + k]); }
		argi = 11;
	}
	for (; argi < argc; argi++) if (!strcmp(argv[argi], "--dynamic-pad")) global_pad = 0;

	g_fma = getenv("SIM_FMA") != NULL;
	g_wide = getenv("SIM_WIDE") != NULL;
	int axis_rays = getenv("SIM_AXIS") != NULL;    /* make some secondary rays (nearly) axis-parallel */
	Tree T;
	build(&T, obj, n, global_pad);
	g_packed = T.packed;

	/* light = first emissive object (main.c:140-146) */
	int light = -1;
	for (int i = 0; i < n && light < 0; i++) if (obj[i].emission_power > 0) light = i;
	float lp[3] = {0, 0, 0};
	if (light >= 0) {
		const float *g = obj[light].geom;
		if (obj[light].type == 1) { lp[0] = g[0]; lp[1] = g[1]; lp[2] = g[2]; }
		else { lp[0] = g[0] + g[3] * 0.5f; lp[1] = g[1] + g[4] * 0.5f; lp[2] = g[2] + g[5] * 0.5f; }
	}

	size_t nr = (size_t) W * H;
	float *rays = malloc(sizeof(float) * 6 * nr);
	if (getenv("SIM_RAYS")) {
		/* W*H rays (origin xyz, direction xyz, binary32) from a file instead of the camera */
		FILE *rf = fopen(getenv("SIM_RAYS"), "rb");
		if (!rf || fread(rays, sizeof(float) * 6, nr, rf) != nr) { fprintf(stderr, "bad SIM_RAYS file\n"); return 2; }
		fclose(rf);
	} else
	for (int j = 0; j < H; j++)
		for (int i = 0; i < W; i++)
			rto_camera_ray(&cam, 1.0f - (float) i / (W - 1), 1.0f - (float) j / (H - 1), (float) W / H, rays + 6 * ((size_t) j * W + i));

	long total_rays = 0, mism = 0, nodes = 0, tests = 0;
	unsigned char *is_shadow = NULL;
	int anyhit = getenv("SIM_ANYHIT") != NULL;      /* light samples in any-hit mode */
	int deepest = 0;
	uint64_t rng = 0x1234;
	for (int g = 0; g <= gens && nr > 0; g++) {
		float *out7 = malloc(sizeof(float) * 7 * nr);
		int32_t *hit = malloc(sizeof(int32_t) * nr);
		Best *bv = malloc(sizeof(Best) * nr);
		long gn = 0, gt = 0, gm = 0, sn = 0, ns = 0;
		int gd = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : gn, gt, gm, sn, ns) reduction(max : gd)
		for (size_t r = 0; r < nr; r++) {
			rto_trace_many(obj, n, rays + 6 * r, 1, out7 + 7 * r, hit + r);
			long a = 0, b = 0;
			int d = 0;
			int shadow_mode = anyhit && is_shadow && is_shadow[r];
			walk(&T, obj, rays + 6 * r, &bv[r], &a, &b, &d, shadow_mode ? light : -1);
			gn += a; gt += b;
			if (shadow_mode) { sn += a; ns++; }
			if (d > gd) gd = d;
			int ok = bv[r].obj == hit[r] && (hit[r] < 0 || bv[r].t == out7[7 * r]);
			if (shadow_mode)        /* all that matters: is the light the nearest hit? */
				ok = (bv[r].obj == light) == (hit[r] == light) && (hit[r] != light || bv[r].t == out7[7 * r]);
			if (!ok) {
				gm++;
				if (gm <= 3) fprintf(stderr, "MISMATCH gen %d ray %zu: scan obj %d t %a, lbvh obj %d t %a\n", g, r, hit[r], out7[7 * r], bv[r].obj, bv[r].t);
			}
		}
		total_rays += nr; nodes += gn; tests += gt; mism += gm;
		if (gd > deepest) deepest = gd;
		fprintf(stderr, "gen %d: %zu rays, %.1f nodes/ray, %.2f tests/ray, %ld mismatches", g, nr, (double) gn / nr, (double) gt / nr, gm);
		if (ns) fprintf(stderr, "; %ld shadow rays in any-hit mode: %.1f nodes/ray", ns, (double) sn / ns);
		fprintf(stderr, "\n");
		/* next generation: a light sample and a bounce from every hit */
		size_t cap = 2 * nr, m = 0;
		float *next = malloc(sizeof(float) * 6 * cap);
		unsigned char *next_shadow = calloc(cap, 1);
		for (size_t r = 0; r < nr; r++) {
			if (hit[r] < 0) continue;
			const float *h = out7 + 7 * r;
			float nrm[3] = {h[4], h[5], h[6]}, pt[3] = {h[1], h[2], h[3]};
			float rd[3];
			rto_random_direction(&rng, rd);
			if (light >= 0) {
				float sd[3], len = 0;
				for (int k = 0; k < 3; k++) { sd[k] = rd[k] * 0.5f + (lp[k] - pt[k]); len += sd[k] * sd[k]; }
				len = sqrtf(len);
				next_shadow[m] = 1;
				float *q = next + 6 * m++;
				for (int k = 0; k < 3; k++) { sd[k] /= len; q[k] = pt[k] + sd[k] * 0.001f; q[3 + k] = sd[k]; }
			}
			rto_random_direction(&rng, rd);
			float dn = rd[0] * nrm[0] + rd[1] * nrm[1] + rd[2] * nrm[2];
			float *q = next + 6 * m++;
			for (int k = 0; k < 3; k++) { float v = dn < 0 ? -rd[k] : rd[k]; q[k] = pt[k] + v * 0.001f; q[3 + k] = v; }
			if (axis_rays) {
				int ax = (int) (m % 3);
				switch (m & 7) {
				case 1: q[3 + ax] = 0.0f; break;
				case 2: q[3 + ax] *= 1e-7f; break;
				case 3: q[3 + ax] = -0.0f; q[3 + (ax + 1) % 3] *= 1e-12f; break;
				case 4: q[3 + ax] *= 1e-30f; break;
				default: break;
				}
			}
		}
		free(rays); free(out7); free(hit); free(bv); free(is_shadow);
		rays = next;
		is_shadow = next_shadow;
		nr = m;
	}
	if (g_wide) fprintf(stderr, "wide walk: %.2f boxes tested per ray\n", (double) g_wide_boxes / total_rays);
	printf("{\"objects\": %d, \"rays\": %ld, \"mismatches\": %ld, \"nodes_per_ray\": %.2f, \"tests_per_ray\": %.3f, \"deepest_stack\": %d, \"tree_depth\": %d, \"rule\": \"%s\"}\n",
	       n, total_rays, mism, (double) nodes / total_rays, (double) tests / total_rays, deepest, T.depth, global_pad ? "static pad (product)" : "per-node pad (experiment)");
	return mism != 0;
}
