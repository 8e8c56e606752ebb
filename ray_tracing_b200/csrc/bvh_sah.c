/*
 * bvh_sah.c -- host-side topology builder for the BVH of large scenes
 * (BASELINE.json config 5): top-down binned surface-area heuristic over the
 * PADDED primitive boxes (rt_lbvh_rule.h), as an alternative to the device's
 * Morton/Karras hierarchy (rt_lbvh.cu).
 *
 * The reference has no acceleration structure (scene.c:156-173 tests every
 * object), so nothing here can change a result: the walk (rt_device.cuh) applies
 * the reference's own per-primitive test and breaks ties by primitive index,
 * and the boxes come from the same refit (rt_lbvh.cu: leaf_box_kernel /
 * refit_kernel) whatever the topology.  What the topology decides is how many
 * nodes a ray visits: on config 5 (100 000 random spheres, boxes padded by
 * ~0.2) a Karras tree costs 44.4 internal nodes per ray, this one 39.5
 * (tests/lbvh_sim.c, SIM_TOPOLOGY=karras|sah, identical hits).
 *
 * Output conventions are the device build's (rt_lbvh.cu: hierarchy_kernel):
 *   prim_index[s]            primitive of leaf slot s (slots in depth-first order)
 *   children[2i], [2i+1]     >= 0 internal node, < 0 leaf (~slot)
 *   parent[i], parent[n-1+s] parent of internal node i / of leaf slot s; root: -1
 * Internal nodes are numbered in preorder (the root is node 0, a left child that
 * is internal is its parent's number + 1).  The depth is capped below
 * RT_SAH_MAX_DEPTH by falling back to median splits, so the walk's stack bound
 * (RT_BVH_STACK) holds for any input.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "rt_host.h"

#define RT_SAH_BINS      16
#ifndef RT_SAH_MAX_DEPTH
#define RT_SAH_MAX_DEPTH 60        /* < RT_BVH_STACK (rt_params.h); tests build this file with a lower cap */
#endif
#define RT_SAH_PAR_DEPTH 3        /* the first levels fork: up to 8 threads */
#define RT_SAH_PAR_MIN   16384    /* ... for nodes of at least this many primitives */

typedef struct { float lo[3], hi[3]; } SahBox;

/* one record per primitive, permuted in place as the build partitions (so every pass over a
 * node's primitives is a sequential read: with an index array and the boxes left where they were
 * the build of 100 000 spheres took 220 ms, most of it cache misses) */
typedef struct { SahBox box; float cen[3]; int prim; } SahPrim;

typedef struct {
	SahPrim *prims;
	int     *children, *parent;
	int      n;
} SahCtx;

static float sah_area(const SahBox *b)
{
	float dx = b->hi[0] - b->lo[0], dy = b->hi[1] - b->lo[1], dz = b->hi[2] - b->lo[2];
	return dx * dy + dy * dz + dz * dx;
}

static void sah_empty(SahBox *b)
{
	for (int k = 0; k < 3; k++) { b->lo[k] = FLT_MAX; b->hi[k] = -FLT_MAX; }
}

/* (plain comparisons: gcc does not inline fminf/fmaxf without -ffinite-math-only; a NaN
 * coordinate is simply never taken) */
static void sah_add(SahBox *b, const SahBox *o)
{
	for (int k = 0; k < 3; k++) {
		if (o->lo[k] < b->lo[k]) b->lo[k] = o->lo[k];
		if (o->hi[k] > b->hi[k]) b->hi[k] = o->hi[k];
	}
}

/* NaN and out-of-range centroids land in a valid bin */
static int sah_bin(float c, float lo, float scale)
{
	float f = (c - lo) * scale;
	return f >= 0.0f ? (f < (float) RT_SAH_BINS ? (int) f : RT_SAH_BINS - 1) : 0;
}

static int ilog2_ceil(int v)
{
	int l = 0;
	while ((1 << l) < v) l++;
	return l;
}

static int sah_split(SahCtx *C, int begin, int end, int depth)
{
	const int count = end - begin;
	/* a median split from here on still ends above the depth cap? then do not risk a lopsided one */
	if (depth + ilog2_ceil(count) + 1 >= RT_SAH_MAX_DEPTH) return (begin + end) / 2;
	float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
	for (int i = begin; i < end; i++) {
		const float *c = C->prims[i].cen;
		for (int k = 0; k < 3; k++) {
			if (c[k] < clo[k]) clo[k] = c[k];
			if (c[k] > chi[k]) chi[k] = c[k];
		}
	}
	/* one pass over the primitives fills the bins of all three axes */
	SahBox bb[3][RT_SAH_BINS];
	int bc[3][RT_SAH_BINS];
	float sc[3];
	int usable[3];
	for (int ax = 0; ax < 3; ax++) {
		float ext = chi[ax] - clo[ax];
		usable[ax] = ext > 0.0f && ext < FLT_MAX;
		sc[ax] = usable[ax] ? (float) RT_SAH_BINS / ext : 0.0f;
		for (int b = 0; b < RT_SAH_BINS; b++) { sah_empty(&bb[ax][b]); bc[ax][b] = 0; }
	}
	for (int i = begin; i < end; i++) {
		const float *c = C->prims[i].cen;
		const SahBox *pb = &C->prims[i].box;
		for (int ax = 0; ax < 3; ax++) {
			int b = sah_bin(c[ax], clo[ax], sc[ax]);
			sah_add(&bb[ax][b], pb);
			bc[ax][b]++;
		}
	}
	int best_axis = -1, best_bin = 0;
	float best_cost = FLT_MAX;
	for (int ax = 0; ax < 3; ax++) {
		if (!usable[ax]) continue;
		float ra[RT_SAH_BINS];
		int rc[RT_SAH_BINS];
		SahBox acc;
		int cnt = 0;
		sah_empty(&acc);
		for (int b = RT_SAH_BINS - 1; b >= 1; b--) {
			if (bc[ax][b]) sah_add(&acc, &bb[ax][b]);
			cnt += bc[ax][b];
			ra[b] = cnt ? sah_area(&acc) : 0.0f;
			rc[b] = cnt;
		}
		sah_empty(&acc);
		cnt = 0;
		for (int b = 0; b < RT_SAH_BINS - 1; b++) {
			if (bc[ax][b]) sah_add(&acc, &bb[ax][b]);
			cnt += bc[ax][b];
			if (cnt == 0 || rc[b + 1] == 0) continue;
			float cost = sah_area(&acc) * (float) cnt + ra[b + 1] * (float) rc[b + 1];
			if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; }
		}
	}
	if (best_axis < 0) return (begin + end) / 2;       /* coincident (or non-finite) centroids */
	int i = begin, j = end - 1;
	while (i <= j) {
		if (sah_bin(C->prims[i].cen[best_axis], clo[best_axis], sc[best_axis]) <= best_bin) i++;
		else { SahPrim t = C->prims[i]; C->prims[i] = C->prims[j]; C->prims[j] = t; j--; }
	}
	return (i == begin || i == end) ? (begin + end) / 2 : i;
}

/* Builds the subtree over prims[begin, end) whose root, if internal, is node `node`.  A subtree of c
 * leaves has c - 1 internal nodes, so in preorder the left subtree takes node + 1 .. node + cl - 1
 * and the right one starts at node + cl: the numbering needs no shared counter, and the two halves
 * of the first RT_SAH_PAR_DEPTH levels are built by separate threads. */
typedef struct { SahCtx *C; int begin, end, node, par, depth, result, deepest; } SahTask;

static int sah_build(SahCtx *C, int begin, int end, int node, int par, int depth, int *deepest);

static void *sah_thread(void *arg)
{
	SahTask *t = (SahTask *) arg;
	t->deepest = 0;
	t->result = sah_build(t->C, t->begin, t->end, t->node, t->par, t->depth, &t->deepest);
	return NULL;
}

static int sah_build(SahCtx *C, int begin, int end, int node, int par, int depth, int *deepest)
{
	if (end - begin == 1) {
		C->parent[(C->n - 1) + begin] = par;
		if (depth > *deepest) *deepest = depth;
		return ~begin;
	}
	C->parent[node] = par;
	int mid = sah_split(C, begin, end, depth);
	int l, r;
	pthread_t th;
	SahTask t = {C, mid, end, node + (mid - begin), node, depth + 1, 0, 0};
	if (depth < RT_SAH_PAR_DEPTH && end - begin >= RT_SAH_PAR_MIN && pthread_create(&th, NULL, sah_thread, &t) == 0) {
		l = sah_build(C, begin, mid, node + 1, node, depth + 1, deepest);
		pthread_join(th, NULL);
		r = t.result;
		if (t.deepest > *deepest) *deepest = t.deepest;
	} else {
		l = sah_build(C, begin, mid, node + 1, node, depth + 1, deepest);
		r = sah_build(C, mid, end, node + (mid - begin), node, depth + 1, deepest);
	}
	C->children[2 * node] = l;
	C->children[2 * node + 1] = r;
	return node;
}

/* fuzz_r2, cube_pad, extra: rt_lbvh_pads() for the d_max the device refit will use.
 * prim_index: n ints; children: 2(n-1); parent: 2n-1.  Returns 0, or -1 when out of memory. */
int rt_host_bvh_sah(const RtF4 *A, const RtF4 *B, int n, double fuzz_r2, float cube_pad, float extra,
                    int *prim_index, int *children, int *parent, int *depth_out)
{
	*depth_out = 0;
	if (n <= 0) return 0;
	if (n == 1) { prim_index[0] = 0; parent[0] = -1; return 0; }
	SahPrim *prims = (SahPrim *) malloc(sizeof(SahPrim) * (size_t) n);
	if (!prims) return -1;
	for (int i = 0; i < n; i++) {
		int ty;
		memcpy(&ty, &B[i].w, sizeof ty);
		const float a[3] = {A[i].x, A[i].y, A[i].z}, b[3] = {B[i].x, B[i].y, B[i].z};
		SahBox *box = &prims[i].box;
		if (ty == RT_OBJECT_SPHERE) {
			/* rt_lbvh.cu: leaf_box_kernel */
			float rp = (float) sqrt((double) fmaxf(A[i].w, 0.0f) + fuzz_r2) * 1.000001f + extra;
			for (int k = 0; k < 3; k++) { box->lo[k] = a[k] - rp; box->hi[k] = a[k] + rp; }
		} else if (ty == RT_OBJECT_CUBE) {
			float pad = cube_pad + extra;
			for (int k = 0; k < 3; k++) { box->lo[k] = fminf(a[k], b[k]) - pad; box->hi[k] = fmaxf(a[k], b[k]) + pad; }
		} else {
			/* never intersects (scene.c:138-153); its box is empty on the device: a point here */
			for (int k = 0; k < 3; k++) box->lo[k] = box->hi[k] = 0.0f;
		}
		for (int k = 0; k < 3; k++) prims[i].cen[k] = 0.5f * (box->lo[k] + box->hi[k]);
		prims[i].prim = i;
	}
	SahCtx C = {prims, children, parent, n};
	sah_build(&C, 0, n, 0, -1, 0, depth_out);
	for (int i = 0; i < n; i++) prim_index[i] = prims[i].prim;
	free(prims);
	return 0;
}
