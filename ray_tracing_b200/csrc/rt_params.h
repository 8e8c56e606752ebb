/*
 * rt_params.h -- kernel parameter blocks (POD, passed by value as
 * __grid_constant__) shared by rt_api.cu and the two builds of rt_render.cu.
 */
#ifndef RT_PARAMS_H
#define RT_PARAMS_H

#include <cuda_runtime.h>
#include <stdint.h>
#include "rt_host.h"

#define RT_BVH_STACK 64
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 32         /* most traversal-stack entries per thread in shared memory; deeper trees use the local-memory build */
#endif
/* A launch reserves what ITS tree needs: one entry per level (near-first traversal) and the sentinel.
 * BASELINE config 5 (depth 20): 10.5 instead of 16.5 KiB of stacks per CTA next to the 10.6 KiB head, so
 * eight CTAs (plus 1 KiB each) need 177 instead of 225 KiB: the carve-out steps from 228 down to 196 KB
 * and the L1 that serves the node fetches grows from 28 to 60 KB: 28.02 -> 27.81 ms per 4K frame */
#define RT_SMEM_STACK_ENTRIES(depth) (((depth) < RT_SMEM_STACK ? (depth) : RT_SMEM_STACK) + 1)
#define RT_TILE_W 8           /* a warp covers an 8x4 tile of low-res pixels */
#define RT_TILE_H 4
#define RT_BLOCK_THREADS 128
#define RT_SMEM_MAX_OBJECTS 1024   /* linear-scan scenes are staged in shared memory */

struct RtSceneView {
	const float4 *geomA;      /* see rt_host.h for the record layout */
	const float4 *geomB;
	const float4 *mat;
	int           n;
	int           light_index;
	RtVector3     light_pos;
	const int2   *runs;       /* maximal runs of same-type objects: (first, count | type << 24) */
	int           num_runs;
	int           div_safe;   /* every box coordinate is zero or in [2^-37, 2^59] (rt_device.cuh: ray_div) */
};

struct RtSkyView {
	const uchar4 *texels;     /* 6 faces, CubeFace order, RGBA8, row-major top row first */
	int           w, h;
	size_t        face_stride;/* w*h */
};

/*
 * LBVH (Karras 2012) over primitive AABBs.  What the walk reads is the PACKED tree: 32 bytes per
 * internal node, fetched with one 256-bit load (the lanes of a warp sit on different nodes, so the
 * L1 spends one wavefront per lane and load instruction whatever the width: that pipe was the
 * limiter of the walk):
 *   words 0-2   left  child box  x, y, z: lo | hi << 16     16-bit fixed point in the tree's own frame
 *   words 3-5   right child box  (same)                      q = (x - center) * scale + 32768
 *   word  6, 7  left, right child
 * lo is rounded down and hi up, plus one quantum each (rt_lbvh.cu: pack_lo), so a packed box contains
 * the binary32 box it came from (which is itself padded, rt_lbvh_rule.h) with room for the rounding
 * of the walk's slab arithmetic; scale is a power of two chosen so that the bounds map to +-32000.
 * One PRMT per plane turns a half word into the binary32 number 2^23 + q AND picks the near or
 * the far plane of its axis by the ray's sign (rt_device.cuh: walk_nodes), so a box costs 6 PRMT,
 * 6 FFMA and 4 min/max.  (Until round 2 the boxes were binary16: 12 conversions + 12 FFMA + 20
 * min/max per node, and 8 times coarser at the rim of the scene.)
 * child >= 0: internal node index; child < 0: leaf, ~child = slot in the Morton-sorted order.
 * rt_lbvh.cu keeps the binary32 boxes (4 float4 per node) as the refit's working copy.
 */
struct RtBvhView {
	const uint4  *nodes;      /* 2 uint4 per internal node */
	const float4 *leaves;     /* 2 float4 per leaf slot (Morton order): geomA; geomB.xyz, prim index | type << 30 */
	int           num_prims;
	int           depth;      /* deepest leaf (levels below the root) = most stack entries a walk can hold */
	float         t_slack;    /* see rt_lbvh_rule.h: cull only if t_entry > best + slack */
	float         cx, cy, cz; /* the packed boxes' frame */
	float         scale, inv_scale;
	/* light samples in any-hit mode (rt_render.cu: warp_step): the scene's only emitting primitive
	 * and its leaf slot, or -1, -1 */
	int           emitter_prim, emitter_slot;
};

struct RtRenderParams {
	RtCameraFrame cam;
	RtSceneView   scene;
	RtBvhView     bvh;
	RtSkyView     sky;
	const float  *byte_lut;   /* 256 floats: (float)i/255 */

	/* render_column geometry (main.c:278-296) */
	int W, H, scale;
	int num_columns, column_w;
	int lw, lh;               /* W/scale, H/scale */
	int cells_per_col;        /* ceil(column_w / scale): visible low-res pixels per column row */
	int cells_per_row;        /* num_columns * cells_per_col */
	int lrow0, lrow1;         /* low-res row band [lrow0, lrow1) rendered by this launch */
	int il_n, il_i, il_shift; /* this launch owns row blocks b (of 1<<il_shift low-res rows) with b % il_n == il_i */
	int local_rows;           /* low-res rows of the band owned by this launch */
	int tiles_x, tiles_y;     /* 8x4 tiles covering cells_per_row x local_rows */
	uint64_t pass_mix;        /* splitmix64(pass_index) */
	unsigned long long magic_tiles_x, magic_cells_per_col;   /* ceil(2^40 / d): index math without integer division */
	float    sweep_tau2;      /* rt_device.cuh: sample_faces_surface */

	/* output */
	int    compact;           /* 1: one value per low-res CELL, row-major over cells_per_row x lh (the concurrent
	                           * sweep, rt_api.cu: the passes of a sweep run side by side and a resolve kernel folds
	                           * them into the frame in pass order); 0: the tile replication of main.c:305-310 */
	int    store_scale;       /* output pixels per cell side: scale, or 1 when compact */
	int    store_stride;      /* pixels per output row: W, or cells_per_row when compact */
	void  *fb;                /* RT_FB_F32X3: float[3] per pixel; RT_FB_U8X4: uchar4 */
	int    fb_format;
	int    fb_row_offset;     /* output row r is stored at fb row (r - fb_row_offset) */
	float *accum;             /* optional W*H*3 accumulation buffer (row offset applies too) */
	int    accum_row_offset;
	float  accum_weight;      /* 1.0f/(scale*scale)  (main.c:278,394) */
	float  inv_count;         /* 1.0f/accum_count after this pass (main.c:476) */

	unsigned long long *ray_counter;
	unsigned int       *work_counter;   /* persistent kernel: next tile-ordered pixel */

	/* render_queued_kernel: longest-tiles-first schedule (rt_api.cu: tile order).
	 * tile_order[k] = tile handed out k-th (NULL: natural order); tile_cost[t] =
	 * largest bounce count of a path of tile t, recorded when non-NULL. */
	const unsigned int *tile_order;
	unsigned int       *tile_cost;
};

#endif
