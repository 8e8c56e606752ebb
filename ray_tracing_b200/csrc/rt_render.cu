/*
 * rt_render.cu -- the render kernels (sm_100a).  Compiled twice, see
 * rt_device.cuh: -DRT_NS=rt_exact -fmad=false and -DRT_NS=rt_fast -fmad=true -DRT_FAST_MATH.
 *
 * Replaces: worker()/render_column()/pixel() (src/main.c:131-414) -- the
 * pthread-per-column pool becomes one launch over low-res pixels.
 *
 * Three kernels, same device functions, bit-identical frames:
 *   render_pixel_kernel       one thread per low-res pixel, 8x4 pixel tile per
 *                             warp, each thread runs its whole path (baseline);
 *   render_persistent_kernel  (default) resident warps; a lane whose path ended
 *                             stores its pixel and immediately starts the next
 *                             pixel of the warp's batch, so the convergent
 *                             nearest-hit scan keeps all 32 lanes busy while
 *                             paths are 1..40 rays long (SURVEY.md 8(a)
 *                             divergence data);
 *   render_wavefront_kernel   per-warp pools of 64 paths in shared memory, every
 *                             phase over a compacted list of the paths needing it.
 * Scenes up to RT_SMEM_MAX_OBJECTS are scanned linearly from shared memory
 * (every lane reads the same primitive: broadcast, no bank conflicts); larger
 * ones walk the LBVH in global memory.
 */
#include <type_traits>

#include "rt_device.cuh"

namespace RT_NS {

/* ---------------------------------------------------------------- helpers */

/* fixed part of a block's scene area: byte LUT, sweep lists, direction caches */
/* The walk's loop wants ~12 more registers than the 64 a thread has at 8 CTAs per SM (ray 6, plane
 * selectors 6, cull limit, node base); without them the compiler re-derives the selectors and the
 * limit for every visited node (10 of 69 instructions, all on the ALU pipe, which is what bounds
 * the walk).  What a path only touches between rays -- contrib and result -- therefore waits in
 * shared memory while the lane walks (persistent kernel over the LBVH only). */
#ifndef RT_PARK_WORDS
#define RT_PARK_WORDS 6
#endif
#define RT_PARK_BYTES (RT_PARK_WORDS * sizeof(float) * RT_BLOCK_THREADS)
#ifndef RT_QUEUED_PARK
#define RT_QUEUED_PARK 0      /* the same for the queued kernel over the shared-memory scene (experiment) */
#endif
#ifndef RT_PARK_PATH
#define RT_PARK_PATH 1
#endif

#define RT_SCENE_HEAD_BYTES (256 * sizeof(float) + RT_BLOCK_THREADS + RT_PARK_BYTES + (RT_BLOCK_THREADS / 32) * RT_DIR_CACHE_BYTES)

struct SharedScene {
	float4 *A;
	float4 *B;
	float  *lut;
	unsigned char *sweep;     /* 32 bytes per warp: lane list of warp_sweep() */
	float  *dirs;             /* per warp: direction cache of warp_sweep_cached() (RT_DIR_ROW floats per lane) */
	int2   *runs;             /* type runs of the scene (rt_device.cuh: nearest_linear) */
	int    *stack;            /* LBVH: this thread's traversal-stack column (rt_device.cuh: SharedStack) */
	volatile float *park;     /* this thread's column of RT_PARK_WORDS words (path_park / path_unpark) */
};

/* `smem` = start of the scene area: kernels that keep per-warp queues in shared
 * memory put them FIRST, so that every address below is the block's base plus a
 * constant (only `runs` depends on the scene size) and costs no registers. */
__device__ __forceinline__ SharedScene stage_scene(const RtRenderParams &P, unsigned char *smem, bool linear)
{
	SharedScene s;
	s.lut = reinterpret_cast<float *>(smem);
	s.sweep = smem + 256 * sizeof(float) + 32 * (threadIdx.x >> 5);
	s.park = reinterpret_cast<float *>(smem + 256 * sizeof(float) + RT_BLOCK_THREADS) + threadIdx.x;
	s.dirs = reinterpret_cast<float *>(smem + RT_SCENE_HEAD_BYTES - (RT_BLOCK_THREADS / 32) * RT_DIR_CACHE_BYTES) + (threadIdx.x >> 5) * (32 * RT_DIR_ROW);
	s.A = reinterpret_cast<float4 *>(smem + RT_SCENE_HEAD_BYTES);
	s.B = s.A + 1;            /* records interleaved: A[2*i], B[2*i] are neighbours (one address per object) */
	s.stack = reinterpret_cast<int *>(s.A) + threadIdx.x;   /* LBVH scenes stage no objects: the area holds the stacks */
	s.runs = reinterpret_cast<int2 *>(s.A + (linear ? 2 * P.scene.n : 0));
	for (int i = threadIdx.x; i < 256; i += blockDim.x) s.lut[i] = __ldg(&P.byte_lut[i]);
	if (linear)
		for (int i = threadIdx.x; i < P.scene.n; i += blockDim.x) {
			s.A[2 * i] = __ldg(&P.scene.geomA[i]);
			s.B[2 * i] = __ldg(&P.scene.geomB[i]);
		}
	if (linear)
		for (int i = threadIdx.x; i < P.scene.num_runs; i += blockDim.x) s.runs[i] = P.scene.runs[i];
	__syncthreads();
	return s;
}

/* n / d for n*d < 2^40 with magic = ceil(2^40 / d) computed on the host */
__device__ __forceinline__ unsigned div_magic(unsigned n, unsigned long long magic)
{
	return (unsigned) (((unsigned long long) n * magic) >> 40);
}

/* tile-ordered work index -> low-res cell (cx, cy) inside the band */
__device__ __forceinline__ bool cell_of(const RtRenderParams &P, unsigned idx, int &cx, int &cy)
{
	unsigned tile = idx >> 5, lane = idx & 31;
	unsigned tyu = div_magic(tile, P.magic_tiles_x);
	int tx = (int) (tile - tyu * (unsigned) P.tiles_x), ty = (int) tyu;
	cx = tx * RT_TILE_W + (int) (lane & (RT_TILE_W - 1));
	int ly = ty * RT_TILE_H + (int) (lane / RT_TILE_W);     /* row among the rows this launch owns */
	/* owned rows -> band rows: blocks of 1 << il_shift rows dealt round robin */
	int blk = ly >> P.il_shift;
	cy = ((blk * P.il_n + P.il_i) << P.il_shift) + (ly - (blk << P.il_shift));
	return cx < P.cells_per_row && ly < P.local_rows && cy < (P.lrow1 - P.lrow0);
}

struct Cell {
	int   x0, y0;     /* first output pixel of the tile */
	int   tw;         /* tile width after column clipping (main.c:302-303) */
	float u, v;
};

/* main.c:280-303 */
__device__ __forceinline__ Cell cell_geometry(const RtRenderParams &P, int cx, int cy)
{
	Cell c;
	int col = 0, i = cx, column_x = 0, lcx = 0;
	if (P.num_columns > 1) {            /* warp-uniform */
		col = (int) div_magic((unsigned) cx, P.magic_cells_per_col);
		i = cx - col * P.cells_per_col;
		column_x = P.column_w * col;
		lcx = column_x / P.scale;
	}
	int j = P.lrow0 + cy;
	float u = (float) (lcx + i) / (float) (P.lw - 1);
	float v = (float) j / (float) (P.lh - 1);
	c.u = 1.0f - u;
	c.v = 1.0f - v;
	c.x0 = column_x + i * P.scale;
	c.y0 = j * P.scale;
	c.tw = min(P.scale, P.column_w - i * P.scale);
	if (P.compact) { c.x0 = cx; c.y0 = j; c.tw = 1; }       /* warp-uniform */
	return c;
}

/* main.c:305-310 (tile replication) fused with main.c:394 (accumulate) and
 * main.c:476 (resolve) when an accumulation buffer is attached. */
__device__ __forceinline__ void store_cell(const RtRenderParams &P, const Cell &c, f3 color)
{
	for (int g = 0; g < P.store_scale; g++) {
		int y = c.y0 + g;
		for (int t = 0; t < c.tw; t++) {
			int x = c.x0 + t;
			f3 out = color;
			if (P.accum) {
				float *a = P.accum + 3 * ((size_t) (y - P.accum_row_offset) * P.store_stride + x);
				f3 acc = mk(a[0] + color.x * P.accum_weight, a[1] + color.y * P.accum_weight,
				            a[2] + color.z * P.accum_weight);
				a[0] = acc.x; a[1] = acc.y; a[2] = acc.z;
				out = scl3(acc, P.inv_count);
			}
			size_t p = (size_t) (y - P.fb_row_offset) * P.store_stride + x;
			if (P.fb_format == RT_FB_F32X3) {
				float *f = reinterpret_cast<float *>(P.fb) + 3 * p;
				f[0] = out.x; f[1] = out.y; f[2] = out.z;
			} else {
				/* main.c:666-670: (uint8_t)(x*255) */
				uchar4 q;
				q.x = (unsigned char) __float2uint_rz(out.x * 255.0f);
				q.y = (unsigned char) __float2uint_rz(out.y * 255.0f);
				q.z = (unsigned char) __float2uint_rz(out.z * 255.0f);
				q.w = 255;
				reinterpret_cast<uchar4 *>(P.fb)[p] = q;
			}
		}
	}
}

/* One output pixel of a tile (shared by the per-lane and the warp-wide store). */
__device__ __forceinline__ void store_pixel(const RtRenderParams &P, int x, int y, f3 color)
{
	f3 out = color;
	if (P.accum) {
		float *a = P.accum + 3 * ((size_t) (y - P.accum_row_offset) * P.store_stride + x);
		f3 acc = mk(a[0] + color.x * P.accum_weight, a[1] + color.y * P.accum_weight, a[2] + color.z * P.accum_weight);
		a[0] = acc.x; a[1] = acc.y; a[2] = acc.z;
		out = scl3(acc, P.inv_count);
	}
	size_t p = (size_t) (y - P.fb_row_offset) * P.store_stride + x;
	if (P.fb_format == RT_FB_F32X3) {
		float *f = reinterpret_cast<float *>(P.fb) + 3 * p;
		f[0] = out.x; f[1] = out.y; f[2] = out.z;
	} else {
		uchar4 q;
		q.x = (unsigned char) __float2uint_rz(out.x * 255.0f);
		q.y = (unsigned char) __float2uint_rz(out.y * 255.0f);
		q.z = (unsigned char) __float2uint_rz(out.z * 255.0f);
		q.w = 255;
		reinterpret_cast<uchar4 *>(P.fb)[p] = q;
	}
}

/* Warp-wide tile replication for low-res passes (scale >= 4): the finished lanes'
 * tiles are written one after the other by all 32 lanes, consecutive lanes on
 * consecutive pixels of a tile row.  A single lane replicating a 16x16 tile does
 * 256 dependent read-modify-writes (a scale-16 accumulate pass took 0.76 ms
 * against 0.21 ms for the tracing itself).  All lanes must call. */
__device__ __forceinline__ void store_cells_warp(const RtRenderParams &P, bool has, const Cell &c, f3 color)
{
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	unsigned todo = __ballot_sync(full, has);
	while (todo) {
		int src = __ffs(todo) - 1;
		todo &= todo - 1;
		int x0 = __shfl_sync(full, c.x0, src), y0 = __shfl_sync(full, c.y0, src), tw = __shfl_sync(full, c.tw, src);
		f3 col = mk(__shfl_sync(full, color.x, src), __shfl_sync(full, color.y, src), __shfl_sync(full, color.z, src));
		int n = tw * P.store_scale;
		for (int i = lane; i < n; i += 32) {
			int g = i / tw, t = i - g * tw;
			store_pixel(P, x0 + t, y0 + g, col);
		}
	}
}

__device__ __forceinline__ void count_rays(const RtRenderParams &P, unsigned rays)
{
	for (int o = 16; o > 0; o >>= 1) rays += __shfl_xor_sync(0xffffffffu, rays, o);
	if ((threadIdx.x & 31) == 0 && rays) atomicAdd(P.ray_counter, (unsigned long long) rays);
}

/* counter build: ray_counter[1], [2] = internal nodes visited, primitives tested */
__device__ __forceinline__ void walk_counters_init(Walk &w)
{
#ifdef RT_COUNT_WALK
	w.nodes = w.tests = 0;
#else
	(void) w;
#endif
}

__device__ __forceinline__ void walk_counters_flush(const RtRenderParams &P, const Walk &w)
{
#ifdef RT_COUNT_WALK
	unsigned a = w.nodes, b = w.tests;
	for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(P.ray_counter + 1, (unsigned long long) a);
		atomicAdd(P.ray_counter + 2, (unsigned long long) b);
	}
#else
	(void) P; (void) w;
#endif
}

__device__ __forceinline__ void stack_init(SharedStack &st, const SharedScene &S) { st.init(S.stack); }

template <bool PARK>
__device__ __forceinline__ void path_park(const Path &p, const SharedScene &S)
{
	if (!PARK) return;
	volatile float *c = S.park;
	c[0 * RT_BLOCK_THREADS] = p.contrib.x; c[1 * RT_BLOCK_THREADS] = p.contrib.y; c[2 * RT_BLOCK_THREADS] = p.contrib.z;
	c[3 * RT_BLOCK_THREADS] = p.result.x;  c[4 * RT_BLOCK_THREADS] = p.result.y;  c[5 * RT_BLOCK_THREADS] = p.result.z;
}

template <bool PARK>
__device__ __forceinline__ void path_unpark(Path &p, const SharedScene &S)
{
	if (!PARK) return;
	volatile float *c = S.park;
	p.contrib = mk(c[0 * RT_BLOCK_THREADS], c[1 * RT_BLOCK_THREADS], c[2 * RT_BLOCK_THREADS]);
	p.result = mk(c[3 * RT_BLOCK_THREADS], c[4 * RT_BLOCK_THREADS], c[5 * RT_BLOCK_THREADS]);
}
__device__ __forceinline__ void stack_init(LocalStack &, const SharedScene &) {}

/*
 * One warp step: every lane with a pending ray traces it and consumes the hit
 * (classify), the warp shares out the light-sample tests of the new surfaces
 * (sweep), then every lane that still has a surface to work on builds its next
 * ray (launch).  Must be called by all 32 lanes.  Returns 1 for lanes that
 * started a ray.
 *
 * LBVH scenes: a ray's walk (rt_device.cuh: Walk) is spread over as many warp
 * steps as it needs.  One step = up to RT_WALK_ITERS internal nodes per lane,
 * then -- the warp reconverged in between -- the parked leaves, then the sphere
 * roots.  Lanes whose walk ended wait (MODE_HIT) until RT_WALK_HOLD of them do, or
 * nobody walks any more, so that classify / sweep / launch run with a fuller warp.
 */
#ifndef RT_WALK_ITERS
#define RT_WALK_ITERS 9     /* 4K config 5, same box, Karras tree: 8 / 9 / 10 / 11 / 12 / 14 nodes: 30.63 / 30.12 / 29.76 / 29.93 / 30.22 / 31.31 ms;
                             * SAH tree (fewer nodes per ray): 9 / 10 / 11 / 12: 27.78 / 28.02 / 28.49 / 29.10 ms (profiles/r02_lbvh_ab_iters.jsonl) */
#endif
#ifndef RT_WALK_HOLD
#define RT_WALK_HOLD 12     /* 4K config 5: 8 / 12 / 16 / 20 lanes: +3 % / 31.4 / 31.7 / 32.7 ms; 1 (no waiting): +17 % */
#endif

template <bool LBVH, bool DEFER_SKY, bool PARK = false, class Stack>
__device__ __forceinline__ unsigned warp_step(Path &p, Walk &w, Stack &st, const RtRenderParams &P, const SharedScene &S)
{
	unsigned traced = 0;
	if (!LBVH) {
		if (p.mode == MODE_TRACE) {
			f3 ro = p.ray_o;
			f3 dn = unit3(p.ray_d);             /* scene.c:158 */
			RayQ q = ray_quadratic(dn);
			Hit h = nearest_linear(S.A, S.B, S.runs, P.scene.num_runs, P.scene.n, ro, dn, q, P.scene.div_safe);
			traced = 1;
			path_unpark<PARK>(p, S);
			path_classify<DEFER_SKY>(p, h, dn, P.scene, P.sky, S.lut,
			              [&](const Hit &hh, f3 d, f3 &point, f3 &normal) {
				              surface_of(hh, S.A[2 * hh.obj], S.B[2 * hh.obj], ro, d, point, normal);
			              });
		}
	} else {
		const unsigned full = 0xffffffffu;
		/* Light samples in any-hit mode.  A sample adds what its nearest hit EMITS (main.c:199-204),
		 * so when one primitive is the scene's only emitter (bvh.emitter_prim) the walk only has
		 * to decide whether that primitive is the nearest hit: its leaf is tested first (parked
		 * here), its distance bounds the walk, and the first primitive accepted in front of it
		 * (or tying with a lower index) ends the walk -- as does missing the emitter. */
		const bool anyhit = p.shadow && P.bvh.emitter_slot >= 0;
		if (p.mode == MODE_TRACE) {
			p.ray_d = unit3(p.ray_d);           /* scene.c:158; kept for the steps the walk lasts */
			walk_begin(w, P.bvh, st);
			if (anyhit) w.leaf = ~P.bvh.emitter_slot;
			p.mode = MODE_WALK;
			traced = 1;
		}
		const bool walking = p.mode == MODE_WALK;
		const f3 ro = p.ray_o, dn = p.ray_d;
		/* (the ray's slab constants are re-derived every step: parking them in shared memory per
		 * ray was measured, 31.0 vs 30.5 ms -- the arithmetic is off the ALU pipe that bounds the walk) */
		if (walking) walk_nodes(P.bvh, walk_ray(P.bvh, ro, dn), w, st, RT_WALK_ITERS);
		__syncwarp();
		int prim = 0;
		float nb = 0.0f, discr = 0.0f;
		bool roots = false;
		const bool leafed = walking && w.leaf;
		if (leafed) roots = walk_leaf_screen(P.bvh, ro, dn, w, prim, nb, discr);
		__syncwarp();
		if (roots) walk_leaf_root(dn, prim, nb, discr, w);
		__syncwarp();
		if (leafed && anyhit && w.best.obj != P.bvh.emitter_prim) { w.node = RT_WALK_DONE; w.sp = st.bottom(); }
		if (walking && walk_over(w)) p.mode = MODE_HIT;
		const unsigned hits = __ballot_sync(full, p.mode == MODE_HIT);
		if (hits == 0) return traced;
		if (__popc(hits) < RT_WALK_HOLD && __any_sync(full, p.mode == MODE_WALK)) return traced;
		path_unpark<PARK>(p, S);
		if (p.mode == MODE_HIT)
			path_classify<DEFER_SKY>(p, w.best, dn, P.scene, P.sky, S.lut,
			              [&](const Hit &hh, f3 d, f3 &point, f3 &normal) {
				              surface_of(hh, __ldg(&P.scene.geomA[hh.obj]), __ldg(&P.scene.geomB[hh.obj]), ro, d, point, normal);
			              });
	}
	warp_sweep_cached(p, S.sweep, S.dirs);
	if (p.mode == MODE_LAUNCH) path_launch(p, P.scene, S.dirs + (threadIdx.x & 31) * RT_DIR_ROW);
	path_park<PARK>(p, S);
	return traced;
}

/* ---------------------------------------------------------------- kernels */

template <bool LBVH>
__global__ void __launch_bounds__(RT_BLOCK_THREADS)
render_pixel_kernel(const __grid_constant__ RtRenderParams P)
{
	extern __shared__ __align__(16) unsigned char smem[];
	SharedScene S = stage_scene(P, smem, !LBVH);

	const unsigned full = 0xffffffffu;
	unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned rays = 0;
	int cx, cy;
	Path p;
	p.mode = MODE_IDLE;
	Walk w;
	SharedStack st;
	if (LBVH) stack_init(st, S);        /* writes the stack sentinel into the area linear-scan scenes stage their objects in */
	Cell c;
	bool owns = idx < (unsigned) (P.tiles_x * P.tiles_y) * 32u && cell_of(P, idx, cx, cy);
	if (owns) {
		c = cell_geometry(P, cx, cy);
		path_begin(p, P.cam, c.u, c.v, P.pass_mix);
	}
	while (__any_sync(full, p.mode != MODE_IDLE))
		rays += warp_step<LBVH, false>(p, w, st, P, S);
	if (owns) store_cell(P, c, path_final(p));
	count_rays(P, rays);
}

#define RT_WARP_BATCH 8     /* most tiles (of 32 pixels) a warp claims per global atomic */

#ifndef RT_PERSISTENT_MIN_BLOCKS
#define RT_PERSISTENT_MIN_BLOCKS 6   /* 80 registers, 24 warps/SM: best of 5/6/8 on 4K scene_0 (2.39 / 2.37 / 2.51 ms) */
#endif

/* LBVH kernels.  Once the node fetch was down to one 32-byte load the walk became latency bound
 * and resident warps started to pay: BASELINE config 5 at 4K, persistent kernel, 5 / 6 / 7 / 8 / 9 /
 * 10 CTAs per SM (96 / 80 / 72 / 64 / 56 / 48 registers): 47.3 / 42.3 / 39.8 / 39.3 / 40.1 / 40.7 ms.
 * The queued kernel's extra shared memory (finish / refill queues) caps it at 5-6 CTAs: 44.5 ms. */
/* Re-measured once the stacks were sized by the tree's depth (rt_params.h) and the tree came from the SAH
 * builder: 7 CTAs x 72 registers (no spills; 155 KiB of shared memory: carve-out 164 KB, 92 KB of L1 for the
 * node fetches) against 8 x 64 (62 bytes of spills, 177 KiB: carve-out 196 KB, 60 KB of L1): 26.64 against
 * 27.55 ms (profiles/r02_lbvh_ab_iters.jsonl, tag b7). */
#ifndef RT_LBVH_MIN_BLOCKS
#define RT_LBVH_MIN_BLOCKS 7
#endif
#ifndef RT_LBVH_QUEUED_MIN_BLOCKS
#define RT_LBVH_QUEUED_MIN_BLOCKS 5
#endif

/* TRAV: 0 = linear scan from shared memory, 1 = LBVH with the traversal stacks in
 * shared memory, 2 = LBVH with local-memory stacks (trees deeper than RT_SMEM_STACK) */
template <int TRAV>
__global__ void __launch_bounds__(RT_BLOCK_THREADS, TRAV ? RT_LBVH_MIN_BLOCKS : RT_PERSISTENT_MIN_BLOCKS)
render_persistent_kernel(const __grid_constant__ RtRenderParams P)
{
	constexpr bool LBVH = TRAV != 0;
	constexpr bool PARK = TRAV == 1 && RT_PARK_PATH;
	extern __shared__ __align__(16) unsigned char smem[];
	SharedScene S = stage_scene(P, smem, !LBVH);

	const unsigned full = 0xffffffffu;
	const unsigned lane = threadIdx.x & 31;
	const unsigned total = (unsigned) (P.tiles_x * P.tiles_y) * 32u;

	Path p;
	p.mode = MODE_IDLE;
	Walk w;
	typename std::conditional<TRAV == 2, LocalStack, SharedStack>::type st;
	if (LBVH) stack_init(st, S);        /* writes the stack sentinel into the area linear-scan scenes stage their objects in */
	walk_counters_init(w);
	Cell c;
	bool owns = false;          /* lane holds a pixel whose path is running or just ended */
	unsigned rays = 0;
	unsigned batch_next = 0, batch_end = 0;   /* warp-uniform: tile-ordered pixel indices */
	unsigned batch_delta = 0;                 /* warp-uniform: (tile handed out << 5) - batch base, when P.tile_order deals the tiles */
	bool exhausted = false;                   /* warp-uniform */

	for (;;) {
		unsigned idle = __ballot_sync(full, p.mode == MODE_IDLE);
		if (idle) {
			path_unpark<PARK>(p, S);
			const bool ended = p.mode == MODE_IDLE && owns;
			/* longest tiles first (rt_api.cu: tile_schedule): what this pixel's path cost, for the pose's next pass */
			if (ended && P.tile_cost && p.bounce) atomicMax(P.tile_cost + ((unsigned) c.tw >> 12), (unsigned) p.bounce);
			Cell out = c;
			out.tw = c.tw & 4095;                  /* the tile index rides in the upper bits */
			if (P.store_scale >= 4) {       /* warp-uniform */
				store_cells_warp(P, ended, out, path_final(p));
				if (p.mode == MODE_IDLE) owns = false;
			} else if (ended) {
				store_cell(P, out, path_final(p));
				owns = false;
			}
			if (batch_next == batch_end && !exhausted) {
				/* guided self-scheduling: big batches while plenty of work is left,
				 * single tiles at the end.  A path is up to 40 rays long and a warp
				 * step takes microseconds, so a warp that claims 8 tiles of a
				 * mirror-heavy region late in the launch would otherwise BE the
				 * launch's tail (measured: 0.75 ms floor on a 3.2 ms frame).  With a
				 * longest-first order every claim is ONE tile (see render_queued_kernel). */
				unsigned base = 0, claim = 32u;
				if (lane == 0) {
					if (!P.tile_order) {
						unsigned seen = *(volatile unsigned *) P.work_counter;
						unsigned left = seen < total ? (total - seen) >> 5 : 0;
						unsigned warps = gridDim.x * (RT_BLOCK_THREADS / 32);
						claim = min(max(left / (4u * warps), 1u), (unsigned) RT_WARP_BATCH) * 32u;
					}
					base = atomicAdd(P.work_counter, claim);
				}
				base = __shfl_sync(full, base, 0);
				claim = __shfl_sync(full, claim, 0);
				if (base >= total) exhausted = true;
				else {
					batch_next = base;
					batch_end = min(base + claim, total);
					batch_delta = P.tile_order ? (__ldg(P.tile_order + (base >> 5)) << 5) - base : 0u;
				}
			}
			/* hand the idle lanes the next pixels of the warp's batch */
			unsigned avail = batch_end - batch_next;
			unsigned rank = __popc(idle & ((1u << lane) - 1u));
			int cx, cy;
			const unsigned idx = batch_next + rank + batch_delta;
			if (p.mode == MODE_IDLE && rank < avail && cell_of(P, idx, cx, cy)) {
				c = cell_geometry(P, cx, cy);
				c.tw |= (int) ((idx >> 5) << 12);
				path_begin(p, P.cam, c.u, c.v, P.pass_mix);
				owns = true;
			}
			batch_next += min((unsigned) __popc(idle), avail);
			path_park<PARK>(p, S);
		}
		if (__ballot_sync(full, p.mode != MODE_IDLE) == 0) {
			if (exhausted && batch_next == batch_end) break;
			continue;       /* only clipped cells were handed out; fetch more */
		}
		rays += warp_step<LBVH, false, PARK>(p, w, st, P, S);
	}
	count_rays(P, rays);
	walk_counters_flush(P, w);
}

/* ------------------------------------------- persistent kernel with queues */

/*
 * render_queued_kernel: the persistent kernel, except that the two things a
 * lane does BETWEEN paths no longer run with whichever few lanes happen to be
 * between paths in that warp step (ncu on the plain persistent kernel: sky lookup
 * 7.5, pixel store 8.5, cell geometry / camera ray / RNG key 8.2 of 32 lanes per
 * instruction, together 14 % of the issue slots):
 *
 *   finish   a lane whose path ended pushes (sky direction, contrib, result,
 *            pixel) onto its warp's stack in shared memory and is free at once;
 *            whenever 32 entries wait, the whole warp does the sky lookups
 *            (gpu_and_windowing.c:42-112), the clamp (main.c:267-269) and the
 *            stores (main.c:305-310, 394, 476) together;
 *   refill   the warp prepares a whole tile of 32 pixels at a time (cell
 *            geometry main.c:293-303, camera ray camera.c:121, RNG key) and parks
 *            them in shared memory; lanes pop one as their paths end.
 *
 * Paths stay in registers (unlike the wavefront kernel), the per-path arithmetic
 * is the same device code, every pixel is still finished exactly once: frames
 * and ray counts are identical to the other kernels'.
 */
#define RQ_FIN_CAP    64        /* at most 31 waiting + 32 pushed in one step */
#define RQ_FIN_WORDS  12
#define RQ_PREP_WORDS 8
#define RQ_WARP_WORDS (RQ_FIN_WORDS * RQ_FIN_CAP + RQ_PREP_WORDS * 32)
#define RQ_BLOCK_BYTES (sizeof(float) * RQ_WARP_WORDS * (RT_BLOCK_THREADS / 32))

enum { RQ_DIR = 0, RQ_CONTRIB = 3, RQ_RESULT = 6, RQ_X0 = 9, RQ_Y0 = 10, RQ_FLAGS = 11 };
enum { RQP_D = 0, RQP_RNGLO = 3, RQP_RNGHI = 4, RQP_X0 = 5, RQP_Y0 = 6, RQP_TW = 7 };
/* RQ_FLAGS / RQP_TW word: output tile width (bits 0-6), RQ_ESCAPED when the sky
 * lookup is still due, bounces of the path (bits 8-11, for the tile cost), index
 * of the 8x4 tile the pixel came from (bits 12-31; rt_api.cu keeps the tile
 * schedule off for frames with 2^20 tiles or more) */
#define RQ_ESCAPED     128u
#define RQ_TW(f)       ((f) & 127u)
#define RQ_BOUNCES(f)  (((f) >> 8) & 15u)
#define RQ_TILE(f)     ((f) >> 12)

/* finish `n` (<= 32) entries from the top of the warp's stack; all lanes call */
__device__ __forceinline__ void rq_drain(const RtRenderParams &P, const SharedScene &S, const float *fin, int top, int n)
{
	const int lane = threadIdx.x & 31;
	const bool has = lane < n;
	const int e = top - n + lane;
	Cell c;
	c.x0 = c.y0 = c.tw = 0;
	c.u = c.v = 0.0f;
	f3 color = mk(0.0f, 0.0f, 0.0f);
	if (has) {
		const unsigned *fu = reinterpret_cast<const unsigned *>(fin);
		unsigned flags = fu[RQ_FLAGS * RQ_FIN_CAP + e];
		f3 res = mk(fin[(RQ_RESULT + 0) * RQ_FIN_CAP + e], fin[(RQ_RESULT + 1) * RQ_FIN_CAP + e], fin[(RQ_RESULT + 2) * RQ_FIN_CAP + e]);
		if (flags & RQ_ESCAPED) {                      /* main.c:162-173 */
			f3 dn = mk(fin[(RQ_DIR + 0) * RQ_FIN_CAP + e], fin[(RQ_DIR + 1) * RQ_FIN_CAP + e], fin[(RQ_DIR + 2) * RQ_FIN_CAP + e]);
			f3 contrib = mk(fin[(RQ_CONTRIB + 0) * RQ_FIN_CAP + e], fin[(RQ_CONTRIB + 1) * RQ_FIN_CAP + e], fin[(RQ_CONTRIB + 2) * RQ_FIN_CAP + e]);
			f3 skyc = sky_lookup(P.sky, S.lut, dn);
			res = add3(res, mul3(skyc, contrib));
		}
		color = mk(clamp01(res.x), clamp01(res.y), clamp01(res.z));      /* main.c:267-269 */
		c.x0 = (int) fu[RQ_X0 * RQ_FIN_CAP + e];
		c.y0 = (int) fu[RQ_Y0 * RQ_FIN_CAP + e];
		c.tw = (int) RQ_TW(flags);
		/* a tile lasts as long as its longest path: the next pass hands out the
		 * long tiles first (sky tiles, cost 0, never touch the counter) */
		if (P.tile_cost && RQ_BOUNCES(flags)) atomicMax(P.tile_cost + RQ_TILE(flags), RQ_BOUNCES(flags));
	}
	if (P.store_scale >= 4) store_cells_warp(P, has, c, color);   /* warp-uniform */
	else if (has) store_cell(P, c, color);
}

#ifndef RT_QUEUED_BATCH
#define RT_QUEUED_BATCH RT_WARP_BATCH    /* most tiles per claim while tiles are in image order */
#endif
#ifndef RT_QUEUED_MIN_BLOCKS
#define RT_QUEUED_MIN_BLOCKS 6
#endif
#ifndef RT_QUEUED_DENSE_BLOCKS
#define RT_QUEUED_DENSE_BLOCKS 7
#endif

/* DENSE: the build for big launches of linear-scan scenes: 7 CTAs per SM at 72 registers, contrib /
 * result parked in shared memory while a lane traces (path_park: without it 72 registers spill in
 * the scan).  4K scene_0 1.887 -> 1.842 ms alone, 1.877 -> 1.831 ms per frame with three frames in
 * flight; small launches (a 720p frame, 1/8 of a 4K frame) are 4 % slower alone -- more warps, fewer
 * tiles each, a longer tail -- and the same in flight, so rt_api.cu picks this build from ~1080p up. */
template <bool LBVH, bool DENSE = false>
__global__ void __launch_bounds__(RT_BLOCK_THREADS, LBVH ? RT_LBVH_QUEUED_MIN_BLOCKS : (DENSE ? RT_QUEUED_DENSE_BLOCKS : RT_QUEUED_MIN_BLOCKS))
render_queued_kernel(const __grid_constant__ RtRenderParams P)
{
	extern __shared__ __align__(16) unsigned char smem[];
	SharedScene S = stage_scene(P, smem + RQ_BLOCK_BYTES, !LBVH);
	float *fin = reinterpret_cast<float *>(smem) + (size_t) (threadIdx.x >> 5) * RQ_WARP_WORDS;
	float *prep = fin + RQ_FIN_WORDS * RQ_FIN_CAP;
	unsigned *finu = reinterpret_cast<unsigned *>(fin), *prepu = reinterpret_cast<unsigned *>(prep);

	const unsigned full = 0xffffffffu;
	const unsigned lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
	const unsigned total = (unsigned) (P.tiles_x * P.tiles_y) * 32u;
	constexpr bool QPARK = !LBVH && (DENSE || RT_QUEUED_PARK);

	Path p;
	p.mode = MODE_IDLE;
	p.obj = 0;
	Walk w;
	SharedStack st;
	if (LBVH) stack_init(st, S);        /* writes the stack sentinel into the area linear-scan scenes stage their objects in */
	walk_counters_init(w);
	int cx0 = 0, cy0 = 0, ctw = 0;  /* output tile of the lane's pixel */
	bool owns = false;              /* lane holds a pixel whose path is running or just ended */
	unsigned rays = 0;
	int fin_n = 0, prep_n = 0;                /* warp-uniform stack heights */
	unsigned batch_next = 0, batch_end = 0;   /* warp-uniform: tile-ordered pixel indices, multiples of 32 */
	bool exhausted = false;                   /* warp-uniform */

	for (;;) {
		unsigned idle = __ballot_sync(full, p.mode == MODE_IDLE);
		if (idle) {
			path_unpark<QPARK>(p, S);
			/* ---- finish: push the ended paths, drain when a full warp's worth waits ---- */
			const bool ended = p.mode == MODE_IDLE && owns;
			unsigned em = __ballot_sync(full, ended);
			if (em) {
				if (ended) {
					int e = fin_n + (int) __popc(em & lt);
					const bool escaped = p.obj < 0;
					if (escaped) {
						fin[(RQ_DIR + 0) * RQ_FIN_CAP + e] = p.point.x;
						fin[(RQ_DIR + 1) * RQ_FIN_CAP + e] = p.point.y;
						fin[(RQ_DIR + 2) * RQ_FIN_CAP + e] = p.point.z;
						fin[(RQ_CONTRIB + 0) * RQ_FIN_CAP + e] = p.contrib.x;
						fin[(RQ_CONTRIB + 1) * RQ_FIN_CAP + e] = p.contrib.y;
						fin[(RQ_CONTRIB + 2) * RQ_FIN_CAP + e] = p.contrib.z;
					}
					fin[(RQ_RESULT + 0) * RQ_FIN_CAP + e] = p.result.x;
					fin[(RQ_RESULT + 1) * RQ_FIN_CAP + e] = p.result.y;
					fin[(RQ_RESULT + 2) * RQ_FIN_CAP + e] = p.result.z;
					finu[RQ_X0 * RQ_FIN_CAP + e] = (unsigned) cx0;
					finu[RQ_Y0 * RQ_FIN_CAP + e] = (unsigned) cy0;
					finu[RQ_FLAGS * RQ_FIN_CAP + e] = (unsigned) ctw | (escaped ? RQ_ESCAPED : 0u) | ((unsigned) p.bounce << 8);
					owns = false;
				}
				fin_n += (int) __popc(em);
				__syncwarp();
			}
			/* ---- refill: idle lanes pop prepared pixels; a tile is prepared when none is left ---- */
			unsigned want = idle;
			while (want) {
				if (prep_n == 0) {
					if (batch_next == batch_end) {
						if (exhausted) break;
						/* guided self-scheduling: big batches while plenty of work is left,
						 * single tiles at the end (see render_persistent_kernel).  With a
						 * longest-first order every claim is ONE tile: a batch of the first
						 * tiles would be several of the longest tiles in a row for one warp
						 * (scene_1 at 1080p: 0.48 ms with batches, 0.34 ms in image order). */
						unsigned base = 0, claim = 32u;
						if (lane == 0) {
							if (!P.tile_order) {
								unsigned seen = *(volatile unsigned *) P.work_counter;
								unsigned left = seen < total ? (total - seen) >> 5 : 0;
								unsigned warps = gridDim.x * (RT_BLOCK_THREADS / 32);
								claim = min(max(left / (4u * warps), 1u), (unsigned) RT_QUEUED_BATCH) * 32u;
							}
							base = atomicAdd(P.work_counter, claim);
						}
						base = __shfl_sync(full, base, 0);
						claim = __shfl_sync(full, claim, 0);
						if (base >= total) { exhausted = true; break; }
						batch_next = base;
						batch_end = min(base + claim, total);
					}
					unsigned tile = batch_next >> 5;
					if (P.tile_order) tile = __ldg(P.tile_order + tile);    /* longest tiles first */
					int cx, cy;
					const bool ok = cell_of(P, tile * 32u + lane, cx, cy);
					batch_next += 32;
					unsigned om = __ballot_sync(full, ok);
					if (ok) {
						Cell c = cell_geometry(P, cx, cy);
						f3 d = camera_dir(P.cam, c.u, c.v);
						uint64_t key = pixel_key(c.u, c.v, P.pass_mix);
						int e = (int) __popc(om & lt);
						prep[(RQP_D + 0) * 32 + e] = d.x;
						prep[(RQP_D + 1) * 32 + e] = d.y;
						prep[(RQP_D + 2) * 32 + e] = d.z;
						prepu[RQP_RNGLO * 32 + e] = (unsigned) key;
						prepu[RQP_RNGHI * 32 + e] = (unsigned) (key >> 32);
						prepu[RQP_X0 * 32 + e] = (unsigned) c.x0;
						prepu[RQP_Y0 * 32 + e] = (unsigned) c.y0;
						prepu[RQP_TW * 32 + e] = (unsigned) c.tw | (tile << 12);
					}
					prep_n = (int) __popc(om);
					__syncwarp();
					if (prep_n == 0) continue;          /* a tile of clipped cells only */
				}
				const bool mine = (want >> lane) & 1u;
				const int rank = (int) __popc(want & lt);
				const int n = min((int) __popc(want), prep_n);
				if (mine && rank < n) {
					int e = prep_n - 1 - rank;
					p.d = mk(prep[(RQP_D + 0) * 32 + e], prep[(RQP_D + 1) * 32 + e], prep[(RQP_D + 2) * 32 + e]);
					p.rng = ((uint64_t) prepu[RQP_RNGHI * 32 + e] << 32) | prepu[RQP_RNGLO * 32 + e];
					cx0 = (int) prepu[RQP_X0 * 32 + e];
					cy0 = (int) prepu[RQP_Y0 * 32 + e];
					ctw = (int) prepu[RQP_TW * 32 + e];
					/* path_begin (main.c:135-156) with the prepared ray and key */
					p.ray_o = mk(P.cam.origin);
					p.ray_d = p.d;
					p.contrib = mk(1.0f, 1.0f, 1.0f);
					p.result = mk(0.0f, 0.0f, 0.0f);
					p.bounce = 0;
					p.shadow = false;
					p.obj = 0;
					p.mode = MODE_TRACE;
					owns = true;
				}
				want = __ballot_sync(full, mine && rank >= n);
				prep_n -= n;
				__syncwarp();
			}
			path_park<QPARK>(p, S);
		}
		/* nothing runs and nothing is left to hand out: the launch is over for this warp */
		const bool over = __ballot_sync(full, p.mode != MODE_IDLE) == 0;
		while (fin_n >= 32 || (over && fin_n > 0)) {    /* the ONE finishing site */
			int n = min(fin_n, 32);
			rq_drain(P, S, fin, fin_n, n);
			fin_n -= n;
			__syncwarp();
		}
		if (over) break;
		rays += warp_step<LBVH, true, QPARK>(p, w, st, P, S);
	}
	count_rays(P, rays);
	walk_counters_flush(P, w);
}

/* ------------------------------------------------------- wavefront kernel */

/*
 * render_wavefront_kernel: the same per-path state machine, but every WARP
 * keeps a pool of 64 paths in SHARED MEMORY (structure of arrays, 32 words per
 * path) and runs each phase of a round over a compacted list of exactly the
 * paths that need it, 32 at a time:
 *
 *   idle     escaped paths finish (sky lookup), pixels are stored, new pixels taken
 *   trace    paths with a pending ray: nearest-hit scan + classify
 *   sweep    three rd.n > 0 tests per fresh surface (one task per lane)
 *   launch   next shadow ray / shade and bounce, as separate homogeneous lists
 *
 * In the persistent kernel a lane owns one path, so whenever some lanes of a
 * warp need a phase the whole warp walks through it (ncu: 15 of 32 lanes active
 * per instruction; the non-trace phases ran at 25-40 % of lanes).  Here a warp
 * executes a phase only with lanes that need it.  Pools are per warp, so the
 * phases need no CTA barrier (a first version with one pool per CTA executed
 * 30 % fewer instructions but lost it all to barrier stalls).  Per-path
 * arithmetic and the order of every path's own draws are unchanged, so the
 * results are identical to the other kernels'.
 */
#define WF_WARPS   4            /* warps per CTA */
#define WF_THREADS (32 * WF_WARPS)
#ifndef WF_SLOTS
#define WF_SLOTS   2            /* pool slots per lane: lane, lane + 32, ... */
#endif
#define WF_PATHS   (32 * WF_SLOTS)   /* paths per warp */
#define WF_WORDS   32

enum { WF_RO = 0, WF_RD = 3, WF_D = 6, WF_CONTRIB = 9, WF_RESULT = 12, WF_POINT = 15, WF_NORMAL = 18,
       WF_TOLIGHT = 21, WF_SAMPLED = 24, WF_RNGLO = 27, WF_RNGHI = 28, WF_OBJ = 29, WF_STATE = 30, WF_PIXEL = 31 };

/* WF_STATE bits */
#define WF_MODE(st)      ((st) & 3u)
#define WF_SHADOW        (1u << 2)
#define WF_OWNS          (1u << 3)
#define WF_BOUNCE(st)    (((st) >> 4) & 15u)
#define WF_GOT(st)       (((st) >> 8) & 3u)
#define WF_PENDING(st)   (((st) >> 10) & 7u)
#define WF_FRESH         (1u << 13)
#define WF_SKY           (1u << 14)     /* idle slot whose path escaped: sky lookup pending (direction in WF_POINT) */
#define WF_TW(st)        (((st) >> 16) & 31u)

struct WfPool {
	float *w;                   /* this warp's pool: WF_WORDS x WF_PATHS words */
	__device__ __forceinline__ float &f(int field, int slot) const { return w[field * WF_PATHS + slot]; }
	__device__ __forceinline__ unsigned &u(int field, int slot) const { return reinterpret_cast<unsigned *>(w)[field * WF_PATHS + slot]; }
	__device__ __forceinline__ f3 get3(int field, int slot) const { return mk(f(field, slot), f(field + 1, slot), f(field + 2, slot)); }
	__device__ __forceinline__ void set3(int field, int slot, f3 v) const { f(field, slot) = v.x; f(field + 1, slot) = v.y; f(field + 2, slot) = v.z; }
};

/* Append this lane's slots (lane, lane + 32, ...) whose predicate holds to a
 * warp-local list; returns the list length.  Warp-convergent. */
__device__ __forceinline__ int wf_list(const bool (&p)[WF_SLOTS], unsigned char *list)
{
	const unsigned full = 0xffffffffu;
	const unsigned lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
	int n = 0;
#pragma unroll
	for (int k = 0; k < WF_SLOTS; k++) {
		unsigned m = __ballot_sync(full, p[k]);
		if (p[k]) list[n + __popc(m & lt)] = (unsigned char) (lane + 32 * k);
		n += __popc(m);
	}
	return n;
}

/* launch for one pool slot: next shadow ray, or shade and bounce */
__device__ __forceinline__ void wf_launch_slot(const WfPool &pool, int s, const RtSceneView &scene)
{
	unsigned st = pool.u(WF_STATE, s);
	Path p;
	p.mode = MODE_LAUNCH;
	p.rng = ((uint64_t) pool.u(WF_RNGHI, s) << 32) | pool.u(WF_RNGLO, s);
	p.pending = (int) WF_PENDING(st);
	p.got = (st & WF_FRESH) ? __popc(WF_PENDING(st)) : (int) WF_GOT(st);   /* main.c:206 */
	p.bounce = (int) WF_BOUNCE(st);
	p.shadow = (st & WF_SHADOW) != 0;
	p.normal = pool.get3(WF_NORMAL, s);
	p.point = pool.get3(WF_POINT, s);
	const bool shade = p.pending == 0;
	if (shade) {
		p.obj = (int) pool.u(WF_OBJ, s);
		p.sampled = pool.get3(WF_SAMPLED, s);
		p.d = pool.get3(WF_D, s);
		p.contrib = pool.get3(WF_CONTRIB, s);
		p.result = pool.get3(WF_RESULT, s);
	}
	uint64_t rng_before = p.rng;
	path_launch(p, scene);
	pool.set3(WF_RO, s, p.ray_o);
	pool.set3(WF_RD, s, p.ray_d);
	if (shade) {
		pool.set3(WF_D, s, p.d);
		pool.set3(WF_CONTRIB, s, p.contrib);
		pool.set3(WF_RESULT, s, p.result);
		if (p.rng != rng_before) {
			pool.u(WF_RNGLO, s) = (unsigned) p.rng;
			pool.u(WF_RNGHI, s) = (unsigned) (p.rng >> 32);
		}
	}
	pool.u(WF_STATE, s) = (st & (WF_OWNS | (31u << 16))) | (unsigned) p.mode | (p.shadow ? WF_SHADOW : 0u) |
	                      ((unsigned) p.bounce << 4) | ((unsigned) p.got << 8) | ((unsigned) p.pending << 10);
}

#ifndef RT_WAVEFRONT_MIN_BLOCKS
#define RT_WAVEFRONT_MIN_BLOCKS 5
#endif

template <bool LBVH>
__global__ void __launch_bounds__(WF_THREADS, RT_WAVEFRONT_MIN_BLOCKS)
render_wavefront_kernel(const __grid_constant__ RtRenderParams P)
{
	extern __shared__ __align__(16) unsigned char smem[];
	SharedScene S = stage_scene(P, smem, !LBVH);
	size_t scene_bytes = RT_SCENE_HEAD_BYTES +
	                     (LBVH ? sizeof(int) * RT_SMEM_STACK_ENTRIES(P.bvh.depth) * RT_BLOCK_THREADS
	                           : 2 * sizeof(float4) * (size_t) P.scene.n + sizeof(int2) * (size_t) P.scene.num_runs);
	scene_bytes = (scene_bytes + 15) & ~(size_t) 15;
	const unsigned full = 0xffffffffu;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	WfPool pool;
	pool.w = reinterpret_cast<float *>(smem + scene_bytes) + (size_t) warp * WF_WORDS * WF_PATHS;
	unsigned char *lists = smem + scene_bytes + sizeof(float) * WF_WORDS * WF_PATHS * WF_WARPS + (size_t) warp * 3 * 2 * WF_PATHS;
	unsigned char *list0 = lists, *list1 = lists + 2 * WF_PATHS, *list2 = lists + 4 * WF_PATHS;   /* list0 may hold trace + refilled */

	const unsigned total = (unsigned) (P.tiles_x * P.tiles_y) * 32u;
	const bool lit = P.scene.light_index >= 0;
	unsigned rays = 0;
	bool exhausted = false;          /* warp-uniform */

#pragma unroll
	for (int k = 0; k < WF_SLOTS; k++) pool.u(WF_STATE, lane + 32 * k) = 0u;
	__syncwarp();

	for (;;) {
		/* ---- lists of this round: slots with a pending ray (0) and idle slots (1) ---- */
		unsigned st[WF_SLOTS];
		bool pa[WF_SLOTS], pb[WF_SLOTS];
		bool mine_launching = false;
#pragma unroll
		for (int k = 0; k < WF_SLOTS; k++) {
			st[k] = pool.u(WF_STATE, lane + 32 * k);
			pa[k] = WF_MODE(st[k]) == MODE_TRACE;
			pb[k] = WF_MODE(st[k]) == MODE_IDLE;
			mine_launching = mine_launching || WF_MODE(st[k]) == MODE_LAUNCH;
		}
		int n_trace = wf_list(pa, list0);
		int n_idle = wf_list(pb, list1);
		bool launching = __any_sync(full, mine_launching);
		unsigned base = 0;
		if (!exhausted && n_idle > 0) {
			if (lane == 0) base = atomicAdd(P.work_counter, (unsigned) n_idle);
			base = __shfl_sync(full, base, 0);
			if (base >= total) exhausted = true;
		}
		__syncwarp();

		/* ---- idle slots: finish the escaped path (sky), hand in the pixel, take the next one ---- */
		int dealt = 0;
		for (int c0 = 0; c0 < n_idle; c0 += 32) {
			int i = c0 + lane;
			bool ok = false;
			int s = 0;
			if (i < n_idle) {
				s = list1[i];
				unsigned st = pool.u(WF_STATE, s);
				if (st & WF_OWNS) {
					f3 res = pool.get3(WF_RESULT, s);
					if (st & WF_SKY) {                      /* main.c:170-171 */
						f3 skyc = sky_lookup(P.sky, S.lut, pool.get3(WF_POINT, s));
						res = add3(res, mul3(skyc, pool.get3(WF_CONTRIB, s)));
					}
					Cell c;
					unsigned px = pool.u(WF_PIXEL, s);
					c.x0 = (int) (px & 0xffffu); c.y0 = (int) (px >> 16); c.tw = (int) WF_TW(st);
					store_cell(P, c, mk(clamp01(res.x), clamp01(res.y), clamp01(res.z)));
				}
				unsigned idx = base + (unsigned) i;
				int cx, cy;
				ok = !exhausted && idx < total && cell_of(P, idx, cx, cy);
				if (ok) {
					Cell c = cell_geometry(P, cx, cy);
					Path p;
					path_begin(p, P.cam, c.u, c.v, P.pass_mix);
					pool.set3(WF_RO, s, p.ray_o);
					pool.set3(WF_RD, s, p.ray_d);
					pool.set3(WF_D, s, p.d);
					pool.set3(WF_CONTRIB, s, p.contrib);
					pool.set3(WF_RESULT, s, p.result);
					pool.u(WF_RNGLO, s) = (unsigned) p.rng;
					pool.u(WF_RNGHI, s) = (unsigned) (p.rng >> 32);
					pool.u(WF_PIXEL, s) = (unsigned) c.x0 | ((unsigned) c.y0 << 16);
					pool.u(WF_STATE, s) = (unsigned) MODE_TRACE | WF_OWNS | ((unsigned) c.tw << 16);
				} else
					pool.u(WF_STATE, s) = 0u;
			}
			/* refilled slots join this round's trace list */
			unsigned m = __ballot_sync(full, ok);
			if (ok) list0[n_trace + dealt + __popc(m & ((1u << lane) - 1u))] = (unsigned char) s;
			dealt += __popc(m);
		}
		n_trace += dealt;
		__syncwarp();
		if (n_trace == 0 && !launching) {
			if (exhausted) break;
			continue;               /* only clipped cells were dealt: claim more */
		}

		/* ---------------- trace + classify ---------------- */
		for (int c0 = 0; c0 < n_trace; c0 += 32) {
			int i = c0 + lane;
			if (i < n_trace) {
				int s = list0[i];
				unsigned st = pool.u(WF_STATE, s);
				f3 ro = pool.get3(WF_RO, s);
				f3 dn = unit3(pool.get3(WF_RD, s));  /* scene.c:158 */
				RayQ q = ray_quadratic(dn);
				Hit h;
				if (LBVH) { LocalStack ls; h = nearest_lbvh(P.bvh, ro, dn, ls); }
				else      h = nearest_linear(S.A, S.B, S.runs, P.scene.num_runs, P.scene.n, ro, dn, q, P.scene.div_safe);
				rays++;
				if (st & WF_SHADOW) {
					if (h.obj >= 0) {                      /* main.c:201-204 */
						float4 m2 = __ldg(P.scene.mat + (size_t) h.obj * RT_MAT_STRIDE + 2);
						pool.set3(WF_SAMPLED, s, add3(pool.get3(WF_SAMPLED, s), mk(m2.x, m2.y, m2.z)));
					}
					pool.u(WF_STATE, s) = (st & ~3u) | (unsigned) MODE_LAUNCH;
				} else if (h.obj < 0) {
					/* escaped: the sky lookup runs with the idle list of the next round */
					pool.set3(WF_POINT, s, dn);
					pool.u(WF_STATE, s) = (st & ~3u) | (unsigned) MODE_IDLE | WF_SKY;
				} else {
					f3 point, normal;
					if (LBVH) surface_of(h, __ldg(&P.scene.geomA[h.obj]), __ldg(&P.scene.geomB[h.obj]), ro, dn, point, normal);
					else      surface_of(h, S.A[2 * h.obj], S.B[2 * h.obj], ro, dn, point, normal);
					pool.u(WF_OBJ, s) = (unsigned) h.obj;
					pool.set3(WF_POINT, s, point);
					pool.set3(WF_NORMAL, s, normal);
					pool.set3(WF_SAMPLED, s, mk(0.0f, 0.0f, 0.0f));
					/* got = 0, pending = 0; fresh asks the sweep phase for the three tests */
					pool.u(WF_STATE, s) = (st & ~(3u | (3u << 8) | (7u << 10) | WF_FRESH)) | (unsigned) MODE_LAUNCH | (lit ? WF_FRESH : 0u);
				}
			}
		}
		__syncwarp();

		/* ---- after the trace: fresh surfaces get their three tests first ... ---- */
#pragma unroll
		for (int k = 0; k < WF_SLOTS; k++) {
			st[k] = pool.u(WF_STATE, lane + 32 * k);
			pa[k] = (st[k] & WF_FRESH) != 0 && WF_MODE(st[k]) == MODE_LAUNCH;
		}
		int n_fresh = wf_list(pa, list0);
		__syncwarp();
		for (int t = lane; t < 3 * n_fresh; t += 32) {
			int j = (t * 171) >> 9;                         /* t / 3 for t < 256 */
			int k = t - 3 * j;
			int s = list0[j];
			uint64_t x0 = ((uint64_t) pool.u(WF_RNGHI, s) << 32) | pool.u(WF_RNGLO, s);
			if (sample_faces_surface(x0, k, pool.get3(WF_NORMAL, s), P.sweep_tau2))
				atomicOr(&pool.u(WF_STATE, s), 1u << (10 + k));
		}
		__syncwarp();

		/* ---- ... then one launch list: shadow rays first, shading from the next
		 * multiple of 32, so that every chunk of 32 is homogeneous ---- */
#pragma unroll
		for (int k = 0; k < WF_SLOTS; k++) {
			st[k] = pool.u(WF_STATE, lane + 32 * k);
			bool l = WF_MODE(st[k]) == MODE_LAUNCH;
			pa[k] = l && WF_PENDING(st[k]) != 0;
			pb[k] = l && WF_PENDING(st[k]) == 0;
		}
		int n_shadow = wf_list(pa, list1);
		int shade_at = (n_shadow + 31) & ~31;
		int n_shade = wf_list(pb, list1 + shade_at);
		__syncwarp();
		for (int i = lane; i < shade_at + n_shade; i += 32)
			if (i < n_shadow || i >= shade_at) wf_launch_slot(pool, list1[i], P.scene);
		__syncwarp();
	}
	count_rays(P, rays);
}

/* ------------------------------------------------------------ unit probes */

template <bool LBVH>
__global__ void probe_trace_kernel(RtRenderParams P, const float *rays6, int n, float *out7, int *obj)
{
	extern __shared__ __align__(16) unsigned char smem[];
	SharedScene S = stage_scene(P, smem, !LBVH);
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	f3 o = mk(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]);
	f3 d = unit3(mk(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]));
	RayQ q = ray_quadratic(d);
	Hit h;
	if (LBVH) { LocalStack ls; h = nearest_lbvh(P.bvh, o, d, ls); }
	else      h = nearest_linear(S.A, S.B, S.runs, P.scene.num_runs, P.scene.n, o, d, q, P.scene.div_safe);
	float *r = out7 + 7 * (size_t) i;
	obj[i] = h.obj;
	if (h.obj < 0) {                        /* scene.c:175-181 */
		r[0] = -1.0f;
		for (int k = 1; k < 7; k++) r[k] = 0.0f;
		return;
	}
	f3 pt, nm;
	surface_of(h, __ldg(&P.scene.geomA[h.obj]), __ldg(&P.scene.geomB[h.obj]), o, d, pt, nm);
	r[0] = h.t;
	r[1] = pt.x; r[2] = pt.y; r[3] = pt.z;
	r[4] = nm.x; r[5] = nm.y; r[6] = nm.z;
}

__global__ void probe_sky_kernel(RtSkyView sky, const float *lut, const float *dirs3, int n, float *out3)
{
	__shared__ float s_lut[256];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut[i];
	__syncthreads();
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	f3 c = sky_lookup(sky, s_lut, mk(dirs3[3 * i], dirs3[3 * i + 1], dirs3[3 * i + 2]));
	out3[3 * i] = c.x; out3[3 * i + 1] = c.y; out3[3 * i + 2] = c.z;
}

__global__ void probe_camera_kernel(RtCameraFrame cam, const float *pxpy, int n, float *rays6)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	f3 d = camera_dir(cam, pxpy[2 * i], pxpy[2 * i + 1]);
	float *r = rays6 + 6 * (size_t) i;
	r[0] = cam.origin.x; r[1] = cam.origin.y; r[2] = cam.origin.z;
	r[3] = d.x; r[4] = d.y; r[5] = d.z;
}

__global__ void probe_rng_kernel(uint64_t state, int n, uint64_t *u64_out, float *f32_out, float *dir_out)
{
	if (blockIdx.x || threadIdx.x) return;
	uint64_t s = state;
	if (u64_out) for (int i = 0; i < n; i++) u64_out[i] = wyhash64(s);
	s = state;
	if (f32_out) for (int i = 0; i < n; i++) f32_out[i] = random_float(s);
	s = state;
	if (dir_out) for (int i = 0; i < n; i++) {
		f3 d = random_direction(s);
		dir_out[3 * i] = d.x; dir_out[3 * i + 1] = d.y; dir_out[3 * i + 2] = d.z;
	}
}

/* div_hoisted(a, b, recip_refine(b)) against nvcc's IEEE a / b on pseudo-random
 * operand pairs inside the guarded range; counts bit mismatches. */
__global__ void probe_div_kernel(uint64_t seed, unsigned per_thread, int lo_exp_b, int hi_exp_b, int lo_exp_a,
                                 int hi_exp_a, unsigned long long *mismatches)
{
	uint64_t st = splitmix64(seed ^ ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * 0x9e3779b97f4a7c15ull);
	unsigned bad = 0;
	for (unsigned i = 0; i < per_thread; i++) {
		uint64_t r1 = wyhash64(st), r2 = wyhash64(st);
		unsigned eb = (unsigned) (lo_exp_b + 127) + (unsigned) ((r1 >> 40) % (unsigned) (hi_exp_b - lo_exp_b + 1));
		unsigned ea = (unsigned) (lo_exp_a + 127) + (unsigned) ((r2 >> 40) % (unsigned) (hi_exp_a - lo_exp_a + 1));
		unsigned mb = (unsigned) r1 & 0x7fffffu, ma = (unsigned) r2 & 0x7fffffu;
		/* bias some mantissas towards the hard cases: all ones / all zeros / few bits */
		unsigned sel = (unsigned) (r1 >> 60);
		if (sel == 0) mb = 0x7fffffu; else if (sel == 1) mb = 0; else if (sel == 2) mb &= 0x7f0000u;
		sel = (unsigned) (r2 >> 60);
		if (sel == 0) ma = 0x7fffffu; else if (sel == 1) ma = 0; else if (sel == 2) ma &= 0x7f0000u;
		float b = __uint_as_float(((unsigned) (r1 >> 63) << 31) | (eb << 23) | mb);
		float a = __uint_as_float(((unsigned) (r2 >> 63) << 31) | (ea << 23) | ma);
		if (i % 97 == 0) a = 0.0f;          /* +0 numerators are inside the guard, -0 is not */
		float want = a / b;
		float got = div_hoisted(a, b, recip_refine(b));
		bad += __float_as_uint(want) != __float_as_uint(got);
	}
	if (bad) atomicAdd(mismatches, (unsigned long long) bad);
}

} // namespace RT_NS

/* ------------------------------------------------------------ launchers */

#define RT_CAT2(a, b) a##_##b
#define RT_CAT(a, b) RT_CAT2(a, b)
#define RT_FN(name) RT_CAT(RT_NS, name)

static size_t smem_bytes(const RtRenderParams &P, bool lbvh)
{
	return RT_SCENE_HEAD_BYTES +
	       (lbvh ? sizeof(int) * RT_SMEM_STACK_ENTRIES(P.bvh.depth) * RT_BLOCK_THREADS      /* traversal stacks */
	             : 2 * sizeof(float4) * (size_t) P.scene.n + sizeof(int2) * (size_t) P.scene.num_runs);
}

template <class K>
static cudaError_t allow_smem(K kernel, size_t bytes)
{
	if (bytes <= 48 * 1024) return cudaSuccess;
	return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
}

static size_t wavefront_smem_bytes(const RtRenderParams &P, bool lbvh)
{
	size_t scene = (smem_bytes(P, lbvh) + 15) & ~(size_t) 15;
	return scene + (sizeof(float) * WF_WORDS * WF_PATHS + 3 * 2 * WF_PATHS) * WF_WARPS;
}

static size_t queued_smem_bytes(const RtRenderParams &P, bool lbvh)
{
	return RQ_BLOCK_BYTES + smem_bytes(P, lbvh);
}

extern "C" cudaError_t RT_FN(launch_render)(const RtRenderParams *P, int lbvh, int persistent,
                                             int grid_blocks, cudaStream_t stream)
{
	using namespace RT_NS;
	size_t sm = smem_bytes(*P, lbvh != 0);
	unsigned total = (unsigned) (P->tiles_x * P->tiles_y) * 32u;
	if (total == 0) return cudaSuccess;
	cudaError_t e;
	if (lbvh && P->bvh.depth > RT_SMEM_STACK) {
		/* a tree deeper than the shared-memory stacks (degenerate scenes): the
		 * local-memory-stack build of the persistent kernel, whatever was asked for */
		if ((e = allow_smem(render_persistent_kernel<2>, sm)) != cudaSuccess) return e;
		render_persistent_kernel<2><<<grid_blocks, RT_BLOCK_THREADS, sm, stream>>>(*P);
		return cudaGetLastError();
	}
	if (persistent == 3) {      /* persistent kernel with finish / refill queues */
		size_t qsm = queued_smem_bytes(*P, lbvh != 0);
		if (lbvh) {
			if ((e = allow_smem(render_queued_kernel<true>, qsm)) != cudaSuccess) return e;
			render_queued_kernel<true><<<grid_blocks, RT_BLOCK_THREADS, qsm, stream>>>(*P);
		} else {
			if ((e = allow_smem(render_queued_kernel<false>, qsm)) != cudaSuccess) return e;
			render_queued_kernel<false><<<grid_blocks, RT_BLOCK_THREADS, qsm, stream>>>(*P);
		}
		return cudaGetLastError();
	}
	if (persistent == 4) {      /* the queued kernel's build for big launches of linear-scan scenes */
		size_t qsm = queued_smem_bytes(*P, false);
		if ((e = allow_smem(render_queued_kernel<false, true>, qsm)) != cudaSuccess) return e;
		render_queued_kernel<false, true><<<grid_blocks, RT_BLOCK_THREADS, qsm, stream>>>(*P);
		return cudaGetLastError();
	}
	if (persistent == 2) {      /* wavefront kernel: one CTA per SM, paths pooled in shared memory */
		size_t wsm = wavefront_smem_bytes(*P, lbvh != 0);
		if (lbvh) {
			if ((e = allow_smem(render_wavefront_kernel<true>, wsm)) != cudaSuccess) return e;
			render_wavefront_kernel<true><<<grid_blocks, WF_THREADS, wsm, stream>>>(*P);
		} else {
			if ((e = allow_smem(render_wavefront_kernel<false>, wsm)) != cudaSuccess) return e;
			render_wavefront_kernel<false><<<grid_blocks, WF_THREADS, wsm, stream>>>(*P);
		}
		return cudaGetLastError();
	}
	if (persistent) {
		if (lbvh) {
			if ((e = allow_smem(render_persistent_kernel<1>, sm)) != cudaSuccess) return e;
			render_persistent_kernel<1><<<grid_blocks, RT_BLOCK_THREADS, sm, stream>>>(*P);
		} else {
			if ((e = allow_smem(render_persistent_kernel<0>, sm)) != cudaSuccess) return e;
			render_persistent_kernel<0><<<grid_blocks, RT_BLOCK_THREADS, sm, stream>>>(*P);
		}
	} else {
		unsigned blocks = (total + RT_BLOCK_THREADS - 1) / RT_BLOCK_THREADS;
		if (lbvh) {
			if ((e = allow_smem(render_pixel_kernel<true>, sm)) != cudaSuccess) return e;
			render_pixel_kernel<true><<<blocks, RT_BLOCK_THREADS, sm, stream>>>(*P);
		} else {
			if ((e = allow_smem(render_pixel_kernel<false>, sm)) != cudaSuccess) return e;
			render_pixel_kernel<false><<<blocks, RT_BLOCK_THREADS, sm, stream>>>(*P);
		}
	}
	return cudaGetLastError();
}

extern "C" cudaError_t RT_FN(wavefront_blocks_per_sm)(const RtRenderParams *P, int lbvh, int *out)
{
	using namespace RT_NS;
	size_t sm = wavefront_smem_bytes(*P, lbvh != 0);
	cudaError_t e;
	if (lbvh) {
		if ((e = allow_smem(render_wavefront_kernel<true>, sm)) != cudaSuccess) return e;
		return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, render_wavefront_kernel<true>, WF_THREADS, sm);
	}
	if ((e = allow_smem(render_wavefront_kernel<false>, sm)) != cudaSuccess) return e;
	return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, render_wavefront_kernel<false>, WF_THREADS, sm);
}

extern "C" int RT_FN(wavefront_paths_per_block)(void) { return WF_PATHS * WF_WARPS; }

/* occupancy query for sizing the persistent grid (queued != 0: render_queued_kernel) */
extern "C" cudaError_t RT_FN(persistent_blocks_per_sm)(const RtRenderParams *P, int lbvh, int queued, int *out)
{
	using namespace RT_NS;
	size_t sm = smem_bytes(*P, lbvh != 0);
	cudaError_t e;
	if (queued) {
		sm = queued_smem_bytes(*P, lbvh != 0);
		if (lbvh) {
			if ((e = allow_smem(render_queued_kernel<true>, sm)) != cudaSuccess) return e;
			return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, render_queued_kernel<true>, RT_BLOCK_THREADS, sm);
		}
		if (queued == 2) {
			if ((e = allow_smem(render_queued_kernel<false, true>, sm)) != cudaSuccess) return e;
			return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, render_queued_kernel<false, true>, RT_BLOCK_THREADS, sm);
		}
		if ((e = allow_smem(render_queued_kernel<false>, sm)) != cudaSuccess) return e;
		return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, render_queued_kernel<false>, RT_BLOCK_THREADS, sm);
	}
	if (lbvh) {
		if ((e = allow_smem(render_persistent_kernel<1>, sm)) != cudaSuccess) return e;
		return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, render_persistent_kernel<1>, RT_BLOCK_THREADS, sm);
	}
	if ((e = allow_smem(render_persistent_kernel<0>, sm)) != cudaSuccess) return e;
	return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, render_persistent_kernel<0>, RT_BLOCK_THREADS, sm);
}

extern "C" cudaError_t RT_FN(launch_probe_trace)(const RtRenderParams *P, int lbvh, const float *rays6, int n,
                                                  float *out7, int *obj, cudaStream_t stream)
{
	using namespace RT_NS;
	if (n <= 0) return cudaSuccess;
	size_t sm = smem_bytes(*P, lbvh != 0);
	int blocks = (n + 127) / 128;
	cudaError_t e;
	if (lbvh) {
		if ((e = allow_smem(probe_trace_kernel<true>, sm)) != cudaSuccess) return e;
		probe_trace_kernel<true><<<blocks, 128, sm, stream>>>(*P, rays6, n, out7, obj);
	} else {
		if ((e = allow_smem(probe_trace_kernel<false>, sm)) != cudaSuccess) return e;
		probe_trace_kernel<false><<<blocks, 128, sm, stream>>>(*P, rays6, n, out7, obj);
	}
	return cudaGetLastError();
}

extern "C" cudaError_t RT_FN(launch_probe_sky)(const RtSkyView *sky, const float *lut, const float *dirs3, int n,
                                                float *out3, cudaStream_t stream)
{
	if (n <= 0) return cudaSuccess;
	RT_NS::probe_sky_kernel<<<(n + 127) / 128, 128, 0, stream>>>(*sky, lut, dirs3, n, out3);
	return cudaGetLastError();
}

extern "C" cudaError_t RT_FN(launch_probe_camera)(const RtCameraFrame *cam, const float *pxpy, int n,
                                                   float *rays6, cudaStream_t stream)
{
	if (n <= 0) return cudaSuccess;
	RT_NS::probe_camera_kernel<<<(n + 127) / 128, 128, 0, stream>>>(*cam, pxpy, n, rays6);
	return cudaGetLastError();
}

extern "C" cudaError_t RT_FN(launch_probe_div)(uint64_t seed, unsigned blocks, unsigned per_thread, int lo_b, int hi_b,
                                                int lo_a, int hi_a, unsigned long long *mismatches, cudaStream_t stream)
{
	RT_NS::probe_div_kernel<<<blocks, 256, 0, stream>>>(seed, per_thread, lo_b, hi_b, lo_a, hi_a, mismatches);
	return cudaGetLastError();
}

extern "C" cudaError_t RT_FN(launch_probe_rng)(uint64_t state, int n, uint64_t *u64_out, float *f32_out,
                                                float *dir_out, cudaStream_t stream)
{
	RT_NS::probe_rng_kernel<<<1, 32, 0, stream>>>(state, n, u64_out, f32_out, dir_out);
	return cudaGetLastError();
}
