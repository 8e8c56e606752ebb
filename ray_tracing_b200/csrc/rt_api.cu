/*
 * rt_api.cu -- the C ABI of include/rt_cuda.h: device lifecycle, scene/skybox
 * upload (AoS -> SoA in HBM, LBVH for large scenes), the render entry points
 * that replace the reference's worker pool (src/main.c:324-414, 450-482) with
 * kernel launches, row-band multi-GPU rendering with a fused P2P composite, and
 * the unit probes used by the parity tests.
 *
 * There is deliberately no CPU path in this file: every rendering entry point
 * fails with RT_ERR_NO_DEVICE when no CUDA device is usable.
 */
#include <cuda_runtime.h>

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "rt_cuda.h"
#include "rt_host.h"
#include "rt_params.h"
#include "rt_lbvh.h"

/* launchers exported by the two builds of rt_render.cu */
#define DECLARE_VARIANT(ns)                                                                              \
	extern "C" cudaError_t ns##_launch_render(const RtRenderParams *, int, int, int, cudaStream_t);      \
	extern "C" cudaError_t ns##_persistent_blocks_per_sm(const RtRenderParams *, int, int, int *);            \
	extern "C" cudaError_t ns##_wavefront_blocks_per_sm(const RtRenderParams *, int, int *);             \
	extern "C" int ns##_wavefront_paths_per_block(void);                                                \
	extern "C" cudaError_t ns##_launch_probe_trace(const RtRenderParams *, int, const float *, int,      \
	                                               float *, int *, cudaStream_t);                        \
	extern "C" cudaError_t ns##_launch_probe_sky(const RtSkyView *, const float *, const float *, int,   \
	                                             float *, cudaStream_t);                                 \
	extern "C" cudaError_t ns##_launch_probe_camera(const RtCameraFrame *, const float *, int, float *,  \
	                                                cudaStream_t);                                       \
	extern "C" cudaError_t ns##_launch_probe_rng(uint64_t, int, uint64_t *, float *, float *, cudaStream_t);      \
	extern "C" cudaError_t ns##_launch_probe_div(uint64_t, unsigned, unsigned, int, int, int, int,             \
	                                             unsigned long long *, cudaStream_t);
DECLARE_VARIANT(rt_exact)
DECLARE_VARIANT(rt_fast)

/* ------------------------------------------------------------------ errors */

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

extern "C" const char *rt_cuda_last_error(void) { return g_err; }

#define CU(call)                                                                                     \
	do {                                                                                             \
		cudaError_t e_ = (call);                                                                     \
		if (e_ != cudaSuccess)                                                                       \
			return fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),        \
			            __FILE__, __LINE__);                                                         \
	} while (0)

/* ----------------------------------------------------------------- context */

#define RT_MAX_GPUS 16
#define RT_WORK_SLOTS 8       /* tile counters: launches that may be in flight at once on one GPU */
#define RT_SWEEP_MAX_COARSE 6 /* passes of a sweep above scale 1 (init_scale <= 64) */
#define RT_SYNC_BANDS_MAX 8   /* row bands of a synchronous call with a host frame (render_pass) */

/* What the tile costs of a pose are valid for: the same scene seen from the same
 * camera through the same frame, band and interleave -- at any scale. */
struct TileKey {
	int      w, h, ncols, r0, r1, il_n, il_i, lbvh;
	unsigned scene_epoch;
	RtCamera cam;
};

/* Longest-tiles-first schedule of the queued kernel (see tile_schedule()). */
struct TileSched {
	TileKey       key;
	bool          have_key = false;
	unsigned int *cost[2] = {nullptr, nullptr};     /* per-tile largest bounce count, double buffered */
	unsigned int *order[2] = {nullptr, nullptr};
	size_t        cost_cap[2] = {0, 0}, order_cap[2] = {0, 0};
	int           cost_cur = -1, cost_scale = 0, cost_tiles_x = 0, cost_tiles_y = 0;
	int           order_cur = -1, order_scale = 0, order_from_scale = 0;
	size_t        order_tiles = 0;
	cudaEvent_t   fence = nullptr;                  /* after the last launch that touched these buffers */
	unsigned int *hist = nullptr;                   /* scratch of launch_tile_order() */
	cudaStream_t  last_stream = nullptr;
	/* Launches of a pose on DIFFERENT streams may overlap (the tail of one frame under the head of the
	 * next: a 1/8 share of a 4K frame takes 0.27 ms alone and 0.19 ms per frame on three streams) as
	 * long as they only READ the order.  A launch that WRITES (records costs, builds an order) waits
	 * for the last use on every other stream; a reader waits for the last writer. */
	struct Use { cudaStream_t st = nullptr; cudaEvent_t ev = nullptr; } uses[4];
	int           nuses = 0;
	cudaEvent_t   write_fence = nullptr;
	cudaStream_t  write_stream = nullptr;
	bool          launch_writes = false;            /* the launch being issued */
};

struct DeviceCtx {
	int          device = -1;
	int          sm_count = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t  ev[4] = {nullptr, nullptr, nullptr, nullptr};
	/* scene */
	float4 *geomA = nullptr, *geomB = nullptr, *mat = nullptr;
	int2   *runs = nullptr;
	RtLbvh  bvh;
	/* skybox */
	uchar4 *sky = nullptr;
	float  *lut = nullptr;
	/* outputs */
	void   *fb = nullptr;       size_t fb_bytes = 0;      /* internal framebuffer (host destinations) */
	float  *accum = nullptr;    size_t accum_bytes = 0;
	bool    accum_needs_clear = false;    /* zero it on the stream of the next accumulating pass */
	int     accum_w = 0, accum_h = 0, accum_row0 = 0, accum_rows = 0;
	unsigned long long *ray_counter = nullptr;
	unsigned int       *work_counter = nullptr;   /* RT_WORK_SLOTS counters, one per launch in flight */
	unsigned            launch_seq = 0;
	unsigned long long *host_rays = nullptr;              /* pinned */
	TileSched    sched;
	/* pipelined host read-back: two staging frames + a copy stream */
	cudaStream_t copy_stream = nullptr;
	void        *stage[2] = {nullptr, nullptr};
	size_t       stage_bytes[2] = {0, 0};
	cudaEvent_t  stage_rendered[2] = {nullptr, nullptr}, stage_copied[2] = {nullptr, nullptr};
	int          stage_next = 0;
	cudaEvent_t  band_ev[RT_SYNC_BANDS_MAX] = {};     /* banded host read-back (render_pass): band k rendered */
	/* concurrent sweep (sweep_concurrent): side streams for the coarse passes, their cell buffers */
	cudaStream_t sweep_stream[RT_SWEEP_MAX_COARSE] = {};
	cudaEvent_t  sweep_fork = nullptr, sweep_join[RT_SWEEP_MAX_COARSE] = {};
	float       *sweep_cells = nullptr;   size_t sweep_cells_floats = 0;
	float       *sweep_fine = nullptr;    size_t sweep_fine_floats = 0;     /* the scale-1 pass, W x H float3 */
};

struct Context {
	bool      ready = false;
	int       ngpu = 0;
	DeviceCtx dev[RT_MAX_GPUS];
	/* scene meta (identical on every device) */
	int       n = 0, light_index = -1;
	RtVector3 light_pos = {0, 0, 0};
	bool      have_scene = false, have_bvh = false;
	int       div_safe = 0, num_runs = 0;
	int       sky_w = 0, sky_h = 0;
	bool      have_sky = false;
	RtScene  *scene_cache = nullptr;    /* copy of the last RtScene given to render_frame_cuda */
	float     accum_count = 0.0f;       /* accum_counts[] of main.c:89 (all columns advance together) */
	float     sweep_tau2 = 4e-12f;      /* rt_device.cuh: sample_faces_surface; tests may override */
	unsigned  scene_epoch = 0;          /* bumped by every scene upload (tile schedules die with the scene) */
	int       tile_schedule = 1;        /* 0: never reorder tiles (tests / A-B) */
	int       concurrent_sweep = 1;     /* 0: rt_cuda_render_sweep runs its passes one after the other (tests / A-B) */
	int       queued_dense = 1;         /* 0: never take the queued kernel's 7-CTA build (tests / A-B) */
	int       sync_bands = 4;           /* most row bands of a synchronous call with a host frame; 1 = render, then copy */
	bool      sync_bands_forced = false;/* tests: exactly that many, whatever the frame size */
};

static Context g;

static int select_device(const DeviceCtx &d)
{
	CU(cudaSetDevice(d.device));
	return RT_OK;
}

static void free_device(DeviceCtx &d)
{
	if (d.device < 0) return;
	cudaSetDevice(d.device);
	cudaFree(d.geomA); cudaFree(d.geomB); cudaFree(d.mat); cudaFree(d.runs);
	rt_lbvh_free(&d.bvh);
	cudaFree(d.sky); cudaFree(d.lut);
	cudaFree(d.fb); cudaFree(d.accum);
	cudaFree(d.ray_counter); cudaFree(d.work_counter);
	for (int k = 0; k < 2; k++) { cudaFree(d.sched.cost[k]); cudaFree(d.sched.order[k]); }
	cudaFree(d.sched.hist);
	if (d.sched.fence) cudaEventDestroy(d.sched.fence);
	if (d.sched.write_fence) cudaEventDestroy(d.sched.write_fence);
	for (auto &u : d.sched.uses) if (u.ev) cudaEventDestroy(u.ev);
	d.sched = TileSched();
	if (d.host_rays) cudaFreeHost(d.host_rays);
	for (auto &e : d.ev) if (e) cudaEventDestroy(e);
	for (int k = 0; k < 2; k++) {
		cudaFree(d.stage[k]);
		if (d.stage_rendered[k]) cudaEventDestroy(d.stage_rendered[k]);
		if (d.stage_copied[k]) cudaEventDestroy(d.stage_copied[k]);
	}
	if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
	for (int k = 0; k < RT_SWEEP_MAX_COARSE; k++) {
		if (d.sweep_stream[k]) cudaStreamDestroy(d.sweep_stream[k]);
		if (d.sweep_join[k]) cudaEventDestroy(d.sweep_join[k]);
	}
	if (d.sweep_fork) cudaEventDestroy(d.sweep_fork);
	for (auto &e : d.band_ev) if (e) cudaEventDestroy(e);
	cudaFree(d.sweep_cells); cudaFree(d.sweep_fine);
	if (d.stream) cudaStreamDestroy(d.stream);
	d = DeviceCtx();
}

static int init_devices(const int *devices, int count)
{
	int available = 0;
	cudaError_t e = cudaGetDeviceCount(&available);
	if (e != cudaSuccess || available <= 0)
		return fail(RT_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
		            e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
	if (count > RT_MAX_GPUS) return fail(RT_ERR_ARG, "at most %d GPUs", RT_MAX_GPUS);
	for (int i = 0; i < count; i++)
		if (devices[i] < 0 || devices[i] >= available)
			return fail(RT_ERR_ARG, "device %d not present (%d visible)", devices[i], available);

	rt_cuda_shutdown();
	g.ngpu = count;
	for (int i = 0; i < count; i++) {
		DeviceCtx &d = g.dev[i];
		d.device = devices[i];
		CU(cudaSetDevice(d.device));
		CU(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.device));
		CU(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
		{
			/* the copy stream also runs the tiny flag kernels of the pipelined composite: highest
			 * priority, so that they take the first free warp slot next to a resident render grid */
			int lo_prio = 0, hi_prio = 0;
			CU(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
			CU(cudaStreamCreateWithPriority(&d.copy_stream, cudaStreamNonBlocking, hi_prio));
		}
		for (int k = 0; k < 2; k++) {
			CU(cudaEventCreateWithFlags(&d.stage_rendered[k], cudaEventDisableTiming));
			CU(cudaEventCreateWithFlags(&d.stage_copied[k], cudaEventDisableTiming));
		}
		for (auto &ev : d.ev) CU(cudaEventCreate(&ev));
		CU(cudaMalloc(&d.ray_counter, 4 * sizeof(unsigned long long)));     /* [0] rays; [1], [2]: walk counters of the counter build */
		CU(cudaMalloc(&d.work_counter, RT_WORK_SLOTS * sizeof(unsigned int)));
		CU(cudaMemset(d.ray_counter, 0, 4 * sizeof(unsigned long long)));
		CU(cudaMallocHost(&d.host_rays, sizeof(unsigned long long)));
		float lut[256];
		rt_host_byte_lut(lut);
		CU(cudaMalloc(&d.lut, sizeof(lut)));
		CU(cudaMemcpy(d.lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
	}
	/* band GPUs store straight into GPU 0's framebuffer over NVLink */
	for (int i = 1; i < count; i++) {
		int can = 0;
		CU(cudaDeviceCanAccessPeer(&can, g.dev[i].device, g.dev[0].device));
		if (!can) return fail(RT_ERR_CUDA, "GPU %d cannot access GPU %d peer memory", g.dev[i].device, g.dev[0].device);
		CU(cudaSetDevice(g.dev[i].device));
		cudaError_t pe = cudaDeviceEnablePeerAccess(g.dev[0].device, 0);
		if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
			return fail(RT_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(pe));
		cudaGetLastError();
	}
	CU(cudaSetDevice(g.dev[0].device));
	g.ready = true;
	return RT_OK;
}

extern "C" int rt_cuda_init(int num_gpus)
{
	if (num_gpus <= 0) num_gpus = 1;
	int devs[RT_MAX_GPUS];
	for (int i = 0; i < num_gpus && i < RT_MAX_GPUS; i++) devs[i] = i;
	return init_devices(devs, num_gpus);
}

extern "C" int rt_cuda_init_device(int device) { return init_devices(&device, 1); }

static void gl_release(void);      /* CUDA-OpenGL presenter, end of this file */

extern "C" void rt_cuda_shutdown(void)
{
	gl_release();
	for (int i = 0; i < g.ngpu; i++) free_device(g.dev[i]);
	free(g.scene_cache);
	g = Context();
}

extern "C" int rt_cuda_num_gpus(void) { return g.ready ? g.ngpu : 0; }

static int require_ready(void)
{
	if (!g.ready) {
		int rc = rt_cuda_init(1);       /* lazy single-GPU init, like the north_star one-call API */
		if (rc != RT_OK) return rc;
	}
	return RT_OK;
}

extern "C" int rt_cuda_synchronize(void)
{
	if (!g.ready) return RT_OK;
	for (int i = 0; i < g.ngpu; i++) {
		CU(cudaSetDevice(g.dev[i].device));
		CU(cudaStreamSynchronize(g.dev[i].stream));
		CU(cudaStreamSynchronize(g.dev[i].copy_stream));
	}
	CU(cudaSetDevice(g.dev[0].device));
	return RT_OK;
}

/* ------------------------------------------------------------------ upload */

extern "C" int rt_cuda_upload_objects(const RtObject *objects, int n)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (n < 0 || (n > 0 && !objects)) return fail(RT_ERR_ARG, "bad object array");
	RtPackedScene ps;
	rc = rt_host_pack_scene(objects, n, &ps);
	if (rc != RT_OK) return fail(rc, "out of host memory packing %d objects", n);

	size_t cnt = n > 0 ? (size_t) n : 1;
	bool want_bvh = n > RT_LBVH_THRESHOLD;
	/* maximal runs of consecutive same-type objects, in scene order */
	std::vector<int2> runs;
	for (int i = 0; i < n;) {
		int ty = (int) objects[i].type, j = i + 1;
		while (j < n && (int) objects[j].type == ty && j - i < 0xffffff) j++;
		runs.push_back(make_int2(i, (j - i) | ((ty & 0x7f) << 24)));
		i = j;
	}
	/* from here on the old scene is gone on at least one device: a failure below
	 * must not leave the context claiming it still has one */
	g.have_scene = false;
	g.have_bvh = false;
	g.scene_epoch++;
	rt_lbvh_drop_topology_cache();      /* keyed by this call's packed scene: never reuse an older one */
	for (int i = 0; i < g.ngpu; i++) {
		DeviceCtx &d = g.dev[i];
		if ((rc = select_device(d)) != RT_OK) { rt_host_free_packed(&ps); return rc; }
		cudaStreamSynchronize(d.stream);
		cudaFree(d.geomA); cudaFree(d.geomB); cudaFree(d.mat); cudaFree(d.runs);
		d.geomA = d.geomB = d.mat = nullptr;
		d.runs = nullptr;
		if (cudaMalloc(&d.runs, sizeof(int2) * (runs.size() + 1)) != cudaSuccess ||
		    (!runs.empty() && cudaMemcpy(d.runs, runs.data(), sizeof(int2) * runs.size(), cudaMemcpyHostToDevice) != cudaSuccess)) {
			rt_host_free_packed(&ps);
			return fail(RT_ERR_CUDA, "scene upload (runs): %s", cudaGetErrorString(cudaGetLastError()));
		}
		rt_lbvh_free(&d.bvh);
		cudaError_t e;
		if ((e = cudaMalloc(&d.geomA, cnt * sizeof(float4))) != cudaSuccess ||
		    (e = cudaMalloc(&d.geomB, cnt * sizeof(float4))) != cudaSuccess ||
		    (e = cudaMalloc(&d.mat, cnt * RT_MAT_STRIDE * sizeof(float4))) != cudaSuccess ||
		    (e = cudaMemcpy(d.geomA, ps.geomA, cnt * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess ||
		    (e = cudaMemcpy(d.geomB, ps.geomB, cnt * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess ||
		    (e = cudaMemcpy(d.mat, ps.mat, cnt * RT_MAT_STRIDE * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess) {
			rt_host_free_packed(&ps);
			return fail(RT_ERR_CUDA, "scene upload: %s", cudaGetErrorString(e));
		}
		if (want_bvh) {
			rc = rt_lbvh_build(&d.bvh, d.geomA, d.geomB, n, &ps, d.stream);
			if (rc != RT_OK) {
				rt_host_free_packed(&ps);
				return fail(rc, "LBVH build failed: %s", rt_lbvh_last_error());
			}
		}
	}
	g.n = n;
	g.light_index = ps.light_index;
	g.light_pos = ps.light_pos;
	g.div_safe = ps.div_safe;
	g.num_runs = (int) runs.size();
	g.have_scene = true;
	g.have_bvh = want_bvh;
	rt_lbvh_drop_topology_cache();
	rt_host_free_packed(&ps);
	cudaSetDevice(g.dev[0].device);
	return RT_OK;
}

/* Objects changed in place (moved, resized, re-coloured; same count, same types in
 * the same order): refresh the device records and REFIT the LBVH instead of
 * rebuilding it (~1 ms instead of ~20 ms for 100 000 spheres). */
extern "C" int rt_cuda_update_objects(const RtObject *objects, int n)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!g.have_scene) return fail(RT_ERR_STATE, "no scene uploaded (rt_cuda_upload_objects)");
	if (n != g.n || (n > 0 && !objects)) return fail(RT_ERR_ARG, "rt_cuda_update_objects: %d objects, the uploaded scene has %d", n, g.n);
	std::vector<int2> runs;
	for (int i = 0; i < n;) {
		int ty = (int) objects[i].type, j = i + 1;
		while (j < n && (int) objects[j].type == ty && j - i < 0xffffff) j++;
		runs.push_back(make_int2(i, (j - i) | ((ty & 0x7f) << 24)));
		i = j;
	}
	if ((int) runs.size() != g.num_runs) return fail(RT_ERR_ARG, "rt_cuda_update_objects: object types changed; upload the scene again");
	RtPackedScene ps;
	rc = rt_host_pack_scene(objects, n, &ps);
	if (rc != RT_OK) return fail(rc, "out of host memory packing %d objects", n);
	size_t cnt = n > 0 ? (size_t) n : 1;
	g.scene_epoch++;                      /* tile costs of the old positions are void */
	for (int i = 0; i < g.ngpu; i++) {
		DeviceCtx &d = g.dev[i];
		if ((rc = select_device(d)) != RT_OK) { rt_host_free_packed(&ps); return rc; }
		cudaError_t e;
		if ((e = cudaStreamSynchronize(d.stream)) != cudaSuccess ||
		    (e = cudaMemcpy(d.geomA, ps.geomA, cnt * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess ||
		    (e = cudaMemcpy(d.geomB, ps.geomB, cnt * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess ||
		    (e = cudaMemcpy(d.mat, ps.mat, cnt * RT_MAT_STRIDE * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess ||
		    (!runs.empty() && (e = cudaMemcpy(d.runs, runs.data(), sizeof(int2) * runs.size(), cudaMemcpyHostToDevice)) != cudaSuccess)) {
			rt_host_free_packed(&ps);
			g.have_scene = false;
			return fail(RT_ERR_CUDA, "scene update: %s", cudaGetErrorString(e));
		}
		if (g.have_bvh && (rc = rt_lbvh_update(&d.bvh, d.geomA, d.geomB, &ps, d.stream)) != RT_OK) {
			rt_host_free_packed(&ps);
			g.have_scene = false;
			return fail(rc, "LBVH refit failed: %s", rt_lbvh_last_error());
		}
	}
	g.light_index = ps.light_index;
	g.light_pos = ps.light_pos;
	g.div_safe = ps.div_safe;
	rt_host_free_packed(&ps);
	if (g.scene_cache) g.scene_cache->num_objects = -1;      /* render_frame_cuda(scene, ...) compares afresh */
	cudaSetDevice(g.dev[0].device);
	return RT_OK;
}

extern "C" int rt_cuda_upload_scene(const RtScene *scene)
{
	if (!scene) return fail(RT_ERR_ARG, "scene is NULL");
	if (scene->num_objects < 0 || scene->num_objects > RT_MAX_OBJECTS)
		return fail(RT_ERR_ARG, "scene->num_objects = %d out of range", scene->num_objects);
	int rc = rt_cuda_upload_objects(scene->objects, scene->num_objects);
	if (rc != RT_OK) {
		if (g.scene_cache) g.scene_cache->num_objects = -1;    /* never equal to a caller's scene */
		return rc;
	}
	if (!g.scene_cache) g.scene_cache = (RtScene *) malloc(sizeof(RtScene));
	if (g.scene_cache) {
		memcpy(g.scene_cache->objects, scene->objects, sizeof(RtObject) * (size_t) scene->num_objects);
		g.scene_cache->num_objects = scene->num_objects;
	}
	return RT_OK;
}

extern "C" int rt_cuda_upload_skybox(const RtCubemap *sky)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!sky || sky->w <= 0 || sky->h <= 0 || sky->chan < 3)
		return fail(RT_ERR_ARG, "skybox must have w,h > 0 and >= 3 channels (sample_cubemap reads RGB)");
	for (int f = 0; f < 6; f++)
		if (!sky->data[f]) return fail(RT_ERR_ARG, "skybox face %d is NULL", f);

	/* RGB(A) rows -> RGBA8 so a texel is one aligned 4-byte load */
	size_t face = (size_t) sky->w * sky->h;
	std::vector<uchar4> staged;
	try { staged.resize(6 * face); } catch (...) { return fail(RT_ERR_NOMEM, "out of host memory staging the skybox"); }
	for (int f = 0; f < 6; f++) {
		const uint8_t *src = sky->data[f];
		uchar4 *dst = staged.data() + (size_t) f * face;
		for (size_t p = 0; p < face; p++) {
			const uint8_t *t = src + p * (size_t) sky->chan;
			dst[p] = make_uchar4(t[0], t[1], t[2], 255);
		}
	}
	for (int i = 0; i < g.ngpu; i++) {
		DeviceCtx &d = g.dev[i];
		if ((rc = select_device(d)) != RT_OK) return rc;
		cudaStreamSynchronize(d.stream);
		cudaFree(d.sky);
		d.sky = nullptr;
		CU(cudaMalloc(&d.sky, 6 * face * sizeof(uchar4)));
		CU(cudaMemcpy(d.sky, staged.data(), 6 * face * sizeof(uchar4), cudaMemcpyHostToDevice));
	}
	g.sky_w = sky->w;
	g.sky_h = sky->h;
	g.have_sky = true;
	cudaSetDevice(g.dev[0].device);
	return RT_OK;
}

/* ------------------------------------------------------------------ render */

extern "C" void rt_render_opts_default(RtRenderOpts *o)
{
	memset(o, 0, sizeof(*o));
	o->struct_size = sizeof(*o);
	o->scale = 1;
	o->num_columns = 1;
	o->fb_format = RT_FB_F32X3;
	o->fb_memory = RT_MEM_AUTO;
	o->variant = RT_VARIANT_EXACT;
	o->traversal = RT_TRAVERSAL_AUTO;
	o->kernel = RT_KERNEL_AUTO;
}

static size_t bytes_per_pixel(int fmt) { return fmt == RT_FB_U8X4 ? 4 : 12; }

static bool is_device_pointer(const void *p)
{
	cudaPointerAttributes a;
	cudaError_t e = cudaPointerGetAttributes(&a, p);
	if (e != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static void fill_views(DeviceCtx &d, RtRenderParams &P)
{
	P.scene.geomA = d.geomA;
	P.scene.geomB = d.geomB;
	P.scene.mat = d.mat;
	P.scene.n = g.n;
	P.scene.light_index = g.light_index;
	P.scene.light_pos = g.light_pos;
	P.scene.div_safe = g.div_safe;
	P.scene.runs = d.runs;
	P.scene.num_runs = g.num_runs;
	P.bvh = rt_lbvh_view(&d.bvh);
	P.sky.texels = d.sky;
	P.sky.w = g.sky_w;
	P.sky.h = g.sky_h;
	P.sky.face_stride = (size_t) g.sky_w * g.sky_h;
	P.byte_lut = d.lut;
	P.ray_counter = d.ray_counter;
	/* every launch gets its own tile counter (zeroed on its stream just before it),
	 * so that frames rendered on different caller streams may overlap */
	P.work_counter = d.work_counter + (d.launch_seq++ % RT_WORK_SLOTS);
}

static int pick_traversal(int requested, bool *lbvh)
{
	if (requested == RT_TRAVERSAL_LBVH) {
		if (!g.have_bvh) return fail(RT_ERR_STATE, "LBVH traversal requested but the scene has <= %d objects (no LBVH built)", RT_LBVH_THRESHOLD);
		*lbvh = true;
	} else if (requested == RT_TRAVERSAL_LINEAR) {
		if (g.n > RT_SMEM_MAX_OBJECTS)
			return fail(RT_ERR_ARG, "linear scan is limited to %d objects (shared-memory staging)", RT_SMEM_MAX_OBJECTS);
		*lbvh = false;
	} else
		*lbvh = g.have_bvh;
	return RT_OK;
}

static int ensure_accum(DeviceCtx &d, int w, int h, int row0, int rows, bool *fresh)
{
	size_t need = (size_t) w * rows * 3 * sizeof(float);
	if (d.accum && d.accum_w == w && d.accum_h == h && d.accum_row0 == row0 && d.accum_rows == rows) return RT_OK;
	*fresh = true;
	CU(cudaFree(d.accum));
	d.accum = nullptr;
	CU(cudaMalloc(&d.accum, need ? need : 4));
	d.accum_needs_clear = true;
	d.accum_bytes = need;
	d.accum_w = w; d.accum_h = h; d.accum_row0 = row0; d.accum_rows = rows;
	return RT_OK;
}

/* The buffers are zeroed lazily, on the stream the next accumulating pass is
 * launched on: a reset issued on the library's stream would race with passes a
 * caller runs on its own stream (opts->stream). */
extern "C" int rt_cuda_accum_reset(void)
{
	g.accum_count = 0.0f;
	if (!g.ready) return RT_OK;
	for (int i = 0; i < g.ngpu; i++)
		if (g.dev[i].accum) g.dev[i].accum_needs_clear = true;
	return RT_OK;
}

extern "C" float rt_cuda_accum_count(void) { return g.accum_count; }

struct PassPlan {
	int  w, h, scale, ncols;
	int  row0, row1;          /* output band of the whole call */
	bool lbvh, persistent, exact, wavefront, queued;
};

/* Rows of a band are dealt to GPUs (or ranks) in blocks of RT_INTERLEAVE_ROWS
 * output rows, round robin: sky-heavy and scene-heavy rows cost very different
 * amounts (1 vs up to 40 rays per pixel), so contiguous bands of H/N rows leave
 * most GPUs idle.  The block size is fixed in OUTPUT rows so that a GPU owns the
 * same pixels at every scale of a progressive sweep (its accumulation buffer
 * stays valid across passes). */
#define RT_INTERLEAVE_ROWS 16

static int interleave_shift(int scale)
{
	int rows = (scale >= 1 && scale <= RT_INTERLEAVE_ROWS && RT_INTERLEAVE_ROWS % scale == 0) ? RT_INTERLEAVE_ROWS / scale : 4;
	int sh = 0;
	while ((1 << sh) < rows) sh++;
	return sh;       /* low-res rows per block = 1 << shift */
}

/* ---- pipelined composite: flag words in front of a shared frame ---------- */

/* rt_cuda_shared_frame_create() allocates this header in the owner's memory,
 * RT_SHARED_HEADER_BYTES before the address it returns; the other ranks see it
 * through their cudaIpc mapping of the same allocation. */
#define RT_SHARED_HEADER_BYTES 4096
struct SharedHeader {
	unsigned int arrived[RT_MAX_GPUS];   /* arrived[r] = last frame whose blocks rank r has copied in */
	unsigned int consumed;               /* last frame the owner released */
	unsigned int error;                  /* a polling kernel gave up (bit 0: wait, bit 1: ack) */
};

static SharedHeader *header_of(void *frame) { return (SharedHeader *) ((char *) frame - RT_SHARED_HEADER_BYTES); }

__device__ __forceinline__ unsigned long long global_ns(void)
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

/* seq numbers wrap: a is at or past b */
__device__ __forceinline__ bool seq_reached(unsigned a, unsigned b) { return (int) (a - b) >= 0; }

#define RT_FLAG_TIMEOUT_NS 2000000000ull

/* after this rank's blocks of frame `seq` were copied (same stream) */
__global__ void flag_arrive_kernel(SharedHeader *h, int rank, unsigned seq)
{
	__threadfence_system();
	*(volatile unsigned int *) &h->arrived[rank] = seq;
}

/* owner: returns once every rank's blocks of frame `seq` have landed */
__global__ void flag_wait_kernel(SharedHeader *h, int num_ranks, unsigned seq)
{
	int r = threadIdx.x;
	if (r >= num_ranks) return;
	unsigned long long t0 = global_ns();
	while (!seq_reached(*(volatile unsigned int *) &h->arrived[r], seq)) {
		__nanosleep(200);
		if (global_ns() - t0 > RT_FLAG_TIMEOUT_NS) { atomicOr(&h->error, 1u); break; }
	}
	__threadfence_system();
}

__global__ void flag_release_kernel(SharedHeader *h, unsigned seq)
{
	__threadfence_system();
	*(volatile unsigned int *) &h->consumed = seq;
}

/* any rank, before overwriting the shared frame with frame `seq`: the owner is done with seq - 1 */
__global__ void flag_ack_kernel(SharedHeader *h, unsigned seq)
{
	unsigned long long t0 = global_ns();
	while (!seq_reached(*(volatile unsigned int *) &h->consumed, seq - 1u)) {
		__nanosleep(500);
		if (global_ns() - t0 > RT_FLAG_TIMEOUT_NS) { atomicOr(&h->error, 2u); break; }
	}
}

/* Copy the row blocks this GPU owns (interleave il_i of il_n) from its local
 * frame to the same rows of `dst` (GPU 0's frame, peer memory) with the copy
 * engine.  The render kernel stores 4-byte words pixel by pixel as paths end;
 * done straight over NVLink those small scattered writes ran at ~12 GB/s per
 * GPU (8-GPU 4K frame: 1.09 ms kernel vs 0.40 ms on local memory), so remote
 * GPUs render locally and ship their blocks as one strided 2D copy. */
static int copy_owned_blocks(void *dst, const void *src, const PassPlan &pl, int fb_row_offset, int il_n, int il_i,
                             size_t bpp, cudaStream_t stream, cudaMemcpyKind kind = cudaMemcpyDeviceToDevice)
{
	int rows_per_block = (1 << interleave_shift(pl.scale)) * pl.scale;
	int band = pl.row1 - pl.row0;
	size_t row_bytes = (size_t) pl.w * bpp;
	size_t chunk = (size_t) rows_per_block * row_bytes;
	int nblocks = (band + rows_per_block - 1) / rows_per_block;
	int last = nblocks - 1;
	int last_rows = band - last * rows_per_block;
	int mine = il_i < nblocks ? (nblocks - 1 - il_i) / il_n + 1 : 0;          /* blocks il_i, il_i + il_n, ... */
	if (mine <= 0) return RT_OK;
	bool own_partial_last = last_rows != rows_per_block && last % il_n == il_i;
	int full = own_partial_last ? mine - 1 : mine;
	size_t base = (size_t) (pl.row0 - fb_row_offset) * row_bytes + (size_t) il_i * chunk;
	if (full > 0)
		CU(cudaMemcpy2DAsync((char *) dst + base, (size_t) il_n * chunk, (const char *) src + base, (size_t) il_n * chunk,
		                     chunk, (size_t) full, kind, stream));
	if (own_partial_last) {
		size_t off = (size_t) (pl.row0 - fb_row_offset) * row_bytes + (size_t) last * chunk;
		CU(cudaMemcpyAsync((char *) dst + off, (const char *) src + off, (size_t) last_rows * row_bytes, kind, stream));
	}
	return RT_OK;
}

/* Pipelined composite, second half: staging frame `slot` of this rank holds its blocks of frame
 * o.frame_seq (rendered on `st`); the copy stream ships them into the shared frame `fb` and raises
 * the rank's arrived flag.  Returns after queueing. */
static int ship_piped_peer(DeviceCtx &d0, const PassPlan &pl, const RtRenderOpts &o, void *fb, int fb_row_offset, int slot,
                           int il_n, int il_i, size_t bpp, cudaStream_t st)
{
	int rc;
	SharedHeader *hdr = header_of(fb);
	CU(cudaEventRecord(d0.stage_rendered[slot], st));
	CU(cudaStreamWaitEvent(d0.copy_stream, d0.stage_rendered[slot], 0));
	if (o.frame_ack) {
		flag_ack_kernel<<<1, 1, 0, d0.copy_stream>>>(hdr, o.frame_seq);
		CU(cudaGetLastError());
	}
	if ((rc = copy_owned_blocks(fb, d0.stage[slot], pl, fb_row_offset, il_n, il_i, bpp, d0.copy_stream)) != RT_OK) return rc;
	flag_arrive_kernel<<<1, 1, 0, d0.copy_stream>>>(hdr, il_i, o.frame_seq);
	CU(cudaGetLastError());
	CU(cudaEventRecord(d0.stage_copied[slot], d0.copy_stream));
	return RT_OK;
}

/* Pixels the reference's pass never writes (main.c:285-290: rows >= lh*scale;
 * main.c:363: columns >= T*column_w) hold 0.  Every GPU clears them in ITS
 * render target and only inside the row blocks it owns (copy_owned_blocks ships
 * whole rows of exactly those blocks), so no two GPUs ever write the same byte. */
static int clear_uncovered_owned(void *fb, const PassPlan &pl, int fb_row_offset, int il_n, int il_i, size_t bpp,
                                 int covered_rows_end, bool columns_uncovered, cudaStream_t stream, int *launches)
{
	int rows_per_block = (1 << interleave_shift(pl.scale)) * pl.scale;
	int band = pl.row1 - pl.row0;
	size_t row_bytes = (size_t) pl.w * bpp;
	size_t chunk = (size_t) rows_per_block * row_bytes;
	int nblocks = (band + rows_per_block - 1) / rows_per_block;
	int last = nblocks - 1;
	int last_rows = band - last * rows_per_block;
	char *base = (char *) fb + (size_t) (pl.row0 - fb_row_offset) * row_bytes;
	if (columns_uncovered) {
		/* rare (W % T != 0): zero every owned row, the launch then fills the covered pixels */
		int mine = il_i < nblocks ? (nblocks - 1 - il_i) / il_n + 1 : 0;
		bool own_last = mine > 0 && last % il_n == il_i;
		int full = own_last && last_rows != rows_per_block ? mine - 1 : mine;
		if (full > 0) {
			CU(cudaMemset2DAsync(base + (size_t) il_i * chunk, (size_t) il_n * chunk, 0, chunk, (size_t) full, stream));
			(*launches)++;
		}
		if (own_last && last_rows != rows_per_block) {
			CU(cudaMemsetAsync(base + (size_t) last * chunk, 0, (size_t) last_rows * row_bytes, stream));
			(*launches)++;
		}
		return RT_OK;
	}
	if (covered_rows_end < pl.row1 && last % il_n == il_i) {
		/* rows [lh*scale, H) lie in the last block */
		int s0 = std::max(covered_rows_end, pl.row0);
		CU(cudaMemsetAsync((char *) fb + (size_t) (s0 - fb_row_offset) * row_bytes, 0, (size_t) (pl.row1 - s0) * row_bytes, stream));
		(*launches)++;
	}
	return RT_OK;
}

/*
 * Longest tiles first.  A warp of the queued kernel works on 8x4-pixel tiles and
 * a tile lasts as long as its longest path (1 to 40 rays, each a warp step of
 * several microseconds), so whatever tiles are handed out last decide how long
 * the launch drags on after the work has run out (~0.1 ms per launch with tiles
 * in image order: 5 % of a 4K frame, 30 % of one GPU's share of it at 8 GPUs).
 * What a tile costs is a property of the POSE, not of the pass: the reference's
 * loop renders a pose at scale 16, 8, 4, 2, 1, 1, ... (main.c:354-403), and a
 * coarse pass samples the same surfaces as the fine ones.  Every launch of a
 * pose that is finer than any before it records the largest bounce count per
 * tile (one atomicMax per non-sky pixel: +0.5 % on a 4K frame), and a launch is
 * ordered -- stable counting sort, costly tiles first, image order within a
 * class -- by the costs recorded at ITS scale: the first scale-1 pass of a pose
 * records, every later one is ordered (1.97 -> 1.915 ms at 4K).  Ordering a pass
 * by the costs of a COARSER pass of the pose (a fine tile takes the cost of the
 * coarse tile covering it) was built and measured for the moving camera, whose
 * poses never repeat: first scale-1 pass 2.00 ms seeded by scale 2 against
 * 1.97 ms in image order -- single-tile claims in a merely approximate order
 * cost more than the shorter tail saves.  Kept as rt_cuda_debug_set_tile_schedule(2).
 * Scheduling only -- every pixel is computed by the same code from the same
 * key, so frames are bit-identical with and without it (test_tile_schedule).
 */
#define RT_ORDER_THREADS 256
#define RT_ORDER_MAX_CTAS 128
#define RT_ORDER_CHUNK_MIN 2048u        /* tiles per CTA at least */
/* Stable counting sort of `n` tiles (tiles_x per row) by descending cost class (costs above 15
 * share the top class), image order within a class.  Tile (tx, ty) takes the cost of tile
 * (tx >> shift, ty >> shift) of the cost map (cost_tiles_x per row).  Two kernels over the same
 * grid, CTA c owning the contiguous chunk [c * chunk, (c + 1) * chunk):
 *   tile_hist_kernel     class histogram of every chunk -> hist[c][16]
 *   tile_scatter_kernel  base of (class, chunk) from the histograms, then the same stable sort
 *                        inside the chunk: classes gathered into shared memory with coalesced
 *                        loads, thread t counts and scatters its own contiguous run.
 * (A single-CTA version took 0.33 ms for the 259 200 tiles of a 4K frame -- three times what the
 * order then saved; this one takes a few microseconds.) */
__device__ __forceinline__ unsigned tile_class(const unsigned int *cost, unsigned i, unsigned tiles_x, unsigned cost_tiles_x,
                                               unsigned cost_tiles_y, unsigned shift)
{
	unsigned c;
	if (shift == 0 && cost_tiles_x == tiles_x) c = __ldg(cost + i);
	else {
		unsigned ty = i / tiles_x, tx = i - ty * tiles_x;
		unsigned cx = min(tx >> shift, cost_tiles_x - 1u), cy = min(ty >> shift, cost_tiles_y - 1u);
		c = __ldg(cost + (size_t) cy * cost_tiles_x + cx);
	}
	return min(c, 15u);
}

__global__ void __launch_bounds__(RT_ORDER_THREADS)
tile_hist_kernel(const unsigned int *cost, unsigned n, unsigned chunk, unsigned tiles_x, unsigned cost_tiles_x,
                 unsigned cost_tiles_y, unsigned shift, unsigned int *hist)
{
	__shared__ unsigned h[16];
	if (threadIdx.x < 16) h[threadIdx.x] = 0;
	__syncthreads();
	const unsigned lo = blockIdx.x * chunk, hi = min(lo + chunk, n);
	for (unsigned i = lo + threadIdx.x; i < hi; i += RT_ORDER_THREADS)
		atomicAdd(&h[tile_class(cost, i, tiles_x, cost_tiles_x, cost_tiles_y, shift)], 1u);
	__syncthreads();
	if (threadIdx.x < 16) hist[blockIdx.x * 16 + threadIdx.x] = h[threadIdx.x];
}

/* dynamic shared memory: `chunk` class bytes */
__global__ void __launch_bounds__(RT_ORDER_THREADS)
tile_scatter_kernel(const unsigned int *cost, unsigned int *order, unsigned n, unsigned chunk, unsigned tiles_x,
                    unsigned cost_tiles_x, unsigned cost_tiles_y, unsigned shift, const unsigned int *hist)
{
	extern __shared__ unsigned char cls[];
	__shared__ unsigned cnt[16][RT_ORDER_THREADS];
	__shared__ unsigned base[16];
	const unsigned t = threadIdx.x;
	const unsigned lo = blockIdx.x * chunk, hi = min(lo + chunk, n), m = hi > lo ? hi - lo : 0;
	for (unsigned i = t; i < m; i += RT_ORDER_THREADS) cls[i] = (unsigned char) tile_class(cost, lo + i, tiles_x, cost_tiles_x, cost_tiles_y, shift);
	for (int b = 0; b < 16; b++) cnt[b][t] = 0;
	if (t < 16) {
		/* where class t of this chunk starts: all tiles of costlier classes, then class t of earlier chunks */
		unsigned before = 0;
		for (unsigned c = 0; c < gridDim.x; c++) {
			for (unsigned b = t + 1; b < 16; b++) before += hist[c * 16 + b];
			if (c < blockIdx.x) before += hist[c * 16 + t];
		}
		base[t] = before;
	}
	__syncthreads();
	const unsigned run_len = (m + RT_ORDER_THREADS - 1) / RT_ORDER_THREADS;
	const unsigned a = min(t * run_len, m), z = min(a + run_len, m);
	for (unsigned i = a; i < z; i++) cnt[cls[i]][t]++;
	__syncthreads();
	/* exclusive scan over the threads, per class: warp w scans classes w and w + 8 */
	for (unsigned b = t >> 5; b < 16; b += RT_ORDER_THREADS / 32) {
		const unsigned lane = t & 31;
		unsigned run = 0;
		for (unsigned k0 = 0; k0 < RT_ORDER_THREADS; k0 += 32) {
			unsigned c = cnt[b][k0 + lane], x = c;
			for (int o = 1; o < 32; o <<= 1) {
				unsigned y = __shfl_up_sync(0xffffffffu, x, o);
				if (lane >= (unsigned) o) x += y;
			}
			cnt[b][k0 + lane] = run + x - c;
			run += __shfl_sync(0xffffffffu, x, 31);
		}
	}
	__syncthreads();
	for (unsigned i = a; i < z; i++) {
		unsigned b = cls[i];
		order[base[b] + cnt[b][t]++] = lo + i;
	}
}

/* hist: RT_ORDER_MAX_CTAS * 16 counters of scratch */
static cudaError_t launch_tile_order(const unsigned int *cost, unsigned int *order, size_t n, int tiles_x, int cost_tiles_x,
                                     int cost_tiles_y, int shift, unsigned int *hist, cudaStream_t stream)
{
	unsigned ctas = (unsigned) std::min<size_t>(RT_ORDER_MAX_CTAS, (n + RT_ORDER_CHUNK_MIN - 1) / RT_ORDER_CHUNK_MIN);
	if (ctas == 0) ctas = 1;
	unsigned chunk = (unsigned) ((n + ctas - 1) / ctas);
	tile_hist_kernel<<<ctas, RT_ORDER_THREADS, 0, stream>>>(cost, (unsigned) n, chunk, (unsigned) tiles_x, (unsigned) cost_tiles_x,
	                                                          (unsigned) cost_tiles_y, (unsigned) shift, hist);
	tile_scatter_kernel<<<ctas, RT_ORDER_THREADS, (chunk + 15) & ~15u, stream>>>(cost, order, (unsigned) n, chunk, (unsigned) tiles_x,
	                                                                            (unsigned) cost_tiles_x, (unsigned) cost_tiles_y, (unsigned) shift, hist);
	return cudaGetLastError();
}

static bool sched_reserve(unsigned int **buf, size_t *cap, size_t need)
{
	if (*cap >= need) return true;
	cudaFree(*buf);          /* waits for everything that may still read it */
	*buf = nullptr; *cap = 0;
	if (cudaMalloc(buf, need * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); return false; }
	*cap = need;
	return true;
}

/* Decide what this launch does about the schedule: sets P.tile_order (hand the tiles out in this
 * order) and / or P.tile_cost (record costs).  Returns true when tile_schedule_done() must follow
 * the launch. */
static bool tile_schedule(DeviceCtx &d, const TileKey &key, int scale, int tiles_x, int tiles_y, RtRenderParams &P, cudaStream_t stream, int *launches)
{
	TileSched &S = d.sched;
	P.tile_order = nullptr;
	P.tile_cost = nullptr;
	const size_t tiles = (size_t) tiles_x * tiles_y;
	if (!g.tile_schedule || tiles >= (1u << 20)) { S.have_key = false; return false; }
	if (!S.have_key || memcmp(&key, &S.key, sizeof(TileKey)) != 0) {
		S.key = key;
		S.have_key = true;
		S.cost_cur = S.order_cur = -1;      /* a new pose: nothing is known about it */
	}
	if (!S.fence && cudaEventCreateWithFlags(&S.fence, cudaEventDisableTiming) != cudaSuccess) return false;
	if (!S.write_fence && cudaEventCreateWithFlags(&S.write_fence, cudaEventDisableTiming) != cudaSuccess) return false;
	/* what this launch will do (decided before anything is queued, so that it can wait first) */
	bool keeps = false, builds = false, records = false;
	if (tiles >= 2048) {
		const bool exact_costs = S.cost_cur >= 0 && S.cost_scale == scale;
		keeps = S.order_cur >= 0 && S.order_scale == scale && S.order_tiles == tiles && (S.order_from_scale == scale || !exact_costs);
		builds = !keeps && S.cost_cur >= 0 && S.cost_scale >= scale && S.cost_scale % scale == 0 && (g.tile_schedule == 2 || S.cost_scale == scale);
	}
	records = tiles >= 64 && (S.cost_cur < 0 || S.cost_scale > scale);
	S.launch_writes = builds || records;
	if (S.launch_writes) {
		for (int i = 0; i < S.nuses; i++)
			if (S.uses[i].st != stream) cudaStreamWaitEvent(stream, S.uses[i].ev, 0);
	} else if (S.write_stream && S.write_stream != stream)
		cudaStreamWaitEvent(stream, S.write_fence, 0);
	S.last_stream = stream;
	bool touched = false;

	/* ---- the order of this launch ---- */
	if (tiles >= 2048) {
		const bool exact_costs = S.cost_cur >= 0 && S.cost_scale == scale;
		if (S.order_cur >= 0 && S.order_scale == scale && S.order_tiles == tiles && (S.order_from_scale == scale || !exact_costs)) {
			P.tile_order = S.order[S.order_cur];            /* same pose, same scale: keep the order */
			touched = true;
		} else if (S.cost_cur >= 0 && S.cost_scale >= scale && S.cost_scale % scale == 0 && (g.tile_schedule == 2 || S.cost_scale == scale)) {
			int ratio = S.cost_scale / scale, shift = 0;
			while ((1 << shift) < ratio) shift++;
			int o = S.order_cur == 0 ? 1 : 0;
			if ((1 << shift) == ratio && shift <= 4 && sched_reserve(&S.order[o], &S.order_cap[o], tiles)) {
				cudaError_t le = S.hist || cudaMalloc(&S.hist, RT_ORDER_MAX_CTAS * 16 * sizeof(unsigned)) == cudaSuccess
				                     ? launch_tile_order(S.cost[S.cost_cur], S.order[o], tiles, tiles_x, S.cost_tiles_x, S.cost_tiles_y, shift, S.hist, stream)
				                     : cudaErrorMemoryAllocation;
				(*launches) += 2;
				if (le != cudaSuccess) fprintf(stderr, "rt_cuda: tile_order_kernel launch failed (%zu tiles): %s\n", tiles, cudaGetErrorString(le));
				if (le == cudaSuccess) {
					S.order_cur = o; S.order_scale = scale; S.order_from_scale = S.cost_scale; S.order_tiles = tiles;
					P.tile_order = S.order[o];
					touched = true;
				}
			}
		}
	}
	/* ---- record, when this launch sees the pose finer than any launch before it ---- */
	if (tiles >= 64 && (S.cost_cur < 0 || S.cost_scale > scale)) {
		int c = S.cost_cur == 0 ? 1 : 0;
		if (sched_reserve(&S.cost[c], &S.cost_cap[c], tiles) &&
		    cudaMemsetAsync(S.cost[c], 0, tiles * sizeof(unsigned), stream) == cudaSuccess) {
			P.tile_cost = S.cost[c];
			(*launches)++;
			touched = true;
		}
	}
	return touched;
}

/* after the launch: the recorded costs become the pose's cost map */
static void tile_schedule_done(DeviceCtx &d, const RtRenderParams &P, int scale, int tiles_x, int tiles_y, cudaStream_t stream)
{
	TileSched &S = d.sched;
	if (P.tile_cost) {
		S.cost_cur = P.tile_cost == S.cost[0] ? 0 : 1;
		S.cost_scale = scale; S.cost_tiles_x = tiles_x; S.cost_tiles_y = tiles_y;
	}
	cudaEventRecord(S.fence, stream);
	int slot = -1;
	for (int i = 0; i < S.nuses; i++) if (S.uses[i].st == stream) slot = i;
	if (slot < 0 && S.nuses < 4 && cudaEventCreateWithFlags(&S.uses[S.nuses].ev, cudaEventDisableTiming) == cudaSuccess)
		slot = S.nuses++;
	if (slot < 0) {
		/* more streams than slots: take over the slot of a stream whose last use is over (waiting for
		 * one if need be: callers that rotate through many streams are rare) */
		for (int i = 0; i < S.nuses && slot < 0; i++) if (cudaEventQuery(S.uses[i].ev) == cudaSuccess) slot = i;
		if (slot < 0) { slot = 0; cudaEventSynchronize(S.uses[0].ev); }
		cudaGetLastError();
	}
	S.uses[slot].st = stream;
	cudaEventRecord(S.uses[slot].ev, stream);
	if (S.launch_writes) {
		cudaEventRecord(S.write_fence, stream);
		S.write_stream = stream;
	}
}

/* Launch one pass for one device over output rows [r0, r1) (scale aligned). */
struct LaunchExtra {
	bool  compact = false;       /* one value per low-res cell instead of the replicated tiles (RtRenderParams::compact) */
	bool  no_schedule = false;   /* leave the pose's tile schedule alone (launches that overlap one another) */
	float grid_share = 1.0f;     /* persistent kernels: fraction of the resident CTA slots this launch may take */
	bool  no_clear = false;      /* the never-written pixels of the CALL's band were cleared by an earlier launch of the call */
	bool  no_dense = false;      /* keep the queued kernel's 6-CTA build (launches that share the SMs with other launches) */
};

static int launch_band(DeviceCtx &d, const RtCamera *cam, const PassPlan &pl, const RtRenderOpts *o,
                       void *fb, int fb_row_offset, int r0, int r1, int il_n, int il_i, cudaStream_t stream,
                       bool accumulate, float accum_weight, float inv_count, int *launches, const LaunchExtra *extra = nullptr)
{
	const LaunchExtra none;
	const LaunchExtra &X = extra ? *extra : none;
	RtRenderParams P;
	memset(&P, 0, sizeof(P));
	fill_views(d, P);
	rt_host_camera_frame(cam, (float) pl.w / pl.h, &P.cam);       /* main.c:281 aspect */
	P.W = pl.w; P.H = pl.h; P.scale = pl.scale;
	P.num_columns = pl.ncols;
	P.column_w = pl.w / pl.ncols;                                   /* main.c:363 */
	P.lw = pl.w / pl.scale;                                         /* main.c:284-285 */
	P.lh = pl.h / pl.scale;
	P.cells_per_col = (P.column_w + pl.scale - 1) / pl.scale;
	P.cells_per_row = pl.ncols * P.cells_per_col;
	P.lrow0 = r0 / pl.scale;
	P.lrow1 = std::min((r1 + pl.scale - 1) / pl.scale, P.lh);
	if (P.lrow1 < P.lrow0) P.lrow1 = P.lrow0;
	P.il_n = il_n > 1 ? il_n : 1;
	P.il_i = il_n > 1 ? il_i : 0;
	P.il_shift = interleave_shift(pl.scale);
	{
		/* low-res rows of [lrow0, lrow1) owned by this GPU: blocks b with b % il_n == il_i */
		int per = 1 << P.il_shift, total = P.lrow1 - P.lrow0, mine = 0;
		for (int b = P.il_i, start = P.il_i * per; start < total; b += P.il_n, start += P.il_n * per)
			mine += std::min(per, total - start);
		P.local_rows = mine;
	}
	P.tiles_x = (P.cells_per_row + RT_TILE_W - 1) / RT_TILE_W;
	P.tiles_y = (P.local_rows + RT_TILE_H - 1) / RT_TILE_H;
	P.pass_mix = rt_host_splitmix64(o->pass_index);
	P.magic_tiles_x = ((1ull << 40) + (unsigned long long) std::max(P.tiles_x, 1) - 1) / (unsigned long long) std::max(P.tiles_x, 1);
	P.magic_cells_per_col = ((1ull << 40) + (unsigned long long) std::max(P.cells_per_col, 1) - 1) / (unsigned long long) std::max(P.cells_per_col, 1);
	P.sweep_tau2 = g.sweep_tau2;
	P.fb = fb;
	P.fb_format = o->fb_format;
	P.fb_row_offset = fb_row_offset;
	P.compact = X.compact ? 1 : 0;
	P.store_scale = X.compact ? 1 : pl.scale;
	P.store_stride = X.compact ? P.cells_per_row : pl.w;
	P.accum = accumulate ? d.accum : nullptr;
	P.accum_row_offset = d.accum_row0;
	P.accum_weight = accum_weight;
	P.inv_count = inv_count;

	/* pixels the reference's pass never writes hold 0 (clear_uncovered_owned) */
	int covered_rows_end = std::min(P.lh * pl.scale, pl.row1);    /* of the whole call (r1 == pl.row1 except for the bands of a banded read-back) */
	bool columns_uncovered = P.column_w * pl.ncols < pl.w;
	if (!X.compact && !X.no_clear && (covered_rows_end < pl.row1 || columns_uncovered)) {
		int rc = clear_uncovered_owned(fb, pl, fb_row_offset, P.il_n, P.il_i, bytes_per_pixel(o->fb_format),
		                               covered_rows_end, columns_uncovered, stream, launches);
		if (rc != RT_OK) return rc;
	}

	if (P.tiles_x > 0 && P.tiles_y > 0) {
		int grid = 0;
		if (pl.wavefront) {
			CU(cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned int), stream));
			static int wf_per_sm[2][2] = {{0, 0}, {0, 0}}, wf_n[2][2] = {{-1, -1}, {-1, -1}};
			int &per_sm = wf_per_sm[pl.exact ? 0 : 1][pl.lbvh ? 1 : 0];
			int &for_n = wf_n[pl.exact ? 0 : 1][pl.lbvh ? 1 : 0];
			if (per_sm < 1 || for_n != P.scene.n * 4096 + P.scene.num_runs) {
				CU((pl.exact ? rt_exact_wavefront_blocks_per_sm : rt_fast_wavefront_blocks_per_sm)(&P, pl.lbvh, &per_sm));
				if (per_sm < 1) return fail(RT_ERR_ARG, "the wavefront kernel does not fit: %d objects staged next to the path pool", P.scene.n);
				for_n = P.scene.n * 4096 + P.scene.num_runs;
			}
			unsigned paths = (unsigned) rt_exact_wavefront_paths_per_block();
			unsigned pixels = (unsigned) P.tiles_x * P.tiles_y * 32u;
			grid = (int) std::min<unsigned>((unsigned) (d.sm_count * per_sm), (pixels + paths - 1) / paths);
			CU((pl.exact ? rt_exact_launch_render : rt_fast_launch_render)(&P, pl.lbvh, 2, grid, stream));
			(*launches)++;
		} else {
		bool dense = false;
		if (pl.persistent) {
			CU(cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned int), stream));
			/* occupancy of the persistent kernel, queried once per (variant, traversal, scene size) */
			/* big launches of linear-scan scenes take the queued kernel's 7-CTA build (rt_render.cu) */
			dense = pl.queued && !pl.lbvh && g.queued_dense && !X.no_dense && (size_t) P.tiles_x * (size_t) P.tiles_y >= 60000;
			const int qk = pl.queued ? (dense ? 2 : 1) : 0;
			static int cached_per_sm[3][2][2] = {};
			static int cached_n[3][2][2] = {{{-1, -1}, {-1, -1}}, {{-1, -1}, {-1, -1}}, {{-1, -1}, {-1, -1}}};
			int &per_sm = cached_per_sm[qk][pl.exact ? 0 : 1][pl.lbvh ? 1 : 0];
			int &for_n = cached_n[qk][pl.exact ? 0 : 1][pl.lbvh ? 1 : 0];
			if (per_sm < 1 || for_n != P.scene.n * 4096 + P.scene.num_runs) {
				CU((pl.exact ? rt_exact_persistent_blocks_per_sm : rt_fast_persistent_blocks_per_sm)(&P, pl.lbvh, qk, &per_sm));
				if (per_sm < 1) per_sm = 1;
				for_n = P.scene.n * 4096 + P.scene.num_runs;
			}
			unsigned warps_needed = (unsigned) P.tiles_x * P.tiles_y;
			unsigned blocks_needed = (warps_needed + (RT_BLOCK_THREADS / 32) - 1) / (RT_BLOCK_THREADS / 32);
			unsigned slots = (unsigned) std::max(1, (int) ((float) (d.sm_count * per_sm) * X.grid_share));
			grid = (int) std::min<unsigned>(slots, blocks_needed);
		}
		bool sched = false;
		if ((pl.queued || (pl.lbvh && pl.persistent && !pl.wavefront)) && !X.no_schedule) {     /* the kernels that read P.tile_order */
			TileKey key;
			memset(&key, 0, sizeof(key));
			key.w = pl.w; key.h = pl.h; key.ncols = pl.ncols; key.r0 = r0; key.r1 = r1;
			key.il_n = P.il_n; key.il_i = P.il_i; key.lbvh = pl.lbvh; key.scene_epoch = g.scene_epoch;
			key.cam.pos = cam->pos; key.cam.front = cam->front; key.cam.up = cam->up; key.cam.fov = cam->fov;
			sched = tile_schedule(d, key, pl.scale, P.tiles_x, P.tiles_y, P, stream, launches);
			if (getenv("RT_SCHED_DEBUG"))
				fprintf(stderr, "sched: %dx%d tiles scale %d order %p cost %p (cost_cur %d cost_scale %d order_cur %d order_scale %d from %d)\n",
				        P.tiles_x, P.tiles_y, pl.scale, (void *) P.tile_order, (void *) P.tile_cost, d.sched.cost_cur, d.sched.cost_scale,
				        d.sched.order_cur, d.sched.order_scale, d.sched.order_from_scale);
		}
		CU((pl.exact ? rt_exact_launch_render : rt_fast_launch_render)(&P, pl.lbvh, pl.queued ? (dense ? 4 : 3) : (pl.persistent ? 1 : 0), grid, stream));
		(*launches)++;
		if (sched) tile_schedule_done(d, P, pl.scale, P.tiles_x, P.tiles_y, stream);
		}
	}

	return RT_OK;
}

static int validate_common(const RtCamera *cam, void *fb, int w, int h, const RtRenderOpts *o)
{
	if (!cam || !fb) return fail(RT_ERR_ARG, "camera and framebuffer must be non-NULL");
	if (w <= 0 || h <= 0) return fail(RT_ERR_ARG, "bad frame size %dx%d", w, h);
	if (!o || o->struct_size != sizeof(RtRenderOpts)) return fail(RT_ERR_ARG, "opts->struct_size mismatch (ABI)");
	if (o->scale < 1) return fail(RT_ERR_ARG, "scale must be >= 1");
	if (o->num_columns < 1 || o->num_columns > w) return fail(RT_ERR_ARG, "num_columns out of range");
	if (w / o->scale < 2 || h / o->scale < 2)
		return fail(RT_ERR_ARG, "frame %dx%d is too small for scale %d: u = i/(w/scale - 1), v = j/(h/scale - 1) divide by zero (main.c:293-294)", w, h, o->scale);
	if (o->fb_format != RT_FB_F32X3 && o->fb_format != RT_FB_U8X4) return fail(RT_ERR_ARG, "unknown fb_format");
	if (o->variant != RT_VARIANT_EXACT && o->variant != RT_VARIANT_FAST) return fail(RT_ERR_ARG, "unknown variant");
	if (!g.have_scene) return fail(RT_ERR_STATE, "no scene uploaded (rt_cuda_upload_scene)");
	if (!g.have_sky) return fail(RT_ERR_STATE, "no skybox uploaded (rt_cuda_upload_skybox)");
	return RT_OK;
}

/* traversal, build and kernel of a pass (pl.w, pl.h, pl.scale set by the caller) */
static int plan_kernel(PassPlan &pl, const RtRenderOpts *o)
{
	int rc = pick_traversal(o->traversal, &pl.lbvh);
	if (rc != RT_OK) return rc;
	pl.exact = o->variant == RT_VARIANT_EXACT;
	/* AUTO: the queued kernel for linear-scan scenes (4K scene_0: 1.88 ms against 2.20 ms persistent),
	 * the persistent kernel over the LBVH (100k spheres 4K: 39.3 ms against 44.5 ms: its smaller shared
	 * memory footprint lets 8 CTAs per SM hide the node-fetch latency) */
	pl.queued = (o->kernel == RT_KERNEL_QUEUED || (o->kernel == RT_KERNEL_AUTO && !pl.lbvh)) && o->scale <= 64;   /* tile width is packed into 7 bits */
	pl.persistent = o->kernel == RT_KERNEL_PERSISTENT || o->kernel == RT_KERNEL_AUTO || pl.queued;
	pl.wavefront = o->kernel == RT_KERNEL_WAVEFRONT;
	/* a tree deeper than the shared-memory traversal stacks is walked by the local-stack build
	 * of the persistent kernel (rt_render.cu: launch_render), whatever was asked for */
	if (pl.lbvh && g.dev[0].bvh.depth > RT_SMEM_STACK) {
		pl.queued = pl.wavefront = false;
		pl.persistent = true;
	}
	/* the wavefront kernel packs the tile width into 5 bits and a pixel's x, y into 16 bits each */
	if (pl.wavefront && (o->scale > 31 || pl.w > 65535 || pl.h > 65535))
		return fail(RT_ERR_ARG, "RT_KERNEL_WAVEFRONT supports scale <= 31 and frames up to 65535x65535");

	return RT_OK;
}

/* ship = false: a pass of a sweep whose frame nobody will look at (only the accumulation
 * matters): a rank with a remote frame renders it locally and sends nothing */
static int render_pass(const RtCamera *cam, void *fb, int w, int h, const RtRenderOpts *o,
                       bool accumulate, RtRenderStats *stats, bool sync_and_copy, bool ship = true)
{
	PassPlan pl;
	pl.w = w; pl.h = h; pl.scale = o->scale; pl.ncols = o->num_columns;
	pl.row0 = o->row_begin; pl.row1 = o->row_end;
	if (pl.row0 == 0 && pl.row1 == 0) pl.row1 = h;
	if (pl.row0 < 0 || pl.row1 > h || pl.row0 >= pl.row1) return fail(RT_ERR_ARG, "bad row band [%d,%d)", pl.row0, pl.row1);
	if (pl.row0 % pl.scale != 0 || (pl.row1 % pl.scale != 0 && pl.row1 != h))
		return fail(RT_ERR_ARG, "row band must be aligned to scale");
	int rc = plan_kernel(pl, o);
	if (rc != RT_OK) return rc;

	size_t bpp = bytes_per_pixel(o->fb_format);
	int band_rows = pl.row1 - pl.row0;
	int fb_row_offset = o->band_only_fb ? pl.row0 : 0;
	size_t fb_rows = o->band_only_fb ? (size_t) band_rows : (size_t) h;

	bool dev_fb = o->fb_memory == RT_MEM_DEVICE || (o->fb_memory == RT_MEM_AUTO && is_device_pointer(fb));
	DeviceCtx &d0 = g.dev[0];
	void *target = fb;
	/* Pipelined host read-back (opts->pipeline): the frame is rendered into one of
	 * two staging buffers and copied to `fb` by the copy stream while the next
	 * call already renders into the other one; the call returns without waiting.
	 * `fb` is valid after rt_cuda_synchronize() (or once two later calls returned). */
	bool pipelined = !dev_fb && o->pipeline && g.ngpu == 1 && !stats;
	/* Pipelined composite into a shared frame (opts->frame_seq): same two staging
	 * frames, but the copy stream ships the owned blocks to peer memory and then
	 * raises this rank's `arrived` flag in the frame's header. */
	bool piped_peer = ship && dev_fb && o->interleave_count > 1 && o->remote_fb && o->frame_seq != 0 && g.ngpu == 1 && !stats;
	int slot = 0;
	if (pipelined || piped_peer) {
		size_t need = fb_rows * (size_t) w * bpp;
		if ((rc = select_device(d0)) != RT_OK) return rc;
		slot = d0.stage_next;
		d0.stage_next ^= 1;
		if (pipelined) CU(cudaEventSynchronize(d0.stage_copied[slot]));       /* the copy that last used this slot */
		if (d0.stage_bytes[slot] < need) {
			CU(cudaFree(d0.stage[slot]));
			d0.stage[slot] = nullptr; d0.stage_bytes[slot] = 0;
			CU(cudaMalloc(&d0.stage[slot], need));
			d0.stage_bytes[slot] = need;
		}
		if (pipelined) target = d0.stage[slot];
	} else
	if (!dev_fb) {
		size_t need = fb_rows * (size_t) w * bpp;
		if ((rc = select_device(d0)) != RT_OK) return rc;
		if (d0.fb_bytes < need) {
			CU(cudaFree(d0.fb));
			d0.fb = nullptr; d0.fb_bytes = 0;
			CU(cudaMalloc(&d0.fb, need));
			d0.fb_bytes = need;
		}
		target = d0.fb;
	}

	int ngpu = g.ngpu;
	bool use_user_stream = ngpu == 1 && o->stream != nullptr;
	int launches = 0;

	/* Several GPUs in this process: every GPU gets the whole call band and renders
	 * the row blocks it owns (round robin), storing straight into GPU 0's frame. */
	int il_n = 1, il_base = 0;
	if (o->interleave_count > 1) {
		if (o->interleave_index < 0 || o->interleave_index >= o->interleave_count)
			return fail(RT_ERR_ARG, "interleave_index out of range");
		if (ngpu > 1) return fail(RT_ERR_ARG, "interleave_count is for one-GPU-per-process ranks; this context already spans %d GPUs", ngpu);
		if (o->interleave_count > RT_MAX_GPUS) return fail(RT_ERR_ARG, "interleave_count > %d", RT_MAX_GPUS);
		il_n = o->interleave_count;
		il_base = o->interleave_index;
	} else if (ngpu > 1)
		il_n = ngpu;
	bool fresh_accum = false;
	for (int i = 0; i < ngpu; i++) {
		DeviceCtx &d = g.dev[i];
		if ((rc = select_device(d)) != RT_OK) return rc;
		if (accumulate && (rc = ensure_accum(d, w, h, pl.row0, band_rows, &fresh_accum)) != RT_OK) return rc;
		if (pl.lbvh) {
			/* the LBVH padding covers ray origins up to d_max away (rt_lbvh.cu) */
			float need = rt_lbvh_required_dmax(&d.bvh, cam->pos);
			if (need > d.bvh.d_max) {
				cudaStream_t st = use_user_stream ? (cudaStream_t) o->stream : d.stream;
				if ((rc = rt_lbvh_refit(&d.bvh, d.geomA, d.geomB, need * 1.05f, st)) != RT_OK)
					return fail(rc, "LBVH refit failed: %s", rt_lbvh_last_error());
			}
		}
	}

	/* banded read-back: one device, host frame, nothing else overlapping already */
	/* a band is worth its launch when it holds ~1.5 M pixels (4K pinned frame: 3.71 / 3.14 / 2.83 / 2.89 / 3.04 ms
	 * with 1 / 2 / 4 / 6 / 8 bands; 1080p: 1.00 / 0.94 / 1.06 ms with 1 / 2 / 4) */
	int nbands = (int) std::min<size_t>((size_t) std::min(g.sync_bands, RT_SYNC_BANDS_MAX), ((size_t) w * (size_t) (pl.row1 - pl.row0)) / 1500000);
	if (g.sync_bands_forced) nbands = std::min(g.sync_bands, RT_SYNC_BANDS_MAX);
	const bool banded = !dev_fb && !pipelined && !piped_peer && ngpu == 1 && o->interleave_count <= 1 && sync_and_copy &&
	                    nbands > 1 && (pl.persistent || pl.queued) && !pl.wavefront &&
	                    band_rows / (pl.scale * RT_TILE_H * 4) >= 2 * nbands;

	/* accumulation weights (main.c:278, 394-396, 476) */
	float wgt = 1.0f / (float) (pl.scale * pl.scale);
	float inv = 1.0f;
	if (accumulate) {
		if (fresh_accum) g.accum_count = 0.0f;      /* buffers were (re)allocated: start over */
		g.accum_count += wgt;
		inv = 1.0f / g.accum_count;
	}

	for (int i = 0; i < ngpu; i++) {
		DeviceCtx &d = g.dev[i];
		if ((rc = select_device(d)) != RT_OK) return rc;
		cudaStream_t st = use_user_stream ? (cudaStream_t) o->stream : d.stream;
		if (stats) {
			CU(cudaMemsetAsync(d.ray_counter, 0, 4 * sizeof(unsigned long long), st));
			CU(cudaEventRecord(d.ev[0], st));
		}
		/* a GPU whose destination frame lives on another GPU renders into its own
		 * memory and ships the blocks it owns afterwards (copy_owned_blocks) */
		bool remote = (ngpu > 1 && i > 0) || (o->interleave_count > 1 && o->remote_fb);
		void *render_to = target;
		if (piped_peer) {
			/* the render may reuse this staging frame once its last copy has left */
			CU(cudaStreamWaitEvent(st, d.stage_copied[slot], 0));
			render_to = d.stage[slot];
		} else if (remote) {
			size_t need = fb_rows * (size_t) w * bpp;
			if (d.fb_bytes < need) {
				CU(cudaFree(d.fb));
				d.fb = nullptr; d.fb_bytes = 0;
				CU(cudaMalloc(&d.fb, need));
				d.fb_bytes = need;
			}
			render_to = d.fb;
		}
		int il_i = ngpu > 1 ? i : il_base;
		if (accumulate && d.accum_needs_clear) {
			CU(cudaMemsetAsync(d.accum, 0, d.accum_bytes, st));
			d.accum_needs_clear = false;
			launches++;
		}
		if (banded) {
			/* Synchronous call with a HOST frame (what INTEGRATION.md's update_frame() binding makes):
			 * the frame is rendered as a few row bands, one launch each, and the copy stream sends band
			 * k to the host while band k+1 renders: 1.9 ms of render + 1.8 ms of PCIe copy per 4K frame
			 * become ~2.6 ms instead of 3.7.  Bands change nothing in the frame (pixels are independent). */
			int unit = pl.scale * RT_TILE_H * 4;                  /* whole tiles, 4 tile rows at least */
			int rows = pl.row1 - pl.row0;
			int per = ((rows + nbands - 1) / nbands + unit - 1) / unit * unit;
			/* every launch first, then the copies: a copy into pageable memory blocks the host, and the
			 * bands behind it must already be queued for the GPU to stay busy meanwhile */
			int k = 0;
			for (int r0 = pl.row0; r0 < pl.row1; r0 += per, k++) {
				int r1 = std::min(r0 + per, pl.row1);
				if (!d.band_ev[k]) CU(cudaEventCreateWithFlags(&d.band_ev[k], cudaEventDisableTiming));
				LaunchExtra X;
				X.no_schedule = true;                               /* one pose, several keys: leave its schedule alone */
				X.no_clear = k > 0;                                 /* the first launch clears for the whole call */
				rc = launch_band(d, cam, pl, o, render_to, fb_row_offset, r0, r1, il_n, il_i, st,
				                 accumulate, wgt, inv, &launches, &X);
				if (rc != RT_OK) return rc;
				CU(cudaEventRecord(d.band_ev[k], st));
			}
			if (stats) CU(cudaEventRecord(d.ev[1], st));
			k = 0;
			for (int r0 = pl.row0; r0 < pl.row1; r0 += per, k++) {
				int r1 = std::min(r0 + per, pl.row1);
				CU(cudaStreamWaitEvent(d.copy_stream, d.band_ev[k], 0));
				if (k == 0 && stats) CU(cudaEventRecord(d.ev[2], d.copy_stream));
				size_t off = (size_t) (r0 - fb_row_offset) * (size_t) w * bpp;
				CU(cudaMemcpyAsync((char *) fb + off, (const char *) render_to + off, (size_t) (r1 - r0) * (size_t) w * bpp,
				                   cudaMemcpyDeviceToHost, d.copy_stream));
			}
			if (stats) CU(cudaEventRecord(d.ev[3], d.copy_stream));
		} else {
		rc = launch_band(d, cam, pl, o, render_to, fb_row_offset, pl.row0, pl.row1, il_n, il_i, st,
		                 accumulate, wgt, inv, &launches);
		if (rc != RT_OK) return rc;
		}
		if (remote && ship && !piped_peer && (rc = copy_owned_blocks(target, render_to, pl, fb_row_offset, il_n, il_i, bpp, st)) != RT_OK) return rc;
		if (stats && !banded) CU(cudaEventRecord(d.ev[1], st));
	}

	if (piped_peer) {
		cudaStream_t st = use_user_stream ? (cudaStream_t) o->stream : d0.stream;
		return ship_piped_peer(d0, pl, *o, fb, fb_row_offset, slot, il_n, il_base, bpp, st);
	}
	if (pipelined) {
		cudaStream_t st = use_user_stream ? (cudaStream_t) o->stream : d0.stream;
		CU(cudaEventRecord(d0.stage_rendered[slot], st));
		CU(cudaStreamWaitEvent(d0.copy_stream, d0.stage_rendered[slot], 0));
		if (o->interleave_count > 1) {
			if ((rc = copy_owned_blocks(fb, d0.stage[slot], pl, fb_row_offset, il_n, il_base, bpp, d0.copy_stream,
			                            cudaMemcpyDeviceToHost)) != RT_OK) return rc;
		} else
			CU(cudaMemcpyAsync(fb, d0.stage[slot], fb_rows * (size_t) w * bpp, cudaMemcpyDeviceToHost, d0.copy_stream));
		CU(cudaEventRecord(d0.stage_copied[slot], d0.copy_stream));
		return RT_OK;
	}
	if (!sync_and_copy && !stats && dev_fb) {
		cudaSetDevice(d0.device);
		return RT_OK;       /* fully asynchronous call */
	}

	/* wait for the bands; band GPUs wrote into GPU 0 memory directly (P2P) */
	float render_ms = 0.0f;
	unsigned long long rays = 0;
	for (int i = 0; i < ngpu; i++) {
		DeviceCtx &d = g.dev[i];
		if ((rc = select_device(d)) != RT_OK) return rc;
		cudaStream_t st = use_user_stream ? (cudaStream_t) o->stream : d.stream;
		if (stats) {
			CU(cudaMemcpyAsync(d.host_rays, d.ray_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
		}
		CU(cudaStreamSynchronize(st));
		if (stats) {
			float ms = 0.0f;
			CU(cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
			render_ms = std::max(render_ms, ms);
			rays += *d.host_rays;
		}
	}
	if ((rc = select_device(d0)) != RT_OK) return rc;

	float copy_ms = 0.0f;
	if (banded) {
		CU(cudaStreamSynchronize(d0.copy_stream));
		if (stats) CU(cudaEventElapsedTime(&copy_ms, d0.ev[2], d0.ev[3]));
	} else
	if (!dev_fb) {
		cudaStream_t st = use_user_stream ? (cudaStream_t) o->stream : d0.stream;
		CU(cudaEventRecord(d0.ev[2], st));
		if (o->interleave_count > 1) {
			/* one-GPU-per-process ranks sharing a HOST frame: every rank copies only the
			 * row blocks it rendered, over its own PCIe link, straight into the frame */
			if ((rc = copy_owned_blocks(fb, d0.fb, pl, fb_row_offset, il_n, il_base, bpp, st, cudaMemcpyDeviceToHost)) != RT_OK) return rc;
		} else
			CU(cudaMemcpyAsync(fb, d0.fb, fb_rows * (size_t) w * bpp, cudaMemcpyDeviceToHost, st));
		CU(cudaEventRecord(d0.ev[3], st));
		CU(cudaStreamSynchronize(st));
		if (stats) CU(cudaEventElapsedTime(&copy_ms, d0.ev[2], d0.ev[3]));
	}
	if (stats) {
		stats->rays = rays;
		int colw = w / pl.ncols;
		int cells_per_row = pl.ncols * ((colw + pl.scale - 1) / pl.scale);
		int lh = h / pl.scale;
		int l0 = pl.row0 / pl.scale, l1 = std::min((pl.row1 + pl.scale - 1) / pl.scale, lh);
		stats->pixels = (uint64_t) cells_per_row * (uint64_t) std::max(l1 - l0, 0);
		stats->render_ms = render_ms;
		stats->composite_ms = 0.0f;    /* fused: band kernels store into GPU 0 over P2P */
		stats->copy_ms = copy_ms;
		stats->kernel_launches = launches;
	}
	return RT_OK;
}

extern "C" int render_frame_cuda_ex(const RtCamera *cam, void *fb, int w, int h, const RtRenderOpts *opts,
                                    RtRenderStats *stats)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	RtRenderOpts def;
	if (!opts) { rt_render_opts_default(&def); opts = &def; }
	if ((rc = validate_common(cam, fb, w, h, opts)) != RT_OK) return rc;
	/* a caller that names a stream and asks for no statistics gets stream-ordered
	 * behaviour for a device frame: the call returns after queueing the launch */
	bool wait = !(opts->stream && !stats && g.ngpu == 1);
	return render_pass(cam, fb, w, h, opts, opts->accumulate != 0, stats, wait);
}

extern "C" int render_frame_cuda(const RtScene *scene, const RtCamera *cam, void *fb, int w, int h, int scale)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (scene) {
		bool same = g.have_scene && g.scene_cache && g.scene_cache->num_objects == scene->num_objects &&
		            memcmp(g.scene_cache->objects, scene->objects, sizeof(RtObject) * (size_t) scene->num_objects) == 0;
		if (!same && (rc = rt_cuda_upload_scene(scene)) != RT_OK) return rc;
	}
	RtRenderOpts o;
	rt_render_opts_default(&o);
	o.scale = scale;
	if ((rc = validate_common(cam, fb, w, h, &o)) != RT_OK) return rc;
	return render_pass(cam, fb, w, h, &o, false, nullptr, true);
}

/* ------------------------------------------------------ concurrent sweep */

/*
 * The passes of a progressive sweep (main.c:354, 402-403: scale 16, 8, 4, 2, 1 after an
 * invalidation) are independent path-tracing jobs -- each has its own RNG key (pass index) and
 * its own pixel grid -- and the coarse ones are tiny: 8 040 paths at scale 16 of a 1080p frame,
 * which keep 1/14 of the GPU busy for as long as their longest path lasts (up to 40 dependent
 * rays, ~0.15 ms).  Run one after the other the five passes take 1.28 ms, of which the three
 * coarsest are 0.45 ms of a nearly idle GPU.  What orders them is only the accumulation
 * (main.c:394: accum = accum * 1 + column_data / scale^2, a binary32 sum taken in pass order).
 *
 * So: every coarse pass runs on its own stream and writes ONE value per low-res cell
 * (RtRenderParams::compact) into its own small buffer, the scale-1 pass writes a plain frame,
 * all five side by side; then one resolve kernel folds them per output pixel in pass order --
 * the same additions main.c:394 performs (weights are powers of two, a pixel a pass does not
 * cover adds nothing: rows >= (H / scale) * scale, main.c:285-290) -- and writes the
 * accumulation buffer and the resolved frame (main.c:476).  Bit-identical to the sequential
 * passes (tests/test_gpu_parity.py: sweeps vs the oracle), which stay as the path for
 * several GPUs, row bands and the other kernels.
 */
struct SweepCoarse {
	const float *cells;       /* cells_per_row x lh float3 */
	int   scale, lh, cells_per_row, cells_per_col;
	float weight;             /* 1.0f / (scale * scale) */
};
struct SweepResolve {
	SweepCoarse coarse[RT_SWEEP_MAX_COARSE];
	int   ncoarse;
	const float *fine;        /* scale-1 pass: W x H float3 */
	float *accum;
	void  *fb;
	int    fb_format;
	int    W, H, column_w, covered_w;
	int    il_n, il_i;        /* this rank owns the blocks b of RT_INTERLEAVE_ROWS output rows with b % il_n == il_i */
	float  inv_count;
};

__global__ void __launch_bounds__(256) sweep_resolve_kernel(const __grid_constant__ SweepResolve R)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
	if (x >= R.W) return;
	if (R.il_n > 1 && (y / RT_INTERLEAVE_ROWS) % R.il_n != R.il_i) return;     /* another rank's rows */
	const size_t p = (size_t) y * R.W + x;
	float ax = 0.0f, ay = 0.0f, az = 0.0f;
	if (x < R.covered_w) {            /* main.c:363: columns beyond T * column_w are never rendered */
		const int col = x / R.column_w, xi = x - col * R.column_w;
		for (int k = 0; k < R.ncoarse; k++) {
			const SweepCoarse &c = R.coarse[k];
			const int j = y / c.scale;
			if (j >= c.lh) continue;                                   /* main.c:285-290 */
			const float *v = c.cells + 3 * ((size_t) j * c.cells_per_row + col * c.cells_per_col + xi / c.scale);
			ax = __fadd_rn(ax, __fmul_rn(v[0], c.weight));             /* main.c:394 combine(accum, data, 1, 1/scale^2) */
			ay = __fadd_rn(ay, __fmul_rn(v[1], c.weight));
			az = __fadd_rn(az, __fmul_rn(v[2], c.weight));
		}
		const float *f = R.fine + 3 * p;
		ax = __fadd_rn(ax, f[0]); ay = __fadd_rn(ay, f[1]); az = __fadd_rn(az, f[2]);
	}
	float *a = R.accum + 3 * p;
	a[0] = ax; a[1] = ay; a[2] = az;
	const float ox = __fmul_rn(ax, R.inv_count), oy = __fmul_rn(ay, R.inv_count), oz = __fmul_rn(az, R.inv_count);   /* main.c:476 */
	if (R.fb_format == RT_FB_F32X3) {
		float *o = reinterpret_cast<float *>(R.fb) + 3 * p;
		o[0] = ox; o[1] = oy; o[2] = oz;
	} else {
		uchar4 q;                                                      /* main.c:666-670 */
		q.x = (unsigned char) __float2uint_rz(__fmul_rn(ox, 255.0f));
		q.y = (unsigned char) __float2uint_rz(__fmul_rn(oy, 255.0f));
		q.z = (unsigned char) __float2uint_rz(__fmul_rn(oz, 255.0f));
		q.w = 255;
		if (x >= R.covered_w) q = make_uchar4(0, 0, 0, 0);             /* never written: the cleared frame (clear_uncovered_owned) */
		reinterpret_cast<uchar4 *>(R.fb)[p] = q;
	}
}

static bool sweep_can_run_concurrently(int w, int h, int init_scale, const RtRenderOpts &o)
{
	if (!g.concurrent_sweep || g.ngpu != 1 || init_scale < 2 || init_scale > 64) return false;
	if (o.pipeline || o.band_only_fb) return false;
	/* ranks of a one-process-per-GPU run: only in the pipelined composite (the sweep's frame goes to a
	 * staging frame and is shipped by the copy stream); every other split takes the passes in turn */
	const bool rank_mode = o.interleave_count > 1;
	if (rank_mode && !(o.remote_fb && o.frame_seq != 0 && init_scale <= RT_INTERLEAVE_ROWS)) return false;
	if (!rank_mode && (o.remote_fb || o.frame_seq)) return false;
	if (!((o.row_begin == 0 && o.row_end == 0) || (o.row_begin == 0 && o.row_end == h))) return false;
	if (o.kernel == RT_KERNEL_WAVEFRONT || o.kernel == RT_KERNEL_PIXEL) return false;
	(void) w;
	return true;
}

/* CTA slots of the resident grid a coarse pass may take, so that the finer passes launched after it
 * still find room: roughly the pass's share of the sweep's pixels, with a floor for the latency-bound
 * coarsest ones.  The scale-1 pass comes last, asks for everything and soaks up what the others free. */
static float sweep_grid_share(int scale)
{
	static float env[3] = {-1.0f, -1.0f, -1.0f};
	if (env[0] < 0.0f) {
		const char *e = getenv("RT_SWEEP_SHARES");      /* "s2,s4,coarser" for experiments */
		if (!e || sscanf(e, "%f,%f,%f", &env[0], &env[1], &env[2]) != 3) { env[0] = 0.5f; env[1] = 0.25f; env[2] = 0.125f; }
	}
	return scale == 2 ? env[0] : (scale == 4 ? env[1] : env[2]);
}

static int sweep_concurrent(const RtCamera *cam, void *fb, int w, int h, int init_scale, uint64_t first_pass,
                            const RtRenderOpts &o, RtRenderStats *stats, bool dev_fb)
{
	DeviceCtx &d = g.dev[0];
	int rc = select_device(d);
	if (rc != RT_OK) return rc;
	cudaStream_t main_st = o.stream ? (cudaStream_t) o.stream : d.stream;
	const size_t bpp = bytes_per_pixel(o.fb_format);
	int launches = 0;
	const bool piped = o.interleave_count > 1;             /* a rank of the pipelined composite (sweep_can_run_concurrently) */
	const int il_n = piped ? o.interleave_count : 1, il_i = piped ? o.interleave_index : 0;
	if (piped && (stats || !dev_fb)) return fail(RT_ERR_ARG, "a pipelined composite (frame_seq) takes a device frame and no statistics");
	if (piped && (il_i < 0 || il_i >= il_n || il_n > RT_MAX_GPUS)) return fail(RT_ERR_ARG, "interleave_index out of range");

	/* the passes */
	int scales[RT_SWEEP_MAX_COARSE + 1], npass = 0;
	for (int sc = init_scale; sc >= 1; sc >>= 1) scales[npass++] = sc;
	const int ncoarse = npass - 1;

	/* buffers: one cell grid per coarse pass, the scale-1 frame, the accumulation, a staging frame for a host fb */
	const int ncols = o.num_columns, column_w = w / ncols;
	size_t cell_off[RT_SWEEP_MAX_COARSE], cells_total = 0;
	for (int k = 0; k < ncoarse; k++) {
		int cpr = ncols * ((column_w + scales[k] - 1) / scales[k]), lh = h / scales[k];
		cell_off[k] = cells_total;
		cells_total += 3 * (size_t) cpr * lh;
	}
	if (d.sweep_cells_floats < cells_total) {
		CU(cudaFree(d.sweep_cells));
		d.sweep_cells = nullptr; d.sweep_cells_floats = 0;
		CU(cudaMalloc(&d.sweep_cells, cells_total * sizeof(float)));
		d.sweep_cells_floats = cells_total;
	}
	const size_t fine_floats = 3 * (size_t) w * h;
	if (d.sweep_fine_floats < fine_floats) {
		CU(cudaFree(d.sweep_fine));
		d.sweep_fine = nullptr; d.sweep_fine_floats = 0;
		CU(cudaMalloc(&d.sweep_fine, fine_floats * sizeof(float)));
		d.sweep_fine_floats = fine_floats;
	}
	bool fresh = false;
	if ((rc = ensure_accum(d, w, h, 0, h, &fresh)) != RT_OK) return rc;
	d.accum_needs_clear = false;                    /* the resolve writes every pixel */
	void *target = fb;
	int slot = 0;
	if (piped) {
		/* the resolved frame goes to one of the two staging frames; the copy stream ships it (render_pass) */
		size_t need = (size_t) w * h * bpp;
		slot = d.stage_next;
		d.stage_next ^= 1;
		if (d.stage_bytes[slot] < need) {
			CU(cudaFree(d.stage[slot]));
			d.stage[slot] = nullptr; d.stage_bytes[slot] = 0;
			CU(cudaMalloc(&d.stage[slot], need));
			d.stage_bytes[slot] = need;
		}
		CU(cudaStreamWaitEvent(main_st, d.stage_copied[slot], 0));
		target = d.stage[slot];
	} else
	if (!dev_fb) {
		size_t need = (size_t) w * h * bpp;
		if (d.fb_bytes < need) {
			CU(cudaFree(d.fb));
			d.fb = nullptr; d.fb_bytes = 0;
			CU(cudaMalloc(&d.fb, need));
			d.fb_bytes = need;
		}
		target = d.fb;
	}
	if (!d.sweep_fork) CU(cudaEventCreateWithFlags(&d.sweep_fork, cudaEventDisableTiming));
	for (int k = 0; k < ncoarse; k++) {
		if (!d.sweep_stream[k]) CU(cudaStreamCreateWithFlags(&d.sweep_stream[k], cudaStreamNonBlocking));
		if (!d.sweep_join[k]) CU(cudaEventCreateWithFlags(&d.sweep_join[k], cudaEventDisableTiming));
	}

	PassPlan pl;
	pl.w = w; pl.h = h; pl.ncols = ncols; pl.row0 = 0; pl.row1 = h;
	pl.scale = 1;
	RtRenderOpts oo = o;
	oo.scale = 1;
	if ((rc = plan_kernel(pl, &oo)) != RT_OK) return rc;
	if (pl.lbvh) {
		float need = rt_lbvh_required_dmax(&d.bvh, cam->pos);
		if (need > d.bvh.d_max && (rc = rt_lbvh_refit(&d.bvh, d.geomA, d.geomB, need * 1.05f, main_st)) != RT_OK)
			return fail(rc, "LBVH refit failed: %s", rt_lbvh_last_error());
	}

	if (stats) {
		CU(cudaMemsetAsync(d.ray_counter, 0, 4 * sizeof(unsigned long long), main_st));
		CU(cudaEventRecord(d.ev[0], main_st));
	}
	CU(cudaEventRecord(d.sweep_fork, main_st));

	SweepResolve R;
	memset(&R, 0, sizeof(R));
	float count = 0.0f;
	uint64_t pixels = 0;
	for (int k = 0; k < npass; k++) {
		const int sc = scales[k];
		const bool fine = sc == 1;
		const float wgt = 1.0f / (float) (sc * sc);           /* main.c:278 */
		count += wgt;                                          /* main.c:395 */
		pl.scale = sc;
		oo = o;
		oo.scale = sc;
		oo.pass_index = first_pass + (uint64_t) k;
		oo.fb_format = RT_FB_F32X3;
		oo.accumulate = 0;
		LaunchExtra X;
		X.no_schedule = true;
		X.no_dense = true;            /* measured: the 1080p sweep takes 0.98 ms with the scale-1 pass on the 7-CTA build, 0.83 ms without */
		X.compact = !fine;
		X.grid_share = fine ? 1.0f : sweep_grid_share(sc);
		const int cpc = (column_w + sc - 1) / sc, cpr = ncols * cpc, lh = h / sc;
		pixels += (uint64_t) cpr * (uint64_t) lh;
		cudaStream_t st = fine ? main_st : d.sweep_stream[k];
		void *out = fine ? (void *) d.sweep_fine : (void *) (d.sweep_cells + cell_off[k]);
		if (!fine) CU(cudaStreamWaitEvent(st, d.sweep_fork, 0));
		rc = launch_band(d, cam, pl, &oo, out, 0, 0, h, il_n, il_i, st, false, wgt, 1.0f, &launches, &X);
		if (rc != RT_OK) return rc;
		if (!fine) {
			CU(cudaEventRecord(d.sweep_join[k], st));
			SweepCoarse &c = R.coarse[k];
			c.cells = d.sweep_cells + cell_off[k];
			c.scale = sc; c.lh = lh; c.cells_per_row = cpr; c.cells_per_col = cpc; c.weight = wgt;
		}
	}
	for (int k = 0; k < ncoarse; k++) CU(cudaStreamWaitEvent(main_st, d.sweep_join[k], 0));
	R.ncoarse = ncoarse;
	R.fine = d.sweep_fine;
	R.accum = d.accum;
	R.fb = target;
	R.fb_format = o.fb_format;
	R.W = w; R.H = h; R.column_w = column_w; R.covered_w = column_w * ncols;
	R.il_n = il_n; R.il_i = il_i;
	g.accum_count = count;
	R.inv_count = 1.0f / count;                                /* main.c:476 */
	sweep_resolve_kernel<<<dim3((unsigned) ((w + 255) / 256), (unsigned) h), 256, 0, main_st>>>(R);
	CU(cudaGetLastError());
	launches++;
	if (stats) CU(cudaEventRecord(d.ev[1], main_st));
	if (piped) {
		pl.scale = 1;
		return ship_piped_peer(d, pl, o, fb, 0, slot, il_n, il_i, bpp, main_st);
	}

	/* with a caller stream, a device frame and no statistics the whole sweep is stream-ordered */
	if (dev_fb && o.stream && !stats) return RT_OK;
	if (stats) CU(cudaMemcpyAsync(d.host_rays, d.ray_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, main_st));
	float copy_ms = 0.0f;
	if (!dev_fb) {
		CU(cudaEventRecord(d.ev[2], main_st));
		CU(cudaMemcpyAsync(fb, d.fb, (size_t) w * h * bpp, cudaMemcpyDeviceToHost, main_st));
		CU(cudaEventRecord(d.ev[3], main_st));
	}
	CU(cudaStreamSynchronize(main_st));
	if (stats) {
		memset(stats, 0, sizeof(*stats));
		float ms = 0.0f;
		CU(cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
		if (!dev_fb) CU(cudaEventElapsedTime(&copy_ms, d.ev[2], d.ev[3]));
		stats->rays = *d.host_rays;
		stats->pixels = pixels;
		stats->render_ms = ms;
		stats->copy_ms = copy_ms;
		stats->kernel_launches = launches;
	}
	return RT_OK;
}

extern "C" int rt_cuda_render_sweep(const RtCamera *cam, void *fb, int w, int h, int init_scale,
                                    uint64_t first_pass, const RtRenderOpts *opts, RtRenderStats *stats)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (init_scale < 1 || (init_scale & (init_scale - 1))) return fail(RT_ERR_ARG, "init_scale must be a power of two");
	RtRenderOpts o;
	if (opts) o = *opts; else rt_render_opts_default(&o);
	/* rows are dealt to GPUs / ranks in blocks of RT_INTERLEAVE_ROWS output rows so that a GPU owns the
	 * same pixels (and its accumulation buffer stays valid) at every scale of the sweep: that holds
	 * for scales up to the block size, which is also the reference's largest --init-scale (main.c:598) */
	if (init_scale > RT_INTERLEAVE_ROWS && (g.ngpu > 1 || o.interleave_count > 1))
		return fail(RT_ERR_ARG, "init_scale %d > %d is not supported when the frame is split over GPUs", init_scale, RT_INTERLEAVE_ROWS);
	o.scale = init_scale;
	if ((rc = validate_common(cam, fb, w, h, &o)) != RT_OK) return rc;
	if ((rc = rt_cuda_accum_reset()) != RT_OK) return rc;          /* invalidate_accumulation() */
	RtRenderStats total;
	memset(&total, 0, sizeof(total));
	uint64_t pass = first_pass;
	bool dev_fb = o.fb_memory == RT_MEM_DEVICE || (o.fb_memory == RT_MEM_AUTO && is_device_pointer(fb));
	if (sweep_can_run_concurrently(w, h, init_scale, o)) return sweep_concurrent(cam, fb, w, h, init_scale, first_pass, o, stats, dev_fb);
	for (int s = init_scale; s >= 1; s >>= 1, pass++) {             /* main.c:402-403 */
		o.scale = s;
		o.pass_index = pass;
		RtRenderStats st;
		memset(&st, 0, sizeof(st));
		bool last = s == 1;
		/* intermediate passes of a host-destination sweep stay on the device */
		RtRenderOpts oo = o;
		void *dst = fb;
		if (!dev_fb && !last) {
			size_t need = (size_t) w * h * bytes_per_pixel(o.fb_format);
			DeviceCtx &d0 = g.dev[0];
			if ((rc = select_device(d0)) != RT_OK) return rc;
			if (d0.fb_bytes < need) {
				CU(cudaFree(d0.fb));
				d0.fb = nullptr; d0.fb_bytes = 0;
				CU(cudaMalloc(&d0.fb, need));
				d0.fb_bytes = need;
			}
			dst = d0.fb;
			oo.fb_memory = RT_MEM_DEVICE;
		}
		/* with a caller stream, a device frame and no statistics the whole sweep is stream-ordered */
		bool wait = last && !(dev_fb && o.stream && !stats && g.ngpu == 1);
		rc = render_pass(cam, dst, w, h, &oo, true, stats ? &st : nullptr, wait, last || !oo.remote_fb);
		if (rc != RT_OK) return rc;
		total.rays += st.rays; total.pixels += st.pixels;
		total.render_ms += st.render_ms; total.copy_ms += st.copy_ms;
		total.kernel_launches += st.kernel_launches;
	}
	if (stats) *stats = total;
	return RT_OK;
}

/* ------------------------------------------------------------------ probes */

struct TempBuf {
	void *p = nullptr;
	~TempBuf() { cudaFree(p); }
	cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 4); }
};

extern "C" int rt_cuda_debug_trace(const float *rays6, int n, float *out7, int32_t *obj, int variant, int traversal)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!g.have_scene) return fail(RT_ERR_STATE, "no scene uploaded");
	bool lbvh = false;
	if ((rc = pick_traversal(traversal, &lbvh)) != RT_OK) return rc;
	DeviceCtx &d = g.dev[0];
	if ((rc = select_device(d)) != RT_OK) return rc;
	TempBuf r, o, ob;
	CU(r.alloc(sizeof(float) * 6 * (size_t) n));
	CU(o.alloc(sizeof(float) * 7 * (size_t) n));
	CU(ob.alloc(sizeof(int) * (size_t) n));
	CU(cudaMemcpyAsync(r.p, rays6, sizeof(float) * 6 * (size_t) n, cudaMemcpyHostToDevice, d.stream));
	RtRenderParams P;
	memset(&P, 0, sizeof(P));
	fill_views(d, P);
	CU((variant == RT_VARIANT_FAST ? rt_fast_launch_probe_trace : rt_exact_launch_probe_trace)(
	    &P, lbvh, (const float *) r.p, n, (float *) o.p, (int *) ob.p, d.stream));
	CU(cudaMemcpyAsync(out7, o.p, sizeof(float) * 7 * (size_t) n, cudaMemcpyDeviceToHost, d.stream));
	CU(cudaMemcpyAsync(obj, ob.p, sizeof(int) * (size_t) n, cudaMemcpyDeviceToHost, d.stream));
	CU(cudaStreamSynchronize(d.stream));
	return RT_OK;
}

extern "C" int rt_cuda_debug_sample_cubemap(const float *dirs3, int n, float *out3)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!g.have_sky) return fail(RT_ERR_STATE, "no skybox uploaded");
	DeviceCtx &d = g.dev[0];
	if ((rc = select_device(d)) != RT_OK) return rc;
	TempBuf in, out;
	CU(in.alloc(sizeof(float) * 3 * (size_t) n));
	CU(out.alloc(sizeof(float) * 3 * (size_t) n));
	CU(cudaMemcpyAsync(in.p, dirs3, sizeof(float) * 3 * (size_t) n, cudaMemcpyHostToDevice, d.stream));
	RtRenderParams P;
	memset(&P, 0, sizeof(P));
	fill_views(d, P);
	CU(rt_exact_launch_probe_sky(&P.sky, d.lut, (const float *) in.p, n, (float *) out.p, d.stream));
	CU(cudaMemcpyAsync(out3, out.p, sizeof(float) * 3 * (size_t) n, cudaMemcpyDeviceToHost, d.stream));
	CU(cudaStreamSynchronize(d.stream));
	return RT_OK;
}

extern "C" int rt_cuda_debug_camera_rays(const RtCamera *cam, const float *pxpy, int n, float aspect, float *rays6)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	DeviceCtx &d = g.dev[0];
	if ((rc = select_device(d)) != RT_OK) return rc;
	TempBuf in, out;
	CU(in.alloc(sizeof(float) * 2 * (size_t) n));
	CU(out.alloc(sizeof(float) * 6 * (size_t) n));
	CU(cudaMemcpyAsync(in.p, pxpy, sizeof(float) * 2 * (size_t) n, cudaMemcpyHostToDevice, d.stream));
	RtCameraFrame cf;
	rt_host_camera_frame(cam, aspect, &cf);
	CU(rt_exact_launch_probe_camera(&cf, (const float *) in.p, n, (float *) out.p, d.stream));
	CU(cudaMemcpyAsync(rays6, out.p, sizeof(float) * 6 * (size_t) n, cudaMemcpyDeviceToHost, d.stream));
	CU(cudaStreamSynchronize(d.stream));
	return RT_OK;
}

static int rng_probe(uint64_t state, int n, uint64_t *u64_out, float *f32_out, float *dir_out)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	DeviceCtx &d = g.dev[0];
	if ((rc = select_device(d)) != RT_OK) return rc;
	TempBuf u, f, v;
	CU(u.alloc(sizeof(uint64_t) * (size_t) n));
	CU(f.alloc(sizeof(float) * (size_t) n));
	CU(v.alloc(sizeof(float) * 3 * (size_t) n));
	CU(rt_exact_launch_probe_rng(state, n, u64_out ? (uint64_t *) u.p : nullptr, f32_out ? (float *) f.p : nullptr,
	                             dir_out ? (float *) v.p : nullptr, d.stream));
	if (u64_out) CU(cudaMemcpyAsync(u64_out, u.p, sizeof(uint64_t) * (size_t) n, cudaMemcpyDeviceToHost, d.stream));
	if (f32_out) CU(cudaMemcpyAsync(f32_out, f.p, sizeof(float) * (size_t) n, cudaMemcpyDeviceToHost, d.stream));
	if (dir_out) CU(cudaMemcpyAsync(dir_out, v.p, sizeof(float) * 3 * (size_t) n, cudaMemcpyDeviceToHost, d.stream));
	CU(cudaStreamSynchronize(d.stream));
	return RT_OK;
}

extern "C" int rt_cuda_debug_rng(uint64_t state, int n, uint64_t *u64_out, float *f32_out)
{
	return rng_probe(state, n, u64_out, f32_out, nullptr);
}

extern "C" int rt_cuda_debug_random_directions(uint64_t state, int n, float *out3)
{
	return rng_probe(state, n, nullptr, nullptr, out3);
}

/* ------------------------------------------------------ FP32 peak probe */

/* Register-only FMA (or MUL+ADD) chains: the measured FP32 issue-rate ceiling
 * the roofline fraction is quoted against.  8 independent
 * chains per thread, 2 flops per FMA / per MUL+ADD pair. */
template <bool FMA>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float a, float b)
{
	float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f;
	float x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int k = 0; k < 16; k++) {
			if (FMA) {
				x0 = __fmaf_rn(x0, a, b); x1 = __fmaf_rn(x1, a, b); x2 = __fmaf_rn(x2, a, b); x3 = __fmaf_rn(x3, a, b);
				x4 = __fmaf_rn(x4, a, b); x5 = __fmaf_rn(x5, a, b); x6 = __fmaf_rn(x6, a, b); x7 = __fmaf_rn(x7, a, b);
			} else {
				x0 = __fadd_rn(__fmul_rn(x0, a), b); x1 = __fadd_rn(__fmul_rn(x1, a), b);
				x2 = __fadd_rn(__fmul_rn(x2, a), b); x3 = __fadd_rn(__fmul_rn(x3, a), b);
				x4 = __fadd_rn(__fmul_rn(x4, a), b); x5 = __fadd_rn(__fmul_rn(x5, a), b);
				x6 = __fadd_rn(__fmul_rn(x6, a), b); x7 = __fadd_rn(__fmul_rn(x7, a), b);
			}
		}
	}
	float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
	if (s == 12345.678f) out[0] = s;    /* keep the chains alive */
}

/* Returns measured TFLOP/s (2 flops per FMA, or per MUL+ADD pair when fma == 0). */
extern "C" int rt_cuda_debug_fp32_peak(int fma, float *tflops_out)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	DeviceCtx &d = g.dev[0];
	if ((rc = select_device(d)) != RT_OK) return rc;
	TempBuf out;
	CU(out.alloc(sizeof(float)));
	const int iters = 2048, threads = 256;
	int blocks = d.sm_count * 8;
	float best = 0.0f;
	for (int rep = 0; rep < 5; rep++) {
		CU(cudaEventRecord(d.ev[0], d.stream));
		if (fma) fp32_peak_kernel<true><<<blocks, threads, 0, d.stream>>>((float *) out.p, iters, 0.999f, 0.001f);
		else     fp32_peak_kernel<false><<<blocks, threads, 0, d.stream>>>((float *) out.p, iters, 0.999f, 0.001f);
		CU(cudaGetLastError());
		CU(cudaEventRecord(d.ev[1], d.stream));
		CU(cudaStreamSynchronize(d.stream));
		float ms = 0.0f;
		CU(cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
		double flops = 2.0 * 8.0 * 16.0 * (double) iters * (double) threads * (double) blocks;
		float tf = (float) (flops / (ms * 1e-3) / 1e12);
		if (rep > 0 && tf > best) best = tf;
	}
	*tflops_out = best;
	return RT_OK;
}

/* Bit-compare the hoisted division (rt_device.cuh: div_hoisted) with the IEEE
 * `/` on blocks*256*per_thread pseudo-random operand pairs whose exponents lie
 * in [lo_b, hi_b] (divisor) and [lo_a, hi_a] (dividend). */
extern "C" int rt_cuda_debug_div_check(uint64_t seed, unsigned blocks, unsigned per_thread, int lo_b, int hi_b,
                                       int lo_a, int hi_a, uint64_t *mismatches)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	DeviceCtx &d = g.dev[0];
	if ((rc = select_device(d)) != RT_OK) return rc;
	TempBuf cnt;
	CU(cnt.alloc(sizeof(unsigned long long)));
	CU(cudaMemsetAsync(cnt.p, 0, sizeof(unsigned long long), d.stream));
	CU(rt_exact_launch_probe_div(seed, blocks, per_thread, lo_b, hi_b, lo_a, hi_a, (unsigned long long *) cnt.p, d.stream));
	unsigned long long h = 0;
	CU(cudaMemcpyAsync(&h, cnt.p, sizeof(h), cudaMemcpyDeviceToHost, d.stream));
	CU(cudaStreamSynchronize(d.stream));
	*mismatches = h;
	return RT_OK;
}

/* ------------------------------------------- cross-process frame sharing */

/*
 * One-process-per-GPU composite without a collective: rank 0 allocates the
 * frame with rt_cuda_shared_frame_create() and hands the 64-byte handle to the
 * other ranks (any transport; bench.py uses torch.distributed); they map it
 * with rt_cuda_shared_frame_open() and pass the mapped address as `fb` with
 * their interleave index and remote_fb = 1: each rank renders the row blocks it
 * owns locally and copy_owned_blocks() ships them into GPU 0's memory over
 * NVLink with one strided peer copy.
 */
extern "C" int rt_cuda_shared_frame_create(size_t bytes, void **dev_ptr, void *handle64)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!dev_ptr || !handle64 || bytes == 0) return fail(RT_ERR_ARG, "bad shared frame request");
	if ((rc = select_device(g.dev[0])) != RT_OK) return rc;
	void *p = nullptr;
	CU(cudaMalloc(&p, bytes + RT_SHARED_HEADER_BYTES));
	cudaError_t e = cudaMemset(p, 0, RT_SHARED_HEADER_BYTES);
	cudaIpcMemHandle_t h;
	if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
	if (e != cudaSuccess) {
		cudaFree(p);
		return fail(RT_ERR_CUDA, "shared frame (cudaIpcGetMemHandle): %s", cudaGetErrorString(e));
	}
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	static_assert(sizeof(SharedHeader) <= RT_SHARED_HEADER_BYTES, "header size");
	memcpy(handle64, &h, 64);
	*dev_ptr = (char *) p + RT_SHARED_HEADER_BYTES;
	return RT_OK;
}

extern "C" int rt_cuda_shared_frame_open(const void *handle64, void **dev_ptr)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!dev_ptr || !handle64) return fail(RT_ERR_ARG, "bad shared frame handle");
	if ((rc = select_device(g.dev[0])) != RT_OK) return rc;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, 64);
	void *p = nullptr;
	CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	*dev_ptr = (char *) p + RT_SHARED_HEADER_BYTES;
	return RT_OK;
}

extern "C" int rt_cuda_shared_frame_close(void *dev_ptr, int owner)
{
	if (!dev_ptr) return RT_OK;
	void *base = (char *) dev_ptr - RT_SHARED_HEADER_BYTES;
	if (owner) CU(cudaFree(base));
	else CU(cudaIpcCloseMemHandle(base));
	return RT_OK;
}

extern "C" int rt_cuda_shared_frame_wait(void *dev_ptr, int num_ranks, uint32_t seq, void *stream)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!dev_ptr || num_ranks < 1 || num_ranks > RT_MAX_GPUS) return fail(RT_ERR_ARG, "bad shared frame wait");
	if ((rc = select_device(g.dev[0])) != RT_OK) return rc;
	cudaStream_t st = stream ? (cudaStream_t) stream : g.dev[0].stream;
	flag_wait_kernel<<<1, 32, 0, st>>>(header_of(dev_ptr), num_ranks, seq);
	CU(cudaGetLastError());
	return RT_OK;
}

extern "C" int rt_cuda_shared_frame_release(void *dev_ptr, uint32_t seq, void *stream)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!dev_ptr) return fail(RT_ERR_ARG, "bad shared frame");
	if ((rc = select_device(g.dev[0])) != RT_OK) return rc;
	cudaStream_t st = stream ? (cudaStream_t) stream : g.dev[0].stream;
	flag_release_kernel<<<1, 1, 0, st>>>(header_of(dev_ptr), seq);
	CU(cudaGetLastError());
	return RT_OK;
}

extern "C" int rt_cuda_shared_frame_error(void *dev_ptr, uint32_t *error_out)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!dev_ptr || !error_out) return fail(RT_ERR_ARG, "bad shared frame");
	CU(cudaMemcpy(error_out, &header_of(dev_ptr)->error, sizeof(uint32_t), cudaMemcpyDeviceToHost));
	return RT_OK;
}

/* Plain device->host copy of a library-owned frame (rank 0's read-back). */
extern "C" int rt_cuda_copy_to_host(void *host_dst, const void *dev_src, size_t bytes, void *stream)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	cudaStream_t st = stream ? (cudaStream_t) stream : g.dev[0].stream;
	CU(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return RT_OK;
}

/* Stream-ordered copy between any two addresses of the unified address space
 * (a consumer draining the shared frame on its own stream). */
extern "C" int rt_cuda_copy_async(void *dst, const void *src, size_t bytes, void *stream)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	cudaStream_t st = stream ? (cudaStream_t) stream : g.dev[0].stream;
	CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
	return RT_OK;
}

/* Rays traced on GPU 0 since the last call that returned statistics (those calls reset the
 * counter): lets a caller count the rays of a run of asynchronous launches exactly. */
extern "C" int rt_cuda_ray_counter(uint64_t *rays)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!rays) return fail(RT_ERR_ARG, "rays is NULL");
	DeviceCtx &d = g.dev[0];
	if ((rc = select_device(d)) != RT_OK) return rc;
	CU(cudaDeviceSynchronize());
	unsigned long long h = 0;
	CU(cudaMemcpy(&h, d.ray_counter, sizeof(h), cudaMemcpyDeviceToHost));
	*rays = h;
	return RT_OK;
}

/* Unit probe of the tile-order kernel: the order of `tiles_x * tiles_y` tiles for a cost map of
 * `cost_tiles_x * cost_tiles_y` tiles that is `1 << shift` times coarser. */
extern "C" int rt_cuda_debug_tile_order(const uint32_t *cost, int cost_tiles_x, int cost_tiles_y, int shift,
                                        int tiles_x, int tiles_y, uint32_t *order_out)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	DeviceCtx &d = g.dev[0];
	if ((rc = select_device(d)) != RT_OK) return rc;
	size_t n = (size_t) tiles_x * tiles_y, nc = (size_t) cost_tiles_x * cost_tiles_y;
	if (n == 0 || n >= (1u << 20) || nc == 0 || shift < 0 || shift > 4) return fail(RT_ERR_ARG, "bad tile order request");
	TempBuf c, o, h;
	CU(c.alloc(nc * sizeof(unsigned)));
	CU(o.alloc(n * sizeof(unsigned)));
	CU(h.alloc(RT_ORDER_MAX_CTAS * 16 * sizeof(unsigned)));
	CU(cudaMemcpyAsync(c.p, cost, nc * sizeof(unsigned), cudaMemcpyHostToDevice, d.stream));
	CU(cudaMemsetAsync(o.p, 0xff, n * sizeof(unsigned), d.stream));
	CU(launch_tile_order((const unsigned *) c.p, (unsigned *) o.p, n, tiles_x, cost_tiles_x, cost_tiles_y, shift, (unsigned *) h.p, d.stream));
	CU(cudaMemcpyAsync(order_out, o.p, n * sizeof(unsigned), cudaMemcpyDeviceToHost, d.stream));
	CU(cudaStreamSynchronize(d.stream));
	return RT_OK;
}

/* Internal LBVH nodes visited and primitives tested by the last call that asked for
 * statistics (GPU 0).  Zeros unless the library was built with -DRT_COUNT_WALK
 * (tools/build_variant.sh): the counters cost registers in the hot loop. */
extern "C" int rt_cuda_debug_walk_counts(uint64_t *nodes, uint64_t *tests)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	DeviceCtx &d = g.dev[0];
	if ((rc = select_device(d)) != RT_OK) return rc;
	unsigned long long h[4] = {0, 0, 0, 0};
	CU(cudaMemcpy(h, d.ray_counter, sizeof(h), cudaMemcpyDeviceToHost));
	if (nodes) *nodes = h[1];
	if (tests) *tests = h[2];
	return RT_OK;
}

/* Test knob: threshold of the sign shortcut in the light-sample sweep
 * (rt_device.cuh: sample_faces_surface).  A huge value forces the literal
 * normalise-then-dot path for every sample; results must not change. */
extern "C" int rt_cuda_debug_set_sweep_threshold(float tau2)
{
	g.sweep_tau2 = tau2 >= 0.0f ? tau2 : 4e-12f;
	return RT_OK;
}

/* Test / A-B knob: 0 keeps the queued kernel's tiles in image order (no cost
 * recording, no reordering), 1 (default) lets repeated poses be scheduled
 * longest tiles first.  Results never depend on it. */
extern "C" int rt_cuda_debug_set_tile_schedule(int on)
{
	g.tile_schedule = on == 2 ? 2 : (on ? 1 : 0);
	for (int i = 0; i < g.ngpu; i++) g.dev[i].sched.have_key = false;
	return RT_OK;
}

/* Test / A-B knob: 0 = rt_cuda_render_sweep() runs its passes one after the other. */
extern "C" int rt_cuda_debug_set_concurrent_sweep(int on)
{
	g.concurrent_sweep = on ? 1 : 0;
	return RT_OK;
}

/* Test / A-B knob: 0 = big launches keep the queued kernel's 6-CTA build. */
extern "C" int rt_cuda_debug_set_queued_dense(int on)
{
	g.queued_dense = on ? 1 : 0;
	return RT_OK;
}

/* Test / A-B knob: row bands of a synchronous call with a host frame (1 = render, then copy). */
extern "C" int rt_cuda_debug_set_sync_bands(int bands)
{
	g.sync_bands_forced = bands < 0;            /* negative: exactly -bands bands whatever the frame size (tests) */
	if (bands < 0) bands = -bands;
	g.sync_bands = bands < 1 ? 1 : (bands > RT_SYNC_BANDS_MAX ? RT_SYNC_BANDS_MAX : bands);
	return RT_OK;
}

extern "C" void rt_lbvh_set_builder(int builder);
extern "C" int rt_cuda_set_bvh_builder(int builder)
{
	if (builder != RT_BVH_BUILDER_SAH && builder != RT_BVH_BUILDER_LBVH) return fail(RT_ERR_ARG, "unknown BVH builder %d", builder);
	rt_lbvh_set_builder(builder);
	return RT_OK;
}

extern "C" void rt_lbvh_debug_set_anyhit(int on);
extern "C" int rt_cuda_debug_set_light_anyhit(int on)
{
	rt_lbvh_debug_set_anyhit(on);
	return RT_OK;
}

/* size of the kernel-argument block that goes host -> device with every launch */
extern "C" size_t rt_cuda_param_bytes(void) { return sizeof(RtRenderParams); }

/* ------------------------------------------------ interactive frame loop */

/*
 * The reference's frame scheduler in three calls: workers start
 * at --init-scale and halve the scale after every published pass
 * (main.c:354, 402-403); any camera/window change bumps the generation counter,
 * zeroes accum and sends them back to init_scale (main.c:115-124, 405-408);
 * update_frame() shows accum / count (main.c:450-482).
 */
static struct {
	int      init_scale = 8;            /* main.c:589 default */
	int      num_columns = 1;
	int      scale = 8;
	uint64_t pass = 0;
	uint32_t generation = 0;            /* accum_generation, main.c:59 */
	int      w = 0, h = 0;              /* frame size of the last update_frame() */
	float    ladder_ms = 0.0f;          /* device time of the last concurrent init_scale -> 1 ladder at this size */
} g_loop;

extern "C" int rt_cuda_set_progressive(int init_scale, int num_columns)
{
	if (init_scale != 1 && init_scale != 2 && init_scale != 4 && init_scale != 8 && init_scale != 16)
		return fail(RT_ERR_ARG, "init_scale must be a power of 2 between 1 and 16 (main.c:598)");
	if (num_columns < 1) return fail(RT_ERR_ARG, "num_columns must be >= 1");
	g_loop.init_scale = init_scale;
	g_loop.num_columns = num_columns > 32 ? 32 : num_columns;          /* MAX_COLUMNS, main.c:632 */
	g_loop.scale = init_scale;
	return RT_OK;
}

extern "C" int rt_cuda_invalidate_accumulation(void)
{
	g_loop.scale = g_loop.init_scale;
	g_loop.generation++;
	return rt_cuda_accum_reset();
}

extern "C" uint32_t rt_cuda_accum_generation(void) { return g_loop.generation; }
/* the pass index (RNG key) the next pass of rt_cuda_update_frame() will use; it keeps counting across invalidations */
extern "C" uint64_t rt_cuda_next_pass_index(void) { return g_loop.pass; }

/* One update_frame(): at least one pass at the current scale, then further
 * passes (scale halving down to 1, then more scale-1 samples) while the time
 * already spent stays below budget_ms -- what the reference's free-running
 * workers achieve between two redraws.  fb receives accum / count. */
extern "C" int rt_cuda_update_frame(const RtCamera *cam, void *fb, int w, int h, double budget_ms,
                                    const RtRenderOpts *opts, RtRenderStats *stats)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	RtRenderOpts o;
	if (opts) o = *opts; else rt_render_opts_default(&o);
	/* a resized frame restarts the progressive sweep: realloc_frame_buffer() bumps
	 * accum_generation and the workers go back to init_scale (main.c:416-444, 405-408) */
	if ((g_loop.w || g_loop.h) && (g_loop.w != w || g_loop.h != h) && (rc = rt_cuda_invalidate_accumulation()) != RT_OK) return rc;
	g_loop.w = w; g_loop.h = h;
	o.num_columns = g_loop.num_columns;
	o.scale = g_loop.scale;
	if ((rc = validate_common(cam, fb, w, h, &o)) != RT_OK) return rc;
	bool dev_fb = o.fb_memory == RT_MEM_DEVICE || (o.fb_memory == RT_MEM_AUTO && is_device_pointer(fb));
	RtRenderStats total;
	memset(&total, 0, sizeof(total));
	DeviceCtx &d0 = g.dev[0];
	cudaEvent_t t0 = d0.ev[2], t1 = d0.ev[3];
	if ((rc = select_device(d0)) != RT_OK) return rc;
	cudaStream_t st0 = (g.ngpu == 1 && o.stream) ? (cudaStream_t) o.stream : d0.stream;
	CU(cudaEventRecord(t0, st0));
	/* budget_ms < 0: refine the pose to full resolution -- every pass of the ladder down to scale 1
	 * (exactly one pass when it is there already) -- and return */
	const bool to_full = budget_ms < 0.0;
	/* A fresh pose (right after an invalidation) whose whole init_scale -> 1 ladder fits the budget:
	 * the passes run side by side (sweep_concurrent: 0.83 ms instead of 1.26 ms at 1080p, and one
	 * synchronisation instead of five), with the same pass indices and the same accumulation. */
	if (budget_ms != 0.0 && g_loop.scale == g_loop.init_scale && g_loop.scale >= 2 && g.accum_count == 0.0f) {
		RtRenderOpts so = o;
		so.scale = g_loop.init_scale;
		double guess = g_loop.ladder_ms > 0.0f ? 1.1 * g_loop.ladder_ms : (double) w * h / 2.5e6;
		if (sweep_can_run_concurrently(w, h, g_loop.init_scale, so) && so.interleave_count <= 1 && (to_full || budget_ms >= guess)) {
			RtRenderStats st;
			memset(&st, 0, sizeof(st));
			g.accum_count = 0.0f;
			rc = sweep_concurrent(cam, fb, w, h, g_loop.init_scale, g_loop.pass, so, &st, dev_fb);
			if (rc != RT_OK) return rc;
			for (int sc = g_loop.init_scale; sc >= 1; sc >>= 1) g_loop.pass++;
			g_loop.scale = 1;
			g_loop.ladder_ms = st.render_ms;
			total = st;
			/* more passes at scale 1 while the budget lasts (about 60 % of the ladder's time each) */
			if (to_full || st.render_ms + 0.6 * st.render_ms > budget_ms) {
				if (stats) *stats = total;
				return RT_OK;
			}
		}
	}
	for (int passes = 0;; passes++) {
		o.scale = g_loop.scale;
		o.pass_index = g_loop.pass++;
		RtRenderStats st;
		memset(&st, 0, sizeof(st));
		/* decide after this pass whether another one fits: needs the elapsed device time */
		RtRenderOpts oo = o;
		void *dst = fb;
		bool last_possible = budget_ms == 0.0 || (to_full && g_loop.scale == 1);
		if (!dev_fb && !last_possible) {
			/* keep intermediate passes on the device; the final one copies to the host */
			size_t need = (size_t) w * h * bytes_per_pixel(o.fb_format);
			if (d0.fb_bytes < need) {
				CU(cudaFree(d0.fb));
				d0.fb = nullptr; d0.fb_bytes = 0;
				CU(cudaMalloc(&d0.fb, need));
				d0.fb_bytes = need;
			}
			dst = d0.fb;
			oo.fb_memory = RT_MEM_DEVICE;
		}
		rc = render_pass(cam, dst, w, h, &oo, true, &st, true);
		if (rc != RT_OK) return rc;
		total.rays += st.rays; total.pixels += st.pixels; total.render_ms += st.render_ms;
		total.kernel_launches += st.kernel_launches;
		if (g_loop.scale > 1) g_loop.scale >>= 1;                      /* main.c:402-403 */
		if (last_possible) break;
		if (to_full) continue;
		if ((rc = select_device(d0)) != RT_OK) return rc;
		CU(cudaEventRecord(t1, st0));
		CU(cudaEventSynchronize(t1));
		float ms = 0.0f;
		CU(cudaEventElapsedTime(&ms, t0, t1));
		/* another pass costs about what the last one did (x4 while the scale still halves) */
		double next = st.render_ms * (o.scale > 1 ? 4.0 : 1.0);
		if (ms + next > budget_ms) {
			if (!dev_fb) {
				CU(cudaMemcpyAsync(fb, d0.fb, (size_t) w * h * bytes_per_pixel(o.fb_format), cudaMemcpyDeviceToHost, st0));
				CU(cudaStreamSynchronize(st0));
			}
			break;
		}
	}
	if (stats) *stats = total;
	return RT_OK;
}

/* ------------------------------------------------ CUDA-OpenGL presenter */

/*
 * the reference shows a frame with
 *     glTexImage2D(GL_TEXTURE_2D, 0, GL_RGB, w, h, 0, GL_RGB, GL_FLOAT, data)
 * from a host Vector3 array (gpu_and_windowing.c:371-376), i.e. a device->host
 * copy here plus a host->device upload there per displayed frame.  With a
 * registered pixel-unpack buffer the render kernels store straight into GL-owned
 * device memory.  The two cudart entry points are declared here instead of
 * including <cuda_gl_interop.h>, which wants a system GL header this library has
 * no other use for (GLuint is `unsigned int` on every platform GL runs on).
 */
extern "C" cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource **resource, unsigned int buffer, unsigned int flags);

static struct {
	cudaGraphicsResource *res = nullptr;
	size_t bytes = 0;
} g_gl;

static void gl_release(void)
{
	if (g_gl.res) { cudaGraphicsUnregisterResource(g_gl.res); cudaGetLastError(); }
	g_gl.res = nullptr;
	g_gl.bytes = 0;
}

extern "C" int rt_cuda_gl_register_buffer(unsigned int gl_buffer, size_t bytes)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (bytes == 0) return fail(RT_ERR_ARG, "GL buffer size must be given");
	if (g.ngpu != 1) return fail(RT_ERR_ARG, "the GL presenter needs a single-GPU context (the GPU that owns the GL context)");
	if ((rc = select_device(g.dev[0])) != RT_OK) return rc;
	if (g_gl.res) { cudaGraphicsUnregisterResource(g_gl.res); g_gl.res = nullptr; }
	cudaError_t e = cudaGraphicsGLRegisterBuffer(&g_gl.res, gl_buffer, cudaGraphicsRegisterFlagsWriteDiscard);
	if (e != cudaSuccess) {
		cudaGetLastError();
		g_gl.res = nullptr;
		return fail(RT_ERR_CUDA, "cudaGraphicsGLRegisterBuffer(%u): %s (is the GL context current on this thread?)", gl_buffer, cudaGetErrorString(e));
	}
	g_gl.bytes = bytes;
	return RT_OK;
}

extern "C" int rt_cuda_gl_unregister_buffer(void)
{
	if (!g_gl.res) return RT_OK;
	cudaError_t e = cudaGraphicsUnregisterResource(g_gl.res);
	g_gl.res = nullptr;
	g_gl.bytes = 0;
	if (e != cudaSuccess) { cudaGetLastError(); return fail(RT_ERR_CUDA, "cudaGraphicsUnregisterResource: %s", cudaGetErrorString(e)); }
	return RT_OK;
}

/* map -> fn(device pointer) -> unmap, on the library stream */
template <class Fn>
static int with_mapped_gl_buffer(int w, int h, const RtRenderOpts *opts, Fn fn)
{
	int rc = require_ready();
	if (rc != RT_OK) return rc;
	if (!g_gl.res) return fail(RT_ERR_STATE, "no GL buffer registered (rt_cuda_gl_register_buffer)");
	if (w <= 0 || h <= 0) return fail(RT_ERR_ARG, "bad frame size %dx%d", w, h);
	RtRenderOpts o;
	if (opts) o = *opts; else rt_render_opts_default(&o);
	if (o.struct_size != sizeof(RtRenderOpts)) return fail(RT_ERR_ARG, "opts->struct_size mismatch (ABI)");
	if ((rc = select_device(g.dev[0])) != RT_OK) return rc;
	cudaStream_t st = o.stream ? (cudaStream_t) o.stream : g.dev[0].stream;
	CU(cudaGraphicsMapResources(1, &g_gl.res, st));
	void *ptr = nullptr;
	size_t mapped = 0;
	cudaError_t e = cudaGraphicsResourceGetMappedPointer(&ptr, &mapped, g_gl.res);
	size_t need = (size_t) w * h * bytes_per_pixel(o.fb_format);
	if (e != cudaSuccess || mapped < need) {
		cudaGetLastError();
		cudaGraphicsUnmapResources(1, &g_gl.res, st);
		if (e != cudaSuccess) return fail(RT_ERR_CUDA, "cudaGraphicsResourceGetMappedPointer: %s", cudaGetErrorString(e));
		return fail(RT_ERR_ARG, "the registered GL buffer holds %zu bytes, a %dx%d frame needs %zu", mapped, w, h, need);
	}
	o.fb_memory = RT_MEM_DEVICE;
	rc = fn(ptr, &o);
	/* unmapping orders the GL commands that follow after the work on `st` */
	cudaError_t u = cudaGraphicsUnmapResources(1, &g_gl.res, st);
	if (rc != RT_OK) return rc;
	if (u != cudaSuccess) { cudaGetLastError(); return fail(RT_ERR_CUDA, "cudaGraphicsUnmapResources: %s", cudaGetErrorString(u)); }
	return RT_OK;
}

extern "C" int rt_cuda_gl_update_frame(const RtCamera *cam, int w, int h, double budget_ms,
                                       const RtRenderOpts *opts, RtRenderStats *stats)
{
	return with_mapped_gl_buffer(w, h, opts, [&](void *ptr, const RtRenderOpts *o) {
		return rt_cuda_update_frame(cam, ptr, w, h, budget_ms, o, stats);
	});
}

extern "C" int rt_cuda_gl_render_frame(const RtCamera *cam, int w, int h, const RtRenderOpts *opts, RtRenderStats *stats)
{
	return with_mapped_gl_buffer(w, h, opts, [&](void *ptr, const RtRenderOpts *o) {
		return render_frame_cuda_ex(cam, ptr, w, h, o, stats);
	});
}
