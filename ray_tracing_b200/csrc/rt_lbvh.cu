/*
 * rt_lbvh.cu -- BVH build for scenes above RT_LBVH_THRESHOLD objects.  Topology: by
 * default the host's binned-SAH builder (bvh_sah.c; 11 % fewer nodes per ray on
 * BASELINE config 5), or entirely on the device (RT_BVH_BUILDER_LBVH: Morton codes
 * -> radix sort -> Karras 2012 hierarchy).  Boxes (bottom-up refit), 16-bit packing
 * and leaf records are made on the device either way.
 *
 * Parity contract: traversal (rt_device.cuh: walk_nodes / walk_leaf) must return
 * what the reference's linear scan (scene.c:156-173) returns, bit for bit.  The
 * per-primitive arithmetic is the same device function as the linear scan, and
 * ties go to the lower primitive index; what the hierarchy must guarantee is
 * that no primitive the reference would accept is ever culled.  The padding
 * rule that guarantees it, and why, is stated once in rt_lbvh_rule.h and
 * checked on the CPU against the O(N) scan by tests/lbvh_sim.c.
 */
#include <cub/device/device_radix_sort.cuh>

#include <math.h>
#include <stdarg.h>
#include <stdio.h>

#include "rt_lbvh.h"
#include "rt_lbvh_rule.h"
#include "rt_cuda.h"

static thread_local char lbvh_err[256] = "";
const char *rt_lbvh_last_error(void) { return lbvh_err; }

static int lfail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(lbvh_err, sizeof(lbvh_err), fmt, ap);
	va_end(ap);
	return code;
}

#define LCU(call)                                                                                  \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess)                                                                     \
			return lfail(RT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                    \
	} while (0)

static double g_fuzz_k = RT_LBVH_FUZZ_K;
static double g_slack = RT_LBVH_SLACK;
extern "C" void rt_lbvh_debug_set(double k, double slack) { g_fuzz_k = k; g_slack = slack; }
/* test / A-B knob: 0 = light samples walk to their nearest hit like every other ray */
static int g_anyhit = 1;
extern "C" void rt_lbvh_debug_set_anyhit(int on) { g_anyhit = on ? 1 : 0; }
/* topology of the next build: RT_BVH_BUILDER_* (include/rt_cuda.h) */
static int g_builder = RT_BVH_BUILDER_SAH;
extern "C" void rt_lbvh_set_builder(int builder) { g_builder = builder; }
static struct { const void *key; int n; double fuzz_r2; int *prim; int *links; int depth; } g_topo = {nullptr, 0, 0.0, nullptr, nullptr, 0};
void rt_lbvh_drop_topology_cache(void)
{
	free(g_topo.prim); free(g_topo.links);
	g_topo.prim = g_topo.links = nullptr;
	g_topo.key = nullptr; g_topo.n = 0;
}
extern "C" int rt_host_bvh_sah(const RtF4 *A, const RtF4 *B, int n, double fuzz_r2, float cube_pad, float extra,
                               int *prim_index, int *children, int *parent, int *depth_out);

/* ---- Morton keys -------------------------------------------------------- */

__device__ __forceinline__ unsigned int spread10(unsigned int v)
{
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}

__global__ void morton_kernel(const float4 *A, const float4 *B, int n, float3 lo, float3 inv_extent,
                              unsigned long long *keys)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 a = A[i], b = B[i];
	float3 c;
	if (__float_as_int(b.w) == RT_OBJECT_SPHERE) c = make_float3(a.x, a.y, a.z);
	else c = make_float3(0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z));
	float x = fminf(fmaxf((c.x - lo.x) * inv_extent.x, 0.0f), 1.0f);
	float y = fminf(fmaxf((c.y - lo.y) * inv_extent.y, 0.0f), 1.0f);
	float z = fminf(fmaxf((c.z - lo.z) * inv_extent.z, 0.0f), 1.0f);
	unsigned int xi = min((unsigned int) (x * 1024.0f), 1023u);
	unsigned int yi = min((unsigned int) (y * 1024.0f), 1023u);
	unsigned int zi = min((unsigned int) (z * 1024.0f), 1023u);
	unsigned int code = (spread10(xi) << 2) | (spread10(yi) << 1) | spread10(zi);
	/* the index in the low word makes every key unique (Karras needs that) */
	keys[i] = ((unsigned long long) code << 32) | (unsigned int) i;
}

/* The record the walk loads for leaf slot i with ONE 32-byte load: geomA of the primitive, then
 * geomB.xyz and the primitive index with its type in the top two bits. */
__device__ __forceinline__ void write_leaf(float4 *leaves, int slot, int p, const float4 *A, const float4 *B)
{
	float4 a = A[p], b = B[p];
	int ty = __float_as_int(b.w);
	ty = ty == RT_OBJECT_CUBE || ty == RT_OBJECT_SPHERE ? ty : 2;
	leaves[2 * (size_t) slot] = a;
	leaves[2 * (size_t) slot + 1] = make_float4(b.x, b.y, b.z, __int_as_float(p | (ty << 30)));
}

/* sorted keys -> primitive index per leaf slot, and the leaf records in that order */
__global__ void unpack_index_kernel(const unsigned long long *keys, int n, const float4 *A, const float4 *B,
                                    int *prim_index, float4 *leaves)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int p = (int) (unsigned int) (keys[i] & 0xffffffffull);
	prim_index[i] = p;
	write_leaf(leaves, i, p, A, B);
}

/* ---- Karras hierarchy ---------------------------------------------------- */

__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j)
{
	if (j < 0 || j >= n) return -1;
	return __clzll((long long) (keys[i] ^ keys[j]));
}

/* children[2i], children[2i+1]: >= 0 internal, < 0 leaf (~slot).  parent[] is
 * indexed by internal node i in [0,n-1) and by leaf slot s at (n-1)+s. */
__global__ void hierarchy_kernel(const unsigned long long *keys, int n, int *children, int *parent)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1) return;
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
	for (int t = (l + 1) / 2; ; t = (t + 1) / 2) {
		if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
		if (t == 1) break;
	}
	int gamma = i + s * d + min(d, 0);
	int lo = min(i, j), hi = max(i, j);
	int left = (lo == gamma) ? ~gamma : gamma;
	int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
	children[2 * i] = left;
	children[2 * i + 1] = right;
	parent[left >= 0 ? left : (n - 1) + ~left] = i;
	parent[right >= 0 ? right : (n - 1) + ~right] = i;
	if (i == 0) parent[0] = -1;
}

/* deepest leaf: what a near-first walk can have on its stack at most */
__global__ void depth_kernel(const int *parent, int n, int *depth_out)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	int d = 0;
	for (int cur = parent[(n - 1) + s]; cur >= 0; cur = parent[cur]) d++;
	atomicMax(depth_out, d);
}

/* ---- padded leaf boxes ---------------------------------------------------- */

__global__ void leaf_box_kernel(const float4 *A, const float4 *B, const int *prim_index, int n,
                                double fuzz_r2, float cube_pad, float extra, float4 *leaf_lo, float4 *leaf_hi)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	int p = prim_index[s];
	float4 a = A[p], b = B[p];
	float3 lo, hi;
	int ty = __float_as_int(b.w);
	if (ty == RT_OBJECT_SPHERE) {
		/* a.w = r*r as the reference computes it (scene.c:112) */
		float rp = (float) sqrt((double) fmaxf(a.w, 0.0f) + fuzz_r2);
		rp = rp * 1.000001f + extra;
		lo = make_float3(a.x - rp, a.y - rp, a.z - rp);
		hi = make_float3(a.x + rp, a.y + rp, a.z + rp);
	} else if (ty == RT_OBJECT_CUBE) {
		float pad = cube_pad + extra;
		lo = make_float3(fminf(a.x, b.x) - pad, fminf(a.y, b.y) - pad, fminf(a.z, b.z) - pad);
		hi = make_float3(fmaxf(a.x, b.x) + pad, fmaxf(a.y, b.y) + pad, fmaxf(a.z, b.z) + pad);
	} else {
		/* unknown type never intersects (scene.c:138-153): empty box */
		lo = make_float3(1.0f, 1.0f, 1.0f);
		hi = make_float3(-1.0f, -1.0f, -1.0f);
	}
	leaf_lo[s] = make_float4(lo.x, lo.y, lo.z, 0.0f);
	leaf_hi[s] = make_float4(hi.x, hi.y, hi.z, 0.0f);
}

/* ---- bottom-up refit ------------------------------------------------------ */

/* node_box[2i], node_box[2i+1] = lo/hi of internal node i (scratch) */
__global__ void refit_kernel(const int *children, const int *parent, const float4 *leaf_lo, const float4 *leaf_hi,
                             int n, unsigned int *visit, float4 *node_box, float4 *nodes)
{
	int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	int cur = parent[(n - 1) + s];
	while (cur >= 0) {
		/* the second thread to arrive owns the node: both subtrees are complete */
		__threadfence();
		if (atomicAdd(&visit[cur], 1u) == 0u) return;
		__threadfence();
		int cl = children[2 * cur], cr = children[2 * cur + 1];
		float4 llo, lhi, rlo, rhi;
		if (cl < 0) { llo = leaf_lo[~cl]; lhi = leaf_hi[~cl]; }
		else { llo = __ldcg(&node_box[2 * cl]); lhi = __ldcg(&node_box[2 * cl + 1]); }
		if (cr < 0) { rlo = leaf_lo[~cr]; rhi = leaf_hi[~cr]; }
		else { rlo = __ldcg(&node_box[2 * cr]); rhi = __ldcg(&node_box[2 * cr + 1]); }
		float4 *nd = nodes + 4 * (size_t) cur;
		nd[0] = make_float4(llo.x, llo.y, llo.z, __int_as_float(cl));
		nd[1] = make_float4(lhi.x, lhi.y, lhi.z, 0.0f);
		nd[2] = make_float4(rlo.x, rlo.y, rlo.z, __int_as_float(cr));
		nd[3] = make_float4(rhi.x, rhi.y, rhi.z, 0.0f);
		/* union; an empty child box (lo > hi) is ignored */
		bool le = llo.x > lhi.x, re = rlo.x > rhi.x;
		float4 ulo, uhi;
		if (le && re) { ulo = llo; uhi = lhi; }
		else if (le) { ulo = rlo; uhi = rhi; }
		else if (re) { ulo = llo; uhi = lhi; }
		else {
			ulo = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
			uhi = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
		}
		__stcg(&node_box[2 * cur], ulo);
		__stcg(&node_box[2 * cur + 1], uhi);
		cur = parent[cur];
	}
}

/* ---- packed tree ---------------------------------------------------------- */

/* binary32 box coordinate -> 16-bit fixed point in the tree's frame q = (x - center) * scale + 32768,
 * rounded outwards and then moved ONE more quantum outwards: the walk evaluates a slab distance as
 * fma(2^23 + q, inv, -(2^23 + 32768 + o') * inv) (rt_device.cuh: walk_ray, walk_nodes), whose second
 * term is rounded at the magnitude 2^23 |inv|, i.e. it places the plane up to ~0.51 quanta off.
 * The frame change is computed in binary32 first; its rounding (at most 2^-9 quanta here, and that of
 * the walk's own o - center) is covered by the `extra` pad of the boxes.  The frame (rt_lbvh_refit)
 * leaves 700 quanta of margin at both ends, so the clamps never bind for a box inside the bounds. */
__device__ __forceinline__ unsigned pack_lo(float v, float c, float s)
{
	float t = floorf((v - c) * s) - 1.0f + 32768.0f;
	return (unsigned) fminf(fmaxf(t, 0.0f), 65535.0f);
}

__device__ __forceinline__ unsigned pack_hi(float v, float c, float s)
{
	float t = ceilf((v - c) * s) + 1.0f + 32768.0f;
	return (unsigned) fminf(fmaxf(t, 0.0f), 65535.0f);
}

__global__ void pack_nodes_kernel(const float4 *nodes, int num_nodes, float cx, float cy, float cz, float s, uint4 *packed)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= num_nodes) return;
	const float4 *nd = nodes + 4 * (size_t) i;
	float4 llo = nd[0], lhi = nd[1], rlo = nd[2], rhi = nd[3];
	/* one word per axis: lo | hi << 16.  An empty box (lo > hi: a primitive of unknown type)
	 * stays empty: 65535 | 0 << 16 on every axis. */
	auto axis = [](float lo, float hi, float c, float sc) {
		if (lo > hi) return 65535u;
		return pack_lo(lo, c, sc) | (pack_hi(hi, c, sc) << 16);
	};
	uint4 q0, q1;
	q0.x = axis(llo.x, lhi.x, cx, s);
	q0.y = axis(llo.y, lhi.y, cy, s);
	q0.z = axis(llo.z, lhi.z, cz, s);
	q0.w = axis(rlo.x, rhi.x, cx, s);
	q1.x = axis(rlo.y, rhi.y, cy, s);
	q1.y = axis(rlo.z, rhi.z, cz, s);
	q1.z = (unsigned) __float_as_int(llo.w);
	q1.w = (unsigned) __float_as_int(rlo.w);
	packed[2 * (size_t) i] = q0;
	packed[2 * (size_t) i + 1] = q1;
}

/* ---- host side ------------------------------------------------------------ */

struct Scratch {
	void *p = nullptr;
	~Scratch() { cudaFree(p); }
};

static int *g_children_of(RtLbvh *bvh) { return bvh->parent + (2 * (size_t) bvh->num_prims - 1); }

void rt_lbvh_free(RtLbvh *bvh)
{
	cudaFree(bvh->nodes); cudaFree(bvh->packed); cudaFree(bvh->prim_index); cudaFree(bvh->parent);
	cudaFree(bvh->leaf_lo); cudaFree(bvh->leaf_hi); cudaFree(bvh->visit);
	cudaFree(bvh->leaves);
	*bvh = RtLbvh();
}

RtBvhView rt_lbvh_view(const RtLbvh *bvh)
{
	RtBvhView v;
	v.nodes = bvh->packed;
	v.cx = bvh->cx; v.cy = bvh->cy; v.cz = bvh->cz;
	v.scale = bvh->scale; v.inv_scale = bvh->inv_scale;
	v.leaves = bvh->leaves;
	v.num_prims = bvh->num_prims;
	v.depth = bvh->depth;
	v.t_slack = bvh->t_slack;
	v.emitter_prim = g_anyhit ? bvh->emitter_prim : -1;
	v.emitter_slot = g_anyhit ? bvh->emitter_slot : -1;
	if (v.emitter_slot < 0) v.emitter_prim = -1;
	return v;
}

float rt_lbvh_required_dmax(const RtLbvh *bvh, RtVector3 p)
{
	double dx = fmax(fabs((double) p.x - bvh->lo.x), fabs((double) p.x - bvh->hi.x));
	double dy = fmax(fabs((double) p.y - bvh->lo.y), fabs((double) p.y - bvh->hi.y));
	double dz = fmax(fabs((double) p.z - bvh->lo.z), fabs((double) p.z - bvh->hi.z));
	return (float) sqrt(dx * dx + dy * dy + dz * dz);
}

int rt_lbvh_refit(RtLbvh *bvh, const float4 *geomA, const float4 *geomB, float d_max, cudaStream_t stream)
{
	int n = bvh->num_prims;
	if (n <= 0) return RT_OK;
	double mag = fmax(fmax(fabs((double) bvh->lo.x), fabs((double) bvh->hi.x)),
	                  fmax(fmax(fabs((double) bvh->lo.y), fabs((double) bvh->hi.y)),
	                       fmax(fabs((double) bvh->lo.z), fabs((double) bvh->hi.z))));
	RtLbvhPads pads = rt_lbvh_pads(mag, (double) d_max, g_fuzz_k, g_slack);
	int blocks = (n + 255) / 256;
	leaf_box_kernel<<<blocks, 256, 0, stream>>>(geomA, geomB, bvh->prim_index, n, pads.fuzz_r2, pads.cube_pad, pads.extra,
	                                            bvh->leaf_lo, bvh->leaf_hi);
	LCU(cudaGetLastError());
	if (n >= 2) {
		Scratch node_box;
		LCU(cudaMalloc(&node_box.p, sizeof(float4) * 2 * (size_t) (n - 1)));
		LCU(cudaMemsetAsync(bvh->visit, 0, sizeof(unsigned int) * (size_t) (n - 1), stream));
		refit_kernel<<<blocks, 256, 0, stream>>>(g_children_of(bvh), bvh->parent, bvh->leaf_lo, bvh->leaf_hi, n,
		                                         bvh->visit, (float4 *) node_box.p, bvh->nodes);
		LCU(cudaGetLastError());
		/* the frame of the packed boxes: centred on the (padded) bounds, scaled by a power of two so
		 * that they span at most +-32000 of the +-32768 quanta (config 5: 1/512 of a unit per quantum) */
		double pad = (double) sqrt(pads.fuzz_r2) + pads.extra + pads.cube_pad;
		double hx = 0.5 * ((double) bvh->hi.x - bvh->lo.x) + pad, hy = 0.5 * ((double) bvh->hi.y - bvh->lo.y) + pad,
		       hz = 0.5 * ((double) bvh->hi.z - bvh->lo.z) + pad;
		double half = fmax(fmax(hx, hy), fmax(hz, 1e-30));
		int e = (int) floor(log2(32000.0 / half));
		if (e > 100) e = 100;
		if (e < -100) e = -100;
		bvh->scale = (float) ldexp(1.0, e);
		bvh->inv_scale = (float) ldexp(1.0, -e);
		bvh->cx = (float) (0.5 * ((double) bvh->hi.x + bvh->lo.x));
		bvh->cy = (float) (0.5 * ((double) bvh->hi.y + bvh->lo.y));
		bvh->cz = (float) (0.5 * ((double) bvh->hi.z + bvh->lo.z));
		pack_nodes_kernel<<<(n - 1 + 255) / 256, 256, 0, stream>>>(bvh->nodes, n - 1, bvh->cx, bvh->cy, bvh->cz, bvh->scale, bvh->packed);
		LCU(cudaGetLastError());
		LCU(cudaStreamSynchronize(stream));
	}
	bvh->d_max = d_max;
	bvh->t_slack = pads.t_slack;
	return RT_OK;
}

/* Morton slot of primitive `prim` (the scene's only emitter) */
__global__ void find_slot_kernel(const int *prim_index, int n, int prim, int *slot_out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && prim_index[i] == prim) *slot_out = i;
}

static int locate_emitter(RtLbvh *bvh, int prim, cudaStream_t stream)
{
	bvh->emitter_prim = -1;
	bvh->emitter_slot = -1;
	int n = bvh->num_prims;
	if (prim < 0 || prim >= n) return RT_OK;
	/* visit[0] is free between refits (it doubles as a result cell, as for the depth) */
	int slot = -1;
	LCU(cudaMemcpyAsync(bvh->visit, &slot, sizeof(int), cudaMemcpyHostToDevice, stream));
	find_slot_kernel<<<(n + 255) / 256, 256, 0, stream>>>(bvh->prim_index, n, prim, (int *) bvh->visit);
	LCU(cudaGetLastError());
	LCU(cudaMemcpyAsync(&slot, bvh->visit, sizeof(int), cudaMemcpyDeviceToHost, stream));
	LCU(cudaStreamSynchronize(stream));
	if (slot >= 0) { bvh->emitter_prim = prim; bvh->emitter_slot = slot; }
	return RT_OK;
}

__global__ void gather_leaves_kernel(const int *prim_index, int n, const float4 *A, const float4 *B, float4 *leaves)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	write_leaf(leaves, i, prim_index[i], A, B);
}

/* Objects moved (same count): keep the topology, refresh the leaf records and refit
 * every box bottom-up.  The tree stays correct for any motion -- the boxes are
 * recomputed from the new primitives -- but its quality degrades with large
 * displacements; rebuild (rt_cuda_upload_objects) when that matters. */
int rt_lbvh_update(RtLbvh *bvh, const float4 *geomA, const float4 *geomB, const RtPackedScene *hs, cudaStream_t stream)
{
	int n = bvh->num_prims;
	if (n <= 0) return RT_OK;
	if (hs->n != n) return lfail(RT_ERR_ARG, "LBVH update: %d objects, the tree was built for %d", hs->n, n);
	bvh->lo = hs->bounds_lo;
	bvh->hi = hs->bounds_hi;
	gather_leaves_kernel<<<(n + 255) / 256, 256, 0, stream>>>(bvh->prim_index, n, geomA, geomB, bvh->leaves);
	LCU(cudaGetLastError());
	int rc = locate_emitter(bvh, hs->only_emitter, stream);
	if (rc != RT_OK) return rc;
	float d_max = rt_lbvh_default_dmax((double) hs->bounds_hi.x - hs->bounds_lo.x, (double) hs->bounds_hi.y - hs->bounds_lo.y,
	                                   (double) hs->bounds_hi.z - hs->bounds_lo.z);
	return rt_lbvh_refit(bvh, geomA, geomB, d_max, stream);
}

int rt_lbvh_build(RtLbvh *bvh, const float4 *geomA, const float4 *geomB, int n,
                  const RtPackedScene *hs, cudaStream_t stream)
{
	rt_lbvh_free(bvh);
	if (n <= 0) return RT_OK;
	bvh->num_prims = n;
	bvh->lo = hs->bounds_lo;
	bvh->hi = hs->bounds_hi;
	size_t nn = (size_t) n;
	LCU(cudaMalloc(&bvh->nodes, sizeof(float4) * 4 * (nn > 1 ? nn - 1 : 1)));
	LCU(cudaMalloc(&bvh->packed, sizeof(uint4) * 2 * (nn > 1 ? nn - 1 : 1)));
	LCU(cudaMalloc(&bvh->prim_index, sizeof(int) * nn));
	/* parent[0 .. 2n-1) followed by children[0 .. 2(n-1)) */
	LCU(cudaMalloc(&bvh->parent, sizeof(int) * ((2 * nn - 1) + 2 * (nn > 1 ? nn - 1 : 1))));
	LCU(cudaMalloc(&bvh->leaf_lo, sizeof(float4) * nn));
	LCU(cudaMalloc(&bvh->leaf_hi, sizeof(float4) * nn));
	LCU(cudaMalloc(&bvh->visit, sizeof(unsigned int) * (nn > 1 ? nn - 1 : 1)));
	if (nn >= (1u << 30)) return lfail(RT_ERR_ARG, "the LBVH holds at most 2^30 primitives");
	LCU(cudaMalloc(&bvh->leaves, sizeof(float4) * 2 * nn));

	float3 lo = make_float3(hs->bounds_lo.x, hs->bounds_lo.y, hs->bounds_lo.z);
	float3 ext = make_float3(hs->bounds_hi.x - lo.x, hs->bounds_hi.y - lo.y, hs->bounds_hi.z - lo.z);
	int blocks = (n + 255) / 256;
	if (g_builder == RT_BVH_BUILDER_SAH) {
		/* topology from the host (bvh_sah.c: binned SAH over the padded boxes the refit below will
		 * compute); boxes, packing and everything the walk reads are made on the device as for the
		 * Karras tree */
		double mag = fmax(fmax(fabs((double) bvh->lo.x), fabs((double) bvh->hi.x)),
		                  fmax(fmax(fabs((double) bvh->lo.y), fabs((double) bvh->hi.y)),
		                       fmax(fabs((double) bvh->lo.z), fabs((double) bvh->hi.z))));
		RtLbvhPads pads = rt_lbvh_pads(mag, (double) rt_lbvh_default_dmax((double) ext.x, (double) ext.y, (double) ext.z), g_fuzz_k, g_slack);
		size_t links = (2 * nn - 1) + 2 * (nn > 1 ? nn - 1 : 1);
		/* one host build serves every device of an upload (rt_api.cu calls this once per GPU with the
		 * same packed scene and drops the cache when the upload is over) */
		if (g_topo.key != hs->geomA || g_topo.n != n || g_topo.fuzz_r2 != pads.fuzz_r2) {
			rt_lbvh_drop_topology_cache();
			g_topo.prim = (int *) malloc(sizeof(int) * nn);
			g_topo.links = (int *) calloc(links, sizeof(int));
			if (!g_topo.prim || !g_topo.links ||
			    rt_host_bvh_sah(hs->geomA, hs->geomB, n, pads.fuzz_r2, pads.cube_pad, pads.extra, g_topo.prim,
			                    g_topo.links + (2 * nn - 1), g_topo.links, &g_topo.depth) != 0) {
				rt_lbvh_drop_topology_cache();
				return lfail(RT_ERR_NOMEM, "out of host memory building the BVH topology of %d primitives", n);
			}
			g_topo.key = hs->geomA; g_topo.n = n; g_topo.fuzz_r2 = pads.fuzz_r2;
		}
		LCU(cudaMemcpyAsync(bvh->prim_index, g_topo.prim, sizeof(int) * nn, cudaMemcpyHostToDevice, stream));
		LCU(cudaMemcpyAsync(bvh->parent, g_topo.links, sizeof(int) * links, cudaMemcpyHostToDevice, stream));
		LCU(cudaStreamSynchronize(stream));
		bvh->depth = g_topo.depth;
		gather_leaves_kernel<<<blocks, 256, 0, stream>>>(bvh->prim_index, n, geomA, geomB, bvh->leaves);
		LCU(cudaGetLastError());
	} else {
		Scratch keys_in, keys_out, temp;
		LCU(cudaMalloc(&keys_in.p, sizeof(unsigned long long) * nn));
		LCU(cudaMalloc(&keys_out.p, sizeof(unsigned long long) * nn));

		float3 inv = make_float3(ext.x > 0 ? 1.0f / ext.x : 0.0f, ext.y > 0 ? 1.0f / ext.y : 0.0f,
		                         ext.z > 0 ? 1.0f / ext.z : 0.0f);
		morton_kernel<<<blocks, 256, 0, stream>>>(geomA, geomB, n, lo, inv, (unsigned long long *) keys_in.p);
		LCU(cudaGetLastError());

		size_t temp_bytes = 0;
		LCU(cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, (const unsigned long long *) keys_in.p,
		                                   (unsigned long long *) keys_out.p, n, 0, 64, stream));
		LCU(cudaMalloc(&temp.p, temp_bytes ? temp_bytes : 4));
		LCU(cub::DeviceRadixSort::SortKeys(temp.p, temp_bytes, (const unsigned long long *) keys_in.p,
		                                   (unsigned long long *) keys_out.p, n, 0, 64, stream));
		unpack_index_kernel<<<blocks, 256, 0, stream>>>((const unsigned long long *) keys_out.p, n, geomA, geomB,
		                                                bvh->prim_index, bvh->leaves);
		LCU(cudaGetLastError());
		if (n >= 2) {
			hierarchy_kernel<<<blocks, 256, 0, stream>>>((const unsigned long long *) keys_out.p, n,
			                                             g_children_of(bvh), bvh->parent);
			LCU(cudaGetLastError());
		}
		if (n >= 2) {
			/* visit[0] doubles as the result cell (refit clears the array afterwards) */
			LCU(cudaMemsetAsync(bvh->visit, 0, sizeof(unsigned int), stream));
			depth_kernel<<<blocks, 256, 0, stream>>>(bvh->parent, n, (int *) bvh->visit);
			LCU(cudaGetLastError());
			LCU(cudaMemcpyAsync(&bvh->depth, bvh->visit, sizeof(int), cudaMemcpyDeviceToHost, stream));
		}
		LCU(cudaStreamSynchronize(stream));
	}
	int rc = locate_emitter(bvh, hs->only_emitter, stream);
	if (rc != RT_OK) return rc;

	/* secondary rays start on surfaces, i.e. inside the primitive bounds: their
	 * distance to any primitive is at most the bounds' diagonal.  The renderer
	 * re-pads when the camera is farther than that (rt_api.cu). */
	float d_max = rt_lbvh_default_dmax((double) ext.x, (double) ext.y, (double) ext.z);
	return rt_lbvh_refit(bvh, geomA, geomB, d_max, stream);
}
