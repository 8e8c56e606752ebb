/*
 * rt_lbvh.h -- device LBVH for scenes too large for the shared-memory linear
 * scan (BASELINE.json config 5: 100 000 spheres).  The reference has no
 * acceleration structure (scene.c:163 is an O(N) loop); the LBVH must return
 * exactly what that loop returns, so it is built around the reference's
 * *computed* intersection test, not the geometric primitives (see rt_lbvh.cu).
 */
#ifndef RT_LBVH_H
#define RT_LBVH_H

#include <cuda_runtime.h>
#include "rt_host.h"
#include "rt_params.h"

struct RtLbvh {
	float4 *nodes = nullptr;        /* binary32 working copy: 4 float4 per internal node (both child boxes + children) */
	uint4  *packed = nullptr;       /* what the walk reads: 2 uint4 per internal node, 16-bit fixed-point boxes (rt_params.h) */
	float   cx = 0, cy = 0, cz = 0, scale = 1, inv_scale = 1;   /* frame of the packed boxes */
	int    *prim_index = nullptr;   /* Morton order -> primitive index */
	int    *parent = nullptr;       /* internal-node parents; leaves at [n-1, 2n-1) */
	float4 *leaf_lo = nullptr, *leaf_hi = nullptr;   /* padded primitive boxes, Morton order */
	float4 *leaves = nullptr;       /* 2 float4 per leaf slot, Morton order: what walk_leaf_screen() loads with one 32-byte load */
	unsigned int *visit = nullptr;
	int     num_prims = 0;
	int     depth = 0;              /* deepest leaf, in levels below the root */
	int     emitter_prim = -1;      /* RtPackedScene::only_emitter and its Morton slot (-1: none or several) */
	int     emitter_slot = -1;
	float   d_max = 0.0f;           /* largest origin-to-primitive distance the padding covers */
	float   t_slack = 0.0f;
	RtVector3 lo = {0, 0, 0}, hi = {0, 0, 0};   /* unpadded bounds of all primitives */
};

int  rt_lbvh_build(RtLbvh *bvh, const float4 *geomA, const float4 *geomB, int n,
                   const RtPackedScene *host_scene, cudaStream_t stream);
/* Re-pad and refit for ray origins up to `d_max` away from any primitive
 * (topology unchanged).  Called when the camera leaves the region the build
 * assumed. */
int  rt_lbvh_refit(RtLbvh *bvh, const float4 *geomA, const float4 *geomB, float d_max, cudaStream_t stream);
int  rt_lbvh_update(RtLbvh *bvh, const float4 *geomA, const float4 *geomB, const RtPackedScene *host_scene, cudaStream_t stream);
void rt_lbvh_free(RtLbvh *bvh);
/* the host-built topology is kept between the per-device builds of one upload; call when the upload is over */
void rt_lbvh_drop_topology_cache(void);
RtBvhView rt_lbvh_view(const RtLbvh *bvh);
const char *rt_lbvh_last_error(void);

/* distance from `p` to the farthest corner of the primitive bounds */
float rt_lbvh_required_dmax(const RtLbvh *bvh, RtVector3 p);

#endif
