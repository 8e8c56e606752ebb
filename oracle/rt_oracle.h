/*
 * rt_oracle.h -- TEST INFRASTRUCTURE (CPU restatement of the reference's
 * per-pixel render path).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load librt_oracle.so; the product
 * (ray_tracing_b200/) never does.
 *
 * Parity status: PINNED against the unmodified reference compiled here
 * (oracle/_ref/libref_pixel.so, see tests/test_oracle_vs_ref.py) and against the
 * golden frames generated from it (tests/golden/, tests/golden/make_golden.py).
 * The reference itself ships no tests or golden vectors.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stdint.h>
#include <stddef.h>

/* 68-byte record, byte-compatible with the reference `Object` (scene.h:24-31):
 * type@0 (0 = cube, 1 = sphere), geometry union@4 (sphere: center xyz, radius;
 * cube: origin xyz, size xyz), material@28. */
typedef struct {
	int32_t type;
	float   geom[6];
	float   albedo[3];
	float   roughness;
	float   reflectance;
	float   metallic;
	float   emission_power;
	float   emission_color[3];
} RtoObject;

typedef struct {
	float pos[3];
	float front[3];
	float up[3];
	float fov;
} RtoCamera;

typedef struct {
	const uint8_t *face[6];   /* CubeFace order: front, back, left, right, top, bottom */
	int w, h, chan;
} RtoSky;

typedef struct {
	const RtoObject *objects;
	int              num_objects;
	RtoCamera        camera;
	RtoSky           sky;
} RtoWorld;

/* RNG (utils.c:60-75) */
uint64_t rto_wyhash64(uint64_t *state);
float    rto_random_float(uint64_t *state);
void     rto_random_direction(uint64_t *state, float out[3]);
uint64_t rto_pixel_key(float px, float py, uint64_t pass);

/* camera.c:95-125 */
void rto_camera_ray(const RtoCamera *cam, float px, float py, float aspect, float out_ray[6]);

/* scene.c:156-190 ; out7 = distance, point xyz, normal xyz ; obj = index or -1 */
void rto_trace_many(const RtoObject *objects, int n, const float *rays6, int nrays,
                    float *out7, int32_t *obj);

/* gpu_and_windowing.c:42-112 */
void rto_sample_cubemap_many(const RtoSky *sky, const float *dirs3, int n, float *out3);

/* main.c:131-272 with an explicit RNG state */
void rto_pixel(const RtoWorld *w, float px, float py, float aspect, uint64_t rng_state,
               float out[3], uint64_t *rays);

/*
 * One pass of render_column() over all columns (main.c:274-322), RNG re-keyed
 * per pixel with rto_pixel_key(u, v, pass).  out = W*H*3 floats, row 0 = bottom
 * row.  Only rows [row0,row1) are produced (others untouched).  Pixels the
 * reference never writes are set to 0.  nthreads host threads split the rows.
 * Returns rays traced for the visible low-res pixels of those rows.
 */
uint64_t rto_render(const RtoWorld *w, float *out, int W, int H, int scale,
                    int num_columns, uint64_t pass, int row0, int row1, int nthreads);

/* accumulate/resolve (main.c:387-396, 467-477) */
void rto_accumulate(float *accum, const float *pass_data, size_t n_floats, int scale);
void rto_resolve(float *frame, const float *accum, size_t n_floats, float count);

#endif
