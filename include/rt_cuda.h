/*
 * rt_cuda.h -- C ABI of the B200-native render path for cozis/ray_tracing.
 *
 * This library replaces, behind plain C entry points, the reference's
 *   worker()/render_column()/pixel()        src/main.c:131-414
 *   trace_ray() + intersectors              src/scene.c:10-190
 *   ray_through_screen_at()                 src/camera.c:95-125
 *   sample_cubemap()                        src/gpu_and_windowing.c:42-112
 *   the accumulate/resolve of update_frame  src/main.c:387-396, 467-477
 * and keeps the reference's host-side API for the callers either side of it:
 *   parse_scene_file()                      src/scene.h:47, scene.c:611-624
 *   move_camera/rotate_camera/get_camera_pos src/camera.h:25-29
 * There is no plugin/FFI layer in the reference; the seam is those C functions
 * (SURVEY.md section 8(b)).  INTEGRATION.md shows the lines a maintainer changes
 * in main.c.
 *
 * Conventions: every call returns RT_OK (0) or a negative RT_ERR_* code and
 * never aborts; rt_cuda_last_error() gives the message.  All structs are POD.
 * Calls are made from one host thread, and the library holds ONE renderer per
 * process (a single context: devices, scene, skybox, accumulation, frame
 * scheduler), as the reference does.  There is NO CPU fallback: without a
 * CUDA device every rendering call fails with RT_ERR_NO_DEVICE.
 */
#ifndef RT_CUDA_H
#define RT_CUDA_H

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* Types, layout-identical to the reference's (sizes probed in SURVEY.md R11) */
/* ------------------------------------------------------------------------- */

#ifdef RT_CUDA_REFERENCE_TYPES
/* Building inside the reference tree: vector.h / scene.h /
 * gpu_and_windowing.h were included first; reuse their types. */
typedef Vector3  RtVector3;
typedef Material RtMaterial;
typedef Object   RtObject;
typedef Scene    RtScene;
typedef Cubemap  RtCubemap;
#define RT_MAX_OBJECTS MAX_OBJECTS
#else

#define RT_MAX_OBJECTS 1024                 /* scene.h:3 */

typedef struct { float x, y, z; } RtVector3; /* vector.h:32-36, 12 B */

typedef struct {                            /* scene.h:5-12, 40 B */
	RtVector3 albedo;
	float     roughness;
	float     reflectance;
	float     metallic;
	float     emission_power;
	RtVector3 emission_color;
} RtMaterial;

typedef enum { RT_OBJECT_CUBE = 0, RT_OBJECT_SPHERE = 1 } RtObjectType;   /* scene.h:19-22 */

typedef struct {                            /* scene.h:24-31, 68 B */
	RtObjectType type;
	union {
		struct { RtVector3 center; float radius; } sphere;   /* vector.h:58-61 */
		struct { RtVector3 origin; RtVector3 size; } cube;    /* scene.h:14-17 */
	};
	RtMaterial material;
} RtObject;

typedef struct {                            /* scene.h:33-36, 69 636 B */
	RtObject objects[RT_MAX_OBJECTS];
	int      num_objects;
} RtScene;

typedef struct {                            /* gpu_and_windowing.h:4-7 */
	uint8_t *data[6];                       /* CubeFace order (gpu_and_windowing.h:9-16) */
	int w, h, chan;
} RtCubemap;
#endif

enum { RT_CF_FRONT = 0, RT_CF_BACK, RT_CF_LEFT, RT_CF_RIGHT, RT_CF_TOP, RT_CF_BOTTOM };

/* The reference keeps the pose in file statics (camera.c:23-35); this is the
 * snapshot handed to the renderer. */
typedef struct {
	RtVector3 pos;
	RtVector3 front;                        /* not normalised by the caller */
	RtVector3 up;
	float     fov;                          /* passed to tan() as radians, camera.c:107 */
} RtCamera;

typedef enum { RT_UP = 0, RT_DOWN, RT_LEFT, RT_RIGHT } RtDirection;   /* camera.h:20-22 */

/* ------------------------------------------------------------------------- */
/* Error codes                                                               */
/* ------------------------------------------------------------------------- */
enum {
	RT_OK              = 0,
	RT_ERR_NO_DEVICE   = -1,   /* no CUDA device / driver: there is no CPU fallback */
	RT_ERR_CUDA        = -2,   /* a CUDA runtime call failed */
	RT_ERR_ARG         = -3,   /* invalid argument */
	RT_ERR_STATE       = -4,   /* missing init / scene / skybox */
	RT_ERR_NOMEM       = -5,
	RT_ERR_PARSE       = -6,
	RT_ERR_IO          = -7,
};

const char *rt_cuda_last_error(void);

/* ------------------------------------------------------------------------- */
/* Host side kept from the reference                                          */
/* ------------------------------------------------------------------------- */

/* scene.c:611-624 -- same grammar, same float construction, same messages on
 * stderr, same return convention.  Writes the same bytes the reference writes
 * (fields of each Object it assigns; union tail / padding are left untouched). */
bool rt_parse_scene_file(const char *file, RtScene *scene);
bool rt_parse_scene_string(const char *src, size_t len, RtScene *scene);

/* Large-scene variant: same grammar, heap array, no 1024 cap.
 * *objects is malloc'ed; release with rt_free_objects(). */
int  rt_parse_scene_file_large(const char *file, RtObject **objects, int *num_objects);
int  rt_parse_scene_string_large(const char *src, size_t len, RtObject **objects, int *num_objects);
void rt_free_objects(RtObject *objects);

/* camera.c:37-93 -- process-global pose with the reference's mutators. */
void      rt_camera_reset(void);
void      rt_move_camera(RtDirection dir, float speed);          /* camera.c:80-88  */
void      rt_rotate_camera(double mouse_x, double mouse_y);      /* camera.c:42-78  */
RtVector3 rt_get_camera_pos(void);                               /* camera.c:37-40  */
RtCamera  rt_camera_snapshot(void);

/* main.c:666-670 quantisation rule: (uint8_t)(x*255), row order unchanged. */
void rt_quantize_frame(const float *frame_rgb, size_t num_pixels, uint8_t *out_rgb);
/* screenshot() (main.c:637-681) without the file-name search: quantise, flip vertically, write an
 * 8-bit RGB PNG (or a binary PPM when the path ends in ".ppm"). */
int  rt_save_screenshot(const char *path, const float *frame_rgb, int w, int h);

/* ------------------------------------------------------------------------- */
/* Device lifecycle                                                          */
/* ------------------------------------------------------------------------- */

/* Single-process mode: use GPUs [0, num_gpus) (num_gpus <= 0 -> 1).  Rows are
 * split into contiguous bands, one per GPU, composited into GPU 0 over
 * NVLink P2P. */
int  rt_cuda_init(int num_gpus);
/* One-process-per-GPU mode (torchrun ranks): bind to exactly this device. */
int  rt_cuda_init_device(int device);
void rt_cuda_shutdown(void);
int  rt_cuda_num_gpus(void);

/* AoS -> device SoA.  Scenes above RT_LBVH_THRESHOLD objects also get a device
 * LBVH whose hits tie-break by primitive index (== the reference's linear scan). */
int  rt_cuda_upload_scene(const RtScene *scene);
int  rt_cuda_upload_objects(const RtObject *objects, int num_objects);
/* The same objects changed in place (same count, same types in the same order): refresh the
 * device records and refit the LBVH (topology kept) instead of rebuilding it. */
int  rt_cuda_update_objects(const RtObject *objects, int num_objects);
int  rt_cuda_upload_skybox(const RtCubemap *sky);

#define RT_LBVH_THRESHOLD 64

/* Who shapes the tree of the NEXT rt_cuda_upload_scene / rt_cuda_upload_objects of a large scene.
 * Either way the frames are the reference's, bit for bit (the walk applies scene.c:79-134 per
 * primitive and breaks ties by index); the topology only decides how many nodes a ray visits. */
enum { RT_BVH_BUILDER_SAH  = 0,   /* default: binned surface-area heuristic on the host (threads), ~40 ms per 100 000
                                     primitives; BASELINE config 5 renders 8-9 % faster than with the Karras tree */
       RT_BVH_BUILDER_LBVH = 1 }; /* Morton order + Karras hierarchy, entirely on the device (~1 ms of kernels):
                                     for scenes rebuilt every frame */
int  rt_cuda_set_bvh_builder(int builder);

/* ------------------------------------------------------------------------- */
/* Rendering                                                                 */
/* ------------------------------------------------------------------------- */

enum { RT_FB_F32X3 = 0,    /* reference `Vector3*` frame: 12 B/px, bottom row first */
       RT_FB_U8X4  = 1 };  /* (uint8_t)(x*255) per channel + alpha 255, 4 B/px      */

enum { RT_VARIANT_EXACT = 0,  /* no FMA contraction, IEEE div/sqrt, f64 where C promotes: bit-exact */
       RT_VARIANT_FAST  = 1 };/* FMA contraction allowed: <= 1 LSB (8-bit) on >= 99.9 % of pixels */

enum { RT_TRAVERSAL_AUTO = 0, RT_TRAVERSAL_LINEAR = 1, RT_TRAVERSAL_LBVH = 2 };

enum { RT_MEM_AUTO = 0, RT_MEM_HOST = 1, RT_MEM_DEVICE = 2 };

enum { RT_KERNEL_AUTO = 0,
       RT_KERNEL_PIXEL = 1,       /* one thread per low-res pixel, runs its whole path */
       RT_KERNEL_PERSISTENT = 2,  /* persistent warps, lanes refill with new pixels as paths end */
       RT_KERNEL_WAVEFRONT = 3,   /* per-warp path pool in shared memory, phases over compacted lists */
       RT_KERNEL_QUEUED = 4 };    /* persistent warps + per-warp shared-memory queues for finishing / preparing pixels */

typedef struct {
	uint32_t struct_size;   /* = sizeof(RtRenderOpts) */
	int      scale;         /* render_column's scale: 1,2,4,8,16 (any >= 1) */
	int      num_columns;   /* reference --threads; reproduces its column artefacts. default 1 */
	uint64_t pass_index;    /* RNG key: state = key(u, v, pass_index) at the top of each pixel */
	int      fb_format;     /* RT_FB_* */
	int      fb_memory;     /* RT_MEM_*: where `fb` lives (AUTO = cudaPointerGetAttributes) */
	int      row_begin;     /* band [row_begin,row_end) of output rows; 0,0 = whole frame. */
	int      row_end;       /*   must be multiples of scale (except row_end == h) */
	int      accumulate;    /* 0: fb <- this pass.  1: accum += pass/scale^2; fb <- accum/count */
	int      variant;       /* RT_VARIANT_* */
	int      traversal;     /* RT_TRAVERSAL_* */
	int      kernel;        /* RT_KERNEL_* */
	int      band_only_fb;  /* 1: fb holds only the band's rows (row_begin maps to fb row 0) */
	void    *stream;        /* cudaStream_t to launch on in one-device mode; NULL = library stream */
	int      interleave_count; /* one-GPU-per-process ranks sharing a frame: this call renders the row   */
	int      interleave_index; /* blocks (16 output rows) b of the band with b % count == index; 0/1 = all */
	int      pipeline;         /* host `fb` only: return before the device->host copy has finished; frames of  */
	                           /*    consecutive calls overlap (render k+1 || copy k).  `fb` should be pinned;    */
	                           /*    it is valid after rt_cuda_synchronize()                                      */
	int      remote_fb;        /* 1: `fb` is peer memory (another GPU's frame): render locally, then copy    */
	                           /*    the owned blocks there with one strided device-to-device copy           */
	uint32_t frame_seq;        /* remote_fb only, `fb` from rt_cuda_shared_frame_*: 0 = the copy follows the  */
	                           /*    render on the same stream.  s >= 1 (consecutive per frame) = pipelined   */
	                           /*    composite: the call returns after queueing; the copy of frame s runs on  */
	                           /*    the copy stream while frame s+1 renders, then marks this rank's blocks   */
	                           /*    of frame s as arrived (rt_cuda_shared_frame_wait)                        */
	int      frame_ack;        /* frame_seq only: do not overwrite the shared frame with frame s before the  */
	                           /*    owner released frame s-1 (rt_cuda_shared_frame_release)                  */
} RtRenderOpts;

typedef struct {
	uint64_t rays;          /* trace_ray-equivalent invocations of this call */
	uint64_t pixels;        /* low-res pixels evaluated */
	float    render_ms;     /* device time of the render kernels (max over GPUs) */
	float    composite_ms;  /* device time of the band composite to GPU 0 (0 for 1 GPU) */
	float    copy_ms;       /* device->host copy if fb is host memory */
	int      kernel_launches;
} RtRenderStats;

void rt_render_opts_default(RtRenderOpts *opts);

/* north_star signature.  `scene` may be NULL to reuse the uploaded scene; when
 * non-NULL it is (re)uploaded if its contents changed.  fb: w*h RtVector3
 * (host or device memory), bottom row first, values clamped to [0,1]. */
int render_frame_cuda(const RtScene *scene, const RtCamera *cam, void *fb, int w, int h, int scale);

int render_frame_cuda_ex(const RtCamera *cam, void *fb, int w, int h,
                         const RtRenderOpts *opts, RtRenderStats *stats);

/* invalidate_accumulation() (main.c:115-124): zero accum and the weight. */
int   rt_cuda_accum_reset(void);
float rt_cuda_accum_count(void);

/* Progressive refinement as the reference's workers do after an invalidation
 * (main.c:354,402-403): passes at init_scale, init_scale/2, ..., 1 with
 * pass_index first_pass, first_pass+1, ..., each accumulated and resolved;
 * fb holds the resolved frame after the last pass.  opts->scale/accumulate
 * are ignored. */
int rt_cuda_render_sweep(const RtCamera *cam, void *fb, int w, int h, int init_scale,
                         uint64_t first_pass, const RtRenderOpts *opts, RtRenderStats *stats);

/* One-process-per-GPU composite over NVLink without a data-path collective:
 * rank 0 creates the frame and shares the 64-byte handle; the other ranks open
 * it and pass the mapped address as `fb` with opts->interleave_* and
 * opts->remote_fb = 1: they render the row blocks they own into local memory
 * and ship them with one strided peer copy into the shared frame (direct peer
 * stores from the kernel were measured first: small scattered writes over
 * NVLink, slower).  close: owner = 1 on the creating rank, 0 elsewhere. */
int rt_cuda_shared_frame_create(size_t bytes, void **dev_ptr, void *handle64);
int rt_cuda_shared_frame_open(const void *handle64, void **dev_ptr);
int rt_cuda_shared_frame_close(void *dev_ptr, int owner);
/* Pipelined composite (opts->frame_seq): the frame carries a small header of
 * flag words in the owner's memory.  Every rank -- the owner too, with
 * remote_fb = 1 -- renders frame s into one of two local frames and ships its
 * blocks on its copy stream while it already renders frame s+1; after the copy
 * it writes s into its `arrived` word.
 *   wait     owner: work queued on `stream` after this call starts once the
 *            blocks of frame `seq` of all `num_ranks` ranks have landed;
 *   release  owner: frame `seq` has been consumed; ranks rendering with
 *            opts->frame_ack may overwrite it with frame seq+1.
 * Both are stream operations (tiny polling kernels with a 2 s timeout that
 * raises the header's error word; rt_cuda_shared_frame_error reads it). */
int rt_cuda_shared_frame_wait(void *dev_ptr, int num_ranks, uint32_t seq, void *stream);
int rt_cuda_shared_frame_release(void *dev_ptr, uint32_t seq, void *stream);
int rt_cuda_shared_frame_error(void *dev_ptr, uint32_t *error_out);
int rt_cuda_copy_to_host(void *host_dst, const void *dev_src, size_t bytes, void *stream);
/* stream-ordered copy (no wait) between any two addresses, e.g. a consumer draining the shared frame */
int rt_cuda_copy_async(void *dst, const void *src, size_t bytes, void *stream);

/* The reference's frame scheduler (main.c:324-482) in three calls:
 *   rt_cuda_set_progressive(init_scale, num_columns)   the --init-scale / --threads flags
 *   rt_cuda_invalidate_accumulation()                  invalidate_accumulation() (main.c:115-124):
 *                                                      accum = 0, generation++, back to init_scale
 *   rt_cuda_update_frame(cam, fb, w, h, budget_ms,...)  update_frame() (main.c:450-482): at least one
 *       pass at the current scale (halving after each, main.c:402-403), then more passes while the
 *       device time spent stays within budget_ms (0 = exactly one pass; negative = refine the pose to
 *       full resolution: every remaining pass of the ladder down to scale 1, or one pass when it is
 *       there already); fb = accum / count.  A fresh pose whose whole ladder fits the budget runs
 *       its passes side by side (as rt_cuda_render_sweep does): same pass indices, same frame. */
int      rt_cuda_set_progressive(int init_scale, int num_columns);
int      rt_cuda_invalidate_accumulation(void);
uint32_t rt_cuda_accum_generation(void);
/* pass index (the RNG key of SURVEY.md R7's per-pixel stream) of the next rt_cuda_update_frame() pass;
 * it keeps counting across invalidations, so a pose revisited draws fresh samples */
uint64_t rt_cuda_next_pass_index(void);
int      rt_cuda_update_frame(const RtCamera *cam, void *fb, int w, int h, double budget_ms,
                              const RtRenderOpts *opts, RtRenderStats *stats);

/* CUDA-OpenGL interop presenter: replaces the per-frame host
 * upload of move_frame_to_the_gpu() (gpu_and_windowing.c:371-376,
 * glTexImage2D(GL_RGB, GL_FLOAT, host pointer)).  The application creates one
 * GL_PIXEL_UNPACK_BUFFER of w*h*12 bytes (RT_FB_F32X3) or w*h*4 (RT_FB_U8X4) in
 * the thread that owns the GL context and registers it once;
 * rt_cuda_gl_update_frame() then maps the buffer, runs rt_cuda_update_frame()
 * with the mapped device pointer as the frame and unmaps it, after which
 * glTexImage2D(..., 0) with the buffer bound sources the texture from device
 * memory: the frame never visits the host.  Must be called with the GL context
 * current; fails with RT_ERR_CUDA (and the CUDA error text) when there is none.
 * GL names are plain unsigned ints (GLuint); no GL header is needed here. */
int rt_cuda_gl_register_buffer(unsigned int gl_buffer, size_t bytes);
int rt_cuda_gl_update_frame(const RtCamera *cam, int w, int h, double budget_ms,
                            const RtRenderOpts *opts, RtRenderStats *stats);
/* same for a single explicit pass (render_frame_cuda_ex into the mapped buffer) */
int rt_cuda_gl_render_frame(const RtCamera *cam, int w, int h, const RtRenderOpts *opts, RtRenderStats *stats);
int rt_cuda_gl_unregister_buffer(void);

/* Rays traced on GPU 0 since the last call that returned RtRenderStats (such calls reset the
 * counter); waits for the device.  Counts a run of asynchronous launches exactly. */
int rt_cuda_ray_counter(uint64_t *rays);

/* Block until all work issued by the library has finished. */
int rt_cuda_synchronize(void);

/* ------------------------------------------------------------------------- */
/* Unit-level device probes (used by the parity tests; thin wrappers over the */
/* same device functions the render kernels inline)                          */
/* ------------------------------------------------------------------------- */
/* rays6: n x (origin xyz, direction xyz); out7: n x (distance, point, normal) */
int rt_cuda_debug_trace(const float *rays6, int n, float *out7, int32_t *obj, int variant, int traversal);
int rt_cuda_debug_sample_cubemap(const float *dirs3, int n, float *out3);
int rt_cuda_debug_camera_rays(const RtCamera *cam, const float *pxpy, int n, float aspect, float *rays6);
int rt_cuda_debug_rng(uint64_t state, int n, uint64_t *u64_out, float *f32_out);
int rt_cuda_debug_random_directions(uint64_t state, int n, float *out3);
uint64_t rt_pixel_key(float px, float py, uint64_t pass_index);
/* Register-only FP32 issue-rate probe: TFLOP/s of FMA chains (fma=1) or of
 * MUL+ADD pairs (fma=0, the ceiling of the no-contraction exact build). */
int rt_cuda_debug_fp32_peak(int fma, float *tflops_out);
/* LBVH scenes, library built with -DRT_COUNT_WALK only (zeros otherwise): internal nodes
 * visited and primitives tested by the last call that returned statistics. */
int rt_cuda_debug_walk_counts(uint64_t *nodes, uint64_t *tests);
/* Bytes of kernel arguments (camera frame, views, sizes) sent host -> device per launch. */
size_t rt_cuda_param_bytes(void);
/* Test knob: tau^2 of the sign shortcut in the light-sample sweep (negative =
 * default 4e-12).  1e30 forces the literal path; frames must be identical. */
int rt_cuda_debug_set_sweep_threshold(float tau2);
/* Test knob: 0 = RT_KERNEL_QUEUED keeps tiles in image order; 1 (default) = a pose rendered
 * repeatedly is scheduled longest tiles first from the costs its previous pass recorded.
 * Scheduling only: frames must be identical either way.  2 = as 1, and a pass whose own scale has
 * no costs yet is ordered by the costs a coarser pass of the pose recorded (measured: the first
 * scale-1 pass of a new 4K pose takes 2.00 ms seeded by its scale-2 pass, 1.97 ms in image order). */
int rt_cuda_debug_set_tile_schedule(int on);
/* Test knob, LBVH scenes: 1 (default) = when ONE object of the scene emits, a light sample
 * (main.c:186-205) is walked in any-hit mode -- the emitter first, then only until something is
 * accepted in front of it -- because all the sample adds is the emission of its nearest hit
 * (main.c:201-204); 0 = every ray is walked to its nearest hit.  Frames must be identical. */
int rt_cuda_debug_set_light_anyhit(int on);
/* Test knob: 1 (default) = rt_cuda_render_sweep() on one GPU runs the passes of the sweep side by
 * side on separate streams and folds them into the frame in pass order with one resolve kernel;
 * 0 = one pass after the other.  Frames, accumulation and ray counts must be identical. */
int rt_cuda_debug_set_concurrent_sweep(int on);
/* Test knob: a synchronous call with a HOST frame on one GPU renders the frame as `bands` row bands
 * at most (default 4; one band per ~1.5 M pixels) and copies band k to the host while band k+1
 * renders; 1 = render, then copy; a negative count forces exactly that many bands whatever the
 * frame size.  Frames must be identical. */
int rt_cuda_debug_set_sync_bands(int bands);
/* Test knob: 1 (default) = launches of 60 000 tiles or more (about 1080p) over a linear-scan scene use the
 * queued kernel's build with 7 CTAs per SM; 0 = always the 6-CTA build.  Frames must be identical. */
int rt_cuda_debug_set_queued_dense(int on);
/* Unit probe of the longest-tiles-first order: tiles_x*tiles_y tiles ordered by the costs of a map
 * that is 1 << shift times coarser (stable: costly classes first, image order within a class). */
int rt_cuda_debug_tile_order(const uint32_t *cost, int cost_tiles_x, int cost_tiles_y, int shift,
                             int tiles_x, int tiles_y, uint32_t *order_out);
/* Bit-compare the render kernels' hoisted-reciprocal division with IEEE `/` on
 * blocks*256*per_thread operand pairs (exponent ranges given); see rt_device.cuh. */
int rt_cuda_debug_div_check(uint64_t seed, unsigned blocks, unsigned per_thread, int lo_exp_b, int hi_exp_b,
                            int lo_exp_a, int hi_exp_a, uint64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* RT_CUDA_H */
