/*
 * ref_driver.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Headless driver over the UNMODIFIED reference translation units of
 * cozis/ray_tracing.  The reference sources are compiled where they lie under
 * /root/reference by oracle/Makefile; only the resulting shared objects land in
 * oracle/_ref/.  Nothing in ray_tracing_b200/ may link or load this.
 *
 * How the reference is reached without editing it:
 *   - main.c is compiled with -Dmain=ref_main so its render_column()/pixel()
 *     (main.c:131-322, both non-static) are callable; they read the globals
 *     `scene`, `skybox` (main.c:54-55) which this driver fills.
 *   - this TU #includes the reference's utils.c and camera.c (found through
 *     -I<ref>/src) to reach the file-static RNG state (utils.c:60) and camera
 *     pose (camera.c:23-35); utils.o / camera.o are then NOT linked separately.
 *   - the "pixel" and "count" variants compile main.c additionally with
 *     -Dray_through_screen_at=hook_rtsa and/or -Dtrace_ray=hook_trace_ray, so
 *     the hooks below run at the top of pixel() (main.c:135) and around every
 *     trace_ray() call made by pixel() (main.c:161,200).
 *
 * Variants (oracle/Makefile):
 *   libref_stream.so  no hooks: the reference as shipped, per-thread RNG stream
 *                     (CPU timing baseline, RNG-free pixel comparisons).
 *   libref_count.so   trace_ray hook only: same image as libref_stream, plus a
 *                     ray counter.
 *   libref_pixel.so   both hooks: RNG state re-keyed per pixel -> partition
 *                     independent image; THE parity target for the CUDA path.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <stdatomic.h>
#include <time.h>

#include "utils.c"              /* reference: wyhash64_x, wyhash64(), random_float(), load_file() */
#include "camera.c"             /* reference: pose statics, ray_through_screen_at(), mutators */
#include "scene.h"              /* reference types: Scene, Object, HitInfo */
#include "gpu_and_windowing.h"  /* reference types: Cubemap; load/sample_cubemap */

/* reference globals and entry points (main.c:50-59, 102-104) */
extern int num_columns;
extern int init_scale;
extern Scene scene;
extern Cubemap skybox;
extern _Atomic uint32_t accum_generation;
Vector3 pixel(float x, float y, float aspect_ratio);
float render_column(Vector3 *data, int scale, int column_w, int column_i,
                    int frame_w, int frame_h, uint64_t cached_generation);

/* ---- per-pixel RNG key (must equal ray_tracing_b200/csrc/rt_rng.h) ---- */
static uint64_t splitmix64(uint64_t z)
{
	z += 0x9e3779b97f4a7c15ull;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}

static uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

static uint64_t pixel_key(float px, float py, uint64_t pass)
{
	uint64_t k = ((uint64_t) f32_bits(px) << 32) | (uint64_t) f32_bits(py);
	return splitmix64(k ^ splitmix64(pass));
}

static int      g_keyed = 0;        /* re-key RNG per pixel? (only effective in libref_pixel) */
static uint64_t g_pass  = 0;
static _Thread_local uint64_t tl_rays = 0; /* trace_ray calls (hooked variants only) */

Ray hook_rtsa(float px, float py, float ar)
{
	if (g_keyed)
		wyhash64_x = pixel_key(px, py, g_pass);
	return ray_through_screen_at(px, py, ar);
}

HitInfo hook_trace_ray(Ray ray, Scene *s)
{
	tl_rays++;
	return trace_ray(ray, s);
}

/* ------------------------------------------------------------------ */
/* scene / skybox / camera plumbing                                   */
/* ------------------------------------------------------------------ */

int refdrv_max_objects(void) { return MAX_OBJECTS; }
size_t refdrv_sizeof_scene(void) { return sizeof(Scene); }
size_t refdrv_sizeof_object(void) { return sizeof(Object); }
void *refdrv_scene_ptr(void) { return &scene; }

int refdrv_parse_scene_file(const char *path)
{
	return parse_scene_file((char *) path, &scene) ? 0 : -1;
}

/* Parse into a caller buffer of refdrv_sizeof_scene() bytes pre-filled by the
 * caller (the reference leaves union/padding bytes of `Object` uninitialised,
 * scene.c:222). */
int refdrv_parse_scene_file_into(const char *path, void *dst)
{
	return parse_scene_file((char *) path, (Scene *) dst) ? 0 : -1;
}

/* objects: n records of sizeof(Object)=68 bytes in the reference layout */
int refdrv_set_scene(const void *objects, int n)
{
	if (n < 0 || n > MAX_OBJECTS) return -1;
	memcpy(scene.objects, objects, (size_t) n * sizeof(Object));
	scene.num_objects = n;
	return 0;
}

static int skybox_owned = 0;

/* Decode the six faces with the reference's own loader (stb_image, as the
 * reference links it).  `dir` must contain right/left/top/bottom/front/back.jpg */
int refdrv_load_skybox(const char *dir)
{
	static char paths[6][4096];
	const char *faces[6];
	const char *names[6];
	names[CF_RIGHT] = "right.jpg"; names[CF_LEFT] = "left.jpg";
	names[CF_TOP] = "top.jpg";     names[CF_BOTTOM] = "bottom.jpg";
	names[CF_FRONT] = "front.jpg"; names[CF_BACK] = "back.jpg";
	for (int i = 0; i < 6; i++) {
		snprintf(paths[i], sizeof(paths[i]), "%s/%s", dir, names[i]);
		FILE *f = fopen(paths[i], "rb");
		if (!f) return -1;       /* load_cubemap would abort() */
		fclose(f);
		faces[i] = paths[i];
	}
	if (skybox_owned) free_cubemap(&skybox);
	load_cubemap(&skybox, faces);
	skybox_owned = 1;
	return 0;
}

/* Point the reference's global skybox at caller-owned face buffers. */
void refdrv_set_skybox(uint8_t *const faces[6], int w, int h, int chan)
{
	if (skybox_owned) { free_cubemap(&skybox); skybox_owned = 0; }
	for (int i = 0; i < 6; i++) skybox.data[i] = faces[i];
	skybox.w = w; skybox.h = h; skybox.chan = chan;
}

void refdrv_get_skybox(uint8_t *faces_out[6], int *w, int *h, int *chan)
{
	for (int i = 0; i < 6; i++) faces_out[i] = skybox.data[i];
	*w = skybox.w; *h = skybox.h; *chan = skybox.chan;
}

void refdrv_set_camera(const float pos[3], const float front[3], const float up[3], float fov_)
{
	camera_pos   = (Vector3) {pos[0], pos[1], pos[2]};
	camera_front = (Vector3) {front[0], front[1], front[2]};
	camera_up    = (Vector3) {up[0], up[1], up[2]};
	fov = fov_;
}

void refdrv_get_camera(float pos[3], float front[3], float up[3], float *fov_)
{
	pos[0] = camera_pos.x; pos[1] = camera_pos.y; pos[2] = camera_pos.z;
	front[0] = camera_front.x; front[1] = camera_front.y; front[2] = camera_front.z;
	up[0] = camera_up.x; up[1] = camera_up.y; up[2] = camera_up.z;
	*fov_ = fov;
}

void refdrv_reset_camera(void)
{
	first_mouse = true; yaw = -90.0f; pitch = 0.0f;
	last_x = 800.0f / 2.0; last_y = 600.0f / 2.0; fov = 30.0f;
	camera_pos   = (Vector3) {5, 5, 5};
	camera_front = (Vector3) {-1, -1, -1};
	camera_up    = (Vector3) {0, 1, 0};
}

void refdrv_move_camera(int dir, float speed) { move_camera((Direction) dir, speed); }
void refdrv_rotate_camera(double mx, double my) { rotate_camera(mx, my); }

/* ------------------------------------------------------------------ */
/* unit-level probes (known-answer tests)                             */
/* ------------------------------------------------------------------ */

void refdrv_rng_seed(uint64_t state) { wyhash64_x = state; }
uint64_t refdrv_rng_state(void) { return wyhash64_x; }
uint64_t refdrv_rng_u64(void) { return wyhash64(); }
float refdrv_random_float(void) { return random_float(); }
void refdrv_random_direction(float out[3])
{
	Vector3 d = random_direction();
	out[0] = d.x; out[1] = d.y; out[2] = d.z;
}
uint64_t refdrv_pixel_key(float px, float py, uint64_t pass) { return pixel_key(px, py, pass); }

void refdrv_camera_ray(float px, float py, float aspect, float out[6])
{
	Ray r = ray_through_screen_at(px, py, aspect);
	out[0] = r.origin.x; out[1] = r.origin.y; out[2] = r.origin.z;
	out[3] = r.direction.x; out[4] = r.direction.y; out[5] = r.direction.z;
}

/* out: distance, point xyz, normal xyz ; returns object index */
int refdrv_trace(const float ray[6], float out[7])
{
	Ray r = {{ray[0], ray[1], ray[2]}, {ray[3], ray[4], ray[5]}};
	HitInfo h = trace_ray(r, &scene);
	out[0] = h.distance;
	out[1] = h.point.x;  out[2] = h.point.y;  out[3] = h.point.z;
	out[4] = h.normal.x; out[5] = h.normal.y; out[6] = h.normal.z;
	return h.object;
}

void refdrv_trace_many(const float *rays, int n, float *out7, int *obj)
{
	for (int i = 0; i < n; i++)
		obj[i] = refdrv_trace(rays + 6 * (size_t) i, out7 + 7 * (size_t) i);
}

void refdrv_sample_cubemap_many(const float *dirs, int n, float *out3)
{
	for (int i = 0; i < n; i++) {
		Vector3 d = {dirs[3*i], dirs[3*i+1], dirs[3*i+2]};
		Vector3 c = sample_cubemap(&skybox, d);
		out3[3*i] = c.x; out3[3*i+1] = c.y; out3[3*i+2] = c.z;
	}
}

/* One pixel() evaluation with the RNG state set explicitly beforehand. */
void refdrv_pixel(float px, float py, float aspect, uint64_t rng_state, float out[3])
{
	int keyed = g_keyed;
	g_keyed = 0;
	wyhash64_x = rng_state;
	Vector3 c = pixel(px, py, aspect);
	g_keyed = keyed;
	out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

/* ------------------------------------------------------------------ */
/* whole-frame pass: T fresh pthreads, one render_column() each        */
/* ------------------------------------------------------------------ */

typedef struct {
	Vector3 *data;
	int scale, column_w, column_i, W, H;
	uint64_t rays;
} ColJob;

static void *column_thread(void *arg)
{
	ColJob *j = (ColJob *) arg;
	tl_rays = 0;
	/* fresh thread => wyhash64_x == 0, exactly like a worker's first pass (main.c:324) */
	render_column(j->data, j->scale, j->column_w, j->column_i, j->W, j->H,
	              (uint64_t) atomic_load(&accum_generation));
	j->rays = tl_rays;
	return NULL;
}

static double now_s(void)
{
	struct timespec ts;
	timespec_get(&ts, TIME_UTC);   /* C11; -std=c11 hides clock_gettime */
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/*
 * Render one pass into out[W*H*3] (row 0 = bottom row, RGB f32 = the
 * reference's Vector3 frame layout).  T = number of columns/threads
 * (reference: --threads).  keyed: re-key RNG per pixel (libref_pixel only).
 * Pixels the reference never writes (rows >= (H/scale)*scale, columns
 * >= (W/T)*T) are left at 0.  Returns wall seconds from first thread create
 * to last join (scene/skybox load excluded); *rays_out = trace_ray calls
 * (0 in the unhooked variant).
 */
double refdrv_render(float *out, int W, int H, int scale, int T, uint64_t pass,
                     int keyed, uint64_t *rays_out)
{
	if (T < 1) T = 1;
	int column_w = W / T;
	ColJob *jobs = calloc((size_t) T, sizeof(ColJob));
	pthread_t *th = calloc((size_t) T, sizeof(pthread_t));
	for (int i = 0; i < T; i++) {
		jobs[i].data = calloc((size_t) column_w * H + 1, sizeof(Vector3));
		jobs[i].scale = scale; jobs[i].column_w = column_w; jobs[i].column_i = i;
		jobs[i].W = W; jobs[i].H = H;
	}
	num_columns = T;
	init_scale = scale;
	g_keyed = keyed;
	g_pass = pass;

	double t0 = now_s();
	for (int i = 0; i < T; i++)
		pthread_create(&th[i], NULL, column_thread, &jobs[i]);
	for (int i = 0; i < T; i++)
		pthread_join(th[i], NULL);
	double t1 = now_s();

	uint64_t rays = 0;
	memset(out, 0, sizeof(float) * 3 * (size_t) W * H);
	for (int c = 0; c < T; c++) {
		rays += jobs[c].rays;
		for (int y = 0; y < H; y++)
			memcpy(out + 3 * ((size_t) y * W + (size_t) c * column_w),
			       jobs[c].data + (size_t) y * column_w,
			       sizeof(Vector3) * (size_t) column_w);
		free(jobs[c].data);
	}
	free(jobs); free(th);
	if (rays_out) *rays_out = rays;
	return t1 - t0;
}
