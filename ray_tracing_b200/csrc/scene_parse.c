/*
 * scene_parse.c -- scene-file loader with the grammar and float construction of
 * the reference's parser (reference: src/scene.c:193-624).
 *
 * Written from the grammar, not from the code: the reference matches each
 * keyword with an unrolled character chain; here one table drives it.  What
 * must stay identical, and is covered by tests/test_scene_parser.py against the
 * compiled reference:
 *   - object keywords `sphere` / `cube` matched by prefix, defaults
 *     (scene.c:231-254; double literals narrowed to float);
 *   - property keywords matched by prefix, with the reference's cursor advance:
 *     `albedo` skips 9 characters and `metallic` 11 regardless of what follows
 *     (scene.c:280,320), so files need >= 3 blanks after those two keywords;
 *   - numbers are `-?digits[.digits]`, built as v = v*10 + d, then
 *     v += q*d with q = 1.0f/10, q /= 10 in binary32 (scene.c:441-461) -- not strtof;
 *   - range checks and messages (scene.c:530-599); objects beyond the capacity
 *     are dropped with a warning (scene.c:602-605);
 *   - on failure num_objects keeps the count parsed so far (scene.c:208).
 */
#include "rt_cuda.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { VAL_SCALAR, VAL_VEC3 };
enum { ANY = -1 };
enum {
	P_ALBEDO, P_ROUGHNESS, P_REFLECTANCE, P_METALLIC, P_EMISSION_POWER,
	P_EMISSION_COLOR, P_RADIUS, P_CENTER, P_ORIGIN, P_SIZE
};

typedef struct {
	const char *word;     /* characters compared */
	int         need;     /* characters that must remain for the keyword to be considered */
	int         skip;     /* cursor advance after a match (the reference's quirk lives here) */
	int         value;    /* VAL_* */
	int         only;     /* object type this property is restricted to, or ANY */
	int         id;
} PropSpec;

/* order matters: first match wins, as in the reference's if/else chain */
static const PropSpec PROPS[] = {
	{"albedo",         7,  9,  VAL_VEC3,   ANY,              P_ALBEDO},
	{"roughness",      9,  9,  VAL_SCALAR, ANY,              P_ROUGHNESS},
	{"reflectance",    11, 11, VAL_SCALAR, ANY,              P_REFLECTANCE},
	{"metallic",       8,  11, VAL_SCALAR, ANY,              P_METALLIC},
	{"emission_power", 14, 14, VAL_SCALAR, ANY,              P_EMISSION_POWER},
	{"emission_color", 14, 14, VAL_VEC3,   ANY,              P_EMISSION_COLOR},
	{"radius",         6,  6,  VAL_SCALAR, RT_OBJECT_SPHERE, P_RADIUS},
	{"center",         6,  6,  VAL_VEC3,   RT_OBJECT_SPHERE, P_CENTER},
	{"origin",         6,  6,  VAL_VEC3,   RT_OBJECT_CUBE,   P_ORIGIN},
	{"size",           4,  4,  VAL_VEC3,   RT_OBJECT_CUBE,   P_SIZE},
};
#define NPROPS ((int) (sizeof(PROPS) / sizeof(PROPS[0])))

typedef struct {
	const char *s;
	size_t      n, i;
	int         line;
} Cursor;

static int blank(char c) { return c == ' ' || c == '\r' || c == '\t' || c == '\n'; }   /* utils.h:34 */
static int digit(char c) { return c >= '0' && c <= '9'; }

static void skip_blanks(Cursor *c)
{
	while (c->i < c->n && blank(c->s[c->i])) {
		if (c->s[c->i] == '\n') c->line++;
		c->i++;
	}
}

/* `need` characters must remain: the whole word, except `albedo` whose guard
 * asks for one more (scene.c:271 `6 < len - i` vs scene.c:224 `5 < len - i`). */
static int at_word(const Cursor *c, const char *word, size_t need)
{
	if (c->i >= c->n || c->n - c->i < need) return 0;
	return memcmp(c->s + c->i, word, strlen(word)) == 0;
}

/* -?digits[.digits] with the reference's binary32 accumulation.
 * `what` selects the message for a missing leading digit. */
static int number(Cursor *c, float *out, int vec_index)
{
	int sign = 1;
	char ch = c->i < c->n ? c->s[c->i] : '\0';
	if (ch == '-') {
		sign = -1;
		c->i++;
		if (c->i >= c->n || !digit(c->s[c->i])) {
			fprintf(stderr, "Error: Missing number after minus sign (line %d)\n", c->line);
			return 0;
		}
	} else if (!digit(ch)) {
		if (vec_index < 0)
			fprintf(stderr, "Error: Missing number after property name (line %d)\n", c->line);
		else
			fprintf(stderr, "Error: Missing number %d in vector value (line %d)\n", vec_index, c->line);
		return 0;
	}
	float v = 0;
	do {
		v = v * 10 + (c->s[c->i] - '0');
		c->i++;
	} while (c->i < c->n && digit(c->s[c->i]));
	if (c->i < c->n && c->s[c->i] == '.') {
		c->i++;
		if (c->i >= c->n || !digit(c->s[c->i])) {
			fprintf(stderr, "Error: Missing decimal part after dot (line %d)\n", c->line);
			return 0;
		}
		float q = 1.0f / 10;
		do {
			v += q * (c->s[c->i] - '0');
			q /= 10;
			c->i++;
		} while (c->i < c->n && digit(c->s[c->i]));
	}
	*out = v * sign;
	return 1;
}

static int braced_vec3(Cursor *c, RtVector3 *out)
{
	if (c->s[c->i] != '{') {
		fprintf(stderr, "Error: Missing '{' after property name (line %d)\n", c->line);
		return 0;
	}
	c->i++;
	float t[3];
	for (int j = 0; j < 3; j++) {
		skip_blanks(c);
		if (!number(c, &t[j], j)) return 0;
	}
	skip_blanks(c);
	if (c->i >= c->n || c->s[c->i] != '}') {
		fprintf(stderr, "Error: Missing '}' after property value (line %d)\n", c->line);
		return 0;
	}
	c->i++;
	out->x = t[0]; out->y = t[1]; out->z = t[2];
	return 1;
}

static int unit_range(float f) { return !(f < 0 || f > 1); }
static int unit_range3(RtVector3 v) { return unit_range(v.x) && unit_range(v.y) && unit_range(v.z); }

static void default_material(RtMaterial *m)
{
	m->albedo.x = 0.44; m->albedo.y = 0.68; m->albedo.z = 0.84;   /* double -> float, scene.c:234 */
	m->roughness = 0;
	m->reflectance = 0.2;
	m->metallic = 0;
	m->emission_power = 0;
	m->emission_color.x = 1; m->emission_color.y = 1; m->emission_color.z = 1;
}

/* Parses one object starting at the cursor.  1 = ok, 0 = error (message printed). */
static int one_object(Cursor *c, RtObject *o)
{
	memset(o, 0, sizeof(*o));
	if (at_word(c, "sphere", 6)) {
		o->type = RT_OBJECT_SPHERE;
		o->sphere.radius = 1;
		c->i += 6;
	} else if (at_word(c, "cube", 4)) {
		o->type = RT_OBJECT_CUBE;
		o->cube.size.x = 1; o->cube.size.y = 1; o->cube.size.z = 1;
		c->i += 4;
	} else {
		fprintf(stderr, "Error: Invalid character (line %d)\n", c->line);
		return 0;
	}
	default_material(&o->material);

	for (;;) {
		skip_blanks(c);
		const PropSpec *p = NULL;
		for (int k = 0; k < NPROPS && !p; k++)
			if (at_word(c, PROPS[k].word, (size_t) PROPS[k].need)) p = &PROPS[k];
		if (!p) return 1;                       /* not a property: the object ends here */
		if (p->only != ANY && p->only != (int) o->type) {
			fprintf(stderr, "Poperty '%s' only allowed on %s (line %d)\n", p->word,
			        p->only == RT_OBJECT_SPHERE ? "spheres" : "cubes", c->line);
			return 0;
		}
		c->i += (size_t) p->skip;
		skip_blanks(c);
		if (c->i >= c->n) {
			fprintf(stderr, "Error: Property value is missing (line %d)\n", c->line);
			return 0;
		}

		float     f = 0;
		RtVector3 v = {0, 0, 0};
		if (p->value == VAL_SCALAR ? !number(c, &f, -1) : !braced_vec3(c, &v))
			return 0;

		switch (p->id) {
		case P_ALBEDO:
			if (!unit_range3(v)) {
				fprintf(stderr, "Error: albedo values must be between 0 and 1 (line %d)\n", c->line);
				return 0;
			}
			o->material.albedo = v;
			break;
		case P_ROUGHNESS:
			if (!unit_range(f)) {
				fprintf(stderr, "Error: Roughness must be between 0 and 1 (line %d)\n", c->line);
				return 0;
			}
			o->material.roughness = f;
			break;
		case P_REFLECTANCE:
			if (!unit_range(f)) {
				fprintf(stderr, "Error: Reflectance must be between 0 and 1 (line %d)\n", c->line);
				return 0;
			}
			o->material.reflectance = f;
			break;
		case P_METALLIC:
			if (!unit_range(f)) {
				fprintf(stderr, "Error: Metallic must be between 0 and 1 (line %d)\n", c->line);
				return 0;
			}
			o->material.metallic = f;
			break;
		case P_EMISSION_POWER:
			o->material.emission_power = f;
			break;
		case P_EMISSION_COLOR:
			if (!unit_range3(v)) {
				fprintf(stderr, "Error: Emission color values must be between 0 and 1 (line %d)\n", c->line);
				return 0;
			}
			o->material.emission_color = v;
			break;
		case P_RADIUS: o->sphere.radius = f; break;
		case P_CENTER: o->sphere.center = v; break;
		case P_ORIGIN: o->cube.origin = v; break;
		case P_SIZE:
			if (v.x < 0 || v.y < 0 || v.z < 0) {
				fprintf(stderr, "Error: Size values must be positive (line %d)\n", c->line);
				return 0;
			}
			o->cube.size = v;
			break;
		}
	}
}

/* Sink: fixed-capacity array (reference Scene) or growable heap array. */
typedef struct {
	RtObject *items;
	int       count, cap;
	int       growable;
} Sink;

/* Parse every object in c->s[c->i .. c->n) and append it to the sink. */
static int parse_objects(Cursor *c, Sink *sink)
{
	for (;;) {
		skip_blanks(c);
		if (c->i >= c->n) return 1;
		RtObject o;
		if (!one_object(c, &o)) return 0;
		if (sink->count == sink->cap && sink->growable) {
			int ncap = sink->cap ? sink->cap * 2 : 1024;
			RtObject *p = (RtObject *) realloc(sink->items, (size_t) ncap * sizeof(RtObject));
			if (!p) {
				fprintf(stderr, "Error: out of memory while parsing scene\n");
				return 0;
			}
			memset(p + sink->cap, 0, (size_t) (ncap - sink->cap) * sizeof(RtObject));
			sink->items = p;
			sink->cap = ncap;
		}
		if (sink->count == sink->cap)
			fprintf(stderr, "Warning: Ignoring object because the scene is too big (line %d)\n", c->line);
		else {
			/* assign field-wise what the reference assigns; the sphere's union
			 * tail stays whatever the destination held */
			RtObject *d = &sink->items[sink->count++];
			d->type = o.type;
			if (o.type == RT_OBJECT_SPHERE) d->sphere = o.sphere;
			else d->cube = o.cube;
			d->material = o.material;
		}
	}
}

static int parse_all(const char *src, size_t len, Sink *sink)
{
	Cursor c = {src, len, 0, 1};
	sink->count = 0;
	return parse_objects(&c, sink);
}

/*
 * Streaming variant for large files (the 100 000-sphere scene is
 * 15 MB of text; the reference reads a whole file into memory, utils.c:32-58 +
 * scene.c:611-624).  The file is read through a window of RT_PARSE_WINDOW bytes
 * that is cut at the last object keyword it holds: no property keyword and no
 * number contains "sphere" or "cube", so in a well-formed file a blank followed
 * by one of them starts an object and everything before it is complete objects.
 * The tail moves to the front of the window and the next read appends to it.
 * Line numbers in messages carry across windows.
 */
#define RT_PARSE_WINDOW (1u << 20)

static size_t last_object_start(const char *s, size_t n)
{
	for (size_t i = n; i-- > 1;) {
		if (!blank(s[i - 1])) continue;
		if ((n - i >= 6 && memcmp(s + i, "sphere", 6) == 0) || (n - i >= 4 && memcmp(s + i, "cube", 4) == 0)) return i;
	}
	return 0;
}

static int parse_stream(FILE *f, Sink *sink)
{
	size_t cap = RT_PARSE_WINDOW, have = 0;
	char *buf = (char *) malloc(cap + 1);
	int line = 1, ok = 1;
	sink->count = 0;
	if (!buf) {
		fprintf(stderr, "Error: out of memory while parsing scene\n");
		return 0;
	}
	for (;;) {
		size_t got = fread(buf + have, 1, cap - have, f);
		have += got;
		int eof = got == 0;
		size_t cut = eof ? have : last_object_start(buf, have);
		if (!eof && cut == 0) {
			/* one object larger than the window (or no keyword yet): grow and read on */
			if (have == cap) {
				char *p = (char *) realloc(buf, 2 * cap + 1);
				if (!p) { fprintf(stderr, "Error: out of memory while parsing scene\n"); ok = 0; break; }
				buf = p;
				cap *= 2;
			}
			continue;
		}
		Cursor c = {buf, cut, 0, line};
		ok = parse_objects(&c, sink);
		line = c.line;
		if (!ok || eof) break;
		memmove(buf, buf + cut, have - cut);
		have -= cut;
	}
	free(buf);
	return ok;
}

static char *slurp(const char *file, size_t *len)
{
	FILE *f = fopen(file, "rb");
	if (!f) return NULL;
	if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return NULL; }
	long sz = ftell(f);
	if (sz < 0 || fseek(f, 0, SEEK_SET) != 0) { fclose(f); return NULL; }
	char *buf = (char *) malloc((size_t) sz + 1);
	if (!buf) { fclose(f); return NULL; }
	size_t got = fread(buf, 1, (size_t) sz, f);
	if (ferror(f)) { free(buf); fclose(f); return NULL; }
	fclose(f);
	buf[got] = '\0';
	*len = (size_t) sz;
	return buf;
}

bool rt_parse_scene_string(const char *src, size_t len, RtScene *scene)
{
	Sink sink = {scene->objects, 0, RT_MAX_OBJECTS, 0};
	int ok = parse_all(src, len, &sink);
	scene->num_objects = sink.count;
	return ok != 0;
}

bool rt_parse_scene_file(const char *file, RtScene *scene)
{
	size_t len = 0;
	char *src = slurp(file, &len);
	if (!src) {
		fprintf(stderr, "Error: Couldn't open scene file\n");
		return false;
	}
	bool ok = rt_parse_scene_string(src, len, scene);
	free(src);
	return ok;
}

int rt_parse_scene_string_large(const char *src, size_t len, RtObject **objects, int *num_objects)
{
	Sink sink = {NULL, 0, 0, 1};
	int ok = parse_all(src, len, &sink);
	if (!ok) {
		free(sink.items);
		*objects = NULL;
		*num_objects = 0;
		return RT_ERR_PARSE;
	}
	*objects = sink.items;
	*num_objects = sink.count;
	return RT_OK;
}

int rt_parse_scene_file_large(const char *file, RtObject **objects, int *num_objects)
{
	*objects = NULL;
	*num_objects = 0;
	FILE *f = fopen(file, "rb");
	if (!f) {
		fprintf(stderr, "Error: Couldn't open scene file\n");
		return RT_ERR_IO;
	}
	Sink sink = {NULL, 0, 0, 1};
	int ok = parse_stream(f, &sink);
	int io_error = ferror(f);
	fclose(f);
	if (!ok || io_error) {
		free(sink.items);
		return io_error ? RT_ERR_IO : RT_ERR_PARSE;
	}
	*objects = sink.items;
	*num_objects = sink.count;
	return RT_OK;
}

void rt_free_objects(RtObject *objects) { free(objects); }
