/*
 * fake_napi_host.c -- a minimal C stand-in for the Node runtime, used to exercise
 * spectrogram_b200/js/build/spectrogram.node where Node itself is not installed.
 *
 * It implements the napi_* functions the shim imports (a tiny value model: numbers, strings,
 * objects with named properties, functions, externals, typed arrays, a pending exception), loads the
 * addon with dlopen, calls napi_register_module_v1 and then drives the exported functions the way
 * index.js does.  Test infrastructure only.
 *
 *   fake_napi_host <spectrogram.node> cpu                 boundary checks that need no GPU
 *   fake_napi_host <spectrogram.node> gpu <out.bin>       runs stftBatch + an AnalyserNode on device 0
 *                                                         and writes the bytes for the Python test
 */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../spectrogram_b200/js/node_api_min.h"

/* ---------------------------------------------------------------- value model */
typedef struct prop { char* name; napi_value value; struct prop* next; } prop;
struct napi_value__ {
  napi_valuetype type;
  double num;
  char* str;
  void* ext;
  napi_callback fn;
  void* fn_data;
  prop* props;
  int is_typed;
  int is_array; size_t arr_len; napi_value* arr;     /* a JS array of values */
  napi_typedarray_type ta_type;
  size_t ta_len;
  void* ta_data;
};
struct napi_env__ {
  int pending;
  char kind[16];   /* Error | TypeError | RangeError */
  char code[64];
  char msg[512];
};
struct napi_callback_info__ { size_t argc; napi_value* argv; void* data; };

static napi_value new_value(napi_valuetype t) {
  napi_value v = (napi_value)calloc(1, sizeof(struct napi_value__));
  v->type = t;
  return v;
}
static napi_value num(double d) { napi_value v = new_value(napi_number); v->num = d; return v; }
static napi_value str(const char* s) { napi_value v = new_value(napi_string); v->str = strdup(s); return v; }
static napi_value typed(napi_typedarray_type t, size_t len, void* data) {
  napi_value v = new_value(napi_object);
  v->is_typed = 1; v->ta_type = t; v->ta_len = len; v->ta_data = data;
  return v;
}

static napi_value array_of(size_t n, napi_value* items) {
  napi_value v = new_value(napi_object);
  v->is_array = 1; v->arr_len = n; v->arr = (napi_value*)malloc(sizeof(napi_value) * (n ? n : 1));
  memcpy(v->arr, items, sizeof(napi_value) * n);
  return v;
}

/* ---------------------------------------------------------------- the napi_* surface */
napi_status napi_create_function(napi_env env, const char* name, size_t length, napi_callback cb, void* data, napi_value* result) {
  (void)env; (void)name; (void)length;
  *result = new_value(napi_function);
  (*result)->fn = cb; (*result)->fn_data = data;
  return napi_ok;
}
napi_status napi_set_named_property(napi_env env, napi_value object, const char* name, napi_value value) {
  (void)env;
  prop* p = (prop*)calloc(1, sizeof(prop));
  p->name = strdup(name); p->value = value; p->next = object->props; object->props = p;
  return napi_ok;
}
static prop* find_prop(napi_value o, const char* name) {
  for (prop* p = o->props; p; p = p->next) if (!strcmp(p->name, name)) return p;
  return NULL;
}
napi_status napi_get_named_property(napi_env env, napi_value object, const char* name, napi_value* result) {
  (void)env;
  prop* p = find_prop(object, name);
  *result = p ? p->value : new_value(napi_undefined);
  return napi_ok;
}
napi_status napi_has_named_property(napi_env env, napi_value object, const char* name, bool* result) {
  (void)env;
  *result = find_prop(object, name) != NULL;
  return napi_ok;
}
napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
  (void)env;
  if (argc) {
    size_t cap = *argc;
    for (size_t i = 0; i < cap; ++i) if (argv) argv[i] = i < info->argc ? info->argv[i] : new_value(napi_undefined);
    *argc = info->argc;
  }
  if (this_arg) *this_arg = NULL;
  if (data) *data = info->data;
  return napi_ok;
}
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result) { (void)env; *result = value->type; return napi_ok; }
napi_status napi_get_value_int32(napi_env env, napi_value v, int32_t* r) { (void)env; if (v->type != napi_number) return napi_number_expected; *r = (int32_t)v->num; return napi_ok; }
napi_status napi_get_value_int64(napi_env env, napi_value v, int64_t* r) { (void)env; if (v->type != napi_number) return napi_number_expected; *r = (int64_t)v->num; return napi_ok; }
napi_status napi_get_value_double(napi_env env, napi_value v, double* r) { (void)env; if (v->type != napi_number) return napi_number_expected; *r = v->num; return napi_ok; }
napi_status napi_get_value_string_utf8(napi_env env, napi_value v, char* buf, size_t n, size_t* result) {
  (void)env;
  if (v->type != napi_string) return napi_string_expected;
  size_t len = strlen(v->str);
  if (buf && n) { size_t c = len < n - 1 ? len : n - 1; memcpy(buf, v->str, c); buf[c] = 0; if (result) *result = c; }
  else if (result) *result = len;
  return napi_ok;
}
napi_status napi_get_value_external(napi_env env, napi_value v, void** r) { (void)env; if (v->type != napi_external) return napi_invalid_arg; *r = v->ext; return napi_ok; }
napi_status napi_create_int32(napi_env env, int32_t v, napi_value* r) { (void)env; *r = num(v); return napi_ok; }
napi_status napi_create_int64(napi_env env, int64_t v, napi_value* r) { (void)env; *r = num((double)v); return napi_ok; }
napi_status napi_create_double(napi_env env, double v, napi_value* r) { (void)env; *r = num(v); return napi_ok; }
napi_status napi_create_string_utf8(napi_env env, const char* s, size_t len, napi_value* r) { (void)env; (void)len; *r = str(s); return napi_ok; }
napi_status napi_create_object(napi_env env, napi_value* r) { (void)env; *r = new_value(napi_object); return napi_ok; }
napi_status napi_create_external(napi_env env, void* data, napi_finalize f, void* hint, napi_value* r) {
  (void)env; (void)f; (void)hint;
  *r = new_value(napi_external); (*r)->ext = data;
  return napi_ok;
}
napi_status napi_get_undefined(napi_env env, napi_value* r) { (void)env; *r = new_value(napi_undefined); return napi_ok; }
napi_status napi_is_array(napi_env env, napi_value v, bool* r) { (void)env; *r = v->is_array != 0; return napi_ok; }
napi_status napi_get_array_length(napi_env env, napi_value v, uint32_t* r) { (void)env; if (!v->is_array) return napi_array_expected; *r = (uint32_t)v->arr_len; return napi_ok; }
napi_status napi_get_element(napi_env env, napi_value v, uint32_t i, napi_value* r) { (void)env; if (!v->is_array || i >= v->arr_len) return napi_invalid_arg; *r = v->arr[i]; return napi_ok; }
napi_status napi_is_typedarray(napi_env env, napi_value v, bool* r) { (void)env; *r = v->is_typed != 0; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env env, napi_value v, napi_typedarray_type* type, size_t* length, void** data, napi_value* ab, size_t* off) {
  (void)env;
  if (!v->is_typed) return napi_invalid_arg;
  if (type) *type = v->ta_type;
  if (length) *length = v->ta_len;
  if (data) *data = v->ta_data;
  if (ab) *ab = NULL;
  if (off) *off = 0;
  return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env env, size_t n, void** data, napi_value* r) {
  (void)env;
  *r = new_value(napi_object); (*r)->ta_data = calloc(1, n ? n : 1); *data = (*r)->ta_data;
  return napi_ok;
}
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type t, size_t len, napi_value ab, size_t off, napi_value* r) {
  (void)env;
  *r = typed(t, len, (char*)ab->ta_data + off);
  return napi_ok;
}
static napi_status set_pending(napi_env env, const char* kind, const char* code, const char* msg) {
  env->pending = 1;
  snprintf(env->kind, sizeof env->kind, "%s", kind);
  snprintf(env->code, sizeof env->code, "%s", code ? code : "");
  snprintf(env->msg, sizeof env->msg, "%s", msg ? msg : "");
  return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char* code, const char* msg) { return set_pending(env, "Error", code, msg); }
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg) { return set_pending(env, "TypeError", code, msg); }
napi_status napi_throw_range_error(napi_env env, const char* code, const char* msg) { return set_pending(env, "RangeError", code, msg); }

/* ---------------------------------------------------------------- driver */
static struct napi_env__ g_env;
static napi_value g_exports;
static int g_failures = 0;

static napi_value call(const char* name, size_t argc, napi_value* argv) {
  prop* p = find_prop(g_exports, name);
  if (!p || p->value->type != napi_function) { fprintf(stderr, "no export %s\n", name); exit(2); }
  struct napi_callback_info__ info = {argc, argv, p->value->fn_data};
  g_env.pending = 0;
  return p->value->fn(&g_env, &info);
}
static void expect(int cond, const char* what) {
  printf("%s %s\n", cond ? "ok  " : "FAIL", what);
  if (!cond) { ++g_failures; if (g_env.pending) printf("     pending %s[%s]: %s\n", g_env.kind, g_env.code, g_env.msg); }
}
static int threw(const char* kind, const char* code) {
  return g_env.pending && !strcmp(g_env.kind, kind) && (!code || !strcmp(g_env.code, code));
}
static napi_value options(int fft, int hop, const char* output, const char* align) {
  napi_value o = new_value(napi_object);
  napi_set_named_property(&g_env, o, "fftSize", num(fft));
  napi_set_named_property(&g_env, o, "hop", num(hop));
  if (output) napi_set_named_property(&g_env, o, "output", str(output));
  if (align) napi_set_named_property(&g_env, o, "align", str(align));
  return o;
}

static void cpu_checks(void) {
  const char* names[] = {"deviceCount", "engineCreate", "engineDestroy", "engineLastKernel", "numFrames", "stftBatch", "stftBatchMulti",
                         "colormapReference", "analyserCreate", "analyserDestroy", "analyserSet", "analyserGet",
                         "analyserPush", "getByteFrequencyData", "getFloatFrequencyData", "getByteTimeDomainData",
                         "getFloatTimeDomainData", "streamCreate", "streamPush", "streamDestroy", "ringCreate", "ringAppend",
                         "ringYoffset", "ringView", "ringDestroy", "wavParse", "pcmDecode", "stftPcm"};
  for (size_t i = 0; i < sizeof names / sizeof *names; ++i) {
    prop* p = find_prop(g_exports, names[i]);
    expect(p && p->value->type == napi_function, names[i]);
  }
  napi_value a[6];
  a[0] = options(2048, 512, NULL, NULL); a[1] = num(441000);
  napi_value r = call("numFrames", 2, a);
  expect(!g_env.pending && r->num == 858, "numFrames(default, 441000) == 858");
  a[0] = options(2048, 512, "u8", "analyser");
  r = call("numFrames", 2, a);
  expect(!g_env.pending && r->num == 861, "numFrames(analyser alignment) == 861");
  a[0] = options(2047, 512, NULL, NULL);
  call("numFrames", 2, a);
  expect(threw("RangeError", "IndexSizeError"), "odd fftSize -> RangeError[IndexSizeError]");
  a[0] = options(2048, 0, NULL, NULL);
  call("numFrames", 2, a);
  expect(threw("RangeError", "IndexSizeError"), "hop 0 -> RangeError[IndexSizeError]");
  a[0] = options(2048, 512, "png", NULL);
  call("numFrames", 2, a);
  expect(threw("TypeError", NULL), "unknown output -> TypeError");
  a[0] = num(3);
  call("numFrames", 2, a);
  expect(threw("TypeError", NULL), "options not an object -> TypeError");
  r = call("colormapReference", 0, NULL);
  expect(!g_env.pending && r->is_typed && r->ta_len == 256 && ((uint32_t*)r->ta_data)[255] == 0xFF1414FFu &&
             ((uint32_t*)r->ta_data)[0] == 0xFF141414u, "colormapReference(): byte 255 -> (255,20,20), byte 0 -> (20,20,20)");
  a[0] = num(42);
  call("analyserCreate", 1, a);
  expect(threw("TypeError", NULL), "analyserCreate(non-engine) -> TypeError");
  /* decodeAudioData's header walk (util/util.js:9): a 44-byte canonical header + 3 stereo s16 frames */
  {
    static unsigned char wav[56] = {'R','I','F','F', 48,0,0,0, 'W','A','V','E', 'f','m','t',' ', 16,0,0,0, 1,0, 2,0,
                                    0x44,0xAC,0,0, 0x10,0xB1,2,0, 4,0, 16,0, 'd','a','t','a', 12,0,0,0};
    a[0] = typed(napi_uint8_array, sizeof wav, wav);
    r = call("wavParse", 1, a);
    prop *pf = find_prop(r, "format"), *pc = find_prop(r, "channels"), *ps = find_prop(r, "sampleRate"),
         *pl = find_prop(r, "length"), *po = find_prop(r, "dataOffset");
    expect(!g_env.pending && pf && !strcmp(pf->value->str, "s16") && pc->value->num == 2 && ps->value->num == 44100 &&
               pl->value->num == 3 && po->value->num == 44, "wavParse(stereo s16 header) -> {s16, 2, 44100, 3 frames, offset 44}");
    wav[20] = 85;                                     /* MPEG layer 3 inside RIFF: not decoded here */
    call("wavParse", 1, a);
    expect(threw("TypeError", NULL), "wavParse(compressed WAVE) -> TypeError");
    a[0] = num(1);
    call("wavParse", 1, a);
    expect(threw("TypeError", NULL), "wavParse(non-array) -> TypeError");
  }
  r = call("deviceCount", 0, NULL);
  if (r->num == 0) {
    a[0] = num(0);
    call("engineCreate", 1, a);
    expect(threw("Error", "ERR_NO_CUDA_DEVICE"), "engineCreate without a GPU -> Error[ERR_NO_CUDA_DEVICE] (no CPU fallback)");
  }
}

static void gpu_run(const char* out_path) {
  napi_value a[6];
  a[0] = num(0);
  napi_value eng = call("engineCreate", 1, a);
  expect(!g_env.pending && eng->type == napi_external, "engineCreate(0)");
  /* 1 s chirp, config 1 */
  const int n = 44100, frames = 1 + (n - 2048) / 512, bins = 1024;
  float* pcm = (float*)malloc(sizeof(float) * n);
  for (int i = 0; i < n; ++i) {
    double t = i / 44100.0;
    pcm[i] = (float)(0.5 * sin(2 * M_PI * (20.0 * t + 0.5 * (20000.0 - 20.0) / 1.0 * t * t)));
  }
  unsigned char* out = (unsigned char*)calloc((size_t)frames * bins, 1);
  a[0] = eng; a[1] = typed(napi_float32_array, n, pcm); a[2] = num(1); a[3] = num(n);
  a[4] = options(2048, 512, "u8", "valid"); a[5] = typed(napi_uint8_array, (size_t)frames * bins, out);
  call("stftBatch", 6, a);
  expect(!g_env.pending, "stftBatch(config 1)");
  a[5] = typed(napi_float32_array, (size_t)frames * bins, out);
  call("stftBatch", 6, a);
  expect(threw("TypeError", NULL), "stftBatch with a Float32Array for u8 output -> TypeError");
  /* AnalyserNode through the shim: player.js:7-11 + visualizer.js:352 */
  a[0] = eng;
  napi_value an = call("analyserCreate", 1, a);
  expect(!g_env.pending && an->type == napi_external, "analyserCreate");
  a[0] = an; a[1] = str("fftSize"); a[2] = num(2048);
  call("analyserSet", 3, a);
  a[1] = str("smoothingTimeConstant"); a[2] = num(0);
  call("analyserSet", 3, a);
  expect(!g_env.pending, "analyser.fftSize = 2048; smoothingTimeConstant = 0");
  a[1] = str("fftSize"); a[2] = num(1000);
  call("analyserSet", 3, a);
  expect(threw("RangeError", "IndexSizeError"), "analyser.fftSize = 1000 -> IndexSizeError");
  a[1] = str("frequencyBinCount");
  napi_value r = call("analyserGet", 2, a);
  expect(!g_env.pending && r->num == 1024, "analyser.frequencyBinCount == 1024");
  a[1] = typed(napi_float32_array, 2048, pcm + 512);      /* frame 1 of the clip */
  call("analyserPush", 2, a);
  unsigned char* row = (unsigned char*)calloc(bins, 1);
  a[1] = typed(napi_uint8_array, bins, row);
  call("getByteFrequencyData", 2, a);
  expect(!g_env.pending && memcmp(row, out + bins, bins) == 0, "analyser.getByteFrequencyData == batch frame 1");
  a[0] = an;
  call("analyserDestroy", 1, a);
  /* the same clip as 16-bit stereo (L = R = the chirp): mono mix through stftPcm == stftBatch on the quantised clip */
  {
    short* s16 = (short*)malloc(sizeof(short) * 2 * n);
    float* q = (float*)malloc(sizeof(float) * n);
    float* dec = (float*)malloc(sizeof(float) * n);
    unsigned char *o1 = (unsigned char*)calloc((size_t)frames * bins, 1), *o2 = (unsigned char*)calloc((size_t)frames * bins, 1);
    napi_value a7[7], desc = new_value(napi_object);
    for (int i = 0; i < n; ++i) { s16[2 * i] = s16[2 * i + 1] = (short)lrint(pcm[i] * 32767.0); q[i] = s16[2 * i] / 32768.0f; }
    napi_set_named_property(&g_env, desc, "format", str("s16"));
    napi_set_named_property(&g_env, desc, "channels", num(2));
    a7[0] = eng; a7[1] = typed(napi_uint8_array, sizeof(short) * 2 * n, s16); a7[2] = desc; a7[3] = num(1); a7[4] = num(1);
    a7[5] = typed(napi_float32_array, n, dec);
    call("pcmDecode", 6, a7);
    expect(!g_env.pending && memcmp(dec, q, sizeof(float) * n) == 0, "pcmDecode(s16 stereo, mix) == 0.5(L+R) of v/32768");
    a7[5] = options(2048, 512, "u8", "valid"); a7[6] = typed(napi_uint8_array, (size_t)frames * bins, o1);
    call("stftPcm", 7, a7);
    expect(!g_env.pending, "stftPcm(s16 stereo, mix)");
    a[0] = eng; a[1] = typed(napi_float32_array, n, q); a[2] = num(1); a[3] = num(n);
    a[4] = options(2048, 512, "u8", "valid"); a[5] = typed(napi_uint8_array, (size_t)frames * bins, o2);
    call("stftBatch", 6, a);
    expect(!g_env.pending && memcmp(o1, o2, (size_t)frames * bins) == 0, "stftPcm bytes == stftBatch bytes of the decoded clip");
    a7[1] = typed(napi_uint8_array, sizeof(short) * 2 * n - 1, s16);
    call("stftPcm", 7, a7);
    expect(threw("TypeError", NULL), "stftPcm with a ragged byte count -> TypeError");
  }
  /* clips sharded over two engines inside the library == one engine (bit for bit); array-length checks of the
   * streaming and ring entry points (a short array must be a TypeError, never a write past its end) */
  {
    const int nc = 5, cl = 2048 + 37 * 512, fr = 38;
    float* many = (float*)malloc(sizeof(float) * nc * cl);
    unsigned char *m1 = (unsigned char*)calloc((size_t)nc * fr * bins, 1), *m2 = (unsigned char*)calloc((size_t)nc * fr * bins, 1);
    for (int i = 0; i < nc * cl; ++i) many[i] = (float)(0.3 * sin(0.01 * i + 0.7 * (i / cl)) + 0.1 * sin(1.3 * i));
    a[0] = num(0);
    napi_value eng2 = call("engineCreate", 1, a);
    napi_value pair[2] = {eng, eng2};
    a[0] = array_of(2, pair); a[1] = typed(napi_float32_array, (size_t)nc * cl, many); a[2] = num(nc); a[3] = num(cl);
    a[4] = options(2048, 512, "u8", "valid"); a[5] = typed(napi_uint8_array, (size_t)nc * fr * bins, m2);
    call("stftBatchMulti", 6, a);
    expect(!g_env.pending, "stftBatchMulti([engine, engine2], 5 clips)");
    a[0] = eng; a[5] = typed(napi_uint8_array, (size_t)nc * fr * bins, m1);
    call("stftBatch", 6, a);
    expect(!g_env.pending && memcmp(m1, m2, (size_t)nc * fr * bins) == 0, "stftBatchMulti bytes == stftBatch bytes");
    napi_value same[2] = {eng, eng};
    a[0] = array_of(2, same);
    call("stftBatchMulti", 6, a);
    expect(threw("TypeError", NULL), "stftBatchMulti with one engine listed twice -> TypeError");
    a[0] = array_of(2, pair); a[5] = typed(napi_uint8_array, (size_t)nc * fr * bins - 1, m2);
    call("stftBatchMulti", 6, a);
    expect(threw("TypeError", NULL), "stftBatchMulti with a short out -> TypeError");
    a[0] = eng2;
    call("engineDestroy", 1, a);
    /* streaming bank: 4 channels, n_fft 1024, hop 128 */
    a[0] = eng; a[1] = num(4); a[2] = options(1024, 128, "u8", NULL); a[3] = num(256);
    napi_value st = call("streamCreate", 4, a);
    expect(!g_env.pending && st->type == napi_external, "streamCreate(4 channels)");
    float* chunk = (float*)calloc(4 * 256, sizeof(float));
    unsigned char* so = (unsigned char*)calloc(4 * 2 * 512, 1);
    napi_value a5[5];
    a5[0] = st; a5[1] = typed(napi_float32_array, 4 * 256, chunk); a5[2] = num(256); a5[3] = typed(napi_uint8_array, 4 * 2 * 512, so);
    call("streamPush", 4, a5);
    expect(!g_env.pending, "streamPush(4 x 256 samples)");
    a5[3] = typed(napi_uint8_array, 4 * 2 * 512 - 1, so);
    call("streamPush", 4, a5);
    expect(threw("TypeError", NULL), "streamPush with a short out -> TypeError");
    a5[3] = typed(napi_uint8_array, 4 * 2 * 512, so); a5[1] = typed(napi_float32_array, 4 * 256 - 1, chunk);
    call("streamPush", 4, a5);
    expect(threw("TypeError", NULL), "streamPush with a short chunk -> TypeError");
    a5[1] = typed(napi_float32_array, 4 * 256, chunk); a5[3] = typed(napi_float32_array, 4 * 2 * 512, so);
    call("streamPush", 4, a5);
    expect(threw("TypeError", NULL), "streamPush with a Float32Array for u8 output -> TypeError");
    a5[3] = typed(napi_uint8_array, 4 * 2 * 512, so); a5[2] = num(100);
    call("streamPush", 4, a5);
    expect(threw("TypeError", NULL), "streamPush with chunkLen not a multiple of hop -> TypeError");
    a5[2] = num(256); a5[4] = typed(napi_uint32_array, 4 * 2 * 512 - 1, so);
    call("streamPush", 5, a5);
    expect(threw("TypeError", NULL), "streamPush with a short outRgba -> TypeError");
    a[0] = st;
    call("streamDestroy", 1, a);
    a[0] = eng; a[1] = num(512); a[2] = num(256);
    napi_value ring = call("ringCreate", 3, a);
    expect(!g_env.pending && ring->type == napi_external, "ringCreate(512 bins, 256 rows)");
    a[0] = ring; a[1] = typed(napi_uint8_array, 2 * 512, so); a[2] = num(2);
    call("ringAppend", 3, a);
    expect(!g_env.pending, "ringAppend(2 rows)");
    a[1] = typed(napi_uint8_array, 2 * 512 - 1, so);
    call("ringAppend", 3, a);
    expect(threw("TypeError", NULL), "ringAppend with frames shorter than nRows * bins -> TypeError");
    a[0] = ring;
    call("ringDestroy", 1, a);
  }
  FILE* f = fopen(out_path, "wb");
  fwrite(out, 1, (size_t)frames * bins, f);
  fclose(f);
  a[0] = eng;
  call("engineDestroy", 1, a);
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s spectrogram.node cpu|gpu [out.bin]\n", argv[0]); return 2; }
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  napi_value (*init)(napi_env, napi_value) = (napi_value(*)(napi_env, napi_value))dlsym(h, "napi_register_module_v1");
  if (!init) { fprintf(stderr, "addon does not export napi_register_module_v1\n"); return 2; }
  g_exports = new_value(napi_object);
  g_exports = init(&g_env, g_exports);
  cpu_checks();
  if (!strcmp(argv[2], "gpu")) gpu_run(argc > 3 ? argv[3] : "/tmp/napi_out.bin");
  printf("%d failure(s)\n", g_failures);
  return g_failures ? 1 : 0;
}
