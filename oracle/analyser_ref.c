/*
 * CPU oracle (float32, "Chromium-faithful" arithmetic) for the frame-producing hot path,
 * and the timed CPU baseline of bench.py.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing under spectrogram_b200/ links, loads or
 * calls this file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs load liboracle.so.
 *
 * PARITY UNPINNED.  The reference (amilajack/spectrogram) has no FFT/window/dB code of
 * its own: it configures a browser AnalyserNode (src/javascripts/UI/player.js:7-11) and
 * polls it (src/javascripts/3D/visualizer.js:346-368).  The arithmetic is the browser's
 * Web Audio engine, unpinned by any lock file and not runnable here.  This file restates
 * the W3C Web Audio API AnalyserNode algorithm with the arithmetic widths Chromium's
 * RealtimeAnalyser uses:
 *   window computed in double, cast to float, multiplied into the float sample;
 *   float32 real FFT; |X| via double hypot, scaled by 1/N in double;
 *   smoothing k*prev + (1-k)*mag in double, stored as float; non-finite -> 0 [SPEC];
 *   dB = 20*log10f(linear) in float; byte = (unsigned char)clamp(255*(dB-min)/(max-min)).
 * It is pinned only by the closed-form known-answer tests in tests/test_oracle_kat.py and
 * by agreement with the float64 numpy oracle (oracle/analyser_oracle.py).
 *
 * The FFT below is written for this repo (Stockham autosort, radix-4/2 for powers of two,
 * generic-radix stages otherwise); it is not Chromium's PFFFT.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define SGO_WINDOW_BLACKMAN 0
#define SGO_WINDOW_HANN 1
#define SGO_WINDOW_RECT 2
#define SGO_WINDOW_CUSTOM 3
#define SGO_OUT_U8 0
#define SGO_OUT_F32_DB 1
#define SGO_OUT_RGBA8 2
#define SGO_OUT_F32_MAG 3
#define SGO_ALIGN_VALID 0
#define SGO_ALIGN_ANALYSER 1

typedef struct sgo_config {
  int32_t n_fft, hop, window, output, align;
  float min_db, max_db, smoothing;
  const float* custom_window; /* n_fft floats when window == CUSTOM */
  const uint32_t* colormap;   /* 256 RGBA8 entries, or NULL for the reference LUT */
} sgo_config;

/* ---------------------------------------------------------------- plan */
typedef struct {
  int n;        /* real length */
  int m;        /* complex length n/2 */
  int nstage;
  int radix[32];
  float* tw_re; /* W_m^k, k < m */
  float* tw_im;
  float* ut_re; /* W_n^k, k <= m/2 */
  float* ut_im;
  float* win;   /* float(window) */
} plan_t;

static int factorize(int m, int* radix) {
  int ns = 0;
  while (m % 4 == 0) { radix[ns++] = 4; m /= 4; }
  while (m % 2 == 0) { radix[ns++] = 2; m /= 2; }
  for (int p = 3; p <= m; p += 2)
    while (m % p == 0) { radix[ns++] = p; m /= p; }
  return ns;
}

static void window_table(int kind, int n, const float* custom, float* out) {
  const double two_pi = 6.283185307179586476925286766559;
  for (int i = 0; i < n; ++i) {
    double x = (double)i / (double)n, w;
    switch (kind) {
      case SGO_WINDOW_BLACKMAN: {
        const double alpha = 0.16, a0 = 0.5 * (1 - alpha), a1 = 0.5, a2 = 0.5 * alpha;
        w = a0 - a1 * cos(two_pi * x) + a2 * cos(two_pi * 2.0 * x);
        break;
      }
      case SGO_WINDOW_HANN: w = 0.5 - 0.5 * cos(two_pi * x); break;
      case SGO_WINDOW_CUSTOM: w = custom[i]; break;
      default: w = 1.0;
    }
    out[i] = (float)w;
  }
}

static plan_t* plan_create(const sgo_config* cfg) {
  plan_t* p = (plan_t*)calloc(1, sizeof(plan_t));
  p->n = cfg->n_fft;
  p->m = cfg->n_fft / 2;
  p->nstage = factorize(p->m, p->radix);
  p->tw_re = (float*)malloc(sizeof(float) * p->m);
  p->tw_im = (float*)malloc(sizeof(float) * p->m);
  p->ut_re = (float*)malloc(sizeof(float) * (p->m / 2 + 1));
  p->ut_im = (float*)malloc(sizeof(float) * (p->m / 2 + 1));
  p->win = (float*)malloc(sizeof(float) * p->n);
  const double two_pi = 6.283185307179586476925286766559;
  for (int k = 0; k < p->m; ++k) {
    p->tw_re[k] = (float)cos(-two_pi * k / p->m);
    p->tw_im[k] = (float)sin(-two_pi * k / p->m);
  }
  for (int k = 0; k <= p->m / 2; ++k) {
    p->ut_re[k] = (float)cos(-two_pi * k / p->n);
    p->ut_im[k] = (float)sin(-two_pi * k / p->n);
  }
  window_table(cfg->window, p->n, cfg->custom_window, p->win);
  return p;
}

static void plan_destroy(plan_t* p) {
  free(p->tw_re); free(p->tw_im); free(p->ut_re); free(p->ut_im); free(p->win); free(p);
}

/* ---------------------------------------------------------------- complex FFT (SoA) */
/* One Stockham stage: length m, radix r, ns = product of previous radices.
 * in/out are split re/im arrays of length m. */
static void stage_generic(const plan_t* p, int r, int ns, const float* ir, const float* ii,
                          float* orr, float* oi) {
  const int m = p->m, cnt = m / r, tstep = m / (ns * r);
  float vr[16], vi[16];
  for (int j = 0; j < cnt; ++j) {
    const int k = j % ns;
    for (int q = 0; q < r; ++q) {
      float xr = ir[j + q * cnt], xi = ii[j + q * cnt];
      int t = (q * k * tstep) % m;
      float wr = p->tw_re[t], wi = p->tw_im[t];
      vr[q] = xr * wr - xi * wi;
      vi[q] = xr * wi + xi * wr;
    }
    const int j0 = (j / ns) * ns * r + k;
    const int rstep = m / r;
    for (int q = 0; q < r; ++q) {
      float sr = 0.f, si = 0.f;
      for (int a = 0; a < r; ++a) {
        int t = ((a * q) % r) * rstep;
        float wr = p->tw_re[t], wi = p->tw_im[t];
        sr += vr[a] * wr - vi[a] * wi;
        si += vr[a] * wi + vi[a] * wr;
      }
      orr[j0 + q * ns] = sr;
      oi[j0 + q * ns] = si;
    }
  }
}

static void stage_radix2(const plan_t* p, int ns, const float* ir, const float* ii, float* orr,
                         float* oi) {
  const int m = p->m, cnt = m / 2, tstep = m / (ns * 2);
  for (int jb = 0; jb < cnt; jb += ns) {
    const int j0 = jb * 2;
    for (int k = 0; k < ns; ++k) {
      const int j = jb + k;
      float ar = ir[j], ai = ii[j];
      float xr = ir[j + cnt], xi = ii[j + cnt];
      float wr = p->tw_re[k * tstep], wi = p->tw_im[k * tstep];
      float br = xr * wr - xi * wi, bi = xr * wi + xi * wr;
      orr[j0 + k] = ar + br; oi[j0 + k] = ai + bi;
      orr[j0 + k + ns] = ar - br; oi[j0 + k + ns] = ai - bi;
    }
  }
}

static void stage_radix4(const plan_t* p, int ns, const float* ir, const float* ii, float* orr,
                         float* oi) {
  const int m = p->m, cnt = m / 4, tstep = m / (ns * 4);
  for (int jb = 0; jb < cnt; jb += ns) {
    const int j0 = jb * 4;
    for (int k = 0; k < ns; ++k) {
      const int j = jb + k;
      const int t1 = k * tstep, t2 = 2 * t1, t3 = 3 * t1;
      float ar = ir[j], ai = ii[j];
      float x1r = ir[j + cnt], x1i = ii[j + cnt];
      float x2r = ir[j + 2 * cnt], x2i = ii[j + 2 * cnt];
      float x3r = ir[j + 3 * cnt], x3i = ii[j + 3 * cnt];
      float br = x1r * p->tw_re[t1] - x1i * p->tw_im[t1], bi = x1r * p->tw_im[t1] + x1i * p->tw_re[t1];
      float cr = x2r * p->tw_re[t2] - x2i * p->tw_im[t2], ci = x2r * p->tw_im[t2] + x2i * p->tw_re[t2];
      float dr = x3r * p->tw_re[t3] - x3i * p->tw_im[t3], di = x3r * p->tw_im[t3] + x3i * p->tw_re[t3];
      float s0r = ar + cr, s0i = ai + ci, s1r = ar - cr, s1i = ai - ci;
      float s2r = br + dr, s2i = bi + di, s3r = br - dr, s3i = bi - di;
      orr[j0 + k] = s0r + s2r;          oi[j0 + k] = s0i + s2i;
      orr[j0 + k + ns] = s1r + s3i;     oi[j0 + k + ns] = s1i - s3r;      /* -i * s3 */
      orr[j0 + k + 2 * ns] = s0r - s2r; oi[j0 + k + 2 * ns] = s0i - s2i;
      orr[j0 + k + 3 * ns] = s1r - s3i; oi[j0 + k + 3 * ns] = s1i + s3r;
    }
  }
}

/* complex FFT of length m; data in (ar, ai); scratch (br, bi).  Returns which buffer holds
 * the result (0 = a, 1 = b). */
static int cfft(const plan_t* p, float* ar, float* ai, float* br, float* bi) {
  int ns = 1, flip = 0;
  for (int s = 0; s < p->nstage; ++s) {
    const int r = p->radix[s];
    const float* ir = flip ? br : ar; const float* ii = flip ? bi : ai;
    float* orr = flip ? ar : br; float* oi = flip ? ai : bi;
    if (r == 4) stage_radix4(p, ns, ir, ii, orr, oi);
    else if (r == 2) stage_radix2(p, ns, ir, ii, orr, oi);
    else stage_generic(p, r, ns, ir, ii, orr, oi);
    ns *= r;
    flip ^= 1;
  }
  return flip;
}

/* One frame: block[n] (already the time-domain block) -> re[k], im[k], k < n/2 (unscaled DFT). */
static void frame_fft(const plan_t* p, const float* block, float* work /* 4*m */, float* xre,
                      float* xim) {
  const int m = p->m;
  float* ar = work; float* ai = work + m; float* br = work + 2 * m; float* bi = work + 3 * m;
  for (int i = 0; i < m; ++i) {
    ar[i] = block[2 * i] * p->win[2 * i];
    ai[i] = block[2 * i + 1] * p->win[2 * i + 1];
  }
  const int flip = cfft(p, ar, ai, br, bi);
  const float* zr = flip ? br : ar; const float* zi = flip ? bi : ai;
  /* real-input untangle: X[k] = E[k] + W_n^k O[k] */
  xre[0] = zr[0] + zi[0];
  xim[0] = 0.f; /* "blow away the packed nyquist component" */
  for (int k = 1; k <= m / 2; ++k) {
    const int mk = m - k;
    float er = 0.5f * (zr[k] + zr[mk]), ei = 0.5f * (zi[k] - zi[mk]);
    float orr = 0.5f * (zi[k] + zi[mk]), oi = 0.5f * (zr[mk] - zr[k]);
    float wr = p->ut_re[k], wi = p->ut_im[k];
    float tr = orr * wr - oi * wi, ti = orr * wi + oi * wr;
    xre[k] = er + tr; xim[k] = ei + ti;
    xre[mk] = er - tr; xim[mk] = -(ei - ti);
  }
}

/* ---------------------------------------------------------------- colour map */
static float hsv_channel(int c, float hue) {
  float hd = hue / 60.0f;
  float x = 1.0f - fabsf(fmodf(hd, 2.0f) - 1.0f);
  float r = 0, g = 0, b = 0;
  if (hd < 1.0f) { r = 1; g = x; }
  else if (hd < 2.0f) { r = x; g = 1; }
  else if (hd < 3.0f) { g = 1; b = x; }
  else if (hd < 4.0f) { g = x; b = 1; }
  else if (hd < 5.0f) { r = x; b = 1; }
  else if (hd < 6.0f) { r = 1; b = x; }
  return c == 0 ? r : (c == 1 ? g : b);
}

void sgo_colormap_lut(uint32_t* lut /* 256 */) {
  for (int b = 0; b < 256; ++b) {
    double a = b / 255.0;
    double hue = 360.0 - a * 360.0;
    uint32_t px = 0xFF000000u;
    for (int c = 0; c < 3; ++c) {
      double v = 0.08 + a * (double)hsv_channel(c, (float)hue);
      if (v < 0) v = 0; if (v > 1) v = 1;
      px |= ((uint32_t)floor(v * 255.0 + 0.5)) << (8 * c);
    }
    lut[b] = px;
  }
}

/* ---------------------------------------------------------------- framing + epilogue */
int64_t sgo_num_frames(const sgo_config* cfg, int64_t clip_len) {
  if (cfg->align == SGO_ALIGN_VALID)
    return clip_len < cfg->n_fft ? 0 : 1 + (clip_len - cfg->n_fft) / cfg->hop;
  return clip_len / cfg->hop;
}

static void gather_block(const sgo_config* cfg, const float* clip, int64_t clip_len, int64_t t,
                         float* block) {
  const int n = cfg->n_fft;
  int64_t start = cfg->align == SGO_ALIGN_VALID ? t * cfg->hop : (t + 1) * (int64_t)cfg->hop - n;
  for (int i = 0; i < n; ++i) {
    int64_t s = start + i;
    block[i] = (s >= 0 && s < clip_len) ? clip[s] : 0.f;
  }
}

static inline float smooth_step(double k, float prev, double mag) {
  float v = (float)(k * (double)prev + (1.0 - k) * mag);
  if (!isfinite(v)) v = 0.f;
  return v;
}

static inline unsigned char db_to_byte(float db, double min_db, double range_scale) {
  double scaled = 255.0 * ((double)db - min_db) * range_scale;
  if (!(scaled > 0)) scaled = 0; /* also catches NaN */
  if (scaled > 255.0) scaled = 255.0;
  return (unsigned char)scaled;
}

static void emit_row(const sgo_config* cfg, const uint32_t* lut, const float* lin, int bins,
                     void* out_row) {
  const double range_scale = 1.0 / ((double)cfg->max_db - (double)cfg->min_db);
  switch (cfg->output) {
    case SGO_OUT_F32_MAG: memcpy(out_row, lin, sizeof(float) * bins); break;
    case SGO_OUT_F32_DB: {
      float* o = (float*)out_row;
      for (int k = 0; k < bins; ++k) o[k] = 20.0f * log10f(lin[k]);
      break;
    }
    case SGO_OUT_U8: {
      unsigned char* o = (unsigned char*)out_row;
      for (int k = 0; k < bins; ++k) o[k] = db_to_byte(20.0f * log10f(lin[k]), cfg->min_db, range_scale);
      break;
    }
    default: {
      uint32_t* o = (uint32_t*)out_row;
      for (int k = 0; k < bins; ++k) o[k] = lut[db_to_byte(20.0f * log10f(lin[k]), cfg->min_db, range_scale)];
    }
  }
}

static size_t elem_bytes(int output) { return output == SGO_OUT_U8 ? 1 : 4; }

int sgo_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (int)n;
}

/* Work is split into independent units: single frames when tau == 0 (every frame is
 * independent), whole clips otherwise (the recurrence runs along a clip's frames). */
typedef struct {
  const float* pcm; int64_t n_clips, clip_len, frames; const sgo_config* cfg; void* out;
  const plan_t* plan; const uint32_t* lut; int64_t lo, hi; /* unit range of this worker */
} job_t;

static void* worker(void* arg) {
  const job_t* j = (const job_t*)arg;
  const sgo_config* cfg = j->cfg; const plan_t* p = j->plan;
  const int n = cfg->n_fft, bins = n / 2;
  const size_t row = elem_bytes(cfg->output) * (size_t)bins;
  const double k = cfg->smoothing, mag_scale = 1.0 / (double)n;
  float* block = (float*)malloc(sizeof(float) * n);
  float* work = (float*)malloc(sizeof(float) * 4 * p->m);
  float* xre = (float*)malloc(sizeof(float) * (bins + 1));
  float* xim = (float*)malloc(sizeof(float) * (bins + 1));
  float* state = (float*)malloc(sizeof(float) * bins);
  if (k == 0.0) {
    for (int64_t f = j->lo; f < j->hi; ++f) {
      const int64_t c = f / j->frames, t = f % j->frames;
      gather_block(cfg, j->pcm + c * j->clip_len, j->clip_len, t, block);
      frame_fft(p, block, work, xre, xim);
      for (int b = 0; b < bins; ++b)
        state[b] = smooth_step(0.0, 0.f, hypot((double)xre[b], (double)xim[b]) * mag_scale);
      emit_row(cfg, j->lut, state, bins, (char*)j->out + (size_t)f * row);
    }
  } else {
    for (int64_t c = j->lo; c < j->hi; ++c) {
      memset(state, 0, sizeof(float) * bins);
      for (int64_t t = 0; t < j->frames; ++t) {
        gather_block(cfg, j->pcm + c * j->clip_len, j->clip_len, t, block);
        frame_fft(p, block, work, xre, xim);
        for (int b = 0; b < bins; ++b)
          state[b] = smooth_step(k, state[b], hypot((double)xre[b], (double)xim[b]) * mag_scale);
        emit_row(cfg, j->lut, state, bins, (char*)j->out + (size_t)(c * j->frames + t) * row);
      }
    }
  }
  free(block); free(work); free(xre); free(xim); free(state);
  return NULL;
}

/* Whole path.  pcm: [n_clips][clip_len] float32.  out: [n_clips][frames][bins] of the output
 * element type.  n_threads <= 0 -> all cores.  Returns 0, or -1 on a bad argument. */
int sgo_stft_batch(const float* pcm, int64_t n_clips, int64_t clip_len, const sgo_config* cfg,
                   void* out, int n_threads) {
  if (!pcm || !cfg || !out || cfg->n_fft < 4 || (cfg->n_fft & 1) || cfg->hop < 1 ||
      !(cfg->min_db < cfg->max_db) || !(cfg->smoothing >= 0.f && cfg->smoothing <= 1.f))
    return -1;
  plan_t* p = plan_create(cfg);
  for (int s = 0; s < p->nstage; ++s)
    if (p->radix[s] > 16) { plan_destroy(p); return -1; }
  const int64_t frames = sgo_num_frames(cfg, clip_len);
  uint32_t lut_local[256];
  const uint32_t* lut = cfg->colormap;
  if (!lut) { sgo_colormap_lut(lut_local); lut = lut_local; }
  const int64_t units = cfg->smoothing == 0.f ? n_clips * frames : n_clips;
  int nt = n_threads > 0 ? n_threads : sgo_max_threads();
  if (nt > units) nt = units > 0 ? (int)units : 1;
  if (nt > 1024) nt = 1024;
  job_t* jobs = (job_t*)calloc((size_t)nt, sizeof(job_t));
  pthread_t* th = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
  for (int i = 0; i < nt; ++i) {
    job_t j = {pcm, n_clips, clip_len, frames, cfg, out, p, lut, units * i / nt, units * (i + 1) / nt};
    jobs[i] = j;
  }
  for (int i = 1; i < nt; ++i) pthread_create(&th[i], NULL, worker, &jobs[i]);
  worker(&jobs[0]);
  for (int i = 1; i < nt; ++i) pthread_join(th[i], NULL);
  free(jobs); free(th);
  plan_destroy(p);
  return 0;
}

/* getByteTimeDomainData: b = (unsigned char)clamp(128*(x+1), 0, 255) */
void sgo_time_domain_byte(const float* x, int64_t n, unsigned char* out) {
  for (int64_t i = 0; i < n; ++i) {
    double v = 128.0 * ((double)x[i] + 1.0);
    if (!(v > 0)) v = 0;
    if (v > 255.0) v = 255.0;
    out[i] = (unsigned char)v;
  }
}
