// AnalyserNode-shaped object and the multi-channel streaming object (included by sgcore.cu).
//
// sg_analyser mirrors the object the reference builds at src/javascripts/UI/player.js:7-11 and
// polls at src/javascripts/3D/visualizer.js:346-368.  The browser's audio render thread is
// replaced by sg_analyser_push().  All arithmetic runs on the GPU; the host side only keeps the
// sample ring (the browser keeps it on the audio thread's side too).

namespace {
constexpr int kRing = 32768;  // AnalyserNode's maximum fftSize; the ring holds that many samples

template <int OUT>
int emit_state(sg_engine* e, const float* state, void* out, long long n, const sg::Epilogue& ep, cudaStream_t st) {
  using T = typename sg::OutElem<OUT>::type;
  sg::emit_state_kernel<OUT><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(state, (T*)out, n, ep);
  e->launches++;
  SG_CUDA(cudaGetLastError());
  return SG_OK;
}
}  // namespace

struct sg_analyser {
  sg_engine* e = nullptr;
  int fft_size = 2048;
  double min_db = -100.0, max_db = -30.0, tau = 0.8;
  float* ring = nullptr;  // pinned host, kRing samples
  size_t write = 0;
  bool dirty = true;      // samples pushed (or attributes changed) since the last analysis
  PinBuf h_block, h_out;
  DevBuf d_block, d_mag, d_state, d_out;

  sg_stft_config cfg() const {
    sg_stft_config c;
    sg_stft_config_default(&c);
    c.n_fft = fft_size; c.hop = fft_size; c.output = SG_OUT_F32_MAG;
    c.min_db = (float)min_db; c.max_db = (float)max_db; c.smoothing = (float)tau;
    return c;
  }

  void linearise(float* dst, int n) const {  // the most recent n samples, oldest first
    for (int i = 0; i < n; ++i) dst[i] = ring[(write + kRing - n + i) % kRing];
  }

  int reset_state() {
    SG_TRY(d_state.reserve(sizeof(float) * (kRing / 2)));
    SG_CUDA(cudaMemsetAsync(d_state.p, 0, sizeof(float) * (kRing / 2), e->stream));
    return SG_OK;
  }

  // steps 1-4: block -> window -> FFT -> |X|/N -> smoothing; leaves X^ in d_state
  int analyse() {
    if (!dirty) return SG_OK;
    const int n = fft_size, bins = n / 2;
    SG_TRY(h_block.reserve(sizeof(float) * n));
    SG_TRY(d_block.reserve(sizeof(float) * n));
    SG_TRY(d_mag.reserve(sizeof(float) * 2 * bins));
    linearise((float*)h_block.p, n);
    SG_CUDA(cudaMemcpyAsync(d_block.p, h_block.p, sizeof(float) * n, cudaMemcpyHostToDevice, e->stream));
    const sg_stft_config c = cfg();
    Plan* pl;
    SG_TRY(e->get_plan(c, &pl));
    sg::FrameGeom g{(const float*)d_block.p, n, n, 1, 1, 0, n, n};
    SG_TRY(e->scratch_acquire(e->stream));
    SG_TRY(launch_frames(e, *pl, g, c, SG_OUT_F32_MAG, e->lut_ref, d_mag.p, e->stream));
    const sg::Epilogue ep = make_epilogue(c, 2.0 * n, e->lut_ref);
    SG_TRY(launch_smooth(e, SG_OUT_F32_MAG, (const float*)d_mag.p, (float*)d_mag.p + bins, (float*)d_state.p, 1, 1, bins, tau, ep,
                         e->stream));
    dirty = false;
    return SG_OK;
  }

  template <int OUT>
  int read_frequency(void* dst, int64_t len) {
    if (len < 0 || (len > 0 && !dst)) return fail(SG_ERR_INVALID_ARG, "bad destination array");
    std::lock_guard<std::mutex> lock(e->mu);
    SG_CUDA(cudaSetDevice(e->device));
    SG_TRY(analyse());
    const int bins = fft_size / 2;
    const int64_t n = std::min<int64_t>(len, bins);
    if (n == 0) return SG_OK;
    const size_t eb = sizeof(typename sg::OutElem<OUT>::type);
    SG_TRY(d_out.reserve(eb * bins));
    SG_TRY(h_out.reserve(eb * bins));
    const sg_stft_config c = cfg();
    const sg::Epilogue ep = make_epilogue(c, 2.0 * fft_size, e->lut_ref);
    SG_TRY(emit_state<OUT>(e, (const float*)d_state.p, d_out.p, n, ep, e->stream));
    SG_CUDA(cudaMemcpyAsync(h_out.p, d_out.p, eb * n, cudaMemcpyDeviceToHost, e->stream));
    SG_CUDA(cudaStreamSynchronize(e->stream));
    std::memcpy(dst, h_out.p, eb * n);
    return SG_OK;
  }
};

struct sg_stream {
  sg_engine* e = nullptr;
  int channels = 0, max_chunk = 0;
  sg_stft_config cfg;
  std::vector<float> custom_window;
  std::vector<uint32_t> colormap;
  long long pitch = 0;  // floats per channel row: n_fft history + max_chunk new samples
  DevBuf hist[2], d_out, d_rgba, d_state;
  PinBuf h_in, h_out, h_rgba;
  int cur = 0;
  int64_t frames_emitted = 0;
  // Off the critical path of a push: the history slide runs on a side stream beside the frame kernel, and the byte
  // rows leave on the copy stream while the colour LUT kernel runs.
  cudaEvent_t ev_in = nullptr, ev_frames = nullptr, ev_slide = nullptr, ev_bytes = nullptr;
  bool slide_pending = false;
  const void* pinned_seen[3] = {nullptr, nullptr, nullptr};   // caller buffers already known to be page-locked
  bool pinned(const void* p, int slot) {
    if (p == pinned_seen[slot]) return true;
    if (!is_pinned(p)) return false;
    pinned_seen[slot] = p;
    return true;
  }
};

struct sg_ring {
  sg_engine* e = nullptr;
  int bins = 0, rows = 0, yoffset = 0;
  DevBuf tex, img;
  PinBuf h_img, h_rows;
};

extern "C" {

// ------------------------------------------------------------------------------------------
// AnalyserNode
// ------------------------------------------------------------------------------------------
int sg_analyser_create(sg_engine* e, sg_analyser** out) {
  if (!e || !out) return fail(SG_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  std::lock_guard<std::mutex> lock(e->mu);
  SG_CUDA(cudaSetDevice(e->device));
  sg_analyser* a = new sg_analyser();
  a->e = e;
  cudaError_t ce = cudaHostAlloc((void**)&a->ring, sizeof(float) * kRing, cudaHostAllocDefault);
  if (ce != cudaSuccess) { delete a; return fail(SG_ERR_OOM, cudaGetErrorString(ce)); }
  std::memset(a->ring, 0, sizeof(float) * kRing);  // the ring starts zero filled
  int rc = a->reset_state();
  if (rc != SG_OK) { cudaFreeHost(a->ring); delete a; return rc; }
  *out = a;
  return SG_OK;
}

int sg_analyser_destroy(sg_analyser* a) {
  if (!a) return SG_OK;
  {
    std::lock_guard<std::mutex> lock(a->e->mu);
    cudaSetDevice(a->e->device);
    cudaStreamSynchronize(a->e->stream);
    cudaFreeHost(a->ring);
    a->h_block.release(); a->h_out.release();
    a->d_block.release(); a->d_mag.release(); a->d_state.release(); a->d_out.release();
  }
  delete a;
  return SG_OK;
}

int sg_analyser_set_fft_size(sg_analyser* a, int fft_size) {
  if (!a) return fail(SG_ERR_INVALID_ARG, "analyser is null");
  if (fft_size < 32 || fft_size > 32768 || (fft_size & (fft_size - 1)))
    return fail(SG_ERR_INDEX_SIZE, "fftSize must be a power of two in [32, 32768]");
  if (fft_size != a->fft_size) {
    std::lock_guard<std::mutex> lock(a->e->mu);
    SG_CUDA(cudaSetDevice(a->e->device));
    a->fft_size = fft_size;
    SG_TRY(a->reset_state());  // a new magnitude buffer starts at zero
    a->dirty = true;
  }
  return SG_OK;
}
int sg_analyser_get_fft_size(const sg_analyser* a) { return a ? a->fft_size : SG_ERR_INVALID_ARG; }
int sg_analyser_get_frequency_bin_count(const sg_analyser* a) { return a ? a->fft_size / 2 : SG_ERR_INVALID_ARG; }

int sg_analyser_set_min_decibels(sg_analyser* a, double v) {
  if (!a) return fail(SG_ERR_INVALID_ARG, "analyser is null");
  if (!(v < a->max_db)) return fail(SG_ERR_INDEX_SIZE, "minDecibels must be < maxDecibels");
  a->min_db = v;
  return SG_OK;
}
int sg_analyser_set_max_decibels(sg_analyser* a, double v) {
  if (!a) return fail(SG_ERR_INVALID_ARG, "analyser is null");
  if (!(v > a->min_db)) return fail(SG_ERR_INDEX_SIZE, "maxDecibels must be > minDecibels");
  a->max_db = v;
  return SG_OK;
}
double sg_analyser_get_min_decibels(const sg_analyser* a) { return a ? a->min_db : 0.0; }
double sg_analyser_get_max_decibels(const sg_analyser* a) { return a ? a->max_db : 0.0; }
int sg_analyser_set_smoothing_time_constant(sg_analyser* a, double tau) {
  if (!a) return fail(SG_ERR_INVALID_ARG, "analyser is null");
  if (!(tau >= 0.0 && tau <= 1.0)) return fail(SG_ERR_INDEX_SIZE, "smoothingTimeConstant must be in [0, 1]");
  a->tau = tau;
  return SG_OK;
}
double sg_analyser_get_smoothing_time_constant(const sg_analyser* a) { return a ? a->tau : 0.0; }

int sg_analyser_push(sg_analyser* a, const float* samples, int64_t n) {
  if (!a) return fail(SG_ERR_INVALID_ARG, "analyser is null");
  if (n < 0 || (n > 0 && !samples)) return fail(SG_ERR_INVALID_ARG, "bad sample array");
  if (n == 0) return SG_OK;
  if (n > kRing) { samples += n - kRing; n = kRing; }
  for (int64_t i = 0; i < n; ++i) a->ring[(a->write + i) % kRing] = samples[i];
  a->write = (a->write + n) % kRing;
  a->dirty = true;
  return SG_OK;
}

int sg_analyser_get_byte_frequency_data(sg_analyser* a, uint8_t* dst, int64_t len) {
  if (!a) return fail(SG_ERR_INVALID_ARG, "analyser is null");
  return a->read_frequency<sg::kOutU8>(dst, len);
}
int sg_analyser_get_float_frequency_data(sg_analyser* a, float* dst, int64_t len) {
  if (!a) return fail(SG_ERR_INVALID_ARG, "analyser is null");
  return a->read_frequency<sg::kOutF32Db>(dst, len);
}

int sg_analyser_get_float_time_domain_data(sg_analyser* a, float* dst, int64_t len) {
  if (!a) return fail(SG_ERR_INVALID_ARG, "analyser is null");
  if (len < 0 || (len > 0 && !dst)) return fail(SG_ERR_INVALID_ARG, "bad destination array");
  const int n = (int)std::min<int64_t>(len, a->fft_size);
  std::vector<float> tmp(a->fft_size);
  a->linearise(tmp.data(), a->fft_size);  // a copy: the most recent fftSize samples, oldest first
  std::memcpy(dst, tmp.data(), sizeof(float) * n);
  return SG_OK;
}

int sg_analyser_get_byte_time_domain_data(sg_analyser* a, uint8_t* dst, int64_t len) {
  if (!a) return fail(SG_ERR_INVALID_ARG, "analyser is null");
  if (len < 0 || (len > 0 && !dst)) return fail(SG_ERR_INVALID_ARG, "bad destination array");
  const int n = a->fft_size;
  const int64_t m = std::min<int64_t>(len, n);
  if (m == 0) return SG_OK;
  sg_engine* e = a->e;
  std::lock_guard<std::mutex> lock(e->mu);
  SG_CUDA(cudaSetDevice(e->device));
  SG_TRY(a->h_block.reserve(sizeof(float) * n));
  SG_TRY(a->d_block.reserve(sizeof(float) * n));
  SG_TRY(a->d_out.reserve(4 * (size_t)n));
  SG_TRY(a->h_out.reserve(4 * (size_t)n));
  a->linearise((float*)a->h_block.p, n);
  SG_CUDA(cudaMemcpyAsync(a->d_block.p, a->h_block.p, sizeof(float) * n, cudaMemcpyHostToDevice, e->stream));
  sg::time_domain_byte_kernel<<<(unsigned)((m + 255) / 256), 256, 0, e->stream>>>((const float*)a->d_block.p,
                                                                                 (uint8_t*)a->d_out.p, m);
  e->launches++;
  SG_CUDA(cudaGetLastError());
  SG_CUDA(cudaMemcpyAsync(a->h_out.p, a->d_out.p, m, cudaMemcpyDeviceToHost, e->stream));
  SG_CUDA(cudaStreamSynchronize(e->stream));
  std::memcpy(dst, a->h_out.p, m);
  return SG_OK;
}

// ------------------------------------------------------------------------------------------
// streaming: n_channels analysers advanced in lock step, hop-spaced frames per chunk
// ------------------------------------------------------------------------------------------
int sg_stream_reset(sg_stream* s) {
  if (!s) return fail(SG_ERR_INVALID_ARG, "stream is null");
  std::lock_guard<std::mutex> lock(s->e->mu);
  SG_CUDA(cudaSetDevice(s->e->device));
  if (s->slide_pending) { SG_CUDA(cudaEventSynchronize(s->ev_slide)); s->slide_pending = false; }
  for (int i = 0; i < 2; ++i) SG_CUDA(cudaMemsetAsync(s->hist[i].p, 0, s->hist[i].cap, s->e->stream));
  SG_CUDA(cudaMemsetAsync(s->d_state.p, 0, s->d_state.cap, s->e->stream));
  s->cur = 0;
  s->frames_emitted = 0;
  return SG_OK;
}

int sg_stream_create(sg_engine* e, int n_channels, const sg_stft_config* cfg, int max_chunk, sg_stream** out) {
  if (!e || !out) return fail(SG_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  SG_TRY(validate_cfg(cfg));
  if (n_channels < 1) return fail(SG_ERR_INVALID_ARG, "n_channels must be >= 1");
  if (max_chunk < cfg->hop || max_chunk % cfg->hop) return fail(SG_ERR_INVALID_ARG, "max_chunk must be a positive multiple of hop");
  sg_stream* s = new sg_stream();
  s->e = e; s->channels = n_channels; s->max_chunk = max_chunk; s->cfg = *cfg;
  s->cfg.align = SG_ALIGN_VALID;  // geometry is set up by push(); see there
  if (cfg->window == SG_WINDOW_CUSTOM) {
    s->custom_window.assign(cfg->custom_window, cfg->custom_window + cfg->n_fft);
    s->cfg.custom_window = s->custom_window.data();
  }
  if (cfg->colormap) {
    s->colormap.assign(cfg->colormap, cfg->colormap + 256);
    s->cfg.colormap = s->colormap.data();
  }
  s->pitch = ((long long)cfg->n_fft + max_chunk + 3) & ~3LL;
  const int bins = cfg->n_fft / 2;
  const long long max_frames = max_chunk / cfg->hop;
  int rc = SG_OK;
  {
    std::lock_guard<std::mutex> lock(e->mu);
    cudaSetDevice(e->device);
    for (int i = 0; i < 2 && rc == SG_OK; ++i) rc = s->hist[i].reserve(sizeof(float) * s->pitch * n_channels);
    if (rc == SG_OK) rc = s->d_state.reserve(sizeof(float) * (size_t)bins * n_channels);
    if (rc == SG_OK) rc = s->d_out.reserve(4 * (size_t)bins * n_channels * max_frames);
    if (rc == SG_OK) rc = s->d_rgba.reserve(4 * (size_t)bins * n_channels * max_frames);
    if (rc == SG_OK) rc = s->h_in.reserve(sizeof(float) * (size_t)max_chunk * n_channels);
    if (rc == SG_OK) rc = s->h_out.reserve(4 * (size_t)bins * n_channels * max_frames);
    if (rc == SG_OK) rc = s->h_rgba.reserve(4 * (size_t)bins * n_channels * max_frames);
  }
  if (rc == SG_OK) rc = sg_stream_reset(s);
  if (rc != SG_OK) { sg_stream_destroy(s); return rc; }
  *out = s;
  return SG_OK;
}

int sg_stream_destroy(sg_stream* s) {
  if (!s) return SG_OK;
  {
    std::lock_guard<std::mutex> lock(s->e->mu);
    cudaSetDevice(s->e->device);
    cudaStreamSynchronize(s->e->stream);
    cudaStreamSynchronize(s->e->s_h2d);
    cudaStreamSynchronize(s->e->s_d2h);
    for (cudaEvent_t ev : {s->ev_in, s->ev_frames, s->ev_slide, s->ev_bytes}) if (ev) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) s->hist[i].release();
    s->d_out.release(); s->d_rgba.release(); s->d_state.release();
    s->h_in.release(); s->h_out.release(); s->h_rgba.release();
  }
  delete s;
  return SG_OK;
}

int64_t sg_stream_frames_emitted(const sg_stream* s) { return s ? s->frames_emitted : 0; }
int sg_stream_info(const sg_stream* s, int* channels, int* hop, int* bins, int* output, int* max_chunk) {
  if (!s) return fail(SG_ERR_INVALID_ARG, "stream is null");
  if (channels) *channels = s->channels;
  if (hop) *hop = s->cfg.hop;
  if (bins) *bins = s->cfg.n_fft / 2;
  if (output) *output = s->cfg.output;
  if (max_chunk) *max_chunk = s->max_chunk;
  return SG_OK;
}

int sg_stream_push(sg_stream* s, const float* chunk, int chunk_len, void* out, uint32_t* out_rgba) {
  if (!s) return fail(SG_ERR_INVALID_ARG, "stream is null");
  if (!chunk || !out) return fail(SG_ERR_INVALID_ARG, "null buffer");
  const sg_stft_config& c = s->cfg;
  if (chunk_len < c.hop || chunk_len > s->max_chunk || chunk_len % c.hop)
    return fail(SG_ERR_INVALID_ARG, "chunk_len must be a multiple of hop in [hop, max_chunk]");
  if (out_rgba && c.output != SG_OUT_U8) return fail(SG_ERR_INVALID_ARG, "out_rgba needs cfg.output == SG_OUT_U8");
  sg_engine* e = s->e;
  std::lock_guard<std::mutex> lock(e->mu);
  SG_CUDA(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  const int n = c.n_fft, bins = n / 2, ch = s->channels;
  const long long frames = chunk_len / c.hop;
  const size_t eb = elem_bytes(c.output);
  float* hist = (float*)s->hist[s->cur].p;
  float* next = (float*)s->hist[s->cur ^ 1].p;
  if (!s->ev_in) {
    for (cudaEvent_t* ev : {&s->ev_in, &s->ev_frames, &s->ev_slide, &s->ev_bytes})
      SG_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  }
  // new samples land behind the n_fft samples of history: row = [history | chunk]
  const float* src = chunk;
  if (!s->pinned(chunk, 0)) {
    std::memcpy(s->h_in.p, chunk, sizeof(float) * (size_t)chunk_len * ch);
    src = (const float*)s->h_in.p;
  }
  if (s->slide_pending) SG_CUDA(cudaStreamWaitEvent(st, s->ev_slide, 0));   // this history row set is being written
  SG_CUDA(cudaMemcpy2DAsync(hist + n, s->pitch * sizeof(float), src, (size_t)chunk_len * sizeof(float),
                            (size_t)chunk_len * sizeof(float), ch, cudaMemcpyHostToDevice, st));
  SG_CUDA(cudaEventRecord(s->ev_in, st));
  // slide the history beside the frame kernel: the last n_fft samples of [history | chunk] become the next history
  SG_CUDA(cudaStreamWaitEvent(e->s_h2d, s->ev_in, 0));
  SG_CUDA(cudaMemcpy2DAsync(next, s->pitch * sizeof(float), hist + chunk_len, s->pitch * sizeof(float),
                            (size_t)n * sizeof(float), ch, cudaMemcpyDeviceToDevice, e->s_h2d));
  SG_CUDA(cudaEventRecord(s->ev_slide, e->s_h2d));
  s->slide_pending = true;
  // frame t ends at history + (t+1)*hop: with the row shifted by hop it is the "valid" geometry
  Plan* pl;
  SG_TRY(e->get_plan(c, &pl));
  if (c.smoothing != 0.f) SG_TRY(e->scratch_acquire(st));   // an asynchronous sg_stft_batch_device call may still hold the scratch
  const uint32_t* lut;
  SG_TRY(e->lut_for(c, st, &lut));
  SG_TRY(run_range(e, *pl, c, hist + c.hop, ch, (long long)n + chunk_len - c.hop, s->pitch, 0, frames, frames,
                   s->d_out.p, (float*)s->d_state.p, lut, st));
  const size_t n_out = (size_t)ch * frames * bins;
  const bool out_pin = s->pinned(out, 1), rgba_pin = out_rgba && s->pinned(out_rgba, 2);
  if (out_rgba) {
    // the byte rows go home on the copy stream while the LUT kernel expands them
    SG_CUDA(cudaEventRecord(s->ev_frames, st));
    SG_CUDA(cudaStreamWaitEvent(e->s_d2h, s->ev_frames, 0));
    SG_CUDA(cudaMemcpyAsync(out_pin ? out : s->h_out.p, s->d_out.p, n_out * eb, cudaMemcpyDeviceToHost, e->s_d2h));
    SG_CUDA(cudaEventRecord(s->ev_bytes, e->s_d2h));
    sg::lut_kernel<<<(unsigned)std::min<size_t>((n_out + 255) / 256, 148 * 8), 256, 0, st>>>(
        (const uint8_t*)s->d_out.p, (uint32_t*)s->d_rgba.p, (long long)n_out, lut);
    e->launches++;
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaMemcpyAsync(rgba_pin ? (void*)out_rgba : s->h_rgba.p, s->d_rgba.p, n_out * 4, cudaMemcpyDeviceToHost, st));
  } else {
    SG_CUDA(cudaMemcpyAsync(out_pin ? out : s->h_out.p, s->d_out.p, n_out * eb, cudaMemcpyDeviceToHost, st));
  }
  s->cur ^= 1;
  SG_CUDA(cudaStreamSynchronize(st));
  if (out_rgba) SG_CUDA(cudaEventSynchronize(s->ev_bytes));
  if (!out_pin) std::memcpy(out, s->h_out.p, n_out * eb);
  if (out_rgba && !rgba_pin) std::memcpy(out_rgba, s->h_rgba.p, n_out * 4);
  s->frames_emitted += frames;
  return SG_OK;
}

// ------------------------------------------------------------------------------------------
// sonogram ring: the reference's bins x 256 byte texture (3D/visualizer.js:301-329, 399-416) and its view
// ------------------------------------------------------------------------------------------
int sg_ring_reset(sg_ring* r) {
  if (!r) return fail(SG_ERR_INVALID_ARG, "ring is null");
  std::lock_guard<std::mutex> lock(r->e->mu);
  SG_CUDA(cudaSetDevice(r->e->device));
  SG_CUDA(cudaMemsetAsync(r->tex.p, 0, (size_t)r->bins * r->rows, r->e->stream));
  r->yoffset = 0;
  return SG_OK;
}

int sg_ring_create(sg_engine* e, int bins, int rows, sg_ring** out) {
  if (!e || !out) return fail(SG_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if (bins < 1 || bins > 16384) return fail(SG_ERR_INVALID_ARG, "bins must be in [1, 16384]");
  if (rows < 2 || rows > 65536) return fail(SG_ERR_INVALID_ARG, "rows must be in [2, 65536]");
  sg_ring* r = new sg_ring();
  r->e = e; r->bins = bins; r->rows = rows;
  int rc;
  {
    std::lock_guard<std::mutex> lock(e->mu);
    cudaSetDevice(e->device);
    rc = r->tex.reserve((size_t)bins * rows);
  }
  if (rc == SG_OK) rc = sg_ring_reset(r);
  if (rc != SG_OK) { sg_ring_destroy(r); return rc; }
  *out = r;
  return SG_OK;
}

int sg_ring_destroy(sg_ring* r) {
  if (!r) return SG_OK;
  {
    std::lock_guard<std::mutex> lock(r->e->mu);
    cudaSetDevice(r->e->device);
    cudaStreamSynchronize(r->e->stream);
    r->tex.release(); r->img.release(); r->h_img.release(); r->h_rows.release();
  }
  delete r;
  return SG_OK;
}

int sg_ring_yoffset(const sg_ring* r) { return r ? r->yoffset : SG_ERR_INVALID_ARG; }
int sg_ring_info(const sg_ring* r, int* bins, int* rows) {
  if (!r) return fail(SG_ERR_INVALID_ARG, "ring is null");
  if (bins) *bins = r->bins;
  if (rows) *rows = r->rows;
  return SG_OK;
}

int sg_ring_append(sg_ring* r, const uint8_t* frames, int n_rows) {
  if (!r) return fail(SG_ERR_INVALID_ARG, "ring is null");
  if (n_rows < 0 || (n_rows > 0 && !frames)) return fail(SG_ERR_INVALID_ARG, "bad rows");
  if (n_rows == 0) return SG_OK;
  sg_engine* e = r->e;
  std::lock_guard<std::mutex> lock(e->mu);
  SG_CUDA(cudaSetDevice(e->device));
  // only the last `rows` frames can survive in the texture
  long long skip = n_rows > r->rows ? n_rows - r->rows : 0;
  int y = (int)((r->yoffset + skip) % r->rows);
  const uint8_t* src = frames + skip * r->bins;
  long long left = n_rows - skip;
  if (!is_pinned(frames)) {
    SG_TRY(r->h_rows.reserve((size_t)left * r->bins));
    std::memcpy(r->h_rows.p, src, (size_t)left * r->bins);
    src = (const uint8_t*)r->h_rows.p;
  }
  while (left > 0) {   // at most two spans: up to the end of the texture, then from row 0
    const long long span = std::min<long long>(left, r->rows - y);
    SG_CUDA(cudaMemcpyAsync((uint8_t*)r->tex.p + (size_t)y * r->bins, src, (size_t)span * r->bins,
                            cudaMemcpyHostToDevice, e->stream));
    src += span * r->bins; left -= span; y = (int)((y + span) % r->rows);
  }
  SG_CUDA(cudaStreamSynchronize(e->stream));   // the caller's buffer is free again (texSubImage2D semantics)
  r->yoffset = (int)((r->yoffset + (long long)n_rows) % r->rows);
  return SG_OK;
}

int sg_ring_read(sg_ring* r, uint8_t* dst) {
  if (!r || !dst) return fail(SG_ERR_INVALID_ARG, "null argument");
  std::lock_guard<std::mutex> lock(r->e->mu);
  SG_CUDA(cudaSetDevice(r->e->device));
  SG_CUDA(cudaMemcpyAsync(dst, r->tex.p, (size_t)r->bins * r->rows, cudaMemcpyDeviceToHost, r->e->stream));
  SG_CUDA(cudaStreamSynchronize(r->e->stream));
  return SG_OK;
}

int sg_ring_view(sg_ring* r, int width, int height, uint32_t* rgba_out) {
  if (!r || !rgba_out) return fail(SG_ERR_INVALID_ARG, "null argument");
  if (width < 1 || height < 1 || (long long)width * height > (1LL << 28))
    return fail(SG_ERR_INVALID_ARG, "image size out of range");
  sg_engine* e = r->e;
  std::lock_guard<std::mutex> lock(e->mu);
  SG_CUDA(cudaSetDevice(e->device));
  const size_t n = (size_t)width * height;
  SG_TRY(r->img.reserve(n * 4));
  sg::sonogram_view_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(
      (const uint8_t*)r->tex.p, r->bins, r->rows, r->yoffset, width, height, 0.08f, (uint32_t*)r->img.p);
  e->launches++;
  e->last_kernel = "sonogram_view";
  SG_CUDA(cudaGetLastError());
  const bool pin = is_pinned(rgba_out);
  if (!pin) SG_TRY(r->h_img.reserve(n * 4));
  SG_CUDA(cudaMemcpyAsync(pin ? (void*)rgba_out : r->h_img.p, r->img.p, n * 4, cudaMemcpyDeviceToHost, e->stream));
  SG_CUDA(cudaStreamSynchronize(e->stream));
  if (!pin) std::memcpy(rgba_out, r->h_img.p, n * 4);
  return SG_OK;
}

}  // extern "C"
