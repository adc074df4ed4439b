// PCM ingestion in front of the path (included by sgcore.cu): RIFF/WAVE header walk on the host, sample
// conversion + speakers down-mix on the GPU (kernel_pcm.cuh).  Stands in for the uncompressed-PCM part of
// context.decodeAudioData (src/javascripts/util/util.js:9-17) and for the mono down-mix the AnalyserNode
// applies to its input ([SPEC] AnalyserNode / "Up-mixing and down-mixing", speakers interpretation).

namespace {

// [SPEC] speaker down-mix to mono.  Channel order as WAVE/Web Audio: L R | L R SL SR | L R C LFE SL SR.
int pcm_mix_weights(int channels, PcmMix* m) {
  for (int c = 0; c < kPcmMaxChannels; ++c) m->w[c] = 0.f;
  switch (channels) {
    case 1: m->w[0] = 1.f; break;
    case 2: m->w[0] = m->w[1] = 0.5f; break;
    case 4: m->w[0] = m->w[1] = m->w[2] = m->w[3] = 0.25f; break;
    case 6: m->w[0] = m->w[1] = 0.70710678118654752440f; m->w[2] = 1.f; m->w[3] = 0.f; m->w[4] = m->w[5] = 0.5f; break;
    default: m->w[0] = 1.f; break;  // no speaker layout: discrete, the first channel survives
  }
  return SG_OK;
}

int validate_pcm(const sg_pcm_info* info, int layout) {
  if (!info) return fail(SG_ERR_INVALID_ARG, "pcm info is null");
  if (sg_pcm_sample_bytes(info->format) == 0) return fail(SG_ERR_INVALID_ARG, "pcm format must be an sg_pcm_format");
  if (info->channels < 1 || info->channels > kPcmMaxChannels) return fail(SG_ERR_INVALID_ARG, "pcm channels must be 1..32");
  if (info->frames < 0) return fail(SG_ERR_INVALID_ARG, "pcm frames must be >= 0");
  if (layout != SG_PCM_MONO_MIX && layout != SG_PCM_PLANAR) return fail(SG_ERR_INVALID_ARG, "layout must be an sg_pcm_layout");
  return SG_OK;
}

RawPcm raw_of(const sg_pcm_info& info, int layout) {
  RawPcm r;
  r.format = info.format;
  r.channels = info.channels;
  r.planes = layout == SG_PCM_PLANAR ? info.channels : 1;
  r.bytes_per_frame = info.channels * sg_pcm_sample_bytes(info.format);
  pcm_mix_weights(info.channels, &r.mix);
  return r;
}

int ingest_on_device(sg_engine* e, const void* src_dev, long long src_bytes, long long clip_bytes, int64_t n_clips,
                     long long frames, const RawPcm& raw, float* out_dev, long long out_stride, cudaStream_t st) {
  if (n_clips == 0 || frames == 0) return SG_OK;
  PcmGeom pg;
  pg.src = (const unsigned char*)src_dev;
  pg.src_bytes = src_bytes;
  pg.clip_bytes = clip_bytes;
  pg.frames = frames;
  pg.out = out_dev;
  pg.out_stride = out_stride;
  pg.tile_frames = pcm_tile_frames(raw.bytes_per_frame);
  pg.tiles_per_clip = (frames + pg.tile_frames - 1) / pg.tile_frames;
  pg.channels = raw.channels;
  pg.planes = raw.planes;
  SG_CUDA((cudaError_t)launch_pcm_ingest(raw.format, pg, n_clips, raw.mix, st));
  e->launches++;
  return SG_OK;
}

uint32_t rd_u32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
uint32_t rd_u16(const unsigned char* p) { return p[0] | (p[1] << 8); }

}  // namespace

extern "C" {

int sg_pcm_sample_bytes(int format) {
  switch (format) {
    case SG_PCM_U8: return 1;
    case SG_PCM_S16: return 2;
    case SG_PCM_S24: return 3;
    case SG_PCM_S32: return 4;
    case SG_PCM_F32: return 4;
    default: return 0;
  }
}

int sg_pcm_num_planes(const sg_pcm_info* info, int layout) {
  if (validate_pcm(info, layout) != SG_OK) return SG_ERR_INVALID_ARG;
  return layout == SG_PCM_PLANAR ? info->channels : 1;
}

int sg_wav_parse(const void* file_bytes, size_t len, sg_pcm_info* out) {
  if (!file_bytes || !out) return fail(SG_ERR_INVALID_ARG, "null argument");
  const unsigned char* p = (const unsigned char*)file_bytes;
  if (len < 12 || std::memcmp(p, "RIFF", 4) != 0 || std::memcmp(p + 8, "WAVE", 4) != 0)
    return fail(SG_ERR_INVALID_ARG, "not a RIFF/WAVE file");
  bool have_fmt = false;
  int tag = 0, bits = 0, channels = 0, rate = 0, block_align = 0;
  size_t pos = 12;
  while (pos + 8 <= len) {
    const unsigned char* ck = p + pos;
    const size_t size = rd_u32(ck + 4);
    const size_t body = pos + 8;
    if (std::memcmp(ck, "fmt ", 4) == 0) {
      if (size < 16 || body + 16 > len) return fail(SG_ERR_INVALID_ARG, "truncated fmt chunk");
      tag = (int)rd_u16(ck + 8);
      channels = (int)rd_u16(ck + 10);
      rate = (int)rd_u32(ck + 12);
      block_align = (int)rd_u16(ck + 20);
      bits = (int)rd_u16(ck + 22);
      if (tag == 0xFFFE) {  // WAVE_FORMAT_EXTENSIBLE: the sub-format GUID starts with the real tag
        if (size < 40 || body + 40 > len) return fail(SG_ERR_INVALID_ARG, "truncated extensible fmt chunk");
        tag = (int)rd_u16(ck + 8 + 24);
      }
      have_fmt = true;
    } else if (std::memcmp(ck, "data", 4) == 0) {
      if (!have_fmt) return fail(SG_ERR_INVALID_ARG, "data chunk before fmt chunk");
      int format;
      if (tag == 1 && bits == 8) format = SG_PCM_U8;
      else if (tag == 1 && bits == 16) format = SG_PCM_S16;
      else if (tag == 1 && bits == 24) format = SG_PCM_S24;
      else if (tag == 1 && bits == 32) format = SG_PCM_S32;
      else if (tag == 3 && bits == 32) format = SG_PCM_F32;
      else return fail(SG_ERR_INVALID_ARG, "unsupported WAVE encoding (format tag " + std::to_string(tag) + ", " +
                                               std::to_string(bits) + " bits): only uncompressed PCM is decoded");
      if (channels < 1 || channels > kPcmMaxChannels) return fail(SG_ERR_INVALID_ARG, "WAVE channel count must be 1..32");
      const int bpf = channels * sg_pcm_sample_bytes(format);
      if (block_align != bpf) return fail(SG_ERR_INVALID_ARG, "WAVE block align does not match channels x sample size");
      const size_t avail = len - body;            // a streamed file may carry 0 or 0xFFFFFFFF here
      const size_t bytes = (size == 0 || size == 0xFFFFFFFFu || size > avail) ? avail : size;
      out->format = format;
      out->channels = channels;
      out->sample_rate = rate;
      out->frames = (int64_t)(bytes / bpf);
      out->data_offset = (int64_t)body;
      return SG_OK;
    }
    pos = body + size + (size & 1);               // chunks are word aligned
  }
  return fail(SG_ERR_INVALID_ARG, have_fmt ? "no data chunk" : "no fmt chunk");
}

int sg_pcm_ingest_device(sg_engine* e, const void* pcm_dev, int64_t n_clips, const sg_pcm_info* info, int layout,
                         float* out_dev, int64_t out_stride, void* cuda_stream) {
  if (!e) return fail(SG_ERR_INVALID_ARG, "engine is null");
  SG_TRY(validate_pcm(info, layout));
  if (n_clips < 0 || out_stride < info->frames) return fail(SG_ERR_INVALID_ARG, "bad clip geometry");
  if (n_clips == 0 || info->frames == 0) return SG_OK;
  if (!pcm_dev || !out_dev) return fail(SG_ERR_INVALID_ARG, "null device buffer");
  std::lock_guard<std::mutex> lock(e->mu);
  SG_CUDA(cudaSetDevice(e->device));
  const RawPcm raw = raw_of(*info, layout);
  const long long clip_bytes = (long long)info->frames * raw.bytes_per_frame;
  return ingest_on_device(e, pcm_dev, clip_bytes * n_clips, clip_bytes, n_clips, info->frames, raw, out_dev, out_stride,
                          cuda_stream ? (cudaStream_t)cuda_stream : e->stream);
}

int sg_pcm_ingest(sg_engine* e, const void* pcm, int64_t n_clips, const sg_pcm_info* info, int layout, float* out) {
  if (!e) return fail(SG_ERR_INVALID_ARG, "engine is null");
  SG_TRY(validate_pcm(info, layout));
  if (n_clips < 0) return fail(SG_ERR_INVALID_ARG, "bad clip geometry");
  if (n_clips == 0 || info->frames == 0) return SG_OK;
  if (!pcm || !out) return fail(SG_ERR_INVALID_ARG, "null host buffer");
  std::lock_guard<std::mutex> lock(e->mu);
  SG_CUDA(cudaSetDevice(e->device));
  const RawPcm raw = raw_of(*info, layout);
  const long long clip_bytes = (long long)info->frames * raw.bytes_per_frame;
  // whole clips per chunk where they fit, else sample-frame ranges of one clip; copies ride the engine's stream
  const long long kChunkBytes = 32LL << 20;
  const long long frames_per_chunk = std::max<long long>(1, kChunkBytes / raw.bytes_per_frame);
  for (int64_t c = 0; c < n_clips;) {
    long long nc = 1, f0 = 0, nf = info->frames;
    if (clip_bytes <= kChunkBytes) nc = std::min<long long>(n_clips - c, std::max<long long>(1, kChunkBytes / std::max<long long>(clip_bytes, 1)));
    for (f0 = 0; f0 < info->frames; f0 += nf) {
      nf = nc > 1 ? info->frames : std::min<long long>(frames_per_chunk, info->frames - f0);
      const size_t in_b = (size_t)nc * nf * raw.bytes_per_frame;       // contiguous: nc whole clips or one range
      const size_t out_f = (size_t)nc * raw.planes * nf;
      SG_TRY(e->d_raw[0].reserve(in_b + 32));
      SG_TRY(e->d_in.reserve(out_f * sizeof(float)));
      const char* src = (const char*)pcm + (size_t)c * clip_bytes + (size_t)f0 * raw.bytes_per_frame;
      SG_CUDA(cudaMemcpyAsync(e->d_raw[0].p, src, in_b, cudaMemcpyHostToDevice, e->stream));
      SG_TRY(ingest_on_device(e, e->d_raw[0].p, (long long)in_b, nf * raw.bytes_per_frame, nc, nf, raw, (float*)e->d_in.p, nf,
                              e->stream));
      // device planes are [clip][plane][nf]; the caller's are [clip][plane][frames]
      float* dst = out + (size_t)c * raw.planes * info->frames + f0;
      SG_CUDA(cudaMemcpy2DAsync(dst, (size_t)info->frames * sizeof(float), e->d_in.p, (size_t)nf * sizeof(float),
                                (size_t)nf * sizeof(float), (size_t)nc * raw.planes, cudaMemcpyDeviceToHost, e->stream));
      SG_CUDA(cudaStreamSynchronize(e->stream));
    }
    c += nc;
  }
  return SG_OK;
}

int sg_stft_pcm(sg_engine* e, const void* pcm, int64_t n_clips, const sg_pcm_info* info, int layout,
                const sg_stft_config* cfg, void* out) {
  SG_TRY(validate_pcm(info, layout));
  const RawPcm raw = raw_of(*info, layout);
  return stft_batch_impl(e, pcm, n_clips, info->frames, cfg, out, &raw);
}

}  // extern "C"
