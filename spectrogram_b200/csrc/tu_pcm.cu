// Translation unit: PCM ingestion kernels.
#include "kernel_pcm.cuh"
#include "plans.cuh"
#include <algorithm>

namespace sg {

int pcm_tile_frames(int bytes_per_frame) {
  const int tf = (kPcmTileBytes / bytes_per_frame) & ~(kPcmThreads - 1);
  return tf < kPcmThreads ? kPcmThreads : tf;      // 32 channels x 4 bytes x 256 frames = the 32 KB tile exactly
}

int launch_pcm_ingest(int format, const PcmGeom& g, long long n_clips, const PcmMix& m, cudaStream_t st) {
  // 16-bit mono / stereo with 16-byte friendly rows: one 16-byte word per thread, no staging
  if (format == kPcmS16 && g.channels <= 2 && n_clips > 0 && g.frames > 0) {
    const int fpw = g.channels == 1 ? 8 : 4;                       // sample frames per 16-byte word
    const bool rows_ok = (reinterpret_cast<uintptr_t>(g.src) & 15) == 0 && (g.clip_bytes & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(g.out) & 15) == 0 && (g.out_stride & 3) == 0 && g.frames % fpw == 0;
    if (rows_ok && (g.clip_bytes == g.frames * 2 * g.channels || n_clips == 1)) {
      const long long wpc = g.frames / fpw;
      const long long total = wpc * n_clips;
      const unsigned grid = (unsigned)std::min<long long>((total + 255) / 256, 148LL * 16);
      const uint4* src = reinterpret_cast<const uint4*>(g.src);
      if (g.channels == 1) pcm_s16_vec_kernel<0><<<grid, 256, 0, st>>>(src, wpc, n_clips, g.out, g.out_stride);
      else if (g.planes == 1) pcm_s16_vec_kernel<1><<<grid, 256, 0, st>>>(src, wpc, n_clips, g.out, g.out_stride);
      else pcm_s16_vec_kernel<2><<<grid, 256, 0, st>>>(src, wpc, n_clips, g.out, g.out_stride);
      return (int)cudaGetLastError();
    }
  }
  const long long blocks = n_clips * g.tiles_per_clip;
  if (blocks <= 0) return 0;
  if (blocks > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  switch (format) {
    case kPcmU8:  pcm_ingest_kernel<kPcmU8><<<(unsigned)blocks, kPcmThreads, 0, st>>>(g, m); break;
    case kPcmS16: pcm_ingest_kernel<kPcmS16><<<(unsigned)blocks, kPcmThreads, 0, st>>>(g, m); break;
    case kPcmS24: pcm_ingest_kernel<kPcmS24><<<(unsigned)blocks, kPcmThreads, 0, st>>>(g, m); break;
    case kPcmS32: pcm_ingest_kernel<kPcmS32><<<(unsigned)blocks, kPcmThreads, 0, st>>>(g, m); break;
    case kPcmF32: pcm_ingest_kernel<kPcmF32><<<(unsigned)blocks, kPcmThreads, 0, st>>>(g, m); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

}  // namespace sg
