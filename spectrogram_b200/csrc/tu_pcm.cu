// Translation unit: PCM ingestion kernels.
#include "kernel_pcm.cuh"
#include "plans.cuh"

namespace sg {

int pcm_tile_frames(int bytes_per_frame) {
  const int tf = (kPcmTileBytes / bytes_per_frame) & ~(kPcmThreads - 1);
  return tf < kPcmThreads ? kPcmThreads : tf;      // 32 channels x 4 bytes x 256 frames = the 32 KB tile exactly
}

int launch_pcm_ingest(int format, const PcmGeom& g, long long n_clips, const PcmMix& m, cudaStream_t st) {
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
