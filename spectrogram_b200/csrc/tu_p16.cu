// Translation unit: half-warp frame-pair kernel for n_fft = 1024 (hop 256 or 128).
#include "kernel_p16.cuh"

namespace sg {

template <int OUT, int HOPJ>
static int launch_one(const FrameGeom& g, const P16Plan& p, const Epilogue& ep, void* out, int sm_count, int device,
                      cudaStream_t st) {
  using T = typename OutElem<OUT>::type;
  const cudaError_t rc = ensure_dynamic_smem<stft_p16_kernel<OUT, HOPJ>>(kP16SmemBytes, device);
  if (rc != cudaSuccess) return (int)rc;
  const long long quads = (g.total_frames + 3) / 4;                 // one warp takes two pairs = four frames
  const int grid = (int)std::min<long long>((quads + kP16Warps - 1) / kP16Warps, sm_count);
  stft_p16_kernel<OUT, HOPJ><<<grid, kP16Warps * 32, kP16SmemBytes, st>>>(g, p, ep, (T*)out);
  return (int)cudaGetLastError();
}

int launch_p16(int out_kind, const FrameGeom& g, const P16Plan& p, const Epilogue& ep, void* out, int sm_count,
               int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    if (g.hop == 256) return launch_one<OUT, 8>(g, p, ep, out, sm_count, device, st);
    return launch_one<OUT, 4>(g, p, ep, out, sm_count, device, st);
  });
}

}  // namespace sg
