// Translation unit: n_fft 4096, one frame per warp as a packed even/odd pair of 1024-point transforms.
#include "kernel_w32eo.cuh"

namespace sg {

template <int OUT, int NW>
static int launch_eo(const FrameGeom& g, const EoPlan& p, const Epilogue& ep, void* out, int sm_count, int device,
                     cudaStream_t st) {
  using T = typename OutElem<OUT>::type;
  constexpr int smem = kEoTableBytes + NW * kXpPlaneBytes;
  const cudaError_t rc = ensure_dynamic_smem<stft_w32eo_kernel<OUT, NW>>(smem, device);
  if (rc != cudaSuccess) return (int)rc;
  const int grid = (int)std::min<long long>((g.total_frames + NW - 1) / NW, sm_count);
  stft_w32eo_kernel<OUT, NW><<<grid, NW * 32, smem, st>>>(g, p, ep, (T*)out);
  return (int)cudaGetLastError();
}

int launch_w32eo(int out_kind, int warps, const FrameGeom& g, const EoPlan& p, const Epilogue& ep, void* out, int sm_count,
                 int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    if (warps == 12) return launch_eo<OUT, 12>(g, p, ep, out, sm_count, device, st);
    return launch_eo<OUT, kEoWarps>(g, p, ep, out, sm_count, device, st);
  });
}

}  // namespace sg
