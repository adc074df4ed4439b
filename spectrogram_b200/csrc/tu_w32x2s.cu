// Translation unit: frame-pair kernel with the smoothing recurrence on chip (n_fft 2048, hop 512 / 256, tau > 0).
#include "kernel_w32x2s.cuh"

namespace sg {

template <int OUT, int HOPJ>
static int launch_xs(const FrameGeom& g, const XsGeom& x, const W32Plan& p, const Epilogue& ep, void* out, int grid,
                     int device, cudaStream_t st) {
  using T = typename OutElem<OUT>::type;
  const cudaError_t rc = ensure_dynamic_smem<stft_w32x2s_kernel<OUT, HOPJ>>(kXsSmemBytes, device);
  if (rc != cudaSuccess) return (int)rc;
  stft_w32x2s_kernel<OUT, HOPJ><<<grid, kXsWarps * 32, kXsSmemBytes, st>>>(g, x, p, ep, (T*)out);
  return (int)cudaGetLastError();
}

int launch_w32x2s(int out_kind, const FrameGeom& g, const XsGeom& x, const W32Plan& p, const Epilogue& ep, void* out,
                  int grid, int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    if (g.hop == 256) return launch_xs<OUT, 4>(g, x, p, ep, out, grid, device, st);
    return launch_xs<OUT, 8>(g, x, p, ep, out, grid, device, st);
  });
}

}  // namespace sg
