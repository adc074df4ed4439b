// Translation unit: frame-pair kernel with the smoothing recurrence fused in (n_fft 2048, hop 512 / 256, tau > 0).
#include <cstdlib>

#include "kernel_w32x2s.cuh"

namespace sg {

template <int OUT, int HOPJ, int NW, bool LATE, int K, bool NOSYNC = false>
static int launch_xs(const FrameGeom& g, const XsGeom& x, const W32Plan& p, const Epilogue& ep, void* out, int grid,
                     int device, cudaStream_t st) {
  using T = typename OutElem<OUT>::type;
  constexpr int smem = XsShape<NW>::kSmemBytes;
  const cudaError_t rc = ensure_dynamic_smem<stft_w32x2s_kernel<OUT, NW, HOPJ, LATE, K, NOSYNC>>(smem, device);
  if (rc != cudaSuccess) return (int)rc;
  stft_w32x2s_kernel<OUT, NW, HOPJ, LATE, K, NOSYNC><<<grid, NW * 32, smem, st>>>(g, x, p, ep, (T*)out);
  return (int)cudaGetLastError();
}

int launch_w32x2s(int out_kind, const FrameGeom& g, const XsGeom& x, const W32Plan& p, const Epilogue& ep, void* out,
                  int grid, int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    static const int variant = [] { const char* v = getenv("SG_XS_VARIANT"); return v ? atoi(v) : 0; }();   // A/B runs
    if (g.hop == 256) return launch_xs<OUT, 4, 8, true, 4>(g, x, p, ep, out, grid, device, st);
    if constexpr (OUT == kOutU8) {
      if (variant == 1) return launch_xs<OUT, 8, 8, true, 1>(g, x, p, ep, out, grid, device, st);
      if (variant == 2) return launch_xs<OUT, 8, 8, true, 2>(g, x, p, ep, out, grid, device, st);
      if (variant == 3) return launch_xs<OUT, 8, 12, true, 4>(g, x, p, ep, out, grid, device, st);
      if (variant == 4) return launch_xs<OUT, 8, 12, true, 2>(g, x, p, ep, out, grid, device, st);
      if (variant == 5) return launch_xs<OUT, 8, 8, false, 4>(g, x, p, ep, out, grid, device, st);
      if (variant == 6) return launch_xs<OUT, 8, 8, true, 1, true>(g, x, p, ep, out, grid, device, st);
      if (variant == 7) return launch_xs<OUT, 8, 12, true, 1, true>(g, x, p, ep, out, grid, device, st);
      if (variant == 8) return launch_xs<OUT, 8, 12, true, 1>(g, x, p, ep, out, grid, device, st);
    }
    return launch_xs<OUT, 8, 8, true, 4>(g, x, p, ep, out, grid, device, st);
  });
}

}  // namespace sg
