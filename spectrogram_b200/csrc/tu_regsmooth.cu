// Translation unit: register family with the smoothing recurrence fused in (n_fft 8192; n_fft 4096 at hops the even/odd
// kernel cannot load), tau > 0.
#include "kernel_wreg_s.cuh"

namespace sg {

template <int LM, int OUT>
static int launch_ws(const FrameGeom& g, const XsGeom& x, const WregPlan& p, const Epilogue& ep, void* out, int grid, int device,
                     cudaStream_t st) {
  using T = typename OutElem<OUT>::type;
  constexpr int smem = WsShape<LM>::kSmemBytes;
  const cudaError_t rc = ensure_dynamic_smem<stft_wreg_s_kernel<LM, OUT>>(smem, device);
  if (rc != cudaSuccess) return (int)rc;
  // CTAs wait for one another (a segment's first frame for its predecessor's carry): cooperative launch, two CTAs per SM
  T* out_t = (T*)out;
  void* args[] = {(void*)&g, (void*)&x, (void*)&p, (void*)&ep, (void*)&out_t};
  return (int)cudaLaunchCooperativeKernel((const void*)stft_wreg_s_kernel<LM, OUT>, dim3(grid), dim3(kWregThreads), args, smem, st);
}

int launch_wreg_s(int out_kind, int log2m, const FrameGeom& g, const XsGeom& x, const WregPlan& p, const Epilogue& ep, void* out,
                  int grid, int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    if (log2m == 11) return launch_ws<11, OUT>(g, x, p, ep, out, grid, device, st);
    return launch_ws<12, OUT>(g, x, p, ep, out, grid, device, st);
  });
}

}  // namespace sg
