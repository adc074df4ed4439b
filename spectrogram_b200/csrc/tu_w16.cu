// Translation unit: 16-points-per-thread kernel for n_fft = 512.
#include "kernel_w16.cuh"

namespace sg {

int launch_w16(int out_kind, const FrameGeom& g, const W16Plan& p, const Epilogue& ep, void* out, int sm_count,
               int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    using T = typename OutElem<OUT>::type;
    const cudaError_t rc = ensure_dynamic_smem<stft_w16_kernel<OUT>>(kW16SmemBytes, device);
    if (rc != cudaSuccess) return (int)rc;
    const long long groups = (g.total_frames + kW16FPC - 1) / kW16FPC;
    const int grid = (int)std::min<long long>(groups, 4LL * sm_count);
    stft_w16_kernel<OUT><<<grid, kW16Threads, kW16SmemBytes, st>>>(g, p, ep, (T*)out);
    return (int)cudaGetLastError();
  });
}

int launch_w16x8(int out_kind, const FrameGeom& g, const W16Plan& p, const Epilogue& ep, void* out, int sm_count,
                 int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    using T = typename OutElem<OUT>::type;
    const cudaError_t rc = ensure_dynamic_smem<stft_w16x8_kernel<OUT>>(kW8SmemBytes, device);
    if (rc != cudaSuccess) return (int)rc;
    const long long groups = (g.total_frames + kW8FPC - 1) / kW8FPC;
    const int grid = (int)std::min<long long>(groups, 4LL * sm_count);
    stft_w16x8_kernel<OUT><<<grid, kW16Threads, kW8SmemBytes, st>>>(g, p, ep, (T*)out);
    return (int)cudaGetLastError();
  });
}

}  // namespace sg
