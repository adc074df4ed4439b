// Translation unit: register-resident n_fft = 400 kernel.
#include "kernel_r400.cuh"

namespace sg {

int launch_r400(int out_kind, const FrameGeom& g, const R400Plan& p, const Epilogue& ep, void* out, int sm_count,
                int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    using T = typename OutElem<OUT>::type;
    const cudaError_t rc = ensure_dynamic_smem<stft_r400_kernel<OUT>>(kR4SmemBytes, device);
    if (rc != cudaSuccess) return (int)rc;
    const long long groups = (g.total_frames + 5) / 6;
    const int grid = (int)std::min<long long>((groups + kR4Warps - 1) / kR4Warps, 2LL * sm_count);
    stft_r400_kernel<OUT><<<grid, kR4Warps * 32, kR4SmemBytes, st>>>(g, p, ep, (T*)out);
    return (int)cudaGetLastError();
  });
}

}  // namespace sg
