// Translation unit: TMA-staged frame-pair kernel (n_fft 2048, any hop).
#include "kernel_w32x2.cuh"

namespace sg {

int launch_w32x2(int out_kind, const FrameGeom& g, const W32Plan& p, const Epilogue& ep, void* out, int sm_count,
                 int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    using T = typename OutElem<OUT>::type;
    const long long pairs = (g.total_frames + 1) / 2;
    const int grid = (int)std::min<long long>((pairs + kX2Warps - 1) / kX2Warps, sm_count);
    cudaError_t rc;
    if (g.hop == 512) {
      rc = ensure_dynamic_smem<stft_w32x2_kernel<OUT, 8>>(kX2SmemBytes, device);
      if (rc != cudaSuccess) return (int)rc;
      stft_w32x2_kernel<OUT, 8><<<grid, kX2Warps * 32, kX2SmemBytes, st>>>(g, p, ep, (T*)out);
    } else if ((g.hop & 3) || (g.clip_stride & 3)) {
      rc = ensure_dynamic_smem<stft_w32x2_kernel<OUT, -1>>(kX2SmemBytes, device);
      if (rc != cudaSuccess) return (int)rc;
      stft_w32x2_kernel<OUT, -1><<<grid, kX2Warps * 32, kX2SmemBytes, st>>>(g, p, ep, (T*)out);
    } else {
      rc = ensure_dynamic_smem<stft_w32x2_kernel<OUT, 0>>(kX2SmemBytes, device);
      if (rc != cudaSuccess) return (int)rc;
      stft_w32x2_kernel<OUT, 0><<<grid, kX2Warps * 32, kX2SmemBytes, st>>>(g, p, ep, (T*)out);
    }
    return (int)cudaGetLastError();
  });
}

}  // namespace sg
