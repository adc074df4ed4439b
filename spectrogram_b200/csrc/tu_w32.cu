// Translation unit: one-frame-per-warp kernel (n_fft 2048) and the generic shared-memory kernel.
#include "kernel_smem.cuh"
#include "kernel_w32.cuh"

namespace sg {

int launch_w32(int out_kind, const FrameGeom& g, const W32Plan& p, const Epilogue& ep, void* out, int sm_count,
               int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    using T = typename OutElem<OUT>::type;
    const cudaError_t rc = ensure_dynamic_smem<stft_w32_kernel<OUT>>(kW32SmemBytes, device);
    if (rc != cudaSuccess) return (int)rc;
    const long long ctas_needed = (g.total_frames + kW32Warps - 1) / kW32Warps;
    const int grid = (int)std::min<long long>(ctas_needed, 2LL * sm_count);
    stft_w32_kernel<OUT><<<grid, kW32Warps * 32, kW32SmemBytes, st>>>(g, p, ep, (T*)out);
    return (int)cudaGetLastError();
  });
}

int launch_smem(int out_kind, const FrameGeom& g, const SmemPlan& p, const Epilogue& ep, void* out, int sm_count,
                int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    using T = typename OutElem<OUT>::type;
    const size_t smem = sizeof(float2) * p.m;
    if (smem > 48 * 1024) {
      const cudaError_t rc = ensure_dynamic_smem<stft_smem_kernel<OUT>>((int)smem, device);
      if (rc != cudaSuccess) return (int)rc;
    }
    const int threads = std::min(1024, std::max(32, ((p.m / 4 + 31) / 32) * 32));
    const int per_sm = std::max(1, std::min<int>(2048 / threads, (int)((200 * 1024) / std::max<size_t>(smem, 1024))));
    const int grid = (int)std::min<long long>(g.total_frames, (long long)sm_count * per_sm);
    stft_smem_kernel<OUT><<<grid, threads, smem, st>>>(g, p, ep, (T*)out);
    return (int)cudaGetLastError();
  });
}

}  // namespace sg
