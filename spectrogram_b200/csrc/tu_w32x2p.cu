// Translation unit: register-pipelined frame-pair kernel (n_fft 2048, hop 512).
#include "kernel_w32x2p.cuh"

namespace sg {

template <int OUT, int NW>
static int launch_xp(const FrameGeom& g, const W32Plan& p, const Epilogue& ep, void* out, int sm_count, int device,
                     cudaStream_t st) {
  using T = typename OutElem<OUT>::type;
  constexpr int smem = XpShape<NW>::kSmemBytes;
  const cudaError_t rc = ensure_dynamic_smem<stft_w32x2p_kernel<OUT, NW>>(smem, device);
  if (rc != cudaSuccess) return (int)rc;
  const long long pairs = (g.total_frames + 1) / 2;
  const int grid = (int)std::min<long long>((pairs + NW - 1) / NW, sm_count);
  stft_w32x2p_kernel<OUT, NW><<<grid, NW * 32, smem, st>>>(g, p, ep, (T*)out);
  return (int)cudaGetLastError();
}

int launch_w32x2p(int out_kind, int warps, const FrameGeom& g, const W32Plan& p, const Epilogue& ep, void* out,
                  int sm_count, int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    if (warps == 12) return launch_xp<OUT, 12>(g, p, ep, out, sm_count, device, st);
    return launch_xp<OUT, 8>(g, p, ep, out, sm_count, device, st);
  });
}

}  // namespace sg
