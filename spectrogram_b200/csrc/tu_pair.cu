// Translation unit: part-warp frame-pair kernels (n_fft 1024 and 512).
#include "kernel_pair.cuh"

namespace sg {

bool pair_kernel_serves(int n_fft, int hop) {
  return (n_fft == 1024 && (hop == 256 || hop == 128)) || (n_fft == 512 && (hop == 160 || hop == 128)) ||
         (n_fft == 256 && hop == 64);
}

template <int OUT, int LOG2L, int HOPJ>
static int launch_one(const FrameGeom& g, const PairPlan& p, const Epilogue& ep, void* out, int sm_count, int device,
                      cudaStream_t st) {
  using T = typename OutElem<OUT>::type;
  using S = PairShape<LOG2L>;
  const cudaError_t rc = ensure_dynamic_smem<stft_pair_kernel<OUT, LOG2L, HOPJ>>(S::kSmemBytes, device);
  if (rc != cudaSuccess) return (int)rc;
  const long long per_warp = 2 * S::PW;                             // frames one warp takes per iteration
  const long long groups = (g.total_frames + per_warp - 1) / per_warp;
  const int grid = (int)std::min<long long>((groups + kPairWarps - 1) / kPairWarps, sm_count);
  stft_pair_kernel<OUT, LOG2L, HOPJ><<<grid, kPairWarps * 32, S::kSmemBytes, st>>>(g, p, ep, (T*)out);
  return (int)cudaGetLastError();
}

int launch_pair(int out_kind, const FrameGeom& g, const PairPlan& p, const Epilogue& ep, void* out, int sm_count,
                int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    if (g.n_fft == 1024) {
      if (g.hop == 256) return launch_one<OUT, 4, 8>(g, p, ep, out, sm_count, device, st);
      return launch_one<OUT, 4, 4>(g, p, ep, out, sm_count, device, st);
    }
    if (g.n_fft == 256) return launch_one<OUT, 2, 8>(g, p, ep, out, sm_count, device, st);
    if (g.hop == 160) return launch_one<OUT, 3, 10>(g, p, ep, out, sm_count, device, st);
    return launch_one<OUT, 3, 8>(g, p, ep, out, sm_count, device, st);
  });
}

}  // namespace sg
