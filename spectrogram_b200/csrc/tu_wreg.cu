// Translation unit: register-resident FFT family, compiled once per output kind (-DSG_TU_OUT=k) so the
// 24 instantiations build in parallel.
#include "kernel_wreg.cuh"

#ifndef SG_TU_OUT
#error "compile with -DSG_TU_OUT=<kOut*>"
#endif

namespace sg {

template <int LM>
static int launch_one(const FrameGeom& g, const WregPlan& p, const Epilogue& ep, void* out, int sm_count, int device,
                      cudaStream_t st) {
  constexpr int OUT = SG_TU_OUT;
  using T = typename OutElem<OUT>::type;
  using S = WregShape<LM>;
  const cudaError_t rc = ensure_dynamic_smem<stft_wreg_kernel<LM, OUT>>(S::kSmemBytes, device);
  if (rc != cudaSuccess) return (int)rc;
  const long long groups = (g.total_frames + S::FPC - 1) / S::FPC;
  const int grid = (int)std::min<long long>(groups, 2LL * sm_count);
  stft_wreg_kernel<LM, OUT><<<grid, kWregThreads, S::kSmemBytes, st>>>(g, p, ep, (T*)out);
  return (int)cudaGetLastError();
}

#define SG_CAT2(a, b) a##b
#define SG_CAT(a, b) SG_CAT2(a, b)
int SG_CAT(launch_wreg_out, SG_TU_OUT)(int log2m, const FrameGeom& g, const WregPlan& p, const Epilogue& ep, void* out,
                                       int sm_count, int device, cudaStream_t st) {
  switch (log2m) {
    case 7: return launch_one<7>(g, p, ep, out, sm_count, device, st);
    case 8: return launch_one<8>(g, p, ep, out, sm_count, device, st);
    case 9: return launch_one<9>(g, p, ep, out, sm_count, device, st);
    case 10: return launch_one<10>(g, p, ep, out, sm_count, device, st);
    case 11: return launch_one<11>(g, p, ep, out, sm_count, device, st);
    default: return launch_one<12>(g, p, ep, out, sm_count, device, st);
  }
}

}  // namespace sg
