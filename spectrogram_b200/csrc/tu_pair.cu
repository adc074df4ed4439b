// Translation unit: part-warp frame-pair kernels, compiled once per lane-group size (-DSG_PAIR_LOG2L=4|3|2:
// n_fft 1024 | 512 | 256) so the instantiations build in parallel.
#include "kernel_pair.cuh"

#ifndef SG_PAIR_LOG2L
#error "compile with -DSG_PAIR_LOG2L=<2|3|4>"
#endif

namespace sg {

template <int OUT, int HOPJ>
static int launch_one(const FrameGeom& g, const PairPlan& p, const Epilogue& ep, void* out, int sm_count, int device,
                      cudaStream_t st) {
  constexpr int LOG2L = SG_PAIR_LOG2L;
  using T = typename OutElem<OUT>::type;
  using S = PairShape<LOG2L>;
  constexpr int NW = pair_warps<OUT>();
  constexpr int smem = S::smem_bytes(NW);
  const cudaError_t rc = ensure_dynamic_smem<stft_pair_kernel<OUT, LOG2L, HOPJ>>(smem, device);
  if (rc != cudaSuccess) return (int)rc;
  const long long per_warp = 2 * S::PW;                             // frames one warp takes per iteration
  const long long groups = (g.total_frames + per_warp - 1) / per_warp;
  const int grid = (int)std::min<long long>((groups + NW - 1) / NW, sm_count);
  stft_pair_kernel<OUT, LOG2L, HOPJ><<<grid, NW * 32, smem, st>>>(g, p, ep, (T*)out);
  return (int)cudaGetLastError();
}

#define SG_CAT2(a, b) a##b
#define SG_CAT(a, b) SG_CAT2(a, b)
// hop = 2 * L * HOPJ samples; returns -1 when this lane-group size has no instantiation for the hop
int SG_CAT(launch_pair_l, SG_PAIR_LOG2L)(int out_kind, const FrameGeom& g, const PairPlan& p, const Epilogue& ep, void* out,
                                        int sm_count, int device, cudaStream_t st) {
  constexpr int L = 1 << SG_PAIR_LOG2L;
  if (g.hop % (2 * L)) return -1;
  const int hopj = g.hop / (2 * L);
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    switch (hopj) {
      case 4: return launch_one<OUT, 4>(g, p, ep, out, sm_count, device, st);     // hop = n_fft / 8
      case 8: return launch_one<OUT, 8>(g, p, ep, out, sm_count, device, st);     // hop = n_fft / 4
      case 16: return launch_one<OUT, 16>(g, p, ep, out, sm_count, device, st);   // hop = n_fft / 2
#if SG_PAIR_LOG2L == 4
      case 5: return launch_one<OUT, 5>(g, p, ep, out, sm_count, device, st);     // n_fft 1024 at hop 160
#elif SG_PAIR_LOG2L == 3
      case 10: return launch_one<OUT, 10>(g, p, ep, out, sm_count, device, st);   // n_fft 512 at hop 160
#endif
      default: return -1;
    }
  });
}

}  // namespace sg
