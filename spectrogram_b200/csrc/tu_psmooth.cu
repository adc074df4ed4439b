// Translation unit: part-warp frame-pair kernels with the smoothing recurrence fused in (n_fft 1024 | 512 | 256, hop
// n_fft/2, n_fft/4 or n_fft/8, tau > 0), compiled once per lane-group size (-DSG_PAIR_LOG2L=4|3|2).
#include "kernel_pair_s.cuh"

#ifndef SG_PAIR_LOG2L
#error "compile with -DSG_PAIR_LOG2L=<2|3|4>"
#endif

namespace sg {

template <int OUT, int HOPJ, bool WARM = false>
static int launch_ps(const FrameGeom& g, const XsGeom& x, const PairPlan& p, const Epilogue& ep, void* out, int grid,
                     int device, cudaStream_t st) {
  constexpr int LOG2L = SG_PAIR_LOG2L;
  using T = typename OutElem<OUT>::type;
  constexpr int smem = PsShape<LOG2L>::kSmemBytes;
  const cudaError_t rc = ensure_dynamic_smem<stft_pair_s_kernel<OUT, LOG2L, HOPJ, WARM>>(smem, device);
  if (rc != cudaSuccess) return (int)rc;
  // CTAs wait for one another (a segment's first step for its predecessor's carry): a cooperative launch guarantees
  // that the whole grid (<= one CTA per SM) is resident at the same time, or fails instead of hanging
  T* out_t = (T*)out;
  void* args[] = {(void*)&g, (void*)&x, (void*)&p, (void*)&ep, (void*)&out_t};
  return (int)cudaLaunchCooperativeKernel((const void*)stft_pair_s_kernel<OUT, LOG2L, HOPJ, WARM>, dim3(grid), dim3(kPsWarps * 32),
                                          args, smem, st);
}

#define SG_CAT2(a, b) a##b
#define SG_CAT(a, b) SG_CAT2(a, b)
// hop = 2 * L * HOPJ samples share the frames' loads; every other hop runs the direct-load instantiation (HOPJ = 0)
int SG_CAT(launch_pair_s_l, SG_PAIR_LOG2L)(int out_kind, const FrameGeom& g, const XsGeom& x, const PairPlan& p,
                                          const Epilogue& ep, void* out, int grid, int device, cudaStream_t st) {
  constexpr int L = 1 << SG_PAIR_LOG2L;
  const int hopj = (g.hop % (2 * L)) ? 0 : g.hop / (2 * L);
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    if (x.mode == 2) {      // independent segments with a warm-up: instantiated for hop = n_fft / 4 only
      if (hopj != 8) return -1;
      return launch_ps<OUT, 8, true>(g, x, p, ep, out, grid, device, st);
    }
    switch (hopj) {
      case 4: return launch_ps<OUT, 4>(g, x, p, ep, out, grid, device, st);     // hop = n_fft / 8
      case 8: return launch_ps<OUT, 8>(g, x, p, ep, out, grid, device, st);     // hop = n_fft / 4
      case 16: return launch_ps<OUT, 16>(g, x, p, ep, out, grid, device, st);   // hop = n_fft / 2
#if SG_PAIR_LOG2L == 4
      case 5: return launch_ps<OUT, 5>(g, x, p, ep, out, grid, device, st);     // n_fft 1024 at hop 160
#elif SG_PAIR_LOG2L == 3
      case 10: return launch_ps<OUT, 10>(g, x, p, ep, out, grid, device, st);   // n_fft 512 at hop 160
#endif
      default: return launch_ps<OUT, 0>(g, x, p, ep, out, grid, device, st);    // any other hop: direct loads
    }
  });
}

#ifdef SG_DEBUG
int SG_CAT(dbg_attach_psmooth_l, SG_PAIR_LOG2L)(const DbgState& st) { return (int)dbg_attach(st); }
#endif

}  // namespace sg
