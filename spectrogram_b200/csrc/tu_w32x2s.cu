// Translation unit: frame-pair kernel with the smoothing recurrence fused in (n_fft 2048, hop 1024 / 512 / 256, tau > 0).
#include <cstdlib>

#include "kernel_w32x2s.cuh"

namespace sg {

template <int OUT, int HOPJ, int NW, bool LATE, int K>
static int launch_xs(const FrameGeom& g, const XsGeom& x, const W32Plan& p, const Epilogue& ep, void* out, int grid,
                     int device, cudaStream_t st) {
  using T = typename OutElem<OUT>::type;
  constexpr int smem = XsShape<NW>::kSmemBytes;
  const cudaError_t rc = ensure_dynamic_smem<stft_w32x2s_kernel<OUT, NW, HOPJ, LATE, K>>(smem, device);
  if (rc != cudaSuccess) return (int)rc;
  // CTAs wait for one another (a segment's first pair for its predecessor's carry): a cooperative launch guarantees
  // that the whole grid (<= one CTA per SM) is resident at the same time, or fails instead of hanging
  T* out_t = (T*)out;
  void* args[] = {(void*)&g, (void*)&x, (void*)&p, (void*)&ep, (void*)&out_t};
  return (int)cudaLaunchCooperativeKernel((const void*)stft_w32x2s_kernel<OUT, NW, HOPJ, LATE, K>, dim3(grid), dim3(NW * 32),
                                          args, smem, st);
}

int launch_w32x2s(int out_kind, const FrameGeom& g, const XsGeom& x, const W32Plan& p, const Epilogue& ep, void* out,
                  int grid, int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    // 8 warps x 255 registers, the next pair's loads issued after the turn has been passed on, one turn chain
    // (measured, 512 clips x 858 frames, u8: 470 M frames/s; 12 warps with the same chain 251 M -- 26 spilled words
    //  and three warps per scheduler waiting on one ring; four chains over slot groups 433 M; loads inside the
    //  untangle 409 M; without waiting at all, wrong results, 518 M: the ordering itself costs 9 %)
    if (g.hop == 256) return launch_xs<OUT, 4, 8, true, 1>(g, x, p, ep, out, grid, device, st);
    if (g.hop == 1024) return launch_xs<OUT, 16, 8, true, 1>(g, x, p, ep, out, grid, device, st);
    if (g.hop == 512) return launch_xs<OUT, 8, 8, true, 1>(g, x, p, ep, out, grid, device, st);
    return launch_xs<OUT, 0, 8, true, 1>(g, x, p, ep, out, grid, device, st);       // any other hop: direct loads
  });
}

#ifdef SG_DEBUG
int dbg_attach_w32x2s(const DbgState& st) { return (int)dbg_attach(st); }
#endif

}  // namespace sg
