// Translation unit: n_fft 4096 with the smoothing recurrence fused in (tau > 0, hops that keep frames 16-byte aligned).
#include "kernel_w32eo_s.cuh"

namespace sg {

int launch_w32eo_s(int out_kind, const FrameGeom& g, const XsGeom& x, const EoPlan& p, const Epilogue& ep, void* out, int grid,
                   int device, cudaStream_t st) {
  return dispatch_out(out_kind, [&](auto tag) {
    constexpr int OUT = decltype(tag)::value;
    using T = typename OutElem<OUT>::type;
    const cudaError_t rc = ensure_dynamic_smem<stft_w32eo_s_kernel<OUT>>(kEsSmemBytes, device);
    if (rc != cudaSuccess) return (int)rc;
    // CTAs wait for one another (a segment's first frame for its predecessor's carry): cooperative launch
    T* out_t = (T*)out;
    void* args[] = {(void*)&g, (void*)&x, (void*)&p, (void*)&ep, (void*)&out_t};
    return (int)cudaLaunchCooperativeKernel((const void*)stft_w32eo_s_kernel<OUT>, dim3(grid), dim3(kEsWarps * 32), args,
                                            kEsSmemBytes, st);
  });
}

#ifdef SG_DEBUG
int dbg_attach_w32eo_s(const DbgState& st) { return (int)dbg_attach(st); }
#endif

}  // namespace sg
