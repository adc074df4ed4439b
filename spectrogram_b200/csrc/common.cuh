// Shared device-side definitions for the spectrogram kernels (sm_100a).
//
// Path restated: Web Audio AnalyserNode "FFT windowing and smoothing over time" as the
// reference drives it (src/javascripts/UI/player.js:7-11, src/javascripts/3D/visualizer.js:346-368):
// time block -> window -> DFT/N -> |X| -> smoothing -> 20 log10 -> byte / colour.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace sg {

enum : int { kOutU8 = 0, kOutF32Db = 1, kOutRgba8 = 2, kOutF32Mag = 3 };

// Where frames come from.  Frame f (global index) belongs to clip f / frames_per_clip and is
// frame t = f % frames_per_clip of that clip; it covers samples [start0 + t*hop, +n_fft) of the
// clip, zero filled outside [0, clip_len)  (AnalyserNode's ring starts zero filled).
struct FrameGeom {
  const float* pcm;          // [n_clips][clip_stride]
  long long clip_len;
  long long clip_stride;
  long long frames_per_clip;
  long long total_frames;
  long long start0;          // 0 (valid alignment), hop - n_fft (analyser alignment), ...
  int n_fft;
  int hop;
};

// Epilogue constants, all derived on the host in double.
//   p = |X'|^2 where X' = norm * X/N is what the kernel holds (norm folds 1/N and any constant
//   factor the butterflies dropped).
//   dB   = db_scale * log2(p) + db_off              (= 20 log10(|X|/N))
//   byte = clamp(byte_a * log2(p) + byte_b, 0, 255) (= 255/(max-min) * (dB - min))
//   mag  = sqrt(p) * mag_scale
// and, for values that are already linear magnitudes m (after smoothing):
//   dB = 2*db_scale*log2(m);  byte = clamp(2*byte_a*log2(m) + byte_b0, 0, 255)
struct Epilogue {
  float db_scale;   // 10*log10(2)
  float db_off;     // -20*log10(norm)
  float byte_a;     // 255/(max-min) * 10*log10(2)
  float byte_b;     // 255/(max-min) * (db_off - min_db)
  float byte_b0;    // 255/(max-min) * (-min_db)
  float mag_scale;  // 1/norm
  const uint32_t* lut;  // device, 256 entries (RGBA8 output only)
};

template <int OUT> struct OutElem { using type = float; };
template <> struct OutElem<kOutU8> { using type = uint8_t; };
template <> struct OutElem<kOutRgba8> { using type = uint32_t; };

// [SPEC] "if X^[k] is NaN or infinite, set it to 0"
__device__ __forceinline__ float finite_or_zero(float v) { return (fabsf(v) <= 3.4028235e38f) ? v : 0.f; }

__device__ __forceinline__ unsigned byte_from_scaled(float v) {
  v = fminf(fmaxf(v, 0.f), 255.f);  // fmaxf(NaN, 0) = 0
  return (unsigned)v;               // truncation, as static_cast<unsigned char> in Chromium
}

__device__ __forceinline__ float sqrt_ftz(float x) {  // MUFU.SQRT alone (2^-23 relative), subnormal inputs read as 0
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_ftz(float x) {  // MUFU.LG2 alone: subnormal inputs read as 0 (-inf)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// cvt.rzi.u8.f32 saturates to [0, 255] and maps NaN to 0: the clamp of step 6 in one instruction
__device__ __forceinline__ unsigned byte_of_scaled(float v) {
  unsigned b;
  asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(b) : "f"(v));
  return b;
}

// from unnormalised power p = re^2 + im^2  (tau == 0 path: no sqrt needed); the caller has already applied the
// non-finite rule (kernels that decide it once per frame)
template <int OUT>
__device__ __forceinline__ typename OutElem<OUT>::type emit_power_finite(float p, const Epilogue& e) {
  if constexpr (OUT == kOutF32Mag) {
    return sqrt_ftz(p) * e.mag_scale;
  } else {
    if constexpr (OUT == kOutF32Db) {
      return fmaf(e.db_scale, lg2_ftz(p), e.db_off);      // powers below 2^-126 (|X|/N < 1e-19) read as 0: -inf dB
    } else {
      const unsigned b = byte_of_scaled(fmaf(e.byte_a, __log2f(p), e.byte_b));   // exact for any minDecibels
      if constexpr (OUT == kOutU8) return (uint8_t)b;
      else return __ldg(e.lut + b);
    }
  }
}

// from unnormalised power p = re^2 + im^2  (tau == 0 path: no sqrt needed)
template <int OUT>
__device__ __forceinline__ typename OutElem<OUT>::type emit_power(float p, const Epilogue& e) {
  p = finite_or_zero(p);
  if constexpr (OUT == kOutF32Mag) {
    return sqrtf(p) * e.mag_scale;
  } else {
    const float l = __log2f(p);
    if constexpr (OUT == kOutF32Db) {
      return fmaf(e.db_scale, l, e.db_off);
    } else {
      const unsigned b = byte_from_scaled(fmaf(e.byte_a, l, e.byte_b));
      if constexpr (OUT == kOutU8) return (uint8_t)b;
      else return __ldg(e.lut + b);
    }
  }
}

// from a linear magnitude m (tau > 0 path, after the recurrence)
template <int OUT>
__device__ __forceinline__ typename OutElem<OUT>::type emit_mag(float m, const Epilogue& e) {
  if constexpr (OUT == kOutF32Mag) {
    return m;
  } else {
    const float l = __log2f(m);
    if constexpr (OUT == kOutF32Db) {
      return 2.f * e.db_scale * l;
    } else {
      const unsigned b = byte_from_scaled(fmaf(2.f * e.byte_a, l, e.byte_b0));
      if constexpr (OUT == kOutU8) return (uint8_t)b;
      else return __ldg(e.lut + b);
    }
  }
}

// ---- SG_DEBUG build (make debug -> libsgcore_debug.so): epoch tags beside the shared-memory hand-offs.
// compute-sanitizer is closed on this pool, so the races it would look for are checked in-tree: every value a kernel
// passes through shared memory to another lane or warp is accompanied by a tag (in global memory, so no kernel's
// shared-memory budget moves) holding the iteration that wrote it, and the reader asserts that it sees the tag of
// the iteration it is in.  A missing barrier, a consumer running a pair ahead of its producer, or a stage reused
// before it was read all show up as a tag from the wrong iteration.  tests/test_debug_build.py runs the shape list
// of tools/sanitize_run.py through the debug library and expects zero mismatches -- and, with the turn chain of the
// fused smoothing kernel deliberately disabled, a non-zero count (the checker does detect a real ordering bug).
#ifdef SG_DEBUG
constexpr int kDbgWarpSlots = 2048, kDbgMaxWarps = 16, kDbgCtaSlots = 512, kDbgSites = 16;
struct DbgState {
  unsigned* tags;                 // [CTA][kDbgMaxWarps][kDbgWarpSlots] then [CTA][kDbgCtaSlots]
  unsigned long long* counts;     // [kDbgSites] mismatches per site; [kDbgSites - 1] = checks performed (per warp iteration)
  int ctas;                       // CTAs the tag buffer was sized for
  int break_chain;                // negative control: the fused smoothing kernel does not wait for its turn
};
static __device__ DbgState g_sg_dbg;   // one copy per translation unit, attached by dbg_attach()
inline cudaError_t dbg_attach(const DbgState& st) { return cudaMemcpyToSymbol(g_sg_dbg, &st, sizeof(st)); }
__device__ __forceinline__ unsigned* dbg_warp_tags(int warp) {
  return g_sg_dbg.tags + ((size_t)(blockIdx.x % g_sg_dbg.ctas) * kDbgMaxWarps + warp) * kDbgWarpSlots;
}
__device__ __forceinline__ unsigned* dbg_cta_tags() {
  return g_sg_dbg.tags + (size_t)g_sg_dbg.ctas * kDbgMaxWarps * kDbgWarpSlots + (size_t)(blockIdx.x % g_sg_dbg.ctas) * kDbgCtaSlots;
}
__device__ __forceinline__ void dbg_write(unsigned* t, int slot, unsigned epoch) { reinterpret_cast<volatile unsigned*>(t)[slot] = epoch; }
__device__ __forceinline__ void dbg_check(const unsigned* t, int slot, unsigned epoch, int site) {
  if (reinterpret_cast<const volatile unsigned*>(t)[slot] != epoch) atomicAdd(g_sg_dbg.counts + site, 1ull);
}
__device__ __forceinline__ void dbg_count_iteration() { atomicAdd(g_sg_dbg.counts + kDbgSites - 1, 1ull); }
#endif

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

}  // namespace sg
