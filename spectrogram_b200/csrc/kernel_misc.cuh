// Small kernels around the frame kernels: the smoothing recurrence along frames (AnalyserNode
// step 4, 3D/visualizer.js:351,357,362 set tau per mode), the byte -> colour LUT
// (bin/shaders/sonogram-*.shader) and the byte time-domain getter (3D/visualizer.js:363).
#pragma once
#include "common.cuh"

namespace sg {

// X^_t[k] = tau*X^_{t-1}[k] + (1-tau)*|X_t[k]|, per (clip, bin), sequential along frames.
// mags: [n_clips][frames][bins] linear magnitudes (already /N).  state: [n_clips][bins], read as
// the initial X^ and written back with the final one (so chunks of one clip can be chained, and
// the streaming objects keep it across calls).  Arithmetic as Chromium: double, stored as float.
// HBM-bound: 4 B read + elem B written per (frame, bin); coalesced across bins.
template <int OUT>
__global__ void __launch_bounds__(256)
smooth_emit_kernel(const float* __restrict__ mags, typename OutElem<OUT>::type* __restrict__ out,
                   float* __restrict__ state, long long n_clips, long long frames, int bins, double tau,
                   Epilogue ep) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_clips * bins) return;
  const long long clip = idx / bins;
  const int b = (int)(idx - clip * bins);
  const float* __restrict__ m = mags + clip * frames * bins + b;
  typename OutElem<OUT>::type* __restrict__ o = out + clip * frames * bins + b;
  float s = state[idx];
  const double k1 = 1.0 - tau;
  long long t = 0;
  for (; t + 4 <= frames; t += 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(m + (t + u) * bins);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s = finite_or_zero((float)(tau * (double)s + k1 * (double)v[u]));
      o[(t + u) * bins] = emit_mag<OUT>(s, ep);
    }
  }
  for (; t < frames; ++t) {
    s = finite_or_zero((float)(tau * (double)s + k1 * (double)__ldg(m + t * bins)));
    o[t * bins] = emit_mag<OUT>(s, ep);
  }
  state[idx] = s;
}

// ---- time-parallel form of the same recurrence for long clips with few (clip, bin) pairs -------------
// The recurrence is linear with a constant coefficient, so a clip's frames are cut into chunks of
// `chunk` frames:  (A) each chunk's zero-state response at its last frame, (B) a short sequential pass
// over chunks that turns those into the true state at every chunk start, (C) the chunks re-run in
// parallel from their true initial state, emitting dB / bytes.  Inputs are finite (the frame kernel
// maps non-finite magnitudes to 0), so the [SPEC] non-finite rule cannot fire inside the scan.
// carry: [n_clips][n_chunks][bins] float.
__global__ void __launch_bounds__(256)
scan_chunk_sums_kernel(const float* __restrict__ mags, float* __restrict__ carry, long long n_clips, long long frames,
                       int bins, int chunk, long long n_chunks, double tau) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_clips * n_chunks * bins) return;
  const int b = (int)(idx % bins);
  const long long cj = idx / bins, clip = cj / n_chunks, j = cj - clip * n_chunks;
  const long long t0 = j * chunk, t1 = min(frames, t0 + (long long)chunk);
  const float* __restrict__ m = mags + clip * frames * bins + b;
  const double k1 = 1.0 - tau;
  double s = 0.0;
  for (long long t = t0; t < t1; ++t) s = tau * s + k1 * (double)__ldg(m + t * bins);
  carry[idx] = (float)s;
}

// carry[j] <- state at the START of chunk j (in place); state[clip][bin] holds the clip's initial state
__global__ void __launch_bounds__(256)
scan_chunk_carry_kernel(float* __restrict__ carry, const float* __restrict__ state, long long n_clips, long long frames,
                        int bins, int chunk, long long n_chunks, double tau) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_clips * bins) return;
  const long long clip = idx / bins;
  const int b = (int)(idx - clip * bins);
  float* __restrict__ c = carry + clip * n_chunks * bins + b;
  double s = (double)state[idx];
  for (long long j = 0; j < n_chunks; ++j) {
    const long long len = min((long long)chunk, frames - j * chunk);
    const double local = (double)c[j * bins];
    c[j * bins] = (float)s;
    s = pow(tau, (double)len) * s + local;
  }
}

template <int OUT>
__global__ void __launch_bounds__(256)
scan_chunk_emit_kernel(const float* __restrict__ mags, typename OutElem<OUT>::type* __restrict__ out,
                       const float* __restrict__ carry, float* __restrict__ state, long long n_clips, long long frames,
                       int bins, int chunk, long long n_chunks, double tau, Epilogue ep) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_clips * n_chunks * bins) return;
  const int b = (int)(idx % bins);
  const long long cj = idx / bins, clip = cj / n_chunks, j = cj - clip * n_chunks;
  const long long t0 = j * chunk, t1 = min(frames, t0 + (long long)chunk);
  const float* __restrict__ m = mags + clip * frames * bins + b;
  typename OutElem<OUT>::type* __restrict__ o = out + clip * frames * bins + b;
  const double k1 = 1.0 - tau;
  float s = carry[idx];
  for (long long t = t0; t < t1; ++t) {
    s = finite_or_zero((float)(tau * (double)s + k1 * (double)__ldg(m + t * bins)));
    o[t * bins] = emit_mag<OUT>(s, ep);
  }
  if (j == n_chunks - 1) state[clip * bins + b] = s;
}

// re-emit a stored state vector (getByte/FloatFrequencyData called twice in one render quantum)
template <int OUT>
__global__ void emit_state_kernel(const float* __restrict__ state, typename OutElem<OUT>::type* __restrict__ out,
                                  long long n, Epilogue ep) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = emit_mag<OUT>(state[i], ep);
}

__global__ void lut_kernel(const uint8_t* __restrict__ in, uint32_t* __restrict__ out, long long n,
                           const uint32_t* __restrict__ lut) {
  __shared__ uint32_t s_lut[256];
  if (threadIdx.x < 256) s_lut[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = s_lut[in[i]];
}

// b = (unsigned char) clamp(128 * (x + 1), 0, 255)   [SPEC getByteTimeDomainData]
__global__ void time_domain_byte_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint8_t)byte_from_scaled(128.f * (x[i] + 1.f));
}

}  // namespace sg
