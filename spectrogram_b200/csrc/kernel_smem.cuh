// Generic frame kernel: one CTA per frame at a time, the whole complex half-length FFT in shared
// memory, in-place mixed-radix (2,3,4,5) decimation in frequency.  Covers every legal n_fft
// (4..32768, n_fft/2 = 2^a 3^b 5^c, e.g. 400) and every alignment; the warp-resident kernels in
// kernel_w32.cuh take over for the shapes they are specialised for.
//
// Roofline: HBM-bound nominally (4*hop + elem*bins bytes per frame) but at these sizes the
// shared-memory passes bind; this kernel is the coverage path, not the headline path.
#pragma once
#include "common.cuh"
#include "plans.cuh"

namespace sg {

template <int R>
__device__ __forceinline__ void dft_small(float2 (&x)[R]) {
  if constexpr (R == 2) {
    float2 a = x[0], b = x[1];
    x[0] = make_float2(a.x + b.x, a.y + b.y);
    x[1] = make_float2(a.x - b.x, a.y - b.y);
  } else if constexpr (R == 4) {
    float2 s0 = make_float2(x[0].x + x[2].x, x[0].y + x[2].y);
    float2 s1 = make_float2(x[0].x - x[2].x, x[0].y - x[2].y);
    float2 s2 = make_float2(x[1].x + x[3].x, x[1].y + x[3].y);
    float2 s3 = make_float2(x[1].x - x[3].x, x[1].y - x[3].y);
    x[0] = make_float2(s0.x + s2.x, s0.y + s2.y);
    x[1] = make_float2(s1.x + s3.y, s1.y - s3.x);  // s1 - i*s3
    x[2] = make_float2(s0.x - s2.x, s0.y - s2.y);
    x[3] = make_float2(s1.x - s3.y, s1.y + s3.x);  // s1 + i*s3
  } else if constexpr (R == 3) {
    const float c = -0.5f, s = -0.86602540378443864676f;  // W_3 = c + i*s
    float2 t1 = make_float2(x[1].x + x[2].x, x[1].y + x[2].y);
    float2 t2 = make_float2(x[1].x - x[2].x, x[1].y - x[2].y);
    float2 m = make_float2(fmaf(c, t1.x, x[0].x), fmaf(c, t1.y, x[0].y));
    x[0] = make_float2(x[0].x + t1.x, x[0].y + t1.y);
    // x1 = m + i*s*t2 ; x2 = m - i*s*t2   (i*s*t2 = (-s*t2.y, s*t2.x))
    x[1] = make_float2(fmaf(-s, t2.y, m.x), fmaf(s, t2.x, m.y));
    x[2] = make_float2(fmaf(s, t2.y, m.x), fmaf(-s, t2.x, m.y));
  } else {
    static_assert(R == 5, "radix");
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = -0.95105651629515357212f, s2 = -0.58778525229247312917f;  // W_5^1, W_5^2 imag
    float2 a1 = make_float2(x[1].x + x[4].x, x[1].y + x[4].y);
    float2 b1 = make_float2(x[1].x - x[4].x, x[1].y - x[4].y);
    float2 a2 = make_float2(x[2].x + x[3].x, x[2].y + x[3].y);
    float2 b2 = make_float2(x[2].x - x[3].x, x[2].y - x[3].y);
    float2 x0 = x[0];
    x[0] = make_float2(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
    float2 m1 = make_float2(fmaf(c1, a1.x, fmaf(c2, a2.x, x0.x)), fmaf(c1, a1.y, fmaf(c2, a2.y, x0.y)));
    float2 m2 = make_float2(fmaf(c2, a1.x, fmaf(c1, a2.x, x0.x)), fmaf(c2, a1.y, fmaf(c1, a2.y, x0.y)));
    // n1 = s1*b1 + s2*b2 ; n2 = s2*b1 - s1*b2 ; X1 = m1 + i*n1, X4 = m1 - i*n1, X2 = m2 + i*n2, X3 = m2 - i*n2
    float2 n1 = make_float2(fmaf(s1, b1.x, s2 * b2.x), fmaf(s1, b1.y, s2 * b2.y));
    float2 n2 = make_float2(fmaf(s2, b1.x, -s1 * b2.x), fmaf(s2, b1.y, -s1 * b2.y));
    x[1] = make_float2(m1.x - n1.y, m1.y + n1.x);
    x[4] = make_float2(m1.x + n1.y, m1.y - n1.x);
    x[2] = make_float2(m2.x - n2.y, m2.y + n2.x);
    x[3] = make_float2(m2.x + n2.y, m2.y - n2.x);
  }
}

template <int R>
__device__ __forceinline__ void dif_stage(float2* buf, const float2* __restrict__ tw, int m, int lb) {
  const int sub = lb / R, tstep = m / lb, cnt = m / R;
  for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
    const int blk = j / sub, jj = j - blk * sub, base = blk * lb + jj;
    float2 x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = buf[base + r * sub];
    dft_small<R>(x);
    buf[base] = x[0];
#pragma unroll
    for (int q = 1; q < R; ++q) buf[base + q * sub] = cmul(x[q], __ldg(tw + jj * q * tstep));
  }
}

template <int OUT>
__global__ void __launch_bounds__(1024) stft_smem_kernel(FrameGeom g, SmemPlan pl, Epilogue ep,
                                                         typename OutElem<OUT>::type* __restrict__ out) {
  extern __shared__ float2 buf[];
  const int m = pl.m, bins = m;
  for (long long f = blockIdx.x; f < g.total_frames; f += gridDim.x) {
    const long long clip = f / g.frames_per_clip, t = f - clip * g.frames_per_clip;
    const float* __restrict__ x = g.pcm + clip * g.clip_stride;
    const long long start = g.start0 + t * g.hop;
    // step 1-2: time block, window, pack two reals per complex: z[i] = xw[2i] + i*xw[2i+1]
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const long long s0 = start + 2 * i, s1 = s0 + 1;
      const float a = (s0 >= 0 && s0 < g.clip_len) ? __ldg(x + s0) : 0.f;
      const float b = (s1 >= 0 && s1 < g.clip_len) ? __ldg(x + s1) : 0.f;
      buf[i] = make_float2(a * __ldg(pl.win + 2 * i), b * __ldg(pl.win + 2 * i + 1));
    }
    __syncthreads();
    // step 3: complex FFT of length m, in place
    int lb = m;
    for (int s = 0; s < pl.nstage; ++s) {
      const int r = pl.radix[s];
      if (r == 4) dif_stage<4>(buf, pl.tw, m, lb);
      else if (r == 2) dif_stage<2>(buf, pl.tw, m, lb);
      else if (r == 5) dif_stage<5>(buf, pl.tw, m, lb);
      else dif_stage<3>(buf, pl.tw, m, lb);
      lb /= r;
      __syncthreads();
    }
    // real-input untangle + steps 4-6 (tau == 0) or linear magnitude for the scan kernel
    typename OutElem<OUT>::type* __restrict__ row = out + f * (long long)bins;
    for (int k = threadIdx.x; k <= m / 2; k += blockDim.x) {
      const int mk = (k == 0) ? 0 : m - k;
      const float2 zk = buf[__ldg(pl.pos + k)], zm = buf[__ldg(pl.pos + mk)];
      const float2 e2 = make_float2(zk.x + zm.x, zk.y - zm.y);    // 2E  = Z[k] + conj Z[m-k]
      const float2 o2 = make_float2(zk.y + zm.y, zm.x - zk.x);    // 2O  = -i (Z[k] - conj Z[m-k])
      const float2 t2 = cmul(o2, __ldg(pl.ut + k));               // 2 W_n^k O
      const float xr = e2.x + t2.x, xi = e2.y + t2.y;             // 2 X[k]
      row[k] = emit_power<OUT>(fmaf(xr, xr, xi * xi), ep);
      if (k != 0 && mk != k) {
        const float yr = e2.x - t2.x, yi = e2.y - t2.y;           // 2 conj X[m-k]
        row[mk] = emit_power<OUT>(fmaf(yr, yr, yi * yi), ep);
      }
    }
    __syncthreads();
  }
}

}  // namespace sg
