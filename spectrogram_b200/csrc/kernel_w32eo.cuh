// n_fft = 4096: one warp per frame, the frame's 2048-point complex FFT split (radix-2 decimation in time) into the
// 1024-point FFTs of its even and odd complex samples, which ride in the two halves of the packed FP32 registers.
//
// The packed frame-pair kernels (kernel_w32x2p.cuh) run two independent 1024-point FFTs per warp, one per
// register half -- there the two halves are two consecutive frames.  Here they are the two sub-transforms of ONE
// frame:  z[m] = x[2m] + i x[2m+1],  E = FFT1024(z[2m']),  O = FFT1024(z[2m'+1]),  and
//     Z[k'] = E[k'] + W_2048^k' O[k'],   Z[k'+1024] = E[k'] - W_2048^k' O[k']          (k' < 1024)
// so the whole 5 + 5 stage network, the 32x32 exchange and the pass-2 twiddle generation are the pair kernel's,
// unchanged, and stay packed.  What differs:
//   loader     one 16-byte load brings (x[4m'], x[4m'+1] | x[4m'+2], x[4m'+3]) = the E and the O input of element
//              m': 32 coalesced LDG.128 per lane per frame, any hop that keeps frames 16-byte aligned
//   combine    the radix-2 step across the two register halves (6 scalar FMAs per element: the same FP32-pipe
//              cycles as one packed stage), twiddles W_2048^lane * W_64^i from one per-lane base
//   untangle   X[k'] pairs with Z[2048-k'] = the partner lane's MINUS half, X[k'+1024] with its PLUS half: the
//              mirror shuffles swap the halves, and the twiddle of the upper half is -i times the lower one's,
//              so one loaded float2 (w.x, w.y) is the packed x-operand and (w.y, -w.x) the packed y-operand
//   output     the halves of |X|^2 are bins k' and k'+1024 of the same row (mirror slots: 2048-k' and 1024-k')
// Frames the loader cannot express (clip edges / zero history, misaligned starts) take guarded scalar loads.
#pragma once
#include "kernel_w32x2p.cuh"

namespace sg {

constexpr int kEoBins = kEoN / 2;
constexpr int kEoWarps = 8;    // 8 warps x 255 registers measured faster than 12 x 168 (u8 242 vs 237 M frames/s, float dB 238 vs 209 M)
constexpr int kEoTableBytes = 1024 * 16 + 5 * 32 * 8 + 16 * 32 * 8;     // window (float4) + 5 base twiddles + untangle
constexpr int kEoSmemBytes = kEoTableBytes + kEoWarps * kXpPlaneBytes;

__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// window + stage 1 for elements j and j + 16 (x0, x1: their float4 loads; w0, w1: the window at the same places)
__device__ __forceinline__ void window_stage1_eo(C2& lo, C2& hi, float4 x0, float4 x1, float4 w0, float4 w1) {
  const float te = x0.x * w0.x, ue = x0.y * w0.y, to = x0.z * w0.z, uo = x0.w * w0.w;
  lo.re = P2(fmaf(x1.x, w1.x, te), fmaf(x1.z, w1.z, to));
  hi.re = P2(fmaf(x1.x, -w1.x, te), fmaf(x1.z, -w1.z, to));
  lo.im = P2(fmaf(x1.y, w1.y, ue), fmaf(x1.w, w1.w, uo));
  hi.im = P2(fmaf(x1.y, -w1.y, ue), fmaf(x1.w, -w1.w, uo));
}

// 2X[k] (lower half: k = k', upper half: k = k' + 1024) and 2 conj X[2048 - k] from zk = (Z+[k'], Z-[k']) and
// zm = (Z-[1024-k'], Z+[1024-k']);  w = W_4096^k'
__device__ __forceinline__ void untangle_eo(const C2& zk, const C2& zm, float2 w, P2& pk, P2& pm) {
  const P2 wx = P2(w.x, w.y), wy = P2(w.y, -w.x);
  const P2 ex = add2(zk.re, zm.re), ey = add2(zk.im, neg(zm.im));          // 2E
  const P2 ox = add2(zk.im, zm.im), oy = add2(zm.re, neg(zk.re));          // 2O
  const P2 xr = fma2(ox, wx, fma2(oy, neg(wy), ex));
  const P2 xi = fma2(ox, wy, fma2(oy, wx, ey));
  const P2 yr = fma2(ex, bc(2.f), neg(xr));
  const P2 yi = fma2(ey, bc(2.f), neg(xi));
  pk = fma2(xr, xr, mul2(xi, xi));
  pm = fma2(yr, yr, mul2(yi, yi));
}

template <int OUT, int NW = kEoWarps>
__global__ void __launch_bounds__(NW * 32, 1) __maxnreg__(XpShape<NW>::kMaxRegs)
stft_w32eo_kernel(FrameGeom g, EoPlan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using T = typename OutElem<OUT>::type;
  extern __shared__ float4 smem_raw[];
  float4* s_win4 = smem_raw;                                           // [1024] (w[4m] .. w[4m+3])
  float2* s_twb = reinterpret_cast<float2*>(s_win4 + 1024);            // [5][32]  W_{32*2^u}^lane
  float2* s_ut = s_twb + 5 * 32;                                       // [16][32] W_4096^{lane + 32 i}
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wbase = reinterpret_cast<unsigned char*>(s_ut + 16 * 32) + warp * kXpPlaneBytes;
  float4* xp = reinterpret_cast<float4*>(wbase);                       // exchange planes
  uint16_t* sb16 = reinterpret_cast<uint16_t*>(wbase);                 // byte stage, aliases the planes

  {
    const float4* w4 = reinterpret_cast<const float4*>(pl.win);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_win4[i] = __ldg(w4 + i);
    for (int i = threadIdx.x; i < 5 * 32; i += blockDim.x) s_twb[i] = __ldg(pl.tw2 + ((1 << (i >> 5)) - 1) * 32 + (i & 31));
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) s_ut[i] = __ldg(pl.tab + 32 + i);
  }
  __syncthreads();

  const int partner = (32 - lane) & 31;
  const bool lane0 = lane == 0;
  const long long fstep = (long long)gridDim.x * NW;
  const bool base_aligned = (reinterpret_cast<uintptr_t>(g.pcm) & 15) == 0;

  // geometry of one frame: clip-relative start, offset of its first sample from g.pcm, and whether the 16-byte
  // loader can express it (inside the clip, aligned)
  struct Fr { long long f, off; bool fast; };
  auto frame_at = [&](long long f) {
    Fr r;
    r.f = f;
    const long long clip = f / g.frames_per_clip;
    const long long start = g.start0 + (f - clip * g.frames_per_clip) * g.hop;
    r.off = clip * g.clip_stride + start;
    r.fast = f < g.total_frames && start >= 0 && start + kEoN <= g.clip_len && base_aligned && (r.off & 3) == 0;
    return r;
  };
  Fr cur = frame_at((long long)blockIdx.x * NW + warp);
  if (cur.f >= g.total_frames) return;

  // Samples of the current frame (fast path): s[m] = float4 #(lane + 32 e(m)) of the span, e(m) = the element order
  // the window stage consumes (pairs (j, j + 16), j ascending).  Loads are unconditional -- a frame the loader
  // cannot express reads the window table instead and ignores it -- and are issued for the NEXT frame while this
  // one is untangled and written, as its FFT registers die.
  float4 s[32];
  const float4* idle_src = reinterpret_cast<const float4*>(pl.win) + lane;   // 4096 readable floats
  auto elem_of = [](int m) { return (m >> 1) + 16 * (m & 1); };              // m = 2j -> j, 2j + 1 -> j + 16
  {
    const float4* src = cur.fast ? reinterpret_cast<const float4*>(g.pcm + cur.off) + lane : idle_src;
    static_for<0, 32>([&](auto mm) { constexpr int m = decltype(mm)::value; s[m] = ldg_nc_f4(src + 32 * elem_of(m)); });
  }

  while (true) {
    const long long f = cur.f;
    // ---- steps 1-2 (+ FFT stage 1): window, even/odd complex samples into the register halves, bit-reversed
    C2 a[32];
    if (cur.fast) {
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);   // r1 == r0 + 1
        window_stage1_eo(a[r0], a[r1], s[2 * j], s[2 * j + 1], s_win4[lane + 32 * j], s_win4[lane + 32 * (j + 16)]);
      });
    } else {
      const long long clip = f / g.frames_per_clip;
      const long long start = g.start0 + (f - clip * g.frames_per_clip) * g.hop;
      const float* __restrict__ x = g.pcm + clip * g.clip_stride;
      auto ld = [&](long long q) { return (q >= 0 && q < g.clip_len) ? __ldg(x + q) : 0.f; };
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
        const long long o0 = start + 4 * (lane + 32 * j), o1 = start + 4 * (lane + 32 * (j + 16));
        window_stage1_eo(a[r0], a[r1], make_float4(ld(o0), ld(o0 + 1), ld(o0 + 2), ld(o0 + 3)),
                         make_float4(ld(o1), ld(o1 + 1), ld(o1 + 2), ld(o1 + 3)), s_win4[lane + 32 * j],
                         s_win4[lane + 32 * (j + 16)]);
      });
    }

    // ---- pass 1: stages 2-5 in registers, compile-time twiddles
    dit2_stage_const<2>(a);
    dit2_stage_const<3>(a);
    dit2_stage_const<4>(a);
    dit2_stage_const<5>(a);

    // ---- exchange (32x32 transpose), as in the frame-pair kernel: row-paired planes, STS.64 in, LDS.128 out
    {
      float2* wre = reinterpret_cast<float2*>(xp) + ((lane >> 1) * kXpStride) * 2 + (lane & 1);
      float2* wim = wre + 16 * kXpStride * 2;
      static_for<0, 32>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        wre[2 * q] = a[q].re.v;
        wim[2 * q] = a[q].im.v;
      });
      asm volatile("bar.sync %0, 32;" ::"r"(warp + 1) : "memory");
      const float4* rre = xp + lane;
      const float4* rim = rre + 16 * kXpStride;
      static_for<0, 16>([&](auto qq) {
        constexpr int q0 = decltype(qq)::value;
        constexpr int j = bitrev(q0, 4);
        const float4 vr = rre[j * kXpStride], vi = rim[j * kXpStride];
        a[q0].re = P2(vr.x, vr.y); a[q0 + 16].re = P2(vr.z, vr.w);
        a[q0].im = P2(vi.x, vi.y); a[q0 + 16].im = P2(vi.z, vi.w);
      });
      __syncwarp();
    }

    // ---- pass 2: stages 6-10, twiddles from the five per-lane bases
    dit2_stage_gen<1>(a, s_twb[0 * 32 + lane]);
    dit2_stage_gen<2>(a, s_twb[1 * 32 + lane]);
    dit2_stage_gen<3>(a, s_twb[2 * 32 + lane]);
    dit2_stage_gen<4>(a, s_twb[3 * 32 + lane]);
    dit2_stage_gen<5>(a, s_twb[4 * 32 + lane]);
    // now a[i] = (E[k'], O[k']), k' = lane + 32 i

    // now a[i] = (E[k'], O[k']), k' = lane + 32 i.
    // a non-finite sample makes every E or every O non-finite: one test per frame. byte scale -> NaN -> byte 0
    const P2 poison2 = fma2(a[0].re, bc(0.f), mul2(a[0].im, bc(0.f)));
    const P2 poison = bc(poison2.v.x + poison2.v.y);

    // ---- next frame: geometry now; its first 16 loads ride in the untangle steps, the other 16 in the epilogue steps
    const Fr nxt = frame_at(f + fstep);
    const bool has_next = nxt.f < g.total_frames;
    const float4* nsrc = nxt.fast ? reinterpret_cast<const float4*>(g.pcm + nxt.off) + lane : idle_src;

    // ---- combine + untangle.  Z[k'] = E + wO and Z[k'+1024] = E - wO (w = W_2048^k') are formed where they are
    //      consumed: a lane's elements 16..31 are only ever used by its partner lane (they are the mirrors
    //      1024 - k' of the partner's k'), so the raw (E, O) pairs are fetched in place exactly as in the frame-pair
    //      kernel and the receiving lane applies the mirror's twiddle W_2048^(1024-k') = (-w.x, w.y) itself.
    //      X[k'] pairs with Z[2048-k'] = the mirror's MINUS value and X[k'+1024] with its PLUS value, so the mirror
    //      is combined straight into swapped halves.  k' = 0 needs no special case (its mirror element is itself with
    //      twiddle -1, which swaps the halves back).  Lane 0's element 16 (k' = 512, its own mirror) is taken first
    //      and rides in lane 0's spare mirror slot.
    auto combine = [&](const C2& z, float wx, float wy, bool swap) {   // (E, O) -> (E + wO, E - wO), or swapped
      const float er = z.re.v.x, ei = z.im.v.x, orr = z.re.v.y, oi = z.im.v.y;
      const float pr = fmaf(orr, wx, fmaf(oi, -wy, er));
      const float pi = fmaf(oi, wx, fmaf(orr, wy, ei));
      const float mr = fmaf(er, 2.f, -pr), mi = fmaf(ei, 2.f, -pi);
      C2 r;
      r.re = swap ? P2(mr, pr) : P2(pr, mr);
      r.im = swap ? P2(mi, pi) : P2(pi, mi);
      return r;
    };
    P2 p512;
    {
      P2 unused;   // W_2048^512 = -i; W_4096^512 = exp(-i pi/4)
      untangle_eo(combine(a[16], 0.f, -1.f, false), combine(a[16], 0.f, -1.f, true),
                  make_float2(0.70710678118654752440f, -0.70710678118654752440f), p512, unused);
    }
    static_for<0, 16>([&](auto ii) {
      constexpr int i = 15 - decltype(ii)::value;
      constexpr int src = 31 - i, own = (32 - i) & 31;
      const float mra = __shfl_sync(0xffffffffu, a[src].re.v.x, partner);
      const float mrb = __shfl_sync(0xffffffffu, a[src].re.v.y, partner);
      const float mia = __shfl_sync(0xffffffffu, a[src].im.v.x, partner);
      const float mib = __shfl_sync(0xffffffffu, a[src].im.v.y, partner);
      a[src].re = P2(lane0 ? a[own].re.v.x : mra, lane0 ? a[own].re.v.y : mrb);
      a[src].im = P2(lane0 ? a[own].im.v.x : mia, lane0 ? a[own].im.v.y : mib);
    });
    P2 pk[16], pm[16];   // pk[i] = |2X|^2 at (k', k' + 1024);  pm[i] at (2048 - k', 1024 - k')
    static_for<0, 16>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      const float2 u = s_ut[i * 32 + lane];                                  // W_4096^k'
      const float wx = fmaf(u.x, u.x, -u.y * u.y), wy = (u.x + u.x) * u.y;    // W_2048^k' = its square
      untangle_eo(combine(a[i], wx, wy, false), combine(a[31 - i], -wx, wy, true), u, pk[i], pm[i]);
      if constexpr (i == 0) {
        // lane 0: the mirrors of k' = 0 are the Nyquist bin (dropped) and bin 1024 again; the slot carries bins
        // 512 / 1536 (stored like every mirror slot: lower half -> upper row)
        pm[0] = P2(lane0 ? p512.v.y : pm[0].v.x, lane0 ? p512.v.x : pm[0].v.y);
      }
      s[i] = ldg_nc_f4(nsrc + 32 * elem_of(i));
    });

    // ---- epilogue: lower halves of pk -> bins k', upper -> k' + 1024; pm the other way round at 1024 - k'
    T* __restrict__ row_lo = out + f * (long long)kEoBins;
    T* __restrict__ row_hi = row_lo + 1024;
    if constexpr (OUT == kOutU8 || OUT == kOutRgba8) {
      const P2 scale = add2(bc(ep.byte_a), poison);
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        const int k = lane + 32 * i;
        int mk = 1024 - k;
        if constexpr (i == 0) { if (lane0) mk = 512; }
        const P2 vk = fma2(P2(lg2_ftz(pk[i].v.x), lg2_ftz(pk[i].v.y)), scale, bc(ep.byte_b));
        const P2 vm = fma2(P2(lg2_ftz(pm[i].v.x), lg2_ftz(pm[i].v.y)), scale, bc(ep.byte_b));
        const unsigned k_lo = byte_of_scaled(vk.v.x), k_hi = byte_of_scaled(vk.v.y);
        const unsigned m_hi = byte_of_scaled(vm.v.x), m_lo = byte_of_scaled(vm.v.y);
        if constexpr (OUT == kOutU8) {
          sb16[k] = (uint16_t)(k_lo | (k_hi << 8));
          sb16[mk] = (uint16_t)(m_lo | (m_hi << 8));
        } else {
          row_lo[k] = __ldg(ep.lut + k_lo); row_lo[mk] = __ldg(ep.lut + m_lo);
          row_hi[k] = __ldg(ep.lut + k_hi); row_hi[mk] = __ldg(ep.lut + m_hi);
        }
        s[16 + i] = ldg_nc_f4(nsrc + 32 * elem_of(16 + i));
      });
      if constexpr (OUT == kOutU8) {
        __syncwarp();
        // de-interleave the (lower, upper) byte pairs: 8 bins per lane per round, 8-byte coalesced stores
        const uint4* s16 = reinterpret_cast<const uint4*>(sb16);
        uint2* ra = reinterpret_cast<uint2*>(row_lo);
        uint2* rb = reinterpret_cast<uint2*>(row_hi);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 w = s16[c * 32 + lane];
          ra[c * 32 + lane] = make_uint2(__byte_perm(w.x, w.y, 0x6420), __byte_perm(w.z, w.w, 0x6420));
          rb[c * 32 + lane] = make_uint2(__byte_perm(w.x, w.y, 0x7531), __byte_perm(w.z, w.w, 0x7531));
        }
      }
    } else {
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        const int k = lane + 32 * i;
        int mk = 1024 - k;
        if constexpr (i == 0) { if (lane0) mk = 512; }
        {
          // packed dB / magnitude; the non-finite rule from the per-frame flag (a poisoned frame reads 0 / -inf)
          const bool bad = !(poison.v.x == 0.f);
          const P2 vk = float_of_power<OUT>(pk[i], ep), vm = float_of_power<OUT>(pm[i], ep);
          const float z = float_of_poisoned<OUT>();
          row_lo[k] = bad ? z : vk.v.x; row_hi[k] = bad ? z : vk.v.y;
          row_hi[mk] = bad ? z : vm.v.x; row_lo[mk] = bad ? z : vm.v.y;
        }
        s[16 + i] = ldg_nc_f4(nsrc + 32 * elem_of(16 + i));
      });
    }
    __syncwarp();
    if (!has_next) break;
    cur = nxt;
  }
}

}  // namespace sg
