// Frame-pair kernels for n_fft = 1024 (L = 16 lanes per pair), 512 (L = 8) and 256 (L = 4): the n_fft 2048 kernel of
// kernel_w32x2p.cuh folded onto part of a warp.
//
// n_fft = 64 L, M = 32 L complex points = 32 x L.  A pair of consecutive frames (A, B) is owned by L lanes -- 32/L
// pairs per warp -- and every lane keeps 32 complex points of both frames in 64-bit register pairs, so every
// butterfly is one packed FFMA2 / FADD2 exactly as in the n_fft 2048 kernel:
//   loader    lane t holds z[t + L j]: 8 L contiguous bytes per pair per load; frame B's element j is element
//             j + HOPJ of the same lane (hop = 2 L HOPJ samples), so the two frames share all but HOPJ of their 32
//             loads; the NEXT pair's loads are issued during the untangle
//   pass 1    32-point DIT in registers (stage 1 fused with the window, stages 2-5 compile-time twiddles)
//   xchg      L x 32 tile per pair in two row-paired planes (STS.64 / LDS.128, conflict free)
//   pass 2    32 columns of L-point FFTs; a lane takes 32/L columns as mirror pairs (c, 32 - c) -- lane 0's first
//             pair is the self-mirrored columns 0 and 16 -- so Z[k] and Z[M - k] meet in the same lane: no shuffles,
//             no second exchange; twiddles W_{2^u}^p * W_{32*2^u}^col are built from per-column bases
//   epilogue  as in the 2048 kernel (per-frame non-finite decision, MUFU.LG2, FFMA2, cvt.sat, staged byte rows)
// Algorithmic bytes per frame: 4*hop + elem*M.
#pragma once
#include "common.cuh"
#include "ct_math.cuh"
#include "kernel_w32.cuh"
#include "kernel_w32x2.cuh"
#include "kernel_w32x2p.cuh"

namespace sg {

constexpr int kPairWarps = 12;        // byte outputs; float rows (store-heavy) run 8 warps x 255 registers
constexpr int kPairWarpsFloat = 8;
template <int OUT> constexpr int pair_warps() { return (OUT == kOutF32Db || OUT == kOutF32Mag) ? kPairWarpsFloat : kPairWarps; }

template <int LOG2L>
struct PairShape {
  static constexpr int L = 1 << LOG2L;               // lanes per frame pair = points of a pass-2 FFT
  static constexpr int N = 64 * L, M = 32 * L;
  static constexpr int PW = 32 / L;                  // pairs per warp
  static constexpr int NCOL = 32 / L, NP = NCOL / 2; // columns / mirror pairs per lane
  static constexpr int kPlaneUnits = (L / 2) * kXpStride;                       // 16-byte units per plane
  // pairs that share a half-warp (L = 8: two, L = 4: four) are offset by 16 / 8 banks so their 8-byte stores do
  // not collide
  static constexpr int kPairBytes = 2 * kPlaneUnits * 16 + (L == 16 ? 0 : L == 8 ? 64 : 96);
  static constexpr int kWarpBytes = PW * kPairBytes;
  static constexpr int kUtEntries = M / 2 + 2;
  static constexpr int kTableBytes = 16 * L * 16 + LOG2L * 32 * 8 + kUtEntries * 8;   // window quads, bases, W_N^k
  static constexpr int smem_bytes(int warps) { return kTableBytes + warps * kWarpBytes; }
  static_assert(kTableBytes % 16 == 0 && kPairBytes % 16 == 0, "16-byte alignment of the planes");
};

// stage U of an NPTS-point pass 2 on a[OFF .. OFF+NPTS): butterflies (i0, i0 + half), twiddle W_{2 half}^p * base
template <int U, int OFF, int NPTS>
__device__ __forceinline__ void dit2_stage_gen_n(C2 (&a)[32], float2 base) {
  constexpr int half = 1 << (U - 1);
  static_for<0, half>([&](auto pp) {
    constexpr int p = decltype(pp)::value;
    const float2 w = twiddle_times<p, 2 * half>(base);
    static_for<0, NPTS / (2 * half)>([&](auto bb) {
      constexpr int i0 = OFF + decltype(bb)::value * 2 * half + p;
      bfly2(a[i0], a[i0 + half], w.x, w.y);
    });
  });
}

template <int OUT, int LOG2L, int HOPJ>
__global__ void __launch_bounds__(pair_warps<OUT>() * 32, 1)
stft_pair_kernel(FrameGeom g, PairPlan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using T = typename OutElem<OUT>::type;
  using S = PairShape<LOG2L>;
  constexpr int L = S::L, N = S::N, M = S::M, NCOL = S::NCOL, NP = S::NP;
  constexpr int HOP = 2 * L * HOPJ, NLOAD = 32 + HOPJ;
  extern __shared__ float4 smem_raw[];
  float4* s_win4 = smem_raw;                                           // [16][L] (w2[t+Lj], w2[t+L(j+16)])
  float2* s_twb = reinterpret_cast<float2*>(s_win4 + 16 * L);          // [LOG2L][32]  W_{32*2^u}^col
  float2* s_ut = s_twb + LOG2L * 32;                                   // [M/2 + 2]    W_N^k
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> LOG2L, t = lane & (L - 1);
  unsigned char* wbase = reinterpret_cast<unsigned char*>(s_ut + S::kUtEntries) + warp * S::kWarpBytes + h * S::kPairBytes;
  float4* xp = reinterpret_cast<float4*>(wbase);                       // this pair's planes (re, then im)
  uint16_t* sb16 = reinterpret_cast<uint16_t*>(wbase);                 // byte stage aliases the planes

  {
    const float2* w2 = reinterpret_cast<const float2*>(pl.win);
    for (int i = threadIdx.x; i < 16 * L; i += blockDim.x) {
      const int j = i >> LOG2L, l = i & (L - 1);
      const float2 lo = __ldg(w2 + l + L * j), hi = __ldg(w2 + l + L * (j + 16));
      s_win4[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    for (int i = threadIdx.x; i < LOG2L * 32; i += blockDim.x) s_twb[i] = __ldg(pl.twb + i);
    for (int i = threadIdx.x; i <= M / 2; i += blockDim.x) s_ut[i] = __ldg(pl.ut + i);
  }
  __syncthreads();

  // pair geometry of THIS lane group, advanced incrementally
  const int fpc = (int)g.frames_per_clip;
  constexpr int NW = pair_warps<OUT>();
  const int step = 2 * S::PW * gridDim.x * NW;   // frames between a lane group's consecutive pairs
  const int step_clip = step / fpc, step_t = step - step_clip * fpc;
  const long long d_off = (long long)step_clip * g.clip_stride + (long long)step_t * HOP;
  const long long wrap_off = g.clip_stride - (long long)fpc * HOP;
  const unsigned pcm_lo = (unsigned)reinterpret_cast<uintptr_t>(g.pcm);
  int t_lo, t_hi;
  {
    const long long lo = g.start0 >= 0 ? 0 : (-g.start0 + HOP - 1) / HOP;
    const long long room = g.clip_len - (HOP + N) - g.start0;
    const long long hi = room < 0 ? -1 : min((long long)fpc - 2, room / HOP);
    t_lo = (int)lo;
    t_hi = (int)hi;
  }
  long long fa = 2 * S::PW * ((long long)blockIdx.x * NW + warp) + 2 * h;
  if (fa - 2 * h >= g.total_frames) return;              // warp-uniform: the warp's first pair has no frame
  int clip = (int)(min(fa, g.total_frames - 1) / fpc);
  int tt = (int)(min(fa, g.total_frames - 1) - (long long)clip * fpc);
  long long off = clip * g.clip_stride + g.start0 + (long long)tt * HOP;
  auto is_fast = [&](long long f, int tq, long long o) {
    return f + 1 < g.total_frames && tq >= t_lo && tq <= t_hi && ((pcm_lo + ((unsigned)o << 2)) & 7) == 0;
  };
  bool cur_fast = is_fast(fa, tt, off);
  const bool c0 = t == 0;
  // mirror pair r of this lane: columns (ka, kb) = (t + L r, 32 - (t + L r)); lane 0's pair 0 is (0, 16)
  int cols[NCOL];
  static_for<0, NP>([&](auto rr) {
    constexpr int r = decltype(rr)::value;
    cols[2 * r] = t + L * r;
    cols[2 * r + 1] = (r == 0 && c0) ? 16 : 32 - (t + L * r);
  });
  auto col_a = [&](int r) { return cols[2 * r]; };
  auto col_b = [&](int r) { return cols[2 * r + 1]; };

  float2 s[NLOAD];
  const float2* idle_src = reinterpret_cast<const float2*>(pl.win) + t;   // readable past the window (build_plan)
  {
    const float2* src = cur_fast ? reinterpret_cast<const float2*>(g.pcm + off) + t : idle_src;
    static_for<0, NLOAD>([&](auto mm) { constexpr int m = decltype(mm)::value; s[m] = ldg_nc_f2(src + L * m); });
  }

  while (true) {
    const bool alive = fa < g.total_frames;   // a lane group past the end keeps marching (warp barrier), stores nothing
    // ---- steps 1-2 (+ FFT stage 1)
    C2 a[32];
    if (cur_fast) {
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
        const float4 w = s_win4[j * L + t];
        window_stage1(a[r0], a[r1], s[j], s[j + 16], s[j + HOPJ], s[j + 16 + HOPJ], make_float2(w.x, w.y),
                      make_float2(w.z, w.w));
      });
    } else {
      const long long fq = alive ? fa : g.total_frames - 1;
      const int cq = (int)(fq / fpc), tq = (int)(fq - (long long)cq * fpc);
      const bool has_b = fq + 1 < g.total_frames;
      int clip_b = cq, tb = tq;
      if (has_b) { if (tq + 1 == fpc) { ++clip_b; tb = 0; } else ++tb; }
      const float* __restrict__ xa = g.pcm + cq * g.clip_stride;
      const float* __restrict__ xb = g.pcm + clip_b * g.clip_stride;
      const long long start_a = g.start0 + (long long)tq * HOP, start_b = g.start0 + (long long)tb * HOP;
      auto ld = [&](const float* __restrict__ x, long long q) { return (q >= 0 && q < g.clip_len) ? __ldg(x + q) : 0.f; };
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
        const long long o0 = 2 * (t + L * j), o1 = 2 * (t + L * (j + 16));
        const float4 w = s_win4[j * L + t];
        window_stage1(a[r0], a[r1], make_float2(ld(xa, start_a + o0), ld(xa, start_a + o0 + 1)),
                      make_float2(ld(xa, start_a + o1), ld(xa, start_a + o1 + 1)),
                      make_float2(ld(xb, start_b + o0), ld(xb, start_b + o0 + 1)),
                      make_float2(ld(xb, start_b + o1), ld(xb, start_b + o1 + 1)), make_float2(w.x, w.y),
                      make_float2(w.z, w.w));
      });
    }

    // ---- pass 1: stages 2-5
    dit2_stage_const<2>(a);
    dit2_stage_const<3>(a);
    dit2_stage_const<4>(a);
    dit2_stage_const<5>(a);

    // ---- exchange: rows = lanes of the pair (L), columns = k1 (32); row pairs interleaved in 16-byte units.
    //      After it, a[ci*L + q'] = tile[row bitrev(q')][column ci] for this lane's NCOL columns.
    {
      float2* wre = reinterpret_cast<float2*>(xp) + ((t >> 1) * kXpStride) * 2 + (t & 1);
      float2* wim = wre + S::kPlaneUnits * 2;
      static_for<0, 32>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        wre[2 * q] = a[q].re.v;
        wim[2 * q] = a[q].im.v;
      });
      asm volatile("bar.sync %0, 32;" ::"r"(warp + 1) : "memory");
      const float4* rre = xp;
      const float4* rim = rre + S::kPlaneUnits;
      static_for<0, NP>([&](auto rr) {
        constexpr int r = decltype(rr)::value;
        const int ka = col_a(r), kb = col_b(r);
        static_for<0, L / 2>([&](auto qq) {   // rows 2j, 2j+1 hold q' = bitrev(j), bitrev(j) + L/2
          constexpr int q0 = decltype(qq)::value;
          constexpr int j = bitrev(q0, LOG2L - 1);
          constexpr int oa = (2 * r) * L, ob = (2 * r + 1) * L;
          const float4 ar = rre[j * kXpStride + ka], ai = rim[j * kXpStride + ka];
          const float4 br = rre[j * kXpStride + kb], bi = rim[j * kXpStride + kb];
          a[oa + q0].re = P2(ar.x, ar.y); a[oa + q0 + L / 2].re = P2(ar.z, ar.w);
          a[oa + q0].im = P2(ai.x, ai.y); a[oa + q0 + L / 2].im = P2(ai.z, ai.w);
          a[ob + q0].re = P2(br.x, br.y); a[ob + q0 + L / 2].re = P2(br.z, br.w);
          a[ob + q0].im = P2(bi.x, bi.y); a[ob + q0 + L / 2].im = P2(bi.z, bi.w);
        });
      });
      __syncwarp();
    }

    // ---- pass 2: L-point DIT on each of this lane's columns
    static_for<1, LOG2L + 1>([&](auto uu) {
      constexpr int u = decltype(uu)::value;
      static_for<0, NP>([&](auto rr) {
        constexpr int r = decltype(rr)::value;
        dit2_stage_gen_n<u, (2 * r) * L, L>(a, s_twb[(u - 1) * 32 + col_a(r)]);
        dit2_stage_gen_n<u, (2 * r + 1) * L, L>(a, s_twb[(u - 1) * 32 + col_b(r)]);
      });
    });
    // now a[(2r)L + q] = Z[ka_r + 32 q], a[(2r+1)L + q] = Z[kb_r + 32 q] of both frames

    const P2 poison = fma2(a[0].re, bc(0.f), mul2(a[0].im, bc(0.f)));   // 0 or NaN per frame
    // lane 0: bin M/2 = conj Z[M/2] is element L/2 of column 0
    const P2 pmid = mul2(bc(4.f), fma2(a[L / 2].re, a[L / 2].re, mul2(a[L / 2].im, a[L / 2].im)));

    // ---- next pair of this lane group
    long long nfa = fa + step, noff = off + d_off;
    int nclip = clip + step_clip, ntt = tt + step_t;
    if (ntt >= fpc) { ntt -= fpc; ++nclip; noff += wrap_off; }
    const bool has_next = nfa < g.total_frames;
    const bool nxt_fast = has_next && is_fast(nfa, ntt, noff);
    const float2* nsrc = nxt_fast ? reinterpret_cast<const float2*>(g.pcm + noff) + t : idle_src;

    // ---- untangle, in-lane.  The lower halves of a mirror pair's columns lead (k < M/2); their mirrors are the upper
    //      halves: Z[M - (ka + 32 q)] = column kb, element L-1-q, and vice versa.  Lane 0's pair 0 is self-mirrored
    //      (column 0: element L-q; column 16: element L-1-q): its mirrors are first moved to where the general rule
    //      looks (in place, descending q keeps every source intact until it is read).
    static_for<0, L / 2>([&](auto qq) {
      constexpr int q = L / 2 - 1 - decltype(qq)::value;
      const C2 na = a[q ? L - q : 0], nb = a[2 * L - 1 - q];
      a[2 * L - 1 - q].re = P2(c0 ? na.re.v.x : a[2 * L - 1 - q].re.v.x, c0 ? na.re.v.y : a[2 * L - 1 - q].re.v.y);
      a[2 * L - 1 - q].im = P2(c0 ? na.im.v.x : a[2 * L - 1 - q].im.v.x, c0 ? na.im.v.y : a[2 * L - 1 - q].im.v.y);
      a[L - 1 - q].re = P2(c0 ? nb.re.v.x : a[L - 1 - q].re.v.x, c0 ? nb.re.v.y : a[L - 1 - q].re.v.y);
      a[L - 1 - q].im = P2(c0 ? nb.im.v.x : a[L - 1 - q].im.v.x, c0 ? nb.im.v.y : a[L - 1 - q].im.v.y);
    });
    P2 pk[16], pm[16];   // slot i = (r * L/2 + q) * 2 + {0: column ka, 1: column kb}
    static_for<0, NP>([&](auto rr) {
      constexpr int r = decltype(rr)::value;
      constexpr int oa = (2 * r) * L, ob = (2 * r + 1) * L;
      const int ka = col_a(r), kb = col_b(r);
      static_for<0, L / 2>([&](auto qq) {
        constexpr int q = decltype(qq)::value;
        constexpr int slot = (r * (L / 2) + q) * 2;
        auto pair = [&](const C2& zk, const C2& zm, int k, P2& opk, P2& opm) {
          const float2 w = s_ut[k];
          const P2 ex = add2(zk.re, zm.re), ey = add2(zk.im, neg(zm.im));      // 2E
          const P2 ox = add2(zk.im, zm.im), oy = add2(zm.re, neg(zk.re));      // 2O
          const P2 xr = fma2(ox, bc(w.x), fma2(oy, bc(-w.y), ex));             // 2X[k]
          const P2 xi = fma2(ox, bc(w.y), fma2(oy, bc(w.x), ey));
          const P2 yr = fma2(ex, bc(2.f), neg(xr));                            // 2 conj X[M-k]
          const P2 yi = fma2(ey, bc(2.f), neg(xi));
          opk = fma2(xr, xr, mul2(xi, xi));
          opm = fma2(yr, yr, mul2(yi, yi));
        };
        pair(a[oa + q], a[ob + L - 1 - q], ka + 32 * q, pk[slot], pm[slot]);
        pair(a[ob + q], a[oa + L - 1 - q], kb + 32 * q, pk[slot + 1], pm[slot + 1]);
        if constexpr (slot == 0) pm[0] = P2(c0 ? pmid.v.x : pm[0].v.x, c0 ? pmid.v.y : pm[0].v.y);
        // this step's share of the next pair's loads
        constexpr int stepi = r * (L / 2) + q;
        static_for<(NLOAD * stepi) / 8, (NLOAD * (stepi + 1)) / 8>([&](auto mm) {
          constexpr int m = decltype(mm)::value;
          s[m] = ldg_nc_f2(nsrc + L * m);
        });
      });
    });

    // ---- epilogue
    const bool has_b_out = fa + 1 < g.total_frames;
    T* __restrict__ row_a = out + fa * (long long)M;
    T* __restrict__ row_b = row_a + M;
    auto bins_of = [&](auto ii, int& k, int& mk) {
      constexpr int i = decltype(ii)::value, st = i >> 1, r = st / (L / 2), q = st - r * (L / 2);
      k = cols[2 * r + (i & 1)] + 32 * q;
      mk = M - k;
      if constexpr (i == 0) mk = c0 ? M / 2 : mk;
    };
    if constexpr (OUT == kOutU8 || OUT == kOutRgba8) {
      const P2 scale = add2(bc(ep.byte_a), poison);
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        int k, mk;
        bins_of(ii, k, mk);
        const P2 vk = fma2(P2(lg2_ftz(pk[i].v.x), lg2_ftz(pk[i].v.y)), scale, bc(ep.byte_b));
        const P2 vm = fma2(P2(lg2_ftz(pm[i].v.x), lg2_ftz(pm[i].v.y)), scale, bc(ep.byte_b));
        const unsigned kA = byte_of_scaled(vk.v.x), kB = byte_of_scaled(vk.v.y);
        const unsigned mA = byte_of_scaled(vm.v.x), mB = byte_of_scaled(vm.v.y);
        if constexpr (OUT == kOutU8) {
          sb16[k] = (uint16_t)(kA | (kB << 8));
          sb16[mk] = (uint16_t)(mA | (mB << 8));
        } else if (alive) {
          row_a[k] = __ldg(ep.lut + kA); row_a[mk] = __ldg(ep.lut + mA);
          if (has_b_out) { row_b[k] = __ldg(ep.lut + kB); row_b[mk] = __ldg(ep.lut + mB); }
        }
      });
      if constexpr (OUT == kOutU8) {
        __syncwarp();
        // de-interleave the (A,B) byte pairs of this pair: 8 bins per lane per round, 8-byte row stores
        const uint4* s16 = reinterpret_cast<const uint4*>(sb16);
        uint2* ra = reinterpret_cast<uint2*>(row_a);
        uint2* rb = reinterpret_cast<uint2*>(row_b);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 w = s16[c * L + t];
          if (alive) ra[c * L + t] = make_uint2(__byte_perm(w.x, w.y, 0x6420), __byte_perm(w.z, w.w, 0x6420));
          if (has_b_out) rb[c * L + t] = make_uint2(__byte_perm(w.x, w.y, 0x7531), __byte_perm(w.z, w.w, 0x7531));
        }
      }
    } else {
      // float rows (dB / magnitude), packed arithmetic, the non-finite rule from the per-frame flag (a poisoned frame
      // reads 0 / -inf).  A lane's bins are scattered over the row (mirror-pair columns), so the two rows are staged in
      // the pair's idle planes and leave as 16-byte coalesced stores: direct 4-byte stores fill a quarter of each
      // 32-byte sector per instruction.
      float* sfa = reinterpret_cast<float*>(wbase);
      float* sfb = sfa + M;
      const bool bad_a = !(poison.v.x == 0.f), bad_b = !(poison.v.y == 0.f);
      const float z = float_of_poisoned<OUT>();
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        int k, mk;
        bins_of(ii, k, mk);
        const P2 vk = float_of_power<OUT>(pk[i], ep), vm = float_of_power<OUT>(pm[i], ep);
        sfa[k] = bad_a ? z : vk.v.x; sfa[mk] = bad_a ? z : vm.v.x;
        sfb[k] = bad_b ? z : vk.v.y; sfb[mk] = bad_b ? z : vm.v.y;
      });
      __syncwarp();
      const uint4* s4 = reinterpret_cast<const uint4*>(sfa);
      uint4* ra = reinterpret_cast<uint4*>(row_a);
      uint4* rb = reinterpret_cast<uint4*>(row_b);
#pragma unroll
      for (int c = 0; c < 8; ++c) {          // M / 4 = 8 L 16-byte words per row
        if (alive) ra[c * L + t] = s4[c * L + t];
        if (has_b_out) rb[c * L + t] = s4[M / 4 + c * L + t];
      }
    }
    __syncwarp();
    if (!__any_sync(0xffffffffu, has_next)) break;   // the lane groups leave together (the exchange barrier is per warp)
    fa = nfa; off = noff; clip = nclip; tt = ntt;
    cur_fast = nxt_fast;
  }
}

}  // namespace sg
