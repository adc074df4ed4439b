// n_fft = 1024 / 512 / 256 with smoothingTimeConstant > 0 in ONE pass: the part-warp frame-pair kernel of
// kernel_pair.cuh with the AnalyserNode recurrence  X^_t[k] = tau X^_{t-1}[k] + (1 - tau) |X_t[k]|
// (3D/visualizer.js:351,357,362 set tau per mode; [SPEC] step 4) fused between the untangle and the dB / byte epilogue,
// organised like the n_fft 2048 kernel of kernel_w32x2s.cuh (chain mode): a CTA owns a SEGMENT of consecutive frames of
// one clip, its warps take the segment's STEPS round robin -- a step is the PW = 32/L frame pairs a warp holds at once,
// 2 PW consecutive frames -- and run window -> FFT -> untangle -> sqrt concurrently; only the state update is passed from
// warp to warp in step order through an mbarrier chain.  X^ of the segment lives in shared memory ([16][L] float2 in the
// lanes' slot order); segments of a clip are chained through a carry vector and a flag in global memory (tasks are dealt
// segment-major, the grid is co-resident: cooperative launch).
//
// Inside a step the 2 PW frames sit in PW lane groups, two per group (the halves of the packed registers), and the
// recurrence over them is a composition of affine maps  s -> d s + g  (d = tau^2, g = tau A + B for a full pair).
// Everything that does not depend on the incoming state is done BEFORE the warp waits for its turn: the exclusive scan
// of (d, g) over the lane groups (log2 PW + 1 shuffles per value).  Inside the turn a lane reads X^, forms its group's
// incoming state with one FMA, runs its two frames and -- last group only -- writes X^ back: 16 x (LDS.64, 6 FMA, STS.64),
// as in the 2048 kernel.  The scan rounds differently from the frame-by-frame recurrence (a few ulp; the tolerance of
// tests/_tol.py, like the two-kernel path's chunk sums).
// [SPEC] "non-finite X^ -> 0" is applied per frame: if any power of the step is Inf/NaN (one vote per step) the groups
// take their turn one after the other through the shared state with the per-value rule.
#pragma once
#include "kernel_pair.cuh"
#include "kernel_w32x2s.cuh"   // XsGeom, acquire/release, mbarrier helpers

namespace sg {

constexpr int kPsWarps = 8;   // 8 warps x 255 registers (pk/pm, the scanned g and the prefetch do not fit 168)

template <int LOG2L>
struct PsShape {
  using P = PairShape<LOG2L>;
  static constexpr int kStateBytes = 16 * P::L * 8;
  static constexpr int kSmemBytes = P::kTableBytes + kStateBytes + kPsWarps * 8 + kPsWarps * P::kWarpBytes;
};

struct PsItem {
  int it, clip, seg, f0, nfr, skip;   // skip: leading frames without output (XsGeom mode 2: warm-up of an independent segment)
  bool valid;
};
// WARM instantiations serve XsGeom mode 2 (independent segments with a warm-up from a zero state, kernel_w32x2s.cuh).  They
// are separate instantiations because the chained kernel cannot afford the mode as a run-time switch -- three more integer
// operations per item and two more store predicates cost it 9 % (785 vs 860 M frames/s at n_fft 1024: at 254 registers the
// loop state is rematerialised from the item every iteration) -- and they keep their `x.mode == 2` tests as run-time tests:
// the same code with those tests folded at compile time ran its turn ring at half speed (404 vs 697 M frames/s; 35 % of the
// samples in the try_wait spin, `no_instruction` 2.3 per issue), for a reason the profile shows but does not explain.
template <bool WARM>
__device__ __forceinline__ PsItem ps_item(const XsGeom& x, int fpc, int it) {
  PsItem c;
  c.it = it;
  const unsigned n_clips = (unsigned)x.n_clips, n_tasks = (unsigned)x.segs * n_clips;   // < 2^31 (host checks)
  const unsigned task = blockIdx.x + (unsigned)it * gridDim.x;
  c.valid = task < n_tasks;
  c.seg = (int)(task / n_clips);
  c.clip = (int)(task - (unsigned)c.seg * n_clips);
  c.f0 = c.seg * x.seg_frames;
  c.nfr = min(x.seg_frames, fpc - c.f0);
  c.skip = 0;
  if (WARM && x.mode == 2) {
    c.skip = min(c.f0, x.warm);
    c.f0 -= c.skip;
    c.nfr += c.skip;
  }
  return c;
}

template <int OUT, int LOG2L, int HOPJ, bool WARM = false>
__global__ void __launch_bounds__(kPsWarps * 32, 1)
stft_pair_s_kernel(FrameGeom g, XsGeom x, PairPlan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using T = typename OutElem<OUT>::type;
  using S = PairShape<LOG2L>;
  constexpr int L = S::L, N = S::N, M = S::M, NCOL = S::NCOL, NP = S::NP, PW = S::PW, NW = kPsWarps;
  // HOPJ > 0: hop = 2 L HOPJ, frame B shares frame A's loads, the next step is prefetched into registers.
  // HOPJ == 0: any hop -- every lane group loads its two frames directly at the top of the iteration (8-byte loads where
  // the frame starts allow, else 4-byte), no prefetch.
  constexpr int NLOAD = HOPJ ? 32 + HOPJ : 1;
  const int HOP = HOPJ ? 2 * L * HOPJ : (int)g.hop;
  extern __shared__ float4 smem_raw[];
  float4* s_win4 = smem_raw;                                           // [16][L] (w2[t+Lj], w2[t+L(j+16)])
  float2* s_twb = reinterpret_cast<float2*>(s_win4 + 16 * L);          // [LOG2L][32]  W_{32*2^u}^col
  float2* s_ut = s_twb + LOG2L * 32;                                   // [M/2 + 2]    W_N^k
  float2* s_state = s_ut + S::kUtEntries;                              // [16][L] (X^[k], X^[mirror k]) in slot order
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_state + 16 * L);     // [NW] the turn of warp w
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> LOG2L, t = lane & (L - 1);
  unsigned char* wbase = reinterpret_cast<unsigned char*>(s_bar + NW) + warp * S::kWarpBytes + h * S::kPairBytes;
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
    if (threadIdx.x < NW) mbar_init(s_bar + threadIdx.x, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(s_bar);      // warp 0 holds the first turn
  unsigned turn = 0;                             // phase parity of this warp's next wait

  const int fpc = (int)g.frames_per_clip;
  int t_lo, t_hi;                                // pairs with t in [t_lo, t_hi] lie wholly inside their clip
  {
    const long long lo = g.start0 >= 0 ? 0 : (-g.start0 + HOP - 1) / HOP;
    const long long room = g.clip_len - (HOP + N) - g.start0;
    const long long hi = room < 0 ? -1 : min((long long)fpc - 2, room / HOP);
    t_lo = (int)lo;
    t_hi = (int)hi;
  }
  const unsigned pcm_lo = (unsigned)reinterpret_cast<uintptr_t>(g.pcm);
  auto steps_of = [&](const PsItem& c) { return ((c.nfr + 1) / 2 + PW - 1) / PW; };
  // this lane group's pair in step u of item c: first frame (clip-relative), whether it exists, whether it has a frame B
  auto pair_off = [&](const PsItem& c, int ta) { return c.clip * g.clip_stride + g.start0 + (long long)ta * HOP; };
  auto pair_fast = [&](const PsItem& c, int u, long long off) {
    const int p = u * PW + h, ta = c.f0 + 2 * p;
    return HOPJ != 0 && c.valid && 2 * p + 1 < c.nfr && ta >= t_lo && ta <= t_hi && ((pcm_lo + ((unsigned)off << 2)) & 7u) == 0;
  };
  auto advance = [&](PsItem& c, int& u) {
    u += NW;
    while (c.valid && u >= steps_of(c)) {
      u -= steps_of(c);
      c = ps_item<WARM>(x, fpc, c.it + 1);
    }
  };

  int it, u = warp - NW;
  bool cur_fast;
  {
    PsItem c0 = ps_item<WARM>(x, fpc, 0);
    advance(c0, u);
    if (!c0.valid) return;
    it = c0.it;
    cur_fast = pair_fast(c0, u, pair_off(c0, c0.f0 + 2 * (u * PW + h)));
  }
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
  // natural bins of slot i: k and its mirror (lane 0's slot 0: bin M/2 in place of the dropped Nyquist bin)
  auto bins_of = [&](auto ii, int& k, int& mk) {
    constexpr int i = decltype(ii)::value, st = i >> 1, r = st / (L / 2), q = st - r * (L / 2);
    k = cols[2 * r + (i & 1)] + 32 * q;
    mk = M - k;
    if constexpr (i == 0) mk = c0 ? M / 2 : mk;
  };
  float2 s[NLOAD];
  const float2* idle_src = reinterpret_cast<const float2*>(pl.win) + t;   // readable past the window (build_plan)
  {
    const PsItem ci = ps_item<WARM>(x, fpc, it);
    const float2* src = cur_fast ? reinterpret_cast<const float2*>(g.pcm + pair_off(ci, ci.f0 + 2 * (u * PW + h))) + t : idle_src;
    if constexpr (HOPJ != 0)
      static_for<0, NLOAD>([&](auto mm) { constexpr int m = decltype(mm)::value; s[m] = ldg_nc_f2(src + L * m); });
  }

#ifdef SG_DEBUG
  unsigned dbg_q = warp;                        // this step's place in the CTA's step sequence (round robin over warps)
  unsigned* dbg_tags = dbg_cta_tags();          // one tag per state slot (i, t): the step that wrote it, plus one
#endif
  while (true) {
    const PsItem cur = ps_item<WARM>(x, fpc, it);
    const int p = u * PW + h;
    const bool active = 2 * p < cur.nfr, has_b = 2 * p + 1 < cur.nfr;
    const bool store_a = active && (!WARM || 2 * p >= cur.skip), store_b = has_b && (!WARM || 2 * p >= cur.skip);   // (skip is even)
    const int ta = cur.f0 + min(2 * p, cur.nfr - 1);          // idle lane groups recompute the segment's last frame
    // ---- steps 1-2 (+ FFT stage 1)
    C2 a[32];
    auto load_guarded = [&] {
      // clip edges / zero history / a segment's odd last frame (frame B reads as frame A and is dropped)
      const float* __restrict__ xa = g.pcm + cur.clip * g.clip_stride;
      const long long start_a = g.start0 + (long long)ta * HOP, start_b = start_a + (has_b ? HOP : 0);
      auto ld = [&](long long q) { return (q >= 0 && q < g.clip_len) ? __ldg(xa + q) : 0.f; };
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
        const long long o0 = 2 * (t + L * j), o1 = 2 * (t + L * (j + 16));
        const float4 w = s_win4[j * L + t];
        window_stage1(a[r0], a[r1], make_float2(ld(start_a + o0), ld(start_a + o0 + 1)),
                      make_float2(ld(start_a + o1), ld(start_a + o1 + 1)),
                      make_float2(ld(start_b + o0), ld(start_b + o0 + 1)),
                      make_float2(ld(start_b + o1), ld(start_b + o1 + 1)), make_float2(w.x, w.y),
                      make_float2(w.z, w.w));
      });
    };
    if constexpr (HOPJ != 0) {
      if (cur_fast) {
        static_for<0, 16>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
          const float4 w = s_win4[j * L + t];
          window_stage1(a[r0], a[r1], s[j], s[j + 16], s[j + HOPJ], s[j + 16 + HOPJ], make_float2(w.x, w.y),
                        make_float2(w.z, w.w));
        });
      } else {
        load_guarded();
      }
    } else {
      // any hop: both frames inside the clip -> unguarded loads, 8 bytes wide when both frame starts are 8-byte aligned
      const float* __restrict__ xa = g.pcm + cur.clip * g.clip_stride;
      const long long start_a = g.start0 + (long long)ta * HOP, start_b = start_a + (has_b ? HOP : 0);
      if (start_a >= 0 && start_b + N <= g.clip_len) {
        const float* __restrict__ fa_ = xa + start_a + 2 * t;
        const float* __restrict__ fb_ = xa + start_b + 2 * t;
        if (((reinterpret_cast<uintptr_t>(fa_) | reinterpret_cast<uintptr_t>(fb_)) & 7) == 0) {
          const float2* __restrict__ pa = reinterpret_cast<const float2*>(fa_);
          const float2* __restrict__ pb = reinterpret_cast<const float2*>(fb_);
          static_for<0, 16>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
            const float4 w = s_win4[j * L + t];
            window_stage1(a[r0], a[r1], __ldg(pa + L * j), __ldg(pa + L * (j + 16)), __ldg(pb + L * j), __ldg(pb + L * (j + 16)),
                          make_float2(w.x, w.y), make_float2(w.z, w.w));
            if constexpr (j == 7) asm volatile("" ::: "memory");   // two batches of 32 loads: 64 at once do not fit the registers
          });
        } else {
          static_for<0, 16>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
            const float4 w = s_win4[j * L + t];
            window_stage1(a[r0], a[r1], make_float2(__ldg(fa_ + 2 * L * j), __ldg(fa_ + 2 * L * j + 1)),
                          make_float2(__ldg(fa_ + 2 * L * (j + 16)), __ldg(fa_ + 2 * L * (j + 16) + 1)),
                          make_float2(__ldg(fb_ + 2 * L * j), __ldg(fb_ + 2 * L * j + 1)),
                          make_float2(__ldg(fb_ + 2 * L * (j + 16)), __ldg(fb_ + 2 * L * (j + 16) + 1)), make_float2(w.x, w.y),
                          make_float2(w.z, w.w));
            if constexpr (j == 7) asm volatile("" ::: "memory");
          });
        }
      } else {
        load_guarded();
      }
    }

    // ---- pass 1, exchange, pass 2: kernel_pair.cuh
    dit2_stage_const<2>(a);
    dit2_stage_const<3>(a);
    dit2_stage_const<4>(a);
    dit2_stage_const<5>(a);
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
    static_for<1, LOG2L + 1>([&](auto uu) {
      constexpr int us = decltype(uu)::value;
      static_for<0, NP>([&](auto rr) {
        constexpr int r = decltype(rr)::value;
        dit2_stage_gen_n<us, (2 * r) * L, L>(a, s_twb[(us - 1) * 32 + col_a(r)]);
        dit2_stage_gen_n<us, (2 * r + 1) * L, L>(a, s_twb[(us - 1) * 32 + col_b(r)]);
      });
    });
    // lane 0: bin M/2 = conj Z[M/2] is element L/2 of column 0
    const P2 pmid = mul2(bc(4.f), fma2(a[L / 2].re, a[L / 2].re, mul2(a[L / 2].im, a[L / 2].im)));

    // ---- next step of this warp
    int nit, nu = u;
    bool has_next, nxt_fast;
    const float2* nsrc;
    {
      PsItem nxt = cur;
      advance(nxt, nu);
      nit = nxt.it;
      has_next = nxt.valid;
      const long long nxt_off = pair_off(nxt, nxt.f0 + 2 * (nu * PW + h));
      nxt_fast = has_next && pair_fast(nxt, nu, nxt_off);
      nsrc = nxt_fast ? reinterpret_cast<const float2*>(g.pcm + nxt_off) + t : idle_src;
    }

    // ---- untangle, in-lane (kernel_pair.cuh), then (1 - tau) |X| / N of both frames
    static_for<0, L / 2>([&](auto qq) {
      constexpr int q = L / 2 - 1 - decltype(qq)::value;
      const C2 na = a[q ? L - q : 0], nb = a[2 * L - 1 - q];
      a[2 * L - 1 - q].re = P2(c0 ? na.re.v.x : a[2 * L - 1 - q].re.v.x, c0 ? na.re.v.y : a[2 * L - 1 - q].re.v.y);
      a[2 * L - 1 - q].im = P2(c0 ? na.im.v.x : a[2 * L - 1 - q].im.v.x, c0 ? na.im.v.y : a[2 * L - 1 - q].im.v.y);
      a[L - 1 - q].re = P2(c0 ? nb.re.v.x : a[L - 1 - q].re.v.x, c0 ? nb.re.v.y : a[L - 1 - q].re.v.y);
      a[L - 1 - q].im = P2(c0 ? nb.im.v.x : a[L - 1 - q].im.v.x, c0 ? nb.im.v.y : a[L - 1 - q].im.v.y);
    });
    P2 pk[16], pm[16];   // slot i = (r * L/2 + q) * 2 + {0: column ka, 1: column kb}
    unsigned worst = 0;  // largest bit pattern among this lane's powers: >= 0x7f800000 means Inf or NaN
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
        static_for<slot, slot + 2>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          worst = max(max(worst, max(__float_as_uint(pk[i].v.x), __float_as_uint(pk[i].v.y))),
                      max(__float_as_uint(pm[i].v.x), __float_as_uint(pm[i].v.y)));
          pk[i] = mul2(P2(sqrt_ftz(pk[i].v.x), sqrt_ftz(pk[i].v.y)), bc(x.mscale));
          pm[i] = mul2(P2(sqrt_ftz(pm[i].v.x), sqrt_ftz(pm[i].v.y)), bc(x.mscale));
        });
      });
    });
    const bool dirty = __any_sync(0xffffffffu, worst >= 0x7f800000u);

    // ---- before the turn: the exclusive scan over the lane groups of the affine maps of their pairs.
    //      Group h maps s -> d s + g with (d, g) = (tau^2, tau A + B) for a full pair, (tau, A) for a lone frame A,
    //      (1, 0) for an idle group; ek/em become the g of everything before this group, dex its d.
    float ek[16], em[16], dex = 1.f;
    if (!dirty) {
      float d = active ? (has_b ? x.tau * x.tau : x.tau) : 1.f;
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        ek[i] = active ? (has_b ? fmaf(x.tau, pk[i].v.x, pk[i].v.y) : pk[i].v.x) : 0.f;
        em[i] = active ? (has_b ? fmaf(x.tau, pm[i].v.x, pm[i].v.y) : pm[i].v.x) : 0.f;
      });
      // inclusive scan, reaching back 2^s groups (two groups: the exclusive values below are the raw maps already)
      static_for<0, (PW > 2 ? 5 - LOG2L : 0)>([&](auto ss) {
        constexpr int sx = decltype(ss)::value, o = 1 << sx;
        const bool take = h >= o;
        const float dm = take ? d : 0.f;            // groups without a partner keep their map
        static_for<0, 16>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          ek[i] = fmaf(dm, __shfl_up_sync(0xffffffffu, ek[i], o * L), ek[i]);
          em[i] = fmaf(dm, __shfl_up_sync(0xffffffffu, em[i], o * L), em[i]);
        });
        const float dp = __shfl_up_sync(0xffffffffu, d, o * L);
        d = take ? d * dp : d;
      });
      // exclusive: what the groups before this one make of the incoming state
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        const float vk = __shfl_up_sync(0xffffffffu, ek[i], L), vm = __shfl_up_sync(0xffffffffu, em[i], L);
        ek[i] = h ? vk : 0.f;
        em[i] = h ? vm : 0.f;
      });
      const float dp = __shfl_up_sync(0xffffffffu, d, L);
      dex = h ? dp : 1.f;
    }

    // ---- the recurrence, in step order
    if (u == 0 && cur.seg > 0 && !(WARM && x.mode == 2)) {
      // the segment this one starts from must have been published
      if (lane == 0)
        while (ld_acquire_u32(x.flags + (long long)(cur.seg - 1) * x.n_clips + cur.clip) != x.epoch) {}
      __syncwarp();
    }
#ifdef SG_DEBUG
    if (!g_sg_dbg.break_chain)
#endif
    while (!mbar_try_wait(s_bar + warp, turn)) {}
#ifdef SG_DEBUG
    if (lane == 0) dbg_count_iteration();
    // every slot must hold the state the step just before this one left
    if (dbg_q != 0) static_for<0, 16>([&](auto ii) { constexpr int i = decltype(ii)::value; dbg_check(dbg_tags, i * L + t, dbg_q, 3); });
#endif
    if (u == 0) {
      // first step of a work item: the state the segment starts from
      if (h == 0) {
        if (cur.seg == 0 || (WARM && x.mode == 2)) {
          // (an independent segment starts from zero unless its warm-up reaches back to the clip's first frame)
          const float* __restrict__ si = (x.state_in && cur.f0 == 0) ? x.state_in + (long long)cur.clip * M : nullptr;
          static_for<0, 16>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            int k, mk;
            bins_of(ii, k, mk);
            s_state[i * L + t] = si ? make_float2(si[k], si[mk & (M - 1)]) : make_float2(0.f, 0.f);
          });
        } else {
          const float2* __restrict__ cv = x.carry + ((long long)(cur.seg - 1) * x.n_clips + cur.clip) * (16 * L) + t;
          static_for<0, 16>([&](auto ii) { constexpr int i = decltype(ii)::value; s_state[i * L + t] = __ldcg(cv + i * L); });
        }
      }
      __syncwarp();
    }
    float2 fin[16];      // the state this step leaves behind (meaningful in the last lane group)
    if (!dirty) {
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        const float2 st = s_state[i * L + t];
        const float sk = fmaf(dex, st.x, ek[i]), sm = fmaf(dex, st.y, em[i]);     // this group's incoming state
        const float ka = fmaf(x.tau, sk, pk[i].v.x), kb = fmaf(x.tau, ka, pk[i].v.y);
        const float ma = fmaf(x.tau, sm, pm[i].v.x), mb = fmaf(x.tau, ma, pm[i].v.y);
        pk[i] = P2(ka, kb);
        pm[i] = P2(ma, mb);
        fin[i] = active ? (has_b ? make_float2(kb, mb) : make_float2(ka, ma)) : make_float2(sk, sm);
        if (h == PW - 1) s_state[i * L + t] = fin[i];
      });
    } else {
      // [SPEC] a non-finite X^ is set to 0: the groups go one after the other through the shared state
#pragma unroll 1
      for (int hh = 0; hh < PW; ++hh) {
        if (h == hh) {
          static_for<0, 16>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const float2 st = s_state[i * L + t];
            const float ka = finite_or_zero(fmaf(x.tau, st.x, pk[i].v.x)), kb = finite_or_zero(fmaf(x.tau, ka, pk[i].v.y));
            const float ma = finite_or_zero(fmaf(x.tau, st.y, pm[i].v.x)), mb = finite_or_zero(fmaf(x.tau, ma, pm[i].v.y));
            pk[i] = P2(ka, kb);
            pm[i] = P2(ma, mb);
            fin[i] = active ? (has_b ? make_float2(kb, mb) : make_float2(ka, ma)) : st;
            s_state[i * L + t] = fin[i];
          });
        }
        __syncwarp();
      }
    }
#ifdef SG_DEBUG
    if (h == PW - 1) static_for<0, 16>([&](auto ii) { constexpr int i = decltype(ii)::value; dbg_write(dbg_tags, i * L + t, dbg_q + 1); });
    __threadfence_block();
    dbg_q += NW;
#endif
    __syncwarp();
    if (lane == 0) mbar_arrive(s_bar + (warp + 1 == NW ? 0 : warp + 1));
    turn ^= 1;

    if (u == steps_of(cur) - 1 && h == PW - 1) {
      // last step of a work item: hand the state to the next segment (or to the caller)
      if (cur.seg + 1 < x.segs) {
        if (!(WARM && x.mode == 2)) {
        const long long me = (long long)cur.seg * x.n_clips + cur.clip;
        float2* __restrict__ cv = x.carry + me * (16 * L) + t;
        static_for<0, 16>([&](auto ii) { constexpr int i = decltype(ii)::value; cv[i * L] = fin[i]; });
        __threadfence();
        __syncwarp((0xffffffffu >> (32 - L)) << (L * (PW - 1)));
        if (t == 0) st_release_u32(x.flags + me, x.epoch);
        }
      } else if (x.state_out != nullptr) {
        float* __restrict__ so = x.state_out + (long long)cur.clip * M;
        static_for<0, 16>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          int k, mk;
          bins_of(ii, k, mk);
          so[k] = fin[i].x;
          so[mk & (M - 1)] = fin[i].y;
        });
      }
    }
    // the next step's loads, after the turn has been passed on
    if constexpr (HOPJ != 0)
      static_for<0, NLOAD>([&](auto mm) { constexpr int m = decltype(mm)::value; s[m] = ldg_nc_f2(nsrc + L * m); });

    // ---- epilogue: X^ -> dB / byte / colour of both frames
    T* __restrict__ row_a = out + ((long long)cur.clip * x.out_clip_rows + ta) * (long long)M;
    T* __restrict__ row_b = row_a + M;
    if constexpr (OUT == kOutU8 || OUT == kOutRgba8) {
      const P2 scale = bc(2.f * ep.byte_a);
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        int k, mk;
        bins_of(ii, k, mk);
        const P2 vk = fma2(P2(lg2_ftz(pk[i].v.x), lg2_ftz(pk[i].v.y)), scale, bc(ep.byte_b0));
        const P2 vm = fma2(P2(lg2_ftz(pm[i].v.x), lg2_ftz(pm[i].v.y)), scale, bc(ep.byte_b0));
        const unsigned kA = byte_of_scaled(vk.v.x), kB = byte_of_scaled(vk.v.y);
        const unsigned mA = byte_of_scaled(vm.v.x), mB = byte_of_scaled(vm.v.y);
        if constexpr (OUT == kOutU8) {
          sb16[k] = (uint16_t)__byte_perm(kA, kB, 0x0040);
          sb16[mk] = (uint16_t)__byte_perm(mA, mB, 0x0040);
        } else if (store_a) {
          row_a[k] = __ldg(ep.lut + kA); row_a[mk] = __ldg(ep.lut + mA);
          if (store_b) { row_b[k] = __ldg(ep.lut + kB); row_b[mk] = __ldg(ep.lut + mB); }
        }
      });
      if constexpr (OUT == kOutU8) {
        __syncwarp();
        const uint4* s16 = reinterpret_cast<const uint4*>(sb16);
        uint2* ra = reinterpret_cast<uint2*>(row_a);
        uint2* rb = reinterpret_cast<uint2*>(row_b);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 w = s16[c * L + t];
          if (store_a) ra[c * L + t] = make_uint2(__byte_perm(w.x, w.y, 0x6420), __byte_perm(w.z, w.w, 0x6420));
          if (store_b) rb[c * L + t] = make_uint2(__byte_perm(w.x, w.y, 0x7531), __byte_perm(w.z, w.w, 0x7531));
        }
      }
    } else {
      // float rows: staged in the pair's idle planes, stored as 16-byte coalesced rows (kernel_pair.cuh)
      float* sfa = reinterpret_cast<float*>(wbase);
      float* sfb = sfa + M;
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        int k, mk;
        bins_of(ii, k, mk);
        P2 vk = pk[i], vm = pm[i];
        if constexpr (OUT == kOutF32Db) {
          vk = mul2(P2(lg2_ftz(vk.v.x), lg2_ftz(vk.v.y)), bc(2.f * ep.db_scale));
          vm = mul2(P2(lg2_ftz(vm.v.x), lg2_ftz(vm.v.y)), bc(2.f * ep.db_scale));
        }
        sfa[k] = vk.v.x; sfa[mk] = vm.v.x;
        sfb[k] = vk.v.y; sfb[mk] = vm.v.y;
      });
      __syncwarp();
      const uint4* s4 = reinterpret_cast<const uint4*>(sfa);
      uint4* ra = reinterpret_cast<uint4*>(row_a);
      uint4* rb = reinterpret_cast<uint4*>(row_b);
#pragma unroll
      for (int c = 0; c < 8; ++c) {          // M / 4 = 8 L 16-byte words per row
        if (store_a) ra[c * L + t] = s4[c * L + t];
        if (store_b) rb[c * L + t] = s4[M / 4 + c * L + t];
      }
    }
    __syncwarp();
    if (!has_next) break;
    it = nit;
    u = nu;
    cur_fast = nxt_fast;
  }
}

}  // namespace sg
