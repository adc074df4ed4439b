// n_fft = 2048 with smoothingTimeConstant > 0 in ONE pass: the frame-pair kernel of kernel_w32x2p.cuh with the
// AnalyserNode recurrence  X^_t[k] = tau X^_{t-1}[k] + (1 - tau) |X_t[k]|  (3D/visualizer.js:351,357,362 set tau per
// mode; [SPEC] step 4) fused between the untangle and the dB / byte epilogue.  No magnitude scratch in HBM.
//
// The FFTs of a clip's frames are independent; only the recurrence is ordered.  So a CTA owns a SEGMENT of consecutive
// frames of one clip, its warps take the segment's frame pairs round robin and run window -> FFT -> untangle -> sqrt
// concurrently, and only the state update -- 16 x (LDS.64, 4 FMA, STS.64) per lane -- is passed from warp to warp in
// pair order through an mbarrier chain (the waiting warp sleeps in try_wait, it takes no issue slots).  X^ of the
// segment lives in 4 KB of shared memory in lane order (slot (i, lane) = bins lane + 32 i and its mirror).
// Segments of one clip are chained through a carry vector in global memory:
//   mode 0 (chain)  many clips: tasks (segment, clip) are dealt to CTAs segment-major, so a task's predecessor was
//                   started ~n_clips / gridDim tasks earlier; its last pair publishes the state and a flag, the
//                   successor's first pair waits for it.  Exactly the sequential arithmetic, bit for bit.
//   mode 1 (few)    fewer clips than CTAs: one task per CTA, run twice -- first without output from a zero state,
//                   which yields the segment's aggregate (the recurrence is linear), then, after a look-back over the
//                   aggregates of the segments before it (all produced concurrently), again from the true state with
//                   output.  Twice the arithmetic, on a GPU that would otherwise idle; one launch.
// [SPEC] "non-finite X^ -> 0" is applied per frame: one integer max over the lane's powers per pair decides whether
// the slow, per-value path is needed.  In the aggregate pass the sign bit of a state value records that a non-finite
// frame wiped that bin inside the segment (X^ itself is never negative), so the look-back drops what came in from
// earlier segments.
// A warp-specialised form (10 producer warps leaving |2X|^2 in their planes, 2 consumer warps doing the recurrence and
// the epilogue in pair order, state in registers) was built and measured at 218-244 M frames/s against 470 M for this
// one: a single ordered warp only ever gets its fair share of a scheduler's issue slots.
#pragma once
#include "kernel_w32x2p.cuh"

namespace sg {

struct XsGeom {
  long long n_clips;
  long long out_clip_rows;  // output rows between consecutive clips (>= frames_per_clip: a frame range of longer clips)
  int seg_frames;          // frames per segment (even); the last segment of a clip may be shorter
  int segs;                // segments per clip
  float tau;
  float mscale;            // (1 - tau) / norm: sqrt(|2X|^2) * mscale = (1 - tau) |X| / N
  float dec;               // tau^seg_frames (mode 1 look-back)
  const float* state_in;   // [n_clips][1024], natural bin order; nullptr = zeros
  float* state_out;        // [n_clips][1024], natural bin order; nullptr = not wanted
  float2* carry;           // [segs * n_clips][16][32]: state at the END of a segment (mode 1: its aggregate)
  unsigned* flags;         // [segs * n_clips]: == epoch once the carry is visible
  unsigned epoch;
  int mode;                // 0 chained segments, 1 aggregate pass + look-back + emit pass, 2 independent segments with warm-up
  int warm;                // mode 2: frames (even, a multiple of the kernel's step) a segment runs ahead of its first output
                           // row, from a zero state, so that nothing of what it missed survives in float32: tau^warm < 2^-149
};

constexpr int kXsStateBytes = 16 * 32 * 8;
template <int NW>
struct XsShape {
  static constexpr int kSmemBytes = XpShape<NW>::kSmemBytes + kXsStateBytes + 4 * NW * 8;   // up to 4 turn chains
};

// the work item a pair belongs to: (segment, clip) and what to do at its first / last pair
struct XsItem {
  int it;        // index in this CTA's item list
  int clip, seg;
  int f0, nfr;   // first frame of the segment within the clip, frames in it (mode 2: including the warm-up frames)
  int skip;      // leading frames that produce no output (mode 2 warm-up)
  int kind;      // 0 chain, 1 aggregate pass (no output, zero state), 2 emit pass after look-back, 3 independent + warm-up
  bool valid;
};
__device__ __forceinline__ XsItem xs_item(const XsGeom& x, int fpc, int it) {
  XsItem c;
  c.it = it;
  const unsigned n_clips = (unsigned)x.n_clips, n_tasks = (unsigned)x.segs * n_clips;   // < 2^31 (host checks)
  const unsigned task = x.mode != 1 ? blockIdx.x + (unsigned)it * gridDim.x : blockIdx.x;
  c.valid = task < n_tasks && (x.mode != 1 || it < 2);
  c.kind = x.mode == 0 ? 0 : x.mode == 1 ? 1 + it : 3;
  c.seg = (int)(task / n_clips);
  c.clip = (int)(task - (unsigned)c.seg * n_clips);
  c.f0 = c.seg * x.seg_frames;
  c.nfr = min(x.seg_frames, fpc - c.f0);
  c.skip = 0;
  if (x.mode == 2) {
    c.skip = min(c.f0, x.warm);
    c.f0 -= c.skip;
    c.nfr += c.skip;
  }
  return c;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// natural bin index of state slot (i, lane, half): half 0 = bin lane + 32 i, half 1 = its mirror (lane 0, i 0: bin 512)
__device__ __forceinline__ int xs_bin(int i, int lane, int half) {
  const int k = lane + 32 * i;
  return half == 0 ? k : ((i == 0 && lane == 0) ? 512 : kW32M - k);
}

// LATE: the next pair's loads are issued after the turn has been passed on (in the epilogue) instead of inside the untangle
template <int OUT, int NW, int HOPJ, bool LATE, int K>
__global__ void __launch_bounds__(NW * 32, 1) __maxnreg__(XpShape<NW>::kMaxRegs)
stft_w32x2s_kernel(FrameGeom g, XsGeom x, W32Plan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using T = typename OutElem<OUT>::type;
  // HOPJ > 0: hop = 64 HOPJ, frame B shares frame A's loads, the next pair is prefetched into registers.
  // HOPJ == 0: any hop (441, 735 ... -- a display or feature cadence turned into a fixed hop): every pair loads its two
  // frames directly at the top of its iteration (8-byte loads where the frame starts allow, else 4-byte), no prefetch.
  constexpr int NLOAD = HOPJ ? 32 + HOPJ : 1;
  const int HOP = HOPJ ? 64 * HOPJ : (int)g.hop;
  extern __shared__ float4 smem_raw[];
  float4* s_win4 = smem_raw;                                           // [16][32] (w2[l+32j], w2[l+32(j+16)])
  float2* s_twb = reinterpret_cast<float2*>(s_win4 + 16 * 32);         // [5][32]  W_{32*2^u}^lane
  float2* s_ut = s_twb + 5 * 32;                                       // [16][32] W_2048^{lane + 32 i}
  float2* s_state = s_ut + 16 * 32;                                    // [16][32] (X^[k], X^[mirror k])
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_state + 16 * 32);    // [K][NW] the turn of warp w on slot group c
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wbase = reinterpret_cast<unsigned char*>(s_bar + 4 * NW) + warp * kXpWarpBytes;
  float4* xp = reinterpret_cast<float4*>(wbase);                       // exchange planes
  uint16_t* sb16 = reinterpret_cast<uint16_t*>(wbase);                 // byte stage: aliases the planes

  {
    const float2* w2 = reinterpret_cast<const float2*>(pl.win);
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) {
      const int j = i >> 5, l = i & 31;
      const float2 lo = __ldg(w2 + l + 32 * j), hi = __ldg(w2 + l + 32 * (j + 16));
      s_win4[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    for (int i = threadIdx.x; i < 5 * 32; i += blockDim.x) {
      const int u = i >> 5, l = i & 31;
      s_twb[i] = __ldg(pl.tw2 + ((1 << u) - 1) * 32 + l);
    }
    for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) s_ut[i] = __ldg(pl.ut + i);
    if (threadIdx.x < K * NW) mbar_init(s_bar + threadIdx.x, 1);
  }
  __syncthreads();
  if (threadIdx.x < K) mbar_arrive(s_bar + threadIdx.x * NW);      // warp 0 holds the first turn of every group
  unsigned turn = 0;                             // phase parity of this warp's next wait

  const int fpc = (int)g.frames_per_clip;
  // pairs with t in [t_lo, t_hi] lie wholly inside their clip (no zero fill)
  int t_lo, t_hi;
  {
    const long long lo = g.start0 >= 0 ? 0 : (-g.start0 + HOP - 1) / HOP;
    const long long room = g.clip_len - (HOP + kW32N) - g.start0;                 // start0 + HOP t <= room
    const long long hi = room < 0 ? -1 : min((long long)fpc - 2, room / HOP);
    t_lo = (int)lo;
    t_hi = (int)hi;
  }
  const unsigned pcm_lo = (unsigned)reinterpret_cast<uintptr_t>(g.pcm);
  auto pair_off = [&](const XsItem& c, int p) { return c.clip * g.clip_stride + g.start0 + (long long)(c.f0 + 2 * p) * HOP; };
  auto pair_fast = [&](const XsItem& c, int p, long long off) {
    const int t = c.f0 + 2 * p;
    return HOPJ != 0 && c.valid && 2 * p + 1 < c.nfr && t >= t_lo && t <= t_hi && ((pcm_lo + ((unsigned)off << 2)) & 7u) == 0;
  };
  // the pair NW places further down this CTA's pair sequence.  Only (item index, pair index) is carried from one
  // iteration to the next; the item's fields are re-derived where they are needed (registers are what this kernel
  // has least of).
  auto advance = [&](XsItem& c, int& p) {
    p += NW;
    while (c.valid && p >= (c.nfr + 1) / 2) {
      p -= (c.nfr + 1) / 2;
      c = xs_item(x, fpc, c.it + 1);
    }
  };

  int it, p = warp - NW;
  bool cur_fast;
  {
    XsItem c0 = xs_item(x, fpc, 0);
    advance(c0, p);
    if (!c0.valid) return;
    it = c0.it;
    cur_fast = pair_fast(c0, p, pair_off(c0, p));
  }
  const int partner = (32 - lane) & 31;
  const bool lane0 = lane == 0;

  float2 s[NLOAD];
  const float2* idle_src = reinterpret_cast<const float2*>(pl.win) + lane;   // 4096 readable floats (build_plan)
  {
    const XsItem c0 = xs_item(x, fpc, it);
    const float2* src = cur_fast ? reinterpret_cast<const float2*>(g.pcm + pair_off(c0, p)) + lane : idle_src;
    if constexpr (HOPJ != 0)
      static_for<0, NLOAD>([&](auto mm) { constexpr int m = decltype(mm)::value; s[m] = ldg_nc_f2(src + 32 * m); });
  }

#ifdef SG_DEBUG
  unsigned dbg_q = warp;                        // this pair's place in the CTA's pair sequence (round robin over warps)
  unsigned* dbg_tags = dbg_cta_tags();          // one tag per state slot (i, lane): the pair that wrote it, plus one
#endif
  while (true) {
    const XsItem cur = xs_item(x, fpc, it);
    const int ta = cur.f0 + 2 * p;
    const bool has_b = 2 * p + 1 < cur.nfr;
    // ---- steps 1-2 (+ FFT stage 1): window both frames, bit-reversed into registers
    C2 a[32];
    auto load_guarded = [&] {
      // clip edges / zero history / a segment's odd last frame (frame B reads as frame A and is dropped)
      const float* __restrict__ xa = g.pcm + cur.clip * g.clip_stride;
      const long long start_a = g.start0 + (long long)ta * HOP, start_b = start_a + (has_b ? HOP : 0);
      auto ld = [&](long long q) { return (q >= 0 && q < g.clip_len) ? __ldg(xa + q) : 0.f; };
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
        const long long o0 = 2 * (lane + 32 * j), o1 = 2 * (lane + 32 * (j + 16));
        const float4 w = s_win4[j * 32 + lane];
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
          constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);   // r1 == r0 + 1
          const float4 w = s_win4[j * 32 + lane];
          window_stage1(a[r0], a[r1], s[j], s[j + 16], s[j + (HOPJ ? HOPJ : 0)], s[j + 16 + (HOPJ ? HOPJ : 0)],
                        make_float2(w.x, w.y), make_float2(w.z, w.w));
        });
      } else {
        load_guarded();
      }
    } else {
      // any hop: both frames inside the clip -> unguarded loads, 8 bytes wide when both frame starts are 8-byte aligned
      const float* __restrict__ xa = g.pcm + cur.clip * g.clip_stride;
      const long long start_a = g.start0 + (long long)ta * HOP, start_b = start_a + (has_b ? HOP : 0);
      if (start_a >= 0 && start_b + kW32N <= g.clip_len) {
        const float* __restrict__ fa_ = xa + start_a + 2 * lane;
        const float* __restrict__ fb_ = xa + start_b + 2 * lane;
        if (((reinterpret_cast<uintptr_t>(fa_) | reinterpret_cast<uintptr_t>(fb_)) & 7) == 0) {
          const float2* __restrict__ pa = reinterpret_cast<const float2*>(fa_);
          const float2* __restrict__ pb = reinterpret_cast<const float2*>(fb_);
          static_for<0, 16>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
            const float4 w = s_win4[j * 32 + lane];
            window_stage1(a[r0], a[r1], __ldg(pa + 32 * j), __ldg(pa + 32 * (j + 16)), __ldg(pb + 32 * j), __ldg(pb + 32 * (j + 16)),
                          make_float2(w.x, w.y), make_float2(w.z, w.w));
          });
        } else {
          static_for<0, 16>([&](auto jj) {
            constexpr int j = decltype(jj)::value;
            constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);
            const float4 w = s_win4[j * 32 + lane];
            window_stage1(a[r0], a[r1], make_float2(__ldg(fa_ + 64 * j), __ldg(fa_ + 64 * j + 1)),
                          make_float2(__ldg(fa_ + 64 * (j + 16)), __ldg(fa_ + 64 * (j + 16) + 1)),
                          make_float2(__ldg(fb_ + 64 * j), __ldg(fb_ + 64 * j + 1)),
                          make_float2(__ldg(fb_ + 64 * (j + 16)), __ldg(fb_ + 64 * (j + 16) + 1)), make_float2(w.x, w.y),
                          make_float2(w.z, w.w));
          });
        }
      } else {
        load_guarded();
      }
    }

    // ---- pass 1, exchange, pass 2: kernel_w32x2p.cuh
    dit2_stage_const<2>(a);
    dit2_stage_const<3>(a);
    dit2_stage_const<4>(a);
    dit2_stage_const<5>(a);
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
    dit2_stage_gen<1>(a, s_twb[0 * 32 + lane]);
    dit2_stage_gen<2>(a, s_twb[1 * 32 + lane]);
    dit2_stage_gen<3>(a, s_twb[2 * 32 + lane]);
    dit2_stage_gen<4>(a, s_twb[3 * 32 + lane]);
    dit2_stage_gen<5>(a, s_twb[4 * 32 + lane]);

    // ---- next pair: geometry now, its loads interleaved with the untangle
    int nit, np = p;
    bool has_next, nxt_fast;
    const float2* nsrc;
    {
      XsItem nxt = cur;
      advance(nxt, np);
      nit = nxt.it;
      has_next = nxt.valid;
      const long long nxt_off = pair_off(nxt, np);
      nxt_fast = has_next && pair_fast(nxt, np, nxt_off);
      nsrc = nxt_fast ? reinterpret_cast<const float2*>(g.pcm + nxt_off) + lane : idle_src;
    }

    // ---- untangle (kernel_w32x2p.cuh): mirrors fetched in place by shuffle, then 16 register steps
    const P2 p512 = mul2(bc(4.f), fma2(a[16].re, a[16].re, mul2(a[16].im, a[16].im)));
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
    P2 pk[16], pm[16];
    unsigned worst = 0;   // largest bit pattern among this lane's powers: >= 0x7f800000 means Inf or NaN
    static_for<0, 16>([&](auto ii) {
      constexpr int i = decltype(ii)::value;
      const P2 zmr = a[31 - i].re, zmi = a[31 - i].im;
      const float2 w = s_ut[i * 32 + lane];
      const C2 zk = a[i];
      const P2 ex = add2(zk.re, zmr), ey = add2(zk.im, neg(zmi));          // 2E
      const P2 ox = add2(zk.im, zmi), oy = add2(zmr, neg(zk.re));          // 2O
      const P2 xr = fma2(ox, bc(w.x), fma2(oy, bc(-w.y), ex));             // 2X[k]
      const P2 xi = fma2(ox, bc(w.y), fma2(oy, bc(w.x), ey));
      const P2 yr = fma2(ex, bc(2.f), neg(xr));                            // 2 conj X[1024-k]
      const P2 yi = fma2(ey, bc(2.f), neg(xi));
      P2 qk = fma2(xr, xr, mul2(xi, xi)), qm = fma2(yr, yr, mul2(yi, yi));
      if constexpr (i == 0) qm = P2(lane0 ? p512.v.x : qm.v.x, lane0 ? p512.v.y : qm.v.y);
      worst = max(max(worst, max(__float_as_uint(qk.v.x), __float_as_uint(qk.v.y))),
                  max(__float_as_uint(qm.v.x), __float_as_uint(qm.v.y)));
      // (1 - tau) |X| / N of both frames
      pk[i] = mul2(P2(sqrt_ftz(qk.v.x), sqrt_ftz(qk.v.y)), bc(x.mscale));
      pm[i] = mul2(P2(sqrt_ftz(qm.v.x), sqrt_ftz(qm.v.y)), bc(x.mscale));
      if constexpr (!LATE && HOPJ != 0) {
        static_for<(NLOAD * i) / 16, (NLOAD * (i + 1)) / 16>([&](auto mm) {
          constexpr int m = decltype(mm)::value;
          s[m] = ldg_nc_f2(nsrc + 32 * m);
        });
      }
    });
    const bool dirty = __any_sync(0xffffffffu, worst >= 0x7f800000u);

    // ---- the recurrence, in pair order.  The 16 state slots of a lane are cut into K groups with a turn of their
    //      own each, so warp w + 1 updates group c while warp w is already on group c + 1: the ordered part of the
    //      kernel is pipelined K deep.  In the aggregate pass (kind 1) the sign bit of a state value records that a
    //      non-finite frame wiped that bin inside this segment (X^ itself is never negative): the look-back must
    //      then drop what came in from earlier segments.
    if (p == 0 && cur.kind != 1 && cur.kind != 3 && cur.seg > 0) {
      // the segments this one starts from must have been published (chain: the previous one; look-back: all of them)
      if (lane0) {
        for (int j = cur.kind == 0 ? cur.seg - 1 : 0; j < cur.seg; ++j)
          while (ld_acquire_u32(x.flags + (long long)j * x.n_clips + cur.clip) != x.epoch) {}
      }
      __syncwarp();
    }
    static_for<0, K>([&](auto cc) {
      constexpr int c = decltype(cc)::value, i0 = c * (16 / K), i1 = i0 + 16 / K;
#ifdef SG_DEBUG
      if (!g_sg_dbg.break_chain)
#endif
      while (!mbar_try_wait(s_bar + c * NW + warp, turn)) {}
#ifdef SG_DEBUG
      if (lane == 0 && c == 0) dbg_count_iteration();
      // every slot of this group must hold the state the pair just before this one left (or this pair's own start state)
      if (p != 0) static_for<i0, i1>([&](auto ii) { constexpr int i = decltype(ii)::value; dbg_check(dbg_tags, i * 32 + lane, dbg_q, 2); });
#endif
      if (p == 0) {
        // first pair of a work item: the state the segment starts from
        // (an independent segment starts from zero unless its warm-up reaches back to the clip's first frame)
        const float* __restrict__ si = (cur.kind != 1 && (cur.kind != 3 || cur.f0 == 0) && x.state_in) ? x.state_in + (long long)cur.clip * kW32M : nullptr;
        if (cur.kind == 1 || cur.kind == 3 || cur.seg == 0) {
          static_for<i0, i1>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            s_state[i * 32 + lane] =
                si ? make_float2(si[xs_bin(i, lane, 0)], si[xs_bin(i, lane, 1) & (kW32M - 1)]) : make_float2(0.f, 0.f);
          });
        } else if (cur.kind == 0) {
          const float2* __restrict__ cv = x.carry + ((long long)(cur.seg - 1) * x.n_clips + cur.clip) * 512 + lane;
          static_for<i0, i1>([&](auto ii) { constexpr int i = decltype(ii)::value; s_state[i * 32 + lane] = __ldcg(cv + i * 32); });
        } else {
          // look-back: state at the start of segment s = dec^s state_in + sum_{j < s} dec^(s-1-j) aggregate_j (Horner);
          // an aggregate with its sign bit set wipes what came before it
          float2 acc[16 / K];
          static_for<i0, i1>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            acc[i - i0] = si ? make_float2(si[xs_bin(i, lane, 0)], si[xs_bin(i, lane, 1) & (kW32M - 1)]) : make_float2(0.f, 0.f);
          });
          for (int j = 0; j < cur.seg; ++j) {
            const float2* __restrict__ cv = x.carry + ((long long)j * x.n_clips + cur.clip) * 512 + lane;
            static_for<i0, i1>([&](auto ii) {
              constexpr int i = decltype(ii)::value;
              const float2 v = __ldcg(cv + i * 32);
              acc[i - i0] = make_float2(fmaf(signbit(v.x) ? 0.f : x.dec, acc[i - i0].x, fabsf(v.x)),
                                        fmaf(signbit(v.y) ? 0.f : x.dec, acc[i - i0].y, fabsf(v.y)));
            });
          }
          static_for<i0, i1>([&](auto ii) { constexpr int i = decltype(ii)::value; s_state[i * 32 + lane] = acc[i - i0]; });
        }
        __syncwarp();
      }
      if (!dirty && cur.kind != 1) {
        static_for<i0, i1>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          const float2 st = s_state[i * 32 + lane];
          const float ka = fmaf(x.tau, st.x, pk[i].v.x), kb = fmaf(x.tau, ka, pk[i].v.y);
          const float ma = fmaf(x.tau, st.y, pm[i].v.x), mb = fmaf(x.tau, ma, pm[i].v.y);
          pk[i] = P2(ka, kb);
          pm[i] = P2(ma, mb);
          s_state[i * 32 + lane] = has_b ? make_float2(kb, mb) : make_float2(ka, ma);
        });
      } else {
        static_for<i0, i1>([&](auto ii) {   // [SPEC] a non-finite X^ is set to 0 (and, kind 1, remembered in the sign)
          constexpr int i = decltype(ii)::value;
          const float2 st = s_state[i * 32 + lane];
          const float ka0 = fmaf(x.tau, fabsf(st.x), pk[i].v.x), ka = finite_or_zero(ka0);
          const float kb0 = fmaf(x.tau, ka, pk[i].v.y), kb = finite_or_zero(kb0);
          const float ma0 = fmaf(x.tau, fabsf(st.y), pm[i].v.x), ma = finite_or_zero(ma0);
          const float mb0 = fmaf(x.tau, ma, pm[i].v.y), mb = finite_or_zero(mb0);
          pk[i] = P2(ka, kb);
          pm[i] = P2(ma, mb);
          float nk = has_b ? kb : ka, nm = has_b ? mb : ma;
          if (cur.kind == 1) {
            const bool wk = signbit(st.x) || !(fabsf(ka0) <= 3.4028235e38f) || (has_b && !(fabsf(kb0) <= 3.4028235e38f));
            const bool wm = signbit(st.y) || !(fabsf(ma0) <= 3.4028235e38f) || (has_b && !(fabsf(mb0) <= 3.4028235e38f));
            nk = wk ? -nk : nk;
            nm = wm ? -nm : nm;
            pk[i] = P2(nk, nk);     // no output in this pass: the registers only feed the carry-out below
            pm[i] = P2(nm, nm);
          }
          s_state[i * 32 + lane] = make_float2(nk, nm);
        });
      }
#ifdef SG_DEBUG
      static_for<i0, i1>([&](auto ii) { constexpr int i = decltype(ii)::value; dbg_write(dbg_tags, i * 32 + lane, dbg_q + 1); });
      __threadfence_block();
#endif
      __syncwarp();
      if (lane0) mbar_arrive(s_bar + c * NW + (warp + 1 == NW ? 0 : warp + 1));
    });
    turn ^= 1;
#ifdef SG_DEBUG
    dbg_q += NW;
#endif
    const bool last = p == (cur.nfr + 1) / 2 - 1;
    if (last && !((cur.kind == 2 || cur.kind == 3) && cur.seg + 1 < x.segs)) {
      // last pair of a work item: hand the state to the next segment (or to the caller)
      if (cur.seg + 1 < x.segs) {
        const long long me = (long long)cur.seg * x.n_clips + cur.clip;
        float2* __restrict__ c = x.carry + me * 512 + lane;
        static_for<0, 16>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          c[i * 32] = has_b ? make_float2(pk[i].v.y, pm[i].v.y) : make_float2(pk[i].v.x, pm[i].v.x);
        });
        __threadfence();
        __syncwarp();
        if (lane0) st_release_u32(x.flags + me, x.epoch);
      } else if (cur.kind != 1 && x.state_out != nullptr) {
        float* __restrict__ so = x.state_out + (long long)cur.clip * kW32M;
        static_for<0, 16>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          so[xs_bin(i, lane, 0)] = has_b ? pk[i].v.y : pk[i].v.x;
          so[xs_bin(i, lane, 1) & (kW32M - 1)] = has_b ? pm[i].v.y : pm[i].v.x;
        });
      }
    }
    if constexpr (LATE && HOPJ != 0) {
      static_for<0, NLOAD>([&](auto mm) { constexpr int m = decltype(mm)::value; s[m] = ldg_nc_f2(nsrc + 32 * m); });
    }

    // ---- epilogue: X^ -> dB / byte / colour of both frames (nothing to write in the aggregate pass)
    if (cur.kind != 1 && 2 * p >= cur.skip) {
      T* __restrict__ row_a = out + ((long long)cur.clip * x.out_clip_rows + ta) * (long long)kW32M;
      T* __restrict__ row_b = row_a + kW32M;
      if constexpr (OUT == kOutU8 || OUT == kOutRgba8) {
        const P2 scale = bc(2.f * ep.byte_a);
        static_for<0, 16>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          const int k = lane + 32 * i;
          int mk = kW32M - k;
          if constexpr (i == 0) { if (lane0) mk = 512; }
          const P2 vk = fma2(P2(lg2_ftz(pk[i].v.x), lg2_ftz(pk[i].v.y)), scale, bc(ep.byte_b0));
          const P2 vm = fma2(P2(lg2_ftz(pm[i].v.x), lg2_ftz(pm[i].v.y)), scale, bc(ep.byte_b0));
          const unsigned ka = byte_of_scaled(vk.v.x), kb = byte_of_scaled(vk.v.y);
          const unsigned ma = byte_of_scaled(vm.v.x), mb = byte_of_scaled(vm.v.y);
          if constexpr (OUT == kOutU8) {
            sb16[k] = (uint16_t)__byte_perm(ka, kb, 0x0040);
            sb16[mk] = (uint16_t)__byte_perm(ma, mb, 0x0040);
          } else {
            row_a[k] = __ldg(ep.lut + ka); row_a[mk] = __ldg(ep.lut + ma);
            if (has_b) { row_b[k] = __ldg(ep.lut + kb); row_b[mk] = __ldg(ep.lut + mb); }
          }
        });
        if constexpr (OUT == kOutU8) {
          __syncwarp();
          const uint4* s16 = reinterpret_cast<const uint4*>(sb16);
          uint2* ra = reinterpret_cast<uint2*>(row_a);
          uint2* rb = reinterpret_cast<uint2*>(row_b);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 w = s16[c * 32 + lane];
            ra[c * 32 + lane] = make_uint2(__byte_perm(w.x, w.y, 0x6420), __byte_perm(w.z, w.w, 0x6420));
            if (has_b) rb[c * 32 + lane] = make_uint2(__byte_perm(w.x, w.y, 0x7531), __byte_perm(w.z, w.w, 0x7531));
          }
        }
      } else {
        static_for<0, 16>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          const int k = lane + 32 * i;
          int mk = kW32M - k;
          if constexpr (i == 0) { if (lane0) mk = 512; }
          P2 vk = pk[i], vm = pm[i];
          if constexpr (OUT == kOutF32Db) {
            vk = mul2(P2(lg2_ftz(vk.v.x), lg2_ftz(vk.v.y)), bc(2.f * ep.db_scale));
            vm = mul2(P2(lg2_ftz(vm.v.x), lg2_ftz(vm.v.y)), bc(2.f * ep.db_scale));
          }
          row_a[k] = vk.v.x; row_a[mk] = vm.v.x;
          if (has_b) { row_b[k] = vk.v.y; row_b[mk] = vm.v.y; }
        });
      }
    }
    __syncwarp();
    if (!has_next) break;
    it = nit;
    p = np;
    cur_fast = nxt_fast;
  }
}

}  // namespace sg
