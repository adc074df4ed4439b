// n_fft = 2048 with smoothingTimeConstant > 0 in ONE pass: the frame-pair kernel of kernel_w32x2p.cuh with the
// AnalyserNode recurrence  X^_t[k] = tau X^_{t-1}[k] + (1 - tau) |X_t[k]|  (3D/visualizer.js:351,357,362 set tau per
// mode; [SPEC] step 4) applied on chip between the untangle and the dB / byte epilogue.  No magnitude scratch in HBM.
//
// The FFTs of a clip's frames are independent; only the recurrence is ordered.  A CTA owns a SEGMENT of consecutive
// frames of one clip and its warps are specialised:
//   producers (10 warps)  take the segment's frame pairs round robin and run loader -> window -> FFT -> untangle exactly
//                         as kernel_w32x2p.cuh, but instead of an epilogue they leave the pair's powers |2X|^2 in their
//                         own (now idle) exchange planes, 16 x STS.128 per lane, and signal a "full" mbarrier;
//   consumers (2 warps)   walk the pairs IN ORDER; consumer c owns state slots 8c .. 8c+7 of every lane (slot (i, lane)
//                         = bins lane + 32 i and its mirror) and keeps X^ of those 512 bins in registers for the whole
//                         segment: wait full, 8 x LDS.128, release the planes ("empty" mbarrier), sqrt, two FMAs per
//                         value, lg2, byte, staged row stores.  A consumer needs a fraction of the time the producers
//                         take to deliver a pair, so the ordered part does not throttle the FFTs, and a producer only
//                         waits if its previous pair has not been read by the time it reaches its next exchange.
// (A first version passed the state update itself from warp to warp through an mbarrier chain, every warp doing its own
//  epilogue: 472 M frames/s with 8 warps, 395 M with 12 -- the ring of turns makes every warp wait for the slowest.)
// Segments of one clip are chained through a carry vector in global memory (per consumer half, with a ready flag):
//   mode 0 (chain)  many clips: tasks (segment, clip) are dealt to CTAs segment-major, so a task's predecessor was
//                   started ~n_clips / gridDim tasks earlier.  Exactly the sequential arithmetic, bit for bit.
//   mode 1 (few)    fewer clips than CTAs: one task per CTA, run twice -- first without output from a zero state,
//                   which yields the segment's aggregate (the recurrence is linear), then, after a look-back over the
//                   aggregates of the segments before it (all produced concurrently), again from the true state with
//                   output.  Twice the arithmetic, on a GPU that would otherwise idle; one launch.
// [SPEC] "non-finite X^ -> 0" is applied per value; in the aggregate pass the sign bit of a state value records that
// a non-finite frame wiped that bin inside the segment (X^ itself is never negative), so the look-back drops what
// came in from earlier segments.
#pragma once
#include "kernel_w32x2p.cuh"

namespace sg {

struct XsGeom {
  long long n_clips;
  long long out_clip_rows;  // output rows between consecutive clips (>= frames_per_clip: a frame range of longer clips)
  int seg_frames;           // frames per segment (even); the last segment of a clip may be shorter
  int segs;                 // segments per clip
  float tau;
  float mscale;             // (1 - tau) / norm: sqrt(|2X|^2) * mscale = (1 - tau) |X| / N
  float dec;                // tau^seg_frames (mode 1 look-back)
  const float* state_in;    // [n_clips][1024], natural bin order; nullptr = zeros
  float* state_out;         // [n_clips][1024], natural bin order; nullptr = not wanted
  float2* carry;            // [segs * n_clips][16][32]: state at the END of a segment (mode 1: its aggregate)
  unsigned* flags;          // [segs * n_clips][2]: == epoch once consumer c's half of the carry is visible
  unsigned epoch;
  int mode;
};

constexpr int kXsProducers = 10, kXsConsumers = 2, kXsWarps = kXsProducers + kXsConsumers;
constexpr int kXsStageBytes = 2 * 1024;                                // per consumer: 512 (A, B) byte pairs
constexpr int kXsSmemBytes = kXpTableBytes + kXsStageBytes + 2 * kXsProducers * 8 + kXsProducers * kXpWarpBytes;
constexpr int kXsMaxRegs = (65536 / (kXsWarps * 32)) / 8 * 8;

// the work item a pair belongs to: (segment, clip) and what to do at its first / last pair
struct XsItem {
  int it;        // index in this CTA's item list
  int clip, seg;
  int f0, nfr;   // first frame of the segment within the clip, frames in it
  int kind;      // 0 chain, 1 aggregate pass (no output, zero state), 2 emit pass after look-back
  bool valid;
};
__device__ __forceinline__ XsItem xs_item(const XsGeom& x, int fpc, int it) {
  XsItem c;
  c.it = it;
  const unsigned n_clips = (unsigned)x.n_clips, n_tasks = (unsigned)x.segs * n_clips;   // < 2^31 (host checks)
  const unsigned task = x.mode == 0 ? blockIdx.x + (unsigned)it * gridDim.x : blockIdx.x;
  c.valid = task < n_tasks && (x.mode == 0 || it < 2);
  c.kind = x.mode == 0 ? 0 : 1 + it;
  c.seg = (int)(task / n_clips);
  c.clip = (int)(task - (unsigned)c.seg * n_clips);
  c.f0 = c.seg * x.seg_frames;
  c.nfr = min(x.seg_frames, fpc - c.f0);
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

template <int OUT, int HOPJ>
__global__ void __launch_bounds__(kXsWarps * 32, 1)
stft_w32x2s_kernel(FrameGeom g, XsGeom x, W32Plan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using T = typename OutElem<OUT>::type;
  constexpr int HOP = 64 * HOPJ, NLOAD = 32 + HOPJ, NP = kXsProducers;
  extern __shared__ float4 smem_raw[];
  float4* s_win4 = smem_raw;                                           // [16][32] (w2[l+32j], w2[l+32(j+16)])
  float2* s_twb = reinterpret_cast<float2*>(s_win4 + 16 * 32);         // [5][32]  W_{32*2^u}^lane
  float2* s_ut = s_twb + 5 * 32;                                       // [16][32] W_2048^{lane + 32 i}
  uint16_t* s_stage = reinterpret_cast<uint16_t*>(s_ut + 16 * 32);     // [2][512] (A, B) byte pairs, one stage per consumer
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_stage + 2 * 512);  // [NP] producer w's powers are in its planes
  uint64_t* s_empty = s_full + NP;                                     // [NP] both consumers have read them
  unsigned char* planes = reinterpret_cast<unsigned char*>(s_empty + NP);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool lane0 = lane == 0;

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
    if (threadIdx.x < NP) {
      mbar_init(s_full + threadIdx.x, 1);
      mbar_init(s_empty + threadIdx.x, kXsConsumers);
    }
  }
  __syncthreads();
  const int fpc = (int)g.frames_per_clip;

  if (warp >= NP) {
    // =================================================================================== consumer
    // consumer 0 owns the bins a lane computes directly (k = lane + 32 i < 512), consumer 1 their mirrors (bins 513 ..
    // 1023, and bin 512 in lane 0's slot 0): each keeps X^ of its 512 bins in registers and writes its half of the rows
    const int c = warp - NP;
    uint16_t* sb16 = s_stage + c * 512;     // this consumer's byte stage: 512 (A, B) pairs
    float st[16];                           // X^ of this consumer's bins: slot i of every lane
    unsigned full_par = 0;
    int w = 0;                              // producer that delivers the current pair
    for (XsItem cur = xs_item(x, fpc, 0); cur.valid; cur = xs_item(x, fpc, cur.it + 1)) {
      const int npairs = (cur.nfr + 1) / 2;
      // ---- first pair of a work item: the state the segment starts from
      {
        const float* __restrict__ si = (cur.kind != 1 && x.state_in) ? x.state_in + (long long)cur.clip * kW32M : nullptr;
        static_for<0, 16>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          st[i] = si ? si[xs_bin(i, lane, c) & (kW32M - 1)] : 0.f;
        });
        if (cur.kind != 1 && cur.seg > 0) {
          // chain: the previous segment's final state; look-back: Horner over the aggregates of all earlier segments,
          // state = dec^s state_in + sum_{j < s} dec^(s-1-j) aggregate_j (an aggregate with its sign set wipes the rest)
          for (int j = cur.kind == 0 ? cur.seg - 1 : 0; j < cur.seg; ++j) {
            const long long tj = (long long)j * x.n_clips + cur.clip;
            if (lane0) while (ld_acquire_u32(x.flags + 2 * tj + c) != x.epoch) {}
            __syncwarp();
            const float* __restrict__ cv = reinterpret_cast<const float*>(x.carry) + (tj * 2 + c) * 512 + lane;
            static_for<0, 16>([&](auto ii) {
              constexpr int i = decltype(ii)::value;
              const float v = __ldcg(cv + i * 32);
              if (cur.kind == 0) st[i] = v;
              else st[i] = fmaf(signbit(v) ? 0.f : x.dec, st[i], fabsf(v));
            });
          }
        }
      }
      for (int p = 0; p < npairs; ++p) {
        const bool has_b = 2 * p + 1 < cur.nfr;
        // ---- the pair's powers from producer w's planes
        while (!mbar_try_wait(s_full + w, full_par)) {}
        const float2* __restrict__ src = reinterpret_cast<const float2*>(planes + w * kXpWarpBytes) + c * 512 + lane;
        float2 pw[16];
        static_for<0, 16>([&](auto ii) { constexpr int i = decltype(ii)::value; pw[i] = src[i * 32]; });
        __syncwarp();
        if (lane0) mbar_arrive(s_empty + w);
        if (++w == NP) { w = 0; full_par ^= 1; }
        // ---- the recurrence for frames A then B.  [SPEC] a non-finite X^ is set to 0: one integer max over the
        //      lane's powers decides whether the per-value path is needed (Inf and NaN have the largest bit patterns)
        unsigned worst = 0;
        static_for<0, 16>([&](auto ii) {
          constexpr int i = decltype(ii)::value;
          worst = max(worst, max(__float_as_uint(pw[i].x), __float_as_uint(pw[i].y)));
        });
        P2 v[16];                             // X^ of (frame A, frame B)
        if (cur.kind != 1 && !__any_sync(0xffffffffu, worst >= 0x7f800000u)) {
          static_for<0, 16>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const float a_ = fmaf(x.tau, st[i], sqrt_ftz(pw[i].x) * x.mscale);
            const float b_ = fmaf(x.tau, a_, sqrt_ftz(pw[i].y) * x.mscale);
            v[i] = P2(a_, b_);
            st[i] = has_b ? b_ : a_;
          });
        } else {
          static_for<0, 16>([&](auto ii) {
            constexpr int i = decltype(ii)::value;
            const float a0 = fmaf(x.tau, fabsf(st[i]), sqrt_ftz(pw[i].x) * x.mscale), a_ = finite_or_zero(a0);
            const float b0 = fmaf(x.tau, a_, sqrt_ftz(pw[i].y) * x.mscale), b_ = finite_or_zero(b0);
            v[i] = P2(a_, b_);
            float n_ = has_b ? b_ : a_;
            if (cur.kind == 1) {   // aggregate pass: remember in the sign that the bin was wiped inside this segment
              const bool wiped = signbit(st[i]) || !(fabsf(a0) <= 3.4028235e38f) || (has_b && !(fabsf(b0) <= 3.4028235e38f));
              n_ = wiped ? -n_ : n_;
            }
            st[i] = n_;
          });
        }
        // ---- epilogue: X^ -> dB / byte / colour of both frames (nothing to write in the aggregate pass)
        if (cur.kind != 1) {
          const int ta = cur.f0 + 2 * p;
          T* __restrict__ row_a = out + ((long long)cur.clip * x.out_clip_rows + ta) * (long long)kW32M;
          T* __restrict__ row_b = row_a + kW32M;
          if constexpr (OUT == kOutU8 || OUT == kOutRgba8) {
            const P2 scale = bc(2.f * ep.byte_a);
            static_for<0, 16>([&](auto ii) {
              constexpr int i = decltype(ii)::value;
              const int bin = xs_bin(i, lane, c);
              const P2 bv = fma2(P2(lg2_ftz(v[i].v.x), lg2_ftz(v[i].v.y)), scale, bc(ep.byte_b0));
              const unsigned ba = byte_of_scaled(bv.v.x), bb = byte_of_scaled(bv.v.y);
              if constexpr (OUT == kOutU8) {
                sb16[bin & 511] = (uint16_t)__byte_perm(ba, bb, 0x0040);
              } else {
                row_a[bin] = __ldg(ep.lut + ba);
                if (has_b) row_b[bin] = __ldg(ep.lut + bb);
              }
            });
            if constexpr (OUT == kOutU8) {
              __syncwarp();
              const uint4* s16 = reinterpret_cast<const uint4*>(sb16);
              uint2* ra = reinterpret_cast<uint2*>(row_a) + c * 64;
              uint2* rb = reinterpret_cast<uint2*>(row_b) + c * 64;
#pragma unroll
              for (int cc = 0; cc < 2; ++cc) {
                const uint4 wv = s16[cc * 32 + lane];
                ra[cc * 32 + lane] = make_uint2(__byte_perm(wv.x, wv.y, 0x6420), __byte_perm(wv.z, wv.w, 0x6420));
                if (has_b) rb[cc * 32 + lane] = make_uint2(__byte_perm(wv.x, wv.y, 0x7531), __byte_perm(wv.z, wv.w, 0x7531));
              }
              __syncwarp();
            }
          } else {
            static_for<0, 16>([&](auto ii) {
              constexpr int i = decltype(ii)::value;
              const int bin = xs_bin(i, lane, c);
              P2 f = v[i];
              if constexpr (OUT == kOutF32Db) f = mul2(P2(lg2_ftz(f.v.x), lg2_ftz(f.v.y)), bc(2.f * ep.db_scale));
              row_a[bin] = f.v.x;
              if (has_b) row_b[bin] = f.v.y;
            });
          }
        }
      }
      // ---- last pair of the work item done: hand the state to the next segment (or to the caller)
      if (!(cur.kind == 2 && cur.seg + 1 < x.segs)) {
        if (cur.seg + 1 < x.segs) {
          const long long me = (long long)cur.seg * x.n_clips + cur.clip;
          float* __restrict__ cv = reinterpret_cast<float*>(x.carry) + (me * 2 + c) * 512 + lane;
          static_for<0, 16>([&](auto ii) { constexpr int i = decltype(ii)::value; cv[i * 32] = st[i]; });
          __threadfence();
          __syncwarp();
          if (lane0) st_release_u32(x.flags + 2 * me + c, x.epoch);
        } else if (cur.kind != 1 && x.state_out != nullptr) {
          float* __restrict__ so = x.state_out + (long long)cur.clip * kW32M;
          static_for<0, 16>([&](auto ii) { constexpr int i = decltype(ii)::value; so[xs_bin(i, lane, c) & (kW32M - 1)] = st[i]; });
        }
      }
    }
    return;
  }

  // ===================================================================================== producer
  float4* xp = reinterpret_cast<float4*>(planes + warp * kXpWarpBytes);   // exchange planes, then the pair's powers
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
    return c.valid && 2 * p + 1 < c.nfr && t >= t_lo && t <= t_hi && ((pcm_lo + ((unsigned)off << 2)) & 7u) == 0;
  };
  // the pair NP places further down this CTA's pair sequence.  Only (item index, pair index) is carried from one
  // iteration to the next; the item's fields are re-derived where they are needed.
  auto advance = [&](XsItem& c, int& p) {
    p += NP;
    while (c.valid && p >= (c.nfr + 1) / 2) {
      p -= (c.nfr + 1) / 2;
      c = xs_item(x, fpc, c.it + 1);
    }
  };

  int it, p = warp - NP;
  bool cur_fast;
  {
    XsItem c0 = xs_item(x, fpc, 0);
    advance(c0, p);
    if (!c0.valid) return;
    it = c0.it;
    cur_fast = pair_fast(c0, p, pair_off(c0, p));
  }
  const int partner = (32 - lane) & 31;
  unsigned empty_par = 1;       // a fresh mbarrier reads as "the phase before has completed": the planes start free

  float2 s[NLOAD];
  const float2* idle_src = reinterpret_cast<const float2*>(pl.win) + lane;   // 4096 readable floats (build_plan)
  {
    const XsItem c0 = xs_item(x, fpc, it);
    const float2* src = cur_fast ? reinterpret_cast<const float2*>(g.pcm + pair_off(c0, p)) + lane : idle_src;
    static_for<0, NLOAD>([&](auto mm) { constexpr int m = decltype(mm)::value; s[m] = ldg_nc_f2(src + 32 * m); });
  }

  while (true) {
    const XsItem cur = xs_item(x, fpc, it);
    // ---- steps 1-2 (+ FFT stage 1): window both frames, bit-reversed into registers
    C2 a[32];
    if (cur_fast) {
      static_for<0, 16>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        constexpr int r0 = bitrev(j, 5), r1 = bitrev(j + 16, 5);   // r1 == r0 + 1
        const float4 w = s_win4[j * 32 + lane];
        window_stage1(a[r0], a[r1], s[j], s[j + 16], s[j + HOPJ], s[j + 16 + HOPJ], make_float2(w.x, w.y),
                      make_float2(w.z, w.w));
      });
    } else {
      // clip edges / zero history / a segment's odd last frame (frame B reads as frame A and is dropped)
      const float* __restrict__ xa = g.pcm + cur.clip * g.clip_stride;
      const long long start_a = g.start0 + (long long)(cur.f0 + 2 * p) * HOP, start_b = start_a + (2 * p + 1 < cur.nfr ? HOP : 0);
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
    }

    // ---- pass 1, exchange, pass 2: kernel_w32x2p.cuh
    dit2_stage_const<2>(a);
    dit2_stage_const<3>(a);
    dit2_stage_const<4>(a);
    dit2_stage_const<5>(a);
    // the planes still hold the previous pair's powers until both consumers have read them
    while (!mbar_try_wait(s_empty + warp, empty_par)) {}
    empty_par ^= 1;
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

    // ---- untangle (kernel_w32x2p.cuh): mirrors fetched in place by shuffle, then 16 register steps; each step's
    //      powers (frames A and B at the bin, A and B at its mirror) go to the planes as two 8-byte stores
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
      const P2 qk = fma2(xr, xr, mul2(xi, xi));
      P2 qm = fma2(yr, yr, mul2(yi, yi));
      if constexpr (i == 0) qm = P2(lane0 ? p512.v.x : qm.v.x, lane0 ? p512.v.y : qm.v.y);
      reinterpret_cast<float2*>(xp)[i * 32 + lane] = qk.v;          // plane 0: the bins this lane computes directly
      reinterpret_cast<float2*>(xp)[512 + i * 32 + lane] = qm.v;    // plane 1: their mirrors
      static_for<(NLOAD * i) / 16, (NLOAD * (i + 1)) / 16>([&](auto mm) {
        constexpr int m = decltype(mm)::value;
        s[m] = ldg_nc_f2(nsrc + 32 * m);
      });
    });
    __syncwarp();
    if (lane0) mbar_arrive(s_full + warp);
    if (!has_next) break;
    it = nit;
    p = np;
    cur_fast = nxt_fast;
  }
}

}  // namespace sg
