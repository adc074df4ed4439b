// n_fft = 8192 (and 4096 where the even/odd kernel cannot load: hops that are not a multiple of 4) with
// smoothingTimeConstant > 0 in ONE pass: the register-family kernel of kernel_wreg.cuh (T = M/32 threads per frame, three
// register passes, last pass merged with the untangle) with the AnalyserNode recurrence
// X^_t[k] = tau X^_{t-1}[k] + (1 - tau) |X_t[k]|  (3D/visualizer.js:351,357,362; [SPEC] step 4) between the untangle and the
// dB / byte epilogue.  A CTA owns a SEGMENT of consecutive frames of one clip (chain mode of kernel_w32x2s.cuh: segments of
// a clip are chained through a carry vector and a flag in global memory, tasks dealt segment-major, cooperative launch); its
// FPC = 256/T frame slots take the segment's frames round robin, run the FFT concurrently and pass the state update --
// 32 x (LDS, FMA, STS) per thread on the M floats of X^ in shared memory -- from slot to slot through named barriers
// (slot k arrives, slot k + 1 waits).  One frame per turn: exactly the sequential arithmetic.
// [SPEC] "non-finite X^ -> 0": a non-finite sample makes every bin of its frame non-finite, so the per-frame flag of the
// register family decides: such a frame leaves X^ = 0 in every bin.
#pragma once
#include "kernel_wreg.cuh"
#include "kernel_w32x2s.cuh"   // XsGeom, acquire/release helpers

namespace sg {

template <int LOG2M>
struct WsShape {
  using W = WregShape<LOG2M>;
  static constexpr int kStateBytes = W::M * 4;
  static constexpr int kSmemBytes = W::kSmemBytes + kStateBytes;
};

struct WsItem {
  int it, clip, seg, f0, nfr, skip;   // skip: leading frames without output (XsGeom mode 2: warm-up of an independent segment)
  bool valid;
};
__device__ __forceinline__ WsItem ws_item(const XsGeom& x, int fpc, int it) {
  WsItem c;
  c.it = it;
  const unsigned n_clips = (unsigned)x.n_clips, n_tasks = (unsigned)x.segs * n_clips;   // < 2^31 (host checks)
  const unsigned task = blockIdx.x + (unsigned)it * gridDim.x;
  c.valid = task < n_tasks;
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

// smoothed, normalised magnitude -> output element
template <int OUT>
__device__ __forceinline__ typename OutElem<OUT>::type emit_smoothed(float m, const Epilogue& e) {
  if constexpr (OUT == kOutF32Mag) {
    return m;
  } else if constexpr (OUT == kOutF32Db) {
    return 2.f * e.db_scale * lg2_ftz(m);
  } else {
    const unsigned b = byte_of_scaled(fmaf(2.f * e.byte_a, lg2_ftz(m), e.byte_b0));
    if constexpr (OUT == kOutU8) return (uint8_t)b;
    else return __ldg(e.lut + b);
  }
}

template <int LOG2M, int OUT>
__global__ void __launch_bounds__(kWregThreads, 2)
stft_wreg_s_kernel(FrameGeom g, XsGeom x, WregPlan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using S = WregShape<LOG2M>;
  using TO = typename OutElem<OUT>::type;
  constexpr int M = S::M, N = S::N, T = S::T, R = S::R, R3 = S::R3, L2 = S::L2, FPC = S::FPC;
  static_assert(R3 > 0 && S::R2 == 5, "M = 2048 or 4096");
  extern __shared__ float4 smem_raw[];
  const int tid = threadIdx.x, fs = tid / T, t = tid % T;
  unsigned char* fbase = reinterpret_cast<unsigned char*>(smem_raw) + fs * S::kFrameBytes;
  float2* A = reinterpret_cast<float2*>(fbase);
  unsigned char* sb = fbase + S::kTileF2 * 8;
  float* s_state = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(smem_raw) + S::kSmemBytes);   // [32][T]
  auto frame_sync = [fs] { asm volatile("bar.sync %0, %1;" ::"r"(fs + 1), "n"(T) : "memory"); };
  // the turn chain: barrier 8 + k is shared by slot k (arrives after its update) and slot k + 1 (waits before its own)
  auto turn_wait = [fs] { asm volatile("bar.sync %0, %1;" ::"r"(8 + (fs + FPC - 1) % FPC), "n"(2 * T) : "memory"); };
  auto turn_pass = [fs] { asm volatile("bar.arrive %0, %1;" ::"r"(8 + fs), "n"(2 * T) : "memory"); };
  const int fpc = (int)g.frames_per_clip;
  const float2* __restrict__ win2 = reinterpret_cast<const float2*>(pl.win);

  // pass-3 ownership of this thread (kernel_wreg.cuh): column pair and q's; the same in every frame slot
  constexpr int S3 = S::S3, NP = 16 / (T / 16), NQ = 2 * NP;
  const int pi = t & 15, qg = t >> 4;
  const bool self = pi == 0, self0 = self && qg == 0;
  const int col[2] = {pi, self ? 16 : 32 - pi};

  bool first_turn = true;          // slot 0 has nobody to wait for at the CTA's very first frame
  for (int it = 0;; ++it) {
    const WsItem cur = ws_item(x, fpc, it);
    if (!cur.valid) break;
    const int ngroups = (cur.nfr + FPC - 1) / FPC;
    if (cur.seg > 0 && x.mode != 2) {
      // the segment this one starts from must have been published (every thread of slot 0 reads the carry below)
      if (tid == 0)
        while (ld_acquire_u32(x.flags + (long long)(cur.seg - 1) * x.n_clips + cur.clip) != x.epoch) {}
    }
    for (int gq = 0; gq < ngroups; ++gq) {
      const int p = gq * FPC + fs;
      const bool live = p < cur.nfr;
      const bool store = live && p >= cur.skip;             // warm-up frames of an independent segment leave no row
      const int tf = cur.f0 + min(p, cur.nfr - 1);            // idle slots recompute the segment's last frame, store nothing
      const long long start = g.start0 + (long long)tf * g.hop;
      const float* __restrict__ xc = g.pcm + cur.clip * g.clip_stride;

      // ---- pass 1: load + window + stages 1-5
      float2 v[32];
      const bool inside = start >= 0 && start + N <= g.clip_len;
      if (inside && ((reinterpret_cast<uintptr_t>(xc + start) & 7) == 0)) {
        const float2* __restrict__ src = reinterpret_cast<const float2*>(xc + start) + t;
        static_for<0, 32>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          const float2 sv = __ldg(src + T * j), w = __ldg(win2 + t + T * j);
          v[bitrev(j, 5)] = make_float2(sv.x * w.x, sv.y * w.y);
        });
      } else if (inside) {
        const float* __restrict__ src = xc + start + 2 * t;
        static_for<0, 32>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          const float a0 = __ldg(src + 2 * T * j), a1 = __ldg(src + 2 * T * j + 1);
          const float2 w = __ldg(win2 + t + T * j);
          v[bitrev(j, 5)] = make_float2(a0 * w.x, a1 * w.y);
        });
      } else {
        static_for<0, 32>([&](auto jj) {
          constexpr int j = decltype(jj)::value;
          const long long s0 = start + 2 * (t + T * j), s1 = s0 + 1;
          const float a0 = (s0 >= 0 && s0 < g.clip_len) ? __ldg(xc + s0) : 0.f;
          const float a1 = (s1 >= 0 && s1 < g.clip_len) ? __ldg(xc + s1) : 0.f;
          const float2 w = __ldg(win2 + t + T * j);
          v[bitrev(j, 5)] = make_float2(a0 * w.x, a1 * w.y);
        });
      }
      dit_stage_const<1>(v);
      dit_stage_const<2>(v);
      dit_stage_const<3>(v);
      dit_stage_const<4>(v);
      dit_stage_const<5>(v);
      static_for<0, 32>([&](auto kk) { constexpr int k = decltype(kk)::value; A[t * kWregStride + k] = v[k]; });
      frame_sync();

      // ---- pass 2: stages 6-10 on column ka of the rows bitrev5(q) * L2 + hi2, twiddles from per-column bases
      {
        const int ka = t & 31, hi2p = t >> 5;
        static_for<0, 32>([&](auto qq) {
          constexpr int q = decltype(qq)::value;
          v[q] = A[(bitrev(q, 5) * L2 + hi2p) * kWregStride + ka];
        });
        static_for<1, 6>([&](auto uu) {
          constexpr int u = decltype(uu)::value, half = 1 << (u - 1);
          const float2 base = __ldg(pl.tw2 + (half - 1) * 32 + ka);
          static_for<0, half>([&](auto pp) {
            constexpr int pq = decltype(pp)::value;
            const float2 w = twiddle_times<pq, 2 * half>(base);
            static_for<0, 16 / half>([&](auto bb) {
              constexpr int i0 = decltype(bb)::value * 2 * half + pq;
              bfly(v[i0], v[i0 + half], w.x, w.y);
            });
          });
        });
        static_for<0, 32>([&](auto qq) {
          constexpr int q = decltype(qq)::value;
          A[(bitrev(q, 5) * L2 + hi2p) * kWregStride + ka] = v[q];
        });
      }
      frame_sync();

      // ---- pass 3 merged with the untangle (kernel_wreg.cuh): the thread's 32 powers
      int qs[2][NQ];
      static_for<0, NP>([&](auto ss) {
        constexpr int sl = decltype(ss)::value;
        const int j = qg * NP + sl;
        qs[0][2 * sl] = j;
        qs[0][2 * sl + 1] = self ? (j == 0 ? 16 : 32 - j) : 31 - j;
        qs[1][2 * sl] = 31 - j;
        qs[1][2 * sl + 1] = j;
      });
      static_for<0, 2 * NQ>([&](auto ee) {
        constexpr int e = decltype(ee)::value, xx = e / NQ, i = e % NQ;
        const int q = qs[xx][i];
        const float2* src = A + (int)(__brev((unsigned)q) >> 27) * S3 * kWregStride + col[xx];
        static_for<0, S3>([&](auto hh) {
          constexpr int h = decltype(hh)::value;
          v[e * S3 + h] = src[bitrev(h, R3) * kWregStride];
        });
      });
      static_for<0, 2 * NQ>([&](auto ee) {
        constexpr int e = decltype(ee)::value, xx = e / NQ, i = e % NQ;
        const float2* tw = pl.tw3 + (qs[xx][i] * (S3 - 1)) * 32 + col[xx];
        const float2 w0 = __ldg(tw);
        if constexpr (S3 == 2) {
          bfly(v[e * 2], v[e * 2 + 1], w0.x, w0.y);
        } else {
          const float2 w1 = __ldg(tw + 32);
          bfly(v[e * 4], v[e * 4 + 1], w0.x, w0.y);
          bfly(v[e * 4 + 2], v[e * 4 + 3], w0.x, w0.y);
          bfly(v[e * 4], v[e * 4 + 2], w1.x, w1.y);
          bfly(v[e * 4 + 1], v[e * 4 + 3], w1.y, -w1.x);
        }
      });
      const bool bad = !(fabsf(v[0].x) <= 3.4028235e38f) || !(fabsf(v[0].y) <= 3.4028235e38f);
      float mk_[16], mm_[16];      // (1 - tau) |X| / N at bin k and at its mirror, unit u = e (S3/2) + h
      static_for<0, 2 * NQ>([&](auto ee) {
        constexpr int e = decltype(ee)::value, xx = e / NQ, i = e % NQ;
        static_for<0, S3 / 2>([&](auto hh) {
          constexpr int h = decltype(hh)::value;
          const int k = col[xx] + 32 * (qs[xx][i] + 32 * h);
          const float2 zk = v[e * S3 + h];
          float2 zm = v[((1 - xx) * NQ + i) * S3 + (S3 - 1 - h)];
          const float2 zs = v[(xx * NQ + (i ^ 1)) * S3 + (S3 - 1 - h)];
          zm.x = self ? zs.x : zm.x;
          zm.y = self ? zs.y : zm.y;
          if constexpr (xx == 0 && i < 2) {
            const float2 z0 = v[i * S3 + (i == 0 ? (S3 - h) % S3 : S3 - 1 - h)];
            zm.x = self0 ? z0.x : zm.x;
            zm.y = self0 ? z0.y : zm.y;
          }
          const float2 w = __ldg(pl.ut + k);
          const float ex = zk.x + zm.x, ey = zk.y - zm.y;       // 2E
          const float ox = zk.y + zm.y, oy = zm.x - zk.x;       // 2O
          const float xr = fmaf(ox, w.x, fmaf(-oy, w.y, ex));   // 2X[k]
          const float xi = fmaf(ox, w.y, fmaf(oy, w.x, ey));
          const float yr = fmaf(2.f, ex, -xr);                  // 2 conj X[M-k]
          const float yi = fmaf(2.f, ey, -xi);
          float pk = fmaf(xr, xr, xi * xi), pm = fmaf(yr, yr, yi * yi);
          if constexpr (e == 0 && h == 0) {
            const float2 zh = v[S3 / 2];
            pm = self0 ? 4.f * fmaf(zh.x, zh.x, zh.y * zh.y) : pm;
          }
          mk_[e * (S3 / 2) + h] = sqrt_ftz(pk) * x.mscale;
          mm_[e * (S3 / 2) + h] = sqrt_ftz(pm) * x.mscale;
        });
      });

      // ---- the recurrence, in frame order: slot fs waits for the slot before it
      if (!(first_turn && fs == 0)) turn_wait();
      first_turn = false;
      if (gq == 0 && fs == 0) {
        // first frame of a work item: the state the segment starts from
        if (cur.seg == 0 || x.mode == 2) {
          // (an independent segment starts from zero unless its warm-up reaches back to the clip's first frame)
          const float* __restrict__ si = (x.state_in && cur.f0 == 0) ? x.state_in + (long long)cur.clip * M : nullptr;
          static_for<0, 2 * NQ>([&](auto ee) {
            constexpr int e = decltype(ee)::value, xx = e / NQ, i = e % NQ;
            static_for<0, S3 / 2>([&](auto hh) {
              constexpr int h = decltype(hh)::value, u = e * (S3 / 2) + h;
              const int k = col[xx] + 32 * (qs[xx][i] + 32 * h);
              int mk = M - k;
              if constexpr (e == 0 && h == 0) mk = self0 ? M / 2 : mk;
              s_state[(2 * u) * T + t] = si ? si[k] : 0.f;
              s_state[(2 * u + 1) * T + t] = si ? si[mk & (M - 1)] : 0.f;
            });
          });
        } else {
          const float* __restrict__ cv = reinterpret_cast<const float*>(x.carry) + ((long long)(cur.seg - 1) * x.n_clips + cur.clip) * M + t;
          static_for<0, 32>([&](auto jj) { constexpr int j = decltype(jj)::value; s_state[j * T + t] = __ldcg(cv + j * T); });
        }
      }
      if (live) {
        static_for<0, 16>([&](auto uu) {
          constexpr int u = decltype(uu)::value;
          const float a = fmaf(x.tau, s_state[(2 * u) * T + t], mk_[u]), b = fmaf(x.tau, s_state[(2 * u + 1) * T + t], mm_[u]);
          mk_[u] = bad ? 0.f : finite_or_zero(a);
          mm_[u] = bad ? 0.f : finite_or_zero(b);
          s_state[(2 * u) * T + t] = mk_[u];
          s_state[(2 * u + 1) * T + t] = mm_[u];
        });
      }
      __threadfence_block();
      turn_pass();

      if (live && p == cur.nfr - 1) {
        // last frame of a work item: hand the state to the next segment (or to the caller)
        if (cur.seg + 1 < x.segs) {
          if (x.mode != 2) {
          const long long me = (long long)cur.seg * x.n_clips + cur.clip;
          float* __restrict__ cv = reinterpret_cast<float*>(x.carry) + me * M + t;
          static_for<0, 16>([&](auto uu) {
            constexpr int u = decltype(uu)::value;
            cv[(2 * u) * T] = mk_[u];
            cv[(2 * u + 1) * T] = mm_[u];
          });
          __threadfence();
          frame_sync();
          if (t == 0) st_release_u32(x.flags + me, x.epoch);
          }
        } else if (x.state_out != nullptr) {
          float* __restrict__ so = x.state_out + (long long)cur.clip * M;
          static_for<0, 2 * NQ>([&](auto ee) {
            constexpr int e = decltype(ee)::value, xx = e / NQ, i = e % NQ;
            static_for<0, S3 / 2>([&](auto hh) {
              constexpr int h = decltype(hh)::value, u = e * (S3 / 2) + h;
              const int k = col[xx] + 32 * (qs[xx][i] + 32 * h);
              int mk = M - k;
              if constexpr (e == 0 && h == 0) mk = self0 ? M / 2 : mk;
              so[k] = mk_[u];
              so[mk & (M - 1)] = mm_[u];
            });
          });
        }
      }

      // ---- epilogue: X^ -> dB / byte / colour
      TO* __restrict__ row = out + ((long long)cur.clip * x.out_clip_rows + tf) * (long long)M;
      static_for<0, 2 * NQ>([&](auto ee) {
        constexpr int e = decltype(ee)::value, xx = e / NQ, i = e % NQ;
        static_for<0, S3 / 2>([&](auto hh) {
          constexpr int h = decltype(hh)::value, u = e * (S3 / 2) + h;
          const int k = col[xx] + 32 * (qs[xx][i] + 32 * h);
          int mk = M - k;
          if constexpr (e == 0 && h == 0) mk = self0 ? M / 2 : mk;
          if constexpr (OUT == kOutU8) {
            sb[k] = emit_smoothed<OUT>(mk_[u], ep);
            sb[mk] = emit_smoothed<OUT>(mm_[u], ep);
          } else if (store) {
            row[k] = emit_smoothed<OUT>(mk_[u], ep);
            row[mk] = emit_smoothed<OUT>(mm_[u], ep);
          }
        });
      });
      frame_sync();
      if constexpr (OUT == kOutU8) {
        if (store) {
          const uint4* s16 = reinterpret_cast<const uint4*>(sb);
          uint4* r16 = reinterpret_cast<uint4*>(row);
          r16[t] = s16[t];
          r16[T + t] = s16[T + t];
        }
        frame_sync();
      }
    }
  }
  // the last slot's final arrival has no partner waiting: complete the barrier so that the CTA can retire cleanly
  if (!first_turn && fs == 0) turn_wait();
}

}  // namespace sg
