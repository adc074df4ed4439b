// Register-resident batched real FFT for the power-of-two sizes n_fft = 256 ... 8192 that do not have
// a dedicated kernel (north_star subsystem 2: "radix-4/radix-2 in shared memory with warp-level
// butterflies for n = 256-8192").
//
// M = n_fft/2 complex points, T = M/32 threads per frame, 32 points per thread.  The network is
// radix-2 decimation in time; its log2(M) stages are run in up to three register passes
//   pass 1  stages 1-5      thread b owns z[b + T j], j < 32; compile-time twiddles
//   pass 2  stages 6-10     min(log2(M)-5, 5) stages; lane-major twiddle table W_{32*2^u}^{32 p + k_a} (all 5 stages:
//                           only the per-column base of each stage is loaded, the rest are compile-time roots times it)
//   pass 3  stages 11-12    only M = 2048, 4096; merged with the untangle: a thread takes a column pair (c, 32 - c) and
//                           the q's whose mirrors it also holds, so Z[k] and Z[M - k] meet in its registers and pass 3
//                           is neither written back nor read again
// with the frame's working set in one shared-memory tile A[row][col], row stride 33 float2, updated in
// place between passes (lanes always walk a row: conflict free but for the one lane of a half-warp that owns the
// self-mirrored columns 0 and 16).  Without pass 3, Z[k] sits at A[bitrev(k >> 5)][k & 31] after the last pass; the
// real-input untangle, |X|^2, dB and byte/colour epilogue follow as in kernel_w32.cuh.  tools/emulate_wreg.py checks the
// index algebra of both forms.
//
// n_fft 8192 on a B200 (64 clips x 60 s, byte rows): 76 M frames/s with pass 3 written back and table twiddles (L1 data
// pipe 77 % busy: 2 900 wavefronts per frame), 87 M as it stands (1 830 wavefronts, issue slots 64 % busy, FP32 pipe 48 %).
// Measured and rejected: the next frame's samples prefetched into registers behind the epilogue (128 registers do not
// hold them: spills, 66 M); FFMA2-packed butterflies on (element i, element i + 16) register pairs (17 % fewer
// instructions, but word-wide tile accesses and twiddle loads double the LSU instructions, and the dependent packed chains
// stall: 71 M).
//
// Frames of a CTA (256 threads = 256/T frames) advance together; for T <= 32 a frame lives inside one
// warp and only __syncwarp() is needed.
#pragma once
#include "common.cuh"
#include "plans.cuh"
#include "ct_math.cuh"
#include "kernel_w32.cuh"   // bfly, bfly_const, dit_stage_const

namespace sg {

constexpr int kWregThreads = 256;
constexpr int kWregStride = 33;

template <int LOG2M>
struct WregShape {
  static constexpr int M = 1 << LOG2M, N = 2 * M, T = M / 32, R = LOG2M - 5;
  static constexpr int R2 = R < 5 ? R : 5, R3 = R - R2, L2 = 1 << R3;
  static constexpr int S2 = 1 << R2, G2 = 32 / S2;   // pass 2: G2 sub-FFTs of S2 points per thread
  static constexpr int S3 = 1 << R3, G3 = 32 / S3;   // pass 3
  static constexpr int FPC = kWregThreads / T;       // frames per CTA
  static constexpr int kTileF2 = T * kWregStride;    // float2 per frame tile
  static constexpr int kFrameBytes = kTileF2 * 8 + M;   // tile + u8 staging
  static constexpr int kSmemBytes = FPC * kFrameBytes;
};

// NS radix-2 DIT stages on the NS-point sub-array v[OFF .. OFF + 2^NS), twiddle row (2^(u-1)-1+p) of a
// lane-major table (tw points at this thread's column)
template <int NS, int OFF, int ROWSTRIDE>
__device__ __forceinline__ void dit_stages_table(float2 (&v)[32], const float2* __restrict__ tw) {
  static_for<1, NS + 1>([&](auto uu) {
    constexpr int u = decltype(uu)::value, half = 1 << (u - 1), n = 1 << NS;
    static_for<0, half>([&](auto pp) {
      constexpr int p = decltype(pp)::value;
      const float2 w = __ldg(tw + (half - 1 + p) * ROWSTRIDE);
      static_for<0, n / (2 * half)>([&](auto bb) {
        constexpr int i0 = OFF + decltype(bb)::value * 2 * half + p;
        bfly(v[i0], v[i0 + half], w.x, w.y);
      });
    });
  });
}

template <int LOG2M, int OUT>
__global__ void __launch_bounds__(kWregThreads, 2)
stft_wreg_kernel(FrameGeom g, WregPlan pl, Epilogue ep, typename OutElem<OUT>::type* __restrict__ out) {
  using S = WregShape<LOG2M>;
  using TO = typename OutElem<OUT>::type;
  constexpr int M = S::M, N = S::N, T = S::T, R = S::R, R2 = S::R2, R3 = S::R3, L2 = S::L2;
  extern __shared__ float4 smem_raw[];
  const int tid = threadIdx.x, fs = tid / T, t = tid % T;
  unsigned char* fbase = reinterpret_cast<unsigned char*>(smem_raw) + fs * S::kFrameBytes;
  float2* A = reinterpret_cast<float2*>(fbase);
  unsigned char* sb = fbase + S::kTileF2 * 8;
  // a frame's T threads synchronise among themselves only: __syncwarp inside a warp, otherwise a named barrier per
  // frame slot (the other frames of the CTA keep running)
  auto frame_sync = [fs] {
    if constexpr (T <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(fs + 1), "n"(T) : "memory");
  };
  const long long groups = (g.total_frames + S::FPC - 1) / S::FPC;
  // where a frame slot's samples are: clip base, first sample, and whether the 8-byte loader can take them
  struct Src { const float* x; long long start, fc; bool live, interior; };
  auto source_of = [&](long long gi) {
    Src r;
    const long long f = gi * S::FPC + fs;
    r.live = f < g.total_frames;
    r.fc = r.live ? f : g.total_frames - 1;                 // idle slots recompute the last frame, store nothing
    const long long clip = r.fc / g.frames_per_clip, tt = r.fc - clip * g.frames_per_clip;
    r.start = g.start0 + tt * g.hop;
    r.x = g.pcm + clip * g.clip_stride;
    r.interior = r.start >= 0 && r.start + N <= g.clip_len && ((reinterpret_cast<uintptr_t>(r.x + r.start) & 7) == 0);
    return r;
  };
  for (long long gi = blockIdx.x; gi < groups; gi += gridDim.x) {
    const Src cur = source_of(gi);
    const bool live = cur.live, interior = cur.interior;
    const long long fc = cur.fc, start = cur.start;
    const float* __restrict__ x = cur.x;
    const float2* __restrict__ win2 = reinterpret_cast<const float2*>(pl.win);

    // ---- pass 1: load + window + stages 1-5
    float2 v[32];
    if (interior) {
      const float2* __restrict__ src = reinterpret_cast<const float2*>(x + start) + t;
      static_for<0, 32>([&](auto jj) {
        constexpr int j = decltype(jj)::value;
        const float2 s = __ldg(src + T * j), w = __ldg(win2 + t + T * j);
        v[bitrev(j, 5)] = make_float2(s.x * w.x, s.y * w.y);
      });
    } else if (start >= 0 && start + N <= g.clip_len) {
      // inside the clip, only misaligned (odd hops): 4-byte loads, no bounds checks
      const float* __restrict__ src = x + start + 2 * t;
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
        const float a0 = (s0 >= 0 && s0 < g.clip_len) ? __ldg(x + s0) : 0.f;
        const float a1 = (s1 >= 0 && s1 < g.clip_len) ? __ldg(x + s1) : 0.f;
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

    TO* __restrict__ row = out + fc * (long long)M;
    if constexpr (R >= 2 && R <= 4) {
      // ---- T <= 16: pass 2 is 32 columns of T-point FFTs, and the mirror of column c is column 32 - c.  Threads
      //      take the columns in mirror pairs (pair 0 = the self-mirrored columns 0 and 16), so Z[k] and Z[M - k]
      //      meet in the SAME thread: no write-back of pass 2 and no second read of the tile for the untangle.
      constexpr int P = 16 / T;                 // column pairs per thread
      static_for<0, P>([&](auto rr) {
        constexpr int r = decltype(rr)::value;
        const int pi = t + T * r;
        const int ka = pi, kb = (pi == 0) ? 16 : 32 - pi;
        static_for<0, T>([&](auto qq) {
          constexpr int q = decltype(qq)::value;
          v[(2 * r) * T + q] = A[bitrev(q, R) * kWregStride + ka];
          v[(2 * r + 1) * T + q] = A[bitrev(q, R) * kWregStride + kb];
        });
      });
      static_for<0, P>([&](auto rr) {
        constexpr int r = decltype(rr)::value;
        const int pi = t + T * r;
        const int ka = pi, kb = (pi == 0) ? 16 : 32 - pi;
        dit_stages_table<R, (2 * r) * T, 32>(v, pl.tw2 + ka);
        dit_stages_table<R, (2 * r + 1) * T, 32>(v, pl.tw2 + kb);
      });
      // now v[(2r)T + q] = Z[ka + 32 q], v[(2r+1)T + q] = Z[kb + 32 q]
      const bool bad = !(fabsf(v[0].x) <= 3.4028235e38f) || !(fabsf(v[0].y) <= 3.4028235e38f);
      auto emit2 = [&](int k, int mk, float pk, float pm) {
        pk = bad ? 0.f : pk;
        pm = bad ? 0.f : pm;
        if constexpr (OUT == kOutU8) {
          sb[k] = emit_power_finite<OUT>(pk, ep);
          sb[mk] = emit_power_finite<OUT>(pm, ep);
        } else if (live) {
          row[k] = emit_power_finite<OUT>(pk, ep);
          row[mk] = emit_power_finite<OUT>(pm, ep);
        }
      };
      auto untangle = [&](float2 zk, float2 zm, int k, float& pk, float& pm) {
        const float2 w = __ldg(pl.ut + k);
        const float ex = zk.x + zm.x, ey = zk.y - zm.y;       // 2E
        const float ox = zk.y + zm.y, oy = zm.x - zk.x;       // 2O
        const float xr = fmaf(ox, w.x, fmaf(-oy, w.y, ex));   // 2X[k]
        const float xi = fmaf(ox, w.y, fmaf(oy, w.x, ey));
        const float yr = fmaf(2.f, ex, -xr);                  // 2 conj X[M-k]
        const float yi = fmaf(2.f, ey, -xi);
        pk = fmaf(xr, xr, xi * xi);
        pm = fmaf(yr, yr, yi * yi);
      };
      static_for<0, P>([&](auto rr) {
        constexpr int r = decltype(rr)::value;
        constexpr int oa = (2 * r) * T, ob = (2 * r + 1) * T;
        const int pi = t + T * r;
        const int ka = pi, kb = (pi == 0) ? 16 : 32 - pi;
        // Lower halves of both columns lead (k < M/2, inside the W_n^k table); their mirrors are the upper halves:
        //   general pair: Z[M - (ka + 32 q)] = column kb, element T-1-q, and vice versa
        //   pair 0:       column 0 mirrors into itself (element T-q), column 16 into itself (element T-1-q)
        const bool self = (r == 0) && (t == 0);
        static_for<0, T / 2>([&](auto qq) {
          constexpr int q = decltype(qq)::value;
          float2 zma = v[ob + T - 1 - q], zmb = v[oa + T - 1 - q];
          if constexpr (r == 0) {
            const float2 sa = v[oa + ((T - q) % T)], sb2 = v[ob + T - 1 - q];
            zma = self ? sa : zma;
            zmb = self ? sb2 : zmb;
          }
          float pk, pm;
          const int k1 = ka + 32 * q;
          untangle(v[oa + q], zma, k1, pk, pm);
          int mk1 = M - k1;
          if constexpr (r == 0 && q == 0) {
            // thread 0: the mirror of k = 0 is the dropped Nyquist bin; the slot carries bin M/2 = conj Z[M/2],
            // element T/2 of column 0
            const float2 zh = v[oa + T / 2];
            pm = self ? 4.f * fmaf(zh.x, zh.x, zh.y * zh.y) : pm;
            mk1 = self ? M / 2 : mk1;
          }
          emit2(k1, mk1, pk, pm);
          const int k2 = kb + 32 * q;
          untangle(v[ob + q], zmb, k2, pk, pm);
          emit2(k2, M - k2, pk, pm);
        });
      });
    } else {
      // ---- pass 2: stages 6 .. 5+R2; sub-FFT c: column k_a, rows bitrev(q)*L2 + hi2'
      if constexpr (R2 > 0) {
        const int hi2p = (R > 5) ? (t >> 5) : 0;
        static_for<0, S::G2>([&](auto cc) {
          constexpr int c = decltype(cc)::value;
          const int ka = (R > 5) ? (t & 31) : (t + T * c);
          static_for<0, S::S2>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            v[c * S::S2 + q] = A[(bitrev(q, R2) * L2 + hi2p) * kWregStride + ka];
          });
        });
        static_for<0, S::G2>([&](auto cc) {
          constexpr int c = decltype(cc)::value;
          const int ka = (R > 5) ? (t & 31) : (t + T * c);
          if constexpr (R2 == 5) {
            // twiddles of a stage from its per-column base (row 2^(u-1) - 1) times compile-time roots: FP32-pipe work
            // instead of 31 table loads per thread (the L1 data pipe is this kernel's busiest unit)
            static_for<1, 6>([&](auto uu) {
              constexpr int u = decltype(uu)::value, half = 1 << (u - 1);
              const float2 base = __ldg(pl.tw2 + (half - 1) * 32 + ka);
              static_for<0, half>([&](auto pp) {
                constexpr int p = decltype(pp)::value;
                const float2 w = twiddle_times<p, 2 * half>(base);
                static_for<0, 16 / half>([&](auto bb) {
                  constexpr int i0 = decltype(bb)::value * 2 * half + p;
                  bfly(v[i0], v[i0 + half], w.x, w.y);
                });
              });
            });
          } else {
            dit_stages_table<R2, c * S::S2, 32>(v, pl.tw2 + ka);
          }
        });
        static_for<0, S::G2>([&](auto cc) {
          constexpr int c = decltype(cc)::value;
          const int ka = (R > 5) ? (t & 31) : (t + T * c);
          static_for<0, S::S2>([&](auto qq) {
            constexpr int q = decltype(qq)::value;
            A[(bitrev(q, R2) * L2 + hi2p) * kWregStride + ka] = v[c * S::S2 + q];
          });
        });
        frame_sync();
      }

      if constexpr (R3 > 0) {
        // ---- pass 3 (stages 11 .. 10+R3) merged with the untangle.  A sub-FFT is (column c, q): the S3 points
        //      Z[c + 32 (q + 32 h)], h < S3, read from rows bitrev5(q)*S3 + bitrev(h).  The mirror of that bin is
        //      (column 32 - c, 31 - q, S3-1 - h), so thread (pi = t & 15, qg = t >> 4) takes the column pair
        //      (pi, 32 - pi) and, in each, the q's {j, 31 - j : j in its range}: every Z[k] meets Z[M - k] in the same
        //      thread's registers -- no write-back of pass 3, no second read of the tile.  Pair 0 is the two
        //      self-mirrored columns: 16 (q <-> 31 - q) and 0 (q <-> 32 - q, with q = 0 and q = 16 mirrored onto
        //      themselves).  tools/emulate_wreg.py (merged_pass3) checks this index algebra.
        constexpr int S3 = S::S3, NP = 16 / (T / 16), NQ = 2 * NP;
        const int pi = t & 15, qg = t >> 4;
        const bool self = pi == 0, self0 = self && qg == 0;
        const int col[2] = {pi, self ? 16 : 32 - pi};
        int qs[2][NQ];
        static_for<0, NP>([&](auto ss) {
          constexpr int sl = decltype(ss)::value;
          const int j = qg * NP + sl;
          qs[0][2 * sl] = j;
          qs[0][2 * sl + 1] = self ? (j == 0 ? 16 : 32 - j) : 31 - j;
          qs[1][2 * sl] = 31 - j;      // (column 16 mirrors slot 2s onto 2s+1 in either order: keep the rows of the
          qs[1][2 * sl + 1] = j;       //  other lanes, so its reads fall into the bank the half-warp leaves free)
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
          const float2 w0 = __ldg(tw);                          // stage 11: W_2048^{32 q + c}
          if constexpr (S3 == 2) {
            bfly(v[e * 2], v[e * 2 + 1], w0.x, w0.y);
          } else {
            const float2 w1 = __ldg(tw + 32);                   // stage 12: W_4096^{32 q + c}; p = 1 is -i times it
            bfly(v[e * 4], v[e * 4 + 1], w0.x, w0.y);
            bfly(v[e * 4 + 2], v[e * 4 + 3], w0.x, w0.y);
            bfly(v[e * 4], v[e * 4 + 2], w1.x, w1.y);
            bfly(v[e * 4 + 1], v[e * 4 + 3], w1.y, -w1.x);
          }
        });
        // now v[(x NQ + i) S3 + h] = Z[col[x] + 32 (qs[x][i] + 32 h)]
        const bool bad = !(fabsf(v[0].x) <= 3.4028235e38f) || !(fabsf(v[0].y) <= 3.4028235e38f);
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
            const float pk = bad ? 0.f : fmaf(xr, xr, xi * xi);
            float pm = bad ? 0.f : fmaf(yr, yr, yi * yi);
            int mk = M - k;
            if constexpr (e == 0 && h == 0) {
              // the mirror of k = 0 is the dropped Nyquist bin; the slot carries bin M/2 = conj Z[M/2]
              const float2 zh = v[S3 / 2];
              mk = self0 ? M / 2 : mk;
              pm = self0 ? (bad ? 0.f : 4.f * fmaf(zh.x, zh.x, zh.y * zh.y)) : pm;
            }
            if constexpr (OUT == kOutU8) {
              sb[k] = emit_power_finite<OUT>(pk, ep);
              sb[mk] = emit_power_finite<OUT>(pm, ep);
            } else if (live) {
              row[k] = emit_power_finite<OUT>(pk, ep);
              row[mk] = emit_power_finite<OUT>(pm, ep);
            }
          });
        });
      } else {
      // ---- untangle + epilogue: thread t owns k = t + T i (i < 16) and the mirror bins M - k
      auto zat = [&](int k) { return A[(int)(__brev((unsigned)(k >> 5)) >> (32 - R)) * kWregStride + (k & 31)]; };
      // [SPEC] "non-finite -> 0" decided once per frame: a non-finite sample makes EVERY Z of its frame non-finite
      // (each output is a sum over all inputs and Inf*0 = NaN), so one Z tells; the per-bin work is one select.
      bool bad;
      {
        const float2 z0 = zat(t);
        bad = !(fabsf(z0.x) <= 3.4028235e38f) || !(fabsf(z0.y) <= 3.4028235e38f);
      }
      static_for<0, 16>([&](auto ii) {
        constexpr int i = decltype(ii)::value;
        const int k = t + T * i;
        const int km = (M - k) & (M - 1);
        const float2 zk = zat(k), zm = zat(km);
        const float2 w = __ldg(pl.ut + k);
        const float ex = zk.x + zm.x, ey = zk.y - zm.y;       // 2E
        const float ox = zk.y + zm.y, oy = zm.x - zk.x;       // 2O
        const float xr = fmaf(ox, w.x, fmaf(-oy, w.y, ex));   // 2X[k]
        const float xi = fmaf(ox, w.y, fmaf(oy, w.x, ey));
        const float yr = fmaf(2.f, ex, -xr);                  // 2 conj X[M-k]
        const float yi = fmaf(2.f, ey, -xi);
        const float pk = bad ? 0.f : fmaf(xr, xr, xi * xi);
        float pm = bad ? 0.f : fmaf(yr, yr, yi * yi);
        int mk = M - k;
        if constexpr (i == 0) {
          if (t == 0) {   // the mirror of k = 0 is the dropped Nyquist bin; the slot carries bin M/2 = conj Z[M/2]
            const float2 zh = zat(M / 2);
            mk = M / 2;
            pm = bad ? 0.f : 4.f * fmaf(zh.x, zh.x, zh.y * zh.y);
          }
        }
        if constexpr (OUT == kOutU8) {
          sb[k] = emit_power_finite<OUT>(pk, ep);
          sb[mk] = emit_power_finite<OUT>(pm, ep);
        } else if (live) {
          row[k] = emit_power_finite<OUT>(pk, ep);
          row[mk] = emit_power_finite<OUT>(pm, ep);
        }
      });
      }
    }
    frame_sync();
    if constexpr (OUT == kOutU8) {
      if (live) {
        const uint4* s16 = reinterpret_cast<const uint4*>(sb);
        uint4* r16 = reinterpret_cast<uint4*>(row);
        r16[t] = s16[t];
        r16[T + t] = s16[T + t];
      }
      frame_sync();
    }
  }
}

}  // namespace sg
