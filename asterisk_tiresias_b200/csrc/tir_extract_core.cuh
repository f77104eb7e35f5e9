// tir_extract_core.cuh -- per-thread phases of the fused extraction kernel.
//
// Replaces, for a batch of PCM16 clips, the hop loop of create_audio_fingerprints()
// (src/fp_handler.c:632-661): aubio_source_do -> aubio_pvoc_do -> aubio_mfcc_do ->
// 10*log10(fabs(c)) -> "%f" (src/db_ctx_handler.c:480).
//
// One CTA works on a TILE of 32 consecutive frames of one clip.  In every phase LANE = FRAME and
// WARP = ROLE (a column pair / a row pair / a filter list), so table values are warp-uniform
// broadcasts, every shared-memory access has lane stride 1 (or one 8-byte unit) and is bank-conflict
// free without padding, and the one irregular role (FFT rows 0 and N1/2) is a warp-uniform code
// path instead of per-lane selects.  Phases (a __syncthreads() between each; the bodies below are
// written per (role, lane) so that the CPU tests can run the very same code, tests/emul):
//   P0 load     33 hops of PCM16 -> shared memory (each sample is read from HBM once)
//   P1 pass1    role = column pair (win 512) / column (win 1024): s16 -> f32, hanningz window,
//               fvec_shift (an index permutation), even/odd packing, DFT_N1 over n1 in registers,
//               twiddle W_M^(n2*k1) -> exchange buffer
//   P2 pass2    role = row pair (k1, N1-k1): load both rows into registers             [sync]
//               DFT16 over n2, real untangling, sqrt(re^2+im^2) -> magnitudes (alias the exchange)
//   P3a mel     role = segment of the filter bank: one sweep over its bins, even filters in lane lo,
//               odd ones in lane hi (sequential float adds in bin order, like fmat_vecmul);
//               the coefficient warps run P4 of the PREVIOUS tile here instead
//   P3b log     every thread: clamp 2e-42 and glibc-exact log10f of ceil(40/NW) mel values [no sync after]
//   P4 dct      role = coefficient: sequential 40-term DCT row, 10*log10|c| in double, "%f"
// All float32 arithmetic of P1..P3 runs on TirP2 pairs (FADD2/FMUL2/FFMA2, tir_fp.cuh); the two
// lanes of a pair are two columns (P1), two rows (P2), two untangle slots (P2) or the even and the odd mel filter covering a bin (P3).
// The float32 FFT is "TIR-FFT" (operation order documented in DESIGN.md, and restated
// independently by the CPU oracle).
#pragma once
#include <type_traits>

#include "tir_fp.cuh"

#define TIR_MAX_FILTERS 40
#define TIR_MAX_COEFS 2
#define TIR_MAX_RUNS 64
#define TIR_MAX_W2 1024
#define TIR_MAX_WARPS 16
#define TIR_TILE 32 // frames per tile == lanes per warp

struct TirC2 { // two complex numbers: lane lo and lane hi
  TirP2 r, i;
};

// ---- kernel-parameter block (lives in the constant bank; warp-uniform reads are cheap LDCs)
// Mel sweep: the Slaney triangles of filters f and f+2 never overlap, so ONE pass over the bins
// with the even filters accumulated in lane lo and the odd ones in lane hi computes every filter,
// each bin being loaded once and broadcast to both lanes.  The live filters are cut into
// contiguous SEGMENTS, one per warp; a segment is a list of RUNS: `run_bins` bins to
// accumulate, then filter `run_emit` is complete (its sum leaves its lane, the lane restarts at 0).
struct TirMelParams {
  int16_t seg_bin0[TIR_MAX_WARPS];  // first bin of the segment
  int16_t seg_run0[TIR_MAX_WARPS];  // first run
  int16_t seg_nruns[TIR_MAX_WARPS];
  int16_t seg_woff[TIR_MAX_WARPS];  // first weight record
  int16_t run_bins[TIR_MAX_RUNS];   // (every segment's list is followed by one unused sentinel entry)
  int8_t run_emit[TIR_MAX_RUNS];    // filter id completed by the run
  uint8_t dead[TIR_MAX_FILTERS];    // filter has no non-zero weight: its log-mel value is lg_dead
  uint8_t live[TIR_MAX_FILTERS];    // ids of the n_live filters that have weights
  uint8_t live_prefix;              // live[i] == i: P3b needs no index table (dead filters lie above Nyquist)
  uint8_t pad_[7];
  // (w_even[bin], w_odd[bin]) * 2^-33 : aubio weights of the even / odd filter active at a bin of the
  // segment (0 when none); 2^-33 undoes the scaled FFT (2x) and the scaled square root (2^32 x), exactly
  float2 w2[TIR_MAX_W2];
  float dct[TIR_MAX_COEFS][TIR_MAX_FILTERS];
  float log_clamp; // (float)2e-42 : aubio VERY_SMALL_NUMBER
  float lg_dead;   // log10f(log_clamp)
  int n_filters, n_coefs, n_segs, n_live;
};

static_assert(sizeof(float4) == 16 && alignof(float4) == 16 && alignof(float2) == 8 && alignof(double2) == 16,
              "vector types must have the CUDA layout in every translation unit");
static_assert(sizeof(TirMelParams) % 16 == 0, "TirMelParams layout");

template <int WIN_>
struct TirCfg {
  static_assert(WIN_ == 512 || WIN_ == 1024, "plans: 512/256 and 1024/512 (src/fp_handler.c:33-36)");
  static constexpr int WIN = WIN_, HOP = WIN_ / 2, M = WIN_ / 2;
  static constexpr int N1 = M / 16;            // FFT_M = N1 x 16: n = 16*n1 + n2, k = k1 + N1*k2
  static constexpr int NW = N1 / 2;            // warps per CTA = roles per phase (8 / 16)
  static constexpr int T = TIR_TILE;           // frames per tile
  static constexpr int NT = NW * 32;           // threads per CTA
  static constexpr int PCH = HOP / 4 + 1;      // 8-byte units per hop chunk (+1: lane stride 2 banks mod 32)
  static constexpr int CTAS_PER_SM = WIN_ == 512 ? 2 : 1;
};

// F32IN: the PCM tile holds float samples (the down-mixed multi-channel input, already scaled by 2^15) instead of s16
template <int WIN, bool F32IN = false>
struct TirSmem {
  using C = TirCfg<WIN>;
  using PcmUnit = typename std::conditional<F32IN, float4, uint2>::type; // four samples
  static constexpr int PCM_UNITS = (C::T + 1) * C::PCH;
  static constexpr int XCH_WORDS = 2 * C::N1 * 16 * 32; // [plane re/im][k1][n2][frame]
  static_assert((C::M + 1 + 16) * 32 <= XCH_WORDS, "magnitudes (+ padded mel reads) alias the exchange buffer");
  PcmUnit pcm[PCM_UNITS];    // P1 is its only reader: the next tile streams in (cp.async) under P2..P3a
  float xch[XCH_WORDS];
  float lg[2][TIR_MAX_FILTERS * 32]; // log-mel values of this and of the previous tile (P4 lags by one)
  float4 win4[16 * C::NW];   // window pairs of the two lanes, pre-scaled by 2^-15
  float4 twp4[16 * C::NW];   // pass-1 twiddles of the two lanes
  float4 twu4[C::NW * 8];    // untangle twiddles [role][slot pair]
  double2 logtab[16];
  float2 w2[TIR_MAX_W2];     // mel sweep weights and run lists (copies of TirMelParams::w2 / run_*: a
  int run_bins[TIR_MAX_RUNS];     // broadcast LDS is much cheaper than an indexed constant-bank load
  int run_emit[TIR_MAX_RUNS];     // that misses the 2 KB constant cache)
};

#define TIR_XI(N1, plane, k1, n2, f) ((((plane) * (N1) + (k1)) * 16 + (n2)) * 32 + (f))
#define TIR_NORM_IDX(bin, f) ((bin) * 32 + (f))

// ---- complex helpers, TIR-FFT operation order, two lanes at a time -----------------------------
TIR_DEV TirC2 tir_cmul(TirC2 x, TirP2 wr, TirP2 wi) {
  TirC2 o;
  o.r = tir_pfma(tir_pneg(x.i), wi, tir_pmul(x.r, wr));
  o.i = tir_pfma(x.i, wr, tir_pmul(x.r, wi));
  return o;
}

TIR_DEV void tir_dft4(TirC2 a0, TirC2 a1, TirC2 a2, TirC2 a3, TirC2 &A0, TirC2 &A1, TirC2 &A2, TirC2 &A3) {
  TirC2 s0 = {tir_padd(a0.r, a2.r), tir_padd(a0.i, a2.i)}, d0 = {tir_psub(a0.r, a2.r), tir_psub(a0.i, a2.i)};
  TirC2 s1 = {tir_padd(a1.r, a3.r), tir_padd(a1.i, a3.i)}, d1 = {tir_psub(a1.r, a3.r), tir_psub(a1.i, a3.i)};
  A0.r = tir_padd(s0.r, s1.r), A0.i = tir_padd(s0.i, s1.i);
  A2.r = tir_psub(s0.r, s1.r), A2.i = tir_psub(s0.i, s1.i);
  A1.r = tir_padd(d0.r, d1.i), A1.i = tir_psub(d0.i, d1.r);
  A3.r = tir_psub(d0.r, d1.i), A3.i = tir_padd(d0.i, d1.r);
}

#define TIR_C1 0.92387953251128674f
#define TIR_S1 0.38268343236508977f
#define TIR_H 0.70710678118654752f

// in-place 16-point DFT of both lanes, x[n] -> X[k], natural order in and out
TIR_DEV void tir_dft16(TirC2 (&x)[16], TirP2 nz) {
  TirC2 y[4][4];
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) tir_dft4(x[n2], x[n2 + 4], x[n2 + 8], x[n2 + 12], y[n2][0], y[n2][1], y[n2][2], y[n2][3]);
  TirC2 v;
  const TirP2 h = tir_pbc(TIR_H);
  // W16^(n2*k1)
  y[1][1] = tir_cmul(y[1][1], tir_pbc(TIR_C1), tir_pbc(-TIR_S1));
  v = y[1][2], y[1][2].r = tir_pmulx(tir_padd(v.r, v.i), h, nz), y[1][2].i = tir_pmulx(tir_psub(v.i, v.r), h, nz);
  y[1][3] = tir_cmul(y[1][3], tir_pbc(TIR_S1), tir_pbc(-TIR_C1));
  v = y[2][1], y[2][1].r = tir_pmulx(tir_padd(v.r, v.i), h, nz), y[2][1].i = tir_pmulx(tir_psub(v.i, v.r), h, nz);
  v = y[2][2], y[2][2].r = v.i, y[2][2].i = tir_pneg(v.r);
  v = y[2][3], y[2][3].r = tir_pmulx(tir_psub(v.i, v.r), h, nz), y[2][3].i = tir_pneg(tir_pmulx(tir_padd(v.r, v.i), h, nz));
  y[3][1] = tir_cmul(y[3][1], tir_pbc(TIR_S1), tir_pbc(-TIR_C1));
  v = y[3][2], y[3][2].r = tir_pmulx(tir_psub(v.i, v.r), h, nz), y[3][2].i = tir_pneg(tir_pmulx(tir_padd(v.r, v.i), h, nz));
  y[3][3] = tir_cmul(y[3][3], tir_pbc(-TIR_C1), tir_pbc(TIR_S1));
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) tir_dft4(y[0][k1], y[1][k1], y[2][k1], y[3][k1], x[k1], x[k1 + 4], x[k1 + 8], x[k1 + 12]);
}

TIR_DEV float tir_s16lo(uint32_t w) { return (float)(int16_t)(w & 0xffffu); }
TIR_DEV float tir_s16hi(uint32_t w) { return (float)(int16_t)(w >> 16); }

// ---- P1, win 512 --------------------------------------------------------------------------------
// role w = columns n2 = 2w (lane lo) and 2w+1 (lane hi); lane f = frame of the tile.
// z[n], n = 16*n1 + n2, is the complex point (x[2n], x[2n+1]) of the windowed, fvec_shift'ed frame:
// sample index (2n + WIN/2) mod WIN, i.e. hop chunk f + 1 - (n1 >> 3), 32-bit word 16*(n1 & 7) + n2.
// the four samples of unit (chunk, 8*(n1&7)+w) as (re lo, re hi, im lo, im hi) = (s0, s2, s1, s3)
#if defined(__CUDACC__) && !defined(TIR_I2F_CONVERT)
// s16 -> f32 without the conversion unit (I2F runs at a quarter of the ALU rate and its latency sat in front of every
// window product): bias both halves of a word to unsigned (one XOR), splice each half under the exponent of 2^23 (one
// PRMT), and take 2^23 + 2^15 off again with ONE packed subtraction for two samples -- all exact.
#define TIR_S16_MAGIC 8421376.0f // 2^23 + 2^15
TIR_DEV void tir_s16x2(uint32_t wa, uint32_t wb, TirP2 &lo, TirP2 &hi) { // (lo half of wa, lo half of wb), (hi, hi)
  const uint32_t a = wa ^ 0x80008000u, b = wb ^ 0x80008000u;
  const TirP2 k = tir_pbc(TIR_S16_MAGIC);
  lo = tir_psub(tir_pmk(__uint_as_float(__byte_perm(a, 0x4B00u, 0x5410)), __uint_as_float(__byte_perm(b, 0x4B00u, 0x5410))), k);
  hi = tir_psub(tir_pmk(__uint_as_float(__byte_perm(a, 0x4B00u, 0x5432)), __uint_as_float(__byte_perm(b, 0x4B00u, 0x5432))), k);
}
#else
TIR_DEV void tir_s16x2(uint32_t wa, uint32_t wb, TirP2 &lo, TirP2 &hi) {
  lo = tir_pmk(tir_s16lo(wa), tir_s16lo(wb)), hi = tir_pmk(tir_s16hi(wa), tir_s16hi(wb));
}
#endif
TIR_DEV void tir_unit4(const uint2 *pcm, int idx, TirP2 &re, TirP2 &im) {
  const uint2 u = pcm[idx];
  tir_s16x2(u.x, u.y, re, im);
}
TIR_DEV void tir_unit4(const float4 *pcm, int idx, TirP2 &re, TirP2 &im) {
  const float4 u = pcm[idx];
  re = tir_pmk(u.x, u.z), im = tir_pmk(u.y, u.w);
}
template <class SM, class U>
TIR_DEV void tir_pass1_512(SM &sm, const U *pcm, int w, int f, TirP2 nz) {
  using C = TirCfg<512>;
  TirC2 x[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++) {
    TirP2 re, im;
    tir_unit4(pcm, (f + 1 - (n1 >> 3)) * C::PCH + 8 * (n1 & 7) + w, re, im);
    const float4 wv = sm.win4[n1 * C::NW + w];
    x[n1].r = tir_pmulx(re, tir_pmk(wv.x, wv.y), nz);
    x[n1].i = tir_pmulx(im, tir_pmk(wv.z, wv.w), nz);
  }
  tir_dft16(x, nz);
  float *xr = sm.xch + TIR_XI(C::N1, 0, 0, 2 * w, f), *xi = sm.xch + TIR_XI(C::N1, 1, 0, 2 * w, f);
#pragma unroll
  for (int k1 = 0; k1 < 16; k1++) {
    TirC2 o = x[k1];
    if (k1 > 0) { // row 0 is stored as it is; the table holds (1, 0) where n2*k1 == 0
      const float4 tw = sm.twp4[k1 * C::NW + w];
      o = tir_cmul(o, tir_pmk(tw.x, tw.y), tir_pmk(tw.z, tw.w));
    }
    xr[k1 * 512] = o.r.lo, xr[k1 * 512 + 32] = o.r.hi;
    xi[k1 * 512] = o.i.lo, xi[k1 * 512 + 32] = o.i.hi;
  }
}

// ---- P1, win 1024 -------------------------------------------------------------------------------
// role c = column n2 = c; the DFT32 over n1 is TIR-FFT's "DFT16 of the even and of the odd inputs,
// then X[k] = E + W32^k O, X[k+16] = E - W32^k O": lane lo = even n1 = 2m, lane hi = odd n1 = 2m+1.
// Sample words: hop chunk f + 1 - (n1 >> 4), word 16*(n1 & 15) + c.
#define TIR_W32R(k)                                                                                             \
  ((k) == 1 ? 0x1.f6297cp-1f : (k) == 2 ? 0x1.d906bcp-1f : (k) == 3 ? 0x1.a9b662p-1f : (k) == 4 ? 0x1.6a09e6p-1f \
   : (k) == 5 ? 0x1.1c73b4p-1f : (k) == 6 ? 0x1.87de2ap-2f : (k) == 7 ? 0x1.8f8b84p-3f                          \
   : (k) == 9 ? -0x1.8f8b84p-3f : (k) == 10 ? -0x1.87de2ap-2f : (k) == 11 ? -0x1.1c73b4p-1f                     \
   : (k) == 12 ? -0x1.6a09e6p-1f : (k) == 13 ? -0x1.a9b662p-1f : (k) == 14 ? -0x1.d906bcp-1f : -0x1.f6297cp-1f)
#define TIR_W32I(k)                                                                                               \
  ((k) == 1 ? -0x1.8f8b84p-3f : (k) == 2 ? -0x1.87de2ap-2f : (k) == 3 ? -0x1.1c73b4p-1f : (k) == 4 ? -0x1.6a09e6p-1f \
   : (k) == 5 ? -0x1.a9b662p-1f : (k) == 6 ? -0x1.d906bcp-1f : (k) == 7 ? -0x1.f6297cp-1f                          \
   : (k) == 9 ? -0x1.f6297cp-1f : (k) == 10 ? -0x1.d906bcp-1f : (k) == 11 ? -0x1.a9b662p-1f                        \
   : (k) == 12 ? -0x1.6a09e6p-1f : (k) == 13 ? -0x1.1c73b4p-1f : (k) == 14 ? -0x1.87de2ap-2f : -0x1.8f8b84p-3f)

// the two sample pairs (words idx and idx + 16 of the tile: an even and an odd n1) as (re lo, re hi), (im lo, im hi)
TIR_DEV void tir_pair2(const uint2 *pcm, int idx, TirP2 &re, TirP2 &im) {
  const uint32_t *p = reinterpret_cast<const uint32_t *>(pcm) + idx;
  const uint32_t ue = p[0], uo = p[16];
  tir_s16x2(ue, uo, re, im);
}
TIR_DEV void tir_pair2(const float4 *pcm, int idx, TirP2 &re, TirP2 &im) {
  const float2 *p = reinterpret_cast<const float2 *>(pcm) + idx;
  const float2 ue = p[0], uo = p[16];
  re = tir_pmk(ue.x, uo.x), im = tir_pmk(ue.y, uo.y);
}
template <class SM, class U>
TIR_DEV void tir_pass1_1024(SM &sm, const U *pcm, int c, int f, TirP2 nz) {
  using C = TirCfg<1024>;
  TirC2 x[16];
#pragma unroll
  for (int m = 0; m < 16; m++) {
    TirP2 re, im;
    tir_pair2(pcm, (f + 1 - (m >> 3)) * (2 * C::PCH) + 32 * (m & 7) + c, re, im);
    const float4 wv = sm.win4[m * C::NW + c];
    x[m].r = tir_pmulx(re, tir_pmk(wv.x, wv.y), nz);
    x[m].i = tir_pmulx(im, tir_pmk(wv.z, wv.w), nz);
  }
  tir_dft16(x, nz); // lane lo: E[k], lane hi: O[k]
  float *xr = sm.xch + TIR_XI(C::N1, 0, 0, c, f), *xi = sm.xch + TIR_XI(C::N1, 1, 0, c, f);
#pragma unroll
  for (int k = 0; k < 16; k++) {
    float tr, ti;
    if (k == 0) {
      tr = x[k].r.hi, ti = x[k].i.hi;
    } else if (k == 8) {
      tr = x[k].i.hi, ti = -x[k].r.hi;
    } else {
      tr = TIR_FFMA(-x[k].i.hi, TIR_W32I(k), TIR_FMUL(x[k].r.hi, TIR_W32R(k)));
      ti = TIR_FFMA(x[k].i.hi, TIR_W32R(k), TIR_FMUL(x[k].r.hi, TIR_W32I(k)));
    }
    // rows k1 = k (lane lo) and k + 16 (lane hi)
    TirC2 o;
    o.r = tir_pmk(TIR_FADD(x[k].r.lo, tr), TIR_FSUB(x[k].r.lo, tr));
    o.i = tir_pmk(TIR_FADD(x[k].i.lo, ti), TIR_FSUB(x[k].i.lo, ti));
    const float4 tw = sm.twp4[k * C::NW + c]; // (1, 0) where c*k1 == 0
    TirC2 q = tir_cmul(o, tir_pmk(tw.x, tw.y), tir_pmk(tw.z, tw.w));
    if (k == 0) q.r.lo = o.r.lo, q.i.lo = o.i.lo; // row 0 is stored as it is
    xr[k * 512] = q.r.lo, xr[(k + 16) * 512] = q.r.hi;
    xi[k * 512] = q.i.lo, xi[(k + 16) * 512] = q.i.hi;
  }
}

// ---- P2 ---------------------------------------------------------------------------------------
// role t: rows kA = t (lane lo) and kB = N1 - t (lane hi); role 0: rows 0 and N1/2.
struct TirPass2Regs {
  TirC2 X[16];
};

template <int WIN, class SM>
TIR_DEV void tir_pass2_load(const SM &sm, int t, int f, TirPass2Regs &rg) {
  using C = TirCfg<WIN>;
  const int kA = t, kB = t ? C::N1 - t : C::N1 / 2;
  const float *ar = sm.xch + TIR_XI(C::N1, 0, kA, 0, f), *ai = sm.xch + TIR_XI(C::N1, 1, kA, 0, f);
  const float *br = sm.xch + TIR_XI(C::N1, 0, kB, 0, f), *bi = sm.xch + TIR_XI(C::N1, 1, kB, 0, f);
#pragma unroll
  for (int n2 = 0; n2 < 16; n2++) {
    rg.X[n2].r = tir_pmk(ar[n2 * 32], br[n2 * 32]);
    rg.X[n2].i = tir_pmk(ai[n2 * 32], bi[n2 * 32]);
  }
}

// Two untangle slots at a time.  With Z[k] = U = (a, b), Z[M-k] = V = (c, d), k <= M/2:
//   E2 = (a+c, b-d), O2 = (b+d, c-a), T = W_{2M}^k O2, 2X[k] = E2 + T, 2X[M-k] = conj(E2 - T)
// -> mk = 2^32 |2X[k]|, mmk = 2^32 |2X[M-k]| (the 2^33 is folded, exactly, into the mel weights).
// U and V hold the two slots in their lanes.
TIR_DEV void tir_untangle_mag2(TirC2 U, TirC2 V, float4 w, TirP2 nz, TirP2 &mk, TirP2 &mmk) {
  TirC2 E2, O2;
  E2.r = tir_padd(U.r, V.r), E2.i = tir_psub(U.i, V.i);
  O2.r = tir_padd(U.i, V.i), O2.i = tir_psub(V.r, U.r);
  const TirC2 Tt = tir_cmul(O2, tir_pmk(w.x, w.y), tir_pmk(w.z, w.w));
  const TirP2 pr = tir_padd(E2.r, Tt.r), pi = tir_padd(E2.i, Tt.i);
  const TirP2 qr = tir_psub(E2.r, Tt.r), qi = tir_psub(E2.i, Tt.i);
  mk = tir_psqrt_scaled64(tir_padd(tir_pmulx(pr, pr, nz), tir_pmulx(pi, pi, nz)));
  mmk = tir_psqrt_scaled64(tir_padd(tir_pmulx(qr, qr, nz), tir_pmulx(qi, qi, nz)));
}

// T0 = (role == 0): warp-uniform, so the irregular pairing of rows 0 and N1/2 costs no selects.
//   role t >= 1, slot pair s:  lane lo  k = t + N1 s         U = A[s]    V = B[15-s]
//                              lane hi  k = (N1-t) + N1 s    U = B[s]    V = A[15-s]
//     i.e. U = X[s] as it is and V = X[15-s] with its lanes SWAPPED -- a free operand form of the
//     packed instructions (R.F32x2.LO_HI), so the row-mixing additions are packed too;
//   role 0:                    lane lo  k = N1 (s+1)         U = A[s+1]  V = A[15-s]   (row 0)
//                              lane hi  k = N1/2 + N1 s      U = B[s]    V = B[15-s]   (row N1/2)
//   (role 0, s = 7, lane lo is k = M/2 paired with itself; bins 0 and M are never produced: no mel
//   filter has weight there)
template <int WIN, bool T0, class SM>
TIR_DEV void tir_pass2_compute(SM &sm, int t, int f, TirPass2Regs &rg, TirP2 nz) {
  using C = TirCfg<WIN>;
  tir_dft16(rg.X, nz); // lane lo: Z[kA + N1 k2], lane hi: Z[kB + N1 k2]
  float *mags = sm.xch + f;
  const int klo0 = T0 ? C::N1 : t, khi0 = T0 ? C::N1 / 2 : C::N1 - t;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    const TirC2 &Xs = rg.X[s], &Xr = rg.X[15 - s], &Xn = rg.X[s + 1];
    TirC2 U, V;
    if (T0) {
      U.r = tir_pmk(Xn.r.lo, Xs.r.hi), U.i = tir_pmk(Xn.i.lo, Xs.i.hi);
      V = Xr;
    } else {
      U = Xs;
      V.r = tir_pmk(Xr.r.hi, Xr.r.lo), V.i = tir_pmk(Xr.i.hi, Xr.i.lo);
    }
    TirP2 mk, mmk;
    tir_untangle_mag2(U, V, sm.twu4[t * 8 + s], nz, mk, mmk);
    const int klo = klo0 + C::N1 * s, khi = khi0 + C::N1 * s;
    mags[klo * 32] = mk.lo;
    if (!(T0 && s == 7)) mags[(C::M - klo) * 32] = mmk.lo;
    mags[khi * 32] = mk.hi;
    mags[(C::M - khi) * 32] = mmk.hi;
  }
}

// ---- P3a: mel sweep -----------------------------------------------------------------------------
// role = segment `seg` (warp-uniform), lane f = frame.  fmat_vecmul order: every filter's sum runs
// over its bins in ascending order from 0.f; bins where a lane's weight is 0 add +0 (x is finite: a
// magnitude).  Weights and run lists are read from shared memory (w2, run_*); the parameters of the
// next run are fetched while the current one accumulates.  Writes the raw sums to lg.
// `NP` pairs of bins starting `k` pairs before the end of the run: the addresses are run-end pointers plus
// immediates, all loads of the block are issued before its sums (the sums themselves are one dependent chain).
template <int NP>
TIR_DEV TirP2 tir_mel_block(TirP2 acc, const float *me, const float4 *wp, int k, TirP2 nz) {
  float4 wv[NP];
  float ma[NP], mb[NP];
#pragma unroll
  for (int j = 0; j < NP; j++) wv[j] = wp[j - k], ma[j] = me[64 * (j - k)], mb[j] = me[64 * (j - k) + 32];
#pragma unroll
  for (int j = 0; j < NP; j++) {
    acc = tir_padd(acc, tir_pmulx(tir_pbc(ma[j]), tir_pmk(wv[j].x, wv[j].y), nz));
    acc = tir_padd(acc, tir_pmulx(tir_pbc(mb[j]), tir_pmk(wv[j].z, wv[j].w), nz));
  }
  return acc;
}
TIR_DEV void tir_mel_sweep(const float *norm, float *lg, const TirMelParams &mp, const float2 *w2, const int *run_bins,
                           const int *run_emit, int seg, int f, TirP2 nz) {
  const int r0 = mp.seg_run0[seg], nr = mp.seg_nruns[seg];
  if (nr == 0) return;
  const float *m = norm + TIR_NORM_IDX(mp.seg_bin0[seg], f);
  const float4 *wp = reinterpret_cast<const float4 *>(w2 + mp.seg_woff[seg]); // two bins' records per load
  TirP2 acc = tir_pbc(0.f);
  int n = run_bins[r0], fe = run_emit[r0];
  for (int r = 0; r < nr; r++) {
    const int n_next = run_bins[r0 + r + 1], fe_next = run_emit[r0 + r + 1]; // the list ends with a sentinel
    // bins in pairs; an odd run's last pair holds a zero-weight record (tir_tables.cpp) and sweeps the
    // next bin with it: acc + (+-0) is exact, and the magnitude buffer is finite everywhere.
    // Runs are short (a mel triangle's rising or falling edge: 2..30 pairs), so a counted loop spends as many
    // instructions on its compare / branch / address updates as on the sums.  The pairs are addressed from the
    // run's END: first the k mod 4 leading pairs, then blocks of four pairs (ascending bins -- the very order of
    // the loop this replaces), every block with immediate addresses and its loads in flight together.  The code
    // is the same for every warp: per-warp straight-line code lost to the instruction cache.
    int k = (n + 1) >> 1;
    const float *me = m + 64 * k; // end of the run's pairs (one bin past the run when n is odd)
    m += 32 * n, wp += k;         // the next run starts at bin + n, after the k records
    switch (k & 3) {
      case 3: acc = tir_mel_block<3>(acc, me + 0, wp, k, nz); break;
      case 2: acc = tir_mel_block<2>(acc, me + 0, wp, k, nz); break;
      case 1: acc = tir_mel_block<1>(acc, me + 0, wp, k, nz); break;
      default: break;
    }
    for (k &= ~3; k > 0; k -= 4) acc = tir_mel_block<4>(acc, me, wp, k, nz);
    if (fe & 1) lg[fe * 32 + f] = acc.hi, acc.hi = 0.f;
    else lg[fe * 32 + f] = acc.lo, acc.lo = 0.f;
    n = n_next, fe = fe_next;
  }
}

// ---- P3b: clamp + log10f, in place ----------------------------------------------------------------
// role w takes live filters w, w + NW, ...: the same number of logarithms for every thread, written
// as straight-line code over all of them so that their (long, double precision) dependency chains
// interleave.  Filters without weights keep the constant lg_dead written once at kernel start.
template <int NW>
TIR_DEV void tir_log_phase(float *lg, const double2 *logtab, const TirMelParams &mp, int w, int f) {
  constexpr int Q = (TIR_MAX_FILTERS + NW - 1) / NW;
  float v[Q];
  int at[Q];
#pragma unroll
  for (int q = 0; q < Q; q++) {
    const int i = w + NW * q;
    at[q] = i < mp.n_live ? mp.live[i] * 32 + f : -1;
    v[q] = at[q] >= 0 ? lg[at[q]] : 1.f;
  }
#pragma unroll
  for (int q = 0; q < Q; q++) v[q] = tir_log10f_glibc(fmaxf(v[q], mp.log_clamp), logtab); // sums are >= +0, never NaN
#pragma unroll
  for (int q = 0; q < Q; q++)
    if (at[q] >= 0) lg[at[q]] = v[q];
}

// the same when the live filters are 0..n_live-1 (every plan of the reference: only filters above
// Nyquist are empty): the addresses are lane + immediate, no constant-bank index loads
template <int NW>
TIR_DEV void tir_log_phase_prefix(float *lg, const double2 *logtab, float log_clamp, int n_live, int w, int f) {
  constexpr int Q = (TIR_MAX_FILTERS + NW - 1) / NW;
  float *p = lg + w * 32 + f;
  float v[Q];
  bool on[Q];
#pragma unroll
  for (int q = 0; q < Q; q++) {
    on[q] = w + NW * q < n_live;
    v[q] = on[q] ? p[q * NW * 32] : 1.f;
  }
#pragma unroll
  for (int q = 0; q < Q; q++) v[q] = tir_log10f_glibc(fmaxf(v[q], log_clamp), logtab);
#pragma unroll
  for (int q = 0; q < Q; q++)
    if (on[q]) p[q * NW * 32] = v[q];
}

// ---- P4 ---------------------------------------------------------------------------------------
TIR_DEV void tir_dct_phase(const float *lg, const TirMelParams &mp, int j, int f, float &c, int32_t &vq) {
  float acc = 0.f;
#pragma unroll 8
  for (int i = 0; i < mp.n_filters; i++) acc = TIR_FADD(acc, TIR_FMUL(lg[i * 32 + f], mp.dct[j][i]));
  c = acc;
  vq = tir_quantize_micro(tir_coef_to_y(acc));
}
