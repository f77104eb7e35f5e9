// tir_extract_core.cuh -- per-thread phases of the fused extraction kernel.
//
// Replaces, for a batch of PCM16 clips, the hop loop of create_audio_fingerprints()
// (src/fp_handler.c:632-661): aubio_source_do -> aubio_pvoc_do -> aubio_mfcc_do ->
// 10*log10(fabs(c)) -> "%f" (src/db_ctx_handler.c:480).
//
// One CTA works on a TILE of T consecutive frames of one clip.  Phases (a __syncthreads()
// between each; the phase bodies below are written per thread so that the CPU tests can run the
// very same code thread by thread, tests/emul):
//   P0 load     (T+1) hops of PCM16 -> shared memory (each sample is read from HBM once)
//   P1 pass1    TPF threads per frame: s16 -> f32, hanningz window, fvec_shift (an index
//               permutation), even/odd packing, DFT_N1 over n1 in registers, twiddle, -> exchange
//   P2 pass2    load two rows (k1, N1-k1) of the exchange buffer into registers         [sync]
//               DFT16 x2, real untangling in registers, sqrt(re^2+im^2) -> magnitudes (aliases
//               the exchange buffer)
//   P3 mel      lane = frame, warp = filter list: banded Slaney matvec (sequential float adds in
//               bin order, like fmat_vecmul), clamp 2e-42, glibc-exact log10f
//   P4 dct      lane = frame, warp = coefficient: sequential 40-term DCT row, 10*log10|c| in
//               double, "%f" quantisation, store
// The float32 FFT is "TIR-FFT" (operation order documented in DESIGN.md, and restated
// independently by the CPU oracle).
#pragma once
#include "tir_fp.cuh"

#define TIR_MAX_FILTERS 40
#define TIR_MAX_COEFS 2
#define TIR_MAX_NNZ 2048
#define TIR_MEL_WARPS 8

struct TirCpx {
  float r, i;
};

// ---- kernel-parameter block (lives in the constant bank; warp-uniform reads are free operands)
struct TirMelParams {
  int16_t start[TIR_MAX_FILTERS];   // first bin with non-zero weight
  int16_t len[TIR_MAX_FILTERS];     // number of bins
  int16_t woff[TIR_MAX_FILTERS];    // offset into w[]
  uint8_t warp_nf[TIR_MEL_WARPS];   // filters handled by mel warp w
  uint8_t warp_filters[TIR_MEL_WARPS][TIR_MAX_FILTERS];
  float w[TIR_MAX_NNZ];             // 0.5 * aubio filter weight (the 0.5 of the scaled FFT)
  float dct[TIR_MAX_COEFS][TIR_MAX_FILTERS];
  float log_clamp;                  // (float)2e-42 : aubio VERY_SMALL_NUMBER
  int n_filters, n_coefs;
};

template <int WIN>
struct TirCfg;

template <>
struct TirCfg<512> {
  static constexpr int WIN = 512, HOP = 256, M = 256, N1 = 16, TPF = 8;
  static constexpr int T = 32;                 // frames per tile
  static constexpr int NT = T * TPF;           // threads per CTA (256)
  static constexpr int PCM_STRIDE_W = 136;     // 32-bit words per hop chunk (128 + 8 pad)
  static constexpr int XCH_ROW = 17;           // float2 per k1 row (16 + 1 pad)
  static constexpr int XCH_FRAME_W = 560;      // words per frame (N1*XCH_ROW*2 = 544, +16)
};

template <int WIN>
struct TirSmem {
  using C = TirCfg<WIN>;
  static constexpr int PCM_WORDS = (C::T + 1) * C::PCM_STRIDE_W;
  static constexpr int XCH_WORDS = C::T * C::XCH_FRAME_W;
  static constexpr int NORM_WORDS = (C::M + 1) * 32;
  static_assert(NORM_WORDS <= XCH_WORDS, "magnitudes alias the exchange buffer");
  uint32_t pcm[PCM_WORDS];
  float xch[XCH_WORDS];
  float2 win2[C::M];            // window pairs in z[] order, pre-scaled by 2^-15
  float2 tw_pass[C::N1 * 16];   // [k1][n2]  W_M^(n2*k1)
  float2 tw_unt[16 * C::TPF];   // [slot][t] W_{2M}^k
  double2 logtab[16];
  float lg[TIR_MAX_FILTERS * 32];
};

// ---- complex helpers, TIR-FFT operation order -------------------------------------------------
TIR_DEV TirCpx tir_cmul(TirCpx x, float wr, float wi) {
  TirCpx o;
  o.r = TIR_FFMA(-x.i, wi, TIR_FMUL(x.r, wr));
  o.i = TIR_FFMA(x.i, wr, TIR_FMUL(x.r, wi));
  return o;
}

TIR_DEV void tir_dft4(TirCpx a0, TirCpx a1, TirCpx a2, TirCpx a3, TirCpx &A0, TirCpx &A1, TirCpx &A2,
                      TirCpx &A3) {
  TirCpx s0 = {TIR_FADD(a0.r, a2.r), TIR_FADD(a0.i, a2.i)}, d0 = {TIR_FSUB(a0.r, a2.r), TIR_FSUB(a0.i, a2.i)};
  TirCpx s1 = {TIR_FADD(a1.r, a3.r), TIR_FADD(a1.i, a3.i)}, d1 = {TIR_FSUB(a1.r, a3.r), TIR_FSUB(a1.i, a3.i)};
  A0.r = TIR_FADD(s0.r, s1.r), A0.i = TIR_FADD(s0.i, s1.i);
  A2.r = TIR_FSUB(s0.r, s1.r), A2.i = TIR_FSUB(s0.i, s1.i);
  A1.r = TIR_FADD(d0.r, d1.i), A1.i = TIR_FSUB(d0.i, d1.r);
  A3.r = TIR_FSUB(d0.r, d1.i), A3.i = TIR_FADD(d0.i, d1.r);
}

#define TIR_C1 0.92387953251128674f
#define TIR_S1 0.38268343236508977f
#define TIR_H 0.70710678118654752f

// in-place 16-point DFT, x[n] -> X[k], natural order in and out
TIR_DEV void tir_dft16(TirCpx (&x)[16]) {
  TirCpx y[4][4];
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) tir_dft4(x[n2], x[n2 + 4], x[n2 + 8], x[n2 + 12], y[n2][0], y[n2][1], y[n2][2], y[n2][3]);
  TirCpx v;
  // W16^(n2*k1)
  y[1][1] = tir_cmul(y[1][1], TIR_C1, -TIR_S1);
  v = y[1][2], y[1][2].r = TIR_FMUL(TIR_FADD(v.r, v.i), TIR_H), y[1][2].i = TIR_FMUL(TIR_FSUB(v.i, v.r), TIR_H);
  y[1][3] = tir_cmul(y[1][3], TIR_S1, -TIR_C1);
  v = y[2][1], y[2][1].r = TIR_FMUL(TIR_FADD(v.r, v.i), TIR_H), y[2][1].i = TIR_FMUL(TIR_FSUB(v.i, v.r), TIR_H);
  v = y[2][2], y[2][2].r = v.i, y[2][2].i = -v.r;
  v = y[2][3], y[2][3].r = TIR_FMUL(TIR_FSUB(v.i, v.r), TIR_H), y[2][3].i = -TIR_FMUL(TIR_FADD(v.r, v.i), TIR_H);
  y[3][1] = tir_cmul(y[3][1], TIR_S1, -TIR_C1);
  v = y[3][2], y[3][2].r = TIR_FMUL(TIR_FSUB(v.i, v.r), TIR_H), y[3][2].i = -TIR_FMUL(TIR_FADD(v.r, v.i), TIR_H);
  y[3][3] = tir_cmul(y[3][3], -TIR_C1, TIR_S1);
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) tir_dft4(y[0][k1], y[1][k1], y[2][k1], y[3][k1], x[k1], x[k1 + 4], x[k1 + 8], x[k1 + 12]);
}

// ---- thread <-> work mapping -------------------------------------------------------------------
// warp w, lane l: frame slot fl = 4*w + (l>>3) for TPF=8 (consecutive frames in a warp so that the
// PCM, exchange and magnitude accesses below are bank-conflict free), t = l & (TPF-1).
template <int WIN>
TIR_DEV int tir_frame_of(int tid) { return tid / TirCfg<WIN>::TPF; }
template <int WIN>
TIR_DEV int tir_t_of(int tid) { return tid % TirCfg<WIN>::TPF; }

// magnitude store slot: bin-major, frame rotated so that writes (8 threads x 4 frames) and reads
// (32 frames, one bin) both hit 32 distinct banks
TIR_DEV int tir_norm_idx(int bin, int fl) {
  int rot = ((fl & 3) << 3) | (fl >> 2);
  return bin * 32 + ((rot + bin) & 31);
}

// ---- P1 ---------------------------------------------------------------------------------------
// `frames_valid`: frames of this tile that exist; threads of other frames still run (zeros).
template <int WIN>
TIR_DEV void tir_pass1(TirSmem<WIN> &sm, int tid) {
  using C = TirCfg<WIN>;
  const int fl = tir_frame_of<WIN>(tid), t = tir_t_of<WIN>(tid);
  float2 *xch = reinterpret_cast<float2 *>(sm.xch) + (size_t)fl * (C::XCH_FRAME_W / 2);
  static_assert(C::N1 == 16, "pass1 is written for N1 == 16 (win 512)");
#pragma unroll 1
  for (int j = 0; j < 2; j++) {
    const int n2 = t + 8 * j;
    TirCpx x[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) {
      // z[n], n = 16*n1 + n2, after fvec_shift: sample index (2n + WIN/2) mod WIN
      const int chunk = fl + 1 - (n1 >> 3);
      const uint32_t word = sm.pcm[chunk * C::PCM_STRIDE_W + 16 * (n1 & 7) + n2];
      const float2 w = sm.win2[16 * n1 + n2];
      x[n1].r = TIR_FMUL((float)(int16_t)(word & 0xffffu), w.x);
      x[n1].i = TIR_FMUL((float)(int16_t)(word >> 16), w.y);
    }
    tir_dft16(x);
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) {
      TirCpx o = x[k1];
      if (k1 > 0) {
        const float2 w = sm.tw_pass[k1 * 16 + n2];
        o = tir_cmul(o, w.x, w.y);
      }
      float2 st;
      st.x = o.r, st.y = o.i;
      xch[k1 * C::XCH_ROW + n2] = st;
    }
  }
}

// ---- P2 ---------------------------------------------------------------------------------------
struct TirPass2Regs {
  TirCpx A[16], B[16];
};

template <int WIN>
TIR_DEV void tir_pass2_load(const TirSmem<WIN> &sm, int tid, TirPass2Regs &rg) {
  using C = TirCfg<WIN>;
  const int fl = tir_frame_of<WIN>(tid), t = tir_t_of<WIN>(tid);
  const float2 *xch = reinterpret_cast<const float2 *>(sm.xch) + (size_t)fl * (C::XCH_FRAME_W / 2);
  const int kA = t, kB = t ? C::N1 - t : C::N1 / 2;
#pragma unroll
  for (int n2 = 0; n2 < 16; n2++) {
    float2 a = xch[kA * C::XCH_ROW + n2], b = xch[kB * C::XCH_ROW + n2];
    rg.A[n2].r = a.x, rg.A[n2].i = a.y;
    rg.B[n2].r = b.x, rg.B[n2].i = b.y;
  }
}

// one untangle slot: U=Z[k], V=Z[M-k] (k <= M/2), -> |2X[k]|, |2X[M-k]|
TIR_DEV void tir_untangle_mag(TirCpx U, TirCpx V, float2 w, float &mk, float &mmk) {
  TirCpx E2 = {TIR_FADD(U.r, V.r), TIR_FSUB(U.i, V.i)};
  TirCpx O2 = {TIR_FADD(U.i, V.i), TIR_FSUB(V.r, U.r)};
  TirCpx Tt = tir_cmul(O2, w.x, w.y);
  float pr = TIR_FADD(E2.r, Tt.r), pi = TIR_FADD(E2.i, Tt.i);
  float qr = TIR_FSUB(E2.r, Tt.r), qi = TIR_FSUB(E2.i, Tt.i);
  mk = TIR_FSQRT(TIR_FADD(TIR_FMUL(pr, pr), TIR_FMUL(pi, pi)));
  mmk = TIR_FSQRT(TIR_FADD(TIR_FMUL(qr, qr), TIR_FMUL(qi, qi)));
}

template <int WIN>
TIR_DEV void tir_pass2_compute(TirSmem<WIN> &sm, int tid, TirPass2Regs &rg) {
  using C = TirCfg<WIN>;
  const int fl = tir_frame_of<WIN>(tid), t = tir_t_of<WIN>(tid);
  const bool t0 = (t == 0);
  tir_dft16(rg.A); // A[k2] = Z[kA + N1*k2]
  tir_dft16(rg.B); // B[k2] = Z[kB + N1*k2]
  float *norm = sm.xch;
  // t >= 1 : rows kA=t, kB=N1-t.   slots s=0..7  : U=A[s]   (k = t + N1*s),        V=B[15-s]
  //                                slots 8+i     : U=B[i]   (k = N1-t + N1*i),     V=A[15-i]
  // t == 0 : rows kA=0, kB=N1/2.   slots s=0..7  : U=B[s]   (k = N1/2 + N1*s),     V=B[15-s]
  //                                slots 8+i,i>0 : U=A[i]   (k = N1*i),            V=A[16-i]
  //                                slot  8       : U=V=A[8] (k = M/2)
#pragma unroll
  for (int s = 0; s < 8; s++) {
    TirCpx U, V = rg.B[15 - s];
    U.r = t0 ? rg.B[s].r : rg.A[s].r, U.i = t0 ? rg.B[s].i : rg.A[s].i;
    const int k = (t0 ? C::N1 / 2 : t) + C::N1 * s;
    float mk, mmk;
    tir_untangle_mag(U, V, sm.tw_unt[s * C::TPF + t], mk, mmk);
    norm[tir_norm_idx(k, fl)] = mk;
    norm[tir_norm_idx(C::M - k, fl)] = mmk;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    TirCpx U, V;
    const TirCpx a_alt = rg.A[i ? i : 8], v_alt = rg.A[i ? 16 - i : 8];
    U.r = t0 ? a_alt.r : rg.B[i].r, U.i = t0 ? a_alt.i : rg.B[i].i;
    V.r = t0 ? v_alt.r : rg.A[15 - i].r, V.i = t0 ? v_alt.i : rg.A[15 - i].i;
    const int k = t0 ? (i ? C::N1 * i : C::M / 2) : (C::N1 - t) + C::N1 * i;
    float mk, mmk;
    tir_untangle_mag(U, V, sm.tw_unt[(8 + i) * C::TPF + t], mk, mmk);
    norm[tir_norm_idx(k, fl)] = mk;
    norm[tir_norm_idx(C::M - k, fl)] = mmk;
  }
}

// ---- P3 ---------------------------------------------------------------------------------------
// warp `w` (0..TIR_MEL_WARPS-1), lane = frame slot
TIR_DEV void tir_mel_phase(const float *norm, float *lg, const double2 *logtab, const TirMelParams &mp, int w,
                           int lane) {
  const int nf = mp.warp_nf[w];
  for (int q = 0; q < nf; q++) {
    const int f = mp.warp_filters[w][q];
    const int b0 = mp.start[f], n = mp.len[f], wo = mp.woff[f];
    float acc = 0.f;
    for (int b = 0; b < n; b++) acc = TIR_FADD(acc, TIR_FMUL(norm[tir_norm_idx(b0 + b, lane)], mp.w[wo + b]));
    const float v = acc < mp.log_clamp ? mp.log_clamp : acc;
    lg[f * 32 + lane] = tir_log10f_glibc(v, logtab);
  }
}

// ---- P4 ---------------------------------------------------------------------------------------
TIR_DEV void tir_dct_phase(const float *lg, const TirMelParams &mp, int j, int lane, float &c, int32_t &vq) {
  float acc = 0.f;
  for (int f = 0; f < mp.n_filters; f++) acc = TIR_FADD(acc, TIR_FMUL(lg[f * 32 + lane], mp.dct[j][f]));
  c = acc;
  vq = tir_quantize_micro(tir_coef_to_y(acc));
}
