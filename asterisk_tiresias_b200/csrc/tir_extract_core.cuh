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
  int16_t len[TIR_MAX_FILTERS];     // number of bins (weights are zero padded to a multiple of 4)
  int16_t woff[TIR_MAX_FILTERS];    // offset into w[], multiple of 4
  uint8_t warp_nf[TIR_MEL_WARPS];   // filters handled by mel warp w
  uint8_t warp_filters[TIR_MEL_WARPS][TIR_MAX_FILTERS];
  float4 w4[TIR_MAX_NNZ / 4];       // 2^-33 * aubio filter weight (scaled FFT and scaled sqrt)
  float dct[TIR_MAX_COEFS][TIR_MAX_FILTERS];
  float log_clamp;                  // (float)2e-42 : aubio VERY_SMALL_NUMBER
  int n_filters, n_coefs;
};

static_assert(sizeof(float4) == 16 && alignof(float4) == 16 && alignof(float2) == 8 && alignof(double2) == 16,
              "vector types must have the CUDA layout in every translation unit");
static_assert(sizeof(TirMelParams) % 16 == 0, "TirMelParams layout");

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
  static constexpr int NORM_STRIDE = 33;       // words per bin row of the magnitude buffer (32 frames + 1 pad)
};

template <int WIN>
struct TirSmem {
  using C = TirCfg<WIN>;
  static constexpr int PCM_WORDS = ((C::T + 1) * C::PCM_STRIDE_W + 3) & ~3; // keep what follows 16-byte aligned
  static constexpr int XCH_WORDS = C::T * C::XCH_FRAME_W;
  static constexpr int NORM_WORDS = (C::M + 1) * C::NORM_STRIDE;
  static_assert(NORM_WORDS <= XCH_WORDS, "magnitudes alias the exchange buffer");
  static_assert(XCH_WORDS % 4 == 0 && C::XCH_FRAME_W % 2 == 0, "float2 / double2 alignment of the members below");
  static_assert(TIR_MAX_FILTERS * 32 <= PCM_WORDS, "log-mel values alias the consumed PCM buffer");
  uint32_t pcm[2][PCM_WORDS];   // double buffered: tile N+1 streams in (cp.async) while tile N computes
  float xch[XCH_WORDS];
  float2 win2[C::M];            // window pairs in z[] order, pre-scaled by 2^-15
  float2 tw_pass[C::N1 * 16];   // [k1][n2]  W_M^(n2*k1)
  float2 tw_unt[16 * C::TPF];   // [slot][t] W_{2M}^k
  double2 logtab[16];
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
// warp w, lane l: t = l % TPF, q = l / TPF, frame slot fl = (32/TPF)*w + q (consecutive frames in a
// warp).  The magnitude / log-mel buffers are indexed by the permuted slot col = (NT/32)*q + w, which
// together with the strides (PCM chunk 136 words, exchange frame 560 words, magnitude row 33 words)
// makes every shared-memory access of P1..P4 bank-conflict free.
template <int WIN>
TIR_DEV int tir_frame_of(int tid) { return tid / TirCfg<WIN>::TPF; }
template <int WIN>
TIR_DEV int tir_t_of(int tid) { return tid % TirCfg<WIN>::TPF; }
template <int WIN>
TIR_DEV int tir_col_of_frame(int fl) {
  using C = TirCfg<WIN>;
  constexpr int FPW = 32 / C::TPF; // frames per warp
  return (C::NT / 32) * (fl % FPW) + fl / FPW;
}

// magnitude buffer: bin-major rows of 33 words, frame slot within the row
#define TIR_NORM_IDX(bin, fl) ((bin) * 33 + (fl))

// ---- P1 ---------------------------------------------------------------------------------------
// `frames_valid`: frames of this tile that exist; threads of other frames still run (zeros).
template <int WIN>
TIR_DEV void tir_pass1(TirSmem<WIN> &sm, const uint32_t *pcm, int tid) {
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
      const uint32_t word = pcm[chunk * C::PCM_STRIDE_W + 16 * (n1 & 7) + n2];
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

// one untangle slot: U=Z[k], V=Z[M-k] (k <= M/2), -> 2^32*|2X[k]|, 2^32*|2X[M-k]|
// (the 2^32 of the scaled square root and the 2 of the scaled FFT are folded, exactly, into the
// mel weights: w4 = filter * 2^-33)
TIR_DEV void tir_untangle_mag(TirCpx U, TirCpx V, float2 w, float &mk, float &mmk) {
  TirCpx E2 = {TIR_FADD(U.r, V.r), TIR_FSUB(U.i, V.i)};
  TirCpx O2 = {TIR_FADD(U.i, V.i), TIR_FSUB(V.r, U.r)};
  TirCpx Tt = tir_cmul(O2, w.x, w.y);
  float pr = TIR_FADD(E2.r, Tt.r), pi = TIR_FADD(E2.i, Tt.i);
  float qr = TIR_FSUB(E2.r, Tt.r), qi = TIR_FSUB(E2.i, Tt.i);
  mk = TIR_FSQRT_SCALED64(TIR_FADD(TIR_FMUL(pr, pr), TIR_FMUL(pi, pi)));
  mmk = TIR_FSQRT_SCALED64(TIR_FADD(TIR_FMUL(qr, qr), TIR_FMUL(qi, qi)));
}

template <int WIN>
TIR_DEV void tir_pass2_compute(TirSmem<WIN> &sm, int tid, TirPass2Regs &rg) {
  using C = TirCfg<WIN>;
  static_assert(C::NORM_STRIDE == 33, "TIR_NORM_IDX");
  const int fl = tir_col_of_frame<WIN>(tir_frame_of<WIN>(tid)), t = tir_t_of<WIN>(tid);
  const bool t0 = (t == 0);
  tir_dft16(rg.A); // A[k2] = Z[kA + N1*k2]
  tir_dft16(rg.B); // B[k2] = Z[kB + N1*k2]
  // rows: t >= 1: kA = t, kB = N1-t ; t == 0: kA = 0, kB = N1/2.  Sixteen (k, M-k) pairs per thread:
  //   slots s=0..7   k = (t ? t : N1/2) + N1*s   U = t ? A[s] : B[s]      V = B[15-s]
  //   slots 8+i      k = (N1 - t) + N1*i         U = t ? B[i] : A[i+1]    V = A[15-i]
  // (for t == 0 the last slot is k = M/2 paired with itself: U = V = A[8])
  float *nlo = sm.xch + TIR_NORM_IDX(t0 ? C::N1 / 2 : t, fl);
  float *nhi = sm.xch + TIR_NORM_IDX(C::M - (t0 ? C::N1 / 2 : t), fl);
#pragma unroll
  for (int s = 0; s < 8; s++) {
    TirCpx U, V = rg.B[15 - s];
    U.r = t0 ? rg.B[s].r : rg.A[s].r, U.i = t0 ? rg.B[s].i : rg.A[s].i;
    float mk, mmk;
    tir_untangle_mag(U, V, sm.tw_unt[s * C::TPF + t], mk, mmk);
    nlo[TIR_NORM_IDX(C::N1 * s, 0)] = mk;
    nhi[-TIR_NORM_IDX(C::N1 * s, 0)] = mmk;
  }
  nlo = sm.xch + TIR_NORM_IDX(C::N1 - t, fl);
  nhi = sm.xch + TIR_NORM_IDX(C::M - (C::N1 - t), fl);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    TirCpx U, V = rg.A[15 - i];
    U.r = t0 ? rg.A[i + 1].r : rg.B[i].r, U.i = t0 ? rg.A[i + 1].i : rg.B[i].i;
    float mk, mmk;
    tir_untangle_mag(U, V, sm.tw_unt[(8 + i) * C::TPF + t], mk, mmk);
    nlo[TIR_NORM_IDX(C::N1 * i, 0)] = mk;
    nhi[-TIR_NORM_IDX(C::N1 * i, 0)] = mmk;
  }
}

// ---- P3 ---------------------------------------------------------------------------------------
// warp `w` (0..TIR_MEL_WARPS-1, warp-uniform), lane = permuted frame slot (tir_col_of_frame)
TIR_DEV void tir_mel_phase(const float *norm, float *lg, const double2 *logtab, const TirMelParams &mp, int w,
                           int lane) {
  const int nf = mp.warp_nf[w];
  for (int q = 0; q < nf; q++) {
    const int f = mp.warp_filters[w][q];
    const int n4 = (mp.len[f] + 3) >> 2;
    const float *p = norm + TIR_NORM_IDX(mp.start[f], lane);
    const float4 *wp = mp.w4 + (mp.woff[f] >> 2);
    float acc = 0.f;
    // zero padded weights: acc + x*0 == acc (x is finite: a magnitude, or stale exchange data above
    // the last bin row, which lies inside sm.xch)
#pragma unroll 2
    for (int b = 0; b < n4; b++) {
      const float4 w = wp[b];
      acc = TIR_FADD(acc, TIR_FMUL(p[TIR_NORM_IDX(4 * b + 0, 0)], w.x));
      acc = TIR_FADD(acc, TIR_FMUL(p[TIR_NORM_IDX(4 * b + 1, 0)], w.y));
      acc = TIR_FADD(acc, TIR_FMUL(p[TIR_NORM_IDX(4 * b + 2, 0)], w.z));
      acc = TIR_FADD(acc, TIR_FMUL(p[TIR_NORM_IDX(4 * b + 3, 0)], w.w));
    }
    const float v = acc < mp.log_clamp ? mp.log_clamp : acc;
    lg[f * 32 + lane] = tir_log10f_glibc(v, logtab);
  }
}

// ---- P4 ---------------------------------------------------------------------------------------
// `col` = permuted slot of the frame this thread finishes
TIR_DEV void tir_dct_phase(const float *lg, const TirMelParams &mp, int j, int col, float &c, int32_t &vq) {
  float acc = 0.f;
#pragma unroll 8
  for (int f = 0; f < mp.n_filters; f++) acc = TIR_FADD(acc, TIR_FMUL(lg[f * 32 + col], mp.dct[j][f]));
  c = acc;
  vq = tir_quantize_micro(tir_coef_to_y(acc));
}
