/*
 * tir_oracle_extract.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Restates, in scalar float32 C, what src/fp_handler.c:577-671 (create_audio_fingerprints)
 * computes through libaubio 0.4.x:
 *     aubio_source_do  -> aubio_pvoc_do (win/hop, "hanningz") -> aubio_mfcc_do (40 Slaney
 *     filters, 2 coefs) -> 10*log10(fabs(c))                            (fp_handler.c:633-651)
 * and the "%f" text marshalling the value goes through on its way into SQLite
 * (src/db_ctx_handler.c:480).
 *
 * libaubio is NOT in /root/reference (link flag only: src/Makefile:24, version unpinned) and not
 * installed here, so every aubio stage below is restated from the published aubio 0.4.5/0.4.6
 * sources; the aubio file each function follows is named in its comment.  PARITY UNPINNED: no
 * reference test or golden vector exists for this boundary (see tir_oracle.h).
 *
 * The one stage aubio itself does not define bit-for-bit is the FFT (it delegates to FFTW3f or
 * to Ooura's rdft depending on how the distro built it).  The oracle therefore fixes one
 * float32 FFT, "TIR-FFT", documented at tir_fft_* below: a 16 x (M/16) Cooley-Tukey complex FFT
 * of the even/odd packed frame with a documented operation order, followed by the usual real
 * untangling.  Its accuracy is pinned against numpy's float64 FFT in tests/test_oracle_extract.py.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma (see oracle/Makefile).  -ffp-contract=off makes every
 * a*b+c below two roundings unless fmaf() is written explicitly -- this is the arithmetic of an
 * x86-64 distro build of aubio (no FMA contraction) for the aubio-defined stages.
 */
#define _GNU_SOURCE
#include "tir_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PI_D 3.14159265358979323846 /* aubio_priv.h: PI */
#define TWO_PI_D (PI_D * 2.)          /* aubio_priv.h: TWO_PI */

struct tiro_plan {
  int win, hop, n_filters, n_coefs, samplerate;
  int L;          /* win/2+1 : new_cvec(win) length, fp_handler.c:614 */
  int M;          /* win/2   : complex FFT size                        */
  float *w;       /* [win]                                              */
  float *filters; /* [n_filters][L]                                     */
  float *dct;     /* [n_coefs][n_filters]                               */
  float *edges;   /* [n_filters+2]                                      */
  float *twM_r, *twM_i; /* W_M^m   , m in [0,M)                         */
  float *twN_r, *twN_i; /* W_{2M}^k, k in [0,M/2]                       */
};

/* ------------------------------------------------------------------------------------------
 * aubio mathutils.c: new_aubio_window("hanningz") -> fvec_set_window
 *     w[i] = 0.5 * (1.0 - COS(TWO_PI * i / size))   with COS = cosf (smpl_t = float)
 * ---------------------------------------------------------------------------------------- */
static void build_window(tiro_plan *p) {
  for (int i = 0; i < p->win; i++) {
    float arg = (float)(TWO_PI_D * i / (double)p->win);
    p->w[i] = (float)(0.5 * (1.0 - (double)cosf(arg)));
  }
}

/* aubio mathutils.c: aubio_bintofreq(bin, samplerate, fftsize) -- all smpl_t */
static float bintofreq(float bin, float samplerate, float fftsize) {
  float freq = samplerate / fftsize;
  return freq * (bin > 0.f ? bin : 0.f);
}

/* ------------------------------------------------------------------------------------------
 * aubio filterbank_mel.c: aubio_filterbank_set_mel_coeffs_slaney + ..._set_triangle_bands
 * (norm = 1: unit-area triangles).  Only n_filters == 40 is what the reference asks for
 * (fp_handler.c:37,615); the edge recipe is Slaney's 13 linear + 27 log bands.
 * ---------------------------------------------------------------------------------------- */
static void build_filterbank(tiro_plan *p) {
  const int nf = p->n_filters, L = p->L;
  const float lowestFrequency = 133.3333f, linearSpacing = 66.66666666f, logSpacing = 1.0711703f;
  const int linearFilters = 13;
  float *freqs = p->edges;
  int fn;
  for (fn = 0; fn < linearFilters && fn < nf + 2; fn++)
    freqs[fn] = lowestFrequency + (float)fn * linearSpacing;
  float lastlinearCF = freqs[fn - 1];
  for (fn = 0; fn + linearFilters < nf + 2; fn++)
    freqs[fn + linearFilters] = lastlinearCF * powf(logSpacing, (float)(fn + 1));

  const float *lower = freqs, *center = freqs + 1, *upper = freqs + 2;
  float *height = (float *)malloc(sizeof(float) * nf);
  float *fftfreq = (float *)malloc(sizeof(float) * L);
  for (fn = 0; fn < nf; fn++) height[fn] = (float)(2. / (double)(upper[fn] - lower[fn]));
  for (int bin = 0; bin < L; bin++)
    fftfreq[bin] = bintofreq((float)bin, (float)p->samplerate, (float)((L - 1) * 2));
  memset(p->filters, 0, sizeof(float) * nf * L);
  for (fn = 0; fn < nf; fn++) {
    float *filt = p->filters + (size_t)fn * L;
    int bin;
    /* skip first elements */
    for (bin = 0; bin < L - 1; bin++) {
      if (fftfreq[bin] <= lower[fn] && fftfreq[bin + 1] > lower[fn]) {
        bin++;
        break;
      }
    }
    float riseInc = height[fn] / (center[fn] - lower[fn]);
    for (; bin < L - 1; bin++) {
      filt[bin] = (fftfreq[bin] - lower[fn]) * riseInc;
      if (fftfreq[bin + 1] >= center[fn]) {
        bin++;
        break;
      }
    }
    float downInc = height[fn] / (upper[fn] - center[fn]);
    for (; bin < L - 1; bin++) {
      filt[bin] += (upper[fn] - fftfreq[bin]) * downInc;
      if (filt[bin] < 0.f) filt[bin] = 0.f;
      if (fftfreq[bin + 1] >= upper[fn]) break;
    }
  }
  free(height);
  free(fftfreq);
}

/* ------------------------------------------------------------------------------------------
 * aubio mfcc.c (<= 0.4.6): new_aubio_mfcc -- explicit n_coefs x n_filters DCT-II matrix
 *     scaling = 1. / SQRT(n_filters / 2.);
 *     dct[j][i] = scaling * COS(j * (i + 0.5) * PI / n_filters);   dct[0][i] *= SQRT(2.) / 2.;
 * ---------------------------------------------------------------------------------------- */
static void build_dct(tiro_plan *p) {
  const int nf = p->n_filters;
  float scaling = (float)(1. / (double)sqrtf((float)(nf / 2.)));
  for (int i = 0; i < nf; i++) {
    for (int j = 0; j < p->n_coefs; j++) {
      float arg = (float)(j * (i + 0.5) * PI_D / nf);
      p->dct[(size_t)j * nf + i] = scaling * cosf(arg);
    }
    p->dct[i] = (float)((double)p->dct[i] * ((double)sqrtf(2.f) / 2.));
  }
}

static void build_twiddles(tiro_plan *p) {
  const int M = p->M;
  for (int m = 0; m < M; m++) {
    p->twM_r[m] = (float)cos(2.0 * PI_D * m / M);
    p->twM_i[m] = (float)(-sin(2.0 * PI_D * m / M));
  }
  for (int k = 0; k <= M / 2; k++) {
    p->twN_r[k] = (float)cos(2.0 * PI_D * k / (2 * M));
    p->twN_i[k] = (float)(-sin(2.0 * PI_D * k / (2 * M)));
  }
}

tiro_plan *tiro_plan_create(int win, int hop, int n_filters, int n_coefs, int samplerate) {
  if (!((win == 512 || win == 1024) && hop > 0 && hop <= win && n_filters >= 2 &&
        n_filters <= 40 && n_coefs >= 1 && n_coefs <= n_filters && samplerate > 0))
    return NULL;
  tiro_plan *p = (tiro_plan *)calloc(1, sizeof(*p));
  p->win = win, p->hop = hop, p->n_filters = n_filters, p->n_coefs = n_coefs;
  p->samplerate = samplerate, p->L = win / 2 + 1, p->M = win / 2;
  p->w = (float *)malloc(sizeof(float) * win);
  p->filters = (float *)malloc(sizeof(float) * n_filters * p->L);
  p->dct = (float *)malloc(sizeof(float) * n_coefs * n_filters);
  p->edges = (float *)malloc(sizeof(float) * (n_filters + 2));
  p->twM_r = (float *)malloc(sizeof(float) * p->M);
  p->twM_i = (float *)malloc(sizeof(float) * p->M);
  p->twN_r = (float *)malloc(sizeof(float) * (p->M / 2 + 1));
  p->twN_i = (float *)malloc(sizeof(float) * (p->M / 2 + 1));
  build_window(p);
  build_filterbank(p);
  build_dct(p);
  build_twiddles(p);
  return p;
}

void tiro_plan_destroy(tiro_plan *p) {
  if (!p) return;
  free(p->w), free(p->filters), free(p->dct), free(p->edges);
  free(p->twM_r), free(p->twM_i), free(p->twN_r), free(p->twN_i);
  free(p);
}

int tiro_plan_spec_len(const tiro_plan *p) { return p->L; }
const float *tiro_plan_window(const tiro_plan *p) { return p->w; }
const float *tiro_plan_filters(const tiro_plan *p) { return p->filters; }
const float *tiro_plan_dct(const tiro_plan *p) { return p->dct; }
const float *tiro_plan_band_edges(const tiro_plan *p) { return p->edges; }

size_t tiro_n_frames(size_t n_samples, int hop) { return (n_samples + (size_t)hop - 1) / (size_t)hop; }

/* ==========================================================================================
 * TIR-FFT : the float32 FFT the oracle fixes (aubio leaves it to FFTW3f / Ooura).
 *
 *   complex multiply by a twiddle w=(wr,wi), everywhere:
 *        re = fmaf(-b, wi, a*wr)      im = fmaf(b, wr, a*wi)          (a*w? rounded first)
 *   DFT4(a0..a3):  s0=a0+a2 d0=a0-a2 s1=a1+a3 d1=a1-a3
 *                  A0=s0+s1  A2=s0-s1  A1=d0-i*d1  A3=d0+i*d1
 *   DFT16: n=4*n1+n2, k=k1+4*k2: DFT4 over n1, twiddle W16^(n2*k1), DFT4 over n2;
 *          W16^2 and W16^6 use the (a+b)*h / (b-a)*h forms, W16^4 = -i is a swap,
 *          W16^1, ^3, ^9 use the general multiply with float constants.
 *   DFT32: DFT16 of even and odd inputs, X[k]=E+W32^k*O, X[k+16]=E-W32^k*O (k=8: -i swap).
 *   FFT_M (M=win/2=16*N1): n=16*n1+n2, k=k1+N1*k2: DFT_N1 over n1 for each n2, general
 *          multiply by W_M^(n2*k1) whenever n2*k1 != 0, DFT16 over n2 for each k1.
 *   real untangle for k=1..M/2-1 with Z[k]=(a,b), Z[M-k]=(c,d):
 *          E2=(a+c, b-d) O2=(b+d, c-a) T=W_{2M}^k*O2, 2X[k]=E2+T, 2X[M-k]=conj(E2-T)
 *          (k=M/2 runs through the same formula with Z[M-k]=Z[k])
 * ======================================================================================== */
typedef struct {
  float r, i;
} cpx;

static inline cpx cmul_tw(cpx x, float wr, float wi) {
  cpx o;
  o.r = fmaf(-x.i, wi, x.r * wr);
  o.i = fmaf(x.i, wr, x.r * wi);
  return o;
}

static inline void dft4(cpx a0, cpx a1, cpx a2, cpx a3, cpx *A0, cpx *A1, cpx *A2, cpx *A3) {
  cpx s0 = {a0.r + a2.r, a0.i + a2.i}, d0 = {a0.r - a2.r, a0.i - a2.i};
  cpx s1 = {a1.r + a3.r, a1.i + a3.i}, d1 = {a1.r - a3.r, a1.i - a3.i};
  A0->r = s0.r + s1.r, A0->i = s0.i + s1.i;
  A2->r = s0.r - s1.r, A2->i = s0.i - s1.i;
  A1->r = d0.r + d1.i, A1->i = d0.i - d1.r;
  A3->r = d0.r - d1.i, A3->i = d0.i + d1.r;
}

#define TIR_C1 0.92387953251128674f /* cos(pi/8) */
#define TIR_S1 0.38268343236508977f /* sin(pi/8) */
#define TIR_H 0.70710678118654752f  /* sqrt(1/2) */

static void dft16(const cpx *x, int stride, cpx *X) {
  cpx y[4][4];
  for (int n2 = 0; n2 < 4; n2++)
    dft4(x[(size_t)(n2)*stride], x[(size_t)(n2 + 4) * stride], x[(size_t)(n2 + 8) * stride],
         x[(size_t)(n2 + 12) * stride], &y[n2][0], &y[n2][1], &y[n2][2], &y[n2][3]);
  for (int n2 = 1; n2 < 4; n2++)
    for (int k1 = 1; k1 < 4; k1++) {
      cpx v = y[n2][k1], o;
      switch (n2 * k1) {
      case 1: o = cmul_tw(v, TIR_C1, -TIR_S1); break;
      case 2: o.r = (v.r + v.i) * TIR_H, o.i = (v.i - v.r) * TIR_H; break;
      case 3: o = cmul_tw(v, TIR_S1, -TIR_C1); break;
      case 4: o.r = v.i, o.i = -v.r; break;
      case 6: o.r = (v.i - v.r) * TIR_H, o.i = -((v.r + v.i) * TIR_H); break;
      default: /* 9 */ o = cmul_tw(v, -TIR_C1, TIR_S1); break;
      }
      y[n2][k1] = o;
    }
  for (int k1 = 0; k1 < 4; k1++)
    dft4(y[0][k1], y[1][k1], y[2][k1], y[3][k1], &X[k1], &X[k1 + 4], &X[k1 + 8], &X[k1 + 12]);
}

static void dft32(const tiro_plan *p, const cpx *x, int stride, cpx *X) {
  /* only used when M == 512: W32^k == W_M^(16k), the same float as (float)cos(2*pi*k/32) */
  cpx E[16], O[16];
  dft16(x, 2 * stride, E);
  dft16(x + stride, 2 * stride, O);
  for (int k = 0; k < 16; k++) {
    cpx t;
    if (k == 0)
      t = O[k];
    else if (k == 8)
      t.r = O[k].i, t.i = -O[k].r;
    else
      t = cmul_tw(O[k], p->twM_r[k * (p->M / 32)], p->twM_i[k * (p->M / 32)]);
    X[k].r = E[k].r + t.r, X[k].i = E[k].i + t.i;
    X[k + 16].r = E[k].r - t.r, X[k + 16].i = E[k].i - t.i;
  }
}

/* complex FFT of size M = 16*N1 (N1 = 16 or 32), z -> Z, out of place */
static void tir_fft_complex(const tiro_plan *p, const cpx *z, cpx *Z) {
  const int M = p->M, N1 = M / 16;
  cpx Y[16][32], col[16], out[32];
  for (int n2 = 0; n2 < 16; n2++) {
    if (N1 == 16)
      dft16(z + n2, 16, out);
    else
      dft32(p, z + n2, 16, out);
    for (int k1 = 0; k1 < N1; k1++) {
      int m = n2 * k1;
      Y[n2][k1] = m ? cmul_tw(out[k1], p->twM_r[m], p->twM_i[m]) : out[k1];
    }
  }
  for (int k1 = 0; k1 < N1; k1++) {
    for (int n2 = 0; n2 < 16; n2++) col[n2] = Y[n2][k1];
    dft16(col, 1, out);
    for (int k2 = 0; k2 < 16; k2++) Z[k1 + N1 * k2] = out[k2];
  }
}

/* P[k] = 2*X[k] for k in 1..M-1 (scaled by two, see header), X[0], X[M] unscaled in P[0], P[M] */
static void tir_rfft_scaled(const tiro_plan *p, const float *frame, cpx *P) {
  const int M = p->M;
  cpx z[512], Z[512];
  for (int n = 0; n < M; n++) z[n].r = frame[2 * n], z[n].i = frame[2 * n + 1];
  tir_fft_complex(p, z, Z);
  P[0].r = Z[0].r + Z[0].i, P[0].i = 0.f;
  P[M].r = Z[0].r - Z[0].i, P[M].i = 0.f;
  for (int k = 1; k < M / 2; k++) {
    float a = Z[k].r, b = Z[k].i, c = Z[M - k].r, d = Z[M - k].i;
    cpx E2 = {a + c, b - d}, O2 = {b + d, c - a};
    cpx T = cmul_tw(O2, p->twN_r[k], p->twN_i[k]);
    P[k].r = E2.r + T.r, P[k].i = E2.i + T.i;
    P[M - k].r = E2.r - T.r, P[M - k].i = -(E2.i - T.i);
  }
  {
    /* k = M/2 pairs with itself; it goes through the same formula (W_{2M}^{M/2} is the float
     * pair ((float)cos(pi/2), -1), not an exact -i) */
    float a = Z[M / 2].r, b = Z[M / 2].i;
    cpx E2 = {a + a, b - b}, O2 = {b + b, a - a};
    cpx T = cmul_tw(O2, p->twN_r[M / 2], p->twN_i[M / 2]);
    P[M / 2].r = E2.r + T.r, P[M / 2].i = E2.i + T.i;
  }
}

static int g_fft_kind;
static void alt_rfft_ooura_style(int M, const float *frame, float *re, float *im);
static void alt_rfft_double(int M, const float *frame, float *re, float *im);

void tiro_rfft(const tiro_plan *p, const float *frame, float *re, float *im) {
  if (g_fft_kind == 1) { alt_rfft_ooura_style(p->M, frame, re, im); return; }
  if (g_fft_kind == 2) { alt_rfft_double(p->M, frame, re, im); return; }
  cpx P[513];
  tir_rfft_scaled(p, frame, P);
  re[0] = P[0].r, im[0] = 0.f, re[p->M] = P[p->M].r, im[p->M] = 0.f;
  for (int k = 1; k < p->M; k++) re[k] = 0.5f * P[k].r, im[k] = 0.5f * P[k].i;
}

/* ==========================================================================================
 * Alternative FFT orders (SENSITIVITY STUDY ONLY, tools/oracle_fft_sensitivity.py, tests/test_oracle_fft_orders.py).
 * aubio delegates the FFT to FFTW3f or to Ooura's rdft depending on the distro build, so two legitimate
 * libaubio builds already differ in float32 rounding.  tiro_set_fft_kind() swaps the FFT under the unchanged
 * rest of the pipeline so that the identity rates between legitimate float32 FFTs can be measured:
 *   0  TIR-FFT (default; what the GPU kernel reproduces operation for operation)
 *   1  "Ooura-style": the frame packed as M complex points, an iterative radix-2 decimation-in-time complex FFT
 *      (bit reversal, float twiddle table), then the real untangling in the form Ooura's rftfsub uses
 *      (wkr = 0.5 - cos/2, wki = sin/2; a[j] -= yr ...).  Written from the published description of the
 *      algorithm; NOT an operation-for-operation copy of aubio's bundled ooura_fft8g.c (radix-8/4/2).
 *   2  float64 reference: the DFT evaluated in double with a radix-2 FFT, rounded to float32 at the end --
 *      the value every float32 FFT approximates.
 * ======================================================================================== */
void tiro_set_fft_kind(int kind) { g_fft_kind = kind; }
int tiro_get_fft_kind(void) { return g_fft_kind; }

static unsigned bitrev(unsigned x, int bits) {
  unsigned r = 0;
  for (int i = 0; i < bits; i++) r = (r << 1) | ((x >> i) & 1u);
  return r;
}

/* kind 1: X[k], k = 0..M, of the real frame (length 2M) */
static void alt_rfft_ooura_style(int M, const float *frame, float *re, float *im) {
  float ar[512], ai[512];
  int bits = 0;
  while ((1 << bits) < M) bits++;
  for (int n = 0; n < M; n++) {
    const unsigned r = bitrev((unsigned)n, bits);
    ar[r] = frame[2 * n], ai[r] = frame[2 * n + 1];
  }
  for (int len = 2; len <= M; len <<= 1) {
    const int half = len / 2;
    for (int j = 0; j < half; j++) {
      const float wr = (float)cos(TWO_PI_D * j / len), wi = (float)(-sin(TWO_PI_D * j / len));
      for (int b = j; b < M; b += len) {
        const int c = b + half;
        const float tr = ar[c] * wr - ai[c] * wi, ti = ar[c] * wi + ai[c] * wr;
        ar[c] = ar[b] - tr, ai[c] = ai[b] - ti;
        ar[b] = ar[b] + tr, ai[b] = ai[b] + ti;
      }
    }
  }
  /* untangling in the "Z[k] - w * (Z[k] - conj Z[M-k])" form Ooura's rftfsub uses, w = ((1 + sin)/2, cos/2):
   *   X[k] = Z[k] - w (Z[k] - conj(Z[M-k])),  k = 1..M-1 */
  re[0] = ar[0] + ai[0], im[0] = 0.f, re[M] = ar[0] - ai[0], im[M] = 0.f;
  for (int k = 1; k < M; k++) {
    const int m = M - k;
    const float wkr = 0.5f + 0.5f * (float)sin(PI_D * k / M), wki = 0.5f * (float)cos(PI_D * k / M);
    const float xr = ar[k] - ar[m], xi = ai[k] + ai[m];
    const float yr = wkr * xr - wki * xi, yi = wkr * xi + wki * xr;
    re[k] = ar[k] - yr, im[k] = ai[k] - yi;
  }
}

/* kind 2: float64 FFT of the float32 frame, rounded at the end */
static void alt_rfft_double(int M, const float *frame, float *re, float *im) {
  const int N = 2 * M;
  double xr[1024], xi[1024];
  int bits = 0;
  while ((1 << bits) < N) bits++;
  for (int n = 0; n < N; n++) {
    const unsigned r = bitrev((unsigned)n, bits);
    xr[r] = frame[n], xi[r] = 0.0;
  }
  for (int len = 2; len <= N; len <<= 1) {
    const int half = len / 2;
    for (int j = 0; j < half; j++) {
      const double wr = cos(TWO_PI_D * j / len), wi = -sin(TWO_PI_D * j / len);
      for (int b = j; b < N; b += len) {
        const int c = b + half;
        const double tr = xr[c] * wr - xi[c] * wi, ti = xr[c] * wi + xi[c] * wr;
        xr[c] = xr[b] - tr, xi[c] = xi[b] - ti;
        xr[b] += tr, xi[b] += ti;
      }
    }
  }
  for (int k = 0; k <= M; k++) re[k] = (float)xr[k], im[k] = (float)xi[k];
}

/* ------------------------------------------------------------------------------------------
 * aubio phasevoc.c: aubio_pvoc_do = swapbuffers (done by the caller) ; fvec_weight ; fvec_shift ;
 * aubio_fft_do -> fft.c: aubio_fft_get_norm
 *     norm[0] = ABS(re0); norm[k] = SQRT(SQR(re_k)+SQR(im_k)); norm[win/2] = ABS(re_{win/2})
 * (0.5f*sqrtf(P.r^2+P.i^2) with P=2X is the same float as sqrtf(X.r^2+X.i^2): scaling by a
 * power of two commutes with every rounding involved.)
 * ---------------------------------------------------------------------------------------- */
void tiro_pvoc_norm(const tiro_plan *p, const float *data, float *norm) {
  const int win = p->win, M = p->M;
  float buf[1024], sh[1024];
  cpx P[513];
  for (int i = 0; i < win; i++) buf[i] = data[i] * p->w[i]; /* fvec_weight */
  for (int i = 0; i < win / 2; i++) sh[i] = buf[i + win / 2], sh[i + win / 2] = buf[i]; /* fvec_shift */
  if (g_fft_kind != 0) { /* sensitivity study: another FFT order under the same pipeline */
    float re[513], im[513];
    if (g_fft_kind == 1) alt_rfft_ooura_style(M, sh, re, im);
    else alt_rfft_double(M, sh, re, im);
    norm[0] = fabsf(re[0]), norm[M] = fabsf(re[M]);
    for (int k = 1; k < M; k++) norm[k] = sqrtf(re[k] * re[k] + im[k] * im[k]);
    return;
  }
  tir_rfft_scaled(p, sh, P);
  norm[0] = fabsf(P[0].r);
  norm[M] = fabsf(P[M].r);
  for (int k = 1; k < M; k++) norm[k] = 0.5f * sqrtf(P[k].r * P[k].r + P[k].i * P[k].i);
}

/* ------------------------------------------------------------------------------------------
 * aubio mfcc.c: aubio_mfcc_do = aubio_filterbank_do (fmat_vecmul, magnitude not power) ;
 * fvec_log10 (SAFE_LOG10: log10f(max(x, VERY_SMALL_NUMBER=2e-42))) ; fmat_vecmul(dct)
 * fmat.c: fmat_vecmul -- out[j] = 0; for k: out[j] += scale[k] * s[j][k]   (sequential float)
 * ---------------------------------------------------------------------------------------- */
void tiro_mfcc(const tiro_plan *p, const float *norm, float *mel_out, float *coef) {
  float mel[40];
  const int nf = p->n_filters, L = p->L;
  for (int j = 0; j < nf; j++) {
    const float *filt = p->filters + (size_t)j * L;
    float acc = 0.f;
    for (int k = 0; k < L; k++) acc += norm[k] * filt[k];
    mel[j] = acc;
  }
  if (mel_out) memcpy(mel_out, mel, sizeof(float) * nf);
  for (int j = 0; j < nf; j++) {
    float v = ((double)mel[j] < 2.e-42) ? (float)2.e-42 : mel[j]; /* CEIL_DENORMAL */
    mel[j] = log10f(v);
  }
  for (int j = 0; j < p->n_coefs; j++) {
    const float *row = p->dct + (size_t)j * nf;
    float acc = 0.f;
    for (int k = 0; k < nf; k++) acc += mel[k] * row[k];
    coef[j] = acc;
  }
}

/* src/db_ctx_handler.c:480  ast_asprintf(&tmp_sub, "%f", real)  -> the decimal text SQLite stores.
 * jansson refuses non-finite reals, so such a value never gets a key -> NULL column. */
int32_t tiro_quantize(double y) {
  if (!isfinite(y)) return TIRO_NULL_V;
  char buf[512];
  int n = snprintf(buf, sizeof(buf), "%f", y);
  if (n <= 0 || n >= (int)sizeof(buf)) return TIRO_NULL_V;
  long long v = 0;
  int neg = 0;
  for (const char *c = buf; *c; c++) {
    if (*c == '-')
      neg = 1;
    else if (*c >= '0' && *c <= '9')
      v = v * 10 + (*c - '0');
  }
  if (neg) v = -v;
  if (v > INT32_MAX) v = INT32_MAX;
  if (v <= INT32_MIN) v = (long long)INT32_MIN + 1;
  return (int32_t)v;
}

/* src/fp_handler.c:632-661 : the hop loop.
 * `channels` interleaved PCM16 channels (pcm[sample frame][channel], n_samples sample frames): aubio_source_do
 * delivers their mean -- aubio 0.4.x source_wavread.c aubio_source_wavread_do (and source_sndfile.c
 * aubio_source_sndfile_do the same way): each channel's sample is scaled to float first (value * 1/32768,
 * aubio_source_wavread_readframe), the output sample starts at 0, the channels are added in order, the sum is divided
 * by (smpl_t)input_channels. */
size_t tiro_extract_interleaved(const tiro_plan *p, const int16_t *pcm, size_t n_samples, int channels, float *coef,
                                double *y, int32_t *vq) {
  const int win = p->win, hop = p->hop, end = win - hop, nc = p->n_coefs;
  float data[1024], dataold[1024], hopbuf[1024], norm[513], c[40];
  memset(dataold, 0, sizeof(dataold)); /* new_aubio_pvoc: dataold = zeros */
  size_t F = tiro_n_frames(n_samples, hop);
  for (size_t t = 0; t < F; t++) {
    /* aubio_source_do (source_wavread.c): PCM16 / 32768, short last block zero filled */
    size_t base = t * (size_t)hop;
    for (int i = 0; i < hop; i++) {
      size_t s = base + (size_t)i;
      if (s >= n_samples) {
        hopbuf[i] = 0.f;
      } else if (channels == 1) {
        hopbuf[i] = (float)pcm[s] / 32768.f;
      } else {
        float acc = 0.f;
        for (int ch = 0; ch < channels; ch++) acc += (float)pcm[s * (size_t)channels + (size_t)ch] / 32768.f;
        hopbuf[i] = acc / (float)channels;
      }
    }
    /* aubio_pvoc_swapbuffers */
    for (int i = 0; i < end; i++) data[i] = dataold[i];
    for (int i = 0; i < hop; i++) data[end + i] = hopbuf[i];
    for (int i = 0; i < end; i++) dataold[i] = data[i + hop];
    tiro_pvoc_norm(p, data, norm);
    tiro_mfcc(p, norm, NULL, c);
    for (int j = 0; j < nc; j++) {
      double yy = 10 * log10(fabs((double)c[j])); /* fp_handler.c:651 */
      if (coef) coef[t * nc + j] = c[j];
      if (y) y[t * nc + j] = yy;
      if (vq) vq[t * nc + j] = tiro_quantize(yy);
    }
  }
  return F;
}

size_t tiro_extract(const tiro_plan *p, const int16_t *pcm, size_t n_samples, float *coef,
                    double *y, int32_t *vq) {
  return tiro_extract_interleaved(p, pcm, n_samples, 1, coef, y, vq);
}

typedef struct {
  const tiro_plan *p;
  const int16_t *pcm;
  const uint64_t *clip_off;
  const uint64_t *frame_off;
  uint32_t n_clips;
  float *coef;
  double *y;
  int32_t *vq;
  volatile uint32_t *next;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  const int nc = j->p->n_coefs;
  for (;;) {
    uint32_t c = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
    if (c >= j->n_clips) break;
    uint64_t fo = j->frame_off[c];
    tiro_extract(j->p, j->pcm + j->clip_off[c], (size_t)(j->clip_off[c + 1] - j->clip_off[c]),
                 j->coef ? j->coef + fo * nc : NULL, j->y ? j->y + fo * nc : NULL,
                 j->vq ? j->vq + fo * nc : NULL);
  }
  return NULL;
}

size_t tiro_extract_batch(const tiro_plan *p, const int16_t *pcm, const uint64_t *clip_off,
                          uint32_t n_clips, float *coef, double *y, int32_t *vq, int n_threads) {
  uint64_t *frame_off = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)n_clips + 1));
  frame_off[0] = 0;
  for (uint32_t c = 0; c < n_clips; c++)
    frame_off[c + 1] = frame_off[c] + tiro_n_frames((size_t)(clip_off[c + 1] - clip_off[c]), p->hop);
  volatile uint32_t next = 0;
  batch_job job = {p, pcm, clip_off, frame_off, n_clips, coef, y, vq, &next};
  if (n_threads <= 1) {
    batch_worker(&job);
  } else {
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, batch_worker, &job);
    for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
  }
  size_t F = (size_t)frame_off[n_clips];
  free(frame_off);
  return F;
}
