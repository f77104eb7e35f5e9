// tir_tables.cpp -- see tir_tables.h.  Plain host C++ (no CUDA calls); float32 arithmetic written
// one operation per statement so that the tables are the same floats an x86-64 build of aubio
// 0.4.x produces (window: mathutils.c fvec_set_window "hanningz"; filterbank: filterbank_mel.c
// aubio_filterbank_set_mel_coeffs_slaney / _set_triangle_bands; DCT: mfcc.c new_aubio_mfcc).
// Compile with -ffp-contract=off.
#include "tir_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

namespace {
constexpr double kPi = 3.14159265358979323846;

void build_window(int win, std::vector<float> &w) {
  w.resize(win);
  for (int i = 0; i < win; i++) {
    float arg = (float)((kPi * 2.) * i / (double)win);
    w[i] = (float)(0.5 * (1.0 - (double)cosf(arg)));
  }
}

void build_filterbank(int n_filters, int L, int samplerate, std::vector<float> &filters, std::vector<float> &edges) {
  const float lowestFrequency = 133.3333f, linearSpacing = 66.66666666f, logSpacing = 1.0711703f;
  const int linearFilters = 13;
  edges.assign(n_filters + 2, 0.f);
  int fn = 0;
  for (; fn < linearFilters && fn < n_filters + 2; fn++) {
    float step = (float)fn * linearSpacing;
    edges[fn] = lowestFrequency + step;
  }
  const float lastlinearCF = edges[fn - 1];
  for (fn = 0; fn + linearFilters < n_filters + 2; fn++) {
    float pw = powf(logSpacing, (float)(fn + 1));
    edges[fn + linearFilters] = lastlinearCF * pw;
  }
  const float *lower = edges.data(), *center = edges.data() + 1, *upper = edges.data() + 2;
  std::vector<float> height(n_filters), fftfreq(L);
  for (fn = 0; fn < n_filters; fn++) {
    float span = upper[fn] - lower[fn];
    height[fn] = (float)(2. / (double)span);
  }
  const float binhz = (float)samplerate / (float)((L - 1) * 2);
  for (int bin = 0; bin < L; bin++) fftfreq[bin] = binhz * (float)bin;
  filters.assign((size_t)n_filters * L, 0.f);
  for (fn = 0; fn < n_filters; fn++) {
    float *filt = filters.data() + (size_t)fn * L;
    int bin = 0;
    for (; bin < L - 1; bin++)
      if (fftfreq[bin] <= lower[fn] && fftfreq[bin + 1] > lower[fn]) {
        bin++;
        break;
      }
    float riseDen = center[fn] - lower[fn];
    float riseInc = height[fn] / riseDen;
    for (; bin < L - 1; bin++) {
      float d = fftfreq[bin] - lower[fn];
      filt[bin] = d * riseInc;
      if (fftfreq[bin + 1] >= center[fn]) {
        bin++;
        break;
      }
    }
    float downDen = upper[fn] - center[fn];
    float downInc = height[fn] / downDen;
    for (; bin < L - 1; bin++) {
      float d = upper[fn] - fftfreq[bin];
      float add = d * downInc;
      filt[bin] = filt[bin] + add;
      if (filt[bin] < 0.f) filt[bin] = 0.f;
      if (fftfreq[bin + 1] >= upper[fn]) break;
    }
  }
}

void build_dct(int n_filters, int n_coefs, std::vector<float> &dct) {
  dct.assign((size_t)n_coefs * n_filters, 0.f);
  float root = sqrtf((float)(n_filters / 2.));
  float scaling = (float)(1. / (double)root);
  for (int i = 0; i < n_filters; i++) {
    for (int j = 0; j < n_coefs; j++) {
      float arg = (float)(j * (i + 0.5) * kPi / n_filters);
      float c = cosf(arg);
      dct[(size_t)j * n_filters + i] = scaling * c;
    }
    dct[i] = (float)((double)dct[i] * ((double)sqrtf(2.f) / 2.));
  }
}
} // namespace

int tir_untangle_bin(int N1, int M, int slot, int t) {
  (void)M;
  if (slot < 8) return (t ? t : N1 / 2) + N1 * slot;
  return (N1 - t) + N1 * (slot - 8);
}

bool tir_build_tables(int win, int hop, int n_filters, int n_coefs, int samplerate, TirHostTables &o) {
  if (!((win == 512 || win == 1024) && hop * 2 == win && n_filters >= 2 && n_filters <= TIR_MAX_FILTERS &&
        n_coefs >= 1 && n_coefs <= TIR_MAX_COEFS && samplerate > 0))
    return false;
  o.win = win, o.hop = hop, o.samplerate = samplerate, o.n_filters = n_filters, o.n_coefs = n_coefs;
  o.M = win / 2, o.N1 = o.M / 16, o.TPF = o.N1 / 2, o.L = win / 2 + 1;
  build_window(win, o.window);
  build_filterbank(n_filters, o.L, samplerate, o.filters, o.edges);
  build_dct(n_filters, n_coefs, o.dct);

  const int M = o.M, N1 = o.N1, TPF = o.TPF;
  o.win2.resize(M);
  for (int n = 0; n < M; n++) {
    // 2^-15 folds aubio_source's sample/32768 into the window: fl((x*2^-15)*w) == fl(x*(w*2^-15))
    o.win2[n].x = o.window[(2 * n + win / 2) % win] * (1.0f / 32768.0f);
    o.win2[n].y = o.window[(2 * n + 1 + win / 2) % win] * (1.0f / 32768.0f);
  }
  o.tw_pass.resize((size_t)N1 * 16);
  for (int k1 = 0; k1 < N1; k1++)
    for (int n2 = 0; n2 < 16; n2++) {
      const int m = n2 * k1;
      float2 w;
      if (m == 0) {
        w.x = 1.f, w.y = 0.f; // multiply by one: exact (up to the sign of a zero)
      } else {
        w.x = (float)cos(2.0 * kPi * m / M);
        w.y = (float)(-sin(2.0 * kPi * m / M));
      }
      o.tw_pass[(size_t)k1 * 16 + n2] = w;
    }
  o.tw_unt.resize((size_t)16 * TPF);
  for (int s = 0; s < 16; s++)
    for (int t = 0; t < TPF; t++) {
      const int k = tir_untangle_bin(N1, M, s, t);
      float2 w;
      w.x = (float)cos(2.0 * kPi * k / (2 * M));
      w.y = (float)(-sin(2.0 * kPi * k / (2 * M)));
      o.tw_unt[(size_t)s * TPF + t] = w;
    }
  o.tw32.resize(16);
  for (int k = 0; k < 16; k++) {
    o.tw32[k].x = (float)cos(2.0 * kPi * k / 32);
    o.tw32[k].y = (float)(-sin(2.0 * kPi * k / 32));
  }

  // banded mel weights
  TirMelParams &mp = o.mel;
  std::memset(&mp, 0, sizeof(mp));
  mp.n_filters = n_filters, mp.n_coefs = n_coefs;
  mp.log_clamp = (float)2.e-42; // aubio_priv.h VERY_SMALL_NUMBER, as the float log10f receives
  int nnz = 0;
  for (int f = 0; f < n_filters; f++) {
    const float *filt = o.filters.data() + (size_t)f * o.L;
    int first = -1, last = -2;
    for (int b = 0; b < o.L; b++)
      if (filt[b] != 0.f) {
        if (first < 0) first = b;
        last = b;
      }
    const int len = first < 0 ? 0 : last - first + 1;
    const int len4 = (len + 3) & ~3; // zero padded to whole float4s
    if (nnz + len4 > TIR_MAX_NNZ) return false;
    mp.start[f] = (int16_t)(first < 0 ? 0 : first), mp.len[f] = (int16_t)len, mp.woff[f] = (int16_t)nnz;
    float *wdst = reinterpret_cast<float *>(mp.w4) + nnz;
    // magnitudes arrive as 2^33 * |X[k]| (FFT scaled by 2, sqrt by 2^32); power-of-two scaling is exact
    for (int b = 0; b < len; b++) wdst[b] = filt[first + b] * (1.0f / 8589934592.0f);
    nnz += len4;
  }
  for (int j = 0; j < n_coefs; j++)
    for (int f = 0; f < n_filters; f++) mp.dct[j][f] = o.dct[(size_t)j * n_filters + f];
  // longest-processing-time assignment of filters to the mel warps
  std::vector<int> order(n_filters);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return mp.len[a] > mp.len[b]; });
  int load[TIR_MEL_WARPS] = {0};
  for (int f : order) {
    int w = (int)(std::min_element(load, load + TIR_MEL_WARPS) - load);
    mp.warp_filters[w][mp.warp_nf[w]++] = (uint8_t)f;
    load[w] += mp.len[f] + 24; // + fixed cost of the clamp/log10f per filter
  }
  return true;
}
