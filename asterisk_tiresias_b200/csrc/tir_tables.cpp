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

void tir_untangle_bins(int N1, int t, int s, int &k_lo, int &k_hi) {
  k_lo = (t ? t : N1) + N1 * s;
  k_hi = (t ? N1 - t : N1 / 2) + N1 * s;
}

bool tir_build_tables(int win, int hop, int n_filters, int n_coefs, int samplerate, TirHostTables &o) {
  if (!((win == 512 || win == 1024) && hop * 2 == win && n_filters >= 2 && n_filters <= TIR_MAX_FILTERS &&
        n_coefs >= 1 && n_coefs <= TIR_MAX_COEFS && samplerate > 0))
    return false;
  o.win = win, o.hop = hop, o.samplerate = samplerate, o.n_filters = n_filters, o.n_coefs = n_coefs;
  o.M = win / 2, o.N1 = o.M / 16, o.NW = o.N1 / 2, o.L = win / 2 + 1;
  build_window(win, o.window);
  build_filterbank(n_filters, o.L, samplerate, o.filters, o.edges);
  build_dct(n_filters, n_coefs, o.dct);

  const int M = o.M, N1 = o.N1, NW = o.NW;
  o.win2.resize(M);
  for (int n = 0; n < M; n++) {
    // 2^-15 folds aubio_source's sample/32768 into the window: fl((x*2^-15)*w) == fl(x*(w*2^-15))
    o.win2[n].x = o.window[(2 * n + win / 2) % win] * (1.0f / 32768.0f);
    o.win2[n].y = o.window[(2 * n + 1 + win / 2) % win] * (1.0f / 32768.0f);
  }
  auto twM = [&](int m, float &wr, float &wi) {
    if (m == 0) {
      wr = 1.f, wi = 0.f; // multiply by one: exact (up to the sign of a zero)
    } else {
      wr = (float)cos(2.0 * kPi * m / M);
      wi = (float)(-sin(2.0 * kPi * m / M));
    }
  };
  o.win4.assign((size_t)16 * NW, float4{0, 0, 0, 0});
  o.twp4.assign((size_t)16 * NW, float4{0, 0, 0, 0});
  for (int i = 0; i < 16; i++)
    for (int r = 0; r < NW; r++) {
      // win 512: i = n1, role r = columns (2r, 2r+1), lanes = the two columns, rows k1 = i
      // win 1024: i = m / k, role r = column r, lanes = n1 (2m, 2m+1) resp. rows k1 (k, k+16)
      const int n_lo = win == 512 ? 16 * i + 2 * r : 16 * (2 * i) + r;
      const int n_hi = win == 512 ? 16 * i + 2 * r + 1 : 16 * (2 * i + 1) + r;
      o.win4[(size_t)i * NW + r] = float4{o.win2[n_lo].x, o.win2[n_hi].x, o.win2[n_lo].y, o.win2[n_hi].y};
      const int m_lo = win == 512 ? (2 * r) * i : r * i;
      const int m_hi = win == 512 ? (2 * r + 1) * i : r * (i + 16);
      float4 t;
      twM(m_lo, t.x, t.z), twM(m_hi, t.y, t.w);
      o.twp4[(size_t)i * NW + r] = t;
    }
  o.twu4.assign((size_t)NW * 8, float4{0, 0, 0, 0});
  for (int t = 0; t < NW; t++)
    for (int s = 0; s < 8; s++) {
      int k_lo, k_hi;
      tir_untangle_bins(N1, t, s, k_lo, k_hi);
      float4 w;
      w.x = (float)cos(2.0 * kPi * k_lo / (2 * M)), w.z = (float)(-sin(2.0 * kPi * k_lo / (2 * M)));
      w.y = (float)cos(2.0 * kPi * k_hi / (2 * M)), w.w = (float)(-sin(2.0 * kPi * k_hi / (2 * M)));
      o.twu4[(size_t)t * 8 + s] = w;
    }

  // mel sweep tables (TirMelParams)
  TirMelParams &mp = o.mel;
  std::memset(&mp, 0, sizeof(mp));
  mp.n_filters = n_filters, mp.n_coefs = n_coefs;
  mp.log_clamp = (float)2.e-42; // aubio_priv.h VERY_SMALL_NUMBER, as the float log10f receives
  {
    const double2 lt[16] = TIR_LOGF_TAB_INIT;
    mp.lg_dead = tir_log10f_glibc(mp.log_clamp, lt);
  }
  for (int j = 0; j < n_coefs; j++)
    for (int f = 0; f < n_filters; f++) mp.dct[j][f] = o.dct[(size_t)j * n_filters + f];
  std::vector<int> first(n_filters, 0), last(n_filters, -1), live;
  for (int f = 0; f < n_filters; f++) {
    const float *filt = o.filters.data() + (size_t)f * o.L;
    int a = -1, b = -2;
    for (int i = 0; i < o.L; i++)
      if (filt[i] != 0.f) {
        if (a < 0) a = i;
        b = i;
      }
    if (a < 0) {
      mp.dead[f] = 1;
    } else {
      first[f] = a, last[f] = b;
      live.push_back(f);
    }
  }
  // the sweep needs: completion order == filter order, and same-parity filters disjoint
  for (size_t i = 0; i + 1 < live.size(); i++)
    if (last[live[i]] > last[live[i + 1]] || first[live[i]] > first[live[i + 1]]) return false;
  for (size_t i = 0; i < live.size(); i++)
    for (size_t j = i + 1; j < live.size(); j++)
      if (((live[i] ^ live[j]) & 1) == 0 && first[live[j]] <= last[live[i]]) return false;
  mp.n_live = (int)live.size();
  mp.live_prefix = 1;
  for (size_t i = 0; i < live.size(); i++) mp.live[i] = (uint8_t)live[i], mp.live_prefix &= live[i] == (int)i;
  // contiguous segments of live filters, one per warp, balanced on issue slots (measured with ncu:
  // 5.25 per bin, 26 per filter, 20 per segment); the n_coefs coefficient warps first run the DCT
  // of the previous tile (clock64 trace: worth about 270 slots), so their segments are smaller
  const int n_sweep = NW;
  auto seg_cost = [&](int ia, int ib) { // live[ia..ib)
    return ia >= ib ? 0 : 16 * (last[live[ib - 1]] - first[live[ia]] + 1) / 4 + 24 * (ib - ia) + 20;
  };
  const int dct_cost = 270;
  const int nl = (int)live.size();
  std::vector<int> cut(n_sweep + 1, nl);
  cut[0] = 0;
  if (nl > 0) {
    // smallest bottleneck by bisection on the cost bound, greedy fill
    int lo = 0, hi = seg_cost(0, nl) + dct_cost;
    auto fits = [&](int bound, std::vector<int> *out) {
      int i = 0;
      for (int s = 0; s < n_sweep; s++) {
        int j = i;
        while (j < nl && seg_cost(i, j + 1) + (s < n_coefs ? dct_cost : 0) <= bound) j++;
        if (out) (*out)[s + 1] = j;
        i = j;
      }
      return i == nl;
    };
    while (lo < hi) {
      const int mid = (lo + hi) / 2;
      if (fits(mid, nullptr)) hi = mid; else lo = mid + 1;
    }
    fits(lo, &cut);
  }
  int nruns = 0, nw2 = 0, nsegs = 0;
  for (int s = 0; s < n_sweep; s++) {
    const int ia = cut[s], ib = cut[s + 1];
    const int seg = nsegs++;
    mp.seg_run0[seg] = (int16_t)nruns, mp.seg_woff[seg] = (int16_t)nw2;
    if (ia >= ib) {
      mp.seg_bin0[seg] = 0, mp.seg_nruns[seg] = 0;
      continue;
    }
    const int b0 = first[live[ia]], b1 = last[live[ib - 1]];
    nw2 += nw2 & 1; // every run's weight records start 16-byte aligned: the sweep loads two bins' records at once
    if (nw2 + (b1 - b0 + 1) + (ib - ia) + 1 > TIR_MAX_W2 || nruns + (ib - ia) + 1 > TIR_MAX_RUNS) return false;
    mp.seg_bin0[seg] = (int16_t)b0;
    mp.seg_woff[seg] = (int16_t)nw2;
    // magnitudes arrive as 2^33 * |X[k]| (FFT scaled by 2, sqrt by 2^32); power-of-two scaling is exact
    const float sc = 1.0f / 8589934592.0f;
    int done = b0; // bins [b0, done) are already covered by earlier runs
    for (int i = ia; i < ib; i++) {
      const int fr = live[i];
      for (int bin = done; bin <= last[fr]; bin++) {
        float2 w{0.f, 0.f};
        for (int k = ia; k < ib; k++) {
          const int f = live[k];
          if (bin < first[f] || bin > last[f]) continue;
          const float v = o.filters[(size_t)f * o.L + bin] * sc;
          if (f & 1) w.y = v; else w.x = v;
        }
        mp.w2[nw2++] = w;
      }
      if (nw2 & 1) mp.w2[nw2++] = float2{0.f, 0.f}; // pad: an odd run sweeps one more bin with zero weights (adds +0: exact)
      mp.run_bins[nruns] = (int16_t)(last[fr] + 1 - done), mp.run_emit[nruns] = (int8_t)fr;
      done = last[fr] + 1, nruns++;
    }
    mp.run_bins[nruns] = 0, mp.run_emit[nruns] = 0, nruns++; // sentinel: the sweep fetches one run ahead
    mp.seg_nruns[seg] = (int16_t)(ib - ia);
  }
  mp.n_segs = nsegs;
  return true;
}
