// tir_tables.h -- host-side construction of the constant tables of one extraction plan
// (what new_aubio_pvoc(win,hop) + new_aubio_mfcc(win,40,2,samplerate) set up once per call in the
// reference, src/fp_handler.c:613-617), in the layouts the kernels read them.
#pragma once
#include <cstdint>
#include <vector>

#include "tir_extract_core.cuh"

struct TirHostTables {
  int win = 0, hop = 0, samplerate = 0, n_filters = 0, n_coefs = 0;
  int M = 0, N1 = 0, NW = 0, L = 0;
  // aubio-layout tables (also exported for the table parity tests)
  std::vector<float> window;   // [win]             hanningz
  std::vector<float> filters;  // [n_filters][L]    Slaney triangles, unit area
  std::vector<float> dct;      // [n_coefs][n_filters]
  std::vector<float> edges;    // [n_filters+2]
  // kernel-layout tables (float4 = the values of the two packed lanes, see tir_extract_core.cuh)
  std::vector<float2> win2;    // [M]      (w[(2n+win/2)%win], w[(2n+1+win/2)%win]) * 2^-15
  std::vector<float4> win4;    // [16][NW] window of the two lanes (x: re lo, y: re hi, z: im lo, w: im hi)
  std::vector<float4> twp4;    // [16][NW] W_M^(n2*k1) of the two lanes (x: wr lo, y: wr hi, z: wi lo, w: wi hi)
  std::vector<float4> twu4;    // [NW][8]  W_{2M}^k of the two untangle slots of a slot pair
  TirMelParams mel;
};

// returns false when (win,hop,...) is not a supported plan
bool tir_build_tables(int win, int hop, int n_filters, int n_coefs, int samplerate, TirHostTables &out);

// bins (k <= M/2) of the two lanes of untangle slot pair `s` of role `t` (tir_pass2_compute)
void tir_untangle_bins(int N1, int t, int s, int &k_lo, int &k_hi);
