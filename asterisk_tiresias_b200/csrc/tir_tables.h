// tir_tables.h -- host-side construction of the constant tables of one extraction plan
// (what new_aubio_pvoc(win,hop) + new_aubio_mfcc(win,40,2,samplerate) set up once per call in the
// reference, src/fp_handler.c:613-617), in the layouts the kernels read them.
#pragma once
#include <cstdint>
#include <vector>

#include "tir_extract_core.cuh"

struct TirHostTables {
  int win = 0, hop = 0, samplerate = 0, n_filters = 0, n_coefs = 0;
  int M = 0, N1 = 0, TPF = 0, L = 0;
  // aubio-layout tables (also exported for the table parity tests)
  std::vector<float> window;   // [win]             hanningz
  std::vector<float> filters;  // [n_filters][L]    Slaney triangles, unit area
  std::vector<float> dct;      // [n_coefs][n_filters]
  std::vector<float> edges;    // [n_filters+2]
  // kernel-layout tables
  std::vector<float2> win2;    // [M]      (w[(2n+win/2)%win], w[(2n+1+win/2)%win]) * 2^-15
  std::vector<float2> tw_pass; // [N1][16] W_M^(n2*k1)
  std::vector<float2> tw_unt;  // [16][TPF] W_{2M}^k(slot,t)
  std::vector<float2> tw32;    // [16] W_32^k (win 1024 only)
  TirMelParams mel;
};

// returns false when (win,hop,...) is not a supported plan
bool tir_build_tables(int win, int hop, int n_filters, int n_coefs, int samplerate, TirHostTables &out);

// bin index handled by untangle slot `slot` of thread `t` (the smaller bin of the pair)
int tir_untangle_bin(int N1, int M, int slot, int t);
