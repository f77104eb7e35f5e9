// emul_extract.cpp -- TEST ONLY: runs the per-thread phase functions of the extraction kernel
// (asterisk_tiresias_b200/csrc/tir_extract_core.cuh) on the host, thread by thread with the
// barriers of the kernel turned into loop boundaries, so that the CPU test-suite can check the
// kernel's arithmetic and shared-memory indexing against the oracle without a GPU.
// This is not a fallback: it is built only by tests/ and is never linked into libtiresias_gpu.so.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../asterisk_tiresias_b200/csrc/tir_extract_core.cuh"
#include "../../asterisk_tiresias_b200/csrc/tir_tables.h"

extern "C" int emul_extract(const int16_t *pcm, uint64_t n_samples, int samplerate, float *coef, int32_t *vq) {
  using C = TirCfg<512>;
  TirHostTables tab;
  if (!tir_build_tables(512, 256, 40, 2, samplerate, tab)) return -1;
  auto *sm = new TirSmem<512>();
  std::memset(sm, 0, sizeof(*sm));
  std::memcpy(sm->win2, tab.win2.data(), sizeof(sm->win2));
  std::memcpy(sm->tw_pass, tab.tw_pass.data(), sizeof(sm->tw_pass));
  std::memcpy(sm->tw_unt, tab.tw_unt.data(), sizeof(sm->tw_unt));
  const double2 lt[16] = TIR_LOGF_TAB_INIT;
  std::memcpy(sm->logtab, lt, sizeof(lt));
  const int64_t nsamp = (int64_t)n_samples;
  const int64_t nframes = (nsamp + C::HOP - 1) / C::HOP;
  std::vector<TirPass2Regs> regs(C::NT);
  for (int64_t f0 = 0; f0 < nframes; f0 += C::T) {
    const int nvalid = (int)std::min<int64_t>(C::T, nframes - f0);
    // P0
    for (int chunk = 0; chunk <= C::T; chunk++)
      for (int i = 0; i < C::HOP; i += 2) {
        int64_t s = (f0 - 1 + chunk) * C::HOP + i;
        uint32_t lo = (s >= 0 && s < nsamp) ? (uint16_t)pcm[s] : 0;
        uint32_t hi = (s + 1 >= 0 && s + 1 < nsamp) ? (uint16_t)pcm[s + 1] : 0;
        sm->pcm[0][chunk * C::PCM_STRIDE_W + i / 2] = lo | (hi << 16);
      }
    for (int tid = 0; tid < C::NT; tid++) tir_pass1<512>(*sm, sm->pcm[0], tid);
    for (int tid = 0; tid < C::NT; tid++) tir_pass2_load<512>(*sm, tid, regs[tid]);
    for (int tid = 0; tid < C::NT; tid++) tir_pass2_compute<512>(*sm, tid, regs[tid]);
    float *lg = reinterpret_cast<float *>(sm->pcm[0]); // as in the kernel: aliases the consumed PCM buffer
    for (int w = 0; w < TIR_MEL_WARPS; w++)
      for (int lane = 0; lane < 32; lane++) tir_mel_phase(sm->xch, lg, sm->logtab, tab.mel, w, lane);
    for (int j = 0; j < 2; j++)
      for (int lane = 0; lane < nvalid; lane++) {
        float c;
        int32_t v;
        tir_dct_phase(lg, tab.mel, j, tir_col_of_frame<512>(lane), c, v);
        coef[(f0 + lane) * 2 + j] = c;
        vq[(f0 + lane) * 2 + j] = v;
      }
  }
  delete sm;
  return 0;
}

extern "C" int emul_tables(int win, int hop, int samplerate, float *window, float *filters, float *dct) {
  TirHostTables tab;
  if (!tir_build_tables(win, hop, 40, 2, samplerate, tab)) return -1;
  std::memcpy(window, tab.window.data(), tab.window.size() * 4);
  std::memcpy(filters, tab.filters.data(), tab.filters.size() * 4);
  std::memcpy(dct, tab.dct.data(), tab.dct.size() * 4);
  return 0;
}

extern "C" float emul_log10f(float x) {
  const double2 lt[16] = TIR_LOGF_TAB_INIT;
  return tir_log10f_glibc(x, lt);
}
extern "C" int32_t emul_quantize(double y) { return tir_quantize_micro(y); }

// count floats in [first, last] (bit patterns, stride `step`) where the model differs from libm
extern "C" uint64_t emul_log10f_sweep(uint32_t first, uint32_t last, uint32_t step) {
  const double2 lt[16] = TIR_LOGF_TAB_INIT;
  uint64_t bad = 0;
  for (uint64_t u = first; u <= last; u += step) {
    float x = TIR_U2F((uint32_t)u);
    float a = tir_log10f_glibc(x, lt), b = log10f(x);
    bad += TIR_F2U(a) != TIR_F2U(b);
  }
  return bad;
}
