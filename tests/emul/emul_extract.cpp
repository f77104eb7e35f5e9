// emul_extract.cpp -- TEST ONLY: runs the per-thread phase functions of the extraction kernel
// (asterisk_tiresias_b200/csrc/tir_extract_core.cuh) on the host, thread by thread with the
// barriers of the kernel turned into loop boundaries, so that the CPU test-suite can check the
// kernel's arithmetic and shared-memory indexing against the oracle without a GPU.
// This is not a fallback: it is built only by tests/ and is never linked into libtiresias_gpu.so.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../asterisk_tiresias_b200/csrc/tir_extract_core.cuh"
#include "../../asterisk_tiresias_b200/csrc/tir_tables.h"

static bool force_indexed_logs = false; // exercise the table-indexed P3b on plans that would take the prefix path
extern "C" void emul_force_indexed_logs(int on) { force_indexed_logs = on != 0; }

template <int WIN>
static int emul_extract_t(const int16_t *pcm, uint64_t n_samples, int samplerate, float *coef, int32_t *vq) {
  using C = TirCfg<WIN>;
  TirHostTables tab;
  if (!tir_build_tables(WIN, WIN / 2, 40, 2, samplerate, tab)) return -1;
  auto *sm = new TirSmem<WIN>();
  std::memset(sm, 0, sizeof(*sm));
  std::memcpy(sm->win4, tab.win4.data(), sizeof(sm->win4));
  std::memcpy(sm->twp4, tab.twp4.data(), sizeof(sm->twp4));
  std::memcpy(sm->twu4, tab.twu4.data(), sizeof(sm->twu4));
  const double2 lt[16] = TIR_LOGF_TAB_INIT;
  std::memcpy(sm->logtab, lt, sizeof(lt));
  std::memcpy(sm->w2, tab.mel.w2, sizeof(sm->w2));
  for (int i = 0; i < TIR_MAX_RUNS; i++) sm->run_bins[i] = tab.mel.run_bins[i], sm->run_emit[i] = tab.mel.run_emit[i];
  const int64_t nsamp = (int64_t)n_samples;
  const int64_t nframes = (nsamp + C::HOP - 1) / C::HOP;
  std::vector<TirPass2Regs> regs(C::NT);
  const TirP2 nz = tir_pbc(-0.0f);
  for (int i = 0; i < 2 * TIR_MAX_FILTERS * 32; i++) // as the kernel prologue does
    if ((i / 32) % TIR_MAX_FILTERS < tab.mel.n_filters && tab.mel.dead[(i / 32) % TIR_MAX_FILTERS]) (&sm->lg[0][0])[i] = tab.mel.lg_dead;
  int b = 0;
  for (int64_t f0 = 0; f0 < nframes; f0 += C::T, b ^= 1) {
    const int nvalid = (int)std::min<int64_t>(C::T, nframes - f0);
    // P0 (same placement as tir_issue_tile_load: chunk c, 8-byte unit p -> pcm[c * PCH + p])
    for (int chunk = 0; chunk <= C::T; chunk++)
      for (int p = 0; p < C::HOP / 4; p++) {
        uint32_t h[4];
        for (int e = 0; e < 4; e++) {
          int64_t s = (f0 - 1 + chunk) * C::HOP + 4 * p + e;
          h[e] = (s >= 0 && s < nsamp) ? (uint16_t)pcm[s] : 0;
        }
        sm->pcm[chunk * C::PCH + p].x = h[0] | (h[1] << 16);
        sm->pcm[chunk * C::PCH + p].y = h[2] | (h[3] << 16);
      }
    for (int w = 0; w < C::NW; w++)
      for (int lane = 0; lane < 32; lane++) {
        if constexpr (WIN == 512) tir_pass1_512(*sm, sm->pcm, w, lane, nz);
        else tir_pass1_1024(*sm, sm->pcm, w, lane, nz);
      }
    for (int w = 0; w < C::NW; w++)
      for (int lane = 0; lane < 32; lane++) tir_pass2_load<WIN>(*sm, w, lane, regs[w * 32 + lane]);
    for (int w = 0; w < C::NW; w++)
      for (int lane = 0; lane < 32; lane++) {
        if (w == 0) tir_pass2_compute<WIN, true>(*sm, w, lane, regs[w * 32 + lane], nz);
        else tir_pass2_compute<WIN, false>(*sm, w, lane, regs[w * 32 + lane], nz);
      }
    for (int w = 0; w < C::NW; w++)
      for (int lane = 0; lane < 32; lane++) tir_mel_sweep(sm->xch, sm->lg[b], tab.mel, sm->w2, sm->run_bins, sm->run_emit, w, lane, nz);
    for (int w = 0; w < C::NW; w++)
      for (int lane = 0; lane < 32; lane++) {
        if (tab.mel.live_prefix && !force_indexed_logs)
          tir_log_phase_prefix<C::NW>(sm->lg[b], sm->logtab, tab.mel.log_clamp, tab.mel.n_live, w, lane);
        else tir_log_phase<C::NW>(sm->lg[b], sm->logtab, tab.mel, w, lane);
      }
    for (int j = 0; j < 2; j++)
      for (int lane = 0; lane < nvalid; lane++) {
        float c;
        int32_t v;
        tir_dct_phase(sm->lg[b], tab.mel, j, lane, c, v);
        coef[(f0 + lane) * 2 + j] = c;
        vq[(f0 + lane) * 2 + j] = v;
      }
  }
  delete sm;
  return 0;
}

extern "C" int emul_extract(const int16_t *pcm, uint64_t n_samples, int samplerate, float *coef, int32_t *vq) {
  return emul_extract_t<512>(pcm, n_samples, samplerate, coef, vq);
}
extern "C" int emul_extract_win(int win, const int16_t *pcm, uint64_t n_samples, int samplerate, float *coef,
                                int32_t *vq) {
  if (win == 512) return emul_extract_t<512>(pcm, n_samples, samplerate, coef, vq);
  if (win == 1024) return emul_extract_t<1024>(pcm, n_samples, samplerate, coef, vq);
  return -1;
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
