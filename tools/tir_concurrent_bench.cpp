// tir_concurrent_bench.cpp -- BASELINE config[4]: N concurrent dialplan channels, each a host thread
// blocking in the fp_search_fingerprint_info() replacement (tir_search_one through the batcher),
// against a synthetic fingerprint DB.  Plain C++ over the C ABI of libtiresias_gpu.so; prints one JSON
// object: queries/s, latency percentiles, batching statistics.
//   tir_concurrent_bench [--threads 1000] [--rounds 5] [--db-fps 1000000] [--frames 94]
//                        [--max-batch 1024] [--wait-us 300] [--seconds 3] [--device 0]
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/tiresias_gpu.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd() {
  uint64_t x = rng_state;
  x ^= x << 13, x ^= x >> 7, x ^= x << 17;
  return rng_state = x;
}

int main(int argc, char **argv) {
  int threads = 1000, rounds = 5, frames = 94, max_batch = 1024, wait_us = 300, device = 0;
  long db_fps = 1000000;
  double seconds = 3.0;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string a = argv[i];
    if (a == "--threads") threads = atoi(argv[i + 1]);
    else if (a == "--rounds") rounds = atoi(argv[i + 1]);
    else if (a == "--db-fps") db_fps = atol(argv[i + 1]);
    else if (a == "--frames") frames = atoi(argv[i + 1]);
    else if (a == "--max-batch") max_batch = atoi(argv[i + 1]);
    else if (a == "--wait-us") wait_us = atoi(argv[i + 1]);
    else if (a == "--seconds") seconds = atof(argv[i + 1]);
    else if (a == "--device") device = atoi(argv[i + 1]);
  }
  tir_cfg cfg;
  tir_cfg_default(&cfg);
  cfg.device = device;
  tir_ctx *ctx = nullptr;
  if (tir_open(&cfg, &ctx) != TIR_OK) {
    fprintf(stderr, "tir_open: %s\n", tir_last_error(ctx));
    return 1;
  }
  // synthetic DB: max1 ~ U(15.5, 18.5), max2 ~ U(-5, 20) in micro-units (ranges of SURVEY.md 8a)
  const uint64_t rows = (uint64_t)db_fps * frames;
  std::vector<uint8_t> uu((size_t)db_fps * 16);
  std::vector<uint64_t> off((size_t)db_fps + 1);
  std::vector<int32_t> v1(rows), v2(rows);
  for (size_t i = 0; i < uu.size(); i += 8) {
    uint64_t r = rnd();
    memcpy(&uu[i], &r, 8);
  }
  for (long a = 0; a <= db_fps; a++) off[a] = (uint64_t)a * frames;
  for (uint64_t r = 0; r < rows; r++) {
    const uint64_t x = rnd();
    v1[r] = 15500000 + (int32_t)((x & 0xffffffffu) % 3000000u);
    v2[r] = -5000000 + (int32_t)((x >> 32) % 25000000u);
  }
  if (tir_db_load(ctx, (uint32_t)db_fps, (const uint8_t(*)[16])uu.data(), off.data(), v1.data(), v2.data()) != TIR_OK) {
    fprintf(stderr, "tir_db_load: %s\n", tir_last_error(ctx));
    return 1;
  }
  // one recording per channel: tone + noise, 8 kHz
  const int n = (int)(seconds * 8000);
  std::vector<std::vector<int16_t>> clips(threads, std::vector<int16_t>(n));
  for (int t = 0; t < threads; t++) {
    const double f = 200.0 + (double)(rnd() % 3200), amp = 2000.0 + (double)(rnd() % 20000);
    for (int i = 0; i < n; i++)
      clips[t][i] = (int16_t)(amp * sin(2 * M_PI * f * i / 8000.0) + (double)((int)(rnd() % 2001) - 1000));
  }
  if (tir_batcher_start(ctx, (uint32_t)max_batch, (uint32_t)wait_us) != TIR_OK) {
    fprintf(stderr, "tir_batcher_start: %s\n", tir_last_error(ctx));
    return 1;
  }
  { // warm-up: buffers, index
    tir_hit h;
    for (int i = 0; i < 3; i++) tir_search_one(ctx, clips[0].data(), n, 1, 0.001, -1, -1, &h);
  }
  std::vector<double> lat((size_t)threads * rounds);
  std::atomic<int> errors{0}, found{0}, warmed{0};
  std::atomic<bool> go{false};
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++)
    th.emplace_back([&, t] {
      { // untimed warm-up round with every channel: staging and device buffers reach their working size
        tir_hit h;
        tir_search_one(ctx, clips[t].data(), n, 1, 0.001, -1, -1, &h);
        warmed++;
      }
      while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
      for (int r = 0; r < rounds; r++) {
        tir_hit h;
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = tir_search_one(ctx, clips[t].data(), n, 1, 0.001, -1, -1, &h);
        const auto t1 = std::chrono::steady_clock::now();
        lat[(size_t)t * rounds + r] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (rc != TIR_OK) errors++;
        else if (h.match_count > 0) found++;
      }
    });
  while (warmed.load() < threads) std::this_thread::yield();
  uint64_t nreq0 = 0, nbatch0 = 0, maxb0 = 0;
  tir_batcher_stats(ctx, &nreq0, &nbatch0, &maxb0);
  const auto w0 = std::chrono::steady_clock::now();
  go.store(true, std::memory_order_release);
  for (auto &x : th) x.join();
  const auto w1 = std::chrono::steady_clock::now();
  const double wall = std::chrono::duration<double>(w1 - w0).count();
  uint64_t nreq = 0, nbatch = 0, maxb = 0;
  tir_batcher_stats(ctx, &nreq, &nbatch, &maxb);
  std::sort(lat.begin(), lat.end());
  auto pct = [&](double p) { return lat[std::min(lat.size() - 1, (size_t)(p * lat.size()))]; };
  printf("{\"threads\": %d, \"rounds\": %d, \"queries\": %zu, \"queries_per_s\": %.1f, \"wall_s\": %.4f, "
         "\"latency_ms\": {\"p50\": %.3f, \"p90\": %.3f, \"p99\": %.3f, \"max\": %.3f}, \"batches\": %llu, "
         "\"mean_batch\": %.1f, \"max_batch_seen\": %llu, \"max_batch\": %d, \"max_wait_us\": %d, \"db_fingerprints\": %ld, "
         "\"db_frames_per_fingerprint\": %d, \"seconds_per_query_clip\": %.1f, \"errors\": %d, \"found\": %d, "
         "\"api\": \"tir_search_one (one host thread per channel, batcher on)\"}\n",
         threads, rounds, lat.size(), lat.size() / wall, wall, pct(0.5), pct(0.9), pct(0.99), lat.back(),
         (unsigned long long)(nbatch - nbatch0), nbatch > nbatch0 ? (double)(nreq - nreq0) / (double)(nbatch - nbatch0) : 0.0,
         (unsigned long long)maxb, max_batch, wait_us, db_fps, frames, seconds, errors.load(), found.load());
  tir_close(ctx);
  return errors.load() ? 2 : 0;
}
