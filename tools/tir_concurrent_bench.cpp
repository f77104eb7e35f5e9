// tir_concurrent_bench.cpp -- BASELINE config[4]: N concurrent dialplan channels, each a host thread
// blocking in the fp_search_fingerprint_info() replacement (tir_search_one through the batcher),
// against a synthetic fingerprint DB.  Plain C++ over the C ABI of libtiresias_gpu.so; prints one JSON
// object: queries/s, latency percentiles, batching statistics.
//   tir_concurrent_bench [--threads 1000] [--rounds 5] [--db-fps 1000000] [--frames 94]
//                        [--max-batch 1024] [--wait-us 300] [--seconds 3] [--device 0]
//                        [--devices N]  N > 1: one process, N GPUs (tir_group_*: table sharded by uuid, tir_group_search_one)
//                        [--stream 1]   feed every recording in 20 ms chunks through tir_stream_* (one device)
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
  int threads = 1000, rounds = 5, frames = 94, max_batch = 1024, wait_us = 300, device = 0, devices = 1, stream = 0;
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
    else if (a == "--devices") devices = atoi(argv[i + 1]);
    else if (a == "--stream") stream = atoi(argv[i + 1]);
  }
  tir_cfg cfg;
  tir_cfg_default(&cfg);
  cfg.device = device;
  tir_ctx *ctx = nullptr;
  tir_group *grp = nullptr;
  std::vector<tir_ctx *> shards;
  if (devices > 1) {
    std::vector<int> devs(devices);
    for (int i = 0; i < devices; i++) devs[i] = i;
    if (tir_group_open(&cfg, devs.data(), devices, &grp) != TIR_OK) {
      fprintf(stderr, "tir_group_open: %s\n", tir_group_last_error(grp));
      return 1;
    }
    for (int i = 0; i < devices; i++) shards.push_back(tir_group_ctx(grp, i));
  } else {
    if (tir_open(&cfg, &ctx) != TIR_OK) {
      fprintf(stderr, "tir_open: %s\n", tir_last_error(ctx));
      return 1;
    }
    shards.push_back(ctx);
  }
  // synthetic DB: max1 ~ U(15.5, 18.5), max2 ~ U(-5, 20) in micro-units (ranges of SURVEY.md 8a); one slice of the
  // fingerprints per device (any partition by uuid gives the same winners), generated and loaded in parallel
  {
    std::vector<std::thread> loaders;
    std::atomic<int> load_err{0};
    const int S = (int)shards.size();
    for (int sh = 0; sh < S; sh++)
      loaders.emplace_back([&, sh] {
        const long n_loc = db_fps / S + (sh < db_fps % S ? 1 : 0);
        const uint64_t rows = (uint64_t)n_loc * frames;
        uint64_t st = 0x9E3779B97F4A7C15ull * (uint64_t)(sh + 1) + 12345;
        auto rn = [&st]() { st ^= st << 13, st ^= st >> 7, st ^= st << 17; return st; };
        std::vector<uint8_t> uu((size_t)n_loc * 16);
        std::vector<uint64_t> off((size_t)n_loc + 1);
        std::vector<int32_t> v1(rows), v2(rows);
        for (size_t i = 0; i < uu.size(); i += 8) {
          uint64_t r = rn();
          memcpy(&uu[i], &r, 8);
        }
        for (long a = 0; a <= n_loc; a++) off[a] = (uint64_t)a * frames;
        for (uint64_t r = 0; r < rows; r++) {
          const uint64_t x = rn();
          v1[r] = 15500000 + (int32_t)((x & 0xffffffffu) % 3000000u);
          v2[r] = -5000000 + (int32_t)((x >> 32) % 25000000u);
        }
        if (tir_db_load(shards[sh], (uint32_t)n_loc, (const uint8_t(*)[16])uu.data(), off.data(), v1.data(), v2.data()) != TIR_OK) {
          fprintf(stderr, "tir_db_load: %s\n", tir_last_error(shards[sh]));
          load_err++;
        }
      });
    for (auto &t : loaders) t.join();
    if (load_err.load()) return 1;
  }
  // one recording per channel: tone + noise, 8 kHz
  const int n = (int)(seconds * 8000);
  std::vector<std::vector<int16_t>> clips(threads, std::vector<int16_t>(n));
  for (int t = 0; t < threads; t++) {
    const double f = 200.0 + (double)(rnd() % 3200), amp = 2000.0 + (double)(rnd() % 20000);
    for (int i = 0; i < n; i++)
      clips[t][i] = (int16_t)(amp * sin(2 * M_PI * f * i / 8000.0) + (double)((int)(rnd() % 2001) - 1000));
  }
  if (grp ? tir_group_batcher_start(grp, (uint32_t)max_batch, (uint32_t)wait_us) != TIR_OK
          : tir_batcher_start(ctx, (uint32_t)max_batch, (uint32_t)wait_us) != TIR_OK) {
    fprintf(stderr, "batcher start: %s\n", grp ? tir_group_last_error(grp) : tir_last_error(ctx));
    return 1;
  }
  // one call of the fp_search_fingerprint_info() replacement, as a channel's PBX thread makes it
  auto search_one = [&](const int16_t *pcm, int len, tir_hit *h) -> int {
    if (grp) return tir_group_search_one(grp, pcm, (uint64_t)len, 1, 0.001, -1, -1, h);
    if (!stream) return tir_search_one(ctx, pcm, (uint64_t)len, 1, 0.001, -1, -1, h);
    tir_stream *st = nullptr; // the recording arrives as 20 ms slinear frames (160 samples at 8 kHz)
    int rc = tir_stream_open(ctx, &st);
    for (int a = 0; rc == TIR_OK && a < len; a += 160) rc = tir_stream_feed(st, pcm + a, (uint32_t)std::min(160, len - a));
    if (rc == TIR_OK) rc = tir_stream_finish(st, 1, 0.001, -1, -1, h);
    tir_stream_close(st);
    return rc;
  };
  { // warm-up: buffers, index
    tir_hit h;
    for (int i = 0; i < 3; i++) search_one(clips[0].data(), n, &h);
  }
  std::vector<double> lat((size_t)threads * rounds);
  std::atomic<int> errors{0}, found{0}, warmed{0};
  std::atomic<bool> go{false};
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++)
    th.emplace_back([&, t] {
      { // untimed warm-up round with every channel: staging and device buffers reach their working size
        tir_hit h;
        search_one(clips[t].data(), n, &h);
        warmed++;
      }
      while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
      for (int r = 0; r < rounds; r++) {
        tir_hit h;
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = search_one(clips[t].data(), n, &h);
        const auto t1 = std::chrono::steady_clock::now();
        lat[(size_t)t * rounds + r] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (rc != TIR_OK) {
          if (errors++ == 0) fprintf(stderr, "first error: rc %d: %s\n", rc, grp ? tir_group_last_error(grp) : tir_last_error(ctx));
        } else if (h.match_count > 0) found++;
      }
    });
  while (warmed.load() < threads) std::this_thread::yield();
  uint64_t nreq0 = 0, nbatch0 = 0, maxb0 = 0;
  auto bstats = [&](uint64_t *a, uint64_t *b, uint64_t *c) {
    if (grp) tir_group_batcher_stats(grp, a, b, c);
    else if (stream) tir_stream_stats(ctx, b, a, c), *a = 0; // (extraction batches / frames / match batches)
    else tir_batcher_stats(ctx, a, b, c);
  };
  bstats(&nreq0, &nbatch0, &maxb0);
  const auto w0 = std::chrono::steady_clock::now();
  go.store(true, std::memory_order_release);
  for (auto &x : th) x.join();
  const auto w1 = std::chrono::steady_clock::now();
  const double wall = std::chrono::duration<double>(w1 - w0).count();
  uint64_t nreq = 0, nbatch = 0, maxb = 0;
  bstats(&nreq, &nbatch, &maxb);
  std::sort(lat.begin(), lat.end());
  auto pct = [&](double p) { return lat[std::min(lat.size() - 1, (size_t)(p * lat.size()))]; };
  printf("{\"threads\": %d, \"rounds\": %d, \"queries\": %zu, \"queries_per_s\": %.1f, \"wall_s\": %.4f, "
         "\"latency_ms\": {\"p50\": %.3f, \"p90\": %.3f, \"p99\": %.3f, \"max\": %.3f}, \"batches\": %llu, "
         "\"mean_batch\": %.1f, \"max_batch_seen\": %llu, \"max_batch\": %d, \"max_wait_us\": %d, \"db_fingerprints\": %ld, "
         "\"db_frames_per_fingerprint\": %d, \"seconds_per_query_clip\": %.1f, \"errors\": %d, \"found\": %d, "
         "\"devices\": %d, \"fused_group_searches\": %llu, \"api\": \"%s\"}\n",
         threads, rounds, lat.size(), lat.size() / wall, wall, pct(0.5), pct(0.9), pct(0.99), lat.back(),
         (unsigned long long)(nbatch - nbatch0), nbatch > nbatch0 ? (double)(nreq - nreq0) / (double)(nbatch - nbatch0) : 0.0,
         (unsigned long long)maxb, max_batch, wait_us, db_fps, frames, seconds, errors.load(), found.load(), devices,
         (unsigned long long)[&] { uint64_t f = 0, c = 0; if (grp) tir_group_stats(grp, &f, &c); return f; }(),
         grp ? "tir_group_search_one (one process, one host thread per channel, table sharded over the devices, batcher on)"
             : stream ? "tir_stream_open/feed/finish (20 ms chunks, hop loop while recording, batched finish)"
                      : "tir_search_one (one host thread per channel, batcher on)");
  if (grp) tir_group_close(grp);
  else tir_close(ctx);
  return errors.load() ? 2 : 0;
}
