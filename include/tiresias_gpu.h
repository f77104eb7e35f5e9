/*
 * tiresias_gpu.h -- C ABI of libtiresias_gpu.so, the B200 (sm_100a) implementation of the
 * fingerprint hot path of asterisk-tiresias.
 *
 * The reference has no FFI: its hot path is two in-process seams inside src/fp_handler.c,
 *   (A) create_audio_fingerprints(filename, uuid)            src/fp_handler.c:577-671
 *       libaubio source -> pvoc -> mfcc per hop, then 10*log10(fabs(c)), later "%f"
 *       (src/db_ctx_handler.c:480) on the way into SQLite;
 *   (B) the per-frame SQL probe + tally in fp_search_fingerprint_info()
 *       src/fp_handler.c:285-374, over the rows written by
 *       create_audio_fingerprint_info() src/fp_handler.c:538-575 and removed by
 *       fp_delete_audio_list_info() src/fp_handler.c:147.
 * Each entry point below names the seam it replaces.  Plain C types only; all buffers are owned
 * by the caller; every function returns TIR_OK (0) or a negative tir_status and never aborts.
 * A CUDA failure is reported as TIR_ERR_CUDA -- there is NO CPU fallback.  INTEGRATION.md shows
 * the fp_handler.c patch that binds these.
 *
 * Thread safety: a tir_ctx may be used from any number of host threads (the dialplan runs
 * fp_search_fingerprint_info on one PBX thread per channel); calls are serialised per context.
 *
 * Units: "micro-units" are the integer the reference's "%f" text denotes, v = y * 1e6 rounded the
 * way printf rounds; TIR_NULL_V marks a NULL column (non-finite y never becomes a JSON real).
 */
#ifndef TIRESIAS_GPU_H_
#define TIRESIAS_GPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIR_ABI_VERSION 1
#define TIR_NULL_V INT32_MIN
#define TIR_N_COEFS 2 /* DEF_AUBIO_COEFS, src/fp_handler.c:39 */

typedef enum {
  TIR_OK = 0,
  TIR_ERR_ARG = -1,     /* bad argument (the reference returns NULL/false + LOG_WARNING) */
  TIR_ERR_CUDA = -2,    /* CUDA runtime error, see tir_last_error() */
  TIR_ERR_NOMEM = -3,
  TIR_ERR_STATE = -4,   /* e.g. match before any tir_db_* call */
  TIR_ERR_NOTFOUND = -5 /* tir_db_remove of an unknown uuid */
} tir_status;

typedef struct tir_ctx tir_ctx;

/* The DSP constants are compile-time #defines in the reference (src/fp_handler.c:33-39):
 * win 512 / hop 256 (or the commented 1024 / 512), 40 filters, 2 coefficients; the sample rate
 * is the file's own (DEF_AUBIO_SAMPLERATE 0). */
typedef struct {
  int device;       /* CUDA device ordinal */
  int win;          /* DEF_AUBIO_BUFSIZE : 512 (or 1024)        */
  int hop;          /* DEF_AUBIO_HOPSIZE : win / 2              */
  int n_filters;    /* DEF_AUBIO_FILTER  : 40                   */
  int samplerate;   /* rate the mel filterbank is built for     */
  void *stream;     /* cudaStream_t to run on, NULL = own stream */
} tir_cfg;

void tir_cfg_default(tir_cfg *cfg); /* 512 / 256 / 40 / 8000 Hz, device 0 */

int tir_open(const tir_cfg *cfg, tir_ctx **out); /* fp_init() device side, src/fp_handler.c:68 */
void tir_close(tir_ctx *ctx);                    /* fp_term(),             src/fp_handler.c:92 */
/* text of the last failure on this context; the pointer belongs to the calling thread and stays
 * valid until that thread calls tir_last_error again (other threads may fail meanwhile) */
const char *tir_last_error(tir_ctx *ctx);
int tir_abi_version(void);

/* ---- seam (A): extraction ------------------------------------------------------------------ */

/* frames aubio_source_do delivers for n_samples (loop src/fp_handler.c:632-636): ceil(n/hop) */
uint64_t tir_n_frames(uint64_t n_samples, int hop);

/*
 * create_audio_fingerprints() for a batch of mono PCM16 clips held in HOST memory.
 *   pcm       all clips back to back; clip c = pcm[clip_off[c] .. clip_off[c+1])
 *   coef      [n_frames_total][2] float   mfcc_out->data[i]                      (may be NULL)
 *   vq        [n_frames_total][2] int32   10*log10(fabs(c)) as stored through "%f", micro-units
 *   n_frames  out: frames written (sum over clips of ceil(len/hop)), frames in clip order
 * Copies pcm host->device, runs the fused kernel, copies the results back (all on ctx's stream)
 * and returns when they are in the caller's buffers.
 */
int tir_extract(tir_ctx *ctx, const int16_t *pcm, const uint64_t *clip_off, uint32_t n_clips, float *coef,
                int32_t *vq, uint64_t *n_frames);

/* Same for clips held as G.711 mu-law bytes (a channel's native ulaw frames, or a .ulaw/.pcm file):
 * decoded to PCM16 on the device with the standard table (== Asterisk's AST_MULAW), i.e. exactly
 * tir_extract of the decoded samples, for half the host->device traffic. */
int tir_extract_ulaw(tir_ctx *ctx, const uint8_t *ulaw, const uint64_t *clip_off, uint32_t n_clips, float *coef,
                     int32_t *vq, uint64_t *n_frames);

/* Same for clips with `channels` interleaved PCM16 channels (a stereo WAV's data chunk as it lies in the file):
 * aubio_source_do (src/fp_handler.c:604,612,633) hands the module the float mean of the channels -- every
 * sample / 32768 in float, summed in channel order from 0.f, divided by the channel count (aubio source_wavread.c /
 * source_sndfile.c) -- reproduced on the device rounding for rounding; the hop loop then runs on those floats.
 *   pcm       [sample frame][channel]; clip c = sample frames clip_off[c] .. clip_off[c+1]
 * channels == 1 is tir_extract. */
int tir_extract_interleaved(tir_ctx *ctx, const int16_t *pcm, int channels, const uint64_t *clip_off, uint32_t n_clips,
                            float *coef, int32_t *vq, uint64_t *n_frames);

/* Same with DEVICE buffers (d_pcm, d_coef, d_vq); clip_off stays a host array.  Asynchronous on
 * ctx's stream. */
int tir_extract_dev(tir_ctx *ctx, const int16_t *d_pcm, const uint64_t *clip_off, uint32_t n_clips,
                    float *d_coef, int32_t *d_vq, uint64_t *n_frames);

/* Diagnostics of the two float primitives the bit-exactness of the extraction rests on.
 *  sqrt_mismatches (may be NULL): the kernel's branch-free square root against the IEEE one for every
 *    float in [2^-149, 2^60) -- must come back 0;
 *  log10f_out (may be NULL): the DEVICE build of the glibc-exact log10f for the `count` floats with bit
 *    patterns first_bits + i*step, for comparison with libm on the host. */
int tir_selftest(tir_ctx *ctx, uint64_t *sqrt_mismatches, uint32_t first_bits, uint32_t step, uint32_t count,
                 float *log10f_out);

/* the plan's aubio-layout tables, for table-level parity tests (host copies) */
int tir_get_tables(tir_ctx *ctx, float *window /*[win]*/, float *filters /*[n_filters][win/2+1]*/,
                   float *dct /*[2][n_filters]*/);

/* ---- seam (B): the device-resident mirror of table audio_fingerprint ------------------------- */

/* Bulk load (fp_init's restore of the backup DB, src/fp_handler.c:82-88): audio a has uuid
 * uuid[a] (16 raw bytes; canonical lower-case text order == byte order) and rows
 * v1/v2[row_off[a] .. row_off[a+1]) in frame_idx order.  Replaces the current contents. */
int tir_db_load(tir_ctx *ctx, uint32_t n_audio, const uint8_t (*uuid)[16], const uint64_t *row_off,
                const int32_t *v1, const int32_t *v2);
/* Same from DEVICE arrays (uuid bytes, row offsets, micro-unit values already on this GPU, e.g. the
 * output of tir_extract_dev); the library keeps its own copy. */
int tir_db_load_dev(tir_ctx *ctx, uint32_t n_audio, const uint8_t (*d_uuid)[16], const uint64_t *d_row_off,
                    const int32_t *d_v1, const int32_t *d_v2, uint64_t n_rows);
/* create_audio_fingerprint_info(): the rows of one new audio, src/fp_handler.c:538-575 (no full re-sort: see
 * tir_db_index_stats) */
int tir_db_add(tir_ctx *ctx, const uint8_t uuid[16], const int32_t *v1, const int32_t *v2, uint32_t n_rows);
/* delete from audio_fingerprint where audio_uuid=..., src/fp_handler.c:147 */
int tir_db_remove(tir_ctx *ctx, const uint8_t uuid[16]);
int tir_db_stats(tir_ctx *ctx, uint64_t *n_audio, uint64_t *n_rows);
/* How the index follows tir_db_add / tir_db_remove: an added audio goes into a small TAIL index (re-sorted in about
 * a millisecond by the next search; its winners are folded with the main index's like those of a second shard), a
 * removed audio gets a tombstone that the kernels consult where they pick a block's winners.  The full sort of the
 * table runs only after a load, or when the tail holds more than max(2^20, rows/16) rows or a quarter of the main
 * index is dead; it first drops the removed audios from the device copy of the table (no growth under churn) and sorts
 * in passes of ~64 M rows (scratch: 2 GB whatever the table's size).  Counters: full sorts so far, tail sorts, audios
 * in the tail, tombstones. */
int tir_db_index_stats(tir_ctx *ctx, uint64_t *n_full_builds, uint64_t *n_tail_builds, uint64_t *tail_audios, uint64_t *tombstones);
/* A caller that keeps searching with the same buffers, batch shape and parameters (the batcher, the stream pump, a
 * benchmark loop) has its match chain -- the copy of the query offsets, the clearing of the scratch, the four kernels --
 * captured once into a CUDA graph and replayed with ONE launch per batch; any change of an argument, of the index or of
 * the table re-captures.  Counters: batches served by a graph launch, graphs captured so far. */
int tir_match_graph_stats(tir_ctx *ctx, uint64_t *n_graph_launches, uint64_t *n_graphs_built);

/* ---- SQLite <-> device table -----------------------------------------------------------------
 * SQLite stays the system of record (schema src/fp_handler.c:686-753).  `sqlite3_db` is the
 * module's own connection (g_db_ctx->db, src/fp_handler.c:45); libsqlite3 is resolved at run time.
 *
 * tir_db_load_sqlite: mirror table audio_fingerprint into the device table, e.g. right after
 * fp_init() restored the backup (src/fp_handler.c:82-88).  Rows whose audio_uuid is not a canonical
 * uuid text are counted in n_skipped.  *_file opens a database file read-only instead. */
int tir_db_load_sqlite(tir_ctx *ctx, void *sqlite3_db, uint64_t *n_audio, uint64_t *n_rows, uint64_t *n_skipped);
int tir_db_load_sqlite_file(tir_ctx *ctx, const char *path, uint64_t *n_audio, uint64_t *n_rows, uint64_t *n_skipped);
/* create_audio_fingerprint_info() (src/fp_handler.c:538-575): the frames of one audio into table
 * audio_fingerprint in one transaction through one prepared statement; stores exactly what the
 * reference's textual INSERTs store ("%f" values, NULL for TIR_NULL_V).  vq is [n_frames][2]
 * micro-units as written by tir_extract. */
int tir_sqlite_insert_fingerprints(tir_ctx *ctx, void *sqlite3_db, const char *context, const char *audio_uuid,
                                   const int32_t *vq, uint32_t n_frames);

/* ---- seam (B): match ------------------------------------------------------------------------- */

typedef struct {
  uint8_t uuid[16];     /* winning audio_uuid (greatest uuid among equal counts, as SQLite 3.45.1
                           orders "group by audio_uuid order by count(*) DESC") */
  int32_t match_count;  /* count(*) : query frames with >=1 row of that uuid in their window;
                           0 => no row matched at all (reference returns NULL / NOTFOUND) */
  int32_t frame_count;  /* all query frames, ignored ones included (src/fp_handler.c:286,403) */
} tir_hit;

/*
 * The probe/tally block of fp_search_fingerprint_info() (src/fp_handler.c:285-374) for a batch of
 * queries whose frames were already extracted.
 *   y          [n_frames_total][2] double : max1,max2 of every query frame (query q owns frames
 *              frame_off[q] .. frame_off[q+1]); NaN = JSON key missing (read back as 0.0)
 *   coefs      1 or 2 (src/fp_handler.c:247); tolerance < 0 -> 0.001 (:252-256)
 *   freq_ignore_low/high  <= 0 disables the filter (:293-306, :324-337)
 * Errors: coefs outside [1,2] -> TIR_ERR_ARG.
 */
int tir_match(tir_ctx *ctx, const double *y, const uint64_t *frame_off, uint32_t n_queries, int coefs,
              double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *hits);

/* Same from device-resident mfcc coefficients as written by tir_extract_dev (y is recomputed as
 * 10*log10(fabs((double)c)) on the device); d_hits is a device array of n_queries tir_hit. */
int tir_match_dev(tir_ctx *ctx, const float *d_coef, const uint64_t *frame_off, uint32_t n_queries, int coefs,
                  double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *d_hits);

/* fp_search_fingerprint_info() minus the audio_list lookup: extract + match for a batch of query
 * clips in HOST memory (what each dialplan Tiresias() call does for its recording). */
int tir_search(tir_ctx *ctx, const int16_t *pcm, const uint64_t *clip_off, uint32_t n_clips, int coefs,
               double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *hits);

/* ---- concurrent callers ------------------------------------------------------------------------
 * The dialplan application runs fp_search_fingerprint_info() on one PBX thread per channel
 * (src/application_handler.c:180); in the reference those threads serialise on the one SQLite
 * connection (src/fp_handler.c:45).  With a batcher running, concurrent tir_search_one() callers
 * that share (coefs, tolerance, freq_ignore_*) are served by ONE batched tir_search: a caller waits
 * at most max_wait_us for company, a batch holds at most max_batch recordings.  Without a batcher
 * tir_search_one() is a batch of one.  Results are those of tir_search on the same recording. */
int tir_batcher_start(tir_ctx *ctx, uint32_t max_batch, uint32_t max_wait_us);
int tir_batcher_stop(tir_ctx *ctx); /* serves the queued requests, then stops; tir_close() does it too */
int tir_search_one(tir_ctx *ctx, const int16_t *pcm, uint64_t n_samples, int coefs, double tolerance,
                   int freq_ignore_low, int freq_ignore_high, tir_hit *hit);
int tir_batcher_stats(tir_ctx *ctx, uint64_t *n_requests, uint64_t *n_batches, uint64_t *max_batch_seen);

/* Streaming recording: replaces record_voice()'s /tmp/tiresias-<uuid>.wav round trip
 * (src/application_handler.c:153-155,248-312).  Feed the channel's slinear frames as ast_read delivers
 * them: the hop loop (src/fp_handler.c:632-661) runs WHILE the call is recorded -- a pump thread of the
 * context gathers, every TIR_STREAM_PERIOD_US (default 2000) microseconds, the completed hops of all live
 * streams into one batched extraction launch; a stream keeps one hop of PCM as state and its coefficients on
 * the device.  tir_stream_finish() adds the final zero-padded hop and waits for ONE batched match over the
 * finishing streams of equal parameters: the time from the last feed to the result does not depend on the
 * length of the recording.  Results equal tir_search_one() on everything fed.  One feeder thread per
 * stream; any number of streams.  tir_stream_frames_done: frames already extracted (monitoring, tests). */
typedef struct tir_stream tir_stream;
int tir_stream_open(tir_ctx *ctx, tir_stream **out);
int tir_stream_feed(tir_stream *s, const int16_t *pcm, uint32_t n_samples);
uint64_t tir_stream_samples(const tir_stream *s);
uint64_t tir_stream_frames_done(const tir_stream *s);
int tir_stream_finish(tir_stream *s, int coefs, double tolerance, int freq_ignore_low, int freq_ignore_high,
                      tir_hit *hit);
int tir_stream_stats(tir_ctx *ctx, uint64_t *n_extract_batches, uint64_t *n_frames, uint64_t *n_match_batches);
void tir_stream_close(tir_stream *s);

/* Multi-GPU (DB sharded by uuid, one context per GPU): fold the per-shard winners of the same
 * queries into the global winner.  d_gathered is [n_shards][n_queries] tir_hit on this device
 * (e.g. the output of an NCCL all-gather of each rank's d_hits); result in d_out[n_queries]. */
int tir_merge_hits_dev(tir_ctx *ctx, const tir_hit *d_gathered, uint32_t n_shards, uint32_t n_queries,
                       tir_hit *d_out);

/* The same exchange WITHOUT a collective library: every rank stores its winners straight into every
 * peer's gather buffer over NVLink peer memory, raises a per-peer flag carrying the batch number, and
 * folds the candidates of all ranks once their flags have arrived (csrc/tir_p2p.cu).  One process per
 * GPU: create, exchange the 64-byte handles (tir_p2p_handle) through the launcher, tir_p2p_connect
 * with all of them in rank order.  Several contexts of one process: tir_p2p_connect_local.
 * tir_p2p_match_dev = tir_match_dev on the local shard + that exchange; d_final receives the global
 * winners (folded by the last CTA of the kernel that produced the local ones).  SPMD: all ranks call it
 * the same number of times.  tir_p2p_error reports (after a stream synchronise) the batch at which a
 * merge gave up waiting for a peer, 0 if none; such a batch leaves match_count = -1 in d_final.  A rank
 * whose local match fails still publishes "no winner" rows so that its peers complete. */
typedef struct tir_p2p tir_p2p;
int tir_p2p_create(tir_ctx *ctx, int rank, int world, uint32_t max_queries, tir_p2p **out);
int tir_p2p_handle(tir_p2p *p, unsigned char handle[64]);
int tir_p2p_connect(tir_p2p *p, const unsigned char *handles /* world x 64 bytes, rank order */);
int tir_p2p_connect_local(tir_p2p *p, tir_p2p *const *all /* world entries, rank order */);
int tir_p2p_match_dev(tir_p2p *p, const float *d_coef, const uint64_t *frame_off, uint32_t n_queries, int coefs,
                      double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *d_final);
int tir_p2p_error(tir_p2p *p, uint32_t *epoch_out);
/* Sharded SEARCH (fp_search_fingerprint_info for a batch, src/fp_handler.c:207-408, over N GPUs): the
 * queries are sharded as well as the table.  Rank r brings only ITS OWN slice of the batch's clips --
 * queries [first_query, first_query + n_local) of n_total, pcm/clip_off in host memory -- uploads and
 * extracts that slice; the extraction kernel stores every coefficient it produces into all ranks'
 * coefficient buffers over NVLink (8 bytes per frame and rank) and its last CTA raises a flag; every
 * rank then matches ALL n_total queries against its shard, and the kernel that produces the winners
 * exchanges them and folds the ranks' candidates (no collective call, no separate merge launch).
 * all_frame_off[n_total + 1] are the frame offsets of the whole batch (every rank knows every clip's
 * length: ceil(samples / hop) frames each).  Results: hits (host, [n_total], may be NULL) and/or
 * d_final (device, may be NULL), identical on every rank.  SPMD like tir_p2p_match_dev; needs
 * tir_p2p_create2 with max_frames >= all_frame_off[n_total].  A merge or wait that timed out leaves
 * match_count = -1 in every hit of the batch and is reported by tir_p2p_error. */
int tir_p2p_create2(tir_ctx *ctx, int rank, int world, uint32_t max_queries, uint64_t max_frames, tir_p2p **out);
/* Pre-size every scratch buffer the context needs for batches of up to max_queries queries / max_frames
 * frames / max_local_samples own samples (and rebuild a dirty index now), so that the exchange calls neither
 * allocate nor free.  REQUIRED when several ranks are driven from ONE host thread (contexts of one process):
 * cudaFree waits for the whole device, and an earlier rank's kernel may be waiting for a rank the thread has
 * not enqueued yet.  One process or thread per rank does not need it. */
int tir_p2p_reserve(tir_p2p *p, uint64_t max_local_samples);
int tir_p2p_search(tir_p2p *p, const int16_t *pcm, const uint64_t *clip_off, uint32_t n_local, uint32_t first_query,
                   const uint64_t *all_frame_off, uint32_t n_total, int coefs, double tolerance, int freq_ignore_low,
                   int freq_ignore_high, tir_hit *hits, tir_hit *d_final);
void tir_p2p_destroy(tir_p2p *p);

/* The same inside ONE process (an Asterisk module cannot be launched one process per GPU): a group
 * owns one context per device and shards the table by uuid over them.  A search is tir_p2p_search driven
 * from this process: the batch's clips are cut into one slice per device, every device uploads and extracts
 * its slice, coefficients and winners cross NVLink inside the kernels (peer stores + flags).  Devices that
 * cannot reach each other take a copy-based path (extraction on the first device, cudaMemcpyPeerAsync).
 * Results equal those of one context holding the whole table. */
typedef struct tir_group tir_group;
int tir_group_open(const tir_cfg *cfg, const int *devices, int n_devices, tir_group **out); /* cfg->device/stream ignored */
void tir_group_close(tir_group *g);
const char *tir_group_last_error(tir_group *g);
int tir_group_size(tir_group *g);
tir_ctx *tir_group_ctx(tir_group *g, int i); /* borrowed: e.g. tir_extract on a particular device */
int tir_group_db_load(tir_group *g, uint32_t n_audio, const uint8_t (*uuid)[16], const uint64_t *row_off,
                      const int32_t *v1, const int32_t *v2);
int tir_group_db_add(tir_group *g, const uint8_t uuid[16], const int32_t *v1, const int32_t *v2, uint32_t n_rows);
int tir_group_db_remove(tir_group *g, const uint8_t uuid[16]);
int tir_group_db_stats(tir_group *g, uint64_t *n_audio, uint64_t *n_rows);
int tir_group_search(tir_group *g, const int16_t *pcm, const uint64_t *clip_off, uint32_t n_clips, int coefs,
                     double tolerance, int freq_ignore_low, int freq_ignore_high, tir_hit *hits);

/* how many searches took the fused NVLink path / the copy path (no peer access between the devices) */
int tir_group_stats(tir_group *g, uint64_t *n_fused, uint64_t *n_copy_path);
/* the concurrent front-end on a group (BASELINE config[4]: 1 000 dialplan channels, 8 GPUs): like
 * tir_batcher_start / tir_search_one, every sealed batch is one tir_group_search */
int tir_group_batcher_start(tir_group *g, uint32_t max_batch, uint32_t max_wait_us);
int tir_group_batcher_stop(tir_group *g);
int tir_group_search_one(tir_group *g, const int16_t *pcm, uint64_t n_samples, int coefs, double tolerance,
                         int freq_ignore_low, int freq_ignore_high, tir_hit *hit);
int tir_group_batcher_stats(tir_group *g, uint64_t *n_requests, uint64_t *n_batches, uint64_t *max_batch_seen);

/* which shard (0..n_shards-1) owns a uuid */
uint32_t tir_shard_of(const uint8_t uuid[16], uint32_t n_shards);

/* bookkeeping for the benches: kernels launched by this context so far; with profiling switched on
 * the main extraction / match kernels are bracketed by CUDA events on ctx's stream and
 * tir_last_kernel_ms() returns the device time of the most recent one (which: 0 = extraction
 * kernel, 1 = match kernel; waits for it to finish; < 0 if none was recorded). */
uint64_t tir_launch_count(tir_ctx *ctx);
void tir_set_profiling(tir_ctx *ctx, int on);
float tir_last_kernel_ms(tir_ctx *ctx, int which);

#ifdef __cplusplus
}
#endif
#endif /* TIRESIAS_GPU_H_ */
