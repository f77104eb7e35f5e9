/*
 * fp_handler.c -- drop-in replacement of the reference's src/fp_handler.c: same file name, the 13
 * prototypes of src/fp_handler.h:13-38 byte for byte, same SQLite schema (src/fp_handler.c:686-753), same
 * JSON results and ownership -- only the bodies of the two hot seams change:
 *
 *   create_audio_fingerprints()      src/fp_handler.c:577-671  libaubio hop loop  -> tir_extract
 *   the probe/tally SQL of search    src/fp_handler.c:285-374  per-frame INSERT..SELECT -> tir_search_one
 *
 * libtiresias_gpu.so (include/tiresias_gpu.h) does both on a B200; there is no CPU fallback: if the GPU
 * cannot be opened fp_init() fails and the module declines to load (src/app_tiresias.c:585-589).
 *
 * SQLite stays the system of record -- tables context_list / audio_list / audio_fingerprint in the module's
 * ":memory:" database, restored from and backed up to DEF_BACKUP_DATABASE exactly as before, through the
 * reference's own unchanged src/db_ctx_handler.c -- and the device holds a mirror of audio_fingerprint that
 * follows every insert and delete.  Build: replace src/fp_handler.c by this file, add -I<repo>/include and
 * link -ltiresias_gpu instead of -laubio.  tests/fake_asterisk compiles it together with the
 * reference's unchanged application_handler.c, cli_handler.c, app_tiresias.c and db_ctx_handler.c.
 */
#include <asterisk.h>
#include <asterisk/logger.h>
#include <asterisk/utils.h>
#include <asterisk/json.h>

#include <libgen.h>
#include <math.h>
#include <openssl/evp.h>
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <uuid/uuid.h>

#include "app_tiresias.h"
#include "db_ctx_handler.h"
#include "fp_handler.h"
#include "tiresias_gpu.h"

#define DEF_DATABASE_NAME ":memory:"
#define DEF_BACKUP_DATABASE "/var/lib/asterisk/third-party/tiresias/audio_recongition.db" /* [sic] the on-disk contract */

/* the DSP plan: compile-time constants in the reference too (src/fp_handler.c:33-39) */
#define DEF_AUBIO_HOPSIZE 256
#define DEF_AUBIO_BUFSIZE 512
#define DEF_AUBIO_FILTER 40
#define DEF_AUBIO_COEFS 2
#define DEF_SEARCH_TOLERANCE 0.001
#define DEF_UUID_STR_LEN 37

/* concurrent Tiresias() calls (one PBX thread per channel) are served in batches */
#define DEF_BATCH_MAX 1024
#define DEF_BATCH_WAIT_US 200
#define DEF_MAX_PLANS 8

db_ctx_t *g_db_ctx; /* the module's one SQLite connection, as in the reference (src/fp_handler.c:45) */

/* One tir_ctx per sample rate met (the mel filterbank depends on the file's rate, src/fp_handler.c:612-615).
 * plan[0] also holds the device mirror of audio_fingerprint and the batcher. */
static struct {
  pthread_mutex_t mu;
  int n;
  int rate[DEF_MAX_PLANS];
  tir_ctx *ctx[DEF_MAX_PLANS];
  int device;
} g_gpu = {PTHREAD_MUTEX_INITIALIZER, 0, {0}, {0}, 0};

/* ---------------------------------------------------------------------------------------------- helpers */

static void release(void *p) {
  if (p) ast_free(p);
}

static const char *backup_path(void) {
  const char *e = getenv("TIRESIAS_BACKUP_DATABASE"); /* deployment override; default = the reference's path */
  return (e && *e) ? e : DEF_BACKUP_DATABASE;
}

/* a private statement handle on the shared connection (what create_db_ctx does, src/fp_handler.c:1161-1169) */
static db_ctx_t *stmt_ctx(void) {
  db_ctx_t *c = ast_calloc(1, sizeof(*c));
  if (c) c->db = g_db_ctx->db;
  return c;
}
static void stmt_done(db_ctx_t *c) {
  if (!c) return;
  db_ctx_free(c);
  ast_free(c);
}

static bool run_sql(const char *fmt, ...) {
  char *sql = NULL;
  va_list ap;
  va_start(ap, fmt);
  const int n = vasprintf(&sql, fmt, ap);
  va_end(ap);
  if (n < 0 || !sql) return false;
  db_ctx_t *c = stmt_ctx();
  const bool ok = c && db_ctx_exec(c, sql);
  stmt_done(c);
  free(sql);
  return ok;
}

/* rows of a SELECT as JSON: the first row (single) or an array of all rows */
static struct ast_json *select_json(bool single, const char *fmt, ...) {
  char *sql = NULL;
  va_list ap;
  va_start(ap, fmt);
  const int n = vasprintf(&sql, fmt, ap);
  va_end(ap);
  if (n < 0 || !sql) return NULL;
  db_ctx_t *c = stmt_ctx();
  struct ast_json *out = NULL;
  if (c && db_ctx_query(c, sql)) {
    if (single) {
      out = db_ctx_get_record(c);
    } else {
      out = ast_json_array_create();
      for (struct ast_json *row; (row = db_ctx_get_record(c)) != NULL;) ast_json_array_append(out, row);
    }
  } else if (!single) {
    out = ast_json_array_create(); /* the reference returns an empty array when the query yields nothing */
  }
  stmt_done(c);
  free(sql);
  return out;
}

static bool uuid_text_to_bytes(const char *text, uint8_t out[16]) {
  uuid_t u;
  if (!text || strlen(text) != 36 || uuid_parse(text, u) != 0) return false;
  memcpy(out, u, 16);
  return true;
}

/* ---------------------------------------------------------------------------------------------- GPU plans */

static tir_ctx *plan_for_rate_locked(int rate) {
  for (int i = 0; i < g_gpu.n; i++)
    if (g_gpu.rate[i] == rate) return g_gpu.ctx[i];
  if (g_gpu.n >= DEF_MAX_PLANS) {
    ast_log(LOG_ERROR, "Too many distinct sample rates. rate[%d]\n", rate);
    return NULL;
  }
  tir_cfg cfg;
  tir_cfg_default(&cfg);
  cfg.device = g_gpu.device;
  cfg.win = DEF_AUBIO_BUFSIZE, cfg.hop = DEF_AUBIO_HOPSIZE, cfg.n_filters = DEF_AUBIO_FILTER, cfg.samplerate = rate;
  tir_ctx *ctx = NULL;
  const int rc = tir_open(&cfg, &ctx);
  if (rc != TIR_OK) {
    ast_log(LOG_ERROR, "Could not open the GPU fingerprint context. rate[%d], err[%d:%s]\n", rate, rc, ctx ? tir_last_error(ctx) : "");
    tir_close(ctx);
    return NULL;
  }
  g_gpu.rate[g_gpu.n] = rate, g_gpu.ctx[g_gpu.n] = ctx, g_gpu.n++;
  return ctx;
}
static tir_ctx *plan_for_rate(int rate) {
  pthread_mutex_lock(&g_gpu.mu);
  tir_ctx *c = plan_for_rate_locked(rate);
  pthread_mutex_unlock(&g_gpu.mu);
  return c;
}
static tir_ctx *main_plan(void) { return g_gpu.n ? g_gpu.ctx[0] : NULL; }

static void close_plans(void) {
  pthread_mutex_lock(&g_gpu.mu);
  for (int i = 0; i < g_gpu.n; i++) tir_close(g_gpu.ctx[i]), g_gpu.ctx[i] = NULL;
  g_gpu.n = 0;
  pthread_mutex_unlock(&g_gpu.mu);
}

/* ---------------------------------------------------------------------------------------------- audio files */

/* What aubio_source hands to the hop loop for the files this module meets -- the recordings the dialplan
 * application writes with ast_writefile(.., "wav") and the WAV files of the context directories: RIFF/WAVE,
 * PCM, 16 bit.  Returns interleaved samples; *frames = samples per channel. */
static int16_t *read_wav_pcm16(const char *filename, uint64_t *frames, int *channels, int *rate) {
  FILE *f = fopen(filename, "rb");
  if (!f) {
    ast_log(LOG_WARNING, "Could not open file. filename[%s]\n", filename);
    return NULL;
  }
  unsigned char hd[12], ck[8], fm[16];
  int16_t *pcm = NULL;
  int format = 0, bits = 0;
  bool have_fmt = false;
  *frames = 0, *channels = 0, *rate = 0;
  if (fread(hd, 1, 12, f) == 12 && memcmp(hd, "RIFF", 4) == 0 && memcmp(hd + 8, "WAVE", 4) == 0) {
    while (fread(ck, 1, 8, f) == 8) {
      const uint32_t len = (uint32_t)ck[4] | ((uint32_t)ck[5] << 8) | ((uint32_t)ck[6] << 16) | ((uint32_t)ck[7] << 24);
      if (memcmp(ck, "fmt ", 4) == 0 && len >= 16) {
        if (fread(fm, 1, 16, f) != 16) break;
        format = fm[0] | (fm[1] << 8), *channels = fm[2] | (fm[3] << 8);
        *rate = (int)((uint32_t)fm[4] | ((uint32_t)fm[5] << 8) | ((uint32_t)fm[6] << 16) | ((uint32_t)fm[7] << 24));
        bits = fm[14] | (fm[15] << 8);
        have_fmt = true;
        fseek(f, (long)(len - 16 + (len & 1)), SEEK_CUR);
      } else if (memcmp(ck, "data", 4) == 0) {
        if (!have_fmt || format != 1 || bits != 16 || *channels < 1 || *rate <= 0) break;
        const long at = ftell(f);
        fseek(f, 0, SEEK_END);
        long avail = ftell(f) - at;
        fseek(f, at, SEEK_SET);
        if ((long)len < avail) avail = (long)len;
        const uint64_t n = (uint64_t)avail / 2 / (uint64_t)*channels; /* whole frames only */
        pcm = ast_calloc(n * (uint64_t)*channels + 1, sizeof(int16_t));
        if (pcm && fread(pcm, sizeof(int16_t) * (size_t)*channels, n, f) == n) *frames = n;
        else release(pcm), pcm = NULL;
        break;
      } else {
        fseek(f, (long)(len + (len & 1)), SEEK_CUR);
      }
    }
  }
  fclose(f);
  if (!pcm) ast_log(LOG_ERROR, "Could not read the file as PCM16 WAV. filename[%s]\n", filename);
  return pcm;
}

/* The hop loop of create_audio_fingerprints() (src/fp_handler.c:577-671) for one file: mfcc coefficients and the
 * "%f" micro-unit values of every frame, computed on the GPU.  Caller frees *vq.  -1 on error. */
static int64_t fingerprint_file(const char *filename, int32_t **vq, int16_t **pcm_out, uint64_t *n_out, tir_ctx **plan_out,
                                int *channels_out) {
  uint64_t frames = 0;
  int channels = 0, rate = 0;
  int16_t *pcm = read_wav_pcm16(filename, &frames, &channels, &rate);
  if (!pcm) return -1;
  tir_ctx *plan = plan_for_rate(rate);
  if (!plan) {
    release(pcm);
    return -1;
  }
  if (plan_out) *plan_out = plan;
  /* more than one channel: aubio's source hands the hop loop the float mean of the channels; tir_extract_interleaved
   * computes exactly that on the device */
  if (channels_out) *channels_out = channels;
  if (!vq) { /* the caller only wants the samples (search: extraction happens inside the batched tir_search) */
    *pcm_out = pcm, *n_out = frames;
    return (int64_t)tir_n_frames(frames, DEF_AUBIO_HOPSIZE);
  }
  const uint64_t off[2] = {0, frames};
  const uint64_t nf = tir_n_frames(frames, DEF_AUBIO_HOPSIZE);
  *vq = ast_calloc(nf * DEF_AUBIO_COEFS + 1, sizeof(int32_t));
  uint64_t got = 0;
  const int rc = *vq ? tir_extract_interleaved(plan, pcm, channels, off, 1, NULL, *vq, &got) : TIR_ERR_NOMEM;
  release(pcm);
  if (rc != TIR_OK || got != nf) {
    ast_log(LOG_ERROR, "GPU extraction failed. filename[%s], err[%d:%s]\n", filename, rc, tir_last_error(plan));
    release(*vq), *vq = NULL;
    return -1;
  }
  return (int64_t)nf;
}

static char *md5_hex_of_file(const char *filename) {
  if (!filename) {
    ast_log(LOG_WARNING, "Wrong input parameter.\n");
    return NULL;
  }
  FILE *f = fopen(filename, "rb");
  if (!f) {
    ast_log(LOG_WARNING, "Could not open file. filename[%s]\n", filename);
    return NULL;
  }
  EVP_MD_CTX *md = EVP_MD_CTX_new();
  unsigned char buf[65536], dig[EVP_MAX_MD_SIZE];
  unsigned int dl = 0;
  bool ok = md && EVP_DigestInit_ex(md, EVP_md5(), NULL) == 1;
  for (size_t n; ok && (n = fread(buf, 1, sizeof buf, f)) > 0;) ok = EVP_DigestUpdate(md, buf, n) == 1;
  ok = ok && EVP_DigestFinal_ex(md, dig, &dl) == 1;
  EVP_MD_CTX_free(md);
  fclose(f);
  if (!ok || dl != 16) return NULL;
  char *hex = ast_calloc(33, 1);
  for (unsigned i = 0; hex && i < 16; i++) sprintf(hex + 2 * i, "%02x", dig[i]);
  return hex;
}

/* ---------------------------------------------------------------------------------------------- schema */

/* init_database(), src/fp_handler.c:673-756: the three tables and three indices of the on-disk contract */
static bool create_schema(void) {
  static const char *const ddl[] = {
      "create table context_list(   name        varchar(255),   directory   varchar(1023),   primary key(name));",
      "create table audio_list(   uuid           varchar(255),   name           varchar(255),   context        varchar(255),"
      "   hash           varchar(1023));",
      "create table audio_fingerprint(   context        varchar(255),   audio_uuid     varchar(255),   frame_idx      integer,"
      "   max1 real,   max2 real);",
      "create index idx_audio_fingerprint_context on audio_fingerprint(context);",
      "create index idx_audio_fingerprint_max1 on audio_fingerprint(max1);",
      "create index idx_audio_fingerprint_max2 on audio_fingerprint(max2);",
  };
  g_db_ctx = db_ctx_init(DEF_DATABASE_NAME);
  if (!g_db_ctx) return false;
  for (size_t i = 0; i < sizeof ddl / sizeof ddl[0]; i++)
    if (!db_ctx_exec(g_db_ctx, ddl[i])) {
      ast_log(LOG_ERROR, "Could not create the database schema. sql[%s]\n", ddl[i]);
      return false;
    }
  return true;
}

/* ---------------------------------------------------------------------------------------------- public API */

bool fp_init(void) {
  if (!create_schema()) {
    ast_log(LOG_ERROR, "Could not initiate database.\n");
    return false;
  }
  db_ctx_t *c = stmt_ctx();
  const bool loaded = c && db_ctx_load_db_data(c, backup_path());
  stmt_done(c);
  if (!loaded) {
    ast_log(LOG_ERROR, "Could not load the database data.\n");
    return false;
  }
  /* the device: no CPU fallback -- without a B200 the module declines to load */
  const char *dev = getenv("TIRESIAS_GPU_DEVICE");
  g_gpu.device = dev ? atoi(dev) : 0;
  tir_ctx *plan = plan_for_rate(8000); /* telephony: the plan the dialplan recordings use holds the table */
  if (!plan) {
    ast_log(LOG_ERROR, "Could not initiate the GPU fingerprint engine.\n");
    return false;
  }
  uint64_t n_audio = 0, n_rows = 0, n_skipped = 0;
  int rc = tir_db_load_sqlite(plan, g_db_ctx->db, &n_audio, &n_rows, &n_skipped);
  if (rc != TIR_OK) {
    ast_log(LOG_ERROR, "Could not mirror audio_fingerprint to the device. err[%d:%s]\n", rc, tir_last_error(plan));
    return false;
  }
  ast_log(LOG_NOTICE, "Device table loaded. audios[%llu], rows[%llu], skipped[%llu]\n", (unsigned long long)n_audio,
          (unsigned long long)n_rows, (unsigned long long)n_skipped);
  rc = tir_batcher_start(plan, DEF_BATCH_MAX, DEF_BATCH_WAIT_US);
  if (rc != TIR_OK) {
    ast_log(LOG_ERROR, "Could not start the search batcher. err[%d:%s]\n", rc, tir_last_error(plan));
    return false;
  }
  return true;
}

bool fp_term(void) {
  db_ctx_t *c = stmt_ctx();
  const bool saved = c && db_ctx_backup(c, backup_path());
  stmt_done(c);
  close_plans(); /* stops the batcher (queued searches are served first), frees the device table */
  if (!saved) {
    ast_log(LOG_ERROR, "Could not write database.\n");
    return false;
  }
  db_ctx_term(g_db_ctx);
  g_db_ctx = NULL;
  return true;
}

static struct ast_json *audio_row(const char *uuid) {
  if (!uuid) {
    ast_log(LOG_WARNING, "Wrong input parameter.\n");
    return NULL;
  }
  return select_json(true, "select * from audio_list where uuid = '%s';", uuid);
}

bool fp_delete_audio_list_info(const char *uuid) {
  if (!uuid) {
    ast_log(LOG_WARNING, "Wrong input parameter.\n");
    return false;
  }
  struct ast_json *row = audio_row(uuid);
  if (!row) {
    ast_log(LOG_NOTICE, "Could not find audio list info.\n");
    return false;
  }
  ast_json_unref(row);
  if (!run_sql("delete from audio_list where uuid='%s';", uuid)) {
    ast_log(LOG_WARNING, "Could not delete audio list info. uuid[%s]\n", uuid);
    return false;
  }
  if (!run_sql("delete from audio_fingerprint where audio_uuid='%s';", uuid)) {
    ast_log(LOG_WARNING, "Could not delete audio fingerprint info. audio_uuid[%s]\n", uuid);
    return false;
  }
  /* the device mirror follows the table (an audio that never got rows is simply unknown there) */
  uint8_t raw[16];
  tir_ctx *plan = main_plan();
  if (plan && uuid_text_to_bytes(uuid, raw)) {
    const int rc = tir_db_remove(plan, raw);
    if (rc != TIR_OK && rc != TIR_ERR_NOTFOUND) ast_log(LOG_WARNING, "Could not drop the audio from the device table. uuid[%s], err[%d]\n", uuid, rc);
  }
  return true;
}

/* 1: listed now, 0: the (context, md5) pair is already there, -1: error   (src/fp_handler.c:479-530) */
static int list_audio(const char *context, const char *filename, const char *uuid) {
  char *hash = md5_hex_of_file(filename);
  if (!hash) {
    ast_log(LOG_WARNING, "Could not create hash info.\n");
    return -1;
  }
  struct ast_json *dup = select_json(true, "select * from audio_list where context = '%s' and hash = '%s';", context, hash);
  if (dup) {
    ast_log(LOG_VERBOSE, "The given file is already fingerprinted. context[%s], filename[%s]\n", context, filename);
    ast_json_unref(dup);
    release(hash);
    return 0;
  }
  char *path = ast_strdup(filename);
  struct ast_json *row = ast_json_pack("{s:s, s:s, s:s, s:s}", "uuid", uuid, "name", basename(path), "context", context, "hash", hash);
  release(path);
  release(hash);
  const bool ok = row && db_ctx_insert(g_db_ctx, "audio_list", row);
  ast_json_unref(row);
  if (!ok) {
    ast_log(LOG_ERROR, "Could not create fingerprint info.\n");
    return -1;
  }
  return 1;
}

bool fp_craete_audio_list_info(const char *context, const char *filename) {
  if (!context || !filename) {
    ast_log(LOG_WARNING, "Wrong input parameter.\n");
    return false;
  }
  char *uuid = fp_generate_uuid();
  const int listed = uuid ? list_audio(context, filename, uuid) : -1;
  if (listed <= 0) {
    if (listed < 0) ast_log(LOG_WARNING, "Could not create audio_list info. context[%s], filename[%s]\n", context, filename);
    else ast_log(LOG_VERBOSE, "The given audio file is already exist in the list. context[%s], filename[%s]\n", context, filename);
    release(uuid);
    return listed == 0;
  }
  /* frames on the GPU; rows into SQLite in one transaction (the very values the reference's per-frame textual
   * INSERTs store, src/fp_handler.c:538-575 + src/db_ctx_handler.c:480) and into the device mirror */
  int32_t *vq = NULL;
  tir_ctx *plan = NULL;
  const int64_t nf = fingerprint_file(filename, &vq, NULL, NULL, &plan, NULL);
  bool ok = nf >= 0;
  uint8_t raw[16];
  if (ok) {
    const int rc = tir_sqlite_insert_fingerprints(plan, g_db_ctx->db, context, uuid, vq, (uint32_t)nf);
    ok = rc == TIR_OK;
    if (!ok) ast_log(LOG_ERROR, "Could not insert the fingerprint rows. err[%d:%s]\n", rc, tir_last_error(plan));
  }
  if (ok && uuid_text_to_bytes(uuid, raw)) {
    int32_t *v1 = ast_calloc((size_t)nf + 1, sizeof(int32_t)), *v2 = ast_calloc((size_t)nf + 1, sizeof(int32_t));
    for (int64_t i = 0; v1 && v2 && i < nf; i++) v1[i] = vq[2 * i], v2[i] = vq[2 * i + 1];
    const int rc = (v1 && v2) ? tir_db_add(main_plan(), raw, v1, v2, (uint32_t)nf) : TIR_ERR_NOMEM;
    release(v1), release(v2);
    ok = rc == TIR_OK;
    if (!ok) ast_log(LOG_ERROR, "Could not add the audio to the device table. err[%d:%s]\n", rc, tir_last_error(main_plan()));
  }
  release(vq);
  if (!ok) {
    ast_log(LOG_NOTICE, "Could not create audio fingerprint info.\n");
    fp_delete_audio_list_info(uuid); /* (the reference passes the file name here, src/fp_handler.c:192, and leaves the row behind) */
    release(uuid);
    return false;
  }
  release(uuid);
  return true;
}

struct ast_json *fp_search_fingerprint_info(const char *context, const char *filename, const int coefs, const double tolerance,
                                            const int freq_ignore_low, const int freq_ignore_high) {
  if (!context || !filename) {
    ast_log(LOG_WARNING, "Wrong input parameter.\n");
    return NULL;
  }
  ast_log(LOG_DEBUG, "Fired fp_search_fingerprint_info. context[%s], filename[%s], coefs[%d], tolerance[%f], freq_ignore_low[%d], freq_ignore_high[%d]\n",
          context, filename, coefs, tolerance, freq_ignore_low, freq_ignore_high);
  if (coefs < 1 || coefs > DEF_AUBIO_COEFS) {
    ast_log(LOG_WARNING, "Wrong coefs count. max[%d], coefs[%d]\n", DEF_AUBIO_COEFS, coefs);
    return NULL;
  }
  double tole = tolerance;
  if (tole < 0) {
    ast_log(LOG_NOTICE, "Wrong tolerance setting. Set to default. tolerance[%f], default[%f]\n", tolerance, DEF_SEARCH_TOLERANCE);
    tole = DEF_SEARCH_TOLERANCE;
  }
  int16_t *pcm = NULL;
  uint64_t n = 0;
  tir_ctx *plan = NULL;
  int channels = 1;
  if (fingerprint_file(filename, NULL, &pcm, &n, &plan, &channels) < 0) {
    ast_log(LOG_ERROR, "Could not create fingerprint info.\n");
    return NULL;
  }
  /* The match runs where the table lives.  The context argument is not part of the reference's match SQL
   * (src/fp_handler.c:308-314): every context's audios compete, here too. */
  tir_hit hit;
  int rc;
  if (plan == main_plan() && channels == 1) {
    rc = tir_search_one(plan, pcm, n, coefs, tole, freq_ignore_low, freq_ignore_high, &hit); /* batched with concurrent callers */
  } else { /* a recording at another sample rate (its own filterbank, the one table), or one with several channels */
    const uint64_t off[2] = {0, n};
    const uint64_t nf = tir_n_frames(n, DEF_AUBIO_HOPSIZE);
    float *coef = ast_calloc(nf * DEF_AUBIO_COEFS + 1, sizeof(float));
    double *y = ast_calloc(nf * DEF_AUBIO_COEFS + 1, sizeof(double));
    uint64_t got = 0;
    rc = (coef && y) ? tir_extract_interleaved(plan, pcm, channels, off, 1, coef, NULL, &got) : TIR_ERR_NOMEM;
    for (uint64_t i = 0; rc == TIR_OK && i < nf * DEF_AUBIO_COEFS; i++) y[i] = 10 * log10(fabs((double)coef[i])); /* :651 */
    const uint64_t foff[2] = {0, nf};
    if (rc == TIR_OK) rc = tir_match(main_plan(), y, foff, 1, coefs, tole, freq_ignore_low, freq_ignore_high, &hit), plan = main_plan();
    release(coef), release(y);
  }
  release(pcm);
  if (rc != TIR_OK) {
    ast_log(LOG_ERROR, "GPU search failed. err[%d:%s]\n", rc, tir_last_error(plan));
    return NULL;
  }
  if (hit.match_count <= 0) {
    ast_log(LOG_NOTICE, "Could not find data.\n");
    return NULL;
  }
  char text[DEF_UUID_STR_LEN];
  uuid_unparse_lower(hit.uuid, text);
  struct ast_json *res = audio_row(text);
  if (!res) {
    ast_log(LOG_WARNING, "Could not find audio list info.\n");
    return NULL;
  }
  ast_json_object_set(res, "frame_count", ast_json_integer_create(hit.frame_count));
  ast_json_object_set(res, "match_count", ast_json_integer_create(hit.match_count));
  return res;
}

struct ast_json *fp_get_audio_lists_all(void) { return select_json(false, "select * from audio_list;"); }

struct ast_json *fp_get_audio_lists_by_contextname(const char *name) {
  if (!name) {
    ast_log(LOG_WARNING, "Wrong input parameter.\n");
    return NULL;
  }
  return select_json(false, "select * from audio_list where context = '%s';", name);
}

struct ast_json *fp_get_context_lists_all(void) { return select_json(false, "select * from context_list;"); }

struct ast_json *fp_get_context_list_info(const char *name) {
  if (!name) {
    ast_log(LOG_WARNING, "Wrong input parameter.\n");
    return NULL;
  }
  return select_json(true, "select * from context_list where name == '%s';", name);
}

bool fp_create_context_list_info(const char *name, const char *directory, bool replace) {
  if (!name) {
    ast_log(LOG_WARNING, "Wrong input parameter.\n");
    return false;
  }
  struct ast_json *row = ast_json_pack("{s:s, s:s}", "name", name, "directory", directory);
  const bool ok = row && (replace ? db_ctx_insert_or_replace(g_db_ctx, "context_list", row) : db_ctx_insert(g_db_ctx, "context_list", row));
  ast_json_unref(row);
  if (!ok) ast_log(LOG_WARNING, "Could not create context list info. name[%s]\n", name);
  return ok;
}

bool fp_delete_context_list_info(const char *name) {
  if (!name) {
    ast_log(LOG_WARNING, "Wrong input parameter.\n");
    return false;
  }
  struct ast_json *ctx = fp_get_context_list_info(name);
  if (!ctx) {
    ast_log(LOG_NOTICE, "Could not find context info. context[%s]\n", name);
    return false;
  }
  ast_json_unref(ctx);
  struct ast_json *audios = fp_get_audio_lists_by_contextname(name); /* the audios of the context go first */
  for (size_t i = 0; i < ast_json_array_size(audios); i++) {
    const char *uuid = ast_json_string_get(ast_json_object_get(ast_json_array_get(audios, i), "uuid"));
    if (uuid && !fp_delete_audio_list_info(uuid)) ast_log(LOG_WARNING, "Could not delete audio_list info. uuid[%s]\n", uuid);
  }
  ast_json_unref(audios);
  if (!run_sql("delete from context_list where name == '%s';", name)) {
    ast_log(LOG_WARNING, "Could not delete context list info.\n");
    return false;
  }
  return true;
}

char *fp_generate_uuid(void) {
  uuid_t u;
  char text[DEF_UUID_STR_LEN];
  uuid_generate(u);
  uuid_unparse_lower(u, text);
  return ast_strdup(text);
}

char *fp_create_hash(const char *filename) {
  if (!filename) {
    ast_log(LOG_WARNING, "Wrong input parameter.\n");
    return NULL;
  }
  return md5_hex_of_file(filename);
}
