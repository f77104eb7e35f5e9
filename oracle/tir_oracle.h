/*
 * tir_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the fingerprint hot path of pchero/asterisk-tiresias:
 *   extraction : src/fp_handler.c:577-671 (create_audio_fingerprints) + the libaubio 0.4.x
 *                stages it calls (source -> pvoc -> mfcc); aubio is a third-party dependency
 *                that is NOT vendored in /root/reference and NOT installed in this image, so
 *                its stages are restated from its published sources.
 *   match      : src/fp_handler.c:207-408 (fp_search_fingerprint_info) -- not restated but RUN:
 *                the literal SQL text of the reference is executed on the real libsqlite3
 *                (tir_oracle_sqlite.c).
 *
 * PARITY UNPINNED at the libaubio boundary: the reference ships no tests, fixtures or golden
 * vectors, and libaubio cannot be run here.  The restatement is pinned instead against
 * independent second opinions (numpy float64 FFT, scipy DCT, closed forms) in tests/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libtiresias_gpu.so) never links or calls it.
 */
#ifndef TIR_ORACLE_H_
#define TIR_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIRO_NULL_V INT32_MIN /* "column is NULL" marker in micro-unit arrays */

typedef struct tiro_plan tiro_plan;

/* new_aubio_pvoc(win,hop) + new_aubio_mfcc(win,n_filters,n_coefs,samplerate): fp_handler.c:613-617 */
tiro_plan *tiro_plan_create(int win, int hop, int n_filters, int n_coefs, int samplerate);
void tiro_plan_destroy(tiro_plan *p);

/* table access for the stage tests */
int tiro_plan_spec_len(const tiro_plan *p);              /* win/2+1                       */
const float *tiro_plan_window(const tiro_plan *p);       /* [win]   "hanningz"            */
const float *tiro_plan_filters(const tiro_plan *p);      /* [n_filters][win/2+1]          */
const float *tiro_plan_dct(const tiro_plan *p);          /* [n_coefs][n_filters]          */
const float *tiro_plan_band_edges(const tiro_plan *p);   /* [n_filters+2] Slaney edges Hz */

/* number of hops aubio_source_do delivers before reads==0 : fp_handler.c:632-636 */
size_t tiro_n_frames(size_t n_samples, int hop);

/* one aubio_pvoc_do magnitude spectrum from an (already slid) un-windowed buffer of win floats */
void tiro_pvoc_norm(const tiro_plan *p, const float *data, float *norm /*[win/2+1]*/);
/* the float32 FFT used by the oracle ("TIR-FFT"): complex spectrum bins 0..win/2 of a real frame */
void tiro_rfft(const tiro_plan *p, const float *frame, float *re, float *im);
/* which float32 FFT runs under the pipeline: 0 = TIR-FFT (default), 1 = Ooura-style radix-2 + rftfsub
 * untangling, 2 = float64 rounded.  Process-wide; for the FFT-order sensitivity study only. */
void tiro_set_fft_kind(int kind);
int tiro_get_fft_kind(void);
/* aubio_mfcc_do on one magnitude spectrum */
void tiro_mfcc(const tiro_plan *p, const float *norm, float *mel /*[n_filters] or NULL*/,
               float *coef /*[n_coefs]*/);

/* "%f" marshalling of db_ctx_handler.c:480 : y -> micro-units (TIRO_NULL_V when not finite) */
int32_t tiro_quantize(double y);

/*
 * create_audio_fingerprints() for one mono PCM16 clip.
 *   coef [F][n_coefs]  float  mfcc_out->data[i]
 *   y    [F][n_coefs]  double 10*log10(fabs(c))            fp_handler.c:651
 *   vq   [F][n_coefs]  int32  value as stored through "%f" (micro-units), TIRO_NULL_V = NULL
 * any of the three may be NULL.  returns F.
 */
size_t tiro_extract(const tiro_plan *p, const int16_t *pcm, size_t n_samples, float *coef,
                    double *y, int32_t *vq);

/* the same for `channels` interleaved channels (pcm[sample frame][channel], n_samples sample frames): the hop loop
 * runs on the float mean of the channels, as aubio's source computes it (aubio source_wavread.c) */
size_t tiro_extract_interleaved(const tiro_plan *p, const int16_t *pcm, size_t n_samples, int channels, float *coef,
                                double *y, int32_t *vq);

/* batch over concatenated clips, clip c = pcm[clip_off[c] .. clip_off[c+1]); frames are
 * written back to back in clip order.  n_threads>1 uses pthreads (one clip per task). */
size_t tiro_extract_batch(const tiro_plan *p, const int16_t *pcm, const uint64_t *clip_off,
                          uint32_t n_clips, float *coef, double *y, int32_t *vq, int n_threads);

/* ------------------------------------------------------------------ match (real SQLite) */
typedef struct tiro_db tiro_db;

tiro_db *tiro_db_open(void); /* schema of fp_handler.c:686-753, ":memory:" */
void tiro_db_close(tiro_db *db);
const char *tiro_db_sqlite_version(void);
/* audio_list row (fp_handler.c:512-522) */
int tiro_db_add_audio(tiro_db *db, const char *uuid, const char *name, const char *context,
                      const char *hash);
/* one textual INSERT per frame (fp_handler.c:559-571, db_ctx_handler.c:413-556);
 * y is [n_frames][2]; a non-finite value leaves the column out (=> NULL). */
int tiro_db_add_fingerprints(tiro_db *db, const char *context, const char *uuid, const double *y,
                             size_t n_frames, int literal_autocommit);
int tiro_db_delete_audio(tiro_db *db, const char *uuid); /* fp_handler.c:135,147 */
long tiro_db_count_rows(tiro_db *db);
void *tiro_db_handle(tiro_db *db); /* the sqlite3* itself (what the module holds in g_db_ctx->db) */
long tiro_db_dump_audio(tiro_db *db, const char *uuid, long cap, long *frame_idx, double *max1, double *max2,
                        int *type1, int *type2, char *context, size_t context_cap);

typedef struct {
  int found;        /* 0 => reference returns NULL (TIRSTATUS=NOTFOUND) */
  char uuid[64];    /* winning audio_uuid                                */
  int match_count;  /* count(*)                                          */
  int frame_count;  /* all query frames (fp_handler.c:286,403)            */
  long rows_in_windows; /* sum over query frames of rows inserted into the temp table (R_k bookkeeping) */
} tiro_hit;

/* fp_search_fingerprint_info() from the already extracted query values y[n_frames][2];
 * has_y[n_frames][2] == 0 marks a missing JSON key (read back as 0.0, like ast_json_real_get(NULL)). */
int tiro_db_search(tiro_db *db, const double *y, const uint8_t *has_y, size_t n_frames, int coefs,
                   double tolerance, int freq_ignore_low, int freq_ignore_high, tiro_hit *out);

#ifdef __cplusplus
}
#endif
#endif
