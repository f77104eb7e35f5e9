/*
 * tir_oracle_sqlite.c -- CPU ORACLE for the match stage (test infrastructure, NOT product code).
 *
 * The match loop of the reference is SQL text executed by SQLite, so the oracle does not restate
 * it: it builds the SAME SQL strings with the SAME printf formats and runs them on the real
 * libsqlite3 of this image (3.45.1; the header is not installed, so the few prototypes used are
 * declared by hand and the library is linked as -l:libsqlite3.so.0).
 *
 *   schema            src/fp_handler.c:686-692, 700-706, 714-753
 *   ingest            src/fp_handler.c:512-522, 559-571 via src/db_ctx_handler.c:413-556
 *                     (reals through "%f" :480, strings quoted and unescaped :469, ints "%ld" :475)
 *   temp table        src/fp_handler.c:857-910
 *   per-frame probe   src/fp_handler.c:287-362
 *   tally / winner    src/fp_handler.c:367-373
 *   delete            src/fp_handler.c:135,147
 */
#define _GNU_SOURCE
#include "tir_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* hand-declared subset of sqlite3.h */
typedef struct sqlite3 sqlite3;
typedef struct sqlite3_stmt sqlite3_stmt;
int sqlite3_open(const char *, sqlite3 **);
int sqlite3_close(sqlite3 *);
int sqlite3_exec(sqlite3 *, const char *, int (*)(void *, int, char **, char **), void *, char **);
int sqlite3_prepare_v2(sqlite3 *, const char *, int, sqlite3_stmt **, const char **);
int sqlite3_step(sqlite3_stmt *);
int sqlite3_finalize(sqlite3_stmt *);
const unsigned char *sqlite3_column_text(sqlite3_stmt *, int);
int sqlite3_column_int(sqlite3_stmt *, int);
double sqlite3_column_double(sqlite3_stmt *, int);
int sqlite3_column_type(sqlite3_stmt *, int);
int sqlite3_column_count(sqlite3_stmt *);
const char *sqlite3_column_name(sqlite3_stmt *, int);
int sqlite3_changes(sqlite3 *);
void sqlite3_free(void *);
const char *sqlite3_libversion(void);
#define SQLITE_OK 0
#define SQLITE_ROW 100

#define DEF_AUBIO_COEFS 2 /* fp_handler.c:39 */

struct tiro_db {
  sqlite3 *db;
  unsigned long search_seq;
};

static int exec_sql(tiro_db *d, const char *sql) {
  char *err = NULL;
  int rc = sqlite3_exec(d->db, sql, NULL, NULL, &err);
  if (rc != SQLITE_OK) {
    fprintf(stderr, "[tir-oracle] sqlite error %d: %s\n  sql: %.200s\n", rc, err ? err : "?", sql);
    if (err) sqlite3_free(err);
    return -1;
  }
  return 0;
}

const char *tiro_db_sqlite_version(void) { return sqlite3_libversion(); }

tiro_db *tiro_db_open(void) {
  tiro_db *d = (tiro_db *)calloc(1, sizeof(*d));
  if (sqlite3_open(":memory:", &d->db) != SQLITE_OK) {
    free(d);
    return NULL;
  }
  char *sql = NULL, *tmp = NULL;
  int rc = 0;
  /* fp_handler.c:686-692 */
  rc |= exec_sql(d, "create table context_list("
                    "   name        varchar(255),"
                    "   directory   varchar(1023),"
                    "   primary key(name)"
                    ");");
  /* fp_handler.c:700-706 */
  rc |= exec_sql(d, "create table audio_list("
                    "   uuid           varchar(255),"
                    "   name           varchar(255),"
                    "   context        varchar(255),"
                    "	hash           varchar(1023)"
                    ");");
  /* fp_handler.c:714-727 */
  if (asprintf(&sql, "%s",
               "create table audio_fingerprint("
               " context        varchar(255),"
               " audio_uuid     varchar(255),"
               " frame_idx      integer") < 0)
    return NULL;
  for (int i = 0; i < DEF_AUBIO_COEFS; i++) {
    if (asprintf(&tmp, "%s, max%d real", sql, i + 1) < 0) return NULL;
    free(sql);
    sql = tmp;
  }
  if (asprintf(&tmp, "%s);", sql) < 0) return NULL;
  free(sql);
  sql = tmp;
  rc |= exec_sql(d, sql);
  free(sql);
  /* fp_handler.c:736, 745-753 */
  rc |= exec_sql(d, "create index idx_audio_fingerprint_context on audio_fingerprint(context);");
  for (int i = 1; i <= DEF_AUBIO_COEFS; i++) {
    if (asprintf(&sql, "create index idx_audio_fingerprint_max%d on audio_fingerprint(max%d);", i, i) < 0)
      return NULL;
    rc |= exec_sql(d, sql);
    free(sql);
  }
  if (rc) {
    tiro_db_close(d);
    return NULL;
  }
  return d;
}

void tiro_db_close(tiro_db *d) {
  if (!d) return;
  if (d->db) sqlite3_close(d->db);
  free(d);
}

int tiro_db_add_audio(tiro_db *d, const char *uuid, const char *name, const char *context,
                      const char *hash) {
  char *sql = NULL;
  /* db_ctx_insert_basic: keys in jansson insertion order of the pack at fp_handler.c:512-517 */
  if (asprintf(&sql, "insert into audio_list(uuid, name, context, hash) values ('%s', '%s', '%s', '%s');",
               uuid, name, context, hash) < 0)
    return -1;
  int rc = exec_sql(d, sql);
  free(sql);
  return rc;
}

int tiro_db_add_fingerprints(tiro_db *d, const char *context, const char *uuid, const double *y,
                             size_t n_frames, int literal_autocommit) {
  int rc = 0;
  if (!literal_autocommit) rc |= exec_sql(d, "begin;");
  for (size_t t = 0; t < n_frames; t++) {
    /* keys: frame_idx, audio_uuid (pack :645-648), max1, max2 (:649-652, only when
     * ast_json_real_create accepted the value, i.e. finite), context (:565) */
    char keys[128], vals[512];
    int kn = snprintf(keys, sizeof(keys), "frame_idx, audio_uuid");
    int vn = snprintf(vals, sizeof(vals), "%ld, '%s'", (long)t, uuid);
    for (int j = 0; j < DEF_AUBIO_COEFS; j++) {
      double v = y[t * DEF_AUBIO_COEFS + j];
      if (!isfinite(v)) continue;
      kn += snprintf(keys + kn, sizeof(keys) - kn, ", max%d", j + 1);
      vn += snprintf(vals + vn, sizeof(vals) - vn, ", %f", v);
    }
    kn += snprintf(keys + kn, sizeof(keys) - kn, ", context");
    vn += snprintf(vals + vn, sizeof(vals) - vn, ", '%s'", context);
    char *sql = NULL;
    if (asprintf(&sql, "insert into %s(%s) values (%s);", "audio_fingerprint", keys, vals) < 0) return -1;
    rc |= exec_sql(d, sql);
    free(sql);
  }
  if (!literal_autocommit) rc |= exec_sql(d, "commit;");
  return rc;
}

int tiro_db_delete_audio(tiro_db *d, const char *uuid) {
  char *sql = NULL;
  int rc = 0;
  if (asprintf(&sql, "delete from audio_list where uuid='%s';", uuid) < 0) return -1;
  rc |= exec_sql(d, sql);
  free(sql);
  if (asprintf(&sql, "delete from audio_fingerprint where audio_uuid='%s';", uuid) < 0) return -1;
  rc |= exec_sql(d, sql);
  free(sql);
  return rc;
}

/* the raw connection, so that tests can hand it to code under test the way the module would hand
 * over g_db_ctx->db (src/fp_handler.c:45) */
void *tiro_db_handle(tiro_db *d) { return d ? (void *)d->db : NULL; }

/* rows of one audio in rowid order: frame_idx, max1, max2 and the storage class of each value
 * (sqlite3_column_type: 1 integer, 2 float, 3 text, 4 blob, 5 null) */
long tiro_db_dump_audio(tiro_db *d, const char *uuid, long cap, long *frame_idx, double *max1, double *max2,
                        int *type1, int *type2, char *context, size_t context_cap) {
  char *sql = NULL;
  if (asprintf(&sql, "select frame_idx, max1, max2, context from audio_fingerprint where audio_uuid='%s' order by rowid", uuid) < 0)
    return -1;
  sqlite3_stmt *st = NULL;
  long n = 0;
  if (sqlite3_prepare_v2(d->db, sql, -1, &st, NULL) != SQLITE_OK) {
    free(sql);
    return -1;
  }
  free(sql);
  while (sqlite3_step(st) == SQLITE_ROW) {
    if (n < cap) {
      frame_idx[n] = sqlite3_column_int(st, 0);
      type1[n] = sqlite3_column_type(st, 1), type2[n] = sqlite3_column_type(st, 2);
      max1[n] = sqlite3_column_double(st, 1), max2[n] = sqlite3_column_double(st, 2);
      if (n == 0 && context) {
        const unsigned char *t = sqlite3_column_text(st, 3);
        snprintf(context, context_cap, "%s", t ? (const char *)t : "");
      }
    }
    n++;
  }
  sqlite3_finalize(st);
  return n;
}

long tiro_db_count_rows(tiro_db *d) {
  sqlite3_stmt *st = NULL;
  long n = -1;
  if (sqlite3_prepare_v2(d->db, "select count(*) from audio_fingerprint", -1, &st, NULL) != SQLITE_OK)
    return -1;
  if (sqlite3_step(st) == SQLITE_ROW) n = sqlite3_column_int(st, 0);
  sqlite3_finalize(st);
  return n;
}

int tiro_db_search(tiro_db *d, const double *y, const uint8_t *has_y, size_t n_frames, int coefs,
                   double tolerance, int freq_ignore_low, int freq_ignore_high, tiro_hit *out) {
  memset(out, 0, sizeof(*out));
  /* fp_handler.c:247-256 */
  if (coefs < 1 || coefs > DEF_AUBIO_COEFS) return -1;
  double tole = tolerance;
  if (tole < 0) tole = 0.001; /* DEF_SEARCH_TOLERANCE */

  /* fp_handler.c:258-272: temp_<uuid with '-' -> '_'>; any unique name behaves the same */
  char tablename[96];
  snprintf(tablename, sizeof(tablename), "temp_%08lx_0000_4000_8000_%012lx", 0x7e57ab1eUL, ++d->search_seq);
  char *sql = NULL, *tmp = NULL;
  if (asprintf(&sql,
               "create table %s("
               " context        varchar(255),"
               " audio_uuid     varchar(255),"
               " frame_idx      integer",
               tablename) < 0)
    return -1;
  for (int i = 0; i < DEF_AUBIO_COEFS; i++) {
    if (asprintf(&tmp, "%s, max%d real", sql, i + 1) < 0) return -1;
    free(sql);
    sql = tmp;
  }
  if (asprintf(&tmp, "%s);", sql) < 0) return -1;
  free(sql);
  sql = tmp;
  int rc = exec_sql(d, sql);
  free(sql);
  if (rc) return -1;

  /* fp_handler.c:286-362 */
  int frame_count = (int)n_frames;
  for (int i = 0; i < frame_count; i++) {
    double v1 = (!has_y || has_y[i * DEF_AUBIO_COEFS + 0]) ? y[i * DEF_AUBIO_COEFS + 0] : 0.0;
    double freq = (int)v1; /* :290 */
    double freq_tmp;
    if (freq_ignore_low > 0) {
      freq_tmp = 10 * log10(freq_ignore_low);
      if (freq < freq_tmp) continue;
    }
    if (freq_ignore_high > 0) {
      freq_tmp = 10 * log10(freq_ignore_high);
      if (freq > freq_tmp) continue;
    }
    if (asprintf(&sql,
                 "insert into %s select * from audio_fingerprint where "
                 " max1 >= %f "
                 " and max1 <= %f ",
                 tablename, freq - tole, freq + tole) < 0)
      return -1;
    for (int j = 1; j < coefs; j++) {
      char tmp_max[16];
      snprintf(tmp_max, sizeof(tmp_max), "max%d", j + 1);
      freq = (!has_y || has_y[i * DEF_AUBIO_COEFS + j]) ? y[i * DEF_AUBIO_COEFS + j] : 0.0; /* :321 */
      if (freq_ignore_low > 0) {
        freq_tmp = 10 * log10(freq_ignore_low);
        if (freq < freq_tmp) continue;
      }
      if (freq_ignore_high > 0) {
        freq_tmp = 10 * log10(freq_ignore_high);
        if (freq > freq_tmp) continue;
      }
      if (asprintf(&tmp, "%s and %s >= %f and %s <= %f", sql, tmp_max, freq - tole, tmp_max, freq + tole) < 0)
        return -1;
      free(sql);
      sql = tmp;
    }
    if (asprintf(&tmp, "%s group by audio_uuid", sql) < 0) return -1;
    free(sql);
    sql = tmp;
    exec_sql(d, sql); /* the reference ignores the result (:357-359) */
    out->rows_in_windows += sqlite3_changes(d->db);
    free(sql);
  }

  /* fp_handler.c:367-373: first record only */
  if (asprintf(&sql, "select *, count(*) from %s group by audio_uuid order by count(*) DESC", tablename) < 0)
    return -1;
  sqlite3_stmt *st = NULL;
  rc = sqlite3_prepare_v2(d->db, sql, -1, &st, NULL);
  free(sql);
  if (rc == SQLITE_OK && sqlite3_step(st) == SQLITE_ROW) {
    int nc = sqlite3_column_count(st);
    for (int c = 0; c < nc; c++) {
      const char *name = sqlite3_column_name(st, c);
      if (strcmp(name, "audio_uuid") == 0) {
        const unsigned char *t = sqlite3_column_text(st, c);
        snprintf(out->uuid, sizeof(out->uuid), "%s", t ? (const char *)t : "");
      } else if (strcmp(name, "count(*)") == 0) {
        out->match_count = sqlite3_column_int(st, c); /* db_ctx_handler.c:297 */
      }
    }
    out->found = 1;
  }
  if (st) sqlite3_finalize(st);

  /* fp_handler.c:377, 904 */
  if (asprintf(&sql, "drop table %s;", tablename) < 0) return -1;
  exec_sql(d, sql);
  free(sql);

  /* fp_handler.c:386-404: the reference then reads audio_list for the winner (NULL if the row is
   * gone) and attaches frame_count (all frames) and match_count */
  out->frame_count = frame_count;
  return 0;
}
