// tir_sqlite.cpp -- keeping the device table coherent with the reference's SQLite database.
//
// SQLite stays the system of record (schema src/fp_handler.c:686-753, in-memory connection
// g_db_ctx restored from / backed up to the file, src/fp_handler.c:68-108).  Two seams:
//   * tir_db_load_sqlite: after fp_init restored the DB, mirror table audio_fingerprint into the
//     device table;
//   * tir_sqlite_insert_fingerprints: the rows of one new audio in ONE transaction through ONE
//     prepared statement instead of create_audio_fingerprint_info's textual INSERT per frame
//     (src/fp_handler.c:538-575 via src/db_ctx_handler.c:413-556) -- the dominant cost of
//     fingerprinting once extraction runs on the GPU.  The stored values are the
//     ones the text path stores: a real goes through "%f" (src/db_ctx_handler.c:480), so SQLite
//     parses a decimal with six digits, v / 10^6 correctly rounded -- which is what is bound here;
//     a non-finite value never became a JSON real, its key is missing from the INSERT: NULL.
// libsqlite3 is resolved at run time (the Asterisk process already has it loaded; its header is
// not needed): dlopen with RTLD_NOLOAD first, then libsqlite3.so.0.
#include <cstring>

#include "tir_internal.h"
#include "tir_sqlite_dl.h"

namespace {

using Sqlite = TirSqlite;
Sqlite &sq() { return tir_sqlite(); }

int hexval(char c) {
  if (c >= '0' && c <= '9') return c - '0';
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  if (c >= 'A' && c <= 'F') return c - 'A' + 10;
  return -1;
}

// canonical 8-4-4-4-12 text -> 16 bytes
bool parse_uuid(const char *t, uint8_t out[16]) {
  if (!t || std::strlen(t) != 36) return false;
  int k = 0;
  for (int i = 0; i < 36;) {
    if (i == 8 || i == 13 || i == 18 || i == 23) {
      if (t[i] != '-') return false;
      i++;
      continue;
    }
    const int a = hexval(t[i]), b = hexval(t[i + 1]);
    if (a < 0 || b < 0) return false;
    out[k++] = (uint8_t)(a * 16 + b);
    i += 2;
  }
  return k == 16;
}

int load_from(tir_ctx *ctx, void *db, uint64_t *n_audio, uint64_t *n_rows, uint64_t *n_skipped) {
  Sqlite &s = sq();
  void *st = nullptr;
  if (s.prepare_v2(db, "select audio_uuid, max1, max2 from audio_fingerprint order by audio_uuid, frame_idx", -1, &st,
                   nullptr) != kSqliteOk)
    return tir_fail(ctx, TIR_ERR_STATE, "sqlite: %s", s.errmsg(db));
  std::vector<uint8_t> uu;
  std::vector<uint64_t> off(1, 0);
  std::vector<int32_t> v1, v2;
  uint64_t skipped = 0;
  uint8_t cur[16], last[16];
  bool have_last = false;
  int rc;
  while ((rc = s.step(st)) == kSqliteRow) {
    const char *t = (const char *)s.column_text(st, 0);
    if (!parse_uuid(t, cur)) { // K2: a failed insert can leave a file name where a uuid belongs (src/fp_handler.c:192)
      skipped++;
      continue;
    }
    if (!have_last || std::memcmp(cur, last, 16) != 0) {
      if (have_last) off.push_back(v1.size());
      uu.insert(uu.end(), cur, cur + 16);
      std::memcpy(last, cur, 16), have_last = true;
    }
    v1.push_back(s.column_type(st, 1) == kSqliteNull ? TIR_NULL_V : tir_quantize_micro(s.column_double(st, 1)));
    v2.push_back(s.column_type(st, 2) == kSqliteNull ? TIR_NULL_V : tir_quantize_micro(s.column_double(st, 2)));
  }
  s.finalize(st);
  if (rc != kSqliteDone) return tir_fail(ctx, TIR_ERR_STATE, "sqlite: %s", s.errmsg(db));
  if (have_last) off.push_back(v1.size());
  const uint32_t n = (uint32_t)(uu.size() / 16);
  if (n_audio) *n_audio = n;
  if (n_rows) *n_rows = v1.size();
  if (n_skipped) *n_skipped = skipped;
  return tir_db_load(ctx, n, (const uint8_t(*)[16])uu.data(), off.data(), v1.data(), v2.data());
}

} // namespace

extern "C" {

int tir_db_load_sqlite(tir_ctx *ctx, void *sqlite3_db, uint64_t *n_audio, uint64_t *n_rows, uint64_t *n_skipped) {
  if (!ctx || !sqlite3_db) return tir_fail(ctx, TIR_ERR_ARG, "null argument");
  if (!sq().ok) return tir_fail(ctx, TIR_ERR_STATE, "libsqlite3 could not be resolved");
  return load_from(ctx, sqlite3_db, n_audio, n_rows, n_skipped);
}

int tir_db_load_sqlite_file(tir_ctx *ctx, const char *path, uint64_t *n_audio, uint64_t *n_rows, uint64_t *n_skipped) {
  if (!ctx || !path) return tir_fail(ctx, TIR_ERR_ARG, "null argument");
  Sqlite &s = sq();
  if (!s.ok) return tir_fail(ctx, TIR_ERR_STATE, "libsqlite3 could not be resolved");
  void *db = nullptr;
  if (s.open_v2(path, &db, kOpenReadOnly, nullptr) != kSqliteOk) {
    const int rc = tir_fail(ctx, TIR_ERR_STATE, "sqlite: cannot open %s: %s", path, db ? s.errmsg(db) : "?");
    if (db) s.close(db);
    return rc;
  }
  const int rc = load_from(ctx, db, n_audio, n_rows, n_skipped);
  s.close(db);
  return rc;
}

int tir_sqlite_insert_fingerprints(tir_ctx *ctx, void *sqlite3_db, const char *context, const char *audio_uuid,
                                   const int32_t *vq, uint32_t n_frames) {
  if (!sqlite3_db || !context || !audio_uuid || (!vq && n_frames)) return tir_fail(ctx, TIR_ERR_ARG, "null argument");
  Sqlite &s = sq();
  if (!s.ok) return tir_fail(ctx, TIR_ERR_STATE, "libsqlite3 could not be resolved");
  void *db = sqlite3_db, *st = nullptr;
  if (s.exec(db, "begin;", nullptr, nullptr, nullptr) != kSqliteOk) return tir_fail(ctx, TIR_ERR_STATE, "sqlite: %s", s.errmsg(db));
  int rc = s.prepare_v2(db, "insert into audio_fingerprint(frame_idx, audio_uuid, max1, max2, context) values (?1, ?2, ?3, ?4, ?5);",
                        -1, &st, nullptr);
  for (uint32_t f = 0; rc == kSqliteOk && f < n_frames; f++) {
    s.bind_int64(st, 1, (long long)f);
    s.bind_text(st, 2, audio_uuid, -1, nullptr);
    for (int j = 0; j < TIR_N_COEFS; j++) {
      const int32_t v = vq[(size_t)f * TIR_N_COEFS + j];
      if (v == TIR_NULL_V) s.bind_null(st, 3 + j);
      else s.bind_double(st, 3 + j, (double)v / 1000000.0); // what SQLite makes of the "%f" text
    }
    s.bind_text(st, 5, context, -1, nullptr);
    rc = s.step(st) == kSqliteDone ? kSqliteOk : 1;
    s.reset(st);
  }
  if (st) s.finalize(st);
  if (rc != kSqliteOk) {
    const int e = tir_fail(ctx, TIR_ERR_STATE, "sqlite: %s", s.errmsg(db));
    s.exec(db, "rollback;", nullptr, nullptr, nullptr);
    return e;
  }
  if (s.exec(db, "commit;", nullptr, nullptr, nullptr) != kSqliteOk) return tir_fail(ctx, TIR_ERR_STATE, "sqlite: %s", s.errmsg(db));
  return TIR_OK;
}

} // extern "C"
