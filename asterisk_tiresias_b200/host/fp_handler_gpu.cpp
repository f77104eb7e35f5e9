// fp_handler_gpu.cpp -- host-side mirror of the reference's src/fp_handler.c on top of the C ABI
// of libtiresias_gpu.so (see include/fp_handler_gpu.h for what differs and why).
//
// What stays the reference's: the in-memory SQLite database with its three tables and their SQL
// text (src/fp_handler.c:686-753), the audio_list bookkeeping (uuid, basename, md5, duplicate check
// on (context, hash), :481-536), the delete order (:115-159), the backup file (:68-108).
// What is replaced: create_audio_fingerprints' aubio hop loop (:577-671) -> tir_extract;
// the per-frame textual INSERTs (:559-571) -> tir_sqlite_insert_fingerprints + tir_db_add;
// the SQL probe/tally of fp_search_fingerprint_info (:258-377) -> tir_match.
#include "../../include/fp_handler_gpu.h"

#include <dirent.h>
#include <libgen.h>
#include <openssl/evp.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tiresias_gpu.h"
#include "../csrc/tir_sqlite_dl.h"

namespace {

enum { LOG_DEBUG_ = 0, LOG_NOTICE_ = 2, LOG_WARNING_ = 3, LOG_ERROR_ = 4 };
void (*g_log)(int, const char *) = nullptr;

void ast_log_(int level, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (g_log) g_log(level, buf);
  else if (level >= LOG_NOTICE_) fputs(buf, stderr);
}

struct Host {
  void *db = nullptr;                   // g_db_ctx->db, ":memory:"
  std::string backup;
  int device = 0;
  std::map<int, tir_ctx *> plans;       // extraction plan per sample rate (the filterbank follows the file, :612-615)
  tir_ctx *main = nullptr;              // owns the device mirror of audio_fingerprint
  std::mutex mu;
} g;

bool exec(const char *sql) {
  TirSqlite &s = tir_sqlite();
  char *err = nullptr;
  if (s.exec(g.db, sql, nullptr, nullptr, &err) != kSqliteOk) {
    ast_log_(LOG_ERROR_, "Could not execute query. sql[%.200s], err[%s]\n", sql, s.errmsg(g.db));
    return false;
  }
  return true;
}

std::string fmt(const char *f, ...) {
  char buf[2048];
  va_list ap;
  va_start(ap, f);
  vsnprintf(buf, sizeof buf, f, ap);
  va_end(ap);
  return buf;
}

void copy_field(char *dst, size_t cap, const unsigned char *src) { snprintf(dst, cap, "%s", src ? (const char *)src : ""); }

// "select * from audio_list where ..." -> rows (columns uuid, name, context, hash: :700-706)
int query_audio(const std::string &sql, fp_audio_info *out, int cap) {
  TirSqlite &s = tir_sqlite();
  void *st = nullptr;
  if (s.prepare_v2(g.db, sql.c_str(), -1, &st, nullptr) != kSqliteOk) return -1;
  int n = 0;
  while (s.step(st) == kSqliteRow) {
    if (out && n < cap) {
      fp_audio_info &a = out[n];
      memset(&a, 0, sizeof a);
      copy_field(a.uuid, sizeof a.uuid, s.column_text(st, 0));
      copy_field(a.name, sizeof a.name, s.column_text(st, 1));
      copy_field(a.context, sizeof a.context, s.column_text(st, 2));
      copy_field(a.hash, sizeof a.hash, s.column_text(st, 3));
    }
    n++;
  }
  s.finalize(st);
  return n;
}

bool get_audio_list_info(const char *uuid, fp_audio_info *out) { // :836-858
  return query_audio(fmt("select * from audio_list where uuid = '%s';", uuid), out, 1) > 0;
}

bool parse_uuid16(const char *t, uint8_t out[16]) {
  if (!t || strlen(t) != 36) return false;
  int k = 0;
  for (int i = 0; i < 36;) {
    if (i == 8 || i == 13 || i == 18 || i == 23) { if (t[i++] != '-') return false; continue; }
    unsigned v;
    if (sscanf(t + i, "%2x", &v) != 1) return false;
    out[k++] = (uint8_t)v, i += 2;
  }
  return k == 16;
}

void unparse_uuid16(const uint8_t u[16], char out[37]) {
  snprintf(out, 37, "%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x", u[0], u[1], u[2], u[3], u[4], u[5],
           u[6], u[7], u[8], u[9], u[10], u[11], u[12], u[13], u[14], u[15]);
}

// What aubio_source delivers for the files this module sees (ast_writefile(..., "wav") recordings and
// WAV directories): RIFF/WAVE, PCM, 16 bit.  `pcm` keeps the channels interleaved as they lie in the file: aubio's
// source averages them in float, which tir_extract_interleaved reproduces on the device.
bool read_wav_pcm16(const char *filename, std::vector<int16_t> &pcm, int &rate, int &channels) {
  FILE *f = fopen(filename, "rb");
  if (!f) {
    ast_log_(LOG_WARNING_, "Could not open file. filename[%s]\n", filename);
    return false;
  }
  unsigned char h[12];
  bool ok = fread(h, 1, 12, f) == 12 && !memcmp(h, "RIFF", 4) && !memcmp(h + 8, "WAVE", 4);
  int bits = 0, format = 0;
  rate = 0, channels = 0;
  bool have_fmt = false, have_data = false;
  while (ok && !have_data) {
    unsigned char ch[8];
    if (fread(ch, 1, 8, f) != 8) break;
    const uint32_t len = ch[4] | (ch[5] << 8) | (ch[6] << 16) | ((uint32_t)ch[7] << 24);
    if (!memcmp(ch, "fmt ", 4)) {
      unsigned char b[16];
      if (len < 16 || fread(b, 1, 16, f) != 16) { ok = false; break; }
      format = b[0] | (b[1] << 8), channels = b[2] | (b[3] << 8);
      rate = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24), bits = b[14] | (b[15] << 8);
      have_fmt = true;
      fseek(f, (long)(len - 16 + (len & 1)), SEEK_CUR);
    } else if (!memcmp(ch, "data", 4)) {
      if (!have_fmt) { ok = false; break; }
      const long pos = ftell(f);
      fseek(f, 0, SEEK_END);
      const long rest = ftell(f) - pos;
      fseek(f, pos, SEEK_SET);
      const size_t bytes = (size_t)std::min<long>(rest, (long)len);
      pcm.resize(bytes / 2);
      ok = fread(pcm.data(), 2, pcm.size(), f) == pcm.size();
      have_data = true;
    } else {
      fseek(f, (long)(len + (len & 1)), SEEK_CUR);
    }
  }
  fclose(f);
  if (!ok || !have_data || format != 1 || bits != 16 || channels < 1 || channels > 64 || rate <= 0) {
    ast_log_(LOG_ERROR_, "Could not read the file as PCM16 WAV. filename[%s]\n", filename);
    return false;
  }
  pcm.resize(pcm.size() / (size_t)channels * (size_t)channels); // whole sample frames only
  return true;
}

tir_ctx *plan_for_rate(int rate) {
  auto it = g.plans.find(rate);
  if (it != g.plans.end()) return it->second;
  tir_cfg cfg;
  tir_cfg_default(&cfg); // 512 / 256 / 40: DEF_AUBIO_BUFSIZE / HOPSIZE / FILTER, :33-37
  cfg.device = g.device, cfg.samplerate = rate;
  tir_ctx *ctx = nullptr;
  if (tir_open(&cfg, &ctx) != TIR_OK) {
    ast_log_(LOG_ERROR_, "Could not open the GPU context. err[%s]\n", tir_last_error(ctx));
    tir_close(ctx);
    return nullptr;
  }
  g.plans[rate] = ctx;
  return ctx;
}

// create_audio_fingerprints(): file -> per-frame coefficients and "%f" micro-units (:577-671)
bool create_audio_fingerprints(const char *filename, std::vector<float> &coef, std::vector<int32_t> &vq) {
  std::vector<int16_t> pcm;
  int rate, channels;
  if (!read_wav_pcm16(filename, pcm, rate, channels)) return false;
  tir_ctx *ctx = plan_for_rate(rate);
  if (!ctx) return false;
  const uint64_t off[2] = {0, pcm.size() / (size_t)channels};
  uint64_t frames = tir_n_frames(off[1], 256);
  coef.assign(frames * 2, 0.f), vq.assign(frames * 2, 0);
  if (tir_extract_interleaved(ctx, pcm.data(), channels, off, 1, coef.data(), vq.data(), &frames) != TIR_OK) {
    ast_log_(LOG_ERROR_, "Could not create fingerprint data. err[%s]\n", tir_last_error(ctx));
    return false;
  }
  return true;
}

char *create_file_hash(const char *filename) { // :758-805 (MD5 of the file bytes, lower-case hex)
  FILE *file = fopen(filename, "rb");
  if (!file) {
    ast_log_(LOG_WARNING_, "Could not open file. filename[%s]\n", filename);
    return nullptr;
  }
  EVP_MD_CTX *md = EVP_MD_CTX_new();
  EVP_DigestInit_ex(md, EVP_md5(), nullptr);
  unsigned char data[1024], hash[EVP_MAX_MD_SIZE];
  size_t n;
  while ((n = fread(data, 1, sizeof data, file)) > 0) EVP_DigestUpdate(md, data, n);
  fclose(file);
  unsigned len = 0;
  EVP_DigestFinal_ex(md, hash, &len);
  EVP_MD_CTX_free(md);
  char *res = (char *)malloc(2 * len + 1);
  for (unsigned i = 0; i < len; i++) sprintf(res + 2 * i, "%02x", hash[i]);
  return res;
}

bool create_tables() { // init_database(), :673-756 (same SQL text)
  bool ok = exec("create table context_list("
                 "   name        varchar(255),"
                 "   directory   varchar(1023),"
                 "   primary key(name)"
                 ");");
  ok = ok && exec("create table audio_list("
                  "   uuid           varchar(255),"
                  "   name           varchar(255),"
                  "   context        varchar(255),"
                  "	hash           varchar(1023)"
                  ");");
  ok = ok && exec("create table audio_fingerprint("
                  " context        varchar(255),"
                  " audio_uuid     varchar(255),"
                  " frame_idx      integer, max1 real, max2 real);");
  ok = ok && exec("create index idx_audio_fingerprint_context on audio_fingerprint(context);");
  ok = ok && exec("create index idx_audio_fingerprint_max1 on audio_fingerprint(max1);");
  ok = ok && exec("create index idx_audio_fingerprint_max2 on audio_fingerprint(max2);");
  return ok;
}

} // namespace

extern "C" {

void fp_set_log(void (*fn)(int, const char *)) { g_log = fn; }
void *fp_sqlite_handle(void) { return g.db; }

static bool fp_init_locked(const char *backup_db, int device);

static void reset_state() {
  for (auto &kv : g.plans) tir_close(kv.second);
  g.plans.clear(), g.main = nullptr;
  if (g.db) tir_sqlite().close(g.db);
  g.db = nullptr;
}

bool fp_init(const char *backup_db, int device) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (g.db) return true;
  if (fp_init_locked(backup_db, device)) return true;
  reset_state(); // a failed load leaves nothing behind: the module declines to load
  return false;
}

static bool fp_init_locked(const char *backup_db, int device) {
  TirSqlite &s = tir_sqlite();
  if (!s.ok) {
    ast_log_(LOG_ERROR_, "Could not initiate database. err[libsqlite3 not found]\n");
    return false;
  }
  g.device = device, g.backup = backup_db ? backup_db : "";
  if (s.open(":memory:", &g.db) != kSqliteOk || !create_tables()) {
    ast_log_(LOG_ERROR_, "Could not initiate database.\n");
    return false;
  }
  if (!g.backup.empty()) { // db_ctx_load_db_data(): ATTACH + copy every table, src/db_ctx_handler.c:745-772
    FILE *f = fopen(g.backup.c_str(), "rb");
    if (f) {
      fclose(f);
      bool ok = exec(fmt("ATTACH DATABASE '%s' as backup", g.backup.c_str()).c_str());
      ok = ok && exec("BEGIN");
      for (const char *t : {"context_list", "audio_list", "audio_fingerprint"})
        ok = ok && exec(fmt("insert into main.'%s' select * from backup.'%s'", t, t).c_str());
      ok = ok && exec("COMMIT");
      exec("DETACH DATABASE backup");
      if (!ok) {
        ast_log_(LOG_ERROR_, "Could not load the database data.\n");
        return false;
      }
    }
  }
  g.main = plan_for_rate(8000);
  if (!g.main) return false; // no CPU fallback: the module declines to load
  uint64_t na = 0, nr = 0, skipped = 0;
  if (tir_db_load_sqlite(g.main, g.db, &na, &nr, &skipped) != TIR_OK) {
    ast_log_(LOG_ERROR_, "Could not mirror the fingerprints to the device. err[%s]\n", tir_last_error(g.main));
    return false;
  }
  ast_log_(LOG_DEBUG_, "Mirrored fingerprints. audios[%llu], rows[%llu], skipped[%llu]\n", (unsigned long long)na,
           (unsigned long long)nr, (unsigned long long)skipped);
  return true;
}

bool fp_term(void) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.db) return false;
  TirSqlite &s = tir_sqlite();
  bool ok = true;
  if (!g.backup.empty()) { // db_ctx_backup(), src/db_ctx_handler.c:673-714
    void *file = nullptr;
    ok = s.open(g.backup.c_str(), &file) == kSqliteOk;
    void *bk = ok ? s.backup_init(file, "main", g.db, "main") : nullptr;
    if (bk) {
      while (s.backup_step(bk, 5) == kSqliteOk) {}
      s.backup_finish(bk);
    } else {
      ok = false;
    }
    if (file) s.close(file);
    if (!ok) ast_log_(LOG_ERROR_, "Could not write database.\n");
  }
  reset_state();
  return ok;
}

bool fp_create_context_list_info(const char *name, const char *directory, bool replace) { // :936-1001
  if (!name || !directory) {
    ast_log_(LOG_WARNING_, "Wrong input parameter.\n");
    return false;
  }
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.db) return false;
  return exec(fmt("insert %sinto context_list(name, directory) values ('%s', '%s');", replace ? "or replace " : "", name, directory).c_str());
}

bool fp_get_context_list_info(const char *name, fp_context_info *out) {
  if (!name) {
    ast_log_(LOG_WARNING_, "Wrong input parameter.\n");
    return false;
  }
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.db) return false;
  TirSqlite &s = tir_sqlite();
  void *st = nullptr;
  if (s.prepare_v2(g.db, fmt("select * from context_list where name = '%s';", name).c_str(), -1, &st, nullptr) != kSqliteOk) return false;
  const bool found = s.step(st) == kSqliteRow;
  if (found && out) {
    copy_field(out->name, sizeof out->name, s.column_text(st, 0));
    copy_field(out->directory, sizeof out->directory, s.column_text(st, 1));
  }
  s.finalize(st);
  return found;
}

int fp_get_context_lists_all(fp_context_info *out, int cap) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.db) return -1;
  TirSqlite &s = tir_sqlite();
  void *st = nullptr;
  if (s.prepare_v2(g.db, "select * from context_list;", -1, &st, nullptr) != kSqliteOk) return -1;
  int n = 0;
  while (s.step(st) == kSqliteRow) {
    if (out && n < cap) {
      copy_field(out[n].name, sizeof out[n].name, s.column_text(st, 0));
      copy_field(out[n].directory, sizeof out[n].directory, s.column_text(st, 1));
    }
    n++;
  }
  s.finalize(st);
  return n;
}

int fp_get_audio_lists_all(fp_audio_info *out, int cap) {
  std::lock_guard<std::mutex> lk(g.mu);
  return g.db ? query_audio("select * from audio_list;", out, cap) : -1;
}

int fp_get_audio_lists_by_contextname(const char *name, fp_audio_info *out, int cap) {
  if (!name) {
    ast_log_(LOG_WARNING_, "Wrong input parameter.\n");
    return -1;
  }
  std::lock_guard<std::mutex> lk(g.mu);
  return g.db ? query_audio(fmt("select * from audio_list where context = '%s';", name), out, cap) : -1;
}

static bool delete_audio_locked(const char *uuid) { // fp_delete_audio_list_info, :115-159
  fp_audio_info a;
  if (!get_audio_list_info(uuid, &a)) {
    ast_log_(LOG_NOTICE_, "Could not find audio list info.\n");
    return false;
  }
  if (!exec(fmt("delete from audio_list where uuid='%s';", uuid).c_str())) return false;
  if (!exec(fmt("delete from audio_fingerprint where audio_uuid='%s';", uuid).c_str())) return false;
  uint8_t u16[16];
  if (parse_uuid16(uuid, u16)) {
    const int rc = tir_db_remove(g.main, u16);
    if (rc != TIR_OK && rc != TIR_ERR_NOTFOUND) { // an audio without frames has no device rows
      ast_log_(LOG_WARNING_, "Could not delete audio fingerprint info. audio_uuid[%s]\n", uuid);
      return false;
    }
  }
  return true;
}

bool fp_delete_audio_list_info(const char *uuid) {
  if (!uuid) {
    ast_log_(LOG_WARNING_, "Wrong input parameter.\n");
    return false;
  }
  std::lock_guard<std::mutex> lk(g.mu);
  return g.db && delete_audio_locked(uuid);
}

bool fp_delete_context_list_info(const char *name) { // :1039-1095: the audios of the context first, then the context
  if (!name) {
    ast_log_(LOG_WARNING_, "Wrong input parameter.\n");
    return false;
  }
  fp_context_info c;
  if (!fp_get_context_list_info(name, &c)) {
    ast_log_(LOG_NOTICE_, "Could not find context info. context[%s]\n", name);
    return false;
  }
  std::lock_guard<std::mutex> lk(g.mu);
  const int n = query_audio(fmt("select * from audio_list where context = '%s';", name), nullptr, 0);
  std::vector<fp_audio_info> v((size_t)std::max(n, 0));
  query_audio(fmt("select * from audio_list where context = '%s';", name), v.data(), n);
  for (const fp_audio_info &a : v)
    if (!delete_audio_locked(a.uuid)) ast_log_(LOG_WARNING_, "Could not delete audio_list info. uuid[%s]\n", a.uuid);
  return exec(fmt("delete from context_list where name='%s';", name).c_str());
}

bool fp_craete_audio_list_info(const char *context, const char *filename) { // :161-205
  if (!context || !filename) {
    ast_log_(LOG_WARNING_, "Wrong input parameter.\n");
    return false;
  }
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.db) return false;
  char *uuid = fp_generate_uuid();
  std::string u = uuid;
  free(uuid);
  // create_audio_list_info(), :481-536
  char *hash = create_file_hash(filename);
  if (!hash) {
    ast_log_(LOG_WARNING_, "Could not create audio_list info. context[%s], filename[%s]\n", context, filename);
    return false;
  }
  const std::string h = hash;
  free(hash);
  if (query_audio(fmt("select * from audio_list where context = '%s' and hash = '%s';", context, h.c_str()), nullptr, 0) > 0)
    return true; // "The given audio file is already exist in the list" (P7)
  std::string tmp = filename;
  const char *name = basename(&tmp[0]);
  if (!exec(fmt("insert into audio_list(uuid, name, context, hash) values ('%s', '%s', '%s', '%s');", u.c_str(), name, context, h.c_str()).c_str()))
    return false;
  // create_audio_fingerprint_info(), :538-575
  std::vector<float> coef;
  std::vector<int32_t> vq;
  uint8_t u16[16];
  bool ok = create_audio_fingerprints(filename, coef, vq) && parse_uuid16(u.c_str(), u16);
  const uint32_t frames = (uint32_t)(vq.size() / 2);
  ok = ok && tir_sqlite_insert_fingerprints(g.main, g.db, context, u.c_str(), vq.data(), frames) == TIR_OK;
  if (ok) {
    std::vector<int32_t> v1(frames), v2(frames);
    for (uint32_t f = 0; f < frames; f++) v1[f] = vq[2 * f], v2[f] = vq[2 * f + 1];
    ok = tir_db_add(g.main, u16, v1.data(), v2.data(), frames) == TIR_OK;
  }
  if (!ok) {
    ast_log_(LOG_NOTICE_, "Could not create audio fingerprint info.\n");
    // the reference calls fp_delete_audio_list_info(filename) here (K2: a file name, which never
    // matches a uuid, so the audio_list row stays); same observable state:
    return false;
  }
  return true;
}

bool fp_search_fingerprint_info(const char *context, const char *filename, const int coefs, const double tolerance,
                                const int freq_ignore_low, const int freq_ignore_high, fp_audio_info *out) {
  if (!context || !filename) { // :234-238
    ast_log_(LOG_WARNING_, "Wrong input parameter.\n");
    return false;
  }
  if (coefs < 1 || coefs > TIR_N_COEFS) { // :247-250
    ast_log_(LOG_WARNING_, "Wrong coefs count. max[%d], coefs[%d]\n", TIR_N_COEFS, coefs);
    return false;
  }
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.db) return false;
  std::vector<float> coef;
  std::vector<int32_t> vq;
  if (!create_audio_fingerprints(filename, coef, vq)) { // :274-279
    ast_log_(LOG_ERROR_, "Could not create fingerprint info.\n");
    return false;
  }
  const uint64_t frames = coef.size() / 2;
  std::vector<double> y(frames * 2);
  for (size_t i = 0; i < y.size(); i++) y[i] = 10 * log10(fabs((double)coef[i])); // :651; inf/nan = key missing
  const uint64_t foff[2] = {0, frames};
  tir_hit hit;
  if (tir_match(g.main, y.data(), foff, 1, coefs, tolerance, freq_ignore_low, freq_ignore_high, &hit) != TIR_OK) {
    ast_log_(LOG_WARNING_, "Could not search. err[%s]\n", tir_last_error(g.main));
    return false;
  }
  if (hit.match_count == 0) { // :386-390
    ast_log_(LOG_NOTICE_, "Could not find data.\n");
    return false;
  }
  char uuid[37];
  unparse_uuid16(hit.uuid, uuid);
  fp_audio_info a;
  if (!get_audio_list_info(uuid, &a)) { // :394-398
    ast_log_(LOG_NOTICE_, "Could not find audio list info. uuid[%s]\n", uuid);
    return false;
  }
  a.frame_count = hit.frame_count, a.match_count = hit.match_count; // :403-404
  if (out) *out = a;
  return true;
}

// init_audio() of the module shell (src/app_tiresias.c:324-551), which is where bulk fingerprinting
// happens: for every context, audios whose file is gone are deleted (delete_removed_audio_info:
// directory hashes vs audio_list.hash), files not yet listed are fingerprinted (create_new_audio_info
// -> fp_craete_audio_list_info per file, scandir + alphasort, "." and ".." skipped).  Same outcome,
// but the new files of a context go through ONE batched extraction launch per sample rate instead of
// one aubio pass per file.  Returns the number of audios added, -1 on error.
static int sync_context_locked(const fp_context_info &c) {
  struct dirent **namelist = nullptr;
  const int count = scandir(c.directory, &namelist, nullptr, alphasort);
  if (count < 0) {
    ast_log_(LOG_NOTICE_, "Could not get directory list info. context[%s]\n", c.name);
    return -1;
  }
  std::vector<std::string> files, hashes;
  for (int i = 0; i < count; i++) {
    const std::string nm = namelist[i]->d_name;
    free(namelist[i]);
    if (nm == "." || nm == "..") continue; // file_select, :552-572
    const std::string path = std::string(c.directory) + "/" + nm;
    char *h = create_file_hash(path.c_str());
    if (!h) continue;
    files.push_back(path), hashes.push_back(h);
    free(h);
  }
  free(namelist);
  // delete_removed_audio_info(), :431-550
  const std::string q = fmt("select * from audio_list where context = '%s';", c.name);
  const int n_listed = query_audio(q, nullptr, 0);
  std::vector<fp_audio_info> listed((size_t)std::max(n_listed, 0));
  if (n_listed > 0) query_audio(q, listed.data(), n_listed);
  for (const fp_audio_info &a : listed) {
    bool found = false;
    for (const std::string &h : hashes) found = found || h == a.hash;
    if (!found && !delete_audio_locked(a.uuid)) ast_log_(LOG_DEBUG_, "Could not delete audio list info.\n");
  }
  // create_new_audio_info(), :364-424, batched: read every new file, one extraction per sample rate
  struct NewAudio {
    std::string path, hash, uuid;
    std::vector<int16_t> pcm;
    int rate, channels;
  };
  std::vector<NewAudio> fresh;
  for (size_t i = 0; i < files.size(); i++) {
    bool known = false;
    for (const fp_audio_info &a : listed) known = known || hashes[i] == a.hash;
    for (const NewAudio &f : fresh) known = known || f.hash == hashes[i]; // the same bytes twice in one directory (P7)
    if (known) continue;
    NewAudio f;
    f.path = files[i], f.hash = hashes[i];
    char *u = fp_generate_uuid();
    f.uuid = u;
    free(u);
    // the reference inserts the audio_list row before it tries to fingerprint, and leaves it there when
    // fingerprinting fails (K2, src/fp_handler.c:186-195)
    std::string tmp = f.path;
    if (!exec(fmt("insert into audio_list(uuid, name, context, hash) values ('%s', '%s', '%s', '%s');", f.uuid.c_str(),
                  basename(&tmp[0]), c.name, f.hash.c_str()).c_str()))
      continue;
    if (!read_wav_pcm16(f.path.c_str(), f.pcm, f.rate, f.channels)) continue;
    fresh.push_back(std::move(f));
  }
  int added = 0;
  std::map<std::pair<int, int>, std::vector<size_t>> by_rate; // one batched extraction per (rate, channel count)
  for (size_t i = 0; i < fresh.size(); i++) by_rate[{fresh[i].rate, fresh[i].channels}].push_back(i);
  for (auto &kv : by_rate) {
    tir_ctx *ctx = plan_for_rate(kv.first.first);
    if (!ctx) continue;
    const int channels = kv.first.second;
    std::vector<int16_t> pcm;
    std::vector<uint64_t> off(1, 0); // in sample frames
    for (size_t i : kv.second) {
      pcm.insert(pcm.end(), fresh[i].pcm.begin(), fresh[i].pcm.end());
      off.push_back(pcm.size() / (size_t)channels);
    }
    uint64_t frames = 0;
    for (size_t k = 0; k + 1 < off.size(); k++) frames += tir_n_frames(off[k + 1] - off[k], 256);
    std::vector<int32_t> vq(frames * 2);
    if (tir_extract_interleaved(ctx, pcm.data(), channels, off.data(), (uint32_t)kv.second.size(), nullptr, vq.data(), &frames) != TIR_OK) {
      ast_log_(LOG_ERROR_, "Could not create fingerprint data. err[%s]\n", tir_last_error(ctx));
      continue;
    }
    uint64_t f0 = 0;
    for (size_t k = 0; k < kv.second.size(); k++) {
      const NewAudio &f = fresh[kv.second[k]];
      const uint32_t nf = (uint32_t)tir_n_frames(off[k + 1] - off[k], 256);
      const int32_t *v = vq.data() + f0 * 2;
      f0 += nf;
      uint8_t u16[16];
      if (!parse_uuid16(f.uuid.c_str(), u16)) continue;
      if (tir_sqlite_insert_fingerprints(g.main, g.db, c.name, f.uuid.c_str(), v, nf) != TIR_OK) continue;
      std::vector<int32_t> v1(nf), v2(nf);
      for (uint32_t t = 0; t < nf; t++) v1[t] = v[2 * t], v2[t] = v[2 * t + 1];
      if (tir_db_add(g.main, u16, v1.data(), v2.data(), nf) == TIR_OK) added++;
    }
  }
  return added;
}

int fp_sync_directories(void) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.db) return -1;
  TirSqlite &s = tir_sqlite();
  std::vector<fp_context_info> ctxs;
  void *st = nullptr;
  if (s.prepare_v2(g.db, "select * from context_list;", -1, &st, nullptr) != kSqliteOk) return -1;
  while (s.step(st) == kSqliteRow) {
    fp_context_info c;
    copy_field(c.name, sizeof c.name, s.column_text(st, 0));
    copy_field(c.directory, sizeof c.directory, s.column_text(st, 1));
    ctxs.push_back(c);
  }
  s.finalize(st);
  int added = 0;
  for (const fp_context_info &c : ctxs) {
    const int n = sync_context_locked(c);
    if (n > 0) added += n;
  }
  return added;
}

char *fp_generate_uuid(void) { // uuid_generate + uuid_unparse_lower, :1103-1115 (libuuid's header is not installed: RFC 4122 v4 from /dev/urandom)
  uint8_t u[16];
  FILE *f = fopen("/dev/urandom", "rb");
  if (!f || fread(u, 1, 16, f) != 16) {
    for (int i = 0; i < 16; i++) u[i] = (uint8_t)rand();
  }
  if (f) fclose(f);
  u[6] = (u[6] & 0x0f) | 0x40, u[8] = (u[8] & 0x3f) | 0x80;
  char *res = (char *)malloc(37);
  unparse_uuid16(u, res);
  return res;
}

char *fp_create_hash(const char *filename) {
  if (!filename) {
    ast_log_(LOG_WARNING_, "Wrong input parameter.\n");
    return nullptr;
  }
  return create_file_hash(filename);
}

} // extern "C"
