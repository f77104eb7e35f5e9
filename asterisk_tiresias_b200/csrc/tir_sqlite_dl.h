// tir_sqlite_dl.h -- libsqlite3 resolved at run time (its header is not installed here; inside
// Asterisk the library is already loaded): dlopen with RTLD_NOLOAD first, then libsqlite3.so.0.
#pragma once
#include <dlfcn.h>

#include <type_traits>

struct TirSqlite {
  void *lib = nullptr;
  int (*open)(const char *, void **) = nullptr;
  int (*open_v2)(const char *, void **, int, const char *) = nullptr;
  int (*close)(void *) = nullptr;
  int (*exec)(void *, const char *, int (*)(void *, int, char **, char **), void *, char **) = nullptr;
  int (*prepare_v2)(void *, const char *, int, void **, const char **) = nullptr;
  int (*step)(void *) = nullptr;
  int (*reset)(void *) = nullptr;
  int (*finalize)(void *) = nullptr;
  int (*bind_int64)(void *, int, long long) = nullptr;
  int (*bind_double)(void *, int, double) = nullptr;
  int (*bind_null)(void *, int) = nullptr;
  int (*bind_text)(void *, int, const char *, int, void (*)(void *)) = nullptr;
  const unsigned char *(*column_text)(void *, int) = nullptr;
  double (*column_double)(void *, int) = nullptr;
  int (*column_int)(void *, int) = nullptr;
  int (*column_type)(void *, int) = nullptr;
  const char *(*errmsg)(void *) = nullptr;
  void *(*backup_init)(void *, const char *, void *, const char *) = nullptr;
  int (*backup_step)(void *, int) = nullptr;
  int (*backup_finish)(void *) = nullptr;
  bool ok = false;
};

enum { kSqliteOk = 0, kSqliteRow = 100, kSqliteDone = 101, kSqliteNull = 5, kOpenReadOnly = 1 };

inline TirSqlite &tir_sqlite() {
  static TirSqlite s = [] {
    TirSqlite t;
    for (const char *name : {"libsqlite3.so.0", "libsqlite3.so"}) {
      t.lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
      if (!t.lib) t.lib = dlopen(name, RTLD_NOW);
      if (t.lib) break;
    }
    if (!t.lib) return t;
    bool all = true;
    auto sym = [&](auto &fn, const char *n) {
      fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(t.lib, n));
      all = all && fn != nullptr;
    };
    sym(t.open, "sqlite3_open"), sym(t.open_v2, "sqlite3_open_v2"), sym(t.close, "sqlite3_close"), sym(t.exec, "sqlite3_exec");
    sym(t.prepare_v2, "sqlite3_prepare_v2"), sym(t.step, "sqlite3_step"), sym(t.reset, "sqlite3_reset");
    sym(t.finalize, "sqlite3_finalize"), sym(t.bind_int64, "sqlite3_bind_int64"), sym(t.bind_double, "sqlite3_bind_double");
    sym(t.bind_null, "sqlite3_bind_null"), sym(t.bind_text, "sqlite3_bind_text"), sym(t.column_text, "sqlite3_column_text");
    sym(t.column_double, "sqlite3_column_double"), sym(t.column_int, "sqlite3_column_int"), sym(t.column_type, "sqlite3_column_type");
    sym(t.errmsg, "sqlite3_errmsg"), sym(t.backup_init, "sqlite3_backup_init"), sym(t.backup_step, "sqlite3_backup_step");
    sym(t.backup_finish, "sqlite3_backup_finish");
    t.ok = all;
    return t;
  }();
  return s;
}
