/* Minimal declarations of the libsqlite3.so.0 functions src/db_ctx_handler.c calls (test infrastructure; the image
 * has the library, 3.45.1, but not its header).  Prototypes and constants as documented at sqlite.org/c3ref. */
#ifndef FAKE_SQLITE3_H_
#define FAKE_SQLITE3_H_
typedef struct sqlite3 sqlite3;
typedef struct sqlite3_stmt sqlite3_stmt;
typedef struct sqlite3_backup sqlite3_backup;
typedef long long sqlite3_int64;
#define SQLITE_OK 0
#define SQLITE_ERROR 1
#define SQLITE_BUSY 5
#define SQLITE_LOCKED 6
#define SQLITE_ROW 100
#define SQLITE_DONE 101
#define SQLITE_INTEGER 1
#define SQLITE_FLOAT 2
#define SQLITE_TEXT 3
#define SQLITE3_TEXT 3
#define SQLITE_BLOB 4
#define SQLITE_NULL 5
int sqlite3_open(const char *filename, sqlite3 **ppDb);
int sqlite3_close(sqlite3 *);
int sqlite3_exec(sqlite3 *, const char *sql, int (*callback)(void *, int, char **, char **), void *, char **errmsg);
void sqlite3_free(void *);
char *sqlite3_mprintf(const char *, ...);
const char *sqlite3_errmsg(sqlite3 *);
int sqlite3_errcode(sqlite3 *);
int sqlite3_busy_handler(sqlite3 *, int (*)(void *, int), void *);
int sqlite3_busy_timeout(sqlite3 *, int ms);
int sqlite3_prepare_v2(sqlite3 *db, const char *zSql, int nByte, sqlite3_stmt **ppStmt, const char **pzTail);
int sqlite3_step(sqlite3_stmt *);
int sqlite3_finalize(sqlite3_stmt *);
int sqlite3_reset(sqlite3_stmt *);
int sqlite3_column_count(sqlite3_stmt *);
const char *sqlite3_column_name(sqlite3_stmt *, int N);
int sqlite3_column_type(sqlite3_stmt *, int iCol);
int sqlite3_column_int(sqlite3_stmt *, int iCol);
sqlite3_int64 sqlite3_column_int64(sqlite3_stmt *, int iCol);
double sqlite3_column_double(sqlite3_stmt *, int iCol);
const unsigned char *sqlite3_column_text(sqlite3_stmt *, int iCol);
int sqlite3_column_bytes(sqlite3_stmt *, int iCol);
sqlite3_backup *sqlite3_backup_init(sqlite3 *pDest, const char *zDestName, sqlite3 *pSource, const char *zSourceName);
int sqlite3_backup_step(sqlite3_backup *p, int nPage);
int sqlite3_backup_finish(sqlite3_backup *p);
int sqlite3_backup_remaining(sqlite3_backup *p);
int sqlite3_backup_pagecount(sqlite3_backup *p);
int sqlite3_sleep(int);
int sqlite3_changes(sqlite3 *);
int sqlite3_threadsafe(void);
#endif
