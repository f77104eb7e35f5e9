/*
 * fp_handler_gpu.h -- host-side mirror of the reference's src/fp_handler.h on top of
 * libtiresias_gpu.so (libtiresias_host.so, source asterisk_tiresias_b200/host/fp_handler_gpu.cpp).
 *
 * Same function names, argument meaning and error behaviour as src/fp_handler.h:13-38 so that the
 * parity tests read like calls into the reference; two differences forced by this environment
 * (Asterisk is not installed):
 *   - struct ast_json* results become the plain struct fp_audio_info / arrays of it
 *     (a NULL result is `false` / a count of 0; the fields are the JSON keys of
 *     src/fp_handler.c:394-404 and of table audio_list);
 *   - fp_init() reads its database path / device from arguments instead of #defines and the
 *     module configuration (DEF_BACKUP_DATABASE src/fp_handler.c:31).
 * SQLite (schema src/fp_handler.c:686-753, ":memory:" + backup file) stays the system of record;
 * the GPU holds extraction and a mirror of table audio_fingerprint.
 */
#ifndef FP_HANDLER_GPU_H_
#define FP_HANDLER_GPU_H_

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  char uuid[40];     /* audio_list.uuid    */
  char name[256];    /* audio_list.name  (basename of the file, src/fp_handler.c:510) */
  char context[256]; /* audio_list.context */
  char hash[40];     /* audio_list.hash  (md5 of the file, src/fp_handler.c:758-805) */
  int frame_count;   /* search results only, src/fp_handler.c:403 */
  int match_count;   /* search results only, src/fp_handler.c:404 */
} fp_audio_info;

typedef struct {
  char name[256];
  char directory[1024];
} fp_context_info;

/* fp_init(): create the in-memory database (init_database, src/fp_handler.c:673-756), load the backup
 * file if it exists (db_ctx_load_db_data, :82), open the GPU context and mirror audio_fingerprint.
 * backup_db may be NULL (nothing to restore, nothing written by fp_term). */
bool fp_init(const char *backup_db, int device);
bool fp_term(void); /* db_ctx_backup + close, src/fp_handler.c:92-108 */

bool fp_create_context_list_info(const char *name, const char *directory, bool replace);
bool fp_delete_context_list_info(const char *name);
int fp_get_context_lists_all(fp_context_info *out, int cap);  /* returns the count (may exceed cap) */
bool fp_get_context_list_info(const char *name, fp_context_info *out);

int fp_get_audio_lists_all(fp_audio_info *out, int cap);
int fp_get_audio_lists_by_contextname(const char *name, fp_audio_info *out, int cap);

bool fp_craete_audio_list_info(const char *context, const char *filename); /* [sic], src/fp_handler.h:25 */
bool fp_delete_audio_list_info(const char *uuid);

/* false == the reference's NULL (wrong arguments, unreadable file, or nothing matched -> the
 * dialplan sets TIRSTATUS=NOTFOUND, src/application_handler.c:187-191) */
bool fp_search_fingerprint_info(const char *context, const char *filename, const int coefs, const double tolerance,
                                const int freq_ignore_low, const int freq_ignore_high, fp_audio_info *out);

/* init_audio() of the module shell (src/app_tiresias.c:324-551): for every context delete the audios
 * whose file disappeared from its directory, fingerprint the files that are not listed yet -- the new
 * files of a context in ONE batched extraction per sample rate.  Returns the audios added, -1 on error. */
int fp_sync_directories(void);

char *fp_generate_uuid(void);              /* malloc'ed, caller frees */
char *fp_create_hash(const char *filename); /* malloc'ed md5 hex, caller frees */

/* not in the reference: where ast_log() lines go (default: stderr) */
void fp_set_log(void (*fn)(int level, const char *msg));
/* the sqlite3* of the in-memory database (g_db_ctx->db), for tests */
void *fp_sqlite_handle(void);

#ifdef __cplusplus
}
#endif
#endif
