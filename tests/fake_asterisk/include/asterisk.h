/*
 * Fake <asterisk.h> -- TEST INFRASTRUCTURE, not product code.
 *
 * Asterisk is not installed in this image.  This tree declares exactly the ast_* / pbx_* surface that the
 * reference's Asterisk-facing files use (src/application_handler.c, src/cli_handler.c, src/app_tiresias.c,
 * src/db_ctx_handler.c; list in SURVEY.md 8b), so that those files compile UNCHANGED, from where they lie
 * under /root/reference/src, against the replacement asterisk_tiresias_b200/host/fp_handler.c.  The
 * implementations are in ../fake_asterisk.c (ast_json_* over the real libjansson.so.4).  Written from the
 * documented behaviour of the Asterisk API; no Asterisk source is copied.
 */
#ifndef FAKE_ASTERISK_H_
#define FAKE_ASTERISK_H_

#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <alloca.h>
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

/* The module shell creates its library directory under /var/lib/asterisk (src/app_tiresias.c:40,153-169):
 * the fake maps that prefix into the test's scratch directory ($FAKE_AST_ROOT).  The macros are defined
 * before <dirent.h> / <sys/stat.h> are seen, so those headers declare the fakes. */
#define opendir fake_ast_opendir
#define mkdir fake_ast_mkdir

#define ARRAY_LEN(a) (sizeof(a) / sizeof((a)[0]))
#define ASTERISK_GPL_KEY "fake-asterisk-test-harness"

/* memory / string helpers: in Asterisk these come with <asterisk.h> (astmm.h) and <asterisk/utils.h> */
#define ast_calloc(n, s) calloc((n), (s))
#define ast_malloc(s) malloc((s))
#define ast_free(p) free((p))
#define ast_strdup(s) strdup((s))
#define ast_strdupa(s) strcpy((char *)alloca(strlen((s)) + 1), (s))
#define ast_asprintf(ret, ...) ((void)(asprintf((ret), __VA_ARGS__) < 0 ? (*(ret) = NULL, 0) : 0))

struct ast_flags {
  unsigned int flags;
};

#endif /* FAKE_ASTERISK_H_ */
