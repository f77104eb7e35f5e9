/* fake <asterisk/file.h> (test infrastructure): "wav" file streams */
#ifndef FAKE_AST_FILE_H_
#define FAKE_AST_FILE_H_
#include <sys/types.h>
struct ast_filestream;
struct ast_frame;
#define AST_FILE_MODE 0666
struct ast_filestream *ast_writefile(const char *filename, const char *type, const char *comment, int flags, int check, mode_t mode);
int ast_writestream(struct ast_filestream *fs, struct ast_frame *f);
int ast_closestream(struct ast_filestream *f);
int ast_filedelete(const char *filename, const char *fmt);
#endif
