/* fake <asterisk/strings.h> (test infrastructure) */
#ifndef FAKE_AST_STRINGS_H_
#define FAKE_AST_STRINGS_H_
static inline int ast_strlen_zero(const char *s) { return (!s || *s == '\0'); }
#endif
