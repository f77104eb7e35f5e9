/* fake <asterisk/app.h> (test infrastructure): standard application argument parsing */
#ifndef FAKE_AST_APP_H_
#define FAKE_AST_APP_H_
#include <stddef.h>
#define AST_APP_ARG(name) char *name
#define AST_DECLARE_APP_ARGS(name, arglist) \
  struct {                                  \
    unsigned int argc;                      \
    char *argv[0];                          \
    arglist                                 \
  } name = {0}
unsigned int fake_ast_separate_args(char *buf, char delim, char **array, int arraylen);
#define AST_STANDARD_APP_ARGS(args, parse) \
  args.argc = fake_ast_separate_args((parse), ',', args.argv, (int)((sizeof(args) - offsetof(__typeof__(args), argv)) / sizeof(args.argv[0])))
#endif
