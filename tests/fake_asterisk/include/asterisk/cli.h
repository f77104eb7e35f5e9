/* fake <asterisk/cli.h> (test infrastructure) */
#ifndef FAKE_AST_CLI_H_
#define FAKE_AST_CLI_H_
#define CLI_SUCCESS ((char *)0)
#define CLI_SHOWUSAGE ((char *)1)
#define CLI_FAILURE ((char *)2)
enum ast_cli_command { CLI_INIT = -2, CLI_GENERATE = -3, CLI_HANDLER = -1 };
struct ast_cli_args {
  int fd;
  int argc;
  const char *const *argv;
  const char *line, *word;
  int pos, n;
};
struct ast_cli_entry {
  const char *cmda[16];
  const char *summary;
  const char *usage;
  const char *command;
  char *(*handler)(struct ast_cli_entry *e, int cmd, struct ast_cli_args *a);
};
#define AST_CLI_DEFINE(fn, txt, ...) {.handler = fn, .summary = txt, ##__VA_ARGS__}
void ast_cli(int fd, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
int ast_cli_register_multiple(struct ast_cli_entry *e, int len);
int ast_cli_unregister_multiple(struct ast_cli_entry *e, int len);
#endif
