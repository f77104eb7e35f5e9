/* fake <asterisk/config.h> (test infrastructure): $FAKE_AST_ROOT/etc/asterisk/<name>, "[category]" / "name = value" */
#ifndef FAKE_AST_CONFIG_H_
#define FAKE_AST_CONFIG_H_
#include <asterisk.h>
struct ast_config;
struct ast_variable {
  const char *name, *value;
  struct ast_variable *next;
};
#define CONFIG_STATUS_FILEMISSING ((struct ast_config *)0)
#define CONFIG_STATUS_FILEUNCHANGED ((struct ast_config *)-1)
#define CONFIG_STATUS_FILEINVALID ((struct ast_config *)-2)
struct ast_config *fake_ast_config_load(const char *filename, struct ast_flags flags);
#define ast_config_load(filename, flags) fake_ast_config_load((filename), (flags))
void ast_config_destroy(struct ast_config *cfg);
char *ast_category_browse(struct ast_config *cfg, const char *prev);
struct ast_variable *ast_variable_browse(const struct ast_config *cfg, const char *category);
#endif
