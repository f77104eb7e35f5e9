/* fake <asterisk/pbx.h> (test infrastructure) */
#ifndef FAKE_AST_PBX_H_
#define FAKE_AST_PBX_H_
struct ast_channel;
struct ast_module;
int pbx_builtin_setvar_helper(struct ast_channel *chan, const char *name, const char *value);
int ast_register_application2(const char *app, int (*execute)(struct ast_channel *, const char *), const char *synopsis,
                              const char *description, void *mod);
int ast_unregister_application(const char *app);
#endif
