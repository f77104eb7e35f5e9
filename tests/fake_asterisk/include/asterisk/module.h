/* fake <asterisk/module.h> (test infrastructure) */
#ifndef FAKE_AST_MODULE_H_
#define FAKE_AST_MODULE_H_
enum ast_module_load_result { AST_MODULE_LOAD_SUCCESS = 0, AST_MODULE_LOAD_DECLINE = 1, AST_MODULE_LOAD_SKIP = 2, AST_MODULE_LOAD_FAILURE = -1 };
enum ast_module_reload_result { AST_MODULE_RELOAD_SUCCESS = 0 };
enum ast_module_support_level { AST_MODULE_SUPPORT_UNKNOWN, AST_MODULE_SUPPORT_CORE, AST_MODULE_SUPPORT_EXTENDED };
enum { AST_MODFLAG_DEFAULT = 0, AST_MODFLAG_GLOBAL_SYMBOLS = 1, AST_MODFLAG_LOAD_ORDER = 2 };
enum { AST_MODPRI_DEFAULT = 128 };
struct ast_module;
struct ast_module_info {
  const char *name, *description, *key;
  unsigned int flags;
  int (*load)(void);
  int (*unload)(void);
  int (*reload)(void);
  int load_pri;
  int support_level;
};
/* the harness finds the module through this symbol (the real macro registers with the loader instead) */
#define AST_MODULE_INFO(keystr, flags_to_set, desc, ...) \
  struct ast_module_info fake_ast_module_info = {.name = AST_MODULE, .description = desc, .key = keystr, .flags = flags_to_set, __VA_ARGS__}
#endif
