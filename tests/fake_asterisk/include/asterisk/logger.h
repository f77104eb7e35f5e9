/* fake <asterisk/logger.h> (test infrastructure) */
#ifndef FAKE_AST_LOGGER_H_
#define FAKE_AST_LOGGER_H_
#define LOG_DEBUG 0
#define LOG_NOTICE 2
#define LOG_WARNING 3
#define LOG_ERROR 4
#define LOG_VERBOSE 5
#define AST_LOG_DEBUG LOG_DEBUG
#define AST_LOG_NOTICE LOG_NOTICE
#define AST_LOG_WARNING LOG_WARNING
#define AST_LOG_ERROR LOG_ERROR
#define AST_LOG_VERBOSE LOG_VERBOSE
void fake_ast_log(int level, const char *file, int line, const char *func, const char *fmt, ...) __attribute__((format(printf, 5, 6)));
#define ast_log(level, ...) fake_ast_log((level), __FILE__, __LINE__, __func__, __VA_ARGS__)
#endif
