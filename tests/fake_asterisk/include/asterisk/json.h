/* fake <asterisk/json.h> (test infrastructure): the ast_json_* wrapper, implemented over libjansson.so.4 */
#ifndef FAKE_AST_JSON_H_
#define FAKE_AST_JSON_H_
#include <stdint.h>
struct ast_json;
struct ast_json_iter;
struct ast_json_error;
enum ast_json_type { AST_JSON_OBJECT, AST_JSON_ARRAY, AST_JSON_STRING, AST_JSON_INTEGER, AST_JSON_REAL, AST_JSON_TRUE, AST_JSON_FALSE, AST_JSON_NULL };
enum ast_json_encoding_format { AST_JSON_COMPACT, AST_JSON_PRETTY };
struct ast_json *ast_json_ref(struct ast_json *value);
void ast_json_unref(struct ast_json *value);
enum ast_json_type ast_json_typeof(const struct ast_json *value);
struct ast_json *ast_json_null(void);
struct ast_json *ast_json_string_create(const char *value);
const char *ast_json_string_get(const struct ast_json *string);
struct ast_json *ast_json_integer_create(intmax_t value);
intmax_t ast_json_integer_get(const struct ast_json *integer);
struct ast_json *ast_json_real_create(double value);
double ast_json_real_get(const struct ast_json *real);
struct ast_json *ast_json_array_create(void);
size_t ast_json_array_size(const struct ast_json *array);
struct ast_json *ast_json_array_get(const struct ast_json *array, size_t index);
int ast_json_array_append(struct ast_json *array, struct ast_json *value); /* steals the reference */
int ast_json_array_remove(struct ast_json *array, size_t index);
struct ast_json *ast_json_object_create(void);
struct ast_json *ast_json_object_get(struct ast_json *object, const char *key);
int ast_json_object_set(struct ast_json *object, const char *key, struct ast_json *value); /* steals the reference */
struct ast_json_iter *ast_json_object_iter(struct ast_json *object);
struct ast_json_iter *ast_json_object_iter_next(struct ast_json *object, struct ast_json_iter *iter);
const char *ast_json_object_iter_key(struct ast_json_iter *iter);
struct ast_json *ast_json_object_iter_value(struct ast_json_iter *iter);
struct ast_json *ast_json_pack(char const *format, ...);
struct ast_json *ast_json_deep_copy(const struct ast_json *value);
struct ast_json *ast_json_load_string(const char *input, struct ast_json_error *error);
char *ast_json_dump_string_format(struct ast_json *root, enum ast_json_encoding_format format);
#define ast_json_dump_string(root) ast_json_dump_string_format((root), AST_JSON_COMPACT)
#endif
