/* Minimal declarations of the libjansson.so.4 functions the fake ast_json layer calls (test infrastructure; the
 * image has the library but not its header).  Prototypes as documented in the jansson 2.x API reference. */
#ifndef FAKE_JANSSON_H_
#define FAKE_JANSSON_H_
#include <stdarg.h>
#include <stddef.h>
typedef enum { JSON_OBJECT, JSON_ARRAY, JSON_STRING, JSON_INTEGER, JSON_REAL, JSON_TRUE, JSON_FALSE, JSON_NULL } json_type;
typedef struct json_t {
  json_type type;
  volatile size_t refcount;
} json_t;
typedef long long json_int_t;
typedef struct json_error_t {
  int line, column, position;
  char source[80], text[160];
} json_error_t;
#define JSON_COMPACT 0x20
#define JSON_ENCODE_ANY 0x200
#define JSON_DECODE_ANY 0x4
#define JSON_INDENT(n) ((n) & 0x1F)
json_t *json_object(void);
json_t *json_array(void);
json_t *json_string(const char *value);
json_t *json_integer(json_int_t value);
json_t *json_real(double value);
json_t *json_null(void);
void json_delete(json_t *json);
json_t *json_object_get(const json_t *object, const char *key);
int json_object_set_new(json_t *object, const char *key, json_t *value);
void *json_object_iter(json_t *object);
void *json_object_iter_next(json_t *object, void *iter);
const char *json_object_iter_key(void *iter);
json_t *json_object_iter_value(void *iter);
size_t json_array_size(const json_t *array);
json_t *json_array_get(const json_t *array, size_t index);
int json_array_append_new(json_t *array, json_t *value);
int json_array_remove(json_t *array, size_t index);
const char *json_string_value(const json_t *string);
json_int_t json_integer_value(const json_t *integer);
double json_real_value(const json_t *real);
json_t *json_vpack_ex(json_error_t *error, size_t flags, const char *fmt, va_list ap);
json_t *json_deep_copy(const json_t *value);
json_t *json_loads(const char *input, size_t flags, json_error_t *error);
char *json_dumps(const json_t *json, size_t flags);
#endif
