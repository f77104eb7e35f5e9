/* Minimal declarations of libuuid.so.1 (test infrastructure; the image has the library but not its header). */
#ifndef FAKE_UUID_H_
#define FAKE_UUID_H_
typedef unsigned char uuid_t[16];
void uuid_generate(uuid_t out);
void uuid_unparse_lower(const uuid_t uu, char *out);
int uuid_parse(const char *in, uuid_t uu);
#endif
