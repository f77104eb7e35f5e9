/* fake <asterisk/utils.h> (test infrastructure): the helpers live in <asterisk.h> */
#ifndef FAKE_AST_UTILS_H_
#define FAKE_AST_UTILS_H_
#include <asterisk.h>
#endif
