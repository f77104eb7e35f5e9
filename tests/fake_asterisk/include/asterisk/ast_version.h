/* fake <asterisk/ast_version.h> (test infrastructure): included by src/app_tiresias.c, nothing of it is used */
