/* fake <asterisk/manager.h> (test infrastructure): included by src/app_tiresias.c, nothing of it is used */
