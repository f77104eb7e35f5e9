/* Placeholder for <aubio/aubio.h> (test infrastructure): the unchanged module shell src/app_tiresias.c:24 still
 * includes it but calls nothing from it; the replacement fp_handler.c does not use libaubio at all. */
