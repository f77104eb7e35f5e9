// placeholder, replaced by the match engine
#include "tir_internal.h"
struct TirDb { int dummy; };
void tir_db_destroy(TirDb *db) { delete db; }
