// htslib-API shim (oracle build only): see ../hts_shim.cpp
#pragma once
#include <stddef.h>
typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
