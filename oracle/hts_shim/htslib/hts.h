// htslib-API shim (oracle build only): see ../hts_shim.cpp
// Declares only the subset of the public htslib API the reference links against.
#pragma once
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct htsFile htsFile;
typedef struct hts_idx_t hts_idx_t;
typedef struct hts_itr_t hts_itr_t;
extern const char seq_nt16_str[];
void hts_idx_destroy(hts_idx_t *idx);
void hts_itr_destroy(hts_itr_t *iter);
#ifdef __cplusplus
}
#endif
