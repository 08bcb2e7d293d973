// htslib-API shim (oracle build only): see ../hts_shim.cpp
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef struct faidx_t faidx_t;
faidx_t *fai_load(const char *fn);
void fai_destroy(faidx_t *fai);
/* "chr:beg-end" (1-based inclusive, commas allowed); returns malloc'ed NUL-terminated sequence */
char *fai_fetch(const faidx_t *fai, const char *reg, int *len);
char *faidx_fetch_seq(const faidx_t *fai, const char *c_name, int p_beg_i, int p_end_i, int *len);
#ifdef __cplusplus
}
#endif
