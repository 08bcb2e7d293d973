/* rabbitvar_b200.h — C ABI of the B200-native pileup-and-score path of RabbitVar.
 *
 * Drop-in boundary.  The reference (LeiHaoa/RabbitVar) has no plugin/FFI layer; the seam this
 * library replaces is the per-region C++ call
 *     Scope<AlignedVarsData>* one_region_run(Region, Configuration*, dataPool*, vector<bamReader>, set<string>*)
 * (reference src/modes/simpleMode.cpp:18-64; somatic twin src/modes/somaticMode.cpp:83-127) and the
 * three stage calls inside it:
 *     CigarParser::process        include/parseCigar.h:47          -> rv_pileup()
 *     VariationRealigner::process include/VariationRealigner.h:132 -> host side, fed by rv_fetch_* / rv_apply_patch()
 *     ToVarsBuilder::process      include/ToVarsBuilder.h:66       -> rv_score()
 * One context = one host thread = one GPU stream set (mirrors "one thread = one dataPool + BAM
 * handles", include/modes/simpleMode.h:28-31).  All functions return 0 on success, <0 on error
 * (message via rv_last_error); no C++ exceptions and no torch types cross this boundary.
 * Inputs are caller-owned, outputs library-owned (valid until the next call that produces the same
 * output on the same context, or rv_destroy).  There is NO CPU fallback: every entry point that
 * computes fails with RV_ERR_CUDA when no sm_100-class device is present.
 */
#ifndef RABBITVAR_B200_H
#define RABBITVAR_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RV_ABI_VERSION 4
#define RV_OK 0
#define RV_ERR_ARG (-1)
#define RV_ERR_CUDA (-2)
#define RV_ERR_NOMEM (-3)
#define RV_ERR_OVERFLOW (-4) /* event / variant / patch buffer exceeded: raise the limits in rv_limits */
#define RV_ERR_STATE (-5)

typedef struct rv_ctx rv_ctx;

/* Subset of the reference Configuration (include/Configuration.h:98-401) the path reads. */
typedef struct rv_params {
  double goodq;            /* -q  phred_score, default 22.5 */
  double freq;             /* -f  allele frequency threshold, default 0.01 */
  double lofreq;           /* -V  default 0.05 */
  double qratio;           /* -o  default 1.5 */
  double mapq;             /* -O  default 0 */
  double bias;             /* strand-bias fraction, Configuration.h:193, 0.05 */
  int32_t vext;            /* -X  default 2 */
  int32_t mismatch;        /* -m  default 8 */
  int32_t minr;            /* -r  default 2 */
  int32_t min_bias_reads;  /* -B  default 2 */
  int32_t read_pos_filter; /* -P  default 5 */
  int32_t minmatch;        /* -M  default 0 */
  int32_t trim_bases_after;/* -T  default 0 */
  int32_t indelsize;       /* -I  default 50 */
  int32_t mapping_quality; /* -Q  default 0 */
  int32_t samfilter;       /* -F  default 0x504 */
  uint8_t local_realign;   /* -k  default 1 */
  uint8_t move3;           /* -3 */
  uint8_t uniq_u;          /* -u */
  uint8_t uniq_un;         /* --UN */
  uint8_t dedup;           /* -t */
  uint8_t pileup;          /* -p */
  uint8_t fisher;          /* --fisher */
  uint8_t has_bam2;        /* somatic: second BAM present (ToVarsBuilder.cpp:170-175,213-233) */
  uint8_t candidates_only; /* scoring emits a position only if one of its variants passes the numeric part of
                              Variant::isGoodVar (freq, hicnt, pmean, qual, qratio: include/Variant.h:205-231) —
                              exact for simple-mode output, which prints nothing else (simpleMode.cpp:176-190).
                              2 = the same cut, but only one bare (region, pos) record per passing position: the
                              candidate list of the paired mode (see rv_score_positions) */
  uint8_t pad_[7];
} rv_params;

void rv_default_params(rv_params* p);

/* Capacity limits of one context (device allocations are sized from these once, at rv_create). */
typedef struct rv_limits {
  int64_t max_reads;        /* reads resident per batch */
  int64_t max_read_bytes;   /* bytes of packed cigar+seq+qual payload per batch */
  int64_t max_positions;    /* sum over regions of (region length + 2*halo) per batch */
  int32_t max_regions;      /* regions per batch */
  int32_t halo;             /* table positions kept either side of a region (soft-clip re-extension,
                               deletions and look-ahead coverage reach outside it: parseCigar.cpp:1170,
                               :1237,:243) */
  int64_t max_events;       /* sparse (complex-key / soft-clip) events per batch */
  int64_t max_variants;     /* scored variant records per batch */
  int64_t max_patch;        /* patch entries per batch */
  int64_t max_ref_bases;    /* reference bases resident */
  int64_t max_sparse_obs;   /* observations of the walked reads that do not travel as gather descriptors (soft-clip
                               re-extension, stretches that are not plain, every base under -T): 0 = 8 x max_reads */
} rv_limits;

void rv_default_limits(rv_limits* l);

/* Fixed 32-byte per-read header.  The variable part of each read lives in one byte pool, laid out
 * exactly as in a BAM record after the read name: u32 cigar[n_cigar] | 4-bit packed seq
 * ((l_seq+1)/2 bytes) | u8 qual[l_seq]; every read starts on a 16-byte boundary. */
typedef struct rv_read {
  int32_t pos;        /* 1-based alignment start (bam pos + 1) */
  int32_t mpos;       /* 1-based mate start */
  uint32_t data_off16;/* offset of the variable part in the pool, in units of 16 bytes */
  int32_t l_seq;
  uint16_t flag;
  uint16_t n_cigar;
  int16_t nm;         /* NM tag value, -1 when the tag is absent */
  uint8_t mapq;
  uint8_t mate_same_tid; /* tid == mtid */
  int32_t end_pos;    /* htslib bam_endpos(): 0-based exclusive end == 1-based inclusive end */
  int32_t mtid;       /* mate's contig id (bam mtid, -1 = none): part of the -t duplicate key */
} rv_read;

typedef struct rv_read_batch {
  int64_t n_reads;
  const rv_read* reads;     /* coordinate-sorted, BAM order */
  const uint8_t* pool;      /* variable parts */
  int64_t pool_bytes;
} rv_read_batch;

/* One region (tile) = the unit the reference processes independently (one_region_run).  Reads of
 * the region are reads[read_lo, read_hi) of the batch that also satisfy the htslib overlap test
 * pos0 < end && end_pos > start-1 (checked on the device). */
typedef struct rv_region {
  int32_t start, end;      /* 1-based inclusive */
  int32_t ref_lo, ref_hi;  /* loaded reference window (recordPreprocessor.cpp:41-55), 1-based inclusive */
  int64_t read_lo, read_hi;
  int32_t chr_len;         /* conf->chrLengths[chr] */
  int32_t max_read_len_in; /* somatic: the normal pass starts from the tumor's value (somaticMode.cpp:109) */
} rv_region;

/* Dense position-major count table: per position 4 alleles (A,C,G,T) x 8 u32 + 1 u32 coverage
 * = 132 bytes.  Row layout of one allele: */
enum { RV_F_FWD = 0, RV_F_REV = 1, RV_F_SUM_TP = 2, RV_F_SUM_Q = 3, RV_F_SUM_MAPQ = 4, RV_F_SUM_NM = 5,
       RV_F_HI = 6, RV_F_STD = 7 /* bits 0-15 first tp, 16-23 first q, 24 pstd, 25 qstd */ };
/* PRECONDITION of pstd / qstd: no base quality of 0.  The reference sets the flags when two observations that follow
 * each other in BAM order differ and the earlier one is non-zero (parseCigar.cpp:902-914); with tp >= 1 and q >= 1 that
 * is "not all observations are equal", which is what the tables hold (order-free).  A Phred-0 base that precedes the
 * others in BAM order would leave the reference's flag unset where this library sets it. */
#define RV_ROW_U32 8
#define RV_POS_U32 (4 * RV_ROW_U32)

/* Capacity of an allele-key string (include/Variant.h:24-33) in an event / in a patch entry.  A key that does not
 * fit is never scored silently: the event carries RV_EVF_KEY_TRUNC, the read counts in rv_pileup_stats.n_unsupported
 * and the host stage refuses the batch (rvh_*: RV_ERR_OVERFLOW). */
#define RV_EVENT_KEY_MAX 112
#define RV_PATCH_KEY_MAX 240

/* Sparse event: one observation that does not fit a dense single-base row (multi-base / indel /
 * complex keys, insertion table, soft-clip accumulators).  160 bytes. */
enum { RV_EV_NI = 0,    /* nonInsertionVariants[pos][key]   (parseCigar.cpp:884-937, :1036-1082) */
       RV_EV_IN = 1,    /* insertionVariants[pos][key]      (parseCigar.cpp:1427-1463) */
       RV_EV_SC5 = 2,   /* softClips5End[pos]               (parseCigar.cpp:1197-1217) */
       RV_EV_SC3 = 3,   /* softClips3End[pos]               (parseCigar.cpp:1268-1295) */
       RV_EV_TTREF = 4  /* insertion right after S/H: extra ref observation (parseCigar.cpp:1497-1515) */ };
enum { RV_EVF_MNP = 1, RV_EVF_PDEL = 2, RV_EVF_PINS = 4, RV_EVF_KEY_TRUNC = 8 };
typedef struct rv_event {
  int32_t region;
  int32_t pos;
  uint32_t read_idx;   /* index in the batch == BAM order */
  uint16_t seq_no;     /* order of the event inside its read */
  uint8_t kind;
  uint8_t flags;
  int32_t tp;
  int32_t nm;          /* numberOfMismatches - nmoff */
  int32_t qsum;        /* quality = qsum / (double) qcnt */
  int32_t qcnt;
  uint8_t dir;         /* 1 = reverse strand */
  uint8_t mapq;
  uint8_t keylen;
  uint8_t pad;
  int32_t aux0;        /* soft clips: remaining clip length m ; TTREF: pstd|qstd<<1 of the insertion */
  int32_t aux1;        /* soft clips: number of high-quality bases kept (n_hi) */
  int32_t aux2;        /* soft clips: read offset of the base nearest the junction */
  char key[RV_EVENT_KEY_MAX];
} rv_event;

/* One accumulator in the reference's own field set (include/Variation.h:12-82) for the host-side
 * realigner hand-off and for positions patched before scoring. */
typedef struct rv_variation {
  int32_t cnt, fwd, rev;
  int32_t lo, hi;
  int32_t extracnt;
  double sum_tp, sum_q, sum_mapq, sum_nm;
  uint8_t pstd, qstd, pad[6];
} rv_variation;

/* A key (allele description string) with its accumulator at a position, for rv_apply_patch. */
typedef struct rv_patch_entry {
  int32_t region;
  int32_t pos;
  uint8_t table;   /* 0 = nonInsertionVariants, 1 = insertionVariants, 2 = tombstone: the dense
                      single-base key `key` no longer exists at this position */
  uint8_t keylen;
  uint8_t pad[2];
  char key[RV_PATCH_KEY_MAX];
  rv_variation v;
} rv_patch_entry;

/* Scored variant record, numeric part (reference include/Variant.h:34-67 minus strings that the host
 * assembles from `key` and the reference window: alleles, genotype, flanks). 152 bytes. */
typedef struct rv_variant {
  int32_t region;
  int32_t pos;
  int32_t cnt, fwd, rev;       /* positionCoverage, varsCountOnForward/Reverse */
  int32_t tcov;                /* totalPosCoverage */
  int32_t hicnt, hicov;
  int32_t ref_fwd, ref_rev;    /* refForwardCoverage / refReverseCoverage */
  int32_t shift3, msint;
  double freq, pmean, qual, mapq, qratio, hifreq, extrafreq, nm, msi;
  double pvalue, oddratio;     /* --fisher: two-sided p and max(ad/bc, bc/ad) (simpleMode.cpp:96-108) */
  uint8_t bias_ref, bias_var;  /* strandBiasFlag "r;v" */
  uint8_t pstd, qstd;
  uint8_t is_ref;              /* this record is the position's reference allele */
  uint8_t key_kind;            /* 0 = dense single base (key_id = allele 0..3), 1 = patch entry (key_id = index) */
  uint8_t rank;                /* order among the position's variants after the CMP_VARI sort */
  uint8_t pad;
  int32_t key_id;
} rv_variant;

/* ---- lifecycle ------------------------------------------------------------------------------- */
int rv_abi_version(void);
int rv_device_count(void);
/* Initialises the CUDA runtime on `device` and loads the library's kernels (what the first rv_create would otherwise
 * pay): callers run it beside their own start-up work. */
int rv_warmup(int device);
int rv_create(rv_ctx** out, int device, const rv_params* params, const rv_limits* limits);
void rv_destroy(rv_ctx* ctx);
const char* rv_last_error(const rv_ctx* ctx);
int rv_sync(rv_ctx* ctx);
/* Lazy mode (off by default): rv_pileup / rv_score / rv_score_positions only enqueue their kernels — no stream
 * synchronisation, no read-back of the statistics.  The host's view is settled by rv_sync or by the first call that
 * reads a result (rv_get_pileup_stats, rv_fetch_*, rv_variant_count, rv_last_*_ms), which then also returns the
 * RV_ERR_OVERFLOW the enqueuing call could not report.  For callers that run several batches' kernels back to back. */
int rv_set_lazy(rv_ctx* ctx, int on);
/* Table halo (limits.halo) of the context, and replacement of its parameter block between batches. */
int32_t rv_ctx_halo(const rv_ctx* ctx);
int rv_set_params(rv_ctx* ctx, const rv_params* params);

/* ---- inputs ---------------------------------------------------------------------------------- */
/* Reference bases [ref_start, ref_start+n) of the contig being processed, upper-case ASCII. */
int rv_set_reference(rv_ctx* ctx, int32_t ref_start, int64_t n, const char* bases);
/* Stage a read batch (host pointers; copied H2D asynchronously on the context stream). */
int rv_push_reads(rv_ctx* ctx, const rv_read_batch* batch);
/* Stage only reads [read_lo, read_hi) of the batch (and their slice of the pool).  Regions set afterwards keep
 * batch-global read indices and must lie inside the staged range; events report batch-global read indices.
 * Lets several contexts work through one large host batch chunk by chunk. */
int rv_push_reads_range(rv_ctx* ctx, const rv_read_batch* batch, int64_t read_lo, int64_t read_hi);
/* Several disjoint ranges at once (paired mode: the tumor tiles' reads and the normal tiles' reads of one chunk);
 * every region must lie inside one of them. */
int rv_push_reads_ranges(rv_ctx* ctx, const rv_read_batch* batch, int32_t n_ranges, const int64_t* read_lo,
                         const int64_t* read_hi);
/* Same, from buffers already resident on the device (all pointers are device pointers; 16-byte aligned pool whose
 * allocation extends at least 32 readable bytes beyond pool_bytes: the gather kernel's look-ahead loads stop there). */
int rv_push_reads_device(rv_ctx* ctx, const rv_read_batch* batch);
/* Regions of this batch. */
int rv_set_regions(rv_ctx* ctx, const rv_region* regions, int32_t n_regions);

/* ---- compute --------------------------------------------------------------------------------- */
/* Read filter + CIGAR rewrite + CIGAR walk + pileup into the dense tables and the event list
 * (CigarParser::process for every region of the batch). */
int rv_pileup(rv_ctx* ctx);
/* Per-position scoring + compaction of candidate variants (ToVarsBuilder::process). */
int rv_score(rv_ctx* ctx);

/* The same scoring for an explicit list of (region, position) pairs only, without the frequency / candidate cuts
 * of the per-sample passes deciding which positions to visit (somatic mode joins the tumor and normal records of
 * the positions where either sample has a candidate: somaticMode.cpp:311-352). */
int rv_score_positions(rv_ctx* ctx, const int32_t* region, const int32_t* pos, int64_t n);

/* ---- outputs --------------------------------------------------------------------------------- */
typedef struct rv_pileup_stats {
  int64_t n_items;          /* (region, read) pairs examined */
  int64_t n_reads_kept;     /* pairs that passed every read filter */
  int64_t n_aligned_bases;  /* M/=/X bases of kept reads (the throughput unit) */
  int64_t n_events;
  int64_t n_overflow;       /* events that did not fit limits.max_events (rv_pileup then returns RV_ERR_OVERFLOW) */
  int64_t n_unsupported;    /* reads that hit a corner the device path refuses (counted, not guessed) */
  int64_t n_walk_items;     /* work items that took the exact CIGAR walk (rv_walk_kernel) ... */
  int64_t n_walk_full;      /* ... of which the whole read was walked (the rest: soft clips of a plain read only) */
  int64_t n_clipped;        /* soft-clip re-extension / deletion / look-ahead coverage observations that fell outside
                               [start - halo, end + halo] and were dropped (raise limits.halo for long reads); the
                               matched bases themselves are only ever counted inside the region (parseCigar.cpp:884) */
  int64_t n_score_unsupported; /* positions where createInsertion would edit the neighbouring position's reference
                               allele (ToVarsBuilder.cpp:405-415): counted by the last scoring call */
  int64_t n_sparse_obs;     /* observations of walked reads added with global atomics (rv_apply_kernel) ... */
  int64_t n_walk_segments;  /* ... and plain matched stretches of walked reads handed to the gather kernel instead */
} rv_pileup_stats;
int rv_get_pileup_stats(rv_ctx* ctx, rv_pileup_stats* out);
/* maxReadLength per region after the pileup (parseCigar.cpp:598-601). */
int rv_fetch_max_read_len(rv_ctx* ctx, const int32_t** out, int32_t* n);
/* Dense tables of region r: table has (end-start+1+2*halo) positions, first position = start-halo.
 * counts: RV_POS_U32 u32 per position; cov: 1 u32 per position. Host copies (pinned, library-owned). */
int rv_fetch_tables(rv_ctx* ctx, int32_t region, const uint32_t** counts, const uint32_t** cov,
                    int32_t* first_pos, int32_t* n_pos);
/* Gather of selected dense rows: for each (region[i], pos[i]) 33 u32 = RV_POS_U32 counts + coverage.
 * Positions outside the region's table come back as zeros. Host pointers in, library-owned pinned buffer out. */
int rv_fetch_rows(rv_ctx* ctx, const int32_t* region, const int32_t* pos, int64_t n, const uint32_t** rows);
int rv_fetch_events(rv_ctx* ctx, const rv_event** events, int64_t* n_events);
/* Replace/insert accumulators before scoring (realigner write-back): dense single-base keys update the
 * dense row, everything else goes to the per-position patch list; cov_pos/cov_val overwrite coverage. */
int rv_apply_patch(rv_ctx* ctx, const rv_patch_entry* entries, int64_t n_entries, const int32_t* cov_region,
                   const int32_t* cov_pos, const int32_t* cov_val, int64_t n_cov);
int rv_fetch_variants(rv_ctx* ctx, const rv_variant** variants, int64_t* n_variants);
/* Per region: sum of the coverage and number of covered positions over [start, end) (add_depth_by_region,
 * somaticMode.cpp:69-81: the numbers behind the paired mode's <out>.info file).  Host arrays of n_regions. */
int rv_cov_summary(rv_ctx* ctx, int64_t* sum, int64_t* covered);
/* Number of records the last rv_score produced (no copy). */
int64_t rv_variant_count(const rv_ctx* ctx);

/* Fisher exact test for a batch of 2x2 tables on the device (call sites simpleMode.cpp:98,
 * somaticMode.cpp:132; algorithm of htslib kfunc.c kt_fisher_exact).  tables: n x 4 ints
 * (n11,n12,n21,n22); out: n x 3 doubles (left, right, two-sided). Host pointers. */
int rv_fisher_exact(rv_ctx* ctx, const int32_t* tables, int64_t n, double* out);

/* Device timing of the last rv_pileup / rv_score in milliseconds (CUDA events on the context stream). */
int rv_last_kernel_ms(rv_ctx* ctx, float* pileup_ms, float* score_ms);
/* The same for the kernels of the last rv_pileup: rv_pileup_kernel (filters, CIGAR rewrite, plain-run proof),
 * rv_tile_index_kernel + rv_gather_kernel (position-major accumulation), rv_walk_kernel (exact CIGAR walks). */
int rv_last_pileup_split_ms(rv_ctx* ctx, float* classify_ms, float* gather_ms, float* walk_ms);
/* All four: out[0] rv_pileup_kernel, out[1] rv_tile_index_kernel + rv_gather4_kernel, out[2] rv_walk_kernel,
 * out[3] rv_apply_kernel (launch order: classify, walk, gather, apply). */
int rv_last_pileup_stage_ms(rv_ctx* ctx, float out[4]);
/* Brackets any sequence of calls with CUDA events recorded on the context's (launching) stream;
 * rv_timer_stop synchronises and returns the elapsed device time in milliseconds. */
int rv_timer_start(rv_ctx* ctx);
int rv_timer_stop(rv_ctx* ctx, float* ms);
/* Number of kernels launched by this context so far. */
int64_t rv_launch_count(const rv_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
