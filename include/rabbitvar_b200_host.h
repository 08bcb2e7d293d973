/* rabbitvar_b200_host.h — C ABI of the host side of the path: BAM/FASTA decode into the staging
 * buffers rv_push_reads consumes, and the batch-level replacement of one_region_run (reference
 * src/modes/simpleMode.cpp:18-64) that returns TSV text in the reference's output format
 * (print_output_variant_simple, simpleMode.cpp:66-142).  Lives in the same shared library as
 * rabbitvar_b200.h. */
#ifndef RABBITVAR_B200_HOST_H
#define RABBITVAR_B200_HOST_H
#include "rabbitvar_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rvh_batch rvh_batch;

/* Decode every read overlapping chr:start-end (1-based inclusive) of an indexed BAM, in file order
 * (replaces sam_itr_querys/sam_itr_next, recordPreprocessor.cpp:10-32,101-103). NULL on failure. */
rvh_batch* rvh_load_bam(const char* bam_path, const char* chr, int32_t start, int32_t end, int32_t* chr_len_out);
/* Concatenate b onto a (reads of a second sample); returns the read-index offset of b's reads inside a. */
int64_t rvh_batch_append(rvh_batch* a, const rvh_batch* b);
int64_t rvh_batch_n_reads(const rvh_batch* b);
const rv_read* rvh_batch_reads(const rvh_batch* b);
const uint8_t* rvh_batch_pool(const rvh_batch* b);
int64_t rvh_batch_pool_bytes(const rvh_batch* b);
int32_t rvh_batch_max_ref_span(const rvh_batch* b);
/* Page-lock the batch's host buffers (cudaHostRegister) so rv_push_reads copies from pinned memory. */
int rvh_batch_pin(rvh_batch* b);
void rvh_batch_free(rvh_batch* b);

/* Region descriptors (read ranges + reference window of RecordPreprocessor::makeReference,
 * recordPreprocessor.cpp:41-78) for n regions of one contig. read_offset/n_reads_sample restrict the
 * search to one sample's slice of a concatenated batch (0, -1 = whole batch). */
int rvh_make_regions(const rvh_batch* b, const int32_t* starts, const int32_t* ends, int32_t n, int32_t chr_len,
                     int32_t ref_extension, int64_t read_offset, int64_t n_reads_sample, rv_region* out);

/* Upper-cased reference bases chr:lo-hi (1-based inclusive) from an indexed FASTA into out (hi-lo+1 bytes).
 * Returns the number of bases written, <0 on failure. */
int64_t rvh_fetch_ref(const char* fasta_path, const char* chr, int32_t lo, int32_t hi, char* out);

/* The whole per-batch path on host buffers: H2D, pileup, realign hand-off, scoring, D2H, TSV formatting
 * (simple mode).  tsv_out is library-owned (valid until the next call on this ctx or rv_destroy). */
typedef struct rvh_timing {
  double push_ms, pileup_ms, fetch_ms, host_ms, patch_ms, score_ms, assemble_ms;
  float pileup_kernel_ms, score_kernel_ms;
  int64_t n_items, n_reads_kept, n_aligned_bases, n_events, n_unsupported, n_variants, n_lines, h2d_bytes, d2h_bytes;
} rvh_timing;
int rvh_call_regions(rv_ctx* ctx, const rv_params* params, const rvh_batch* batch, const rv_region* regions,
                     int32_t n_regions, const char* ref_bases, int32_t ref_lo, int64_t ref_n,
                     int push_flags /* bit0: upload the reference, bit1: upload the reads (else reuse resident) */,
                     const char* sample, const char* chr, const char** tsv_out, int64_t* tsv_len, rvh_timing* timing);
/* The host hand-off alone, for a batch that is already resident and piled up on ctx (after rv_pileup):
 * events/tables D2H, BAM-order reduce, host realigner, patch write-back (rv_apply_patch). */
int rvh_install_patch(rv_ctx* ctx, const rv_params* params, const rvh_batch* batch, const rv_region* regions,
                      int32_t n_regions, const char* ref_bases, int32_t ref_lo, int64_t ref_n);

/* Pipelined form of rvh_call_regions for large batches: the region list is cut into chunks of chunk_regions
 * regions; n_workers host threads, each with its own context (stream + device buffers sized for one chunk) on
 * `device`, pull chunks from a queue, so the H2D copy of one chunk overlaps the kernels of another and the host
 * stages (event reduce, realign hand-off, TSV assembly) of a third.  This is the region loop of the reference
 * (simpleMode.cpp:320-347: one region per OpenMP thread) with chunks in place of regions.  Output = the
 * concatenation of the chunks' TSV in region order, identical to a single rvh_call_regions over all regions. */
typedef struct rvh_pipeline rvh_pipeline;
rvh_pipeline* rvh_pipeline_create(int device, int n_workers);
void rvh_pipeline_destroy(rvh_pipeline* p);
int rvh_pipeline_run(rvh_pipeline* p, const rv_params* params, const rvh_batch* batch, const rv_region* regions,
                     int32_t n_regions, int32_t chunk_regions, const char* ref_bases, int32_t ref_lo, int64_t ref_n,
                     const char* sample, const char* chr, const char** tsv_out, int64_t* tsv_len, rvh_timing* timing);
/* Paired tumor | normal form (one_region_run_somt + SomaticMode::output, somaticMode.cpp:83-127, :311-352):
 * `regions` = the n tumor tiles followed by the same n tiles of the normal sample (n_regions = 2n), both referring to
 * reads of one concatenated batch (rvh_batch_append); a chunk takes chunk_tiles tiles of both samples.  Output = the
 * 55/63-column lines of the reference's somatic mode. */
int rvh_pipeline_run_paired(rvh_pipeline* p, const rv_params* params, const rvh_batch* batch, const rv_region* regions,
                            int32_t n_regions, int32_t chunk_tiles, const char* ref_bases, int32_t ref_lo, int64_t ref_n,
                            const char* sample, const char* chr, const char** tsv_out, int64_t* tsv_len, rvh_timing* timing);
/* Kernels launched by the pipeline's contexts so far. */
int64_t rvh_pipeline_launch_count(const rvh_pipeline* p);
const char* rvh_last_error(void);

/* The loader's block decoder and checksum (csrc/io/fast_inflate.hpp), exposed so that they can be checked against
 * zlib from outside: one raw-DEFLATE stream (a BGZF block's payload) in, bytes out.  Returns the number of bytes
 * written, -1 when the stream is malformed or does not fit out_cap. */
int64_t rvh_inflate_block(const uint8_t* in, int64_t in_len, uint8_t* out, int64_t out_cap);
uint32_t rvh_crc32(const uint8_t* buf, int64_t n);

/* Files in, TSV text out: the region loop of the reference's CLI (Launcher.cpp / simpleMode.cpp:210-387 /
 * somaticMode.cpp:860-930) for a list of regions of one or more contigs — parallel BGZF/BAM decode threads feeding GPU
 * worker contexts (csrc/host/file_pipeline.hpp).  bam2 = NULL or "" selects the single-sample mode.  Regions are
 * (chr[i], start[i], end[i], gene[i]), 1-based inclusive.  tsv_out is malloc'ed by the library: free it with
 * rvh_free.  Returns 0, 2 when some regions failed (see rvh_last_error), 3 without a CUDA device. */
int rvh_run_files(const rv_params* params, const char* fasta, const char* bam, const char* bam2, const char* sample,
                  int32_t n_regions, const char* const* chr, const int32_t* start, const int32_t* end,
                  const char* const* gene, int32_t decode_threads, int32_t gpus, int32_t first_device,
                  char** tsv_out, int64_t* tsv_len, double* cov_info /* [4]: sum T, sites T, sum N, sites N; may be NULL */);
void rvh_free(void* p);


#ifdef __cplusplus
}
#endif
#endif
