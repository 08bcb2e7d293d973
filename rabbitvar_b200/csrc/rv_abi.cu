// rv_abi.cu — CUDA kernels (sm_100a) and the C ABI of rabbitvar_b200 (include/rabbitvar_b200.h).
//
// Data layout in HBM (per context):
//   reads   : rv_read[max_reads] (32 B fixed headers) + byte pool (cigar | 4-bit seq | qual per read)
//   ref     : 1 B / reference base, flat contig slice
//   counts  : position-major dense tables, [position][allele A,C,G,T][8 x u32] = 128 B / position
//   cov     : u32 coverage / position
//   events  : rv_event[max_events] appended through one global atomic cursor
//   patch   : rv_patch_entry[] + u32 patch_first[position] (0 = none, else index+1) + group sizes
//   variants: rv_variant[max_variants] appended through one global atomic cursor
// Regions of a batch own disjoint slices [tab_off, tab_off + len + 2*halo) of counts/cov.
#include "../../include/rabbitvar_b200.h"
#include "kernels/rv_core.cuh"
#include "kernels/rv_score.cuh"
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

using namespace rvk;
namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------
// device-side views
// ------------------------------------------------------------------------------------------------
struct DevRegion {
  rv_region r;
  int64_t tab_off;    // first table position index of this region
  int32_t first_pos;  // reference position of table index tab_off
  int32_t n_pos;
  int64_t item_base;  // first (region, read) work item of this region
  int64_t tile_base;  // first gather tile (G4_W table positions each, halo included) of this region
  // the region's reads live in one uploaded slice of the host batch: device read index = batch index - read_bias,
  // device pool byte = batch pool byte - pool_bias (rv_push_reads_ranges)
  int64_t read_bias, pool_bias;
};


static const size_t COUNTER_BLOCK_BYTES = 320;
struct DevStats {  // (must fit the first 128 bytes of the counter block)
  unsigned long long n_items, n_kept, n_bases, n_events, n_overflow, n_unsupported, n_variants, n_score_unsupported;
  unsigned long long n_walk_items, n_walk_full;
  unsigned long long n_clipped;  // observations outside [start - halo, end + halo]: dropped, like the gather path clips
  unsigned long long n_sparse;   // SparseObs entries the walk kernel left for rv_apply_kernel
  unsigned long long n_segments; // gather descriptors written by the walk kernel
};
static_assert(sizeof(DevStats) <= 128, "DevStats must fit the first 128 bytes of the counter block");

// Gather descriptor of one plain matched segment (rvk::scan_plain_segment / the plain-run proof of rv_pileup_kernel):
// all the gather kernel needs to find the segment's qualities and to give every base its read position.
// A work item owns up to two: descs[item] and, when bit 15 of its m_len is set, descs2[item] (a read with one indel).
struct GDesc {          // 16 bytes, one LDG.128
  int32_t m_start;      // reference position of the first base
  uint16_t m_len;       // bits 0-13: bases (0 = nothing for the gather path); bit 15: descs2[item] holds a second segment
  uint16_t re0;         // read offset of the first base, soft clips excluded: tp of base k = min(re0 + k + 1, rlen - re0 - k)
  uint32_t qual_off;    // byte offset in the device pool of the first base's quality
  uint16_t rlen;        // matched + inserted bases of the read (parseCigar.cpp:591-595)
  uint8_t mapq;
  uint8_t dir_nm;       // bit 7: reverse strand; bits 0-6: nm (0..127)
};
static const uint32_t GD_LEN_MASK = 0x3fffu, GD_HAS_SECOND = 0x8000u;

// Observation that does not go through the gather kernel (soft-clip re-extension, the anchor base an insertion takes
// back, bases of stretches that are not plain, coverage under deletions): rv_walk_kernel appends them, rv_apply_kernel
// adds them to the tables with global atomics once the gather kernel has stored every row.
struct SparseObs {      // 16 bytes
  uint32_t tab;         // table position index (region slice offset + position - first_pos)
  uint16_t tp;
  uint8_t q, mapq;
  int16_t nm;
  uint8_t flags;        // bits 0-1 allele, bit 2 reverse strand, bits 3-4 kind
  uint8_t pad;
  uint32_t pad2;
};
enum { SO_SINGLE = 0, SO_ADD = 1, SO_SUB = 2, SO_COV = 3 };
static const int SPARSE_CHUNK = 32;  // (tab == 0xffffffff marks an unused slot of a chunk)

struct PileupArgs {
  rv_params P;
  const DevRegion* regions;
  int n_regions;
  const int32_t* itemblk_region;  // region of work item 128 b
  int64_t n_items;
  const rv_read* reads;
  const uint8_t* pool;
  const char* ref;
  const uint32_t* ref4;  // the same slice, 4 bits / base (A=1 C=2 G=4 T=8 other=15), base e in bits 28-4*(e&7) of word e>>3
  int32_t ref_start;
  int64_t ref_n;
  uint32_t* counts;
  uint32_t* cov;
  rv_event* events;
  unsigned long long max_events;
  int32_t* max_rl;
  DevStats* stats;
  GDesc* descs;          // one per work item
  uint16_t* desc_mm;     // per work item: bit b set = the plain run has a mismatch in its bases [16b, 16b+16) (bit 15: and beyond)
  uint4* desc_mml;       // per work item with desc_mm != 0: the run's mismatches, up to 8 entries of 16 bits,
                         // 0x8000 | run offset << 2 | allele of the read base (rv_gather4_kernel)
  GDesc* descs2;         // second segment of a work item (GD_HAS_SECOND), with its own mismatch mask / list
  uint16_t* desc_mm2;
  uint4* desc_mml2;
  SparseObs* sparse;     // observations for rv_apply_kernel
  unsigned long long max_sparse;
  unsigned long long* sparse_count;
  int32_t* reach;        // [0] = max(pos - m_start), [1] = max(m_start + m_len - pos) over the descriptors
  int force_exact;       // debugging: every read takes the exact walk
  // reads that need the exact CIGAR walk are not walked by the classifying kernel (one slow lane would stall
  // its 31 neighbours): their work item goes to this queue and rv_walk_kernel runs them densely packed
  // Whole-read walks fill the queue from the front, soft-clip-only walks (WALK_PLAIN_DONE) from the back, so that the
  // lanes of a warp of rv_walk_kernel run the same kind of walk side by side.
  unsigned long long* walk_queue;   // item | flags | class << WALK_CLASS_SHIFT, in item order (rv_pileup_kernel)
  unsigned long long* walk_queue2;  // the same entries grouped by class (rv_walk_sort_kernel): what rv_walk_kernel reads
  unsigned long long* walk_count;   // [0] entries, [2] the walk kernel's cursor, [8..15] entries per class, [16..23] sort cursors
  unsigned long long walk_cap;
};
// Classes of queued reads by CIGAR shape (the lanes of a warp of rv_walk_kernel should run the same code: the kernel is
// bound by instruction fetch): 0 = one M op (dirty ends, mismatch clusters, N bases), 1 = S M, 2 = M S, 3 = S M S,
// 4 = M D M, 5 = M I M, 6 = one insertion or deletion with soft clips, 7 = everything else.  Walked from 7 down to 0
// (longest first).
static const int WALK_CLASSES = 8;
static const int WALK_CLASS_SHIFT = 57;
static const unsigned long long WALK_ITEM_MASK = (1ull << WALK_CLASS_SHIFT) - 1ull;
static const unsigned long long WALK_PLAIN_DONE = 1ull << 62;  // the matched run already left a descriptor

// Sink of prepare_read (read filters, CIGAR rewrite): statistics and maxReadLength only.
struct DeviceSink {
  const PileupArgs* a;
  const DevRegion* dr;
  int kept_bases, n_kept, n_unsup, n_over;
  bool mute;  // rv_walk_kernel re-runs prepare_read: its statistics were already counted by rv_pileup_kernel
  __device__ __forceinline__ void max_read_len(int tlen) {
    if (!mute && tlen > a->max_rl[dr - a->regions]) atomicMax(a->max_rl + (dr - a->regions), tlen);
  }
  __device__ __forceinline__ void kept(int aligned) { if (!mute) { kept_bases += aligned; n_kept++; } }
  __device__ __forceinline__ void unsupported() { if (!mute) n_unsup++; }
};

// Sink of rv_walk_kernel: nothing is added to the tables here.  Plain segments become gather descriptors, everything
// else a SparseObs for rv_apply_kernel or an event for the host stage.
struct ListSink {
  const PileupArgs* a;
  const DevRegion* dr;
  int64_t item;
  int n_seg;             // descriptors this item has written
  bool first_taken;      // descs[item] already holds the matched run of a plain read (WALK_PLAIN_DONE)
  bool want_segments;
  const rv_read* rd;
  int64_t pool_off;      // device pool byte offset of the read's variable part
  int n_unsup, n_over, n_clip, n_ev;
  int first_pos, n_pos;
  int64_t tab_off;
  __device__ __forceinline__ void bind(const DevRegion* r) {
    dr = r;
    first_pos = r->first_pos;
    n_pos = r->n_pos;
    tab_off = r->tab_off;
  }
  // list slots are reserved SPARSE_CHUNK at a time per lane (one atomic on the shared cursor per chunk, nothing to
  // wait for in between); what a lane has left over when the kernel ends is filled with no-op entries (flush)
  unsigned long long slot_next, slot_end;
  uint32_t ridx1;  // -T only: the read's index in the batch + 1 (BAM order inside a region), 0 otherwise; travels with every entry
  __device__ __forceinline__ void put(int pos, uint32_t flags, int tp, int q, int mapq, int nm) {
    const int i = pos - first_pos;
    if (i < 0 || i >= n_pos) { n_clip++; return; }
    if (slot_next == slot_end) {
      slot_next = atomicAdd(a->sparse_count, (unsigned long long)SPARSE_CHUNK);
      slot_end = slot_next + SPARSE_CHUNK;
    }
    const unsigned long long slot = slot_next++;
    if (slot >= a->max_sparse) { n_over++; return; }
    uint4 v;
    v.x = (uint32_t)(tab_off + i);
    v.y = ((uint32_t)tp & 0xffffu) | (((uint32_t)q & 0xffu) << 16) | (((uint32_t)mapq & 0xffu) << 24);
    v.z = ((uint32_t)nm & 0xffffu) | (flags << 16);
    v.w = ridx1;
    *(uint4*)(a->sparse + slot) = v;
  }
  __device__ __forceinline__ void flush() {
    for (; slot_next < slot_end; ++slot_next)
      if (slot_next < a->max_sparse) *(uint4*)(a->sparse + slot_next) = make_uint4(0xffffffffu, 0u, 0u, 0u);
  }
  __device__ __forceinline__ void single(int pos, int allele, bool dir, int tp, int q, int mapq, int nm) {
    put(pos, (uint32_t)allele | (dir ? 4u : 0u) | (SO_SINGLE << 3), tp, q, mapq, nm);
  }
  __device__ __forceinline__ void adj(int pos, int allele, int sign, bool dir, int tp, int q, int mapq, int nm) {
    put(pos, (uint32_t)allele | (dir ? 4u : 0u) | ((sign > 0 ? SO_ADD : SO_SUB) << 3), tp, q, mapq, nm);
  }
  __device__ __forceinline__ void sub_anchor(int pos, int allele, bool dir, int tp, int q, int mapq, int nm) {
    put(pos, (uint32_t)allele | (dir ? 4u : 0u) | (SO_SUB << 3), tp, q, mapq, nm);  // (-T: applied by the second apply pass)
  }
  __device__ __forceinline__ void cov(int pos) { put(pos, SO_COV << 3, 0, 0, 0, 0); }
  __device__ __forceinline__ void event(const rv_event& e) {
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long slot = 0;
    if (g.thread_rank() == 0) slot = atomicAdd(&a->stats->n_events, (unsigned long long)g.size());
    slot = g.shfl(slot, 0) + g.thread_rank();
    if (slot >= a->max_events) { n_over++; return; }
    a->events[slot] = e;
    n_ev++;
  }
  // scan_plain_segment with the nibble-SIMD proof (8 bases per step against the 4-bit reference)
  __device__ __forceinline__ int scan_segment(const rv_params& P, const ReadView& rv, const RefView& ref, int m_start, int rp, int len,
                                              bool indel_follows, SegDesc* out) {
    if (!want_segments || len <= 0 || len > 8192) return 0;
    const int E0 = m_start - rp - a->ref_start;
    const int64_t w_lo = ref.lo > a->ref_start ? ref.lo : a->ref_start;
    const int64_t w_hi = (int64_t)ref.hi < a->ref_start + a->ref_n - 1 ? (int64_t)ref.hi : a->ref_start + a->ref_n - 1;
    if (E0 < 0 || m_start < w_lo || (int64_t)m_start + len - 1 > w_hi) return 0;
    PlainScan ps;
    const int kind = simd_scan_kind(P, (const uint32_t*)rv.seq4, a->ref4, E0, rp, len, indel_follows, &ps);
    if (kind == 0) return 0;
    out->mm_blocks = ps.mm_blocks;
    out->ml[0] = (uint32_t)ps.ml_lo; out->ml[1] = (uint32_t)(ps.ml_lo >> 32);
    out->ml[2] = (uint32_t)ps.ml_hi; out->ml[3] = (uint32_t)(ps.ml_hi >> 32);
    out->n_mm = ps.ml_n;
    out->p_first = ps.p_first;
    out->p_last = ps.p_last;
    return kind;
  }
  __device__ __forceinline__ bool segment(const SegDesc& sd, bool dir, int mapq, int nm) {
    if (!want_segments) return false;
    const int slot = n_seg + (first_taken ? 1 : 0);
    if (slot > 1 || sd.rlen > 1800 || sd.re > 65535) return false;
    GDesc gd;
    gd.m_start = sd.m_start;
    gd.m_len = (uint16_t)sd.len;
    gd.re0 = (uint16_t)sd.re;
    gd.qual_off = (uint32_t)(pool_off + 4 * (int64_t)rd->n_cigar + ((rd->l_seq + 1) >> 1) + sd.rp);
    gd.rlen = (uint16_t)sd.rlen;
    gd.mapq = (uint8_t)mapq;
    gd.dir_nm = (uint8_t)((dir ? 0x80 : 0) | nm);
    const uint4 ml = make_uint4(sd.ml[0], sd.ml[1], sd.ml[2], sd.ml[3]);
    if (slot == 0) {
      *(uint4*)(a->descs + item) = *(const uint4*)&gd;
      a->desc_mm[item] = (uint16_t)sd.mm_blocks;
      if (sd.mm_blocks) a->desc_mml[item] = ml;
    } else {
      *(uint4*)(a->descs2 + item) = *(const uint4*)&gd;
      a->desc_mm2[item] = (uint16_t)sd.mm_blocks;
      if (sd.mm_blocks) a->desc_mml2[item] = ml;
      ((uint16_t*)(a->descs + item))[2] |= (uint16_t)GD_HAS_SECOND;  // m_len of the first descriptor
    }
    n_seg++;
    const int back = rd->pos - sd.m_start, reach = sd.m_start + sd.len - rd->pos;
    if (back > a->reach[0]) atomicMax(a->reach + 0, back);
    if (reach > a->reach[1]) atomicMax(a->reach + 1, reach);
    return true;
  }
  __device__ __forceinline__ void max_read_len(int) {}   // counted by rv_pileup_kernel
  __device__ __forceinline__ void kept(int) {}
  __device__ __forceinline__ void unsupported() { n_unsup++; }
};

__device__ __forceinline__ int find_region(const DevRegion* regs, int n, int64_t item) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (regs[mid].item_base <= item) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}
// The same for a work item at or after the one whose region r0 is known (lane 0 of the warp searched for the warp's first
// item): consecutive items are in the same region or the next, the search over thousands of regions is the exception.
__device__ __forceinline__ int find_region_near(const DevRegion* regs, int n, int64_t item, int r0) {
  if (r0 + 1 >= n || regs[r0 + 1].item_base > item) return r0;
  if (r0 + 2 >= n || regs[r0 + 2].item_base > item) return r0 + 1;
  return find_region(regs, n, item);
}


// 4-bit packing of the reference slice (one thread per 8 bases)
__global__ void rv_pack_ref_kernel(const char* ref, int64_t n, uint32_t* out, int64_t n_words) {
  int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const uint32_t v = pack_ref8(ref, n, w);
  out[w] = v;
}

static const unsigned long long WALK_COUNT_STATS = 1ull << 61;  // the walk kernel's prepare_read counts the read's statistics

// rv_pileup_kernel — classify.  One (region, read) pair per thread.
//   * the record filters of RecordPreprocessor::next_record (recordPreprocessor.cpp:121-176, -t included);
//   * a read whose CIGAR is one M op over the whole read — the bulk of any data set — is handled here completely:
//     NM filter, the plain-run proof over the read, and the observation that CigarModifier (cigarModifier.cpp:365-377)
//     leaves such a read alone when its first three and last three bases equal the reference (with them inside the loaded
//     window the two mismatch walks, :412-439 and :529-568, stop after three matches with nothing to clip).  The
//     remaining tests of parseCigar's preamble (:584-603) are applied and a plain read leaves its gather descriptor;
//   * every other read (soft clips, indels, hard clips, mismatching ends, N bases, clustered mismatches) is appended
//     to the walk queue: rv_walk_kernel runs prepare_read + walk_read on it.
__global__ void __launch_bounds__(128) rv_pileup_kernel(PileupArgs a) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int kept = 0, kept_bases = 0;
  int back = 0, reach = 0;
  bool queue = false;
  unsigned long long queue_entry = 0;
  const int r_first = a.itemblk_region[blockIdx.x];  // (blockDim.x == 128) region of the CTA's first item
  if (item < a.n_items) {
    const int ri = find_region_near(a.regions, a.n_regions, item, r_first);
    const DevRegion* dr = a.regions + ri;
    const int64_t read_idx = dr->r.read_lo + (item - dr->item_base);
    const rv_read rd = a.reads[read_idx - dr->read_bias];
    GDesc gd;
    gd.m_start = 0; gd.m_len = 0; gd.re0 = 0; gd.qual_off = 0; gd.rlen = 0; gd.mapq = 0; gd.dir_nm = 0;
    uint32_t mm_blocks = 0;
    bool store_ml = false;
    PlainScan ps;
    ps.ml_lo = ps.ml_hi = 0;
    // htslib iterator overlap test (sam_itr_next): pos0 < end && endpos > beg0; then recordPreprocessor.cpp:121-146
    bool pass = rd.pos - 1 < dr->r.end && rd.end_pos > dr->r.start - 1 && (rd.flag & a.P.samfilter) == 0 &&
                (int)rd.mapq >= a.P.mapping_quality && rd.l_seq != 1 && rd.n_cigar > 0;
    if (pass && a.P.dedup && is_duplicate_read(a.P, dr->r, a.reads - dr->read_bias, a.pool - dr->pool_bias, read_idx)) pass = false;
    if (pass) {
      const int64_t pool_off = (int64_t)rd.data_off16 * 16 - dr->pool_bias;  // device pool byte of the read's variable part
      const uint32_t c0 = *(const uint32_t*)(a.pool + pool_off);
      bool simple = rd.n_cigar == 1 && c_op(c0) == OP_M && c_len(c0) == rd.l_seq && rd.l_seq >= 8 && rd.l_seq <= 8192 &&
                    !a.force_exact && a.P.trim_bases_after == 0;
      int nm = 0;
      if (simple) {  // parseCigar.cpp:514-534 with no indel bases
        if (rd.nm >= 0) { nm = rd.nm; if (nm > a.P.mismatch) { pass = false; simple = false; } }
        else if (rd.flag & 4) { pass = false; simple = false; }
        if (nm > 127) simple = false;
      }
      if (simple) {
        const int ml = rd.l_seq;
        const int E0 = rd.pos - a.ref_start;
        const int64_t w_lo = dr->r.ref_lo > a.ref_start ? dr->r.ref_lo : a.ref_start;
        const int64_t w_hi = (int64_t)dr->r.ref_hi < a.ref_start + a.ref_n - 1 ? (int64_t)dr->r.ref_hi : a.ref_start + a.ref_n - 1;
        simple = E0 >= 0 && rd.pos >= w_lo && (int64_t)rd.pos + ml - 1 <= w_hi;
        if (simple) {
          const uint32_t* sq = (const uint32_t*)(a.pool + pool_off) + 1;
          const bool plain = simd_plain_scan(sq, a.ref4, E0, 0, ml, a.P.vext + 1, &ps) && ps.ml_n <= 8;
          // CigarModifier leaves the read alone only with clean ends (-k 1); a base that is not A/C/G/T anywhere sends
          // the read to the literal path as well
          const bool ends_dirty = a.P.local_realign && (ps.special != 0 || mismatch_near_ends(ps, ml, 3, 3));
          if (ends_dirty) simple = false;
          else {
            // the rest of parseCigar's preamble for an unchanged "<l_seq>M": -M, maxReadLength, supplementary (:591-603)
            if (a.P.minmatch != 0 && ml < a.P.minmatch) pass = false;
            else {
              if (ml > a.max_rl[ri]) atomicMax(a.max_rl + ri, ml);
              if (rd.flag & 2048) pass = false;
              else {
                kept = 1;
                kept_bases = ml;
                // parseCigar.cpp:630 / skipOverlappingReads :182-206 — the -u / --UN test happens once, before the first op
                const bool dir = (rd.flag & 16) != 0;
                const bool paired_same = (rd.flag & 1) && rd.mate_same_tid;
                bool skip = a.P.uniq_u && paired_same && !dir && rd.pos >= rd.mpos;
                if (!skip && a.P.uniq_un && paired_same && rd.pos >= rd.mpos && rd.pos <= rd.mpos + ml - 1) skip = true;
                if (!skip) {
                  if (plain) {
                    gd.m_start = rd.pos;
                    gd.m_len = (uint16_t)ml;
                    gd.re0 = 0;
                    gd.qual_off = (uint32_t)(pool_off + 4 + ((ml + 1) >> 1));
                    gd.rlen = (uint16_t)ml;
                    gd.mapq = rd.mapq;
                    gd.dir_nm = (uint8_t)((dir ? 0x80 : 0) | nm);
                    mm_blocks = ps.mm_blocks;
                    store_ml = mm_blocks != 0;
                    reach = ml;
                  } else {  // clustered mismatches away from the ends: the literal walk (statistics are counted already)
                    queue = true;
                    queue_entry = (unsigned long long)item;  // (class 0)
                  }
                }
              }
            }
          }
        }
      }
      if (pass && !simple && !kept) {
        queue = true;
        int cls = 7;
        if (rd.n_cigar == 1) cls = 0;
        else if (rd.n_cigar <= 6) {
          const uint32_t* cg = (const uint32_t*)(a.pool + pool_off);
          int n_ins = 0, n_del = 0, n_other = 0, n_m = 0, n_clip = 0;
          for (int k = 0; k < (int)rd.n_cigar; ++k) {
            const int op = c_op(cg[k]);
            if (op == OP_M) n_m++;
            else if (op == OP_I) n_ins++;
            else if (op == OP_D) n_del++;
            else if (op == OP_S) n_clip++;
            else n_other++;
          }
          if (n_other == 0 && n_ins + n_del == 0 && n_m == 1)
            cls = n_clip == 2 ? 3 : c_op(cg[0]) == OP_S ? 1 : 2;
          else if (n_other == 0 && n_ins + n_del == 1 && n_m == 2)
            cls = n_clip ? 6 : n_del ? 4 : 5;
        }
        queue_entry = (unsigned long long)item | WALK_COUNT_STATS | ((unsigned long long)cls << WALK_CLASS_SHIFT);
      }
    }
    *(uint4*)(a.descs + item) = *(const uint4*)&gd;
    a.desc_mm[item] = (uint16_t)mm_blocks;
    if (store_ml)
      a.desc_mml[item] = make_uint4((uint32_t)ps.ml_lo, (uint32_t)(ps.ml_lo >> 32), (uint32_t)ps.ml_hi, (uint32_t)(ps.ml_hi >> 32));
  }
  // ---- statistics and the candidate-window bounds of the gather kernel: one atomic per warp / block ----
  unsigned long long w_kept = kept, w_bases = kept_bases;
  for (int off = 16; off > 0; off >>= 1) {
    w_kept += __shfl_down_sync(0xffffffffu, w_kept, off);
    w_bases += __shfl_down_sync(0xffffffffu, w_bases, off);
    back = max(back, __shfl_down_sync(0xffffffffu, back, off));
    reach = max(reach, __shfl_down_sync(0xffffffffu, reach, off));
  }
  // queue slots: one atomic per CTA, entries of a CTA stay together in item order, so that neighbouring lanes of
  // rv_walk_kernel walk neighbouring reads (same indel site, same CIGAR shape: the lanes stay in step)
  __shared__ unsigned long long sh[2];
  __shared__ unsigned s_qcnt[4], s_cls[WALK_CLASSES];
  __shared__ unsigned long long s_qbase;
  const unsigned b0 = __ballot_sync(0xffffffffu, queue);
  if (lane == 0) s_qcnt[threadIdx.x >> 5] = __popc(b0);
  if (threadIdx.x < 2) sh[threadIdx.x] = 0;
  if (threadIdx.x < WALK_CLASSES) s_cls[threadIdx.x] = 0;
  __syncthreads();
  if (queue) atomicAdd(&s_cls[(queue_entry >> WALK_CLASS_SHIFT) & (unsigned long long)(WALK_CLASSES - 1)], 1u);
  if (threadIdx.x == 0) {
    const unsigned tot = s_qcnt[0] + s_qcnt[1] + s_qcnt[2] + s_qcnt[3];
    s_qbase = tot ? atomicAdd(a.walk_count, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  if (threadIdx.x < WALK_CLASSES && s_cls[threadIdx.x]) atomicAdd(a.walk_count + 8 + threadIdx.x, (unsigned long long)s_cls[threadIdx.x]);
  if (queue) {
    unsigned long long slot = s_qbase;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) slot += s_qcnt[w];
    slot += __popc(b0 & ((1u << lane) - 1u));
    a.walk_queue[slot] = queue_entry;
  }
  if (lane == 0) {
    if (w_kept) atomicAdd(&sh[0], w_kept);
    if (w_bases) atomicAdd(&sh[1], w_bases);
    if (back > a.reach[0]) atomicMax(a.reach + 0, back);
    if (reach > a.reach[1]) atomicMax(a.reach + 1, reach);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sh[0]) atomicAdd(&a.stats->n_kept, sh[0]);
    if (sh[1]) atomicAdd(&a.stats->n_bases, sh[1]);
  }
}

// Groups the walk queue by class (stable inside a chunk of 256 entries, chunks in arbitrary order: neighbouring reads —
// same indel site, same CIGAR shape — stay together).  Class c starts at the sum of the counts of the classes walked
// before it.
__global__ void __launch_bounds__(256) rv_walk_sort_kernel(PileupArgs a) {
  const unsigned long long n = a.walk_count[0];
  __shared__ unsigned s_cnt[WALK_CLASSES];
  __shared__ unsigned long long s_base[WALK_CLASSES], s_off[WALK_CLASSES];
  if (threadIdx.x == 0) {  // class c starts after the classes walked before it (7, 6, ... c + 1)
    unsigned long long off = 0;
    for (int c = WALK_CLASSES - 1; c >= 0; --c) { s_off[c] = off; off += a.walk_count[8 + c]; }
  }
  const int lane = threadIdx.x & 31;
  for (unsigned long long start = (unsigned long long)blockIdx.x * 256; start < n; start += (unsigned long long)gridDim.x * 256) {
    if (threadIdx.x < WALK_CLASSES) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long q = start + threadIdx.x;
    const bool have = q < n;
    const unsigned long long e = have ? a.walk_queue[q] : 0ull;
    const int c = have ? (int)((e >> WALK_CLASS_SHIFT) & (unsigned long long)(WALK_CLASSES - 1)) : -1;
    unsigned rank = 0, wbase = 0;
#pragma unroll
    for (int cc = 0; cc < WALK_CLASSES; ++cc) {
      const unsigned mk = __ballot_sync(0xffffffffu, c == cc);
      if (mk == 0) continue;
      unsigned wb = 0;
      if (lane == 0) wb = atomicAdd(&s_cnt[cc], (unsigned)__popc(mk));
      wb = __shfl_sync(0xffffffffu, wb, 0);
      if (c == cc) { rank = __popc(mk & ((1u << lane) - 1u)); wbase = wb; }
    }
    __syncthreads();
    if (threadIdx.x < WALK_CLASSES)
      s_base[threadIdx.x] = s_off[threadIdx.x] +
                            (s_cnt[threadIdx.x] ? atomicAdd(a.walk_count + 16 + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]) : 0ull);
    __syncthreads();
    if (have) a.walk_queue2[s_base[c] + wbase + rank] = e;
    __syncthreads();
  }
}

// rv_walk_kernel — the queued reads, one per thread: prepare_read (parseCigar's preamble: NM filter, every CigarModifier
// rule, clean-up, the clip / length / supplementary filters), then walk_read with a ListSink: plain matched stretches
// leave gather descriptors, the rest goes to the SparseObs list / the events.  All 32 lanes of a warp go through
// walk_read together (its CIGAR-op rounds are warp-synchronous).  The grid is sized to what is resident (MIN_CTAS per
// SM); a warp takes its next 32 queue entries from a cursor.
template <int MIN_CTAS>
__global__ void __launch_bounds__(128, MIN_CTAS) rv_walk_kernel(PileupArgs a) {
  const unsigned long long n = a.walk_count[0];
  unsigned long long over = 0, unsup = 0, clip = 0, kept = 0, bases = 0, segs = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) { a.stats->n_walk_items = n; a.stats->n_walk_full = n; }
  const int lane = threadIdx.x & 31;
  unsigned long long slot_next = 0, slot_end = 0;  // this lane's reserved stretch of the SparseObs list
  for (;;) {
    unsigned long long q0 = 0;
    if (lane == 0) q0 = atomicAdd(a.walk_count + 2, 32ull);
    q0 = __shfl_sync(0xffffffffu, q0, 0);
    if (q0 >= n) break;
    const unsigned long long q = q0 + lane;
    bool work = q < n;
    const unsigned long long entry = work ? a.walk_queue2[q] : 0ull;
    const int64_t item = (int64_t)(entry & WALK_ITEM_MASK);
    const int ri = find_region(a.regions, a.n_regions, item);
    const DevRegion* dr = a.regions + ri;
    const int64_t read_idx = dr->r.read_lo + (item - dr->item_base);
    rv_read rd;
    if (work) rd = a.reads[read_idx - dr->read_bias];
    else { rd.pos = 0; rd.mpos = 0; rd.data_off16 = 0; rd.l_seq = 0; rd.flag = 0; rd.n_cigar = 0; rd.nm = 0; rd.mapq = 0; rd.mate_same_tid = 0; rd.end_pos = 0; rd.mtid = 0; }
    const uint8_t* pool = a.pool - dr->pool_bias;
    RefView ref;
    ref.bases = a.ref;
    ref.base_pos = a.ref_start;
    ref.n = a.ref_n;
    ref.lo = dr->r.ref_lo;
    ref.hi = dr->r.ref_hi;
    ListSink s;
    s.a = &a;
    s.bind(dr);
    s.item = item;
    s.n_seg = 0;
    s.first_taken = false;
    s.want_segments = !a.force_exact;
    s.rd = &rd;
    s.pool_off = (int64_t)rd.data_off16 * 16 - dr->pool_bias;
    s.n_unsup = s.n_over = s.n_clip = s.n_ev = 0;
    s.slot_next = slot_next;
    s.slot_end = slot_end;
    s.ridx1 = a.P.trim_bases_after != 0 ? (uint32_t)read_idx + 1u : 0u;
    Prep pr;
    pr.ok = false;
    pr.n_cigar = 0;
    const rv_region R = dr->r;  // by value: the walk compares against start / end at every base
    if (work) {
      DeviceSink ps;
      ps.a = &a;
      ps.dr = dr;
      ps.mute = (entry & WALK_COUNT_STATS) == 0;  // counted by rv_pileup_kernel already
      ps.kept_bases = ps.n_kept = ps.n_unsup = ps.n_over = 0;
      prepare_read(a.P, R, rd, pool, ref, ps, false, pr);
      kept += ps.n_kept;
      bases += ps.kept_bases;
      unsup += ps.n_unsup;
    }
    work = work && pr.ok;
    __syncwarp();
    walk_read(a.P, R, ri, rd, pool, ref, (uint32_t)read_idx, s, pr, (FastDesc*)0, 0, work);
    over += s.n_over;
    unsup += s.n_unsup;
    clip += s.n_clip;
    segs += s.n_seg;
    slot_next = s.slot_next;
    slot_end = s.slot_end;
  }
  for (; slot_next < slot_end; ++slot_next)  // unused slots of the last chunk
    if (slot_next < a.max_sparse) *(uint4*)(a.sparse + slot_next) = make_uint4(0xffffffffu, 0u, 0u, 0u);
  if (segs) atomicAdd(&a.stats->n_segments, segs);
  if (clip) atomicAdd(&a.stats->n_clipped, clip);
  if (over) atomicAdd(&a.stats->n_overflow, over);
  if (unsup) atomicAdd(&a.stats->n_unsupported, unsup);
  if (kept) atomicAdd(&a.stats->n_kept, kept);
  if (bases) atomicAdd(&a.stats->n_bases, bases);
}

// The SparseObs list onto the tables (after the gather kernel has stored every row): one entry per thread.
__device__ __forceinline__ void row_observe(uint32_t* row, uint32_t dir, uint32_t tp, uint32_t q, uint32_t mapq, uint32_t nm, int thr);
// -T (pass != 0): the subtraction of an insertion's anchor base applies only if this or an earlier read created the row
// (see the Sink concept in rv_core.cuh).  Pass 1 applies everything else and records every row's first contributor
// (atomicMin of the read index), pass 2 — a second launch — applies the subtractions whose read is not before it.
__global__ void __launch_bounds__(256) rv_apply_kernel(const SparseObs* list, const unsigned long long* count, unsigned long long cap,
                                                        uint32_t* counts, uint32_t* cov, uint8_t* touched, int thr, DevStats* stats,
                                                        uint32_t* first_read, int pass) {
  unsigned long long n = *count;
  if (blockIdx.x == 0 && threadIdx.x == 0 && pass != 2) stats->n_sparse = n;
  if (n > cap) n = cap;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint4 v = *(const uint4*)(list + i);
    if (v.x == 0xffffffffu) continue;  // unused slot of a lane's chunk
    const uint32_t flags = v.z >> 16, kind = (flags >> 3) & 3u;
    if (kind == SO_COV) { if (pass != 2) atomicAdd(cov + v.x, 1u); continue; }
    if (pass == 1) {
      if (kind == SO_SUB) continue;
      if (v.w) atomicMin(first_read + (size_t)v.x * 4 + (flags & 3u), v.w - 1u);
    } else if (pass == 2) {
      if (kind != SO_SUB || !v.w || first_read[(size_t)v.x * 4 + (flags & 3u)] > v.w - 1u) continue;
    }
    uint32_t* row = counts + ((size_t)v.x * 4 + (flags & 3u)) * RV_ROW_U32;
    touched[v.x] = 1;  // (whichever allele: the screen pass then looks at the position's rows)
    const uint32_t dir = (flags >> 2) & 1u, tp = v.y & 0xffffu, q = (v.y >> 16) & 0xffu, mapq = v.y >> 24;
    const int nm = (int)(int16_t)(v.z & 0xffffu);
    if (kind == SO_SINGLE) {
      row_observe(row, dir, tp, q, mapq, (uint32_t)nm, thr);
    } else {
      // addCnt (parseCigar.cpp:300-325) / the subtraction of an insertion's anchor base (:1471-1490): no pstd / qstd
      const uint32_t sg = kind == SO_ADD ? 1u : 0u - 1u;
      atomicAdd(row + (dir ? RV_F_REV : RV_F_FWD), sg);
      atomicAdd(row + RV_F_SUM_TP, sg * tp);
      atomicAdd(row + RV_F_SUM_Q, sg * q);
      atomicAdd(row + RV_F_SUM_MAPQ, sg * mapq);
      if (nm) atomicAdd(row + RV_F_SUM_NM, sg * (uint32_t)nm);
      if ((int)q >= thr) atomicAdd(row + RV_F_HI, sg);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Position-major accumulation of the plain segments (rv_gather4_kernel): arguments and the tile index.
// ------------------------------------------------------------------------------------------------
struct GatherArgs {
  double goodq;
  const DevRegion* regions;
  int n_regions;
  const rv_read* reads;
  const GDesc* descs;
  const uint16_t* desc_mm;
  const uint8_t* pool;
  const char* ref;
  int32_t ref_start;
  int64_t ref_n;
  uint32_t* counts;
  uint32_t* cov;
  const int32_t* reach;
  int64_t* tile_range;  // [2 * tile]: first candidate read of the tile; number of candidates | region << 32
  int64_t n_tiles;
  int tile;             // table positions per tile (G4_W)
  const uint4* desc_mml;
  const GDesc* descs2;  // second segments (GD_HAS_SECOND)
  const uint16_t* desc_mm2;
  const uint4* desc_mml2;
  int64_t pool_bytes;   // bytes of the device pool (rv_gather4_kernel clamps its look-ahead loads to it)
  int thr;              // ceil(goodq)
  int run;              // consecutive tiles per warp (rv_gather4_kernel)
  int alternate;        // odd tiles walk their candidates downward
  uint8_t* touched;     // [table position]: set where a base that differs from the reference was added
};

__device__ __forceinline__ int find_region_by_tile(const DevRegion* regs, int n, int64_t tile) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (regs[mid].tile_base <= tile) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// candidate reads of every tile: read start in (c_lo - max_reach, c_hi + max_back]  (one thread per tile)
__global__ void rv_tile_index_kernel(GatherArgs a) {
  const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= a.n_tiles) return;
  const int ri = find_region_by_tile(a.regions, a.n_regions, tile);
  const DevRegion* dr = a.regions + ri;
  const int p_lo = dr->first_pos + (int)(tile - dr->tile_base) * a.tile;
  const int c_lo = p_lo > dr->r.start ? p_lo : dr->r.start;
  const int c_hi = p_lo + a.tile - 1 < dr->r.end ? p_lo + a.tile - 1 : dr->r.end;
  int64_t lo = dr->r.read_lo, hi = dr->r.read_lo;
  if (c_lo <= c_hi) {
    const int want_lo = c_lo - a.reach[1];  // pos > want_lo
    const int want_hi = c_hi + a.reach[0];  // pos <= want_hi
    int64_t x = dr->r.read_lo, y = dr->r.read_hi;
    const rv_read* rr = a.reads - dr->read_bias;
    while (x < y) { const int64_t m = (x + y) >> 1; if (rr[m].pos <= want_lo) x = m + 1; else y = m; }
    lo = x;
    y = dr->r.read_hi;
    while (x < y) { const int64_t m = (x + y) >> 1; if (rr[m].pos <= want_hi) x = m + 1; else y = m; }
    hi = x;
  }
  a.tile_range[2 * tile] = lo;
  a.tile_range[2 * tile + 1] = (int64_t)(((unsigned long long)(unsigned)ri << 32) | (unsigned)(hi - lo));
}

// ------------------------------------------------------------------------------------------------
// rv_gather4_kernel — the plain matched runs, four table positions per lane.
// A WARP owns G4_W = 128 consecutive table positions (lane l: positions 4l..4l+3); warps share nothing, so the
// kernel has no CTA barrier.  What a plain run adds to the reference allele of a position splits in two:
//   * per-read constants (count, reverse count, mapq, nm) and tp = min(k+1, L-k), a tent over the run: range adds /
//     second-order differences into per-warp shared-memory arrays, 2-5 shared atomics per (read, warp), prefix-summed
//     once per tile;
//   * what depends on the base quality (sum, high-quality count) and the "all observations equal" tests of pstd/qstd
//     (AND- and OR-reductions of tp and q): one pass over the warp's reads with byte / halfword SIMD in registers:
//     tp of two positions per VIADDMNMX.S16x2.RELU (0 = the read does not cover the position), the four qualities
//     as one unaligned 32-bit word (two LDG.32 + PRMT, prefetched G4_PF records ahead).
// Mismatching bases (listed by the classify kernel, rv_pileup_kernel: desc_mml) are handled by the lane that owns
// the READ while it builds the record: it masks them out of the SIMD pass (a 128-bit exclusion map per record),
// takes their share back out of the difference arrays and adds them to the row of their allele with the same
// atomics rv_walk_kernel uses.  The rows those atomics land in are zeroed by the warp before its first read.
// ------------------------------------------------------------------------------------------------
static const int G4_W = 128;
static const int G4_WARPS = 4;
static const int G4_B = 16384;    // bias that keeps tile coordinates non-negative halfwords

template <int G4_PF>  // quality words are fetched this many records ahead
struct G4Warp {
  uint4 rec[32 + 2 * G4_PF];      // {-(s+B) x2, (s+L+1+B) x2, PRMT selector | flag << 16, -}
  uint4 ldr[32 + 2 * G4_PF];      // {word index of the quality under coordinate 0, first lane, last lane the run covers, -}
  uint2 rec2[32 + 2 * G4_PF];     // segments that are not a whole read: {-(s+B) + re0 x2, (s+L+1+B) + tail x2} (tp of the SIMD pass)
  uint8_t excl[32][32];           // [record][lane]: bit j = position j of the lane is masked out of the SIMD pass
  uint32_t d_n[G4_W + 4], c_n[G4_W + 4], d_rev[G4_W + 4], d_mapq[G4_W + 4], d_nm[G4_W + 4], d_tp1[G4_W + 4], d_tp2[G4_W + 4];
  uint32_t g_tp[G4_W];            // sum of tp of the segments that are not a whole read (each lane owns its four)
  int8_t refal_s[G4_W];           // reference allele of every position of the tile (-1 = none)
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// the observation of one base on its allele's row, as DeviceSink::single does it
__device__ __forceinline__ void row_observe(uint32_t* row, uint32_t dir, uint32_t tp, uint32_t q, uint32_t mapq, uint32_t nm, int thr) {
  atomicAdd(row + (dir ? RV_F_REV : RV_F_FWD), 1u);
  atomicAdd(row + RV_F_SUM_TP, tp);
  atomicAdd(row + RV_F_SUM_Q, q);
  atomicAdd(row + RV_F_SUM_MAPQ, mapq);
  if (nm) atomicAdd(row + RV_F_SUM_NM, nm);
  if ((int)q >= thr) atomicAdd(row + RV_F_HI, 1u);
  const uint32_t mine = (tp & 0xffffu) | ((q & 0xffu) << 16) | (1u << 31);
  const uint32_t old = atomicCAS(row + RV_F_STD, 0u, mine);
  if (old != 0u) {
    uint32_t bits = 0;
    if ((old ^ mine) & 0xffffu) bits |= 1u << 24;
    if ((old ^ mine) & 0xff0000u) bits |= 1u << 25;
    if (bits & ~old) atomicOr(row + RV_F_STD, bits);
  }
}

// inclusive prefix sum over the warp's 128 values, four consecutive ones per lane
__device__ __forceinline__ void warp_scan4(uint32_t v[4], int lane) {
  v[1] += v[0]; v[2] += v[1]; v[3] += v[2];
  uint32_t t = v[3];
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, t, off);
    if (lane >= off) t += y;
  }
  const uint32_t ex = t - v[3];
  v[0] += ex; v[1] += ex; v[2] += ex; v[3] += ex;
}

template <int G4_PF, int MIN_CTAS>
__global__ void __launch_bounds__(G4_WARPS * 32, MIN_CTAS) rv_gather4_kernel(GatherArgs a) {
  __shared__ __align__(16) G4Warp<G4_PF> s_w[G4_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  G4Warp<G4_PF>& W = s_w[warp];
  // A warp works through a.run consecutive tiles: the reads a tile shares with the next one are needed again right
  // after they were used (L2 / L1 hits), not a whole tile's duration later by some other warp.
  const int64_t tile0 = ((int64_t)blockIdx.x * G4_WARPS + warp) * a.run;
#pragma unroll 1
  for (int64_t tile = tile0; tile < tile0 + a.run && tile < a.n_tiles; ++tile) {  // (no CTA barrier anywhere below)
  const int4 tinfo = ((const int4*)a.tile_range)[tile];  // {read lo (64 bit), n reads, region}
  const DevRegion* dr = a.regions + tinfo.w;
  const int64_t lo = (int64_t)(((unsigned long long)(unsigned)tinfo.y << 32) | (unsigned)tinfo.x);
  const int64_t hi = lo + tinfo.z;
  const int p_lo = dr->first_pos + (int)(tile - dr->tile_base) * G4_W;
  const int c_lo = p_lo > dr->r.start ? p_lo : dr->r.start;  // positions that can receive observations
  const int c_hi = p_lo + G4_W - 1 < dr->r.end ? p_lo + G4_W - 1 : dr->r.end;
  const int x4 = lane * 4;
  const int thr = a.thr;
  const int64_t t_row0 = dr->tab_off + (p_lo - dr->first_pos);
  uint32_t* const tile_rows = a.counts + (size_t)t_row0 * RV_POS_U32;
  // ---- the lane's four positions: table / live flags, reference alleles; rows the atomics may land in start at zero
  // (kept in shared memory, not in registers: the pass below needs every register it can get)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = p_lo + x4 + j;
    int ral = -1;
    if (p >= c_lo && p <= c_hi && p >= dr->r.ref_lo && p <= dr->r.ref_hi && p >= a.ref_start && (int64_t)(p - a.ref_start) < a.ref_n) {
      const char c = a.ref[p - a.ref_start];
      ral = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
    }
    W.refal_s[x4 + j] = (int8_t)ral;
  }
  for (int i = lane; i < G4_W + 4; i += 32) {
    W.d_n[i] = 0; W.c_n[i] = 0; W.d_rev[i] = 0; W.d_mapq[i] = 0; W.d_nm[i] = 0; W.d_tp1[i] = 0; W.d_tp2[i] = 0;
  }
  for (int i = lane; i < G4_W; i += 32) W.g_tp[i] = 0;
  __syncwarp();
  {
    const int n_in = dr->n_pos - (p_lo - dr->first_pos) < G4_W ? dr->n_pos - (p_lo - dr->first_pos) : G4_W;
    // Every byte of the tile's rows is stored exactly once.  Now: zeros into the rows of the alleles that are not the
    // reference's (they only ever receive the atomics of bases that differ from the reference: from the lanes of this
    // warp below, from rv_apply_kernel later).  At the end of the tile: the reference allele's row.  Whole 32-byte
    // sectors both times, so the two partial writes of a line cost DRAM what one full write does.
    uint4* blk = (uint4*)tile_rows;
    for (int c = lane; c < n_in * (RV_POS_U32 / 4); c += 32)
      if (((c >> 1) & 3) != (int)W.refal_s[c >> 3]) blk[c] = make_uint4(0, 0, 0, 0);
    for (int x = lane; x < n_in; x += 32) a.touched[t_row0 + x] = 0;
  }
  // the zero rows are in place before any lane's atomics on them: the warp barrier orders the lanes' memory accesses
  // (a device-wide fence here made every tile wait for the acknowledgement of its stores)
  __syncwarp();
  // ---- lane constants of the SIMD pass
  const uint32_t X01 = (uint32_t)(x4 + 1 + G4_B) | ((uint32_t)(x4 + 2 + G4_B) << 16);
  const uint32_t X23 = X01 + 0x00020002u;
  const uint32_t bias4 = (uint32_t)(128 - thr) * 0x01010101u;  // 1 <= thr <= 128 (checked by the host)
  const uint32_t* const pool32 = (const uint32_t*)a.pool;
  const int64_t item_shift = dr->item_base - dr->r.read_lo;  // work item of read index i = i + item_shift
  const GDesc* item_desc = a.descs + item_shift;
  const uint16_t* item_mm = a.desc_mm + item_shift;
  uint32_t tp_and01 = 0xffffffffu, tp_and23 = 0xffffffffu, tp_or01 = 0, tp_or23 = 0, q_and = 0xffffffffu, q_or = 0;
  uint32_t sum_q[4] = {0, 0, 0, 0}, n_hi[4] = {0, 0, 0, 0};

  // Neighbouring tiles walk their candidates in opposite directions: the reads two tiles share are then needed by
  // both at the same end of their loops, i.e. close in time (L1 / L2 hits instead of a second trip to DRAM).
  const bool downward = a.alternate && ((tile - dr->tile_base) & 1) != 0;
  const int n_batches = (int)((hi - lo + 31) >> 5);
  uint4 d_nxt = make_uint4(0, 0, 0, 0);
  uint32_t mm_nxt = 0;
  {
    const int64_t i0 = (downward ? lo + 32 * (int64_t)(n_batches - 1) : lo) + lane;
    if (n_batches > 0 && i0 < hi) {
      d_nxt = *(const uint4*)(item_desc + i0);
      mm_nxt = item_mm[i0];
    }
  }
  for (int bi = 0; bi < n_batches; ++bi) {
    const int64_t i = lo + 32 * (int64_t)(downward ? n_batches - 1 - bi : bi) + lane;
    GDesc d;
    *(uint4*)&d = d_nxt;
    uint32_t mm = mm_nxt;
    d_nxt = make_uint4(0, 0, 0, 0);  // (m_len = 0) the next round's descriptor travels while this round computes
    mm_nxt = 0;
    {
      const int64_t i_n = downward ? i - 32 : i + 32;
      if (bi + 1 < n_batches && i_n < hi) {
        d_nxt = *(const uint4*)(item_desc + i_n);
        mm_nxt = item_mm[i_n];
      }
    }
    bool second = (d.m_len & GD_HAS_SECOND) != 0;  // the work item has a second segment (a read with an indel)
    d.m_len &= (uint16_t)GD_LEN_MASK;
    const uint4* cur_mml = a.desc_mml + item_shift;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      if (!__any_sync(0xffffffffu, second)) break;
      d.m_len = 0;
      mm = 0;
      if (second) {
        *(uint4*)&d = *(const uint4*)(a.descs2 + item_shift + i);
        d.m_len &= (uint16_t)GD_LEN_MASK;
        mm = a.desc_mm2[item_shift + i];
      }
      cur_mml = a.desc_mml2 + item_shift;
    }
    // ---- 1. one candidate descriptor per lane -> record of the SIMD pass, range adds, mismatching bases -----------
    bool take = false;
    int k_lo = 0, k_hi = 0;
    if (d.m_len != 0) {
      k_lo = c_lo - d.m_start > 0 ? c_lo - d.m_start : 0;
      k_hi = c_hi + 1 - d.m_start < (int)d.m_len ? c_hi + 1 - d.m_start : (int)d.m_len;
      take = k_lo < k_hi;
    }
    const unsigned bt = __ballot_sync(0xffffffffu, take);
    const int n_rec = __popc(bt);
    int g_n = 0, g_rev = 0, g_mapq = 0, g_nm = 0, g_tp2 = 0, g_tp1 = 0;  // what this lane adds at coordinate 0
    if (take) {
      const int slot = __popc(bt & ((1u << lane) - 1u));
      const int s = d.m_start - p_lo, L = (int)d.m_len;
      const int re0 = (int)d.re0, tail = (int)d.rlen - re0 - L;
      const bool general = (re0 | tail) != 0;  // not a whole read: tp = min(re0 + k + 1, L - k + tail)
      const uint32_t dir = d.dir_nm >> 7, nm = d.dir_nm & 0x7fu, mapq = d.mapq;
      const int64_t qb = (int64_t)d.qual_off;  // pool byte of the segment's first quality
      const int64_t q0 = qb - s;               // ... of the quality under coordinate 0
      bool any_ex = false;
      if (mm) {
        // mismatching bases of the segment that fall on this warp's live positions
        const uint4 ml = cur_mml[i];
        unsigned long long l_lo = (unsigned long long)ml.x | ((unsigned long long)ml.y << 32);
        unsigned long long l_hi = (unsigned long long)ml.z | ((unsigned long long)ml.w << 32);
#pragma unroll 1
        while (l_lo | l_hi) {
          const uint32_t e = (uint32_t)l_lo & 0xffffu;
          l_lo = (l_lo >> 16) | (l_hi << 48);
          l_hi >>= 16;
          const int k = (int)((e >> 2) & 0x1fffu);
          if ((e & 0x8000u) && k >= k_lo && k < k_hi) {
            const int x = s + k;
            if (!any_ex) {
              any_ex = true;
              ((uint4*)W.excl[slot])[0] = make_uint4(0, 0, 0, 0);
              ((uint4*)W.excl[slot])[1] = make_uint4(0, 0, 0, 0);
            }
            W.excl[slot][x >> 2] |= (uint8_t)(1u << (x & 3));
            const uint32_t tp = (uint32_t)min(re0 + k + 1, L - k + tail);
            const uint32_t q = a.pool[qb + k];
            row_observe(tile_rows + (size_t)x * RV_POS_U32 + (e & 3u) * RV_ROW_U32, dir, tp, q, mapq, nm, thr);
            a.touched[t_row0 + x] = 1;
            atomicAdd(&W.c_n[x], 1u);
            if (dir) { atomicAdd(&W.d_rev[x], 0u - 1u); atomicAdd(&W.d_rev[x + 1], 1u); }
            atomicAdd(&W.d_mapq[x], 0u - mapq); atomicAdd(&W.d_mapq[x + 1], mapq);
            if (nm) { atomicAdd(&W.d_nm[x], 0u - nm); atomicAdd(&W.d_nm[x + 1], nm); }
            if (!general) { atomicAdd(&W.d_tp1[x], 0u - tp); atomicAdd(&W.d_tp1[x + 1], tp); }
          }
        }
      }
      const uint32_t flag = (any_ex ? 0x10000u : 0u) | (general ? 0x20000u : 0u);
      W.rec[slot] = make_uint4(((uint32_t)(-(s + G4_B)) & 0xffffu) * 0x10001u, ((uint32_t)(s + L + 1 + G4_B) & 0xffffu) * 0x10001u,
                               (0x3210u + 0x1111u * (uint32_t)(q0 & 3)) | flag, 0u);
      if (general)
        W.rec2[slot] = make_uint2(((uint32_t)(-(s + G4_B) + re0) & 0xffffu) * 0x10001u,
                                  ((uint32_t)(s + L + 1 + G4_B + tail) & 0xffffu) * 0x10001u);
      // lanes outside the run load what its nearest covered lane loads (no sector that no lane needs is fetched)
      W.ldr[slot] = make_uint4((uint32_t)(int)(q0 >> 2), (uint32_t)((s > 0 ? s : 0) >> 2), (uint32_t)((s + L - 1 < G4_W - 1 ? s + L - 1 : G4_W - 1) >> 2), 0u);
      // range adds of the per-read constants over [s, s+L) clipped to the tile (index G4_W is never read)
      const int rb = s + L;
      if (s > 0) {
        atomicAdd(&W.d_n[s], 1u);
        if (dir) atomicAdd(&W.d_rev[s], 1u);
        atomicAdd(&W.d_mapq[s], mapq);
        if (nm) atomicAdd(&W.d_nm[s], nm);
      } else {  // runs that begin at or before coordinate 0 all add there: summed over the warp below
        g_n = 1; g_rev = (int)dir; g_mapq = (int)mapq; g_nm = (int)nm;
      }
      if (rb < G4_W) {
        atomicAdd(&W.d_n[rb], 0u - 1u);
        if (dir) atomicAdd(&W.d_rev[rb], 0u - 1u);
        atomicAdd(&W.d_mapq[rb], 0u - mapq);
        if (nm) atomicAdd(&W.d_nm[rb], 0u - nm);
      }
      if (!general) {
        // tp = min(k+1, L-k): slope +1 on [s, s+h), -1 on [s+L-h+1, s+L], h = ceil(L/2); second differences, the part
        // left of the tile folded into coordinate 0
        const int h = (L + 1) >> 1;
        const int i0 = s, i1 = s + h, i2 = s + L - h + 1, i3 = s + L + 1;
        if (i0 <= 0) g_tp2 += 1; else if (i0 < G4_W) atomicAdd(&W.d_tp2[i0], 1u);
        if (i1 <= 0) g_tp2 -= 1; else if (i1 < G4_W) atomicAdd(&W.d_tp2[i1], 0u - 1u);
        if (i2 <= 0) g_tp2 -= 1; else if (i2 < G4_W) atomicAdd(&W.d_tp2[i2], 0u - 1u);
        if (i3 <= 0) g_tp2 += 1; else if (i3 < G4_W) atomicAdd(&W.d_tp2[i3], 1u);
        if (s < 0) g_tp1 = min(-s, s + L + 1);  // tp under coordinate -1
      }
    }
    if (__any_sync(0xffffffffu, g_n != 0)) {
      const int t_n = __reduce_add_sync(0xffffffffu, g_n), t_rev = __reduce_add_sync(0xffffffffu, g_rev);
      const int t_mapq = __reduce_add_sync(0xffffffffu, g_mapq), t_nm = __reduce_add_sync(0xffffffffu, g_nm);
      const int t_tp2 = __reduce_add_sync(0xffffffffu, g_tp2), t_tp1 = __reduce_add_sync(0xffffffffu, g_tp1);
      if (lane == 0) {
        atomicAdd(&W.d_n[0], (uint32_t)t_n); atomicAdd(&W.d_rev[0], (uint32_t)t_rev); atomicAdd(&W.d_mapq[0], (uint32_t)t_mapq);
        atomicAdd(&W.d_nm[0], (uint32_t)t_nm); atomicAdd(&W.d_tp2[0], (uint32_t)t_tp2); atomicAdd(&W.d_tp1[0], (uint32_t)t_tp1);
      }
    }
    if (lane < 2 * G4_PF) {  // records that cover nothing: the pass runs in steps of G4_PF and prefetches G4_PF ahead
      W.rec[n_rec + lane] = make_uint4(((uint32_t)(-G4_B) & 0xffffu) * 0x10001u, ((uint32_t)(1 + G4_B) & 0xffffu) * 0x10001u, 0x3210u, 0u);
      W.ldr[n_rec + lane] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    // ---- 2. the SIMD pass over the records ----------------------------------------------------------------------
    if (n_rec) {
      uint32_t w0[G4_PF], w1[G4_PF];
#pragma unroll
      for (int u = 0; u < G4_PF; ++u) {
        const uint4 lr = W.ldr[u];
        const int widx = (int)lr.x + max(min(lane, (int)lr.z), (int)lr.y);
        w0[u] = __ldg(pool32 + widx);
        w1[u] = __ldg(pool32 + widx + 1);
      }
      uint32_t sq_lo = 0, sq_hi = 0, hi_acc = 0;  // at most 32 + G4_PF records: no byte / halfword can overflow
      uint32_t stp01 = 0, stp23 = 0;              // (tp <= 900 for the segments summed here: rlen <= 1800)
      for (int r = 0; r < n_rec; r += G4_PF) {
#pragma unroll
        for (int u = 0; u < G4_PF; ++u) {
          const uint4 rc = W.rec[r + u];
          const uint32_t dn01 = rc.y - X01, dn23 = rc.y - X23;
          uint32_t t01 = __viaddmin_s16x2_relu(X01, rc.x, dn01);  // min(k+1, L-k) of positions 0, 1 (0 = not covered)
          uint32_t t23 = __viaddmin_s16x2_relu(X23, rc.x, dn23);
          uint32_t m = prmt(t01 + 0x7fff7fffu, t23 + 0x7fff7fffu, 0xfdb9u);  // 0xff per covered position
          if (rc.z & 0x10000u)  // (warp-uniform) the record has bases that differ from the reference under this warp
            m &= ~((((uint32_t)W.excl[r + u][lane] * 0x00204081u) & 0x01010101u) * 0xffu);
          const uint32_t m01 = prmt(m, 0u, 0x1100u), m23 = prmt(m, 0u, 0x3322u);
          if (rc.z & 0x20000u) {  // (warp-uniform) a segment of a read with indels: its bases' read positions
            const uint2 r2 = W.rec2[r + u];
            t01 = __viaddmin_s16x2_relu(X01, r2.x, r2.y - X01);
            t23 = __viaddmin_s16x2_relu(X23, r2.x, r2.y - X23);
            stp01 += t01 & m01;
            stp23 += t23 & m23;
          }
          tp_and01 &= t01 | ~m01; tp_or01 |= t01 & m01;
          tp_and23 &= t23 | ~m23; tp_or23 |= t23 & m23;
          const uint32_t q4 = prmt(w0[u], w1[u], rc.z);
          const uint32_t qm = q4 & m;
          q_and &= q4 | ~m; q_or |= qm;
          sq_lo += qm & 0x00ff00ffu;
          sq_hi += prmt(qm, 0u, 0x4341u);
          hi_acc += __umulhi((((qm & 0x7f7f7f7fu) + bias4) | qm) & 0x80808080u, 1u << 25);
          // fetch for record r + u + G4_PF
          const uint4 lr = W.ldr[r + u + G4_PF];
          const int widx = (int)lr.x + max(min(lane, (int)lr.z), (int)lr.y);
          w0[u] = __ldg(pool32 + widx);
          w1[u] = __ldg(pool32 + widx + 1);
        }
      }
      sum_q[0] += sq_lo & 0xffffu; sum_q[2] += sq_lo >> 16; sum_q[1] += sq_hi & 0xffffu; sum_q[3] += sq_hi >> 16;
      n_hi[0] += hi_acc & 0xffu; n_hi[1] += (hi_acc >> 8) & 0xffu; n_hi[2] += (hi_acc >> 16) & 0xffu; n_hi[3] += hi_acc >> 24;
      if (stp01 | stp23) {
        W.g_tp[x4] += stp01 & 0xffffu; W.g_tp[x4 + 1] += stp01 >> 16; W.g_tp[x4 + 2] += stp23 & 0xffffu; W.g_tp[x4 + 3] += stp23 >> 16;
      }
    }
    __syncwarp();
    }  // first / second segment of the batch's work items
  }
  // ---- 3. prefix sums of the difference arrays, rows of the reference alleles -------------------------------------
  __syncwarp();
  uint32_t cnt[4], rev[4], smq[4], snm[4], stp[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    cnt[j] = W.d_n[x4 + j]; rev[j] = W.d_rev[x4 + j]; smq[j] = W.d_mapq[x4 + j]; snm[j] = W.d_nm[x4 + j]; stp[j] = W.d_tp2[x4 + j];
  }
  warp_scan4(cnt, lane); warp_scan4(rev, lane); warp_scan4(smq, lane); warp_scan4(snm, lane); warp_scan4(stp, lane);
#pragma unroll
  for (int j = 0; j < 4; ++j) stp[j] += W.d_tp1[x4 + j];
  warp_scan4(stp, lane);
#pragma unroll
  for (int j = 0; j < 4; ++j) stp[j] += W.g_tp[x4 + j];
  uint32_t cn[4];
  int refal[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { cn[j] = W.c_n[x4 + j]; refal[j] = (int)W.refal_s[x4 + j]; }
  __syncwarp();  // the difference arrays are dead: their memory now stages the tile's reference-allele rows
  uint4* const st_rows = (uint4*)&W;  // [128][2]: first / second half of the position's reference-allele row (ends before W.refal_s)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = x4 + j;
    const int p = p_lo + x;
    if (p - dr->first_pos < dr->n_pos) a.cov[t_row0 + x] = (p >= c_lo && p <= c_hi) ? cnt[j] : 0u;
    const uint32_t n_ref = refal[j] < 0 ? 0u : cnt[j] - cn[j];
    uint4 ra = make_uint4(0, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
    if (n_ref) {
      const uint32_t t_and = ((j & 2) ? tp_and23 : tp_and01) >> (16 * (j & 1)) & 0xffffu;
      const uint32_t t_or = ((j & 2) ? tp_or23 : tp_or01) >> (16 * (j & 1)) & 0xffffu;
      const uint32_t qa = (q_and >> (8 * j)) & 0xffu, qo = (q_or >> (8 * j)) & 0xffu;
      uint32_t w = t_and | (qa << 16) | (1u << 31);
      if (t_and != t_or) w |= 1u << 24;
      if (qa != qo) w |= 1u << 25;
      ra = make_uint4(n_ref - rev[j], rev[j], stp[j], sum_q[j]);
      rb = make_uint4(smq[j], snm[j], n_hi[j], w);
    }
    st_rows[2 * x] = ra;
    st_rows[2 * x + 1] = rb;
  }
  __syncwarp();
  {
    // the reference alleles' rows (zeros where no read matched): two lanes per row, i.e. whole 32-byte sectors
    const int n_in = dr->n_pos - (p_lo - dr->first_pos) < G4_W ? dr->n_pos - (p_lo - dr->first_pos) : G4_W;
    uint4* blk = (uint4*)tile_rows;
    for (int c = lane; c < n_in * 2; c += 32) {
      const int x = c >> 1, al = (int)W.refal_s[x];
      if (al >= 0) blk[x * (RV_POS_U32 / 4) + al * 2 + (c & 1)] = st_rows[c];
    }
  }
  __syncwarp();
  }  // tiles of the warp
}

struct ScoreArgs {
  rv_params P;
  const DevRegion* regions;
  int n_regions;
  int64_t n_positions;  // total table positions
  const char* ref;
  int32_t ref_start;
  int64_t ref_n;
  const uint32_t* counts;
  const uint32_t* cov;
  const rv_patch_entry* patch;
  const uint32_t* patch_first;
  const uint8_t* patch_count;
  const uint8_t* touched;
  const int32_t* tabblk_region;  // region of table position 256 b
  rv_variant* variants;
  unsigned long long max_variants;
  const double* lgt;
  int lgt_n;
  DevStats* stats;
  int64_t* patched_queue;             // table positions that carry patch entries
  unsigned long long* patched_count;
};

struct DeviceEmit {
  const ScoreArgs* a;
  __device__ __forceinline__ unsigned long long reserve() {  // one atomic per group of converged lanes
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned long long slot = 0;
    if (g.thread_rank() == 0) slot = atomicAdd(&a->stats->n_variants, (unsigned long long)g.size());
    return g.shfl(slot, 0) + g.thread_rank();
  }
  __device__ __forceinline__ void emit(const rv_variant& v) {
    const unsigned long long slot = reserve();
    if (slot < a->max_variants) a->variants[slot] = v;
  }
};

__device__ __forceinline__ int find_region_by_tab(const DevRegion* regs, int n, int64_t t) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (regs[mid].tab_off <= t) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ int find_region_by_tab_near(const DevRegion* regs, int n, int64_t t, int r0) {  // see find_region_near
  if (r0 + 1 >= n || regs[r0 + 1].tab_off > t) return r0;
  if (r0 + 2 >= n || regs[r0 + 2].tab_off > t) return r0 + 1;
  return find_region_by_tab(regs, n, t);
}

// Scoring, ToVarsBuilder::process, in two kernels:
//   rv_score_kernel         one table position per thread; positions whose keys are all dense single-base
//                           alleles are scored here (score_dense_position), the MSI context + Fisher test of
//                           their non-reference alleles go through a per-block work list so that all lanes
//                           run them together; positions with patch entries are queued
//   rv_score_patched_kernel the general score_position for the queued positions
// Only positions inside the region are scored (simpleMode.cpp:155-159 drops the rest at output time).
static const int SCORE_BLOCK = 128;
struct ScoreWork {  // 32 bytes
  uint32_t slot;
  int32_t pos, region, a11, a12, a21, a22, pad;
};

struct DenseEmit {
  const ScoreArgs* a;
  DeviceEmit base;
  ScoreWork* work;
  int* n_work;
  int pos, region;
  __device__ __forceinline__ void emit(const rv_variant& v) { base.emit(v); }
  __device__ __forceinline__ void emit_variant(const rv_variant& v, int a11, int a12, int a21, int a22) {
    const unsigned long long slot = base.reserve();
    if (slot >= a->max_variants) return;
    a->variants[slot] = v;
    ScoreWork w;
    w.slot = (uint32_t)slot; w.pos = pos; w.region = region; w.a11 = a11; w.a12 = a12; w.a21 = a21; w.a22 = a22; w.pad = 0;
    work[atomicAdd(n_work, 1)] = w;
  }
};

__global__ void __launch_bounds__(SCORE_BLOCK, 4) rv_score_kernel(ScoreArgs a) {
  __shared__ ScoreWork s_work[SCORE_BLOCK * 3];  // at most three non-reference alleles per position
  __shared__ int s_nwork;
  if (threadIdx.x == 0) s_nwork = 0;
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < a.n_positions) {
    const int ri = find_region_by_tab(a.regions, a.n_regions, t);
    const DevRegion* dr = a.regions + ri;
    const int pos = dr->first_pos + (int)(t - dr->tab_off);
    if (pos >= dr->r.start && pos <= dr->r.end) {
      const uint32_t pf = a.patch_first ? a.patch_first[t] : 0u;
      if (pf != 0) {
        cg::coalesced_group g = cg::coalesced_threads();
        unsigned long long slot = 0;
        if (g.thread_rank() == 0) slot = atomicAdd(a.patched_count, (unsigned long long)g.size());
        a.patched_queue[g.shfl(slot, 0) + g.thread_rank()] = t;
      } else {
        uint32_t rows[RV_POS_U32];
        const uint4* r4 = (const uint4*)(a.counts + (size_t)t * RV_POS_U32);
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint4 v = r4[k];
          rows[4 * k] = v.x; rows[4 * k + 1] = v.y; rows[4 * k + 2] = v.z; rows[4 * k + 3] = v.w;
          any |= v.x | v.y | v.z | v.w;
        }
        if (any) {
          char refb = 0;
          if (pos >= dr->r.ref_lo && pos <= dr->r.ref_hi && pos >= a.ref_start && (int64_t)(pos - a.ref_start) < a.ref_n)
            refb = a.ref[pos - a.ref_start];
          DenseEmit em;
          em.a = &a;
          em.base.a = &a;
          em.work = s_work;
          em.n_work = &s_nwork;
          em.pos = pos;
          em.region = ri;
          score_dense_position(a.P, ri, pos, refb, rows, a.cov[t], em);
        }
      }
    }
  }
  __syncthreads();
  // ---- the open part of the non-reference records, one work item per lane ----------------------------
  const int n_work = s_nwork;
  LgTable L;
  L.t = a.lgt;
  L.n = a.lgt_n;
  for (int w = threadIdx.x; w < n_work; w += SCORE_BLOCK) {
    const ScoreWork k = s_work[w];
    const rv_region& R = a.regions[k.region].r;
    double msi, pvalue, oddratio;
    int shift3, msint;
    // the spans [pos-30, pos+70] of findMSI: unchecked loads when they lie inside the loaded window
    const int64_t w_lo = R.ref_lo > a.ref_start ? R.ref_lo : a.ref_start;
    const int64_t w_hi = (int64_t)R.ref_hi < a.ref_start + a.ref_n - 1 ? (int64_t)R.ref_hi : a.ref_start + a.ref_n - 1;
    if (k.pos - 30 >= w_lo && k.pos + 70 <= w_hi) {
      RefRaw raw;
      raw.b = a.ref - a.ref_start;
      finish_dense_variant(a.P, k.pos, R.chr_len, raw, L, k.a11, k.a12, k.a21, k.a22, &msi, &shift3, &msint, &pvalue, &oddratio);
    } else {
      RefView ref;
      ref.bases = a.ref;
      ref.base_pos = a.ref_start;
      ref.n = a.ref_n;
      ref.lo = R.ref_lo;
      ref.hi = R.ref_hi;
      finish_dense_variant(a.P, k.pos, R.chr_len, ref, L, k.a11, k.a12, k.a21, k.a22, &msi, &shift3, &msint, &pvalue, &oddratio);
    }
    rv_variant* o = a.variants + k.slot;
    o->msi = msi;
    o->shift3 = shift3;
    o->msint = msint;
    o->pvalue = pvalue;
    o->oddratio = oddratio;
  }
}

// Candidate mode (rv_params.candidates_only, simple-mode and paired-mode output): integer work only.  A position can print something only if one of its non-reference alleles has hicnt >= minr (a
// necessary condition of Variant::isGoodVar, include/Variant.h:205-231); those positions, and the positions with
// patch entries, are queued for the general scoring kernel.  Everything else ends here after one pass over its row.
// Layout of the pass: a warp owns 128 consecutive table positions.
//   A. four positions per lane, two vector loads: patch flags and the pileup's `touched` bytes (some allele other than the
//      reference's was observed there).  A position with neither can print nothing and ends here: no region lookup, no
//      reference base, no coverage, and above all none of its 128 bytes of rows (6 % of the positions go on at 30x, 30 % at
//      300x).
//   B. the positions that go on, one per lane (32 per round): region, reference allele, coverage.  Patched ones are queued as
//      they are.
//   C. the others' rows, four positions per round, EIGHT LANES PER POSITION: lane s loads bytes [16 s, 16 s + 16) of the
//      position's 128-byte row block (whole lines per load instruction; one thread per position touched 32 lines per
//      instruction and ran at half the HBM rate), the eight 16-byte parts are combined with one shuffle pair and a ballot.
__global__ void __launch_bounds__(256) rv_score_screen_kernel(ScoreArgs a) {
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  const int64_t w_base = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * 128;  // the warp's first position
  if (w_base >= a.n_positions) return;  // (whole warps; the kernel has no CTA barrier)
  // ---- A
  unsigned m4[4];  // m4[j]: lanes whose position 4 * lane + j goes on
  {
    const int64_t t4 = w_base + 4 * lane;
    uint32_t tw = 0;
    uint4 pf = make_uint4(0, 0, 0, 0);
    if (t4 + 3 < a.n_positions) {
      tw = *(const uint32_t*)(a.touched + t4);
      if (a.patch_first) pf = *(const uint4*)(a.patch_first + t4);
    } else {
      if (t4 + 0 < a.n_positions) { tw |= (uint32_t)a.touched[t4 + 0]; if (a.patch_first) pf.x = a.patch_first[t4 + 0]; }
      if (t4 + 1 < a.n_positions) { tw |= (uint32_t)a.touched[t4 + 1] << 8; if (a.patch_first) pf.y = a.patch_first[t4 + 1]; }
      if (t4 + 2 < a.n_positions) { tw |= (uint32_t)a.touched[t4 + 2] << 16; if (a.patch_first) pf.z = a.patch_first[t4 + 2]; }
    }
    m4[0] = __ballot_sync(0xffffffffu, (tw & 0x000000ffu) || pf.x);
    m4[1] = __ballot_sync(0xffffffffu, (tw & 0x0000ff00u) || pf.y);
    m4[2] = __ballot_sync(0xffffffffu, (tw & 0x00ff0000u) || pf.z);
    m4[3] = __ballot_sync(0xffffffffu, (tw & 0xff000000u) || pf.w);
  }
  const int c0 = __popc(m4[0]), c1 = __popc(m4[1]), c2 = __popc(m4[2]), total = c0 + c1 + c2 + __popc(m4[3]);
  const int r_first = a.tabblk_region[w_base >> 8];  // region of a position at or before the warp's first
  for (int r0 = 0; r0 < total; r0 += 32) {
    // ---- B: this lane's position of the round
    int off = -1;  // position - w_base
    {
      int kk = r0 + lane;
      if (kk < total) {
        int j = 0;
        unsigned mk = m4[0];
        if (kk >= c0) { kk -= c0; j = 1; mk = m4[1];
          if (kk >= c1) { kk -= c1; j = 2; mk = m4[2];
            if (kk >= c2) { kk -= c2; j = 3; mk = m4[3]; } } }
        off = 4 * (int)__fns(mk, 0u, kk + 1) + j;
      }
    }
    // bits 0-2: reference allele + 1 (0 = none), 3: inside the region, 4: has patch entries, 5: coverage != 0, 6: touched
    uint32_t m = 0;
    const int64_t t = w_base + (off < 0 ? 0 : off);
    if (off >= 0) {
      const int ri = find_region_by_tab_near(a.regions, a.n_regions, t, r_first);
      const DevRegion* dr = a.regions + ri;
      const int pos = dr->first_pos + (int)(t - dr->tab_off);
      if (pos >= dr->r.start && pos <= dr->r.end) {
        m = 8u;
        if (pos >= dr->r.ref_lo && pos <= dr->r.ref_hi && pos >= a.ref_start && (int64_t)(pos - a.ref_start) < a.ref_n)
          m |= (uint32_t)(allele_of(a.ref[pos - a.ref_start]) + 1);
        if ((a.patch_first ? a.patch_first[t] : 0u) != 0) m |= 16u;
        if (a.cov[t] != 0) m |= 32u;
        if (a.touched[t]) m |= 64u;
      }
    }
    // positions with patch entries are queued as they are; rows are read where they can matter: inside the region, covered,
    // not queued already, touched
    unsigned qmask = __ballot_sync(0xffffffffu, (m & (8u | 16u)) == (8u | 16u));
    const unsigned need = __ballot_sync(0xffffffffu, (m & (8u | 16u | 32u | 64u)) == (8u | 32u | 64u));
    const int n_need = __popc(need);
    // ---- C
    for (int r = 0; 4 * r < n_need; ++r) {
      const int idx = 4 * r + grp;
      const bool valid = idx < n_need;
      const int k = valid ? (int)__fns(need, 0u, idx + 1) : 0;  // the lane (of B) whose position this group reads
      const uint32_t mk = __shfl_sync(0xffffffffu, m, k);
      const int offk = __shfl_sync(0xffffffffu, off, k);
      uint4 x = make_uint4(0, 0, 0, 0);
      if (valid) x = __ldg((const uint4*)(a.counts + (size_t)(w_base + offk) * RV_POS_U32) + sub);
      const uint32_t ex = x.x | x.y | x.z | x.w;
      // even lanes hold {fwd, rev, sum tp, sum q} of allele sub / 2, odd lanes {sum mapq, sum nm, hicnt, std}
      const uint32_t ex_o = __shfl_xor_sync(0xffffffffu, ex, 1);
      const uint32_t hi_o = __shfl_xor_sync(0xffffffffu, x.z, 1);
      const int refal = (int)(mk & 7u) - 1;
      const bool possible = valid && (sub & 1) == 0 && (sub >> 1) != refal && (ex | ex_o) != 0 && x.x + x.y != 0 && (int)hi_o >= a.P.minr;
      const unsigned b_pos = __ballot_sync(0xffffffffu, possible);
      // the group leaders' verdicts, as bits at their lanes of B
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int kg = __shfl_sync(0xffffffffu, k, 8 * g);
        if ((b_pos >> (8 * g)) & 0xffu) qmask |= 1u << kg;
      }
    }
    if ((qmask >> lane) & 1u) {
      cg::coalesced_group g = cg::coalesced_threads();
      unsigned long long slot = 0;
      if (g.thread_rank() == 0) slot = atomicAdd(a.patched_count, (unsigned long long)g.size());
      a.patched_queue[g.shfl(slot, 0) + g.thread_rank()] = t;
    }
  }
}

__global__ void __launch_bounds__(128) rv_score_patched_kernel(ScoreArgs a) {
  const unsigned long long n = *a.patched_count;
  for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < n;
       q += (unsigned long long)gridDim.x * blockDim.x) {
    const int64_t t = a.patched_queue[q];
    const int ri = find_region_by_tab(a.regions, a.n_regions, t);
    const DevRegion* dr = a.regions + ri;
    const int i = (int)(t - dr->tab_off);
    const int pos = dr->first_pos + i;
    const uint32_t* rows = a.counts + (size_t)t * RV_POS_U32;
    const uint32_t pf = a.patch_first ? a.patch_first[t] : 0u;
    const int pn = pf ? (int)a.patch_count[t] : 0;
    RefView ref;
    ref.bases = a.ref;
    ref.base_pos = a.ref_start;
    ref.n = a.ref_n;
    ref.lo = dr->r.ref_lo;
    ref.hi = dr->r.ref_hi;
    const bool has_next = i + 1 < dr->n_pos;
    const uint32_t pfn = (has_next && a.patch_first) ? a.patch_first[t + 1] : 0u;  // ToVarsBuilder.cpp:754-760
    const int pnn = pfn ? (int)a.patch_count[t + 1] : 0;
    LgTable L;
    L.t = a.lgt;
    L.n = a.lgt_n;
    DeviceEmit em;
    em.a = &a;
    int unsup = 0;
    score_position(a.P, dr->r, ri, pos, ref, rows, a.cov[t], has_next, rows + RV_POS_U32, has_next ? a.cov[t + 1] : 0u,
                   a.patch, pf ? (int)(pf - 1) : 0, pn, pfn ? (int)(pfn - 1) : 0, pnn, L, em, &unsup);
    if (unsup) atomicAdd(&a.stats->n_score_unsupported, (unsigned long long)unsup);
  }
}

// The general score_position for an explicit list of table positions (somatic mode: full records of both
// samples at the positions where either sample has a candidate).
__global__ void __launch_bounds__(128) rv_score_list_kernel(ScoreArgs a, const int64_t* list, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = list[q];
    if (t < 0) continue;
    const int ri = find_region_by_tab(a.regions, a.n_regions, t);
    const DevRegion* dr = a.regions + ri;
    const int i = (int)(t - dr->tab_off);
    const int pos = dr->first_pos + i;
    const uint32_t* rows = a.counts + (size_t)t * RV_POS_U32;
    const uint32_t pf = a.patch_first ? a.patch_first[t] : 0u;
    const int pn = pf ? (int)a.patch_count[t] : 0;
    RefView ref;
    ref.bases = a.ref;
    ref.base_pos = a.ref_start;
    ref.n = a.ref_n;
    ref.lo = dr->r.ref_lo;
    ref.hi = dr->r.ref_hi;
    const bool has_next = i + 1 < dr->n_pos;
    const uint32_t pfn = (has_next && a.patch_first) ? a.patch_first[t + 1] : 0u;  // ToVarsBuilder.cpp:754-760
    const int pnn = pfn ? (int)a.patch_count[t + 1] : 0;
    LgTable L;
    L.t = a.lgt;
    L.n = a.lgt_n;
    DeviceEmit em;
    em.a = &a;
    int unsup = 0;
    score_position(a.P, dr->r, ri, pos, ref, rows, a.cov[t], has_next, rows + RV_POS_U32, has_next ? a.cov[t + 1] : 0u,
                   a.patch, pf ? (int)(pf - 1) : 0, pn, pfn ? (int)(pfn - 1) : 0, pnn, L, em, &unsup);
    if (unsup) atomicAdd(&a.stats->n_score_unsupported, (unsigned long long)unsup);
  }
}

// Per-region coverage summary (add_depth_by_region, somaticMode.cpp:69-81): sum and number of covered positions
// over [start, end) — the end is exclusive there.  One CTA per region.
__global__ void rv_cov_summary_kernel(const DevRegion* regions, int n_regions, const uint32_t* cov, unsigned long long* out) {
  const int r = blockIdx.x;
  if (r >= n_regions) return;
  const DevRegion* dr = regions + r;
  unsigned long long sum = 0, cnt = 0;
  for (int p = dr->r.start + (int)threadIdx.x; p < dr->r.end; p += (int)blockDim.x) {
    const uint32_t c = cov[dr->tab_off + (p - dr->first_pos)];
    if (c) { sum += c; cnt++; }
  }
  __shared__ unsigned long long sh[2];
  if (threadIdx.x < 2) sh[threadIdx.x] = 0;
  __syncthreads();
  for (int off = 16; off > 0; off >>= 1) {
    sum += __shfl_down_sync(0xffffffffu, sum, off);
    cnt += __shfl_down_sync(0xffffffffu, cnt, off);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sh[0], sum); atomicAdd(&sh[1], cnt); }
  __syncthreads();
  if (threadIdx.x == 0) { out[2 * r] = sh[0]; out[2 * r + 1] = sh[1]; }
}

__global__ void rv_lgamma_table_kernel(double* t, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = lgamma((double)i + 1.0);
}

__global__ void rv_fisher_kernel(const int32_t* tables, int64_t n, double* out, const double* lgt, int lgt_n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  LgTable L;
  L.t = lgt;
  L.n = lgt_n;
  double l, r, two;
  fisher_exact(L, tables[4 * i], tables[4 * i + 1], tables[4 * i + 2], tables[4 * i + 3], &l, &r, &two);
  out[3 * i] = l;
  out[3 * i + 1] = r;
  out[3 * i + 2] = two;
}

// gather of selected dense rows (33 u32 per position) into a compact buffer
__global__ void rv_gather_rows_kernel(const int64_t* tab, int64_t n, const uint32_t* counts, const uint32_t* cov,
                                      uint32_t* out) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 33;
  int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) % 33);
  if (i >= n) return;
  int64_t t = tab[i];
  uint32_t v = 0;
  if (t >= 0) v = w < RV_POS_U32 ? counts[(size_t)t * RV_POS_U32 + w] : cov[t];
  out[(size_t)i * 33 + w] = v;
}

// scatter of patch groups / coverage overrides / dense-row overrides
__global__ void rv_patch_scatter_kernel(const int64_t* grp_tab, const uint32_t* grp_first, const uint8_t* grp_n, int64_t n,
                                        uint32_t* patch_first, uint8_t* patch_count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  patch_first[grp_tab[i]] = grp_first[i] + 1u;
  patch_count[grp_tab[i]] = grp_n[i];
}
__global__ void rv_cov_scatter_kernel(const int64_t* tab, const int32_t* val, int64_t n, uint32_t* cov) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cov[tab[i]] = (uint32_t)val[i];
}

// ------------------------------------------------------------------------------------------------
// host side of the ABI
// ------------------------------------------------------------------------------------------------
struct rv_ctx {
  int device;
  rv_params P;
  rv_limits L;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1, tev0, tev1, pev0, pev1;  // ev: scoring calls, pev: rv_pileup, tev: rv_timer_*
  cudaEvent_t evs[3];      // boundaries between the kernels of the pileup stage
  float split_ms[4];       // classify, tile index + gather, walk, apply
  std::string err;
  int64_t launches;
  // device buffers
  rv_read* d_reads;
  uint8_t* d_pool;
  char* d_ref;
  int32_t ref_start;
  int64_t ref_n;
  uint32_t* d_counts;
  uint32_t* d_cov;
  int32_t* d_tabblk_region;   // region of table position 256 b (b = block): the kernels start from it instead of searching thousands of regions
  int32_t* d_itemblk_region;  // region of work item 128 b
  std::vector<int32_t> h_blk_region;
  uint8_t* d_touched;  // per table position: a row of an allele other than the reference's has received something (rv_score_screen_kernel reads rows only there)
  rv_event* d_events;
  rv_variant* d_variants;
  rv_patch_entry* d_patch;
  uint32_t* d_patch_first;
  uint8_t* d_patch_count;
  DevRegion* d_regions;
  int32_t* d_max_rl;
  DevStats* d_stats;
  double* d_lgt;
  int lgt_n;
  GDesc* d_descs;
  uint16_t* d_desc_mm;
  int32_t* d_reach;
  uint32_t* d_ref4;
  int64_t* d_tile_range;
  int64_t tile_cap;
  int64_t* d_patched_queue;
  unsigned long long* d_patched_count;
  unsigned long long* d_walk_queue;
  unsigned long long* d_walk_queue2;  // grouped by class (rv_walk_sort_kernel)
  bool walk_sort;
  unsigned long long* d_walk_count;
  int64_t n_tiles;
  bool use_gather;  // false (RV_NO_GATHER=1, a debugging reference): no descriptors, every base through the SparseObs list
  int tile;         // table positions per gather tile
  int g4_run, g4_alt, g4_variant, walk_occ;  // launch shapes (environment overrides are read once, at rv_create)
  GDesc* d_descs2;
  uint16_t* d_desc_mm2;
  uint4* d_desc_mml2;
  SparseObs* d_sparse;
  unsigned long long* d_sparse_count;
  int64_t max_sparse;
  int n_sms;        // SMs of the device (persistent grids)
  uint4* d_desc_mml;
  int64_t pool_dev_bytes;
  // batch state
  const rv_read* reads_dev_view;  // d_reads or a caller-provided device pointer
  const uint8_t* pool_dev_view;
  int64_t n_reads;
  int64_t read_origin;  // first batch read index that is resident (rv_push_reads_range)
  struct Slice { int64_t read_lo, read_hi, read_bias, pool_bias; };
  std::vector<Slice> slices;  // uploaded slices of the host batch (one for rv_push_reads / rv_push_reads_device)
  std::vector<DevRegion> regions;
  int64_t n_positions, n_items;
  bool have_patch;
  // host mirrors (pinned)
  uint32_t* h_counts;
  uint32_t* h_cov;
  size_t h_tab_cap;
  bool tables_fetched;
  rv_event* h_events;
  size_t h_events_cap;
  uint32_t* h_rows;
  size_t h_rows_cap;
  uint8_t* d_scratch;   // persistent device scratch (no cudaMalloc/cudaFree on the per-batch path: both
  size_t scratch_cap;   // stall for ~100 ms once GBs of host memory are page-locked)
  rv_variant* h_variants;
  size_t h_variants_cap;
  int32_t* h_max_rl;
  DevStats h_stats;
  uint8_t* d_arena;  // the one device allocation the buffers above are carved from
  uint32_t* d_first_read;  // -T only: [position][allele] lowest read index that added to the row (allocated on first use)
  bool lazy;     // rv_set_lazy: rv_pileup / rv_score enqueue only; results are settled at rv_sync or by the first getter
  bool unsettled;
  bool unsettled_pileup;
  std::vector<int64_t> keep_list;  // rv_score_positions' position list while its upload may still be in flight
  float pileup_ms, score_ms;
};

static int fail(rv_ctx* c, int code, const std::string& msg);
static int ensure_scratch(rv_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->scratch_cap) return RV_OK;
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  ctx->d_scratch = NULL;
  ctx->scratch_cap = 0;
  size_t cap = bytes + bytes / 4 + (1 << 20);
  if (cudaMalloc(&ctx->d_scratch, cap) != cudaSuccess) return fail(ctx, RV_ERR_NOMEM, "out of device memory (scratch)");
  ctx->scratch_cap = cap;
  return RV_OK;
}
static int fail(rv_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ctx, RV_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
  } while (0)

// Lazy mode: brings the host's view (statistics, maxReadLength, kernel times, variant count) up to date with what has
// been enqueued.  Returns the error a non-lazy rv_pileup / rv_score would have returned.
static int settle(rv_ctx* ctx) {
  if (!ctx->unsettled) return RV_OK;
  ctx->unsettled = false;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(&ctx->h_stats, ctx->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, ctx->stream));
  if (ctx->unsettled_pileup)
    CK(cudaMemcpyAsync(ctx->h_max_rl, ctx->d_max_rl, sizeof(int32_t) * ctx->regions.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->h_stats.n_items = (unsigned long long)ctx->n_items;
  if (ctx->unsettled_pileup) {
    ctx->unsettled_pileup = false;
    CK(cudaEventElapsedTime(&ctx->pileup_ms, ctx->pev0, ctx->pev1));
    CK(cudaEventElapsedTime(&ctx->split_ms[0], ctx->pev0, ctx->evs[0]));
    CK(cudaEventElapsedTime(&ctx->split_ms[2], ctx->evs[0], ctx->evs[1]));
    CK(cudaEventElapsedTime(&ctx->split_ms[1], ctx->evs[1], ctx->evs[2]));
    CK(cudaEventElapsedTime(&ctx->split_ms[3], ctx->evs[2], ctx->pev1));
  }
  CK(cudaEventElapsedTime(&ctx->score_ms, ctx->ev0, ctx->ev1));
  if (ctx->h_stats.n_overflow)
    return fail(ctx, RV_ERR_OVERFLOW, "pileup dropped events or sparse observations (limits.max_events / limits.max_reads too small)");
  if (ctx->h_stats.n_variants > (unsigned long long)ctx->L.max_variants)
    return fail(ctx, RV_ERR_OVERFLOW, "more variants than limits.max_variants");
  return RV_OK;
}

extern "C" {

int rv_abi_version(void) { return RV_ABI_VERSION; }

int rv_set_lazy(rv_ctx* ctx, int on) {
  if (!ctx) return RV_ERR_ARG;
  const int rc = settle(ctx);
  ctx->lazy = on != 0;
  return rc;
}

int rv_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int rv_warmup(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return RV_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) return RV_ERR_CUDA;
  // With lazy module loading (the CUDA 12 default) a kernel's image is loaded at its first launch — under a lock, in the
  // middle of the first jobs.  Asking for the attributes of the kernels a run uses loads them here, on the caller's
  // warm-up thread, beside the decode threads.
  cudaFuncAttributes fa;
  const void* used[] = {(const void*)rv_pack_ref_kernel, (const void*)rv_pileup_kernel, (const void*)rv_walk_sort_kernel,
                        (const void*)rv_walk_kernel<4>, (const void*)rv_tile_index_kernel, (const void*)rv_gather4_kernel<4, 7>,
                        (const void*)rv_apply_kernel, (const void*)rv_score_screen_kernel, (const void*)rv_score_patched_kernel,
                        (const void*)rv_score_list_kernel, (const void*)rv_lgamma_table_kernel, (const void*)rv_gather_rows_kernel,
                        (const void*)rv_patch_scatter_kernel, (const void*)rv_cov_scatter_kernel, (const void*)rv_cov_summary_kernel};
  for (size_t k = 0; k < sizeof(used) / sizeof(used[0]); ++k)
    if (cudaFuncGetAttributes(&fa, used[k]) != cudaSuccess) { cudaGetLastError(); return RV_ERR_CUDA; }
  return RV_OK;
}

void rv_default_params(rv_params* p) {
  memset(p, 0, sizeof(*p));
  p->goodq = 22.5;
  p->freq = 0.01;
  p->lofreq = 0.05;
  p->qratio = 1.5;
  p->mapq = 0;
  p->bias = 0.05;
  p->vext = 2;
  p->mismatch = 8;
  p->minr = 2;
  p->min_bias_reads = 2;
  p->read_pos_filter = 5;
  p->minmatch = 0;
  p->trim_bases_after = 0;
  p->indelsize = 50;
  p->mapping_quality = 0;
  p->samfilter = 0x504;
  p->local_realign = 1;
}

void rv_default_limits(rv_limits* l) {
  memset(l, 0, sizeof(*l));
  l->max_reads = 1 << 20;
  l->max_read_bytes = (int64_t)(1 << 20) * 256;
  l->max_positions = 4 << 20;
  l->max_regions = 8192;
  l->halo = 512;
  l->max_events = 1 << 20;
  l->max_variants = 1 << 20;
  l->max_patch = 1 << 20;
  l->max_ref_bases = 64 << 20;
}

const char* rv_last_error(const rv_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int rv_create(rv_ctx** out, int device, const rv_params* params, const rv_limits* limits) {
  if (!out || !params || !limits) return RV_ERR_ARG;
  *out = NULL;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return RV_ERR_CUDA;
  rv_ctx* ctx = new rv_ctx();
  memset((void*)&ctx->h_stats, 0, sizeof(ctx->h_stats));
  ctx->lazy = false;
  ctx->unsettled = false;
  ctx->unsettled_pileup = false;
  ctx->device = device;
  ctx->P = *params;
  ctx->L = *limits;
  ctx->launches = 0;
  ctx->d_reads = NULL; ctx->d_pool = NULL; ctx->d_ref = NULL; ctx->d_counts = NULL; ctx->d_cov = NULL;
  ctx->d_events = NULL; ctx->d_variants = NULL; ctx->d_patch = NULL; ctx->d_patch_first = NULL;
  ctx->d_patch_count = NULL; ctx->d_regions = NULL; ctx->d_max_rl = NULL; ctx->d_stats = NULL; ctx->d_lgt = NULL; ctx->d_descs = NULL; ctx->d_desc_mm = NULL; ctx->d_reach = NULL; ctx->d_ref4 = NULL; ctx->d_tile_range = NULL; ctx->tile_cap = 0; ctx->d_patched_queue = NULL; ctx->d_patched_count = NULL; ctx->d_walk_queue = NULL; ctx->d_walk_count = NULL; ctx->n_tiles = 0;
  ctx->use_gather = getenv("RV_NO_GATHER") == NULL;
  ctx->tile = G4_W;
  ctx->g4_run = getenv("RV_G4_RUN") ? std::max(1, atoi(getenv("RV_G4_RUN"))) : 1;
  ctx->g4_alt = getenv("RV_G4_ALT") ? atoi(getenv("RV_G4_ALT")) : 1;
  ctx->g4_variant = getenv("RV_G4_VARIANT") ? atoi(getenv("RV_G4_VARIANT")) : 4;  // <4, 7>: 72 registers, fewest spills
  ctx->walk_occ = getenv("RV_WALK_OCC") ? atoi(getenv("RV_WALK_OCC")) : 4;
  ctx->walk_sort = getenv("RV_NO_WALK_SORT") == NULL;
  ctx->d_descs2 = NULL; ctx->d_desc_mm2 = NULL; ctx->d_desc_mml2 = NULL; ctx->d_sparse = NULL; ctx->d_sparse_count = NULL;
  ctx->max_sparse = 0;

  ctx->d_desc_mml = NULL;
  ctx->d_arena = NULL;
  ctx->d_first_read = NULL;
  ctx->pool_dev_bytes = 0;
  ctx->h_counts = NULL; ctx->h_cov = NULL; ctx->h_tab_cap = 0; ctx->h_events = NULL; ctx->h_events_cap = 0;
  ctx->h_variants = NULL; ctx->h_variants_cap = 0; ctx->h_max_rl = NULL; ctx->h_rows = NULL; ctx->h_rows_cap = 0; ctx->d_scratch = NULL; ctx->scratch_cap = 0;
  ctx->n_reads = 0; ctx->read_origin = 0; ctx->n_positions = 0; ctx->n_items = 0; ctx->have_patch = false; ctx->tables_fetched = false;
  ctx->ref_start = 1; ctx->ref_n = 0; ctx->pileup_ms = ctx->score_ms = 0;
  ctx->evs[0] = ctx->evs[1] = ctx->evs[2] = NULL; ctx->split_ms[0] = ctx->split_ms[1] = ctx->split_ms[2] = ctx->split_ms[3] = 0;
  ctx->reads_dev_view = NULL; ctx->pool_dev_view = NULL;
  *out = ctx;  // returned even on failure so the caller can read rv_last_error, then rv_destroy
  {
    // the byte-SIMD quality threshold of rv_gather4_kernel and its 32-bit pool offsets
    const int thr = (int)ceil(params->goodq);
    if (!(thr >= 1 && thr <= 128)) return fail(ctx, RV_ERR_ARG, "phred threshold (-q) outside (0, 128] is not supported");
    if (limits->max_read_bytes >= ((int64_t)1 << 32) - 64) return fail(ctx, RV_ERR_ARG, "limits.max_read_bytes must stay below 4 GiB per context");
    if (limits->max_positions >= ((int64_t)1 << 32) - 2) return fail(ctx, RV_ERR_ARG, "limits.max_positions must stay below 2^32 per context");
  }
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->n_sms = 148;
  CK(cudaDeviceGetAttribute(&ctx->n_sms, cudaDevAttrMultiProcessorCount, ctx->device));
  CK(cudaEventCreate(&ctx->ev0));
  CK(cudaEventCreate(&ctx->ev1));
  CK(cudaEventCreate(&ctx->tev0));
  CK(cudaEventCreate(&ctx->tev1));
  CK(cudaEventCreate(&ctx->pev0));
  CK(cudaEventCreate(&ctx->pev1));
  for (int k = 0; k < 3; ++k) CK(cudaEventCreate(&ctx->evs[k]));
  const rv_limits& L = ctx->L;
  // Every device buffer of the context is carved out of ONE allocation: cudaMalloc takes a device-wide lock, and the
  // ~30 calls a context used to make cost 100-400 ms of a short process's life when four worker threads created
  // their contexts side by side (measured through the drop-in CLI on BASELINE configs[1]).
  ctx->max_sparse = L.max_sparse_obs > 0 ? L.max_sparse_obs : std::max<int64_t>(4 << 20, 8 * L.max_reads);
  if (ctx->max_sparse > ((int64_t)1 << 31)) return fail(ctx, RV_ERR_ARG, "limits.max_sparse_obs too large");
  ctx->tile_cap = L.max_positions / ctx->tile + L.max_regions + 1;  // every region rounds its table up to whole tiles
  ctx->lgt_n = 1 << 20;
  const size_t n_items_cap = (size_t)(2 * L.max_reads + 1024);  // work items are (region, read) pairs
  {
    size_t total = 0;
    uint8_t* base = NULL;
    for (int pass = 0; pass < 2; ++pass) {
      size_t off = 0;
      auto carve = [&](auto** p, size_t bytes) {
        if (pass == 1) *p = reinterpret_cast<std::remove_reference_t<decltype(**p)>*>(base + off);
        off += (bytes + 255) & ~(size_t)255;
      };
      carve(&ctx->d_reads, sizeof(rv_read) * (size_t)L.max_reads);
      carve(&ctx->d_pool, (size_t)L.max_read_bytes + 64);  // the gather kernel's look-ahead loads
      carve(&ctx->d_ref, (size_t)L.max_ref_bases);
      carve(&ctx->d_counts, sizeof(uint32_t) * RV_POS_U32 * (size_t)(L.max_positions + 1));
      carve(&ctx->d_cov, sizeof(uint32_t) * (size_t)(L.max_positions + 1));
      carve(&ctx->d_touched, (size_t)(L.max_positions + 4));
      carve(&ctx->d_tabblk_region, sizeof(int32_t) * (size_t)(L.max_positions / 256 + 4));
      carve(&ctx->d_itemblk_region, sizeof(int32_t) * (size_t)(n_items_cap / 128 + 4));
      carve(&ctx->d_events, sizeof(rv_event) * (size_t)L.max_events);
      carve(&ctx->d_variants, sizeof(rv_variant) * (size_t)L.max_variants);
      carve(&ctx->d_patch, sizeof(rv_patch_entry) * (size_t)L.max_patch);
      carve(&ctx->d_patch_first, sizeof(uint32_t) * (size_t)(L.max_positions + 1));
      carve(&ctx->d_patch_count, (size_t)(L.max_positions + 1));
      carve(&ctx->d_regions, sizeof(DevRegion) * (size_t)L.max_regions);
      carve(&ctx->d_max_rl, sizeof(int32_t) * (size_t)L.max_regions);
      // one block: statistics | walk queue counters [3] | SparseObs cursor | reach bounds [2] (a single memset per pileup)
      uint8_t* blk = NULL;
      carve(&blk, COUNTER_BLOCK_BYTES);
      if (pass == 1) {
        ctx->d_stats = (DevStats*)blk;
        ctx->d_walk_count = (unsigned long long*)(blk + 128);
        ctx->d_sparse_count = (unsigned long long*)(blk + 128 + 24);
        ctx->d_reach = (int32_t*)(blk + 128 + 32);
      }
      carve(&ctx->d_descs, sizeof(GDesc) * n_items_cap);
      carve(&ctx->d_desc_mm, sizeof(uint16_t) * n_items_cap);
      carve(&ctx->d_desc_mml, sizeof(uint4) * n_items_cap);
      carve(&ctx->d_descs2, sizeof(GDesc) * n_items_cap);
      carve(&ctx->d_desc_mm2, sizeof(uint16_t) * n_items_cap);
      carve(&ctx->d_desc_mml2, sizeof(uint4) * n_items_cap);
      // SparseObs list: a walked read leaves a few entries (soft-clip re-extension, coverage under a deletion), a stretch
      // that is not plain one per base; RV_NO_GATHER sends every base here (small debugging batches only)
      carve(&ctx->d_sparse, sizeof(SparseObs) * (size_t)ctx->max_sparse);
      carve(&ctx->d_ref4, sizeof(uint32_t) * (size_t)(L.max_ref_bases / 8 + 16));
      carve(&ctx->d_tile_range, 2 * sizeof(int64_t) * (size_t)ctx->tile_cap);
      // positions queued for the general scoring kernel (patched positions, screened candidates): at most every position
      carve(&ctx->d_patched_queue, sizeof(int64_t) * (size_t)(L.max_positions + 1));
      carve(&ctx->d_patched_count, sizeof(unsigned long long));
      carve(&ctx->d_walk_queue, sizeof(unsigned long long) * n_items_cap);
      carve(&ctx->d_walk_queue2, sizeof(unsigned long long) * n_items_cap);
      carve(&ctx->d_lgt, sizeof(double) * (size_t)ctx->lgt_n);
      if (pass == 0) {
        total = off;
        if (cudaMalloc(&base, total) != cudaSuccess) {
          cudaGetLastError();
          return fail(ctx, RV_ERR_NOMEM, "out of device memory: the context's limits need " + std::to_string(total >> 20) + " MiB");
        }
        ctx->d_arena = base;
      }
    }
  }
  CK(cudaMemsetAsync(ctx->d_stats, 0, sizeof(DevStats), ctx->stream));
  rv_lgamma_table_kernel<<<(ctx->lgt_n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_lgt, ctx->lgt_n);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaMallocHost(&ctx->h_max_rl, sizeof(int32_t) * (size_t)L.max_regions));
  CK(cudaStreamSynchronize(ctx->stream));
  return RV_OK;
}

void rv_destroy(rv_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaFree(ctx->d_arena);  // every device buffer but the scratch
  cudaFree(ctx->d_first_read);
  cudaFree(ctx->d_scratch);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  if (ctx->h_cov) cudaFreeHost(ctx->h_cov);
  if (ctx->h_events) cudaFreeHost(ctx->h_events);
  if (ctx->h_rows) cudaFreeHost(ctx->h_rows);
  if (ctx->h_variants) cudaFreeHost(ctx->h_variants);
  if (ctx->h_max_rl) cudaFreeHost(ctx->h_max_rl);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->tev0) cudaEventDestroy(ctx->tev0);
  if (ctx->tev1) cudaEventDestroy(ctx->tev1);
  if (ctx->pev0) cudaEventDestroy(ctx->pev0);
  if (ctx->pev1) cudaEventDestroy(ctx->pev1);
  for (int k = 0; k < 3; ++k) if (ctx->evs[k]) cudaEventDestroy(ctx->evs[k]);
  delete ctx;
}

int32_t rv_ctx_halo(const rv_ctx* ctx) { return ctx ? ctx->L.halo : 0; }

int rv_set_params(rv_ctx* ctx, const rv_params* params) {
  if (!ctx || !params) return RV_ERR_ARG;
  if (!(ceil(params->goodq) >= 1 && ceil(params->goodq) <= 128))
    return fail(ctx, RV_ERR_ARG, "phred threshold (-q) outside (0, 128] is not supported");
  ctx->P = *params;
  return RV_OK;
}

int rv_sync(rv_ctx* ctx) {
  if (!ctx) return RV_ERR_ARG;
  { const int rcs = settle(ctx); if (rcs != RV_OK) return rcs; }
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return RV_OK;
}

int rv_set_reference(rv_ctx* ctx, int32_t ref_start, int64_t n, const char* bases) {
  if (!ctx || !bases || n < 0) return RV_ERR_ARG;
  if (n > ctx->L.max_ref_bases) return fail(ctx, RV_ERR_OVERFLOW, "reference larger than limits.max_ref_bases");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(ctx->d_ref, bases, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  ctx->ref_start = ref_start;
  ctx->ref_n = n;
  const int64_t n_words = n / 8 + 16;
  rv_pack_ref_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_ref, n, ctx->d_ref4, n_words);
  ctx->launches++;
  CK(cudaGetLastError());
  return RV_OK;
}

int rv_push_reads(rv_ctx* ctx, const rv_read_batch* b) {
  if (!ctx || !b || b->n_reads < 0) return RV_ERR_ARG;
  if (b->n_reads > ctx->L.max_reads) return fail(ctx, RV_ERR_OVERFLOW, "batch has more reads than limits.max_reads");
  if (b->pool_bytes > ctx->L.max_read_bytes) return fail(ctx, RV_ERR_OVERFLOW, "batch pool larger than limits.max_read_bytes");
  CK(cudaSetDevice(ctx->device));
  if (b->n_reads) CK(cudaMemcpyAsync(ctx->d_reads, b->reads, sizeof(rv_read) * (size_t)b->n_reads, cudaMemcpyHostToDevice, ctx->stream));
  if (b->pool_bytes) CK(cudaMemcpyAsync(ctx->d_pool, b->pool, (size_t)b->pool_bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->reads_dev_view = ctx->d_reads;
  ctx->pool_dev_view = ctx->d_pool;
  ctx->pool_dev_bytes = b->pool_bytes;
  ctx->n_reads = b->n_reads;
  ctx->read_origin = 0;
  ctx->slices.assign(1, rv_ctx::Slice{0, b->n_reads, 0, 0});
  return RV_OK;
}

int rv_push_reads_ranges(rv_ctx* ctx, const rv_read_batch* b, int32_t n_ranges, const int64_t* read_lo, const int64_t* read_hi) {
  if (!ctx || !b || n_ranges < 0 || (n_ranges && (!read_lo || !read_hi))) return RV_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  ctx->slices.clear();
  int64_t dev_reads = 0, dev_pool = 0, hi_max = 0;
  for (int k = 0; k < n_ranges; ++k) {
    const int64_t lo = read_lo[k], hi = read_hi[k];
    if (lo < 0 || hi < lo || hi > b->n_reads) return fail(ctx, RV_ERR_ARG, "bad read range");
    const int64_t n = hi - lo;
    if (n == 0) continue;
    const int64_t pool_lo = (int64_t)b->reads[lo].data_off16 * 16;
    const int64_t pool_hi = hi < b->n_reads ? (int64_t)b->reads[hi].data_off16 * 16 : ((b->pool_bytes + 15) & ~(int64_t)15);
    if (pool_hi < pool_lo) return fail(ctx, RV_ERR_ARG, "read headers do not index the pool in order");
    if (dev_reads + n > ctx->L.max_reads) return fail(ctx, RV_ERR_OVERFLOW, "read ranges larger than limits.max_reads");
    if (dev_pool + (pool_hi - pool_lo) > ctx->L.max_read_bytes + 16)
      return fail(ctx, RV_ERR_OVERFLOW, "read ranges' pool larger than limits.max_read_bytes");
    const int64_t copy_hi = pool_hi < b->pool_bytes ? pool_hi : b->pool_bytes;
    CK(cudaMemcpyAsync(ctx->d_reads + dev_reads, b->reads + lo, sizeof(rv_read) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_pool + dev_pool, b->pool + pool_lo, (size_t)(copy_hi - pool_lo), cudaMemcpyHostToDevice, ctx->stream));
    ctx->slices.push_back(rv_ctx::Slice{lo, hi, lo - dev_reads, pool_lo - dev_pool});
    dev_reads += n;
    dev_pool += pool_hi - pool_lo;  // a multiple of 16: every read starts on a 16-byte boundary
    if (hi > hi_max) hi_max = hi;
  }
  ctx->reads_dev_view = ctx->d_reads;
  ctx->pool_dev_view = ctx->d_pool;
  ctx->pool_dev_bytes = dev_pool;
  ctx->n_reads = hi_max;
  ctx->read_origin = 0;
  return RV_OK;
}

int rv_push_reads_range(rv_ctx* ctx, const rv_read_batch* b, int64_t read_lo, int64_t read_hi) {
  return rv_push_reads_ranges(ctx, b, 1, &read_lo, &read_hi);
}

int rv_push_reads_device(rv_ctx* ctx, const rv_read_batch* b) {
  if (!ctx || !b || b->n_reads < 0) return RV_ERR_ARG;
  ctx->reads_dev_view = b->reads;
  ctx->pool_dev_view = b->pool;
  ctx->pool_dev_bytes = b->pool_bytes;
  ctx->n_reads = b->n_reads;
  ctx->read_origin = 0;
  ctx->slices.assign(1, rv_ctx::Slice{0, b->n_reads, 0, 0});
  return RV_OK;
}

int rv_set_regions(rv_ctx* ctx, const rv_region* regs, int32_t n) {
  if (!ctx || (!regs && n) || n < 0) return RV_ERR_ARG;
  if (n > ctx->L.max_regions) return fail(ctx, RV_ERR_OVERFLOW, "more regions than limits.max_regions");
  CK(cudaSetDevice(ctx->device));
  ctx->regions.resize(n);
  int64_t tab = 0, items = 0, tiles = 0;
  for (int i = 0; i < n; ++i) {
    DevRegion& d = ctx->regions[i];
    d.r = regs[i];
    if (d.r.end < d.r.start || d.r.read_hi < d.r.read_lo) return fail(ctx, RV_ERR_ARG, "bad region " + std::to_string(i));
    d.read_bias = 0;
    d.pool_bias = 0;
    if (d.r.read_hi > d.r.read_lo) {
      bool found = false;
      for (size_t k = 0; k < ctx->slices.size() && !found; ++k)
        if (d.r.read_lo >= ctx->slices[k].read_lo && d.r.read_hi <= ctx->slices[k].read_hi) {
          d.read_bias = ctx->slices[k].read_bias;
          d.pool_bias = ctx->slices[k].pool_bias;
          found = true;
        }
      if (!found) return fail(ctx, RV_ERR_ARG, "region " + std::to_string(i) + " refers to reads that are not resident");
    }
    d.first_pos = d.r.start - ctx->L.halo;
    d.n_pos = d.r.end - d.r.start + 1 + 2 * ctx->L.halo;
    d.tab_off = tab;
    d.item_base = items;
    d.tile_base = tiles;
    tiles += (d.n_pos + ctx->tile - 1) / ctx->tile;
    tab += d.n_pos;
    items += d.r.read_hi - d.r.read_lo;
    ctx->h_max_rl[i] = d.r.max_read_len_in;
  }
  if (tab > ctx->L.max_positions) return fail(ctx, RV_ERR_OVERFLOW, "regions need more table positions than limits.max_positions");
  if (tiles > ctx->tile_cap) return fail(ctx, RV_ERR_OVERFLOW, "regions need more gather tiles than the context was sized for");
  if (items > 2 * ctx->L.max_reads + 1024)
    return fail(ctx, RV_ERR_OVERFLOW, "more (region, read) work items than 2 x limits.max_reads");
  ctx->n_positions = tab;
  ctx->n_items = items;
  ctx->n_tiles = tiles;
  ctx->have_patch = false;
  ctx->tables_fetched = false;
  if (n) {
    CK(cudaMemcpyAsync(ctx->d_regions, ctx->regions.data(), sizeof(DevRegion) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_max_rl, ctx->h_max_rl, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    {
      // region of the first position of every 256-position block / of the first item of every 128-item block: the last
      // region whose base is <= it (what the binary searches of the kernels return)
      const size_t nb_tab = (size_t)(tab / 256 + 1), nb_item = (size_t)(items / 128 + 1);
      std::vector<int32_t>& h = ctx->h_blk_region;
      h.resize(nb_tab + nb_item);
      int r = 0;
      for (size_t b = 0; b < nb_tab; ++b) {
        while (r + 1 < n && ctx->regions[(size_t)r + 1].tab_off <= (int64_t)b * 256) ++r;
        h[b] = r;
      }
      r = 0;
      for (size_t b = 0; b < nb_item; ++b) {
        while (r + 1 < n && ctx->regions[(size_t)r + 1].item_base <= (int64_t)b * 128) ++r;
        h[nb_tab + b] = r;
      }
      CK(cudaMemcpyAsync(ctx->d_tabblk_region, h.data(), sizeof(int32_t) * nb_tab, cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaMemcpyAsync(ctx->d_itemblk_region, h.data() + nb_tab, sizeof(int32_t) * nb_item, cudaMemcpyHostToDevice, ctx->stream));
    }
    // the async copies above read host vectors: make sure they are consumed before the caller can mutate them
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return RV_OK;
}

int rv_pileup(rv_ctx* ctx) {
  if (!ctx) return RV_ERR_ARG;
  if (ctx->regions.empty()) return fail(ctx, RV_ERR_STATE, "rv_set_regions has not been called");
  if (!ctx->reads_dev_view) return fail(ctx, RV_ERR_STATE, "rv_push_reads has not been called");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->lazy && ctx->unsettled) { const int rcs = settle(ctx); if (rcs != RV_OK) return rcs; }
  CK(cudaEventRecord(ctx->pev0, ctx->stream));
  // no table memset: rv_gather4_kernel stores every row (halo included) before rv_apply_kernel adds to them.
  // One memset clears the statistics, the walk queue counters, the SparseObs cursor and the reach bounds.
  CK(cudaMemsetAsync(ctx->d_stats, 0, COUNTER_BLOCK_BYTES, ctx->stream));
  PileupArgs a;
  a.P = ctx->P;
  a.regions = ctx->d_regions;
  a.itemblk_region = ctx->d_itemblk_region;
  a.n_regions = (int)ctx->regions.size();
  a.n_items = ctx->n_items;
  a.reads = ctx->reads_dev_view;
  a.pool = ctx->pool_dev_view;
  a.ref = ctx->d_ref;
  a.ref4 = ctx->d_ref4;
  a.ref_start = ctx->ref_start;
  a.ref_n = ctx->ref_n;
  a.counts = ctx->d_counts;
  a.cov = ctx->d_cov;
  a.events = ctx->d_events;
  a.max_events = (unsigned long long)ctx->L.max_events;
  a.max_rl = ctx->d_max_rl;
  a.stats = ctx->d_stats;
  a.descs = ctx->d_descs;
  a.desc_mm = ctx->d_desc_mm;
  a.desc_mml = ctx->d_desc_mml;
  a.descs2 = ctx->d_descs2;
  a.desc_mm2 = ctx->d_desc_mm2;
  a.desc_mml2 = ctx->d_desc_mml2;
  a.sparse = ctx->d_sparse;
  a.max_sparse = (unsigned long long)ctx->max_sparse;
  a.sparse_count = ctx->d_sparse_count;
  a.reach = ctx->d_reach;
  a.force_exact = ctx->use_gather ? 0 : 1;
  a.walk_queue = ctx->d_walk_queue;
  a.walk_queue2 = ctx->walk_sort ? ctx->d_walk_queue2 : ctx->d_walk_queue;
  a.walk_count = ctx->d_walk_count;
  a.walk_cap = (unsigned long long)(2 * ctx->L.max_reads + 1024);
  // 1. classify: filters, CIGAR rewrite, plain-run proof -> descriptors of the plain reads, queue of the others
  if (ctx->n_items > 0) {
    unsigned grid = (unsigned)((ctx->n_items + 127) / 128);
    rv_pileup_kernel<<<grid, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ctx->evs[0], ctx->stream));
  // 2. the queued reads: CIGAR walk -> descriptors of their plain segments, SparseObs list, events
  if (ctx->n_items > 0) {
    // the queue length is only known on the device: a fixed grid of grid-stride threads
    if (ctx->walk_sort) {
      rv_walk_sort_kernel<<<ctx->n_sms * 2, 256, 0, ctx->stream>>>(a);
      ctx->launches++;
    }
    const int occ = ctx->walk_occ;
    if (occ == 8) rv_walk_kernel<8><<<ctx->n_sms * 8, 128, 0, ctx->stream>>>(a);
    else if (occ == 6) rv_walk_kernel<6><<<ctx->n_sms * 6, 128, 0, ctx->stream>>>(a);
    else if (occ == 5) rv_walk_kernel<5><<<ctx->n_sms * 5, 128, 0, ctx->stream>>>(a);
    else rv_walk_kernel<4><<<ctx->n_sms * 4, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ctx->evs[1], ctx->stream));
  // 3. position-major accumulation of every descriptor; stores every table row
  if (ctx->n_tiles > 0) {
    GatherArgs g;
    g.goodq = ctx->P.goodq;
    g.regions = ctx->d_regions;
    g.n_regions = (int)ctx->regions.size();
    g.reads = ctx->reads_dev_view;
    g.descs = ctx->d_descs;
    g.desc_mm = ctx->d_desc_mm;
    g.pool = ctx->pool_dev_view;
    g.ref = ctx->d_ref;
    g.ref_start = ctx->ref_start;
    g.ref_n = ctx->ref_n;
    g.counts = ctx->d_counts;
    g.cov = ctx->d_cov;
    g.reach = ctx->d_reach;
    g.tile_range = ctx->d_tile_range;
    g.n_tiles = ctx->n_tiles;
    g.tile = ctx->tile;
    g.desc_mml = ctx->d_desc_mml;
    g.descs2 = ctx->d_descs2;
    g.desc_mm2 = ctx->d_desc_mm2;
    g.desc_mml2 = ctx->d_desc_mml2;
    g.pool_bytes = ctx->pool_dev_bytes;
    g.thr = (int)ceil(ctx->P.goodq);
    rv_tile_index_kernel<<<(unsigned)((ctx->n_tiles + 127) / 128), 128, 0, ctx->stream>>>(g);
    g.run = ctx->g4_run;
    g.alternate = ctx->g4_alt;
    g.touched = ctx->d_touched;
    const int64_t g4_warps = (ctx->n_tiles + g.run - 1) / g.run;
    const unsigned g4_grid = (unsigned)((g4_warps + G4_WARPS - 1) / G4_WARPS);
    const int variant = ctx->g4_variant;
    if (variant == 1) rv_gather4_kernel<8, 6><<<g4_grid, G4_WARPS * 32, 0, ctx->stream>>>(g);
    else if (variant == 2) rv_gather4_kernel<8, 5><<<g4_grid, G4_WARPS * 32, 0, ctx->stream>>>(g);
    else if (variant == 3) rv_gather4_kernel<4, 6><<<g4_grid, G4_WARPS * 32, 0, ctx->stream>>>(g);
    else if (variant == 4) rv_gather4_kernel<4, 7><<<g4_grid, G4_WARPS * 32, 0, ctx->stream>>>(g);
    ctx->launches += 2;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ctx->evs[2], ctx->stream));
  // 4. the SparseObs list onto the tables
  if (ctx->n_items > 0) {
    if (ctx->P.trim_bases_after == 0) {
      rv_apply_kernel<<<ctx->n_sms * 8, 256, 0, ctx->stream>>>(ctx->d_sparse, ctx->d_sparse_count, (unsigned long long)ctx->max_sparse,
                                                             ctx->d_counts, ctx->d_cov, ctx->d_touched, (int)ceil(ctx->P.goodq), ctx->d_stats,
                                                             (uint32_t*)0, 0);
      ctx->launches++;
    } else {
      // -T: first contributor of every row, then the conditional subtractions (two launches; the array exists only in this mode)
      const size_t need = sizeof(uint32_t) * 4 * (size_t)(ctx->L.max_positions + 1);
      if (!ctx->d_first_read) {
        if (cudaMalloc(&ctx->d_first_read, need) != cudaSuccess) { cudaGetLastError(); return fail(ctx, RV_ERR_NOMEM, "out of device memory (-T: first contributor per row)"); }
      }
      CK(cudaMemsetAsync(ctx->d_first_read, 0xff, sizeof(uint32_t) * 4 * (size_t)ctx->n_positions, ctx->stream));
      for (int pass = 1; pass <= 2; ++pass)
        rv_apply_kernel<<<ctx->n_sms * 8, 256, 0, ctx->stream>>>(ctx->d_sparse, ctx->d_sparse_count, (unsigned long long)ctx->max_sparse,
                                                               ctx->d_counts, ctx->d_cov, ctx->d_touched, (int)ceil(ctx->P.goodq), ctx->d_stats,
                                                               ctx->d_first_read, pass);
      ctx->launches += 2;
    }
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ctx->pev1, ctx->stream));
  ctx->tables_fetched = false;
  if (ctx->lazy) {
    ctx->unsettled = ctx->unsettled_pileup = true;
    return RV_OK;
  }
  CK(cudaMemcpyAsync(&ctx->h_stats, ctx->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_max_rl, ctx->d_max_rl, sizeof(int32_t) * ctx->regions.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaEventElapsedTime(&ctx->pileup_ms, ctx->pev0, ctx->pev1));
  CK(cudaEventElapsedTime(&ctx->split_ms[0], ctx->pev0, ctx->evs[0]));
  CK(cudaEventElapsedTime(&ctx->split_ms[2], ctx->evs[0], ctx->evs[1]));
  CK(cudaEventElapsedTime(&ctx->split_ms[1], ctx->evs[1], ctx->evs[2]));
  CK(cudaEventElapsedTime(&ctx->split_ms[3], ctx->evs[2], ctx->pev1));
  ctx->h_stats.n_items = (unsigned long long)ctx->n_items;
  // the patch list stays attached until rv_set_regions / the next rv_apply_patch: a caller that
  // re-runs the same resident batch may score against it again
  ctx->tables_fetched = false;
  if (ctx->h_stats.n_overflow)
    return fail(ctx, RV_ERR_OVERFLOW, "pileup dropped events or sparse observations (limits.max_events / limits.max_reads too small)");
  return RV_OK;
}

int rv_get_pileup_stats(rv_ctx* ctx, rv_pileup_stats* o) {
  if (!ctx || !o) return RV_ERR_ARG;
  { const int rcs = settle(ctx); if (rcs != RV_OK) return rcs; }
  o->n_items = (int64_t)ctx->h_stats.n_items;
  o->n_reads_kept = (int64_t)ctx->h_stats.n_kept;
  o->n_aligned_bases = (int64_t)ctx->h_stats.n_bases;
  o->n_events = (int64_t)ctx->h_stats.n_events;
  o->n_overflow = (int64_t)ctx->h_stats.n_overflow;
  o->n_unsupported = (int64_t)ctx->h_stats.n_unsupported;
  o->n_walk_items = (int64_t)ctx->h_stats.n_walk_items;
  o->n_walk_full = (int64_t)ctx->h_stats.n_walk_full;
  o->n_clipped = (int64_t)ctx->h_stats.n_clipped;
  o->n_score_unsupported = (int64_t)ctx->h_stats.n_score_unsupported;
  o->n_sparse_obs = (int64_t)ctx->h_stats.n_sparse;
  o->n_walk_segments = (int64_t)ctx->h_stats.n_segments;
  return RV_OK;
}

int rv_fetch_max_read_len(rv_ctx* ctx, const int32_t** out, int32_t* n) {
  if (!ctx || !out || !n) return RV_ERR_ARG;
  { const int rcs = settle(ctx); if (rcs != RV_OK) return rcs; }
  *out = ctx->h_max_rl;
  *n = (int32_t)ctx->regions.size();
  return RV_OK;
}

int rv_fetch_tables(rv_ctx* ctx, int32_t region, const uint32_t** counts, const uint32_t** cov, int32_t* first_pos,
                    int32_t* n_pos) {
  if (!ctx || region < 0 || region >= (int)ctx->regions.size()) return RV_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->tables_fetched) {
    size_t need = (size_t)ctx->n_positions;
    if (need > ctx->h_tab_cap) {
      if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
      if (ctx->h_cov) cudaFreeHost(ctx->h_cov);
      ctx->h_counts = NULL; ctx->h_cov = NULL;
      CK(cudaMallocHost(&ctx->h_counts, sizeof(uint32_t) * RV_POS_U32 * need));
      CK(cudaMallocHost(&ctx->h_cov, sizeof(uint32_t) * need));
      ctx->h_tab_cap = need;
    }
    CK(cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, sizeof(uint32_t) * RV_POS_U32 * need, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_cov, ctx->d_cov, sizeof(uint32_t) * need, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->tables_fetched = true;
  }
  const DevRegion& d = ctx->regions[region];
  if (counts) *counts = ctx->h_counts + (size_t)d.tab_off * RV_POS_U32;
  if (cov) *cov = ctx->h_cov + d.tab_off;
  if (first_pos) *first_pos = d.first_pos;
  if (n_pos) *n_pos = d.n_pos;
  return RV_OK;
}

int rv_fetch_rows(rv_ctx* ctx, const int32_t* region, const int32_t* pos, int64_t n, const uint32_t** rows) {
  if (!ctx || !rows || n < 0 || (n && (!region || !pos))) return RV_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  *rows = NULL;
  if (n == 0) return RV_OK;
  std::vector<int64_t> tab((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    int r = region[i];
    if (r < 0 || r >= (int)ctx->regions.size()) return fail(ctx, RV_ERR_ARG, "rv_fetch_rows: bad region");
    const DevRegion& d = ctx->regions[r];
    int idx = pos[i] - d.first_pos;
    tab[(size_t)i] = (idx >= 0 && idx < d.n_pos) ? d.tab_off + idx : -1;
  }
  if ((size_t)n > ctx->h_rows_cap) {
    if (ctx->h_rows) cudaFreeHost(ctx->h_rows);
    ctx->h_rows = NULL;
    ctx->h_rows_cap = 0;
    const size_t cap = (size_t)n + (size_t)n / 2 + 1024;  // geometric growth: page-locking is slow
    CK(cudaMallocHost(&ctx->h_rows, sizeof(uint32_t) * 33 * cap));
    ctx->h_rows_cap = cap;
  }
  int rcs = ensure_scratch(ctx, (size_t)n * (8 + 33 * 4) + 64);
  if (rcs != RV_OK) return rcs;
  int64_t* d_tab = (int64_t*)ctx->d_scratch;
  uint32_t* d_out = (uint32_t*)(ctx->d_scratch + 8 * (size_t)n);
  cudaError_t e;
  cudaMemcpyAsync(d_tab, tab.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
  int64_t threads = n * 33;
  rv_gather_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(d_tab, n, ctx->d_counts, ctx->d_cov, d_out);
  ctx->launches++;
  cudaMemcpyAsync(ctx->h_rows, d_out, sizeof(uint32_t) * 33 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
  e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(ctx, RV_ERR_CUDA, std::string("rv_fetch_rows: ") + cudaGetErrorString(e));
  *rows = ctx->h_rows;
  return RV_OK;
}

int rv_fetch_events(rv_ctx* ctx, const rv_event** events, int64_t* n_events) {
  if (!ctx || !events || !n_events) return RV_ERR_ARG;
  { const int rcs = settle(ctx); if (rcs != RV_OK) return rcs; }
  CK(cudaSetDevice(ctx->device));
  size_t n = (size_t)std::min<unsigned long long>(ctx->h_stats.n_events, (unsigned long long)ctx->L.max_events);
  if (n > ctx->h_events_cap) {
    if (ctx->h_events) cudaFreeHost(ctx->h_events);
    ctx->h_events = NULL;
    ctx->h_events_cap = 0;
    const size_t cap = n + n / 2 + 1024;
    CK(cudaMallocHost(&ctx->h_events, sizeof(rv_event) * cap));
    ctx->h_events_cap = cap;
  }
  if (n) {
    CK(cudaMemcpyAsync(ctx->h_events, ctx->d_events, sizeof(rv_event) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // restore BAM order: (region, read, order inside the read)
    std::sort(ctx->h_events, ctx->h_events + n, [](const rv_event& a, const rv_event& b) {
      if (a.region != b.region) return a.region < b.region;
      if (a.read_idx != b.read_idx) return a.read_idx < b.read_idx;
      return a.seq_no < b.seq_no;
    });
  }
  *events = ctx->h_events;
  *n_events = (int64_t)n;
  return RV_OK;
}

int rv_apply_patch(rv_ctx* ctx, const rv_patch_entry* entries, int64_t n_entries, const int32_t* cov_region,
                   const int32_t* cov_pos, const int32_t* cov_val, int64_t n_cov) {
  if (!ctx || n_entries < 0 || n_cov < 0) return RV_ERR_ARG;
  if (n_entries > ctx->L.max_patch) return fail(ctx, RV_ERR_OVERFLOW, "more patch entries than limits.max_patch");
  CK(cudaSetDevice(ctx->device));
  // entries must be grouped by (region, pos); groups are located here and scattered on the device
  std::vector<int64_t> grp_tab;
  std::vector<uint32_t> grp_first;
  std::vector<uint8_t> grp_n;
  for (int64_t i = 0; i < n_entries;) {
    int64_t j = i;
    while (j < n_entries && entries[j].region == entries[i].region && entries[j].pos == entries[i].pos) ++j;
    int r = entries[i].region;
    if (r < 0 || r >= (int)ctx->regions.size()) return fail(ctx, RV_ERR_ARG, "patch entry with bad region");
    const DevRegion& d = ctx->regions[r];
    int idx = entries[i].pos - d.first_pos;
    if (idx >= 0 && idx < d.n_pos) {
      if (j - i > 255) return fail(ctx, RV_ERR_OVERFLOW, "more than 255 keys at one position");
      grp_tab.push_back(d.tab_off + idx);
      grp_first.push_back((uint32_t)i);
      grp_n.push_back((uint8_t)(j - i));
    }
    i = j;
  }
  CK(cudaMemsetAsync(ctx->d_patch_first, 0, sizeof(uint32_t) * (size_t)(ctx->n_positions + 1), ctx->stream));
  CK(cudaMemsetAsync(ctx->d_patch_count, 0, (size_t)(ctx->n_positions + 1), ctx->stream));
  if (n_entries) CK(cudaMemcpyAsync(ctx->d_patch, entries, sizeof(rv_patch_entry) * (size_t)n_entries, cudaMemcpyHostToDevice, ctx->stream));
  int64_t ng = (int64_t)grp_tab.size();
  size_t sbytes = (size_t)ng * (8 + 4 + 1) + (size_t)n_cov * (8 + 4) + 64;
  int rcs = ensure_scratch(ctx, sbytes);
  if (rcs != RV_OK) return rcs;
  uint8_t* sp = ctx->d_scratch;
  int64_t* d_tab = (int64_t*)sp; sp += 8 * (size_t)ng;
  int64_t* d_ctab = (int64_t*)sp; sp += 8 * (size_t)n_cov;
  uint32_t* d_first = (uint32_t*)sp; sp += 4 * (size_t)ng;
  int32_t* d_cval = (int32_t*)sp; sp += 4 * (size_t)n_cov;
  uint8_t* d_n = sp;
  std::vector<int64_t> ctab((size_t)n_cov);
  std::vector<int32_t> cval((size_t)n_cov);
  int64_t nc = 0;
  for (int64_t i = 0; i < n_cov; ++i) {
    int r = cov_region[i];
    if (r < 0 || r >= (int)ctx->regions.size()) return fail(ctx, RV_ERR_ARG, "coverage patch with bad region");
    const DevRegion& d = ctx->regions[r];
    int idx = cov_pos[i] - d.first_pos;
    if (idx < 0 || idx >= d.n_pos) continue;
    ctab[nc] = d.tab_off + idx;
    cval[nc] = cov_val[i];
    nc++;
  }
  if (ng) {
    cudaMemcpyAsync(d_tab, grp_tab.data(), 8 * (size_t)ng, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_first, grp_first.data(), 4 * (size_t)ng, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_n, grp_n.data(), (size_t)ng, cudaMemcpyHostToDevice, ctx->stream);
    rv_patch_scatter_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, ctx->stream>>>(d_tab, d_first, d_n, ng, ctx->d_patch_first, ctx->d_patch_count);
    ctx->launches++;
  }
  if (nc) {
    cudaMemcpyAsync(d_ctab, ctab.data(), 8 * (size_t)nc, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_cval, cval.data(), 4 * (size_t)nc, cudaMemcpyHostToDevice, ctx->stream);
    rv_cov_scatter_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>(d_ctab, d_cval, nc, ctx->d_cov);
    ctx->launches++;
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(ctx, RV_ERR_CUDA, std::string("rv_apply_patch: ") + cudaGetErrorString(e));
  ctx->have_patch = true;
  ctx->tables_fetched = false;
  return RV_OK;
}

static void fill_score_args(rv_ctx* ctx, ScoreArgs& a) {
  a.P = ctx->P;
  a.regions = ctx->d_regions;
  a.n_regions = (int)ctx->regions.size();
  a.n_positions = ctx->n_positions;
  a.ref = ctx->d_ref;
  a.ref_start = ctx->ref_start;
  a.ref_n = ctx->ref_n;
  a.counts = ctx->d_counts;
  a.cov = ctx->d_cov;
  a.patch = ctx->d_patch;
  a.patch_first = ctx->have_patch ? ctx->d_patch_first : NULL;
  a.patch_count = ctx->d_patch_count;
  a.touched = ctx->d_touched;
  a.tabblk_region = ctx->d_tabblk_region;
  a.variants = ctx->d_variants;
  a.max_variants = (unsigned long long)ctx->L.max_variants;
  a.lgt = ctx->d_lgt;
  a.lgt_n = ctx->lgt_n;
  a.stats = ctx->d_stats;
  a.patched_queue = ctx->d_patched_queue;
  a.patched_count = ctx->d_patched_count;
}

int rv_score(rv_ctx* ctx) {
  if (!ctx) return RV_ERR_ARG;
  if (ctx->regions.empty()) return fail(ctx, RV_ERR_STATE, "rv_set_regions has not been called");
  CK(cudaSetDevice(ctx->device));
  if (!ctx->lazy && ctx->unsettled) { const int rcs = settle(ctx); if (rcs != RV_OK) return rcs; }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(cudaMemsetAsync(&ctx->d_stats->n_variants, 0, 2 * sizeof(unsigned long long), ctx->stream));
  ScoreArgs a;
  fill_score_args(ctx, a);
  if (ctx->n_positions > 0) {
    CK(cudaMemsetAsync(ctx->d_patched_count, 0, sizeof(unsigned long long), ctx->stream));
    if (ctx->P.candidates_only && !ctx->P.pileup) {
      rv_score_screen_kernel<<<(unsigned)((ctx->n_positions + 1023) / 1024), 256, 0, ctx->stream>>>(a);
      rv_score_patched_kernel<<<148 * 8, 128, 0, ctx->stream>>>(a);
      ctx->launches += 2;
      CK(cudaGetLastError());
    } else {
      unsigned grid = (unsigned)((ctx->n_positions + SCORE_BLOCK - 1) / SCORE_BLOCK);
      rv_score_kernel<<<grid, SCORE_BLOCK, 0, ctx->stream>>>(a);
      ctx->launches++;
      CK(cudaGetLastError());
      if (ctx->have_patch) {
        rv_score_patched_kernel<<<148 * 4, 128, 0, ctx->stream>>>(a);
        ctx->launches++;
        CK(cudaGetLastError());
      }
    }
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (ctx->lazy) { ctx->unsettled = true; return RV_OK; }
  CK(cudaMemcpyAsync(&ctx->h_stats.n_variants, &ctx->d_stats->n_variants, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaEventElapsedTime(&ctx->score_ms, ctx->ev0, ctx->ev1));
  if (ctx->h_stats.n_variants > (unsigned long long)ctx->L.max_variants)
    return fail(ctx, RV_ERR_OVERFLOW, "more variants than limits.max_variants");
  return RV_OK;
}

int rv_score_positions(rv_ctx* ctx, const int32_t* region, const int32_t* pos, int64_t n) {
  if (!ctx || n < 0 || (n && (!region || !pos))) return RV_ERR_ARG;
  if (ctx->regions.empty()) return fail(ctx, RV_ERR_STATE, "rv_set_regions has not been called");
  CK(cudaSetDevice(ctx->device));
  std::vector<int64_t>& tab = ctx->keep_list;
  tab.resize((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    const int r = region[i];
    if (r < 0 || r >= (int)ctx->regions.size()) return fail(ctx, RV_ERR_ARG, "rv_score_positions: bad region");
    const DevRegion& d = ctx->regions[(size_t)r];
    const int idx = pos[i] - d.first_pos;
    tab[(size_t)i] = (pos[i] >= d.r.start && pos[i] <= d.r.end && idx >= 0 && idx < d.n_pos) ? d.tab_off + idx : -1;
  }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  CK(cudaMemsetAsync(&ctx->d_stats->n_variants, 0, 2 * sizeof(unsigned long long), ctx->stream));
  if (n) {
    int rcs = ensure_scratch(ctx, 8 * (size_t)n + 64);
    if (rcs != RV_OK) return rcs;
    int64_t* d_list = (int64_t*)ctx->d_scratch;
    CK(cudaMemcpyAsync(d_list, tab.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    ScoreArgs a;
    fill_score_args(ctx, a);
    unsigned grid = (unsigned)std::min<int64_t>((n + 127) / 128, 148 * 8);
    rv_score_list_kernel<<<grid, 128, 0, ctx->stream>>>(a, d_list, n);
    ctx->launches++;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (ctx->lazy) { ctx->unsettled = true; return RV_OK; }  // (`tab` lives in the context)
  CK(cudaMemcpyAsync(&ctx->h_stats.n_variants, &ctx->d_stats->n_variants, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaEventElapsedTime(&ctx->score_ms, ctx->ev0, ctx->ev1));
  if (ctx->h_stats.n_variants > (unsigned long long)ctx->L.max_variants)
    return fail(ctx, RV_ERR_OVERFLOW, "more variants than limits.max_variants");
  return RV_OK;
}

int rv_fetch_variants(rv_ctx* ctx, const rv_variant** variants, int64_t* n_variants) {
  if (!ctx || !variants || !n_variants) return RV_ERR_ARG;
  { const int rcs = settle(ctx); if (rcs != RV_OK) return rcs; }
  CK(cudaSetDevice(ctx->device));
  size_t n = (size_t)ctx->h_stats.n_variants;
  if (n > ctx->h_variants_cap) {
    if (ctx->h_variants) cudaFreeHost(ctx->h_variants);
    ctx->h_variants = NULL;
    ctx->h_variants_cap = 0;
    const size_t cap = n + n / 2 + 1024;
    CK(cudaMallocHost(&ctx->h_variants, sizeof(rv_variant) * cap));
    ctx->h_variants_cap = cap;
  }
  if (n) {
    CK(cudaMemcpyAsync(ctx->h_variants, ctx->d_variants, sizeof(rv_variant) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::sort(ctx->h_variants, ctx->h_variants + n, [](const rv_variant& a, const rv_variant& b) {
      if (a.region != b.region) return a.region < b.region;
      if (a.pos != b.pos) return a.pos < b.pos;
      return a.rank < b.rank;
    });
  }
  *variants = ctx->h_variants;
  *n_variants = (int64_t)n;
  return RV_OK;
}

int rv_cov_summary(rv_ctx* ctx, int64_t* sum, int64_t* covered) {
  if (!ctx || !sum || !covered) return RV_ERR_ARG;
  const size_t n = ctx->regions.size();
  if (n == 0) return RV_OK;
  CK(cudaSetDevice(ctx->device));
  int rcs = ensure_scratch(ctx, 16 * n + 64);
  if (rcs != RV_OK) return rcs;
  unsigned long long* d_out = (unsigned long long*)ctx->d_scratch;
  rv_cov_summary_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(ctx->d_regions, (int)n, ctx->d_cov, d_out);
  ctx->launches++;
  CK(cudaGetLastError());
  std::vector<unsigned long long> h(2 * n);
  CK(cudaMemcpyAsync(h.data(), d_out, 16 * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (size_t r = 0; r < n; ++r) { sum[r] = (int64_t)h[2 * r]; covered[r] = (int64_t)h[2 * r + 1]; }
  return RV_OK;
}

int64_t rv_variant_count(const rv_ctx* ctx) {
  if (!ctx) return 0;
  settle(const_cast<rv_ctx*>(ctx));
  return (int64_t)ctx->h_stats.n_variants;
}

int rv_fisher_exact(rv_ctx* ctx, const int32_t* tables, int64_t n, double* out) {
  if (!ctx || (!tables && n) || (!out && n) || n < 0) return RV_ERR_ARG;
  if (n == 0) return RV_OK;
  CK(cudaSetDevice(ctx->device));
  int32_t* d_t = NULL;
  double* d_o = NULL;
  CK(cudaMalloc(&d_t, sizeof(int32_t) * 4 * (size_t)n));
  cudaError_t e = cudaMalloc(&d_o, sizeof(double) * 3 * (size_t)n);
  if (e != cudaSuccess) { cudaFree(d_t); return fail(ctx, RV_ERR_NOMEM, "rv_fisher_exact: out of device memory"); }
  cudaMemcpyAsync(d_t, tables, sizeof(int32_t) * 4 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
  rv_fisher_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_t, n, d_o, ctx->d_lgt, ctx->lgt_n);
  ctx->launches++;
  cudaMemcpyAsync(out, d_o, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
  e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_t);
  cudaFree(d_o);
  if (e != cudaSuccess) return fail(ctx, RV_ERR_CUDA, std::string("rv_fisher_exact: ") + cudaGetErrorString(e));
  return RV_OK;
}

int rv_last_kernel_ms(rv_ctx* ctx, float* pileup_ms, float* score_ms) {
  if (!ctx) return RV_ERR_ARG;
  { const int rcs = settle(ctx); if (rcs != RV_OK) return rcs; }
  if (pileup_ms) *pileup_ms = ctx->pileup_ms;
  if (score_ms) *score_ms = ctx->score_ms;
  return RV_OK;
}

int rv_last_pileup_split_ms(rv_ctx* ctx, float* classify_ms, float* gather_ms, float* walk_ms) {
  if (!ctx) return RV_ERR_ARG;
  { const int rcs = settle(ctx); if (rcs != RV_OK) return rcs; }
  if (classify_ms) *classify_ms = ctx->split_ms[0];
  if (gather_ms) *gather_ms = ctx->split_ms[1];
  if (walk_ms) *walk_ms = ctx->split_ms[2] + ctx->split_ms[3];  // rv_walk_kernel + rv_apply_kernel
  return RV_OK;
}

int rv_last_pileup_stage_ms(rv_ctx* ctx, float out[4]) {
  if (!ctx || !out) return RV_ERR_ARG;
  { const int rcs = settle(ctx); if (rcs != RV_OK) return rcs; }
  for (int k = 0; k < 4; ++k) out[k] = ctx->split_ms[k];
  return RV_OK;
}

int rv_timer_start(rv_ctx* ctx) {
  if (!ctx) return RV_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->tev0, ctx->stream));
  return RV_OK;
}

int rv_timer_stop(rv_ctx* ctx, float* ms) {
  if (!ctx || !ms) return RV_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventRecord(ctx->tev1, ctx->stream));
  CK(cudaEventSynchronize(ctx->tev1));
  CK(cudaEventElapsedTime(ms, ctx->tev0, ctx->tev1));
  return RV_OK;
}

int64_t rv_launch_count(const rv_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
