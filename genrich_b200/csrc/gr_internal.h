// gr_internal.h -- host-visible launchers of the CUDA kernels (one per stage of
// the reference's hot path) and the argument blocks they take.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

// Device-resident description of the chromosome table / slot layout.
struct DevLayout {
  int nchrom;
  u64 T;                    // total slots (multiple of GR_BLOCK_SLOTS)
  u64 nblocks;              // T / GR_BLOCK_SLOTS
  const u64* off;           // [nchrom] first slot of each chromosome (UINT64_MAX: no slots)
  const u32* len;           // [nchrom]
  const uint8_t* flags;     // [nchrom] GR_CF_*
  const int* blk2chrom;     // [nblocks]
};

// One RLE array of the reference's Pileup type (Genrich.h:173-176), device side.
struct DevRle {
  u32* end;                 // chromosome-relative exclusive end
  float* val;
  u64* chrom_start;         // [nchrom+1] first interval of each chromosome
  u64* total;               // [1] device copy of the interval count
};

// ---- K1: delta scatter (saveInterval 2516-2591) ------------------------------
void launch_unpack6(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, u64* out, int* err);
void launch_scatter(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, int packed,
                    int32_t* delta, int* err, u64* clamped);

// same, behind a locality pass that first moves the records into ~3000 position buckets
// (scratch_recs: n records; bin_cnt / bin_cursor: 8192 entries each)
// bucketed build of the delta array (large samples): count -> scan -> move -> build (+ spills)
void launch_sb_count(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, int packed,
                     u32* blk_cnt, int* err, u64* clamped);
void launch_sb_scan(cudaStream_t s, u64 nbuckets, const u32* blk_cnt, u32* blk_start, u32* cursor, u32* chunk_sum);
void launch_sb_move(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, int packed, u32* cursor, u32* bucketed,
                    uint2* spill, u32* spill_ctr);
void launch_sb_build(cudaStream_t s, const DevLayout& L, const u32* bucketed, const u32* blk_start,
                     int32_t* delta, uint2* spill, u32* spill_ctr);

// ---- K2: dense prefix sum + break compaction + bitmap (savePileupExpt 2168) ----
struct ScanScratch {
  void* ws; u64 cap;        // workspace of dense_scan_ws_bytes(cap, nchrom); cap = upper bound of the #breaks
};
size_t dense_scan_ws_bytes(u64 cap, int nchrom);
// zero_after: the scan clears every non-zero delta cell behind itself, so the array is all
// zero again when the kernel ends (the next sample then needs no 4 B/bp memset)
void launch_dense_scan(cudaStream_t s, const DevLayout& L, int32_t* delta,
                       const ScanScratch& sc, u32* bitmap, int* err, int zero_after);
// owners: run owners of the scan that filled the pages (0: the warps of launch_dense_scan)
// excl_val: value of the intervals the scan flagged as lying in a -E region (0.0f expt, SKIP ctrl)
void launch_scan_place(cudaStream_t s, const DevLayout& L, const ScanScratch& sc, DevRle out, int* err, u32 owners,
                       float excl_val);

// ---- K1+K2 fused: event buckets -> breaks, the delta cells live in shared memory only ----
// buckets = the 8192-cell blocks of the layout
void launch_fb_count(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, int packed,
                     u32* blk_cnt, int* err, u64* clamped);
// sat_res / skip_bits / seg_base: records k_sat_resolve took out (sat_res: its 3 result words; seg_base:
// arrival index of the segment's first record)
void launch_fb_move(cudaStream_t s, const DevLayout& L, const void* recs, u64 n, int packed, u32* cursor, u32* bucketed,
                    const u32* sat_res, const u32* skip_bits, u64 seg_base);
// the block-count scan in two halves with the reference's int16 saturation rule (saveInterval
// 2558-2573) resolved in between: see k_sat_resolve (gr_dense.cu)
void launch_sb_scan_a(cudaStream_t s, u64 nbuckets, const u32* blk_cnt, u32* chunk_sum, u32* sat_flag);
void launch_sb_scan_b(cudaStream_t s, u64 nbuckets, const u32* blk_cnt, u32* blk_start, u32* cursor, u32* chunk_sum);
#define SAT_MAX_BLOCKS 32                 // blocks with a saturating cell that one sample may hold
struct SatSeg { const void* d; u64 n; u64 base; int packed; int pad_; };   // one pushed segment: records, count, arrival index of the first
void launch_sat_resolve(cudaStream_t s, const u32* flag, const void* segs, int nseg, const DevLayout& L, u32* blk_cnt,
                        u32* chunk_sum, void* cells /* SAT_MAX_BLOCKS * 8192 * 16 bytes */, u32* skip_bits, u64 nbits,
                        u64* list, u32 list_cap, u32* sat_res, int* err);
// returns the number of run owners (warps or CTAs), to be handed to launch_scan_place
// blk_bed: per 8192-cell block, bit 0 = the block starts inside a -E region, bit 1 = it holds
// region boundaries (NULL: no regions); chrom_marks: region boundaries per chromosome; chrom_ever: one word per
// chromosome, set once a sample of the context had reads there (the reference's shared diff array exists from then
// on); is_expt: an experimental sample -- a chromosome that has never held a read is one interval there
// (savePileupExpt 2178-2182)
// stat: the words k_sb_scan1 leaves behind the block counters (launch_sb_scan_a's sat_flag); the scan form is
// chosen from them on the device
u32 launch_fb_scan(cudaStream_t s, const DevLayout& L, const u32* bucketed, const u32* blk_start,
                   const ScanScratch& sc, u32* bitmap, int* err, const uint8_t* blk_bed, const u32* chrom_marks,
                   u32* chrom_ever, int is_expt, const u32* stat);
// -E region boundaries as pseudo entries (cursor == NULL: count pass)
void launch_fb_marks(cudaStream_t s, const u64* marks, u32 n, u32* blk_cnt, u32* cursor, u32* bucketed);

// ---- K2b: per-chromosome sum of (float)(end-start)*val, exact fixed point ------
// acc_int / acc_frac: [nchrom] u64, zeroed by the caller.  sum = int + frac*2^-40.
void launch_rle_moment(cudaStream_t s, const DevRle& r, u64 n_upper, int nchrom,
                       u64* acc_int, u64* acc_frac);

// ---- K3: control sweep max(factor*val, lambda) + RLE re-merge (savePileupCtrl) --
struct CompactScratch { u64* st; u32* ticket; };
// n_upper: capacity of the raw array (the count is read from raw.total); factor_lambda: device float[2]
void launch_ctrl_clamp(cudaStream_t s, const DevLayout& L, const DevRle& raw, u64 n_upper,
                       const float* factor_lambda, const CompactScratch& sc,
                       DevRle out, u32* bitmap /* raw breaks in, surviving breaks out */);
// no-control variant: one interval (len, lambda) per active chromosome (saveLambda 1838)
// ends, chromosome starts and the total are written by the host; here the values (lambda, or SKIP
// where skip[i] != 0: -E regions) and the break bits at the n cell slots
void launch_ctrl_const(cudaStream_t s, const DevLayout& L, const float* lambda_dev, u64 n, DevRle out, u32* bitmap,
                       const u64* slots, const uint8_t* skip);
// per-chromosome fixed-point sums -> doubles; lambda and the scale factor from them, on the device
void launch_sums_double(cudaStream_t s, const u64* acc_int, const u64* acc_frac, int nchrom, double* out);
void launch_lambda_factor(cudaStream_t s, const double* sums, int nchrom, bool has_ctrl, u64 genome_len,
                          float* factor_lambda, int* err);

// ---- K4: breakpoint union of expt and ctrl (savePval 1768-1791) -----------------
struct RankScratch { u64* st[3]; u32* ticket; };
// pass A: per-block exclusive ranks of E, C and E|C; totals[3]
void launch_union_rank(cudaStream_t s, const DevLayout& L, const u32* bmE, const u32* bmC,
                       const RankScratch& sc, u64* rankE, u64* rankC, u64* rankU, u64* totals);
// pass B: emit merged intervals (end, expt value, ctrl value) + union bitmap
void launch_union_emit(cudaStream_t s, const DevLayout& L, const u32* bmE, const u32* bmC,
                       const u64* rankE, const u64* rankC, const u64* rankU,
                       const float* exptVal, const float* ctrlVal,
                       u32* pEnd, float* pExpt, float* pCtrl, u32* bmU, u64* chrom_start,
                       const u64* total);

// ---- K5: -log10 p per interval through a table of distinct (expt, ctrl) pairs ---
struct PairTable {
  u64* keys;      // (expt bits << 32) | ctrl bits ; EMPTY = ~0
  u64* lens;      // total bp per pair (BH histogram, hashPval 300)
  float* pval;    // filled by launch_pair_eval
  float* qval;    // filled after BH
  u32 cap;        // power of two
  u32* count;     // [1] occupied slots
};
// n_upper sizes the launch, the count itself is read from *n_dev
void launch_pair_insert(cudaStream_t s, const float* pExpt, const float* pCtrl, u64 n_upper, const u64* n_dev,
                        const PairTable& t, u32* slot, int* err);
void launch_pair_eval(cudaStream_t s, const PairTable& t);
void launch_gather_f32(cudaStream_t s, const float* table, const u32* slot, u64 n_upper, const u64* n_dev, float* out);

// ---- K6: Fisher combine over replicates (combinePval 612, multPval 567) ----------
struct RepView {            // one replicate's p arrays
  const u32* bmU; const u64* rankU; const float* pval; const uint8_t* present; /* [nchrom] */
};
void launch_or_bitmaps(cudaStream_t s, const u32* a, const u32* b, u32* out, u64 nwords);
void launch_block_rank(cudaStream_t s, const DevLayout& L, const u32* bm, const CompactScratch& sc,
                       u64* rank, u64* total);
// for each combined break: sum of replicate p (double) and df -> sum_out/df_out; end_out
// reps: HOST array (the views travel to the kernel as parameters, 8 replicates per launch)
void launch_fisher_emit(cudaStream_t s, const DevLayout& L, const u32* bmAll, const u64* rankAll,
                        const RepView* reps, int nrep, u32* end_out, double* sum_out,
                        int* df_out, u64* chrom_start, const u64* total);
void launch_fisher_eval(cudaStream_t s, const double* sum, const int* df, u64 n, float* pcomb);   // every interval (small inputs)
// the same through a table of distinct sums: t empty on entry (table_alloc); GR_DE_TABLE in *err: it was too small
void launch_fisher_table(cudaStream_t s, const double* sum, const int* df, u64 n, const PairTable& t, u32* slot,
                         float* pcomb, int* err);

// ---- K7: Benjamini-Hochberg over the histogram of distinct p (computeQval 352) ----
// generic single-key table used when the final p array is the Fisher-combined one
void launch_key_insert(cudaStream_t s, const u32* pEnd, const float* pval, u64 n,
                       const u64* chrom_start, int nchrom, const PairTable& t, u32* slot, int* err);
// single-replicate runs: bp per pair-table slot (the intervals know their slot), then by_pval = 1 below
void launch_slot_hist(cudaStream_t s, const u32* pEnd, const u32* slot, u64 n_upper, const u64* n_dev,
                      const u64* chrom_start, int nchrom, u64* lens);
void launch_table_compact(cudaStream_t s, const PairTable& t, const CompactScratch& sc,
                          u32* keys_out, u64* lens_out, u64* count_out, int by_pval);
// sort (keys,lens) ascending by key, merge equal keys, compute q per distinct key.
// Work arrays must hold n entries each.  Results: dk (distinct keys), dq (their q),
// *dcount.  logN = -log10f(genomeLen) is evaluated by the caller (host libm), like
// every other scalar of the reference's CLI layer.
struct BhWork {
  u32* k0; u32* k1; u64* l0; u64* l1;   // ping-pong
  u32* hist;                            // radix histograms
  u64* ksum;                            // suffix sums
  float* x;                             // per-key statistic
  u32* dk; float* dq; u64* dl; u64* dcount;
  CompactScratch sc;
  u64 cap;
};
void launch_bh(cudaStream_t s, const u32* keys, const u64* lens, u64 n, float logN, const BhWork& w);
void launch_table_q(cudaStream_t s, const PairTable& t, const u32* dk, const float* dq, const u64* dcount, int by_pval);

// ---- K8: peak scan (callPeaks 977) ---------------------------------------------
struct PeakRec {           // mirrors gr_peak
  int32_t chrom; u32 summit; i64 start; i64 end; float auc, pval, qval, reserved;
};
struct PeakWork {
  u32* ev_idx;             // significant / SKIP interval indices, in order
  u64* ev_count;
  u32* head_idx;           // indices into ev_idx where a candidate peak starts
  u64* head_count;
  PeakRec* cand;           // one per head
  uint8_t* cand_ok;
  PeakRec* out; u64* out_count; u64* peak_bp;
  CompactScratch sc;
};
// Counts stay on the device: n_upper / nev_upper size the status words and buffers, the kernels
// are persistent and read the actual counts (*n_dev, *w.ev_count, *w.head_count) themselves.
void launch_peak_events(cudaStream_t s, const float* v, u64 n_upper, const u64* n_dev, float thr, const PeakWork& w);
// heads -> per-candidate walk -> compaction of valid peaks.  head_idx holds nev_upper entries;
// cand / cand_ok / out hold hcap (more candidates than that: GR_DE_CAP in *err, the host retries)
void launch_peak_chain(cudaStream_t s, const u32* pEnd, const float* pval, const float* qval,
                       const u64* chrom_start, int nchrom, float thr, int qopt, int max_gap,
                       float min_auc, int min_len, const PeakWork& w, u64 nev_upper, u64 hcap, int* err);

// ---- small utilities -------------------------------------------------------------
void launch_fill_chrom_start(cudaStream_t s, const DevLayout& L, u64* chrom_start, const u64* total);
void launch_fill_u64(cudaStream_t s, u64* p, u64 v, u64 n);
u64 lookback_tiles_for(u64 n_items, u32 tile);

// number of kernel launches issued by this library (all contexts)
extern unsigned long long g_gr_launches;
#define GR_NOTE_LAUNCH() (++g_gr_launches)
