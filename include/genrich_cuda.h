/* genrich_cuda.h -- C-ABI of libgenrich_cuda.so (sm_100a).
 *
 * The reference (jsh58/Genrich v0.6.2, one C translation unit) has no plugin or
 * FFI surface; its only stable contract is the process boundary.  This header
 * therefore cuts the program at the two seams where its hot path begins and
 * ends, and exports exactly what crosses them:
 *
 *   seam IN   saveInterval(chrom,start,end,qname,count,...)   Genrich.c:2516
 *             -> gr_sample_begin / gr_push_intervals
 *   hot path  savePileupExpt 2168, calcLambda 1817, calcFactor 1980,
 *             savePileupCtrl 2052, savePileupNoCtrl 1883, savePval 1720,
 *             combinePval 612, computeQval 352, callPeaks 977
 *             -> gr_sample_pileup / gr_replicate_finish / gr_call_peaks
 *   seam OUT  printPeak(out,...,name,start,end,count,signal,pval,qval,pos)
 *             Genrich.c:885; printInterval 770 / printPile 1697 for -f/-k
 *             -> gr_peak records / gr_fetch_intervals
 *
 * Plain pointers and sizes only.  Every entry returns 0 on success or a
 * gr_status (gr_strerror() gives the reference's own message text where the
 * failure corresponds to one of its errCode cases, Genrich.h:97-154).
 * One submitting host thread per context; the library never calls back.
 *
 * Multi-GPU: one context per device, each owning a subset of chromosomes
 * (gr_chrom.owned).  Chromosomes are independent except for three scalars and
 * one table, which the HOST exchanges between contexts/ranks:
 *   - per-chromosome sum(len*val) of the experimental pileup  -> lambda
 *   - per-chromosome sum(len*val) of the control pileup       -> scale factor
 *   - the genome-wide histogram of distinct -log10(p) (gr_bh_local_hist /
 *     gr_bh_set_global), which the caller all-gathers (NCCL) between the two.
 */
#ifndef GENRICH_CUDA_H
#define GENRICH_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GR_SKIP (-1.0f)   /* SKIP, Genrich.h:27 */

typedef enum gr_status {
  GR_OK = 0,
  GR_ERR_ARG = 1,        /* bad argument / call order */
  GR_ERR_CUDA = 2,       /* CUDA runtime failure (gr_last_error_detail) */
  GR_ERR_MEM = 3,        /* ERRMEM      Genrich.h:111 */
  GR_ERR_POS = 4,        /* ERRPOS      Genrich.c:2531-2535: start >= chrom len */
  GR_ERR_EXPT = 5,       /* ERREXPT     Genrich.c:2292: no analyzable fragments */
  GR_ERR_GENOME = 6,     /* ERRGEN      Genrich.c:1828: genome length 0 */
  GR_ERR_PILE = 7,       /* ERRPILE     Genrich.c:1921,1969: pileup < 0 */
  GR_ERR_COUNT = 8,      /* ERRALNS     Genrich.c:2400: count not in {1,2,3,4,5,6,8,10} */
  GR_ERR_CHROM = 9,      /* interval on an unknown / unowned chromosome */
  GR_ERR_DF = 10,        /* ERRDF       Genrich.c:556: > 200 replicates */
  GR_ERR_GENLEN = 11,    /* Genrich.c:377-382: histogram length != genome length */
  GR_ERR_NODEVICE = 12,  /* no CUDA device / library built without one */
  GR_ERR_SATURATED = 13  /* the reference's int16 delta counters saturate at 32767 interval starts (32768 ends) on
                            one base and it then SKIPS further intervals, in arrival order (saveInterval,
                            Genrich.c:2558-2573).  The default path reproduces that rule (gr_sample_skipped); this
                            status is left for what it cannot hold: more than 32 such 8192-bp blocks in one sample,
                            or the dense formulation (GR_FUSED=0, a measurement aid), which only detects it. */
} gr_status;

/* One reference sequence (Chrom, Genrich.h:183-201, the fields the path reads). */
typedef struct gr_chrom {
  uint32_t len;      /* Chrom.len, < 2^31 (getInt on LN:, Genrich.c:4299) */
  uint8_t  skip;     /* Chrom.skip (-e list) : ignored everywhere */
  uint8_t  owned;    /* 1 if this context computes this chromosome */
  uint16_t reserved;
} gr_chrom;

/* Peak-calling parameters, as runProgram() passes them (Genrich.c:5386-5395). */
typedef struct gr_params {
  float    min_pqval;   /* -log10 threshold, already converted (Genrich.c:5817) */
  int32_t  qval_opt;    /* 1: threshold applies to q (-q), 0: to p (-p) */
  float    min_auc;     /* -a */
  int32_t  min_len;     /* -l */
  int32_t  max_gap;     /* -g */
  int32_t  keep_pileups;/* 1: retain expt/ctrl columns for -f/-k output */
  uint64_t genome_len;  /* -L; 0 = compute (calcLambda 1819-1830, findPeaks 1091-1101) */
} gr_params;

/* Per-replicate scalars (what -v prints: Genrich.c:1888, 2058, 2063). */
typedef struct gr_sample_stats {
  double   frag_len;     /* sum over expt pileup of len*val (savePileupExpt) */
  double   ctrl_frag;    /* same over the control pileup (calcFactor), 0 if none */
  float    lambda;       /* calcLambda */
  float    factor;       /* calcFactor; 1.0f when no control */
  uint64_t genome_len;   /* denominator used for lambda */
  uint64_t n_expt;       /* RLE intervals, experimental (owned chroms) */
  uint64_t n_ctrl;       /* RLE intervals, control after max(ctrl,lambda) merge */
  uint64_t n_pval;       /* p-value intervals */
  uint64_t n_clamped;    /* intervals clamped to [0,len] (saveInterval 2522-2544) */
} gr_sample_stats;

/* One called peak = the argument list of printPeak (Genrich.c:885). */
typedef struct gr_peak {
  int32_t  chrom;      /* index into the gr_chrom table */
  uint32_t summit;     /* offset of summit from start */
  int64_t  start;
  int64_t  end;
  float    auc;        /* signalValue */
  float    pval;       /* -log10(p) at summit */
  float    qval;       /* -log10(q) at summit, GR_SKIP if !qval_opt */
  float    reserved;
} gr_peak;

typedef struct gr_run_stats {
  uint64_t genome_len;   /* findPeaks 1091-1101 */
  uint64_t n_peaks;
  uint64_t peak_bp;      /* Genrich.c:924 */
  uint64_t n_intervals;  /* final p/q interval count (owned chroms) */
  uint64_t n_distinct_p; /* size of the BH table (0 if !qval_opt) */
  int32_t  all_q_one;    /* "Warning! All q-values are 1" Genrich.c:245 */
  int32_t  n_replicates;
} gr_run_stats;

typedef struct gr_ctx gr_ctx;

/* ---- context ------------------------------------------------------------ */
/* saveChrom 4220 builds Chrom[]; here the table arrives complete. */
int gr_create(gr_ctx** out, const gr_chrom* chroms, int32_t nchrom,
              const gr_params* params, int32_t device);
void gr_destroy(gr_ctx* ctx);
/* -E: genomic regions to exclude (loadBED 5187 -> saveXBed 1144, called from saveChrom 4265).
 * The records as the BED file lists them: any order, overlapping or not, start < end; the
 * library sorts, clamps and merges them per chromosome exactly as saveXBed does.  Inside a
 * region the experimental pileup is 0 (Genrich.c:2248), the control pileup -- and with it p and
 * q -- SKIP (2124, 1632), region boundaries always close an interval (2241, 2122), the regions
 * count neither for lambda (1824) nor for the scale factor (2016) nor for the BH genome length
 * (1097), and no peak extends into one (1031).  Call before the first sample of the context.
 * gr_excluded_bp: excluded bp per chromosome after merging (for callers that pass their own
 * genome length to gr_replicate_finish* / gr_bh_set_global). */
int gr_set_exclusions(gr_ctx* ctx, const int32_t* chrom, const uint32_t* start,
                      const uint32_t* end, uint64_t n);
int gr_excluded_bp(gr_ctx* ctx, uint64_t* per_chrom /* [nchrom] */);
int gr_set_params(gr_ctx* ctx, const gr_params* params);
/* Forget all replicates and results (a new runProgram on the same chromosome table). */
int gr_reset(gr_ctx* ctx);
const char* gr_strerror(int status);
const char* gr_last_error_detail(const gr_ctx* ctx);

/* ---- seam IN: intervals --------------------------------------------------
 * gr_sample_begin zeroes the delta arrays (runProgram 5503-5510).  `save` is the
 * per-chromosome Chrom.save flag for this replicate (Genrich.c:5463, 4231);
 * NULL = all saved.  A control sample inherits the flags of its experimental
 * sample.  Records are int32 x 4 = chrom, start, end, count with
 * count in {1,2,3,4,5,6,8,10} (weight 1/count).  start < 0 and end > len are
 * clamped as saveInterval does (2522-2544); start >= len (or end < start, which the
 * reference's callers never produce) is GR_ERR_POS, reported by gr_sample_pileup.
 * Records are consumed by gr_sample_pileup: host buffers are copied before the push
 * returns, DEVICE buffers must stay unchanged until gr_sample_pileup has returned. */
int gr_sample_begin(gr_ctx* ctx, int32_t is_ctrl, const uint8_t* save);
int gr_push_intervals(gr_ctx* ctx, const int32_t* recs, uint64_t n);        /* host memory (pinned or not) */
int gr_push_intervals_device(gr_ctx* ctx, const int32_t* d_recs, uint64_t n);/* device memory */
/* Optional hint, callable at any time: start copying a PINNED host buffer that a later
 * gr_push_intervals(ctx, recs, n) -- same pointer, same n, contents unchanged -- will
 * consume.  The copy runs on the library's copy stream behind whatever the device is
 * doing (e.g. the control sample travels while the treatment sample is integrated).
 * At most two buffers can be in flight; a third request is ignored (returns 0). */
int gr_prefetch_intervals(gr_ctx* ctx, const int32_t* recs, uint64_t n);

/* Compact wire format of the same four saveInterval arguments, 8 bytes per record:
 *   bits  0-31  start            (0 <= start < 2^32)
 *   bits 32-45  end - start      (< GR_PACK_MAX_LEN)
 *   bits 46-59  chromosome index (< GR_PACK_MAX_CHROM)
 *   bits 60-63  count            (as above)
 * Semantics are those of gr_push_intervals on the unpacked record; a record that does not
 * fit (start < 0, a longer interval, more chromosomes) goes through gr_push_intervals --
 * both may be mixed freely within one sample.  The host -> device copy is what bounds the
 * end-to-end rate of the hot path (PCIe), so producers should pack: half the bytes.
 * gr_push_packed accepts host (pinned or not) and device pointers. */
#define GR_PACK_MAX_LEN   (1u << 14)
#define GR_PACK_MAX_CHROM (1u << 14)
#define GR_PACK(chrom, start, end, count) \
  ((uint64_t)(uint32_t)(start) | ((uint64_t)((end) - (start)) << 32) | ((uint64_t)(chrom) << 46) | ((uint64_t)(count) << 60))
int gr_push_packed(gr_ctx* ctx, const uint64_t* recs, uint64_t n);
int gr_prefetch_packed(gr_ctx* ctx, const uint64_t* recs, uint64_t n);

/* The densest wire format, 6 bytes per record (three little-endian uint16 words): the start is
 * given as a CELL of the context's own layout, so the chromosome index needs no bits of its own.
 *   words 0-1  cell of the start = gr_pack6_layout()[chrom] + start   (32 bits)
 *   word  2    bits 0-11 end - start (< GR_PACK6_MAX_LEN), bits 12-15 count
 * gr_pack6_layout fills cell_offset[nchrom] (UINT64_MAX for chromosomes this context does not hold:
 * skipped or owned by another rank) and fails with GR_ERR_ARG when the context's layout does not
 * fit 32 bits (genomes beyond ~4.29 G cells on one device) -- then the format is not available.
 * A record must lie inside its chromosome (0 <= start <= end <= len): clamping, longer intervals
 * and everything else keep going through gr_push_packed / gr_push_intervals; all three may be
 * mixed within a sample.  The library expands the records to GR_PACK words on the device (one
 * streaming pass, 14 B per record) -- against 25 % fewer bytes over PCIe, which is what bounds the
 * end-to-end rate.  Host (pinned or not) and device pointers are accepted. */
#define GR_PACK6_MAX_LEN (1u << 12)
#define GR_PACK6(dst, cell, len, count) \
  ((dst)[0] = (uint16_t)(cell), (dst)[1] = (uint16_t)((uint32_t)(cell) >> 16), (dst)[2] = (uint16_t)((len) | ((count) << 12)))
int gr_pack6_layout(gr_ctx* ctx, uint64_t* cell_offset /* [nchrom] */);
int gr_push_packed6(gr_ctx* ctx, const uint16_t* recs, uint64_t n);
int gr_prefetch_packed6(gr_ctx* ctx, const uint16_t* recs, uint64_t n);

/* Integrate the current sample (savePileupExpt 2168 / the RLE pass of
 * calcFactor 1980).  chrom_sums[nchrom] receives, per owned chromosome, the
 * double sum of (float)(end-start)*val (0 elsewhere); the caller adds them in
 * chromosome order across contexts to get fragLen / ctrlFrag. */
int gr_sample_pileup(gr_ctx* ctx, double* chrom_sums);
/* The reference's int16 saturation rule (saveInterval, Genrich.c:2558-2573): an interval is dropped, whole,
 * when the delta counter of its start already holds INT16_MAX whole fragments, else when that of its end
 * holds INT16_MIN -- in ARRIVAL order, i.e. the order of the pushes and of the records inside a push (a
 * caller that wants the reference's decisions pushes in file order).  The library applies the same rule
 * while it integrates the sample.  This call reports what was dropped from the last experimental
 * (is_ctrl = 0) or control sample: counts by cause, and for the "Warning! ... skipped due to overflow /
 * underflow" lines (2560-2571) the dropped records as (arrival index << 1) | (1 if underflow), the first
 * 262144 of them, in arrival order.  The array is owned by the context (valid until its next call).
 * Waits for the device. */
int gr_sample_skipped(gr_ctx* ctx, int32_t is_ctrl, uint64_t* n_overflow, uint64_t* n_underflow,
                      const uint64_t** list, uint64_t* n_list);
/* Nothing on this path makes the host wait for the device unless it asks for a result.
 * chrom_sums == NULL above only enqueues the work; the sums of the experimental and the
 * control sample (NULL: not wanted) are then fetched together, one device round trip: */
int gr_sample_sums(gr_ctx* ctx, double* expt_sums, double* ctrl_sums);

/* Close the replicate: lambda, scale factor, max(ctrl*factor, lambda) sweep
 * (savePileupCtrl 2052 / savePileupNoCtrl 1883), breakpoint merge and
 * -log10(p) (savePval 1720).  has_ctrl = 0: no control was pushed.
 * genome_len = 0: use the sum of saved owned chromosome lengths (single
 * context) -- multi-context callers pass the global value. */
int gr_replicate_finish(gr_ctx* ctx, double frag_len, double ctrl_frag,
                        int32_t has_ctrl, uint64_t genome_len,
                        gr_sample_stats* stats /* NULL: do not wait for the counts */);

/* The same with the sums left where they are -- on the device: lambda and the scale factor are
 * computed there (calcLambda 1817 / calcFactor 2043-2045, per-chromosome doubles added in
 * chromosome order) and the host never waits.  gr_sums_device exposes the two double[nchrom]
 * arrays so that a multi-GPU launcher can all-reduce them in place, in gr_stream's order
 * (every chromosome has one owner: the sum is exact).  Statistics afterwards, on request. */
int gr_replicate_finish_device(gr_ctx* ctx, int32_t has_ctrl, uint64_t genome_len);
int gr_sums_device(gr_ctx* ctx, double** d_expt_sums, double** d_ctrl_sums);
void* gr_stream(gr_ctx* ctx);                       /* the context's cudaStream_t */
int gr_replicate_stats(gr_ctx* ctx, int32_t replicate, gr_sample_stats* stats);

/* Convenience for a single context: gr_sample_pileup for the pending
 * sample(s) + gr_replicate_finish with local sums. */
int gr_replicate_end(gr_ctx* ctx, gr_sample_stats* stats);

/* ---- peaks ---------------------------------------------------------------
 * gr_pvalues_finalize: Fisher combine when > 1 replicate (combinePval 612).
 * gr_bh_local_hist / gr_bh_set_global: the exchange step of computeQval 352;
 * pointers are DEVICE pointers (keys = float bits of -log10 p, lens = bp).
 * gr_bh_local_hist returns with the list COMPLETE in device memory (it waits for the
 * context's stream), so the caller may read it on any stream; gr_bh_set_global reads its
 * arguments on gr_stream(ctx) -- a caller that produced them on another stream orders the
 * two itself (an event, or by issuing its collective on gr_stream(ctx)).
 * gr_call_peaks runs whatever of these has not been run (single context),
 * then callPeaks 977.  The returned array is owned by the context (pinned host memory the
 * records were copied into; no second copy) and stays valid until the next gr_call_peaks,
 * gr_reset or gr_destroy on that context. */
int gr_pvalues_finalize(gr_ctx* ctx);
int gr_bh_local_hist(gr_ctx* ctx, const uint32_t** d_keys,
                     const uint64_t** d_lens, uint64_t* n);
int gr_bh_set_global(gr_ctx* ctx, const uint32_t* d_keys,
                     const uint64_t* d_lens, uint64_t n, uint64_t genome_len);
/* The same exchange through HOST memory, for one process that drives several contexts without a
 * device-side collective: keys / lens returned by the first call are host arrays owned by the
 * context (valid until its next call); the second takes the concatenation of every context's list. */
int gr_bh_local_hist_host(gr_ctx* ctx, const uint32_t** keys, const uint64_t** lens, uint64_t* n);
int gr_bh_set_global_host(gr_ctx* ctx, const uint32_t* keys, const uint64_t* lens, uint64_t n,
                          uint64_t genome_len);
/* -P (callPeaksLog, Genrich.c:1277-1470: peaks from an already written -f log): the significance
 * values come from the caller instead of from pileups.  chrom_start[nchrom+1] delimits every
 * chromosome's run inside end / pval / qval (n = chrom_start[nchrom] intervals; an interval starts
 * where the previous one of its chromosome ends, the first at 0; GR_SKIP = "NA" / excluded).
 * qval must be given iff the context was created with qval_opt: the q-values are the log's, no
 * Benjamini-Hochberg pass is run (1343-1345, 1383-1389).  gr_call_peaks follows. */
int gr_load_pvalues(gr_ctx* ctx, const uint64_t* chrom_start, const uint32_t* end,
                    const float* pval, const float* qval, uint64_t n);
int gr_call_peaks(gr_ctx* ctx, const gr_peak** peaks, uint64_t* n,
                  gr_run_stats* stats);
/* The same records where gr_call_peaks left them in DEVICE memory (valid until the next
 * call on the context): lets a multi-GPU launcher all-gather them without a host bounce. */
int gr_peaks_device(gr_ctx* ctx, const gr_peak** d_peaks, uint64_t* n);
/* gr_call_peaks in two halves, for launchers that gather several contexts' peaks with ONE wait for the device
 * (one process per GPU: an NCCL all-gather of every rank's slot).  gr_call_peaks_enqueue runs the peak scan
 * without waiting and hands out the device address of a SLOT: a 64-byte gr_peak_slot header, which the device
 * fills in stream order, followed by record_cap gr_peak records.  The caller moves the first 64 + k * sizeof(gr_peak) bytes
 * wherever it wants them on gr_stream(ctx), waits once, and passes the header as it reads on the host to
 * gr_call_peaks_done.  redo = 1: a buffer was too small, enqueue again; redo = 2: call gr_call_peaks instead
 * (rare: a table overflowed).  With -q the histogram exchange (gr_bh_local_hist / gr_bh_set_global) comes first. */
typedef struct gr_peak_slot {
  uint64_t n_peaks;      /* records that follow the header (may exceed what the caller chose to move) */
  uint64_t peak_bp;      /* Genrich.c:924 */
  int32_t  flags;        /* device-side condition bits, interpreted by gr_call_peaks_done */
  int32_t  reserved;
  uint64_t n_intervals;  /* final p/q interval count (owned chromosomes) */
  uint64_t pad[4];
} gr_peak_slot;
int gr_call_peaks_enqueue(gr_ctx* ctx, const void** d_slot, uint64_t* record_cap);
int gr_call_peaks_done(gr_ctx* ctx, const gr_peak_slot* header, int32_t* redo, gr_run_stats* stats);
/* gr_call_peaks with peaks == NULL leaves the records on the device (n and stats are still
 * filled).  Host utility for launchers that gathered several contexts' lists (each in chromosome
 * order, every chromosome in exactly one of them): one list in chromosome order, the order in
 * which callPeaks numbers the peaks (Genrich.c:986).  `out` holds the sum of the counts. */
int gr_merge_peaks(const gr_peak* const* lists, const uint64_t* counts, int32_t nlists, gr_peak* out);

/* ---- seam OUT for -f / -k (printInterval 770, printPile 1697) -------------
 * which: 0 = experimental pileup, 1 = control pileup (last replicate),
 *        2 = p-value intervals of `replicate` (replicate == n_replicates:
 *            the Fisher-combined array), 3 = q-value intervals.
 * For which 2/3 with keep_pileups, expt/ctrl receive the pileup columns.
 * Host arrays owned by the context, valid until the next fetch. */
int gr_fetch_intervals(gr_ctx* ctx, int32_t which, int32_t replicate,
                       int32_t chrom, const uint32_t** end, const float** val,
                       const float** expt, const float** ctrl, uint64_t* n);

/* Device time (ms, CUDA events on the library's stream) of named stages since
 * the last gr_timing_reset: fills up to `cap` entries. */
typedef struct gr_stage_time { const char* name; double ms; uint64_t launches; uint64_t bytes; } gr_stage_time;
int gr_timing_enable(gr_ctx* ctx, int32_t on);
int gr_timing_get(gr_ctx* ctx, gr_stage_time* out, int32_t cap, int32_t* n);
int gr_timing_reset(gr_ctx* ctx);
uint64_t gr_kernel_launches(const gr_ctx* ctx);
/* Which per-base scan the last sample of the context chose on the device (the choice needs no host round trip:
 * both kernels are launched and the one whose turn it is not returns at once): form 0 = rank form (k_fr_scan),
 * 1 = CTA form (k_fb_scan), -1 = the sample did not go through the bucketed path.  hot_entries / entries: the
 * statistic the choice is made from (entries in 8192-cell blocks that hold 1024 or more, all entries).  Waits for
 * the stream. */
int gr_scan_form(gr_ctx* ctx, int32_t* form, uint64_t* hot_entries, uint64_t* entries);
/* CUDA events on the library's own stream: ms between start and stop as the device saw it */
int gr_timer_start(gr_ctx* ctx);
int gr_timer_stop(gr_ctx* ctx, double* ms);
int gr_synchronize(gr_ctx* ctx);
/* page-locked host memory for the caller-owned interval staging buffers */
void* gr_pinned_alloc(size_t bytes);
void gr_pinned_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
