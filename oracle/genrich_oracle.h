/* genrich_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain sequential C) of the reference's pileup -> p-value ->
 * q-value -> peak path, exposing the same call sequence and the same structs as
 * include/genrich_cuda.h so parity tests can drive both sides identically.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product never does.
 */
#ifndef GENRICH_ORACLE_H
#define GENRICH_ORACLE_H
#include "../include/genrich_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

int  orc_create(orc_ctx** out, const gr_chrom* chroms, int32_t nchrom,
                const gr_params* params);
void orc_destroy(orc_ctx* ctx);
int  orc_set_params(orc_ctx* ctx, const gr_params* params);
int  orc_set_exclusions(orc_ctx* ctx, const int32_t* chrom, const uint32_t* start,
                        const uint32_t* end, uint64_t n);
int  orc_excluded_bp(orc_ctx* ctx, uint64_t* per_chrom);
int  orc_sample_begin(orc_ctx* ctx, int32_t is_ctrl, const uint8_t* save);
int  orc_push_intervals(orc_ctx* ctx, const int32_t* recs, uint64_t n);
int  orc_sample_pileup(orc_ctx* ctx, double* chrom_sums);
int  orc_sample_sums(orc_ctx* ctx, double* expt_sums, double* ctrl_sums);
int  orc_sample_skipped(orc_ctx* ctx, int32_t is_ctrl, uint64_t* n_overflow, uint64_t* n_underflow,
                        const uint64_t** list, uint64_t* n_list);   /* saveInterval 2558-2573 */
int  orc_replicate_finish(orc_ctx* ctx, double frag_len, double ctrl_frag,
                          int32_t has_ctrl, uint64_t genome_len,
                          gr_sample_stats* stats);
int  orc_replicate_end(orc_ctx* ctx, gr_sample_stats* stats);
int  orc_pvalues_finalize(orc_ctx* ctx);
int  orc_bh_local_hist(orc_ctx* ctx, const uint32_t** keys,
                       const uint64_t** lens, uint64_t* n);
int  orc_bh_set_global(orc_ctx* ctx, const uint32_t* keys,
                       const uint64_t* lens, uint64_t n, uint64_t genome_len);
int  orc_load_pvalues(orc_ctx* ctx, const uint64_t* chrom_start, const uint32_t* end,
                      const float* pval, const float* qval, uint64_t n);   /* callPeaksLog 1277 */
int  orc_call_peaks(orc_ctx* ctx, const gr_peak** peaks, uint64_t* n,
                    gr_run_stats* stats);
int  orc_fetch_intervals(orc_ctx* ctx, int32_t which, int32_t replicate,
                         int32_t chrom, const uint32_t** end, const float** val,
                         const float** expt, const float** ctrl, uint64_t* n);

/* function-level restatements (pinned against oracle/_ref/libref_funcs.so) */
float  orc_units_to_val(int32_t units);          /* getVal o updateVal state */
int32_t orc_cell_cov(int64_t units);             /* Diff.cov of a cell holding that many 1/120ths (addFrac 2311 / subFrac 2412) */
float  orc_calc_pval(float expt, float ctrl);    /* calcPval 1628 */
double orc_pchisq(double x, int df);             /* pchisq 555 */
float  orc_mult_pval(const float* vals, int n);  /* multPval 567 */

#ifdef __cplusplus
}
#endif
#endif
