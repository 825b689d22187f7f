/* cli_twin.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The entry points of include/genrich_cuda.h that the host program (genrich_b200/cli/)
 * calls, implemented on the CPU oracle (genrich_oracle.c).  Linking the host program's own
 * sources against this file instead of libgenrich_cuda.so gives `_test/genrich-b200-oracle`:
 * the same option parsing, SAM/BAM decode, mate pairing, duplicate removal, interval
 * transforms and text writers, with the oracle standing in for the GPU.  The CPU test-suite
 * runs it on the SAM view of every seeded case and compares its narrowPeak / -f / -k text
 * with the files the unmodified reference wrote (tests/golden) BYTE FOR BYTE -- the host C
 * code is thereby checked on every round, without a GPU.  The product binary
 * (genrich_b200/bin/genrich-b200) links libgenrich_cuda.so and nothing from oracle/.
 */
#include <stdlib.h>
#include <string.h>
#include "genrich_oracle.h"

struct gr_ctx { orc_ctx* o; };

int orc_set_params(orc_ctx* x, const gr_params* p);

int gr_create(gr_ctx** out, const gr_chrom* chroms, int32_t nchrom, const gr_params* params, int32_t device) {
  (void)device;
  gr_ctx* x = (gr_ctx*)calloc(1, sizeof *x);
  int rc = orc_create(&x->o, chroms, nchrom, params);
  if (rc) { free(x); return rc; }
  *out = x;
  return GR_OK;
}
void gr_destroy(gr_ctx* x) { if (x) { orc_destroy(x->o); free(x); } }
int gr_set_params(gr_ctx* x, const gr_params* p) { return orc_set_params(x->o, p); }
int gr_set_exclusions(gr_ctx* x, const int32_t* chrom, const uint32_t* start, const uint32_t* end, uint64_t n) {
  return orc_set_exclusions(x->o, chrom, start, end, n);
}
int gr_sample_begin(gr_ctx* x, int32_t is_ctrl, const uint8_t* save) { return orc_sample_begin(x->o, is_ctrl, save); }
int gr_push_intervals(gr_ctx* x, const int32_t* recs, uint64_t n) { return orc_push_intervals(x->o, recs, n); }
int gr_push_packed(gr_ctx* x, const uint64_t* recs, uint64_t n) {        /* GR_PACK, include/genrich_cuda.h */
  int32_t r[4 * 1024];
  for (uint64_t i = 0; i < n;) {
    uint64_t k = 0;
    for (; k < 1024 && i < n; k++, i++) {
      const uint64_t w = recs[i];
      const int32_t start = (int32_t)(uint32_t)w;
      r[4 * k] = (int32_t)((w >> 46) & 0x3fff);
      r[4 * k + 1] = start;
      r[4 * k + 2] = start + (int32_t)((w >> 32) & 0x3fff);
      r[4 * k + 3] = (int32_t)(w >> 60);
    }
    int rc = orc_push_intervals(x->o, r, k);
    if (rc) return rc;
  }
  return GR_OK;
}
int gr_sample_pileup(gr_ctx* x, double* sums) { return orc_sample_pileup(x->o, sums); }
int gr_sample_sums(gr_ctx* x, double* e, double* c) { return orc_sample_sums(x->o, e, c); }
int gr_sample_skipped(gr_ctx* x, int32_t is_ctrl, uint64_t* a, uint64_t* b, const uint64_t** list, uint64_t* n) {
  return orc_sample_skipped(x->o, is_ctrl, a, b, list, n);
}
int gr_replicate_end(gr_ctx* x, gr_sample_stats* st) { return orc_replicate_end(x->o, st); }
int gr_load_pvalues(gr_ctx* x, const uint64_t* cs, const uint32_t* end, const float* pval, const float* qval, uint64_t n) {
  return orc_load_pvalues(x->o, cs, end, pval, qval, n);
}
int gr_call_peaks(gr_ctx* x, const gr_peak** peaks, uint64_t* n, gr_run_stats* st) {
  return orc_call_peaks(x->o, peaks, n, st);
}
int gr_fetch_intervals(gr_ctx* x, int32_t which, int32_t replicate, int32_t chrom, const uint32_t** end,
                       const float** val, const float** expt, const float** ctrl, uint64_t* n) {
  return orc_fetch_intervals(x->o, which, replicate, chrom, end, val, expt, ctrl, n);
}
int gr_replicate_finish(gr_ctx* x, double frag_len, double ctrl_frag, int32_t has_ctrl, uint64_t genome_len,
                        gr_sample_stats* st) {
  return orc_replicate_finish(x->o, frag_len, ctrl_frag, has_ctrl, genome_len, st);
}
int gr_excluded_bp(gr_ctx* x, uint64_t* per_chrom) { return orc_excluded_bp(x->o, per_chrom); }
int gr_pvalues_finalize(gr_ctx* x) { return orc_pvalues_finalize(x->o); }
int gr_bh_local_hist_host(gr_ctx* x, const uint32_t** keys, const uint64_t** lens, uint64_t* n) {
  return orc_bh_local_hist(x->o, keys, lens, n);            /* the oracle's lists are host arrays anyway */
}
int gr_bh_set_global_host(gr_ctx* x, const uint32_t* keys, const uint64_t* lens, uint64_t n, uint64_t genome_len) {
  return orc_bh_set_global(x->o, keys, lens, n, genome_len);
}
/* every list is in (chromosome, start) order and a chromosome has one owner (callPeaks 986-987) */
int gr_merge_peaks(const gr_peak* const* lists, const uint64_t* counts, int32_t nlists, gr_peak* out) {
  uint64_t pos[64] = { 0 }, w = 0;
  if (nlists < 0 || nlists > 64) return GR_ERR_ARG;
  for (;;) {
    int best = -1;
    for (int i = 0; i < nlists; i++)
      if (pos[i] < counts[i] && (best < 0 || lists[i][pos[i]].chrom < lists[best][pos[best]].chrom)) best = i;
    if (best < 0) break;
    const int32_t c = lists[best][pos[best]].chrom;
    while (pos[best] < counts[best] && lists[best][pos[best]].chrom == c) out[w++] = lists[best][pos[best]++];
  }
  return GR_OK;
}
void* gr_pinned_alloc(size_t bytes) { return malloc(bytes); }
void gr_pinned_free(void* p) { free(p); }
const char* gr_last_error_detail(const gr_ctx* x) { (void)x; return ""; }
const char* gr_strerror(int status) {
  static const char* text[] = {
    "ok", "bad argument or call order", "CUDA failure", "Cannot allocate memory",
    ": read aligned beyond reference end", "Experimental sample has no analyzable fragments",
    "No analyzable genome (length=0)", "Invalid pileup value (< 0)",
    "Disallowed number of alignments", "interval on an unknown or unowned chromosome",
    "Invalid df in pchisq()", "Genome length does not match p-value length",
    "no CUDA device available",
    "int16 saturation of the delta counters beyond what the path reproduces"
  };
  return status < 0 || status > GR_ERR_SATURATED ? "Unknown error" : text[status];
}
