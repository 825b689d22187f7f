/* oracle/ref_funcs.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Exposes a handful of the reference's own functions through a flat C surface
 * so that tests can pin the restatement in genrich_oracle.c function by
 * function.  The reference translation unit is #included from where it lies
 * (REF_SRC, set by oracle/Makefile to /root/reference/Genrich.c) with `main`
 * renamed; no reference source is copied into this repository.  The resulting
 * library lands in oracle/_ref/ (git-ignored).
 */
#define main genrich_reference_main
#include REF_SRC
#undef main

/* calcPval, Genrich.c:1628 */
float ref_calcPval(float expt, float ctrl) { return calcPval(expt, ctrl); }

/* pchisq, Genrich.c:555 (returns -log10 p) */
double ref_pchisq(double x, int df) { return pchisq(x, df); }

/* multPval, Genrich.c:567, over n replicate values (SKIP = -1 excluded) */
float ref_multPval(const float* vals, int n) {
  Pileup pl[n];
  Pileup* pp[n];
  uint32_t idx[n];
  float cov[n];
  for (int i = 0; i < n; i++) {
    cov[i] = vals[i];
    pl[i].cov = &cov[i];
    pl[i].end = NULL;
    pp[i] = &pl[i];
    idx[i] = 0;
  }
  return multPval(pp, n, idx);
}

/* Drive addFrac/subFrac (Genrich.c:2311, 2412) on one diff cell. */
void ref_diff_add(int16_t* cov, uint8_t* frac, int count, int sign) {
  if (count == 1) {
    *cov += sign > 0 ? 1 : -1;
    return;
  }
  if (sign > 0) addFrac(cov, frac, (uint8_t)count);
  else subFrac(cov, frac, (uint8_t)count);
}

/* updateVal, Genrich.c:1915 */
float ref_updateVal(int16_t dCov, uint8_t dFrac, int32_t* cov, uint8_t* frac) {
  return updateVal(dCov, dFrac, cov, frac);
}

/* saveQval path, Genrich.c:352: one chromosome, p-values `p` with interval
 * ends `end` (n intervals); writes q-values into q. */
void ref_computeQval(const float* p, const uint32_t* end, uint32_t n,
                     uint64_t genomeLen, float* q) {
  Chrom c;
  memset(&c, 0, sizeof c);
  c.name = "x";
  c.len = end[n - 1];
  c.skip = false;
  c.save = true;
  Pileup pv;
  pv.end = (uint32_t*)end;
  pv.cov = (float*)p;
  Pileup* pvp = &pv;
  c.pval = &pvp;
  uint32_t plen = n;
  c.pvalLen = &plen;
  c.sample = 1;
  computeQval(&c, 1, genomeLen, false, 0, false);
  for (uint32_t i = 0; i < n; i++) q[i] = c.qval->cov[i];
  free(c.qval->end);
  free(c.qval->cov);
  free(c.qval);
}
