/* gb_peaksonly.c -- -P: peaks from an already written -f log (findPeaksOnly 5243, callPeaksLog
 * 1277-1470, getIdx 1225, loadBDG 1253 of Genrich.c).
 *
 * The reference streams the log and calls peaks on the text's own -log(p) / -log(q) columns (no
 * pileups, no Benjamini-Hochberg pass).  Here the host parses the log into per-chromosome arrays
 * (interval end, p, q) -- applying -e and the "new -E" regions with the reference's own state
 * machine, so that its quirks survive (an NA record is skipped BEFORE the region bookkeeping,
 * 1364-1374) -- and the peak scan itself runs on the device (gr_load_pvalues + gr_call_peaks, K8),
 * like every other peak call of this program.  A chromosome is a run of records with one name
 * (1325: a name that comes back later is a new run, as in the reference).
 *
 * Difference from the reference, on malformed input only: records must tile a chromosome run
 * (start == previous end, as every -f log does); a hole is bridged by a non-significant filler
 * interval, a step backwards is an error.  The reference ignores the hole (no maxGap test is made
 * for what is not there). */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gb_host.h"

typedef struct { uint32_t* v; int n, cap; } U32List;
static void u32_push(U32List* l, uint32_t x) {
  if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 16; l->v = (uint32_t*)gb_realloc(l->v, (size_t)l->cap * sizeof(uint32_t)); }
  l->v[l->n++] = x;
}

typedef struct { char* name; uint32_t pos[2]; } XBed;

static bool in_name_list(const char* name, const char* list) {       /* checkChrom 1213 */
  if (!list) return false;
  const size_t n = strlen(name);
  for (const char* p = list; *p;) {
    const char* q = strchr(p, ',');
    const size_t m = q ? (size_t)(q - p) : strlen(p);
    if (m == n && !strncmp(p, name, n)) return true;
    if (!q) break;
    p = q + 1;
  }
  return false;
}

static int cmp_bed(const void* a, const void* b) {
  const uint32_t x = ((const uint32_t*)a)[0], y = ((const uint32_t*)b)[0];
  return x < y ? -1 : x > y;
}
/* saveXBed 1144-1205 with len = UINT32_MAX (1338: "cannot check validity of coordinates"):
 * the regions of one chromosome sorted by start, overlapping / touching ones fused */
static void merged_bed(const XBed* xb, int nxb, const char* chr, U32List* out) {
  out->n = 0;
  for (int i = 0; i < nxb; i++)
    if (!strcmp(xb[i].name, chr)) { u32_push(out, xb[i].pos[0]); u32_push(out, xb[i].pos[1]); }
  if (!out->n) return;
  qsort(out->v, (size_t)out->n / 2, 2 * sizeof(uint32_t), cmp_bed);
  int w = 0;
  for (int i = 2; i < out->n; i += 2) {
    if (out->v[i] <= out->v[w + 1]) { if (out->v[i + 1] > out->v[w + 1]) out->v[w + 1] = out->v[i + 1]; }
    else { w += 2; out->v[w] = out->v[i]; out->v[w + 1] = out->v[i + 1]; }
  }
  out->n = w + 2;
}

typedef struct {
  uint32_t* end; float* p; float* q; uint64_t n, cap;
} Ivals;
static void iv_push(Ivals* a, uint32_t end, float p, float q) {
  if (a->n == a->cap) {
    a->cap = a->cap ? 2 * a->cap : (1u << 16);
    a->end = (uint32_t*)gb_realloc(a->end, a->cap * sizeof(uint32_t));
    a->p = (float*)gb_realloc(a->p, a->cap * sizeof(float));
    a->q = (float*)gb_realloc(a->q, a->cap * sizeof(float));
  }
  a->end[a->n] = end; a->p[a->n] = p; a->q[a->n] = q; a->n++;
}

int gb_peaks_only(const HOpts* o, char* xfile, float thr) {
  /* -E regions (loadBED 5181) */
  XBed* xb = NULL;
  int nxb = 0, capxb = 0;
  static char line[65536];
  if (xfile) {
    char* save_list;
    for (char* fname = strtok_r(xfile, ",", &save_list); fname; fname = strtok_r(NULL, ",", &save_list)) {
      HIn in;
      gb_in_open(&in, fname);
      while (gb_in_gets(&in, line, sizeof line)) {
        char copy[256];
        snprintf(copy, sizeof copy, "%.255s", line);
        char* sp;
        char* name = strtok_r(line, "\t", &sp);
        if (!name) gb_die(copy, ": poorly formatted BED record");
        int pos[2];
        for (int i = 0; i < 2; i++) {
          char* val = strtok_r(NULL, i ? "\t\n" : "\t", &sp);
          if (!val) gb_die(copy, ": poorly formatted BED record");
          pos[i] = gb_parse_int(val);
        }
        if (pos[1] <= pos[0] || pos[0] < 0 || pos[1] < 0) {
          char msg[512];
          snprintf(msg, sizeof msg, "%s, %d - %d", name, pos[0], pos[1]);
          gb_die(msg, ": poorly formatted BED record");
        }
        if (nxb == capxb) { capxb = capxb ? 2 * capxb : 64; xb = (XBed*)gb_realloc(xb, (size_t)capxb * sizeof(XBed)); }
        xb[nxb].name = strdup(name); xb[nxb].pos[0] = (uint32_t)pos[0]; xb[nxb].pos[1] = (uint32_t)pos[1];
        nxb++;
      }
      gb_in_close(&in, fname);
    }
  }

  HIn in;
  gb_in_open(&in, o->log_file);
  if (o->verbose) fprintf(stderr, "Peak-calling from log file: %s\n", o->log_file);
  /* getIdx 1225: the LAST header fields that start with -log(p) / -log(q) */
  if (!gb_in_gets(&in, line, sizeof line)) gb_die("<header>", ": cannot find field in header of bedgraph-ish log file");
  int idxP = -1, idxQ = -1;
  {
    int i = 0;
    char* sp;
    for (char* f = strtok_r(line, "\t\n", &sp); f; f = strtok_r(NULL, "\t\n", &sp), i++) {
      if (!strncmp(f, "-log(p)", 7)) idxP = i;
      else if (!strncmp(f, "-log(q)", 7)) idxQ = i;
    }
  }
  if (idxP == -1) gb_die("-log(p)", ": cannot find field in header of bedgraph-ish log file");
  if (o->qval_opt && idxQ == -1) gb_die("-log(q)", ": cannot find field in header of bedgraph-ish log file");
  const int idx = o->qval_opt ? idxQ : idxP;

  /* chromosome runs */
  char** names = NULL;
  uint64_t* cstart = NULL;               /* first interval of every run, +1 */
  uint32_t* clen = NULL;
  int nchr = 0, capchr = 0;
  Ivals iv = { NULL, NULL, NULL, 0, 0 };
  U32List bed = { NULL, 0, 0 };
  int bedIdx = 0;
  uint32_t bedPos = UINT32_MAX;
  bool save = true, warn = false, skip = false;
  uint64_t genomeLen = o->genome_len;
  const bool genomeOpt = o->genome_len == 0;
  char prev[65536];
  prev[0] = '\0';
  uint32_t last_end = 0;
  while (gb_in_gets(&in, line, sizeof line)) {
    /* loadBDG 1253 */
    char *chr = NULL, *pStat = NULL, *qStat = NULL, *sp;
    uint32_t start = 0, end = 0;
    char* f = strtok_r(line, "\t\n", &sp);
    for (int i = 0; i <= idx; i++) {
      if (!f) gb_die("", "Poorly formatted bedgraph-ish log record");
      if (i == 0) chr = f;
      else if (i == 1) start = (uint32_t)gb_parse_int(f);
      else if (i == 2) end = (uint32_t)gb_parse_int(f);
      else if (i == idxP) pStat = f;
      else if (i == idxQ) qStat = f;
      f = strtok_r(NULL, "\t\n", &sp);
    }
    if (strcmp(prev, chr)) {                                         /* 1325-1353 */
      skip = in_name_list(chr, o->xchrom);
      if (o->verbose && skip) {
        fprintf(stderr, "Warning! Skipping chromosome %s --\n  ", chr);
        fprintf(stderr, "Reads aligning to it were used in the background");
        fprintf(stderr, " pileup calculation,\n  and its length was included");
        fprintf(stderr, " in the genome length %scalculation\n", o->qval_opt ? "(and q-value) " : "");
      }
      bed.n = 0;
      if (!skip) {
        merged_bed(xb, nxb, chr, &bed);
        bedIdx = 0;
        bedPos = bedIdx < bed.n ? bed.v[bedIdx] : UINT32_MAX;
        save = true;
        if (nchr == capchr) {
          capchr = capchr ? 2 * capchr : 64;
          names = (char**)gb_realloc(names, (size_t)capchr * sizeof(char*));
          cstart = (uint64_t*)gb_realloc(cstart, ((size_t)capchr + 1) * sizeof(uint64_t));
          clen = (uint32_t*)gb_realloc(clen, (size_t)capchr * sizeof(uint32_t));
        }
        names[nchr] = strdup(chr);
        cstart[nchr] = iv.n;
        clen[nchr] = 0;
        nchr++;
        last_end = 0;
      }
      snprintf(prev, sizeof prev, "%s", chr);
    }
    if (skip) continue;
    if (end <= start || start < last_end) gb_die("", "Poorly formatted bedgraph-ish log record");
    if (start > last_end) iv_push(&iv, start, 0.0f, 0.0f);           /* hole: non-significant filler (see the header) */
    last_end = end;
    clen[nchr - 1] = end;
    const char* stat = o->qval_opt ? qStat : pStat;
    if (!strcmp(stat, "NA")) { iv_push(&iv, end, GR_SKIP, GR_SKIP); continue; }      /* 1364-1374 */
    const float pq = gb_parse_float(stat);
    const float pv = o->qval_opt ? gb_parse_float(pStat) : pq;
    const float qv = o->qval_opt ? pq : GR_SKIP;
    if (bedPos == start) {                                           /* 1379-1392: the interval starts at a region boundary */
      save = !save;
      bedIdx++;
      bedPos = bedIdx < bed.n ? bed.v[bedIdx] : UINT32_MAX;
    }
    uint32_t subStart = start;
    while (bedPos > start && bedPos < end) {                         /* 1397-1428: region boundaries inside it */
      if (save) {
        iv_push(&iv, bedPos, pv, qv);                                /* kept piece; the excluded piece behind it closes the peak */
        if (genomeOpt) genomeLen += bedPos - subStart;
      } else {
        iv_push(&iv, bedPos, GR_SKIP, GR_SKIP);
        warn = true;
      }
      subStart = bedPos;
      save = !save;
      bedIdx++;
      bedPos = bedIdx < bed.n ? bed.v[bedIdx] : UINT32_MAX;
    }
    if (!save) { warn = true; iv_push(&iv, end, GR_SKIP, GR_SKIP); continue; }       /* 1429-1432 */
    if (genomeOpt) genomeLen += end - subStart;
    iv_push(&iv, end, pv, qv);
  }
  gb_in_close(&in, o->log_file);

  /* the device: one context over the runs, the peak scan on the loaded values */
  const gr_peak* peaks = NULL;
  uint64_t npk = 0;
  gr_run_stats rs;
  memset(&rs, 0, sizeof rs);
  gr_ctx* ctx = NULL;
  if (nchr) {
    cstart[nchr] = iv.n;
    gr_chrom* gc = (gr_chrom*)gb_alloc((size_t)nchr * sizeof(gr_chrom));
    for (int i = 0; i < nchr; i++) { gc[i].len = clen[i] ? clen[i] : 1; gc[i].skip = 0; gc[i].owned = 1; gc[i].reserved = 0; }
    gr_params par;
    par.min_pqval = thr; par.qval_opt = o->qval_opt; par.min_auc = o->min_auc; par.min_len = o->min_len;
    par.max_gap = o->max_gap; par.keep_pileups = 0; par.genome_len = o->genome_len;
    int rc = gr_create(&ctx, gc, nchr, &par, o->device);
    if (rc) gb_die(gr_strerror(rc), "");
    rc = gr_load_pvalues(ctx, cstart, iv.end, iv.p, o->qval_opt ? iv.q : NULL, iv.n);
    if (rc) gb_die(gr_strerror(rc), "");
    rc = gr_call_peaks(ctx, &peaks, &npk, &rs);
    if (rc) gb_die(gr_strerror(rc), "");
    free(gc);
  }
  HOut out;
  gb_out_open(&out, o->out_file, o->gz_out);
  for (uint64_t i = 0; i < npk; i++) {                               /* printPeak 885-909 */
    const gr_peak* p = &peaks[i];
    unsigned score = (unsigned)(1000.0f * p->auc / (p->end - p->start) + 0.5f);
    if (score > 1000) score = 1000;
    gb_out_printf(&out, "%s\t%ld\t%ld\tpeak_%d\t%d\t.\t%f\t%f", names[p->chrom], (long)p->start, (long)p->end, (int)i,
                  score, p->auc, p->pval);
    if (p->qval == GR_SKIP) gb_out_printf(&out, "\t-1\t%d\n", p->summit);
    else gb_out_printf(&out, "\t%f\t%d\n", p->qval, p->summit);
  }
  gb_out_close(&out, o->out_file);
  if (o->verbose) {                                                  /* 1447-1467 */
    if (warn) {
      fprintf(stderr, "Warning! Skipping given BED regions --\n  ");
      fprintf(stderr, "Reads aligning to them were used in the background");
      fprintf(stderr, " pileup calculation,\n  and the lengths were included");
      fprintf(stderr, " in the genome length %scalculation\n", o->qval_opt ? "(and q-value) " : "");
    }
    fprintf(stderr, "Peak-calling parameters:\n");
    fprintf(stderr, "  Genome length: %ldbp\n", (long)genomeLen);
    fprintf(stderr, "  Significance threshold: -log(%c) > %.3f\n", o->qval_opt ? 'q' : 'p', thr);
    fprintf(stderr, "  Min. AUC: %.3f\n", o->min_auc);
    if (o->min_len) fprintf(stderr, "  Min. peak length: %dbp\n", o->min_len);
    fprintf(stderr, "  Max. gap between sites: %dbp\n", o->max_gap);
    fprintf(stderr, "Peaks identified: %d (%ldbp)\n", (int)npk, (long)rs.peak_bp);
  }
  if (ctx) gr_destroy(ctx);
  return EXIT_SUCCESS;
}
