/* gb_cli.c -- genrich-b200: the Genrich command line over libgenrich_cuda.so.
 *
 * Same options, inputs, outputs, messages and exit status as the reference
 * (getArgs 5718-5827, runProgram 5386-5695, usage 34-71); the per-chromosome
 * loops of runProgram/findPeaks are replaced by calls into the C-ABI of
 * include/genrich_cuda.h.  -P (peaks from a log file) lives in gb_peaksonly.c.
 */
#include "gb_host.h"
#include <float.h>
#include <getopt.h>
#include <unistd.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

void* gr_pinned_alloc(size_t bytes);     /* exported by libgenrich_cuda.so */
void gr_pinned_free(void* p);

#define GB_OPTIONS "ht:c:o:f:k:b:zyw:xjd:De:E:m:s:p:q:a:l:g:rR:XPSL:vVG:"

static struct option gb_long[] = {
  {"help", no_argument, NULL, 'h'}, {"verbose", no_argument, NULL, 'v'},
  {"version", no_argument, NULL, 'V'}, {"gpu", required_argument, NULL, 'G'},
  {"threads", required_argument, NULL, 'T'}, {"gpus", required_argument, NULL, 'N'}, {0, 0, 0, 0}
};

static void usage(void) {
  fprintf(stderr, "Usage: ./genrich-b200  -t <file>  -o <file>  [optional arguments]\n");
  fprintf(stderr, "Required arguments:\n");
  fprintf(stderr, "  -t  <file>       Input SAM/BAM file(s) for experimental sample(s)\n");
  fprintf(stderr, "  -o  <file>       Output peak file (in ENCODE narrowPeak format)\n");
  fprintf(stderr, "Optional I/O arguments:\n");
  fprintf(stderr, "  -c  <file>       Input SAM/BAM file(s) for control sample(s)\n");
  fprintf(stderr, "  -f  <file>       Output bedgraph-ish file for p/q values\n");
  fprintf(stderr, "  -k  <file>       Output bedgraph-ish file for pileups and p-values\n");
  fprintf(stderr, "  -b  <file>       Output BED file for reads/fragments/intervals\n");
  fprintf(stderr, "  -R  <file>       Output file for PCR duplicates (only with -r)\n");
  fprintf(stderr, "Filtering options:\n");
  fprintf(stderr, "  -r               Remove PCR duplicates\n");
  fprintf(stderr, "  -e  <arg>        Comma-separated list of chromosomes to exclude\n");
  fprintf(stderr, "  -E  <file>       Input BED file(s) of genomic regions to exclude\n");
  fprintf(stderr, "  -m  <int>        Minimum MAPQ to keep an alignment (def. 0)\n");
  fprintf(stderr, "  -s  <float>      Keep sec alns with AS >= bestAS - <float> (def. 0)\n");
  fprintf(stderr, "  -y               Keep unpaired alignments (def. false)\n");
  fprintf(stderr, "  -w  <int>        Keep unpaired alns, lengths changed to <int>\n");
  fprintf(stderr, "  -x               Keep unpaired alns, lengths changed to paired avg\n");
  fprintf(stderr, "Options for ATAC-seq:\n");
  fprintf(stderr, "  -j               Use ATAC-seq mode (def. false)\n");
  fprintf(stderr, "  -d  <int>        Expand cut sites to <int> bp (def. 100)\n");
  fprintf(stderr, "  -D               Skip Tn5 adjustments of cut sites (def. false)\n");
  fprintf(stderr, "Options for peak-calling:\n");
  fprintf(stderr, "  -p  <float>      Maximum p-value (def. 0.01)\n");
  fprintf(stderr, "  -q  <float>      Maximum q-value (FDR-adjusted p-value; def. 1)\n");
  fprintf(stderr, "  -a  <float>      Minimum AUC for a peak (def. 200.0)\n");
  fprintf(stderr, "  -l  <int>        Minimum length of a peak (def. 0)\n");
  fprintf(stderr, "  -g  <int>        Maximum distance between signif. sites (def. 100)\n");
  fprintf(stderr, "Other options:\n");
  fprintf(stderr, "  -X               Skip peak-calling\n");
  fprintf(stderr, "  -z               Option to gzip-compress output(s)\n");
  fprintf(stderr, "  -v               Option to print status updates/counts to stderr\n");
  fprintf(stderr, "  --gpu <int>      CUDA device to use (def. 0)\n");
  fprintf(stderr, "  --gpus <int>     Number of devices, starting at --gpu; chromosomes are sharded over them (def. 1)\n");
  fprintf(stderr, "  --threads <int>  Host threads decoding a plain SAM file (def. all cores, at most 16)\n");
  exit(EXIT_FAILURE);
}

/* logCounts 5295-5374 */
static void chk(gr_ctx* ctx, int rc, const char* what);
/* saveInterval 2558-2573: the engine drops what the reference's int16 counters would have made it drop
 * (gr_sample_skipped).  Under -v the reference names every such alignment as it meets it; here a second,
 * sequential pass over the file, which pushes nothing, finds the names of the listed records. */
static void report_skipped(HDecode* d, gr_ctx** ctxs, int nctx, int is_ctrl, const char* path) {
  uint64_t total = 0;
  HLookup lk;
  lk.list = (const uint64_t**)gb_alloc((size_t)nctx * sizeof(uint64_t*));
  lk.n = (uint64_t*)calloc((size_t)nctx, sizeof(uint64_t));
  lk.pos = (uint64_t*)calloc((size_t)nctx, sizeof(uint64_t));
  lk.arrival = (uint64_t*)calloc((size_t)nctx, sizeof(uint64_t));
  uint64_t** copy = (uint64_t**)calloc((size_t)nctx, sizeof(uint64_t*));
  for (int k = 0; k < nctx; k++) {
    uint64_t a = 0, b = 0, nl = 0;
    const uint64_t* l = NULL;
    chk(ctxs[k], gr_sample_skipped(ctxs[k], is_ctrl, &a, &b, &l, &nl), "gr_sample_skipped");
    total += a + b;
    copy[k] = (uint64_t*)gb_alloc((nl ? nl : 1) * sizeof(uint64_t));      /* the library's array lives until its next call */
    if (nl) memcpy(copy[k], l, nl * sizeof(uint64_t));
    lk.list[k] = copy[k]; lk.n[k] = nl;
  }
  /* Two corners are refused rather than computed differently from the reference: with -x the reference's average
   * fragment length leaves out the fragments it had dropped by then (3174), which is known here only after the
   * unpaired alignments were extended by that average; and -b lists the fragments as they were decoded, the
   * dropped ones included. */
  if (total && (d->opt->avg_ext_opt || d->opt->bed_file))
    gb_die("", "More than 32767 fragments start or end on one base (the reference drops fragments there): "
               "not supported together with -x or -b");
  if (total && d->opt->verbose) {
    HDecode* d2 = (HDecode*)calloc(1, sizeof(HDecode));
    HIvBuf* bufs = (HIvBuf*)calloc((size_t)nctx, sizeof(HIvBuf));
    if (!d2 || !bufs) gb_die("", "Cannot allocate memory");
    d2->opt = d->opt; d2->tab = d->tab; d2->nctx = nctx; d2->ctxs = ctxs; d2->owner = d->owner; d2->bufs = bufs;
    d2->ctrl = d->ctrl; d2->sample = d->sample; d2->lookup = &lk;
    gb_decode_file(d2, path);
    d->cnt.total_len = d2->cnt.total_len;                 /* without the dropped fragments, like the reference's (3174) */
    free(d2->unp); free(d2->rd_pr.r); free(d2->rd_dc.r); free(d2->rd_sn.r);
    free(bufs); free(d2);
  }
  for (int k = 0; k < nctx; k++) free(copy[k]);
  free(copy); free(lk.list); free(lk.n); free(lk.pos); free(lk.arrival);
}

static void log_counts(const HDecode* d, bool bam) {
  const HCounts* c = &d->cnt;
  const HOpts* o = d->opt;
  if (c->err_count > GB_MAX_ALNS)
    fprintf(stderr, "(another %ld warning messages suppressed)\n", (long)(c->err_count - GB_MAX_ALNS));
  const double avg = c->paired_pr ? c->total_len / c->paired_pr : 0.0;
  fprintf(stderr, "  %s records analyzed: %11ld\n", bam ? "BAM" : "SAM", (long)c->count);
  if (c->unmapped) fprintf(stderr, "    Unmapped:           %11ld\n", (long)c->unmapped);
  if (c->supp) fprintf(stderr, "    Supp./dups/lowQual: %11ld\n", (long)c->supp);
  if (c->skipped) {
    fprintf(stderr, "    To skipped refs:    %11ld\n", (long)c->skipped);
    fprintf(stderr, "      (");
    bool first = true;
    for (int i = 0; i < d->tab->n; i++)
      if (d->tab->c[i].skip || !d->tab->c[i].save) {
        fprintf(stderr, "%s%s", first ? "" : ",", d->tab->c[i].name);
        first = false;
      }
    fprintf(stderr, ")\n");
  }
  if (c->low_mapq) fprintf(stderr, "    MAPQ < %-2d:          %11ld\n", o->min_mapq, (long)c->low_mapq);
  fprintf(stderr, "    Paired alignments:  %11ld\n", (long)c->paired);
  if (c->sec_pair) fprintf(stderr, "      secondary alns:   %11ld\n", (long)c->sec_pair);
  if (c->orphan) fprintf(stderr, "      \"orphan\" alns:    %11ld\t** Warning! **\n", (long)c->orphan);
  fprintf(stderr, "    Unpaired alignments:%11ld\n", (long)c->single);
  if (c->sec_single) fprintf(stderr, "      secondary alns:   %11ld\n", (long)c->sec_single);
  if (o->dups_opt) {
    fprintf(stderr, "  PCR duplicates --\n");
    fprintf(stderr, "    Paired aln sets:    %11ld\n", (long)c->count_pr);
    fprintf(stderr, "      duplicates:       %11ld (%.1f%%)\n", (long)c->dups_pr,
            c->count_pr ? 100.0f * c->dups_pr / c->count_pr : 0.0f);
    if (o->single_opt) {
      fprintf(stderr, "    Discordant aln sets:%11ld\n", (long)c->count_dc);
      fprintf(stderr, "      duplicates:       %11ld (%.1f%%)\n", (long)c->dups_dc,
              c->count_dc ? 100.0f * c->dups_dc / c->count_dc : 0.0f);
      fprintf(stderr, "    Singleton aln sets: %11ld\n", (long)c->count_sn);
      fprintf(stderr, "      duplicates:       %11ld (%.1f%%)\n", (long)c->dups_sn,
              c->count_sn ? 100.0f * c->dups_sn / c->count_sn : 0.0f);
    }
  }
  fprintf(stderr, "  Fragments analyzed:   %11ld\n", (long)(c->single_pr + c->paired_pr));
  fprintf(stderr, "    Full fragments:     %11ld\n", (long)c->paired_pr);
  if (c->paired_pr && !o->atac_opt) fprintf(stderr, "      (avg. length: %.1fbp)\n", avg);
  if (o->single_opt) {
    fprintf(stderr, "    Half fragments:     %11ld\n", (long)c->single_pr);
    if (c->single_pr) {
      fprintf(stderr, "      (from unpaired alns");
      if (o->extend_opt) fprintf(stderr, ", extended to %dbp", o->extend);
      else if (o->avg_ext_opt && c->paired_pr) fprintf(stderr, ", extended to %dbp", (int)(avg + 0.5));
      fprintf(stderr, ")\n");
    }
  }
  if (o->atac_opt) {
    fprintf(stderr, "    ATAC-seq cut sites: %11ld\n", (long)(2 * c->paired_pr + c->single_pr));
    fprintf(stderr, "      (expanded to length %dbp)\n", o->atac_len5 + o->atac_len3);
  }
}

static void chk(gr_ctx* ctx, int rc, const char* what) {
  if (!rc) return;
  /* statuses that correspond to one of the reference's errCode cases print its text */
  if (rc == GR_ERR_EXPT || rc == GR_ERR_GENOME || rc == GR_ERR_MEM) gb_die("", gr_strerror(rc));
  char msg[1024];
  snprintf(msg, sizeof msg, "%s: %s %s", what, gr_strerror(rc), ctx ? gr_last_error_detail(ctx) : "");
  gb_die(msg, rc == GR_ERR_PILE || rc == GR_ERR_COUNT || rc == GR_ERR_DF
         ? "\n  (internal error: please open an Issue on https://github.com/jsh58/Genrich)" : "");
}

/* "-" as an input file name: the reference reads uncompressed SAM from stdin (openRead 5135-5165).
 * Here every input is read twice (the engine needs the whole chromosome table before the first
 * record), so stdin is first copied to a temporary file, which stands in for "-" from then on. */
static char* g_spool = NULL;
static void spool_remove(void) { if (g_spool) unlink(g_spool); }
static const char* real_path(const char* name) {
  if (strcmp(name, "-")) return name;
  if (g_spool) return g_spool;
  const char* dir = getenv("TMPDIR");
  if (!dir || !*dir) dir = "/tmp";
  g_spool = (char*)gb_alloc(strlen(dir) + 32);
  sprintf(g_spool, "%s/genrich-b200.XXXXXX", dir);
  const int fd = mkstemp(g_spool);
  if (fd < 0) gb_die(g_spool, ": cannot open file for writing");
  atexit(spool_remove);
  static char buf[1 << 20];
  size_t n, total = 0;
  while ((n = fread(buf, 1, sizeof buf, stdin)) > 0) {
    if (!total && n >= 2 && (unsigned char)buf[0] == 0x1F && (unsigned char)buf[1] == 0x8B)
      gb_die("", "Cannot pipe in gzip-compressed file (use zcat instead)");
    for (size_t off = 0; off < n;) {
      const ssize_t w = write(fd, buf + off, n - off);
      if (w <= 0) gb_die(g_spool, ": cannot write to file");
      off += (size_t)w;
    }
    total += n;
  }
  close(fd);
  if (!total) gb_die("-", ": cannot open file for reading");
  return g_spool;
}

/* split a list on ", " like strtok_r(.., COM, ..) at Genrich.c:5457 */
static int split_list(char* s, char*** out) {
  int n = 0;
  *out = NULL;
  if (!s) return 0;
  char* save;
  for (char* t = strtok_r(s, ", ", &save); t; t = strtok_r(NULL, ", ", &save)) {
    *out = (char**)gb_realloc(*out, (n + 1) * sizeof(char*));
    (*out)[n++] = t;
  }
  return n;
}

/* -k: printPileHeader 1680 + printPile 1697 for one replicate */
static void write_pile(gr_ctx** ctxs, const int* owner, HOut* out, const HChromTab* tab, int rep, const char* ename, const char* cname) {
  gb_out_printf(out, "# experimental file: %s; control file: %s\n", ename,
                cname && strcmp(cname, "null") ? cname : "NA");
  gb_out_printf(out, "chr\tstart\tend\texperimental\tcontrol\t-log(p)\n");
  for (int c = 0; c < tab->n; c++) {
    const uint32_t* end; const float *val, *ex, *ct; uint64_t n;
    gr_ctx* ctx = ctxs[owner[c]];                              /* the device that holds this chromosome */
    chk(ctx, gr_fetch_intervals(ctx, 2, rep, c, &end, &val, &ex, &ct, &n), "fetch");
    uint32_t start = 0;
    for (uint64_t m = 0; m < n; m++) {
      if (ct[m] == GR_SKIP)
        gb_out_printf(out, "%s\t%d\t%d\t%f\t%f\t%s\n", tab->c[c].name, start, end[m], ex[m], 0.0f, "NA");
      else
        gb_out_printf(out, "%s\t%d\t%d\t%f\t%f\t%f\n", tab->c[c].name, start, end[m], ex[m], ct[m], val[m]);
      start = end[m];
    }
  }
}

typedef struct { uint32_t* end; float *val, *ex, *ct; uint64_t n; } Arr;
static Arr fetch_copy(gr_ctx* ctx, int which, int rep, int c, bool cols) {
  const uint32_t* end; const float *val, *ex, *ct; uint64_t n;
  chk(ctx, gr_fetch_intervals(ctx, which, rep, c, &end, &val, &ex, &ct, &n), "fetch");
  Arr a = { NULL, NULL, NULL, NULL, n };
  if (!n || !end) { a.n = 0; return a; }
  a.end = (uint32_t*)gb_alloc(n * 4); memcpy(a.end, end, n * 4);
  a.val = (float*)gb_alloc(n * 4); memcpy(a.val, val, n * 4);
  if (cols && ex) {
    a.ex = (float*)gb_alloc(n * 4); memcpy(a.ex, ex, n * 4);
    a.ct = (float*)gb_alloc(n * 4); memcpy(a.ct, ct, n * 4);
  }
  return a;
}
static void arr_free(Arr* a) { free(a->end); free(a->val); free(a->ex); free(a->ct); }

/* -f: printLogHeader 674 + printInterval 770 / printIntervalN 724 as callPeaks / logIntervals emit them */
static void write_log(gr_ctx** ctxs, const int* owner, HOut* out, const HChromTab* tab, int nrep, const HOpts* o, float thr) {
  const bool sig_col = o->peaks_opt, q = o->qval_opt;
  if (nrep > 1) {
    gb_out_printf(out, "chr\tstart\tend");
    for (int i = 0; i < nrep; i++) gb_out_printf(out, "\t-log(p)_%d", i);
    gb_out_printf(out, "\t-log(p)_comb");
  } else
    gb_out_printf(out, "chr\tstart\tend\texperimental\tcontrol\t-log(p)");
  if (q) gb_out_printf(out, "\t-log(q)");
  if (sig_col) gb_out_printf(out, "\tsignif");
  gb_out_printf(out, "\n");
  for (int c = 0; c < tab->n; c++) {
    gr_ctx* ctx = ctxs[owner[c]];
    Arr fin = fetch_copy(ctx, 2, nrep > 1 ? nrep : 0, c, nrep == 1);
    if (!fin.n) continue;
    Arr qv = { 0 };
    if (q) qv = fetch_copy(ctx, 3, 0, c, false);
    Arr* reps = NULL;
    uint64_t* idx = NULL;
    if (nrep > 1) {
      reps = (Arr*)gb_alloc(nrep * sizeof(Arr));
      idx = (uint64_t*)calloc(nrep, sizeof(uint64_t));
      for (int r = 0; r < nrep; r++) reps[r] = fetch_copy(ctx, 2, r, c, false);
    }
    uint32_t start = 0;
    for (uint64_t m = 0; m < fin.n; m++) {
      const float pv = fin.val[m], qq = q ? qv.val[m] : GR_SKIP;
      const bool sig = sig_col && (q ? qq : pv) > thr;
      if (nrep == 1) {
        if (fin.ct[m] == GR_SKIP) {
          gb_out_printf(out, "%s\t%d\t%d\t%f\t%f\t%s", tab->c[c].name, start, fin.end[m], fin.ex[m], 0.0f, "NA");
          if (q) gb_out_printf(out, "\t%s", "NA");
          gb_out_printf(out, "\n");
        } else {
          gb_out_printf(out, "%s\t%d\t%d\t%f\t%f\t%f", tab->c[c].name, start, fin.end[m], fin.ex[m], fin.ct[m], pv);
          if (q) gb_out_printf(out, "\t%f", qq);
          gb_out_printf(out, "%s\n", sig ? "\t*" : "");
        }
      } else {
        gb_out_printf(out, "%s\t%d\t%d", tab->c[c].name, start, fin.end[m]);
        for (int r = 0; r < nrep; r++) {
          if (!reps[r].n || reps[r].val[idx[r]] == GR_SKIP) gb_out_printf(out, "\t%s", "NA");
          else gb_out_printf(out, "\t%f", reps[r].val[idx[r]]);
        }
        if (pv == GR_SKIP) {
          gb_out_printf(out, "\t%s", "NA");
          if (q) gb_out_printf(out, "\t%s", "NA");
        } else {
          gb_out_printf(out, "\t%f", pv);
          if (q) gb_out_printf(out, "\t%f", qq);
        }
        gb_out_printf(out, "%s\n", sig ? "\t*" : "");
        for (int r = 0; r < nrep; r++)
          if (reps[r].n && reps[r].end[idx[r]] == fin.end[m]) idx[r]++;
      }
      start = fin.end[m];
    }
    arr_free(&fin);
    if (q) arr_free(&qv);
    if (reps) { for (int r = 0; r < nrep; r++) arr_free(&reps[r]); free(reps); free(idx); }
  }
}

/* -E: loadBED 5187-5240 (comma-separated BED files, plain or gzip; start < end, both >= 0) and
 * the -v warnings of saveXBed 1151-1192.  Records of unknown references are ignored, as in the
 * reference (saveXBed only looks for the names of the chromosomes it knows); sorting, clamping
 * and merging happen in the library (gr_set_exclusions), exactly as saveXBed does them. */
/* The reference prints them from saveXBed, i.e. when a chromosome is first met in a file header (saveChrom 4265):
 * behind the "Processing ... file" line of the file that introduces the chromosome, chromosomes in header order,
 * records in BED order.  The table is complete before the first file is processed here, so the texts are kept
 * (bed_warnings_flush prints those of the chromosomes a file introduced). */
typedef struct { int chrom, kind; uint32_t start; size_t seq; char* text; } BedWarn;   /* kind 0: ignored, 1: edited */
static BedWarn* g_bed_warn = NULL;
static size_t g_bed_warn_n = 0, g_bed_warn_cap = 0;
static void bed_warn_keep(int chrom, int kind, uint32_t start, const char* text) {
  if (g_bed_warn_n == g_bed_warn_cap) {
    g_bed_warn_cap = g_bed_warn_cap ? 2 * g_bed_warn_cap : 64;
    g_bed_warn = (BedWarn*)gb_realloc(g_bed_warn, g_bed_warn_cap * sizeof *g_bed_warn);
  }
  BedWarn* w = &g_bed_warn[g_bed_warn_n];
  w->chrom = chrom; w->kind = kind; w->start = start; w->seq = g_bed_warn_n; w->text = strdup(text);
  g_bed_warn_n++;
}
/* saveXBed's order within a chromosome: the "ignored" records as the BED file lists them (first loop, 1153-1161);
 * then the "edited" ones while the list -- kept sorted by start, a record going IN FRONT of one with the same start
 * (1164-1167) -- is merged (1181-1189) */
static int bed_warn_cmp(const void* pa, const void* pb) {
  const BedWarn* a = (const BedWarn*)pa;
  const BedWarn* b = (const BedWarn*)pb;
  if (a->kind != b->kind) return a->kind - b->kind;
  if (a->kind == 0) return a->seq < b->seq ? -1 : 1;
  if (a->start != b->start) return a->start < b->start ? -1 : 1;
  return a->seq > b->seq ? -1 : 1;
}
static void bed_warnings_flush(const int* first_file, int nchrom, int file_ordinal) {
  for (int c = 0; c < nchrom; c++) {
    if (first_file[c] != file_ordinal) continue;
    size_t n = 0;
    for (size_t i = 0; i < g_bed_warn_n; i++) n += g_bed_warn[i].chrom == c && g_bed_warn[i].text;
    if (!n) continue;
    BedWarn* sel = (BedWarn*)gb_alloc(n * sizeof *sel);
    n = 0;
    for (size_t i = 0; i < g_bed_warn_n; i++)
      if (g_bed_warn[i].chrom == c && g_bed_warn[i].text) { sel[n++] = g_bed_warn[i]; g_bed_warn[i].text = NULL; }
    qsort(sel, n, sizeof *sel, bed_warn_cmp);
    for (size_t i = 0; i < n; i++) { fputs(sel[i].text, stderr); free(sel[i].text); }
    free(sel);
  }
}

static void load_exclusions(gr_ctx** ctxs, int nctx, char* xfile, const HChromTab* tab, bool verbose) {
  int32_t* chrom = NULL;
  uint32_t *start = NULL, *end = NULL;
  size_t n = 0, cap = 0;
  static char line[65536];
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
      const int c = gb_chrom_find(tab, name);
      if (c < 0) continue;
      const uint32_t len = tab->c[c].len;
      /* (saveChrom 4264: a chromosome excluded with -e has no region list, hence no warnings) */
      if (verbose && !tab->c[c].skip && (uint32_t)pos[0] >= len) {
        char w[1024];
        snprintf(w, sizeof w, "Warning! BED interval (%s, %d - %d) ignored\n  - located off end of reference %s (length %d)\n",
                 name, pos[0], pos[1], name, (int)len);
        bed_warn_keep(c, 0, (uint32_t)pos[0], w);
      } else if (verbose && !tab->c[c].skip && (uint32_t)pos[1] > len) {
        char w[1024];
        snprintf(w, sizeof w, "Warning! BED interval (%s, %d - %d) extends past end of ref.\n  - edited to (%s, %d - %d)\n",
                 name, pos[0], pos[1], name, pos[0], (int)len);
        bed_warn_keep(c, 1, (uint32_t)pos[0], w);
      }
      if (n == cap) {
        cap = cap ? 2 * cap : 1024;
        chrom = (int32_t*)gb_realloc(chrom, cap * sizeof *chrom);
        start = (uint32_t*)gb_realloc(start, cap * sizeof *start);
        end = (uint32_t*)gb_realloc(end, cap * sizeof *end);
      }
      chrom[n] = c; start[n] = (uint32_t)pos[0]; end[n] = (uint32_t)pos[1];
      n++;
    }
    gb_in_close(&in, fname);
  }
  for (int k = 0; k < nctx; k++)                               /* every device gets every region */
    chk(ctxs[k], gr_set_exclusions(ctxs[k], chrom, start, end, n), "gr_set_exclusions");
  free(chrom); free(start); free(end);
}

int main(int argc, char** argv) {
  HOpts o;
  memset(&o, 0, sizeof o);
  o.min_len = 0; o.max_gap = 100; o.atac_len5 = 100; o.pqvalue = 0.01f; o.min_auc = 200.0f;
  o.atac_adj = true; o.peaks_opt = true; o.sort_opt = true;
  bool peaks_only = false;
  char* xfile = NULL;
  int c;
  while ((c = getopt_long(argc, argv, GB_OPTIONS, gb_long, NULL)) != -1)
    switch (c) {
      case 't': o.in_files = optarg; break;
      case 'c': o.ctrl_files = optarg; break;
      case 'o': o.out_file = optarg; break;
      case 'f': o.log_file = optarg; break;
      case 'k': o.pile_file = optarg; break;
      case 'b': o.bed_file = optarg; break;
      case 'z': o.gz_out = true; break;
      case 'y': o.single_opt = true; break;
      case 'w': o.extend = gb_parse_int(optarg); o.extend_opt = true; break;
      case 'x': o.avg_ext_opt = true; break;
      case 'j': o.atac_opt = true; break;
      case 'd': o.atac_len5 = gb_parse_int(optarg); break;
      case 'D': o.atac_adj = false; break;
      case 'e': o.xchrom = optarg; break;
      case 'E': xfile = optarg; break;
      case 'm': o.min_mapq = gb_parse_int(optarg); break;
      case 's': o.as_diff = gb_parse_float(optarg); break;
      case 'p': o.pqvalue = gb_parse_float(optarg); break;
      case 'q': o.pqvalue = gb_parse_float(optarg); o.qval_opt = true; break;
      case 'a': o.min_auc = gb_parse_float(optarg); break;
      case 'l': o.min_len = gb_parse_int(optarg); break;
      case 'g': o.max_gap = gb_parse_int(optarg); break;
      case 'r': o.dups_opt = true; break;
      case 'R': o.dups_file = optarg; break;
      case 'X': o.peaks_opt = false; break;
      case 'P': peaks_only = true; break;
      case 'S': o.sort_opt = false; break;
      case 'L': { char* e; o.genome_len = (uint64_t)strtol(optarg, &e, 10); if (*e) gb_die(optarg, ": cannot convert to int"); break; }
      case 'v': o.verbose = true; break;
      case 'V': fprintf(stderr, "genrich-b200, version %s\n", GB_VERSION); exit(EXIT_FAILURE);
      case 'G': o.device = gb_parse_int(optarg); break;
      case 'T': o.threads = gb_parse_int(optarg); break;
      case 'N': o.gpus = gb_parse_int(optarg); break;
      case 'h': usage(); break;
      default: exit(EXIT_FAILURE);
    }
  if (optind < argc) gb_die(argv[optind], ": unknown command-line argument");
  if (o.threads <= 0) {
    long nc = sysconf(_SC_NPROCESSORS_ONLN);
    o.threads = nc < 1 ? 1 : nc > 16 ? 16 : (int)nc;
  }
  if ((o.peaks_opt && !o.out_file) || (peaks_only && !o.log_file) || (!peaks_only && !o.in_files)) {   /* 5776-5778 */
    fprintf(stderr, "Error! Need input/output files\n");
    usage();
  }
  if (o.avg_ext_opt) { o.single_opt = true; o.extend_opt = false; }
  if (o.extend_opt) { o.single_opt = true; if (o.extend <= 0) gb_die("", "Extension length must be > 0"); }
  if (o.atac_opt) {
    o.avg_ext_opt = o.extend_opt = false;
    if (o.atac_len5 <= 0) gb_die("", "ATAC-seq interval length must be > 0");
    o.atac_len3 = (int)(o.atac_len5 / 2.0f + 0.5f);
    o.atac_len5 /= 2;
  }
  if (o.min_len < 0) gb_die("", "Minimum peak length must be >= 0");
  if (o.min_auc < 0.0f) gb_die("", "Minimum AUC must be >= 0.0");
  if (o.as_diff < 0.0f) gb_die("", "Secondary alignment score threshold must be >= 0.0");
  if (o.pqvalue <= 0.0f || o.pqvalue > 1.0f) gb_die("", "p-/q-value must be in (0,1]");
  const float thr = -log10f(o.pqvalue);                       /* 5817 */
  if (peaks_only) return gb_peaks_only(&o, xfile, thr);       /* -P: runProgram 5398-5403 */

  /* file lists */
  char **tf, **cf;
  const int nt = split_list(o.in_files, &tf);
  const int ncf = split_list(o.ctrl_files, &cf);
  if (!nt) { fprintf(stderr, "Error! Need input/output files\n"); usage(); }

  /* the engine needs the whole chromosome table first: header-only pass, in the
   * order the reference meets the files (t0, c0, t1, c1, ...) */
  HChromTab tab = { NULL, 0 };
  int* first_file = NULL;                                  /* per chromosome: the file (2 r + is_ctrl) that introduced it */
  for (int r = 0; r < nt; r++)
    for (int s = 0; s < 2; s++) {
      if (s && !(r < ncf && strcmp(cf[r], "null"))) continue;
      const int before = tab.n;
      gb_scan_header(real_path(s ? cf[r] : tf[r]), &tab, s != 0, &o);
      first_file = (int*)gb_realloc(first_file, (size_t)(tab.n ? tab.n : 1) * sizeof(int));
      for (int i = before; i < tab.n; i++) first_file[i] = 2 * r + s;
    }
  if (!tab.n) gb_die("", "No analyzable genome (length=0)");
  /* Chromosomes are sharded over the devices (runProgram's per-chromosome loop, 5460-5607, becomes
   * the dispatcher): greedy longest-first assignment to the least loaded device.  Every context
   * knows the whole table and owns its share; what crosses devices goes through this host: the
   * per-chromosome sums behind lambda and the scale factor, the p-value histogram behind BH, and
   * the order of the peaks. */
  int nctx = o.gpus < 1 ? 1 : o.gpus;
  int* owner = (int*)calloc((size_t)tab.n, sizeof(int));
  if (!owner) gb_die("", "Cannot allocate memory");
  if (nctx > 1) {
    int usable = 0;
    for (int i = 0; i < tab.n; i++) usable += !(tab.c[i].skip || !tab.c[i].ever_saved);
    if (nctx > usable) nctx = usable > 0 ? usable : 1;        /* no device without a chromosome */
  }
  if (nctx > 1) {
    uint64_t* load = (uint64_t*)calloc((size_t)nctx, sizeof(uint64_t));
    bool* done = (bool*)calloc((size_t)tab.n, sizeof(bool));
    for (int i = 0; i < tab.n; i++) done[i] = tab.c[i].skip || !tab.c[i].ever_saved;
    for (;;) {
      int best = -1;
      for (int i = 0; i < tab.n; i++)
        if (!done[i] && (best < 0 || tab.c[i].len > tab.c[best].len)) best = i;
      if (best < 0) break;
      int k = 0;
      for (int g = 1; g < nctx; g++) if (load[g] < load[k]) k = g;
      owner[best] = k;
      load[k] += tab.c[best].len;
      done[best] = true;
    }
    free(load); free(done);
  }
  gr_params par;
  par.min_pqval = thr; par.qval_opt = o.qval_opt; par.min_auc = o.min_auc; par.min_len = o.min_len;
  par.max_gap = o.max_gap; par.keep_pileups = (o.log_file || o.pile_file) ? 1 : 0; par.genome_len = o.genome_len;
  gr_ctx** ctxs = (gr_ctx**)calloc((size_t)nctx, sizeof(gr_ctx*));
  gr_chrom* gc = (gr_chrom*)gb_alloc(tab.n * sizeof(gr_chrom));
  for (int k = 0; k < nctx; k++) {
    for (int i = 0; i < tab.n; i++) {
      gc[i].len = tab.c[i].len;
      gc[i].skip = tab.c[i].skip || !tab.c[i].ever_saved;     /* control-only references are never used (4244) */
      gc[i].owned = owner[i] == k;
      gc[i].reserved = 0;
    }
    chk(NULL, gr_create(&ctxs[k], gc, tab.n, &par, o.device + k), "gr_create");
  }
  gr_ctx* ctx = ctxs[0];
  if (xfile) load_exclusions(ctxs, nctx, xfile, &tab, o.verbose);
  uint64_t* xbp = (uint64_t*)calloc((size_t)tab.n, sizeof(uint64_t));          /* excluded bp per chromosome */
  if (xfile) chk(ctx, gr_excluded_bp(ctx, xbp), "gr_excluded_bp");

  HOut bed = { NULL, NULL }, pile = { NULL, NULL }, dupf = { NULL, NULL };
  const bool dups_verb = o.dups_opt && o.dups_file;              /* 5411-5415 */
  if (dups_verb) gb_out_open(&dupf, o.dups_file, o.gz_out);
  if (o.bed_file) gb_out_open(&bed, o.bed_file, o.gz_out);
  if (o.pile_file) gb_out_open(&pile, o.pile_file, o.gz_out);
  HIvBuf* bufs = (HIvBuf*)calloc((size_t)nctx, sizeof(HIvBuf));
  for (int k = 0; k < nctx; k++) {
    bufs[k].cap = 1u << 20;
    bufs[k].recs = (int32_t*)gr_pinned_alloc(bufs[k].cap * 16);
    bufs[k].cap_pk = 1u << 21;
    bufs[k].pk = (uint64_t*)gr_pinned_alloc(bufs[k].cap_pk * 8);
    if (!bufs[k].recs || !bufs[k].pk) gb_die("", "Cannot allocate memory");
  }
  uint8_t* save = (uint8_t*)gb_alloc(tab.n);
  double* sums = (double*)gb_alloc((size_t)tab.n * sizeof(double));
  double* esum = (double*)gb_alloc((size_t)tab.n * sizeof(double));
  double* csum = (double*)gb_alloc((size_t)tab.n * sizeof(double));
  bool* ever = (bool*)calloc((size_t)tab.n, sizeof(bool));     /* chromosomes that got a p-value array (findPeaks 1091) */

  HDecode d;
  memset(&d, 0, sizeof d);
  d.opt = &o; d.tab = &tab; d.bed = o.bed_file ? &bed : NULL;
  d.nctx = nctx; d.ctxs = ctxs; d.owner = owner; d.bufs = bufs;
  d.dups = dups_verb ? &dupf : NULL;

  for (int r = 0; r < nt; r++) {
    const char* cname = r < ncf ? cf[r] : NULL;
    const bool has_ctrl = cname && strcmp(cname, "null");
    for (int i = 0; i < tab.n; i++) tab.c[i].save = false;          /* 5463-5464 */
    gb_scan_header(real_path(tf[r]), &tab, false, &o);
    for (int i = 0; i < tab.n; i++) { save[i] = tab.c[i].save; if (save[i] && !tab.c[i].skip) ever[i] = true; }
    for (int i = 0; i < tab.n; i++) esum[i] = csum[i] = 0.0;
    for (int s = 0; s < 2; s++) {
      const char* fname = s ? cname : tf[r];
      if (s && !has_ctrl) {
        if (o.verbose) fprintf(stderr, "- control file #%d not provided -\n", r);
        break;
      }
      HIn probe;
      gb_in_open(&probe, real_path(fname));
      const bool bam = probe.is_bam;
      gb_in_close(&probe, real_path(fname));
      if (o.verbose) {
        fprintf(stderr, "Processing %s file #%d: %s\n", s ? "control" : "experimental", r, fname);
        bed_warnings_flush(first_file, tab.n, 2 * r + s);      /* saveXBed's warnings for the chromosomes this file introduces */
      }
      if (dups_verb) gb_out_printf(&dupf, "# %s file #%d: %s\n", s ? "control" : "experimental", r, fname);   /* 5493-5499 */
      for (int k = 0; k < nctx; k++) chk(ctxs[k], gr_sample_begin(ctxs[k], s, s ? NULL : save), "gr_sample_begin");
      memset(&d.cnt, 0, sizeof d.cnt);
      d.ctrl = s; d.sample = r;
      gb_decode_file(&d, real_path(fname));
      /* every device integrates its chromosomes (all enqueued before the first is waited for);
       * a chromosome has one owner, so the per-chromosome sums simply add up */
      for (int k = 0; k < nctx; k++) chk(ctxs[k], gr_sample_pileup(ctxs[k], NULL), "gr_sample_pileup");
      report_skipped(&d, ctxs, nctx, s, real_path(fname));            /* saveInterval 2558-2573 */
      if (o.verbose) log_counts(&d, bam);
      if (nctx > 1)
        for (int k = 0; k < nctx; k++) {
          chk(ctxs[k], gr_sample_sums(ctxs[k], s ? NULL : sums, s ? sums : NULL), "gr_sample_sums");
          for (int i = 0; i < tab.n; i++) (s ? csum : esum)[i] += sums[i];
        }
      if (!s) {
        /* savePileupExpt 2292-2293: an experimental sample without any weight ends the run HERE, before its
         * control file is opened (the library would report the same condition at the end of the replicate) */
        double tot = 0.0;
        if (nctx == 1) {
          chk(ctx, gr_sample_sums(ctx, sums, NULL), "gr_sample_sums");
          for (int i = 0; i < tab.n; i++) tot += sums[i];
        } else
          for (int i = 0; i < tab.n; i++) tot += esum[i];
        if (tot == 0.0) gb_die("", "Experimental sample has no analyzable fragments");
      }
    }
    gr_sample_stats st;
    if (nctx == 1)
      chk(ctx, gr_replicate_end(ctx, &st), "gr_replicate_end");
    else {
      double frag = 0.0, ctl = 0.0;                                   /* chromosome order, like the running sums */
      uint64_t glen = 0;                                              /* calcLambda 1819-1827 */
      for (int i = 0; i < tab.n; i++) {
        frag += esum[i];
        ctl += csum[i];
        if (save[i] && !tab.c[i].skip) glen += tab.c[i].len - xbp[i];
      }
      if (o.genome_len) glen = o.genome_len;
      for (int k = 0; k < nctx; k++)
        chk(ctxs[k], gr_replicate_finish(ctxs[k], frag, ctl, has_ctrl, glen, &st), "gr_replicate_finish");
    }
    if (o.verbose) {
      fprintf(stderr, "  Background pileup value: %f\n", st.lambda);                 /* 1888, 2058 */
      if (has_ctrl) {
        fprintf(stderr, "  Scaling factor for control pileup: %f\n", st.factor);     /* 2063 */
        if (st.factor > 5.0f) fprintf(stderr, "  ** Warning! Large scaling may mask true signal **\n");
      }
    }
    if (o.pile_file) write_pile(ctxs, owner, &pile, &tab, r, tf[r], cname);
  }

  const gr_peak* peaks = NULL;
  gr_peak* merged = NULL;
  uint64_t npk = 0;
  gr_run_stats rs;
  memset(&rs, 0, sizeof rs);
  if (o.peaks_opt || o.log_file) {
    if (!o.peaks_opt) {
      /* -X: p (and q) only; a threshold no value can pass keeps the peak list empty */
      gr_params p2 = par;
      p2.min_pqval = FLT_MAX;
      for (int k = 0; k < nctx; k++) gr_set_params(ctxs[k], &p2);
    }
    if (nctx == 1)
      chk(ctx, gr_call_peaks(ctx, &peaks, &npk, &rs), "gr_call_peaks");
    else {
      uint64_t G = o.genome_len;                                       /* findPeaks 1091-1101 */
      if (!G) for (int i = 0; i < tab.n; i++) if (ever[i]) G += tab.c[i].len - xbp[i];
      if (o.qval_opt) {
        /* computeQval 352 sees ONE histogram: every device gets the concatenation of all local lists */
        uint32_t* keys = NULL; uint64_t* lens = NULL; uint64_t tot = 0;
        for (int k = 0; k < nctx; k++) {
          const uint32_t* hk; const uint64_t* hl; uint64_t hn;
          chk(ctxs[k], gr_bh_local_hist_host(ctxs[k], &hk, &hl, &hn), "gr_bh_local_hist_host");
          keys = (uint32_t*)gb_realloc(keys, (tot + hn + 1) * sizeof(uint32_t));
          lens = (uint64_t*)gb_realloc(lens, (tot + hn + 1) * sizeof(uint64_t));
          memcpy(keys + tot, hk, hn * sizeof(uint32_t));
          memcpy(lens + tot, hl, hn * sizeof(uint64_t));
          tot += hn;
        }
        for (int k = 0; k < nctx; k++)
          chk(ctxs[k], gr_bh_set_global_host(ctxs[k], keys, lens, tot, G), "gr_bh_set_global_host");
        free(keys); free(lens);
      }
      const gr_peak** lists = (const gr_peak**)gb_alloc((size_t)nctx * sizeof(gr_peak*));
      uint64_t* counts = (uint64_t*)gb_alloc((size_t)nctx * sizeof(uint64_t));
      for (int k = 0; k < nctx; k++) {
        gr_run_stats rk;
        memset(&rk, 0, sizeof rk);
        chk(ctxs[k], gr_call_peaks(ctxs[k], &lists[k], &counts[k], &rk), "gr_call_peaks");
        npk += counts[k];
        rs.peak_bp += rk.peak_bp;
        if (!k) rs.all_q_one = rk.all_q_one;
      }
      rs.genome_len = G;
      rs.n_peaks = npk;
      merged = (gr_peak*)gb_alloc((npk + 1) * sizeof(gr_peak));      /* peak_N follows the chromosome order (986-987) */
      chk(ctx, gr_merge_peaks(lists, counts, nctx, merged), "gr_merge_peaks");
      peaks = merged;
      free(lists); free(counts);
    }
  }
  if (o.verbose) {                                                    /* findPeaks 1103-1117 */
    if (o.peaks_opt) {
      fprintf(stderr, "Peak-calling parameters:\n");
      fprintf(stderr, "  Genome length: %ldbp\n", (long)rs.genome_len);
      fprintf(stderr, "  Significance threshold: -log(%c) > %.3f\n", o.qval_opt ? 'q' : 'p', thr);
      fprintf(stderr, "  Min. AUC: %.3f\n", o.min_auc);
      if (o.min_len) fprintf(stderr, "  Min. peak length: %dbp\n", o.min_len);
      fprintf(stderr, "  Max. gap between sites: %dbp\n", o.max_gap);
    } else {
      fprintf(stderr, "- peak-calling skipped -\n");
      fprintf(stderr, "  Genome length: %ldbp\n", (long)rs.genome_len);
    }
    if (o.qval_opt && rs.all_q_one) fprintf(stderr, "Warning! All q-values are 1\n");
  }
  if (o.peaks_opt) {
    HOut out;
    gb_out_open(&out, o.out_file, o.gz_out);
    for (uint64_t i = 0; i < npk; i++) {                              /* printPeak 885-909 */
      const gr_peak* p = &peaks[i];
      unsigned score = (unsigned)(1000.0f * p->auc / (p->end - p->start) + 0.5f);
      if (score > 1000) score = 1000;
      gb_out_printf(&out, "%s\t%ld\t%ld\tpeak_%d\t%d\t.\t%f\t%f", tab.c[p->chrom].name, (long)p->start,
                    (long)p->end, (int)i, score, p->auc, p->pval);
      if (p->qval == GR_SKIP) gb_out_printf(&out, "\t-1\t%d\n", p->summit);
      else gb_out_printf(&out, "\t%f\t%d\n", p->qval, p->summit);
    }
    gb_out_close(&out, o.out_file);
    if (o.verbose) fprintf(stderr, "Peaks identified: %d (%ldbp)\n", (int)npk, (long)rs.peak_bp);
  }
  if (o.log_file) {
    HOut lg;
    gb_out_open(&lg, o.log_file, o.gz_out);
    write_log(ctxs, owner, &lg, &tab, nt, &o, thr);
    gb_out_close(&lg, o.log_file);
  }
  if (o.pile_file) gb_out_close(&pile, o.pile_file);
  if (o.bed_file) gb_out_close(&bed, o.bed_file);
  if (dups_verb) gb_out_close(&dupf, o.dups_file);
  for (int k = 0; k < nctx; k++) {
    gr_pinned_free(bufs[k].recs);
    gr_pinned_free(bufs[k].pk);
    gr_destroy(ctxs[k]);
  }
  free(merged);
  return EXIT_SUCCESS;
}
