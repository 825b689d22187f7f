/* See gen_synth.h.  Build: gcc -O2 -shared -fPIC (library) or with
 * -DGEN_SYNTH_MAIN for the command-line tool. */
#include "gen_synth.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAX_PLACE 16

typedef struct { uint64_t s; } rng_t;

static inline uint64_t rng_next(rng_t* r) {
  uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline double rng_unif(rng_t* r) {           /* [0,1) */
  return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0);
}
static inline rng_t rng_for(uint64_t seed, uint64_t idx) {
  rng_t r;
  r.s = seed * 0xD1342543DE82EF95ULL + idx * 0x2545F4914F6CDD1DULL + 0x1234567ULL;
  rng_next(&r);
  return r;
}

/* kept placement count for k valid placements (Genrich.c:3010, 3113, 3145) */
static inline int kept_count(int k) {
  if (k > 10) return 10;
  if (k == 7 || k == 9) return k - 1;
  return k;
}

typedef struct { int32_t chrom, start, end; } place_t;

/* Draw all placements of template t.  Returns k; *kept = records to emit. */
static int draw_template(const synth_params* p, const uint64_t* cum,
                         uint64_t G, uint64_t t, place_t* pl, int* kept) {
  rng_t r = rng_for(p->seed, t);
  int k = 1;
  if (p->multimap_frac > 0.0 && rng_unif(&r) < p->multimap_frac) {
    int span = p->multimap_max - 1;
    if (span < 1) span = 1;
    k = 2 + (int)(rng_next(&r) % (uint64_t)span);
    if (k > MAX_PLACE) k = MAX_PLACE;
  }
  uint64_t ncentre = p->peak_spacing ? G / p->peak_spacing : 0;
  for (int i = 0; i < k; i++) {
    int32_t flen = p->frag_min +
      (int32_t)(rng_next(&r) % (uint64_t)(p->frag_max - p->frag_min + 1));
    double mid;
    if (ncentre && rng_unif(&r) < p->enrich) {
      uint64_t c = rng_next(&r) % ncentre;
      double u1 = rng_unif(&r), u2 = rng_unif(&r);
      double g = sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586 * u2);
      mid = ((double)c + 0.5) * (double)p->peak_spacing + g * p->peak_sigma;
    } else
      mid = rng_unif(&r) * (double)G;
    if (mid < 0.0) mid = 0.0;
    if (mid >= (double)G) mid = (double)(G - 1);
    uint64_t gm = (uint64_t)mid;
    /* locate chromosome by binary search on cumulative lengths */
    int lo = 0, hi = p->nchrom - 1;
    while (lo < hi) {
      int m = (lo + hi) / 2;
      if (gm < cum[m + 1]) hi = m; else lo = m + 1;
    }
    int64_t clen = p->chrom_len[lo];
    int64_t start = (int64_t)(gm - cum[lo]) - flen / 2;
    if (start < 0) start = 0;
    if (start + flen > clen) start = clen - flen;
    if (start < 0) start = 0;
    int64_t end = start + flen;
    if (end > clen) end = clen;
    pl[i].chrom = lo;
    pl[i].start = (int32_t)start;
    pl[i].end = (int32_t)end;
  }
  *kept = kept_count(k);
  return k;
}

static uint64_t* cum_lengths(const synth_params* p, uint64_t* G) {
  uint64_t* cum = (uint64_t*)malloc((p->nchrom + 1) * sizeof(uint64_t));
  cum[0] = 0;
  for (int i = 0; i < p->nchrom; i++) cum[i + 1] = cum[i] + p->chrom_len[i];
  *G = cum[p->nchrom];
  return cum;
}

uint64_t synth_count_records(const synth_params* p) {
  if (p->multimap_frac <= 0.0) return p->nfrag;
  uint64_t G, n = 0;
  uint64_t* cum = cum_lengths(p, &G);
  place_t pl[MAX_PLACE];
  for (uint64_t t = 0; t < p->nfrag; t++) {
    int kept;
    draw_template(p, cum, G, t, pl, &kept);
    n += kept;
  }
  free(cum);
  return n;
}

uint64_t synth_fragments(const synth_params* p, uint64_t first, uint64_t n,
                         int32_t* out) {
  uint64_t G, w = 0;
  uint64_t* cum = cum_lengths(p, &G);
  place_t pl[MAX_PLACE];
  for (uint64_t t = first; t < first + n; t++) {
    int kept;
    draw_template(p, cum, G, t, pl, &kept);
    for (int i = 0; i < kept; i++) {
      out[4 * w + 0] = pl[i].chrom;
      out[4 * w + 1] = pl[i].start;
      out[4 * w + 2] = pl[i].end;
      out[4 * w + 3] = kept;
      w++;
    }
  }
  free(cum);
  return w;
}

int synth_write_sam(const synth_params* p, FILE* f) {
  uint64_t G;
  uint64_t* cum = cum_lengths(p, &G);
  fprintf(f, "@HD\tVN:1.6\tSO:queryname\n");
  for (int i = 0; i < p->nchrom; i++)
    fprintf(f, "@SQ\tSN:chr%d\tLN:%u\n", i + 1, p->chrom_len[i]);
  place_t pl[MAX_PLACE];
  int rl = p->read_len;
  int tag = p->multimap_frac > 0.0;
  for (uint64_t t = 0; t < p->nfrag; t++) {
    int kept;
    int k = draw_template(p, cum, G, t, pl, &kept);
    for (int i = 0; i < k; i++) {
      int sec = i ? 256 : 0;
      int32_t s = pl[i].start, e = pl[i].end;
      int32_t r2 = e - rl;          /* 0-based leftmost base of the reverse mate */
      if (r2 < 0) r2 = 0;
      /* R1 forward: pos[0] = s; R2 reverse: pos[1] = r2 + rl = e */
      fprintf(f, "f%lu\t%d\tchr%d\t%d\t42\t%dM\t=\t%d\t%d\t*\t*%s\n",
              (unsigned long)t, 99 + sec, pl[i].chrom + 1, s + 1, rl, r2 + 1,
              e - s, tag ? "\tAS:i:0" : "");
      fprintf(f, "f%lu\t%d\tchr%d\t%d\t42\t%dM\t=\t%d\t%d\t*\t*%s\n",
              (unsigned long)t, 147 + sec, pl[i].chrom + 1, r2 + 1, rl, s + 1,
              -(e - s), tag ? "\tAS:i:0" : "");
    }
  }
  free(cum);
  return ferror(f) ? -1 : 0;
}

#ifdef GEN_SYNTH_MAIN
/* gen_synth --chroms N --len L [--lens a,b,c] --frags F --seed S [--enrich x]
 *           [--spacing bp] [--sigma s] [--multimap frac --mmax k]
 *           (--sam out.sam | --bin out.i32) */
static void die(const char* m) { fprintf(stderr, "gen_synth: %s\n", m); exit(2); }
int main(int argc, char** argv) {
  synth_params p;
  memset(&p, 0, sizeof p);
  p.seed = 1; p.enrich = 0.2; p.peak_spacing = 50000; p.peak_sigma = 150.0;
  p.frag_min = 100; p.frag_max = 400; p.read_len = 50; p.multimap_max = 12;
  int nchrom = 1; uint32_t len = 1000000; const char* lens = NULL;
  const char* sam = NULL; const char* bin = NULL;
  for (int i = 1; i < argc; i++) {
    const char* a = argv[i];
    const char* v = i + 1 < argc ? argv[i + 1] : NULL;
#define OPT(n) (!strcmp(a, n) && v && ++i)
    if (OPT("--chroms")) nchrom = atoi(v);
    else if (OPT("--len")) len = (uint32_t)strtoul(v, NULL, 10);
    else if (OPT("--lens")) lens = v;
    else if (OPT("--frags")) p.nfrag = strtoull(v, NULL, 10);
    else if (OPT("--seed")) p.seed = strtoull(v, NULL, 10);
    else if (OPT("--enrich")) p.enrich = atof(v);
    else if (OPT("--spacing")) p.peak_spacing = (uint32_t)strtoul(v, NULL, 10);
    else if (OPT("--sigma")) p.peak_sigma = atof(v);
    else if (OPT("--fmin")) p.frag_min = atoi(v);
    else if (OPT("--fmax")) p.frag_max = atoi(v);
    else if (OPT("--multimap")) p.multimap_frac = atof(v);
    else if (OPT("--mmax")) p.multimap_max = atoi(v);
    else if (OPT("--sam")) sam = v;
    else if (OPT("--bin")) bin = v;
    else die("bad option");
  }
  uint32_t* cl;
  if (lens) {
    nchrom = 1;
    for (const char* c = lens; *c; c++) if (*c == ',') nchrom++;
    cl = (uint32_t*)malloc(nchrom * sizeof(uint32_t));
    const char* c = lens;
    for (int i = 0; i < nchrom; i++) {
      cl[i] = (uint32_t)strtoul(c, NULL, 10);
      c = strchr(c, ','); if (c) c++;
    }
  } else {
    cl = (uint32_t*)malloc(nchrom * sizeof(uint32_t));
    for (int i = 0; i < nchrom; i++) cl[i] = len;
  }
  p.nchrom = nchrom; p.chrom_len = cl;
  if (sam) {
    FILE* f = !strcmp(sam, "-") ? stdout : fopen(sam, "w");
    if (!f) die("cannot open sam output");
    setvbuf(f, NULL, _IOFBF, 1 << 20);
    if (synth_write_sam(&p, f)) die("write error");
    if (f != stdout) fclose(f);
  }
  if (bin) {
    uint64_t n = synth_count_records(&p);
    int32_t* buf = (int32_t*)malloc(n * 16 + 16);
    uint64_t w = synth_fragments(&p, 0, p.nfrag, buf);
    FILE* f = fopen(bin, "wb");
    if (!f) die("cannot open bin output");
    fwrite(buf, 16, w, f);
    fclose(f);
    free(buf);
  }
  return 0;
}
#endif
