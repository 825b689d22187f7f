/* gb_dups.c -- -r: PCR duplicate removal on the host cores.
 *
 * Behaviour follows the reference: with -r the alignment sets of a file are not turned into
 * intervals as they are read but kept (saveAlns, Genrich.c:2940-2978: properly paired sets,
 * discordant sets = both mates aligned but not as a pair, singleton sets), and at the end of
 * the file (findDups 3949-4043) each class is visited in order of decreasing quality-score sum
 * (sortReads 3362: a STABLE sort, ties keep file order).  A set is a duplicate if ANY of its
 * alignments has been seen before -- pairs by (reference, both 5' ends) (checkHashPr 3592),
 * discordant sets by (both references, both 5' ends, both strands) in either order (checkHashDc
 * 3720), singletons by (reference, 5' end, strand) against everything kept so far including
 * the ends of kept pairs and discordant sets (addHashPr 3582-3587, addHashDc 3704-3711,
 * checkHashSn 3861); sets that are not duplicates are entered and processed as usual
 * (processPair 3122 / processSingle 3019).  The reference's chained hash tables are an
 * implementation detail (equality decides, not the hash): here one open-addressing table per
 * class keyed by the same tuples.  -R <file>: the log of logDup 3528-3564.
 */
#include "gb_host.h"
#include <stdlib.h>
#include <string.h>

/* ---- keeping the sets (saveAlns 2940) -------------------------------------------------- */
static HRead* read_new(HReadList* l) {
  if (l->n == l->cap) {
    l->cap = l->cap ? 2 * l->cap : 65536;
    l->r = (HRead*)gb_realloc(l->r, l->cap * sizeof(HRead));
  }
  HRead* r = &l->r[l->n++];
  memset(r, 0, sizeof *r);
  return r;
}

static char* dup_name(const char* s) {
  char* n = (char*)gb_alloc(strlen(s) + 1);
  strcpy(n, s);
  return n;
}

/* copyAlns 2814-2850 */
static void copy_singles(const HDecode* d, float score, bool first, HAln** dest, uint8_t* ndest) {
  if (score != GB_NOSCORE) score -= d->opt->as_diff;
  uint8_t count = 0;
  for (int i = 0; i < d->naln; i++) {
    const HAln* a = &d->aln[i];
    if (!a->paired && a->first == first && a->score >= score) count++;
  }
  *dest = (HAln*)gb_alloc((size_t)(count ? count : 1) * sizeof(HAln));
  *ndest = count;
  uint8_t j = 0;
  for (int i = 0; i < d->naln; i++) {
    const HAln* a = &d->aln[i];
    if (!a->paired && a->first == first && a->score >= score) (*dest)[j++] = *a;
  }
}

static uint16_t qual_sum2(uint16_t a, uint16_t b) {          /* MIN(qualR1 + qualR2, UINT16_MAX) */
  const int s = (int)a + (int)b;
  return s > UINT16_MAX ? UINT16_MAX : (uint16_t)s;
}

void gb_save_alns(HDecode* d, const char* qname, bool pair, bool s1, bool s2, float best_pr, float best_r1,
                  float best_r2) {
  if (pair) {                                                 /* saveAlnsPair 2890-2934 */
    HRead* r = read_new(&d->rd_pr);
    r->name = dup_name(qname);
    r->qual = qual_sum2(d->qual_r1, d->qual_r2);
    r->score = best_pr;
    float floor_ = best_pr;
    if (floor_ != GB_NOSCORE) floor_ -= d->opt->as_diff;
    uint8_t count = 0;
    for (int i = 0; i < d->naln; i++) {
      const HAln* a = &d->aln[i];
      if (a->paired && a->full && a->score >= floor_) count++;
    }
    r->aln = (HAln*)gb_alloc((size_t)(count ? count : 1) * sizeof(HAln));
    r->naln = count;
    uint8_t j = 0;
    for (int i = 0; i < d->naln; i++) {
      const HAln* a = &d->aln[i];
      if (a->paired && a->full && a->score >= floor_) {
        HAln* b = &r->aln[j++];
        *b = *a;
        if (a->pos[0] > a->pos[1]) { b->pos[0] = a->pos[1]; b->pos[1] = a->pos[0]; }   /* positions ordered */
      }
    }
  } else if (d->opt->single_opt) {
    if (s1 && s2) {                                           /* saveAlnsDiscord 2873-2885 */
      HRead* r = read_new(&d->rd_dc);
      r->name = dup_name(qname);
      r->first = true;
      r->score = best_r1;
      r->score_r2 = best_r2;
      copy_singles(d, best_r1, true, &r->aln, &r->naln);
      copy_singles(d, best_r2, false, &r->aln_r2, &r->naln_r2);
      r->qual = qual_sum2(d->qual_r1, d->qual_r2);
    } else if (s1 || s2) {                                    /* saveAlnsSingle 2855-2868 */
      HRead* r = read_new(&d->rd_sn);
      r->name = dup_name(qname);
      r->first = s1;
      r->score = s1 ? best_r1 : best_r2;
      r->qual = s1 ? d->qual_r1 : d->qual_r2;
      copy_singles(d, r->score, s1, &r->aln, &r->naln);
    }
  }
}

/* ---- order of evaluation (sortReads 3362 / johnSort 3331: stable, descending quality sum) ---- */
static uint32_t* qual_order(const HReadList* l) {
  uint32_t* order = (uint32_t*)gb_alloc((l->n ? l->n : 1) * sizeof(uint32_t));
  size_t* start = (size_t*)calloc(65537, sizeof(size_t));
  if (!start) gb_die("", "Cannot allocate memory");
  for (size_t i = 0; i < l->n; i++) start[65535 - l->r[i].qual + 1]++;      /* bucket 0 = highest sum */
  for (int k = 0; k < 65536; k++) start[k + 1] += start[k];
  for (size_t i = 0; i < l->n; i++) order[start[65535 - l->r[i].qual]++] = (uint32_t)i;
  free(start);
  return order;
}

/* ---- the "seen" tables ---------------------------------------------------------------------- */
typedef struct {
  int chrom, chrom1;
  uint32_t pos, pos1;
  uint8_t strand, strand1, used;
  const char* name;      /* read that entered the key (-R log) */
} DKey;

typedef struct {
  DKey* e;
  size_t cap, n;         /* cap: power of two */
} DTable;

static void tab_init(DTable* t, size_t expect) {
  size_t cap = 1024;
  while (cap < 2 * expect + 16) cap <<= 1;
  t->e = (DKey*)calloc(cap, sizeof(DKey));
  if (!t->e) gb_die("", "Cannot allocate memory");
  t->cap = cap;
  t->n = 0;
}

static size_t key_hash(const DKey* k) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
  const uint64_t v[3] = { ((uint64_t)(uint32_t)k->chrom << 32) | (uint32_t)k->chrom1,
                          ((uint64_t)k->pos << 32) | k->pos1, ((uint64_t)k->strand << 1) | k->strand1 };
  for (int i = 0; i < 3; i++) {
    h ^= v[i];
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 32;
  }
  return (size_t)h;
}

static bool key_eq(const DKey* a, const DKey* b) {
  return a->chrom == b->chrom && a->chrom1 == b->chrom1 && a->pos == b->pos && a->pos1 == b->pos1
      && a->strand == b->strand && a->strand1 == b->strand1;
}

static const DKey* tab_find(const DTable* t, const DKey* k) {
  for (size_t i = key_hash(k) & (t->cap - 1);; i = (i + 1) & (t->cap - 1)) {
    if (!t->e[i].used) return NULL;
    if (key_eq(&t->e[i], k)) return &t->e[i];
  }
}

static void tab_grow(DTable* t);
static void tab_add(DTable* t, const DKey* k) {              /* the caller knows the key is absent, or does not care */
  if (2 * (t->n + 1) > t->cap) tab_grow(t);
  size_t i = key_hash(k) & (t->cap - 1);
  while (t->e[i].used) {
    if (key_eq(&t->e[i], k)) return;                         /* first entry wins, like the head-of-chain search */
    i = (i + 1) & (t->cap - 1);
  }
  t->e[i] = *k;
  t->e[i].used = 1;
  t->n++;
}
static void tab_grow(DTable* t) {
  DTable big;
  tab_init(&big, t->cap);
  for (size_t i = 0; i < t->cap; i++)
    if (t->e[i].used) tab_add(&big, &t->e[i]);
  free(t->e);
  *t = big;
}

static DKey key_pair(const HAln* a, const char* name) {
  DKey k = { a->chrom, -1, a->pos[0], a->pos[1], 0, 0, 0, name };
  return k;
}
static DKey key_single(int chrom, uint32_t pos, bool strand, const char* name) {
  DKey k = { chrom, -1, pos, 0, strand, 0, 0, name };
  return k;
}
static uint32_t end5(const HAln* a) { return a->strand ? a->pos[0] : a->pos[1]; }

/* checkAndAdd 3515-3523 */
static void single_check_add(DTable* sn, int chrom, uint32_t pos, bool strand, const char* name) {
  const DKey k = key_single(chrom, pos, strand, name);
  if (!tab_find(sn, &k)) tab_add(sn, &k);
}

/* logDup 3528-3564 */
static void log_pair(HDecode* d, const char* name, const HAln* a, const char* match) {
  gb_out_printf(d->dups, "%s\t%s:%d-%d\t%s\tpaired\n", name, d->tab->c[a->chrom].name, a->pos[0], a->pos[1], match);
}
static void log_single(HDecode* d, const char* name, int chrom, uint32_t pos, bool strand, const char* match) {
  gb_out_printf(d->dups, "%s\t%s:%d,%c\t%s\tsingle\n", name, d->tab->c[chrom].name, pos, strand ? '+' : '-', match);
}
static void log_discord(HDecode* d, const char* name, const DKey* k, const char* match) {
  gb_out_printf(d->dups, "%s\t%s:%d,%c;%s:%d,%c\t%s\tdiscordant\n", name, d->tab->c[k->chrom].name, k->pos,
                k->strand ? '+' : '-', d->tab->c[k->chrom1].name, k->pos1, k->strand1 ? '+' : '-', match);
}

static void read_free(HRead* r) {
  free(r->name);
  free(r->aln);
  free(r->aln_r2);
}

/* findDups 3949-4043 */
void gb_find_dups(HDecode* d) {
  const HOpts* o = d->opt;
  HCounts* c = &d->cnt;
  const bool log = d->dups != NULL;
  /* the singleton table exists only if the file has singleton sets (3968-3985) */
  const bool use_sn = o->single_opt && d->rd_sn.n > 0;
  DTable sn = { NULL, 0, 0 };
  if (use_sn) tab_init(&sn, 2 * d->rd_pr.n + 2 * d->rd_dc.n + d->rd_sn.n);

  if (d->rd_pr.n) {                                           /* findDupsPr 3616-3683 */
    DTable t;
    tab_init(&t, d->rd_pr.n);
    uint32_t* order = qual_order(&d->rd_pr);
    for (size_t i = 0; i < d->rd_pr.n; i++) {
      HRead* r = &d->rd_pr.r[order[i]];
      bool dup = false;
      for (int k = 0; k < r->naln && !dup; k++) {
        const DKey key = key_pair(&r->aln[k], r->name);
        const DKey* h = tab_find(&t, &key);
        if (h) {
          if (log) log_pair(d, r->name, &r->aln[k], h->name);
          dup = true;
        }
      }
      if (dup) c->dups_pr++;
      else {
        for (int k = 0; k < r->naln; k++) {
          const HAln* a = &r->aln[k];
          const DKey key = key_pair(a, r->name);
          tab_add(&t, &key);
          if (use_sn) {                                       /* both ends as singletons, 3582-3587 */
            single_check_add(&sn, a->chrom, a->pos[0], true, r->name);
            single_check_add(&sn, a->chrom, a->pos[1], false, r->name);
          }
        }
        c->paired_pr += gb_do_pairs(d, r->name, r->aln, r->naln, r->score);
      }
      c->count_pr++;
    }
    free(order);
    free(t.e);
  }

  if (o->single_opt) {
    bool extend_opt = o->extend_opt;
    int extend = o->extend;
    if (o->avg_ext_opt) {                                     /* 4013-4019, calcAvgLen 2597 */
      extend = 0;
      if (!c->paired_pr) {
        if (o->verbose) {
          fprintf(stderr, "Warning! No paired alignments to calculate avg frag ");
          fprintf(stderr, "length --\n  Printing unpaired alignments \"as is\"\n");
        }
      } else
        extend = (int)(c->total_len / c->paired_pr + 0.5);
      if (extend) extend_opt = true;
    }

    if (d->rd_dc.n) {                                         /* findDupsDc 3761-3840 */
      DTable t;
      tab_init(&t, d->rd_dc.n);
      uint32_t* order = qual_order(&d->rd_dc);
      for (size_t i = 0; i < d->rd_dc.n; i++) {
        HRead* r = &d->rd_dc.r[order[i]];
        bool dup = false;
        for (int k = 0; k < r->naln && !dup; k++) {           /* checkHashDc 3720-3756: either order */
          const HAln* a = &r->aln[k];
          for (int j = 0; j < r->naln_r2 && !dup; j++) {
            const HAln* b = &r->aln_r2[j];
            const DKey fw = { a->chrom, b->chrom, end5(a), end5(b), a->strand, b->strand, 0, r->name };
            const DKey rv = { b->chrom, a->chrom, end5(b), end5(a), b->strand, a->strand, 0, r->name };
            const DKey* h = tab_find(&t, &fw);
            if (h) { if (log) log_discord(d, r->name, &fw, h->name); dup = true; break; }
            h = tab_find(&t, &rv);
            if (h) { if (log) log_discord(d, r->name, &rv, h->name); dup = true; }
          }
        }
        if (dup) c->dups_dc++;
        else {
          for (int k = 0; k < r->naln; k++) {                 /* addHashDc 3689-3714 */
            const HAln* a = &r->aln[k];
            for (int j = 0; j < r->naln_r2; j++) {
              const HAln* b = &r->aln_r2[j];
              const DKey fw = { a->chrom, b->chrom, end5(a), end5(b), a->strand, b->strand, 0, r->name };
              tab_add(&t, &fw);
              if (use_sn) {
                if (!j) single_check_add(&sn, a->chrom, end5(a), a->strand, r->name);
                if (!k) single_check_add(&sn, b->chrom, end5(b), b->strand, r->name);
              }
            }
          }
          c->single_pr += gb_do_singles(d, r->name, r->aln, r->naln, r->score, true, extend_opt, extend, false);
          c->single_pr += gb_do_singles(d, r->name, r->aln_r2, r->naln_r2, r->score_r2, false, extend_opt, extend, false);
        }
        c->count_dc++;
      }
      free(order);
      free(t.e);
    }

    if (d->rd_sn.n) {                                         /* findDupsSn 3886-3943 */
      uint32_t* order = qual_order(&d->rd_sn);
      for (size_t i = 0; i < d->rd_sn.n; i++) {
        HRead* r = &d->rd_sn.r[order[i]];
        bool dup = false;
        for (int k = 0; k < r->naln && !dup; k++) {
          const HAln* a = &r->aln[k];
          const DKey key = key_single(a->chrom, end5(a), a->strand, r->name);
          const DKey* h = tab_find(&sn, &key);
          if (h) {
            if (log) log_single(d, r->name, a->chrom, end5(a), a->strand, h->name);
            dup = true;
          }
        }
        if (dup) c->dups_sn++;
        else {
          for (int k = 0; k < r->naln; k++) {                 /* addHashSn 3845-3855 */
            const HAln* a = &r->aln[k];
            const DKey key = key_single(a->chrom, end5(a), a->strand, r->name);
            tab_add(&sn, &key);
          }
          c->single_pr += gb_do_singles(d, r->name, r->aln, r->naln, r->score, r->first, extend_opt, extend, false);
        }
        c->count_sn++;
      }
      free(order);
    }
  }

  /* the names logged as matches live in the read lists: release everything last */
  HReadList* lists[3] = { &d->rd_pr, &d->rd_dc, &d->rd_sn };
  for (int k = 0; k < 3; k++) {
    for (size_t i = 0; i < lists[k]->n; i++) read_free(&lists[k]->r[i]);
    lists[k]->n = 0;
  }
  free(sn.e);
}
