/* gb_frag.c -- mate pairing, multimap weighting and interval transforms on the
 * host cores; the output is the stream of (chrom, start, end, count) records the
 * device consumes.  Behaviour follows parseAlign (Genrich.c:4141-4212),
 * processAlns (3187-3265), processPair (3122-3176), processSingle (3019-3083),
 * subsamplePair/Single (3089, 2985), saveFragment (2754), saveFragAtac (2728),
 * saveUnpair (2689), processAvgExt (2614) and the clamping half of saveInterval
 * (2522-2544); -r duplicate removal is gb_dups.c. */
#include "gb_host.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* the engine takes records from one submitter at a time (include/genrich_cuda.h): decode workers
 * hand their full buffers over one after the other */
static pthread_mutex_t push_mu = PTHREAD_MUTEX_INITIALIZER;

#define ATAC_ADJ_F 5      /* ATACADJF, Genrich.h:35 */
#define ATAC_ADJ_R (-5)   /* ATACADJR, Genrich.h:36 */

static void push_buf(HDecode* d, int k, int kind, const void* data, size_t n) {
  const int rc = kind ? gr_push_intervals(d->ctxs[k], (const int32_t*)data, n) : gr_push_packed(d->ctxs[k], (const uint64_t*)data, n);
  if (rc) gb_die("pushing intervals: ", gr_strerror(rc));
}
static void push_parked(HDecode* d) {
  for (size_t i = 0; i < d->n_parked; i++) {
    push_buf(d, d->parked[i].ctx, d->parked[i].kind, d->parked[i].data, d->parked[i].n);
    free(d->parked[i].data);
  }
  d->n_parked = 0;
}
static void park(HDecode* d, int k, int kind, const void* data, size_t n) {
  if (d->n_parked == d->cap_parked) {
    d->cap_parked = d->cap_parked ? 2 * d->cap_parked : 16;
    d->parked = (HParked*)gb_realloc(d->parked, d->cap_parked * sizeof(HParked));
  }
  const size_t bytes = n * (kind ? 16 : 8);
  HParked* q = &d->parked[d->n_parked++];
  q->ctx = k; q->kind = kind; q->n = n;
  q->data = gb_alloc(bytes);
  memcpy(q->data, data, bytes);
}

/* gb_emit_interval never lets both buffers of a context fill at once (it flushes on a change of record
 * form), so whatever is flushed here leaves in the order it was emitted */
static void flush_one(HDecode* d, int k) {
  HIvBuf* b = &d->bufs[k];
  if (!b->npk && !b->n) return;
  static int drop = -1;                        /* GB_DECODE_ONLY: measurement aid, the records are discarded */
  if (drop < 0) drop = getenv("GB_DECODE_ONLY") != NULL;
  if (drop || d->lookup) { b->npk = 0; b->n = 0; return; }
  bool mine = true;
  if (d->order) {
    pthread_mutex_lock(&d->order->mu);
    mine = d->order->turn == d->piece;
    pthread_mutex_unlock(&d->order->mu);
  }
  if (!mine) {                                 /* an earlier piece of the file is still being pushed: park */
    if (b->npk) park(d, k, 0, b->pk, b->npk);
    if (b->n) park(d, k, 1, b->recs, b->n);
    b->npk = 0; b->n = 0;
    return;
  }
  pthread_mutex_lock(&push_mu);
  push_parked(d);
  if (b->npk) { push_buf(d, k, 0, b->pk, b->npk); b->npk = 0; }
  if (b->n) { push_buf(d, k, 1, b->recs, b->n); b->n = 0; }
  pthread_mutex_unlock(&push_mu);
}

/* end of a worker's piece: wait for its turn, send what is parked and what is left, pass the turn on */
void gb_finish_piece(HDecode* d) {
  if (!d->order) { gb_flush_intervals(d); return; }
  pthread_mutex_lock(&d->order->mu);
  while (d->order->turn != d->piece) pthread_cond_wait(&d->order->cv, &d->order->mu);
  pthread_mutex_unlock(&d->order->mu);
  pthread_mutex_lock(&push_mu);
  push_parked(d);
  pthread_mutex_unlock(&push_mu);
  gb_flush_intervals(d);
  free(d->parked);
  d->parked = NULL; d->cap_parked = 0;
  pthread_mutex_lock(&d->order->mu);
  d->order->turn++;
  pthread_cond_broadcast(&d->order->cv);
  pthread_mutex_unlock(&d->order->mu);
}

void gb_flush_intervals(HDecode* d) {
  for (int k = 0; k < d->nctx; k++) flush_one(d, k);
}

/* "counted" warnings are the ones saveInterval stops printing after MAX_ALNS of them (2524, 2538) */
void gb_warn(HDecode* d, bool counted, const char* msg) {
  if (d->lookup) return;
  if (counted && d->cnt.err_count++ >= GB_MAX_ALNS) return;
  if (!d->wlog) { fputs(msg, stderr); return; }
  HWarnLog* w = d->wlog;
  if (w->n == w->cap) {
    w->cap = w->cap ? 2 * w->cap : 64;
    w->msg = (char**)gb_realloc(w->msg, w->cap * sizeof(char*));
    w->counted = (uint8_t*)gb_realloc(w->counted, w->cap);
  }
  w->msg[w->n] = (char*)gb_alloc(strlen(msg) + 1);
  strcpy(w->msg[w->n], msg);
  w->counted[w->n++] = counted;
}

/* saveInterval 2516-2591, host half: clamp, messages, BED line, enqueue */
bool gb_emit_interval(HDecode* d, int chrom, int64_t start, int64_t end, const char* qname, uint8_t count) {
  const HChrom* c = &d->tab->c[chrom];
  if (start < 0) {
    if (d->opt->verbose) {
      if (d->cnt.err_count >= GB_MAX_ALNS) d->cnt.err_count++;
      else {
        char w[2 * GB_MAX_ALNS + 96];
        snprintf(w, sizeof w, "Warning! Read %s prevented from extending below 0 on %s\n", qname, c->name);
        gb_warn(d, true, w);
      }
    }
    start = 0;
  }
  if (start >= (int64_t)c->len) {
    char msg[2 * GB_MAX_ALNS + 32];
    snprintf(msg, sizeof msg, "Read %s, ref. %s", qname, c->name);
    gb_die(msg, ": read aligned beyond reference end");
  }
  if (end > (int64_t)c->len) {
    if (d->opt->verbose) {
      if (d->cnt.err_count >= GB_MAX_ALNS) d->cnt.err_count++;
      else {
        char w[2 * GB_MAX_ALNS + 96];
        snprintf(w, sizeof w, "Warning! Read %s prevented from extending past %d on %s\n", qname, c->len, c->name);
        gb_warn(d, true, w);
      }
    }
    end = c->len;
  }
  const int k = d->nctx > 1 ? d->owner[chrom] : 0;         /* the device that holds this chromosome */
  if (d->lookup) {                                         /* naming pass: saveInterval 2558-2573's warnings */
    HLookup* lk = d->lookup;
    const uint64_t idx = lk->arrival[k]++;
    if (lk->pos[k] < lk->n[k] && (lk->list[k][lk->pos[k]] >> 1) == idx) {
      fprintf(stderr, "Warning! Read %s, alignment at (%s, %ld-%ld) skipped due to %s\n", qname, c->name, (long)start,
              (long)end, (lk->list[k][lk->pos[k]] & 1) ? "underflow" : "overflow");
      lk->pos[k]++;
      return false;                                         /* saveInterval returned 0 for it (2564, 2572) */
    }
    return true;
  }
  if (d->bed)
    gb_out_printf(d->bed, "%s\t%ld\t%ld\t%s_%d_%c_%d\n", c->name, (long)start, (long)end, qname, count,
                  d->ctrl ? 'C' : 'E', d->sample);
  HIvBuf* b = &d->bufs[k];
  if (end - start >= 0 && end - start < (int64_t)GR_PACK_MAX_LEN && (uint32_t)chrom < GR_PACK_MAX_CHROM) {
    if (b->npk == b->cap_pk || b->n) flush_one(d, k);      /* full, or records of the other form wait: keep the order */
    b->pk[b->npk++] = GR_PACK(chrom, start, end, count);
    return true;
  }
  if (b->n == b->cap || b->npk) flush_one(d, k);
  int32_t* r = b->recs + 4 * b->n++;
  r[0] = chrom; r[1] = (int32_t)start; r[2] = (int32_t)end; r[3] = count;
  return true;
}

static bool usable(const HDecode* d, const HAln* a) {
  const HChrom* c = &d->tab->c[a->chrom];
  return c->save && !c->skip;
}

/* sumQual 4127-4134: the bytes are added as (signed) chars, like the reference does */
static uint16_t sum_qual(const char* qual, int len, int offset) {
  if ((int)qual[0] == 0xFF) return 0;         /* never true where char is signed -- kept as written (4128) */
  int sum = 0;
  for (int i = 0; i < len; i++) sum += qual[i] - offset;
  return sum > UINT16_MAX ? UINT16_MAX : (uint16_t)sum;
}

/* parseAlign 4141-4212 */
bool gb_parse_align(HDecode* d, uint16_t flag, int chrom, uint32_t pos, int length, uint32_t pnext, float score,
                    const char* qual, int qual_len, int qual_offset) {
  if (flag & 0x1) {
    if ((flag & 0xC0) == 0xC0) gb_die("", "Linear template with >2 reads -- not allowed");
    if (!(flag & 0xC0)) gb_die("", "Unknown index of paired alignment");
  }
  if (d->opt->dups_opt) {                       /* 4157-4165: first record of each mate that has qualities */
    uint16_t* q = (flag & 0x40) ? &d->qual_r1 : &d->qual_r2;
    if (!*q && strcmp(qual, "*")) *q = sum_qual(qual, qual_len, qual_offset);
  }
  const HChrom* c = &d->tab->c[chrom];
  const bool ignored = c->skip || !c->save;
  const bool rev = flag & 0x10, r1 = flag & 0x40, secondary = flag & 0x100;
  const uint32_t end5 = rev ? pos + length : pos;      /* 5' end of this read */
  if ((flag & 0x3) == 0x3) {
    if (ignored) d->cnt.skipped++;
    else { d->cnt.paired++; if (secondary) d->cnt.sec_pair++; }
    for (int i = 0; i < d->naln; i++) {                  /* look for the waiting mate, 4180-4190 */
      HAln* a = &d->aln[i];
      if (a->paired && !a->full && a->chrom == chrom
          && (r1 ? (!a->first && a->pos[0] == pos) : (a->first && a->pos[1] == pos))
          && (secondary ? !a->primary : a->primary)) {
        if (r1) a->pos[0] = end5; else a->pos[1] = end5;   /* updatePairedAln 4049-4060 */
        if (score == GB_NOSCORE) a->score = GB_NOSCORE;
        else if (a->score != GB_NOSCORE) a->score += score;
        a->full = true;
        return true;
      }
    }
    if (d->naln == GB_MAX_ALNS) return false;            /* savePairedAln 4066-4096 */
    HAln* a = &d->aln[d->naln++];
    a->chrom = chrom; a->score = score; a->primary = !secondary; a->full = false; a->paired = true;
    a->strand = false;
    if (r1) { a->pos[0] = end5; a->pos[1] = pnext; a->first = true; }
    else { a->pos[0] = pnext; a->pos[1] = end5; a->first = false; }
    return true;
  }
  if (ignored) d->cnt.skipped++;
  else { d->cnt.single++; if (secondary) d->cnt.sec_single++; }
  if (!d->opt->single_opt) return true;
  if (d->naln == GB_MAX_ALNS) return false;              /* saveSingleAln 4102-4122 */
  HAln* a = &d->aln[d->naln++];
  a->chrom = chrom; a->score = score; a->primary = !secondary; a->paired = false; a->full = false;
  a->strand = !rev; a->first = r1;
  a->pos[0] = pos; a->pos[1] = pos + length;
  return true;
}

/* counts 7, 9 and > 10 are cut back to 6, 8, 10 by raising the score floor to the
 * (new count)-th best score (subsamplePair 3089-3115 / subsampleSingle 2985-3012) */
static void tighten(float* scores, int k, uint8_t* count, float* floor_) {
  for (int i = 1; i < k; i++) {             /* insertion sort, descending */
    float v = scores[i];
    int j = i - 1;
    while (j >= 0 && scores[j] < v) { scores[j + 1] = scores[j]; j--; }
    scores[j + 1] = v;
  }
  *count = *count > 10 ? 10 : (uint8_t)(*count - 1);
  *floor_ = scores[*count - 1];
}

static void emit_fragment(HDecode* d, const char* qname, const HAln* a, uint8_t count, uint64_t* frag_len) {
  const HOpts* o = d->opt;
  uint32_t s = a->pos[0], e = a->pos[1];
  if (s > e) { uint32_t t = s; s = e; e = t; }          /* saveFragment 2759-2766 */
  if (!o->atac_opt) {
    /* false only in the naming pass of report_skipped, for a record the engine dropped: the reference's
     * saveInterval returned 0 for it, so it is not in its average fragment length either */
    if (gb_emit_interval(d, a->chrom, s, e, qname, count))
      *frag_len += (e > (uint32_t)d->tab->c[a->chrom].len ? d->tab->c[a->chrom].len : e) - s;
    return;
  }
  if (o->atac_adj) { s += ATAC_ADJ_F; e += ATAC_ADJ_R; }   /* saveFragAtac 2733-2748, uint32 arithmetic */
  if (s + o->atac_len3 >= (uint32_t)(int)(e - o->atac_len3))
    gb_emit_interval(d, a->chrom, (int)(s - o->atac_len5), (int64_t)e + o->atac_len5, qname, count);
  else {
    gb_emit_interval(d, a->chrom, (int)(s - o->atac_len5), (int64_t)s + o->atac_len3, qname, count);
    gb_emit_interval(d, a->chrom, (int)(e - o->atac_len3), (int64_t)e + o->atac_len5, qname, count);
  }
}

static void emit_unpaired(HDecode* d, const char* qname, HAln* a, uint8_t count, bool extend_opt, int extend) {   /* saveUnpair 2689-2721 */
  const HOpts* o = d->opt;
  if (extend_opt) {
    if (a->strand) gb_emit_interval(d, a->chrom, a->pos[0], (int64_t)a->pos[0] + extend, qname, count);
    else gb_emit_interval(d, a->chrom, (int)(a->pos[1] - extend), a->pos[1], qname, count);
  } else if (o->atac_opt) {
    if (a->strand) {
      if (o->atac_adj) a->pos[0] += ATAC_ADJ_F;
      gb_emit_interval(d, a->chrom, (int)(a->pos[0] - o->atac_len5), (int64_t)a->pos[0] + o->atac_len3, qname, count);
    } else {
      if (o->atac_adj) a->pos[1] += ATAC_ADJ_R;
      gb_emit_interval(d, a->chrom, (int)(a->pos[1] - o->atac_len3), (int64_t)a->pos[1] + o->atac_len5, qname, count);
    }
  } else
    gb_emit_interval(d, a->chrom, a->pos[0], a->pos[1], qname, count);
}

static void defer_unpaired(HDecode* d, const char* qname, const HAln* a, uint8_t count) {   /* saveAvgExt 2654 */
  if (d->n_unp == d->cap_unp) {
    d->cap_unp = d->cap_unp ? 2 * d->cap_unp : 65536;
    d->unp = (HUnpaired*)gb_realloc(d->unp, d->cap_unp * sizeof(HUnpaired));
  }
  HUnpaired* u = &d->unp[d->n_unp++];
  u->chrom = a->chrom; u->pos[0] = a->pos[0]; u->pos[1] = a->pos[1]; u->strand = a->strand; u->count = count;
  u->name = (char*)gb_alloc(strlen(qname) + 1);
  strcpy(u->name, qname);
}

/* processPair 3122-3176 */
int gb_do_pairs(HDecode* d, const char* qname, const HAln* aln, int naln, float best) {
  float floor_ = best;
  if (floor_ != GB_NOSCORE) floor_ -= d->opt->as_diff;
  float sc[GB_MAX_ALNS];
  int k = 0;
  for (int i = 0; i < naln; i++) {
    const HAln* a = &aln[i];
    if (a->paired && a->full && a->score >= floor_ && usable(d, a)) sc[k++] = a->score;
  }
  if (!k) return 0;
  uint8_t count = (uint8_t)k;
  if (k > 10 || k == 7 || k == 9) tighten(sc, k, &count, &floor_);
  uint64_t frag_len = 0;
  uint8_t saved = 0;
  for (int i = 0; i < naln && saved < count; i++) {
    const HAln* a = &aln[i];
    if (a->paired && a->full && a->score >= floor_ && usable(d, a)) {
      emit_fragment(d, qname, a, count, &frag_len);
      saved++;
    }
  }
  d->cnt.total_len += (double)frag_len / count;
  return 1;
}

/* processSingle 3019-3083.  extend_opt / extend: -w, or the average fragment length found after the
 * paired sets of a -r run (findDups 4015-4019); defer: -x without -r (saveAvgExt 2654) */
int gb_do_singles(HDecode* d, const char* qname, HAln* aln, int naln, float best, bool first,
                  bool extend_opt, int extend, bool defer) {
  float floor_ = best;
  if (floor_ != GB_NOSCORE) floor_ -= d->opt->as_diff;
  float sc[GB_MAX_ALNS];
  int k = 0;
  for (int i = 0; i < naln; i++) {
    const HAln* a = &aln[i];
    if (!a->paired && a->first == first && a->score >= floor_ && usable(d, a)) sc[k++] = a->score;
  }
  if (!k) return 0;
  uint8_t count = (uint8_t)k;
  if (k > 10 || k == 7 || k == 9) tighten(sc, k, &count, &floor_);
  uint8_t saved = 0;
  for (int i = 0; i < naln && saved < count; i++) {
    HAln* a = &aln[i];
    if (!a->paired && a->first == first && a->score >= floor_ && usable(d, a)) {
      if (defer) defer_unpaired(d, qname, a, count);
      else emit_unpaired(d, qname, a, count, extend_opt, extend);
      saved++;
    }
  }
  return 1;
}

/* processAlns 3187-3265 */
void gb_process_alns(HDecode* d, const char* qname) {
  float best_pr = GB_NOSCORE, best_r1 = GB_NOSCORE, best_r2 = GB_NOSCORE;
  bool pair = false, s1 = false, s2 = false;
  for (int i = 0; i < d->naln; i++) {
    const HAln* a = &d->aln[i];
    if (a->paired) {
      if (a->full) {
        if (!pair || best_pr < a->score) best_pr = a->score;
        pair = true;
      } else
        d->cnt.orphan++;
    } else if (d->opt->single_opt && !pair) {
      if (a->first && best_r1 <= a->score) { best_r1 = a->score; s1 = true; }
      else if (!a->first && best_r2 <= a->score) { best_r2 = a->score; s2 = true; }
    }
  }
  if (d->opt->dups_opt) {                       /* 3228-3235: kept for the end of the file */
    gb_save_alns(d, qname, pair, s1, s2, best_pr, best_r1, best_r2);
    return;
  }
  const HOpts* o = d->opt;
  if (pair)
    d->cnt.paired_pr += gb_do_pairs(d, qname, d->aln, d->naln, best_pr);
  else if (o->single_opt) {
    if (s1) d->cnt.single_pr += gb_do_singles(d, qname, d->aln, d->naln, best_r1, true, o->extend_opt, o->extend, o->avg_ext_opt);
    if (s2) d->cnt.single_pr += gb_do_singles(d, qname, d->aln, d->naln, best_r2, false, o->extend_opt, o->extend, o->avg_ext_opt);
  }
}

/* processAvgExt 2614-2647 (calcAvgLen 2597) */
void gb_process_avg_ext(HDecode* d) {
  int avg = 0;
  if (!d->cnt.paired_pr) {
    if (d->opt->verbose) {
      fprintf(stderr, "Warning! No paired alignments to calculate avg frag ");
      fprintf(stderr, "length --\n  Printing unpaired alignments \"as is\"\n");
    }
  } else
    avg = (int)(d->cnt.total_len / d->cnt.paired_pr + 0.5);
  for (size_t i = 0; i < d->n_unp; i++) {
    HUnpaired* u = &d->unp[i];
    if (!avg) gb_emit_interval(d, u->chrom, u->pos[0], u->pos[1], u->name, u->count);
    else if (u->strand) gb_emit_interval(d, u->chrom, u->pos[0], (int64_t)u->pos[0] + avg, u->name, u->count);
    else gb_emit_interval(d, u->chrom, (int)(u->pos[1] - avg), u->pos[1], u->name, u->count);
    free(u->name);
  }
  d->n_unp = 0;
}
