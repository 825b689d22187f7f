/* gb_decode.c -- SAM / BAM decode on the host cores (readSAM 4468, loadFields
 * 4350, parseCigar 4408, calcDist 4451, getScore 4383, checkHeader 4307,
 * loadChrom 4275, saveChrom 4220; readBAM 4983, parseBAM 4826, loadBAMfields
 * 4665, calcDistBAM 4697, getBAMscore 4751).  The decoder hands each record to
 * gb_parse_align() and each completed read name to gb_process_alns().  Plain SAM files are
 * decoded by several threads (decode_sam_threads), BGZF members inflated by several threads
 * (bgzf_*); both give the byte stream / record sequence of the sequential path. */
#include "gb_host.h"
#include <fcntl.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

int gb_chrom_find(const HChromTab* t, const char* name) {
  for (int i = 0; i < t->n; i++)
    if (!strcmp(t->c[i].name, name)) return i;
  return -1;
}

static bool in_list(const char* name, const char* list) {   /* checkChrom 1212: -e names, "," or " " separated */
  if (!list) return false;
  size_t n = strlen(name);
  const char* p = list;
  while (*p) {
    while (*p == ',' || *p == ' ') p++;
    const char* q = p;
    while (*q && *q != ',' && *q != ' ') q++;
    if ((size_t)(q - p) == n && !strncmp(p, name, n)) return true;
    p = q;
  }
  return false;
}

/* saveChrom 4220-4270 (no BED exclusions) */
int gb_chrom_add(HChromTab* t, const char* name, uint32_t len, bool ctrl, const HOpts* opt) {
  int i = gb_chrom_find(t, name);
  if (i >= 0) {
    if (t->c[i].len != len) gb_die(name, ": reference sequence has different lengths in BAM/SAM files");
    if (!ctrl) { t->c[i].save = true; t->c[i].ever_saved = true; }
    return i;
  }
  t->c = (HChrom*)gb_realloc(t->c, (t->n + 1) * sizeof(HChrom));
  HChrom* c = &t->c[t->n];
  c->name = (char*)gb_alloc(strlen(name) + 1);
  strcpy(c->name, name);
  c->len = len;
  c->skip = in_list(name, opt->xchrom);
  c->save = !ctrl;
  c->ever_saved = !ctrl;
  return t->n++;
}

/* one SAM header line (checkHeader 4307-4342, loadChrom 4275-4301) */
static void sam_header_line(char* line, HChromTab* tab, bool ctrl, const HOpts* opt) {
  char* save;
  char* tag = strtok_r(line, "\t", &save);
  if (!tag) return;
  if (!strcmp(tag, "@HD")) {
    char* order = NULL;
    for (char* f = strtok_r(NULL, "\t", &save); f; f = strtok_r(NULL, "\t", &save))
      if (!strncmp(f, "SO:", 3)) order = f + 3;
    if (order) order[strcspn(order, "\n")] = '\0';
    if (opt->sort_opt && (!order || strcmp(order, "queryname")))
      gb_die("", "SAM/BAM file not sorted by queryname (samtools sort -n)");
  } else if (!strcmp(tag, "@SQ")) {
    char *name = NULL, *len = NULL;
    for (char* f = strtok_r(NULL, "\t", &save); f; f = strtok_r(NULL, "\t", &save)) {
      if (!strncmp(f, "SN:", 3)) name = f + 3;
      else if (!strncmp(f, "LN:", 3)) len = f + 3;
    }
    if (!name || !len) return;
    name[strcspn(name, "\n")] = '\0';
    len[strcspn(len, "\n")] = '\0';
    gb_chrom_add(tab, name, (uint32_t)gb_parse_int(len), ctrl, opt);
  }
}

/* ---- BGZF inflate on several host threads ---------------------------------------------------
 * A BAM file is a series of independent gzip members of at most 64 KB (BGZF); each carries its
 * compressed size in a 'BC' extra subfield and its uncompressed size in its last four bytes, so
 * the members can be found by hopping over the mapped file without inflating anything.  Batches
 * of members (64 MB of uncompressed data) are inflated by all threads at once into one of two
 * buffers, while the record parser reads the other: the single-threaded gzread of the reference's
 * readBAM (4983) becomes a read from memory.  The byte stream, and so everything decoded from it,
 * is the same. */
typedef struct {
  unsigned char* buf;
  size_t cap, len;
  size_t* coff;            /* per member: offset of the deflate data in the file */
  uint32_t *csz, *usz, *crc;
  size_t* uoff;
  int nblk, capblk;
} BgzfBatch;

typedef struct {
  const unsigned char* base;
  size_t size, pos;        /* mapped file, next member to schedule */
  int nthreads;
  size_t batch_bytes;
  BgzfBatch b[2];
  int cur, filling;        /* batch being read / batch the background filler works on */
  size_t rd;
  pthread_t bg;
  bool bg_running;
} BgzfMT;

typedef struct { BgzfMT* mt; BgzfBatch* b; int t; } BgzfJob;

static void* bgzf_inflate_part(void* arg) {
  BgzfJob* j = (BgzfJob*)arg;
  BgzfBatch* b = j->b;
  const int n = j->mt->nthreads;
  const int lo = (int)((long)b->nblk * j->t / n), hi = (int)((long)b->nblk * (j->t + 1) / n);
  for (int i = lo; i < hi; i++) {
    if (!b->usz[i]) continue;                              /* the empty end-of-file member */
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) gb_die("", "Cannot parse BAM file");
    zs.next_in = (Bytef*)(j->mt->base + b->coff[i]);
    zs.avail_in = b->csz[i];
    zs.next_out = b->buf + b->uoff[i];
    zs.avail_out = b->usz[i];
    const int rc = inflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END || zs.avail_out != 0) gb_die("", "Cannot parse BAM file");
    inflateEnd(&zs);
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), b->buf + b->uoff[i], b->usz[i]) != b->crc[i])
      gb_die("", "Cannot parse BAM file");
  }
  return NULL;
}

/* schedule the next batch of members and inflate it; false: the file is not BGZF */
static bool bgzf_fill(BgzfMT* m, BgzfBatch* b) {
  b->nblk = 0;
  b->len = 0;
  while (m->pos < m->size && b->len < m->batch_bytes) {
    const unsigned char* h = m->base + m->pos;
    if (m->size - m->pos < 28 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return false;
    const unsigned xlen = h[10] | (h[11] << 8);
    if (m->size - m->pos < 12 + (size_t)xlen + 8) return false;
    unsigned bsize = 0;
    for (unsigned x = 0; x + 4 <= xlen;) {
      const unsigned char* sf = h + 12 + x;
      const unsigned slen = sf[2] | (sf[3] << 8);
      if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = (sf[4] | (sf[5] << 8)) + 1u;
      x += 4 + slen;
    }
    if (!bsize || bsize < 12 + xlen + 8 || m->size - m->pos < bsize) return false;
    if (b->nblk == b->capblk) {
      b->capblk = b->capblk ? 2 * b->capblk : 2048;
      b->coff = (size_t*)gb_realloc(b->coff, (size_t)b->capblk * sizeof(size_t));
      b->uoff = (size_t*)gb_realloc(b->uoff, (size_t)b->capblk * sizeof(size_t));
      b->csz = (uint32_t*)gb_realloc(b->csz, (size_t)b->capblk * 4);
      b->usz = (uint32_t*)gb_realloc(b->usz, (size_t)b->capblk * 4);
      b->crc = (uint32_t*)gb_realloc(b->crc, (size_t)b->capblk * 4);
    }
    const unsigned char* tail = h + bsize - 8;
    const int i = b->nblk++;
    b->coff[i] = m->pos + 12 + xlen;
    b->csz[i] = bsize - 12 - xlen - 8;
    b->crc[i] = (uint32_t)tail[0] | ((uint32_t)tail[1] << 8) | ((uint32_t)tail[2] << 16) | ((uint32_t)tail[3] << 24);
    b->usz[i] = (uint32_t)tail[4] | ((uint32_t)tail[5] << 8) | ((uint32_t)tail[6] << 16) | ((uint32_t)tail[7] << 24);
    if (b->usz[i] > 65536) return false;
    b->uoff[i] = b->len;
    b->len += b->usz[i];
    m->pos += bsize;
  }
  if (b->len > b->cap) {
    b->cap = b->len;
    b->buf = (unsigned char*)gb_realloc(b->buf, b->cap);
  }
  const int n = m->nthreads;
  pthread_t* th = (pthread_t*)gb_alloc((size_t)n * sizeof(pthread_t));
  BgzfJob* jobs = (BgzfJob*)gb_alloc((size_t)n * sizeof(BgzfJob));
  for (int t = 0; t < n; t++) {
    jobs[t].mt = m; jobs[t].b = b; jobs[t].t = t;
    if (pthread_create(&th[t], NULL, bgzf_inflate_part, &jobs[t])) gb_die("", "Cannot start a decode thread");
  }
  for (int t = 0; t < n; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
  return true;
}

static void* bgzf_bg(void* arg) {
  BgzfMT* m = (BgzfMT*)arg;
  if (!bgzf_fill(m, &m->b[m->filling])) gb_die("", "Cannot parse BAM file");
  return NULL;
}

static BgzfMT* bgzf_open(const char* path, int nthreads) {
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return NULL;
  struct stat st;
  off_t min_bytes = 4 << 20;
  { const char* e = getenv("GB_THREAD_MIN_BYTES"); if (e) min_bytes = (off_t)atoll(e); }   /* test knob */
  if (fstat(fd, &st) || !S_ISREG(st.st_mode) || st.st_size < min_bytes) { close(fd); return NULL; }
  void* base = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (base == MAP_FAILED) return NULL;
  madvise(base, (size_t)st.st_size, MADV_SEQUENTIAL);
  BgzfMT* m = (BgzfMT*)calloc(1, sizeof *m);
  if (!m) gb_die("", "Cannot allocate memory");
  m->base = (const unsigned char*)base;
  m->size = (size_t)st.st_size;
  m->nthreads = nthreads;
  m->batch_bytes = 64u << 20;
  { const char* e = getenv("GB_BGZF_BATCH_BYTES"); if (e && atoll(e) > 0) m->batch_bytes = (size_t)atoll(e); }  /* test knob */
  if (!bgzf_fill(m, &m->b[0])) {                            /* not BGZF (plain gzip): the caller keeps gzread */
    munmap(base, m->size);
    free(m->b[0].buf); free(m->b[0].coff); free(m->b[0].uoff); free(m->b[0].csz); free(m->b[0].usz); free(m->b[0].crc);
    free(m);
    return NULL;
  }
  m->cur = 0;
  m->filling = 1;
  m->bg_running = pthread_create(&m->bg, NULL, bgzf_bg, m) == 0;
  if (!m->bg_running) gb_die("", "Cannot start a decode thread");
  return m;
}

static int bgzf_read(BgzfMT* m, void* dst, unsigned n) {
  unsigned got = 0;
  while (got < n) {
    BgzfBatch* b = &m->b[m->cur];
    if (m->rd == b->len) {                                  /* this batch is used up: take the one filled meanwhile */
      if (!m->bg_running) break;
      pthread_join(m->bg, NULL);
      m->bg_running = false;
      m->cur ^= 1;
      m->rd = 0;
      if (!m->b[m->cur].len) break;                         /* end of file */
      m->filling = m->cur ^ 1;
      m->bg_running = pthread_create(&m->bg, NULL, bgzf_bg, m) == 0;
      if (!m->bg_running) gb_die("", "Cannot start a decode thread");
      continue;
    }
    size_t k = b->len - m->rd;
    if (k > n - got) k = n - got;
    memcpy((char*)dst + got, b->buf + m->rd, k);
    m->rd += k;
    got += (unsigned)k;
  }
  return (int)got;
}

static void bgzf_close(BgzfMT* m) {
  if (m->bg_running) pthread_join(m->bg, NULL);
  for (int i = 0; i < 2; i++) {
    free(m->b[i].buf); free(m->b[i].coff); free(m->b[i].uoff); free(m->b[i].csz); free(m->b[i].usz); free(m->b[i].crc);
  }
  munmap((void*)m->base, m->size);
  free(m);
}

/* where the BAM byte stream comes from: zlib's gzread, or the threaded inflater above */
typedef struct { gzFile gz; BgzfMT* mt; } BamSrc;
static int bam_read(BamSrc* s, void* dst, unsigned n) { return s->mt ? bgzf_read(s->mt, dst, n) : gzread(s->gz, dst, n); }

static int32_t src_i32(BamSrc* g, bool must) {            /* readInt32 4633 */
  unsigned char b[4];
  int n = bam_read(g, b, 4);
  if (n != 4) {
    if (must || n > 0) gb_die("", "Cannot parse BAM file");
    return -1;
  }
  return (int32_t)(b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24));
}

/* BAM header: text (first line = @HD) + reference table (readBAM 5007-5055).
 * idx_out (malloc'd) maps BAM refID -> table index. */
static int* bam_header(BamSrc* in, HChromTab* tab, bool ctrl, const HOpts* opt, int* n_ref_out) {
  int32_t l_text = src_i32(in, true);
  if (l_text < 0) gb_die("", "Cannot parse BAM file");
  char* text = (char*)gb_alloc((size_t)l_text + 1);
  if (bam_read(in, text, (unsigned)l_text) != l_text) gb_die("", "Cannot parse BAM file");
  text[l_text] = '\0';
  char* nl = strpbrk(text, "\n");
  if (nl) *nl = '\0';
  char* save;
  char* tag = strtok_r(text, "\t", &save);
  if (!tag || strcmp(tag, "@HD")) gb_die("", "Cannot parse BAM file");
  char* order = NULL;
  for (char* f = strtok_r(NULL, "\t", &save); f; f = strtok_r(NULL, "\t", &save))
    if (!strncmp(f, "SO:", 3)) order = f + 3;
  if (opt->sort_opt && (!order || strcmp(order, "queryname")))
    gb_die("", "SAM/BAM file not sorted by queryname (samtools sort -n)");
  free(text);
  int32_t n_ref = src_i32(in, true);
  if (n_ref < 0) gb_die("", "Cannot parse BAM file");
  int* idx = (int*)gb_alloc((size_t)(n_ref ? n_ref : 1) * sizeof(int));
  char name[GB_MAX_LINE];
  for (int i = 0; i < n_ref; i++) {
    int32_t l = src_i32(in, true);
    if (l < 1 || l > GB_MAX_LINE) gb_die("", "Cannot parse BAM file");
    if (bam_read(in, name, (unsigned)l) != l || name[l - 1] != '\0') gb_die("", "Cannot parse BAM file");
    idx[i] = gb_chrom_add(tab, name, (uint32_t)src_i32(in, true), ctrl, opt);
  }
  *n_ref_out = n_ref;
  return idx;
}

/* header-only pass: the engine needs the complete chromosome table up front */
void gb_scan_header(const char* path, HChromTab* tab, bool ctrl, const HOpts* opt) {
  HIn in;
  gb_in_open(&in, path);
  if (in.is_bam) {
    int n_ref;
    BamSrc src = { in.gz, NULL };
    free(bam_header(&src, tab, ctrl, opt, &n_ref));
  } else {
    char* line = (char*)gb_alloc(GB_MAX_LINE);
    while (gb_in_gets(&in, line, GB_MAX_LINE)) {
      if (line[0] != '@') break;
      sam_header_line(line, tab, ctrl, opt);
    }
    free(line);
  }
  gb_in_close(&in, path);
}

static void new_read_name(HDecode* d, const char* qname) {
  if (d->read_name[0] == '\0' || strcmp(qname, d->read_name)) {
    if (d->read_name[0] != '\0') gb_process_alns(d, d->read_name);
    d->naln = 0;
    d->qual_r1 = d->qual_r2 = 0;                            /* 4575 */
    strncpy(d->read_name, qname, GB_MAX_ALNS);
    d->read_name[GB_MAX_ALNS] = '\0';
  }
}

/* length on the reference to the 3' end from a CIGAR string (parseCigar 4408, calcDist 4451) */
static int sam_ref_dist(const char* qname, const char* seq, const char* cigar) {
  int length = strcmp(seq, "*") ? (int)strlen(seq) : 0;
  int offset = 0;
  if (strcmp(cigar, "*")) {
    int qlen = 0, num = 0;
    bool have = false;
    for (const char* p = cigar; *p; p++) {
      if (*p >= '0' && *p <= '9') { num = num * 10 + (*p - '0'); have = true; continue; }
      if (!have) gb_die(cigar, ": cannot convert to int");
      switch (*p) {
        case 'M': case '=': case 'X': qlen += num; break;
        case 'I': case 'S': qlen += num; offset -= num; break;
        case 'D': offset += num; break;
        case 'N': case 'H': case 'P': break;
        default: { char msg[4] = "' '"; msg[1] = *p; gb_die(msg, ": unknown Op in CIGAR"); }
      }
      num = 0;
      have = false;
    }
    if (!length) length = qlen;
    else if (length != qlen) gb_die(qname, ": mismatch between sequence length and CIGAR");
  } else if (!length)
    gb_die(qname, ": no sequence information (SEQ or CIGAR)");
  return length + offset;
}

static float sam_score(char* extra) {                     /* getScore 4383 */
  if (!extra) return GB_NOSCORE;
  char* save;
  for (char* f = strtok_r(extra, "\t\n", &save); f; f = strtok_r(NULL, "\t\n", &save))
    if (f[0] == 'A' && f[1] == 'S' && f[2] == ':') {
      char* v = strchr(f + 3, ':');
      if (!v) return GB_NOSCORE;
      /* "AS:i:<int>": a short decimal integer converts exactly like strtof does; the rest goes to it */
      const char* p = v + 1;
      const bool neg = *p == '-';
      if (neg) p++;
      int iv = 0, nd = 0;
      while (*p >= '0' && *p <= '9' && nd < 8) { iv = iv * 10 + (*p - '0'); p++; nd++; }
      if (nd > 0 && nd < 8 && *p == '\0') return neg ? -(float)iv : (float)iv;
      return gb_parse_float(v + 1);
    }
  return GB_NOSCORE;
}

/* one alignment record of a SAM file (the body of readSAM's loop, 4518-4590); `line` is modified */
static void sam_record(HDecode* d, char* line) {
  const HOpts* o = d->opt;
  char* f[12];
  char* p = line;
  int nf = 0;
  while (nf < 11) {                               /* 11 mandatory fields, loadFields 4350 */
    f[nf++] = p;
    char* t = strchr(p, '\t');
    if (!t) break;
    *t = '\0';
    p = t + 1;
  }
  if (nf < 11) gb_die(f[0], ": poorly formatted SAM/BAM record");
  char* extra = NULL;
  {
    char* t = strchr(f[10], '\t');
    if (t) { *t = '\0'; extra = t + 1; }
    else f[10][strcspn(f[10], "\n")] = '\0';
  }
  const char* qname = f[0];
  const uint16_t flag = (uint16_t)gb_parse_int(f[1]);
  const char* rname = f[2];
  const uint32_t pos = (uint32_t)(gb_parse_int(f[3]) - 1);
  const int mapq = (uint8_t)gb_parse_int(f[4]);
  const uint32_t pnext = (uint32_t)(gb_parse_int(f[7]) - 1);
  (void)gb_parse_int(f[8]);
  d->cnt.count++;
  if (flag & 0x4) { d->cnt.unmapped++; return; }
  if (!strcmp(qname, "*") || !strcmp(rname, "*")) gb_die(qname, ": poorly formatted SAM/BAM record");
  if (flag & 0xE00) { d->cnt.supp++; return; }
  int chrom = d->last_chrom;
  if (chrom < 0 || chrom >= d->tab->n || strcmp(d->tab->c[chrom].name, rname)) {
    chrom = gb_chrom_find(d->tab, rname);
    if (chrom < 0) gb_die(rname, ": cannot find reference sequence name in SAM header");
    d->last_chrom = chrom;
  }
  if (mapq < o->min_mapq) { d->cnt.low_mapq++; return; }
  new_read_name(d, qname);
  const int length = sam_ref_dist(qname, f[9], f[5]);
  const float score = sam_score(extra);
  if (!gb_parse_align(d, flag, chrom, pos, length, pnext, score, f[10], (int)strlen(f[10]), 33) && o->verbose) {
    char w[GB_MAX_ALNS + 96];
    snprintf(w, sizeof w, "Warning! Read %s has more than %d alignments\n", qname, GB_MAX_ALNS);
    gb_warn(d, false, w);
  }
}

static void decode_sam(HDecode* d, HIn* in) {
  char* line = (char*)gb_alloc(GB_MAX_LINE);
  bool past_header = false;
  while (gb_in_gets(in, line, GB_MAX_LINE)) {
    if (line[0] == '@') {
      if (past_header) gb_die(line, ": misplaced SAM header line");
      sam_header_line(line, d->tab, d->ctrl, d->opt);   /* idempotent: the table was pre-scanned */
      continue;
    }
    past_header = true;
    sam_record(d, line);
  }
  free(line);
}

/* ---- several host threads on one plain SAM file --------------------------------------------
 * The file is mapped and its body cut into one piece per thread, each cut moved forward to the
 * next line that starts a new read name (the file is grouped by name), so every alignment set is
 * seen whole by one worker.  A worker is an ordinary decoder with its own alignment array,
 * counters, interval buffers and lists; full interval buffers go to the engine one at a time
 * (gb_flush_intervals).  What depends on file order is put back in file order afterwards: the
 * -x and -r lists are concatenated piece by piece and the -v warnings replayed with the
 * reference's cap (saveInterval 2524: the first MAX_ALNS of them).  The interval records reach the
 * engine in file order as well (HOrder: the pieces push one after the other) -- the pileups are
 * integer adds and would not care, but the reference's int16 saturation rule (saveInterval 2558-2573),
 * which the engine replays, is an arrival-order rule.  Not order-free in the last bits: the double sum of
 * fragment lengths behind the printed average length (and the -x extension derived from it) is
 * added per piece and then over the pieces. */
typedef struct {
  HDecode d;
  HIvBuf* bufs;            /* one per engine context */
  HWarnLog wl;
  const char* beg;
  const char* end;
} SamWorker;

static void* sam_worker(void* arg) {
  SamWorker* w = (SamWorker*)arg;
  HDecode* d = &w->d;
  char* line = (char*)gb_alloc(GB_MAX_LINE);
  for (const char* p = w->beg; p < w->end;) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(w->end - p));
    size_t n = nl ? (size_t)(nl - p) + 1 : (size_t)(w->end - p);
    if (n > GB_MAX_LINE - 1) n = GB_MAX_LINE - 1;        /* what fgets would have handed over */
    memcpy(line, p, n);
    line[n] = '\0';
    p += n;
    if (line[0] == '@') gb_die(line, ": misplaced SAM header line");
    sam_record(d, line);
  }
  if (d->read_name[0] != '\0') gb_process_alns(d, d->read_name);
  d->naln = 0;
  gb_finish_piece(d);                                    /* its records follow those of the pieces before it */
  free(line);
  return NULL;
}

static size_t qname_len(const char* p, const char* end) {
  const char* t = (const char*)memchr(p, '\t', (size_t)(end - p));
  return t ? (size_t)(t - p) : (size_t)(end - p);
}

/* would sam_record keep this line for an alignment set?  (flag 0x4 / 0xE00 and MAPQ as in loadFields' callers, 4541-4557) */
static bool line_kept(const char* p, const char* end, int min_mapq) {
  const char* f = p;
  long v[5] = { 0, 0, 0, 0, 0 };                         /* fields 2 (flag) and 5 (MAPQ) are numbers */
  for (int i = 0; i < 5; i++) {
    const char* t = (const char*)memchr(f, '\t', (size_t)(end - f));
    if (!t) return false;
    if (i == 1 || i == 4) v[i] = strtol(f, NULL, 10);
    f = t + 1;
  }
  if (v[1] & 0x4) return false;
  if (v[1] & 0xE00) return false;
  return v[4] >= min_mapq;
}

static void list_append(HReadList* dst, HReadList* src) {
  if (!src->n) { free(src->r); return; }
  if (dst->n + src->n > dst->cap) {
    dst->cap = dst->n + src->n;
    dst->r = (HRead*)gb_realloc(dst->r, dst->cap * sizeof(HRead));
  }
  memcpy(dst->r + dst->n, src->r, src->n * sizeof(HRead));
  dst->n += src->n;
  free(src->r);
}

/* returns false if the file is not worth / not fit for the threaded path */
static bool decode_sam_threads(HDecode* d, const char* path, int nthreads) {
  /* -x: the average fragment length is the reference's RUNNING double sum of fragLen / count (processPair 3174);
   * sums per piece, added up afterwards, could round differently in the last bit (and the average, rounded to an
   * integer, extends every unpaired alignment): such files are decoded in file order by one thread */
  if (d->opt->avg_ext_opt) return false;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  off_t min_bytes = 8 << 20;                             /* smaller files are not worth the threads */
  { const char* e = getenv("GB_THREAD_MIN_BYTES"); if (e) min_bytes = (off_t)atoll(e); }   /* test knob */
  if (fstat(fd, &st) || !S_ISREG(st.st_mode) || st.st_size < min_bytes) { close(fd); return false; }
  const size_t size = (size_t)st.st_size;
  const char* base = (const char*)mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (base == MAP_FAILED) return false;
  madvise((void*)base, size, MADV_SEQUENTIAL);
  const char* end = base + size;
  /* header lines (the table was pre-scanned; this repeats readSAM's checks) */
  const char* p = base;
  char* line = (char*)gb_alloc(GB_MAX_LINE);
  while (p < end && *p == '@') {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    size_t n = nl ? (size_t)(nl - p) + 1 : (size_t)(end - p);
    if (n > GB_MAX_LINE - 1) n = GB_MAX_LINE - 1;
    memcpy(line, p, n);
    line[n] = '\0';
    p += n;
    sam_header_line(line, d->tab, d->ctrl, d->opt);
  }
  free(line);
  /* cuts: between two alignment sets.  readSAM starts a new set only at a KEPT record whose name differs from the
   * previous kept record's (4559-4566: unmapped, supplementary and low-MAPQ lines are dropped before the name is
   * looked at), so a cut is moved forward until the nearest kept lines on its two sides carry different names --
   * also in a file whose lines are only loosely grouped (-S). */
  const char** cut = (const char**)gb_alloc((size_t)(nthreads + 1) * sizeof(char*));
  cut[0] = p;
  cut[nthreads] = end;
  for (int k = 1; k < nthreads; k++) {
    const char* q = p + (size_t)(end - p) / (size_t)nthreads * (size_t)k;
    if (q < cut[k - 1]) q = cut[k - 1];
    while (q > p && q[-1] != '\n') q--;                  /* the start of the line that holds q */
    while (q < end && q > p) {
      const char* prev = q;                              /* nearest kept line before the cut */
      bool have_prev = false;
      while (prev > p) {
        const char* s0 = prev - 1;
        while (s0 > p && s0[-1] != '\n') s0--;
        prev = s0;
        if (line_kept(prev, end, d->opt->min_mapq)) { have_prev = true; break; }
      }
      const char* next = q;                              /* nearest kept line at or after the cut */
      while (next < end && !line_kept(next, end, d->opt->min_mapq)) {
        const char* nl = (const char*)memchr(next, '\n', (size_t)(end - next));
        next = nl ? nl + 1 : end;
      }
      if (!have_prev || next >= end) break;
      const size_t a = qname_len(prev, end), b = qname_len(next, end);
      if (a != b || memcmp(prev, next, a)) break;        /* different sets on the two sides: cut here */
      const char* nl = (const char*)memchr(next, '\n', (size_t)(end - next));
      q = nl ? nl + 1 : end;                             /* same set: the cut moves behind that line */
    }
    cut[k] = q;
  }
  SamWorker* ws = (SamWorker*)calloc((size_t)nthreads, sizeof(SamWorker));
  pthread_t* th = (pthread_t*)gb_alloc((size_t)nthreads * sizeof(pthread_t));
  if (!ws) gb_die("", "Cannot allocate memory");
  gb_flush_intervals(d);                                 /* nothing of the caller's may arrive after the pieces' records */
  HOrder order;
  pthread_mutex_init(&order.mu, NULL);
  pthread_cond_init(&order.cv, NULL);
  order.turn = 0;
  for (int k = 0; k < nthreads; k++) {
    SamWorker* w = &ws[k];
    w->d.order = &order; w->d.piece = k;
    w->d.opt = d->opt; w->d.tab = d->tab; w->d.bed = NULL; w->d.dups = NULL;
    w->d.nctx = d->nctx; w->d.ctxs = d->ctxs; w->d.owner = d->owner;
    w->d.ctrl = d->ctrl; w->d.sample = d->sample;
    w->d.last_chrom = -1;
    w->d.wlog = &w->wl;
    w->bufs = (HIvBuf*)calloc((size_t)d->nctx, sizeof(HIvBuf));
    if (!w->bufs) gb_die("", "Cannot allocate memory");
    for (int g = 0; g < d->nctx; g++) {
      HIvBuf* b = &w->bufs[g];
      b->cap = 1u << 16; b->cap_pk = d->nctx > 1 ? 1u << 18 : 1u << 19;
      b->recs = (int32_t*)gr_pinned_alloc(b->cap * 16);
      b->pk = (uint64_t*)gr_pinned_alloc(b->cap_pk * 8);
      if (!b->recs || !b->pk) gb_die("", "Cannot allocate memory");
    }
    w->d.bufs = w->bufs;
    w->beg = cut[k]; w->end = cut[k + 1];
    if (pthread_create(&th[k], NULL, sam_worker, w)) gb_die("", "Cannot start a decode thread");
  }
  for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
  /* file order again */
  for (int k = 0; k < nthreads; k++) {
    SamWorker* w = &ws[k];
    const HCounts* c = &w->d.cnt;
    for (size_t i = 0; i < w->wl.n; i++) {                 /* replay with the cap on the counted kind */
      if (!w->wl.counted[i]) fputs(w->wl.msg[i], stderr);
      else if (d->cnt.err_count < GB_MAX_ALNS) { fputs(w->wl.msg[i], stderr); d->cnt.err_count++; }
      else d->cnt.err_count++;
      free(w->wl.msg[i]);
    }
    /* counted warnings a worker did not log (its own 129th and later) */
    {
      uint64_t logged = 0;
      for (size_t i = 0; i < w->wl.n; i++) logged += w->wl.counted[i];
      d->cnt.err_count += c->err_count - logged;
    }
    free(w->wl.msg); free(w->wl.counted);
    d->cnt.count += c->count; d->cnt.unmapped += c->unmapped; d->cnt.supp += c->supp; d->cnt.skipped += c->skipped;
    d->cnt.low_mapq += c->low_mapq; d->cnt.paired += c->paired; d->cnt.sec_pair += c->sec_pair;
    d->cnt.orphan += c->orphan; d->cnt.single += c->single; d->cnt.sec_single += c->sec_single;
    d->cnt.single_pr += c->single_pr; d->cnt.paired_pr += c->paired_pr; d->cnt.total_len += c->total_len;
    if (w->d.n_unp) {
      if (d->n_unp + w->d.n_unp > d->cap_unp) {
        d->cap_unp = d->n_unp + w->d.n_unp;
        d->unp = (HUnpaired*)gb_realloc(d->unp, d->cap_unp * sizeof(HUnpaired));
      }
      memcpy(d->unp + d->n_unp, w->d.unp, w->d.n_unp * sizeof(HUnpaired));
      d->n_unp += w->d.n_unp;
    }
    free(w->d.unp);
    list_append(&d->rd_pr, &w->d.rd_pr);
    list_append(&d->rd_dc, &w->d.rd_dc);
    list_append(&d->rd_sn, &w->d.rd_sn);
    for (int g = 0; g < d->nctx; g++) { gr_pinned_free(w->bufs[g].recs); gr_pinned_free(w->bufs[g].pk); }
    free(w->bufs);
  }
  free(ws); free(th); free(cut);
  pthread_mutex_destroy(&order.mu);
  pthread_cond_destroy(&order.cv);
  munmap((void*)base, size);
  d->read_name[0] = '\0';
  return true;
}

static float bam_score(const unsigned char* x, int len) {  /* getBAMscore 4751 */
  int i = 0;
  while (i < len - 4) {
    const char t0 = (char)x[i], t1 = (char)x[i + 1], ty = (char)x[i + 2];
    i += 3;
    if (t0 == 'A' && t1 == 'S') {
      const unsigned char* v = x + i;
      switch (ty) {
        case 'c': return (float)(int8_t)v[0];
        case 'C': return (float)(uint8_t)v[0];
        case 's': return (float)(int16_t)(v[0] | (v[1] << 8));
        case 'S': return (float)(uint16_t)(v[0] | (v[1] << 8));
        case 'i': return (float)(int32_t)(v[0] | (v[1] << 8) | (v[2] << 16) | ((uint32_t)v[3] << 24));
        case 'I': return (float)(uint32_t)(v[0] | (v[1] << 8) | (v[2] << 16) | ((uint32_t)v[3] << 24));
        default: { char msg[4] = "' '"; msg[1] = ty; gb_die(msg, ": unknown value type in BAM auxiliary field"); }
      }
    }
    switch (ty) {
      case 'A': case 'c': case 'C': i += 1; break;
      case 's': case 'S': i += 2; break;
      case 'i': case 'I': case 'f': i += 4; break;
      case 'Z': while (i < len && x[i]) i++; i++; break;
      case 'H': while (i < len && x[i]) i += 2; i++; break;
      case 'B': {
        int sz = 0;
        switch ((char)x[i]) {
          case 'c': case 'C': sz = 1; break;
          case 's': case 'S': sz = 2; break;
          case 'i': case 'I': case 'f': sz = 4; break;
          default: { char msg[4] = "' '"; msg[1] = (char)x[i]; gb_die(msg, ": unknown value type in BAM auxiliary field"); }
        }
        const int32_t cnt = (int32_t)(x[i + 1] | (x[i + 2] << 8) | (x[i + 3] << 16) | ((uint32_t)x[i + 4] << 24));
        i += 1 + 4 + sz * cnt;
        break;
      }
      default: { char msg[4] = "' '"; msg[1] = ty; gb_die(msg, ": unknown value type in BAM auxiliary field"); }
    }
    if (i > len) gb_die("", "Poorly formatted BAM auxiliary field");
  }
  return GB_NOSCORE;
}

static inline int32_t le32(const unsigned char* b) {
  return (int32_t)(b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24));
}

static void decode_bam(HDecode* d, BamSrc* in) {            /* parseBAM 4826-4977 */
  const HOpts* o = d->opt;
  int n_ref;
  int* idx = bam_header(in, d->tab, d->ctrl, o, &n_ref);
  unsigned char* blk = (unsigned char*)gb_alloc(GB_MAX_LINE * 4);
  size_t cap = GB_MAX_LINE * 4;
  for (;;) {
    const int32_t bs = src_i32(in, false);
    if (bs < 0) break;
    if (bs < 32) gb_die("", "Cannot parse BAM file");
    if ((size_t)bs > cap) { cap = (size_t)bs; blk = (unsigned char*)gb_realloc(blk, cap); }
    if (bam_read(in, blk, (unsigned)bs) != bs) gb_die("", "Cannot parse BAM file");
    const int32_t refID = le32(blk), pos = le32(blk + 4);
    const uint32_t bin_mq_nl = (uint32_t)le32(blk + 8), flag_nc = (uint32_t)le32(blk + 12);
    const int l_name = bin_mq_nl & 0xFF, mapq = (bin_mq_nl >> 8) & 0xFF;
    const int n_cigar = flag_nc & 0xFFFF;
    const uint16_t flag = (uint16_t)(flag_nc >> 16);
    const int32_t l_seq = le32(blk + 16), next_pos = le32(blk + 24);
    const char* qname = (const char*)blk + 32;
    const unsigned char* cig = blk + 32 + l_name;
    const unsigned char* extra = cig + 4 * (size_t)n_cigar + (size_t)(l_seq + 1) / 2 + (size_t)l_seq;
    if (extra > blk + bs) gb_die("", "Cannot parse BAM file");
    d->cnt.count++;
    if (flag & 0x4) { d->cnt.unmapped++; continue; }
    if (!strcmp(qname, "*") || refID < 0 || refID >= n_ref || pos < 0)
      gb_die(qname, ": poorly formatted SAM/BAM record");
    if (flag & 0xE00) { d->cnt.supp++; continue; }
    if (mapq < o->min_mapq) { d->cnt.low_mapq++; continue; }
    new_read_name(d, qname);
    int length = l_seq;                                      /* calcDistBAM 4697 */
    for (int i = 0; i < n_cigar; i++) {
      const uint32_t cg = (uint32_t)le32(cig + 4 * i);
      const int op = cg & 0xF, ol = (int)(cg >> 4);
      if (op == 1 || op == 4) length -= ol;
      else if (op == 2) length += ol;
    }
    const float score = bam_score(extra, (int)(blk + bs - extra));
    /* the raw quality bytes follow the packed sequence; they are handed over as the reference does
     * (parseBAM 4911-4915: a char* into the block, length l_seq, offset 0) */
    const char* qual = (const char*)(cig + 4 * (size_t)n_cigar + (size_t)(l_seq + 1) / 2);
    if (!gb_parse_align(d, flag, idx[refID], (uint32_t)pos, length, (uint32_t)next_pos, score, qual, l_seq, 0) && o->verbose)
      fprintf(stderr, "Warning! Read %s has more than %d alignments\n", qname, GB_MAX_ALNS);
    (void)o;
  }
  free(blk);
  free(idx);
}

void gb_decode_file(HDecode* d, const char* path) {
  HIn in;
  gb_in_open(&in, path);
  d->naln = 0;
  d->qual_r1 = d->qual_r2 = 0;
  d->last_chrom = -1;
  d->read_name[0] = '\0';
  /* plain SAM files are decoded by several threads; -b wants its lines in file order */
  const bool threaded = !in.is_gz && !in.is_bam && !d->bed && !d->lookup && d->opt->threads > 1 && strcmp(path, "-")
      && decode_sam_threads(d, path, d->opt->threads);
  if (!threaded) {
    if (in.is_bam) {
      /* BGZF members are inflated by all threads, two batches deep; plain gzip keeps gzread */
      BamSrc src = { in.gz, d->opt->threads > 1 ? bgzf_open(path, d->opt->threads) : NULL };
      if (src.mt) {
        char magic[4];
        if (bgzf_read(src.mt, magic, 4) != 4 || memcmp(magic, "BAM\1", 4)) gb_die("", "Cannot parse BAM file");
      }
      decode_bam(d, &src);
      if (src.mt) bgzf_close(src.mt);
    } else
      decode_sam(d, &in);
    if (d->read_name[0] != '\0') gb_process_alns(d, d->read_name);   /* last set, 4593 */
  }
  d->naln = 0;
  if (d->opt->dups_opt) gb_find_dups(d);                            /* 4605-4615 */
  else if (d->opt->avg_ext_opt) gb_process_avg_ext(d);
  gb_flush_intervals(d);
  gb_in_close(&in, path);
}
