/* gb_decode.c -- SAM / BAM decode on the host cores (readSAM 4468, loadFields
 * 4350, parseCigar 4408, calcDist 4451, getScore 4383, checkHeader 4307,
 * loadChrom 4275, saveChrom 4220; readBAM 4983, parseBAM 4826, loadBAMfields
 * 4665, calcDistBAM 4697, getBAMscore 4751).  The decoder hands each record to
 * gb_parse_align() and each completed read name to gb_process_alns(). */
#include "gb_host.h"
#include <stdlib.h>
#include <string.h>

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

static int32_t gz_i32(gzFile g, bool must) {              /* readInt32 4633 */
  unsigned char b[4];
  int n = gzread(g, b, 4);
  if (n != 4) {
    if (must || n > 0) gb_die("", "Cannot parse BAM file");
    return -1;
  }
  return (int32_t)(b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24));
}

/* BAM header: text (first line = @HD) + reference table (readBAM 5007-5055).
 * idx_out (malloc'd) maps BAM refID -> table index. */
static int* bam_header(HIn* in, HChromTab* tab, bool ctrl, const HOpts* opt, int* n_ref_out) {
  int32_t l_text = gz_i32(in->gz, true);
  if (l_text < 0) gb_die("", "Cannot parse BAM file");
  char* text = (char*)gb_alloc((size_t)l_text + 1);
  if (gzread(in->gz, text, (unsigned)l_text) != l_text) gb_die("", "Cannot parse BAM file");
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
  int32_t n_ref = gz_i32(in->gz, true);
  if (n_ref < 0) gb_die("", "Cannot parse BAM file");
  int* idx = (int*)gb_alloc((size_t)(n_ref ? n_ref : 1) * sizeof(int));
  char name[GB_MAX_LINE];
  for (int i = 0; i < n_ref; i++) {
    int32_t l = gz_i32(in->gz, true);
    if (l < 1 || l > GB_MAX_LINE) gb_die("", "Cannot parse BAM file");
    if (gzread(in->gz, name, (unsigned)l) != l || name[l - 1] != '\0') gb_die("", "Cannot parse BAM file");
    idx[i] = gb_chrom_add(tab, name, (uint32_t)gz_i32(in->gz, true), ctrl, opt);
  }
  *n_ref_out = n_ref;
  return idx;
}

/* header-only pass: the engine needs the complete chromosome table up front */
void gb_scan_header(const char* path, HChromTab* tab, bool ctrl, const HOpts* opt) {
  if (!strcmp(path, "-")) gb_die(path, ": reading alignments from stdin is not supported (the chromosome table is scanned first)");
  HIn in;
  gb_in_open(&in, path);
  if (in.is_bam) {
    int n_ref;
    free(bam_header(&in, tab, ctrl, opt, &n_ref));
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
      return gb_parse_float(v + 1);
    }
  return GB_NOSCORE;
}

static void decode_sam(HDecode* d, HIn* in) {
  const HOpts* o = d->opt;
  char* line = (char*)gb_alloc(GB_MAX_LINE);
  bool past_header = false;
  while (gb_in_gets(in, line, GB_MAX_LINE)) {
    if (line[0] == '@') {
      if (past_header) gb_die(line, ": misplaced SAM header line");
      sam_header_line(line, d->tab, d->ctrl, o);   /* idempotent: the table was pre-scanned */
      continue;
    }
    past_header = true;
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
    if (flag & 0x4) { d->cnt.unmapped++; continue; }
    if (!strcmp(qname, "*") || !strcmp(rname, "*")) gb_die(qname, ": poorly formatted SAM/BAM record");
    if (flag & 0xE00) { d->cnt.supp++; continue; }
    const int chrom = gb_chrom_find(d->tab, rname);
    if (chrom < 0) gb_die(rname, ": cannot find reference sequence name in SAM header");
    if (mapq < o->min_mapq) { d->cnt.low_mapq++; continue; }
    new_read_name(d, qname);
    const int length = sam_ref_dist(qname, f[9], f[5]);
    const float score = sam_score(extra);
    if (!gb_parse_align(d, flag, chrom, pos, length, pnext, score, f[10], (int)strlen(f[10]), 33) && o->verbose)
      fprintf(stderr, "Warning! Read %s has more than %d alignments\n", qname, GB_MAX_ALNS);
  }
  free(line);
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

static void decode_bam(HDecode* d, HIn* in) {               /* parseBAM 4826-4977 */
  const HOpts* o = d->opt;
  int n_ref;
  int* idx = bam_header(in, d->tab, d->ctrl, o, &n_ref);
  unsigned char* blk = (unsigned char*)gb_alloc(GB_MAX_LINE * 4);
  size_t cap = GB_MAX_LINE * 4;
  for (;;) {
    const int32_t bs = gz_i32(in->gz, false);
    if (bs < 0) break;
    if (bs < 32) gb_die("", "Cannot parse BAM file");
    if ((size_t)bs > cap) { cap = (size_t)bs; blk = (unsigned char*)gb_realloc(blk, cap); }
    if (gzread(in->gz, blk, (unsigned)bs) != bs) gb_die("", "Cannot parse BAM file");
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
  }
  free(blk);
  free(idx);
}

void gb_decode_file(HDecode* d, const char* path) {
  HIn in;
  gb_in_open(&in, path);
  d->naln = 0;
  d->qual_r1 = d->qual_r2 = 0;
  if (in.is_bam) decode_bam(d, &in);
  else decode_sam(d, &in);
  if (d->read_name[0] != '\0') gb_process_alns(d, d->read_name);   /* last set, 4593 */
  d->naln = 0;
  if (d->opt->dups_opt) gb_find_dups(d);                            /* 4605-4615 */
  else if (d->opt->avg_ext_opt) gb_process_avg_ext(d);
  gb_flush_intervals(d);
  gb_in_close(&in, path);
}
