/* gb_host.h -- host side (plain C) of the genrich-b200 command-line program.
 *
 * Everything the reference does on the host stays on the host: option parsing
 * (getArgs, Genrich.c:5718), SAM/BAM decode (readSAM 4468, readBAM 4983), mate
 * pairing and fragment inference (parseAlign 4141, processAlns 3187), interval
 * transforms (saveFragment 2754, saveFragAtac 2728, saveUnpair 2689) and the text
 * writers (printPeak 885, printInterval 770, printPile 1697, printBED 2497).
 * The hot path between them is libgenrich_cuda.so (include/genrich_cuda.h).
 */
#ifndef GB_HOST_H
#define GB_HOST_H
#include <stdbool.h>
#include <stdint.h>
#include <pthread.h>
#include <stdio.h>
#include <zlib.h>
#include "../../include/genrich_cuda.h"

#define GB_VERSION   "0.6.2-b200"
#define GB_MAX_LINE  65520     /* MAX_SIZE, Genrich.h:16 */
#define GB_MAX_ALNS  128       /* MAX_ALNS, Genrich.h:17 */
#define GB_NOSCORE   (-3.402823466e+38F)   /* NOSCORE = -FLT_MAX, Genrich.h:43 */

typedef struct {
  char* name;
  uint32_t len;
  bool skip;        /* -e */
  bool save;        /* present in the current replicate's experimental header */
  bool ever_saved;
} HChrom;

typedef struct {
  HChrom* c;
  int n;
} HChromTab;

/* one alignment of the current read name (Aln, Genrich.h:203-214) */
typedef struct {
  int chrom;
  uint32_t pos[2];
  float score;
  bool primary, paired, full, first, strand;
} HAln;

typedef struct {        /* unpaired alignment deferred for -x */
  int chrom;
  uint32_t pos[2];
  bool strand;
  uint8_t count;
  char* name;
} HUnpaired;

/* one read name's alignment set, kept until the end of the file for -r (Read, Genrich.h:216-229) */
typedef struct {
  char* name;
  uint16_t qual;        /* sum of the quality scores (both mates for pairs / discordant sets) */
  bool first;           /* singleton sets: which mate */
  float score, score_r2;
  HAln* aln;            /* paired: complete pairs within as_diff of the best, positions ordered;
                           discordant: the R1 alignments; singleton: the alignments */
  HAln* aln_r2;         /* discordant: the R2 alignments */
  uint8_t naln, naln_r2;
} HRead;

typedef struct {
  HRead* r;
  size_t n, cap;
} HReadList;

typedef struct {
  /* files */
  char *in_files, *ctrl_files, *out_file, *log_file, *pile_file, *bed_file, *xchrom, *dups_file;
  bool dups_opt;        /* -r */
  /* options (same meaning and defaults as getArgs 5720-5733) */
  bool gz_out, single_opt, extend_opt, avg_ext_opt, atac_opt, atac_adj, qval_opt;
  bool peaks_opt, sort_opt, verbose;
  int extend, min_mapq, min_len, max_gap, atac_len5, atac_len3;
  float as_diff, pqvalue, min_auc;
  uint64_t genome_len;
  int device;
  int threads;          /* host threads decoding one plain SAM file (--threads; 1 = the sequential path) */
  int gpus;             /* devices device .. device + gpus - 1, chromosomes sharded over them (--gpus) */
} HOpts;

typedef struct {        /* per-file counters (logCounts 5295) */
  uint64_t count, unmapped, supp, skipped, low_mapq, paired, sec_pair, orphan;
  uint64_t single, sec_single, single_pr, paired_pr, err_count;
  uint64_t count_pr, dups_pr, count_dc, dups_dc, count_sn, dups_sn;   /* -r */
  double total_len;
} HCounts;

/* pinned buffers of interval records: the 8-byte GR_PACK form for whatever fits it
 * (half the PCIe bytes), int32 x 4 for the rest */
typedef struct {
  int32_t* recs;
  size_t n, cap;
  uint64_t* pk;
  size_t npk, cap_pk;
} HIvBuf;

/* text sink: plain FILE or gzip */
typedef struct {
  FILE* f;
  gzFile gz;
} HOut;

/* input stream: plain FILE or gzip/BGZF */
typedef struct {
  FILE* f;
  gzFile gz;
  bool is_gz, is_bam;
} HIn;

/* -v warnings of a decode worker, replayed in file order once the workers are done */
typedef struct {
  char** msg;
  uint8_t* counted;     /* 1: one of the messages the reference stops printing after MAX_ALNS of them */
  size_t n, cap;
} HWarnLog;

/* Arrival order.  The engine applies the reference's int16 saturation rule (saveInterval 2558-2573) in
 * the order the records reach it, so they reach it in FILE order: decode workers push one after the
 * other (the worker whose turn it is streams its full buffers, the later ones park theirs). */
typedef struct {
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int turn;             /* piece whose records may go to the engine now */
} HOrder;
typedef struct { int ctx; int kind; void* data; size_t n; } HParked;   /* kind 0: GR_PACK words, 1: int32 x 4 */

/* -v lines for the records the engine dropped (gr_sample_skipped): a second, sequential decode of the
 * file that pushes nothing and names the records whose arrival index is listed */
typedef struct {
  const uint64_t** list;   /* per context: (arrival index << 1) | underflow, ascending */
  uint64_t* n;             /* per context: entries */
  uint64_t* pos;           /* per context: next entry */
  uint64_t* arrival;       /* per context: records emitted so far */
} HLookup;

/* state of one input file being decoded (or of one worker's share of it) */
typedef struct {
  const HOpts* opt;
  HChromTab* tab;
  /* one engine context per device; owner[chrom] says which one holds a chromosome (all 0 with one
   * device), bufs[k] collects the records bound for context k */
  int nctx;
  gr_ctx** ctxs;
  const int* owner;
  HIvBuf* bufs;
  HOut* bed;            /* -b, may be NULL */
  bool ctrl;
  int sample;
  HCounts cnt;
  HAln aln[GB_MAX_ALNS];
  int naln;
  char read_name[GB_MAX_ALNS + 1];
  HUnpaired* unp;       /* -x */
  size_t n_unp, cap_unp;
  /* -r: quality sums of the current read name, the alignment sets of the file, the -R sink */
  uint16_t qual_r1, qual_r2;
  HReadList rd_pr, rd_dc, rd_sn;
  HOut* dups;
  HWarnLog* wlog;       /* NULL: warnings go to stderr as they arise */
  HOrder* order;        /* NULL: the only submitter (sequential decode) */
  int piece;            /* this worker's place in the file */
  HParked* parked; size_t n_parked, cap_parked;
  HLookup* lookup;      /* not NULL: the naming pass for dropped records, nothing is pushed or printed but its lines */
  int last_chrom;       /* one-entry cache of the reference-name lookup */
} HDecode;

/* gb_util.c */
void gb_die(const char* msg, const char* suffix);          /* "Error! <msg><suffix>" + exit(1), like error() 78 */
void* gb_alloc(size_t n);
void* gb_realloc(void* p, size_t n);
int gb_parse_int(const char* s);
float gb_parse_float(const char* s);
void gb_out_open(HOut* o, const char* path, bool gz);      /* openWrite 5076 */
void gb_out_close(HOut* o, const char* path);
void gb_out_printf(HOut* o, const char* fmt, ...);
bool gb_in_open(HIn* in, const char* path);                /* openRead 5132 + checkBAM 5107 */
void gb_in_close(HIn* in, const char* path);
char* gb_in_gets(HIn* in, char* line, int size);

/* gb_decode.c */
int gb_chrom_add(HChromTab* t, const char* name, uint32_t len, bool ctrl, const HOpts* opt);  /* saveChrom 4220 */
int gb_chrom_find(const HChromTab* t, const char* name);
int gb_peaks_only(const HOpts* o, char* xfile, float thr);  /* -P: findPeaksOnly 5243 / callPeaksLog 1277 (gb_peaksonly.c) */
void gb_scan_header(const char* path, HChromTab* tab, bool ctrl, const HOpts* opt);  /* header-only pass */
void gb_decode_file(HDecode* d, const char* path);          /* readSAM 4468 / readBAM 4983 */
void gb_flush_intervals(HDecode* d);
void gb_finish_piece(HDecode* d);                             /* a decode worker is done: its parked records go out in turn */

/* gb_frag.c */
bool gb_parse_align(HDecode* d, uint16_t flag, int chrom, uint32_t pos, int length, uint32_t pnext, float score,
                    const char* qual, int qual_len, int qual_offset);
void gb_process_alns(HDecode* d, const char* qname);
void gb_process_avg_ext(HDecode* d);
bool gb_emit_interval(HDecode* d, int chrom, int64_t start, int64_t end, const char* qname, uint8_t count);
void gb_warn(HDecode* d, bool counted, const char* msg);   /* -v warning: now, or logged for the file-order replay */
int gb_do_pairs(HDecode* d, const char* qname, const HAln* aln, int naln, float best);      /* processPair 3122 */
int gb_do_singles(HDecode* d, const char* qname, HAln* aln, int naln, float best, bool first,
                  bool extend_opt, int extend, bool defer);                                /* processSingle 3019 */

/* gb_dups.c: -r PCR duplicate removal (saveAlns 2940, findDups 3949) */
void gb_save_alns(HDecode* d, const char* qname, bool pair, bool s1, bool s2, float best_pr, float best_r1, float best_r2);
void gb_find_dups(HDecode* d);

#endif
