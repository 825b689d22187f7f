/* Deterministic synthetic workload generator for genrich-b200.
 *
 * Produces, from one counter-based RNG stream, either
 *   (a) fragment records (chrom, start, end, count) as int32 x 4 -- what the
 *       reference hands to saveFragment()/saveInterval() (Genrich.c:2754, 2516)
 *       after pairing and multimap weighting, or
 *   (b) the queryname-sorted SAM text that makes the reference derive exactly
 *       those fragments (flags 99/147, +256 for secondary placements,
 *       SEQ=* QUAL=* CIGAR=<read_len>M, PNEXT of R1 == POS of R2 as the mate
 *       match at Genrich.c:4182-4185 requires).
 * Both views come from the same per-fragment draw, so a file written with
 * synth_write_sam() and a buffer filled by synth_fragments() agree record for
 * record.  This is workload tooling, not part of the product path.
 */
#ifndef GEN_SYNTH_H
#define GEN_SYNTH_H
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct synth_params {
  uint64_t seed;
  int32_t  nchrom;
  const uint32_t* chrom_len;   /* nchrom lengths, each < 2^31 */
  uint64_t nfrag;              /* read pairs (templates) */
  double   enrich;             /* fraction of templates drawn around peak centres */
  uint32_t peak_spacing;       /* bp between peak centres (global coordinate) */
  double   peak_sigma;         /* std-dev of fragment midpoint around a centre */
  int32_t  frag_min, frag_max; /* fragment length range, inclusive */
  int32_t  read_len;           /* CIGAR <read_len>M, <= frag_min */
  double   multimap_frac;      /* fraction of templates with k > 1 placements */
  int32_t  multimap_max;       /* k drawn uniformly in [2, multimap_max] */
} synth_params;

/* Number of int32x4 records synth_fragments() will write (sum over templates of
 * the kept placement count: k -> k for k in {1,2,3,4,5,6,8,10}, 7->6, 9->8,
 * >10 -> 10; Genrich.c:3010, 3113). */
uint64_t synth_count_records(const synth_params* p);

/* Fill out[4*i..4*i+3] = chrom, start, end, count for templates
 * [first, first+n).  Returns records written.  Thread-safe; any sub-range gives
 * the same records as the full run (counter-based RNG). */
uint64_t synth_fragments(const synth_params* p, uint64_t first, uint64_t n,
                         int32_t* out);

/* Write header (@HD SO:queryname, @SQ chr1..) and all alignment lines. */
int synth_write_sam(const synth_params* p, FILE* f);

#ifdef __cplusplus
}
#endif
#endif
