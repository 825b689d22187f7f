// gr_api.cu -- the C-ABI of include/genrich_cuda.h: context, device memory,
// stage orchestration (the multi-stage pipeline that replaces runProgram's
// per-chromosome loop, Genrich.c:5460-5607), timing.  Host code only; every stage
// is a CUDA kernel from gr_dense.cu / gr_interval.cu / gr_bh.cu / gr_peaks.cu.
// There is no CPU path: without a device gr_create fails with GR_ERR_NODEVICE.
#include "../../include/genrich_cuda.h"
#include "gr_common.cuh"
#include "gr_internal.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <chrono>
#include <vector>

unsigned long long g_gr_launches = 0;

// ---------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {              // retry with the exact size
      cudaGetLastError();
      want = bytes;
      e = cudaMalloc(&p, want);
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};

struct Replicate {
  DevBuf bmU, rankU, pEnd, pVal, pExpt, pCtrl, chrom_start, present;
  DevBuf cnt;                 // device: [0] number of p intervals, [1] number of control intervals
  u64 n = 0;                  // host mirror of cnt[0] (valid when the context is not lagging)
  u64 n_upper = 0;            // what the arrays were sized for
  u64 n_ctrl = 0;
  bool has_ctrl = false;
  std::vector<u64> chrom_start_h;
  std::vector<uint8_t> present_h;
  bool has_cols = false;
  void release() {
    bmU.release(); rankU.release(); pEnd.release(); pVal.release(); pExpt.release();
    pCtrl.release(); chrom_start.release(); present.release(); cnt.release();
  }
};

struct StageRec { const char* name; cudaEvent_t a, b; u64 bytes; };

enum Filling { FILL_NONE = -1, FILL_EXPT = 0, FILL_CTRL = 1 };

struct gr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, copy = nullptr;
  gr_params par;
  std::string detail;

  // chromosome table / layout
  int nchrom = 0;
  std::vector<u32> len;
  std::vector<uint8_t> skip, owned, save, flags;
  std::vector<u64> off;
  std::vector<int> blk2chrom;
  u64 T = 0, nblocks = 0;
  DevBuf d_off, d_len, d_flags, d_blk2chrom;
  DevLayout L;

  // dense working set
  DevBuf delta, bmE, bmC, rankE, rankC, rankTmp;
  DevBuf lb0, lb1, lb2, ticket, scanWs;
  DevBuf small;                  // err(int) | pad | clamped(u64) | totals[3] | misc counters
  int* d_err = nullptr; u64* d_clamped = nullptr; u64* d_totals = nullptr; u64* d_cnt = nullptr;
  void* h_small = nullptr;       // pinned mirror
  DevBuf accI, accF;             // [2][nchrom]
  void* h_acc = nullptr;         // pinned, 2*nchrom u64

  // RLE arrays of the current replicate
  DevBuf exptEnd, exptVal, exptCS, exptTot;
  DevBuf rawEnd, rawVal, rawCS, rawTot;
  DevBuf ctrlEnd, ctrlVal, ctrlCS, ctrlTot;
  u64 n_expt = 0, n_raw = 0, n_ctrl = 0;
  std::vector<u64> exptCS_h, ctrlCS_h;

  // Interval records of the sample being filled.  A push only makes the records device-resident
  // and remembers where they are; gr_sample_pileup consumes all segments at once (it has to
  // know the per-block record counts of the whole sample before it can bucket them).
  struct SegBuf { DevBuf buf; cudaEvent_t freed = nullptr; };          // device copies of host pushes, recycled
  struct Prefetch { DevBuf buf; const void* host = nullptr; u64 n = 0; cudaEvent_t ready = nullptr, freed = nullptr;
                    bool live = false, in_use = false; };
  struct Segment { const void* d; u64 n; int rb; SegBuf* own; Prefetch* pf; };
  std::vector<Segment> segs;
  DevBuf unpack6;                      // GR_PACK6 segments expanded to GR_PACK words (gr_sample_pileup)
  // int16 saturation rule (k_sat_resolve): per-cell sums of the suspect blocks, one bit per record, segment table,
  // and per sample the list of dropped records (arrival index << 1 | underflow) the host fetches for its warnings
  DevBuf satCells, satBits, satSegs, satList[2];
  static const u32 SAT_LIST_CAP = 1u << 18;
  u64 n_sat[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } };   // host mirror: dropped for overflow / underflow, list entries
  std::vector<u64> sat_list_h;
  std::vector<SegBuf*> seg_free, seg_used;
  Prefetch pf[2];                      // buffers on their way ahead of their push (gr_prefetch_*)
  cudaEvent_t h_ready[2] = { nullptr, nullptr };   // pinned bounce buffers for pageable sources
  void* h_stage[2] = { nullptr, nullptr };
  int stage_next = 0;
  static const u64 STAGE_BYTES = 64ull << 20;
  cudaEvent_t ev_copy = nullptr;
  // bucketed build (large samples)
  DevBuf sbCnt, sbStart, sbCursor, sbBucket, sbSpill, sbSpillCtr;
  u64 sb_min = 1ull << 20;             // samples with fewer records use the plain scatter (GR_SB_MIN)
  int fused = 1;                       // buckets -> breaks in shared memory, no delta array in HBM (GR_FUSED=0: dense array)
  u64 fused_min = 32767;               // ... for samples of at least this many records (GR_FUSED_MIN): fewer cannot saturate a cell (k_sat_resolve)
  // -E regions (saveXBed 1144): per chromosome the merged, clamped boundary list start0,end0,start1,...
  std::vector<std::vector<u32>> bed;
  std::vector<u64> bed_bp;             // excluded bp per chromosome
  bool has_bed = false;
  DevBuf bedMarks, blkBed;             // boundary cell slots (u64); per 8192-cell block: bit 0 starts inside a region, bit 1 holds boundaries
  DevBuf bedChromMarks;                // boundaries per chromosome (u32)
  u32 n_marks = 0;
  DevBuf ccSlots, ccSkip;              // no-control pileup: break slots and SKIP flags of its intervals

  // Host mirrors of device-side results lag behind while `lag` is set: nothing on the hot path
  // waits for the device between gr_sample_begin and the peak records; whoever needs a mirror
  // (sums, interval counts, per-chromosome starts, device error bits) calls materialize().
  bool lag = false;
  bool pend_pile[2] = { false, false };
  std::vector<Replicate*> pend_reps;
  int retry_flags = 0;              // GR_DE_TABLE / GR_DE_CAP seen by the last materialize()
  u64 cap_expt = 0, cap_raw = 0;    // capacities (upper bounds of the interval counts) of the current sample arrays
  u32 pair_cap = 1u << 20;          // pair-table capacity, grown on overflow and remembered
  u32 fisher_cap = 1u << 22;        // table of distinct Fisher sums (gr_pvalues_finalize), grown on overflow and remembered
  bool pair_valid = false;          // the tables hold the last replicate's (expt, ctrl) pairs and x->slot its intervals' slots
  bool hist_by_pairs = false;       // the BH histogram was read off the pair table (gr_bh_local_hist)
  u64 head_cap = 0;                 // candidate-peak capacity of the last peak call
  DevBuf dpar;                      // device: float factor, lambda (+ pad)
  DevBuf dsums;                     // device: double[2][nchrom], per-chromosome sum(len*val) of expt / ctrl
  u64* h_mat = nullptr;             // pinned landing area of materialize(): 2 + 4 chromosome-start tables
  char* h_up = nullptr;             // pinned ring for small host -> device uploads (a pageable source would make
  size_t up_pos = 0;                // cudaMemcpyAsync wait for the stream)
  static const size_t UP_BYTES = 1u << 20;
  gr_peak* h_peaks = nullptr;       // pinned: the first PEAK_SPEC records come back with the counts
  static const u64 PEAK_SPEC = 1u << 16;

  // sample state
  int filling = FILL_NONE;
  bool have_expt = false, have_ctrl = false;
  bool delta_clean = false;         // the delta array is known to be all zero
  int zero_after = 1;               // dense scan clears behind itself (GR_SCAN_ZERO=0: memset per sample)
  u64 n_pushed = 0, n_clamped = 0;
  int last_built = 0;                  // how the last sample was laid out (2: bucketed for the fused scan)
  std::vector<double> expt_sums, ctrl_sums;

  // replicates and final arrays
  std::vector<Replicate*> reps;
  std::vector<Replicate*> pool;     // released replicates, buffers kept for reuse
  bool finalized = false;
  Replicate* comb = nullptr;       // Fisher-combined (nrep > 1)
  Replicate* fin = nullptr;        // points at reps[0] or comb
  DevBuf qVal;
  bool have_q = false;

  // tables / BH / peaks
  DevBuf tKeys, tLens, tPval, tQval, tCount, slot;
  DevBuf hk, hl, hcount;           // local histogram list
  DevBuf ghk, ghl;                 // global histogram handed over from host memory (gr_bh_set_global_host)
  std::vector<uint32_t> hk_h;      // host copies of the local list (gr_bh_local_hist_host)
  std::vector<uint64_t> hl_h;
  u64 hn = 0;
  u32 hist_cap = 0;
  DevBuf bk0, bk1, bl0, bl1, bhist, bksum, bx, bdk, bdq, bdl, bdcount;
  u64 n_distinct = 0;
  int all_q_one = 0;
  DevBuf fsum, fdf, repviews;
  DevBuf evIdx, evCount, headIdx, headCount, cand, candOk, peakOut;     // peakOut: gr_peak_slot header + records
  std::vector<gr_peak> peaks_h;
  u64 n_peaks = 0;

  // fetch caches
  std::vector<u32> f_end; std::vector<float> f_val, f_expt, f_ctrl;

  // timing
  bool timing = false;
  std::vector<StageRec> stages;
  std::vector<cudaEvent_t> ev_pool;
  u64 launches0 = 0;
  cudaEvent_t tm_a = nullptr, tm_b = nullptr;
};

static const char* kStatusText[] = {
  "ok", "bad argument or call order", "CUDA failure", "Cannot allocate memory",
  ": read aligned beyond reference end", "Experimental sample has no analyzable fragments",
  "No analyzable genome (length=0)", "Invalid pileup value (< 0)",
  "Disallowed number of alignments", "interval on an unknown or unowned chromosome",
  "Invalid df in pchisq()", "Genome length does not match p-value length",
  "no CUDA device available",
  "int16 saturation of the delta counters beyond what the path reproduces"
};

extern "C" const char* gr_strerror(int status) {
  if (status < 0 || status > GR_ERR_SATURATED) return "Unknown error";
  return kStatusText[status];
}
extern "C" const char* gr_last_error_detail(const gr_ctx* ctx) { return ctx ? ctx->detail.c_str() : ""; }

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      char b__[512];                                                                    \
      snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call,           \
               cudaGetErrorString(e__));                                                \
      x->detail = b__;                                                                  \
      cudaGetLastError();                                                               \
      return e__ == cudaErrorMemoryAllocation ? GR_ERR_MEM : GR_ERR_CUDA;               \
    }                                                                                   \
  } while (0)

#define CKL() CK(cudaGetLastError())

// GR_GAP_DEBUG: host time between probes (where does the host keep the device waiting?)
static bool g_gap_debug = getenv("GR_GAP_DEBUG") != nullptr;
static std::chrono::steady_clock::time_point g_ht_last = std::chrono::steady_clock::now();
static inline void ht_probe(const char* label) {
  if (!g_gap_debug) return;
  const auto t = std::chrono::steady_clock::now();
  fprintf(stderr, "  host +%8.3f ms  %s\n", std::chrono::duration<double, std::milli>(t - g_ht_last).count(), label);
  g_ht_last = t;
}
#define HT(label) ht_probe(label)

// ---- timing ------------------------------------------------------------------
static void stage_begin(gr_ctx* x, const char* name, u64 bytes = 0) {
  if (!x->timing) return;
  StageRec r;
  r.name = name; r.bytes = bytes;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, x->stream);
  x->stages.push_back(r);
}
static void stage_end(gr_ctx* x) {
  if (!x->timing || x->stages.empty()) return;
  cudaEventRecord(x->stages.back().b, x->stream);
}

// ---- layout ------------------------------------------------------------------
static bool chrom_active(const gr_ctx* x, int c) { return x->owned[c] && !x->skip[c] && x->save[c]; }

// small host -> device copy that does not stall the host: the bytes are parked in a pinned ring
static int upload(gr_ctx* x, void* dst, const void* src, size_t bytes) {
  if (!bytes) return GR_OK;
  if (bytes > gr_ctx::UP_BYTES / 4) {                    // too big for the ring: the plain (waiting) copy
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, x->stream));
    return GR_OK;
  }
  const size_t need = (bytes + 63) & ~(size_t)63;
  if (x->up_pos + need > gr_ctx::UP_BYTES) {             // wrap: everything parked so far must have left
    CK(cudaStreamSynchronize(x->stream));
    x->up_pos = 0;
  }
  char* slot = x->h_up + x->up_pos;
  x->up_pos += need;
  memcpy(slot, src, bytes);
  CK(cudaMemcpyAsync(dst, slot, bytes, cudaMemcpyHostToDevice, x->stream));
  return GR_OK;
}

static int upload_flags(gr_ctx* x) {
  for (int c = 0; c < x->nchrom; c++)
    x->flags[c] = (uint8_t)(((x->owned[c] && !x->skip[c]) ? GR_CF_OWNED : 0) | (x->save[c] ? GR_CF_SAVE : 0));
  return upload(x, x->d_flags.p, x->flags.data(), x->nchrom);
}

static DevRle rle_view(DevBuf& e, DevBuf& v, DevBuf& cs, DevBuf& tot) {
  DevRle r;
  r.end = e.as<u32>(); r.val = v.as<float>(); r.chrom_start = cs.as<u64>(); r.total = tot.as<u64>();
  return r;
}

extern "C" int gr_create(gr_ctx** out, const gr_chrom* chroms, int32_t nchrom,
                         const gr_params* params, int32_t device) {
  if (!out || !chroms || nchrom <= 0 || !params) return GR_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return GR_ERR_NODEVICE; }
  if (device < 0 || device >= ndev) return GR_ERR_ARG;
  gr_ctx* x = new gr_ctx();
  x->device = device;
  x->par = *params;
  x->nchrom = nchrom;
  int rc = [&]() -> int {
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&x->copy, cudaStreamNonBlocking));
    x->len.resize(nchrom); x->skip.resize(nchrom); x->owned.resize(nchrom);
    x->save.assign(nchrom, 1); x->flags.resize(nchrom); x->off.assign(nchrom, ~0ull);
    u64 T = 0;
    for (int c = 0; c < nchrom; c++) {
      x->len[c] = chroms[c].len;
      x->skip[c] = chroms[c].skip;
      x->owned[c] = chroms[c].owned;
      if (chroms[c].len >= 0x80000000u) return GR_ERR_ARG;
      if (x->owned[c] && !x->skip[c]) {
        x->off[c] = T;
        const u64 cells = (u64)x->len[c] + 1;
        const u64 blocks = (cells + GR_BLOCK_SLOTS - 1) / GR_BLOCK_SLOTS;
        for (u64 b = 0; b < blocks; b++) x->blk2chrom.push_back(c);
        T += blocks * GR_BLOCK_SLOTS;
      }
    }
    if (!T) return GR_ERR_GENOME;
    x->T = T;
    x->nblocks = T / GR_BLOCK_SLOTS;
    if (T / GR_SCAN_TILE >= 0xffffffffull) return GR_ERR_ARG;
    CK(x->d_off.ensure(nchrom * sizeof(u64)));
    CK(x->d_len.ensure(nchrom * sizeof(u32)));
    CK(x->d_flags.ensure(nchrom));
    CK(x->d_blk2chrom.ensure(x->nblocks * sizeof(int)));
    CK(cudaMemcpy(x->d_off.p, x->off.data(), nchrom * sizeof(u64), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(x->d_len.p, x->len.data(), nchrom * sizeof(u32), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(x->d_blk2chrom.p, x->blk2chrom.data(), x->nblocks * sizeof(int), cudaMemcpyHostToDevice));
    x->L.nchrom = nchrom; x->L.T = T; x->L.nblocks = x->nblocks;
    x->L.off = x->d_off.as<u64>(); x->L.len = x->d_len.as<u32>();
    x->L.flags = x->d_flags.as<uint8_t>(); x->L.blk2chrom = x->d_blk2chrom.as<int>();
    CK(cudaMallocHost((void**)&x->h_up, gr_ctx::UP_BYTES));
    int r = upload_flags(x);
    if (r) return r;

    CK(x->delta.ensure(T * sizeof(int32_t)));
    CK(x->bmE.ensure(T / 8));
    CK(x->bmC.ensure(T / 8));
    CK(x->rankE.ensure(x->nblocks * sizeof(u64)));
    CK(x->rankC.ensure(x->nblocks * sizeof(u64)));
    CK(x->rankTmp.ensure(x->nblocks * sizeof(u64)));
    const u64 ntile = T / GR_SCAN_TILE;
    CK(x->lb0.ensure(ntile * sizeof(u64)));
    CK(x->lb1.ensure(ntile * sizeof(u64)));
    CK(x->lb2.ensure(ntile * sizeof(u64)));
    CK(x->ticket.ensure(64));
    CK(x->small.ensure(256));
    CK(cudaMemset(x->small.p, 0, 256));
    x->d_err = x->small.as<int>();
    x->d_clamped = (u64*)((char*)x->small.p + 8);
    x->d_totals = (u64*)((char*)x->small.p + 16);       // 3 entries
    x->d_cnt = (u64*)((char*)x->small.p + 64);          // 8 scratch counters
    CK(cudaMallocHost(&x->h_small, 256));
    CK(x->dpar.ensure(64));
    CK(x->dsums.ensure(2 * nchrom * sizeof(double)));
    CK(cudaMemsetAsync(x->dsums.p, 0, 2 * nchrom * sizeof(double), x->stream));
    CK(cudaMallocHost((void**)&x->h_peaks, gr_ctx::PEAK_SPEC * sizeof(gr_peak)));
    CK(cudaMallocHost((void**)&x->h_mat, 6 * (size_t)(nchrom + 3) * sizeof(u64)));
    CK(x->accI.ensure(2 * nchrom * sizeof(u64)));
    CK(x->accF.ensure(2 * nchrom * sizeof(u64)));
    CK(cudaMallocHost(&x->h_acc, 4 * nchrom * sizeof(u64)));
    const size_t cs = (nchrom + 1) * sizeof(u64);
    CK(x->exptCS.ensure(cs)); CK(x->rawCS.ensure(cs)); CK(x->ctrlCS.ensure(cs));
    CK(x->exptTot.ensure(8)); CK(x->rawTot.ensure(8)); CK(x->ctrlTot.ensure(8));
    x->expt_sums.assign(nchrom, 0.0);
    x->ctrl_sums.assign(nchrom, 0.0);
    { const char* e = getenv("GR_SCAN_ZERO"); if (e) x->zero_after = atoi(e) != 0; }
    { const char* e = getenv("GR_SB_MIN"); if (e) x->sb_min = strtoull(e, nullptr, 10); }
    { const char* e = getenv("GR_FUSED"); if (e) x->fused = atoi(e) != 0; }
    { const char* e = getenv("GR_FUSED_MIN"); if (e) x->fused_min = strtoull(e, nullptr, 10); }
    // test knobs: start the optimistic capacities small enough to exercise the retry paths
    { const char* e = getenv("GR_PAIR_CAP"); if (e) { u32 v = (u32)strtoul(e, nullptr, 10); u32 c = 64; while (c < v) c <<= 1; x->pair_cap = c; } }
    { const char* e = getenv("GR_HEAD_CAP"); if (e) x->head_cap = strtoull(e, nullptr, 10); }
    { const char* e = getenv("GR_FISHER_CAP"); if (e) { u32 v = (u32)strtoul(e, nullptr, 10); u32 c = 64; while (c < v) c <<= 1; x->fisher_cap = c; } }
    CK(cudaEventCreateWithFlags(&x->ev_copy, cudaEventDisableTiming));
    CK(cudaStreamSynchronize(x->stream));
    return GR_OK;
  }();
  if (rc != GR_OK) {
    // keep the detail reachable: hand the context back only on success
    fprintf(stderr, "gr_create: %s %s\n", gr_strerror(rc), x->detail.c_str());
    gr_destroy(x);
    return rc;
  }
  *out = x;
  return GR_OK;
}

static void free_reps(gr_ctx* x, bool keep_buffers = false) {
  for (auto* r : x->reps) {
    if (keep_buffers) x->pool.push_back(r);
    else { r->release(); delete r; }
  }
  x->reps.clear();
  if (x->comb) {
    if (keep_buffers) x->pool.push_back(x->comb);
    else { x->comb->release(); delete x->comb; }
    x->comb = nullptr;
  }
  if (!keep_buffers) {
    for (auto* r : x->pool) { r->release(); delete r; }
    x->pool.clear();
  }
  x->fin = nullptr;
}
static Replicate* new_replicate(gr_ctx* x) {
  if (!x->pool.empty()) {
    Replicate* r = x->pool.back();
    x->pool.pop_back();
    r->n = 0; r->has_cols = false;
    return r;
  }
  return new Replicate();
}

extern "C" void gr_destroy(gr_ctx* x) {
  if (!x) return;
  cudaSetDevice(x->device);
  if (x->stream) cudaStreamSynchronize(x->stream);
  free_reps(x);
  DevBuf* all[] = { &x->d_off, &x->d_len, &x->d_flags, &x->d_blk2chrom, &x->delta, &x->bmE, &x->bmC,
    &x->rankE, &x->rankC, &x->rankTmp, &x->lb0, &x->lb1, &x->lb2, &x->ticket, &x->scanWs, &x->small, &x->accI,
    &x->accF, &x->exptEnd, &x->exptVal, &x->exptCS, &x->exptTot, &x->rawEnd, &x->rawVal, &x->rawCS,
    &x->rawTot, &x->ctrlEnd, &x->ctrlVal, &x->ctrlCS, &x->ctrlTot, &x->sbCnt, &x->sbStart, &x->sbCursor, &x->sbBucket, &x->sbSpill, &x->sbSpillCtr, &x->bedMarks, &x->blkBed, &x->bedChromMarks, &x->ccSlots, &x->ccSkip,
    &x->qVal, &x->tKeys, &x->tLens, &x->tPval, &x->tQval, &x->tCount, &x->slot, &x->hk, &x->hl,
    &x->hcount, &x->bk0, &x->bk1, &x->bl0, &x->bl1, &x->bhist, &x->bksum, &x->bx, &x->bdk, &x->bdq,
    &x->bdl, &x->bdcount, &x->fsum, &x->fdf, &x->repviews, &x->evIdx, &x->evCount, &x->headIdx,
    &x->headCount, &x->cand, &x->candOk, &x->peakOut };
  for (DevBuf* b : all) b->release();
  if (x->h_small) cudaFreeHost(x->h_small);
  if (x->h_peaks) cudaFreeHost(x->h_peaks);
  if (x->h_mat) cudaFreeHost(x->h_mat);
  if (x->h_up) cudaFreeHost(x->h_up);
  x->dpar.release();
  x->dsums.release();
  x->unpack6.release();
  x->satCells.release(); x->satBits.release(); x->satSegs.release(); x->satList[0].release(); x->satList[1].release();
  x->ghk.release();
  x->ghl.release();
  if (x->h_acc) cudaFreeHost(x->h_acc);
  for (int i = 0; i < 2; i++) {
    if (x->h_stage[i]) cudaFreeHost(x->h_stage[i]);
    if (x->h_ready[i]) cudaEventDestroy(x->h_ready[i]);
  }
  if (x->ev_copy) cudaEventDestroy(x->ev_copy);
  for (auto* v : { &x->seg_free, &x->seg_used })
    for (auto* sb : *v) { sb->buf.release(); if (sb->freed) cudaEventDestroy(sb->freed); delete sb; }
  for (auto& s : x->stages) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
  if (x->tm_a) { cudaEventDestroy(x->tm_a); cudaEventDestroy(x->tm_b); }
  for (auto& p : x->pf) { p.buf.release(); if (p.ready) { cudaEventDestroy(p.ready); cudaEventDestroy(p.freed); } }
  if (x->stream) cudaStreamDestroy(x->stream);
  if (x->copy) cudaStreamDestroy(x->copy);
  delete x;
}

extern "C" int gr_set_params(gr_ctx* x, const gr_params* p) {
  if (!x || !p) return GR_ERR_ARG;
  x->par = *p;
  x->have_q = false;
  return GR_OK;
}

// -E: the BED records as read (loadBED 5187 has checked start < end), any order, overlapping or
// not.  Per chromosome exactly what saveXBed 1144-1206 does: records starting beyond the end
// are dropped, the rest sorted by start, ends clamped to the chromosome length, overlapping or
// touching regions merged.  Must precede the first sample.
extern "C" int gr_set_exclusions(gr_ctx* x, const int32_t* chrom, const uint32_t* start, const uint32_t* end, uint64_t n) {
  if (!x || (n && (!chrom || !start || !end))) return GR_ERR_ARG;
  if (!x->reps.empty() || x->filling != FILL_NONE || x->have_expt) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  for (u64 i = 0; i < n; i++)
    if (chrom[i] < 0 || chrom[i] >= x->nchrom || end[i] <= start[i]) return GR_ERR_ARG;
  const int nc = x->nchrom;
  x->bed.assign(nc, std::vector<u32>());
  x->bed_bp.assign(nc, 0);
  x->has_bed = false;
  std::vector<u64> marks;
  std::vector<u32> cmarks(nc, 0);
  std::vector<uint8_t> blk(x->nblocks, 0);
  for (int c = 0; c < nc; c++) {
    std::vector<u32>& b = x->bed[c];
    const u32 len = x->len[c];
    for (u64 i = 0; i < n; i++) {
      if (chrom[i] != c || start[i] >= len) continue;          // 1151-1160
      size_t j = 0;
      while (j < b.size() && start[i] > b[j]) j += 2;          // 1163-1176: before the first start >= this one
      b.insert(b.begin() + j, { start[i], end[i] });
    }
    size_t i = 0;
    while (i < b.size()) {                                     // 1181-1204
      if (b[i + 1] > len) b[i + 1] = len;
      if (i && b[i] <= b[i - 1]) {
        if (b[i + 1] > b[i - 1]) b[i - 1] = b[i + 1];
        b.erase(b.begin() + i, b.begin() + i + 2);
      } else
        i += 2;
    }
    for (size_t k = 0; k < b.size(); k += 2) x->bed_bp[c] += b[k + 1] - b[k];
    if (b.empty()) continue;
    x->has_bed = true;
    if (x->off[c] == ~0ull) continue;                          // not computed here
    // boundaries that close an interval (1 <= b < len) become marks; per block the region state at its start
    const u64 blk0 = x->off[c] >> GR_BLOCK_SHIFT, nblk = ((u64)len + 1 + GR_BLOCK_SLOTS - 1) / GR_BLOCK_SLOTS;
    size_t k = 0;                                              // boundaries before the block's first counted position
    for (u64 q = 0; q < nblk; q++) {
      const u64 first = q ? q * GR_BLOCK_SLOTS : 1;            // boundary 0 opens a region but closes no interval
      while (k < b.size() && b[k] < first) k++;
      uint8_t v = (uint8_t)(k & 1);
      for (size_t m = k; m < b.size() && b[m] < (q + 1) * GR_BLOCK_SLOTS; m++)
        if (b[m] >= 1 && b[m] < len) { v |= 2; break; }
      blk[blk0 + q] = v;
    }
    for (u32 pos : b)
      if (pos >= 1 && pos < len) { marks.push_back(x->off[c] + pos); cmarks[c]++; }
  }
  x->n_marks = (u32)marks.size();
  if (x->has_bed) {
    CK(x->blkBed.ensure(x->nblocks));
    CK(cudaMemcpy(x->blkBed.p, blk.data(), x->nblocks, cudaMemcpyHostToDevice));
    CK(x->bedChromMarks.ensure(2 * nc * sizeof(u32)));           // boundaries per chromosome, then the "has had reads" words
    CK(cudaMemcpy(x->bedChromMarks.p, cmarks.data(), nc * sizeof(u32), cudaMemcpyHostToDevice));
    CK(cudaMemset((char*)x->bedChromMarks.p + nc * sizeof(u32), 0, nc * sizeof(u32)));
    if (x->n_marks) {
      CK(x->bedMarks.ensure(marks.size() * sizeof(u64)));
      CK(cudaMemcpy(x->bedMarks.p, marks.data(), marks.size() * sizeof(u64), cudaMemcpyHostToDevice));
    }
  }
  return GR_OK;
}

extern "C" int gr_excluded_bp(gr_ctx* x, uint64_t* per_chrom) {
  if (!x || !per_chrom) return GR_ERR_ARG;
  for (int c = 0; c < x->nchrom; c++) per_chrom[c] = x->has_bed ? x->bed_bp[c] : 0;
  return GR_OK;
}

extern "C" int gr_reset(gr_ctx* x) {
  if (!x) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  // no device round trip: the replicate buffers go back to the pool and are reused in stream order
  free_reps(x, true);
  x->pend_reps.clear();
  x->pend_pile[0] = x->pend_pile[1] = false;
  x->finalized = false; x->have_q = false; x->have_expt = x->have_ctrl = false;
  x->filling = FILL_NONE; x->hist_cap = 0; x->peaks_h.clear(); x->pair_valid = false;
  if (x->has_bed)                                        // a new run: no chromosome has had reads yet (k_fb_scan)
    CK(cudaMemsetAsync((char*)x->bedChromMarks.p + (size_t)x->nchrom * sizeof(u32), 0, (size_t)x->nchrom * sizeof(u32), x->stream));
  return GR_OK;
}

static int map_dev_err(int e) {
  if (e & GR_DE_CHROM) return GR_ERR_CHROM;
  if (e & GR_DE_COUNT) return GR_ERR_COUNT;
  if (e & GR_DE_POS) return GR_ERR_POS;
  if (e & (GR_DE_PILE | GR_DE_TAIL)) return GR_ERR_PILE;
  if (e & GR_DE_SAT) return GR_ERR_SATURATED;
  if (e & GR_DE_EXPT) return GR_ERR_EXPT;
  return GR_OK;
}

// ---- seam IN -------------------------------------------------------------------
extern "C" int gr_sample_begin(gr_ctx* x, int32_t is_ctrl, const uint8_t* save) {
  HT("sample_begin: enter");
  if (!x) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  if (is_ctrl && !x->have_expt) return GR_ERR_ARG;
  if (!is_ctrl) {
    for (int c = 0; c < x->nchrom; c++) x->save[c] = save ? (save[c] != 0) : 1;
    int r = upload_flags(x);
    if (r) return r;
    x->have_expt = false;
    x->n_clamped = 0;
    CK(cudaMemsetAsync(x->d_clamped, 0, sizeof(u64), x->stream));
  }
  x->have_ctrl = false;
  // records of an abandoned sample are dropped
  for (auto& g : x->segs) if (g.pf) g.pf->in_use = false;
  for (auto* b : x->seg_used) x->seg_free.push_back(b);
  x->seg_used.clear();
  x->segs.clear();
  x->filling = is_ctrl ? FILL_CTRL : FILL_EXPT;
  x->n_pushed = 0;
  return GR_OK;
}

// rb: bytes per record -- 16 (int32 x 4) or 8 (GR_PACK)
static int push_device(gr_ctx* x, const void* d_recs, u64 n, int rb) {
  if (!x || x->filling == FILL_NONE || (!d_recs && n)) return GR_ERR_ARG;
  if (n) x->segs.push_back({ d_recs, n, rb, nullptr, nullptr });      // the caller keeps it unchanged until the pileup
  x->n_pushed += n;
  return GR_OK;
}

static int prefetch_any(gr_ctx* x, const void* recs, u64 n, int rb) {
  if (!x || !recs || !n) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, recs) != cudaSuccess || at.type != cudaMemoryTypeHost) {
    cudaGetLastError();
    return GR_OK;                              // not pinned: nothing to gain, the push will stage it
  }
  for (auto& p : x->pf) {
    if (p.live || p.in_use) continue;
    if (!p.ready) {
      CK(cudaEventCreateWithFlags(&p.ready, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&p.freed, cudaEventDisableTiming));
    } else
      CK(cudaStreamWaitEvent(x->copy, p.freed, 0));     // the pileup that last read this buffer is done
    CK(p.buf.ensure(n * rb));
    CK(cudaMemcpyAsync(p.buf.p, recs, n * rb, cudaMemcpyHostToDevice, x->copy));
    CK(cudaEventRecord(p.ready, x->copy));
    p.host = recs; p.n = n; p.live = true;
    return GR_OK;
  }
  return GR_OK;                                // both slots busy: ignored
}

static int push_any(gr_ctx* x, const void* recs, u64 n, int rb) {
  if (!x || x->filling == FILL_NONE || (!recs && n)) return GR_ERR_ARG;
  if (!n) return GR_OK;
  CK(cudaSetDevice(x->device));
  for (auto& p : x->pf)
    if (p.live && p.host == recs && p.n == n) {          // already on its way (gr_prefetch_*)
      CK(cudaStreamWaitEvent(x->stream, p.ready, 0));
      p.live = false; p.in_use = true;
      x->segs.push_back({ p.buf.p, n, rb, nullptr, &p });
      x->n_pushed += n;
      return GR_OK;
    }
  cudaPointerAttributes at;
  bool pinned = false;
  if (cudaPointerGetAttributes(&at, recs) == cudaSuccess) {
    if (at.type == cudaMemoryTypeDevice) return push_device(x, recs, n, rb);
    pinned = at.type == cudaMemoryTypeHost;
  } else
    cudaGetLastError();
  // a device copy of our own, kept until the pileup
  gr_ctx::SegBuf* sb = nullptr;
  if (!x->seg_free.empty()) { sb = x->seg_free.back(); x->seg_free.pop_back(); }
  else { sb = new gr_ctx::SegBuf(); if (cudaEventCreateWithFlags(&sb->freed, cudaEventDisableTiming) != cudaSuccess) { delete sb; return GR_ERR_CUDA; } CK(cudaEventRecord(sb->freed, x->stream)); }
  x->seg_used.push_back(sb);
  CK(cudaStreamWaitEvent(x->copy, sb->freed, 0));        // the pileup that last read it is done
  CK(sb->buf.ensure(n * rb));
  const u64 bytes = n * rb;
  if (pinned) {
    CK(cudaMemcpyAsync(sb->buf.p, recs, bytes, cudaMemcpyHostToDevice, x->copy));
  } else {
    for (u64 done = 0; done < bytes; done += gr_ctx::STAGE_BYTES) {
      const u64 m = bytes - done < gr_ctx::STAGE_BYTES ? bytes - done : gr_ctx::STAGE_BYTES;
      const int b = x->stage_next;
      x->stage_next ^= 1;
      if (!x->h_stage[b]) {
        CK(cudaMallocHost(&x->h_stage[b], gr_ctx::STAGE_BYTES));
        CK(cudaEventCreateWithFlags(&x->h_ready[b], cudaEventDisableTiming));
      } else
        CK(cudaEventSynchronize(x->h_ready[b]));         // previous H2D out of this pinned slot is done
      memcpy(x->h_stage[b], (const char*)recs + done, m);
      CK(cudaMemcpyAsync((char*)sb->buf.p + done, x->h_stage[b], m, cudaMemcpyHostToDevice, x->copy));
      CK(cudaEventRecord(x->h_ready[b], x->copy));
    }
  }
  CK(cudaEventRecord(x->ev_copy, x->copy));
  CK(cudaStreamWaitEvent(x->stream, x->ev_copy, 0));
  if (pinned) CK(cudaStreamSynchronize(x->copy));        // the caller may reuse its pinned buffer when we return
  x->segs.push_back({ sb->buf.p, n, rb, sb, nullptr });
  x->n_pushed += n;
  return GR_OK;
}

// all records of the sample -> the delta array (called by gr_sample_pileup)
// *built: 0 = plain scatter, 1 = delta array built block by block, 2 = events bucketed for the
// fused scan (the delta array is not touched)
static int consume_segments(gr_ctx* x, int* built) {
  int32_t* delta = x->delta.as<int32_t>();
  {
    // 6-byte records become 8-byte words first (one streaming pass); the bucket / scatter
    // kernels below then see nothing new.  The source buffers are released with the others at
    // the end of this function, i.e. after the expansion has read them (stream order).
    u64 n6 = 0;
    for (auto& g : x->segs) if (g.rb == 6) n6 += g.n;
    if (n6) {
      CK(x->unpack6.ensure(n6 * 8));
      stage_begin(x, "unpack6", n6 * 14);
      u64 at = 0;
      for (auto& g : x->segs) {
        if (g.rb != 6) continue;
        u64* out = x->unpack6.as<u64>() + at;
        launch_unpack6(x->stream, x->L, g.d, g.n, out, x->d_err);
        g.d = out; g.rb = 8;
        at += g.n;
      }
      CKL();
      stage_end(x);
    }
  }
  const bool fb_ok = x->n_pushed < (1ull << 31) && (x->T >> 11) < (1ull << 32);
  // -E regions are built into the fused scan only: every sample, whatever its size, goes that way
  if (x->has_bed && !fb_ok) { x->detail = "-E regions with more than 2^31 records in one sample"; return GR_ERR_ARG; }
  const bool fb = fb_ok && (x->has_bed || (x->fused && x->n_pushed >= x->fused_min));
  const bool sb = !fb && x->n_pushed >= x->sb_min && x->T < (1ull << 32);
  u64 bytes = 0;
  for (auto& g : x->segs) bytes += g.n * g.rb;
  if (fb) {
    const u64 nbk = x->nblocks;
    const int ctrl = x->filling == FILL_CTRL;
    CK(x->sbCnt.ensure(nbk * 4 + 64));                   // the counters, then k_sb_scan1's words (saturation flag, entries in full blocks)
    CK(x->sbStart.ensure((nbk + 1) * 4));
    CK(x->sbCursor.ensure(nbk * 4));
    CK(x->sbBucket.ensure(x->n_pushed * 8 + (u64)x->n_marks * 4 + 16));   // at most two event entries per record
    CK(x->sbSpillCtr.ensure(4 + (nbk / 4096 + 2) * 4));  // (unused word), then the scan's chunk sums
    CK(x->satCells.ensure((size_t)SAT_MAX_BLOCKS * GR_BLOCK_SLOTS * 16));
    CK(x->satBits.ensure((x->n_pushed + 31) / 32 * 4 + 4));
    CK(x->satList[ctrl].ensure((size_t)gr_ctx::SAT_LIST_CAP * 8));
    CK(x->satSegs.ensure(x->segs.size() * sizeof(SatSeg) + 16));
    HT("consume: bucket buffers ensured");
    u32* sat_flag = x->sbCnt.as<u32>() + nbk;
    u32* sat_res = (u32*)((char*)x->small.p + 40) + 3 * ctrl;
    // the segments in arrival order (the saturation rule replays it)
    {
      std::vector<SatSeg> sg(x->segs.size());
      u64 base = 0;
      for (size_t i = 0; i < sg.size(); i++) {
        sg[i].d = x->segs[i].d; sg[i].n = x->segs[i].n; sg[i].base = base; sg[i].packed = x->segs[i].rb == 8; sg[i].pad_ = 0;
        base += x->segs[i].n;
      }
      int r = upload(x, x->satSegs.p, sg.data(), sg.size() * sizeof(SatSeg));
      if (r) return r;
    }
    const int nseg = (int)x->segs.size();
    stage_begin(x, "bucket", bytes);
    {
      CK(cudaMemsetAsync(x->sbCnt.p, 0, nbk * 4 + 64, x->stream));
      for (auto& g : x->segs)
        launch_fb_count(x->stream, x->L, g.d, g.n, g.rb == 8, x->sbCnt.as<u32>(), x->d_err, x->d_clamped);
      HT("consume: memset + count launched");
      launch_fb_marks(x->stream, x->bedMarks.as<u64>(), x->n_marks, x->sbCnt.as<u32>(), nullptr, nullptr);
      launch_sb_scan_a(x->stream, nbk, x->sbCnt.as<u32>(), x->sbSpillCtr.as<u32>() + 1, sat_flag);
      launch_sat_resolve(x->stream, sat_flag, x->satSegs.p, nseg, x->L, x->sbCnt.as<u32>(), x->sbSpillCtr.as<u32>() + 1,
                         x->satCells.p, x->satBits.as<u32>(), x->n_pushed, x->satList[ctrl].as<u64>(),
                         gr_ctx::SAT_LIST_CAP, sat_res, x->d_err);
      launch_sb_scan_b(x->stream, nbk, x->sbCnt.as<u32>(), x->sbStart.as<u32>(), x->sbCursor.as<u32>(),
                       x->sbSpillCtr.as<u32>() + 1);
      {
        u64 base = 0;
        for (auto& g : x->segs) {
          launch_fb_move(x->stream, x->L, g.d, g.n, g.rb == 8, x->sbCursor.as<u32>(), x->sbBucket.as<u32>(),
                         sat_res, x->satBits.as<u32>(), base);
          base += g.n;
        }
      }
      launch_fb_marks(x->stream, x->bedMarks.as<u64>(), x->n_marks, nullptr, x->sbCursor.as<u32>(), x->sbBucket.as<u32>());
      HT("consume: scans + move launched");
      CKL();
    }
    stage_end(x);
  } else if (sb) {
    CK(cudaMemsetAsync((char*)x->small.p + 40 + 12 * (x->filling == FILL_CTRL), 0, 12, x->stream));   // no record is dropped on this path
    CK(x->sbCnt.ensure(x->nblocks * 4));
    CK(x->sbStart.ensure((x->nblocks + 1) * 4));
    CK(x->sbCursor.ensure(x->nblocks * 4));
    CK(x->sbBucket.ensure(x->n_pushed * 4));
    CK(x->sbSpill.ensure(x->n_pushed * 16));             // worst case: every interval is long (two entries each)
    CK(x->sbSpillCtr.ensure(4 + (x->nblocks / 4096 + 2) * 4));      // spill counter, then the scan's chunk sums
    stage_begin(x, "bucket", bytes);
    CK(cudaMemsetAsync(x->sbCnt.p, 0, x->nblocks * 4, x->stream));
    for (auto& g : x->segs)
      launch_sb_count(x->stream, x->L, g.d, g.n, g.rb == 8, x->sbCnt.as<u32>(), x->d_err, x->d_clamped);
    launch_sb_scan(x->stream, x->nblocks, x->sbCnt.as<u32>(), x->sbStart.as<u32>(), x->sbCursor.as<u32>(),
                   x->sbSpillCtr.as<u32>() + 1);
    CK(cudaMemsetAsync(x->sbSpillCtr.p, 0, 4, x->stream));
    for (auto& g : x->segs)
      launch_sb_move(x->stream, x->L, g.d, g.n, g.rb == 8, x->sbCursor.as<u32>(), x->sbBucket.as<u32>(),
                     x->sbSpill.as<uint2>(), x->sbSpillCtr.as<u32>());
    CKL();
    stage_end(x);
    stage_begin(x, "build", x->T * 4);
    launch_sb_build(x->stream, x->L, x->sbBucket.as<u32>(), x->sbStart.as<u32>(), delta,
                    x->sbSpill.as<uint2>(), x->sbSpillCtr.as<u32>());
    CKL();
    stage_end(x);
    x->delta_clean = false;
  } else {
    CK(cudaMemsetAsync((char*)x->small.p + 40 + 12 * (x->filling == FILL_CTRL), 0, 12, x->stream));
    if (!x->delta_clean) {
      stage_begin(x, "memset_delta", x->T * 4);
      CK(cudaMemsetAsync(delta, 0, x->T * sizeof(int32_t), x->stream));   // runProgram 5503-5510
      stage_end(x);
    }
    stage_begin(x, "scatter", bytes);
    for (auto& g : x->segs)
      launch_scatter(x->stream, x->L, g.d, g.n, g.rb == 8, delta, x->d_err, x->d_clamped);
    CKL();
    stage_end(x);
    x->delta_clean = false;
  }
  // the buffers may be refilled once the kernels above have read them
  for (auto& g : x->segs) {
    if (g.own) CK(cudaEventRecord(g.own->freed, x->stream));
    if (g.pf) { CK(cudaEventRecord(g.pf->freed, x->stream)); g.pf->in_use = false; }
  }
  for (auto* b : x->seg_used) x->seg_free.push_back(b);
  x->seg_used.clear();
  x->segs.clear();
  *built = fb ? 2 : sb ? 1 : 0;
  return GR_OK;
}

extern "C" int gr_push_intervals_device(gr_ctx* x, const int32_t* d_recs, uint64_t n) {
  if (x) { CK(cudaSetDevice(x->device)); }
  return push_device(x, d_recs, n, 16);
}
extern "C" int gr_prefetch_intervals(gr_ctx* x, const int32_t* recs, uint64_t n) { return prefetch_any(x, recs, n, 16); }
extern "C" int gr_push_intervals(gr_ctx* x, const int32_t* recs, uint64_t n) { return push_any(x, recs, n, 16); }
extern "C" int gr_prefetch_packed(gr_ctx* x, const uint64_t* recs, uint64_t n) { return prefetch_any(x, recs, n, 8); }
extern "C" int gr_push_packed(gr_ctx* x, const uint64_t* recs, uint64_t n) { return push_any(x, recs, n, 8); }
extern "C" int gr_prefetch_packed6(gr_ctx* x, const uint16_t* recs, uint64_t n) { return prefetch_any(x, recs, n, 6); }
extern "C" int gr_push_packed6(gr_ctx* x, const uint16_t* recs, uint64_t n) {
  if (x && (x->T > (1ull << 32) || x->nchrom > 16384)) return GR_ERR_ARG;      // gr_pack6_layout said so
  return push_any(x, recs, n, 6);
}
extern "C" int gr_pack6_layout(gr_ctx* x, uint64_t* cell_offset) {
  if (!x || !cell_offset) return GR_ERR_ARG;
  if (x->T > (1ull << 32) || x->nchrom > 16384) return GR_ERR_ARG;             // the layout does not fit the format
  for (int c = 0; c < x->nchrom; c++) cell_offset[c] = x->off[c];
  return GR_OK;
}

// ---- host mirrors ------------------------------------------------------------------
// One device round trip that brings every lagging host mirror up to date and reports the
// device-side error bits.  GR_DE_TABLE / GR_DE_CAP (an optimistically sized table or buffer
// was too small) are not errors: they are left in x->retry_flags for the caller that can redo
// the stage with more room.
static int map_dev_err(int e);
static int materialize(gr_ctx* x) {
  if (!x->lag) return GR_OK;
  const int nc = x->nchrom;
  const size_t row = (size_t)nc + 3;                           // nc+1 starts, then two counts
  u64* hI = (u64*)x->h_acc;
  CK(cudaMemcpyAsync(x->h_small, x->small.p, 64, cudaMemcpyDeviceToHost, x->stream));
  for (int k = 0; k < 2; k++) {
    if (!x->pend_pile[k]) continue;
    CK(cudaMemcpyAsync(hI + 2 * k * nc, x->accI.as<u64>() + k * nc, nc * sizeof(u64), cudaMemcpyDeviceToHost, x->stream));
    CK(cudaMemcpyAsync(hI + (2 * k + 1) * nc, x->accF.as<u64>() + k * nc, nc * sizeof(u64), cudaMemcpyDeviceToHost, x->stream));
    u64* m = x->h_mat + k * row;
    CK(cudaMemcpyAsync(m, (k ? x->rawCS : x->exptCS).p, (nc + 1) * sizeof(u64), cudaMemcpyDeviceToHost, x->stream));
    CK(cudaMemcpyAsync(m + nc + 1, (k ? x->rawTot : x->exptTot).p, 8, cudaMemcpyDeviceToHost, x->stream));
  }
  // up to four replicates land in pinned memory; more (many replicates finished without asking
  // for statistics) take the slower pageable way
  for (size_t i = 0; i < x->pend_reps.size(); i++) {
    Replicate* r = x->pend_reps[i];
    if (i < 4) {
      u64* m = x->h_mat + (2 + i) * row;
      CK(cudaMemcpyAsync(m, r->chrom_start.p, (nc + 1) * sizeof(u64), cudaMemcpyDeviceToHost, x->stream));
      CK(cudaMemcpyAsync(m + nc + 1, r->cnt.p, 16, cudaMemcpyDeviceToHost, x->stream));
    } else {
      r->chrom_start_h.resize(nc + 1);
      CK(cudaMemcpy(r->chrom_start_h.data(), r->chrom_start.p, (nc + 1) * sizeof(u64), cudaMemcpyDeviceToHost));
      u64 c2[2];
      CK(cudaMemcpy(c2, r->cnt.p, 16, cudaMemcpyDeviceToHost));
      r->n = c2[0]; r->n_ctrl = c2[1];
    }
  }
  CK(cudaMemsetAsync(x->d_err, 0, sizeof(int), x->stream));     // the bits are ours now
  CK(cudaStreamSynchronize(x->stream));
  const int derr = *(int*)x->h_small;
  x->n_clamped = *(u64*)((char*)x->h_small + 8);
  for (int k = 0; k < 2; k++) {
    if (!x->pend_pile[k]) continue;
    for (int j = 0; j < 3; j++) x->n_sat[k][j] = ((const u32*)((char*)x->h_small + 40))[3 * k + j];
    std::vector<double>& sums = k ? x->ctrl_sums : x->expt_sums;
    for (int c = 0; c < nc; c++)
      sums[c] = (double)hI[2 * k * nc + c] + (double)hI[(2 * k + 1) * nc + c] * (1.0 / 1099511627776.0);
    const u64* m = x->h_mat + k * row;
    std::vector<u64>& csh = k ? x->ctrlCS_h : x->exptCS_h;
    csh.assign(m, m + nc + 1);
    (k ? x->n_raw : x->n_expt) = m[nc + 1];
    x->pend_pile[k] = false;
  }
  for (size_t i = 0; i < x->pend_reps.size() && i < 4; i++) {
    Replicate* r = x->pend_reps[i];
    const u64* m = x->h_mat + (2 + i) * row;
    r->chrom_start_h.assign(m, m + nc + 1);
    r->n = m[nc + 1];
    r->n_ctrl = m[nc + 2];
  }
  x->pend_reps.clear();
  x->lag = false;
  x->retry_flags = (x->retry_flags & GR_DE_TABLE) | (derr & (GR_DE_TABLE | GR_DE_CAP));   // a pending p-value redo stays pending
  const int hard = derr & ~(GR_DE_TABLE | GR_DE_CAP);
  if (hard) { x->filling = FILL_NONE; return map_dev_err(hard); }
  return GR_OK;
}

// ---- pileup integration ----------------------------------------------------------
// Enqueues everything; the host learns the sums / counts at the next materialize().
static int pileup_enqueue(gr_ctx* x) {
  HT("pileup_enqueue: enter");
  const bool ctrl = x->filling == FILL_CTRL;
  int nact = 0;
  for (int c = 0; c < x->nchrom; c++) nact += chrom_active(x, c);
  const u64 cap = 2 * x->n_pushed + (u64)nact + 1 + x->n_marks;
  DevBuf& E = ctrl ? x->rawEnd : x->exptEnd;
  DevBuf& V = ctrl ? x->rawVal : x->exptVal;
  DevBuf& CS = ctrl ? x->rawCS : x->exptCS;
  DevBuf& TT = ctrl ? x->rawTot : x->exptTot;
  CK(E.ensure(cap * sizeof(u32)));
  CK(V.ensure(cap * sizeof(float)));
  DevRle out = rle_view(E, V, CS, TT);
  CK(x->scanWs.ensure(dense_scan_ws_bytes(cap, x->nchrom)));
  HT("pileup_enqueue: buffers ensured");
  int built = 0;
  { int r = consume_segments(x, &built); if (r) return r; }
  x->last_built = built;
  HT("pileup_enqueue: segments consumed");
  // after the plain scatter the scan clears the cells behind itself (the next small sample
  // finds the array zero); a built array is overwritten as a whole by the next build anyway
  const int zero_after = built ? 0 : x->zero_after;
  ScanScratch sc;
  sc.ws = x->scanWs.p; sc.cap = cap;
  u32 owners = 0;
  if (built == 2) {
    stage_begin(x, "fused_scan", x->T * 4);
    owners = launch_fb_scan(x->stream, x->L, x->sbBucket.as<u32>(), x->sbStart.as<u32>(), sc,
                            (ctrl ? x->bmC : x->bmE).as<u32>(), x->d_err,
                            x->has_bed ? x->blkBed.as<uint8_t>() : nullptr,
                            x->has_bed ? x->bedChromMarks.as<u32>() : nullptr,
                            x->has_bed ? x->bedChromMarks.as<u32>() + x->nchrom : nullptr, ctrl ? 0 : 1,
                            x->sbCnt.as<u32>() + x->nblocks);
    CKL();
    stage_end(x);
  } else {
    stage_begin(x, "dense_scan", x->T * 4);
    launch_dense_scan(x->stream, x->L, x->delta.as<int32_t>(), sc,
                      (ctrl ? x->bmC : x->bmE).as<u32>(), x->d_err, zero_after);
    CKL();
    stage_end(x);
  }
  u64* aI = x->accI.as<u64>() + (ctrl ? x->nchrom : 0);
  u64* aF = x->accF.as<u64>() + (ctrl ? x->nchrom : 0);
  CK(cudaMemsetAsync(aI, 0, x->nchrom * sizeof(u64), x->stream));
  CK(cudaMemsetAsync(aF, 0, x->nchrom * sizeof(u64), x->stream));
  stage_begin(x, "scan_place", cap * 16);
  launch_scan_place(x->stream, x->L, sc, out, x->d_err, owners, ctrl ? GR_SKIP : 0.0f);
  HT("pileup_enqueue: scan + place launched");
  CKL();
  stage_end(x);
  stage_begin(x, "rle_moment", 0);
  launch_rle_moment(x->stream, out, cap, x->nchrom, aI, aF);
  launch_sums_double(x->stream, aI, aF, x->nchrom, x->dsums.as<double>() + (ctrl ? x->nchrom : 0));
  HT("pileup_enqueue: moments launched");
  CKL();
  stage_end(x);
  if (built != 2) x->delta_clean = zero_after != 0;          // the scan left the array all zero
  (ctrl ? x->cap_raw : x->cap_expt) = cap;
  x->pend_pile[ctrl ? 1 : 0] = true;
  x->lag = true;
  if (ctrl) x->have_ctrl = true; else x->have_expt = true;
  x->filling = FILL_NONE;
  return GR_OK;
}

// chrom_sums == NULL: nothing is waited for (the sums come with gr_sample_sums, or never leave
// the device: gr_replicate_finish_device)
extern "C" int gr_sample_pileup(gr_ctx* x, double* chrom_sums) {
  if (!x || x->filling == FILL_NONE) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  const bool ctrl = x->filling == FILL_CTRL;
  { int r = pileup_enqueue(x); if (r) return r; }
  if (!chrom_sums) return GR_OK;
  { int r = materialize(x); if (r) return r; }
  memcpy(chrom_sums, (ctrl ? x->ctrl_sums : x->expt_sums).data(), x->nchrom * sizeof(double));
  return GR_OK;
}

extern "C" int gr_sample_sums(gr_ctx* x, double* expt_sums, double* ctrl_sums) {
  if (!x || !x->have_expt) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  { int r = materialize(x); if (r) return r; }
  if (expt_sums) memcpy(expt_sums, x->expt_sums.data(), x->nchrom * sizeof(double));
  if (ctrl_sums) {
    if (x->have_ctrl) memcpy(ctrl_sums, x->ctrl_sums.data(), x->nchrom * sizeof(double));
    else memset(ctrl_sums, 0, x->nchrom * sizeof(double));
  }
  return GR_OK;
}

// saveInterval 2558-2573: what the reference's int16 counters made it drop from the sample, in arrival order
extern "C" int gr_sample_skipped(gr_ctx* x, int32_t is_ctrl, uint64_t* n_overflow, uint64_t* n_underflow,
                                 const uint64_t** list, uint64_t* n_list) {
  if (!x) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  { int r = materialize(x); if (r) return r; }
  const int k = is_ctrl ? 1 : 0;
  if (n_overflow) *n_overflow = x->n_sat[k][0];
  if (n_underflow) *n_underflow = x->n_sat[k][1];
  const u64 nl = x->n_sat[k][2];
  if (list) {
    x->sat_list_h.resize(nl ? nl : 1);
    if (nl) {
      CK(cudaMemcpyAsync(x->sat_list_h.data(), x->satList[k].p, nl * 8, cudaMemcpyDeviceToHost, x->stream));
      CK(cudaStreamSynchronize(x->stream));
    }
    *list = (const uint64_t*)x->sat_list_h.data();
  }
  if (n_list) *n_list = nl;
  return GR_OK;
}

// ---- p-values through the pair table ------------------------------------------------
static int table_alloc(gr_ctx* x, u32 cap) {
  CK(x->tKeys.ensure((size_t)cap * 8));
  CK(x->tLens.ensure((size_t)cap * 8));
  CK(x->tPval.ensure((size_t)cap * 4));
  CK(x->tQval.ensure((size_t)cap * 4));
  CK(x->tCount.ensure(8));
  CK(cudaMemsetAsync(x->tKeys.p, 0xff, (size_t)cap * 8, x->stream));
  CK(cudaMemsetAsync(x->tLens.p, 0, (size_t)cap * 8, x->stream));
  CK(cudaMemsetAsync(x->tCount.p, 0, 8, x->stream));
  return GR_OK;
}
static PairTable table_view(gr_ctx* x, u32 cap) {
  PairTable t;
  t.keys = x->tKeys.as<u64>(); t.lens = x->tLens.as<u64>(); t.pval = x->tPval.as<float>();
  t.qval = x->tQval.as<float>(); t.cap = cap; t.count = x->tCount.as<u32>();
  return t;
}

// Host-counted builds (BH histogram): run an insert kernel, growing the table until it stays
// under half full.
template <class F>
static int table_build(gr_ctx* x, u64 n, u32& cap_io, F insert) {
  u32 cap = cap_io;
  { int r = materialize(x); if (r) return r; }
  for (;;) {
    int r = table_alloc(x, cap);
    if (r) return r;
    insert(table_view(x, cap));
    CKL();
    x->lag = true;
    r = materialize(x);
    if (r) return r;
    if (!(x->retry_flags & GR_DE_TABLE)) break;
    x->retry_flags &= ~GR_DE_TABLE;                      // this table's, and dealt with here
    if (cap >= (1u << 30)) { x->detail = "distinct-value table overflow"; return GR_ERR_MEM; }
    cap <<= 2;
  }
  (void)n;
  cap_io = cap;
  return GR_OK;
}

// K5 for one replicate: -log10 p through the table of distinct (expt, ctrl) pairs.  The table
// capacity is a remembered guess; an overflow shows up as GR_DE_TABLE at the next materialize()
// and the stage is simply run again with a larger table (its inputs are still in place).
static int rep_stage_pvals(gr_ctx* x, Replicate* rep) {
  CK(x->slot.ensure((rep->n_upper + 1) * sizeof(u32)));
  const u64* n_dev = rep->cnt.as<u64>();
  stage_begin(x, "pval", rep->n_upper * 16);
  { int r = table_alloc(x, x->pair_cap); if (r) return r; }
  PairTable t = table_view(x, x->pair_cap);
  launch_pair_insert(x->stream, rep->pExpt.as<float>(), rep->pCtrl.as<float>(), rep->n_upper, n_dev, t,
                     x->slot.as<u32>(), x->d_err);
  launch_pair_eval(x->stream, t);
  launch_gather_f32(x->stream, t.pval, x->slot.as<u32>(), rep->n_upper, n_dev, rep->pVal.as<float>());
  CKL();
  stage_end(x);
  x->lag = true;
  x->pair_valid = true;
  return GR_OK;
}

// the part of a replicate that follows the two pileups; factor and lambda are already in x->dpar
static int replicate_tail(gr_ctx* x, bool has_ctrl) {
  HT("replicate_tail: enter");
  const int nc = x->nchrom;
  int nact = 0;
  for (int c = 0; c < nc; c++) nact += chrom_active(x, c);
  Replicate* rep = new_replicate(x);
  x->reps.push_back(rep);
  rep->has_ctrl = has_ctrl;
  CK(rep->cnt.ensure(16));
  u64 n_ctrl_upper;
  if (has_ctrl) {
    n_ctrl_upper = x->cap_raw;
    CK(x->ctrlEnd.ensure((x->cap_raw + 1) * sizeof(u32)));
    CK(x->ctrlVal.ensure((x->cap_raw + 1) * sizeof(float)));
    DevRle raw = rle_view(x->rawEnd, x->rawVal, x->rawCS, x->rawTot);
    DevRle out = rle_view(x->ctrlEnd, x->ctrlVal, x->ctrlCS, x->ctrlTot);
    CompactScratch cs;
    CK(x->lb0.ensure(((x->cap_raw + 1023) / 1024 + 1) * 8 > x->lb0.cap ? ((x->cap_raw + 1023) / 1024 + 1) * 8 : x->lb0.cap));
    cs.st = x->lb0.as<u64>(); cs.ticket = x->ticket.as<u32>();
    stage_begin(x, "ctrl_clamp", x->cap_raw * 8);
    launch_ctrl_clamp(x->stream, x->L, raw, x->cap_raw, x->dpar.as<float>(), cs, out, x->bmC.as<u32>());
    HT("replicate_tail: ctrl_clamp launched");
    CKL();
    stage_end(x);
  } else {
    // saveLambda 1838-1877: lambda over every saved chromosome, SKIP inside its -E regions
    std::vector<u32> e; std::vector<u64> cs(nc + 1), slots; std::vector<uint8_t> skp;
    for (int c = 0; c < nc; c++) {
      cs[c] = e.size();
      if (!chrom_active(x, c)) continue;
      const u32 len = x->len[c];
      bool save = true;
      if (x->has_bed)
        for (u32 b : x->bed[c]) {                              // boundary 0 opens a region and closes no interval;
          if (b >= len) break;                                 // boundary len: the last interval is closed below
          if (b >= 1) { e.push_back(b); slots.push_back(x->off[c] + b); skp.push_back(save ? 0 : 1); }
          save = !save;
        }
      e.push_back(len); slots.push_back(x->off[c] + len); skp.push_back(save ? 0 : 1);
    }
    cs[nc] = e.size();
    n_ctrl_upper = e.size();
    CK(x->ctrlEnd.ensure((e.size() + 1) * sizeof(u32)));
    CK(x->ctrlVal.ensure((e.size() + 1) * sizeof(float)));
    CK(x->ccSlots.ensure((e.size() + 1) * sizeof(u64)));
    CK(x->ccSkip.ensure(e.size() + 1));
    { int r = upload(x, x->ctrlEnd.p, e.data(), e.size() * sizeof(u32)); if (r) return r; }
    { int r = upload(x, x->ccSlots.p, slots.data(), slots.size() * sizeof(u64)); if (r) return r; }
    { int r = upload(x, x->ccSkip.p, skp.data(), skp.size()); if (r) return r; }
    { int r = upload(x, x->ctrlCS.p, cs.data(), (nc + 1) * sizeof(u64)); if (r) return r; }
    u64 tot = e.size();
    { int r = upload(x, x->ctrlTot.p, &tot, 8); if (r) return r; }
    stage_begin(x, "ctrl_const", x->T / 8);
    DevRle out = rle_view(x->ctrlEnd, x->ctrlVal, x->ctrlCS, x->ctrlTot);
    launch_ctrl_const(x->stream, x->L, x->dpar.as<float>() + 1, e.size(), out, x->bmC.as<u32>(),
                      x->ccSlots.as<u64>(), x->ccSkip.as<uint8_t>());
    CKL();
    stage_end(x);
  }

  // K4: union ranks
  RankScratch rs;
  rs.st[0] = x->lb0.as<u64>(); rs.st[1] = x->lb1.as<u64>(); rs.st[2] = x->lb2.as<u64>();
  rs.ticket = x->ticket.as<u32>();
  CK(rep->bmU.ensure(x->T / 8));
  CK(rep->rankU.ensure(x->nblocks * sizeof(u64)));
  CK(rep->chrom_start.ensure((nc + 1) * sizeof(u64)));
  CK(rep->present.ensure(nc));
  stage_begin(x, "union_rank", x->T / 4);
  launch_union_rank(x->stream, x->L, x->bmE.as<u32>(), x->bmC.as<u32>(), rs, x->rankE.as<u64>(),
                    x->rankC.as<u64>(), rep->rankU.as<u64>(), x->d_totals);
  HT("replicate_tail: union_rank launched");
  CKL();
  stage_end(x);
  // the counts of this replicate stay on the device (cnt[0] = #p intervals, cnt[1] = #control intervals)
  CK(cudaMemcpyAsync(rep->cnt.p, x->d_totals + 2, 8, cudaMemcpyDeviceToDevice, x->stream));
  CK(cudaMemcpyAsync((char*)rep->cnt.p + 8, x->ctrlTot.p, 8, cudaMemcpyDeviceToDevice, x->stream));
  const u64 np_upper = x->cap_expt + n_ctrl_upper;
  if (np_upper >= 0xfffffff0ull) { x->detail = "more than 2^32 intervals on one device"; return GR_ERR_MEM; }
  rep->n_upper = np_upper;
  CK(rep->pEnd.ensure((np_upper + 1) * sizeof(u32)));
  CK(rep->pVal.ensure((np_upper + 1) * sizeof(float)));
  CK(rep->pExpt.ensure((np_upper + 1) * sizeof(float)));
  CK(rep->pCtrl.ensure((np_upper + 1) * sizeof(float)));
  stage_begin(x, "union_emit", x->T / 4 + np_upper * 12);
  launch_union_emit(x->stream, x->L, x->bmE.as<u32>(), x->bmC.as<u32>(), x->rankE.as<u64>(),
                    x->rankC.as<u64>(), rep->rankU.as<u64>(), x->exptVal.as<float>(),
                    x->ctrlVal.as<float>(), rep->pEnd.as<u32>(), rep->pExpt.as<float>(),
                    rep->pCtrl.as<float>(), rep->bmU.as<u32>(), rep->chrom_start.as<u64>(),
                    x->d_totals + 2);
  CKL();
  HT("replicate_tail: union_emit launched");
  stage_end(x);
  { int r = rep_stage_pvals(x, rep); if (r) return r; }
  HT("replicate_tail: pvals launched");

  rep->present_h.resize(nc);
  for (int c = 0; c < nc; c++) rep->present_h[c] = chrom_active(x, c);
  { int r = upload(x, rep->present.p, rep->present_h.data(), nc); if (r) return r; }
  HT("replicate_tail: present uploaded");
  rep->has_cols = x->par.keep_pileups != 0;
  x->pend_reps.push_back(rep);
  x->lag = true;
  x->have_expt = x->have_ctrl = false;
  x->finalized = false;
  x->have_q = false;
  return GR_OK;
}

// a table that overflowed is rebuilt, four times larger, for every replicate
static int redo_pvals(gr_ctx* x) {
  while (x->retry_flags & GR_DE_TABLE) {
    if (x->pair_cap >= (1u << 30)) { x->detail = "distinct-value table overflow"; return GR_ERR_MEM; }
    x->pair_cap <<= 2;
    x->retry_flags &= ~GR_DE_TABLE;
    for (Replicate* r : x->reps) { int rc = rep_stage_pvals(x, r); if (rc) return rc; }
    int rc = materialize(x);
    if (rc) return rc;
  }
  return GR_OK;
}

static u64 replicate_genome_len(const gr_ctx* x, u64 genome_len) {
  u64 G = x->par.genome_len ? x->par.genome_len : genome_len;
  if (!G)
    for (int c = 0; c < x->nchrom; c++)
      if (chrom_active(x, c)) G += x->len[c] - (x->has_bed ? x->bed_bp[c] : 0);   // calcLambda 1819-1827
  return G;
}

static void fill_stats(gr_ctx* x, gr_sample_stats* st, double frag_len, double ctrl_frag, bool has_ctrl,
                       float lambda, float factor, u64 G) {
  const Replicate* rep = x->reps.back();
  memset(st, 0, sizeof *st);
  st->frag_len = frag_len;
  st->ctrl_frag = has_ctrl ? ctrl_frag : 0.0;
  st->lambda = lambda;
  st->factor = factor;
  st->genome_len = G;
  st->n_expt = x->n_expt;
  st->n_ctrl = rep->n_ctrl;
  st->n_pval = rep->n;
  st->n_clamped = x->n_clamped;
}

// st == NULL: nothing is waited for
extern "C" int gr_replicate_finish(gr_ctx* x, double frag_len, double ctrl_frag, int32_t has_ctrl,
                                   uint64_t genome_len, gr_sample_stats* st) {
  if (!x || !x->have_expt) return GR_ERR_ARG;
  if (has_ctrl && !x->have_ctrl) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  if (frag_len == 0.0) return GR_ERR_EXPT;                     // Genrich.c:2292
  const u64 G = replicate_genome_len(x, genome_len);
  if (!G) return GR_ERR_GENOME;
  const float lambda = (float)(frag_len / (double)G);          // 1831
  float factor = 1.0f;
  if (has_ctrl && ctrl_frag != 0.0) factor = (float)(frag_len / ctrl_frag);   // 2043-2045
  const float fl[2] = { factor, lambda };
  { int r = upload(x, x->dpar.p, fl, sizeof fl); if (r) return r; }
  { int r = replicate_tail(x, has_ctrl != 0); if (r) return r; }
  if (st) {
    int r = materialize(x);
    if (r) return r;
    r = redo_pvals(x);
    if (r) return r;
    fill_stats(x, st, frag_len, ctrl_frag, has_ctrl != 0, lambda, factor, G);
  }
  return GR_OK;
}

// Same, with the sums taken where they are: on the device (gr_sums_device), possibly after the
// caller all-reduced them across ranks in the library's stream (gr_stream).  lambda and the
// scale factor are computed by one thread exactly as calcLambda / calcFactor do (per-chromosome
// doubles added in chromosome order, Genrich.c:1831, 2043-2045); the host is not involved.
extern "C" int gr_replicate_finish_device(gr_ctx* x, int32_t has_ctrl, uint64_t genome_len) {
  if (!x || !x->have_expt) return GR_ERR_ARG;
  if (has_ctrl && !x->have_ctrl) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  const u64 G = replicate_genome_len(x, genome_len);
  if (!G) return GR_ERR_GENOME;
  launch_lambda_factor(x->stream, x->dsums.as<double>(), x->nchrom, has_ctrl != 0, G, x->dpar.as<float>(), x->d_err);
  CKL();
  return replicate_tail(x, has_ctrl != 0);
}

extern "C" int gr_sums_device(gr_ctx* x, double** d_expt_sums, double** d_ctrl_sums) {
  if (!x) return GR_ERR_ARG;
  if (d_expt_sums) *d_expt_sums = x->dsums.as<double>();
  if (d_ctrl_sums) *d_ctrl_sums = x->dsums.as<double>() + x->nchrom;
  return GR_OK;
}

extern "C" void* gr_stream(gr_ctx* x) { return x ? (void*)x->stream : nullptr; }

extern "C" int gr_replicate_stats(gr_ctx* x, int32_t replicate, gr_sample_stats* st) {
  if (!x || !st || replicate < 0 || replicate >= (int)x->reps.size()) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  { int r = materialize(x); if (r) return r; }
  { int r = redo_pvals(x); if (r) return r; }
  const Replicate* rep = x->reps[replicate];
  float fl[2];
  CK(cudaMemcpy(fl, x->dpar.p, sizeof fl, cudaMemcpyDeviceToHost));   // the last replicate's scalars
  memset(st, 0, sizeof *st);
  st->n_ctrl = rep->n_ctrl;
  st->n_pval = rep->n;
  st->n_clamped = x->n_clamped;
  if (replicate == (int)x->reps.size() - 1) { st->factor = fl[0]; st->lambda = fl[1]; st->n_expt = x->n_expt; }
  return GR_OK;
}

extern "C" int gr_replicate_end(gr_ctx* x, gr_sample_stats* st) {
  if (!x) return GR_ERR_ARG;
  const bool pending_ctrl = x->filling == FILL_CTRL;
  if (x->filling != FILL_NONE) {
    int r = gr_sample_pileup(x, nullptr);
    if (r) return r;
  }
  const bool has_ctrl = pending_ctrl || x->have_ctrl;
  { int r = materialize(x); if (r) return r; }                 // the sums of both samples, one round trip
  double f = 0.0, g = 0.0;
  for (int c = 0; c < x->nchrom; c++) { f += x->expt_sums[c]; if (has_ctrl) g += x->ctrl_sums[c]; }
  return gr_replicate_finish(x, f, g, has_ctrl, 0, st);
}

// ---- Fisher combine ---------------------------------------------------------------
extern "C" int gr_pvalues_finalize(gr_ctx* x) {
  if (!x || x->reps.empty()) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  const int nrep = (int)x->reps.size();
  if (nrep > 200) return GR_ERR_DF;                            // pchisq 556
  if (x->comb) { x->pool.push_back(x->comb); x->comb = nullptr; }
  x->have_q = false;
  if (nrep == 1) {
    x->fin = x->reps[0];
    x->finalized = true;
    return GR_OK;
  }
  { int r = materialize(x); if (r) return r; }
  { int r = redo_pvals(x); if (r) return r; }
  const int nc = x->nchrom;
  Replicate* cb = new_replicate(x);
  x->comb = cb;
  CK(cb->bmU.ensure(x->T / 8));
  CK(cb->rankU.ensure(x->nblocks * sizeof(u64)));
  CK(cb->chrom_start.ensure((nc + 1) * sizeof(u64)));
  CK(cb->present.ensure(nc));
  const u64 nwords = x->T / 32;
  stage_begin(x, "fisher_union", (u64)nrep * x->T / 8);
  launch_or_bitmaps(x->stream, x->reps[0]->bmU.as<u32>(), x->reps[1]->bmU.as<u32>(), cb->bmU.as<u32>(), nwords);
  for (int r = 2; r < nrep; r++)
    launch_or_bitmaps(x->stream, cb->bmU.as<u32>(), x->reps[r]->bmU.as<u32>(), cb->bmU.as<u32>(), nwords);
  CK(x->lb0.ensure(3 * x->nblocks * sizeof(u64) > x->lb0.cap ? 3 * x->nblocks * sizeof(u64) : x->lb0.cap));
  CompactScratch cs;
  cs.st = x->lb0.as<u64>(); cs.ticket = x->ticket.as<u32>();
  launch_block_rank(x->stream, x->L, cb->bmU.as<u32>(), cs, cb->rankU.as<u64>(), x->d_totals);
  CKL();
  stage_end(x);
  CK(cudaMemcpyAsync(x->h_small, x->small.p, 64, cudaMemcpyDeviceToHost, x->stream));
  CK(cudaStreamSynchronize(x->stream));
  const u64 np = ((const u64*)((char*)x->h_small + 16))[2];
  cb->n = np; cb->n_upper = np; cb->n_ctrl = 0; cb->has_ctrl = false;
  CK(cb->cnt.ensure(16));
  CK(cudaMemcpyAsync(cb->cnt.p, x->d_totals + 2, 8, cudaMemcpyDeviceToDevice, x->stream));
  CK(cb->pEnd.ensure((np + 1) * sizeof(u32)));
  CK(cb->pVal.ensure((np + 1) * sizeof(float)));
  CK(x->fsum.ensure((np + 1) * sizeof(double)));
  CK(x->fdf.ensure((np + 1) * sizeof(int)));
  std::vector<RepView> views(nrep);
  for (int r = 0; r < nrep; r++) {
    views[r].bmU = x->reps[r]->bmU.as<u32>();
    views[r].rankU = x->reps[r]->rankU.as<u64>();
    views[r].pval = x->reps[r]->pVal.as<float>();
    views[r].present = x->reps[r]->present.as<uint8_t>();
  }
  stage_begin(x, "fisher_emit", np * 16 * nrep);
  launch_fisher_emit(x->stream, x->L, cb->bmU.as<u32>(), cb->rankU.as<u64>(), views.data(),
                     nrep, cb->pEnd.as<u32>(), x->fsum.as<double>(), x->fdf.as<int>(),
                     cb->chrom_start.as<u64>(), x->d_totals + 2);
  CKL();
  stage_end(x);
  // chi-square tails through a table of distinct sums (grown until it stays under half full)
  stage_begin(x, "fisher_eval", np * 16);
  CK(x->slot.ensure((np + 1) * sizeof(u32)));
  x->pair_valid = false;                       // the tables are reused
  for (;;) {
    int r = table_alloc(x, x->fisher_cap);
    if (r) return r;
    launch_fisher_table(x->stream, x->fsum.as<double>(), x->fdf.as<int>(), np, table_view(x, x->fisher_cap),
                        x->slot.as<u32>(), cb->pVal.as<float>(), x->d_err);
    CKL();
    x->lag = true;
    r = materialize(x);
    if (r) return r;
    if (!(x->retry_flags & GR_DE_TABLE)) break;
    x->retry_flags &= ~GR_DE_TABLE;
    if (x->fisher_cap >= (1u << 30)) { x->detail = "distinct-sum table overflow"; return GR_ERR_MEM; }
    x->fisher_cap <<= 2;
  }
  stage_end(x);
  cb->chrom_start_h.resize(nc + 1);
  CK(cudaMemcpyAsync(cb->chrom_start_h.data(), cb->chrom_start.p, (nc + 1) * sizeof(u64),
                     cudaMemcpyDeviceToHost, x->stream));
  CK(cudaStreamSynchronize(x->stream));      // also keeps `views` alive until the copy is done
  cb->present_h.resize(nc);
  for (int c = 0; c < nc; c++) cb->present_h[c] = cb->chrom_start_h[c + 1] > cb->chrom_start_h[c];
  x->fin = cb;
  x->finalized = true;
  return GR_OK;
}

// ---- BH ------------------------------------------------------------------------------
static u64 final_genome_len(const gr_ctx* x) {        // findPeaks 1091-1101
  if (x->par.genome_len) return x->par.genome_len;
  u64 G = 0;
  const Replicate* f = x->fin;
  for (int c = 0; c < x->nchrom; c++)
    if (!x->skip[c] && f->chrom_start_h[c + 1] > f->chrom_start_h[c]) G += x->len[c] - (x->has_bed ? x->bed_bp[c] : 0);
  return G;
}

extern "C" int gr_bh_local_hist(gr_ctx* x, const uint32_t** d_keys, const uint64_t** d_lens, uint64_t* n) {
  if (!x) return GR_ERR_ARG;
  if (!x->finalized) { int r = gr_pvalues_finalize(x); if (r) return r; }
  CK(cudaSetDevice(x->device));
  { int r = materialize(x); if (r) return r; }
  { int r = redo_pvals(x); if (r) return r; }
  Replicate* f = x->fin;
  const u64 np = f->n;
  u32 cap = 1u << 20;
  // One replicate, p-values computed here (not loaded, -P): the intervals already know the pair-table slot of
  // their (expt, ctrl) pair (k_pair_insert), so the bp are summed per slot and the list is read off that table.
  const bool by_pairs = x->reps.size() == 1 && f == x->reps[0] && x->pair_valid;
  x->hist_by_pairs = by_pairs;
  stage_begin(x, "bh_hist", np * 12);
  if (by_pairs) {
    cap = x->pair_cap;
    PairTable pt = table_view(x, cap);
    CK(cudaMemsetAsync(pt.lens, 0, (size_t)cap * 8, x->stream));
    launch_slot_hist(x->stream, f->pEnd.as<u32>(), x->slot.as<u32>(), f->n_upper, f->cnt.as<u64>(),
                     f->chrom_start.as<u64>(), x->nchrom, pt.lens);
    CKL();
  } else {
    CK(x->slot.ensure((f->n_upper + 1) * sizeof(u32)));
    int r = table_build(x, np, cap, [&](const PairTable& t) {
      launch_key_insert(x->stream, f->pEnd.as<u32>(), f->pVal.as<float>(), np, f->chrom_start.as<u64>(),
                        x->nchrom, t, x->slot.as<u32>(), x->d_err);
    });
    if (r) return r;
  }
  PairTable t = table_view(x, cap);
  // occupied count -> list
  CK(cudaMemcpyAsync(x->h_small, x->tCount.p, 4, cudaMemcpyDeviceToHost, x->stream));
  CK(cudaStreamSynchronize(x->stream));
  const u64 occ = *(u32*)x->h_small;
  CK(x->hk.ensure((occ + 1) * sizeof(u32)));
  CK(x->hl.ensure((occ + 1) * sizeof(u64)));
  CK(x->hcount.ensure(8));
  const u64 ntile = (cap + 255) / 256;
  CK(x->lb0.ensure(ntile * sizeof(u64) > x->lb0.cap ? ntile * sizeof(u64) : x->lb0.cap));
  CompactScratch cs;
  cs.st = x->lb0.as<u64>(); cs.ticket = x->ticket.as<u32>();
  launch_table_compact(x->stream, t, cs, x->hk.as<u32>(), x->hl.as<u64>(), x->hcount.as<u64>(), by_pairs ? 1 : 0);
  CKL();
  stage_end(x);
  // The caller reads hk / hl on a stream of its own (an NCCL all-gather, a copy): the list must be
  // complete when the pointers are handed out, not merely enqueued on x->stream.
  CK(cudaMemcpyAsync(x->h_small, x->hcount.p, 8, cudaMemcpyDeviceToHost, x->stream));
  CK(cudaStreamSynchronize(x->stream));
  const u64 listed = *(u64*)x->h_small;                  // <= occ: pairs that evaluate to SKIP are not listed
  if (!by_pairs) x->pair_valid = false;                  // the key table took the pair table's place
  x->hn = listed;
  // remember the table capacity for the q lookup
  x->n_distinct = 0;
  x->hist_cap = cap;
  if (d_keys) *d_keys = x->hk.as<u32>();
  if (d_lens) *d_lens = (const uint64_t*)x->hl.as<u64>();
  if (n) *n = listed;
  return GR_OK;
}

extern "C" int gr_bh_set_global(gr_ctx* x, const uint32_t* d_keys, const uint64_t* d_lens,
                                uint64_t n, uint64_t genome_len) {
  if (!x || !x->finalized || !genome_len) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  const u32 cap = x->hist_cap;
  if (!cap) return GR_ERR_ARG;
  Replicate* f = x->fin;
  const u64 m = n ? n : 1;
  CK(x->bk0.ensure(m * 4)); CK(x->bk1.ensure(m * 4));
  CK(x->bl0.ensure(m * 8)); CK(x->bl1.ensure(m * 8));
  const u64 nblk = (m + 4095) / 4096;
  CK(x->bhist.ensure(256 * nblk * 4));
  CK(x->bdk.ensure(m * 4)); CK(x->bdq.ensure(m * 4)); CK(x->bdl.ensure(m * 8));
  CK(x->bdcount.ensure(8));
  const u64 ntile = (m + 255) / 256;
  CK(x->lb1.ensure(ntile * sizeof(u64) > x->lb1.cap ? ntile * sizeof(u64) : x->lb1.cap));
  BhWork w;
  w.k0 = x->bk0.as<u32>(); w.k1 = x->bk1.as<u32>(); w.l0 = x->bl0.as<u64>(); w.l1 = x->bl1.as<u64>();
  w.hist = x->bhist.as<u32>(); w.ksum = nullptr; w.x = nullptr;
  w.dk = x->bdk.as<u32>(); w.dq = x->bdq.as<float>(); w.dl = x->bdl.as<u64>();
  w.dcount = x->bdcount.as<u64>();
  w.sc.st = x->lb1.as<u64>(); w.sc.ticket = x->ticket.as<u32>();
  w.cap = m;
  const float logN = -log10f((float)genome_len);               // saveQval 221
  stage_begin(x, "bh", n * 12 * 8);
  launch_bh(x->stream, d_keys, (const u64*)d_lens, n, logN, w);
  PairTable t = table_view(x, cap);
  launch_table_q(x->stream, t, w.dk, w.dq, w.dcount, x->hist_by_pairs ? 1 : 0);
  CK(x->qVal.ensure((f->n + 1) * sizeof(float)));
  launch_gather_f32(x->stream, t.qval, x->slot.as<u32>(), f->n, f->cnt.as<u64>(), x->qVal.as<float>());
  CKL();
  stage_end(x);
  CK(cudaMemcpyAsync(x->h_small, x->bdcount.p, 8, cudaMemcpyDeviceToHost, x->stream));
  CK(cudaStreamSynchronize(x->stream));
  x->n_distinct = *(u64*)x->h_small;
  x->all_q_one = 0;
  if (x->n_distinct) {
    float top;
    CK(cudaMemcpy(&top, x->bdq.as<float>() + (x->n_distinct - 1), 4, cudaMemcpyDeviceToHost));
    x->all_q_one = top == 0.0f;                                // Genrich.c:245
  }
  x->have_q = true;
  return GR_OK;
}

// The same exchange with the lists in HOST memory: for a caller that drives several contexts from
// one process without a device-side collective (the host program's --gpus: computeQval 352 sees
// one histogram, so every context gets the concatenation of all local lists).
extern "C" int gr_bh_local_hist_host(gr_ctx* x, const uint32_t** keys, const uint64_t** lens, uint64_t* n) {
  if (!x || !keys || !lens || !n) return GR_ERR_ARG;
  const uint32_t* dk; const uint64_t* dl; uint64_t hn = 0;
  { int r = gr_bh_local_hist(x, &dk, &dl, &hn); if (r) return r; }
  x->hk_h.resize(hn ? hn : 1);
  x->hl_h.resize(hn ? hn : 1);
  if (hn) {
    CK(cudaMemcpyAsync(x->hk_h.data(), dk, hn * sizeof(uint32_t), cudaMemcpyDeviceToHost, x->stream));
    CK(cudaMemcpyAsync(x->hl_h.data(), dl, hn * sizeof(uint64_t), cudaMemcpyDeviceToHost, x->stream));
  }
  CK(cudaStreamSynchronize(x->stream));
  *keys = x->hk_h.data(); *lens = x->hl_h.data(); *n = hn;
  return GR_OK;
}
extern "C" int gr_bh_set_global_host(gr_ctx* x, const uint32_t* keys, const uint64_t* lens, uint64_t n,
                                     uint64_t genome_len) {
  if (!x || (n && (!keys || !lens))) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  CK(x->ghk.ensure((n ? n : 1) * sizeof(uint32_t)));
  CK(x->ghl.ensure((n ? n : 1) * sizeof(uint64_t)));
  if (n) {
    CK(cudaMemcpyAsync(x->ghk.p, keys, n * sizeof(uint32_t), cudaMemcpyHostToDevice, x->stream));
    CK(cudaMemcpyAsync(x->ghl.p, lens, n * sizeof(uint64_t), cudaMemcpyHostToDevice, x->stream));
  }
  return gr_bh_set_global(x, x->ghk.as<uint32_t>(), x->ghl.as<uint64_t>(), n, genome_len);   // waits for the device itself
}

// -P: the final p (and q) arrays come from the caller (callPeaksLog 1277); K8 runs on them as on computed ones
extern "C" int gr_load_pvalues(gr_ctx* x, const uint64_t* chrom_start, const uint32_t* end, const float* pval,
                               const float* qval, uint64_t n) {
  if (!x || !chrom_start || (n && (!end || !pval))) return GR_ERR_ARG;
  if ((x->par.qval_opt != 0) != (qval != nullptr)) return GR_ERR_ARG;
  if (chrom_start[0] != 0 || chrom_start[x->nchrom] != n || n >= 0xfffffff0ull || !x->reps.empty()) return GR_ERR_ARG;
  for (int c = 0; c < x->nchrom; c++) if (chrom_start[c + 1] < chrom_start[c]) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  { int r = materialize(x); if (r) return r; }
  const int nc = x->nchrom;
  Replicate* rep = new_replicate(x);
  x->reps.push_back(rep);
  rep->has_ctrl = false;
  rep->has_cols = false;
  x->pair_valid = false;
  rep->n = rep->n_upper = n;
  rep->n_ctrl = 0;
  CK(rep->cnt.ensure(16));
  CK(rep->pEnd.ensure((n + 1) * sizeof(u32)));
  CK(rep->pVal.ensure((n + 1) * sizeof(float)));
  CK(rep->chrom_start.ensure((nc + 1) * sizeof(u64)));
  CK(rep->present.ensure(nc));
  const u64 cnt2[2] = { n, 0 };
  { int r = upload(x, rep->cnt.p, cnt2, 16); if (r) return r; }
  { int r = upload(x, rep->chrom_start.p, chrom_start, (nc + 1) * sizeof(u64)); if (r) return r; }
  if (n) {
    CK(cudaMemcpyAsync(rep->pEnd.p, end, n * sizeof(u32), cudaMemcpyHostToDevice, x->stream));
    CK(cudaMemcpyAsync(rep->pVal.p, pval, n * sizeof(float), cudaMemcpyHostToDevice, x->stream));
  }
  rep->chrom_start_h.assign(chrom_start, chrom_start + nc + 1);
  rep->present_h.resize(nc);
  for (int c = 0; c < nc; c++) rep->present_h[c] = chrom_start[c + 1] > chrom_start[c];
  { int r = upload(x, rep->present.p, rep->present_h.data(), nc); if (r) return r; }
  if (qval) {
    CK(x->qVal.ensure((n + 1) * sizeof(float)));
    if (n) CK(cudaMemcpyAsync(x->qVal.p, qval, n * sizeof(float), cudaMemcpyHostToDevice, x->stream));
    x->have_q = true;
    x->n_distinct = 0;
    x->all_q_one = 0;
  }
  CK(cudaStreamSynchronize(x->stream));          // the caller's arrays may go away when we return
  x->fin = rep;
  x->finalized = true;
  return GR_OK;
}

// ---- peaks ---------------------------------------------------------------------------
// events -> heads -> walk -> compaction, all sized by upper bounds with the counts on the device
// want_host: 0 the records stay on the device, 1 the first PEAK_SPEC come back with the counts, 2 nothing is copied
// to the host at all (the caller gathers the slot: gr_call_peaks_enqueue)
static int peaks_enqueue(gr_ctx* x, Replicate* f, int qopt, int want_host) {
  const u64 nu = f->n_upper;
  CK(x->evIdx.ensure((nu + 1) * sizeof(u32)));
  CK(x->headIdx.ensure((nu + 1) * sizeof(u32)));
  CK(x->evCount.ensure(8)); CK(x->headCount.ensure(8));
  if (!x->head_cap) x->head_cap = nu / 16 > (1u << 20) ? nu / 16 : (1u << 20);
  if (x->head_cap > nu + 1) x->head_cap = nu + 1;
  const u64 hc = x->head_cap;
  CK(x->cand.ensure(hc * sizeof(PeakRec)));
  CK(x->candOk.ensure(hc));
  CK(x->peakOut.ensure(64 + hc * sizeof(PeakRec)));             // gr_peak_slot header, then the records
  const u64 nt = (nu + 2047) / 2048 + (hc + 255) / 256 + 2;     // status words of the largest tiling
  CK(x->lb0.ensure(nt * sizeof(u64) > x->lb0.cap ? nt * sizeof(u64) : x->lb0.cap));
  PeakWork w;
  w.ev_idx = x->evIdx.as<u32>(); w.ev_count = x->evCount.as<u64>();
  w.head_idx = x->headIdx.as<u32>(); w.head_count = x->headCount.as<u64>();
  w.cand = x->cand.as<PeakRec>(); w.cand_ok = x->candOk.as<uint8_t>();
  w.out = (PeakRec*)((char*)x->peakOut.p + 64); w.out_count = x->peakOut.as<u64>(); w.peak_bp = x->peakOut.as<u64>() + 1;
  w.sc.st = x->lb0.as<u64>(); w.sc.ticket = x->ticket.as<u32>();
  const float* v = qopt ? x->qVal.as<float>() : f->pVal.as<float>();
  stage_begin(x, "peak_events", nu * 4);
  launch_peak_events(x->stream, v, nu, f->cnt.as<u64>(), x->par.min_pqval, w);
  CKL();
  stage_end(x);
  stage_begin(x, "peak_scan", nu);
  launch_peak_chain(x->stream, f->pEnd.as<u32>(), f->pVal.as<float>(), x->qVal.as<float>(),
                    f->chrom_start.as<u64>(), x->nchrom, x->par.min_pqval, qopt, x->par.max_gap,
                    x->par.min_auc, x->par.min_len, w, nu, hc, x->d_err);
  CKL();
  stage_end(x);
  // the counts and the first PEAK_SPEC records come back together
  if (want_host == 2) {
    // gr_call_peaks_enqueue: the rest of the slot header, filled in stream order -- the device error bits as they
    // stand behind the peak scan (then cleared: they are the caller's now) and the interval count
    CK(cudaMemcpyAsync((char*)x->peakOut.p + 16, x->d_err, 4, cudaMemcpyDeviceToDevice, x->stream));
    CK(cudaMemcpyAsync((char*)x->peakOut.p + 24, f->cnt.p, 8, cudaMemcpyDeviceToDevice, x->stream));
    CK(cudaMemsetAsync(x->d_err, 0, sizeof(int), x->stream));
    return GR_OK;
  }
  CK(cudaMemcpyAsync((char*)x->h_small + 192, x->peakOut.p, 16, cudaMemcpyDeviceToHost, x->stream));
  const u64 spec = hc < gr_ctx::PEAK_SPEC ? hc : gr_ctx::PEAK_SPEC;
  if (want_host)
    CK(cudaMemcpyAsync(x->h_peaks, (char*)x->peakOut.p + 64, spec * sizeof(gr_peak), cudaMemcpyDeviceToHost, x->stream));
  x->lag = true;
  return GR_OK;
}

extern "C" int gr_call_peaks(gr_ctx* x, const gr_peak** peaks, uint64_t* n, gr_run_stats* st) {
  if (!x || x->reps.empty()) return GR_ERR_ARG;
  if (!x->finalized) { int r = gr_pvalues_finalize(x); if (r) return r; }
  CK(cudaSetDevice(x->device));
  Replicate* f = x->fin;
  const int qopt = x->par.qval_opt != 0;
  if (qopt && !x->have_q) {
    const uint32_t* k; const uint64_t* l; uint64_t hn;
    int r = gr_bh_local_hist(x, &k, &l, &hn);
    if (r) return r;
    const u64 G = final_genome_len(x);
    if (!G) return GR_ERR_GENOME;
    r = gr_bh_set_global(x, k, l, hn, G);
    if (r) return r;
  }
  static_assert(sizeof(PeakRec) == sizeof(gr_peak), "peak record layout");
  static const bool gap_debug = getenv("GR_GAP_DEBUG") != nullptr;
  u64 npk = 0, peak_bp = 0;
  for (;;) {
    const auto t_a = std::chrono::steady_clock::now();
    { int r = peaks_enqueue(x, f, qopt, peaks != nullptr ? 1 : 0); if (r) return r; }
    const auto t_b = std::chrono::steady_clock::now();
    { int r = materialize(x); if (r) return r; }               // the one round trip of a peak call
    HT("call_peaks: device waited for");
    if (gap_debug) {
      const auto t_c = std::chrono::steady_clock::now();
      fprintf(stderr, "gr_call_peaks: enqueue %.3f ms, wait for the device %.3f ms\n",
              std::chrono::duration<double, std::milli>(t_b - t_a).count(),
              std::chrono::duration<double, std::milli>(t_c - t_b).count());
    }
    if (x->retry_flags & GR_DE_TABLE) {                        // -log10 p came from an overflowed table
      int r = redo_pvals(x);
      if (r) return r;
      if (qopt) { x->have_q = false; return gr_call_peaks(x, peaks, n, st); }
      continue;
    }
    if (x->retry_flags & GR_DE_CAP) {                          // more candidate peaks than room: once more, with room
      x->retry_flags &= ~GR_DE_CAP;
      x->head_cap = x->head_cap * 8 < f->n_upper + 1 ? x->head_cap * 8 : f->n_upper + 1;
      continue;
    }
    npk = *(u64*)((char*)x->h_small + 192);
    peak_bp = *(u64*)((char*)x->h_small + 200);
    break;
  }
  x->n_peaks = npk;
  if (peaks) {                                                 // peaks == NULL: the records stay on the device (gr_peaks_device)
    const u64 spec = npk < gr_ctx::PEAK_SPEC ? npk : gr_ctx::PEAK_SPEC;
    if (npk <= gr_ctx::PEAK_SPEC) {
      // the usual case: every record came back with the counts -- hand out the pinned buffer itself
      // (valid until the next peak call, like the reference's printPeak arguments it stands for)
      *peaks = x->h_peaks;
    } else {
      x->peaks_h.resize(npk);
      memcpy(x->peaks_h.data(), x->h_peaks, spec * sizeof(gr_peak));
      CK(cudaMemcpy(x->peaks_h.data() + spec, (const gr_peak*)((char*)x->peakOut.p + 64) + spec, (npk - spec) * sizeof(gr_peak),
                    cudaMemcpyDeviceToHost));
      *peaks = x->peaks_h.data();
    }
  }
  if (n) *n = npk;
  HT("call_peaks: peaks on the host");
  if (gap_debug) {
    static auto t_last = std::chrono::steady_clock::now();
    const auto t_d = std::chrono::steady_clock::now();
    fprintf(stderr, "gr_call_peaks: returns %.3f ms after the previous return\n",
            std::chrono::duration<double, std::milli>(t_d - t_last).count());
    t_last = t_d;
  }
  if (st) {
    memset(st, 0, sizeof *st);
    st->genome_len = final_genome_len(x);
    st->n_peaks = npk;
    st->peak_bp = peak_bp;
    st->n_intervals = f->n;
    st->n_distinct_p = qopt ? x->n_distinct : 0;
    st->all_q_one = qopt ? x->all_q_one : 0;
    st->n_replicates = (int)x->reps.size();
  }
  return GR_OK;
}

// gr_call_peaks in two halves for launchers that gather several contexts' peaks with ONE wait: the first half
// enqueues, the caller moves the slot (an all-gather, a copy) on gr_stream and waits once, the second half takes
// this context's header as the host now sees it.
extern "C" int gr_call_peaks_enqueue(gr_ctx* x, const void** d_slot, uint64_t* record_cap) {
  if (!x || x->reps.empty() || !d_slot) return GR_ERR_ARG;
  if (!x->finalized) { int r = gr_pvalues_finalize(x); if (r) return r; }
  CK(cudaSetDevice(x->device));
  const int qopt = x->par.qval_opt != 0;
  if (x->retry_flags & GR_DE_TABLE) {                          // left by gr_call_peaks_done: -log10 p came from an overflowed table
    { int r = materialize(x); if (r) return r; }
    { int r = redo_pvals(x); if (r) return r; }
    if (qopt) x->have_q = false;
  }
  if (qopt && !x->have_q) return GR_ERR_ARG;                   // the launcher runs the histogram exchange first
  { int r = peaks_enqueue(x, x->fin, qopt, 2); if (r) return r; }
  *d_slot = x->peakOut.p;
  if (record_cap) *record_cap = x->head_cap;
  return GR_OK;
}
extern "C" int gr_call_peaks_done(gr_ctx* x, const gr_peak_slot* hdr, int32_t* redo, gr_run_stats* st) {
  if (!x || !hdr || !redo) return GR_ERR_ARG;
  *redo = 0;
  const int derr = hdr->flags;
  const int hard = derr & ~(GR_DE_TABLE | GR_DE_CAP);
  if (hard) return map_dev_err(hard);
  if (derr & GR_DE_TABLE) { x->retry_flags |= GR_DE_TABLE; *redo = 2; return GR_OK; }   // p-values have to be redone: gr_call_peaks does that
  if (derr & GR_DE_CAP) {                                      // more candidate peaks than room: once more, with room
    const u64 lim = x->fin->n_upper + 1;
    x->head_cap = x->head_cap * 8 < lim ? x->head_cap * 8 : lim;
    *redo = 1;
    return GR_OK;
  }
  x->n_peaks = hdr->n_peaks;
  if (st) {
    memset(st, 0, sizeof *st);
    st->genome_len = x->par.genome_len;                        // the launcher computed it for the exchange (0: not given)
    st->n_peaks = hdr->n_peaks;
    st->peak_bp = hdr->peak_bp;
    st->n_intervals = hdr->n_intervals;
    st->n_distinct_p = x->par.qval_opt ? x->n_distinct : 0;
    st->all_q_one = x->par.qval_opt ? x->all_q_one : 0;
    st->n_replicates = (int)x->reps.size();
  }
  return GR_OK;
}

// ---- seam OUT for -f / -k --------------------------------------------------------------
extern "C" int gr_fetch_intervals(gr_ctx* x, int32_t which, int32_t replicate, int32_t chrom,
                                  const uint32_t** end, const float** val, const float** expt,
                                  const float** ctrl, uint64_t* n) {
  if (!x || chrom < 0 || chrom >= x->nchrom) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  { int r = materialize(x); if (r) return r; }
  { int r = redo_pvals(x); if (r) return r; }
  CK(cudaStreamSynchronize(x->stream));
  const u32* dE = nullptr; const float* dV = nullptr; const float* dX = nullptr; const float* dC = nullptr;
  u64 a = 0, b = 0;
  if (which == 0 || which == 1) {
    const std::vector<u64>& cs = which ? x->ctrlCS_h : x->exptCS_h;
    std::vector<u64> tmp;
    const std::vector<u64>* use = &cs;
    if (which == 1) {            // control start table lives on the device after the sweep
      tmp.resize(x->nchrom + 1);
      CK(cudaMemcpy(tmp.data(), x->ctrlCS.p, (x->nchrom + 1) * sizeof(u64), cudaMemcpyDeviceToHost));
      use = &tmp;
    }
    if (use->size() != (size_t)x->nchrom + 1) return GR_ERR_ARG;
    a = (*use)[chrom]; b = (*use)[chrom + 1];
    dE = which ? x->ctrlEnd.as<u32>() : x->exptEnd.as<u32>();
    dV = which ? x->ctrlVal.as<float>() : x->exptVal.as<float>();
  } else if (which == 2 || which == 3) {
    const int nrep = (int)x->reps.size();
    Replicate* r = nullptr;
    if (which == 3) {
      if (!x->have_q) return GR_ERR_ARG;
      r = x->fin;
    } else if (replicate == nrep) {
      if (!x->finalized) return GR_ERR_ARG;
      r = x->fin;
    } else if (replicate >= 0 && replicate < nrep)
      r = x->reps[replicate];
    else
      return GR_ERR_ARG;
    a = r->chrom_start_h[chrom]; b = r->chrom_start_h[chrom + 1];
    dE = r->pEnd.as<u32>();
    dV = which == 3 ? x->qVal.as<float>() : r->pVal.as<float>();
    if (which == 2 && r->has_cols) { dX = r->pExpt.as<float>(); dC = r->pCtrl.as<float>(); }
  } else
    return GR_ERR_ARG;
  const u64 m = b - a;
  if (end) *end = nullptr;
  if (val) *val = nullptr;
  if (expt) *expt = nullptr;
  if (ctrl) *ctrl = nullptr;
  if (n) *n = 0;
  if (!m) return GR_OK;
  x->f_end.resize(m); x->f_val.resize(m);
  CK(cudaMemcpy(x->f_end.data(), dE + a, m * sizeof(u32), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(x->f_val.data(), dV + a, m * sizeof(float), cudaMemcpyDeviceToHost));
  if (dX) {
    x->f_expt.resize(m); x->f_ctrl.resize(m);
    CK(cudaMemcpy(x->f_expt.data(), dX + a, m * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(x->f_ctrl.data(), dC + a, m * sizeof(float), cudaMemcpyDeviceToHost));
  }
  if (end) *end = x->f_end.data();
  if (val) *val = x->f_val.data();
  if (expt && dX) *expt = x->f_expt.data();
  if (ctrl && dX) *ctrl = x->f_ctrl.data();
  if (n) *n = m;
  return GR_OK;
}

// ---- timing / misc ------------------------------------------------------------------------
extern "C" int gr_timing_enable(gr_ctx* x, int32_t on) {
  if (!x) return GR_ERR_ARG;
  x->timing = on != 0;
  return GR_OK;
}
extern "C" int gr_timing_reset(gr_ctx* x) {
  if (!x) return GR_ERR_ARG;
  cudaSetDevice(x->device);
  cudaStreamSynchronize(x->stream);
  for (auto& s : x->stages) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
  x->stages.clear();
  return GR_OK;
}
extern "C" int gr_timing_get(gr_ctx* x, gr_stage_time* out, int32_t cap, int32_t* n) {
  if (!x || !out || !n) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  CK(cudaStreamSynchronize(x->stream));
  if (getenv("GR_GAP_DEBUG")) {            // time between the end of a stage and the start of the next one
    for (size_t i = 0; i + 1 < x->stages.size() && i < 40; i++) {
      float d = 0.f, g = 0.f;
      cudaEventElapsedTime(&d, x->stages[i].a, x->stages[i].b);
      cudaEventElapsedTime(&g, x->stages[i].b, x->stages[i + 1].a);
      fprintf(stderr, "stage %-12s %8.3f ms, then idle/untimed %8.3f ms\n", x->stages[i].name, d, g);
    }
    cudaGetLastError();
  }
  int m = 0;
  for (auto& s : x->stages) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) != cudaSuccess) { cudaGetLastError(); continue; }
    int k = 0;
    for (; k < m; k++) if (!strcmp(out[k].name, s.name)) break;
    if (k == m) {
      if (m == cap) continue;
      out[m].name = s.name; out[m].ms = 0; out[m].launches = 0; out[m].bytes = 0;
      m++;
    }
    out[k].ms += ms; out[k].launches += 1; out[k].bytes += s.bytes;
  }
  *n = m;
  return GR_OK;
}
extern "C" uint64_t gr_kernel_launches(const gr_ctx* x) { (void)x; return g_gr_launches; }
extern "C" int gr_scan_form(gr_ctx* x, int32_t* form, uint64_t* hot_entries, uint64_t* entries) {
  if (!x || !form) return GR_ERR_ARG;
  *form = -1;
  if (hot_entries) *hot_entries = 0;
  if (entries) *entries = 0;
  if (x->last_built != 2) return GR_OK;
  CK(cudaSetDevice(x->device));
  u32 stat[2] = {0, 0}, total = 0;
  CK(cudaMemcpyAsync(stat, x->sbCnt.as<u32>() + x->nblocks, 8, cudaMemcpyDeviceToHost, x->stream));
  CK(cudaMemcpyAsync(&total, x->sbStart.as<u32>() + x->nblocks, 4, cudaMemcpyDeviceToHost, x->stream));
  CK(cudaStreamSynchronize(x->stream));
  if (hot_entries) *hot_entries = stat[1];
  if (entries) *entries = total;
  const bool forced_cta = x->has_bed || (getenv("GR_FUSED_CTA") && atoi(getenv("GR_FUSED_CTA")));
  const bool forced_rank = !forced_cta && getenv("GR_FUSED_RANK") && atoi(getenv("GR_FUSED_RANK"));
  *form = forced_cta ? 1 : forced_rank ? 0 : ((u64)stat[1] * 4 > (u64)total ? 1 : 0);
  return GR_OK;
}
extern "C" int gr_synchronize(gr_ctx* x) {
  if (!x) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  CK(cudaStreamSynchronize(x->copy));
  CK(cudaStreamSynchronize(x->stream));
  return GR_OK;
}

extern "C" int gr_timer_start(gr_ctx* x) {
  if (!x) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  if (!x->tm_a) { CK(cudaEventCreate(&x->tm_a)); CK(cudaEventCreate(&x->tm_b)); }
  CK(cudaEventRecord(x->tm_a, x->stream));
  return GR_OK;
}
extern "C" int gr_timer_stop(gr_ctx* x, double* ms) {
  if (!x || !ms || !x->tm_a) return GR_ERR_ARG;
  CK(cudaSetDevice(x->device));
  CK(cudaEventRecord(x->tm_b, x->stream));
  CK(cudaEventSynchronize(x->tm_b));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, x->tm_a, x->tm_b));
  *ms = f;
  return GR_OK;
}

extern "C" void* gr_pinned_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
extern "C" void gr_pinned_free(void* p) { if (p) cudaFreeHost(p); }

// Host utility for multi-context callers: the peak lists of the contexts (each in chromosome
// order, every chromosome in exactly one list) -> one list in chromosome order (callPeaks
// numbers peaks in that order, Genrich.c:986).  Runs are moved with memcpy.
extern "C" int gr_merge_peaks(const gr_peak* const* lists, const uint64_t* counts, int32_t nlists, gr_peak* out) {
  if (nlists < 0 || (nlists && (!lists || !counts)) || !out) return GR_ERR_ARG;
  std::vector<u64> pos(nlists, 0);
  u64 w = 0;
  for (;;) {
    int best = -1;
    for (int i = 0; i < nlists; i++)
      if (pos[i] < counts[i] && (best < 0 || lists[i][pos[i]].chrom < lists[best][pos[best]].chrom)) best = i;
    if (best < 0) break;
    const gr_peak* l = lists[best];
    const int32_t c = l[pos[best]].chrom;
    // end of the chromosome's run: gallop, then bisect (the lists lie in pinned host memory: every record looked
    // at is a cache line fetched, and the copy below fetches them all once more)
    u64 lo = pos[best], step = 1, hi = counts[best];
    while (lo + step < counts[best] && l[lo + step].chrom == c) { lo += step; step <<= 1; }
    if (lo + step < hi) hi = lo + step;                  // l[lo].chrom == c, l[hi] (if any) is past the run
    while (lo + 1 < hi) {
      const u64 mid = lo + (hi - lo) / 2;
      if (l[mid].chrom == c) lo = mid; else hi = mid;
    }
    const u64 e = lo + 1;
    memcpy(out + w, l + pos[best], (e - pos[best]) * sizeof(gr_peak));
    w += e - pos[best];
    pos[best] = e;
  }
  return GR_OK;
}

extern "C" int gr_peaks_device(gr_ctx* x, const gr_peak** d_peaks, uint64_t* n) {
  if (!x || !d_peaks || !n) return GR_ERR_ARG;
  *d_peaks = x->n_peaks ? (const gr_peak*)((char*)x->peakOut.p + 64) : nullptr;
  *n = x->n_peaks;
  return GR_OK;
}
