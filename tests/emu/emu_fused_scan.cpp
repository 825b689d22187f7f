// emu_fused_scan.cpp -- TEST INFRASTRUCTURE (CPU): the fused-scan kernels of gr_dense.cu, compiled
// for the host through cuda_emu.h, on seeded interval records:
//   k_fb_count -> k_sb_scan1..3 -> k_fb_move -> k_fb_scan  (the path validated on the B200) -> k_scan_fix -> k_scan_place
//   the same buckets -> k_fr_scan<CAP> (rank form) -> k_scan_fix -> k_scan_place
//   k_fb_move_slot -> k_fr_scan<CAP, ., SLOT> (fixed-capacity buckets, no count pass), with roomy slots
//   and with slots so small that the exact chain behind the gate has to take over
//   k_p1_count -> k_p1_scan -> k_p1_move -> k_p2 (two-level partition): same buckets as count -> scan -> move
// and a plain per-cell prefix sum written here.  All of them must give the same RLE pileup
// (interval ends, float bits, chromosome starts) and the same break bitmap.
#include "cuda_emu.h"
#include "gen_kernels.h"
#include <random>

template <typename T> static T* dalloc(size_t n) { return (T*)aligned_alloc(64, ((n * sizeof(T) + 63) / 64 + 1) * 64); }

struct Layout {
  std::vector<u32> len;
  std::vector<u64> off;
  std::vector<uint8_t> flags;
  std::vector<int> blk2chrom;
  u64 T = 0;
  DevLayout dev() const {
    DevLayout L;
    L.nchrom = (int)len.size(); L.T = T; L.nblocks = T / GR_BLOCK_SLOTS;
    L.off = off.data(); L.len = len.data(); L.flags = flags.data(); L.blk2chrom = blk2chrom.data();
    return L;
  }
};
static Layout make_layout(const std::vector<u32>& len, const std::vector<uint8_t>& flags) {
  Layout L;
  L.len = len; L.flags = flags;
  for (size_t c = 0; c < len.size(); c++) {
    if (!(flags[c] & GR_CF_OWNED)) { L.off.push_back(~0ull); continue; }
    L.off.push_back(L.T);
    const u64 nb = ((u64)len[c] + 1 + GR_BLOCK_SLOTS - 1) / GR_BLOCK_SLOTS;
    for (u64 b = 0; b < nb; b++) L.blk2chrom.push_back((int)c);
    L.T += nb * GR_BLOCK_SLOTS;
  }
  return L;
}

struct Result {
  std::vector<u32> end; std::vector<u32> valbits; std::vector<u64> cs; std::vector<u32> bitmap;
  u64 total = 0; int err = 0;
  bool operator==(const Result& o) const {
    return end == o.end && valbits == o.valbits && cs == o.cs && bitmap == o.bitmap && total == o.total && err == o.err;
  }
};

struct Ws {
  StreamWs W; void* mem; u64 cap;
  Ws(u64 cap_, int nchrom) : cap(cap_) {
    const u64 max_pages = (cap / SS_PAGE + SS_SPARE_PAGES + 3) & ~1ull;
    const size_t bytes = (size_t)(max_pages * SS_PAGE * 8 + max_pages * 8 + SS_MAX_WARPS * (8 + 16) + (u64)nchrom * 16 + 256);
    mem = aligned_alloc(64, (bytes + 63) / 64 * 64);
    // stale garbage, like a reused device buffer (the page area is 268 MB: only its first 16 MB, which is
    // more than any case here fills, and everything behind it)
    const size_t pent_bytes = (size_t)max_pages * SS_PAGE * 8;
    memset(mem, 0xA5, std::min<size_t>(pent_bytes, 16u << 20));
    memset((char*)mem + pent_bytes, 0xA5, bytes - pent_bytes);
    char* p = (char*)mem;
    W.pent = (uint2*)p; p += max_pages * SS_PAGE * 8;
    W.page_meta = (uint2*)p; p += max_pages * 8;
    W.warp_base = (ulonglong2*)p; p += SS_MAX_WARPS * 16;
    W.warp_tot = (uint2*)p; p += SS_MAX_WARPS * 8;
    W.marks = (uint4*)p; p += (u64)nchrom * 16;
    W.page_ctr = (u32*)p;
    W.max_pages = (u32)max_pages;
    *W.page_ctr = 0;
  }
  ~Ws() { free(mem); }
};

// K2b + K2c, then collect
static Result finish(const Layout& Lh, Ws& ws, u32 owners, const u32* bitmap, int* err) {
  const DevLayout L = Lh.dev();
  Result R;
  const u64 cap = ws.cap;
  u32* end = dalloc<u32>(cap + 1); float* val = dalloc<float>(cap + 1);
  u64* cs = dalloc<u64>(L.nchrom + 2); u64* tot = dalloc<u64>(1);
  memset(end, 0xEE, (cap + 1) * 4); memset(val, 0xEE, (cap + 1) * 4);
  DevRle out{end, val, cs, tot};
  StreamWs W = ws.W;
  emu::launch(1, 1024, [&] { k_scan_fix(L, W, out, err, owners); });
  emu::launch(8, 2 * SS_PAGE, [&] { k_scan_place(W, out, err, 0.0f); });
  R.total = *tot;
  R.end.assign(end, end + R.total);
  R.valbits.resize(R.total);
  memcpy(R.valbits.data(), val, R.total * 4);
  R.cs.assign(cs, cs + L.nchrom + 1);
  R.bitmap.assign(bitmap, bitmap + L.T / 32);
  R.err = *err;
  free(end); free(val); free(cs); free(tot);
  return R;
}

// the reference semantics, cell by cell (savePileupExpt 2239-2273 without -E)
static Result plain(const Layout& Lh, const std::vector<int4>& recs) {
  Result R;
  R.bitmap.assign(Lh.T / 32, 0);
  R.cs.assign(Lh.len.size() + 1, 0);
  float4 lut[120];
  units_lut_fill(lut, 0, 1);
  for (size_t c = 0; c < Lh.len.size(); c++) {
    R.cs[c] = R.end.size();
    if (!(Lh.flags[c] & GR_CF_OWNED)) continue;
    const u32 len = Lh.len[c];
    std::vector<int> d(len + 1, 0);
    const bool act = (Lh.flags[c] & GR_CF_SAVE) != 0;
    for (const int4& r : recs) {
      if (r.x != (int)c || !act) continue;
      long s = r.y, e = r.z;
      if (s < 0) s = 0;
      if (s >= (long)len || e < s) continue;
      if (e > (long)len) e = len;
      d[s] += 120 / r.w; d[e] -= 120 / r.w;
    }
    if (!act) continue;
    int h = 0;
    for (u32 j = 0; j <= len; j++) {
      const bool brk = j == len || (j >= 1 && d[j] != 0);
      if (brk) {
        R.end.push_back(j);
        R.valbits.push_back(__float_as_uint(units_to_val_lut(lut, h)));
        const u64 g = Lh.off[c] + j;
        R.bitmap[g >> 5] |= 1u << (g & 31);
      }
      h += d[j];
    }
  }
  R.cs[Lh.len.size()] = R.end.size();
  // chromosomes without slots start where the next one does (k_scan_fix)
  for (int c = (int)Lh.len.size() - 1; c >= 0; c--)
    if (!(Lh.flags[c] & GR_CF_OWNED)) R.cs[c] = R.cs[c + 1];
  R.total = R.end.size();
  return R;
}

static void diff(const char* what, const Result& a, const Result& b) {
  fprintf(stderr, "MISMATCH %s: total %llu vs %llu, err %d vs %d\n", what, a.total, b.total, a.err, b.err);
  for (size_t i = 0; i < std::min(a.end.size(), b.end.size()); i++)
    if (a.end[i] != b.end[i] || a.valbits[i] != b.valbits[i]) {
      fprintf(stderr, "  first difference at interval %zu: end %u vs %u, val %08x vs %08x\n", i, a.end[i], b.end[i],
              a.valbits[i], b.valbits[i]);
      break;
    }
  for (size_t i = 0; i < a.cs.size(); i++)
    if (a.cs[i] != b.cs[i]) { fprintf(stderr, "  chrom_start[%zu] %llu vs %llu\n", i, a.cs[i], b.cs[i]); break; }
  for (size_t i = 0; i < a.bitmap.size(); i++)
    if (a.bitmap[i] != b.bitmap[i]) { fprintf(stderr, "  bitmap word %zu: %08x vs %08x\n", i, a.bitmap[i], b.bitmap[i]); break; }
}

template <int CAP>
static Result run_rank(const Layout& Lh, const u32* bucketed, const u32* blk_start, u64 cap, u32 ctas, int err0) {
  const DevLayout L = Lh.dev();
  Ws ws(cap, L.nchrom);
  StreamWs W = ws.W;
  u32* bitmap = dalloc<u32>(L.T / 32);
  memset(bitmap, 0x5A, L.T / 8);
  int err = err0;                                     // what the bucket passes flagged
  const u32 owners = ctas * 4, nb = (u32)L.nblocks, R = (nb + owners - 1) / owners;
  emu::launch(ctas, 128, [&] { k_fr_scan<CAP, 1, (CAP == 512 ? 4 : 8), false>(bucketed, blk_start, L, W, bitmap, &err, nb, R, 0u, nullptr, 0); });
  Result r = finish(Lh, ws, owners, bitmap, &err);
  free(bitmap);
  return r;
}

// The slot path as gr_api.cu enqueues it: one pass into fixed-capacity buckets, the exact chain
// behind it gated on the overflow flag, then both scans (only one of them does anything).
template <int CAP>
static Result run_slots(const Layout& Lh, const int4* recs, u64 n, u32 slot_cap, u64 cap, u32 ctas, bool* overflowed, u64* n_clamped) {
  const DevLayout L = Lh.dev();
  const u64 nbk = L.nblocks;
  u32* scnt = dalloc<u32>(nbk + 1); u32* cnt = dalloc<u32>(nbk + 1); u32* start = dalloc<u32>(nbk + 2);
  u32* cursor = dalloc<u32>(nbk + 1); u32* chunk = dalloc<u32>(nbk / SB_CHUNK + 4);
  u32* bucketed = dalloc<u32>(std::max<u64>(nbk * slot_cap, 2 * n + 16));
  memset(bucketed, 0xC3, std::max<u64>(nbk * slot_cap, 2 * n + 16) * 4);
  memset(scnt, 0, (nbk + 1) * 4);
  memset(cnt, 0, (nbk + 1) * 4);
  int err = 0, gate = 0; u64 clamped = 0;
  emu::launch(4, 256, [&] { k_fb_move_slot<false>(recs, n, L, scnt, bucketed, slot_cap, &gate, &err, &clamped); });
  const int* g = &gate;
  emu::launch(4, 256, [&] { k_fb_count<false>(recs, n, L, cnt, &err, nullptr, GR_BLOCK_SHIFT, g); });
  const u32 nchunks = (u32)((nbk + SB_CHUNK - 1) / SB_CHUNK);
  emu::launch(nchunks, 256, [&] { k_sb_scan1(cnt, chunk, nbk); });
  emu::launch(1, 1024, [&] { k_sb_scan2(chunk, nchunks, start, nbk); });
  emu::launch(nchunks, 256, [&] { k_sb_scan3(cnt, chunk, start, cursor, nbk); });
  emu::launch(4, 256, [&] { k_fb_move<false>(recs, n, L, cursor, bucketed, GR_BLOCK_SHIFT, g); });
  Ws ws(cap, L.nchrom);
  StreamWs W = ws.W;
  u32* bitmap = dalloc<u32>(L.T / 32);
  memset(bitmap, 0x5A, L.T / 8);
  const u32 owners = ctas * 4, nb = (u32)nbk, R = (nb + owners - 1) / owners;
  emu::launch(ctas, 128, [&] { k_fr_scan<CAP, 1, (CAP == 512 ? 4 : 8), true>(bucketed, scnt, L, W, bitmap, &err, nb, R, slot_cap, g, 0); });
  emu::launch(ctas, 128, [&] { k_fr_scan<CAP, 1, (CAP == 512 ? 4 : 8), false>(bucketed, start, L, W, bitmap, &err, nb, R, 0u, g, 1); });
  Result r = finish(Lh, ws, owners, bitmap, &err);
  *overflowed = gate != 0;
  *n_clamped = clamped;
  free(scnt); free(cnt); free(start); free(cursor); free(chunk); free(bucketed); free(bitmap);
  return r;
}

static int run_case(const char* name, const std::vector<u32>& len, const std::vector<uint8_t>& flags,
                    const std::vector<int4>& recs_in, u32 ctas) {
  const Layout Lh = make_layout(len, flags);
  const DevLayout L = Lh.dev();
  const u64 n = recs_in.size();
  int4* recs = dalloc<int4>(n + 1);
  memcpy(recs, recs_in.data(), n * sizeof(int4));
  const u64 nbk = L.nblocks;
  u32* cnt = dalloc<u32>(nbk + 1); u32* start = dalloc<u32>(nbk + 2); u32* cursor = dalloc<u32>(nbk + 1);
  u32* chunk = dalloc<u32>(nbk / SB_CHUNK + 4);
  u32* bucketed = dalloc<u32>(2 * n + 16);
  memset(cnt, 0, (nbk + 1) * 4);
  int err = 0; u64 clamped = 0;
  // ---- the validated chain
  emu::launch(4, 256, [&] { k_fb_count<false>(recs, n, L, cnt, &err, &clamped, GR_BLOCK_SHIFT); });
  const u32 nchunks = (u32)((nbk + SB_CHUNK - 1) / SB_CHUNK);
  emu::launch(nchunks, 256, [&] { k_sb_scan1(cnt, chunk, nbk); });
  emu::launch(1, 1024, [&] { k_sb_scan2(chunk, nchunks, start, nbk); });
  emu::launch(nchunks, 256, [&] { k_sb_scan3(cnt, chunk, start, cursor, nbk); });
  emu::launch(4, 256, [&] { k_fb_move<false>(recs, n, L, cursor, bucketed, GR_BLOCK_SHIFT); });
  const u64 cap = 2 * n + len.size() + 1;
  Result base;
  {
    Ws ws(cap, L.nchrom);
    StreamWs W = ws.W;
    u32* bitmap = dalloc<u32>(L.T / 32);
    memset(bitmap, 0x5A, L.T / 8);
    int e2 = err;
    const u32 owners = ctas * 2, nb = (u32)nbk, R = (nb + owners - 1) / owners;
    emu::launch(owners, 128, [&] { k_fb_scan<6, 128, false>(bucketed, start, L, W, bitmap, &e2, nb, R, nullptr); });
    base = finish(Lh, ws, owners, bitmap, &e2);
    free(bitmap);
  }
  int bad = 0;
  const Result ref = plain(Lh, recs_in);
  Result refe = ref; refe.err = base.err;            // the plain walk does not model error flags
  if (!(base == refe)) { diff("k_fb_scan vs per-cell walk", base, refe); bad++; }
  // ---- rank form on the same buckets, three capacities (64: many rounds per block)
  const Result r64 = run_rank<64>(Lh, bucketed, start, cap, ctas, err);
  const Result r512 = run_rank<512>(Lh, bucketed, start, cap, ctas, err);
  const Result r1024 = run_rank<1024>(Lh, bucketed, start, cap, ctas + 1, err);
  if (!(r64 == base)) { diff("k_fr_scan<64> vs k_fb_scan", r64, base); bad++; }
  if (!(r512 == base)) { diff("k_fr_scan<512> vs k_fb_scan", r512, base); bad++; }
  if (!(r1024 == base)) { diff("k_fr_scan<1024> vs k_fb_scan", r1024, base); bad++; }
  // ---- slot path: roomy slots (no overflow), and slots so small that the gated exact chain takes over
  {
    u64 mx = 0;
    for (u64 b = 0; b < nbk; b++) mx = std::max<u64>(mx, cnt[b]);
    u32 roomy = 4; while (roomy < mx) roomy <<= 1;
    bool ov = false;
    u64 cl = 0;
    const Result s1 = run_slots<512>(Lh, recs, n, roomy, cap, ctas, &ov, &cl);
    if (!(s1 == base) || ov || cl != clamped) { diff("slot path (roomy) vs k_fb_scan", s1, base); bad++; }
    const Result s2 = run_slots<64>(Lh, recs, n, roomy, cap, ctas + 2, &ov, &cl);
    if (!(s2 == base) || ov || cl != clamped) { diff("slot path (roomy, 64 per round) vs k_fb_scan", s2, base); bad++; }
    if (mx > 4) {
      const Result s3 = run_slots<512>(Lh, recs, n, 4, cap, ctas, &ov, &cl);
      if (!(s3 == base) || !ov || cl != clamped) { diff("slot path (overflow -> exact chain) vs k_fb_scan", s3, base); bad++; }
    }
  }
  // ---- two-level partition (GR_FB_P2=1): same blk_start, same entries per bucket (order inside a bucket is free)
  {
    const int fsh = 9;
    const u32 nb1 = (u32)((nbk + (1ull << fsh) - 1) >> fsh);
    u32* cnt1 = dalloc<u32>(1024); u32* base1 = dalloc<u32>(1025); u32* cur1 = dalloc<u32>(1024);
    u64* pairs = dalloc<u64>(2 * n + 16);
    u32* start2 = dalloc<u32>(nbk + 2); u32* bucket2 = dalloc<u32>(2 * n + 16);
    memset(cnt1, 0, 1024 * 4); memset(start2, 0xEE, (nbk + 2) * 4); memset(bucket2, 0xEE, (2 * n + 16) * 4);
    int err2 = 0; u64 cl2 = 0;
    emu::launch(3, 256, [&] { k_p1_count<false>(recs, n, L, cnt1, fsh, &err2, &cl2); });
    emu::launch(1, 1024, [&] { k_p1_scan(cnt1, nb1, base1, cur1); });
    emu::launch(2, 256, [&] { k_p1_move<false>(recs, n, L, cur1, pairs, fsh, nb1); });
    emu::launch(nb1, 512, [&] { k_p2(pairs, base1, nb1, fsh, (u32)nbk, start2, bucket2); });
    bool ok = err2 == err && cl2 == clamped;
    for (u64 b = 0; b <= nbk; b++) ok = ok && start2[b] == start[b];
    for (u64 b = 0; ok && b < nbk; b++) {
      std::vector<u32> x(bucketed + start[b], bucketed + start[b + 1]), y(bucket2 + start2[b], bucket2 + start2[b + 1]);
      std::sort(x.begin(), x.end()); std::sort(y.begin(), y.end());
      ok = x == y;
    }
    if (!ok) { fprintf(stderr, "MISMATCH two-level partition vs count/scan/move (err %d vs %d, clamped %llu vs %llu)\n", err2, err, cl2, clamped); bad++; }
    free(cnt1); free(base1); free(cur1); free(pairs); free(start2); free(bucket2);
  }
  printf("%-28s %8llu records %7llu blocks %9llu intervals  err %d  %s\n", name, (unsigned long long)n,
         (unsigned long long)nbk, (unsigned long long)base.total, base.err, bad ? "FAIL" : "ok");
  free(recs); free(cnt); free(start); free(cursor); free(chunk); free(bucketed);
  return bad;
}

int main() {
  int bad = 0;
  const uint8_t A = GR_CF_OWNED | GR_CF_SAVE;
  {  // edge inputs: chromosome ends on block boundaries, one-base chromosome, empty interval, equal records
    std::vector<u32> len = {5000, 8192, 8191, 1, 20000, 16384, 16383, 30000, 40000};   // 7: its end cell is alone in a block; 8: no record at all
    std::vector<int4> r = {
        {0, 0, 5000, 1}, {0, -50, 10, 2}, {0, 4990, 6000, 3}, {1, 0, 1, 1}, {1, 8191, 8192, 1},
        {2, 8190, 8191, 10}, {2, 0, 8191, 8}, {3, 0, 1, 1}, {4, 100, 100, 5},
        {4, 300, 900, 6}, {4, 300, 900, 6}, {4, 300, 900, 6}, {4, 300, 900, 6}, {4, 300, 900, 6}, {4, 300, 900, 6},
        {4, 19999, 25000, 4}, {5, 0, 16384, 1}, {5, 8191, 8192, 2}, {5, 8192, 8193, 2}, {5, 16383, 16384, 3},
        {6, 0, 16383, 1}, {6, 8100, 16383, 2}, {6, 16382, 16383, 3},
        {4, 1000, 1500, 2}, {4, 1500, 2000, 2},          // an end and a start of equal weight on one cell: no break
        {4, 3000, 3100, 3}, {4, 3100, 3200, 5}, {7, 100, 200, 1}};
    bad += run_case("edge", len, std::vector<uint8_t>(len.size(), A), r, 2);
  }
  std::mt19937_64 rng(20261017);
  auto gen = [&](const std::vector<u32>& len, const std::vector<uint8_t>& fl, size_t n, double hot, bool weights) {
    std::vector<int4> r;
    static const int CNT[8] = {1, 2, 3, 4, 5, 6, 8, 10};
    for (size_t i = 0; i < n; i++) {
      const int c = (int)(rng() % len.size());
      (void)fl;
      long s, e;
      const u32 L = len[c];
      if ((rng() % 1000) < hot * 1000) {                 // hot spots: hundreds of events inside a few hundred cells
        const long centre = (long)((rng() % 7 + 1) * (u64)L / 8);
        s = centre + (long)(rng() % 600) - 300;
        e = s + 50 + (long)(rng() % 300);
      } else {
        s = (long)(rng() % L);
        e = s + 100 + (long)(rng() % 300);
        if (rng() % 50 == 0) e = s + (long)(rng() % 30000);        // spans several blocks
        if (rng() % 97 == 0) s -= 200;                             // clamped at 0 now and then
      }
      if (s >= (long)L) s = L - 1;
      r.push_back(int4{c, (int)s, (int)e, weights ? CNT[rng() % 8] : 1});
    }
    return r;
  };
  {
    std::vector<u32> len = {300000, 70000, 8192 * 3 - 1, 123457};
    std::vector<uint8_t> fl(len.size(), A);
    bad += run_case("uniform, weight 1", len, fl, gen(len, fl, 20000, 0.0, false), 3);
    bad += run_case("hot spots, weights", len, fl, gen(len, fl, 30000, 0.5, true), 3);
    bad += run_case("dense (multi-round at 512)", {40000, 9000}, {A, A}, gen({40000, 9000}, {A, A}, 60000, 0.3, true), 2);
    fl[1] = GR_CF_OWNED;                                 // in the header, not in this replicate (Chrom.save false)
    fl[2] = 0;                                           // not owned by this context
    bad += run_case("unsaved / foreign chromosome", len, fl, gen(len, fl, 20000, 0.2, true), 5);
  }
  {  // more than two coarse bins of the two-level partition (512 blocks each), several tiles of 16384 records
    std::vector<u32> len = {8192u * 700 - 5, 8192u * 610 + 77, 50000};
    std::vector<uint8_t> fl(len.size(), A);
    bad += run_case("1317 blocks, 3 coarse bins", len, fl, gen(len, fl, 40000, 0.1, true), 4);
  }
  printf("fiber switches: %llu\n", emu::n_switches);
  return bad ? 1 : 0;
}
