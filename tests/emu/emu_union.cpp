// emu_union.cpp -- TEST INFRASTRUCTURE (CPU): k_union_emit_w (warp per bitmap block, GR_UE_WARP=1)
// against k_union_emit<4> and <2> (the default, validated on the B200) and a plain walk over the
// bits, on random break bitmaps: sparse blocks, empty blocks, blocks with more union breaks than
// one list round holds, a partial last CTA.
#include "cuda_emu.h"
#include "gen_kernels_union.h"
#include <random>

template <typename T> static T* dalloc(size_t n) { return (T*)aligned_alloc(64, ((n * sizeof(T) + 63) / 64 + 1) * 64); }

struct Out {
  std::vector<u32> pEnd, bmU; std::vector<float> pExpt, pCtrl; std::vector<u64> cs;
  bool operator==(const Out& o) const {
    return pEnd == o.pEnd && bmU == o.bmU && cs == o.cs &&
           !memcmp(pExpt.data(), o.pExpt.data(), pExpt.size() * 4) && !memcmp(pCtrl.data(), o.pCtrl.data(), pCtrl.size() * 4);
  }
};

int main() {
  std::mt19937_64 rng(777);
  int bad = 0;
  for (int trial = 0; trial < 4; trial++) {
    // chromosomes of 3, 1, 7, 2 (+trial) blocks
    std::vector<u32> nblk = {3, 1, 7, (u32)(2 + trial), (u32)(trial * 37)};
    if (!trial) nblk.pop_back();
    std::vector<u64> off; std::vector<u32> len; std::vector<uint8_t> flags; std::vector<int> b2c;
    u64 T = 0;
    for (size_t c = 0; c < nblk.size(); c++) {
      off.push_back(T); len.push_back(nblk[c] * GR_BLOCK_SLOTS - 1 - (u32)(rng() % 100)); flags.push_back(3);
      for (u32 b = 0; b < nblk[c]; b++) b2c.push_back((int)c);
      T += (u64)nblk[c] * GR_BLOCK_SLOTS;
    }
    DevLayout L; L.nchrom = (int)nblk.size(); L.T = T; L.nblocks = T / GR_BLOCK_SLOTS;
    L.off = off.data(); L.len = len.data(); L.flags = flags.data(); L.blk2chrom = b2c.data();
    const u32 nb = (u32)L.nblocks;
    u32* bmE = dalloc<u32>(T / 32); u32* bmC = dalloc<u32>(T / 32);
    for (u32 b = 0; b < nb; b++) {
      const int kind = (int)(rng() % 5);              // 0 empty, 1-2 sparse, 3 medium, 4 dense (several list rounds)
      const double pe = kind == 0 ? 0 : kind <= 2 ? 0.02 : kind == 3 ? 0.08 : 0.5;
      const double pc = kind == 0 ? 0 : kind <= 2 ? 0.01 : kind == 3 ? 0.05 : 0.4;
      for (u32 w = 0; w < 256; w++) {
        u32 e = 0, c = 0;
        for (int i = 0; i < 32; i++) {
          if ((rng() % 10000) < pe * 10000) e |= 1u << i;
          if ((rng() % 10000) < pc * 10000) c |= 1u << i;
        }
        bmE[(u64)b * 256 + w] = e; bmC[(u64)b * 256 + w] = c;
      }
    }
    std::vector<u64> rE(nb + 1), rC(nb + 1), rU(nb + 1);
    u64 e = 0, c = 0, u = 0;
    for (u32 b = 0; b < nb; b++) {
      rE[b] = e; rC[b] = c; rU[b] = u;
      for (u32 w = 0; w < 256; w++) {
        const u32 E = bmE[(u64)b * 256 + w], C = bmC[(u64)b * 256 + w];
        e += __popc(E); c += __popc(C); u += __popc(E | C);
      }
    }
    rE[nb] = e; rC[nb] = c; rU[nb] = u;
    // K4 pass A: per-block ranks through the look-back, 16 / 32 / 64 blocks per tile
    for (int G : {1, 2, 4}) {
      const u32 per = UR_BLOCKS * G, ntiles = (nb + per - 1) / per;
      std::vector<u64> st0(ntiles + 1, 0), st1(ntiles + 1, 0), st2(ntiles + 1, 0), gE(nb + 1, ~0ull), gC(nb + 1, ~0ull), gU(nb + 1, ~0ull);
      u64 totals[3] = {~0ull, ~0ull, ~0ull};
      u32 ticket = 0;
      Lookback<3> lb;
      lb.st[0] = st0.data(); lb.st[1] = st1.data(); lb.st[2] = st2.data(); lb.ticket = &ticket;
      if (G == 1) emu::launch(ntiles, 256, [&] { k_union_rank(bmE, bmC, lb, gE.data(), gC.data(), gU.data(), totals, nb, ntiles); });
      else if (G == 2) emu::launch(ntiles, 256, [&] { k_union_rank_g<2>(bmE, bmC, lb, gE.data(), gC.data(), gU.data(), totals, nb, ntiles); });
      else emu::launch(ntiles, 256, [&] { k_union_rank_g<4>(bmE, bmC, lb, gE.data(), gC.data(), gU.data(), totals, nb, ntiles); });
      bool ok = totals[0] == e && totals[1] == c && totals[2] == u;
      for (u32 b = 0; b < nb; b++) ok = ok && gE[b] == rE[b] && gC[b] == rC[b] && gU[b] == rU[b];
      if (!ok) { bad++; fprintf(stderr, "MISMATCH trial %d: union rank, %d blocks per tile\n", trial, UR_BLOCKS * G); }
      printf("trial %d: union rank, %2d blocks per tile  %s\n", trial, UR_BLOCKS * G, ok ? "ok" : "FAIL");
    }
    std::vector<float> ev(e + 2), cv(c + 2);
    for (auto& x : ev) x = (float)(rng() % 100000) / 7.0f;
    for (auto& x : cv) x = (float)(rng() % 100000) / 3.0f;
    // plain walk
    Out ref;
    ref.pEnd.resize(u); ref.pExpt.resize(u); ref.pCtrl.resize(u); ref.bmU.resize(T / 32); ref.cs.assign(L.nchrom, 0);
    {
      u64 ke = 0, kc = 0, ku = 0;
      for (u64 g = 0; g < T; g++) {
        const bool E = (bmE[g >> 5] >> (g & 31)) & 1, C = (bmC[g >> 5] >> (g & 31)) & 1;
        const int ch = b2c[g >> GR_BLOCK_SHIFT];
        if (g == off[ch]) ref.cs[ch] = ku;
        if (E || C) {
          ref.pEnd[ku] = (u32)(g - off[ch]); ref.pExpt[ku] = ev[ke]; ref.pCtrl[ku] = cv[kc];
          ref.bmU[g >> 5] |= 1u << (g & 31);
          ku++;
        }
        ke += E; kc += C;
      }
    }
    auto run = [&](int which) {
      Out o;
      u32* pEnd = dalloc<u32>(u + 1); float* pE = dalloc<float>(u + 1); float* pC = dalloc<float>(u + 1);
      u32* bmU = dalloc<u32>(T / 32); u64* cs = dalloc<u64>(L.nchrom + 1);
      memset(pEnd, 0xEE, (u + 1) * 4); memset(pE, 0xEE, (u + 1) * 4); memset(pC, 0xEE, (u + 1) * 4);
      memset(bmU, 0xEE, T / 8); memset(cs, 0xEE, (L.nchrom + 1) * 8);
      if (which == 4)
        emu::launch((nb + 3) / 4, 256, [&] { k_union_emit<4>(L, bmE, bmC, rE.data(), rC.data(), rU.data(), ev.data(), cv.data(),
                                                               pEnd, pE, pC, bmU, cs, nb); });
      else if (which == 2)
        emu::launch((nb + 1) / 2, 256, [&] { k_union_emit<2>(L, bmE, bmC, rE.data(), rC.data(), rU.data(), ev.data(), cv.data(),
                                                               pEnd, pE, pC, bmU, cs, nb); });
      else
        emu::launch((nb + 7) / 8, 256, [&] { k_union_emit_w(L, bmE, bmC, rE.data(), rC.data(), rU.data(), ev.data(), cv.data(),
                                                            pEnd, pE, pC, bmU, cs, nb); });
      o.pEnd.assign(pEnd, pEnd + u); o.pExpt.assign(pE, pE + u); o.pCtrl.assign(pC, pC + u);
      o.bmU.assign(bmU, bmU + T / 32); o.cs.assign(cs, cs + L.nchrom);
      free(pEnd); free(pE); free(pC); free(bmU); free(cs);
      return o;
    };
    const Out a = run(4), b2 = run(2), w = run(0);
    const bool ok = (a == ref) && (b2 == ref) && (w == ref);
    if (!ok) {
      bad++;
      fprintf(stderr, "MISMATCH trial %d: <4> %d <2> %d warp %d\n", trial, a == ref, b2 == ref, w == ref);
      for (size_t i = 0; i < u; i++)
        if (w.pEnd[i] != ref.pEnd[i] || memcmp(&w.pExpt[i], &ref.pExpt[i], 4) || memcmp(&w.pCtrl[i], &ref.pCtrl[i], 4)) {
          fprintf(stderr, "  warp form, first difference at %zu: end %u vs %u\n", i, w.pEnd[i], ref.pEnd[i]);
          break;
        }
    }
    printf("trial %d: %u blocks, %llu union intervals  %s\n", trial, nb, (unsigned long long)u, ok ? "ok" : "FAIL");
    free(bmE); free(bmC);
  }
  return bad ? 1 : 0;
}
