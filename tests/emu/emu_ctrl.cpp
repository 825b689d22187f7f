// emu_ctrl.cpp -- TEST INFRASTRUCTURE (CPU): k_ctrl_clamp_m<4> (GR_CL_TILES=4: 32768 raw intervals per
// look-back tile, 128-wide window) against k_ctrl_clamp (the default, validated on the B200) and a
// plain loop: clamped control pileup (savePileupCtrl 2107-2141), chromosome starts, break bitmap.
#include "cuda_emu.h"
#include "gen_kernels_ctrl.h"
#include <random>

template <typename T> static T* dalloc(size_t n) { return (T*)aligned_alloc(64, ((n * sizeof(T) + 63) / 64 + 1) * 64); }

int main() {
  std::mt19937_64 rng(4242);
  int bad = 0;
  for (int trial = 0; trial < 3; trial++) {
    // chromosomes with many / few / no raw intervals (the third has no slots at all in trial 1)
    std::vector<u32> want = {(u32)(60000 + trial * 45000), 1, 0, 9000, (u32)(30000 * trial)};
    const int nc = (int)want.size();
    std::vector<u64> off(nc); std::vector<u32> len(nc); std::vector<uint8_t> flags(nc, 3); std::vector<int> b2c;
    std::vector<u32> end; std::vector<float> val; std::vector<u64> cs(nc + 1);
    u64 T = 0;
    static const float V[8] = {0.0f, 0.5f, 1.0f, 1.0f, 2.0f, 3.5f, 7.25f, -1.0f};
    for (int c = 0; c < nc; c++) {
      cs[c] = end.size();
      u32 pos = 0;
      for (u32 i = 0; i < want[c]; i++) {
        pos += 1 + (u32)(rng() % 40);
        end.push_back(pos);
        val.push_back(V[rng() % 8]);
      }
      len[c] = want[c] ? pos : 1000;
      if (!want[c] && trial == 1) { off[c] = ~0ull; continue; }
      off[c] = T;
      const u64 nb = ((u64)len[c] + 1 + GR_BLOCK_SLOTS - 1) / GR_BLOCK_SLOTS;
      for (u64 b = 0; b < nb; b++) b2c.push_back(c);
      T += nb * GR_BLOCK_SLOTS;
    }
    cs[nc] = end.size();
    const u64 n = end.size();
    DevLayout L; L.nchrom = nc; L.T = T; L.nblocks = T / GR_BLOCK_SLOTS;
    L.off = off.data(); L.len = len.data(); L.flags = flags.data(); L.blk2chrom = b2c.data();
    val.push_back(123.0f);                              // val[n]: readable, never decisive
    const float fl[2] = {1.5f, 1.4f};                   // scale factor, lambda
    std::vector<u32> bm0(T / 32, 0);
    for (int c = 0; c < nc; c++)
      for (u64 i = cs[c]; i < cs[c + 1]; i++) { const u64 g = off[c] + end[i]; bm0[g >> 5] |= 1u << (g & 31); }
    // plain loop
    std::vector<u32> rEnd; std::vector<float> rVal; std::vector<u64> rCs(nc + 1); std::vector<u32> rBm = bm0;
    for (int c = 0; c < nc; c++) {
      rCs[c] = rEnd.size();
      for (u64 i = cs[c]; i < cs[c + 1]; i++) {
        const float net = clamp_net(fl[0], val[i], fl[1]);
        const bool last = i + 1 == cs[c + 1];
        if (last || net != clamp_net(fl[0], val[i + 1], fl[1])) { rEnd.push_back(end[i]); rVal.push_back(net); }
        else { const u64 g = off[c] + end[i]; rBm[g >> 5] &= ~(1u << (g & 31)); }
      }
    }
    rCs[nc] = rEnd.size();
    auto run = [&](int m, std::vector<u32>& oEnd, std::vector<float>& oVal, std::vector<u64>& oCs, std::vector<u32>& oBm, u64& oTot) {
      u64 tot_in = n;
      DevRle raw{end.data(), val.data(), cs.data(), &tot_in};
      oEnd.assign(n + 1, 0xEEEEEEEEu); oVal.assign(n + 1, -77.0f); oCs.assign(nc + 1, ~0ull); oBm = bm0; oTot = 0;
      DevRle out{oEnd.data(), oVal.data(), oCs.data(), &oTot};
      const u64 ntiles = (n + (u64)CL_TILE * m - 1) / ((u64)CL_TILE * m);
      std::vector<u64> st(ntiles + 1, 0); u32 ticket = 0;
      Lookback<1> lb; lb.st[0] = st.data(); lb.ticket = &ticket;
      if (m == 1) emu::launch((unsigned)ntiles, 256, [&] { k_ctrl_clamp(L, raw, fl, lb, out, oBm.data()); });
      else emu::launch((unsigned)ntiles, 256, [&] { k_ctrl_clamp_m<4>(L, raw, fl, lb, out, oBm.data()); });
      emu::launch(1, 32, [&] { k_fill_forward(nc, cs.data(), oCs.data()); });
      oEnd.resize(oTot); oVal.resize(oTot);
    };
    std::vector<u32> aE, bE, aB, bB; std::vector<float> aV, bV; std::vector<u64> aC, bC; u64 aT, bT;
    run(1, aE, aV, aC, aB, aT);
    run(4, bE, bV, bC, bB, bT);
    auto same = [&](std::vector<u32>& E, std::vector<float>& Vv, std::vector<u64>& C, std::vector<u32>& B, u64 Tt) {
      return Tt == rEnd.size() && E == rEnd && !memcmp(Vv.data(), rVal.data(), rVal.size() * 4) && C == rCs && B == rBm;
    };
    const bool ok1 = same(aE, aV, aC, aB, aT), ok4 = same(bE, bV, bC, bB, bT);
    if (!ok1 || !ok4) bad++;
    printf("trial %d: %llu raw -> %zu clamped control intervals   k_ctrl_clamp %s   k_ctrl_clamp_m<4> %s\n", trial,
           (unsigned long long)n, rEnd.size(), ok1 ? "ok" : "FAIL", ok4 ? "ok" : "FAIL");
  }
  return bad ? 1 : 0;
}
