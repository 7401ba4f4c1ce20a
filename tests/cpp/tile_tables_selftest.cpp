// CPU check of libkriging_b200/csrc/tile_tables.hpp: every table covers its tile set exactly once (sizes 1 .. 313 panels,
// several outer-block and band sizes), serpentine() is a permutation that reverses exactly the odd rounds.
#include <cstdio>
#include <set>
#include <tuple>

#include "../../libkriging_b200/csrc/tile_tables.hpp"

using lk::TileDesc;
using Key = std::tuple<int, int, int, int>;

static bool same_set(const std::vector<TileDesc>& got, const std::set<Key>& want, const char* what, int nb) {
  std::set<Key> s;
  for (const TileDesc& t : got) s.insert(Key{t.c_row, t.c_col, t.k_begin, t.k_end});
  if (s.size() != got.size() || s != want) {
    printf("FAIL %s nb=%d: %zu tiles, %zu distinct, %zu expected\n", what, nb, got.size(), s.size(), want.size());
    return false;
  }
  return true;
}

int main() {
  using namespace lk;
  bool ok = true;
  long long checked = 0;
  const int sizes[] = {1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 24, 40, 59, 79, 95, 96, 97, 141, 157, 158, 313};
  for (int nb : sizes) {
    const int N = nb * BLK;
    // lower tiles: LAUUM (k from the row) and the LOO product (whole k range)
    for (int kfr = 0; kfr < 2; ++kfr) {
      std::set<Key> want;
      for (int rt = 0; rt < 2 * nb; ++rt)
        for (int ct = 0; 2 * ct <= rt; ++ct) want.insert(Key{rt * TM, ct * TN, kfr ? rt * TM : 0, N});
      ok = same_set(tables::lower_tiles_by_row(nb, N, kfr), want, "lower_tiles_by_row", nb) && ok;
      for (int band : {1, 4, 8, 16, 32, 1000})
        ok = same_set(tables::lower_tiles_in_bands(nb, N, band, kfr), want, "lower_tiles_in_bands", nb) && ok;
      checked += 7;
    }
    // Cholesky trailing updates against the closed-form trapezoid of gemm_get_tile (SCHED_TRAP): tm >= 2 tn
    for (int OB : {3, 4, 6, 8})
      for (int J0 = 0; J0 < nb; J0 += OB) {
        const int J1 = std::min(nb, J0 + OB), rem = nb - J1, c0 = J0 * BLK, c1 = J1 * BLK;
        if (rem <= OB) continue;
        std::set<Key> la, rest;
        for (int tn = 0; tn < OB; ++tn)
          for (int tm = 2 * tn; tm < 2 * rem; ++tm) la.insert(Key{c1 + tm * TM, c1 + tn * TN, c0, c1});
        const int r0 = c1 + OB * BLK, mt = 2 * (rem - OB), nt = rem - OB;
        for (int tn = 0; tn < nt; ++tn)
          for (int tm = 2 * tn; tm < mt; ++tm) rest.insert(Key{r0 + tm * TM, r0 + tn * TN, c0, c1});
        std::vector<TileDesc> g1, g2;
        tables::chol_lookahead_tiles(g1, c0, c1, rem, OB);
        ok = same_set(g1, la, "chol_lookahead_tiles", nb) && ok;
        for (int band : {8, 32, 64}) {
          g2.clear();
          tables::chol_rest_tiles(g2, c0, c1, r0, mt, nt, band);
          ok = same_set(g2, rest, "chol_rest_tiles", nb) && ok;
        }
        checked += 4;
      }
  }
  // serpentine: even rounds untouched, odd rounds reversed
  for (size_t len : {0u, 5u, 296u, 297u, 592u, 1000u, 12403u}) {
    std::vector<TileDesc> t(len);
    for (size_t i = 0; i < len; ++i) t[i] = {(int)i, 0, 0, 0};
    tables::serpentine(t, 296);
    for (size_t i = 0; i < len; ++i) {
      const size_t round = i / 296, lo = round * 296, hi = std::min(len, lo + 296);
      const size_t want = (round & 1) ? (lo + (hi - 1 - i)) : i;
      if ((size_t)t[i].c_row != want) { printf("FAIL serpentine len=%zu i=%zu\n", len, i); ok = false; break; }
    }
    ++checked;
  }
  printf("{\"ok\": %s, \"tables_checked\": %lld}\n", ok ? "true" : "false", checked);
  return ok ? 0 : 1;
}
