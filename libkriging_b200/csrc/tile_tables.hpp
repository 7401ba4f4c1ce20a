// tile_tables.hpp -- host-side builders of the tile tables that the DMMA tile engine (gemm_dmma.cuh) walks.
// Plain C++ (no CUDA): compiled into engine.cu and, on the CPU, into tests/cpp/tile_tables_selftest.cpp, which checks
// that every table covers its tile set exactly once for many sizes.  All orders are pure reorderings: a tile computes
// the same sums whatever its position, so results do not depend on them bit for bit; what they change is which operand
// strips the concurrently running CTAs share in L2 (profiles/r02c_table_orders.log).
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>

namespace lk {

constexpr int TM = 64;    // C rows per tile
constexpr int TN = 128;   // C cols per tile
constexpr int BLK = 128;  // factorisation block size; every n*n buffer is padded to a multiple of it

struct TileDesc {
  int c_row, c_col, k_begin, k_end;
};

namespace tables {

// A persistent grid of G CTAs walks a table in rounds (CTA b takes tiles b, b + G, ...).  In a table sorted by k-length
// CTA 0 would get the longest tile of EVERY round and CTA G - 1 the shortest: the CTAs drift apart by the length spread
// of a round per round, and tiles that share an operand strip (neighbours in the table, started together) stop
// meeting in L2.  Reversing every other round cancels the drift over two rounds.
inline void serpentine(std::vector<TileDesc>& t, size_t G) {
  if (G == 0) return;
  for (size_t c0 = G; c0 < t.size(); c0 += 2 * G) std::reverse(t.begin() + c0, t.begin() + std::min(t.size(), c0 + G));
}

// Lower tiles (2 ct <= rt) of an nb-block square matrix in bands of `band` row tiles, column by column inside a band.
// k range of tile (rt, ct): [rt * TM, N) (LAUUM: k_from_row) or [0, N) (the LOO product).
inline std::vector<TileDesc> lower_tiles_in_bands(int nb, int N, int band, bool k_from_row) {
  std::vector<TileDesc> t;
  band = std::max(1, band);
  for (int b0 = 0; b0 < 2 * nb; b0 += band) {
    const int b1 = std::min(2 * nb, b0 + band);
    for (int ct = 0; 2 * ct < b1; ++ct)
      for (int rt = std::max(b0, 2 * ct); rt < b1; ++rt) t.push_back({rt * TM, ct * TN, k_from_row ? rt * TM : 0, N});
  }
  return t;
}

// Lower tiles row by row (longest k first for LAUUM).
inline std::vector<TileDesc> lower_tiles_by_row(int nb, int N, bool k_from_row) {
  std::vector<TileDesc> t;
  for (int rt = 0; rt < 2 * nb; ++rt)
    for (int ct = 0; 2 * ct <= rt; ++ct) t.push_back({rt * TM, ct * TN, k_from_row ? rt * TM : 0, N});
  return t;
}

// Cholesky, look-ahead update of an outer block: region origin c1 (rows and columns), 2 * rem row tiles, OB column
// tiles, lower trapezoid tm >= 2 tn, k = [c0, c1); row by row (the OB N-side strips stay resident).
inline void chol_lookahead_tiles(std::vector<TileDesc>& out, int c0, int c1, int rem, int OB) {
  for (int tm = 0; tm < 2 * rem; ++tm)
    for (int tn = 0; tn < OB && 2 * tn <= tm; ++tn) out.push_back({c1 + tm * TM, c1 + tn * TN, c0, c1});
}

// Cholesky, rest of the trailing update: square region at r0 with mt row tiles and nt column tiles, lower trapezoid,
// k = [c0, c1); bands of `band` row tiles, column by column inside a band (the band's M-side strips stay resident).
inline void chol_rest_tiles(std::vector<TileDesc>& out, int c0, int c1, int r0, int mt, int nt, int band) {
  band = std::max(1, band);
  for (int b0 = 0; b0 < mt; b0 += band) {
    const int b1 = std::min(mt, b0 + band);
    for (int tn = 0; tn < nt && 2 * tn < b1; ++tn)
      for (int tm = std::max(b0, 2 * tn); tm < b1; ++tm) out.push_back({r0 + tm * TM, r0 + tn * TN, c0, c1});
  }
}

}  // namespace tables
}  // namespace lk
