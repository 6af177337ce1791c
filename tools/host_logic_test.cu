// Host-side checks of the GEMM planning code in gemm.cuh (no GPU needed): the CT32 index map is a
// bijection onto the padded matrix, the accumulator plan fits tensor memory, and the split-K choice
// respects the SM budget, the ring-stage budget of the exchange slots and the minimum chunk count.
// Built by __graft_entry__.build(), run by tests/test_host_logic.py.
#include <cstdio>
#include <cstring>
#include <vector>
#include "../oprl_b200/csrc/gemm.cuh"

using namespace oprl;

static int fails = 0;
#define CHECK(c)                                              \
  do {                                                        \
    if (!(c)) {                                               \
      printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);   \
      ++fails;                                                \
    }                                                         \
  } while (0)

static GemmOp op(int M, int N, int K) {
  GemmOp o;
  memset(&o, 0, sizeof(o));
  o.M = M; o.N = N; o.K = K;
  return o;
}

int main() {
  // ---- ct_index: every (r, c) of a padded matrix maps to a distinct offset inside it
  for (int rows : {32, 128, 256}) {
    for (int cols : {32, 96, 256}) {
      std::vector<char> hit(static_cast<size_t>(rows) * cols, 0);
      for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
          const size_t i = ct_index(rows, r, c);
          CHECK(i < hit.size());
          if (i < hit.size()) {
            CHECK(!hit[i]);
            hit[i] = 1;
          }
        }
      // a [8 rows x 32 cols] block is 1 KB contiguous, blocks of one 32-column chunk are consecutive
      CHECK(ct_index(rows, 8, 0) == 256);
      CHECK(ct_index(rows, 0, 4) == 32);
      if (cols > 32) CHECK(ct_index(rows, 0, 32) == static_cast<size_t>(rows / 8) * 256);
    }
  }
  CHECK(pad32(30) == 32 && pad32(32) == 32 && pad32(33) == 64 && pad128(6) == 128 && pad128(256) == 256);

  // ---- accumulator plan: at most 7 hi*hi accumulators, groups of >= 2 chunks, everything in 512 columns
  for (int K = 32; K <= 4096; K += 32) {
    GemmOp o = op(128, 32, K);
    gemm_finalize(o);
    const int nchunks = K / kBK;
    CHECK(o.group >= 2);
    CHECK(o.n_big >= 1 && o.n_big <= 7 || nchunks > 7 * o.group);
    CHECK(o.n_big * o.group >= nchunks);
    if (nchunks <= 49) CHECK(32 * (o.n_big + 1) + kASlots * kATmemCols <= 512);
  }

  // ---- split-K choice
  {
    GemmOp one = op(256, 256, 256);  // 16 tiles, 8 chunks
    CHECK(gemm_choose_ksplit(&one, 1, 148) == 4);
    GemmOp three[3] = {one, one, one};  // 48 tiles: 4-way would need 192 SMs
    CHECK(gemm_choose_ksplit(three, 3, 148) == 2);
    GemmOp k32 = op(256, 256, 32);  // a single chunk: nothing to split
    CHECK(gemm_choose_ksplit(&k32, 1, 148) == 1);
    GemmOp k96 = op(256, 256, 96);  // 3 chunks: fewer than 2 per CTA even at 2-way
    CHECK(gemm_choose_ksplit(&k96, 1, 148) == 1);
    GemmOp k128 = op(256, 256, 128);  // 4 chunks: 2-way yes, 4-way no
    CHECK(gemm_choose_ksplit(&k128, 1, 148) == 2);
    GemmOp big = op(1024, 256, 256);  // 64 tiles: 2-way fits (128), 4-way does not
    CHECK(gemm_choose_ksplit(&big, 1, 148) == 2);
    GemmOp five[5] = {big, big, big, big, big};  // 320 tiles: more than one wave already
    CHECK(gemm_choose_ksplit(five, 5, 148) == 1);
    GemmOp deep = op(128, 32, 1024);  // 32 chunks: any depth splits (the partial tiles land behind a shortened ring)
    CHECK(gemm_choose_ksplit(&deep, 1, 148) == 4);
    GemmOp k448 = op(128, 32, 448);  // 14 chunks
    CHECK(gemm_choose_ksplit(&k448, 1, 148) == 4);
    GemmOp k32s = op(128, 256, 32);  // 8 tiles
    GemmOp mixed[2] = {one, k32s};  // the deepest op decides; the shallow one's extra ranks leave at once (24 tiles x 4 = 96 CTAs)
    CHECK(gemm_choose_ksplit(mixed, 2, 148) == 4);
    GemmOp two[2] = {one, one};  // 32 tiles: 4-way would be 128 CTAs = 32 clusters, above the placement cap of 112
    CHECK(gemm_choose_ksplit(two, 2, 148) == 2);
    GemmOp longk = op(256, 256, 1024);  // ... but a 32-chunk K loop keeps the 4-way split
    GemmOp twolong[2] = {longk, longk};
    CHECK(gemm_choose_ksplit(twolong, 2, 148) == 4);
    CHECK(gemm_choose_ksplit(&one, 1, 32) == 2);  // a small part: SM budget caps the split
    CHECK(gemm_choose_ksplit(&one, 1, 16) == 1);
  }
  // ---- shared-memory budget of one CTA (227 KB opt-in limit) and the exchange slots
  CHECK(kGemmSmemBytes + 2048 <= 227 * 1024);
  CHECK(kGemmSmemBytesWide + 2048 <= 227 * 1024);
  CHECK(kStageFloats * 4 == kStageBytes && kAFloats * 4 <= kStageBytes);
  CHECK(3 * kAFloats <= 2 * kStageFloats);  // three 16 KB partial-tile slots fit the two ring stages a split-K CTA gives up
  // epilogue staging (transposed tile, dW_0 buffers, row-major tile) stays inside the shortest ring (split-K: 6 stages)
  CHECK((32 * kTTPitch + 2 * kBM * 36 + 8 * 32 * 36 + kBM * 33) * 4 <= (kStages - 2) * kStageBytes);
  CHECK((32 * kTTPitch + 2 * kBM * 36 + 8 * 32 * 36 + kBM * 33) * 4 <= kWideStages * (kAFloats + 4 * kBFloats) * 4);
  // ---- wide tiles: tile counts (ragged last tile), accumulator plan within 512 tensor-memory columns
  {
    GemmOp w = op(256, 512, 512);
    CHECK(gemm_tiles(w) == 32 && gemm_tiles(w, 2) == 16);
    GemmOp r = op(256, 96, 256);  // 3 sub-tiles of 32 columns: two wide tiles, the second ragged
    CHECK(gemm_tiles(r, 2) == 4);
    GemmOp n32 = op(256, 32, 256);
    CHECK(gemm_tiles(n32, 2) == 2 && gemm_wide_ok(n32));
    for (int K = 32; K <= 4096; K += 32) {
      GemmOp o = op(128, 64, K);
      gemm_finalize(o, kWideMaxBig);
      CHECK(o.n_big <= kWideMaxBig && o.n_big * o.group >= K / kBK);
      CHECK(64 * (o.n_big + 1) + kWideASlots * kATmemCols <= 512);
    }
  }
  printf("host logic: %s (%d failures)\n", fails ? "FAIL" : "PASS", fails);
  return fails ? 1 : 0;
}
