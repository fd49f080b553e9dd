// Host-side check of the blocked row order of the CSR product (spmm_common.cuh: choose_row_order,
// chunk_row0): for a set of grids the chunk -> first-row map must be a permutation of the chunk
// starts 0, R, 2R, ..., and the same row range of consecutive planes must be adjacent in it.
// Compiled and run by tests/test_host.py (nvcc, no GPU needed: no CUDA call is made).
#include <cstdio>
#include <vector>

#include "spmm_common.cuh"

using namespace mf;

static int check(int64_t n, int64_t bandwidth, int R, int64_t ld, bool expect_blocked) {
  SpmmParams p{(int)ld, R, 0, 0, 0, 0, 0, 0, 0, 0};
  const double avg = 7.0;
  choose_row_order(&p, n, avg, bandwidth, ld, MF_F32);
  const bool blocked = p.block_rows != 0;
  if (blocked != expect_blocked) {
    std::printf("FAIL n=%lld bw=%lld R=%d ld=%lld: blocked=%d expected %d\n", (long long)n,
                (long long)bandwidth, R, (long long)ld, (int)blocked, (int)expect_blocked);
    return 1;
  }
  const int64_t nchunks = (n + R - 1) / R;
  std::vector<char> seen((size_t)nchunks, 0);
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t r0 = chunk_row0(c, p);
    if (r0 < 0 || r0 >= n || r0 % R != 0 || seen[(size_t)(r0 / R)]) {
      std::printf("FAIL n=%lld bw=%lld R=%d: chunk %lld -> row %lld\n", (long long)n,
                  (long long)bandwidth, R, (long long)c, (long long)r0);
      return 1;
    }
    seen[(size_t)(r0 / R)] = 1;
  }
  if (blocked) {
    // consecutive blocks of the order are the same row range of consecutive planes
    const int64_t cpb = p.block_rows / R;  // chunks per block
    if (p.block_rows % R != 0 || bandwidth % p.block_rows != 0) return 1;
    const int64_t a = chunk_row0(0, p), b = chunk_row0(cpb, p);
    if (b - a != bandwidth) {
      std::printf("FAIL n=%lld bw=%lld R=%d: second block starts %lld rows after the first\n",
                  (long long)n, (long long)bandwidth, R, (long long)(b - a));
      return 1;
    }
  }
  return 0;
}

int main() {
  int bad = 0;
  bad += check(256ll * 256 * 256, 256 * 256, 64, 256, true);   // the 3-D target, row-group kernel
  bad += check(256ll * 256 * 256, 256 * 256, 16, 256, true);   // ... TMA kernel chunks
  bad += check(6ll * 256 * 256, 256 * 256, 64, 256, true);     // the GPU test's grid
  bad += check(128ll * 384 * 384, 384 * 384, 64, 256, true);   // planes that are not powers of two
  bad += check(256ll * 256 * 256, 256 * 256, 64, 32, false);   // narrow tile: planes fit L2
  bad += check(4096ll * 4096, 4096, 64, 256, false);           // 2-D: ascending order
  bad += check(2ll * 256 * 256, 256 * 256, 64, 256, false);    // two planes only
  std::printf(bad ? "row order: %d failures\n" : "row order: ok\n", bad);
  return bad ? 1 : 0;
}
