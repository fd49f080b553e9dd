// K2s, TMA route -- CSR times probe block for 5- / 7-diagonal band matrices (2-D 5-point and 3-D
// 7-point stencils) on the 256-wide fp32 tile, with the X rows staged in shared memory by TMA bulk copies.
//
// Why (profiles/r2f_spmm_2d_*.txt): the row-group kernel (spmm_csr.cu) is bound
// by the L1 data pipe (83.7 % of peak: 6 global-load requests and 12 shared-memory wavefronts per
// warp-row); the band kernel (spmm_strip.cu) halves that but keeps only 3 gathers per warp in flight,
// in registers, and becomes latency-bound.  Here the in-flight window lives in SHARED MEMORY instead.
// Two kernels, both producer / consumer pipelines (warp 8 produces, warps 0..7 consume, full / empty
// mbarriers, no CTA barrier in the steady state), both verifying band-ness per chunk of 16 rows
// from the CSR arrays as they are and falling back to a row-by-row gather path otherwise; the FMA
// order per row is the CSR order on every path, so W is bit-identical to the other kernels:
//   * spmm_tma_kernel ("chunked"): a CTA takes chunks of 16 consecutive rows in the row-group
//     kernel's order (static chunk -> CTA map, completed-chunk window; blocked order for 3-D); the
//     X rows a band chunk needs are three contiguous runs -- fetched by three cp.async.bulk copies
//     (SASS UBLKCP) a whole chunk ahead; (3R + 2) KB per stage, two stages, two CTAs per SM.  For 7
//     diagonals the two +-plane diagonals are gathered by the consumers instead of staged.
//   * spmm_walk_kernel ("strip walk", the default for 2-D stencils whose line length is known): a
//     CTA walks down a strip of the grid, so two of the three runs of a chunk are already in its
//     ring of line segments: ONE copy of R + 2 rows per chunk, 1.1 instead of 3.1 KB of L2 -> SM
//     traffic per row.
//   * the consumers read the staged rows with conflict-free LDS.128 (a warp reads 512 contiguous
//     bytes); a row-group walks consecutive rows, so the -1 / 0 / +1 diagonals slide through
//     registers and X[row] for the fused alpha dot (matfree/decomp.py:288) is the diagonal's register.
#include "internal.h"
#include "spmm_common.cuh"

namespace mf {
namespace {


__device__ __forceinline__ void tma_mbar_init(uint64_t* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint64_t* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU
__device__ __forceinline__ void tma_mbar_wait(uint64_t* bar, unsigned int parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
#pragma unroll 1
  for (unsigned int spin = 0; spin < (1u << 26); ++spin) {
    unsigned int done;
    asm volatile(
        "{\n .reg .pred p;\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned int bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          (uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
      "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

template <int VEC>
struct TmaRowDot {
  double d[VEC];
};
// one row on the gather path, out of line (metadata in shared memory)
template <typename T, int VEC, int LD, bool FUSE_DOT>
__device__ __noinline__ TmaRowDot<VEC> tma_gather_row(const int32_t* cols, const T* vals, int len,
                                                      const T* __restrict__ Xc, T* __restrict__ Wc,
                                                      int64_t off, const T* sc) {
  T sum[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) sum[q] = T(0);
  int u = 0;
  for (; u + 4 <= len; u += 4) {
    T x[4][VEC];
#pragma unroll
    for (int v = 0; v < 4; ++v) ldx<T, VEC>(Xc, (int64_t)cols[u + v] * LD, x[v]);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const T av = vals[u + v];
#pragma unroll
      for (int q = 0; q < VEC; ++q) sum[q] += av * x[v][q];
    }
  }
  for (; u < len; ++u) {
    const T av = vals[u];
    T x[VEC];
    ldx<T, VEC>(Xc, (int64_t)cols[u] * LD, x);
#pragma unroll
    for (int q = 0; q < VEC; ++q) sum[q] += av * x[q];
  }
  TmaRowDot<VEC> out;
  T w[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) w[q] = sum[q] * sc[q];
  stw<T, VEC>(Wc, off, w);
#pragma unroll
  for (int q = 0; q < VEC; ++q) out.d[q] = 0.0;
  if (FUSE_DOT) {
    T xo[VEC];
    ldx<T, VEC>(Xc, off, xo);
#pragma unroll
    for (int q = 0; q < VEC; ++q) out.d[q] = (double)(xo[q] * sc[q]) * (double)w[q];
  }
  return out;
}

__device__ __forceinline__ void tma_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}

// The producer warp's match of a chunk's rows against the diagonals `o` (lane r <-> row r): fills
// the row's coefficients by diagonal (zero where the entry is absent) and returns whether the row is
// on the band.  `full_rows`: every row of the chunk has all SEGL entries (the common case), then
// entry u must sit on diagonal u.  Otherwise (the x-boundary rows of a grid line miss an entry)
// every entry must sit on one of the diagonals and the columns must ascend strictly -- branch-free,
// all candidate entries in flight at once: that path runs for one chunk in eight on a 256^3 grid,
// and a serial match per entry made the producer late for the chunks behind it.
template <typename T, int SEGL>
__device__ __forceinline__ bool tma_match_row(bool ok, bool full_rows, const int32_t* colb,
                                              const T* valb, int jb, int len, int32_t row,
                                              const int32_t (&o)[SEGL], T (&dv)[SEGL]) {
#pragma unroll
  for (int u = 0; u < SEGL; ++u) dv[u] = T(0);
  if (!ok) return false;
  if (full_rows) {
#pragma unroll
    for (int u = 0; u < SEGL; ++u) {
      ok = ok && colb[jb + u] - row == o[u];
      dv[u] = valb[jb + u];
    }
    return ok;
  }
  ok = len <= SEGL;
  int matched = 0;
  int32_t dprev = 0;
#pragma unroll
  for (int e = 0; e < SEGL; ++e) {
    const bool live = e < len;
    const int32_t d = live ? colb[jb + e] - row : 0;
    const T a = live ? valb[jb + e] : T(0);
    ok = ok && (e == 0 || !live || d > dprev);
    dprev = d;
#pragma unroll
    for (int w = 0; w < SEGL; ++w) {
      const bool hit = live && o[w] == d;
      dv[w] = hit ? a : dv[w];
      matched += hit ? 1 : 0;
    }
  }
  return ok && matched == len;
}

// Dynamic shared memory layout (per CTA):
//   [NS stages][3R + 2 rows][LD] T     X rows: run A (R), run B (R + 2), run C (R)
//   [8][R + 1] int32                   producer-private: row pointers of chunks t .. t+5
//   [4][R * 8] int32, [4][R * 8] T     producer-private: column indices / values of chunks t .. t+3
//   [LD] T                             column scales
//   [NS] full + [NS] empty mbarriers, [NS] band flags, [NS][2] far offsets, [NS][R * SEGL] T dense coefficients
// 7-diagonal bands (3-D stencils) stage only the five inner diagonals: the two outermost ones
// (+- one plane) are gathered from global memory / L2 by the consumers, two loads per row issued
// before the wait on the stage.  Staging all seven costs (5R + 2) KB per stage, which leaves one
// CTA per SM at R = 16 and was slower than the row-group kernel (profiles/r2o_instep.jsonl).
template <typename T, int VEC, int LD, int SEGL, int ROWS>
struct TmaLayout {
  static constexpr int R = ROWS;  // rows per chunk
  static constexpr int FAR = SEGL == 7 ? 1 : 0;  // outermost diagonals per side that are gathered
  // the staged far diagonals (R rows each) + the run of the three adjacent middle ones (R + 2 rows)
  static constexpr int kStageRows = (SEGL - 3 - 2 * FAR) * R + R + 2;
  static constexpr int kEntCap = R * 8;
  static constexpr int NP = 8, NE = 4;  // ring depths (pointers: chunks t .. t+5, entries: t .. t+3), powers of two
  // stages of the X ring: two of 50 KB at R = 16, four of 26 KB at R = 8 (the same ~100 KB per CTA)
  static constexpr int NS = R <= 8 ? 4 : 2;
  static constexpr size_t kStageBytes = (size_t)kStageRows * LD * sizeof(T);
  static constexpr size_t kPtrOff = NS * kStageBytes;
  static constexpr size_t kColOff = (kPtrOff + NP * (R + 1) * sizeof(int32_t) + 15) / 16 * 16;
  static constexpr size_t kValOff = kColOff + NE * kEntCap * sizeof(int32_t);
  static constexpr size_t kSvOff = kValOff + NE * kEntCap * sizeof(T);
  static constexpr size_t kBarOff = (kSvOff + LD * sizeof(T) + 15) / 16 * 16;
  static constexpr size_t kFlagOff = kBarOff + 2 * NS * sizeof(uint64_t);
  static constexpr size_t kFarOff = kFlagOff + NS * sizeof(int);
  static constexpr size_t kDenseOff = kFarOff + 2 * NS * sizeof(int);
  static constexpr size_t kBytes = kDenseOff + NS * (size_t)R * SEGL * sizeof(T);
};

// Producer / consumer pipeline without CTA barriers in the steady state.  Warps 0..7 (kBlock
// threads) consume; warp 8 produces.
//   producer, chunk t: its CSR metadata was prefetched into private shared-memory rings by the
//     producer's own cp.async (entries three chunks ahead, row pointers five), so the verification
//     below reads shared memory only -- with the metadata loaded on demand the single warp's
//     dependent global loads were the critical path (7.7 vs 7.1 ms, tools/runs_r2_tma_v3.sh).
//     Band check (lane r <-> row r): the chunk has R rows whose columns all lie on the SEGL
//     diagonals row + o[u] (o from any row of the chunk that has all SEGL entries; o = .., -1, 0,
//     +1, ..); rows may miss entries (the x-boundary rows of a grid line): the lane files the row's
//     coefficients by diagonal into the stage's dense table, zero where the entry is absent --
//     adding 0 * x is exact, so W keeps the bits of the CSR-order sum.  Then: wait until the
//     consumers have released the stage (`empty`), publish table + flags, arm `full` with the
//     byte count and issue the TMA copies.  Every chunk arms `full` (a chunk that is not a band
//     with 0 bytes), so the phase parities depend on the chunk count alone.
//   consumer warp, chunk t: wait on `full`, compute its rows (band: LDS.128 + dense coefficients;
//     otherwise the gather path with metadata straight from global memory), arrive on `empty`.
template <typename T, int VEC, int LD, int SEGL, int ROWS, bool FUSE_DOT, bool BLOCKED>
__global__ void __launch_bounds__(kBlock + 32, ROWS <= 16 ? 2 : 1)
spmm_tma_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                const T* __restrict__ data, int64_t n, const T* __restrict__ X,
                const T* __restrict__ s, T* __restrict__ W, SpmmParams p,
                unsigned int* __restrict__ progress, double* __restrict__ partial, Finalize fin) {
  using L = TmaLayout<T, VEC, LD, SEGL, ROWS>;
  constexpr int R = L::R;
  constexpr int UD = SEGL / 2;
  constexpr int FAR = L::FAR;
  constexpr int MID = (UD - 1 - FAR) * R;  // first stage row of the middle run
  constexpr int NP = L::NP, NE = L::NE, NS = L::NS;
  static_assert(R < 32, "the producer warp checks one row per lane");
  static_assert(SEGL == 5 || SEGL == 7, "5- or 7-diagonal bands");
  constexpr int ld = LD;
  constexpr int tpr = LD / VEC;
  constexpr int rps = kBlock / tpr;  // row-groups per CTA
  constexpr int S = R / rps;         // consecutive rows per row-group
  constexpr int kConsumerWarps = kBlock / 32;
  static_assert(tpr >= 32 && tpr % 32 == 0 && R % rps == 0 && S >= 1, "tile / chunk geometry");
  extern __shared__ __align__(128) unsigned char smem[];
  T* const s_x = reinterpret_cast<T*>(smem);
  int32_t* const s_ptr = reinterpret_cast<int32_t*>(smem + L::kPtrOff);   // [NP][R + 1]
  int32_t* const s_col = reinterpret_cast<int32_t*>(smem + L::kColOff);   // [NE][kEntCap]
  T* const s_val = reinterpret_cast<T*>(smem + L::kValOff);               // [NE][kEntCap]
  T* const s_sv = reinterpret_cast<T*>(smem + L::kSvOff);
  uint64_t* const s_full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* const s_empty = s_full + NS;
  int* const s_band = reinterpret_cast<int*>(smem + L::kFlagOff);         // [2] per stage
  int* const s_far = reinterpret_cast<int*>(smem + L::kFarOff);           // [2][2] offsets of the gathered diagonals
  T* const s_dense = reinterpret_cast<T*>(smem + L::kDenseOff);           // [2][R][SEGL] coefficients by diagonal

  const bool producer = threadIdx.x >= kBlock;
  const int grp = threadIdx.x / tpr;
  const int lane = threadIdx.x & 31;
  const int c0 = (threadIdx.x % tpr) * VEC;
  const T* __restrict__ Xc = X + c0;
  T* __restrict__ Wc = W + c0;
  T sv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) sv[i] = (s && !producer) ? s[c0 + i] : T(1);
  for (int i = threadIdx.x; i < LD; i += kBlock + 32) s_sv[i] = s ? s[i] : T(1);
  double acc[1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = 0.0;

  const int64_t nchunks = (n + R - 1) / R;
  const int64_t G = gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      tma_mbar_init(&s_full[i], 1);
      tma_mbar_init(&s_empty[i], kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // mbarriers initialised; s_sv filled

  // first row of chunk c: ascending order, or the blocked order of SpmmParams (3-D stencils)
  auto row0_of = [&](int64_t c) -> int64_t {
    if constexpr (BLOCKED) return chunk_row0(c, p);
    else return c * (int64_t)R;
  };
  auto rows_of = [&](int64_t r0c) -> int { return (int)((n - r0c) < R ? (n - r0c) : R); };

  if (producer) {
    // ------------------------------------------------------------------ producer warp
    // my chunks are c_t = blockIdx.x + t G; ring slots by t
    auto issue_ptr = [&](int64_t t) {
      const int64_t c = blockIdx.x + t * G;
      if (c < nchunks) {
        const int64_t r0c = row0_of(c);
        const int nr = rows_of(r0c);
        if (lane <= nr)
          cp_async<4>((uint32_t)__cvta_generic_to_shared(&s_ptr[(int)(t & (NP - 1)) * (R + 1) + lane]),
                      indptr + r0c + lane);
      }
    };
    auto issue_ent = [&](int64_t t) {  // needs the pointers of chunk t visible
      const int64_t c = blockIdx.x + t * G;
      if (c < nchunks) {
        const int nr = rows_of(row0_of(c));
        const int32_t* ptrb = s_ptr + (int)(t & (NP - 1)) * (R + 1);
        const int32_t base = ptrb[0];
        const int total = ptrb[nr] - base;
        if (total <= L::kEntCap) {
          int32_t* cb = s_col + (int)(t & (NE - 1)) * L::kEntCap;
          T* vb = s_val + (int)(t & (NE - 1)) * L::kEntCap;
          for (int i = lane; i < total; i += 32) {
            cp_async<4>((uint32_t)__cvta_generic_to_shared(&cb[i]), indices + base + i);
            cp_async<(int)sizeof(T)>((uint32_t)__cvta_generic_to_shared(&vb[i]), data + base + i);
          }
        }
      }
    };
    for (int64_t t = 0; t < 5; ++t) issue_ptr(t);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    for (int64_t t = 0; t < 3; ++t) issue_ent(t);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    cp_async_commit();  // an empty group, so that "all but the newest group" below is uniform

    unsigned int seen_done = 0;
    int64_t ch = blockIdx.x;
    for (int64_t t = 0; ch < nchunks; ++t, ch += G) {
      const int stage = (int)(t & (NS - 1));
      // group G_{t-1} = {entries of chunk t+2, pointers of chunk t+4} may still fly; everything
      // older has landed: entries up to chunk t+1, pointers up to chunk t+3
      cp_async_wait<1>();
      __syncwarp();
      issue_ent(t + 3);
      issue_ptr(t + 5);
      cp_async_commit();
      if (progress != nullptr && lane == 0) {  // window throttle (see spmm_csr.cu)
        while ((int64_t)seen_done + p.window <= ch) {
          seen_done = *reinterpret_cast<volatile unsigned int*>(progress);
          if ((int64_t)seen_done + p.window <= ch) __nanosleep(200);
        }
      }
      __syncwarp();
      // ---- band check of chunk t
      const int64_t r0c = row0_of(ch);
      const int nr = rows_of(r0c);
      const int32_t* ptrb = s_ptr + (int)(t & (NP - 1)) * (R + 1);
      const int32_t base = ptrb[0];
      const int32_t* colb = s_col + (int)(t & (NE - 1)) * L::kEntCap;
      const T* valb = s_val + (int)(t & (NE - 1)) * L::kEntCap;
      int32_t o[SEGL];  // diagonal offsets
#pragma unroll
      for (int u = 0; u < SEGL; ++u) o[u] = 0;
      bool ok = nr == R && ptrb[nr] - base <= L::kEntCap;
      int jb = 0, len = 0;
      if (ok && lane < R) {
        jb = ptrb[lane] - base;
        len = ptrb[lane + 1] - base - jb;
      }
      const unsigned int fullm = __ballot_sync(0xffffffffu, ok && lane < R && len == SEGL);
      ok = ok && fullm != 0u;
      const int f = fullm ? __ffs((int)fullm) - 1 : 0;
      const int jbf = __shfl_sync(0xffffffffu, jb, f);
      if (ok) {
#pragma unroll
        for (int u = 0; u < SEGL; ++u) o[u] = colb[jbf + u] - (int32_t)(r0c + f);
        ok = o[UD] == 0 && o[UD - 1] == -1 && o[UD + 1] == 1;
#pragma unroll
        for (int u = 1; u < SEGL; ++u) ok = ok && o[u] > o[u - 1];
        // every X row the chunk touches must exist
        ok = ok && r0c + o[0] >= 0 && r0c + R - 1 + o[SEGL - 1] < n && r0c >= 1 && r0c + R < n;
      }
      // the row's coefficients by diagonal, in registers until the stage is free
      T dv[SEGL];
      ok = tma_match_row<T, SEGL>(ok && lane < R, fullm == (1u << R) - 1u, colb, valb, jb, len,
                                  (int32_t)(r0c + lane), o, dv) || (ok && lane >= R);
      const bool band = __all_sync(0xffffffffu, ok);
      // the consumers have released this stage (its use NS chunks ago)
      if (t >= NS) tma_mbar_wait(&s_empty[stage], (unsigned int)(((t / NS) + 1) & 1));
      if (band && lane < R) {
        T* dr = s_dense + (size_t)stage * R * SEGL + lane * SEGL;
#pragma unroll
        for (int u = 0; u < SEGL; ++u) dr[u] = dv[u];
      }
      if (lane == 0) {
        s_band[stage] = band ? 1 : 0;
        s_far[stage * 2 + 0] = o[0];
        s_far[stage * 2 + 1] = o[SEGL - 1];
      }
      __syncwarp();
      if (lane == 0) {
        T* xs = s_x + (size_t)stage * L::kStageRows * LD;
        constexpr unsigned int row_bytes = LD * sizeof(T);
        tma_mbar_expect_tx(&s_full[stage], band ? (unsigned int)L::kStageRows * row_bytes : 0u);
        if (band) {
#pragma unroll
          for (int u = FAR; u < UD - 1; ++u)  // staged far diagonals below the middle run
            tma_bulk_g2s(xs + (size_t)((u - FAR) * R) * LD, X + (r0c + o[u]) * LD, R * row_bytes, &s_full[stage]);
          tma_bulk_g2s(xs + (size_t)MID * LD, X + (r0c - 1) * LD, (R + 2) * row_bytes, &s_full[stage]);
#pragma unroll
          for (int u = UD + 2; u < SEGL - FAR; ++u)  // staged far diagonals above it
            tma_bulk_g2s(xs + (size_t)(MID + R + 2 + (u - UD - 2) * R) * LD, X + (r0c + o[u]) * LD,
                         R * row_bytes, &s_full[stage]);
        }
      }
    }
    cp_async_wait<0>();
  } else {
    // ------------------------------------------------------------------ consumer warps
    const int lr0 = grp * S;
    int far_lo = 0, far_hi = 0;  // offsets of the gathered diagonals in the last band chunk
    bool far_known = false;
    int64_t ch = blockIdx.x;
    for (int64_t t = 0; ch < nchunks; ++t, ch += G) {
      const int stage = (int)(t & (NS - 1));
      const int64_t r0 = row0_of(ch);
      const int64_t coff = r0 * ld;
      // The two outermost diagonals of a 7-diagonal band are gathered from global memory / L2, two
      // rows ahead of their use.  Their offsets come with the stage, but a stencil has the same ones
      // in every chunk: the first two rows are requested BEFORE the wait with the offsets of the last
      // band chunk (in flight while the stage lands) and again after it only if the offsets differ.
      constexpr int FD = S < 2 ? S : 2;
      T xf[FAR ? 2 : 1][FD][VEC];
      bool spec = false;
      if constexpr (FAR > 0) {
        const int64_t rlo = r0 + lr0 + far_lo, rhi = r0 + lr0 + far_hi;
        if (far_known && rlo >= 0 && rhi + S <= n) {
          spec = true;
#pragma unroll
          for (int i = 0; i < FD; ++i) {
            ldx<T, VEC>(Xc, (rlo + i) * LD, xf[0][i]);
            ldx<T, VEC>(Xc, (rhi + i) * LD, xf[1][i]);
          }
        }
      }
      tma_mbar_wait(&s_full[stage], (unsigned int)((t / NS) & 1));  // X rows / table of chunk t landed
      if (s_band[stage]) {
        int64_t rlo = 0, rhi = 0;
        if constexpr (FAR > 0) {
          const int flo = s_far[stage * 2 + 0], fhi = s_far[stage * 2 + 1];
          rlo = r0 + lr0 + flo;
          rhi = r0 + lr0 + fhi;
          if (!spec || flo != far_lo || fhi != far_hi) {
#pragma unroll
            for (int i = 0; i < FD; ++i) {
              ldx<T, VEC>(Xc, (rlo + i) * LD, xf[0][i]);
              ldx<T, VEC>(Xc, (rhi + i) * LD, xf[1][i]);
            }
          }
          far_lo = flo;
          far_hi = fhi;
          far_known = true;
        }
        const T* __restrict__ xs = s_x + (size_t)stage * L::kStageRows * LD + c0;
        const T* __restrict__ xb = xs + (size_t)(MID + lr0) * LD;  // middle run: rows lr, lr+1, lr+2
        T x[SEGL][VEC];
        vec_load<T>(xb, x[UD - 1]);
        vec_load<T>(xb + LD, x[UD]);
        const T* __restrict__ vrow = s_dense + (size_t)stage * R * SEGL + lr0 * SEGL;
        int64_t off = coff + (int64_t)lr0 * ld;
#pragma unroll
        for (int i = 0; i < S; ++i) {
          if constexpr (FAR > 0) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
              x[0][q] = xf[0][i % FD][q];
              x[SEGL - 1][q] = xf[1][i % FD][q];
            }
            if (i + FD < S) {
              ldx<T, VEC>(Xc, (rlo + i + FD) * LD, xf[0][i % FD]);
              ldx<T, VEC>(Xc, (rhi + i + FD) * LD, xf[1][i % FD]);
            }
          }
#pragma unroll
          for (int u = FAR; u < UD - 1; ++u) vec_load<T>(xs + (size_t)((u - FAR) * R + lr0 + i) * LD, x[u]);
          vec_load<T>(xb + (size_t)(i + 2) * LD, x[UD + 1]);
#pragma unroll
          for (int u = UD + 2; u < SEGL - FAR; ++u)
            vec_load<T>(xs + (size_t)(MID + R + 2 + (u - UD - 2) * R + lr0 + i) * LD, x[u]);
          T sum[VEC];
#pragma unroll
          for (int q = 0; q < VEC; ++q) sum[q] = T(0);
#pragma unroll
          for (int u = 0; u < SEGL; ++u) {
            const T av = vrow[i * SEGL + u];
#pragma unroll
            for (int q = 0; q < VEC; ++q) sum[q] += av * x[u][q];
          }
#pragma unroll
          for (int q = 0; q < VEC; ++q) sum[q] *= sv[q];
          if (FUSE_DOT) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc[0][q] += (double)(x[UD][q] * sv[q]) * (double)sum[q];
          }
          stw<T, VEC>(Wc, off, sum);
          off += ld;
#pragma unroll
          for (int q = 0; q < VEC; ++q) {
            x[UD - 1][q] = x[UD][q];
            x[UD][q] = x[UD + 1][q];
          }
        }
      } else {
        // not a band (first / last rows, anything irregular): gather path, metadata from global
        const int nr = rows_of(r0);
        for (int i = 0; i < S; ++i) {
          const int lr = lr0 + i;
          if (lr >= nr) break;
          const int32_t jb = __ldg(indptr + r0 + lr), len = __ldg(indptr + r0 + lr + 1) - jb;
          const TmaRowDot<VEC> d = tma_gather_row<T, VEC, LD, FUSE_DOT>(
              indices + jb, data + jb, len, Xc, Wc, coff + (int64_t)lr * ld, s_sv + c0);
          if (FUSE_DOT) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc[0][q] += d.d[q];
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        tma_mbar_arrive(&s_empty[stage]);  // this warp is done with the stage
        if (progress != nullptr && threadIdx.x == 0) atomicAdd(progress, 1u);
      }
    }
  }
  __syncthreads();  // all chunks of this CTA done; no copy is in flight (every armed phase was waited on)
  if (progress != nullptr && threadIdx.x == 0) {
    // the last CTA to leave re-arms the counters for the next launch
    __threadfence();
    const unsigned int left = atomicAdd(progress + 1, 1u);
    if (left == gridDim.x - 1) {
      progress[0] = 0u;
      progress[1] = 0u;
    }
  }
  if (FUSE_DOT) {
    // the X stages are dead: their memory serves the CTA-level reduction (no static array, so two
    // CTAs of ~105 KB fit one SM)
    cta_reduce_columns_smem<VEC, 1>(acc, ld, partial, 0, reinterpret_cast<double*>(smem));
    finalize_if_last<T>(ld, partial, 0, 1, fin);
  }
}

// ---------------------------------------------------------------- column walk (2-D 5-point bands)
// The kernel above moves every X row from L2 into shared memory three times (as part of the -line
// run, the middle run and the +line run of three different chunks): 54.6 GB of L2 -> SM traffic for
// 17.2 GB of X (profiles/r2zb_spmm_2d_tma.txt), and its consumers spend 43 % of their time waiting
// for a stage.  Here a CTA walks DOWN a strip of the grid instead: chunk (line l, strip s) = rows
// l L + s R .. + R (L = the line length = the operator's bandwidth hint), then (l + 1, s), ... for
// SEG lines.  Shared memory holds a ring of line segments X[l' L + s R - 1 .. + R + 2): the chunk
// of line l needs the segments l - 1, l, l + 1, two of which the previous chunk already used, so
// ONE TMA copy of R + 2 rows per chunk replaces three runs: 1.125 instead of 3.1 KB of L2 -> SM
// traffic per row.  Work items = (segment of SEG lines, strip), item i -> CTA i mod grid, strips
// of one segment adjacent in i, so the CTAs resident at any time read adjacent strips of the same
// lines and X comes from DRAM once.  Everything else (producer warp with private metadata rings,
// band check per chunk, dense coefficients, gather path for chunks that are not bands or touch
// the first / last line, FMA order) is the kernel above; no completed-chunk window is needed.
template <typename T, int LD, int SEGL, int ROWS>
struct WalkLayout {
  static constexpr int R = ROWS;
  static constexpr int NSL = 5;  // line-segment slots: the chunk's three + two in flight
  static constexpr int kSlotRows = R + 2;
  static constexpr int kEntCap = R * 8;
  static constexpr int NP = 8, NE = 4;
  static constexpr size_t kSlotBytes = (size_t)kSlotRows * LD * sizeof(T);
  static constexpr size_t kPtrOff = NSL * kSlotBytes;
  static constexpr size_t kColOff = (kPtrOff + NP * (R + 1) * sizeof(int32_t) + 15) / 16 * 16;
  static constexpr size_t kValOff = kColOff + NE * kEntCap * sizeof(int32_t);
  static constexpr size_t kSvOff = kValOff + NE * kEntCap * sizeof(T);
  static constexpr size_t kBarOff = (kSvOff + LD * sizeof(T) + 15) / 16 * 16;
  static constexpr size_t kFlagOff = kBarOff + 2 * NSL * sizeof(uint64_t);
  static constexpr size_t kDenseOff = kFlagOff + 8 * sizeof(int);
  static constexpr size_t kBytes = kDenseOff + NSL * (size_t)R * SEGL * sizeof(T);
};

// SEGL = 7 (3-D stencils: offsets -plane, -line, -1, 0, 1, line, plane): the five inner diagonals
// as above; the +-plane rows are gathered from global memory / L2 by the consumers, issued before
// the waits on the segments.  One work item = one plane's strip (seg = lines per plane), so the
// CTAs resident at a time walk the same lines of ~18 consecutive planes and the +-plane rows a
// CTA gathers are the rows its neighbours in z stage at about the same time.
template <typename T, int VEC, int LD, int SEGL, int ROWS, bool FUSE_DOT>
__global__ void __launch_bounds__(kBlock + 32, 2)
spmm_walk_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                 const T* __restrict__ data, int64_t n, const T* __restrict__ X,
                 const T* __restrict__ s, T* __restrict__ W, int64_t line, int64_t plane, int seg,
                 unsigned int* __restrict__ progress, int window, double* __restrict__ partial,
                 Finalize fin) {
  using L = WalkLayout<T, LD, SEGL, ROWS>;
  constexpr int R = L::R, UD = SEGL / 2, FAR = SEGL == 7 ? 1 : 0;
  static_assert(SEGL == 5 || SEGL == 7, "5- or 7-diagonal bands");
  constexpr int NP = L::NP, NE = L::NE, NSL = L::NSL;
  static_assert(R < 32, "the producer warp checks one row per lane");
  constexpr int ld = LD;
  constexpr int tpr = LD / VEC;
  constexpr int rps = kBlock / tpr;
  constexpr int S = R / rps;
  constexpr int kConsumerWarps = kBlock / 32;
  static_assert(tpr >= 32 && tpr % 32 == 0 && R % rps == 0 && S >= 1, "tile / chunk geometry");
  extern __shared__ __align__(128) unsigned char smem[];
  T* const s_x = reinterpret_cast<T*>(smem);                              // [NSL][R + 2][LD]
  int32_t* const s_ptr = reinterpret_cast<int32_t*>(smem + L::kPtrOff);   // [NP][R + 1]
  int32_t* const s_col = reinterpret_cast<int32_t*>(smem + L::kColOff);   // [NE][kEntCap]
  T* const s_val = reinterpret_cast<T*>(smem + L::kValOff);               // [NE][kEntCap]
  T* const s_sv = reinterpret_cast<T*>(smem + L::kSvOff);
  uint64_t* const s_full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* const s_empty = s_full + NSL;
  int* const s_band = reinterpret_cast<int*>(smem + L::kFlagOff);         // [NSL] by the slot of the newest segment
  T* const s_dense = reinterpret_cast<T*>(smem + L::kDenseOff);           // [NSL][R][SEGL]

  const bool producer = threadIdx.x >= kBlock;
  const int grp = threadIdx.x / tpr;
  const int lane = threadIdx.x & 31;
  const int c0 = (threadIdx.x % tpr) * VEC;
  const T* __restrict__ Xc = X + c0;
  T* __restrict__ Wc = W + c0;
  T sv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) sv[i] = (s && !producer) ? s[c0 + i] : T(1);
  for (int i = threadIdx.x; i < LD; i += kBlock + 32) s_sv[i] = s ? s[i] : T(1);
  double acc[1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = 0.0;

  const int64_t lines = n / line;
  const int strips = (int)(line / R);
  const int64_t items = (lines / seg) * strips;
  const int64_t G = gridDim.x;
  const int64_t my_items = (int64_t)blockIdx.x < items ? (items - blockIdx.x + G - 1) / G : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSL; ++i) {
      tma_mbar_init(&s_full[i], 1);
      tma_mbar_init(&s_empty[i], kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // first row of this CTA's chunk (item k, line y of its segment)
  auto chunk_row = [&](int64_t k, int y) -> int64_t {
    const int64_t item = blockIdx.x + k * G;
    const int64_t g = item / strips;
    const int sidx = (int)(item - g * strips);
    return (g * seg + y) * line + (int64_t)sidx * R;
  };

  if (producer) {
    // ------------------------------------------------------------------ producer warp
    const int64_t my_chunks = my_items * seg;
    auto row_of_t = [&](int64_t t) -> int64_t { return chunk_row(t / seg, (int)(t % seg)); };
    auto issue_ptr = [&](int64_t t) {
      if (t < my_chunks && lane <= R)
        cp_async<4>((uint32_t)__cvta_generic_to_shared(&s_ptr[(int)(t & (NP - 1)) * (R + 1) + lane]),
                    indptr + row_of_t(t) + lane);
    };
    auto issue_ent = [&](int64_t t) {  // needs the pointers of chunk t visible
      if (t < my_chunks) {
        const int32_t* ptrb = s_ptr + (int)(t & (NP - 1)) * (R + 1);
        const int32_t base = ptrb[0];
        const int total = ptrb[R] - base;
        if (total <= L::kEntCap) {
          int32_t* cb = s_col + (int)(t & (NE - 1)) * L::kEntCap;
          T* vb = s_val + (int)(t & (NE - 1)) * L::kEntCap;
          for (int i = lane; i < total; i += 32) {
            cp_async<4>((uint32_t)__cvta_generic_to_shared(&cb[i]), indices + base + i);
            cp_async<(int)sizeof(T)>((uint32_t)__cvta_generic_to_shared(&vb[i]), data + base + i);
          }
        }
      }
    };
    for (int64_t t = 0; t < 5; ++t) issue_ptr(t);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    for (int64_t t = 0; t < 3; ++t) issue_ent(t);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    cp_async_commit();  // an empty group, so that "all but the newest group" below is uniform

    int slot = 0;
    unsigned int wraps = 0;  // how often the slot ring has wrapped
    int64_t t = 0;           // chunk counter of this CTA
    unsigned int seen_done = 0;
    for (int64_t k = 0; k < my_items; ++k) {
      const int64_t item = blockIdx.x + k * G;
      const int64_t g = item / strips;
      const int sidx = (int)(item - g * strips);
      for (int j = 0; j < seg + 2; ++j) {
        // line segment j of the item: line g seg - 1 + j, rows first .. first + R + 2
        const int64_t lj = g * seg - 1 + j;
        const int64_t first = lj * line + (int64_t)sidx * R - 1;
        const bool valid = lj >= 0 && lj < lines && first >= 0 && first + R + 2 <= n;
        bool band = false;
        T dv[SEGL];
#pragma unroll
        for (int u = 0; u < SEGL; ++u) dv[u] = T(0);
        if (j >= 2) {
          // ---- chunk t = (k, y = j - 2).  Step throttle (3-D): the +-plane rows a CTA gathers are
          // in L2 only while its neighbours in z walk the same lines, so no CTA may run more than
          // `window` chunks (counted over the whole grid) ahead of the completed ones
          // (only over the steps every CTA has: in the last, partial wave of items the completed
          // count no longer grows by the whole grid per step)
          if (progress != nullptr && lane == 0 && t < (items / G) * seg) {
            const int64_t vch = t * G + blockIdx.x;
            unsigned int spins = 0;
            while ((int64_t)seen_done + window <= vch) {
              seen_done = *reinterpret_cast<volatile unsigned int*>(progress);
              if ((int64_t)seen_done + window <= vch) {
                __nanosleep(200);
                if (++spins > (1u << 26)) __trap();  // a protocol bug must not hang the GPU
              }
            }
          }
          __syncwarp();
          // metadata pipeline + band check
          cp_async_wait<1>();
          __syncwarp();
          issue_ent(t + 3);
          issue_ptr(t + 5);
          cp_async_commit();
          const int y = j - 2;
          const int64_t ly = g * seg + y;
          const int64_t r0c = ly * line + (int64_t)sidx * R;
          const int32_t* ptrb = s_ptr + (int)(t & (NP - 1)) * (R + 1);
          const int32_t base = ptrb[0];
          const int32_t* colb = s_col + (int)(t & (NE - 1)) * L::kEntCap;
          const T* valb = s_val + (int)(t & (NE - 1)) * L::kEntCap;
          // the three line segments of the chunk must exist (not the first / last line; the last
          // strip of the last but one line reaches one row past the matrix)
          bool ok = ly >= 1 && ly + 1 < lines && r0c - line - 1 >= 0 && r0c + line + R + 1 <= n &&
                    ptrb[R] - base <= L::kEntCap;
          if (FAR) ok = ok && r0c - plane >= 0 && r0c + R - 1 + plane < n;  // the gathered rows exist
          int jb = 0, len = 0;
          if (ok && lane < R) {
            jb = ptrb[lane] - base;
            len = ptrb[lane + 1] - base - jb;
          }
          int32_t o[SEGL];
          o[UD] = 0;
          o[UD - 1] = -1;
          o[UD + 1] = 1;
          o[UD - 2] = (int32_t)-line;
          o[UD + 2] = (int32_t)line;
          if constexpr (FAR > 0) {
            o[0] = (int32_t)-plane;
            o[SEGL - 1] = (int32_t)plane;
          }
          const unsigned int fullm = __ballot_sync(0xffffffffu, ok && lane < R && len == SEGL);
          ok = tma_match_row<T, SEGL>(ok && lane < R, fullm == (1u << R) - 1u, colb, valb, jb, len,
                                      (int32_t)(r0c + lane), o, dv) || (ok && lane >= R);
          band = __all_sync(0xffffffffu, ok);
          ++t;
        }
        // the consumers have released this slot (the segment NSL loads ago)
        if (wraps > 0) tma_mbar_wait(&s_empty[slot], (wraps - 1) & 1u);
        if (j >= 2) {
          if (band && lane < R) {
            T* dr = s_dense + (size_t)slot * R * SEGL + lane * SEGL;
#pragma unroll
            for (int u = 0; u < SEGL; ++u) dr[u] = dv[u];
          }
          if (lane == 0) s_band[slot] = band ? 1 : 0;
        }
        __syncwarp();
        if (lane == 0) {
          constexpr unsigned int seg_bytes = (unsigned int)L::kSlotBytes;
          tma_mbar_expect_tx(&s_full[slot], valid ? seg_bytes : 0u);
          if (valid) tma_bulk_g2s(s_x + (size_t)slot * L::kSlotRows * LD, X + first * LD, seg_bytes, &s_full[slot]);
        }
        if (++slot == NSL) {
          slot = 0;
          ++wraps;
        }
      }
    }
    cp_async_wait<0>();
  } else {
    // ------------------------------------------------------------------ consumer warps
    const int lr0 = grp * S;
    int sl0 = 0;             // slot of the chunk's oldest segment (line l - 1)
    unsigned int w0 = 0;     // ring wraps at sl0
    for (int64_t k = 0; k < my_items; ++k) {
      for (int y = 0; y < seg; ++y) {
        const int64_t r0 = chunk_row(k, y);
        const int64_t coff = r0 * ld;
        int sl1 = sl0 + 1, sl2 = sl0 + 2;
        unsigned int w1 = w0, w2 = w0;
        if (sl1 >= NSL) { sl1 -= NSL; ++w1; }
        if (sl2 >= NSL) { sl2 -= NSL; ++w2; }
        // the +-plane rows of a 7-diagonal band: gathered, in flight during the waits below
        // (two rows ahead: all S rows at once cost 32 registers and spilled)
        constexpr int FD = S < 2 ? S : 2;
        T xf[FAR ? 2 : 1][FD][VEC];
        const bool far_ok = FAR > 0 && r0 - plane >= 0 && r0 + R - 1 + plane < n;
        const int64_t rlo = r0 + lr0 - plane, rhi = r0 + lr0 + plane;
        if constexpr (FAR > 0) {
          if (far_ok) {
#pragma unroll
            for (int i = 0; i < FD; ++i) {
              ldx<T, VEC>(Xc, (rlo + i) * LD, xf[0][i]);
              ldx<T, VEC>(Xc, (rhi + i) * LD, xf[1][i]);
            }
          }
        }
        tma_mbar_wait(&s_full[sl0], w0 & 1u);
        tma_mbar_wait(&s_full[sl1], w1 & 1u);
        tma_mbar_wait(&s_full[sl2], w2 & 1u);  // also publishes the chunk's table
        if (s_band[sl2]) {
          const T* __restrict__ xa = s_x + ((size_t)sl0 * L::kSlotRows + 1 + lr0) * LD + c0;  // line l - 1
          const T* __restrict__ xb = s_x + ((size_t)sl1 * L::kSlotRows + lr0) * LD + c0;      // rows lr - 1 ..
          const T* __restrict__ xc = s_x + ((size_t)sl2 * L::kSlotRows + 1 + lr0) * LD + c0;  // line l + 1
          T x[SEGL][VEC];
          vec_load<T>(xb, x[UD - 1]);
          vec_load<T>(xb + LD, x[UD]);
          const T* __restrict__ vrow = s_dense + (size_t)sl2 * R * SEGL + lr0 * SEGL;
          int64_t off = coff + (int64_t)lr0 * ld;
#pragma unroll
          for (int i = 0; i < S; ++i) {
            if constexpr (FAR > 0) {
#pragma unroll
              for (int q = 0; q < VEC; ++q) {
                x[0][q] = xf[0][i % FD][q];
                x[SEGL - 1][q] = xf[1][i % FD][q];
              }
              if (i + FD < S) {  // the band flag implies far_ok
                ldx<T, VEC>(Xc, (rlo + i + FD) * LD, xf[0][i % FD]);
                ldx<T, VEC>(Xc, (rhi + i + FD) * LD, xf[1][i % FD]);
              }
            }
            vec_load<T>(xa + (size_t)i * LD, x[UD - 2]);
            vec_load<T>(xb + (size_t)(i + 2) * LD, x[UD + 1]);
            vec_load<T>(xc + (size_t)i * LD, x[UD + 2]);
            T sum[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q) sum[q] = T(0);
#pragma unroll
            for (int u = 0; u < SEGL; ++u) {
              const T av = vrow[i * SEGL + u];
#pragma unroll
              for (int q = 0; q < VEC; ++q) sum[q] += av * x[u][q];
            }
#pragma unroll
            for (int q = 0; q < VEC; ++q) sum[q] *= sv[q];
            if (FUSE_DOT) {
#pragma unroll
              for (int q = 0; q < VEC; ++q) acc[0][q] += (double)(x[UD][q] * sv[q]) * (double)sum[q];
            }
            stw<T, VEC>(Wc, off, sum);
            off += ld;
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
              x[UD - 1][q] = x[UD][q];
              x[UD][q] = x[UD + 1][q];
            }
          }
        } else {
          // not a band, or a chunk of the first / last line: gather path, metadata from global
          for (int i = 0; i < S; ++i) {
            const int lr = lr0 + i;
            const int32_t jb = __ldg(indptr + r0 + lr), len = __ldg(indptr + r0 + lr + 1) - jb;
            const TmaRowDot<VEC> d = tma_gather_row<T, VEC, LD, FUSE_DOT>(
                indices + jb, data + jb, len, Xc, Wc, coff + (int64_t)lr * ld, s_sv + c0);
            if (FUSE_DOT) {
#pragma unroll
              for (int q = 0; q < VEC; ++q) acc[0][q] += d.d[q];
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          if (progress != nullptr && threadIdx.x == 0) atomicAdd(progress, 1u);
          tma_mbar_arrive(&s_empty[sl0]);  // the segment of line l - 1 has had its last use
          if (y == seg - 1) {              // end of the item: so have the other two
            tma_mbar_arrive(&s_empty[sl1]);
            tma_mbar_arrive(&s_empty[sl2]);
          }
        }
        if (y == seg - 1) {
          // the next item starts three segments further
          sl0 += 3;
          if (sl0 >= NSL) { sl0 -= NSL; ++w0; }
        } else {
          sl0 = sl1;
          w0 = w1;
        }
      }
    }
  }
  __syncthreads();  // all chunks of this CTA done; every armed phase was waited on
  if (progress != nullptr && threadIdx.x == 0) {
    // the last CTA to leave re-arms the counters for the next launch
    __threadfence();
    const unsigned int left = atomicAdd(progress + 1, 1u);
    if (left == gridDim.x - 1) {
      progress[0] = 0u;
      progress[1] = 0u;
    }
  }
  if (FUSE_DOT) {
    cta_reduce_columns_smem<VEC, 1>(acc, ld, partial, 0, reinterpret_cast<double*>(smem));
    finalize_if_last<T>(ld, partial, 0, 1, fin);
  }
}

// On by default for 5-diagonal matrices: 6.8-7.6 ms per product inside the Lanczos step of BASELINE
// config 2 against 7.8-9.0 ms for the row-group kernel on the same pods, step 0.87-0.91 instead of
// 0.80-0.86 of the HBM roofline (profiles/r2n_instep.jsonl, r2o_instep.jsonl, r2w_sweep.jsonl).  The
// 7-diagonal variant is opt-in (MF_SPMM_TMA=2, mf_spmm_config(3, ...)): on the 3-D workload it is slower
// than the row-group kernel -- all seven diagonals staged: 18.7 ms; five staged + two gathered (this
// version): 15.8 vs 12.3 ms in the step; ncu (profiles/r2v_spmm_3d_tma.txt): L1 pipe 36 % instead of
// 65 %, but 30 GB of DRAM reads instead of 19 (the staged +-line runs and the gathered +-plane rows
// miss L2 more often than the row-group kernel's window) and 3.3 cycles per instruction at the
// per-chunk barrier.
std::atomic<int> g_tma{env_int("MF_SPMM_TMA", 1)};

}  // namespace

void spmm_tma_config(int use_tma) {
  if (use_tma >= 0) g_tma.store(use_tma);
}

int32_t launch_spmm_tma(const int32_t* indptr, const int32_t* indices, const void* data, int64_t n,
                        int64_t nnz, int32_t dtype, const void* X, const void* s, void* W,
                        int64_t ld, const Reduce* red, unsigned int* progress, cudaStream_t st,
                        bool* taken, int64_t bandwidth, int32_t num_diagonals, int64_t line_stride) {
  *taken = false;
  if (!g_tma.load(std::memory_order_relaxed) || n <= 0) return MF_OK;
  if (dtype != MF_F32 || ld != 256) return MF_OK;  // fp32, one 1 KB row per probe-tile row
  const double avg = (double)nnz / (double)n;
  static const int env_rows = env_int("MF_SPMM_TMA_ROWS", 16);
  // 7 diagonals: on by default where it was measured faster than the row-group kernel -- 3-D
  // stencils walked in the blocked row order (11.1 vs 11.8 ms per product on the 256^3 target,
  // profiles/r2zc_window.jsonl); elsewhere opt-in (MF_SPMM_TMA=2, mf_spmm_config(3, ...))
  bool tma7 = g_tma.load(std::memory_order_relaxed) >= 2;  // modes 2, 3, 4
  if (!tma7 && avg > 6.0 && avg <= 7.0) {
    SpmmParams probe{(int)ld, env_rows <= 8 ? 8 : 16, 0, 0, 0, 0, 0, 0, 0, 0};
    choose_row_order(&probe, n, avg, bandwidth, ld, dtype);
    tma7 = probe.block_rows != 0;
  }
  const int segl = (avg > 4.0 && avg <= 5.0) ? 5 : ((tma7 && avg > 6.0 && avg <= 7.0) ? 7 : 0);
  if (segl == 0 || n >= (1ll << 31) - 1 || nnz >= (1ll << 31) - 1) return MF_OK;
  // the host knows the diagonal count: only true 5- / 7-diagonal matrices (an irregular matrix that
  // averages 5 entries per row would run every chunk on the slow gather path of this kernel)
  if (num_diagonals > 0 && num_diagonals != segl) return MF_OK;
  if (((uintptr_t)X & 15) != 0) return MF_OK;
  static const int env_throttle = env_int("MF_SPMM_THROTTLE", 1);
  // completed-chunk window beyond the resident grid, in rows: wide in ascending order (C2: 6.6 ms at
  // 24576 rows, 7.6 at 8192, 10.8 at 0), two blocks in the blocked order (3-D target: 9.9 ms at 8192
  // or 12288 rows, 10.1 at 6144, 10.4 at 4096, 11.3 at 16384 -- a wider window spreads the CTAs over
  // more planes than L2 keeps; profiles/r2zc_window.jsonl, r2zp_sweep.jsonl, r2zq_sweep.jsonl)
  static const int env_window_rows = env_int("MF_SPMM_TMA_WINDOW", 0);
  Finalize fin{};
  double* partial = nullptr;
  if (red) {
    fin = red->fin;
    partial = red->partial;
  }
  unsigned int* prog = env_throttle ? progress : nullptr;
  *taken = true;
  // stencils whose line length is known (2-D: the bandwidth hint; 3-D: csr_line_stride): the strip walk
  static const int env_walk = env_int("MF_SPMM_WALK", 1);
  static const int env_walk_slack = env_int("MF_SPMM_WALK_SLACK", 3);  // line steps a CTA may run ahead (3-D)
  const int mode = g_tma.load(std::memory_order_relaxed);  // 3: walk whatever the size; 4: never
  const int64_t wline = segl == 5 ? bandwidth : line_stride;
  const int64_t wplane = segl == 5 ? 0 : bandwidth;
  // (3-D: measured slower than the chunked kernel, 13.1 vs 10.7 ms on the 256^3 target whatever the
  // step throttle -- the +-plane gathers sit in the consumers' critical path and shared memory has no
  // room to stage them as well; profiles/r2zg_walk.jsonl, r2zh_walk.jsonl -- so only on request)
  if (env_walk && mode != 4 && (segl == 5 || mode == 3) && wline >= 64 && wline % 16 == 0 && n % wline == 0 &&
      bandwidth < (1ll << 30) &&
      (segl == 5 ? (line_stride == 0 || line_stride == bandwidth)
                 : (wplane > wline && wplane % wline == 0 && n % wplane == 0))) {
    const int64_t lines = n / wline;
    const int64_t strips = wline / 16;
    // lines per work item.  2-D: the longest segment (a divisor of the line count, 16 .. 512) that
    // still leaves four items per resident CTA, else the shortest one; 3-D: one plane
    int seg = 0;
    if (segl == 7) {
      const int64_t lpp = wplane / wline;
      if (lpp >= 16 && lpp <= 4096) seg = (int)lpp;
    } else {
      for (int d = 512; d >= 16 && seg == 0; --d)
        if (lines % d == 0 && (lines / d) * strips >= 8 * (int64_t)num_sms()) seg = d;
      for (int d = 16; d <= 512 && seg == 0; ++d)
        if (lines % d == 0) seg = d;
    }
    const int64_t items = seg > 0 ? (lines / seg) * strips : 0;
    if (seg > 0 && (items >= 2 * num_sms() || mode == 3)) {
#define MF_WALK_K(SEGL, DOT)                                                                        \
  do {                                                                                              \
    using WL = WalkLayout<float, 256, SEGL, 16>;                                                    \
    auto kern = spmm_walk_kernel<float, 4, 256, SEGL, 16, DOT>;                                     \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WL::kBytes) != \
        cudaSuccess) {                                                                              \
      cudaGetLastError();                                                                           \
      *taken = false;                                                                               \
      return MF_OK;                                                                                 \
    }                                                                                               \
    int grid = resident_grid((const void*)kern, kBlock + 32, WL::kBytes, items);                    \
    /* 3-D: whole planes per wave of CTAs, and the step throttle (see the kernel) */                \
    if ((SEGL) == 7 && grid > strips) grid = (int)(grid / strips * strips);                         \
    kern<<<grid, kBlock + 32, WL::kBytes, st>>>(indptr, indices, (const float*)data, n,             \
                                                (const float*)X, (const float*)s, (float*)W,        \
                                                wline, wplane, seg, (SEGL) == 7 ? prog : nullptr,   \
                                                grid * (1 + env_walk_slack), partial, fin);         \
    return check_launch("spmm_walk");                                                               \
  } while (0)
      if (segl == 5) {
        if (red) MF_WALK_K(5, true);
        else MF_WALK_K(5, false);
      } else {
        if (red) MF_WALK_K(7, true);
        else MF_WALK_K(7, false);
      }
#undef MF_WALK_K
    }
  }
#define MF_TMA_K(SEGL, ROWS, DOT, BLK)                                                             \
  do {                                                                                             \
    using L = TmaLayout<float, 4, 256, SEGL, ROWS>;                                                \
    auto kern = spmm_tma_kernel<float, 4, 256, SEGL, ROWS, DOT, BLK>;                              \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kBytes) != \
        cudaSuccess) {                                                                             \
      cudaGetLastError();                                                                          \
      *taken = false;                                                                              \
      return MF_OK;                                                                                \
    }                                                                                              \
    const int grid = resident_grid((const void*)kern, kBlock + 32, L::kBytes, nchunks);            \
    /* the rows in flight stay one contiguous window (see env_window_rows above) */              \
    const int wrows = env_window_rows > 0 ? env_window_rows                                        \
                                          : (prm.block_rows != 0 ? 2 * prm.block_rows : 24576);    \
    prm.window = grid + wrows / L::R;                                                              \
    kern<<<grid, kBlock + 32, L::kBytes, st>>>(indptr, indices, (const float*)data, n,             \
                                               (const float*)X, (const float*)s, (float*)W, prm,   \
                                               prog, partial, fin);                                \
    return check_launch("spmm_tma");                                                               \
  } while (0)
#define MF_TMA_R(SEGL, ROWS)                                                                       \
  do {                                                                                             \
    const int64_t nchunks = (n + (ROWS) - 1) / (ROWS);                                             \
    SpmmParams prm{(int)ld, ROWS, 0, 0, 0, 0, 0, 0, 0, 0};                                         \
    choose_row_order(&prm, n, avg, bandwidth, ld, dtype);                                          \
    if (prm.block_rows != 0) {                                                                     \
      if (red) MF_TMA_K(SEGL, ROWS, true, true);                                                   \
      else MF_TMA_K(SEGL, ROWS, false, true);                                                      \
    } else {                                                                                       \
      if (red) MF_TMA_K(SEGL, ROWS, true, false);                                                  \
      else MF_TMA_K(SEGL, ROWS, false, false);                                                     \
    }                                                                                              \
  } while (0)
  if (segl == 5) {
    if (env_rows <= 8) MF_TMA_R(5, 8);
    else MF_TMA_R(5, 16);
  } else {
    if (env_rows <= 8) MF_TMA_R(7, 8);
    else MF_TMA_R(7, 16);
  }
#undef MF_TMA_R
#undef MF_TMA_K
}

}  // namespace mf
