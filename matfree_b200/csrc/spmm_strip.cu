// K2s, band route -- CSR times probe block for banded / stencil matrices on wide tiles.
//
// profiles/r1m_*, r1o_*: on BASELINE config 2 the row-group kernel of spmm_csr.cu is bound by the
// L1 data pipe (l1tex__data_pipe_lsu_wavefronts 86 %): per 1 KB row 40 wavefronts of gathers, 8 of
// the fused dot's reload of X[row], 8 of stores, 24 of shared-memory metadata reads -- 80 against
// a budget of 78 clocks per row at the HBM roof.  This kernel removes the avoidable ones:
//   * a row-group walks a STRIP of S consecutive rows, so the X rows shared by consecutive rows of
//     a band (the -1 / 0 / +1 diagonals of a stencil: row r+1 needs X[r], X[r+1], X[r+2]) stay in
//     REGISTERS: per row only the top of every run of adjacent diagonals is loaded (3 of 5 gathers
//     for a 2-D 5-point row, 5 of 7 for a 3-D 7-point row), and X[row] for the fused alpha dot
//     (matfree/decomp.py:288) is the diagonal's register, not a reload;
//   * the band structure is found per strip on the fly from the CSR arrays as they are (no format
//     change, no preprocessing, nothing cached): every row of the strip has SEGL entries, the
//     columns of row i are those of row 0 shifted by i, the middle three are adjacent with the
//     diagonal in the centre.  A strip that is not such a band (boundary rows, anything else)
//     takes the gather path row by row.  The FMA order per row is the CSR order either way, so W
//     is bit-identical to the row-group kernel's;
//   * the column indices / values of the next chunk arrive by ONE bulk copy each (TMA:
//     cp.async.bulk + mbarrier, SASS UBLKCP) issued by a single thread, instead of ~10 four-byte
//     LDGSTS per row through the LSU; column indices are only read by the strip detection.
// Chunk scheduling (static chunk -> CTA map, completed-chunk window that keeps the rows in flight
// contiguous so a stencil's neighbours stay in L2) is the row-group kernel's.
#include "internal.h"
#include "spmm_common.cuh"

namespace mf {
namespace {

constexpr int kStripPad = 8;  // staging slack for the 16-byte alignment of the bulk copies

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned int parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n .reg .pred p;\n"
      "MF_WAIT:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra MF_DONE;\n"
      " bra MF_WAIT;\n"
      "MF_DONE:\n}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned int bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          (uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
      "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// One row on the gather path (strips that are not a full band, chunks whose entries did not fit
// the staging buffer): kept out of line so that its registers do not compete with the band
// loop's.  `cols` / `vals` may point to shared or global memory.
template <int VEC>
struct RowDot {
  double d[VEC];
};
template <typename T, int VEC, int LD, bool FUSE_DOT>
__device__ __noinline__ RowDot<VEC> strip_gather_row(const int32_t* cols, const T* vals, int len,
                                                     const T* __restrict__ Xc, T* __restrict__ Wc,
                                                     int64_t off, const T* sc) {
  T sum[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) sum[q] = T(0);
  int u = 0;
  for (; u + 4 <= len; u += 4) {  // four independent gathers in flight
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
  RowDot<VEC> out;
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

// SEGL diagonals d_0 < ... < d_{SEGL-1} with the three in the middle adjacent (-1, 0, +1): a 2-D
// 5-point (SEGL = 5) or 3-D 7-point (SEGL = 7) stencil, or any matrix with that local structure.
template <typename T, int VEC, int LD, int SEGL, bool FUSE_DOT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB)
spmm_strip_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                  const T* __restrict__ data, int64_t n, int64_t nnz, const T* __restrict__ X,
                  const T* __restrict__ s, T* __restrict__ W, SpmmParams p,
                  unsigned int* __restrict__ progress, double* __restrict__ partial,
                  Finalize fin) {
  __shared__ int32_t s_ptr[3][kMaxRows + 1];
  __shared__ __align__(16) int32_t s_col[2][kCap + kStripPad];
  __shared__ __align__(16) T s_val[2][kCap + kStripPad];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ int s_skew[2][2];  // [buffer][0: columns, 1: values] element offset of entry 0
  __shared__ T s_sv[LD];        // the column scales, for the out-of-line gather rows

  constexpr int ld = LD;
  constexpr int tpr = LD / VEC;      // threads per row (>= 32: a row-group is whole warps)
  constexpr int rps = kBlock / tpr;  // row-groups per CTA
  constexpr int CA = 4;              // int32 elements per 16 bytes
  constexpr int VA = 16 / (int)sizeof(T);
  constexpr int UD = SEGL / 2;       // position of the diagonal in a full band row
  static_assert(tpr >= 32 && tpr % 32 == 0, "strip kernel: a row-group must be whole warps");
  static_assert(SEGL == 5 || SEGL == 7, "strip kernel: 5- or 7-diagonal bands");
  const int grp = threadIdx.x / tpr;
  const int lane = threadIdx.x & 31;
  const int c0 = (threadIdx.x % tpr) * VEC;
  const T* __restrict__ Xc = X + c0;
  T* __restrict__ Wc = W + c0;
  const int R = p.rows_per_chunk;
  const int S = R / rps;  // rows per strip (<= 32)
  T sv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) sv[i] = s ? s[c0 + i] : T(1);
  for (int i = threadIdx.x; i < LD; i += kBlock) s_sv[i] = s ? s[i] : T(1);
  double acc[1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = 0.0;

  const int64_t nchunks = (n + R - 1) / R;
  const int64_t G = gridDim.x;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  auto issue_ptr = [&](int64_t c, int buf) {  // row pointers of chunk c -> s_ptr[buf]
    if (c < nchunks) {
      const int64_t r0 = chunk_row0(c, p);
      const int nr = (int)((n - r0) < R ? (n - r0) : R);
      for (int i = threadIdx.x; i <= nr; i += kBlock)
        cp_async<4>((uint32_t)__cvta_generic_to_shared(&s_ptr[buf][i]), indptr + r0 + i);
    }
  };
  // entries of chunk c -> s_col / s_val [ebuf]; needs s_ptr[pbuf] visible.  One thread arms the
  // buffer's mbarrier with the byte count and issues two bulk copies whose source is rounded
  // down to 16 bytes (entry 0 then sits at element `skew`); every chunk arms its barrier
  // (oversized chunks with 0 bytes) so the phase parity is a function of the chunk count alone.
  auto issue_ent = [&](int64_t c, int pbuf, int ebuf) {
    if (c < nchunks && threadIdx.x == 0) {
      const int64_t r0 = chunk_row0(c, p);
      const int nr = (int)((n - r0) < R ? (n - r0) : R);
      const int32_t base = s_ptr[pbuf][0];
      const int total = s_ptr[pbuf][nr] - base;
      unsigned int bytes_c = 0, bytes_v = 0;
      int skc = 0, skv = 0;
      if (total > 0 && total <= kCap) {
        skc = base & (CA - 1);
        skv = base & (VA - 1);
        int64_t cnt_c = (int64_t)(skc + total + CA - 1) / CA * CA;
        int64_t cnt_v = (int64_t)(skv + total + VA - 1) / VA * VA;
        // never read past the arrays: the (at most 15) bytes of rounding at the very end of
        // indices / data are dropped from the bulk copy and fetched by plain loads
        const int64_t lim_c = nnz - (base - skc), lim_v = nnz - (base - skv);
        int tail_c = 0, tail_v = 0;
        if (cnt_c > lim_c) {
          tail_c = (int)(lim_c % CA);
          cnt_c = lim_c - tail_c;
        }
        if (cnt_v > lim_v) {
          tail_v = (int)(lim_v % VA);
          cnt_v = lim_v - tail_v;
        }
        bytes_c = (unsigned int)(cnt_c * 4);
        bytes_v = (unsigned int)(cnt_v * (int64_t)sizeof(T));
        for (int i = 0; i < tail_c; ++i)
          s_col[ebuf][cnt_c + i] = __ldg(indices + base - skc + cnt_c + i);
        for (int i = 0; i < tail_v; ++i)
          s_val[ebuf][cnt_v + i] = __ldg(data + base - skv + cnt_v + i);
      }
      s_skew[ebuf][0] = skc;
      s_skew[ebuf][1] = skv;
      mbar_expect_tx(&s_bar[ebuf], bytes_c + bytes_v);
      if (bytes_c) bulk_g2s(&s_col[ebuf][0], indices + base - skc, bytes_c, &s_bar[ebuf]);
      if (bytes_v) bulk_g2s(&s_val[ebuf][0], data + base - skv, bytes_v, &s_bar[ebuf]);
    }
  };

  int64_t ch = blockIdx.x;
  issue_ptr(ch, 0);
  issue_ptr(ch + G, 1);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();  // row pointers of the first chunk visible; mbarriers initialised; s_sv filled
  issue_ent(ch, 0, 0);

  unsigned int seen_done = 0;
  for (int64_t t = 0; ch < nchunks; ++t, ch += G) {
    const int pb = (int)(t % 3), eb = (int)(t & 1);
    if (progress != nullptr && threadIdx.x == 0) {
      while ((int64_t)seen_done + p.window <= ch) {
        seen_done = *reinterpret_cast<volatile unsigned int*>(progress);
        if ((int64_t)seen_done + p.window <= ch) __nanosleep(200);
      }
    }
    __syncthreads();  // row pointers of chunks t, t+1 visible; everyone is done with chunk t-1
    if (progress != nullptr && threadIdx.x == 0 && t > 0) atomicAdd(progress, 1u);
    issue_ent(ch + G, (int)((t + 1) % 3), (int)((t + 1) & 1));
    issue_ptr(ch + 2 * G, (int)((t + 2) % 3));
    cp_async_commit();

    const int64_t r0 = chunk_row0(ch, p);
    const int nr = (int)((n - r0) < R ? (n - r0) : R);
    const int32_t* __restrict__ ptrb = s_ptr[pb];
    const int32_t base = ptrb[0];
    const int total = ptrb[nr] - base;
    const int64_t coff = r0 * ld;

    mbar_wait(&s_bar[eb], (unsigned int)((t >> 1) & 1));  // entries of chunk t have landed
    const int32_t* __restrict__ colb = s_col[eb] + s_skew[eb][0];
    const T* __restrict__ valb = s_val[eb] + s_skew[eb][1];

    // ---- this row-group's strip: rows [a, a + ns) of the chunk
    const int a = grp * S;
    const int ns = nr - a < S ? (nr - a > 0 ? nr - a : 0) : S;
    bool band = false;
    int jb0 = 0;
    if (total <= kCap && ns > 0) {
      // Is the strip a full band?  Lane i checks row i.
      jb0 = ptrb[a] - base;
      bool ok = true;
      if (lane < ns) {
        const int jb = ptrb[a + lane] - base;
        ok = (ptrb[a + lane + 1] - base - jb == SEGL) && (jb == jb0 + lane * SEGL);
        if (ok) {
#pragma unroll
          for (int u = 0; u < SEGL; ++u) ok = ok && (colb[jb + u] == colb[jb0 + u] + lane);
        }
      }
      if (lane == 0 && ok) {
        const int cd = colb[jb0 + UD];
        ok = (int64_t)cd == r0 + a && colb[jb0 + UD - 1] == cd - 1 && colb[jb0 + UD + 1] == cd + 1;
      }
      band = __all_sync(0xffffffffu, ok);
    }
    if (band) {
      // x[u] = X[col_u(row)] for the current row; the middle three slide, the others are loaded
      T x[SEGL][VEC];
      int cb[SEGL];  // column of entry u in the strip's first row
#pragma unroll
      for (int u = 0; u < SEGL; ++u) {
        cb[u] = colb[jb0 + u];
        ldx<T, VEC>(Xc, (int64_t)cb[u] * LD, x[u]);
      }
      const T* __restrict__ vrow = valb + jb0;
      int64_t off = coff + (int64_t)a * ld;
      for (int i = 0; i < ns; ++i, vrow += SEGL, off += ld) {
        T sum[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) sum[q] = T(0);
#pragma unroll
        for (int u = 0; u < SEGL; ++u) {
          const T av = vrow[u];
#pragma unroll
          for (int q = 0; q < VEC; ++q) sum[q] += av * x[u][q];
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q) sum[q] *= sv[q];
        if (FUSE_DOT) {
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[0][q] += (double)(x[UD][q] * sv[q]) * (double)sum[q];
        }
        if (i + 1 < ns) {
          // next row: slide the middle run, load the top of every run
#pragma unroll
          for (int q = 0; q < VEC; ++q) {
            x[UD - 1][q] = x[UD][q];
            x[UD][q] = x[UD + 1][q];
          }
#pragma unroll
          for (int u = 0; u < SEGL; ++u)
            if (u < UD - 1 || u > UD) ldx<T, VEC>(Xc, (int64_t)(cb[u] + i + 1) * LD, x[u]);
          if (p.pfd < 0) {
            // L1 prefetch of this thread's own segment of the X rows the strip loads -pfd rows
            // from now (every eighth lane starts a 128-byte line): the later LDG hits L1, so the
            // few gathers a warp keeps in flight no longer bound the kernel by their latency
            if ((lane & 7) == 0) {
#pragma unroll
              for (int u = 0; u < SEGL; ++u)
                if (u < UD - 1 || u > UD) {
                  int64_t pr = (int64_t)cb[u] + i + 1 - p.pfd;
                  pr = pr < n ? pr : n - 1;
                  asm volatile("prefetch.global.L1 [%0];" ::"l"(Xc + pr * ld));
                }
            }
          } else if (p.pfd) {
            // L2 prefetch of the X rows this strip loads `pfd` rows from now: one 128-byte line
            // per lane of the group's first lanes
            constexpr int lines = (LD * (int)sizeof(T) + 127) / 128;
            const int lig = threadIdx.x % tpr;
            if (lig < lines) {
#pragma unroll
              for (int u = 0; u < SEGL; ++u)
                if (u < UD - 1 || u > UD) {
                  int64_t pr = (int64_t)cb[u] + i + 1 + p.pfd;
                  pr = pr < n ? pr : n - 1;
                  asm volatile("prefetch.global.L2 [%0];" ::"l"(
                      reinterpret_cast<const char*>(X + pr * ld) + lig * 128));
                }
            }
          }
        }
        stw<T, VEC>(Wc, off, sum);
      }
    } else {
      // gather path, row by row (metadata from shared memory, or straight from global memory
      // for a chunk whose entries did not fit the staging buffer)
      for (int i = 0; i < ns; ++i) {
        const int lr = a + i;
        const int32_t jb = ptrb[lr], len = ptrb[lr + 1] - jb;
        const int32_t* cols = total <= kCap ? colb + (jb - base) : indices + jb;
        const T* vals = total <= kCap ? valb + (jb - base) : data + jb;
        const RowDot<VEC> d = strip_gather_row<T, VEC, LD, FUSE_DOT>(
            cols, vals, len, Xc, Wc, coff + (int64_t)lr * ld, s_sv + c0);
        if (FUSE_DOT) {
#pragma unroll
          for (int q = 0; q < VEC; ++q) acc[0][q] += d.d[q];
        }
      }
    }
    cp_async_wait<0>();  // row pointers of chunk t + 2 landed (visible after the next barrier)
  }
  cp_async_wait<0>();
  if (progress != nullptr) {
    // count my last chunk; the last CTA to leave re-arms the counters for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
      if (blockIdx.x < nchunks) atomicAdd(progress, 1u);
      __threadfence();
      const unsigned int left = atomicAdd(progress + 1, 1u);
      if (left == gridDim.x - 1) {
        progress[0] = 0u;
        progress[1] = 0u;
      }
    }
  }
  if (FUSE_DOT) cta_reduce_finalize<T, VEC, 1>(acc, ld, partial, 0, 1, fin);
}

// process-wide knobs (mf_spmm_config; initial values from the environment)
// Off by default: measured on BASELINE config 2 (profiles/r2b_spmm_sweep.jsonl, r2c_spmm_instep.jsonl)
// the band kernel moves 40 % fewer bytes through L1 but holds only 3 instead of 5 gathers per warp
// in flight (36 KB instead of 80 KB per SM at its register budget), and the product is bound by
// the latency of the gathers: 8.7 ms per launch inside the Lanczos step against 8.6 ms for the
// row-group kernel (8.5 against 6.7 ms stand-alone).  Results are bit-identical either way.
std::atomic<int> g_strip{env_int("MF_SPMM_STRIP", 0)};
std::atomic<int> g_rows{env_int("MF_SPMM_STRIP_ROWS", 64)};
std::atomic<int> g_pfd{env_int("MF_SPMM_STRIP_PFD", 2)};
std::atomic<int> g_minb{env_int("MF_SPMM_STRIP_MINB", 3)};

}  // namespace

void spmm_strip_config(int use_strip, int rows, int pfd, int minb) {
  if (use_strip >= 0) g_strip.store(use_strip);
  if (rows > 0) g_rows.store(rows);
  if (pfd > -100) g_pfd.store(pfd);  // > 0: L2 prefetch distance, < 0: L1 prefetch distance, 0: off
  if (minb > 0) g_minb.store(minb);
}

int32_t launch_spmm_strip(const int32_t* indptr, const int32_t* indices, const void* data,
                          int64_t n, int64_t nnz, int32_t dtype, const void* X, const void* s,
                          void* W, int64_t ld, const Reduce* red, unsigned int* progress,
                          cudaStream_t st, bool* taken, int64_t bandwidth) {
  *taken = false;
  const int env_strip = g_strip.load(std::memory_order_relaxed);
  const int env_rows = g_rows.load(std::memory_order_relaxed);
  const int env_pfd = g_pfd.load(std::memory_order_relaxed);
  const int env_minb = g_minb.load(std::memory_order_relaxed);
  static const int env_throttle = env_int("MF_SPMM_THROTTLE", 1);
  if (!env_strip || n <= 0) return MF_OK;
  const int nv = dtype == MF_F64 ? 2 : 4;
  if (ld < 32 * nv) return MF_OK;  // a row-group must be whole warps
  const double avg = (double)nnz / (double)n;
  const int segl = avg <= 5.0 ? 5 : (avg <= 7.0 ? 7 : 0);
  const bool aligned = ((uintptr_t)indices % 16 == 0) && ((uintptr_t)data % 16 == 0);
  if (segl == 0 || !aligned || n >= (1ll << 31) - 1 || nnz >= (1ll << 31) - 1) return MF_OK;
  const int rps = kBlock / (int)(ld / nv);
  int64_t R = env_rows;
  const int64_t fit = (int64_t)(kCap / (avg > 1.0 ? avg : 1.0));
  if (R > fit) R = fit;
  if (R > 32 * rps) R = 32 * rps;  // one lane per strip row in the detection
  if (R > kMaxRows) R = kMaxRows;
  R = R / rps * rps;
  if (R < rps) R = rps;
  const int64_t nchunks = (n + R - 1) / R;
  Finalize fin{};
  double* partial = nullptr;
  if (red) {
    fin = red->fin;
    partial = red->partial;
  }
  SpmmParams prm{(int)ld, (int)R, 0, 0, 0, env_pfd, 0, 0, 0, 0};
  choose_row_order(&prm, n, avg, bandwidth, ld, dtype);
  unsigned int* prog = env_throttle ? progress : nullptr;
  *taken = true;  // timed under the caller's MF_KC_SPMM_CSR scope
#define MF_STRIP_L(T, VEC, LD, SEGL, DOT, MINB)                                                  \
  do {                                                                                           \
    auto kern = spmm_strip_kernel<T, VEC, LD, SEGL, DOT, MINB>;                                  \
    const int grid = resident_grid((const void*)kern, kBlock, 0, nchunks);                       \
    prm.window = grid + (grid / 4 > 8 ? grid / 4 : 8);                                           \
    kern<<<grid, kBlock, 0, st>>>(indptr, indices, (const T*)data, n, nnz, (const T*)X,          \
                                  (const T*)s, (T*)W, prm, prog, partial, fin);                  \
    return check_launch("spmm_strip");                                                           \
  } while (0)
#define MF_STRIP_D(T, VEC, LD, SEGL, MINB)                                                       \
  do {                                                                                           \
    if (red) MF_STRIP_L(T, VEC, LD, SEGL, true, MINB);                                           \
    else MF_STRIP_L(T, VEC, LD, SEGL, false, MINB);                                              \
  } while (0)
#define MF_STRIP_S(T, VEC, LD)                                                                   \
  do {                                                                                           \
    if (segl == 5) {                                                                             \
      if (env_minb >= 4) MF_STRIP_D(T, VEC, LD, 5, 4);                                           \
      else MF_STRIP_D(T, VEC, LD, 5, 3);                                                         \
    } else {                                                                                     \
      MF_STRIP_D(T, VEC, LD, 7, 3);                                                              \
    }                                                                                            \
  } while (0)
  if (dtype == MF_F32) {
    if (ld == 256) MF_STRIP_S(float, 4, 256);
    if (ld == 128) MF_STRIP_S(float, 4, 128);
  } else {
    if (ld == 256) MF_STRIP_S(double, 2, 256);
    if (ld == 128) MF_STRIP_S(double, 2, 128);
    if (ld == 64) MF_STRIP_S(double, 2, 64);
  }
#undef MF_STRIP_S
#undef MF_STRIP_D
#undef MF_STRIP_L
  *taken = false;
  return MF_OK;
}

}  // namespace mf
