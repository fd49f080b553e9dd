// K2s -- CSR sparse matrix times probe block, W[n][ld] = s * (A @ X[n][ld]),
// with the Lanczos alpha (column sums of (X*s) .* W, matfree/decomp.py:288)
// reduced in the same kernel (last CTA writes the scalars).
//
// The user matvec of the reference (matfree/stochtrace.py:47-49) has no sparse
// implementation; under vmap XLA would gather per probe.  Here all probes of a
// tile advance together: a group of ld/VEC threads owns one row, every
// non-zero costs one 16-byte copy per thread of the contiguous segment
// X[col][c0..c0+VEC), i.e. ld*sizeof(T) contiguous bytes per non-zero per row
// (1 KB at ld = 256).
//
// Structure (measured on B200, profiles/: after the fixes below the kernel is bound by the
// latency of the gathers -- not by issue, L2 or DRAM bandwidth):
//   * one wave of persistent CTAs; a CTA takes chunks of R consecutive rows.
//     Chunk c goes to CTA c % grid (static => the per-CTA reduction order is
//     fixed => bit-reproducible results); a completed-chunk counter keeps every
//     CTA within one window of the slowest, so the rows in flight are always one
//     contiguous window and the stencil neighbours of a row stay in L2 between
//     their uses (measured: DRAM reads drop from 1.6x to 1.0x of the algorithmic
//     bytes; without it the CTAs drift apart by more than the L2 can hold);
//   * the CSR metadata of the NEXT chunks (row pointers two chunks ahead, column
//     indices / values one chunk ahead) streams into multi-buffered shared
//     memory with cp.async while the current chunk is processed, so the only
//     global loads on the per-row critical path are the gathers of X themselves;
//   * per row all gathers (a whole 5- or 7-point stencil row) are issued
//     back to back into registers before the first FMA, one IMAD.WIDE + one
//     LDG.128 per non-zero (tile width is a template constant).
// Variants that were measured and rejected (profiles/README.md, "SpMM tuning log"):
// dynamic chunk tickets (same speed, not reproducible), a cp.async ring in shared
// memory as the landing zone (2x the instructions, fewer warps: slower), register
// software pipelining across rows (spills), L1/L2 software prefetch (no gain).
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "internal.h"
#include "spmm_common.cuh"

namespace mf {
namespace {

// Irregular matrices (power-law graphs, BASELINE config 5): rows longer than kLongRow are not
// multiplied by the row-group that meets them; they are cut into segments of kSegNnz non-zeros
// and appended to a work list that spmm_long_rows_kernel (one CTA per segment) drains afterwards.
constexpr int kLongRow = kSpmmLongRow;
constexpr int kSegNnz = kSpmmSegNnz;
struct LongSeg {
  int32_t row;       // global row
  int32_t j0, j1;    // non-zero range of this segment
  int32_t nseg;      // segments of this row
  int32_t seg;       // index of this segment within the row
  int32_t slot0;     // first partial-sum slot of the row (multi-segment rows), else -1
  int32_t head;      // list index of the row's first segment (its `ticket` counts arrivals)
  unsigned int ticket;
};
struct LongList {
  LongSeg* segs;          // capacity max_segs
  unsigned int* counts;   // [0] segments, [1] slots   (zeroed before the launch)
  int64_t max_segs, max_slots;
};
static_assert(sizeof(LongSeg) == 32, "scratch sizing assumes 32-byte list entries");

// The gathers of one row held in registers: up to SEGL non-zeros (a whole stencil row);
// longer rows finish in a serial tail.
template <typename T, int VEC, int SEGL>
struct RowRegs {
  T x[SEGL][VEC];
  int jb, len;
  unsigned int dmask;  // bit u set: entry u (< SEGL) is the diagonal; filled when `rowg` >= 0
};

template <typename T, int VEC, int SEGL>
__device__ __forceinline__ void issue_row(const int32_t* __restrict__ ptrb, int32_t base,
                                          const int32_t* __restrict__ colb,
                                          const T* __restrict__ Xc, int ld, int lr,
                                          RowRegs<T, VEC, SEGL>& r, int32_t rowg = -1) {
  r.jb = ptrb[lr] - base;
  r.len = ptrb[lr + 1] - base - r.jb;
  int32_t c[SEGL];
  r.dmask = 0u;
#pragma unroll
  for (int u = 0; u < SEGL; ++u) {
    c[u] = u < r.len ? colb[r.jb + u] : -2;
    r.dmask |= (unsigned int)(c[u] == rowg) << u;
  }
#pragma unroll
  for (int u = 0; u < SEGL; ++u)
    if (u < r.len) ldx<T, VEC>(Xc, (int64_t)c[u] * ld, r.x[u]);
}

template <typename T, int VEC, int SEGL>
__device__ __forceinline__ void prefetch_row_l1(const int32_t* __restrict__ ptrb, int32_t base,
                                                const int32_t* __restrict__ colb,
                                                const T* __restrict__ Xc, int ld, int lr) {
  const int jb = ptrb[lr] - base;
  const int len = ptrb[lr + 1] - base - jb;
#pragma unroll
  for (int u = 0; u < SEGL; ++u)
    if (u < len)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(Xc + (int64_t)colb[jb + u] * ld));
}

// LD > 0: the tile width is a compile-time constant (address arithmetic folds into one
// IMAD.WIDE per gather); LD == 0: any power of two, read from p.ld.
// PIPE: the gathers of the next row are issued before the current row is multiplied.
template <typename T, int VEC, int LD, int SEGL, bool PIPE, bool FUSE_DOT, bool DEFER = false,
          bool BLOCKED = false>
__global__ void __launch_bounds__(kBlock, (PIPE || DEFER || (SEGL == 7 && LD >= 128)) ? 3 : 4)
spmm_csr_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                const T* __restrict__ data, int64_t n, const T* __restrict__ X,
                const T* __restrict__ s, T* __restrict__ W, SpmmParams p,
                unsigned int* __restrict__ progress, double* __restrict__ partial,
                Finalize fin, LongList longs) {
  __shared__ int32_t s_ptr[3][kMaxRows + 1];
  __shared__ int32_t s_col[2][kCap];
  __shared__ T s_val[2][kCap];

  const int ld = LD > 0 ? LD : p.ld;
  const int tpr = ld / VEC;      // threads per row
  const int rps = kBlock / tpr;  // rows per sweep of the CTA
  const int my_row = threadIdx.x / tpr;
  const int c0 = (threadIdx.x % tpr) * VEC;
  const T* __restrict__ Xc = X + c0;
  T* __restrict__ Wc = W + c0;
  const int R = p.rows_per_chunk;
  T sv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) sv[i] = s ? s[c0 + i] : T(1);
  double acc[1][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[0][i] = 0.0;

  const int64_t nchunks = (n + R - 1) / R;
  const int64_t G = gridDim.x;

  // first row of chunk c: ascending order, or the blocked order of SpmmParams (3-D stencils
  // whose planes exceed the L2; a separate instantiation so the common path pays nothing for it)
  auto row0_of = [&](int64_t c) -> int64_t {
    if constexpr (BLOCKED) return chunk_row0(c, p);
    else return c * (int64_t)R;
  };
  // first row of this CTA's next chunk (>= n if there is none)
  auto next_row0 = [&](int64_t c, int64_t r0c) -> int64_t {
    if constexpr (BLOCKED) return c + G < nchunks ? chunk_row0(c + G, p) : n;
    else return r0c + G * R;
  };
  // ---- metadata pipeline (all threads): cp.async into multi-buffered shared memory --------
  auto issue_ptr = [&](int64_t c, int buf) {  // row pointers of chunk c -> s_ptr[buf]
    if (c < nchunks) {
      const int64_t r0 = row0_of(c);
      const int nr = (int)((n - r0) < R ? (n - r0) : R);
      for (int i = threadIdx.x; i <= nr; i += kBlock)
        cp_async<4>((uint32_t)__cvta_generic_to_shared(&s_ptr[buf][i]), indptr + r0 + i);
    }
  };
  auto issue_ent = [&](int64_t c, int pbuf, int ebuf) {  // needs s_ptr[pbuf] visible
    if (c < nchunks) {
      const int64_t r0 = row0_of(c);
      const int nr = (int)((n - r0) < R ? (n - r0) : R);
      const int32_t base = s_ptr[pbuf][0];
      const int total = s_ptr[pbuf][nr] - base;
      if (total <= kCap) {  // larger chunks read their entries straight from global memory
        for (int i = threadIdx.x; i < total; i += kBlock) {
          cp_async<4>((uint32_t)__cvta_generic_to_shared(&s_col[ebuf][i]), indices + base + i);
          cp_async<(int)sizeof(T)>((uint32_t)__cvta_generic_to_shared(&s_val[ebuf][i]),
                                   data + base + i);
        }
      }
    }
  };

  // finish one row: scale, store, fused dot
  // The fused dot needs X[row]: for the 7-point rows (80-register kernel) it is taken from the
  // gather of the diagonal entry when the row has one, instead of a reload through L1.
  // (measured on the 3-D workload: 14.5 instead of 13.1 ms per product -- the extra live registers
  // spill in the 80-register kernel -- so it is compiled out; profiles/r2j_instep.jsonl)
  constexpr bool kDiagFromGather = false;
  auto finish_row_x = [&](T (&sum)[VEC], int64_t off, T (&xo)[VEC], bool have_x) {
    T w[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) w[i] = sum[i] * sv[i];
    stw<T, VEC>(Wc, off, w);
    if (FUSE_DOT) {
      if (!have_x) ldx<T, VEC>(Xc, off, xo);
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[0][i] += (double)(xo[i] * sv[i]) * (double)w[i];
    }
  };
  auto finish_row = [&](T (&sum)[VEC], int64_t off) {
    T xo[VEC];
    finish_row_x(sum, off, xo, false);
  };

  // hand a long row to the segment list (one thread of the row-group)
  auto defer_row = [&](int64_t row, int32_t jb, int32_t je) {
    if ((threadIdx.x % tpr) != 0) return;
    const int len = je - jb;
    const int nseg = (len + kSegNnz - 1) / kSegNnz;
    const unsigned int first = atomicAdd(longs.counts, (unsigned int)nseg);
    int slot0 = -1;
    if (nseg > 1) slot0 = (int)atomicAdd(longs.counts + 1, (unsigned int)nseg);
    if ((int64_t)first + nseg > longs.max_segs || (nseg > 1 && (int64_t)slot0 + nseg > longs.max_slots))
      __trap();  // cannot happen: capacities are upper bounds derived from nnz
    for (int q = 0; q < nseg; ++q) {
      LongSeg e;
      e.row = (int32_t)row;
      e.j0 = jb + q * kSegNnz;
      e.j1 = e.j0 + kSegNnz < je ? e.j0 + kSegNnz : je;
      e.nseg = nseg;
      e.seg = q;
      e.slot0 = slot0;
      e.head = (int32_t)first;
      e.ticket = 0u;
      longs.segs[first + q] = e;
    }
  };

  int64_t ch = blockIdx.x;
  issue_ptr(ch, 0);
  issue_ptr(ch + G, 1);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  issue_ent(ch, 0, 0);
  cp_async_commit();
  cp_async_wait<0>();

  unsigned int seen_done = 0;  // thread 0: last value read from progress[0]
  for (int64_t t = 0; ch < nchunks; ++t, ch += G) {
    const int pb = (int)(t % 3), eb = (int)(t & 1);
    // Throttle: do not run more than `window` chunks ahead of the completed count.
    if (progress != nullptr && threadIdx.x == 0) {
      while ((int64_t)seen_done + p.window <= ch) {
        seen_done = *reinterpret_cast<volatile unsigned int*>(progress);
        if ((int64_t)seen_done + p.window <= ch) __nanosleep(200);
      }
    }
    __syncthreads();  // metadata of chunk t visible; everyone is done with chunk t-1
    if (progress != nullptr && threadIdx.x == 0 && t > 0) atomicAdd(progress, 1u);
    // stream in the metadata of the next chunks while this one is processed
    issue_ent(ch + G, (int)((t + 1) % 3), (int)((t + 1) & 1));
    issue_ptr(ch + 2 * G, (int)((t + 2) % 3));
    cp_async_commit();

    const int64_t r0 = row0_of(ch);
    const int nr = (int)((n - r0) < R ? (n - r0) : R);
    if (p.prefetch) {
      const int64_t pr0 = next_row0(ch, r0);
      if (pr0 < n) {
        const int64_t pnr = (n - pr0) < R ? (n - pr0) : R;
        const char* pbase = reinterpret_cast<const char*>(X + pr0 * ld);
        const int64_t pbytes = pnr * ld * (int64_t)sizeof(T);
        for (int64_t o = (int64_t)threadIdx.x * 128; o < pbytes; o += (int64_t)kBlock * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pbase + o));
      }
    }
    const int32_t* __restrict__ ptrb = s_ptr[pb];
    const int32_t base = ptrb[0];
    const int total = ptrb[nr] - base;
    const int32_t* __restrict__ colb = s_col[eb];
    const T* __restrict__ valb = s_val[eb];
    const int64_t coff = r0 * ld;

    if (total > kCap) {
      // oversized chunk (very long rows): plain serial loop over global metadata
      for (int lr = my_row; lr < nr; lr += rps) {
        const int32_t jb = ptrb[lr], je = ptrb[lr + 1];
        if (DEFER && je - jb > kLongRow) {
          defer_row(r0 + lr, jb, je);
          continue;
        }
        T sum[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) sum[i] = T(0);
        int32_t j = jb;
        if constexpr (DEFER) {
          for (; j + 8 <= je; j += 8) {  // eight independent gathers in flight
            int32_t c[8];
            T a[8], x[8][VEC];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              c[u] = __ldg(indices + j + u);
              a[u] = __ldg(data + j + u);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) ldx<T, VEC>(Xc, (int64_t)c[u] * ld, x[u]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
              for (int i = 0; i < VEC; ++i) sum[i] += a[u] * x[u][i];
            }
          }
        }
        for (; j < je; ++j) {
          const int32_t c = __ldg(indices + j);
          const T a = __ldg(data + j);
          T x[VEC];
          ldx<T, VEC>(Xc, (int64_t)c * ld, x);
#pragma unroll
          for (int i = 0; i < VEC; ++i) sum[i] += a * x[i];
        }
        finish_row(sum, coff + (int64_t)lr * ld);
      }
    } else {
      // multiply the gathers held in r (and the serial tail of a long row)
      auto consume = [&](RowRegs<T, VEC, SEGL>& r, int lr) {
        T sum[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) sum[i] = T(0);
#pragma unroll
        for (int u = 0; u < SEGL; ++u) {
          if (u < r.len) {
            const T a = valb[r.jb + u];
#pragma unroll
            for (int i = 0; i < VEC; ++i) sum[i] += a * r.x[u][i];
          }
        }
        int u = SEGL;
        if constexpr (DEFER) {
          for (; u + 8 <= r.len; u += 8) {  // long tail: eight independent gathers in flight
            T x[8][VEC];
#pragma unroll
            for (int q = 0; q < 8; ++q) ldx<T, VEC>(Xc, (int64_t)colb[r.jb + u + q] * ld, x[q]);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const T a = valb[r.jb + u + q];
#pragma unroll
              for (int i = 0; i < VEC; ++i) sum[i] += a * x[q][i];
            }
          }
        }
        for (; u < r.len; ++u) {
          const T a = valb[r.jb + u];
          T x[VEC];
          ldx<T, VEC>(Xc, (int64_t)colb[r.jb + u] * ld, x);
#pragma unroll
          for (int i = 0; i < VEC; ++i) sum[i] += a * x[i];
        }
        if constexpr (kDiagFromGather) {
          T xo[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) xo[i] = T(0);
#pragma unroll
          for (int u = 0; u < SEGL; ++u)
            if ((r.dmask >> u) & 1u) {  // independent predicates: stays in registers
#pragma unroll
              for (int i = 0; i < VEC; ++i) xo[i] = r.x[u][i];
            }
          finish_row_x(sum, coff + (int64_t)lr * ld, xo, r.dmask != 0u);
        } else {
          finish_row(sum, coff + (int64_t)lr * ld);
        }
      };
      if (PIPE) {
        RowRegs<T, VEC, SEGL> ra, rb;
        int lr = my_row;
        if (lr < nr) issue_row<T, VEC, SEGL>(ptrb, base, colb, Xc, ld, lr, ra);
        while (lr < nr) {
          const int lr1 = lr + rps;
          if (lr1 < nr) issue_row<T, VEC, SEGL>(ptrb, base, colb, Xc, ld, lr1, rb);
          consume(ra, lr);
          if (lr1 >= nr) break;
          const int lr2 = lr1 + rps;
          if (lr2 < nr) issue_row<T, VEC, SEGL>(ptrb, base, colb, Xc, ld, lr2, ra);
          consume(rb, lr1);
          lr = lr2;
        }
      } else {
        for (int lr = my_row; lr < nr; lr += rps) {
          if (DEFER && ptrb[lr + 1] - ptrb[lr] > kLongRow) {
            defer_row(r0 + lr, ptrb[lr], ptrb[lr + 1]);
            continue;
          }
          RowRegs<T, VEC, SEGL> ra;
          issue_row<T, VEC, SEGL>(ptrb, base, colb, Xc, ld, lr, ra,
                                  kDiagFromGather ? (int32_t)(r0 + lr) : -1);
          if (p.pfd) {
            // Short-distance L2 prefetch of the X row this row-group owns `pfd` sweeps from now
            // (possibly in the CTA's next chunk).  All CTAs advance through one contiguous window
            // in step, and for a banded matrix the rows a sweep gathers are the rows other CTAs
            // own at the same sweep: without this every first touch of an X row -- and the up to
            // nnz/row requesters that pile onto it -- waits a full DRAM latency with only one row
            // per warp in flight.  One prefetch per 128-byte line by the first lanes of the group.
            const int lane_in_row = threadIdx.x % tpr;
            const int lines = (ld * (int)sizeof(T) + 127) / 128;
            if (lane_in_row < lines) {
              const int lp = lr + p.pfd * rps;
              const int64_t prow = lp < R ? r0 + lp : next_row0(ch, r0) + (lp - R);
              if (prow < n)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(
                    reinterpret_cast<const char*>(X + prow * ld) + lane_in_row * 128));
            }
          }
          if (p.l1pf) {
            const int lp = lr + p.l1pf * rps;
            if (lp < nr) prefetch_row_l1<T, VEC, SEGL>(ptrb, base, colb, Xc, ld, lp);
          }
          consume(ra, lr);
        }
      }
    }
    cp_async_wait<0>();  // next chunk's metadata landed
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

// One CTA per segment of a long row: the CTA's row-groups take the segment's non-zeros
// round-robin (eight gathers in flight each), their sums are added in group order in shared
// memory; a multi-segment row parks its segment sums in `slots` and the CTA that completes the
// row (ticket on the row's first list entry) adds them in segment order -> deterministic.
template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock)
spmm_long_rows_kernel(const int32_t* __restrict__ indices, const T* __restrict__ data,
                      const T* __restrict__ X, const T* __restrict__ s, T* __restrict__ W, int ld,
                      LongList longs, T* __restrict__ slots) {
  __shared__ T red[kBlock * VEC];
  __shared__ bool last;
  const int tpr = ld / VEC;
  const int rps = kBlock / tpr;
  const int g = threadIdx.x / tpr;
  const int c0 = (threadIdx.x % tpr) * VEC;
  const T* __restrict__ Xc = X + c0;
  const unsigned int nsegs = *reinterpret_cast<volatile unsigned int*>(longs.counts);
  for (unsigned int e = blockIdx.x; e < nsegs; e += gridDim.x) {
    const LongSeg sg = longs.segs[e];
    T sum[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) sum[i] = T(0);
    int32_t j = sg.j0 + g;
    for (; j + 7 * rps < sg.j1; j += 8 * rps) {
      int32_t c[8];
      T a[8], x[8][VEC];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        c[u] = __ldg(indices + j + u * rps);
        a[u] = __ldg(data + j + u * rps);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) ldx<T, VEC>(Xc, (int64_t)c[u] * ld, x[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) sum[i] += a[u] * x[u][i];
      }
    }
    for (; j < sg.j1; j += rps) {
      const int32_t c = __ldg(indices + j);
      const T a = __ldg(data + j);
      T x[VEC];
      ldx<T, VEC>(Xc, (int64_t)c * ld, x);
#pragma unroll
      for (int i = 0; i < VEC; ++i) sum[i] += a * x[i];
    }
    __syncthreads();  // previous iteration's readers of `red` are done
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[threadIdx.x * VEC + i] = sum[i];
    __syncthreads();
    // group 0 adds the groups in order: element (g, c0 + i) sits at red[(g * tpr + c0 / VEC) * VEC + i]
    if (g == 0) {
      for (int q = 1; q < rps; ++q) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) sum[i] += red[(q * tpr + threadIdx.x) * VEC + i];
      }
      if (sg.nseg == 1) {
        T w[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) w[i] = sum[i] * (s ? s[c0 + i] : T(1));
        stw<T, VEC>(W + c0, (int64_t)sg.row * ld, w);
      } else {
        T* slot = slots + (int64_t)(sg.slot0 + sg.seg) * ld + c0;
#pragma unroll
        for (int i = 0; i < VEC; ++i) slot[i] = sum[i];
      }
    }
    if (sg.nseg > 1) {
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0)
        last = atomicAdd(&longs.segs[sg.head].ticket, 1u) == (unsigned int)sg.nseg - 1;
      __syncthreads();
      if (last && g == 0) {
        __threadfence();
        T tot[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) tot[i] = T(0);
        for (int q = 0; q < sg.nseg; ++q) {
          const T* slot = slots + (int64_t)(sg.slot0 + q) * ld + c0;
#pragma unroll
          for (int i = 0; i < VEC; ++i) tot[i] += __ldcg(slot + i);
        }
        T w[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) w[i] = tot[i] * (s ? s[c0 + i] : T(1));
        stw<T, VEC>(W + c0, (int64_t)sg.row * ld, w);
      }
    }
  }
}

}  // namespace

int resident_grid(const void* kernel, int block, size_t smem, int64_t want) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> cache;
  int per_sm = 0;
  if (smem == 0) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(kernel);
    if (it != cache.end()) per_sm = it->second;
  }
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess ||
        per_sm < 1) {
      cudaGetLastError();
      per_sm = 1;
    }
    if (smem == 0) {  // dynamic-smem kernels change occupancy per call: do not cache
      std::lock_guard<std::mutex> lk(mu);
      cache[kernel] = per_sm;
    }
  }
  int64_t g = (int64_t)per_sm * num_sms();
  if (g > kMaxPartialCtas) g = kMaxPartialCtas;
  if (g > want) g = want;
  if (g < 1) g = 1;
  return (int)g;
}

int64_t spmm_irregular_scratch_bytes(int64_t nnz, int64_t ld, int32_t dtype) {
  const int64_t max_segs = nnz / kSegNnz + nnz / kLongRow + 2;
  const int64_t max_slots = 2 * (nnz / kSegNnz) + 2;
  return 256 + align_up(max_segs * (int64_t)sizeof(LongSeg), 256) +
         max_slots * ld * (int64_t)dtype_size(dtype) + 256;
}

namespace {
// Irregular route: the row-group kernel with long rows deferred, then the segment kernel.
template <typename T, int VEC>
int32_t launch_irregular(const int32_t* indptr, const int32_t* indices, const T* data, int64_t n,
                         int64_t nnz, const T* X, const T* s, T* W, int64_t ld, void* scratch,
                         cudaStream_t st) {
  char* base = (char*)scratch;
  LongList longs;
  longs.counts = (unsigned int*)base;
  longs.max_segs = nnz / kSegNnz + nnz / kLongRow + 2;
  longs.max_slots = 2 * (nnz / kSegNnz) + 2;
  longs.segs = (LongSeg*)(base + 256);
  T* slots = (T*)(base + 256 + align_up(longs.max_segs * (int64_t)sizeof(LongSeg), 256));
  if (cudaMemsetAsync(longs.counts, 0, 256, st) != cudaSuccess) {
    set_error("spmm: counter memset failed");
    return MF_ERR_CUDA;
  }
  const int rps = kBlock / (int)(ld / VEC);
  // chunks sized for the average row; chunks that still overflow the staging buffer take the
  // direct-from-global path of the kernel
  const double avg = (double)nnz / (double)n;
  int64_t R = 64;
  const int64_t fit = (int64_t)(kCap / (avg > 1.0 ? avg : 1.0));
  if (R > fit) R = fit;
  R = R / rps * rps;
  if (R < rps) R = rps;
  const int64_t nchunks = (n + R - 1) / R;
  SpmmParams prm{(int)ld, (int)R, 0, 0, 0, 0, 0, 0, 0, 0};
  {
    auto kern = spmm_csr_kernel<T, VEC, 0, 8, false, false, true>;
    const int grid = resident_grid((const void*)kern, kBlock, 0, nchunks);
    kern<<<grid, kBlock, 0, st>>>(indptr, indices, data, n, X, s, W, prm, nullptr, nullptr,
                                  Finalize{}, longs);
    MF_TRY(check_launch("spmm_csr (irregular)"));
  }
  {
    auto kern = spmm_long_rows_kernel<T, VEC>;
    const int grid = resident_grid((const void*)kern, kBlock, 0, 1 << 20);
    kern<<<grid, kBlock, 0, st>>>(indices, data, X, s, W, (int)ld, longs, slots);
  }
  return check_launch("spmm_long_rows");
}
// One thread per row, metadata straight from global memory, no staging and no barriers: the route
// for a single vector (ld = 1: BASELINE config 4, where the product is bound by streaming the
// matrix).  The staged row-group kernel above moves 14 KB of metadata per 256-row chunk between
// two CTA barriers and is latency-bound there (2.5-3.1 TB/s of matrix stream on the 256^3 7-point
// Laplacian); here every thread keeps its row's (column, value) pairs and X gathers in flight at
// once, 2048 threads per SM.  Rows of up to 8 entries are fully unrolled; the FMA order is the CSR
// order, so the results are those of the other kernels bit for bit.
template <typename T>
__global__ void __launch_bounds__(256)
spmm_row_thread_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const T* __restrict__ data, int64_t n, const T* __restrict__ X,
                       const T* __restrict__ s, T* __restrict__ W) {
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  const T sc = s ? s[0] : T(1);
  const int32_t jb = __ldg(indptr + r), je = __ldg(indptr + r + 1);
  const int len = je - jb;
  T sum = T(0);
  if (len <= 8) {
    int32_t c[8];
    T a[8], x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      c[u] = 0;
      a[u] = T(0);
      if (u < len) {
        c[u] = __ldg(indices + jb + u);
        a[u] = __ldg(data + jb + u);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = u < len ? __ldg(X + c[u]) : T(0);
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (u < len) sum += a[u] * x[u];
  } else {
    for (int32_t j = jb; j < je; ++j) sum += __ldg(data + j) * __ldg(X + __ldg(indices + j));
  }
  W[r] = sum * sc;
}

}  // namespace

int32_t launch_spmm_csr(const int32_t* indptr, const int32_t* indices, const void* data,
                        int64_t n, int64_t nnz, int32_t dtype, const void* X, const void* s,
                        void* W, int64_t ld, const Reduce* red, unsigned int* progress,
                        cudaStream_t st, void* irregular_scratch, int64_t bandwidth,
                        int32_t num_diagonals, int64_t line_stride) {
  MF_KSCOPE(MF_KC_SPMM_CSR, st);
  if (n <= 0) return MF_OK;
  if (irregular_scratch != nullptr) {
    if (red != nullptr) {
      set_error("spmm: the irregular route does not fuse the dot product");
      return MF_ERR_INVALID_ARGUMENT;
    }
    if (dtype == MF_F32) {
      if (ld >= 4)
        return launch_irregular<float, 4>(indptr, indices, (const float*)data, n, nnz,
                                          (const float*)X, (const float*)s, (float*)W, ld,
                                          irregular_scratch, st);
      return launch_irregular<float, 1>(indptr, indices, (const float*)data, n, nnz,
                                        (const float*)X, (const float*)s, (float*)W, ld,
                                        irregular_scratch, st);
    }
    if (ld >= 2)
      return launch_irregular<double, 2>(indptr, indices, (const double*)data, n, nnz,
                                         (const double*)X, (const double*)s, (double*)W, ld,
                                         irregular_scratch, st);
    return launch_irregular<double, 1>(indptr, indices, (const double*)data, n, nnz,
                                       (const double*)X, (const double*)s, (double*)W, ld,
                                       irregular_scratch, st);
  }
  if (ld == 1 && red == nullptr && n < (1ll << 31) * 256) {
    // a single vector without the fused dot: one thread per row (see the kernel)
    static const int env_row_thread = env_int("MF_SPMM_ROW_THREAD", 1);
    if (env_row_thread) {
      const unsigned int grid = (unsigned int)((n + 255) / 256);
      if (dtype == MF_F32)
        spmm_row_thread_kernel<float><<<grid, 256, 0, st>>>(indptr, indices, (const float*)data, n,
                                                           (const float*)X, (const float*)s, (float*)W);
      else
        spmm_row_thread_kernel<double><<<grid, 256, 0, st>>>(indptr, indices, (const double*)data, n,
                                                            (const double*)X, (const double*)s, (double*)W);
      return check_launch("spmm_row_thread");
    }
  }
  const int nv = dtype == MF_F64 ? 2 : 4;
  const int vec = ld >= nv ? nv : 1;
  const int rps = kBlock / (int)(ld / vec);
  {
    // 5-diagonal band matrices on the 256-wide fp32 tile: X rows staged by TMA (spmm_tma.cu)
    bool taken = false;
    const int32_t rc = launch_spmm_tma(indptr, indices, data, n, nnz, dtype, X, s, W, ld, red, progress,
                                       st, &taken, bandwidth, num_diagonals, line_stride);
    if (taken) return rc;
  }
  {
    // banded / stencil matrices on wide tiles: the strip kernel (spmm_strip.cu)
    bool taken = false;
    const int32_t rc = launch_spmm_strip(indptr, indices, data, n, nnz, dtype, X, s, W, ld, red,
                                         progress, st, &taken, bandwidth);
    if (taken) return rc;
  }
  static const int env_rows = env_int("MF_SPMM_ROWS", 0);
  static const int env_prefetch = env_int("MF_SPMM_PREFETCH", 0);
  static const int env_l1pf = env_int("MF_SPMM_L1PF", 0);
  static const int env_throttle = env_int("MF_SPMM_THROTTLE", 1);
  static const int env_pipe = env_int("MF_SPMM_PIPE", 0);
  static const int env_slack = env_int("MF_SPMM_SLACK", 0);
  if (!env_throttle) progress = nullptr;
  const double avg = n > 0 ? (double)nnz / (double)n : 1.0;
  // rows per chunk: a multiple of the CTA sweep, sized so the chunk's non-zeros fit the
  // staging buffer on average, at most kMaxRows.
  int64_t R = env_rows > 0 ? env_rows : 64;
  const int64_t fit = (int64_t)(kCap / (avg > 1.0 ? avg : 1.0));
  if (R > fit) R = fit;
  if (R > kMaxRows) R = kMaxRows;
  R = R / rps * rps;
  if (R < rps) R = rps;  // rps <= 256 == kMaxRows
  const int64_t nchunks = (n + R - 1) / R;
  Finalize fin{};
  double* partial = nullptr;
  if (red) {
    fin = red->fin;
    partial = red->partial;
  }
  // measured on C2 (tools/bench_spmm.py, profiles/r1j_spmm_sweep.jsonl): 7.62 ms without,
  // 7.40 / 7.18 / 7.11 ms at 1 / 2 / 3 sweeps ahead, worse from 4 on (the prefetched rows then
  // leave the window the L2 holds)
  static const int env_pfd = env_int("MF_SPMM_PFD", 3);
  SpmmParams prm{(int)ld, (int)R, env_prefetch, env_l1pf, 0, env_pfd, 0, 0, 0, 0};
  choose_row_order(&prm, n, avg, bandwidth, ld, dtype);
#define MF_SPMM_K(T, VEC, LD, SEGL, PIPE, DOT, BLK)                                            \
  do {                                                                                         \
    auto kern = spmm_csr_kernel<T, VEC, LD, SEGL, PIPE, DOT, false, BLK>;                      \
    const int grid = resident_grid((const void*)kern, kBlock, 0, nchunks);                     \
    prm.window = grid + (env_slack > 0 ? env_slack : (grid / 4 > 8 ? grid / 4 : 8));           \
    kern<<<grid, kBlock, 0, st>>>(indptr, indices, (const T*)data, n, (const T*)X,             \
                                  (const T*)s, (T*)W, prm, progress, partial, fin, LongList{}); \
  } while (0)
  // the blocked row order exists for the wide tiles only (compile-time LD): elsewhere a plane of
  // X is small enough for L2 anyway
#define MF_SPMM_L(T, VEC, LD, SEGL, PIPE, DOT)                                                 \
  do {                                                                                         \
    if constexpr ((LD) > 0 && !(PIPE)) {                                                       \
      if (prm.block_rows != 0) MF_SPMM_K(T, VEC, LD, SEGL, PIPE, DOT, true);                   \
      else MF_SPMM_K(T, VEC, LD, SEGL, PIPE, DOT, false);                                      \
    } else {                                                                                   \
      prm.block_rows = 0;                                                                      \
      MF_SPMM_K(T, VEC, LD, SEGL, PIPE, DOT, false);                                           \
    }                                                                                          \
  } while (0)
#define MF_SPMM_D(T, VEC, LD, SEGL, PIPE)                     \
  do {                                                        \
    if (red) MF_SPMM_L(T, VEC, LD, SEGL, PIPE, true);         \
    else MF_SPMM_L(T, VEC, LD, SEGL, PIPE, false);            \
  } while (0)
  // segment length: a whole row of a 5-point (2-D) or 7-point (3-D) stencil, else 8
  const int segl = avg <= 5.0 ? 5 : (avg <= 7.0 ? 7 : 8);
#define MF_SPMM_S(T, VEC, LD)                                           \
  do {                                                                  \
    if (segl == 5) {                                                    \
      if (env_pipe) MF_SPMM_D(T, VEC, LD, 5, true);                     \
      else MF_SPMM_D(T, VEC, LD, 5, false);                             \
    } else if (segl == 7) {                                             \
      MF_SPMM_D(T, VEC, LD, 7, false);                                  \
    } else {                                                            \
      MF_SPMM_D(T, VEC, LD, 8, false);                                  \
    }                                                                   \
  } while (0)
  if (dtype == MF_F32) {
    if (ld == 256) MF_SPMM_S(float, 4, 256);
    else if (ld == 128) MF_SPMM_S(float, 4, 128);
    else if (vec == 4) MF_SPMM_S(float, 4, 0);
    else MF_SPMM_S(float, 1, 0);
  } else {
    if (ld == 256) MF_SPMM_S(double, 2, 256);
    else if (vec == 2) MF_SPMM_S(double, 2, 0);
    else MF_SPMM_S(double, 1, 0);
  }
#undef MF_SPMM_S
#undef MF_SPMM_D
#undef MF_SPMM_L
#undef MF_SPMM_K
  return check_launch("spmm_csr");
}

}  // namespace mf
