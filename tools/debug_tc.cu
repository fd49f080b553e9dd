// Bring-up harness for gemm_tcgen05.cu (not part of the library): one CTA, one stage.
// Dumps the TMA-written shared-memory tiles and the TMEM accumulator for exact small-integer
// inputs so that a wrong layout / descriptor shows up as a specific permutation.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --expt-relaxed-constexpr \
//        -I matfree_b200/csrc tools/debug_tc.cu -o gpurun_out/debug_tc -L matfree_b200/_lib \
//        -lmatfree_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../matfree_b200/_lib'
#include "../matfree_b200/csrc/gemm_tcgen05.cu"

#include <cstdio>
#include <cstring>
#include <vector>

using namespace mf;

template <int SWB, bool A_MN, bool B_K = false>
__global__ void __launch_bounds__(128, 1)
dbg_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
           float* dumpA, float* dumpB, float* dumpD, int mode) {
  constexpr int BM = 128, BN = 32;
  using Cfg = TcCfg<BN, SWB>;
  constexpr int CH = Cfg::CH, BK = Cfg::BK, KSTEPS = Cfg::KSTEPS;
  constexpr uint32_t kSbo = 8 * SWB, kLboMn = BK * 128, kSboMn = 512, kStepMn = 1024;
  constexpr int KL = Cfg::KL, ML = kLayoutSw128Atom32;
  constexpr uint32_t kIdesc = instr_desc_tf32(128, BN, A_MN, !B_K);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + 2 * Cfg::A_PLANE;
  const uint32_t bars = base + Cfg::STAGE;
  const uint32_t full = bars, done = bars + 8, slot = bars + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + Cfg::STAGE + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(full, Cfg::STAGE);
    if (!A_MN) {
      tma_load_3d(sA, &tmA, full, 0, 0, 0);
      tma_load_3d(sA + Cfg::A_PLANE, &tmA, full, 0, 0, 1);
    } else {
      for (int c = 0; c < BM / CH; ++c) {
        tma_load_3d(sA + c * kLboMn, &tmA, full, c * CH, 0, 0);
        tma_load_3d(sA + Cfg::A_PLANE + c * kLboMn, &tmA, full, c * CH, 0, 1);
      }
    }
    if (B_K) {
      tma_load_4d(sB, &tmB, full, 0, 0, 0, 0);
      tma_load_4d(sB + Cfg::B_PLANE, &tmB, full, 0, 0, 1, 0);
    } else {
      for (int c = 0; c < BN / CH; ++c) {
        tma_load_4d(sB + c * kLboMn, &tmB, full, c * CH, 0, 0, 0);
        tma_load_4d(sB + Cfg::B_PLANE + c * kLboMn, &tmB, full, c * CH, 0, 1, 0);
      }
    }
  }
  mbar_wait(full, 0);
  // dump raw tiles (generic-proxy reads of TMA-written smem are fine after the barrier)
  const float* fa = reinterpret_cast<const float*>(base_ptr);
  const float* fb = reinterpret_cast<const float*>(base_ptr + 2 * Cfg::A_PLANE);
  for (int i = threadIdx.x; i < Cfg::A_PLANE / 4; i += 128) dumpA[i] = fa[i];
  for (int i = threadIdx.x; i < Cfg::B_PLANE / 4; i += 128) dumpB[i] = fb[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    for (int kk = 0; kk < KSTEPS; ++kk) {
      const uint64_t bhi = B_K ? smem_desc<KL>(sB + kk * 32, 16, kSbo)
                               : smem_desc<ML>(sB + kk * kStepMn, kLboMn, kSboMn);
      uint64_t ahi;
      if (!A_MN) ahi = smem_desc<KL>(sA + kk * 32, 16, kSbo);
      else ahi = smem_desc<ML>(sA + kk * kStepMn, kLboMn, kSboMn);
      umma_tf32(tmem, ahi, bhi, kIdesc, kk != 0 ? 1u : 0u);
    }
    umma_commit(done);
  }
  uint32_t spins = 0;
  if (mode == 1) {
    // no MMA consumer: write a pattern with tcgen05.st, read it back below
    if (threadIdx.x == 0) { /* MMA was still issued above; wait for it */ }
    while (!mbar_try_wait(done, 0)) ++spins;
    tc_fence_after();
    const uint32_t val = 1000u * warp + lane;
    for (int j = 0; j < 32; ++j) {
      const uint32_t bits = __float_as_uint((float)(val * 100 + j));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(
                       tmem + ((uint32_t)(warp * 32) << 16) + j),
                   "r"(bits)
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  } else {
    while (!mbar_try_wait(done, 0)) ++spins;
    if (mode == 2) {
      for (int i = 0; i < 2000; ++i) __nanosleep(100);
    }
    tc_fence_after();
  }
  if (threadIdx.x == 0) {
    dumpA[0] = __uint_as_float(tmem);
    dumpA[1] = (float)spins;
  }
  uint32_t r[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
  const int row = warp * 32 + lane;
  for (int j = 0; j < 32; ++j) dumpD[row * 32 + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 32);
}

template <int SWB, bool A_MN, bool B_K = false>
int run(const char* name, int mode = 0) {
  constexpr int BM = 128, BN = 32;
  using Cfg = TcCfg<BN, SWB>;
  constexpr int CH = Cfg::CH, BK = Cfg::BK;
  const int M = 128, K = BK, ld = BN;
  // A planes [2][rows][lda]; lo plane = 0
  const int a_rows = A_MN ? K : M, lda = A_MN ? M : K;
  std::vector<float> hA(2 * a_rows * lda, 0.f), hB(2 * K * ld, 0.f);
  auto Aval = [](int m, int k) { return (float)((m * 7 + k * 3) % 13 - 6); };
  auto Bval = [](int k, int n) { return (float)((k * 5 + n * 11) % 9 - 4); };
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      if (A_MN) hA[k * lda + m] = Aval(m, k);
      else hA[m * lda + k] = Aval(m, k);
    }
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < ld; ++n) {
      if (B_K) hB[n * K + k] = Bval(k, n);
      else hB[k * ld + n] = Bval(k, n);
    }
  float *dA, *dB, *dumpA, *dumpB, *dumpD;
  cudaMalloc(&dA, hA.size() * 4);
  cudaMalloc(&dB, hB.size() * 4);
  cudaMalloc(&dumpA, Cfg::A_PLANE);
  cudaMalloc(&dumpB, Cfg::B_PLANE);
  cudaMalloc(&dumpD, 128 * 32 * 4);
  cudaMemset(dumpD, 0xff, 128 * 32 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  const uint64_t a_plane_bytes = (uint64_t)a_rows * lda * 4;
  if (!A_MN) {
    const uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, 2};
    const uint64_t str[2] = {(uint64_t)lda * 4u, a_plane_bytes};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BM, 1};
    if (make_map(&tmA, dA, 3, dims, str, box, SWB)) { printf("mapA: %s\n", mf_last_error()); return 1; }
  } else {
    const uint64_t dims[3] = {(uint64_t)M, (uint64_t)K, 2};
    const uint64_t str[2] = {(uint64_t)lda * 4u, a_plane_bytes};
    const uint32_t box[3] = {(uint32_t)CH, (uint32_t)BK, 1};
    if (make_map(&tmA, dA, 3, dims, str, box, 32)) { printf("mapA: %s\n", mf_last_error()); return 1; }
  }
  {
    const uint64_t plane = (uint64_t)K * ld * 4u;
    const uint64_t dims[4] = {(uint64_t)(B_K ? K : ld), (uint64_t)(B_K ? ld : K), 2, 1};
    const uint64_t str[3] = {(uint64_t)(B_K ? K : ld) * 4u, plane, 2 * plane};
    const uint32_t box[4] = {(uint32_t)(B_K ? BK : CH), (uint32_t)(B_K ? BN : BK), 1, 1};
    if (make_map(&tmB, dB, 4, dims, str, box, B_K ? SWB : 32)) { printf("mapB: %s\n", mf_last_error()); return 1; }
  }
  auto kern = dbg_kernel<SWB, A_MN, B_K>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::STAGE + 2048);
  kern<<<1, 128, Cfg::STAGE + 2048>>>(tmA, tmB, dumpA, dumpB, dumpD, mode);
  cudaError_t e = cudaDeviceSynchronize();
  printf("== %s: sync -> %s\n", name, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> sa(Cfg::A_PLANE / 4), sb(Cfg::B_PLANE / 4), d(128 * 32);
  cudaMemcpy(sa.data(), dumpA, Cfg::A_PLANE, cudaMemcpyDeviceToHost);
  cudaMemcpy(sb.data(), dumpB, Cfg::B_PLANE, cudaMemcpyDeviceToHost);
  cudaMemcpy(d.data(), dumpD, 128 * 32 * 4, cudaMemcpyDeviceToHost);
  // smem A check against the expected swizzled layout
  int badA = 0, badB = 0, badD = 0;
  auto swz = [](uint32_t off) {  // byte offset -> swizzled byte offset (Swizzle<3,4,3> / <2,4,3>)
    const uint32_t bits = SWB == 128 ? 7u : 3u;
    return off ^ (((off >> 7) & bits) << 4);
  };
  auto swz_mn = [](uint32_t off) {  // Swizzle<2,5,2>: 32-byte chunks ^ (row % 4)
    return off ^ (((off >> 7) & 3u) << 5);
  };
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      uint32_t off;
      if (!A_MN) off = swz(m * SWB + k * 4);
      else off = swz_mn((m / CH) * (BK * 128) + k * 128 + (m % CH) * 4);
      if (sa[off / 4] != Aval(m, k)) ++badA;
    }
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < ld; ++n) {
      const uint32_t off = B_K ? swz((uint32_t)(n * SWB + k * 4))
                               : swz_mn((uint32_t)((n / CH) * (BK * 128) + k * 128 + (n % CH) * 4));
      if (sb[off / 4] != Bval(k, n)) ++badB;
    }
  double maxd = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < ld; ++n) {
      float want = 0;
      for (int k = 0; k < K; ++k) want += Aval(m, k) * Bval(k, n);
      if (d[m * 32 + n] != want) ++badD;
      if (fabs(d[m * 32 + n]) > maxd) maxd = fabs(d[m * 32 + n]);
    }
  printf("   smemA mismatches %d / %d, smemB mismatches %d / %d, D mismatches %d / %d (max|D| %g)\n",
         badA, M * K, badB, K * ld, badD, M * ld, maxd);
  {
    uint32_t tb; memcpy(&tb, &sa[0], 4);
    printf("   mode %d: tmem base 0x%08x, spins on done barrier %g\n", mode, tb, sa[1]);
    printf("   D[33][0..4): %g %g %g %g\n", d[33 * 32], d[33 * 32 + 1], d[33 * 32 + 2], d[33 * 32 + 3]);
  }
  printf("   smemA[0..8): ");
  for (int i = 0; i < 8; ++i) printf("%g ", sa[i]);
  printf("| D[0][0..8): ");
  for (int i = 0; i < 8; ++i) printf("%g ", d[i]);
  float w0 = 0; for (int k = 0; k < K; ++k) w0 += Aval(0, k) * Bval(k, 0);
  printf("| want D[0][0] = %g\n", w0);
  return badD != 0;
}

int main() {
  int rc = 0;
  rc |= run<128, false>("SW128 K-major A");
  run<128, false, true>("SW128 K-major A, K-major B");
  run<64, false, true>("SW64 K-major A, K-major B");
  rc |= run<128, true>("SW128 MN-major A");
  rc |= run<64, false>("SW64 K-major A");
  rc |= run<64, true>("SW64 MN-major A");
  return rc;
}
