// Tensor-core form of the covariance downdate of the EKF measurement update.
//
// Reference: Estimator::UpdateJosephForm (/root/reference/src/estimator.cpp:1257-1288).  After the gain
// kernel (ekf_kernels.cu) has produced HP = H P (M x N) and K^T = S^-1 HP (M x N), the only dense
// contraction left in the update is the rank-M downdate
//        P  <-  P - K (H P),      D[i][j] = sum_k Kt[k][i] * HP[k][j],   i, j < N, k < M.
// ekf_cov_kernel evaluates D in fp64 on the CUDA cores.  This file evaluates it on the 5th-generation
// tensor cores: tcgen05.mma kind::tf32 with the fp32 accumulator tile in TMEM, operands split as
// x = hi + lo (both TF32) and D = Alo*Bhi + Ahi*Blo + Ahi*Bhi ("3xTF32": ~21 mantissa bits per product,
// fp32 accumulation), then  P(fp64) -= D  in the epilogue.  There is no fp64 kind of tcgen05.mma, so this
// path is the "fp32 covariance" mode of SURVEY.md §8d (config 3): the correction carries fp32 accuracy,
// P itself stays fp64 in HBM so every other kernel is unchanged.  It is opt-in ("covariance_update":
// "tf32x3" in the estimator config, or flags bit 0 of xivo_ekf_update_ex); the default remains fp64.
//
// One CTA (128 threads = the 128 TMEM lanes) per (filter, 128-row tile, 32-column chunk) of the upper
// triangle (narrow chunks: the tile is latency-, not math-bound, so more CTAs per filter win).  Operands
// are staged by the CTA's threads (fp64 -> hi/lo TF32, transposed) into the canonical K-major no-swizzle
// shared-memory layout of the UMMA matrix descriptor: 8-row x 16-byte core matrices,
//   byte(row, k) = ((k / 4) * ROWS + row) * 16 + (k % 4) * 4
// i.e. SBO (8-row group stride) = 128 B and LBO (stride between the two 16-byte K chunks of one MMA) =
// ROWS * 16 B.  K (= measurement rows) is consumed in blocks of 32 (four K=8 MMAs per pass).
#include <atomic>
#include <cstdint>
#include <cstdlib>

#include "kernels.h"
#include "prof.h"

namespace xb {
namespace {

constexpr int TC_MT = 128;       // rows per tile = TMEM lanes
constexpr int TC_NT_MAX = 32;    // columns per chunk: narrow chunks = more CTAs per filter (the tile work is latency, not math)
constexpr int TC_KB = 32;        // K elements staged per block
constexpr int TC_THREADS = 128;

__device__ int g_tc_fault;  // set when an mbarrier wait timed out (never expected; keeps a bad build from hanging the GPU)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// UMMA shared-memory matrix descriptor (SM100): start address, LBO, SBO in 16-byte units, version 1,
// layout type 0 (no swizzle), base offset 0.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// Instruction descriptor, kind::tf32: D = F32, A = B = TF32, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_MT >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// four consecutive k of one row -> 16-byte chunk `chunk` of the hi array and of the lo array (x = hi + lo, both TF32)
__device__ __forceinline__ void stage_chunk(const double (&x)[4], uint32_t* hi, uint32_t* lo, int chunk) {
  uint4 h, l;
  uint32_t* hp = &h.x;
  uint32_t* lp = &l.x;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t t = to_tf32((float)x[q]);
    hp[q] = t;
    lp[q] = to_tf32((float)(x[q] - (double)__uint_as_float(t)));
  }
  reinterpret_cast<uint4*>(hi)[chunk] = h;
  reinterpret_cast<uint4*>(lo)[chunk] = l;
}

__global__ void __launch_bounds__(TC_THREADS) ekf_cov_tc_kernel(int N, const int* __restrict__ nsel, int Mdense, int Mmax,
                                                                const double* __restrict__ HP, const double* __restrict__ Kt,
                                                                double* __restrict__ P, int swap_lbo_sbo) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;

  const int b = blockIdx.z;
  const int M = nsel ? 2 * nsel[b] : Mdense;
  const int m0 = blockIdx.y * TC_MT;
  const int n0 = m0 + blockIdx.x * TC_NT_MAX;  // only the upper triangle: columns start at the tile's first row
  if (M == 0 || m0 >= N || n0 >= N) return;    // uniform over the CTA
  const int ncols = min(TC_NT_MAX, N - n0);
  const int NT = (ncols + 15) & ~15;           // UMMA N: multiple of 16 for M = 128
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint32_t* Ah = reinterpret_cast<uint32_t*>(tc_smem);
  uint32_t* Al = Ah + TC_KB * TC_MT;
  uint32_t* Bh = Al + TC_KB * TC_MT;
  uint32_t* Bl = Bh + TC_KB * NT;
  const double* __restrict__ HPb = HP + (size_t)b * Mmax * N;
  const double* __restrict__ Ktb = Kt + (size_t)b * Mmax * N;
  double* __restrict__ Pb = P + (size_t)b * N * N;

  const uint32_t tmem_cols = NT <= 32 ? 32u : NT <= 64 ? 64u : NT <= 128 ? 128u : 256u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  const uint32_t idesc = umma_idesc_tf32(NT);
  const uint32_t a_chunk = TC_MT * 16, b_chunk = (uint32_t)NT * 16;  // bytes between consecutive 16-byte K chunks
  const uint32_t a_lbo = swap_lbo_sbo ? 128u : a_chunk, a_sbo = swap_lbo_sbo ? a_chunk : 128u;
  const uint32_t b_lbo = swap_lbo_sbo ? 128u : b_chunk, b_sbo = swap_lbo_sbo ? b_chunk : 128u;
  const int nkb = (M + TC_KB - 1) / TC_KB;
  bool ok = true;
  uint32_t first = 1;

  for (int kb = 0; kb < nkb; ++kb) {
    const int k0 = kb * TC_KB;
    // ---- stage A(i, k) = Kt[k][m0 + i] and B(j, k) = HP[k][n0 + j] as hi/lo TF32, zero padded.  A thread owns one row and
    // converts four consecutive k at a time: four coalesced loads -> one 16-byte chunk of the hi array and one of the lo array.
    const int kreal = min(TC_KB, M - k0);
    const int kgroups = ((kreal + 7) >> 3) << 1;  // 16-byte chunks the MMAs of this block read
    {
      const int gi = m0 + tid;
      const double* __restrict__ src = Ktb + (size_t)k0 * N + gi;
#pragma unroll 2
      for (int g = 0; g < kgroups; ++g) {
        double x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = (gi < N && 4 * g + q < kreal) ? src[(size_t)(4 * g + q) * N] : 0.0;
        stage_chunk(x, Ah, Al, g * TC_MT + tid);
      }
    }
    for (int u = tid; u < kgroups * NT; u += TC_THREADS) {  // NT <= 32 rows: the k groups are spread over the warps
      const int g = u / NT, r = u - g * NT;
      const int gj = n0 + r;
      const double* __restrict__ src = HPb + (size_t)k0 * N + gj;
      double x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = (gj < N && 4 * g + q < kreal) ? src[(size_t)(4 * g + q) * N] : 0.0;
      stage_chunk(x, Bh, Bl, g * NT + r);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int ksteps = kgroups >> 1;  // K = 8 per tf32 MMA; all-zero steps are skipped
      const uint32_t aH = smem_u32(Ah), aL = smem_u32(Al), bH = smem_u32(Bh), bL = smem_u32(Bl);
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {  // small terms first
        const uint32_t sa = pass == 0 ? aL : aH, sb = pass == 1 ? bL : bH;
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = umma_smem_desc(sa + (uint32_t)ks * 2u * a_chunk, a_lbo, a_sbo);
          const uint64_t db = umma_smem_desc(sb + (uint32_t)ks * 2u * b_chunk, b_lbo, b_sbo);
          umma_tf32(tmem, da, db, idesc, first ? 0u : 1u);
          first = 0;
        }
      }
      // arrives on the mbarrier when every MMA issued so far has completed (implies fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    {  // wait for phase kb of the barrier (bounded: a broken build must not hang the device)
      const uint32_t parity = (uint32_t)kb & 1u, addr = smem_u32(&bar);
      uint32_t done = 0;
      const long long t_start = clock64();
      while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && clock64() - t_start > 400000000LL) break;  // ~0.2 s: the MMAs of one block take microseconds
      }
      if (!done) ok = false;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();  // every thread is past the wait before the operands are overwritten
  }

  // ---- epilogue: TMEM lane = row, 8 consecutive columns per tcgen05.ld; P -= D on the upper triangle, mirrored
  const int row = m0 + warp * 32 + lane;
  if (!__all_sync(0xffffffffu, ok)) {
    if (lane == 0) atomicExch(&g_tc_fault, 1);
  } else {
    for (int c = 0; c < NT; c += 8) {
      uint32_t v[8];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < N) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = n0 + c + j;
          if (col < N && col >= row) {
            const double val = Pb[(size_t)row * N + col] - (double)__uint_as_float(v[j]);
            Pb[(size_t)row * N + col] = val;
            if (col != row) Pb[(size_t)col * N + row] = val;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------------------
// Second formulation (default for "covariance_update": "tf32x3"; XIVO_TC_V1=1 keeps the kernel above for A/B parity).
// What changed and why (profiles/r01d_ncu_ekf_cov_tc.txt: the first kernel spent its time staging, 0.2-0.7 % tensor-pipe active):
//  * operands are not converted here: the gain kernel, which owns one state column per thread when it finishes K^T, writes K^T and HP
//    once more as TF32 hi/lo words directly in the UMMA canonical K-major layout  [filter][hi|lo][k/4][state column][4 words]
//    (column stride 16 B = one core-matrix row, 8 columns = one 128-byte core matrix, k-chunk stride = Npad * 16 B);
//  * a CTA owns a 128-row tile of one filter and walks ALL its 32-column chunks of the upper triangle: the A operand (128 rows x
//    all K, hi and lo) arrives once by TMA box loads {4, 128, 8} and stays in shared memory, the B chunks {4, 32, 8} are double
//    buffered and prefetched two chunks ahead;
//  * two fp32 accumulators in TMEM (2 x 32 columns): the MMAs of chunk i+1 are issued before the epilogue of chunk i starts;
//  * the epilogue goes through shared memory so that both the direct update P[i][j] and the mirrored one P[j][i] are row-contiguous
//    (the first kernel wrote 8-byte words 8 N bytes apart).  The mirrored element is recomputed from its own old value, which equals
//    P[i][j] bit for bit because P is kept exactly symmetric, so it stays exactly symmetric.
// The contraction is bound by the fp64 traffic of P (2 N^2 8 B per filter against 3 * 2 N^2 M flop): see DESIGN.md section 4.
constexpr int T2_MT = 128, T2_NT = 32, T2_KB = 32, T2_THREADS = 512;  // warps 0-3 own the 128 TMEM lanes; all 16 warps stream P in the epilogue
constexpr int T2_A_BLOCK = (T2_KB / 4) * T2_MT * 16;  // bytes of one k-block of A (hi or lo): 8 chunks x 128 rows x 16 B = 16 KB
constexpr int T2_B_BLOCK = (T2_KB / 4) * T2_NT * 16;  // 4 KB

__device__ __forceinline__ bool mbar_wait(uint32_t addr, uint32_t parity) {  // bounded (a broken build must not hang the device)
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 400000000LL) return false;
  }
  return true;
}
__device__ __forceinline__ void tma_box3(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(T2_THREADS) ekf_cov_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int N,
                                                                 const int* __restrict__ nsel, int Mdense, int KCmax, double* __restrict__ P) {
  extern __shared__ __align__(128) unsigned char t2_smem[];
  __shared__ __align__(8) unsigned long long bars[5];  // 0: A landed; 1, 2: B buffer landed; 3, 4: accumulator complete
  __shared__ uint32_t tmem_slot;
  __shared__ int fault;
  const int b = blockIdx.y, m0 = blockIdx.x * T2_MT;
  const int M = nsel ? 2 * nsel[b] : Mdense;
  if (M == 0 || m0 >= N) return;  // uniform over the CTA
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kblocks = (M + T2_KB - 1) / T2_KB;
  const int kbmax = KCmax / (T2_KB / 4);
  // shared: A hi [kbmax][16 KB] | A lo | B buffers 2 x (hi [kbmax][4 KB] | lo) | D staging 128 x 33 floats
  unsigned char* sAh = t2_smem;
  unsigned char* sAl = sAh + (size_t)kbmax * T2_A_BLOCK;
  unsigned char* sB = sAl + (size_t)kbmax * T2_A_BLOCK;
  const uint32_t b_buf_bytes = (uint32_t)kbmax * 2u * T2_B_BLOCK;
  float* sD = reinterpret_cast<float*>(sB + 2 * (size_t)b_buf_bytes);
  const uint32_t barA = smem_u32(&bars[0]);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int i = 0; i < 5; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fault = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int ncc = (N - m0 + T2_NT - 1) / T2_NT;                // column chunks of this row tile: columns [m0, N)
  const int zA = (b * 2) * KCmax, zB = zA;                      // first k-chunk of filter b's hi slab inside the maps (lo slab: + KCmax)
  const uint32_t idesc = umma_idesc_tf32(T2_NT);
  auto load_b = [&](int chunk) {                                // thread 0: TMA of column chunk `chunk` into buffer chunk & 1
    const int buf = chunk & 1;
    const uint32_t bar = smem_u32(&bars[1 + buf]);
    const uint32_t dst = smem_u32(sB) + (uint32_t)buf * b_buf_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kblocks * 2u * T2_B_BLOCK) : "memory");
    for (int kb = 0; kb < kblocks; ++kb) {
      tma_box3(dst + (uint32_t)kb * T2_B_BLOCK, &mapB, 0, m0 + chunk * T2_NT, zB + kb * (T2_KB / 4), bar);
      tma_box3(dst + (uint32_t)(kbmax + kb) * T2_B_BLOCK, &mapB, 0, m0 + chunk * T2_NT, zB + KCmax + kb * (T2_KB / 4), bar);
    }
  };
  auto issue_mma = [&](int chunk) -> bool {                     // thread 0: 3 passes x kblocks x 4 MMAs into accumulator chunk & 1, then commit
    const int buf = chunk & 1;
    if (!mbar_wait(smem_u32(&bars[1 + buf]), (uint32_t)(chunk >> 1) & 1u)) return false;
    if (chunk == 0 && !mbar_wait(barA, 0u)) return false;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t bH = smem_u32(sB) + (uint32_t)buf * b_buf_bytes, bL = bH + (uint32_t)kbmax * T2_B_BLOCK;
    const uint32_t aH = smem_u32(sAh), aL = smem_u32(sAl);
    const uint32_t acc = tmem + (uint32_t)buf * T2_NT;
    uint32_t first = 1;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {                      // small terms first: lo*hi, hi*lo, hi*hi
      const uint32_t sa = pass == 0 ? aL : aH, sb = pass == 1 ? bL : bH;
#pragma unroll 1
      for (int kb = 0; kb < kblocks; ++kb) {
        const int ksteps = min(T2_KB / 8, (M - kb * T2_KB + 7) / 8);  // K = 8 per MMA; all-zero tail steps are skipped
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = umma_smem_desc(sa + (uint32_t)kb * T2_A_BLOCK + (uint32_t)ks * 2u * (T2_MT * 16), T2_MT * 16, 128u);
          const uint64_t db = umma_smem_desc(sb + (uint32_t)kb * T2_B_BLOCK + (uint32_t)ks * 2u * (T2_NT * 16), T2_NT * 16, 128u);
          umma_tf32(acc, da, db, idesc, first ? 0u : 1u);
          first = 0;
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[3 + buf])) : "memory");
    return true;
  };
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barA), "r"((uint32_t)kblocks * 2u * T2_A_BLOCK) : "memory");
    for (int kb = 0; kb < kblocks; ++kb) {
      tma_box3(smem_u32(sAh) + (uint32_t)kb * T2_A_BLOCK, &mapA, 0, m0, zA + kb * (T2_KB / 4), barA);
      tma_box3(smem_u32(sAl) + (uint32_t)kb * T2_A_BLOCK, &mapA, 0, m0, zA + KCmax + kb * (T2_KB / 4), barA);
    }
    load_b(0);
    if (ncc > 1) load_b(1);
    if (!issue_mma(0)) fault = 1;
  }
  double* __restrict__ Pb = P + (size_t)b * N * N;
  for (int i = 0; i < ncc; ++i) {
    const int buf = i & 1;
    if (tid == 0 && i + 1 < ncc && !fault) {                    // next chunk's MMAs run while this chunk's epilogue does
      if (!issue_mma(i + 1)) fault = 1;
    }
    bool ok = mbar_wait(smem_u32(&bars[3 + buf]), (uint32_t)(i >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    ok = __all_sync(0xffffffffu, ok);
    if (!ok) {
      if (lane == 0) fault = 1;
    } else if (warp < 4) {
      // TMEM lane = row of the tile; 4 x 8 consecutive columns -> shared staging (row stride 33 floats: conflict-free both ways)
#pragma unroll
      for (int c = 0; c < T2_NT; c += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * T2_NT + c);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) sD[(warp * 32 + lane) * (T2_NT + 1) + c + j] = __uint_as_float(v[j]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (!fault) {
      const int n0 = m0 + i * T2_NT;
      // 4096 tile elements, 512 threads: every thread issues its 8 + 8 loads of P before the first dependent store, so that the SM keeps
      // ~8 K loads in flight (the first version of this loop, one load per thread at a time, ran at 5 % of the HBM rate).
      constexpr int PER = T2_MT * T2_NT / T2_THREADS;
      double pv[PER];
      size_t po[PER];
      float dv[PER];
      // direct part: P[m0 + r][n0 + c] for r <= c (global indices), 32 consecutive doubles per row
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int idx = tid + k * T2_THREADS;
        const int r = idx >> 5, c = idx & 31;
        const int gi = m0 + r, gj = n0 + c;
        const bool on = gi < N && gj < N && gi <= gj;
        po[k] = on ? (size_t)gi * N + gj : ~(size_t)0;
        dv[k] = sD[r * (T2_NT + 1) + c];
        pv[k] = on ? Pb[po[k]] : 0.0;
      }
#pragma unroll
      for (int k = 0; k < PER; ++k)
        if (po[k] != ~(size_t)0) Pb[po[k]] = pv[k] - (double)dv[k];
      // mirrored part: P[n0 + c][m0 + r] for r < c, 128 consecutive doubles per row; the old value read here equals P[m0 + r][n0 + c]
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int idx = tid + k * T2_THREADS;
        const int c = idx >> 7, r = idx & 127;
        const int gi = m0 + r, gj = n0 + c;
        const bool on = gi < N && gj < N && gi < gj;
        po[k] = on ? (size_t)gj * N + gi : ~(size_t)0;
        dv[k] = sD[r * (T2_NT + 1) + c];
        pv[k] = on ? Pb[po[k]] : 0.0;
      }
#pragma unroll
      for (int k = 0; k < PER; ++k)
        if (po[k] != ~(size_t)0) Pb[po[k]] = pv[k] - (double)dv[k];
    }
    __syncthreads();                                            // staging and B buffer `buf` are free again
    if (tid == 0 && i + 2 < ncc && !fault) load_b(i + 2);
  }
  if (fault && tid == 0) atomicExch(&g_tc_fault, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

}  // namespace

size_t ekf_cov_tc_smem(int N) {
  const int nt = std::min(TC_NT_MAX, (N + 15) & ~15);
  return (size_t)2 * TC_KB * (TC_MT + nt) * sizeof(uint32_t);
}

int launch_ekf_cov_tc(cudaStream_t st, int N, const int* nsel, int Mdense, int Mmax, const double* HP, const double* Kt, double* P,
                      int batch) {
  const int variant = 0;  // LBO = stride between the two 16-byte K chunks, SBO = stride between 8-row groups (confirmed on the B200, profiles/r01d_tc_probe.txt)
  const size_t smem = ekf_cov_tc_smem(N);
  XB_CUDA(cudaFuncSetAttribute(ekf_cov_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tm = (N + TC_MT - 1) / TC_MT, tn = (N + TC_NT_MAX - 1) / TC_NT_MAX;
  ProfScope ps("ekf_cov", st);
  ekf_cov_tc_kernel<<<dim3(tn, tm, batch), TC_THREADS, smem, st>>>(N, nsel, Mdense, Mmax, HP, Kt, P, variant);
  XB_CUDA(cudaGetLastError());
  return 0;
}

// ---- operand buffers of the second formulation -------------------------------------------------------------------------------
int tc_npad(int N) { return (N + 7) & ~7; }
int tc_kcmax(int Mmax) { return ((Mmax + T2_KB - 1) / T2_KB) * (T2_KB / 4); }
size_t tc_operand_words(int N, int Mmax, int batch) { return (size_t)batch * 2 * tc_kcmax(Mmax) * tc_npad(N) * 4; }

static int make_operand_map(CUtensorMap* out, const uint32_t* base, int Npad, int KCmax, int batch, int box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    XB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    XB_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from this driver");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  // {4 words, state column, k-chunk of (filter, hi|lo)}: column stride 16 B, chunk stride Npad * 16 B
  const cuuint64_t dims[3] = {4, (cuuint64_t)Npad, (cuuint64_t)batch * 2 * KCmax};
  const cuuint64_t strides[2] = {16, (cuuint64_t)Npad * 16};
  const cuuint32_t box[3] = {4, (cuuint32_t)box_rows, (cuuint32_t)(T2_KB / 4)};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint32_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed for the update operands");
  return 0;
}

int tc_operands_init(TcOperands* t, int N, int Mmax, int batch, uint32_t* kt32, uint32_t* hp32) {
  t->kt32 = kt32; t->hp32 = hp32;
  t->Npad = tc_npad(N); t->KCmax = tc_kcmax(Mmax); t->batch = batch;
  if (int rc = make_operand_map(&t->mapA, kt32, t->Npad, t->KCmax, batch, T2_MT)) return rc;
  return make_operand_map(&t->mapB, hp32, t->Npad, t->KCmax, batch, T2_NT);
}

size_t ekf_cov_tc2_smem(int KCmax) {
  const size_t kbmax = (size_t)KCmax / (T2_KB / 4);
  return kbmax * 2 * T2_A_BLOCK + 2 * kbmax * 2 * T2_B_BLOCK + (size_t)T2_MT * (T2_NT + 1) * sizeof(float);
}

int launch_ekf_cov_tc2(cudaStream_t st, int N, const int* nsel, int Mdense, const TcOperands& t, double* P, int batch) {
  const size_t smem = ekf_cov_tc2_smem(t.KCmax);
  XB_REQUIRE(smem <= 227 * 1024, "tensor-core downdate: measurement dimension too large for the resident A operand");
  XB_REQUIRE(batch <= t.batch, "tensor-core downdate: operand buffers hold fewer filters");
  static std::atomic<size_t> attr{0};
  if (smem > attr.load(std::memory_order_relaxed)) {
    XB_CUDA(cudaFuncSetAttribute(ekf_cov_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr.store(smem, std::memory_order_relaxed);
  }
  ProfScope ps("ekf_cov", st);
  ekf_cov_tc2_kernel<<<dim3((N + T2_MT - 1) / T2_MT, batch), T2_THREADS, smem, st>>>(t.mapA, t.mapB, N, nsel, Mdense, t.KCmax, P);
  XB_CUDA(cudaGetLastError());
  return 0;
}

int ekf_cov_tc_fault(cudaStream_t st) {
  int f = 0;
  if (cudaMemcpyFromSymbolAsync(&f, g_tc_fault, sizeof(int), 0, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
  return f;
}

}  // namespace xb
