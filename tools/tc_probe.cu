// tcgen05 probe for the tensor-core formulation of batched Hamming distance (north_star: "tensor cores only if a
// +-1 int8 tcgen05 GEMM formulation beats the popcount path").  Three measurements on one B200:
//   1. check   D[128 codes][256 queries] = A * B^T through tcgen05.mma (kind::i8 and kind::f8f6f4) from operands the
//              threads expand themselves into the no-swizzle K-major canonical shared-memory layout, against the host:
//              A[c][i] = bit i of code c (0/1), A[c][pad] = 1;  B[q][i] = 1 - 2 * (bit i of query q), B[q][pad] =
//              popc(q) - tau_q - 1   =>   D[c][q] = hamming(c, q) - tau_q - 1  (negative <=> candidate).
//   2. ldtm    tcgen05.ld throughput in bytes/clk/SM for 4 / 8 / 16 warps (the epilogue bound of that formulation).
//   3. mma     clocks per 128 x 256 x 32-byte tcgen05.mma for kind::i8 and kind::f8f6f4.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int KIND>   // 0: i8, 1: f8f6f4
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
#define TMEM_LD32(r, taddr)                                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"  \
               "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                             \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                        \
               : "r"(taddr))
#define TMEM_LD32_PACK(r, taddr)                                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"  \
               "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                             \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                        \
               : "r"(taddr))
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes (128 contiguous bytes);
// LBO = byte distance of core matrices adjacent in K, SBO = of 8-row groups adjacent in M/N
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int kind, int M, int N) {
  // c_format [4,6): 1 = F32, 2 = S32; a_format [7,10), b_format [10,13): i8: 1 = signed; f8f6f4: 0 = E4M3; K-major both
  return (kind == 0 ? (2u << 4) | (1u << 7) | (1u << 10) : (1u << 4)) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int M_ = 128, N_ = 256, KB = 96;           // 64 code bits + 32 pad bytes
constexpr uint32_t LBO = 128, SBO = (KB / 16) * 128;

__device__ __forceinline__ uint32_t op_off(uint32_t row, uint32_t kbyte) {
  return (kbyte >> 4) * LBO + (row >> 3) * SBO + (row & 7) * 16 + (kbyte & 15);
}


// ---- 4. the MMA issuer's loop: what does a unit cost? ------------------------------------------------------------
// MODE 0: 2 MMAs + commit;  1: + try_wait on an already completed barrier before;  2: + tcgen05.fence::after_thread_sync;
// MODE 3: like 2 but the wait is on a barrier completed by this loop's own commit of 4 units ago (the real dependency)
template <int MODE>
__global__ void __launch_bounds__(128) mma_loop_kernel(int iters, long long* clk, int n_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_done, bar_sink[4];
  __shared__ uint32_t tmem_base;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < (M_ + N_) * KB; i += 128) smem[i] = 0;
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_sink[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_done)) : "memory");     // phase 0 complete for good
  }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  long long t0 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(0, M_, n_cols);
    const uint64_t ad = make_desc(smem_u32(smem), LBO, SBO), bd = make_desc(smem_u32(smem + M_ * KB), LBO, SBO);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s4 = it & 3;
      if (MODE == 1 || MODE == 2) mbar_wait(&bar_done, 0);
      if (MODE == 3 && it >= 4) mbar_wait(&bar_sink[s4], ((it >> 2) - 1) & 1);
      if (MODE >= 2) tc_fence_after();
      tc_mma<0>(tb + s4 * 128, ad, bd, idesc, 0);
      tc_mma<0>(tb + s4 * 128, ad + 16, bd + 16, idesc, 1);
      tc_commit(&bar_sink[s4]);
    }
    tc_commit(&bar_done);    // harmless extra arrival; just to have something to wait on below
  }
  __syncthreads();
  if (tid == 0) { mbar_wait(&bar_sink[(iters - 1) & 3], ((iters - 1) >> 2) & 1); clk[blockIdx.x] = clock64() - t0; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}


// ---- 5. what slows the MMAs down in the real kernel?  MMA issue loop (2 x 128x64x32B + commit) with company: --------------
// bit 0: 8 warps store 128-bit words to shared memory flat out (the expanders); bit 1: 8 warps read TMEM flat out (the epilogue);
// bit 2: 8 warps spin on mbarrier try_wait
__global__ void __launch_bounds__(25 * 32) mma_company_kernel(int iters, long long* clk, int company, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_sink[4], bar_never;
  __shared__ uint32_t tmem_base;
  __shared__ volatile int stop;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t i = tid; i < 100 * 1024; i += blockDim.x) smem[i] = 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar_sink[i], 1);
    mbar_init(&bar_never, 1);
    stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 24) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(0, M_, 64);
      const uint64_t ad = make_desc(smem_u32(smem), 128, 512), bd = make_desc(smem_u32(smem + 8192), 128, 512);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const int s4 = it & 3;
        tc_mma<0>(tb + s4 * 128, ad, bd, idesc, 0);
        tc_mma<0>(tb + s4 * 128, ad + 16, bd + 16, idesc, 1);
        tc_commit(&bar_sink[s4]);
      }
      mbar_wait(&bar_sink[(iters - 1) & 3], ((iters - 1) >> 2) & 1);
      clk[blockIdx.x] = clock64() - t0;
      stop = 1;
    }
  } else if (warp < 8) {
    if (company & 1) {
      uint4* dst = reinterpret_cast<uint4*>(smem + 16384 + warp * 8192);
      uint32_t v = tid;
      while (!stop) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[i * 32 + lane] = make_uint4(v, v + 1, v + 2, v + 3);
        v += 4;
        if (company & 8) proxy_fence();
        if (company & 16) __nanosleep(500);
      }
    }
  } else if (warp < 16) {
    if (company & 2) {
      uint32_t acc = 0;
      const uint32_t ta = tb + (((warp & 3) * 32) << 16) + 256;
      while (!stop) {
        uint32_t r[32];
        TMEM_LD32_PACK(r, ta);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) acc = __vimax3_s16x2(acc, r[j], r[j + 1]);
      }
      if (acc == 0x12345678u) sink[0] = acc;
    }
  } else if (warp == 23) {
    // latency probe: dependent shared-memory loads while the others are busy
    volatile uint32_t* p32 = reinterpret_cast<volatile uint32_t*>(smem + 90 * 1024);
    long long tsum = 0; int n = 0;
    uint32_t idx = 0;
    while (!stop && n < 100000) {
      const long long a = clock64();
#pragma unroll
      for (int i = 0; i < 8; ++i) idx = p32[idx & 63];
      tsum += clock64() - a; ++n;
    }
    if (lane == 0) { clk[gridDim.x + blockIdx.x] = n ? tsum / n / 8 : 0; if (idx == 12345) sink[0] = idx; }
  } else {
    if ((company & 4) && lane == 0) {
      while (!stop) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar_never)), "r"(0u) : "memory");
        if (ok) break;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}


// ---- 6. operand setup: tcgen05.mma takes its operands from UNIFORM registers.  What does it cost when they change? ---------
// MODE 0: lane 0 alone, operands change every unit (like the first production loop: R2UR per operand per MMA)
// MODE 1: whole warp runs the loop, operands derived from the loop counter only, MMA issued under elect.sync
// MODE 2: like 0, but the varying part of the operands goes through a reduction (redux.sync -> uniform register) ... n/a for one lane
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
template <int MODE>
__global__ void __launch_bounds__(128) mma_operand_kernel(int iters, long long* clk, const uint32_t* perm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_sink[4];
  __shared__ uint32_t tmem_base;
  __shared__ uint32_t s_perm[64];
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t i = tid; i < 100 * 1024; i += blockDim.x) smem[i] = 0;
  if (tid < 64) s_perm[tid] = perm[tid];
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar_sink[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 1) {
    const uint32_t idesc = make_idesc(0, M_, 64);
    const uint64_t ad0 = make_desc(smem_u32(smem), 128, 512), bd = make_desc(smem_u32(smem + 65536), 128, 512);
    long long t0 = 0;
    if (MODE == 0) {
      if (lane == 0) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
          const uint32_t as = s_perm[it & 63] & 7;                    // a value the compiler cannot prove uniform
          const uint64_t ad = ad0 + (uint64_t)(as * 512);
          const uint32_t dcol = tb + (it & 3) * 128;
          tc_mma<0>(dcol, ad, bd, idesc, 0);
          tc_mma<0>(dcol, ad + 16, bd + 16, idesc, 1);
          if ((it & 3) == 3) tc_commit(&bar_sink[0]);
        }
        tc_commit(&bar_sink[1]);
        mbar_wait(&bar_sink[1], 0);
        clk[blockIdx.x] = clock64() - t0;
      }
    } else {
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t as = MODE == 1 ? (uint32_t)(it * 5) & 7 : __reduce_max_sync(0xffffffffu, s_perm[it & 63] & 7);
        const uint64_t ad = ad0 + (uint64_t)(as * 512);
        const uint32_t dcol = tb + (it & 3) * 128;
        if (elect_one()) {
          tc_mma<0>(dcol, ad, bd, idesc, 0);
          tc_mma<0>(dcol, ad + 16, bd + 16, idesc, 1);
          if ((it & 3) == 3) tc_commit(&bar_sink[0]);
        }
        __syncwarp();
      }
      if (lane == 0) {
        tc_commit(&bar_sink[1]);
        mbar_wait(&bar_sink[1], 0);
        clk[blockIdx.x] = clock64() - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}


// ---- 7. how can the MMA thread learn that its inputs are ready without stalling behind its own MMAs? ---------------------
// (a shared-memory load or an mbarrier try_wait by the issuing thread costs ~100-300 clocks: it waits for the MMAs in flight)
// MODE 0: no synchronisation (reference);  MODE 1: bar.sync with a partner warp (named barrier 1) before every unit;
// MODE 2: ld.volatile.shared of a flag before every unit;  MODE 3: ld.global of a flag before every unit
template <int MODE>
__global__ void __launch_bounds__(128) mma_sync_kernel(int iters, long long* clk, const uint32_t* gflag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_sink[4];
  __shared__ uint32_t tmem_base;
  __shared__ volatile uint32_t s_flag;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t i = tid; i < 100 * 1024; i += blockDim.x) smem[i] = 0;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar_sink[i], 1);
    s_flag = 1;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 1) {
    const uint32_t idesc = make_idesc(0, M_, 64);
    const uint64_t ad = make_desc(smem_u32(smem), 128, 512), bd = make_desc(smem_u32(smem + 65536), 128, 512);
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
      if (MODE == 1) asm volatile("bar.sync 1, 64;" ::: "memory");
      if (MODE == 2) acc += s_flag;
      if (MODE == 3) acc += __ldcg(gflag);
      if (MODE == 4) { acc += s_flag; tc_fence_after(); }
      if (MODE == 5) { uint32_t v; asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)&s_flag)) : "memory"); acc += v; tc_fence_after(); }
      if (MODE == 6) { if (lane == 0) { uint32_t v; do { asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)&s_flag)) : "memory"); } while (v != 1); acc += v; tc_fence_after(); } }
      if (lane == 0 && acc != 0xFFFFFFFFu) {
        const uint32_t dcol = tb + (it & 3) * 128;
        tc_mma<0>(dcol, ad, bd, idesc, 0);
        tc_mma<0>(dcol, ad + 16, bd + 16, idesc, 1);
        tc_commit(&bar_sink[it & 3]);
      }
      __syncwarp();
    }
    if (lane == 0) {
      mbar_wait(&bar_sink[(iters - 1) & 3], ((iters - 1) >> 2) & 1);
      clk[blockIdx.x] = clock64() - t0;
    }
  } else if (warp == 2 && MODE == 1) {
    for (int it = 0; it < iters; ++it) asm volatile("bar.sync 1, 64;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

static uint64_t sm64(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
// ---- 1. correctness ------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(128) check_kernel(const uint64_t* codes, const uint64_t* queries, const int* pad, int32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                      // 128 x 96
  uint8_t* sB = smem + M_ * KB;            // 256 x 96
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  const uint8_t one = KIND == 0 ? 1 : 0x38, mone = KIND == 0 ? 0xFF : 0xB8;
  for (uint32_t i = tid; i < (M_ + N_) * KB; i += 128) smem[i] = 0;
  __syncthreads();
  {
    const uint64_t c = codes[tid];
    for (int b = 0; b < 64; ++b) sA[op_off(tid, b)] = (c >> b) & 1 ? one : 0;
    sA[op_off(tid, 64)] = one;
    for (uint32_t q = tid; q < N_; q += 128) {
      const uint64_t v = queries[q];
      for (int b = 0; b < 64; ++b) sB[op_off(q, b)] = (v >> b) & 1 ? mone : one;
      if (KIND == 0) sB[op_off(q, 64)] = (uint8_t)(int8_t)pad[q];
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(KIND, M_, N_);
    for (int k = 0; k < KB / 32; ++k)
      tc_mma<KIND>(tb, make_desc(smem_u32(sA) + k * 2 * LBO, LBO, SBO), make_desc(smem_u32(sB) + k * 2 * LBO, LBO, SBO), idesc, k > 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N_; c0 += 32) {
    uint32_t r[32];
    TMEM_LD32(r, tb + ((warp * 32) << 16) + c0);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)tid * N_ + c0 + j] = KIND == 0 ? (int32_t)r[j] : (int32_t)__uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}


// ---- 1c. the production encoding: codes as A, bytes +-a_j from two instructions per 4 bits (LOP3 + IMAD); queries as B,
// bytes +-64/a_j; D = 64 * (64 - 2 * hamming); epilogue: packed 16-bit loads + VIMNMX3 row maximum ----------------------
__device__ __forceinline__ void expand_code_word(uint32_t x, uint32_t* out8) {
  // out8[j] holds bits j, j+8, j+16, j+24 of x as bytes +a_j (bit clear) / -a_j (bit set); a = 1,1,2,4,8,16,32,64
  out8[0] = (x & 0x01010101u) * 254u + 0x01010101u;
  out8[1] = (x & 0x02020202u) * 127u + 0x01010101u;
  out8[2] = (x & 0x04040404u) * 63u + 0x02020202u;
  out8[3] = (x & 0x08080808u) * 31u + 0x04040404u;
  out8[4] = (x & 0x10101010u) * 15u + 0x08080808u;
  out8[5] = (x & 0x20202020u) * 7u + 0x10101010u;
  out8[6] = (x & 0x40404040u) * 3u + 0x20202020u;
  out8[7] = (x & 0x80808080u) * 1u + 0x40404040u;
}
__device__ __forceinline__ void expand_query_word(uint32_t x, uint32_t* out8) {
  const int bw[8] = {64, 64, 32, 16, 8, 4, 2, 1};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int v = (x >> (j + 8 * i)) & 1 ? -bw[j] : bw[j];
      r |= (uint32_t)(uint8_t)(int8_t)v << (8 * i);
    }
    out8[j] = r;
  }
}
constexpr int KB2 = 64;
constexpr uint32_t SBO2 = (KB2 / 16) * 128;
__global__ void __launch_bounds__(128) check2_kernel(const uint64_t* codes, const uint64_t* queries, int n_cols, int32_t* out, int32_t* rowmax) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                      // 128 x 64
  uint8_t* sB = smem + M_ * KB2;           // n_cols x 64
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  {
    const uint64_t c = codes[tid];
    uint32_t o[16];
    expand_code_word((uint32_t)c, o); expand_code_word((uint32_t)(c >> 32), o + 8);
    for (int g = 0; g < 4; ++g)
      *reinterpret_cast<uint4*>(sA + g * LBO + (tid >> 3) * SBO2 + (tid & 7) * 16) = make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
    for (uint32_t q = tid; q < (uint32_t)n_cols; q += 128) {
      const uint64_t v = queries[q];
      expand_query_word((uint32_t)v, o); expand_query_word((uint32_t)(v >> 32), o + 8);
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(sB + g * LBO + (q >> 3) * SBO2 + (q & 7) * 16) = make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(0, M_, n_cols);
    for (int k = 0; k < KB2 / 32; ++k)
      tc_mma<0>(tb, make_desc(smem_u32(sA) + k * 2 * LBO, LBO, SBO2), make_desc(smem_u32(sB) + k * 2 * LBO, LBO, SBO2), idesc, k > 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < n_cols; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tb + ((warp * 32) << 16) + c0));
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(size_t)tid * 256 + c0 + j] = (int32_t)r[j];
  }
  // packed row maximum over all columns, 16 columns (8 registers) at a time
  uint32_t acc = 0x80008000u;
  for (int c0 = 0; c0 < n_cols; c0 += 16) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tb + ((warp * 32) << 16) + c0));
    tmem_ld_wait();
    for (int j = 0; j < 8; j += 2) acc = __vimax3_s16x2(acc, r[j], r[j + 1]);
  }
  rowmax[tid] = max((int32_t)(int16_t)(acc & 0xFFFF), (int32_t)(int16_t)(acc >> 16));
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}
static int run_check2(int n_cols) {
  std::vector<uint64_t> codes(M_), queries(256);
  uint64_t s = 4242 + n_cols;
  for (auto& c : codes) c = sm64(s);
  for (auto& q : queries) q = sm64(s);
  for (int i = 0; i < 8; ++i) queries[i] = codes[5 * i] ^ (0x8000000000000001ull << i);      // a few near neighbours
  uint64_t *dc, *dq; int32_t *dout, *drm;
  CK(cudaMalloc(&dc, M_ * 8)); CK(cudaMalloc(&dq, 256 * 8)); CK(cudaMalloc(&dout, M_ * 256 * 4)); CK(cudaMalloc(&drm, M_ * 4));
  CK(cudaMemcpy(dc, codes.data(), M_ * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dq, queries.data(), 256 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, M_ * 256 * 4));
  const size_t smem = (M_ + 256) * KB2;
  CK(cudaFuncSetAttribute(check2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  check2_kernel<<<1, 128, smem>>>(dc, dq, n_cols, dout, drm);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> out(M_ * 256), rm(M_);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rm.data(), drm, rm.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0, badmax = 0;
  for (int c = 0; c < M_; ++c) {
    int best = -100000;
    for (int q = 0; q < n_cols; ++q) {
      const int want = 64 * (64 - 2 * __builtin_popcountll(codes[c] ^ queries[q]));
      best = want > best ? want : best;
      if (out[c * 256 + q] != want && bad++ < 5) printf("  check2 mismatch c=%d q=%d got %d want %d\n", c, q, out[c * 256 + q], want);
    }
    if (rm[c] != best && badmax++ < 5) printf("  check2 rowmax mismatch c=%d got %d want %d\n", c, rm[c], best);
  }
  printf("check2 weighted encoding N=%d: %s (%d value mismatches, %d row-max mismatches)\n", n_cols, bad || badmax ? "FAIL" : "OK", bad, badmax);
  return bad + badmax;
}


// ---- 8. A operand from tensor memory: the expanded codes never touch shared memory ----------------------------------------
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
#define TMEM_ST16(taddr, r)                                                                                                          \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"             \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),      \
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")
__global__ void __launch_bounds__(128) check3_kernel(const uint64_t* codes, const uint64_t* queries, int n_cols, int32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;                      // n_cols x 64
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  uint32_t o[16];
  for (uint32_t q = tid; q < (uint32_t)n_cols; q += 128) {
    const uint64_t v = queries[q];
    expand_query_word((uint32_t)v, o); expand_query_word((uint32_t)(v >> 32), o + 8);
    for (int g = 0; g < 4; ++g)
      *reinterpret_cast<uint4*>(sB + g * LBO + (q >> 3) * SBO2 + (q & 7) * 16) = make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  {
    const uint64_t c = codes[tid];
    expand_code_word((uint32_t)c, o); expand_code_word((uint32_t)(c >> 32), o + 8);
    TMEM_ST16(tb + ((warp * 32) << 16) + 256, o);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = make_idesc(0, M_, n_cols);
    for (int k = 0; k < 2; ++k)
      tc_mma_ts(tb, tb + 256 + 8 * k, make_desc(smem_u32(sB) + k * 2 * LBO, LBO, SBO2), idesc, k > 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < n_cols; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tb + ((warp * 32) << 16) + c0));
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(size_t)tid * 256 + c0 + j] = (int32_t)r[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}
static int run_check3(int n_cols) {
  std::vector<uint64_t> codes(M_), queries(256);
  uint64_t s = 777 + n_cols;
  for (auto& c : codes) c = sm64(s);
  for (auto& q : queries) q = sm64(s);
  uint64_t *dc, *dq; int32_t* dout;
  CK(cudaMalloc(&dc, M_ * 8)); CK(cudaMalloc(&dq, 256 * 8)); CK(cudaMalloc(&dout, M_ * 256 * 4));
  CK(cudaMemcpy(dc, codes.data(), M_ * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dq, queries.data(), 256 * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, M_ * 256 * 4));
  const size_t smem = 256 * KB2;
  CK(cudaFuncSetAttribute(check3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  check3_kernel<<<1, 128, smem>>>(dc, dq, n_cols, dout);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> out(M_ * 256);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int c = 0; c < M_; ++c)
    for (int q = 0; q < n_cols; ++q) {
      const int want = 64 * (64 - 2 * __builtin_popcountll(codes[c] ^ queries[q]));
      if (out[c * 256 + q] != want && bad++ < 5) printf("  check3 mismatch c=%d q=%d got %d want %d\n", c, q, out[c * 256 + q], want);
    }
  printf("check3 A operand from TMEM (tcgen05.st 32x32b.x16 per row) N=%d: %s (%d mismatches)\n", n_cols, bad ? "FAIL" : "OK", bad);
  return bad;
}
// MMA rate with A from TMEM
__global__ void __launch_bounds__(128) mma_ts_rate_kernel(int iters, long long* clk, int n_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < 256 * KB2; i += 128) smem[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  long long t0 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(0, M_, n_cols);
    const uint64_t bd = make_desc(smem_u32(smem), LBO, SBO2);
    t0 = clock64();
    for (int it = 0; it < iters; it += 4) {
      tc_mma_ts(tb, tb + 448, bd, idesc, 1); tc_mma_ts(tb, tb + 456, bd + 16, idesc, 1);
      tc_mma_ts(tb + 256, tb + 464, bd, idesc, 1); tc_mma_ts(tb + 256, tb + 472, bd + 16, idesc, 1);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) clk[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// ---- 1b. f16 accumulators: how are they laid out in TMEM? ------------------------------------------------------------
__global__ void __launch_bounds__(128) f16_layout_kernel(const uint64_t* codes, const uint64_t* queries, uint32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + M_ * KB;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < (M_ + N_) * KB; i += 128) smem[i] = 0;
  __syncthreads();
  const uint64_t c = codes[tid];
  for (int b = 0; b < 64; ++b) sA[op_off(tid, b)] = (c >> b) & 1 ? 0x38 : 0;
  for (uint32_t q = tid; q < N_; q += 128) {
    const uint64_t v = queries[q];
    for (int b = 0; b < 64; ++b) sB[op_off(q, b)] = (v >> b) & 1 ? 0xB8 : 0x38;
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  {  // poison the columns first so that untouched cells are recognisable
    uint32_t r[32];
    for (int j = 0; j < 32; ++j) r[j] = 0xDEADBEEFu;
    for (int c0 = 0; c0 < 256; c0 += 32)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
                   "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(tb + ((warp * 32) << 16) + c0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                   "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
                   "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
                   "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = (0u << 4) | ((uint32_t)(N_ >> 3) << 17) | ((uint32_t)(M_ >> 4) << 24);      // f8f6f4, E4M3 x E4M3, D = F16
    for (int k = 0; k < 2; ++k)
      tc_mma<1>(tb, make_desc(smem_u32(sA) + k * 2 * LBO, LBO, SBO), make_desc(smem_u32(sB) + k * 2 * LBO, LBO, SBO), idesc, k > 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N_; c0 += 32) {
    uint32_t r[32];
    TMEM_LD32(r, tb + ((warp * 32) << 16) + c0);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)tid * N_ + c0 + j] = r[j];
  }
  {
    uint32_t r[32];
    TMEM_LD32_PACK(r, tb + ((warp * 32) << 16));
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(size_t)M_ * N_ + tid * 32 + j] = r[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 256);
}

// ---- 2. tcgen05.ld throughput --------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(512) ldtm_kernel(int iters, uint32_t* sink, long long* clk) {
  __shared__ uint32_t tmem_base;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base + (((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 2) {
      uint32_t a[32], b[32];
      const uint32_t col = ((it + warp) & 3) * 128;
      TMEM_LD32_PACK(a, tb + col);
      TMEM_LD32_PACK(b, tb + col + 64);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 2) acc = __vimax3_s16x2(acc, a[j], a[j + 1]);
#pragma unroll
      for (int j = 0; j < 32; j += 2) acc = __vimax3_s16x2(acc, b[j], b[j + 1]);
      continue;
    }
    uint32_t a[32], b[32], c[32], d[32];
    const uint32_t col = ((it + warp) & 3) * 128;
    TMEM_LD32(a, tb + col);
    TMEM_LD32(b, tb + col + 32);
    TMEM_LD32(c, tb + col + 64);
    TMEM_LD32(d, tb + col + 96);
    tmem_ld_wait();
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) acc |= (a[j] | a[j + 1]) | (b[j] | b[j + 1]) | ((c[j] | c[j + 1]) | (d[j] | d[j + 1]));
    } else {
      acc |= a[0] | b[1] | c[2] | d[3];
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (tid == 0) clk[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- 3. MMA rate ---------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(128) mma_rate_kernel(int iters, long long* clk, int n_cols = N_, int n_acc = 2, int acc_stride = 256) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < (M_ + N_) * KB; i += 128) smem[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  long long t0 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(KIND, M_, n_cols);
    const uint64_t ad = make_desc(smem_u32(smem), LBO, SBO), bd = make_desc(smem_u32(smem + M_ * KB), LBO, SBO);
    t0 = clock64();
    if (n_acc == 1) { for (int it = 0; it < iters; it += 4) { tc_mma<KIND>(tb, ad, bd, idesc, 1); tc_mma<KIND>(tb, ad, bd, idesc, 1); tc_mma<KIND>(tb, ad, bd, idesc, 1); tc_mma<KIND>(tb, ad, bd, idesc, 1); } }
    else if (n_acc == 2) { for (int it = 0; it < iters; it += 4) { tc_mma<KIND>(tb, ad, bd, idesc, 1); tc_mma<KIND>(tb + acc_stride, ad, bd, idesc, 1); tc_mma<KIND>(tb, ad, bd, idesc, 1); tc_mma<KIND>(tb + acc_stride, ad, bd, idesc, 1); } }
    else { for (int it = 0; it < iters; it += 4) { tc_mma<KIND>(tb, ad, bd, idesc, 1); tc_mma<KIND>(tb + acc_stride, ad, bd, idesc, 1); tc_mma<KIND>(tb + 2 * acc_stride, ad, bd, idesc, 1); tc_mma<KIND>(tb + 3 * acc_stride, ad, bd, idesc, 1); } }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) clk[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}


template <int KIND>
static int run_check() {
  std::vector<uint64_t> codes(M_), queries(N_);
  std::vector<int> pad(N_);
  uint64_t s = 42 + KIND;
  for (auto& c : codes) c = sm64(s);
  for (int q = 0; q < N_; ++q) { queries[q] = sm64(s); pad[q] = KIND == 0 ? __builtin_popcountll(queries[q]) - (20 + q % 20) - 1 : 0; }
  uint64_t *dc, *dq; int* dp; int32_t* dout;
  CK(cudaMalloc(&dc, M_ * 8)); CK(cudaMalloc(&dq, N_ * 8)); CK(cudaMalloc(&dp, N_ * 4)); CK(cudaMalloc(&dout, M_ * N_ * 4));
  CK(cudaMemcpy(dc, codes.data(), M_ * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dq, queries.data(), N_ * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dp, pad.data(), N_ * 4, cudaMemcpyHostToDevice));
  const size_t smem = (M_ + N_) * KB;
  CK(cudaFuncSetAttribute(check_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  check_kernel<KIND><<<1, 128, smem>>>(dc, dq, dp, dout);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> out(M_ * N_);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int c = 0; c < M_; ++c)
    for (int q = 0; q < N_; ++q) {
      const int ham = __builtin_popcountll(codes[c] ^ queries[q]);
      const int want = KIND == 0 ? ham - (20 + q % 20) - 1 : ham - __builtin_popcountll(queries[q]);
      if (out[c * N_ + q] != want && bad++ < 5) printf("  mismatch c=%d q=%d got %d want %d\n", c, q, out[c * N_ + q], want);
    }
  printf("check kind=%s: %s (%d mismatches of %d)\n", KIND == 0 ? "i8" : "f8f6f4(e4m3)", bad ? "FAIL" : "OK", bad, M_ * N_);
  return bad;
}

int main(int argc, char** argv) {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  int rc = 0;
  rc |= run_check<0>();
  rc |= run_check<1>();
  for (int n : {16, 48, 64, 256}) rc |= run_check2(n);
  for (int n : {16, 64, 192}) rc |= run_check3(n);
  {
    long long* dck; CK(cudaMalloc(&dck, 148 * 8 * 2)); std::vector<long long> ck(sms);
    CK(cudaFuncSetAttribute(mma_ts_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * KB2));
    for (int n : {16, 32, 64, 128, 256}) {
      const int iters = 4000;
      mma_ts_rate_kernel<<<sms, 128, 256 * KB2>>>(iters, dck, n);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(ck.data(), dck, sms * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (auto c : ck) avg += (double)c; avg /= sms;
      printf("mma kind=i8 A from TMEM 128x%dx32B: %.1f clk per instruction\n", n, avg / iters);
    }
  }
  long long* dclk; uint32_t* dsink;
  CK(cudaMalloc(&dclk, sms * 8)); CK(cudaMalloc(&dsink, 4));
  std::vector<long long> clk(sms);
  for (int threads : {128, 256, 512}) {
    const int iters = 4000;
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) ldtm_kernel<0><<<sms, threads>>>(iters, dsink, dclk); else if (mode == 1) ldtm_kernel<1><<<sms, threads>>>(iters, dsink, dclk); else ldtm_kernel<2><<<sms, threads>>>(iters, dsink, dclk);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(clk.data(), dclk, sms * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (auto c : clk) avg += (double)c; avg /= sms;
      const double bytes = (double)iters * (threads / 32) * 4 * 32 * 32 * 4;
      printf("ldtm %2d warps/SM %s: %.1f bytes/clk/SM = %.1f 32-bit results/clk/SM\n", threads / 32, mode == 1 ? "loads only      " : mode == 2 ? "pack::16b (128 columns) + 0.25 VIMNMX3.S16x2/result: 16-bit results counted as 4 bytes" : "loads + 0.5 LOP3/result",
             bytes / avg, bytes / avg / 4);
    }
  }
  {
    // f16 accumulator layout
    std::vector<uint64_t> codes(M_), queries(N_);
    uint64_t s = 7;
    for (auto& c : codes) c = sm64(s);
    for (auto& q : queries) q = sm64(s);
    uint64_t *dc, *dq; uint32_t* dout;
    CK(cudaMalloc(&dc, M_ * 8)); CK(cudaMalloc(&dq, N_ * 8)); CK(cudaMalloc(&dout, M_ * (N_ + 32) * 4));
    CK(cudaMemcpy(dc, codes.data(), M_ * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dq, queries.data(), N_ * 8, cudaMemcpyHostToDevice));
    const size_t smem2 = (M_ + N_) * KB;
    CK(cudaFuncSetAttribute(f16_layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    f16_layout_kernel<<<1, 128, smem2>>>(dc, dq, dout);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> out(M_ * (N_ + 32));
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    auto h2f = [](uint16_t h) { int e = (h >> 10) & 31, m = h & 1023; float v = e ? ldexpf(1.0f + m / 1024.0f, e - 15) : ldexpf(m / 1024.0f, -14); return (h & 0x8000) ? -v : v; };
    int packed_ok = 0, unpacked_ok = 0, poison = 0;
    for (int c = 0; c < M_; ++c)
      for (int q = 0; q < N_; ++q) {
        const float want = (float)(__builtin_popcountll(codes[c] ^ queries[q]) - __builtin_popcountll(queries[q]));
        const uint32_t wp = out[c * N_ + q / 2];
        if (h2f((uint16_t)(q & 1 ? wp >> 16 : wp & 0xFFFF)) == want) ++packed_ok;
        if (h2f((uint16_t)(out[c * N_ + q] & 0xFFFF)) == want) ++unpacked_ok;
      }
    for (int c = 0; c < M_; ++c) for (int j = 0; j < N_; ++j) poison += out[c * N_ + j] == 0xDEADBEEFu;
    {
      // .x32.pack::16b: register j = columns (2j, 2j+1) [64 columns], or something else?
      int ok64 = 0, ok32 = 0;
      for (int c = 0; c < M_; ++c)
        for (int j = 0; j < 32; ++j) {
          const uint32_t v = out[M_ * N_ + c * 32 + j];
          const uint32_t lo2 = out[c * N_ + 2 * j] & 0xFFFF, hi2 = out[c * N_ + 2 * j + 1] & 0xFFFF;
          ok64 += v == (lo2 | (hi2 << 16));
          if (j < 16) ok32 += v == (lo2 | (hi2 << 16));
        }
      printf("pack::16b .x32: registers j = columns (2j | 2j+1 << 16) for %d / %d (first 16 registers: %d / %d)\n", ok64, M_ * 32, ok32, M_ * 16);
    }
    printf("f16 accumulators: packed-hypothesis matches %d / %d, unpacked-hypothesis %d / %d, untouched cells %d; row0 cols0-3: %08x %08x %08x %08x col128: %08x\n",
           packed_ok, M_ * N_, unpacked_ok, M_ * N_, poison, out[0], out[1], out[2], out[3], out[128]);
  }
  const size_t smem = (M_ + N_) * KB;
  CK(cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int kind = 0; kind < 2; ++kind) {
    const int iters = 3000;
    if (kind == 0) mma_rate_kernel<0><<<sms, 128, smem>>>(iters, dclk); else mma_rate_kernel<1><<<sms, 128, smem>>>(iters, dclk);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(clk.data(), dclk, sms * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto c : clk) avg += (double)c; avg /= sms;
    printf("mma kind=%s 128x256x32B: %.1f clk per instruction = %.0f MAC/clk/SM, %.0f 64-bit code-query tests/clk/SM at 3 instr per tile\n",
           kind == 0 ? "i8" : "f8f6f4", avg / iters, 128.0 * 256 * 32 / (avg / iters), 128.0 * 256 / (3 * avg / iters));
  }
  {
    uint32_t hperm[64]; for (int i = 0; i < 64; ++i) hperm[i] = (i * 5 + 3) & 7;
    uint32_t* dperm; CK(cudaMalloc(&dperm, sizeof hperm)); CK(cudaMemcpy(dperm, hperm, sizeof hperm, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(mma_operand_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(mma_operand_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(mma_operand_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int mode = 0; mode < 3; ++mode) {
      const int iters = 4000;
      if (mode == 0) mma_operand_kernel<0><<<sms, 128, 100 * 1024>>>(iters, dclk, dperm);
      else if (mode == 1) mma_operand_kernel<1><<<sms, 128, 100 * 1024>>>(iters, dclk, dperm);
      else mma_operand_kernel<2><<<sms, 128, 100 * 1024>>>(iters, dclk, dperm);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(clk.data(), dclk, sms * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (auto c : clk) avg += (double)c; avg /= sms;
      printf("mma operand setup mode %d (0: one lane, operands from memory; 1: whole warp, operands from the loop counter, elect; 2: whole warp, redux): %.1f clk per tile of 2 MMAs\n", mode, avg / iters);
    }
  }
  {
    uint32_t* dflag; CK(cudaMalloc(&dflag, 4)); CK(cudaMemset(dflag, 0, 4));
    CK(cudaFuncSetAttribute(mma_sync_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(mma_sync_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(mma_sync_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(mma_sync_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(mma_sync_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(mma_sync_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(mma_sync_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int mode = 0; mode < 7; ++mode) {
      const int iters = 4000;
      if (mode == 4) mma_sync_kernel<4><<<sms, 128, 100 * 1024>>>(iters, dclk, dflag);
      else if (mode == 5) mma_sync_kernel<5><<<sms, 128, 100 * 1024>>>(iters, dclk, dflag);
      else if (mode == 6) mma_sync_kernel<6><<<sms, 128, 100 * 1024>>>(iters, dclk, dflag);
      else if (mode == 0) mma_sync_kernel<0><<<sms, 128, 100 * 1024>>>(iters, dclk, dflag);
      else if (mode == 1) mma_sync_kernel<1><<<sms, 128, 100 * 1024>>>(iters, dclk, dflag);
      else if (mode == 2) mma_sync_kernel<2><<<sms, 128, 100 * 1024>>>(iters, dclk, dflag);
      else mma_sync_kernel<3><<<sms, 128, 100 * 1024>>>(iters, dclk, dflag);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(clk.data(), dclk, sms * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (auto c : clk) avg += (double)c; avg /= sms;
      printf("mma + sync mode %d (0: none, 1: bar.sync with a partner warp, 2: shared-memory flag load, 3: global flag load, 4: 2 + tcgen05.fence::after, 5: ld.acquire + fence, 6: lane 0 polls ld.acquire + fence): %.1f clk per unit of 2 MMAs + commit\n", mode, avg / iters);
    }
  }
  CK(cudaFuncSetAttribute(mma_company_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  {
    long long* dclk2; CK(cudaMalloc(&dclk2, 2 * sms * 8));
    std::vector<long long> clk2(2 * sms);
    for (int company : {0, 1, 2, 3, 4, 7, 9, 11, 25, 27}) {
      const int iters = 4000;
      mma_company_kernel<<<sms, 25 * 32, 100 * 1024>>>(iters, dclk2, company, dsink);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(clk2.data(), dclk2, 2 * sms * 8, cudaMemcpyDeviceToHost));
      double avg = 0, lat = 0; for (int i = 0; i < sms; ++i) { avg += (double)clk2[i]; lat += (double)clk2[sms + i]; } avg /= sms; lat /= sms;
      printf("mma issue loop with company %2d (1: 8 warps STS.128, 2: 8 warps tcgen05.ld, 4: try_wait spinners, 8: + fence.proxy.async per 16 STS, 16: + 500 ns pause): %.1f clk per unit of 2 MMAs + commit; dependent LDS latency seen by another warp %.0f clk\n", company, avg / iters, lat);
    }
  }
  for (int mode = 0; mode < 4; ++mode) {
    const int iters = 4000;
    CK(cudaFuncSetAttribute(mma_loop_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(mma_loop_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(mma_loop_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(mma_loop_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (mode == 0) mma_loop_kernel<0><<<sms, 128, smem>>>(iters, dclk, 64); else if (mode == 1) mma_loop_kernel<1><<<sms, 128, smem>>>(iters, dclk, 64);
    else if (mode == 2) mma_loop_kernel<2><<<sms, 128, smem>>>(iters, dclk, 64); else mma_loop_kernel<3><<<sms, 128, smem>>>(iters, dclk, 64);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(clk.data(), dclk, sms * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto c : clk) avg += (double)c; avg /= sms;
    printf("mma issue loop (2 x 128x64x32B + commit per unit) mode %d: %.1f clk per unit\n", mode, avg / iters);
  }
  for (int n : {16, 32, 64, 128}) {
    const int iters = 3000;
    for (int n_acc : {1, 2, 4}) {
      mma_rate_kernel<0><<<sms, 128, smem>>>(iters, dclk, n, n_acc, 128);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(clk.data(), dclk, sms * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (auto c : clk) avg += (double)c; avg /= sms;
      printf("mma kind=i8 128x%dx32B, %d accumulators round-robin: %.1f clk per instruction\n", n, n_acc, avg / iters);
    }
  }
  return rc;
}
