// Tensor-core verification of batched-MIH / batched-scan work items (tcgen05.mma kind::i8, accumulators in TMEM).
//
// Same work items, thresholds, de-duplication and candidate buffers as bmih_verify_kernel (bmih.cuh) - i.e. the
// full-code verification of src/search_worker.cc:249-257 (compute_hamming_dist, Pilaf/image_tools.h:21-33) for a
// bucket x query-list rectangle - but the distance FILTER runs on the tensor cores:
//
//   Hamming distance as an int8 GEMM.  Code bit c and query bit q become bytes A = (c ? -a : +a), B = (q ? -b : +b)
//   with a * b = 64 for every bit position, so that  D[code][query] = sum A * B = 64 * (bits - 2 * hamming).
//   (The weights a_j = 1,1,2,4,8,16,32,64 make the expansion of a code two instructions per four bits - one LOP3, one
//   IMAD: ((x & (0x01010101 << j)) * M_j + C_j - instead of shift + permute + or; the queries carry 64 / a_j.)
//   A tile = 128 codes (MMA M) x the item's queries (MMA N = 16..256) x bits (K = 64 bytes per 64 code bits).
//
//   Epilogue: each thread owns one code (one TMEM lane), reads the row with packed 16-bit tcgen05.ld and keeps the row
//   maximum (VIMNMX3.S16x2, 1/4 instruction per code-query pair).  Only if max >= 64 * (bits - 2 * tau_max) - some
//   query of the item is within the largest threshold of the item - the code takes the exact path: XOR + POPC against
//   every query of the item, threshold test, first-discoverer de-duplication, append (bmih_append).  With k = 100 of
//   10^9 codes that is one row in thousands, so the answers are those of the POPC kernel, bit for bit.
//
// One persistent CTA per SM, warp-specialised, mbarrier pipelines:
//   producer (1 thread)   claims work items, streams the codes tile by tile into a raw ring (cp.async.bulk, 1 KB * W)
//   expanders (<= 8 warps) stage the item's queries (B operand + exact-path records); each takes every NE-th tile and
//                         expands its raw codes into an A stage of its own
//   MMA (1 thread)        per UNIT = up to G consecutive tiles of an item (G * queries <= one TMEM stage): K / 32
//                         tcgen05.mma per tile, ONE tcgen05.commit per unit
//   epilogue (8 warps)    two groups of four warps (TMEM lane quarters) take alternate units
//   relay (1 thread)      follows the units' completion barriers and hands their A stages back to the expanders (a second
//                         tcgen05.commit per unit in the MMA thread would do the same, for ~40 clocks of that thread)
// Measured on B200 (tools/tc_probe.cu):
//   - one 128 x N x 32-byte kind::i8 MMA takes max(45.5, N / 2) clocks: a 64-bit tile costs >= 91 clocks whatever the
//     number of queries up to 91, against 128 * N / 15.4 clocks on the POPC pipe;
//   - in the issuing thread a tcgen05.commit costs ~40 clocks and every mbarrier try_wait ~100 clocks that do NOT overlap
//     with the MMAs in flight (2 MMAs + commit: 138 clocks per loop; + one wait on a completed barrier: 237).  Hence the
//     units (one commit per G tiles), and the issuing thread never touches an mbarrier: it polls plain shared-memory
//     sequence words (ld.acquire) that the expanders / epilogue warps / producer publish with st.release / red.release.
#pragma once
#include "bmih.cuh"

namespace vc {

constexpr int kTcEpiWarps = 8;
constexpr int kTcTile = 128;                                         // codes per tile (MMA M)
constexpr uint32_t kTcCpi = 16384;                                   // codes per work item

// CS = accumulator columns per TMEM stage = most queries per work item (128: batched MIH, 256: large-batch scan)
template <int W, int CS> struct TcCfg {
  static_assert(CS == 128, "TMEM plan: 2 accumulator stages of 128 columns + 256 columns of A stages");
  static constexpr int QT = CS;
  static constexpr int KB = 64 * W;                  // expanded bytes per code / per query
  static constexpr int KSTEPS = KB / 32;             // MMAs per tile
  static constexpr uint32_t LBO = 128;               // B operand: K-adjacent core matrices (8 rows x 16 bytes) are contiguous
  static constexpr uint32_t SBO = (KB / 16) * 128;   // next group of 8 rows
  static constexpr int B_BYTES = QT * KB;
  static constexpr int RAW_BYTES = kTcTile * 8 * W;
  static constexpr int A_COLS = 16 * W;              // TMEM columns of one expanded tile (A operand): 64 W bytes per lane
  static constexpr int AS = 256 / A_COLS;            // A stages in TMEM: 16 / 8 / 4
  static constexpr int NE = 8;                       // expander warps: two sets of four (one warp per TMEM lane quarter)
  static constexpr int GM = AS / 2 > 8 ? 8 : AS / 2; // most tiles per unit
  static constexpr int THREADS = (kTcEpiWarps + NE + 3) * 32;     // + MMA warp + producer warp + relay warp
  static constexpr int NI = 4;                       // item slots (B operand, query records)
  static constexpr int RS = 32 / W;                  // raw stages: 32 KB of codes in flight per SM
  static constexpr int TS = 2;                       // accumulator stages
  static constexpr uint32_t A_COL0 = TS * CS;        // first TMEM column of the A stages
  static constexpr int QS = BmihCfg<W>::QS;          // u32 per staged query record (words, tau, pad)
  static constexpr int NBAR = 2 * RS + AS + TS + 3 * NI;
  static constexpr size_t OFF_B = 0;
  static constexpr size_t OFF_RAW = OFF_B + (size_t)NI * B_BYTES;
  static constexpr size_t OFF_QREC = OFF_RAW + (size_t)RS * RAW_BYTES;
  static constexpr size_t OFF_QID = OFF_QREC + (size_t)NI * QT * QS * 4;
  static constexpr size_t OFF_BAR = OFF_QID + (size_t)NI * QT * 4;
  static constexpr size_t OFF_ITEM = OFF_BAR + (size_t)NBAR * 8;
  static constexpr size_t OFF_FLAGS = OFF_ITEM + (size_t)NI * 32;          // a_cnt[AS] t_rel[TS] it_seq[NI] it_qc[NI] taumax[NI] tmem
  static constexpr size_t SMEM = OFF_FLAGS + (size_t)(AS + TS + 3 * NI + 4) * 4 + 1024;   // + alignment slack
  static_assert(AS % 2 == 0 && RS % 2 == 0, "the two expander sets take alternate tiles, stages and raw stages");
  static_assert(SMEM <= 227 * 1024, "shared memory");
};

struct TcItem { uint32_t t, c0, c1, qbeg, qn, a0, ntiles, npad /* = stride: columns per tile, a power of two >= 16 */; };
static_assert(sizeof(TcItem) == 32, "TcItem");

// ---- PTX ----------------------------------------------------------------------------------------------------------------
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the hint expires) instead
// of returning at once - polling loops that spin flat out take the issue slots the working warps need (measured:
// 1.5 instructions / clock / SM, two thirds of them polls, and every role of the pipeline 5x slower than on its own)
__device__ __forceinline__ bool tc_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
  return ok != 0;
}
__device__ __noinline__ void tc_timeout() {
  printf("bmih_verify_tc_kernel: wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
  __trap();
}
// bounded waits: a broken pipeline traps (the host sees a launch failure) instead of hanging the device
template <int SLEEP_NS>
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity) {
  if (tc_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  do {
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    if (++spins > (1u << 22)) tc_timeout();
  } while (!tc_try_wait(bar, parity));
}
// a whole warp waits: lane 0 polls, the others join once the phase is complete (one poll per warp instead of 32)
template <int SLEEP_NS>
__device__ __forceinline__ void tc_wait_warp(uint64_t* bar, uint32_t parity, uint32_t lane) {
  if (lane == 0) tc_wait<SLEEP_NS>(bar, parity);
  __syncwarp();
  if (lane != 0) tc_wait<0>(bar, parity);
}
// sequence words in shared memory, for the MMA thread (see the header)
__device__ __forceinline__ uint32_t tc_ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void tc_st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void tc_red_release(uint32_t* p, uint32_t v) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void tc_poll_eq(const uint32_t* p, uint32_t want) {
  uint32_t spins = 0;
  while (tc_ld_acquire(p) != want) { __nanosleep(20); if (++spins > (1u << 24)) tc_timeout(); }
}
__device__ __forceinline__ void tc_poll_ge(const uint32_t* p, uint32_t want) {
  uint32_t spins = 0;
  while ((int32_t)(tc_ld_acquire(p) - want) < 0) { __nanosleep(20); if (++spins > (1u << 24)) tc_timeout(); }
}
__device__ __forceinline__ uint32_t tc_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory: lane = code (row), 4 K-bytes per 32-bit column
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
#define VC_TC_ST16(taddr, r)                                                                                                         \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"             \
               ::"r"(taddr), "r"((r)[0]), "r"((r)[1]), "r"((r)[2]), "r"((r)[3]), "r"((r)[4]), "r"((r)[5]), "r"((r)[6]), "r"((r)[7]), "r"((r)[8]), \
                 "r"((r)[9]), "r"((r)[10]), "r"((r)[11]), "r"((r)[12]), "r"((r)[13]), "r"((r)[14]), "r"((r)[15]) : "memory")
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// shared-memory matrix descriptor: no swizzle, K-major, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::i8: D = S32, A and B signed 8-bit, both K-major, M = 128
__device__ __forceinline__ uint32_t tc_idesc(uint32_t n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(kTcTile >> 4) << 24);
}
// packed loads: register j = low 16 bits of columns (2j, 2j + 1) of this thread's TMEM lane
#define VC_TC_LD32P(r, taddr)                                                                                                        \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"  \
               "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                             \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                        \
               : "r"(taddr))
#define VC_TC_LD64P(r, taddr)                                                                                                        \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"  \
               "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,"    \
               "%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"                                                     \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),      \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),           \
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),          \
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),          \
                 "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),          \
                 "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),          \
                 "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),          \
                 "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])                        \
               : "r"(taddr))
#define VC_TC_LD8P(r, taddr)                                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                  \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr))
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bit -> byte expansion --------------------------------------------------------------------------------------------
// o[j] = bits j, j+8, j+16, j+24 of x as four bytes; K order (the same for codes and queries): word w, shift j, byte i
__device__ __forceinline__ void tc_expand_code_word(uint32_t x, uint32_t* o) {      // bytes +a_j / -a_j, a = 1,1,2,4,8,16,32,64
  o[0] = (x & 0x01010101u) * 254u + 0x01010101u;
  o[1] = (x & 0x02020202u) * 127u + 0x01010101u;
  o[2] = (x & 0x04040404u) * 63u + 0x02020202u;
  o[3] = (x & 0x08080808u) * 31u + 0x04040404u;
  o[4] = (x & 0x10101010u) * 15u + 0x08080808u;
  o[5] = (x & 0x20202020u) * 7u + 0x10101010u;
  o[6] = (x & 0x40404040u) * 3u + 0x20202020u;
  o[7] = (x & 0x80808080u) + 0x40404040u;
}
__device__ __forceinline__ void tc_expand_query_word(uint32_t x, uint32_t* o) {     // bytes +b_j / -b_j, b = 64 / a_j
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t b = j == 0 ? 64u : (128u >> j);
    o[j] = ((x >> j) & 0x01010101u) * (256u - 2u * b) + b * 0x01010101u;
  }
}
// one expanded row (code or query) into the canonical no-swizzle K-major layout
template <int W, uint32_t SBO>
__device__ __forceinline__ void tc_store_row(uint8_t* base, uint32_t row, const uint32_t* words, bool is_query) {
  uint8_t* rp = base + (row >> 3) * SBO + (row & 7) * 16;
#pragma unroll
  for (int w = 0; w < 2 * W; ++w) {
    uint32_t o[8];
    if (is_query) tc_expand_query_word(words[w], o); else tc_expand_code_word(words[w], o);
    *reinterpret_cast<uint4*>(rp + (2 * w) * 128) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(rp + (2 * w + 1) * 128) = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

// pass 1 of the epilogue: packed row maxima of the unit's accumulator rows, 64 columns at a time; tile i of the unit
// owns columns [i * STRIDE, (i + 1) * STRIDE).  Everything about the column -> tile mapping is static.
template <int CS, int GM, int STRIDE>
__device__ __forceinline__ void tc_pass1(uint32_t taddr, uint32_t ncols, uint32_t* tmaxp) {
  // 128 columns per load (one TMEM round trip of ~250 clocks each: the loads are latency-bound, not bandwidth-bound)
#pragma unroll
  for (int c = 0; c < CS / 128; ++c) {
    if ((c * 128) / STRIDE >= GM) break;                      // static: no unit reaches these columns
    if ((uint32_t)c * 128 < ncols) {
      uint32_t r[64];
      if ((uint32_t)c * 128 + 128 <= ncols) {
        VC_TC_LD64P(r, taddr + c * 128);
      } else {
        // 16 .. 112 columns left: the rest of the registers stay at the minimum
#pragma unroll
        for (int jj = 0; jj < 64; ++jj) r[jj] = 0x80008000u;
#pragma unroll
        for (int s16 = 0; s16 < 7; ++s16)
          if ((uint32_t)c * 128 + s16 * 16 < ncols) { VC_TC_LD8P((r + 8 * s16), taddr + c * 128 + s16 * 16); }
      }
      tc_ld_wait();
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) {
        uint32_t m = __vimax3_s16x2(r[8 * k8], r[8 * k8 + 1], r[8 * k8 + 2]);
        m = __vimax3_s16x2(m, r[8 * k8 + 3], r[8 * k8 + 4]);
        m = __vimax3_s16x2(m, r[8 * k8 + 5], r[8 * k8 + 6]);
        const int tile = (c * 128 + k8 * 16) / STRIDE;        // static after unrolling
        if (tile < GM) tmaxp[tile] = __vimax3_s16x2(tmaxp[tile], m, r[8 * k8 + 7]);
      }
    }
  }
}

#define VC_TC_TRACE(kind, idx, slot4) do { if (trace && (threadIdx.x & 31) == 0 && (uint32_t)((idx) - (kind == 2 ? 4096u : 2048u)) < 256u) trace[((kind) * 256 + ((idx) & 255u)) * 4 + (slot4)] = clock64(); } while (0)   // kinds 0 .. 2

// ---- the kernel ---------------------------------------------------------------------------------------------------------
// one hit of the exact path: code j is within tau of query q (both decided from the accumulator, which is exact)
template <int W>
__device__ __noinline__ void tc_hit(const BmihParams* pp, uint32_t t, const uint64_t* codes, uint32_t* qrec, const uint32_t* qids,
                                    uint32_t q, uint32_t dist, uint32_t j) {
  constexpr int QS = BmihCfg<W>::QS;
  CodeRegs<W> code;
  const uint64_t* cp = codes + (size_t)j * W;
#pragma unroll
  for (int i = 0; i < W; ++i) { const uint64_t v = cp[i]; code.w[2 * i] = (uint32_t)v; code.w[2 * i + 1] = (uint32_t)(v >> 32); }
  bmih_append<W>(pp, qids[q], t, dist, j, code, qrec + q * QS, qrec[q * QS + 2 * W]);
}

template <int W, int CS>
__global__ void __launch_bounds__(TcCfg<W, CS>::THREADS, 1) bmih_verify_tc_kernel(const __grid_constant__ BmihParams p) {
  using Cfg = TcCfg<W, CS>;
  constexpr int AS = Cfg::AS, NE = Cfg::NE, NI = Cfg::NI, RS = Cfg::RS, TS = Cfg::TS, QS = Cfg::QS, QT = Cfg::QT, GM = Cfg::GM;
  constexpr uint32_t BITS = 64 * W;
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw_ + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem + Cfg::OFF_B;
  uint8_t* sRaw = smem + Cfg::OFF_RAW;
  uint32_t* s_qrec = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_QREC);
  uint32_t* s_qid = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_QID);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* raw_full = bars;            uint64_t* raw_empty = raw_full + RS;
  uint64_t* a_empty = raw_empty + RS;   uint64_t* t_full = a_empty + AS;
  uint64_t* it_desc = t_full + TS;      uint64_t* it_q = it_desc + NI;     uint64_t* it_empty = it_q + NI;
  TcItem* s_item = reinterpret_cast<TcItem*>(smem + Cfg::OFF_ITEM);
  uint32_t* a_cnt = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_FLAGS);     // [AS] expander warps done with the stage (4 per use)
  uint32_t* t_rel = a_cnt + AS;                                             // [TS] epilogue warps done with the stage (4 per use)
  uint32_t* it_seq = t_rel + TS;                                            // [NI] item li published: li + 1
  uint32_t* it_qc = it_seq + NI;                                            // [NI] expander warps done staging (NE per use)
  volatile uint32_t* s_taumax = reinterpret_cast<volatile uint32_t*>(it_qc + NI);    // [NI]
  uint32_t* s_tmem = const_cast<uint32_t*>(s_taumax) + NI;
  __shared__ const uint64_t* s_codes[kMaxTables];

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < p.m) s_codes[tid] = p.tables[tid].codes;
  if (tid == 0) {
    for (int i = 0; i < RS; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 4); }
    for (int i = 0; i < AS; ++i) { mbar_init(&a_empty[i], 1); a_cnt[i] = 0; }
    for (int i = 0; i < TS; ++i) { mbar_init(&t_full[i], 1); t_rel[i] = 0; }
    for (int i = 0; i < NI; ++i) { mbar_init(&it_desc[i], 1); mbar_init(&it_q[i], NE); mbar_init(&it_empty[i], kTcEpiWarps); it_seq[i] = 0; it_qc[i] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t kMmaWarp = kTcEpiWarps + NE, kProdWarp = kMmaWarp + 1;
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t n_items = *p.n_items;
  long long* trace = blockIdx.x == 0 ? p.tc_trace : nullptr;

  if (warp == kProdWarp) {
    // ===== producer ===========================================================================================
    if (lane == 0) {
      uint32_t rg = 0;
      for (uint32_t li = 0;; ++li) {
        const uint32_t slot = li % NI;
        tc_wait<0>(&it_empty[slot], ((li / NI) & 1) ^ 1);
        const uint32_t idx = atomicAdd(p.item_cursor, 1u);
        TcItem d;
        d.qn = 0; d.t = 0; d.c0 = 0; d.c1 = 0; d.qbeg = 0; d.a0 = 0; d.ntiles = 0; d.npad = 16;
        if (idx < n_items) {
          const BmihItem bi = p.items[idx];
          d.t = bi.t; d.c0 = bi.c0; d.c1 = bi.c1; d.qbeg = bi.qbeg; d.qn = bi.qn;
          d.a0 = W == 1 ? (bi.c0 & ~1u) : bi.c0;                       // 16-byte aligned source
          d.ntiles = (bi.c1 - d.a0 + kTcTile - 1) / kTcTile;
          d.npad = max(16u, next_pow2(bi.qn));
        }
        s_item[slot] = d;
        s_taumax[slot] = 0;
        mbar_arrive(&it_desc[slot]);
        tc_st_release(&it_seq[slot], li + 1);
        if (d.qn == 0) break;
        const uint64_t* src = s_codes[d.t] + (size_t)d.a0 * W;
        for (uint32_t tile = 0; tile < d.ntiles; ++tile, ++rg) {
          const uint32_t rs = rg % RS;
          tc_wait<0>(&raw_empty[rs], ((rg / RS) & 1) ^ 1);
          const uint32_t ncodes = min((uint32_t)kTcTile, d.c1 - (d.a0 + tile * kTcTile));
          const uint32_t bytes = (ncodes * 8 * W + 15u) & ~15u;
          mbar_arrive_expect_tx(&raw_full[rs], bytes);
          bulk_g2s(sRaw + (size_t)rs * Cfg::RAW_BYTES, src + (size_t)tile * kTcTile * W, bytes, &raw_full[rs]);
        }
      }
    }
  } else if (warp == kProdWarp + 1) {
    // ===== relay: unit complete -> its A stages are free ======================================================
    if (lane == 0) {
      uint32_t g = 0, u = 0;
      for (uint32_t li = 0;; ++li) {
        const uint32_t slot = li % NI;
        tc_wait<0>(&it_desc[slot], (li / NI) & 1);
        const uint32_t qn = s_item[slot].qn, ntiles = s_item[slot].ntiles, stride = s_item[slot].npad;
        if (qn == 0) break;
        const uint32_t G = min((uint32_t)GM, (uint32_t)CS / stride);
        for (uint32_t tile0 = 0; tile0 < ntiles; tile0 += G, ++u) {
          const uint32_t gcount = min(G, ntiles - tile0);
          tc_wait<0>(&t_full[u % TS], (u / TS) & 1);
          for (uint32_t i = 0; i < gcount; ++i, ++g) mbar_arrive(&a_empty[g % AS]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer =========================================================================================
    // The whole warp runs the loop converged; every lane reads the (same) sequence words with plain volatile loads and
    // one elected lane issues.  Measured (tools/tc_probe.cu, 2 MMAs per unit): this costs ~20 clocks per unit over the
    // bare issue loop, where one lane polling in a divergent loop costs ~340 and an mbarrier try_wait ~100.
    {
      uint32_t g = 0, u = 0;
      const uint64_t b_desc0 = tc_desc(smem_u32(sB), Cfg::LBO, Cfg::SBO);
      const volatile uint32_t* v_a_cnt = a_cnt;
      const volatile uint32_t* v_t_rel = t_rel;
      const volatile uint32_t* v_it_seq = it_seq;
      const volatile uint32_t* v_it_qc = it_qc;
      const volatile TcItem* v_item = s_item;
      for (uint32_t li = 0;; ++li) {
        const uint32_t slot = li % NI;
        { uint32_t spins = 0; while (v_it_seq[slot] != li + 1) { __nanosleep(40); if (++spins > (1u << 24)) tc_timeout(); } }
        const uint32_t qn = v_item[slot].qn, ntiles = v_item[slot].ntiles, stride = v_item[slot].npad;
        if (qn == 0) break;
        { uint32_t spins = 0; while ((int32_t)(v_it_qc[slot] - NE * (li / NI + 1)) < 0) { __nanosleep(40); if (++spins > (1u << 24)) tc_timeout(); } }
        const uint32_t G = min((uint32_t)GM, (uint32_t)CS / stride);
        const uint32_t idesc = tc_idesc(stride);
        const uint64_t bd = b_desc0 + (uint64_t)((slot * Cfg::B_BYTES) >> 4);
        for (uint32_t tile0 = 0; tile0 < ntiles; tile0 += G, ++u) {
          const uint32_t gcount = min(G, ntiles - tile0);
          const uint32_t ts = u % TS;
          { uint32_t spins = 0; while ((int32_t)(v_t_rel[ts] - 4 * (u / TS)) < 0) { __nanosleep(40); if (++spins > (1u << 24)) tc_timeout(); } }
          VC_TC_TRACE(0, u, 0);
          for (uint32_t i = 0; i < gcount; ++i) {
            uint32_t spins = 0;
            while ((int32_t)(v_a_cnt[(g + i) % AS] - 4 * ((g + i) / AS + 1)) < 0) { __nanosleep(40); if (++spins > (1u << 24)) tc_timeout(); }
          }
          VC_TC_TRACE(0, u, 1);
          tc_fence_after();
          if (tc_elect_one()) {
            uint32_t dcol = tmem + ts * CS;
            for (uint32_t i = 0; i < gcount; ++i, dcol += stride) {
              const uint32_t at = tmem + Cfg::A_COL0 + ((g + i) % AS) * Cfg::A_COLS;
#pragma unroll
              for (int k = 0; k < Cfg::KSTEPS; ++k)
                tc_mma_i8_ts(dcol, at + 8 * k, bd + (uint64_t)((k * 2 * Cfg::LBO) >> 4), idesc, k > 0 ? 1u : 0u);
            }
            tc_commit(&t_full[ts]);
          }
          __syncwarp();
          VC_TC_TRACE(0, u, 3);
          g += gcount;
        }
      }
    }
  } else if (warp >= kTcEpiWarps) {
    // ===== expanders: two sets of four warps (one per TMEM lane quarter) take alternate tiles; lane = code ========
    const uint32_t we = warp - kTcEpiWarps;
    const uint32_t quarter = warp & 3, set = we >> 2;
    uint32_t g = 0;
    for (uint32_t li = 0;; ++li) {
      const uint32_t slot = li % NI, ph = (li / NI) & 1;
      tc_wait_warp<0>(&it_desc[slot], ph, lane);
      const TcItem d = s_item[slot];
      if (d.qn == 0) break;
      uint32_t* qrec = s_qrec + (size_t)slot * QT * QS;
      uint8_t* bB = sB + (size_t)slot * Cfg::B_BYTES;
      uint32_t tmax = 0;
      for (uint32_t row = we * 32 + lane; row < d.npad; row += NE * 32) {
        const uint32_t qid = p.qlist[d.qbeg + min(row, d.qn - 1)];      // pad rows repeat the last query
        uint32_t qw[2 * W];
#pragma unroll
        for (int i = 0; i < 2 * W; ++i) qw[i] = p.queries[(size_t)qid * 2 * W + i];
        if (row < d.qn) {
          const uint32_t tau = __ldcg(&p.gtau[qid]);
#pragma unroll
          for (int i = 0; i < 2 * W; ++i) qrec[row * QS + i] = qw[i];
          qrec[row * QS + 2 * W] = tau;
          s_qid[slot * QT + row] = qid;
          tmax = max(tmax, min(tau, BITS));
        }
        tc_store_row<W, Cfg::SBO>(bB, row, qw, true);
      }
      tmax = __reduce_max_sync(0xffffffffu, tmax);
      if (lane == 0) atomicMax(const_cast<uint32_t*>(&s_taumax[slot]), tmax);
      tc_proxy_fence();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&it_q[slot]); tc_red_release(&it_qc[slot], 1u); }
      for (uint32_t tile = 0; tile < d.ntiles; ++tile, ++g) {
        if ((g & 1) != set) continue;
        const uint32_t rs = g % RS, as = g % AS;
        if (we == 0 && lane == 0) VC_TC_TRACE(2, g, 0);
        tc_wait_warp<0>(&raw_full[rs], (g / RS) & 1, lane);
        if (we == 0 && lane == 0) VC_TC_TRACE(2, g, 1);
        uint32_t cw[2 * W];
        {
          const uint2* rp = reinterpret_cast<const uint2*>(sRaw + (size_t)rs * Cfg::RAW_BYTES) + (size_t)(quarter * 32 + lane) * W;
#pragma unroll
          for (int i = 0; i < W; ++i) { const uint2 v = rp[i]; cw[2 * i] = v.x; cw[2 * i + 1] = v.y; }
        }
        tc_proxy_fence();              // the raw stage is refilled by cp.async.bulk (async proxy): these generic-proxy reads first (scan.cuh)
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[rs]);
        uint32_t o[16 * W];
#pragma unroll
        for (int w = 0; w < 2 * W; ++w) tc_expand_code_word(cw[w], o + 8 * w);
        tc_wait_warp<0>(&a_empty[as], ((g / AS) & 1) ^ 1, lane);
        if (we == 0 && lane == 0) VC_TC_TRACE(2, g, 2);
        const uint32_t at = tmem + ((quarter * 32) << 16) + Cfg::A_COL0 + as * Cfg::A_COLS;
#pragma unroll
        for (int w = 0; w < W; ++w) VC_TC_ST16(at + 16 * w, (o + 16 * w));
        tc_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { tc_red_release(&a_cnt[as], 1u); if (we == 0) VC_TC_TRACE(2, g, 3); }
      }
    }
  } else {
    // ===== epilogue ===========================================================================================
    const uint32_t quarter = warp & 3, group = warp >> 2;
    const uint32_t lane_addr = (quarter * 32) << 16;
    uint32_t u = 0, g = 0, mine = 0;
    uint32_t st_units = 0, st_flagged = 0, st_hits = 0;
    for (uint32_t li = 0;; ++li) {
      const uint32_t slot = li % NI, ph = (li / NI) & 1;
      tc_wait_warp<0>(&it_desc[slot], ph, lane);
      const TcItem d = s_item[slot];
      if (d.qn == 0) break;
      tc_wait_warp<0>(&it_q[slot], ph, lane);
      uint32_t* qrec = s_qrec + (size_t)slot * QT * QS;
      const uint32_t* qids = s_qid + slot * QT;
      const uint64_t* tcodes = s_codes[d.t];
      const uint32_t stride = d.npad;
      const uint32_t G = min((uint32_t)GM, (uint32_t)CS / stride);
      for (uint32_t tile0 = 0; tile0 < d.ntiles; tile0 += G, ++u) {
        const uint32_t gcount = min(G, d.ntiles - tile0);
        g += gcount;
        if ((u & 1) != group) continue;
        const uint32_t ts = u % TS;
        // keep the staged thresholds current (other CTAs - and, sharded, other GPUs - lower them all the time)
        if (warp == 0 && (++mine & 7) == 0) {
          uint32_t m = 0;
          for (uint32_t q = lane; q < d.qn; q += 32) {
            const uint32_t f = __ldcg(&p.gtau[qids[q]]);
            const uint32_t old = atomicMin(&qrec[q * QS + 2 * W], f);
            m = max(m, min(min(f, old), BITS));
          }
          m = __reduce_max_sync(0xffffffffu, m);
          if (lane == 0) s_taumax[slot] = m;
        }
        tc_wait_warp<0>(&t_full[ts], (u / TS) & 1, lane);
        tc_fence_after();
        if (quarter == 0 && lane == 0) VC_TC_TRACE(1, u, 0);
        const uint32_t taddr = tmem + lane_addr + ts * CS;
        const uint32_t ncols = gcount * stride;
        // pass 1: packed row maxima, 64 columns at a time; tile i of the unit owns columns [i * stride, (i + 1) * stride)
        uint32_t tmaxp[GM];
#pragma unroll
        for (int i = 0; i < GM; ++i) tmaxp[i] = 0x80008000u;
        switch (stride) {
          case 16: tc_pass1<CS, GM, 16>(taddr, ncols, tmaxp); break;
          case 32: tc_pass1<CS, GM, 32>(taddr, ncols, tmaxp); break;
          case 64: tc_pass1<CS, GM, 64>(taddr, ncols, tmaxp); break;
          case 128: tc_pass1<CS, GM, 128>(taddr, ncols, tmaxp); break;
          default: if constexpr (CS >= 256) tc_pass1<CS, GM, 256>(taddr, ncols, tmaxp); break;
        }
        if (quarter == 0 && lane == 0) VC_TC_TRACE(1, u, 1);
        const int thr = 64 * ((int)BITS - 2 * (int)min(s_taumax[slot], BITS));
        const uint32_t jbase = d.a0 + tile0 * kTcTile + quarter * 32 + lane;
        uint32_t fmask = 0;                                     // bit i: my row of tile i is flagged
#pragma unroll
        for (int i = 0; i < GM; ++i) {
          const int rowmax = max((int)(int16_t)(tmaxp[i] & 0xFFFFu), (int)(int16_t)(tmaxp[i] >> 16));
          const uint32_t j = jbase + i * kTcTile;
          if ((uint32_t)i < gcount && rowmax >= thr && j >= d.c0 && j < d.c1) fmask |= 1u << i;
        }
        ++st_units;
        // exact path: some query of the item is within the item's largest threshold of a code.  That tile's accumulator
        // row is read again, 16 columns at a time by the whole warp; flagged lanes pick the columns that pass and turn D
        // back into the distance.  Up to two hits per lane are kept for after the TMEM stage has been released.
        uint32_t hq0 = 0, hd0 = 0, hj0 = 0, hq1 = 0, hd1 = 0, hj1 = 0, nh = 0;
        uint32_t anymask = __reduce_or_sync(0xffffffffu, fmask);
        if (anymask) {
          ++st_flagged;
          while (anymask) {
            const uint32_t i = __ffs(anymask) - 1;
            anymask &= anymask - 1;
            const bool flagged = (fmask >> i) & 1u;
            const uint32_t j = jbase + i * kTcTile;
            for (uint32_t c2 = 0; c2 < stride; c2 += 16) {
              uint32_t r[8];
              VC_TC_LD8P(r, taddr + i * stride + c2);
              tc_ld_wait();
              if (flagged) {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
                  for (int h = 0; h < 2; ++h) {
                    const int dv = h ? (int)(int16_t)(r[jj] >> 16) : (int)(int16_t)(r[jj] & 0xFFFFu);
                    if (dv >= thr) {
                      const uint32_t q = c2 + 2 * jj + h;
                      const uint32_t dist = (uint32_t)(((int)BITS - (dv >> 6)) >> 1);
                      if (q < d.qn && dist <= qrec[q * QS + 2 * W]) {
                        if (nh == 0) { hq0 = q; hd0 = dist; hj0 = j; }
                        else if (nh == 1) { hq1 = q; hd1 = dist; hj1 = j; }
                        else tc_hit<W>(&p, d.t, tcodes, qrec, qids, q, dist, j);
                        ++nh;
                      }
                    }
                  }
                }
              }
              __syncwarp();
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { tc_red_release(&t_rel[ts], 1u); if (quarter == 0) VC_TC_TRACE(1, u, 2); }
        st_hits += nh;
        if (nh > 0) tc_hit<W>(&p, d.t, tcodes, qrec, qids, hq0, hd0, hj0);
        if (nh > 1) tc_hit<W>(&p, d.t, tcodes, qrec, qids, hq1, hd1, hj1);
        __syncwarp();
        if (quarter == 0 && lane == 0) VC_TC_TRACE(1, u, 3);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&it_empty[slot]);
    }
    if (p.tc_stats) {
      st_hits = __reduce_add_sync(0xffffffffu, st_hits);
      if (lane == 0) { atomicAdd(&p.tc_stats[0], (unsigned long long)st_units); atomicAdd(&p.tc_stats[1], (unsigned long long)st_flagged); atomicAdd(&p.tc_stats[2], (unsigned long long)st_hits); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

}  // namespace vc
