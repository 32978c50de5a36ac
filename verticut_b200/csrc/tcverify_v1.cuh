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
//   expanders (4 warps)   stage the item's queries (B operand + exact-path records), expand raw codes into A stages
//   MMA (1 thread)        K / 32 tcgen05.mma per tile into one of 512 / QT TMEM stages, tcgen05.commit -> mbarriers
//   epilogue (8 warps)    two groups of four warps (TMEM lane quarters) take alternate tiles
// Measured on B200 (tools/tc_probe.cu): one 128 x N x 32-byte kind::i8 MMA takes max(45.5, N / 2) clocks, so a 64-bit
// tile costs >= 91 clocks whatever the number of queries up to 91, against 128 * N / 15.4 clocks on the POPC pipe.
#pragma once
#include "bmih.cuh"

namespace vc {

constexpr int kTc1EpiWarps = 8;
constexpr int kTc1ExpWarps = 4;
constexpr int kTc1Threads = (kTc1EpiWarps + kTc1ExpWarps + 2) * 32;     // + MMA warp + producer warp
constexpr int kTc1Tile = 128;                                         // codes per tile (MMA M)
constexpr uint32_t kTc1Cpi = 16384;                                   // codes per work item

template <int W, int QT> struct Tc1Cfg {
  static constexpr int KB = 64 * W;                  // expanded bytes per code / per query
  static constexpr int KSTEPS = KB / 32;             // MMAs per tile
  static constexpr uint32_t LBO = 128;               // K-adjacent core matrices (8 rows x 16 bytes) are contiguous
  static constexpr uint32_t SBO = (KB / 16) * 128;   // next group of 8 rows
  static constexpr int A_BYTES = kTc1Tile * KB;
  static constexpr int B_BYTES = QT * KB;
  static constexpr int RAW_BYTES = kTc1Tile * 8 * W;
  static constexpr int AS = W == 4 ? 2 : 4;          // A stages
  static constexpr int NI = QT > 64 ? 2 : 4;         // item slots (B operand, query records)
  static constexpr int RS = 32 / W;                  // raw stages: 32 KB of codes in flight per SM
  static constexpr int TS = 512 / QT;                // TMEM stages
  static constexpr int QS = BmihCfg<W>::QS;          // u32 per staged query record (words, tau, pad)
  static constexpr int NBAR = 2 * RS + 2 * AS + 2 * TS + 3 * NI;
  static constexpr size_t OFF_B = (size_t)AS * A_BYTES;
  static constexpr size_t OFF_RAW = OFF_B + (size_t)NI * B_BYTES;
  static constexpr size_t OFF_QREC = OFF_RAW + (size_t)RS * RAW_BYTES;
  static constexpr size_t OFF_QID = OFF_QREC + (size_t)NI * QT * QS * 4;
  static constexpr size_t OFF_BAR = OFF_QID + (size_t)NI * QT * 4;
  static constexpr size_t OFF_ITEM = OFF_BAR + (size_t)NBAR * 8;
  static constexpr size_t SMEM = OFF_ITEM + (size_t)NI * 48 + 64 + 1024;   // + alignment slack
  static_assert(TS >= 2 && TS % 2 == 0, "two epilogue groups take alternate TMEM stages");
};

struct Tc1Item { uint32_t t, c0, c1, qbeg, qn, a0, ntiles, npad; };
static_assert(sizeof(Tc1Item) == 32, "Tc1Item");

// ---- PTX ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool tc1_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a broken pipeline traps (the host sees a launch failure) instead of hanging the device
__device__ __forceinline__ void tc1_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!tc1_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) { printf("bmih_verify_tc1_kernel: barrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tc1_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc1_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc1_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc1_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc1_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor: no swizzle, K-major, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t tc1_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::i8: D = S32, A and B signed 8-bit, both K-major, M = 128
__device__ __forceinline__ uint32_t tc1_idesc(uint32_t n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((uint32_t)(kTc1Tile >> 4) << 24);
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
#define VC_TC_LD8P(r, taddr)                                                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                  \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr))
__device__ __forceinline__ void tc1_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- bit -> byte expansion --------------------------------------------------------------------------------------------
// o[j] = bits j, j+8, j+16, j+24 of x as four bytes; K order (the same for codes and queries): word w, shift j, byte i
__device__ __forceinline__ void tc1_expand_code_word(uint32_t x, uint32_t* o) {      // bytes +a_j / -a_j, a = 1,1,2,4,8,16,32,64
  o[0] = (x & 0x01010101u) * 254u + 0x01010101u;
  o[1] = (x & 0x02020202u) * 127u + 0x01010101u;
  o[2] = (x & 0x04040404u) * 63u + 0x02020202u;
  o[3] = (x & 0x08080808u) * 31u + 0x04040404u;
  o[4] = (x & 0x10101010u) * 15u + 0x08080808u;
  o[5] = (x & 0x20202020u) * 7u + 0x10101010u;
  o[6] = (x & 0x40404040u) * 3u + 0x20202020u;
  o[7] = (x & 0x80808080u) + 0x40404040u;
}
__device__ __forceinline__ void tc1_expand_query_word(uint32_t x, uint32_t* o) {     // bytes +b_j / -b_j, b = 64 / a_j
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t b = j == 0 ? 64u : (128u >> j);
    o[j] = ((x >> j) & 0x01010101u) * (256u - 2u * b) + b * 0x01010101u;
  }
}
// one expanded row (code or query) into the canonical no-swizzle K-major layout
template <int W, uint32_t SBO>
__device__ __forceinline__ void tc1_store_row(uint8_t* base, uint32_t row, const uint32_t* words, bool is_query) {
  uint8_t* rp = base + (row >> 3) * SBO + (row & 7) * 16;
#pragma unroll
  for (int w = 0; w < 2 * W; ++w) {
    uint32_t o[8];
    if (is_query) tc1_expand_query_word(words[w], o); else tc1_expand_code_word(words[w], o);
    *reinterpret_cast<uint4*>(rp + (2 * w) * 128) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(rp + (2 * w + 1) * 128) = make_uint4(o[4], o[5], o[6], o[7]);
  }
}

// ---- the kernel ---------------------------------------------------------------------------------------------------------
template <int W, int QT>
__global__ void __launch_bounds__(kTc1Threads, 1) bmih_verify_tc1_kernel(const __grid_constant__ BmihParams p) {
  using Cfg = Tc1Cfg<W, QT>;
  constexpr int AS = Cfg::AS, NI = Cfg::NI, RS = Cfg::RS, TS = Cfg::TS, QS = Cfg::QS;
  constexpr uint32_t BITS = 64 * W;
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw_ + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::OFF_B;
  uint8_t* sRaw = smem + Cfg::OFF_RAW;
  uint32_t* s_qrec = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_QREC);
  uint32_t* s_qid = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_QID);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* raw_full = bars;            uint64_t* raw_empty = raw_full + RS;
  uint64_t* a_full = raw_empty + RS;    uint64_t* a_empty = a_full + AS;
  uint64_t* t_full = a_empty + AS;      uint64_t* t_empty = t_full + TS;
  uint64_t* it_desc = t_empty + TS;     uint64_t* it_q = it_desc + NI;     uint64_t* it_empty = it_q + NI;
  Tc1Item* s_item = reinterpret_cast<Tc1Item*>(smem + Cfg::OFF_ITEM);
  volatile uint32_t* s_taumax = reinterpret_cast<volatile uint32_t*>(smem + Cfg::OFF_ITEM + NI * 32);    // [NI]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_ITEM + NI * 32 + NI * 4);
  __shared__ const uint64_t* s_codes[kMaxTables];

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < p.m) s_codes[tid] = p.tables[tid].codes;
  if (tid == 0) {
    for (int i = 0; i < RS; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], kTc1ExpWarps * 32); }
    for (int i = 0; i < AS; ++i) { mbar_init(&a_full[i], kTc1ExpWarps * 32); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < TS; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128); }
    for (int i = 0; i < NI; ++i) { mbar_init(&it_desc[i], 1); mbar_init(&it_q[i], kTc1ExpWarps * 32); mbar_init(&it_empty[i], 1 + kTc1EpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t kMmaWarp = kTc1EpiWarps + kTc1ExpWarps, kProdWarp = kMmaWarp + 1;
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc1_fence_before();
  __syncthreads();
  tc1_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t n_items = *p.n_items;

  if (warp == kProdWarp) {
    // ===== producer ===========================================================================================
    if (lane == 0) {
      uint32_t rg = 0;
      for (uint32_t li = 0;; ++li) {
        const uint32_t slot = li % NI;
        tc1_wait(&it_empty[slot], ((li / NI) & 1) ^ 1);
        const uint32_t idx = atomicAdd(p.item_cursor, 1u);
        Tc1Item d;
        d.qn = 0; d.t = 0; d.c0 = 0; d.c1 = 0; d.qbeg = 0; d.a0 = 0; d.ntiles = 0; d.npad = 0;
        if (idx < n_items) {
          const BmihItem bi = p.items[idx];
          d.t = bi.t; d.c0 = bi.c0; d.c1 = bi.c1; d.qbeg = bi.qbeg; d.qn = bi.qn;
          d.a0 = W == 1 ? (bi.c0 & ~1u) : bi.c0;                       // 16-byte aligned source
          d.ntiles = (bi.c1 - d.a0 + kTc1Tile - 1) / kTc1Tile;
          d.npad = (bi.qn + 15u) & ~15u;
        }
        s_item[slot] = d;
        s_taumax[slot] = 0;
        mbar_arrive(&it_desc[slot]);
        if (d.qn == 0) break;
        const uint64_t* src = s_codes[d.t] + (size_t)d.a0 * W;
        for (uint32_t tile = 0; tile < d.ntiles; ++tile, ++rg) {
          const uint32_t rs = rg % RS;
          tc1_wait(&raw_empty[rs], ((rg / RS) & 1) ^ 1);
          const uint32_t ncodes = min((uint32_t)kTc1Tile, d.c1 - (d.a0 + tile * kTc1Tile));
          const uint32_t bytes = (ncodes * 8 * W + 15u) & ~15u;
          mbar_arrive_expect_tx(&raw_full[rs], bytes);
          bulk_g2s(sRaw + (size_t)rs * Cfg::RAW_BYTES, src + (size_t)tile * kTc1Tile * W, bytes, &raw_full[rs]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===== MMA issuer =========================================================================================
    if (lane == 0) {
      uint32_t g = 0;
      for (uint32_t li = 0;; ++li) {
        const uint32_t slot = li % NI, ph = (li / NI) & 1;
        tc1_wait(&it_desc[slot], ph);
        const uint32_t qn = s_item[slot].qn, ntiles = s_item[slot].ntiles, npad = s_item[slot].npad;
        if (qn == 0) break;
        tc1_wait(&it_q[slot], ph);
        const uint32_t idesc = tc1_idesc(npad);
        const uint32_t b_addr = smem_u32(sB + (size_t)slot * Cfg::B_BYTES);
        for (uint32_t tile = 0; tile < ntiles; ++tile, ++g) {
          const uint32_t as = g % AS, ts = g % TS;
          tc1_wait(&a_full[as], (g / AS) & 1);
          tc1_wait(&t_empty[ts], ((g / TS) & 1) ^ 1);
          tc1_fence_after();
          const uint32_t a_addr = smem_u32(sA + (size_t)as * Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < Cfg::KSTEPS; ++k)
            tc1_mma_i8(tmem + ts * QT, tc1_desc(a_addr + k * 2 * Cfg::LBO, Cfg::LBO, Cfg::SBO), tc1_desc(b_addr + k * 2 * Cfg::LBO, Cfg::LBO, Cfg::SBO),
                      idesc, k > 0 ? 1u : 0u);
          tc1_commit(&a_empty[as]);
          tc1_commit(&t_full[ts]);
        }
        tc1_commit(&it_empty[slot]);
      }
    }
  } else if (warp >= kTc1EpiWarps) {
    // ===== expanders ==========================================================================================
    const uint32_t e = tid - kTc1EpiWarps * 32;
    uint32_t g = 0;
    for (uint32_t li = 0;; ++li) {
      const uint32_t slot = li % NI, ph = (li / NI) & 1;
      tc1_wait(&it_desc[slot], ph);
      const Tc1Item d = s_item[slot];
      if (d.qn == 0) break;
      uint32_t* qrec = s_qrec + (size_t)slot * QT * QS;
      uint8_t* bB = sB + (size_t)slot * Cfg::B_BYTES;
      uint32_t tmax = 0;
      for (uint32_t row = e; row < d.npad; row += kTc1ExpWarps * 32) {
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
      tc1_proxy_fence();
      mbar_arrive(&it_q[slot]);
      for (uint32_t tile = 0; tile < d.ntiles; ++tile, ++g) {
        const uint32_t rs = g % RS, as = g % AS;
        tc1_wait(&raw_full[rs], (g / RS) & 1);
        uint32_t cw[2 * W];
        {
          const uint2* rp = reinterpret_cast<const uint2*>(sRaw + (size_t)rs * Cfg::RAW_BYTES) + (size_t)e * W;
#pragma unroll
          for (int i = 0; i < W; ++i) { const uint2 v = rp[i]; cw[2 * i] = v.x; cw[2 * i + 1] = v.y; }
        }
        tc1_proxy_fence();             // the raw stage is refilled by cp.async.bulk (async proxy): these generic-proxy reads first (scan.cuh)
        mbar_arrive(&raw_empty[rs]);
        tc1_wait(&a_empty[as], ((g / AS) & 1) ^ 1);
        tc_store_row<W, Cfg::SBO>(sA + (size_t)as * Cfg::A_BYTES, e, cw, false);
        tc1_proxy_fence();
        mbar_arrive(&a_full[as]);
      }
    }
  } else {
    // ===== epilogue ===========================================================================================
    const uint32_t quarter = warp & 3, group = warp >> 2;
    const uint32_t lane_addr = (quarter * 32) << 16;
    uint32_t g = 0;
    for (uint32_t li = 0;; ++li) {
      const uint32_t slot = li % NI, ph = (li / NI) & 1;
      tc1_wait(&it_desc[slot], ph);
      const Tc1Item d = s_item[slot];
      if (d.qn == 0) break;
      tc1_wait(&it_q[slot], ph);
      uint32_t* qrec = s_qrec + (size_t)slot * QT * QS;
      const uint32_t* qids = s_qid + slot * QT;
      uint32_t mine = 0;
      for (uint32_t tile = 0; tile < d.ntiles; ++tile, ++g) {
        if ((g & 1) != group) continue;
        const uint32_t ts = g % TS;
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
        tc1_wait(&t_full[ts], (g / TS) & 1);
        tc1_fence_after();
        const uint32_t taddr = tmem + lane_addr + ts * QT;
        uint32_t acc = 0x80008000u;
        uint32_t c = 0;
        for (; c + 64 <= d.npad; c += 64) {
          uint32_t r[32];
          VC_TC_LD32P(r, taddr + c);
          tc1_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) acc = __vimax3_s16x2(acc, r[j], r[j + 1]);
        }
        for (; c < d.npad; c += 16) {
          uint32_t r[8];
          VC_TC_LD8P(r, taddr + c);
          tc1_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; j += 2) acc = __vimax3_s16x2(acc, r[j], r[j + 1]);
        }
        tc1_fence_before();
        mbar_arrive(&t_empty[ts]);
        const int rowmax = max((int)(int16_t)(acc & 0xFFFFu), (int)(int16_t)(acc >> 16));
        const int thr = 64 * ((int)BITS - 2 * (int)min(s_taumax[slot], BITS));
        const uint32_t j = d.a0 + tile * kTc1Tile + quarter * 32 + lane;
        if (rowmax >= thr && j >= d.c0 && j < d.c1) {
          // exact path: this code is within the item's largest threshold of at least one of its queries
          CodeRegs<W> code;
          const uint64_t* cp = s_codes[d.t] + (size_t)j * W;
#pragma unroll
          for (int i = 0; i < W; ++i) { const uint64_t v = cp[i]; code.w[2 * i] = (uint32_t)v; code.w[2 * i + 1] = (uint32_t)(v >> 32); }
          for (uint32_t q = 0; q < d.qn; ++q) {
            const QRec<W> cur = load_qrec<W, QS>(qrec, q);
            const uint32_t dist = hamming_exact<W>(code.w, cur.qw);
            if (dist <= cur.tau) bmih_append<W>(&p, qids[q], d.t, dist, j, code, qrec + q * QS, cur.tau);
          }
        }
        __syncwarp();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&it_empty[slot]);
    }
  }
  tc1_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

}  // namespace vc
