// Peer exchange between the id-shards of one box over NVLink / NVSwitch peer memory.
//
// Replaces, for the sharded search, the reference's per-radius MPI_Gatherv of all candidates + MPI_Bcast of the stop flag
// (src/search_worker.cc:177,207, src/mpi_coordinator.cc:26-69) - and this project's own first version of it, an NCCL
// all-reduce per search step plus an NCCL all-gather of the results - by stores into the other GPUs' memory:
//
//   every rank owns a WINDOW (device memory, opened by all peers through CUDA IPC): two sets of G slots plus G flags per set.
//   Exchange number s (1, 2, ... - all ranks run the same sequence) uses set s & 1:
//     push   the producing kernel itself (bmih_settle_kernel: histogram rows, bmih_finish_kernel: top-k rows), or
//            xchg_push_kernel for anything else, writes this rank's payload into slot [rank] of EVERY rank's window with plain
//            16-byte stores; the last CTA to finish fences (system scope) and stores s into flag [rank] of every window
//            (st.release.sys);
//     wait   the consumer spins on its own window's G flags (ld.acquire.sys) until all show s, then reads the G slots:
//            xchg_sum_kernel (all-reduce of histograms) or the merge kernel (all-gather of top-k lists).
//   Set s & 1 is written again by exchange s + 2, which a rank starts only after it finished exchange s + 1 - and that needed
//   every peer's flag s + 1, which a peer stores (stream order) after it consumed exchange s.  So two sets are enough.
//
// No NCCL kernel, no host round trip and no extra copy sits between "local result ready" and "peer can merge"; the transfer is
// the producing kernel's own stores.  A peer that never arrives (crashed rank) ends the spin after kXchgTimeoutNs with an error
// flag instead of a hang.
#pragma once
#include "common.cuh"

namespace vc {

constexpr int kXchgMaxWorld = 16;
constexpr unsigned long long kXchgTimeoutNs = 20ull * 1000 * 1000 * 1000;
constexpr uint32_t kXchgFlagStride = 8;        // u32 per flag (32 bytes: one sector each)
constexpr uint64_t kXchgHeaderBytes = 2 * kXchgMaxWorld * kXchgFlagStride * 4;

// what a kernel needs to push into, or wait on, the windows
struct XchgDev {
  unsigned char* peer[kXchgMaxWorld];   // window base of every rank (peer[rank] = the local one); world == 0: no exchange
  uint32_t rank, world;
  uint32_t seq;                         // number of this exchange
  uint64_t slot_off;                    // byte offset of slot 0 of this exchange's set in a window
  uint64_t stride;                      // bytes between the slots of this exchange
  uint32_t* done;                       // [1] local counter of CTAs that finished a push (running total over all exchanges)
  uint32_t done_target;                 // ... and its value once the last CTA of this push has been counted
  uint32_t* err;                        // [1] local: set to 1 if a wait timed out
};

__device__ __forceinline__ uint32_t* xchg_flag(unsigned char* window, uint32_t seq, uint32_t src_rank) {
  return reinterpret_cast<uint32_t*>(window) + ((seq & 1u) * kXchgMaxWorld + src_rank) * kXchgFlagStride;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Called by every thread of a CTA after its own stores into the peers' slots: the last CTA of the grid publishes the
// exchange (flag [rank] = seq in every window).  Every CTA the host counted into done_target must call this exactly once.
__device__ __forceinline__ void xchg_publish(const XchgDev& x) {
  __threadfence_system();                              // this thread's slot stores are visible system-wide ...
  __syncthreads();
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0) s_last = atomicAdd(x.done, 1u) + 1 == x.done_target ? 1u : 0u;       // ... before the CTA is counted
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    for (uint32_t g = threadIdx.x; g < x.world; g += blockDim.x) st_release_sys(xchg_flag(x.peer[g], x.seq, x.rank), x.seq);
  }
}

// generic push: n16 16-byte units from src into slot [rank] of every window
__global__ void __launch_bounds__(256) xchg_push_kernel(const XchgDev x, const uint4* __restrict__ src, uint64_t n16) {
  const uint64_t off = x.slot_off + (uint64_t)x.rank * x.stride;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
    for (uint32_t g = 0; g < x.world; ++g) reinterpret_cast<uint4*>(x.peer[g] + off)[i] = v;
  }
  xchg_publish(x);
}

// every thread that is going to read the slots calls this first (thread g < world polls flag g; the CTA then syncs)
__device__ __forceinline__ void xchg_wait_cta(const XchgDev& x) {
  if (threadIdx.x < x.world) {
    const uint32_t* f = xchg_flag(x.peer[x.rank], x.seq, threadIdx.x);
    if (ld_acquire_sys(f) != x.seq) {
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys(f) != x.seq) {
        __nanosleep(200);
        if (global_timer_ns() - t0 > kXchgTimeoutNs) { atomicExch(x.err, 1u); break; }
      }
    }
  }
  __syncthreads();
}

// all-reduce (sum of u32) of the G slots into dst, n words
__global__ void __launch_bounds__(256) xchg_sum_kernel(const XchgDev x, uint32_t* __restrict__ dst, uint64_t n_words) {
  xchg_wait_cta(x);
  const unsigned char* base = x.peer[x.rank] + x.slot_off;
  const uint64_t n4 = n_words / 4;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
    uint4 a = make_uint4(0, 0, 0, 0);
    for (uint32_t g = 0; g < x.world; ++g) {
      const uint4 v = __ldcg(reinterpret_cast<const uint4*>(base + (uint64_t)g * x.stride) + i);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<uint4*>(dst)[i] = a;
  }
  for (uint64_t i = n4 * 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t a = 0;
    for (uint32_t g = 0; g < x.world; ++g) a += __ldcg(reinterpret_cast<const uint32_t*>(base + (uint64_t)g * x.stride) + i);
    dst[i] = a;
  }
}

// all-gather: nothing to compute - one small CTA waits, the consumer (merge kernel) follows in stream order
__global__ void xchg_wait_kernel(const XchgDev x) { xchg_wait_cta(x); }

}  // namespace vc
