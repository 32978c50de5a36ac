// Integer-pipe micro-benchmark for the roofline denominators of the scan / verify kernels (B200, sm_100a):
// sustained POPC, LOP3 and the per-pair instruction mixes of the distance loops, in ops/clk/SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, int iters) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
  uint32_t q0 = seed ^ 0x12345678u, q1 = seed * 77u;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i] = __popc(a[i]) + a[i] * 3u; }                 // POPC + IMAD (fma pipe) dependent chain x8
      else if (MODE == 1) { acc += __popc(a[i] ^ q0); a[i] += acc; }        // xor + popc + add
      else if (MODE == 2) { acc = min(acc + 0u, (uint32_t)__popc((a[i] ^ q0) | ((a[i] * 5u) ^ q1))); a[i] ^= it; }   // prefilter mix
      else if (MODE == 3) { a[i] = (a[i] ^ q0) | (a[(i + 1) & 7] & q1); }   // LOP3 only
      else if (MODE == 4) { acc += __popc(a[i] ^ q0) + __popc(a[(i + 1) & 7] ^ q1); a[i] += it; }  // exact 64-bit distance mix
    }
    q0 += 0x01010101u;
  }
  uint32_t r = acc;
#pragma unroll
  for (int i = 0; i < 8; ++i) r ^= a[i];
  if (r == 0xdeadbeef) out[0] = r;
}

template <int MODE>
void run(const char* name, double ops_per_inner) {
  uint32_t* d;
  cudaMalloc(&d, 4);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int iters = 20000, blocks = sms * 8;
  k<MODE><<<blocks, 256>>>(d, 3, 100);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(d, 3, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double inner = (double)blocks * 256 * iters * 8;
  double per_s = inner * ops_per_inner / (ms * 1e-3);
  printf("%-28s %8.3f ms  %.3e ops/s  = %.2f ops/clk/SM at max clock %d MHz\n", name, ms, per_s,
         per_s / sms / (clk_khz * 1e3), clk_khz / 1000);
  cudaFree(d);
}

int main() {
  run<0>("popc+imad chain", 1);
  run<1>("xor+popc+add (1 popc)", 1);
  run<2>("prefilter pair (1 popc)", 1);
  run<3>("lop3 only", 1);
  run<4>("exact pair (2 popc)", 2);
  return 0;
}
