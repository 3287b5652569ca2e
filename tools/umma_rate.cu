// tools/umma_rate.cu -- how fast can one SM issue tcgen05.mma kind::tf32 with BOTH operands in shared
// memory (SS mode) as a function of N?  Decides whether tap stacking along N (DESIGN.md section 7,
// tc::k_conv_stack) can pay off: at M=128, K=8 one UMMA reads 4 KB of A + N*32 B of B; if the SM's
// shared-memory operand path is the limit (128 B/clk), N=64 caps at 67 % of the tensor peak, N=128
// at 100 %, N=192/256 stay at 100 % with headroom.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate tools/umma_rate.cu && /tmp/umma_rate
//
// One CTA per SM; warp 0 allocates TMEM, one elected lane issues ITERS x 36 UMMAs (9 row-shifted A
// views x 4 k-steps of a 32-channel slab, like one pipeline stage of the convolution) back to back
// into one accumulator and commits once; the kernel time is taken with CUDA events.  No global
// memory traffic: operands are whatever shared memory holds (zero-filled).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k_umma_rate(int N, int iters, int two_issuers, int nacc, int blk, unsigned long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  // A: 8 channel groups x 200 rows x 16 B (one 32-channel slab with halo); B: 3 taps x 8 groups x N x 16 B (reused by the 9 views)
  const int a_rows = 200;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 8 * a_rows * 16;
  for (int i = tid; i < (8 * a_rows * 16 + 3 * 8 * N * 16) / 16; i += blockDim.x) ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t d_hi = (128u >> 4) | (1u << 14);
  const uint32_t a16 = smem_u32(sA) >> 4, b16 = smem_u32(sB) >> 4;
  const uint32_t a_lo_c = ((uint32_t)a_rows & 0x3fff) << 16, b_lo_c = ((uint32_t)N & 0x3fff) << 16;
  long long t0 = 0;
  if ((warp == 1 || (two_issuers && warp == 2)) && (tid & 31) == 0) {
    const int me = warp - 1;
    const uint32_t d0 = tmem + (uint32_t)(me * 256);         // each issuer its own accumulator(s)
    // nacc > 1: consecutive UMMAs of ONE issuer rotate over nacc accumulators (independent dependency chains)
    const uint32_t dstep = (nacc > 1) ? (uint32_t)(256 / nacc) : 0u;
    uint32_t cnt = 0;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const uint32_t a_t = a16 + 36 + (uint32_t)((t / 3 - 1) * 34 + (t % 3 - 1));   // row-shifted views, r = 32
        const uint32_t b_t = b16 + (t % 3) * 8 * N;
#pragma unroll
        for (int k2 = 0; k2 < 8; k2 += 2) {
          uint64_t ad = ((uint64_t)d_hi << 32) | a_lo_c | ((a_t + k2 * a_rows) & 0x3fff);
          uint64_t bd = ((uint64_t)d_hi << 32) | b_lo_c | ((b_t + k2 * N) & 0x3fff);
          uint32_t acc = (it | t | k2) ? 1u : 0u;
          const uint32_t d = d0 + ((cnt / (uint32_t)blk) % (uint32_t)nacc) * dstep;   // blk = 36: one accumulator per burst of 36 (block-sequential)
          ++cnt;
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                       ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[me])) : "memory");
    // wait for the MMAs to retire
    uint32_t ok = 0;
    for (long long spin = 0; !ok && spin < 400000000LL; ++spin) {          // bounded: never hang the box
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(ok) : "r"(smem_u32(&bar[me])), "r"(0u) : "memory");
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && me == 0) *cycles = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  CK(cudaFuncSetAttribute(k_umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  unsigned long long* d_cyc;
  CK(cudaMalloc(&d_cyc, 8));
  const int iters = 2000;
  printf("SMs %d, max clock %.0f MHz; M=128, K=8, kind::tf32, SS mode, %d x 36 UMMAs per issuer\n", sms, khz / 1000.0, iters);
  for (int blk : {1, 36})
  for (int nacc : {1, 2, 4})
  for (int two = 0; two <= 1; ++two)
    for (int N : {32, 64, 96, 128, 192, 256}) {
      if ((nacc > 1 && N * nacc > 256) || (blk > 1 && nacc == 1)) continue;                      // each issuer owns 256 TMEM columns
      size_t smem = 8 * 200 * 16 + (size_t)3 * 8 * N * 16;
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      k_umma_rate<<<sms, 128, smem>>>(N, 10, two, nacc, blk, d_cyc);           // warm-up
      CK(cudaEventRecord(e0));
      k_umma_rate<<<sms, 128, smem>>>(N, iters, two, nacc, blk, d_cyc);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      unsigned long long cyc = 0;
      CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
      double n_mma = (double)iters * 36 * (two ? 2 : 1);
      double flops = 2.0 * 128 * N * 8 * n_mma * sms;
      printf("rotate-every %2d  accumulators/issuer %d  issuers %d  N %3d : %8.3f ms  %7.1f TFLOP/s  %6.1f clk/UMMA (SM0)  operand bytes/clk %.0f\n", blk, nacc, two + 1, N, ms,
             flops / (ms * 1e-3) / 1e12, (double)cyc / (iters * 36.0), (4096.0 + 32.0 * N) * (two ? 2 : 1) / ((double)cyc / (iters * 36.0)));
    }
  return 0;
}
