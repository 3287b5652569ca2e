// tools/umma_issue.cu -- how many cycles does ONE warp need to ISSUE a tcgen05.mma (kind::tf32, SS mode)?
// Round-2 finding (tools/umma_rate.cu): a single issuing thread tops out at one UMMA per ~64 clk whatever N is,
// i.e. the convolution's issuers are instruction-issue bound (per UMMA: ELECT + 5 R2UR.BROADCAST + 2 VOTEU in the
// SASS), not tensor-pipe bound.  This benchmark compares issue-loop formulations on the convolution's own
// stage shape (9 row-shifted A views x 4 k-steps of a 32-channel slab, N = 64):
//   v0  one asm statement per UMMA under `if (lane == 0)`                      (round-1 microbenchmark)
//   v1  one asm statement per UMMA, warp-uniform with elect.sync inside        (round-1 convolution)
//   v2  ONE asm block per stage: elect once, all 36 descriptors derived by in-asm adds from two bases
//   v3  as v2 but the whole block sits under `if (elect_one)`: no predicate per instruction
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_issue tools/umma_issue.cu && /tmp/umma_issue
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one UMMA inside an asm block: descriptors {lo, %5}; registers ad/bd/q/pt declared by the enclosing block
#define MMA1(AL, BL, PRED) "mov.b64 ad, {" AL ", %5};\n mov.b64 bd, {" BL ", %5};\n @q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, " PRED ";\n"
#define TAP4(TOP, FIRSTPRED)                                                   \
  "add.u32 a0, %1, " TOP ";\n" MMA1("a0", "bt", FIRSTPRED)                     \
  "add.u32 a1, a0, %6;\n add.u32 b1, bt, %7;\n" MMA1("a1", "b1", "pt")         \
  "add.u32 a2, a1, %6;\n add.u32 b2, b1, %7;\n" MMA1("a2", "b2", "pt")         \
  "add.u32 a3, a2, %6;\n add.u32 b3, b2, %7;\n" MMA1("a3", "b3", "pt")         \
  "add.u32 bt, bt, %8;\n"
#define MMA1U(AL, BL, PRED) "mov.b64 ad, {" AL ", %5};\n mov.b64 bd, {" BL ", %5};\n tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, " PRED ";\n"
#define TAP4U(TOP, FIRSTPRED)                                                  \
  "add.u32 a0, %1, " TOP ";\n" MMA1U("a0", "bt", FIRSTPRED)                    \
  "add.u32 a1, a0, %6;\n add.u32 b1, bt, %7;\n" MMA1U("a1", "b1", "pt")        \
  "add.u32 a2, a1, %6;\n add.u32 b2, b1, %7;\n" MMA1U("a2", "b2", "pt")        \
  "add.u32 a3, a2, %6;\n add.u32 b3, b2, %7;\n" MMA1U("a3", "b3", "pt")        \
  "add.u32 bt, bt, %8;\n"

// 36 UMMAs of one stage (9 taps x 4 k-steps), warp-uniform, one elect
__device__ __forceinline__ void stage36_block(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc, uint32_t hi,
                                              uint32_t a_k2, uint32_t b_k2, uint32_t b_tap, const int* to) {
  asm volatile(
      "{\n.reg .pred q, p, pt;\n.reg .b32 a0, a1, a2, a3, b1, b2, b3, bt;\n.reg .b64 ad, bd;\n"
      "elect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\nsetp.eq.b32 pt, 0, 0;\nmov.b32 bt, %2;\n"
      TAP4("%9", "p") TAP4("%10", "pt") TAP4("%11", "pt") TAP4("%12", "pt") TAP4("%13", "pt") TAP4("%14", "pt") TAP4("%15", "pt")
      TAP4("%16", "pt") TAP4("%17", "pt")
      "}\n"
      ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(hi), "r"(a_k2), "r"(b_k2), "r"(b_tap),
        "r"(to[0]), "r"(to[1]), "r"(to[2]), "r"(to[3]), "r"(to[4]), "r"(to[5]), "r"(to[6]), "r"(to[7]), "r"(to[8]) : "memory");
}
// the same, to be called by ONE thread (no predicates on the MMAs)
__device__ __forceinline__ void stage36_single(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc, uint32_t hi,
                                               uint32_t a_k2, uint32_t b_k2, uint32_t b_tap, const int* to) {
  asm volatile(
      "{\n.reg .pred p, pt;\n.reg .b32 a0, a1, a2, a3, b1, b2, b3, bt;\n.reg .b64 ad, bd;\n"
      "setp.ne.b32 p, %4, 0;\nsetp.eq.b32 pt, 0, 0;\nmov.b32 bt, %2;\n"
      TAP4U("%9", "p") TAP4U("%10", "pt") TAP4U("%11", "pt") TAP4U("%12", "pt") TAP4U("%13", "pt") TAP4U("%14", "pt") TAP4U("%15", "pt")
      TAP4U("%16", "pt") TAP4U("%17", "pt")
      "}\n"
      ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(hi), "r"(a_k2), "r"(b_k2), "r"(b_tap),
        "r"(to[0]), "r"(to[1]), "r"(to[2]), "r"(to[3]), "r"(to[4]), "r"(to[5]), "r"(to[6]), "r"(to[7]), "r"(to[8]) : "memory");
}

__device__ __forceinline__ void umma_w(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, q;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
               "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_1(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

struct TapOff { int v[9]; };

__global__ void __launch_bounds__(128, 1) k_issue(int N, int iters, int variant, int issuers, TapOff to, unsigned long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a_rows = 200;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 8 * a_rows * 16;
  for (int i = tid; i < (8 * a_rows * 16 + 9 * 8 * N * 16) / 16; i += blockDim.x) ((float4*)smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
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
  const uint32_t a16 = (smem_u32(sA) >> 4) + 36, b16 = smem_u32(sB) >> 4;
  const uint32_t a_lo_c = ((uint32_t)a_rows & 0x3fff) << 16, b_lo_c = ((uint32_t)N & 0x3fff) << 16;
  if (warp >= 1 && warp <= issuers) {
    const int me = warp - 1;
    const uint32_t d = tmem + (uint32_t)(me * 256);
    long long t0 = clock64();
    if (variant == 0) {
      if (lane == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t a_t = a16 + (uint32_t)to.v[t], b_t = b16 + t * 8 * N;
#pragma unroll
            for (int k2 = 0; k2 < 8; k2 += 2) {
              uint64_t ad = ((uint64_t)d_hi << 32) | a_lo_c | ((a_t + k2 * a_rows) & 0x3fff);
              uint64_t bd = ((uint64_t)d_hi << 32) | b_lo_c | ((b_t + k2 * N) & 0x3fff);
              umma_1(d, ad, bd, idesc, (it | t | k2) ? 1u : 0u);
            }
          }
        }
      }
    } else if (variant == 1) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint32_t a_t = a16 + (uint32_t)to.v[t], b_t = b16 + t * 8 * N;
#pragma unroll
          for (int k2 = 0; k2 < 8; k2 += 2) {
            uint64_t ad = ((uint64_t)d_hi << 32) | a_lo_c | ((a_t + k2 * a_rows) & 0x3fff);
            uint64_t bd = ((uint64_t)d_hi << 32) | b_lo_c | ((b_t + k2 * N) & 0x3fff);
            umma_w(d, ad, bd, idesc, (it | t | k2) ? 1u : 0u);
          }
        }
      }
    } else if (variant == 2) {
      for (int it = 0; it < iters; ++it)
        stage36_block(d, a_lo_c | a16, b_lo_c | b16, idesc, it ? 1u : 0u, d_hi, 2u * a_rows, 2u * N, 8u * N, to.v);
    } else {
      if (lane == 0)
        for (int it = 0; it < iters; ++it)
          stage36_single(d, a_lo_c | a16, b_lo_c | b16, idesc, it ? 1u : 0u, d_hi, 2u * a_rows, 2u * N, 8u * N, to.v);
    }
    __syncwarp();
    if (lane == 0) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[me])) : "memory");
      uint32_t ok = 0;
      for (long long spin = 0; !ok && spin < 400000000LL; ++spin)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(&bar[me])), "r"(0u) : "memory");
      long long t1 = clock64();
      if (blockIdx.x == 0 && me == 0) *cycles = (unsigned long long)(t1 - t0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  int sms = 0;
  CK(cudaSetDevice(0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(k_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  unsigned long long* d_cyc;
  CK(cudaMalloc(&d_cyc, 8));
  TapOff to;
  for (int t = 0; t < 9; ++t) to.v[t] = (t / 3 - 1) * 34 + (t % 3 - 1);
  const int iters = 2000;
  printf("issue-loop formulations; M=128, K=8, kind::tf32, SS mode, %d x 36 UMMAs per issuer; tensor time per UMMA = N/2 clk\n", iters);
  for (int N : {64, 128, 192})
    for (int issuers = 1; issuers <= 2; ++issuers)
      for (int variant = 0; variant < 4; ++variant) {
        size_t smem = 8 * 200 * 16 + (size_t)9 * 8 * N * 16;
        if (smem > 200 * 1024) continue;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        k_issue<<<sms, 128, smem>>>(N, 10, variant, issuers, to, d_cyc);
        CK(cudaEventRecord(e0));
        k_issue<<<sms, 128, smem>>>(N, iters, variant, issuers, to, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long cyc = 0;
        CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
        double n_mma = (double)iters * 36 * issuers;
        printf("N %3d  issuers %d  v%d : %8.3f ms  %7.1f TFLOP/s  %6.1f clk per UMMA per issuer (SM0)\n", N, issuers, variant, ms,
               2.0 * 128 * N * 8 * n_mma * sms / (ms * 1e-3) / 1e12, (double)cyc / (iters * 36.0));
      }
  return 0;
}
