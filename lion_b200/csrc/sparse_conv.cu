// lion_b200 -- first half of the SPARSE first convolution of a PVConv (sm_100a).
//
// A PVConv's first 3x3x3 convolution reads a voxel grid with at most N occupied voxels (2048 of 32768 at r = 32).
// Instead of 27 dense taps over all r^3 positions, the occupied voxels' features x[v] (compact list, k_scatter_compact)
// are multiplied by all 27 taps at once,
//     y[v][t * C + c] = sum_ci  W[t][ci][c] * x[v][ci]                       (this file: one GEMM, K = Cin, 27*C columns)
// and every output voxel then sums the rows y[v(p + off(t))][t] of its occupied neighbours (k_sparse_conv_gather,
// packed_kernels.cuh).  16x fewer FLOPs at r = 32; the cost is streaming y once out and once in.
//
// GEMM orientation: D^T = W^T x^T -- the WEIGHTS are the UMMA A operand (M = 128 tap-channels n), the voxels the B
// operand (N = 256 voxel rows, K-major = the packed [C/4][rows][4] layout as it lies in HBM).  TMEM lanes are then
// tap-channels and TMEM columns voxels, so a tcgen05.ld register holds one voxel for 32 consecutive n across the warp:
// every store instruction writes 128 contiguous bytes of a y row.  (The convolution kernel's own epilogue, lanes = rows,
// wrote 16-byte pieces 6.9 KB apart: 220 us for the same 453 MB.)
// One CTA = 256 threads, one n-tile, a strided range of 256-row blocks; operands by cp.async.bulk; the next block's
// voxels are requested as soon as the MMAs of the current one retire, i.e. under the epilogue; 2 CTAs per SM.
#include "common.cuh"
#include "model.cuh"
#include <cstdlib>

namespace lion {
namespace spc {

constexpr int ROWS = 256;     // voxel rows per block (UMMA N)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: traps instead of hanging
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}

struct Params {
  const float4* x;       // compact voxel features, PF [B][G][N]
  const float* w;        // conv_tc packing of the wide 1x1 convolution: [n-tile][G][128][4], tf32-rounded
  float* y;              // [B][N][ld]
  const int* nocc;       // [B] occupied voxels per shape (rows beyond are never read back: skipped)
  int G, N, B, ld, blocks_per_shape, nblocks;
};

__global__ void __launch_bounds__(256, 2) k_ygemm(Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  float4* sW = (float4*)smem;                              // [G][128]
  float4* sX = sW + (size_t)P.G * 128;                     // [G][ROWS]
  uint64_t* bars = (uint64_t*)(sX + (size_t)P.G * ROWS);
  uint32_t* s_tmem = (uint32_t*)(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = blockIdx.x;
  const uint32_t bar_w = smem_u32(bars), bar_x = smem_u32(bars + 1), bar_mma = smem_u32(bars + 2);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_w));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_x));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_mma));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;

  // the first block this CTA owns that holds occupied voxels
  auto block_rows = [&](int rb, int& b, int& r0) -> int {       // rows of block rb that matter (0 = skip)
    b = rb / P.blocks_per_shape;
    r0 = (rb % P.blocks_per_shape) * ROWS;
    int n = min(P.nocc ? __ldg(P.nocc + b) : P.N, P.N) - r0;
    return n < 0 ? 0 : (n > ROWS ? ROWS : n);
  };
  auto load_x = [&](int b, int r0, int nrows) {                 // thread 0 only
    const uint32_t bytes = (uint32_t)nrows * 16u;
    mbar_expect_tx(bar_x, bytes * (uint32_t)P.G);
    for (int g = 0; g < P.G; ++g)
      bulk_g2s(smem_u32(sX + (size_t)g * ROWS), P.x + ((size_t)b * P.G + g) * P.N + r0, bytes, bar_x);
  };
  int rb = blockIdx.y;
  int b = 0, r0 = 0, nrows = 0;
  while (rb < P.nblocks && (nrows = block_rows(rb, b, r0)) == 0) rb += gridDim.y;
  if (tid == 0) {
    const uint32_t wbytes = (uint32_t)P.G * 128u * 16u;
    mbar_expect_tx(bar_w, wbytes);
    bulk_g2s(smem_u32(sW), P.w + (size_t)nt * P.G * 128 * 4, wbytes, bar_w);
    if (rb < P.nblocks) load_x(b, r0, nrows);
  }
  // (rows a partial block does not load hold stale bytes: UMMA columns are independent and those columns are never stored)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ROWS >> 3) << 17) | ((128u >> 4) << 24);
  if (tid == 0) mbar_wait(bar_w, 0);      // only the issuing thread consumes the operand buffers
  uint32_t phase = 0;
  const int half = warp >> 2, wq = warp & 3;
  const int n = nt * 128 + wq * 32 + lane;                      // this lane's tap-channel
  const bool n_ok = nt * 128 + wq * 32 < P.ld;                  // warp-uniform (ld is a multiple of 32)
  const uint32_t tl = tmem + ((uint32_t)(wq * 32) << 16);
  while (rb < P.nblocks) {
    if (tid == 0) {
      mbar_wait(bar_x, phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t ad = make_desc(smem_u32(sW), 128 * 16, 128), bd = make_desc(smem_u32(sX), ROWS * 16, 128);
      for (int ks = 0; ks < P.G / 2; ++ks)
        umma_tf32(tmem, ad + (uint64_t)((ks * 2 * 128 * 16) >> 4), bd + (uint64_t)((ks * 2 * ROWS * 16) >> 4), idesc, ks ? 1u : 0u);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, phase);
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // the voxel buffer is free again: request the next block now, under this block's epilogue
    const int cb = b, cr0 = r0, cn = nrows;
    int nb = rb + gridDim.y, b2 = 0, r2 = 0, n2 = 0;
    while (nb < P.nblocks && (n2 = block_rows(nb, b2, r2)) == 0) nb += gridDim.y;
    if (tid == 0 && nb < P.nblocks) load_x(b2, r2, n2);
    rb = nb; b = b2; r0 = r2; nrows = n2;
    // ---- epilogue: lane = tap-channel n, register i of chunk c = voxel row cr0 + 16 c + i.  Warps w and w + 4 share a
    // TMEM lane quarter and take alternate 16-voxel chunks.  ~3 instructions per 128-byte store (pointer bump + STG):
    // the first version spent 12 (64-bit index arithmetic and a row predicate per store) and was issue-bound at 98 us.
    const size_t ldb = (size_t)P.ld;
    const int nchunk = (cn + 15) >> 4;
    if (n_ok) {
      for (int c = half; c < nchunk; c += 2) {
        uint32_t ra[16];
        tmem_ld16_issue(tl + (uint32_t)(c * 16), ra);
        tmem_ld16_wait(ra);
        float* yp = P.y + ((size_t)cb * P.N + cr0 + c * 16) * ldb + n;
        if (c * 16 + 16 <= cn) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { *yp = __uint_as_float(ra[i]); yp += ldb; }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) { if (c * 16 + i < cn) *yp = __uint_as_float(ra[i]); yp += ldb; }
        }
      }
    }
    // all TMEM reads of this block are complete before the next block's MMAs overwrite the accumulator
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

}  // namespace spc

bool ygemm_usable(const ConvW& y) {
  return y.tc.w && y.ntaps == 1 && y.tc.n == 128 && y.cin_pad % 8 == 0 && y.cin_pad <= 128 &&
         y.tc.nchunk * (y.tc.ck / 4) == y.cin_pad / 4;
}

// y[b][v][0..ld) = x[b][v][:] * Wy for the first nocc[b] rows of every shape
int ygemm_run(Ctx* c, const ConvW& y, const float4* xc, float* out, int ld, const int* nocc, int B, int N) {
  if (c->dry) return 0;
  if (ld % 32) { set_error("ygemm: row pitch %d is not a multiple of 32", ld); return LION_ERR_ARG; }
  spc::Params P{};
  P.x = xc; P.w = y.tc.w; P.y = out; P.nocc = nocc;
  P.G = y.tc.nchunk * (y.tc.ck / 4);        // group slots of the packing (>= cin_pad / 4; extra slots hold zero weights)
  if (P.G != y.cin_pad / 4) { set_error("ygemm: packed groups %d != input groups %d", P.G, y.cin_pad / 4); return LION_ERR_STATE; }
  P.N = N; P.B = B; P.ld = ld;
  P.blocks_per_shape = (N + spc::ROWS - 1) / spc::ROWS;
  P.nblocks = B * P.blocks_per_shape;
  const int n_tiles = y.cout_pad / 128;
  int slices = (2 * c->num_sms + n_tiles - 1) / n_tiles;
  if (slices > P.nblocks) slices = P.nblocks;
  if (slices < 1) slices = 1;
  const size_t smem = (size_t)P.G * (128 + spc::ROWS) * 16 + 64;
  static DevOnce attr_once;
  if (attr_once.need()) {
    LION_CHECK_CUDA(cudaFuncSetAttribute(spc::k_ygemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LION_CHECK_CUDA(cudaFuncSetAttribute(spc::k_ygemm, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  if (smem > 200 * 1024) { set_error("ygemm: %d input channels do not fit shared memory", y.cin_pad); return LION_ERR_ARG; }   // > 113 KB: one CTA per SM
  spc::k_ygemm<<<dim3(n_tiles, slices), 256, smem, c->stream>>>(P);
  c->launches++;
  return check_launch(c, "ygemm");
}

}  // namespace lion
