// lion_b200 -- tensor-core convolution for sm_100a: tcgen05.mma (kind::tf32) with TMEM
// accumulators, operands staged by cp.async.bulk + mbarrier pipelines, warp-specialised.
//
// One kernel serves the 3x3x3 voxel convolutions (reference: nn.Conv3d in
// models/pvcnn2_ada.py:211-222 -> cuDNN, 98 % of the step's FLOPs) and the 1x1 point
// convolutions (SharedMLP / attention projections, pvcnn2_ada.py:125-137, :51-52).
//
// Formulation (im2col-free implicit GEMM on the zero-haloed packed layouts, see
// packed_kernels.cuh):  out[p][co] = sum_tap sum_ci in[p + off(tap)][ci] * W[tap][ci][co].
//   * M = 128 consecutive rows p of one shape (voxel positions incl. y/z halo rows, masked at
//     the store), N = output channels (<= 128 per CTA), K = 8 channels per UMMA.
//   * A operand: the activation tensor is [C/4][rows][4] in HBM, i.e. for 4 channels all rows
//     are contiguous at a 16-byte pitch.  That *is* the canonical no-swizzle K-major UMMA
//     layout (core matrix = 8 rows x 16 B = 128 contiguous bytes, SBO = 128 B between 8-row
//     groups, LBO = distance between channel groups), so one contiguous cp.async.bulk per
//     channel group stages 128 + 2*halo rows, and every one of the 9 (dy,dz) taps of an x-plane
//     is the SAME shared-memory bytes viewed through a descriptor whose start address is
//     shifted by (dy*(r+2)+dz) rows.  L2->SM traffic per MAC drops 9x versus reloading per tap.
//   * B operand: weights pre-packed [n-tile][chunk][x-plane][tap][C/4][co][4] so one bulk copy
//     brings the 9 taps of a (channel chunk, x-plane); a CTA reuses it for up to 8 row tiles
//     (8 x 64 fp32 accumulator columns = all 512 TMEM columns).
//   * persistent CTAs (one per SM, 352 threads): warp 0 = bulk-copy producer, warps 1-2 = UMMA
//     issuers (even / odd row tiles), warps 3-10 = epilogue (tcgen05.ld -> +bias -> halo mask ->
//     coalesced float4 stores, GroupNorm sum / sum-of-squares reduced in registers / shared memory,
//     one fp64 atomic per channel per work item).
#include "common.cuh"
#include "model.cuh"
#include <cstdlib>

namespace lion {
namespace tc {

constexpr int EPI_WARPS = 8;
constexpr int THREADS = 96 + 32 * EPI_WARPS;   // warp 0 producer, warps 1-2 MMA issuers (even / odd row tiles), warps 3-10 epilogue
constexpr int MAX_A_STAGES = 16;   // the A ring is as deep as shared memory allows (Params::a_stages)
constexpr int B_STAGES = 2;
constexpr int MAX_ACC = 8;
constexpr int OCC_SMEM = 1024;       // bytes of per-item occupancy flags kept in shared memory (r <= 38)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (and surfaces as a CUDA error) instead of hanging the GPU
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}
// Called by all 32 lanes of a warp (every role below is warp-collective).  Lanes can leave the
// polling loop at different times, so reconverge explicitly: the elect.sync / .sync.aligned
// tcgen05 instructions that follow require the full warp.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);      // common case: already complete
  __syncwarp();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Warp-uniform issue: every lane executes these with identical operands and elect.sync picks
// the issuing lane inside the asm.  Keeping the surrounding control flow warp-uniform lets the
// compiler hold descriptors / loop counters in uniform registers (a divergent `if (lane == 0)`
// around tcgen05 instructions costs an R2UR + ELECT/BRA.U.ANY sequence per MMA).
__device__ __forceinline__ void umma_tf32_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p, q;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
               "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint32_t bar) {
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
               "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_w(uint32_t bar) {
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
               "@q mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n}" ::"r"(bar) : "memory");
}
// ---- one pipeline stage's UMMAs as ONE asm block ------------------------------------------------------------
// Round-2 finding (tools/umma_issue.cu, profiles/r02_umma_issue.txt): issuing tcgen05.mma one asm statement at a
// time costs ~64-140 clk per UMMA per warp -- every MMA drags an ELECT, five R2UR.BROADCAST (descriptor halves and
// the TMEM address have to reach uniform registers) and two VOTEU through the issue slot, serialised by the
// volatile statement boundaries -- which is MORE than the 32 clk an M128 x N64 x K8 UMMA occupies the tensor pipe:
// the round-1 kernel was instruction-issue bound, not operand- or tensor-bound.  Emitting a whole stage (TPG taps
// x KS k-steps) as one block with a single elect.sync and descriptors derived by in-asm adds lets ptxas software-
// pipeline the R2URs of later MMAs under earlier UTCHMMAs (~6 instructions per MMA).
//   operands: %0 TMEM accumulator, %1 A descriptor low word (LBO<<16 | addr16 of the un-shifted view), %2 B ditto,
//   %3 idesc, %4 accumulate flag of the FIRST MMA, %5 descriptor high word, %6 A k-step (2 groups), %7 B k-step,
//   %8 B tap pitch, %9.. row offset of tap t.  Low-word adds never carry out of the 14-bit address field
//   (all views lie inside this CTA's shared memory).
#define LION_MMA1(AL, BL, PRED) "mov.b64 ad, {" AL ", %5};\n mov.b64 bd, {" BL ", %5};\n @q tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %3, " PRED ";\n"
#define LION_TAP4(TOP, FP)                                                                 \
  "add.u32 a0, %1, " TOP ";\n" LION_MMA1("a0", "bt", FP)                                   \
  "add.u32 a1, a0, %6;\n add.u32 b1, bt, %7;\n" LION_MMA1("a1", "b1", "pt")                \
  "add.u32 a2, a1, %6;\n add.u32 b2, b1, %7;\n" LION_MMA1("a2", "b2", "pt")                \
  "add.u32 a3, a2, %6;\n add.u32 b3, b2, %7;\n" LION_MMA1("a3", "b3", "pt")                \
  "add.u32 bt, bt, %8;\n"
#define LION_TAP2(TOP, FP)                                                                 \
  "add.u32 a0, %1, " TOP ";\n" LION_MMA1("a0", "bt", FP)                                   \
  "add.u32 a1, a0, %6;\n add.u32 b1, bt, %7;\n" LION_MMA1("a1", "b1", "pt")                \
  "add.u32 bt, bt, %8;\n"
#define LION_TAP1(TOP, FP)                                                                 \
  "add.u32 a0, %1, " TOP ";\n" LION_MMA1("a0", "bt", FP)                                   \
  "add.u32 bt, bt, %8;\n"
#define LION_STAGE_HEAD                                                                    \
  "{\n.reg .pred q, p, pt;\n.reg .b32 a0, a1, a2, a3, b1, b2, b3, bt;\n.reg .b64 ad, bd;\n" \
  "elect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\nsetp.eq.b32 pt, 0, 0;\nmov.b32 bt, %2;\n"
#define LION_STAGE9(TAP)                                                                   \
  asm volatile(LION_STAGE_HEAD TAP("%9", "p") TAP("%10", "pt") TAP("%11", "pt") TAP("%12", "pt") TAP("%13", "pt")     \
               TAP("%14", "pt") TAP("%15", "pt") TAP("%16", "pt") TAP("%17", "pt") "}\n"                                 \
               ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(hi), "r"(a_k2), "r"(b_k2), "r"(b_tap),          \
                 "r"(to[0]), "r"(to[1]), "r"(to[2]), "r"(to[3]), "r"(to[4]), "r"(to[5]), "r"(to[6]), "r"(to[7]), "r"(to[8]) : "memory")
#define LION_STAGE1(TAP)                                                                   \
  asm volatile(LION_STAGE_HEAD TAP("%9", "p") "}\n"                                                                     \
               ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(hi), "r"(a_k2), "r"(b_k2), "r"(b_tap), "r"(to[0]) : "memory")

// all UMMAs of one (row tile, channel chunk, tap group): TPG taps x KG/2 k-steps.  Warp-collective (elect inside).
template <int KG, int TPG>
__device__ __forceinline__ void issue_stage(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc, uint32_t hi,
                                            uint32_t a_k2, uint32_t b_k2, uint32_t b_tap, const int* to) {
  static_assert((KG == 2 || KG == 4 || KG == 8) && (TPG == 1 || TPG == 9), "unsupported stage shape");
  if constexpr (TPG == 9) {
    if constexpr (KG == 8) LION_STAGE9(LION_TAP4);
    else if constexpr (KG == 4) LION_STAGE9(LION_TAP2);
    else LION_STAGE9(LION_TAP1);
  } else {
    if constexpr (KG == 8) LION_STAGE1(LION_TAP4);
    else if constexpr (KG == 4) LION_STAGE1(LION_TAP2);
    else LION_STAGE1(LION_TAP1);
  }
}

// no-swizzle K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// split form: issue the load, then make the registers depend on the wait (the "+r" operands keep
// the compiler from scheduling consumers above tcgen05.wait::ld)
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
  const float4* in; const float* w; const float* bias; float4* out; double* ssum; double* ssq;
  int Gin, Gout_store, cout_pad;
  int rows;              // rows per (b, group)
  int p_begin, p_end;
  int rp;                // r+2 (halo mask) or 0
  int ntg, tpg;          // tap groups (x planes) and taps per group: (3,9) or (1,1)
  int tg_off[3];         // row offset of each tap group (dx * rp^2)
  int tap_off[9];        // row offset of each tap inside a group (dy*rp + dz)
  int halo;              // rp+1 or 0
  int KG, nchunk;        // channel groups per chunk, chunks
  int NT;                // output channels per CTA (UMMA N)
  int G;                 // row tiles (accumulators) per work item
  int B;                 // shapes
  int a_stage_bytes, b_stage_bytes, stage_rows;
  int a_stages;          // depth of the A ring
  const unsigned char* occ;   // 64-row occupancy flags of the input (sparse first conv of a PVConv) or null
  int occ_stride;
  // pooled 1x1 (last layer of a set-abstraction MLP): instead of the [rows][C] result, write per 32 consecutive rows (the
  // neighbours of one centre = the 32 TMEM lanes of one epilogue warp) the per-channel minimum and maximum,
  // pool_mm[b][C/4][rows/32][2][4].  AdaGN's affine is monotonic and Swish is quasi-convex (one minimum), so
  // max_i swish(s*x_i + t) = max(swish(s*min + t), swish(s*max + t)): the consumer needs 2 of the 32 values.
  float* pool_mm;
  // row-major output (the y = x W GEMM of the sparse first convolution): out_rm[(b*rows + p) * ld_rm + n], no PF store
  float* out_rm; int ld_rm;
  int sched;             // work distribution: 0 contiguous range per CTA, 1 interleaved items (see k_conv_tc)
};

// per 32 channels: butterfly that leaves in lane l the sum over the warp's 32 rows of channel l.
// in: v[32] (this lane's row, 32 channels); cost 31 shuffles instead of 160.
__device__ __forceinline__ float warp_transpose_sum(float* v, int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float keep = upper ? v[i + half] : v[i];
      float send = upper ? v[i] : v[i + half];
      float got = __shfl_xor_sync(0xffffffffu, send, half);
      v[i] = keep + got;
    }
  }
  return v[0];
}

// per 16 channels: butterfly that leaves in lane l the sum over the warp's 32 rows of channel
// ch16(l) = 8*bit4 + 4*bit3 + 2*bit2 + bit1 of l (lanes 2k and 2k+1 hold the same channel): 16 shuffles.
__device__ __forceinline__ float warp_transpose_sum16(float* v, int lane) {
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const int bit = half * 2;               // lane bit 16, 8, 4, 2
    bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float keep = upper ? v[i + half] : v[i];
      float send = upper ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}
// same butterfly with min / max: v is destroyed; returns in lane l the extreme over the warp's 32 rows of channel ch16(l)
template <bool MAX>
__device__ __forceinline__ float warp_transpose_ext16(float* v, int lane) {
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const int bit = half * 2;
    bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float keep = upper ? v[i + half] : v[i];
      float send = upper ? v[i] : v[i + half];
      float got = __shfl_xor_sync(0xffffffffu, send, bit);
      v[i] = MAX ? fmaxf(keep, got) : fminf(keep, got);
    }
  }
  float o = __shfl_xor_sync(0xffffffffu, v[0], 1);
  return MAX ? fmaxf(v[0], o) : fminf(v[0], o);
}
__device__ __forceinline__ int ch16_of_lane(int lane) { return (lane >> 1) & 15; }
// pooled epilogue: v[16] = this lane's row, columns col..col+15 of the item's n-tile (bias added); centre = row / 32
__device__ __forceinline__ void pool_store16(const Params& P, int b, int n0, int col, long long centre, const float* v, int lane) {
  float lo[16], hi[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { lo[i] = v[i]; hi[i] = v[i]; }
  const float mn = warp_transpose_ext16<false>(lo, lane);
  const float mx = warp_transpose_ext16<true>(hi, lane);
  const int c = n0 + col + ch16_of_lane(lane);
  const long long ncent = (long long)P.rows / 32;
  float* dst = P.pool_mm + ((((size_t)b * (P.cout_pad / 4) + (c >> 2)) * ncent + centre) * 2) * 4 + (c & 3);
  if ((lane & 1) == 0) { dst[0] = mn; dst[4] = mx; }
}

template <int KG, int TPG>
__global__ void __launch_bounds__(THREADS, 1) k_conv_tc(Params P) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: see LION_LAUNCH
  extern __shared__ __align__(128) uint8_t smem[];
  // layout: [A stages][B stages][bias 128 floats][stat 8*2*64 floats][barriers][tmem ptr, skip flags]
  uint8_t* sA = smem;
  const int A_STAGES = P.a_stages;
  uint8_t* sB = sA + (size_t)A_STAGES * P.a_stage_bytes;
  float* s_bias = (float*)(sB + (size_t)B_STAGES * P.b_stage_bytes);
  float* s_stat = s_bias + 128;                 // [8 epilogue warps][2][64]
  uint64_t* bars = (uint64_t*)(s_stat + 8 * 2 * 64);
  uint32_t* s_tmem = (uint32_t*)(bars + 64);
  volatile uint32_t* s_skip = s_tmem + 1;        // [MAX_A_STAGES] stage holds no data (all-zero input slab)
  uint8_t* s_occ = (uint8_t*)(s_tmem + 32);      // [OCC_SMEM] this item's 64-row occupancy flags (sparse first convolution)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const uint32_t bar_full_a = smem_u32(bars), bar_empty_a = smem_u32(bars + MAX_A_STAGES);
  const uint32_t bar_full_b = smem_u32(bars + 2 * MAX_A_STAGES), bar_empty_b = smem_u32(bars + 2 * MAX_A_STAGES + B_STAGES);
  const uint32_t bar_accf = smem_u32(bars + 2 * MAX_A_STAGES + 2 * B_STAGES);   // MAX_ACC: accumulator j complete
  const uint32_t bar_tfree = bar_accf + 8 * MAX_ACC;                            // MAX_ACC: accumulator j drained

  // zero the A stages once: channel-group slots that a partial chunk does not load must hold
  // finite values (their weights are zero)
  if (P.Gin % KG != 0)
    for (int i = tid; i < A_STAGES * P.a_stage_bytes / 16; i += THREADS) ((float4*)sA)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(bar_full_a + 8 * i, 1); mbar_init(bar_empty_a + 8 * i, 2); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(bar_full_b + 8 * i, 1); mbar_init(bar_empty_b + 8 * i, 2); }
    for (int i = 0; i < MAX_ACC; ++i) { mbar_init(bar_accf + 8 * i, 1); mbar_init(bar_tfree + 8 * i, EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zero fill -> async proxy readers
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *s_tmem;
  // everything above touched only shared / tensor memory and overlapped the previous kernel's
  // tail; from here on global memory written by that kernel is consumed
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // Work distribution.  The (n-tile, shape, row-tile) space is flattened (row tile fastest) and cut into work items = up to G
  // consecutive row tiles sharing the weight slabs (they may belong to two shapes, never to two n-tiles).  Every CTA gets
  // the same number of tiles (+-1).  Round 1 dealt fixed G-tile items round-robin: with 41 tiles per shape (r = 16) that
  // left some CTAs 12 tiles and others 8 -- the kernel ran at the pace of the 12 (profiles/r02_conv_balanced_ranges.txt).
  //   sched 0: one contiguous range per CTA, cut into evenly sized items (9 tiles -> 3+3+3, not 4+4+1).  Balanced, but at
  //            r = 32 the 147 CTAs then stream 147 distant windows of a 268 MB input at once: the 3 x-plane sweeps of a tile
  //            (9 tiles apart) miss the L2 and the launch reads 975 MB from DRAM instead of 268 (ncu, profiles/).
  //   sched 1: the same tiles per CTA T_i (= the contiguous range's size) and the same number of items m = ceil(max T / G), but
  //            dealt in m ROUNDS: round k is one contiguous stretch of the tile space, cut into one item per CTA (CTA i's item
  //            has floor(T_i (k+1) / m) - floor(T_i k / m) tiles).  At any moment the CTAs work on ~grid ADJACENT items
  //            (a few shapes), so a tile's x-plane neighbours are in flight in a neighbouring CTA and hit the L2.
  //            With T_i in {q, q+1} every offset has a closed form: A_k = floor(q k / m), B_k = floor((q+1) k / m),
  //            c_i = CTAs before i that own q+1 tiles;  start(i, k) = n_q A_k + n_p B_k + (i - c_i)(A_k+1 - A_k) + c_i (B_k+1 - B_k).
  const int ntile_total = (P.p_end - P.p_begin + 127) / 128;
  const int n_nt = P.cout_pad / P.NT;
  const long long U = (long long)n_nt * P.B * ntile_total;
  const long long u_begin = U * blockIdx.x / gridDim.x, u_end = U * (blockIdx.x + 1) / gridDim.x;
  struct Items {
    long long u, u_end; int ntile_total, B, G;
    int mode, m, q, n_p, n_q, c_i, i, big, k;   // sched 1: [u, u_end) is the rest of the current item (cut at n-tile boundaries)
    // next item: n-tile nt, first tile v0 in the n-tile's flat (shape, row tile) space, ntile tiles.  An item may run
    // across a shape boundary -- its tiles share the weight slabs whatever shape they belong to -- but not across n-tiles.
    __device__ __forceinline__ bool next(int& nt, long long& v0, int& ntile) {
      const long long per_nt = (long long)B * ntile_total;
      if (mode) {
        while (u >= u_end) {
          if (k >= m) return false;
          const int a0 = q * k / m, a1 = q * (k + 1) / m, b0 = (q + 1) * k / m, b1 = (q + 1) * (k + 1) / m;
          u = (long long)n_q * a0 + (long long)n_p * b0 + (long long)(i - c_i) * (a1 - a0) + (long long)c_i * (b1 - b0);
          u_end = u + (big ? b1 - b0 : a1 - a0);
          ++k;
        }
        nt = (int)(u / per_nt);
        v0 = u - (long long)nt * per_nt;
        const long long lim = (long long)(nt + 1) * per_nt;
        const long long e = u_end < lim ? u_end : lim;
        ntile = (int)(e - u);          // <= ceil((q + 1) / m) <= G
        u = e;
        return true;
      }
      if (u >= u_end) return false;
      nt = (int)(u / per_nt);
      v0 = u - (long long)nt * per_nt;
      long long run = per_nt - v0;
      if (u_end - u < run) run = u_end - u;
      const long long k2 = (run + G - 1) / G;
      ntile = (int)((run + k2 - 1) / k2);
      u += ntile;
      return true;
    }
  };
  Items items0{u_begin, u_end, ntile_total, P.B, P.G, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (P.sched) {
    const int grid = (int)gridDim.x, q = (int)(U / grid), n_p = (int)(U - (long long)q * grid);
    const int tmax = q + (n_p ? 1 : 0);
    items0.mode = 1; items0.u = 0; items0.u_end = 0;
    items0.m = (tmax + P.G - 1) / P.G; items0.q = q; items0.n_p = n_p; items0.n_q = grid - n_p;
    items0.i = (int)blockIdx.x; items0.c_i = (int)(u_begin - (long long)q * blockIdx.x);
    items0.big = (int)(u_end - u_begin) > q; items0.k = 0;
  }

  if (warp == 0) {
    // ===================== producer (whole warp; lane kg issues the copy of channel group kg) ====
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0;          // ring positions and phase bits
    const uint32_t bytes = (uint32_t)P.stage_rows * 16u;
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    Items items = items0;
    int nt, ntile;
    long long v0;
    int occ_shape = -1;                                // shape whose occupancy flags s_occ holds
    while (items.next(nt, v0, ntile)) {
      const float* wsrc = P.w + (size_t)nt * P.nchunk * P.ntg * (P.b_stage_bytes / 4);
      // lane l keeps (shape, row tile) of the item's tile l: one division per item, a shuffle per stage
      int my_b = 0, my_t = 0;
      if (lane < ntile) { const int vl = (int)v0 + lane; my_b = vl / ntile_total; my_t = vl - my_b * ntile_total; }
      for (int cc = 0; cc < P.nchunk; ++cc) {
        const int kg_real = min(KG, P.Gin - cc * KG);
        const size_t grp_lane = (size_t)(cc * KG + (lane < kg_real ? lane : 0));
        for (int tg = 0; tg < P.ntg; ++tg) {
          // the last sweep of an item is never skipped: every accumulator then receives at
          // least one (zero-initialising) MMA per item and needs no "was it touched" bookkeeping
          const bool may_skip = P.occ && !(cc == P.nchunk - 1 && tg == P.ntg - 1);
          mbar_wait(bar_empty_b + 8 * sb, pb ^ 1);
          if (lane == 0) {
            mbar_expect_tx(bar_full_b + 8 * sb, P.b_stage_bytes);
            bulk_g2s(sB_addr + sb * (uint32_t)P.b_stage_bytes, wsrc + (size_t)(cc * P.ntg + tg) * (P.b_stage_bytes / 4),
                     P.b_stage_bytes, bar_full_b + 8 * sb);
          }
          if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
          for (int j = 0; j < ntile; ++j) {
            const int bj = __shfl_sync(0xffffffffu, my_b, j), tj = __shfl_sync(0xffffffffu, my_t, j);
            const long long row0 = (long long)P.p_begin + (long long)tj * 128 - P.halo + P.tg_off[tg];
            const float4* in_lane = P.in + ((size_t)bj * P.Gin + grp_lane) * P.rows;
            mbar_wait(bar_empty_a + 8 * sa, pa ^ 1);
            bool empty = false;
            if (may_skip) {
              // sparse input: the shape's occupancy flags (<= 1 KB) live in shared memory -- a broadcast LDS per
              // check instead of dependent global loads on the producer's critical path
              const unsigned char* occ_b = P.occ + (size_t)bj * P.occ_stride;
              if (P.occ_stride <= OCC_SMEM) {
                if (occ_shape != bj) {
                  __syncwarp();
                  for (int k = lane; k < P.occ_stride; k += 32) s_occ[k] = __ldg(occ_b + k);
                  __syncwarp();
                  occ_shape = bj;
                }
                occ_b = s_occ;
              }
              long long lo = row0 < 0 ? 0 : row0, hi = row0 + P.stage_rows - 1;
              if (hi > P.rows - 1) hi = P.rows - 1;
              unsigned any = 0;
              for (int k = (int)(lo >> 6); k <= (int)(hi >> 6); ++k) any |= occ_b[k];
              empty = (any == 0);
            }
            const uint32_t full = bar_full_a + 8 * sa;
            if (lane == 0) {
              s_skip[sa] = empty ? 1u : 0u;
              if (empty) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(full) : "memory");
              else mbar_expect_tx(full, bytes * kg_real);
            }
            __syncwarp();
            if (!empty && lane < kg_real)
              bulk_g2s(sA_addr + sa * (uint32_t)P.a_stage_bytes + lane * bytes, in_lane + row0, bytes, full);
            if (++sa == (uint32_t)A_STAGES) { sa = 0; pa ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: warp 1 owns even row tiles, warp 2 odd ones ==========
    // (each whole warp runs the loop; one elected lane issues).  Two issuers because a single
    // thread cannot generate descriptors + issue one UMMA every 16-32 cycles (N <= 64).
    const int me = warp - 1;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.NT >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_pitch16 = (uint32_t)P.stage_rows;             // (bytes between channel groups) >> 4
    const uint32_t b_pitch16 = (uint32_t)P.NT;
    const uint32_t b_tap16 = (uint32_t)KG * b_pitch16;
    // descriptor high words are constant: SBO = 128 B, version 1, no swizzle; LBO = group pitch
    const uint32_t d_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo_c = (a_pitch16 & 0x3fff) << 16, b_lo_c = (b_pitch16 & 0x3fff) << 16;
    const uint32_t a_stage16 = (uint32_t)P.a_stage_bytes >> 4;
    const uint32_t a_ring16 = (smem_u32(sA) >> 4) + (uint32_t)P.halo;
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
    Items items = items0;
    int nt, ntile;
    long long v0;
    for (; items.next(nt, v0, ntile); ++it) {
      uint32_t started = 0;
      // accumulators this item does not use still take part in the per-item phase bookkeeping:
      // wait until the epilogue released them (previous item) before re-arming them below
      for (int j = me; j < MAX_ACC; j += 2)
        if (j >= ntile) mbar_wait(bar_tfree + 8 * j, (it & 1) ^ 1);
      for (int cc = 0; cc < P.nchunk; ++cc) {
        for (int tg = 0; tg < P.ntg; ++tg) {
          mbar_wait(bar_full_b + 8 * sb, pb);
          const uint32_t b_base16 = smem_u32(sB + (size_t)sb * P.b_stage_bytes) >> 4;
          const bool first = (cc | tg) == 0;
          const bool last = (cc == P.nchunk - 1) && (tg == P.ntg - 1);
          for (int jt = 0; jt < ntile; ++jt) {
            const uint32_t my_sa = sa, my_pa = pa;
            if (++sa == (uint32_t)A_STAGES) { sa = 0; pa ^= 1; }
            // BOTH issuers wait on every stage and both release it (empty count 2).  A parity wait is
            // only sound for a waiter that observes every phase of a barrier: bulk copies complete out
            // of order, so an issuer that skipped the other's stages could see a slot's barrier one
            // phase behind and mistake "not yet loaded" for "loaded" (seen with odd ring depths).
            mbar_wait(bar_full_a + 8 * my_sa, my_pa);
            bool issued = false;
            if ((jt & 1) == me) {                                         // else: the other issuer's tile
              if (first) mbar_wait(bar_tfree + 8 * jt, (it & 1) ^ 1);    // accumulator jt drained (previous item)
              if (!s_skip[my_sa]) {                                      // (all-zero input slab: nothing to accumulate)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_base16 = a_ring16 + my_sa * a_stage16;
                const uint32_t d = tmem_base + (uint32_t)(jt * P.NT);
                const bool fresh = ((started >> jt) & 1u) == 0;
                started |= 1u << jt;
                issue_stage<KG, TPG>(d, a_lo_c | (a_base16 & 0x3fff), b_lo_c | (b_base16 & 0x3fff), idesc, fresh ? 0u : 1u, d_hi,
                                     2u * a_pitch16, 2u * b_pitch16, b_tap16, P.tap_off);
                if (last) umma_commit_w(bar_accf + 8 * jt);   // accumulator jt complete -> epilogue may drain it
                issued = true;
              }
            }
            // hand the stage back: when this warp's MMAs retire, or at once if it issued none
            if (issued) umma_commit_w(bar_empty_a + 8 * my_sa);
            else mbar_arrive_w(bar_empty_a + 8 * my_sa);
          }
          umma_commit_w(bar_empty_b + 8 * sb);            // (count 2: both issuers)
          if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
        }
      }
      for (int j = me; j < MAX_ACC; j += 2)
        if (j >= ntile) umma_commit_w(bar_accf + 8 * j);  // exactly one arrival per accumulator per item
    }
  } else {
    // ===================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves =========
    const int ew = warp - 3, et = tid - 96;
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int hcol = (ew >> 2) * (P.NT / 2);       // first accumulator column of this warp's half
    const int CH = P.NT / 2;
    uint32_t it = 0;
    Items items = items0;
    int nt, ntile;
    long long v0;
    for (; items.next(nt, v0, ntile); ++it) {
      const int n0 = nt * P.NT;
      asm volatile("bar.sync 1, 256;" ::: "memory");            // previous item's s_bias / s_stat readers are done
      if (et < P.NT) s_bias[et] = P.bias ? P.bias[n0 + et] : 0.0f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float run_s[4] = {0, 0, 0, 0}, run_q[4] = {0, 0, 0, 0};    // per 16-column chunk of this warp's half
      // GroupNorm statistics are per shape: flushed whenever the item moves on to the next shape, and at its end
      // (warp partials -> shared memory -> one fp64 atomic per channel; block-uniform control flow: bar.sync inside)
      auto flush_stats = [&](int bb) {
        if ((lane & 1) == 0) {
          for (int k = 0; k < CH / 16; ++k) {
            s_stat[(ew * 2 + 0) * 64 + k * 16 + ch16_of_lane(lane)] = run_s[k];
            s_stat[(ew * 2 + 1) * 64 + k * 16 + ch16_of_lane(lane)] = run_q[k];
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et < P.NT) {
          const int half = et / CH, ch = et % CH;
          float s = 0.f, qq = 0.f;
#pragma unroll
          for (int ww = 0; ww < 4; ++ww) { s += s_stat[((half * 4 + ww) * 2 + 0) * 64 + ch]; qq += s_stat[((half * 4 + ww) * 2 + 1) * 64 + ch]; }
          atomicAdd(P.ssum + (size_t)bb * P.cout_pad + n0 + et, (double)s);
          atomicAdd(P.ssq + (size_t)bb * P.cout_pad + n0 + et, (double)qq);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");          // s_stat is free again
#pragma unroll
        for (int k = 0; k < 4; ++k) { run_s[k] = 0.0f; run_q[k] = 0.0f; }
      };
      int b = (int)(v0 / ntile_total);
      if (TPG == 1 && CH <= 32) {
        // 1x1 convolutions, N <= 64 (compiled out of the 3x3x3 instantiations, whose epilogue is off the
        // critical path and whose issue loop suffered from the extra register pressure): per-lane (= per-row) running sums over the item's tiles, ONE cross-lane reduction
        // per item instead of one per tile; both 16-column loads of a tile are in flight together.
        // (1x1 convolutions have almost no MMA work per tile, so this loop is their critical path.)
        float acc_s[2][16], acc_q[2][16];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
          for (int i = 0; i < 16; ++i) { acc_s[cc][i] = 0.0f; acc_q[cc][i] = 0.0f; }
        auto fold_acc = [&]() {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            if (cc * 16 < CH) {
              run_s[cc] = warp_transpose_sum16(acc_s[cc], lane);
              run_q[cc] = warp_transpose_sum16(acc_q[cc], lane);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) { acc_s[cc][i] = 0.0f; acc_q[cc][i] = 0.0f; }
          }
        };
        for (int j = 0; j < ntile; ++j) {
          const long long vj = v0 + j;
          const int bj = (int)(vj / ntile_total), tj = (int)(vj - (long long)bj * ntile_total);
          if (bj != b) { if (P.ssum) { fold_acc(); flush_stats(b); } b = bj; }
          int p = P.p_begin + tj * 128 + q * 32 + lane;
          bool inrange = p < P.p_end;
          bool valid = inrange;
          if (P.rp > 0 && inrange) {
            int z = p % P.rp, y = (p / P.rp) % P.rp;
            valid = (z >= 1 && z <= P.rp - 2 && y >= 1 && y <= P.rp - 2);
          }
          mbar_wait(bar_accf + 8 * j, it & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          uint32_t rr[2][16];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * P.NT + hcol);
          tmem_ld16_issue(taddr, rr[0]);
          if (CH > 16) tmem_ld16_issue(taddr + 16, rr[1]);
          tmem_ld16_wait(rr[0]);
          if (CH > 16) tmem_ld16_wait(rr[1]);
          // accumulator j is in registers: let the issuers start the next item's tile j
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tfree + 8 * j) : "memory");
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            if (cc * 16 < CH) {
              const int col = hcol + cc * 16;
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = valid ? __uint_as_float(rr[cc][i]) + s_bias[col + i] : 0.0f;
              if (P.pool_mm) pool_store16(P, b, n0, col, (long long)(p >> 5), v, lane);     // all rows valid (rows % 128 == 0)
              else if (inrange) {
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                  int g = (n0 + col) / 4 + g4;
                  if (g < P.Gout_store)
                    P.out[((size_t)b * P.Gout_store + g) * P.rows + p] = make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
                }
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) { acc_s[cc][i] += v[i]; acc_q[cc][i] = fmaf(v[i], v[i], acc_q[cc][i]); }
            }
          }
        }
        if (P.ssum) fold_acc();
      } else
      for (int j = 0; j < ntile; ++j) {
        const long long vj = v0 + j;
        const int bj = (int)(vj / ntile_total), tj = (int)(vj - (long long)bj * ntile_total);
        if (bj != b) { if (P.ssum) flush_stats(b); b = bj; }
        int p = P.p_begin + tj * 128 + q * 32 + lane;
        bool inrange = p < P.p_end;
        bool valid = inrange;
        if (P.rp > 0 && inrange) {
          int z = p % P.rp, y = (p / P.rp) % P.rp;
          valid = (z >= 1 && z <= P.rp - 2 && y >= 1 && y <= P.rp - 2);
        }
        mbar_wait(bar_accf + 8 * j, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c16 = 0; c16 < CH; c16 += 16) {
          float v[16];
          const int col = hcol + c16;
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * P.NT + col), v);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = valid ? v[i] + s_bias[col + i] : 0.0f;
          if (TPG == 1 && P.pool_mm) pool_store16(P, b, n0, col, (long long)(p >> 5), v, lane);
          else if (TPG == 1 && P.out_rm) {
            if (inrange && n0 + col < P.ld_rm) {
              float4* d = reinterpret_cast<float4*>(P.out_rm + ((size_t)b * P.rows + p) * P.ld_rm + n0 + col);
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) d[g4] = make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
            }
          } else if (inrange) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              int g = (n0 + col) / 4 + g4;
              if (g < P.Gout_store)
                P.out[((size_t)b * P.Gout_store + g) * P.rows + p] = make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
            }
          }
          if (P.ssum) {
            float sq[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) sq[i] = v[i] * v[i];
            run_s[c16 >> 4] += warp_transpose_sum16(v, lane);
            run_q[c16 >> 4] += warp_transpose_sum16(sq, lane);
          }
        }
        // accumulator j is drained (this warp's part): let the issuers start the next item's tile j
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tfree + 8 * j) : "memory");
      }
      if (P.ssum) flush_stats(b);
      // accumulators this item did not use: release them too so that the phase bookkeeping of
      // bar_tfree stays in step with the item counter
      for (int j = ntile; j < MAX_ACC; ++j) {
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tfree + 8 * j) : "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// weights: from the SIMT packing wt[tap][cin_pad][cout_pad] to
//   w[nt][chunk][tg][t][kg][n][4]  (tf32-rounded, round-to-nearest-away like cuDNN's conversion)
__global__ void k_pack_tc(const float* __restrict__ wt, float* __restrict__ w, int ntaps, int cin_pad, int cout_pad,
                          int NT, int nchunk, int ntg, int tpg, int KG) {
  pdl_prologue();
  size_t total = (size_t)(cout_pad / NT) * nchunk * ntg * tpg * KG * NT * 4;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int jj = i % 4;
  size_t r = i / 4;
  int n = r % NT; r /= NT;
  int kg = r % KG; r /= KG;
  int t = r % tpg; r /= tpg;
  int tg = r % ntg; r /= ntg;
  int cc = r % nchunk; r /= nchunk;
  int nt = (int)r;
  int tap = tg * tpg + t;
  int ci = (cc * KG + kg) * 4 + jj;
  float v = 0.0f;
  if (ci < cin_pad) v = wt[((size_t)tap * cin_pad + ci) * cout_pad + nt * NT + n];
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  w[i] = __uint_as_float(u);
}

}  // namespace tc

static int tc_mode() {      // 1 = tensor cores (default), 0 = SIMT only (LION_CONV_IMPL=simt; bring-up/debug)
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("LION_CONV_IMPL");
    mode = (e && strcmp(e, "simt") == 0) ? 0 : 1;
  }
  return mode;
}

static void tc_shape(const ConvW& w, int& NT, int& KG, int& nchunk, int& ntg, int& tpg) {
  NT = w.cout_pad < 128 ? w.cout_pad : 128;
  ntg = w.ntaps == 27 ? 3 : 1;
  tpg = w.ntaps == 27 ? 9 : 1;
  int G = w.cin_pad / 4;
  // 3x3x3: 32-channel chunks (36 UMMAs per activation stage) for N <= 64 -- the per-stage barrier
  // round trip is amortised over more MMAs, ~15 % faster than 16-channel chunks -- and 16-channel
  // chunks for N = 128, where the 9-tap weight slab (2 x 72 KB) leaves room for a 16-channel ring only.
  // 1x1: 32-channel chunks.
  if (w.ntaps == 27) KG = (NT > 64) ? 4 : 8;
  else KG = 8;
  { static int kg64 = -1; if (kg64 < 0) { const char* e = getenv("LION_TC_KG64"); kg64 = e ? atoi(e) : 0; }
    if (kg64 && w.ntaps == 27 && NT == 64) KG = kg64; }
  if (G < KG) KG = (G <= 2) ? 2 : ((G <= 4) ? 4 : 8);
  nchunk = (G + KG - 1) / KG;
}

int conv_tc_prepare(Model* m, ConvW& w) {
  w.tc = ConvTcW();
  if (!(w.ntaps == 27 || w.ntaps == 1)) return 0;
  if (w.cout_pad < 32 || w.cout_pad % 32) return 0;      // epilogue splits N into two halves of 16-column chunks
  int NT, KG, nchunk, ntg, tpg;
  tc_shape(w, NT, KG, nchunk, ntg, tpg);
  if (w.cout_pad % NT) return 0;
  size_t total = (size_t)(w.cout_pad / NT) * nchunk * ntg * tpg * KG * NT * 4;
  LION_TRY(m->dmalloc(&w.tc.w, total));
  w.tc.ck = KG * 4; w.tc.nchunk = nchunk; w.tc.n = NT;
  PackJob j{2, w.wt, nullptr, w.tc.w, w.ntaps, w.cin_pad, w.cout_pad, NT, 0};
  j.kmap = nullptr;
  m->jobs.push_back(j);
  return 0;
}

int conv_tc_pack_job(const PackJob& j) {
  ConvW tmp;
  tmp.ntaps = j.a; tmp.cin_pad = j.b; tmp.cout_pad = j.c;
  int NT, KG, nchunk, ntg, tpg;
  tc_shape(tmp, NT, KG, nchunk, ntg, tpg);
  size_t total = (size_t)(j.c / NT) * nchunk * ntg * tpg * KG * NT * 4;
  tc::k_pack_tc<<<(unsigned)cdivz(total, 256), 256>>>(j.src, j.dst, j.a, j.b, j.c, NT, nchunk, ntg, tpg, KG);
  return 0;
}

bool conv_tc_usable(const ConvW& w, const ConvGeom& geo) {
  if (!tc_mode() || !w.tc.w) return false;
  if (geo.ntaps != w.ntaps) return false;
  return true;
}

int conv_tc_run(Ctx* c, const ConvW& w, const float4* in, int Gin, float4* out, int Gout_store, double* ssum, double* ssq,
                const ConvGeom& geo, int B, float* pool_mm, float* out_rm, int ld_rm) {
  tc::Params P{};
  if (out_rm && (w.ntaps != 1 || w.cout_pad < 128 || ld_rm % 16)) {
    set_error("conv_tc: the row-major epilogue needs a 1x1 convolution with >= 128 output channels"); return LION_ERR_ARG;
  }
  P.out_rm = out_rm; P.ld_rm = ld_rm;
  if (pool_mm && (w.ntaps != 1 || geo.p_begin != 0 || geo.p_end != geo.rows || geo.rows % 128)) {
    set_error("conv_tc: the pooled epilogue needs a 1x1 convolution over a multiple of 128 rows"); return LION_ERR_ARG;
  }
  P.pool_mm = pool_mm;
  int NT, KG, nchunk, ntg, tpg;
  tc_shape(w, NT, KG, nchunk, ntg, tpg);
  P.in = in; P.w = w.tc.w; P.bias = w.bias; P.out = out; P.ssum = ssum; P.ssq = ssq;
  P.Gin = Gin; P.Gout_store = Gout_store; P.cout_pad = w.cout_pad;
  P.rows = geo.rows; P.p_begin = geo.p_begin; P.p_end = geo.p_end; P.rp = geo.rp;
  P.ntg = ntg; P.tpg = tpg; P.KG = KG; P.nchunk = nchunk; P.NT = NT;
  if (w.ntaps == 27) {
    int rp = geo.rp;
    for (int dx = 0; dx < 3; ++dx) P.tg_off[dx] = (dx - 1) * rp * rp;
    for (int dy = 0; dy < 3; ++dy) for (int dz = 0; dz < 3; ++dz) P.tap_off[dy * 3 + dz] = (dy - 1) * rp + (dz - 1);
    P.halo = rp + 1;
  } else {
    P.tg_off[0] = 0; P.tap_off[0] = 0; P.halo = 0;
  }
  P.stage_rows = 128 + 2 * P.halo;
  P.a_stage_bytes = KG * P.stage_rows * 16;
  P.b_stage_bytes = tpg * KG * NT * 16;
  int ntile = cdiv(geo.p_end - geo.p_begin, 128);
  int n_tiles_n = w.cout_pad / NT;
  int gmax = 512 / NT;
  if (gmax > tc::MAX_ACC) gmax = tc::MAX_ACC;
  // row tiles per work item: as many as TMEM holds (the weight slab of a stage is then shared by that many MMA groups);
  // parallelism does not depend on G any more -- the kernel cuts the flat tile space into equal ranges per CTA
  int G = gmax;
  { static int ge = -1; if (ge < 0) { const char* e = getenv("LION_TC_G"); ge = e ? atoi(e) : 0; } if (ge > 0 && ge < G) G = ge; }
  P.G = G;
  P.B = B;
  // Round-based items (sched 1) for the 64-channel grids at r = 32 (268 MB of input at B = 32, twice the L2): DRAM reads of
  // the 64 -> 64 launch drop from 975 MB to 326 MB and it runs 2-3 % faster (0.400 vs 0.4125 ms).  Launches whose input is
  // L2-resident (r = 16 / r = 8) measured 5-14 % slower with it and the 32-channel r = 32 grids (134 MB) 1-4 % slower
  // (gpurun calls 32-33, profiles/r02_conv_sched_ab.txt): they keep one contiguous range per CTA.  LION_CONV_SCHED=0|1 forces.
  { static int sc = -2; if (sc == -2) { const char* e = getenv("LION_CONV_SCHED"); sc = e ? atoi(e) : -1; }
    const double in_bytes = (double)B * Gin * geo.rows * 16.0;
    P.sched = sc >= 0 ? sc : (w.ntaps == 27 && in_bytes > 200e6 ? 1 : 0); }
  P.occ = geo.occ; P.occ_stride = geo.occ_stride;
  { static int ns = -1; if (ns < 0) { const char* e = getenv("LION_TC_NOSKIP"); ns = e ? atoi(e) : 0; } if (ns) P.occ = nullptr; }
  const size_t fixed = 128 * 4 + 8 * 2 * 64 * 4 + 64 * 8 + 128 + tc::OCC_SMEM;
  long long room = 227LL * 1024 - (long long)fixed - (long long)tc::B_STAGES * P.b_stage_bytes;
  int a_stages = (int)(room / P.a_stage_bytes);
  // Sharing an SM with the side stream.  FPS and the neighbour searches run on the side stream during the first
  // ~1 ms of a step (32 CTAs x <= 26 KB of shared memory, latency-bound); a persistent convolution CTA that claims all
  // 227 KB cannot become resident on their SMs, and with the static item split the whole convolution then takes two
  // waves (tools/timeline_step.py: the first PVConv's second convolution 340 us instead of 178).  While the caller
  // flags side-stream work (Ctx::conv_smem_cap) the ring gives up a slot or two -- never below 4 -- and the side
  // kernels opt into the maximum shared-memory carve-out, because an SM only hosts kernels of one carve-out at a time.
  if (c->conv_smem_cap > 0 && a_stages >= 5) {
    int capped = (int)(((long long)c->conv_smem_cap - (long long)fixed - (long long)tc::B_STAGES * P.b_stage_bytes) / P.a_stage_bytes);
    if (capped >= 4 && capped < a_stages) a_stages = capped;
  }
  if (a_stages > tc::MAX_A_STAGES) a_stages = tc::MAX_A_STAGES;
  { static int as = -1; if (as < 0) { const char* e = getenv("LION_TC_ASTAGES"); as = e ? atoi(e) : 0; } if (as > 0 && as < a_stages) a_stages = as; }
  if (a_stages < 2) { set_error("conv_tc: shared memory cannot hold the operand pipeline (N=%d, KG=%d)", NT, KG); return LION_ERR_ARG; }
  P.a_stages = a_stages;
  size_t smem = (size_t)a_stages * P.a_stage_bytes + (size_t)tc::B_STAGES * P.b_stage_bytes + fixed;
  if (smem > 227 * 1024) { set_error("conv_tc: %zu bytes of shared memory needed", smem); return LION_ERR_ARG; }
  // persistent, at most one CTA per SM: the fewest CTAs that still reach the minimal maximum of tiles per CTA
  // (224 tiles on 148 SMs: 112 CTAs x 2 tiles, not 148 CTAs x 1-2 -- every CTA streams the whole weight tensor from L2,
  // and the r = 8 layers are bound by exactly that)
  long long n_units = (long long)ntile * n_tiles_n * B;
  long long per_cta = (n_units + c->num_sms - 1) / c->num_sms;
  int grid = (int)((n_units + per_cta - 1) / per_cta);
#define LION_TC_CASE(kg, tpg_)                                                                            \
  if (KG == kg && tpg == tpg_) {                                                                          \
    static DevOnce attr_once;                                                                         \
    if (attr_once.need()) {                                                                                      \
      LION_CHECK_CUDA(cudaFuncSetAttribute(tc::k_conv_tc<kg, tpg_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
    }                                                                                                     \
    LION_LAUNCH(c, (tc::k_conv_tc<kg, tpg_>), grid, tc::THREADS, smem, P);                                \
  }
  LION_TC_CASE(2, 1) LION_TC_CASE(4, 1) LION_TC_CASE(8, 1) LION_TC_CASE(2, 9) LION_TC_CASE(4, 9) LION_TC_CASE(8, 9)
#undef LION_TC_CASE
  return check_launch(c, "conv_tc");
}

}  // namespace lion
