// lion_b200 -- the global-latent prior (PriorSEDrop / PriorSEClip): a 2048-wide residual MLP
// with squeeze-excite cells evaluated on a handful of rows (B <= 64 shapes).
//
// Reference: models/score_sde/resnet.py:195-218 (Prior.forward), :60-90 (ResBlockSEDrop),
// :29-56 (ResBlockSEClip), :16-27 (SE), models/utils.py:16-31 (PositionalEmbedding).
// The work is weight streaming (309 MB of fp32 weights per step for 0.15 GFLOP/shape), so the
// kernel is a skinny GEMM: every weight is read once, coalesced, and applied to all B rows
// held in shared memory; bias / ReLU / residual / SE gate are fused into the epilogue.
#include "common.cuh"
#include "model.cuh"
#include "../../include/lion_b200.h"
#include <cstdlib>

namespace lion {

constexpr int GP_MAXB = 32;     // rows (shapes) per call
constexpr int GP_BT = 32;
constexpr int GP_KS = 256;      // K slice per block (split-K)
constexpr int GP_PITCH = GP_KS + 4;
constexpr int GP_WARPS = 8;
constexpr int GP_OW = 16;       // outputs per warp (processed two at a time)
constexpr int GP_OB = GP_WARPS * GP_OW;   // 128 outputs per block
constexpr int GP_MAXSPLIT = 16;

__device__ __forceinline__ float warp_transpose_sum32(float* v, int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float keep = upper ? v[i + half] : v[i];
      float send = upper ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// Split-K skinny GEMM, phase 1:  part[ks][b][o] = sum_{k in slice ks} W[o][k] * (x[b][k] + add[b][k])
// grid = (O/128, K/256).  M = 32 shapes is far below a tcgen05 tile (M >= 128 rows of the SAME
// operand), so this layer uses warp-level mma.sync.m16n8k8 TF32 (the reference's cuDNN 1x1
// convolutions run TF32 as well): A = 16 weight rows x 8 k, B = 8 k x 8 shapes.
//   * the block's weight tile [128 x 256] (128 KB) streams HBM -> shared memory with cp.async in
//     four 64-column commit groups, so the tensor work on group g overlaps the arrival of g+1;
//   * the activation slice [32 x 256] is staged once per block (rounded to TF32, round-to-nearest);
//   * row pitch 260 floats makes every fragment load bank-conflict free (bank = 4*row + col).
// Deterministic: partials are summed in a fixed order by k_gp_reduce.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t tf32_bits(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
struct GpEpi {
  const float* bias; float* out; int out_stride; const float* mul; int mul_stride; const float* res; int res_stride; int act;
};
__global__ void __launch_bounds__(GP_WARPS * 32, 1)
k_gp_partial(const float* __restrict__ W, const float* __restrict__ x, int x_stride, const float* __restrict__ add,
             int add_stride, float* __restrict__ part, int B, int K, int O, GpEpi epi) {
  pdl_prologue();
  extern __shared__ __align__(16) float s_mem[];
  float* s_w = s_mem;                          // [GP_OB][GP_PITCH]
  float* s_x = s_mem + GP_OB * GP_PITCH;       // [GP_BT][GP_PITCH]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k0 = blockIdx.y * GP_KS, kt = min(GP_KS, K - k0);
  const int o0 = blockIdx.x * GP_OB;
  // weights: 4 commit groups of 64 columns; chunk c of a group = (row, 16-byte column piece)
  const uint32_t s_w_addr = (uint32_t)__cvta_generic_to_shared(s_w);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
#pragma unroll
    for (int u = 0; u < (GP_OB * 16) / (GP_WARPS * 32); ++u) {       // 128 rows x 16 pieces / 256 threads = 8
      int c = tid + u * (GP_WARPS * 32);
      int row = c >> 4, piece = c & 15;
      int col = g * 64 + piece * 4;
      if (o0 + row < O && col < kt)
        cp_async16(s_w_addr + (uint32_t)(row * GP_PITCH + col) * 4u, W + (size_t)(o0 + row) * K + k0 + col);
      else   // keep out-of-range rows / columns finite: they meet zero activations or unused outputs
        *reinterpret_cast<float4*>(s_w + row * GP_PITCH + col) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // activations (+ add), rounded to TF32
  {
    constexpr int PER_THREAD = GP_BT * (GP_KS / 4) / (GP_WARPS * 32);      // 8
    float4 v[PER_THREAD];
#pragma unroll
    for (int u = 0; u < PER_THREAD; ++u) {
      int i = tid + u * (GP_WARPS * 32);
      int b = i / (GP_KS / 4), k4 = i % (GP_KS / 4);
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < B && k4 * 4 < kt) {
        v[u] = *reinterpret_cast<const float4*>(x + (size_t)b * x_stride + k0 + k4 * 4);
        if (add) {
          float4 a = *reinterpret_cast<const float4*>(add + (size_t)b * add_stride + k0 + k4 * 4);
          v[u].x += a.x; v[u].y += a.y; v[u].z += a.z; v[u].w += a.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < PER_THREAD; ++u) {
      int i = tid + u * (GP_WARPS * 32);
      int b = i / (GP_KS / 4), k4 = i % (GP_KS / 4);
      float4 r = make_float4(__uint_as_float(tf32_bits(v[u].x)), __uint_as_float(tf32_bits(v[u].y)),
                             __uint_as_float(tf32_bits(v[u].z)), __uint_as_float(tf32_bits(v[u].w)));
      *reinterpret_cast<float4*>(s_x + b * GP_PITCH + k4 * 4) = r;
    }
  }
  const int g8 = lane >> 2, t4 = lane & 3;
  float acc[4][4];
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[n][i] = 0.0f;
  const float* wa = s_w + (wid * 16 + g8) * GP_PITCH + t4;       // rows g8 / g8+8 of this warp's 16 outputs
  const float* xb = s_x + g8 * GP_PITCH + t4;                    // shape g8 of each 8-shape tile
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g == 0) asm volatile("cp.async.wait_group 3;" ::: "memory");
    else if (g == 1) asm volatile("cp.async.wait_group 2;" ::: "memory");
    else if (g == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (g * 64 < kt) {
#pragma unroll
      for (int kk = 0; kk < 64; kk += 8) {
        const int k = g * 64 + kk;
        uint32_t a0 = __float_as_uint(wa[k]), a1 = __float_as_uint(wa[8 * GP_PITCH + k]);
        uint32_t a2 = __float_as_uint(wa[k + 4]), a3 = __float_as_uint(wa[8 * GP_PITCH + k + 4]);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          uint32_t b0 = __float_as_uint(xb[n * 8 * GP_PITCH + k]), b1 = __float_as_uint(xb[n * 8 * GP_PITCH + k + 4]);
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(acc[n][0]), "+f"(acc[n][1]), "+f"(acc[n][2]), "+f"(acc[n][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
      }
    }
  }
  // C fragment: c0,c1 -> (row g8, shapes 2*t4, 2*t4+1); c2,c3 -> (row g8+8, same shapes)
  float* pout = part + (size_t)blockIdx.y * B * O;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int o = o0 + wid * 16 + g8 + (i >= 2 ? 8 : 0);
      int b = n * 8 + 2 * t4 + (i & 1);
      if (b < B && o < O) pout[(size_t)b * O + o] = acc[n][i];
    }
  }
}

// phase 2: out[b][o] = epi( sum_ks part[ks][b][o] + bias[o] );  act: 0 none, 1 relu, 2 sigmoid;
// mul: result *= mul[b][o] (SE gate);  res: result += res[b][o] (residual shortcut)
__global__ void k_gp_reduce(const float* __restrict__ part, int nsplit, const float* __restrict__ bias, float* __restrict__ out,
                            int out_stride, const float* __restrict__ mul, int mul_stride, const float* __restrict__ res,
                            int res_stride, int B, int O, int act) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * O) return;
  int b = i / O, o = i % O;
  float v = 0.0f;
  for (int s = 0; s < nsplit; ++s) v += part[((size_t)s * B + b) * O + o];
  v += bias ? bias[o] : 0.0f;
  if (act == 1) v = fmaxf(v, 0.0f);
  else if (act == 2) v = 1.0f / (1.0f + expf(-v));
  if (mul) v *= mul[(size_t)b * mul_stride + o];
  if (res) v += res[(size_t)b * res_stride + o];
  out[(size_t)b * out_stride + o] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// The whole network as ONE persistent kernel (round 2, second design; the first -- no split-K, every CTA re-reading
// the full activation matrix -- lost to the two-kernels-per-Linear form: profiles/r02_global_prior_persistent_ab.txt).
//   * grid = 128 CTAs in thread-block clusters of CL (8, 4 or 2; the largest size whose clusters are all co-resident:
//     a B200 hosts only 15 clusters of 8 CTAs with this much shared memory, so CL = 4 is what runs there).  A cluster
//     owns 16 * CL output rows of every Linear.  K is always split 8 ways -- the SAME 256-wide slices and summation
//     order as k_gp_partial / k_gp_reduce for K = 2048 --: CL ways across the cluster's CTAs and HS = 8 / CL ways across
//     the warp groups of a CTA, so each CTA reads only its [32 shapes x K / CL] slice of the activations.
//   * warp w = (row group w % CL, k-part w / CL): 16 rows x K / 8 columns of the weight matrix, streamed with
//     cp.async.bulk into two private ring slots ([16 rows x <= 128 columns] each); a slot is refilled with the warp's
//     NEXT stage -- usually the next layer's rows -- the moment it has been consumed, so the HBM stream (309 MB per
//     evaluation) runs ahead across the grid barriers: weights never wait for activations.
//   * split-K reduction through distributed shared memory: every CTA leaves HS partial tiles [32 x 16 CL] in its own
//     shared memory, one barrier.cluster, then CTA r sums columns [16r, 16r+16) over the 8 partials with
//     ld.shared::cluster (fixed order -> bit-reproducible), applies bias / ReLU / sigmoid / SE gate / residual and
//     stores the final values.
//   * one grid barrier per Linear (monotonic counter, zeroed by k_gp_posemb): 36 instead of 73 kernel boundaries.
// All 128 CTAs must be co-resident (checked once per device with cudaOccupancyMaxActiveClusters; otherwise the
// two-kernel form runs).  LION_GP_PERSIST=0 selects the two-kernel form, LION_GP_PERSIST=8|4|2 pins the cluster size.
namespace gpp {
constexpr int NCTA = 128;             // 16 output rows x 8 K slices per CTA-warp; 8 warps per CTA
constexpr int CW = 128;               // columns per weight stage
constexpr int WP = CW + 4;            // stage row pitch in floats (bank = 4 * row + column: conflict-free fragments)
constexpr int STAGE_FLOATS = 16 * WP; // one stage = one warp's 16 rows
constexpr int NSLOT = 16;             // two per warp
constexpr int XW = 512;               // activation columns staged per pass (all k-parts of the CTA together)
constexpr int XP = XW + 4;
constexpr int THREADS = 256;
constexpr int MAXL = 38;
constexpr int PART_FLOATS = 4 * 32 * (32 + 4);   // HS tiles of [32][16 CL + 4] floats: 4224 / 4352 / 4608 for CL = 8 / 4 / 2
constexpr size_t SMEM = (size_t)(NSLOT * STAGE_FLOATS + 32 * XP + PART_FLOATS) * sizeof(float) + NSLOT * 8;

struct Layer {
  const float* w; const float* bias; const float* x; const float* add; float* out; float* out2; const float* mul; const float* res;
  int K, O, xs, as, os, os2, ms, rs, act, pad;
};
struct Prog { int nl, B; unsigned* counter; Layer l[MAXL]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded waits: a protocol bug traps (CUDA error) instead of hanging the GPU
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > 2000000000LL) __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
  __syncwarp();
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// every CTA of the grid has finished the previous layer (its global stores included)
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    long long t0 = clock64();
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > 2000000000LL) __trap();
    }
  }
  __syncthreads();
}

struct Cur { int l, cs; };     // (layer, column stage) of a warp's weight stream

template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1) k_gp_persist(const __grid_constant__ Prog P) {
  constexpr int HS = 8 / CL;            // k-parts per CTA
  constexpr int TO = 16 * CL;           // output rows per cluster
  constexpr int PP = TO + 4;            // partial-tile row pitch
  constexpr int PW = XW / HS;           // activation columns per k-part and pass
  extern __shared__ __align__(128) float gsm[];
  float* s_ring = gsm;                                   // [NSLOT][16][WP]
  float* s_x = s_ring + NSLOT * STAGE_FLOATS;            // [32][XP]  activations slice of one pass: HS k-parts x PW columns, TF32-rounded
  float* s_part = s_x + 32 * XP;                         // [HS][32][PP]  this CTA's partial sums [k-part][shape][output]
  uint64_t* s_bar = (uint64_t*)(s_part + PART_FLOATS);   // [NSLOT]   "stage landed"
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int rw = w % CL, h = w / CL;                     // this warp's row group and k-part
  const int cr = blockIdx.x % CL, cid = blockIdx.x / CL;
  const int ks = cr * HS + h;                            // this warp's K slice (0..7)
  const int B = P.B;
  const uint32_t bar0 = smem_u32(s_bar), ring0 = smem_u32(s_ring);
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto active = [&](int l) { return cid * TO < P.l[l].O; };
  auto ncs_of = [&](int l) { return (P.l[l].K / 8 + CW - 1) / CW; };
  // warp-collective: start the copy of stage `c` (this warp's 16 rows x <= 128 columns) as fetch number fi of this warp
  auto fetch = [&](Cur c, int fi) {
    const Layer& L = P.l[c.l];
    const int kw = L.K / 8, cw = min(CW, kw - c.cs * CW);
    const int slot = (fi & 1) * 8 + w;
    const uint32_t bar = bar0 + 8 * slot;
    const uint32_t bytes = (uint32_t)cw * 4u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // this slot was read through the generic proxy
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * 16u) : "memory");
    __syncwarp();
    if (lane < 16) {
      const float* src = L.w + (size_t)(cid * TO + rw * 16 + lane) * L.K + ks * kw + c.cs * CW;
      const uint32_t dst = ring0 + (uint32_t)(slot * STAGE_FLOATS + lane * WP) * 4u;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
    }
  };
  auto advance = [&](Cur& c) {
    if (c.l >= P.nl) return;
    if (++c.cs >= ncs_of(c.l)) {
      c.cs = 0;
      do { ++c.l; } while (c.l < P.nl && !active(c.l));
    }
  };
  Cur fc{0, 0};
  while (fc.l < P.nl && !active(fc.l)) ++fc.l;
  int fi = 0, ci = 0;
  for (; fi < 2 && fc.l < P.nl; ++fi) { fetch(fc, fi); advance(fc); }

  const int g8 = lane >> 2, t4 = lane & 3;
  for (int l = 0; l < P.nl; ++l) {
    const Layer& L = P.l[l];
    if (l > 0) grid_barrier(P.counter, (unsigned)l * gridDim.x);
    if (!active(l)) continue;                            // (the whole cluster skips together)
    const int kw = L.K / 8;                              // columns per K slice
    float acc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[n][i] = 0.0f;
    for (int p0 = 0; p0 < kw; p0 += PW) {                // passes: PW columns of every k-part at a time
      const int pw = min(PW, kw - p0), pw4 = pw >> 2;
      if (p0 > 0) __syncthreads();                       // the previous pass has been consumed
      // activations [32][HS][pw] (+ add), rounded to TF32; .cg loads: other CTAs wrote these buffers in this launch
      for (int i = tid; i < 32 * HS * pw4; i += THREADS) {
        const int k4 = i % pw4, hb = i / pw4, hh = hb % HS, b = hb / HS;
        const int col = (cr * HS + hh) * kw + p0;        // first column of k-part hh in this pass
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < B) {
          v = __ldcg(reinterpret_cast<const float4*>(L.x + (size_t)b * L.xs + col) + k4);
          if (L.add) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(L.add + (size_t)b * L.as + col) + k4);
            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
          }
        }
        *reinterpret_cast<float4*>(s_x + b * XP + hh * PW + k4 * 4) =
            make_float4(__uint_as_float(tf32_bits(v.x)), __uint_as_float(tf32_bits(v.y)), __uint_as_float(tf32_bits(v.z)),
                        __uint_as_float(tf32_bits(v.w)));
      }
      __syncthreads();
      for (int c0 = 0; c0 < pw; c0 += CW, ++ci) {
        const int cw = min(CW, pw - c0);
        const int slot = (ci & 1) * 8 + w;
        mbar_wait(bar0 + 8 * slot, (uint32_t)(ci >> 1) & 1u);
        const float* wa = s_ring + slot * STAGE_FLOATS + g8 * WP + t4;     // rows g8 / g8+8 of this warp's 16 outputs
        const float* xb = s_x + g8 * XP + h * PW + c0 + t4;                // shape g8 of each 8-shape tile
#pragma unroll 4
        for (int k = 0; k < cw; k += 8) {
          const uint32_t a0 = __float_as_uint(wa[k]), a1 = __float_as_uint(wa[8 * WP + k]);
          const uint32_t a2 = __float_as_uint(wa[k + 4]), a3 = __float_as_uint(wa[8 * WP + k + 4]);
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const uint32_t b0 = __float_as_uint(xb[n * 8 * XP + k]), b1 = __float_as_uint(xb[n * 8 * XP + k + 4]);
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[n][0]), "+f"(acc[n][1]), "+f"(acc[n][2]), "+f"(acc[n][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
          }
        }
        __syncwarp();
        if (fc.l < P.nl) { fetch(fc, fi); advance(fc); ++fi; }             // refill the slot just consumed
      }
    }
    // C fragment: c0,c1 -> (row g8, shapes 2*t4, 2*t4+1); c2,c3 -> (row g8+8, same shapes)
    {
      float* sp = s_part + h * (32 * PP);
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          sp[(n * 8 + 2 * t4 + (i & 1)) * PP + rw * 16 + g8 + (i >= 2 ? 8 : 0)] = acc[n][i];
    }
    cluster_sync_all();
    {
      // CTA cr: columns [16 cr, 16 cr + 16) of the cluster's tile; thread -> (shape b, two adjacent outputs);
      // the 8 partials are summed in K-slice order (peer CTA major, k-part minor)
      const int b = tid >> 3, oc = cr * 16 + (tid & 7) * 2;
      const uint32_t local = smem_u32(s_part + b * PP + oc);
      float v0 = 0.0f, v1 = 0.0f;
#pragma unroll
      for (int p = 0; p < CL; ++p) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(p));
#pragma unroll
        for (int hh = 0; hh < HS; ++hh) {
          float x0, x1;
          asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(x0), "=f"(x1) : "r"(remote + (uint32_t)(hh * 32 * PP * 4)) : "memory");
          v0 += x0; v1 += x1;
        }
      }
      const int o = cid * TO + oc;
      if (b < B) {
        if (L.bias) { v0 += __ldg(L.bias + o); v1 += __ldg(L.bias + o + 1); }
        if (L.act == 1) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
        else if (L.act == 2) { v0 = 1.0f / (1.0f + expf(-v0)); v1 = 1.0f / (1.0f + expf(-v1)); }
        if (L.mul) { const float2 m = __ldcg(reinterpret_cast<const float2*>(L.mul + (size_t)b * L.ms + o)); v0 *= m.x; v1 *= m.y; }
        if (L.res) { const float2 r = __ldcg(reinterpret_cast<const float2*>(L.res + (size_t)b * L.rs + o)); v0 += r.x; v1 += r.y; }
        *reinterpret_cast<float2*>(L.out + (size_t)b * L.os + o) = make_float2(v0, v1);
        if (L.out2) *reinterpret_cast<float2*>(L.out2 + (size_t)b * L.os2 + o) = make_float2(v0, v1);
      }
    }
    // (the next layer's grid barrier also orders these s_part reads before anyone overwrites its tile)
  }
  cluster_sync_all();      // no CTA leaves while a peer may still read its shared memory
}
}  // namespace gpp

// PositionalEmbedding (models/utils.py:16-31): fp32 frequencies exp(i * -log(1e4)/(half-1))
__global__ void k_gp_posemb(const float* __restrict__ t, const float* __restrict__ freqs, float* __restrict__ out,
                            int half, float scale, unsigned* __restrict__ zero_me) {
  pdl_prologue();
  int b = blockIdx.x, i = threadIdx.x;
  if (zero_me && b == 0 && i == 0) *zero_me = 0u;      // grid-barrier counter of the persistent kernel that follows
  if (i >= half) return;
  float e = __fmul_rn(__fmul_rn(t[b], scale), freqs[i]);
  out[(size_t)b * 2 * half + i] = sinf(e);
  out[(size_t)b * 2 * half + half + i] = cosf(e);
}

struct GPLin { const float* w; const float* b; int K, O; };
struct GlobalPriorBlk {
  int D = 128, nf = 2048, emb = 128, ncell = 8, clip = 0, clip_dim = 512;
  float scale = 1.0f;
  float* d_freqs = nullptr;
  GPLin t0, t1, cmap, in, outl;
  struct Cell { GPLin c1, c2, se0, se2; };
  std::vector<Cell> cells;
};
void global_prior_free(GlobalPriorBlk* g) { delete g; }

// desc: [D, nf, emb_dim, ncell, clip, clip_dim, scale_bits]; params in state_dict order:
//   [clip_feat_mapping.w,b] temb_layer.0.w,b temb_layer.1.w,b input_layer.w,b
//   all_modules.k.{conv1.w,b conv2.w,b SE.fc.0.w SE.fc.2.w} output_layer.w,b
int global_prior_build(Model* m, Cursor& cur) {
  const std::vector<int>& d = m->desc;
  if (d.size() < 7) { set_error("global prior descriptor: [D, nf, emb, ncell, clip, clip_dim, scale_bits]"); return LION_ERR_ARG; }
  GlobalPriorBlk* g = new GlobalPriorBlk();
  m->gp = g;
  g->D = d[0]; g->nf = d[1]; g->emb = d[2]; g->ncell = d[3]; g->clip = d[4]; g->clip_dim = d[5];
  memcpy(&g->scale, &d[6], 4);
  auto lin = [&](GPLin& l, int K, int O, bool bias) { l.w = cur.next(); l.b = bias ? cur.next() : nullptr; l.K = K; l.O = O; };
  if (g->clip) lin(g->cmap, g->clip_dim, g->nf, true);
  lin(g->t0, g->emb, g->emb * 4, true);
  lin(g->t1, g->emb * 4, g->nf, true);
  lin(g->in, g->D, g->nf, true);
  g->cells.resize(g->ncell);
  for (auto& c : g->cells) {
    lin(c.c1, g->clip ? 2 * g->nf : g->nf, g->nf, true);
    lin(c.c2, g->nf, g->nf, true);
    lin(c.se0, g->nf, g->nf / 8, false);
    lin(c.se2, g->nf / 8, g->nf, false);
  }
  lin(g->outl, g->nf, g->D, true);
  if (cur.bad) { set_error("global prior: parameter list too short (%d given)", cur.n); return LION_ERR_ARG; }
  int half = g->emb / 2;
  std::vector<float> fr(half);
  float step = (float)(std::log(10000.0) / (half - 1));      // python float -> fp32 tensor multiply
  for (int i = 0; i < half; ++i) fr[i] = expf((float)i * -step);
  LION_TRY(m->dmalloc(&g->d_freqs, (size_t)half));
  LION_CHECK_CUDA(cudaMemcpy(g->d_freqs, fr.data(), half * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

static int gp_linear(Ctx* c, const GPLin& l, const float* x, int xs, const float* add, int as, float* out, int os,
                     const float* mul, int ms, const float* res, int rs, int B, int act) {
  if (l.K % 4) { set_error("global prior: K=%d must be a multiple of 4", l.K); return LION_ERR_ARG; }
  int nsplit = cdiv(l.K, GP_KS);
  if (nsplit > GP_MAXSPLIT) { set_error("global prior: K=%d too large", l.K); return LION_ERR_ARG; }
  const size_t smem = (size_t)(GP_OB + GP_BT) * GP_PITCH * sizeof(float);
  static DevOnce attr_once;
  if (attr_once.need()) LION_CHECK_CUDA(cudaFuncSetAttribute(k_gp_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GpEpi epi{l.b, out, os, mul, ms, res, rs, act};
  size_t mk = c->mark();
  float* part = c->alloc_n<float>((size_t)nsplit * B * l.O);
  LION_LAUNCH(c, k_gp_partial, dim3(cdiv(l.O, GP_OB), nsplit), GP_WARPS * 32, smem, l.w, x, xs, add, as, part, B, l.K, l.O, epi);
  LION_LAUNCH(c, k_gp_reduce, cdiv(B * l.O, 256), 256, 0, part, nsplit, l.b, out, os, mul, ms, res, rs, B, l.O, act);
  c->release(mk);     // stream order makes reuse by the next layer safe
  return 0;
}

// one chunk of <= 32 shapes: two kernels per Linear (split-K partial sums + deterministic reduce with fused epilogue)
static int global_prior_forward_layers(Model* m, const float* x, const float* t, const float* clip, float* out, int B) {
  GlobalPriorBlk* g = m->gp;
  Ctx* c = m->ctx;
  int nf = g->nf, tw = g->clip ? 2 * nf : nf;
  float* pe = c->alloc_n<float>((size_t)B * g->emb);
  float* t0 = c->alloc_n<float>((size_t)B * g->emb * 4);
  float* tadd = c->alloc_n<float>((size_t)B * tw);     // [temb | 0]: what is added to the cell input
  float* cat = c->alloc_n<float>((size_t)B * tw);      // [h | clip-mapped] (clip variant only)
  float* h = c->alloc_n<float>((size_t)B * nf);
  float* h2 = c->alloc_n<float>((size_t)B * nf);
  float* a = c->alloc_n<float>((size_t)B * nf);
  float* bb = c->alloc_n<float>((size_t)B * nf);
  float* s0 = c->alloc_n<float>((size_t)B * nf / 8);
  LION_LAUNCH(c, k_gp_posemb, B, 64, 0, t, g->d_freqs, pe, g->emb / 2, g->scale, (unsigned*)nullptr);
  // temb_layer: two 1x1 convs, no nonlinearity in between (resnet.py:181-184)
  LION_TRY(gp_linear(c, g->t0, pe, g->emb, nullptr, 0, t0, g->emb * 4, nullptr, 0, nullptr, 0, B, 0));
  if (g->clip) LION_TRY(memset_async(c, tadd, 0, sizeof(float) * B * tw));
  LION_TRY(gp_linear(c, g->t1, t0, g->emb * 4, nullptr, 0, tadd, tw, nullptr, 0, nullptr, 0, B, 0));
  // clip_feat_mapping output is concatenated behind temb (resnet.py:203-208) and reaches every
  // cell's conv1 un-added (ResBlockSEClip.forward, resnet.py:41-46)
  if (g->clip) LION_TRY(gp_linear(c, g->cmap, clip, g->clip_dim, nullptr, 0, cat + nf, tw, nullptr, 0, nullptr, 0, B, 0));
  LION_TRY(gp_linear(c, g->in, x, g->D, nullptr, 0, h, nf, nullptr, 0, nullptr, 0, B, 0));
  for (auto& cell : g->cells) {
    // conv1(x + t [| clip]) -> ReLU -> (dropout: identity in eval) -> conv2 -> ReLU -> SE -> + x
    if (g->clip) {
      if (!c->dry)
        LION_CHECK_CUDA(cudaMemcpy2DAsync(cat, tw * sizeof(float), h, nf * sizeof(float), nf * sizeof(float), B, cudaMemcpyDeviceToDevice, c->stream));
      LION_TRY(gp_linear(c, cell.c1, cat, tw, tadd, tw, a, nf, nullptr, 0, nullptr, 0, B, 1));
    } else {
      LION_TRY(gp_linear(c, cell.c1, h, nf, tadd, tw, a, nf, nullptr, 0, nullptr, 0, B, 1));
    }
    LION_TRY(gp_linear(c, cell.c2, a, nf, nullptr, 0, bb, nf, nullptr, 0, nullptr, 0, B, 1));
    LION_TRY(gp_linear(c, cell.se0, bb, nf, nullptr, 0, s0, nf / 8, nullptr, 0, nullptr, 0, B, 1));
    LION_TRY(gp_linear(c, cell.se2, s0, nf / 8, nullptr, 0, h2, nf, bb, nf, h, nf, B, 2));   // sigmoid(.) * bb + h
    float* tmp = h; h = h2; h2 = tmp;
  }
  LION_TRY(gp_linear(c, g->outl, h, nf, nullptr, 0, out, g->D, nullptr, 0, nullptr, 0, B, 0));
  return check_launch(c, "global_prior_forward");
}

// The persistent form is usable when every Linear fits the kernel's tiling and all 128 CTAs can be co-resident.
// Returns the cluster size to launch (8, 4 or 2), or 0 for the two-kernel form.
template <int CL>
static int gp_clusters_resident() {
  if (cudaFuncSetAttribute(gpp::k_gp_persist<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gpp::SMEM) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gpp::NCTA); cfg.blockDim = dim3(gpp::THREADS); cfg.dynamicSmemBytes = gpp::SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gpp::k_gp_persist<CL>, &cfg) != cudaSuccess) n = 0;
  return n;
}
static int gp_persist_cluster(const GlobalPriorBlk* g) {
  static int want = -1;
  if (want < 0) { const char* e = getenv("LION_GP_PERSIST"); want = e ? atoi(e) : 1; }
  if (!want) return 0;
  auto ok = [](const GPLin& l) { return l.K % 64 == 0 && l.K <= 8 * 4096 && l.O % 128 == 0 && l.O <= 2048; };
  bool all = ok(g->t0) && ok(g->t1) && ok(g->in) && ok(g->outl) && (!g->clip || ok(g->cmap));
  for (auto& c : g->cells) all = all && ok(c.c1) && ok(c.c2) && ok(c.se0) && ok(c.se2);
  if (!all || 4 + 4 * (int)g->cells.size() + (g->clip ? 1 : 0) > gpp::MAXL) return 0;
  // co-residency of all clusters, per device (a grid barrier deadlocks otherwise): the largest cluster size that fits
  static int chosen[64];
  static bool asked[64];
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return 0;
  if (!asked[d]) {
    asked[d] = true;
    chosen[d] = 0;
    const int n8 = gp_clusters_resident<8>(), n4 = gp_clusters_resident<4>(), n2 = gp_clusters_resident<2>();
    (void)cudaGetLastError();
    if (n8 >= gpp::NCTA / 8 && (want == 1 || want == 8)) chosen[d] = 8;
    else if (n4 >= gpp::NCTA / 4 && (want == 1 || want == 4)) chosen[d] = 4;
    else if (n2 >= gpp::NCTA / 2 && (want == 1 || want == 2)) chosen[d] = 2;
    if (getenv("LION_VERBOSE") || !chosen[d])
      fprintf(stderr, "lion_b200: global prior on device %d: co-resident clusters %d x8, %d x4, %d x2 -> %s\n", d, n8, n4, n2,
              chosen[d] == 8 ? "persistent, clusters of 8" : chosen[d] == 4 ? "persistent, clusters of 4"
              : chosen[d] == 2 ? "persistent, clusters of 2" : "two-kernel form");
  }
  return chosen[d];
}

// one chunk of <= 32 shapes through the persistent kernel (k_gp_posemb zeroes its barrier counter)
static int global_prior_forward_persist(Model* m, const float* x, const float* t, const float* clip, float* out, int B, int CL) {
  GlobalPriorBlk* g = m->gp;
  Ctx* c = m->ctx;
  const int nf = g->nf, tw = g->clip ? 2 * nf : nf;
  float* pe = c->alloc_n<float>((size_t)B * g->emb);
  float* t0 = c->alloc_n<float>((size_t)B * g->emb * 4);
  float* tadd = c->alloc_n<float>((size_t)B * tw);     // [temb | 0]: what is added to the cell input
  float* cat = c->alloc_n<float>((size_t)B * tw);      // [h | clip-mapped] (clip variant only)
  float* h = c->alloc_n<float>((size_t)B * nf);
  float* h2 = c->alloc_n<float>((size_t)B * nf);
  float* a = c->alloc_n<float>((size_t)B * nf);
  float* bb = c->alloc_n<float>((size_t)B * nf);
  float* s0 = c->alloc_n<float>((size_t)B * nf / 8);
  unsigned* counter = c->alloc_n<unsigned>(1);
  LION_LAUNCH(c, k_gp_posemb, B, 64, 0, t, g->d_freqs, pe, g->emb / 2, g->scale, counter);
  if (g->clip) LION_TRY(memset_async(c, tadd, 0, sizeof(float) * B * tw));
  gpp::Prog P;
  memset(&P, 0, sizeof(P));
  P.B = B; P.counter = counter;
  auto add = [&](const GPLin& l, const float* xin, int xs, const float* ad, int as, float* o, int os, float* o2, int os2,
                 const float* mul, int ms, const float* res, int rs, int act) {
    gpp::Layer& L = P.l[P.nl++];
    L.w = l.w; L.bias = l.b; L.x = xin; L.add = ad; L.out = o; L.out2 = o2; L.mul = mul; L.res = res;
    L.K = l.K; L.O = l.O; L.xs = xs; L.as = as; L.os = os; L.os2 = os2; L.ms = ms; L.rs = rs; L.act = act;
  };
  // same layer sequence as global_prior_forward_layers; the clip variant's copy of h into [h | clip] is a second store
  add(g->t0, pe, g->emb, nullptr, 0, t0, g->emb * 4, nullptr, 0, nullptr, 0, nullptr, 0, 0);
  add(g->t1, t0, g->emb * 4, nullptr, 0, tadd, tw, nullptr, 0, nullptr, 0, nullptr, 0, 0);
  if (g->clip) add(g->cmap, clip, g->clip_dim, nullptr, 0, cat + nf, tw, nullptr, 0, nullptr, 0, nullptr, 0, 0);
  add(g->in, x, g->D, nullptr, 0, h, nf, g->clip ? cat : nullptr, tw, nullptr, 0, nullptr, 0, 0);
  for (auto& cell : g->cells) {
    if (g->clip) add(cell.c1, cat, tw, tadd, tw, a, nf, nullptr, 0, nullptr, 0, nullptr, 0, 1);
    else add(cell.c1, h, nf, tadd, tw, a, nf, nullptr, 0, nullptr, 0, nullptr, 0, 1);
    add(cell.c2, a, nf, nullptr, 0, bb, nf, nullptr, 0, nullptr, 0, nullptr, 0, 1);
    add(cell.se0, bb, nf, nullptr, 0, s0, nf / 8, nullptr, 0, nullptr, 0, nullptr, 0, 1);
    add(cell.se2, s0, nf / 8, nullptr, 0, h2, nf, g->clip ? cat : nullptr, tw, bb, nf, h, nf, 2);   // sigmoid(.) * bb + h
    float* tmp = h; h = h2; h2 = tmp;
  }
  add(g->outl, h, nf, nullptr, 0, out, g->D, nullptr, 0, nullptr, 0, nullptr, 0, 0);
  if (!c->dry) {
    if (CL == 8) gpp::k_gp_persist<8><<<gpp::NCTA, gpp::THREADS, gpp::SMEM, c->stream>>>(P);
    else if (CL == 4) gpp::k_gp_persist<4><<<gpp::NCTA, gpp::THREADS, gpp::SMEM, c->stream>>>(P);
    else gpp::k_gp_persist<2><<<gpp::NCTA, gpp::THREADS, gpp::SMEM, c->stream>>>(P);
    c->launches++;
  }
  return check_launch(c, "global_prior_forward (persistent)");
}

int global_prior_forward(Model* m, const float* x, const float* t, const float* clip, float* out, int B) {
  GlobalPriorBlk* g = m->gp;
  Ctx* c = m->ctx;
  if (g->clip && !clip) { set_error("global prior: this network needs clip_feat"); return LION_ERR_ARG; }
  // any batch size: chunks of 32 shapes (the reference takes any B, resnet.py:195-218)
  for (int b0 = 0; b0 < B; b0 += GP_MAXB) {
    const int nb = B - b0 < GP_MAXB ? B - b0 : GP_MAXB;
    const size_t mk = c->mark();
    const float* xc = x + (size_t)b0 * g->D;
    const float* cc = clip ? clip + (size_t)b0 * g->clip_dim : nullptr;
    float* oc = out + (size_t)b0 * g->D;
    const int CL = gp_persist_cluster(g);
    if (CL) LION_TRY(global_prior_forward_persist(m, xc, t + b0, cc, oc, nb, CL));
    else LION_TRY(global_prior_forward_layers(m, xc, t + b0, cc, oc, nb));
    c->release(mk);
  }
  return 0;
}

}  // namespace lion
