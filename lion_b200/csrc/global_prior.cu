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

// =====================================================================================
// Persistent form: the whole network (36 Linears) as ONE cooperative kernel per denoising step.
// Round 1 ran it as 75 graph nodes (~5 us each: launch-latency bound, 0.37 ms/step against a 47 us weight-streaming
// bound).  Here 128 CTAs stay resident and walk a device-side layer table; phases (layers that depend on each other)
// are separated by a grid barrier (sense-reversing, one atomic per CTA).
//   * no split-K: a CTA owns 16 output rows of a layer over the full K, so there is no cross-CTA reduction and the
//     summation order is fixed (bit-reproducible);  mma.sync.m16n8k8 TF32, A = 16 weight rows, B = the 32 shapes;
//     the 8 warps split the k-steps and fold their partial tiles through shared memory in warp order;
//   * W [16 x 256] and x [32 x 256] chunks stream through a 4-deep cp.async ring; the weights of the NEXT phase do
//     not depend on the barrier, so their first four chunks are issued BEFORE waiting on it and the activations
//     after: HBM streaming continues across the barrier;
//   * the epilogue (bias / ReLU / sigmoid gate * mul + residual, optional second output out2 = out + add2, i.e. the
//     `x + temb` that the next cell's conv1 consumes, resnet.py:78) is fused.
// =====================================================================================
constexpr int GPP_ROWS = 16;                 // output rows per work item
constexpr int GPP_KC = 256;                  // K chunk
constexpr int GPP_PITCH = GPP_KC + 4;
constexpr int GPP_STAGES = 4;
constexpr int GPP_THREADS = 256;
constexpr int GPP_MAXLAYERS = 48;
constexpr int GPP_STAGE_FLOATS = (GPP_ROWS + GP_BT) * GPP_PITCH;

struct GpLayer {
  const float* W; const float* bias;
  const float* x; int xs;
  float* out; int os;
  const float* mul; int ms;          // out *= mul (after the activation)
  const float* res; int rs;          // out += res
  float* out2; int o2s; const float* add2; int a2s;   // out2 = out + add2 (optional)
  int K, O, act;                     // act: 0 none, 1 relu, 2 sigmoid
  int phase;                         // layers of one phase are independent; a grid barrier separates phases
};
struct GpProgram { GpLayer l[GPP_MAXLAYERS]; int n; unsigned* bar; };   // bar[0] arrivals, bar[1] generation

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// stage the W rows (and / or the x rows) of chunk `ck` of an item into ring slot `slot`
__device__ __forceinline__ void gpp_load(const GpLayer& L, int o0, int ck, float* stage, int B, bool do_w, bool do_x) {
  const int tid = threadIdx.x;
  const int k0 = ck * GPP_KC, kt = min(GPP_KC, L.K - k0);
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(stage);
  if (do_w) {
#pragma unroll
    for (int u = 0; u < GPP_ROWS * (GPP_KC / 4) / GPP_THREADS; ++u) {          // 16 rows x 64 pieces / 256 = 4
      const int c = tid + u * GPP_THREADS, row = c >> 6, piece = c & 63, col = piece * 4;
      float* dst = stage + row * GPP_PITCH + col;
      if (o0 + row < L.O && col < kt) cp_async16(base + (uint32_t)(row * GPP_PITCH + col) * 4u, L.W + (size_t)(o0 + row) * L.K + k0 + col);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (do_x) {
    float* sx = stage + GPP_ROWS * GPP_PITCH;
#pragma unroll
    for (int u = 0; u < GP_BT * (GPP_KC / 4) / GPP_THREADS; ++u) {             // 32 rows x 64 pieces / 256 = 8
      const int c = tid + u * GPP_THREADS, b = c >> 6, piece = c & 63, col = piece * 4;
      if (b < B && col < kt) cp_async16(base + (uint32_t)((GPP_ROWS + b) * GPP_PITCH + col) * 4u, L.x + (size_t)b * L.xs + k0 + col);
      else *reinterpret_cast<float4*>(sx + b * GPP_PITCH + col) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

__global__ void __launch_bounds__(GPP_THREADS, 1) k_gp_persistent(GpProgram P, int B) {
  extern __shared__ __align__(16) float s_ring[];            // [GPP_STAGES][16 + 32][GPP_PITCH]
  __shared__ float s_red[8][GPP_ROWS][GP_BT + 1];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g8 = lane >> 2, t4 = lane & 3;
  unsigned my_gen = 0;
  if (tid == 0) my_gen = ld_acquire_u32(P.bar + 1);
  // work list: items (layer, row block) in table order; this CTA takes items cta, cta + gridDim, ... of each phase
  int li = 0;                                                  // first layer of the current phase
  bool prefetched = false;                                     // the W chunks 0..3 of this phase's first item are in flight
  while (li < P.n) {
    const int phase = P.l[li].phase;
    int lend = li;
    while (lend < P.n && P.l[lend].phase == phase) ++lend;
    // ---- this CTA's items of the phase -------------------------------------------------------------
    int item = blockIdx.x, first = 1;
    for (int l = li; l < lend; ++l) {
      const GpLayer& L = P.l[l];
      const int nblk = (L.O + GPP_ROWS - 1) / GPP_ROWS;
      for (; item < nblk; item += gridDim.x, first = 0) {
        const int o0 = item * GPP_ROWS;
        const int nck = (L.K + GPP_KC - 1) / GPP_KC;
        // prologue: chunks 0..3.  If the weights were prefetched before the barrier only x is still missing.
        const bool pf = prefetched && first;
        if (!pf) {
          for (int c = 0; c < GPP_STAGES; ++c) {
            if (c < nck) gpp_load(L, o0, c, s_ring + c * GPP_STAGE_FLOATS, B, true, false);
            asm volatile("cp.async.commit_group;" ::: "memory");
          }
        }
        for (int c = 0; c < GPP_STAGES; ++c) {
          if (c < nck) gpp_load(L, o0, c, s_ring + c * GPP_STAGE_FLOATS, B, false, true);
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        float acc[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[n][i] = 0.0f;
        for (int c = 0; c < nck; ++c) {
          asm volatile("cp.async.wait_group 3;" ::: "memory");      // group of chunk c (and everything older) has landed
          __syncthreads();
          const float* st = s_ring + (c % GPP_STAGES) * GPP_STAGE_FLOATS;
          const float* wa = st + g8 * GPP_PITCH + t4;
          const float* xb = st + (GPP_ROWS + g8) * GPP_PITCH + t4;
          const int kt = min(GPP_KC, L.K - c * GPP_KC);
#pragma unroll
          for (int ks = 0; ks < GPP_KC / 64; ++ks) {                // warp w takes k-steps w, w+8, w+16, w+24 of the chunk
            const int k = (wid + 8 * ks) * 8;
            if (k < kt) {
              const uint32_t a0 = __float_as_uint(wa[k]), a1 = __float_as_uint(wa[8 * GPP_PITCH + k]);
              const uint32_t a2 = __float_as_uint(wa[k + 4]), a3 = __float_as_uint(wa[8 * GPP_PITCH + k + 4]);
#pragma unroll
              for (int n = 0; n < 4; ++n) {
                const uint32_t b0 = tf32_bits(xb[n * 8 * GPP_PITCH + k]), b1 = tf32_bits(xb[n * 8 * GPP_PITCH + k + 4]);
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[n][0]), "+f"(acc[n][1]), "+f"(acc[n][2]), "+f"(acc[n][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
              }
            }
          }
          __syncthreads();                                          // every warp is done with this slot
          if (c + GPP_STAGES < nck) gpp_load(L, o0, c + GPP_STAGES, s_ring + (c % GPP_STAGES) * GPP_STAGE_FLOATS, B, true, true);
          asm volatile("cp.async.commit_group;" ::: "memory");      // (possibly empty: keeps the group count uniform)
        }
        // fold the 8 warps' partial [16 x 32] tiles in warp order, then the fused epilogue
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int i = 0; i < 4; ++i) s_red[wid][g8 + (i >= 2 ? 8 : 0)][n * 8 + 2 * t4 + (i & 1)] = acc[n][i];
        __syncthreads();
        for (int e = tid; e < GPP_ROWS * GP_BT; e += GPP_THREADS) {
          const int r = e & (GPP_ROWS - 1), b = e / GPP_ROWS, o = o0 + r;
          if (b < B && o < L.O) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) v += s_red[w][r][b];
            v += L.bias ? L.bias[o] : 0.0f;
            if (L.act == 1) v = fmaxf(v, 0.0f);
            else if (L.act == 2) v = 1.0f / (1.0f + expf(-v));
            if (L.mul) v *= L.mul[(size_t)b * L.ms + o];
            if (L.res) v += L.res[(size_t)b * L.rs + o];
            L.out[(size_t)b * L.os + o] = v;
            if (L.out2) L.out2[(size_t)b * L.o2s + o] = v + L.add2[(size_t)b * L.a2s + o];
          }
        }
        __syncthreads();                                            // s_red and the ring are free again
      }
      item -= nblk;                                                 // continue the strided walk in the next layer of the phase
    }
    // drain the (empty) tail groups so that the next phase starts from a known group count
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    li = lend;
    if (li >= P.n) break;
    // ---- prefetch the next phase's first weight chunks, then the grid barrier -----------------------
    prefetched = false;
    {
      const GpLayer& L = P.l[li];
      const int nblk = (L.O + GPP_ROWS - 1) / GPP_ROWS;
      if ((int)blockIdx.x < nblk) {
        const int nck = (L.K + GPP_KC - 1) / GPP_KC;
        for (int c = 0; c < GPP_STAGES; ++c) {
          if (c < nck) gpp_load(L, blockIdx.x * GPP_ROWS, c, s_ring + c * GPP_STAGE_FLOATS, B, true, false);
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        prefetched = true;
      }
    }
    __threadfence();                                                // this CTA's outputs are visible device-wide
    __syncthreads();
    if (tid == 0) {
      const unsigned arrived = atomicAdd(P.bar, 1u);
      if (arrived == gridDim.x - 1) {
        P.bar[0] = 0;
        __threadfence();
        atomicAdd(P.bar + 1, 1u);
      } else {
        long long t0 = clock64();
        while (ld_acquire_u32(P.bar + 1) == my_gen) {
          if (clock64() - t0 > 4000000000LL) __trap();              // a protocol bug must not hang the GPU
        }
      }
      ++my_gen;
      __threadfence();
    }
    __syncthreads();
  }
}

// PositionalEmbedding (models/utils.py:16-31): fp32 frequencies exp(i * -log(1e4)/(half-1))
__global__ void k_gp_posemb(const float* __restrict__ t, const float* __restrict__ freqs, float* __restrict__ out,
                            int half, float scale) {
  pdl_prologue();
  int b = blockIdx.x, i = threadIdx.x;
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
  unsigned* d_bar = nullptr;     // grid-barrier state of the persistent kernel {arrivals, generation}
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
  LION_TRY(m->dmalloc(&g->d_bar, (size_t)4));
  LION_CHECK_CUDA(cudaMemset(g->d_bar, 0, 4 * sizeof(unsigned)));
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

// one chunk of <= 32 shapes through the round-1 form: two kernels per Linear (fallback when the persistent kernel
// cannot be co-resident, and the A/B reference: LION_GP_PERSISTENT=0)
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
  LION_LAUNCH(c, k_gp_posemb, B, 64, 0, t, g->d_freqs, pe, g->emb / 2, g->scale);
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

// one chunk of <= 32 shapes: positional embedding + ONE cooperative launch of k_gp_persistent
static int global_prior_forward_persistent(Model* m, const float* x, const float* t, const float* clip, float* out, int B, int n_cta) {
  GlobalPriorBlk* g = m->gp;
  Ctx* c = m->ctx;
  const int nf = g->nf, tw = g->clip ? 2 * nf : nf, e4 = g->emb * 4;
  float* pe = c->alloc_n<float>((size_t)B * g->emb);
  float* t0 = c->alloc_n<float>((size_t)B * e4);
  float* tadd = c->alloc_n<float>((size_t)B * nf);
  float* xt = c->alloc_n<float>((size_t)B * tw);       // conv1 input of the next cell: [h + temb | clip-mapped]
  float* h = c->alloc_n<float>((size_t)B * nf);
  float* h2 = c->alloc_n<float>((size_t)B * nf);
  float* a = c->alloc_n<float>((size_t)B * nf);
  float* bb = c->alloc_n<float>((size_t)B * nf);
  float* s0 = c->alloc_n<float>((size_t)B * nf / 8);
  LION_LAUNCH(c, k_gp_posemb, B, 64, 0, t, g->d_freqs, pe, g->emb / 2, g->scale);
  GpProgram P{};
  int n = 0, phase = 0;
  auto add = [&](const GPLin& l, const float* xin, int xs, float* o, int os, int act, const float* mul, int ms, const float* res, int rs,
                 float* out2, int o2s, const float* add2, int a2s) -> int {
    if (n >= GPP_MAXLAYERS) { set_error("global prior: more than %d layers", GPP_MAXLAYERS); return LION_ERR_ARG; }
    if (l.K % 8) { set_error("global prior: K=%d must be a multiple of 8", l.K); return LION_ERR_ARG; }
    GpLayer& L = P.l[n++];
    L.W = l.w; L.bias = l.b; L.x = xin; L.xs = xs; L.out = o; L.os = os; L.mul = mul; L.ms = ms; L.res = res; L.rs = rs;
    L.out2 = out2; L.o2s = o2s; L.add2 = add2; L.a2s = a2s; L.K = l.K; L.O = l.O; L.act = act; L.phase = phase;
    return 0;
  };
  // phase 0: everything that depends on the inputs only
  LION_TRY(add(g->t0, pe, g->emb, t0, e4, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0));
  LION_TRY(add(g->in, x, g->D, h, nf, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0));
  if (g->clip) LION_TRY(add(g->cmap, clip, g->clip_dim, xt + nf, tw, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0));
  // phase 1: temb (two 1x1 convs, no nonlinearity in between, resnet.py:181-184); also xt = h + temb for the first cell
  ++phase;
  LION_TRY(add(g->t1, t0, e4, tadd, nf, 0, nullptr, 0, nullptr, 0, xt, tw, h, nf));
  for (auto& cell : g->cells) {
    // conv1(x + t [| clip]) -> ReLU -> (dropout: identity in eval) -> conv2 -> ReLU -> SE -> + x   (resnet.py:60-90, :29-56)
    ++phase; LION_TRY(add(cell.c1, xt, tw, a, nf, 1, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0));
    ++phase; LION_TRY(add(cell.c2, a, nf, bb, nf, 1, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0));
    ++phase; LION_TRY(add(cell.se0, bb, nf, s0, nf / 8, 1, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0));
    ++phase; LION_TRY(add(cell.se2, s0, nf / 8, h2, nf, 2, bb, nf, h, nf, xt, tw, tadd, nf));     // h' = sigmoid(.)*bb + h;  xt = h' + temb
    float* tmp = h; h = h2; h2 = tmp;
  }
  ++phase;
  LION_TRY(add(g->outl, h, nf, out, g->D, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0));
  P.n = n; P.bar = g->d_bar;
  if (!c->dry) {
    const size_t smem = (size_t)GPP_STAGES * GPP_STAGE_FLOATS * sizeof(float);
    int Bv = B;
    void* args[] = {(void*)&P, (void*)&Bv};
    LION_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)k_gp_persistent, dim3(n_cta), dim3(GPP_THREADS), args, smem, c->stream));
    c->launches++;
  }
  return check_launch(c, "global_prior_forward (persistent)");
}

// can the persistent kernel be co-resident on this device?  (cooperative launch, one 200 KB CTA per SM)
static int gp_persistent_ctas(Ctx* c) {
  static int cached[64];
  static DevOnce once;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (once.need()) {
    cached[dev] = 0;
    const char* e = getenv("LION_GP_PERSISTENT");
    int coop = 0, per_sm = 0;
    const size_t smem = (size_t)GPP_STAGES * GPP_STAGE_FLOATS * sizeof(float);
    if (!(e && atoi(e) == 0) && cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop &&
        cudaFuncSetAttribute(k_gp_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gp_persistent, GPP_THREADS, smem) == cudaSuccess && per_sm >= 1) {
      int n = c->num_sms < 128 ? c->num_sms : 128;       // 2048 outputs / 16 rows = 128 work items per big layer
      cached[dev] = n >= 32 ? n : 0;
    }
    cudaGetLastError();
  }
  return cached[dev];
}

int global_prior_forward(Model* m, const float* x, const float* t, const float* clip, float* out, int B) {
  GlobalPriorBlk* g = m->gp;
  Ctx* c = m->ctx;
  if (g->clip && !clip) { set_error("global prior: this network needs clip_feat"); return LION_ERR_ARG; }
  const int n_cta = gp_persistent_ctas(c);
  // any batch size: chunks of 32 shapes (the reference takes any B, resnet.py:195-218)
  for (int b0 = 0; b0 < B; b0 += GP_MAXB) {
    const int nb = B - b0 < GP_MAXB ? B - b0 : GP_MAXB;
    const size_t mk = c->mark();
    const float* xc = x + (size_t)b0 * g->D;
    const float* cc = clip ? clip + (size_t)b0 * g->clip_dim : nullptr;
    float* oc = out + (size_t)b0 * g->D;
    if (n_cta) LION_TRY(global_prior_forward_persistent(m, xc, t + b0, cc, oc, nb, n_cta));
    else LION_TRY(global_prior_forward_layers(m, xc, t + b0, cc, oc, nb));
    c->release(mk);
  }
  return 0;
}

}  // namespace lion
