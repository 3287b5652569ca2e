// lion_b200 -- fused set-abstraction MLP for sm_100a (PointNetSAModule, models/pvcnn2_ada.py:98-114, :140-164, :354-382):
//   ball-query gather -> 1x1 conv -> AdaGN + Swish -> 1x1 conv -> max over the 32 neighbours
// without ever materialising the [B, C, M, 32] tensors in HBM.  GroupNorm needs whole-tensor statistics before the
// activation that follows it, so the module is two passes over the same gather (the gathered rows come from an L2-resident
// feature tensor: 262 KB per shape at level 0):
//   pass 1: gather -> conv1 (tcgen05, TF32) -> per-channel sum / sum of squares of layer 1.         Writes nothing else.
//   pass 2: gather -> conv1 -> AdaGN-1 + Swish in registers -> shared memory -> conv2 (tcgen05) -> statistics of layer 2
//           + per (centre, channel) min and max over the 32 neighbours (conv_tc.cu: Params::pool_mm explains why the
//           two extremes are enough for max_i swish(affine(x_i))); k_act_pool_minmax finishes the module.
// One CTA = 128 threads = 128 rows = 4 centres x 32 neighbours per tile; warp w owns TMEM lanes 32w..32w+31 = the 32
// neighbours of one centre, so the max-pool is a warp reduction.  No warp specialisation: a CTA is a strictly serial
// gather -> MMA -> drain chain and the overlap comes from 4 CTAs per SM (50 KB of shared memory, 128 TMEM columns each).
// DRAM traffic of SA level 0 at B = 32: ~1.36 GB per step with the unfused kernels, indices + 2 x 17 MB with this one.
#include "common.cuh"
#include "model.cuh"
#include <cstdlib>

namespace lion {
namespace saf {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) __trap();
  }
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
// transposing butterflies over a warp's 32 rows x 16 channels: lane l ends with channel ch16(l) = (l >> 1) & 15
template <int OP>   // 0 sum, 1 min, 2 max
__device__ __forceinline__ float warp_red16(float* v, int lane) {
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const int bit = half * 2;
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float keep = upper ? v[i + half] : v[i];
      float send = upper ? v[i] : v[i + half];
      float got = __shfl_xor_sync(0xffffffffu, send, bit);
      v[i] = OP == 0 ? keep + got : (OP == 1 ? fminf(keep, got) : fmaxf(keep, got));
    }
  }
  float o = __shfl_xor_sync(0xffffffffu, v[0], 1);
  return OP == 0 ? v[0] + o : (OP == 1 ? fminf(v[0], o) : fmaxf(v[0], o));
}

struct Params {
  const float4* feat;      // PF [B][GF][N]
  const float4* points;    // [B][N] xyz
  const float4* centers;   // [B][M]
  const int* nidx;         // [B][M][32]
  const float* w1;         // conv_tc packing [G1P][N1][4], tf32-rounded; groups >= 1+GF are zero
  const float* b1;         // [N1]
  const float* w2;         // [N1/4][N2][4]
  const float* b2;         // [N2]
  const float* scale1;     // [B][N1] folded AdaGN-1 (pass 2)
  const float* shift1;
  double* ssum; double* ssq;   // [B][stat_stride]: layer 1 (pass 1) or layer 2 (pass 2)
  int stat_stride;
  float* pool_mm;          // [B][N2/4][M][2][4]  (pass 2)
  int N, M, tiles_per_cta;
};

// GF feature groups (+1 coordinate group, padded to an even count G1), N1 / N2 output channels of the two layers
template <int GF, int N1, int N2, int PASS>
__global__ void __launch_bounds__(128, 4) k_sa_fused(Params P) {
  constexpr int G1 = (GF + 1 + 1) & ~1;        // groups of the layer-1 operand (K = 4*G1, multiple of 8)
  constexpr int G2 = N1 / 4;
  constexpr int NS = PASS == 1 ? N1 : N2;      // channels whose statistics this pass accumulates
  static_assert(N1 % 16 == 0 && N2 % 16 == 0 && N1 + N2 <= 128 && G2 % 2 == 0, "unsupported fused SA shape");
  extern __shared__ __align__(128) uint8_t smem[];
  float4* sA1 = (float4*)smem;                               // [G1][128]
  float4* sA2 = sA1 + G1 * 128;                              // [G2][128]
  float4* sW1 = sA2 + G2 * 128;                              // [G1][N1]
  float4* sW2 = sW1 + G1 * N1;                               // [G2][N2]
  float* s_b1 = (float*)(sW2 + G2 * N2);                     // [N1]
  float* s_b2 = s_b1 + N1;                                   // [N2]
  float* s_sc = s_b2 + N2;                                   // [N1]
  float* s_sh = s_sc + N1;                                   // [N1]
  float* s_stat = s_sh + N1;                                 // [4 warps][2][NS] partial sums
  uint64_t* bars = (uint64_t*)(s_stat + 8 * (N1 > N2 ? N1 : N2));
  uint32_t* s_tmem = (uint32_t*)(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_per_shape = P.M / 4;
  const int ctas_per_shape = (tiles_per_shape + P.tiles_per_cta - 1) / P.tiles_per_cta;
  const int b = blockIdx.x / ctas_per_shape;
  const int tile_begin = (blockIdx.x % ctas_per_shape) * P.tiles_per_cta;
  const int tile_end = min(tile_begin + P.tiles_per_cta, tiles_per_shape);

  const uint32_t bar1 = smem_u32(bars), bar2 = smem_u32(bars + 1);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights / biases / folded affine of this shape; the zero padding group of A1
  for (int i = tid; i < G1 * N1; i += 128) sW1[i] = ((const float4*)P.w1)[i];
  if (PASS == 2) for (int i = tid; i < G2 * N2; i += 128) sW2[i] = ((const float4*)P.w2)[i];
  if (tid < N1) {
    s_b1[tid] = P.b1 ? P.b1[tid] : 0.0f;
    if (PASS == 2) { s_sc[tid] = P.scale1[(size_t)b * N1 + tid]; s_sh[tid] = P.shift1[(size_t)b * N1 + tid]; }
  }
  if (PASS == 2 && tid < N2) s_b2[tid] = P.b2 ? P.b2[tid] : 0.0f;
  for (int g = GF + 1; g < G1; ++g) sA1[g * 128 + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;
  const uint32_t tmem_lane = tmem + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N1 >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N2 >> 3) << 17) | ((128u >> 4) << 24);

  const float4* feat_b = P.feat + (size_t)b * GF * P.N;
  const float4* pts_b = P.points + (size_t)b * P.N;
  float run_s[NS / 16], run_q[NS / 16];
#pragma unroll
  for (int k = 0; k < NS / 16; ++k) { run_s[k] = 0.0f; run_q[k] = 0.0f; }

  uint32_t phase = 0;
  for (int tile = tile_begin; tile < tile_end; ++tile, phase ^= 1) {
    // ---- gather: thread = row = (centre tile*4 + warp, neighbour lane)
    const int centre = tile * 4 + warp;
    {
      const int k = __ldg(P.nidx + ((size_t)b * P.M + centre) * 32 + lane);
      const float4 c = __ldg(P.centers + (size_t)b * P.M + centre);
      const float4 p = __ldg(pts_b + k);
      float4 f[GF];
#pragma unroll
      for (int g = 0; g < GF; ++g) f[g] = __ldg(feat_b + (size_t)g * P.N + k);
      sA1[tid] = make_float4(p.x - c.x, p.y - c.y, p.z - c.z, 0.0f);
#pragma unroll
      for (int g = 0; g < GF; ++g) sA1[(g + 1) * 128 + tid] = f[g];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t ad = make_desc(smem_u32(sA1), 128 * 16, 128), bd = make_desc(smem_u32(sW1), N1 * 16, 128);
#pragma unroll
      for (int ks = 0; ks < G1 / 2; ++ks)       // K = 8 = two channel groups per MMA
        umma_tf32(tmem, ad + (uint64_t)((ks * 2 * 128 * 16) >> 4), bd + (uint64_t)((ks * 2 * N1 * 16) >> 4), idesc1, ks ? 1u : 0u);
      umma_commit(bar1);
    }
    mbar_wait(bar1, phase);
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- layer 1 out of TMEM
#pragma unroll
    for (int cc = 0; cc < N1 / 16; ++cc) {
      float v[16];
      tmem_ld16(tmem_lane + (uint32_t)(cc * 16), v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += s_b1[cc * 16 + i];
      if (PASS == 1) {
        float sq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) sq[i] = v[i] * v[i];
        run_s[cc] += warp_red16<0>(v, lane);
        run_q[cc] += warp_red16<0>(sq, lane);
      } else {
        // AdaGN-1 + Swish, rounded to TF32 exactly like k_act_rows does for a convolution's input
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) {
          float4 a = make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
          const int c0 = cc * 16 + 4 * g4;
          const float4 s = make_float4(s_sc[c0], s_sc[c0 + 1], s_sc[c0 + 2], s_sc[c0 + 3]);
          const float4 t = make_float4(s_sh[c0], s_sh[c0 + 1], s_sh[c0 + 2], s_sh[c0 + 3]);
          sA2[(cc * 4 + g4) * 128 + tid] = f4_tf32(f4_swish(f4_affine(a, s, t)));
        }
      }
    }
    if (PASS == 2) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t ad = make_desc(smem_u32(sA2), 128 * 16, 128), bd = make_desc(smem_u32(sW2), N2 * 16, 128);
#pragma unroll
        for (int ks = 0; ks < G2 / 2; ++ks)
          umma_tf32(tmem + (uint32_t)N1, ad + (uint64_t)((ks * 2 * 128 * 16) >> 4), bd + (uint64_t)((ks * 2 * N2 * 16) >> 4), idesc2,
                    ks ? 1u : 0u);
        umma_commit(bar2);
      }
      mbar_wait(bar2, phase);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < N2 / 16; ++cc) {
        float v[16], w[16];
        tmem_ld16(tmem_lane + (uint32_t)(N1 + cc * 16), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] += s_b2[cc * 16 + i]; w[i] = v[i]; }
        const float mn = warp_red16<1>(w, lane);
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = v[i];
        const float mx = warp_red16<2>(w, lane);
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = v[i] * v[i];
        run_s[cc] += warp_red16<0>(v, lane);
        run_q[cc] += warp_red16<0>(w, lane);
        if ((lane & 1) == 0) {
          const int c = cc * 16 + ((lane >> 1) & 15);
          float* dst = P.pool_mm + ((((size_t)b * (N2 / 4) + (c >> 2)) * P.M + centre) * 2) * 4 + (c & 3);
          dst[0] = mn; dst[4] = mx;
        }
      }
    }
    // the next tile overwrites A1 (its MMA has completed: bar1) and re-uses both accumulators: order this tile's
    // tcgen05.ld before the next MMAs across the barrier inside the next iteration
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  // ---- statistics: warp partials -> CTA (fixed order: bit-reproducible) -> one fp64 atomic per channel
  if ((lane & 1) == 0) {
#pragma unroll
    for (int k = 0; k < NS / 16; ++k) {
      s_stat[(warp * 2 + 0) * NS + k * 16 + ((lane >> 1) & 15)] = run_s[k];
      s_stat[(warp * 2 + 1) * NS + k * 16 + ((lane >> 1) & 15)] = run_q[k];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < NS) {
    float a = 0.0f, q = 0.0f;
#pragma unroll
    for (int w = 0; w < 4; ++w) { a += s_stat[(w * 2 + 0) * NS + tid]; q += s_stat[(w * 2 + 1) * NS + tid]; }
    atomicAdd(P.ssum + (size_t)b * P.stat_stride + tid, (double)a);
    atomicAdd(P.ssq + (size_t)b * P.stat_stride + tid, (double)q);
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
  }
}

template <int GF, int N1, int N2>
constexpr size_t smem_bytes() {
  constexpr int G1 = (GF + 2) & ~1, G2 = N1 / 4;
  return (size_t)(G1 * 128 + G2 * 128 + G1 * N1 + G2 * N2) * 16 + (size_t)(N1 + N2 + 2 * N1 + 8 * (N1 > N2 ? N1 : N2)) * 4 + 64;
}

}  // namespace saf

// shapes this file is instantiated for: (feature channels, layer-1 width, layer-2 width)
bool sa_fused_usable(const SABlk& s) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("LION_SA_FUSED"); on = e ? atoi(e) : 1; }
  if (!on || s.mlp.conv.size() != 2 || s.k != 32 || s.m % 4) return false;
  const ConvW &c1 = s.mlp.conv[0], &c2 = s.mlp.conv[1];
  if (!c1.tc.w || !c2.tc.w || c1.cout != c1.cout_pad || c2.cout != c2.cout_pad) return false;
  return s.cfeat == 32 && c1.cout == 32 && c2.cout == 64 && c1.tc.n == 32 && c2.tc.n == 64 && c1.tc.ck == 32 && c2.tc.ck == 32;
}

// pass 1 (scale1 == nullptr): layer-1 statistics; pass 2: layer-2 statistics + pooled min / max
int sa_fused_run(Ctx* c, const SABlk& s, const float4* feat, const float4* points, const float4* centers, const int* nidx,
                 const float* scale1, const float* shift1, double* ssum, double* ssq, int stat_stride, float* pool_mm,
                 int B, int N) {
  if (c->dry) return 0;
  saf::Params P{};
  const ConvW &c1 = s.mlp.conv[0], &c2 = s.mlp.conv[1];
  P.feat = feat; P.points = points; P.centers = centers; P.nidx = nidx;
  P.w1 = c1.tc.w; P.b1 = c1.bias; P.w2 = c2.tc.w; P.b2 = c2.bias;
  P.scale1 = scale1; P.shift1 = shift1; P.ssum = ssum; P.ssq = ssq; P.stat_stride = stat_stride; P.pool_mm = pool_mm;
  P.N = N; P.M = s.m;
  // CTAs: ~4 per SM resident; aim at about two waves so that the tail is short and the weight loads amortise
  const int tiles = s.m / 4;
  int tpc = 1;
  while (tpc < 16 && (long long)B * ((tiles + tpc - 1) / tpc) > 8LL * c->num_sms) tpc <<= 1;
  P.tiles_per_cta = tpc;
  const int grid = B * ((tiles + tpc - 1) / tpc);
  constexpr size_t smem = saf::smem_bytes<8, 32, 64>();
  static DevOnce attr_once;
  if (attr_once.need()) {
    LION_CHECK_CUDA(cudaFuncSetAttribute(saf::k_sa_fused<8, 32, 64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LION_CHECK_CUDA(cudaFuncSetAttribute(saf::k_sa_fused<8, 32, 64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (!scale1) saf::k_sa_fused<8, 32, 64, 1><<<grid, 128, smem, c->stream>>>(P);
  else saf::k_sa_fused<8, 32, 64, 2><<<grid, 128, smem, c->stream>>>(P);
  c->launches++;
  return check_launch(c, "sa_fused");
}

}  // namespace lion
