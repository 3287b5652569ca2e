// lion_b200 -- Chamfer nearest-neighbour kernels for the generation metrics that follow sampling
// (SURVEY.md 8f rank 2).
//
// Reference: third_party/ChamferDistancePytorch/chamfer3D/chamfer3D.cu:12-143 (NmDistanceKernel,
// launched twice as <<<dim3(32,16,1),512>>> by chamfer_cuda_forward) and its use in
// utils/evaluation_metrics_fast.py:272-340 (_pairwise_EMD_CD_: one sample cloud expanded against a
// batch of reference clouds, dl.mean(1) + dr.mean(1)).
//
// Semantics kept bit for bit:
//   * d = (x2-x1)^2 + (y2-y1)^2 + (z2-z1)^2 with the contraction nvcc applies to the reference
//     source (t = dy*dy; t = fma(dx,dx,t); t = fma(dz,dz,t) -- read off the reference's SASS, the
//     same pattern as the pvcnn kernels, common.cuh: sqdist_ref);
//   * the first candidate is always taken, later ones only when strictly smaller, also across the
//     reference's 512-point chunks (`result > best`): the lowest index wins exact ties.
// Design: the candidate cloud is staged once in shared memory as SoA (three broadcast LDS per
// candidate), every thread keeps Q query points and their running (best, index) in registers, so
// a candidate costs 3 LDS + Q x 8 FP32/select instructions; the pairwise kernel handles both
// directions of one (sample, reference) pair per CTA and reduces the two means in a fixed order
// (deterministic, no atomics, no [Nr, N, 3] expansion of the sample cloud).
#include "common.cuh"
#include "../../include/lion_b200.h"

namespace lion {

constexpr int CD_THREADS = 128;
constexpr int CD_Q = 4;            // query points per thread
constexpr int CD_CHUNK = 2048;     // candidates staged per pass (24 KB)

// one direction: queries q[b][n][3] against candidates c[b][m][3] -> dist[b][n], idx[b][n]
__global__ void __launch_bounds__(CD_THREADS)
k_chamfer_nn(const float* __restrict__ q, const float* __restrict__ c, float* __restrict__ dist, int* __restrict__ idx,
             int n, int m) {
  pdl_prologue();
  __shared__ float sx[CD_CHUNK], sy[CD_CHUNK], sz[CD_CHUNK];
  const int b = blockIdx.y;
  const float* qb = q + (size_t)b * n * 3;
  const float* cb = c + (size_t)b * m * 3;
  float x1[CD_Q], y1[CD_Q], z1[CD_Q], best[CD_Q];
  int bi[CD_Q];
#pragma unroll
  for (int u = 0; u < CD_Q; ++u) {
    int j = (blockIdx.x * CD_Q + u) * CD_THREADS + threadIdx.x;
    int jj = j < n ? j : 0;
    x1[u] = qb[jj * 3 + 0]; y1[u] = qb[jj * 3 + 1]; z1[u] = qb[jj * 3 + 2];
    best[u] = 0.0f; bi[u] = 0;
  }
  for (int k0 = 0; k0 < m; k0 += CD_CHUNK) {
    const int kn = min(CD_CHUNK, m - k0);
    __syncthreads();
    for (int k = threadIdx.x; k < kn; k += CD_THREADS) {
      sx[k] = cb[(size_t)(k0 + k) * 3 + 0]; sy[k] = cb[(size_t)(k0 + k) * 3 + 1]; sz[k] = cb[(size_t)(k0 + k) * 3 + 2];
    }
    __syncthreads();
    int k = 0;
    if (k0 == 0) {            // the first candidate is taken unconditionally (reference: `k==0 || d<best`)
#pragma unroll
      for (int u = 0; u < CD_Q; ++u) { best[u] = sqdist_ref(sx[0] - x1[u], sy[0] - y1[u], sz[0] - z1[u]); bi[u] = 0; }
      k = 1;
    }
#pragma unroll 4
    for (; k < kn; ++k) {
      const float cx = sx[k], cy = sy[k], cz = sz[k];
#pragma unroll
      for (int u = 0; u < CD_Q; ++u) {
        float d = sqdist_ref(cx - x1[u], cy - y1[u], cz - z1[u]);
        if (d < best[u]) { best[u] = d; bi[u] = k0 + k; }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < CD_Q; ++u) {
    int j = (blockIdx.x * CD_Q + u) * CD_THREADS + threadIdx.x;
    if (j < n) { dist[(size_t)b * n + j] = best[u]; if (idx) idx[(size_t)b * n + j] = bi[u]; }
  }
}

// pairwise Chamfer matrix: out[i][j] = mean_n min_m d(s_i[n], r_j[m]) + mean_m min_n d(r_j[m], s_i[n])
// one CTA per (i, j); both clouds live in shared memory (n, m <= CD_PW_MAX points each)
constexpr int CD_PW_MAX = 2048;
constexpr int CD_PW_THREADS = 256;

__device__ __forceinline__ float cd_direction_sum(const float* qx, const float* qy, const float* qz, int n,
                                                  const float* cx, const float* cy, const float* cz, int m) {
  // every thread owns queries t, t + T, ... (<= CD_PW_MAX / T = 8) and scans all candidates
  constexpr int QP = CD_PW_MAX / CD_PW_THREADS;
  float x1[QP], y1[QP], z1[QP], best[QP];
#pragma unroll
  for (int u = 0; u < QP; ++u) {
    int j = u * CD_PW_THREADS + threadIdx.x;
    int jj = j < n ? j : 0;
    x1[u] = qx[jj]; y1[u] = qy[jj]; z1[u] = qz[jj];
    best[u] = sqdist_ref(cx[0] - x1[u], cy[0] - y1[u], cz[0] - z1[u]);
  }
#pragma unroll 2
  for (int k = 1; k < m; ++k) {
    const float ax = cx[k], ay = cy[k], az = cz[k];
#pragma unroll
    for (int u = 0; u < QP; ++u) {
      float d = sqdist_ref(ax - x1[u], ay - y1[u], az - z1[u]);
      best[u] = d < best[u] ? d : best[u];
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int u = 0; u < QP; ++u) s += (u * CD_PW_THREADS + threadIdx.x < n) ? best[u] : 0.0f;
  return s;
}

__global__ void __launch_bounds__(CD_PW_THREADS)
k_chamfer_pairwise(const float* __restrict__ samples, const float* __restrict__ refs, float* __restrict__ out,
                   int n, int m, int n_ref) {
  pdl_prologue();
  extern __shared__ float sm[];     // s: x,y,z [n] ; r: x,y,z [m] ; reduction scratch
  float* sxp = sm; float* syp = sxp + CD_PW_MAX; float* szp = syp + CD_PW_MAX;
  float* rxp = szp + CD_PW_MAX; float* ryp = rxp + CD_PW_MAX; float* rzp = ryp + CD_PW_MAX;
  float* red = rzp + CD_PW_MAX;     // [2][warps]
  const int i = blockIdx.y, j = blockIdx.x;
  const float* s = samples + (size_t)i * n * 3;
  const float* r = refs + (size_t)j * m * 3;
  for (int k = threadIdx.x; k < n; k += CD_PW_THREADS) { sxp[k] = s[k * 3]; syp[k] = s[k * 3 + 1]; szp[k] = s[k * 3 + 2]; }
  for (int k = threadIdx.x; k < m; k += CD_PW_THREADS) { rxp[k] = r[k * 3]; ryp[k] = r[k * 3 + 1]; rzp[k] = r[k * 3 + 2]; }
  __syncthreads();
  float a = cd_direction_sum(sxp, syp, szp, n, rxp, ryp, rzp, m);     // sample -> reference ("dl")
  float c = cd_direction_sum(rxp, ryp, rzp, m, sxp, syp, szp, n);     // reference -> sample ("dr")
  a = warp_sum(a); c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[warp] = a; red[CD_PW_THREADS / 32 + warp] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.0f, tc = 0.0f;
    for (int w = 0; w < CD_PW_THREADS / 32; ++w) { ta += red[w]; tc += red[CD_PW_THREADS / 32 + w]; }
    out[(size_t)i * n_ref + j] = ta / (float)n + tc / (float)m;
  }
}

}  // namespace lion

using namespace lion;

extern "C" int lion_chamfer_forward(const float* xyz1, const float* xyz2, float* dist1, float* dist2, int* idx1, int* idx2,
                                    int B, int N, int M, void* stream) {
  LION_REQUIRE(xyz1 && xyz2 && dist1 && dist2 && B > 0 && N > 0 && M > 0, "lion_chamfer_forward: bad arguments");
  LION_REQUIRE(B <= 65535, "lion_chamfer_forward: at most 65535 cloud pairs per call (got %d)", B);
  Ctx c;
  c.stream = (cudaStream_t)stream;
  LION_LAUNCH(&c, k_chamfer_nn, dim3(cdiv(N, CD_THREADS * CD_Q), B), CD_THREADS, 0, xyz1, xyz2, dist1, idx1, N, M);
  LION_LAUNCH(&c, k_chamfer_nn, dim3(cdiv(M, CD_THREADS * CD_Q), B), CD_THREADS, 0, xyz2, xyz1, dist2, idx2, M, N);
  return check_launch(&c, "lion_chamfer_forward");
}

extern "C" int lion_chamfer_pairwise(const float* samples, const float* refs, float* out, int n_sample, int n_ref, int N, int M,
                                     void* stream) {
  LION_REQUIRE(samples && refs && out && n_sample > 0 && n_ref > 0 && N > 0 && M > 0, "lion_chamfer_pairwise: bad arguments");
  LION_REQUIRE(N <= CD_PW_MAX && M <= CD_PW_MAX, "lion_chamfer_pairwise: clouds of at most %d points (got %d, %d)", CD_PW_MAX, N, M);
  LION_REQUIRE(n_sample <= 65535, "lion_chamfer_pairwise: at most 65535 sample clouds per call (got %d)", n_sample);
  Ctx c;
  c.stream = (cudaStream_t)stream;
  const size_t smem = (6 * CD_PW_MAX + 2 * (CD_PW_THREADS / 32)) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    LION_CHECK_CUDA(cudaFuncSetAttribute(k_chamfer_pairwise, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  LION_LAUNCH(&c, k_chamfer_pairwise, dim3(n_ref, n_sample), CD_PW_THREADS, smem, samples, refs, out, N, M, n_ref);
  return check_launch(&c, "lion_chamfer_pairwise");
}
