// lion_b200 -- device-side cores of the index-producing operators, shared by the
// reference-layout entry points (point_ops.cu) and the fused network path (unet.cu) so that
// both produce bit-identical indices.  Each core is templated on a coordinate loader
// `load(k, x, y, z)` so it works on [3,N] planes and on packed float4 points alike.
//
// Semantics follow the reference kernels exactly (citations into
// /root/reference/third_party/pvcnn/functional/src/); only the parallel decomposition differs.
#pragma once
#include "common.cuh"

namespace lion {

// ---------------------------------------------------------------------------------------
// trilinear corners (interpolate/trilinear_devox.cu:38-76): floor / frac, the hi-corner
// offset is 0 when frac == 0, corner order 000,001,010,011,100,101,110,111 (x major).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void trilinear_corners(float x, float y, float z, int r, int idx[8], float w[8]) {
  float xl = floorf(x), yl = floorf(y), zl = floorf(z);
  float x1 = x - xl, y1 = y - yl, z1 = z - zl;
  float x0 = 1.0f - x1, y0 = 1.0f - y1, z0 = 1.0f - z1;
  w[0] = x0 * y0 * z0; w[1] = x0 * y0 * z1; w[2] = x0 * y1 * z0; w[3] = x0 * y1 * z1;
  w[4] = x1 * y0 * z0; w[5] = x1 * y0 * z1; w[6] = x1 * y1 * z0; w[7] = x1 * y1 * z1;
  int hx = (x1 > 0.0f) ? r * r : 0;
  int hy = (y1 > 0.0f) ? r : 0;
  int hz = (z1 > 0.0f) ? 1 : 0;
  int i0 = (int)xl * r * r + (int)yl * r + (int)zl;
  idx[0] = i0;           idx[1] = i0 + hz;
  idx[2] = i0 + hy;      idx[3] = i0 + hy + hz;
  idx[4] = i0 + hx;      idx[5] = i0 + hx + hz;
  idx[6] = i0 + hx + hy; idx[7] = i0 + hx + hy + hz;
}

// ---------------------------------------------------------------------------------------
// Voxelization.forward (models/pvcnn2_ada.py:173-188): statistics + normalisation.
// Every fp32 operation is a separate torch op in the reference, so nothing may be contracted
// into an FMA here: explicit _rn intrinsics.
// ---------------------------------------------------------------------------------------
constexpr int VOX_THREADS = 256;

// --- coords.mean(2) with the BITS of torch's CUDA reduction ---------------------------------------------
// The voxel index is round(f(coords - mean)): a 1-ulp difference in the mean moves points that sit on a
// rounding boundary into the neighbouring voxel, so "bit-exact voxel index assignment" needs the reference's
// summation ORDER, which is torch's at::native::reduce_kernel<512, 1, ReduceOp<float, MeanOps>> (ATen/native/
// cuda/Reduce.cuh, vt0 = 4; restated in oracle/point_ops.py::cuda_mean_lastdim and pinned against torch on the
// GPU by tests/test_point_ops_gpu.py).  For a contiguous [n_out = 3B, N] fp32 input reduced over N:
//   vectorised by 4 when N >= 128 (dim0 = N/4, else N);  W0 = min(last_pow2(dim0), 32),
//   H = min(last_pow2(n_out), 512 / W0),  W = min(last_pow2(dim0), 512 / H) lanes share one row;
//   lane x keeps 4 accumulators (vectorised: accumulator j <- elements 4*(x + k*W) + j; scalar: accumulator i <-
//   element x + (4k + i)*W), folds them ((a0 + a1) + a2) + a3; lanes fold through a shared-memory tree
//   (offsets W/2 .. 32) and a shuffle-down tree (16 .. 1); result * (float(n_out) / float(n_out * N)).
//   Rows that do not start on a 16-byte boundary (N % 4 != 0) take Reduce.cuh's head / tail path.
// Yes: the bits depend on the batch size.  Called by all threads of the block (blockDim.x >= 256);
// elem(a, i) returns axis a of point i;  row0 = index of this shape's first row (3 * b).
__device__ __forceinline__ int vox_last_pow2(int n) { return 1 << (31 - __clz(n)); }

template <typename Elem>
__device__ __forceinline__ void vox_mean_torch_cuda(Elem elem, int N, int n_out, long long row0, float* s_mean /*[3]*/) {
  __shared__ float s_tree[3][256];
  const bool vec = N >= 128;
  const int dim0 = vec ? N / 4 : N;
  const int d0p = dim0 < 512 ? vox_last_pow2(dim0) : 512;
  const int d1p = n_out < 512 ? vox_last_pow2(n_out) : 512;
  int W = min(d0p, 32);
  const int H = min(d1p, 512 / W);
  W = min(d0p, 512 / H);
  const int x = threadIdx.x;
  float v[3] = {0.f, 0.f, 0.f};
  if (x < W) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (vec) {
        int shift = (int)(((row0 + a) * (long long)N) & 3);
        int base = 0, end = N;
        if (shift > 0) {
          if (x >= shift && x < 4) a0 = __fadd_rn(a0, elem(a, x - shift));
          base = 4 - shift;
          end = N + shift - 4;
        }
        for (int idx = x; idx * 4 + 3 < end; idx += W) {
          a0 = __fadd_rn(a0, elem(a, base + 4 * idx));
          a1 = __fadd_rn(a1, elem(a, base + 4 * idx + 1));
          a2 = __fadd_rn(a2, elem(a, base + 4 * idx + 2));
          a3 = __fadd_rn(a3, elem(a, base + 4 * idx + 3));
        }
        const int i = end - (end & 3) + x;
        if (i < end) a0 = __fadd_rn(a0, elem(a, base + i));
      } else {
        int idx = x;
        while (idx + 3 * W < N) {
          a0 = __fadd_rn(a0, elem(a, idx));
          a1 = __fadd_rn(a1, elem(a, idx + W));
          a2 = __fadd_rn(a2, elem(a, idx + 2 * W));
          a3 = __fadd_rn(a3, elem(a, idx + 3 * W));
          idx += 4 * W;
        }
        if (idx < N) { a0 = __fadd_rn(a0, elem(a, idx)); idx += W; }
        if (idx < N) { a1 = __fadd_rn(a1, elem(a, idx)); idx += W; }
        if (idx < N) { a2 = __fadd_rn(a2, elem(a, idx)); idx += W; }
        if (idx < N) { a3 = __fadd_rn(a3, elem(a, idx)); idx += W; }
      }
      v[a] = __fadd_rn(__fadd_rn(__fadd_rn(a0, a1), a2), a3);
    }
  }
  if (W > 32) {
    if (x < W) { s_tree[0][x] = v[0]; s_tree[1][x] = v[1]; s_tree[2][x] = v[2]; }
    for (int off = W >> 1; off >= 32; off >>= 1) {
      __syncthreads();
      if (x < off) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { v[a] = __fadd_rn(v[a], s_tree[a][x + off]); s_tree[a][x] = v[a]; }
      }
    }
  }
  if (x < 32) {
    const int lim = W < 32 ? W : 32;
    for (int off = lim >> 1; off > 0; off >>= 1) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float o = __shfl_down_sync(0xffffffffu, v[a], off);
        if (x + off >= 32) o = 0.f;      // (never reaches lane 0's dependency cone; keeps the values finite)
        v[a] = __fadd_rn(v[a], o);
      }
    }
    if (x == 0) {
      const float factor = __fdiv_rn((float)n_out, (float)((long long)n_out * N));
#pragma unroll
      for (int a = 0; a < 3; ++a) s_mean[a] = __fmul_rn(v[a], factor);
    }
  }
  __syncthreads();
}

// s_out[0..2] = mean over points (torch-CUDA summation order, see above), s_out[3] = max_k ||c_k - mean||_2.
// load(k, x, y, z) returns point k;  MAXW = blockDim.x / 32 upper bound.
template <int MAXW, typename Load>
__device__ __forceinline__ void vox_stats_block(Load load, int N, int n_out, long long row0, float* s_out) {
  __shared__ float s_max[MAXW];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  vox_mean_torch_cuda([&](int a, int i) { float x, y, z; load(i, x, y, z); return a == 0 ? x : (a == 1 ? y : z); }, N, n_out, row0, s_out);
  float mx = s_out[0], my = s_out[1], mz = s_out[2];
  float best = 0.0f;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    float x, y, z;
    load(k, x, y, z);
    float dx = __fsub_rn(x, mx), dy = __fsub_rn(y, my), dz = __fsub_rn(z, mz);
    float n2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    best = fmaxf(best, __fsqrt_rn(n2));
  }
  best = warp_max(best);
  if (lane == 0) s_max[wid] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmaxf(t, s_max[w]);
    s_out[3] = t;
  }
  __syncthreads();
}

// (dx,dy,dz) already centred.  v in voxel units, clamped to [0, r-1] (not yet rounded).
__device__ __forceinline__ void vox_normalize(float dx, float dy, float dz, float nrm, int r, int normalize,
                                              float eps, float v[3]) {
  float d[3] = {dx, dy, dz};
  float rf = (float)r;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float t;
    if (normalize) {
      float den = __fadd_rn(__fmul_rn(nrm, 2.0f), eps);
      t = __fadd_rn(__fdiv_rn(d[a], den), 0.5f);
    } else {
      t = __fdiv_rn(__fadd_rn(d[a], 1.0f), 2.0f);
    }
    t = __fmul_rn(t, rf);
    v[a] = fminf(fmaxf(t, 0.0f), rf - 1.0f);
  }
}

// ---------------------------------------------------------------------------------------
// furthest point sampling (sampling/sampling.cu:86-167).
// One CTA of 128 threads per shape.  The reference always launches VT = 512 threads
// (sampling.cu:171), thread v owning points v, v+VT, v+2VT, ..., and breaks ties by (max distance,
// then lowest thread, then lowest index within the thread).  Here thread t plays the reference threads
// v = t, t+128, t+256, t+384 (A of them) with up to C points each, visited in (v, index) order,
// so the same total order falls out of a strict '>' scan plus an integer key (v << 20 | index).
// Coordinates and running min-distances live in registers (A*C <= 32 points per thread); the
// arg-max is two warp `redux` operations (max of the distance bits, min of the key among the
// maxima) per level and ONE __syncthreads per round -- the reference: 9 barriers plus global-
// memory distance traffic per round; the first version of this kernel (512 threads, 15-shuffle
// butterflies, 16 warps at the barrier) needed 1600 cycles per round, this one about a third.
// ---------------------------------------------------------------------------------------
constexpr int FPS_THREADS = 128;
constexpr int FPS_MAX_N = 4096;          // A = 4 virtual threads x C = 8 points

inline int fps_virtual_threads(int) { return 512; }   // sampling.cu:171: <<<b, 512>>> whatever n is

// Emit(j, k, x, y, z) is called by one thread for every selected point j=0..M-1.
// s_pts: dynamic shared memory, 3*N floats (SoA copy of the coordinates: the winner's position is
// one LDS away, so the scan carries only (key, slot) -- two selects per point, branch-free, which
// lets the 16 independent distance chains of a thread overlap).
// FULL: every (a, c) slot of every thread holds a point (N = 128*A*C); otherwise slots are masked.
template <int A, int C, bool FULL, typename Load, typename Emit>
__device__ __forceinline__ void fps_block_emit(Load load, Emit emit, int N, int M, int VT, float* s_pts) {
  constexpr unsigned ALL = 0xffffffffu;
  __shared__ uint4 s_m[2], s_id[2];        // per warp: max key, id of the warp's winner (double-buffered by round parity)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float* sx = s_pts; float* sy = s_pts + N; float* sz = s_pts + 2 * N;
  for (int k = tid; k < N; k += FPS_THREADS) { float x, y, z; load(k, x, y, z); sx[k] = x; sy[k] = y; sz[k] = z; }
  __syncthreads();
  float px[A * C], py[A * C], pz[A * C], pd[A * C];
  unsigned vmask = 0u;
#pragma unroll
  for (int a = 0; a < A; ++a)
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int v = tid + a * FPS_THREADS, k = v + c * VT, s = a * C + c;
      const bool ok = FULL || (v < VT && k < N);
      px[s] = ok ? sx[k] : 0.0f; py[s] = ok ? sy[k] : 0.0f; pz[s] = ok ? sz[k] : 0.0f;
      pd[s] = 1e38f;                         // sampling.cpp:54
      if (ok) vmask |= 1u << s;
    }
  float lx = sx[0], ly = sy[0], lz = sz[0];  // first pick is point 0 (sampling.cu:104-106)
  if (tid == 0) emit(0, 0, lx, ly, lz);
  for (int j = 1; j < M; ++j) {
    unsigned bm = 0u; int bs = 0;            // threads without points keep key 0 and never win (sampling.cu:117-118)
#pragma unroll
    for (int s = 0; s < A * C; ++s) {
      float d2 = fminf(sqdist_ref(px[s] - lx, py[s] - ly, pz[s] - lz), pd[s]);
      pd[s] = d2;
      unsigned key = __float_as_uint(d2) + 1u;           // d2 >= 0: the bit pattern orders like the value
      if (!FULL) key = ((vmask >> s) & 1u) ? key : 0u;
      const bool better = key > bm;                      // strict: slots are visited in (thread, index) order
      bm = better ? key : bm;
      bs = better ? s : bs;
    }
    const int bv = tid + (bs / C) * FPS_THREADS;
    const unsigned bid = bm ? (((unsigned)bv << 20) | (unsigned)(bv + (bs % C) * VT)) : ALL;
    const unsigned wm = __reduce_max_sync(ALL, bm);
    const unsigned wi = __reduce_min_sync(ALL, bm == wm ? bid : ALL);
    const int buf = j & 1;
    if (lane == 0) { ((unsigned*)&s_m[buf])[wid] = wm; ((unsigned*)&s_id[buf])[wid] = wi; }
    __syncthreads();
    const uint4 gm = s_m[buf], gi = s_id[buf];
    unsigned tm = gm.x, ti = gi.x;
    if (gm.y > tm || (gm.y == tm && gi.y < ti)) { tm = gm.y; ti = gi.y; }
    if (gm.z > tm || (gm.z == tm && gi.z < ti)) { tm = gm.z; ti = gi.z; }
    if (gm.w > tm || (gm.w == tm && gi.w < ti)) { tm = gm.w; ti = gi.w; }
    const int kb = (int)(ti & 0xfffffu);
    lx = sx[kb]; ly = sy[kb]; lz = sz[kb];
    if (tid == 0) emit(j, kb, lx, ly, lz);
  }
}
static_assert(FPS_THREADS == 128, "fps_block_emit packs one uint4 of per-warp results");

// pick the instantiation: A = reference threads per real thread, C = points per reference thread
#define LION_FPS_DISPATCH(N, VT, CALL)                                                   \
  do {                                                                                   \
    const int _a = (N) <= lion::FPS_THREADS ? 1 : ((N) <= 2 * lion::FPS_THREADS ? 2 : 4); /* reference threads that own a point */ \
    const int _c = ((N) + (VT)-1) / (VT);                                                \
    if (_a == 4 && _c > 4) { if ((N) == 4096) { CALL(4, 8, true); } else { CALL(4, 8, false); } }       \
    else if (_a == 4 && _c > 2) { if ((N) == 2048) { CALL(4, 4, true); } else { CALL(4, 4, false); } }  \
    else if (_a == 4 && _c > 1) { if ((N) == 1024) { CALL(4, 2, true); } else { CALL(4, 2, false); } }  \
    else if (_a == 4) { if ((N) == 512) { CALL(4, 1, true); } else { CALL(4, 1, false); } }             \
    else if (_a == 2) { if ((N) == 256) { CALL(2, 1, true); } else { CALL(2, 1, false); } }             \
    else { if ((N) == 128) { CALL(1, 1, true); } else { CALL(1, 1, false); } }                          \
  } while (0)
// dynamic shared memory of the FPS kernels: the SoA coordinate copy (49 KB at N = 4096: opt-in above 48 KB)
inline size_t fps_smem_bytes(int N) { return (size_t)3 * N * sizeof(float); }

// ---------------------------------------------------------------------------------------
// ball query (ball_query/ball_query.cu:19-50): first K point indices in ascending order with
// d^2 < r^2 (strict), padded with the first hit, all zeros when there is no hit.
// One warp per centre: 32 candidates per step, ballot + prefix popcount keeps the order.
// ---------------------------------------------------------------------------------------
template <typename Load>
__device__ __forceinline__ void ball_query_warp(Load load, float cx, float cy, float cz, float r2, int N, int K,
                                                int* out) {
  int lane = threadIdx.x & 31;
  int cnt = 0, first = 0;
  for (int k0 = 0; k0 < N && cnt < K; k0 += 32) {
    int k = k0 + lane;
    bool hit = false;
    if (k < N) {
      float x, y, z;
      load(k, x, y, z);
      hit = sqdist_ref(cx - x, cy - y, cz - z) < r2;
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
      if (cnt == 0) first = k0 + __ffs(m) - 1;
      int pos = cnt + __popc(m & ((1u << lane) - 1u));
      if (hit && pos < K) out[pos] = k;
      cnt += __popc(m);
    }
  }
  if (cnt > K) cnt = K;
  if (lane >= cnt && lane < K) out[lane] = (cnt > 0) ? first : 0;
}

// ---------------------------------------------------------------------------------------
// 3 nearest centres (interpolate/neighbor_interpolate.cu:36-73): strict '<' cascade over
// centres in index order; weights from distances clamped to [1e-10, 1e10].
// ---------------------------------------------------------------------------------------
struct ThreeNN {
  float b0, b1, b2;
  int i0, i1, i2;
  __device__ __forceinline__ void init() {
    b0 = b1 = b2 = __int_as_float(0x7f800000);   // the reference starts at 1e40 (double): > any float
    i0 = i1 = i2 = 0;
  }
  __device__ __forceinline__ void push(float d, int k) {
    if (d < b2) {
      b2 = d; i2 = k;
      if (d < b1) {
        b2 = b1; i2 = i1; b1 = d; i1 = k;
        if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = k; }
      }
    }
  }
  __device__ __forceinline__ void weights(float& w0, float& w1, float& w2) const {
    float c0 = fmaxf(fminf(1e10f, b0), 1e-10f);
    float c1 = fmaxf(fminf(1e10f, b1), 1e-10f);
    float c2 = fmaxf(fminf(1e10f, b2), 1e-10f);
    float d0d1 = __fmul_rn(c0, c1), d0d2 = __fmul_rn(c0, c2), d1d2 = __fmul_rn(c1, c2);
    float inv = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(d0d1, d0d2), d1d2));
    w0 = __fmul_rn(d1d2, inv); w1 = __fmul_rn(d0d2, inv); w2 = __fmul_rn(d0d1, inv);
  }
};

}  // namespace lion
