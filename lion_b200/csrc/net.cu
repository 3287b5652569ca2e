// lion_b200 -- host orchestration of the PVCNN2-AdaGN U-Net and its blocks on the packed
// layouts of packed_kernels.cuh, plus the network/block C ABI.
//
// Reference being replaced (paths relative to /root/reference):
//   models/latent_points_ada.py:117-173      PVCNN2Unet.forward
//   models/pvcnn2_ada.py:235-280             PVConv.forward
//   models/pvcnn2_ada.py:354-382, :98-114    PointNetSAModule.forward, BallQuery.forward
//   models/pvcnn2_ada.py:393-411             PointNetFPModule.forward
//   models/pvcnn2_ada.py:54-71               LinearAttention.forward
//   models/pvcnn2_ada.py:140-164             SharedMLP.forward
//   models/adagn.py:45-65                    AdaGN.forward
// What is restructured relative to the reference (same arithmetic, fewer passes):
//   * the 61 AdaGN style Linears run as ONE kernel per forward (style is step-invariant);
//   * GroupNorm statistics come out of the producing convolution's epilogue; GroupNorm,
//     the style affine and (after the 2nd conv) the SE gate fold into one per-(b,c) affine
//     that the next consumer applies on load (activation pass / devoxelisation);
//   * voxel indices, counts and normalised coordinates are computed once per distinct
//     (coords, resolution) -- 4x per step instead of 14x;
//   * no permutes: the latent [B,N,4] is already a packed-feature tensor.
#include <memory>
#include <cstdlib>
#include <cmath>
#include "common.cuh"
#include "packed_kernels.cuh"
#include "model.cuh"
#include "../../include/lion_b200.h"

namespace lion {

// =====================================================================================
// error state, context
// =====================================================================================
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int ctx_reserve(Ctx* c, size_t bytes) {
  if (bytes <= c->cap) return 0;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(c->stream, &st);
  if (st != cudaStreamCaptureStatusNone) {
    set_error("scratch arena too small (%zu > %zu bytes) during stream capture: run one eager warm-up call first", bytes, c->cap);
    return LION_ERR_STATE;
  }
  LION_CHECK_CUDA(cudaDeviceSynchronize());
  if (c->base) LION_CHECK_CUDA(cudaFree(c->base));
  c->base = nullptr;
  c->cap = 0;
  size_t want = bytes + bytes / 8 + (size_t(1) << 20);
  cudaError_t e = cudaMalloc((void**)&c->base, want);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) for the scratch arena failed: %s", want, cudaGetErrorString(e));
    return LION_ERR_OOM;
  }
  c->cap = want;
  c->generation++;           // every CUDA graph captured on this context so far has the old addresses baked in
  return 0;
}

// persistent zero grid: (re)allocated and zeroed outside graph capture; users keep it all-zero
int ctx_reserve_zgrid(Ctx* c, size_t bytes) {
  if (bytes <= c->zgrid_cap) return 0;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(c->stream, &st);
  if (st != cudaStreamCaptureStatusNone) {
    set_error("zero grid too small during stream capture: run one eager warm-up call first");
    return LION_ERR_STATE;
  }
  LION_CHECK_CUDA(cudaDeviceSynchronize());
  if (c->zgrid) LION_CHECK_CUDA(cudaFree(c->zgrid));
  c->zgrid = nullptr; c->zgrid_cap = 0;
  cudaError_t e = cudaMalloc((void**)&c->zgrid, bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu) for the zero grid failed: %s", bytes, cudaGetErrorString(e)); return LION_ERR_OOM; }
  LION_CHECK_CUDA(cudaMemset(c->zgrid, 0, bytes));
  c->zgrid_cap = bytes;
  c->generation++;
  return 0;
}

// weight packing kernels
__global__ void k_pack_conv_w(const float* __restrict__ w_ref, const int* __restrict__ kmap, float* __restrict__ wt,
                              int ntaps, int cin_ref, int cin_pad, int cout, int cout_pad) {
  pdl_prologue();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)ntaps * cin_pad * cout_pad;
  if (i >= total) return;
  int co = i % cout_pad;
  int ci = (i / cout_pad) % cin_pad;
  int t = i / ((size_t)cout_pad * cin_pad);
  int src = kmap[ci];
  float v = 0.0f;
  if (src >= 0 && co < cout) v = w_ref[((size_t)co * cin_ref + src) * ntaps + t];
  wt[i] = v;
}
// sparse first convolution: the 27 taps side by side, wy[ci][t * cout_pad + co] = wt[t][ci][co] (zero beyond 27 * cout_pad)
__global__ void k_pack_taps_wide(const float* __restrict__ wt, float* __restrict__ wy, int cin_pad, int cout_pad, int ny_pad) {
  pdl_prologue();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)cin_pad * ny_pad) return;
  int n = i % ny_pad, ci = i / ny_pad;
  int t = n / cout_pad, co = n % cout_pad;
  wy[i] = t < 27 ? wt[((size_t)t * cin_pad + ci) * cout_pad + co] : 0.0f;
}
__global__ void k_pad_vec(const float* __restrict__ src, float* __restrict__ dst, int n, int n_pad) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) dst[i] = i < n ? src[i] : 0.0f;
}


// build a conv's packed forms. kmap: packed input channel -> reference input channel (-1 = zero)
int make_conv(Model* m, ConvW& w, const float* w_ref, const float* b_ref, int ntaps, int cin_ref, int cout,
              const std::vector<int>& kmap) {
  w.ntaps = ntaps; w.cin_ref = cin_ref; w.cin_pad = (int)kmap.size(); w.cout = cout;
  w.cout_pad = (cout <= 4) ? 4 : roundup(cout, 8);
  w.w_ref = w_ref; w.b_ref = b_ref;
  if (!w_ref) { set_error("model parameters exhausted while building a convolution"); return LION_ERR_ARG; }
  if (w.cin_pad % 4) { set_error("packed input channels must be a multiple of 4 (got %d)", w.cin_pad); return LION_ERR_ARG; }
  LION_TRY(m->dmalloc(&w.d_kmap, kmap.size()));
  LION_CHECK_CUDA(cudaMemcpy(w.d_kmap, kmap.data(), kmap.size() * sizeof(int), cudaMemcpyHostToDevice));
  size_t nw = (size_t)ntaps * w.cin_pad * w.cout_pad;
  LION_TRY(m->dmalloc(&w.wt, nw));
  m->jobs.push_back({0, w_ref, w.d_kmap, w.wt, ntaps, cin_ref, w.cin_pad, cout, w.cout_pad});
  if (b_ref) {
    LION_TRY(m->dmalloc(&w.bias, (size_t)w.cout_pad));
    m->jobs.push_back({1, b_ref, nullptr, w.bias, cout, w.cout_pad, 0, 0, 0});
  }
  LION_TRY(conv_tc_prepare(m, w));     // optional tensor-core packing (adds its own job)
  return 0;
}
std::vector<int> ident_map(int c) {
  std::vector<int> k(roundup(c, 4), -1);
  for (int i = 0; i < c; ++i) k[i] = i;
  return k;
}

int run_jobs(Model* m) {
  for (auto& j : m->jobs) {
    if (j.type == 0) {
      size_t total = (size_t)j.a * j.c * j.e;
      k_pack_conv_w<<<(unsigned)cdivz(total, 256), 256>>>(j.src, j.kmap, j.dst, j.a, j.b, j.c, j.d, j.e);
    } else if (j.type == 1) {
      k_pad_vec<<<cdiv(j.b, 128), 128>>>(j.src, j.dst, j.a, j.b);
    } else if (j.type == 3) {
      k_pack_taps_wide<<<(unsigned)cdivz((size_t)j.a * j.c, 256), 256>>>(j.src, j.dst, j.a, j.b, j.c);
    } else {
      LION_TRY(conv_tc_pack_job(j));
    }
  }
  LION_CHECK_CUDA(cudaGetLastError());
  LION_CHECK_CUDA(cudaDeviceSynchronize());
  return 0;
}

// plain: nn.GroupNorm(8, C) of the non-Ada blocks (models/pvcnn2.py) = AdaGN whose style Linear is identically
// (factor, bias) = (1, 0): no `emd` parameters are consumed and k_style_linear writes the constants.
int make_adagn(Model* m, AdaGNW& g, Cursor& cur, int C, bool plain = false) {
  g.C = C;
  g.gamma = cur.next(); g.beta = cur.next();
  const float* ew = plain ? nullptr : cur.next(); const float* eb = plain ? nullptr : cur.next();
  if (cur.bad) { set_error("model parameters exhausted while building AdaGN(%d)", C); return LION_ERR_ARG; }
  if (C % 8 || C > 512) { set_error("AdaGN channels must be a multiple of 8 and <= 512 (got %d)", C); return LION_ERR_ARG; }
  g.style_off = m->style_total;
  m->style_layers.push_back({ew, eb, 2 * C, m->style_total});
  m->style_total += 2 * C;
  return 0;
}

// SharedMLP: n x (1x1 conv, AdaGN, Swish); first conv's input mapping is given
int make_shared_mlp(Model* m, SharedMLPBlk& b, Cursor& cur, int cin_ref, const std::vector<int>& kmap0,
                    const std::vector<int>& outs, bool plain = false) {
  int cin = cin_ref;
  b.cin_pad = (int)kmap0.size();
  b.conv.resize(outs.size());
  b.gn.resize(outs.size());
  for (size_t i = 0; i < outs.size(); ++i) {
    const float* w = cur.next(); const float* bi = cur.next();
    if (cur.bad) { set_error("model parameters exhausted in SharedMLP"); return LION_ERR_ARG; }
    LION_TRY(make_conv(m, b.conv[i], w, bi, 1, cin, outs[i], i == 0 ? kmap0 : ident_map(cin)));
    LION_TRY(make_adagn(m, b.gn[i], cur, outs[i], plain));
    cin = outs[i];
  }
  return 0;
}
int make_attn(Model* m, AttnBlk& a, Cursor& cur, int C, int heads) {
  a.C = C; a.heads = heads;
  int hid = heads * 32;
  const float* qw = cur.next(); const float* ow = cur.next(); const float* ob = cur.next();
  if (cur.bad) { set_error("model parameters exhausted in LinearAttention"); return LION_ERR_ARG; }
  LION_TRY(make_conv(m, a.qkv, qw, nullptr, 1, C, 3 * hid, ident_map(C)));
  LION_TRY(make_conv(m, a.out, ow, ob, 1, hid, C, ident_map(hid)));
  return 0;
}
// cin need not be a multiple of 4: the input PF is padded to whole groups and the padding meets zero weights
int make_pvconv(Model* m, PVConvBlk& p, Cursor& cur, int cin, int cout, int r, bool attn, bool plain = false) {
  p.cin = roundup(cin, 4); p.cout = cout; p.r = r; p.has_attn = attn;
  const float* w1 = cur.next(); const float* b1 = cur.next();
  if (cur.bad) { set_error("model parameters exhausted in PVConv"); return LION_ERR_ARG; }
  LION_TRY(make_conv(m, p.c1, w1, b1, 27, cin, cout, ident_map(cin)));
  if ((p.c1.cout_pad == 32 || p.c1.cout_pad == 64) && p.c1.cout == p.c1.cout_pad && p.c1.cin_pad >= 16) {
    // weights of the sparse form of this convolution (pvconv_fwd): one GEMM x[v] -> y[v][27 taps][cout]
    ConvW& y = p.c1y;
    y.ntaps = 1; y.cin_ref = p.c1.cin_pad; y.cin_pad = p.c1.cin_pad;
    y.cout = y.cout_pad = roundup(27 * p.c1.cout_pad, 128);
    LION_TRY(m->dmalloc(&y.wt, (size_t)y.cin_pad * y.cout_pad));
    m->jobs.push_back({3, p.c1.wt, nullptr, y.wt, y.cin_pad, p.c1.cout_pad, y.cout_pad, 0, 0});
    LION_TRY(conv_tc_prepare(m, y));
  }
  LION_TRY(make_adagn(m, p.g1, cur, cout, plain));
  const float* w2 = cur.next(); const float* b2 = cur.next();
  if (cur.bad) { set_error("model parameters exhausted in PVConv"); return LION_ERR_ARG; }
  LION_TRY(make_conv(m, p.c2, w2, b2, 27, cout, cout, ident_map(cout)));
  LION_TRY(make_adagn(m, p.g2, cur, cout, plain));
  p.se1 = cur.next(); p.se2 = cur.next();
  // state_dict order of the reference PVConv: voxel_layers, attn, point_features (pvcnn2_ada.py:227-233)
  if (attn) LION_TRY(make_attn(m, p.attn, cur, cout, 4));
  LION_TRY(make_shared_mlp(m, p.point, cur, cin, ident_map(cin), {cout}, plain));
  if (cur.bad) { set_error("model parameters exhausted in PVConv"); return LION_ERR_ARG; }
  return 0;
}
int make_sa(Model* m, SABlk& s, Cursor& cur, int cfeat, int mcent, float radius, int k, const std::vector<int>& outs, bool plain = false) {
  s.cfeat = cfeat; s.m = mcent; s.radius = radius; s.k = k;
  if (cfeat % 4) { set_error("SA feature channels must be a multiple of 4 (got %d)", cfeat); return LION_ERR_ARG; }
  if (k != 32) { set_error("SA module: num_neighbors must be 32 (got %d)", k); return LION_ERR_ARG; }
  // reference input order: [rel-xyz(3) | features]  (pvcnn2_ada.py:113) -> packed [xyz,0 | features]
  std::vector<int> kmap(4 + cfeat, -1);
  kmap[0] = 0; kmap[1] = 1; kmap[2] = 2;
  for (int i = 0; i < cfeat; ++i) kmap[4 + i] = 3 + i;
  return make_shared_mlp(m, s.mlp, cur, cfeat + 3, kmap, outs, plain);
}
int make_fp(Model* m, FPBlk& f, Cursor& cur, int cc, int cp, const std::vector<int>& outs) {
  f.cc = cc; f.cp = cp;
  if (cc % 4) { set_error("FP interpolated channels must be a multiple of 4 (got %d)", cc); return LION_ERR_ARG; }
  // reference input order: [interpolated(cc) | skip(cp)] (pvcnn2_ada.py:402-406); skip padded to a group boundary
  std::vector<int> kmap(cc + roundup(cp, 4), -1);
  for (int i = 0; i < cc + cp; ++i) kmap[i] = i;
  return make_shared_mlp(m, f.mlp, cur, cc + cp, kmap, outs);
}

// =====================================================================================
// forward-time helpers
// =====================================================================================
struct PF { float4* p = nullptr; int G = 0; int R = 0; };
struct VoxPrep { const float4* c4; int N, r; float4* nc; int* order; int* ppos; int* len; unsigned char* occ; int occ_stride;
                 int* cidx; int* nocc; int* vgrid; };   // compact ids of the occupied voxels (sparse first convolution) or null
struct Fwd {
  Ctx* c; Model* m; int B;
  char* stat_pool = nullptr;     // all GroupNorm statistics of a forward: zeroed by ONE memset
  size_t stat_off = 0, stat_cap = 0;
  float* aff = nullptr;          // [B][style_total] all AdaGN (factor|bias) vectors of this forward
  std::vector<VoxPrep> vox;
};

static PF alloc_pf(Fwd& f, int G, int R) {
  PF t; t.G = G; t.R = R;
  t.p = f.c->alloc_n<float4>((size_t)f.B * G * R);
  return t;
}
// VG with guard rows on both sides (shifted conv reads may touch up to rp*rp+rp+1 rows outside)
static float4* alloc_vg(Fwd& f, int G, int r) {
  int rp = r + 2;
  size_t P = (size_t)rp * rp * rp, guard = (size_t)rp * rp + rp + 8;
  float4* base = f.c->alloc_n<float4>((size_t)f.B * G * P + 2 * guard);
  return base + guard;
}

static int style_affine_all(Fwd& f, const float* style) {
  Model* m = f.m;
  if (m->style_layers.empty()) return 0;
  f.aff = f.c->alloc_n<float>((size_t)f.B * m->style_total);
  LION_LAUNCH(f.c, k_style_linear, dim3((unsigned)m->style_layers.size(), f.B), 256, m->S * sizeof(float),
              m->d_style_layers, style, m->S, f.aff, m->style_total);
  return check_launch(f.c, "style_affine_all");
}

static ConvGeom geom_rows(int R) {
  ConvGeom g{};
  g.ntaps = 1; g.off[0] = 0; g.rp = 0; g.rows = R; g.p_begin = 0; g.p_end = R;
  g.occ = nullptr; g.occ_stride = 0;
  return g;
}
static ConvGeom geom_grid(int r) {
  ConvGeom g{};
  int rp = r + 2;
  g.ntaps = 27; g.rp = rp; g.rows = rp * rp * rp;
  for (int kx = 0; kx < 3; ++kx) for (int ky = 0; ky < 3; ++ky) for (int kz = 0; kz < 3; ++kz)
    g.off[(kx * 3 + ky) * 3 + kz] = (kx - 1) * rp * rp + (ky - 1) * rp + (kz - 1);
  g.p_begin = rp * rp; g.p_end = (rp - 1) * rp * rp;
  g.occ = nullptr; g.occ_stride = 0;
  return g;
}

// out rows in [p_begin,p_end) of every (b, group < Gout_store); statistics optional
static int run_conv(Fwd& f, const ConvW& w, const float4* in, int Gin, float4* out, int Gout_store,
                    double* ssum, double* ssq, const ConvGeom& geo, float* pool_mm = nullptr, float* out_rm = nullptr, int ld_rm = 0) {
  if (Gin * 4 != w.cin_pad) { set_error("conv: input has %d channels, weights expect %d", Gin * 4, w.cin_pad); return LION_ERR_ARG; }
  if (conv_tc_usable(w, geo))
    return conv_tc_run(f.c, w, in, Gin, out, Gout_store, ssum, ssq, geo, f.B, pool_mm, out_rm, ld_rm);
  if (pool_mm || out_rm) { set_error("conv: the pooled epilogue exists on the tensor-core path only"); return LION_ERR_STATE; }
  int span = geo.p_end - geo.p_begin;
  if (w.cout_pad == 4) {
    LION_LAUNCH(f.c, k_conv_simt<4>, dim3(cdiv(span, 128), 1, f.B), 128, geo.ntaps * 16 * sizeof(float),
                in, w.wt, w.bias, out, ssum, ssq, Gin, w.cin_pad, w.cout_pad, Gout_store, geo);
  } else {
    LION_LAUNCH(f.c, k_conv_simt<8>, dim3(cdiv(span, 128), w.cout_pad / 8, f.B), 128, geo.ntaps * 32 * sizeof(float),
                in, w.wt, w.bias, out, ssum, ssq, Gin, w.cin_pad, w.cout_pad, Gout_store, geo);
  }
  return check_launch(f.c, "conv");
}

// AdaGN (+SE) folded into y = scale*x + shift: k_affine_prep materialises the two [B][C] arrays per layer.
static PrepJob prep_job(Fwd& f, const AdaGNW& g, const double* ssum, const double* ssq, int stat_stride, double count,
                        const float* se1, const float* se2, AffSrc& a) {
  a = AffSrc{nullptr, nullptr, ssum, ssq, stat_stride, g.gamma, g.beta, f.aff + g.style_off, f.m->style_total, count};
  float* scale = f.c->alloc_n<float>((size_t)f.B * g.C);
  float* shift = f.c->alloc_n<float>((size_t)f.B * g.C);
  a.scale = scale; a.shift = shift;
  return PrepJob{ssum, ssq, stat_stride, g.gamma, g.beta, f.aff + g.style_off, f.m->style_total, se1, se2, scale, shift, g.C, count};
}
static int run_prep(Fwd& f, const PrepJob& j0, const PrepJob* j1) {
  int C = j0.C, njobs = 1;
  size_t smem = j0.se_w1 ? (j0.C + j0.C / 8) * sizeof(float) : 0;
  if (j1) {
    njobs = 2;
    if (j1->C > C) C = j1->C;
    size_t s1 = j1->se_w1 ? (j1->C + j1->C / 8) * sizeof(float) : 0;
    if (s1 > smem) smem = s1;
  }
  LION_LAUNCH(f.c, k_affine_prep, dim3(f.B, njobs), C, smem, j0, j1 ? *j1 : j0);
  return check_launch(f.c, "affine_prep");
}
static int run_affine(Fwd& f, const AdaGNW& g, const double* ssum, const double* ssq, int stat_stride, double count,
                      const float* se1, const float* se2, AffSrc& a) {
  PrepJob j = prep_job(f, g, ssum, ssq, stat_stride, count, se1, se2, a);
  return run_prep(f, j, nullptr);
}
static int stat_pool_begin(Fwd& f, size_t bytes) {
  f.stat_pool = (char*)f.c->alloc(bytes);
  f.stat_cap = bytes; f.stat_off = 0;
  return memset_async(f.c, f.stat_pool, 0, bytes);
}
static int alloc_stats(Fwd& f, int stride, double** ssum, double** ssq) {
  size_t bytes = sizeof(double) * 2 * f.B * stride;
  double* s;
  if (f.stat_pool && f.stat_off + bytes <= f.stat_cap) {        // pooled: already zero
    s = (double*)(f.stat_pool + f.stat_off);
    f.stat_off += bytes;
    *ssum = s; *ssq = s + (size_t)f.B * stride;
    return 0;
  }
  s = f.c->alloc_n<double>((size_t)2 * f.B * stride);
  *ssum = s; *ssq = s + (size_t)f.B * stride;
  return memset_async(f.c, s, 0, bytes);
}

// convolution + GroupNorm statistics (epilogue) + AdaGN(/SE) fold (k_affine_prep).
// Round 2 tried computing the fold in the LAST CTA of the convolution (arrival ticket, no extra launch): measured
// 1.3-2.0 ms per step SLOWER at B = 32 (profiles/r02_affine_fold_ab.json) -- a single SM folding 2048-4096 (b, c) pairs
// with cold code on the critical path loses to a 32-CTA kernel whose launch latency the graph mostly hides -- so the
// separate launch stays.
static int conv_gn(Fwd& f, const ConvW& w, const float4* in, int Gin, float4* out, int Gout_store, const ConvGeom& geo,
                   const AdaGNW& g, double count, const float* se1, const float* se2, AffSrc& a) {
  double *ssum, *ssq;
  LION_TRY(alloc_stats(f, w.cout_pad, &ssum, &ssq));
  LION_TRY(run_conv(f, w, in, Gin, out, Gout_store, ssum, ssq, geo));
  return run_affine(f, g, ssum, ssq, w.cout_pad, count, se1, se2, a);
}
// same, but the fold is left to the caller (who merges it with another layer's: run_prep)
static int conv_gn_deferred(Fwd& f, const ConvW& w, const float4* in, int Gin, float4* out, int Gout_store, const ConvGeom& geo,
                            const AdaGNW& g, double count, AffSrc& a, PrepJob& job) {
  double *ssum, *ssq;
  LION_TRY(alloc_stats(f, w.cout_pad, &ssum, &ssq));
  LION_TRY(run_conv(f, w, in, Gin, out, Gout_store, ssum, ssq, geo));
  job = prep_job(f, g, ssum, ssq, w.cout_pad, count, nullptr, nullptr, a);
  return 0;
}

// SharedMLP on a PF.  pool: 1, or 32 = max over neighbour rows after the last activation.
// The activated result goes to dst (Gd groups, offset g_off, R/pool rows).
static int shared_mlp_fwd(Fwd& f, const SharedMLPBlk& m, PF in, int pool, float4* dst, int Gd, int g_off) {
  int n = (int)m.conv.size();
  PF cur = in;
  for (int i = 0; i < n; ++i) {
    const ConvW& w = m.conv[i];
    int Gout = w.cout / 4;
    bool last = (i == n - 1);
    AffSrc a;
    if (last && pool == 32 && cur.R % 128 == 0 && w.cout == w.cout_pad && conv_tc_usable(w, geom_rows(cur.R))) {
      // pooled last layer: the convolution's epilogue keeps, per centre and channel, the minimum and the maximum over the
      // 32 neighbours (enough to evaluate max_i swish(affine(x_i)) exactly, see conv_tc.cu: Params::pool_mm); the
      // [B, C, M, 32] tensor is never written
      int Ro = cur.R / 32;
      float* mm = f.c->alloc_n<float>((size_t)f.B * Gout * Ro * 8);
      double *ssum, *ssq;
      LION_TRY(alloc_stats(f, w.cout_pad, &ssum, &ssq));
      LION_TRY(run_conv(f, w, cur.p, cur.G, nullptr, Gout, ssum, ssq, geom_rows(cur.R), mm));
      LION_TRY(run_affine(f, m.gn[i], ssum, ssq, w.cout_pad, (double)cur.R, nullptr, nullptr, a));
      LION_LAUNCH(f.c, k_act_pool_minmax, dim3(cdiv(Ro, 256), Gout, f.B), 256, 0, (const float4*)mm, dst, a, Gout, w.cout, Ro, Gd, g_off);
      LION_TRY(check_launch(f.c, "shared_mlp pooled"));
      continue;
    }
    PF raw = alloc_pf(f, Gout, cur.R);
    LION_TRY(conv_gn(f, w, cur.p, cur.G, raw.p, Gout, geom_rows(cur.R), m.gn[i], (double)cur.R, nullptr, nullptr, a));
    if (last && pool > 1) {
      if (pool != 32 || cur.R % 32) { set_error("shared_mlp: unsupported pooling %d", pool); return LION_ERR_ARG; }
      int Ro = cur.R / 32;
      LION_LAUNCH(f.c, k_act_rows_pool32, dim3(cdiv(Ro, 8 * 4), Gout, f.B), 256, 0, raw.p, dst, a, Gout, w.cout, Ro, Gd, g_off);
    } else {
      PF nxt;
      float4* o; int gd, go;
      if (last) { o = dst; gd = Gd; go = g_off; }
      else { nxt = alloc_pf(f, Gout, cur.R); o = nxt.p; gd = Gout; go = 0; }
      LION_LAUNCH(f.c, k_act_rows<1>, dim3(cdiv(cur.R, 256), Gout, f.B), 256, 0, raw.p, o, a, Gout, w.cout, cur.R, gd, go, last ? 0 : 1);
      cur = nxt;
    }
    LION_TRY(check_launch(f.c, "shared_mlp act"));
  }
  return 0;
}

static int attn_fwd(Fwd& f, const AttnBlk& a, PF x, float4* dst, int Gd, int g_off) {
  int hid = a.heads * 32, N = x.R;
  PF qkv = alloc_pf(f, 3 * hid / 4, N);
  LION_TRY(run_conv(f, a.qkv, x.p, x.G, qkv.p, qkv.G, nullptr, nullptr, geom_rows(N)));
  const int S = cdiv(N, ATTN_CHUNK);
  if (S > 32) { set_error("attention: N=%d too large (max %d points)", N, 32 * ATTN_CHUNK); return LION_ERR_ARG; }
  float* part = f.c->alloc_n<float>((size_t)f.B * a.heads * S * ATTN_PART);
  LION_LAUNCH(f.c, k_attn_ctx, dim3(a.heads, f.B, S), 256, 0, qkv.p, part, a.heads, N);
  PF o = alloc_pf(f, hid / 4, N);
  LION_LAUNCH(f.c, k_attn_apply, dim3(cdiv(N, 128), a.heads, f.B), 128, 0, qkv.p, part, o.p, a.heads, N, S);
  LION_TRY(check_launch(f.c, "attention"));
  if (Gd != a.C / 4 || g_off != 0) { set_error("attention: destination must be a plain PF"); return LION_ERR_ARG; }
  return run_conv(f, a.out, o.p, o.G, dst, a.C / 4, nullptr, nullptr, geom_rows(N));
}

// The first convolution of a PVConv reads a grid with at most N occupied voxels.  When that is a small fraction of r^3
// (<= 25 %) it is cheaper to multiply only the occupied voxels by all 27 taps (one GEMM, 27 * N rows of output instead of r^3 * 27
// taps of dense work) and let every output voxel gather its neighbours' rows (k_sparse_conv_gather): at r = 32, N = 2048
// that is 16x fewer FLOPs and the convolution becomes a ~0.7 GB streaming problem.  LION_SPARSE_CONV1=0 disables it.
static bool sparse_conv1_wanted(int N, int r) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("LION_SPARSE_CONV1"); on = e ? atoi(e) : 1; }
  return on && (long long)N * 4 <= (long long)r * r * r;
}
static int get_vox(Fwd& f, const float4* c4, int N, int r, VoxPrep** out) {
  for (auto& v : f.vox) if (v.c4 == c4 && v.N == N && v.r == r) { *out = &v; return 0; }
  VoxPrep v{c4, N, r, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr};
  if (N > VOXP_MAXN || r > 32) { set_error("voxelisation: N=%d (max %d) or r=%d (max 32) unsupported", N, VOXP_MAXN, r); return LION_ERR_ARG; }
  v.nc = f.c->alloc_n<float4>((size_t)f.B * N);
  v.order = f.c->alloc_n<int>((size_t)f.B * N);
  v.ppos = f.c->alloc_n<int>((size_t)f.B * N);
  v.len = f.c->alloc_n<int>((size_t)f.B * N);
  int P = (r + 2) * (r + 2) * (r + 2);
  v.occ_stride = (P + 63) / 64 + 4;
  v.occ = f.c->alloc_n<unsigned char>((size_t)f.B * v.occ_stride);
  LION_TRY(memset_async(f.c, v.occ, 0, (size_t)f.B * v.occ_stride));
  if (sparse_conv1_wanted(N, r)) {
    v.cidx = f.c->alloc_n<int>((size_t)f.B * N);
    v.nocc = f.c->alloc_n<int>((size_t)f.B);
    v.vgrid = f.c->alloc_n<int>((size_t)f.B * P);
    LION_TRY(memset_async(f.c, v.vgrid, 0xff, (size_t)f.B * P * sizeof(int)));     // -1 = empty voxel
  }
  LION_LAUNCH(f.c, k_vox_prep, f.B, VOXP_THREADS, 0, c4, v.nc, v.order, v.ppos, v.len, v.occ, v.occ_stride, N, r, v.cidx, v.nocc, v.vgrid);
  LION_TRY(check_launch(f.c, "vox_prep"));
  f.vox.push_back(v);
  *out = &f.vox.back();
  return 0;
}

// PVConv: features PF [cin] + coords -> dst PF (cout) at (Gd, g_off)
static int pvconv_fwd(Fwd& f, const PVConvBlk& p, PF feat, const float4* c4, float4* dst, int Gd, int g_off) {
  int N = feat.R, r = p.r, rp = r + 2, P = rp * rp * rp, Gin = p.cin / 4, Gout = p.cout / 4;
  if (feat.G != Gin) { set_error("PVConv: got %d input channels, expected %d", feat.G * 4, p.cin); return LION_ERR_ARG; }
  VoxPrep* vp;
  LION_TRY(get_vox(f, c4, N, r, &vp));
  size_t mk = f.c->mark();
  ConvGeom geo = geom_grid(r);
  static int sparse_minc = -1;
  if (sparse_minc < 0) { const char* e = getenv("LION_SPARSE_MINC"); sparse_minc = e ? atoi(e) : 32; }
  const bool sparse1 = vp->cidx && ygemm_usable(p.c1y) && feat.G == p.c1y.cin_pad / 4 && p.c1.cout_pad >= sparse_minc;
  float4* g_in = nullptr;
  PF xc;
  if (sparse1) {
    // compact list of the occupied voxels' mean features (same values k_scatter would store into the grid)
    xc = alloc_pf(f, Gin, N);
    LION_LAUNCH(f.c, k_scatter_compact, dim3(cdiv(N, 128), Gin, f.B), 128, 0, feat.p, vp->order, vp->cidx, vp->len, vp->nocc, xc.p, Gin, N);
  } else {
    // point -> voxel scatter-mean into the context's persistent all-zero grid (no per-call memset)
    size_t zbytes = sizeof(float4) * ((size_t)f.B * Gin * P + 2 * ((size_t)rp * rp + rp + 8));
    if (zbytes > f.c->zgrid_need) f.c->zgrid_need = zbytes;
    g_in = f.c->dry ? (float4*)(uintptr_t)0x1000 : (float4*)f.c->zgrid + ((size_t)rp * rp + rp + 8);
    LION_LAUNCH(f.c, k_scatter, dim3(cdiv(N, 128), Gin, f.B), 128, 0, feat.p, vp->order, vp->ppos, vp->len, g_in, Gin, N, P);
  }
  stamp(f.c, f.c->stream, " scatter");
  // point branch first: conv1x1 -> stats (its fold shares a launch with conv1's below; the activation is applied inside
  // the devox kernel)
  const ConvW& pw = p.point.conv[0];
  PF rawp = alloc_pf(f, Gout, N);
  AffSrc ap;
  PrepJob jp, j1;
  LION_TRY(conv_gn_deferred(f, pw, feat.p, feat.G, rawp.p, Gout, geom_rows(N), p.point.gn[0], (double)N, ap, jp));
  // conv1 -> (stats) -> AdaGN + Swish
  float4* raw1 = alloc_vg(f, Gout, r);
  AffSrc a1;
  double V = (double)r * r * r;
  if (sparse1) {
    const int ld = 27 * p.c1.cout_pad;
    float* y = f.c->alloc_n<float>((size_t)f.B * N * ld);
    LION_TRY(ygemm_run(f.c, p.c1y, xc.p, y, ld, vp->nocc, f.B, N));
    double *s1, *q1;
    LION_TRY(alloc_stats(f, p.c1.cout_pad, &s1, &q1));
    const int nwarp = 4;
    const size_t smem = (size_t)nwarp * (32 * (p.c1.cout_pad + 2) * sizeof(float) + SPG_LIST * sizeof(unsigned));
    static DevOnce attr_once;
    if (attr_once.need()) {
      LION_CHECK_CUDA(cudaFuncSetAttribute(k_sparse_conv_gather<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      LION_CHECK_CUDA(cudaFuncSetAttribute(k_sparse_conv_gather<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      // level 0 runs next to the side stream: same (maximum) shared-memory carve-out as its kernels, or these blocks
      // cannot become resident on the SMs FPS occupies (see unet_forward)
      LION_CHECK_CUDA(cudaFuncSetAttribute(k_sparse_conv_gather<32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      LION_CHECK_CUDA(cudaFuncSetAttribute(k_sparse_conv_gather<64>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      LION_CHECK_CUDA(cudaFuncSetAttribute(k_act_grid, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      LION_CHECK_CUDA(cudaFuncSetAttribute(k_scatter_compact, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    const dim3 grid(cdiv(r * r * r, 32 * nwarp), f.B);
    if (p.c1.cout_pad == 32)
      LION_LAUNCH(f.c, k_sparse_conv_gather<32>, grid, 32 * nwarp, smem, y, ld, vp->vgrid, p.c1.bias, raw1, s1, q1, p.c1.cout_pad, r, N);
    else
      LION_LAUNCH(f.c, k_sparse_conv_gather<64>, grid, 32 * nwarp, smem, y, ld, vp->vgrid, p.c1.bias, raw1, s1, q1, p.c1.cout_pad, r, N);
    LION_TRY(check_launch(f.c, "sparse conv1"));
    j1 = prep_job(f, p.g1, s1, q1, p.c1.cout_pad, V, nullptr, nullptr, a1);
  } else {
    // (the tensor-core kernel still skips operand slabs whose 64-row occupancy flags are all clear)
    ConvGeom geo1 = geo;
    geo1.occ = vp->occ; geo1.occ_stride = vp->occ_stride;
    LION_TRY(conv_gn_deferred(f, p.c1, g_in, Gin, raw1, Gout, geo1, p.g1, V, a1, j1));
  }
  LION_TRY(run_prep(f, j1, &jp));
  stamp(f.c, f.c->stream, " conv1");
  // AdaGN-1 + Swish as a stand-alone pass over the grid (HBM-bound).  Round 2 tried to fold it into conv2's operand
  // staging ("transform on load"): parity-green but 3.5x slower convolutions, deleted --
  // profiles/r02_xf_transform_on_load_experiment.txt.  Its extra blocks re-zero the scatter grid (was: k_unscatter).
  float4* act1 = alloc_vg(f, Gout, r);
  {
    const int nb_act = cdiv(P, 256 * ACT_U);
    LION_LAUNCH(f.c, k_act_grid, dim3(nb_act + (sparse1 ? 0 : cdiv(N, 256)), Gout, f.B), 256, 0, raw1, act1, a1, Gout, p.cout, rp, P, nb_act,
                vp->ppos, g_in, Gin, N);
  }
  stamp(f.c, f.c->stream, " act1");
  // conv2 -> (stats) -> AdaGN + SE folded into one affine
  float4* raw2 = alloc_vg(f, Gout, r);
  AffSrc a2;
  LION_TRY(conv_gn(f, p.c2, act1, Gout, raw2, Gout, geo, p.g2, V, p.se1, p.se2, a2));
  stamp(f.c, f.c->stream, " conv2");
  // voxel -> point gather (+ point branch)
  if (p.has_attn) {
    PF fused = alloc_pf(f, Gout, N);
    LION_LAUNCH(f.c, k_devox_fuse, dim3(cdiv(N, 128), Gout, f.B), 128, 0, raw2, vp->nc, a2.scale, a2.shift, rawp.p, ap,
                fused.p, Gout, p.cout, N, r, P, Gout, 0);
    LION_TRY(check_launch(f.c, "pvconv"));
    if (Gd != Gout || g_off != 0) {
      PF t = alloc_pf(f, Gout, N);
      LION_TRY(attn_fwd(f, p.attn, fused, t.p, Gout, 0));
      LION_LAUNCH(f.c, k_copy_groups, dim3(cdiv(N, 256), Gout, f.B), 256, 0, t.p, dst, Gout, Gd, g_off, N);
    } else {
      LION_TRY(attn_fwd(f, p.attn, fused, dst, Gd, g_off));
    }
  } else {
    LION_LAUNCH(f.c, k_devox_fuse, dim3(cdiv(N, 128), Gout, f.B), 128, 0, raw2, vp->nc, a2.scale, a2.shift, rawp.p, ap,
                dst, Gout, p.cout, N, r, P, Gd, g_off);
  }
  LION_TRY(check_launch(f.c, "pvconv"));
  f.c->release(mk);   // grids are dead once the output PF is written (stream order keeps this safe)
  return 0;
}

// SA module: (features PF, coords) -> (dst PF with Gd groups at g_off, centres C4)
// pre_fps >= 0: the centres were already sampled on the side stream (event ev[pre_fps])
static int sa_fwd(Fwd& f, const SABlk& s, PF feat, const float4* c4, float4* centers, float4* dst, int Gd, int g_off,
                  int pre_fps = -1, const int* pre_nidx = nullptr) {
  int N = feat.R, M = s.m, U = s.k, Gf = s.cfeat / 4;
  if (feat.G != Gf) { set_error("SA: got %d feature channels, expected %d", feat.G * 4, s.cfeat); return LION_ERR_ARG; }
  if (N > FPS_MAX_N) { set_error("SA: N=%d too large for FPS", N); return LION_ERR_ARG; }
  if (M > N) { set_error("SA: more centres (%d) than points (%d)", M, N); return LION_ERR_ARG; }
  size_t mk = f.c->mark();
  if (pre_fps >= 0) {
    if (!f.c->dry) LION_CHECK_CUDA(cudaStreamWaitEvent(f.c->stream, f.c->ev[pre_fps], 0));
  } else {
    int* fidx = f.c->alloc_n<int>((size_t)f.B * M);
    const int VT = fps_virtual_threads(N);
#define LION_FPS_CALL(A_, C_, F_)                                                                                         \
  do {                                                                                                                    \
    if (fps_smem_bytes(N) > 48 * 1024)                                                                                    \
      LION_CHECK_CUDA(cudaFuncSetAttribute(k_fps_c4<A_, C_, F_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); \
    LION_LAUNCH(f.c, (k_fps_c4<A_, C_, F_>), f.B, FPS_THREADS, fps_smem_bytes(N), c4, fidx, centers, N, M, VT);           \
  } while (0)
    LION_FPS_DISPATCH(N, VT, LION_FPS_CALL);
#undef LION_FPS_CALL
  }
  const int* nidx = pre_nidx;                     // ball query already ran on the side stream (behind event ev[pre_fps])
  if (!nidx) {
    int* ni = f.c->alloc_n<int>((size_t)f.B * M * U);
    float r2 = s.radius * s.radius;
    LION_LAUNCH(f.c, k_ball_query_c4, dim3(cdiv(M * 32, 256), f.B), 256, 0, centers, c4, ni, N, M, r2, U);
    nidx = ni;
  }
  if (sa_fused_usable(s)) {
    // gather -> conv -> AdaGN/Swish -> conv -> max-pool in two fused passes (sa_fused.cu): no [B, C, M, 32] round trips
    const ConvW &c1 = s.mlp.conv[0], &c2 = s.mlp.conv[1];
    double *s1, *q1, *s2, *q2;
    AffSrc a1, a2;
    LION_TRY(alloc_stats(f, c1.cout_pad, &s1, &q1));
    LION_TRY(sa_fused_run(f.c, s, feat.p, c4, centers, nidx, nullptr, nullptr, s1, q1, c1.cout_pad, nullptr, f.B, N));
    LION_TRY(run_affine(f, s.mlp.gn[0], s1, q1, c1.cout_pad, (double)M * U, nullptr, nullptr, a1));
    LION_TRY(alloc_stats(f, c2.cout_pad, &s2, &q2));
    float* mm = f.c->alloc_n<float>((size_t)f.B * (c2.cout / 4) * M * 8);
    LION_TRY(sa_fused_run(f.c, s, feat.p, c4, centers, nidx, a1.scale, a1.shift, s2, q2, c2.cout_pad, mm, f.B, N));
    LION_TRY(run_affine(f, s.mlp.gn[1], s2, q2, c2.cout_pad, (double)M * U, nullptr, nullptr, a2));
    LION_LAUNCH(f.c, k_act_pool_minmax, dim3(cdiv(M, 256), c2.cout / 4, f.B), 256, 0, (const float4*)mm, dst, a2, c2.cout / 4, c2.cout, M,
                Gd, g_off);
    LION_TRY(check_launch(f.c, "sa fused"));
    f.c->release(mk);
    return 0;
  }
  PF grp = alloc_pf(f, Gf + 1, M * U);
  LION_LAUNCH(f.c, k_group_gather, dim3(cdiv(M * U, 256), Gf + 1, f.B), 256, 0, feat.p, c4, centers, nidx, grp.p, Gf, N, M, U);
  LION_TRY(check_launch(f.c, "sa grouping"));
  LION_TRY(shared_mlp_fwd(f, s.mlp, grp, 32, dst, Gd, g_off));
  f.c->release(mk);
  return 0;
}

// FP module: interpolate centres' features to the points, concat skip, SharedMLP
static int fp_fwd(Fwd& f, const FPBlk& b, const float4* pts_c4, int N, const float4* ctr_c4, int M, PF cfeat, PF skip,
                  float4* dst, int Gd, int g_off, const int* pre_idx = nullptr, const float* pre_wgt = nullptr, int pre_ev = -1) {
  int Gc = b.cc / 4, Gs = roundup(b.cp, 4) / 4;
  if (cfeat.G != Gc || cfeat.R != M) { set_error("FP: centre features mismatch"); return LION_ERR_ARG; }
  if (Gs && (skip.G != Gs || skip.R != N)) { set_error("FP: skip features mismatch (%d groups, expected %d)", skip.G, Gs); return LION_ERR_ARG; }
  size_t mk = f.c->mark();
  const int* idx = pre_idx;
  const float* wgt = pre_wgt;
  if (idx) {                                      // 3-NN search already ran on the side stream
    if (!f.c->dry) LION_CHECK_CUDA(cudaStreamWaitEvent(f.c->stream, f.c->ev[pre_ev], 0));
  } else {
    int* ii = f.c->alloc_n<int>((size_t)f.B * N * 3);
    float* ww = f.c->alloc_n<float>((size_t)f.B * N * 3);
    LION_LAUNCH(f.c, k_three_nn_c4, dim3(cdiv(N, 128), f.B), 128, 1024 * sizeof(float4), pts_c4, ctr_c4, ii, ww, N, M);
    idx = ii; wgt = ww;
  }
  PF cat = alloc_pf(f, Gc + Gs, N);
  LION_LAUNCH(f.c, k_interp_rows, dim3(cdiv(N, 128), Gc, f.B), 128, 0, cfeat.p, idx, wgt, cat.p, Gc, M, N, Gc + Gs, 0);
  if (Gs) LION_LAUNCH(f.c, k_copy_groups, dim3(cdiv(N, 256), Gs, f.B), 256, 0, skip.p, cat.p, Gs, Gc + Gs, Gc, N);
  LION_TRY(check_launch(f.c, "fp interpolate"));
  LION_TRY(shared_mlp_fwd(f, b.mlp, cat, 1, dst, Gd, g_off));
  f.c->release(mk);
  return 0;
}

// =====================================================================================
// U-Net
// =====================================================================================
static float bits_to_float(int v) { float f; memcpy(&f, &v, 4); return f; }

// desc: [num_classes, embed_dim, extra, input_dim, use_att, clip, clip_dim, S,
//        n_sa, {has_conv, oc, nblk, res, m, radius_bits, k, n_mlp, mlp...}*,
//        n_fp, {n_mlp, mlp..., has_conv, oc, nblk, res}*]
// The level/block structure restates create_pointnet2_sa_components / create_pointnet2_fp_modules
// (models/pvcnn2_ada.py:448-567) including their quirks.
// the set-abstraction half of a PVCNN2 network: restates create_pointnet2_sa_components (models/pvcnn2_ada.py:448-517 and
// the non-Ada twin models/pvcnn2.py:440-509) including their quirk that levels > 0 keep only their first PVConv.
// Reads n_sa records {has_conv, oc, nblk, res, m, radius_bits, k, n_mlp, mlp...} from d at q.
static int build_sa_levels(Model* m, Cursor& cur, const std::vector<int>& d, size_t& q, int n_sa, int E, bool use_att, bool plain,
                           std::vector<std::vector<Block>>& levels, std::vector<int>& sa_in, int& in_ch) {
  auto rd = [&](int& v) { if (q >= d.size()) return false; v = d[q++]; return true; };
  for (int c = 0; c < n_sa; ++c) {
    int has_conv, oc, nblk, res, mc, rbits, kk, nm;
    if (!(rd(has_conv) && rd(oc) && rd(nblk) && rd(res) && rd(mc) && rd(rbits) && rd(kk) && rd(nm))) { set_error("descriptor truncated (sa)"); return LION_ERR_ARG; }
    std::vector<int> mlp(nm);
    for (int i = 0; i < nm; ++i) if (!rd(mlp[i])) { set_error("descriptor truncated (sa mlp)"); return LION_ERR_ARG; }
    std::vector<Block> blocks;
    sa_in.push_back(in_ch);
    int k = 0;
    if (has_conv) {
      for (int p = 0; p < nblk; ++p) {
        bool att = ((c + 1) % 2 == 0) && use_att && p == 0;
        if (c == 0 || k == 0) {
          blocks.emplace_back();
          blocks.back().kind = LION_KIND_PVCONV;
          LION_TRY(make_pvconv(m, blocks.back().pv, cur, (c == 0 || k > 0) ? in_ch : in_ch + E, oc, res, att, plain));
        }
        in_ch = oc;
        k++;
      }
    }
    int cfeat = in_ch + (k == 0 ? E : 0);
    blocks.emplace_back();
    blocks.back().kind = LION_KIND_SA;
    float radius; memcpy(&radius, &rbits, 4);
    LION_TRY(make_sa(m, blocks.back().sa, cur, cfeat, mc, radius, kk, mlp, plain));
    in_ch = mlp.back();
    levels.push_back(std::move(blocks));
  }
  return 0;
}

static int build_unet(Model* m, Cursor& cur) {
  const std::vector<int>& d = m->desc;
  size_t q = 0;
  auto rd = [&](int& v) { if (q >= d.size()) return false; v = d[q++]; return true; };
  m->unet.reset(new UnetBlk());
  UnetBlk& u = *m->unet;
  int n_sa = 0;
  if (!(rd(u.num_classes) && rd(u.embed_dim) && rd(u.extra) && rd(u.input_dim) && rd(u.use_att) && rd(u.clip) &&
        rd(u.clip_dim) && rd(u.S) && rd(n_sa))) { set_error("unet descriptor too short"); return LION_ERR_ARG; }
  m->S = u.S;
  int E = u.embed_dim;
  if (u.input_dim != 3 || u.extra < 0 || u.extra > 1) { set_error("unet: points must be xyz + at most one extra feature channel"); return LION_ERR_ARG; }
  if (E % 4) { set_error("unet: embed_dim must be a multiple of 4"); return LION_ERR_ARG; }
  if (E > 0) {
    u.e0w = cur.next(); u.e0b = cur.next(); u.e2w = cur.next(); u.e2b = cur.next();
    int half = E / 2;
    std::vector<float> fr(half);
    for (int i = 0; i < half; ++i) fr[i] = (float)std::exp((double)i * -(std::log(10000.0) / (half - 1)));
    LION_TRY(m->dmalloc(&u.d_freqs, (size_t)half));
    LION_CHECK_CUDA(cudaMemcpy(u.d_freqs, fr.data(), half * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (u.clip) { u.cfw = cur.next(); u.cfb = cur.next(); u.scw = cur.next(); u.scb = cur.next(); }
  int in_ch = u.extra + u.input_dim;
  std::vector<int> sa_in;
  LION_TRY(build_sa_levels(m, cur, d, q, n_sa, E, u.use_att != 0, false, u.sa, sa_in, in_ch));
  int ch_sa = in_ch;
  sa_in[0] = u.extra + u.input_dim - 3;
  if (u.use_att) LION_TRY(make_attn(m, u.gatt, cur, ch_sa, 8));
  int n_fp = 0;
  if (!rd(n_fp)) { set_error("unet descriptor truncated (fp)"); return LION_ERR_ARG; }
  for (int i = 0; i < n_fp; ++i) {
    int nm;
    if (!rd(nm)) { set_error("unet descriptor truncated (fp)"); return LION_ERR_ARG; }
    std::vector<int> mlp(nm);
    for (int j = 0; j < nm; ++j) if (!rd(mlp[j])) { set_error("unet descriptor truncated (fp mlp)"); return LION_ERR_ARG; }
    int has_conv, oc, nblk, res;
    if (!(rd(has_conv) && rd(oc) && rd(nblk) && rd(res))) { set_error("unet descriptor truncated (fp conv)"); return LION_ERR_ARG; }
    std::vector<Block> blocks;
    blocks.emplace_back();
    blocks.back().kind = LION_KIND_FP;
    LION_TRY(make_fp(m, blocks.back().fp, cur, in_ch + E, sa_in[n_sa - 1 - i], mlp));
    in_ch = mlp.back();
    if (has_conv) {
      for (int p = 0; p < nblk; ++p) {
        blocks.emplace_back();
        blocks.back().kind = LION_KIND_PVCONV;
        LION_TRY(make_pvconv(m, blocks.back().pv, cur, in_ch, oc, res, false));
        in_ch = oc;
      }
    }
    u.fp.push_back(std::move(blocks));
  }
  // classifier: SharedMLP(ch_fp -> 128), Dropout, Conv1d(128 -> num_classes) (latent_points_ada.py:94-99)
  LION_TRY(make_shared_mlp(m, u.cls0, cur, in_ch, ident_map(in_ch), {128}));
  const float* cw = cur.next(); const float* cb = cur.next();
  if (cur.bad) { set_error("unet: parameter list too short (%d given)", cur.n); return LION_ERR_ARG; }
  LION_TRY(make_conv(m, u.cls2, cw, cb, 1, 128, u.num_classes, ident_map(128)));
  if (cur.i != cur.n) { set_error("unet: %d parameters given, %d consumed", cur.n, cur.i); return LION_ERR_ARG; }
  return 0;
}

__global__ void k_extract_extra(const float4* __restrict__ x, float4* __restrict__ o, int total) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) o[i] = make_float4(x[i].w, 0.f, 0.f, 0.f);
}

// descriptor [input_dim, zdim, use_att, n_sa, levels...]; parameters: the SA levels in state_dict order, then mlp.weight/bias
static int build_style_enc(Model* m, Cursor& cur) {
  const std::vector<int>& d = m->desc;
  if (d.size() < 4) { set_error("style encoder descriptor: [input_dim, zdim, use_att, n_sa, levels...]"); return LION_ERR_ARG; }
  m->senc.reset(new StyleEncBlk());
  StyleEncBlk& e = *m->senc;
  e.input_dim = d[0]; e.zdim = d[1];
  if (e.input_dim != 3) { set_error("style encoder: input_dim must be 3"); return LION_ERR_ARG; }
  size_t q = 4;
  int in_ch = e.input_dim;
  std::vector<int> sa_in;
  LION_TRY(build_sa_levels(m, cur, d, q, d[3], 0, d[2] != 0, true, e.sa, sa_in, in_ch));
  e.cfeat = in_ch;
  e.mlp_w = cur.next(); e.mlp_b = cur.next();
  if (cur.bad || cur.i != cur.n) { set_error("style encoder: %d parameters given, %d consumed", cur.n, cur.i); return LION_ERR_ARG; }
  if (e.cfeat % 4) { set_error("style encoder: feature width must be a multiple of 4"); return LION_ERR_ARG; }
  m->S = 4;   // no style input: every normalisation is a plain GroupNorm
  return 0;
}

static int style_enc_forward(Fwd& f, const float* x, float* out, int N) {
  StyleEncBlk& e = *f.m->senc;
  Ctx* c = f.c;
  int B = f.B;
  float4* c0 = c->alloc_n<float4>((size_t)B * N);
  LION_LAUNCH(c, k_pad3, cdiv(B * N, 256), 256, 0, x, c0, B * N);
  float* dummy_style = c->alloc_n<float>((size_t)B * 4);      // never read: all style layers are constant
  LION_TRY(style_affine_all(f, dummy_style));
  LION_TRY(stat_pool_begin(f, (size_t)f.m->style_total * B * sizeof(double) + 4096));
  PF feat; feat.p = c0; feat.G = 1; feat.R = N;
  const float4* coords = c0;
  int Ncur = N;
  for (auto& lvl : e.sa) {
    for (auto& blk : lvl) {
      if (blk.kind == LION_KIND_PVCONV) {
        PF o = alloc_pf(f, blk.pv.cout / 4, Ncur);
        LION_TRY(pvconv_fwd(f, blk.pv, feat, coords, o.p, o.G, 0));
        feat = o;
      } else {
        PF o = alloc_pf(f, blk.sa.mlp.cout() / 4, blk.sa.m);
        float4* ctr = c->alloc_n<float4>((size_t)B * blk.sa.m);
        LION_TRY(sa_fwd(f, blk.sa, feat, coords, ctr, o.p, o.G, 0));
        feat = o; coords = ctr; Ncur = blk.sa.m;
      }
    }
  }
  float* pooled = c->alloc_n<float>((size_t)B * e.cfeat);
  LION_LAUNCH(c, k_max_rows, dim3(feat.G, B), 256, 0, feat.p, pooled, feat.G, Ncur);
  LION_LAUNCH(c, k_small_linear, B, 128, e.cfeat * sizeof(float), e.mlp_w, e.mlp_b, pooled, e.cfeat, out, 2 * e.zdim, e.cfeat, 2 * e.zdim, 0);
  return check_launch(c, "style encoder");
}

// style -> (CLIP mixing, latent_points_ada.py:132-137) -> all 61 AdaGN style Linears in one launch; result in f.aff
static int unet_style_affine(Fwd& f, const float* style, const float* clip) {
  UnetBlk& u = *f.m->unet;
  Ctx* c = f.c;
  int B = f.B, E = u.embed_dim;
  if (u.clip) {
    if (!clip) { set_error("unet: this network needs clip_feat"); return LION_ERR_ARG; }
    float* cat = c->alloc_n<float>((size_t)B * (u.S + E));
    float* st2 = c->alloc_n<float>((size_t)B * u.S);
    if (!c->dry) LION_CHECK_CUDA(cudaMemcpy2DAsync(cat, (u.S + E) * sizeof(float), style, u.S * sizeof(float), u.S * sizeof(float), B, cudaMemcpyDeviceToDevice, c->stream));
    LION_LAUNCH(c, k_small_linear, B, 128, u.clip_dim * sizeof(float), u.cfw, u.cfb, clip, u.clip_dim, cat + u.S, u.S + E, u.clip_dim, E, 0);
    LION_LAUNCH(c, k_small_linear, B, 128, (u.S + E) * sizeof(float), u.scw, u.scb, cat, u.S + E, st2, u.S, u.S + E, u.S, 0);
    style = st2;
  }
  LION_TRY(check_launch(c, "unet style"));
  return style_affine_all(f, style);
}

static int unet_forward(Fwd& f, const float* x, const float* t, const float* style, const float* clip, float* out, int N) {
  UnetBlk& u = *f.m->unet;
  int B = f.B, E = u.embed_dim;
  Ctx* c = f.c;
  // time embedding: sinusoid -> Linear -> LeakyReLU(0.1) -> Linear  (latent_points_ada.py:53-57, :101-128)
  float* temb = nullptr;
  float *temb_sinu = nullptr, *temb_h = nullptr;
  bool temb_pending = false;
  if (E > 0) {
    if (!t) { set_error("unet: this network needs timesteps"); return LION_ERR_ARG; }
    float* sinu = c->alloc_n<float>((size_t)B * E);
    float* h = c->alloc_n<float>((size_t)B * E);
    temb = c->alloc_n<float>((size_t)B * E);
    temb_sinu = sinu; temb_h = h;       // launched below, on the side stream: first used at SA level 1
  }
  // AdaGN style Linears (and the CLIP mixing in front of them) depend on the style only, which is constant over the
  // 1000 steps of a sampling run: style == nullptr means "use what lion_unet_cache_style computed"
  if (style) {
    LION_TRY(unet_style_affine(f, style, clip));
  } else {
    if (!f.m->aff_cache || f.m->aff_cache_B != B) { set_error("unet: no cached style for B=%d (call lion_unet_cache_style first)", B); return LION_ERR_STATE; }
    f.aff = f.m->aff_cache;
  }
  LION_TRY(check_launch(c, "unet prologue"));
  LION_TRY(stat_pool_begin(f, (size_t)f.m->style_total * f.B * sizeof(double) + 4096));   // sum(2*C) doubles per shape

  int n_sa = (int)u.sa.size();
  std::vector<const float4*> coords_list(n_sa);
  std::vector<PF> feats_list(n_sa);
  std::vector<int> n_list(n_sa);
  // level-0 inputs: coords = xyz, features = all input channels (the latent x[B,N,4] itself is a PF with G=1; a
  // 3-channel cloud x[B,N,3] (PointTransPVC) is padded to (x, y, z, 0), which serves as coordinates AND features)
  float4* c0 = c->alloc_n<float4>((size_t)B * N);
  PF feat; feat.G = 1; feat.R = N;
  if (u.extra == 0) {
    LION_LAUNCH(c, k_pad3, cdiv(B * N, 256), 256, 0, x, c0, B * N);
    feat.p = c0;
  } else {
    LION_LAUNCH(c, k_make_coords, cdiv(B * N, 256), 256, 0, (const float4*)x, c0, B * N);
    feat.p = (float4*)x;
  }
  // furthest-point sampling of all levels depends on coordinates only: a chain of 1360 latency-
  // bound rounds on 32 SMs.  Fork it onto the side stream so it hides under the first PVConvs.
  std::vector<float4*> fps_centers(n_sa, nullptr);
  // Neighbour searches depend on coordinates only as well: the ball query of level i follows FPS i on the side stream,
  // the four 3-NN searches of the FP half follow the last FPS (LION_AUX_NN=0: keep them on the main stream).
  static int aux_nn = -1;
  if (aux_nn < 0) { const char* e = getenv("LION_AUX_NN"); aux_nn = e ? atoi(e) : 1; }
  const bool side_nn = aux_nn != 0 && n_sa <= 4;
  std::vector<int*> sa_nidx(n_sa, nullptr), fp_idx(n_sa, nullptr);
  std::vector<char> vox_pending(n_sa + 1, 0);
  std::vector<float*> fp_wgt(n_sa, nullptr);
  {
    const float4* src = c0;
    int ncur = N;
    if (!c->dry) {
      LION_CHECK_CUDA(cudaEventRecord(c->ev_fork, c->stream));
      LION_CHECK_CUDA(cudaStreamWaitEvent(c->aux, c->ev_fork, 0));
      static DevOnce carve_once;
      if (carve_once.need()) {     // side-stream kernels share SMs with the convolutions: same (maximum) carve-out
        LION_CHECK_CUDA(cudaFuncSetAttribute(k_ball_query_c4, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        LION_CHECK_CUDA(cudaFuncSetAttribute(k_three_nn_c4, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      }
    }
    for (int i = 0; i < n_sa && i < 8; ++i) {
      const SABlk& sb = u.sa[i].back().sa;
      if (ncur > FPS_MAX_N || sb.m > ncur) { set_error("unet: FPS sizes unsupported"); return LION_ERR_ARG; }
      fps_centers[i] = c->alloc_n<float4>((size_t)B * sb.m);
      int* fidx = c->alloc_n<int>((size_t)B * sb.m);
      if (!c->dry) {
        const int VT = fps_virtual_threads(ncur);
#define LION_FPS_CALL(A_, C_, F_)                                                                                         \
  do {                                                                                                                    \
    if (fps_smem_bytes(ncur) > 48 * 1024)                                                                                 \
      LION_CHECK_CUDA(cudaFuncSetAttribute(k_fps_c4<A_, C_, F_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); \
    LION_CHECK_CUDA(cudaFuncSetAttribute(k_fps_c4<A_, C_, F_>, cudaFuncAttributePreferredSharedMemoryCarveout,            \
                                         cudaSharedmemCarveoutMaxShared));                                               \
    k_fps_c4<A_, C_, F_><<<B, FPS_THREADS, fps_smem_bytes(ncur), c->aux>>>(src, fidx, fps_centers[i], ncur, sb.m, VT);    \
  } while (0)
        LION_FPS_DISPATCH(ncur, VT, LION_FPS_CALL);
#undef LION_FPS_CALL
        c->launches++;
        stamp(c, c->aux, "aux:fps", i);
      }
      if (side_nn) {
        sa_nidx[i] = c->alloc_n<int>((size_t)B * sb.m * sb.k);
        if (!c->dry) {
          k_ball_query_c4<<<dim3(cdiv(sb.m * 32, 256), B), 256, 0, c->aux>>>(fps_centers[i], src, sa_nidx[i], ncur, sb.m,
                                                                             sb.radius * sb.radius, sb.k);
          c->launches++;
          stamp(c, c->aux, "aux:ballq", i);
        }
      }
      if (!c->dry) LION_CHECK_CUDA(cudaEventRecord(c->ev[i], c->aux));
      if (i == 0 && temb && !c->dry) {
        // three tiny dependent launches (~45 us of latency) that nothing needs before level 1
        k_time_sinusoid<<<B, 64, 0, c->aux>>>(t, u.d_freqs, temb_sinu, E / 2, 1.0f);
        k_small_linear<<<B, 128, E * sizeof(float), c->aux>>>(u.e0w, u.e0b, temb_sinu, E, temb_h, E, E, E, 1);
        k_small_linear<<<B, 128, E * sizeof(float), c->aux>>>(u.e2w, u.e2b, temb_h, E, temb, E, E, E, 0);
        c->launches += 3;
        stamp(c, c->aux, "aux:temb");
        LION_CHECK_CUDA(cudaEventRecord(c->ev_temb, c->aux));
        temb_pending = true;
      }
      // voxelisation prep of the PVConvs that work on these centres (the next SA level and, mirrored, an FP level):
      // it depends on coordinates only -- one CTA per shape, 13-25 us each on the critical path otherwise
      if (side_nn) {
        int rs[2] = {0, 0};
        if (i + 1 < n_sa) for (auto& blk : u.sa[i + 1]) if (blk.kind == LION_KIND_PVCONV) rs[0] = blk.pv.r;
        const int fi = n_sa - 2 - i;                  // FP stage whose PVConvs run on level i + 1's points
        if (fi >= 0 && fi < (int)u.fp.size()) for (auto& blk : u.fp[fi]) if (blk.kind == LION_KIND_PVCONV) rs[1] = blk.pv.r;
        bool any = false;
        for (int k = 0; k < 2; ++k) {
          if (rs[k] <= 0 || (k == 1 && rs[1] == rs[0])) continue;
          cudaStream_t main_stream = c->stream;
          c->stream = c->aux;
          VoxPrep* vp = nullptr;
          int rc = get_vox(f, fps_centers[i], sb.m, rs[k], &vp);
          c->stream = main_stream;
          if (rc) return rc;
          any = true;
        }
        if (any && !c->dry) {
          stamp(c, c->aux, "aux:voxprep", i + 1);
          LION_CHECK_CUDA(cudaEventRecord(c->ev_vox[i], c->aux));
          vox_pending[i + 1] = true;
        }
      }
      src = fps_centers[i];
      ncur = sb.m;
    }
    if (side_nn && u.fp.size() == (size_t)n_sa) {
      for (int lvl = n_sa - 1; lvl >= 0; --lvl) {
        const float4* pts = lvl == 0 ? c0 : fps_centers[lvl - 1];
        const int npts = lvl == 0 ? N : u.sa[lvl - 1].back().sa.m, nctr = u.sa[lvl].back().sa.m;
        fp_idx[lvl] = c->alloc_n<int>((size_t)B * npts * 3);
        fp_wgt[lvl] = c->alloc_n<float>((size_t)B * npts * 3);
        if (!c->dry) {
          k_three_nn_c4<<<dim3(cdiv(npts, 128), B), 128, 1024 * sizeof(float4), c->aux>>>(pts, fps_centers[lvl], fp_idx[lvl],
                                                                                        fp_wgt[lvl], npts, nctr);
          c->launches++;
          stamp(c, c->aux, "aux:3nn", lvl);
          LION_CHECK_CUDA(cudaEventRecord(c->ev[4 + lvl], c->aux));
        }
      }
    }
  }
  stamp(c, c->stream, "start");
  const float4* coords = c0;
  int Ncur = N;
  bool has_t = temb != nullptr;
  auto with_temb = [&](PF src, PF* dstp) -> int {   // cat(features, temb expanded) (:145)
    if (temb_pending) { LION_CHECK_CUDA(cudaStreamWaitEvent(c->stream, c->ev_temb, 0)); temb_pending = false; }
    PF d = alloc_pf(f, src.G + E / 4, src.R);
    LION_LAUNCH(c, k_copy_groups, dim3(cdiv(src.R, 256), src.G, B), 256, 0, src.p, d.p, src.G, d.G, 0, src.R);
    LION_LAUNCH(c, k_fill_groups, dim3(cdiv(src.R, 256), E / 4, B), 256, 0, temb, E, d.p, d.G, src.G, src.R);
    *dstp = d;
    return check_launch(c, "concat temb");
  };
  static int share_kb = -1;
  if (share_kb < 0) { const char* e = getenv("LION_TC_SHARE_KB"); share_kb = e ? atoi(e) : 196; }
  for (int i = 0; i < n_sa; ++i) {
    c->conv_smem_cap = (i == 0 && share_kb > 0) ? share_kb * 1024 : 0;      // the side stream is busy during level 0
    feats_list[i] = feat; coords_list[i] = coords; n_list[i] = Ncur;
    if (vox_pending[i]) { LION_CHECK_CUDA(cudaStreamWaitEvent(c->stream, c->ev_vox[i - 1], 0)); vox_pending[i] = 0; }
    if (i > 0 && has_t) LION_TRY(with_temb(feat, &feat));
    for (auto& blk : u.sa[i]) {
      if (blk.kind == LION_KIND_PVCONV) {
        PF o = alloc_pf(f, blk.pv.cout / 4, Ncur);
        LION_TRY(pvconv_fwd(f, blk.pv, feat, coords, o.p, o.G, 0));
        feat = o;
        stamp(c, c->stream, "sa.pvconv", i);
      } else {
        PF o = alloc_pf(f, blk.sa.mlp.cout() / 4, blk.sa.m);
        float4* ctr = fps_centers[i];
        LION_TRY(sa_fwd(f, blk.sa, feat, coords, ctr, o.p, o.G, 0, i, sa_nidx[i]));
        feat = o; coords = ctr; Ncur = blk.sa.m;
        stamp(c, c->stream, "sa.module", i);
      }
    }
  }
  c->conv_smem_cap = 0;
  // skip features of level 0 are the extra channels only (inputs[:, 3:], :153): packed as
  // one group [f, 0, 0, 0]
  if (u.extra == 1) {
    PF s0 = alloc_pf(f, 1, N);
    LION_LAUNCH(c, k_extract_extra, cdiv(B * N, 256), 256, 0, (const float4*)x, s0.p, B * N);
    feats_list[0] = s0;
  } else {
    feats_list[0] = PF();          // no skip features at level 0
  }
  if (u.use_att) {
    PF o = alloc_pf(f, feat.G, Ncur);
    LION_TRY(attn_fwd(f, u.gatt, feat, o.p, o.G, 0));
    feat = o;
    stamp(c, c->stream, "global_att");
  }
  for (size_t i = 0; i < u.fp.size(); ++i) {
    int lvl = n_sa - 1 - (int)i;
    for (auto& blk : u.fp[i]) {
      if (blk.kind == LION_KIND_FP) {
        PF cf = feat;
        if (has_t) LION_TRY(with_temb(feat, &cf));          // torch.cat([features, temb]) (:160)
        PF o = alloc_pf(f, blk.fp.mlp.cout() / 4, n_list[lvl]);
        LION_TRY(fp_fwd(f, blk.fp, coords_list[lvl], n_list[lvl], coords, Ncur, cf, feats_list[lvl], o.p, o.G, 0,
                        fp_idx[lvl], fp_wgt[lvl], 4 + lvl));
        feat = o; coords = coords_list[lvl]; Ncur = n_list[lvl];
        stamp(c, c->stream, "fp.module", (int)i);
      } else {
        PF o = alloc_pf(f, blk.pv.cout / 4, Ncur);
        LION_TRY(pvconv_fwd(f, blk.pv, feat, coords, o.p, o.G, 0));
        feat = o;
        stamp(c, c->stream, "fp.pvconv", (int)i);
      }
    }
  }
  PF h = alloc_pf(f, 32, Ncur);
  LION_TRY(shared_mlp_fwd(f, u.cls0, feat, 1, h.p, 32, 0));
  if (u.num_classes == 4) {
    LION_TRY(run_conv(f, u.cls2, h.p, h.G, (float4*)out, 1, nullptr, nullptr, geom_rows(Ncur)));
  } else {
    PF o4 = alloc_pf(f, (u.num_classes + 3) / 4, Ncur);
    LION_TRY(run_conv(f, u.cls2, h.p, h.G, o4.p, o4.G, nullptr, nullptr, geom_rows(Ncur)));
    LION_LAUNCH(c, k_pf_to_pm, dim3(cdiv(Ncur, 256), o4.G, B), 256, 0, o4.p, out, o4.G, u.num_classes, Ncur);
  }
  if (temb_pending) LION_CHECK_CUDA(cudaStreamWaitEvent(c->stream, c->ev_temb, 0));   // never consumed: still join the side stream
  for (int i = 1; i <= n_sa; ++i) if (vox_pending[i]) LION_CHECK_CUDA(cudaStreamWaitEvent(c->stream, c->ev_vox[i - 1], 0));
  stamp(c, c->stream, "end");
  return check_launch(c, "unet epilogue");
}


// run `body` twice: a dry pass measuring the arena, then (after growing it) the real pass
template <typename F>
static int two_pass(Model* m, void* stream, int B, F body) {
  Ctx* c = m->ctx;
  std::unique_lock<std::mutex> lock;
  if (c->mu) lock = std::unique_lock<std::mutex>(*c->mu);
  c->stream = (cudaStream_t)stream;
  for (int pass = 0; pass < 2; ++pass) {
    c->dry = (pass == 0);
    c->reset();
    Fwd f{c, m, B};
    int rc = body(f);
    if (rc) { c->dry = false; return rc; }
    if (pass == 0) {
      LION_TRY(ctx_reserve(c, c->peak));
      LION_TRY(ctx_reserve_zgrid(c, c->zgrid_need));
    }
  }
  return 0;
}

}  // namespace lion

// =====================================================================================
// C ABI
// =====================================================================================
using namespace lion;

struct LionCtx { Ctx c; std::mutex mu; };
struct LionModel { Model m; };

extern "C" int lion_version(void) { return 100; }
extern "C" const char* lion_last_error(void) { return get_error(); }

extern "C" int lion_ctx_create(int device, LionCtx** out) {
  LION_REQUIRE(out, "lion_ctx_create: null out");
  LION_CHECK_CUDA(cudaSetDevice(device));
  LionCtx* h = new LionCtx();
  h->c.device = device;
  h->c.mu = &h->mu;
  cudaDeviceProp prop;
  LION_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  h->c.num_sms = prop.multiProcessorCount;
  { const char* e = getenv("LION_PDL"); h->c.pdl = (e && atoi(e) != 0); }   // measured: no gain inside CUDA graphs; off by default
  LION_CHECK_CUDA(cudaStreamCreateWithFlags(&h->c.aux, cudaStreamNonBlocking));
  { const char* e = getenv("LION_TIMELINE");
    if (e && atoi(e) != 0) LION_CHECK_CUDA(cudaMalloc(&h->c.d_stamps, LION_MAX_STAMPS * sizeof(unsigned long long))); }
  LION_CHECK_CUDA(cudaEventCreateWithFlags(&h->c.ev_fork, cudaEventDisableTiming));
  LION_CHECK_CUDA(cudaEventCreateWithFlags(&h->c.ev_temb, cudaEventDisableTiming));
  for (int i = 0; i < 4; ++i) LION_CHECK_CUDA(cudaEventCreateWithFlags(&h->c.ev_vox[i], cudaEventDisableTiming));
  for (int i = 0; i < 8; ++i) LION_CHECK_CUDA(cudaEventCreateWithFlags(&h->c.ev[i], cudaEventDisableTiming));
  if (prop.major != 10) {
    set_error("lion_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    delete h;
    return LION_ERR_STATE;
  }
  *out = h;
  return 0;
}
extern "C" int lion_ctx_destroy(LionCtx* h) {
  if (!h) return 0;
  if (h->c.base) cudaFree(h->c.base);
  if (h->c.zgrid) cudaFree(h->c.zgrid);
  if (h->c.aux) cudaStreamDestroy(h->c.aux);
  if (h->c.d_stamps) cudaFree(h->c.d_stamps);
  if (h->c.ev_fork) cudaEventDestroy(h->c.ev_fork);
  if (h->c.ev_temb) cudaEventDestroy(h->c.ev_temb);
  for (int i = 0; i < 4; ++i) if (h->c.ev_vox[i]) cudaEventDestroy(h->c.ev_vox[i]);
  for (int i = 0; i < 8; ++i) if (h->c.ev[i]) cudaEventDestroy(h->c.ev[i]);
  delete h;
  return 0;
}
extern "C" int lion_ctx_last_launches(LionCtx* h) { return h ? h->c.launches : 0; }
extern "C" unsigned lion_ctx_generation(LionCtx* h) { return h ? h->c.generation : 0; }
extern "C" int lion_ctx_timeline(LionCtx* h, unsigned long long* t_ns, char* names, int max_entries) {
  LION_REQUIRE(h && t_ns && names && max_entries >= 0, "lion_ctx_timeline: null argument");
  if (!h->c.d_stamps) { set_error("lion_ctx_timeline: the context was created without LION_TIMELINE=1"); return LION_ERR_STATE; }
  int n = h->c.n_stamps < max_entries ? h->c.n_stamps : max_entries;
  LION_CHECK_CUDA(cudaSetDevice(h->c.device));
  LION_CHECK_CUDA(cudaDeviceSynchronize());
  LION_CHECK_CUDA(cudaMemcpy(t_ns, h->c.d_stamps, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) memcpy(names + (size_t)i * 24, h->c.stamp_names[i], 24);
  return n;
}
extern "C" size_t lion_ctx_arena_bytes(LionCtx* h) { return h ? h->c.cap : 0; }
extern "C" size_t lion_workspace_bytes(LionCtx* h) { return h ? h->c.cap + h->c.zgrid_cap : 0; }

extern "C" int lion_model_create(LionCtx* ctx, int kind, const int* desc, int ndesc, const float* const* params,
                                 int nparams, LionModel** out) {
  LION_REQUIRE(ctx && out && (desc || ndesc == 0) && (params || nparams == 0), "lion_model_create: null argument");
  LION_CHECK_CUDA(cudaSetDevice(ctx->c.device));
  std::unique_ptr<LionModel> h(new LionModel());
  Model* m = &h->m;
  m->ctx = &ctx->c; m->kind = kind;
  m->desc.assign(desc, desc + ndesc);
  m->params.assign(params, params + nparams);
  for (int i = 0; i < nparams; ++i) LION_REQUIRE(params[i], "lion_model_create: parameter %d is null", i);
  Cursor cur{m->params.data(), nparams};
  const std::vector<int>& d = m->desc;
  auto need = [&](int n) { return (int)d.size() >= n; };
  switch (kind) {
    case LION_KIND_UNET: LION_TRY(build_unet(m, cur)); break;
    // style_dim == 0 selects the non-Ada blocks of models/pvcnn2.py (plain GroupNorm(8), no `emd` parameters, no style)
    case LION_KIND_PVCONV: {   // [cin, cout, r, attn, S]
      LION_REQUIRE(need(5), "pvconv descriptor: [cin, cout, r, attn, style_dim]");
      m->S = d[4] > 0 ? d[4] : 4;
      m->block.reset(new Block()); m->block->kind = kind;
      LION_TRY(make_pvconv(m, m->block->pv, cur, d[0], d[1], d[2], d[3] != 0, d[4] == 0));
      break;
    }
    case LION_KIND_SA: {       // [cfeat, m, radius_bits, k, S, n_mlp, mlp...]
      LION_REQUIRE(need(6) && need(6 + d[5]), "sa descriptor: [cfeat, m, radius_bits, k, style_dim, n, outs...]");
      m->S = d[4] > 0 ? d[4] : 4;
      m->block.reset(new Block()); m->block->kind = kind;
      LION_TRY(make_sa(m, m->block->sa, cur, d[0], d[1], bits_to_float(d[2]), d[3], std::vector<int>(d.begin() + 6, d.begin() + 6 + d[5]), d[4] == 0));
      break;
    }
    case LION_KIND_FP: {       // [cc, cp, S, n_mlp, mlp...]
      LION_REQUIRE(need(4) && need(4 + d[3]), "fp descriptor: [cc, cp, style_dim, n, outs...]");
      m->S = d[2];
      m->block.reset(new Block()); m->block->kind = kind;
      LION_TRY(make_fp(m, m->block->fp, cur, d[0], d[1], std::vector<int>(d.begin() + 4, d.begin() + 4 + d[3])));
      break;
    }
    case LION_KIND_ATTN: {     // [C, heads]
      LION_REQUIRE(need(2), "attention descriptor: [C, heads]");
      LION_REQUIRE(d[0] % 4 == 0, "attention: C must be a multiple of 4");
      m->attn.reset(new AttnBlk());
      LION_TRY(make_attn(m, *m->attn, cur, d[0], d[1]));
      break;
    }
    case LION_KIND_SHARED_MLP: {   // [cin, S, n, outs...]
      LION_REQUIRE(need(3) && need(3 + d[2]), "shared_mlp descriptor: [cin, style_dim, n, outs...]");
      m->S = d[1] > 0 ? d[1] : 4;
      m->mlp.reset(new SharedMLPBlk());
      LION_TRY(make_shared_mlp(m, *m->mlp, cur, d[0], ident_map(d[0]), std::vector<int>(d.begin() + 3, d.begin() + 3 + d[2]), d[1] == 0));
      break;
    }
    case LION_KIND_GLOBAL_PRIOR: LION_TRY(global_prior_build(m, cur)); break;
    case LION_KIND_STYLE_ENC: LION_TRY(build_style_enc(m, cur)); break;
    case LION_KIND_ADAGN: {        // [C, S]
      LION_REQUIRE(need(2), "adagn descriptor: [C, style_dim]");
      m->S = d[1];
      LION_TRY(make_adagn(m, m->gn_single, cur, d[0]));
      break;
    }
    case LION_KIND_CONV3D: {       // [cin, cout, r]
      LION_REQUIRE(need(3) && d[0] > 0 && d[1] > 0 && d[2] > 0 && d[2] <= 64, "conv3d descriptor: [cin, cout, r]");
      const float* w = cur.next(); const float* bi = cur.next();
      LION_REQUIRE(!cur.bad, "conv3d: parameters are [weight, bias]");
      LION_TRY(make_conv(m, m->conv_single, w, bi, 27, d[0], d[1], ident_map(d[0])));
      break;
    }
    default: LION_REQUIRE(false, "lion_model_create: unknown kind %d", kind);
  }
  LION_REQUIRE(!cur.bad && cur.i == cur.n, "lion_model_create(kind %d): %d parameters given, %d consumed", kind, cur.n, cur.i);
  if (!m->style_layers.empty()) {
    LION_TRY(m->dmalloc(&m->d_style_layers, m->style_layers.size()));
    LION_CHECK_CUDA(cudaMemcpy(m->d_style_layers, m->style_layers.data(), m->style_layers.size() * sizeof(StyleLayer), cudaMemcpyHostToDevice));
  }
  LION_TRY(run_jobs(m));
  *out = h.release();
  return 0;
}
extern "C" int lion_model_destroy(LionModel* h) { delete h; return 0; }
extern "C" int lion_model_refresh(LionModel* h) {
  LION_REQUIRE(h, "lion_model_refresh: null model");
  LION_CHECK_CUDA(cudaSetDevice(h->m.ctx->device));
  return run_jobs(&h->m);
}

extern "C" int lion_unet_forward(LionModel* h, const float* x, const float* t, const float* style, const float* clip,
                                 float* out, int B, int N, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_UNET, "lion_unet_forward: not a unet model");
  LION_REQUIRE(x && out && B > 0 && N > 0, "lion_unet_forward: bad arguments");
  Model* m = &h->m;
  return two_pass(m, stream, B, [&](Fwd& f) { return unet_forward(f, x, t, style, clip, out, N); });
}

extern "C" int lion_style_encoder_forward(LionModel* h, const float* x, float* out, int B, int N, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_STYLE_ENC, "lion_style_encoder_forward: not a style-encoder model");
  LION_REQUIRE(x && out && B > 0 && N > 0, "lion_style_encoder_forward: bad arguments");
  Model* m = &h->m;
  return two_pass(m, stream, B, [&](Fwd& f) { return style_enc_forward(f, x, out, N); });
}

// Hoists everything that depends on the style only out of the denoising loop (the reference recomputes the 61 AdaGN
// Linears and the CLIP mixing every step, models/adagn.py:59-61, latent_points_ada.py:132-137): computes them once
// into a buffer owned by the model; later lion_unet_forward calls with style == NULL use it.  Not capturable when the
// buffer has to grow (first call / larger B).
extern "C" int lion_unet_cache_style(LionModel* h, const float* style, const float* clip, int B, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_UNET, "lion_unet_cache_style: not a unet model");
  LION_REQUIRE(style && B > 0, "lion_unet_cache_style: bad arguments");
  Model* m = &h->m;
  if (m->aff_cache_B < B) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing((cudaStream_t)stream, &st);
    LION_REQUIRE(st == cudaStreamCaptureStatusNone, "lion_unet_cache_style: the cache must be sized outside stream capture");
    LION_CHECK_CUDA(cudaSetDevice(m->ctx->device));
    LION_CHECK_CUDA(cudaDeviceSynchronize());
    if (m->aff_cache) LION_CHECK_CUDA(cudaFree(m->aff_cache));
    m->aff_cache = nullptr; m->aff_cache_B = 0;
    LION_CHECK_CUDA(cudaMalloc((void**)&m->aff_cache, (size_t)B * m->style_total * sizeof(float) + 16));
  }
  m->aff_cache_B = B;
  return two_pass(m, stream, B, [&](Fwd& f) -> int {
    LION_TRY(unet_style_affine(f, style, clip));
    if (!f.c->dry) LION_CHECK_CUDA(cudaMemcpyAsync(m->aff_cache, f.aff, (size_t)B * m->style_total * sizeof(float), cudaMemcpyDeviceToDevice, f.c->stream));
    return 0;
  });
}

// ---- block-level entry points on the reference's channel-major layouts -------------------
namespace {
PF to_pf(Fwd& f, const float* src, int C, int R) {
  PF p = alloc_pf(f, (C + 3) / 4, R);
  LION_LAUNCH(f.c, k_cm_to_pf, dim3(cdiv(R, 256), p.G, f.B), 256, 0, src, p.p, C, p.G, R);
  return p;
}
void from_pf(Fwd& f, PF p, float* dst, int C) {
  LION_LAUNCH(f.c, k_pf_to_cm, dim3(cdiv(p.R, 256), p.G, f.B), 256, 0, p.p, dst, C, p.G, p.R);
}
float4* to_c4(Fwd& f, const float* src, int N) {
  float4* c = f.c->alloc_n<float4>((size_t)f.B * N);
  LION_LAUNCH(f.c, k_cm_to_c4, dim3(cdiv(N, 256), f.B), 256, 0, src, c, N);
  return c;
}
}  // namespace

extern "C" int lion_pvconv_fwd(LionModel* h, const float* features, const float* coords, const float* style, float* out,
                               int B, int N, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_PVCONV, "lion_pvconv_fwd: not a pvconv model");
  LION_REQUIRE(features && coords && (style || h->m.desc[4] == 0) && out && B > 0 && N > 0, "lion_pvconv_fwd: bad arguments");
  Model* m = &h->m;
  return two_pass(m, stream, B, [&](Fwd& f) {
    const PVConvBlk& p = m->block->pv;
    LION_TRY(style_affine_all(f, style ? style : f.c->alloc_n<float>((size_t)B * 4)));
    PF x = to_pf(f, features, m->desc[0], N);
    float4* c4 = to_c4(f, coords, N);
    PF o = alloc_pf(f, p.cout / 4, N);
    LION_TRY(pvconv_fwd(f, p, x, c4, o.p, o.G, 0));
    from_pf(f, o, out, p.cout);
    return check_launch(f.c, "lion_pvconv_fwd");
  });
}

extern "C" int lion_sa_module_fwd(LionModel* h, const float* features, const float* coords, const float* style,
                                  float* out_features, float* out_coords, int B, int N, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_SA, "lion_sa_module_fwd: not an SA model");
  LION_REQUIRE(features && coords && (style || h->m.desc[4] == 0) && out_features && out_coords && B > 0 && N > 0, "lion_sa_module_fwd: bad arguments");
  Model* m = &h->m;
  return two_pass(m, stream, B, [&](Fwd& f) {
    const SABlk& s = m->block->sa;
    LION_TRY(style_affine_all(f, style ? style : f.c->alloc_n<float>((size_t)B * 4)));
    PF x = to_pf(f, features, s.cfeat, N);
    float4* c4 = to_c4(f, coords, N);
    float4* ctr = f.c->alloc_n<float4>((size_t)B * s.m);
    PF o = alloc_pf(f, s.mlp.cout() / 4, s.m);
    LION_TRY(sa_fwd(f, s, x, c4, ctr, o.p, o.G, 0));
    from_pf(f, o, out_features, s.mlp.cout());
    LION_LAUNCH(f.c, k_c4_to_cm, dim3(cdiv(s.m, 256), B), 256, 0, ctr, out_coords, s.m);
    return check_launch(f.c, "lion_sa_module_fwd");
  });
}

extern "C" int lion_fp_module_fwd(LionModel* h, const float* points_coords, const float* centers_coords,
                                  const float* centers_features, const float* points_features, const float* style,
                                  float* out, int B, int N, int M, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_FP, "lion_fp_module_fwd: not an FP model");
  LION_REQUIRE(points_coords && centers_coords && centers_features && style && out && B > 0 && N > 0 && M > 0, "lion_fp_module_fwd: bad arguments");
  Model* m = &h->m;
  LION_REQUIRE((m->block->fp.cp > 0) == (points_features != nullptr), "lion_fp_module_fwd: points_features presence does not match the module");
  return two_pass(m, stream, B, [&](Fwd& f) {
    const FPBlk& b = m->block->fp;
    LION_TRY(style_affine_all(f, style));
    float4* pc = to_c4(f, points_coords, N);
    float4* cc = to_c4(f, centers_coords, M);
    PF cf = to_pf(f, centers_features, b.cc, M);
    PF sk;
    if (b.cp) sk = to_pf(f, points_features, b.cp, N);
    PF o = alloc_pf(f, b.mlp.cout() / 4, N);
    LION_TRY(fp_fwd(f, b, pc, N, cc, M, cf, sk, o.p, o.G, 0));
    from_pf(f, o, out, b.mlp.cout());
    return check_launch(f.c, "lion_fp_module_fwd");
  });
}

extern "C" int lion_linear_attention_fwd(LionModel* h, const float* x, float* out, int B, int N, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_ATTN, "lion_linear_attention_fwd: not an attention model");
  LION_REQUIRE(x && out && B > 0 && N > 0, "lion_linear_attention_fwd: bad arguments");
  Model* m = &h->m;
  return two_pass(m, stream, B, [&](Fwd& f) {
    const AttnBlk& a = *m->attn;
    PF xi = to_pf(f, x, a.C, N);
    PF o = alloc_pf(f, a.C / 4, N);
    LION_TRY(attn_fwd(f, a, xi, o.p, o.G, 0));
    from_pf(f, o, out, a.C);
    return check_launch(f.c, "lion_linear_attention_fwd");
  });
}

extern "C" int lion_shared_mlp_fwd(LionModel* h, const float* x, const float* style, float* out, int B, int R, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_SHARED_MLP, "lion_shared_mlp_fwd: not a shared-mlp model");
  LION_REQUIRE(x && (style || h->m.desc[1] == 0) && out && B > 0 && R > 0, "lion_shared_mlp_fwd: bad arguments");
  Model* m = &h->m;
  return two_pass(m, stream, B, [&](Fwd& f) {
    const SharedMLPBlk& s = *m->mlp;
    LION_TRY(style_affine_all(f, style ? style : f.c->alloc_n<float>((size_t)B * 4)));
    PF xi = to_pf(f, x, s.conv[0].cin_ref, R);
    PF o = alloc_pf(f, s.cout() / 4, R);
    LION_TRY(shared_mlp_fwd(f, s, xi, 1, o.p, o.G, 0));
    from_pf(f, o, out, s.cout());
    return check_launch(f.c, "lion_shared_mlp_fwd");
  });
}

// ---- stand-alone AdaGN / SE3d / Swish (reference interfaces models/adagn.py:45-65,
// models/pvcnn2_ada.py:27-41, :74-83); the fused path never calls these -------------------------
extern "C" int lion_adagn_fwd(LionModel* h, const float* x, const float* style, float* out, int B, int R, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_ADAGN, "lion_adagn_fwd: not an AdaGN model");
  LION_REQUIRE(x && style && out && B > 0 && R > 0, "lion_adagn_fwd: bad arguments");
  Model* m = &h->m;
  return two_pass(m, stream, B, [&](Fwd& f) {
    const AdaGNW& g = m->gn_single;
    LION_TRY(style_affine_all(f, style));
    PF xi = to_pf(f, x, g.C, R);
    double *ssum, *ssq;
    LION_TRY(alloc_stats(f, g.C, &ssum, &ssq));
    LION_LAUNCH(f.c, k_row_stats, dim3(cdiv(R, 1024) > 64 ? 64 : cdiv(R, 1024), xi.G, B), 256, 0, xi.p, ssum, ssq, xi.G, R, g.C);
    AffSrc a;
    LION_TRY(run_affine(f, g, ssum, ssq, g.C, (double)R, nullptr, nullptr, a));
    PF o = alloc_pf(f, xi.G, R);
    LION_LAUNCH(f.c, k_act_rows<1>, dim3(cdiv(R, 256), xi.G, B), 256, 0, xi.p, o.p, a, xi.G, g.C, R, xi.G, 0, 2);
    from_pf(f, o, out, g.C);
    return check_launch(f.c, "lion_adagn_fwd");
  });
}
extern "C" int lion_se3d_fwd(LionCtx* ctx, const float* w1, const float* w2, const float* x, float* out, int B, int C, int V,
                             void* stream) {
  LION_REQUIRE(ctx && w1 && w2 && x && out && B > 0 && C > 0 && C % 8 == 0 && C <= 1024 && V > 0, "lion_se3d_fwd: bad arguments");
  Model tmp;
  tmp.ctx = &ctx->c;
  return two_pass(&tmp, stream, B, [&](Fwd& f) {
    PF xi = to_pf(f, x, C, V);
    double *ssum, *ssq;
    LION_TRY(alloc_stats(f, xi.G * 4, &ssum, &ssq));
    LION_LAUNCH(f.c, k_row_stats, dim3(cdiv(V, 1024) > 64 ? 64 : cdiv(V, 1024), xi.G, B), 256, 0, xi.p, ssum, ssq, xi.G, V, xi.G * 4);
    float* gate = f.c->alloc_n<float>((size_t)B * C);
    LION_LAUNCH(f.c, k_se_gate, B, C, (C + C / 8) * sizeof(float), ssum, xi.G * 4, w1, w2, gate, C, (double)V);
    size_t total = (size_t)B * C * V;
    LION_LAUNCH(f.c, k_scale_or_swish, (unsigned)cdivz(total, 256), 256, 0, x, gate, out, (size_t)V, total);
    return check_launch(f.c, "lion_se3d_fwd");
  });
}
extern "C" int lion_swish_fwd(const float* x, float* out, size_t n, void* stream) {
  LION_REQUIRE(x && out && n > 0, "lion_swish_fwd: bad arguments");
  Ctx c;
  c.stream = (cudaStream_t)stream;
  LION_LAUNCH(&c, k_scale_or_swish, (unsigned)cdivz(n, 256), 256, 0, x, (const float*)nullptr, out, (size_t)1, n);
  return check_launch(&c, "lion_swish_fwd");
}

// ---- stand-alone Conv3d 3x3x3 + fused GroupNorm statistics (reference: nn.Conv3d in
// models/pvcnn2_ada.py:211-222 followed by AdaGN's GroupNorm, models/adagn.py:36) -----------------
extern "C" int lion_conv3d_gn_fwd(LionModel* h, const float* x, float* out, double* gn_sum, double* gn_sqsum, int B, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_CONV3D, "lion_conv3d_gn_fwd: not a conv3d model");
  LION_REQUIRE(x && out && B > 0 && ((gn_sum == nullptr) == (gn_sqsum == nullptr)), "lion_conv3d_gn_fwd: bad arguments");
  Model* m = &h->m;
  return two_pass(m, stream, B, [&](Fwd& f) -> int {
    const ConvW& w = m->conv_single;
    const int cin = m->desc[0], cout = m->desc[1], r = m->desc[2];
    const int Gin = w.cin_pad / 4, Gout = (cout + 3) / 4, rp = r + 2, V = r * r * r;
    const size_t P = (size_t)rp * rp * rp;
    float4* gi = alloc_vg(f, Gin, r);
    LION_TRY(memset_async(f.c, gi, 0, sizeof(float4) * (size_t)B * Gin * P));
    LION_LAUNCH(f.c, k_cm_to_vg, dim3(cdiv(V, 256), Gin, B), 256, 0, x, gi, cin, Gin, r, 1);
    float4* go = alloc_vg(f, Gout, r);
    double *ssum = nullptr, *ssq = nullptr;
    if (gn_sum) LION_TRY(alloc_stats(f, w.cout_pad, &ssum, &ssq));
    LION_TRY(run_conv(f, w, gi, Gin, go, Gout, ssum, ssq, geom_grid(r)));
    LION_LAUNCH(f.c, k_vg_to_cm, dim3(cdiv(V, 256), Gout, B), 256, 0, go, out, cout, Gout, r);
    if (gn_sum && !f.c->dry) {
      LION_CHECK_CUDA(cudaMemcpy2DAsync(gn_sum, sizeof(double) * cout, ssum, sizeof(double) * w.cout_pad, sizeof(double) * cout, B,
                                        cudaMemcpyDeviceToDevice, f.c->stream));
      LION_CHECK_CUDA(cudaMemcpy2DAsync(gn_sqsum, sizeof(double) * cout, ssq, sizeof(double) * w.cout_pad, sizeof(double) * cout, B,
                                        cudaMemcpyDeviceToDevice, f.c->stream));
    }
    return check_launch(f.c, "lion_conv3d_gn_fwd");
  });
}

extern "C" int lion_global_prior_step(LionModel* h, const float* x, const float* t, const float* clip, float* out, int B,
                                      void* stream) {
  return lion_global_prior_forward(h, x, t, clip, out, B, stream);
}

extern "C" int lion_global_prior_forward(LionModel* h, const float* x, const float* t, const float* clip, float* out,
                                         int B, void* stream) {
  LION_REQUIRE(h && h->m.kind == LION_KIND_GLOBAL_PRIOR, "lion_global_prior_forward: not a global-prior model");
  LION_REQUIRE(x && t && out && B > 0, "lion_global_prior_forward: bad arguments");
  Model* m = &h->m;
  return two_pass(m, stream, B, [&](Fwd& f) { (void)f; return global_prior_forward(m, x, t, clip, out, B); });
}

// ---- measurement hook: time the convolution kernel alone (bench.py roofline leg) ------------
__global__ void k_fill_pattern(float* p, size_t n, float scale) {
  pdl_prologue();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { unsigned h = (unsigned)(i * 2654435761u) ^ 0x9e3779b9u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
               p[i] = scale * ((float)(h & 0xffff) / 32768.0f - 1.0f); }
}
extern "C" int lion_bench_conv(LionCtx* ctx, int ntaps, int cin, int cout, int r_or_rows, int B, int iters, int warmup,
                               float* ms_out, double* flops_out, void* stream) {
  LION_REQUIRE(ctx && ms_out && flops_out && (ntaps == 27 || ntaps == 1) && cin % 4 == 0 && cout % 8 == 0 && B > 0 && iters > 0,
               "lion_bench_conv: bad arguments");
  LION_CHECK_CUDA(cudaSetDevice(ctx->c.device));
  LionModel h;
  Model* m = &h.m;
  m->ctx = &ctx->c;
  float *wref = nullptr, *bref = nullptr;
  size_t nw = (size_t)cout * cin * ntaps;
  LION_TRY(m->dmalloc(&wref, nw));
  LION_TRY(m->dmalloc(&bref, (size_t)cout));
  k_fill_pattern<<<(unsigned)cdivz(nw, 256), 256>>>(wref, nw, 0.05f);
  k_fill_pattern<<<cdiv(cout, 256), 256>>>(bref, cout, 0.1f);
  ConvW w;
  LION_TRY(make_conv(m, w, wref, bref, ntaps, cin, cout, ident_map(cin)));
  LION_TRY(run_jobs(m));
  ConvGeom geo = ntaps == 27 ? geom_grid(r_or_rows) : geom_rows(r_or_rows);
  size_t guard = ntaps == 27 ? (size_t)(r_or_rows + 2) * (r_or_rows + 2) + (r_or_rows + 2) + 8 : 256;
  size_t n_in = (size_t)B * (cin / 4) * geo.rows + 2 * guard, n_out = (size_t)B * (cout / 4) * geo.rows + 2 * guard;
  float4 *din = nullptr, *dout = nullptr;
  double* stats = nullptr;
  LION_TRY(m->dmalloc(&din, n_in));
  LION_TRY(m->dmalloc(&dout, n_out));
  LION_TRY(m->dmalloc(&stats, (size_t)2 * B * w.cout_pad));
  k_fill_pattern<<<(unsigned)cdivz(n_in * 4, 256), 256>>>((float*)din, n_in * 4, 1.0f);
  LION_CHECK_CUDA(cudaMemset(stats, 0, sizeof(double) * 2 * B * w.cout_pad));
  LION_CHECK_CUDA(cudaDeviceSynchronize());
  Ctx* c = &ctx->c;
  c->stream = (cudaStream_t)stream;
  c->dry = false;
  Fwd f{c, m, B};
  cudaEvent_t e0, e1;
  LION_CHECK_CUDA(cudaEventCreate(&e0));
  LION_CHECK_CUDA(cudaEventCreate(&e1));
  for (int i = 0; i < warmup; ++i)
    LION_TRY(run_conv(f, w, din + guard, cin / 4, dout + guard, cout / 4, stats, stats + (size_t)B * w.cout_pad, geo));
  LION_CHECK_CUDA(cudaEventRecord(e0, c->stream));
  for (int i = 0; i < iters; ++i)
    LION_TRY(run_conv(f, w, din + guard, cin / 4, dout + guard, cout / 4, stats, stats + (size_t)B * w.cout_pad, geo));
  LION_CHECK_CUDA(cudaEventRecord(e1, c->stream));
  LION_CHECK_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  LION_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms / iters;
  // algorithmic FLOPs of the dense convolution (interior voxels / rows only, no halo work counted)
  double rows = ntaps == 27 ? (double)r_or_rows * r_or_rows * r_or_rows : (double)r_or_rows;
  *flops_out = 2.0 * B * rows * ntaps * (double)cin * cout;
  return 0;
}
