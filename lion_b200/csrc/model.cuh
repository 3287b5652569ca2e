// lion_b200 -- host-side model description shared by net.cu, conv_tc.cu and global_prior.cu.
#pragma once
#include <memory>
#include <vector>
#include "common.cuh"

namespace lion {

struct StyleLayer { const float* w; const float* b; int n_out; int out_off; };

// tensor-core packing of a convolution's weights (conv_tc.cu); w == nullptr -> SIMT kernel only
struct ConvTcW {
  float* w = nullptr;
  int ck = 0;        // input channels per pipeline chunk (8, 16 or 32)
  int nchunk = 0;
  int n = 0;         // UMMA N (= padded output channels)
};

struct Cursor {
  const float* const* p;
  int n, i = 0;
  bool bad = false;
  const float* next() {
    if (i >= n) { bad = true; return nullptr; }
    return p[i++];
  }
};

struct ConvW {
  int ntaps = 1, cin_ref = 0, cin_pad = 0, cout = 0, cout_pad = 0;
  const float* w_ref = nullptr;
  const float* b_ref = nullptr;
  int* d_kmap = nullptr;
  float* wt = nullptr;      // [ntaps][cin_pad][cout_pad]   (SIMT kernel)
  float* bias = nullptr;    // [cout_pad]
  ConvTcW tc;               // tensor-core packing (conv_tc.cuh); tc.w == nullptr when unsupported
};
struct AdaGNW {
  const float* gamma = nullptr; const float* beta = nullptr;
  int C = 0; int style_off = 0;
};

struct Model;
struct Fwd;

struct SharedMLPBlk {
  std::vector<ConvW> conv;
  std::vector<AdaGNW> gn;
  int cin_pad = 0;
  int cout() const { return conv.back().cout; }
};
struct AttnBlk {
  ConvW qkv, out;
  int C = 0, heads = 0;
};
struct PVConvBlk {
  int cin = 0, cout = 0, r = 0;
  ConvW c1, c2;
  ConvW c1y;            // sparse first convolution: 1x1 "convolution" x[v] -> y[v][tap][cout] (c1y.tc.w set when available)
  AdaGNW g1, g2;
  const float* se1 = nullptr; const float* se2 = nullptr;
  SharedMLPBlk point;
  bool has_attn = false;
  AttnBlk attn;
};
struct SABlk {
  int cfeat = 0, m = 0, k = 0;
  float radius = 0;
  SharedMLPBlk mlp;
};
struct FPBlk {
  int cc = 0, cp = 0;   // interpolated channels, skip channels
  SharedMLPBlk mlp;
};
struct Block {
  int kind;   // LION_KIND_PVCONV / SA / FP
  PVConvBlk pv; SABlk sa; FPBlk fp;
};
struct UnetBlk {
  int num_classes, embed_dim, extra, input_dim, use_att, clip, clip_dim, S;
  const float *e0w = nullptr, *e0b = nullptr, *e2w = nullptr, *e2b = nullptr;
  const float *cfw = nullptr, *cfb = nullptr, *scw = nullptr, *scb = nullptr;
  float* d_freqs = nullptr;
  std::vector<std::vector<Block>> sa, fp;
  AttnBlk gatt;
  SharedMLPBlk cls0;
  ConvW cls2;
};
// PointNetPlusEncoder (models/shapelatent_modules.py:13-52): SA levels on the non-Ada blocks, max over points, Linear
struct StyleEncBlk {
  int input_dim = 3, zdim = 128, cfeat = 0;
  std::vector<std::vector<Block>> sa;
  const float* mlp_w = nullptr; const float* mlp_b = nullptr;
};
struct GlobalPriorBlk;   // global_prior.cu
int global_prior_build(Model* m, Cursor& cur);
int global_prior_forward(Model* m, const float* x, const float* t, const float* clip, float* out, int B);
void global_prior_free(GlobalPriorBlk*);

struct PackJob { int type; const float* src; const int* kmap; float* dst; int a, b, c, d, e; };

struct Model {
  Ctx* ctx = nullptr;
  int kind = 0;
  std::vector<int> desc;
  std::vector<const float*> params;
  std::vector<void*> owned;
  std::vector<PackJob> jobs;
  std::vector<StyleLayer> style_layers;
  StyleLayer* d_style_layers = nullptr;
  int style_total = 0;
  int S = 128;
  float* aff_cache = nullptr;  // [aff_cache_B][style_total]: the AdaGN style Linears of a step-invariant style (lion_unet_cache_style)
  int aff_cache_B = 0;
  std::unique_ptr<UnetBlk> unet;
  std::unique_ptr<Block> block;
  std::unique_ptr<AttnBlk> attn;
  std::unique_ptr<SharedMLPBlk> mlp;
  std::unique_ptr<StyleEncBlk> senc;
  GlobalPriorBlk* gp = nullptr;
  AdaGNW gn_single;            // LION_KIND_ADAGN
  ConvW conv_single;           // LION_KIND_CONV3D

  template <typename T> int dmalloc(T** p, size_t n) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 16);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return LION_ERR_OOM; }
    owned.push_back(q);
    *p = (T*)q;
    return 0;
  }
  ~Model() {
    for (void* q : owned) cudaFree(q);
    if (aff_cache) cudaFree(aff_cache);
    if (gp) global_prior_free(gp);
  }
};


int make_conv(Model* m, ConvW& w, const float* w_ref, const float* b_ref, int ntaps, int cin_ref, int cout,
              const std::vector<int>& kmap);
std::vector<int> ident_map(int c);
static inline int roundup(int a, int b) { return (a + b - 1) / b * b; }

// geometry of one convolution launch (rows, tap offsets, halo mask)
struct ConvGeom {
  int ntaps;
  int off[27];
  int rp;        // r+2 for a VG, 0 for a PF (no halo mask)
  int rows;      // rows per (b, group): P or R
  int p_begin, p_end;
  const unsigned char* occ;   // optional 64-row occupancy flags of the INPUT ([B][occ_stride]); null = dense
  int occ_stride;
};

// conv_tc.cu
int conv_tc_prepare(Model* m, ConvW& w);
int conv_tc_pack_job(const PackJob& j);
bool conv_tc_usable(const ConvW& w, const ConvGeom& geo);
// sparse first convolution, GEMM half (sparse_conv.cu)
bool ygemm_usable(const ConvW& y);
int ygemm_run(Ctx* c, const ConvW& y, const float4* xc, float* out, int ld, const int* nocc, int B, int N);
// fused set-abstraction MLP (sa_fused.cu)
bool sa_fused_usable(const SABlk& s);
int sa_fused_run(Ctx* c, const SABlk& s, const float4* feat, const float4* points, const float4* centers, const int* nidx,
                 const float* scale1, const float* shift1, double* ssum, double* ssq, int stat_stride, float* pool_mm,
                 int B, int N);
int conv_tc_run(Ctx* c, const ConvW& w, const float4* in, int Gin, float4* out, int Gout_store, double* ssum,
                double* ssq, const ConvGeom& geo, int B, float* pool_mm = nullptr, float* out_rm = nullptr, int ld_rm = 0);

}  // namespace lion
