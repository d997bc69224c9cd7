// Edgewise (Mixture-of-Products) attention core, fp32 mode.
//
// One persistent CTA per (batch, head) problem; every N x N map of the
// algorithm lives in a per-CTA scratch slot (global memory, L1/L2 resident for
// the Edgewise shapes N <= 196) and is touched only by that CTA.  Contractions
// are fp32 CUDA-core GEMMs (simt_blas.cuh); softmax / log / gate statistics are
// fp32.  This path meets the 1e-5 fp32-mode parity bar and covers every
// configuration (any N, dk, V, shared or per-view projections, low-rank or
// dense gate head with or without the 3x3 stage); the tcgen05 kernels in
// edgewise_tc.cuh cover the bf16 hot shapes.
//
// Math: SURVEY.md appendix A / D.1; reference attention_variants.py:500-562
// (forward) and :311-331 (gate head).  The executable specification is
// oracle/edgewise_manual.py.
#pragma once
#include "simt_blas.cuh"

namespace mop {
namespace ew {

constexpr int kMaxViews = 8;
constexpr int kMaxLens = 4;                                           // dilations of the S lens bank
constexpr int kMaxFeat = 2 * kMaxViews + 2 + kMaxViews * kMaxLens;    // feature channels of the gate head
constexpr int kMaxRank = 8;
constexpr int kMaxHidden = 16;

// Per-CTA scratch layout, in floats.  Shared by host (sizing) and device.
struct Layout {
  int N, dk, V, Vp, C, r, hid, dense, k3, bwd;
  int nl;      // S lens bank (attention_variants.py:427-442, :523-533): number of dilations; V * nl extra feature channels
  int mL;      // first lens feature map
  int hops;    // length of the chain product: V for the Edgewise modes, `hops` for the fixed-gate multi-hop mode (views 0,1,1,..)
  int cgate;   // fixed scalar gates (MultiHopMSA): no gate head, no reverse chain
  size_t N2, Nd;
  int n_maps;
  // map indices
  int mS, mA, mP, mR, mM, mG, mD, mdA, mX0, mX1, mdG, mZ1, mH2, mDH2, mX2;
  // vector offsets
  size_t o_maps, o_qb, o_kb, o_vb, o_yacc, o_rho, o_kap, o_a, o_b, o_cvec, o_vs1, o_vsL;
  size_t o_dyf, o_dv1, o_dvl, o_tt, o_dqa, o_dka, o_da, o_db, o_drho, o_dkap, o_z;
  size_t total;

  __host__ __device__ void build(int N_, int dk_, int V_, int Vp_, int r_, int hid_, int dense_, int k3_, int bwd_, int hops_ = 0,
                                 int cgate_ = 0, int nl_ = 0) {
    N = N_; dk = dk_; V = V_; Vp = Vp_; r = r_; hid = hid_; dense = dense_; k3 = k3_; bwd = bwd_;
    hops = hops_ > 0 ? hops_ : V_; cgate = cgate_; nl = cgate_ ? 0 : nl_;
    C = 2 * V + 2 + V * nl;
    N2 = (size_t)N * N; Nd = (size_t)N * dk;
    int m = 0;
    mS = m; m += V;
    mA = m; m += V;
    mP = m; m += hops - 1;
    mR = m; m += cgate ? 0 : V - 1;
    mM = m; m += 1;
    mG = m; m += 4;
    mL = m; m += V * nl;
    mZ1 = mH2 = mDH2 = mX2 = -1;
    if (dense) { mZ1 = m; m += hid; if (k3) { mH2 = m; m += hid; mX2 = m; m += hid; } }   // X2 = gelu(gelu(z1)): the input of the 3x3 stage
    mD = mdA = mX0 = mX1 = mdG = -1;
    if (bwd) {
      mD = m; m += 1;
      mdA = m; m += V;
      mX0 = m; m += 1;
      mX1 = m; m += 1;
      mdG = m; m += 4;
      if (dense && k3) { mDH2 = m; m += hid; }
    }
    n_maps = m;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r0 = o; o += (n + 3) / 4 * 4; return r0; };
    o_maps = take((size_t)n_maps * N2);
    o_qb = take((size_t)Vp * Nd);
    o_kb = take((size_t)Vp * Nd);
    o_vb = take(2 * Nd);
    o_yacc = take(Nd);
    o_rho = take((size_t)C * N);
    o_kap = take((size_t)C * N);
    o_a = take((size_t)4 * r * N);
    o_b = take((size_t)4 * r * N);
    o_cvec = take((size_t)V * dk);
    o_vs1 = take(dk);
    o_vsL = take(dk);
    o_dyf = o_dv1 = o_dvl = o_tt = o_dqa = o_dka = o_da = o_db = o_drho = o_dkap = o_z = 0;
    if (bwd) {
      o_dyf = take(Nd); o_dv1 = take(Nd); o_dvl = take(Nd); o_tt = take(Nd);
      o_dqa = take((size_t)Vp * Nd); o_dka = take((size_t)Vp * Nd);
      o_da = take((size_t)4 * r * N); o_db = take((size_t)4 * r * N);
      o_drho = take((size_t)C * N); o_dkap = take((size_t)C * N);
      o_z = take((size_t)V * dk);
    }
    total = o;
  }
};

__host__ __device__ inline size_t head_param_count(int gate_mode, int V, int r, int hid, int k3, int nl = 0) {
  int C = 2 * V + 2 + V * nl;
  if (gate_mode == MOP_GATE_LOWRANK) return (size_t)2 * (4 * r * C + 4 * r);
  return (size_t)hid * C + hid + (k3 ? (size_t)hid * hid * 9 + hid : 0) + 4 * hid + 4;
}

struct Ctx {
  const MopEdgewiseParams& p;
  const Layout& L;
  float* ws;
  simt::GemmSmem& gs;
  float* red;
  int b, h, g;
  float w;  // sigmoid(chain_value_logit)
  __device__ float* map(int idx) const { return ws + L.o_maps + (size_t)idx * L.N2; }
  __device__ float* S(int i) const { return map(L.mS + i); }
  __device__ float* A(int i) const { return map(L.mA + i); }
  // P(k) = A_{cv(0)}..A_{cv(k)}, R(k) = A_V..A_{k+1}; cv(k) = view at chain position k
  __device__ int cv(int k) const { return k < L.V ? k : L.V - 1; }
  __device__ float* P(int k) const { return k == 0 ? A(0) : map(L.mP + k - 1); }
  __device__ float* R(int k) const { return k == L.V - 1 ? A(L.V - 1) : map(L.mR + k); }
  __device__ float* F() const { return P(L.hops - 1); }
  __device__ float* Rf() const { return R(0); }
  __device__ float* Gm(int t) const { return map(L.mG + t); }
  __device__ float* qb(int i) const { return ws + L.o_qb + (size_t)(L.Vp == 1 ? 0 : i) * L.Nd; }
  __device__ float* kb(int i) const { return ws + L.o_kb + (size_t)(L.Vp == 1 ? 0 : i) * L.Nd; }
  __device__ float* vb(int which) const { return ws + L.o_vb + (size_t)which * L.Nd; }
  __device__ float* cvec(int i) const { return ws + L.o_cvec + (size_t)i * L.dk; }
};

template <typename T>
__device__ void stage_inputs(const Ctx& c) {
  const auto& p = c.p; const auto& L = c.L;
  const int N = L.N, dk = L.dk, H = p.H, Vp = L.Vp, V = L.V;
  const T* qkv = reinterpret_cast<const T*>(p.qkv);
  const size_t hd = (size_t)H * dk;
  for (int vi = 0; vi < Vp; ++vi) {
    float* qd = c.ws + L.o_qb + (size_t)vi * L.Nd;
    float* kd = c.ws + L.o_kb + (size_t)vi * L.Nd;
    for (int idx = threadIdx.x; idx < N * dk; idx += simt::kThreads) {
      int n = idx / dk, d = idx % dk;
      size_t base = ((((size_t)c.b * N + n) * Vp + vi) * 3) * hd + (size_t)c.h * dk + d;
      qd[idx] = to_f32<T>(qkv[base]);
      kd[idx] = to_f32<T>(qkv[base + hd]);
    }
  }
  const int vlast = (Vp == 1) ? 0 : V - 1;
  for (int idx = threadIdx.x; idx < N * dk; idx += simt::kThreads) {
    int n = idx / dk, d = idx % dk;
    size_t b0 = ((((size_t)c.b * N + n) * Vp + 0) * 3 + 2) * hd + (size_t)c.h * dk + d;
    size_t b1 = ((((size_t)c.b * N + n) * Vp + vlast) * 3 + 2) * hd + (size_t)c.h * dk + d;
    c.vb(0)[idx] = to_f32<T>(qkv[b0]);
    c.vb(1)[idx] = to_f32<T>(qkv[b1]);
  }
  for (int idx = threadIdx.x; idx < V * dk; idx += simt::kThreads) {
    int i = idx / dk, d = idx % dk;
    float v = 1.f;
    if (p.q_scale) v = p.q_scale[((size_t)i * H + c.h) * dk + d] * p.k_scale[((size_t)i * H + c.h) * dk + d];
    c.ws[L.o_cvec + idx] = v;
  }
  for (int d = threadIdx.x; d < dk; d += simt::kThreads) {
    c.ws[L.o_vs1 + d] = p.v_scale ? p.v_scale[((size_t)0 * H + c.h) * dk + d] : 1.f;
    c.ws[L.o_vsL + d] = p.v_scale ? p.v_scale[((size_t)(V - 1) * H + c.h) * dk + d] : 1.f;
  }
  __syncthreads();
}

// S lens bank (reference :523-533): depthwise 3x3 convolutions of the score maps with dilation d_l and zero padding d_l,
// channel l * V + v = conv_l(S_v).  Stored as maps: the dense head reads them per pixel, the low-rank head their means.
__device__ void lens_forward(const Ctx& c) {
  const auto& p = c.p; const auto& L = c.L;
  const int N = L.N, V = L.V;
  for (int ch = 0; ch < V * L.nl; ++ch) {
    const int l = ch / V, v = ch % V, d = p.lens_dil[l];
    const float* wt = p.lens_w + (size_t)ch * 9;
    const float* S = c.S(v);
    float* out = c.map(L.mL + ch);
    for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) {
      const int i = idx / N, j = idx % N;
      float acc = 0.f;
      for (int u = 0; u < 3; ++u) {
        const int ii = i + (u - 1) * d;
        if (ii < 0 || ii >= N) continue;
        for (int x = 0; x < 3; ++x) {
          const int jj = j + (x - 1) * d;
          if (jj < 0 || jj >= N) continue;
          acc = fmaf(wt[u * 3 + x], S[(size_t)ii * N + jj], acc);
        }
      }
      out[idx] = acc;
    }
  }
  __syncthreads();
}

// Dense head: z1 = W1 feat + b1 for every pixel (stored), h2 (stored iff k3), gates -> G maps.
__device__ void dense_head_forward(const Ctx& c) {
  const auto& p = c.p; const auto& L = c.L;
  const int N = L.N, V = L.V, C = L.C, hid = L.hid;
  const float eps = p.eps;
  const float* Fm = c.F(); const float* Rm = c.Rf();
  for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) {
    int i = idx / N, j = idx % N;
    float feat[kMaxFeat];
    for (int v = 0; v < V; ++v) { feat[v] = c.S(v)[idx]; feat[V + v] = c.S(v)[(size_t)j * N + i]; }
    feat[2 * V] = logf(Fm[idx] + eps);
    feat[2 * V + 1] = logf(Rm[idx] + eps);
    for (int ch = 0; ch < V * L.nl; ++ch) feat[2 * V + 2 + ch] = c.map(L.mL + ch)[idx];
    float h[kMaxHidden];
    for (int o = 0; o < hid; ++o) {
      float z = p.conv1_b[o];
      for (int ch = 0; ch < C; ++ch) z = fmaf(p.conv1_w[o * C + ch], feat[ch], z);
      c.map(L.mZ1 + o)[idx] = z;
      h[o] = gelu_tanh_(z);
      if (L.k3) c.map(L.mX2 + o)[idx] = gelu_tanh_(h[o]);   // evaluated once per pixel (the 3x3 stage reads it at 9 neighbours)
    }
    if (!L.k3) {
      for (int t = 0; t < 4; ++t) {
        float z = p.conv2_b[t];
        for (int o = 0; o < hid; ++o) z = fmaf(p.conv2_w[t * hid + o], h[o], z);
        c.Gm(t)[idx] = sigmoidf_(z);
      }
    }
  }
  __syncthreads();
  if (!L.k3) return;
  // 3x3 stage on gelu(gelu(z1)) with zero padding (attention_variants.py:314-316)
  for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) {
    int i = idx / N, j = idx % N;
    float h2[kMaxHidden];
    for (int o = 0; o < hid; ++o) h2[o] = p.mid3_b[o];
    for (int u = 0; u < 3; ++u) {
      int ii = i + u - 1;
      if (ii < 0 || ii >= N) continue;
      for (int v = 0; v < 3; ++v) {
        int jj = j + v - 1;
        if (jj < 0 || jj >= N) continue;
        size_t q = (size_t)ii * N + jj;
        for (int ch = 0; ch < hid; ++ch) {
          const float x = c.map(L.mX2 + ch)[q];
          for (int o = 0; o < hid; ++o) h2[o] = fmaf(p.mid3_w[((o * hid + ch) * 3 + u) * 3 + v], x, h2[o]);
        }
      }
    }
    for (int o = 0; o < hid; ++o) c.map(L.mH2 + o)[idx] = h2[o];
    for (int t = 0; t < 4; ++t) {
      float z = p.conv2_b[t];
      for (int o = 0; o < hid; ++o) z = fmaf(p.conv2_w[t * hid + o], h2[o], z);
      c.Gm(t)[idx] = sigmoidf_(z);
    }
  }
  __syncthreads();
}

__device__ void lowrank_head_forward(const Ctx& c) {
  const auto& p = c.p; const auto& L = c.L;
  const int N = L.N, V = L.V, C = L.C, r = L.r;
  float* rho = c.ws + L.o_rho; float* kap = c.ws + L.o_kap;
  for (int v = 0; v < V; ++v) simt::row_col_means<false>(c.S(v), N, 0.f, rho + (size_t)v * N, kap + (size_t)v * N);
  for (int idx = threadIdx.x; idx < V * N; idx += simt::kThreads) {  // S_i^T channels swap the roles
    rho[(size_t)V * N + idx] = kap[idx];
    kap[(size_t)V * N + idx] = rho[idx];
  }
  simt::row_col_means<true>(c.F(), N, p.eps, rho + (size_t)(2 * V) * N, kap + (size_t)(2 * V) * N);
  simt::row_col_means<true>(c.Rf(), N, p.eps, rho + (size_t)(2 * V + 1) * N, kap + (size_t)(2 * V + 1) * N);
  for (int ch = 0; ch < V * L.nl; ++ch)
    simt::row_col_means<false>(c.map(L.mL + ch), N, 0.f, rho + (size_t)(2 * V + 2 + ch) * N, kap + (size_t)(2 * V + 2 + ch) * N);
  float* a = c.ws + L.o_a; float* bb = c.ws + L.o_b;
  for (int idx = threadIdx.x; idx < 4 * r * N; idx += simt::kThreads) {
    int q = idx / N, i = idx % N;
    float sa = p.row_b[q], sb = p.col_b[q];
    for (int ch = 0; ch < C; ++ch) {
      sa = fmaf(p.row_w[q * C + ch], rho[(size_t)ch * N + i], sa);
      sb = fmaf(p.col_w[q * C + ch], kap[(size_t)ch * N + i], sb);
    }
    a[idx] = sa; bb[idx] = sb;
  }
  __syncthreads();
  for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) {
    int i = idx / N, j = idx % N;
    for (int t = 0; t < 4; ++t) {
      float z = 0.f;
      for (int k = 0; k < r; ++k) z = fmaf(a[(size_t)(t * r + k) * N + i], bb[(size_t)(t * r + k) * N + j], z);
      c.Gm(t)[idx] = sigmoidf_(z);
    }
  }
  __syncthreads();
}

// Everything of the forward except writing y; leaves all maps in scratch.
template <typename T>
__device__ void forward_maps(Ctx& c) {
  const auto& p = c.p; const auto& L = c.L;
  const int N = L.N, dk = L.dk, V = L.V;
  const float s = rsqrtf((float)dk);
  stage_inputs<T>(c);
  c.w = sigmoidf_(p.chain_value_logit[0]);
  for (int i = 0; i < V; ++i) {
    simt::gemm(c.S(i), N, c.qb(i), dk, 1, c.kb(i), 1, dk, N, N, dk, c.cvec(i), nullptr, s, false, c.gs);
    simt::softmax_rows(c.A(i), c.S(i), N, N);
  }
  for (int k = 1; k < L.hops; ++k) simt::gemm(c.P(k), N, c.P(k - 1), N, 1, c.A(c.cv(k)), N, 1, N, N, N, nullptr, nullptr, 1.f, false, c.gs);
  if (L.cgate) {   // fixed scalar gates (MultiHopMSA, attention_variants.py:207-217): no gate head, no reverse chain
    for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads)
      for (int t = 0; t < 4; ++t) c.Gm(t)[idx] = p.const_gates[t];
    __syncthreads();
  } else {
    for (int k = V - 2; k >= 0; --k) simt::gemm(c.R(k), N, c.R(k + 1), N, 1, c.A(k), N, 1, N, N, N, nullptr, nullptr, 1.f, false, c.gs);
    if (L.nl) lens_forward(c);
    if (L.dense) dense_head_forward(c); else lowrank_head_forward(c);
  }
  // mix (attention_variants.py:537-547) then row softmax
  const float bn = p.beta_not / (float)max(1, V - 1);
  float* M = c.map(L.mM);
  const float* Fm = c.F();
  for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) {
    float sv[kMaxViews];
    float sum = 0.f, mx = -INFINITY;
    for (int v = 0; v < V; ++v) { sv[v] = c.S(v)[idx]; sum += sv[v]; mx = fmaxf(mx, sv[v]); }
    float se = 0.f;
    for (int v = 0; v < V; ++v) se += expf(sv[v] - mx);
    float lse = mx + logf(se);
    float U = sum - sv[0], O = lse - sv[0];
    float lf = logf(Fm[idx] + p.eps);
    M[idx] = sv[0] + c.Gm(0)[idx] * U + c.Gm(1)[idx] * O - c.Gm(2)[idx] * bn * U + c.Gm(3)[idx] * lf;
  }
  __syncthreads();
  simt::softmax_rows(M, M, N, N);
}

template <typename T>
static __global__ void __launch_bounds__(simt::kThreads, 2) fwd_kernel(MopEdgewiseParams p, Layout L, float* ws_base) {
  __shared__ simt::GemmSmem gs;
  __shared__ float red[32];
  const int G = p.B * p.H;
  const int N = L.N, dk = L.dk;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    Ctx c{p, L, ws_base + (size_t)blockIdx.x * L.total, gs, red, g / p.H, g % p.H, g, 0.f};
    forward_maps<T>(c);
    float* yacc = c.ws + L.o_yacc;
    simt::gemm(yacc, dk, c.map(L.mM), N, 1, c.vb(0), dk, 1, N, dk, N, nullptr, c.ws + L.o_vs1, 1.f, false, gs);
    simt::gemm(yacc, dk, c.F(), N, 1, c.vb(1), dk, 1, N, dk, N, nullptr, c.ws + L.o_vsL, c.w, true, gs);
    T* y = reinterpret_cast<T*>(p.y);
    for (int idx = threadIdx.x; idx < N * dk; idx += simt::kThreads) {
      int n = idx / dk, d = idx % dk;
      y[(((size_t)c.b * N + n) * p.H + c.h) * dk + d] = from_f32<T>(yacc[idx]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------
__device__ void lowrank_head_backward(const Ctx& c, float* dhead) {
  const auto& p = c.p; const auto& L = c.L;
  const int N = L.N, V = L.V, C = L.C, r = L.r;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = simt::kThreads / 32;
  float* a = c.ws + L.o_a; float* bb = c.ws + L.o_b;
  float* da = c.ws + L.o_da; float* db = c.ws + L.o_db;
  float* rho = c.ws + L.o_rho; float* kap = c.ws + L.o_kap;
  float* drho = c.ws + L.o_drho; float* dkap = c.ws + L.o_dkap;
  // da[(t,k), i] = sum_j dG_t[i,j] b[(t,k), j]
  for (int ti = warp; ti < 4 * N; ti += nw) {
    int t = ti / N, i = ti % N;
    const float* dg = c.map(L.mdG + t) + (size_t)i * N;
    float acc[kMaxRank];
    for (int k = 0; k < r; ++k) acc[k] = 0.f;
    for (int j = lane; j < N; j += 32) {
      float d = dg[j];
      for (int k = 0; k < r; ++k) acc[k] = fmaf(d, bb[(size_t)(t * r + k) * N + j], acc[k]);
    }
    for (int k = 0; k < r; ++k) {
      float v = warp_sum(acc[k]);
      if (lane == 0) da[(size_t)(t * r + k) * N + i] = v;
    }
  }
  // db[(t,k), j] = sum_i dG_t[i,j] a[(t,k), i]
  for (int tj = threadIdx.x; tj < 4 * N; tj += simt::kThreads) {
    int t = tj / N, j = tj % N;
    const float* dg = c.map(L.mdG + t);
    float acc[kMaxRank];
    for (int k = 0; k < r; ++k) acc[k] = 0.f;
    for (int i = 0; i < N; ++i) {
      float d = dg[(size_t)i * N + j];
      for (int k = 0; k < r; ++k) acc[k] = fmaf(d, a[(size_t)(t * r + k) * N + i], acc[k]);
    }
    for (int k = 0; k < r; ++k) db[(size_t)(t * r + k) * N + j] = acc[k];
  }
  __syncthreads();
  const float invN = 1.f / (float)N;
  for (int idx = threadIdx.x; idx < C * N; idx += simt::kThreads) {
    int ch = idx / N, i = idx % N;
    float sr = 0.f, sc = 0.f;
    for (int q = 0; q < 4 * r; ++q) {
      sr = fmaf(p.row_w[q * C + ch], da[(size_t)q * N + i], sr);
      sc = fmaf(p.col_w[q * C + ch], db[(size_t)q * N + i], sc);
    }
    drho[idx] = sr * invN; dkap[idx] = sc * invN;
  }
  // parameter partials: [row_w (4r*C), row_b (4r), col_w, col_b]
  const int nW = 4 * r * C;
  for (int idx = threadIdx.x; idx < 2 * (nW + 4 * r); idx += simt::kThreads) {
    int half = idx / (nW + 4 * r), rem = idx % (nW + 4 * r);
    const float* dv = half ? db : da;
    const float* ft = half ? kap : rho;
    float s = 0.f;
    if (rem < nW) {
      int q = rem / C, ch = rem % C;
      for (int i = 0; i < N; ++i) s = fmaf(dv[(size_t)q * N + i], ft[(size_t)ch * N + i], s);
    } else {
      int q = rem - nW;
      for (int i = 0; i < N; ++i) s += dv[(size_t)q * N + i];
    }
    dhead[idx] = s;
  }
  __syncthreads();
}

// Dense head backward.  In: dG maps (grad wrt pre-sigmoid z2).  Out: parameter
// partials, and dz1 written IN PLACE over the z1 maps (dfeat is gathered from
// them by the caller).
__device__ void dense_head_backward(const Ctx& c, float* dhead) {
  const auto& p = c.p; const auto& L = c.L;
  const int N = L.N, C = L.C, hid = L.hid, V = L.V;
  const float eps = p.eps;
  const int nW1 = hid * C, nW3 = L.k3 ? hid * hid * 9 : 0;
  float* dW1 = dhead; float* db1 = dW1 + nW1;
  float* dW3 = db1 + hid; float* db3 = dW3 + nW3;
  float* dW2 = L.k3 ? db3 + hid : db1 + hid; float* db2 = dW2 + 4 * hid;
  // --- conv2 grads: dW2[t,o] = sum_p dz2[t,p] h2[o,p]
  for (int idx = threadIdx.x; idx < 4 * hid + 4; idx += simt::kThreads) {
    float s = 0.f;
    if (idx < 4 * hid) {
      int t = idx / hid, o = idx % hid;
      const float* dz = c.map(L.mdG + t);
      if (L.k3) {
        const float* h2 = c.map(L.mH2 + o);
        for (size_t q = 0; q < L.N2; ++q) s = fmaf(dz[q], h2[q], s);
      } else {
        const float* z1 = c.map(L.mZ1 + o);
        for (size_t q = 0; q < L.N2; ++q) s = fmaf(dz[q], gelu_tanh_(z1[q]), s);
      }
      dW2[idx] = s;
    } else {
      const float* dz = c.map(L.mdG + (idx - 4 * hid));
      for (size_t q = 0; q < L.N2; ++q) s += dz[q];
      db2[idx - 4 * hid] = s;
    }
  }
  __syncthreads();
  if (L.k3) {
    // dh2[o,p] = sum_t W2[t,o] dz2[t,p]  (stored: needed at 3x3 neighbours)
    for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) {
      float dz[4];
      for (int t = 0; t < 4; ++t) dz[t] = c.map(L.mdG + t)[idx];
      for (int o = 0; o < hid; ++o) {
        float s = 0.f;
        for (int t = 0; t < 4; ++t) s = fmaf(p.conv2_w[t * hid + o], dz[t], s);
        c.map(L.mDH2 + o)[idx] = s;
      }
    }
    __syncthreads();
    // dW3[o,ch,u,v] = sum_p dh2[o,p] * X2[ch, p + (u-1,v-1)];  db3[o] = sum_p dh2[o,p].  One warp per (o, ch) pair, lanes
    // stride over the pixels, nine accumulators (the first version recomputed gelu(gelu(z1)) per output and pixel: 330 k
    // instructions per output element)
    {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = simt::kThreads / 32;
      for (int pair = warp; pair < hid * hid + hid; pair += nw) {
        if (pair < hid * hid) {
          const int o = pair / hid, ch = pair % hid;
          const float* dh = c.map(L.mDH2 + o);
          const float* x2 = c.map(L.mX2 + ch);
          float acc[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) acc[t] = 0.f;
          for (int q = lane; q < N * N; q += 32) {
            const int i = q / N, j = q % N;
            const float d = dh[q];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
              const int ii = i + u - 1;
              if (ii < 0 || ii >= N) continue;
#pragma unroll
              for (int v = 0; v < 3; ++v) {
                const int jj = j + v - 1;
                if (jj < 0 || jj >= N) continue;
                acc[u * 3 + v] = fmaf(d, x2[(size_t)ii * N + jj], acc[u * 3 + v]);
              }
            }
          }
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const float sv = warp_sum(acc[t]);
            if (lane == 0) dW3[(size_t)pair * 9 + t] = sv;
          }
        } else {
          const float* dh = c.map(L.mDH2 + (pair - hid * hid));
          float sv = 0.f;
          for (int q = lane; q < N * N; q += 32) sv += dh[q];
          sv = warp_sum(sv);
          if (lane == 0) db3[pair - hid * hid] = sv;
        }
      }
    }
    __syncthreads();
  }
  // --- dz1 in place.  dh1[ch,p] = k3 ? gelu'(h1) * sum_{o,u,v} W3[o,ch,u,v] dh2[o, p-(u-1,v-1)] : dh2
  for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) {
    int i = idx / N, j = idx % N;
    float dh1[kMaxHidden];
    if (L.k3) {
      for (int ch = 0; ch < hid; ++ch) dh1[ch] = 0.f;
      for (int u = 0; u < 3; ++u) {
        int ii = i - (u - 1);
        if (ii < 0 || ii >= N) continue;
        for (int v = 0; v < 3; ++v) {
          int jj = j - (v - 1);
          if (jj < 0 || jj >= N) continue;
          size_t q = (size_t)ii * N + jj;
          for (int o = 0; o < hid; ++o) {
            float d = c.map(L.mDH2 + o)[q];
            for (int ch = 0; ch < hid; ++ch) dh1[ch] = fmaf(p.mid3_w[((o * hid + ch) * 3 + u) * 3 + v], d, dh1[ch]);
          }
        }
      }
    } else {
      float dz[4];
      for (int t = 0; t < 4; ++t) dz[t] = c.map(L.mdG + t)[idx];
      for (int o = 0; o < hid; ++o) {
        float s = 0.f;
        for (int t = 0; t < 4; ++t) s = fmaf(p.conv2_w[t * hid + o], dz[t], s);
        dh1[o] = s;
      }
    }
    for (int ch = 0; ch < hid; ++ch) {
      float z = c.map(L.mZ1 + ch)[idx];
      float g = dh1[ch];
      if (L.k3) g *= dgelu_tanh_(gelu_tanh_(z));
      c.map(L.mZ1 + ch)[idx] = g * dgelu_tanh_(z);
    }
  }
  __syncthreads();
  // --- conv1 grads: dW1[o,ch] = sum_p dz1[o,p] feat[ch,p]
  const float* Fm = c.F(); const float* Rm = c.Rf();
  for (int idx = threadIdx.x; idx < nW1 + hid; idx += simt::kThreads) {
    float s = 0.f;
    if (idx < nW1) {
      int o = idx / C, ch = idx % C;
      const float* dz = c.map(L.mZ1 + o);
      if (ch < V) {
        const float* f = c.S(ch);
        for (size_t q = 0; q < L.N2; ++q) s = fmaf(dz[q], f[q], s);
      } else if (ch < 2 * V) {
        const float* f = c.S(ch - V);
        for (int i = 0; i < N; ++i)
          for (int j = 0; j < N; ++j) s = fmaf(dz[(size_t)i * N + j], f[(size_t)j * N + i], s);
      } else if (ch < 2 * V + 2) {
        const float* f = (ch == 2 * V) ? Fm : Rm;
        for (size_t q = 0; q < L.N2; ++q) s = fmaf(dz[q], logf(f[q] + eps), s);
      } else {
        const float* f = c.map(L.mL + ch - (2 * V + 2));
        for (size_t q = 0; q < L.N2; ++q) s = fmaf(dz[q], f[q], s);
      }
      dW1[idx] = s;
    } else {
      const float* dz = c.map(L.mZ1 + (idx - nW1));
      for (size_t q = 0; q < L.N2; ++q) s += dz[q];
      db1[idx - nW1] = s;
    }
  }
  __syncthreads();
}

// grad of feature channel `ch` at pixel (i,j) (flat idx) for either head
__device__ __forceinline__ float dfeat_at(const Ctx& c, int ch, int i, int j) {
  const auto& L = c.L;
  if (L.cgate) return 0.f;
  if (L.dense) {
    size_t q = (size_t)i * L.N + j;
    float s = 0.f;
    for (int o = 0; o < L.hid; ++o) s = fmaf(c.p.conv1_w[o * L.C + ch], c.map(L.mZ1 + o)[q], s);
    return s;
  }
  return c.ws[L.o_drho + (size_t)ch * L.N + i] + c.ws[L.o_dkap + (size_t)ch * L.N + j];
}

template <typename T>
static __global__ void __launch_bounds__(simt::kThreads, 2) bwd_kernel(MopEdgewiseParams p, Layout L, float* ws_base) {
  __shared__ simt::GemmSmem gs;
  __shared__ float red[32];
  const int G = p.B * p.H;
  const int N = L.N, dk = L.dk, V = L.V, H = p.H;
  const float s = rsqrtf((float)dk);
  const float bn = p.beta_not / (float)max(1, V - 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = simt::kThreads / 32;
  const size_t nhead = head_param_count(p.gate_mode, V, L.r, L.hid, L.k3, L.nl);
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    Ctx c{p, L, ws_base + (size_t)blockIdx.x * L.total, gs, red, g / H, g % H, g, 0.f};
    forward_maps<T>(c);
    const float w = c.w;
    float* M = c.map(L.mM);
    float* D = c.map(L.mD);
    float* X0 = c.map(L.mX0);
    float* X1 = c.map(L.mX1);
    float* dyf = c.ws + L.o_dyf; float* dv1 = c.ws + L.o_dv1; float* dvl = c.ws + L.o_dvl;
    float* tt = c.ws + L.o_tt;
    const float* vs1 = c.ws + L.o_vs1; const float* vsL = c.ws + L.o_vsL;
    const T* dy = reinterpret_cast<const T*>(p.dy);
    for (int idx = threadIdx.x; idx < N * dk; idx += simt::kThreads) {
      int n = idx / dk, d = idx % dk;
      dyf[idx] = to_f32<T>(dy[(((size_t)c.b * N + n) * H + c.h) * dk + d]);
    }
    __syncthreads();
    // dA = dY V1^T ; D = A (.) (dA - rowsum(dA (.) A))
    simt::gemm(D, N, dyf, dk, 1, c.vb(0), 1, dk, N, N, dk, vs1, nullptr, 1.f, false, gs);
    simt::softmax_bwd_rows(D, M, N, N);
    simt::gemm(dv1, dk, M, 1, N, dyf, dk, 1, N, dk, N, nullptr, nullptr, 1.f, false, gs);       // A^T dY
    simt::gemm(dvl, dk, c.F(), 1, N, dyf, dk, 1, N, dk, N, nullptr, nullptr, w, false, gs);     // w F^T dY
    simt::gemm(X0, N, dyf, dk, 1, c.vb(1), 1, dk, N, N, dk, vsL, nullptr, w, false, gs);        // dF = w dY VL^T
    {
      float part = 0.f;
      const float* Fm = c.F();
      for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) part = fmaf(Fm[idx], X0[idx], part);
      float tot = simt::block_sum(part, red);
      if (threadIdx.x == 0) p.dlogit_part[g] = (1.f - w) * tot;
    }
    // gate pre-activation grads dG_t
    if (!L.cgate) {
      const float* Fm = c.F();
      for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) {
        float sum = 0.f, mx = -INFINITY, s0 = c.S(0)[idx];
        for (int v = 0; v < V; ++v) { float x = c.S(v)[idx]; sum += x; mx = fmaxf(mx, x); }
        float se = 0.f;
        for (int v = 0; v < V; ++v) se += expf(c.S(v)[idx] - mx);
        float U = sum - s0, O = mx + logf(se) - s0, lf = logf(Fm[idx] + p.eps), d = D[idx];
        float dg[4] = {d * U, d * O, -bn * d * U, d * lf};
        for (int t = 0; t < 4; ++t) { float gt = c.Gm(t)[idx]; c.map(L.mdG + t)[idx] = dg[t] * gt * (1.f - gt); }
      }
      __syncthreads();
    }
    float* dhead = p.dhead_part + (size_t)g * nhead;
    if (!L.cgate) { if (L.dense) dense_head_backward(c, dhead); else lowrank_head_backward(c, dhead); }
    // dF += (D g_chain + dfeat_{2V}) / (F + eps)
    {
      const float* Fm = c.F();
      for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) {
        int i = idx / N, j = idx % N;
        float dlf = D[idx] * c.Gm(3)[idx] + dfeat_at(c, 2 * V, i, j);
        X0[idx] += dlf / (Fm[idx] + p.eps);
      }
      __syncthreads();
    }
    // F chain: F = A_{cv(0)} .. A_{cv(hops-1)}   (a view that occupies several chain positions accumulates)
    {
      float* X = X0; float* Xn = X1;
      unsigned written = 0;
      for (int k = L.hops - 1; k >= 1; --k) {
        const int vk = c.cv(k);
        simt::gemm(c.map(L.mdA + vk), N, c.P(k - 1), 1, N, X, N, 1, N, N, N, nullptr, nullptr, 1.f, (written >> vk) & 1u, gs);
        written |= 1u << vk;
        simt::gemm(Xn, N, X, N, 1, c.A(vk), 1, N, N, N, N, nullptr, nullptr, 1.f, false, gs);  // X A_k^T
        float* t = X; X = Xn; Xn = t;
      }
      float* dA0 = c.map(L.mdA + 0);
      for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) dA0[idx] = ((written & 1u) ? dA0[idx] : 0.f) + X[idx];
      __syncthreads();
    }
    // R chain: R = A_V .. A_1 ; dR = dfeat_{2V+1} / (R + eps)
    if (!L.cgate) {
      const float* Rm = c.Rf();
      for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) {
        int i = idx / N, j = idx % N;
        X0[idx] = dfeat_at(c, 2 * V + 1, i, j) / (Rm[idx] + p.eps);
      }
      __syncthreads();
      float* X = X0; float* Xn = X1;
      for (int k = 0; k <= V - 2; ++k) {
        simt::gemm(c.map(L.mdA + k), N, c.R(k + 1), 1, N, X, N, 1, N, N, N, nullptr, nullptr, 1.f, true, gs);
        simt::gemm(Xn, N, X, N, 1, c.A(k), 1, N, N, N, N, nullptr, nullptr, 1.f, false, gs);
        float* t = X; X = Xn; Xn = t;
      }
      float* dAl = c.map(L.mdA + V - 1);
      for (size_t idx = threadIdx.x; idx < L.N2; idx += simt::kThreads) dAl[idx] += X[idx];
      __syncthreads();
    }
    // S lens bank: weight gradients dW[l,v,u,x] = sum_p dfeat_ch[p] S_v[p + (u-1,x-1) d_l]
    if (L.nl) {
      float* dl = p.dlens_part + (size_t)g * V * L.nl * 9;
      for (int idx = threadIdx.x; idx < V * L.nl * 9; idx += simt::kThreads) {
        const int x = idx % 3, u = (idx / 3) % 3, ch = idx / 9, l = ch / V, v = ch % V, d = p.lens_dil[l];
        const float* S = c.S(v);
        float acc = 0.f;
        for (int i = 0; i < N; ++i) {
          const int ii = i + (u - 1) * d;
          if (ii < 0 || ii >= N) continue;
          for (int j = 0; j < N; ++j) {
            const int jj = j + (x - 1) * d;
            if (jj < 0 || jj >= N) continue;
            acc = fmaf(dfeat_at(c, 2 * V + 2 + ch, i, j), S[(size_t)ii * N + jj], acc);
          }
        }
        dl[idx] = acc;
      }
    }
    // dS_k (in place over dA_k): softmax backward + direct mix terms + feature terms
    for (int i = warp; i < N; i += nw) {
      float rs[kMaxViews];
      for (int k = 0; k < V; ++k) {
        const float* da = c.map(L.mdA + k) + (size_t)i * N;
        const float* ak = c.A(k) + (size_t)i * N;
        float dot = 0.f;
        for (int j = lane; j < N; j += 32) dot = fmaf(da[j], ak[j], dot);
        rs[k] = warp_sum(dot);
      }
      for (int j = lane; j < N; j += 32) {
        size_t idx = (size_t)i * N + j;
        float sv[kMaxViews];
        float mx = -INFINITY;
        for (int v = 0; v < V; ++v) { sv[v] = c.S(v)[idx]; mx = fmaxf(mx, sv[v]); }
        float se = 0.f;
        for (int v = 0; v < V; ++v) se += expf(sv[v] - mx);
        float lse = mx + logf(se);
        float d = D[idx], g_and = c.Gm(0)[idx], g_or = c.Gm(1)[idx], g_not = c.Gm(2)[idx];
        for (int k = 0; k < V; ++k) {
          float pi = expf(sv[k] - lse);
          float direct = (k == 0) ? d * (1.f + g_or * (pi - 1.f)) : d * (g_and + g_or * pi - g_not * bn);
          float* da = c.map(L.mdA + k);
          float v = c.A(k)[idx] * (da[idx] - rs[k]) + direct + dfeat_at(c, k, i, j) + dfeat_at(c, V + k, j, i);
          for (int l = 0; l < L.nl; ++l) {   // through the lens convolutions: S_k[i,j] feeds pixel (i,j) - (u-1,x-1) d_l of channel l*V+k
            const int d = p.lens_dil[l], ch = 2 * V + 2 + l * V + k;
            const float* wt = p.lens_w + (size_t)(l * V + k) * 9;
            for (int u = 0; u < 3; ++u) {
              const int ii = i - (u - 1) * d;
              if (ii < 0 || ii >= N) continue;
              for (int x = 0; x < 3; ++x) {
                const int jj = j - (x - 1) * d;
                if (jj < 0 || jj >= N) continue;
                v = fmaf(wt[u * 3 + x], dfeat_at(c, ch, ii, jj), v);
              }
            }
          }
          da[idx] = v;
        }
      }
    }
    __syncthreads();
    // projections
    float* zk = c.ws + L.o_z;
    for (int k = 0; k < V; ++k) {
      float* dSk = c.map(L.mdA + k);
      const int vi = (L.Vp == 1) ? 0 : k;
      float* dqa = c.ws + L.o_dqa + (size_t)vi * L.Nd;
      float* dka = c.ws + L.o_dka + (size_t)vi * L.Nd;
      const bool accv = (L.Vp == 1) && k > 0;
      // dK_k = s dS_k^T Qb (.) c_k
      simt::gemm(dka, dk, dSk, 1, N, c.qb(k), dk, 1, N, dk, N, nullptr, c.cvec(k), s, accv, gs);
      // T = s dS_k Kb (unscaled): needed both for dQ and for the scale grads
      simt::gemm(tt, dk, dSk, N, 1, c.kb(k), dk, 1, N, dk, N, nullptr, nullptr, s, false, gs);
      const float* cv = c.cvec(k);
      for (int idx = threadIdx.x; idx < N * dk; idx += simt::kThreads) {
        float v = tt[idx] * cv[idx % dk];
        dqa[idx] = accv ? dqa[idx] + v : v;
      }
      const float* qbk = c.qb(k);
      for (int d = threadIdx.x; d < dk; d += simt::kThreads) {
        float z = 0.f;
        for (int n = 0; n < N; ++n) z = fmaf(tt[(size_t)n * dk + d], qbk[(size_t)n * dk + d], z);
        zk[(size_t)k * dk + d] = z;
      }
      __syncthreads();
    }
    // write dqkv and the per-(b,h) scale partials
    T* dqkv = reinterpret_cast<T*>(p.dqkv);
    const size_t hd = (size_t)H * dk;
    for (int vi = 0; vi < L.Vp; ++vi) {
      const float* dqa = c.ws + L.o_dqa + (size_t)vi * L.Nd;
      const float* dka = c.ws + L.o_dka + (size_t)vi * L.Nd;
      for (int idx = threadIdx.x; idx < N * dk; idx += simt::kThreads) {
        int n = idx / dk, d = idx % dk;
        size_t base = ((((size_t)c.b * N + n) * L.Vp + vi) * 3) * hd + (size_t)c.h * dk + d;
        float dvv;
        if (L.Vp == 1) dvv = dv1[idx] * vs1[d] + dvl[idx] * vsL[d];
        else dvv = (vi == 0 ? dv1[idx] * vs1[d] : 0.f) + (vi == V - 1 ? dvl[idx] * vsL[d] : 0.f);
        dqkv[base] = from_f32<T>(dqa[idx]);
        dqkv[base + hd] = from_f32<T>(dka[idx]);
        dqkv[base + 2 * hd] = from_f32<T>(dvv);
      }
    }
    if (p.dscale_part) {
      float* ds = p.dscale_part + (size_t)g * 3 * V * dk;
      for (int idx = threadIdx.x; idx < V * dk; idx += simt::kThreads) {
        int k = idx / dk, d = idx % dk;
        size_t pi = ((size_t)k * H + c.h) * dk + d;
        ds[idx] = p.k_scale[pi] * zk[idx];
        ds[(size_t)V * dk + idx] = p.q_scale[pi] * zk[idx];
        float dvs = 0.f;
        if (k == 0 || k == V - 1) {
          const float* dv = (k == 0) ? dv1 : dvl;
          const float* vbb = c.vb(k == 0 ? 0 : 1);
          for (int n = 0; n < N; ++n) dvs = fmaf(dv[(size_t)n * dk + d], vbb[(size_t)n * dk + d], dvs);
        }
        ds[(size_t)2 * V * dk + idx] = dvs;
      }
    }
    __syncthreads();
  }
}

}  // namespace ew
}  // namespace mop
