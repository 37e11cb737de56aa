// Multi-scale deformable attention forward for sm_100a.
//
// Replaces ms_deformable_im2col_gpu_kernel (reference ops/src/cuda/ms_deform_im2col_cuda.cuh:242-304,
// bilinear helper :38-89, launcher :928-959).  The op is a gather: per (query, head) 12 samples x 4 corners
// x 128 B head vectors; it is bound by L2/HBM bandwidth, not FLOPs, so the design is:
//   * 8 lanes x float4 cover the 32 channels of a head  -> every corner is one 128-bit load per lane and
//     one fully-used 128 B line per (sample corner);
//   * all 48 corner loads of a thread are unconditional (clamped index, zeroed weight) so they are issued
//     back-to-back (memory-level parallelism instead of the reference's dependent branchy loads);
//   * offsets / logits are read as 128-bit broadcast loads; softmax over L*P and the sampling-location
//     arithmetic of MSDeformAttn.forward (ops/modules/ms_deform_attn.py:101-108) are fused in the
//     encoder variant, so loc/weight tensors are never materialised.
#include "rowwise.cuh"

namespace univs {

constexpr int kMaxLevels = 8;
struct LevelTable {
  int H[kMaxLevels];
  int W[kMaxLevels];
  int start[kMaxLevels];
};

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x = fmaf(w, v.x, acc.x);
  acc.y = fmaf(w, v.y, acc.y);
  acc.z = fmaf(w, v.z, acc.z);
  acc.w = fmaf(w, v.w, acc.w);
}

// One bilinear sample of a 4-channel slice.  `base` points at channel slice of pixel (0,0) of the level;
// pix_stride = M*D floats.  Semantics of ms_deform_attn_im2col_bilinear (:38-89) + the in-range test (:293).
__device__ __forceinline__ void sample4(float4& acc, const float* __restrict__ base, int pix_stride, int H, int W,
                                        float x, float y, float aw) {
  const bool inside = (y > -1.f) && (x > -1.f) && (y < (float)H) && (x < (float)W);
  const float xf = floorf(x), yf = floorf(y);
  const int x0 = (int)xf, y0 = (int)yf;
  const float lx = x - xf, ly = y - yf;
  const float hx = 1.f - lx, hy = 1.f - ly;
  const bool x0ok = x0 >= 0, x1ok = x0 + 1 <= W - 1, y0ok = y0 >= 0, y1ok = y0 + 1 <= H - 1;
  const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
  const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
  const float w1 = (inside && y0ok && x0ok) ? hy * hx : 0.f;
  const float w2 = (inside && y0ok && x1ok) ? hy * lx : 0.f;
  const float w3 = (inside && y1ok && x0ok) ? ly * hx : 0.f;
  const float w4 = (inside && y1ok && x1ok) ? ly * lx : 0.f;
  const float4 v1 = ldg_f4(base + (size_t)(yc0 * W + xc0) * pix_stride);
  const float4 v2 = ldg_f4(base + (size_t)(yc0 * W + xc1) * pix_stride);
  const float4 v3 = ldg_f4(base + (size_t)(yc1 * W + xc0) * pix_stride);
  const float4 v4 = ldg_f4(base + (size_t)(yc1 * W + xc1) * pix_stride);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  fma4(s, w1, v1);
  fma4(s, w2, v2);
  fma4(s, w3, v3);
  fma4(s, w4, v4);
  fma4(acc, aw, s);
}

__device__ __forceinline__ float sample1(const float* __restrict__ base, int pix_stride, int H, int W, float x,
                                         float y) {
  if (!((y > -1.f) && (x > -1.f) && (y < (float)H) && (x < (float)W))) return 0.f;
  const float xf = floorf(x), yf = floorf(y);
  const int x0 = (int)xf, y0 = (int)yf;
  const float lx = x - xf, ly = y - yf, hx = 1.f - lx, hy = 1.f - ly;
  float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
  if (y0 >= 0 && x0 >= 0) v1 = __ldg(base + (size_t)(y0 * W + x0) * pix_stride);
  if (y0 >= 0 && x0 + 1 <= W - 1) v2 = __ldg(base + (size_t)(y0 * W + x0 + 1) * pix_stride);
  if (y0 + 1 <= H - 1 && x0 >= 0) v3 = __ldg(base + (size_t)((y0 + 1) * W + x0) * pix_stride);
  if (y0 + 1 <= H - 1 && x0 + 1 <= W - 1) v4 = __ldg(base + (size_t)((y0 + 1) * W + x0 + 1) * pix_stride);
  return hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
}

// Level tables of the reference ABI are DEVICE int64 tensors (ms_deform_attn_cuda.cu:25-32): when the caller hands device
// pointers the kernels read them here, so the entry point stays stream-ordered with no host synchronisation (and can be
// captured in a CUDA graph); host tables arrive by value in `lt`.
__device__ __forceinline__ void device_levels(LevelTable& lt, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi, int L) {
  if (shapes == nullptr) return;
  for (int l = 0; l < L; ++l) {
    lt.H[l] = (int)__ldg(shapes + 2 * l);
    lt.W[l] = (int)__ldg(shapes + 2 * l + 1);
    lt.start[l] = (int)__ldg(lsi + l);
  }
}

// ---- reference-ABI kernel: explicit sampling_loc / attn_weight tensors, any L, P; D % 4 == 0 ------------
__global__ void __launch_bounds__(256)
msda_generic_vec4_kernel(const float* __restrict__ value, LevelTable lt, const int64_t* __restrict__ dev_shapes,
                         const int64_t* __restrict__ dev_lsi, const float* __restrict__ loc,
                         const float* __restrict__ aw, int N, int S, int M, int D, int L, int Lq, int P,
                         float* __restrict__ out) {
  device_levels(lt, dev_shapes, dev_lsi, L);
  const int groups = D >> 2;
  const long long total = (long long)N * Lq * M * groups;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % groups);
    const long long pair = idx / groups;  // (n*Lq + q)*M + m
    const int m = (int)(pair % M);
    const int n = (int)(pair / ((long long)M * Lq));
    const float* lp = loc + pair * (L * P * 2);
    const float* wp = aw + pair * (L * P);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int pix_stride = M * D;
    for (int l = 0; l < L; ++l) {
      const int H = lt.H[l], W = lt.W[l];
      const float* base = value + ((size_t)n * S + lt.start[l]) * pix_stride + m * D + cg * 4;
      for (int p = 0; p < P; ++p) {
        const float lx = __ldg(lp + (l * P + p) * 2), ly = __ldg(lp + (l * P + p) * 2 + 1);
        const float w = __ldg(wp + l * P + p);
        sample4(acc, base, pix_stride, H, W, lx * W - 0.5f, ly * H - 0.5f, w);
      }
    }
    *reinterpret_cast<float4*>(out + pair * D + cg * 4) = acc;
  }
}

__global__ void __launch_bounds__(256)
msda_generic_scalar_kernel(const float* __restrict__ value, LevelTable lt, const int64_t* __restrict__ dev_shapes,
                           const int64_t* __restrict__ dev_lsi, const float* __restrict__ loc,
                           const float* __restrict__ aw, int N, int S, int M, int D, int L, int Lq, int P,
                           float* __restrict__ out) {
  device_levels(lt, dev_shapes, dev_lsi, L);
  const long long total = (long long)N * Lq * M * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const long long pair = idx / D;
    const int m = (int)(pair % M);
    const int n = (int)(pair / ((long long)M * Lq));
    const float* lp = loc + pair * (L * P * 2);
    const float* wp = aw + pair * (L * P);
    float acc = 0.f;
    const int pix_stride = M * D;
    for (int l = 0; l < L; ++l) {
      const int H = lt.H[l], W = lt.W[l];
      const float* base = value + ((size_t)n * S + lt.start[l]) * pix_stride + m * D + c;
      for (int p = 0; p < P; ++p) {
        const float lx = __ldg(lp + (l * P + p) * 2), ly = __ldg(lp + (l * P + p) * 2 + 1);
        acc += __ldg(wp + l * P + p) * sample1(base, pix_stride, H, W, lx * W - 0.5f, ly * H - 0.5f);
      }
    }
    out[idx] = acc;
  }
}

// ---- encoder kernel: D = 32, queries = pyramid pixels, softmax + location arithmetic fused ---------------
// 8 lanes per (n, q, m); a 256-thread CTA covers 32 consecutive (q, m) pairs = 4 queries x 8 heads.
template <int L, int P>
__global__ void __launch_bounds__(256)
msda_encoder_kernel(const float* __restrict__ value, LevelTable lt, const float* __restrict__ ol, int N, int S,
                    int M, float* __restrict__ out) {
  constexpr int LP = L * P;
  static_assert((LP * 2) % 4 == 0 && LP % 4 == 0, "vector loads need L*P % 4 == 0");
  const int lane8 = threadIdx.x & 7;
  const long long pair = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  if (pair >= (long long)N * S * M) return;
  const int m = (int)(pair % M);
  const long long nq = pair / M;
  const int q = (int)(nq % S);
  const int n = (int)(nq / S);

  int lq = 0;
#pragma unroll
  for (int l = 1; l < L; ++l)
    if (q >= lt.start[l]) lq = l;
  const int rel = q - lt.start[lq];
  const int qy = rel / lt.W[lq], qx = rel - qy * lt.W[lq];
  const float refx = ((float)qx + 0.5f) / (float)lt.W[lq];
  const float refy = ((float)qy + 0.5f) / (float)lt.H[lq];

  const float* row = ol + nq * (long long)(M * LP * 3);
  float off[LP * 2], lg[LP];
  {
    const float4* po = reinterpret_cast<const float4*>(row + m * (LP * 2));
#pragma unroll
    for (int i = 0; i < LP * 2 / 4; ++i) {
      const float4 t = __ldg(po + i);
      off[4 * i] = t.x; off[4 * i + 1] = t.y; off[4 * i + 2] = t.z; off[4 * i + 3] = t.w;
    }
    const float4* pl = reinterpret_cast<const float4*>(row + M * LP * 2 + m * LP);
#pragma unroll
    for (int i = 0; i < LP / 4; ++i) {
      const float4 t = __ldg(pl + i);
      lg[4 * i] = t.x; lg[4 * i + 1] = t.y; lg[4 * i + 2] = t.z; lg[4 * i + 3] = t.w;
    }
  }
  float mx = lg[0];
#pragma unroll
  for (int i = 1; i < LP; ++i) mx = fmaxf(mx, lg[i]);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LP; ++i) {
    lg[i] = expf(lg[i] - mx);
    sum += lg[i];
  }
  const float inv = 1.f / sum;

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int pix_stride = M * 32;
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int H = lt.H[l], W = lt.W[l];
    const float* base = value + ((size_t)n * S + lt.start[l]) * pix_stride + m * 32 + lane8 * 4;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float locx = refx + off[(l * P + p) * 2] / (float)W;
      const float locy = refy + off[(l * P + p) * 2 + 1] / (float)H;
      sample4(acc, base, pix_stride, H, W, locx * (float)W - 0.5f, locy * (float)H - 0.5f, lg[l * P + p] * inv);
    }
  }
  *reinterpret_cast<float4*>(out + pair * 32 + lane8 * 4) = acc;
}

// The work of one (frame n, query q, head m) triple for one 8-lane group (lane8 owns channels [4*lane8, 4*lane8+4)):
// a verbatim restatement of the body of msda_encoder_kernel above, which is left untouched (validated SASS).
// FUSED adds what surrounds the operator in the encoder layer (ops/modules/ms_deform_attn.py:98-120): the biases of the
// value_proj / sampling_offsets / attention_weights linears (so those GEMMs run bias-free, without the broadcast copy a
// library GEMM needs for a bias), and emission of the result in the GEMM operand format of output_proj.
//   value bias: sum_i w_i (v_i + b) = sum_i w_i v_i + b * sum_i w_i over the IN-BOUNDS corners i (out-of-range corners
//   read zero in the reference, not b: ms_deform_im2col_cuda.cuh:38-89), so b is scaled by the surviving corner weights.
__device__ __forceinline__ void sample4_biased(float4& acc, const float* __restrict__ base, int pix_stride, int H, int W,
                                               float x, float y, float aw, const float4& vb) {
  const bool inside = (y > -1.f) && (x > -1.f) && (y < (float)H) && (x < (float)W);
  const float xf = floorf(x), yf = floorf(y);
  const int x0 = (int)xf, y0 = (int)yf;
  const float lx = x - xf, ly = y - yf;
  const float hx = 1.f - lx, hy = 1.f - ly;
  const bool x0ok = x0 >= 0, x1ok = x0 + 1 <= W - 1, y0ok = y0 >= 0, y1ok = y0 + 1 <= H - 1;
  const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
  const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
  const float w1 = (inside && y0ok && x0ok) ? hy * hx : 0.f;
  const float w2 = (inside && y0ok && x1ok) ? hy * lx : 0.f;
  const float w3 = (inside && y1ok && x0ok) ? ly * hx : 0.f;
  const float w4 = (inside && y1ok && x1ok) ? ly * lx : 0.f;
  const float4 v1 = ldg_f4(base + (size_t)(yc0 * W + xc0) * pix_stride);
  const float4 v2 = ldg_f4(base + (size_t)(yc0 * W + xc1) * pix_stride);
  const float4 v3 = ldg_f4(base + (size_t)(yc1 * W + xc0) * pix_stride);
  const float4 v4 = ldg_f4(base + (size_t)(yc1 * W + xc1) * pix_stride);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  fma4(s, w1, v1);
  fma4(s, w2, v2);
  fma4(s, w3, v3);
  fma4(s, w4, v4);
  fma4(s, (w1 + w2) + (w3 + w4), vb);
  fma4(acc, aw, s);
}

template <int L, int P, bool FUSED>
__device__ __forceinline__ void msda_encoder_pair(const float* __restrict__ value, const LevelTable& lt,
                                                  const float* __restrict__ ol, int n, int q, int m, int lane8, int S,
                                                  int M, float* __restrict__ out, const float* __restrict__ value_bias,
                                                  const float* __restrict__ ol_bias, int split) {
  constexpr int LP = L * P;
  static_assert((LP * 2) % 4 == 0 && LP % 4 == 0, "vector loads need L*P % 4 == 0");
  const long long nq = (long long)n * S + q;
  const long long pair = nq * M + m;

  int lq = 0;
#pragma unroll
  for (int l = 1; l < L; ++l)
    if (q >= lt.start[l]) lq = l;
  const int rel = q - lt.start[lq];
  const int qy = rel / lt.W[lq], qx = rel - qy * lt.W[lq];
  const float refx = ((float)qx + 0.5f) / (float)lt.W[lq];
  const float refy = ((float)qy + 0.5f) / (float)lt.H[lq];

  const float* row = ol + nq * (long long)(M * LP * 3);
  float off[LP * 2], lg[LP];
  {
    const float4* po = reinterpret_cast<const float4*>(row + m * (LP * 2));
#pragma unroll
    for (int i = 0; i < LP * 2 / 4; ++i) {
      const float4 t = __ldg(po + i);
      off[4 * i] = t.x; off[4 * i + 1] = t.y; off[4 * i + 2] = t.z; off[4 * i + 3] = t.w;
    }
    const float4* pl = reinterpret_cast<const float4*>(row + M * LP * 2 + m * LP);
#pragma unroll
    for (int i = 0; i < LP / 4; ++i) {
      const float4 t = __ldg(pl + i);
      lg[4 * i] = t.x; lg[4 * i + 1] = t.y; lg[4 * i + 2] = t.z; lg[4 * i + 3] = t.w;
    }
  }
  if (FUSED && ol_bias != nullptr) {
#pragma unroll
    for (int i = 0; i < LP * 2; ++i) off[i] += __ldg(ol_bias + m * (LP * 2) + i);
#pragma unroll
    for (int i = 0; i < LP; ++i) lg[i] += __ldg(ol_bias + M * LP * 2 + m * LP + i);
  }
  float mx = lg[0];
#pragma unroll
  for (int i = 1; i < LP; ++i) mx = fmaxf(mx, lg[i]);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LP; ++i) {
    lg[i] = expf(lg[i] - mx);
    sum += lg[i];
  }
  const float inv = 1.f / sum;

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int pix_stride = M * 32;
  float4 vb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (FUSED && value_bias != nullptr) vb = ldg_f4(value_bias + m * 32 + lane8 * 4);
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int H = lt.H[l], W = lt.W[l];
    const float* base = value + ((size_t)n * S + lt.start[l]) * pix_stride + m * 32 + lane8 * 4;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float locx = refx + off[(l * P + p) * 2] / (float)W;
      const float locy = refy + off[(l * P + p) * 2 + 1] / (float)H;
      if (FUSED)
        sample4_biased(acc, base, pix_stride, H, W, locx * (float)W - 0.5f, locy * (float)H - 0.5f, lg[l * P + p] * inv, vb);
      else
        sample4(acc, base, pix_stride, H, W, locx * (float)W - 0.5f, locy * (float)H - 0.5f, lg[l * P + p] * inv);
    }
  }
  if (FUSED)
    store_maybe_split(out, (size_t)nq, M * 32, m * 32 + lane8 * 4, acc, split);
  else
    *reinterpret_cast<float4*>(out + pair * 32 + lane8 * 4) = acc;
}

// Tiled variant (opt-in, results bit-identical to msda_encoder_kernel: same per-(n,q,m) arithmetic): a CTA covers a
// tile_w x (32 / tile_w) block of neighbouring queries of ONE head instead of 4 consecutive queries x 8 heads.  The
// sampling footprints of neighbouring queries overlap, and all 32 queries of the CTA now touch the same 128-byte head
// slice of each pixel, so the gathered lines are reused from L1 instead of being re-fetched from L2 (the ncu capture of
// the untiled kernel: L2->L1 traffic 8x the size of the value tensor, L1 hit rate 62 %).
struct TileTable {
  int first[kMaxLevels + 1];   // first tile index of each level (+ total)
  int tiles_x[kMaxLevels];
};

template <int L, int P, bool FUSED>
__global__ void __launch_bounds__(256)
msda_encoder_tiled_kernel(const float* __restrict__ value, LevelTable lt, TileTable tt, const float* __restrict__ ol,
                          int S, int M, int tile_w_log2, float* __restrict__ out, const float* __restrict__ value_bias,
                          const float* __restrict__ ol_bias, int split) {
  const int lane8 = threadIdx.x & 7;
  const int tq = threadIdx.x >> 3;                 // query inside the tile
  const int tile = blockIdx.x, m = blockIdx.y, n = blockIdx.z;
  int l = 0;
#pragma unroll
  for (int i = 1; i < L; ++i)
    if (tile >= tt.first[i]) l = i;
  const int t = tile - tt.first[l];
  const int ty = t / tt.tiles_x[l], tx = t - ty * tt.tiles_x[l];
  const int tile_w = 1 << tile_w_log2, tile_h = 32 >> tile_w_log2;
  const int qx = tx * tile_w + (tq & (tile_w - 1));
  const int qy = ty * tile_h + (tq >> tile_w_log2);
  if (qx >= lt.W[l] || qy >= lt.H[l]) return;
  msda_encoder_pair<L, P, FUSED>(value, lt, ol, n, lt.start[l] + qy * lt.W[l] + qx, m, lane8, S, M, out, value_bias, ol_bias, split);
}

// Staged variant of the fused tiled kernel (round 2).  ncu of msda_encoder_tiled_kernel<3,4,true>: issue-bound (SM 69 %), 1890
// instructions per lane and (query, head) -- of which the gather itself (48 x {address, 128-bit load, 4 FMA}) is ~350: every
// one of the 8 lanes of a (query, head) group recomputed the whole set-up (36 bias loads, 12 expf, 24 IEEE divisions, 12
// floor / weight / address computations).  Here the 8 lanes SHARE it: lane j prepares sample points j and j + 8 (softmax
// normaliser by two 8-lane shuffle reductions), writes each point's four corner offsets and four weights (attention weight
// folded in) to shared memory, and after a __syncwarp every lane walks the 12 points with two broadcast 128-bit shared loads
// per point.  The value bias enters once per (query, head): sum_p aw_p * (w1 + w2 + w3 + w4)_p * b, reduced over the 8 lanes.
// Same mathematics as sample4_biased (reference: ms_deform_im2col_cuda.cuh:38-89, 242-304; ms_deform_attn.py:101-108);
// the attention weight multiplies the corner weights before instead of after the corner sum (1 ulp-level differences).
template <int L, int P>
__global__ void __launch_bounds__(256)
msda_encoder_staged_kernel(const float* __restrict__ value, LevelTable lt, TileTable tt, const float* __restrict__ ol,
                           int S, int M, int tile_w_log2, float* __restrict__ out, const float* __restrict__ value_bias,
                           const float* __restrict__ ol_bias, int split) {
  constexpr int LP = L * P;
  static_assert(LP <= 16, "two sample points per lane");
  __shared__ int4 s_off[32][LP];
  __shared__ float4 s_w[32][LP];
  const int lane8 = threadIdx.x & 7;
  const int tq = threadIdx.x >> 3;                 // query inside the tile = (query, head) unit of this CTA
  const int tile = blockIdx.x, m = blockIdx.y, n = blockIdx.z;
  // level of this tile and its geometry by explicit selects (run-time indexing of the by-value tables compiles to long
  // uniform compare / select chains)
  int Wq = lt.W[0], Hq = lt.H[0], startq = lt.start[0], firstq = tt.first[0], tilesxq = tt.tiles_x[0];
#pragma unroll
  for (int i = 1; i < L; ++i) {
    if (tile >= tt.first[i]) {
      Wq = lt.W[i];
      Hq = lt.H[i];
      startq = lt.start[i];
      firstq = tt.first[i];
      tilesxq = tt.tiles_x[i];
    }
  }
  const int t = tile - firstq;
  const int ty = t / tilesxq, tx = t - ty * tilesxq;
  const int tile_w = 1 << tile_w_log2, tile_h = 32 >> tile_w_log2;
  const int qx0 = tx * tile_w + (tq & (tile_w - 1));
  const int qy0 = ty * tile_h + (tq >> tile_w_log2);
  const bool valid = qx0 < Wq && qy0 < Hq;
  // out-of-tile units run on a clamped query (the shuffles below need every lane) and skip the store
  const int qx = min(qx0, Wq - 1), qy = min(qy0, Hq - 1);
  const int q = startq + qy * Wq + qx;
  const long long nq = (long long)n * S + q;
  const float refx = ((float)qx + 0.5f) / (float)Wq;
  const float refy = ((float)qy + 0.5f) / (float)Hq;
  const float* row = ol + nq * (long long)(M * LP * 3);
  const int pix_stride = M * 32;

  // ---- phase 1: lane j prepares points j and j + 8 -------------------------------------------------------------------
  float lg[2], ox[2], oy[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int pp = lane8 + 8 * k;
    lg[k] = -INFINITY;
    ox[k] = oy[k] = 0.f;
    if (pp < LP) {
      const float2 o2 = __ldg(reinterpret_cast<const float2*>(row + m * (LP * 2) + pp * 2));
      ox[k] = o2.x;
      oy[k] = o2.y;
      lg[k] = __ldg(row + M * LP * 2 + m * LP + pp);
      if (ol_bias != nullptr) {
        ox[k] += __ldg(ol_bias + m * (LP * 2) + pp * 2);
        oy[k] += __ldg(ol_bias + m * (LP * 2) + pp * 2 + 1);
        lg[k] += __ldg(ol_bias + M * LP * 2 + m * LP + pp);
      }
    }
  }
  float mx = fmaxf(lg[0], lg[1]);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float e[2];
  e[0] = expf(lg[0] - mx);                          // expf(-inf) = 0 for the unused slot
  e[1] = expf(lg[1] - mx);
  float sum = e[0] + e[1];
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  float bpart = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int pp = lane8 + 8 * k;
    if (pp < LP) {
      const int l = pp / P;
      // explicit selects over the L levels: indexing the by-value table with a run-time level compiles to a long
      // compare / select chain over all kMaxLevels entries (ncu: 264 uniform compares in this kernel)
      int H = lt.H[0], W = lt.W[0], lbase = lt.start[0];
#pragma unroll
      for (int i = 1; i < L; ++i) {
        if (l == i) {
          H = lt.H[i];
          W = lt.W[i];
          lbase = lt.start[i];
        }
      }
      const float locx = refx + ox[k] / (float)W;
      const float locy = refy + oy[k] / (float)H;
      const float x = locx * (float)W - 0.5f, y = locy * (float)H - 0.5f;
      const bool inside = (y > -1.f) && (x > -1.f) && (y < (float)H) && (x < (float)W);
      const float xf = floorf(x), yf = floorf(y);
      const int x0 = (int)xf, y0 = (int)yf;
      const float lx = x - xf, ly = y - yf;
      const float hx = 1.f - lx, hy = 1.f - ly;
      const bool x0ok = x0 >= 0, x1ok = x0 + 1 <= W - 1, y0ok = y0 >= 0, y1ok = y0 + 1 <= H - 1;
      const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
      const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
      const float w1 = (inside && y0ok && x0ok) ? hy * hx : 0.f;
      const float w2 = (inside && y0ok && x1ok) ? hy * lx : 0.f;
      const float w3 = (inside && y1ok && x0ok) ? ly * hx : 0.f;
      const float w4 = (inside && y1ok && x1ok) ? ly * lx : 0.f;
      const float aw = e[k] * inv;
      s_off[tq][pp] = make_int4((lbase + yc0 * W + xc0) * pix_stride, (lbase + yc0 * W + xc1) * pix_stride,
                                (lbase + yc1 * W + xc0) * pix_stride, (lbase + yc1 * W + xc1) * pix_stride);
      s_w[tq][pp] = make_float4(aw * w1, aw * w2, aw * w3, aw * w4);
      bpart = fmaf(aw, (w1 + w2) + (w3 + w4), bpart);
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) bpart += __shfl_xor_sync(0xffffffffu, bpart, o);
  __syncwarp();                                     // the 8 lanes of a unit sit in one warp

  // ---- phase 2: every lane gathers its 4 channels over the 12 points -------------------------------------------------
  const float* base = value + (size_t)n * S * pix_stride + m * 32 + lane8 * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int pp = 0; pp < LP; ++pp) {
    const int4 o = s_off[tq][pp];
    const float4 w = s_w[tq][pp];
    const float4 v1 = ldg_f4(base + o.x), v2 = ldg_f4(base + o.y), v3 = ldg_f4(base + o.z), v4 = ldg_f4(base + o.w);
    fma4(acc, w.x, v1);
    fma4(acc, w.y, v2);
    fma4(acc, w.z, v3);
    fma4(acc, w.w, v4);
  }
  if (value_bias != nullptr) fma4(acc, bpart, ldg_f4(value_bias + m * 32 + lane8 * 4));
  if (valid) store_maybe_split(out, (size_t)nq, M * 32, m * 32 + lane8 * 4, acc, split);
}

static int fill_levels(LevelTable& lt, const int64_t* shapes_h, const int64_t* lsi_h, int L) {
  for (int l = 0; l < L; ++l) {
    lt.H[l] = (int)shapes_h[2 * l];
    lt.W[l] = (int)shapes_h[2 * l + 1];
    lt.start[l] = (int)lsi_h[l];
  }
  return 0;
}

}  // namespace univs

using namespace univs;

static bool is_device_pointer(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice;
}

// Host-side read of a small level table (the encoder entry points of this library take host tables; a device pointer is
// still accepted there through a blocking 64-byte copy -- the reference-ABI entry below never takes this path).
static int load_small_i64(const int64_t* p, int n, int64_t* dst) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    for (int i = 0; i < n; ++i) dst[i] = p[i];
    return 0;
  }
  if (attr.type == cudaMemoryTypeDevice) {
    e = cudaMemcpy(dst, p, sizeof(int64_t) * n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
      set_error("cudaMemcpy of level table failed: %s", cudaGetErrorString(e));
      return UNIVS_E_LAUNCH;
    }
  } else {
    for (int i = 0; i < n; ++i) dst[i] = p[i];
  }
  return 0;
}

extern "C" int univs_ms_deform_attn_forward_f32(void* stream, const float* value, const int64_t* spatial_shapes,
                                                const int64_t* level_start_index, const float* sampling_loc,
                                                const float* attn_weight, int batch, int spatial_size,
                                                int num_heads, int channels, int num_levels, int num_query,
                                                int num_point, float* out) {
  UNIVS_REQUIRE(batch >= 0 && spatial_size >= 0 && num_query >= 0, "ms_deform_attn_forward: negative size");
  if (batch == 0 || num_query == 0) return UNIVS_OK;   // empty output: nothing to do (pointers may be null)
  UNIVS_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out,
                "ms_deform_attn_forward: null pointer");
  UNIVS_REQUIRE(num_heads > 0 && channels > 0 && num_point > 0, "ms_deform_attn_forward: bad head/channel/point");
  UNIVS_REQUIRE(num_levels > 0 && num_levels <= kMaxLevels, "ms_deform_attn_forward: num_levels must be 1..%d",
                kMaxLevels);
  // spatial_shapes / level_start_index: DEVICE int64 tensors in the reference ABI (ms_deform_attn_cuda.cu:25-32).  Device
  // tables are read by the kernel itself -- stream-ordered, no device->host copy, capturable -- and, like in the reference,
  // not validated; host tables (this library's own callers) are validated here and passed by value.
  LevelTable lt = {};
  const int64_t *dev_shapes = nullptr, *dev_lsi = nullptr;
  if (is_device_pointer(spatial_shapes) || is_device_pointer(level_start_index)) {
    UNIVS_REQUIRE(is_device_pointer(spatial_shapes) && is_device_pointer(level_start_index),
                  "ms_deform_attn_forward: spatial_shapes and level_start_index must live in the same memory space");
    dev_shapes = spatial_shapes;
    dev_lsi = level_start_index;
  } else {
    fill_levels(lt, spatial_shapes, level_start_index, num_levels);
    long long tot = 0;
    for (int l = 0; l < num_levels; ++l) {
      UNIVS_REQUIRE(lt.H[l] > 0 && lt.W[l] > 0, "ms_deform_attn_forward: empty level %d", l);
      tot += (long long)lt.H[l] * lt.W[l];
    }
    UNIVS_REQUIRE(tot == spatial_size, "ms_deform_attn_forward: sum(H*W)=%lld != spatial_size=%d", tot, spatial_size);
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (channels % 4 == 0) {
    const long long total = (long long)batch * num_query * num_heads * (channels / 4);
    const int grid = (int)min((total + 255) / 256, (long long)148 * 64);
    msda_generic_vec4_kernel<<<grid, 256, 0, st>>>(value, lt, dev_shapes, dev_lsi, sampling_loc, attn_weight, batch, spatial_size,
                                                   num_heads, channels, num_levels, num_query, num_point, out);
  } else {
    const long long total = (long long)batch * num_query * num_heads * channels;
    const int grid = (int)min((total + 255) / 256, (long long)148 * 64);
    msda_generic_scalar_kernel<<<grid, 256, 0, st>>>(value, lt, dev_shapes, dev_lsi, sampling_loc, attn_weight, batch, spatial_size,
                                                     num_heads, channels, num_levels, num_query, num_point, out);
  }
  return check_launch("ms_deform_attn_forward");
}

extern "C" int univs_ms_deform_attn_backward_f32(void) {
  set_error("ms_deform_attn_backward: training path is out of scope (inference-only build)");
  return UNIVS_E_NOTIMPL;
}

extern "C" int univs_ms_deform_attn_encoder_f32(void* stream, const float* value, const int64_t* spatial_shapes,
                                                const int64_t* level_start_index, const float* offs_logits,
                                                int batch, int spatial_size, int num_heads, int num_levels,
                                                int num_point, float* out) {
  UNIVS_REQUIRE(value && spatial_shapes && level_start_index && offs_logits && out,
                "ms_deform_attn_encoder: null pointer");
  UNIVS_REQUIRE(num_levels == 3 && num_point == 4,
                "ms_deform_attn_encoder: only L=3, P=4 is instantiated (got L=%d P=%d)", num_levels, num_point);
  UNIVS_REQUIRE(num_heads > 0 && batch >= 0 && spatial_size >= 0, "ms_deform_attn_encoder: bad sizes");
  if (batch == 0 || spatial_size == 0) return UNIVS_OK;
  int64_t sh[2 * kMaxLevels], ls[kMaxLevels];
  int rc = load_small_i64(spatial_shapes, 2 * num_levels, sh);
  if (rc) return rc;
  rc = load_small_i64(level_start_index, num_levels, ls);
  if (rc) return rc;
  LevelTable lt;
  fill_levels(lt, sh, ls, num_levels);
  long long tot = 0;
  for (int l = 0; l < num_levels; ++l) tot += (long long)lt.H[l] * lt.W[l];
  UNIVS_REQUIRE(tot == spatial_size, "ms_deform_attn_encoder: sum(H*W)=%lld != spatial_size=%d", tot, spatial_size);
  const long long pairs = (long long)batch * spatial_size * num_heads;
  const long long grid = (pairs + 31) / 32;
  UNIVS_REQUIRE(grid < (1ll << 31), "ms_deform_attn_encoder: problem too large");
  msda_encoder_kernel<3, 4><<<(int)grid, 256, 0, (cudaStream_t)stream>>>(value, lt, offs_logits, batch,
                                                                         spatial_size, num_heads, out);
  return check_launch("ms_deform_attn_encoder");
}

extern "C" int univs_ms_deform_attn_encoder_tiled_f32(void* stream, const float* value, const int64_t* spatial_shapes,
                                                      const int64_t* level_start_index, const float* offs_logits,
                                                      int batch, int spatial_size, int num_heads, int num_levels,
                                                      int num_point, int tile_width, const float* value_bias,
                                                      const float* offs_logits_bias, int split, float* out) {
  UNIVS_REQUIRE(value && spatial_shapes && level_start_index && offs_logits && out,
                "ms_deform_attn_encoder_tiled: null pointer");
  UNIVS_REQUIRE(num_levels == 3 && num_point == 4,
                "ms_deform_attn_encoder_tiled: only L=3, P=4 is instantiated (got L=%d P=%d)", num_levels, num_point);
  UNIVS_REQUIRE(num_heads > 0 && num_heads <= 65535 && batch >= 0 && batch <= 65535 && spatial_size >= 0,
                "ms_deform_attn_encoder_tiled: bad sizes");
  UNIVS_REQUIRE(tile_width == 1 || tile_width == 2 || tile_width == 4 || tile_width == 8 || tile_width == 16 || tile_width == 32,
                "ms_deform_attn_encoder_tiled: tile_width must be a power of two <= 32 (got %d)", tile_width);
  if (batch == 0 || spatial_size == 0) return UNIVS_OK;
  int64_t sh[2 * kMaxLevels], ls[kMaxLevels];
  int rc = load_small_i64(spatial_shapes, 2 * num_levels, sh);
  if (rc) return rc;
  rc = load_small_i64(level_start_index, num_levels, ls);
  if (rc) return rc;
  LevelTable lt;
  fill_levels(lt, sh, ls, num_levels);
  int log2w = 0;
  while ((1 << log2w) < tile_width) ++log2w;
  const int tile_h = 32 / tile_width;
  TileTable tt;
  long long tot = 0, tiles = 0;
  for (int l = 0; l < num_levels; ++l) {
    UNIVS_REQUIRE(lt.H[l] > 0 && lt.W[l] > 0, "ms_deform_attn_encoder_tiled: empty level %d", l);
    tot += (long long)lt.H[l] * lt.W[l];
    tt.first[l] = (int)tiles;
    tt.tiles_x[l] = (lt.W[l] + tile_width - 1) / tile_width;
    tiles += (long long)tt.tiles_x[l] * ((lt.H[l] + tile_h - 1) / tile_h);
  }
  tt.first[num_levels] = (int)tiles;
  UNIVS_REQUIRE(tot == spatial_size, "ms_deform_attn_encoder_tiled: sum(H*W)=%lld != spatial_size=%d", tot, spatial_size);
  UNIVS_REQUIRE(tiles < (1ll << 31), "ms_deform_attn_encoder_tiled: problem too large");
  const dim3 grid((unsigned)tiles, (unsigned)num_heads, (unsigned)batch);
  if (value_bias == nullptr && offs_logits_bias == nullptr && split == 0) {
    msda_encoder_tiled_kernel<3, 4, false><<<grid, 256, 0, (cudaStream_t)stream>>>(value, lt, tt, offs_logits, spatial_size,
                                                                                 num_heads, log2w, out, nullptr, nullptr, 0);
  } else {
    const int C = num_heads * 32;
    UNIVS_REQUIRE(split == 0 || split == UNIVS_SPLIT_F16U || split == UNIVS_SPLIT_F16C ||
                      (split != -1 && (split > 0 ? split : -split) % 4 == 0 && C % (split > 0 ? split : -split) == 0),
                  "ms_deform_attn_encoder_tiled: split chunk must divide heads*32");
    msda_encoder_staged_kernel<3, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(value, lt, tt, offs_logits, spatial_size,
                                                                                num_heads, log2w, out, value_bias,
                                                                                offs_logits_bias, split);
  }
  return check_launch("ms_deform_attn_encoder_tiled");
}
