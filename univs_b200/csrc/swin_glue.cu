// Two gather-fused row kernels at the edges of the Swin stages:
//
//   patchify : frames [N,3,H,W] (uint8 or fp32, values 0..255) -> patch matrix [N*(Hp/P)*(Wp/P), 3*P*P] fp32 with the
//              pixel normalisation (x - mean) / std and the zero padding to (Hp, Wp) applied on the fly, so the 4x4/stride-4
//              patch projection (swin.py:456-495 PatchEmbed.proj) becomes one GEMM on it.  Replaces: cast, subtract, divide,
//              pad (univs_prompt.py:165-168 / ImageList.from_tensors) + the strided convolution + its NCHW->NHWC copy.
//   merge2x2 : PatchMerging (swin.py:298-337): x [N,H,W,C] -> LayerNorm over the 4C channels of each 2x2 neighbourhood
//              [x(2i,2j) | x(2i+1,2j) | x(2i,2j+1) | x(2i+1,2j+1)] (zero beyond odd borders), written plain or in the GEMM
//              operand format of the reduction linear.  Replaces the concatenated [N,H/2,W/2,4C] copy + its re-read.
// Both are pure HBM streams; one warp per output row in merge2x2 (row held in registers, as in layernorm_kernel).
#include "rowwise.cuh"

namespace univs {

template <typename T>
__device__ __forceinline__ float4 load4_as_float(const T* p);
template <>
__device__ __forceinline__ float4 load4_as_float<float>(const float* p) {
  return make_float4(p[0], p[1], p[2], p[3]);     // rows of arbitrary width: no 16-byte alignment guarantee
}
template <>
__device__ __forceinline__ float4 load4_as_float<unsigned char>(const unsigned char* p) {
  return make_float4((float)p[0], (float)p[1], (float)p[2], (float)p[3]);
}

template <typename T, int P>
__global__ void __launch_bounds__(256)
patchify_kernel(const T* __restrict__ frames, int N, int H, int W, int Hp, int Wp, float m0, float m1, float m2, float s0,
                float s1, float s2, float* __restrict__ out, int split) {
  static_assert(P == 4, "one thread handles the P = 4 pixels of one patch row");
  constexpr int K = 3 * P * P;        // 48
  constexpr int QUADS = K / 4;        // 12 = (channel, ky) pairs
  const int pw = Wp / P, ph = Hp / P;
  const long long total = (long long)N * ph * pw * QUADS;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int quad = (int)(i % QUADS);
    const long long row = i / QUADS;
    const int px = (int)(row % pw);
    const long long r2 = row / pw;
    const int py = (int)(r2 % ph);
    const int n = (int)(r2 / ph);
    const int c = quad / P, ky = quad - c * P;
    const int y = py * P + ky, x0 = px * P;
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < H && x0 < W) {
      const T* src = frames + (((size_t)n * 3 + c) * H + y) * W + x0;
      if (x0 + 3 < W) {
        const float4 a = load4_as_float<T>(src);
        v = make_float4(__fdiv_rn(a.x - mean, sd), __fdiv_rn(a.y - mean, sd), __fdiv_rn(a.z - mean, sd), __fdiv_rn(a.w - mean, sd));
      } else {        // right border of a width that is not a multiple of 4: the rest is zero padding
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < 4 && x0 + k < W; ++k) t[k] = __fdiv_rn((float)src[k] - mean, sd);
        v = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    store_maybe_split(out, (size_t)row, K, quad * 4, v, split);
  }
}

template <int MAXV>  // float4 per lane; 4*C <= MAXV*128
__global__ void __launch_bounds__(256)
layernorm_merge_kernel(const float* __restrict__ x, int N, int H, int W, int C, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, float* __restrict__ out, int split) {
  const int oh = (H + 1) >> 1, ow = (W + 1) >> 1;
  const long long rows = (long long)N * oh * ow;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int ox = (int)(row % ow);
  const long long r2 = row / ow;
  const int oy = (int)(r2 % oh);
  const int n = (int)(r2 / oh);
  const int C4 = 4 * C;
  const int nv = C4 >> 2;   // float4 per merged row; C % 4 == 0 so a float4 never straddles two source pixels
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < nv) {
      const int col = idx * 4;
      const int seg = col / C, cin = col - seg * C;
      const int sy = 2 * oy + (seg & 1), sx = 2 * ox + (seg >> 1);      // segment order of the reference's torch.cat
      if (sy < H && sx < W) v[i] = *reinterpret_cast<const float4*>(x + (((size_t)n * H + sy) * W + sx) * C + cin);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)C4;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C4 + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float4 gm = ldg_f4(gamma + idx * 4), bt = ldg_f4(beta + idx * 4);
      float4 o;
      o.x = (v[i].x - mean) * rstd * gm.x + bt.x;
      o.y = (v[i].y - mean) * rstd * gm.y + bt.y;
      o.z = (v[i].z - mean) * rstd * gm.z + bt.z;
      o.w = (v[i].w - mean) * rstd * gm.w + bt.w;
      store_maybe_split(out, (size_t)row, C4, idx * 4, o, split);
    }
  }
}

// LayerNorm with several consumers (post-norm transformer layers, msdeformattn.py:126-133: the normalised tokens are the
// residual stream AND the input of the next GEMMs, one of them after adding the positional embedding):
//   s = x (+ residual (+ residual_bias));  y = LN(s) * gamma + beta
//   out_f32 (nullable) = y;  out_split (nullable) = operand(y);  out_split_pos (nullable) = operand(y + pos[row % pos_rows])
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_multi_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ res_bias,
                       const float* __restrict__ gamma, const float* __restrict__ beta, long long rows, int C, float eps,
                       float* __restrict__ out_f32, float* __restrict__ out_split, int split, const float* __restrict__ pos,
                       long long pos_rows, float* __restrict__ out_split_pos) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = C >> 2;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < nv) {
      float4 a = *reinterpret_cast<const float4*>(x + row * C + idx * 4);
      if (res != nullptr) {
        const float4 b = *reinterpret_cast<const float4*>(res + row * C + idx * 4);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        if (res_bias != nullptr) {
          const float4 rb = ldg_f4(res_bias + idx * 4);
          a.x += rb.x; a.y += rb.y; a.z += rb.z; a.w += rb.w;
        }
      }
      v[i] = a;
      s += (a.x + a.y) + (a.z + a.w);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  const long long prow = pos != nullptr ? row % pos_rows : 0;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + i * 32;
    if (idx < nv) {
      const float4 gm = ldg_f4(gamma + idx * 4), bt = ldg_f4(beta + idx * 4);
      float4 o;
      o.x = (v[i].x - mean) * rstd * gm.x + bt.x;
      o.y = (v[i].y - mean) * rstd * gm.y + bt.y;
      o.z = (v[i].z - mean) * rstd * gm.z + bt.z;
      o.w = (v[i].w - mean) * rstd * gm.w + bt.w;
      if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + row * C + idx * 4) = o;
      if (out_split != nullptr) store_maybe_split(out_split, (size_t)row, C, idx * 4, o, split);
      if (out_split_pos != nullptr) {
        const float4 pe = ldg_f4(pos + prow * C + idx * 4);
        store_maybe_split(out_split_pos, (size_t)row, C, idx * 4, make_float4(o.x + pe.x, o.y + pe.y, o.z + pe.z, o.w + pe.w), split);
      }
    }
  }
}

}  // namespace univs

using namespace univs;

static bool split_ok(int split, int channels) {
  if (split == 0 || split == UNIVS_SPLIT_F16U || split == UNIVS_SPLIT_F16C) return true;
  const int a = split > 0 ? split : -split;
  return split != -1 && a % 4 == 0 && channels % a == 0;
}

extern "C" int univs_patchify_normalize(void* stream, const void* frames, int is_uint8, int num_frames, int height, int width,
                                        int padded_height, int padded_width, int patch, const float* mean3, const float* std3,
                                        float* out, int split) {
  UNIVS_REQUIRE(patch == 4, "patchify: only the 4x4 patch embedding of Swin is built (got %d)", patch);
  UNIVS_REQUIRE(num_frames >= 0 && height > 0 && width > 0 && padded_height >= height && padded_width >= width &&
                    padded_height % patch == 0 && padded_width % patch == 0,
                "patchify: padded size must cover the frame and be a multiple of the patch size");
  UNIVS_REQUIRE(split_ok(split, 3 * patch * patch), "patchify: split chunk must divide 48");
  if (num_frames == 0) return UNIVS_OK;
  UNIVS_REQUIRE(frames && mean3 && std3 && out, "patchify: null pointer");
  UNIVS_REQUIRE(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "patchify: zero pixel std");
  const long long total = (long long)num_frames * (padded_height / patch) * (padded_width / patch) * 12;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_uint8)
    patchify_kernel<unsigned char, 4><<<(unsigned)blocks, 256, 0, st>>>(
        static_cast<const unsigned char*>(frames), num_frames, height, width, padded_height, padded_width, mean3[0], mean3[1],
        mean3[2], std3[0], std3[1], std3[2], out, split);
  else
    patchify_kernel<float, 4><<<(unsigned)blocks, 256, 0, st>>>(
        static_cast<const float*>(frames), num_frames, height, width, padded_height, padded_width, mean3[0], mean3[1], mean3[2],
        std3[0], std3[1], std3[2], out, split);
  return check_launch("patchify");
}

extern "C" int univs_layernorm_merge2x2_f32(void* stream, const float* x, int num_frames, int height, int width, int channels,
                                            const float* gamma, const float* beta, float eps, float* out, int split) {
  UNIVS_REQUIRE(num_frames >= 0 && height >= 0 && width >= 0 && channels > 0, "layernorm_merge2x2: bad sizes");
  UNIVS_REQUIRE(channels % 4 == 0 && 4 * channels <= 4096, "layernorm_merge2x2: channels %% 4 == 0 and 4*channels <= 4096 (got %d)", channels);
  UNIVS_REQUIRE(split_ok(split, 4 * channels), "layernorm_merge2x2: split chunk must divide 4*channels");
  if (num_frames == 0 || height == 0 || width == 0) return UNIVS_OK;
  UNIVS_REQUIRE(x && gamma && beta && out, "layernorm_merge2x2: null pointer");
  const long long rows = (long long)num_frames * ((height + 1) / 2) * ((width + 1) / 2);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  const int c4 = 4 * channels;
  cudaStream_t st = (cudaStream_t)stream;
#define LNM_LAUNCH(MV) layernorm_merge_kernel<MV><<<grid, 256, 0, st>>>(x, num_frames, height, width, channels, gamma, beta, eps, out, split)
  if (c4 <= 128) LNM_LAUNCH(1);
  else if (c4 <= 256) LNM_LAUNCH(2);
  else if (c4 <= 512) LNM_LAUNCH(4);
  else if (c4 <= 1024) LNM_LAUNCH(8);
  else if (c4 <= 2048) LNM_LAUNCH(16);
  else LNM_LAUNCH(32);
#undef LNM_LAUNCH
  return check_launch("layernorm_merge2x2");
}

extern "C" int univs_layernorm_multi_f32(void* stream, const float* x, const float* residual, const float* residual_bias,
                                         const float* gamma, const float* beta, int64_t rows, int channels, float eps,
                                         float* out_f32, void* out_split, int split, const float* pos, int64_t pos_rows,
                                         void* out_split_pos) {
  UNIVS_REQUIRE(rows >= 0 && channels > 0, "layernorm_multi: bad sizes");
  if (rows == 0) return UNIVS_OK;
  UNIVS_REQUIRE(x && gamma && beta, "layernorm_multi: null pointer");
  UNIVS_REQUIRE(out_f32 || out_split || out_split_pos, "layernorm_multi: no output requested");
  UNIVS_REQUIRE(residual_bias == nullptr || residual != nullptr, "layernorm_multi: residual_bias needs a residual");
  UNIVS_REQUIRE(channels % 4 == 0 && channels <= 4096, "layernorm_multi: channels must be a multiple of 4 and <= 4096 (got %d)", channels);
  UNIVS_REQUIRE((out_split == nullptr && out_split_pos == nullptr) || (split != 0 && split_ok(split, channels)),
                "layernorm_multi: operand outputs need a valid split code");
  UNIVS_REQUIRE(out_split_pos == nullptr || (pos != nullptr && pos_rows > 0 && rows % pos_rows == 0),
                "layernorm_multi: pos [pos_rows, channels] must tile the rows");
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
#define LNX_LAUNCH(MV)                                                                                                    \
  layernorm_multi_kernel<MV><<<grid, 256, 0, st>>>(x, residual, residual_bias, gamma, beta, rows, channels, eps, out_f32,      \
                                                   reinterpret_cast<float*>(out_split), split, pos, pos_rows,                  \
                                                   reinterpret_cast<float*>(out_split_pos))
  if (channels <= 128) LNX_LAUNCH(1);
  else if (channels <= 256) LNX_LAUNCH(2);
  else if (channels <= 512) LNX_LAUNCH(4);
  else if (channels <= 1024) LNX_LAUNCH(8);
  else if (channels <= 2048) LNX_LAUNCH(16);
  else LNX_LAUNCH(32);
#undef LNX_LAUNCH
  return check_launch("layernorm_multi");
}
