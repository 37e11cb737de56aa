// GroupNorm on channel-last activations, fused with what surrounds it in the FPN of the pixel decoder
// (reference: detectron2 Conv2d = conv -> GroupNorm(32) -> [ReLU], msdeformattn.py:218-221 (input_proj), :249-283
//  (lateral / output convs) and the top-down path  y = lateral(x) + interpolate(y_prev, bilinear)  (:345-354)).
//
// The convolutions run as library GEMMs on channel-last tokens, so GroupNorm sees x[n, y, xw, c] with c contiguous.
// torch's GroupNorm wants NCHW: on this path that cost two transposing copies, a moments pass, an apply pass, a
// separate ReLU, a separate upsample + add, a separate operand split and a separate zero-pad copy per FPN level --
// sixteen HBM passes over the 1/4-resolution map (~4.6 ms per clip at the north-star shape).  Here it is
//   1. partial : per (frame, pixel chunk) and group, sum and sum of squares in fp64        (one read of x)
//   2. finalize: per (frame, group) mean and 1/sqrt(var + eps)
//   3. apply   : y = (x - mean) * rstd * gamma + beta  [+ bilinear(lowres)]  [ReLU], written as fp32 and / or straight in
//                the GEMM operand format of the next layer, optionally at a zero-padded position (the 3x3 convolution
//                reads a spatially padded token matrix)                              (one read of x, one write per output)
// All three are HBM streams.  `x` may live inside a padded row buffer (the convolution's output): it is addressed as
// x[n * img_stride + y * row_stride + xw * C + c].
#include "rowwise.cuh"

namespace univs {

constexpr int kGnThreads = 256;
constexpr int kGnMaxGroups = 256;

__global__ void __launch_bounds__(kGnThreads)
gn_partial_kernel(const float* __restrict__ x, int H, int W, int C, long long img_stride, long long row_stride, int cpg,
                  int groups, int pixels_per_chunk, double2* __restrict__ partial) {
  __shared__ double s_sum[kGnMaxGroups], s_sq[kGnMaxGroups];
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int nq = C >> 2;                  // channel quads per pixel; kGnThreads % nq == 0 (checked on the host)
  const int lanes = kGnThreads / nq;      // pixels in flight per block
  const int q = threadIdx.x % nq, lane = threadIdx.x / nq;
  for (int g = threadIdx.x; g < groups; g += kGnThreads) { s_sum[g] = 0.0; s_sq[g] = 0.0; }
  __syncthreads();
  const int HW = H * W;
  const int p0 = chunk * pixels_per_chunk;
  const int p1 = min(p0 + pixels_per_chunk, HW);
  const float* base = x + (size_t)n * img_stride + (size_t)q * 4;
  double S = 0.0, Q = 0.0;
#pragma unroll 4
  for (int p = p0 + lane; p < p1; p += lanes) {
    const int y = p / W, xw = p - y * W;
    const float4 a = *reinterpret_cast<const float4*>(base + (size_t)y * row_stride + (size_t)xw * C);
    S += (double)((a.x + a.y) + (a.z + a.w));
    Q += (double)((a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w));
  }
  const int g = (q * 4) / cpg;            // cpg % 4 == 0: a quad never straddles two groups
  atomicAdd(&s_sum[g], S);
  atomicAdd(&s_sq[g], Q);
  __syncthreads();
  for (int gg = threadIdx.x; gg < groups; gg += kGnThreads)
    partial[((size_t)n * gridDim.x + chunk) * groups + gg] = make_double2(s_sum[gg], s_sq[gg]);
}

__global__ void gn_finalize_kernel(const double2* __restrict__ partial, int N, int groups, int chunks, double count,
                                   float eps, float2* __restrict__ mean_rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * groups) return;
  const int n = i / groups, g = i - n * groups;
  double S = 0.0, Q = 0.0;
  for (int c = 0; c < chunks; ++c) {
    const double2 v = partial[((size_t)n * chunks + c) * groups + g];
    S += v.x;
    Q += v.y;
  }
  const double mean = S / count;
  double var = Q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_rstd[i] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

// source index of torch's bilinear resize with align_corners=False (UpSample.h area_pixel_compute_source_index)
__device__ __forceinline__ void bilinear_tap(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(const float* __restrict__ x, int N, int H, int W, int C, long long img_stride, long long row_stride,
                const float2* __restrict__ mean_rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                int cpg, int groups, const float* __restrict__ lowres, long long lowres_img_stride, int h2, int w2, int relu,
                float* __restrict__ out_f32, float* __restrict__ out_split, int split, int pad) {
  const int nq = C >> 2;
  const int n = blockIdx.y;                                  // one frame per grid row: 32-bit index math inside
  const int total = H * W * nq;
  const float sy = lowres ? (float)h2 / (float)H : 0.f, sx = lowres ? (float)w2 / (float)W : 0.f;
  const int Wp = W + 2 * pad, Hp = H + 2 * pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int q = i % nq;
    const int pix_in = i / nq;
    const int y = pix_in / W, xw = pix_in - y * W;
    const size_t pix = (size_t)n * H * W + pix_in;
    const int col = q * 4;
    const float4 a = *reinterpret_cast<const float4*>(x + (size_t)n * img_stride + (size_t)y * row_stride + (size_t)xw * C + col);
    const float2 mr = mean_rstd[n * groups + col / cpg];
    const float4 gm = ldg_f4(gamma + col), bt = ldg_f4(beta + col);
    float4 o;
    o.x = (a.x - mr.x) * mr.y * gm.x + bt.x;
    o.y = (a.y - mr.x) * mr.y * gm.y + bt.y;
    o.z = (a.z - mr.x) * mr.y * gm.z + bt.z;
    o.w = (a.w - mr.x) * mr.y * gm.w + bt.w;
    if (lowres != nullptr) {
      int y0, y1, x0, x1;
      float ly, lx;
      bilinear_tap(y, sy, h2, y0, y1, ly);
      bilinear_tap(xw, sx, w2, x0, x1, lx);
      const float hy = 1.f - ly, hx = 1.f - lx;
      const float* lr = lowres + (size_t)n * lowres_img_stride + col;
      const float4 v00 = *reinterpret_cast<const float4*>(lr + ((size_t)y0 * w2 + x0) * C);
      const float4 v01 = *reinterpret_cast<const float4*>(lr + ((size_t)y0 * w2 + x1) * C);
      const float4 v10 = *reinterpret_cast<const float4*>(lr + ((size_t)y1 * w2 + x0) * C);
      const float4 v11 = *reinterpret_cast<const float4*>(lr + ((size_t)y1 * w2 + x1) * C);
      o.x += hy * (hx * v00.x + lx * v01.x) + ly * (hx * v10.x + lx * v11.x);
      o.y += hy * (hx * v00.y + lx * v01.y) + ly * (hx * v10.y + lx * v11.y);
      o.z += hy * (hx * v00.z + lx * v01.z) + ly * (hx * v10.z + lx * v11.z);
      o.w += hy * (hx * v00.w + lx * v01.w) + ly * (hx * v10.w + lx * v11.w);
    }
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + (size_t)pix * C + col) = o;
    if (out_split != nullptr) {
      const size_t row = ((size_t)n * Hp + (y + pad)) * Wp + (xw + pad);
      store_maybe_split(out_split, row, C, col, o, split);
    }
  }
}

static int gn_chunks(long long hw) {
  long long c = (hw + 255) / 256;       // >= 256 pixels (64 per pixel lane at C = 256) per block
  if (c < 1) c = 1;
  if (c > 128) c = 128;
  return (int)c;
}

}  // namespace univs

using namespace univs;

static bool gn_shape_ok(int C, int groups) {
  if (C <= 0 || groups <= 0 || C % groups) return false;
  const int cpg = C / groups, nq = C / 4;
  return C % 4 == 0 && cpg % 4 == 0 && groups <= kGnMaxGroups && nq <= kGnThreads && kGnThreads % nq == 0;
}

extern "C" int64_t univs_groupnorm_workspace_bytes(int frames, int height, int width, int groups) {
  if (frames <= 0 || height <= 0 || width <= 0 || groups <= 0) return 0;
  return (int64_t)frames * gn_chunks((long long)height * width) * groups * (int64_t)sizeof(double2);
}

extern "C" int univs_groupnorm_stats_f32(void* stream, const float* x, int frames, int height, int width, int channels,
                                         int64_t img_stride, int64_t row_stride, int groups, float eps, void* workspace,
                                         float* mean_rstd) {
  UNIVS_REQUIRE(frames >= 0 && frames <= 65535 && height >= 0 && width >= 0, "groupnorm_stats: bad sizes");
  UNIVS_REQUIRE(gn_shape_ok(channels, groups),
                "groupnorm_stats: need C %% groups == 0, (C/groups) %% 4 == 0, 256 %% (C/4) == 0 (got C=%d, groups=%d)", channels, groups);
  if (frames == 0 || height == 0 || width == 0) return UNIVS_OK;
  UNIVS_REQUIRE(x && workspace && mean_rstd, "groupnorm_stats: null pointer");
  UNIVS_REQUIRE(row_stride >= (int64_t)width * channels && img_stride >= (int64_t)(height - 1) * row_stride + (int64_t)width * channels &&
                    row_stride % 4 == 0 && img_stride % 4 == 0,
                "groupnorm_stats: strides must cover the image and keep 16-byte alignment");
  const long long hw = (long long)height * width;
  const int chunks = gn_chunks(hw);
  const int ppc = (int)((hw + chunks - 1) / chunks);
  cudaStream_t st = (cudaStream_t)stream;
  gn_partial_kernel<<<dim3(chunks, frames), kGnThreads, 0, st>>>(x, height, width, channels, img_stride, row_stride,
                                                                   channels / groups, groups, ppc,
                                                                   reinterpret_cast<double2*>(workspace));
  int rc = check_launch("groupnorm_partial");
  if (rc) return rc;
  const int tot = frames * groups;
  gn_finalize_kernel<<<(tot + 127) / 128, 128, 0, st>>>(reinterpret_cast<const double2*>(workspace), frames, groups, chunks,
                                                        (double)hw * (channels / groups), eps,
                                                        reinterpret_cast<float2*>(mean_rstd));
  return check_launch("groupnorm_finalize");
}

extern "C" int univs_groupnorm_apply_f32(void* stream, const float* x, int frames, int height, int width, int channels,
                                         int64_t img_stride, int64_t row_stride, const float* mean_rstd, const float* gamma,
                                         const float* beta, int groups, const float* lowres, int64_t lowres_img_stride, int low_height,
                                         int low_width, int relu, float* out_f32, void* out_split, int split, int pad) {
  UNIVS_REQUIRE(frames >= 0 && height >= 0 && width >= 0 && pad >= 0, "groupnorm_apply: bad sizes");
  UNIVS_REQUIRE(gn_shape_ok(channels, groups),
                "groupnorm_apply: need C %% groups == 0, (C/groups) %% 4 == 0, 256 %% (C/4) == 0 (got C=%d, groups=%d)", channels, groups);
  if (frames == 0 || height == 0 || width == 0) return UNIVS_OK;
  UNIVS_REQUIRE(x && mean_rstd && gamma && beta, "groupnorm_apply: null pointer");
  UNIVS_REQUIRE(out_f32 != nullptr || out_split != nullptr, "groupnorm_apply: no output requested");
  UNIVS_REQUIRE(lowres == nullptr || (low_height > 0 && low_width > 0 && lowres_img_stride % 4 == 0 &&
                                      lowres_img_stride >= (int64_t)low_height * low_width * channels),
                "groupnorm_apply: bad low-resolution size / stride");
  UNIVS_REQUIRE(out_split == nullptr || split == UNIVS_SPLIT_F16U || split == UNIVS_SPLIT_F16C ||
                    (split != 0 && split != -1 && (split > 0 ? split : -split) % 4 == 0 &&
                     channels % (split > 0 ? split : -split) == 0),
                "groupnorm_apply: split chunk must divide channels");
  UNIVS_REQUIRE(row_stride % 4 == 0 && img_stride % 4 == 0, "groupnorm_apply: strides must keep 16-byte alignment");
  UNIVS_REQUIRE((long long)height * width * (channels / 4) < (1ll << 30) && frames <= 65535, "groupnorm_apply: image too large");
  const long long total = (long long)height * width * (channels / 4);
  long long blocks = (total + kGnThreads - 1) / kGnThreads;
  const long long cap = (148ll * 16 + frames - 1) / frames;
  if (blocks > cap) blocks = cap;
  gn_apply_kernel<<<dim3((unsigned)blocks, (unsigned)frames), kGnThreads, 0, (cudaStream_t)stream>>>(
      x, frames, height, width, channels, img_stride, row_stride, reinterpret_cast<const float2*>(mean_rstd), gamma, beta,
      channels / groups, groups, lowres, lowres_img_stride, low_height, low_width, relu, out_f32, reinterpret_cast<float*>(out_split), split, pad);
  return check_launch("groupnorm_apply");
}
