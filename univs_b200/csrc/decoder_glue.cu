// Attention masks of the intermediate decoder layers from POOLED mask features (opt-in: UNIVS_POOLED_MASKS=1).
//
// Reference (..._univs.py:527-566): every prediction head computes the full-resolution mask logits
// "btqc,btchw->btqhw" and then, for the NEXT layer's cross-attention, resizes them bilinearly to that layer's memory size
// and thresholds: blocked = sigmoid(resize(logits)) < 0.5.  SIZE_DIVISIBILITY = 32 makes every resize ratio an even
// integer, for which the bilinear resize (align_corners=False) is the mean of the centre 2x2 block of each r x r cell
// (what attn_mask_bits_kernel in mha.cu computes from the logits).  The resize is linear in the logits and the logits are
// linear in the mask features, so   resize(E . F) = E . resize(F):   pooling the mask features ONCE per clip to the three
// memory sizes lets the nine intermediate heads run the einsum at 1/4, 1/16 and 1/64 of the pixels and never write
// full-resolution logits (235 MB per head at the north-star size); only the last head, whose logits are the output, runs
// at full resolution.  Same mathematics; the fp32 rounding differs (the mean is taken before instead of after the dot
// product), so a logit within rounding distance of zero may flip its mask bit -- the class of deviation DESIGN.md 2
// quantifies with tests/tools/parity_at_scale.py.
//   mask_feature_pool : F [T,H,W,C] channel-last fp32 -> [T, h*w, C] pooled (plain fp32 or einsum operand format)
//   mask_bits_direct  : logits [Q,T,S] at the memory resolution -> bits [T,Q,ceil(S/32)] (bit set = blocked) + row flag
#include "rowwise.cuh"

namespace univs {

__global__ void __launch_bounds__(256)
mask_feature_pool_kernel(const float* __restrict__ f, int T, int H, int W, int C, int h, int w, int ry, int rx,
                         float* __restrict__ out, int split) {
  const int nq = C >> 2;
  const long long total = (long long)T * h * w * nq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % nq);
    const long long pix = i / nq;                     // t * h * w + ky * w + kx
    const int kx = (int)(pix % w);
    const long long r2 = pix / w;
    const int ky = (int)(r2 % h);
    const int t = (int)(r2 / h);
    const int y0 = ky * ry + (ry >> 1) - 1, x0 = kx * rx + (rx >> 1) - 1;
    const float* p = f + (((size_t)t * H + y0) * W + x0) * C + q * 4;
    const float4 a = ldg_f4(p), b = ldg_f4(p + C), c = ldg_f4(p + (size_t)W * C), d = ldg_f4(p + (size_t)W * C + C);
    float4 v;   // the bilinear weights of attn_mask_bits_kernel (both lambdas 0.5), applied to the features
    v.x = 0.5f * (0.5f * a.x + 0.5f * b.x) + 0.5f * (0.5f * c.x + 0.5f * d.x);
    v.y = 0.5f * (0.5f * a.y + 0.5f * b.y) + 0.5f * (0.5f * c.y + 0.5f * d.y);
    v.z = 0.5f * (0.5f * a.z + 0.5f * b.z) + 0.5f * (0.5f * c.z + 0.5f * d.z);
    v.w = 0.5f * (0.5f * a.w + 0.5f * b.w) + 0.5f * (0.5f * c.w + 0.5f * d.w);
    store_maybe_split(out, (size_t)pix, C, q * 4, v, split);
  }
}

__global__ void __launch_bounds__(256)
mask_bits_direct_kernel(const float* __restrict__ logits, int Q, int T, int S, uint32_t* __restrict__ bits,
                        int32_t* __restrict__ row_open) {
  const int words = (S + 31) >> 5;
  const long long total_words = (long long)T * Q * words;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= total_words) return;
  const int lane = threadIdx.x & 31;
  const int word = (int)(wid % words);
  const long long tq = wid / words;   // t * Q + q
  const int qi = (int)(tq % Q), ti = (int)(tq / Q);
  const int key = word * 32 + lane;
  bool blocked = true;
  if (key < S) {
    const float val = __ldg(logits + ((size_t)qi * T + ti) * S + key);
    const float sg = 1.f / (1.f + expf(-val));       // same test as attn_mask_bits_kernel
    blocked = sg < 0.5f;
  }
  const uint32_t ballot = __ballot_sync(0xffffffffu, blocked);
  if (lane == 0) {
    bits[wid] = ballot;
    const int valid = min(32, S - word * 32);
    const uint32_t vmask = valid == 32 ? 0xffffffffu : ((1u << valid) - 1u);
    if ((ballot & vmask) != vmask) atomicOr(row_open + tq, 1);
  }
}

}  // namespace univs

using namespace univs;

extern "C" int univs_mask_feature_pool_f32(void* stream, const float* feats_cl, int frames, int height, int width, int channels,
                                           int tgt_h, int tgt_w, void* out, int split) {
  UNIVS_REQUIRE(frames >= 0 && height > 0 && width > 0 && tgt_h > 0 && tgt_w > 0, "mask_feature_pool: bad sizes");
  UNIVS_REQUIRE(channels > 0 && channels % 4 == 0, "mask_feature_pool: channels %% 4 == 0");
  UNIVS_REQUIRE(height % tgt_h == 0 && width % tgt_w == 0, "mask_feature_pool: target size must divide the feature size");
  const int ry = height / tgt_h, rx = width / tgt_w;
  UNIVS_REQUIRE(ry % 2 == 0 && rx % 2 == 0, "mask_feature_pool: resize ratios must be even (got %d, %d)", ry, rx);
  UNIVS_REQUIRE(split == 0 || split == UNIVS_SPLIT_F16U, "mask_feature_pool: output is plain fp32 (0) or fp16 [hi|lo] (-2)");
  if (frames == 0) return UNIVS_OK;
  UNIVS_REQUIRE(feats_cl && out, "mask_feature_pool: null pointer");
  const long long total = (long long)frames * tgt_h * tgt_w * (channels / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mask_feature_pool_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(feats_cl, frames, height, width, channels, tgt_h,
                                                                               tgt_w, ry, rx, reinterpret_cast<float*>(out), split);
  return check_launch("mask_feature_pool");
}

extern "C" int univs_attn_mask_bits_direct_f32(void* stream, const float* logits, int queries, int frames, int keys,
                                               uint32_t* bits, int32_t* row_open) {
  UNIVS_REQUIRE(queries >= 0 && frames >= 0 && keys > 0, "attn_mask_bits_direct: bad sizes");
  if (queries == 0 || frames == 0) return UNIVS_OK;
  UNIVS_REQUIRE(logits && bits && row_open, "attn_mask_bits_direct: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(row_open, 0, sizeof(int32_t) * (size_t)queries * frames, st);
  if (e != cudaSuccess) { set_error("attn_mask_bits_direct: memset failed: %s", cudaGetErrorString(e)); return UNIVS_E_LAUNCH; }
  const int words = (keys + 31) / 32;
  const long long total = (long long)frames * queries * words;
  mask_bits_direct_kernel<<<(unsigned)((total + 7) / 8), 256, 0, st>>>(logits, queries, frames, keys, bits, row_open);
  return check_launch("attn_mask_bits_direct");
}
