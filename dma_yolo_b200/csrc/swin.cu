// 8f-1: the memory-bound pieces of SwinTransformerLayer (models/common.py:538-634) on NHWC bf16.  The four Linear
// layers (qkv, proj, fc1, fc2 — models/common.py:472-474, 105-108) are 1x1 convolutions and run on the tcgen05
// implicit-GEMM kernel (conv_tcgen05.cu); what is left is
//   layernorm        : nn.LayerNorm over the channel vector of every pixel (norm1 / norm2)
//   window_attention : W-MSA / SW-MSA of WindowAttention.forward (models/common.py:483-515) including the layer's
//                      data movement — zero padding to a multiple of the window, the cyclic shift (torch.roll),
//                      window_partition / window_reverse, the crop — as index arithmetic: nothing is permuted or
//                      copied, q/k/v are gathered straight from the qkv tensor and the result is scattered back.
// The reference reads the NCHW tensor as (b, c, w, h) and permutes it to (b, h, w, c) (models/common.py:596-597):
// its "rows" are OUR x (dim 3) and its "columns" OUR y (dim 2).  Window membership, the shift mask and the relative
// position bias are all defined in that transposed frame, so the kernel works in it: token (r, s) = pixel (y=s, x=r).
#include "common.cuh"

namespace dmay {

// ---- LayerNorm over C per pixel: one warp per pixel, values stay in registers between the two passes ----
template <int VPL>   // 16-byte vectors per lane: C <= 32 * 8 * VPL
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        long long npix, int C, int ldx, int ldy, float eps) {
  const int lane = threadIdx.x & 31;
  const int cvec = C >> 3;
  const float invC = 1.0f / (float)C;
  for (long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); pix < npix; pix += (long long)gridDim.x * 8) {
    float v[VPL][8];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + 32 * k;
      if (vi < cvec) {
        unpack8(ld_nc16(x + pix * ldx + vi * 8), v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += v[k][j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * invC;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (lane + 32 * k < cvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[k][j] - mean;
          sq += d * d;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * invC + eps);          // biased variance, as nn.LayerNorm
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + 32 * k;
      if (vi < cvec) {
        const float4 g0 = reinterpret_cast<const float4*>(gamma + vi * 8)[0], g1 = reinterpret_cast<const float4*>(gamma + vi * 8)[1];
        const float4 b0 = reinterpret_cast<const float4*>(beta + vi * 8)[0], b1 = reinterpret_cast<const float4*>(beta + vi * 8)[1];
        float f[8];
        f[0] = (v[k][0] - mean) * rstd * g0.x + b0.x; f[1] = (v[k][1] - mean) * rstd * g0.y + b0.y;
        f[2] = (v[k][2] - mean) * rstd * g0.z + b0.z; f[3] = (v[k][3] - mean) * rstd * g0.w + b0.w;
        f[4] = (v[k][4] - mean) * rstd * g1.x + b1.x; f[5] = (v[k][5] - mean) * rstd * g1.y + b1.y;
        f[6] = (v[k][6] - mean) * rstd * g1.z + b1.z; f[7] = (v[k][7] - mean) * rstd * g1.w + b1.w;
        st16(y + pix * ldy + vi * 8, pack8(f));
      }
    }
  }
}

// ---- window attention: CTA = (image, window, head); 64 threads, thread i owns query token i ----
// qkv: [N, H, W, ld] bf16, channel = which*C + head*32 + d (the layout nn.Linear(dim, 3*dim) + reshape(B_, N, 3, nH, hd)
// produces).  rel_bias: [heads][64][64] fp32 (table gathered through relative_position_index on the host, once).
// mask: [nW][64][64] fp32 or nullptr.  Tokens that fall into the zero padding have q = k = v = 0 (the reference pads
// AFTER norm1 and qkv has no bias): they still take part in every softmax, exactly as in the reference.
constexpr int kWs = 8, kTok = 64, kHd = 32;

__global__ void __launch_bounds__(kTok) window_attention_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                __nv_bfloat16* __restrict__ out,
                                                                const float* __restrict__ rel_bias,
                                                                const float* __restrict__ mask, int H, int W, int C, int heads,
                                                                int ldq, int ldo, int shift, int nWr, int nWc, float scale) {
  __shared__ float ks[kTok][kHd + 1];
  __shared__ float vs[kTok][kHd + 1];
  int b = blockIdx.x;
  const int head = b % heads;
  b /= heads;
  const int wc = b % nWc;
  b /= nWc;
  const int wr = b % nWr;
  const int n = b / nWr;
  const int R = W, S = H;                 // transposed frame: rows = our x, columns = our y
  const int Rp = nWr * kWs, Sp = nWc * kWs;
  const int i = threadIdx.x;
  const int r_in = i >> 3, s_in = i & 7;
  int r = wr * kWs + r_in + shift, s = wc * kWs + s_in + shift;   // position before the roll(-shift)
  if (r >= Rp) r -= Rp;
  if (s >= Sp) s -= Sp;
  const bool real = r < R && s < S;
  const long long pix = ((long long)n * H + s) * W + r;          // y = s, x = r
  float q[kHd];
  if (real) {
    const __nv_bfloat16* base = qkv + pix * ldq + head * kHd;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f[8];
      unpack8(ld16(base + k * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) q[k * 8 + j] = f[j] * scale;
      unpack8(ld16(base + C + k * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ks[i][k * 8 + j] = f[j];
      unpack8(ld16(base + 2 * C + k * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) vs[i][k * 8 + j] = f[j];
    }
  } else {
#pragma unroll
    for (int d = 0; d < kHd; ++d) {
      q[d] = 0.f;
      ks[i][d] = 0.f;
      vs[i][d] = 0.f;
    }
  }
  __syncthreads();
  const float* bias = rel_bias + ((long long)head * kTok + i) * kTok;
  const float* mrow = mask != nullptr ? mask + (((long long)wr * nWc + wc) * kTok + i) * kTok : nullptr;
  float sc[kTok];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kTok; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < kHd; ++d) a = fmaf(q[d], ks[j][d], a);
    a += bias[j];
    if (mrow != nullptr) a += mrow[j];
    sc[j] = a;
    mx = fmaxf(mx, a);
  }
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < kTok; ++j) {
    sc[j] = __expf(sc[j] - mx);
    den += sc[j];
  }
  const float inv = 1.0f / den;
  float o[kHd];
#pragma unroll
  for (int d = 0; d < kHd; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < kTok; ++j) {
    const float p = sc[j] * inv;
#pragma unroll
    for (int d = 0; d < kHd; ++d) o[d] = fmaf(p, vs[j][d], o[d]);
  }
  if (real) {
    __nv_bfloat16* ob = out + pix * ldo + head * kHd;
#pragma unroll
    for (int k = 0; k < 4; ++k) st16(ob + k * 8, pack8(o + k * 8));
  }
}

}  // namespace dmay

using namespace dmay;

extern "C" {

int dmay_layernorm(const dmay_layernorm_params* p, dmay_stream_t stream) {
  if (!p || !p->x || !p->y || !p->gamma || !p->beta || p->npix <= 0 || p->C <= 0) return DMAY_EINVAL;
  if ((p->C | p->ldx | p->ldy) & 7) return DMAY_EUNSUPPORTED;
  if (!aligned16(p->x) || !aligned16(p->y) || !aligned16(p->gamma) || !aligned16(p->beta)) return DMAY_EINVAL;
  if (p->C > 32 * 8 * 8) return DMAY_EUNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const long long need = (p->npix + 7) / 8;
  const long long cap = (long long)sm_count() * 16;
  const int grid = (int)(need < cap ? need : cap);
  const int vpl = (p->C / 8 + 31) / 32;
#define DMAY_LN(V)                                                                                                    \
  layernorm_kernel<V><<<grid, 256, 0, s>>>((const __nv_bfloat16*)p->x, (__nv_bfloat16*)p->y, (const float*)p->gamma, \
                                           (const float*)p->beta, p->npix, p->C, p->ldx, p->ldy, p->eps)
  if (vpl <= 1) DMAY_LN(1);
  else if (vpl <= 2) DMAY_LN(2);
  else if (vpl <= 4) DMAY_LN(4);
  else DMAY_LN(8);
#undef DMAY_LN
  return finish_launch();
}

int dmay_window_attention(const dmay_winattn_params* p, dmay_stream_t stream) {
  if (!p || !p->qkv || !p->out || !p->rel_bias || p->N <= 0 || p->H <= 0 || p->W <= 0 || p->C <= 0 || p->heads <= 0)
    return DMAY_EINVAL;
  if (p->window != kWs || p->C != p->heads * kHd) return DMAY_EUNSUPPORTED;   // C3STR: window 8, head_dim = 32
  if ((p->ldq | p->ldo) & 7 || p->ldq < 3 * p->C || p->ldo < p->C) return DMAY_EUNSUPPORTED;
  if (!aligned16(p->qkv) || !aligned16(p->out)) return DMAY_EINVAL;
  if (p->shift < 0 || p->shift >= kWs) return DMAY_EINVAL;
  const int nWr = (p->W + kWs - 1) / kWs, nWc = (p->H + kWs - 1) / kWs;       // transposed frame: rows = x, columns = y
  const long long grid = (long long)p->N * nWr * nWc * p->heads;
  if (grid > 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  window_attention_kernel<<<(int)grid, kTok, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)p->qkv, (__nv_bfloat16*)p->out, (const float*)p->rel_bias, (const float*)p->mask, p->H, p->W,
      p->C, p->heads, p->ldq, p->ldo, p->shift, nWr, nWc, p->scale);
  return finish_launch();
}

}  // extern "C"
