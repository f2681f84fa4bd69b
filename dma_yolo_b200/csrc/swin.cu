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

// ---- window attention on mma.sync (bf16 m16n8k16, fp32 accumulate): one warp per (image, window, head) ----------
// A 64-token x 32-dim head is far below a tcgen05 tile (and there are tens of thousands of them), so this uses the
// warp-level tensor path: S = (Q K^T) per 16-query block in registers, softmax on the fragments (rows live in lane
// quads), P re-used as the A operand of P V straight from the accumulator registers (no shared-memory round trip).
// The scalar kernel above stays as the reference implementation (tests compare the two).
constexpr int kAttnWarps = 3;   // 3 x 15 KB of staged q, k, v stays under the 48 KB static limit
constexpr int kPitch = 40;   // bf16 per staged row (80 bytes): ldmatrix rows fall on distinct bank groups

__global__ void __launch_bounds__(kAttnWarps * 32) window_attention_mma_kernel(
    const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, const float* __restrict__ rel_bias,
    const float* __restrict__ mask, int H, int W, int C, int heads, int ldq, int ldo, int shift, int nWr, int nWc,
    float scale, long long items) {
  __shared__ __align__(16) __nv_bfloat16 sm[kAttnWarps][3][kTok][kPitch];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * kAttnWarps + warp;
  if (item >= items) return;                       // whole warp leaves together; no block-level sync below
  long long b = item;
  const int head = (int)(b % heads);
  b /= heads;
  const int wc = (int)(b % nWc);
  b /= nWc;
  const int wr = (int)(b % nWr);
  const int n = (int)(b / nWr);
  const int R = W, S = H, Rp = nWr * kWs, Sp = nWc * kWs;
  // ---- stage q, k, v of the 64 tokens: lane handles tokens lane and lane + 32 ----
  long long pixs[2];
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const int t = lane + 32 * h2;
    int r = wr * kWs + (t >> 3) + shift, s_ = wc * kWs + (t & 7) + shift;
    if (r >= Rp) r -= Rp;
    if (s_ >= Sp) s_ -= Sp;
    const bool real = r < R && s_ < S;
    pixs[h2] = real ? ((long long)n * H + s_) * W + r : -1;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (real) v = ld16(qkv + pixs[h2] * ldq + m * C + head * kHd + k * 8);
        *reinterpret_cast<uint4*>(&sm[warp][m][t][k * 8]) = v;
      }
    }
  }
  __syncwarp();
  const int g = lane >> 2, tq = lane & 3;
  const uint32_t q_base = (uint32_t)__cvta_generic_to_shared(&sm[warp][0][0][0]);
  const uint32_t k_base = (uint32_t)__cvta_generic_to_shared(&sm[warp][1][0][0]);
  const uint32_t v_base = (uint32_t)__cvta_generic_to_shared(&sm[warp][2][0][0]);
  const float* bias_h = rel_bias + (long long)head * kTok * kTok;
  const float* mask_w = mask != nullptr ? mask + ((long long)wr * nWc + wc) * kTok * kTok : nullptr;
#pragma unroll 1
  for (int mt = 0; mt < 4; ++mt) {
    // Q fragments of this 16-row block, both k-steps
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, kof = ks * 16 + (lane >> 4) * 8;
      ldsm_x4(q_base + (uint32_t)(row * kPitch + kof) * 2u, qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
    float sacc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[nt][e] = 0.f;
    }
#pragma unroll
    for (int np = 0; np < 4; ++np) {       // pairs of key tiles (16 keys)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t b0, b1, b2, b3;
        const int krow = np * 16 + (lane & 7) + (lane >> 4) * 8, kof = ks * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(k_base + (uint32_t)(krow * kPitch + kof) * 2u, b0, b1, b2, b3);
        mma_bf16(sacc[2 * np], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b0, b1);
        mma_bf16(sacc[2 * np + 1], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3], b2, b3);
      }
    }
    // scale + relative position bias (+ shift mask), row-wise softmax (rows g and g+8 of the block)
    const int i0 = mt * 16 + g, i1 = i0 + 8;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int j = nt * 8 + tq * 2;
      float2 bb0 = *reinterpret_cast<const float2*>(bias_h + i0 * kTok + j);
      float2 bb1 = *reinterpret_cast<const float2*>(bias_h + i1 * kTok + j);
      if (mask_w != nullptr) {
        const float2 m0 = *reinterpret_cast<const float2*>(mask_w + i0 * kTok + j);
        const float2 m1 = *reinterpret_cast<const float2*>(mask_w + i1 * kTok + j);
        bb0.x += m0.x; bb0.y += m0.y; bb1.x += m1.x; bb1.y += m1.y;
      }
      sacc[nt][0] = fmaf(sacc[nt][0], scale, bb0.x);
      sacc[nt][1] = fmaf(sacc[nt][1], scale, bb0.y);
      sacc[nt][2] = fmaf(sacc[nt][2], scale, bb1.x);
      sacc[nt][3] = fmaf(sacc[nt][3], scale, bb1.y);
      mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      sacc[nt][0] = __expf(sacc[nt][0] - mx0); sacc[nt][1] = __expf(sacc[nt][1] - mx0);
      sacc[nt][2] = __expf(sacc[nt][2] - mx1); sacc[nt][3] = __expf(sacc[nt][3] - mx1);
      d0 += sacc[nt][0] + sacc[nt][1];
      d1 += sacc[nt][2] + sacc[nt][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    const float inv0 = 1.0f / d0, inv1 = 1.0f / d1;
    // O = P V: P from the accumulator registers (bf16), V through transposed ldmatrix
    float oacc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) oacc[nt][e] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {       // 16 keys per step
      const uint32_t a0 = pack_bf2(sacc[2 * kk][0] * inv0, sacc[2 * kk][1] * inv0);
      const uint32_t a1 = pack_bf2(sacc[2 * kk][2] * inv1, sacc[2 * kk][3] * inv1);
      const uint32_t a2 = pack_bf2(sacc[2 * kk + 1][0] * inv0, sacc[2 * kk + 1][1] * inv0);
      const uint32_t a3 = pack_bf2(sacc[2 * kk + 1][2] * inv1, sacc[2 * kk + 1][3] * inv1);
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {     // pairs of 8-dim tiles
        uint32_t b0, b1, b2, b3;
        const int vrow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dof = dp * 16 + (lane >> 4) * 8;
        ldsm_x4_t(v_base + (uint32_t)(vrow * kPitch + dof) * 2u, b0, b1, b2, b3);
        mma_bf16(oacc[2 * dp], a0, a1, a2, a3, b0, b1);
        mma_bf16(oacc[2 * dp + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    // rows i0 / i1 of the window -> pixels; lanes of a quad cover 8 consecutive dims of one 8-dim tile
    const long long p0 = __shfl_sync(0xffffffffu, pixs[i0 >> 5], i0 & 31);
    const long long p1 = __shfl_sync(0xffffffffu, pixs[i1 >> 5], i1 & 31);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int d = head * kHd + nt * 8 + tq * 2;
      if (p0 >= 0) *reinterpret_cast<uint32_t*>(out + p0 * ldo + d) = pack_bf2(oacc[nt][0], oacc[nt][1]);
      if (p1 >= 0) *reinterpret_cast<uint32_t*>(out + p1 * ldo + d) = pack_bf2(oacc[nt][2], oacc[nt][3]);
    }
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
  if (p->variant == 1) {   // scalar fp32 reference kernel
    window_attention_kernel<<<(int)grid, kTok, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)p->qkv, (__nv_bfloat16*)p->out, (const float*)p->rel_bias, (const float*)p->mask, p->H, p->W,
        p->C, p->heads, p->ldq, p->ldo, p->shift, nWr, nWc, p->scale);
  } else {
    const long long blocks = (grid + kAttnWarps - 1) / kAttnWarps;
    window_attention_mma_kernel<<<(int)blocks, kAttnWarps * 32, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)p->qkv, (__nv_bfloat16*)p->out, (const float*)p->rel_bias, (const float*)p->mask, p->H, p->W,
        p->C, p->heads, p->ldq, p->ldo, p->shift, nWr, nWc, p->scale, grid);
  }
  return finish_launch();
}

}  // extern "C"
