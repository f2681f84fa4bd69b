// 8f-2: HorBlock / GnConv (models/common.py:1318-1426; C3HB in spdconv.yaml) on NHWC bf16 — the pieces that are not
// GEMMs.  proj_in, the pws chain, proj_out and the MLP run on the tcgen05 implicit-GEMM kernel (the recursive gating
// x_{i+1} = pws_i(x_i) * dw_{i+1} is its EPI_LINEAR_MUL epilogue); what is left is
//   dwconv7        : the 7x7 depth-wise convolution over the `abc` part of proj_in's output (+ bias, * scale).  Its
//                    output is written segment by segment (one segment per gating order) at 8-channel-aligned
//                    offsets, so that every segment is addressable with 16-byte vectors / TMA boxes even when the
//                    reference's split points are not (dims = c/16, c/8, ... : 4, 8, 16, 32, 64 for c = 64);
//   mul_channels   : x_0 = pwa * dw_0 (first gating step, d0 = c/16 channels, zero-padded to the next conv's K);
//   axpy_channels  : x + gamma[c] * y (layer scale + residual after proj_out, whose SiLU sits between the GEMM and gamma).
#include "common.cuh"

namespace dmay {

struct DwSeg {
  int n;
  int start[8];   // first abc channel of the segment
  int out[8];     // channel offset of the segment in y
};

// CTA = 8x8 output pixels x 64 channels of one image.  The 14x14x64 input halo (25 KB bf16) and the 49x64 weights
// (12.5 KB fp32) are staged in shared memory once; warp = tile row, lane = channel pair, and a thread slides over its
// row's 8 pixels: per kernel row it reads 14 inputs + 7 weight pairs and issues 112 FMAs, so every input element is
// fetched from HBM/L2 once instead of 49 times and the loop is FMA-bound.  Weights are tap-major [49][Cd] fp32.
constexpr int kDwT = 8, kDwC = 64, kDwHalo = kDwT + 6;

__global__ void __launch_bounds__(256) dwconv7_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int N, int H,
                                                      int W, int Cd, int ldx, int ldy, float scale, DwSeg seg, int tiles_x,
                                                      int tiles_y) {
  __shared__ __nv_bfloat162 in_s[kDwHalo][kDwHalo][kDwC / 2];
  __shared__ float2 w_s[49][kDwC / 2];
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int n = t / tiles_y;
  const int cbase = blockIdx.y * kDwC;
  const int x0 = tx * kDwT - 3, y0 = ty * kDwT - 3;
  for (int i = threadIdx.x; i < kDwHalo * kDwHalo * (kDwC / 2); i += blockDim.x) {
    const int pr = i & (kDwC / 2 - 1);
    const int p = i >> 5;
    const int px = p % kDwHalo, py = p / kDwHalo;
    const int xx = x0 + px, yy = y0 + py, j = cbase + 2 * pr;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
    if (xx >= 0 && xx < W && yy >= 0 && yy < H && j < Cd)
      v = *reinterpret_cast<const __nv_bfloat162*>(x + (((long long)n * H + yy) * W + xx) * ldx + j);
    in_s[py][px][pr] = v;
  }
  for (int i = threadIdx.x; i < 49 * (kDwC / 2); i += blockDim.x) {
    const int pr = i & (kDwC / 2 - 1), tap = i >> 5;
    const int j = cbase + 2 * pr;
    w_s[tap][pr] = j < Cd ? *reinterpret_cast<const float2*>(w + tap * Cd + j) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = cbase + 2 * lane;
  if (j >= Cd) return;
  // the channel pair of a thread is one packed fp32x2 value: 392 FFMA2 per thread instead of 784 FFMA (same IEEE roundings):
  // cfg-4a 13.5 -> 11.8 ms per step for the eight launches.  (Staging the halo as fp32 pairs to drop the per-row bf16 -> fp32
  // conversions was measured SLOWER, 13.6 ms: twice the shared-memory bytes per load and half the CTAs per SM.)
  f32x2_t acc2[kDwT];
  const float2 b2 = *reinterpret_cast<const float2*>(bias + j);
#pragma unroll
  for (int px = 0; px < kDwT; ++px) acc2[px] = f2_pack(b2.x, b2.y);
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    f32x2_t v[kDwHalo];
#pragma unroll
    for (int c = 0; c < kDwHalo; ++c) v[c] = f2_from_bf2(*reinterpret_cast<const uint32_t*>(&in_s[r + dy][c][lane]));
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) {
      const f32x2_t ww = *reinterpret_cast<const f32x2_t*>(&w_s[dy * 7 + dx][lane]);
#pragma unroll
      for (int px = 0; px < kDwT; ++px) acc2[px] = f2_fma(v[px + dx], ww, acc2[px]);
    }
  }
  float2 acc[kDwT];
#pragma unroll
  for (int px = 0; px < kDwT; ++px) f2_unpack(acc2[px], acc[px].x, acc[px].y);
  int s = 0;
  while (s + 1 < seg.n && j >= seg.start[s + 1]) ++s;
  const int oc = seg.out[s] + (j - seg.start[s]);
  const int oy = ty * kDwT + r;
  if (oy >= H) return;
#pragma unroll
  for (int px = 0; px < kDwT; ++px) {
    const int ox = tx * kDwT + px;
    if (ox < W)
      *reinterpret_cast<__nv_bfloat162*>(y + (((long long)n * H + oy) * W + ox) * ldy + oc) =
          __floats2bfloat162_rn(acc[px].x * scale, acc[px].y * scale);
  }
}

__global__ void __launch_bounds__(256) mul_channels_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                           __nv_bfloat16* __restrict__ y, long long npix, int d, int lda, int ldb,
                                                           int ldy) {
  const int pairs = d >> 1;
  const long long items = npix * pairs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % pairs) * 2;
    const long long pix = i / pairs;
    const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a + pix * lda + j));
    const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(b + pix * ldb + j));
    *reinterpret_cast<__nv_bfloat162*>(y + pix * ldy + j) = __floats2bfloat162_rn(fa.x * fb.x, fa.y * fb.y);
  }
}

__global__ void __launch_bounds__(256) axpy_channels_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g,
                                                            const float* __restrict__ gamma, __nv_bfloat16* __restrict__ y,
                                                            long long npix, int C, int ldx, int ldg, int ldy) {
  const int cv = C >> 3;
  const long long items = npix * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    const long long pix = i / cv;
    float fx[8], fg[8];
    unpack8(ld_nc16(x + pix * ldx + v * 8), fx);
    unpack8(ld_nc16(g + pix * ldg + v * 8), fg);
    const float4 g0 = reinterpret_cast<const float4*>(gamma + v * 8)[0], g1 = reinterpret_cast<const float4*>(gamma + v * 8)[1];
    fx[0] = fmaf(g0.x, fg[0], fx[0]); fx[1] = fmaf(g0.y, fg[1], fx[1]); fx[2] = fmaf(g0.z, fg[2], fx[2]); fx[3] = fmaf(g0.w, fg[3], fx[3]);
    fx[4] = fmaf(g1.x, fg[4], fx[4]); fx[5] = fmaf(g1.y, fg[5], fx[5]); fx[6] = fmaf(g1.z, fg[6], fx[6]); fx[7] = fmaf(g1.w, fg[7], fx[7]);
    st16(y + pix * ldy + v * 8, pack8(fx));
  }
}

}  // namespace dmay

using namespace dmay;

extern "C" {

int dmay_dwconv7(const dmay_dwconv7_params* p, dmay_stream_t stream) {
  if (!p || !p->x || !p->w || !p->bias || !p->y || p->N <= 0 || p->H <= 0 || p->W <= 0 || p->Cd <= 0) return DMAY_EINVAL;
  if ((p->Cd & 1) || (p->ldx & 1) || (p->ldy & 1) || p->n_seg < 1 || p->n_seg > 5) return DMAY_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(p->x) & 3u) || (reinterpret_cast<uintptr_t>(p->y) & 3u) || (reinterpret_cast<uintptr_t>(p->w) & 7u))
    return DMAY_EINVAL;
  DwSeg seg;
  seg.n = p->n_seg;
  const int st[5] = {p->seg_start0, p->seg_start1, p->seg_start2, p->seg_start3, p->seg_start4};
  const int ou[5] = {p->seg_out0, p->seg_out1, p->seg_out2, p->seg_out3, p->seg_out4};
  for (int i = 0; i < 8; ++i) {
    seg.start[i] = i < 5 ? st[i] : 0;
    seg.out[i] = i < 5 ? ou[i] : 0;
  }
  for (int i = 0; i < p->n_seg; ++i)
    if ((seg.start[i] & 1) || (seg.out[i] & 1) || seg.start[i] < 0 || seg.start[i] >= p->Cd) return DMAY_EINVAL;
  const int tiles_x = (p->W + kDwT - 1) / kDwT, tiles_y = (p->H + kDwT - 1) / kDwT;
  const long long gx = (long long)p->N * tiles_x * tiles_y;
  if (gx > 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  dwconv7_kernel<<<dim3((unsigned)gx, (unsigned)((p->Cd + kDwC - 1) / kDwC)), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)p->x, (const float*)p->w, (const float*)p->bias, (__nv_bfloat16*)p->y, p->N, p->H, p->W, p->Cd,
      p->ldx, p->ldy, p->scale, seg, tiles_x, tiles_y);
  return finish_launch();
}

int dmay_mul_channels(const dmay_mulch_params* p, dmay_stream_t stream) {
  if (!p || !p->a || !p->b || !p->y || p->npix <= 0 || p->d <= 0) return DMAY_EINVAL;
  if ((p->d | p->lda | p->ldb | p->ldy) & 1) return DMAY_EUNSUPPORTED;
  mul_channels_kernel<<<grid_for(p->npix * (p->d / 2), 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)p->a, (const __nv_bfloat16*)p->b, (__nv_bfloat16*)p->y, p->npix, p->d, p->lda, p->ldb, p->ldy);
  return finish_launch();
}

int dmay_axpy_channels(const dmay_axpych_params* p, dmay_stream_t stream) {
  if (!p || !p->x || !p->g || !p->gamma || !p->y || p->npix <= 0 || p->C <= 0) return DMAY_EINVAL;
  if ((p->C | p->ldx | p->ldg | p->ldy) & 7) return DMAY_EUNSUPPORTED;
  if (!aligned16(p->x) || !aligned16(p->g) || !aligned16(p->y) || !aligned16(p->gamma)) return DMAY_EINVAL;
  axpy_channels_kernel<<<grid_for(p->npix * (p->C / 8), 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)p->x, (const __nv_bfloat16*)p->g, (const float*)p->gamma, (__nv_bfloat16*)p->y, p->npix, p->C,
      p->ldx, p->ldg, p->ldy);
  return finish_launch();
}

}  // extern "C"
