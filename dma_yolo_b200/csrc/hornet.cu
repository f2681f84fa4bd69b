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

// thread -> (pixel, channel pair).  Weights are tap-major [49][Cd] fp32 so that a warp reads consecutive floats.
__global__ void __launch_bounds__(256) dwconv7_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int N, int H,
                                                      int W, int Cd, int ldx, int ldy, float scale, DwSeg seg) {
  const int pairs = Cd >> 1;
  const long long items = (long long)N * H * W * pairs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i % pairs) * 2;
    long long pix = i / pairs;
    const int wx = (int)(pix % W);
    const long long t = pix / W;
    const int hy = (int)(t % H);
    const int n = (int)(t / H);
    float a0 = bias[j], a1 = bias[j + 1];
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
      const int yy = hy + dy - 3;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const int xx = wx + dx - 3;
        if (xx < 0 || xx >= W) continue;
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(x + (((long long)n * H + yy) * W + xx) * ldx + j);
        const float2 f = __bfloat1622float2(v);
        const float2 ww = *reinterpret_cast<const float2*>(w + (dy * 7 + dx) * Cd + j);
        a0 = fmaf(f.x, ww.x, a0);
        a1 = fmaf(f.y, ww.y, a1);
      }
    }
    int s = 0;
    while (s + 1 < seg.n && j >= seg.start[s + 1]) ++s;
    const int oc = seg.out[s] + (j - seg.start[s]);
    *reinterpret_cast<__nv_bfloat162*>(y + pix * ldy + oc) = __floats2bfloat162_rn(a0 * scale, a1 * scale);
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
  const long long items = (long long)p->N * p->H * p->W * (p->Cd / 2);
  dwconv7_kernel<<<grid_for(items, 256, 16), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)p->x, (const float*)p->w, (const float*)p->bias, (__nv_bfloat16*)p->y, p->N, p->H, p->W, p->Cd,
      p->ldx, p->ldy, p->scale, seg);
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
