// Memory-bound single-pass kernels over NHWC bf16 activations: space_to_depth, AdConcat (+fused
// nearest upsample), Adapt_Add, Upsample, SCConv gate, AvgPool, layout/input conversion, copy.
// Every kernel moves 128-bit vectors (8 bf16 channels) with a grid-stride loop sized to the SM
// count; pixel rows are 16-byte aligned by the ld/offset contract in include/dmayolo.h.
#include "common.cuh"

namespace dmay {

std::atomic<long long> g_launches{0};

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    n = v;
  }
  return n;
}

constexpr int kThreads = 256;

// Thread layout of the NHWC streaming kernels: blockDim = (VX, PY) with VX = channel vectors (16 B each)
// handled side by side (a power of two <= 32) and PY = 256/VX pixels per CTA step.  Pixels are indexed
// with 32-bit integers (host checks N*H*W < 2^31) and decoded into (n,h,w) once per pixel, so the
// per-vector cost is one add — 64-bit div/mod per 16-byte item made the first version ALU-bound.
struct PixGeom {
  unsigned npix;
  int H, W;
};
__device__ __forceinline__ void decode_pix(unsigned pix, int H, int W, int& n, int& h, int& w) {
  const unsigned t = pix / (unsigned)W;
  w = (int)(pix - t * (unsigned)W);
  n = (int)(t / (unsigned)H);
  h = (int)(t - (unsigned)n * (unsigned)H);
}
#define FOR_EACH_PIXEL(pix, npix) \
  for (unsigned pix = blockIdx.x * blockDim.y + threadIdx.y; pix < (npix); pix += gridDim.x * blockDim.y)
#define FOR_EACH_VEC(v, cv) for (int v = threadIdx.x; v < (cv); v += blockDim.x)

static inline dim3 pix_block(int cv) {
  int vx = 1;
  while (vx * 2 <= cv && vx < 32) vx *= 2;
  return dim3(vx, kThreads / vx);
}
static inline int pix_grid(long long npix, dim3 block, int ctas_per_sm = 8) {
  long long need = (npix + block.y - 1) / block.y;
  long long cap = (long long)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- a4 space_to_depth ----------------------------------------------------------------------
// reference: models/common.py:1457-1458.  One item = one 16-byte vector of the OUTPUT.
__global__ void __launch_bounds__(kThreads) spd_kernel(const __nv_bfloat16* __restrict__ x,
                                                       __nv_bfloat16* __restrict__ y, unsigned npix, int H, int W,
                                                       int C, int ldx, int ldy) {
  const int Ho = H >> 1, Wo = W >> 1, cv = C >> 3;
  FOR_EACH_PIXEL(pix, npix) {  // output pixel
    int n, ho, wo;
    decode_pix(pix, Ho, Wo, n, ho, wo);
    const __nv_bfloat16* src0 = x + (((long long)n * H + 2 * ho) * W + 2 * wo) * ldx;
    __nv_bfloat16* dst0 = y + (long long)pix * ldy;
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // q = dy + 2*dx
      const __nv_bfloat16* src = src0 + ((long long)(q & 1) * W + (q >> 1)) * ldx;
      FOR_EACH_VEC(v, cv) st_na16(dst0 + q * C + v * 8, ld_nc16(src + v * 8));
    }
  }
}

// ---- a6 AdConcat2/3 (+ nearest upsample of any input), Concat ---------------------------------
// reference: models/common.py:1003-1008, 1021-1026 (weights normalised by the caller).
struct CatArgs {
  const __nv_bfloat16* x[3];
  int C[3];
  int ld[3];
  int up[3];
  float w[3];
};
__global__ void __launch_bounds__(kThreads) adconcat_kernel(CatArgs a, __nv_bfloat16* __restrict__ y, int n_in,
                                                            unsigned npix, int H, int W, int ldy) {
  FOR_EACH_PIXEL(pix, npix) {
    int n, h_, w_;
    decode_pix(pix, H, W, n, h_, w_);
    int coff = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < n_in) {
        const int u = a.up[k];
        const __nv_bfloat16* src =
            a.x[k] + (((long long)n * (H >> u) + (h_ >> u)) * (W >> u) + (w_ >> u)) * a.ld[k];
        __nv_bfloat16* dst = y + (long long)pix * ldy + coff;
        const float wk = a.w[k];
        const int cv = a.C[k] >> 3;
        FOR_EACH_VEC(v, cv) {
          float f[8];
          unpack8(u ? ld16(src + v * 8) : ld_nc16(src + v * 8), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __fmul_rn(wk, f[j]);
          st_na16(dst + v * 8, pack8(f));
        }
        coff += a.C[k];
      }
    }
  }
}

// ---- Adapt_Add2/3: y = silu(sum_i w_i x_i), models/common.py:1040-1061 -------------------------
__global__ void __launch_bounds__(kThreads) adaptadd_kernel(CatArgs a, __nv_bfloat16* __restrict__ y, int n_in,
                                                            unsigned npix, int C, int ldy) {
  const int cv = C >> 3;
  FOR_EACH_PIXEL(pix, npix) {
    FOR_EACH_VEC(v, cv) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int k = 0; k < n_in; ++k) {
        float f[8];
        unpack8(ld_nc16(a.x[k] + (long long)pix * a.ld[k] + v * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += a.w[k] * f[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = acc[j] * sigmoid_acc(acc[j]);
      st_na16(y + (long long)pix * ldy + v * 8, pack8(acc));
    }
  }
}

// ---- nn.Upsample(None, f, 'nearest') -----------------------------------------------------------
__global__ void __launch_bounds__(kThreads) upsample_kernel(const __nv_bfloat16* __restrict__ x,
                                                            __nv_bfloat16* __restrict__ y, unsigned npix, int H, int W,
                                                            int C, int ldx, int ldy, int f) {
  // one item = one SOURCE pixel: a 16-byte vector is loaded once and stored f * f times (the output-pixel form issued one load
  // per store and two integer divisions per 16 bytes: 3.1 TB/s on the two replicate launches of cfg-2)
  const int Wo = W * f, cv = C >> 3;
  FOR_EACH_PIXEL(pix, npix) {  // source pixel
    int n, h, w;
    decode_pix(pix, H, W, n, h, w);
    const __nv_bfloat16* src = x + (long long)pix * ldx;
    __nv_bfloat16* dst = y + (((long long)n * H + h) * f * Wo + (long long)w * f) * ldy;
    FOR_EACH_VEC(v, cv) {
      const uint4 u = ld_nc16(src + v * 8);
      for (int dy = 0; dy < f; ++dy)
        for (int dx = 0; dx < f; ++dx) st_na16(dst + ((long long)dy * Wo + dx) * ldy + v * 8, u);
    }
  }
}

// ---- a5 SCConv: AvgPool2d(r,r) and the calibration gate ------------------------------------------
// reference: models/common.py:1281-1287 (k2 pooling), 1310-1314 (gate).
__global__ void __launch_bounds__(kThreads) avgpool_kernel(const __nv_bfloat16* __restrict__ x,
                                                           __nv_bfloat16* __restrict__ y, unsigned npix, int H, int W,
                                                           int C, int ldx, int ldy, int r) {
  const int Ho = H / r, Wo = W / r, cv = C >> 3;
  const float inv = 1.0f / (float)(r * r);
  FOR_EACH_PIXEL(pix, npix) {  // output pixel
    int n, ho, wo;
    decode_pix(pix, Ho, Wo, n, ho, wo);
    const __nv_bfloat16* src0 = x + (((long long)n * H + ho * r) * W + (long long)wo * r) * ldx;
    FOR_EACH_VEC(v, cv) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int dy = 0; dy < r; ++dy) {
        const __nv_bfloat16* row = src0 + (long long)dy * W * ldx + v * 8;
#pragma unroll 4
        for (int dx = 0; dx < r; ++dx) {
          float f[8];
          unpack8(ld_nc16(row + (long long)dx * ldx), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= inv;
      st16(y + (long long)pix * ldy + v * 8, pack8(acc));
    }
  }
}

__device__ __forceinline__ float tanh_apx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads) scgate_kernel(const __nv_bfloat16* __restrict__ x,
                                                          const __nv_bfloat16* __restrict__ k3,
                                                          const __nv_bfloat16* __restrict__ k2,
                                                          __nv_bfloat16* __restrict__ y, unsigned npix, int H, int W, int C,
                                                          int Hk, int Wk, int ldx, int ld3, int ld2, int ldy) {
  const int cv = C >> 3;
  const float sh = (float)Hk / (float)H, sw = (float)Wk / (float)W;
  FOR_EACH_PIXEL(pix, npix) {
    int n, h_, w_;
    decode_pix(pix, H, W, n, h_, w_);
    const int hs = nearest_src(h_, Hk, H, sh), ws = nearest_src(w_, Wk, W, sw);
    const __nv_bfloat16* k2p = k2 + (((long long)n * Hk + hs) * Wk + ws) * ld2;
    FOR_EACH_VEC(v, cv) {
      float fx[8], f3[8], f2[8];
      const uint4 ux = ld_nc16(x + (long long)pix * ldx + v * 8);
      const uint4 u3 = ld_nc16(k3 + (long long)pix * ld3 + v * 8);
      const uint4 u2 = ld16(k2p + v * 8);
      unpack8(ux, fx);
      unpack8(u3, f3);
      unpack8(u2, f2);
#pragma unroll
      for (int j = 0; j < 8; ++j) f3[j] *= fmaf(0.5f, tanh_apx(0.5f * (fx[j] + f2[j])), 0.5f);  // sigmoid
      st_na16(y + (long long)pix * ldy + v * 8, pack8(f3));
    }
  }
}

// ---- layout glue -------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f<unsigned char>(unsigned char v) { return (float)v; }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// 32x32 smem-tiled transpose between [N, HW, ld] (NHWC slice) and [N, C, HW] (NCHW).
template <typename T, int DIR>
__global__ void __launch_bounds__(256) layout_kernel(const void* __restrict__ xin, void* __restrict__ yout, int C,
                                                     int HW, int ld) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (DIR == 0) {  // nhwc(bf16) -> nchw(T)
    const __nv_bfloat16* x = (const __nv_bfloat16*)xin;
    T* y = (T*)yout;
    for (int r = ty; r < 32; r += 8) {
      int p = p0 + r, c = c0 + tx;
      tile[r][tx] = (p < HW && c < C) ? __bfloat162float(x[((long long)n * HW + p) * ld + c]) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      int c = c0 + r, p = p0 + tx;
      if (p < HW && c < C) y[((long long)n * C + c) * HW + p] = from_f<T>(tile[tx][r]);
    }
  } else {  // nchw(T) -> nhwc(bf16)
    const T* x = (const T*)xin;
    __nv_bfloat16* y = (__nv_bfloat16*)yout;
    for (int r = ty; r < 32; r += 8) {
      int c = c0 + r, p = p0 + tx;
      tile[r][tx] = (p < HW && c < C) ? to_f<T>(x[((long long)n * C + c) * HW + p]) : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      int p = p0 + r, c = c0 + tx;
      if (p < HW && c < C) y[((long long)n * HW + p) * ld + c] = __float2bfloat16_rn(tile[tx][r]);
    }
  }
}

// NCHW image (C small) -> NHWC bf16 with channel padding, optionally 2x2 pixel-unshuffled.
// One thread per output pixel; reads are coalesced per input channel plane.
template <typename T>
__global__ void __launch_bounds__(kThreads) prep_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                        int N, int C, int H, int W, int Cpad, int spd, float mul) {
  const int Ho = spd ? H >> 1 : H, Wo = spd ? W >> 1 : W;
  const long long items = (long long)N * Ho * Wo;
  const long long plane = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    int wo = (int)(i % Wo);
    long long t = i / Wo;
    int ho = (int)(t % Ho);
    int n = (int)(t / Ho);
    __nv_bfloat16* dst = y + i * Cpad;
    const T* src = x + (long long)n * C * plane;
    int cnt = 0;
    if (Cpad == 16 && C == 3) {  // RGB stem fast path: build the 32-byte pixel in registers, two 16-byte stores
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
      if (spd) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int dy = q & 1, dx = q >> 1;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            v[q * 3 + c] = to_f<T>(src[c * plane + (long long)(2 * ho + dy) * W + (2 * wo + dx)]) * mul;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = to_f<T>(src[c * plane + (long long)ho * W + wo]) * mul;
      }
      st_na16(dst, pack8(v));
      st_na16(dst + 8, pack8(v + 8));
      continue;
    }
    if (spd) {
      for (int q = 0; q < 4; ++q) {
        int dy = q & 1, dx = q >> 1;
        for (int c = 0; c < C; ++c, ++cnt)
          dst[cnt] = __float2bfloat16_rn(to_f<T>(src[c * plane + (long long)(2 * ho + dy) * W + (2 * wo + dx)]) * mul);
      }
    } else {
      for (int c = 0; c < C; ++c, ++cnt)
        dst[cnt] = __float2bfloat16_rn(to_f<T>(src[c * plane + (long long)ho * W + wo]) * mul);
    }
    for (; cnt < Cpad; ++cnt) dst[cnt] = __float2bfloat16_rn(0.f);
  }
}

__global__ void __launch_bounds__(kThreads) copy_kernel(const uint4* __restrict__ s, uint4* __restrict__ d,
                                                        long long n16) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
       i += (long long)gridDim.x * blockDim.x)
    st_na16(d + i, ld_nc16(s + i));
}

}  // namespace dmay

using namespace dmay;

extern "C" {

int dmay_version(void) { return 100; }
long long dmay_launch_count(void) { return g_launches.load(); }
const char* dmay_strerror(int code) {
  switch (code) {
    case DMAY_OK: return "ok";
    case DMAY_EINVAL: return "invalid argument (null/misaligned pointer or non-positive size)";
    case DMAY_EUNSUPPORTED: return "unsupported shape for this kernel";
    case DMAY_EDRIVER: return "CUDA driver entry point (cuTensorMapEncode*) unavailable or failed";
    case DMAY_ETOOBIG: return "workspace or capacity too small";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

#define REQ(cond)                 \
  do {                            \
    if (!(cond)) return DMAY_EINVAL; \
  } while (0)

int dmay_spd(const dmay_spd_params* p, dmay_stream_t stream) {
  REQ(p && p->x && p->y && p->N > 0 && p->H > 0 && p->W > 0 && p->C > 0);
  REQ(aligned16(p->x) && aligned16(p->y));
  if ((p->H | p->W) & 1) return DMAY_EUNSUPPORTED;
  if ((p->C | p->ldx | p->ldy) & 7) return DMAY_EUNSUPPORTED;
  const long long npix = (long long)p->N * (p->H / 2) * (p->W / 2);
  if (npix >= 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  const dim3 blk = pix_block(p->C / 8);
  spd_kernel<<<pix_grid(npix, blk), blk, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)p->x, (__nv_bfloat16*)p->y, (unsigned)npix, p->H, p->W, p->C, p->ldx, p->ldy);
  return finish_launch();
}

int dmay_adconcat(const dmay_adconcat_params* p, dmay_stream_t stream) {
  REQ(p && p->x0 && p->x1 && p->y && p->N > 0 && p->H > 0 && p->W > 0);
  REQ(p->n_in == 2 || (p->n_in == 3 && p->x2));
  CatArgs a;
  a.x[0] = (const __nv_bfloat16*)p->x0; a.x[1] = (const __nv_bfloat16*)p->x1; a.x[2] = (const __nv_bfloat16*)p->x2;
  a.C[0] = p->C0; a.C[1] = p->C1; a.C[2] = p->n_in > 2 ? p->C2 : 0;
  a.ld[0] = p->ld0; a.ld[1] = p->ld1; a.ld[2] = p->n_in > 2 ? p->ld2 : 8;
  a.up[0] = p->up0; a.up[1] = p->up1; a.up[2] = p->n_in > 2 ? p->up2 : 0;
  a.w[0] = p->w0; a.w[1] = p->w1; a.w[2] = p->w2;
  int ctot = 0;
  for (int i = 0; i < p->n_in; ++i) {
    REQ(a.C[i] > 0 && aligned16(a.x[i]));
    if ((a.C[i] | a.ld[i]) & 7) return DMAY_EUNSUPPORTED;
    if (a.up[i] < 0 || a.up[i] > 3) return DMAY_EUNSUPPORTED;
    if ((p->H & ((1 << a.up[i]) - 1)) || (p->W & ((1 << a.up[i]) - 1))) return DMAY_EUNSUPPORTED;
    ctot += a.C[i];
  }
  REQ(aligned16(p->y));
  if (p->ldy & 7) return DMAY_EUNSUPPORTED;
  const long long npix = (long long)p->N * p->H * p->W;
  if (npix >= 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  int cmin = a.C[0];
  for (int i = 1; i < p->n_in; ++i) cmin = a.C[i] < cmin ? a.C[i] : cmin;
  (void)ctot;
  const dim3 blk = pix_block(cmin / 8);
  adconcat_kernel<<<pix_grid(npix, blk), blk, 0, (cudaStream_t)stream>>>(
      a, (__nv_bfloat16*)p->y, p->n_in, (unsigned)npix, p->H, p->W, p->ldy);
  return finish_launch();
}

int dmay_adaptadd(const dmay_adaptadd_params* p, dmay_stream_t stream) {
  REQ(p && p->x0 && p->x1 && p->y && p->npix > 0 && p->C > 0);
  REQ(p->n_in == 2 || (p->n_in == 3 && p->x2));
  if ((p->C | p->ld0 | p->ld1 | p->ldy) & 7) return DMAY_EUNSUPPORTED;
  if (p->n_in == 3 && (p->ld2 & 7)) return DMAY_EUNSUPPORTED;
  CatArgs a;
  a.x[0] = (const __nv_bfloat16*)p->x0; a.x[1] = (const __nv_bfloat16*)p->x1; a.x[2] = (const __nv_bfloat16*)p->x2;
  a.C[0] = a.C[1] = a.C[2] = p->C;
  a.ld[0] = p->ld0; a.ld[1] = p->ld1; a.ld[2] = p->ld2;
  a.up[0] = a.up[1] = a.up[2] = 0;
  a.w[0] = p->w0; a.w[1] = p->w1; a.w[2] = p->w2;
  REQ(aligned16(p->x0) && aligned16(p->x1) && aligned16(p->y) && (p->n_in < 3 || aligned16(p->x2)));
  if (p->npix >= 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  const dim3 blk = pix_block(p->C / 8);
  adaptadd_kernel<<<pix_grid(p->npix, blk), blk, 0, (cudaStream_t)stream>>>(
      a, (__nv_bfloat16*)p->y, p->n_in, (unsigned)p->npix, p->C, p->ldy);
  return finish_launch();
}

int dmay_upsample_nearest(const dmay_upsample_params* p, dmay_stream_t stream) {
  REQ(p && p->x && p->y && p->N > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->factor >= 1);
  REQ(aligned16(p->x) && aligned16(p->y));
  if ((p->C | p->ldx | p->ldy) & 7) return DMAY_EUNSUPPORTED;
  if ((long long)p->N * p->H * p->factor * p->W * p->factor >= 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  const long long npix = (long long)p->N * p->H * p->W;   // source pixels
  const dim3 blk = pix_block(p->C / 8);
  upsample_kernel<<<pix_grid(npix, blk), blk, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)p->x, (__nv_bfloat16*)p->y, (unsigned)npix, p->H, p->W, p->C, p->ldx, p->ldy, p->factor);
  return finish_launch();
}

int dmay_avgpool(const dmay_avgpool_params* p, dmay_stream_t stream) {
  REQ(p && p->x && p->y && p->N > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->r >= 1);
  REQ(aligned16(p->x) && aligned16(p->y));
  if ((p->C | p->ldx | p->ldy) & 7) return DMAY_EUNSUPPORTED;
  if (p->H / p->r < 1 || p->W / p->r < 1) return DMAY_EUNSUPPORTED;
  const long long npix = (long long)p->N * (p->H / p->r) * (p->W / p->r);
  if (npix >= 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  const dim3 blk = pix_block(p->C / 8);
  avgpool_kernel<<<pix_grid(npix, blk), blk, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)p->x, (__nv_bfloat16*)p->y, (unsigned)npix, p->H, p->W, p->C, p->ldx, p->ldy, p->r);
  return finish_launch();
}

int dmay_scconv_gate(const dmay_scgate_params* p, dmay_stream_t stream) {
  REQ(p && p->x && p->k3 && p->k2 && p->y && p->N > 0 && p->H > 0 && p->W > 0 && p->C > 0 && p->Hk > 0 && p->Wk > 0);
  REQ(aligned16(p->x) && aligned16(p->k3) && aligned16(p->k2) && aligned16(p->y));
  if ((p->C | p->ldx | p->ld3 | p->ld2 | p->ldy) & 7) return DMAY_EUNSUPPORTED;
  const long long npix = (long long)p->N * p->H * p->W;
  if (npix >= 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  const dim3 blk = pix_block(p->C / 8);
  scgate_kernel<<<pix_grid(npix, blk), blk, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)p->x, (const __nv_bfloat16*)p->k3, (const __nv_bfloat16*)p->k2, (__nv_bfloat16*)p->y,
      (unsigned)npix, p->H, p->W, p->C, p->Hk, p->Wk, p->ldx, p->ld3, p->ld2, p->ldy);
  return finish_launch();
}

int dmay_layout_convert(const dmay_layout_params* p, dmay_stream_t stream) {
  REQ(p && p->x && p->y && p->N > 0 && p->C > 0 && p->H > 0 && p->W > 0 && p->ld >= p->C);
  const int HW = p->H * p->W;
  dim3 grid((HW + 31) / 32, (p->C + 31) / 32, p->N);
  if (p->N > 65535 || grid.y > 65535) return DMAY_EUNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH(T)                                                                     \
  do {                                                                                \
    if (p->dir == 0) layout_kernel<T, 0><<<grid, 256, 0, s>>>(p->x, p->y, p->C, HW, p->ld); \
    else layout_kernel<T, 1><<<grid, 256, 0, s>>>(p->x, p->y, p->C, HW, p->ld);       \
  } while (0)
  switch (p->dtype) {
    case DMAY_DT_F32: LAUNCH(float); break;
    case DMAY_DT_BF16: LAUNCH(__nv_bfloat16); break;
    case DMAY_DT_F16: LAUNCH(__half); break;
    default: return DMAY_EUNSUPPORTED;
  }
#undef LAUNCH
  return finish_launch();
}

int dmay_input_prep(const dmay_prep_params* p, dmay_stream_t stream) {
  REQ(p && p->x && p->y && p->N > 0 && p->C > 0 && p->H > 0 && p->W > 0);
  const int creal = p->spd ? 4 * p->C : p->C;
  if (p->Cpad < creal || (p->Cpad & 7)) return DMAY_EUNSUPPORTED;
  if (p->spd && ((p->H | p->W) & 1)) return DMAY_EUNSUPPORTED;
  long long items = (long long)p->N * (p->spd ? p->H / 2 : p->H) * (p->spd ? p->W / 2 : p->W);
  int grid = grid_for(items, kThreads);
  cudaStream_t s = (cudaStream_t)stream;
  __nv_bfloat16* y = (__nv_bfloat16*)p->y;
  switch (p->in_dtype) {
    case DMAY_DT_F32:
      prep_kernel<float><<<grid, kThreads, 0, s>>>((const float*)p->x, y, p->N, p->C, p->H, p->W, p->Cpad, p->spd, p->mul);
      break;
    case DMAY_DT_BF16:
      prep_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>((const __nv_bfloat16*)p->x, y, p->N, p->C, p->H, p->W, p->Cpad, p->spd, p->mul);
      break;
    case DMAY_DT_F16:
      prep_kernel<__half><<<grid, kThreads, 0, s>>>((const __half*)p->x, y, p->N, p->C, p->H, p->W, p->Cpad, p->spd, p->mul);
      break;
    case DMAY_DT_U8:
      prep_kernel<unsigned char><<<grid, kThreads, 0, s>>>((const unsigned char*)p->x, y, p->N, p->C, p->H, p->W, p->Cpad, p->spd, p->mul);
      break;
    default: return DMAY_EUNSUPPORTED;
  }
  return finish_launch();
}

int dmay_copy(const dmay_copy_params* p, dmay_stream_t stream) {
  REQ(p && p->src && p->dst && p->bytes > 0 && (p->bytes & 15) == 0 && aligned16(p->src) && aligned16(p->dst));
  long long n16 = p->bytes / 16;
  copy_kernel<<<grid_for(n16, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>((const uint4*)p->src, (uint4*)p->dst, n16);
  return finish_launch();
}

}  // extern "C"
