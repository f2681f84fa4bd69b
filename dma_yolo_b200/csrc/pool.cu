// a7: cascaded stride-1 max-pools of SPPF / SPPFCSPC (models/common.py:1272-1274, 252-258) and the
// single k x k stride-1 max-pool of SPP / SPPCSPC (models/common.py:212-227, 1237-1255).
//
// One CTA owns one image x one group of CB channels and keeps the whole H x W x CB plane in shared
// memory (20x20x64 bf16 = 51 KB at cfg-2): x is read from HBM exactly once and y1..y3 are written
// exactly once — the kernel's traffic is the algorithmic (1 in + 3 out).  Each pool is done
// separably (row max then column max) on packed bf16x2 with __hmax2, which is bit-exact for a max.
#include "common.cuh"

namespace dmay {

__device__ __forceinline__ uint4 max16(const uint4& a, const uint4& b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// smem: bufA[H*Wp*VL] + bufB[H*Wp*VL] uint4 (VL = CB/8 vectors per pixel, Wp = W|1: an odd row pitch keeps the
// row-walking threads of a warp on different banks).
// K5 = true: k == 5 with a rotating 5-element register window — one shared-memory read and one write per
// element per pass (the straightforward form re-reads each element k times).
template <int STAGES, bool K5>
__global__ void __launch_bounds__(256) pool_plane_kernel(const __nv_bfloat16* __restrict__ x,
                                                         __nv_bfloat16* __restrict__ y1, __nv_bfloat16* __restrict__ y2,
                                                         __nv_bfloat16* __restrict__ y3, int H, int W, int C, int ldx,
                                                         int ldy, int k, int VL) {
  extern __shared__ uint4 smem[];
  const int HW = H * W;
  const int Wp = W | 1;
  uint4* A = smem;
  uint4* B = smem + (size_t)H * Wp * VL;
  const int groups = (C / 8 + VL - 1) / VL;
  const int n = blockIdx.x / groups, g = blockIdx.x % groups;
  const int v0 = g * VL;                              // first channel-vector of this CTA
  const int vl = min(VL, C / 8 - v0);                 // active vectors per pixel
  const int items = HW * vl;
  const int r = k >> 1;
  const long long pbase = (long long)n * HW;
  const uint4 NEG = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);  // bf16 -inf x8

  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int p = i / vl, v = i - p * vl;
    const int h_ = p / W, w_ = p - h_ * W;
    A[(h_ * Wp + w_) * VL + v] = ld_nc16(x + (pbase + p) * ldx + (v0 + v) * 8);
  }
  __syncthreads();
  __nv_bfloat16* outs[3] = {y1, y2, y3};
#pragma unroll
  for (int s = 0; s < STAGES; ++s) {
    if (K5) {
      // row pass: thread = (row h, vector v) walks w with the window in registers
      for (int i = threadIdx.x; i < H * vl; i += blockDim.x) {
        const int h_ = i / vl, v = i - h_ * vl;
        const uint4* in = A + (size_t)h_ * Wp * VL + v;
        uint4* out = B + (size_t)h_ * Wp * VL + v;
        uint4 a = NEG, b = NEG, c = in[0], d = W > 1 ? in[VL] : NEG, e;
#pragma unroll 5
        for (int w_ = 0; w_ < W; ++w_) {
          e = w_ + 2 < W ? in[(w_ + 2) * VL] : NEG;
          out[w_ * VL] = max16(max16(max16(a, b), max16(c, d)), e);
          a = b; b = c; c = d; d = e;
        }
      }
      __syncthreads();
      // column pass: thread = (column w, vector v) walks h; writes the stage output and the next stage's input
      for (int i = threadIdx.x; i < W * vl; i += blockDim.x) {
        const int w_ = i / vl, v = i - w_ * vl;
        const uint4* in = B + (size_t)w_ * VL + v;
        uint4* nxt = A + (size_t)w_ * VL + v;
        __nv_bfloat16* go = outs[s] + (pbase + w_) * ldy + (v0 + v) * 8;
        const int pitch = Wp * VL;
        uint4 a = NEG, b = NEG, c = in[0], d = H > 1 ? in[pitch] : NEG, e;
#pragma unroll 5
        for (int h_ = 0; h_ < H; ++h_) {
          e = h_ + 2 < H ? in[(h_ + 2) * pitch] : NEG;
          const uint4 m = max16(max16(max16(a, b), max16(c, d)), e);
          st_na16(go + (long long)h_ * W * ldy, m);
          if (s + 1 < STAGES) nxt[h_ * pitch] = m;
          a = b; b = c; c = d; d = e;
        }
      }
      __syncthreads();
    } else {
      for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int p = i / vl, v = i - p * vl;
        const int h_ = p / W, w_ = p - h_ * W;
        const int lo = max(w_ - r, 0), hi = min(w_ + r, W - 1);
        uint4 m = A[(h_ * Wp + lo) * VL + v];
        for (int ww = lo + 1; ww <= hi; ++ww) m = max16(m, A[(h_ * Wp + ww) * VL + v]);
        B[(h_ * Wp + w_) * VL + v] = m;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < items; i += blockDim.x) {
        const int p = i / vl, v = i - p * vl;
        const int h_ = p / W, w_ = p - h_ * W;
        const int lo = max(h_ - r, 0), hi = min(h_ + r, H - 1);
        uint4 m = B[(lo * Wp + w_) * VL + v];
        for (int hh = lo + 1; hh <= hi; ++hh) m = max16(m, B[(hh * Wp + w_) * VL + v]);
        st_na16(outs[s] + (pbase + p) * ldy + (v0 + v) * 8, m);
        if (s + 1 < STAGES) A[(h_ * Wp + w_) * VL + v] = m;
      }
      __syncthreads();
    }
  }
}

// ---- register-resident variant for maps up to 20 x 20 (every 640-class input at stride 32), k == 5 ----------
// CTA = (image, 64 channels).  A thread owns one column (or one row) of one 16-byte channel vector — 20 values in
// registers — so a 5-window pass is 20 independent sliding-window updates with no shared-memory traffic at all;
// shared memory is touched only to TRANSPOSE between column owners and row owners, once per stage:
//   load x (column owners, 20 independent 16-byte loads each)
//   stage 1: vertical pass | transpose | horizontal pass -> y1 (stored by the row owners)
//   stage 2: horizontal pass | transpose | vertical pass -> y2 (column owners)
//   stage 3: vertical pass | transpose | horizontal pass -> y3 (row owners)
// 6 shared-memory accesses per element instead of 12, and 20-way instruction-level parallelism per thread instead
// of a serial walk with a shared-memory round trip per step (the plane kernel ran at 47 % of HBM peak).
constexpr int kRegD = 20;
constexpr int kRegVL = 8;

__device__ __forceinline__ void window5(uint4 (&a)[kRegD + 4]) {
  // in place: a[i] <- max(a[i-2 .. i+2]); entries beyond the map hold -inf (a[kRegD .. kRegD+3] are padding)
  const uint4 NEG = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
  uint4 m2_prev2 = NEG, m2_prev1 = a[0];      // pairs (i-2, i-1) and (i-1, i) of the ORIGINAL values; (-1, 0) = a[0]
#pragma unroll
  for (int i = 0; i < kRegD; ++i) {
    const uint4 m2 = max16(a[i], a[i + 1]);   // pair (i, i+1)
    const uint4 out = max16(max16(m2_prev2, m2), a[i + 2]);
    m2_prev2 = m2_prev1;
    m2_prev1 = m2;
    a[i] = out;                                // a[i] is not read again (later outputs use a[i+1..])
  }
}

template <int STAGES>
__global__ void __launch_bounds__(kRegD* kRegVL, 3) pool_reg_kernel(const __nv_bfloat16* __restrict__ x,
                                                                 __nv_bfloat16* __restrict__ y1,
                                                                 __nv_bfloat16* __restrict__ y2,
                                                                 __nv_bfloat16* __restrict__ y3, int H, int W, int C,
                                                                 int ldx, int ldy) {
  extern __shared__ uint4 tbuf[];   // [H][W|1][kRegVL]: odd pitch keeps both access directions conflict-light
  const uint4 NEG = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
  const int Wp = W | 1;
  const int groups = (C / 8 + kRegVL - 1) / kRegVL;
  const int n = blockIdx.x / groups, g = blockIdx.x % groups;
  const int v = threadIdx.x & (kRegVL - 1), line = threadIdx.x >> 3;   // line = column index (column owner) / row index (row owner)
  const int vg = g * kRegVL + v;
  const bool von = vg < C / 8;
  const bool col_on = von && line < W, row_on = von && line < H;
  const long long pbase = (long long)n * H * W;
  __nv_bfloat16* outs[3] = {y1, y2, y3};
  uint4 a[kRegD + 4];
#pragma unroll
  for (int i = 0; i < kRegD + 4; ++i) a[i] = NEG;
  // column owners load their column
  if (col_on) {
#pragma unroll
    for (int h = 0; h < kRegD; ++h)
      if (h < H) a[h] = ld_nc16(x + (pbase + (long long)h * W + line) * ldx + vg * 8);
  }
  bool col_owner = true;
#pragma unroll
  for (int s = 0; s < STAGES; ++s) {
    // pass 1 in the current ownership
    window5(a);
    // transpose through shared memory
    __syncthreads();
    if (col_owner) {
      if (col_on) {
#pragma unroll
        for (int h = 0; h < kRegD; ++h)
          if (h < H) tbuf[(h * Wp + line) * kRegVL + v] = a[h];
      }
    } else if (row_on) {
#pragma unroll
      for (int w = 0; w < kRegD; ++w)
        if (w < W) tbuf[(line * Wp + w) * kRegVL + v] = a[w];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kRegD + 4; ++i) a[i] = NEG;
    col_owner = !col_owner;
    if (col_owner) {
      if (col_on) {
#pragma unroll
        for (int h = 0; h < kRegD; ++h)
          if (h < H) a[h] = tbuf[(h * Wp + line) * kRegVL + v];
      }
    } else if (row_on) {
#pragma unroll
      for (int w = 0; w < kRegD; ++w)
        if (w < W) a[w] = tbuf[(line * Wp + w) * kRegVL + v];
    }
    // pass 2 in the new ownership, then store the stage output from registers
    window5(a);
    __nv_bfloat16* go = outs[s];
    if (col_owner) {
      if (col_on) {
#pragma unroll
        for (int h = 0; h < kRegD; ++h)
          if (h < H) st_na16(go + (pbase + (long long)h * W + line) * ldy + vg * 8, a[h]);
      }
    } else if (row_on) {
#pragma unroll
      for (int w = 0; w < kRegD; ++w)
        if (w < W) st_na16(go + (pbase + (long long)line * W + w) * ldy + vg * 8, a[w]);
    }
    // entries beyond the map must be -inf again before the next window pass
    const int lim = col_owner ? H : W;
#pragma unroll
    for (int i = 0; i < kRegD; ++i)
      if (i >= lim) a[i] = NEG;
  }
}

// Fallback for planes that do not fit in shared memory: nested windows straight from global/L2.
__global__ void __launch_bounds__(256) pool_direct_kernel(const __nv_bfloat16* __restrict__ x,
                                                          __nv_bfloat16* __restrict__ y1, __nv_bfloat16* __restrict__ y2,
                                                          __nv_bfloat16* __restrict__ y3, int N, int H, int W, int C,
                                                          int ldx, int ldy, int k, int stages) {
  const int cv = C >> 3, r = k >> 1;
  const long long items = (long long)N * H * W * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items;
       i += (long long)gridDim.x * blockDim.x) {
    int v = (int)(i % cv);
    long long pix = i / cv;
    int w_ = (int)(pix % W);
    long long t = pix / W;
    int h_ = (int)(t % H);
    int n = (int)(t / H);
    const uint4 c0 = ld16(x + pix * ldx + v * 8);
    uint4 m1 = c0, m2 = c0, m3 = c0;
    const int R = r * stages;
    for (int dy = -R; dy <= R; ++dy) {
      int hh = h_ + dy;
      if (hh < 0 || hh >= H) continue;
      for (int dx = -R; dx <= R; ++dx) {
        int ww = w_ + dx;
        if (ww < 0 || ww >= W) continue;
        uint4 val = ld16(x + (((long long)n * H + hh) * W + ww) * ldx + v * 8);
        int d = max(abs(dy), abs(dx));
        if (d <= r) m1 = max16(m1, val);
        if (d <= 2 * r) m2 = max16(m2, val);
        m3 = max16(m3, val);
      }
    }
    st_na16(y1 + pix * ldy + v * 8, m1);
    if (stages > 1) st_na16(y2 + pix * ldy + v * 8, m2);
    if (stages > 2) st_na16(y3 + pix * ldy + v * 8, m3);
  }
}

static int launch_pool(const void* x, void* y1, void* y2, void* y3, int N, int H, int W, int C, int ldx, int ldy,
                       int k, int stages, cudaStream_t s) {
  if (!x || !y1 || N <= 0 || H <= 0 || W <= 0 || C <= 0) return DMAY_EINVAL;
  if (stages == 3 && (!y2 || !y3)) return DMAY_EINVAL;
  if (!(k & 1) || k < 1) return DMAY_EUNSUPPORTED;
  if ((C | ldx | ldy) & 7) return DMAY_EUNSUPPORTED;
  if (!aligned16(x) || !aligned16(y1) || (y2 && !aligned16(y2)) || (y3 && !aligned16(y3))) return DMAY_EINVAL;
  const long long HW = (long long)H * W;
  const int cvec = C / 8;
  if (k == 5 && H <= kRegD && W <= kRegD) {   // register-resident variant
    const int groups = (cvec + kRegVL - 1) / kRegVL;
    const long long grid = (long long)N * groups;
    if (grid > 0x7fffffffLL) return DMAY_EUNSUPPORTED;
    const size_t smem = (size_t)H * (W | 1) * kRegVL * 16;
    const __nv_bfloat16* xi = (const __nv_bfloat16*)x;
    __nv_bfloat16 *o1 = (__nv_bfloat16*)y1, *o2 = (__nv_bfloat16*)y2, *o3 = (__nv_bfloat16*)y3;
    if (stages == 3) {
      cudaError_t e = cudaFuncSetAttribute(pool_reg_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      cudaFuncSetAttribute(pool_reg_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      pool_reg_kernel<3><<<(int)grid, kRegD * kRegVL, smem, s>>>(xi, o1, o2, o3, H, W, C, ldx, ldy);
    } else {
      cudaError_t e = cudaFuncSetAttribute(pool_reg_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      cudaFuncSetAttribute(pool_reg_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      pool_reg_kernel<1><<<(int)grid, kRegD * kRegVL, smem, s>>>(xi, o1, o2, o3, H, W, C, ldx, ldy);
    }
    return finish_launch();
  }
  // largest VL (vectors per pixel per CTA) whose two planes fit in ~200 KB; prefer >=2 CTAs per SM
  int VL = 0;
  for (int cand : {8, 4, 2, 1}) {
    if (cand > cvec && cand != 1) continue;
    if ((long long)H * (W | 1) * cand * 16 * 2 <= 100 * 1024) { VL = cand; break; }
  }
  if (VL == 0)
    for (int cand : {8, 4, 2, 1}) {
      if (cand > cvec && cand != 1) continue;
      if ((long long)H * (W | 1) * cand * 16 * 2 <= 200 * 1024) { VL = cand; break; }
    }
  if (VL == 0) {
    long long items = (long long)N * HW * cvec;
    pool_direct_kernel<<<grid_for(items, 256), 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y1,
                                                            (__nv_bfloat16*)y2, (__nv_bfloat16*)y3, N, H, W, C, ldx,
                                                            ldy, k, stages);
    return finish_launch();
  }
  const size_t smem = (size_t)H * (W | 1) * VL * 16 * 2;
  const int groups = (cvec + VL - 1) / VL;
  const long long grid = (long long)N * groups;
  if (grid > 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  const __nv_bfloat16* xi = (const __nv_bfloat16*)x;
  __nv_bfloat16 *o1 = (__nv_bfloat16*)y1, *o2 = (__nv_bfloat16*)y2, *o3 = (__nv_bfloat16*)y3;
#define DMAY_POOL_LAUNCH(ST, K5)                                                                                  \
  do {                                                                                                            \
    cudaError_t e = cudaFuncSetAttribute(pool_plane_kernel<ST, K5>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         (int)smem);                                                              \
    if (e != cudaSuccess) return (int)e;                                                                          \
    pool_plane_kernel<ST, K5><<<(int)grid, 256, smem, s>>>(xi, o1, o2, o3, H, W, C, ldx, ldy, k, VL);             \
  } while (0)
  if (stages == 3) {
    if (k == 5) DMAY_POOL_LAUNCH(3, true);
    else DMAY_POOL_LAUNCH(3, false);
  } else {
    if (k == 5) DMAY_POOL_LAUNCH(1, true);
    else DMAY_POOL_LAUNCH(1, false);
  }
#undef DMAY_POOL_LAUNCH
  return finish_launch();
}

}  // namespace dmay

extern "C" {
int dmay_sppf_pool3(const dmay_sppf_params* p, dmay_stream_t stream) {
  if (!p) return DMAY_EINVAL;
  return dmay::launch_pool(p->x, p->y1, p->y2, p->y3, p->N, p->H, p->W, p->C, p->ldx, p->ldy, p->k, 3,
                           (cudaStream_t)stream);
}
int dmay_maxpool_s1(const dmay_maxpool_params* p, dmay_stream_t stream) {
  if (!p) return DMAY_EINVAL;
  return dmay::launch_pool(p->x, p->y, nullptr, nullptr, p->N, p->H, p->W, p->C, p->ldx, p->ldy, p->k, 1,
                           (cudaStream_t)stream);
}
}
