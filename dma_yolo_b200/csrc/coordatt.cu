// a3: CoordAtt (models/common.py:1183-1207) on NHWC bf16.
//   pool : pooled[n, h, c] = mean_w x ; pooled[n, H+w, c] = mean_h x            (fp32)
//   mlp  : y = hardswish(s1*(W1.pooled + b1) + t1) ; gates = sigmoid(W{h,w}.y + b{h,w})  (fp32)
//   apply: out = (x * a_w) * a_h                                                 (bf16)
// x is read once by `pool` from HBM and once by `apply` (L2-resident at cfg-2 sizes: 52 MB < 126 MB
// L2); the gate tensors are (H+W)/(H*W) of the activation.  All reductions are deterministic
// (no atomics): a CTA owns complete rows (for mean_w) and complete columns (for mean_h).
#include "common.cuh"

namespace dmay {

// ---- pool, plane-in-smem variant: CTA = (image, group of VL channel vectors); the H x W x VL plane is
// staged with fully coalesced, fully parallel loads, then row / column sums run out of shared memory.
__global__ void __launch_bounds__(256) ca_pool_plane_kernel(const __nv_bfloat16* __restrict__ x,
                                                            float* __restrict__ pooled, int H, int W, int C, int ldx,
                                                            int VL) {
  extern __shared__ uint4 plane[];  // [H*W][VL]
  const int cvec = C >> 3;
  const int groups = (cvec + VL - 1) / VL;
  const int n = blockIdx.x / groups, g = blockIdx.x % groups;
  const int v0 = g * VL;
  const int vl = min(VL, cvec - v0);
  const int HW = H * W;
  const __nv_bfloat16* xb = x + (long long)n * HW * ldx + v0 * 8;
  for (int i = threadIdx.x; i < HW * vl; i += blockDim.x) {
    const int p = i / vl, v = i - p * vl;
    plane[p * VL + v] = ld16(xb + (long long)p * ldx + v * 8);  // default caching: `apply` re-reads x from L2
  }
  __syncthreads();
  float* pb = pooled + (long long)n * (H + W) * C + v0 * 8;
  const float invW = 1.0f / (float)W, invH = 1.0f / (float)H;
  // items: (position p in [0, H+W), vector v)
  for (int i = threadIdx.x; i < (H + W) * vl; i += blockDim.x) {
    const int p = i / vl, v = i - p * vl;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    int start, step, cnt;
    float inv;
    if (p < H) { start = p * W; step = 1; cnt = W; inv = invW; }
    else { start = p - H; step = W; cnt = H; inv = invH; }
    for (int k = 0; k < cnt; ++k) {
      float f[8];
      unpack8(plane[(start + k * step) * VL + v], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    float4* o = reinterpret_cast<float4*>(pb + (long long)p * C + v * 8);
    o[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    o[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
  }
}

// ---- pool, direct variant for planes that do not fit in shared memory.
// grid = N * groups * bands.  CTA (n, g, b): channel vectors [g*8, g*8+8) (64 channels),
// rows h === b (mod bands) reduced over w, columns w === b (mod bands) reduced over h.
// thread = (slot 0..31, lane 0..7): lane picks the 16-byte channel vector, slot the row/column.
__global__ void __launch_bounds__(256) ca_pool_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ pooled,
                                                      int H, int W, int C, int ldx, int bands) {
  const int cvec = C >> 3;
  const int groups = (cvec + 7) >> 3;
  int bid = blockIdx.x;
  const int b = bid % bands;
  bid /= bands;
  const int g = bid % groups;
  const int n = bid / groups;
  const int lane = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const int v = g * 8 + lane;
  if (v >= cvec) return;
  const __nv_bfloat16* xb = x + (long long)n * H * W * ldx + v * 8;
  float* pb = pooled + (long long)n * (H + W) * C + v * 8;
  const float invW = 1.0f / (float)W, invH = 1.0f / (float)H;
  for (int h = b + slot * bands; h < H; h += 32 * bands) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const __nv_bfloat16* row = xb + (long long)h * W * ldx;
#pragma unroll 4
    for (int w = 0; w < W; ++w) {
      float f[8];
      unpack8(ld16(row + (long long)w * ldx), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    float4* o = reinterpret_cast<float4*>(pb + (long long)h * C);
    o[0] = make_float4(acc[0] * invW, acc[1] * invW, acc[2] * invW, acc[3] * invW);
    o[1] = make_float4(acc[4] * invW, acc[5] * invW, acc[6] * invW, acc[7] * invW);
  }
  for (int w = b + slot * bands; w < W; w += 32 * bands) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const __nv_bfloat16* col = xb + (long long)w * ldx;
#pragma unroll 4
    for (int h = 0; h < H; ++h) {
      float f[8];
      unpack8(ld16(col + (long long)h * W * ldx), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    float4* o = reinterpret_cast<float4*>(pb + (long long)(H + w) * C);
    o[0] = make_float4(acc[0] * invH, acc[1] * invH, acc[2] * invH, acc[3] * invH);
    o[1] = make_float4(acc[4] * invH, acc[5] * invH, acc[6] * invH, acc[7] * invH);
  }
}

// ---- mlp.  grid = N * (ceil(H/PG) + ceil(W/PG)); a CTA handles PG consecutive positions of one image that
// all lie in the H part or all in the W part (so it needs only one of Wh / Ww).
// w1T: [C][Cm] fp32, whT / wwT: [Cm][Cout] fp32 (transposed: consecutive threads read consecutive floats).
// Hidden layer: warp w owns the channel range [w*C/8, (w+1)*C/8) for every hidden unit (lane = j) and all PG
// positions (register accumulators), so W1 is streamed ONCE per CTA with 8 independent loads in flight per
// lane; partial sums are combined through shared memory.
constexpr int PG = 8;
constexpr int kMlpWarps = 8;
__global__ void __launch_bounds__(kMlpWarps * 32) ca_mlp_kernel(const float* __restrict__ pooled, float* __restrict__ gates,
                                                                const float* __restrict__ w1T, const float* __restrict__ b1,
                                                                const float* __restrict__ s1, const float* __restrict__ t1,
                                                                const float* __restrict__ whT, const float* __restrict__ bh,
                                                                const float* __restrict__ wwT, const float* __restrict__ bw,
                                                                int H, int W, int C, int Cm, int Cout) {
  extern __shared__ float sm[];
  float* sp = sm;                         // [PG][C]
  float* spart = sm + PG * C;             // [kMlpWarps][PG][Cm] partial hidden sums
  float* sy = spart + kMlpWarps * PG * Cm;  // [PG][Cm]
  const int P = H + W;
  const int gh = (H + PG - 1) / PG, gw = (W + PG - 1) / PG;
  const int n = blockIdx.x / (gh + gw);
  const int g = blockIdx.x % (gh + gw);
  const bool is_h = g < gh;
  const int p0 = is_h ? g * PG : H + (g - gh) * PG;
  const int np = min(PG, (is_h ? H : P) - p0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < PG * C; i += blockDim.x) {
    const int pp = i / C, c = i - pp * C;
    sp[i] = pp < np ? pooled[((long long)n * P + p0 + pp) * C + c] : 0.f;
  }
  __syncthreads();
  const int cper = (C + kMlpWarps - 1) / kMlpWarps;
  const int c0 = warp * cper, c1 = min(C, c0 + cper);
  for (int j = lane; j < Cm; j += 32) {
    float acc[PG];
#pragma unroll
    for (int q = 0; q < PG; ++q) acc[q] = 0.f;
#pragma unroll 8
    for (int c = c0; c < c1; ++c) {
      const float wv = w1T[(long long)c * Cm + j];
#pragma unroll
      for (int q = 0; q < PG; ++q) acc[q] = fmaf(wv, sp[q * C + c], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < PG; ++q) spart[(warp * PG + q) * Cm + j] = acc[q];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < PG * Cm; o += blockDim.x) {
    const int j = o % Cm;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kMlpWarps; ++w) acc += spart[w * PG * Cm + o];
    sy[o] = hardswish(s1[j] * (acc + b1[j]) + t1[j]);
  }
  __syncthreads();
  const float* wT = is_h ? whT : wwT;
  const float* bias = is_h ? bh : bw;
  for (int c = threadIdx.x; c < Cout; c += blockDim.x) {
    float acc[PG];
    const float bc = bias[c];
#pragma unroll
    for (int q = 0; q < PG; ++q) acc[q] = bc;
#pragma unroll 8
    for (int j = 0; j < Cm; ++j) {
      const float a = wT[(long long)j * Cout + c];
#pragma unroll
      for (int q = 0; q < PG; ++q) acc[q] = fmaf(a, sy[q * Cm + j], acc[q]);
    }
    for (int q = 0; q < np; ++q) gates[((long long)n * P + p0 + q) * Cout + c] = sigmoid_acc(acc[q]);
  }
}

// ---- apply: blockDim = (VX channel vectors, PY pixels), 32-bit pixel index.
__global__ void __launch_bounds__(256) ca_apply_kernel(const __nv_bfloat16* __restrict__ x,
                                                       const float* __restrict__ gates, __nv_bfloat16* __restrict__ y,
                                                       unsigned npix, int H, int W, int C, int ldx, int ldy) {
  const int cv = C >> 3;
  for (unsigned pix = blockIdx.x * blockDim.y + threadIdx.y; pix < npix; pix += gridDim.x * blockDim.y) {
    const unsigned t = pix / (unsigned)W;
    const int w_ = (int)(pix - t * (unsigned)W);
    const int n = (int)(t / (unsigned)H);
    const int h_ = (int)(t - (unsigned)n * (unsigned)H);
    const float* gh = gates + ((long long)n * (H + W) + h_) * C;
    const float* gw = gates + ((long long)n * (H + W) + H + w_) * C;
    for (int v = threadIdx.x; v < cv; v += blockDim.x) {
      float f[8];
      unpack8(ld_nc16(x + (long long)pix * ldx + v * 8), f);
      const float4 h0 = reinterpret_cast<const float4*>(gh + v * 8)[0], h1 = reinterpret_cast<const float4*>(gh + v * 8)[1];
      const float4 w0 = reinterpret_cast<const float4*>(gw + v * 8)[0], w1 = reinterpret_cast<const float4*>(gw + v * 8)[1];
      f[0] = (f[0] * w0.x) * h0.x; f[1] = (f[1] * w0.y) * h0.y; f[2] = (f[2] * w0.z) * h0.z; f[3] = (f[3] * w0.w) * h0.w;
      f[4] = (f[4] * w1.x) * h1.x; f[5] = (f[5] * w1.y) * h1.y; f[6] = (f[6] * w1.z) * h1.z; f[7] = (f[7] * w1.w) * h1.w;
      st_na16(y + (long long)pix * ldy + v * 8, pack8(f));
    }
  }
}

// ---- fast path (plane of 64 channels fits in shared memory): TWO launches ---------------------------------
// K1  ca_pool_hidden_kernel : CTA = (image, 64-channel group).  Stages its H x W x 64 plane (one coalesced read of x),
//     reduces it to the pooled row / column means, multiplies them with its 64-row slice of W1 (partial hidden
//     layer, [H+W][Cm] per CTA) and writes the partial.  The LAST CTA of an image to arrive (atomic ticket) sums
//     the image's partials and applies bias + folded BN + hardswish: y[n][H+W][Cm].  No spin-waiting.
// K2  ca_gate_apply_kernel  : CTA = (image, 64-channel group).  Computes the sigmoid gates of its 64 channels from
//     y (shared memory, 10 KB), then streams its plane once more (L2-resident at cfg-2: 52 MB) and writes
//     out = (x * a_w) * a_h.  Gates / pooled means go to global memory only when the caller asks for them.
constexpr int kCaVL = 8;   // 16-byte channel vectors per CTA (64 channels)

__global__ void __launch_bounds__(256) ca_pool_hidden_kernel(const __nv_bfloat16* __restrict__ x,
                                                             float* __restrict__ pooled_out, float* __restrict__ partial,
                                                             float* __restrict__ yhid, unsigned* __restrict__ counters,
                                                             const float* __restrict__ w1T, const float* __restrict__ b1,
                                                             const float* __restrict__ s1, const float* __restrict__ t1,
                                                             int H, int W, int C, int Cm, int ldx, int G) {
  extern __shared__ uint4 plane[];                    // [H*W][kCaVL]
  const int HW = H * W, P = H + W;
  float* pooled_s = reinterpret_cast<float*>(plane + (size_t)HW * kCaVL);   // [P][64]
  __shared__ unsigned ticket_s;
  const int n = blockIdx.x / G, g = blockIdx.x % G;
  const int cvec = C >> 3;
  const int v0 = g * kCaVL;
  const int vl = min(kCaVL, cvec - v0);
  const __nv_bfloat16* xb = x + (long long)n * HW * ldx + v0 * 8;
  // four independent 16-byte loads in flight per thread (a load -> store loop left one: ~12 KB in flight per SM)
  for (int i0 = threadIdx.x; i0 < HW * kCaVL; i0 += 4 * blockDim.x) {
    uint4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      const int p = i >> 3, v = i & 7;
      r[u] = (i < HW * kCaVL && v < vl) ? ld16(xb + (long long)p * ldx + v * 8) : make_uint4(0, 0, 0, 0);   // default caching: K2 re-reads x from L2
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < HW * kCaVL) plane[i] = r[u];
    }
  }
  __syncthreads();
  const float invW = 1.0f / (float)W, invH = 1.0f / (float)H;
  for (int i = threadIdx.x; i < P * kCaVL; i += blockDim.x) {
    const int p = i >> 3, v = i & 7;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    int start, step, cnt;
    float inv;
    if (p < H) { start = p * W; step = 1; cnt = W; inv = invW; }
    else { start = p - H; step = W; cnt = H; inv = invH; }
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      float f[8];
      unpack8(plane[(start + k * step) * kCaVL + v], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] *= inv;
      pooled_s[p * 64 + v * 8 + j] = acc[j];
    }
    if (pooled_out != nullptr && v < vl) {
      float4* o = reinterpret_cast<float4*>(pooled_out + ((long long)n * P + p) * C + (v0 + v) * 8);
      o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
  __syncthreads();
  // partial hidden layer of this channel group: [P][Cm]; thread -> (position p, hidden unit j).  The group's 64-row
  // slice of W1 is staged in shared memory: strided L2 loads inside the dot product were latency-bound.
  const int nch = vl * 8;
  float* w1s = pooled_s + P * 64;                               // [nch][Cm]
  for (int i = threadIdx.x; i < 64 * Cm; i += blockDim.x) w1s[i] = i < nch * Cm ? w1T[(long long)(v0 * 8) * Cm + i] : 0.f;
  __syncthreads();
  float* part = partial + (long long)(n * G + g) * P * Cm;
  // register tile: a thread owns hidden unit j for up to 8 positions (p = pg, pg + PGS, ...), so one W1 element and
  // one float4 of pooled values per position feed 4 FMAs each (the naive form issued 2 loads per FMA)
  {
    const int PGS = blockDim.x / Cm > 0 ? blockDim.x / Cm : 1;     // position groups covered by the block at once
    const int j = threadIdx.x % Cm, pg = threadIdx.x / Cm;
    if (pg < PGS) {
      for (int p0 = pg; p0 < P; p0 += 8 * PGS) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int c = 0; c < 64; c += 4) {
          const float w0 = w1s[(c + 0) * Cm + j], w1 = w1s[(c + 1) * Cm + j], w2 = w1s[(c + 2) * Cm + j], w3 = w1s[(c + 3) * Cm + j];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int p = p0 + k * PGS;
            if (p < P) {
              const float4 pv = *reinterpret_cast<const float4*>(pooled_s + p * 64 + c);
              acc[k] = fmaf(w3, pv.w, fmaf(w2, pv.z, fmaf(w1, pv.y, fmaf(w0, pv.x, acc[k]))));
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int p = p0 + k * PGS;
          if (p < P) part[p * Cm + j] = acc[k];
        }
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket_s = atomicAdd(counters + n, 1u);
  __syncthreads();
  if (ticket_s == (unsigned)(G - 1)) {   // every other group of this image has published its partial
    __threadfence();
    const float* pn = partial + (long long)n * G * P * Cm;
    for (int o = threadIdx.x; o < P * Cm; o += blockDim.x) {
      const int j = o % Cm;
      float acc = 0.f;
#pragma unroll 8
      for (int gg = 0; gg < G; ++gg) acc += __ldcg(pn + (long long)gg * P * Cm + o);
      yhid[(long long)n * P * Cm + o] = hardswish(s1[j] * (acc + b1[j]) + t1[j]);
    }
    if (threadIdx.x == 0) counters[n] = 0u;   // ready for the next launch
  }
}

__global__ void __launch_bounds__(256) ca_gate_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ yhid,
                                                            __nv_bfloat16* __restrict__ out, float* __restrict__ gates_out,
                                                            const float* __restrict__ whT, const float* __restrict__ bh,
                                                            const float* __restrict__ wwT, const float* __restrict__ bw,
                                                            int H, int W, int C, int Cm, int ldx, int ldy, int G) {
  extern __shared__ float sm[];
  const int HW = H * W, P = H + W;
  float* ys = sm;                 // [P][Cm]
  float* gs = sm + P * Cm;        // [P][64]
  const int n = blockIdx.x / G, g = blockIdx.x % G;
  const int cvec = C >> 3;
  const int v0 = g * kCaVL;
  const int vl = min(kCaVL, cvec - v0);
  const int nch = vl * 8;
  float* whs = gs + P * 64;       // [Cm][64]
  float* wws = whs + Cm * 64;     // [Cm][64]
  for (int i = threadIdx.x; i < P * Cm; i += blockDim.x) ys[i] = yhid[(long long)n * P * Cm + i];
  for (int i = threadIdx.x; i < Cm * 64; i += blockDim.x) {
    const int j = i >> 6, c = i & 63;
    const bool in = c < nch;
    whs[i] = in ? whT[(long long)j * C + v0 * 8 + c] : 0.f;
    wws[i] = in ? wwT[(long long)j * C + v0 * 8 + c] : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < P * 64; o += blockDim.x) {
    const int p = o >> 6, c = o & 63;
    float gate = 0.f;
    if (c < nch) {
      const bool is_h = p < H;
      const float* wT = (is_h ? whs : wws) + c;
      float acc = (is_h ? bh : bw)[v0 * 8 + c];
      const float* yr = ys + p * Cm;
#pragma unroll 8
      for (int j = 0; j < Cm; ++j) acc = fmaf(wT[j * 64], yr[j], acc);
      gate = sigmoid_acc(acc);
      if (gates_out != nullptr) gates_out[((long long)n * P + p) * C + v0 * 8 + c] = gate;
    }
    gs[o] = gate;
  }
  __syncthreads();
  const __nv_bfloat16* xb = x + (long long)n * HW * ldx + v0 * 8;
  __nv_bfloat16* ob = out + (long long)n * HW * ldy + v0 * 8;
  for (int i0 = threadIdx.x; i0 < HW * kCaVL; i0 += 4 * blockDim.x) {
    uint4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      const int p = i >> 3, v = i & 7;
      r[u] = (i < HW * kCaVL && v < vl) ? ld_nc16(xb + (long long)p * ldx + v * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      const int p = i >> 3, v = i & 7;
      if (i >= HW * kCaVL || v >= vl) continue;
      const int h_ = p / W, w_ = p - h_ * W;
      float f[8];
      unpack8(r[u], f);
      const float4* gh = reinterpret_cast<const float4*>(gs + h_ * 64 + v * 8);
      const float4* gw = reinterpret_cast<const float4*>(gs + (H + w_) * 64 + v * 8);
      const float4 h0 = gh[0], h1 = gh[1], w0 = gw[0], w1 = gw[1];
      f[0] = (f[0] * w0.x) * h0.x; f[1] = (f[1] * w0.y) * h0.y; f[2] = (f[2] * w0.z) * h0.z; f[3] = (f[3] * w0.w) * h0.w;
      f[4] = (f[4] * w1.x) * h1.x; f[5] = (f[5] * w1.y) * h1.y; f[6] = (f[6] * w1.z) * h1.z; f[7] = (f[7] * w1.w) * h1.w;
      st_na16(ob + (long long)p * ldy + v * 8, pack8(f));
    }
  }
}

}  // namespace dmay

using namespace dmay;

extern "C" long long dmay_coordatt_ws(int N, int H, int W, int C, int Cm) {
  if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || Cm <= 0) return DMAY_EINVAL;
  const long long G = (C / 8 + kCaVL - 1) / kCaVL, P = H + W;
  return ((long long)N * G * P * Cm + (long long)N * P * Cm) * 4 + (long long)N * 4;
}

extern "C" int dmay_coordatt(const dmay_coordatt_params* p, dmay_stream_t stream) {
  if (!p || !p->x || !p->y || !p->w1 || !p->b1 || !p->s1 || !p->t1 || !p->wh || !p->bh || !p->ww || !p->bw)
    return DMAY_EINVAL;
  if (p->N <= 0 || p->H <= 0 || p->W <= 0 || p->C <= 0 || p->Cm <= 0) return DMAY_EINVAL;
  if ((p->C | p->ldx | p->ldy) & 7) return DMAY_EUNSUPPORTED;
  if (!aligned16(p->x) || !aligned16(p->y) || (p->pooled && !aligned16(p->pooled)) || (p->gates && !aligned16(p->gates)))
    return DMAY_EINVAL;
  const long long npix = (long long)p->N * p->H * p->W;
  if (npix >= 0x7fffffffLL) return DMAY_EUNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const int cvec = p->C / 8;
  const int sms = p->num_sms > 0 ? p->num_sms : sm_count();
  const long long HW = (long long)p->H * p->W;
  // fast path: two launches (see ca_pool_hidden_kernel); needs the partial / y / ticket workspace
  {
    const int G = (cvec + kCaVL - 1) / kCaVL, P = p->H + p->W;
    const size_t smem1 = (size_t)HW * kCaVL * 16 + (size_t)P * 64 * 4 + (size_t)64 * p->Cm * 4;
    const size_t smem2 = ((size_t)P * p->Cm + (size_t)P * 64 + 2 * (size_t)p->Cm * 64) * 4;
    const long long need = ((long long)p->N * G * P * p->Cm + (long long)p->N * P * p->Cm) * 4 + (long long)p->N * 4;
    if (p->ws != nullptr && p->ws_bytes >= need && smem1 <= 100 * 1024 && smem2 <= 100 * 1024 && aligned16(p->ws) &&
        p->Cm <= 256) {
      float* partial = (float*)p->ws;
      float* yhid = partial + (long long)p->N * G * P * p->Cm;
      unsigned* counters = (unsigned*)(yhid + (long long)p->N * P * p->Cm);   // zero on first use, self-resetting
      if (smem1 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(ca_pool_hidden_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess) return (int)e;
      }
      if (smem2 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(ca_gate_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return (int)e;
      }
      ca_pool_hidden_kernel<<<p->N * G, 256, smem1, s>>>((const __nv_bfloat16*)p->x, (float*)p->pooled, partial, yhid, counters,
                                                        (const float*)p->w1, (const float*)p->b1, (const float*)p->s1,
                                                        (const float*)p->t1, p->H, p->W, p->C, p->Cm, p->ldx, G);
      ca_gate_apply_kernel<<<p->N * G, 256, smem2, s>>>((const __nv_bfloat16*)p->x, yhid, (__nv_bfloat16*)p->y, (float*)p->gates,
                                                       (const float*)p->wh, (const float*)p->bh, (const float*)p->ww,
                                                       (const float*)p->bw, p->H, p->W, p->C, p->Cm, p->ldx, p->ldy, G);
      return finish_launch(2);
    }
  }
  if (!p->pooled || !p->gates) return DMAY_EINVAL;   // the three-launch path needs both workspaces
  // 1. pool
  int VL = 0;
  for (int cand : {8, 4, 2, 1}) {
    if (cand > cvec && cand != 1) continue;
    if (HW * cand * 16 <= 100 * 1024) { VL = cand; break; }
  }
  if (VL > 0) {
    const int groups = (cvec + VL - 1) / VL;
    const size_t smem = (size_t)HW * VL * 16;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(ca_pool_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    ca_pool_plane_kernel<<<p->N * groups, 256, smem, s>>>((const __nv_bfloat16*)p->x, (float*)p->pooled, p->H, p->W,
                                                          p->C, p->ldx, VL);
  } else {
    const int groups = (cvec + 7) / 8;
    int bands = 1;
    while ((long long)p->N * groups * bands < 2LL * sms && bands * 2 <= (p->H < p->W ? p->H : p->W)) bands *= 2;
    long long g1 = (long long)p->N * groups * bands;
    if (g1 > 0x7fffffffLL) return DMAY_EUNSUPPORTED;
    ca_pool_kernel<<<(int)g1, 256, 0, s>>>((const __nv_bfloat16*)p->x, (float*)p->pooled, p->H, p->W, p->C, p->ldx,
                                           bands);
  }
  // 2. mlp
  const int pgroups = (p->H + PG - 1) / PG + (p->W + PG - 1) / PG;
  const size_t smem = ((size_t)PG * p->C + (size_t)(kMlpWarps + 1) * PG * p->Cm) * sizeof(float);
  if (smem > 200 * 1024) return DMAY_EUNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ca_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  ca_mlp_kernel<<<p->N * pgroups, kMlpWarps * 32, smem, s>>>((const float*)p->pooled, (float*)p->gates, (const float*)p->w1,
                                                  (const float*)p->b1, (const float*)p->s1, (const float*)p->t1,
                                                  (const float*)p->wh, (const float*)p->bh, (const float*)p->ww,
                                                  (const float*)p->bw, p->H, p->W, p->C, p->Cm, p->C);
  // 3. apply
  int vx = 1;
  while (vx * 2 <= cvec && vx < 32) vx *= 2;
  const dim3 blk(vx, 256 / vx);
  long long need = (npix + blk.y - 1) / blk.y;
  const long long cap = (long long)sms * 8;
  ca_apply_kernel<<<(int)(need < cap ? need : cap), blk, 0, s>>>((const __nv_bfloat16*)p->x, (const float*)p->gates,
                                                                 (__nv_bfloat16*)p->y, (unsigned)npix, p->H, p->W, p->C,
                                                                 p->ldx, p->ldy);
  return finish_launch(3);
}
