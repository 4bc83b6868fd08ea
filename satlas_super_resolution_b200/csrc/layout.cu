// HBM-bound layout / resampling kernels: NCHW<->NHWC ingest & egress, nearest and bilinear resampling,
// batched weight packing.  All are coalesced on the NHWC (channel-contiguous) side and vectorised 16 B
// where the channel count allows; grids are sized in multiples of the SM count (persistent grid-stride).
#include "common.cuh"

namespace ssr {

static int g_sms = 0;
static int grid_for(long work_items, int threads) {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  long blocks = (work_items + threads - 1) / threads;
  long cap = (long)g_sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------- ingest: planar (NCHW) -> NHWC bf16
// One thread per (pixel, 8-channel group): reads 8 planar values (each coalesced across the warp, which
// walks consecutive pixels), writes one 16-byte NHWC vector.  Channels >= C up to c_pad are zero-filled.
template <typename T>
__global__ void nchw_to_nhwc_bf16_kernel(const T* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int C,
                                         int H, int W, int dst_stride, int c_pad, float scale,
                                         const float* __restrict__ mean, const float* __restrict__ inv_std) {
  const long HW = (long)H * W;
  const int groups = c_pad / 8;
  const long total = (long)B * HW * groups;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i % ((long)B * HW);
    const int g = (int)(i / ((long)B * HW));
    const long n = pix / HW;
    const long hw = pix - n * HW;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float f = 0.f;
      if (c < C) {
        f = (float)src[(n * C + c) * HW + hw] * scale;
        if (mean) f = (f - mean[c]) * inv_std[c];
      }
      v[j] = f;
    }
    uint4 o;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(dst + pix * dst_stride + g * 8) = o;
  }
}

// ---------------------------------------------------------------- ingest with pixel_unshuffle (space-to-depth)
// pixel_unshuffle(x, s) of ssr/archs/arch_util.py:769-785 (the scale 1 / 2 front end of SSR_RRDBNet, rrdbnet_arch.py:117-120):
// out[n, c*s*s + i*s + j, y, x] = in[n, c, y*s + i, x*s + j], written NHWC bf16.  One thread per (output pixel, 8-channel group).
__global__ void nchw_unshuffle_to_nhwc_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int C,
                                                   int H, int W, int s, int dst_stride, int c_pad, float scale) {
  const int h = H / s, w = W / s, ss = s * s;
  const long hw_out = (long)h * w;
  const int groups = c_pad / 8;
  const long total = (long)B * hw_out * groups;
  const int C_out = C * ss;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i % ((long)B * hw_out);
    const int g = (int)(i / ((long)B * hw_out));
    const long n = pix / hw_out;
    const long r = pix - n * hw_out;
    const int y = (int)(r / w), x = (int)(r - (long)y * w);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = g * 8 + j;
      float f = 0.f;
      if (co < C_out) {
        const int c = co / ss, ij = co - c * ss;
        f = src[((n * C + c) * H + (y * s + ij / s)) * (long)W + (x * s + ij % s)] * scale;
      }
      v[j] = f;
    }
    uint4 o;
    __nv_bfloat162 hh;
    hh = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&hh);
    hh = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&hh);
    hh = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&hh);
    hh = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&hh);
    *reinterpret_cast<uint4*>(dst + pix * dst_stride + g * 8) = o;
  }
}

// ---------------------------------------------------------------- egress: NHWC bf16 -> planar f32
__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ src, int src_stride,
                                             float* __restrict__ dst, int B, int C, int H, int W, float scale,
                                             int accumulate, const float* __restrict__ ch_scale) {
  const long HW = (long)H * W;
  const long total = (long)B * C * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long hw = i % HW;
    const long t = i / HW;
    const int c = (int)(t % C);
    const long n = t / C;
    float v = __bfloat162float(src[(n * HW + hw) * src_stride + c]) * scale;
    if (ch_scale) v *= ch_scale[c];
    if (accumulate) dst[i] += v; else dst[i] = v;
  }
}

// ---------------------------------------------------------------- inference egress: f32 NCHW -> uint8 HWC canvas
// out = uint8(clamp(v, 0, 1) * 255) with truncation (np.astype(np.uint8) of ssr/infer.py:61-64), image i of the batch
// pasted at tile (i / grid_cols, i % grid_cols) of a canvas with `canvas_w` pixels per row (ssr/utils/infer_utils.py:41-60).
__global__ void f32_nchw_to_u8_canvas_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int B, int C, int H, int W,
                                             int canvas_w, int grid_cols, int first_index) {
  const long HW = (long)H * W;
  const long total = (long)B * HW;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / HW;
    const long hw = i - n * HW;
    const int y = (int)(hw / W), x = (int)(hw % W);
    const long idx = n + first_index;
    const long ty = idx / grid_cols, tx = idx % grid_cols;
    uint8_t* o = dst + ((ty * H + y) * (long)canvas_w + tx * W + x) * C;
    for (int c = 0; c < C; ++c) {
      float v = src[(n * C + c) * HW + hw];
      v = fminf(fmaxf(v, 0.f), 1.f) * 255.f;
      o[c] = (uint8_t)v;
    }
  }
}

// ---------------------------------------------------------------- nearest upsample (NHWC bf16), factor f
// out[y, x, :] = in[y / f, x / f, :]   (F.interpolate(mode='nearest'), rrdbnet_arch.py:127-128,
// ssr_esrgan_model.py:133).  One thread per 16-byte channel vector of an output pixel.
__global__ void upsample_nearest_kernel(const __nv_bfloat16* __restrict__ src, int src_stride,
                                        __nv_bfloat16* __restrict__ dst, int dst_stride, int B, int H, int W, int C,
                                        int f) {
  const int groups = C / 8;
  const int OH = H * f, OW = W * f;
  const long total = (long)B * OH * OW * groups;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long p = i / groups;
    const int ox = (int)(p % OW);
    p /= OW;
    const int oy = (int)(p % OH);
    const long n = p / OH;
    const long sp = (n * H + oy / f) * W + ox / f;
    const uint4 v = *reinterpret_cast<const uint4*>(src + sp * src_stride + g * 8);
    *reinterpret_cast<uint4*>(dst + ((n * OH + oy) * OW + ox) * (long)dst_stride + g * 8) = v;
  }
}

// backward of nearest upsample: sum over each f x f block (fp32 accumulate, bf16 out)
__global__ void upsample_nearest_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_stride,
                                            __nv_bfloat16* __restrict__ dx, int dx_stride, int B, int H, int W,
                                            int C, int f, const __nv_bfloat16* __restrict__ mask, int mask_stride) {
  const int groups = C / 8;
  const int OH = H * f, OW = W * f;
  const long total = (long)B * H * W * groups;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long p = i / groups;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const long n = p / H;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int dyy = 0; dyy < f; ++dyy)
      for (int dxx = 0; dxx < f; ++dxx) {
        const long sp = (n * OH + y * f + dyy) * OW + x * f + dxx;
        const uint4 v = *reinterpret_cast<const uint4*>(dy + sp * dy_stride + g * 8);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(u[j] << 16);
          acc[2 * j + 1] += __uint_as_float(u[j] & 0xFFFF0000u);
        }
      }
    if (mask) {  // LeakyReLU(0.2) derivative from the saved activation
      const uint4 mv = *reinterpret_cast<const uint4*>(mask + ((n * H + y) * W + x) * (long)mask_stride + g * 8);
      const uint32_t mu[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] *= (__uint_as_float(mu[j] << 16) > 0.f ? 1.f : 0.2f);
        acc[2 * j + 1] *= (__uint_as_float(mu[j] & 0xFFFF0000u) > 0.f ? 1.f : 0.2f);
      }
    }
    uint4 o;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(acc[0], acc[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(acc[2], acc[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(acc[4], acc[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(acc[6], acc[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(dx + ((n * H + y) * W + x) * (long)dx_stride + g * 8) = o;
  }
}

// ---------------------------------------------------------------- bilinear x2, align_corners=False
// F.interpolate(scale_factor=2, mode='bilinear', align_corners=False), discriminator_arch.py:50,55,60.
// Source coordinate (o + 0.5)/2 - 0.5 clamped at 0: taps (0.25, 0.75) in the interior, replicate at edges.
// src2 (optional) is added to src before interpolating: the U-Net skip `x4 = x4 + x2` of discriminator_arch.py:53-64 (summed in
// f32; only the interpolated result is rounded to bf16).
// One thread = one INPUT pixel x 8 channels: it loads the 3 x 3 neighbourhood once (18 independent 16-byte loads with the skip)
// and writes the 2 x 2 output pixels it owns.  ncu (profiles/r02b_bilinear_ncu.md) showed the previous form issue-bound (43 % of
// the stalls "selected / not selected", the bf16 -> f32 shifts on top of the SASS list): every output re-expanded its four
// inputs and re-derived its taps.  Now every input word is expanded ONCE, the taps are constants and the interpolation is
// separable: with the clamped neighbourhood (index 0 = max(. - 1, 0), 2 = min(. + 1, size - 1)) output 2i reads (0, 1) with
// weights (0.25, 0.75) and output 2i + 1 reads (1, 2) with (0.75, 0.25) -- at an image edge both taps are the same pixel and
// fma(0.75, a, 0.25 a) = a exactly, which is what ATen's clamped source index gives.  Same operation order as ATen's
// upsample_bilinear2d (horizontal inside vertical): 62 instead of 92 instructions per 2 channels x 4 outputs.
__device__ __forceinline__ void bf16x2_to_f32(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xFFFF0000u);
}

__global__ void upsample_bilinear2x_kernel(const __nv_bfloat16* __restrict__ src, int src_stride,
                                           const __nv_bfloat16* __restrict__ src2, int src2_stride,
                                           __nv_bfloat16* __restrict__ dst, int dst_stride, int B, int H, int W, int C) {
  const int groups = C / 8;
  const int OW = W * 2;
  const long total = (long)B * H * W * groups;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long p = i / groups;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const long n = p / H;
    // rows / columns of the neighbourhood: index 0 = max(. - 1, 0), 1 = itself, 2 = min(. + 1, size - 1)
    const int ys[3] = {max(y - 1, 0), y, min(y + 1, H - 1)};
    const int xs[3] = {max(x - 1, 0), x, min(x + 1, W - 1)};
    uint32_t v[3][3][4];
    const __nv_bfloat16* base = src + n * H * W * (long)src_stride + g * 8;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint4 t = *reinterpret_cast<const uint4*>(base + ((long)ys[r] * W + xs[c]) * src_stride);
        v[r][c][0] = t.x, v[r][c][1] = t.y, v[r][c][2] = t.z, v[r][c][3] = t.w;
      }
    uint32_t u[3][3][4];
    if (src2) {
      const __nv_bfloat16* b2 = src2 + n * H * W * (long)src2_stride + g * 8;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint4 t = *reinterpret_cast<const uint4*>(b2 + ((long)ys[r] * W + xs[c]) * src2_stride);
          u[r][c][0] = t.x, u[r][c][1] = t.y, u[r][c][2] = t.z, u[r][c][3] = t.w;
        }
    }
    uint32_t o[2][2][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {          // two channels at a time
      float f[3][3][2];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          bf16x2_to_f32(v[r][c][j], f[r][c][0], f[r][c][1]);
          if (src2) {
            float a0, a1;
            bf16x2_to_f32(u[r][c][j], a0, a1);
            f[r][c][0] += a0;
            f[r][c][1] += a1;
          }
        }
      float hx[3][2][2];                   // [row][output column parity][channel]
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          hx[r][0][k] = fmaf(0.75f, f[r][1][k], 0.25f * f[r][0][k]);
          hx[r][1][k] = fmaf(0.25f, f[r][2][k], 0.75f * f[r][1][k]);
        }
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const float t0 = fmaf(0.75f, hx[1][dx][0], 0.25f * hx[0][dx][0]), t1 = fmaf(0.75f, hx[1][dx][1], 0.25f * hx[0][dx][1]);
        const float b0 = fmaf(0.25f, hx[2][dx][0], 0.75f * hx[1][dx][0]), b1 = fmaf(0.25f, hx[2][dx][1], 0.75f * hx[1][dx][1]);
        __nv_bfloat162 h = __floats2bfloat162_rn(t0, t1);
        o[0][dx][j] = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(b0, b1);
        o[1][dx][j] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx)
        *reinterpret_cast<uint4*>(dst + ((n * (2 * H) + 2 * y + dy) * OW + 2 * x + dx) * (long)dst_stride + g * 8) =
            make_uint4(o[dy][dx][0], o[dy][dx][1], o[dy][dx][2], o[dy][dx][3]);
  }
}

// backward (gather form): input pixel (y, x) collects from the 4 x 4 output pixels (2y - 1 .. 2y + 2) x (2x - 1 .. 2x + 2) whose
// taps touch it.  The weights are constants -- rows 2y - 1, 2y, 2y + 1, 2y + 2 contribute 0.25, 0.75, 0.75, 0.25; at the image
// edge the clamped source index of the forward folds the missing neighbour's weight onto the edge pixel: row 2y counts 1.0 at
// y = 0 and row 2y + 1 counts 1.0 at y = H - 1, rows outside the image count 0 -- and the sum is separable: horizontal inside
// vertical, 16 loads, every word expanded once (the loop over bilin_src() of round 1 spent 46 instructions per tap).
__global__ void upsample_bilinear2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_stride,
                                               __nv_bfloat16* __restrict__ dx, int dx_stride, int B, int H, int W,
                                               int C) {
  const int groups = C / 8;
  const int OH = H * 2, OW = W * 2;
  const long total = (long)B * H * W * groups;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long p = i / groups;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const long n = p / H;
    const float cy[4] = {y > 0 ? 0.25f : 0.f, y == 0 ? 1.f : 0.75f, y == H - 1 ? 1.f : 0.75f, y < H - 1 ? 0.25f : 0.f};
    const float cx[4] = {x > 0 ? 0.25f : 0.f, x == 0 ? 1.f : 0.75f, x == W - 1 ? 1.f : 0.75f, x < W - 1 ? 0.25f : 0.f};
    // rows / columns with weight 0 lie outside the image: read a valid address instead
    const int oys[4] = {max(2 * y - 1, 0), 2 * y, 2 * y + 1, min(2 * y + 2, OH - 1)};
    const int oxs[4] = {max(2 * x - 1, 0), 2 * x, 2 * x + 1, min(2 * x + 2, OW - 1)};
    const __nv_bfloat16* base = dy + n * OH * OW * (long)dy_stride + g * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      uint4 v[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) v[b] = *reinterpret_cast<const uint4*>(base + ((long)oys[a] * OW + oxs[b]) * dy_stride);
      float row[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t w4[4] = {v[b].x, v[b].y, v[b].z, v[b].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float lo, hi;
          bf16x2_to_f32(w4[j], lo, hi);
          row[2 * j] = fmaf(cx[b], lo, row[2 * j]);
          row[2 * j + 1] = fmaf(cx[b], hi, row[2 * j + 1]);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(cy[a], row[k], acc[k]);
    }
    uint4 o;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(acc[0], acc[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(acc[2], acc[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(acc[4], acc[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(acc[6], acc[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(dx + ((n * H + y) * W + x) * (long)dx_stride + g * 8) = o;
  }
}

// ---------------------------------------------------------------- batched weight packing
struct PackDesc {
  const float* w;          // OIHW f32
  __nv_bfloat16* dst;      // [chunks][R][R][n_pad][64]
  const float* inv_scale;  // sigma (device) or null
  int cout, cin, r, mode, k_pad, n_pad;
};

// One block = one tile of the packed operand: 64 K entries (one chunk) x kPackNT output rows x all r*r taps.  The f32 OIHW source
// is read in ITS order (forward: 64 * r*r consecutive floats per output channel; input-gradient form: kPackNT * r*r consecutive
// floats per K entry), transposed through shared memory, and written as whole 128-byte rows (kPackNT consecutive rows per tap =
// 1 KB contiguous) -- both sides coalesced, where the element-wise form read one 4-byte value per 36-byte stride.
static constexpr int kPackNT = 8;
// RT > 0: the filter size as a compile-time constant -- the index arithmetic below divides by r and r * r per element, and ncu
// (profiles/r02b_pack_ncu.md: 246 us for the generator's 702 operands at 5 % of the DRAM roofline, the stalls "math / not selected")
// showed the run-time divisions, not memory, bounding the kernel
template <int RT>
__device__ __forceinline__ void pack_tiles(const PackDesc& d, __nv_bfloat16* tile, int tl0, int tl_step) {
  const int r = RT > 0 ? RT : d.r;
  const bool s2 = d.mode == SSR_PACK_DGRAD_S2;        // r == 4: sixteen taps -> four parity classes of a 2 x 2 kernel
  const int chunks = d.k_pad / 64;
  const int T = r * r;
  const int n_tiles = (d.n_pad + kPackNT - 1) / kPackNT;
  const float sc = d.inv_scale ? 1.f / *d.inv_scale : 1.f;
  const bool fwd = d.mode == SSR_PACK_FWD;
  const int n_valid = fwd ? d.cout : d.cin;     // rows (n) that exist
  const int k_valid = fwd ? d.cin : d.cout;     // K entries that exist
  for (int tl = tl0; tl < chunks * n_tiles; tl += tl_step) {
    const int c = tl / n_tiles, n0 = (tl - c * n_tiles) * kPackNT;
    const int total = kPackNT * 64 * T;
    // six source elements per thread in flight (one dependent load -> shared-memory store per iteration made this kernel
    // latency bound: 17 us per 4608-element tile)
    constexpr int U = 6;
    for (int e0 = threadIdx.x; e0 < total; e0 += blockDim.x * U) {
      float vals[U];
      int slots[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + u * (int)blockDim.x;
        vals[u] = 0.f;
        slots[u] = -1;
        if (e >= total) continue;
        int nn, kl, tap;
        if (fwd) {            // e = (nn, kl, tap): source w[n][k][ky][kx] contiguous in (kl, tap)
          nn = e / (64 * T);
          const int rem = e - nn * 64 * T;
          kl = rem / T;
          tap = rem - kl * T;
        } else {              // e = (kl, nn, tap): source w[k][n][ky][kx] contiguous in (nn, tap)
          kl = e / (kPackNT * T);
          const int rem = e - kl * kPackNT * T;
          nn = rem / T;
          tap = rem - nn * T;
        }
        const int n = n0 + nn, k = c * 64 + kl;
        if (n < n_valid && k < k_valid) vals[u] = fwd ? __ldg(d.w + ((long)n * d.cin + k) * T + tap) : __ldg(d.w + ((long)k * d.cin + n) * T + tap);
        int ky = tap / r, kx = tap - ky * r;              // source tap (ky, kx); the input-gradient operand mirrors both
        int slot;
        if (s2) {
          // dx[2u + oy] collects w[ky] * dy[u - (1 - oy) + a] with ky = oy ? 2 - 2a : 3 - 2a  <=>  oy = 1 - (ky & 1), a = 1 - (ky >> 1)
          const int oy = 1 - (ky & 1), a = 1 - (ky >> 1), ox = 1 - (kx & 1), b = 1 - (kx >> 1);
          slot = (oy * 2 + ox) * 4 + b * 2 + a;
        } else {
          if (!fwd) {
            ky = r - 1 - ky;
            kx = r - 1 - kx;
          }
          slot = kx * r + ky;
        }
        slots[u] = (slot * kPackNT + nn) * 64 + kl;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (slots[u] >= 0) tile[slots[u]] = __float2bfloat16(vals[u] * sc);
    }
    __syncthreads();
    // rows out: dst[((c * r + kx) * r + ky) * n_pad + n][64], 16 bytes per thread
    const uint4* t4 = reinterpret_cast<const uint4*>(tile);
    for (int o = threadIdx.x; o < T * kPackNT * 8; o += blockDim.x) {
      const int tp = o / (kPackNT * 8), rem = o - tp * kPackNT * 8, nn = rem >> 3;
      const int n = n0 + nn;
      // plain: [chunk][tap][n][64]; parity classes: [class][chunk][2 x 2 tap][n][64]
      const long row = s2 ? (((long)(tp >> 2) * chunks + c) * 4 + (tp & 3)) * d.n_pad + n : ((long)c * T + tp) * d.n_pad + n;
      if (n < d.n_pad) reinterpret_cast<uint4*>(d.dst + row * 64)[rem & 7] = t4[o];
    }
    __syncthreads();
  }
}

__global__ void pack_batched_kernel(const PackDesc* __restrict__ descs) {
  const PackDesc d = descs[blockIdx.y];
  if (d.mode != SSR_PACK_FWD && d.mode != SSR_PACK_DGRAD && d.mode != SSR_PACK_DGRAD_S2) return;
  __shared__ __nv_bfloat16 tile[16 * kPackNT * 64];   // [tap (kx * r + ky)][row][k]
  if (d.r == 3) pack_tiles<3>(d, tile, (int)blockIdx.x, (int)gridDim.x);               // (block-uniform branch: one descriptor per block row)
  else if (d.r == 4) pack_tiles<4>(d, tile, (int)blockIdx.x, (int)gridDim.x);
  else pack_tiles<0>(d, tile, (int)blockIdx.x, (int)gridDim.x);
}

// the same with ONE tile per block from a host-built work list {descriptor, tile}: the generator's 702 operands have 4 .. 24 tiles each,
// so a fixed number of blocks per operand either idles most blocks or serialises three latency-bound tiles in one (174 us for 134 MB)
__global__ void pack_tiled_kernel(const PackDesc* __restrict__ descs, const int2* __restrict__ work) {
  const int2 e = work[blockIdx.x];
  const PackDesc d = descs[e.x];
  if (d.mode != SSR_PACK_FWD && d.mode != SSR_PACK_DGRAD && d.mode != SSR_PACK_DGRAD_S2) return;
  __shared__ __nv_bfloat16 tile[16 * kPackNT * 64];
  if (d.r == 3) pack_tiles<3>(d, tile, e.y, 1 << 30);
  else if (d.r == 4) pack_tiles<4>(d, tile, e.y, 1 << 30);
  else pack_tiles<0>(d, tile, e.y, 1 << 30);
}

// GEMM (1x1) forms for a conv computed through im2col: K' = (ky*R + kx)*cin + ci
__global__ void pack_gemm_kernel(const PackDesc* __restrict__ descs) {
  const PackDesc d = descs[blockIdx.y];
  if (d.mode != SSR_PACK_FWD_GEMM && d.mode != SSR_PACK_DGRAD_GEMM) return;
  const int chunks = d.k_pad / 64;
  const long total = (long)chunks * d.n_pad * 64;
  const float sc = d.inv_scale ? 1.f / *d.inv_scale : 1.f;
  const int kk = d.r * d.r * d.cin;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i % 64);
    const int nn = (int)((i / 64) % d.n_pad);
    const int c = (int)(i / (64L * d.n_pad));
    const int k = c * 64 + j;
    float v = 0.f;
    if (d.mode == SSR_PACK_FWD_GEMM) {
      if (nn < d.cout && k < kk) {
        const int tap = k / d.cin, ci = k % d.cin;
        v = d.w[(((long)nn * d.cin + ci) * d.r + tap / d.r) * d.r + tap % d.r];
      }
    } else {
      if (nn < kk && k < d.cout) {
        const int tap = nn / d.cin, ci = nn % d.cin;
        v = d.w[(((long)k * d.cin + ci) * d.r + tap / d.r) * d.r + tap % d.r];
      }
    }
    d.dst[i] = __float2bfloat16(v * sc);
  }
}

// ------------------------------------------------------------------ split-bf16 (tight parity) mode: the conv epilogue as a kernel
// v = act(s1 + s2 + s3 + bias) * s0 + w1 * r1 + w2 * r2 (every operand an NHWC f32 channel slice with its own pixel stride);
// out_f32 = v; then v *= LeakyReLU'(mask) when a mask (bf16 forward activation) is given; the bf16 pair (hi, lo) with hi + lo = v
// to ~2^-17 goes to hi / lo, channels [c, c_pad) zero-filled
__global__ void split_finish_kernel(const float* __restrict__ s1, const float* __restrict__ s2, const float* __restrict__ s3, int sum_stride,
                                    long npix, int c, const float* __restrict__ bias, int act, float s0, const float* __restrict__ r1,
                                    int r1_stride, float w1, const float* __restrict__ r2, int r2_stride, float w2,
                                    const __nv_bfloat16* __restrict__ mask, int mask_stride, float* __restrict__ out_f32, int out32_stride,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int out_stride, int c_pad) {
  const long total = npix * c_pad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long pix = i / c_pad;
    const int ch = (int)(i - pix * c_pad);
    float v = 0.f;
    if (ch < c) {
      v = s1[pix * sum_stride + ch];
      if (s2) v += s2[pix * sum_stride + ch];
      if (s3) v += s3[pix * sum_stride + ch];
      if (bias) v += bias[ch];
      if (act) v = v > 0.f ? v : 0.2f * v;
      v *= s0;
      if (r1) v = fmaf(w1, r1[pix * r1_stride + ch], v);
      if (r2) v = fmaf(w2, r2[pix * r2_stride + ch], v);
      if (out_f32) out_f32[pix * out32_stride + ch] = v;
      if (mask) v *= __bfloat162float(mask[pix * mask_stride + ch]) > 0.f ? 1.f : 0.2f;
    }
    const __nv_bfloat16 h = __float2bfloat16(v);
    if (hi) hi[pix * out_stride + ch] = h;
    if (lo) lo[pix * out_stride + ch] = __float2bfloat16(v - __bfloat162float(h));
  }
}

// the adjoint of F.interpolate(scale_factor=2, mode='nearest') in f32 (rrdbnet_arch.py:127-134 backwards): dst[n, y, x, :] = sum of
// the 2 x 2 block of src[n, 2y .. 2y+1, 2x .. 2x+1, :]; NHWC f32, c channels
__global__ void sum_pool2x2_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int h, int w, int c) {
  const long total = (long)B * h * w * c;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c);
    long p = i / c;
    const int x = (int)(p % w);
    p /= w;
    const int y = (int)(p % h);
    const long n = p / h;
    const float* s = src + ((n * (2 * h) + 2 * y) * (long)(2 * w) + 2 * x) * c + ch;
    dst[i] = (s[0] + s[c]) + (s[(long)2 * w * c] + s[(long)2 * w * c + c]);
  }
}
}  // namespace ssr

using namespace ssr;

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int ssr_ingest_nchw(const void* src, int32_t src_kind /*0 = u8, 2 = f32*/, void* dst_bf16, int32_t dst_pix_stride,
                               int32_t b, int32_t c, int32_t h, int32_t w, int32_t c_pad, float scale,
                               const float* mean, const float* inv_std, void* stream) {
  SSR_REQUIRE(src && dst_bf16, "ssr_ingest_nchw: null pointer");
  SSR_REQUIRE(c_pad % 8 == 0 && c_pad >= c && dst_pix_stride % 8 == 0 && dst_pix_stride >= c_pad, "ssr_ingest_nchw: c_pad/stride");
  SSR_REQUIRE((reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0, "ssr_ingest_nchw: dst alignment");
  const long total = (long)b * h * w * (c_pad / 8);
  const int threads = 256;
  const int blocks = grid_for(total, threads);
  if (src_kind == 0)
    nchw_to_nhwc_bf16_kernel<uint8_t><<<blocks, threads, 0, STREAM(stream)>>>(
        reinterpret_cast<const uint8_t*>(src), reinterpret_cast<__nv_bfloat16*>(dst_bf16), b, c, h, w, dst_pix_stride,
        c_pad, scale, mean, inv_std);
  else if (src_kind == SSR_F32)
    nchw_to_nhwc_bf16_kernel<float><<<blocks, threads, 0, STREAM(stream)>>>(
        reinterpret_cast<const float*>(src), reinterpret_cast<__nv_bfloat16*>(dst_bf16), b, c, h, w, dst_pix_stride,
        c_pad, scale, mean, inv_std);
  else
    SSR_REQUIRE(false, "ssr_ingest_nchw: src_kind must be 0 (u8) or 2 (f32)");
  count_launch();
  return check_last("ingest launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_ingest_nchw_unshuffle(const float* src, void* dst_bf16, int32_t dst_pix_stride, int32_t b, int32_t c, int32_t h,
                                         int32_t w, int32_t factor, int32_t c_pad, float scale, void* stream) {
  SSR_REQUIRE(src && dst_bf16, "ssr_ingest_nchw_unshuffle: null pointer");
  SSR_REQUIRE(factor >= 1 && h % factor == 0 && w % factor == 0, "ssr_ingest_nchw_unshuffle: size not divisible by the factor");
  SSR_REQUIRE(c_pad % 8 == 0 && c_pad >= c * factor * factor && dst_pix_stride % 8 == 0 && dst_pix_stride >= c_pad,
              "ssr_ingest_nchw_unshuffle: c_pad/stride");
  SSR_REQUIRE((reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0, "ssr_ingest_nchw_unshuffle: dst alignment");
  const long total = (long)b * (h / factor) * (w / factor) * (c_pad / 8);
  nchw_unshuffle_to_nhwc_bf16_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), b, c, h,
                                                                                         w, factor, dst_pix_stride, c_pad, scale);
  count_launch();
  return check_last("ingest unshuffle launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_egress_nchw(const void* src_bf16, int32_t src_pix_stride, float* dst, int32_t b, int32_t c, int32_t h,
                               int32_t w, float scale, int32_t accumulate, const float* ch_scale, void* stream) {
  SSR_REQUIRE(src_bf16 && dst, "ssr_egress_nchw: null pointer");
  const long total = (long)b * c * h * w;
  nhwc_bf16_to_nchw_f32_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src_bf16), src_pix_stride, dst, b, c, h, w, scale, accumulate, ch_scale);
  count_launch();
  return check_last("egress launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_f32_nchw_to_u8_canvas(const float* src, void* dst_u8, int32_t b, int32_t c, int32_t h, int32_t w, int32_t canvas_w,
                                         int32_t grid_cols, int32_t first_index, void* stream) {
  SSR_REQUIRE(src && dst_u8 && grid_cols > 0 && canvas_w >= grid_cols * w, "ssr_f32_nchw_to_u8_canvas: bad args");
  const long total = (long)b * h * w;
  f32_nchw_to_u8_canvas_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(src, reinterpret_cast<uint8_t*>(dst_u8), b, c, h, w,
                                                                                canvas_w, grid_cols, first_index);
  count_launch();
  return check_last("f32_nchw_to_u8_canvas launch") ? SSR_OK : SSR_E_CUDA;
}

static int check_vec(const void* a, int sa, const void* b, int sb, int c, const char* who) {
  SSR_REQUIRE(a && b, "%s: null pointer", who);
  SSR_REQUIRE(c % 8 == 0 && sa % 8 == 0 && sb % 8 == 0, "%s: channels and strides must be multiples of 8", who);
  SSR_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0, "%s: 16-byte alignment", who);
  return SSR_OK;
}

extern "C" int ssr_upsample_nearest(const void* src, int32_t src_pix_stride, void* dst, int32_t dst_pix_stride, int32_t b,
                                    int32_t h, int32_t w, int32_t c, int32_t factor, void* stream) {
  if (int rc = check_vec(src, src_pix_stride, dst, dst_pix_stride, c, "ssr_upsample_nearest")) return rc;
  const long total = (long)b * h * factor * w * factor * (c / 8);
  upsample_nearest_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), src_pix_stride, reinterpret_cast<__nv_bfloat16*>(dst), dst_pix_stride, b,
      h, w, c, factor);
  count_launch();
  return check_last("upsample_nearest launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_upsample_nearest_bwd(const void* dy, int32_t dy_pix_stride, void* dx, int32_t dx_pix_stride, int32_t b,
                                        int32_t h, int32_t w, int32_t c, int32_t factor, const void* lrelu_mask,
                                        int32_t mask_pix_stride, void* stream) {
  if (int rc = check_vec(dy, dy_pix_stride, dx, dx_pix_stride, c, "ssr_upsample_nearest_bwd")) return rc;
  const long total = (long)b * h * w * (c / 8);
  upsample_nearest_bwd_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), dy_pix_stride, reinterpret_cast<__nv_bfloat16*>(dx), dx_pix_stride, b, h,
      w, c, factor, reinterpret_cast<const __nv_bfloat16*>(lrelu_mask), mask_pix_stride);
  count_launch();
  return check_last("upsample_nearest_bwd launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_upsample_bilinear2x(const void* src, int32_t src_pix_stride, const void* src2, int32_t src2_pix_stride,
                                       void* dst, int32_t dst_pix_stride, int32_t b, int32_t h, int32_t w, int32_t c,
                                       void* stream) {
  if (int rc = check_vec(src, src_pix_stride, dst, dst_pix_stride, c, "ssr_upsample_bilinear2x")) return rc;
  const long total = (long)b * h * w * (c / 8);   // one thread per input pixel and 8 channels
  upsample_bilinear2x_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), src_pix_stride, reinterpret_cast<const __nv_bfloat16*>(src2),
      src2_pix_stride, reinterpret_cast<__nv_bfloat16*>(dst), dst_pix_stride, b, h, w, c);
  count_launch();
  return check_last("upsample_bilinear2x launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_upsample_bilinear2x_bwd(const void* dy, int32_t dy_pix_stride, void* dx, int32_t dx_pix_stride, int32_t b,
                                           int32_t h, int32_t w, int32_t c, void* stream) {
  if (int rc = check_vec(dy, dy_pix_stride, dx, dx_pix_stride, c, "ssr_upsample_bilinear2x_bwd")) return rc;
  const long total = (long)b * h * w * (c / 8);
  upsample_bilinear2x_bwd_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy), dy_pix_stride, reinterpret_cast<__nv_bfloat16*>(dx), dx_pix_stride, b, h,
      w, c);
  count_launch();
  return check_last("upsample_bilinear2x_bwd launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_pack_conv_weights_batched(const ssr_pack_desc* descs_device, int32_t n_layers, int32_t has_gemm_forms,
                                             void* stream) {
  SSR_REQUIRE(descs_device && n_layers > 0, "ssr_pack_conv_weights_batched: bad args");
  static_assert(sizeof(ssr_pack_desc) == sizeof(PackDesc), "ssr_pack_desc layout");
  // blocks per layer: the generator has 702 small operands (4 .. 24 tiles of 4608 elements each) -- more than a few blocks per
  // layer only adds block-launch overhead (22 k mostly empty blocks cost 0.2 ms); the discriminator has 20 large ones
  dim3 grid(n_layers > 64 ? 8u : 96u, (unsigned)n_layers);
  pack_batched_kernel<<<grid, 256, 0, STREAM(stream)>>>(reinterpret_cast<const PackDesc*>(descs_device));
  count_launch();
  if (has_gemm_forms) {
    dim3 grid_g(148, (unsigned)n_layers);
    pack_gemm_kernel<<<grid_g, 256, 0, STREAM(stream)>>>(reinterpret_cast<const PackDesc*>(descs_device));
    count_launch();
  }
  return check_last("pack_batched launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int32_t ssr_pack_tile_count(int32_t k_pad, int32_t n_pad) { return (k_pad / 64) * ((n_pad + kPackNT - 1) / kPackNT); }

extern "C" int ssr_pack_conv_weights_tiled(const ssr_pack_desc* descs_device, const int32_t* work_device, int32_t n_work, void* stream) {
  SSR_REQUIRE(descs_device && work_device && n_work > 0, "ssr_pack_conv_weights_tiled: bad args");
  pack_tiled_kernel<<<(unsigned)n_work, 256, 0, STREAM(stream)>>>(reinterpret_cast<const PackDesc*>(descs_device),
                                                                 reinterpret_cast<const int2*>(work_device));
  count_launch();
  return check_last("pack_tiled launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_split_finish(const float* s1, const float* s2, const float* s3, int32_t sum_stride, int64_t npix, int32_t c,
                                const float* bias, int32_t act, float s0, const float* r1, int32_t r1_stride, float w1, const float* r2,
                                int32_t r2_stride, float w2, const void* mask_bf16, int32_t mask_stride, float* out_f32,
                                int32_t out32_stride, void* hi, void* lo, int32_t out_stride, int32_t c_pad, void* stream) {
  SSR_REQUIRE(s1 && npix > 0 && c > 0 && c_pad >= c && sum_stride >= c, "ssr_split_finish: bad args");
  SSR_REQUIRE((hi == nullptr && lo == nullptr) || out_stride >= c_pad, "ssr_split_finish: out_stride %d < c_pad %d", out_stride, c_pad);
  SSR_REQUIRE((!r1 || r1_stride >= c) && (!r2 || r2_stride >= c) && (!out_f32 || out32_stride >= c) && (!mask_bf16 || mask_stride >= c),
              "ssr_split_finish: an operand's pixel stride is smaller than the channel count %d", c);
  const long total = (long)npix * c_pad;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 16 * 148) blocks = 16 * 148;
  split_finish_kernel<<<blocks, 256, 0, STREAM(stream)>>>(s1, s2, s3, sum_stride, npix, c, bias, act, s0, r1, r1_stride, w1, r2, r2_stride, w2,
                                                        reinterpret_cast<const __nv_bfloat16*>(mask_bf16), mask_stride, out_f32, out32_stride,
                                                        reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), out_stride, c_pad);
  count_launch();
  return check_last("split_finish") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_sum_pool2x2_f32(const float* src, float* dst, int32_t b, int32_t h, int32_t w, int32_t c, void* stream) {
  SSR_REQUIRE(src && dst && b > 0 && h > 0 && w > 0 && c > 0, "ssr_sum_pool2x2_f32: bad args");
  const long total = (long)b * h * w * c;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 16 * 148) blocks = 16 * 148;
  sum_pool2x2_f32_kernel<<<blocks, 256, 0, STREAM(stream)>>>(src, dst, b, h, w, c);
  count_launch();
  return check_last("sum_pool2x2_f32") ? SSR_OK : SSR_E_CUDA;
}
