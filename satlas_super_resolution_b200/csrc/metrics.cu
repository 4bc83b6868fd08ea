// Validation metrics on the GPU: what nondist_validation of ssr/models/ssr_esrgan_model.py:269-352 computes per
// image on the host -- basicsr tensor2img, calculate_psnr / calculate_ssim, and the 81-offset brute-force search of
// ssr/metrics/cpsnr.py:7-59 -- batched over the whole validation batch.  PSNR and cPSNR are reduced to EXACT integer sums on the
// device (differences of uint8 images), so the host's float64 formula reproduces the reference bit for bit up to its own rounding.
#include "common.cuh"

namespace ssr {

static int grid_for_m(long work_items, int threads) {
  long blocks = (work_items + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// basicsr tensor2img(rgb2bgr, out_type=uint8, min_max=(0, 1)): clamp, * 255, round half to even (np.round), HWC, optional RGB -> BGR
__global__ void f32_nchw_to_u8_hwc_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int B, int C, int H, int W, int reverse) {
  const long HW = (long)H * W;
  const long total = (long)B * HW * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long pix = i / C;
    const long n = pix / HW, hw = pix - n * HW;
    const int cs = reverse ? C - 1 - c : c;
    float v = src[(n * C + cs) * HW + hw];
    v = fminf(fmaxf(v, 0.f), 1.f);
    dst[i] = (uint8_t)__float2int_rn(v * 255.0f);
  }
}

// For image n, offset o = (ro, co) in [0, M]^2 and channel ch, over the (h - 2cb - M) x (w - 2cb - M) crop:
//   d = a[y + ro, x + co] - b[y + M - ro, x + M - co]   (both after removing `cb` border pixels);  S1 = sum d,  S2 = sum d^2
// out[((n * (M+1)^2 + o) * C + ch) * 2 + {0, 1}] (int64).  M = 0: the plain PSNR sums.
__global__ void u8_shift_diff_sums_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int H, int W, int C, int cb,
                                          int M, long long* __restrict__ out) {
  const int o = blockIdx.x, n = blockIdx.y;
  const int ro = o / (M + 1), co = o - ro * (M + 1);
  const int ch_ = H - 2 * cb - M, cw_ = W - 2 * cb - M;
  const uint8_t* pa = a + (long)n * H * W * C;
  const uint8_t* pb = b + (long)n * H * W * C;
  long long s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  for (int p = threadIdx.x; p < ch_ * cw_; p += blockDim.x) {
    const int y = p / cw_, x = p - y * cw_;
    const uint8_t* qa = pa + ((long)(y + cb + ro) * W + (x + cb + co)) * C;
    const uint8_t* qb = pb + ((long)(y + cb + M - ro) * W + (x + cb + M - co)) * C;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < C) {
        const int d = (int)qa[c] - (int)qb[c];
        s1[c] += d;
        s2[c] += (long long)(d * d);
      }
    }
  }
  // warp sums -> one 64-bit atomic per warp, channel and moment (the caller zeroes `out`; two's complement makes the unsigned add exact)
  const int lane = threadIdx.x & 31;
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(out) + (((long)n * (M + 1) * (M + 1) + o) * C) * 2;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c < C) {
      long long v1 = s1[c], v2 = s2[c];
      for (int off = 16; off; off >>= 1) {
        v1 += __shfl_xor_sync(0xffffffffu, v1, off);
        v2 += __shfl_xor_sync(0xffffffffu, v2, off);
      }
      if (lane == 0) {
        atomicAdd(dst + 2 * c, (unsigned long long)v1);
        atomicAdd(dst + 2 * c + 1, (unsigned long long)v2);
      }
    }
  }
}

// basicsr calculate_ssim / _ssim: 11 x 11 Gaussian window (sigma 1.5) over the 'valid' region, per channel; out[n * C + c] += sum of
// the SSIM map (float64), the caller divides by the (h - 2cb - 10) * (w - 2cb - 10) map size.
__global__ void u8_ssim_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int B, int H, int W, int C, int cb,
                               const double* __restrict__ win /* [11] */, double* __restrict__ out) {
  const int vh = H - 2 * cb - 10, vw = W - 2 * cb - 10;
  const long per = (long)vh * vw;
  const long total = (long)B * C * per;
  const double c1 = (0.01 * 255) * (0.01 * 255), c2 = (0.03 * 255) * (0.03 * 255);
  __shared__ double g[11];
  if (threadIdx.x < 11) g[threadIdx.x] = win[threadIdx.x];
  __syncthreads();
  for (long i0 = blockIdx.x * (long)blockDim.x; i0 < total; i0 += (long)gridDim.x * blockDim.x) {
    const long i = i0 + threadIdx.x;
    double val = 0.0;
    long key = -1;
    if (i < total) {
      const long nc = i / per, p = i - nc * per;
      const int n = (int)(nc / C), c = (int)(nc - (long)n * C);
      const int y = (int)(p / vw), x = (int)(p - (long)y * vw);
      const uint8_t* pa = a + ((long)n * H + y + cb) * W * C + (long)(x + cb) * C + c;
      const uint8_t* pb = b + ((long)n * H + y + cb) * W * C + (long)(x + cb) * C + c;
      double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
      for (int dy = 0; dy < 11; ++dy) {
        for (int dx = 0; dx < 11; ++dx) {
          const double wgt = g[dy] * g[dx];
          const double u = (double)pa[((long)dy * W + dx) * C], v = (double)pb[((long)dy * W + dx) * C];
          m1 += wgt * u;
          m2 += wgt * v;
          s11 += wgt * u * u;
          s22 += wgt * v * v;
          s12 += wgt * u * v;
        }
      }
      const double m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
      val = ((2 * m12 + c1) * (2 * (s12 - m12) + c2)) / ((m11 + m22 + c1) * ((s11 - m11) + (s22 - m22) + c2));
      key = nc;
    }
    // consecutive threads mostly share (n, c): reduce the warp when they all do, else add one by one
    const long k0 = __shfl_sync(0xffffffffu, key, 0);
    const bool same = __all_sync(0xffffffffu, key == k0);
    if (same) {
      for (int off = 16; off; off >>= 1) val += __shfl_xor_sync(0xffffffffu, val, off);
      if ((threadIdx.x & 31) == 0 && key >= 0) atomicAdd(out + key, val);
    } else if (key >= 0) {
      atomicAdd(out + key, val);
    }
  }
}

}  // namespace ssr

using namespace ssr;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int ssr_f32_nchw_to_u8_hwc(const float* src, void* dst_u8, int32_t b, int32_t c, int32_t h, int32_t w, int32_t reverse_channels,
                                      void* stream) {
  SSR_REQUIRE(src && dst_u8 && b > 0 && c > 0 && h > 0 && w > 0, "ssr_f32_nchw_to_u8_hwc: bad args");
  const long total = (long)b * c * h * w;
  f32_nchw_to_u8_hwc_kernel<<<grid_for_m(total, 256), 256, 0, STREAM(stream)>>>(src, reinterpret_cast<uint8_t*>(dst_u8), b, c, h, w,
                                                                               reverse_channels);
  count_launch();
  return check_last("f32_nchw_to_u8_hwc launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_u8_shift_diff_sums(const void* a_u8, const void* b_u8, int32_t b, int32_t h, int32_t w, int32_t c, int32_t crop_border,
                                      int32_t max_offset, long long* out, void* stream) {
  SSR_REQUIRE(a_u8 && b_u8 && out && b > 0 && c > 0 && c <= 4, "ssr_u8_shift_diff_sums: bad args (1..4 channels)");
  SSR_REQUIRE(crop_border >= 0 && max_offset >= 0 && max_offset <= 16 && h - 2 * crop_border - max_offset > 0 && w - 2 * crop_border - max_offset > 0,
              "ssr_u8_shift_diff_sums: crop larger than the image");
  dim3 grid((unsigned)((max_offset + 1) * (max_offset + 1)), (unsigned)b);
  u8_shift_diff_sums_kernel<<<grid, 256, 0, STREAM(stream)>>>(reinterpret_cast<const uint8_t*>(a_u8), reinterpret_cast<const uint8_t*>(b_u8), h, w,
                                                            c, crop_border, max_offset, out);
  count_launch();
  return check_last("u8_shift_diff_sums launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_u8_ssim_sums(const void* a_u8, const void* b_u8, int32_t b, int32_t h, int32_t w, int32_t c, int32_t crop_border,
                                const double* window11_device, double* out, void* stream) {
  SSR_REQUIRE(a_u8 && b_u8 && out && window11_device && b > 0 && c > 0, "ssr_u8_ssim_sums: bad args");
  SSR_REQUIRE(h - 2 * crop_border > 10 && w - 2 * crop_border > 10, "ssr_u8_ssim_sums: image smaller than the 11 x 11 window");
  const long total = (long)b * c * (h - 2 * crop_border - 10) * (w - 2 * crop_border - 10);
  u8_ssim_kernel<<<grid_for_m(total, 128), 128, 0, STREAM(stream)>>>(reinterpret_cast<const uint8_t*>(a_u8), reinterpret_cast<const uint8_t*>(b_u8),
                                                                    b, h, w, c, crop_border, window11_device, out);
  count_launch();
  return check_last("u8_ssim launch") ? SSR_OK : SSR_E_CUDA;
}
