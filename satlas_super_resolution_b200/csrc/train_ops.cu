// Training-step kernels around the tensor-core convs (all HBM-bound, vectorised / coalesced):
// im2col / col2im for the strided discriminator convs, max-pool for VGG, the L1 / BCE / feature-L1 losses with
// their gradients, spectral-norm power iteration and its backward, USM sharpening, fused Adam + EMA.
#include "common.cuh"

namespace ssr {

static int g_sms2 = 0;
static int grid_for2(long work_items, int threads) {
  if (g_sms2 == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms2, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms2 <= 0) g_sms2 = 148;
  }
  long blocks = (work_items + threads - 1) / threads;
  long cap = (long)g_sms2 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(u[j] << 16);
    f[2 * j + 1] = __uint_as_float(u[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  __nv_bfloat162 h;
  h = __floats2bfloat162_rn(f[0], f[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[2], f[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[4], f[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[6], f[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
  return o;
}

__device__ __forceinline__ float block_reduce_sum(float v) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  return v;  // valid in thread 0
}

// ------------------------------------------------------------------ im2col / col2im (k x k, stride s, pad p)
// col[m][(ky*k + kx)*C + c] = x[n, oy*s + ky - p, ox*s + kx - p, c]   (zero outside), m = (n*OH + oy)*OW + ox
__global__ void im2col_kernel(const __nv_bfloat16* __restrict__ x, int x_stride, __nv_bfloat16* __restrict__ col, int B, int H,
                              int W, int C, int k, int s, int p, int OH, int OW) {
  const int groups = C / 8;
  const long K = (long)k * k * C;
  const long total = (long)B * OH * OW * k * k * groups;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long t = i / groups;
    const int tap = (int)(t % (k * k));
    const long m = t / (k * k);
    const int ox = (int)(m % OW);
    const long t2 = m / OW;
    const int oy = (int)(t2 % OH);
    const long n = t2 / OH;
    const int iy = oy * s + tap / k - p, ix = ox * s + tap % k - p;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = *reinterpret_cast<const uint4*>(x + ((n * H + iy) * W + ix) * (long)x_stride + g * 8);
    *reinterpret_cast<uint4*>(col + m * K + (long)tap * C + g * 8) = v;
  }
}

// dx[n, iy, ix, c] = sum over taps with (iy + p - ky) % s == 0 ... of dcol[m(oy, ox)][tap][c]   (gather form)
__global__ void col2im_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx, int dx_stride, int B, int H,
                              int W, int C, int k, int s, int p, int OH, int OW, const __nv_bfloat16* __restrict__ add,
                              int add_stride, const __nv_bfloat16* __restrict__ mask, int mask_stride) {
  const int groups = C / 8;
  const long K = (long)k * k * C;
  const long total = (long)B * H * W * groups;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long t = i / groups;
    const int ix = (int)(t % W);
    t /= W;
    const int iy = (int)(t % H);
    const long n = t / H;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int ky = 0; ky < k; ++ky) {
      const int ny = iy + p - ky;
      if (ny < 0 || ny % s) continue;
      const int oy = ny / s;
      if (oy >= OH) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int nx = ix + p - kx;
        if (nx < 0 || nx % s) continue;
        const int ox = nx / s;
        if (ox >= OW) continue;
        const long m = (n * OH + oy) * OW + ox;
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(dcol + m * K + (long)(ky * k + kx) * C + g * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    const long pix = (n * H + iy) * W + ix;
    if (add) {
      float ad[8];
      unpack8(*reinterpret_cast<const uint4*>(add + pix * add_stride + g * 8), ad);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += ad[j];
    }
    if (mask) {
      float mk[8];
      unpack8(*reinterpret_cast<const uint4*>(mask + pix * mask_stride + g * 8), mk);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= (mk[j] > 0.f ? 1.f : 0.2f);
    }
    *reinterpret_cast<uint4*>(dx + pix * dx_stride + g * 8) = pack8(acc);
  }
}

// ------------------------------------------------------------------ elementwise helpers on NHWC bf16
// y = a*x1 (+ b*x2), optionally multiplied by the LeakyReLU(0.2) / ReLU derivative taken from `mask`
__global__ void axpby_kernel(const __nv_bfloat16* __restrict__ x1, int s1, float a, const __nv_bfloat16* __restrict__ x2, int s2,
                             float b, const __nv_bfloat16* __restrict__ mask, int sm, int mask_relu, __nv_bfloat16* __restrict__ y,
                             int sy, long npix, int C) {
  const int groups = C / 8;
  const long total = npix * groups;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long pix = i / groups;
    float f[8], f2[8];
    unpack8(*reinterpret_cast<const uint4*>(x1 + pix * s1 + g * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= a;
    if (x2) {
      unpack8(*reinterpret_cast<const uint4*>(x2 + pix * s2 + g * 8), f2);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(b, f2[j], f[j]);
    }
    if (mask) {
      unpack8(*reinterpret_cast<const uint4*>(mask + pix * sm + g * 8), f2);
      const float neg = mask_relu ? 0.f : 0.2f;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= (f2[j] > 0.f ? 1.f : neg);
    }
    *reinterpret_cast<uint4*>(y + pix * sy + g * 8) = pack8(f);
  }
}

// ------------------------------------------------------------------ VGG: relu(maxpool2x2) and its backward
// y = maxpool2x2(relu(x)) (= relu(maxpool(x))).  x is the PRE-ReLU conv output (the perceptual feature).
__global__ void maxpool_relu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H, int W, int C) {
  const int groups = C / 8;
  const int OH = H / 2, OW = W / 2;
  const long total = (long)B * OH * OW * groups;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long t = i / groups;
    const int ox = (int)(t % OW);
    t /= OW;
    const int oy = (int)(t % OH);
    const long n = t / OH;
    float m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(x + ((n * H + oy * 2 + (d >> 1)) * W + ox * 2 + (d & 1)) * (long)C + g * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
    }
    *reinterpret_cast<uint4*>(y + ((n * OH + oy) * OW + ox) * (long)C + g * 8) = pack8(m);
  }
}

// Gradient w.r.t. the pre-ReLU feature x of a perceptual layer:
//   dx = [x > 0 and x is the first maximum of its 2x2 window] * dpool[window]     (when dpool != null)
//      + lw/numel * sign(x - x_gt)                                                (the layer's own L1 term)
// x holds 2B images (generated | ground truth); only the first B get a gradient.
__global__ void feat_grad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dpool,
                                 __nv_bfloat16* __restrict__ dx, int B, int H, int W, int C, float l1_scale) {
  const int groups = C / 8;
  const long total = (long)B * H * W * groups;
  const long half = (long)B * H * W * C;  // offset of the ground-truth half
  const int OH = H / 2, OW = W / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long t = i / groups;
    const int ix = (int)(t % W);
    t /= W;
    const int iy = (int)(t % H);
    const long n = t / H;
    const long off = ((n * H + iy) * W + ix) * (long)C + g * 8;
    float f[8], fg[8], out[8];
    unpack8(*reinterpret_cast<const uint4*>(x + off), f);
    unpack8(*reinterpret_cast<const uint4*>(x + half + off), fg);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = f[j] - fg[j];
      out[j] = d > 0.f ? l1_scale : (d < 0.f ? -l1_scale : 0.f);
    }
    if (dpool) {
      const int oy = iy >> 1, ox = ix >> 1;
      float dp[8];
      unpack8(*reinterpret_cast<const uint4*>(dpool + ((n * OH + oy) * OW + ox) * (long)C + g * 8), dp);
      const int me = ((iy & 1) << 1) | (ix & 1);
      bool win[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) win[j] = f[j] > 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        if (d == me) continue;
        float o[8];
        unpack8(*reinterpret_cast<const uint4*>(x + ((n * H + oy * 2 + (d >> 1)) * W + ox * 2 + (d & 1)) * (long)C + g * 8), o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // an earlier element wins ties, a later one must be strictly larger
          if (d < me ? (o[j] >= f[j]) : (o[j] > f[j])) win[j] = false;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (win[j]) out[j] += dp[j];
    }
    *reinterpret_cast<uint4*>(dx + off) = pack8(out);
  }
}

// loss += lw * mean |x - x_gt| over the first B images' features
__global__ void feat_l1_kernel(const __nv_bfloat16* __restrict__ x, long n_half, float scale, float* __restrict__ loss) {
  float acc = 0.f;
  const long nvec = n_half / 8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
    float f[8], fg[8];
    unpack8(*reinterpret_cast<const uint4*>(x + i * 8), f);
    unpack8(*reinterpret_cast<const uint4*>(x + n_half + i * 8), fg);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += fabsf(f[j] - fg[j]);
  }
  acc = block_reduce_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * scale);
}

// ------------------------------------------------------------------ pixel losses on f32 tensors
// L1Loss(mean)*w : loss += w*mean|a-b| ; grad (+)= w*sign(a-b)/n
__global__ void l1_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, float w, float* __restrict__ loss,
                               float* __restrict__ grad, int accumulate) {
  float acc = 0.f;
  const float gs = w / (float)n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    acc += fabsf(d);
    if (grad) {
      const float gv = d > 0.f ? gs : (d < 0.f ? -gs : 0.f);
      if (accumulate) grad[i] += gv; else grad[i] = gv;
    }
  }
  acc = block_reduce_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * gs);
}

// BCEWithLogitsLoss(mean) against a constant target t: loss += w*mean(...), grad = w*(sigmoid(x)-t)/n,
// mean_logit += mean(x)    (GANLoss 'vanilla', ssr_esrgan_model.py:182,218-226)
__global__ void bce_logits_kernel(const float* __restrict__ x, long n, float target, float w, float* __restrict__ loss,
                                  float* __restrict__ mean_logit, float* __restrict__ grad) {
  float acc = 0.f, accx = 0.f;
  const float inv_n = 1.f / (float)n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = x[i];
    // (1-t)*x + softplus(-x), computed the way ATen does: max(-x,0) + log(exp(-max(-x,0)) + exp(-x-max(-x,0)))
    const float mx = fmaxf(-v, 0.f);
    acc += (1.f - target) * v + mx + logf(expf(-mx) + expf(-v - mx));
    accx += v;
    if (grad) grad[i] = w * inv_n * (1.f / (1.f + expf(-v)) - target);
  }
  acc = block_reduce_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_n * w);
  if (mean_logit) {
    accx = block_reduce_sum(accx);
    if (threadIdx.x == 0) atomicAdd(mean_logit, accx * inv_n);
  }
}

// ------------------------------------------------------------------ SSIMLoss (ssr/losses/basic_loss.py:50-60)
// kornia.losses.ssim_loss(x, gt, window_size=5, reduction="none"): 5x5 Gaussian window (sigma 1.5), reflect-padded 'same' filtering,
//   mu = F(img), sigma = F(img * img') - mu mu', ssim = (2 mu_x mu_y + C1)(2 sigma_xy + C2) / ((mu_x^2 + mu_y^2 + C1)(sigma_xx + sigma_yy + C2) + eps),
//   loss map = clamp((1 - ssim) / 2, 0, 1); the module takes the mean over (C, H, W) and then over the batch = the mean over everything.
// Planes are [n, h, w] f32 (n = batch x channels).  Forward: one thread per pixel reads its 5 x 5 window of x and y (reflected at the
// border), adds weight * mean(loss map) into *loss and -- when the gradient is wanted -- stores dLoss/d(mu_x), dLoss/d(E[x^2]),
// dLoss/d(E[xy]) of its window into `part` ([3][n*h*w]).  Backward: the adjoint of the reflect-padded filter applied to those
// three maps, gathered per pixel (no atomics: deterministic).
struct SsimWin {
  float k[5];
};
__device__ __forceinline__ int reflect_idx(int i, int n) {   // torch 'reflect' padding (also used by the USM blur below)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__global__ void ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int n, int h, int w, SsimWin win, float weight,
                                float* __restrict__ loss, float* __restrict__ part) {
  const long total = (long)n * h * w;
  const float gs = weight / (float)total;
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f, eps = 1e-12f;
  float acc = 0.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w), py = (int)((i / w) % h);
    const float* xp = x + (i - (long)py * w - px);
    const float* yp = y + (i - (long)py * w - px);
    float mx = 0.f, my = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
      const int ry = reflect_idx(py + dy - 2, h) * w;
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
        const int rx = reflect_idx(px + dx - 2, w);
        const float kw = win.k[dy] * win.k[dx];
        const float a = xp[ry + rx], b = yp[ry + rx];
        mx = fmaf(kw, a, mx);
        my = fmaf(kw, b, my);
        exx = fmaf(kw, a * a, exx);
        eyy = fmaf(kw, b * b, eyy);
        exy = fmaf(kw, a * b, exy);
      }
    }
    const float sxx = exx - mx * mx, syy = eyy - my * my, sxy = exy - mx * my;
    const float A1 = 2.f * mx * my + C1, A2 = 2.f * sxy + C2, B1 = mx * mx + my * my + C1, B2 = sxx + syy + C2;
    const float D = B1 * B2 + eps;
    const float s = A1 * A2 / D;
    const float l = 0.5f * (1.f - s);
    acc += fminf(fmaxf(l, 0.f), 1.f);
    if (part) {
      const float gl = (l > 0.f && l < 1.f) ? -0.5f * gs : 0.f;   // d(clamped loss)/d(ssim), times weight / N
      const float sd = s / D;
      part[i] = gl * (2.f * my * (A2 - A1) / D - sd * 2.f * mx * (B2 - B1));   // via mu_x (directly and through sigma_xy, sigma_xx)
      part[total + i] = gl * (-sd * B1);                                       // via E[x^2]
      part[2 * total + i] = gl * (2.f * A1 / D);                               // via E[xy]
    }
  }
  acc = block_reduce_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * gs);
}

// grad[p] (+)= F^T(g_mu)[p] + 2 x[p] F^T(g_xx)[p] + y[p] F^T(g_xy)[p];  F^T(g)[p] = sum over (q, d) with reflect(q + d) == p of k[d] g[q]
__global__ void ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ part, int n, int h, int w,
                                SsimWin win, float* __restrict__ grad, int accumulate) {
  const long total = (long)n * h * w;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int px = (int)(i % w), py = (int)((i / w) % h);
    const long base = i - (long)py * w - px;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int dy = 0; dy < 5; ++dy) {
      // the (up to three) rows q with reflect(q + dy - 2) == py: the direct one and the mirror images about either border
      int qy[3];
      qy[0] = py - (dy - 2);
      qy[1] = py >= 1 ? -py - (dy - 2) : -1;
      qy[2] = py <= h - 2 ? 2 * (h - 1) - py - (dy - 2) : -1;
      for (int cy = 0; cy < 3; ++cy) {
        const int ry = qy[cy];
        if (ry < 0 || ry >= h) continue;
        const int ly = ry + dy - 2;   // the unreflected coordinate: direct inside the image, mirrors strictly outside
        if (cy == 0 ? (ly < 0 || ly >= h) : (cy == 1 ? ly >= 0 : ly <= h - 1)) continue;
        for (int dx = 0; dx < 5; ++dx) {
          int qx[3];
          qx[0] = px - (dx - 2);
          qx[1] = px >= 1 ? -px - (dx - 2) : -1;
          qx[2] = px <= w - 2 ? 2 * (w - 1) - px - (dx - 2) : -1;
          const float kw = win.k[dy] * win.k[dx];
          for (int cx = 0; cx < 3; ++cx) {
            const int rx = qx[cx];
            if (rx < 0 || rx >= w) continue;
            const int lx = rx + dx - 2;
            if (cx == 0 ? (lx < 0 || lx >= w) : (cx == 1 ? lx >= 0 : lx <= w - 1)) continue;
            const long q = base + (long)ry * w + rx;
            t0 = fmaf(kw, part[q], t0);
            t1 = fmaf(kw, part[total + q], t1);
            t2 = fmaf(kw, part[2 * total + q], t2);
          }
        }
      }
    }
    const float g = t0 + 2.f * x[i] * t1 + y[i] * t2;
    if (accumulate) grad[i] += g; else grad[i] = g;
  }
}

// ------------------------------------------------------------------ discriminator input assembly
// out[n, y, x, :] = [ img[n, 0:ci, y, x] (planar f32) | lr[n, y/f, x/f, 0:cl] (NHWC bf16, nearest x f) | extra[n, 0:ce, y, x] (planar f32) | 0 ... ]
// = torch.cat((output | gt, F.interpolate(lr, scale_factor=4), old_hr), 1) of ssr_esrgan_model.py:133,171-176,202-210.
// One thread per pixel: the planar reads are coalesced across the warp (consecutive pixels), the NHWC row of the pixel
// (out_stride bf16, a multiple of 8) leaves as consecutive 16-byte stores.
__global__ void disc_input_kernel(const float* __restrict__ img, int ci, const __nv_bfloat16* __restrict__ lr, int lr_stride, int cl,
                                  int f, const float* __restrict__ extra, int ce, __nv_bfloat16* __restrict__ out, int out_stride,
                                  int B, int H, int W) {
  const long HW = (long)H * W;
  const long total = (long)B * HW;
  const int h = H / f, w = W / f;
  const int groups = out_stride >> 3;
  for (long pix = blockIdx.x * (long)blockDim.x + threadIdx.x; pix < total; pix += (long)gridDim.x * blockDim.x) {
    const long n = pix / HW;
    const long hw = pix - n * HW;
    const int y = (int)(hw / W), x = (int)(hw - (long)y * W);
    const __nv_bfloat16* lrp = cl ? lr + ((n * h + y / f) * w + x / f) * (long)lr_stride : nullptr;
    uint4* dst = reinterpret_cast<uint4*>(out + pix * out_stride);
    for (int g = 0; g < groups; ++g) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        float t = 0.f;
        if (c < ci) t = img[(n * ci + c) * HW + hw];
        else if (c < ci + cl) t = __bfloat162float(lrp[c - ci]);
        else if (c < ci + cl + ce) t = extra[(n * ce + (c - ci - cl)) * HW + hw];
        v[j] = t;
      }
      uint4 o;
      __nv_bfloat162 hh;
      hh = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&hh);
      hh = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&hh);
      hh = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&hh);
      hh = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&hh);
      dst[g] = o;
    }
  }
}

// ------------------------------------------------------------------ spectral norm (batched over layers)
struct SnDesc {
  const float* w;  // weight_orig viewed [rows][cols]
  float* u;        // [rows]
  float* v;        // [cols]
  float* sigma;    // [1]
  float* scratch;  // [cols + rows + 4]: t = W^T u | s = W t | norms
  float* geff;     // gradient w.r.t. the normalised weight, [rows][cols]
  float* grad;     // gradient of weight_orig (accumulated)
  int rows, cols;
};

// phase 1: t = W^T u ; nt2 = |t|^2        (threads walk columns: coalesced rows of W)
__global__ void sn_phase1(const SnDesc* __restrict__ descs) {
  // block = 32 columns x 8 row groups: a warp reads 128 contiguous bytes of one row, the eight warps walk the rows interleaved
  // (the first version gave every thread a whole column: up to 512 dependent loads per thread on 128 blocks)
  const SnDesc d = descs[blockIdx.y];
  float* t = d.scratch;
  float* norms = d.scratch + d.cols + d.rows;
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float local = 0.f;
  for (int j0 = blockIdx.x * 32; j0 < d.cols; j0 += gridDim.x * 32) {
    const int j = j0 + tx;
    float acc = 0.f;
    if (j < d.cols)
      for (int i = ty; i < d.rows; i += 8) acc = fmaf(d.w[(long)i * d.cols + j], d.u[i], acc);
    part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && j < d.cols) {
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += part[k][tx];
      t[j] = sum;
      local += sum * sum;
    }
    __syncthreads();
  }
  if (ty == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (tx == 0 && local != 0.f) atomicAdd(&norms[0], local);
  }
}
// phase 2: v = t / max(|t|, eps) ; s = W v ; ns2 = |s|^2     (one warp per row)
__global__ void sn_phase2(const SnDesc* __restrict__ descs, float eps) {
  const SnDesc d = descs[blockIdx.y];
  const float* t = d.scratch;
  float* s = d.scratch + d.cols;
  float* norms = d.scratch + d.cols + d.rows;
  const float inv = 1.f / fmaxf(sqrtf(norms[0]), eps);
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  // 16-byte loads when the rows allow it (every spectral-norm layer of the discriminator: cols = cin * k * k, a multiple of 16; the
  // flat parameter buffer aligns tensors to 256 bytes): four independent products per load instead of one dependent fma per 4 bytes
  const bool v4 = (d.cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(d.w) | reinterpret_cast<uintptr_t>(t)) & 15) == 0;
  for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < d.rows; i += gridDim.x * warps_per_block) {
    float acc = 0.f;
    if (v4) {
      const float4* w4 = reinterpret_cast<const float4*>(d.w + (long)i * d.cols);
      const float4* t4 = reinterpret_cast<const float4*>(t);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int j = lane; j < (d.cols >> 2); j += 32) {
        const float4 a = w4[j], b = t4[j];
        a0 = fmaf(a.x, b.x, a0);
        a1 = fmaf(a.y, b.y, a1);
        a2 = fmaf(a.z, b.z, a2);
        a3 = fmaf(a.w, b.w, a3);
      }
      acc = ((a0 + a1) + (a2 + a3)) * inv;
    } else {
      for (int j = lane; j < d.cols; j += 32) acc = fmaf(d.w[(long)i * d.cols + j], t[j] * inv, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      s[i] = acc;
      atomicAdd(&norms[1], acc * acc);
    }
  }
  // v is written by block 0 of each layer (t is complete: phase 1 finished before this launch)
  if (blockIdx.x == 0)
    for (int j = threadIdx.x; j < d.cols; j += blockDim.x) d.v[j] = t[j] * inv;
}
// phase 3: u = s / max(|s|, eps) ; sigma = u . s ; reset the norm accumulators for the next call
__global__ void sn_phase3(const SnDesc* __restrict__ descs, float eps) {
  const SnDesc d = descs[blockIdx.x];
  const float* s = d.scratch + d.cols;
  float* norms = d.scratch + d.cols + d.rows;
  const float ns2 = norms[1];
  const float inv = 1.f / fmaxf(sqrtf(ns2), eps);
  for (int i = threadIdx.x; i < d.rows; i += blockDim.x) d.u[i] = s[i] * inv;
  __syncthreads();
  if (threadIdx.x == 0) {
    *d.sigma = ns2 * inv;
    norms[0] = 0.f;
    norms[1] = 0.f;
    norms[2] = 0.f;
  }
}
// eval mode: sigma = u^T W v with the stored u, v (no iteration)
__global__ void sn_sigma_only(const SnDesc* __restrict__ descs) {
  const SnDesc d = descs[blockIdx.x];
  float acc = 0.f;
  for (long idx = threadIdx.x; idx < (long)d.rows * d.cols; idx += blockDim.x) {
    const int i = (int)(idx / d.cols), j = (int)(idx % d.cols);
    acc = fmaf(d.w[idx] * d.u[i], d.v[j], acc);
  }
  acc = block_reduce_sum(acc);
  if (threadIdx.x == 0) *d.sigma = acc;
}
// backward of w_eff = w / sigma, sigma = u^T w v:  grad += geff / sigma - (<geff, w> / sigma^2) * u v^T
__global__ void sn_bwd_dot(const SnDesc* __restrict__ descs) {
  const SnDesc d = descs[blockIdx.y];
  float* norms = d.scratch + d.cols + d.rows;
  const long n = (long)d.rows * d.cols;
  float acc = 0.f;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(d.geff) | reinterpret_cast<uintptr_t>(d.w)) & 15) == 0) {
    const float4* g4 = reinterpret_cast<const float4*>(d.geff);
    const float4* w4 = reinterpret_cast<const float4*>(d.w);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int i = blockIdx.x * (int)blockDim.x + threadIdx.x; i < (int)(n >> 2); i += (int)(gridDim.x * blockDim.x)) {
      const float4 a = g4[i], b = w4[i];
      a0 = fmaf(a.x, b.x, a0);
      a1 = fmaf(a.y, b.y, a1);
      a2 = fmaf(a.z, b.z, a2);
      a3 = fmaf(a.w, b.w, a3);
    }
    acc = (a0 + a1) + (a2 + a3);
  } else {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) acc = fmaf(d.geff[i], d.w[i], acc);
  }
  acc = block_reduce_sum(acc);
  if (threadIdx.x == 0) atomicAdd(&norms[2], acc);
}
__global__ void sn_bwd_apply(const SnDesc* __restrict__ descs) {
  const SnDesc d = descs[blockIdx.y];
  const float* norms = d.scratch + d.cols + d.rows;
  const float sigma = *d.sigma;
  const float coef = norms[2] / (sigma * sigma);
  const float inv = 1.f / sigma;
  const long n = (long)d.rows * d.cols;
  if ((d.cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(d.geff) | reinterpret_cast<uintptr_t>(d.grad) | reinterpret_cast<uintptr_t>(d.v)) & 15) == 0) {
    // a float4 never straddles two rows (cols % 4 == 0); 32-bit index arithmetic (n < 2^31), one division per 4 elements
    const int q = d.cols >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(d.geff);
    const float4* v4 = reinterpret_cast<const float4*>(d.v);
    float4* o4 = reinterpret_cast<float4*>(d.grad);
    for (int i = blockIdx.x * (int)blockDim.x + threadIdx.x; i < (int)(n >> 2); i += (int)(gridDim.x * blockDim.x)) {
      const int r = i / q, c = i - r * q;
      const float cu = coef * d.u[r];
      const float4 g = g4[i], v = v4[c];
      float4 o = o4[i];
      o.x += g.x * inv - cu * v.x;
      o.y += g.y * inv - cu * v.y;
      o.z += g.z * inv - cu * v.z;
      o.w += g.w * inv - cu * v.w;
      o4[i] = o;
    }
    return;
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / d.cols), c = (int)(i % d.cols);
    d.grad[i] += d.geff[i] * inv - coef * d.u[r] * d.v[c];
  }
}
__global__ void sn_bwd_reset(const SnDesc* __restrict__ descs) {
  const SnDesc d = descs[blockIdx.x];
  if (threadIdx.x == 0) d.scratch[d.cols + d.rows + 2] = 0.f;
}

// ------------------------------------------------------------------ USM sharpening (basicsr USMSharp, radius 51, sigma 8)
__constant__ float c_gauss[64];
// horizontal (dir=0) or vertical (dir=1) 1-D blur with reflect padding, planes of H x W
__global__ void blur1d_kernel(const float* __restrict__ src, float* __restrict__ dst, long planes, int H, int W, int taps, int dir) {
  const long total = planes * H * W;
  const int r = taps / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long t = i / W;
    const int y = (int)(t % H);
    const long pl = t / H;
    const float* base = src + pl * H * W;
    float acc = 0.f;
    // interior pixels (61 % of a 128-wide row at radius 25, whole warps of them): no reflection, a running pointer -- the reflected
    // index cost six instructions per tap and made the four blur passes instruction-bound (2 % of the HBM roofline)
    if (dir == 0) {
      if (x >= r && x + r < W) {
        const float* q = base + (long)y * W + (x - r);
#pragma unroll 17
        for (int k = 0; k < taps; ++k) acc = fmaf(c_gauss[k], q[k], acc);
      } else {
        for (int k = 0; k < taps; ++k) acc = fmaf(c_gauss[k], base[(long)y * W + reflect_idx(x + k - r, W)], acc);
      }
    } else {
      if (y >= r && y + r < H) {
        const float* q = base + (long)(y - r) * W + x;
#pragma unroll 17
        for (int k = 0; k < taps; ++k) acc = fmaf(c_gauss[k], q[(long)k * W], acc);
      } else {
        for (int k = 0; k < taps; ++k) acc = fmaf(c_gauss[k], base[(long)reflect_idx(y + k - r, H) * W + x], acc);
      }
    }
    dst[i] = acc;
  }
}
__global__ void usm_mask_kernel(const float* __restrict__ img, const float* __restrict__ blur, float* __restrict__ mask, long n,
                                float threshold) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    mask[i] = (fabsf(img[i] - blur[i]) * 255.f > threshold) ? 1.f : 0.f;
}
__global__ void usm_blend_kernel(const float* __restrict__ img, const float* __restrict__ blur, const float* __restrict__ soft,
                                 float* __restrict__ out, long n, float weight) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float im = img[i];
    const float sharp = fminf(fmaxf(im + weight * (im - blur[i]), 0.f), 1.f);
    const float s = soft[i];
    out[i] = s * sharp + (1.f - s) * im;
  }
}
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long n, float scale) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) dst[i] = (float)src[i] * scale;
}

// ------------------------------------------------------------------ fused Adam (+ EMA) over flat f32 buffers
// torch.optim.Adam (no amsgrad): m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= step_size * m / (sqrt(v)/sqrt(bc2) + eps)
// then, when ema != null: ema = decay*ema + (1-decay)*p    (basicsr model_ema, ssr_esrgan_model.py:230-231)
__global__ void adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                float* __restrict__ ema, long n, float lr, float b1, float b2, float eps, float wd, float bc1,
                                float bc2_sqrt, float ema_decay, float grad_scale, const float* __restrict__ dev_hyper) {
  if (dev_hyper) {  // CUDA-graph replay: the step-dependent scalars live in device memory
    lr = dev_hyper[0];
    bc1 = dev_hyper[1];
    bc2_sqrt = dev_hyper[2];
  }
  const float step_size = lr / bc1;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi;
    if (ema) ema[i] = ema_decay * ema[i] + (1.f - ema_decay) * pi;
  }
}

// basicsr model_ema on an iteration without a generator step (net_d_iters > 1 / net_d_init_iters): ema = decay*ema + (1-decay)*p
__global__ void ema_update_kernel(float* __restrict__ ema, const float* __restrict__ p, long n, float decay) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    ema[i] = decay * ema[i] + (1.f - decay) * p[i];
}

// hyper = [lr, 1 - b1^t, sqrt(1 - b2^t), t, b1, b2]: advance t on the device (CUDA-graph replays cannot take new kernel
// arguments, and a host-staged copy could be overwritten by a CPU that runs several steps ahead)
__global__ void adam_tick_kernel(float* __restrict__ hyper) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float t = hyper[3] + 1.f;
    hyper[3] = t;
    hyper[1] = 1.f - powf(hyper[4], t);
    hyper[2] = sqrtf(1.f - powf(hyper[5], t));
  }
}

}  // namespace ssr

using namespace ssr;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define LAUNCH_OK(name) (count_launch(), check_last(name) ? SSR_OK : SSR_E_CUDA)

extern "C" int ssr_im2col(const void* x, int32_t x_pix_stride, void* col, int32_t b, int32_t h, int32_t w, int32_t c, int32_t k,
                          int32_t s, int32_t p, void* stream) {
  SSR_REQUIRE(x && col && c % 8 == 0 && x_pix_stride % 8 == 0, "ssr_im2col: bad args");
  const int oh = (h + 2 * p - k) / s + 1, ow = (w + 2 * p - k) / s + 1;
  const long total = (long)b * oh * ow * k * k * (c / 8);
  im2col_kernel<<<grid_for2(total, 256), 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), x_pix_stride,
                                                                  reinterpret_cast<__nv_bfloat16*>(col), b, h, w, c, k, s, p, oh, ow);
  return LAUNCH_OK("im2col");
}

extern "C" int ssr_col2im(const void* dcol, void* dx, int32_t dx_pix_stride, int32_t b, int32_t h, int32_t w, int32_t c, int32_t k,
                          int32_t s, int32_t p, const void* add, int32_t add_pix_stride, const void* lrelu_mask,
                          int32_t mask_pix_stride, void* stream) {
  SSR_REQUIRE(dcol && dx && c % 8 == 0 && dx_pix_stride % 8 == 0, "ssr_col2im: bad args");
  const int oh = (h + 2 * p - k) / s + 1, ow = (w + 2 * p - k) / s + 1;
  const long total = (long)b * h * w * (c / 8);
  col2im_kernel<<<grid_for2(total, 256), 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(dcol),
                                                                  reinterpret_cast<__nv_bfloat16*>(dx), dx_pix_stride, b, h, w, c, k, s, p,
                                                                  oh, ow, reinterpret_cast<const __nv_bfloat16*>(add), add_pix_stride,
                                                                  reinterpret_cast<const __nv_bfloat16*>(lrelu_mask), mask_pix_stride);
  return LAUNCH_OK("col2im");
}

extern "C" int ssr_axpby(const void* x1, int32_t s1, float a, const void* x2, int32_t s2, float b, const void* mask, int32_t sm,
                         int32_t mask_relu, void* y, int32_t sy, int64_t npix, int32_t c, void* stream) {
  SSR_REQUIRE(x1 && y && c % 8 == 0 && s1 % 8 == 0 && sy % 8 == 0, "ssr_axpby: bad args");
  axpby_kernel<<<grid_for2(npix * (c / 8), 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x1), s1, a, reinterpret_cast<const __nv_bfloat16*>(x2), s2, b,
      reinterpret_cast<const __nv_bfloat16*>(mask), sm, mask_relu, reinterpret_cast<__nv_bfloat16*>(y), sy, npix, c);
  return LAUNCH_OK("axpby");
}

extern "C" int ssr_maxpool_relu(const void* x, void* y, int32_t b, int32_t h, int32_t w, int32_t c, void* stream) {
  SSR_REQUIRE(x && y && c % 8 == 0 && h % 2 == 0 && w % 2 == 0, "ssr_maxpool_relu: bad args");
  const long total = (long)b * (h / 2) * (w / 2) * (c / 8);
  maxpool_relu_kernel<<<grid_for2(total, 256), 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                        reinterpret_cast<__nv_bfloat16*>(y), b, h, w, c);
  return LAUNCH_OK("maxpool_relu");
}

extern "C" int ssr_feat_grad(const void* x, const void* dpool, void* dx, int32_t b, int32_t h, int32_t w, int32_t c, float l1_scale,
                             void* stream) {
  SSR_REQUIRE(x && dx && c % 8 == 0, "ssr_feat_grad: bad args");
  const long total = (long)b * h * w * (c / 8);
  feat_grad_kernel<<<grid_for2(total, 256), 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                     reinterpret_cast<const __nv_bfloat16*>(dpool),
                                                                     reinterpret_cast<__nv_bfloat16*>(dx), b, h, w, c, l1_scale);
  return LAUNCH_OK("feat_grad");
}

extern "C" int ssr_feat_l1(const void* x, int64_t n_half, float scale, float* loss, void* stream) {
  SSR_REQUIRE(x && loss && n_half % 8 == 0, "ssr_feat_l1: bad args");
  feat_l1_kernel<<<grid_for2(n_half / 8, 256), 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), n_half, scale, loss);
  return LAUNCH_OK("feat_l1");
}

extern "C" int ssr_l1_loss(const float* a, const float* b, int64_t n, float weight, float* loss, float* grad, int32_t accumulate,
                           void* stream) {
  SSR_REQUIRE(a && b && loss && n > 0, "ssr_l1_loss: bad args");
  l1_loss_kernel<<<grid_for2(n, 256), 256, 0, STREAM(stream)>>>(a, b, n, weight, loss, grad, accumulate);
  return LAUNCH_OK("l1_loss");
}

extern "C" int ssr_ssim_loss(const float* x, const float* y, int32_t planes, int32_t h, int32_t w, float weight, float* loss, float* grad,
                             int32_t accumulate, float* scratch, void* stream) {
  SSR_REQUIRE(x && y && loss && planes > 0 && h >= 3 && w >= 3, "ssr_ssim_loss: bad args (planes of at least 3 x 3)");
  SSR_REQUIRE(grad == nullptr || scratch != nullptr, "ssr_ssim_loss: the gradient needs 3 * planes * h * w floats of scratch");
  SsimWin win;
  double g[5], sum = 0.0;
  for (int i = 0; i < 5; ++i) sum += (g[i] = exp(-(double)((i - 2) * (i - 2)) / (2.0 * 1.5 * 1.5)));   // kornia get_gaussian_kernel1d(5, 1.5)
  for (int i = 0; i < 5; ++i) win.k[i] = (float)(g[i] / sum);
  const long total = (long)planes * h * w;
  ssim_fwd_kernel<<<grid_for2(total, 256), 256, 0, STREAM(stream)>>>(x, y, planes, h, w, win, weight, loss, grad ? scratch : nullptr);
  if (grad) {
    count_launch();
    if (!check_last("ssim_fwd")) return SSR_E_CUDA;
    ssim_bwd_kernel<<<grid_for2(total, 256), 256, 0, STREAM(stream)>>>(x, y, scratch, planes, h, w, win, grad, accumulate);
  }
  return LAUNCH_OK("ssim_loss");
}

extern "C" int ssr_bce_logits(const float* x, int64_t n, float target, float weight, float* loss, float* mean_logit, float* grad,
                              void* stream) {
  SSR_REQUIRE(x && loss && n > 0, "ssr_bce_logits: bad args");
  bce_logits_kernel<<<grid_for2(n, 256), 256, 0, STREAM(stream)>>>(x, n, target, weight, loss, mean_logit, grad);
  return LAUNCH_OK("bce_logits");
}

extern "C" int ssr_disc_input_ex(const float* img, int32_t ci, const void* lr, int32_t lr_pix_stride, int32_t cl, int32_t factor,
                                 const float* extra, int32_t ce, void* out, int32_t out_pix_stride, int32_t b, int32_t h, int32_t w,
                                 void* stream) {
  SSR_REQUIRE(img && out && (cl == 0 || lr) && (ce == 0 || extra) && ci + cl + ce <= out_pix_stride, "ssr_disc_input: bad args");
  SSR_REQUIRE(out_pix_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "ssr_disc_input: out must be 16-byte aligned rows");
  SSR_REQUIRE(cl == 0 || (factor > 0 && h % factor == 0 && w % factor == 0), "ssr_disc_input: size not divisible by factor");
  const long total = (long)b * h * w;
  disc_input_kernel<<<grid_for2(total, 128), 128, 0, STREAM(stream)>>>(img, ci, reinterpret_cast<const __nv_bfloat16*>(lr), lr_pix_stride,
                                                                      cl, factor > 0 ? factor : 1, extra, ce,
                                                                      reinterpret_cast<__nv_bfloat16*>(out), out_pix_stride, b, h, w);
  return LAUNCH_OK("disc_input");
}

extern "C" int ssr_disc_input(const float* img, int32_t ci, const void* lr, int32_t lr_pix_stride, int32_t cl, int32_t factor, void* out,
                              int32_t out_pix_stride, int32_t b, int32_t h, int32_t w, void* stream) {
  return ssr_disc_input_ex(img, ci, lr, lr_pix_stride, cl, factor, nullptr, 0, out, out_pix_stride, b, h, w, stream);
}

extern "C" int ssr_spectral_norm(const ssr_sn_desc* descs_device, int32_t n_layers, int32_t power_iteration, float eps, void* stream) {
  static_assert(sizeof(ssr_sn_desc) == sizeof(SnDesc), "ssr_sn_desc layout");
  SSR_REQUIRE(descs_device && n_layers > 0, "ssr_spectral_norm: bad args");
  const SnDesc* d = reinterpret_cast<const SnDesc*>(descs_device);
  if (power_iteration) {
    sn_phase1<<<dim3(144, n_layers), 256, 0, STREAM(stream)>>>(d);   // 144 x 32 columns covers the widest layer (4608) in one pass
    count_launch();
    sn_phase2<<<dim3(64, n_layers), 256, 0, STREAM(stream)>>>(d, eps);   // 512 warps per layer: one row each for the widest (512 rows)
    count_launch();
    sn_phase3<<<n_layers, 256, 0, STREAM(stream)>>>(d, eps);
  } else {
    sn_sigma_only<<<n_layers, 1024, 0, STREAM(stream)>>>(d);
  }
  return LAUNCH_OK("spectral_norm");
}

extern "C" int ssr_spectral_norm_bwd(const ssr_sn_desc* descs_device, int32_t n_layers, void* stream) {
  SSR_REQUIRE(descs_device && n_layers > 0, "ssr_spectral_norm_bwd: bad args");
  const SnDesc* d = reinterpret_cast<const SnDesc*>(descs_device);
  sn_bwd_dot<<<dim3(128, n_layers), 256, 0, STREAM(stream)>>>(d);
  count_launch();
  sn_bwd_apply<<<dim3(256, n_layers), 256, 0, STREAM(stream)>>>(d);
  count_launch();
  sn_bwd_reset<<<n_layers, 32, 0, STREAM(stream)>>>(d);
  return LAUNCH_OK("spectral_norm_bwd");
}

extern "C" int ssr_usm_sharp(const float* img, float* out, float* scratch /* 3 x numel */, int32_t planes, int32_t h, int32_t w,
                             const float* gauss_host, int32_t taps, float weight, float threshold, void* stream) {
  SSR_REQUIRE(img && out && scratch && gauss_host && taps > 0 && taps <= 64 && (taps & 1), "ssr_usm_sharp: bad args");
  SSR_REQUIRE(h > taps / 2 && w > taps / 2, "ssr_usm_sharp: reflect padding needs size > radius");
  static float cached[64];
  static int cached_taps = 0;
  bool same = cached_taps == taps;
  for (int i = 0; same && i < taps; ++i) same = cached[i] == gauss_host[i];
  if (!same) {
    if (!check_cuda(cudaMemcpyToSymbolAsync(c_gauss, gauss_host, sizeof(float) * taps, 0, cudaMemcpyHostToDevice, STREAM(stream)),
                    "usm kernel upload"))
      return SSR_E_CUDA;
    for (int i = 0; i < taps; ++i) cached[i] = gauss_host[i];
    cached_taps = taps;
  }
  const long n = (long)planes * h * w;
  float* t0 = scratch;
  float* blur = scratch + n;
  float* t1 = scratch + 2 * n;
  const int blocks = grid_for2(n, 256);
  blur1d_kernel<<<blocks, 256, 0, STREAM(stream)>>>(img, t0, planes, h, w, taps, 0);
  blur1d_kernel<<<blocks, 256, 0, STREAM(stream)>>>(t0, blur, planes, h, w, taps, 1);
  usm_mask_kernel<<<blocks, 256, 0, STREAM(stream)>>>(img, blur, t0, n, threshold);
  blur1d_kernel<<<blocks, 256, 0, STREAM(stream)>>>(t0, t1, planes, h, w, taps, 0);
  blur1d_kernel<<<blocks, 256, 0, STREAM(stream)>>>(t1, t0, planes, h, w, taps, 1);
  usm_blend_kernel<<<blocks, 256, 0, STREAM(stream)>>>(img, blur, t0, out, n, weight);
  count_launch(5);
  return LAUNCH_OK("usm_sharp");
}

extern "C" int ssr_u8_to_f32(const void* src, float* dst, int64_t n, float scale, void* stream) {
  SSR_REQUIRE(src && dst && n > 0, "ssr_u8_to_f32: bad args");
  u8_to_f32_kernel<<<grid_for2(n, 256), 256, 0, STREAM(stream)>>>(reinterpret_cast<const uint8_t*>(src), dst, n, scale);
  return LAUNCH_OK("u8_to_f32");
}

extern "C" int ssr_adam_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr, float beta1, float beta2,
                            float eps, float weight_decay, int32_t step, float ema_decay, float grad_scale, const float* dev_hyper,
                            void* stream) {
  SSR_REQUIRE(p && g && m && v && n > 0 && (step >= 1 || dev_hyper), "ssr_adam_ema: bad args");
  if (step < 1) step = 1;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_ema_kernel<<<grid_for2(n, 256), 256, 0, STREAM(stream)>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                                 (float)sqrt(bc2), ema_decay, grad_scale, dev_hyper);
  return LAUNCH_OK("adam_ema");
}

extern "C" int ssr_ema_update(float* ema, const float* p, int64_t n, float decay, void* stream) {
  SSR_REQUIRE(ema && p && n > 0, "ssr_ema_update: bad args");
  ema_update_kernel<<<grid_for2(n, 256), 256, 0, STREAM(stream)>>>(ema, p, n, decay);
  return LAUNCH_OK("ema_update");
}

extern "C" int ssr_adam_tick(float* hyper_dev, void* stream) {
  SSR_REQUIRE(hyper_dev, "ssr_adam_tick: null pointer");
  adam_tick_kernel<<<1, 32, 0, STREAM(stream)>>>(hyper_dev);
  return LAUNCH_OK("adam_tick");
}
