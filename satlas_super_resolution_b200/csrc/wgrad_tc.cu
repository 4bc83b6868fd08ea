// Weight-gradient ("pixel-reduction") GEMM on tcgen05 tensor cores (sm_100a):
//
//   dW[ky][kx][cx][cy] += scale * sum_{pixels p} X[p + (ky - pad, kx - pad), cx] * dY[p, cy]
//
// Both operands are read straight from the NHWC bf16 activation / gradient buffers: a TMA box of 128
// pixels x 64 channels lands in shared memory as 128-byte rows (one row per pixel = one row per K index), which
// is exactly the MN-major 128B-swizzle operand layout of tcgen05.mma, so no transposed copy is ever made.
//   A (M = 128 input channels = two 64-channel boxes, LBO apart) = X window shifted by the tap,
//   B (N = cy, 64-byte rows / SWIZZLE_64B when cy <= 32)          = dY tile,
//   K = the 128 pixels of the tile (8 MMAs of K = 16); a CTA owns one (128-channel block, kx) pair, keeps
//   the R vertical taps in R TMEM accumulators and walks a range of pixel tiles (split over gridDim.x);
//   image borders are zero-filled by TMA.  The epilogue reduces the per-CTA partial sums with red.add.f32.
//
// Replaces the cuDNN wgrad behind autograd of every nn.Conv2d on the path
// (/root/reference/ssr/archs/rrdbnet_arch.py:26-30,99-112; discriminator_arch.py:28-40).
#include "common.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace ssr {

struct WgradK {
  int n_img, H, W, R, pad;
  int TW, TH, tiles_x, tiles_y, total_tiles, ksteps;
  int cx, cx_rows, cy, n_tile, y_blocks, y_rowbytes, out_stride;
  int stages, splits;
  uint32_t x_chunk_bytes, y_blk_bytes, stage_bytes, tmem_cols;
  float* out;
  float scale;
};

static constexpr int kWThreads = 192;

template <int R>
__global__ void __launch_bounds__(kWThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const WgradK p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* bar_empty = bar_full + p.stages;
  uint64_t* bar_tmem = bar_empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tmem + 1);

  // R = 1, 3: stride-1 conv, a CTA owns (128-channel block, kx) and keeps the R vertical taps in R accumulators.
  // R = 4: the 4 x 4 stride-2 conv of the discriminator -- X is gathered with element strides (2, 2); a CTA owns
  //        (128-channel block, kx, parity of ky) and keeps the two taps ky = parity, parity + 2 (consecutive box rows).
  constexpr int NH = R == 4 ? 8 : R;
  constexpr int NV = R == 4 ? 2 : R;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mtile = blockIdx.y / NH;
  const int hs = blockIdx.y % NH;
  const int kx = R == 4 ? (hs >> 1) : hs;
  const int par = R == 4 ? (hs & 1) : 0;
  const int n0 = blockIdx.z * p.n_tile;
  const int per = (p.total_tiles + p.splits - 1) / p.splits;
  const int t_begin = blockIdx.x * per;
  const int t_end = min(p.total_tiles, t_begin + per);
  const int iters = t_end - t_begin;
  if (iters <= 0) return;
  griddep_launch_dependents();
  const int nchunks = min(2, (p.cx - mtile * 128 + 63) / 64);
  const uint32_t x_bytes = 2 * p.x_chunk_bytes;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmX);
      prefetch_tmap(&tmY);
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&bar_full[s], 1);
        mbar_init(&bar_empty[s], 1);
      }
      mbar_init(bar_tmem, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 0) {
    // TMA producer: whole warp converged, one elected lane issues
    for (int it = 0; it < iters; ++it) {
      int t = t_begin + it;
      const int tx = t % p.tiles_x;
      t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int n = t / p.tiles_y;
      const int x0 = tx * p.TW, y0 = ty * p.TH;
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(&bar_empty[s], ph ^ 1);
      if (elect_one()) {
        uint8_t* xs = smem + (size_t)s * p.stage_bytes;
        uint8_t* ys = xs + x_bytes;
        mbar_expect_tx(&bar_full[s], nchunks * p.x_chunk_bytes + p.y_blocks * p.y_blk_bytes);
        for (int ch = 0; ch < nchunks; ++ch) {
          if (R == 4) tma_load_4d(xs + (size_t)ch * p.x_chunk_bytes, &tmX, &bar_full[s], mtile * 128 + ch * 64, 2 * x0 + kx - 1, 2 * y0 + par - 1, n);
          else tma_load_4d(xs + (size_t)ch * p.x_chunk_bytes, &tmX, &bar_full[s], mtile * 128 + ch * 64, x0 + kx - p.pad, y0 - p.pad, n);
        }
        for (int yb = 0; yb < p.y_blocks; ++yb)
          tma_load_4d(ys + (size_t)yb * p.y_blk_bytes, &tmY, &bar_full[s], n0 + yb * 64, x0, y0, n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // MMA issuer: descriptors differ only in the 14-bit start-address field -> one per stage + constant offsets
    const uint32_t idesc = umma_idesc_bf16(128u, (uint32_t)p.n_tile, 1u, 1u);
    const uint32_t y_layout = p.y_rowbytes == 128 ? 2u : 4u;
    const uint32_t a_tap = (uint32_t)(p.TW * 128) >> 4;         // one tile row down
    const uint32_t a_k = (16u * 128u) >> 4;                      // next 16 pixels of K
    const uint32_t b_k = (uint32_t)(16 * p.y_rowbytes) >> 4;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(&bar_full[s], ph);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t xs = smem_u32(smem + (size_t)s * p.stage_bytes);
        const uint64_t da0 = umma_desc(xs, p.x_chunk_bytes, 1024u, 2u);
        const uint64_t db0 = umma_desc(xs + x_bytes, p.y_blk_bytes, 8u * (uint32_t)p.y_rowbytes, y_layout);
        if (p.ksteps == 8) {
#pragma unroll
          for (int ky = 0; ky < NV; ++ky) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_bf16_ss(tmem_base + (uint32_t)(ky * p.n_tile), da0 + (ky * a_tap + ks * a_k), db0 + ks * b_k, idesc,
                           ks == 0 ? acc : 1u);
          }
        } else {
          for (int ky = 0; ky < NV; ++ky)
            for (int ks = 0; ks < p.ksteps; ++ks)
              umma_bf16_ss(tmem_base + (uint32_t)(ky * p.n_tile), da0 + (ky * a_tap + ks * a_k), db0 + ks * b_k, idesc,
                           ks == 0 ? acc : 1u);
        }
        umma_commit(&bar_empty[s]);
        if (it == iters - 1) umma_commit(bar_tmem);
      }
      __syncwarp();
      acc = 1;
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int cxi = mtile * 128 + row;
    const bool valid = cxi < p.cx;
    mbar_wait(bar_tmem, 0);
    tc_fence_after_sync();
    // accumulator layout [cy / 4][R*R taps * cx_rows][4]: 32 consecutive cx rows of one channel quad are 512 contiguous bytes
    const long plane = (long)R * R * p.cx_rows * 4;
#pragma unroll 1
    for (int vt = 0; vt < NV; ++vt) {
      const int ky = R == 4 ? par + 2 * vt : vt;
      float* dst = p.out + ((long)(ky * R + kx) * p.cx_rows + cxi) * 4;
#pragma unroll 1
      for (int cb = 0; cb < p.n_tile; cb += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(vt * p.n_tile + cb), v);
        tmem_ld_wait();
        if (!valid) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int co = n0 + cb + 4 * j;   // channels past cy inside the last quad get exact zeros (TMA zero fill of dY)
          if (co < p.out_stride)
            red_add_v4(dst + (long)(co >> 2) * plane, p.scale * __uint_as_float(v[4 * j]), p.scale * __uint_as_float(v[4 * j + 1]),
                       p.scale * __uint_as_float(v[4 * j + 2]), p.scale * __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// acc[cy / 4][taps * cx_rows][4] (f32) -> grad OIHW [cy][cx][R][R] += scale * acc
__global__ void wgrad_unpack_kernel(const float* __restrict__ acc, int cx_rows, int stride, float* __restrict__ grad,
                                    int cy, int cx, int r, float scale, int accumulate) {
  const long total = (long)cy * cx * r * r;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int kxx = t % r;
    t /= r;
    const int kyy = t % r;
    t /= r;
    const int ci = t % cx;
    const int co = t / cx;
    const float v = scale * acc[((long)(co >> 2) * (r * r * cx_rows) + (long)(kyy * r + kxx) * cx_rows + ci) * 4 + (co & 3)];
    if (accumulate) grad[i] += v; else grad[i] = v;
  }
}

struct UnpackDesc {
  const float* acc;
  float* grad;
  int cx_rows, acc_stride, cout, cin, r, accumulate;
  float scale;
  int pad_;
};
// RT > 0: the filter size as a compile-time constant (the index arithmetic divides by r twice per element; with run-time 64-bit
// divisions the kernel was instruction-bound: 0.19 ms per step at 16 % of the HBM roofline)
template <int RT>
__device__ __forceinline__ void wgrad_unpack_body(const UnpackDesc& d) {
  const int r = RT > 0 ? RT : d.r;
  const int T = r * r;
  const int total = d.cout * d.cin * T;              // < 2^31: the largest conv of the path has 2.4 M weights
  const int plane = T * d.cx_rows;                   // accumulator floats / 4 per output-channel quad
  for (int i = blockIdx.x * (int)blockDim.x + threadIdx.x; i < total; i += (int)gridDim.x * (int)blockDim.x) {
    const int tap = i % T;                           // OIHW order: (co, ci, ky, kx)
    const int t = i / T;
    const int ci = t % d.cin;
    const int co = t / d.cin;
    const float v = d.scale * d.acc[((long)(co >> 2) * plane + (long)tap * d.cx_rows + ci) * 4 + (co & 3)];
    if (d.accumulate) d.grad[i] += v; else d.grad[i] = v;
  }
}

__global__ void wgrad_unpack_batched_kernel(const UnpackDesc* __restrict__ descs) {
  const UnpackDesc d = descs[blockIdx.y];
  if (d.r == 3) wgrad_unpack_body<3>(d);             // (block-uniform branch: one descriptor per block row)
  else if (d.r == 4) wgrad_unpack_body<4>(d);
  else wgrad_unpack_body<0>(d);
}

// bias gradient: out[c] += scale * sum_p dy[p*stride + c]
__global__ void bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, int stride, long npix, int C, float* __restrict__ out,
                                 float scale) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C)
    for (long pth = blockIdx.y * 8L + threadIdx.y; pth < npix; pth += (long)gridDim.y * 8) acc += __bfloat162float(dy[pth * stride + c]);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    atomicAdd(out + c, scale * s);
  }
}

// the same with 16-byte loads (C and the pixel stride multiples of 8, C <= 256): a thread owns 8 channels of every (blockDim / groups)-th
// pixel; the scalar form above reads 2 bytes per thread and ran at 7 % of the HBM roofline (profiles/r02b_hbm_table.md)
__global__ void bias_grad_vec_kernel(const __nv_bfloat16* __restrict__ dy, int stride, long npix, int C, float* __restrict__ out, float scale) {
  __shared__ float red[256 * 8];
  const int groups = C >> 3;
  const int ppb = (int)blockDim.x / groups;             // pixels one block pass covers
  const int g = (int)threadIdx.x % groups, pl = (int)threadIdx.x / groups;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pl < ppb)
    for (long p = blockIdx.x * (long)ppb + pl; p < npix; p += (long)gridDim.x * ppb) {
      const uint4 v = *reinterpret_cast<const uint4*>(dy + p * stride + g * 8);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += __uint_as_float(u[j] << 16);
        acc[2 * j + 1] += __uint_as_float(u[j] & 0xFFFF0000u);
      }
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];   // thread t = (pl, g): channels g * 8 + j
  __syncthreads();
  if ((int)threadIdx.x < C) {
    const int c = (int)threadIdx.x;
    float sum = 0.f;
    for (int q = 0; q < ppb; ++q) sum += red[(q * groups + (c >> 3)) * 8 + (c & 7)];
    atomicAdd(out + c, scale * sum);
  }
}

// grouped form: channel c of dy belongs to output tensor outs[c / group_ch], element c % group_ch
__global__ void bias_grad_groups_kernel(const __nv_bfloat16* __restrict__ dy, int stride, long npix, int C, int group_ch,
                                        float* const* __restrict__ outs, float scale) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C)
    for (long pth = blockIdx.y * 8L + threadIdx.y; pth < npix; pth += (long)gridDim.y * 8) acc += __bfloat162float(dy[pth * stride + c]);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    atomicAdd(outs[c / group_ch] + (c % group_ch), scale * s);
  }
}

static int g_w_smem_optin = -1;

int launch_wgrad9(const ssr_wgrad_tc_args* a, cudaStream_t stream);  // wgrad9_tc.cu
int launch_wgrad9_batched(const ssr_wgrad_tc_args* args, int n, int* fallback, cudaStream_t stream);

}  // namespace ssr

using namespace ssr;

extern "C" int ssr_wgrad_tc(const ssr_wgrad_tc_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SSR_REQUIRE(a && a->x && a->dy && a->out, "ssr_wgrad_tc: null pointer");
  SSR_REQUIRE(a->r == 1 || a->r == 3 || a->r == 4, "ssr_wgrad_tc: r must be 1, 3 or 4 (the 4 x 4 stride-2 conv)");
  const bool s2 = a->r == 4;
  if (s2) SSR_REQUIRE(a->h % 2 == 0 && a->w % 2 == 0 && a->w >= 16, "ssr_wgrad_tc: r == 4: (h, w) is the even INPUT size of the stride-2 conv");
  SSR_REQUIRE(a->w >= 8 && a->h > 0 && a->n_img > 0, "ssr_wgrad_tc: geometry");
  SSR_REQUIRE(a->cx > 0 && a->cy > 0, "ssr_wgrad_tc: channels");
  SSR_REQUIRE(a->x_pix_stride % 8 == 0 && a->dy_pix_stride % 8 == 0, "ssr_wgrad_tc: strides must be multiples of 8");
  SSR_REQUIRE(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->dy)) & 15) == 0, "ssr_wgrad_tc: alignment");
  if (!s2) {
    // fast path: all nine taps from one halo tile per CTA (cy <= 64, power-of-two tile widths)
    const int rc = launch_wgrad9(a, stream);
    if (rc <= 0) return rc;
  }
  if (g_w_smem_optin < 0) {
    int dev = 0, v = 0;
    if (!check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return SSR_E_CUDA;
    if (!check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev), "smem attr")) return SSR_E_CUDA;
    g_w_smem_optin = v;
    if (!check_cuda(cudaFuncSetAttribute(wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, v), "cudaFuncSetAttribute(wgrad)"))
      return SSR_E_CUDA;
    if (!check_cuda(cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, v), "cudaFuncSetAttribute(wgrad)"))
      return SSR_E_CUDA;
    if (!check_cuda(cudaFuncSetAttribute(wgrad_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, v), "cudaFuncSetAttribute(wgrad)"))
      return SSR_E_CUDA;
  }
  WgradK p{};
  const int oh = s2 ? a->h / 2 : a->h, ow = s2 ? a->w / 2 : a->w;   // the dY grid the pixel tiles walk
  const int nh = s2 ? 8 : a->r, nv = s2 ? 2 : a->r;
  p.n_img = a->n_img; p.H = oh; p.W = ow; p.R = a->r; p.pad = a->r == 3 ? 1 : 0;
  p.TW = ow >= 128 ? 128 : round_up(ow, 8);
  p.TH = 128 / p.TW;
  while (p.TH > 1 && (p.TW * p.TH) % 16) --p.TH;
  SSR_REQUIRE((p.TW * p.TH) % 16 == 0, "ssr_wgrad_tc: tile %dx%d not a multiple of 16 pixels", p.TW, p.TH);
  p.ksteps = p.TW * p.TH / 16;
  p.tiles_x = (ow + p.TW - 1) / p.TW;
  p.tiles_y = (oh + p.TH - 1) / p.TH;
  p.total_tiles = p.tiles_x * p.tiles_y * a->n_img;
  p.cx = a->cx; p.cx_rows = a->out_cx_rows; p.cy = a->cy; p.out_stride = a->out_stride;
  SSR_REQUIRE(p.cx_rows >= p.cx && p.out_stride >= p.cy && p.out_stride % 4 == 0, "ssr_wgrad_tc: output too small / cy stride not a multiple of 4");
  SSR_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "ssr_wgrad_tc: out must be 16-byte aligned");
  if (a->cy <= 32) { p.n_tile = 32; p.y_blocks = 1; p.y_rowbytes = 64; }
  else if (a->cy <= 64) { p.n_tile = 64; p.y_blocks = 1; p.y_rowbytes = 128; }
  else { p.n_tile = 128; p.y_blocks = 2; p.y_rowbytes = 128; }
  const int n_tiles = (a->cy + p.n_tile - 1) / p.n_tile;
  const int mtiles = (a->cx + 127) / 128;
  p.x_chunk_bytes = (uint32_t)(p.TW * (p.TH + nv - 1)) * 128u;
  p.y_blk_bytes = (uint32_t)(p.TW * p.TH * p.y_rowbytes);
  p.stage_bytes = (uint32_t)round_up((int)(2 * p.x_chunk_bytes + p.y_blocks * p.y_blk_bytes), 1024);
  int stages = (g_w_smem_optin - 1280) / (int)p.stage_bytes;
  SSR_REQUIRE(stages >= 1, "ssr_wgrad_tc: stage does not fit shared memory");
  if (stages > 4) stages = 4;
  int units = mtiles * nh * n_tiles;
  static int target_ctas = -1;
  if (target_ctas < 0) {
    const char* e = getenv("SSR_WGRAD_CTAS");
    target_ctas = e ? atoi(e) : 148;
  }
  int splits = a->splits > 0 ? a->splits : (target_ctas + units - 1) / units;
  if (splits > p.total_tiles) splits = p.total_tiles;
  { int per = (p.total_tiles + splits - 1) / splits; splits = (p.total_tiles + per - 1) / per; }
  p.splits = splits;
  { int per = (p.total_tiles + splits - 1) / splits; if (stages > per) stages = per; }
  p.stages = stages;
  uint32_t cols = 32;
  while (cols < (uint32_t)(nv * p.n_tile)) cols <<= 1;
  p.tmem_cols = cols;
  p.out = a->out;
  p.scale = a->scale;

  CUtensorMap tmX, tmY;
  {
    uint64_t dims[4] = {(uint64_t)a->cx, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n_img};
    uint64_t str[3] = {(uint64_t)a->x_pix_stride * 2, (uint64_t)a->x_pix_stride * 2 * a->w, (uint64_t)a->x_pix_stride * 2 * a->w * a->h};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)(p.TH + nv - 1), 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (s2) {   // every second pixel of X: bounding box twice the tile
      box[1] *= 2;
      box[2] *= 2;
      es[1] = es[2] = 2;
    }
    if (!encode_tmap_tiled(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, es)) return SSR_E_CUDA;
  }
  {
    uint64_t dims[4] = {(uint64_t)a->cy, (uint64_t)ow, (uint64_t)oh, (uint64_t)a->n_img};
    uint64_t str[3] = {(uint64_t)a->dy_pix_stride * 2, (uint64_t)a->dy_pix_stride * 2 * ow, (uint64_t)a->dy_pix_stride * 2 * ow * oh};
    uint32_t box[4] = {(uint32_t)(p.y_rowbytes / 2), (uint32_t)p.TW, (uint32_t)p.TH, 1};
    if (!encode_tmap_tiled(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->dy, dims, str, box,
                           p.y_rowbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B)) return SSR_E_CUDA;
  }
  const size_t smem_bytes = (size_t)stages * p.stage_bytes + 1024 + 256;
  dim3 grid((unsigned)splits, (unsigned)(mtiles * nh), (unsigned)n_tiles);
  prof_before(1, stream);
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kWThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t e = p.R == 3 ? cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<3>, tmX, tmY, p)
                    : p.R == 4 ? cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<4>, tmX, tmY, p)
                               : cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<1>, tmX, tmY, p);
    if (!check_cuda(e, "wgrad_tc launch")) return SSR_E_CUDA;
  }
  prof_after(stream);
  count_launch();
  return check_last("wgrad_tc launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_wgrad_unpack(const float* acc, int32_t cx_rows, int32_t acc_stride, float* grad_oihw, int32_t cout, int32_t cin,
                                int32_t r, float scale, int32_t accumulate, void* stream) {
  SSR_REQUIRE(acc && grad_oihw, "ssr_wgrad_unpack: null pointer");
  const long total = (long)cout * cin * r * r;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  wgrad_unpack_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(acc, cx_rows, acc_stride, grad_oihw, cout, cin, r,
                                                                                   scale, accumulate);
  count_launch();
  return check_last("wgrad_unpack launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_wgrad_unpack_batched(const ssr_unpack_desc* descs_device, int32_t n_layers, void* stream) {
  static_assert(sizeof(ssr_unpack_desc) == sizeof(UnpackDesc), "ssr_unpack_desc layout");
  SSR_REQUIRE(descs_device && n_layers > 0, "ssr_wgrad_unpack_batched: bad args");
  // few layers = the discriminator (up to 2.4 M elements per layer): spread each over the whole machine; the generator's 351 small
  // layers (18 k .. 110 k elements) are served by 16 blocks each
  dim3 grid(n_layers <= 32 ? 148u : 16u, (unsigned)n_layers);
  wgrad_unpack_batched_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const UnpackDesc*>(descs_device));
  count_launch();
  return check_last("wgrad_unpack_batched launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_bias_grad(const void* dy_bf16, int32_t dy_pix_stride, int64_t npix, int32_t c, float* out, float scale,
                             void* stream) {
  SSR_REQUIRE(dy_bf16 && out && c > 0, "ssr_bias_grad: bad args");
  if (c % 8 == 0 && c <= 256 && dy_pix_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(dy_bf16) & 15) == 0) {
    const int ppb = 256 / (c / 8);
    long blocks = (npix + ppb * 8L - 1) / (ppb * 8L);   // >= 8 pixels per thread
    if (blocks > 2 * 148) blocks = 2 * 148;
    if (blocks < 1) blocks = 1;
    bias_grad_vec_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(dy_bf16),
                                                                                            dy_pix_stride, npix, c, out, scale);
    count_launch();
    return check_last("bias_grad launch") ? SSR_OK : SSR_E_CUDA;
  }
  dim3 block(32, 8);
  long slabs = (npix + 8 * 64 - 1) / (8 * 64);
  if (slabs > 296) slabs = 296;
  if (slabs < 1) slabs = 1;
  dim3 grid((unsigned)((c + 31) / 32), (unsigned)slabs);
  bias_grad_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(dy_bf16),
                                                                               dy_pix_stride, npix, c, out, scale);
  count_launch();
  return check_last("bias_grad launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_bias_grad_groups(const void* dy_bf16, int32_t dy_pix_stride, int64_t npix, int32_t c, int32_t group_ch,
                                    float* const* outs_device, float scale, void* stream) {
  SSR_REQUIRE(dy_bf16 && outs_device && c > 0 && group_ch > 0, "ssr_bias_grad_groups: bad args");
  dim3 block(32, 8);
  long slabs = (npix + 8 * 64 - 1) / (8 * 64);
  if (slabs > 148) slabs = 148;
  if (slabs < 1) slabs = 1;
  dim3 grid((unsigned)((c + 31) / 32), (unsigned)slabs);
  bias_grad_groups_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy_bf16), dy_pix_stride, npix, c, group_ch, outs_device, scale);
  count_launch();
  return check_last("bias_grad_groups launch") ? SSR_OK : SSR_E_CUDA;
}

extern "C" int ssr_wgrad_tc_batched(const ssr_wgrad_tc_args* args, int32_t n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SSR_REQUIRE(args && n > 0 && n <= 64, "ssr_wgrad_tc_batched: bad args");
  int fallback[64];
  for (int i = 0; i < n; ++i) {
    const ssr_wgrad_tc_args* a = &args[i];
    SSR_REQUIRE(a->x && a->dy && a->out, "ssr_wgrad_tc_batched: null pointer in problem %d", i);
    SSR_REQUIRE(a->x_pix_stride % 8 == 0 && a->dy_pix_stride % 8 == 0, "ssr_wgrad_tc_batched: strides must be multiples of 8");
    SSR_REQUIRE(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->dy)) & 15) == 0, "ssr_wgrad_tc_batched: alignment");
    fallback[i] = 1;
  }
  const int rc = launch_wgrad9_batched(args, n, fallback, stream);
  if (rc != SSR_OK) return rc;
  for (int i = 0; i < n; ++i)
    if (fallback[i]) {
      const int r2 = ssr_wgrad_tc(&args[i], stream_);
      if (r2 != SSR_OK) return r2;
    }
  return SSR_OK;
}
