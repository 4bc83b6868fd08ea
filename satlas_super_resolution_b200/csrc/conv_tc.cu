// Implicit-GEMM R x R (R = 1, 3) stride-1 convolution on tcgen05 tensor cores (sm_100a).
//
// GEMM view per CTA: D[128*MT pixels, n_tile] = sum_{chunk c, kx, ky} A_{c,kx,ky}[pixels, 64] * W_{c,kx,ky}[n_tile, 64]^T
//
//   * one pipeline stage = (64-channel chunk c, horizontal tap kx): ONE 4-D TMA box
//     {64 ch, TW cols, MT*TH + R-1 rows, 1 image} whose x origin is shifted by kx - pad, so the image
//     border is zero-filled by TMA itself (no padding pass, no masks) and the R vertical taps are
//     whole-row offsets (multiples of TW*128 bytes = whole 1024-byte swizzle atoms) into that box;
//   * weights for the stage's R taps arrive by R 2-D TMA boxes {64, n_tile};
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2..5 = epilogue
//     (tcgen05.ld -> bias / LeakyReLU / scaled residuals / activation-derivative mask -> bf16 / f32).
//
// Reference arithmetic replaced: nn.Conv2d(3,1,1) + lrelu + residual-scale of
// ssr/archs/rrdbnet_arch.py:37-44, :63-68, :122-136 and ssr/archs/discriminator_arch.py:44-69.
#include "common.cuh"
#include "ptx.cuh"
#include <stdlib.h>

namespace ssr {

struct ConvTcK {
  int n_img, H, W, R, pad;
  int TW, TH, tiles_x, tiles_y;
  int chunks, cin;
  int n_tile, n_pad, cout;
  int stages;
  uint32_t a_box_bytes, a_alloc, b_bytes, tmem_cols;
  int splits;
  // epilogue
  const float* bias;
  int act;
  float s0, s1, s2;
  const void* res1;
  const void* res2;
  int res1_kind, res2_kind;
  int res1_stride, res2_stride;
  int res1_cmax;  // res1 applies to channels < res1_cmax (0 = all)
  const __nv_bfloat16* mask;
  int mask_stride, mask_lo, mask_relu;
  __nv_bfloat16* out_bf16;
  int out_stride;
  float* out_f32;
  int out32_mode, out32_stride;
  int dbg_aoff;  // experiment: extra row offset (x128 B) of the A descriptor, see scripts/probe_swizzle.py
};

static constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void load16_bf16(const __nv_bfloat16* p, float (&f)[16]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[2 * i] = bf16_lo(u[i]);
    f[2 * i + 1] = bf16_hi(u[i]);
  }
}
__device__ __forceinline__ void load16_f32(const float* p, float (&f)[16]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 t = q[i];
    f[4 * i] = t.x;
    f[4 * i + 1] = t.y;
    f[4 * i + 2] = t.z;
    f[4 * i + 3] = t.w;
  }
}

// Persistent, warp-specialised kernel.  CTA c owns M-tile groups c, c + gridDim.x, ...; its three roles run as
// independent loops coupled only by mbarriers:
//   warp 0     TMA producer   : streams (chunk, kx) stages for tile after tile (runs ahead across tile boundaries)
//   warp 1     MMA issuer     : accumulates a tile into TMEM buffer b = tile & 1, then hands it to the epilogue
//   warps 2..9 epilogue       : drains buffer b (bias / activation / residuals / mask / stores) while the MMAs of the NEXT
//                               tile already fill buffer b ^ 1 -- prologue, first-load latency and epilogue are paid once
//                               per CTA instead of once per tile.
template <int MT, int R>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ConvTcK p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const uint32_t stage_bytes = p.a_alloc + p.b_bytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* bar_empty = bar_full + p.stages;
  uint64_t* bar_acc_full = bar_empty + p.stages;   // [2] accumulator buffer b complete (MMA -> epilogue)
  uint64_t* bar_acc_empty = bar_acc_full + 2;      // [2] accumulator buffer b drained  (epilogue -> MMA), 8 warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.n_tile;

  // split-K range over the 64-channel chunks
  const int per = (p.chunks + p.splits - 1) / p.splits;
  const int c_begin = blockIdx.z * per;
  const int c_end = min(p.chunks, c_begin + per);
  const int iters = (c_end - c_begin) * R;
  const int total_tiles = p.tiles_x * p.tiles_y * p.n_img;
  const int my_tiles = ((int)blockIdx.x < total_tiles) ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (iters <= 0 || my_tiles <= 0) return;
  griddep_launch_dependents();  // let the next kernel's prologue overlap this kernel (it still waits for our completion)
  const uint32_t acc_cols = (uint32_t)(MT * p.n_tile);  // TMEM columns of one accumulator buffer

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmA);
      prefetch_tmap(&tmB);
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&bar_full[s], 1);
        mbar_init(&bar_empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bar_acc_full[b], 1);
        mbar_init(&bar_acc_empty[b], 8);
      }
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
  griddep_wait();  // everything above touched only this CTA's smem / TMEM; global memory of earlier kernels is read below

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    int g = 0;  // running stage counter across tiles
    for (int lt = 0; lt < my_tiles; ++lt) {
      int t = (int)blockIdx.x + lt * (int)gridDim.x;
      const int tx = t % p.tiles_x;
      t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int n = t / p.tiles_y;
      const int x0 = tx * p.TW, y0 = ty * (MT * p.TH);
      for (int it = 0; it < iters; ++it, ++g) {
        const int c = c_begin + it / R;
        const int kx = it - (it / R) * R;
        const int s = g % p.stages;
        const uint32_t ph = (g / p.stages) & 1;
        mbar_wait(&bar_empty[s], ph ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + (size_t)s * stage_bytes;
          uint8_t* b_dst = a_dst + p.a_alloc;
          mbar_expect_tx(&bar_full[s], p.a_box_bytes + p.b_bytes);
          tma_load_4d(a_dst, &tmA, &bar_full[s], c * 64, x0 + kx - p.pad, y0 - p.pad, n);
#pragma unroll
          for (int ky = 0; ky < R; ++ky)
            tma_load_2d(b_dst + (size_t)ky * p.n_tile * 128, &tmB, &bar_full[s], 0, ((c * R + kx) * R + ky) * p.n_pad + n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
    // Descriptors differ only in their 14-bit start-address field: build one per stage, then add constant offsets.
    const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)p.n_tile);
    const uint32_t a_tap = (uint32_t)(p.TW * 128) >> 4;        // one tile row down  (descriptor address units of 16 B)
    const uint32_t a_mt = (uint32_t)(p.TH * p.TW * 128) >> 4;  // next stacked M tile
    const uint32_t b_tap = (uint32_t)(p.n_tile * 128) >> 4;    // next vertical tap's weight tile
    const uint32_t a_dbg = (uint32_t)(p.dbg_aoff * 128) >> 4;  // hardware probe only (scripts/probe_swizzle.py)
    int g = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int b = lt & 1;
      mbar_wait(&bar_acc_empty[b], ((lt >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator buffer
      tc_fence_after_sync();
      const uint32_t d_base = tmem_base + (uint32_t)b * acc_cols;
      uint32_t acc = 0;
      for (int it = 0; it < iters; ++it, ++g) {
        const int c = c_begin + it / R;
        const int s = g % p.stages;
        const uint32_t ph = (g / p.stages) & 1;
        mbar_wait(&bar_full[s], ph);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t a_base = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t da0 = umma_desc_k128(a_base) + a_dbg;
          const uint64_t db0 = umma_desc_k128(a_base + p.a_alloc);
          const int ks = min(4, (p.cin - c * 64) >> 4);
          if (ks == 4) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
              for (int ky = 0; ky < R; ++ky) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_ss(d_base + (uint32_t)(m * p.n_tile), da0 + (m * a_mt + ky * a_tap + 2 * k), db0 + (ky * b_tap + 2 * k),
                               idesc, (ky == 0 && k == 0) ? acc : 1u);
              }
            }
          } else {
#pragma unroll
            for (int m = 0; m < MT; ++m)
              for (int ky = 0; ky < R; ++ky)
                for (int k = 0; k < ks; ++k)
                  umma_bf16_ss(d_base + (uint32_t)(m * p.n_tile), da0 + (m * a_mt + ky * a_tap + 2 * k), db0 + (ky * b_tap + 2 * k),
                               idesc, (ky == 0 && k == 0) ? acc : 1u);
          }
          umma_commit(&bar_empty[s]);                              // frees this smem stage once the MMAs above have read it
          if (it == iters - 1) umma_commit(&bar_acc_full[b]);      // accumulators of this tile complete
        }
        __syncwarp();
        acc = 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9: two warps per TMEM lane quarter) =====================
    // Work items = (M tile, 16-channel chunk); the two warps of a lane quarter take alternate chunks.  Operands that do NOT
    // depend on the accumulator -- bias (staged once in shared memory), residuals, the derivative mask -- are fetched for the
    // first item while the MMAs are still running and for later items before the TMEM load is issued.
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // 0: even chunks, 1: odd chunks
    const int m = q * 32 + lane;
    const int et = (int)threadIdx.x - 64;
    const bool add_bias = (p.bias != nullptr) && (blockIdx.z == 0);
    for (int i = et; i < p.n_tile; i += kThreads - 64) s_bias[i] = (add_bias && n0 + i < p.cout) ? p.bias[n0 + i] : 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int tyy = m / p.TW;
    const int txx = m - tyy * p.TW;
    const int nchunks = p.n_tile >> 4;
    const bool use_r1 = p.res1_kind != SSR_NONE, use_r2 = p.res2_kind != SSR_NONE, use_mk = p.mask != nullptr;

    struct Ops {
      uint4 r1[4], r2[4], mk[2];
    };
    auto fetch = [&](long pix, int c0, Ops& o) {
      if (use_r1 && (p.res1_cmax == 0 || c0 < p.res1_cmax)) {
        if (p.res1_kind == SSR_BF16) {
          const uint4* s4 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res1) + pix * p.res1_stride + c0);
          o.r1[0] = s4[0];
          o.r1[1] = s4[1];
        } else {
          const uint4* s4 = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.res1) + pix * p.res1_stride + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) o.r1[j] = s4[j];
        }
      }
      if (use_r2) {
        if (p.res2_kind == SSR_BF16) {
          const uint4* s4 = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.res2) + pix * p.res2_stride + c0);
          o.r2[0] = s4[0];
          o.r2[1] = s4[1];
        } else {
          const uint4* s4 = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.res2) + pix * p.res2_stride + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) o.r2[j] = s4[j];
        }
      }
      if (use_mk && c0 >= p.mask_lo) {
        const uint4* s4 = reinterpret_cast<const uint4*>(p.mask + pix * p.mask_stride + c0);
        o.mk[0] = s4[0];
        o.mk[1] = s4[1];
      }
    };
    auto expand = [&](const uint4* src, int kind, float (&r)[16]) {
      if (kind == SSR_BF16) {
        const uint32_t u[8] = {src[0].x, src[0].y, src[0].z, src[0].w, src[1].x, src[1].y, src[1].z, src[1].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          r[2 * j] = bf16_lo(u[j]);
          r[2 * j + 1] = bf16_hi(u[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          r[4 * j] = __uint_as_float(src[j].x);
          r[4 * j + 1] = __uint_as_float(src[j].y);
          r[4 * j + 2] = __uint_as_float(src[j].z);
          r[4 * j + 3] = __uint_as_float(src[j].w);
        }
      }
    };

#pragma unroll 1
    for (int lt = 0; lt < my_tiles; ++lt) {
      int t = (int)blockIdx.x + lt * (int)gridDim.x;
      const int tx = t % p.tiles_x;
      t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int n = t / p.tiles_y;
      const int x = tx * p.TW + txx;
      const int y0 = ty * (MT * p.TH);
      const int b = lt & 1;
      const uint32_t d_base = tmem_base + (uint32_t)b * acc_cols;
      int ys[MT];
      long pixs[MT];
      bool oks[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        ys[mt] = y0 + mt * p.TH + tyy;
        pixs[mt] = ((long)n * p.H + ys[mt]) * p.W + x;
        oks[mt] = (tyy < p.TH) && (ys[mt] < p.H) && (x < p.W);
      }
      Ops o;
      bool first = true;
      if (half < nchunks && oks[0] && n0 + half * 16 + 16 <= p.cout) fetch(pixs[0], n0 + half * 16, o);   // overlaps the MMAs
      mbar_wait(&bar_acc_full[b], (lt >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int y = ys[mt];
        const long pix = pixs[mt];
        const bool in_img = oks[mt];
#pragma unroll 1
        for (int ci = half; ci < nchunks; ci += 2) {
          const int c0 = n0 + ci * 16;
          const bool live = in_img && (c0 + 16 <= p.cout);
          if (!first && live) fetch(pix, c0, o);
          first = false;
          uint32_t v[16];
          __syncwarp();
          tmem_ld16(d_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * p.n_tile + ci * 16), v);
          tmem_ld_wait();
          if (!in_img || c0 >= p.cout) continue;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          if (live) {
            if (add_bias) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] += s_bias[ci * 16 + j];
            }
            if (p.act) {
              const float neg = p.act == 2 ? 0.f : 0.2f;  // 1 = LeakyReLU(0.2), 2 = ReLU
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = f[j] > 0.f ? f[j] : neg * f[j];
            }
            if (p.s0 != 1.f) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] *= p.s0;
            }
            if (use_r1 && (p.res1_cmax == 0 || c0 < p.res1_cmax)) {
              float r[16];
              expand(o.r1, p.res1_kind, r);
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaf(p.s1, r[j], f[j]);
            }
            if (use_r2) {
              float r[16];
              expand(o.r2, p.res2_kind, r);
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = fmaf(p.s2, r[j], f[j]);
            }
            // the f32 output is the UNMASKED value (a running gradient sum); the derivative mask only shapes the bf16 copy
            if (p.out32_mode == SSR_OUT32_NHWC) {
              float4* dst = reinterpret_cast<float4*>(p.out_f32 + pix * p.out32_stride + c0);
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else if (p.out32_mode == SSR_OUT32_NHWC_ATOMIC) {
              float* dst = p.out_f32 + pix * p.out32_stride + c0;
#pragma unroll
              for (int j = 0; j < 16; ++j) atomicAdd(dst + j, f[j]);
            } else if (p.out32_mode == SSR_OUT32_NCHW) {
#pragma unroll
              for (int j = 0; j < 16; ++j) p.out_f32[(((long)n * p.cout + c0 + j) * p.H + y) * p.W + x] = f[j];
            }
            if (use_mk && c0 >= p.mask_lo) {
              float r[16];
              expand(o.mk, SSR_BF16, r);
              const float neg = p.mask_relu ? 0.f : 0.2f;
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] *= (r[j] > 0.f ? 1.f : neg);
            }
            if (p.out_bf16 != nullptr) {
              uint4 o0, o1;
              o0.x = pack_bf16(f[0], f[1]);
              o0.y = pack_bf16(f[2], f[3]);
              o0.z = pack_bf16(f[4], f[5]);
              o0.w = pack_bf16(f[6], f[7]);
              o1.x = pack_bf16(f[8], f[9]);
              o1.y = pack_bf16(f[10], f[11]);
              o1.z = pack_bf16(f[12], f[13]);
              o1.w = pack_bf16(f[14], f[15]);
              uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + pix * p.out_stride + c0);
              dst[0] = o0;
              dst[1] = o1;
            }
          } else {
            // ragged tail of the channel dimension (cout not a multiple of 16): scalar path, fully unrolled so that the
            // accumulator array is never indexed dynamically (a dynamic index would force it into local memory)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int c = c0 + j;
              if (c >= p.cout) continue;
              float val = f[j] + s_bias[c - n0];
              if (p.act) val = val > 0.f ? val : (p.act == 2 ? 0.f : 0.2f * val);
              val *= p.s0;
              if (p.res1_cmax == 0 || c < p.res1_cmax) {
                if (p.res1_kind == SSR_BF16)
                  val += p.s1 * __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.res1)[pix * p.res1_stride + c]);
                else if (p.res1_kind == SSR_F32)
                  val += p.s1 * reinterpret_cast<const float*>(p.res1)[pix * p.res1_stride + c];
              }
              if (p.res2_kind == SSR_BF16)
                val += p.s2 * __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.res2)[pix * p.res2_stride + c]);
              else if (p.res2_kind == SSR_F32)
                val += p.s2 * reinterpret_cast<const float*>(p.res2)[pix * p.res2_stride + c];
              if (p.out32_mode == SSR_OUT32_NHWC)
                p.out_f32[pix * p.out32_stride + c] = val;
              else if (p.out32_mode == SSR_OUT32_NHWC_ATOMIC)
                atomicAdd(p.out_f32 + pix * p.out32_stride + c, val);
              else if (p.out32_mode == SSR_OUT32_NCHW)
                p.out_f32[(((long)n * p.cout + c) * p.H + y) * p.W + x] = val;
              if (p.mask != nullptr && c >= p.mask_lo) {
                const float mv = __bfloat162float(p.mask[pix * p.mask_stride + c]);
                val *= (mv > 0.f ? 1.f : (p.mask_relu ? 0.f : 0.2f));
              }
              if (p.out_bf16 != nullptr) p.out_bf16[pix * p.out_stride + c] = __float2bfloat16(val);
            }
          }
        }
      }
      // this warp has finished reading accumulator buffer b: hand it back to the MMA issuer
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_empty[b]);
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

static int g_smem_optin = -1;
static int g_num_sms = 0;

static bool device_limits() {
  if (g_smem_optin >= 0) return true;
  int dev = 0;
  if (!check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return false;
  int v = 0;
  if (!check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev), "smem attr"))
    return false;
  int sms = 0;
  if (!check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "sm count"))
    return false;
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    set_error("libssr_b200 needs an sm_100 GPU (found compute capability major %d)", major);
    return false;
  }
  g_smem_optin = v;
  g_num_sms = sms;
  return true;
}

}  // namespace ssr

using namespace ssr;

// output channels are cut into equal N tiles of at most 128 (192 -> 2 x 96, 160 -> 2 x 80): no half-empty last tile
static int balanced_n_tile(int cout) {
  const int tiles = (cout + 127) / 128;
  return round_up((cout + tiles - 1) / tiles, 16);
}

extern "C" int64_t ssr_packed_weight_bytes(int32_t cin, int32_t cout, int32_t r, int32_t* n_pad) {
  int np = balanced_n_tile(cout) * ((cout + 127) / 128);
  if (n_pad) *n_pad = np;
  int chunks = (cin + 63) / 64;
  return (int64_t)chunks * r * r * np * 64 * 2;
}

extern "C" int ssr_conv_tc(const ssr_conv_tc_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SSR_REQUIRE(a != nullptr, "ssr_conv_tc: null args");
  SSR_REQUIRE(a->r == 1 || a->r == 3, "ssr_conv_tc: r must be 1 or 3 (got %d)", a->r);
  SSR_REQUIRE(a->n_img > 0 && a->h > 0 && a->w > 0, "ssr_conv_tc: bad geometry");
  SSR_REQUIRE(a->cin > 0 && a->cin % 16 == 0, "ssr_conv_tc: cin must be a positive multiple of 16 (got %d)", a->cin);
  SSR_REQUIRE(a->x_pix_stride % 8 == 0 && a->x_pix_stride >= a->cin, "ssr_conv_tc: x_pix_stride %d", a->x_pix_stride);
  SSR_REQUIRE((reinterpret_cast<uintptr_t>(a->x) & 15) == 0, "ssr_conv_tc: x must be 16-byte aligned");
  SSR_REQUIRE((reinterpret_cast<uintptr_t>(a->w_packed) & 127) == 0, "ssr_conv_tc: w_packed must be 128-byte aligned");
  SSR_REQUIRE(a->cout > 0 && a->n_pad >= a->cout && a->n_pad % 16 == 0, "ssr_conv_tc: cout/n_pad");
  SSR_REQUIRE(a->w >= 8, "ssr_conv_tc: width < 8 unsupported");
  if (!device_limits()) return SSR_E_CUDA;

  ConvTcK p{};
  p.n_img = a->n_img;
  p.H = a->h;
  p.W = a->w;
  p.R = a->r;
  p.pad = (a->r - 1) / 2;
  // tile width: narrower tiles have less vertical-halo overhead per pixel (box rows = MT*TH + 2), and every pixel is its own
  // 128-byte TMA segment anyway, so wide images are cut into 32-column tiles (SSR_CONV_TW overrides for experiments)
  {
    static int tw_max = -1;
    if (tw_max < 0) {
      const char* e = getenv("SSR_CONV_TW");
      tw_max = e ? atoi(e) : 32;
      if (tw_max != 16 && tw_max != 32 && tw_max != 64 && tw_max != 128) tw_max = 32;
    }
    p.TW = a->w >= tw_max && a->r == 3 ? tw_max : (a->w >= 128 ? 128 : round_up(a->w, 8));
  }
  p.TH = 128 / p.TW;
  int mt = a->mt;
  if (mt == 0) {
    // two stacked M tiles per CTA when one-tile CTAs would spill past a single co-resident wave: the weight tiles
    // and the halo rows are then shared by 256 pixels (less L2 traffic per MMA) and the grid fits the 148 SMs
    static int forced = -1;
    if (forced < 0) {
      const char* e = getenv("SSR_CONV_MT");
      forced = e ? atoi(e) : 0;
    }
    const int nt_guess = a->n_tile ? a->n_tile : (a->n_pad <= 128 ? a->n_pad : a->n_pad / ((a->n_pad + 127) / 128));
    const long tiles1 = (long)((a->w + p.TW - 1) / p.TW) * ((a->h + p.TH - 1) / p.TH) * a->n_img;
    mt = 1;
    if (forced == 1 || forced == 2) mt = forced;
    else if (a->h >= 2 * p.TH && tiles1 >= 200 && 2 * nt_guess <= 512) mt = 2;
  }
  SSR_REQUIRE(mt == 1 || mt == 2, "ssr_conv_tc: mt must be 1 or 2");
  p.tiles_x = (a->w + p.TW - 1) / p.TW;
  p.tiles_y = (a->h + mt * p.TH - 1) / (mt * p.TH);
  p.chunks = (a->cin + 63) / 64;
  p.cin = a->cin;
  p.n_pad = a->n_pad;
  p.cout = a->cout;
  p.n_tile = a->n_tile ? a->n_tile : (a->n_pad <= 128 ? a->n_pad : a->n_pad / ((a->n_pad + 127) / 128));
  SSR_REQUIRE(p.n_tile % 16 == 0 && p.n_tile <= 256 && p.n_pad % p.n_tile == 0,
              "ssr_conv_tc: n_tile %d incompatible with n_pad %d", p.n_tile, p.n_pad);
  SSR_REQUIRE(2 * mt * p.n_tile <= 512, "ssr_conv_tc: two accumulator buffers of mt*n_tile columns exceed TMEM");
  p.splits = a->splits > 0 ? a->splits : 1;
  if (p.splits > p.chunks) p.splits = p.chunks;
  // make sure no split is empty
  {
    int per = (p.chunks + p.splits - 1) / p.splits;
    p.splits = (p.chunks + per - 1) / per;
  }
  if (p.splits > 1)
    SSR_REQUIRE(a->out32_mode == SSR_OUT32_NHWC_ATOMIC && a->out_bf16 == nullptr && !a->act &&
                    a->res1_kind == SSR_NONE && a->res2_kind == SSR_NONE && a->mask == nullptr,
                "ssr_conv_tc: split-K needs a pure atomic f32 epilogue");

  const int rows = mt * p.TH + p.R - 1;
  p.a_box_bytes = (uint32_t)p.TW * rows * 128u;
  // the M=128 operand window of the last tap may run past the box when TW*TH < 128: keep it inside the stage
  uint32_t need = (uint32_t)(((mt - 1) * p.TH + p.R - 1) * p.TW) * 128u + 16384u;
  p.a_alloc = (uint32_t)round_up((int)max(p.a_box_bytes, need), 1024);
  p.b_bytes = (uint32_t)(p.R * p.n_tile * 128);
  const uint32_t stage_bytes = p.a_alloc + p.b_bytes;
  const int iters_max = ((p.chunks + p.splits - 1) / p.splits) * p.R;
  const int budget = g_smem_optin - 1024 - 256 - 1024;
  int stages = budget / (int)stage_bytes;
  SSR_REQUIRE(stages >= 1, "ssr_conv_tc: stage of %u bytes does not fit shared memory", stage_bytes);
  // prefer two co-resident CTAs per SM when that still leaves a >= 3 deep pipeline
  int stages_half = (budget / 2 - 1024) / (int)stage_bytes;
  if (stages_half >= 3) stages = stages_half;
  if (stages > 8) stages = 8;
  if (stages > iters_max) stages = iters_max;
  p.stages = stages;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * mt * p.n_tile)) cols <<= 1;   // double-buffered accumulators
  p.tmem_cols = cols;

  p.bias = a->bias;
  p.act = a->act;
  p.s0 = a->s0;
  p.s1 = a->s1;
  p.s2 = a->s2;
  p.res1 = a->res1;
  p.res2 = a->res2;
  p.res1_kind = a->res1 ? a->res1_kind : SSR_NONE;
  p.res2_kind = a->res2 ? a->res2_kind : SSR_NONE;
  p.res1_stride = a->res1_pix_stride;
  p.res1_cmax = a->res1_cmax;
  p.res2_stride = a->res2_pix_stride;
  p.mask = reinterpret_cast<const __nv_bfloat16*>(a->mask);
  p.mask_stride = a->mask_pix_stride;
  p.mask_lo = a->mask_lo;
  p.mask_relu = a->mask_relu;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16);
  p.out_stride = a->out_pix_stride;
  p.out_f32 = a->out_f32;
  p.out32_mode = a->out_f32 ? a->out32_mode : SSR_OUT32_NONE;
  p.out32_stride = a->out32_pix_stride;
  {
    const char* e = getenv("SSR_DBG_AOFF");
    p.dbg_aoff = e ? atoi(e) : 0;
  }
  if (a->cout % 16 == 0) {
    // vector epilogue alignment contract
    if (p.out_bf16) SSR_REQUIRE(p.out_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out_bf16) & 15) == 0, "ssr_conv_tc: out_bf16 alignment");
    if (p.res1_kind == SSR_BF16) SSR_REQUIRE(p.res1_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.res1) & 15) == 0, "ssr_conv_tc: res1 alignment");
    if (p.res2_kind == SSR_BF16) SSR_REQUIRE(p.res2_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.res2) & 15) == 0, "ssr_conv_tc: res2 alignment");
    if (p.res1_kind == SSR_F32) SSR_REQUIRE(p.res1_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(p.res1) & 15) == 0, "ssr_conv_tc: res1 alignment");
    if (p.res2_kind == SSR_F32) SSR_REQUIRE(p.res2_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(p.res2) & 15) == 0, "ssr_conv_tc: res2 alignment");
    if (p.mask) SSR_REQUIRE(p.mask_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0, "ssr_conv_tc: mask alignment");
    if (p.out32_mode == SSR_OUT32_NHWC) SSR_REQUIRE(p.out32_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0, "ssr_conv_tc: out_f32 alignment");
    if (p.bias) SSR_REQUIRE((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0, "ssr_conv_tc: bias alignment");
  }

  // tensor maps
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)a->cin, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n_img};
    uint64_t str[3] = {(uint64_t)a->x_pix_stride * 2, (uint64_t)a->x_pix_stride * 2 * a->w,
                       (uint64_t)a->x_pix_stride * 2 * a->w * a->h};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)rows, 1};
    if (!encode_tmap_tiled(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->x, dims, str, box,
                           CU_TENSOR_MAP_SWIZZLE_128B))
      return SSR_E_CUDA;
  }
  {
    uint64_t dims[2] = {64, (uint64_t)p.chunks * p.R * p.R * p.n_pad};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, (uint32_t)p.n_tile};
    if (!encode_tmap_tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->w_packed, dims, str, box,
                           CU_TENSOR_MAP_SWIZZLE_128B))
      return SSR_E_CUDA;
  }

  const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/ + 1024 /*bias*/;
  // persistent: at most one CTA per SM along x, each walking tiles x, x + gridDim.x, ...
  const int total_tiles = p.tiles_x * p.tiles_y * p.n_img;
  int ctas_x = total_tiles < g_num_sms ? total_tiles : g_num_sms;
  {
    // keep the per-CTA tile counts balanced (e.g. 2048 tiles on 148 SMs -> 14 tiles each on 147 CTAs, not 13.8 ragged)
    const int waves = (total_tiles + ctas_x - 1) / ctas_x;
    ctas_x = (total_tiles + waves - 1) / waves;
  }
  dim3 grid((unsigned)ctas_x, (unsigned)(p.n_pad / p.n_tile), (unsigned)p.splits);
  auto kern = mt == 1 ? (p.R == 3 ? conv_tc_kernel<1, 3> : conv_tc_kernel<1, 1>) : (p.R == 3 ? conv_tc_kernel<2, 3> : conv_tc_kernel<2, 1>);
  static size_t configured[6] = {0, 0, 0, 0, 0, 0};
  const int cfg_idx = mt * 2 + (p.R == 3 ? 1 : 0);
  if (configured[cfg_idx] < smem_bytes) {
    if (!check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin),
                    "cudaFuncSetAttribute(conv_tc)"))
      return SSR_E_CUDA;
    configured[cfg_idx] = (size_t)g_smem_optin;
  }
  prof_before(0, stream);
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    if (!check_cuda(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p), "conv_tc launch")) return SSR_E_CUDA;
  }
  prof_after(stream);
  count_launch();
  if (!check_last("conv_tc launch")) return SSR_E_CUDA;
  return SSR_OK;
}

// ------------------------------------------------------------------ weight packing
namespace ssr {
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin, int r, int mode,
                                   const float* __restrict__ inv_scale, __nv_bfloat16* __restrict__ out,
                                   int chunks, int n_pad) {
  // out[c][kx][ky][n][j]
  const long total = (long)chunks * r * r * n_pad * 64;
  const float sc = inv_scale ? 1.f / *inv_scale : 1.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long t = i;
    const int j = t % 64;
    t /= 64;
    const int nn = t % n_pad;
    t /= n_pad;
    const int ky = t % r;
    t /= r;
    const int kx = t % r;
    const int c = t / r;
    const int k = c * 64 + j;
    float v = 0.f;
    if (mode == SSR_PACK_FWD) {
      if (nn < cout && k < cin) v = w[(((long)nn * cin + k) * r + ky) * r + kx];
    } else {
      // dgrad: n indexes the conv's input channel, k its output channel, taps mirrored
      if (nn < cin && k < cout) v = w[(((long)k * cin + nn) * r + (r - 1 - ky)) * r + (r - 1 - kx)];
    }
    out[i] = __float2bfloat16(v * sc);
  }
}
}  // namespace ssr

extern "C" int ssr_pack_conv_weight(const float* w_oihw, int32_t cout, int32_t cin, int32_t r, int32_t mode,
                                    const float* inv_scale, void* packed, int32_t k_pad, int32_t n_pad,
                                    void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SSR_REQUIRE(w_oihw && packed, "ssr_pack_conv_weight: null pointer");
  SSR_REQUIRE(k_pad > 0 && k_pad % 64 == 0 && n_pad % 16 == 0, "ssr_pack_conv_weight: k_pad must be a multiple of 64, n_pad of 16");
  const int red = mode == SSR_PACK_FWD ? cin : cout;
  const int outc = mode == SSR_PACK_FWD ? cout : cin;
  SSR_REQUIRE(k_pad >= red && n_pad >= outc, "ssr_pack_conv_weight: padded sizes too small");
  const int chunks = k_pad / 64;
  const long total = (long)chunks * r * r * n_pad * 64;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4096) blocks = 4096;
  pack_weight_kernel<<<blocks, 256, 0, stream>>>(w_oihw, cout, cin, r, mode, inv_scale,
                                                 reinterpret_cast<__nv_bfloat16*>(packed), chunks, n_pad);
  count_launch();
  if (!check_last("pack_weight launch")) return SSR_E_CUDA;
  return SSR_OK;
}
