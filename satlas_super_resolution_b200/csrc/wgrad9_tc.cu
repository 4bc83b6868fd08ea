// Weight gradient of a 3x3 conv with ALL NINE taps per CTA (tcgen05, sm_100a) -- the fast path of ssr_wgrad_tc for cy <= 64.
//
//   dW[ky][kx][cx][cy] += scale * sum_p X[p + (ky-1, kx-1), cx] * dY[p, cy]
//
// One TMA box {64 ch, TW+2 cols, TH+2 rows} per 64-channel chunk holds the tile WITH its halo (borders zero-filled by TMA).
// Because the 128B swizzle of both TMA and the UMMA descriptors is a function of the absolute shared-memory address
// (probed on the B200, scripts/probe_swizzle.py), an operand window may start at any 128-byte row: tap (ky, kx) of K-step s
// (16 consecutive pixels of image row ry) is just the start address ((ry+ky)*(TW+2) + 16*hx + kx) * 128 -- so the activation
// tile is loaded ONCE for nine taps (the first version reloaded it per kx from three CTAs).  Operands are MN-major (one
// 128-byte row per pixel = per K index); nine f32 accumulators [128 cx, 32 cy] live in 288 TMEM columns; a CTA walks a range
// of pixel tiles and reduces its partial sums with vector red.global.add.v4.f32.
//
// fuse3 (default): the shift is moved from X to dY -- dW[ky][kx] = sum_q X[q] * dY[q - (ky-1, kx-1)] over the pixels q of the
// tile, with the dY tile carrying the halo -- so that the three horizontal taps of one ky share ONE MMA: their dY windows are
// the same 16 pixel rows shifted by one pixel (64 bytes), i.e. three N blocks of an MN-major operand with LBO = 64 B.  One
// 128 x 96 x 16 MMA (4 KB of X + 3 KB of dY from shared memory) replaces three 128 x 32 x 16 MMAs (3 x 5 KB): the operand
// reads that bound this kernel drop 2.1x.  Three accumulators [128 cx, 3 x 32 cy] in 288 TMEM columns.
#include "common.cuh"
#include "ptx.cuh"

namespace ssr {

struct Wg9K {
  int n_img, H, W, TW, TH, tiles_x, tiles_y, total_tiles, kpr;
  int cx, cx_rows, cy, out_stride;
  int stages, splits;
  int fuse3;   // dY carries the halo, the three kx taps of a ky are one N = 96 MMA
  uint32_t x_chunk_bytes, x_chunk_alloc, y_bytes, stage_bytes;
  float* out;
  float scale;
};

static constexpr int kW9Threads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue


__device__ __forceinline__ void wgrad9_body(const CUtensorMap& tmX, const CUtensorMap& tmY, const Wg9K& p, int bx, int by, int bz, int splits) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* bar_empty = bar_full + p.stages;
  uint64_t* bar_tmem = bar_empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tmem + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int mtile = by;
  const int n0 = bz * 32;
  const int per = (p.total_tiles + splits - 1) / splits;
  const int t_begin = bx * per;
  const int t_end = min(p.total_tiles, t_begin + per);
  const int iters = t_end - t_begin;
  if (iters <= 0) return;
  griddep_launch_dependents();
  const int nchunks = min(2, (p.cx - mtile * 128 + 63) / 64);
  const uint32_t x_bytes = 2 * p.x_chunk_alloc;
  constexpr uint32_t kCols = 512;

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
    tmem_alloc(tmem_slot, kCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();

  if (warp == 0) {
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
        mbar_expect_tx(&bar_full[s], nchunks * p.x_chunk_bytes + p.y_bytes);
        const int hx = p.fuse3 ? 0 : 1, hy = p.fuse3 ? 1 : 0;   // which operand carries the halo
        for (int ch = 0; ch < nchunks; ++ch)
          tma_load_4d(xs + (size_t)ch * p.x_chunk_alloc, &tmX, &bar_full[s], mtile * 128 + ch * 64, x0 - hx, y0 - hx, n);
        tma_load_4d(xs + x_bytes, &tmY, &bar_full[s], n0, x0 - hy, y0 - hy, n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128u, p.fuse3 ? 96u : 32u, 1u, 1u);
    const int pitch = p.TW + 2;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
      const int s = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      mbar_wait(&bar_full[s], ph);
      tc_fence_after_sync();
      if (elect_one()) {
        const uint32_t xs = smem_u32(smem + (size_t)s * p.stage_bytes);
        const uint64_t da0 = umma_desc(xs, p.x_chunk_alloc, 1024u, 2u);               // X: 128-byte rows, SWIZZLE_128B
        if (p.fuse3) {
          // dY with halo, 64-byte rows, SWIZZLE_64B; N blocks (kx = 2, 1, 0) are one pixel row = 64 B apart
          const uint64_t db0 = umma_desc(xs + x_bytes, 64u, 512u, 4u);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const int ry = ks / p.kpr, hx = ks - ry * p.kpr;
            const uint64_t da = da0 + (uint32_t)(ry * p.TW + hx * 16) * 8u;            // 16 interior pixels of X
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const uint64_t db = db0 + (uint32_t)((ry + 2 - ky) * pitch + hx * 16) * 4u;   // 64 B >> 4 per pixel row
              umma_bf16_ss(tmem_base + (uint32_t)(ky * 96), da, db, idesc, ks == 0 ? acc : 1u);
            }
          }
        } else {
          const uint64_t db0 = umma_desc(xs + x_bytes, 0u, 512u, 4u);                  // dY: 64-byte rows, SWIZZLE_64B
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const int ry = ks / p.kpr, hx = ks - ry * p.kpr;
            const uint32_t row0 = (uint32_t)(ry * pitch + hx * 16);
            const uint64_t db = db0 + (uint32_t)(ks * 64);                             // 16 pixels * 64 B >> 4
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const uint64_t da = da0 + (row0 + (uint32_t)(ky * pitch + kx)) * 8u;   // 128 B >> 4 per pixel row
                umma_bf16_ss(tmem_base + (uint32_t)((ky * 3 + kx) * 32), da, db, idesc, ks == 0 ? acc : 1u);
              }
            }
          }
        }
        umma_commit(&bar_empty[s]);
        if (it == iters - 1) umma_commit(bar_tmem);
      }
      __syncwarp();
      acc = 1;
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int cxi = mtile * 128 + row;
    const bool valid = cxi < p.cx;
    mbar_wait(bar_tmem, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int item = half; item < 18; item += 2) {
      const int tap = item >> 1, cb = (item & 1) * 16;
      // fuse3: accumulator ky holds the N blocks in the order kx = 2, 1, 0
      const int col = p.fuse3 ? (tap / 3) * 96 + (2 - tap % 3) * 32 + cb : tap * 32 + cb;
      uint32_t v[16];
      __syncwarp();
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
      tmem_ld_wait();
      if (!valid) continue;
      const int c0 = n0 + cb;
      if (c0 >= p.cy) continue;
      // accumulator layout [cy / 4][9 taps * cx_rows][4]: the warp's 32 consecutive cx rows are 512 contiguous bytes per
      // channel quad (channels past cy inside the last quad receive exact zeros: their dY columns are TMA zero fill)
      const long plane = 9L * p.cx_rows * 4;
      float* dst = p.out + (long)(c0 >> 2) * plane + ((long)tap * p.cx_rows + cxi) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + 4 * j < p.out_stride)
          red_add_v4(dst + j * plane, p.scale * __uint_as_float(v[4 * j]), p.scale * __uint_as_float(v[4 * j + 1]),
                     p.scale * __uint_as_float(v[4 * j + 2]), p.scale * __uint_as_float(v[4 * j + 3]));
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kCols);
  }
}

__global__ void __launch_bounds__(kW9Threads, 1)
wgrad9_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const Wg9K p) {
  wgrad9_body(tmX, tmY, p, blockIdx.x, blockIdx.y, blockIdx.z, p.splits);
}

// Horizontal fusion: up to kW9Batch independent weight-gradient problems (the five convs of a ResidualDenseBlock) in ONE launch
// of about one CTA per SM.  A problem splits into units (128 cx rows x 32 cy columns of output); every unit gets a share of
// the CTAs proportional to its work (pixel tiles; the MMA operand reads dominate), so all CTAs finish together and each pays
// prologue, first-load latency and the reduction epilogue once for a long run of pixel tiles.
static constexpr int kW9Batch = 8;
static constexpr int kW9Units = 32;
struct Wg9Unit {
  short prob, by, bz, pad_;
  int splits, cta_begin;
};
struct Wg9BatchK {
  CUtensorMap tmX[kW9Batch];
  CUtensorMap tmY[kW9Batch];
  Wg9K k[kW9Batch];
  Wg9Unit u[kW9Units];
  int n_units;
};

__global__ void __launch_bounds__(kW9Threads, 1) wgrad9_tc_batched_kernel(const __grid_constant__ Wg9BatchK b) {
  int j = 0;
  while (j + 1 < b.n_units && (int)blockIdx.x >= b.u[j + 1].cta_begin) ++j;
  const Wg9Unit& u = b.u[j];
  wgrad9_body(b.tmX[u.prob], b.tmY[u.prob], b.k[u.prob], (int)blockIdx.x - u.cta_begin, u.by, u.bz, u.splits);
}

static int g_w9_smem = -1;

struct Wg9Prepared {
  Wg9K p;
  CUtensorMap tmX, tmY;
  int mtiles, halves;
  size_t smem_bytes;
};

// returns SSR_OK / error, or 1 when the shape is not eligible (caller falls back to the per-kx kernel)
static int prepare_wgrad9(const ssr_wgrad_tc_args* a, Wg9Prepared* out, int target_override = 0) {
  if (a->r != 3 || a->cy > 64) return 1;
  int TW = a->w >= 128 ? 128 : a->w;
  if (TW != 16 && TW != 32 && TW != 64 && TW != 128) return 1;
  if (a->w % TW) return 1;
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("SSR_WGRAD9");
    disabled = (e && e[0] == '0') ? 1 : 0;
  }
  if (disabled) return 1;
  if (g_w9_smem < 0) {
    int dev = 0, v = 0;
    if (!check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return SSR_E_CUDA;
    if (!check_cuda(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev), "smem attr")) return SSR_E_CUDA;
    if (!check_cuda(cudaFuncSetAttribute(wgrad9_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v), "cudaFuncSetAttribute(wgrad9)"))
      return SSR_E_CUDA;
    if (!check_cuda(cudaFuncSetAttribute(wgrad9_tc_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v), "cudaFuncSetAttribute(wgrad9b)"))
      return SSR_E_CUDA;
    g_w9_smem = v;
  }
  Wg9K p{};
  p.n_img = a->n_img; p.H = a->h; p.W = a->w;
  p.TW = TW; p.TH = 128 / TW; p.kpr = TW / 16;
  p.tiles_x = a->w / TW;
  p.tiles_y = (a->h + p.TH - 1) / p.TH;
  p.total_tiles = p.tiles_x * p.tiles_y * a->n_img;
  p.cx = a->cx; p.cx_rows = a->out_cx_rows; p.cy = a->cy; p.out_stride = a->out_stride;
  SSR_REQUIRE(p.cx_rows >= p.cx && p.out_stride >= p.cy && p.out_stride % 4 == 0, "ssr_wgrad_tc: output layout");
  SSR_REQUIRE((reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "ssr_wgrad_tc: out must be 16-byte aligned");
  static int fuse3 = -1;
  if (fuse3 < 0) {
    const char* e = getenv("SSR_WGRAD9_FUSE3");
    fuse3 = e ? atoi(e) : 1;
  }
  p.fuse3 = fuse3;
  if (fuse3) {
    // X: the 128 interior pixels; dY: the tile with its halo (operand windows end exactly at its last row)
    p.x_chunk_bytes = (uint32_t)(TW * p.TH) * 128u;
    p.x_chunk_alloc = p.x_chunk_bytes;
    p.y_bytes = (uint32_t)((TW + 2) * (p.TH + 2)) * 64u;
  } else {
    p.x_chunk_bytes = (uint32_t)((TW + 2) * (p.TH + 2)) * 128u;
    // the last tap of the last K-step reads up to ((TH+1)*(TW+2) + TW + 2) rows: keep it inside the chunk allocation
    p.x_chunk_alloc = (uint32_t)round_up((int)(((p.TH + 2) * (TW + 2) + 16) * 128), 1024);
    p.y_bytes = (uint32_t)(TW * p.TH * 64);
  }
  p.stage_bytes = (uint32_t)round_up((int)(2 * p.x_chunk_alloc + p.y_bytes), 1024);
  int stages = (g_w9_smem - 1280) / (int)p.stage_bytes;
  if (stages > 4) stages = 4;
  SSR_REQUIRE(stages >= 1, "ssr_wgrad_tc: stage does not fit shared memory");
  const int mtiles = (a->cx + 127) / 128;
  const int halves = (a->cy + 31) / 32;
  const int units = mtiles * halves;
  static int target_ctas = -1;
  if (target_ctas < 0) {
    const char* e = getenv("SSR_WGRAD_CTAS");
    target_ctas = e ? atoi(e) : 148;
  }
  const int tgt = target_override > 0 ? target_override : target_ctas;
  int splits = a->splits > 0 ? a->splits : (tgt + units - 1) / units;
  if (splits < 1) splits = 1;
  if (splits > p.total_tiles) splits = p.total_tiles;
  { int per = (p.total_tiles + splits - 1) / splits; splits = (p.total_tiles + per - 1) / per; if (stages > per) stages = per; }
  p.splits = splits;
  p.stages = stages;
  p.out = a->out;
  p.scale = a->scale;
  CUtensorMap tmX, tmY;
  {
    uint64_t dims[4] = {(uint64_t)a->cx, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n_img};
    uint64_t str[3] = {(uint64_t)a->x_pix_stride * 2, (uint64_t)a->x_pix_stride * 2 * a->w, (uint64_t)a->x_pix_stride * 2 * a->w * a->h};
    uint32_t box[4] = {64, (uint32_t)(TW + (fuse3 ? 0 : 2)), (uint32_t)(p.TH + (fuse3 ? 0 : 2)), 1};
    if (!encode_tmap_tiled(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return SSR_E_CUDA;
  }
  {
    uint64_t dims[4] = {(uint64_t)a->cy, (uint64_t)a->w, (uint64_t)a->h, (uint64_t)a->n_img};
    uint64_t str[3] = {(uint64_t)a->dy_pix_stride * 2, (uint64_t)a->dy_pix_stride * 2 * a->w, (uint64_t)a->dy_pix_stride * 2 * a->w * a->h};
    uint32_t box[4] = {32, (uint32_t)(TW + (fuse3 ? 2 : 0)), (uint32_t)(p.TH + (fuse3 ? 2 : 0)), 1};
    if (!encode_tmap_tiled(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->dy, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return SSR_E_CUDA;
  }
  out->p = p;
  out->tmX = tmX;
  out->tmY = tmY;
  out->mtiles = mtiles;
  out->halves = halves;
  out->smem_bytes = (size_t)stages * p.stage_bytes + 1024 + 256;
  return SSR_OK;
}

static cudaLaunchAttribute pdl_attr() {
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  return attr;
}

int launch_wgrad9(const ssr_wgrad_tc_args* a, cudaStream_t stream) {
  Wg9Prepared w;
  const int rc = prepare_wgrad9(a, &w);
  if (rc != SSR_OK) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)w.p.splits, (unsigned)w.mtiles, (unsigned)w.halves);
  cfg.blockDim = dim3(kW9Threads);
  cfg.dynamicSmemBytes = w.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1] = {pdl_attr()};
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  prof_before(1, stream);
  if (!check_cuda(cudaLaunchKernelEx(&cfg, wgrad9_tc_kernel, w.tmX, w.tmY, w.p), "wgrad9_tc launch")) return SSR_E_CUDA;
  prof_after(stream);
  count_launch();
  return SSR_OK;
}

// n problems in as few launches as possible; problems that are not eligible for the nine-tap kernel return 1 in `fallback[i]`
int launch_wgrad9_batched(const ssr_wgrad_tc_args* args, int n, int* fallback, cudaStream_t stream) {
  static int budget = -1;
  if (budget < 0) {
    const char* e = getenv("SSR_WGRAD_BATCH_CTAS");
    budget = e ? atoi(e) : 0;
    if (budget <= 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&budget, cudaDevAttrMultiProcessorCount, dev);   // one wave: one CTA per SM
    }
  }
  int i = 0;
  while (i < n) {
    static Wg9BatchK b;
    b.n_units = 0;
    int n_prob = 0;
    size_t smem = 0;
    double cost[kW9Units];
    double total_cost = 0;
    while (i < n && n_prob < kW9Batch) {
      Wg9Prepared w;
      const int rc = prepare_wgrad9(&args[i], &w, 1);
      if (rc < 0) return rc;
      fallback[i] = rc == 1 ? 1 : 0;
      if (rc == SSR_OK) {
        if (b.n_units + w.mtiles * w.halves > kW9Units) break;   // next launch takes it
        const int j = n_prob++;
        b.k[j] = w.p;
        b.tmX[j] = w.tmX;
        b.tmY[j] = w.tmY;
        for (int bz = 0; bz < w.halves; ++bz)
          for (int by = 0; by < w.mtiles; ++by) {
            Wg9Unit& u = b.u[b.n_units];
            u.prob = (short)j;
            u.by = (short)by;
            u.bz = (short)bz;
            // a pixel tile costs 72 MMAs whose shared-memory operand reads (4 KB of X per MMA) bound the SM whatever the
            // number of valid cx rows; units that stream one 64-channel chunk instead of two are a little cheaper
            const int ch = w.p.cx - by * 128 < 128 ? w.p.cx - by * 128 : 128;   // valid X channels of this unit
            cost[b.n_units] = (double)w.p.total_tiles * (ch > 64 ? 1.0 : 0.85);
            total_cost += cost[b.n_units];
            ++b.n_units;
          }
        if (w.smem_bytes > smem) smem = w.smem_bytes;
      }
      ++i;
    }
    if (b.n_units == 0) continue;
    // CTAs per unit, proportional to cost; leftovers go to the units with the most work per CTA
    int used = 0;
    for (int u = 0; u < b.n_units; ++u) {
      int sp = (int)(budget * cost[u] / total_cost);
      const int tiles = b.k[b.u[u].prob].total_tiles;
      if (sp < 1) sp = 1;
      if (sp > tiles) sp = tiles;
      b.u[u].splits = sp;
      used += sp;
    }
    while (used < budget) {
      int best = -1;
      double best_load = 0;
      for (int u = 0; u < b.n_units; ++u) {
        const double load = cost[u] / b.u[u].splits;
        if (b.u[u].splits < b.k[b.u[u].prob].total_tiles && load > best_load) {
          best_load = load;
          best = u;
        }
      }
      if (best < 0) break;
      ++b.u[best].splits;
      ++used;
    }
    int ctas = 0;
    for (int u = 0; u < b.n_units; ++u) {
      // no empty split: per = ceil(tiles / splits), splits = ceil(tiles / per)
      const int tiles = b.k[b.u[u].prob].total_tiles;
      const int per = (tiles + b.u[u].splits - 1) / b.u[u].splits;
      b.u[u].splits = (tiles + per - 1) / per;
      b.u[u].cta_begin = ctas;
      ctas += b.u[u].splits;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kW9Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1] = {pdl_attr()};
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    prof_before(4, stream);   // profile class 4: the batched dense-block weight gradient
    if (!check_cuda(cudaLaunchKernelEx(&cfg, wgrad9_tc_batched_kernel, b), "wgrad9_tc_batched launch")) return SSR_E_CUDA;
    prof_after(stream);
    count_launch();
  }
  return SSR_OK;
}

}  // namespace ssr
