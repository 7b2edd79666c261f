// opp_gemm.cuh — the one tensor-core engine of the hot path.
//
// D[M,N] = A[M,K] * W[N,K]^T on tcgen05 (fp16 operands, fp32 accumulators in TMEM), operands
// staged by TMA into 128B-swizzled shared memory through an mbarrier ring, persistent over
// output tiles with a double-buffered TMEM accumulator so the epilogue of tile i overlaps the
// MMAs of tile i+1.
//
// Precision: the reference computes in fp32 and the parity bar is 1e-3 on outputs whose logits
// reach ~1e2, so single fp16 operands (2^-11) are not enough.  In `split` mode every operand is a
// pair of fp16 planes  x = hi + lo  (see opp_common.cuh) and each K-step issues three MMAs into
// the same accumulator:  hi*hi + hi*lo + lo*hi  (the dropped lo*lo term is ~2^-22), i.e. an
// fp32-grade GEMM at 3 tensor-core passes and 2x operand bytes.  split = 0 is plain fp16.
//
// The A operand is either
//   A_ROWS : token rows  [batch][rows][planes*K]  (up to two arrays concatenated along K), or
//   A_CONV : an NHWC feature map read as an implicit-GEMM im2col: for every filter tap the TMA
//            box is the output tile shifted by the tap offset; out-of-bounds coordinates are
//            zero-filled by the TMA unit, which implements the convolution padding for free.
//            Stride-2 convolutions use four parity views (y%2, x%2) of the same tensor.
// Everything after the accumulator (bias / BN / activation / residual / LayerNorm / elu+1 /
// linear-attention normaliser / dual-softmax statistics) is a fused epilogue functor.
//
// Warp roles (192 threads): warp 0 = TMA producer (lane 0), warp 1 = TMEM owner + MMA issuer
// (lane 0), warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31, one row per thread).
//
// Reference semantics implemented by the epilogues are cited at each functor
// (paths relative to the reference repo zju3dv/OnePose_Plus_Plus).
#pragma once

#include "opp_common.cuh"

namespace opp {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;               // 64 fp16 = 128 B = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kMaxStages = 8;
constexpr int kEpiSmemBytes = 8192;
constexpr int kGemmThreads = 192;

enum AMode : int { A_ROWS = 0, A_CONV = 1 };

struct TensorMaps {
  CUtensorMap a[4];
  CUtensorMap b;
};

struct GemmShape {
  int batches;      // tiles never straddle a batch
  int rows;         // A_ROWS: valid rows per batch; A_CONV: out_h*out_w
  int m_tiles;      // M tiles per batch
  int n_tiles;
  int n_total;      // valid output columns
  int block_n;      // UMMA N (multiple of 16, <= 256)
  int k_chunks;     // number of 64-wide K chunks per tile (per plane)
  int stages;
  int b_batched;    // W operand has a leading batch dim
  int split;        // operands are (hi|lo) plane pairs; 3 MMAs per K-step
  int b_lo;         // element offset of W's lo plane inside a row (= K total)
  // A_ROWS
  int k_chunks_a0;  // chunks read through maps.a[0]; the rest through maps.a[1]
  int a0_lo, a1_lo; // element offsets of the lo planes of the two A arrays (= their K)
  // A_CONV
  int conv_cchunks; // K chunks per filter tap
  int conv_c;       // padded input channels (multiple of 16); also the lo-plane offset
  int conv_kw;      // filter width/height (1 or 3)
  int conv_pad;
  int conv_stride;  // 1 or 2
  int tile_w, tile_h;    // output tile, tile_w*tile_h == 128
  int tiles_x, tiles_y;  // tiles per image
  int out_w, out_h;
};

struct EpiCtx {
  uint32_t tmem;   // accumulator address of this thread's lane quarter, column 0 of the tile
  int b, m_tile, n_tile;
  int row;         // row within the batch (pixel index within the image for A_CONV)
  long long grow;  // b*rows + row
  bool valid;      // row is inside the tensor
  int n0;          // first global column of the tile
  int ncols;       // valid columns in this tile (multiple of 8)
  int etid;        // 0..127 within the epilogue group
  float* smem;     // kEpiSmemBytes of scratch shared by the epilogue group
};

__device__ __forceinline__ void epi_sync() { named_bar_sync(1, 128); }

// =============================================================================================
// Epilogues.  Outputs that feed later GEMMs are written as (hi|lo) plane pairs when
// `out_lo` != 0: row layout [hi(n_total) | lo(n_total)], out_lo = n_total.
// =============================================================================================

// Plain store with an optional activation on the leading `act_cols` columns.
//   act 1 = ReLU  (transformer.py:41-45 mlp ReLU), act 2 = elu(x)+1 (linear_attention.py:10-11)
struct EpiStoreF16 {
  struct Params {
    __half* out;
    long long ld;   // row stride in elements
    int out_lo;     // 0 or n_total
    int act;
    int act_cols;
  };
  __device__ static void run(const Params& p, const GemmShape& s, const EpiCtx& c) {
    for (int col = 0; col < c.ncols; col += 32) {
      float v[32];
      tmem_ld32(c.tmem + col, v);
      const int g0 = c.n0 + col;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (g0 + j < p.act_cols) {
          if (p.act == 1) v[j] = fmaxf(v[j], 0.f);
          else if (p.act == 2) v[j] = elu_plus_one(v[j]);
        }
      }
      if (c.valid) {
        __half* row = p.out + c.grow * p.ld;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (col + g * 8 < c.ncols) store_split8(row, g0 + g * 8, v + g * 8, p.out_lo);
      }
    }
  }
};

// Query side of linear attention (linear_attention.py:45,58-59): Q = elu(q)+1,
// Z = 1/(Q . Ksum + eps), output Q * Z * v_length per head of 32 channels.  The matching KV
// state is pre-divided by v_length (linear_attention.py:55-56), so (Q*Z*v_length) @ (KV/v_length)
// reproduces the reference product.
struct EpiQ {
  struct Params {
    __half* out;
    long long ld;
    int out_lo;
    const float* ksum;  // [batches][n_total]
    float v_len;
    float eps;
  };
  __device__ static void run(const Params& p, const GemmShape& s, const EpiCtx& c) {
    epi_sync();
    for (int i = c.etid; i < c.ncols; i += 128)
      c.smem[i] = p.ksum[(long long)c.b * s.n_total + c.n0 + i];
    epi_sync();
    for (int col = 0; col < c.ncols; col += 32) {
      float v[32];
      tmem_ld32(c.tmem + col, v);
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = elu_plus_one(v[j]);
        dot = fmaf(v[j], c.smem[col + j], dot);
      }
      const float z = p.v_len / (dot + p.eps);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= z;
      if (c.valid) {
        __half* row = p.out + c.grow * p.ld;
#pragma unroll
        for (int g = 0; g < 4; ++g) store_split8(row, c.n0 + col + g * 8, v + g * 8, p.out_lo);
      }
    }
  }
};

// LayerNorm over the full output row (the tile spans all N columns), optional residual add
// (transformer.py:86-94: norm1 after merge; norm2 then x + msg).
struct EpiLN {
  struct Params {
    const float* gamma;
    const float* beta;
    float eps;
    const __half* resid;  // same layout as out16 (ld, out_lo) or null
    __half* out16;        // or null
    long long ld;
    int out_lo;
    float* out32;         // fp32 [rows][n_total] or null
  };
  __device__ static void run(const Params& p, const GemmShape& s, const EpiCtx& c) {
    float* g_s = c.smem;
    float* b_s = c.smem + 256;
    epi_sync();
    for (int i = c.etid; i < c.ncols; i += 128) {
      g_s[i] = p.gamma[i];
      b_s[i] = p.beta[i];
    }
    epi_sync();
    const float inv_n = 1.f / (float)c.ncols;
    float sum = 0.f;
    for (int col = 0; col < c.ncols; col += 32) {
      float v[32];
      tmem_ld32(c.tmem + col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += v[j];
    }
    const float mean = sum * inv_n;
    float sq = 0.f;
    for (int col = 0; col < c.ncols; col += 32) {
      float v[32];
      tmem_ld32(c.tmem + col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float d = v[j] - mean;
        sq = fmaf(d, d, sq);
      }
    }
    const float rstd = 1.f / sqrtf(sq * inv_n + p.eps);
    for (int col = 0; col < c.ncols; col += 32) {
      float v[32];
      tmem_ld32(c.tmem + col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * g_s[col + j] + b_s[col + j];
      if (!c.valid) continue;
      if (p.resid) {
        const __half* rrow = p.resid + c.grow * p.ld;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float r[8];
          load_split8(rrow, col + g * 8, r, p.out_lo);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[g * 8 + j] += r[j];
        }
      }
      if (p.out32) {
        float4* o4 = reinterpret_cast<float4*>(p.out32 + c.grow * (long long)s.n_total + col);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          o4[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
      }
      if (p.out16) {
        __half* row = p.out16 + c.grow * p.ld;
#pragma unroll
        for (int g = 0; g < 4; ++g) store_split8(row, col + g * 8, v + g * 8, p.out_lo);
      }
    }
  }
};

// Convolution epilogue: folded-BN bias, residual add, ReLU / LeakyReLU (resnet.py:36-45,112-124,
// 141-147).  Optionally also emits the coarse tokens  x3_out + pe  in token-major order
// (position_encoding.py:37-42 + OnePosePlusModel.py:137-142: NHWC *is* 'n (h w) c').
struct EpiConv {
  struct Params {
    __half* out;          // NHWC, pixel stride ld (or null)
    long long ld;
    int out_lo;
    const float* bias;    // [n_total]
    const __half* resid;  // same layout as out, or null
    int act;              // 0 none, 1 relu, 2 leaky relu
    float slope;
    __half* tok;          // [B*H*W] rows with the same (ld, out_lo) layout, or null
    const float* pe;      // [H*W][n_total]
  };
  __device__ static void run(const Params& p, const GemmShape& s, const EpiCtx& c) {
    epi_sync();
    for (int i = c.etid; i < c.ncols; i += 128) c.smem[i] = p.bias[c.n0 + i];
    epi_sync();
    for (int col = 0; col < c.ncols; col += 32) {
      float v[32];
      tmem_ld32(c.tmem + col, v);
      if (!c.valid) continue;
      const int g0 = c.n0 + col;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (col + g * 8 >= c.ncols) break;
        float* vv = v + g * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) vv[j] += c.smem[col + g * 8 + j];
        if (p.resid) {
          float r[8];
          load_split8(p.resid + c.grow * p.ld, g0 + g * 8, r, p.out_lo);
#pragma unroll
          for (int j = 0; j < 8; ++j) vv[j] += r[j];
        }
        if (p.act == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) vv[j] = fmaxf(vv[j], 0.f);
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) vv[j] = vv[j] > 0.f ? vv[j] : vv[j] * p.slope;
        }
        if (p.out) store_split8(p.out + c.grow * p.ld, g0 + g * 8, vv, p.out_lo);
        if (p.tok) {
          const float* pe = p.pe + (long long)c.row * s.n_total + g0 + g * 8;
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) t[j] = vv[j] + pe[j];
          store_split8(p.tok + c.grow * p.ld, g0 + g * 8, t, p.out_lo);
        }
      }
    }
  }
};

// Dual-softmax statistics (coarse_matching.py:102-115): per row, over this tile's columns,
// (max, sum exp) of sim = acc*scale.  Partials [grow][n_tile] are merged by a finalize kernel.
struct EpiLse {
  struct Params {
    float* part_m;
    float* part_s;
    float scale;
  };
  __device__ static void run(const Params& p, const GemmShape& s, const EpiCtx& c) {
    float m = -INFINITY;
    for (int col = 0; col < c.ncols; col += 32) {
      float v[32];
      tmem_ld32(c.tmem + col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col + j < c.ncols) m = fmaxf(m, v[j] * p.scale);
    }
    float sum = 0.f;
    for (int col = 0; col < c.ncols; col += 32) {
      float v[32];
      tmem_ld32(c.tmem + col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col + j < c.ncols) sum += expf(v[j] * p.scale - m);
    }
    if (c.valid) {
      p.part_m[c.grow * s.n_tiles + c.n_tile] = m;
      p.part_s[c.grow * s.n_tiles + c.n_tile] = sum;
    }
  }
};

// conf = softmax_dim1(sim) * softmax_dim2(sim) = exp((2*sim - lse_pt) - lse_px)
// (coarse_matching.py:115) evaluated per element, optional fp32 store of conf_matrix, and the
// per-row (max, first argmax) over this tile's columns for the mutual-nearest test
// (coarse_matching.py:157-165).  `own_is_pt` says whether rows are 3D points (pass A) or
// query cells (pass B); the expression is evaluated in the same order in both passes.
struct EpiConf {
  struct Params {
    const float* lse_own;    // [batches*rows]
    const float* lse_other;  // [batches][n_total]
    float scale;
    int own_is_pt;
    float* conf;             // [batches*rows][n_total] or null
    float* part_val;         // [batches*rows][n_tiles]
    int* part_idx;
  };
  __device__ static void run(const Params& p, const GemmShape& s, const EpiCtx& c) {
    epi_sync();
    for (int i = c.etid; i < c.ncols; i += 128)
      c.smem[i] = p.lse_other[(long long)c.b * s.n_total + c.n0 + i];
    epi_sync();
    const float lown = c.valid ? p.lse_own[c.grow] : 0.f;
    float best = -1.f;
    int best_idx = c.n0;
    for (int col = 0; col < c.ncols; col += 32) {
      float v[32];
      tmem_ld32(c.tmem + col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x2 = 2.f * (v[j] * p.scale);
        const float lo = c.smem[(col + j) & 255];
        const float e = p.own_is_pt ? (x2 - lown) - lo : (x2 - lo) - lown;
        v[j] = expf(e);
        if (col + j < c.ncols && v[j] > best) {
          best = v[j];
          best_idx = c.n0 + col + j;
        }
      }
      if (c.valid && p.conf) {
        float* dst = p.conf + c.grow * (long long)s.n_total + c.n0 + col;
        if ((s.n_total & 3) == 0) {
          float4* o4 = reinterpret_cast<float4*>(dst);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            if (col + g * 4 < c.ncols)
              o4[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col + j < c.ncols) dst[j] = v[j];
        }
      }
    }
    if (c.valid) {
      p.part_val[c.grow * s.n_tiles + c.n_tile] = best;
      p.part_idx[c.grow * s.n_tiles + c.n_tile] = best_idx;
    }
  }
};

// =============================================================================================
// The kernel
// =============================================================================================
template <int A_MODE, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ TensorMaps maps, const GemmShape s,
            const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int planes = s.split ? 2 : 1;
  const int b_bytes = s.block_n * kBlockK * 2;   // one plane of the W tile
  const int a_stage = kABytes * planes;
  const int b_stage = b_bytes * planes;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + s.stages * a_stage;
  float* epi_smem = reinterpret_cast<float*>(smem_b + s.stages * b_stage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(epi_smem) +
                                               kEpiSmemBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int acc_stride = s.block_n <= 128 ? 128 : 256;
  const uint32_t tmem_cols = 2 * acc_stride;
  const int tiles_per_batch = s.m_tiles * s.n_tiles;
  const int total_tiles = s.batches * tiles_per_batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
  }
  if (warp == 2 && lane == 0) {
    for (int i = 0; i < s.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int b = t / tiles_per_batch;
        const int r = t - b * tiles_per_batch;
        const int m_tile = r / s.n_tiles;
        const int n_tile = r - m_tile * s.n_tiles;
        int ox0 = 0, oy0 = 0;
        if (A_MODE == A_CONV) {
          const int ty = m_tile / s.tiles_x;
          oy0 = ty * s.tile_h;
          ox0 = (m_tile - ty * s.tiles_x) * s.tile_w;
        }
        for (int chunk = 0; chunk < s.k_chunks; ++chunk) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], a_stage + b_stage);
          uint8_t* sa = smem_a + stage * a_stage;
          uint8_t* sb = smem_b + stage * b_stage;
          int kb;
          if (A_MODE == A_ROWS) {
            const bool first = chunk < s.k_chunks_a0;
            const int kc = (first ? chunk : chunk - s.k_chunks_a0) * kBlockK;
            const CUtensorMap* am = first ? &maps.a[0] : &maps.a[1];
            tma_load_3d(am, &full[stage], sa, kc, m_tile * kBlockM, b);
            if (s.split)
              tma_load_3d(am, &full[stage], sa + kABytes, kc + (first ? s.a0_lo : s.a1_lo),
                          m_tile * kBlockM, b);
            kb = chunk * kBlockK;
          } else {
            const int tap = chunk / s.conv_cchunks;
            const int cc = chunk - tap * s.conv_cchunks;
            const int ky = tap / s.conv_kw;
            const int kx = tap - ky * s.conv_kw;
            int dy = ky - s.conv_pad, dx = kx - s.conv_pad, mi = 0;
            if (s.conv_stride == 2) {
              const int py = dy & 1, px = dx & 1;
              dy = (dy - py) >> 1;
              dx = (dx - px) >> 1;
              mi = py * 2 + px;
            }
            tma_load_4d(&maps.a[mi], &full[stage], sa, cc * kBlockK, ox0 + dx, oy0 + dy, b);
            if (s.split)
              tma_load_4d(&maps.a[mi], &full[stage], sa + kABytes, s.conv_c + cc * kBlockK,
                          ox0 + dx, oy0 + dy, b);
            kb = tap * s.conv_c + cc * kBlockK;
          }
          const int bb = s.b_batched ? b : 0;
          tma_load_3d(&maps.b, &full[stage], sb, kb, n_tile * s.block_n, bb);
          if (s.split)
            tma_load_3d(&maps.b, &full[stage], sb + b_bytes, s.b_lo + kb, n_tile * s.block_n, bb);
          if (++stage == s.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(kBlockM, s.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        for (int chunk = 0; chunk < s.k_chunks; ++chunk) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          int ksteps = 4;
          if (A_MODE == A_CONV) {
            const int cc = chunk % s.conv_cchunks;
            const int rem = (s.conv_c - cc * kBlockK) >> 4;
            ksteps = rem < 4 ? rem : 4;
          }
          const uint32_t sa = smem_u32(smem_a + stage * a_stage);
          const uint32_t sb = smem_u32(smem_b + stage * b_stage);
          const uint64_t a_hi = make_kmajor_sw128_desc(sa);
          const uint64_t b_hi = make_kmajor_sw128_desc(sb);
          // advance 16 fp16 = 32 B along K inside the 128 B swizzle row: +2 in (addr >> 4)
          for (int k = 0; k < ksteps; ++k)
            tc_mma_f16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, idesc, (chunk | k) != 0);
          if (s.split) {
            const uint64_t a_lo = make_kmajor_sw128_desc(sa + kABytes);
            const uint64_t b_lo = make_kmajor_sw128_desc(sb + b_bytes);
            for (int k = 0; k < ksteps; ++k) tc_mma_f16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, idesc, 1);
            for (int k = 0; k < ksteps; ++k) tc_mma_f16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, idesc, 1);
          }
          tc_commit(&empty[stage]);
          if (++stage == s.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&tfull[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    EpiCtx c;
    c.etid = (warp - 2) * 32 + lane;
    c.smem = epi_smem;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      c.b = t / tiles_per_batch;
      const int r = t - c.b * tiles_per_batch;
      c.m_tile = r / s.n_tiles;
      c.n_tile = r - c.m_tile * s.n_tiles;
      c.n0 = c.n_tile * s.block_n;
      const int rem = s.n_total - c.n0;
      c.ncols = rem < s.block_n ? rem : s.block_n;
      if (A_MODE == A_ROWS) {
        c.row = c.m_tile * kBlockM + row_in_tile;
        c.valid = c.row < s.rows;
      } else {
        const int ty = c.m_tile / s.tiles_x;
        const int tx = c.m_tile - ty * s.tiles_x;
        const int ly = row_in_tile / s.tile_w;
        const int oy = ty * s.tile_h + ly;
        const int ox = tx * s.tile_w + (row_in_tile - ly * s.tile_w);
        c.valid = oy < s.out_h && ox < s.out_w;
        c.row = oy * s.out_w + ox;
      }
      c.grow = (long long)c.b * s.rows + c.row;
      c.tmem = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * acc_stride;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      Epi::run(ep, s, c);
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// dynamic shared memory a launch needs (ring + epilogue scratch + barriers + alignment slack)
inline int gemm_stage_bytes(int block_n, int split) {
  return (kABytes + block_n * kBlockK * 2) * (split ? 2 : 1);
}
inline int gemm_smem_bytes(int stages, int block_n, int split) {
  return stages * gemm_stage_bytes(block_n, split) + kEpiSmemBytes + (2 * kMaxStages + 4) * 8 + 16 +
         1024;
}
inline int gemm_pick_stages(int block_n, int k_chunks, int split) {
  int st = (227 * 1024 - kEpiSmemBytes - 2048) / gemm_stage_bytes(block_n, split);
  if (st > kMaxStages) st = kMaxStages;
  if (st > k_chunks * 2 && k_chunks * 2 >= 2) st = k_chunks * 2;
  return st < 2 ? 2 : st;
}

}  // namespace opp
