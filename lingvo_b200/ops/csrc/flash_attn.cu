// tcgen05 flash attention with a learned Toeplitz (T5) relative-position bias
// (reference semantics `gshard_builder.py:2697-2804,1431-1514`).
//
//   O[b,i,h,:] = softmax_j( scale·q_i·k_j + rel[h, i-j+L-1] + mask(b,i,j) ) · v_j
//
// mask(b,i,j) is built in registers from packed-input metadata: keys are visible iff they
// are in the query's (non-zero) segment and, for decoders, not in its future
// (segment_pos[j] <= segment_pos[i]).  Neither the [B,H,L,L] bias, nor S, P, dP, dS ever
// reach HBM; the gradient of the bias table is folded into the backward kernel.
//
// Both kernels use the same recipe: operand tiles [128 x 128] bf16 arrive by TMA as two
// [128 rows x 64 cols] SWIZZLE_128B chunks; one elected thread issues tcgen05.mma
// (128x128x16, fp32 accumulators in TMEM); eight epilogue warps (one query row per TMEM
// lane, two column halves) do the softmax math between the GEMMs. The same physical smem
// layout serves as a K-major operand (rows = M/N index) or an MN-major operand (rows = K
// index) by changing only the descriptor, so no transposes are ever materialised.
//
//   forward  (CTA = (i-block, h, b)):  loop over key blocks:
//       S = Q·K^T -> TMEM | online softmax, P (bf16) -> smem | PV = P·V -> TMEM |
//       O = alpha·O + PV in registers.  K/V double-buffered; S(j+1) is issued right
//       behind PV(j) so the tensor pipe works while the warps rescale O.
//   backward (CTA = (j-block, h, b)):  loop over query blocks (FlashAttention-2 order):
//       S = Q·K^T, dP = dO·V^T | P = exp2(S - lse), dS = P∘(dP - delta) |
//       dV += P^T·dO | dK += dS^T·Q | dQ_tile = dS·K -> fp32 red.add into dq_acc;
//       d rel[h, i-j+L-1] += Σ dS along diagonals (warp-shuffle skew, no smem atomics).
//       dK, dV stay in TMEM for the whole loop (S | dP | dV | dK = 512 columns).
//
// Layouts: q,k,v,dO [B, L, H, D=128] bf16 views (strides (*, *, D, 1));  O, dK, dV
// [B, L, H, D] contiguous bf16;  dq_acc [B, L, H, D] fp32;  lse, delta [B, H, L] fp32;
// rel / drel [H, 2L-1] fp32;  segment ids / positions [B, L] int32 (optional).

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <torch/extension.h>

#include "ptx.cuh"
#include "registry.h"

namespace lb {

CUtensorMap MakeMap(const void* base, int64_t inner, int64_t rows, int64_t groups,
                    int64_t row_stride, int64_t group_stride, int box_rows,
                    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, int elem_bytes = 2);

namespace {

constexpr int kT = 128;                      // tile edge (queries / keys)
constexpr int kD = 128;                      // head dim
constexpr int kChunk = kT * 64 * 2;          // [128 x 64] SW128 chunk = 16 KiB
constexpr int kTile = 2 * kChunk;            // [128 x 128] bf16 operand = 32 KiB
constexpr int kThreads = 320;                // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kMasked = -1.0e9f;           // additive mask value (log2 domain, finite)

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}
// Explicit shared-space accesses: the tile base is re-aligned through integer arithmetic,
// after which the compiler can only emit *generic* LD/ST for C++ pointers into it.
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s16(uint32_t a) {
  int v;
  asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_s32(uint32_t a, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_s16(uint32_t a, int v) {
  asm volatile("st.shared.s16 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// smem byte offset of the 16-byte unit holding columns [8u, 8u+8) of row r in a
// [128 x 128] tile stored as two SW128 chunks of 64 columns.
__device__ __forceinline__ uint32_t tile_unit_off(int r, int col8) {
  const int chunk = col8 >> 3;               // 64 columns = 8 units per chunk
  const int u = col8 & 7;
  return static_cast<uint32_t>(chunk * kChunk + r * 128 + ((u ^ (r & 7)) << 4));
}

// Issue one 128x128x128 GEMM as 8 UMMAs. kAMn / kBMn: operand is MN-major (rows = K index).
template <bool kAMn, bool kBMn>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t sa, uint32_t sb,
                                           bool accumulate) {
  constexpr uint32_t idesc = make_idesc(1, 1, kAMn, kBMn, kT, kT);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    // K-major: 64-wide K chunks are kChunk apart, 16-element steps are 32 B inside the
    //          128 B swizzle row.  MN-major: 16 K-rows of 128 B per step; the two 64-wide
    //          MN chunks are kChunk apart (LBO), 8-row groups 1 KiB (SBO).
    const uint64_t adesc = kAMn ? make_smem_desc_sw128(sa + k * (16 * 128), kChunk, 1024)
                                : make_smem_desc_sw128(sa + (k >> 2) * kChunk + (k & 3) * 32, 16, 1024);
    const uint64_t bdesc = kBMn ? make_smem_desc_sw128(sb + k * (16 * 128), kChunk, 1024)
                                : make_smem_desc_sw128(sb + (k >> 2) * kChunk + (k & 3) * 32, 16, 1024);
    umma_f16(tmem_d, adesc, bdesc, idesc, (accumulate || k != 0) ? 1u : 0u);
  }
}

__device__ __forceinline__ void load_tile(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                          int h, int row0, int b) {
  tma_load_3d(dst, map, bar, h * kD, row0, b);
  tma_load_3d(dst + kChunk, map, bar, h * kD + 64, row0, b);
}

struct MaskInfo {
  const int* seg;        // [B, L] or nullptr (single segment, nothing padded)
  const int* pos;        // [B, L] or nullptr (position = index)
  const int4* blk;       // [B, L/128] (pos_min, pos_max, seg_min, seg_max) or nullptr
  int causal;
};

// Per-128-row summary of the packed-input metadata; lets every CTA skip (query block, key
// block) pairs in which nothing can be visible — the causal upper triangle, but also
// blocks that only hold other segments — without assuming anything about the positions.
__device__ __forceinline__ int4 block_meta(const MaskInfo& m, int b, int nblk, int blk) {
  if (m.blk != nullptr) return m.blk[b * nblk + blk];
  return make_int4(blk * kT, blk * kT + kT - 1, 1, 1);
}
// Every (query, key) pair of the two blocks is visible: no per-element mask needed.
__device__ __forceinline__ bool pair_full(const MaskInfo& m, const int4& qm, const int4& km) {
  return qm.z == qm.w && km.z == km.w && qm.z == km.z && qm.z != 0 &&
         (!m.causal || qm.x >= km.y);
}
__device__ __forceinline__ bool pair_visible(const MaskInfo& m, const int4& qm, const int4& km) {
  if (qm.w == 0 || km.w == 0) return false;                  // a block of pure padding
  if (km.w < qm.z || km.z > qm.w) return false;              // disjoint segment-id ranges
  if (m.causal && qm.y < km.x) return false;                 // every key lies in the future
  return true;
}

// ================================================================== forward ==
struct FwdParams {
  const float* rel;      // [H, 2L-1] or nullptr
  MaskInfo mask;
  __nv_bfloat16* out;    // [B, L, H, D]
  float* lse;            // [B, H, L] natural log
  int B, H, L;
  float scale;
};

// smem: Q | K x2 | V | P x2 (each 32 KiB) + 8 KiB of barriers / per-block tables.
constexpr size_t kFwdSmem = kTile * 6 + 8192 + 1024;

// Software pipeline (steady state, block j):
//   tensor pipe : ... PV(j-1) | S(j+1) ...          (issued while the warps work on block j)
//   warps       : softmax(j) -> P(j) to smem -> signal -> O = alpha·O + PV(j-1)
// S and PV are double-buffered in TMEM (4 x 128 columns), P in smem, K has two stages and
// V one (V(j+1) is only needed after softmax(j+1)); the per-block bias/mask tables for
// block j+1 are fetched while block j is processed. The warps never wait for a GEMM that
// was not issued a whole softmax earlier.
__global__ void __launch_bounds__(kThreads, 1)
flash_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const FwdParams p) {
  const int nblk = p.L / kT;
  const int i_blk = nblk - 1 - static_cast<int>(blockIdx.x);      // heavy (long) rows first
  const int h = blockIdx.y, b = blockIdx.z;
  const int i0 = i_blk * kT;
  const int4 qmeta = block_meta(p.mask, b, nblk, i_blk);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t s_q = smem_u32(smem);
  const uint32_t s_k = s_q + kTile;                          // 2 stages
  const uint32_t s_v = s_k + 2 * kTile;                      // 1 stage
  const uint32_t s_p = s_v + kTile;                          // 2 buffers
  uint8_t* meta = smem + kTile * 6;
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta);
  uint64_t* q_bar = bars;                 // 1
  uint64_t* k_full = bars + 1;            // 2
  uint64_t* k_empty = bars + 3;           // 2
  uint64_t* v_full = bars + 5;            // 1
  uint64_t* v_empty = bars + 6;           // 1
  uint64_t* s_full = bars + 7;            // 2
  uint64_t* p_ready = bars + 9;           // 2
  uint64_t* pv_full = bars + 11;          // 2
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 13);
  const uint32_t s_xch = smem_u32(meta + 128);              // float [2][128]
  const uint32_t s_tab = s_xch + 1024;                      // 2 x {rel[256] f32, seg[128], pos[128]}
  constexpr int kTableBytes = 1024 + 512 + 512;

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(smem_u32(q_bar), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&k_full[s]), 1);
      mbar_init(smem_u32(&k_empty[s]), 1);
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&p_ready[s]), 8);
      mbar_init(smem_u32(&pv_full[s]), 1);
    }
    mbar_init(smem_u32(v_full), 1);
    mbar_init(smem_u32(v_empty), 1);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc<512>(smem_u32(tmem_ptr_smem));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // S buffers: columns [0,128), [128,256);  PV buffers: [256,384), [384,512)

  // ordered list of visible key blocks is implicit: every role walks jb = 0..nblk-1 and
  // skips with the same predicate.
  int n_vis = 0;
  for (int jb = 0; jb < nblk; ++jb)
    n_vis += pair_visible(p.mask, qmeta, block_meta(p.mask, b, nblk, jb)) ? 1 : 0;

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(q_bar), kTile);
      load_tile(s_q, &map_q, smem_u32(q_bar), h, i0, b);
      // K runs one block ahead of V: K(0), then for each visible block it: K(it+1), V(it).
      int jk = 0;                                  // scan position of the K stream
      int itk = 0;
      auto next_k = [&]() {
        while (jk < nblk && !pair_visible(p.mask, qmeta, block_meta(p.mask, b, nblk, jk))) ++jk;
        if (jk >= nblk) return;
        const int s = itk & 1;
        mbar_wait(smem_u32(&k_empty[s]), ((itk >> 1) & 1) ^ 1);
        const uint32_t fb = smem_u32(&k_full[s]);
        mbar_arrive_expect_tx(fb, kTile);
        load_tile(s_k + s * kTile, &map_k, fb, h, jk * kT, b);
        ++itk;
        ++jk;
      };
      next_k();
      int it = 0;
      for (int jb = 0; jb < nblk; ++jb) {
        if (!pair_visible(p.mask, qmeta, block_meta(p.mask, b, nblk, jb))) continue;
        next_k();
        mbar_wait(smem_u32(v_empty), (it & 1) ^ 1);
        mbar_arrive_expect_tx(smem_u32(v_full), kTile);
        load_tile(s_v, &map_v, smem_u32(v_full), h, jb * kT, b);
        ++it;
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ============================= MMA issuer =============================
    if (lane == 0) {
      mbar_wait(smem_u32(q_bar), 0);        // always: the Q load must land before the CTA exits
      auto issue_s = [&](int it) {          // S(it) into S buffer it & 1 from K stage it & 1
        const int s = it & 1;
        mbar_wait(smem_u32(&k_full[s]), (it >> 1) & 1);
        tc_fence_after();
        issue_gemm<false, false>(tmem_base + s * 128, s_q, s_k + s * kTile, false);
        umma_commit(smem_u32(&s_full[s]));
        umma_commit(smem_u32(&k_empty[s]));
      };
      if (n_vis > 0) issue_s(0);
      if (n_vis > 1) issue_s(1);
      for (int it = 0; it < n_vis; ++it) {
        const int s = it & 1;
        mbar_wait(smem_u32(&p_ready[s]), (it >> 1) & 1);
        mbar_wait(smem_u32(v_full), it & 1);
        tc_fence_after();
        issue_gemm<false, true>(tmem_base + 256 + s * 128, s_p + s * kTile, s_v, false);   // P·V
        umma_commit(smem_u32(&pv_full[s]));
        umma_commit(smem_u32(v_empty));
        if (it + 2 < n_vis) issue_s(it + 2);      // S buffer `s` was drained before p_ready(it)
      }
    }
    __syncwarp();
  } else {
    // ================= softmax / rescale warps (8 x 32 threads) =================
    const int q4 = warp_idx & 3;                       // TMEM lane quarter of this warp
    const int hf = (warp_idx - 2) >> 2;                // column half
    const int r = q4 * 32 + lane;                      // query row inside the tile
    const int et = (hf * 4 + ((q4 + 2) & 3)) * 32 + lane;   // 0..255, dense id for loads
    const int i = i0 + r;
    const long long bl = static_cast<long long>(b) * p.L;
    const int segq = p.mask.seg ? p.mask.seg[bl + i] : 1;
    const int posq = p.mask.pos ? p.mask.pos[bl + i] : i;
    const float sl2 = p.scale * kLog2e;
    const uint32_t lane_off = static_cast<uint32_t>(q4 * 32) << 16;
    float o[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) o[c] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    // per-block tables → smem buffer `buf` (published by the next epi_bar_sync)
    auto fetch_tables = [&](int jb, int buf) {
      const uint32_t tb = s_tab + buf * kTableBytes;
      const int j0 = jb * kT;
      if (et < 255)
        sts_f32(tb + et * 4,
                p.rel ? p.rel[static_cast<long long>(h) * (2 * p.L - 1) + (i0 - j0 + p.L - 1) - 127 + et] * kLog2e
                      : 0.f);
      if (et < 128) {
        sts_s32(tb + 1024 + et * 4, p.mask.seg ? p.mask.seg[bl + j0 + et] : 1);
        sts_s32(tb + 1536 + et * 4, p.mask.pos ? p.mask.pos[bl + j0 + et] : j0 + et);
      }
    };
    auto next_vis = [&](int jb) {
      while (jb < nblk && !pair_visible(p.mask, qmeta, block_meta(p.mask, b, nblk, jb))) ++jb;
      return jb;
    };
    // O = alpha·O + PV(buffer)
    auto o_update = [&](int it, float alpha) {
      const int s = it & 1;
      mbar_wait(smem_u32(&pv_full[s]), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + 256 + s * 128 + lane_off + hf * 64 + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k)
          o[cc * 32 + k] = fmaf(o[cc * 32 + k], alpha, __uint_as_float(v[k]));
      }
      tc_fence_before();
    };

    int jb = next_vis(0);
    if (jb < nblk) fetch_tables(jb, 0);
    float alpha_prev = 0.f;
    for (int it = 0; it < n_vis; ++it) {
      const int s = it & 1;
      const uint32_t rel_a = s_tab + s * kTableBytes + (r + 127) * 4;   // rel_s[r - c + 127]
      const uint32_t seg_a = s_tab + s * kTableBytes + 1024;
      const uint32_t pos_a = seg_a + 512;
      const bool full = pair_full(p.mask, qmeta, block_meta(p.mask, b, nblk, jb));
      const int jb_next = next_vis(jb + 1);
      epi_bar_sync();                                  // tables of block `it` are in place
      if (jb_next < nblk) fetch_tables(jb_next, s ^ 1);    // buffer s^1: readers left it at the barrier
      mbar_wait(smem_u32(&s_full[s]), (it >> 1) & 1);
      tc_fence_after();
      float sv[64];
      float mloc = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + s * 128 + lane_off + hf * 64 + cc * 32, v);
        tmem_ld_wait();
        if (full) {                    // interior block: bias only, no mask arithmetic
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int c = hf * 64 + cc * 32 + k;
            const float x = fmaf(__uint_as_float(v[k]), sl2, lds_f32(rel_a - c * 4));
            sv[cc * 32 + k] = x;
            mloc = fmaxf(mloc, x);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const int c = hf * 64 + cc * 32 + k;
            float x = fmaf(__uint_as_float(v[k]), sl2, lds_f32(rel_a - c * 4));
            const bool vis = (segq == lds_s32(seg_a + c * 4)) && (segq != 0) &&
                             (!p.mask.causal || posq >= lds_s32(pos_a + c * 4));
            x = vis ? x : kMasked;
            sv[cc * 32 + k] = x;
            mloc = fmaxf(mloc, x);
          }
        }
      }
      sts_f32(s_xch + (hf * 128 + r) * 4, mloc);
      epi_bar_sync();
      const float m_new = fmaxf(m_run, fmaxf(mloc, lds_f32(s_xch + ((1 - hf) * 128 + r) * 4)));
      const float alpha = ex2(m_run - m_new);
      float lsum = 0.f;
      const uint32_t s_pb = s_p + s * kTile;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = ex2(sv[u * 8 + 2 * e] - m_new);
          const float p1 = ex2(sv[u * 8 + 2 * e + 1] - m_new);
          lsum += p0 + p1;
          w[e] = pack_bf16x2(p0, p1);
        }
        st_shared_v4(s_pb + tile_unit_off(r, hf * 8 + u), w[0], w[1], w[2], w[3]);
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      fence_proxy_async();          // P (generic-proxy stores) → visible to the MMA's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_ready[s]));
      // the previous block's P·V has been running during this softmax: fold it in now
      if (it > 0) o_update(it - 1, alpha_prev);
      alpha_prev = alpha;
      jb = jb_next;
    }
    if (n_vis > 0) o_update(n_vis - 1, alpha_prev);
    // ---- finalize: O / l, log-sum-exp ----
    epi_bar_sync();
    sts_f32(s_xch + (hf * 128 + r) * 4, l_run);
    epi_bar_sync();
    const float l_tot = l_run + lds_f32(s_xch + ((1 - hf) * 128 + r) * 4);
    const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;      // query block that sees nothing: 0
    __nv_bfloat16* dst = p.out + ((bl + i) * p.H + h) * kD + hf * 64;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      int4 w;
      w.x = pack_bf16x2(o[u * 8 + 0] * inv, o[u * 8 + 1] * inv);
      w.y = pack_bf16x2(o[u * 8 + 2] * inv, o[u * 8 + 3] * inv);
      w.z = pack_bf16x2(o[u * 8 + 4] * inv, o[u * 8 + 5] * inv);
      w.w = pack_bf16x2(o[u * 8 + 6] * inv, o[u * 8 + 7] * inv);
      st_v4(dst + u * 8, w);
    }
    if (hf == 0)
      p.lse[(static_cast<long long>(b) * p.H + h) * p.L + i] =
          l_tot > 0.f ? (m_run + lg2(l_tot)) * kLn2 : 0.f;
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================= backward ==
struct BwdParams {
  const float* rel;      // [H, 2L-1] or nullptr
  MaskInfo mask;
  const float* lse;      // [B, H, L]
  const float* delta;    // [B, H, L]
  float* dq_acc;         // [B, L, H, D] fp32, zero-initialised
  __nv_bfloat16* dk;     // [B, L, H, D]
  __nv_bfloat16* dv;     // [B, L, H, D]
  float* drel;           // [H, 2L-1] fp32, zero-initialised (or nullptr)
  int B, H, L;
  float scale;
};

constexpr int kBwdStages = 2;
// K, V resident + 2 x (Q, dO) + one P/dS tile + 2 KiB of barriers / tables.
constexpr size_t kBwdSmem = kTile * (2 + 2 * kBwdStages + 1) + 2048 + 1008;

__global__ void __launch_bounds__(kThreads, 1)
flash_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                 const BwdParams p) {
  const int nblk = p.L / kT;
  const int j_blk = blockIdx.x;                               // key block (small j = long loop first)
  const int h = blockIdx.y, b = blockIdx.z;
  const int j0 = j_blk * kT;
  const int4 kmeta = block_meta(p.mask, b, nblk, j_blk);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t s_k = smem_u32(smem);
  const uint32_t s_v = s_k + kTile;
  const uint32_t s_qdo = s_v + kTile;                        // stage s: Q at +2s·kTile, dO behind it
  const uint32_t s_pds = s_qdo + 2 * kBwdStages * kTile;
  uint8_t* meta = smem + kTile * (3 + 2 * kBwdStages);
  uint64_t* bars = reinterpret_cast<uint64_t*>(meta);
  uint64_t* kv_bar = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = bars + 3;
  uint64_t* sdp_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* p_consumed = bars + 7;
  uint64_t* ds_ready = bars + 8;
  uint64_t* dq_full = bars + 9;
  uint64_t* dq_read = bars + 10;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 11);
  const uint32_t s_rel = smem_u32(meta + 128);                // float [256]
  const uint32_t s_seg = s_rel + 1024;                        // short [128]
  const uint32_t s_pos = s_seg + 256;                         // short [128]

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    tma_prefetch_desc(&map_do);
    mbar_init(smem_u32(kv_bar), 1);
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(sdp_full), 1);
    mbar_init(smem_u32(p_ready), 8);
    mbar_init(smem_u32(p_consumed), 1);
    mbar_init(smem_u32(ds_ready), 8);
    mbar_init(smem_u32(dq_full), 1);
    mbar_init(smem_u32(dq_read), 8);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc<512>(smem_u32(tmem_ptr_smem));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t t_s = tmem_base;            // S, later dQ tile
  const uint32_t t_dp = tmem_base + 128;
  const uint32_t t_dv = tmem_base + 256;
  const uint32_t t_dk = tmem_base + 384;

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(smem_u32(kv_bar), 2 * kTile);
      load_tile(s_k, &map_k, smem_u32(kv_bar), h, j0, b);
      load_tile(s_v, &map_v, smem_u32(kv_bar), h, j0, b);
      int it = 0;
      for (int ib = 0; ib < nblk; ++ib) {
        if (!pair_visible(p.mask, block_meta(p.mask, b, nblk, ib), kmeta)) continue;
        const int s = it % kBwdStages;
        mbar_wait(smem_u32(&empty_bar[s]), ((it / kBwdStages) & 1) ^ 1);
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_arrive_expect_tx(fb, 2 * kTile);
        load_tile(s_qdo + (2 * s) * kTile, &map_q, fb, h, ib * kT, b);
        load_tile(s_qdo + (2 * s + 1) * kTile, &map_do, fb, h, ib * kT, b);
        ++it;
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ============================= MMA issuer =============================
    if (lane == 0) {
      int n_q = 0;
      for (int ib = 0; ib < nblk; ++ib)
        n_q += pair_visible(p.mask, block_meta(p.mask, b, nblk, ib), kmeta) ? 1 : 0;
      mbar_wait(smem_u32(kv_bar), 0);
      for (int it = 0; it < n_q; ++it) {
        const int s = it % kBwdStages;
        const uint32_t sq = s_qdo + (2 * s) * kTile, sdo = sq + kTile;
        mbar_wait(smem_u32(&full_bar[s]), (it / kBwdStages) & 1);
        if (it > 0) mbar_wait(smem_u32(dq_read), (it - 1) & 1);     // dQ tile left columns [0,128)
        tc_fence_after();
        issue_gemm<false, false>(t_s, sq, s_k, false);              // S  = Q·K^T
        issue_gemm<false, false>(t_dp, sdo, s_v, false);            // dP = dO·V^T
        umma_commit(smem_u32(sdp_full));
        mbar_wait(smem_u32(p_ready), it & 1);
        tc_fence_after();
        issue_gemm<true, true>(t_dv, s_pds, sdo, it > 0);           // dV += P^T·dO
        umma_commit(smem_u32(p_consumed));
        mbar_wait(smem_u32(ds_ready), it & 1);
        tc_fence_after();
        issue_gemm<true, true>(t_dk, s_pds, sq, it > 0);            // dK += dS^T·Q
        issue_gemm<false, true>(t_s, s_pds, s_k, false);            // dQ tile = dS·K
        umma_commit(smem_u32(dq_full));
        umma_commit(smem_u32(&empty_bar[s]));
      }
    }
    __syncwarp();
  } else {
    // ============== softmax-recompute / dS / dQ warps (8 x 32 threads) ==============
    const int q4 = warp_idx & 3;
    const int hf = (warp_idx - 2) >> 2;
    const int r = q4 * 32 + lane;                      // query row (S, dP, dQ) / key row (dK, dV)
    const int et = (hf * 4 + ((q4 + 2) & 3)) * 32 + lane;
    const long long bl = static_cast<long long>(b) * p.L;
    const long long bh = static_cast<long long>(b) * p.H + h;
    const float sl2 = p.scale * kLog2e;
    const uint32_t lane_off = static_cast<uint32_t>(q4 * 32) << 16;
    if (et < 128) {
      sts_s16(s_seg + et * 2, p.mask.seg ? p.mask.seg[bl + j0 + et] : 1);
      sts_s16(s_pos + et * 2, p.mask.pos ? p.mask.pos[bl + j0 + et] : j0 + et);
    }

    int it = 0;
    for (int ib = 0; ib < nblk; ++ib) {
      if (!pair_visible(p.mask, block_meta(p.mask, b, nblk, ib), kmeta)) continue;
      const int i0 = ib * kT;
      const int i = i0 + r;
      if (et < 255)
        sts_f32(s_rel + et * 4,
                p.rel ? p.rel[static_cast<long long>(h) * (2 * p.L - 1) + (i0 - j0 + p.L - 1) - 127 + et] * kLog2e
                      : 0.f);
      const uint32_t rel_a = s_rel + (r + 127) * 4;
      const int segq = p.mask.seg ? p.mask.seg[bl + i] : 1;
      const int posq = p.mask.pos ? p.mask.pos[bl + i] : i;
      const float lse2 = p.lse[bh * p.L + i] * kLog2e;
      const float delta = p.delta[bh * p.L + i];
      epi_bar_sync();
      mbar_wait(smem_u32(sdp_full), it & 1);
      tc_fence_after();
      const bool full = pair_full(p.mask, block_meta(p.mask, b, nblk, ib), kmeta);
      uint32_t dsp[32];                                 // scale·dS, packed bf16x2 (64 values)
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;               // diagonal sums: lane, lane-32, lane-64
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t vs[32], vp[32];
        tmem_ld_32x32b_x32(t_s + lane_off + hf * 64 + cc * 32, vs);
        tmem_ld_32x32b_x32(t_dp + lane_off + hf * 64 + cc * 32, vp);
        tmem_ld_wait();
        float ds[32];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float pr[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int k = u * 8 + 2 * e + t;
              const int c = hf * 64 + cc * 32 + k;
              const float x = fmaf(__uint_as_float(vs[k]), sl2, lds_f32(rel_a - c * 4));
              const bool vis = full || ((segq == lds_s16(s_seg + c * 2)) && (segq != 0) &&
                                        (!p.mask.causal || posq >= lds_s16(s_pos + c * 2)));
              pr[t] = vis ? ex2(x - lse2) : 0.f;
              ds[k] = pr[t] * (__uint_as_float(vp[k]) - delta);
            }
            w[e] = pack_bf16x2(pr[0], pr[1]);
          }
          st_shared_v4(s_pds + tile_unit_off(r, hf * 8 + cc * 4 + u), w[0], w[1], w[2], w[3]);
        }
        // Toeplitz reduction without smem atomics: lane l takes element k of lane (l+k)%32,
        // which lies on diagonal (row - col) = r0 - c0 + l  (or l - 32 after the wrap).
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float y = __shfl_sync(0xffffffffu, ds[k], (lane + k) & 31);
          const bool wrap = lane + k >= 32;
          if (cc == 0) { if (wrap) a1 += y; else a0 += y; }
          else         { if (wrap) a2 += y; else a1 += y; }
        }
#pragma unroll
        for (int e = 0; e < 16; ++e)
          dsp[cc * 16 + e] = pack_bf16x2(ds[2 * e] * p.scale, ds[2 * e + 1] * p.scale);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(p_ready));
      if (p.drel != nullptr) {
        // (row - col) of this warp's sub-tile starts at q4*32 - hf*64
        // index of diagonal (row - col) = q4*32 - hf*64 + lane; the -32 / -64 neighbours of
        // the outermost lanes are structurally empty and may fall outside [0, 2L-2].
        const int idx = (i0 - j0) + (p.L - 1) + (q4 * 32 - hf * 64) + lane;
        float* dr = p.drel + static_cast<long long>(h) * (2 * p.L - 1);
        if (idx <= 2 * p.L - 2) atomicAdd(dr + idx, a0);
        if (idx - 32 >= 0) atomicAdd(dr + idx - 32, a1);
        if (idx - 64 >= 0) atomicAdd(dr + idx - 64, a2);
      }
      mbar_wait(smem_u32(p_consumed), it & 1);          // dV MMA has read P: overwrite with dS
#pragma unroll
      for (int u = 0; u < 8; ++u)
        st_shared_v4(s_pds + tile_unit_off(r, hf * 8 + u), dsp[u * 4], dsp[u * 4 + 1],
                     dsp[u * 4 + 2], dsp[u * 4 + 3]);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(ds_ready));
      mbar_wait(smem_u32(dq_full), it & 1);
      tc_fence_after();
      float* dq = p.dq_acc + ((bl + i) * p.H + h) * kD + hf * 64;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_s + lane_off + hf * 64 + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e)
          red_add_v4_f32(dq + cc * 32 + e * 4, __uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]),
                         __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(dq_read));
      ++it;
    }
    // ---- dK, dV: TMEM → bf16 → global (row = key) ----
    // (the last dq_full commit covers every MMA issued before it, including dV/dK)
    const long long orow = ((bl + j0 + r) * p.H + h) * kD + hf * 64;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t t_src = which ? t_dk : t_dv;
      __nv_bfloat16* dst = (which ? p.dk : p.dv) + orow;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_src + lane_off + hf * 64 + cc * 32, v);
        tmem_ld_wait();
        if (it == 0) {                       // no query block sees this key block: zero grads
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          int4 w;
          w.x = pack_bf16x2(__uint_as_float(v[u * 8 + 0]), __uint_as_float(v[u * 8 + 1]));
          w.y = pack_bf16x2(__uint_as_float(v[u * 8 + 2]), __uint_as_float(v[u * 8 + 3]));
          w.z = pack_bf16x2(__uint_as_float(v[u * 8 + 4]), __uint_as_float(v[u * 8 + 5]));
          w.w = pack_bf16x2(__uint_as_float(v[u * 8 + 6]), __uint_as_float(v[u * 8 + 7]));
          st_v4(dst + cc * 32 + u * 8, w);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// (pos_min, pos_max, seg_min, seg_max) of every 128-row block; one warp per block.
__global__ void __launch_bounds__(128)
flash_blkmeta_kernel(const int* __restrict__ seg, const int* __restrict__ pos,
                     int4* __restrict__ out, int n_blocks, int blocks_per_row) {
  const int blk = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (blk >= n_blocks) return;
  const int row0 = (blk % blocks_per_row) * kT;
  int pmin = 0x7fffffff, pmax = -0x7fffffff, smin = 0x7fffffff, smax = -0x7fffffff;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const long long idx = static_cast<long long>(blk) * kT + lane * 4 + t;
    const int sv = seg ? seg[idx] : 1;
    const int pv = pos ? pos[idx] : row0 + lane * 4 + t;
    pmin = min(pmin, pv); pmax = max(pmax, pv);
    smin = min(smin, sv); smax = max(smax, sv);
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    pmin = min(pmin, __shfl_xor_sync(0xffffffffu, pmin, m));
    pmax = max(pmax, __shfl_xor_sync(0xffffffffu, pmax, m));
    smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, m));
    smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, m));
  }
  if (lane == 0) out[blk] = make_int4(pmin, pmax, smin, smax);
}

// delta[b,h,l] = Σ_d dO[b,l,h,d]·O[b,l,h,d]   (both [B, L, H, D] contiguous)
__global__ void __launch_bounds__(256)
flash_delta_kernel(const __nv_bfloat16* __restrict__ d_o, const __nv_bfloat16* __restrict__ o,
                   float* __restrict__ delta, int B, int L, int H) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<long long>(B) * L * H) return;
  const int h = static_cast<int>(row % H);
  const long long bl = row / H;
  const int l = static_cast<int>(bl % L);
  const int b = static_cast<int>(bl / L);
  const uint2 a = *reinterpret_cast<const uint2*>(d_o + row * 128 + lane * 4);
  const uint2 c = *reinterpret_cast<const uint2*>(o + row * 128 + lane * 4);
  const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y);
  const float2 c0 = unpack_bf16x2(c.x), c1 = unpack_bf16x2(c.y);
  float s = a0.x * c0.x + a0.y * c0.y + a1.x * c1.x + a1.y * c1.y;
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if (lane == 0) delta[(static_cast<long long>(b) * H + h) * L + l] = s;
}

// dq (bf16) = dq_acc (fp32)
__global__ void __launch_bounds__(256)
flash_cast_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long long n4) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = src[i];
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    dst[i] = o;
  }
}

void CheckQkv(const torch::Tensor& t, const torch::Tensor& q, const char* name) {
  TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kBFloat16 && t.dim() == 4 &&
                  t.sizes() == q.sizes() && t.stride(3) == 1 && t.stride(2) == kD &&
                  t.stride(1) % 8 == 0 && t.stride(0) % 8 == 0 &&
                  reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0,
              "flash_attn: ", name, " must be a bf16 [B,L,H,128] view with strides (*, *, 128, 1)");
}

const int* IntPtr(const c10::optional<torch::Tensor>& t, int64_t B, int64_t L, const char* name) {
  if (!t.has_value() || !t->defined()) return nullptr;
  TORCH_CHECK(t->is_cuda() && t->scalar_type() == torch::kInt32 && t->is_contiguous() &&
                  t->numel() == B * L, "flash_attn: ", name, " must be int32 [B, L]");
  return t->data_ptr<int>();
}

// [B, L/128, 4] int32 block summaries (empty tensor when neither seg nor pos is given).
torch::Tensor BlockMeta(const int* seg, const int* pos, int64_t B, int64_t L,
                        const torch::TensorOptions& opts) {
  if (seg == nullptr && pos == nullptr) return torch::empty({0}, opts.dtype(torch::kInt32));
  auto out = torch::empty({B, L / kT, 4}, opts.dtype(torch::kInt32));
  const int n = static_cast<int>(B * (L / kT));
  flash_blkmeta_kernel<<<(n + 3) / 4, 128, 0, at::cuda::getCurrentCUDAStream()>>>(
      seg, pos, reinterpret_cast<int4*>(out.data_ptr<int>()), n, static_cast<int>(L / kT));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  CountLaunch();
  return out;
}

void BindContext() {
  static thread_local bool bound = false;
  if (bound) return;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(at::cuda::getCurrentCUDAStream(), &cap);
  if (cap == cudaStreamCaptureStatusNone) {
    C10_CUDA_CHECK(cudaFree(nullptr));
    bound = true;
  }
}

}  // namespace

// Returns (out [B,L,H,D] bf16, lse [B,H,L] fp32).
std::vector<torch::Tensor> flash_attn_fwd(const torch::Tensor& q, const torch::Tensor& k,
                                          const torch::Tensor& v,
                                          const c10::optional<torch::Tensor>& rel,
                                          const c10::optional<torch::Tensor>& seg,
                                          const c10::optional<torch::Tensor>& pos, double scale,
                                          bool causal) {
  TORCH_CHECK(q.dim() == 4 && q.size(3) == kD && q.size(1) % kT == 0,
              "flash_attn_fwd: needs D == 128 and L % 128 == 0");
  CheckQkv(q, q, "q"); CheckQkv(k, q, "k"); CheckQkv(v, q, "v");
  const int64_t B = q.size(0), L = q.size(1), H = q.size(2);
  TORCH_CHECK(L <= 32768, "flash_attn: L <= 32768");
  const c10::cuda::CUDAGuard guard(q.device());
  BindContext();
  FwdParams p;
  p.rel = nullptr;
  torch::Tensor relc;
  if (rel.has_value() && rel->defined()) {
    TORCH_CHECK(rel->scalar_type() == torch::kFloat32 && rel->dim() == 2 && rel->size(0) == H &&
                    rel->size(1) == 2 * L - 1, "flash_attn: rel must be fp32 [H, 2L-1]");
    relc = rel->contiguous();
    p.rel = relc.data_ptr<float>();
  }
  p.mask.seg = IntPtr(seg, B, L, "segment_ids");
  p.mask.pos = IntPtr(pos, B, L, "segment_pos");
  p.mask.causal = causal ? 1 : 0;
  auto blk = BlockMeta(p.mask.seg, p.mask.pos, B, L, q.options());
  p.mask.blk = blk.numel() ? reinterpret_cast<const int4*>(blk.data_ptr<int>()) : nullptr;
  auto out = torch::empty({B, L, H, kD}, q.options());
  auto lse = torch::empty({B, H, L}, q.options().dtype(torch::kFloat32));
  p.out = reinterpret_cast<__nv_bfloat16*>(out.data_ptr());
  p.lse = lse.data_ptr<float>();
  p.B = static_cast<int>(B); p.H = static_cast<int>(H); p.L = static_cast<int>(L);
  p.scale = static_cast<float>(scale);
  auto mk = [&](const torch::Tensor& t) {
    return MakeMap(t.data_ptr(), H * kD, L, B, t.stride(1), t.stride(0), kT);
  };
  const CUtensorMap mq = mk(q), mkk = mk(k), mv = mk(v);
  static bool configured = false;
  if (!configured) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(kFwdSmem)));
    configured = true;
  }
  dim3 grid(static_cast<unsigned>(L / kT), static_cast<unsigned>(H), static_cast<unsigned>(B));
  flash_fwd_kernel<<<grid, kThreads, kFwdSmem, at::cuda::getCurrentCUDAStream()>>>(mq, mkk, mv, p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  CountLaunch();
  return {out, lse};
}

// Returns (dq, dk, dv [B,L,H,D] bf16, drel [H,2L-1] fp32 or an empty tensor).
std::vector<torch::Tensor> flash_attn_bwd(const torch::Tensor& q, const torch::Tensor& k,
                                          const torch::Tensor& v, const torch::Tensor& out,
                                          const torch::Tensor& d_out, const torch::Tensor& lse,
                                          const c10::optional<torch::Tensor>& rel,
                                          const c10::optional<torch::Tensor>& seg,
                                          const c10::optional<torch::Tensor>& pos, double scale,
                                          bool causal, bool need_drel) {
  CheckQkv(q, q, "q"); CheckQkv(k, q, "k"); CheckQkv(v, q, "v");
  const int64_t B = q.size(0), L = q.size(1), H = q.size(2);
  TORCH_CHECK(q.size(3) == kD && L % kT == 0 && L <= 32768);
  TORCH_CHECK(out.is_contiguous() && out.sizes() == q.sizes() && out.scalar_type() == torch::kBFloat16);
  TORCH_CHECK(lse.scalar_type() == torch::kFloat32 && lse.is_contiguous() && lse.numel() == B * H * L);
  const c10::cuda::CUDAGuard guard(q.device());
  BindContext();
  auto d_o = d_out.contiguous();
  CheckQkv(d_o, q, "dO");
  auto stream = at::cuda::getCurrentCUDAStream();
  auto delta = torch::empty({B, H, L}, lse.options());
  {
    const long long rows = B * L * H;
    flash_delta_kernel<<<static_cast<int>((rows * 32 + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(d_o.data_ptr()),
        reinterpret_cast<const __nv_bfloat16*>(out.data_ptr()), delta.data_ptr<float>(),
        static_cast<int>(B), static_cast<int>(L), static_cast<int>(H));
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    CountLaunch();
  }
  BwdParams p;
  p.rel = nullptr;
  torch::Tensor relc, drel;
  if (rel.has_value() && rel->defined()) {
    TORCH_CHECK(rel->scalar_type() == torch::kFloat32 && rel->dim() == 2 && rel->size(0) == H &&
                rel->size(1) == 2 * L - 1);
    relc = rel->contiguous();
    p.rel = relc.data_ptr<float>();
  }
  p.drel = nullptr;
  if (need_drel) {
    drel = torch::zeros({H, 2 * L - 1}, lse.options());
    p.drel = drel.data_ptr<float>();
  } else {
    drel = torch::empty({0}, lse.options());
  }
  p.mask.seg = IntPtr(seg, B, L, "segment_ids");
  p.mask.pos = IntPtr(pos, B, L, "segment_pos");
  p.mask.causal = causal ? 1 : 0;
  auto blk = BlockMeta(p.mask.seg, p.mask.pos, B, L, q.options());
  p.mask.blk = blk.numel() ? reinterpret_cast<const int4*>(blk.data_ptr<int>()) : nullptr;
  auto dq_acc = torch::zeros({B, L, H, kD}, lse.options());
  auto dk = torch::empty({B, L, H, kD}, q.options());
  auto dv = torch::empty({B, L, H, kD}, q.options());
  p.lse = lse.data_ptr<float>();
  p.delta = delta.data_ptr<float>();
  p.dq_acc = dq_acc.data_ptr<float>();
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk.data_ptr());
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv.data_ptr());
  p.B = static_cast<int>(B); p.H = static_cast<int>(H); p.L = static_cast<int>(L);
  p.scale = static_cast<float>(scale);
  auto mk = [&](const torch::Tensor& t) {
    return MakeMap(t.data_ptr(), H * kD, L, B, t.stride(1), t.stride(0), kT);
  };
  const CUtensorMap mq = mk(q), mkk = mk(k), mv = mk(v), mdo = mk(d_o);
  static bool configured = false;
  if (!configured) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(flash_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(kBwdSmem)));
    configured = true;
  }
  dim3 grid(static_cast<unsigned>(L / kT), static_cast<unsigned>(H), static_cast<unsigned>(B));
  flash_bwd_kernel<<<grid, kThreads, kBwdSmem, stream>>>(mq, mkk, mv, mdo, p);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  CountLaunch();
  auto dq = torch::empty({B, L, H, kD}, q.options());
  {
    const long long n4 = dq.numel() / 4;
    const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    long long blocks = (n4 + 255) / 256;
    if (blocks > sms * 8) blocks = sms * 8;
    flash_cast_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(dq_acc.data_ptr<float>()),
        reinterpret_cast<uint2*>(dq.data_ptr()), n4);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    CountLaunch();
  }
  return {dq, dk, dv, drel};
}

}  // namespace lb

LB_REGISTER(flash_attn) {
  m.attr("_has_flash_attn") = true;
  m.def("flash_attn_fwd", &lb::flash_attn_fwd, py::arg("q"), py::arg("k"), py::arg("v"),
        py::arg("rel"), py::arg("seg"), py::arg("pos"), py::arg("scale"), py::arg("causal"));
  m.def("flash_attn_bwd", &lb::flash_attn_bwd, py::arg("q"), py::arg("k"), py::arg("v"),
        py::arg("out"), py::arg("d_out"), py::arg("lse"), py::arg("rel"), py::arg("seg"),
        py::arg("pos"), py::arg("scale"), py::arg("causal"), py::arg("need_drel"));
}
