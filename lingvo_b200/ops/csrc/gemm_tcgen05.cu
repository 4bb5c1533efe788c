// Persistent, warp-specialised bf16 GEMM for sm_100a.
//
//   C[g, m, n] = epilogue( sum_k A[g, m, k] * B[g, n, k] )        g in [0, G)
//
// * operands are staged by TMA (cp.async.bulk.tensor, 128B swizzle) into a
//   4-stage shared-memory ring;  either operand may be K-major ([.., MN, K]
//   storage) or MN-major ([.., K, MN] storage), so the same kernel serves the
//   forward (x @ w), dgrad (dy @ w^T) and wgrad (x^T @ dy) of every linear /
//   expert-FFN layer without any transposes in HBM;
// * one elected thread issues tcgen05.mma (UMMA 128 x BN x 16, fp32 accumulate)
//   into one of two TMEM accumulator stages (2 x BN columns), so the epilogue of
//   tile i overlaps the main loop of tile i+1;
// * four epilogue warps read the accumulator with tcgen05.ld, apply the fused
//   epilogue (bias, activation, relu-gradient mask / residual add, per-row scale,
//   accumulate) and store 16-byte vectors straight to global memory.  The store
//   target of every output row can be redirected through a row-pointer table,
//   which is how the MoE combine (expert GEMM -> NVLink peer store) is fused.
//
// Grid: one CTA per SM (persistent), 256 threads:
//   warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
//   warps 4..7 = epilogue (TMEM lane quarter = warp_idx % 4).

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <mutex>
#include <unordered_map>

#include "ptx.cuh"
#include "registry.h"

namespace lb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kNumThreads = 384;   // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue

enum Act : int { kActNone = 0, kActRelu = 1, kActGelu = 2, kActSilu = 3, kActGeluTanh = 4,
                 kActSquaredRelu = 5 };
enum AuxMode : int { kAuxNone = 0, kAuxReluMask = 1, kAuxAdd = 2 };

// A operand may be split along K over several tensor maps (one per NVLink peer):
// the TMA producer then *is* the all-gather (tiles are pulled straight from the
// owning rank's memory), SURVEY K5/K6 "all-gather -> GEMM".
constexpr int kMaxAMaps = 8;
struct TmapArray {
  CUtensorMap m[kMaxAMaps];
};

struct GemmParams {
  int G, M, N, K;
  void* c;                 // [G, M, N] out (bf16 or fp32)
  long long ldc;           // row stride of C (elements)
  long long c_batch;       // batch stride of C (elements)
  const float* bias;       // [G, N] fp32 or nullptr
  const __nv_bfloat16* aux;  // [G, M, N] bf16 (same strides as C) or nullptr
  void* pre_act;           // optional second output: pre-activation (bf16), same strides
  const float* row_scale;  // [G, M] or nullptr
  void* const* row_ptrs;   // [G * M] destination row pointers (peer memory) or nullptr
  void* const* nblk_ptrs;  // [tiles_n] destination base per N-tile (peer slab) or nullptr;
                           // element (row, col) goes to base[n_blk] + row*ldc + (col - n0)
  int a_k_per_map;         // K elements covered by each A tensor map (all-gather prologue)
  int act;
  int aux_mode;
  int accumulate;          // C += result (fp32 out only)
};

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case kActRelu: return fmaxf(x, 0.f);
    case kActGelu: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    case kActSilu: return x / (1.f + __expf(-x));
    case kActGeluTanh: {
      float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.f + tanhf(u));
    }
    case kActSquaredRelu: { float r = fmaxf(x, 0.f); return r * r; }
    default: return x;
  }
}

struct EpiRow {
  void* c_row;
  long long row_off;
  const float* bias;
  float rscale;
  bool store_ok;
};

template <int kAct>
__device__ __forceinline__ void act32(float (&f)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = apply_act(f[i], kAct);
}

// One 32-column chunk of one output row: bias / pre-act tap / activation /
// aux (ReLU-mask or residual add) / row scale / convert / store.
template <typename OutT>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, const EpiRow& er,
                                               const uint32_t (&v)[32], int col0) {
  if (!er.store_ok || col0 >= p.N) return;
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  if (er.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if (col0 + i < p.N) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(er.bias + col0 + i));
        f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
      }
    }
  }
  if (p.pre_act != nullptr) {
    __nv_bfloat16* pr = reinterpret_cast<__nv_bfloat16*>(p.pre_act) + er.row_off + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      if (col0 + i < p.N) {
        int4 o;
        o.x = pack_bf16x2(f[i], f[i + 1]);
        o.y = pack_bf16x2(f[i + 2], f[i + 3]);
        o.z = pack_bf16x2(f[i + 4], f[i + 5]);
        o.w = pack_bf16x2(f[i + 6], f[i + 7]);
        st_v4(pr + i, o);
      }
    }
  }
  switch (p.act) {   // hoisted: one uniform branch per chunk, not per element
    case kActRelu: act32<kActRelu>(f); break;
    case kActGelu: act32<kActGelu>(f); break;
    case kActSilu: act32<kActSilu>(f); break;
    case kActGeluTanh: act32<kActGeluTanh>(f); break;
    case kActSquaredRelu: act32<kActSquaredRelu>(f); break;
    default: break;
  }
  if (p.aux_mode != kAuxNone) {
    const __nv_bfloat16* ax = p.aux + er.row_off + col0;
    const bool mask = p.aux_mode == kAuxReluMask;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      if (col0 + i < p.N) {
        const int4 a = *reinterpret_cast<const int4*>(ax + i);
        const uint32_t w[4] = {static_cast<uint32_t>(a.x), static_cast<uint32_t>(a.y),
                               static_cast<uint32_t>(a.z), static_cast<uint32_t>(a.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = unpack_bf16x2(w[j]);
          if (mask) {
            f[i + 2 * j] = t.x > 0.f ? f[i + 2 * j] : 0.f;
            f[i + 2 * j + 1] = t.y > 0.f ? f[i + 2 * j + 1] : 0.f;
          } else {
            f[i + 2 * j] += t.x;
            f[i + 2 * j + 1] += t.y;
          }
        }
      }
    }
  }
  if (p.row_scale != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] *= er.rscale;
  }
  if constexpr (sizeof(OutT) == 2) {
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(er.c_row) + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      if (col0 + i < p.N) {
        int4 o;
        o.x = pack_bf16x2(f[i], f[i + 1]);
        o.y = pack_bf16x2(f[i + 2], f[i + 3]);
        o.z = pack_bf16x2(f[i + 4], f[i + 5]);
        o.w = pack_bf16x2(f[i + 6], f[i + 7]);
        st_v4(out + i, o);
      }
    }
  } else {
    float* out = reinterpret_cast<float*>(er.c_row) + col0;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      if (col0 + i < p.N) {
        float4 o = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
        if (p.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(out + i);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(out + i) = o;
      }
    }
  }
}

template <int BN, bool kAK, bool kBK, typename OutT>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ TmapArray tmaps_a,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  constexpr int kStages = (BN == 256) ? 4 : 6;
  constexpr uint32_t kABytes = kBlockM * kBlockK * 2;   // 16 KiB
  constexpr uint32_t kBBytes = BN * kBlockK * 2;        // 32 KiB (BN=256)
  constexpr uint32_t kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = 2 * BN;                // two accumulator stages
  static_assert(kTmemCols <= 512, "TMEM overflow");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024B alignment for SWIZZLE_128B atoms.
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (p.M + kBlockM - 1) / kBlockM;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_per_group = tiles_m * tiles_n;
  const int num_tiles = tiles_per_group * p.G;
  const int num_k_blocks = (p.K + kBlockK - 1) / kBlockK;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmaps_a.m[0]);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tmem_full_bar[i]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[i]), 8);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc<kTmemCols>(smem_u32(tmem_ptr_smem));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int g = tile / tiles_per_group;
        const int r = tile - g * tiles_per_group;
        const int n_blk = r / tiles_m;
        const int m_blk = r - n_blk * tiles_m;
        const int m0 = m_blk * kBlockM, n0 = n_blk * BN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_arrive_expect_tx(fb, kStageBytes);
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
          const int k0 = kb * kBlockK;
          const int amap = k0 / p.a_k_per_map;
          const int ka = k0 - amap * p.a_k_per_map;
          if constexpr (kAK) {
            tma_load_3d(sa, &tmaps_a.m[amap], fb, ka, m0, g);
          } else {
#pragma unroll
            for (int i = 0; i < kBlockM / 64; ++i)
              tma_load_3d(sa + i * (64 * kBlockK * 2), &tmaps_a.m[amap], fb, m0 + 64 * i, ka, g);
          }
          if constexpr (kBK) {
            tma_load_3d(sb, &tmap_b, fb, k0, n0, g);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_3d(sb + i * (64 * kBlockK * 2), &tmap_b, fb, n0 + 64 * i, k0, g);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ============================= MMA issuer =============================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(1, 1, !kAK, !kBK, kBlockM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&tmem_empty_bar[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int j = 0; j < kBlockK / kUmmaK; ++j) {
            // K-major: advance 32 B inside the 128 B swizzle row.
            // MN-major: advance 16 K-rows of 128 B; 64-wide MN chunks are
            //           64*kBlockK*2 bytes apart (LBO), 8-row groups 1 KiB (SBO).
            const uint64_t adesc =
                kAK ? make_smem_desc_sw128(sa + j * 32, 16, 1024)
                    : make_smem_desc_sw128(sa + j * (kUmmaK * 128), 64 * kBlockK * 2, 1024);
            const uint64_t bdesc =
                kBK ? make_smem_desc_sw128(sb + j * 32, 16, 1024)
                    : make_smem_desc_sw128(sb + j * (kUmmaK * 128), 64 * kBlockK * 2, 1024);
            umma_f16(tmem_d, adesc, bdesc, idesc, (kb | j) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty_bar[stage]));   // frees the smem slot when MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&tmem_full_bar[acc]));   // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp_idx >= 4) {
    // ============================== epilogue ==============================
    // 8 warps: warp w owns TMEM lanes 32*(w%4).. and one half of the tile's columns.
    const int q = warp_idx & 3;                 // TMEM lane quarter
    const int half = (warp_idx - 4) >> 2;       // column half
    constexpr int kChunks = BN / 64;            // 32-column chunks per warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int g = tile / tiles_per_group;
      const int r = tile - g * tiles_per_group;
      const int n_blk = r / tiles_m;
      const int m_blk = r - n_blk * tiles_m;
      const int row = m_blk * kBlockM + q * 32 + lane;
      const int n0 = n_blk * BN;
      const bool row_ok = row < p.M;

      EpiRow er;
      er.row_off = static_cast<long long>(g) * p.c_batch + static_cast<long long>(row) * p.ldc;
      if (p.row_ptrs != nullptr && row_ok) {
        er.c_row = p.row_ptrs[static_cast<long long>(g) * p.M + row];
      } else if (p.nblk_ptrs != nullptr) {
        er.c_row = reinterpret_cast<OutT*>(p.nblk_ptrs[n_blk]) + static_cast<long long>(row) * p.ldc - n0;
      } else {
        er.c_row = reinterpret_cast<OutT*>(p.c) + er.row_off;
      }
      er.store_ok = row_ok && (er.c_row != nullptr);
      er.rscale = (p.row_scale != nullptr && row_ok)
                      ? p.row_scale[static_cast<long long>(g) * p.M + row] : 1.f;
      er.bias = p.bias ? p.bias + static_cast<long long>(g) * p.N : nullptr;

      mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_phase);
      tc_fence_after();

      // Software-pipelined TMEM reads: chunk c+1 is in flight while chunk c is
      // converted and stored.
      const uint32_t tbase = tmem_base + acc * BN + half * (BN / 2) +
                             (static_cast<uint32_t>(q * 32) << 16);
      const int cbase = n0 + half * (BN / 2);
      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32(tbase, va);
#pragma unroll 1
      for (int c = 0; c < kChunks; c += 2) {
        tmem_ld_wait();
        tmem_ld_32x32b_x32(tbase + (c + 1) * 32, vb);
        epilogue_chunk<OutT>(p, er, va, cbase + c * 32);
        tmem_ld_wait();
        if (c + 2 < kChunks) tmem_ld_32x32b_x32(tbase + (c + 2) * 32, va);
        epilogue_chunk<OutT>(p, er, vb, cbase + (c + 1) * 32);
      }
      // Release this accumulator stage back to the MMA warp.
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ host ----
CUtensorMap MakeMap(const void* base, int64_t inner, int64_t rows, int64_t groups,
                    int64_t row_stride, int64_t group_stride, int box_rows,
                    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, int elem_bytes = 2);

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                              const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn GetEncodeFn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    TORCH_CHECK(err == cudaSuccess && ptr != nullptr, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeFn>(ptr);
  });
  return fn;
}

// 3-D bf16 map over a [G, rows, inner] view; box = [1, box_rows, 64], 128B swizzle.
CUtensorMap MakeMap(const void* base, int64_t inner, int64_t rows, int64_t groups,
                    int64_t row_stride, int64_t group_stride, int box_rows,
                    CUtensorMapDataType dt, int elem_bytes) {
  CUtensorMap m;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(groups)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_stride * elem_bytes),
                           static_cast<cuuint64_t>(group_stride * elem_bytes)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / elem_bytes), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = GetEncodeFn()(&m, dt, 3, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TORCH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: ", static_cast<int>(r),
              " inner=", inner, " rows=", rows, " groups=", groups, " row_stride=", row_stride,
              " group_stride=", group_stride);
  return m;
}

template <int BN, bool kAK, bool kBK, typename OutT>
static void Launch(const TmapArray& ta, const CUtensorMap& tb, const GemmParams& p,
                   cudaStream_t stream, int num_sms) {
  constexpr int kStages = (BN == 256) ? 4 : 6;
  constexpr size_t smem = kStages * (kBlockM * kBlockK * 2 + BN * kBlockK * 2) + 1024 + 256;
  auto kern = gemm_tcgen05_kernel<BN, kAK, kBK, OutT>;
  static bool configured = false;
  if (!configured) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    configured = true;
  }
  const int tiles = p.G * ((p.M + kBlockM - 1) / kBlockM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, kNumThreads, smem, stream>>>(ta, tb, p);
  CountLaunch();
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

template <int BN, typename OutT>
static void Dispatch(bool ak, bool bk, const TmapArray& ta, const CUtensorMap& tb,
                     const GemmParams& p, cudaStream_t s, int sms) {
  if (ak && bk) Launch<BN, true, true, OutT>(ta, tb, p, s, sms);
  else if (ak && !bk) Launch<BN, true, false, OutT>(ta, tb, p, s, sms);
  else if (!ak && bk) Launch<BN, false, true, OutT>(ta, tb, p, s, sms);
  else Launch<BN, false, false, OutT>(ta, tb, p, s, sms);
}

// Wave-quantisation heuristic switch. Off by default: measured on B200
// (profiles/gemm_bn_probe_r1.jsonl) the 128-wide tile runs at ≈75 % of the 256-wide tile's
// per-FLOP rate (operand traffic per FLOP is 1.33× higher), which costs more than the
// 3.46 → 4 wave rounding it avoids. LINGVO_B200_GEMM_AUTO_BN=1 turns it on for experiments.
static bool GemmAutoBn128() {
  static const bool on = [] {
    const char* e = getenv("LINGVO_B200_GEMM_AUTO_BN");
    return e != nullptr && atoi(e) != 0;
  }();
  return on;
}

// a: [G, M, K] if a_kmajor else [G, K, M];  b: [G, N, K] if b_kmajor else [G, K, N]
// (2-D inputs are treated as G = 1).  Returns / fills out [G, M, N].
torch::Tensor gemm_bf16(const torch::Tensor& a_in, const torch::Tensor& b_in, bool a_kmajor,
                        bool b_kmajor, const c10::optional<torch::Tensor>& bias,
                        int64_t act, const c10::optional<torch::Tensor>& aux, int64_t aux_mode,
                        const c10::optional<torch::Tensor>& row_scale,
                        const c10::optional<torch::Tensor>& out_opt, bool out_fp32,
                        bool accumulate, const c10::optional<torch::Tensor>& pre_act,
                        const c10::optional<torch::Tensor>& row_ptrs,
                        const c10::optional<torch::Tensor>& nblk_ptrs,
                        const c10::optional<torch::Tensor>& a_peer_ptrs, int64_t nblk_ld) {
  TORCH_CHECK(a_in.is_cuda() && b_in.is_cuda(), "gemm_bf16: CUDA tensors required");
  TORCH_CHECK(a_in.scalar_type() == torch::kBFloat16 && b_in.scalar_type() == torch::kBFloat16,
              "gemm_bf16: bf16 inputs required");
  const c10::cuda::CUDAGuard guard(a_in.device());
  // cuTensorMapEncodeTiled is a *driver* call: make sure this thread (e.g. an
  // autograd worker that has not issued a runtime call yet) has the primary
  // context bound, otherwise it fails with CUDA_ERROR_INVALID_CONTEXT.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    // (skipped while a stream capture is in progress: cudaFree is not capturable, and a
    // thread that reaches this point during capture already ran the warm-up steps)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(at::cuda::getCurrentCUDAStream(), &cap);
    if (cap == cudaStreamCaptureStatusNone) {
      C10_CUDA_CHECK(cudaFree(nullptr));
      ctx_bound = true;
    }
  }
  torch::Tensor a = a_in.dim() == 2 ? a_in.unsqueeze(0) : a_in;
  torch::Tensor b = b_in.dim() == 2 ? b_in.unsqueeze(0) : b_in;
  TORCH_CHECK(a.dim() == 3 && b.dim() == 3, "gemm_bf16: 2-D or 3-D operands");
  if (a.stride(2) != 1) a = a.contiguous();
  if (b.stride(2) != 1) b = b.contiguous();
  const int64_t G = a.size(0);
  TORCH_CHECK(b.size(0) == G, "gemm_bf16: group mismatch");
  const int64_t M = a_kmajor ? a.size(1) : a.size(2);
  const int64_t K = a_kmajor ? a.size(2) : a.size(1);
  const int64_t N = b_kmajor ? b.size(1) : b.size(2);
  const int64_t Kb = b_kmajor ? b.size(2) : b.size(1);
  const bool ag_prologue = a_peer_ptrs.has_value() && a_peer_ptrs->defined();
  TORCH_CHECK(ag_prologue || K == Kb, "gemm_bf16: K mismatch ", K, " vs ", Kb);
  TORCH_CHECK(N % 8 == 0, "gemm_bf16: N must be a multiple of 8");
  auto check_op = [](const torch::Tensor& t) {
    TORCH_CHECK(t.stride(1) % 8 == 0 && (t.size(0) == 1 || t.stride(0) % 8 == 0) &&
                    reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0,
                "gemm_bf16: operand rows must be 16-byte aligned");
  };
  check_op(a);
  check_op(b);

  torch::Tensor out;
  if (out_opt.has_value()) {
    out = out_opt.value();
    TORCH_CHECK(out.stride(-1) == 1, "gemm_bf16: out must be row-contiguous");
  } else {
    out = torch::empty({G, M, N}, a.options().dtype(out_fp32 ? torch::kFloat32 : torch::kBFloat16));
  }
  const bool fp32 = out.scalar_type() == torch::kFloat32;
  TORCH_CHECK(fp32 || out.scalar_type() == torch::kBFloat16, "gemm_bf16: out dtype");
  torch::Tensor out3 = out.dim() == 2 ? out.unsqueeze(0) : out;
  TORCH_CHECK(out3.size(0) == G && out3.size(1) == M && out3.size(2) == N, "gemm_bf16: out shape");

  GemmParams p;
  p.G = static_cast<int>(G); p.M = static_cast<int>(M); p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.c = out3.data_ptr();
  p.ldc = out3.stride(1);
  p.c_batch = G > 1 ? out3.stride(0) : 0;
  p.bias = nullptr; p.aux = nullptr; p.row_scale = nullptr; p.pre_act = nullptr;
  p.row_ptrs = nullptr;
  p.nblk_ptrs = nullptr;
  p.a_k_per_map = static_cast<int>(K);
  torch::Tensor bias_f;
  if (bias.has_value() && bias->defined()) {
    bias_f = bias->to(torch::kFloat32).contiguous();
    TORCH_CHECK(bias_f.numel() == G * N, "gemm_bf16: bias must be [G, N]");
    p.bias = bias_f.data_ptr<float>();
  }
  if (aux.has_value() && aux->defined() && aux_mode != 0) {
    torch::Tensor ax = aux.value();
    ax = ax.dim() == 2 ? ax.unsqueeze(0) : ax;
    TORCH_CHECK(ax.scalar_type() == torch::kBFloat16 && ax.stride(1) == p.ldc &&
                    (G == 1 || ax.stride(0) == p.c_batch) && ax.stride(2) == 1,
                "gemm_bf16: aux must be bf16 with the same strides as out");
    p.aux = reinterpret_cast<const __nv_bfloat16*>(ax.data_ptr());
  }
  if (pre_act.has_value() && pre_act->defined()) {
    torch::Tensor pa = pre_act.value();
    pa = pa.dim() == 2 ? pa.unsqueeze(0) : pa;
    TORCH_CHECK(pa.scalar_type() == torch::kBFloat16 && pa.stride(1) == p.ldc &&
                    (G == 1 || pa.stride(0) == p.c_batch), "gemm_bf16: pre_act strides");
    p.pre_act = pa.data_ptr();
  }
  torch::Tensor rs_f;
  if (row_scale.has_value() && row_scale->defined()) {
    rs_f = row_scale->to(torch::kFloat32).contiguous();
    TORCH_CHECK(rs_f.numel() == G * M, "gemm_bf16: row_scale must be [G, M]");
    p.row_scale = rs_f.data_ptr<float>();
  }
  if (row_ptrs.has_value() && row_ptrs->defined()) {
    TORCH_CHECK(row_ptrs->scalar_type() == torch::kInt64 && row_ptrs->numel() == G * M &&
                    row_ptrs->is_contiguous(), "gemm_bf16: row_ptrs must be int64 [G, M]");
    p.row_ptrs = reinterpret_cast<void* const*>(row_ptrs->data_ptr());
  }
  if (nblk_ptrs.has_value() && nblk_ptrs->defined()) {
    TORCH_CHECK(nblk_ptrs->scalar_type() == torch::kInt64 && nblk_ptrs->is_cuda() &&
                    nblk_ptrs->is_contiguous(), "gemm_bf16: nblk_ptrs must be CUDA int64");
    TORCH_CHECK(G == 1, "gemm_bf16: nblk_ptrs needs G == 1");
    TORCH_CHECK(!fp32 && N > 128 && nblk_ld > 0 && nblk_ld % 8 == 0 &&
                    nblk_ptrs->numel() == (N + 255) / 256,
                "gemm_bf16: nblk_ptrs needs bf16 out, one pointer per 256-column tile and nblk_ld");
    p.nblk_ptrs = reinterpret_cast<void* const*>(nblk_ptrs->data_ptr());
    p.ldc = nblk_ld;
    TORCH_CHECK(p.aux == nullptr && p.pre_act == nullptr, "gemm_bf16: nblk_ptrs excludes aux/pre_act");
  }
  p.act = static_cast<int>(act);
  p.aux_mode = p.aux ? static_cast<int>(aux_mode) : 0;
  p.accumulate = accumulate ? 1 : 0;
  TORCH_CHECK(!accumulate || fp32, "gemm_bf16: accumulate needs fp32 out");

  if (M == 0 || N == 0 || G == 0) return out;

  // Tile-N: 256 unless N is small — or unless 128-wide tiles quantise better onto the SMs:
  // an [8192, 2048] output is 512 tiles of 128×256 = 3.46 waves on 148 SMs (4 are paid for),
  // but 1024 tiles of 128×128 = 6.92 half-cost waves. LINGVO_B200_GEMM_BN=128|256 forces one.
  bool bn256 = N > 128;
  if (bn256) {
    static const int forced = [] {
      const char* e = getenv("LINGVO_B200_GEMM_BN");
      return e ? atoi(e) : 0;
    }();
    const int sms_q = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
    const int64_t mt = (M + kBlockM - 1) / kBlockM;
    const int64_t t256 = G * mt * ((N + 255) / 256), t128 = G * mt * ((N + 127) / 128);
    const int64_t w256 = 2 * ((t256 + sms_q - 1) / sms_q), w128 = (t128 + sms_q - 1) / sms_q;
    if (forced == 128) bn256 = false;
    else if (forced == 0 && GemmAutoBn128() && w128 * 100 <= w256 * 90) bn256 = false;
  }
  const int bn = bn256 ? 256 : 128;
  TmapArray ta;
  if (a_peer_ptrs.has_value() && a_peer_ptrs->defined()) {
    // `a` is the local K-shard [M, K/W] (K-major); peers hold the other shards at
    // the given base pointers with identical strides.
    TORCH_CHECK(a_kmajor && G == 1, "all-gather prologue needs a K-major 2-D A");
    TORCH_CHECK(a_peer_ptrs->device().is_cpu() && a_peer_ptrs->scalar_type() == torch::kInt64);
    const int W = static_cast<int>(a_peer_ptrs->numel());
    TORCH_CHECK(W <= kMaxAMaps);
    const int64_t k_shard = a.size(2);
    TORCH_CHECK(k_shard % kBlockK == 0, "K shard must be a multiple of 64");
    TORCH_CHECK(Kb == k_shard * W, "B's K must equal W * shard K");
    p.K = static_cast<int>(k_shard * W);
    p.a_k_per_map = static_cast<int>(k_shard);
    for (int r = 0; r < W; ++r)
      ta.m[r] = MakeMap(reinterpret_cast<void*>(a_peer_ptrs->data_ptr<int64_t>()[r]), k_shard, M,
                        1, a.stride(1), a.stride(1) * a.size(1), kBlockM);
  } else {
    ta.m[0] = a_kmajor
        ? MakeMap(a.data_ptr(), K, M, G, a.stride(1), G > 1 ? a.stride(0) : a.stride(1) * a.size(1), kBlockM)
        : MakeMap(a.data_ptr(), M, K, G, a.stride(1), G > 1 ? a.stride(0) : a.stride(1) * a.size(1), kBlockK);
  }
  CUtensorMap tb = b_kmajor
      ? MakeMap(b.data_ptr(), Kb, N, G, b.stride(1), G > 1 ? b.stride(0) : b.stride(1) * b.size(1), bn)
      : MakeMap(b.data_ptr(), N, Kb, G, b.stride(1), G > 1 ? b.stride(0) : b.stride(1) * b.size(1), kBlockK);

  cudaStream_t stream = at::cuda::getCurrentCUDAStream();
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  if (bn256) {
    if (fp32) Dispatch<256, float>(a_kmajor, b_kmajor, ta, tb, p, stream, sms);
    else Dispatch<256, __nv_bfloat16>(a_kmajor, b_kmajor, ta, tb, p, stream, sms);
  } else {
    if (fp32) Dispatch<128, float>(a_kmajor, b_kmajor, ta, tb, p, stream, sms);
    else Dispatch<128, __nv_bfloat16>(a_kmajor, b_kmajor, ta, tb, p, stream, sms);
  }
  return out;
}

}  // namespace lb
