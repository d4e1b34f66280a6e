// Dense projection GEMM on the 5th-gen tensor cores (tcgen05 + TMEM accumulators + TMA-fed operands).
//
//   out[m, n] = epilogue( sum_k x[m, k] * w[n, k] )        x: [M, K] row-major, w: [N, K] row-major
//
// This is the "nn.Linear" contraction used by every dense projection on the hot path
// (reference call sites: HF Qwen3Attention/Qwen3MLP q/k/v/o/gate/up/down invoked from
// modeling_asteroid.py:226,273-284; the 8 lm_heads modeling_asteroid.py:412; codec nn.Linear /
// ConvTranspose1d-as-GEMM XY_Tokenizer/xy_tokenizer/nn/modules.py:84-87,181-182,494-500,1111-1115).
//
// B200 design (not a translation of anything in the reference, which only calls aten::linear):
//   * swap-AB: the WEIGHT tile (128 rows of w) is the UMMA "A"/M operand and the activation tile
//     (BN rows of x, BN in {16..256}) is the UMMA "B"/N operand, so a batch-1 decode step still
//     issues full 128-lane MMAs and the kernel degenerates to a pure weight-streaming pass that is
//     HBM-bound; large M (prefill, codec) uses BN = 128/256.
//   * operands arrive by TMA (cp.async.bulk.tensor, 128B swizzle) into a kStages-deep mbarrier ring;
//     one elected thread issues tcgen05.mma; accumulators live in TMEM; four epilogue warps read
//     them back with tcgen05.ld and apply the fused epilogue (bias / GELU / layer-scale / residual /
//     SwiGLU / bf16 rounding points of the reference).
//   * split-K (grid.z) keeps all 148 SMs pulling weights when N/128 tiles are few; the partial sums
//     are reduced in fixed split order by the last CTA to arrive (no spin-waits, deterministic).
#include "common.cuh"
#include "sm100.cuh"
#include "mtts_internal.h"

#include <mutex>
#include <unordered_map>
#include <string>
#include <string.h>

using namespace sm100;

namespace {

constexpr int kBlockW = 128;      // weight rows per tile == UMMA M
constexpr int kSwizzleBytes = 128;
constexpr int kNumThreads = 192;  // warp0: TMA, warp1: MMA + TMEM alloc, warps 2-5: epilogue

struct GemmParams {
  int M, N, K;
  int kb_total;      // ceil(K / BLOCK_K)
  int kb_per_split;  // k-blocks handled by each grid.z slice
  int splits;
  void* out;
  long long ldo;
  int out_bf16;  // 1: bf16 output, 0: fp32
  const float* bias;
  const float* gamma;
  const void* residual;
  long long ldr;
  int flags;
  float* ws;      // split-K partials [splits][tiles][BN][128]
  int* counters;  // one per output tile, zero on entry, zero on exit
};

template <typename T>
struct Traits;
template <>
struct Traits<bf16> {
  static constexpr int kBlockK = 64, kUmmaK = 16, kFmt = 1;
};
template <>
struct Traits<float> {
  static constexpr int kBlockK = 32, kUmmaK = 8, kFmt = 2;
};

template <int BN>
constexpr uint32_t tmem_cols() {
  return BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
}

template <int BN, int kStages>
constexpr int smem_bytes() {
  return kStages * (kBlockW * kSwizzleBytes + BN * kSwizzleBytes) + 1024 /*align slack*/ + 256 /*barriers*/;
}

__device__ __forceinline__ float apply_epilogue_scalar(float v, int n, int m, const GemmParams& p) {
  if (p.flags & MTTS_EPI_BIAS) v += __ldg(p.bias + n);
  if (p.flags & MTTS_EPI_GELU) v = gelu_erf(v);
  if (p.out_bf16) v = bf16_round(v);  // the reference materialises the bf16 linear output first
  if (p.flags & MTTS_EPI_GAMMA) v *= __ldg(p.gamma + n);
  if (p.flags & MTTS_EPI_RESIDUAL) {
    float r = p.out_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p.residual)[(long long)m * p.ldr + n])
                         : reinterpret_cast<const float*>(p.residual)[(long long)m * p.ldr + n];
    v = r + v;
  }
  return v;
}

template <typename T, int BN, int kStages>
__global__ void __launch_bounds__(kNumThreads) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_w,
                                                              const __grid_constant__ CUtensorMap tmap_x,
                                                              const GemmParams p) {
  constexpr int BK = Traits<T>::kBlockK;
  constexpr int UK = Traits<T>::kUmmaK;
  constexpr uint32_t kABytes = kBlockW * kSwizzleBytes;
  constexpr uint32_t kBBytes = BN * kSwizzleBytes;
  constexpr uint32_t kCols = tmem_cols<BN>();
  constexpr uint32_t kIdesc = make_idesc(Traits<T>::kFmt, kBlockW, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kStages * kBBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  int* is_last_smem = reinterpret_cast<int*>(tmem_ptr_smem + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, split = blockIdx.z;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
  const int num_kb = kb_end - kb_begin;  // host guarantees >= 1

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<kCols>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      // weights are streamed once per launch when M fits one tile; keep activations resident in L2.
      const uint64_t pol_w = (gridDim.y == 1) ? kEvictFirst : kEvictNormal;
      const uint64_t pol_x = kEvictLast;
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], kABytes + kBBytes);
        const int kc = (kb_begin + it) * BK;
        tma_load_2d(smem_a + s * kABytes, &tmap_w, &full_bar[s], kc, n_tile * kBlockW, pol_w);
        tma_load_2d(smem_b + s * kBBytes, &tmap_x, &full_bar[s], kc, m_tile * BN, pol_x);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + s * kABytes);
        const uint32_t b_addr = smem_u32(smem_b + s * kBBytes);
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * UK * (int)sizeof(T));
          const uint64_t db = make_smem_desc_sw128(b_addr + k * UK * (int)sizeof(T));
          if constexpr (sizeof(T) == 2)
            umma_bf16(tmem_base, da, db, kIdesc, (it > 0 || k > 0) ? 1u : 0u);
          else
            umma_tf32(tmem_base, da, db, kIdesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ================= epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =================
    const int quarter = warp & 3;
    const int n_local = quarter * 32 + lane;
    const int n = n_tile * kBlockW + n_local;
    const int epi_tid = threadIdx.x - 64;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int m_base = m_tile * BN;
    const int tile_id = m_tile * gridDim.x + n_tile;
    const int num_tiles = gridDim.x * gridDim.y;
    bool do_final = true;

    if (p.splits > 1) {
      float* wsp = p.ws + ((long long)(split * num_tiles + tile_id) * BN) * kBlockW + n_local;
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(taddr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) __stcg(wsp + (c + j) * kBlockW, __uint_as_float(r[j]));
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (epi_tid == 0) {
        const int prev = atomicAdd(p.counters + tile_id, 1);
        const int last = (prev == p.splits - 1);
        if (last) p.counters[tile_id] = 0;  // leave the counter clean for the next launch
        *is_last_smem = last;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      do_final = (*is_last_smem != 0);
      if (do_final) __threadfence();
    }

    if (do_final) {
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        float acc[16];
        if (p.splits > 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = 0.f;
          for (int s = 0; s < p.splits; ++s) {  // fixed order -> bitwise deterministic
            const float* wsp = p.ws + ((long long)(s * num_tiles + tile_id) * BN) * kBlockW + n_local;
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] += __ldcg(wsp + (c + j) * kBlockW);
          }
        } else {
          uint32_t r[16];
          tmem_ld_32x32b_x16(taddr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(r[j]);
        }
        if (p.flags & MTTS_EPI_SWIGLU) {
          // weight rows are interleaved (2j = gate_j, 2j+1 = up_j): neighbouring lanes pair up.
          // Rounding points follow Qwen3MLP in bf16: bf16(gate), bf16(silu), bf16(up), bf16(product).
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m_base + c + j;
            float v = p.out_bf16 ? bf16_round(acc[j]) : acc[j];
            float other = __shfl_xor_sync(0xffffffffu, v, 1);
            if ((lane & 1) == 0 && m < p.M && n < p.N) {
              float s = silu_f(v);
              if (p.out_bf16) s = bf16_round(s);
              float h = s * other;
              const long long o = (long long)m * p.ldo + (n >> 1);
              if (p.out_bf16)
                reinterpret_cast<bf16*>(p.out)[o] = __float2bfloat16_rn(h);
              else
                reinterpret_cast<float*>(p.out)[o] = h;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m_base + c + j;
            if (m < p.M && n < p.N) {
              float v = apply_epilogue_scalar(acc[j], n, m, p);
              const long long o = (long long)m * p.ldo + n;
              if (p.out_bf16)
                reinterpret_cast<bf16*>(p.out)[o] = __float2bfloat16_rn(v);
              else
                reinterpret_cast<float*>(p.out)[o] = v;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor-map cache + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  long long rows, cols, ld;
  int box_rows, elem_bytes;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows &&
           elem_bytes == o.elem_bytes;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix((size_t)k.rows);
    mix((size_t)k.cols);
    mix((size_t)k.ld);
    mix((size_t)k.box_rows);
    mix((size_t)k.elem_bytes);
    return h;
  }
};

std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

int get_tmap(const void* ptr, long long rows, long long cols, long long ld, int box_rows, int elem_bytes,
             CUtensorMap* out) {
  TmapKey key{ptr, rows, cols, ld, box_rows, elem_bytes};
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      return MTTS_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return mtts_set_error(MTTS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(kSwizzleBytes / elem_bytes), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return mtts_set_error(MTTS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%d", (int)r,
                          rows, cols, ld, box_rows);
  {
    std::lock_guard<std::mutex> g(g_tmap_mu);
    if (g_tmap_cache.size() > 65536) g_tmap_cache.clear();
    g_tmap_cache[key] = m;
  }
  *out = m;
  return MTTS_OK;
}

template <typename T, int BN, int kStages>
int launch(const CUtensorMap& tw, const CUtensorMap& tx, const GemmParams& p, dim3 grid, cudaStream_t stream) {
  constexpr int smem = smem_bytes<BN, kStages>();
  gemm_tc_kernel<T, BN, kStages><<<grid, kNumThreads, smem, stream>>>(tw, tx, p);
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}

template <typename T>
int dispatch(int bn, const CUtensorMap& tw, const CUtensorMap& tx, const GemmParams& p, dim3 grid,
             cudaStream_t stream) {
  switch (bn) {
    case 16: return launch<T, 16, 8>(tw, tx, p, grid, stream);
    case 32: return launch<T, 32, 8>(tw, tx, p, grid, stream);
    case 64: return launch<T, 64, 6>(tw, tx, p, grid, stream);
    case 128: return launch<T, 128, 3>(tw, tx, p, grid, stream);
    default: return launch<T, 256, 4>(tw, tx, p, grid, stream);
  }
}

template <typename T, int BN, int kStages>
int configure_one() {
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<T, BN, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       smem_bytes<BN, kStages>()));
  return MTTS_OK;
}

}  // namespace

// Opt every instantiation into its dynamic shared-memory size (called from mtts_init, outside any graph capture).
int mtts_configure_gemm_tc() {
  int rc = 0;
  if ((rc = configure_one<bf16, 16, 8>())) return rc;
  if ((rc = configure_one<bf16, 32, 8>())) return rc;
  if ((rc = configure_one<bf16, 64, 6>())) return rc;
  if ((rc = configure_one<bf16, 128, 3>())) return rc;
  if ((rc = configure_one<bf16, 256, 4>())) return rc;
  if ((rc = configure_one<float, 16, 8>())) return rc;
  if ((rc = configure_one<float, 32, 8>())) return rc;
  if ((rc = configure_one<float, 64, 6>())) return rc;
  if ((rc = configure_one<float, 128, 3>())) return rc;
  if ((rc = configure_one<float, 256, 4>())) return rc;
  return MTTS_OK;
}

int mtts_gemm_tc_pick_bn(int M) {
  if (M <= 16) return 16;
  if (M <= 32) return 32;
  if (M <= 64) return 64;
  if (M <= 128) return 128;
  return 256;
}

// Split-K heuristic: about one CTA per SM (the small-BN variants keep ~145 KB of TMA loads in flight per CTA,
// which is what saturates HBM), each split keeping at least 4 k-blocks so the ring still pipelines.
static int pick_splits(int tiles, int kb_total, int bn) {
  if (bn > 64) return 1;
  int target = mtts_num_sms();
  int s = target / tiles;
  if (s < 1) s = 1;
  int max_by_k = kb_total / 4;
  if (max_by_k < 1) max_by_k = 1;
  if (s > max_by_k) s = max_by_k;
  if (s > 16) s = 16;
  return s;
}

// Workspace layout: [kCounterBytes of per-tile arrival counters | split-K partial sums].
// The counter area must be zero before the first launch (the caller allocates it zeroed once); every
// launch leaves it zero again, so one workspace can be shared by all GEMMs issued on one stream.
static constexpr size_t kCounterBytes = 16384;

extern "C" size_t mtts_gemm_workspace_bytes(int M, int N, int K, int dtype) {
  int bn = mtts_gemm_tc_pick_bn(M);
  long long tiles = (long long)ceil_div(N, kBlockW) * ceil_div(M, bn);
  int bk = dtype == MTTS_DTYPE_BF16 ? 64 : 32;
  int splits = pick_splits((int)tiles, ceil_div(K, bk), bn);
  size_t ws = splits > 1 ? (size_t)splits * tiles * bn * kBlockW * sizeof(float) : 0;
  return kCounterBytes + ws;
}

extern "C" int mtts_gemm(const void* x, long long ldx, const void* w, long long ldw, void* out, long long ldo,
                         int M, int N, int K, int in_dtype, int out_dtype, int flags, const float* bias,
                         const float* gamma, const void* residual, long long ldr, void* workspace,
                         size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(M > 0 && N > 0 && K > 0, "mtts_gemm: empty problem M=%d N=%d K=%d", M, N, K);
  MTTS_REQUIRE(in_dtype == MTTS_DTYPE_BF16 || in_dtype == MTTS_DTYPE_F32, "mtts_gemm: bad in_dtype %d", in_dtype);
  MTTS_REQUIRE(out_dtype == MTTS_DTYPE_BF16 || out_dtype == MTTS_DTYPE_F32, "mtts_gemm: bad out_dtype %d", out_dtype);
  const int eb = in_dtype == MTTS_DTYPE_BF16 ? 2 : 4;
  MTTS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
               "mtts_gemm: x and w must be 16-byte aligned");
  MTTS_REQUIRE((ldx * eb) % 16 == 0 && (ldw * eb) % 16 == 0, "mtts_gemm: row strides must be multiples of 16 bytes");
  MTTS_REQUIRE(ldx >= K && ldw >= K, "mtts_gemm: leading dimensions smaller than K");
  if (flags & MTTS_EPI_BIAS) MTTS_REQUIRE(bias != nullptr, "mtts_gemm: EPI_BIAS without bias");
  if (flags & MTTS_EPI_GAMMA) MTTS_REQUIRE(gamma != nullptr, "mtts_gemm: EPI_GAMMA without gamma");
  if (flags & MTTS_EPI_RESIDUAL) MTTS_REQUIRE(residual != nullptr, "mtts_gemm: EPI_RESIDUAL without residual");
  if (flags & MTTS_EPI_SWIGLU)
    MTTS_REQUIRE((N % 2) == 0 && !(flags & ~MTTS_EPI_SWIGLU), "mtts_gemm: SWIGLU needs even N and no other flags");

  const int bn = mtts_gemm_tc_pick_bn(M);
  const int bk = kSwizzleBytes / eb;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.kb_total = ceil_div(K, bk);
  const int tiles_n = ceil_div(N, kBlockW), tiles_m = ceil_div(M, bn);
  const int tiles = tiles_n * tiles_m;
  int splits = pick_splits(tiles, p.kb_total, bn);
  p.kb_per_split = ceil_div(p.kb_total, splits);
  splits = ceil_div(p.kb_total, p.kb_per_split);  // every slice gets >= 1 k-block
  p.splits = splits;
  p.out = out; p.ldo = ldo; p.out_bf16 = out_dtype == MTTS_DTYPE_BF16;
  p.bias = bias; p.gamma = gamma; p.residual = residual; p.ldr = ldr; p.flags = flags;
  if (splits > 1 && (size_t)tiles * sizeof(int) > kCounterBytes) {
    splits = 1;
    p.splits = 1;
    p.kb_per_split = p.kb_total;
  }
  if (splits > 1) {
    const size_t ws_need = kCounterBytes + (size_t)splits * tiles * bn * kBlockW * sizeof(float);
    MTTS_REQUIRE(workspace != nullptr && workspace_bytes >= ws_need,
                 "mtts_gemm: workspace too small (%zu given, %zu needed)", workspace_bytes, ws_need);
    p.counters = reinterpret_cast<int*>(workspace);
    p.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kCounterBytes);
  }
  CUtensorMap tw, tx;
  int rc = get_tmap(w, N, K, ldw, kBlockW, eb, &tw);
  if (rc) return rc;
  rc = get_tmap(x, M, K, ldx, bn, eb, &tx);
  if (rc) return rc;
  dim3 grid(tiles_n, tiles_m, splits);
  if (in_dtype == MTTS_DTYPE_BF16) return dispatch<bf16>(bn, tw, tx, p, grid, stream);
  return dispatch<float>(bn, tw, tx, p, grid, stream);
}
