// Non-causal multi-head attention over padded batches on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), head_dim 64,
// fp16 operands: the VarLenAttention core of the codec's transformer stacks (XY_Tokenizer/xy_tokenizer/nn/modules.py:117-160,
// called per layer from the post-RVQ adapter and decoder stacks, model.py:103-128) on the fp16-operand decode path.
//
//   qkv [B*T, 3*H*64] fp16 (q | k | v incl. biases, q not scaled), out [B*T, H*64] fp16; keys >= lengths[b] are masked.
//
// B200 design. One CTA (128 threads) owns 128 query rows of one (item, head) and walks the keys in tiles of 64:
//   S  = Q K^T      tcgen05.mma M=128 N=64 K=64, Q and the K tile K-major in shared memory (TMA, 128B swizzle), S in TMEM
//   P  = softmax    thread = query row (tcgen05.ld 32x32b: a row's 64 scores arrive in one thread's registers, so the row
//                   maximum and sum need no shuffles); exp2 with the scale folded in; P written to shared memory as the
//                   fp16 K-major A operand of the second MMA (manual 128B swizzle)
//   O += P V        tcgen05.mma M=128 N=64 K=64, the V tile [keys][dims] is the MN-major B operand exactly as TMA stored
//                   it; O stays in TMEM for the whole key loop and is rescaled in place (tcgen05.ld / st) only when some
//                   row of the warp saw a new maximum
// K and V live in separate two-stage rings (a K tile is free once S is computed, a V tile once O += P V has retired);
// the next S is issued as soon as every thread has read the current one, so it runs under the softmax arithmetic. 64 KB
// of shared memory and 128 TMEM columns per CTA: three CTAs per SM, which is what overlaps one CTA's exponentials
// (MUFU: 16 per clock and SM — the real bound of head_dim-64 attention) with another's MMAs and loads.
// The mma.sync kernel this replaces (mha_varlen_h_kernel, codec_ops.cu) ran at 141 TFLOP/s.
#include "common.cuh"
#include "sm100.cuh"
#include "mtts_internal.h"

#include <cuda_fp16.h>

using namespace sm100;

namespace {

constexpr int kQ = 128;   // query rows per CTA
constexpr int kK = 64;    // keys per tile
constexpr int kD = 64;    // head dim
constexpr uint32_t kTileBytes = kK * kD * 2;  // 8 KB: one K or V tile; Q is two of them
constexpr uint32_t kSmemBytes = 2 * kTileBytes /*Q*/ + 2 * kTileBytes /*K ring*/ + 2 * kTileBytes /*V ring*/ + 2 * kTileBytes /*P*/ +
                                1024 /*align*/ + 128 /*barriers*/;  // barriers: 11 x 8 B + the TMEM pointer

struct MhaParams {
  __half* out;
  const int* lengths;
  int T, H, B;
  float scale_log2;
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major operand, 128-byte swizzle, 64 elements (128 B) wide: rows of the other dimension (keys) are 128 B apart, groups
// of 8 keys 1024 B apart (SBO); the leading-dimension offset between 64-element column blocks is unused here.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(kTileBytes >> 4) << 16;  // LBO (one tile; not reached with N = 64)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;        // SBO
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;                // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Warp roles: warps 0..3 = softmax (thread = query row = TMEM lane), warp 4 = control (one thread issues every TMA copy and
// every MMA). The two sides talk through mbarriers only — s_full / pv_done (MMA commit -> softmax), s_free / p_ready (one
// arrival per softmax warp -> control) — so no softmax thread ever waits for the issue of a copy or an MMA, and there is no
// block-wide barrier in the key loop. (First form: thread 0 issued everything between two __syncthreads per tile; ncu put
// 36 % of all stall samples in those two barriers, 4400 cycles per key tile and CTA.)
constexpr int kThreadsMha = 160;

__global__ void __launch_bounds__(kThreadsMha, 3) mha_varlen_tc5_kernel(const __grid_constant__ CUtensorMap tmap, const MhaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                        // [128 q][64 d] fp16, K-major, swizzled (two 64-row TMA boxes)
  uint8_t* sK = sQ + 2 * kTileBytes;         // [2][64 keys][64 d]
  uint8_t* sV = sK + 2 * kTileBytes;         // [2][64 keys][64 d]
  uint8_t* sP = sV + 2 * kTileBytes;         // [128 q][64 keys] fp16, K-major, swizzled
  uint64_t* q_full = reinterpret_cast<uint64_t*>(sP + 2 * kTileBytes);
  uint64_t* k_full = q_full + 1;             // [2]
  uint64_t* v_full = k_full + 2;             // [2]
  uint64_t* s_full = v_full + 2;
  uint64_t* pv_done = s_full + 1;
  uint64_t* s_free = pv_done + 1;            // 4 arrivals: every softmax warp holds its scores of the current tile
  uint64_t* p_ready = s_free + 1;            // 4 arrivals: every softmax warp has written its rows of P (and rescaled O)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(p_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQ;
  const int E = p.H * kD;
  const int len = p.lengths ? min(p.lengths[b], p.T) : p.T;
  const int kv_len = len > 0 ? len : p.T;  // all-masked item: uniform over every key, as the reference computes it
  const int n_tiles = (kv_len + kK - 1) / kK;
  const int row_base = b * p.T;            // first row of this item in the [B*T, 3E] matrix

  if (tid == 0) {
    prefetch_tmap(&tmap);
    mbar_init(q_full, 1);
    mbar_init(&k_full[0], 1);
    mbar_init(&k_full[1], 1);
    mbar_init(&v_full[0], 1);
    mbar_init(&v_full[1], 1);
    mbar_init(s_full, 1);
    mbar_init(pv_done, 1);
    mbar_init(s_free, 4);
    mbar_init(p_ready, 4);
    fence_barrier_init();
  }
  if (warp == 4) {
    __syncwarp();
    tmem_alloc<128>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();  // qkv is the predecessor's output

  constexpr uint32_t kIdescS = make_idesc(0, kQ, kK);               // f16 x f16 -> f32, both K-major
  constexpr uint32_t kIdescO = make_idesc(0, kQ, kD) | (1u << 16);  // B (the V tile) MN-major

  if (warp == 4) {
    // ================= control: TMA copies + MMA issue =================
    if (lane == 0) {
      const int cq = h * kD, ck = E + h * kD, cv = 2 * E + h * kD;
      auto load_k = [&](int t) {
        mbar_arrive_expect_tx(&k_full[t & 1], kTileBytes);
        tma_load_2d(sK + (t & 1) * kTileBytes, &tmap, &k_full[t & 1], ck, row_base + t * kK, kEvictLast);
      };
      auto load_v = [&](int t) {
        mbar_arrive_expect_tx(&v_full[t & 1], kTileBytes);
        tma_load_2d(sV + (t & 1) * kTileBytes, &tmap, &v_full[t & 1], cv, row_base + t * kK, kEvictLast);
      };
      auto issue_s = [&](int t) {
        mbar_wait(&k_full[t & 1], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + (t & 1) * kTileBytes);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16(tmem_base, make_smem_desc_sw128(qa + k * 32), make_smem_desc_sw128(ka + k * 32), kIdescS, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      tma_load_2d(sQ, &tmap, q_full, cq, row_base + q0, kEvictNormal);
      tma_load_2d(sQ + kTileBytes, &tmap, q_full, cq, row_base + q0 + 64, kEvictNormal);
      for (int t = 0; t < 2 && t < n_tiles; ++t) {
        load_k(t);
        load_v(t);
      }
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(s_free, j & 1);  // S_j is in registers: its TMEM columns and K stage j % 2 are free
        tc_fence_after();
        if (j + 2 < n_tiles) load_k(j + 2);
        if (j + 1 < n_tiles) issue_s(j + 1);  // runs under tile j's softmax
        mbar_wait(p_ready, j & 1);
        mbar_wait(&v_full[j & 1], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV + (j & 1) * kTileBytes);
#pragma unroll
        for (int k = 0; k < kK / 16; ++k)  // 16 keys per MMA: +32 B along P's rows, +16 rows (2 KB) down the V tile
          umma_bf16(tmem_base + 64, make_smem_desc_sw128(pa + k * 32), make_smem_desc_mn_sw128(va + k * 2048), kIdescO,
                    (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(pv_done);
        if (j + 2 < n_tiles) {  // V stage j % 2 is free once this O += P V has retired
          mbar_wait(pv_done, j & 1);
          load_v(j + 2);
        }
      }
    }
  } else {
    // ================= softmax: thread = query row =================
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);  // this warp's lanes, columns 0..63: S
    const uint32_t tO = tS + 64;                                               // columns 64..127: O
    float m_run = -INFINITY, l_run = 0.f;
    const float c = p.scale_log2;
    const int row = warp * 32 + lane;
    uint8_t* p_row = sP + row * 128;
    const int sw = row & 7;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32b_x32(tS, s0);
      tmem_ld_32x32b_x32(tS + 32, s1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      const int valid = kv_len - j * kK;  // keys of this tile inside the item (>= 1)
      float mx = -INFINITY;
      if (valid < kK) {  // uniform: only the item's last key tile is masked (ncu: the softmax warps are issue-bound, 62 %)
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= valid) s0[i] = __float_as_uint(-INFINITY);
          if (32 + i >= valid) s1[i] = __float_as_uint(-INFINITY);
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s1[i])));
      const float m_new = fmaxf(m_run, mx);
      // LAZY rescaling: the reference maximum of a row moves only when the true maximum has outgrown it by more than 2^8
      // (in the exp2 domain) somewhere in the warp; until then probabilities may exceed 1 (<= 256: harmless in fp16 / fp32)
      // and O, l keep their scale — the result is the same quotient. Only rows of this item vote: a row's arithmetic must
      // not depend on the item that follows it in the batch.
      const bool move = j == 0 || __any_sync(0xffffffffu, q0 + row < p.T && (m_new - m_run) * c > 8.0f);
      const float m_use = move ? m_new : m_run;
      const float alpha = ex2((m_run - m_use) * c);  // 0 on the first tile (m_run = -inf), 1 when nothing moves
      const float mc = m_use * c;
      float sum = 0.f;
      uint32_t ph[32];  // 64 probabilities as fp16 pairs
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a0 = ex2(fmaf(__uint_as_float(s0[2 * i]), c, -mc)), a1 = ex2(fmaf(__uint_as_float(s0[2 * i + 1]), c, -mc));
        const float b0 = ex2(fmaf(__uint_as_float(s1[2 * i]), c, -mc)), b1 = ex2(fmaf(__uint_as_float(s1[2 * i + 1]), c, -mc));
        const __half2 ha = __floats2half2_rn(a0, a1), hb = __floats2half2_rn(b0, b1);
        // the sum runs over the ROUNDED probabilities, the values the second MMA multiplies with V
        sum += (__low2float(ha) + __high2float(ha)) + (__low2float(hb) + __high2float(hb));
        ph[i] = *reinterpret_cast<const uint32_t*>(&ha);
        ph[16 + i] = *reinterpret_cast<const uint32_t*>(&hb);
      }
      l_run = l_run * alpha + sum;
      m_run = m_use;
      if (j > 0) {  // O += P V of the previous tile has retired: P and O are ours
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
        if (move) {  // warp-uniform: rescale this warp's 32 rows of O in place
          uint32_t o[32];
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {
            tmem_ld_32x32b_x32(tO + half * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(tO + half * 32, o);
          }
          tmem_st_wait();
        }
      }
      // P -> shared memory, K-major with the 128-byte swizzle the MMA descriptor expects (16-byte chunk ^= row % 8)
#pragma unroll
      for (int ch = 0; ch < 8; ++ch)
        *reinterpret_cast<uint4*>(p_row + ((ch ^ sw) << 4)) = make_uint4(ph[4 * ch], ph[4 * ch + 1], ph[4 * ch + 2], ph[4 * ch + 3]);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // ---- O / l -> out
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const int q = q0 + row;
    __half* op = p.out + ((long long)(row_base + q)) * E + h * kD;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO + half * 32, o);
      tmem_ld_wait();
      if (q < p.T) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __half2 hh = __floats2half2_rn(__uint_as_float(o[v4 * 8 + 2 * i]) * inv, __uint_as_float(o[v4 * 8 + 2 * i + 1]) * inv);
            w[i] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          *reinterpret_cast<uint4*>(op + half * 32 + v4 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

}  // namespace

int mtts_configure_mha_tc5() {
  MTTS_CUDA_CHECK(cudaFuncSetAttribute(mha_varlen_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  return MTTS_OK;
}

extern "C" int mtts_mha_varlen_tc(const void* qkv_f16, void* out_f16, const int* lengths, int B, int T, int num_heads, int head_dim,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(head_dim == kD, "mtts_mha_varlen_tc: head_dim must be 64 (got %d)", head_dim);
  if (B <= 0 || T <= 0) return MTTS_OK;
  MTTS_REQUIRE(qkv_f16 && out_f16, "mtts_mha_varlen_tc: null pointer");
  MTTS_REQUIRE((reinterpret_cast<uintptr_t>(qkv_f16) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_f16) & 15) == 0,
               "mtts_mha_varlen_tc: qkv and out must be 16-byte aligned");
  const int E = num_heads * kD;
  CUtensorMap tm;
  int rc = mtts_get_tmap_2d(qkv_f16, (long long)B * T, 3LL * E, 3LL * E, kK, -2, &tm);
  if (rc) return rc;
  MhaParams p;
  p.out = reinterpret_cast<__half*>(out_f16);
  p.lengths = lengths;
  p.T = T; p.H = num_heads; p.B = B;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
  dim3 grid((T + kQ - 1) / kQ, num_heads, B);
  MTTS_CUDA_CHECK(mtts_launch(mha_varlen_tc5_kernel, grid, dim3(kThreadsMha), kSmemBytes, stream, tm, p));
  MTTS_LAUNCH_CHECK();
  return MTTS_OK;
}
