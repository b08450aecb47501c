// Mixture-of-experts hot path for sm_100a: peer-memory all-to-all dispatch fused with fp8 quantisation, fp8 (e4m3)
// grouped expert GEMM on tcgen05 (kind::f8f6f4) with per-token / per-expert scales applied in the epilogue, and a
// peer-memory combine that pulls expert outputs back and applies the gate weights.
//
// The reference implements expert parallelism as einsum dispatch/combine + a hand-rolled ncclSend/ncclRecv all-to-all
// (DAPPLEAllToAllThunk; examples/gpt_moe/layers/moe_layers.py:425-446).  Here a token row travels exactly
// once over NVLink, already quantised, straight into the owning expert's input buffer:
//   dispatch : x[t, :] (bf16) --row amax--> e4m3 + fp32 scale --P2P st--> rank(e).xin[e_local, src*C + slot, :]
//   expert FC: h = gelu((xin . w1^T) * row_scale * w_scale + b1)      tcgen05 fp8, TMEM accumulators, bf16 out
//   combine  : out[t, :] = sum_k gate[t,k] * rank(e_k).y[e_local, src*C + slot, :]        P2P loads
// Slots are assigned per (expert, source rank) region, so ranks never contend for a slot (GShard local groups).
#include "sm100_ptx.cuh"
#include <cuda_fp8.h>
#include <stdio.h>

using namespace sm100;

namespace {

constexpr int MAXP = 8;
struct Peers {
  void* p[MAXP];
};

// ------------------------------------------------------------------ dispatch (quantise + all-to-all push)
// route[t*K + k] = e * 65536 + slot, or -1 when the token was dropped for that choice.
__global__ void __launch_bounds__(256) moe_dispatch_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ route, Peers xin,
                                                          Peers xscale, int T, int M, int top_k, int experts_per_rank, int cap_per_src,
                                                          int n_ranks, int src_rank) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= T * top_k) return;
  const int r = route[warp];
  if (r < 0) return;
  const int t = warp / top_k, e = r >> 16, slot = r & 0xFFFF;
  const int owner = e / experts_per_rank, el = e % experts_per_rank;
  const __nv_bfloat16* row = x + (size_t)t * M;
  // pass 1: row amax (M <= 8192: each lane keeps its values in registers, 8 per 16-byte vector)
  float amax = 0.f;
  for (int v = lane; v < (M >> 3); v += 32) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(row) + v);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  const float scale = amax > 0.f ? amax / 448.f : 1.f, inv = 1.f / scale;
  const size_t drow = ((size_t)el * n_ranks + src_rank) * cap_per_src + slot;
  uint8_t* dst = reinterpret_cast<uint8_t*>(xin.p[owner]) + drow * M;
  for (int v = lane; v < (M >> 3); v += 32) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(row) + v);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    uint2 q;
    uint8_t* qb = reinterpret_cast<uint8_t*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      qb[2 * i] = (uint8_t)__nv_cvt_float_to_fp8(f.x * inv, __NV_SATFINITE, __NV_E4M3);
      qb[2 * i + 1] = (uint8_t)__nv_cvt_float_to_fp8(f.y * inv, __NV_SATFINITE, __NV_E4M3);
    }
    reinterpret_cast<uint2*>(dst)[v] = q;  // 8 fp8 values per store, straight into the expert owner's memory
  }
  if (lane == 0) reinterpret_cast<float*>(xscale.p[owner])[drow] = scale;
}

// ------------------------------------------------------------------ combine (pull + gate)
__global__ void __launch_bounds__(256) moe_combine_kernel(Peers y, const int* __restrict__ route, const float* __restrict__ gate,
                                                         __nv_bfloat16* __restrict__ out, int T, int M, int top_k, int experts_per_rank,
                                                         int cap_per_src, int n_ranks, int src_rank) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= T) return;
  for (int v = lane; v < (M >> 3); v += 32) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < top_k; ++k) {
      const int r = route[warp * top_k + k];
      if (r < 0) continue;
      const int e = r >> 16, slot = r & 0xFFFF;
      const int owner = e / experts_per_rank, el = e % experts_per_rank;
      const size_t srow = ((size_t)el * n_ranks + src_rank) * cap_per_src + slot;
      const uint4 u = *(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(y.p[owner]) + srow * M) + v);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
      const float g = gate[warp * top_k + k];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        acc[2 * i] += g * f.x;
        acc[2 * i + 1] += g * f.y;
      }
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    reinterpret_cast<uint4*>(out + (size_t)warp * M)[v] = o;
  }
}

// ------------------------------------------------------------------ fp8 grouped GEMM (tcgen05 kind::f8f6f4)
constexpr int BM = 128, BN = 128, BK = 128 /*bytes == e4m3 elements*/, UK = 32, STAGES = 6, THREADS = 192;
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

struct Fp8Params {
  int M, N, K, batch, m_blocks, n_blocks, k_blocks;
  long long ldd, stride_d;
  void* D;                 // bf16 [batch, M, N]
  const float* row_scale;  // [batch, M]   per-token activation scale
  const float* w_scale;    // [batch]      per-expert weight scale
  const float* bias;       // [batch, N] or nullptr
  int act;                 // 1: tanh-GELU
};

__device__ __forceinline__ float gelu_t(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k0 * (x + k1 * x * x * x)));
  return 0.5f * x * (1.0f + t);
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_fp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Fp8Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 2 * BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total = p.m_blocks * p.n_blocks * p.batch;
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int t = tile;
        const int m_blk = t % p.m_blocks; t /= p.m_blocks;
        const int n_blk = t % p.n_blocks; const int b = t / p.n_blocks;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_3d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM, b);
          tma_load_3d(sa + A_BYTES, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN, b);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(UMMA_E4M3, UMMA_E4M3, BM, BN, 0, 0);
      int stage = 0, acc = 0; uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            umma_f8f6f4_ss(tmem_base + acc * BN, make_smem_desc_sw128(sa + k * UK, 0, 1024), make_smem_desc_sw128(sb + k * UK, 0, 1024),
                           idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int quarter = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      int t = tile;
      const int m_blk = t % p.m_blocks; t /= p.m_blocks;
      const int n_blk = t % p.n_blocks; const int b = t / p.n_blocks;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const float rs = (row_ok ? p.row_scale[(size_t)b * p.M + row] : 0.f) * p.w_scale[b];
      const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row_ok && col0 < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * rs;
          if (p.bias) {
            const float* bp = p.bias + (size_t)b * p.N + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __ldg(bp + j);
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_t(v[j]);
          }
          uint4* dp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.D) + (long long)b * p.stride_d + (long long)row * p.ldd + col0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (col0 + q * 8 < p.N) {
              uint4 u;
              u.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]); u.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
              u.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]); u.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
              dp[q] = u;
            }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * BN); }
}

// per-tensor (per-expert) weight quantisation: w[b] (bf16 [rows, K]) -> e4m3 + scale[b]
__global__ void quant_weight_kernel(const __nv_bfloat16* __restrict__ w, uint8_t* __restrict__ q, const float* __restrict__ scale_inv,
                                    long long per_batch, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    q[i] = (uint8_t)__nv_cvt_float_to_fp8(__bfloat162float(w[i]) * scale_inv[i / per_batch], __NV_SATFINITE, __NV_E4M3);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_tmap_u8(CUtensorMap* out, const void* ptr, long long inner, long long rows, long long batch, int box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)inner, (cuuint64_t)(inner * rows)};
  cuuint32_t box[3] = {128, (cuuint32_t)box_rows, 1}, es[3] = {1, 1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : -2;
}
Peers mk(void* const* ptrs, int n) {
  Peers p;
  for (int i = 0; i < MAXP; ++i) p.p[i] = i < n ? ptrs[i] : nullptr;
  return p;
}

}  // namespace

#define CS(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int tepd_moe_dispatch(const void* x, const void* route, void* const* xin_ptrs, void* const* xscale_ptrs, int T, int M,
                                 int top_k, int experts_per_rank, int cap_per_src, int n_ranks, int src_rank, void* stream) {
  if (M % 8 || n_ranks > MAXP) return -2;
  const int warps = T * top_k;
  moe_dispatch_kernel<<<(warps * 32 + 255) / 256, 256, 0, CS(stream)>>>((const __nv_bfloat16*)x, (const int*)route, mk(xin_ptrs, n_ranks),
                                                                      mk(xscale_ptrs, n_ranks), T, M, top_k, experts_per_rank, cap_per_src,
                                                                      n_ranks, src_rank);
  return (int)cudaGetLastError();
}
extern "C" int tepd_moe_combine(void* const* y_ptrs, const void* route, const void* gate, void* out, int T, int M, int top_k,
                                int experts_per_rank, int cap_per_src, int n_ranks, int src_rank, void* stream) {
  if (M % 8 || n_ranks > MAXP) return -2;
  moe_combine_kernel<<<(T * 32 + 255) / 256, 256, 0, CS(stream)>>>(mk(y_ptrs, n_ranks), (const int*)route, (const float*)gate,
                                                                  (__nv_bfloat16*)out, T, M, top_k, experts_per_rank, cap_per_src, n_ranks,
                                                                  src_rank);
  return (int)cudaGetLastError();
}
extern "C" int tepd_quant_weight_fp8(const void* w, void* q, const void* scale_inv, long long per_batch, long long total, void* stream) {
  quant_weight_kernel<<<148 * 4, 256, 0, CS(stream)>>>((const __nv_bfloat16*)w, (uint8_t*)q, (const float*)scale_inv, per_batch, total);
  return (int)cudaGetLastError();
}
// D[b] (bf16 [M,N]) = act((A[b] (e4m3 [M,K]) . B[b]^T (e4m3 [N,K])) * row_scale[b,:] * w_scale[b] + bias[b])
extern "C" int tepd_gemm_fp8(const void* A, const void* B, void* D, const void* row_scale, const void* w_scale, const void* bias, int M,
                             int N, int K, int batch, int act, int num_sms, void* stream) {
  if (K % 16 || N % 8) return -2;
  Fp8Params p;
  p.M = M; p.N = N; p.K = K; p.batch = batch;
  p.m_blocks = (M + BM - 1) / BM; p.n_blocks = (N + BN - 1) / BN; p.k_blocks = (K + BK - 1) / BK;
  p.ldd = N; p.stride_d = (long long)M * N; p.D = D;
  p.row_scale = (const float*)row_scale; p.w_scale = (const float*)w_scale; p.bias = (const float*)bias; p.act = act;
  CUtensorMap ta, tb;
  if (make_tmap_u8(&ta, A, K, M, batch, BM) || make_tmap_u8(&tb, B, K, N, batch, BN)) return -3;
  static bool cfg = false;
  if (!cfg) { if (cudaFuncSetAttribute(gemm_fp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -4; cfg = true; }
  const int total = p.m_blocks * p.n_blocks * batch;
  if (num_sms <= 0) num_sms = 148;
  gemm_fp8_kernel<<<total < num_sms ? total : num_sms, THREADS, SMEM_BYTES, CS(stream)>>>(ta, tb, p);
  return (int)cudaGetLastError();
}

// ================================================================================================ route-table dispatch / combine
// GShard writes dispatch and combine as dense einsums with a [G, S, E, C] one-hot mask (and so does the planner graph, which is
// what lets the planner see the G <-> E re-distribution).  Every slot (e, c) of a group holds at most ONE token and every token
// occupies at most top_k slots, so at run time the four dense einsums (and the four their gradients need) are row gathers:
//   gather_scale : out[e, g, c, :] = w[g, e, c] * src[g, slot_src[g, e, c], :]                 (dispatch; d combine / d y)
//   combine_sum  : out[g, s, :]    = sum_k gw[g, s, k] * y[e_k, g, c_k, :]                      (combine; d dispatch / d x)
//   route_dots   : dots[g, s, k]   = < a[g, s, :], b[e_k, g, c_k, :] >                          (d / d mask weights)
// 2 * G*S*E*C*M FLOPs per einsum (25.8 GFLOP per group at the reference shape) become G*E*C*M bytes of row traffic.
namespace {

__global__ void __launch_bounds__(256) moe_gather_scale_kernel(const __nv_bfloat16* __restrict__ src, const int* __restrict__ slot_src,
                                                              const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int G, int S,
                                                              int E, int C, int M) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= (long long)E * G * C) return;
  const int c = (int)(row % C), g = (int)((row / C) % G), e = (int)(row / ((long long)G * C));
  const long long si = ((long long)g * E + e) * C + c;
  const int s = slot_src[si];
  const float ws = s >= 0 ? w[si] : 0.f;
  uint4* dst = reinterpret_cast<uint4*>(out + row * M);
  const uint4* sp = reinterpret_cast<const uint4*>(src + ((long long)g * S + (s >= 0 ? s : 0)) * M);
  for (int v = lane; v < (M >> 3); v += 32) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (s >= 0) {
      const uint4 u = __ldg(sp + v);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
      float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]), f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
      o.x = pack_bf16x2(ws * f0.x, ws * f0.y); o.y = pack_bf16x2(ws * f1.x, ws * f1.y);
      o.z = pack_bf16x2(ws * f2.x, ws * f2.y); o.w = pack_bf16x2(ws * f3.x, ws * f3.y);
    }
    dst[v] = o;
  }
}

__global__ void __launch_bounds__(256) moe_combine_sum_kernel(const __nv_bfloat16* __restrict__ y, const int* __restrict__ re,
                                                             const int* __restrict__ rc, const float* __restrict__ gw,
                                                             __nv_bfloat16* __restrict__ out, int G, int S, int E, int C, int M, int K) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;   // (g, s)
  const int lane = threadIdx.x & 31;
  if (row >= (long long)G * S) return;
  const int g = (int)(row / S);
  for (int v = lane; v < (M >> 3); v += 32) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < K; ++k) {
      const int e = re[row * K + k];
      if (e < 0) continue;
      const int c = rc[row * K + k];
      const float wk = gw[row * K + k];
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(y + (((long long)e * G + g) * C + c) * M) + v);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        acc[2 * i] += wk * f.x;
        acc[2 * i + 1] += wk * f.y;
      }
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    reinterpret_cast<uint4*>(out + row * M)[v] = o;
  }
}

__global__ void __launch_bounds__(256) moe_route_dots_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                            const int* __restrict__ re, const int* __restrict__ rc,
                                                            float* __restrict__ dots, int G, int S, int E, int C, int M, int K) {
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;   // (g, s, k)
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)G * S * K) return;
  const long long row = wid / K;
  const int g = (int)(row / S);
  const int e = re[wid];
  float acc = 0.f;
  if (e >= 0) {
    const int c = rc[wid];
    const uint4* ap = reinterpret_cast<const uint4*>(a + row * M);
    const uint4* bp = reinterpret_cast<const uint4*>(b + (((long long)e * G + g) * C + c) * M);
    for (int v = lane; v < (M >> 3); v += 32) {
      const uint4 ua = __ldg(ap + v), ub = __ldg(bp + v);
      const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&ua);
      const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&ub);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 fa = __bfloat1622float2(ha[i]), fb = __bfloat1622float2(hb[i]);
        acc += fa.x * fb.x + fa.y * fb.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) dots[wid] = acc;
}

}  // namespace

extern "C" int tepd_moe_gather_scale(const void* src, const void* slot_src, const void* w, void* out, int G, int S, int E, int C, int M,
                                     void* stream) {
  if (M % 8) return -2;
  const long long rows = (long long)E * G * C;
  moe_gather_scale_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, CS(stream)>>>((const __nv_bfloat16*)src, (const int*)slot_src,
                                                                                        (const float*)w, (__nv_bfloat16*)out, G, S, E, C, M);
  return (int)cudaGetLastError();
}
extern "C" int tepd_moe_combine_sum(const void* y, const void* re, const void* rc, const void* gw, void* out, int G, int S, int E, int C,
                                    int M, int K, void* stream) {
  if (M % 8) return -2;
  const long long rows = (long long)G * S;
  moe_combine_sum_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, CS(stream)>>>((const __nv_bfloat16*)y, (const int*)re, (const int*)rc,
                                                                                       (const float*)gw, (__nv_bfloat16*)out, G, S, E, C, M, K);
  return (int)cudaGetLastError();
}
extern "C" int tepd_moe_route_dots(const void* a, const void* b, const void* re, const void* rc, void* dots, int G, int S, int E, int C,
                                   int M, int K, void* stream) {
  if (M % 8) return -2;
  const long long warps = (long long)G * S * K;
  moe_route_dots_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, CS(stream)>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                                                       (const int*)re, (const int*)rc, (float*)dots, G, S, E, C, M, K);
  return (int)cudaGetLastError();
}
