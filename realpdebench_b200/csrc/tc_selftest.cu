// Hardware self-test of the tensor-core building blocks used by tc_layer.cu:
// one CTA computes D[128 x N] = A[128 x K] * B[N x K]^T with tcgen05.mma kind::tf32 and returns D,
// for every operand-staging variant the production kernels use (manual swizzled stores, TMA
// SWIZZLE_128B, K-major and MN-major descriptors, A operand in TMEM, TMA store of the result).
// tests/test_gpu_tc_primitives.py compares D with an fp32 reference; if a descriptor or swizzle
// assumption were wrong this is where it shows, not inside the fused kernels.
#include <stdio.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200fno {

int encode_tensor_map(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
      return B200FNO_ECUDA;
    }
    fn = (EncodeFn)p;
  }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i], es[i] = 1;
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(gptr), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return B200FNO_ECUDA;
  }
  return 0;
}

using namespace tc;

struct SelfTestArgs {
  const float *A, *B;
  float* D;
  int N, K, mode_a, mode_b, out_tma;
};

__global__ void __launch_bounds__(128) selftest_kernel(SelfTestArgs a, const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB,
                                                       const __grid_constant__ CUtensorMap tmD) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int N = a.N, K = a.K;
  uint8_t* sA = smem;                        // 128 x K fp32 = K*512 B
  uint8_t* sB = sA + (size_t)K * 512;        // N x K fp32
  uint8_t* sD = sB + (size_t)N * K * 4;      // staging 128 x N fp32 (N/32 sub-tiles of 16 KB)
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_A = tmem_base + 128;  // columns 128.. hold the A operand in TS mode

  // ---------------- stage operands ----------------
  uint32_t tx_bytes = 0;
  if (a.mode_a == 0) {  // manual K-major, 128B swizzle: sub-tile s = k/32, row r, chunk c = (k%32)/4
    for (int idx = tid; idx < 128 * K / 4; idx += 128) {
      int r = idx / (K / 4), kc = idx % (K / 4);
      float4 v = *reinterpret_cast<const float4*>(a.A + (size_t)r * K + kc * 4);
      int s = kc / 8, c = kc % 8;
      *reinterpret_cast<float4*>(sA + (size_t)s * 16384 + sw128_off(r, c)) = v;
    }
  } else if (a.mode_a == 3) {  // A in TMEM: thread = row
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t r[8];
      for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(a.A[(size_t)tid * K + k0 + j]);
      tmem_st8(tmem_A + ((uint32_t)(warp * 32) << 16) + k0, r);
    }
    tmem_st_wait();
  } else {
    tx_bytes += 128 * K * 4;
  }
  if (a.mode_b == 0) {
    for (int idx = tid; idx < N * K / 4; idx += 128) {
      int r = idx / (K / 4), kc = idx % (K / 4);
      float4 v = *reinterpret_cast<const float4*>(a.B + (size_t)r * K + kc * 4);
      int s = kc / 8, c = kc % 8;
      *reinterpret_cast<float4*>(sB + (size_t)s * N * 128 + sw128_off(r, c)) = v;
    }
  } else {
    tx_bytes += N * K * 4;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tx_bytes) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_tma, tx_bytes);
      if (a.mode_a == 1)  // K-major: global [128][K] -> sub-tiles of 32 k
        for (int s = 0; s < K / 32; ++s) tma_load_2d(sA + (size_t)s * 16384, &tmA, &bar_tma, 32 * s, 0);
      if (a.mode_a == 2)  // MN-major: global [K][128] -> 4 blocks of 32 m, each K rows x 128 B
        for (int mb = 0; mb < 4; ++mb) tma_load_2d(sA + (size_t)mb * K * 128, &tmA, &bar_tma, 32 * mb, 0);
      if (a.mode_b == 1)
        for (int s = 0; s < K / 32; ++s) tma_load_2d(sB + (size_t)s * N * 128, &tmB, &bar_tma, 32 * s, 0);
      if (a.mode_b == 2)
        for (int nb = 0; nb < N / 32; ++nb) tma_load_2d(sB + (size_t)nb * K * 128, &tmB, &bar_tma, 32 * nb, 0);
    }
    mbar_wait(&bar_tma, 0);
  }
  // ---------------- MMA ----------------
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, N, a.mode_a == 2, a.mode_b == 2);
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t bd;
      if (a.mode_b == 2) bd = make_smem_desc(smem_u32(sB) + ks * 1024, K * 128, 512, LAYOUT_SW128_BASE32B);
      else bd = make_smem_desc(smem_u32(sB) + (ks / 4) * N * 128 + (ks % 4) * 32, 0, 1024);
      if (a.mode_a == 3) {
        umma_tf32_ts(tmem_base, tmem_A + ks * 8, bd, idesc, ks > 0);
      } else {
        uint64_t ad;
        if (a.mode_a == 2) ad = make_smem_desc(smem_u32(sA) + ks * 1024, K * 128, 512, LAYOUT_SW128_BASE32B);
        else ad = make_smem_desc(smem_u32(sA) + (ks / 4) * 16384 + (ks % 4) * 32, 0, 1024);
        umma_tf32_ss(tmem_base, ad, bd, idesc, ks > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  // ---------------- epilogue ----------------
  const int row = tid;
  for (int n0 = 0; n0 < N; n0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + n0, r);
    tmem_ld_wait();
    if (a.out_tma) {
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(sD + (size_t)(n0 / 32) * 16384 + sw128_off(row, c)) =
            make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
    } else {
      for (int j = 0; j < 32; ++j) a.D[(size_t)row * N + n0 + j] = __uint_as_float(r[j]);
    }
  }
  if (a.out_tma) {
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      for (int s = 0; s < N / 32; ++s) tma_store_2d(&tmD, sD + (size_t)s * 16384, 32 * s, 0);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// MMA issue-rate probe: one CTA issues `iters` back-to-back tcgen05.mma (M=128, N, K=8, kind::tf32) into one
// accumulator and reports the elapsed SM cycles.  a_in_tmem: .ts form (A from TMEM) vs .ss (A from smem).
__global__ void __launch_bounds__(128) mma_rate_kernel(int N, int iters, int a_in_tmem, int nacc, int accumulate,
                                                       long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) mbar_init(&bar, 1), fence_barrier_init();
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const uint32_t tb = tmem_base_s;
    const uint64_t bd0 = make_smem_desc(sb, 0, 1024), ad0 = make_smem_desc(sa, 0, 1024);
    const long long t0 = clock64();
    if (elect_one_sync()) {
      if (nacc == 1) {
#pragma unroll 8
        for (int i = 0; i < iters; ++i) {
          if (a_in_tmem) umma_tf32_ts(tb, tb + 448 + (i & 7) * 8, bd0 + 2 * (i & 3), idesc, accumulate);
          else umma_tf32_ss(tb, ad0 + 2 * (i & 3), bd0 + 2 * (i & 3), idesc, accumulate);
        }
      } else {
#pragma unroll 8
        for (int i = 0; i < iters; ++i) {
          const uint32_t d = tb + (uint32_t)((i & (nacc - 1)) * N);
          if (a_in_tmem) umma_tf32_ts(d, tb + 448 + (i & 7) * 8, bd0 + 2 * (i & 3), idesc, accumulate);
          else umma_tf32_ss(d, ad0 + 2 * (i & 3), bd0 + 2 * (i & 3), idesc, accumulate);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (tid == 32) out_cycles[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

}  // namespace b200fno

using namespace b200fno;

extern "C" int b200fno_selftest_mma_rate(int32_t N, int32_t iters, int32_t a_in_tmem, int32_t nacc, int32_t accumulate,
                                         long long* out_cycles_dev, void* stream) {
  if (N % 16 || N < 16 || N > 256 || iters < 1 || !out_cycles_dev || nacc < 1 || nacc * N > 448) {
    set_error("mma_rate: bad argument");
    return B200FNO_EINVAL;
  }
  const int smem = 16384 + 32768 + 1024;
  B2_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_rate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(N, iters, a_in_tmem, nacc, accumulate, out_cycles_dev);
  B2_LAUNCHED("mma_rate_kernel");
  return 0;
}

extern "C" int b200fno_selftest_umma(int32_t mode_a, int32_t mode_b, int32_t out_tma, int32_t N, int32_t K,
                                     const float* A, const float* B, float* D, void* stream) {
  if (mode_a < 0 || mode_a > 3 || mode_b < 0 || mode_b > 2 || (N != 32 && N != 64 && N != 128) ||
      (K != 32 && K != 64 && K != 96 && K != 128) || !A || !B || !D) {
    set_error("selftest: bad argument");
    return B200FNO_EINVAL;
  }
  CUtensorMap tmA, tmB, tmD;
  memset(&tmA, 0, sizeof(tmA)), memset(&tmB, 0, sizeof(tmB)), memset(&tmD, 0, sizeof(tmD));
  if (mode_a == 1 || mode_a == 0 || mode_a == 3) {  // A global [128][K]
    uint64_t d[2] = {(uint64_t)K, 128}, s[1] = {(uint64_t)K * 4};
    uint32_t b[2] = {32, 128};
    B2_TRY(encode_tensor_map(&tmA, A, 2, d, s, b, 1));
  } else {  // A global [K][128]
    uint64_t d[2] = {128, (uint64_t)K}, s[1] = {128 * 4};
    uint32_t b[2] = {32, (uint32_t)K};
    B2_TRY(encode_tensor_map(&tmA, A, 2, d, s, b, 2));
  }
  if (mode_b != 2) {  // B global [N][K]
    uint64_t d[2] = {(uint64_t)K, (uint64_t)N}, s[1] = {(uint64_t)K * 4};
    uint32_t b[2] = {32, (uint32_t)N};
    B2_TRY(encode_tensor_map(&tmB, B, 2, d, s, b, 1));
  } else {  // B global [K][N]
    uint64_t d[2] = {(uint64_t)N, (uint64_t)K}, s[1] = {(uint64_t)N * 4};
    uint32_t b[2] = {32, (uint32_t)K};
    B2_TRY(encode_tensor_map(&tmB, B, 2, d, s, b, 2));
  }
  {
    uint64_t d[2] = {(uint64_t)N, 128}, s[1] = {(uint64_t)N * 4};
    uint32_t b[2] = {32, 128};
    B2_TRY(encode_tensor_map(&tmD, D, 2, d, s, b, 1));
  }
  SelfTestArgs a{A, B, D, N, K, mode_a, mode_b, out_tma};
  size_t smem = (size_t)K * 512 + (size_t)N * K * 4 + (size_t)128 * N * 4 + 1024;
  B2_CUDA(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, tmA, tmB, tmD);
  B2_LAUNCHED("selftest_kernel");
  return 0;
}
