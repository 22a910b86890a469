// fp32 FFMA (SIMT) stage kernels.  Every stage of the truncated-DFT pipeline is a
// small dense GEMM whose contiguous dimension is the channel axis; all of them
// share one register-tiled micro-kernel (4x4 outputs per thread, operands staged
// in shared memory, next K-chunk prefetched into registers while the current one
// is multiplied).  These kernels handle every shape; the tcgen05 kernels in
// tc_layer.cu replace the two activation-sized ones when width is 64/128.
#include <algorithm>

#include "common.cuh"

namespace b200fno {

constexpr int KC = 32;       // K chunk
constexpr int LDA = KC + 4;  // padded row of the A tile: conflict-free float4 reads
constexpr int TN = 64;       // tile columns (channels)

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// acc[r][c] += sum_k As[r][k] * Bs[k][c]   for one K chunk resident in shared memory.
// The 4x4 register tile is held as 4x2 packed fp32 pairs and updated with FFMA2 (sm_100 packed fp32:
// two FMAs per instruction, the A value broadcast to both halves), halving the FMA instruction count.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float a, float b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void ffma2(f32x2_t& acc, f32x2_t a, f32x2_t b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void mma_chunk(float (&acc)[4][4], const float* __restrict__ As, int lda,
                                          const float* __restrict__ Bs, int ldb) {
  f32x2_t c[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) c[r][0] = pk2(acc[r][0], acc[r][1]), c[r][1] = pk2(acc[r][2], acc[r][3]);
#pragma unroll
  for (int k = 0; k < KC; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(As + r * lda + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bs + (k + j) * ldb);
    f32x2_t blo[4], bhi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) blo[j] = pk2(b[j].x, b[j].y), bhi[j] = pk2(b[j].z, b[j].w);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const f32x2_t a0 = pk2(a[r].x, a[r].x), a1 = pk2(a[r].y, a[r].y), a2 = pk2(a[r].z, a[r].z),
                    a3 = pk2(a[r].w, a[r].w);
      ffma2(c[r][0], a0, blo[0]), ffma2(c[r][1], a0, bhi[0]);
      ffma2(c[r][0], a1, blo[1]), ffma2(c[r][1], a1, bhi[1]);
      ffma2(c[r][0], a2, blo[2]), ffma2(c[r][1], a2, bhi[2]);
      ffma2(c[r][0], a3, blo[3]), ffma2(c[r][1], a3, bhi[3]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[r][0]), "=f"(acc[r][1]) : "l"(c[r][0]));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[r][2]), "=f"(acc[r][3]) : "l"(c[r][1]));
  }
}

// bf16 compute mode (autocast semantics, tc_common.cuh): operand rounded to nearest-even bf16
__device__ __forceinline__ float bf16r(float x) {
  const uint32_t u = __float_as_uint(x);
  return __uint_as_float((u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u);
}
__device__ __forceinline__ float4 bf16r4(float4 v) { return make_float4(bf16r(v.x), bf16r(v.y), bf16r(v.z), bf16r(v.w)); }

__device__ __forceinline__ float gelu_erf(float v) {  // F.gelu default (fno.py:119,124)
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}

// ---------------------------------------------------------------------------
// lmul: Out[g][m][n] = sum_k L[m][k] R[g][k][n]
// grid (G, n tiles, m tiles); block TM*4 threads = (TM/4 row groups) x 16 column quads
// ---------------------------------------------------------------------------
// 16-byte async copy global -> shared; src_bytes == 0 zero-fills (out-of-range rows / columns)
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// K chunks in flight (template parameter NSTG): 4 for long contractions (latency-bound K loop), 2 when the whole
// contraction is at most two chunks (inverse transforms, K = 2*kept modes <= 64): half the shared memory, twice the
// resident CTAs per SM to overlap one CTA's stores with another's loads

template <int TM, int LM_STAGES>
__global__ void __launch_bounds__(TM * 4) lmul_kernel(const float* __restrict__ L, int ldl, int M, int K,
                                                      const float* __restrict__ R, long long strideRg,
                                                      long long strideRk, float* __restrict__ Out,
                                                      long long strideOg, long long strideOm, int N, int mdiv,
                                                      long long strideOmLo, long long split_off) {
  pdl_launch_dependents();
  pdl_wait();  // R is the previous kernel's output
  constexpr int NT = TM * 4;
  constexpr int A4 = TM * KC / 4 / NT;  // float4 per thread for the A tile (= 2)
  constexpr int B4 = KC * TN / 4 / NT;  // for the B tile (2 or 4)
  constexpr int STAGE = TM * LDA + KC * TN;
  extern __shared__ __align__(16) float lm_smem[];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long g = blockIdx.x;
  const int n0 = blockIdx.y * TN, m0 = blockIdx.z * TM;
  const float* Rg = R + g * strideRg;
  float acc[4][4] = {};
  auto issue = [&](int chunk) {
    const int k0 = chunk * KC;
    float* As = lm_smem + (chunk % LM_STAGES) * STAGE;
    float* Bs = As + TM * LDA;
#pragma unroll
    for (int i = 0; i < A4; ++i) {
      int idx = tid + i * NT, mm = idx / (KC / 4), kk = (idx % (KC / 4)) * 4;
      int m = m0 + mm, k = k0 + kk;
      const bool ok = m < M && k < ldl;
      cp_async16(As + mm * LDA + kk, ok ? L + (size_t)m * ldl + k : L, ok ? 16 : 0);
    }
#pragma unroll
    for (int i = 0; i < B4; ++i) {
      int idx = tid + i * NT, kk = idx / (TN / 4), nn = (idx % (TN / 4)) * 4;
      int k = k0 + kk, n = n0 + nn;
      const bool ok = k < K && n < N;
      cp_async16(Bs + kk * TN + nn, ok ? Rg + (long long)k * strideRk + n : R, ok ? 16 : 0);
    }
  };
  const int nchunks = (K + KC - 1) / KC;
#pragma unroll
  for (int s = 0; s < LM_STAGES - 1; ++s) {
    if (s < nchunks) issue(s);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<LM_STAGES - 2>();  // chunk c has landed
    __syncthreads();                 // ... for every thread, and chunk c-1's buffer is free
    if (c + LM_STAGES - 1 < nchunks) issue(c + LM_STAGES - 1);
    cp_async_commit();
    const float* As = lm_smem + (c % LM_STAGES) * STAGE;
    mma_chunk(acc, As + (ty * 4) * LDA, LDA, As + TM * LDA + tx * 4, TN);
  }
  const int n = n0 + tx * 4;
  if (n < N) {
    float* Og = Out + g * strideOg;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int m = m0 + ty * 4 + r;
      if (m < M) {
        // output row address: (m / mdiv) * strideOm + (m % mdiv) * strideOmLo  (mdiv == 1: plain m * strideOm)
        float* dst = Og + (long long)(m / mdiv) * strideOm + (long long)(m % mdiv) * strideOmLo + n;
        float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        if (split_off) {  // 3xTF32 operand planes for the tensor-core layer kernel: hi | lo
          // hi rounded to nearest tf32 (tc_common.cuh: tf32_hi)
          float4 hi = make_float4(__uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u),
                                  __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u),
                                  __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u),
                                  __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u));
          *reinterpret_cast<float4*>(dst) = hi;
          *reinterpret_cast<float4*>(dst + split_off) = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
        } else {
          *reinterpret_cast<float4*>(dst) = v;
        }
      }
    }
  }
}

int launch_lmul(const float* L, int ldl, int M, int K, const float* R, long long strideRg, long long strideRk,
                float* Out, long long strideOg, long long strideOm, int N, int G, cudaStream_t st, int mdiv,
                long long strideOmLo, long long split_off) {
  if (G <= 0 || M <= 0) return 0;
  const int nstg = K <= 2 * KC ? 2 : 4;
  const int SM32 = nstg * (32 * LDA + KC * TN) * 4, SM64 = nstg * (64 * LDA + KC * TN) * 4;
  // 32-row tiles when M is small or when 64-row tiles would leave most SMs without a CTA
  const long long ctas64 = (long long)G * ceil_div(N, TN) * ceil_div(M, 64);
  // ... or when 64-row tiles would spend more than a fifth of their rows on padding that 32-row tiles avoid
  // (3-D inverse-H: M = 2*70 = 140 -> 192 rows in 64-row tiles, 160 in 32-row tiles)
  const int pad64 = ceil_div(M, 64) * 64, pad32 = ceil_div(M, 32) * 32;
  const bool rows32 = M <= 32 || ctas64 < 2 * 148 || (pad32 < pad64 && (pad64 - M) * 5 > M);
#define B2_LMUL_LAUNCH(TMv, NSv, SMv)                                                                              \
  do {                                                                                                            \
    dim3 grid(G, ceil_div(N, TN), ceil_div(M, TMv));                                                              \
    B2_CUDA(cudaFuncSetAttribute(lmul_kernel<TMv, NSv>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMv));       \
    B2_CUDA(launch_kernel(lmul_kernel<TMv, NSv>, grid, dim3(TMv * 4), (size_t)SMv, st, L, ldl, M, K, R, strideRg,   \
                          strideRk, Out, strideOg, strideOm, N, mdiv, strideOmLo, split_off));                    \
  } while (0)
  if (rows32) {
    if (nstg == 2) B2_LMUL_LAUNCH(32, 2, SM32);
    else B2_LMUL_LAUNCH(32, 4, SM32);
  } else {
    if (nstg == 2) B2_LMUL_LAUNCH(64, 2, SM64);
    else B2_LMUL_LAUNCH(64, 4, SM64);
  }
#undef B2_LMUL_LAUNCH
  B2_LAUNCHED("lmul_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// Per-mode complex channel mixing (fno.py:41-43, einsum "bixyz,ioxyz->boxyz").
// One CTA per kept mode streams that mode's [Cp][2][Cp] weights exactly once.
// S, O: [B][2][NM][Cp];  Wpk: [NM][Cp(i)][2][Cp(o)]
// ---------------------------------------------------------------------------
constexpr int MODES_BCH = 32;  // batch entries staged per pass

// Thread layout: threadIdx.x % nq = output-channel quad, the remaining 256/nq thread groups are split
// between batch PAIRS and slices of the input-channel (i) loop.  With a small batch (the rollout runs
// at B = 8) most groups would otherwise idle while a few threads walk all Cp input channels with two
// dependent-latency weight loads per step; slicing i keeps every thread busy and 4x more loads in
// flight.  Partial sums of the slices are reduced through shared memory.
__global__ void __launch_bounds__(256, 4) modes_kernel(const float* __restrict__ S, const float* __restrict__ Wpk,
                                                    float* __restrict__ O, int B, int NM, int Cp) {
  extern __shared__ __align__(16) float Ss[];  // [MODES_BCH][2][Cp] inputs, then [nsl][nb][2][Cp] partials
  pdl_launch_dependents();
  pdl_wait();  // S is the previous kernel's output
  const int mode = blockIdx.x;
  const int OQ = Cp >> 2;      // column quads
  const int nq = min(OQ, 32);  // quads handled per sweep by threadIdx.x % nq
  const int tid = threadIdx.x;
  const int tx = tid % nq, ty = tid / nq, nty = 256 / nq;
  float* Ps = Ss + MODES_BCH * 2 * Cp;
  const float* Wm = Wpk + (size_t)mode * Cp * 2 * Cp;
  for (int b0 = 0; b0 < B; b0 += MODES_BCH) {
    const int nb = min(MODES_BCH, B - b0);
    const int nbp = (nb + 1) >> 1;                 // batch pairs in this pass
    const int nsl = max(1, min(nty / nbp, 8));     // i-slices
    const int isl = (Cp + nsl - 1) / nsl;          // input channels per slice
    __syncthreads();
    for (int idx = tid; idx < nb * 2 * OQ; idx += 256) {
      int q = idx % OQ, ri = (idx / OQ) & 1, bb = idx / (2 * OQ);
      *reinterpret_cast<float4*>(Ss + (bb * 2 + ri) * Cp + q * 4) =
          ldg4(S + (((size_t)(b0 + bb) * 2 + ri) * NM + mode) * Cp + q * 4);
    }
    __syncthreads();
    // work items = (batch pair, i-slice); with few thread groups (Cp >= 128: 256 / 32 = 8) and a full batch pass
    // (16 pairs) there are more pairs than groups, so every group walks its items
    for (int item = ty; item < nbp * nsl; item += nty) {
      const int pair = item % nbp, sl = item / nbp;
      const int bb = pair * 2;
      const bool two = bb + 1 < nb;
      const float* s0 = Ss + (bb * 2) * Cp;
      const float* s1 = Ss + ((two ? bb + 1 : bb) * 2) * Cp;
      const int i0 = sl * isl, i1 = min(Cp, i0 + isl);
      for (int q = tx; q < OQ; q += nq) {
        float4 r0 = zero4(), im0 = zero4(), r1 = zero4(), im1 = zero4();
#pragma unroll 8
        for (int i = i0; i < i1; ++i) {
          const float4 wr = ldg4(Wm + ((size_t)i * 2 + 0) * Cp + q * 4);
          const float4 wi = ldg4(Wm + ((size_t)i * 2 + 1) * Cp + q * 4);
          const float ar = s0[i], ai = s0[Cp + i], br = s1[i], bi = s1[Cp + i];
          r0.x = fmaf(ar, wr.x, r0.x); r0.y = fmaf(ar, wr.y, r0.y); r0.z = fmaf(ar, wr.z, r0.z); r0.w = fmaf(ar, wr.w, r0.w);
          r0.x = fmaf(-ai, wi.x, r0.x); r0.y = fmaf(-ai, wi.y, r0.y); r0.z = fmaf(-ai, wi.z, r0.z); r0.w = fmaf(-ai, wi.w, r0.w);
          im0.x = fmaf(ar, wi.x, im0.x); im0.y = fmaf(ar, wi.y, im0.y); im0.z = fmaf(ar, wi.z, im0.z); im0.w = fmaf(ar, wi.w, im0.w);
          im0.x = fmaf(ai, wr.x, im0.x); im0.y = fmaf(ai, wr.y, im0.y); im0.z = fmaf(ai, wr.z, im0.z); im0.w = fmaf(ai, wr.w, im0.w);
          r1.x = fmaf(br, wr.x, r1.x); r1.y = fmaf(br, wr.y, r1.y); r1.z = fmaf(br, wr.z, r1.z); r1.w = fmaf(br, wr.w, r1.w);
          r1.x = fmaf(-bi, wi.x, r1.x); r1.y = fmaf(-bi, wi.y, r1.y); r1.z = fmaf(-bi, wi.z, r1.z); r1.w = fmaf(-bi, wi.w, r1.w);
          im1.x = fmaf(br, wi.x, im1.x); im1.y = fmaf(br, wi.y, im1.y); im1.z = fmaf(br, wi.z, im1.z); im1.w = fmaf(br, wi.w, im1.w);
          im1.x = fmaf(bi, wr.x, im1.x); im1.y = fmaf(bi, wr.y, im1.y); im1.z = fmaf(bi, wr.z, im1.z); im1.w = fmaf(bi, wr.w, im1.w);
        }
        float* pp = Ps + ((size_t)(sl * nb + bb) * 2) * Cp + q * 4;
        *reinterpret_cast<float4*>(pp) = r0;
        *reinterpret_cast<float4*>(pp + Cp) = im0;
        if (two) {
          *reinterpret_cast<float4*>(pp + 2 * Cp) = r1;
          *reinterpret_cast<float4*>(pp + 3 * Cp) = im1;
        }
      }
    }
    __syncthreads();
    // reduce the i-slices and write O[b][ri][mode][:]
    for (int idx = tid; idx < nb * 2 * OQ; idx += 256) {
      int q = idx % OQ, ri = (idx / OQ) & 1, bb = idx / (2 * OQ);
      float4 acc = zero4();
      for (int s2 = 0; s2 < nsl; ++s2) {
        const float4 v = *reinterpret_cast<const float4*>(Ps + ((size_t)(s2 * nb + bb) * 2 + ri) * Cp + q * 4);
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
      }
      *reinterpret_cast<float4*>(O + (((size_t)(b0 + bb) * 2 + ri) * NM + mode) * Cp + q * 4) = acc;
    }
  }
}

int launch_modes(const float* S, const float* Wpk, float* O, int B, int NM, int Cp, cudaStream_t st) {
  // inputs + partial sums: nsl * nb <= max(MODES_BCH, 2 * thread groups) entries of 2*Cp floats
  const int nq = std::min(Cp / 4, 32), nty = 256 / nq;
  size_t smem = (size_t)(MODES_BCH + std::max(MODES_BCH, 2 * nty)) * 2 * Cp * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("modes kernel: width %d too large", Cp);
    return B200FNO_EINVAL;
  }
  B2_CUDA(cudaFuncSetAttribute(modes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B2_CUDA(launch_kernel(modes_kernel, dim3(NM), dim3(256), smem, st, S, Wpk, O, B, NM, Cp));
  B2_LAUNCHED("modes_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// Fused Fourier-layer body (fno.py:114-119):
//   out = act( bn( conv1x1(in) + irfft_W(D) ) )       per row (b,t,h), 64 points x 64 channels per CTA
// K loop = [Cp input channels of the bypass conv] ++ [2*m3 rows of the inverse W transform]
// grid (rows, point tiles, channel tiles); block 256
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layer_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                    const float* __restrict__ convT, const float* __restrict__ Gt,
                                                    const float* __restrict__ D, const float* __restrict__ scale,
                                                    const float* __restrict__ shift, int Wp, int Cp, int K2, int K2p,
                                                    int gelu, int bf16) {
  __shared__ __align__(16) float As[64 * LDA];
  __shared__ __align__(16) float Bs[KC * TN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long row = blockIdx.x;
  const int p0 = blockIdx.y * 64, o0 = blockIdx.z * TN;
  const float* in_row = in + (size_t)row * Wp * Cp;
  const float* D_row = D + (size_t)row * K2 * Cp;
  float acc[4][4] = {};
  float4 ra[2], rb[2];
  const int nc_conv = convT ? (Cp + KC - 1) / KC : 0;
  const int nc_inv = (K2p + KC - 1) / KC;
  auto fetch = [&](int c) {
    if (c < nc_conv) {
      const int k0 = c * KC;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int idx = tid + i * 256, pp = idx >> 3, kk = (idx & 7) * 4;
        int p = p0 + pp, k = k0 + kk;
        ra[i] = (p < Wp && k < Cp) ? ldg4(in_row + (size_t)p * Cp + k) : zero4();
        if (bf16) ra[i] = bf16r4(ra[i]);  // conv input cast to bf16 (the weights were rounded when packed)
        int kb = idx >> 4, nn = (idx & 15) * 4;
        int kr = k0 + kb, o = o0 + nn;
        rb[i] = (kr < Cp && o < Cp) ? ldg4(convT + (size_t)kr * Cp + o) : zero4();
      }
    } else {
      const int k0 = (c - nc_conv) * KC;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int idx = tid + i * 256, pp = idx >> 3, kk = (idx & 7) * 4;
        int p = p0 + pp, k = k0 + kk;
        ra[i] = (p < Wp && k < K2p) ? ldg4(Gt + (size_t)p * K2p + k) : zero4();
        int kb = idx >> 4, nn = (idx & 15) * 4;
        int kr = k0 + kb, o = o0 + nn;
        rb[i] = (kr < K2 && o < Cp) ? ldg4(D_row + (size_t)kr * Cp + o) : zero4();
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256;
      *reinterpret_cast<float4*>(As + (idx >> 3) * LDA + (idx & 7) * 4) = ra[i];
      *reinterpret_cast<float4*>(Bs + (idx >> 4) * TN + (idx & 15) * 4) = rb[i];
    }
  };
  const int nchunks = nc_conv + nc_inv;
  fetch(0);
  for (int c = 0; c < nchunks; ++c) {
    stash();
    __syncthreads();
    if (c + 1 < nchunks) fetch(c + 1);
    mma_chunk(acc, As + (ty * 4) * LDA, LDA, Bs + tx * 4, TN);
    __syncthreads();
  }
  const int o = o0 + tx * 4;
  if (o < Cp) {
    float4 sc = scale ? ldg4(scale + o) : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 sh = shift ? ldg4(shift + o) : zero4();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int p = p0 + ty * 4 + r;
      if (p < Wp) {
        float4 v = make_float4(fmaf(acc[r][0], sc.x, sh.x), fmaf(acc[r][1], sc.y, sh.y), fmaf(acc[r][2], sc.z, sh.z),
                               fmaf(acc[r][3], sc.w, sh.w));
        if (gelu) v = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
        *reinterpret_cast<float4*>(out + ((size_t)row * Wp + p) * Cp + o) = v;
      }
    }
  }
}

int launch_layer(const float* act_in, float* act_out, const float* convT, const float* Gt, const float* D,
                 const float* scale, const float* shift, long long rows, int Wp, int Cp, int K2, int K2p, int gelu,
                 cudaStream_t st, int bf16) {
  dim3 grid((unsigned)rows, ceil_div(Wp, 64), ceil_div(Cp, TN));
  layer_kernel<<<grid, 256, 0, st>>>(act_in, act_out, convT, Gt, D, scale, shift, Wp, Cp, K2, K2p, gelu, bf16);
  B2_LAUNCHED("layer_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// Lift (fno.py:106-111): cat(x, grid) -> fc0 -> channels-last, zero padded.
// A row of the K dimension = [input features | grid coordinates | 1 (bias)].
// grid (padded rows (b,t,h), point tiles, channel tiles)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lift_kernel(LiftArgs a) {
  __shared__ __align__(16) float As[64 * LDA];
  __shared__ __align__(16) float Bs[KC * TN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long row = blockIdx.x;
  const int h = (int)(row % a.Hp);
  const int t = (int)((row / a.Hp) % a.Tp);
  const int b = (int)(row / ((long long)a.Hp * a.Tp));
  const int p0 = blockIdx.y * 64, o0 = blockIdx.z * TN;
  float* out_row = a.act + (size_t)row * a.Wp * a.Cp;
  const int o = o0 + tx * 4;
  const bool row_valid = (h < a.H) && (t < a.T);
  float acc[4][4] = {};
  if (row_valid && p0 < a.W) {
    const float* xb = a.x + (size_t)b * a.x_sB + (size_t)t * a.x_sT + (size_t)h * a.W * a.c_in;
    const float gtv = a.gt ? a.gt[t] : 0.f, ghv = a.gh[h];
    const int Kl = a.Fin + a.ng + 1;
    for (int k0 = 0; k0 < Kl; k0 += KC) {
      for (int idx = tid; idx < 64 * KC; idx += 256) {
        int pp = idx / KC, kk = idx % KC;
        int w = p0 + pp, j = k0 + kk;
        float v = 0.f;
        if (w < a.W) {
          if (j < a.Fin) v = __ldg(xb + (size_t)w * a.c_in + a.in_off[j]);
          else if (j < a.Fin + a.ng) {
            int gi = j - a.Fin + (a.gt ? 0 : 1);  // 0: t, 1: h, 2: w
            v = gi == 0 ? gtv : (gi == 1 ? ghv : a.gw[w]);
          } else if (j == a.Fin + a.ng) v = 1.f;
        }
        As[pp * LDA + kk] = a.bf16 ? bf16r(v) : v;
      }
      for (int idx = tid; idx < KC * TN / 4; idx += 256) {
        int kb = idx >> 4, nn = (idx & 15) * 4;
        int kr = k0 + kb, oo = o0 + nn;
        *reinterpret_cast<float4*>(Bs + kb * TN + nn) =
            (kr < a.Klp && oo < a.Cp) ? ldg4(a.W0T + (size_t)kr * a.Cp + oo) : zero4();
      }
      __syncthreads();
      mma_chunk(acc, As + (ty * 4) * LDA, LDA, Bs + tx * 4, TN);
      __syncthreads();
    }
  }
  if (o < a.Cp) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int p = p0 + ty * 4 + r;
      if (p < a.Wp)
        *reinterpret_cast<float4*>(out_row + (size_t)p * a.Cp + o) =
            make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
  }
}

int launch_lift(const LiftArgs& a, cudaStream_t st) {
  dim3 grid((unsigned)((long long)a.B * a.Tp * a.Hp), ceil_div(a.Wp, 64), ceil_div(a.Cp, TN));
  lift_kernel<<<grid, 256, 0, st>>>(a);
  B2_LAUNCHED("lift_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// Projection (fno.py:121-128) + rollout glue (eval.py:315-318 as one affine):
//   crop -> fc1 -> GELU -> fc2 -> unfold -> p*a[c]+b[c] -> prediction slice and next model input.
// grid (valid rows (b,t,h), point tiles over W); block 256; 64 points per CTA
// ---------------------------------------------------------------------------
constexpr int PH = 128;        // proj_hidden (fno.py:102)
constexpr int LDH = PH + 4;

__global__ void __launch_bounds__(256) proj_kernel(ProjArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* As = sm;                  // [64][LDA]
  float* Bs = As + 64 * LDA;       // [KC][128]
  float* Hs = Bs + KC * PH;        // [64][LDH]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long row = blockIdx.x;
  const int h = (int)(row % a.H);
  const int t = (int)((row / a.H) % a.T);
  const int b = (int)(row / ((long long)a.H * a.T));
  const int p0 = blockIdx.y * 64;
  const float* in_row = a.act + (((size_t)b * a.Tp + t) * a.Hp + h) * (size_t)a.Wp * a.Cp;
  // ---- GEMM 1: [64 x Cp] . [Cp x 128]
  float acc0[4][4] = {}, acc1[4][4] = {};
  for (int k0 = 0; k0 < a.Cp; k0 += KC) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256, pp = idx >> 3, kk = (idx & 7) * 4;
      int p = p0 + pp, k = k0 + kk;
      float4 xv = (p < a.W && k < a.Cp) ? ldg4(in_row + (size_t)p * a.Cp + k) : zero4();
      if (a.bf16) xv = bf16r4(xv);
      *reinterpret_cast<float4*>(As + pp * LDA + kk) = xv;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256, kb = idx >> 5, nn = (idx & 31) * 4;
      int kr = k0 + kb;
      *reinterpret_cast<float4*>(Bs + kb * PH + nn) = (kr < a.Cp) ? ldg4(a.fc1T + (size_t)kr * PH + nn) : zero4();
    }
    __syncthreads();
    mma_chunk(acc0, As + (ty * 4) * LDA, LDA, Bs + tx * 4, PH);
    mma_chunk(acc1, As + (ty * 4) * LDA, LDA, Bs + 64 + tx * 4, PH);
    __syncthreads();
  }
  {
    const float4 b0 = ldg4(a.fc1b + tx * 4), b1 = ldg4(a.fc1b + 64 + tx * 4);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float* hrow = Hs + (ty * 4 + r) * LDH;
      float4 h0 = make_float4(gelu_erf(acc0[r][0] + b0.x), gelu_erf(acc0[r][1] + b0.y), gelu_erf(acc0[r][2] + b0.z),
                              gelu_erf(acc0[r][3] + b0.w));
      float4 h1 = make_float4(gelu_erf(acc1[r][0] + b1.x), gelu_erf(acc1[r][1] + b1.y), gelu_erf(acc1[r][2] + b1.z),
                              gelu_erf(acc1[r][3] + b1.w));
      if (a.bf16) h0 = bf16r4(h0), h1 = bf16r4(h1);  // the fc2 operand is a bf16 tensor under autocast
      *reinterpret_cast<float4*>(hrow + tx * 4) = h0;
      *reinterpret_cast<float4*>(hrow + 64 + tx * 4) = h1;
    }
  }
  __syncthreads();
  // ---- GEMM 2: [64 x 128] . [128 x Fp], 64 features per sweep
  const size_t pt_out = (size_t)b * a.out_sB + (size_t)t * a.out_sT + (size_t)h * a.W * a.c_out;
  const size_t pt_st = (size_t)b * a.st_sB + (size_t)t * a.st_sT + (size_t)h * a.W * a.c_in;
  for (int f0 = 0; f0 < a.Fp; f0 += TN) {
    float acc[4][4] = {};
    for (int k0 = 0; k0 < PH; k0 += KC) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int idx = tid + i * 256, kb = idx >> 4, nn = (idx & 15) * 4;
        int f = f0 + nn;
        *reinterpret_cast<float4*>(Bs + kb * TN + nn) =
            (f < a.Fp) ? ldg4(a.fc2T + (size_t)(k0 + kb) * a.Fp + f) : zero4();
      }
      __syncthreads();
      mma_chunk(acc, Hs + (ty * 4) * LDH + k0, LDH, Bs + tx * 4, TN);
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = f0 + tx * 4 + j;
      if (f >= a.Fout) continue;
      const int ch = a.chan[f];
      const float bias = __ldg(a.fc2b + f);
      const float sa = a.aff_a ? __ldg(a.aff_a + ch) : 1.f, sb = a.aff_b ? __ldg(a.aff_b + ch) : 0.f;
      const int oo = a.out_off[f], so = a.st_off[f];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int w = p0 + ty * 4 + r;
        if (w >= a.W) continue;
        const float y = acc[r][j] + bias;
        const float v = a.aff_a ? fmaf(y, sa, sb) : y;
        a.out[pt_out + (size_t)w * a.c_out + oo] = v;
        if (a.state) a.state[pt_st + (size_t)w * a.c_in + so] = v;
      }
    }
  }
}

constexpr size_t PROJ_SMEM = (size_t)(64 * LDA + KC * PH + 64 * LDH) * sizeof(float);

int launch_proj(const ProjArgs& a, cudaStream_t st) {
  B2_CUDA(cudaFuncSetAttribute(proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROJ_SMEM));
  dim3 grid((unsigned)((long long)a.B * a.T * a.H), ceil_div(a.W, 64));
  proj_kernel<<<grid, 256, PROJ_SMEM, st>>>(a);
  B2_LAUNCHED("proj_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------
// x [B][C][S] -> act [B][S][Cp] (zero-padded channels)
__global__ void nchw_to_cl_kernel(const float* __restrict__ x, float* __restrict__ act, int C, long long S, int Cp) {
  __shared__ float tile[32][33];
  const long long s0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32, b = blockIdx.z;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long long s = s0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && s < S) ? x[((size_t)b * C + c) * S + s] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long s = s0 + i;
    int c = c0 + threadIdx.x;
    if (s < S && c < Cp) act[((size_t)b * S + s) * Cp + c] = tile[threadIdx.x][i];
  }
}
__global__ void cl_to_nchw_kernel(const float* __restrict__ act, float* __restrict__ y, int C, long long S, int Cp) {
  __shared__ float tile[32][33];
  const long long s0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32, b = blockIdx.z;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long s = s0 + i;
    int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (s < S && c < Cp) ? act[((size_t)b * S + s) * Cp + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long long s = s0 + threadIdx.x;
    if (c < C && s < S) y[((size_t)b * C + c) * S + s] = tile[threadIdx.x][i];
  }
}
int launch_nchw_to_cl(const float* x, float* act, int B, int C, long long S, int Cp, cudaStream_t st) {
  dim3 grid((unsigned)((S + 31) / 32), ceil_div(Cp, 32), B);
  nchw_to_cl_kernel<<<grid, dim3(32, 8), 0, st>>>(x, act, C, S, Cp);
  B2_LAUNCHED("nchw_to_cl_kernel");
  return 0;
}
int launch_cl_to_nchw(const float* act, float* y, int B, int C, long long S, int Cp, cudaStream_t st) {
  dim3 grid((unsigned)((S + 31) / 32), ceil_div(Cp, 32), B);
  cl_to_nchw_kernel<<<grid, dim3(32, 8), 0, st>>>(act, y, C, S, Cp);
  B2_LAUNCHED("cl_to_nchw_kernel");
  return 0;
}

// state[s][point][c_out..c_in) = x0[point][c_out..c_in) for both ping-pong states (eval.py:317)
__global__ void copy_params_kernel(const float* __restrict__ x0, float* __restrict__ state, long long points, int c_in,
                                   int c_out) {
  const int pc = c_in - c_out;
  const long long n = points * pc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long p = i / pc;
    int c = c_out + (int)(i % pc);
    float v = x0[p * c_in + c];
    state[p * c_in + c] = v;
    state[(points + p) * c_in + c] = v;
  }
}
int launch_copy_params(const float* x0, float* state, long long points, int c_in, int c_out, cudaStream_t st) {
  long long n = points * (c_in - c_out);
  if (n <= 0) return 0;
  int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
  copy_params_kernel<<<blocks, 256, 0, st>>>(x0, state, points, c_in, c_out);
  B2_LAUNCHED("copy_params_kernel");
  return 0;
}

}  // namespace b200fno
