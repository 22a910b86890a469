// Training-path kernels (train.py:321-334): train-mode BatchNorm
// (batch statistics over the PADDED tensor, fno.py:111,117), and the backward pass of every stage.
//
// Every stage of the forward is a real linear map (truncated DFT tables, 1x1 conv, per-mode mixing) or a
// pointwise function, so the backward re-uses the forward kernels of simt.cu with TRANSPOSED tables /
// weights (api.cu: b200fno_train_backward); this file adds what has no forward counterpart:
//   colstats / bn_finalize / bn_apply          train-mode BatchNorm forward
//   bn_bwd_reduce / bn_bwd_finalize / bn_bwd_apply   GELU' + BatchNorm backward
//   wgrad                                       dW[m][n] = sum_points A[p][m] B[p][n]   (fc0, conv, fc1, fc2)
//   colsum                                      bias gradients
//   proj_bwd                                    backward of crop -> fc1 -> GELU -> fc2 -> unfold
//   lift_features                               the lift's K rows [features | grid | 1] per padded point
//   modes_wgrad / pack_spectral_adj / unpack_spectral_grad    spectral weight gradients
// fp32 FFMA throughout; reductions accumulate per-CTA partial sums in fp32 and combine them with
// double (statistics) or float (weight gradients) atomics.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace b200fno {

namespace {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }
// d/dv [ v * Phi(v) ] = Phi(v) + v * phi(v)
__device__ __forceinline__ float gelu_grad(float v) {
  const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
  return fmaf(v, pdf, cdf);
}

constexpr int KC = 32;
constexpr int LDA = KC + 4;
constexpr int TN = 64;
constexpr int PH = 128;
constexpr int LDH = PH + 4;

// acc[r][c] += sum_k As[r][k] * Bs[k][c]   (As row-major with k contiguous)
__device__ __forceinline__ void mma_chunk(float (&acc)[4][4], const float* __restrict__ As, int lda,
                                          const float* __restrict__ Bs, int ldb) {
#pragma unroll
  for (int k = 0; k < KC; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(As + r * lda + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bs + (k + j) * ldb);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float av[4] = {a[r].x, a[r].y, a[r].z, a[r].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[r][0] = fmaf(av[j], b[j].x, acc[r][0]);
        acc[r][1] = fmaf(av[j], b[j].y, acc[r][1]);
        acc[r][2] = fmaf(av[j], b[j].z, acc[r][2]);
        acc[r][3] = fmaf(av[j], b[j].w, acc[r][3]);
      }
    }
  }
}

// Per-channel reduction of two float4 partial sums held by the threads of a CTA laid out as
// (tx = channel quad, ty = row group); one double atomicAdd per channel per CTA.
__device__ __forceinline__ void block_colreduce(float4 s0, float4 s1, int OQ, int nty, int tx, int ty, int Cp,
                                                double* __restrict__ out0, double* __restrict__ out1, float* sh) {
  // sh: [2][nty][Cp]
  if (ty < nty) {
    *reinterpret_cast<float4*>(sh + (size_t)ty * Cp + tx * 4) = s0;
    *reinterpret_cast<float4*>(sh + (size_t)(nty + ty) * Cp + tx * 4) = s1;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * Cp; c += blockDim.x) {
    const int which = c / Cp, ch = c % Cp;
    double a = 0.0;
    for (int y = 0; y < nty; ++y) a += (double)sh[(size_t)(which * nty + y) * Cp + ch];
    atomicAdd((which ? out1 : out0) + ch, a);
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// train-mode BatchNorm forward
// ---------------------------------------------------------------------------
// stats[0][c] += sum_p x[p][c], stats[1][c] += sum_p x[p][c]^2
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ x, long long P, int Cp,
                                                       double* __restrict__ stats) {
  extern __shared__ __align__(16) float cs_sh[];
  const int OQ = Cp >> 2, nty = blockDim.x / OQ, tx = threadIdx.x % OQ, ty = threadIdx.x / OQ;
  float4 s = zero4(), ss = zero4();
  if (ty < nty)
    for (long long r = (long long)blockIdx.x * nty + ty; r < P; r += (long long)gridDim.x * nty) {
      const float4 v = ldg4(x + r * Cp + tx * 4);
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
      ss.x = fmaf(v.x, v.x, ss.x), ss.y = fmaf(v.y, v.y, ss.y), ss.z = fmaf(v.z, v.z, ss.z), ss.w = fmaf(v.w, v.w, ss.w);
    }
  block_colreduce(s, ss, OQ, nty, tx, ty, Cp, stats, stats + Cp, cs_sh);
}

// bnc: [mean | rstd | a | b | gamma*rstd | unused] x Cp floats;  y = a*z + b  with a = gamma*rstd, b = beta - mean*a.
// Running statistics are updated like nn.BatchNorm*d in training mode (momentum, unbiased variance).
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double n, float eps, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int C, int Cp, float* __restrict__ bnc,
                                   float* __restrict__ run_mean, float* __restrict__ run_var, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  float mean = 0.f, rstd = 0.f, a = 0.f, b = 0.f;
  if (c < C) {
    const double m = stats[c] / n;
    double var = stats[Cp + c] / n - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
    a = gamma[c] * rstd;
    b = beta[c] - mean * a;
    if (run_mean) run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
    if (run_var) run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)(var * (n / (n > 1.0 ? n - 1.0 : 1.0)));
  }
  bnc[c] = mean, bnc[Cp + c] = rstd, bnc[2 * Cp + c] = a, bnc[3 * Cp + c] = b;
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ z, float* __restrict__ y,
                                                       long long nquads, int OQ, const float* __restrict__ bnc, int Cp,
                                                       int gelu) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nquads; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % OQ);
    const float4 a = ldg4(bnc + 2 * Cp + q * 4), b = ldg4(bnc + 3 * Cp + q * 4);
    const float4 v = ldg4(z + i * 4);
    float4 o = make_float4(fmaf(v.x, a.x, b.x), fmaf(v.y, a.y, b.y), fmaf(v.z, a.z, b.z), fmaf(v.w, a.w, b.w));
    if (gelu) o = make_float4(gelu_erf(o.x), gelu_erf(o.y), gelu_erf(o.z), gelu_erf(o.w));
    *reinterpret_cast<float4*>(y + i * 4) = o;
  }
}

// ---------------------------------------------------------------------------
// GELU' + BatchNorm backward
//   dyh = gy * gelu'(a*z+b)  (or gy for the last layer);  xh = (z - mean) * rstd
//   s1 = sum dyh, s2 = sum dyh*xh;  dz = gamma*rstd * (dyh - s1/n - xh*s2/n)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ gy, const float* __restrict__ z,
                                                            long long P, int Cp, const float* __restrict__ bnc,
                                                            int gelu, double* __restrict__ sums) {
  extern __shared__ __align__(16) float cs_sh[];
  const int OQ = Cp >> 2, nty = blockDim.x / OQ, tx = threadIdx.x % OQ, ty = threadIdx.x / OQ;
  float4 s1 = zero4(), s2 = zero4();
  if (ty < nty) {
    const float4 mean = ldg4(bnc + tx * 4), rstd = ldg4(bnc + Cp + tx * 4);
    const float4 a = ldg4(bnc + 2 * Cp + tx * 4), b = ldg4(bnc + 3 * Cp + tx * 4);
    for (long long r = (long long)blockIdx.x * nty + ty; r < P; r += (long long)gridDim.x * nty) {
      const float4 v = ldg4(z + r * Cp + tx * 4);
      float4 g = ldg4(gy + r * Cp + tx * 4);
      if (gelu) {
        g.x *= gelu_grad(fmaf(v.x, a.x, b.x)), g.y *= gelu_grad(fmaf(v.y, a.y, b.y));
        g.z *= gelu_grad(fmaf(v.z, a.z, b.z)), g.w *= gelu_grad(fmaf(v.w, a.w, b.w));
      }
      s1.x += g.x, s1.y += g.y, s1.z += g.z, s1.w += g.w;
      s2.x = fmaf(g.x, (v.x - mean.x) * rstd.x, s2.x), s2.y = fmaf(g.y, (v.y - mean.y) * rstd.y, s2.y);
      s2.z = fmaf(g.z, (v.z - mean.z) * rstd.z, s2.z), s2.w = fmaf(g.w, (v.w - mean.w) * rstd.w, s2.w);
    }
  }
  block_colreduce(s1, s2, OQ, nty, tx, ty, Cp, sums, sums + Cp, cs_sh);
}

// d(bn.weight) = s2, d(bn.bias) = s1;  bnc[4] = s1/n, bnc[5] = s2/n for bn_bwd_apply
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, double n, int C, int Cp, float* __restrict__ bnc,
                                       float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  const double s1 = c < C ? sums[c] : 0.0, s2 = c < C ? sums[Cp + c] : 0.0;
  bnc[4 * Cp + c] = (float)(s1 / n);
  bnc[5 * Cp + c] = (float)(s2 / n);
  if (c < C) {
    if (d_gamma) d_gamma[c] = (float)s2;
    if (d_beta) d_beta[c] = (float)s1;
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ gy, const float* __restrict__ z,
                                                           float* __restrict__ dz, long long nquads, int OQ,
                                                           const float* __restrict__ bnc, int Cp, int gelu) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nquads; i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % OQ);
    const float4 mean = ldg4(bnc + q * 4), rstd = ldg4(bnc + Cp + q * 4);
    const float4 a = ldg4(bnc + 2 * Cp + q * 4), b = ldg4(bnc + 3 * Cp + q * 4);
    const float4 c1 = ldg4(bnc + 4 * Cp + q * 4), c2 = ldg4(bnc + 5 * Cp + q * 4);
    const float4 v = ldg4(z + i * 4);
    float4 g = ldg4(gy + i * 4);
    if (gelu) {
      g.x *= gelu_grad(fmaf(v.x, a.x, b.x)), g.y *= gelu_grad(fmaf(v.y, a.y, b.y));
      g.z *= gelu_grad(fmaf(v.z, a.z, b.z)), g.w *= gelu_grad(fmaf(v.w, a.w, b.w));
    }
    float4 o;
    o.x = a.x * (g.x - c1.x - (v.x - mean.x) * rstd.x * c2.x);
    o.y = a.y * (g.y - c1.y - (v.y - mean.y) * rstd.y * c2.y);
    o.z = a.z * (g.z - c1.z - (v.z - mean.z) * rstd.z * c2.z);
    o.w = a.w * (g.w - c1.w - (v.w - mean.w) * rstd.w * c2.w);
    *reinterpret_cast<float4*>(dz + i * 4) = o;
  }
}

static int colreduce_cfg(int Cp, long long P, int* blocks, size_t* smem) {
  const int OQ = Cp / 4;
  if (OQ < 1 || OQ > 256) {
    set_error("training kernels need 4 <= padded width <= 1024, got %d", Cp);
    return B200FNO_EINVAL;
  }
  const int nty = 256 / OQ;
  *smem = (size_t)2 * nty * Cp * sizeof(float);
  *blocks = (int)std::max<long long>(1, std::min<long long>((P + nty - 1) / nty, 148 * 4));
  return 0;
}

int launch_colstats(const float* x, long long P, int Cp, double* stats, cudaStream_t st) {
  int blocks;
  size_t smem;
  B2_TRY(colreduce_cfg(Cp, P, &blocks, &smem));
  B2_CUDA(cudaMemsetAsync(stats, 0, 2 * (size_t)Cp * sizeof(double), st));
  colstats_kernel<<<blocks, 256, smem, st>>>(x, P, Cp, stats);
  B2_LAUNCHED("colstats_kernel");
  return 0;
}

int launch_bn_finalize(const double* stats, long long P, float eps, const float* gamma, const float* beta, int C, int Cp,
                       float* bnc, float* run_mean, float* run_var, float momentum, cudaStream_t st) {
  bn_finalize_kernel<<<ceil_div(Cp, 128), 128, 0, st>>>(stats, (double)P, eps, gamma, beta, C, Cp, bnc, run_mean,
                                                         run_var, momentum);
  B2_LAUNCHED("bn_finalize_kernel");
  return 0;
}

int launch_bn_apply(const float* z, float* y, long long P, int Cp, const float* bnc, int gelu, cudaStream_t st) {
  const long long nq = P * (Cp / 4);
  const int blocks = (int)std::max<long long>(1, std::min<long long>((nq + 255) / 256, 148 * 16));
  bn_apply_kernel<<<blocks, 256, 0, st>>>(z, y, nq, Cp / 4, bnc, Cp, gelu);
  B2_LAUNCHED("bn_apply_kernel");
  return 0;
}

int launch_bn_backward(const float* gy, const float* z, float* dz, long long P, int C, int Cp, float* bnc, int gelu,
                       double* sums, float* d_gamma, float* d_beta, cudaStream_t st) {
  int blocks;
  size_t smem;
  B2_TRY(colreduce_cfg(Cp, P, &blocks, &smem));
  B2_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)Cp * sizeof(double), st));
  bn_bwd_reduce_kernel<<<blocks, 256, smem, st>>>(gy, z, P, Cp, bnc, gelu, sums);
  B2_LAUNCHED("bn_bwd_reduce_kernel");
  bn_bwd_finalize_kernel<<<ceil_div(Cp, 128), 128, 0, st>>>(sums, (double)P, C, Cp, bnc, d_gamma, d_beta);
  B2_LAUNCHED("bn_bwd_finalize_kernel");
  const long long nq = P * (Cp / 4);
  const int b2 = (int)std::max<long long>(1, std::min<long long>((nq + 255) / 256, 148 * 16));
  bn_bwd_apply_kernel<<<b2, 256, 0, st>>>(gy, z, dz, nq, Cp / 4, bnc, Cp, gelu);
  B2_LAUNCHED("bn_bwd_apply_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// bf16 compute mode (torch.autocast semantics: both operands of every Linear / Conv GEMM, forward AND backward, are bf16
// tensors; accumulation fp32): round to nearest-even bf16
__device__ __forceinline__ float tr_bf16r(float x) {
  const uint32_t u = __float_as_uint(x);
  return __uint_as_float((u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u);
}
__device__ __forceinline__ float4 tr_bf16r4(float4 v) {
  return make_float4(tr_bf16r(v.x), tr_bf16r(v.y), tr_bf16r(v.z), tr_bf16r(v.w));
}

// wgrad: out[m*ldo + n] += sum_p A[p*lda + m] * B[p*ldb + n]      m < Mv, n < Nv
// (column n == Nv goes to extra[m] when extra != nullptr: the lift's bias column).
// Split over points: grid.x CTAs each own every grid.x-th chunk of 32 points; 64x64 output tile per CTA.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ A, int lda, int M, int Mv,
                                                    const float* __restrict__ B, int ldb, int N, int Nv, long long P,
                                                    float* __restrict__ out, int ldo, float* __restrict__ extra,
                                                    int bf16) {
  __shared__ __align__(16) float As[2][KC * 64];
  __shared__ __align__(16) float Bs[2][KC * 64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.z * 64;
  const long long chunks = (P + KC - 1) / KC;
  float acc[4][4] = {};
  float4 ra[2], rb[2];
  auto fetch = [&](long long c) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256, kk = idx >> 4, cc = (idx & 15) * 4;
      const long long p = c * KC + kk;
      ra[i] = (p < P && m0 + cc < M) ? ldg4(A + p * lda + m0 + cc) : zero4();
      rb[i] = (p < P && n0 + cc < N) ? ldg4(B + p * ldb + n0 + cc) : zero4();
      if (bf16) ra[i] = tr_bf16r4(ra[i]), rb[i] = tr_bf16r4(rb[i]);  // grad_output and the layer input as bf16 tensors
    }
  };
  long long c = blockIdx.x;
  if (c < chunks) fetch(c);
  int buf = 0;
  for (; c < chunks; c += gridDim.x, buf ^= 1) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      *reinterpret_cast<float4*>(&As[buf][idx * 4]) = ra[i];
      *reinterpret_cast<float4*>(&Bs[buf][idx * 4]) = rb[i];
    }
    __syncthreads();  // double-buffered: the other buffer was last read before the previous barrier
    if (c + gridDim.x < chunks) fetch(c + gridDim.x);
#pragma unroll 8
    for (int k = 0; k < KC; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k * 64 + ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k * 64 + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r][0] = fmaf(av[r], b.x, acc[r][0]);
        acc[r][1] = fmaf(av[r], b.y, acc[r][1]);
        acc[r][2] = fmaf(av[r], b.z, acc[r][2]);
        acc[r][3] = fmaf(av[r], b.w, acc[r][3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = m0 + ty * 4 + r;
    if (m >= Mv) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < Nv) atomicAdd(out + (size_t)m * ldo + n, acc[r][j]);
      else if (n == Nv && extra) atomicAdd(extra + m, acc[r][j]);
    }
  }
}

int launch_wgrad(const float* A, int lda, int M, int Mv, const float* B, int ldb, int N, int Nv, long long P, float* out,
                 int ldo, float* extra, cudaStream_t st, int bf16) {
  if (P <= 0) return 0;
  if ((lda | ldb | M | N) & 3) {
    set_error("internal: wgrad operands must be padded to multiples of 4");
    return B200FNO_EINVAL;
  }
  const long long chunks = (P + KC - 1) / KC;
  const int tiles = ceil_div(M, 64) * ceil_div(N, 64);
  const int gx = (int)std::max<long long>(1, std::min<long long>(chunks, std::max(1, 148 * 4 / tiles)));
  dim3 grid(gx, ceil_div(M, 64), ceil_div(N, 64));
  wgrad_kernel<<<grid, 256, 0, st>>>(A, lda, M, Mv, B, ldb, N, Nv, P, out, ldo, extra, bf16);
  B2_LAUNCHED("wgrad_kernel");
  return 0;
}

// out[n] += sum_p A[p*lda + n],  n < Nv
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int lda, int Nv, long long P,
                                                     float* __restrict__ out) {
  __shared__ float sh[256];
  const int nx = min(Nv, 64);  // columns per sweep handled by threadIdx.x % nx
  const int tx = threadIdx.x % nx, ty = threadIdx.x / nx, nty = 256 / nx;
  for (int n0 = 0; n0 < Nv; n0 += nx) {
    const int n = n0 + tx;
    float s = 0.f;
    if (ty < nty && n < Nv)
      for (long long p = (long long)blockIdx.x * nty + ty; p < P; p += (long long)gridDim.x * nty) s += __ldg(A + p * lda + n);
    sh[threadIdx.x] = s;
    __syncthreads();
    if (ty == 0 && n < Nv) {
      float t = 0.f;
      for (int y = 0; y < nty; ++y) t += sh[y * nx + tx];
      atomicAdd(out + n, t);
    }
    __syncthreads();
  }
}

int launch_colsum(const float* A, int lda, int Nv, long long P, float* out, cudaStream_t st) {
  if (P <= 0 || Nv <= 0) return 0;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((P + 63) / 64, 148 * 2));
  colsum_kernel<<<blocks, 256, 0, st>>>(A, lda, Nv, P, out);
  B2_LAUNCHED("colsum_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// Projection backward (fno.py:121-128 reversed) over the PADDED point grid so that every per-point
// buffer is flat [P][.] and the weight-gradient reductions need no cropping logic:
//   h1 = x.fc1^T + b1; g = gelu(h1);  dF = gather(dy) (0 at pad points);  dg = dF.fc2;  dh1 = dg*gelu'(h1);
//   dx = dh1.fc1          ->  G [P][128], dH [P][128], dF [P][Fp], dact [P][Cp]
// grid (padded rows (b,t,h), point tiles of 64 over Wp); block 256
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) proj_bwd_kernel(ProjBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* As = sm;                // [64][LDA]
  float* Bs = As + 64 * LDA;     // [KC][128]
  float* Hs = Bs + KC * PH;      // [64][LDH]  g, then dh1
  float* Gp = Hs + 64 * LDH;     // [64][LDH]  gelu'(h1)
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long row = blockIdx.x;
  const int h = (int)(row % a.Hp);
  const int t = (int)((row / a.Hp) % a.Tp);
  const int b = (int)(row / ((long long)a.Hp * a.Tp));
  const int p0 = blockIdx.y * 64;
  const int npts = min(64, a.Wp - p0);
  const long long P0 = row * a.Wp + p0;
  const bool live = (h < a.H) && (t < a.T) && (p0 < a.W);
  if (!live) {  // pad rows / pad tiles: every per-point output is zero
    for (int idx = tid; idx < npts * (PH / 4); idx += 256) {
      const int pp = idx / (PH / 4), cc = (idx % (PH / 4)) * 4;
      *reinterpret_cast<float4*>(a.G + (P0 + pp) * PH + cc) = zero4();
      *reinterpret_cast<float4*>(a.dH + (P0 + pp) * PH + cc) = zero4();
    }
    for (int idx = tid; idx < npts * (a.Fp / 4); idx += 256) {
      const int pp = idx / (a.Fp / 4), cc = (idx % (a.Fp / 4)) * 4;
      *reinterpret_cast<float4*>(a.dF + (P0 + pp) * a.Fp + cc) = zero4();
    }
    for (int idx = tid; idx < npts * (a.Cp / 4); idx += 256) {
      const int pp = idx / (a.Cp / 4), cc = (idx % (a.Cp / 4)) * 4;
      *reinterpret_cast<float4*>(a.dact + (P0 + pp) * a.Cp + cc) = zero4();
    }
    return;
  }
  const float* in_row = a.act + (size_t)row * a.Wp * a.Cp;
  // ---- GEMM 1: h1 = [64 x Cp] . [Cp x 128]
  float acc0[4][4] = {}, acc1[4][4] = {};
  for (int k0 = 0; k0 < a.Cp; k0 += KC) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256, pp = idx >> 3, kk = (idx & 7) * 4;
      const int p = p0 + pp, k = k0 + kk;
      float4 xv = (p < a.Wp && k < a.Cp) ? ldg4(in_row + (size_t)p * a.Cp + k) : zero4();
      if (a.bf16) xv = tr_bf16r4(xv);
      *reinterpret_cast<float4*>(As + pp * LDA + kk) = xv;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256, kb = idx >> 5, nn = (idx & 31) * 4;
      const int kr = k0 + kb;
      *reinterpret_cast<float4*>(Bs + kb * PH + nn) = (kr < a.Cp) ? ldg4(a.fc1T + (size_t)kr * PH + nn) : zero4();
    }
    __syncthreads();
    mma_chunk(acc0, As + (ty * 4) * LDA, LDA, Bs + tx * 4, PH);
    mma_chunk(acc1, As + (ty * 4) * LDA, LDA, Bs + 64 + tx * 4, PH);
    __syncthreads();
  }
  {
    const float4 b0 = ldg4(a.fc1b + tx * 4), b1 = ldg4(a.fc1b + 64 + tx * 4);
    const float bb0[4] = {b0.x, b0.y, b0.z, b0.w}, bb1[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float* hrow = Hs + (ty * 4 + r) * LDH;
      float* grow = Gp + (ty * 4 + r) * LDH;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float v0 = acc0[r][j] + bb0[j], v1 = acc1[r][j] + bb1[j];
        const float g0 = gelu_erf(v0), g1 = gelu_erf(v1);  // bf16 mode: the fc2 operand is a bf16 tensor (as in the forward)
        hrow[tx * 4 + j] = a.bf16 ? tr_bf16r(g0) : g0, grow[tx * 4 + j] = gelu_grad(v0);
        hrow[64 + tx * 4 + j] = a.bf16 ? tr_bf16r(g1) : g1, grow[64 + tx * 4 + j] = gelu_grad(v1);
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < npts * (PH / 4); idx += 256) {
    const int pp = idx / (PH / 4), cc = (idx % (PH / 4)) * 4;
    *reinterpret_cast<float4*>(a.G + (P0 + pp) * PH + cc) = *reinterpret_cast<const float4*>(Hs + pp * LDH + cc);
  }
  // ---- GEMM 2: dg = dF [64 x Fp] . fc2 [Fp x 128]
  const size_t pt_out = (size_t)b * a.out_sB + (size_t)t * a.out_sT + (size_t)h * a.W * a.c_out;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc0[r][j] = 0.f, acc1[r][j] = 0.f;
  for (int f0 = 0; f0 < a.Fp; f0 += KC) {
    for (int idx = tid; idx < 64 * KC; idx += 256) {
      const int pp = idx / KC, kk = idx % KC;
      const int w = p0 + pp, f = f0 + kk;
      float v = 0.f;
      if (w < a.W && f < a.Fout) v = __ldg(a.dy + pt_out + (size_t)w * a.c_out + a.out_off[f]);
      if (a.bf16) v = tr_bf16r(v);  // grad_output of fc2 (its output is a bf16 tensor under autocast)
      As[pp * LDA + kk] = v;
      if (pp < npts && f < a.Fp) a.dF[(P0 + pp) * a.Fp + f] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256, kb = idx >> 5, nn = (idx & 31) * 4;
      const int f = f0 + kb;
      *reinterpret_cast<float4*>(Bs + kb * PH + nn) = (f < a.Fp) ? ldg4(a.fc2W + (size_t)f * PH + nn) : zero4();
    }
    __syncthreads();
    mma_chunk(acc0, As + (ty * 4) * LDA, LDA, Bs + tx * 4, PH);
    mma_chunk(acc1, As + (ty * 4) * LDA, LDA, Bs + 64 + tx * 4, PH);
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float* hrow = Hs + (ty * 4 + r) * LDH;
    const float* grow = Gp + (ty * 4 + r) * LDH;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d0 = acc0[r][j] * grow[tx * 4 + j], d1 = acc1[r][j] * grow[64 + tx * 4 + j];
      hrow[tx * 4 + j] = a.bf16 ? tr_bf16r(d0) : d0;  // grad_output of fc1
      hrow[64 + tx * 4 + j] = a.bf16 ? tr_bf16r(d1) : d1;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < npts * (PH / 4); idx += 256) {
    const int pp = idx / (PH / 4), cc = (idx % (PH / 4)) * 4;
    *reinterpret_cast<float4*>(a.dH + (P0 + pp) * PH + cc) = *reinterpret_cast<const float4*>(Hs + pp * LDH + cc);
  }
  // ---- GEMM 3: dx = dh1 [64 x 128] . fc1 [128 x Cp], 64 channels per sweep
  for (int o0 = 0; o0 < a.Cp; o0 += TN) {
    float acc[4][4] = {};
    for (int k0 = 0; k0 < PH; k0 += KC) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = tid + i * 256, kb = idx >> 4, nn = (idx & 15) * 4;
        const int o = o0 + nn;
        *reinterpret_cast<float4*>(Bs + kb * TN + nn) = (o < a.Cp) ? ldg4(a.fc1W + (size_t)(k0 + kb) * a.Cp + o) : zero4();
      }
      __syncthreads();
      mma_chunk(acc, Hs + (ty * 4) * LDH + k0, LDH, Bs + tx * 4, TN);
      __syncthreads();
    }
    const int o = o0 + tx * 4;
    if (o < a.Cp) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int pp = ty * 4 + r;
        if (pp < npts)
          *reinterpret_cast<float4*>(a.dact + (P0 + pp) * a.Cp + o) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      }
    }
  }
}

constexpr size_t PROJ_BWD_SMEM = (size_t)(64 * LDA + KC * PH + 2 * 64 * LDH) * sizeof(float);

int launch_proj_bwd(const ProjBwdArgs& a, cudaStream_t st) {
  B2_CUDA(cudaFuncSetAttribute(proj_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PROJ_BWD_SMEM));
  dim3 grid((unsigned)((long long)a.B * a.Tp * a.Hp), ceil_div(a.Wp, 64));
  proj_bwd_kernel<<<grid, 256, PROJ_BWD_SMEM, st>>>(a);
  B2_LAUNCHED("proj_bwd_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// feat[p][j] = K row of the lift GEMM at padded point p: [input features | grid coords | 1 | 0 pad];
// all zero at pad points (fno.py:106-111: the pad region of the lifted tensor is constant zero).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lift_features_kernel(LiftArgs a, float* __restrict__ feat) {
  const long long total = (long long)a.B * a.Tp * a.Hp * a.Wp * a.Klp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % a.Klp);
    long long p = idx / a.Klp;
    const int w = (int)(p % a.Wp);
    p /= a.Wp;
    const int h = (int)(p % a.Hp);
    p /= a.Hp;
    const int t = (int)(p % a.Tp), b = (int)(p / a.Tp);
    float v = 0.f;
    if (w < a.W && h < a.H && t < a.T) {
      if (j < a.Fin)
        v = __ldg(a.x + (size_t)b * a.x_sB + (size_t)t * a.x_sT + ((size_t)h * a.W + w) * a.c_in + a.in_off[j]);
      else if (j < a.Fin + a.ng) {
        const int gi = j - a.Fin + (a.gt ? 0 : 1);
        v = gi == 0 ? a.gt[t] : (gi == 1 ? a.gh[h] : a.gw[w]);
      } else if (j == a.Fin + a.ng) v = 1.f;
    }
    feat[idx] = v;
  }
}

int launch_lift_features(const LiftArgs& a, float* feat, cudaStream_t st) {
  const long long total = (long long)a.B * a.Tp * a.Hp * a.Wp * a.Klp;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, 148 * 16));
  lift_features_kernel<<<blocks, 256, 0, st>>>(a, feat);
  B2_LAUNCHED("lift_features_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// spectral weights: adjoint pack, gradient, un-pack to the reference layout
// ---------------------------------------------------------------------------
// Wadj[mode][o][ri][i] = conj(W)[i][o]: feeding it to modes_kernel computes dS = dO (x) conj(W)^T,
// the adjoint of the forward mixing (real-linear in (re, im)).  32x32 tiles through shared memory:
// grid (mode, i-tile * o-tile, ri), block (32, 8); both the read and the write are 128-byte coalesced.
__global__ void __launch_bounds__(256) pack_spectral_adj_kernel(const float* __restrict__ Wpk, float* __restrict__ Wadj,
                                                                int Cp) {
  __shared__ float tile[32][33];
  const int nt = (Cp + 31) / 32;
  const int i0 = (blockIdx.y / nt) * 32, o0 = (blockIdx.y % nt) * 32, ri = blockIdx.z;
  const size_t base = (size_t)blockIdx.x * Cp * 2 * Cp;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, o = o0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < Cp && o < Cp) ? Wpk[base + ((size_t)i * 2 + ri) * Cp + o] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int o = o0 + r, i = i0 + threadIdx.x;
    if (o < Cp && i < Cp) {
      const float v = tile[threadIdx.x][r];
      Wadj[base + ((size_t)o * 2 + ri) * Cp + i] = ri ? -v : v;
    }
  }
}

int launch_pack_spectral_adj(const float* Wpk, float* Wadj, int NM, int Cp, cudaStream_t st) {
  const int nt = ceil_div(Cp, 32);
  pack_spectral_adj_kernel<<<dim3(NM, nt * nt, 2), dim3(32, 8), 0, st>>>(Wpk, Wadj, Cp);
  B2_LAUNCHED("pack_spectral_adj_kernel");
  return 0;
}

// dW[mode][i][0][o] = sum_b  Sr[b,i] dOr[b,o] + Si[b,i] dOi[b,o]
// dW[mode][i][1][o] = sum_b -Si[b,i] dOr[b,o] + Sr[b,i] dOi[b,o]
// (gradient w.r.t. the real and imaginary parts = torch's .grad of a complex parameter, viewed as real)
// S, dO: [B][2][NM][Cp]; one CTA per mode.
constexpr int MW_BCH = 32;
__global__ void __launch_bounds__(256) modes_wgrad_kernel(const float* __restrict__ S, const float* __restrict__ dO,
                                                          float* __restrict__ dW, int B, int NM, int Cp) {
  extern __shared__ __align__(16) float mw_sh[];  // S [MW_BCH][2][Cp] | dO [MW_BCH][2][Cp]
  float* Ss = mw_sh;
  float* Os = mw_sh + MW_BCH * 2 * Cp;
  const int mode = blockIdx.x, OQ = Cp >> 2, tid = threadIdx.x;
  float* dWm = dW + (size_t)mode * Cp * 2 * Cp;
  for (int b0 = 0; b0 < B; b0 += MW_BCH) {
    const int nb = min(MW_BCH, B - b0);
    __syncthreads();
    for (int idx = tid; idx < nb * 2 * OQ; idx += 256) {
      const int q = idx % OQ, ri = (idx / OQ) & 1, bb = idx / (2 * OQ);
      const size_t g = (((size_t)(b0 + bb) * 2 + ri) * NM + mode) * Cp + q * 4;
      *reinterpret_cast<float4*>(Ss + (bb * 2 + ri) * Cp + q * 4) = ldg4(S + g);
      *reinterpret_cast<float4*>(Os + (bb * 2 + ri) * Cp + q * 4) = ldg4(dO + g);
    }
    __syncthreads();
    for (int item = tid; item < Cp * OQ; item += 256) {
      const int q = item % OQ, i = item / OQ;
      float4 wr = zero4(), wi = zero4();
      for (int bb = 0; bb < nb; ++bb) {
        const float sr = Ss[(bb * 2) * Cp + i], si = Ss[(bb * 2 + 1) * Cp + i];
        const float4 orr = *reinterpret_cast<const float4*>(Os + (bb * 2) * Cp + q * 4);
        const float4 oi = *reinterpret_cast<const float4*>(Os + (bb * 2 + 1) * Cp + q * 4);
        wr.x = fmaf(sr, orr.x, wr.x), wr.y = fmaf(sr, orr.y, wr.y), wr.z = fmaf(sr, orr.z, wr.z), wr.w = fmaf(sr, orr.w, wr.w);
        wr.x = fmaf(si, oi.x, wr.x), wr.y = fmaf(si, oi.y, wr.y), wr.z = fmaf(si, oi.z, wr.z), wr.w = fmaf(si, oi.w, wr.w);
        wi.x = fmaf(sr, oi.x, wi.x), wi.y = fmaf(sr, oi.y, wi.y), wi.z = fmaf(sr, oi.z, wi.z), wi.w = fmaf(sr, oi.w, wi.w);
        wi.x = fmaf(-si, orr.x, wi.x), wi.y = fmaf(-si, orr.y, wi.y), wi.z = fmaf(-si, orr.z, wi.z), wi.w = fmaf(-si, orr.w, wi.w);
      }
      float4* pr = reinterpret_cast<float4*>(dWm + ((size_t)i * 2 + 0) * Cp + q * 4);
      float4* pi = reinterpret_cast<float4*>(dWm + ((size_t)i * 2 + 1) * Cp + q * 4);
      if (b0 > 0) {  // later batch passes accumulate (this thread owns the element)
        const float4 r0 = *pr, i0 = *pi;
        wr.x += r0.x, wr.y += r0.y, wr.z += r0.z, wr.w += r0.w;
        wi.x += i0.x, wi.y += i0.y, wi.z += i0.z, wi.w += i0.w;
      }
      *pr = wr, *pi = wi;
    }
  }
}

int launch_modes_wgrad(const float* S, const float* dO, float* dW, int B, int NM, int Cp, cudaStream_t st) {
  const size_t smem = (size_t)2 * MW_BCH * 2 * Cp * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("modes wgrad: width %d too large", Cp);
    return B200FNO_EINVAL;
  }
  B2_CUDA(cudaFuncSetAttribute(modes_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  modes_wgrad_kernel<<<NM, 256, smem, st>>>(S, dO, dW, B, NM, Cp);
  B2_LAUNCHED("modes_wgrad_kernel");
  return 0;
}

// Inverse of pack_spectral_kernel: gradient of corner tensor `corner` in the reference layout
// complex64 [ci][co][m1][m2][m3].  A corner element whose frequency was overwritten by a later corner
// assignment (fno.py:53-60, overlapping corners) never reaches the output: its gradient is zero.
__global__ void __launch_bounds__(256) unpack_spectral_grad_kernel(const float* __restrict__ dWpk, float* __restrict__ g,
                                                                   int corner, int ndim, int Tp, int Hp, int m1, int m2,
                                                                   int m3, int KH, int ci, int co, int Cp,
                                                                   const int* __restrict__ slot_t,
                                                                   const int* __restrict__ slot_h) {
  // one CTA per (input channel i, corner element (x, y)): rows dWpk[mode(kw)][i][ri][0..co) in, rows
  // g[i][o][x][y][0..m3)[re,im] out, transposed through shared memory (both sides coalesced)
  extern __shared__ float tile[];  // [2*m3][co + 1]
  const int mm1 = ndim == 3 ? m1 : 1;
  const int i = blockIdx.y, y = blockIdx.x % m2, x = blockIdx.x / m2;
  const int W2 = 2 * m3, ldt = co + 1;
  const bool h_hi = ndim == 3 ? (corner >= 2) : (corner == 1);
  const bool t_hi = ndim == 3 ? (corner & 1) : false;
  const int fH = h_hi ? Hp - m2 + y : y;
  const int fT = ndim == 3 ? (t_hi ? Tp - m1 + x : x) : 0;
  // winner rule of pack_spectral_kernel: "high" owns every frequency >= N - m
  const bool own = ((fH >= Hp - m2) == h_hi) && (ndim == 3 ? ((fT >= Tp - m1) == t_hi) : true);
  if (own) {
    const size_t mode0 = ((size_t)(ndim == 3 ? slot_t[fT] : 0) * KH + slot_h[fH]) * m3;
    for (int idx = threadIdx.x; idx < W2 * co; idx += blockDim.x) {
      const int o = idx % co, k = idx / co, kw = k >> 1, ri = k & 1;
      tile[k * ldt + o] = dWpk[(((mode0 + kw) * Cp + i) * 2 + ri) * Cp + o];
    }
  }
  __syncthreads();
  const size_t row_stride = (size_t)mm1 * m2 * m3 * 2;
  const size_t base = ((size_t)i * co * mm1 * m2 + (size_t)x * m2 + y) * m3 * 2;
  for (int idx = threadIdx.x; idx < co * W2; idx += blockDim.x) {
    const int o = idx / W2, k = idx % W2;
    g[base + (size_t)o * row_stride + k] = own ? tile[k * ldt + o] : 0.f;
  }
}

int launch_unpack_spectral_grad(const float* dWpk, float* const* corners, int ncorner, const Geom& g, int ci, int co,
                                int m1, int m2, const int* slot_t, const int* slot_h, cudaStream_t st) {
  const size_t smem = (size_t)2 * g.m3 * (co + 1) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("spectral gradient unpack: width %d x modes3 %d too large", co, g.m3);
    return B200FNO_EINVAL;
  }
  B2_CUDA(cudaFuncSetAttribute(unpack_spectral_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((g.ndim == 3 ? m1 : 1) * m2, ci);
  for (int c = 0; c < ncorner; ++c) {
    if (!corners[c]) continue;
    unpack_spectral_grad_kernel<<<grid, 256, smem, st>>>(dWpk, corners[c], c, g.ndim, g.Tp, g.Hp, m1, m2, g.m3, g.KH, ci,
                                                         co, g.Cp, slot_t, slot_h);
    B2_LAUNCHED("unpack_spectral_grad_kernel");
  }
  return 0;
}

// ---------------------------------------------------------------------------
// Lift backward w.r.t. the input field (fno.py:106-109 reversed): dx[b,t,h,w,c] = sum_o d act0[p][o] * fc0_w[o][f],
// f = the lift feature the input element feeds (LiftArgs::in_off maps feature -> offset inside a point's input).
// The grid-coordinate columns of fc0 carry no gradient.  grid (valid rows (b,t,h), 64-point tiles over W); block 256.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lift_bwd_input_kernel(LiftArgs a, const float* __restrict__ dact,
                                                             float* __restrict__ dx) {
  extern __shared__ __align__(16) float lb_sh[];  // dact tile [64][Cp + 1]
  const int ld = a.Cp + 1;
  const long long row = blockIdx.x;
  const int h = (int)(row % a.H), t = (int)((row / a.H) % a.T), b = (int)(row / ((long long)a.H * a.T));
  const int p0 = blockIdx.y * 64, npts = min(64, a.W - p0);
  const float* drow = dact + ((((size_t)b * a.Tp + t) * a.Hp + h) * a.Wp + p0) * a.Cp;
  for (int idx = threadIdx.x; idx < npts * a.Cp; idx += 256)
    lb_sh[(idx / a.Cp) * ld + idx % a.Cp] = a.bf16 ? tr_bf16r(drow[idx]) : drow[idx];
  __syncthreads();
  float* xb = dx + (size_t)b * a.x_sB + (size_t)t * a.x_sT + ((size_t)h * a.W + p0) * a.c_in;
  for (int item = threadIdx.x; item < npts * a.Fin; item += 256) {
    const int pp = item % npts, f = item / npts;  // consecutive threads: consecutive points, same feature (weights broadcast)
    const float* w = a.W0T + (size_t)f * a.Cp;    // W0T[f][o] = fc0_w[o][f]
    const float* d = lb_sh + pp * ld;
    float acc = 0.f;
    for (int o = 0; o < a.Cp; ++o) acc = fmaf(d[o], __ldg(w + o), acc);
    xb[(size_t)pp * a.c_in + a.in_off[f]] = acc;
  }
}

int launch_lift_bwd_input(const LiftArgs& a, const float* dact, float* dx, cudaStream_t st) {
  const size_t smem = (size_t)64 * (a.Cp + 1) * sizeof(float);
  B2_CUDA(cudaFuncSetAttribute(lift_bwd_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((long long)a.B * a.T * a.H), ceil_div(a.W, 64));
  lift_bwd_input_kernel<<<grid, 256, smem, st>>>(a, dact, dx);
  B2_LAUNCHED("lift_bwd_input_kernel");
  return 0;
}

// ---------------------------------------------------------------------------
// Fused Adam step (torch.optim.Adam defaults: no weight decay, no amsgrad) in ONE pass over p, g, m, v.
// Same elementwise formulas, in the same order, as torch's foreach implementation:
//   m <- m + (1-b1)(g - m);  v <- v*b2 + (1-b2) g*g;  p <- p - step_size * (m / (sqrt(v)/sqrt(bc2) + eps))
// Complex parameters are passed as their real views (re and im are independent reals, as in torch).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n, float w1,
                                                   float b2, float w2, float step_size, float bc2_sqrt, float eps) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i];
    const float mi = fmaf(w1, gi - m[i], m[i]);
    const float vi = fmaf(w2 * gi, gi, v[i] * b2);
    m[i] = mi, v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                long long step, cudaStream_t st) {
  if (n <= 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const int blocks = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 16));
  adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, 1.0f - beta1, beta2, 1.0f - beta2, (float)((double)lr / bc1),
                                      (float)sqrt(bc2), eps);
  B2_LAUNCHED("adam_kernel");
  return 0;
}

// dst[r][c] = src[r][c] for r < rows, c < cols; zero elsewhere (dst is [dst_rows][dst_cols])
__global__ void pad2d_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst, int dst_rows,
                             int dst_cols) {
  const int total = dst_rows * dst_cols;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx % dst_cols, r = idx / dst_cols;
    dst[idx] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.f;
  }
}
int launch_pad2d(const float* src, int rows, int cols, float* dst, int dst_rows, int dst_cols, cudaStream_t st) {
  const int total = dst_rows * dst_cols;
  pad2d_kernel<<<std::max(1, std::min((total + 255) / 256, 1024)), 256, 0, st>>>(src, rows, cols, dst, dst_rows, dst_cols);
  B2_LAUNCHED("pad2d_kernel");
  return 0;
}

}  // namespace b200fno
