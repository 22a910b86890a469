// One-time re-layout of the reference parameters (state_dict layout, SURVEY
// section 5) into the engine layout.  Runs on the caller's stream.
#include <algorithm>

#include "common.cuh"

namespace b200fno {

struct Corners {
  const float* w[4];
};

// Wpk[mode][i][ri][o], mode = (kt_slot*KH + kh_slot)*m3 + kw.
// Source corner tensors are complex64 [ci][co][m1][m2][m3] (3-D) / [ci][co][m2][m3] (2-D).
// Which corner feeds a kept frequency follows the assignment order fno.py:53-60:
// weights1 (low t, low h), weights2 (high t, low h), weights3 (low t, high h),
// weights4 (high t, high h); where corners overlap the later assignment wins,
// i.e. "high" beats "low" on each axis.
// One CTA per (kept (T,H) frequency slot, input channel i): the source rows src[i][o][x][y][0..m3)[re,im] are
// 2*m3 contiguous floats per output channel o, the destination rows Wpk[mode(kw)][i][ri][0..Cp) are contiguous in o:
// transposed through shared memory so that both sides are coalesced (this runs after every optimiser step).
__global__ void __launch_bounds__(256) pack_spectral_kernel(Corners c, float* __restrict__ Wpk, int ndim, int Tp, int Hp,
                                                            int m1, int m2, int m3, int KH, int ci, int co, int Cp,
                                                            const int* __restrict__ ft, const int* __restrict__ fh,
                                                            int m3s, int kw0) {  // source W modes, first one taken
  extern __shared__ float tile[];  // [Cp][2*m3 + 1]
  const int slot = blockIdx.x, i = blockIdx.y;
  const int khs = slot % KH, kts = slot / KH;
  const int W2 = 2 * m3, ldt = W2 + 1;
  const int fH = fh[khs];
  const bool h_hi = fH >= Hp - m2;
  const int y = h_hi ? fH - (Hp - m2) : fH;
  const float* src = nullptr;
  size_t row_stride = 0, base = 0;  // element (o, kw, ri) at src[base + o*row_stride + kw*2 + ri]
  if (i < ci) {
    if (ndim == 3) {
      const int fT = ft[kts];
      const bool t_hi = fT >= Tp - m1;
      const int x = t_hi ? fT - (Tp - m1) : fT;
      src = c.w[(h_hi ? 2 : 0) + (t_hi ? 1 : 0)];
      row_stride = (size_t)m1 * m2 * m3s * 2;
      base = ((size_t)i * co * m1 * m2 + (size_t)x * m2 + y) * m3s * 2 + (size_t)kw0 * 2;
    } else {
      src = c.w[h_hi ? 1 : 0];
      row_stride = (size_t)m2 * m3s * 2;
      base = ((size_t)i * co * m2 + y) * m3s * 2 + (size_t)kw0 * 2;
    }
  }
  for (int idx = threadIdx.x; idx < Cp * W2; idx += blockDim.x) {
    const int o = idx / W2, k = idx % W2;
    tile[o * ldt + k] = (src && o < co) ? src[base + (size_t)o * row_stride + k] : 0.f;
  }
  __syncthreads();
  const size_t mode0 = (size_t)slot * m3;
  for (int idx = threadIdx.x; idx < W2 * Cp; idx += blockDim.x) {
    const int o = idx % Cp, k = idx / Cp, kw = k >> 1, ri = k & 1;
    Wpk[(((mode0 + kw) * Cp + i) * 2 + ri) * Cp + o] = tile[o * ldt + k];
  }
}

int launch_pack_spectral(const float* const* corners, int ncorner, float* Wpk, const Geom& g, int ci, int co, int m1,
                         int m2, const int* d_ft, const int* d_fh, cudaStream_t st, int m3_src, int kw0) {
  Corners c{};
  for (int i = 0; i < ncorner; ++i) c.w[i] = corners[i];
  const size_t smem = (size_t)g.Cp * (2 * g.m3 + 1) * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("spectral pack: width %d x modes3 %d too large", g.Cp, g.m3);
    return B200FNO_EINVAL;
  }
  B2_CUDA(cudaFuncSetAttribute(pack_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pack_spectral_kernel<<<dim3(g.KT * g.KH, g.Cp), 256, smem, st>>>(c, Wpk, g.ndim, g.Tp, g.Hp, m1, m2, g.m3, g.KH, ci, co,
                                                                   g.Cp, d_ft, d_fh, m3_src > 0 ? m3_src : g.m3, kw0);
  B2_LAUNCHED("pack_spectral_kernel");
  return 0;
}

// dst[c][r] = src[r][c] for r < rows, c < cols; zero elsewhere.  dst is [dst_rows][dst_cols].
__global__ void transpose_pad_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst,
                                     int dst_rows, int dst_cols) {
  const int total = dst_rows * dst_cols;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int r = idx % dst_cols, c = idx / dst_cols;  // dst[c][r]
    dst[idx] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.f;
  }
}
int launch_transpose_pad(const float* src, int rows, int cols, float* dst, int dst_rows, int dst_cols,
                         cudaStream_t st) {
  int total = dst_rows * dst_cols;
  transpose_pad_kernel<<<std::max(1, std::min((total + 255) / 256, 1024)), 256, 0, st>>>(src, rows, cols, dst,
                                                                                          dst_rows, dst_cols);
  B2_LAUNCHED("transpose_pad_kernel");
  return 0;
}

// conv bias + eval BatchNorm (fno.py:115-117) as y = v*scale + shift
__global__ void fold_bn_kernel(const float* conv_b, const float* bn_w, const float* bn_b, const float* bn_m,
                               const float* bn_v, float eps, int C, int Cp, float* scale, float* shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  if (c < C) {
    float s = bn_w[c] / sqrtf(bn_v[c] + eps);
    scale[c] = s;
    shift[c] = (conv_b[c] - bn_m[c]) * s + bn_b[c];
  } else {
    scale[c] = 0.f;
    shift[c] = 0.f;
  }
}
int launch_fold_bn(const float* conv_b, const float* bn_w, const float* bn_b, const float* bn_m, const float* bn_v,
                   float eps, int C, int Cp, float* scale, float* shift, cudaStream_t st) {
  fold_bn_kernel<<<ceil_div(Cp, 128), 128, 0, st>>>(conv_b, bn_w, bn_b, bn_m, bn_v, eps, C, Cp, scale, shift);
  B2_LAUNCHED("fold_bn_kernel");
  return 0;
}

__global__ void pad_copy_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int np) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) dst[i] = i < n ? src[i] : 0.f;
}
int launch_pad_copy(const float* src, int n, float* dst, int np, cudaStream_t st) {
  pad_copy_kernel<<<ceil_div(np, 256), 256, 0, st>>>(src, n, dst, np);
  B2_LAUNCHED("pad_copy_kernel");
  return 0;
}

// Lift weights for the tensor-core kernel, K-major [64 ch][64 k]: k < Fin input features, the last four
// columns of the last K step (k = nkl*8-4 ..) = grid t, grid h, grid w, bias (fc0 columns Fin.., fno.py:107-108)
__global__ void pack_w0k_kernel(const float* __restrict__ fc0_w, const float* __restrict__ fc0_b, int C, int Fin,
                                int ng, int nkl, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * 64) return;
  const int c = idx >> 6, k = idx & 63, e0 = nkl * 8 - 4, ld = Fin + ng;
  float v = 0.f;
  if (c < C) {
    if (k < Fin) v = fc0_w[(size_t)c * ld + k];
    else if (k >= e0 && k < e0 + 3) {
      const int gi = k - e0 - (3 - ng);  // ng == 2: (h, w) only, the t slot stays zero
      if (gi >= 0) v = fc0_w[(size_t)c * ld + Fin + gi];
    } else if (k == e0 + 3) v = fc0_b[c];
  }
  out[idx] = v;
}
int launch_pack_w0k(const float* fc0_w, const float* fc0_b, int C, int Fin, int ng, int nkl, float* out,
                    cudaStream_t st) {
  pack_w0k_kernel<<<16, 256, 0, st>>>(fc0_w, fc0_b, C, Fin, ng, nkl, out);
  B2_LAUNCHED("pack_w0k_kernel");
  return 0;
}

__device__ __forceinline__ float round_bf16(float x) {  // nearest-even, as tc::bf16_rn
  const uint32_t u = __float_as_uint(x);
  return __uint_as_float((u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u);
}
// bf16 != 0 (bf16 compute mode): hi = the weight cast to bf16, lo = 0 (the kernels run a single pass)
__global__ void split_hl_kernel(const float* __restrict__ src, int n, float* __restrict__ hi, float* __restrict__ lo,
                                int bf16) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float x = src[i];
    float h = bf16 ? round_bf16(x) : __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);  // rn, as tc::tf32_hi
    hi[i] = h;
    lo[i] = bf16 ? 0.f : x - h;
  }
}
int launch_split_hl(const float* src, int n, float* dst_hi, float* dst_lo, cudaStream_t st, int bf16) {
  split_hl_kernel<<<ceil_div(n, 256), 256, 0, st>>>(src, n, dst_hi, dst_lo, bf16);
  B2_LAUNCHED("split_hl_kernel");
  return 0;
}
__global__ void round_bf16_kernel(float* __restrict__ p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = round_bf16(p[i]);
}
int launch_round_bf16(float* p, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  round_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n);
  B2_LAUNCHED("round_bf16_kernel");
  return 0;
}

}  // namespace b200fno
