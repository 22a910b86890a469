// eval_metrics on the GPU (reference realpdebench/utils/metrics.py:24-131): the 13 scalars the
// evaluation loop logs per chunk of predictions.  The reference bins |FFT(pred - target)|^2 by radial wavenumber with
// two Python triple loops over (t/2, h/2, w/2) (:75-81, :93-99) that skip every bin beyond nb = min(t,h,w)/2 - so only
// wavenumbers i, j, k < nb are ever used and the full fftn is, once more, a TRUNCATED separable DFT (here complex
// output of a real field, three axes).  Everything is fp32 arithmetic with double accumulators for the reductions.
//
//   pointwise_kernel   per-sample sum (p-g)^2, |p-g|, g^2;  per-(sample, frame) sum of (p-g)      one pass over p, g
//   r2_kernel          sum over positions of the batch variance of the target (:59)
//   ke_kernel          |KE(pred) - KE(target)| with KE = half the temporal variance of u, v (:15-22, :62-67)
//   dftw / dft_axis    truncated DFT along W, then H, then T of (p - g) and of g, nb modes per axis
//   bin_kernel         radial binning of |F|^2 (:75-81, :93-99)
//   finalize_kernel    the 13 outputs (incl. the length-t DFT of the frame sums for freq_error, :107-112)
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace b200fno {

namespace {

struct MetricsDims {
  int b, t, h, w, ct, c, nb;  // ct: channels in memory, c: channels evaluated
};

__device__ __forceinline__ double block_sum(double v, double* sh) {  // blockDim.x == 256
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 8) r = sh[threadIdx.x];
  if (warp == 0)
    for (int o = 4; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  return r;  // valid in thread 0
}

// twiddle tables: tw[n] = (cos, -sin)(2 pi n / N) for each axis length
__global__ void tables_kernel(float2* twW, int W, float2* twH, int H, float2* twT, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W) {
    double s, c;
    sincospi(2.0 * i / W, &s, &c);
    twW[i] = make_float2((float)c, (float)-s);
  }
  if (i < H) {
    double s, c;
    sincospi(2.0 * i / H, &s, &c);
    twH[i] = make_float2((float)c, (float)-s);
  }
  if (i < T) {
    double s, c;
    sincospi(2.0 * i / T, &s, &c);
    twT[i] = make_float2((float)c, (float)-s);
  }
}

// grid (hw chunks, t, b); acc[b][0..2] += (sum d^2, sum |d|, sum g^2); sig[b][t] += sum d
__global__ void __launch_bounds__(256) pointwise_kernel(const float* __restrict__ p, const float* __restrict__ g,
                                                        MetricsDims d, double* __restrict__ acc, double* __restrict__ sig) {
  __shared__ double sh[8];
  const int tt = blockIdx.y, b = blockIdx.z, hw = d.h * d.w;
  const size_t base = ((size_t)b * d.t + tt) * hw * d.ct;
  double sq = 0.0, ab = 0.0, gg = 0.0, sd = 0.0;
  for (int pt = blockIdx.x * 256 + threadIdx.x; pt < hw; pt += gridDim.x * 256) {
    float fsq = 0.f, fab = 0.f, fgg = 0.f, fsd = 0.f;
    for (int cc = 0; cc < d.c; ++cc) {
      const float pv = __ldg(p + base + (size_t)pt * d.ct + cc), gv = __ldg(g + base + (size_t)pt * d.ct + cc);
      const float e = pv - gv;
      fsq = fmaf(e, e, fsq), fab += fabsf(e), fgg = fmaf(gv, gv, fgg), fsd += e;
    }
    sq += fsq, ab += fab, gg += fgg, sd += fsd;
  }
  sq = block_sum(sq, sh), ab = block_sum(ab, sh), gg = block_sum(gg, sh), sd = block_sum(sd, sh);
  if (threadIdx.x == 0) {
    atomicAdd(acc + b * 4 + 0, sq), atomicAdd(acc + b * 4 + 1, ab), atomicAdd(acc + b * 4 + 2, gg);
    atomicAdd(sig + (size_t)b * d.t + tt, sd);
  }
}

// out[0] += sum over (position, channel) of sum_b (g - mean_b g)^2 = sum_b g^2 - (sum_b g)^2 / B
__global__ void __launch_bounds__(256) r2_kernel(const float* __restrict__ g, MetricsDims d, double* __restrict__ out) {
  __shared__ double sh[8];
  const size_t npos = (size_t)d.t * d.h * d.w, stride_b = npos * d.ct;
  double tot = 0.0;
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < npos * d.c; i += (size_t)gridDim.x * 256) {
    const size_t pos = i / d.c;
    const int cc = (int)(i % d.c);
    double s = 0.0, ss = 0.0;
    for (int b = 0; b < d.b; ++b) {
      const double v = (double)__ldg(g + (size_t)b * stride_b + pos * d.ct + cc);
      s += v, ss += v * v;
    }
    tot += ss - s * s / d.b;
  }
  tot = block_sum(tot, sh);
  if (threadIdx.x == 0) atomicAdd(out, tot);
}

// out[0] += sum over (b,h,w) of |KE(pred) - KE(target)|, KE = 0.5 (var_t u + var_t v)  (population variance)
__global__ void __launch_bounds__(256) ke_kernel(const float* __restrict__ p, const float* __restrict__ g, MetricsDims d,
                                                 double* __restrict__ out) {
  __shared__ double sh[8];
  const int hw = d.h * d.w;
  double tot = 0.0;
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < (size_t)d.b * hw; i += (size_t)gridDim.x * 256) {
    const int b = (int)(i / hw), pt = (int)(i % hw);
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};  // (pred u, pred v, target u, target v)
    for (int tt = 0; tt < d.t; ++tt) {
      const size_t at = (((size_t)b * d.t + tt) * hw + pt) * d.ct;
      const double pu = __ldg(p + at), pv = __ldg(p + at + 1), gu = __ldg(g + at), gv = __ldg(g + at + 1);
      s[0] += pu, q[0] += pu * pu, s[1] += pv, q[1] += pv * pv;
      s[2] += gu, q[2] += gu * gu, s[3] += gv, q[3] += gv * gv;
    }
    double var[4];
    for (int k = 0; k < 4; ++k) var[k] = q[k] / d.t - (s[k] / d.t) * (s[k] / d.t);
    tot += fabs(0.5 * (var[0] + var[1]) - 0.5 * (var[2] + var[3]));
  }
  tot = block_sum(tot, sh);
  if (threadIdx.x == 0) atomicAdd(out, tot);
}

// Truncated DFT along W of the two fields (0: p - g, 1: g).  One block per (b,t,h) row.
//   AW[f][row][k][cc] = sum_w x_f[row][w][cc] * tw[(k*w) mod W],  k < nb
__global__ void __launch_bounds__(256) dftw_kernel(const float* __restrict__ p, const float* __restrict__ g, MetricsDims d,
                                                   const float2* __restrict__ tw, float2* __restrict__ AW, size_t rows) {
  extern __shared__ __align__(16) float ms[];
  float2* stw = reinterpret_cast<float2*>(ms);  // [W]
  float* x0 = ms + 2 * d.w;                     // [W][c]  p - g
  float* x1 = x0 + d.w * d.c;                   // [W][c]  g
  const size_t row = blockIdx.x;
  const size_t base = row * d.w * d.ct;
  for (int i = threadIdx.x; i < d.w; i += 256) stw[i] = tw[i];
  for (int i = threadIdx.x; i < d.w * d.c; i += 256) {
    const int wv = i / d.c, cc = i % d.c;
    const float pv = __ldg(p + base + (size_t)wv * d.ct + cc), gv = __ldg(g + base + (size_t)wv * d.ct + cc);
    x0[i] = pv - gv, x1[i] = gv;
  }
  __syncthreads();
  const int nout = 2 * d.nb * d.c;
  for (int o = threadIdx.x; o < nout; o += 256) {
    const int cc = o % d.c, k = (o / d.c) % d.nb, f = o / (d.c * d.nb);
    const float* x = f ? x1 : x0;
    float re = 0.f, im = 0.f;
    int ang = 0;
    for (int wv = 0; wv < d.w; ++wv) {
      const float v = x[wv * d.c + cc];
      const float2 e = stw[ang];
      re = fmaf(v, e.x, re), im = fmaf(v, e.y, im);
      ang += k;
      if (ang >= d.w) ang -= d.w;
    }
    AW[((f * rows + row) * d.nb + k) * d.c + cc] = make_float2(re, im);
  }
}

// Complex truncated DFT along an axis of length N:  out[f][o][m][e] = sum_n in[f][o][n][e] * tw[(m*n) mod N], m < nb
__global__ void __launch_bounds__(256) dft_axis_kernel(const float2* __restrict__ in, float2* __restrict__ out, size_t outer,
                                                       int N, size_t inner, int nb, const float2* __restrict__ tw) {
  const size_t total = 2 * outer * nb * inner;
  for (size_t idx = blockIdx.x * (size_t)256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const size_t e = idx % inner;
    const int m = (int)((idx / inner) % nb);
    const size_t fo = idx / (inner * nb);  // f * outer + o
    const float2* src = in + fo * N * inner + e;
    float re = 0.f, im = 0.f;
    int ang = 0;
    for (int n = 0; n < N; ++n) {
      const float2 v = __ldg(src + (size_t)n * inner);
      const float2 w = __ldg(tw + ang);
      re = fmaf(v.x, w.x, fmaf(-v.y, w.y, re));
      im = fmaf(v.x, w.y, fmaf(v.y, w.x, im));
      ang += m;
      if (ang >= N) ang -= N;
    }
    out[idx] = make_float2(re, im);
  }
}

// spec[f][b][it][cc] += |F[f][b][i][j][k][cc]|^2,  it = floor(sqrt(i^2+j^2+k^2)) if it < nb
__global__ void __launch_bounds__(256) bin_kernel(const float2* __restrict__ F, MetricsDims d, double* __restrict__ spec) {
  const size_t per_b = (size_t)d.nb * d.nb * d.nb * d.c, total = 2 * (size_t)d.b * per_b;
  for (size_t idx = blockIdx.x * (size_t)256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const int cc = (int)(idx % d.c);
    size_t r = idx / d.c;
    const int k = (int)(r % d.nb);
    r /= d.nb;
    const int j = (int)(r % d.nb);
    r /= d.nb;
    const int i = (int)(r % d.nb);
    const size_t fb = r / d.nb;  // f * b + b
    if (i >= d.t / 2 || j >= d.h / 2 || k >= d.w / 2) continue;  // loop bounds of metrics.py:75-77
    const int s2 = i * i + j * j + k * k;
    int it = (int)floor(sqrt((double)s2));
    while (it * it > s2) --it;
    while ((it + 1) * (it + 1) <= s2) ++it;
    if (it > d.nb - 1) continue;
    const float2 v = F[idx];
    atomicAdd(spec + (fb * d.nb + it) * d.c + cc, (double)(v.x * v.x + v.y * v.y));
  }
}

__global__ void __launch_bounds__(256) finalize_kernel(MetricsDims d, const double* __restrict__ acc,
                                                       const double* __restrict__ sig, const double* __restrict__ misc,
                                                       const double* __restrict__ spec, const float2* __restrict__ twT,
                                                       int i_low, int i_high, float* __restrict__ out) {
  __shared__ double sh[8];
  // freq_error: mean over (b, f) of | sum_tau sig[b][tau] e^{-2 pi i f tau / t} |   (all t frequencies, metrics.py:107-112)
  double fsum = 0.0;
  for (int idx = threadIdx.x; idx < d.b * d.t; idx += 256) {
    const int b = idx / d.t, f = idx % d.t;
    double re = 0.0, im = 0.0;
    int ang = 0;
    for (int tau = 0; tau < d.t; ++tau) {
      const double v = sig[(size_t)b * d.t + tau];
      const float2 w = twT[ang];
      re += v * w.x, im += v * w.y;
      ang += f;
      if (ang >= d.t) ang -= d.t;
    }
    fsum += sqrt(re * re + im * im);
  }
  fsum = block_sum(fsum, sh);
  if (threadIdx.x != 0) return;
  const double n_elem = (double)d.b * d.t * d.h * d.w * d.c, thw = (double)d.t * d.h * d.w;
  double sq = 0.0, ab = 0.0, rel = 0.0;
  for (int b = 0; b < d.b; ++b) {
    sq += acc[b * 4 + 0], ab += acc[b * 4 + 1];
    rel += sqrt(acc[b * 4 + 0]) / sqrt(acc[b * 4 + 2]);
  }
  out[0] = (float)sqrt(sq / n_elem);
  out[1] = (float)(ab / n_elem);
  out[2] = (float)(rel / d.b);
  out[3] = (float)(1.0 - sq / misc[0]);
  out[4] = d.c < 2 ? 0.f : (float)(misc[1] / ((double)d.b * d.h * d.w));
  // spectra: err[it][cc] = sqrt(mean_b spec0) / (t h w), nrm likewise from spec1
  double s_all = 0.0, s_lo = 0.0, s_mid = 0.0, s_hi = 0.0, r_lo = 0.0, r_mid = 0.0, r_hi = 0.0;
  for (int it = 0; it < d.nb; ++it)
    for (int cc = 0; cc < d.c; ++cc) {
      double e = 0.0, n = 0.0;
      for (int b = 0; b < d.b; ++b) {
        e += spec[((size_t)b * d.nb + it) * d.c + cc];
        n += spec[(((size_t)d.b + b) * d.nb + it) * d.c + cc];
      }
      const double err = sqrt(e / d.b) / thw, nrm = sqrt(n / d.b) / thw, ratio = err / nrm;
      s_all += err;
      if (it < i_low) s_lo += err, r_lo += ratio;
      else if (it < i_high) s_mid += err, r_mid += ratio;
      else s_hi += err, r_hi += ratio;
    }
  const double n_lo = (double)i_low * d.c, n_mid = (double)(i_high - i_low) * d.c, n_hi = (double)(d.nb - i_high) * d.c;
  out[5] = (float)(s_all / ((double)d.nb * d.c));
  out[6] = (float)(s_lo / n_lo), out[7] = (float)(s_mid / n_mid), out[8] = (float)(s_hi / n_hi);  // empty slice: 0/0 = NaN
  out[9] = (float)(r_lo / n_lo), out[10] = (float)(r_mid / n_mid), out[11] = (float)(r_hi / n_hi);
  out[12] = (float)(fsum / ((double)d.b * d.t));
}

struct MetricsLayout {
  size_t tw, acc, sig, misc, spec, AW, AH, F, total;  // byte offsets
};
MetricsLayout metrics_layout(const MetricsDims& d) {
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  MetricsLayout L;
  size_t o = 0;
  L.tw = o, o += up((size_t)(d.w + d.h + d.t) * sizeof(float2));
  L.acc = o, o += up((size_t)d.b * 4 * sizeof(double));
  L.sig = o, o += up((size_t)d.b * d.t * sizeof(double));
  L.misc = o, o += up(4 * sizeof(double));
  L.spec = o, o += up((size_t)2 * d.b * d.nb * d.c * sizeof(double));
  const size_t zero_end = o;
  (void)zero_end;
  L.AW = o, o += up((size_t)2 * d.b * d.t * d.h * d.nb * d.c * sizeof(float2));
  L.AH = o, o += up((size_t)2 * d.b * d.t * d.nb * d.nb * d.c * sizeof(float2));
  L.F = o, o += up((size_t)2 * d.b * d.nb * d.nb * d.nb * d.c * sizeof(float2));
  L.total = o;
  return L;
}

int make_dims(int b, int t, int h, int w, int ct, int c, MetricsDims* d) {
  if (b < 1 || t < 2 || h < 2 || w < 2 || ct < 1 || c < 1 || c > ct) {
    set_error("eval_metrics: need b >= 1, t, h, w >= 2 and 1 <= c <= channels (got b=%d t=%d h=%d w=%d c=%d of %d)", b, t, h,
              w, c, ct);
    return B200FNO_EINVAL;
  }
  *d = MetricsDims{b, t, h, w, ct, c, std::min(t / 2, std::min(h / 2, w / 2))};
  return 0;
}

}  // namespace

size_t metrics_workspace_bytes(int b, int t, int h, int w, int ct, int c) {
  MetricsDims d;
  if (make_dims(b, t, h, w, ct, c, &d)) return 0;
  return metrics_layout(d).total;
}

int launch_eval_metrics(const float* pred, const float* target, int b, int t, int h, int w, int ct, int c, void* ws,
                        size_t ws_bytes, float* out13, cudaStream_t st) {
  MetricsDims d;
  B2_TRY(make_dims(b, t, h, w, ct, c, &d));
  const MetricsLayout L = metrics_layout(d);
  if (!ws || ws_bytes < L.total || ((uintptr_t)ws & 255)) {
    set_error("eval_metrics: workspace too small (%zu < %zu) or not 256-byte aligned", ws_bytes, L.total);
    return B200FNO_EINVAL;
  }
  const size_t smem_w = (size_t)(2 * d.w + 2 * d.w * d.c) * sizeof(float);
  if (smem_w > 200 * 1024) {
    set_error("eval_metrics: row of %d points x %d channels does not fit shared memory", d.w, d.c);
    return B200FNO_EINVAL;
  }
  uint8_t* base = (uint8_t*)ws;
  float2 *twW = (float2*)(base + L.tw), *twH = twW + d.w, *twT = twH + d.h;
  double *acc = (double*)(base + L.acc), *sig = (double*)(base + L.sig), *misc = (double*)(base + L.misc),
         *spec = (double*)(base + L.spec);
  float2 *AW = (float2*)(base + L.AW), *AH = (float2*)(base + L.AH), *F = (float2*)(base + L.F);
  B2_CUDA(cudaMemsetAsync(base + L.acc, 0, L.AW - L.acc, st));  // all accumulators
  const int nmax = std::max(d.w, std::max(d.h, d.t));
  tables_kernel<<<ceil_div(nmax, 256), 256, 0, st>>>(twW, d.w, twH, d.h, twT, d.t);
  B2_LAUNCHED("metrics tables_kernel");
  const int hw = d.h * d.w;
  pointwise_kernel<<<dim3(std::max(1, std::min(ceil_div(hw, 1024), 64)), d.t, d.b), 256, 0, st>>>(pred, target, d, acc, sig);
  B2_LAUNCHED("metrics pointwise_kernel");
  r2_kernel<<<148 * 8, 256, 0, st>>>(target, d, misc + 0);
  B2_LAUNCHED("metrics r2_kernel");
  if (d.c >= 2) {
    ke_kernel<<<148 * 4, 256, 0, st>>>(pred, target, d, misc + 1);
    B2_LAUNCHED("metrics ke_kernel");
  }
  const size_t rows = (size_t)d.b * d.t * d.h;
  B2_CUDA(cudaFuncSetAttribute(dftw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
  dftw_kernel<<<(unsigned)rows, 256, smem_w, st>>>(pred, target, d, twW, AW, rows);
  B2_LAUNCHED("metrics dftw_kernel");
  dft_axis_kernel<<<148 * 16, 256, 0, st>>>(AW, AH, (size_t)d.b * d.t, d.h, (size_t)d.nb * d.c, d.nb, twH);
  B2_LAUNCHED("metrics dft_axis_kernel");
  dft_axis_kernel<<<148 * 16, 256, 0, st>>>(AH, F, (size_t)d.b, d.t, (size_t)d.nb * d.nb * d.c, d.nb, twT);
  B2_LAUNCHED("metrics dft_axis_kernel");
  bin_kernel<<<148 * 8, 256, 0, st>>>(F, d, spec);
  B2_LAUNCHED("metrics bin_kernel");
  // int(np.round(x)): round half to even (metrics.py:84-85)
  const int i_low = (int)nearbyint(d.nb / 3.0), i_high = (int)nearbyint(d.nb * 2 / 3.0);
  finalize_kernel<<<1, 256, 0, st>>>(d, acc, sig, misc, spec, twT, i_low, i_high, out13);
  B2_LAUNCHED("metrics finalize_kernel");
  return 0;
}

}  // namespace b200fno
