// Shared declarations for the b200fno engine (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/b200fno.h"

namespace b200fno {

// ---- error plumbing ----------------------------------------------------------
void set_error(const char* fmt, ...);
int64_t& launch_counter();

#define B2_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      b200fno::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                         \
      return B200FNO_ECUDA;                                                                 \
    }                                                                                       \
  } while (0)

#define B2_LAUNCHED(name)                                                                     \
  do {                                                                                        \
    ++b200fno::launch_counter();                                                              \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      b200fno::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e),    \
                         __FILE__, __LINE__);                                                 \
      return B200FNO_ECUDA;                                                                   \
    }                                                                                         \
  } while (0)

#define B2_TRY(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

// ---- programmatic dependent launch (PDL) ----------------------------------------
// Inside b200fno_forward / b200fno_rollout every kernel after the lift is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may become resident, set up barriers / TMEM /
// tensor-map prefetches while the previous kernel drains, and block in pdl_wait() until that kernel has completed
// and flushed.  Every kernel calls pdl_launch_dependents() first thing so its successor can be scheduled as soon
// as all of its own CTAs have started.  Both instructions are no-ops for a normally launched kernel.
bool& pdl_enabled();  // thread-local switch read by the launchers (api.cu)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int ceil_div(int x, int m) { return (x + m - 1) / m; }

// ---- geometry of the truncated-DFT pipeline ----------------------------------
// Activations are channels-last fp32 [B][Tp][Hp][Wp][Cp] (Tp == 1 for ndim 2).
// Spectral intermediates (ri = 0 real, 1 imaginary; n = contiguous tail):
//   A  = fwdW(act)   [B][Tp][Hp][ri][m3][Cp]
//   Bh = fwdH(A)     [B][Tp][ri][KH][m3][Cp]
//   S  = fwdT(Bh)    [B][ri][KT][KH][m3][Cp]     (ndim 2: S == Bh, KT == 1)
//   O  = modes(S)    [B][ri][KT][KH][m3][Cp]
//   Ct = invT(O)     [B][Tp][ri][KH][m3][Cp]
//   D  = invH(Ct)    [B][Tp][Hp][ri][m3][Cp]
//   act' = act(bn(conv(act) + invW(D)))
struct Geom {
  int ndim;
  int Tp, Hp, Wp;  // padded (FFT) extents
  int Cp;          // channels rounded up to a multiple of 4
  int m3;          // kept W modes
  int KT, KH;      // distinct kept T / H frequencies (<= 2*m1, 2*m2)
  int K2;          // 2*m3 (re/im rows of the W transform)
  int K2p;         // K2 rounded up to 4
  int NM;          // KT*KH*m3 modes
  size_t act_elems(int B) const { return (size_t)B * Tp * Hp * Wp * Cp; }
  size_t a_elems(int B) const { return (size_t)B * Tp * Hp * K2 * Cp; }
  size_t b_elems(int B) const { return (size_t)B * Tp * 2 * KH * m3 * Cp; }
  size_t s_elems(int B) const { return (size_t)B * 2 * NM * Cp; }
};

// One small axis transform on the tensor cores (tc_tmul.cu): Out[g][m][n] = sum_k L[m][k] R[g][k][n]
struct TmulPlan {
  bool ok = false;
  int M = 0, K = 0, N = 0, MT = 0, n_mt = 0, Mpad = 0, Kpad = 0, nchunk = 0, NS = 0, stage_bytes = 0;
  float* table = nullptr;  // device, owned: hi|lo planes [2][Mpad][Kpad]
  CUtensorMap tmL;
};
int tmul_plan_build(TmulPlan* tp, const std::vector<float>& L, int ldl, int M, int K, int N);
void tmul_plan_free(TmulPlan* tp);
// Is the tensor-core kernel the better choice for G groups?  A single-chunk plan (K <= 64) runs one (g, column tile) unit
// per CTA: it needs enough units to fill the GPU (C2 inverse H, 64 units: 1.96 ms on FFMA vs 3.3 ms; C4, 560 units: 5.1 -> 3.0)
bool tmul_use(const TmulPlan& tp, int G);
int tmul_make_data_map(CUtensorMap* m, const float* R, int G, int K, int N, long long strideRg);
int launch_tmul_tc(const TmulPlan& tp, const CUtensorMap& tmR, float* out, int G, long long sOg, long long sOm,
                   int mdiv, long long sOmLo, long long split_off, cudaStream_t st);

// Constant tables (device, fp32) for one geometry; see tables.cu for the values.
struct Tables {
  float* base = nullptr;  // one cudaMalloc
  size_t bytes = 0;
  const float *LF = nullptr, *LH = nullptr, *LT = nullptr, *LTi = nullptr, *LHi = nullptr, *Gt = nullptr;
  int ldLF = 0, ldLH = 0, ldLT = 0, ldLTi = 0, ldLHi = 0;
  std::vector<int> ft, fh;  // actual frequency index of each kept T / H slot
  int *d_ft = nullptr, *d_fh = nullptr;
  float* LF_hl = nullptr;  // forward-W table as hi|lo planes [2][K2][wpad] (tc_fwdw.cu)
  TmulPlan tm_fwdH, tm_fwdT, tm_invT, tm_invH;  // tensor-core versions of the small axis transforms
};
int compute_tables_host(const Geom& g, int m1, int m2, Tables* t, std::vector<float> (&host)[6], int kw0 = 0);
int build_tables(const Geom& g, int m1, int m2, Tables* t, int kw0 = 0);
void free_tables(Tables* t);

// ---- SIMT stage launchers (simt.cu) ------------------------------------------
// Out[g][m][n] = sum_k L[m][k] * R[g][k][n]   (n contiguous; N % 4 == 0)
int launch_lmul(const float* L, int ldl, int M, int K, const float* R, long long strideRg, long long strideRk,
                float* Out, long long strideOg, long long strideOm, int N, int G, cudaStream_t st, int mdiv = 1,
                long long strideOmLo = 0, long long split_off = 0);
// O[b][ri][mode][o] = sum_i S[b][.][mode][i] (x) W[mode][i][.][o]   (complex)
int launch_modes(const float* S, const float* Wpk, float* O, int B, int NM, int Cp, cudaStream_t st);
// out[row][w][o] = f( (sum_i in[row][w][i] convT[i][o] + sum_k Gt[w][k] D[row][k][o]) * scale[o] + shift[o] )
int launch_layer(const float* act_in, float* act_out, const float* convT, const float* Gt, const float* D,
                 const float* scale, const float* shift, long long rows, int Wp, int Cp, int K2, int K2p, int gelu,
                 cudaStream_t st, int bf16 = 0);

struct LiftArgs {
  const float* x;       // [B][T][H][W][c_in]
  float* act;           // [B][Tp][Hp][Wp][Cp]
  const float* W0T;     // [Klp][Cp]  rows: features, grid coords, bias, zero pad
  const int* in_off;    // [Fin]
  const float *gt, *gh, *gw;  // grid coordinate tables (gt == nullptr for ndim 2)
  int B, T, H, W, Tp, Hp, Wp, Cp, c_in, Fin, ng, Klp;
  long long x_sB, x_sT;  // element strides of x for batch and (3-D) frame
  int bf16 = 0;          // bf16 compute mode: features rounded to bf16 (weights rounded when packed)
};
int launch_lift(const LiftArgs& a, cudaStream_t st);
int launch_lift_bwd_input(const LiftArgs& a, const float* dact, float* dx, cudaStream_t st);  // train.cu

struct ProjArgs {
  const float* act;     // [B][Tp][Hp][Wp][Cp]
  const float *fc1T, *fc1b, *fc2T, *fc2b;  // [Cp][128], [128], [128][Fp], [Fp]
  const float *aff_a, *aff_b;              // per physical channel, may be nullptr (identity)
  const int *chan, *out_off, *st_off;      // [Fout]
  float* out;                              // prediction tensor (pre-offset to this step)
  float* state;                            // next model input or nullptr
  int B, T, H, W, Tp, Hp, Wp, Cp, Fout, Fp, c_out, c_in;
  long long out_sB, out_sT, st_sB, st_sT;
  int bf16 = 0;  // bf16 compute mode: x and the hidden activations rounded to bf16
};
int launch_proj(const ProjArgs& a, cudaStream_t st);

// layout helpers for the stand-alone spectral operator
int launch_nchw_to_cl(const float* x, float* act, int B, int C, long long S, int Cp, cudaStream_t st);
int launch_cl_to_nchw(const float* act, float* y, int B, int C, long long S, int Cp, cudaStream_t st);
int launch_copy_params(const float* x0, float* state, long long points, int c_in, int c_out, cudaStream_t st);

// ---- weight packing (pack.cu) --------------------------------------------------
int launch_pack_spectral(const float* const* corners_dev, int ncorner, float* Wpk, const Geom& g, int ci, int co,
                         int m1, int m2, const int* d_ft, const int* d_fh, cudaStream_t st, int m3_src = 0, int kw0 = 0);
// m3_src / kw0: the corner tensors hold m3_src (> g.m3) W modes and this pack takes [kw0, kw0 + g.m3) of them
int launch_transpose_pad(const float* src, int rows, int cols, float* dst, int dst_rows, int dst_cols,
                         cudaStream_t st);  // dst[c][r] = src[r][c], zero elsewhere
int launch_fold_bn(const float* conv_b, const float* bn_w, const float* bn_b, const float* bn_m, const float* bn_v,
                   float eps, int C, int Cp, float* scale, float* shift, cudaStream_t st);
int launch_pad_copy(const float* src, int n, float* dst, int np, cudaStream_t st);
int launch_pack_w0k(const float* fc0_w, const float* fc0_b, int C, int Fin, int ng, int nkl, float* out,
                    cudaStream_t st);
int launch_split_hl(const float* src, int n, float* dst_hi, float* dst_lo, cudaStream_t st, int bf16 = 0);  // 3xTF32 planes
int launch_round_bf16(float* p, size_t n, cudaStream_t st);  // in place: bf16 compute mode weights

// ---- training path (train.cu) ----------------------------------------------------
// train-mode BatchNorm over a flat channels-last tensor [P][Cp]; bnc = [mean|rstd|a|b|s1/n|s2/n] x Cp
int launch_colstats(const float* x, long long P, int Cp, double* stats, cudaStream_t st);
int launch_bn_finalize(const double* stats, long long P, float eps, const float* gamma, const float* beta, int C, int Cp,
                       float* bnc, float* run_mean, float* run_var, float momentum, cudaStream_t st);
int launch_bn_apply(const float* z, float* y, long long P, int Cp, const float* bnc, int gelu, cudaStream_t st);
int launch_bn_backward(const float* gy, const float* z, float* dz, long long P, int C, int Cp, float* bnc, int gelu,
                       double* sums, float* d_gamma, float* d_beta, cudaStream_t st);
// out[m*ldo+n] += sum_p A[p*lda+m] B[p*ldb+n], m < Mv, n < Nv (column Nv -> extra[m]); caller zeroes out
int launch_wgrad(const float* A, int lda, int M, int Mv, const float* B, int ldb, int N, int Nv, long long P, float* out,
                 int ldo, float* extra, cudaStream_t st, int bf16 = 0);
int launch_colsum(const float* A, int lda, int Nv, long long P, float* out, cudaStream_t st);

struct ProjBwdArgs {
  const float* act;                          // x_L [P][Cp] (padded grid, flat)
  const float* dy;                           // [B][t_out][H][W][c_out]
  const float *fc1T, *fc1b, *fc2W, *fc1W;    // [Cp][128], [128], [Fp][128], [128][Cp]
  const int* out_off;                        // [Fout]
  float *G, *dH, *dF, *dact;                 // [P][128], [P][128], [P][Fp], [P][Cp]
  int B, T, H, W, Tp, Hp, Wp, Cp, Fout, Fp, c_out;
  long long out_sB, out_sT;
  int bf16 = 0;  // bf16 compute mode: every GEMM operand rounded to bf16 (train.cu)
};
int launch_proj_bwd(const ProjBwdArgs& a, cudaStream_t st);
int launch_lift_features(const LiftArgs& a, float* feat, cudaStream_t st);
int launch_pack_spectral_adj(const float* Wpk, float* Wadj, int NM, int Cp, cudaStream_t st);
int launch_modes_wgrad(const float* S, const float* dO, float* dW, int B, int NM, int Cp, cudaStream_t st);
int launch_unpack_spectral_grad(const float* dWpk, float* const* corners, int ncorner, const Geom& g, int ci, int co,
                                int m1, int m2, const int* slot_t, const int* slot_h, cudaStream_t st);
int launch_pad2d(const float* src, int rows, int cols, float* dst, int dst_rows, int dst_cols, cudaStream_t st);
// ---- evaluation metrics (metrics.cu) ----------------------------------------------
size_t metrics_workspace_bytes(int b, int t, int h, int w, int ct, int c);
int launch_eval_metrics(const float* pred, const float* target, int b, int t, int h, int w, int ct, int c, void* ws,
                        size_t ws_bytes, float* out13, cudaStream_t st);

int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                long long step, cudaStream_t st);

}  // namespace b200fno
