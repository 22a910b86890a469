// Plan object and the extern "C" entry points declared in include/b200fno.h.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <tuple>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200fno {

// tc_layer.cu
bool tc_layer_supported(const Geom& g);
int tc_make_act_map(CUtensorMap* m, const float* act, long long rows, const Geom& g);
int tc_make_w_map(CUtensorMap* m, const float* w_hl);
int tc_make_d_map(CUtensorMap* m, const float* d, long long rows, const Geom& g);
bool tc_fwdw_supported(const Geom& g);
// tc_modes.cu
bool tc_modes_supported(const Geom& g, int B);
int tc_make_modes_map(CUtensorMap* m, const float* Wpk, int NM);
int launch_modes_tc(const CUtensorMap& tmW, const float* S, float* O, int B, int NM, cudaStream_t st);
int tc_make_fwdw_maps(CUtensorMap* tmX, CUtensorMap* tmF, const float* act, const float* table, long long rows,
                      const Geom& g);
int launch_fwdw_tc(const CUtensorMap& tmX, const CUtensorMap& tmF, float* out, long long rows, const Geom& g,
                   cudaStream_t st, long long row0 = 0);
bool tc_proj_supported(const Geom& g, int Fout);
int tc_proj_n2(int Fout);
int tc_make_proj_act_map(CUtensorMap* m, const float* act, long long rows, const Geom& g, int W);
int tc_make_fc1_map(CUtensorMap* m, const float* w);
int tc_make_fc2_map(CUtensorMap* m, const float* w, int N2);
int launch_proj_tc(const ProjArgs& pa, const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2,
                   cudaStream_t st, int b0 = 0, int nb = -1, int bf16 = 0);
int tc_lift_nkl(int Fin);
int launch_lift_tc(const LiftArgs& la, const CUtensorMap& tmOut, const CUtensorMap& tmW0, const Geom& g,
                   cudaStream_t st, int b0 = 0, int nb = -1, int bf16 = 0);
int launch_layer_tc(const CUtensorMap& tmX, const CUtensorMap& tmOut, const CUtensorMap& tmW, const CUtensorMap& tmD,
                    const float* Gt, const float* scale, const float* shift, long long rows, const Geom& g, int gelu,
                    cudaStream_t st, long long row0 = 0, int bf16 = 0);

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int64_t& launch_counter() { return g_launches; }
static thread_local bool g_pdl = false;
bool& pdl_enabled() { return g_pdl; }

}  // namespace b200fno

using namespace b200fno;

// Diagnostics (B200FNO_DEBUG_FINITE=1): b200fno_train_backward follows every kernel with a scan of that kernel's output
// for non-finite values and keeps the smallest stage id that had one (device side, no host synchronisation, so the
// timing relative to other streams is barely changed).  Stage id = 100 * (layer + 1) + kernel index in the layer loop;
// 1..9 = projection backward, 9000+ = lift backward.  Read with b200fno_debug_first_nonfinite().
__global__ void finite_check_kernel(const float* __restrict__ x, size_t n, int stage, int* __restrict__ first) {
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    bad |= !(fabsf(v) <= 3.0e38f);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicMin(first, stage);
}

// Activation megabytes per chunk of the chunked launch order (run_network); 0 = whole batch per launch.
#ifndef B200FNO_L2_CHUNK_MB_DEFAULT
#define B200FNO_L2_CHUNK_MB_DEFAULT 0.0
#endif

// Optional per-stage CUDA-event timing (bench.py's roofline numbers are measured with this,
// on the caller's stream, around the real launches).
enum Stage { ST_LIFT = 0, ST_FWD_W, ST_FWD_H, ST_FWD_T, ST_MODES, ST_INV_T, ST_INV_H, ST_LAYER, ST_PROJ, ST_COUNT };
struct Timing {
  bool enabled = false;
  std::vector<cudaEvent_t> ev;  // pairs
  std::vector<int> stage;
  size_t used = 0;
  ~Timing() {
    for (auto e : ev) cudaEventDestroy(e);
  }
};
struct StageScope {
  Timing* t;
  cudaStream_t st;
  size_t slot = (size_t)-1;
  // count = false: a further launch of the same logical stage (batch chunks): its time is added, not its launch
  StageScope(Timing* t_, int stage, cudaStream_t s, bool count = true) : t(t_), st(s) {
    if (!t || !t->enabled) return;
    if (!count) stage |= 0x100;
    if (t->used + 2 > t->ev.size()) {
      for (int i = 0; i < 2; ++i) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        t->ev.push_back(e);
      }
      t->stage.push_back(stage);
    } else {
      t->stage[t->used / 2] = stage;
    }
    slot = t->used;
    t->used += 2;
    cudaEventRecord(t->ev[slot], st);
  }
  ~StageScope() {
    if (slot != (size_t)-1) cudaEventRecord(t->ev[slot + 1], st);
  }
};

struct LayerPacked {
  float *convT, *scale, *shift, *spec;
  float *cbias, *gamma, *beta, *convW;  // training: conv bias, BN affine (unfolded), conv weight [o][i]
  float* convHL;  // [2][Cp][Cp]: conv weight [o][i] as 3xTF32 hi | lo planes (tensor-core path)
  CUtensorMap tmW;
  CUtensorMap tmModes;  // packed spectral weights, one mode per box (tc_modes.cu)
};

// modes3 in (32, 64] at width 64: the tensor-core kernels hold at most 32 W modes (TMEM: table rows of the inverse-W
// term; forward-W accumulator width), so the kept W frequencies are handled as TWO slices [0, m3/2) and [m3/2, m3).
// Each slice is a complete truncated-DFT pipeline of its own (tables with a frequency offset, its own packed weights);
// the inverse-W terms add up because the layer kernel runs once per slice:
//   t   = conv(x)  + invW_0(D_0)                       (no BatchNorm / GELU, written to the other activation buffer)
//   out = f(I . t  + invW_1(D_1))                      (identity "conv" - exact in 3xTF32 -, then BatchNorm + GELU)
// i.e. one extra activation pass per layer instead of the FFMA kernels (C5 k = 48 / 64: 6.9 / 9.5 ms -> see profiles).
struct ModeSlice {
  int kw0 = 0;
  Tables tab;
  std::vector<float*> spec;          // per layer: packed weights of this slice [NMs][Cp][2][Cp] (device, owned)
  std::vector<CUtensorMap> tmModes;  // per layer
  CUtensorMap tmFwF;                 // forward-W table of this slice
};

struct b200fno_plan {
  b200fno_desc_t d;
  Geom g;
  Tables tab;
  Geom gs;                        // geometry of one mode slice (== g when the plan is not split)
  std::vector<ModeSlice> slices;  // empty unless split
  float* ident = nullptr;         // split: [identity hi | zero lo] conv planes, unit scale, zero shift (device, owned)
  CUtensorMap tmWI;
  int device = 0;
  int impl_request = B200FNO_IMPL_AUTO;
  // derived feature bookkeeping
  int Fin, ng, Klp, Fout, Fp, Tv;  // Tv: valid t rows of the activation grid (t_in for 3-D, 1 for 2-D)
  int ncorner;
  // device tables owned by the plan
  float* d_grid = nullptr;  // gt | gh | gw
  int* d_int = nullptr;     // in_off | chan | out_off | st_off
  const float *gt = nullptr, *gh = nullptr, *gw = nullptr;
  const int *in_off = nullptr, *chan = nullptr, *out_off = nullptr, *st_off = nullptr;
  // caller-owned
  float* ws = nullptr;
  size_t ws_bytes = 0;
  float* packed = nullptr;
  size_t packed_bytes = 0;
  bool weights_ready = false;
  Timing timing;
  // views into ws / packed
  float *act[2], *bufA, *bufAD, *bufBC, *bufS, *bufO;  // bufA: forward-W output; bufAD: inverse-H output D
  bool split = false;         // modes3 in (32, 64]: two mode slices on the tensor-core kernels (struct ModeSlice)
  bool use_tc_modes = false;  // per-mode mixing on the tensor cores (width 64; per call: batch <= 32)
  int bf16 = 0;  // compute mode (b200fno_plan_set_compute): 1 = torch.autocast(bfloat16) semantics for Linear / Conv
  int* dbg_first = nullptr;  // B200FNO_DEBUG_FINITE: smallest stage id of b200fno_train_backward with a non-finite output
  int chunk_b = 0;  // samples per launch of the activation-sized kernels (0: whole batch), see run_network
  float *W0T, *fc1T, *fc1b, *fc2T, *fc2b;
  std::vector<LayerPacked> layers;
  // tensor-core path
  bool use_pdl = true;  // B200FNO_NO_PDL=1 in the environment disables it (A/B measurements)
  bool use_tc = false, use_tc_lift = false;
  CUtensorMap tmAct[2], tmD, tmW0;
  float* W0K = nullptr;  // [2][64][64] lift weights as K-major hi|lo planes
  bool use_tc_tmul = false;
  CUtensorMap tmR_fwdH, tmR_fwdT, tmR_invT, tmR_invH;
  bool use_tc_fwdw = false;
  CUtensorMap tmFwX[2], tmFwF;
  bool use_tc_proj = false;
  CUtensorMap tmActProj[2], tmFc1, tmFc2;
  float *fc1HL = nullptr, *fc2HL = nullptr;  // [2][128][64], [2][N2][128]
  // training path (b200fno_train_*): untransposed projection weights + the caller-bound training workspace
  float *fc1W = nullptr, *fc2W = nullptr;  // [128][Cp], [Fp][128]
  struct Train {
    bool bound = false, fwd_done = false;
    int last_batch = 0;
    std::vector<float*> xs, zs, Ssave, bnc;  // saved layer inputs x_0..x_L, pre-BN sums z_l, spectra S_l, BN coefficients
    float *g0 = nullptr, *g1 = nullptr;      // gradient ping-pong [P][Cp]
    float *Wadj = nullptr, *dWpk = nullptr;  // adjoint-packed spectral weights, packed spectral gradient
    float *G = nullptr, *dH = nullptr, *dF = nullptr;  // projection scratch [P][128], [P][128], [P][Fp] (G also = lift features)
    double* stats = nullptr;                 // [2][Cp]
    // transposed tables (device, owned by the plan)
    float* tbase = nullptr;
    const float *GtT = nullptr, *LHiT = nullptr, *LTiT = nullptr, *LTT = nullptr, *LHT = nullptr, *LFT = nullptr;
    int *slot_t = nullptr, *slot_h = nullptr;
  } tr;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int make_geom(int ndim, int T, int H, int W, int C, int m1, int m2, int m3, Geom* g) {
  g->ndim = ndim;
  g->Tp = ndim == 3 ? T : 1;
  g->Hp = H;
  g->Wp = W;
  g->Cp = round_up(C, 4);
  g->m3 = m3;
  if (m2 < 1 || m3 < 1 || m2 > H || m3 > W / 2 + 1 || (ndim == 3 && (m1 < 1 || m1 > T))) {
    set_error("modes (%d,%d,%d) do not fit the transformed grid (%d,%d,%d): need m1<=T', m2<=H', m3<=W'/2+1", m1, m2,
              m3, g->Tp, H, W);
    return B200FNO_EINVAL;
  }
  g->KH = std::min(2 * m2, H);
  g->KT = ndim == 3 ? std::min(2 * m1, T) : 1;
  g->K2 = 2 * m3;
  g->K2p = round_up(g->K2, 4);
  g->NM = g->KT * g->KH * m3;
  return 0;
}

// bufAD holds A = fwdW(act) and later D = invH(.); on the tensor-core path D is stored as hi|lo planes
static size_t ad_elems(const Geom& g, int B) {
  return std::max(g.a_elems(B), (size_t)B * g.Tp * g.Hp * 2 * g.K2p * g.Cp);
}
static size_t spectral_scratch_floats(const Geom& g, int B) {
  return align_up(ad_elems(g, B), 64) + align_up(g.b_elems(B), 64) + 2 * align_up(g.s_elems(B), 64);
}
// the plan keeps A in its own buffer: with the chunked launch order (run_network) the forward-W stage of layer l+1
// writes A while later chunks of layer l still read D
static size_t plan_scratch_floats(const Geom& g, int B) { return spectral_scratch_floats(g, B) + align_up(g.a_elems(B), 64); }

// The truncated-DFT spectral operator on a channels-last activation:
//   act -> D   (everything of SpectralConv3d.forward except the last inverse-W stage,
//               which the layer kernel fuses with the bypass conv)
static int run_spectral(const Geom& g, const Tables& tab, int B, const float* act, const float* Wpk, float* bufAD,
                        float* bufBC, float* bufS, float* bufO, cudaStream_t st, Timing* tm = nullptr,
                        bool tc_planes = false, const CUtensorMap* tmFwX = nullptr,
                        const CUtensorMap* tmFwF = nullptr, const CUtensorMap* tmR4 = nullptr, float* bufA = nullptr,
                        bool fwdw_done = false, const CUtensorMap* tmModes = nullptr) {
  // tmR4: data maps {fwdH, fwdT, invT, invH} when the H/T axis transforms run on the tensor cores
  // bufA: where A = fwdW(act) lives (default: it shares bufAD with D, which is written after A has been consumed);
  // fwdw_done: the caller has already launched the forward-W stage into bufA (run_network's chunked order)
  if (!bufA) bufA = bufAD;
  const long long n_hw = (long long)g.m3 * g.Cp;  // contiguous tail after (h,ri) / (ri,kh)
  if (!fwdw_done) {  // forward W: rows (b,t,h): [K2 x Wp] . [Wp x Cp]
    StageScope sc(tm, ST_FWD_W, st);
    if (tmFwX)
      B2_TRY(launch_fwdw_tc(*tmFwX, *tmFwF, bufA, (long long)B * g.Tp * g.Hp, g, st));
    else
      B2_TRY(launch_lmul(tab.LF, tab.ldLF, g.K2, g.Wp, act, (long long)g.Wp * g.Cp, g.Cp, bufA,
                         (long long)g.K2 * g.Cp, g.Cp, g.Cp, B * g.Tp * g.Hp, st));
  }
  {  // forward H: g=(b,t): [2KH x 2Hp] . [2Hp x m3*Cp]
    StageScope sc(tm, ST_FWD_H, st);
    float* fwdH_out = g.ndim == 3 ? bufBC : bufS;
    if (tmR4 && tmul_use(tab.tm_fwdH, B * g.Tp))
      B2_TRY(launch_tmul_tc(tab.tm_fwdH, tmR4[0], fwdH_out, B * g.Tp, 2LL * g.KH * n_hw, n_hw, 1, 0, 0, st));
    else
      B2_TRY(launch_lmul(tab.LH, tab.ldLH, 2 * g.KH, 2 * g.Hp, bufA, (long long)g.Hp * 2 * n_hw, n_hw, fwdH_out,
                         2LL * g.KH * n_hw, n_hw, (int)n_hw, B * g.Tp, st));
  }
  if (g.ndim == 3) {
    StageScope sc(tm, ST_FWD_T, st);
    const long long n_t = (long long)g.KH * n_hw;
    if (tmR4 && tmul_use(tab.tm_fwdT, B))
      B2_TRY(launch_tmul_tc(tab.tm_fwdT, tmR4[1], bufS, B, 2LL * g.KT * n_t, n_t, 1, 0, 0, st));
    else
      B2_TRY(launch_lmul(tab.LT, tab.ldLT, 2 * g.KT, 2 * g.Tp, bufBC, (long long)g.Tp * 2 * n_t, n_t, bufS,
                         2LL * g.KT * n_t, n_t, (int)n_t, B, st));
  }
  {
    StageScope sc(tm, ST_MODES, st);
    if (tmModes && tc_modes_supported(g, B))  // width 64, batch <= 32: the mixing GEMM on the tensor cores
      B2_TRY(launch_modes_tc(*tmModes, bufS, bufO, B, g.NM, st));
    else
      B2_TRY(launch_modes(bufS, Wpk, bufO, B, g.NM, g.Cp, st));
  }
  const float* invH_in = bufO;
  if (g.ndim == 3) {
    StageScope sc(tm, ST_INV_T, st);
    const long long n_t = (long long)g.KH * n_hw;
    if (tmR4 && tmul_use(tab.tm_invT, B))
      B2_TRY(launch_tmul_tc(tab.tm_invT, tmR4[2], bufBC, B, (long long)g.Tp * 2 * n_t, n_t, 1, 0, 0, st));
    else
      B2_TRY(launch_lmul(tab.LTi, tab.ldLTi, 2 * g.Tp, 2 * g.KT, bufO, 2LL * g.KT * n_t, n_t, bufBC,
                         (long long)g.Tp * 2 * n_t, n_t, (int)n_t, B, st));
    invH_in = bufBC;
  }
  {
    StageScope sc(tm, ST_INV_H, st);
    if (tc_planes) {
      // D[row=(g,h)][hl][k=(ri,kw)][o]: m = h*2+ri -> h * (2*K2p*Cp) + ri * (m3*Cp); lo plane at +K2p*Cp
      const long long plane = (long long)g.K2p * g.Cp;
      if (tmR4 && tmul_use(tab.tm_invH, B * g.Tp))
        B2_TRY(launch_tmul_tc(tab.tm_invH, tmR4[3], bufAD, B * g.Tp, (long long)g.Hp * 2 * plane, 2 * plane, 2, n_hw,
                              plane, st));
      else
        B2_TRY(launch_lmul(tab.LHi, tab.ldLHi, 2 * g.Hp, 2 * g.KH, invH_in, 2LL * g.KH * n_hw, n_hw, bufAD,
                           (long long)g.Hp * 2 * plane, 2 * plane, (int)n_hw, B * g.Tp, st, 2, n_hw, plane));
    } else {
      B2_TRY(launch_lmul(tab.LHi, tab.ldLHi, 2 * g.Hp, 2 * g.KH, invH_in, 2LL * g.KH * n_hw, n_hw, bufAD,
                         (long long)g.Hp * 2 * n_hw, n_hw, (int)n_hw, B * g.Tp, st));
    }
  }
  return 0;
}

static int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return B200FNO_ENODEV;
  }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess || p.major != 10) {
    set_error("b200fno is built for sm_100a only; device %d is sm_%d%d", dev, p.major, p.minor);
    return B200FNO_ENODEV;
  }
  return 0;
}

extern "C" {

const char* b200fno_last_error(void) { return g_err; }
int b200fno_abi_version(void) { return B200FNO_ABI_VERSION; }
int64_t b200fno_launch_count(void) { return g_launches; }
void b200fno_launch_count_reset(void) { g_launches = 0; }

int b200fno_plan_create(const b200fno_desc_t* d, b200fno_plan_t** out) {
  if (!d || !out) {
    set_error("null argument");
    return B200FNO_EINVAL;
  }
  *out = nullptr;
  if (d->abi_version != B200FNO_ABI_VERSION) {
    set_error("descriptor abi_version %d != library %d", d->abi_version, B200FNO_ABI_VERSION);
    return B200FNO_EINVAL;
  }
  if ((d->ndim != 2 && d->ndim != 3) || d->max_batch < 1 || d->t_in < 1 || d->t_out < 1 || d->h < 1 || d->w < 1 ||
      d->c_in < 1 || d->c_out < 1 || d->width < 1 || d->n_layers < 1 || d->padding < 0) {
    set_error("bad descriptor field (ndim must be 2|3, sizes >= 1)");
    return B200FNO_EINVAL;
  }
  if (d->proj_hidden != 128) {
    set_error("proj_hidden must be 128 (fno.py:102), got %d", d->proj_hidden);
    return B200FNO_EINVAL;
  }
  if (d->ndim == 3 && d->t_out % d->t_in != 0) {
    set_error("t_out (%d) must be a multiple of t_in (%d) (fno.py:86)", d->t_out, d->t_in);
    return B200FNO_EINVAL;
  }
  B2_TRY(check_device());
  b200fno_plan* p = new (std::nothrow) b200fno_plan();
  if (!p) {
    set_error("out of host memory");
    return B200FNO_EINVAL;
  }
  p->d = *d;
  p->use_pdl = getenv("B200FNO_NO_PDL") == nullptr;
  cudaGetDevice(&p->device);
  const int pad = d->padding;
  int rc = make_geom(d->ndim, d->t_in + pad, d->h + pad, d->w + pad, d->width, d->modes1, d->modes2, d->modes3, &p->g);
  if (rc) {
    delete p;
    return rc;
  }
  rc = build_tables(p->g, d->modes1, d->modes2, &p->tab);
  if (rc) {
    delete p;
    return rc;
  }
  const int T = d->t_in, H = d->h, W = d->w, Ci = d->c_in, Co = d->c_out;
  if (d->ndim == 3) {
    p->Fin = Ci, p->ng = 3, p->Fout = Co * (d->t_out / T), p->Tv = T, p->ncorner = 4;
  } else {
    p->Fin = T * Ci, p->ng = 2, p->Fout = d->t_out * Co, p->Tv = 1, p->ncorner = 2;
  }
  p->Klp = round_up(p->Fin + p->ng + 1, 4);
  p->Fp = round_up(p->Fout, 4);
  // grid coordinates: fp32(np.linspace(0,1,n)) (fno.py:137-141)
  std::vector<float> grid(T + H + W);
  auto lin = [](float* o, int n) {
    for (int i = 0; i < n; ++i) o[i] = (n == 1) ? 0.f : (i == n - 1 ? 1.f : (float)((double)i * (1.0 / (double)(n - 1))));
  };
  lin(grid.data(), T), lin(grid.data() + T, H), lin(grid.data() + T + H, W);
  std::vector<int> ints(p->Fin + 3 * p->Fout);
  int* in_off = ints.data();
  int *chan = in_off + p->Fin, *out_off = chan + p->Fout, *st_off = out_off + p->Fout;
  const long long HW = (long long)H * W;
  if (d->ndim == 3) {
    const int r = d->t_out / T;
    for (int j = 0; j < p->Fin; ++j) in_off[j] = j;
    for (int f = 0; f < p->Fout; ++f) {  // fno.py:127-128: feature f = c*r + rho -> frame t*r + rho, channel c
      int c = f / r, rho = f % r;
      chan[f] = c;
      out_off[f] = (int)(rho * HW * Co + c);
      st_off[f] = (int)(rho * HW * Ci + c);
    }
  } else {
    for (int j = 0; j < p->Fin; ++j) in_off[j] = (int)((j / Ci) * HW * Ci + (j % Ci));
    for (int f = 0; f < p->Fout; ++f) {
      int to = f / Co, c = f % Co;
      chan[f] = c;
      out_off[f] = (int)(to * HW * Co + c);
      st_off[f] = (int)(to * HW * Ci + c);
    }
  }
  auto fail = [&](const char* what, cudaError_t e) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    b200fno_plan_destroy(p);
    return B200FNO_ECUDA;
  };
  cudaError_t e;
  if ((e = cudaMalloc((void**)&p->d_grid, grid.size() * sizeof(float))) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMalloc((void**)&p->d_int, ints.size() * sizeof(int))) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMemcpy(p->d_grid, grid.data(), grid.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess)
    return fail("cudaMemcpy", e);
  if ((e = cudaMemcpy(p->d_int, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess)
    return fail("cudaMemcpy", e);
  p->gt = d->ndim == 3 ? p->d_grid : nullptr;
  p->gh = p->d_grid + T;
  p->gw = p->d_grid + T + H;
  p->in_off = p->d_int;
  p->chan = p->d_int + p->Fin;
  p->out_off = p->chan + p->Fout;
  p->st_off = p->out_off + p->Fout;
  *out = p;
  return 0;
}

int b200fno_plan_destroy(b200fno_plan_t* p) {
  if (!p) return 0;
  free_tables(&p->tab);
  for (auto& sl : p->slices) {
    free_tables(&sl.tab);
    for (float* q : sl.spec)
      if (q) cudaFree(q);
  }
  if (p->ident) cudaFree(p->ident);
  if (p->tr.tbase) cudaFree(p->tr.tbase);
  if (p->d_grid) cudaFree(p->d_grid);
  if (p->d_int) cudaFree(p->d_int);
  delete p;
  return 0;
}

// width 64, modes3 in (32, 64], two equal slices the forward-W kernel supports (2 * slice in {16, 32, 48, 64})
static bool split_geom(const b200fno_plan* p, Geom* gs) {
  const Geom& g = p->g;
  if (g.Cp != 64 || p->d.width != 64 || g.m3 <= 32 || g.m3 > 64 || g.m3 % 16 != 0) return false;
  Geom s = g;
  s.m3 = g.m3 / 2, s.K2 = 2 * s.m3, s.K2p = s.K2, s.NM = g.KT * g.KH * s.m3;
  if (!tc_layer_supported(s) || !tc_fwdw_supported(s)) return false;
  if (gs) *gs = s;
  return true;
}
static bool plan_can_tc(const b200fno_plan* p) {
  return (tc_layer_supported(p->g) && p->g.K2 == p->g.K2p && p->d.width == p->g.Cp) || split_geom(p, nullptr);
}

int b200fno_plan_set_impl(b200fno_plan_t* p, int impl) {
  if (!p || impl < B200FNO_IMPL_AUTO || impl > B200FNO_IMPL_TC) {
    set_error("bad impl selector");
    return B200FNO_EINVAL;
  }
  const bool can_tc = plan_can_tc(p);
  if (impl == B200FNO_IMPL_TC && !can_tc) {
    set_error("tensor-core layer kernel needs width 64 and modes3 a multiple of 4 up to 32 (or 48 / 64, run as two slices); got width %d, modes3 %d",
              p->d.width, p->d.modes3);
    return B200FNO_EINVAL;
  }
  if (p->ws) {
    set_error("b200fno_plan_set_impl must be called before b200fno_plan_bind");
    return B200FNO_ESTATE;
  }
  p->impl_request = impl;
  return 0;
}
int b200fno_plan_set_compute(b200fno_plan_t* p, int compute) {
  if (!p || (compute != B200FNO_COMPUTE_F32 && compute != B200FNO_COMPUTE_BF16)) {
    set_error("bad compute selector");
    return B200FNO_EINVAL;
  }
  if (p->bf16 != (compute == B200FNO_COMPUTE_BF16)) p->weights_ready = false;  // the packed copies depend on it
  p->bf16 = compute == B200FNO_COMPUTE_BF16;
  return 0;
}
int b200fno_plan_get_compute(const b200fno_plan_t* p) { return p ? (p->bf16 ? B200FNO_COMPUTE_BF16 : B200FNO_COMPUTE_F32) : B200FNO_EINVAL; }
int b200fno_plan_get_impl(const b200fno_plan_t* p) {
  if (!p) return B200FNO_EINVAL;
  return (p->impl_request != B200FNO_IMPL_SIMT && plan_can_tc(p)) ? B200FNO_IMPL_TC : B200FNO_IMPL_SIMT;
}

size_t b200fno_plan_workspace_bytes(const b200fno_plan_t* p) {
  if (!p) return 0;
  const int B = p->d.max_batch;
  return (2 * align_up(p->g.act_elems(B), 64) + plan_scratch_floats(p->g, B)) * sizeof(float);
}

static size_t packed_floats(const b200fno_plan* p) {
  const Geom& g = p->g;
  size_t n = align_up((size_t)p->Klp * g.Cp, 64) + 2 * 4096 + 2 * 128 * 64 + 2 * 64 * 128;
  n += (size_t)p->d.n_layers *
       (3 * align_up((size_t)g.Cp * g.Cp, 64) + 2 * align_up(g.Cp, 64) + align_up((size_t)g.NM * g.Cp * 2 * g.Cp, 64));
  n += align_up((size_t)g.Cp * 128, 64) + 128 + align_up((size_t)128 * p->Fp, 64) + align_up(p->Fp, 64);
  // training copies: per layer conv bias, BN weight, BN bias, conv weight [o][i]; fc1 [128][Cp], fc2 [Fp][128]
  n += (size_t)p->d.n_layers * (3 * align_up(g.Cp, 64) + align_up((size_t)g.Cp * g.Cp, 64));
  n += align_up((size_t)128 * g.Cp, 64) + align_up((size_t)p->Fp * 128, 64);
  return n;
}
size_t b200fno_plan_packed_bytes(const b200fno_plan_t* p) { return p ? packed_floats(p) * sizeof(float) : 0; }

int b200fno_plan_bind(b200fno_plan_t* p, void* workspace, size_t workspace_bytes, void* packed, size_t packed_bytes) {
  if (!p || !workspace || !packed) {
    set_error("null argument");
    return B200FNO_EINVAL;
  }
  if (workspace_bytes < b200fno_plan_workspace_bytes(p) || packed_bytes < b200fno_plan_packed_bytes(p)) {
    set_error("buffers too small: workspace %zu < %zu or packed %zu < %zu", workspace_bytes,
              b200fno_plan_workspace_bytes(p), packed_bytes, b200fno_plan_packed_bytes(p));
    return B200FNO_EINVAL;
  }
  if (((uintptr_t)workspace | (uintptr_t)packed) & 255) {
    set_error("workspace and packed buffers must be 256-byte aligned");
    return B200FNO_EINVAL;
  }
  const Geom& g = p->g;
  const int B = p->d.max_batch;
  float* w = (float*)workspace;
  p->ws = w, p->ws_bytes = workspace_bytes;
  p->act[0] = w, w += align_up(g.act_elems(B), 64);
  p->act[1] = w, w += align_up(g.act_elems(B), 64);
  p->bufAD = w, w += align_up(ad_elems(g, B), 64);
  p->bufBC = w, w += align_up(g.b_elems(B), 64);
  p->bufS = w, w += align_up(g.s_elems(B), 64);
  p->bufO = w, w += align_up(g.s_elems(B), 64);
  p->bufA = w;
  float* q = (float*)packed;
  p->packed = q, p->packed_bytes = packed_bytes;
  p->W0T = q, q += align_up((size_t)p->Klp * g.Cp, 64);
  p->W0K = q, q += 2 * 4096;
  p->fc1HL = q, q += 2 * 128 * 64;
  p->fc2HL = q, q += 2 * 64 * 128;
  p->layers.resize(p->d.n_layers);
  for (auto& L : p->layers) {
    L.convT = q, q += align_up((size_t)g.Cp * g.Cp, 64);
    L.scale = q, q += align_up(g.Cp, 64);
    L.shift = q, q += align_up(g.Cp, 64);
    L.spec = q, q += align_up((size_t)g.NM * g.Cp * 2 * g.Cp, 64);
    L.convHL = q, q += 2 * align_up((size_t)g.Cp * g.Cp, 64);
  }
  p->fc1T = q, q += align_up((size_t)g.Cp * 128, 64);
  p->fc1b = q, q += 128;
  p->fc2T = q, q += align_up((size_t)128 * p->Fp, 64);
  p->fc2b = q, q += align_up(p->Fp, 64);
  for (auto& L : p->layers) {
    L.cbias = q, q += align_up(g.Cp, 64);
    L.gamma = q, q += align_up(g.Cp, 64);
    L.beta = q, q += align_up(g.Cp, 64);
    L.convW = q, q += align_up((size_t)g.Cp * g.Cp, 64);
  }
  p->fc1W = q, q += align_up((size_t)128 * g.Cp, 64);
  p->fc2W = q;
  p->weights_ready = false;
  p->use_tc = p->impl_request != B200FNO_IMPL_SIMT && tc_layer_supported(g) && g.K2 == g.K2p && p->d.width == g.Cp;
  p->gs = g;
  const bool split = !p->use_tc && p->impl_request != B200FNO_IMPL_SIMT && split_geom(p, &p->gs) &&
                     getenv("B200FNO_NO_MODE_SPLIT") == nullptr;
  if (split && p->slices.empty()) {  // two mode slices (struct ModeSlice): tables, packed-weight buffers, identity planes
    p->slices.resize(2);
    const size_t spec_floats = align_up((size_t)p->gs.NM * g.Cp * 2 * g.Cp, 64);
    for (int s = 0; s < 2; ++s) {
      ModeSlice& sl = p->slices[s];
      sl.kw0 = s * p->gs.m3;
      B2_TRY(build_tables(p->gs, p->d.modes1, p->d.modes2, &sl.tab, sl.kw0));
      sl.spec.assign(p->d.n_layers, nullptr);
      sl.tmModes.resize(p->d.n_layers);
      for (int l = 0; l < p->d.n_layers; ++l) {
        B2_CUDA(cudaMalloc((void**)&sl.spec[l], spec_floats * sizeof(float)));
        B2_TRY(tc_make_modes_map(&sl.tmModes[l], sl.spec[l], p->gs.NM));
      }
    }
    std::vector<float> id(2 * 4096 + 128, 0.f);
    for (int i = 0; i < 64; ++i) id[(size_t)i * 64 + i] = 1.f, id[2 * 4096 + i] = 1.f;  // hi = I, lo = 0; scale = 1; shift = 0
    B2_CUDA(cudaMalloc((void**)&p->ident, id.size() * sizeof(float)));
    B2_CUDA(cudaMemcpy(p->ident, id.data(), id.size() * sizeof(float), cudaMemcpyHostToDevice));
    B2_TRY(tc_make_w_map(&p->tmWI, p->ident));
  }
  p->split = split;
  if (split) p->use_tc = true;
  const Geom& gt = p->gs;                                      // geometry the tensor-core maps are built for
  const Tables& tbt = split ? p->slices[0].tab : p->tab;       // (the H / T tables do not depend on the W slice)
  // the lift and the projection do not depend on the mode count: they run on the tensor cores for every width-64
  // model, also where the layer kernel itself has to fall back to the FFMA version (modes3 > 32)
  const bool tc64 = p->impl_request != B200FNO_IMPL_SIMT && g.Cp == 64 && p->d.width == 64;
  const long long rows = (long long)B * g.Tp * g.Hp;
  p->use_tc_lift = p->use_tc_proj = false;
  if (p->use_tc || tc64) {
    B2_TRY(tc_make_act_map(&p->tmAct[0], p->act[0], rows, g));
    B2_TRY(tc_make_act_map(&p->tmAct[1], p->act[1], rows, g));
    p->use_tc_lift = tc_lift_nkl(p->Fin) > 0;
    if (p->use_tc_lift) B2_TRY(tc_make_w_map(&p->tmW0, p->W0K));
    p->use_tc_proj = tc_proj_supported(g, p->Fout);
    if (p->use_tc_proj) {
      B2_TRY(tc_make_proj_act_map(&p->tmActProj[0], p->act[0], rows, g, p->d.w));
      B2_TRY(tc_make_proj_act_map(&p->tmActProj[1], p->act[1], rows, g, p->d.w));
      B2_TRY(tc_make_fc1_map(&p->tmFc1, p->fc1HL));
      B2_TRY(tc_make_fc2_map(&p->tmFc2, p->fc2HL, tc_proj_n2(p->Fout)));
    }
  }
  p->use_tc_modes = tc64 && getenv("B200FNO_NO_TC_MODES") == nullptr;
  if (p->use_tc_modes)
    for (auto& L : p->layers) B2_TRY(tc_make_modes_map(&L.tmModes, L.spec, g.NM));
  if (p->use_tc) {
    B2_TRY(tc_make_d_map(&p->tmD, p->bufAD, rows, gt));
    for (auto& L : p->layers) B2_TRY(tc_make_w_map(&L.tmW, L.convHL));
    {  // H / T axis transforms on the tensor cores
      const Tables& tb = tbt;
      const int n_hw = gt.m3 * g.Cp, n_t = g.KH * n_hw, GT = B * g.Tp;
      p->use_tc_tmul = tb.tm_fwdH.ok || tb.tm_invH.ok || tb.tm_fwdT.ok || tb.tm_invT.ok;
      if (tb.tm_fwdH.ok)
        B2_TRY(tmul_make_data_map(&p->tmR_fwdH, p->bufA, GT, 2 * g.Hp, n_hw, (long long)g.Hp * 2 * n_hw));
      if (tb.tm_invH.ok)
        B2_TRY(tmul_make_data_map(&p->tmR_invH, g.ndim == 3 ? p->bufBC : p->bufO, GT, 2 * g.KH, n_hw,
                                  2LL * g.KH * n_hw));
      if (tb.tm_fwdT.ok)
        B2_TRY(tmul_make_data_map(&p->tmR_fwdT, p->bufBC, B, 2 * g.Tp, n_t, (long long)g.Tp * 2 * n_t));
      if (tb.tm_invT.ok)
        B2_TRY(tmul_make_data_map(&p->tmR_invT, p->bufO, B, 2 * g.KT, n_t, 2LL * g.KT * n_t));
    }
    p->use_tc_fwdw = tc_fwdw_supported(gt);
    {  // samples per launch of the activation-sized kernels (run_network): B200FNO_L2_CHUNK_MB = activation bytes per
       // chunk that should stay L2-resident between the producing kernel and its consumer (0 disables chunking)
      const char* e = getenv("B200FNO_L2_CHUNK_MB");
      const double mb = e ? atof(e) : B200FNO_L2_CHUNK_MB_DEFAULT;
      const double per_sample = (double)g.Tp * g.Hp * g.Wp * g.Cp * 4 / 1e6;
      p->chunk_b = mb > 0 ? std::max(1, (int)(mb / per_sample)) : 0;
    }
    if (p->use_tc_fwdw) {
      B2_TRY(tc_make_fwdw_maps(&p->tmFwX[0], &p->tmFwF, p->act[0], tbt.LF_hl, rows, gt));
      B2_TRY(tc_make_fwdw_maps(&p->tmFwX[1], &p->tmFwF, p->act[1], tbt.LF_hl, rows, gt));
      CUtensorMap tmp;
      for (auto& sl : p->slices) B2_TRY(tc_make_fwdw_maps(&tmp, &sl.tmFwF, p->act[0], sl.tab.LF_hl, rows, gt));
    }
  }
  return 0;
}

int b200fno_pack_weights(b200fno_plan_t* p, const b200fno_weights_t* w, void* stream) {
  if (!p || !w) {
    set_error("null argument");
    return B200FNO_EINVAL;
  }
  if (!p->packed) {
    set_error("b200fno_plan_bind must be called before b200fno_pack_weights");
    return B200FNO_ESTATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const Geom& g = p->g;
  const int C = p->d.width, nf = p->Fin + p->ng;
  B2_TRY(launch_transpose_pad(w->fc0_w, C, nf, p->W0T, p->Klp, g.Cp, st));
  B2_TRY(launch_pad_copy(w->fc0_b, C, p->W0T + (size_t)nf * g.Cp, g.Cp, st));
  if (p->use_tc_lift) {  // K-major lift weights, then split in place into hi | lo
    B2_TRY(launch_pack_w0k(w->fc0_w, w->fc0_b, C, p->Fin, p->ng, tc_lift_nkl(p->Fin), p->W0K, st));
    B2_TRY(launch_split_hl(p->W0K, 4096, p->W0K, p->W0K + 4096, st, p->bf16));
  }
  for (int l = 0; l < p->d.n_layers; ++l) {
    LayerPacked& L = p->layers[l];
    B2_TRY(launch_transpose_pad(w->conv_w[l], C, C, L.convT, g.Cp, g.Cp, st));
    if (g.Cp == C)  // tensor-core operand planes (K-major = the reference [o][i] layout)
      B2_TRY(launch_split_hl(w->conv_w[l], C * C, L.convHL, L.convHL + align_up((size_t)g.Cp * g.Cp, 64), st, p->bf16));
    if (p->bf16) B2_TRY(launch_round_bf16(L.convT, (size_t)g.Cp * g.Cp, st));
    B2_TRY(launch_fold_bn(w->conv_b[l], w->bn_weight[l], w->bn_bias[l], w->bn_mean[l], w->bn_var[l], p->d.bn_eps, C,
                          g.Cp, L.scale, L.shift, st));
    B2_TRY(launch_pack_spectral(w->spec_w + (size_t)l * p->ncorner, p->ncorner, L.spec, g, C, C, p->d.modes1,
                                p->d.modes2, p->tab.d_ft, p->tab.d_fh, st));
    for (auto& sl : p->slices)  // the same weights, W modes [kw0, kw0 + m3 / 2) only (the full pack serves training)
      B2_TRY(launch_pack_spectral(w->spec_w + (size_t)l * p->ncorner, p->ncorner, sl.spec[l], p->gs, C, C, p->d.modes1,
                                  p->d.modes2, sl.tab.d_ft, sl.tab.d_fh, st, g.m3, sl.kw0));
    B2_TRY(launch_pad_copy(w->conv_b[l], C, L.cbias, g.Cp, st));
    B2_TRY(launch_pad_copy(w->bn_weight[l], C, L.gamma, g.Cp, st));
    B2_TRY(launch_pad_copy(w->bn_bias[l], C, L.beta, g.Cp, st));
    B2_TRY(launch_pad2d(w->conv_w[l], C, C, L.convW, g.Cp, g.Cp, st));
  }
  B2_TRY(launch_pad2d(w->fc1_w, 128, C, p->fc1W, 128, g.Cp, st));
  B2_TRY(launch_pad2d(w->fc2_w, p->Fout, 128, p->fc2W, p->Fp, 128, st));
  B2_TRY(launch_transpose_pad(w->fc1_w, 128, C, p->fc1T, g.Cp, 128, st));
  B2_TRY(launch_pad_copy(w->fc1_b, 128, p->fc1b, 128, st));
  B2_TRY(launch_transpose_pad(w->fc2_w, p->Fout, 128, p->fc2T, 128, p->Fp, st));
  B2_TRY(launch_pad_copy(w->fc2_b, p->Fout, p->fc2b, p->Fp, st));
  if (p->use_tc_proj) {  // K-major (= reference [out][in]) hi | lo planes for the tensor-core projection
    const int N2 = tc_proj_n2(p->Fout);
    B2_TRY(launch_split_hl(w->fc1_w, 128 * 64, p->fc1HL, p->fc1HL + 128 * 64, st, p->bf16));
    B2_TRY(launch_pad_copy(w->fc2_w, p->Fout * 128, p->fc2HL, N2 * 128, st));
    B2_TRY(launch_split_hl(p->fc2HL, N2 * 128, p->fc2HL, p->fc2HL + N2 * 128, st, p->bf16));
  }
  if (p->bf16) {  // FFMA-path copies of the Linear weights (bias row of W0T included: autocast casts the bias too)
    B2_TRY(launch_round_bf16(p->W0T, (size_t)p->Klp * g.Cp, st));
    B2_TRY(launch_round_bf16(p->fc1T, (size_t)g.Cp * 128, st));
    B2_TRY(launch_round_bf16(p->fc2T, (size_t)128 * p->Fp, st));
    // training copies (backward GEMMs: grad_input = grad_output . W with W as a bf16 tensor)
    for (auto& L : p->layers) B2_TRY(launch_round_bf16(L.convW, (size_t)g.Cp * g.Cp, st));
    B2_TRY(launch_round_bf16(p->fc1W, (size_t)128 * g.Cp, st));
    B2_TRY(launch_round_bf16(p->fc2W, (size_t)p->Fp * 128, st));
  }
  p->weights_ready = true;
  return 0;
}

static LiftArgs make_lift_args(const b200fno_plan* p, int B, const float* x, float* act) {
  const Geom& g = p->g;
  const b200fno_desc_t& d = p->d;
  LiftArgs la{};
  la.x = x, la.act = act, la.W0T = p->W0T, la.in_off = p->in_off;
  la.gt = p->gt, la.gh = p->gh, la.gw = p->gw;
  la.B = B, la.T = p->Tv, la.H = d.h, la.W = d.w, la.Tp = g.Tp, la.Hp = g.Hp, la.Wp = g.Wp, la.Cp = g.Cp;
  la.c_in = d.c_in, la.Fin = p->Fin, la.ng = p->ng, la.Klp = p->Klp;
  la.x_sB = (long long)d.t_in * d.h * d.w * d.c_in;
  la.x_sT = d.ndim == 3 ? (long long)d.h * d.w * d.c_in : 0;
  la.bf16 = p->bf16;
  return la;
}

struct ProjSpec {  // what the projection kernel writes (one rollout step or a plain forward)
  const float *aff_a = nullptr, *aff_b = nullptr;
  float* out = nullptr;
  long long out_sB = 0;
  float* state = nullptr;
};

static ProjArgs make_proj_args(const b200fno_plan* p, int B, const float* act, const ProjSpec& ps) {
  const Geom& g = p->g;
  const b200fno_desc_t& d = p->d;
  ProjArgs pa{};
  pa.act = act, pa.fc1T = p->fc1T, pa.fc1b = p->fc1b, pa.fc2T = p->fc2T, pa.fc2b = p->fc2b;
  pa.aff_a = ps.aff_a, pa.aff_b = ps.aff_b, pa.chan = p->chan, pa.out_off = p->out_off, pa.st_off = p->st_off;
  pa.out = ps.out, pa.state = ps.state;
  pa.B = B, pa.T = p->Tv, pa.H = d.h, pa.W = d.w, pa.Tp = g.Tp, pa.Hp = g.Hp, pa.Wp = g.Wp, pa.Cp = g.Cp;
  pa.Fout = p->Fout, pa.Fp = p->Fp, pa.c_out = d.c_out, pa.c_in = d.c_in;
  const long long HW = (long long)d.h * d.w;
  pa.out_sB = ps.out_sB;
  pa.out_sT = d.ndim == 3 ? (long long)(d.t_out / d.t_in) * HW * d.c_out : 0;
  pa.st_sB = (long long)d.t_in * HW * d.c_in;
  pa.st_sT = d.ndim == 3 ? HW * d.c_in : 0;
  pa.bf16 = p->bf16;
  return pa;
}

struct PdlScope {
  bool prev;
  explicit PdlScope(bool on) : prev(pdl_enabled()) { pdl_enabled() = on; }
  ~PdlScope() { pdl_enabled() = prev; }
};

// Projection alone (the training forward uses it after its own trunk): crop -> fc1 -> GELU -> fc2 -> unfold [-> affine]
static int run_proj(b200fno_plan* p, int B, const float* act, const ProjSpec& ps, cudaStream_t st, bool allow_tc = true) {
  const ProjArgs pa = make_proj_args(p, B, act, ps);
  StageScope sc(&p->timing, ST_PROJ, st);
  PdlScope pdl(p->use_pdl && !p->timing.enabled && allow_tc);  // follows the last layer kernel
  if (p->use_tc_proj && allow_tc) return launch_proj_tc(pa, p->tmActProj[act == p->act[0] ? 0 : 1], p->tmFc1, p->tmFc2, st, 0, -1, p->bf16);
  return launch_proj(pa, st);
}

// FNO3d.forward (fno.py:105-129) = lift, L Fourier layers, projection.
//
// Launch order.  Every Fourier layer needs two passes over its input activation - the forward-W transform and the
// fused layer kernel - and the FFT is global over a sample, so the second pass cannot be fused into the first.  At
// the headline size an activation (278 MB for 8 samples) is larger than the 126 MB L2, so launched batch-wide the
// forward-W kernel re-reads from HBM what the previous layer kernel has just written.  On the tensor-core path the
// activation-sized kernels therefore run per CHUNK of `chunk_b` samples (B200FNO_L2_CHUNK_MB of activation), each
// producer immediately followed by its consumer on the same chunk:
//     for c: lift(c), fwdW_0(c)
//     for l: fwdH, modes, invH on the whole batch;  for c: layer_l(c), then fwdW_{l+1}(c)  (or proj(c) after the last)
// so the forward-W kernel (and the projection) find their input in L2 and HBM sees the SURVEY 8d traffic: one read
// and one write of the activation per layer.  The small L2-resident stages stay batch-wide (they are latency bound).
// A and D live in separate buffers: fwdW_{l+1}(c) writes A while layer_l(c+1) still reads D.
static int run_network(b200fno_plan* p, int B, const float* x, const ProjSpec& ps, cudaStream_t st,
                       long long x_sB = 0) {  // x_sB: element stride between the samples of x (0: contiguous samples)
  const Geom& g = p->g;
  const b200fno_desc_t& d = p->d;
  const int L = d.n_layers;
  // the whole network is chained with programmatic dependent launches (common.cuh); per-stage event timing sits
  // between the launches and would break the chain, so it keeps plain launches.  The lift may overlap the tail of
  // whatever kernel precedes it (projection of the previous rollout step, a weight-pack or a torch kernel): before
  // its pdl_wait() it only touches plan constants.
  PdlScope pdl_scope(p->use_pdl && !p->timing.enabled);
  Timing* tm = &p->timing;
  const long long rows_s = (long long)g.Tp * g.Hp;  // activation rows per sample
  const CUtensorMap tmR4[4] = {p->tmR_fwdH, p->tmR_fwdT, p->tmR_invT, p->tmR_invH};
  LiftArgs la = make_lift_args(p, B, x, p->act[0]);
  if (x_sB) la.x_sB = x_sB;
  if (p->split) {  // modes3 in (32, 64]: two mode slices per layer (struct ModeSlice); the layer output returns to
                   // the buffer its input came from, so `cur` never flips
    const Geom& gs = p->gs;
    const long long rows = (long long)B * rows_s;
    const float *unit = p->ident + 2 * 4096, *zero = unit + 64;
    {
      StageScope sc(tm, ST_LIFT, st);
      if (p->use_tc_lift)
        B2_TRY(launch_lift_tc(la, p->tmAct[0], p->tmW0, gs, st, 0, -1, p->bf16));
      else
        B2_TRY(launch_lift(la, st));
    }
    for (int l = 0; l < L; ++l) {
      const LayerPacked& Lp = p->layers[l];
      for (int s = 0; s < 2; ++s) {
        const ModeSlice& sl = p->slices[s];
        // both slices transform the layer INPUT act[0]; slice 0's layer pass leaves it intact (it writes act[1])
        B2_TRY(run_spectral(gs, sl.tab, B, p->act[0], sl.spec[l], p->bufAD, p->bufBC, p->bufS, p->bufO, st, tm, true,
                            &p->tmFwX[0], &sl.tmFwF, p->use_tc_tmul ? tmR4 : nullptr, p->bufA, false,
                            p->use_tc_modes ? &sl.tmModes[l] : nullptr));
        StageScope sc(tm, ST_LAYER, st);
        if (s == 0)
          B2_TRY(launch_layer_tc(p->tmAct[0], p->tmAct[1], Lp.tmW, p->tmD, sl.tab.Gt, unit, zero, rows, gs, 0, st, 0,
                                 p->bf16));
        else  // identity weights: no operand rounding in bf16 mode either (t is a partial sum, not a conv operand)
          B2_TRY(launch_layer_tc(p->tmAct[1], p->tmAct[0], p->tmWI, p->tmD, sl.tab.Gt, Lp.scale, Lp.shift, rows, gs,
                                 l < L - 1, st, 0, 0));
      }
    }
    return run_proj(p, B, p->act[0], ps, st);
  }
  const bool all_tc = p->use_tc && p->use_tc_lift && p->use_tc_fwdw && p->use_tc_proj;
  const int cb = (all_tc && p->chunk_b > 0 && p->chunk_b < B) ? p->chunk_b : B;
  if (cb < B) {
    for (int b0 = 0; b0 < B; b0 += cb) {
      const int nb = std::min(cb, B - b0);
      {
        StageScope sc(tm, ST_LIFT, st, b0 == 0);
        B2_TRY(launch_lift_tc(la, p->tmAct[0], p->tmW0, g, st, b0, nb, p->bf16));
      }
      StageScope sc(tm, ST_FWD_W, st, b0 == 0);
      B2_TRY(launch_fwdw_tc(p->tmFwX[0], p->tmFwF, p->bufA, nb * rows_s, g, st, b0 * rows_s));
    }
    int cur = 0;
    for (int l = 0; l < L; ++l) {
      const LayerPacked& Lp = p->layers[l];
      B2_TRY(run_spectral(g, p->tab, B, p->act[cur], Lp.spec, p->bufAD, p->bufBC, p->bufS, p->bufO, st, tm, true,
                          nullptr, nullptr, p->use_tc_tmul ? tmR4 : nullptr, p->bufA, /*fwdw_done=*/true,
                          p->use_tc_modes ? &Lp.tmModes : nullptr));
      for (int b0 = 0; b0 < B; b0 += cb) {
        const int nb = std::min(cb, B - b0);
        {
          StageScope sc(tm, ST_LAYER, st, b0 == 0);
          B2_TRY(launch_layer_tc(p->tmAct[cur], p->tmAct[cur ^ 1], Lp.tmW, p->tmD, p->tab.Gt, Lp.scale, Lp.shift,
                                 nb * rows_s, g, l < L - 1, st, b0 * rows_s, p->bf16));
        }
        if (l < L - 1) {
          StageScope sc(tm, ST_FWD_W, st, b0 == 0);
          B2_TRY(launch_fwdw_tc(p->tmFwX[cur ^ 1], p->tmFwF, p->bufA, nb * rows_s, g, st, b0 * rows_s));
        } else {
          StageScope sc(tm, ST_PROJ, st, b0 == 0);
          const ProjArgs pa = make_proj_args(p, B, p->act[cur ^ 1], ps);
          B2_TRY(launch_proj_tc(pa, p->tmActProj[cur ^ 1], p->tmFc1, p->tmFc2, st, b0, nb, p->bf16));
        }
      }
      cur ^= 1;
    }
    return 0;
  }
  {
    StageScope sc(tm, ST_LIFT, st);
    if (p->use_tc_lift) {
      B2_TRY(launch_lift_tc(la, p->tmAct[0], p->tmW0, g, st, 0, -1, p->bf16));
    } else {
      B2_TRY(launch_lift(la, st));
    }
  }
  int cur = 0;
  const long long rows = (long long)B * rows_s;
  for (int l = 0; l < L; ++l) {
    const LayerPacked& Lp = p->layers[l];
    B2_TRY(run_spectral(g, p->tab, B, p->act[cur], Lp.spec, p->bufAD, p->bufBC, p->bufS, p->bufO, st, tm, p->use_tc,
                        p->use_tc && p->use_tc_fwdw ? &p->tmFwX[cur] : nullptr, &p->tmFwF,
                        p->use_tc && p->use_tc_tmul ? tmR4 : nullptr, p->bufA, false,
                        p->use_tc_modes ? &Lp.tmModes : nullptr));
    {
      StageScope sc(tm, ST_LAYER, st);
      if (p->use_tc)
        B2_TRY(launch_layer_tc(p->tmAct[cur], p->tmAct[cur ^ 1], Lp.tmW, p->tmD, p->tab.Gt, Lp.scale, Lp.shift, rows, g,
                               l < L - 1, st, 0, p->bf16));
      else
        B2_TRY(launch_layer(p->act[cur], p->act[cur ^ 1], Lp.convT, p->tab.Gt, p->bufAD, Lp.scale, Lp.shift, rows,
                            g.Wp, g.Cp, g.K2, g.K2p, l < L - 1, st, p->bf16));
    }
    cur ^= 1;
  }
  return run_proj(p, B, p->act[cur], ps, st);
}

static int check_ready(b200fno_plan* p, int batch) {
  if (!p) {
    set_error("null plan");
    return B200FNO_EINVAL;
  }
  if (!p->ws || !p->weights_ready) {
    set_error("plan not ready: bind buffers and pack weights first");
    return B200FNO_ESTATE;
  }
  if (batch < 1 || batch > p->d.max_batch) {
    set_error("batch %d outside [1, max_batch=%d]", batch, p->d.max_batch);
    return B200FNO_EINVAL;
  }
  return 0;
}

int b200fno_forward(b200fno_plan_t* p, int32_t batch, const float* x, float* y, void* stream) {
  B2_TRY(check_ready(p, batch));
  if (!x || !y) {
    set_error("null tensor");
    return B200FNO_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ProjSpec ps;
  ps.out = y, ps.out_sB = (long long)p->d.t_out * p->d.h * p->d.w * p->d.c_out;
  return run_network(p, batch, x, ps, st);
}

int b200fno_rollout(b200fno_plan_t* p, int32_t batch, const float* x0, const float* affine_a, const float* affine_b,
                    int32_t n_steps, float* state, float* pred, void* stream) {
  B2_TRY(check_ready(p, batch));
  const b200fno_desc_t& d = p->d;
  if (!x0 || !pred || !affine_a || !affine_b || n_steps < 1) {
    set_error("null tensor or n_steps < 1");
    return B200FNO_EINVAL;
  }
  const long long HW0 = (long long)d.h * d.w;
  // c_in == c_out: the next model input IS the prediction slice just written (eval.py:315-319: preds.append(p) holds the
  // re-normalised prediction, and the same tensor is fed back), so the lift of step i + 1 reads it in place - sample
  // stride n_steps * slice - and neither a state buffer nor its 252 MB store per step (C2) exist
  const bool feed_from_pred = d.c_in == d.c_out && d.t_out == d.t_in && ((long long)d.t_out * HW0 * d.c_out) % 4 == 0 &&
                              ((uintptr_t)pred & 15) == 0;
  if (n_steps > 1 && (d.t_out != d.t_in || (!state && !feed_from_pred))) {
    set_error("n_steps > 1 needs t_out == t_in (got %d, %d) and, with parameter channels, a state buffer", d.t_out, d.t_in);
    return B200FNO_EINVAL;
  }
  if (d.c_in < d.c_out) {
    set_error("c_in (%d) < c_out (%d): prediction cannot be fed back", d.c_in, d.c_out);
    return B200FNO_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long HW = (long long)d.h * d.w;
  const long long state_elems = (long long)batch * d.t_in * HW * d.c_in;
  const long long step_elems = (long long)d.t_out * HW * d.c_out;
  const long long out_sB = (long long)n_steps * step_elems;
  if (n_steps > 1 && d.c_in > d.c_out)
    B2_TRY(launch_copy_params(x0, state, (long long)batch * d.t_in * HW, d.c_in, d.c_out, st));
  const float* cur = x0;
  for (int i = 0; i < n_steps; ++i) {
    ProjSpec ps;
    ps.aff_a = affine_a, ps.aff_b = affine_b, ps.out = pred + (size_t)i * step_elems, ps.out_sB = out_sB;
    if (feed_from_pred) {
      B2_TRY(run_network(p, batch, cur, ps, st, i == 0 ? 0 : out_sB));
      cur = ps.out;
    } else {
      float* next = (i + 1 < n_steps) ? state + (size_t)(i & 1) * state_elems : nullptr;
      ps.state = next;
      B2_TRY(run_network(p, batch, cur, ps, st));
      cur = next;
    }
  }
  return 0;
}

// Process-wide cache of the constant tables of the stand-alone spectral operator, keyed by device and geometry.
// Entries live until b200fno_spectral_cache_clear() (kernels in flight read them: clear only after synchronising).
namespace {
struct SpectralKey {
  int dev, ndim, T, H, W, Cp, m1, m2, m3;
  bool operator<(const SpectralKey& o) const {
    return std::tie(dev, ndim, T, H, W, Cp, m1, m2, m3) < std::tie(o.dev, o.ndim, o.T, o.H, o.W, o.Cp, o.m1, o.m2, o.m3);
  }
};
std::mutex g_spec_mu;
std::map<SpectralKey, std::unique_ptr<Tables>> g_spec_tables;
}  // namespace

static int spectral_tables_cached(const Geom& g, int m1, int m2, const Tables** out) {
  int dev = 0;
  B2_CUDA(cudaGetDevice(&dev));
  const SpectralKey key{dev, g.ndim, g.Tp, g.Hp, g.Wp, g.Cp, m1, m2, g.m3};
  std::lock_guard<std::mutex> lock(g_spec_mu);
  auto it = g_spec_tables.find(key);
  if (it == g_spec_tables.end()) {
    std::unique_ptr<Tables> t(new Tables());
    B2_TRY(build_tables(g, m1, m2, t.get()));
    it = g_spec_tables.emplace(key, std::move(t)).first;
  }
  *out = it->second.get();
  return 0;
}

void b200fno_spectral_cache_clear(void) {
  std::lock_guard<std::mutex> lock(g_spec_mu);
  for (auto& kv : g_spec_tables) {
    cudaSetDevice(kv.first.dev);
    free_tables(kv.second.get());
  }
  g_spec_tables.clear();
}

size_t b200fno_spectral_workspace_bytes(int32_t ndim, int32_t batch, int32_t ci, int32_t co, int32_t t, int32_t h,
                                        int32_t w, int32_t m1, int32_t m2, int32_t m3) {
  Geom g;
  if (make_geom(ndim, t, h, w, std::max(ci, co), m1, m2, m3, &g)) return 0;
  size_t n = 2 * align_up(g.act_elems(batch), 64) + spectral_scratch_floats(g, batch) +
             align_up((size_t)g.NM * g.Cp * 2 * g.Cp, 64);
  return n * sizeof(float);
}

int b200fno_spectral_conv(int32_t ndim, int32_t batch, int32_t ci, int32_t co, int32_t t, int32_t h, int32_t w,
                          int32_t m1, int32_t m2, int32_t m3, const float* const* weights, const float* x, float* y,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if ((ndim != 2 && ndim != 3) || batch < 1 || ci < 1 || co < 1 || !weights || !x || !y || !workspace) {
    set_error("bad argument");
    return B200FNO_EINVAL;
  }
  B2_TRY(check_device());
  Geom g;
  B2_TRY(make_geom(ndim, t, h, w, std::max(ci, co), m1, m2, m3, &g));
  const size_t need = b200fno_spectral_workspace_bytes(ndim, batch, ci, co, t, h, w, m1, m2, m3);
  if (workspace_bytes < need || ((uintptr_t)workspace & 255)) {
    set_error("spectral workspace too small (%zu < %zu) or not 256-byte aligned", workspace_bytes, need);
    return B200FNO_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // The constant tables of a geometry are built once per device and kept (spectral_table_cache): the call enqueues
  // kernels only - no allocation, no host synchronisation, CUDA-graph capturable after the first call of a geometry.
  const Tables* tabp = nullptr;
  B2_TRY(spectral_tables_cached(g, m1, m2, &tabp));
  const Tables& tab = *tabp;
  float* ws = (float*)workspace;
  float* a0 = ws;
  ws += align_up(g.act_elems(batch), 64);
  float* a1 = ws;
  ws += align_up(g.act_elems(batch), 64);
  float* bufAD = ws;
  ws += align_up(ad_elems(g, batch), 64);
  float* bufBC = ws;
  ws += align_up(g.b_elems(batch), 64);
  float* bufS = ws;
  ws += align_up(g.s_elems(batch), 64);
  float* bufO = ws;
  ws += align_up(g.s_elems(batch), 64);
  float* Wpk = ws;
  const long long S = (long long)g.Tp * g.Hp * g.Wp, rows = (long long)batch * g.Tp * g.Hp;
  B2_TRY(launch_pack_spectral(weights, ndim == 3 ? 4 : 2, Wpk, g, ci, co, m1, m2, tab.d_ft, tab.d_fh, st));
  B2_TRY(launch_nchw_to_cl(x, a0, batch, ci, S, g.Cp, st));
  // width 64: the forward W / H / T transforms run on the tensor cores (tensor maps are encoded on the host per call -
  // the buffers are the caller's); D stays a plain fp32 plane because the inverse-W stage below is the FFMA kernel
  // (the tcgen05 layer kernel always carries the bypass convolution, which SpectralConv alone does not have)
  CUtensorMap tmFwX, tmFwF, tmR4[4];
  const bool tc_fw = tc_fwdw_supported(g);
  bool tc_tm = false;
  if (tc_fw) B2_TRY(tc_make_fwdw_maps(&tmFwX, &tmFwF, a0, tab.LF_hl, rows, g));
  if (g.Cp == 64) {
    const int n_hw = g.m3 * g.Cp, n_t = g.KH * n_hw;
    if (tab.tm_fwdH.ok) B2_TRY(tmul_make_data_map(&tmR4[0], bufAD, batch * g.Tp, 2 * g.Hp, n_hw, (long long)g.Hp * 2 * n_hw));
    if (tab.tm_fwdT.ok) B2_TRY(tmul_make_data_map(&tmR4[1], bufBC, batch, 2 * g.Tp, n_t, (long long)g.Tp * 2 * n_t));
    if (tab.tm_invT.ok) B2_TRY(tmul_make_data_map(&tmR4[2], bufO, batch, 2 * g.KT, n_t, 2LL * g.KT * n_t));
    if (tab.tm_invH.ok)
      B2_TRY(tmul_make_data_map(&tmR4[3], g.ndim == 3 ? bufBC : bufO, batch * g.Tp, 2 * g.KH, n_hw, 2LL * g.KH * n_hw));
    tc_tm = tab.tm_fwdH.ok || tab.tm_fwdT.ok || tab.tm_invT.ok || tab.tm_invH.ok;
  }
  B2_TRY(run_spectral(g, tab, batch, a0, Wpk, bufAD, bufBC, bufS, bufO, st, nullptr, /*tc_planes=*/false,
                      tc_fw ? &tmFwX : nullptr, &tmFwF, tc_tm ? tmR4 : nullptr));
  B2_TRY(launch_layer(a0, a1, nullptr, tab.Gt, bufAD, nullptr, nullptr, rows, g.Wp, g.Cp, g.K2, g.K2p, 0, st));
  return launch_cl_to_nchw(a1, y, batch, co, S, g.Cp, st);
}

int b200fno_timing_enable(b200fno_plan_t* p, int on) {
  if (!p) return B200FNO_EINVAL;
  p->timing.enabled = on != 0;
  p->timing.used = 0;
  return 0;
}

int b200fno_timing_collect(b200fno_plan_t* p, double* ms, int64_t* count) {
  if (!p || !ms || !count) return B200FNO_EINVAL;
  for (int i = 0; i < ST_COUNT; ++i) ms[i] = 0.0, count[i] = 0;
  Timing& t = p->timing;
  for (size_t s = 0; s + 1 < t.used; s += 2) {
    B2_CUDA(cudaEventSynchronize(t.ev[s + 1]));
    float f = 0.f;
    B2_CUDA(cudaEventElapsedTime(&f, t.ev[s], t.ev[s + 1]));
    ms[t.stage[s / 2] & 0xff] += f;
    if (!(t.stage[s / 2] & 0x100)) count[t.stage[s / 2] & 0xff] += 1;
  }
  t.used = 0;
  return 0;
}

int64_t b200fno_host_table(int32_t ndim, int32_t t, int32_t h, int32_t w, int32_t m1, int32_t m2, int32_t m3,
                           int32_t which, float* out, int64_t cap, int32_t* ld, int32_t* freqs_t, int32_t* freqs_h) {
  return b200fno_host_table_slice(ndim, t, h, w, m1, m2, m3, 0, which, out, cap, ld, freqs_t, freqs_h);
}

int64_t b200fno_host_table_slice(int32_t ndim, int32_t t, int32_t h, int32_t w, int32_t m1, int32_t m2, int32_t m3,
                                 int32_t kw0, int32_t which, float* out, int64_t cap, int32_t* ld, int32_t* freqs_t,
                                 int32_t* freqs_h) {
  if ((ndim != 2 && ndim != 3) || which < 0 || which > 5 || kw0 < 0) {
    set_error("bad argument");
    return B200FNO_EINVAL;
  }
  Geom g;
  B2_TRY(make_geom(ndim, t, h, w, 4, m1, m2, kw0 + m3, &g));  // the LAST kept frequency must fit the grid
  g.m3 = m3, g.K2 = 2 * m3, g.K2p = round_up(g.K2, 4), g.NM = g.KT * g.KH * m3;
  Tables tab;
  std::vector<float> host[6];
  B2_TRY(compute_tables_host(g, m1, m2, &tab, host, kw0));
  const int lds[6] = {tab.ldLF, tab.ldLH, tab.ldLT, tab.ldLTi, tab.ldLHi, g.K2p};
  if (ld) *ld = lds[which];
  if (freqs_t) std::copy(tab.ft.begin(), tab.ft.end(), freqs_t);
  if (freqs_h) std::copy(tab.fh.begin(), tab.fh.end(), freqs_h);
  const int64_t n = (int64_t)host[which].size();
  if (out) std::copy(host[which].begin(), host[which].begin() + std::min(n, cap), out);
  return n;
}

int b200fno_plan_stage_impl(const b200fno_plan_t* p, int32_t stage) {
  if (!p || !p->ws || stage < 0 || stage >= ST_COUNT) {
    set_error("b200fno_plan_stage_impl: bound plan and a stage in [0, %d) required", (int)ST_COUNT);
    return B200FNO_EINVAL;
  }
  const Tables& tb = p->split ? p->slices[0].tab : p->tab;
  switch (stage) {
    case ST_LIFT: return p->use_tc_lift;
    case ST_FWD_W: return p->use_tc && p->use_tc_fwdw;  // (tb below: the H / T plans of slice 0 when the plan is split)
    case ST_FWD_H: return p->use_tc && p->use_tc_tmul && tmul_use(tb.tm_fwdH, p->d.max_batch * p->g.Tp);
    case ST_FWD_T: return p->use_tc && p->use_tc_tmul && tmul_use(tb.tm_fwdT, p->d.max_batch);
    case ST_MODES: return p->use_tc_modes;
    case ST_INV_T: return p->use_tc && p->use_tc_tmul && tmul_use(tb.tm_invT, p->d.max_batch);
    case ST_INV_H: return p->use_tc && p->use_tc_tmul && tmul_use(tb.tm_invH, p->d.max_batch * p->g.Tp);
    case ST_LAYER: return p->use_tc;
    case ST_PROJ: return p->use_tc_proj;
  }
  return 0;
}

double b200fno_algorithmic_bytes(const b200fno_plan_t* p, int32_t batch) {
  if (!p) return 0.0;
  const b200fno_desc_t& d = p->d;
  const Geom& g = p->g;
  // B*[in + out + L*2*C*T'H'W'*4] + L*C^2*K*8, K = corner modes as configured
  const double pts = (double)d.h * d.w;
  const double in = (double)d.t_in * pts * d.c_in * 4, out = (double)d.t_out * pts * d.c_out * 4;
  const double act = (double)d.width * g.Tp * g.Hp * g.Wp * 4;
  const double K = (d.ndim == 3 ? 4.0 * d.modes1 : 2.0) * d.modes2 * d.modes3;
  return batch * (in + out + d.n_layers * 2.0 * act) + (double)d.n_layers * d.width * d.width * K * 8.0;
}

// ---------------------------------------------------------------------------------------------
// Training path (train.py:321-334): train-mode forward + backward.  fp32 FFMA kernels throughout.
// ---------------------------------------------------------------------------------------------
struct TrainLayout {
  size_t A, S, BNC, WSP, G, DH, DF, ST, total;
};
static TrainLayout train_layout(const b200fno_plan* p) {
  const Geom& g = p->g;
  const int B = p->d.max_batch, L = p->d.n_layers;
  const size_t P = (size_t)B * g.Tp * g.Hp * g.Wp;
  TrainLayout t;
  t.A = align_up(g.act_elems(B), 64);
  t.S = align_up(g.s_elems(B), 64);
  t.BNC = align_up((size_t)6 * g.Cp, 64);
  t.WSP = align_up((size_t)g.NM * g.Cp * 2 * g.Cp, 64);
  t.G = align_up(P * (size_t)std::max(128, p->Klp), 64);
  t.DH = align_up(P * 128, 64);
  t.DF = align_up(P * (size_t)p->Fp, 64);
  t.ST = align_up((size_t)4 * g.Cp, 64);  // 2*Cp doubles
  t.total = (size_t)(2 * L + 3) * t.A + (size_t)L * (t.S + t.BNC) + 2 * t.WSP + t.G + t.DH + t.DF + t.ST;
  return t;
}

size_t b200fno_train_workspace_bytes(const b200fno_plan_t* p) { return p ? train_layout(p).total * sizeof(float) : 0; }

static int build_train_tables(b200fno_plan* p) {
  if (p->tr.tbase) return 0;
  const Geom& g = p->g;
  Tables tmp;
  std::vector<float> h[6];
  B2_TRY(compute_tables_host(g, p->d.modes1, p->d.modes2, &tmp, h));
  const int KH = g.KH, KT = g.KT, Hp = g.Hp, Tp = g.Tp, Wp = g.Wp, K2 = g.K2;
  auto transpose = [](const std::vector<float>& src, int ld_src, int rows, int cols, int ld_dst) {
    std::vector<float> out((size_t)cols * ld_dst, 0.f);  // out[c][r] = src[r][c]
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) out[(size_t)c * ld_dst + r] = src[(size_t)r * ld_src + c];
    return out;
  };
  std::vector<float> t[6];
  t[0] = transpose(h[5], g.K2p, Wp, K2, tmp.ldLF);            // GtT  [K2][ldLF]
  t[1] = transpose(h[4], tmp.ldLHi, 2 * Hp, 2 * KH, tmp.ldLH);  // LHiT [2KH][ldLH]
  t[4] = transpose(h[1], tmp.ldLH, 2 * KH, 2 * Hp, tmp.ldLHi);  // LHT  [2Hp][ldLHi]
  t[5] = transpose(h[0], tmp.ldLF, K2, Wp, g.K2p);             // LFT  [Wp][K2p]
  if (g.ndim == 3) {
    t[2] = transpose(h[3], tmp.ldLTi, 2 * Tp, 2 * KT, tmp.ldLT);  // LTiT [2KT][ldLT]
    t[3] = transpose(h[2], tmp.ldLT, 2 * KT, 2 * Tp, tmp.ldLTi);  // LTT  [2Tp][ldLTi]
  }
  size_t off[7] = {0};
  for (int i = 0; i < 6; ++i) off[i + 1] = off[i] + (size_t)round_up((int)t[i].size() + 4, 64);
  std::vector<int> slots(Tp + Hp, -1);
  for (int s = 0; s < KT; ++s) slots[tmp.ft[s]] = s;
  for (int s = 0; s < KH; ++s) slots[Tp + tmp.fh[s]] = s;
  const size_t bytes = off[6] * sizeof(float) + slots.size() * sizeof(int);
  B2_CUDA(cudaMalloc((void**)&p->tr.tbase, bytes));
  B2_CUDA(cudaMemset(p->tr.tbase, 0, bytes));
  for (int i = 0; i < 6; ++i)
    if (!t[i].empty())
      B2_CUDA(cudaMemcpy(p->tr.tbase + off[i], t[i].data(), t[i].size() * sizeof(float), cudaMemcpyHostToDevice));
  p->tr.GtT = p->tr.tbase + off[0], p->tr.LHiT = p->tr.tbase + off[1], p->tr.LTiT = p->tr.tbase + off[2];
  p->tr.LTT = p->tr.tbase + off[3], p->tr.LHT = p->tr.tbase + off[4], p->tr.LFT = p->tr.tbase + off[5];
  p->tr.slot_t = (int*)(p->tr.tbase + off[6]);
  p->tr.slot_h = p->tr.slot_t + Tp;
  B2_CUDA(cudaMemcpy(p->tr.slot_t, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

int b200fno_train_bind(b200fno_plan_t* p, void* workspace, size_t workspace_bytes) {
  if (!p || !workspace) {
    set_error("null argument");
    return B200FNO_EINVAL;
  }
  if (!p->ws) {
    set_error("b200fno_plan_bind must be called before b200fno_train_bind");
    return B200FNO_ESTATE;
  }
  const TrainLayout t = train_layout(p);
  if (workspace_bytes < t.total * sizeof(float) || ((uintptr_t)workspace & 255)) {
    set_error("training workspace too small (%zu < %zu) or not 256-byte aligned", workspace_bytes,
              t.total * sizeof(float));
    return B200FNO_EINVAL;
  }
  B2_TRY(build_train_tables(p));
  const int L = p->d.n_layers;
  float* w = (float*)workspace;
  auto& tr = p->tr;
  tr.xs.resize(L + 1), tr.zs.resize(L), tr.Ssave.resize(L), tr.bnc.resize(L);
  for (int l = 0; l <= L; ++l) tr.xs[l] = w, w += t.A;
  for (int l = 0; l < L; ++l) tr.zs[l] = w, w += t.A;
  tr.g0 = w, w += t.A;
  tr.g1 = w, w += t.A;
  for (int l = 0; l < L; ++l) tr.Ssave[l] = w, w += t.S;
  for (int l = 0; l < L; ++l) tr.bnc[l] = w, w += t.BNC;
  tr.Wadj = w, w += t.WSP;
  tr.dWpk = w, w += t.WSP;
  tr.G = w, w += t.G;
  tr.dH = w, w += t.DH;
  tr.dF = w, w += t.DF;
  tr.stats = (double*)w;
  tr.bound = true, tr.fwd_done = false;
  return 0;
}

int b200fno_train_forward(b200fno_plan_t* p, int32_t batch, const float* x, float* y, float* const* bn_running_mean,
                          float* const* bn_running_var, float momentum, void* stream) {
  B2_TRY(check_ready(p, batch));
  if (!x || !y) {
    set_error("null tensor");
    return B200FNO_EINVAL;
  }
  if (!p->tr.bound) {
    set_error("b200fno_train_bind must be called before b200fno_train_forward");
    return B200FNO_ESTATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const Geom& g = p->g;
  const b200fno_desc_t& d = p->d;
  auto& tr = p->tr;
  const int B = batch, L = d.n_layers, C = d.width;
  const long long rows = (long long)B * g.Tp * g.Hp, P = rows * g.Wp;
  tr.fwd_done = false;
  B2_TRY(launch_lift(make_lift_args(p, B, x, tr.xs[0]), st));
  for (int l = 0; l < L; ++l) {
    const LayerPacked& Lp = p->layers[l];
    B2_TRY(run_spectral(g, p->tab, B, tr.xs[l], Lp.spec, p->bufAD, p->bufBC, tr.Ssave[l], p->bufO, st));
    // z = conv1x1(x) + bias + irfft_W(D)   (fno.py:114-116), BatchNorm and GELU follow as separate passes
    B2_TRY(launch_layer(tr.xs[l], tr.zs[l], Lp.convT, p->tab.Gt, p->bufAD, nullptr, Lp.cbias, rows, g.Wp, g.Cp, g.K2,
                        g.K2p, 0, st, p->bf16));
    B2_TRY(launch_colstats(tr.zs[l], P, g.Cp, tr.stats, st));
    B2_TRY(launch_bn_finalize(tr.stats, P, d.bn_eps, Lp.gamma, Lp.beta, C, g.Cp, tr.bnc[l],
                              bn_running_mean ? bn_running_mean[l] : nullptr,
                              bn_running_var ? bn_running_var[l] : nullptr, momentum, st));
    B2_TRY(launch_bn_apply(tr.zs[l], tr.xs[l + 1], P, g.Cp, tr.bnc[l], l < L - 1, st));
  }
  const long long out_sB = (long long)d.t_out * d.h * d.w * d.c_out;
  ProjSpec ps;
  ps.out = y, ps.out_sB = out_sB;
  B2_TRY(run_proj(p, B, tr.xs[L], ps, st, /*allow_tc=*/false));
  tr.fwd_done = true, tr.last_batch = B;
  return 0;
}

int b200fno_train_backward(b200fno_plan_t* p, int32_t batch, const float* x, const float* dy,
                           const b200fno_grads_t* gr, void* const* grads_ready, void* stream) {
  B2_TRY(check_ready(p, batch));
  if (!x || !dy || !gr) {
    set_error("null argument");
    return B200FNO_EINVAL;
  }
  auto& tr = p->tr;
  if (!tr.bound || !tr.fwd_done || tr.last_batch != batch) {
    set_error("b200fno_train_backward needs a preceding b200fno_train_forward of the same batch (%d)", batch);
    return B200FNO_ESTATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const Geom& g = p->g;
  const b200fno_desc_t& d = p->d;
  const int B = batch, L = d.n_layers, C = d.width, Cp = g.Cp, nf = p->Fin + p->ng;
  const long long rows = (long long)B * g.Tp * g.Hp, P = rows * g.Wp;
  const long long n_hw = (long long)g.m3 * Cp, n_t = (long long)g.KH * n_hw;
  const Tables& tab = p->tab;
  auto zero = [&](float* ptr, size_t n) -> int {
    if (ptr) B2_CUDA(cudaMemsetAsync(ptr, 0, n * sizeof(float), st));
    return 0;
  };
  static const bool dbg_on = getenv("B200FNO_DEBUG_FINITE") != nullptr;
  if (dbg_on) {
    if (!p->dbg_first) B2_CUDA(cudaMalloc((void**)&p->dbg_first, sizeof(int)));
    B2_CUDA(cudaMemsetAsync(p->dbg_first, 0x7f, sizeof(int), st));
  }
  int dbg_seq = 0;  // checks are numbered in launch order so that the minimum is the first in time
  auto chk = [&](int stage, const float* ptr, size_t n) -> int {
    if (dbg_on && ptr) {
      finite_check_kernel<<<296, 256, 0, st>>>(ptr, n, (++dbg_seq) * 100000 + stage, p->dbg_first);
      B2_CUDA(cudaGetLastError());
    }
    return 0;
  };
  B2_TRY(zero(gr->fc2_w, (size_t)p->Fout * 128));
  B2_TRY(zero(gr->fc2_b, p->Fout));
  B2_TRY(zero(gr->fc1_w, (size_t)128 * C));
  B2_TRY(zero(gr->fc1_b, 128));
  B2_TRY(zero(gr->fc0_w, (size_t)C * nf));
  B2_TRY(zero(gr->fc0_b, C));
  for (int l = 0; l < L; ++l) {
    if (gr->conv_w) B2_TRY(zero(gr->conv_w[l], (size_t)C * C));
    if (gr->conv_b) B2_TRY(zero(gr->conv_b[l], C));
  }
  // ---- projection backward (fno.py:121-128)
  ProjBwdArgs pb{};
  pb.act = tr.xs[L], pb.dy = dy, pb.fc1T = p->fc1T, pb.fc1b = p->fc1b, pb.fc2W = p->fc2W, pb.fc1W = p->fc1W;
  pb.out_off = p->out_off, pb.G = tr.G, pb.dH = tr.dH, pb.dF = tr.dF, pb.dact = tr.g0;
  pb.B = B, pb.T = p->Tv, pb.H = d.h, pb.W = d.w, pb.Tp = g.Tp, pb.Hp = g.Hp, pb.Wp = g.Wp, pb.Cp = Cp;
  pb.Fout = p->Fout, pb.Fp = p->Fp, pb.c_out = d.c_out, pb.bf16 = p->bf16;
  const long long HW = (long long)d.h * d.w;
  pb.out_sB = (long long)d.t_out * HW * d.c_out;
  pb.out_sT = d.ndim == 3 ? (long long)(d.t_out / d.t_in) * HW * d.c_out : 0;
  B2_TRY(chk(1, dy, (size_t)B * pb.out_sB));
  B2_TRY(launch_proj_bwd(pb, st));
  B2_TRY(chk(2, tr.g0, (size_t)P * Cp));
  B2_TRY(chk(3, tr.dH, (size_t)P * 128));
  if (gr->fc2_w) B2_TRY(launch_wgrad(tr.dF, p->Fp, p->Fp, p->Fout, tr.G, 128, 128, 128, P, gr->fc2_w, 128, nullptr, st, p->bf16));
  if (gr->fc2_b) B2_TRY(launch_colsum(tr.dF, p->Fp, p->Fout, P, gr->fc2_b, st));
  if (gr->fc1_w) B2_TRY(launch_wgrad(tr.dH, 128, 128, 128, tr.xs[L], Cp, Cp, C, P, gr->fc1_w, C, nullptr, st, p->bf16));
  if (gr->fc1_b) B2_TRY(launch_colsum(tr.dH, 128, 128, P, gr->fc1_b, st));
  // grads_ready[i]: caller's cudaEvent_t recorded as soon as the gradients of group i are final, so that a
  // data-parallel caller can start their all-reduce on another stream under the rest of the backward pass
  auto ready = [&](int i) -> int {
    if (grads_ready && grads_ready[i]) B2_CUDA(cudaEventRecord((cudaEvent_t)grads_ready[i], st));
    return 0;
  };
  B2_TRY(ready(L));  // fc1, fc2
  float *gy = tr.g0, *other = tr.g1;
  for (int l = L - 1; l >= 0; --l) {
    const LayerPacked& Lp = p->layers[l];
    // GELU' (not after the last layer, fno.py:118-119) and BatchNorm backward, in place: gy becomes dz
    const int sid = 100 * (l + 1);
    B2_TRY(chk(sid + 0, gy, (size_t)P * Cp));
    B2_TRY(chk(sid + 1, tr.zs[l], (size_t)P * Cp));
    B2_TRY(launch_bn_backward(gy, tr.zs[l], gy, P, C, Cp, tr.bnc[l], l < L - 1, tr.stats,
                              gr->bn_weight ? gr->bn_weight[l] : nullptr, gr->bn_bias ? gr->bn_bias[l] : nullptr, st));
    B2_TRY(chk(sid + 2, gy, (size_t)P * Cp));
    if (gr->conv_w && gr->conv_w[l])
      B2_TRY(launch_wgrad(gy, Cp, Cp, C, tr.xs[l], Cp, Cp, C, P, gr->conv_w[l], C, nullptr, st, p->bf16));
    if (gr->conv_b && gr->conv_b[l]) B2_TRY(launch_colsum(gy, Cp, C, P, gr->conv_b[l], st));
    // adjoint of irfft_W, ifft_H, ifft_T: the forward kernels with transposed tables
    B2_TRY(launch_lmul(tr.GtT, tab.ldLF, g.K2, g.Wp, gy, (long long)g.Wp * Cp, Cp, p->bufAD, (long long)g.K2 * Cp, Cp, Cp,
                       (int)rows, st));
    B2_TRY(chk(sid + 3, p->bufAD, g.a_elems(B)));
    float* dO = p->bufO;
    B2_TRY(launch_lmul(tr.LHiT, tab.ldLH, 2 * g.KH, 2 * g.Hp, p->bufAD, (long long)g.Hp * 2 * n_hw, n_hw,
                       g.ndim == 3 ? p->bufBC : dO, 2LL * g.KH * n_hw, n_hw, (int)n_hw, B * g.Tp, st));
    if (g.ndim == 3)
      B2_TRY(launch_lmul(tr.LTiT, tab.ldLT, 2 * g.KT, 2 * g.Tp, p->bufBC, (long long)g.Tp * 2 * n_t, n_t, dO,
                         2LL * g.KT * n_t, n_t, (int)n_t, B, st));
    // spectral weights: dW = conj(S) (x) dO per mode;  dS = dO (x) conj(W)^T
    B2_TRY(chk(sid + 4, dO, g.s_elems(B)));
    B2_TRY(chk(sid + 5, tr.Ssave[l], g.s_elems(B)));
    if (gr->spec_w) {
      B2_TRY(launch_modes_wgrad(tr.Ssave[l], dO, tr.dWpk, B, g.NM, Cp, st));
      B2_TRY(chk(sid + 6, tr.dWpk, (size_t)g.NM * Cp * 2 * Cp));
      B2_TRY(launch_unpack_spectral_grad(tr.dWpk, gr->spec_w + (size_t)l * p->ncorner, p->ncorner, g, C, C, d.modes1,
                                         d.modes2, tr.slot_t, tr.slot_h, st));
    }
    B2_TRY(ready(l));  // spectral, conv and BatchNorm gradients of layer l
    B2_TRY(chk(sid + 7, Lp.spec, (size_t)g.NM * Cp * 2 * Cp));
    B2_TRY(launch_pack_spectral_adj(Lp.spec, tr.Wadj, g.NM, Cp, st));
    B2_TRY(chk(sid + 8, tr.Wadj, (size_t)g.NM * Cp * 2 * Cp));
    B2_TRY(chk(sid + 9, dO, g.s_elems(B)));
    B2_TRY(launch_modes(dO, tr.Wadj, p->bufS, B, g.NM, Cp, st));
    B2_TRY(chk(sid + 10, p->bufS, g.s_elems(B)));
    // adjoint of fft_T, fft_H
    const float* dBh = p->bufS;
    if (g.ndim == 3) {
      B2_TRY(launch_lmul(tr.LTT, tab.ldLTi, 2 * g.Tp, 2 * g.KT, p->bufS, 2LL * g.KT * n_t, n_t, p->bufBC,
                         (long long)g.Tp * 2 * n_t, n_t, (int)n_t, B, st));
      dBh = p->bufBC;
    }
    B2_TRY(launch_lmul(tr.LHT, tab.ldLHi, 2 * g.Hp, 2 * g.KH, dBh, 2LL * g.KH * n_hw, n_hw, p->bufAD,
                       (long long)g.Hp * 2 * n_hw, n_hw, (int)n_hw, B * g.Tp, st));
    // dx = dz . conv_w  +  rfft_W^T(dA): the layer kernel with the untransposed conv weight and LF^T
    B2_TRY(chk(sid + 11, p->bufAD, g.a_elems(B)));
    B2_TRY(chk(sid + 12, Lp.convW, (size_t)Cp * Cp));
    B2_TRY(launch_layer(gy, other, Lp.convW, tr.LFT, p->bufAD, nullptr, nullptr, rows, g.Wp, Cp, g.K2, g.K2p, 0, st,
                        p->bf16));  // bf16 mode: dz enters the conv adjoint as a bf16 tensor, the DFT adjoint in fp32
    B2_TRY(chk(sid + 13, other, (size_t)P * Cp));
    std::swap(gy, other);
  }
  // ---- lift backward (fno.py:106-109): d fc0 = dx_0^T . [features | grid | 1]
  if (gr->fc0_w) {
    B2_TRY(launch_lift_features(make_lift_args(p, B, x, nullptr), tr.G, st));
    B2_TRY(chk(9001, tr.G, (size_t)P * p->Klp));
    B2_TRY(chk(9002, gy, (size_t)P * Cp));
    // column nf of the feature matrix is the constant 1 at valid points (0 in the pad region): the bias gradient
    B2_TRY(launch_wgrad(gy, Cp, Cp, C, tr.G, p->Klp, p->Klp, nf, P, gr->fc0_w, nf, gr->fc0_b, st, p->bf16));
  }
  // gradient w.r.t. the input field (every element of x feeds exactly one lift feature: no zero fill needed)
  if (gr->x) B2_TRY(launch_lift_bwd_input(make_lift_args(p, B, x, nullptr), gy, gr->x, st));
  return 0;
}

int b200fno_debug_first_nonfinite(b200fno_plan_t* p) {
  if (!p || !p->dbg_first) return -1;
  int v = 0;
  if (cudaMemcpy(&v, p->dbg_first, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  return v == 0x7f7f7f7f ? 0 : v % 100000;
}

int b200fno_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                      float beta2, float eps, int64_t step, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0 || step < 1) {
    set_error("bad argument");
    return B200FNO_EINVAL;
  }
  return launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, (cudaStream_t)stream);
}

size_t b200fno_metrics_workspace_bytes(int32_t b, int32_t t, int32_t h, int32_t w, int32_t channels, int32_t c) {
  return metrics_workspace_bytes(b, t, h, w, channels, c);
}

int b200fno_eval_metrics(const float* pred, const float* target, int32_t b, int32_t t, int32_t h, int32_t w,
                         int32_t channels, int32_t c, void* workspace, size_t workspace_bytes, float* out13,
                         void* stream) {
  if (!pred || !target || !out13) {
    set_error("null tensor");
    return B200FNO_EINVAL;
  }
  B2_TRY(check_device());
  return launch_eval_metrics(pred, target, b, t, h, w, channels, c, workspace, workspace_bytes, out13,
                             (cudaStream_t)stream);
}

}  // extern "C"
