// Truncated-DFT constant tables.
//
// torch.fft.rfftn (fno.py:48) followed by the corner slicing (fno.py:53-60)
// only ever reads the kept low modes; torch.fft.irfftn (fno.py:63) of a
// spectrum that is zero outside those modes only ever sums them.  Both are
// therefore exact separable truncated DFTs, one small dense matrix per axis.
// irfftn semantics: complex inverse on all but the last axis, then
// a C2R on the last one, which drops the imaginary part of the k_w = 0 (and
// Nyquist) bin and doubles the others.  Twiddles are evaluated in double with
// an exact integer reduction of the angle and rounded once to fp32.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace b200fno {

static double tw_cos(long long k, long long n, long long N) { return cos(2.0 * M_PI * (double)((k * n) % N) / (double)N); }
static double tw_sin(long long k, long long n, long long N) { return sin(2.0 * M_PI * (double)((k * n) % N) / (double)N); }

// distinct kept frequencies of an axis of length N with `m` low and `m` high modes
static std::vector<int> kept(int N, int m) {
  std::vector<char> on(N, 0);
  for (int i = 0; i < m; ++i) on[i] = 1, on[N - m + i] = 1;
  std::vector<int> f;
  for (int i = 0; i < N; ++i)
    if (on[i]) f.push_back(i);
  return f;
}

// forward complex axis transform as a real matrix:
//   rows m = ri_o*KF + slot,  cols k = pos*2 + ri_i,  e^{-i theta}
static void fwd_complex(std::vector<float>& L, int ld, const std::vector<int>& f, int N) {
  int KF = (int)f.size();
  for (int s = 0; s < KF; ++s)
    for (int p = 0; p < N; ++p) {
      double c = tw_cos(f[s], p, N), sn = tw_sin(f[s], p, N);
      L[(size_t)(0 * KF + s) * ld + p * 2 + 0] = (float)c;
      L[(size_t)(0 * KF + s) * ld + p * 2 + 1] = (float)sn;
      L[(size_t)(1 * KF + s) * ld + p * 2 + 0] = (float)-sn;
      L[(size_t)(1 * KF + s) * ld + p * 2 + 1] = (float)c;
    }
}
// inverse: rows m = pos*2 + ri_o, cols k = ri_i*KF + slot, e^{+i theta} (unscaled)
static void inv_complex(std::vector<float>& L, int ld, const std::vector<int>& f, int N) {
  int KF = (int)f.size();
  for (int p = 0; p < N; ++p)
    for (int s = 0; s < KF; ++s) {
      double c = tw_cos(f[s], p, N), sn = tw_sin(f[s], p, N);
      L[(size_t)(p * 2 + 0) * ld + 0 * KF + s] = (float)c;
      L[(size_t)(p * 2 + 0) * ld + 1 * KF + s] = (float)-sn;
      L[(size_t)(p * 2 + 1) * ld + 0 * KF + s] = (float)sn;
      L[(size_t)(p * 2 + 1) * ld + 1 * KF + s] = (float)c;
    }
}

int compute_tables_host(const Geom& g, int m1, int m2, Tables* t, std::vector<float> (&host)[6], int kw0) {
  // kw0: first kept W frequency of this table set (0 unless the plan splits modes3 > 32 into slices, api.cu)
  t->fh = kept(g.Hp, m2);
  t->ft = (g.ndim == 3) ? kept(g.Tp, m1) : std::vector<int>{0};
  const int KH = (int)t->fh.size(), KT = (int)t->ft.size();
  if (KH != g.KH || KT != g.KT) {
    set_error("internal: kept-frequency count mismatch");
    return B200FNO_EINVAL;
  }
  const int m3 = g.m3, Wp = g.Wp, Hp = g.Hp, Tp = g.Tp;
  t->ldLF = round_up(Wp, 4);
  t->ldLH = round_up(2 * Hp, 4);
  t->ldLT = round_up(2 * Tp, 4);
  t->ldLTi = round_up(2 * KT, 4);
  t->ldLHi = round_up(2 * KH, 4);
  std::vector<float> LF((size_t)g.K2 * t->ldLF, 0.f), LH((size_t)2 * KH * t->ldLH, 0.f),
      LT((size_t)2 * KT * t->ldLT, 0.f), LTi((size_t)2 * Tp * t->ldLTi, 0.f), LHi((size_t)2 * Hp * t->ldLHi, 0.f),
      Gt((size_t)Wp * g.K2p, 0.f);
  // forward W (real -> complex): rows m = ri*m3 + kw
  for (int kw = 0; kw < m3; ++kw)
    for (int w = 0; w < Wp; ++w) {
      LF[(size_t)(0 * m3 + kw) * t->ldLF + w] = (float)tw_cos(kw0 + kw, w, Wp);
      LF[(size_t)(1 * m3 + kw) * t->ldLF + w] = (float)-tw_sin(kw0 + kw, w, Wp);
    }
  fwd_complex(LH, t->ldLH, t->fh, Hp);
  inv_complex(LHi, t->ldLHi, t->fh, Hp);
  if (g.ndim == 3) {
    fwd_complex(LT, t->ldLT, t->ft, Tp);
    inv_complex(LTi, t->ldLTi, t->ft, Tp);
  }
  // inverse W (C2R) with the irfftn 1/N scaling of all transformed axes folded in
  const double scale = 1.0 / ((double)Wp * Hp * (g.ndim == 3 ? Tp : 1));
  for (int w = 0; w < Wp; ++w)
    for (int kw = 0; kw < m3; ++kw) {
      const int fw = kw0 + kw;
      const bool self_conj = (fw == 0) || (Wp % 2 == 0 && fw == Wp / 2);
      const double c = self_conj ? 1.0 : 2.0;
      Gt[(size_t)w * g.K2p + 0 * m3 + kw] = (float)(c * scale * tw_cos(fw, w, Wp));
      Gt[(size_t)w * g.K2p + 1 * m3 + kw] = (float)(-c * scale * tw_sin(fw, w, Wp));
    }
  host[0].swap(LF), host[1].swap(LH), host[2].swap(LT), host[3].swap(LTi), host[4].swap(LHi), host[5].swap(Gt);
  return 0;
}

int build_tables(const Geom& g, int m1, int m2, Tables* t, int kw0) {
  std::vector<float> host[6];
  B2_TRY(compute_tables_host(g, m1, m2, t, host, kw0));
  {  // forward-W table as 3xTF32 planes for the tensor-core kernel: [hi|lo][K2m][wpad], zero padded (K2m = K2 rounded
     // up to 16, the MMA N step)
    const int wpad = 2 * ceil_div(g.Wp, 64) * 32, K2m = round_up(g.K2, 16);
    std::vector<float> hl((size_t)2 * K2m * wpad, 0.f);
    for (int k = 0; k < g.K2; ++k)
      for (int w = 0; w < g.Wp; ++w) {
        const float x = host[0][(size_t)k * t->ldLF + w];
        uint32_t u;
        memcpy(&u, &x, 4);
        u = (u + 0x1000u) & 0xFFFFE000u;  // round to nearest tf32 (tc_common.cuh: tf32_hi)
        float hi;
        memcpy(&hi, &u, 4);
        hl[(size_t)k * wpad + w] = hi;
        hl[((size_t)K2m + k) * wpad + w] = x - hi;
      }
    B2_CUDA(cudaMalloc((void**)&t->LF_hl, hl.size() * sizeof(float)));
    B2_CUDA(cudaMemcpy(t->LF_hl, hl.data(), hl.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (g.Cp == 64) {  // tensor-core plans for the H / T axis transforms (column counts are multiples of 128)
    const int n_hw = g.m3 * g.Cp, n_t = g.KH * n_hw;
    B2_TRY(tmul_plan_build(&t->tm_fwdH, host[1], t->ldLH, 2 * g.KH, 2 * g.Hp, n_hw));
    B2_TRY(tmul_plan_build(&t->tm_invH, host[4], t->ldLHi, 2 * g.Hp, 2 * g.KH, n_hw));
    if (g.ndim == 3) {
      B2_TRY(tmul_plan_build(&t->tm_fwdT, host[2], t->ldLT, 2 * g.KT, 2 * g.Tp, n_t));
      B2_TRY(tmul_plan_build(&t->tm_invT, host[3], t->ldLTi, 2 * g.Tp, 2 * g.KT, n_t));
    }
  }
  const int KT = g.KT, KH = g.KH;
  const std::vector<float>* all[6] = {&host[0], &host[1], &host[2], &host[3], &host[4], &host[5]};
  size_t off[7] = {0};
  for (int i = 0; i < 6; ++i) off[i + 1] = off[i] + (size_t)round_up((int)all[i]->size() + 4, 64);
  size_t ints = (size_t)round_up(KT + KH, 64);
  t->bytes = off[6] * sizeof(float) + ints * sizeof(int);
  B2_CUDA(cudaMalloc((void**)&t->base, t->bytes));
  B2_CUDA(cudaMemset(t->base, 0, t->bytes));
  for (int i = 0; i < 6; ++i)
    if (!all[i]->empty())
      B2_CUDA(cudaMemcpy(t->base + off[i], all[i]->data(), all[i]->size() * sizeof(float), cudaMemcpyHostToDevice));
  t->LF = t->base + off[0];
  t->LH = t->base + off[1];
  t->LT = t->base + off[2];
  t->LTi = t->base + off[3];
  t->LHi = t->base + off[4];
  t->Gt = t->base + off[5];
  t->d_ft = (int*)(t->base + off[6]);
  t->d_fh = t->d_ft + KT;
  B2_CUDA(cudaMemcpy(t->d_ft, t->ft.data(), KT * sizeof(int), cudaMemcpyHostToDevice));
  B2_CUDA(cudaMemcpy(t->d_fh, t->fh.data(), KH * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

void free_tables(Tables* t) {
  if (t->base) cudaFree(t->base);
  if (t->LF_hl) cudaFree(t->LF_hl);
  tmul_plan_free(&t->tm_fwdH), tmul_plan_free(&t->tm_fwdT), tmul_plan_free(&t->tm_invT), tmul_plan_free(&t->tm_invH);
  t->base = nullptr;
  t->LF_hl = nullptr;
}

}  // namespace b200fno
