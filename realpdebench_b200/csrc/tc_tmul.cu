// Small axis transforms of the truncated DFT on the tensor cores (H and T axes, forward and inverse):
//
//   Out[g][m][n] = sum_k L[m][k] * R[g][k][n]        n contiguous, N % 128 == 0
//
// (complex axis transforms written as real GEMMs, see csrc/tables.cu).  Same scheme as tc_fwdw.cu: the
// contraction index k is NOT the contiguous one, so a tile of 128 columns n is transposed by writing it
// to TMEM with lane = n; M_mma = 128 columns, N_mma = MT table rows (a tile of m), K = chunks of 64.
// Work item = (g, 128-column tile, m tile).  The table chunk [MT x 64] (hi|lo planes, K-major) is streamed
// with the data chunk.  3xTF32: data hi = raw fp32, lo = x - trunc(x); table planes pre-split on the host.
//
// The epilogue writes Out[m][n] rows (coalesced over n); the output row address is
//   g*sOg + (m / mdiv)*sOm + (m % mdiv)*sOmLo        and an optional second plane (+split_off) receives
// x - trunc(x) (the hi|lo D planes the layer kernel consumes).
//
//   warp 0     TMA producer      warp 1  MMA issuer (.ts)      warp 2  TMEM allocation
//   warps 4-7  transpose + split: smem -> TMEM (hi | lo)       warps 8-15  epilogue (column halves)
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200fno {
using namespace tc;

constexpr int TM_THREADS = 512;
constexpr int TM_CH = 64;      // k per chunk
constexpr int TM_XS = 32768;   // data part of a stage: 4 quarters x 64 k x 128 B

struct TmulArgs {
  float* out;
  int G, NT, n_mt, MT, M, nchunk, Mpad, NS, stage_bytes;  // NT = N/128 column tiles, MT rows per m tile
  int share;  // single-chunk contraction with several m tiles: the data tile is staged ONCE per (g, column tile) and
              // every m tile multiplies it (work unit = (g, nt), inner loop over m tiles)
  long long sOg, sOm, sOmLo, split_off;
  int mdiv;
};

__device__ __forceinline__ int round_up_dev(int x, int m) { return (x + m - 1) / m * m; }

__global__ void __launch_bounds__(TM_THREADS, 1)
    tc_tmul_kernel(TmulArgs a, const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmL) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  __shared__ uint64_t x_full[4], x_empty[4], a_full[2], a_empty[2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int MT = a.MT, nchunk = a.nchunk, NS = a.NS, SB = a.stage_bytes;
  const int n_items = a.G * a.NT * a.n_mt;
  const bool share = a.share != 0;
  const int n_units = a.G * a.NT;  // share mode: (g, nt) units, a.n_mt accumulator-level items each
  const int n_my_u = (int)blockIdx.x < n_units ? (n_units - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int n_my = share ? n_my_u * a.n_mt
                         : ((int)blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&x_full[i], 1), mbar_init(&x_empty[i], 5);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128), mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, 512);
  if (warp == 0 && lane == 0) prefetch_tensormap(&tmR), prefetch_tensormap(&tmL);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t T_ACC = tmem, T_A = tmem + 256;  // acc: 2 x (<=128) cols; A: 2 x (64 hi | 64 lo)
  pdl_wait();  // everything below may touch the previous kernel's output (PDL, common.cuh)

  auto decode = [&](int ip, int& g, int& nt, int& mt) {
    if (share) {  // item ip of this CTA = m tile (ip % n_mt) of its unit (ip / n_mt)
      const int wu = blockIdx.x + (ip / a.n_mt) * gridDim.x;
      mt = ip % a.n_mt, nt = wu % a.NT, g = wu / a.NT;
      return;
    }
    const int wi = blockIdx.x + ip * gridDim.x;  // m tile fastest: the data tile is re-read from L2
    mt = wi % a.n_mt;
    nt = (wi / a.n_mt) % a.NT;
    g = wi / (a.n_mt * a.NT);
  };

  if (warp == 0) {
    int sx = 0, px = 0;  // ring stage / phase advanced incrementally (NS is a run-time value)
    for (int ip = 0; ip < n_my; ++ip) {
      int g, nt, mt;
      decode(ip, g, nt, mt);
      for (int ch = 0; ch < nchunk; ++ch, sx = (sx + 1 == NS ? 0 : sx + 1), px ^= (sx == 0)) {
        mbar_wait(&x_empty[sx], px ^ 1);
        if (elect_one_sync()) {
          uint8_t* st = smem + sx * SB;
          const bool with_data = !share || mt == 0;  // share mode: only the first m tile of a unit carries the data
          mbar_arrive_expect_tx(&x_full[sx], (uint32_t)((with_data ? TM_XS : 0) + 4 * MT * 128));
          // data: columns n = nt*128 + q*32 + lane; viewed as [g][k][nb = n/64][c = n%64], box (32 c, 64 k, 2 nb)
          if (with_data) tma_load_4d(st, &tmR, &x_full[sx], 0, ch * TM_CH, 2 * nt, g);
          if (with_data) tma_load_4d(st + 16384, &tmR, &x_full[sx], 32, ch * TM_CH, 2 * nt, g);
          for (int hl = 0; hl < 2; ++hl)
            for (int s = 0; s < 2; ++s)
              tma_load_2d(st + TM_XS + (hl * 2 + s) * MT * 128, &tmL, &x_full[sx], 32 * (2 * ch + s),
                          hl * a.Mpad + mt * MT);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(128, MT, 0, 0);
    const uint64_t sub = (uint64_t)(MT * 128 >> 4);
    int cc = 0, sx = 0, px = 0;
    for (int ip = 0; ip < n_my; ++ip) {
      const int ab = ip & 1, pab = (ip >> 1) & 1;
      mbar_wait(&acc_empty[ab], pab ^ 1);
      for (int ch = 0; ch < nchunk; ++ch, ++cc, sx = (sx + 1 == NS ? 0 : sx + 1), px ^= (sx == 0)) {
        // A buffer: one per chunk; in share mode one per UNIT (n_mt consecutive items use the same staged data)
        const int ca = share ? ip / a.n_mt : cc, t = ca & 1, pt = (ca >> 1) & 1;
        const bool first_of_a = !share || ip % a.n_mt == 0, last_of_a = !share || ip % a.n_mt == a.n_mt - 1;
        mbar_wait(&x_full[sx], px);
        if (first_of_a) mbar_wait(&a_full[t], pt);
        tc_fence_after();
        // m tiles of at most 64 rows keep the low-order cross terms in their own accumulator (columns 64.. of the
        // slot): the tensor core's adder truncates, see tc_fwdw.cu
        const uint32_t acc = T_ACC + ab * 128, acc_lo = MT <= 64 ? acc + 64 : acc, Ahi = T_A + t * 128, Alo = Ahi + 64;
        const uint64_t dL_hi = make_smem_desc(smem_u32(smem) + sx * SB + TM_XS, 0, 1024);
        const uint64_t dL_lo = dL_hi + 2 * sub;
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(acc_lo, Alo + ks * 8, dL_hi + (uint64_t)(ks >> 2) * sub + (uint64_t)((ks & 3) * 2), idesc,
                         (ch | ks) != 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(acc_lo, Ahi + ks * 8, dL_lo + (uint64_t)(ks >> 2) * sub + (uint64_t)((ks & 3) * 2), idesc, 1);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(acc, Ahi + ks * 8, dL_hi + (uint64_t)(ks >> 2) * sub + (uint64_t)((ks & 3) * 2), idesc,
                         MT <= 64 ? (ch | ks) != 0 : 1);
          if (last_of_a) umma_commit(&a_empty[t]);
          umma_commit(&x_empty[sx]);
          if (ch == nchunk - 1) umma_commit(&acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // TMEM lane = column n of the tile: warp q holds nb = q/2, channel half q%2, lane = c%32
    const int q = warp - 4;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t col_base = (uint32_t)((q & 1) * 16384 + (q >> 1) * 8192 + lane * 4);
    int cc = 0, sx = 0, px = 0;
    for (int ip = 0; ip < n_my; ++ip) {
      for (int ch = 0; ch < nchunk; ++ch, ++cc, sx = (sx + 1 == NS ? 0 : sx + 1), px ^= (sx == 0)) {
        const int ca = share ? ip / a.n_mt : cc, t = ca & 1, pt = (ca >> 1) & 1;
        mbar_wait(&x_full[sx], px);
        if (share && ip % a.n_mt != 0) {  // a later m tile of the unit: nothing to stage, the stage still needs our arrival
          __syncwarp();
          if (lane == 0) mbar_arrive(&x_empty[sx]);
          continue;
        }
        mbar_wait(&a_empty[t], pt ^ 1);
        tc_fence_after();
        const uint32_t src = smem_u32(smem) + sx * SB + col_base;
        const uint32_t Ahi = T_A + t * 128 + lane_addr, Alo = Ahi + 64;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v[i]) : "r"(src + (uint32_t)((half * 32 + i) * 128)));
          if (half == 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_empty[sx]);
          }
          uint32_t hv[32];
#pragma unroll
          for (int i = 0; i < 32; i += 2) tf32_split2(v[i], v[i + 1], hv[i], hv[i + 1]);
          tmem_st32(Ahi + half * 32, hv);
          tmem_st32(Alo + half * 32, v);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&a_full[t]);
      }
    }
  } else if (warp >= 8) {
    // 8 epilogue warps: TMEM lane quarter q = warp % 4, column (= m) halves eh = (warp - 8) / 4
    const int q = warp & 3, eh = (warp - 8) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int n_in_tile = (q >> 1) * 64 + (q & 1) * 32 + lane;
    for (int ip = 0; ip < n_my; ++ip) {
      int g, nt, mt;
      decode(ip, g, nt, mt);
      const int ab = ip & 1, pab = (ip >> 1) & 1;
      mbar_wait(&acc_full[ab], pab);
      tc_fence_after();
      float* og = a.out + (size_t)g * a.sOg + (size_t)nt * 128 + n_in_tile;
      const int mh = round_up_dev((MT + 1) / 2, 32);  // this warp's rows: [eh*mh, min(MT, eh*mh + mh))
      const int c_end = min(MT, eh * mh + mh);
      if (eh * mh >= c_end) {  // this warp has no rows of the tile (MT <= 32): it still owes its arrival
        tc_fence_before();
        mbar_arrive(&acc_empty[ab]);
      }
      for (int c0 = eh * mh; c0 < c_end; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(T_ACC + ab * 128 + lane_addr + c0, v);
        if (MT <= 64) {  // + the low-order accumulator
          uint32_t vl[32];
          tmem_ld32(T_ACC + ab * 128 + 64 + lane_addr + c0, vl);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(vl[i]));
        }
        tmem_ld_wait();
        if (c0 + 32 >= c_end) {  // last read of this accumulator by this warp: the MMA warp may reuse it while we store
          tc_fence_before();
          mbar_arrive(&acc_empty[ab]);
        }
        // output row address of m: (m / mdiv) * sOm + (m % mdiv) * sOmLo.  mdiv is 1 (plain rows) or 2 (the (h, ri)
        // rows of the inverse-H output) in every plan: a run-time division per element made this loop the kernel's
        // critical path for many-row outputs, so the two cases are spelled out
        const int m0 = mt * MT + c0, nrow = min(min(32, c_end - c0), a.M - m0);
        auto store_rows = [&](auto MD) {
          constexpr int md = decltype(MD)::value;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < nrow) {
              const int m = m0 + i;
              const size_t off = md == 1 ? (size_t)m * a.sOm
                                 : md == 2 ? (size_t)(m >> 1) * a.sOm + (size_t)(m & 1) * a.sOmLo
                                           : (size_t)(m / a.mdiv) * a.sOm + (size_t)(m % a.mdiv) * a.sOmLo;
              const float x = __uint_as_float(v[i]);
              if (a.split_off) {
                const float hi = tf32_hi(x);
                og[off] = hi;
                og[off + a.split_off] = x - hi;
              } else {
                og[off] = x;
              }
            }
          }
        };
        if (a.mdiv == 1) store_rows(std::integral_constant<int, 1>{});
        else if (a.mdiv == 2) store_rows(std::integral_constant<int, 2>{});
        else store_rows(std::integral_constant<int, 0>{});
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// Plan for one left-multiply stage (TmulPlan, common.cuh): the table re-laid out as hi|lo planes
// [2][Mpad][Kpad], zero padded.
int tmul_plan_build(TmulPlan* tp, const std::vector<float>& L, int ldl, int M, int K, int N) {
  tp->ok = false;
  // Measured on B200: the transposed scheme wins for long contractions with few output rows (forward H: 25 ->
  // 19 us at C2, 8.9 -> 5.7 ms per rollout in 3-D) and loses for the short-K inverse transforms, where a work
  // item is a single chunk and the streamed table outweighs the data; those stay on the FFMA kernel.
  // Short contractions (K <= 64: a work item is a single chunk): every m tile re-stages the same data tile, so the
  // tensor-core kernel wins only with at most two m tiles (3-D inverse H, 2 * H' = 140 rows: 10.3 -> 9.2 ms per cylinder
  // rollout) and loses with more (C2 inverse H, 524 rows = 5 tiles: 1.96 -> 2.47 ms); B200FNO_TMUL_SHORTK=0|1 overrides.
  {
    const char* e = getenv("B200FNO_TMUL_SHORTK");
    const bool shortk_ok = e ? atoi(e) != 0 : M > 128;
    if (N % 128 != 0 || M < 1 || M > 1024 || (K <= TM_CH && !shortk_ok)) return 0;
  }
  // m tiles: as few as fit 128 accumulator columns, rows balanced over them (M = 140 -> 2 x 80, not 128 + 12)
  const int n_mt0 = ceil_div(M, 128);
  const int MT = round_up(ceil_div(M, n_mt0), 16);
  tp->M = M, tp->K = K, tp->N = N, tp->MT = MT, tp->n_mt = ceil_div(M, MT), tp->Mpad = tp->n_mt * MT;
  tp->nchunk = ceil_div(K, TM_CH), tp->Kpad = tp->nchunk * TM_CH;
  tp->stage_bytes = TM_XS + round_up(4 * MT * 128, 1024);
  tp->NS = std::min(4, (226 * 1024) / tp->stage_bytes);
  if (tp->NS < 2) return 0;
  std::vector<float> hl((size_t)2 * tp->Mpad * tp->Kpad, 0.f);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      const float x = L[(size_t)m * ldl + k];
      uint32_t u;
      memcpy(&u, &x, 4);
      u = (u + 0x1000u) & 0xFFFFE000u;  // round to nearest tf32 (tc_common.cuh: tf32_hi)
      float hi;
      memcpy(&hi, &u, 4);
      hl[(size_t)m * tp->Kpad + k] = hi;
      hl[((size_t)tp->Mpad + m) * tp->Kpad + k] = x - hi;
    }
  B2_CUDA(cudaMalloc((void**)&tp->table, hl.size() * sizeof(float)));
  B2_CUDA(cudaMemcpy(tp->table, hl.data(), hl.size() * sizeof(float), cudaMemcpyHostToDevice));
  uint64_t dims[2] = {(uint64_t)tp->Kpad, (uint64_t)2 * tp->Mpad};
  uint64_t strides[1] = {(uint64_t)tp->Kpad * 4};
  uint32_t box[2] = {32, (uint32_t)MT};
  B2_TRY(encode_tensor_map(&tp->tmL, tp->table, 2, dims, strides, box, 1));
  tp->ok = true;
  return 0;
}
bool tmul_use(const TmulPlan& tp, int G) {
  if (!tp.ok) return false;
  if (tp.nchunk > 1) return true;
  return (long long)G * (tp.N / 128) >= 2 * 148;
}
void tmul_plan_free(TmulPlan* tp) {
  if (tp->table) cudaFree(tp->table);
  tp->table = nullptr, tp->ok = false;
}

// R viewed as [G][K][N/64][64]: box = 32 channels x 64 k x 2 column blocks
int tmul_make_data_map(CUtensorMap* m, const float* R, int G, int K, int N, long long strideRg) {
  uint64_t dims[4] = {64, (uint64_t)K, (uint64_t)N / 64, (uint64_t)G};
  uint64_t strides[3] = {(uint64_t)N * 4, 64 * 4, (uint64_t)strideRg * 4};
  uint32_t box[4] = {32, TM_CH, 2, 1};
  return encode_tensor_map(m, R, 4, dims, strides, box, 0);
}

int launch_tmul_tc(const TmulPlan& tp, const CUtensorMap& tmR, float* out, int G, long long sOg, long long sOm,
                   int mdiv, long long sOmLo, long long split_off, cudaStream_t st) {
  TmulArgs a{};
  a.out = out, a.G = G, a.NT = tp.N / 128, a.n_mt = tp.n_mt, a.MT = tp.MT, a.M = tp.M, a.nchunk = tp.nchunk;
  a.Mpad = tp.Mpad, a.NS = tp.NS, a.stage_bytes = tp.stage_bytes;
  a.sOg = sOg, a.sOm = sOm, a.sOmLo = sOmLo, a.split_off = split_off, a.mdiv = mdiv;
  static const bool no_share = getenv("B200FNO_TMUL_NO_SHARE") != nullptr;  // read once, not per launch
  a.share = (tp.nchunk == 1 && tp.n_mt > 1 && !no_share) ? 1 : 0;
  const int items = a.share ? G * a.NT : G * a.NT * a.n_mt;
  const int smem = tp.NS * tp.stage_bytes + 1024;
  B2_CUDA(cudaFuncSetAttribute(tc_tmul_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  B2_CUDA(launch_kernel(tc_tmul_kernel, dim3(std::min(148, items)), dim3(TM_THREADS), (size_t)smem, st, a, tmR, tp.tmL));
  B2_LAUNCHED("tc_tmul_kernel");
  return 0;
}

}  // namespace b200fno
