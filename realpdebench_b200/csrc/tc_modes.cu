// Per-mode complex channel mixing (fno.py:41-43, einsum "bixyz,ioxyz->boxyz") on the tensor cores, width 64.
//
// For one kept mode, with W = W_re + i W_im [i][o] and S = S_re + i S_im [b][i]:
//
//   O_re = S_re.W_re - S_im.W_im          O_im = S_re.W_im + S_im.W_re                     b < B, i, o < 64
//
// All four real products come out of ONE GEMM per mode: M = 128 = (W_re | W_im, o), N = (S_re | S_im, b), K = 64 = i -
// "a dense batched GEMM in k-space" - and the epilogue combines the quadrants.
// The weights are the big operand (NM x 32 KB, streamed from HBM exactly once per forward: the SURVEY 8d term
// L*C^2*K*8); they enter as the A operand THROUGH TMEM: the TMA lands the mode's [i][re|im][o] block in shared memory
// as it lies in HBM, thread (re|im, o) reads its column (conflict free), splits it into 3xTF32 hi | lo and writes TMEM
// lane = (re|im, o) - the transpose costs no extra pass and every weight is written once per plane (64 KB of TMEM
// stores per mode; tcgen05.st moves ~32 B/clk per SM, which is what bounds this kernel - a first version that wrote
// the real 2x2 embedding [[W_re, W_im], [-W_im, W_re]] (128 KB per mode) ran at half the rate).  The spectra S are
// tiny (2*Nb x 64 floats per mode): four warps fetch them from L2, split them and lay them out as the K-major
// 128B-swizzled B operand.
//
//   warp 0      TMA: weight block of the mode into a 3-stage ring (32 KB per stage)
//   warp 1      MMA issuer: per mode 8 k-steps x (hi*hi -> acc, lo*hi + hi*lo -> acc_lo), N = 2*Nb
//   warp 2      TMEM allocation (512 columns: A 2 x 128, acc 2 x 64, acc_lo 2 x 64)
//   warps 4-7   weight staging: smem -> split -> TMEM A operand
//   warps 8-11  epilogue: acc + acc_lo, quadrant exchange through shared memory -> O[b][re|im][mode][o]
//   warps 12-15 spectra staging: S[b][re|im][mode][i] -> hi | lo swizzled operand tiles, row = (re|im, b)
#include "common.cuh"
#include "tc_common.cuh"

namespace b200fno {
using namespace tc;

constexpr int MD_THREADS = 512;
constexpr int MD_NSW = 5;             // weight ring stages (at most; ModesArgs::nsw): 4 x 32 KB in flight per SM - the stream is
                                      // latency bound: with a 3-stage ring it ran at 2.7 TB/s (profiles/r02_modes3d_ncu_full.csv)
constexpr int MD_WS = 32768;          // one mode's weights: 64 i x 2 x 64 o fp32
constexpr int MD_NMAX = 32;           // batch entries per launch handled by the tensor-core kernel
constexpr int MD_NSR = 8;             // raw-spectra ring stages (at most)

struct ModesArgs {
  const float* S;  // [B][2][NM][64]
  float* O;        // [B][2][NM][64]
  int B, Nb, NM;   // Nb = batch columns per re|im block (16 | 32) >= B; MMA N = 2 * Nb
  int nsw;         // weight ring stages in use
  int nsr;         // raw-spectra ring stages in use (each 2 * B rows of 256 B)
};

__global__ void __launch_bounds__(MD_THREADS, 1)
    tc_modes_kernel(ModesArgs a, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmS) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* sW = smem;                      // MD_NSW weight stages
  uint8_t* sS = smem + a.nsw * MD_WS;      // 2 buffers x [hi | lo] x 2 k-subtiles x [N rows][128 B]
  const int Nb = a.Nb, N = 2 * a.Nb, NM = a.NM, NSW = a.nsw, MD_XLD = 2 * a.Nb + 1;
  const uint32_t S_PLANE = (uint32_t)(2 * N * 128), S_BUF = 2 * S_PLANE;
  float* sX = reinterpret_cast<float*>(sS + 2 * S_BUF);  // epilogue exchange: [128 lanes][N + 1]
  const uint32_t R_STAGE = (uint32_t)(2 * a.B * 256);     // raw spectra of one mode: rows (b, ri) of 64 floats
  uint8_t* sR = reinterpret_cast<uint8_t*>(sX) + ((128 * MD_XLD * 4 + 127) & ~127);  // a.nsr stages, TMA destination
  __shared__ uint64_t w_full[MD_NSW], w_empty[MD_NSW], a_full[2], a_empty[2], s_full[2], s_empty[2], acc_full[2],
      acc_empty[2], r_full[MD_NSR], r_empty[MD_NSR];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my = (int)blockIdx.x < NM ? (NM - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < MD_NSW; ++i) mbar_init(&w_full[i], 1), mbar_init(&w_empty[i], 4);
    for (int i = 0; i < MD_NSR; ++i) mbar_init(&r_full[i], 1), mbar_init(&r_empty[i], 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128), mbar_init(&a_empty[i], 1);
      mbar_init(&s_full[i], 128), mbar_init(&s_empty[i], 1);
      mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, 512);
  if (warp == 0 && lane == 0) prefetch_tensormap(&tmW), prefetch_tensormap(&tmS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t T_A = tmem, T_ACC = tmem + 256, T_ACCLO = tmem + 384;

  if (warp == 0) {
    pdl_wait();  // the packed weights and the spectra are written by earlier kernels of the chain
    // two rings fed by this warp: the raw spectra of a mode (small, a.nsr deep: their L2 / HBM round trip is hidden by
    // depth, not by registers) and its weights.  The spectra run ahead by up to a.nsr modes.
    int itr = 0;  // next mode whose raw spectra are to be requested
    auto request_spectra = [&](bool block) {
      while (itr < n_my) {
        const int k = itr % a.nsr, pk = (itr / a.nsr) & 1;
        if (!block && !mbar_test(&r_empty[k], pk ^ 1)) return;
        mbar_wait(&r_empty[k], pk ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&r_full[k], R_STAGE);
          tma_load_3d(sR + k * R_STAGE, &tmS, &r_full[k], 0, blockIdx.x + itr * gridDim.x, 0);
        }
        __syncwarp();
        ++itr;
        if (block) return;
      }
    };
    for (int it = 0; it < n_my; ++it) {
      request_spectra(false);           // as many as the ring takes right now
      if (itr <= it) request_spectra(true);  // the spectra of this mode at the latest
      const int m = blockIdx.x + it * gridDim.x, s = it % NSW, ps = (it / NSW) & 1;
      mbar_wait(&w_empty[s], ps ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&w_full[s], MD_WS);
        // ONE box = the mode's contiguous 32 KB block (64 rows of 512 B): a single sequential DRAM burst per mode
        tma_load_2d(sW + s * MD_WS, &tmW, &w_full[s], 0, m * 64);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint64_t sub = (uint64_t)(N * 128 >> 4);  // one 32-float k-subtile of the spectra, in 16-byte units
    for (int it = 0; it < n_my; ++it) {
      const int t = it & 1, pt = (it >> 1) & 1;
      mbar_wait(&s_full[t], pt);
      mbar_wait(&acc_empty[t], pt ^ 1);
      mbar_wait(&a_full[t], pt);
      tc_fence_after();
      const uint64_t dS_hi = make_smem_desc(smem_u32(sS) + t * S_BUF, 0, 1024);
      const uint64_t dS_lo = make_smem_desc(smem_u32(sS) + t * S_BUF + S_PLANE, 0, 1024);
      const uint32_t acc = T_ACC + t * 64, acc_lo = T_ACCLO + t * 64, Ahi = T_A + t * 128, Alo = Ahi + 64;
      if (elect_one_sync()) {
        // low-order cross terms in their own accumulator: the adder of the tensor core truncates (tc_fwdw.cu)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t o = (uint64_t)(ks >> 2) * sub + (uint64_t)((ks & 3) * 2);
          umma_tf32_ts(acc_lo, Alo + ks * 8, dS_hi + o, idesc, ks != 0);
          umma_tf32_ts(acc_lo, Ahi + ks * 8, dS_lo + o, idesc, 1);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_tf32_ts(acc, Ahi + ks * 8, dS_hi + (uint64_t)(ks >> 2) * sub + (uint64_t)((ks & 3) * 2), idesc, ks != 0);
        umma_commit(&a_empty[t]);
        umma_commit(&s_empty[t]);
        umma_commit(&acc_full[t]);
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    // weight staging: TMEM lane L = (rw, o) holds W_re (rw = 0) or W_im (rw = 1) [i = 0..63][o]
    const int q = warp - 4, L = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int it = 0; it < n_my; ++it) {
      const int s = it % NSW, ps = (it / NSW) & 1, t = it & 1, pt = (it >> 1) & 1;
      mbar_wait(&w_full[s], ps);
      mbar_wait(&a_empty[t], pt ^ 1);
      tc_fence_after();
      // column (rw, o) = float L of a weight row: stage layout [64 i][128 floats] as in HBM
      const uint32_t src = smem_u32(sW) + s * MD_WS + (uint32_t)(L * 4);
      const uint32_t Ahi = T_A + t * 128 + lane_addr, Alo = Ahi + 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32], hv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = lds32(src + (uint32_t)((half * 32 + i) * 512));
        if (half == 1) {  // the stage has been read completely
          __syncwarp();
          if (lane == 0) mbar_arrive(&w_empty[s]);
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) tf32_split2(v[i], v[i + 1], hv[i], hv[i + 1]);
        tmem_st32(Ahi + half * 32, hv);
        tmem_st32(Alo + half * 32, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&a_full[t]);
    }
  } else if (warp >= 8 && warp < 12) {
    pdl_wait();  // stores to O must not pass the completion of the chain's earlier kernels
    // lane L = (rw, o) holds  W_rw . S_re  in columns [0, Nb)  and  W_rw . S_im  in columns [Nb, 2 Nb):
    //   O_re[b][o] = (W_re.S_re)[b] - (W_im.S_im)[b]     -> lane (0,o) col b      minus  lane (1,o) col Nb + b
    //   O_im[b][o] = (W_im.S_re)[b] + (W_re.S_im)[b]     -> lane (1,o) col b      plus   lane (0,o) col Nb + b
    // each lane publishes its S_im half; the partner lane (other rw, same o) picks it up
    // every lane publishes its N accumulator columns; then the 128 threads walk the B live batch entries
    const int q = warp - 8, L = q * 32 + lane, rw = L >> 6, o = L & 63, partner = L ^ 64;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int XLD = N + 1;  // odd row stride: conflict-free both ways
    const float sg = rw == 0 ? -1.f : 1.f;
    const size_t ob = (size_t)2 * NM * 64;  // batch stride of O
    for (int it = 0; it < n_my; ++it) {
      const int m = blockIdx.x + it * gridDim.x, t = it & 1, pt = (it >> 1) & 1;
      mbar_wait(&acc_full[t], pt);
      tc_fence_after();
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16], vl[16];
        tmem_ld16(T_ACC + t * 64 + lane_addr + c0, v);
        tmem_ld16(T_ACCLO + t * 64 + lane_addr + c0, vl);
        tmem_ld_wait();
        if (c0 + 16 >= N) {
          tc_fence_before();
          mbar_arrive(&acc_empty[t]);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) sX[L * XLD + c0 + k] = __uint_as_float(v[k]) + __uint_as_float(vl[k]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float* og = a.O + ((size_t)rw * NM + m) * 64 + o;
      const float* mine = sX + L * XLD;
      const float* theirs = sX + partner * XLD + Nb;
#pragma unroll 4
      for (int b = 0; b < a.B; ++b) og[b * ob] = fmaf(sg, theirs[b], mine[b]);
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the exchange buffer is reused by the next mode
    }
  } else if (warp >= 12) {
    const int st = tid - 384;  // 0..127
    // rows of batch entries >= B are zero in every mode: written once, both buffers, both planes
    for (uint32_t off = (uint32_t)st * 16; off < 2 * S_BUF; off += 128 * 16) sts128(smem_u32(sS) + off, 0.f, 0.f, 0.f, 0.f);
    asm volatile("bar.sync 2, 128;" ::: "memory");
    pdl_wait();  // S is the previous kernel's output
    // work items of this thread: float4 f = st + 128 j of the 2 * B * 16 live ones, f -> (ri, b, i4).  Everything that
    // does not depend on the mode is computed once: in a first version the index arithmetic (two integer divisions per
    // item) ran per mode and made THIS role the critical path of the kernel - every other role waited on it
    // (profiles/r02_modes3d_ncu_full.csv: 2 us per mode).  All loads of a mode are issued before the first is used.
    constexpr int MAXJ = (2 * MD_NMAX * 16 + 127) / 128;  // 8
    const int n_live = 2 * a.B * 16;
    uint32_t roff[MAXJ], soff[MAXJ];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
      const int f = min(st + 128 * j, n_live - 1);
      const int rb = f >> 4, i4 = f & 15, ri = rb / a.B, b = rb - ri * a.B, r = ri * Nb + b;
      roff[j] = (uint32_t)((b * 2 + ri) * 256 + i4 * 16);  // raw stage: rows (b, ri) of 256 B as they lie in S
      soff[j] = (uint32_t)((i4 >> 3) * N * 128) + sw128_off((uint32_t)r, (uint32_t)(i4 & 7));
    }
    const int nj = (n_live - st + 127) / 128;  // live items of this thread (<= MAXJ)
    for (int it = 0; it < n_my; ++it) {
      const int t = it & 1, pt = (it >> 1) & 1, k = it % a.nsr, pk = (it / a.nsr) & 1;
      mbar_wait(&r_full[k], pk);
      const uint32_t raw = smem_u32(sR) + k * R_STAGE;
      uint4 x[MAXJ];
#pragma unroll
      for (int j = 0; j < MAXJ; ++j)
        if (j < nj) x[j] = lds128(raw + roff[j]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&r_empty[k]);
      mbar_wait(&s_empty[t], pt ^ 1);
      const uint32_t base = smem_u32(sS) + t * S_BUF;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j)
        if (j < nj) {
          const float x0 = __uint_as_float(x[j].x), x1 = __uint_as_float(x[j].y), x2 = __uint_as_float(x[j].z),
                      x3 = __uint_as_float(x[j].w);
          const float4 h = make_float4(tf32_hi(x0), tf32_hi(x1), tf32_hi(x2), tf32_hi(x3));
          sts128(base + soff[j], h.x, h.y, h.z, h.w);
          sts128(base + S_PLANE + soff[j], x0 - h.x, x1 - h.y, x2 - h.z, x3 - h.w);
        }
      fence_proxy_async_smem();
      mbar_arrive(&s_full[t]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
bool tc_modes_supported(const Geom& g, int B) { return g.Cp == 64 && B >= 1 && B <= MD_NMAX; }

// packed weights [NM][64 i][2][64 o] viewed as a matrix of NM*64 rows x 128 floats; box = one mode
int tc_make_modes_map(CUtensorMap* m, const float* Wpk, int NM) {
  uint64_t dims[2] = {128, (uint64_t)NM * 64};
  uint64_t strides[1] = {128 * 4};
  uint32_t box[2] = {128, 64};
  return encode_tensor_map(m, Wpk, 2, dims, strides, box, 0);
}

// S [B][2][NM][64] viewed as (i, mode, row = (b, ri)); box = every row of one mode
static int make_spectra_map(CUtensorMap* m, const float* S, int B, int NM) {
  uint64_t dims[3] = {64, (uint64_t)NM, (uint64_t)2 * B};
  uint64_t strides[2] = {64 * 4, (uint64_t)NM * 64 * 4};
  uint32_t box[3] = {64, 1, (uint32_t)(2 * B)};
  return encode_tensor_map(m, S, 3, dims, strides, box, 0);
}

int launch_modes_tc(const CUtensorMap& tmW, const float* S, float* O, int B, int NM, cudaStream_t st) {
  ModesArgs a{};
  a.S = S, a.O = O, a.B = B, a.Nb = B <= 16 ? 16 : 32, a.NM = NM;
  // weight ring + 2 spectra buffers x (hi | lo) x 2 k-subtiles x N rows x 128 B + the epilogue exchange buffer
  // + the raw-spectra ring: as deep as 16 KB allow (8 stages up to B = 4, 2 at B = 32)
  a.nsr = std::max(2, std::min(MD_NSR, 16384 / (2 * B * 256)));
  const int rest = 2 * 2 * 2 * (2 * a.Nb) * 128 + ((128 * (2 * a.Nb + 1) * 4 + 127) & ~127) + a.nsr * 2 * B * 256 + 1024;
  a.nsw = std::min(MD_NSW, (226 * 1024 - rest) / MD_WS);
  const int smem = a.nsw * MD_WS + rest;
  CUtensorMap tmS;
  B2_TRY(make_spectra_map(&tmS, S, B, NM));
  B2_CUDA(cudaFuncSetAttribute(tc_modes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  B2_CUDA(launch_kernel(tc_modes_kernel, dim3(std::min(148, NM)), dim3(MD_THREADS), (size_t)smem, st, a, tmW, tmS));
  B2_LAUNCHED("tc_modes_kernel");
  return 0;
}

}  // namespace b200fno
