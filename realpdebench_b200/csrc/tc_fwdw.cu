// Forward W transform of the truncated DFT on the tensor cores (width 64, 2*m3 in {16, 32, 48, 64}):
//
//   A[row][k][c] = sum_w LF[k][w] * x[row][w][c]          (rfftn's last axis, kept modes only; fno.py:48,53)
//
// The contraction runs over the points of a row, so the activation tile has to enter the MMA
// transposed (M = channel, K = point).  Each CTA handles PAIRS of rows: M = 128 = 2 rows x 64 channels.
// A chunk of 64 points x 64 channels x 2 rows is TMA-loaded (no swizzle); thread (row r, channel c)
// reads its column of the tile (conflict-free, one 128 B shared-memory row per warp access), and
// writes it as TMEM lane r*64+c - the transpose costs no extra shared-memory pass.  hi = raw fp32
// (the MMA truncates to tf32), lo = x - trunc(x); B = the DFT table hi|lo, K-major; its 64-point chunk
// rides in the same stage as the x chunk (an L2 hit shared by every CTA), which leaves shared memory for
// a 4-deep ring: 128 KB of activation reads in flight per SM.
//
//   warp 0     TMA: x chunk + table chunk into a 4-stage ring
//   warp 1     MMA issuer (.ts form), 24 MMAs (N = 2*m3) per chunk, accumulating over the row
//   warp 2     TMEM allocation
//   warps 4-7  transpose + split: smem -> TMEM (x_hi | x_lo)
//   warps 8-11 epilogue: TMEM accumulator -> A[row][k][c] (coalesced over c)
#include "common.cuh"
#include "tc_common.cuh"

namespace b200fno {
using namespace tc;

constexpr int FW_THREADS = 384;
constexpr int FW_NS = 4;          // ring stages
constexpr int FW_CH = 64;         // points per chunk
constexpr int FW_XS = 32768;      // x part of a stage: 2 channel halves x 2 rows x 64 points x 128 B
// table part of a stage: [hi|lo][2 sub-tiles of 32 points][K2 rows][128 B] = 512*K2 bytes (16-32 KB)

struct FwdWArgs {
  float* out;  // [rows][K2][64]
  int rows, row0, npairs, nchunk, nsub, K2, K2m, ns, stage_bytes;  // K2 output rows; K2m = K2 rounded up to the MMA N step (16)  // ns ring stages of stage_bytes (x chunk + table chunk)
};

__global__ void __launch_bounds__(FW_THREADS, 1)
    tc_fwdw_kernel(FwdWArgs a, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmF) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* sX = smem;  // FW_NS stages of [x chunk | table chunk]
  __shared__ uint64_t x_full[FW_NS], x_empty[FW_NS], a_full[2], a_empty[2], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K2 = a.K2m, K2out = a.K2, nchunk = a.nchunk, NS = a.ns, FW_STAGE = a.stage_bytes;  // K2: MMA N / table rows
  const int n_my = (int)blockIdx.x < a.npairs ? (a.npairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    // a stage is free once the 4 transpose warps have read x AND the MMAs that read its table chunk are done
    for (int i = 0; i < FW_NS; ++i) mbar_init(&x_full[i], 1), mbar_init(&x_empty[i], 5);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 128), mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, 512);
  if (warp == 0 && lane == 0) prefetch_tensormap(&tmX), prefetch_tensormap(&tmF);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // everything below may touch the previous kernel's output (PDL, common.cuh)
  const uint32_t tmem = tmem_base_s;
  // acc: 2 x 64 cols (K2 <= 64 used); A: 2 x (64 hi | 64 lo); low-order accumulators: 2 x 64 cols
  const uint32_t T_ACC = tmem, T_A = tmem + 128, T_ACCLO = tmem + 384;

  if (warp == 0) {
    int sx = 0, px = 0;  // ring stage and its phase, advanced incrementally (NS is a run-time value: no div / mod)
    for (int ip = 0; ip < n_my; ++ip) {
      const int pair = a.npairs - 1 - (blockIdx.x + ip * gridDim.x);  // reverse order: see launch_fwdw_tc
      for (int ch = 0; ch < nchunk; ++ch, sx = (sx + 1 == NS ? 0 : sx + 1), px ^= (sx == 0)) {
        mbar_wait(&x_empty[sx], px ^ 1);
        if (elect_one_sync()) {
          uint8_t* st = sX + sx * FW_STAGE;
          mbar_arrive_expect_tx(&x_full[sx], (uint32_t)(FW_XS + 4 * K2 * 128));
          tma_load_3d(st, &tmX, &x_full[sx], 0, ch * FW_CH, a.row0 + 2 * pair);
          tma_load_3d(st + 16384, &tmX, &x_full[sx], 32, ch * FW_CH, a.row0 + 2 * pair);
          for (int hl = 0; hl < 2; ++hl)
            for (int s = 0; s < 2; ++s)
              tma_load_2d(st + FW_XS + (hl * 2 + s) * K2 * 128, &tmF, &x_full[sx], 32 * (2 * ch + s), K2 * hl);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_tf32(128, K2, 0, 0);
    const uint64_t sub = (uint64_t)(K2 * 128 >> 4);  // one 32-point sub-tile of the table, 16-byte units
    int cc = 0, sx = 0, px = 0;
    for (int ip = 0; ip < n_my; ++ip) {
      const int ab = ip & 1, pab = (ip >> 1) & 1;
      mbar_wait(&acc_empty[ab], pab ^ 1);
      for (int ch = 0; ch < nchunk; ++ch, ++cc, sx = (sx + 1 == NS ? 0 : sx + 1), px ^= (sx == 0)) {
        const int t = cc & 1, pt = (cc >> 1) & 1;
        mbar_wait(&x_full[sx], px);  // table chunk of this stage has landed
        mbar_wait(&a_full[t], pt);
        tc_fence_after();
        // Two accumulators per row pair: the hi*hi products and the two low-order cross terms.  The adder of the
        // tensor core TRUNCATES when it folds a K step into the fp32 accumulator (test_accumulator_rounding_mode),
        // a bias of ~0.3 ulp of the running sum per MMA; kept apart, the 2 x 8 cross-term MMAs of every chunk work
        // on a sum 2^-11 the size, so only the hi*hi chain (1/3 of the MMAs) pays it.  The epilogue adds the two.
        const uint32_t acc = T_ACC + ab * 64, acc_lo = T_ACCLO + ab * 64, Ahi = T_A + t * 128, Alo = Ahi + 64;
        const uint64_t dF_hi = make_smem_desc(smem_u32(sX) + sx * FW_STAGE + FW_XS, 0, 1024);
        const uint64_t dF_lo = dF_hi + 2 * sub;
        const uint64_t o0 = 0;
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(acc_lo, Alo + ks * 8, dF_hi + o0 + (uint64_t)(ks >> 2) * sub + (uint64_t)((ks & 3) * 2), idesc,
                         (ch | ks) != 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(acc_lo, Ahi + ks * 8, dF_lo + o0 + (uint64_t)(ks >> 2) * sub + (uint64_t)((ks & 3) * 2), idesc, 1);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(acc, Ahi + ks * 8, dF_hi + o0 + (uint64_t)(ks >> 2) * sub + (uint64_t)((ks & 3) * 2), idesc,
                         (ch | ks) != 0);
          umma_commit(&a_empty[t]);
          umma_commit(&x_empty[sx]);
          if (ch == nchunk - 1) umma_commit(&acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // TMEM lane m = r*64 + c: warp q holds row r = q/2, channel half q%2, lane = c%32
    const int q = warp - 4;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t col_base = (uint32_t)((q & 1) * 16384 + (q >> 1) * 8192 + lane * 4);
    int cc = 0, sx = 0, px = 0;
    for (int ip = 0; ip < n_my; ++ip) {
      for (int ch = 0; ch < nchunk; ++ch, ++cc, sx = (sx + 1 == NS ? 0 : sx + 1), px ^= (sx == 0)) {
        const int t = cc & 1, pt = (cc >> 1) & 1;
        mbar_wait(&x_full[sx], px);
        mbar_wait(&a_empty[t], pt ^ 1);
        tc_fence_after();
        const uint32_t src = smem_u32(sX) + sx * FW_STAGE + col_base;
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
    const int q = warp - 8;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int r = q >> 1, c = (q & 1) * 32 + lane;
    for (int ip = 0; ip < n_my; ++ip) {
      const int pair = a.npairs - 1 - (blockIdx.x + ip * gridDim.x), ab = ip & 1, pab = (ip >> 1) & 1;
      mbar_wait(&acc_full[ab], pab);
      tc_fence_after();
      const int row = 2 * pair + r;
      float* o = a.out + (size_t)(a.row0 + row) * K2out * 64 + c;
      for (int c0 = 0; c0 < K2; c0 += 32) {  // 32 accumulator columns (= output rows k) per pass
        uint32_t v[32], vl[32];
        tmem_ld32(T_ACC + ab * 64 + lane_addr + c0, v);
        tmem_ld32(T_ACCLO + ab * 64 + lane_addr + c0, vl);
        tmem_ld_wait();
        if (c0 + 32 >= K2) {  // last pass: the accumulator buffers are free again
          tc_fence_before();
          mbar_arrive(&acc_empty[ab]);
        }
        if (row < a.rows) {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c0 + k < K2out) o[(size_t)(c0 + k) * 64] = __uint_as_float(v[k]) + __uint_as_float(vl[k]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
int tc_fwdw_k2m(const Geom& g) { return round_up(g.K2, 16); }  // MMA N: table rows padded with zeros
bool tc_fwdw_supported(const Geom& g) { return g.Cp == 64 && g.K2 == g.K2p && g.K2 <= 64; }  // two 64-column accumulators
int tc_fwdw_nsub(const Geom& g) { return 2 * ceil_div(g.Wp, FW_CH); }
// table planes [2 (hi|lo)][K2 rows][nsub*32 points], zero padded
size_t tc_fwdw_table_floats(const Geom& g) { return (size_t)2 * tc_fwdw_k2m(g) * tc_fwdw_nsub(g) * 32; }

int tc_make_fwdw_maps(CUtensorMap* tmX, CUtensorMap* tmF, const float* act, const float* table, long long rows,
                      const Geom& g) {
  {
    uint64_t dims[3] = {64, (uint64_t)g.Wp, (uint64_t)rows};
    uint64_t strides[2] = {64 * 4, (uint64_t)g.Wp * 64 * 4};
    uint32_t box[3] = {32, FW_CH, 2};
    B2_TRY(encode_tensor_map(tmX, act, 3, dims, strides, box, 0));
  }
  const int wpad = tc_fwdw_nsub(g) * 32;
  uint64_t dims[2] = {(uint64_t)wpad, (uint64_t)2 * tc_fwdw_k2m(g)};
  uint64_t strides[1] = {(uint64_t)wpad * 4};
  uint32_t box[2] = {32, (uint32_t)tc_fwdw_k2m(g)};
  return encode_tensor_map(tmF, table, 2, dims, strides, box, 1);
}

// Rows are visited from the LAST pair to the first: the kernel that produced this activation (lift / layer)
// wrote it front to back, so its tail is still in the 126 MB L2 when this kernel starts; and this kernel
// leaves the head of the activation in L2 for the layer kernel that reads it next, front to back.
int launch_fwdw_tc(const CUtensorMap& tmX, const CUtensorMap& tmF, float* out, long long rows, const Geom& g,
                   cudaStream_t st, long long row0) {
  // rows [row0, row0 + rows) of the activation (a pair whose second row lies beyond the range is loaded - the tensor
  // map zero-fills past the end of the activation - but only its first row is written)
  FwdWArgs a{};
  a.out = out, a.rows = (int)rows, a.row0 = (int)row0, a.npairs = (int)((rows + 1) / 2);
  a.nchunk = ceil_div(g.Wp, FW_CH), a.nsub = tc_fwdw_nsub(g), a.K2 = g.K2, a.K2m = tc_fwdw_k2m(g);
  a.stage_bytes = FW_XS + round_up(512 * a.K2m, 1024);
  a.ns = std::min(FW_NS, (226 * 1024) / a.stage_bytes);
  const int smem = a.ns * a.stage_bytes + 1024;
  B2_CUDA(cudaFuncSetAttribute(tc_fwdw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  B2_CUDA(launch_kernel(tc_fwdw_kernel, dim3(std::min(148, a.npairs)), dim3(FW_THREADS), (size_t)smem, st, a, tmX, tmF));
  B2_LAUNCHED("tc_fwdw_kernel");
  return 0;
}

}  // namespace b200fno
