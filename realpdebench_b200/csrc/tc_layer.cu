// Fused Fourier-layer body on the 5th-generation tensor cores (width 64):
//
//   out[p][o] = act( ( sum_i x[p][i] Wc[o][i]  +  sum_k G[w(p)][k] D[row][k][o] ) * scale[o] + shift[o] )
//               \_____ bypass 1x1 conv ______/   \____ inverse-W DFT of the mixed modes ____/
//
// (fno.py:114-119: x1 + x2 -> BatchNorm (eval) -> GELU).  One CTA per SM, persistent over the tiles
// (row, j) of its fixed W-tile index j; a tile is PT <= 128 consecutive points of one (b,t,h) row.
//
//   warp 0     TMA producer: x tile (2 boxes of 32 channels) into a 3-stage ring
//   warp 3     TMA producer: D[row] (hi|lo planes) into a 2-stage ring
//   warp 1     MMA issuer: tcgen05.mma kind::tf32, A from TMEM (.ts form), B from shared memory
//   warp 2     TMEM allocation (512 columns)
//   warps 4-7  split: x tile smem -> registers -> TMEM as the A operand, hi = raw fp32 (the MMA truncates
//              to tf32, measured in tests/test_gpu_tc_primitives.py) and lo = x - trunc(x)      [3xTF32]
//   warps 8-15 epilogue: TMEM -> registers -> affine + GELU -> swizzled staging tile -> TMA store
//              (warp w: TMEM lane quarter w%4, channel half (w-8)/4)
//
// TMEM columns: [0,128) two accumulators, [128,384) two (x_hi | x_lo) A buffers, [384,512) G_hi | G_lo
// (the inverse-W table rows of this CTA's W tile, loaded once).  Shared memory: 3 x 32 KB x ring,
// 2 x 32 KB output staging, 32 KB conv weights (hi|lo, K-major), 2 x 16 KB D stages (MN-major).
// fp32 parity: a*b = a_hi*b_hi + a_lo*b_hi + a_hi*b_lo with fp32 accumulation in TMEM.
#include "common.cuh"
#include "tc_common.cuh"

namespace b200fno {
using namespace tc;

// 8 control / split warps + the epilogue warps, in groups of 8 (256 threads: TMEM lane quarter x channel half).
// The epilogue of one tile is a latency chain (accumulator ready -> tcgen05.ld -> barrier -> affine + GELU -> staging ->
// proxy fence -> barrier -> TMA store) that a single group runs one tile at a time; the layer kernel was bound by it
// (ncu: producer, split and MMA warps asleep on their barriers, DRAM 53-64 %).  Layer mode therefore runs TWO
// epilogue groups in ping-pong: group e drains accumulator buffer e (even / odd tiles) into its own staging buffer,
// so two tiles' chains overlap.  The lift keeps one group (its split warps need the registers).
constexpr int EPI_THREADS = 256;
__host__ __device__ constexpr int tcl_groups(int mode) { return mode == 0 ? 2 : 1; }
// The lift is bound by its split warps (the feature gather): it runs a second group of four (warps 16-19) in
// ping-pong with warps 4-7, one group per A-operand buffer (even / odd tiles).
__host__ __device__ constexpr int tcl_split_groups(int mode) { return mode == 0 ? 1 : 2; }
__host__ __device__ constexpr int tcl_threads(int mode) {
  return 256 + EPI_THREADS * tcl_groups(mode) + 128 * (tcl_split_groups(mode) - 1);
}
constexpr int NSX_MAX = 4;
constexpr int XS_MAX = 32768;     // x stage: 2 sub-tiles (32 ch) x PT <= 128 rows x 128 B = PT*256 bytes
constexpr int TCL_BUDGET = 225 * 1024;  // dynamic shared memory the ring / staging layout may use
constexpr int W_BYTES = 32768;    // conv weights hi | lo, each 2 sub-tiles x 64 rows x 128 B
constexpr int DS_BYTES = 16384;   // minimum D stage (also the lift's table area); a stage is 2 N-blocks x 2*K2p k-rows x 128 B
__host__ __device__ constexpr int tcl_ds_bytes(int K2p) { return 512 * K2p > DS_BYTES ? 512 * K2p : DS_BYTES; }
constexpr int TCL_SMEM = TCL_BUDGET + 1024;

enum { MODE_LAYER = 0, MODE_LIFT = 1 };

struct TcLayerArgs {
  const float* Gt;  // [Wp][K2p] inverse-W table (scaled), fp32
  const float *scale, *shift;
  int rows, row0, Wp, PT, NTW, G, K2p, gelu, nsx;  // rows [row0, row0 + rows) of the activation are processed
  int bf16;  // bf16 compute mode: conv / fc0 operands rounded to bf16, one MMA pass (the inverse-W term stays 3xTF32)
  // MODE_LIFT only (fno.py:106-111): A tile = [input features | grid coordinates | 1] built from x
  const float* x;
  const int* in_off;
  const float *gt, *gh, *gw;
  int Tv, H, W, Tp, Hp, c_in, Fin, ng, nkl;  // nkl = K steps (Klp / 8)
  long long x_sB, x_sT;
  // input tile staged by TMA (tmX = map over x as [B][T][H][W*c_in]): nb boxes of IB floats x NF frames
  int in_tma, in_nb, in_IB, in_NF, in_box_floats;
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }



template <int MODE, int NKL>  // NKL: K steps of the lift GEMM (0 in layer mode)
__global__ void __launch_bounds__(tcl_threads(MODE), 1)
    tc_layer_kernel(TcLayerArgs a, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmD) {
  constexpr int TCL_THREADS = tcl_threads(MODE), NGRP = tcl_groups(MODE), NSG = tcl_split_groups(MODE);
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  // pointer arithmetic on the __shared__ array (no integer round trip): every derived pointer keeps its address
  // space, so plain C++ accesses below compile to LDS/STS instead of generic LD.E/ST.E
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sX = smem;
  // stages are sized to the tile (PT points), so a 104-point tile gets a 4-deep x ring where a 128-point one gets 3
  const int SUB = a.PT * 128, XS_BYTES = 2 * SUB, OS_BYTES = 2 * SUB, NSX = a.nsx;
  uint8_t* sOut = sX + NSX * XS_BYTES;
  uint8_t* sW = sOut + 2 * OS_BYTES;
  uint8_t* sD = sW + W_BYTES;
  __shared__ uint64_t x_full[NSX_MAX], x_empty[NSX_MAX], d_full[2], d_empty[2], a_full[2], a_empty[2], acc_full[2],
      acc_empty[2], w_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_scale[64], s_shift[64];
  __shared__ int s_inoff[64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = blockIdx.x % a.NTW, g = blockIdx.x / a.NTW;
  const int n_my = g < a.rows ? (a.rows - g + a.G - 1) / a.G : 0;
  const int PT = a.PT, K2p = a.K2p, DSB = tcl_ds_bytes(a.K2p);  // D stage bytes

  if (tid == 0) {
    for (int i = 0; i < NSX_MAX; ++i) mbar_init(&x_full[i], 1), mbar_init(&x_empty[i], 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1), mbar_init(&d_empty[i], 1);
      mbar_init(&a_full[i], 128), mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], EPI_THREADS);
    }
    mbar_init(&w_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, 512);
  if (tid < 64) s_scale[tid] = a.scale ? a.scale[tid] : 1.f, s_shift[tid] = a.shift ? a.shift[tid] : 0.f;
  if (MODE == MODE_LIFT && tid < 64) s_inoff[tid] = tid < a.Fin ? a.in_off[tid] : 0;
  int* s_fchan = reinterpret_cast<int*>(sD);  // [64] channel of feature f
  int* s_ffoff = s_fchan + 64;                // [64] frame offset of feature f inside a staged box
  int* s_pbase = s_ffoff + 64;                // [c_in <= 8][128] offset of element (point p, channel c)
  if (MODE == MODE_LIFT && a.in_tma) {
    if (tid < 64) {  // entries past the last input feature point at element 0 (the packed W0K column is zero there)
      const int fr = tid / a.c_in;
      s_fchan[tid] = tid < a.Fin ? tid - fr * a.c_in : 0;
      s_ffoff[tid] = tid < a.Fin ? (a.x_sT ? 0 : fr) * a.in_IB : 0;
    }
    for (int i = tid; i < a.c_in * 128; i += TCL_THREADS) {
      const int c = i >> 7, pp = i & 127, e = min(pp, PT - 1) * a.c_in + c, bx = e / a.in_IB;
      s_pbase[i] = bx * a.in_box_floats + (e - bx * a.in_IB);
    }
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmX), prefetch_tensormap(&tmOut), prefetch_tensormap(&tmW), prefetch_tensormap(&tmD);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t T_ACC = tmem, T_A = tmem + 128, T_GHI = tmem + 384, T_GLO = tmem + 448;
  pdl_wait();  // everything below may touch the previous kernel's output (PDL, common.cuh)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&w_full, W_BYTES);
      for (int hl = 0; hl < 2; ++hl)
        for (int s = 0; s < 2; ++s) tma_load_2d(sW + hl * 16384 + s * 8192, &tmW, &w_full, 32 * s, 64 * hl);
    }
    __syncwarp();
    for (int it = 0; MODE == MODE_LIFT && a.in_tma && it < n_my; ++it) {
      const int row = a.row0 + g + it * a.G;
      const int h = row % a.Hp, tt = (row / a.Hp) % a.Tp, b = row / (a.Hp * a.Tp);
      const int sx = it % NSX, px = (it / NSX) & 1;
      mbar_wait(&x_empty[sx], px ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&x_full[sx], (uint32_t)(a.in_nb * a.in_NF * a.in_IB * 4));
        for (int bx = 0; bx < a.in_nb; ++bx)  // OOB coordinates (w >= W, h >= H, t >= T) are zero-filled
          tma_load_4d(sX + sx * XS_BYTES + bx * a.in_box_floats * 4, &tmX, &x_full[sx],
                      PT * j * a.c_in + bx * a.in_IB, h, a.x_sT ? tt : 0, b);
      }
      __syncwarp();
    }
    for (int it = 0; MODE == MODE_LAYER && it < n_my; ++it) {
      const int row = a.row0 + g + it * a.G;
      const int sx = it % NSX, px = (it / NSX) & 1;
      mbar_wait(&x_empty[sx], px ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&x_full[sx], (uint32_t)PT * 256u);
        tma_load_3d(sX + sx * XS_BYTES, &tmX, &x_full[sx], 0, PT * j, row);
        tma_load_3d(sX + sx * XS_BYTES + SUB, &tmX, &x_full[sx], 32, PT * j, row);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ D producer (own warp: the x ring
    // must be able to run its full depth ahead, not be tied to the 2-deep D ring)
    for (int it = 0; MODE == MODE_LAYER && it < n_my; ++it) {
      const int row = a.row0 + g + it * a.G;
      const int sd = it & 1, pd = (it >> 1) & 1;
      mbar_wait(&d_empty[sd], pd ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&d_full[sd], (uint32_t)(2 * 2 * K2p * 128));
        tma_load_2d(sD + sd * DSB, &tmD, &d_full[sd], 0, row * 2 * K2p);
        tma_load_2d(sD + sd * DSB + 2 * K2p * 128, &tmD, &d_full[sd], 32, row * 2 * K2p);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the loop (uniform control flow and descriptors); one elected lane issues.
    const uint32_t idesc_w = make_idesc_tf32(128, 64, 0, 0), idesc_d = make_idesc_tf32(128, 64, 0, 1);
    const uint64_t dW_hi = make_smem_desc(smem_u32(sW), 0, 1024), dW_lo = make_smem_desc(smem_u32(sW) + 16384, 0, 1024);
    const int nk2 = K2p / 8;
    mbar_wait(&w_full, 0);
    for (int it = 0; it < n_my; ++it) {
      const int t = it & 1, pt = (it >> 1) & 1;
      mbar_wait(&a_full[t], pt);
      if (MODE == MODE_LAYER) mbar_wait(&d_full[t], pt);
      mbar_wait(&acc_empty[t], pt ^ 1);
      tc_fence_after();
      const uint32_t acc = T_ACC + t * 64, Ahi = T_A + t * 128, Alo = Ahi + 64;
      // descriptor "lo word" offsets are in 16-byte units: k-step ks of a K-major operand sits at
      // sub-tile (ks/4) * 8192 B + (ks%4) * 32 B; of the MN-major D operand at ks * 8 rows * 128 B
      const uint64_t dD_hi = make_smem_desc(smem_u32(sD) + t * DSB, 2 * K2p * 128, 512, LAYOUT_SW128_BASE32B);
      const uint64_t dD_lo = dD_hi + (uint64_t)(K2p * 128 >> 4);
      if (elect_one_sync()) {
        if (MODE == MODE_LIFT) {
#pragma unroll
          for (int ks = 0; ks < NKL; ++ks) {
              const uint64_t o = (uint64_t)((ks >> 2) * 512 + (ks & 3) * 2);
              if (!a.bf16) {
                umma_tf32_ts(acc, Alo + ks * 8, dW_hi + o, idesc_w, ks > 0);
                umma_tf32_ts(acc, Ahi + ks * 8, dW_lo + o, idesc_w, 1);
              }
              umma_tf32_ts(acc, Ahi + ks * 8, dW_hi + o, idesc_w, a.bf16 ? ks > 0 : 1);
            }
        } else {
          // small (lo) terms first, then the hi*hi terms; bf16 mode: the bypass conv is one pass on rounded operands
          if (!a.bf16) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_tf32_ts(acc, Alo + ks * 8, dW_hi + (uint64_t)((ks >> 2) * 512 + (ks & 3) * 2), idesc_w, ks > 0);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_tf32_ts(acc, Ahi + ks * 8, dW_lo + (uint64_t)((ks >> 2) * 512 + (ks & 3) * 2), idesc_w, 1);
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            if (ks < nk2) umma_tf32_ts(acc, T_GLO + ks * 8, dD_hi + (uint64_t)(ks * 64), idesc_d, (a.bf16 && ks == 0) ? 0 : 1);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            if (ks < nk2) umma_tf32_ts(acc, T_GHI + ks * 8, dD_lo + (uint64_t)(ks * 64), idesc_d, 1);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(acc, Ahi + ks * 8, dW_hi + (uint64_t)((ks >> 2) * 512 + (ks & 3) * 2), idesc_w, 1);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            if (ks < nk2) umma_tf32_ts(acc, T_GHI + ks * 8, dD_hi + (uint64_t)(ks * 64), idesc_d, 1);
          umma_commit(&d_empty[t]);
        }
        umma_commit(&a_empty[t]);
        umma_commit(&acc_full[t]);
      }
      __syncwarp();
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 8 + 8 * NGRP) {
    // ------------------------------------------------------------------ split warps (A operand producers)
    const int q = warp & 3, p = q * 32 + lane;  // TMEM lane quarter = warp % 4
    const int sgrp = warp >= 8 ? 1 : 0;          // split group (lift: warps 16-19 take the odd tiles)
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    if (MODE == MODE_LAYER) {  // inverse-W table rows of this W tile -> TMEM, once
      const int w = PT * j + p;
      const bool valid = p < PT && w < a.Wp;
      for (int k0 = 0; k0 < K2p; k0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float x = valid ? __ldg(a.Gt + (size_t)w * K2p + k0 + i) : 0.f;
          const float xh = tf32_hi(x);
          hi[i] = __float_as_uint(xh);
          lo[i] = __float_as_uint(x - xh);
        }
        tmem_st8(T_GHI + lane_addr + k0, hi);
        tmem_st8(T_GLO + lane_addr + k0, lo);
      }
    }
    if (MODE == MODE_LIFT && n_my > 0) {
      // A row of point p: columns [0, Fin) = input features gathered from x, the last four columns of the
      // last K step = (grid t, grid h, grid w, 1 for the bias); W0K is packed to match.  The features of
      // tile it+1 are fetched before tile it is converted, so their latency is off the critical path.
      constexpr int NIN = NKL * 8 - 4;  // columns available to input features
      uint32_t r[NIN > 0 ? NIN : 1];
      float ex[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int f = 0; f < NIN; ++f) r[f] = 0u;
      auto gather = [&](int it) {
        const int row = a.row0 + g + it * a.G;
        const int h = row % a.Hp, tt = (row / a.Hp) % a.Tp, b = row / (a.Hp * a.Tp);
        const int w = PT * j + p;
        const bool valid = p < PT && w < a.W && h < a.H && tt < a.Tv;
        const float vf = valid ? 1.f : 0.f;
        ex[0] = a.gt ? vf * __ldg(a.gt + min(tt, a.Tv - 1)) : 0.f;
        ex[1] = vf * __ldg(a.gh + min(h, a.H - 1));
        ex[2] = vf * __ldg(a.gw + min(w, a.W - 1));
        ex[3] = vf;
        if (a.in_tma) {  // TMA-staged tile; out-of-range points / rows were zero-filled by the copy engine
          const int sx = it % NSX, px = (it / NSX) & 1;
          mbar_wait(&x_full[sx], px);
          const float* tile = reinterpret_cast<const float*>(sX + sx * XS_BYTES);
#pragma unroll
          for (int f = 0; f < NIN; ++f)
            r[f] = __float_as_uint(tile[s_pbase[s_fchan[f] * 128 + p] + s_ffoff[f]]);  // f >= Fin: table entry 0, W0K column 0
          __syncwarp();
          if (lane == 0) mbar_arrive(&x_empty[sx]);
          return;
        }
        const float* xp = a.x + (size_t)b * a.x_sB + (size_t)min(tt, a.Tv - 1) * a.x_sT +
                          ((size_t)min(h, a.H - 1) * a.W + min(w, a.W - 1)) * a.c_in;
#pragma unroll
        for (int f = 0; f < NIN; ++f)
          if (f < a.Fin) r[f] = valid ? __float_as_uint(__ldg(xp + s_inoff[f])) : 0u;
      };
      if (sgrp < n_my) gather(sgrp);
      for (int it = sgrp; it < n_my; it += NSG) {
        const int t = it & 1, pt = (it >> 1) & 1;
        mbar_wait(&a_empty[t], pt ^ 1);
        tc_fence_after();
        const uint32_t Ahi = T_A + t * 128 + lane_addr, Alo = Ahi + 64;
#pragma unroll
        for (int k0 = 0; k0 < NKL * 8; k0 += 8) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int f = k0 + i;
            const float x = f < NIN ? __uint_as_float(r[f < NIN ? f : 0]) : ex[f >= NIN ? f - NIN : 0];
            const float xh = a.bf16 ? bf16_rn(x) : tf32_hi(x);
            hi[i] = __float_as_uint(xh);
            lo[i] = __float_as_uint(x - xh);
          }
          tmem_st8(Ahi + k0, hi);
          if (!a.bf16) tmem_st8(Alo + k0, lo);
        }
        if (it + NSG < n_my) gather(it + NSG);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&a_full[t]);
      }
    }
    for (int it = 0; MODE == MODE_LAYER && it < n_my; ++it) {
      const int sx = it % NSX, px = (it / NSX) & 1, t = it & 1, pt = (it >> 1) & 1;
      mbar_wait(&x_full[sx], px);
      mbar_wait(&a_empty[t], pt ^ 1);
      tc_fence_after();
      const uint32_t Ahi = T_A + t * 128 + lane_addr, Alo = Ahi + 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        const uint32_t base = smem_u32(sX) + sx * XS_BYTES + half * SUB;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 u = lds128(base + sw128_off(p, c));
          v[4 * c] = u.x, v[4 * c + 1] = u.y, v[4 * c + 2] = u.z, v[4 * c + 3] = u.w;
        }
        if (p >= PT) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (half == 1) {  // every shared-memory read of this stage has been consumed
          __syncwarp();
          if (lane == 0) mbar_arrive(&x_empty[sx]);
        }
        if (a.bf16) {  // operand of the bypass conv rounded to bf16; no low-order plane
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = bf16_rn_bits(v[i]);
          tmem_st32(Ahi + half * 32, v);
          continue;
        }
#pragma unroll
        for (int q16 = 0; q16 < 2; ++q16) {  // 16 values at a time: 768 threads leave 80 registers each
          uint32_t hv[16], lv[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            lv[i] = v[q16 * 16 + i], lv[i + 1] = v[q16 * 16 + i + 1];
            tf32_split2(lv[i], lv[i + 1], hv[i], hv[i + 1]);
          }
          tmem_st16(Ahi + half * 32 + q16 * 16, hv);
          tmem_st16(Alo + half * 32 + q16 * 16, lv);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&a_full[t]);
    }
  } else if (warp >= 8 && warp < 8 + 8 * NGRP) {
    // ------------------------------------------------------------------ epilogue warps
    const int ew = warp - 8, grp = ew >> 3, w8 = ew & 7;  // epilogue group (0 when there is only one), warp in group
    const int q = w8 & 3, half = w8 >> 2, p = q * 32 + lane, etid = tid - 256 - grp * EPI_THREADS;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    for (int it = grp; it < n_my; it += NGRP) {
      // two groups: group e owns accumulator buffer e and staging buffer e; one group: both alternate per tile
      const int row = a.row0 + g + it * a.G, t = it & 1, pt = (it >> 1) & 1, buf = it & 1;
      mbar_wait(&acc_full[t], pt);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(T_ACC + t * 64 + lane_addr + half * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&acc_empty[t]);
      if (etid == 0) {  // the store that last used staging[buf] (issued by this very thread) has read it
        if (NGRP == 2) tma_store_wait_read<0>();
        else tma_store_wait_read<1>();
      }
      named_bar_sync(1 + grp, EPI_THREADS);
      const uint32_t stage = smem_u32(sOut) + buf * OS_BYTES + half * SUB;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        // per-channel affine from shared memory (broadcast LDS.128): keeps the epilogue under the 85-register cap
        const float4 sc = *reinterpret_cast<const float4*>(&s_scale[half * 32 + 4 * c]);
        const float4 sh = *reinterpret_cast<const float4*>(&s_shift[half * 32 + 4 * c]);
        float y0, y1, y2, y3;
        f2_unpack(f2_fma(f2_pack(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1])), f2_pack(sc.x, sc.y),
                         f2_pack(sh.x, sh.y)), y0, y1);
        f2_unpack(f2_fma(f2_pack(__uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3])), f2_pack(sc.z, sc.w),
                         f2_pack(sh.z, sh.w)), y2, y3);
        if (a.gelu) gelu_erf_fast2(y0, y1), gelu_erf_fast2(y2, y3);
        if (p < PT) sts128(stage + sw128_off(p, c), y0, y1, y2, y3);  // rows >= PT lie outside the tile-sized buffer
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + grp, EPI_THREADS);
      if (etid == 0) {
        const uint8_t* st = sOut + buf * OS_BYTES;
        tma_store_3d(&tmOut, st, 0, PT * j, row);
        tma_store_3d(&tmOut, st + SUB, 32, PT * j, row);
        tma_store_commit();
      }
    }
    if (etid == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
bool tc_layer_supported(const Geom& g) {
  return g.Cp == 64 && g.K2p % 8 == 0 && g.K2p <= 64 && ceil_div(g.Wp, 128) <= 148;  // G rows: 2 x 64 TMEM columns
}

// x-ring depth that fits next to 2 staging buffers, the weights and the D ring
int tc_layer_nsx(int PT, int K2p) {
  const int stage = PT * 256;
  return std::max(2, std::min(NSX_MAX, (TCL_BUDGET - 2 * stage - W_BYTES - 2 * tcl_ds_bytes(K2p)) / stage));
}

int tc_layer_tile(const Geom& g, int* PT, int* NTW) {
  *NTW = ceil_div(g.Wp, 128);
  *PT = round_up(ceil_div(g.Wp, *NTW), 8);
  return 0;
}

// tensor maps over channels-last activations [rows][Wp][64]: box = 32 channels x PT points
int tc_make_act_map(CUtensorMap* m, const float* act, long long rows, const Geom& g) {
  int PT, NTW;
  tc_layer_tile(g, &PT, &NTW);
  uint64_t dims[3] = {64, (uint64_t)g.Wp, (uint64_t)rows};
  uint64_t strides[2] = {64 * 4, (uint64_t)g.Wp * 64 * 4};
  uint32_t box[3] = {32, (uint32_t)PT, 1};
  return encode_tensor_map(m, act, 3, dims, strides, box, 1);
}
// conv weights hi|lo: [2*64 rows (hl, o)][64 i], K-major boxes of 32 i x 64 o
int tc_make_w_map(CUtensorMap* m, const float* w_hl) {
  uint64_t dims[2] = {64, 128};
  uint64_t strides[1] = {64 * 4};
  uint32_t box[2] = {32, 64};
  return encode_tensor_map(m, w_hl, 2, dims, strides, box, 1);
}
// D planes: [rows * 2 * K2p k-rows][64 o], MN-major boxes of 32 o x 2*K2p k-rows, 32-byte swizzle atoms
int tc_make_d_map(CUtensorMap* m, const float* d, long long rows, const Geom& g) {
  uint64_t dims[2] = {64, (uint64_t)rows * 2 * g.K2p};
  uint64_t strides[1] = {64 * 4};
  uint32_t box[2] = {32, (uint32_t)(2 * g.K2p)};
  return encode_tensor_map(m, d, 2, dims, strides, box, 2);
}

int launch_layer_tc(const CUtensorMap& tmX, const CUtensorMap& tmOut, const CUtensorMap& tmW, const CUtensorMap& tmD,
                    const float* Gt, const float* scale, const float* shift, long long rows, const Geom& g, int gelu,
                    cudaStream_t st, long long row0, int bf16) {
  TcLayerArgs a{};
  a.bf16 = bf16;
  tc_layer_tile(g, &a.PT, &a.NTW);
  a.Gt = Gt, a.scale = scale, a.shift = shift;
  a.rows = (int)rows, a.row0 = (int)row0, a.Wp = g.Wp, a.K2p = g.K2p, a.gelu = gelu;
  a.G = std::max(1, std::min(148 / a.NTW, (int)rows));
  a.nsx = tc_layer_nsx(a.PT, a.K2p);
  B2_CUDA(cudaFuncSetAttribute(tc_layer_kernel<MODE_LAYER, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCL_SMEM));
  B2_CUDA(launch_kernel(tc_layer_kernel<MODE_LAYER, 0>, dim3(a.NTW * a.G), dim3(tcl_threads(MODE_LAYER)), TCL_SMEM, st, a,
                        tmX, tmOut, tmW, tmD));
  B2_LAUNCHED("tc_layer_kernel");
  return 0;
}

int tc_lift_nkl(int Fin);

// Lift on tensor cores: act0 = [x | grid | 1] * W0K^T, zero in the pad region.  W0K: [2 (hi|lo)][64 ch][64 k].
int launch_lift_tc(const LiftArgs& la, const CUtensorMap& tmOut, const CUtensorMap& tmW0, const Geom& g,
                   cudaStream_t st, int b0, int nb, int bf16) {
  // samples [b0, b0 + nb) of the batch la.B (nb < 0: all of them)
  TcLayerArgs a{};
  a.bf16 = bf16;
  tc_layer_tile(g, &a.PT, &a.NTW);
  if (nb < 0) b0 = 0, nb = la.B;
  a.rows = nb * g.Tp * g.Hp, a.row0 = b0 * g.Tp * g.Hp, a.Wp = g.Wp, a.K2p = g.K2p, a.gelu = 0;
  a.G = std::max(1, std::min(148 / a.NTW, a.rows));
  a.nsx = tc_layer_nsx(a.PT, 0);  // the lift has no D ring (that area holds its gather tables)
  a.x = la.x, a.in_off = la.in_off, a.gt = la.gt, a.gh = la.gh, a.gw = la.gw;
  a.Tv = la.T, a.H = la.H, a.W = la.W, a.Tp = g.Tp, a.Hp = g.Hp, a.c_in = la.c_in, a.Fin = la.Fin, a.ng = la.ng;
  a.nkl = tc_lift_nkl(la.Fin);
  a.x_sB = la.x_sB, a.x_sT = la.x_sT;
  // Stage the input tile with TMA when its geometry allows (x viewed as [B][T][H][W*c_in] fp32).
  CUtensorMap tmIn = tmOut;
  {
    const int seg = a.PT * la.c_in;                      // floats of one frame of one tile
    const int nb = ceil_div(seg, 256), IB = seg / nb;    // boxes of IB <= 256 floats
    const int NF = la.x_sT ? 1 : la.Fin / la.c_in;       // frames per box (2-D: all of them)
    const int box_floats = round_up(NF * IB, 32);        // 128-byte aligned box regions
    const bool ok = IB * nb == seg && IB % 4 == 0 && (la.W * la.c_in) % 4 == 0 && la.c_in <= 8 && NF <= 256 &&
                    nb * box_floats * 4 <= 2 * a.PT * 128 && ((uintptr_t)la.x & 15) == 0 &&
                    (la.c_in * 128 + 128) * 4 <= 2 * DS_BYTES;
    if (ok) {
      const int T_frames = la.x_sT ? la.T : NF;
      uint64_t dims[4] = {(uint64_t)la.W * la.c_in, (uint64_t)la.H, (uint64_t)T_frames, (uint64_t)la.B};
      uint64_t strides[3] = {(uint64_t)la.W * la.c_in * 4, (uint64_t)la.H * la.W * la.c_in * 4,
                             (uint64_t)la.x_sB * 4};  // samples need not be contiguous (rollout: slices of the prediction)
      uint32_t box[4] = {(uint32_t)IB, 1, (uint32_t)NF, 1};
      B2_TRY(encode_tensor_map(&tmIn, la.x, 4, dims, strides, box, 0));
      a.in_tma = 1, a.in_nb = nb, a.in_IB = IB, a.in_NF = NF, a.in_box_floats = box_floats;
    }
  }
  const int grid = a.NTW * a.G;
#define B2_LIFT_CASE(N)                                                                                          \
  case N:                                                                                                        \
    B2_CUDA(cudaFuncSetAttribute(tc_layer_kernel<MODE_LIFT, N>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                 TCL_SMEM));                                                                     \
    B2_CUDA(launch_kernel(tc_layer_kernel<MODE_LIFT, N>, dim3(grid), dim3(tcl_threads(MODE_LIFT)), TCL_SMEM, st, a,  \
                          tmIn, tmOut, tmW0, tmW0));                                                             \
    break;
  switch (a.nkl) {
    B2_LIFT_CASE(1)
    B2_LIFT_CASE(2)
    B2_LIFT_CASE(3)
    B2_LIFT_CASE(8)
    default:
      set_error("tensor-core lift: unsupported K steps %d", a.nkl);
      return B200FNO_EINVAL;
  }
#undef B2_LIFT_CASE
  B2_LAUNCHED("tc_lift_kernel");
  return 0;
}

// K steps of the lift GEMM for Fin input features (+ 4 fixed columns: grid t, h, w and the bias), or 0 if the
// tensor-core lift has no instantiation for it
int tc_lift_nkl(int Fin) {
  const int n = ceil_div(Fin + 4, 8);
  if (n <= 3) return n;
  return n <= 8 ? 8 : 0;
}

}  // namespace b200fno
