// Projection + rollout glue on the tensor cores (width 64):
//
//   crop -> fc1 (64 -> 128) -> GELU -> fc2 (128 -> F) -> unfold -> p*a[c] + b[c] -> prediction slice / next input
//   (fno.py:121-128; eval.py:315-318 as one per-channel affine)
//
// Two chained 3xTF32 GEMMs per tile of PT <= 128 valid points; the hidden activations never leave
// the SM: GEMM-1 accumulates in TMEM, the epilogue warps apply bias + GELU and write the result
// straight back into TMEM as the (hi | lo) A operand of GEMM-2.
//
//   warp 0     TMA: fc1 / fc2 weights once (hi|lo, K-major), x tiles into a 3-stage ring
//   warp 1     MMA issuer (A from TMEM, B from shared memory)
//   warp 2     TMEM allocation
//   warps 4-7  split: x tile -> TMEM (x_hi | x_lo), then 2 of the 8 sixteen-column chunks of epilogue 1
//   warps 8-15 epilogue 1 (bias, GELU, hi/lo -> TMEM; 3 chunks each) and epilogue 2 (affine, staged TMA stores)
//
// TMEM columns: [0,128) x_hi|x_lo, [128,256) accumulator (GEMM-2 reuses it), [256,512) h_hi|h_lo.
#include "common.cuh"
#include "tc_common.cuh"

namespace b200fno {
using namespace tc;

constexpr int TCP_THREADS = 512;
constexpr int TCP_EPI = 256;
constexpr int TCP_NSX = 2;
constexpr int TCP_XS = 32768;
constexpr int TCP_ST = 32768;  // output staging tile (+ its index tables) for the TMA-store epilogue
constexpr int TCP_W1 = 65536;  // [hl][2 k-subtiles][128 rows][128 B]
constexpr int TCP_W2 = 65536;  // [hl][4 k-subtiles][N2 <= 64 rows][128 B]
constexpr int TCP_SMEM = TCP_NSX * TCP_XS + TCP_ST + TCP_W1 + TCP_W2 + 1024;

struct TcProjArgs {
  const float *fc1b, *fc2b;      // [128], [>= Fout]
  const float *aff_a, *aff_b;    // per physical channel or nullptr
  const int *chan, *out_off, *st_off;
  float *out, *state;
  int rows, row0, Tv, H, W, Tp, Hp, PT, NTW, G, Fout, N2, c_out, c_in;  // valid rows [row0, row0 + rows)
  int bf16;  // bf16 compute mode: x, hidden activations and weights rounded to bf16, one MMA pass per GEMM
  long long out_sB, out_sT, st_sB, st_sT;
  // TMA-store epilogue: the output tile is staged as [box][frame][OB floats] and stored with 4-D boxes over
  // out viewed as [B][frames][H][W*c_out] (and the next-input state when c_in == c_out)
  int out_tma, o_nb, o_OB, o_NF, o_box_floats, o_r, ndim3, st_tma;
};


__device__ __forceinline__ void tcp_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <bool OUT_TMA>
__global__ void __launch_bounds__(TCP_THREADS, 1)
    tc_proj_kernel(TcProjArgs a, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                   const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
                   const __grid_constant__ CUtensorMap tmState) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* sX = smem;
  uint8_t* sSt = sX + TCP_NSX * TCP_XS;
  uint8_t* sW1 = sSt + TCP_ST;
  int* s_pbase = reinterpret_cast<int*>(sSt + TCP_ST - 2048) + 128;  // [c_out <= 3][128] offset of (point, channel)
  uint8_t* sW2 = sW1 + TCP_W1;
  __shared__ uint64_t x_full[TCP_NSX], x_empty[TCP_NSX], w_full, xa_full, xa_empty, acc1_full, h_full[4], acc2_full,
      acc_free;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_b1[128], s_b2[64];
  // per output feature, one LDS.128: TMA-store epilogue {frame offset in a box, channel, affine a, affine b + b2*a};
  // scattered-store epilogue {offset in the prediction, offset in the next input, affine a, affine b}
  __shared__ int4 s_fpk[64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = blockIdx.x % a.NTW, g = blockIdx.x / a.NTW;
  const int n_my = g < a.rows ? (a.rows - g + a.G - 1) / a.G : 0;
  const int PT = a.PT, N2 = a.N2;

  if (tid == 0) {
    for (int i = 0; i < TCP_NSX; ++i) mbar_init(&x_full[i], 1), mbar_init(&x_empty[i], 4);
    mbar_init(&w_full, 1);
    mbar_init(&xa_full, 128), mbar_init(&xa_empty, 1);
    mbar_init(&acc1_full, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&h_full[i], 256);  // 2 sixteen-column chunks x 128 lanes per hidden quarter
    mbar_init(&acc2_full, 1), mbar_init(&acc_free, TCP_EPI);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, 512);
  if (tid < 128) s_b1[tid] = a.fc1b[tid];
  if (tid < 64) {
    const bool on = tid < a.Fout;
    const int ch = on ? a.chan[tid] : 0;
    const float b2 = on ? a.fc2b[tid] : 0.f;
    const float fa = (on && a.aff_a) ? a.aff_a[ch] : 1.f, fb = (on && a.aff_b) ? a.aff_b[ch] : 0.f;
    s_b2[tid] = b2;
    if (OUT_TMA) {  // fold the fc2 bias into the affine: (acc + b2)*a + b = acc*a + (b2*a + b)
      const int fr = a.ndim3 ? tid % a.o_r : tid / a.c_out;
      s_fpk[tid] = make_int4(on ? fr * a.o_OB : 0, on ? ch : 0, __float_as_int(fa), __float_as_int(fmaf(b2, fa, fb)));
    } else {
      s_fpk[tid] = make_int4(on ? a.out_off[tid] : 0, on ? a.st_off[tid] : 0, __float_as_int(fa), __float_as_int(fb));
    }
  }
  if (OUT_TMA)
    for (int i = tid; i < a.c_out * 128; i += TCP_THREADS) {
      const int c = i >> 7, pp = i & 127, e = min(pp, a.PT - 1) * a.c_out + c, bx = e / a.o_OB;
      s_pbase[i] = bx * a.o_box_floats + (e - bx * a.o_OB);
    }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmX), prefetch_tensormap(&tmW1), prefetch_tensormap(&tmW2);
    if (OUT_TMA) prefetch_tensormap(&tmOut), prefetch_tensormap(&tmState);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t T_X = tmem, T_ACC = tmem + 128, T_H = tmem + 256;
  pdl_wait();  // everything below may touch the previous kernel's output (PDL, common.cuh)

  // epilogue 1 on 16 hidden columns of this thread's TMEM lane: GELU(acc + b1) -> (hi | lo) A operand of GEMM-2.
  // The tcgen05.ld of all of a warp's chunks are issued up front (one wait), so only the first pays the load latency.
  auto epi1_load = [&](uint32_t lane_addr, int col0, uint32_t (&v)[16]) { tmem_ld16(T_ACC + lane_addr + col0, v); };
  auto epi1_store = [&](uint32_t lane_addr, int col0, uint32_t (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      float y0 = __uint_as_float(v[i]) + s_b1[col0 + i], y1 = __uint_as_float(v[i + 1]) + s_b1[col0 + i + 1];
      gelu_erf_fast2(y0, y1);
      v[i] = __float_as_uint(y0), v[i + 1] = __float_as_uint(y1);
    }
    if (a.bf16) {  // fc1 output and GELU are bf16 tensors under autocast: the fc2 operand is the rounded value
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = bf16_rn_bits(v[i]);
      tmem_st16(T_H + lane_addr + col0, v);
      return;
    }
    uint32_t hv[16];
#pragma unroll
    for (int i = 0; i < 16; i += 2) tf32_split2(v[i], v[i + 1], hv[i], hv[i + 1]);
    tmem_st16(T_H + lane_addr + col0, hv);
    tmem_st16(T_H + 128 + lane_addr + col0, v);
  };
  auto epi1_publish = [&](int chunk) {  // call after tmem_st_wait(): the chunk's hidden quarter may feed GEMM-2
    mbar_arrive(&h_full[chunk >> 1]);
  };

  auto padded_row = [&](int row) {  // valid row (b, t, h) -> row of the padded activation grid
    const int h = row % a.H, t = (row / a.H) % a.Tv, b = row / (a.H * a.Tv);
    return (b * a.Tp + t) * a.Hp + h;
  };

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&w_full, (uint32_t)(TCP_W1 + 2 * 4 * N2 * 128));
      for (int hl = 0; hl < 2; ++hl)
        for (int s = 0; s < 2; ++s) tma_load_2d(sW1 + (hl * 2 + s) * 16384, &tmW1, &w_full, 32 * s, 128 * hl);
      for (int hl = 0; hl < 2; ++hl)
        for (int s = 0; s < 4; ++s) tma_load_2d(sW2 + (hl * 4 + s) * N2 * 128, &tmW2, &w_full, 32 * s, N2 * hl);
    }
    __syncwarp();
    for (int it = 0; it < n_my; ++it) {
      const int prow = padded_row(a.row0 + a.rows - 1 - (g + it * a.G));  // back to front: the tail is still in L2
      const int sx = it % TCP_NSX, px = (it / TCP_NSX) & 1;
      mbar_wait(&x_empty[sx], px ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&x_full[sx], (uint32_t)PT * 256u);
        tma_load_3d(sX + sx * TCP_XS, &tmX, &x_full[sx], 0, PT * j, prow);
        tma_load_3d(sX + sx * TCP_XS + 16384, &tmX, &x_full[sx], 32, PT * j, prow);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // whole warp runs the loop; one elected lane issues (see tc_common.cuh: elect_one_sync)
    const uint32_t idesc1 = make_idesc_tf32(128, 128, 0, 0), idesc2 = make_idesc_tf32(128, N2, 0, 0);
    const uint64_t d1_hi = make_smem_desc(smem_u32(sW1), 0, 1024), d1_lo = make_smem_desc(smem_u32(sW1) + 32768, 0, 1024);
    const uint64_t d2_hi = make_smem_desc(smem_u32(sW2), 0, 1024),
                   d2_lo = make_smem_desc(smem_u32(sW2) + 4 * N2 * 128, 0, 1024);
    const uint64_t sub2 = (uint64_t)(N2 * 128 >> 4);  // one k-subtile of fc2 in 16-byte units
    mbar_wait(&w_full, 0);
    for (int it = 0; it < n_my; ++it) {
      const uint32_t ph = it & 1;
      mbar_wait(&xa_full, ph);
      mbar_wait(&acc_free, ph ^ 1);
      tc_fence_after();
      if (elect_one_sync()) {
        if (!a.bf16) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(T_ACC, T_X + 64 + ks * 8, d1_hi + (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2), idesc1, ks > 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_tf32_ts(T_ACC, T_X + ks * 8, d1_lo + (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2), idesc1, 1);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_tf32_ts(T_ACC, T_X + ks * 8, d1_hi + (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2), idesc1,
                       a.bf16 ? ks > 0 : 1);
        umma_commit(&xa_empty);
        umma_commit(&acc1_full);
      }
      __syncwarp();
      // GEMM-2 is issued per hidden quarter (4 K-steps) as soon as epilogue 1 has produced it, so it overlaps
      // the rest of epilogue 1.  Its accumulator re-uses columns [0,64) of GEMM-1's: quarters 0 and 1 together
      // guarantee that every epilogue-1 read of those columns (chunks 0-3) has completed.
      mbar_wait(&h_full[0], ph);
#pragma unroll
      for (int Q = 0; Q < 4; ++Q) {
        if (Q > 0) mbar_wait(&h_full[Q], ph);
        if (Q == 0) continue;  // quarter 0 is issued together with quarter 1
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int ks = (Q == 1 ? 0 : 4 * Q); ks < 4 * Q + 4; ++ks) {
            const uint64_t o = (uint64_t)(ks >> 2) * sub2 + (uint64_t)((ks & 3) * 2);
            if (!a.bf16) {
              umma_tf32_ts(T_ACC, T_H + 128 + ks * 8, d2_hi + o, idesc2, ks > 0);
              umma_tf32_ts(T_ACC, T_H + ks * 8, d2_lo + o, idesc2, 1);
            }
            umma_tf32_ts(T_ACC, T_H + ks * 8, d2_hi + o, idesc2, a.bf16 ? ks > 0 : 1);
          }
          if (Q == 3) umma_commit(&acc2_full);
        }
        __syncwarp();
      }
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp - 4, p = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // x tile `it` -> TMEM (x_hi | x_lo) as the A operand of GEMM-1
    auto stage_x = [&](int it) {
      const int sx = it % TCP_NSX, px = (it / TCP_NSX) & 1;
      mbar_wait(&x_full[sx], px);
      mbar_wait(&xa_empty, (it & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        const uint32_t base = smem_u32(sX) + sx * TCP_XS + half * 16384;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 u = lds128(base + sw128_off(p, c));
          v[4 * c] = u.x, v[4 * c + 1] = u.y, v[4 * c + 2] = u.z, v[4 * c + 3] = u.w;
        }
        if (p >= PT) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        if (half == 1) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&x_empty[sx]);
        }
        if (a.bf16) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = bf16_rn_bits(v[i]);
          tmem_st32(T_X + lane_addr + half * 32, v);
          continue;
        }
        uint32_t hv[32];
#pragma unroll
        for (int i = 0; i < 32; i += 2) tf32_split2(v[i], v[i + 1], hv[i], hv[i + 1]);
        tmem_st32(T_X + lane_addr + half * 32, hv);
        tmem_st32(T_X + 64 + lane_addr + half * 32, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&xa_full);
    };
    if (n_my > 0) stage_x(0);
    for (int it = 0; it < n_my; ++it) {
      mbar_wait(&acc1_full, it & 1);
      // GEMM-1 of this tile has finished with the x operand (xa_empty is committed together with acc1_full): stage
      // the NEXT tile now, so that GEMM-1(it+1) can start the moment epilogue 2 has drained the accumulator, instead
      // of after this warp's epilogue-1 chunks
      if (it + 1 < n_my) stage_x(it + 1);
      tc_fence_after();
      // this warp's share of epilogue 1: round-robin over the three warps of a lane quarter, chunks 2 and 5
      {
        uint32_t v0[16], v1[16];
        epi1_load(lane_addr, 2 * 16, v0), epi1_load(lane_addr, 5 * 16, v1);
        tmem_ld_wait();
        epi1_store(lane_addr, 2 * 16, v0);
        tmem_st_wait();
        tc_fence_before();
        epi1_publish(2);
        epi1_store(lane_addr, 5 * 16, v1);
        tmem_st_wait();
        tc_fence_before();
        epi1_publish(5);
      }
    }
  } else if (warp >= 8) {
    const int q = warp & 3, hh = (warp - 8) >> 2, p = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // staging offset of (this thread's point, channel c): fixed for the whole kernel, kept in registers
    int pb0 = 0, pb1 = 0, pb2 = 0;
    if (OUT_TMA) {
      const uint32_t a_pbase = smem_u32(s_pbase);  // byte addresses inside the staging tile
      pb0 = (int)smem_u32(sSt) + 4 * (int)lds32(a_pbase + 4 * p);
      if (a.c_out > 1) pb1 = (int)smem_u32(sSt) + 4 * (int)lds32(a_pbase + 4 * (128 + p));
      if (a.c_out > 2) pb2 = (int)smem_u32(sSt) + 4 * (int)lds32(a_pbase + 4 * (256 + p));
    }
    for (int it = 0; it < n_my; ++it) {
      const uint32_t ph = it & 1;
      const int row = a.row0 + a.rows - 1 - (g + it * a.G);
      const int h = row % a.H, t = (row / a.H) % a.Tv, b = row / (a.H * a.Tv);
      const int w = PT * j + p;
      // ---- epilogue 1: hidden = GELU(acc + b1) -> TMEM as the A operand of GEMM-2 (hi | lo)
      mbar_wait(&acc1_full, ph);
      tc_fence_after();
      // sixteen-column chunks in increasing order across the three warps of a lane quarter (hh = 0: 0,3,6;
      // hh = 1: 1,4,7; split warp: 2,5): low hidden quarters complete first and GEMM-2 starts on them early
      {
        uint32_t v0[16], v1[16], v2[16];
        epi1_load(lane_addr, hh * 16, v0), epi1_load(lane_addr, (hh + 3) * 16, v1), epi1_load(lane_addr, (hh + 6) * 16, v2);
        tmem_ld_wait();
        epi1_store(lane_addr, hh * 16, v0);
        tmem_st_wait();
        tc_fence_before();
        epi1_publish(hh);
        epi1_store(lane_addr, (hh + 3) * 16, v1);
        tmem_st_wait();
        tc_fence_before();
        epi1_publish(hh + 3);
        epi1_store(lane_addr, (hh + 6) * 16, v2);
        tmem_st_wait();
        tc_fence_before();
        epi1_publish(hh + 6);
      }
      // ---- epilogue 2: + b2, rollout affine, scatter to the prediction slice and the next model input
      mbar_wait(&acc2_full, ph);
      tc_fence_after();
      uint32_t v[32];
      const bool have = hh * 32 < N2;
      if (have) {
        tmem_ld32(T_ACC + lane_addr + hh * 32, v);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(&acc_free);
      if (OUT_TMA) {
        const int etid = tid - 256;
        if (etid == 0) tma_store_wait_read<0>();  // the previous tile's stores have read the staging buffer
        tcp_named_bar(1, TCP_EPI);
        if (have && p < PT) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int f = hh * 32 + i;
            if (f < a.Fout) {
              const int4 k = s_fpk[f];  // warp-uniform address: one broadcast LDS.128 per feature
              const int base = k.y == 0 ? pb0 : (k.y == 1 ? pb1 : pb2);
              sts32((uint32_t)(base + 4 * k.x), fmaf(__uint_as_float(v[i]), __int_as_float(k.z), __int_as_float(k.w)));
            }
          }
        }
        fence_proxy_async_smem();
        tcp_named_bar(1, TCP_EPI);
        if (etid == 0) {
          const int frame0 = a.ndim3 ? t * a.o_r : 0;
          for (int bx = 0; bx < a.o_nb; ++bx) {
            tma_store_4d(&tmOut, sSt + bx * a.o_box_floats * 4, PT * j * a.c_out + bx * a.o_OB, h, frame0, b);
            if (a.st_tma)
              tma_store_4d(&tmState, sSt + bx * a.o_box_floats * 4, PT * j * a.c_out + bx * a.o_OB, h, frame0, b);
          }
          tma_store_commit();
        }
      } else if (have && p < PT && w < a.W) {
        const size_t po = (size_t)b * a.out_sB + (size_t)t * a.out_sT + ((size_t)h * a.W + w) * a.c_out;
        const size_t ps = (size_t)b * a.st_sB + (size_t)t * a.st_sT + ((size_t)h * a.W + w) * a.c_in;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int f = hh * 32 + i;
          if (f < a.Fout) {
            const int4 k = s_fpk[f];
            const float y = fmaf(__uint_as_float(v[i]) + s_b2[f], __int_as_float(k.z), __int_as_float(k.w));
            a.out[po + k.x] = y;
            if (a.state) a.state[ps + k.y] = y;
          }
        }
      }
    }
    if (OUT_TMA && tid == 256) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
bool tc_proj_supported(const Geom& g, int Fout) { return g.Cp == 64 && Fout <= 64; }
int tc_proj_n2(int Fout) { return round_up(Fout, 16); }
void tc_proj_tile(int W, int* PT, int* NTW) {
  *NTW = ceil_div(W, 128);
  *PT = round_up(ceil_div(W, *NTW), 8);
}
int tc_make_proj_act_map(CUtensorMap* m, const float* act, long long rows, const Geom& g, int W) {
  int PT, NTW;
  tc_proj_tile(W, &PT, &NTW);
  uint64_t dims[3] = {64, (uint64_t)g.Wp, (uint64_t)rows};
  uint64_t strides[2] = {64 * 4, (uint64_t)g.Wp * 64 * 4};
  uint32_t box[3] = {32, (uint32_t)PT, 1};
  return encode_tensor_map(m, act, 3, dims, strides, box, 1);
}
int tc_make_fc1_map(CUtensorMap* m, const float* w) {  // [2*128 rows (hl, hid)][64]
  uint64_t dims[2] = {64, 256};
  uint64_t strides[1] = {64 * 4};
  uint32_t box[2] = {32, 128};
  return encode_tensor_map(m, w, 2, dims, strides, box, 1);
}
int tc_make_fc2_map(CUtensorMap* m, const float* w, int N2) {  // [2*N2 rows (hl, f)][128]
  uint64_t dims[2] = {128, (uint64_t)2 * N2};
  uint64_t strides[1] = {128 * 4};
  uint32_t box[2] = {32, (uint32_t)N2};
  return encode_tensor_map(m, w, 2, dims, strides, box, 1);
}

int launch_proj_tc(const ProjArgs& pa, const CUtensorMap& tmX, const CUtensorMap& tmW1, const CUtensorMap& tmW2,
                   cudaStream_t st, int b0, int nb, int bf16) {
  // samples [b0, b0 + nb) of the batch pa.B (nb < 0: all of them)
  TcProjArgs a{};
  a.bf16 = bf16;
  a.fc1b = pa.fc1b, a.fc2b = pa.fc2b, a.aff_a = pa.aff_a, a.aff_b = pa.aff_b;
  a.chan = pa.chan, a.out_off = pa.out_off, a.st_off = pa.st_off, a.out = pa.out, a.state = pa.state;
  if (nb < 0) b0 = 0, nb = pa.B;
  a.rows = nb * pa.T * pa.H, a.row0 = b0 * pa.T * pa.H, a.Tv = pa.T, a.H = pa.H, a.W = pa.W, a.Tp = pa.Tp, a.Hp = pa.Hp;
  tc_proj_tile(pa.W, &a.PT, &a.NTW);
  a.G = std::max(1, std::min(148 / a.NTW, a.rows));
  a.Fout = pa.Fout, a.N2 = tc_proj_n2(pa.Fout), a.c_out = pa.c_out, a.c_in = pa.c_in;
  a.out_sB = pa.out_sB, a.out_sT = pa.out_sT, a.st_sB = pa.st_sB, a.st_sT = pa.st_sT;
  // TMA-store epilogue when the output geometry allows: out viewed as [B][frames][H][W*c_out] fp32
  CUtensorMap tmOut = tmX, tmState = tmX;
  {
    const bool ndim3 = pa.out_sT != 0;
    const int r = ndim3 ? (int)(pa.out_sT / ((long long)pa.H * pa.W * pa.c_out)) : 1;
    const int NF = ndim3 ? r : pa.Fout / pa.c_out;            // frames written per tile
    const int seg = a.PT * pa.c_out;                          // floats of one frame of one tile
    const int nb = ceil_div(seg, 256), OB = seg / nb;
    const int box_floats = round_up(NF * OB, 32);
    const long long frame_elems = (long long)pa.H * pa.W * pa.c_out;
    const bool ok = OB * nb == seg && OB % 4 == 0 && (pa.W * pa.c_out) % 4 == 0 && pa.c_out <= 3 && NF <= 256 &&
                    NF * pa.c_out == (ndim3 ? pa.Fout : pa.Fout) && nb * box_floats * 4 <= TCP_ST - 2048 &&
                    ((uintptr_t)pa.out & 15) == 0 && (pa.out_sB % 4) == 0 && pa.out_sB % frame_elems == 0;
    if (ok) {
      const long long frames_total = ndim3 ? (long long)pa.T * r : NF;  // frames this launch may touch per sample
      uint64_t dims[4] = {(uint64_t)pa.W * pa.c_out, (uint64_t)pa.H, (uint64_t)frames_total, (uint64_t)pa.B};
      uint64_t strides[3] = {(uint64_t)pa.W * pa.c_out * 4, (uint64_t)frame_elems * 4, (uint64_t)pa.out_sB * 4};
      uint32_t box[4] = {(uint32_t)OB, 1, (uint32_t)NF, 1};
      B2_TRY(encode_tensor_map(&tmOut, pa.out, 4, dims, strides, box, 0));
      a.out_tma = 1, a.o_nb = nb, a.o_OB = OB, a.o_NF = NF, a.o_box_floats = box_floats, a.o_r = r, a.ndim3 = ndim3;
      if (pa.state && pa.c_in == pa.c_out && ((uintptr_t)pa.state & 15) == 0 && pa.st_sB % 4 == 0) {
        uint64_t sstr[3] = {(uint64_t)pa.W * pa.c_in * 4, (uint64_t)frame_elems * 4, (uint64_t)pa.st_sB * 4};
        B2_TRY(encode_tensor_map(&tmState, pa.state, 4, dims, sstr, box, 0));
        a.st_tma = 1;
      } else if (pa.state) {
        a.out_tma = 0;  // parameter channels interleave with the prediction: keep the scattered stores
      }
    }
  }
  if (a.out_tma) {
    B2_CUDA(cudaFuncSetAttribute(tc_proj_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM));
    B2_CUDA(launch_kernel(tc_proj_kernel<true>, dim3(a.NTW * a.G), dim3(TCP_THREADS), (size_t)TCP_SMEM, st, a, tmX, tmW1,
                          tmW2, tmOut, tmState));
  } else {
    B2_CUDA(cudaFuncSetAttribute(tc_proj_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM));
    B2_CUDA(launch_kernel(tc_proj_kernel<false>, dim3(a.NTW * a.G), dim3(TCP_THREADS), (size_t)TCP_SMEM, st, a, tmX, tmW1,
                          tmW2, tmOut, tmState));
  }
  B2_LAUNCHED("tc_proj_kernel");
  return 0;
}

}  // namespace b200fno
