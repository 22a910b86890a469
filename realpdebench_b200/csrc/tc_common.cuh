// sm_100a building blocks: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (TMEM alloc / mma / ld / st /
// commit) as inline PTX, plus the UMMA shared-memory and instruction descriptors for kind::tf32.
// Every wait is bounded: a barrier that never completes traps instead of hanging the GPU.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200fno {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (TMA store, tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware (no issue slots burnt) until the phase
// completes or the hint expires, instead of spinning on the barrier.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)  // suspend-time hint, ns
      : "memory");
  return ok != 0;
}
// non-blocking phase test (no suspend): for producers that feed a second ring opportunistically
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: far beyond any legitimate wait in these kernels; on expiry the kernel traps
// (the launch fails, the box stays healthy) instead of hanging the GPU.
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint64_t t0 = 0;
#pragma unroll 1
  for (uint32_t i = 0;; ++i) {
    if (mbar_try_wait(bar, parity)) return;
    if ((i & 63u) == 63u) {  // wall-clock bound: 2 s without progress is a deadlock, not a wait
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) __trap();
    }
  }
}

// explicit shared-state-space vector accesses (keeps ptxas from falling back to generic LD/ST)
// Explicit shared-space scalar accesses: pointers derived from the aligned dynamic-shared base lose their
// address space (ptxas then emits generic LD.E / ST.E, which go through the address-space check).
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void lds64(uint32_t addr, uint32_t& a, uint32_t& b) {
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
}
// volatile keeps it ordered against the (volatile) barrier / fence asm statements; no "memory" clobber, so the
// compiler may still hoist ordinary loads (e.g. the next element's table entry) above it
__device__ __forceinline__ void sts32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One lane of a converged warp.  ptxas recognises elect.sync-guarded regions as single-thread uniform
// code and emits the tcgen05 / TMA instructions directly (a plain `lane == 0` test makes it wrap every
// UTCHMMA in a per-lane serialisation loop: ~120 cycles per MMA instead of the tensor-pipe rate).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

// ---- TMA ----------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {  // smem sources of all but the N newest groups are reusable
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- tcgen05 --------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all prior tcgen05.mma of this thread complete -> arrive(1) on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this thread's lane, 32 / 16 / 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,"
      "%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors ------------------------------------------------------------------------------------
constexpr uint32_t LAYOUT_SW128 = 2;  // UMMA::LayoutType::SWIZZLE_128B

// Shared-memory matrix descriptor (sm_100 format, version bit 46).
// K-major SW128 tile: rows of 128 B (32 tf32), 8-row groups every `sbo` bytes (1024 when dense), lbo unused.
// MN-major SW128 tile: each K row is 128 B of 32 consecutive M/N elements, 8 K rows = one 1024 B atom;
//   `lbo` = bytes between successive 32-element M/N blocks, `sbo` = bytes between successive 8-row K groups.
constexpr uint32_t LAYOUT_SW128_BASE32B = 1;  // 32-bit MN-major operands: Swizzle<2,5,2>, 4-row atoms
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout = LAYOUT_SW128) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// Instruction descriptor, kind::tf32, fp32 accumulate.  a_mn / b_mn: 1 = MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4)                    // c_format = F32
         | (2u << 7) | (2u << 10)     // a_format = b_format = TF32
         | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// byte offset of 16-byte chunk `c` of row `r` inside a 128B-swizzled tile whose rows are 128 B
__device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

// 3xTF32 split: hi = x ROUNDED TO NEAREST onto the 19 bits tf32 keeps (so the MMA's own truncation of the operand is a
// no-op), lo = x - hi exact in fp32 and signed, |lo| <= 2^-11 |x|.  With a truncated hi (what the MMA would make of the raw
// fp32 value) lo is one-signed and twice as large, and the dropped lo*lo term becomes a coherent bias: one spectral
// convolution then carries 1.0e-6 relative error against 1.4e-7 with the rounded split (plain fp32 FFMA: 4e-7;
// profiles/split_study.py).
__device__ __forceinline__ uint32_t tf32_rn_bits(uint32_t u) { return (u + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(tf32_rn_bits(__float_as_uint(x))); }

// GELU (erf form, the F.gelu default of fno.py:119,124) without branches:
//   v*Phi(v) = max(v,0) - |v| * 0.5*erfc(|v|/sqrt2),  erfc from Abramowitz-Stegun 7.1.26 (|eps| <= 1.5e-7).
// Max abs error vs the exact function 5.3e-7 over [-12,12] (torch's own fp32 gelu: 1.2e-6), relative L2
// 9e-8 on N(0,1) inputs.  14 instructions per element, two of them MUFU (rcp.approx, ex2.approx).
__device__ __forceinline__ float gelu_erf_fast(float v) {
  const float av = fabsf(v);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, av, 1.0f)));
  float poly = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  poly = fmaf(t, poly, 0.5f * 1.421413741f);
  poly = fmaf(t, poly, 0.5f * -0.284496736f);
  poly = fmaf(t, poly, 0.5f * 0.254829592f);
  poly *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * v * (-0.5f * 1.4426950408889634f)));
  return fmaf(-av * poly, e, fmaxf(v, 0.0f));
}

// ---- packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2): two elements per instruction ----------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 f2_splat(float c) { return f2_pack(c, c); }

// GELU (erf form, F.gelu default; fno.py:119,124) on two values at once, ONE MUFU op per element:
//   gelu(v) = v * Phi(v) = max(v, 0) - |v| * Phi(-|v|),   Phi(-a) = 2^Q(a)
// with Q a degree-6 polynomial fit of log2(Phi(-a)) on [0, 5.6] (weighted by a * Phi(-a), the size of the term it
// feeds; beyond 5.6 the argument is clamped, the term is < 6e-8 * |v| there).  Max abs error 2.5e-7 in fp32 (torch's
// own fp32 gelu: 1.2e-6).  The previous Abramowitz-Stegun 7.1.26 form needed rcp + ex2 per element; the epilogues
// are MUFU / issue bound, so this halves their special-function traffic (7 packed FMA2 + 2 ex2 per pair).
__device__ __forceinline__ void gelu_erf_fast2(float& x0, float& x1) {
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  const f32x2 ac = f2_pack(fminf(a0, 5.6f), fminf(a1, 5.6f));
  f32x2 q = f2_fma(ac, f2_splat(3.457919228821993e-05f), f2_splat(-0.0007809283561073244f));
  q = f2_fma(ac, q, f2_splat(0.008115331642329693f));
  q = f2_fma(ac, q, f2_splat(-0.053460296243429184f));
  q = f2_fma(ac, q, f2_splat(-0.4587385654449463f));
  q = f2_fma(ac, q, f2_splat(-1.1512112617492676f));
  q = f2_fma(ac, q, f2_splat(-0.9999921321868896f));
  float s0, s1, e0, e1;
  f2_unpack(q, s0, s1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(s0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(s1));
  f2_unpack(f2_fma(f2_pack(-a0, -a1), f2_pack(e0, e1), f2_pack(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f))), x0, x1);
}
// bf16 compute mode (torch.autocast(bfloat16) semantics of the reference: Linear / Conv operands are cast to bf16,
// products accumulate in fp32): the operand rounded to nearest-even bf16.  A bf16 value is exactly representable in
// tf32, so ONE kind::tf32 MMA pass on rounded operands gives the bf16 x bf16 -> fp32 products exactly.
__device__ __forceinline__ uint32_t bf16_rn_bits(uint32_t u) { return (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u; }
__device__ __forceinline__ float bf16_rn(float x) { return __uint_as_float(bf16_rn_bits(__float_as_uint(x))); }

// 3xTF32 split of two values: h = rn_tf32(x), r = x - h (in place)
__device__ __forceinline__ void tf32_split2(uint32_t& r0, uint32_t& r1, uint32_t& h0, uint32_t& h1) {
  const f32x2 x = f2_pack(__uint_as_float(r0), __uint_as_float(r1));
  h0 = tf32_rn_bits(r0), h1 = tf32_rn_bits(r1);
  float a, b;
  f2_unpack(f2_fma(f2_pack(__uint_as_float(h0), __uint_as_float(h1)), f2_splat(-1.0f), x), a, b);
  r0 = __float_as_uint(a), r1 = __float_as_uint(b);
}

}  // namespace tc

// host: cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
int encode_tensor_map(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle /*0 none, 1 128B, 2 128B with 32B atoms*/);

}  // namespace b200fno
