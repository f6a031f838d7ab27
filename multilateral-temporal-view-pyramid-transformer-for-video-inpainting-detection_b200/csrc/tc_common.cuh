// PTX wrappers for the Blackwell tensor-core path: mbarrier, TMA (tiled and im2col), tcgen05 MMA / commit / TMEM load.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mumpy {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Parity wait with a watchdog: a pipeline bug becomes a trapped launch (reported through the C ABI) instead
// of a hung GPU.  The timer is only read every 4096 failed probes, so the fast path is untouched.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (++spins & 4095u) == 0) {
      const uint64_t now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();     // 4 s without progress
    }
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
// im2col-mode TMA over an NHWC tensor: `pixels` consecutive output positions starting at base pixel (w,h,n), channel
// block starting at c, filter tap (off_w, off_h); out-of-image taps are zero filled (the convolution's padding).
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants: the two CTAs of a (2,1,1) cluster run one M=256 MMA; the leader (rank 0) issues ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// (default .release.cta semantics, as CUTLASS uses: an explicit .release.cluster makes ptxas emit MEMBAR.ALL.GPU in front of
// every arrive, which serialises the TMA pipeline -- measured 1.1 us per k-block)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// TMA loads of a CTA pair: data lands in the issuing CTA's shared memory, the bytes are counted on the LEADER's barrier
// (`bar_cluster` = shared::cluster address of that barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c, int w, int h, int n,
                                                        uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One lane of the (converged) warp.  The single-thread instructions of the tensor-core path (tcgen05.mma / commit, TMA loads) are
// issued inside `if (elect_one())` by a warp that runs its whole loop converged: ptxas then keeps the descriptors / addresses in
// uniform registers.  Under a `lane == 0` branch it cannot (the region is divergent): every operand goes through R2UR and every
// issue through an ELECT waterfall loop, which made the old main loops issue-bound (measured with tools/umma_bench.cu: 230-290
// clk per MMA instead of the 128 clk floor of a 128 x 256 x 16 MMA; with elect.sync the same loop runs at the floor).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred px;\n\t"
      "elect.sync _|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128B-swizzled operand tile (rows of 128 B, 8-row groups 1024 B apart): UMMA::SmemDescriptor with
// start>>4 [0,14), LBO [16,30) (unused for swizzled K-major, 1), SBO=1024>>4 [32,46), version=1 [46,48),
// layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// instruction descriptor: D=f32 [4,6)=1, A format [7,10) and B format [10,13) (0 = f16, 1 = bf16), K-major both,
// N>>3 [17,23), M>>4 [24,29)
__host__ __device__ inline uint32_t make_idesc_16_f32(int M, int N, bool f16) {
  const uint32_t fmt = f16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// GELU(x) = relu(x) - |x| * Phi(-|x|), with Phi(-u) = 0.5 erfc(u / sqrt2) = 2^Q(u) on u = min(|x|, 5 sqrt2): degree-6
// minimax fit of log2 Phi(-u) evaluated in s = -u (so that the last step is one FMA: relu(x) + s * 2^Q), |error| <= 5.5e-7
// absolute in fp32 over the whole real line (below erff's own rounding at 16-bit output precision).  Two elements per call on
// the packed fp32x2 pipe of sm_100 (FFMA2): per pair 4 FMNMX + 7 FFMA2 + 2 MUFU.EX2, i.e. 6.5 issue slots per element
// instead of ~40 for erff (which made the fc1 epilogue 2.4x slower than its MMAs) and 12 for the scalar form.
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  const float2 s = make_float2(fmaxf(-fabsf(x.x), -7.0710678f), fmaxf(-fabsf(x.y), -7.0710678f));
  float2 p = make_float2(1.9175164197804406e-05f, 1.9175164197804406e-05f);
  p = __ffma2_rn(p, s, make_float2(0.0006586086819879711f, 0.0006586086819879711f));
  p = __ffma2_rn(p, s, make_float2(0.0077544208616018295f, 0.0077544208616018295f));
  p = __ffma2_rn(p, s, make_float2(0.05296541005373001f, 0.05296541005373001f));
  p = __ffma2_rn(p, s, make_float2(-0.4590602517127991f, -0.4590602517127991f));
  p = __ffma2_rn(p, s, make_float2(1.1511220932006836f, 1.1511220932006836f));
  p = __ffma2_rn(p, s, make_float2(-0.9999997019767761f, -0.9999997019767761f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(p.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(p.y));
  return __ffma2_rn(s, e, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}
__device__ __forceinline__ float gelu_fast(float x) { return gelu_fast2(make_float2(x, x)).x; }

// Four pairs at once, every Horner step written across the four before the next step: the same operations per element (bit-identical
// to gelu_fast2), but four independent dependency chains in flight.  Back-to-back gelu_fast2 calls compile to one pair's seven
// dependent FFMA2 + MUFU after the other (cuobjdump: ILP 1, ~85 clk per pair and warp) -- the GELU epilogues were bound by that
// latency chain, not by issue slots.
__device__ __forceinline__ void gelu_fast2_x4(float2 (&x)[4]) {
  float2 s[4], p[4], e[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) s[k] = make_float2(fmaxf(-fabsf(x[k].x), -7.0710678f), fmaxf(-fabsf(x[k].y), -7.0710678f));
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] = __ffma2_rn(make_float2(1.9175164197804406e-05f, 1.9175164197804406e-05f), s[k], make_float2(0.0006586086819879711f, 0.0006586086819879711f));
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] = __ffma2_rn(p[k], s[k], make_float2(0.0077544208616018295f, 0.0077544208616018295f));
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] = __ffma2_rn(p[k], s[k], make_float2(0.05296541005373001f, 0.05296541005373001f));
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] = __ffma2_rn(p[k], s[k], make_float2(-0.4590602517127991f, -0.4590602517127991f));
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] = __ffma2_rn(p[k], s[k], make_float2(1.1511220932006836f, 1.1511220932006836f));
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] = __ffma2_rn(p[k], s[k], make_float2(-0.9999997019767761f, -0.9999997019767761f));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[k].x) : "f"(p[k].x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[k].y) : "f"(p[k].y));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = __ffma2_rn(s[k], e[k], make_float2(fmaxf(x[k].x, 0.0f), fmaxf(x[k].y, 0.0f)));
}

}  // namespace mumpy
