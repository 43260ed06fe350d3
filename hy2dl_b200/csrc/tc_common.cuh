// Blackwell (sm_100a) primitives used by the tensor-core LSTM path: mbarrier, TMEM alloc,
// tcgen05.mma (kind::tf32, A from TMEM or smem, B from smem), tcgen05.ld/st, descriptors, and
// 1-D bulk async copies (TMA without a tensor map).  Inline PTX only; no CUTLASS dependency.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hy2dl {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// Non-suspending poll (test_wait): for waits on the critical path of a latency-bound chain.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "SPIN_LOOP:\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra SPIN_DONE;\n\t"
      "bra SPIN_LOOP;\n\t"
      "SPIN_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// Same for warps that are not on the critical path: back off between polls so that the spin does not
// take issue slots from the warps sharing the scheduler.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 1000;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(256);
  }
}

// One lane of a fully converged warp.  Code under `if (elect_one())` is what the compiler recognises as a
// single-thread region: operands of tcgen05.mma / commit stay in uniform registers, where a region guarded by
// `lane == 0` is divergent code and pays an ELECT + 4 x R2UR.BROADCAST waterfall per MMA instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- proxies / fences ------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one warp, all 32 lanes) ------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}

// ---- descriptors -----------------------------------------------------------------------------
// Instruction descriptor, kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor bit layout):
//   [4,6) c_format = 1 (F32); [7,10) a_format = 2 (TF32); [10,13) b_format = 2 (TF32);
//   [15] a_major, [16] b_major (0 = K-major, 1 = MN-major); [17,23) N >> 3; [24,29) M >> 4.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts, 16-byte units:
//   K-major : core matrix = 8 MN-rows x 16 B (rows 16 B apart); LBO = step between the two 16-B
//             K-chunks of one MMA; SBO = step between 8-row groups along MN.
//   MN-major: core matrix = 16 B along MN x 8 K-rows (rows 16 B apart); SBO = step between 16-B
//             groups along MN; LBO = step between 8-row groups along K.
// bits [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) layout = 0.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}

// Same, SWIZZLE_128B_BASE32B (layout type 1): the only MN-major layout kind::tf32 accepts.  Atom = 32 MN
// elements (128 B) x 4 K-rows 128 B apart; LBO = step between 32-element MN blocks, SBO = step between
// 4-row K groups; in K-row k the 32-byte unit u of the 128-byte row sits at unit position u ^ (k & 3)
// (byte-address bits [7,9) XORed into bits [5,7); verified on hardware, tools/dbg_wgrad.py).
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return smem_desc(saddr, lbo_bytes, sbo_bytes) | (1ull << 61);
}

// ---- RI32 stream layout -------------------------------------------------------------------------
// [t][G = ceil(B/32)][F/32][32 rows][32 floats]; in row r the 8-float unit u is stored at unit u ^ (r & 3):
// a (t, group, 32-feature block) slab is 4 KB contiguous and is exactly the SWIZZLE_128B_BASE32B MN-major
// operand tile for K = 32 rows.  A row thread reads/writes 16-byte chunks c = feature / 4; the 8 chunks
// of a block are one 128-byte line of its row, two consecutive chunks one 32-byte sector.
__device__ __forceinline__ int64_t ri32_idx4(int64_t t, int64_t G, int64_t gidx, int FB, int c, int row) {
  const int pos = ((((c & 7) >> 1) ^ (row & 3)) << 1) | (c & 1);
  return ((t * G + gidx) * FB + (c >> 3)) * 256 + row * 8 + pos;         // float4 index; FB = F / 32
}
__device__ __forceinline__ float4 ri32_ld(const float4* base, int64_t t, int64_t G, int64_t gidx, int FB, int c, int row) {
  return __ldg(base + ri32_idx4(t, G, gidx, FB, c, row));
}
__device__ __forceinline__ void ri32_st(float4* base, int64_t t, int64_t G, int64_t gidx, int FB, int c, int row, float4 v) {
  base[ri32_idx4(t, G, gidx, FB, c, row)] = v;
}
// 32-byte unit U = feature / 8 of row `row`: float index.  One unit = one sector = one 256-bit access.
__device__ __forceinline__ int64_t ri32_unit(int64_t t, int64_t G, int64_t gidx, int FB, int U, int row) {
  return ((t * G + gidx) * FB + (U >> 2)) * 1024 + row * 32 + (((U & 3) ^ (row & 3)) << 3);
}
// ---- RB8 tape layout of the H = 128 / 256 tensor-core chain: [t][G = ceil(B/32)][F/8][32 rows][8 floats].
// The row-owning threads of those kernels (TMEM lane = sequence) move 8 floats = one 32-byte sector per access; in a
// row-major tape the 32 lanes of a warp then touch 32 different lines (the LSU serialises them: the first version of
// the BPTT kernel spent 34 of 52 ms on tape accesses), here they touch one contiguous 1 KB run.  Float offset of
// the 8-float block `blk` of sequence b at step t; the next block of the same sequence is 256 floats further.
__host__ __device__ __forceinline__ int64_t rb8_off(int64_t t, int64_t G, int64_t b, int F, int blk) {
  return ((t * G + (b >> 5)) * (F >> 3) + blk) * 256 + (b & 31) * 8;
}
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ldg256(const float* p) {                    // read-only path, 32-byte aligned
  f8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
// coherent 256-bit load (data this thread stored earlier in the kernel: the read-only path must not be used)
__device__ __forceinline__ f8 ld256(const float* p) {
  f8 r;
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p)
               : "memory");
  return r;
}
// plain (write-back) 256-bit store: data that is re-read soon
__device__ __forceinline__ void st256(float* p, const f8& r) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
               "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
               : "memory");
}
__device__ __forceinline__ void stg256(float* p, const f8& r) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
               "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
               : "memory");
}
__device__ __forceinline__ f8 f8_zero() { f8 r; for (int i = 0; i < 8; ++i) r.v[i] = 0.f; return r; }

// ---- MMA issue (single thread) ---------------------------------------------------------------
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- CTA pair (cta_group::2): two CTAs of a cluster on the SMs of one TPC execute one MMA with M = 256 -- each CTA supplies
// its 128 rows of A (shared memory or TMEM, same addresses in both) and HALF of the B tile (N/2 columns, CTA rank 0 the
// first half), and receives its 128 rows of D in its own TMEM.  Only the leader (rank 0) issues; completion is multicast
// to the mbarriers at the same offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  // (the form CUTLASS's ClusterBarrier::arrive(cta_id) uses; with .release.cluster every arrival cost ~0.4 us, measured)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire: the arrivals may come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITC_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAITC_DONE;\n\t"
      "bra WAITC_LOOP;\n\t"
      "WAITC_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {      // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void mma2_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void mma2_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// all previously issued MMAs of this (leader) thread arrive on `bar` of BOTH CTAs when they complete
__device__ __forceinline__ void mma2_commit(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

// ---- TMEM <-> registers (warp-wide; warp w of a warpgroup touches lanes 32w .. 32w+31) --------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- 1-D bulk async copy global -> shared (TMA engine, no tensor map); completes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- 3xTF32 operand split: x = hi + lo with hi = x rounded to tf32 (nearest, ties away: the integer
// add carries into the exponent exactly like cvt.rna.tf32.f32), lo = x - hi exact in fp32, |lo| <= 2^-11 |x|.
// The tensor core reads the top 19 bits of lo, i.e. drops <= 2^-21 |x|.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// same with lo rounded to tf32 as well (used where the split is computed once: packed weights)
__device__ __forceinline__ void split_tf32_rr(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xFFFFE000u;
}

// fast, accurate-enough activations on the MUFU ex2/rcp path (abs error ~1e-7)
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// One reciprocal for the four gates of a hidden unit.  The accumulator holds z' = -log2(e) z (i, f, o) and
// -2 log2(e) z (g), bias included (see lstm_tc_pack_w_kernel).  With A = 1 + 2^zi', F = 1 + 2^zf', G = 1 + 2^zg',
// O = 1 + 2^zo' and r = 1 / (A F G O):  i = r F G O, f = r A G O, o = r A F G, g = tanh(zg) = (2 - G) r A F O.
// Arguments are clamped (sigmoid below -20 / tanh below -10 differ from their limits by < 4e-9) so that the
// product stays finite.  5 MUFU (4 ex2 + rcp) instead of 8: the XU pipe is the epilogue's narrowest resource.
__device__ __forceinline__ void lstm_gates(float zi, float zf, float zg, float zo, float& gi, float& gf, float& gg, float& go) {
  constexpr float kClamp = 28.853900817779268f;   // 20 log2(e)
  const float A = 1.f + fast_ex2(fminf(zi, kClamp));
  const float F = 1.f + fast_ex2(fminf(zf, kClamp));
  const float G = 1.f + fast_ex2(fminf(zg, kClamp));
  const float O = 1.f + fast_ex2(fminf(zo, kClamp));
  const float AF = A * F, GO = G * O;
  const float r = fast_rcp(AF * GO);
  const float rAF = r * AF, rGO = r * GO;
  gi = rGO * F;
  gf = rGO * A;
  go = rAF * G;
  gg = (2.f - G) * (rAF * O);
}
// The same with one reciprocal per gate (8 MUFU): every gate carries the rounding of its own ex2 / rcp only, where
// the shared reciprocal above carries the errors of all four exponentials.
__device__ __forceinline__ void lstm_gates_sep(float zi, float zf, float zg, float zo, float& gi, float& gf, float& gg, float& go) {
  constexpr float kClamp = 28.853900817779268f;
  gi = fast_rcp(1.f + fast_ex2(fminf(zi, kClamp)));
  gf = fast_rcp(1.f + fast_ex2(fminf(zf, kClamp)));
  go = fast_rcp(1.f + fast_ex2(fminf(zo, kClamp)));
  gg = fmaf(2.f, fast_rcp(1.f + fast_ex2(fminf(zg, 2.f * kClamp))), -1.f);
}
// tanh(x) = 1 - 2 / (1 + e^2x): 2 MUFU + 3 FMA-pipe instructions; saturates correctly at +-inf
__device__ __forceinline__ float tanh_mufu(float x) {
  return fmaf(-2.f, fast_rcp(1.f + fast_ex2(2.885390081777927f * x)), 1.f);
}


}  // namespace tc
}  // namespace hy2dl
