// TF32 tensor-core GEMM for sm_100a: TMA-staged operand tiles, tcgen05.mma with the accumulator in
// TMEM, mbarrier producer/consumer pipeline, fused epilogue straight out of TMEM.
//
//   C[M,N] = opA(A)[M,K] * opB(B)[N,K]^T  (+ epilogue of gemm_epilogue.cuh)
//
// One CTA computes one BLOCK_M x BLOCK_N output tile (BLOCK_M = 128, BLOCK_N = 128 or 64) for one
// K-slice (grid.z = split-K slices; partial tiles go to a workspace and a reduce kernel applies the
// epilogue, which keeps weight-gradient GEMMs -- tiny output, very long K -- on all SMs and
// deterministic).  Warp roles (192 threads):
//     warp 0   : TMA producer   (cp.async.bulk.tensor.2d -> 128B-swizzled smem, 4-stage ring)
//     warp 1   : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma.kind::tf32)
//     warps 2-5: epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// Operands are fp32 in HBM; the tensor maps use the TFLOAT32 data type and the MMA reads them as
// TF32, so no conversion pass or shadow copy exists.  Both operand majors are supported because the
// backward pass needs them (dgrad: B is [K,N]; wgrad: A is [K,M] and B is [K,N]):
//     K-major  (trans = 0): smem tile = rows x 32 floats, one 128-byte swizzle row per tile row
//                           (TMA SWIZZLE_128B, UMMA layout SWIZZLE_128B)
//     MN-major (trans = 1): smem tile = (rows/32) boxes of [BLOCK_K x 32 floats]; 32-bit MN-major
//                           operands only exist in the 32-byte-atom flavour of the 128-byte swizzle
//                           (TMA SWIZZLE_128B_ATOM_32B, UMMA layout SWIZZLE_128B_BASE32B, atoms of 4 k-rows)
// Rows / K tails are handled by TMA out-of-bounds zero fill and predicated stores.
//
// X3 = true ("bf16x3", gb_gemm_args.precision = 3): fp32-class accuracy on the tensor cores.  TF32 keeps 10 mantissa
// bits per operand; through the ~45 dependent GEMMs of the model that is 2e-3 .. 3.5e-3 on the gated torsion amplitudes,
// outside the 1e-3 contract.  Here four converter warps split every staged fp32 operand element into bf16 hi + bf16 lo
// (hi = rn(a), lo = rn(a - hi): 16 mantissa bits together) and the MMA warp issues hi*hi + lo*hi + hi*lo with
// kind::f16 -- three bf16 MMAs run at 1.5x the time of one TF32 MMA and the split operands take exactly the bytes of the
// fp32 ones, so HBM / L2 traffic is unchanged and nothing but this kernel knows about the format:
//     TMA (fp32 tile, as above) -> ring stage -> converter warps rewrite the stage in place -> tcgen05.mma kind::f16 x3 -> TMEM
// converted tile = rows x 128 B, row = [32 bf16 hi | 32 bf16 lo] of one 32-float k-block, ALWAYS K-major SWIZZLE_128B: the
// same shared-memory descriptor form the TF32 path uses for K-major fp32 (hi part at +0 / +32, lo part at +64 / +96
// bytes inside the swizzle atom).  MN-major fp32 operands (dgrad / wgrad) are transposed by the converter on the fly
// (un-swizzled TMA boxes, conflict-free 4-byte column reads), so the MMA only ever sees K-major bf16.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "f32x2.cuh"
#include "gemm_epilogue.cuh"
#include "mbar.cuh"

#ifndef GB_X3_SPLIT
#define GB_X3_SPLIT 1   // 0: Veltkamp split in packed fp32 + integer rounding of the remainder, 1: cvt.rn.bf16x2.f32 for both planes
#endif

namespace gb {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;          // floats per stage along K = one 128-byte swizzle atom
constexpr int TC_THREADS = 320;        // TMA warp + MMA warp + 8 epilogue warps
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 r;\n\t"
      "elect.sync r|P, %1;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred)
      : "r"(0xffffffffu));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- CTA-pair (cta_group::2) helpers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// Execution barrier over the CTAs of the cluster.  Relaxed arrival (what cute::cluster_arrive_relaxed + cluster_wait issue
// after fence_barrier_init): the release / acquire forms put MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of the arrival,
// ~1500 cycles at the start and at the end of every pair kernel.  What has to be visible across the pair at these two
// points is published separately: the mbarrier initialisation by fence.mbarrier_init.release.cluster, and at the end of
// the kernel nothing is exchanged -- the barrier only keeps shared / tensor memory alive until both CTAs are done.
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
// Arrival on a barrier of another CTA of the cluster.  The default form (what cutlass::arch::ClusterBarrier::arrive(cta_id)
// issues) is ONE SYNCS.ARRIVE; spelling out .release.cluster makes ptxas emit MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front
// of it -- measured with clock64 stamps (tools/gemm_trace.py): ~1500 cycles per arrival, which serialised the bf16x3
// pair pipeline at one k-block per 1600 cycles.  Data handed over with it is already fenced by its writers
// (fence.proxy.async + CTA-scope barrier) or read out of tensor memory (tcgen05.wait::ld).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// data lands in THIS CTA's shared memory, the transaction bytes complete on the barrier at `mbar_cluster_addr`
// (the leader CTA's full barrier): what cute::SM100_TMA_2SM_LOAD does
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t mbar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the barrier at the same offset in every CTA of `mask` once all prior MMAs of this thread retire
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// generic-proxy shared-memory writes (the converter's st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Two fp32 values (one 64-bit register) -> packed bf16 pair of the leading parts and packed bf16 pair of the remainders:
//   hi = rn_bf16(a);  lo = rn_bf16(a - hi)  (a - hi is exact in fp32), so hi + lo carries 16 significant bits
//   (|error| <= 2^-17 |a|, unbiased).
// GB_X3_SPLIT = 1 (default): cvt.rn.bf16x2.f32 (SASS F2FP.BF16.F32.PACK_AB) for both planes + shift / mask / one packed
//   subtract -- 5 instructions per pair.
// GB_X3_SPLIT = 0: Veltkamp's splitting in PACKED fp32 arithmetic (c = a * (2^16 + 1); hi = c - (c - a), identical to
//   cvt.rn on every finite input tested) + integer half-up rounding of the remainder: 4 packed FP + 2 integer adds +
//   2 PRMT per pair.  `z` is a packed zero the compiler cannot see through: c must be ROUNDED before c - a is formed,
//   but ptxas contracts a packed multiply into the dependent add / sub even with explicit .rn (c - a became
//   fma(a, 65537, -a) = 65536 a exactly and results fell back to bf16 accuracy); with c = fma(a, K, z) there is no bare
//   multiply left to contract.
// History (profiles/r2_summary.md): the first F2FP version looked XU-bound under ncu and was replaced by integer rounding
// and then by the Veltkamp form -- all three were measured while the pipeline was serialised by a cluster-scope release
// on the peer CTA's hand-off (see mbar_arrive_cluster), which hid the real ranking.  Measured once that was fixed
// (tools/microbench/pipes.cu: F2FP issues at the same 1.9 warp instructions / clk / SM as PRMT / LOP3): training step
// 7.17 ms with GB_X3_SPLIT = 0, 6.88 ms with 1.
__device__ __forceinline__ void split_pair2(F2 a, F2 z, uint32_t& hi, uint32_t& lo) {
#if GB_X3_SPLIT == 1
  // cvt.rn.bf16x2.f32 (F2FP) for both planes: 2 F2FP + shift + mask + one packed subtract per pair
  const float a0 = gb::lo(a), a1 = gb::hi(a);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(a1), "f"(a0));
  const F2 l = a - f2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xffff0000u));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(gb::hi(l)), "f"(gb::lo(l)));
  (void)z;
#else
  const F2 c = fma2(a, f2(65537.0f), z);
  const F2 h = c - (c - a);
  const F2 l = a - h;
  hi = __byte_perm((uint32_t)(h.v & 0xffffffffull), (uint32_t)(h.v >> 32), 0x7632);   // {h1[31:16], h0[31:16]}
  const unsigned long long lr = l.v + 0x0000800000008000ull;   // round the remainders half-up to bf16 (no carry between the halves
                                                                // for finite values); truncation alone left 2^-16 relative error
  lo = __byte_perm((uint32_t)(lr & 0xffffffffull), (uint32_t)(lr >> 32), 0x7632);
#endif
}
__device__ __forceinline__ void split8(const float (&v)[8], F2 z, uint4& hi, uint4& lo) {
  split_pair2(f2(v[0], v[1]), z, hi.x, lo.x);
  split_pair2(f2(v[2], v[3]), z, hi.y, lo.y);
  split_pair2(f2(v[4], v[5]), z, hi.z, lo.z);
  split_pair2(f2(v[6], v[7]), z, hi.w, lo.w);
}

// bf16x3 warp roles (512 threads, 128 registers): TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter, half of
// the BN columns each -- with four the epilogue, 18-20 k cycles per 128 x 256 tile, was longer than the tile's mainloop),
// 6 converter warps in three groups of two: group g rewrites the stages of k-blocks g, g + 3, ..., so three stages are
// being converted at any time.  (18 warps = 8 + 8 cap the kernel at 96 registers: the epilogue spills and the step is
// 0.8 ms slower; A/B builds: tools/gemm_trace.py build <name>.so -DGB_X3_CONV_WARPS=.. -DGB_X3_CONV_GROUPS=.. -DGB_X3_EPI_WARPS=..)
#ifndef GB_X3_CONV_WARPS   // -D overrides: A/B builds through tools/gemm_trace.py
#define GB_X3_CONV_WARPS 2
#define GB_X3_CONV_GROUPS 3
#define GB_X3_EPI_WARPS 8
#endif
constexpr int TC_CONV_WARPS = GB_X3_CONV_WARPS;    // converter warps per group (= warps that share one stage)
constexpr int TC_CONV_GROUPS = GB_X3_CONV_GROUPS;
constexpr int TC_EPI_WARPS_X3 = GB_X3_EPI_WARPS;
constexpr int TC_THREADS_X3 = 32 * (2 + TC_EPI_WARPS_X3 + TC_CONV_WARPS * TC_CONV_GROUPS);

// One operand tile of one k-block, converted IN PLACE: fp32 (as TMA staged it) -> [32 bf16 hi | 32 bf16 lo] rows, K-major
// SWIZZLE_128B (16-byte chunk j of row r lives at chunk j ^ (r & 7) of the row's 128 bytes).  The converted tile takes
// exactly the bytes of the fp32 tile, so the operand ring keeps the depth of the TF32 kernel (a separate ring of
// converted tiles left room for only 3 raw stages -- too little data in flight to cover the TMA latency).  Every warp
// reads ALL the fp32 data of the rows / boxes it owns, __syncwarp()s, then overwrites them.  `cw` = converter warp 0..3.
//   K-major source : the fp32 tile has the same swizzle; lane -> (row 8 i + lane % 8, floats 8 c .. 8 c + 7, c = lane / 8):
//                    two 16-byte reads, two 16-byte writes into the same row, every quarter-warp touches 8 distinct
//                    chunk positions; the four lanes of a row sit in one warp.
//   MN-major source: un-swizzled boxes [32 k][32 mn] = the 4 KB the 32 converted rows of those mn occupy; a warp owns a
//                    whole box: lane = mn column, 32 four-byte reads down k (one 128-byte row per instruction), eight
//                    16-byte writes into row mn (8 consecutive rows per quarter-warp -> 8 distinct chunk positions).
template <int ROWS, bool MN_MAJOR>
__device__ __forceinline__ void convert_tile(uint8_t* __restrict__ tile, int cw, int lane, F2 z) {
  if (!MN_MAJOR) {
    const int x = lane & 7, c = lane >> 3;
    constexpr int IT = ROWS / 8 / TC_CONV_WARPS;       // 4 (128 rows) or 2 (64 rows)
    float4 v0[IT], v1[IT];
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      const uint32_t rowoff = (uint32_t)(8 * (cw + i * TC_CONV_WARPS) + x) * 128u;
      v0[i] = *reinterpret_cast<const float4*>(tile + rowoff + (((2 * c) ^ x) << 4));
      v1[i] = *reinterpret_cast<const float4*>(tile + rowoff + (((2 * c + 1) ^ x) << 4));
    }
    __syncwarp();    // every lane of the rows has its fp32 values before any lane overwrites the rows
#pragma unroll
    for (int i = 0; i < IT; ++i) {
      const uint32_t rowoff = (uint32_t)(8 * (cw + i * TC_CONV_WARPS) + x) * 128u;
      const float v[8] = {v0[i].x, v0[i].y, v0[i].z, v0[i].w, v1[i].x, v1[i].y, v1[i].z, v1[i].w};
      uint4 hi, lo;
      split8(v, z, hi, lo);
      *reinterpret_cast<uint4*>(tile + rowoff + ((c ^ x) << 4)) = hi;
      *reinterpret_cast<uint4*>(tile + rowoff + (((4 + c) ^ x) << 4)) = lo;
    }
  } else {
#pragma unroll 1
    for (int j = cw; j < ROWS / 32; j += TC_CONV_WARPS) {
      uint8_t* box = tile + j * (TC_BK * 128);
      const float* src = reinterpret_cast<const float*>(box) + lane;
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = src[k * 32];
      __syncwarp();
      uint8_t* row = box + (uint32_t)lane * 128u;      // converted row of mn = 32 j + lane
      const int x = lane & 7;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float w[8] = {v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3], v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]};
        uint4 hi, lo;
        split8(w, z, hi, lo);
        *reinterpret_cast<uint4*>(row + ((q ^ x) << 4)) = hi;
        *reinterpret_cast<uint4*>(row + (((4 + q) ^ x) << 4)) = lo;
      }
    }
  }
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 128-byte swizzle, version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;   // LayoutType: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

// Epilogue transpose patch: 32 rows x 32 floats, the 16-byte chunk index XOR-ed with (row & 7).  Conflict-free for the
// row-per-lane writes and the 8-lanes-per-row reads without padding: 4 KB per warp, so eight epilogue warps fit next to a
// six-stage operand ring.
__device__ __forceinline__ int patch_off(int row, int col) { return row * 32 + ((((col >> 2) ^ row) & 7) << 2); }

// Rows sub_r, sub_r + 4, ... of one transposed 32 x 32 accumulator patch: bias, feature-specialised epilogue, 16-byte stores.
template <int MASK>
__device__ __forceinline__ void epi_patch_rows(const Epilogue& ep, const float* __restrict__ patch, int sub_r, int sub_c, int m0,
                                               int M, int col, const float4* res4, const float4* elu4) {
  const float4 b4 = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
  const uint64_t sd = (MASK & 4) ? ep.seed() : 0ull;
  float* crow = ep.C + (size_t)(m0 + sub_r) * ep.ldc + col;
  float* arow = ((MASK & 1) && ep.act_out) ? ep.act_out + (size_t)(m0 + sub_r) * ep.ldact + col : nullptr;
  uint64_t didx = (uint64_t)(m0 + sub_r) * (uint64_t)ep.N + (uint64_t)col;
  const size_t cstep = (size_t)4 * ep.ldc, astep = (size_t)4 * ep.ldact;
  const uint64_t dstep = (uint64_t)4 * (uint64_t)ep.N;
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (m0 + sub_r + 4 * i < M) {
      float4 acc = *reinterpret_cast<const float4*>(patch + patch_off(sub_r + 4 * i, sub_c));
      acc.x += b4.x; acc.y += b4.y; acc.z += b4.z; acc.w += b4.w;
      const float4 o = ep.template store4_masked<MASK>(acc, crow, arow, didx, sd, res4[i], elu4[i]);
      if (MASK & 32) { csum.x += o.x; csum.y += o.y; csum.z += o.z; csum.w += o.w; }
    }
    crow += cstep;
    if ((MASK & 1) && arow) arow += astep;
    didx += dstep;
  }
  if (MASK & 32) {
    // lanes l, l + 8, l + 16, l + 24 own the same four columns (rows sub_r = 0..3 mod 4): fold them in a fixed order
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o);
      csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
      csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o);
      csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
    }
    if (sub_r == 0 && m0 < M) *reinterpret_cast<float4*>(ep.colsum + (size_t)(m0 >> 5) * ep.ldcs + col) = csum;
  }
}

// One launch serves up to TC_MAX_GROUP independent problems that share the tile configuration (grouped GEMM: the four
// weight gradients of a transformer layer run as ONE persistent kernel, so a work item's K slice is 3-4x longer and the
// split-K partials 3-4x fewer than with one launch per GEMM).  Work items are numbered problem by problem.
constexpr int TC_MAX_GROUP = 4;
struct TcProblem {
  int M, N, K;
  int k_blocks_per_split;   // K blocks (of TC_BK) handled by one split-K slice
  int splits;               // number of split-K slices (1 = none)
  int tiles_m, tiles_n;     // output tile grid
  int item_begin;           // first work item of this problem
  float* partial;           // split-K workspace or NULL
  Epilogue ep;
};
struct TcParams {
  int n_problems, n_items;
  TcProblem pr[TC_MAX_GROUP];
};
struct TcMaps {
  CUtensorMap a[TC_MAX_GROUP], b[TC_MAX_GROUP];
};
__device__ __forceinline__ int find_problem(const TcParams& p, int item) {
  int g = 0;
#pragma unroll
  for (int i = 1; i < TC_MAX_GROUP; ++i)
    if (i < p.n_problems && item >= p.pr[i].item_begin) g = i;
  return g;
}

#ifdef GB_GEMM_TRACE
// Debug build only (tools/gemm_trace.py): per-CTA clock64() stamps of the pipeline events of the first 128 k-blocks / tiles.
__device__ long long* g_gemm_trace = nullptr;
#define GB_TRACE(ev, idx)                                                                                  \
  do {                                                                                                     \
    if (g_gemm_trace && (idx) < 128u) g_gemm_trace[((size_t)blockIdx.x * 128 + (idx)) * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define GB_TRACE(ev, idx)
#endif

// CTAS = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) computes a 256 x BN tile; each CTA stages its own 128 rows
// of A but only HALF of the B tile, so a pair moves 32 KB per k-block and CTA where two independent 128 x 256 CTAs move
// 48 KB for the same FLOPs -- the kernel is bound by exactly that L2 -> SM operand traffic (profiles/r1_summary.md).
template <int BN, int CTAS = 1, bool X3 = false>
struct TcCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
  static constexpr int B_ROWS = BN / CTAS;            // rows of the B tile this CTA stages
  static constexpr int B_BYTES = B_ROWS * TC_BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // 192 KB of operand ring next to 32 KB of epilogue patches.  bf16x3 converts every stage in place (same bytes), so both
  // arithmetics share the ring depth (the TF32 kernel is bound by the latency of its operand stream: one more stage than in
  // round 1 took the step from 5.94 to 5.86 ms).
  static constexpr int STAGES = CTAS == 2 ? (BN == 256 ? 6 : 8) : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int CONV_STAGES = X3 ? STAGES : 0;   // one "converted" barrier per stage
  static constexpr int THREADS = X3 ? TC_THREADS_X3 : TC_THREADS;
  static constexpr int EPI_WARPS = X3 ? TC_EPI_WARPS_X3 : 8;
  static constexpr int EPI_LD = 32;                     // floats per staged row (16-byte chunks swizzled, see patch_off)
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_LD * 4; // one 32x32 transpose patch per epilogue warp
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static_assert(RING_BYTES + EPI_BYTES + 1024 + 512 <= 232448, "shared memory budget");
  static constexpr int SMEM = RING_BYTES + EPI_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
};

// Persistent kernel: grid = min(#work items, #SMs); work item = (split-K slice, m tile, n tile), n fastest.
// Two TMEM accumulator stages: the epilogue of item i overlaps the mainloop of item i+1.
// GROUPED = false: exactly one problem, indexed statically (its fields stay immediate constant-bank operands).
template <int BN, int TA, int TB, int CTAS, bool GROUPED, bool X3>
__global__ void __launch_bounds__(X3 ? TC_THREADS_X3 : TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  using Cfg = TcCfg<BN, CTAS, X3>;
  constexpr bool PAIR = CTAS == 2;
  constexpr int B_ROWS = Cfg::B_ROWS;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;      // 0 = leader (issues the MMAs)
  const int first_item = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int A_BYTES = Cfg::A_BYTES;
  constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CONV = Cfg::CONV_STAGES;
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (an integer round trip would lose the address
  // space: the compiler then emits GENERIC ld / st for every shared-memory access derived from it -- the converter warps
  // of the bf16x3 path ran 4x slower that way)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* epi = (float*)(smem + Cfg::RING_BYTES);
  uint64_t* full_bar = (uint64_t*)(smem + Cfg::RING_BYTES + Cfg::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;    // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;         // [2] epilogue -> MMA
  uint64_t* cfull = acc_empty + 2;            // [STAGES] converter -> MMA (bf16x3; pair: both CTAs' converters, on the leader)
  uint32_t* tmem_slot = (uint32_t*)(cfull + CONV);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.n_items;

  if (threadIdx.x == 0) pdl_trigger();   // the next kernel's CTAs may start their prologue as SMs free up
  if (warp == 0 && lane == 0) {
    for (int g = 0; g < (GROUPED ? p.n_problems : 1); ++g) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.a[g]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.b[g]) : "memory");
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], Cfg::EPI_WARPS * CTAS);   // one arrival per epilogue warp (of both CTAs of a pair, on the leader)
    }
    if (X3) {
      // pair: the converter warps of BOTH CTAs arrive on the leader's cfull (its MMAs read both shared memories)
      for (int c = 0; c < CONV; ++c) mbar_init(&cfull[c], TC_CONV_WARPS * CTAS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: 2 accumulator stages of BN fp32 columns (pair: one collective allocation, same address in both)
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers must be initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      pdl_wait();        // operands are written by the preceding kernel(s)
      uint32_t it = 0;   // running k-block counter across work items (ring position)
      for (int item = first_item; item < n_items; item += item_stride) {
        const int gi = GROUPED ? find_problem(p, item) : 0;
        const TcProblem& q = p.pr[gi];
        const CUtensorMap& map_a = maps.a[gi];
        const CUtensorMap& map_b = maps.b[gi];
        const int local = item - q.item_begin;
        const int nb = local % q.tiles_n;
        const int rest = local / q.tiles_n;
        const int mb = rest % q.tiles_m;
        const int z = rest / q.tiles_m;
        const int m0 = mb * (TC_BM * CTAS) + (int)rank * TC_BM, n0 = nb * BN + (int)rank * B_ROWS;
        const int total_kb = (q.K + TC_BK - 1) / TC_BK;
        const int kb0 = z * q.k_blocks_per_split;
        const int num_kb = min(total_kb, kb0 + q.k_blocks_per_split) - kb0;
        for (int i = 0; i < num_kb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          GB_TRACE(0, it);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k = (kb0 + i) * TC_BK;
          if (PAIR && !X3) {
            // both CTAs' bytes are counted on the LEADER's full barrier (the leader's MMA reads both shared memories)
            if (rank == 0) mbar_expect_tx(&full_bar[s], CTAS * STAGE_BYTES);
            const uint32_t fb = mapa_rank(smem_u32(&full_bar[s]), 0);
            if (TA == 0) {
              tma_load_2d_pair(&map_a, fb, sa, k, m0);
            } else {
#pragma unroll
              for (int j = 0; j < TC_BM / 32; ++j) tma_load_2d_pair(&map_a, fb, sa + j * (TC_BK * 128), m0 + j * 32, k);
            }
            if (TB == 0) {
              tma_load_2d_pair(&map_b, fb, sb, k, n0);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 32; ++j) tma_load_2d_pair(&map_b, fb, sb + j * (TC_BK * 128), n0 + j * 32, k);
            }
          } else {
            // single CTA, or a bf16x3 pair: every CTA's bytes land behind its OWN full barrier (its converter warps
            // consume them; the leader's MMA is signalled through cfull)
            mbar_expect_tx(&full_bar[s], STAGE_BYTES);
            if (TA == 0) {
              tma_load_2d(&map_a, &full_bar[s], sa, k, m0);                     // box {32 k, 128 rows}
            } else {
#pragma unroll
              for (int j = 0; j < TC_BM / 32; ++j)                              // boxes {32 rows, 32 k}
                tma_load_2d(&map_a, &full_bar[s], sa + j * (TC_BK * 128), m0 + j * 32, k);
            }
            if (TB == 0) {
              tma_load_2d(&map_b, &full_bar[s], sb, k, n0);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 32; ++j)
                tma_load_2d(&map_b, &full_bar[s], sb + j * (TC_BK * 128), n0 + j * 32, k);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair: the leader CTA only) =====================
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, majors, N >> 3, M >> 4
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)TA << 15) | ((uint32_t)TB << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TC_BM * CTAS) >> 4) << 24);
    uint32_t it = 0, lt = 0;
    for (int item = first_item; item < n_items && rank == 0; item += item_stride, ++lt) {
      const TcProblem& q = p.pr[GROUPED ? find_problem(p, item) : 0];
      const int z = (item - q.item_begin) / (q.tiles_n * q.tiles_m);
      const int total_kb = (q.K + TC_BK - 1) / TC_BK;
      const int kb0 = z * q.k_blocks_per_split;
      const int num_kb = min(total_kb, kb0 + q.k_blocks_per_split) - kb0;
      const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
      mbar_wait(&acc_empty[as], aph ^ 1);               // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int i = 0; i < num_kb; ++i, ++it) {
        if (X3) {
          // ---- bf16x3: operands come from the converter warps' ring; hi*hi + lo*hi + hi*lo per 16-wide k step
          constexpr uint32_t idesc16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)((TC_BM * CTAS) >> 4) << 24);   // D = F32, A = B = BF16, both K-major
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&cfull[s], ph);
          tc_fence_after();
          if (elect_one()) {
            GB_TRACE(3, it);
            const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
            const uint32_t sb = sa + A_BYTES;
#pragma unroll
            for (int kk = 0; kk < TC_BK / 16; ++kk) {   // UMMA_K = 16 for bf16; hi part at +0, lo part at +64 bytes of the row
              const uint64_t a_hi = make_smem_desc(sa + kk * 32, 16, 1024, 2), a_lo = make_smem_desc(sa + 64 + kk * 32, 16, 1024, 2);
              const uint64_t b_hi = make_smem_desc(sb + kk * 32, 16, 1024, 2), b_lo = make_smem_desc(sb + 64 + kk * 32, 16, 1024, 2);
              const uint32_t acc0 = (i > 0 || kk > 0) ? 1u : 0u;
              if (PAIR) {
                tc_mma_bf16_pair(d_tmem, a_hi, b_hi, idesc16, acc0);
                tc_mma_bf16_pair(d_tmem, a_lo, b_hi, idesc16, 1u);
                tc_mma_bf16_pair(d_tmem, a_hi, b_lo, idesc16, 1u);
              } else {
                tc_mma_bf16(d_tmem, a_hi, b_hi, idesc16, acc0);
                tc_mma_bf16(d_tmem, a_lo, b_hi, idesc16, 1u);
                tc_mma_bf16(d_tmem, a_hi, b_lo, idesc16, 1u);
              }
            }
            if (PAIR) {
              tc_commit_pair(&empty_bar[s], 3);                       // frees this stage in BOTH CTAs
              if (i == num_kb - 1) tc_commit_pair(&acc_full[as], 3);
            } else {
              tc_commit(&empty_bar[s]);
              if (i == num_kb - 1) tc_commit(&acc_full[as]);
            }
            GB_TRACE(4, it);
          }
          __syncwarp();
          continue;
        }
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (elect_one()) {
          GB_TRACE(3, it);
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < TC_BK / 8; ++kk) {   // UMMA_K = 8 for TF32
            // K-major : 8 rows x 128 B atoms, SBO = 1024 B between 8-row groups; k step = 32 B inside the atom
            // MN-major: atom = 4 k-rows x 128 B (32-byte swizzle base); SBO = 512 B between the two atoms of one
            //           UMMA_K = 8 step, LBO = box size between 32-wide MN chunks; k step = 1024 B
            const uint64_t da = TA == 0 ? make_smem_desc(sa + kk * 32, 16, 1024, 2) : make_smem_desc(sa + kk * 1024, TC_BK * 128, 512, 1);
            const uint64_t db = TB == 0 ? make_smem_desc(sb + kk * 32, 16, 1024, 2) : make_smem_desc(sb + kk * 1024, TC_BK * 128, 512, 1);
            if (PAIR) tc_mma_tf32_pair(d_tmem, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
            else tc_mma_tf32(d_tmem, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          }
          if (PAIR) {
            tc_commit_pair(&empty_bar[s], 3);                       // frees this stage in BOTH CTAs
            if (i == num_kb - 1) tc_commit_pair(&acc_full[as], 3);  // accumulator complete (both halves)
          } else {
            tc_commit(&empty_bar[s]);                       // frees the smem stage when these MMAs retire
            if (i == num_kb - 1) tc_commit(&acc_full[as]);  // accumulator complete
          }
        }
        __syncwarp();
      }
    }
  } else if (X3 && warp >= 2 + Cfg::EPI_WARPS) {
    // ===================== converter warps (bf16x3): fp32 stage -> [bf16 hi | bf16 lo] stage, in place =====================
    const int cw = (warp - (2 + Cfg::EPI_WARPS)) % TC_CONV_WARPS;      // position inside the group
    const int cgrp = (warp - (2 + Cfg::EPI_WARPS)) / TC_CONV_WARPS;    // group: k-blocks with it % TC_CONV_GROUPS == cgrp
    const F2 zero2 = f2(__int_as_float(p.n_items >> 30));   // 0.0f the compiler cannot fold (see split_pair2)
    const uint32_t cfull_leader = PAIR ? mapa_rank(smem_u32(&cfull[0]), 0) : 0u;
    uint32_t it = 0;
    for (int item = first_item; item < n_items; item += item_stride) {
      const TcProblem& q = p.pr[GROUPED ? find_problem(p, item) : 0];
      const int z = (item - q.item_begin) / (q.tiles_n * q.tiles_m);
      const int total_kb = (q.K + TC_BK - 1) / TC_BK;
      const int kb0 = z * q.k_blocks_per_split;
      const int num_kb = min(total_kb, kb0 + q.k_blocks_per_split) - kb0;
      for (int i = 0; i < num_kb; ++i, ++it) {
        if ((int)(it % TC_CONV_GROUPS) != cgrp) continue;
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);          // TMA bytes of this stage have landed
        if (cw == 0 && lane == 0) GB_TRACE(1, it);
        uint8_t* st = smem + s * STAGE_BYTES;
        convert_tile<TC_BM, TA != 0>(st, cw, lane, zero2);
        convert_tile<B_ROWS, TB != 0>(st + A_BYTES, cw, lane, zero2);
        // generic-proxy writes -> visible to the tensor core's operand reads.  The .shared::cta form is one FENCE.VIEW.ASYNC.S;
        // the unqualified fence also issues MEMBAR.ALL.GPU (ncu: 9 % of all stall samples, the converter warps 3x slower)
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && rank != 0) mbar_arrive_cluster(cfull_leader + (uint32_t)s * 8u);
          else mbar_arrive(&cfull[s]);
        }
        if (cw == 0 && lane == 0) GB_TRACE(2, it);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // TMEM -> registers (one row per lane, 32 columns) -> per-warp smem transpose patch -> every lane owns
    // 4 consecutive columns of a row, so all global traffic of the fused epilogue is 128-byte coalesced.
    // Two warps share each TMEM lane quarter and split the tile's columns between them.
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    constexpr int EPI_COLS = BN / (Cfg::EPI_WARPS / 4);   // columns of the tile this warp handles (TF32: half, bf16x3: all)
    const int half = (warp - 2) >> 2;                  // which slice of the tile's columns
    float* patch = epi + (warp - 2) * (32 * Cfg::EPI_LD);
    const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
    const uint32_t acc_empty_remote = PAIR ? mapa_rank(smem_u32(&acc_empty[0]), 0) : 0u;   // leader's acc_empty[0]
    pdl_wait();          // side inputs are read and C / the split-K workspace written only after the preceding grid is done
    uint32_t lt = 0;
    for (int item = first_item; item < n_items; item += item_stride, ++lt) {
      const TcProblem& pq = p.pr[GROUPED ? find_problem(p, item) : 0];
      const bool vec_ok = pq.partial ? ((pq.N & 3) == 0) : pq.ep.vec_ok();
      const int local = item - pq.item_begin;
      const int nb = local % pq.tiles_n;
      const int rest = local / pq.tiles_n;
      const int mb = rest % pq.tiles_m;
      const int z = rest / pq.tiles_m;
      const int m0 = mb * (TC_BM * CTAS) + (int)rank * TC_BM + q * 32, n0 = nb * BN + half * EPI_COLS;
      const uint32_t as = lt & 1, aph = (lt >> 1) & 1;
      const uint32_t t_addr = tmem_base + as * BN + half * EPI_COLS + ((uint32_t)(q * 32) << 16);
      const bool side_inputs = !pq.partial && vec_ok && (pq.ep.residual != nullptr || pq.ep.mul_elu_out != nullptr);
      const int fmask = pq.ep.feature_mask();      // warp-uniform: selects one specialised row loop per tile
#pragma unroll 1
      for (int c0 = 0; c0 < EPI_COLS; c0 += 32) {
        const int col = n0 + c0 + sub_c;
        // Side inputs of this 32x32 patch (residual / saved activation) are requested BEFORE the accumulator is
        // touched: 8 independent 16-byte loads per lane are in flight while the MMAs finish, the TMEM read and the
        // transpose run.  (Loading them one row at a time inside the store loop made a residual epilogue
        // latency-bound: 60 us instead of 28 us for M=14848, N=K=512.)
        float4 res4[8], elu4[8];
        if (side_inputs && col < pq.N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = m0 + sub_r + 4 * i;
            if (r < pq.M) {
              if (pq.ep.residual) res4[i] = __ldg(reinterpret_cast<const float4*>(pq.ep.residual + (size_t)r * pq.ep.ldr + col));
              if (pq.ep.mul_elu_out) elu4[i] = __ldg(reinterpret_cast<const float4*>(pq.ep.mul_elu_out + (size_t)r * pq.ep.ldm + col));
            }
          }
        }
        if (c0 == 0) {
          if (X3) mbar_wait_backoff(&acc_full[as], aph);   // the converter warps share these schedulers: do not spin next to them
          else mbar_wait(&acc_full[as], aph);
          tc_fence_after();
          if (warp == 2 && lane == 0) GB_TRACE(6, lt);
        }
        float v[32];
        tc_ld16(t_addr + (uint32_t)c0, v);
        tc_ld16(t_addr + (uint32_t)c0 + 16u, v + 16);
        if (c0 + 32 >= EPI_COLS) {   // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR && rank != 0) mbar_arrive_cluster(acc_empty_remote + as * 8u);
            else mbar_arrive(&acc_empty[as]);
          }
        }
        if (n0 + c0 >= pq.N) continue;   // warp-uniform
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(patch + patch_off(lane, 4 * j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        if (col < pq.N) {
          if (pq.partial) {
            float* dst0 = pq.partial + ((size_t)z * pq.M + m0) * pq.N + col;
            if (vec_ok) {
              float* dst = dst0 + (size_t)sub_r * pq.N;
              const size_t step = (size_t)4 * pq.N;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (m0 + sub_r + 4 * i < pq.M)
                  *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(patch + patch_off(sub_r + 4 * i, sub_c));
                dst += step;
              }
            } else {
#pragma unroll 1
              for (int i = 0; i < 8; ++i) {
                const int r = sub_r + 4 * i;
                if (m0 + r >= pq.M) break;
                const float4 acc = *reinterpret_cast<const float4*>(patch + patch_off(r, sub_c));
                float* dst = dst0 + (size_t)r * pq.N;
                const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
                for (int e = 0; e < 4; ++e)
                  if (col + e < pq.N) dst[e] = a4[e];
              }
            }
          } else if (vec_ok) {
            if (fmask == 0) epi_patch_rows<0>(pq.ep, patch, sub_r, sub_c, m0, pq.M, col, res4, elu4);
            else if (fmask == 1) epi_patch_rows<1>(pq.ep, patch, sub_r, sub_c, m0, pq.M, col, res4, elu4);
            else if (fmask == 12) epi_patch_rows<12>(pq.ep, patch, sub_r, sub_c, m0, pq.M, col, res4, elu4);
            else if (fmask == 10) epi_patch_rows<10>(pq.ep, patch, sub_r, sub_c, m0, pq.M, col, res4, elu4);
            else if (fmask == 42) epi_patch_rows<42>(pq.ep, patch, sub_r, sub_c, m0, pq.M, col, res4, elu4);
            else if (fmask == 63) epi_patch_rows<63>(pq.ep, patch, sub_r, sub_c, m0, pq.M, col, res4, elu4);
            else epi_patch_rows<31>(pq.ep, patch, sub_r, sub_c, m0, pq.M, col, res4, elu4);
          } else {
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const int r = sub_r + 4 * i;
              if (m0 + r >= pq.M) break;
              const float4 acc = *reinterpret_cast<const float4*>(patch + patch_off(r, sub_c));
              const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
              for (int e = 0; e < 4; ++e)
                if (col + e < pq.N) pq.ep.store(a4[e], m0 + r, col + e);
            }
          }
        }
        __syncwarp();
      }
      if (warp == 2 && lane == 0) GB_TRACE(7, lt);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // neither CTA may release shared / tensor memory while the pair still uses it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)));
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
    cudaGetLastError();
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch ld; box = {32 floats, box_rows}.  TF32 path: TFLOAT32 elements, 128-byte
// swizzle (32-byte atoms for MN-major operands).  bf16x3 path (`exact`): FLOAT32 elements (every bit reaches the
// converter warps), 128-byte swizzle for K-major operands, NO swizzle for MN-major ones (the converter transposes them
// with plain 4-byte column reads).
static bool make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows, bool mn_major, bool exact) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  const CUtensorMapSwizzle sw = mn_major ? (exact ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(map, exact ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int TA, int TB, int CTAS, bool GROUPED, bool X3>
static int launch(const TcMaps& maps, const TcParams& p, int grid, cudaStream_t stream) {
  using Cfg = TcCfg<BN, CTAS, X3>;
  constexpr int smem = Cfg::SMEM;
  static unsigned long long configured = 0;   // per device ordinal
  if (first_use_on_device(configured))
    GB_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, TA, TB, CTAS, GROUPED, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  static const bool use_pdl = [] { const char* e = getenv("GRAPPA_B200_PDL"); return e ? atoi(e) != 0 : true; }();
  if (use_pdl) {   // start while the preceding kernel of this stream drains (see pdl_trigger / pdl_wait in common.cuh)
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  if (CTAS > 1) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = CTAS;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  GB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, TA, TB, CTAS, GROUPED, X3>, maps, p));
  GB_CHECK_LAUNCH();
  return GB_OK;
}

// tile configurations (both arithmetics): 128 x {64,128,256} single CTA, 256 x {128,256} pairs
template <bool X3>
static int dispatch_x(int BN, bool pair, int ta, int tb, const TcMaps& maps, const TcParams& p, int grid, cudaStream_t stream) {
  int rc = GB_OK;
  if (p.n_problems > 1) {   // grouped launches exist for weight gradients only (both operands MN-major)
    if (!(ta && tb) || BN == 64) {
      set_error("gemm: grouped launch needs trans_a = trans_b = 1 and 128/256-wide tiles");
      return GB_ERR_INVALID;
    }
    if (pair) rc = BN == 256 ? launch<256, 1, 1, 2, true, X3>(maps, p, grid, stream) : launch<128, 1, 1, 2, true, X3>(maps, p, grid, stream);
    else rc = BN == 256 ? launch<256, 1, 1, 1, true, X3>(maps, p, grid, stream) : launch<128, 1, 1, 1, true, X3>(maps, p, grid, stream);
    return rc;
  }
#define GB_TC(BN_, TA_, TB_, C_) rc = launch<BN_, TA_, TB_, C_, false, X3>(maps, p, grid, stream)
#define GB_TC4(BN_, C_)                     \
  do {                                      \
    if (!ta && !tb) GB_TC(BN_, 0, 0, C_);   \
    else if (!ta && tb) GB_TC(BN_, 0, 1, C_); \
    else if (ta && !tb) GB_TC(BN_, 1, 0, C_); \
    else GB_TC(BN_, 1, 1, C_);              \
  } while (0)
  if (pair) {
    if (BN == 256) GB_TC4(256, 2);
    else GB_TC4(128, 2);
  } else if (BN == 256) {
    GB_TC4(256, 1);
  } else if (BN == 128) {
    GB_TC4(128, 1);
  } else {
    GB_TC4(64, 1);
  }
#undef GB_TC4
#undef GB_TC
  return rc;
}

static int dispatch(bool x3, int BN, bool pair, int ta, int tb, const TcMaps& maps, const TcParams& p, int grid, cudaStream_t stream) {
  return x3 ? dispatch_x<true>(BN, pair, ta, tb, maps, p, grid, stream) : dispatch_x<false>(BN, pair, ta, tb, maps, p, grid, stream);
}

int launch_splitk_reduce(const float* partial, int splits, int M, int N, const Epilogue& ep, cudaStream_t stream);

// the fused column sums need the vector epilogue (16-byte aligned pointers / pitches) and whole 32-row patches
static bool colsum_legal(const gb_gemm_args* a) {
  if (!a->colsum) return true;
  const Epilogue e = make_epilogue(a);
  return ((uintptr_t)a->colsum & 15) == 0 && (a->ld_colsum & 3) == 0 && a->ld_colsum >= a->N && (a->N & 3) == 0 &&
         ((uintptr_t)e.bias & 15) == 0 && ((uintptr_t)e.C & 15) == 0 && (e.ldc & 3) == 0 &&
         ((uintptr_t)e.residual & 15) == 0 && (!e.residual || (e.ldr & 3) == 0) &&
         ((uintptr_t)e.mul_elu_out & 15) == 0 && (!e.mul_elu_out || (e.ldm & 3) == 0) &&
         ((uintptr_t)e.act_out & 15) == 0 && (!e.act_out || (e.ldact & 3) == 0);
}

static bool tma_legal(const gb_gemm_args* a) {
  // 16-byte aligned bases and row pitches (TMA), enough work to fill a tile
  if (((uintptr_t)a->A & 15) || ((uintptr_t)a->B & 15) || (a->lda & 3) || (a->ldb & 3)) return false;
  return a->K >= 8 && a->N >= 16 && a->M >= 1;
}

static bool make_maps(const gb_gemm_args* a, int b_rows, bool x3, CUtensorMap* ma, CUtensorMap* mb) {
  bool ok = a->trans_a ? make_map(ma, a->A, a->K, a->M, a->lda, TC_BK, true, x3) : make_map(ma, a->A, a->M, a->K, a->lda, TC_BM, false, x3);
  return ok && (a->trans_b ? make_map(mb, a->B, a->K, a->N, a->ldb, TC_BK, true, x3) : make_map(mb, a->B, a->N, a->K, a->ldb, b_rows, false, x3));
}

int gemm_tcgen05(const gb_gemm_args* a, cudaStream_t stream, bool* handled) {
  *handled = false;
  const int M = a->M, N = a->N, K = a->K;
  if (!tma_legal(a) || !colsum_legal(a)) return GB_OK;
  const bool x3 = a->precision == 3;
  // tile shape.  pair = a cluster of two CTAs computes 256 x BN (cta_group::2), each staging half of B.  Measured on
  // B200 (tools/gemm_one.py): 16384 x 4096 x 4096 runs at 526 / 626 TFLOP/s with single-CTA 128 x 128 / 128 x 256 tiles
  // and at 764 TFLOP/s with 256 x 256 pair tiles (the cuBLAS-measured TF32-equivalent peak), so the pair wins whenever
  // the K loop is long enough for the steady state to matter (K >= 1024: in_proj dgrad / every weight gradient,
  // -10..-17 %); with K = 512 a tile is 16 k-blocks, fill / drain dominate and the variants tie.
  // Width without split-K: 256 when the grid still covers every SM, 64 (single CTA) when even 128-wide tiles cannot
  // fill half the machine.  With split-K available (weight gradients: small output, very long K) the SMs are filled
  // by K slices instead, so the widest tile that divides N is always the cheapest in operand traffic.
  static const int max_sms_env = [] { const char* e = getenv("GRAPPA_B200_GEMM_MAX_SMS"); return e ? atoi(e) : -1; }();   // tuning aid
  const int max_sms = max_sms_env >= 0 ? max_sms_env : a->max_sms;
  const int sms_ = max_sms > 0 && max_sms < sm_count() ? max_sms : sm_count();
  static const int forced_pair = [] { const char* e = getenv("GRAPPA_B200_GEMM_PAIR"); return e ? atoi(e) : -1; }();   // tuning aid
  const bool can_split = a->workspace != nullptr && a->colsum == nullptr && (K + TC_BK - 1) / TC_BK >= 16;
  static const int pair_min_k = [] { const char* e = getenv("GRAPPA_B200_GEMM_PAIR_MINK"); return e ? atoi(e) : 1024; }();   // tuning aids
  static const int pair_min_m = [] { const char* e = getenv("GRAPPA_B200_GEMM_PAIR_MINM"); return e ? atoi(e) : 0; }();
  // bf16x3: the converter doubles the shared-memory traffic per k-block (TMA 32 KB + converter 64 KB + three MMAs' operand
  // reads 48 KB against 64 KB for TF32) and the kernel is bound by exactly that (ncu: LSU + tensor wavefronts 71 % of the
  // shared-memory pipe); a CTA pair computes twice the FLOPs per staged byte, so pairs are used from K = 128 on.
  static const int pair_min_k_x3 = [] { const char* e = getenv("GRAPPA_B200_GEMM_PAIR_MINK_X3"); return e ? atoi(e) : 128; }();
  const int min_k = x3 ? pair_min_k_x3 : pair_min_k;
  bool pair = M > TC_BM && N >= 128 && forced_pair != 0 && ((K >= min_k && (K >= 1024 || x3 || M >= pair_min_m)) || forced_pair == 1);
  int units = 0, tiles_m = 0, BN = 128;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int bm = pair ? 2 * TC_BM : TC_BM;
    units = pair ? sms_ / 2 : sms_;          // concurrently resident tiles
    tiles_m = (M + bm - 1) / bm;
    BN = 128;
    if (can_split && ((N + 127) / 128) * tiles_m * 2 <= units) BN = (N % 256 == 0) ? 256 : (N >= 128 ? 128 : 64);
    else if (N <= 64 || ((N + 127) / 128) * tiles_m < units / 2) BN = 64;
    else if (N % 256 == 0 && (N / 256) * tiles_m >= (x3 && pair ? units / 2 : units)) BN = 256;   // bf16x3 pairs: one wave of
    // 256-wide tiles on half the machine beats two waves of 128-wide ones (half the shared-memory traffic per FLOP)
    static const int forced = [] { const char* e = getenv("GRAPPA_B200_GEMM_BN"); return e ? atoi(e) : 0; }();   // tuning aid
    if ((forced == 64 || forced == 128 || forced == 256) && (forced != 256 || N % 256 == 0)) BN = forced;
    if (pair && BN == 64) {
      if (forced_pair == 1) { BN = 128; break; }
      pair = false;   // too little work for 256-row tiles: single CTAs with 128 x 64 tiles
      continue;
    }
    break;
  }
  TcMaps maps;
  if (!make_maps(a, pair ? BN / 2 : BN, x3, &maps.a[0], &maps.b[0])) return GB_OK;   // no driver entry point: FFMA path

  TcParams p;
  p.n_problems = 1;
  TcProblem& q = p.pr[0];
  q.M = M; q.N = N; q.K = K;
  q.ep = make_epilogue(a);
  const int gx = (N + BN - 1) / BN, gy = tiles_m;
  const int total_kb = (K + TC_BK - 1) / TC_BK;
  int splits = 1;
  if (can_split && gx * gy * 2 <= units && total_kb >= 16) {
    static const int min_kb = [] { const char* e = getenv("GRAPPA_B200_GEMM_MINKB"); return e ? atoi(e) : 4; }();   // tuning aid
    splits = units / (gx * gy);
    if (splits > total_kb / min_kb) splits = total_kb / min_kb;
    long long by_ws = a->workspace_bytes / ((long long)M * N * 4);
    if (splits > by_ws) splits = (int)by_ws;
    if (splits < 1) splits = 1;
  }
  q.k_blocks_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + q.k_blocks_per_split - 1) / q.k_blocks_per_split;
  q.splits = splits;
  q.tiles_m = gy;
  q.tiles_n = gx;
  q.item_begin = 0;
  q.partial = splits > 1 ? a->workspace : nullptr;
  const int n_items = gx * gy * splits;
  p.n_items = n_items;
  const int grid = pair ? 2 * (n_items < units ? n_items : units) : (n_items < units ? n_items : units);
  int rc = dispatch(x3, BN, pair, a->trans_a, a->trans_b, maps, p, grid, stream);
  if (rc) return rc;
  if (splits > 1) {
    rc = launch_splitk_reduce(q.partial, splits, M, N, q.ep, stream);
    if (rc) return rc;
  }
  *handled = true;
  return GB_OK;
}

// Grouped launch: n <= TC_MAX_GROUP problems with the same operand layouts, all TMA-legal, one shared workspace.
// Splits are chosen so that ALL problems' work items together fill the machine exactly once (one wave of equally long
// K slices) -- the weight gradients of one transformer layer: 4 launches + 4 reduces become 1 + (0 or 1).
int launch_splitk_reduce_grouped(int n, const float* const* partial, const int* splits, const int* Ms, const int* Ns,
                                 const Epilogue* eps, cudaStream_t stream);

bool gemm_tcgen05_can_fuse_colsum(const gb_gemm_args* a) { return tma_legal(a) && colsum_legal(a); }

int gemm_tcgen05_grouped(const gb_gemm_args* list, int n, cudaStream_t stream, bool* handled) {
  *handled = false;
  if (n < 1 || n > TC_MAX_GROUP) return GB_OK;
  for (int i = 0; i < n; ++i)
    if (list[i].colsum) return GB_OK;   // fused column sums are a single-launch feature
  const int ta = list[0].trans_a, tb = list[0].trans_b;
  if (!(ta && tb)) return GB_OK;
  const bool x3 = list[0].precision == 3;
  bool all256 = true, all128 = true, big_m = true, long_k = true;
  for (int i = 0; i < n; ++i) {
    const gb_gemm_args* a = &list[i];
    if (!tma_legal(a) || a->trans_a != ta || a->trans_b != tb || (a->precision == 3) != x3) return GB_OK;
    if (a->workspace != list[0].workspace || a->workspace_bytes != list[0].workspace_bytes) return GB_OK;
    all256 = all256 && (a->N % 256 == 0);
    all128 = all128 && (a->N >= 128);
    big_m = big_m && (a->M > TC_BM);
    long_k = long_k && (a->K >= 1024);
  }
  if (!all128) return GB_OK;
  static const int forced_pair = [] { const char* e = getenv("GRAPPA_B200_GEMM_PAIR"); return e ? atoi(e) : -1; }();
  const bool pair = big_m && long_k && forced_pair != 0;
  const int BN = all256 ? 256 : 128;
  const int bm = pair ? 2 * TC_BM : TC_BM;
  // Grouped launches carry weight gradients, which run on a side stream NEXT TO the latency-critical backward chain: a
  // persistent kernel on all 148 SMs makes every small kernel of that chain wait for a free SM (14 us gaps per GNN block in
  // the CUPTI timeline).  Leaving a third of the machine to the chain is faster overall: 148 / 112 / 96 / 80 / 64 SMs ->
  // 6.17 / 6.05 / 6.02-6.06 / 6.07 / 6.08 ms per step.  GRAPPA_B200_GEMM_GROUP_SMS overrides (tuning aid).
  static const int group_sms = [] { const char* e = getenv("GRAPPA_B200_GEMM_GROUP_SMS"); return e ? atoi(e) : 96; }();
  const int sms_avail = group_sms > 0 && group_sms < sm_count() ? group_sms : sm_count();
  const int units = pair ? sms_avail / 2 : sms_avail;
  TcMaps maps;
  TcParams p;
  p.n_problems = n;
  int total_tiles = 0;
  for (int i = 0; i < n; ++i) {
    const gb_gemm_args* a = &list[i];
    if (!make_maps(a, pair ? BN / 2 : BN, x3, &maps.a[i], &maps.b[i])) return GB_OK;
    TcProblem& q = p.pr[i];
    q.M = a->M; q.N = a->N; q.K = a->K;
    q.ep = make_epilogue(a);
    q.tiles_m = (a->M + bm - 1) / bm;
    q.tiles_n = (a->N + BN - 1) / BN;
    total_tiles += q.tiles_m * q.tiles_n;
  }
  // slices per tile: minimise rounds x (k-blocks per item + fill/drain allowance) + reduce cost over S = 1..16
  int max_kb = 0;
  for (int i = 0; i < n; ++i) max_kb = max_kb > (list[i].K + TC_BK - 1) / TC_BK ? max_kb : (list[i].K + TC_BK - 1) / TC_BK;
  int S = 1;
  if (list[0].workspace) {
    long long best = -1;
    for (int s_try = 1; s_try <= 16; ++s_try) {
      const int kb = (max_kb + s_try - 1) / s_try;
      if (s_try > 1 && kb < 4) break;
      const long long rounds = ((long long)total_tiles * s_try + units - 1) / units;
      const long long cost = rounds * (kb + 8) + (s_try > 1 ? 4 + s_try : 0);
      if (best < 0 || cost < best) { best = cost; S = s_try; }
    }
  }
  const int kb_target = (max_kb + S - 1) / S > 4 ? (max_kb + S - 1) / S : 4;
  long long ws_used = 0;
  int items = 0, any_split = 0;
  const float* partials[TC_MAX_GROUP];
  int splits_v[TC_MAX_GROUP], Ms[TC_MAX_GROUP], Ns[TC_MAX_GROUP];
  Epilogue eps[TC_MAX_GROUP];
  for (int i = 0; i < n; ++i) {
    TcProblem& q = p.pr[i];
    const int total_kb = (q.K + TC_BK - 1) / TC_BK;
    int splits = (total_kb + kb_target - 1) / kb_target;
    const long long bytes_per_split = (long long)q.M * q.N * 4;
    if (splits > 1 && ws_used + (long long)splits * bytes_per_split > list[0].workspace_bytes) splits = 1;
    q.k_blocks_per_split = (total_kb + splits - 1) / splits;
    splits = (total_kb + q.k_blocks_per_split - 1) / q.k_blocks_per_split;
    q.splits = splits;
    q.item_begin = items;
    items += q.tiles_m * q.tiles_n * splits;
    q.partial = nullptr;
    if (splits > 1) {
      q.partial = list[0].workspace + ws_used / 4;
      ws_used += (long long)splits * bytes_per_split;
      ws_used = (ws_used + 255) / 256 * 256;
      any_split = 1;
    }
    partials[i] = q.partial; splits_v[i] = splits; Ms[i] = q.M; Ns[i] = q.N; eps[i] = q.ep;
  }
  p.n_items = items;
  const int grid = pair ? 2 * (items < units ? items : units) : (items < units ? items : units);
  int rc = dispatch(x3, BN, pair, ta, tb, maps, p, grid, stream);
  if (rc) return rc;
  if (any_split) {
    rc = launch_splitk_reduce_grouped(n, partials, splits_v, Ms, Ns, eps, stream);
    if (rc) return rc;
  }
  *handled = true;
  return GB_OK;
}

}  // namespace gb

#ifdef GB_GEMM_TRACE
extern "C" int grappa_b200_debug_set_gemm_trace(long long* buf) {
  return cudaMemcpyToSymbol(gb::g_gemm_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -1;
}
#endif
