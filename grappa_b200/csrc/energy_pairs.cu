// K13, packed-pair kernel: MM bonded energy + analytic forces with TWO conformations per lane (f32x2 arithmetic).
//
// Replaces reference src/grappa/models/internal_coordinates.py:15-125, models/energy.py:8-71 and the autograd call at
// models/energy.py:139, like energy.cu's kernels; math in energy_math2.cuh.
//
// Why: the round-scheduled kernel of round 1 (energy_rounds_kernel) ran at 5.7 % of the HBM roofline with the issue
// slots 67 % busy -- ~2570 warp instructions per (molecule, conformation) evaluation, 78 % lane utilisation, one block
// barrier per 8 tuples.  Here
//   * a lane owns a PAIR of conformations held in one 64-bit register; subtract / cross / dot / FMA run as FADD2 / FMUL2 /
//     FFMA2 (one issue slot for both), shared-memory traffic is 64-bit per lane, indices / parameters / address
//     arithmetic are paid once per pair;
//   * a warp carries two tuple slots (half-warps of 16 lanes = 32 conformations per tile): every 64-bit shared-memory
//     access of a warp is two conflict-free 128-byte wavefronts, one per half;
//   * each half-warp adds into its OWN copy of the force tile (copy 0 / copy 1, summed at write-back), so one block
//     barrier covers TWO rounds of the host's conflict-free schedule: half the barriers, and a backbone atom that sits in
//     25 torsions no longer forces 25 serial rounds;
//   * theta = atan2(|a x b|, a.b) is a packed polynomial instead of two ~40-instruction atan2f calls.
// CTA = (molecule, tile of <= 32 conformations), 8 warps = 16 tuple slots per barrier; shared memory = xyz tile + 2 force
// tiles = 3 * 384 B per atom (60 KB for a 52-atom peptide: 3 CTAs / SM).  Forces and energies are bit-reproducible (fixed
// schedule, fixed summation order).
#include <stdlib.h>

#include "common.cuh"
#include "energy_math2.cuh"
#include "mbar.cuh"

namespace gb {

constexpr int EP_W = 32;       // conformations per tile = 16 lanes x 2
constexpr int EP_WARPS = 8;    // = sched_groups of the host schedule (tuples per conflict-free round)

// shared-memory access through 32-bit shared-window addresses (one IMAD per atom row, immediate offsets per component)
__device__ __forceinline__ F2 lds2(uint32_t addr) {
  F2 r;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r.v) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts2(uint32_t addr, F2 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v.v) : "memory"); }
constexpr uint32_t EP_ROW = EP_W * 4;          // bytes between the x / y / z rows of an atom
constexpr uint32_t EP_ATOM = 3 * EP_ROW;       // bytes per atom
__device__ __forceinline__ W3 ld_pos(uint32_t row) { return {lds2(row), lds2(row + EP_ROW), lds2(row + 2 * EP_ROW)}; }
// g += s * v   /   g -= s * v   on one atom's force row
__device__ __forceinline__ void gadd(uint32_t row, F2 s, W3 v) {
  const F2 gx = lds2(row), gy = lds2(row + EP_ROW), gz = lds2(row + 2 * EP_ROW);
  sts2(row, fma2(s, v.x, gx));
  sts2(row + EP_ROW, fma2(s, v.y, gy));
  sts2(row + 2 * EP_ROW, fma2(s, v.z, gz));
}
__device__ __forceinline__ void gsub(uint32_t row, F2 s, W3 v) {
  const F2 gx = lds2(row), gy = lds2(row + EP_ROW), gz = lds2(row + 2 * EP_ROW);
  sts2(row, gx - s * v.x);
  sts2(row + EP_ROW, gy - s * v.y);
  sts2(row + 2 * EP_ROW, gz - s * v.z);
}
// the pair (c, c + 1) of a (T, C) output row; c + 1 may lie beyond the tile
__device__ __forceinline__ void st_pair(float* p, F2 v, bool second) {
  p[0] = lo(v);
  if (second) p[1] = hi(v);
}

// One molecule's tuple records of one level, staged in shared memory by the whole CTA before the rounds start: the
// round loop then depends on ~30-cycle shared loads instead of a chain of three dependent global loads (schedule ->
// indices -> positions; 22 % of all stall samples in the first version of this kernel).
struct EpLevel {
  uint32_t sched;   // [n_rounds][8] tuple ids RELATIVE to the molecule's first tuple (or -1)
  uint32_t idx;     // [n][L] atom indices RELATIVE to the molecule's first atom (16-byte aligned rows for L = 4)
  uint32_t k;       // [n][width]
  uint32_t eq;      // [n] (bonds / angles)
  int n_rounds, t0;
};

struct EpCtx {
  uint32_t xl, gd;      // shared address of (atom 0 of the molecule, lane's pair) in the position tile; offset to this half's force tile
  uint32_t bar;         // shared address of the round mbarrier
  uint32_t it;          // rounds this warp has completed (mbarrier phase)
  int b, C, c, warp, h, lane;
  bool active, second, want_grad;
};

// words of shared memory the records of a level need (each array padded to 16 bytes)
__host__ __device__ inline int ep_level_words(int n_rounds, int n, int L, int width, bool has_eq) {
  auto pad = [](int w) { return (w + 3) & ~3; };
  return pad(n_rounds * EP_WARPS) + pad(n * L) + pad(n * width) + (has_eq ? pad(n) : 0);
}

template <int L>
__device__ __forceinline__ EpLevel stage_level(const gb_energy_args& a, int lv, int b, int a0, int width, bool has_eq, float*& cursor) {
  EpLevel e;
  const int t0 = __ldg(a.tup_off[lv] + b), n = __ldg(a.tup_off[lv] + b + 1) - t0;
  const int r0 = __ldg(a.round_off[lv] + b), nr = __ldg(a.round_off[lv] + b + 1) - r0;
  e.n_rounds = nr;
  e.t0 = t0;
  auto pad = [](int w) { return (w + 3) & ~3; };
  int32_t* s_sched = reinterpret_cast<int32_t*>(cursor);
  int32_t* s_idx = s_sched + pad(nr * EP_WARPS);
  float* s_k = reinterpret_cast<float*>(s_idx + pad(n * L));
  float* s_eq = s_k + pad(n * width);
  cursor = s_eq + (has_eq ? pad(n) : 0);
  const int tid = threadIdx.x;
  for (int i = tid; i < nr * EP_WARPS; i += 32 * EP_WARPS) {
    const int t = __ldg(a.sched[lv] + (size_t)r0 * EP_WARPS + i);
    s_sched[i] = t < 0 ? -1 : t - t0;
  }
  for (int i = tid; i < n * L; i += 32 * EP_WARPS) s_idx[i] = __ldg(a.idx[lv] + (size_t)t0 * L + i) - a0;
  for (int i = tid; i < n * width; i += 32 * EP_WARPS) s_k[i] = __ldg(a.k[lv] + (size_t)t0 * width + i);
  if (has_eq)
    for (int i = tid; i < n; i += 32 * EP_WARPS) s_eq[i] = __ldg(a.eq[lv] + t0 + i);
  e.sched = smem_u32(s_sched);
  e.idx = smem_u32(s_idx);
  e.k = smem_u32(s_k);
  e.eq = smem_u32(s_eq);
  return e;
}

__device__ __forceinline__ int lds_i(uint32_t addr) {
  int r;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ float lds_f(uint32_t addr) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
  return r;
}
__device__ __forceinline__ int4 lds_i4(uint32_t addr) {
  int4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

// Round synchronisation.  Tuples of consecutive rounds may touch the same atoms, so the force updates (read-modify-write
// of shared rows) of round r + 1 must follow those of round r -- but only the UPDATES: geometry and energies of round
// r + 1 read nothing that round r writes.  Instead of a block barrier per round (32 % of all stall samples), every warp
// ARRIVES on an mbarrier after its updates and WAITS for the previous round's phase only right before its next updates:
// the wait overlaps the ~200 instructions of geometry in between.
__device__ __forceinline__ void round_wait(EpCtx& x) {
  if (x.it > 0) {
    uint32_t done = 0, spins = 0;
    const uint32_t parity = (x.it - 1) & 1;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(x.bar), "r"(parity)
          : "memory");
      if (!done) {
        // back off instead of spinning: in the first version of this wait 40 % of all executed instructions were the
        // try_wait loop of warps whose round slot was idle, taking issue slots from the warps everyone waits for
        __nanosleep(64);
        if (++spins > (SPIN_LIMIT >> 6)) __trap();
      }
    }
  }
}
__device__ __forceinline__ void round_arrive(EpCtx& x) {
  __syncwarp();
  if (x.lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(x.bar) : "memory");
  ++x.it;
}

// One torsion level (LV = 2 propers, 3 impropers) with a compile-time periodicity (NPER = 3: grappa-1.1 / 1.2; 6: generic,
// unused amplitudes are zero).
template <bool FULL, int LV, int NPER>
__device__ __forceinline__ F2 torsion_level(const gb_energy_args& a, EpCtx& x, const EpLevel& lv) {
  const int nper = a.n_per[LV - 2];
  F2 e_acc = f2(0.f);
  uint32_t sp = lv.sched + (uint32_t)(x.h * EP_WARPS + x.warp) * 4u;
  for (int r = x.h; r < lv.n_rounds + x.h; r += 2, sp += 2 * EP_WARPS * 4) {   // both halves run the same number of iterations
    const int t = r < lv.n_rounds ? lds_i(sp) : -1;                             // uniform per half-warp
    const bool on = t >= 0 && x.active;
    uint32_t a0r = 0, a1r = 0, a2r = 0, a3r = 0;
    F2 dedphi = f2(0.f);
    TorsionGeom2 g;
    if (on) {
      const int4 id = lds_i4(lv.idx + (uint32_t)t * 16u);
      a0r = x.xl + (uint32_t)id.x * EP_ATOM; a1r = x.xl + (uint32_t)id.y * EP_ATOM;
      a2r = x.xl + (uint32_t)id.z * EP_ATOM; a3r = x.xl + (uint32_t)id.w * EP_ATOM;
      g = torsion_geom2(ld_pos(a0r), ld_pos(a1r), ld_pos(a2r), ld_pos(a3r));
      float kk[NPER];
      if (NPER == 3) {
        kk[0] = lds_f(lv.k + (uint32_t)t * 12u); kk[1] = lds_f(lv.k + (uint32_t)t * 12u + 4u); kk[2] = lds_f(lv.k + (uint32_t)t * 12u + 8u);
      } else {
#pragma unroll
        for (int n = 0; n < NPER; ++n) kk[n] = n < nper ? lds_f(lv.k + (uint32_t)(t * nper + n) * 4u) : 0.f;
      }
      F2 e;
      torsion_series2<NPER>(kk, g.cphi, g.sphi, e, dedphi);
      if (a.offset_torsion) {
        float off = 0.f;
#pragma unroll
        for (int n = 0; n < NPER; ++n) off += fabsf(kk[n]);
        e = e + f2(off);
      }
      e_acc = e_acc + e;
      if (FULL) {
        const size_t o = (size_t)(t + lv.t0) * x.C + x.c;
        if (a.x[LV]) st_pair(a.x[LV] + o, atan2_2(g.sphi, g.cphi), x.second);
        if (a.tuple_energy[LV]) st_pair(a.tuple_energy[LV] + o, e, x.second);
      }
    }
    round_wait(x);
    if (on && x.want_grad) {
      gadd(a0r + x.gd, dedphi, g.p0);                                     // dphi/dx0 = p0
      gsub(a3r + x.gd, dedphi, g.p3);                                     // dphi/dx3 = -p3
      const W3 m1 = {g.p0.x + g.u.x, g.p0.y + g.u.y, g.p0.z + g.u.z};
      const W3 m2 = {g.p3.x + g.u.x, g.p3.y + g.u.y, g.p3.z + g.u.z};
      gsub(a1r + x.gd, dedphi, m1);                                       // dphi/dx1 = -(p0 + u)
      gadd(a2r + x.gd, dedphi, m2);                                       // dphi/dx2 = p3 + u
    }
    round_arrive(x);
  }
  return e_acc;
}

template <bool FULL, int MINB>
__global__ void __launch_bounds__(32 * EP_WARPS, MINB) energy_pairs_kernel(const __grid_constant__ gb_energy_args a, int n_tiles, int wtile) {
  pdl_trigger();
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x / n_tiles;
  const int tile = blockIdx.x - b * n_tiles;
  const int a0 = __ldg(a.atom_off + b);
  const int n_at = __ldg(a.atom_off + b + 1) - a0;
  const int C = a.n_confs;
  const int c0 = tile * wtile;
  const int wc = min(wtile, C - c0);            // valid conformations of this tile (<= 32)
  const int tile_floats = n_at * 3 * EP_W;
  float* xs = smem;                             // [n_at][3][32] positions
  float* gs = smem + tile_floats;               // [2][n_at][3][32] force accumulators, one copy per half-warp
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = lane >> 4, pl = lane & 15;      // half-warp = tuple slot within the warp, lane's pair of conformations
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 3 * tile_floats);   // round barrier (16-byte slot), then the tuple records
  float* cursor = smem + 3 * tile_floats + 4;

  // One warp per atom row: the tile slice of an atom is 3 * wc contiguous floats.  Columns beyond wc replicate the last
  // valid conformation, so padded lanes compute ordinary finite numbers (never stored).
  for (int at = warp; at < n_at; at += EP_WARPS) {
    const float* src = a.xyz + ((size_t)(a0 + at) * C + c0) * 3;
    float* xrow = xs + at * 3 * EP_W;
    float* g0 = gs + at * 3 * EP_W;
#pragma unroll
    for (int j = lane; j < 3 * EP_W; j += 32) {
      const int cc = j / 3, comp = j - cc * 3;
      xrow[comp * EP_W + cc] = __ldg(src + min(cc, wc - 1) * 3 + comp);
      g0[comp * EP_W + cc] = 0.f;
      g0[tile_floats + comp * EP_W + cc] = 0.f;
    }
  }
  const bool on0 = (a.level_mask & 1) && a.n_tuples[0] > 0, on1 = (a.level_mask & 2) && a.n_tuples[1] > 0;
  const bool on2 = ((a.level_mask >> 2) & 1) && a.n_tuples[2] > 0, on3 = ((a.level_mask >> 3) & 1) && a.n_tuples[3] > 0;
  EpLevel l0 = {}, l1 = {}, l2 = {}, l3 = {};
  if (on0) l0 = stage_level<2>(a, 0, b, a0, 1, true, cursor);
  if (on1) l1 = stage_level<3>(a, 1, b, a0, 1, true, cursor);
  if (on2) l2 = stage_level<4>(a, 2, b, a0, a.n_per[0], false, cursor);
  if (on3) l3 = stage_level<4>(a, 3, b, a0, a.n_per[1], false, cursor);
  if (tid == 0) mbar_init(bar, EP_WARPS);
  __syncthreads();

  EpCtx x;
  x.b = b; x.C = C; x.warp = warp; x.h = h; x.lane = lane;
  x.active = 2 * pl < wc;
  x.second = 2 * pl + 1 < wc;
  x.c = c0 + 2 * pl;
  x.want_grad = a.grad != nullptr;
  x.xl = smem_u32(xs) + (uint32_t)(2 * pl) * 4u;               // + local atom index * EP_ATOM = that atom's row, lane's pair
  x.gd = (uint32_t)(tile_floats * 4) * (uint32_t)(1 + h);      // position row -> the same row of force copy h
  x.bar = smem_u32(bar);
  x.it = 0;
  F2 e_lvl[4] = {f2(0.f), f2(0.f), f2(0.f), f2(0.f)};

  // ---- bonds
  if (on0) {
    uint32_t sp = l0.sched + (uint32_t)(h * EP_WARPS + warp) * 4u;
    for (int r = h; r < l0.n_rounds + h; r += 2, sp += 2 * EP_WARPS * 4) {
      const int t = r < l0.n_rounds ? lds_i(sp) : -1;
      const bool on = t >= 0 && x.active;
      uint32_t a0r = 0, a1r = 0;
      F2 kd = f2(0.f);
      BondGeom2 g;
      if (on) {
        const int i0 = lds_i(l0.idx + (uint32_t)t * 8u), i1 = lds_i(l0.idx + (uint32_t)t * 8u + 4u);
        const float k = lds_f(l0.k + (uint32_t)t * 4u), eq = lds_f(l0.eq + (uint32_t)t * 4u);
        a0r = x.xl + (uint32_t)i0 * EP_ATOM; a1r = x.xl + (uint32_t)i1 * EP_ATOM;
        g = bond_geom2(ld_pos(a0r), ld_pos(a1r));
        const F2 d = g.r - f2(eq);
        kd = f2(k) * d;
        const F2 e = f2(0.5f) * kd * d;
        e_lvl[0] = e_lvl[0] + e;
        if (FULL) {
          const size_t o = (size_t)(t + l0.t0) * C + x.c;
          if (a.x[0]) st_pair(a.x[0] + o, g.r, x.second);
          if (a.tuple_energy[0]) st_pair(a.tuple_energy[0] + o, e, x.second);
        }
      }
      round_wait(x);
      if (on && x.want_grad) {
        gadd(a0r + x.gd, kd, g.d0);
        gsub(a1r + x.gd, kd, g.d0);
      }
      round_arrive(x);
    }
  }
  // ---- angles
  if (on1) {
    uint32_t sp = l1.sched + (uint32_t)(h * EP_WARPS + warp) * 4u;
    for (int r = h; r < l1.n_rounds + h; r += 2, sp += 2 * EP_WARPS * 4) {
      const int t = r < l1.n_rounds ? lds_i(sp) : -1;
      const bool on = t >= 0 && x.active;
      uint32_t a0r = 0, a1r = 0, a2r = 0;
      F2 kd = f2(0.f);
      AngleGeom2 g;
      if (on) {
        const uint32_t ia = l1.idx + (uint32_t)t * 12u;
        const int i0 = lds_i(ia), i1 = lds_i(ia + 4u), i2 = lds_i(ia + 8u);
        const float k = lds_f(l1.k + (uint32_t)t * 4u), eq = lds_f(l1.eq + (uint32_t)t * 4u);
        a0r = x.xl + (uint32_t)i0 * EP_ATOM; a1r = x.xl + (uint32_t)i1 * EP_ATOM; a2r = x.xl + (uint32_t)i2 * EP_ATOM;
        g = angle_geom2(ld_pos(a0r), ld_pos(a1r), ld_pos(a2r));
        const F2 d = g.theta - f2(eq);
        kd = f2(k) * d;
        const F2 e = f2(0.5f) * kd * d;
        e_lvl[1] = e_lvl[1] + e;
        if (FULL) {
          const size_t o = (size_t)(t + l1.t0) * C + x.c;
          if (a.x[1]) st_pair(a.x[1] + o, g.theta, x.second);
          if (a.tuple_energy[1]) st_pair(a.tuple_energy[1] + o, e, x.second);
        }
      }
      round_wait(x);
      if (on && x.want_grad) {
        gadd(a0r + x.gd, kd, g.d0);
        gadd(a2r + x.gd, kd, g.d2);
        const W3 dm = {g.d0.x + g.d2.x, g.d0.y + g.d2.y, g.d0.z + g.d2.z};
        gsub(a1r + x.gd, kd, dm);
      }
      round_arrive(x);
    }
  }
  // ---- torsions
  if (on2) e_lvl[2] = a.n_per[0] == 3 ? torsion_level<FULL, 2, 3>(a, x, l2) : torsion_level<FULL, 2, GB_MAX_PERIODICITY>(a, x, l2);
  if (on3) e_lvl[3] = a.n_per[1] == 3 ? torsion_level<FULL, 3, 3>(a, x, l3) : torsion_level<FULL, 3, GB_MAX_PERIODICITY>(a, x, l3);
  __syncthreads();

  // forces back to global, coalesced: the two copies are summed in a fixed order
  if (x.want_grad) {
    for (int at = warp; at < n_at; at += EP_WARPS) {
      float* dst = a.grad + ((size_t)(a0 + at) * C + c0) * 3;
      const float* g0 = gs + at * 3 * EP_W;
#pragma unroll
      for (int j = lane; j < 3 * EP_W; j += 32) {
        const int cc = j / 3, comp = j - cc * 3;
        if (cc < wc) dst[j] = g0[comp * EP_W + cc] + g0[tile_floats + comp * EP_W + cc];
      }
    }
  }
  __syncthreads();
  // per-level energies: 16 tuple slots per conformation, folded in slot order
  float* red = smem;   // [4][16][32] floats (the host sizes shared memory for it)
  const int slot = warp * 2 + h;
#pragma unroll
  for (int lv = 0; lv < 4; ++lv)
    *reinterpret_cast<unsigned long long*>(red + (lv * 16 + slot) * EP_W + 2 * pl) = x.active ? e_lvl[lv].v : 0ull;
  __syncthreads();
  if (tid < wc) {
    float tot = 0.f;
#pragma unroll
    for (int lv = 0; lv < 4; ++lv) {
      float s2 = 0.f;
#pragma unroll
      for (int g2 = 0; g2 < 16; ++g2) s2 += red[(lv * 16 + g2) * EP_W + tid];
      if (a.term_energy[lv]) a.term_energy[lv][(size_t)b * C + c0 + tid] = s2;
      tot += s2;
    }
    if (a.energy) a.energy[(size_t)b * C + c0 + tid] = tot;
  }
}

// handled = false: the pack carries no schedule / per-molecule maxima, or the molecule's tiles + records do not fit
int launch_energy_pairs(const gb_energy_args* a, cudaStream_t stream, bool* handled) {
  *handled = false;
  const int C = a->n_confs, B = a->n_mols;
  bool have = a->sched_groups == EP_WARPS;
  int rec_words = 4;   // the round barrier
  const int Ls[4] = {2, 3, 4, 4};
  for (int l = 0; l < 4 && have; ++l) {
    if (!(a->n_tuples[l] > 0 && ((a->level_mask >> l) & 1))) continue;
    have = a->sched[l] && a->round_off[l] && a->max_tuples_per_mol[l] > 0 && a->max_rounds_per_mol[l] > 0;
    const int width = l < 2 ? 1 : a->n_per[l - 2];
    rec_words += ep_level_words(a->max_rounds_per_mol[l], a->max_tuples_per_mol[l], Ls[l], width, l < 2);
  }
  size_t smem = (size_t)a->max_atoms_per_mol * 3 * EP_W * 3 * sizeof(float) + (size_t)rec_words * 4;
  if (smem < (size_t)4 * 16 * EP_W * sizeof(float)) smem = (size_t)4 * 16 * EP_W * sizeof(float);
  if (!have || a->max_atoms_per_mol <= 0 || smem > 200 * 1024) return GB_OK;
  bool full = false;
  for (int l = 0; l < 4; ++l) full = full || a->x[l] || a->tuple_energy[l];
  const int nt = (C + EP_W - 1) / EP_W;
  int wtile = (C + nt - 1) / nt;             // balanced tiles, an even number of conformations each (pairs)
  wtile = (wtile + 1) & ~1;
  static const int minb = [] { const char* e = getenv("GRAPPA_B200_ENERGY_MINB"); return e ? atoi(e) : 3; }();   // tuning aid
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    GB_CHECK_CUDA(cudaFuncSetAttribute(energy_pairs_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    GB_CHECK_CUDA(cudaFuncSetAttribute(energy_pairs_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    GB_CHECK_CUDA(cudaFuncSetAttribute(energy_pairs_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    GB_CHECK_CUDA(cudaFuncSetAttribute(energy_pairs_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const dim3 grid((unsigned)(B * nt));
  if (full) {
    if (minb == 2) energy_pairs_kernel<true, 2><<<grid, 32 * EP_WARPS, smem, stream>>>(*a, nt, wtile);
    else energy_pairs_kernel<true, 3><<<grid, 32 * EP_WARPS, smem, stream>>>(*a, nt, wtile);
  } else {
    if (minb == 2) energy_pairs_kernel<false, 2><<<grid, 32 * EP_WARPS, smem, stream>>>(*a, nt, wtile);
    else energy_pairs_kernel<false, 3><<<grid, 32 * EP_WARPS, smem, stream>>>(*a, nt, wtile);
  }
  GB_CHECK_LAUNCH();
  *handled = true;
  return GB_OK;
}

}  // namespace gb
