// K13 / K14: MM bonded energy, analytic forces and parameter gradients over conformations.
//
// Replaces reference src/grappa/models/internal_coordinates.py:15-125 (gathered geometry),
// models/energy.py:8-71 (harmonic / torsion energies, per-molecule pooling) and the autograd call at
// models/energy.py:139 (forces) -- plus the double-backward through it that training on forces needs
// (training/loss.py:64-68).  Math lives in energy_math.cuh.
//
// Forward, two kernels:
//   energy_tiled_kernel   one CTA per (molecule, tile of W conformations).  The molecule's xyz tile
//                         is staged once in shared memory as [atom][xyz][conf] (coalesced 12*W-byte
//                         rows in, bank-conflict-free columns out), W*G threads = W conformations x
//                         G tuple groups; every thread keeps its conformation for the whole kernel
//                         so per-level energies accumulate in registers, forces accumulate in a
//                         shared-memory tile and are written back once, coalesced.  No global atomics.
//   energy_global_kernel  fallback for molecules whose tile does not fit in shared memory
//                         (> ~250 atoms): threads over (tuple, conformation), global RED atomics.
// Backward: one warp per tuple, lanes stride over conformations, warp-shuffle reduction, one
// deterministic store per parameter (no atomics, no workspace).
#include "common.cuh"
#include "energy_math.cuh"

namespace gb {

static __device__ __forceinline__ V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
static __device__ __forceinline__ V3 ld3g(const float* p) { return v3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }

// molecule that owns tuple t (off has n_mols+1 monotone entries)
static __device__ __forceinline__ int find_segment(const int32_t* __restrict__ off, int n, int t) {
  int lo = 0, hi = n;  // invariant: off[lo] <= t < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= t) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------
// tiled forward
// ---------------------------------------------------------------------------------------------
template <bool ATOMIC>
static __device__ __forceinline__ void sm_add(float* p, float v) {
  if (ATOMIC) atomicAdd(p, v); else *p += v;
}

template <bool ATOMIC>
__global__ void __launch_bounds__(512) energy_tiled_kernel(gb_energy_args a, int W, int G, int n_tiles) {
  pdl_trigger();
  extern __shared__ float smem[];
  const int b = blockIdx.x / n_tiles;
  const int tile = blockIdx.x % n_tiles;
  const int a0 = a.atom_off[b];
  const int n_at = a.atom_off[b + 1] - a0;
  const int C = a.n_confs;
  const int c0 = tile * W;
  const int wc = min(W, C - c0);           // valid conformations in this tile
  float* xs = smem;                        // [n_at][3][W]
  float* gs = smem + (size_t)n_at * 3 * W; // [n_at][3][W]
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;

  // stage xyz: global row of atom = 3*C floats, tile slice = 3*wc contiguous floats
  for (int i = tid; i < n_at * 3 * W; i += nthr) {
    int at = i / (3 * W), rem = i - at * 3 * W;
    int cl = rem / 3, comp = rem - cl * 3;
    float v = 0.f;
    if (cl < wc) v = __ldg(a.xyz + ((size_t)(a0 + at) * C + c0) * 3 + rem);
    xs[(at * 3 + comp) * W + cl] = v;
    gs[(at * 3 + comp) * W + cl] = 0.f;
  }
  __syncthreads();

  const int cl = tid % W;
  const int grp = tid / W;
  const bool active = (grp < G) && (cl < wc);
  const int c = c0 + cl;
  float e_lvl[4] = {0.f, 0.f, 0.f, 0.f};
  const bool want_grad = a.grad != nullptr;

  if (active) {
#define XS(at) v3(xs[((at) * 3 + 0) * W + cl], xs[((at) * 3 + 1) * W + cl], xs[((at) * 3 + 2) * W + cl])
#define GADD(at, vec)                                       \
  do {                                                      \
    sm_add<ATOMIC>(&gs[((at) * 3 + 0) * W + cl], (vec).x);  \
    sm_add<ATOMIC>(&gs[((at) * 3 + 1) * W + cl], (vec).y);  \
    sm_add<ATOMIC>(&gs[((at) * 3 + 2) * W + cl], (vec).z);  \
  } while (0)
    // ---- bonds
    if ((a.level_mask & 1) && a.n_tuples[0] > 0) {
      const int t1 = a.tup_off[0][b + 1];
      for (int t = a.tup_off[0][b] + grp; t < t1; t += G) {
        int i0 = __ldg(a.idx[0] + 2 * t) - a0, i1 = __ldg(a.idx[0] + 2 * t + 1) - a0;
        float k = __ldg(a.k[0] + t), eq = __ldg(a.eq[0] + t);
        BondGeom g = bond_geom(XS(i0), XS(i1));
        float d = g.r - eq;
        float e = 0.5f * k * d * d;
        e_lvl[0] += e;
        if (a.x[0]) a.x[0][(size_t)t * C + c] = g.r;
        if (a.tuple_energy[0]) a.tuple_energy[0][(size_t)t * C + c] = e;
        if (want_grad) {
          V3 f = (k * d) * g.d0;
          GADD(i0, f);
          GADD(i1, v3(-f.x, -f.y, -f.z));
        }
      }
    }
    // ---- angles
    if ((a.level_mask & 2) && a.n_tuples[1] > 0) {
      const int t1 = a.tup_off[1][b + 1];
      for (int t = a.tup_off[1][b] + grp; t < t1; t += G) {
        int i0 = __ldg(a.idx[1] + 3 * t) - a0, i1 = __ldg(a.idx[1] + 3 * t + 1) - a0,
            i2 = __ldg(a.idx[1] + 3 * t + 2) - a0;
        float k = __ldg(a.k[1] + t), eq = __ldg(a.eq[1] + t);
        AngleGeom g = angle_geom(XS(i0), XS(i1), XS(i2));
        float d = g.theta - eq;
        float e = 0.5f * k * d * d;
        e_lvl[1] += e;
        if (a.x[1]) a.x[1][(size_t)t * C + c] = g.theta;
        if (a.tuple_energy[1]) a.tuple_energy[1][(size_t)t * C + c] = e;
        if (want_grad) {
          float s = k * d;
          V3 f0 = s * g.d0, f2 = s * g.d2;
          GADD(i0, f0);
          GADD(i2, f2);
          GADD(i1, v3(-f0.x - f2.x, -f0.y - f2.y, -f0.z - f2.z));
        }
      }
    }
    // ---- torsions (propers, impropers)
#pragma unroll
    for (int lv = 2; lv < 4; ++lv) {
      if (!((a.level_mask >> lv) & 1) || a.n_tuples[lv] == 0) continue;
      const int nper = a.n_per[lv - 2];
      const int t1 = a.tup_off[lv][b + 1];
      for (int t = a.tup_off[lv][b] + grp; t < t1; t += G) {
        const int32_t* ip = a.idx[lv] + 4 * t;
        int i0 = __ldg(ip) - a0, i1 = __ldg(ip + 1) - a0, i2 = __ldg(ip + 2) - a0, i3 = __ldg(ip + 3) - a0;
        float kk[GB_MAX_PERIODICITY];
#pragma unroll
        for (int n = 0; n < GB_MAX_PERIODICITY; ++n) kk[n] = n < nper ? __ldg(a.k[lv] + (size_t)t * nper + n) : 0.f;
        TorsionGeom g = torsion_geom(XS(i0), XS(i1), XS(i2), XS(i3));
        float e, dedphi;
        if (nper == 3) torsion_series<3>(kk, g.cphi, g.sphi, e, dedphi, nullptr, nullptr);
        else torsion_series_dyn(nper, kk, g.cphi, g.sphi, e, dedphi, nullptr, nullptr);
        if (a.offset_torsion) {
#pragma unroll
          for (int n = 0; n < GB_MAX_PERIODICITY; ++n) e += fabsf(kk[n]);
        }
        e_lvl[lv] += e;
        if (a.x[lv]) a.x[lv][(size_t)t * C + c] = atan2f(g.sphi, g.cphi);
        if (a.tuple_energy[lv]) a.tuple_energy[lv][(size_t)t * C + c] = e;
        if (want_grad) {
          GADD(i0, dedphi * g.d0);
          GADD(i1, dedphi * g.d1);
          GADD(i2, dedphi * g.d2);
          GADD(i3, dedphi * g.d3);
        }
      }
    }
#undef XS
#undef GADD
  }
  __syncthreads();

  // forces back to global, coalesced
  if (want_grad) {
    for (int i = tid; i < n_at * 3 * W; i += nthr) {
      int at = i / (3 * W), rem = i - at * 3 * W;
      int cl2 = rem / 3, comp = rem - cl2 * 3;
      if (cl2 < wc) a.grad[((size_t)(a0 + at) * C + c0) * 3 + rem] = gs[(at * 3 + comp) * W + cl2];
    }
  }
  __syncthreads();
  // per-level energies: reduce the G groups of each conformation through shared memory (reuse xs)
  float* red = xs;  // [4][G][W]  (fits: 4*G*W <= 4*512 floats, checked on the host)
  if (grp < G) {
#pragma unroll
    for (int lv = 0; lv < 4; ++lv) red[(lv * G + grp) * W + cl] = active ? e_lvl[lv] : 0.f;
  }
  __syncthreads();
  if (tid < wc) {
    float tot = 0.f;
#pragma unroll
    for (int lv = 0; lv < 4; ++lv) {
      float s = 0.f;
      for (int g2 = 0; g2 < G; ++g2) s += red[(lv * G + g2) * W + tid];
      if (a.term_energy[lv]) a.term_energy[lv][(size_t)b * C + c0 + tid] = s;
      tot += s;
    }
    if (a.energy) a.energy[(size_t)b * C + c0 + tid] = tot;
  }
}

// ---------------------------------------------------------------------------------------------
// round-scheduled tiled forward (default when the pack carries a conflict-free schedule)
//
// CTA = (molecule, tile of <= 32 conformations); 32 lanes = conformations, G warps = tuple slots of a round.
// The host packs the tuples of every molecule/level into rounds of <= G tuples that share no atom
// (grappa_b200_conflict_free_rounds), so inside a round the G warps update DISJOINT rows of the shared force
// tile: plain read-modify-write instead of float shared atomics (which are CAS loops in SASS), one
// __syncthreads between rounds, and a fixed summation order -> bit-reproducible forces.  The molecule's
// coordinates are staged once in shared memory as [atom][xyz][32] (bank-conflict-free columns).
// ---------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(32 * G, 5) energy_rounds_kernel(gb_energy_args a, int n_tiles, int wtile) {
  pdl_trigger();
  extern __shared__ float smem[];
  constexpr int W = 32;
  const int b = blockIdx.x / n_tiles;
  const int tile = blockIdx.x - b * n_tiles;
  const int a0 = __ldg(a.atom_off + b);
  const int n_at = __ldg(a.atom_off + b + 1) - a0;
  const int C = a.n_confs;
  const int c0 = tile * wtile;
  const int wc = min(wtile, C - c0);       // valid conformations in this tile (<= 32)
  float* xs = smem;                        // [n_at][3][W]
  float* gs = smem + (size_t)n_at * 3 * W; // [n_at][3][W]
  const int tid = threadIdx.x;
  constexpr int nthr = 32 * G;
  const int cl = tid & 31, grp = tid >> 5;

  // One warp per atom row: the tile slice of an atom is 3 * wc contiguous floats (coalesced), lane j -> (conformation j / 3,
  // component j % 3).  (A flat loop over n_at * 96 elements spent 16 % of the kernel's instructions on index arithmetic.)
  for (int at = grp; at < n_at; at += G) {
    const float* src = a.xyz + ((size_t)(a0 + at) * C + c0) * 3;
    float* xrow = xs + at * 3 * W;
    float* grow = gs + at * 3 * W;
#pragma unroll
    for (int j = cl; j < 3 * W; j += 32) {
      const int cc = j / 3, comp = j - cc * 3;
      xrow[comp * W + cc] = cc < wc ? __ldg(src + j) : 0.f;
      grow[comp * W + cc] = 0.f;
    }
  }
  __syncthreads();

  const bool active = cl < wc;
  const int c = c0 + cl;
  const bool want_grad = a.grad != nullptr;
  float e_lvl[4] = {0.f, 0.f, 0.f, 0.f};
#define XS(at) v3(xs[((at) * 3 + 0) * W + cl], xs[((at) * 3 + 1) * W + cl], xs[((at) * 3 + 2) * W + cl])
#define GADD(at, vec)                         \
  do {                                        \
    float* g_ = gs + ((at) * 3) * W + cl;     \
    g_[0] += (vec).x;                         \
    g_[W] += (vec).y;                         \
    g_[2 * W] += (vec).z;                     \
  } while (0)
  // ---- bonds
  if ((a.level_mask & 1) && a.n_tuples[0] > 0) {
    const int r1 = __ldg(a.round_off[0] + b + 1);
    const int r0 = __ldg(a.round_off[0] + b);
    int t_next = r0 < r1 ? __ldg(a.sched[0] + (size_t)r0 * G + grp) : -1;   // the next round's tuple is fetched one round ahead
    for (int r = r0; r < r1; ++r) {
      const int t = t_next;                                                 // warp-uniform
      if (r + 1 < r1) t_next = __ldg(a.sched[0] + (size_t)(r + 1) * G + grp);
      if (t >= 0 && active) {
        const int i0 = __ldg(a.idx[0] + 2 * t) - a0, i1 = __ldg(a.idx[0] + 2 * t + 1) - a0;
        const float k = __ldg(a.k[0] + t), eq = __ldg(a.eq[0] + t);
        BondGeom g = bond_geom(XS(i0), XS(i1));
        const float d = g.r - eq;
        const float e = 0.5f * k * d * d;
        e_lvl[0] += e;
        if (a.x[0]) a.x[0][(size_t)t * C + c] = g.r;
        if (a.tuple_energy[0]) a.tuple_energy[0][(size_t)t * C + c] = e;
        if (want_grad) {
          V3 f = (k * d) * g.d0;
          GADD(i0, f);
          GADD(i1, v3(-f.x, -f.y, -f.z));
        }
      }
      __syncthreads();
    }
  }
  // ---- angles
  if ((a.level_mask & 2) && a.n_tuples[1] > 0) {
    const int r1 = __ldg(a.round_off[1] + b + 1);
    const int r0 = __ldg(a.round_off[1] + b);
    int t_next = r0 < r1 ? __ldg(a.sched[1] + (size_t)r0 * G + grp) : -1;
    for (int r = r0; r < r1; ++r) {
      const int t = t_next;
      if (r + 1 < r1) t_next = __ldg(a.sched[1] + (size_t)(r + 1) * G + grp);
      if (t >= 0 && active) {
        const int i0 = __ldg(a.idx[1] + 3 * t) - a0, i1 = __ldg(a.idx[1] + 3 * t + 1) - a0,
                  i2 = __ldg(a.idx[1] + 3 * t + 2) - a0;
        const float k = __ldg(a.k[1] + t), eq = __ldg(a.eq[1] + t);
        AngleGeom g = angle_geom(XS(i0), XS(i1), XS(i2));
        const float d = g.theta - eq;
        const float e = 0.5f * k * d * d;
        e_lvl[1] += e;
        if (a.x[1]) a.x[1][(size_t)t * C + c] = g.theta;
        if (a.tuple_energy[1]) a.tuple_energy[1][(size_t)t * C + c] = e;
        if (want_grad) {
          const float s = k * d;
          V3 f0 = s * g.d0, f2 = s * g.d2;
          GADD(i0, f0);
          GADD(i2, f2);
          GADD(i1, v3(-f0.x - f2.x, -f0.y - f2.y, -f0.z - f2.z));
        }
      }
      __syncthreads();
    }
  }
  // ---- torsions (propers, then impropers): one instantiation of the loop body for both levels
#pragma unroll 1
  for (int lv = 2; lv < 4; ++lv) {
    if (!((a.level_mask >> lv) & 1) || a.n_tuples[lv] == 0) continue;   // uniform
    const int nper = a.n_per[lv - 2];
    const int32_t* __restrict__ sched = a.sched[lv];
    const int32_t* __restrict__ idx = a.idx[lv];
    const float* __restrict__ kp = a.k[lv];
    float* __restrict__ xo = a.x[lv];
    float* __restrict__ teo = a.tuple_energy[lv];
    const int r1 = __ldg(a.round_off[lv] + b + 1);
    float e_acc = 0.f;
    const int r0 = __ldg(a.round_off[lv] + b);
    int t_next = r0 < r1 ? __ldg(sched + (size_t)r0 * G + grp) : -1;
    for (int r = r0; r < r1; ++r) {
      const int t = t_next;
      if (r + 1 < r1) t_next = __ldg(sched + (size_t)(r + 1) * G + grp);
      if (t >= 0 && active) {
        const int32_t* ip = idx + 4 * t;
        const int i0 = __ldg(ip) - a0, i1 = __ldg(ip + 1) - a0, i2 = __ldg(ip + 2) - a0, i3 = __ldg(ip + 3) - a0;
        TorsionGeom g = torsion_geom(XS(i0), XS(i1), XS(i2), XS(i3));
        float e, dedphi;
        float kk[GB_MAX_PERIODICITY];
        if (nper == 3) {
          kk[0] = __ldg(kp + 3 * t); kk[1] = __ldg(kp + 3 * t + 1); kk[2] = __ldg(kp + 3 * t + 2);
          kk[3] = kk[4] = kk[5] = 0.f;
          torsion_series<3>(kk, g.cphi, g.sphi, e, dedphi, nullptr, nullptr);
        } else {
#pragma unroll
          for (int n = 0; n < GB_MAX_PERIODICITY; ++n) kk[n] = n < nper ? __ldg(kp + (size_t)t * nper + n) : 0.f;
          torsion_series<GB_MAX_PERIODICITY>(kk, g.cphi, g.sphi, e, dedphi, nullptr, nullptr);
        }
        if (a.offset_torsion) {
#pragma unroll
          for (int n = 0; n < GB_MAX_PERIODICITY; ++n) e += fabsf(kk[n]);
        }
        e_acc += e;
        if (xo) xo[(size_t)t * C + c] = atan2f(g.sphi, g.cphi);
        if (teo) teo[(size_t)t * C + c] = e;
        if (want_grad) {
          GADD(i0, dedphi * g.d0);
          GADD(i1, dedphi * g.d1);
          GADD(i2, dedphi * g.d2);
          GADD(i3, dedphi * g.d3);
        }
      }
      __syncthreads();
    }
    if (lv == 2) e_lvl[2] = e_acc; else e_lvl[3] = e_acc;
  }
#undef XS
#undef GADD
  // forces back to global, coalesced (the last round ended with a barrier)
  if (want_grad) {
    for (int at = grp; at < n_at; at += G) {
      float* dst = a.grad + ((size_t)(a0 + at) * C + c0) * 3;
      const float* grow = gs + at * 3 * W;
#pragma unroll
      for (int j = cl; j < 3 * W; j += 32) {
        const int cc = j / 3, comp = j - cc * 3;
        if (cc < wc) dst[j] = grow[comp * W + cc];
      }
    }
  }
  __syncthreads();
  float* red = xs;  // [4][G][W]
#pragma unroll
  for (int lv = 0; lv < 4; ++lv) red[(lv * G + grp) * W + cl] = active ? e_lvl[lv] : 0.f;
  __syncthreads();
  if (tid < wc) {
    float tot = 0.f;
#pragma unroll
    for (int lv = 0; lv < 4; ++lv) {
      float s2 = 0.f;
#pragma unroll
      for (int g2 = 0; g2 < G; ++g2) s2 += red[(lv * G + g2) * W + tid];
      if (a.term_energy[lv]) a.term_energy[lv][(size_t)b * C + c0 + tid] = s2;
      tot += s2;
    }
    if (a.energy) a.energy[(size_t)b * C + c0 + tid] = tot;
  }
}

// ---------------------------------------------------------------------------------------------
// conformation-per-thread forward (default for molecules up to a few hundred atoms)
//
// One CTA = (molecule, tile of blockDim.x conformations); a thread owns ONE conformation and walks the
// molecule's whole tuple list, so
//   * the force accumulator column gs[atom][xyz][thread] in shared memory is private to the thread: plain
//     read-modify-write, no atomics (float shared atomics are CAS loops in SASS), no __syncthreads, and the
//     summation order is the tuple order -> bit-reproducible forces and energies;
//   * per-level energies never leave registers (no cross-thread reduction at all);
//   * tuple indices / parameters are warp-uniform loads (one transaction per warp);
//   * coordinates come straight from global memory through L1 (lanes = consecutive conformations of one atom =
//     384 contiguous bytes per warp); runs of consecutive tuples that share their central atom (angles) or
//     central bond (torsions) -- the order the reference's tuple builder emits -- keep those positions and
//     force accumulators in registers, halving the L1/shared traffic of the torsion loop.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) energy_conf_kernel(gb_energy_args a, int n_tiles) {
  pdl_trigger();
  extern __shared__ float gs[];            // [n_at * 3][W]
  const int W = blockDim.x, tid = threadIdx.x;
  const int b = blockIdx.x / n_tiles;
  const int tile = blockIdx.x - b * n_tiles;
  const int a0 = __ldg(a.atom_off + b);
  const int n_at = __ldg(a.atom_off + b + 1) - a0;
  const int C = a.n_confs;
  const int c = tile * W + tid;
  if (c >= C) return;                      // threads are independent: no barrier below
  const bool want_grad = a.grad != nullptr;
  if (want_grad)
    for (int i = 0; i < n_at * 3; ++i) gs[i * W + tid] = 0.f;
  const float* __restrict__ xp = a.xyz + ((size_t)a0 * C + c) * 3;
  const size_t astride = (size_t)C * 3;
#define XG(i) ld3g(xp + (size_t)(i) * astride)
#define GADD(i, vec)                         \
  do {                                       \
    float* g_ = gs + ((i) * 3) * W + tid;    \
    g_[0] += (vec).x;                        \
    g_[W] += (vec).y;                        \
    g_[2 * W] += (vec).z;                    \
  } while (0)
  float e_lvl[4] = {0.f, 0.f, 0.f, 0.f};

  // ---- bonds
  if ((a.level_mask & 1) && a.n_tuples[0] > 0) {
    const int t1 = __ldg(a.tup_off[0] + b + 1);
    for (int t = __ldg(a.tup_off[0] + b); t < t1; ++t) {
      const int i0 = __ldg(a.idx[0] + 2 * t) - a0, i1 = __ldg(a.idx[0] + 2 * t + 1) - a0;
      const float k = __ldg(a.k[0] + t), eq = __ldg(a.eq[0] + t);
      BondGeom g = bond_geom(XG(i0), XG(i1));
      const float d = g.r - eq;
      const float e = 0.5f * k * d * d;
      e_lvl[0] += e;
      if (a.x[0]) a.x[0][(size_t)t * C + c] = g.r;
      if (a.tuple_energy[0]) a.tuple_energy[0][(size_t)t * C + c] = e;
      if (want_grad) {
        V3 f = (k * d) * g.d0;
        GADD(i0, f);
        GADD(i1, v3(-f.x, -f.y, -f.z));
      }
    }
  }
  // ---- angles: the central atom's position and force stay in registers across a run
  if ((a.level_mask & 2) && a.n_tuples[1] > 0) {
    const int t1 = __ldg(a.tup_off[1] + b + 1);
    int cur = -1;
    V3 xc = v3(0.f, 0.f, 0.f), fc = v3(0.f, 0.f, 0.f);
    for (int t = __ldg(a.tup_off[1] + b); t < t1; ++t) {
      const int i0 = __ldg(a.idx[1] + 3 * t) - a0, i1 = __ldg(a.idx[1] + 3 * t + 1) - a0,
                i2 = __ldg(a.idx[1] + 3 * t + 2) - a0;
      const float k = __ldg(a.k[1] + t), eq = __ldg(a.eq[1] + t);
      if (i1 != cur) {                      // warp-uniform
        if (cur >= 0 && want_grad) GADD(cur, fc);
        cur = i1;
        xc = XG(i1);
        fc = v3(0.f, 0.f, 0.f);
      }
      AngleGeom g = angle_geom(XG(i0), xc, XG(i2));
      const float d = g.theta - eq;
      const float e = 0.5f * k * d * d;
      e_lvl[1] += e;
      if (a.x[1]) a.x[1][(size_t)t * C + c] = g.theta;
      if (a.tuple_energy[1]) a.tuple_energy[1][(size_t)t * C + c] = e;
      if (want_grad) {
        const float s = k * d;
        V3 f0 = s * g.d0, f2 = s * g.d2;
        GADD(i0, f0);
        GADD(i2, f2);
        fc = v3(fc.x - f0.x - f2.x, fc.y - f0.y - f2.y, fc.z - f0.z - f2.z);
      }
    }
    if (cur >= 0 && want_grad) GADD(cur, fc);
  }
  // ---- torsions (propers, impropers): the central bond's two atoms stay in registers across a run
#pragma unroll
  for (int lv = 2; lv < 4; ++lv) {
    if (!((a.level_mask >> lv) & 1) || a.n_tuples[lv] == 0) continue;
    const int nper = a.n_per[lv - 2];
    const int t1 = __ldg(a.tup_off[lv] + b + 1);
    int c1 = -1, c2 = -1;
    V3 x1 = v3(0.f, 0.f, 0.f), x2 = x1, f1 = x1, f2 = x1;
    for (int t = __ldg(a.tup_off[lv] + b); t < t1; ++t) {
      const int32_t* ip = a.idx[lv] + 4 * t;
      const int i0 = __ldg(ip) - a0, i1 = __ldg(ip + 1) - a0, i2 = __ldg(ip + 2) - a0, i3 = __ldg(ip + 3) - a0;
      float kk[GB_MAX_PERIODICITY];
#pragma unroll
      for (int n = 0; n < GB_MAX_PERIODICITY; ++n) kk[n] = n < nper ? __ldg(a.k[lv] + (size_t)t * nper + n) : 0.f;
      if (i1 != c1 || i2 != c2) {           // warp-uniform
        if (c1 >= 0 && want_grad) { GADD(c1, f1); GADD(c2, f2); }
        c1 = i1; c2 = i2;
        x1 = XG(i1); x2 = XG(i2);
        f1 = v3(0.f, 0.f, 0.f); f2 = f1;
      }
      TorsionGeom g = torsion_geom(XG(i0), x1, x2, XG(i3));
      float e, dedphi;
      if (nper == 3) torsion_series<3>(kk, g.cphi, g.sphi, e, dedphi, nullptr, nullptr);
      else torsion_series_dyn(nper, kk, g.cphi, g.sphi, e, dedphi, nullptr, nullptr);
      if (a.offset_torsion) {
#pragma unroll
        for (int n = 0; n < GB_MAX_PERIODICITY; ++n) e += fabsf(kk[n]);
      }
      e_lvl[lv] += e;
      if (a.x[lv]) a.x[lv][(size_t)t * C + c] = atan2f(g.sphi, g.cphi);
      if (a.tuple_energy[lv]) a.tuple_energy[lv][(size_t)t * C + c] = e;
      if (want_grad) {
        GADD(i0, dedphi * g.d0);
        GADD(i3, dedphi * g.d3);
        f1 = f1 + dedphi * g.d1;
        f2 = f2 + dedphi * g.d2;
      }
    }
    if (c1 >= 0 && want_grad) { GADD(c1, f1); GADD(c2, f2); }
  }
#undef XG
#undef GADD
  if (want_grad) {
    float* gp = a.grad + ((size_t)a0 * C + c) * 3;
    for (int i = 0; i < n_at; ++i) {
      float* q = gp + (size_t)i * astride;
      q[0] = gs[(i * 3 + 0) * W + tid];
      q[1] = gs[(i * 3 + 1) * W + tid];
      q[2] = gs[(i * 3 + 2) * W + tid];
    }
  }
  float tot = 0.f;
#pragma unroll
  for (int lv = 0; lv < 4; ++lv) {
    if (a.term_energy[lv]) a.term_energy[lv][(size_t)b * C + c] = e_lvl[lv];
    tot += e_lvl[lv];
  }
  if (a.energy) a.energy[(size_t)b * C + c] = tot;
}

// ---------------------------------------------------------------------------------------------
// global-atomics forward (any molecule size)
// ---------------------------------------------------------------------------------------------
template <int LV>
__global__ void __launch_bounds__(256) energy_global_kernel(gb_energy_args a) {
  pdl_trigger();
  const int C = a.n_confs;
  const long long n_items = (long long)a.n_tuples[LV] * C;
  const bool want_grad = a.grad != nullptr;
  constexpr int L = LV == 0 ? 2 : (LV == 1 ? 3 : 4);
  for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < n_items;
       it += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(it / C);
    const int c = (int)(it - (long long)t * C);
    int id[L];
#pragma unroll
    for (int j = 0; j < L; ++j) id[j] = __ldg(a.idx[LV] + (size_t)L * t + j);
    V3 p[L];
#pragma unroll
    for (int j = 0; j < L; ++j) p[j] = ld3(a.xyz + ((size_t)id[j] * C + c) * 3);
    float e, xval;
    V3 f[L];
    if (LV == 0) {
      float k = __ldg(a.k[0] + t), eq = __ldg(a.eq[0] + t);
      BondGeom g = bond_geom(p[0], p[1]);
      float d = g.r - eq;
      e = 0.5f * k * d * d;
      xval = g.r;
      f[0] = (k * d) * g.d0;
      f[1] = v3(-f[0].x, -f[0].y, -f[0].z);
    } else if (LV == 1) {
      float k = __ldg(a.k[1] + t), eq = __ldg(a.eq[1] + t);
      AngleGeom g = angle_geom(p[0], p[1], p[2]);
      float d = g.theta - eq;
      e = 0.5f * k * d * d;
      xval = g.theta;
      f[0] = (k * d) * g.d0;
      f[2] = (k * d) * g.d2;
      f[1] = v3(-f[0].x - f[2].x, -f[0].y - f[2].y, -f[0].z - f[2].z);
    } else {
      const int nper = a.n_per[LV - 2];
      float kk[GB_MAX_PERIODICITY];
#pragma unroll
      for (int n = 0; n < GB_MAX_PERIODICITY; ++n) kk[n] = n < nper ? __ldg(a.k[LV] + (size_t)t * nper + n) : 0.f;
      TorsionGeom g = torsion_geom(p[0], p[1], p[2], p[L - 1]);
      float dedphi;
      torsion_series_dyn(nper, kk, g.cphi, g.sphi, e, dedphi, nullptr, nullptr);
      if (a.offset_torsion) {
#pragma unroll
        for (int n = 0; n < GB_MAX_PERIODICITY; ++n) e += fabsf(kk[n]);
      }
      xval = (a.x[LV] != nullptr) ? atan2f(g.sphi, g.cphi) : 0.f;
      f[0] = dedphi * g.d0;
      f[1] = dedphi * g.d1;
      f[2] = dedphi * g.d2;
      f[L - 1] = dedphi * g.d3;
    }
    if (a.x[LV]) a.x[LV][(size_t)t * C + c] = xval;
    if (a.tuple_energy[LV]) a.tuple_energy[LV][(size_t)t * C + c] = e;
    const int b = find_segment(a.tup_off[LV], a.n_mols, t);
    if (a.energy) atomicAdd(a.energy + (size_t)b * C + c, e);
    if (a.term_energy[LV]) atomicAdd(a.term_energy[LV] + (size_t)b * C + c, e);
    if (want_grad) {
#pragma unroll
      for (int j = 0; j < L; ++j) {
        float* gp = a.grad + ((size_t)id[j] * C + c) * 3;
        atomicAdd(gp, f[j].x);
        atomicAdd(gp + 1, f[j].y);
        atomicAdd(gp + 2, f[j].z);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward: one warp per tuple
// ---------------------------------------------------------------------------------------------
template <int LV>
__global__ void __launch_bounds__(256) energy_bwd_kernel(gb_energy_bwd_args ba) {
  pdl_trigger();
  const gb_energy_args& a = ba.fwd;
  const int C = a.n_confs;
  constexpr int L = LV == 0 ? 2 : (LV == 1 ? 3 : 4);
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int t = blockIdx.x * warps_per_block + (threadIdx.x >> 5); t < a.n_tuples[LV];
       t += gridDim.x * warps_per_block) {
    int id[L];
#pragma unroll
    for (int j = 0; j < L; ++j) id[j] = __ldg(a.idx[LV] + (size_t)L * t + j);
    const int b = find_segment(a.tup_off[LV], a.n_mols, t);
    const bool on = (a.level_mask >> LV) & 1;
    if (LV < 2) {
      const float k = __ldg(a.k[LV] + t), eq = __ldg(a.eq[LV] + t);
      float acc_k = 0.f, acc_eq = 0.f;
      for (int c = lane; c < C && on; c += 32) {
        V3 p[L], gf[L];
#pragma unroll
        for (int j = 0; j < L; ++j) {
          p[j] = ld3(a.xyz + ((size_t)id[j] * C + c) * 3);
          gf[j] = ba.g_grad ? ld3(ba.g_grad + ((size_t)id[j] * C + c) * 3) : v3(0.f, 0.f, 0.f);
        }
        float q, s;
        if (LV == 0) {
          BondGeom g = bond_geom(p[0], p[1]);
          q = g.r;
          s = dot(gf[0] - gf[1], g.d0);
        } else {
          AngleGeom g = angle_geom(p[0], p[1], p[2]);
          q = g.theta;
          s = dot(gf[0] - gf[1], g.d0) + dot(gf[2] - gf[1], g.d2);
        }
        const float ge = ba.g_energy ? __ldg(ba.g_energy + (size_t)b * C + c) : 0.f;
        const float d = q - eq;
        acc_k += ge * 0.5f * d * d + d * s;
        acc_eq += -ge * k * d - k * s;
      }
      acc_k = warp_sum(acc_k);
      acc_eq = warp_sum(acc_eq);
      if (lane == 0) {
        if (ba.dk[LV]) ba.dk[LV][t] = acc_k;
        if (ba.deq[LV < 2 ? LV : 0]) ba.deq[LV < 2 ? LV : 0][t] = acc_eq;
      }
    } else {
      const int nper = a.n_per[LV - 2];
      float acc[GB_MAX_PERIODICITY];
#pragma unroll
      for (int n = 0; n < GB_MAX_PERIODICITY; ++n) acc[n] = 0.f;
      float zero_k[GB_MAX_PERIODICITY] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int c = lane; c < C && on; c += 32) {
        V3 p[4], gf[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          p[j] = ld3(a.xyz + ((size_t)id[j] * C + c) * 3);
          gf[j] = ba.g_grad ? ld3(ba.g_grad + ((size_t)id[j] * C + c) * 3) : v3(0.f, 0.f, 0.f);
        }
        TorsionGeom g = torsion_geom(p[0], p[1], p[2], p[3]);
        const float s = dot(gf[0], g.d0) + dot(gf[1], g.d1) + dot(gf[2], g.d2) + dot(gf[3], g.d3);
        const float ge = ba.g_energy ? __ldg(ba.g_energy + (size_t)b * C + c) : 0.f;
        float cn[GB_MAX_PERIODICITY], sn[GB_MAX_PERIODICITY], e, de;
        torsion_series<GB_MAX_PERIODICITY>(zero_k, g.cphi, g.sphi, e, de, cn, sn);
#pragma unroll
        for (int n = 0; n < GB_MAX_PERIODICITY; ++n) acc[n] += ge * cn[n] - float(n + 1) * sn[n] * s;
        if (a.offset_torsion) {
#pragma unroll
          for (int n = 0; n < GB_MAX_PERIODICITY; ++n) {
            if (n < nper) {
              float kv = __ldg(a.k[LV] + (size_t)t * nper + n);
              acc[n] += ge * (kv > 0.f ? 1.f : (kv < 0.f ? -1.f : 0.f));
            }
          }
        }
      }
#pragma unroll
      for (int n = 0; n < GB_MAX_PERIODICITY; ++n) acc[n] = warp_sum(acc[n]);
      if (lane == 0 && ba.dk[LV]) {
#pragma unroll
        for (int n = 0; n < GB_MAX_PERIODICITY; ++n)
          if (n < nper) ba.dk[LV][(size_t)t * nper + n] = acc[n];
      }
    }
  }
}

int launch_energy_pairs(const gb_energy_args* a, cudaStream_t stream, bool* handled);   // energy_pairs.cu

static int validate(const gb_energy_args* a) {
  GB_REQUIRE(a != nullptr, "energy: args is NULL");
  GB_REQUIRE(a->xyz != nullptr, "energy: xyz is NULL (xyz coordinates must be stored in g.nodes['n1'].data['xyz'])");
  GB_REQUIRE(a->n_atoms >= 0 && a->n_confs >= 0 && a->n_mols >= 0, "energy: negative size");
  GB_REQUIRE(a->atom_off != nullptr || a->n_mols == 0, "energy: atom_off is NULL");
  for (int l = 0; l < 4; ++l) {
    GB_REQUIRE(a->n_tuples[l] >= 0, "energy: negative tuple count at level %d", l);
    if (a->n_tuples[l] > 0 && ((a->level_mask >> l) & 1)) {
      GB_REQUIRE(a->idx[l] && a->tup_off[l], "energy: level %d has tuples but idx/tup_off is NULL", l);
      GB_REQUIRE(a->k[l] != nullptr, "energy: level %d has no k attribute", l);
      if (l < 2) GB_REQUIRE(a->eq[l] != nullptr, "energy: level %d has no eq attribute", l);
    }
  }
  for (int l = 0; l < 2; ++l)
    GB_REQUIRE(a->n_per[l] >= 0 && a->n_per[l] <= GB_MAX_PERIODICITY, "energy: n_periodicity %d > %d", a->n_per[l],
               GB_MAX_PERIODICITY);
  return GB_OK;
}

}  // namespace gb

using namespace gb;

extern "C" int grappa_b200_energy_fwd(const gb_energy_args* a, int variant, void* stream_) {
  int rc = validate(a);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int C = a->n_confs, B = a->n_mols;
  if (C == 0 || B == 0) return GB_OK;
  const size_t bc = (size_t)B * C * sizeof(float);
  const int mode = variant;
  const int max_atoms = a->max_atoms_per_mol;
  int W = (C + ((C + 31) / 32) - 1) / ((C + 31) / 32);  // ceil(C / ceil(C/32)) <= 32
  int n_tiles = (C + W - 1) / W;
  int G = 256 / W;
  if (G < 1) G = 1;
  if (G > 16) G = 16;
  size_t smem_floats = (size_t)max_atoms * 3 * W * 2;
  if (smem_floats < (size_t)4 * G * W) smem_floats = (size_t)4 * G * W;
  const size_t smem = smem_floats * sizeof(float);
  const bool tile_ok = max_atoms > 0 && smem <= 200 * 1024;
  GB_REQUIRE(mode != 2 || tile_ok, "energy: tiled variant requested but max_atoms=%d does not fit", max_atoms);
  // packed-pair kernel (two conformations per lane, f32x2 arithmetic): the default whenever the pack carries the
  // conflict-free schedule and the molecule's tiles fit in shared memory
  if (mode == 0 || mode == 5) {
    bool handled = false;
    rc = launch_energy_pairs(a, stream, &handled);
    if (rc) return rc;
    if (handled) return GB_OK;
    GB_REQUIRE(mode != 5, "energy: packed-pair variant needs sched/round_off with sched_groups == 8, 16-byte aligned index "
               "tables and a molecule tile that fits in shared memory (max_atoms=%d)", max_atoms);
  }
  // round-scheduled kernel: needs the host schedule and the molecule tile (xyz + forces, 32 conformations) in smem
  {
    bool have = a->sched_groups == 8;
    for (int l = 0; l < 4 && have; ++l)
      if (a->n_tuples[l] > 0 && ((a->level_mask >> l) & 1)) have = a->sched[l] && a->round_off[l];
    size_t smem_r = (size_t)max_atoms * 3 * 32 * 2 * sizeof(float);
    if (smem_r < (size_t)4 * 8 * 32 * sizeof(float)) smem_r = (size_t)4 * 8 * 32 * sizeof(float);
    const bool fits = max_atoms > 0 && smem_r <= 200 * 1024;
    GB_REQUIRE(mode != 4 || (have && fits), "energy: round-scheduled variant needs sched/round_off with sched_groups == 8 "
               "and a molecule tile that fits in shared memory (max_atoms=%d)", max_atoms);
    if (mode == 4 || (mode == 0 && have && fits)) {
      static unsigned long long configured = 0;   // per device: a process-wide flag left a second GPU unconfigured
      if (first_use_on_device(configured))
        GB_CHECK_CUDA(cudaFuncSetAttribute(energy_rounds_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      const int nt = (C + 31) / 32;
      const int wtile = (C + nt - 1) / nt;          // balanced tiles of <= 32 conformations
      energy_rounds_kernel<8><<<B * nt, 256, smem_r, stream>>>(*a, nt, wtile);
      GB_CHECK_LAUNCH();
      return GB_OK;
    }
  }
  // conformation-per-thread kernel: force columns of one tile must leave room for >= 2 CTAs per SM
  {
    const int Wc = C <= 32 ? 32 : 64;
    const size_t smem_c = (size_t)max_atoms * 3 * Wc * sizeof(float);
    const bool conf_ok = max_atoms > 0 && smem_c <= 100 * 1024;
    GB_REQUIRE(mode != 3 || conf_ok, "energy: conformation-per-thread variant requested but max_atoms=%d does not fit", max_atoms);
    if (mode == 3) {
      static unsigned long long configured = 0;
      if (first_use_on_device(configured))
        GB_CHECK_CUDA(cudaFuncSetAttribute(energy_conf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      const int nt = (C + Wc - 1) / Wc;
      energy_conf_kernel<<<B * nt, Wc, smem_c, stream>>>(*a, nt);
      GB_CHECK_LAUNCH();
      return GB_OK;
    }
  }
  const bool tiled = (mode == 2) || (mode == 0 && tile_ok);
  if (tiled) {
    auto kern = G > 1 ? energy_tiled_kernel<true> : energy_tiled_kernel<false>;
    if (smem > 48 * 1024) GB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = ((W * G + 31) / 32) * 32;
    kern<<<B * n_tiles, threads, smem, stream>>>(*a, W, G, n_tiles);
    GB_CHECK_LAUNCH();
    return GB_OK;
  }
  if (a->energy) GB_CHECK_CUDA(cudaMemsetAsync(a->energy, 0, bc, stream));
  for (int l = 0; l < 4; ++l)
    if (a->term_energy[l]) GB_CHECK_CUDA(cudaMemsetAsync(a->term_energy[l], 0, bc, stream));
  if (a->grad) GB_CHECK_CUDA(cudaMemsetAsync(a->grad, 0, (size_t)a->n_atoms * C * 3 * sizeof(float), stream));
  const int sms = sm_count();
  for (int l = 0; l < 4; ++l) {
    if (!((a->level_mask >> l) & 1) || a->n_tuples[l] == 0) continue;
    long long items = (long long)a->n_tuples[l] * C;
    int blocks = (int)((items + 255) / 256);
    if (blocks > sms * 16) blocks = sms * 16;
    switch (l) {
      case 0: energy_global_kernel<0><<<blocks, 256, 0, stream>>>(*a); break;
      case 1: energy_global_kernel<1><<<blocks, 256, 0, stream>>>(*a); break;
      case 2: energy_global_kernel<2><<<blocks, 256, 0, stream>>>(*a); break;
      default: energy_global_kernel<3><<<blocks, 256, 0, stream>>>(*a); break;
    }
    GB_CHECK_LAUNCH();
  }
  return GB_OK;
}

extern "C" int64_t grappa_b200_energy_bwd_workspace(const gb_energy_args*) { return 0; }

extern "C" int grappa_b200_energy_bwd(const gb_energy_bwd_args* ba, void* stream_) {
  GB_REQUIRE(ba != nullptr, "energy_bwd: args is NULL");
  int rc = validate(&ba->fwd);
  if (rc) return rc;
  cudaStream_t stream = (cudaStream_t)stream_;
  const gb_energy_args* a = &ba->fwd;
  if (a->n_mols == 0) return GB_OK;
  const int sms = sm_count();
  for (int l = 0; l < 4; ++l) {
    if (a->n_tuples[l] == 0 || ba->dk[l] == nullptr) continue;
    GB_REQUIRE(a->idx[l] && a->tup_off[l] && a->k[l], "energy_bwd: level %d inputs missing", l);
    int blocks = (a->n_tuples[l] + 7) / 8;
    if (blocks > sms * 8) blocks = sms * 8;
    switch (l) {
      case 0: energy_bwd_kernel<0><<<blocks, 256, 0, stream>>>(*ba); break;
      case 1: energy_bwd_kernel<1><<<blocks, 256, 0, stream>>>(*ba); break;
      case 2: energy_bwd_kernel<2><<<blocks, 256, 0, stream>>>(*ba); break;
      default: energy_bwd_kernel<3><<<blocks, 256, 0, stream>>>(*ba); break;
    }
    GB_CHECK_LAUNCH();
  }
  return GB_OK;
}
