// Host-side tuple index construction (angles, propers, impropers) from a bond list.
//
// Reproduces the ORDER of the reference's pure-Python routines bit-exactly
// (reference src/grappa/utils/tuple_indices.py:7-63 get_idx_tuples, :66-83 get_neighbor_dict,
//  :87-140 is_improper / is_proper, :144-216 get_torsions) with a different mechanism: a CSR
// adjacency with sorted rows plus a "first appearance" atom order instead of Python dicts.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <set>
#include <unordered_map>
#include <climits>
#include <cstdint>
#include <vector>

#include "../../include/grappa_b200.h"

namespace gb {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace gb

extern "C" const char* grappa_b200_last_error(void) { return gb::g_err; }
extern "C" int grappa_b200_abi_version(void) { return GB_ABI_VERSION; }

namespace {

struct Adjacency {
  std::vector<int64_t> order;                       // atoms in order of first appearance in `bonds`
  std::unordered_map<int64_t, int64_t> slot;        // atom id -> row
  std::vector<std::vector<int64_t>> nbr;            // sorted neighbour lists (duplicates kept)

  const std::vector<int64_t>& of(int64_t atom) const { return nbr[slot.at(atom)]; }
  bool has(int64_t atom) const { return slot.find(atom) != slot.end(); }
  bool bonded(int64_t a, int64_t b) const {
    const auto& r = of(a);
    return std::binary_search(r.begin(), r.end(), b);
  }
};

int build_adjacency(const int64_t* bonds, int64_t n_bonds, Adjacency& adj) {
  for (int64_t b = 0; b < n_bonds; ++b) {
    for (int i = 0; i < 2; ++i) {
      int64_t a = bonds[2 * b + i], o = bonds[2 * b + 1 - i];
      if (a == o) {
        gb::set_error("Encountered self-bond: (%lld, %lld)", (long long)a, (long long)o);
        return GB_ERR_INVALID;
      }
      auto it = adj.slot.find(a);
      if (it == adj.slot.end()) {
        adj.slot.emplace(a, (int64_t)adj.order.size());
        adj.order.push_back(a);
        adj.nbr.emplace_back();
        adj.nbr.back().push_back(o);
      } else {
        adj.nbr[it->second].push_back(o);
      }
    }
  }
  for (auto& r : adj.nbr) std::sort(r.begin(), r.end());
  return GB_OK;
}

// Visits angles and propers in the reference's emission order.
template <class FA, class FP>
void walk(const Adjacency& adj, FA&& on_angle, FP&& on_proper) {
  for (int64_t a1 : adj.order) {
    for (int64_t a2 : adj.of(a1)) {
      for (int64_t a3 : adj.of(a2)) {
        if (a3 == a1) continue;
        if (a1 < a3) on_angle(a1, a2, a3);
        for (int64_t a4 : adj.of(a3)) {
          if (a4 >= a1) break;      // rows are sorted: enforces proper[0] < proper[3]
          if (a4 == a2) continue;
          on_proper(a4, a3, a2, a1);
        }
      }
    }
  }
}

}  // namespace

extern "C" int grappa_b200_tuples_count(const int64_t* bonds, int64_t n_bonds, int64_t* n_angles,
                                        int64_t* n_propers) {
  Adjacency adj;
  int rc = build_adjacency(bonds, n_bonds, adj);
  if (rc) return rc;
  int64_t na = 0, np = 0;
  walk(adj, [&](int64_t, int64_t, int64_t) { ++na; }, [&](int64_t, int64_t, int64_t, int64_t) { ++np; });
  *n_angles = na;
  *n_propers = np;
  return GB_OK;
}

extern "C" int grappa_b200_tuples_build(const int64_t* bonds, int64_t n_bonds, int64_t* bonds_sorted,
                                        int64_t* angles, int64_t* propers) {
  Adjacency adj;
  int rc = build_adjacency(bonds, n_bonds, adj);
  if (rc) return rc;
  for (int64_t b = 0; b < n_bonds; ++b) {
    int64_t u = bonds[2 * b], v = bonds[2 * b + 1];
    bonds_sorted[2 * b] = std::min(u, v);
    bonds_sorted[2 * b + 1] = std::max(u, v);
  }
  int64_t na = 0, np = 0;
  walk(
      adj,
      [&](int64_t a, int64_t b, int64_t c) {
        angles[3 * na] = a; angles[3 * na + 1] = b; angles[3 * na + 2] = c; ++na;
      },
      [&](int64_t a, int64_t b, int64_t c, int64_t d) {
        propers[4 * np] = a; propers[4 * np + 1] = b; propers[4 * np + 2] = c; propers[4 * np + 3] = d; ++np;
      });
  return GB_OK;
}

extern "C" int grappa_b200_torsions_classify(const int64_t* bonds, int64_t n_bonds, const int64_t* torsions,
                                             int64_t n_torsions, int central_pos, int64_t* propers_out,
                                             int64_t* n_propers_out, int64_t* impropers_out,
                                             int64_t* n_impropers_out) {
  if (central_pos < 0 || central_pos > 3) {
    gb::set_error("central_pos must be in [0,3], got %d", central_pos);
    return GB_ERR_INVALID;
  }
  Adjacency adj;
  int rc = build_adjacency(bonds, n_bonds, adj);
  if (rc) return rc;
  std::set<std::array<int64_t, 4>> seen;   // sorted atom sets already emitted (proper or improper)
  int64_t np = 0, ni = 0;
  static const int probe[4] = {2, 1, 0, 3};  // central-atom candidates in the reference's order
  for (int64_t t = 0; t < n_torsions; ++t) {
    std::array<int64_t, 4> ids = {torsions[4 * t], torsions[4 * t + 1], torsions[4 * t + 2], torsions[4 * t + 3]};
    std::array<int64_t, 4> key = ids;
    std::sort(key.begin(), key.end());
    if (seen.count(key)) continue;
    for (int i = 0; i < 4; ++i) {
      if (!adj.has(ids[i])) {
        gb::set_error("torsion %lld references atom %lld that has no bond", (long long)t, (long long)ids[i]);
        return GB_ERR_INVALID;
      }
    }
    int central = -1;
    for (int p : probe) {
      int64_t c = ids[p];
      bool all = true;
      for (int i = 0; i < 4; ++i)
        if (ids[i] != c && !adj.bonded(c, ids[i])) { all = false; break; }
      if (all) {
        // position of the first occurrence of the central atom (python tuple.index)
        for (int i = 0; i < 4; ++i)
          if (ids[i] == c) { central = i; break; }
        break;
      }
    }
    bool is_proper = adj.bonded(ids[1], ids[0]) && adj.bonded(ids[2], ids[1]) && adj.bonded(ids[3], ids[2]);
    bool is_improper = central >= 0 && !is_proper;   // both -> proper (tuple_indices.py:168-171)
    if (!is_proper && !is_improper) {
      gb::set_error("Encountered torsion that is neither proper nor improper: (%lld, %lld, %lld, %lld)",
                    (long long)ids[0], (long long)ids[1], (long long)ids[2], (long long)ids[3]);
      return GB_ERR_INVALID;
    }
    if (!is_improper) {
      for (int i = 0; i < 4; ++i) propers_out[4 * np + i] = ids[i];
      ++np;
      seen.insert(key);
    } else {
      int64_t others[3];
      int j = 0;
      for (int i = 0; i < 4; ++i)
        if (i != central) others[j++] = ids[i];
      static const int rot[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}};
      for (int v = 0; v < 3; ++v) {
        int o = 0;
        for (int pos = 0; pos < 4; ++pos) {
          impropers_out[4 * ni + pos] = (pos == central_pos) ? ids[central] : others[rot[v][o++]];
        }
        ++ni;
      }
      seen.insert(key);
    }
  }
  *n_propers_out = np;
  *n_impropers_out = ni;
  return GB_OK;
}


// ---------------------------------------------------------------------------------------------
// Conflict-free rounds for the energy kernel (see include/grappa_b200.h)
// ---------------------------------------------------------------------------------------------
// Ring-membership features without rdkit: for every bond (a, b) the shortest cycle through it is found by a breadth-first
// search from a to b that may not use the bond itself (depth <= 7, i.e. rings of <= 8 atoms); every atom on that cycle is
// flagged "in a ring" (column 0) and "in a ring of that size" (column size - 2).  For the ring systems force fields see
// (isolated and fused 3..8-rings) this marks the same atoms as rdkit's IsInRing / IsInRingSize(3..8), which the
// reference reads (utils/rdkit_utils.py:7-24).  Neighbours are visited in bond-list order, so the result does not
// depend on anything but the bond list.
extern "C" int grappa_b200_ring_encoding(int64_t n_atoms, const int64_t* bonds, int64_t n_bonds, float* enc) {
  if (n_atoms < 0 || n_bonds < 0 || (n_bonds > 0 && !bonds) || (n_atoms > 0 && !enc)) {
    gb::set_error("ring_encoding: bad arguments (n_atoms=%lld n_bonds=%lld)", (long long)n_atoms, (long long)n_bonds);
    return GB_ERR_INVALID;
  }
  for (int64_t i = 0; i < n_atoms * 7; ++i) enc[i] = 0.f;
  // CSR adjacency in bond-list order
  std::vector<int64_t> ptr((size_t)n_atoms + 1, 0);
  for (int64_t e = 0; e < 2 * n_bonds; ++e) {
    if (bonds[e] < 0 || bonds[e] >= n_atoms) {
      gb::set_error("ring_encoding: bond %lld references atom %lld outside [0, %lld)", (long long)(e / 2),
                    (long long)bonds[e], (long long)n_atoms);
      return GB_ERR_INVALID;
    }
    ++ptr[(size_t)bonds[e] + 1];
  }
  for (int64_t i = 0; i < n_atoms; ++i) ptr[(size_t)i + 1] += ptr[(size_t)i];
  std::vector<int64_t> nbr((size_t)(2 * n_bonds)), fillp(ptr.begin(), ptr.end() - 1);
  for (int64_t e = 0; e < n_bonds; ++e) {
    const int64_t a = bonds[2 * e], b = bonds[2 * e + 1];
    nbr[(size_t)fillp[(size_t)a]++] = b;
    nbr[(size_t)fillp[(size_t)b]++] = a;
  }
  std::vector<int32_t> dist((size_t)n_atoms, -1);
  std::vector<int64_t> parent((size_t)n_atoms, -1), queue;
  for (int64_t e = 0; e < n_bonds; ++e) {
    const int64_t a = bonds[2 * e], b = bonds[2 * e + 1];
    queue.clear();
    queue.push_back(a);
    dist[(size_t)a] = 0;
    parent[(size_t)a] = -1;
    bool found = false;
    for (size_t h = 0; h < queue.size() && !found; ++h) {
      const int64_t u = queue[h];
      if (dist[(size_t)u] >= 7) break;
      for (int64_t j = ptr[(size_t)u]; j < ptr[(size_t)u + 1]; ++j) {
        const int64_t v = nbr[(size_t)j];
        if (u == a && v == b) continue;            // the bond itself (every parallel copy of it)
        if (dist[(size_t)v] >= 0) continue;
        dist[(size_t)v] = dist[(size_t)u] + 1;
        parent[(size_t)v] = u;
        queue.push_back(v);
        if (v == b) {
          found = true;
          break;
        }
      }
    }
    if (found) {
      const int size = dist[(size_t)b] + 1;
      if (size >= 3 && size <= 8)
        for (int64_t v = b; v != -1; v = parent[(size_t)v]) {
          enc[v * 7] = 1.f;
          enc[v * 7 + size - 2] = 1.f;
        }
    }
    for (int64_t v : queue) dist[(size_t)v] = -1;  // reset only what this search touched
  }
  // Column 0 also has to be set for atoms whose smallest ring is larger than 8 (macrocycles: rdkit's IsInRing is true
  // there): an atom lies on a cycle iff one of its bonds is not a bridge.  Bridges by one iterative depth-first search
  // (low-link values; the tree edge is skipped by edge id so that a doubled bond counts as a 2-cycle, not as a bridge).
  std::vector<int64_t> eid((size_t)(2 * n_bonds));
  {
    std::vector<int64_t> fp(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < n_bonds; ++e) {
      eid[(size_t)fp[(size_t)bonds[2 * e]]++] = e;
      eid[(size_t)fp[(size_t)bonds[2 * e + 1]]++] = e;
    }
  }
  std::vector<int64_t> tin((size_t)n_atoms, -1), low((size_t)n_atoms, 0), it((size_t)n_atoms, 0), via((size_t)n_atoms, -1);
  std::vector<char> bridge((size_t)n_bonds, 0);
  std::vector<int64_t> stack;
  int64_t timer = 0;
  for (int64_t root = 0; root < n_atoms; ++root) {
    if (tin[(size_t)root] >= 0) continue;
    tin[(size_t)root] = low[(size_t)root] = timer++;
    it[(size_t)root] = ptr[(size_t)root];
    stack.assign(1, root);
    while (!stack.empty()) {
      const int64_t u = stack.back();
      if (it[(size_t)u] < ptr[(size_t)u + 1]) {
        const int64_t j = it[(size_t)u]++;
        const int64_t v = nbr[(size_t)j];
        if (eid[(size_t)j] == via[(size_t)u]) continue;           // the edge we came in by
        if (tin[(size_t)v] >= 0) {
          low[(size_t)u] = std::min(low[(size_t)u], tin[(size_t)v]);
        } else {
          tin[(size_t)v] = low[(size_t)v] = timer++;
          via[(size_t)v] = eid[(size_t)j];
          it[(size_t)v] = ptr[(size_t)v];
          stack.push_back(v);
        }
      } else {
        stack.pop_back();
        if (!stack.empty()) {
          const int64_t p = stack.back();
          low[(size_t)p] = std::min(low[(size_t)p], low[(size_t)u]);
          if (low[(size_t)u] > tin[(size_t)p]) bridge[(size_t)via[(size_t)u]] = 1;
        }
      }
    }
  }
  for (int64_t e = 0; e < n_bonds; ++e)
    if (!bridge[(size_t)e] && bonds[2 * e] != bonds[2 * e + 1]) enc[bonds[2 * e] * 7] = enc[bonds[2 * e + 1] * 7] = 1.f;
  return GB_OK;
}

extern "C" int64_t grappa_b200_conflict_free_rounds(const int32_t* idx, const int32_t* tup_off, int32_t n_mols, int32_t L,
                                                    int32_t groups, int32_t* round_off, int32_t* sched,
                                                    int64_t capacity_rounds) {
  if (!tup_off || n_mols < 0 || L < 1 || L > 4 || groups < 1 || groups > 32) {
    gb::set_error("conflict_free_rounds: bad arguments (n_mols=%d L=%d groups=%d)", n_mols, L, groups);
    return GB_ERR_INVALID;
  }
  int64_t total = 0;
  // busy[a * words + w]: rounds (bit set) in which atom a of the current molecule is already used.  A molecule's atoms form
  // a contiguous index range, so a dense table replaces the hash map this started with (3.3 ms -> well under 1 ms per
  // 32-molecule batch: the schedule is built on the data-loader path, once per batch).
  std::vector<uint64_t> busy, full, tmp;                     // full: rounds that already hold `groups` tuples
  std::vector<int32_t> fill;                                 // tuples per round
  for (int32_t b = 0; b < n_mols; ++b) {
    const int32_t t0 = tup_off[b], t1 = tup_off[b + 1];
    if (t1 < t0) {
      gb::set_error("conflict_free_rounds: tup_off is not monotone at molecule %d", b);
      return GB_ERR_INVALID;
    }
    if (t1 > t0 && !idx) {
      gb::set_error("conflict_free_rounds: idx is NULL");
      return GB_ERR_INVALID;
    }
    if (round_off) round_off[b] = (int32_t)total;
    int32_t amin = INT32_MAX, amax = INT32_MIN;
    for (int64_t i = (int64_t)t0 * L; i < (int64_t)t1 * L; ++i) {
      amin = idx[i] < amin ? idx[i] : amin;
      amax = idx[i] > amax ? idx[i] : amax;
    }
    const int64_t n_at = t1 > t0 ? (int64_t)amax - amin + 1 : 0;
    // a round holds <= groups disjoint tuples, so (t1 - t0) rounds always suffice: size the bitsets for that
    const size_t words = (size_t)((t1 - t0) + 63) / 64 + 1;
    busy.assign((size_t)n_at * words, 0);
    full.assign(words, 0);
    tmp.resize(words);
    fill.clear();
    for (int32_t t = t0; t < t1; ++t) {
      for (size_t w = 0; w < words; ++w) tmp[w] = full[w];
      for (int j = 0; j < L; ++j) {
        const uint64_t* row = &busy[(size_t)(idx[(int64_t)t * L + j] - amin) * words];
        for (size_t w = 0; w < words; ++w) tmp[w] |= row[w];
      }
      int64_t r = -1;
      for (size_t w = 0; w < words && r < 0; ++w)
        if (~tmp[w]) r = (int64_t)w * 64 + __builtin_ctzll(~tmp[w]);
      if (r < 0 || r > (int64_t)fill.size()) r = (int64_t)fill.size();   // first never-used round
      if (r == (int64_t)fill.size()) fill.push_back(0);
      const size_t w = (size_t)r / 64;
      const uint64_t bit = 1ull << (r % 64);
      for (int j = 0; j < L; ++j) busy[(size_t)(idx[(int64_t)t * L + j] - amin) * words + w] |= bit;
      if (sched) {
        if (total + r >= capacity_rounds) {
          gb::set_error("conflict_free_rounds: schedule buffer too small");
          return GB_ERR_INVALID;
        }
        sched[(total + r) * groups + fill[r]] = t;
      }
      if (++fill[r] == groups) full[w] |= bit;
    }
    if (sched)
      for (size_t r = 0; r < fill.size(); ++r)
        for (int g = fill[r]; g < groups; ++g) sched[(total + (int64_t)r) * groups + g] = -1;
    total += (int64_t)fill.size();
  }
  if (round_off) round_off[n_mols] = (int32_t)total;
  return total;
}
