// Host-side (CPU, C++) bookkeeping of the dynamic 3D token memory: the integer state the reference keeps in Python dicts
// (patch -> instance map, instance -> patch lists, zone keys / ids / member lists; feature_fields.py:164-177) and the
// per-view planner that turns the device results of a view (centroids, K-NN proposals, merge logits) into the packed index
// arrays the pooling kernels consume.  Pure host code behind the same C ABI (host pointers, suffix-free here: EVERYTHING in this
// file is host memory).  Mirrors FF:362-393 (cull), FF:433-475 (lowest free ids), FF:623-756 / 759-812 (update) literally,
// including quirks Q2, Q3, Q5, Q9 of SURVEY.md.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace {

typedef long long i64;
const i64 VOFF = 1 << 20, VM = 1 << 21;

// insertion-ordered map<id, list> with Python-dict semantics (assignment keeps the slot, delete + insert moves to the end)
struct OMap {
  std::vector<i64> ids;
  std::vector<std::vector<i64>> vals;
  std::vector<unsigned char> live;
  std::unordered_map<i64, size_t> pos;
  size_t n_live = 0;
  void clear() { ids.clear(); vals.clear(); live.clear(); pos.clear(); n_live = 0; }
  bool has(i64 id) const { return pos.find(id) != pos.end(); }
  std::vector<i64>& at(i64 id) { return vals[pos[id]]; }
  void assign(i64 id, std::vector<i64>&& v) {
    auto it = pos.find(id);
    if (it != pos.end()) { vals[it->second] = std::move(v); return; }
    pos[id] = ids.size();
    ids.push_back(id); vals.push_back(std::move(v)); live.push_back(1); ++n_live;
  }
  void erase(i64 id) {
    auto it = pos.find(id);
    if (it == pos.end()) return;
    live[it->second] = 0; vals[it->second].clear(); vals[it->second].shrink_to_fit();
    pos.erase(it); --n_live;
    if (ids.size() > 64 && n_live * 2 < ids.size()) compact();
  }
  void compact() {
    size_t w = 0;
    for (size_t r = 0; r < ids.size(); ++r)
      if (live[r]) {
        if (w != r) { ids[w] = ids[r]; vals[w] = std::move(vals[r]); live[w] = 1; }
        pos[ids[w]] = w; ++w;
      }
    ids.resize(w); vals.resize(w); live.resize(w);
  }
};

struct Episode {
  std::vector<float> patch_pos;  // host mirror [n_patch*3]
  std::vector<i64> p2i;          // patch id -> instance id, -1 = not a key
  i64 n_patch = 0, n_p2i = 0;
  OMap i2p;
  std::vector<unsigned char> gone;  // scratch flags of the cull (all zero between calls)
  std::vector<unsigned char> inst_alive;
  std::vector<float> inst_pos;   // host mirror [n_inst*3]
  i64 n_inst = 0;
  std::unordered_map<i64, i64> zone_code_to_id;
  OMap z2i;
  std::vector<unsigned char> zone_alive;
  i64 n_zone = 0;
  bool tree = false;
  // trace of the last processed view (parity tests)
  int last_K = -1, last_G = 0;
  std::vector<float> last_d2;
  std::vector<int> last_idx;
  std::vector<unsigned char> last_merge;
};

struct ViewPlan {  // results of finish_view, fetched by the caller
  std::vector<int> new_src, new_owner; std::vector<i64> new_iid;
  std::vector<int> mg_owner; std::vector<i64> mg_iid; std::vector<float> mg_pos; std::vector<int> mg_len; std::vector<int> mg_members;
  std::vector<int> zn_owner; std::vector<i64> zn_slot; std::vector<int> zn_keys; std::vector<float> zn_pos; std::vector<int> zn_len;
  std::vector<int> zn_members;
  void clear() {
    new_src.clear(); new_owner.clear(); new_iid.clear(); mg_owner.clear(); mg_iid.clear(); mg_pos.clear(); mg_len.clear(); mg_members.clear();
    zn_owner.clear(); zn_slot.clear(); zn_keys.clear(); zn_pos.clear(); zn_len.clear(); zn_members.clear();
  }
};

struct FFH {
  std::vector<Episode> eps;
  int num_proposal = 2;
  float zone_len = 2.0f;
  // view in flight
  int P = 0;
  std::vector<std::vector<std::vector<int>>> splits;  // [b][g] -> view-local patch indices (ascending)
  std::vector<int> seq_start;                          // [B+1]
  int res_base = 0;                                    // sequence index of the first row of the `res` array handed to finish_view
  // whole-step plan (begin_step): splits / sequence starts of every view, sequences numbered globally in (view, episode, segment) order
  std::vector<std::vector<std::vector<std::vector<int>>>> step_splits;  // [V][b][g]
  std::vector<std::vector<int>> step_seq_start;                        // [V][B+1]
  // per-view runtime (d3d_ff_view_pre / _post): upload ring cursor and the zone pooling pass waiting to be issued
  size_t stage_cursor = 0;
  struct Deferred {
    bool pending = false;
    const int64_t* ptrs; const float* centre; const int* tok_seq; const int* tok_src; const int* cu; const int64_t* fts_dst; const int64_t* pos_dst;
    float* out; int T, n_seq, max_len;
  } zone;
  ViewPlan plan;
};

inline i64 voxel_code(const float* p, float L) {
  const i64 x = (i64)floorf(p[0] / L) + VOFF, y = (i64)floorf(p[1] / L) + VOFF, z = (i64)floorf(p[2] / L) + VOFF;
  return (x * VM + y) * VM + z;
}
inline void code_to_key(i64 c, float L, float* out) {
  const i64 z = c % VM; c /= VM;
  const i64 y = c % VM; const i64 x = c / VM;
  out[0] = (float)(x - VOFF) * L + L / 2; out[1] = (float)(y - VOFF) * L + L / 2; out[2] = (float)(z - VOFF) * L + L / 2;
}

void lowest_free(const std::vector<unsigned char>& alive, size_t n_keys, size_t n, std::vector<i64>& out) {
  out.clear();
  for (size_t i = 0; i < n_keys + n && out.size() < n; ++i)
    if (i >= alive.size() || !alive[i]) out.push_back((i64)i);
}

void mean64(const std::vector<float>& pos, const i64* ids, size_t n, float* out) {
  if (n == 0) { out[0] = out[1] = out[2] = NAN; return; }
  double s[3] = {0, 0, 0};
  for (size_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) s[c] += (double)pos[(size_t)ids[i] * 3 + c];
  for (int c = 0; c < 3; ++c) out[c] = (float)(s[c] / (double)n);
}

int update_episode(FFH& H, int b, const float* cen, const int* idx, const float* d2, const float* logits, int s0) {
  Episode& ep = H.eps[b];
  const auto& splits = H.splits[b];
  const int G = (int)splits.size();
  const int P = H.P;
  ViewPlan& pl = H.plan;
  // FF:433-445: the P lowest patch ids that are not keys of the patch->instance map
  std::vector<i64> patch_ids;
  patch_ids.reserve(P);
  for (i64 i = 0; i < ep.n_p2i + P && (int)patch_ids.size() < P; ++i)
    if (ep.p2i[(size_t)i] < 0) patch_ids.push_back(i);
  std::vector<i64> members;
  if (ep.tree) {
    int K = (int)std::min<i64>((i64)ep.i2p.n_live, (i64)H.num_proposal);
    if (K > 0) {  // Q9: K-shrink heuristic (FF:607-610)
      double tot = 0;
      for (int g = 0; g < G; ++g)
        for (int j = 0; j < K; ++j) tot += (double)d2[g * 2 + j];
      if (tot > 1e6) {
        int k2 = 0;
        for (int j = 0; j < K; ++j) {
          double cs = 0;
          for (int g = 0; g < G; ++g) cs += (double)d2[g * 2 + j];
          if (cs < 1e6) ++k2;
        }
        K = k2;
      }
    }
    ep.last_K = K; ep.last_G = G;
    ep.last_d2.assign((size_t)G * K, 0.f); ep.last_idx.assign((size_t)G * K, 0); ep.last_merge.assign((size_t)G * K, 0);
    std::vector<int> first(G, -1);
    int n_new = 0;
    for (int g = 0; g < G; ++g) {
      for (int j = 0; j < K; ++j) {
        const bool mt = logits[(g * 2 + j) * 2 + 1] > logits[(g * 2 + j) * 2];  // argmax of the 2-way softmax, first max wins
        ep.last_d2[(size_t)g * K + j] = d2[g * 2 + j]; ep.last_idx[(size_t)g * K + j] = idx[g * 2 + j]; ep.last_merge[(size_t)g * K + j] = mt;
        if (mt && first[g] < 0) first[g] = j;  // only the nearest accepted proposal (FF:653,691)
      }
      if (first[g] < 0) ++n_new;
    }
    std::vector<i64> new_ids;
    lowest_free(ep.inst_alive, ep.i2p.n_live, (size_t)n_new, new_ids);
    std::vector<i64> touched;
    int ni = 0;
    for (int g = 0; g < G; ++g) {
      members.clear();
      for (int p : splits[g]) members.push_back(patch_ids[(size_t)p]);
      i64 iid;
      if (first[g] < 0) {
        iid = new_ids[(size_t)ni++];
        if (iid >= ep.n_inst) ep.n_inst = iid + 1;
        if ((size_t)ep.n_inst * 3 > ep.inst_pos.size()) ep.inst_pos.resize((size_t)ep.n_inst * 3, 0.f);
        if ((size_t)ep.n_inst > ep.inst_alive.size()) ep.inst_alive.resize((size_t)ep.n_inst, 0);
        ep.i2p.assign(iid, std::vector<i64>(members));
        ep.inst_alive[(size_t)iid] = 1;
        for (int c = 0; c < 3; ++c) ep.inst_pos[(size_t)iid * 3 + c] = cen[g * 3 + c];
        pl.new_src.push_back(s0 + g); pl.new_owner.push_back(b); pl.new_iid.push_back(iid);
      } else {
        iid = (i64)idx[g * 2 + first[g]];
        if (iid < 0 || !ep.i2p.has(iid)) {
          d3d_set_error("merge target instance %lld of episode %d is not alive (the reference raises KeyError here, FF:658)", iid, b);
          return D3D_EINVAL;
        }
        auto& lst = ep.i2p.at(iid);
        lst.insert(lst.end(), members.begin(), members.end());
        if (std::find(touched.begin(), touched.end(), iid) == touched.end()) touched.push_back(iid);
      }
      for (i64 m : members) ep.p2i[(size_t)m] = iid;
    }
    ep.n_p2i += P;
    for (i64 iid : touched) {  // only the state after the last merge survives (FF:663,688 overwrite)
      const auto& ids = ep.i2p.at(iid);
      float pos[3];
      mean64(ep.patch_pos, ids.data(), ids.size(), pos);  // Q2: ids index the patch arrays directly
      for (int c = 0; c < 3; ++c) ep.inst_pos[(size_t)iid * 3 + c] = pos[c];
      pl.mg_owner.push_back(b); pl.mg_iid.push_back(iid);
      pl.mg_pos.insert(pl.mg_pos.end(), pos, pos + 3);
      pl.mg_len.push_back((int)ids.size());
      for (i64 v : ids) pl.mg_members.push_back((int)v);
    }
  } else {
    ep.last_K = -1;
    std::vector<i64> ids;
    lowest_free(ep.inst_alive, ep.i2p.n_live, (size_t)G, ids);
    ep.n_inst = G;
    ep.inst_pos.assign(cen, cen + (size_t)G * 3);
    if ((size_t)G > ep.inst_alive.size()) ep.inst_alive.resize((size_t)G, 0);
    for (int g = 0; g < G; ++g) {
      members.clear();
      for (int p : splits[g]) members.push_back(patch_ids[(size_t)p]);
      const i64 iid = ids[(size_t)g];
      ep.i2p.assign(iid, std::vector<i64>(members));
      ep.inst_alive[(size_t)iid] = 1;
      for (i64 m : members) ep.p2i[(size_t)m] = iid;
      pl.new_src.push_back(s0 + g); pl.new_owner.push_back(b); pl.new_iid.push_back(iid);
    }
    ep.n_p2i += P;
  }
  // zones (FF:693-756 / 777-812): group the instance slots by voxel, then visit the view's voxels in key order
  const float L = H.zone_len;
  const size_t NI = (size_t)ep.n_inst;
  std::vector<i64> uniq(G);
  for (int g = 0; g < G; ++g) uniq[(size_t)g] = voxel_code(cen + g * 3, L);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  // members of the view's voxels: ONE ascending pass over the instance slots (no sort of all slots: the memory holds thousands of them in
  // a long rollout while a view touches a handful of voxels); ascending slot order inside a voxel, as the reference's boolean mask gives
  std::vector<std::vector<i64>> members_of(uniq.size());
  for (size_t i = 0; i < NI; ++i) {
    const i64 code = voxel_code(&ep.inst_pos[i * 3], L);
    auto it = std::lower_bound(uniq.begin(), uniq.end(), code);
    if (it != uniq.end() && *it == code) members_of[(size_t)(it - uniq.begin())].push_back((i64)i);
  }
  std::vector<i64> zone_ids;
  lowest_free(ep.zone_alive, ep.z2i.n_live, uniq.size(), zone_ids);
  size_t zi = 0;
  for (size_t ui = 0; ui < uniq.size(); ++ui) {
    const i64 code = uniq[ui];
    std::vector<i64>& mem = members_of[ui];
    float pos[3];
    auto zit = ep.zone_code_to_id.find(code);
    i64 slot;
    int use_keys;
    if (zit == ep.zone_code_to_id.end()) {
      const i64 zid = zone_ids[zi++];
      ep.zone_code_to_id[code] = zid;
      if ((size_t)zid >= ep.zone_alive.size()) ep.zone_alive.resize((size_t)zid + 1, 0);
      ep.zone_alive[(size_t)zid] = 1;
      mean64(ep.inst_pos, mem.data(), mem.size(), pos);  // empty voxel -> NaN (Q5)
      slot = ep.n_zone++;                                 // Q3: a new zone is always appended, whatever its id
      use_keys = 0;
      ep.z2i.assign(zid, std::vector<i64>(mem));
    } else {
      const i64 zid = zit->second;
      if (mem.empty()) pos[0] = pos[1] = pos[2] = NAN; else code_to_key(code, L, pos);  // Q5: mean of identical voxel-centre keys
      slot = zid;
      use_keys = 1;
      ep.z2i.assign(zid, std::vector<i64>(mem));
    }
    pl.zn_owner.push_back(b); pl.zn_slot.push_back(slot); pl.zn_keys.push_back(use_keys);
    pl.zn_pos.insert(pl.zn_pos.end(), pos, pos + 3);
    pl.zn_len.push_back((int)mem.size());
    for (i64 v : mem) pl.zn_members.push_back((int)v);
  }
  ep.tree = ep.n_inst > 0;
  return 0;
}

}  // namespace

#define HH(h) (*reinterpret_cast<FFH*>(h))

extern "C" void* d3d_ffh_create(int batch_size, int num_proposal, float zone_len) {
  FFH* h = new FFH();
  h->eps.resize((size_t)batch_size);
  h->num_proposal = num_proposal;
  h->zone_len = zone_len;
  return h;
}
extern "C" void d3d_ffh_destroy(void* h) { delete reinterpret_cast<FFH*>(h); }
// a step that raised between view_post and the next run_deferred leaves a pending zone pass / plan behind: a new rollout must not see them
static void clear_runtime(FFH& H) {
  H.zone.pending = false;
  H.plan.clear();
  H.splits.clear(); H.seq_start.clear();
  H.step_splits.clear(); H.step_seq_start.clear();
  H.stage_cursor = 0;
  H.res_base = 0;
}
extern "C" int d3d_ffh_reset(void* h, int batch_size) {
  HH(h).eps.clear();
  HH(h).eps.resize((size_t)batch_size);
  clear_runtime(HH(h));
  return 0;
}
extern "C" int d3d_ffh_pop(void* h, int index) {
  FFH& H = HH(h);
  D3D_REQUIRE(index >= 0 && index < (int)H.eps.size(), "episode index");
  H.eps.erase(H.eps.begin() + index);
  clear_runtime(H);
  return 0;
}

// counts[8] = n_patch, n_p2i, n_inst, live instances, n_zone, live zones, tree, last_K
extern "C" int d3d_ffh_counts(void* h, int b, int64_t* counts) {
  FFH& H = HH(h);
  D3D_REQUIRE(b >= 0 && b < (int)H.eps.size(), "episode index");
  const Episode& e = H.eps[(size_t)b];
  counts[0] = e.n_patch; counts[1] = e.n_p2i; counts[2] = e.n_inst; counts[3] = (int64_t)e.i2p.n_live; counts[4] = e.n_zone;
  counts[5] = (int64_t)e.z2i.n_live; counts[6] = e.tree ? 1 : 0; counts[7] = e.last_K;
  return 0;
}

// FF:362-393 for the rows culled in this step (any order).  dead_inst / dead_zone must hold n_inst / n_zone entries.
static int cull_rows(FFH& H, int b, const int32_t* rows, int64_t n_rows, const uint8_t* mask, int64_t n_mask, int64_t* dead_inst, int* n_dead_inst,
                     int64_t* dead_zone, int* n_dead_zone) {
  Episode& ep = H.eps[(size_t)b];
  *n_dead_inst = 0; *n_dead_zone = 0;
  std::vector<i64> owners;
  std::vector<i64> hit;  // rows that were keys of the patch -> instance map
  auto visit = [&](i64 r) {
    ep.patch_pos[(size_t)r * 3] = ep.patch_pos[(size_t)r * 3 + 1] = ep.patch_pos[(size_t)r * 3 + 2] = -10000.0f;
    const i64 own = ep.p2i[(size_t)r];  // Q2: array index used as patch id
    if (own < 0) return;
    if (ep.gone.size() < (size_t)ep.n_patch) ep.gone.resize((size_t)ep.n_patch, 0);
    ep.gone[(size_t)r] = 1;
    hit.push_back(r);
    ep.p2i[(size_t)r] = -1;
    --ep.n_p2i;
    owners.push_back(own);
  };
  if (rows) {
    for (int64_t k = 0; k < n_rows; ++k) {
      D3D_REQUIRE(rows[k] >= 0 && rows[k] < ep.n_patch, "culled row out of range");
      visit(rows[k]);
    }
  } else {
    for (i64 r = 0; r < n_mask; ++r)
      if (mask[r]) visit(r);
  }
  if (!owners.empty()) {
    std::sort(owners.begin(), owners.end());
    owners.erase(std::unique(owners.begin(), owners.end()), owners.end());
    for (i64 iid : owners) {
      auto& m = ep.i2p.at(iid);
      size_t w = 0;
      for (size_t i = 0; i < m.size(); ++i)
        if (!ep.gone[(size_t)m[i]]) m[w++] = m[i];
      m.resize(w);
      if (w) continue;
      ep.i2p.erase(iid);
      ep.inst_alive[(size_t)iid] = 0;
      const i64 code = voxel_code(&ep.inst_pos[(size_t)iid * 3], H.zone_len);
      ep.inst_pos[(size_t)iid * 3] = ep.inst_pos[(size_t)iid * 3 + 1] = ep.inst_pos[(size_t)iid * 3 + 2] = -10000.0f;
      dead_inst[(*n_dead_inst)++] = iid;
      auto zit = ep.zone_code_to_id.find(code);
      if (zit == ep.zone_code_to_id.end()) continue;
      const i64 zid = zit->second;
      auto& z = ep.z2i.at(zid);
      z.erase(std::remove(z.begin(), z.end(), iid), z.end());
      if (!z.empty()) continue;
      ep.zone_code_to_id.erase(zit);
      ep.z2i.erase(zid);
      ep.zone_alive[(size_t)zid] = 0;
      dead_zone[(*n_dead_zone)++] = zid;
    }
    for (i64 r : hit) ep.gone[(size_t)r] = 0;  // the scratch flags stay all-zero between calls: no O(n_patch) clear per step
  }
  ep.tree = ep.n_inst > 0;
  return 0;
}

// mask [n_patch] (1 = culled now): the per-episode form (d3d_frustum_cull / d3d_frustum_cull_matrix)
extern "C" int d3d_ffh_cull(void* h, int b, const uint8_t* mask, int64_t n, int64_t* dead_inst, int* n_dead_inst, int64_t* dead_zone,
                            int* n_dead_zone) {
  FFH& H = HH(h);
  D3D_REQUIRE(b >= 0 && b < (int)H.eps.size(), "episode index");
  D3D_REQUIRE(n == H.eps[(size_t)b].n_patch, "mask length != number of stored patches");
  return cull_rows(H, b, nullptr, 0, mask, n, dead_inst, n_dead_inst, dead_zone, n_dead_zone);
}
// rows [n] = the compacted list of d3d_frustum_cull_batched
extern "C" int d3d_ffh_cull_list(void* h, int b, const int32_t* rows, int64_t n, int64_t* dead_inst, int* n_dead_inst, int64_t* dead_zone,
                                 int* n_dead_zone) {
  FFH& H = HH(h);
  D3D_REQUIRE(b >= 0 && b < (int)H.eps.size(), "episode index");
  D3D_REQUIRE(rows || n == 0, "rows");
  static const int32_t none = 0;
  return cull_rows(H, b, rows ? rows : &none, n, nullptr, 0, dead_inst, n_dead_inst, dead_zone, n_dead_zone);
}
extern "C" int d3d_ffh_set_tree(void* h) {
  for (auto& ep : HH(h).eps) ep.tree = ep.n_inst > 0;
  return 0;
}

// First half of a view (before the device results exist), all episodes in lock step.
//   xyz [B,P,3] this view's unprojected patches (host mirror), segm [B,P] dense labels, stage_off[B] = first stage row of (b, view).
// Appends the patches to the mirrors and emits the packed arrays of the patch->instance pooling pass:
//   base_rows[B], n_seg[B], seq_owner[n_seq], members[B*P] (stage rows, grouped by segment), cu_m[n_seq+1],
//   tok_src[B*P+n_seq], tok_seq[B*P+n_seq], cu_tok[n_seq+1], n_ref[n_seq];  info[0]=n_seq, info[1]=max sequence length (+1 token)
extern "C" int d3d_ffh_begin_view(void* h, const float* xyz, const int64_t* segm, int P, const int64_t* stage_off, int64_t* base_rows, int* n_seg,
                                  int* seq_owner, int* members, int* cu_m, int* tok_src, int* tok_seq, int* cu_tok, int* n_ref, int* info) {
  FFH& H = HH(h);
  const int B = (int)H.eps.size();
  H.P = P;
  H.splits.assign((size_t)B, {});
  H.seq_start.assign((size_t)B + 1, 0);
  int n_seq = 0, max_len = 0;
  size_t mpos = 0, tpos = 0;
  cu_m[0] = 0; cu_tok[0] = 0;
  for (int b = 0; b < B; ++b) {
    Episode& ep = H.eps[(size_t)b];
    base_rows[b] = ep.n_patch;
    ep.patch_pos.insert(ep.patch_pos.end(), xyz + (size_t)b * P * 3, xyz + (size_t)(b + 1) * P * 3);
    ep.n_patch += P;
    if (ep.p2i.size() < (size_t)ep.n_patch) ep.p2i.resize((size_t)ep.n_patch, -1);
    const int64_t* sg = segm + (size_t)b * P;
    int G = 0;
    for (int p = 0; p < P; ++p) {
      D3D_REQUIRE(sg[p] >= 0 && sg[p] < P, "segment label out of range");
      G = std::max(G, (int)sg[p] + 1);
    }
    auto& sp = H.splits[(size_t)b];
    sp.assign((size_t)G, {});
    for (int p = 0; p < P; ++p) sp[(size_t)sg[p]].push_back(p);  // stable: ascending patch order inside a segment
    n_seg[b] = G;
    H.seq_start[(size_t)b] = n_seq;
    const int nref = ep.tree ? (int)ep.n_inst : 0;
    for (int g = 0; g < G; ++g) {
      D3D_REQUIRE(!sp[(size_t)g].empty(), "patch_segm labels must be dense 0..G-1 (FF:411-422 relabels them)");
      seq_owner[n_seq] = b;
      n_ref[n_seq] = nref;
      tok_src[tpos] = -1; tok_seq[tpos] = n_seq; ++tpos;
      for (int p : sp[(size_t)g]) {
        const int row = (int)(stage_off[b] + p);
        members[mpos++] = row;
        tok_src[tpos] = row; tok_seq[tpos] = n_seq; ++tpos;
      }
      max_len = std::max(max_len, (int)sp[(size_t)g].size() + 1);
      ++n_seq;
      cu_m[n_seq] = (int)mpos;
      cu_tok[n_seq] = (int)tpos;
    }
  }
  H.seq_start[(size_t)B] = n_seq;
  H.res_base = 0;
  info[0] = n_seq; info[1] = max_len;
  return 0;
}

// Whole-step variant of begin_view: the patch -> instance pooling of a view does not depend on the memory state, so ALL views of a step
// are planned (and pooled on the device) in one packed batch; only the K-NN / merge / zone part stays per view (begin_view_refs).
//   xyz [V,B,P,3], segm [V,B,P]; stage row of (episode b, view ix, patch p) = (b*V + ix)*P + p.
// Outputs: base_rows[B] (first pool row of the step's patches; view ix starts at base + ix*P), view_seq_start[V+1], seq_owner[n_seq],
//   members[V*B*P], cu_m[n_seq+1], tok_src / tok_seq [V*B*P + n_seq], cu_tok[n_seq+1]; info[0] = n_seq, info[1] = max sequence length.
extern "C" int d3d_ffh_begin_step(void* h, const float* xyz, const int64_t* segm, int P, int V, int64_t* base_rows, int* view_seq_start,
                                  int* seq_owner, int* members, int* cu_m, int* tok_src, int* tok_seq, int* cu_tok, int* info) {
  FFH& H = HH(h);
  const int B = (int)H.eps.size();
  H.P = P;
  H.step_splits.assign((size_t)V, {});
  H.step_seq_start.assign((size_t)V, std::vector<int>((size_t)B + 1, 0));
  for (int b = 0; b < B; ++b) {
    Episode& ep = H.eps[(size_t)b];
    base_rows[b] = ep.n_patch;
    for (int ix = 0; ix < V; ++ix) {
      const float* src = xyz + ((size_t)ix * B + b) * P * 3;
      ep.patch_pos.insert(ep.patch_pos.end(), src, src + (size_t)P * 3);
    }
    ep.n_patch += (i64)P * V;
    if (ep.p2i.size() < (size_t)ep.n_patch) ep.p2i.resize((size_t)ep.n_patch, -1);
  }
  int n_seq = 0, max_len = 0;
  size_t mpos = 0, tpos = 0;
  cu_m[0] = 0; cu_tok[0] = 0;
  for (int ix = 0; ix < V; ++ix) {
    view_seq_start[ix] = n_seq;
    auto& vs = H.step_splits[(size_t)ix];
    vs.assign((size_t)B, {});
    for (int b = 0; b < B; ++b) {
      const int64_t* sg = segm + ((size_t)ix * B + b) * P;
      int G = 0;
      for (int p = 0; p < P; ++p) {
        D3D_REQUIRE(sg[p] >= 0 && sg[p] < P, "segment label out of range");
        G = std::max(G, (int)sg[p] + 1);
      }
      auto& sp = vs[(size_t)b];
      sp.assign((size_t)G, {});
      for (int p = 0; p < P; ++p) sp[(size_t)sg[p]].push_back(p);
      H.step_seq_start[(size_t)ix][(size_t)b] = n_seq;
      const long long stage0 = ((long long)b * V + ix) * P;
      for (int g = 0; g < G; ++g) {
        D3D_REQUIRE(!sp[(size_t)g].empty(), "patch_segm labels must be dense 0..G-1 (FF:411-422 relabels them)");
        seq_owner[n_seq] = b;
        tok_src[tpos] = -1; tok_seq[tpos] = n_seq; ++tpos;
        for (int p : sp[(size_t)g]) {
          const int row = (int)(stage0 + p);
          members[mpos++] = row;
          tok_src[tpos] = row; tok_seq[tpos] = n_seq; ++tpos;
        }
        max_len = std::max(max_len, (int)sp[(size_t)g].size() + 1);
        ++n_seq;
        cu_m[n_seq] = (int)mpos;
        cu_tok[n_seq] = (int)tpos;
      }
    }
    H.step_seq_start[(size_t)ix][(size_t)B] = n_seq;
  }
  view_seq_start[V] = n_seq;
  info[0] = n_seq; info[1] = max_len;
  return 0;
}

// Makes view `ix` of the step planned by begin_step the view in flight: n_ref[s] (instance slots the K-NN of sequence s searches, 0 = no
// tree yet) for the view's sequences, in order.  finish_view then expects `res` rows of exactly these sequences.
extern "C" int d3d_ffh_begin_view_refs(void* h, int ix, int* n_ref) {
  FFH& H = HH(h);
  D3D_REQUIRE(ix >= 0 && ix < (int)H.step_splits.size(), "view index");
  const int B = (int)H.eps.size();
  H.splits = H.step_splits[(size_t)ix];
  H.seq_start = H.step_seq_start[(size_t)ix];
  H.res_base = H.seq_start[0];
  for (int b = 0; b < B; ++b) {
    const Episode& ep = H.eps[(size_t)b];
    const int nref = ep.tree ? (int)ep.n_inst : 0;
    for (int s = H.seq_start[(size_t)b]; s < H.seq_start[(size_t)b + 1]; ++s) n_ref[s - H.res_base] = nref;
  }
  return 0;
}

// Second half: res [n_seq, 12] fp32 rows = [centre(3) | d2(2) | idx(2, int32 bits) | logits(2x2) | pad] copied back from the device.
// sizes[10] = n_new, n_merged, merged members, n_zones, zone members, max merged len, max zone len, upload KiB d3d_ff_view_post needs, -, -
// after[B*3] = (n_inst, n_zone, needs_key_array) per episode.
extern "C" int d3d_ffh_finish_view(void* h, const float* res, int* sizes, int64_t* after) {
  FFH& H = HH(h);
  const int B = (int)H.eps.size();
  H.plan.clear();
  for (int b = 0; b < B; ++b) {
    const int s0 = H.seq_start[(size_t)b], s1 = H.seq_start[(size_t)b + 1];
    const int G = s1 - s0;
    std::vector<float> cen((size_t)G * 3), d2((size_t)G * 2), lg((size_t)G * 4);
    std::vector<int> idx((size_t)G * 2);
    for (int g = 0; g < G; ++g) {
      const float* r = res + (size_t)(s0 - H.res_base + g) * 12;
      memcpy(&cen[(size_t)g * 3], r, 12);
      memcpy(&d2[(size_t)g * 2], r + 3, 8);
      memcpy(&idx[(size_t)g * 2], r + 5, 8);
      memcpy(&lg[(size_t)g * 4], r + 7, 16);
    }
    D3D_TRY(update_episode(H, b, cen.data(), idx.data(), d2.data(), lg.data(), s0));
  }
  const ViewPlan& pl = H.plan;
  sizes[0] = (int)pl.new_src.size(); sizes[1] = (int)pl.mg_owner.size(); sizes[2] = (int)pl.mg_members.size();
  sizes[3] = (int)pl.zn_owner.size(); sizes[4] = (int)pl.zn_members.size();
  int ml = 0, zl = 0;
  for (int v : pl.mg_len) ml = std::max(ml, v);
  for (int v : pl.zn_len) zl = std::max(zl, v);
  sizes[5] = ml + 1; sizes[6] = zl + 1;
  for (int b = 0; b < B; ++b) {
    after[b * 3] = H.eps[(size_t)b].n_inst; after[b * 3 + 1] = H.eps[(size_t)b].n_zone; after[b * 3 + 2] = 0;
  }
  for (size_t i = 0; i < pl.zn_owner.size(); ++i)
    if (pl.zn_keys[i]) after[pl.zn_owner[i] * 3 + 2] = 1;
  // bytes d3d_ff_view_post stages for this plan (every array rounded up to 16 B): the caller grows the upload ring BEFORE the device writes
  // are issued, so an oversized plan cannot fail after the host state was already mutated
  size_t up = 0;
  auto add = [&up](size_t bytes) { up += (bytes + 15) / 16 * 16; };
  const size_t n_new = pl.new_src.size(), n_mg = pl.mg_owner.size(), n_zn = pl.zn_owner.size();
  add(n_new * 4); add(n_new * 8); add(n_new * 8);
  const size_t t_mg = pl.mg_members.size() + n_mg, t_zn = pl.zn_members.size() + n_zn;
  add(t_mg * 4); add(t_mg * 4); add((n_mg + 1) * 4); add(n_mg * 32); add(n_mg * 12); add(n_mg * 8); add(n_mg * 8);
  add(t_zn * 4); add(t_zn * 4); add((n_zn + 1) * 4); add(n_zn * 32); add(n_zn * 12); add(n_zn * 8); add(n_zn * 8);
  sizes[7] = (int)std::min<size_t>((up + 1023) / 1024, (size_t)INT32_MAX);
  return 0;
}

// Copies the plan of the last finish_view.  Token arrays for the merged / zone pooling passes are emitted ready to upload.
extern "C" int d3d_ffh_fetch_view(void* h, int* new_src, int* new_owner, int64_t* new_iid, int* mg_owner, int64_t* mg_iid, float* mg_pos,
                                  int* mg_tok_src, int* mg_tok_seq, int* mg_cu, int* zn_owner, int64_t* zn_slot, int* zn_keys, float* zn_pos,
                                  int* zn_tok_src, int* zn_tok_seq, int* zn_cu) {
  const ViewPlan& pl = HH(h).plan;
  auto cp = [](auto* dst, const auto& v) { if (!v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
  cp(new_src, pl.new_src); cp(new_owner, pl.new_owner); cp(new_iid, pl.new_iid);
  cp(mg_owner, pl.mg_owner); cp(mg_iid, pl.mg_iid); cp(mg_pos, pl.mg_pos);
  cp(zn_owner, pl.zn_owner); cp(zn_slot, pl.zn_slot); cp(zn_keys, pl.zn_keys); cp(zn_pos, pl.zn_pos);
  auto toks = [](const std::vector<int>& len, const std::vector<int>& mem, int* src, int* seq, int* cu) {
    size_t t = 0, m = 0;
    cu[0] = 0;
    for (size_t s = 0; s < len.size(); ++s) {
      src[t] = -1; seq[t] = (int)s; ++t;
      for (int i = 0; i < len[s]; ++i) { src[t] = mem[m++]; seq[t] = (int)s; ++t; }
      cu[s + 1] = (int)t;
    }
  };
  toks(pl.mg_len, pl.mg_members, mg_tok_src, mg_tok_seq, mg_cu);
  toks(pl.zn_len, pl.zn_members, zn_tok_src, zn_tok_seq, zn_cu);
  return 0;
}

// voxel-centre keys of all instance slots of episode b (FF:694), [n_inst,3] fp32 -- the position source of UPDATED zones (Q5)
extern "C" int d3d_ffh_zone_key_array(void* h, int b, float* out) {
  FFH& H = HH(h);
  D3D_REQUIRE(b >= 0 && b < (int)H.eps.size(), "episode index");
  const Episode& ep = H.eps[(size_t)b];
  for (i64 i = 0; i < ep.n_inst; ++i) {
    const float* p = &ep.inst_pos[(size_t)i * 3];
    for (int c = 0; c < 3; ++c) out[(size_t)i * 3 + c] = floorf(p[c] / H.zone_len) * H.zone_len + H.zone_len / 2;
  }
  return 0;
}

// ---- state export (reference-style views, parity tests, get_environment_features) ----
// which: 0 = instance -> patches, 1 = zone -> instances.  Pass NULL buffers to query sizes: sizes[0] = live entries, sizes[1] = total members.
extern "C" int d3d_ffh_get_map(void* h, int b, int which, int64_t* ids, int64_t* lens, int64_t* cat, int64_t* sizes) {
  FFH& H = HH(h);
  D3D_REQUIRE(b >= 0 && b < (int)H.eps.size(), "episode index");
  const OMap& m = which == 0 ? H.eps[(size_t)b].i2p : H.eps[(size_t)b].z2i;
  size_t n = 0, tot = 0;
  for (size_t r = 0; r < m.ids.size(); ++r) {
    if (!m.live[r]) continue;
    if (ids) { ids[n] = m.ids[r]; lens[n] = (int64_t)m.vals[r].size(); memcpy(cat + tot, m.vals[r].data(), m.vals[r].size() * sizeof(i64)); }
    ++n; tot += m.vals[r].size();
  }
  if (sizes) { sizes[0] = (int64_t)n; sizes[1] = (int64_t)tot; }
  return 0;
}
// dict-order keys only (get_environment_features FF:825,844): ids may be NULL to query the count
extern "C" int d3d_ffh_live_ids(void* h, int b, int which, int64_t* ids, int64_t* n_out) {
  FFH& H = HH(h);
  D3D_REQUIRE(b >= 0 && b < (int)H.eps.size(), "episode index");
  const OMap& m = which == 0 ? H.eps[(size_t)b].i2p : H.eps[(size_t)b].z2i;
  size_t n = 0;
  for (size_t r = 0; r < m.ids.size(); ++r) {
    if (!m.live[r]) continue;
    if (ids) ids[n] = m.ids[r];
    ++n;
  }
  *n_out = (int64_t)n;
  return 0;
}
extern "C" int d3d_ffh_get_p2i(void* h, int b, int64_t* out) {
  const Episode& ep = HH(h).eps[(size_t)b];
  if (ep.n_patch) memcpy(out, ep.p2i.data(), (size_t)ep.n_patch * sizeof(i64));
  return 0;
}
extern "C" int d3d_ffh_get_patch_pos(void* h, int b, float* out) {
  const Episode& ep = HH(h).eps[(size_t)b];
  if (ep.n_patch) memcpy(out, ep.patch_pos.data(), (size_t)ep.n_patch * 3 * sizeof(float));
  return 0;
}
// zone keys as the reference's float triples + ids, in no particular order; returns the count through *n
extern "C" int d3d_ffh_get_zone_keys(void* h, int b, float* keys, int64_t* ids, int64_t* n) {
  FFH& H = HH(h);
  const Episode& ep = H.eps[(size_t)b];
  size_t i = 0;
  for (const auto& kv : ep.zone_code_to_id) {
    if (keys) { code_to_key(kv.first, H.zone_len, keys + i * 3); ids[i] = kv.second; }
    ++i;
  }
  *n = (int64_t)i;
  return 0;
}
// last view's proposals of episode b: K columns; d2 / idx / merge [G*K]; n[0] = G*K, n[1] = G
extern "C" int d3d_ffh_get_last(void* h, int b, float* d2, int* idx, uint8_t* merge, int64_t* n) {
  const Episode& ep = HH(h).eps[(size_t)b];
  n[0] = (int64_t)ep.last_d2.size(); n[1] = ep.last_G;
  if (d2 && !ep.last_d2.empty()) {
    memcpy(d2, ep.last_d2.data(), ep.last_d2.size() * 4); memcpy(idx, ep.last_idx.data(), ep.last_idx.size() * 4);
    memcpy(merge, ep.last_merge.data(), ep.last_merge.size());
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Per-view runtime: the state-dependent part of a panorama view (FF:604-756) behind TWO host calls, so the interpreter is not on the
// critical path of the host-synchronised loop:  d3d_ff_view_pre  = K-NN proposals + merge discriminator + result copy + (previous view's
// deferred zone pass) + wait + planner;   [caller grows the episode pools if the plan needs more slots]   d3d_ff_view_post = slot
// writes, merged-instance pooling pass, upload of the zone pass (issued by the NEXT pre call or d3d_ff_run_deferred).
// ------------------------------------------------------------------------------------------------
extern "C" int d3d_knn2_batched(const int64_t* ref_ptr, const int* n_ref, const float* queries, int n_q, float* out_d2, int* out_idx, void* stream);
extern "C" int d3d_disc_input_batched(const int64_t* fts_ptr, const int64_t* pos_ptr, const int* idx, const float* view_fts, const float* centre,
                                      int Q, int K, int D, int ldo, void* out16, int kind, void* stream);
extern "C" int d3d_scatter_rows_ptr(const float* src, int64_t lds, const int* src_idx, const int64_t* dst_row_ptr, int n, int D, void* stream);

namespace {

__global__ void pack_view_result_kernel(const float* __restrict__ centre, const float* __restrict__ d2, const int* __restrict__ idx,
                                        const float* __restrict__ logits4, int n, int has_refs, float* __restrict__ res) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* r = res + (size_t)i * 12;
  r[0] = centre[i * 3]; r[1] = centre[i * 3 + 1]; r[2] = centre[i * 3 + 2];
  if (has_refs) {
    r[3] = d2[i * 2]; r[4] = d2[i * 2 + 1];
    r[5] = __int_as_float(idx[i * 2]); r[6] = __int_as_float(idx[i * 2 + 1]);
    // logits4 rows are [2*n, 4] (ld 4): proposal j of segment i = row 2i+j, columns 0..1
    r[7] = logits4[(2 * i) * 4]; r[8] = logits4[(2 * i) * 4 + 1]; r[9] = logits4[(2 * i + 1) * 4]; r[10] = logits4[(2 * i + 1) * 4 + 1];
  } else {
    for (int k = 3; k < 11; ++k) r[k] = 0.f;
  }
  r[11] = 0.f;
}

// ---- per-stage device timing of the view runtime (bench.py `stages`: ff.knn / ff.disc / ff.new_slots / ff.merge_pool / ff.zone_pool /
// ff.result_copy): CUDA events on the launch stream around each phase, only while d3d_ff_profile_begin() ... _end() is active ----
enum { PF_KNN = 0, PF_DISC, PF_NEW, PF_MERGE, PF_ZONE, PF_RESULT, PF_N };
struct ProfRec { int stage; cudaEvent_t e0, e1; double work; };
struct Prof { bool on = false; std::vector<ProfRec> recs; } g_prof;
struct ProfScope {
  bool on; int stage; double work; cudaStream_t st; cudaEvent_t e0 = nullptr;
  ProfScope(int stage_, double work_, cudaStream_t st_) : on(g_prof.on), stage(stage_), work(work_), st(st_) {
    if (on && cudaEventCreate(&e0) == cudaSuccess) cudaEventRecord(e0, st); else on = false;
  }
  ~ProfScope() {
    if (!on) return;
    cudaEvent_t e1 = nullptr;
    if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return; }
    cudaEventRecord(e1, st);
    g_prof.recs.push_back({stage, e0, e1, work});
  }
};

// bump allocator over the upload ring: returns the offset of `bytes` (16-byte aligned) in both the pinned host and the device mirror
struct Stage {
  FFH& H; const d3d_ff_runtime& rt; size_t begin, end;
  Stage(FFH& h, const d3d_ff_runtime& r) : H(h), rt(r) {
    // a view's uploads are < 1/4 of the ring and every view ends with a host wait on the device, so a region is never
    // rewritten while an earlier copy or a kernel that reads it (the deferred zone pass: issued one view later) is in flight
    if (H.stage_cursor + rt.stage_bytes / 4 > rt.stage_bytes) H.stage_cursor = 0;
    begin = end = H.stage_cursor;
  }
  template <typename T>
  size_t put(const T* src, size_t n) {
    const size_t off = end;
    if (n) memcpy((char*)rt.stage_host + off, src, n * sizeof(T));
    end += (n * sizeof(T) + 15) / 16 * 16;
    return off;
  }
  bool fits() const { return end <= rt.stage_bytes && end - begin <= rt.stage_bytes / 4; }
  int flush(cudaStream_t st) {
    if (end > begin) D3D_CHECK_CUDA(cudaMemcpyAsync((char*)rt.stage_dev + begin, (char*)rt.stage_host + begin, end - begin, cudaMemcpyHostToDevice, st));
    H.stage_cursor = end;
    return 0;
  }
  template <typename T>
  T* dev(size_t off) const { return (T*)((char*)rt.stage_dev + off); }
};

int run_deferred(FFH& H, const d3d_ff_runtime& rt, void* stream) {
  if (!H.zone.pending) return 0;
  const FFH::Deferred& z = H.zone;
  ProfScope pf(PF_ZONE, (double)z.T * 29.5e6, (cudaStream_t)stream);  // 4->768 MLP + 2 encoder layers: ~29.5 MFLOP per token
  D3D_TRY(d3d_pool_tokens(rt.level_zone, z.ptrs, z.centre, z.tok_seq, z.tok_src, z.cu, z.T, z.n_seq, z.max_len, 1, 1, rt.workspace, rt.workspace_bytes,
                          z.out, stream));
  D3D_TRY(d3d_scatter_rows_ptr(z.out, rt.level_zone->d_model, nullptr, z.fts_dst, z.n_seq, rt.level_zone->d_model, stream));
  D3D_TRY(d3d_scatter_rows_ptr(z.centre, 3, nullptr, z.pos_dst, z.n_seq, 3, stream));
  H.zone.pending = false;
  return 0;
}

}  // namespace

extern "C" int d3d_ff_run_deferred(void* h, const d3d_ff_runtime* rt, void* stream) { return run_deferred(HH(h), *rt, stream); }

extern "C" int d3d_ff_profile_begin(void) {
  for (auto& r : g_prof.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.recs.clear();
  g_prof.on = true;
  return 0;
}
// ms / work / launches: [6] arrays indexed knn, disc, new_slots, merge_pool, zone_pool, result_copy.  Synchronises the device.
extern "C" int d3d_ff_profile_end(float* ms, double* work, int* launches) {
  g_prof.on = false;
  D3D_CHECK_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < PF_N; ++i) { ms[i] = 0.f; work[i] = 0.0; launches[i] = 0; }
  for (auto& r : g_prof.recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { ms[r.stage] += t; work[r.stage] += r.work; launches[r.stage] += 1; }
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  g_prof.recs.clear();
  return 0;
}
extern "C" void* d3d_event_create(void) {
  cudaEvent_t e = nullptr;
  if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { d3d_set_error("cudaEventCreate failed"); return nullptr; }
  return (void*)e;
}
extern "C" void d3d_event_destroy(void* e) { if (e) cudaEventDestroy((cudaEvent_t)e); }

extern "C" int d3d_ff_view_pre(void* h, int ix, const d3d_ff_runtime* rt_p, const d3d_ff_pools* pools, const float* centres_step,
                               const float* view_fts_step, int* sizes10, int64_t* after3, void* stream) {
  FFH& H = HH(h);
  const d3d_ff_runtime& rt = *rt_p;
  D3D_REQUIRE(ix >= 0 && ix < (int)H.step_splits.size(), "view index");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = (int)H.eps.size();
  H.splits = H.step_splits[(size_t)ix];
  H.seq_start = H.step_seq_start[(size_t)ix];
  H.res_base = H.seq_start[0];
  const int s_lo = H.seq_start[0], n_seq = H.seq_start[(size_t)B] - s_lo;
  D3D_REQUIRE(n_seq <= rt.max_seq, "more sequences in a view than the runtime buffers hold");
  const int Dm = rt.level_inst->d_model;
  std::vector<int64_t> pos_ptr((size_t)n_seq), fts_ptr((size_t)n_seq);
  std::vector<int> nref((size_t)n_seq);
  bool any = false;
  for (int b = 0; b < B; ++b) {
    const Episode& ep = H.eps[(size_t)b];
    const int nr = ep.tree ? (int)ep.n_inst : 0;
    any |= nr > 0;
    for (int s = H.seq_start[(size_t)b]; s < H.seq_start[(size_t)b + 1]; ++s) {
      nref[(size_t)(s - s_lo)] = nr;
      pos_ptr[(size_t)(s - s_lo)] = pools[b].inst_pos;
      fts_ptr[(size_t)(s - s_lo)] = pools[b].inst_fts;
    }
  }
  const float* centres = centres_step + (size_t)s_lo * 3;
  const float* view_fts = view_fts_step + (size_t)s_lo * Dm;
  if (any) {
    Stage sg(H, rt);
    const size_t o_pos = sg.put(pos_ptr.data(), (size_t)n_seq), o_fts = sg.put(fts_ptr.data(), (size_t)n_seq), o_nr = sg.put(nref.data(), (size_t)n_seq);
    D3D_REQUIRE(sg.fits(), "upload ring too small");
    D3D_TRY(sg.flush(st));
    {
      double knn_bytes = 0;  // 12 B per reference slot searched + query / result rows
      for (int s = 0; s < n_seq; ++s) knn_bytes += 12.0 * nref[(size_t)s] + 12.0 + 16.0;
      ProfScope pf(PF_KNN, knn_bytes, st);
      D3D_TRY(d3d_knn2_batched(sg.dev<int64_t>(o_pos), sg.dev<int>(o_nr), centres, n_seq, rt.d2, rt.idx, stream));
    }
    {
      ProfScope pf(PF_DISC, 2.0 * (2.0 * n_seq) * ((double)rt.disc->k_pad * rt.disc->d_hidden + (double)rt.disc->d_hidden * rt.disc->d_out), st);
      D3D_TRY(d3d_disc_input_batched(sg.dev<int64_t>(o_fts), sg.dev<int64_t>(o_pos), rt.idx, view_fts, centres, n_seq, 2, Dm, rt.disc->k_pad, rt.disc_in,
                                     rt.disc->kind, stream));
      D3D_TRY(d3d_mlp_ln_gelu(rt.disc, rt.disc_in, rt.disc->k_pad, 2 * n_seq, rt.disc_h32, rt.disc_h16, rt.disc_out, 4, stream));
    }
  }
  {
    ProfScope pf(PF_RESULT, (double)n_seq * 12 * sizeof(float) * 2, st);
    pack_view_result_kernel<<<d3d_cdiv(n_seq, 128), 128, 0, st>>>(centres, rt.d2, rt.idx, rt.disc_out, n_seq, any ? 1 : 0, rt.res_dev);
    D3D_CHECK_LAUNCH();
    D3D_CHECK_CUDA(cudaMemcpyAsync(rt.res_host, rt.res_dev, (size_t)n_seq * 12 * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  D3D_CHECK_CUDA(cudaEventRecord((cudaEvent_t)rt.event, st));
  D3D_TRY(run_deferred(H, rt, stream));  // the previous view's zone pass executes while the host waits for / plans this view
  D3D_CHECK_CUDA(cudaEventSynchronize((cudaEvent_t)rt.event));
  return d3d_ffh_finish_view(h, rt.res_host, sizes10, after3);
}

extern "C" int d3d_ff_view_post(void* h, const d3d_ff_runtime* rt_p, const d3d_ff_pools* pools, float* centres_step, float* view_fts_step,
                                void* stream) {
  FFH& H = HH(h);
  const d3d_ff_runtime& rt = *rt_p;
  cudaStream_t st = (cudaStream_t)stream;
  const ViewPlan& pl = H.plan;
  const int Dm = rt.level_inst->d_model;
  const int n_new = (int)pl.new_src.size(), n_mg = (int)pl.mg_owner.size(), n_zn = (int)pl.zn_owner.size();
  D3D_REQUIRE(n_mg <= rt.max_seq && n_zn <= rt.max_seq, "more merged instances / zones than the runtime buffers hold");
  auto toks = [](const std::vector<int>& len, const std::vector<int>& mem, std::vector<int>& src, std::vector<int>& seq, std::vector<int>& cu,
                 int& max_len) {
    src.clear(); seq.clear(); cu.assign(1, 0);
    max_len = 0;
    size_t m = 0;
    for (size_t s = 0; s < len.size(); ++s) {
      src.push_back(-1); seq.push_back((int)s);
      for (int i = 0; i < len[s]; ++i) { src.push_back(mem[m++]); seq.push_back((int)s); }
      cu.push_back((int)src.size());
      max_len = std::max(max_len, len[s] + 1);
    }
  };
  Stage sg(H, rt);
  // ---- new instances: view token / centroid -> their slots (FF:643-648) ----
  size_t o_ns = 0, o_nf = 0, o_np = 0;
  if (n_new) {
    std::vector<int64_t> fd((size_t)n_new), pd((size_t)n_new);
    for (int i = 0; i < n_new; ++i) {
      const d3d_ff_pools& p = pools[pl.new_owner[(size_t)i]];
      fd[(size_t)i] = p.inst_fts + 4LL * Dm * pl.new_iid[(size_t)i];
      pd[(size_t)i] = p.inst_pos + 12LL * pl.new_iid[(size_t)i];
    }
    o_ns = sg.put(pl.new_src.data(), (size_t)n_new); o_nf = sg.put(fd.data(), (size_t)n_new); o_np = sg.put(pd.data(), (size_t)n_new);
  }
  // ---- merged instances: re-encode ALL member patches (FF:658-688) ----
  size_t o_ms = 0, o_mq = 0, o_mc = 0, o_mp = 0, o_mx = 0, o_mf = 0, o_md = 0;
  int t_mg = 0, ml_mg = 0;
  if (n_mg) {
    std::vector<int> src, seq, cu;
    toks(pl.mg_len, pl.mg_members, src, seq, cu, ml_mg);
    t_mg = (int)src.size();
    std::vector<int64_t> ptrs((size_t)4 * n_mg), fd((size_t)n_mg), pd((size_t)n_mg);
    for (int i = 0; i < n_mg; ++i) {
      const d3d_ff_pools& p = pools[pl.mg_owner[(size_t)i]];
      ptrs[(size_t)i] = p.patch_pos; ptrs[(size_t)n_mg + i] = p.patch_dir; ptrs[(size_t)2 * n_mg + i] = p.patch_scale; ptrs[(size_t)3 * n_mg + i] = p.patch_fts;
      fd[(size_t)i] = p.inst_fts + 4LL * Dm * pl.mg_iid[(size_t)i];
      pd[(size_t)i] = p.inst_pos + 12LL * pl.mg_iid[(size_t)i];
    }
    o_ms = sg.put(src.data(), src.size()); o_mq = sg.put(seq.data(), seq.size()); o_mc = sg.put(cu.data(), cu.size());
    o_mp = sg.put(ptrs.data(), ptrs.size()); o_mx = sg.put(pl.mg_pos.data(), pl.mg_pos.size());
    o_mf = sg.put(fd.data(), fd.size()); o_md = sg.put(pd.data(), pd.size());
  }
  // ---- zones (FF:693-756): uploaded now, issued behind the next view's result copy ----
  size_t o_zs = 0, o_zq = 0, o_zc = 0, o_zp = 0, o_zx = 0, o_zf = 0, o_zd = 0;
  int t_zn = 0, ml_zn = 0;
  if (n_zn) {
    std::vector<int> src, seq, cu;
    toks(pl.zn_len, pl.zn_members, src, seq, cu, ml_zn);
    t_zn = (int)src.size();
    // Q5: an updated zone is embedded from its members' voxel-centre keys, derived on the device from the instance positions
    // (pool_features_kernel: row 1 of the pointer table = 1 selects it, row 2 carries the voxel length as float bits)
    int64_t len_bits = 0;
    { const float L = H.zone_len; int32_t bits; memcpy(&bits, &L, 4); len_bits = (int64_t)(uint32_t)bits; }
    std::vector<int64_t> ptrs((size_t)4 * n_zn), fd((size_t)n_zn), pd((size_t)n_zn);
    for (int i = 0; i < n_zn; ++i) {
      const int b = pl.zn_owner[(size_t)i];
      const d3d_ff_pools& p = pools[b];
      ptrs[(size_t)i] = p.inst_pos; ptrs[(size_t)n_zn + i] = pl.zn_keys[(size_t)i] ? 1 : 0; ptrs[(size_t)2 * n_zn + i] = len_bits;
      ptrs[(size_t)3 * n_zn + i] = p.inst_fts;
      fd[(size_t)i] = p.zone_fts + 4LL * Dm * pl.zn_slot[(size_t)i];
      pd[(size_t)i] = p.zone_pos + 12LL * pl.zn_slot[(size_t)i];
    }
    o_zs = sg.put(src.data(), src.size()); o_zq = sg.put(seq.data(), seq.size()); o_zc = sg.put(cu.data(), cu.size());
    o_zp = sg.put(ptrs.data(), ptrs.size()); o_zx = sg.put(pl.zn_pos.data(), pl.zn_pos.size());
    o_zf = sg.put(fd.data(), fd.size()); o_zd = sg.put(pd.data(), pd.size());
  }
  D3D_REQUIRE(sg.fits(), "upload ring too small for this view's plan");
  D3D_TRY(sg.flush(st));
  if (n_new) {
    ProfScope pf(PF_NEW, (double)n_new * (Dm * 4 + 12) * 2, st);
    D3D_TRY(d3d_scatter_rows_ptr(view_fts_step, Dm, sg.dev<int>(o_ns), sg.dev<int64_t>(o_nf), n_new, Dm, stream));
    D3D_TRY(d3d_scatter_rows_ptr(centres_step, 3, sg.dev<int>(o_ns), sg.dev<int64_t>(o_np), n_new, 3, stream));
  }
  if (n_mg) {
    ProfScope pf(PF_MERGE, (double)t_mg * 29.5e6, st);  // re-encode of ALL member patches of every merged instance (FF:662-688)
    D3D_TRY(d3d_pool_tokens(rt.level_inst, sg.dev<int64_t>(o_mp), sg.dev<float>(o_mx), sg.dev<int>(o_mq), sg.dev<int>(o_ms), sg.dev<int>(o_mc), t_mg, n_mg,
                            ml_mg, 0, 0, rt.workspace, rt.workspace_bytes, rt.out_merge, stream));
    D3D_TRY(d3d_scatter_rows_ptr(rt.out_merge, Dm, nullptr, sg.dev<int64_t>(o_mf), n_mg, Dm, stream));
    D3D_TRY(d3d_scatter_rows_ptr(sg.dev<float>(o_mx), 3, nullptr, sg.dev<int64_t>(o_md), n_mg, 3, stream));
  }
  if (n_zn) {
    FFH::Deferred& z = H.zone;
    z.pending = true;
    z.ptrs = sg.dev<int64_t>(o_zp); z.centre = sg.dev<float>(o_zx); z.tok_seq = sg.dev<int>(o_zq); z.tok_src = sg.dev<int>(o_zs); z.cu = sg.dev<int>(o_zc);
    z.fts_dst = sg.dev<int64_t>(o_zf); z.pos_dst = sg.dev<int64_t>(o_zd); z.out = rt.out_zone; z.T = t_zn; z.n_seq = n_zn; z.max_len = ml_zn;
  }
  return 0;
}
