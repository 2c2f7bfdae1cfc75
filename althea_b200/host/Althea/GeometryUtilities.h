// Host-side mirror of Include/Althea/GeometryUtilities.h:32-70: flat normals and tangent-space generation for
// de-indexed triangle lists (three vertices per face), which is what Primitive.cpp:147-191 feeds it when a glTF
// primitive ships without NORMAL or TANGENT.
//
// The reference hands the work to the MikkTSpace library (Extern/MikkTSpace, genTangSpaceDefault, 180 degree
// threshold) and keeps the "basic" result: one unit tangent and a handedness per face corner, bitangent =
// sign * cross(normal, tangent). This header computes the same tangent space from scratch, as published by the
// algorithm's author (M. Mikkelsen, "Simulation of Wrinkled Surfaces Revisited", 2008) and as observable from the
// reference's build of the library (tests/test_tangent_space.py compares against it, oracle/_ref):
//   1. corners with identical position, normal and uv are one vertex;
//   2. every triangle gets a first-order tangent / bitangent from its uv mapping, and an orientation flag (sign of the
//      uv area); triangles with a vanishing uv area or tangent are "wild" and adopt the orientation of whichever
//      neighbourhood reaches them first;
//   3. triangles are neighbours across an edge when they traverse it in opposite directions;
//   4. around each vertex, the corners reachable through neighbour edges with one orientation form a group;
//   5. a corner's tangent is the angle-weighted sum of its group's triangle tangents, each projected into the plane of
//      the vertex normal, normalised;
//   6. triangles with coincident positions take the tangent space of any healthy triangle sharing the vertex.
// Arithmetic is fp32 in the order the library uses, so results agree to rounding and the flags agree exactly.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace AltheaEngine {
namespace tangent_space_detail {

struct F3 {
  float x, y, z;
};
inline F3 operator+(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline F3 operator*(float s, F3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(F3 a) { return sqrtf(dot(a, a)); }
inline bool same(F3 a, F3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool nonzero(float v) { return fabsf(v) > FLT_MIN; }
inline bool nonzero(F3 a) { return nonzero(a.x) || nonzero(a.y) || nonzero(a.z); }
inline F3 unit(F3 a) { return (1.0f / length(a)) * a; }
inline F3 unitIfNonzero(F3 a) { return nonzero(a) ? unit(a) : a; }
// component of a in the plane perpendicular to the unit vector n, normalised when it does not vanish
inline F3 projectedUnit(F3 a, F3 n) { return unitIfNonzero(a - dot(n, a) * n); }

struct Triangle {
  uint32_t corner0;          // index of the first of its three corners in the caller's arrays
  uint32_t v[3];             // welded vertex ids
  int32_t across[3];         // neighbour over edge (v[i], v[i+1]), -1 when open
  int32_t group[3];          // group of each corner, -1 when none
  F3 os, ot;                 // first-order tangent / bitangent, unit, flipped for mirrored mappings
  float magS, magT;
  bool preserving, wild, collapsed;
};

struct Space {
  F3 os{1.0f, 0.0f, 0.0f};
  F3 ot{0.0f, 1.0f, 0.0f};
  bool preserving = false;
};

struct Mesh {
  const float* position;
  const float* normal;
  const float* uv;
  F3 P(uint32_t c) const { return {position[3 * c], position[3 * c + 1], position[3 * c + 2]}; }
  F3 N(uint32_t c) const { return {normal[3 * c], normal[3 * c + 1], normal[3 * c + 2]}; }
  float U(uint32_t c) const { return uv[2 * c]; }
  float V(uint32_t c) const { return uv[2 * c + 1]; }
};

struct CornerKey {
  uint32_t w[8];
  bool operator==(const CornerKey& o) const { return std::memcmp(w, o.w, sizeof w) == 0; }
};
struct CornerKeyHash {
  size_t operator()(const CornerKey& k) const {
    uint64_t h = 0x9e3779b97f4a7c15ull;
    for (uint32_t x : k.w) h = (h ^ x) * 0x100000001b3ull + (h >> 29);
    return (size_t)h;
  }
};

// step 1: one id per distinct (position, normal, uv); the id is the first corner carrying those values
inline std::vector<uint32_t> weld(const Mesh& m, size_t corners) {
  std::vector<uint32_t> id(corners);
  std::unordered_map<CornerKey, uint32_t, CornerKeyHash> seen;
  seen.reserve(corners * 2);
  for (size_t c = 0; c < corners; ++c) {
    float f[8] = {m.position[3 * c], m.position[3 * c + 1], m.position[3 * c + 2], m.normal[3 * c],
                  m.normal[3 * c + 1], m.normal[3 * c + 2], m.uv[2 * c], m.uv[2 * c + 1]};
    CornerKey k;
    for (int i = 0; i < 8; ++i) {
      float v = f[i] == 0.0f ? 0.0f : f[i];  // -0 and +0 compare equal
      std::memcpy(&k.w[i], &v, 4);
    }
    id[c] = seen.emplace(k, (uint32_t)c).first->second;
  }
  return id;
}

// step 2
inline void firstOrder(const Mesh& m, Triangle& t) {
  const uint32_t a = t.v[0], b = t.v[1], c = t.v[2];
  const float s1 = m.U(b) - m.U(a), t1 = m.V(b) - m.V(a);
  const float s2 = m.U(c) - m.U(a), t2 = m.V(c) - m.V(a);
  const F3 e1 = m.P(b) - m.P(a), e2 = m.P(c) - m.P(a);
  const float area2 = s1 * t2 - t1 * s2;
  const F3 os = t2 * e1 - t1 * e2;
  const F3 ot = (-s2) * e1 + s1 * e2;
  t.preserving = area2 > 0.0f;
  t.wild = true;
  t.os = t.ot = F3{0.0f, 0.0f, 0.0f};
  t.magS = t.magT = 0.0f;
  if (!nonzero(area2)) return;
  const float absArea = fabsf(area2), lenS = length(os), lenT = length(ot);
  const float sign = t.preserving ? 1.0f : -1.0f;
  if (nonzero(lenS)) t.os = (sign / lenS) * os;
  if (nonzero(lenT)) t.ot = (sign / lenT) * ot;
  t.magS = lenS / absArea;
  t.magT = lenT / absArea;
  if (nonzero(t.magS) && nonzero(t.magT)) t.wild = false;
}

// step 3: among the triangles on one undirected edge, taken in list order, each still-open edge pairs with the first
// later one that runs the other way and is still open (so a fin with three or more triangles pairs deterministically)
inline void linkNeighbours(std::vector<Triangle>& tris, size_t healthy) {
  struct Side {
    uint32_t lo, hi, tri;
    uint8_t edge, forward;
  };
  std::vector<Side> sides;
  sides.reserve(healthy * 3);
  for (uint32_t f = 0; f < healthy; ++f)
    for (int e = 0; e < 3; ++e) {
      uint32_t a = tris[f].v[e], b = tris[f].v[(e + 1) % 3];
      sides.push_back({std::min(a, b), std::max(a, b), f, (uint8_t)e, (uint8_t)(a < b)});
    }
  std::sort(sides.begin(), sides.end(), [](const Side& p, const Side& q) {
    if (p.lo != q.lo) return p.lo < q.lo;
    if (p.hi != q.hi) return p.hi < q.hi;
    if (p.tri != q.tri) return p.tri < q.tri;
    return p.edge < q.edge;
  });
  for (size_t i = 0; i < sides.size(); ++i) {
    const Side& s = sides[i];
    if (tris[s.tri].across[s.edge] != -1) continue;
    for (size_t j = i + 1; j < sides.size() && sides[j].lo == s.lo && sides[j].hi == s.hi; ++j) {
      const Side& o = sides[j];
      if (o.forward != s.forward && tris[o.tri].across[o.edge] == -1) {
        tris[s.tri].across[s.edge] = (int32_t)o.tri;
        tris[o.tri].across[o.edge] = (int32_t)s.tri;
        break;
      }
    }
  }
}

struct Group {
  uint32_t vertex;
  bool preserving;
  std::vector<uint32_t> members;
};

// step 4: flood around `vertex` through the two edges that meet there, first the edge leaving the corner, then the edge
// arriving at it, depth first; a wild triangle takes the orientation of the first group that reaches it
inline void flood(std::vector<Triangle>& tris, Group& g, int32_t groupIndex, int32_t first, int32_t second) {
  std::vector<int32_t> pending;
  if (second >= 0) pending.push_back(second);
  if (first >= 0) pending.push_back(first);
  while (!pending.empty()) {
    Triangle& t = tris[pending.back()];
    const uint32_t ti = (uint32_t)pending.back();
    pending.pop_back();
    int c = t.v[0] == g.vertex ? 0 : t.v[1] == g.vertex ? 1 : 2;
    if (t.group[c] != -1) continue;
    if (t.wild && t.group[0] == -1 && t.group[1] == -1 && t.group[2] == -1) t.preserving = g.preserving;
    if (t.preserving != g.preserving) continue;
    g.members.push_back(ti);
    t.group[c] = groupIndex;
    const int32_t leaving = t.across[c], arriving = t.across[(c + 2) % 3];
    if (arriving >= 0) pending.push_back(arriving);
    if (leaving >= 0) pending.push_back(leaving);
  }
}

// step 5 for one set of triangles around `vertex`
inline Space blend(const Mesh& m, const std::vector<Triangle>& tris, const std::vector<uint32_t>& members, uint32_t vertex) {
  Space r;
  r.os = r.ot = F3{0.0f, 0.0f, 0.0f};
  const F3 n = m.N(vertex);
  for (uint32_t ti : members) {
    const Triangle& t = tris[ti];
    if (t.wild) continue;
    const int c = t.v[0] == vertex ? 0 : t.v[1] == vertex ? 1 : 2;
    const F3 os = projectedUnit(t.os, n), ot = projectedUnit(t.ot, n);
    const F3 here = m.P(t.v[c]);
    const F3 toPrev = projectedUnit(m.P(t.v[(c + 2) % 3]) - here, n);
    const F3 toNext = projectedUnit(m.P(t.v[(c + 1) % 3]) - here, n);
    float cosine = dot(toPrev, toNext);
    cosine = cosine > 1.0f ? 1.0f : (cosine < -1.0f ? -1.0f : cosine);
    const float angle = (float)acos((double)cosine);
    r.os = r.os + angle * os;
    r.ot = r.ot + angle * ot;
  }
  r.os = unitIfNonzero(r.os);
  r.ot = unitIfNonzero(r.ot);
  return r;
}

// tangentOut: 3 floats per corner, signOut: +1 / -1 per corner
inline void generate(const float* position, const float* normal, const float* uv, size_t triangleCount,
                     float* tangentOut, float* signOut) {
  const Mesh m{position, normal, uv};
  const size_t corners = triangleCount * 3;
  const std::vector<uint32_t> id = weld(m, corners);

  // healthy triangles first, in their original order; collapsed ones (two coincident positions) after them
  std::vector<Triangle> tris;
  tris.reserve(triangleCount);
  for (int pass = 0; pass < 2; ++pass)
    for (size_t f = 0; f < triangleCount; ++f) {
      const uint32_t c0 = (uint32_t)(3 * f);
      const F3 p0 = m.P(c0), p1 = m.P(c0 + 1), p2 = m.P(c0 + 2);
      const bool collapsed = same(p0, p1) || same(p0, p2) || same(p1, p2);
      if (collapsed != (pass == 1)) continue;
      Triangle t{};
      t.corner0 = c0;
      for (int i = 0; i < 3; ++i) {
        t.v[i] = id[c0 + i];
        t.across[i] = -1;
        t.group[i] = -1;
      }
      t.collapsed = collapsed;
      tris.push_back(t);
    }
  size_t healthy = 0;
  while (healthy < tris.size() && !tris[healthy].collapsed) ++healthy;

  for (size_t f = 0; f < healthy; ++f) firstOrder(m, tris[f]);
  linkNeighbours(tris, healthy);

  std::vector<Group> groups;
  for (uint32_t f = 0; f < healthy; ++f)
    for (int c = 0; c < 3; ++c) {
      if (tris[f].wild || tris[f].group[c] != -1) continue;
      const int32_t gi = (int32_t)groups.size();
      groups.push_back(Group{tris[f].v[c], tris[f].preserving, {f}});
      tris[f].group[c] = gi;
      flood(tris, groups.back(), gi, tris[f].across[c], tris[f].across[(c + 2) % 3]);
    }

  std::vector<Space> spaces(corners);
  const float threshold = (float)cos((double)((180.0f * (float)M_PI) / 180.0f));  // the library's default: 180 degrees
  std::vector<uint32_t> subset, whole;
  for (size_t gi = 0; gi < groups.size(); ++gi) {
    const Group& g = groups[gi];
    const F3 n = m.N(g.vertex);
    whole = g.members;
    std::sort(whole.begin(), whole.end());
    bool haveWhole = false;
    Space wholeSpace;
    for (uint32_t f : g.members) {
      const Triangle& tf = tris[f];
      const int c = tf.group[0] == (int32_t)gi ? 0 : tf.group[1] == (int32_t)gi ? 1 : 2;
      const F3 os = projectedUnit(tf.os, n), ot = projectedUnit(tf.ot, n);
      // a corner blends with the group's triangles whose tangent and bitangent are not exactly opposed to its own
      subset.clear();
      for (uint32_t o : whole) {
        const Triangle& to = tris[o];
        const bool wild = tf.wild || to.wild;
        const F3 os2 = projectedUnit(to.os, n), ot2 = projectedUnit(to.ot, n);
        if (wild || o == f || (dot(os, os2) > threshold && dot(ot, ot2) > threshold)) subset.push_back(o);
      }
      Space s;
      if (subset.size() == whole.size()) {
        if (!haveWhole) {
          wholeSpace = blend(m, tris, whole, g.vertex);
          haveWhole = true;
        }
        s = wholeSpace;
      } else {
        s = blend(m, tris, subset, g.vertex);
      }
      s.preserving = g.preserving;
      spaces[tf.corner0 + c] = s;
    }
  }

  // step 6: the first healthy corner, in list order, that is the same welded vertex
  if (healthy < tris.size()) {
    std::unordered_map<uint32_t, uint32_t> firstCorner;
    firstCorner.reserve(healthy * 3);
    for (size_t f = 0; f < healthy; ++f)
      for (int c = 0; c < 3; ++c) firstCorner.emplace(tris[f].v[c], tris[f].corner0 + c);
    for (size_t f = healthy; f < tris.size(); ++f)
      for (int c = 0; c < 3; ++c) {
        auto it = firstCorner.find(tris[f].v[c]);
        if (it != firstCorner.end()) spaces[tris[f].corner0 + c] = spaces[it->second];
      }
  }

  for (size_t c = 0; c < corners; ++c) {
    tangentOut[3 * c] = spaces[c].os.x;
    tangentOut[3 * c + 1] = spaces[c].os.y;
    tangentOut[3 * c + 2] = spaces[c].os.z;
    signOut[c] = spaces[c].preserving ? 1.0f : -1.0f;
  }
}

} // namespace tangent_space_detail

class GeometryUtilities {
public:
  // Include/Althea/GeometryUtilities.h:32-48. TVertex has float position[3], normal[3] (Model.h's Vertex).
  template <typename TVertex> static void computeFlatNormals(std::vector<TVertex>& vertices) {
    using tangent_space_detail::F3;
    for (size_t i = 0; i < vertices.size() / 3; ++i) {
      TVertex& a = vertices[3 * i];
      TVertex& b = vertices[3 * i + 1];
      TVertex& c = vertices[3 * i + 2];
      const F3 ab{b.position[0] - a.position[0], b.position[1] - a.position[1], b.position[2] - a.position[2]};
      const F3 ac{c.position[0] - a.position[0], c.position[1] - a.position[1], c.position[2] - a.position[2]};
      F3 n{ab.y * ac.z - ab.z * ac.y, ab.z * ac.x - ab.x * ac.z, ab.x * ac.y - ab.y * ac.x};
      n = tangent_space_detail::unit(n);
      for (TVertex* v : {&a, &b, &c}) {
        v->normal[0] = n.x;
        v->normal[1] = n.y;
        v->normal[2] = n.z;
      }
    }
  }

  // Include/Althea/GeometryUtilities.h:51-70,137-155: tangent from the generator, bitangent = sign * cross(normal, tangent).
  // TVertex has float position[3], tangent[3], bitangent[3], normal[3], uvs[4][2].
  template <typename TVertex> static void computeTangentSpace(std::vector<TVertex>& vertices, uint32_t uvIndex) {
    const size_t faces = vertices.size() / 3, corners = faces * 3;
    std::vector<float> p(corners * 3), n(corners * 3), t(corners * 2), tang(corners * 3), sign(corners);
    for (size_t c = 0; c < corners; ++c) {
      for (int k = 0; k < 3; ++k) {
        p[3 * c + k] = vertices[c].position[k];
        n[3 * c + k] = vertices[c].normal[k];
      }
      t[2 * c] = vertices[c].uvs[uvIndex][0];
      t[2 * c + 1] = vertices[c].uvs[uvIndex][1];
    }
    if (faces == 0) return;
    tangent_space_detail::generate(p.data(), n.data(), t.data(), faces, tang.data(), sign.data());
    for (size_t c = 0; c < corners; ++c) {
      TVertex& v = vertices[c];
      for (int k = 0; k < 3; ++k) v.tangent[k] = tang[3 * c + k];
      const float* N = v.normal;
      const float* T = v.tangent;
      v.bitangent[0] = sign[c] * (N[1] * T[2] - N[2] * T[1]);
      v.bitangent[1] = sign[c] * (N[2] * T[0] - N[0] * T[2]);
      v.bitangent[2] = sign[c] * (N[0] * T[1] - N[1] * T[0]);
    }
  }
};

} // namespace AltheaEngine
