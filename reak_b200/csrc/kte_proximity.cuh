// kte_proximity.cuh — minimum distance between two proximity models (included by kte_generic.cu
// inside its anonymous namespace: uses V3, Q4, qmul, qconj of that file).
//
// What it stands in for (paths relative to ReaK's source tree):
//   proxy_query_pair_3D::createProxFinderList / findMinimumDistance   geometry/proximity/proxy_query_model.cpp:212-412
//   the pair finders                                                   geometry/proximity/prox_*_*.cpp, prox_fundamentals_3D.cpp
//   shape poses (anchor frame, then the shape's own pose)              geometry/shapes/geometry_3D.cpp:33-52,
//                                                                      core/kinetostatics/pose_3D.hpp:102-198
// which is the test manip_dk_proxy_env_impl::is_free (ctrl/topologies/manip_free_workspace.hpp:77-99)
// and the steering loops (examples/misc/MEAQR_topology.hpp:921-940) run on every propagated state.
// The finders are followed branch for branch, including where they are not the geometric answer
// (the always-true overlap test of parallel capsules, plane_box using the box x axis three times,
// planes treated as unbounded by the sphere / cylinder / box finders but bounded by the culling
// test): the planner's accept / reject decisions depend on them.
#pragma once

struct Pose { V3 p; Q4 q; };

GD V3 qrotv(Q4 Q, V3 V) {  // quaternion * vect, rotations_3D.hpp:1137-1151
  const double t0 = Q.w * Q.x, t1 = Q.w * Q.y, t2 = Q.w * Q.z, t3 = -Q.x * Q.x, t4 = Q.x * Q.y, t5 = Q.x * Q.z, t6 = -Q.y * Q.y,
               t7 = Q.y * Q.z, t8 = -Q.z * Q.z;
  return v3(2.0 * ((t6 + t8) * V.x + (t4 - t2) * V.y + (t1 + t5) * V.z) + V.x,
            2.0 * ((t2 + t4) * V.x + (t3 + t8) * V.y + (t7 - t0) * V.z) + V.y,
            2.0 * ((t5 - t1) * V.x + (t0 + t7) * V.y + (t3 + t6) * V.z) + V.z);
}
// World pose of a shape.  A finder rotates with the same quaternion five to ten times, so the nine sums of
// products that quaternion * vect forms first (rotations_3D.hpp:1139-1150) are kept: m = (t6+t8, t4-t2, t1+t5;
// t2+t4, t3+t8, t7-t0; t5-t1, t0+t7, t3+t6).  For invert(Q) = (w, -x, -y, -z) the t0..t2 change sign and the
// sums become the transposed set, bit for bit — so both directions come from one table and the results
// are the ones qrotv gives.
struct SPose { V3 p; double m[9]; };
GD void rot_table(Q4 Q, double* m) {
  const double t0 = Q.w * Q.x, t1 = Q.w * Q.y, t2 = Q.w * Q.z, t3 = -Q.x * Q.x, t4 = Q.x * Q.y, t5 = Q.x * Q.z, t6 = -Q.y * Q.y,
               t7 = Q.y * Q.z, t8 = -Q.z * Q.z;
  m[0] = t6 + t8; m[1] = t4 - t2; m[2] = t1 + t5;
  m[3] = t2 + t4; m[4] = t3 + t8; m[5] = t7 - t0;
  m[6] = t5 - t1; m[7] = t0 + t7; m[8] = t3 + t6;
}
GD V3 rot_fwd(const double* m, V3 V) {
  return v3(2.0 * (m[0] * V.x + m[1] * V.y + m[2] * V.z) + V.x, 2.0 * (m[3] * V.x + m[4] * V.y + m[5] * V.z) + V.y,
            2.0 * (m[6] * V.x + m[7] * V.y + m[8] * V.z) + V.z);
}
GD V3 rot_inv(const double* m, V3 V) {
  return v3(2.0 * (m[0] * V.x + m[3] * V.y + m[6] * V.z) + V.x, 2.0 * (m[1] * V.x + m[4] * V.y + m[7] * V.z) + V.y,
            2.0 * (m[2] * V.x + m[5] * V.y + m[8] * V.z) + V.z);
}
GD V3 to_global(const SPose& P, V3 v) { return P.p + rot_fwd(P.m, v); }              // pose_3D.hpp:175-184
GD V3 from_global(const SPose& P, V3 v) { return rot_inv(P.m, v - P.p); }            // pose_3D.hpp:189-198
GD V3 rot_to_global(const SPose& P, V3 v) { return rot_fwd(P.m, v); }
GD V3 rot_from_global(const SPose& P, V3 v) { return rot_inv(P.m, v); }
GD double norm3(V3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
GD V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }

struct ProxRecord { V3 p1, p2; double d; };

// prox_sphere_sphere.cpp:50-66
template <bool PTS>
GD ProxRecord prox_sphere_sphere(const SPose& S1, double r1, const SPose& S2, double r2) {
  ProxRecord R;
  const V3 c1 = S1.p, c2 = S2.p;
  const V3 diff = c2 - c1;
  const double dist = norm3(diff);
  R.d = dist - r1 - r2;
  if (PTS) {
    R.p1 = c1 + (r1 / dist) * diff;
    R.p2 = c2 - (r2 / dist) * diff;
  }
  return R;
}

// prox_sphere_ccylinder.cpp:50-88
template <bool PTS>
GD ProxRecord prox_sphere_ccylinder(const SPose& S, double rs, const SPose& C, double len, double rc) {
  ProxRecord R;
  const V3 sp_c = S.p;
  const V3 rel = from_global(C, sp_c);
  if (fabs(rel.z) <= 0.5 * len) {
    const V3 proj = v3(rel.x, rel.y, 0.0);
    const double pd = norm3(proj);
    if (PTS) {
      R.p2 = to_global(C, v3(0.0, 0.0, rel.z) + (rc / pd) * proj);
      R.p1 = to_global(C, rel - (rs / pd) * proj);
    }
    R.d = pd - rs - rc;
  } else {
    const double fact = rel.z < 0.0 ? -1.0 : 1.0;
    const V3 c2 = to_global(C, v3(0.0, 0.0, fact * 0.5 * len));
    const V3 diff = c2 - sp_c;
    const double dist = norm3(diff);
    R.d = dist - rs - rc;
    if (PTS) {
      R.p1 = sp_c + (rs / dist) * diff;
      R.p2 = c2 - (rc / dist) * diff;
    }
  }
  return R;
}

// prox_sphere_cylinder.cpp:50-103
template <bool PTS>
GD ProxRecord prox_sphere_cylinder(const SPose& S, double rs, const SPose& C, double len, double rc) {
  ProxRecord R;
  const V3 sp_c = S.p;
  const V3 rel = from_global(C, sp_c);
  const double rad = sqrt(rel.x * rel.x + rel.y * rel.y);
  if (fabs(rel.z) <= 0.5 * len) {
    const V3 proj = v3(rel.x, rel.y, 0.0);
    const double pd = norm3(proj);
    if (PTS) {
      R.p2 = to_global(C, v3(0.0, 0.0, rel.z) + (rc / pd) * proj);
      R.p1 = to_global(C, rel - (rs / pd) * proj);
    }
    R.d = pd - rs - rc;
  } else if (rad < rc) {
    const double fact = rel.z < 0.0 ? -1.0 : 1.0;
    if (PTS) {
      R.p2 = to_global(C, v3(rel.x, rel.y, fact * 0.5 * len));
      R.p1 = to_global(C, v3(rel.x, rel.y, rel.z - fact * rs));
    }
    R.d = fact * rel.z - 0.5 * len - rs;
  } else {
    V3 proj = v3(rel.x, rel.y, 0.0);
    double pd = norm3(proj);
    const double fact = rel.z < 0.0 ? -1.0 : 1.0;
    const V3 rim = (rc / pd) * proj + v3(0.0, 0.0, fact * 0.5 * len);
    R.p2 = to_global(C, rim);
    proj = R.p2 - sp_c;
    pd = norm3(proj);
    if (PTS) R.p1 = sp_c + (rs / pd) * proj;
    R.d = pd - rs;
  }
  return R;
}

// prox_plane_sphere.cpp:128-144 (the live definition; the bounded-plane variant above it is commented out)
template <bool PTS>
GD ProxRecord prox_plane_sphere(const SPose& PL, const SPose& S, double rs) {
  ProxRecord R;
  const V3 rel = from_global(PL, S.p);
  if (PTS) {
    R.p1 = to_global(PL, v3(rel.x, rel.y, 0.0));
    R.p2 = to_global(PL, v3(rel.x, rel.y, rel.z - rs));
  }
  R.d = rel.z - rs;
  return R;
}

// prox_plane_ccylinder.cpp:51-81
template <bool PTS>
GD ProxRecord prox_plane_ccylinder(const SPose& PL, const SPose& C, double len, double rc) {
  ProxRecord R;
  const V3 cy_t = rot_to_global(C, v3(0.0, 0.0, 1.0));
  const V3 c_rel = from_global(PL, C.p);
  V3 t_rel = rot_from_global(PL, cy_t);
  if (fabs(t_rel.z) < 1e-6) {
    if (PTS) {
      R.p1 = to_global(PL, v3(c_rel.x, c_rel.y, 0.0));
      R.p2 = to_global(PL, v3(c_rel.x, c_rel.y, c_rel.z - rc));
    }
    R.d = c_rel.z - rc;
  } else {
    if (t_rel.z > 0.0) t_rel = neg(t_rel);
    const V3 pt = c_rel + (0.5 * len) * t_rel + v3(0.0, 0.0, -rc);
    if (PTS) {
      R.p1 = to_global(PL, v3(pt.x, pt.y, 0.0));
      R.p2 = to_global(PL, pt);
    }
    R.d = pt.z;
  }
  return R;
}

// prox_plane_cylinder.cpp:51-88
template <bool PTS>
GD ProxRecord prox_plane_cylinder(const SPose& PL, const SPose& C, double len, double rc) {
  ProxRecord R;
  const V3 cy_t = rot_to_global(C, v3(0.0, 0.0, 1.0));
  const V3 c_rel = from_global(PL, C.p);
  V3 t_rel = rot_from_global(PL, cy_t);
  if (fabs(t_rel.z) < 1e-6) {
    if (PTS) {
      R.p1 = to_global(PL, v3(c_rel.x, c_rel.y, 0.0));
      R.p2 = to_global(PL, v3(c_rel.x, c_rel.y, c_rel.z - rc));
    }
    R.d = c_rel.z - rc;
  } else if (sqrt(t_rel.x * t_rel.x + t_rel.y * t_rel.y) < 1e-6) {
    if (PTS) {
      R.p1 = to_global(PL, v3(c_rel.x, c_rel.y, 0.0));
      R.p2 = to_global(PL, v3(c_rel.x, c_rel.y, c_rel.z - 0.5 * len));
    }
    R.d = c_rel.z - 0.5 * len;
  } else {
    if (t_rel.z > 0.0) t_rel = neg(t_rel);
    // unit(): vect_alg.hpp — v / norm_2(v)
    V3 r_rel = v3(0.0, 0.0, -1.0) + t_rel.z * t_rel;
    const double rn = norm3(r_rel);
    r_rel = v3(r_rel.x / rn, r_rel.y / rn, r_rel.z / rn);
    const V3 pt = c_rel + (0.5 * len) * t_rel + rc * r_rel;
    if (PTS) {
      R.p1 = to_global(PL, v3(pt.x, pt.y, 0.0));
      R.p2 = to_global(PL, pt);
    }
    R.d = pt.z;
  }
  return R;
}

// prox_plane_box.cpp:51-79 — bx_x, bx_y and bx_z are all the image of the box's x axis
template <bool PTS>
GD ProxRecord prox_plane_box(const SPose& PL, const SPose& B, V3 dims) {
  ProxRecord R;
  V3 bx = rot_from_global(PL, rot_to_global(B, v3(1.0, 0.0, 0.0)));
  if (bx.z > 0.0) bx = neg(bx);
  const V3 c_rel = from_global(PL, B.p);
  const V3 pt = c_rel + 0.5 * (dims.x * bx + dims.y * bx + dims.z * bx);
  if (PTS) {
    R.p1 = to_global(PL, v3(pt.x, pt.y, 0.0));
    R.p2 = to_global(PL, pt);
  }
  R.d = pt.z;
  return R;
}

// prox_plane_plane.cpp:44-96
GD void plane_point(const SPose& PL, V3 dims, V3 pt, V3& rec, double& dist) {
  const V3 rel = from_global(PL, pt);
  const bool in_x = rel.x > -0.5 * dims.x && rel.x < 0.5 * dims.x;
  const bool in_y = rel.y > -0.5 * dims.y && rel.y < 0.5 * dims.y;
  if (in_x && in_y) {
    const double fact = rel.z < 0.0 ? -1.0 : 1.0;
    rec = to_global(PL, v3(rel.x, rel.y, 0.0));
    dist = fact * rel.z;
  } else {
    V3 rim;
    if (in_x) rim = v3(rel.x, (rel.y < 0.0 ? -1.0 : 1.0) * 0.5 * dims.y, 0.0);
    else if (in_y) rim = v3((rel.x < 0.0 ? -1.0 : 1.0) * 0.5 * dims.x, rel.y, 0.0);
    else {
      rim = v3(0.5 * dims.x, 0.5 * dims.y, 0.0);
      if (rel.x < 0.0) rim.x = -rim.x;
      if (rel.y < 0.0) rim.y = -rim.y;
    }
    rec = to_global(PL, rim);
    dist = norm3(rec - pt);
  }
}

// prox_plane_plane.cpp:99-197: the four corners of plane 2 against plane 1, then of plane 1 against plane 2
template <bool PTS>
GD ProxRecord prox_plane_plane(const SPose& P1, V3 d1, const SPose& P2, V3 d2) {
  ProxRecord R;
  R.d = INFINITY; R.p1 = v3(0, 0, 0); R.p2 = v3(0, 0, 0);
  for (int side = 0; side < 2; ++side) {
    const SPose& own = side == 0 ? P2 : P1;
    const SPose& other = side == 0 ? P1 : P2;
    const V3 od = side == 0 ? d2 : d1, pd = side == 0 ? d1 : d2;
    V3 corner = v3(0.5 * od.x, 0.5 * od.y, 0.0);
    for (int c = 0; c < 4; ++c) {
      if (c == 1 || c == 3) corner.y = -corner.y;
      if (c == 2) corner.x = -corner.x;
      const V3 g = to_global(own, corner);
      V3 rec; double dist;
      plane_point(other, pd, g, rec, dist);
      if (dist < R.d) {
        R.d = dist;
        if (side == 0) { R.p1 = rec; R.p2 = g; } else { R.p2 = rec; R.p1 = g; }
      }
    }
  }
  return R;
}

// findProximityBoxToPoint, prox_fundamentals_3D.cpp:35-83: p1 on the box, p2 the point
template <bool PTS>
GD ProxRecord box_point(const SPose& B, V3 dims, V3 pt) {
  const V3 rel = from_global(B, pt);
  bool in_x = rel.x > -0.5 * dims.x && rel.x < 0.5 * dims.x;
  bool in_y = rel.y > -0.5 * dims.y && rel.y < 0.5 * dims.y;
  bool in_z = rel.z > -0.5 * dims.z && rel.z < 0.5 * dims.z;
  const bool inside = in_x && in_y && in_z;
  if (inside) {
    const double bx = 0.5 * dims.x - fabs(rel.x), by = 0.5 * dims.y - fabs(rel.y), bz = 0.5 * dims.z - fabs(rel.z);
    if (bx <= by && bx <= bz) in_x = false;
    else if (by <= bx && by <= bz) in_y = false;
    else in_z = false;
  }
  V3 corner = 0.5 * dims;
  if (in_x) corner.x = rel.x; else if (rel.x < 0.0) corner.x = -corner.x;
  if (in_y) corner.y = rel.y; else if (rel.y < 0.0) corner.y = -corner.y;
  if (in_z) corner.z = rel.z; else if (rel.z < 0.0) corner.z = -corner.z;
  ProxRecord R;
  if (PTS) R.p1 = to_global(B, corner);
  const double dd = norm3(corner - rel);
  R.p2 = pt;
  R.d = inside ? -dd : dd;
  return R;
}

// prox_sphere_box.cpp:50-73
template <bool PTS>
GD ProxRecord prox_sphere_box(const SPose& S, double rs, const SPose& B, V3 dims) {
  const ProxRecord b = box_point<PTS>(B, dims, S.p);
  ProxRecord R;
  if (PTS) {
    const V3 diff = b.p1 - b.p2;
    const double dd = norm3(diff);
    if (b.d < 0.0) R.p1 = b.p2 - (rs / dd) * diff;
    else R.p1 = b.p2 + (rs / dd) * diff;
    R.p2 = b.p1;
  }
  R.d = b.d - rs;
  return R;
}

// findProximityBoxToLine, prox_fundamentals_3D.cpp:110-118, with golden_section_search_impl of
// core/optimization/line_search.hpp:71-95 (the bracket flips direction when the probe is not better;
// the record returned is the one of the last evaluation, at the middle of the final bracket).
template <bool PTS>
GD ProxRecord box_line(const SPose& B, V3 dims, V3 centre, V3 tangent, double half) {
  const double phi = 1.618033988;
  double lo = -half, hi = half;
  const double tol = 1e-3 * half;
  double mid = lo + (hi - lo) / phi;
  ProxRecord R = box_point<false>(B, dims, centre + mid * tangent);
  double mid_cost = R.d;
  for (int guard = 0; guard < 200; ++guard) {
    if (fabs(lo - hi) < tol) break;
    const double test = mid + (hi - mid) / phi;
    R = box_point<false>(B, dims, centre + test * tangent);
    if (R.d < mid_cost) { lo = mid; mid = test; mid_cost = R.d; }
    else { hi = lo; lo = test; }
  }
  return box_point<PTS>(B, dims, centre + ((lo + hi) * 0.5) * tangent);
}

// prox_ccylinder_box.cpp:51-75
template <bool PTS>
GD ProxRecord prox_ccylinder_box(const SPose& C, double len, double rc, const SPose& B, V3 dims) {
  const V3 cy_t = rot_to_global(C, v3(0.0, 0.0, 1.0));
  const ProxRecord b = box_line<PTS>(B, dims, C.p, cy_t, 0.5 * len);
  ProxRecord R;
  if (PTS) {
    const V3 diff = b.p1 - b.p2;
    const double dd = norm3(diff);
    if (b.d < 0.0) R.p1 = b.p2 - (rc / dd) * diff;
    else R.p1 = b.p2 + (rc / dd) * diff;
    R.p2 = b.p1;
  }
  R.d = b.d - rc;
  return R;
}

// prox_ccylinder_ccylinder.cpp:43-129
template <bool PTS>
GD ProxRecord prox_ccylinder_ccylinder(const SPose& C1, double len1, double r1, const SPose& C2, double len2, double r2) {
  ProxRecord R;
  const V3 c2 = C2.p;
  const V3 t2 = rot_to_global(C2, v3(0.0, 0.0, 1.0));
  const V3 c = from_global(C1, c2);
  const V3 t = rot_from_global(C1, t2);
  if (sqrt(t.x * t.x + t.y * t.y) < 1e-5) {
    // parallel axes; the overlap test of the reference is an `||` of two conditions that cannot
    // both fail, so this branch always returns here
    if ((c.z + 0.5 * len2 > -0.5 * len1) || (c.z - 0.5 * len2 < 0.5 * len1)) {
      const double max_z = (c.z + 0.5 * len2 < 0.5 * len1) ? (c.z + 0.5 * len2) : (0.5 * len1);
      const double min_z = (c.z - 0.5 * len2 > -0.5 * len1) ? (c.z - 0.5 * len2) : (-0.5 * len1);
      const double avg_z = (max_z + min_z) * 0.5;
      V3 rr = v3(c.x, c.y, 0.0);
      const double rn = norm3(rr);
      rr = v3(rr.x / rn, rr.y / rn, rr.z / rn);
      if (PTS) {
        R.p1 = to_global(C1, v3(r1 * rr.x, r1 * rr.y, avg_z));
        R.p2 = to_global(C1, v3(c.x - r2 * rr.x, c.y - r2 * rr.y, avg_z));
      }
      R.d = sqrt(c.x * c.x + c.y * c.y) - r1 - r2;
      return R;
    }
    V3 s1 = v3(0.0, 0.0, 0.0), s2 = c;
    if (c.z < 0.0) { s1.z -= 0.5 * len1; s2.z += 0.5 * len2; }
    else { s1.z += 0.5 * len1; s2.z -= 0.5 * len2; }
    const V3 diff = s2 - s1;
    const double dist = norm3(diff);
    if (PTS) {
      R.p1 = to_global(C1, s1 + (r1 / dist) * diff);
      R.p2 = to_global(C1, s2 - (r2 / dist) * diff);
    }
    R.d = dist - r1 - r2;
    return R;
  }
  const double d = dot(t, c);
  const double denom = 1.0 - t.z * t.z;
  double s_c = (t.z * c.z - d) / denom;
  double t_c = (c.z - t.z * d) / denom;
  if (s_c < -0.5 * len2) { s_c = -0.5 * len2; t_c = c.z - 0.5 * len2 * t.z; }
  else if (s_c > 0.5 * len2) { s_c = 0.5 * len2; t_c = c.z + 0.5 * len2 * t.z; }
  if (t_c < -0.5 * len1) { t_c = -0.5 * len1; s_c = -0.5 * len1 * t.z - d; }
  else if (t_c > 0.5 * len1) { t_c = 0.5 * len1; s_c = 0.5 * len1 * t.z - d; }
  if (s_c < -0.5 * len2) s_c = -0.5 * len2;
  else if (s_c > 0.5 * len2) s_c = 0.5 * len2;
  const V3 p1c = v3(0.0, 0.0, t_c);
  const V3 p2c = c + s_c * t;
  const V3 diff = p2c - p1c;
  const double dist = norm3(diff);
  if (PTS) {
    R.p1 = to_global(C1, p1c + (r1 / dist) * diff);
    R.p2 = to_global(C1, p2c - (r2 / dist) * diff);
  }
  R.d = dist - r1 - r2;
  return R;
}

// One finder of createProxFinderList (proxy_query_model.cpp:212-384) for shapes a (model 1) and b
// (model 2): the kind that comes first in plane > sphere > capped cylinder takes the first slot.
// Returns false for the pairs the reference has no finder for.
GD bool prox_has_finder(int ka, int kb) {
  const int lo = ka < kb ? ka : kb, hi = ka < kb ? kb : ka;
  if (lo == RKB_SHAPE_PLANE || lo == RKB_SHAPE_SPHERE) return true;
  if (lo == RKB_SHAPE_CCYLINDER) return hi == RKB_SHAPE_CCYLINDER || hi == RKB_SHAPE_BOX;
  return false;
}

template <bool PTS>
GD ProxRecord prox_compute_kd(int ka, V3 da, const SPose& Pa, int kb, V3 db, const SPose& Pb) {
  // first = the shape whose kind is listed first; on equal kinds model 1's shape
  const bool swap = kb < ka;
  const int k1 = swap ? kb : ka, k2 = swap ? ka : kb;
  const SPose& P1 = swap ? Pb : Pa;
  const SPose& P2 = swap ? Pa : Pb;
  const V3 d1 = swap ? db : da, d2 = swap ? da : db;
  if (k1 == RKB_SHAPE_PLANE) {
    if (k2 == RKB_SHAPE_PLANE) return prox_plane_plane<PTS>(P1, d1, P2, d2);
    if (k2 == RKB_SHAPE_SPHERE) return prox_plane_sphere<PTS>(P1, P2, d2.x);
    if (k2 == RKB_SHAPE_CCYLINDER) return prox_plane_ccylinder<PTS>(P1, P2, d2.x, d2.y);
    if (k2 == RKB_SHAPE_CYLINDER) return prox_plane_cylinder<PTS>(P1, P2, d2.x, d2.y);
    return prox_plane_box<PTS>(P1, P2, d2);
  }
  if (k1 == RKB_SHAPE_SPHERE) {
    if (k2 == RKB_SHAPE_SPHERE) return prox_sphere_sphere<PTS>(P1, d1.x, P2, d2.x);
    if (k2 == RKB_SHAPE_CCYLINDER) return prox_sphere_ccylinder<PTS>(P1, d1.x, P2, d2.x, d2.y);
    if (k2 == RKB_SHAPE_CYLINDER) return prox_sphere_cylinder<PTS>(P1, d1.x, P2, d2.x, d2.y);
    return prox_sphere_box<PTS>(P1, d1.x, P2, d2);
  }
  if (k2 == RKB_SHAPE_CCYLINDER) return prox_ccylinder_ccylinder<PTS>(P1, d1.x, d1.y, P2, d2.x, d2.y);
  return prox_ccylinder_box<PTS>(P1, d1.x, d1.y, P2, d2);
}
template <bool PTS>
GD ProxRecord prox_compute(const ProxShape& a, const SPose& Pa, const ProxShape& b, const SPose& Pb) {
  return prox_compute_kd<PTS>(a.kind, v3(a.dims[0], a.dims[1], a.dims[2]), Pa, b.kind, v3(b.dims[0], b.dims[1], b.dims[2]), Pb);
}

// world pose of a shape riding on a frame with pose F: pose_3D::getGlobalPose, pose_3D.hpp:102-110
GD SPose prox_shape_pose(const ProxProgram& P, const ProxShape& S, const Pose* slots) {
  const V3 lp = v3(S.pos[0], S.pos[1], S.pos[2]);
  Q4 lq;
  lq.w = S.quat[0]; lq.x = S.quat[1]; lq.y = S.quat[2]; lq.z = S.quat[3];
  SPose G;
  if (S.anchor < 0) {
    G.p = lp;
    for (int k = 0; k < 9; ++k) G.m[k] = S.rot[k];
  } else {
    const Pose F = slots[P.slot_of[S.anchor]];
    G.p = F.p + qrotv(F.q, lp);
    rot_table(qmul(F.q, lq), G.m);
  }
  return G;
}

// proxy_query_pair_3D::findMinimumDistance, proxy_query_model.cpp:388-412: the first finder is always
// evaluated; a later one is skipped when the distance between the two shape origins minus the two
// bounding radii exceeds the running minimum (for planes that radius is the half diagonal of the
// extents although their finders treat the plane as unbounded — followed as is).
// slots[P.slot_of[f]]: world pose of chain frame f (for the frames shapes ride on).  Returns the finder index, -1 without finders.  The search itself
// only forms distances; the record of the winner (its two points) is evaluated once at the end when wanted.
GD int prox_min_distance(const ProxProgram& P, const Pose* slots, bool want_points, ProxRecord& bestR) {
  int f = 0, best = -1, best_a = 0, best_b = 0;
  double min_d = INFINITY;
  bestR.p1 = v3(0, 0, 0); bestR.p2 = v3(0, 0, 0); bestR.d = INFINITY;
  for (int a = 0; a < P.n1; ++a) {
    const ProxShape& Sa = P.s[a];
    const SPose Pa = prox_shape_pose(P, Sa, slots);
    for (int b = 0; b < P.n2; ++b) {
      const ProxShape& Sb = P.s[P.n1 + b];
      if (!prox_has_finder(Sa.kind, Sb.kind)) continue;
      const SPose Pb = prox_shape_pose(P, Sb, slots);
      if (f > 0 && norm3(Pb.p - Pa.p) - Sa.brad - Sb.brad > min_d) { ++f; continue; }
      const double d = prox_compute<false>(Sa, Pa, Sb, Pb).d;
      if (f == 0 || min_d > d) { best = f; min_d = d; best_a = a; best_b = b; }
      ++f;
    }
  }
  bestR.d = min_d;
  if (best >= 0 && want_points) {
    const ProxShape& Sa = P.s[best_a];
    const ProxShape& Sb = P.s[P.n1 + best_b];
    bestR = prox_compute<true>(Sa, prox_shape_pose(P, Sa, slots), Sb, prox_shape_pose(P, Sb, slots));
  }
  return best;
}

// proxy_query_pair_3D::gatherCollisionPoints (proxy_query_model.cpp:402-421): every finder, in createProxFinderList
// order, whose bounding spheres overlap (origin distance minus the two radii not above 0) is evaluated, and the record
// of each one that reports a negative distance is kept.  rec[r] (r < max_records): d, p1 (3), p2 (3); fnd[r]: the finder
// index.  Returns the number of colliding finders (all of them are counted, the first max_records are stored).
template <class Store>
GD int prox_gather_collisions(const ProxProgram& P, const Pose* slots, int max_records, Store store) {
  int f = 0, n = 0;
  for (int a = 0; a < P.n1; ++a) {
    const ProxShape& Sa = P.s[a];
    const SPose Pa = prox_shape_pose(P, Sa, slots);
    for (int b = 0; b < P.n2; ++b) {
      const ProxShape& Sb = P.s[P.n1 + b];
      if (!prox_has_finder(Sa.kind, Sb.kind)) continue;
      const SPose Pb = prox_shape_pose(P, Sb, slots);
      if (!(norm3(Pb.p - Pa.p) - Sa.brad - Sb.brad > 0.0)) {
        const ProxRecord R = prox_compute<true>(Sa, Pa, Sb, Pb);
        if (R.d < 0.0) {
          if (n < max_records) store(n, f, R);
          ++n;
        }
      }
      ++f;
    }
  }
  return n;
}

