// kte_proximity2d.cuh — minimum distance between two planar proximity models (included by kte_generic.cu
// inside its anonymous namespace after kte_math.cuh: uses V2, R2, rmul, rtmul of that file).
//
// What it stands in for (paths relative to ReaK's source tree):
//   proxy_query_pair_2D::createProxFinderList / findMinimumDistance   geometry/proximity/proxy_query_model.cpp:73-195
//   the pair finders                                                   geometry/proximity/prox_circle_circle.cpp, prox_circle_crect.cpp,
//                                                                      prox_circle_rectangle.cpp, prox_crect_crect.cpp,
//                                                                      prox_crect_rectangle.cpp, prox_rectangle_rectangle.cpp
//   shape poses (anchor frame, then the shape's own pose)              geometry/shapes/geometry_2D.cpp, core/kinetostatics/pose_2D.hpp:98-200
// The finders are followed branch for branch, including where they are not the geometric answer: the overlap
// tests of parallel capped rectangles and of axis-parallel lines are an `||` of two conditions that cannot both
// fail (prox_crect_crect.cpp:57-58, prox_crect_rectangle.cpp:48-49, :87-88), rectangle-rectangle only looks at corners and
// never reports a negative distance (prox_rectangle_rectangle.cpp:42-76), a line end inside the rectangle's slab takes the
// face distance (prox_crect_rectangle.cpp:148-154).  A planner's accept / reject decisions depend on them.
#pragma once

struct Pose2 { V2 p; R2 R; };          // world pose: position, (cos, sin)
struct ProxRecord2 { V2 p1, p2; double d; };

GD V2 to_global2(const Pose2& P, V2 v) { return P.p + rmul(P.R, v); }      // pose_2D.hpp:177-186
GD V2 from_global2(const Pose2& P, V2 v) { return rtmul(P.R, v - P.p); }   // pose_2D.hpp:191-200 ((V - Position) * Rotation)
GD double norm2v(V2 a) { return sqrt(a.x * a.x + a.y * a.y); }

// prox_circle_circle.cpp:40-57
GD ProxRecord2 prox_circle_circle(const Pose2& C1, double r1, const Pose2& C2, double r2) {
  ProxRecord2 R;
  const V2 c1 = C1.p, c2 = C2.p;
  const V2 diff = c2 - c1;
  const double dist = norm2v(diff);
  R.d = dist - r1 - r2;
  R.p1 = c1 + (r1 / dist) * diff;
  R.p2 = c2 - (r2 / dist) * diff;
  return R;
}

// prox_circle_crect.cpp:40-86 (dims: length along x, width = cap diameter)
GD ProxRecord2 prox_circle_crect(const Pose2& CI, double r, const Pose2& CR, V2 dims) {
  ProxRecord2 R;
  const V2 rel = from_global2(CR, CI.p);
  const bool in_x = (rel.x > -0.5 * dims.x) && (rel.x < 0.5 * dims.x);
  if (in_x) {
    if (rel.y > 0.0) {
      R.p1 = to_global2(CR, v2(rel.x, rel.y - r));
      R.p2 = to_global2(CR, v2(rel.x, 0.5 * dims.y));
      R.d = rel.y - r - 0.5 * dims.y;
    } else {
      R.p1 = to_global2(CR, v2(rel.x, rel.y + r));
      R.p2 = to_global2(CR, v2(rel.x, -0.5 * dims.y));
      R.d = -0.5 * dims.y - rel.y - r;
    }
    return R;
  }
  V2 endc = v2(0.0, 0.0);
  if (rel.x > 0.0) endc.x += 0.5 * dims.x;
  else endc.x -= 0.5 * dims.x;
  const V2 dv = rel - endc;
  const double dd = norm2v(dv);
  R.p1 = to_global2(CR, rel - (r / dd) * dv);
  R.p2 = to_global2(CR, endc + (0.5 * dims.y / dd) * dv);
  R.d = dd - 0.5 * dims.y - r;
  return R;
}

// the point of a rectangle's boundary a point is referred to: prox_circle_rectangle.cpp:55-79,
// prox_rectangle_rectangle.cpp:42-76 (the same code twice in the reference)
GD V2 rect_corner_pt(V2 rel, V2 dims) {
  bool in_x = (rel.x > -0.5 * dims.x) && (rel.x < 0.5 * dims.x);
  bool in_y = (rel.y > -0.5 * dims.y) && (rel.y < 0.5 * dims.y);
  if (in_x && in_y) {
    const double bx = 0.5 * dims.x - fabs(rel.x), by = 0.5 * dims.y - fabs(rel.y);
    if (bx <= by) in_x = false;
    else in_y = false;
  }
  V2 c = v2(0.5 * dims.x, 0.5 * dims.y);
  if (in_x) c.x = rel.x;
  else if (rel.x < 0.0) c.x = -c.x;
  if (in_y) c.y = rel.y;
  else if (rel.y < 0.0) c.y = -c.y;
  return c;
}

// prox_circle_rectangle.cpp:40-86
GD ProxRecord2 prox_circle_rectangle(const Pose2& CI, double r, const Pose2& RE, V2 dims) {
  ProxRecord2 R;
  const V2 ci_c = CI.p;
  const V2 rel = from_global2(RE, ci_c);
  R.p2 = to_global2(RE, rect_corner_pt(rel, dims));
  const V2 dv = R.p2 - ci_c;
  const double dd = norm2v(dv);
  R.p1 = ci_c + (r / dd) * dv;
  R.d = dd - r;
  return R;
}

// prox_crect_crect.cpp:40-127
GD ProxRecord2 prox_crect_crect(const Pose2& C1, V2 d1, const Pose2& C2, V2 d2) {
  ProxRecord2 R;
  const V2 c_rel = from_global2(C1, C2.p);
  const V2 t_rel = rtmul(C1.R, rmul(C2.R, v2(1.0, 0.0)));
  if (fabs(t_rel.y) < 1e-5) {
    if ((c_rel.x + 0.5 * d2.x > -0.5 * d1.x) || (c_rel.x - 0.5 * d2.x < 0.5 * d1.x)) {
      const double max_x = (c_rel.x + 0.5 * d2.x < 0.5 * d1.x) ? (c_rel.x + 0.5 * d2.x) : (0.5 * d1.x);
      const double min_x = (c_rel.x - 0.5 * d2.x > -0.5 * d1.x) ? (c_rel.x - 0.5 * d2.x) : (-0.5 * d1.x);
      const double avg_x = (max_x + min_x) * 0.5;
      const double ry = c_rel.y < 0.0 ? -1.0 : 1.0;
      R.p1 = to_global2(C1, v2(avg_x, 0.5 * d1.y * ry));
      R.p2 = to_global2(C1, v2(avg_x, c_rel.y - 0.5 * d2.y * ry));
      R.d = fabs(c_rel.y) - 0.5 * d1.y - 0.5 * d2.y;
      return R;
    }
    V2 a = v2(0.0, 0.0), b = c_rel;  // (unreachable in the reference too: the test above cannot fail)
    if (c_rel.x < 0.0) { a.x -= 0.5 * d1.x; b.x += 0.5 * d2.x; }
    else { a.x += 0.5 * d1.x; b.x -= 0.5 * d2.x; }
    const V2 dv = b - a;
    const double dd = norm2v(dv);
    R.p1 = to_global2(C1, a + (0.5 * d1.y / dd) * dv);
    R.p2 = to_global2(C1, b - (0.5 * d2.y / dd) * dv);
    R.d = dd - 0.5 * d1.y - 0.5 * d2.y;
    return R;
  }
  const double d = dot(t_rel, c_rel);
  const double denom = 1.0 - t_rel.x * t_rel.x;
  double s_c = (t_rel.x * c_rel.x - d) / denom;
  double t_c = (c_rel.x - t_rel.x * d) / denom;
  if (s_c < -0.5 * d2.x) { s_c = -0.5 * d2.x; t_c = c_rel.x - 0.5 * d2.x * t_rel.x; }
  else if (s_c > 0.5 * d2.x) { s_c = 0.5 * d2.x; t_c = c_rel.x + 0.5 * d2.x * t_rel.x; }
  if (t_c < -0.5 * d1.x) { t_c = -0.5 * d1.x; s_c = -0.5 * d1.x * t_rel.x - d; }
  else if (t_c > 0.5 * d1.x) { t_c = 0.5 * d1.x; s_c = 0.5 * d1.x * t_rel.x - d; }
  if (s_c < -0.5 * d2.x) s_c = -0.5 * d2.x;
  else if (s_c > 0.5 * d2.x) s_c = 0.5 * d2.x;
  const V2 a = v2(t_c, 0.0);
  const V2 b = c_rel + s_c * t_rel;
  const V2 dv = b - a;
  const double dd = norm2v(dv);
  R.p1 = to_global2(C1, a + (0.5 * d1.y / dd) * dv);
  R.p2 = to_global2(C1, b - (0.5 * d2.y / dd) * dv);
  R.d = dd - 0.5 * d1.y - 0.5 * d2.y;
  return R;
}

// prox_crect_rectangle.cpp:40-178: the centre line of the capped rectangle against the rectangle
GD ProxRecord2 rect_line(const Pose2& RE, V2 dims, V2 ln_c, V2 ln_t, double half) {
  ProxRecord2 R;
  const V2 c = from_global2(RE, ln_c);
  const V2 t = rtmul(RE.R, ln_t);
  if (fabs(t.x) < 1e-5) {  // vertical
    if ((c.y + half > -0.5 * dims.y) || (c.y - half < 0.5 * dims.y)) {
      const double max_y = (c.y + half < 0.5 * dims.y) ? (c.y + half) : (0.5 * dims.y);
      const double min_y = (c.y - half > -0.5 * dims.y) ? (c.y - half) : (-0.5 * dims.y);
      const double avg_y = (max_y + min_y) * 0.5;
      const double rx = c.x < 0.0 ? -1.0 : 1.0;
      R.p1 = to_global2(RE, v2(c.x, avg_y));
      R.p2 = to_global2(RE, v2(0.5 * dims.x * rx, avg_y));
      R.d = fabs(c.x) - 0.5 * dims.x;
      return R;
    }
    V2 re_pt = v2(0.0, 0.0), ln_pt = c;  // (unreachable, as in the reference)
    if (c.x < 0.0) re_pt.x -= 0.5 * dims.x; else re_pt.x += 0.5 * dims.x;
    if (c.y < 0.0) { re_pt.y -= 0.5 * dims.y; ln_pt.y += half; }
    else { re_pt.y += 0.5 * dims.y; ln_pt.y -= half; }
    R.p1 = to_global2(RE, ln_pt);
    R.p2 = to_global2(RE, re_pt);
    R.d = norm2v(ln_pt - re_pt);
    return R;
  }
  if (fabs(t.y) < 1e-5) {  // horizontal
    if ((c.x + half > -0.5 * dims.x) || (c.x - half < 0.5 * dims.x)) {
      const double max_x = (c.x + half < 0.5 * dims.x) ? (c.x + half) : (0.5 * dims.x);
      const double min_x = (c.x - half > -0.5 * dims.x) ? (c.x - half) : (-0.5 * dims.x);
      const double avg_x = (max_x + min_x) * 0.5;
      const double ry = c.y < 0.0 ? -1.0 : 1.0;
      R.p1 = to_global2(RE, v2(avg_x, c.y));
      R.p2 = to_global2(RE, v2(avg_x, 0.5 * dims.y * ry));
      R.d = fabs(c.y) - 0.5 * dims.y;
      return R;
    }
    V2 re_pt = v2(0.0, 0.0), ln_pt = c;  // (unreachable, as in the reference)
    if (c.y < 0.0) re_pt.y -= 0.5 * dims.y; else re_pt.y += 0.5 * dims.y;
    if (c.x < 0.0) { re_pt.x -= 0.5 * dims.x; ln_pt.x += half; }
    else { re_pt.x += 0.5 * dims.x; ln_pt.x -= half; }
    R.p1 = to_global2(RE, ln_pt);
    R.p2 = to_global2(RE, re_pt);
    R.d = norm2v(ln_pt - re_pt);
    return R;
  }
  V2 n = crs(1.0, t);  // 1.0 % ln_t_rel, vect_alg.hpp:1171
  if (dot(n, c) < 0.0) n = v2(-n.x, -n.y);
  V2 corner = v2(-0.5 * dims.x, -0.5 * dims.y);
  if (n.x > 0.0) corner.x = 0.5 * dims.x;
  if (n.y > 0.0) corner.y = 0.5 * dims.y;
  const V2 cd = c - corner;
  double dist = dot(cd, n);
  double tt = -dot(cd, t);
  if (fabs(tt) > half) {
    tt = tt < 0.0 ? -half : half;
    const V2 ln_pt = c + tt * t;
    const double in_x = fabs(ln_pt.x) - 0.5 * dims.x;
    const double in_y = fabs(ln_pt.y) - 0.5 * dims.y;
    if ((in_x < 0.0) && (in_y > in_x)) {
      corner.x = ln_pt.x;
      dist = fabs(ln_pt.y) - 0.5 * dims.y;
    } else if ((in_y < 0.0) && (in_x > in_y)) {
      corner.y = ln_pt.y;
      dist = fabs(ln_pt.x) - 0.5 * dims.x;
    } else {
      corner.x = ln_pt.x < 0.0 ? -0.5 * dims.x : 0.5 * dims.x;
      corner.y = ln_pt.y < 0.0 ? -0.5 * dims.y : 0.5 * dims.y;
      dist = norm2v(ln_pt - corner);
    }
    R.p1 = to_global2(RE, ln_pt);
    R.p2 = to_global2(RE, corner);
    R.d = dist;
  } else {
    R.p1 = to_global2(RE, corner + dist * n);
    R.p2 = to_global2(RE, corner);
    R.d = dist;
  }
  return R;
}

// prox_crect_rectangle.cpp:181-207: the line solution, then a circle swept along it
GD ProxRecord2 prox_crect_rectangle(const Pose2& CR, V2 dcr, const Pose2& RE, V2 dre) {
  const V2 cr_t = rmul(CR.R, v2(1.0, 0.0));
  ProxRecord2 R = rect_line(RE, dre, CR.p, cr_t, 0.5 * dcr.x);
  const V2 dv = R.p2 - R.p1;
  const double dd = norm2v(dv);
  if (R.d < 0.0) R.p1 = R.p1 - (0.5 * dcr.y / dd) * dv;
  else R.p1 = R.p1 + (0.5 * dcr.y / dd) * dv;
  R.d -= 0.5 * dcr.y;
  return R;
}

// prox_rectangle_rectangle.cpp:78-160: the four corners of each against the other, first strict minimum wins
GD ProxRecord2 prox_rectangle_rectangle(const Pose2& R1, V2 d1, const Pose2& R2, V2 d2) {
  ProxRecord2 R;
  R.d = INFINITY;
  R.p1 = v2(0.0, 0.0);
  R.p2 = v2(0.0, 0.0);
  V2 corner = v2(0.5 * d2.x, 0.5 * d2.y);
  for (int k = 0; k < 4; ++k) {
    if (k == 1 || k == 3) corner.y = -corner.y;
    if (k == 2) corner.x = -corner.x;
    const V2 g = to_global2(R2, corner);
    const V2 pt = to_global2(R1, rect_corner_pt(from_global2(R1, g), d1));
    const double dist = norm2v(pt - g);
    if (dist < R.d) { R.d = dist; R.p1 = pt; R.p2 = g; }
  }
  corner = v2(0.5 * d1.x, 0.5 * d1.y);
  for (int k = 0; k < 4; ++k) {
    if (k == 1 || k == 3) corner.y = -corner.y;
    if (k == 2) corner.x = -corner.x;
    const V2 g = to_global2(R1, corner);
    const V2 pt = to_global2(R2, rect_corner_pt(from_global2(R2, g), d2));
    const double dist = norm2v(pt - g);
    if (dist < R.d) { R.d = dist; R.p2 = pt; R.p1 = g; }
  }
  return R;
}

// createProxFinderList (proxy_query_model.cpp:73-160): every pair of planar shapes has a finder; the kind listed first in
// circle > capped rectangle > rectangle takes the first slot, on equal kinds model 1's shape.
GD ProxRecord2 prox_compute2(const ProxShape& a, const Pose2& Pa, const ProxShape& b, const Pose2& Pb) {
  const bool swap = b.kind < a.kind;
  const ProxShape& s1 = swap ? b : a;
  const ProxShape& s2 = swap ? a : b;
  const Pose2& P1 = swap ? Pb : Pa;
  const Pose2& P2 = swap ? Pa : Pb;
  const V2 d1 = v2(s1.dims[0], s1.dims[1]), d2 = v2(s2.dims[0], s2.dims[1]);
  if (s1.kind == RKB_SHAPE_CIRCLE) {
    if (s2.kind == RKB_SHAPE_CIRCLE) return prox_circle_circle(P1, d1.x, P2, d2.x);
    if (s2.kind == RKB_SHAPE_CRECT) return prox_circle_crect(P1, d1.x, P2, d2);
    return prox_circle_rectangle(P1, d1.x, P2, d2);
  }
  if (s1.kind == RKB_SHAPE_CRECT) {
    if (s2.kind == RKB_SHAPE_CRECT) return prox_crect_crect(P1, d1, P2, d2);
    return prox_crect_rectangle(P1, d1, P2, d2);
  }
  return prox_rectangle_rectangle(P1, d1, P2, d2);
}

// world pose of a planar shape riding on a frame: pose_2D::getGlobalPose, pose_2D.hpp:98-106.
// ProxShape of a planar shape: pos[0..1], quat[0..1] = (cos, sin) of its own rotation.
GD Pose2 prox_shape_pose2(const ProxProgram& P, const ProxShape& S, const Pose2* slots) {
  Pose2 G;
  const V2 lp = v2(S.pos[0], S.pos[1]);
  R2 lr;
  lr.c = S.quat[0]; lr.s = S.quat[1];
  if (S.anchor < 0) {
    G.p = lp;
    G.R = lr;
  } else {
    const Pose2 F = slots[P.slot_of[S.anchor]];
    G.p = F.p + rmul(F.R, lp);
    G.R = rr(F.R, lr);
  }
  return G;
}

// proxy_query_pair_2D::findMinimumDistance (proxy_query_model.cpp:163-190): as the 3D one
GD int prox_min_distance2(const ProxProgram& P, const Pose2* slots, ProxRecord2& bestR) {
  int f = 0, best = -1;
  bestR.p1 = v2(0.0, 0.0); bestR.p2 = v2(0.0, 0.0); bestR.d = INFINITY;
  for (int a = 0; a < P.n1; ++a) {
    const ProxShape& Sa = P.s[a];
    const Pose2 Pa = prox_shape_pose2(P, Sa, slots);
    for (int b = 0; b < P.n2; ++b) {
      const ProxShape& Sb = P.s[P.n1 + b];
      const Pose2 Pb = prox_shape_pose2(P, Sb, slots);
      // (getShape1 / getShape2 of the finder may be swapped with respect to (a, b): the test is symmetric up to the sign of
      // the difference vector, whose norm is what is read)
      const bool swap = Sb.kind < Sa.kind;
      const V2 p1 = swap ? Pb.p : Pa.p, p2 = swap ? Pa.p : Pb.p;
      const double r1 = swap ? Sb.brad : Sa.brad, r2 = swap ? Sa.brad : Sb.brad;
      if (f > 0 && norm2v(p2 - p1) - r1 - r2 > bestR.d) { ++f; continue; }
      const ProxRecord2 R = prox_compute2(Sa, Pa, Sb, Pb);
      if (f == 0 || R.d < bestR.d) { best = f; bestR = R; }
      ++f;
    }
  }
  return best;
}

// proxy_query_pair_2D::gatherCollisionPoints (proxy_query_model.cpp:192-212)
template <class Store>
GD int prox_gather_collisions2(const ProxProgram& P, const Pose2* slots, int max_records, Store store) {
  int f = 0, n = 0;
  for (int a = 0; a < P.n1; ++a) {
    const ProxShape& Sa = P.s[a];
    const Pose2 Pa = prox_shape_pose2(P, Sa, slots);
    for (int b = 0; b < P.n2; ++b) {
      const ProxShape& Sb = P.s[P.n1 + b];
      const Pose2 Pb = prox_shape_pose2(P, Sb, slots);
      const bool swap = Sb.kind < Sa.kind;
      const V2 p1 = swap ? Pb.p : Pa.p, p2 = swap ? Pa.p : Pb.p;
      const double r1 = swap ? Sb.brad : Sa.brad, r2 = swap ? Sa.brad : Sb.brad;
      if (!(norm2v(p2 - p1) - r1 - r2 > 0.0)) {
        const ProxRecord2 R = prox_compute2(Sa, Pa, Sb, Pb);
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
