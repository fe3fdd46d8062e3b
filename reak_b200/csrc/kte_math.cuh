// kte_math.cuh — vectors, quaternions and rotation matrices of the interpreter kernels and the proximity code
// (included inside the includer's anonymous namespace after it has defined GD; also compiled by NVRTC as part of
// the sources rkb_prox_jit.cu generates, and for the host by tests/host_build).
#pragma once

struct V3 { double x, y, z; };
GD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
GD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
GD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
GD V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
GD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GD V3 ldv(const double* p) { return v3(p[0], p[1], p[2]); }

struct Q4 { double w, x, y, z; };
GD Q4 qmul(Q4 a, Q4 b) {  // rotations_3D.hpp:1093-1098
  Q4 r;
  r.w = b.w * a.w - b.x * a.x - b.y * a.y - b.z * a.z;
  r.x = b.w * a.x + b.z * a.y - b.y * a.z + b.x * a.w;
  r.y = b.w * a.y - b.z * a.x + b.x * a.z + b.y * a.w;
  r.z = b.w * a.z + b.y * a.x - b.x * a.y + b.z * a.w;
  return r;
}
GD Q4 qconj(Q4 a) { Q4 r; r.w = a.w; r.x = -a.x; r.y = -a.y; r.z = -a.z; return r; }
struct M3 { double m[9]; };  // row-major, v_parent = R v_local
GD M3 qrot(Q4 q) {  // quaternion::getRotMat, rotations_3D.hpp:986-1000
  const double t01 = 2.0 * q.w * q.x, t02 = 2.0 * q.w * q.y, t03 = 2.0 * q.w * q.z;
  const double t11 = 2.0 * q.x * q.x, t12 = 2.0 * q.x * q.y, t13 = 2.0 * q.x * q.z;
  const double t22 = 2.0 * q.y * q.y, t23 = 2.0 * q.y * q.z, t33 = 2.0 * q.z * q.z;
  M3 R;
  R.m[0] = 1.0 - t22 - t33; R.m[1] = t12 - t03; R.m[2] = t02 + t13;
  R.m[3] = t12 + t03; R.m[4] = 1.0 - t11 - t33; R.m[5] = t23 - t01;
  R.m[6] = t13 - t02; R.m[7] = t01 + t23; R.m[8] = 1.0 - t11 - t22;
  return R;
}
GD V3 mul(const M3& R, V3 v) {
  return v3(R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z, R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z);
}
GD V3 tmul(const M3& R, V3 v) {
  return v3(R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z, R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z);
}
GD V3 unit_axis(V3 a) {  // axis_angle ctor, rotations_3D.hpp:1962-1974
  const double n = sqrt(dot(a, a));
  if (n > 0.0000001) return v3(a.x / n, a.y / n, a.z / n);
  return v3(1.0, 0.0, 0.0);
}
GD M3 aa_rot(double angle, V3 a) {  // axis_angle::getRotMat, rotations_3D.hpp:2159-2178 (a normalised)
  double sa, ca;
  sincos(angle, &sa, &ca);
  const double omc = 1.0 - ca;
  const double t12 = omc * a.x * a.y, t13 = omc * a.x * a.z, t23 = omc * a.y * a.z;
  const double t01 = sa * a.x, t02 = sa * a.y, t03 = sa * a.z;
  M3 R;
  R.m[0] = ca + omc * a.x * a.x; R.m[1] = t12 - t03; R.m[2] = t13 + t02;
  R.m[3] = t12 + t03; R.m[4] = ca + omc * a.y * a.y; R.m[5] = t23 - t01;
  R.m[6] = t13 - t02; R.m[7] = t23 + t01; R.m[8] = ca + omc * a.z * a.z;
  return R;
}

// planar vectors and rotations
struct V2 { double x, y; };
GD V2 v2(double x, double y) { V2 r; r.x = x; r.y = y; return r; }
GD V2 operator+(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
GD V2 operator-(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
GD V2 operator*(double s, V2 a) { return v2(s * a.x, s * a.y); }
GD double dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
GD double cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }   // vect_alg.hpp:1142
GD V2 crs(double s, V2 v) { return v2(-v.y * s, v.x * s); }     // vect_alg.hpp:1171
struct R2 { double c, s; };
GD V2 rmul(R2 R, V2 v) { return v2(v.x * R.c - v.y * R.s, v.x * R.s + v.y * R.c); }    // rotations_2D.hpp:292
GD V2 rtmul(R2 R, V2 v) { return v2(v.x * R.c + v.y * R.s, v.y * R.c - v.x * R.s); }   // rotations_2D.hpp:300 (v * R)
GD R2 rr(R2 a, R2 b) { R2 r; r.c = a.c * b.c - a.s * b.s; r.s = a.s * b.c + a.c * b.s; return r; }
