// rkb_rkx.cu — host code: read a kte_nl_system from a ReaK XML archive (`.rkx`) into the flat chain descriptor.
//
// The archive is what ReaK::serialization::xml_oarchive writes (core/serialization/xml_archiver.cpp:380-706; the model
// files of examples/robot_airship/build_P3R3R_model.cpp:78): a tree of <field attr...> ... </field> records, leaf values
// as quoted text, every shared object carrying an object_ID and written in full only where it is met first (later
// references are empty records with the same object_ID).  Nothing of ReaK is needed to read it: classes are recognised by
// their RTTI numbers (the RK_RTTI_MAKE_* lines of ctrl/mbd_kte/*.hpp: 0xC2100004 revolute_joint_3D, ...), fields by the
// names their save() functions give them.  The descriptor is built the way include/reak_b200/reak_bridge.hpp builds it
// from the live objects (same element order, same frame / coordinate numbering), so that "ReaK loads the file, the bridge
// flattens it" and "this reader flattens the file" give the same chain — the parity test of tests/test_rkx.py.
//
// Numbers: the archiver prints doubles with the default stream precision (6 significant digits, xml_archiver.cpp:452);
// what is read is what the file holds, exactly as ReaK's own xml_iarchive reads it (strtod / operator>>).
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/reak_b200.h"
#include "rkb_internal.h"

namespace {

struct Node {
  std::string name;        // field name ("mBase", "mKTEs_q[3]", ...)
  std::string type;        // type_ID attribute, "" for leaves
  long object_id = -1;     // object_ID attribute, -1 when absent
  std::string text;        // leaf value (quotes removed)
  std::vector<Node*> kids;
  ~Node() { for (Node* k : kids) delete k; }
  const Node* kid(const std::string& n) const {
    for (const Node* k : kids) if (k->name == n) return k;
    return nullptr;
  }
};

struct fail : std::runtime_error {
  explicit fail(const std::string& w) : std::runtime_error(w) {}
};

// ---- the record syntax ---------------------------------------------------------------------------
struct Parser {
  const std::string& s;
  size_t i = 0;
  explicit Parser(const std::string& src) : s(src) {}
  void skip_ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\r' || s[i] == '\t')) ++i; }
  static std::string attr(const std::string& tag, const char* key) {
    const std::string k = std::string(key) + "=\"";
    const size_t a = tag.find(k);
    if (a == std::string::npos) return "";
    const size_t b = tag.find('"', a + k.size());
    return b == std::string::npos ? "" : tag.substr(a + k.size(), b - a - k.size());
  }
  // parses one record starting at '<'; returns nullptr at a closing tag
  Node* record() {
    skip_ws();
    if (i >= s.size() || s[i] != '<') throw fail("malformed archive: '<' expected at offset " + std::to_string(i));
    const size_t e = s.find('>', i);
    if (e == std::string::npos) throw fail("malformed archive: unterminated tag");
    const std::string tag = s.substr(i + 1, e - i - 1);
    if (tag.empty()) throw fail("malformed archive: empty tag");
    if (tag[0] == '/') return nullptr;
    i = e + 1;
    Node* n = new Node();
    try {
      const size_t sp = tag.find(' ');
      n->name = tag.substr(0, sp);
      if (sp != std::string::npos) {
        n->type = attr(tag, "type_ID");
        const std::string oid = attr(tag, "object_ID");
        if (!oid.empty()) n->object_id = std::strtol(oid.c_str(), nullptr, 10);
      }
      skip_ws();
      if (i < s.size() && s[i] == '"') {  // leaf: "value"</name>
        const size_t q = s.find('"', i + 1);
        if (q == std::string::npos) throw fail("malformed archive: unterminated value");
        n->text = s.substr(i + 1, q - i - 1);
        i = q + 1;
      } else {
        while (true) {
          Node* k = record();
          if (!k) break;
          n->kids.push_back(k);
        }
      }
      skip_ws();
      const std::string close = "</" + n->name + ">";
      if (s.compare(i, close.size(), close) != 0) throw fail("malformed archive: " + close + " expected");
      i += close.size();
    } catch (...) { delete n; throw; }
    return n;
  }
};

// ---- the object graph ----------------------------------------------------------------------------
struct Archive {
  Node* root = nullptr;
  std::map<long, const Node*> objects;  // object_ID -> the record that holds the object's fields
  ~Archive() { delete root; }
  void index(const Node* n) {
    if (n->object_id > 0 && !n->kids.empty() && !objects.count(n->object_id)) objects[n->object_id] = n;
    for (const Node* k : n->kids) index(k);
  }
  // the record holding the fields of the object a field refers to (itself, or the first full occurrence)
  const Node* deref(const Node* n) const {
    if (!n) return nullptr;
    if (n->object_id > 0) {
      std::map<long, const Node*>::const_iterator it = objects.find(n->object_id);
      if (it != objects.end()) return it->second;
    }
    return n;
  }
};

double number(const Node* n, const char* what) {
  if (!n) throw fail(std::string("field missing: ") + what);
  errno = 0;
  char* end = nullptr;
  const double v = std::strtod(n->text.c_str(), &end);
  if (end == n->text.c_str()) throw fail(std::string("not a number in field ") + what);
  return v;
}
long count_of(const Node* obj, const std::string& base) {
  const Node* c = obj->kid(base + "_count");
  return c ? std::strtol(c->text.c_str(), nullptr, 10) : 0;
}
// vect<double,N>: N <q> leaves
void vect(const Node* n, int dim, double* out, const char* what) {
  if (!n) throw fail(std::string("field missing: ") + what);
  int k = 0;
  for (const Node* c : n->kids) if (c->name == "q" && k < dim) out[k++] = number(c, what);
  if (k != dim) throw fail(std::string("vector of the wrong size in field ") + what);
}
void quat(const Node* n, double* out) {
  if (!n) throw fail("field missing: Quat");
  for (int k = 0; k < 4; ++k) out[k] = number(n->kid("q[" + std::to_string(k) + "]"), "Quat");
}

// class numbers (second argument of RK_RTTI_MAKE_CONCRETE_*; the archive prints them in decimal followed by the
// template arguments, "3255828484.0" = 0xC2100004 = revolute_joint_3D)
unsigned long class_of(const Node* n) { return n ? std::strtoul(n->type.c_str(), nullptr, 10) : 0ul; }
enum : unsigned long {
  T_NL_SYSTEM = 0xC2300002ul, T_CHAIN = 0xC2100002ul, T_MASS_CALC = 0xC2000001ul,
  T_REV_2D = 0xC2100003ul, T_REV_3D = 0xC2100004ul, T_PRI_2D = 0xC2100005ul, T_PRI_3D = 0xC2100006ul, T_FREE_2D = 0xC2100041ul, T_FREE_3D = 0xC2100042ul,
  T_LINK_GEN = 0xC2100007ul, T_LINK_2D = 0xC2100008ul, T_LINK_3D = 0xC2100009ul,
  T_IN_GEN = 0xC210000Aul, T_IN_2D = 0xC210000Bul, T_IN_3D = 0xC210000Cul,
  T_SPRING_GEN = 0xC210000Dul, T_SPRING_2D = 0xC210000Eul, T_SPRING_3D = 0xC210000Ful,
  T_DAMPER_GEN = 0xC2100010ul, T_DAMPER_2D = 0xC2100011ul, T_DAMPER_3D = 0xC2100012ul,
  T_ACTUATOR_GEN = 0xC2100023ul, T_TSPRING_2D = 0xC210002Cul, T_TSPRING_3D = 0xC210002Dul, T_TDAMPER_2D = 0xC210002Eul, T_TDAMPER_3D = 0xC210002Ful
};

struct Builder {
  const Archive& A;
  int dim = 0;
  std::vector<rkb_element> el;
  std::map<long, int> coords, free3, inputs, frames, aux, elem_of;
  std::map<int, const Node*> frame_node;
  std::map<int, int> written;
  int n_coords = 0, n_aux = 0;
  explicit Builder(const Archive& a) : A(a) {}

  int push(int kind, int fa, int fb, int coord, int aux_, uint64_t up, const double* p, int np) {
    rkb_element e;
    std::memset(&e, 0, sizeof e);
    e.kind = kind; e.frame_a = fa; e.frame_b = fb; e.coord = coord; e.aux = aux_; e.upstream = up;
    for (int i = 0; i < np && i < 12; ++i) e.p[i] = p[i];
    el.push_back(e);
    return (int)el.size() - 1;
  }
  long oid(const Node* field, const char* what) const {
    if (!field || field->object_id <= 0) throw fail(std::string("null or missing object in field ") + what);
    return field->object_id;
  }
  int fid(const Node* field, const char* what) {  // frames are numbered in the order the chain walk meets them
    const long id = oid(field, what);
    std::map<long, int>::const_iterator it = frames.find(id);
    if (it != frames.end()) return it->second;
    const int f = (int)frames.size();
    frames[id] = f;
    frame_node[f] = A.deref(field);
    return f;
  }
  int cid(const Node* field, const char* what) const {
    std::map<long, int>::const_iterator it = coords.find(oid(field, what));
    if (it == coords.end()) throw fail("an element refers to a coordinate that is not a system dof");
    return it->second;
  }
  int gid(const Node* field, const char* what) {  // any gen_coord: a dof, or an auxiliary one declared at its first use
    const long id = oid(field, what);
    std::map<long, int>::const_iterator it = coords.find(id);
    if (it != coords.end()) return it->second;
    it = aux.find(id);
    if (it != aux.end()) return it->second;
    const Node* g = A.deref(field);
    const double p[3] = {number(g->kid("q"), "q"), number(g->kid("q_dot"), "q_dot"), number(g->kid("q_ddot"), "q_ddot")};
    const int idx = n_coords + n_aux++;
    push(RKB_COORD_GEN, -1, -1, idx, 0, 0, p, 3);
    aux[id] = idx;
    return idx;
  }
  uint64_t upstream(const Node* dep) const {
    uint64_t m = 0;
    const long n = count_of(dep, "mUpStreamJoints");
    for (long k = 0; k < n; ++k) m |= uint64_t(1) << cid(dep->kid("mUpStreamJoints_key[" + std::to_string(k) + "]"), "mUpStreamJoints");
    // the free joints of a chain are all 2D or all 3D (read_system): one index space
    const char* maps[2] = {"mUpStream2DJoints", "mUpStream3DJoints"};
    for (int w = 0; w < 2; ++w) {
      const long n3 = count_of(dep, maps[w]);
      for (long k = 0; k < n3; ++k) {
        std::map<long, int>::const_iterator it = free3.find(oid(dep->kid(std::string(maps[w]) + "_key[" + std::to_string(k) + "]"), maps[w]));
        if (it == free3.end()) throw fail("an inertia depends on a free-joint frame that is not among the system's free-frame dofs");
        m |= uint64_t(1) << (32 + it->second);
      }
    }
    return m;
  }
};

void read_system(const Archive& A, const Node* sys, Builder& B, rkb_chain_desc& d) {
  if (class_of(sys) != T_NL_SYSTEM) throw fail("the archive's first object is not a kte_nl_system (type " + sys->type + ")");
  const long n = count_of(sys, "dofs_gen"), nf3 = count_of(sys, "dofs_3D"), nf2 = count_of(sys, "dofs_2D"), nu = count_of(sys, "inputs");
  if (nf3 != 0 && nf2 != 0) throw fail("2D and 3D free-frame dofs in one system");
  const long nf = nf3 + nf2;
  const std::string fdofs = nf2 ? "dofs_2D" : "dofs_3D", fcalc = nf2 ? "mFrames2D" : "mFrames3D";
  if (n > RKB_MAX_COORDS) throw fail("more generalized coordinates than RKB_MAX_COORDS");
  for (long i = 0; i < n; ++i) B.coords[B.oid(sys->kid("dofs_gen_q[" + std::to_string(i) + "]"), "dofs_gen")] = (int)i;
  for (long i = 0; i < nf; ++i) B.free3[B.oid(sys->kid(fdofs + "_q[" + std::to_string(i) + "]"), "free-frame dofs")] = (int)i;
  for (long i = 0; i < nu; ++i) B.inputs[B.oid(sys->kid("inputs_q[" + std::to_string(i) + "]"), "inputs")] = (int)i;
  B.n_coords = (int)n;
  const Node* chain = A.deref(sys->kid("chain"));
  const Node* mcalc = A.deref(sys->kid("mass_calc"));
  if (!chain || class_of(chain) != T_CHAIN || !mcalc || class_of(mcalc) != T_MASS_CALC) throw fail("kte_nl_system without chain or mass_calc");
  // mass_matrix_calc must list the system's coordinates and free frames in the system's order (kte_nl_system.hpp:271)
  if (count_of(mcalc, "mCoords") != n || count_of(mcalc, "mFrames3D") != nf3 || count_of(mcalc, "mFrames2D") != nf2)
    throw fail("mass_matrix_calc coordinates differ from the system dofs");
  for (long i = 0; i < n; ++i)
    if (B.cid(mcalc->kid("mCoords_q[" + std::to_string(i) + "]"), "mCoords") != (int)i) throw fail("mass_matrix_calc coordinates differ from the system dofs");
  for (long i = 0; i < nf; ++i) {
    std::map<long, int>::const_iterator it = B.free3.find(B.oid(mcalc->kid(fcalc + "_q[" + std::to_string(i) + "]"), "mass_calc free frames"));
    if (it == B.free3.end() || it->second != (int)i) throw fail("mass_matrix_calc free frames differ from the system dofs");
  }
  const long nk = count_of(chain, "mKTEs");
  std::vector<const Node*> ktes;
  for (long e = 0; e < nk; ++e) {
    const Node* f = chain->kid("mKTEs_q[" + std::to_string(e) + "]");
    if (!f) throw fail("chain element missing");
    ktes.push_back(f);
  }
  for (const Node* f : ktes) {  // 2D or 3D: decided by the first joint / link / inertia, as the bridge does
    const unsigned long t = class_of(f);
    if (t == T_REV_3D || t == T_PRI_3D || t == T_LINK_3D || t == T_IN_3D || t == T_FREE_3D) { B.dim = 3; break; }
    if (t == T_REV_2D || t == T_PRI_2D || t == T_LINK_2D || t == T_IN_2D || t == T_FREE_2D) { B.dim = 2; break; }
  }
  if (!B.dim) throw fail("chain has no 2D or 3D element");
  std::vector<std::pair<int, long> > pending;  // (actuator element, object id of its joint)
  for (const Node* f : ktes) {
    const Node* k = A.deref(f);
    const unsigned long t = class_of(f);
    int idx = -1;
    double p[12] = {0};
    switch (t) {
      case T_REV_3D: case T_PRI_3D: case T_REV_2D: case T_PRI_2D: {
        const int fa = B.fid(k->kid("mBase"), "mBase"), fb = B.fid(k->kid("mEnd"), "mEnd");
        B.written[fb] = 1;
        const bool rev = (t == T_REV_3D || t == T_REV_2D), is3 = (t == T_REV_3D || t == T_PRI_3D);
        const int c = B.cid(k->kid(rev ? "mAngle" : "mCoord"), "joint coordinate");
        if (t != T_REV_2D) vect(k->kid("mAxis"), is3 ? 3 : 2, p, "mAxis");
        idx = B.push(t == T_REV_3D ? RKB_REVOLUTE_3D : t == T_PRI_3D ? RKB_PRISMATIC_3D : t == T_REV_2D ? RKB_REVOLUTE_2D : RKB_PRISMATIC_2D,
                     fa, fb, c, 0, 0, p, 3);
        break;
      }
      case T_FREE_3D: case T_FREE_2D: {
        const int fa = B.fid(k->kid("mBase"), "mBase"), fb = B.fid(k->kid("mEnd"), "mEnd");
        B.written[fb] = 1;
        std::map<long, int>::const_iterator it = B.free3.find(B.oid(k->kid("mCoord"), "mCoord"));
        if (it == B.free3.end()) throw fail("a free joint's coordinate frame is not listed in the system's free-frame dofs");
        idx = B.push(t == T_FREE_3D ? RKB_FREE_3D : RKB_FREE_2D, fa, fb, it->second, 0, 0, p, 0);
        break;
      }
      case T_LINK_3D: {
        const int fa = B.fid(k->kid("mBase"), "mBase"), fb = B.fid(k->kid("mEnd"), "mEnd");
        B.written[fb] = 1;
        const Node* o = k->kid("mPoseOffset");
        if (!o) throw fail("rigid_link_3D without mPoseOffset");
        vect(o->kid("Position"), 3, p, "mPoseOffset.Position");
        quat(o->kid("Quat"), p + 3);
        idx = B.push(RKB_RIGID_LINK_3D, fa, fb, -1, 0, 0, p, 7);
        break;
      }
      case T_LINK_2D: {
        const int fa = B.fid(k->kid("mBase"), "mBase"), fb = B.fid(k->kid("mEnd"), "mEnd");
        B.written[fb] = 1;
        const Node* o = k->kid("mPoseOffset");
        if (!o || !o->kid("Rotation")) throw fail("rigid_link_2D without mPoseOffset");
        vect(o->kid("Position"), 2, p, "mPoseOffset.Position");
        p[2] = std::atan2(number(o->kid("Rotation")->kid("sin"), "sin"), number(o->kid("Rotation")->kid("cos"), "cos"));  // rot_mat_2D::getAngle
        idx = B.push(RKB_RIGID_LINK_2D, fa, fb, -1, 0, 0, p, 3);
        break;
      }
      case T_IN_3D: {
        const Node* dep = A.deref(k->kid("mCenterOfMass"));
        if (!dep) throw fail("inertia_3D without mCenterOfMass");
        const int fr = B.fid(dep->kid("mFrame"), "mFrame");
        p[0] = number(k->kid("mMass"), "mMass");
        const Node* I = k->kid("mInertiaTensor");  // mat<double, symmetric>: packed lower triangle by rows (a00 a10 a11 a20 a21 a22)
        if (!I || count_of(I, "q") != 6) throw fail("inertia_3D: mInertiaTensor is not a 3 x 3 symmetric matrix");
        double q[6];
        for (int i = 0; i < 6; ++i) q[i] = number(I->kid("q_q[" + std::to_string(i) + "]"), "mInertiaTensor");
        p[1] = q[0]; p[2] = q[1]; p[3] = q[3]; p[4] = q[2]; p[5] = q[4]; p[6] = q[5];
        idx = B.push(RKB_INERTIA_3D, fr, -1, -1, 0, B.upstream(dep), p, 7);
        break;
      }
      case T_IN_2D: {
        const Node* dep = A.deref(k->kid("mCenterOfMass"));
        if (!dep) throw fail("inertia_2D without mCenterOfMass");
        const int fr = B.fid(dep->kid("mFrame"), "mFrame");
        p[0] = number(k->kid("mMass"), "mMass");
        p[1] = number(k->kid("mMomentOfInertia"), "mMomentOfInertia");
        idx = B.push(RKB_INERTIA_2D, fr, -1, -1, 0, B.upstream(dep), p, 2);
        break;
      }
      case T_IN_GEN: {
        const Node* dep = A.deref(k->kid("mCenterOfMass"));
        if (!dep) throw fail("inertia_gen without mCenterOfMass");
        const int c = B.cid(dep->kid("mFrame"), "mFrame");
        if (B.upstream(dep) != (uint64_t(1) << c)) throw fail("inertia_gen must depend on its own coordinate only");
        p[0] = number(k->kid("mMass"), "mMass");
        idx = B.push(RKB_INERTIA_GEN, -1, -1, c, 0, uint64_t(1) << c, p, 1);
        break;
      }
      case T_ACTUATOR_GEN: {
        std::map<long, int>::const_iterator it = B.inputs.find(B.oid(f, "actuator"));
        if (it == B.inputs.end()) throw fail("a driving actuator of the chain is not listed in the system inputs");
        idx = B.push(RKB_ACTUATOR_GEN, -1, -1, B.cid(k->kid("mFrame"), "mFrame"), it->second, 0, p, 0);
        pending.push_back(std::make_pair(idx, B.oid(k->kid("mJoint"), "mJoint")));
        break;
      }
      case T_TSPRING_3D: case T_TSPRING_2D: {
        const int a = B.fid(k->kid("mAnchor1"), "mAnchor1"), b = B.fid(k->kid("mAnchor2"), "mAnchor2");
        p[0] = number(k->kid("mStiffness"), "mStiffness"); p[1] = number(k->kid("mSaturation"), "mSaturation");
        idx = B.push(B.dim == 3 ? RKB_TORSION_SPRING_3D : RKB_TORSION_SPRING_2D, a, b, -1, 0, 0, p, 2);
        break;
      }
      case T_TDAMPER_3D: case T_TDAMPER_2D: {
        const int a = B.fid(k->kid("mAnchor1"), "mAnchor1"), b = B.fid(k->kid("mAnchor2"), "mAnchor2");
        p[0] = number(k->kid("mDamping"), "mDamping");
        idx = B.push(B.dim == 3 ? RKB_TORSION_DAMPER_3D : RKB_TORSION_DAMPER_2D, a, b, -1, 0, 0, p, 1);
        break;
      }
      case T_SPRING_3D: case T_SPRING_2D: {
        const int a = B.fid(k->kid("mAnchor1"), "mAnchor1"), b = B.fid(k->kid("mAnchor2"), "mAnchor2");
        p[0] = number(k->kid("mRestLength"), "mRestLength"); p[1] = number(k->kid("mStiffness"), "mStiffness"); p[2] = number(k->kid("mSaturation"), "mSaturation");
        idx = B.push(B.dim == 3 ? RKB_SPRING_3D : RKB_SPRING_2D, a, b, -1, 0, 0, p, 3);
        break;
      }
      case T_DAMPER_3D: case T_DAMPER_2D: {
        const int a = B.fid(k->kid("mAnchor1"), "mAnchor1"), b = B.fid(k->kid("mAnchor2"), "mAnchor2");
        p[0] = number(k->kid("mDamping"), "mDamping");
        idx = B.push(B.dim == 3 ? RKB_DAMPER_3D : RKB_DAMPER_2D, a, b, -1, 0, 0, p, 1);
        break;
      }
      case T_LINK_GEN: {
        const int a = B.gid(k->kid("mBase"), "mBase"), b = B.gid(k->kid("mEnd"), "mEnd");
        if (b < B.n_coords) throw fail("rigid_link_gen ends on a system dof");
        p[0] = number(k->kid("mOffset"), "mOffset");
        idx = B.push(RKB_RIGID_LINK_GEN, -1, -1, a, b, 0, p, 1);
        break;
      }
      case T_SPRING_GEN: {
        const int a = B.gid(k->kid("mAnchor1"), "mAnchor1"), b = B.gid(k->kid("mAnchor2"), "mAnchor2");
        p[0] = number(k->kid("mRestLength"), "mRestLength"); p[1] = number(k->kid("mStiffness"), "mStiffness"); p[2] = number(k->kid("mSaturation"), "mSaturation");
        idx = B.push(RKB_SPRING_GEN, -1, -1, a, b, 0, p, 3);
        break;
      }
      case T_DAMPER_GEN: {
        const int a = B.gid(k->kid("mAnchor1"), "mAnchor1"), b = B.gid(k->kid("mAnchor2"), "mAnchor2");
        p[0] = number(k->kid("mDamping"), "mDamping");
        idx = B.push(RKB_DAMPER_GEN, -1, -1, a, b, 0, p, 1);
        break;
      }
      default: {
        const Node* nm = k->kid("name");
        throw fail("KTE '" + (nm ? nm->text : std::string("?")) + "' (class " + f->type + ") is outside the compiled element set");
      }
    }
    if (f->object_id > 0) B.elem_of[f->object_id] = idx;
  }
  for (size_t i = 0; i < pending.size(); ++i) {
    std::map<long, int>::const_iterator it = B.elem_of.find(pending[i].second);
    if (it == B.elem_of.end()) throw fail("an actuator drives a joint that is not in the chain");
    B.el[pending[i].first].frame_b = it->second;
  }
  int root = -1, n_roots = 0;
  for (std::map<long, int>::const_iterator it = B.frames.begin(); it != B.frames.end(); ++it)
    if (!B.written.count(it->second)) { root = it->second; ++n_roots; }
  if (n_roots != 1) throw fail("the chain must have exactly one un-driven base frame");
  std::memset(&d, 0, sizeof d);
  d.dim = B.dim; d.n_elements = (int)B.el.size(); d.n_frames = (int)B.frames.size(); d.n_coords = (int)n; d.n_inputs = (int)nu; d.base_frame = root;
  const Node* bf = B.frame_node[root];
  if (B.dim == 3) {
    vect(bf->kid("Position"), 3, d.base.position, "Position");
    quat(bf->kid("Quat"), d.base.quat);
    vect(bf->kid("Velocity"), 3, d.base.velocity, "Velocity");
    vect(bf->kid("AngVelocity"), 3, d.base.ang_velocity, "AngVelocity");
    vect(bf->kid("Acceleration"), 3, d.base.acceleration, "Acceleration");
    vect(bf->kid("AngAcceleration"), 3, d.base.ang_acceleration, "AngAcceleration");
  } else {
    vect(bf->kid("Position"), 2, d.base.position, "Position");
    const Node* R = bf->kid("Rotation");
    if (!R) throw fail("field missing: Rotation");
    d.base.quat[0] = std::atan2(number(R->kid("sin"), "sin"), number(R->kid("cos"), "cos"));
    vect(bf->kid("Velocity"), 2, d.base.velocity, "Velocity");
    d.base.ang_velocity[0] = number(bf->kid("AngVelocity"), "AngVelocity");
    vect(bf->kid("Acceleration"), 2, d.base.acceleration, "Acceleration");
    d.base.ang_acceleration[0] = number(bf->kid("AngAcceleration"), "AngAcceleration");
  }
}

int read_file(const char* path, rkb_chain_desc& d, std::vector<rkb_element>& el, std::string& err) {
  std::FILE* fp = std::fopen(path, "rb");
  if (!fp) { err = std::string("cannot open ") + path; return RKB_ERR_INVALID; }
  std::string src;
  char buf[1 << 16];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, fp)) > 0) src.append(buf, got);
  std::fclose(fp);
  try {
    // prolog: <?xml ... ?>, <!DOCTYPE ...>, then <reak_serialization version="..."> holding the saved objects
    size_t at = src.find("<reak_serialization");
    if (at == std::string::npos) throw fail("not a ReaK XML archive (no <reak_serialization>)");
    Parser P(src);
    P.i = at;
    Archive A;
    A.root = P.record();
    if (!A.root || A.root->kids.empty()) throw fail("empty archive");
    A.index(A.root);
    Builder B(A);
    read_system(A, A.deref(A.root->kids[0]), B, d);
    el.swap(B.el);
    d.elements = el.data();
    return RKB_OK;
  } catch (std::exception& e) {
    err = e.what();
    return RKB_ERR_UNSUPPORTED;
  }
}

void put_err(const std::string& e, char* err, size_t err_len) {
  if (err && err_len > 0) { std::strncpy(err, e.c_str(), err_len - 1); err[err_len - 1] = 0; }
}

}  // namespace

extern "C" {

int rkb_rkx_read(const char* path, rkb_chain_desc* desc, rkb_element* elements, int max_elements, char* err, size_t err_len) {
  if (!path || !desc) return RKB_ERR_INVALID;
  rkb_chain_desc d;
  std::vector<rkb_element> el;
  std::string e;
  const int rc = read_file(path, d, el, e);
  if (rc != RKB_OK) { put_err(e, err, err_len); return rc; }
  if (elements) {
    if ((int)el.size() > max_elements) { put_err("more elements than the caller's buffer holds", err, err_len); return RKB_ERR_NOMEM; }
    for (size_t i = 0; i < el.size(); ++i) elements[i] = el[i];
  }
  *desc = d;
  desc->elements = elements;
  return (int)el.size();
}

int rkb_rkx_load(const char* path, unsigned create_flags, rkb_chain** out, char* err, size_t err_len) {
  if (!path || !out) return RKB_ERR_INVALID;
  *out = nullptr;
  rkb_chain_desc d;
  std::vector<rkb_element> el;
  std::string e;
  int rc = read_file(path, d, el, e);
  if (rc != RKB_OK) { put_err(e, err, err_len); return rc; }
  rc = rkb_chain_create_ex(&d, create_flags, out);
  if (rc != RKB_OK) put_err(std::string("the chain of the archive was rejected: ") + rkb_strerror(rc), err, err_len);
  return rc;
}

}  // extern "C"
