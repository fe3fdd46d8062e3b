#!/usr/bin/env python
"""flop_count.py — ALGORITHMIC FP64 operation count of the formulation the serial-chain kernels ship
(reak_b200/csrc/kte_serial.cuh; derivation in DESIGN.md section 5).

The count is obtained by running the formulation itself — sweep 1 (kinematics outward, d'Alembert wrenches),
sweep 2 (wrenches inward, generalised forces), sweep 3 (composite-inertia mass matrix), the packed L D L^T solve,
the sine / cosine evaluation and the RK4 combination — on SYMBOLIC scalars that only know whether a value is a
structural zero, a literal +-1, or something else, and counting every multiplication and addition whose operands
are not trivial.  What is structural: the promises of the kernel's SHAPE template argument (a revolute axis that is
+-e_D, a link offset along one axis without rotation, a diagonal tensor, a literal axis sign), the zero
initialisation of the composite inertia and the zero component of `a x h`.  What is NOT: the values of chain
constants (unit masses, zero offsets, a base frame at rest are run-time numbers in constant memory) — the count
belongs to the path, not to the benchmark's constants.

  flops        = multiplications + additions        (a fused multiply-add counts 2: the roofline unit)
  instructions = multiplications + additions - fusable pairs (an addition one of whose operands is a product
                 formed for it) — the lower bound on DFMA/DMUL/DADD issue slots

This is a developer / bench tool (bench.py imports it for roofline.algorithmic_flop_per_state_step); it is pure
Python and needs neither the GPU nor the oracle.

  python tools/flop_count.py [preset ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class Counter(object):
    """Collects the expression graph; only operations an output depends on are counted (the code is fully unrolled,
    so whatever the outputs do not need is dead code for the compiler too: e.g. the composite inertia is carried
    into the first joint's base frame only as far as M(0,0) needs it)."""

    def __init__(self):
        self.phase = ""
        self.extra = {}  # closed-form pieces: phase -> [mul, add, fused]

    def bump(self, mul, add, fused):
        e = self.extra.setdefault(self.phase, [0, 0, 0])
        e[0] += mul; e[1] += add; e[2] += fused

    def count(self, outputs):
        """{phase: {'mul', 'add', 'fused', 'flops', 'instr'}} over the operations reachable from `outputs`"""
        seen, order, stack = set(), [], [o for o in outputs if isinstance(o, V)]
        while stack:
            v = stack.pop()
            if id(v) in seen:
                continue
            seen.add(id(v))
            order.append(v)
            stack.extend(v.args)
        uses = {}
        for v in order:
            for a in v.args:
                uses[id(a)] = uses.get(id(a), 0) + 1
        res = {}
        for v in order:
            if v.op is None:
                continue
            r = res.setdefault(v.phase, [0, 0, 0])
            if v.op == "mul":
                r[0] += 1
            else:
                r[1] += 1
                # an addition absorbs ONE product formed only for it: a fused multiply-add
                if any(a.op == "mul" and uses.get(id(a), 0) == 1 for a in v.args):
                    r[2] += 1
        for ph, e in self.extra.items():
            r = res.setdefault(ph, [0, 0, 0])
            for i in range(3):
                r[i] += e[i]
        return {ph: {"mul": m, "add": a, "fused": f, "flops": m + a, "instr": m + a - f} for ph, (m, a, f) in res.items()}

    def depth(self, outputs):
        """length of the longest chain of dependent FP64 instructions behind `outputs` (a product formed only for one
        addition rides in that addition's fused multiply-add): the floor one sample's evaluation cannot go below,
        whatever the number of lanes or warps that share it, times the ~8-cycle DFMA latency"""
        uses, stack, seen = {}, [o for o in outputs if isinstance(o, V)], set()
        while stack:
            v = stack.pop()
            if id(v) in seen:
                continue
            seen.add(id(v))
            for a in v.args:
                uses[id(a)] = uses.get(id(a), 0) + 1
                stack.append(a)
        memo = {}

        def d(v):
            # iterative post-order (the chains are thousands of nodes long)
            st = [v]
            while st:
                n = st[-1]
                if id(n) in memo:
                    st.pop(); continue
                pend = [a for a in n.args if id(a) not in memo]
                if pend:
                    st.extend(pend); continue
                if n.op is None:
                    memo[id(n)] = 0
                elif n.op == "mul":
                    memo[id(n)] = 1 + max(memo[id(a)] for a in n.args)
                else:
                    best = 0
                    for a in n.args:
                        if a.op == "mul" and uses.get(id(a), 0) == 1:
                            best = max(best, max(memo[id(b)] for b in a.args))  # fused: the product costs no extra level
                        else:
                            best = max(best, memo[id(a)])
                    memo[id(n)] = 1 + best
                st.pop()
            return memo[id(v)]

        return max([d(o) for o in outputs if isinstance(o, V)] or [0])


class V(object):
    """kind: 'z' structural zero, 'p' literal +1, 'm' literal -1, 'v' any other value; op / args: how it was formed"""
    __slots__ = ("kind", "c", "op", "args", "phase")

    def __init__(self, c, kind="v", op=None, args=()):
        self.c, self.kind, self.op, self.args, self.phase = c, kind, op, args, c.phase

    def _lift(self, o):
        if isinstance(o, V):
            return o
        if o == 0:
            return V(self.c, "z")
        if o == 1:
            return V(self.c, "p")
        if o == -1:
            return V(self.c, "m")
        return V(self.c, "v")

    def __mul__(self, o):
        o = self._lift(o)
        if self.kind == "z" or o.kind == "z":
            return V(self.c, "z")
        if self.kind in "pm" and o.kind in "pm":
            return V(self.c, "p" if self.kind == o.kind else "m")
        if self.kind in "pm":
            return o      # a sign change is an operand modifier, not an instruction
        if o.kind in "pm":
            return self
        return V(self.c, "v", "mul", (self, o))

    __rmul__ = __mul__

    def __add__(self, o):
        o = self._lift(o)
        if self.kind == "z":
            return o
        if o.kind == "z":
            return self
        return V(self.c, "v", "add", (self, o))

    __radd__ = __add__

    def __neg__(self):
        if self.kind in "zpm":
            return V(self.c, {"z": "z", "p": "m", "m": "p"}[self.kind])
        return self

    def __sub__(self, o):
        return self + (-self._lift(o))

    def __rsub__(self, o):
        return self._lift(o) + (-self)


def vec(c, kinds="vvv"):
    return [V(c, k) for k in kinds]


def vadd(a, b):
    return [a[i] + b[i] for i in range(3)]


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def scale(s, a):
    return [s * a[i] for i in range(3)]


def rot_axis(D, c, s, v):  # R_D v
    D1, D2 = (D + 1) % 3, (D + 2) % 3
    r = [None] * 3
    r[D] = v[D]
    r[D1] = c * v[D1] - s * v[D2]
    r[D2] = s * v[D1] + c * v[D2]
    return r


def rotT_axis(D, c, s, v):  # R_D^T v
    D1, D2 = (D + 1) % 3, (D + 2) % 3
    r = [None] * 3
    r[D] = v[D]
    r[D1] = c * v[D1] + s * v[D2]
    r[D2] = c * v[D2] - s * v[D1]
    return r


def matvec(R, v, transpose=False):
    if transpose:
        return [R[0][i] * v[0] + R[1][i] * v[1] + R[2][i] * v[2] for i in range(3)]
    return [R[i][0] * v[0] + R[i][1] * v[1] + R[i][2] * v[2] for i in range(3)]


def rodrigues(c, cs, sn):
    """axis_angle(q, axis).getRotMat() from cos / sin and the constants an, an an^T (kte_serial.cuh: rodrigues)"""
    K = lambda: V(c)
    omc = 1 - cs
    t12, t13, t23 = omc * K(), omc * K(), omc * K()
    t01, t02, t03 = sn * K(), sn * K(), sn * K()
    return [[cs + omc * K(), t12 - t03, t13 + t02], [t12 + t03, cs + omc * K(), t23 - t01], [t13 - t02, t23 + t01, cs + omc * K()]]


SYM = lambda i, j: (min(i, j), max(i, j))


def sym_get(I, i, j):
    return I[SYM(i, j)]


def sym_rotate_axis(D, c, s, I):
    """I <- R I R^T for a rotation about e_D (kte_serial.cuh: sym_rotate_axis_t)"""
    p, q = (D + 1) % 3, (D + 2) % 3
    ipr, iqr = I[SYM(p, D)], I[SYM(q, D)]
    I[SYM(p, D)] = c * ipr - s * iqr
    I[SYM(q, D)] = s * ipr + c * iqr
    ipp, ipq, iqq = I[SYM(p, p)], I[SYM(p, q)], I[SYM(q, q)]
    app, apq = c * ipp - s * ipq, c * ipq - s * iqq
    aqp, aqq = s * ipp + c * ipq, s * ipq + c * iqq
    I[SYM(p, p)] = c * app - s * apq
    I[SYM(p, q)] = s * app + c * apq
    I[SYM(q, q)] = s * aqp + c * aqq


def sym_rotate(R, I):
    A = [[R[i][0] * sym_get(I, 0, j) + R[i][1] * sym_get(I, 1, j) + R[i][2] * sym_get(I, 2, j) for j in range(3)] for i in range(3)]
    for i in range(3):
        for j in range(i, 3):
            I[(i, j)] = A[i][0] * R[j][0] + A[i][1] * R[j][1] + A[i][2] * R[j][2]


def sym_shift(c, p, mcp, h, I):
    """parallel-axis move by the constant offset p (kte_serial.cuh: sym_shift_c)"""
    w = [h[i] + 0.5 * mcp[i] for i in range(3)]  # (0.5 * constant is a constant: the product below is counted once)
    d = 2 * dot(p, w)
    for i in range(3):
        I[(i, i)] = I[(i, i)] + d - 2 * (w[i] * p[i])
    for (i, j) in ((0, 1), (0, 2), (1, 2)):
        I[(i, j)] = I[(i, j)] - (w[i] * p[j] + w[j] * p[i])
    for i in range(3):
        h[i] = h[i] + mcp[i]


class Stage(object):
    """joint: 'R' revolute / 'P' prismatic; D: 0..2 when the axis is +-e_D else None; signed: the axis direction is a
    literal; link: None, 'axis' (offset along e_LD, no rotation), 'general', 'rotated'; inertia: None, 'diag', 'full'."""

    def __init__(self, joint="R", D=None, signed=False, link=None, LD=2, inertia=None, spring=False, damper=False):
        self.joint, self.D, self.signed, self.link, self.LD, self.inertia, self.spring, self.damper = joint, D, signed, link, LD, inertia, spring, damper


def sweeps(c, stages, want_f=True, want_m=True):
    """one evaluation: f and the packed M, as kte_serial.cuh: serial_sweeps + mass_sweep"""
    N = len(stages)
    K = lambda: V(c)           # a run-time value (state, input or chain constant)
    cs, sn = [K() for _ in stages], [K() for _ in stages]
    f_out = [None] * N
    if want_f:
        w, al, a = vec(c), vec(c), vec(c)
        park = []
        for k, S in enumerate(stages):
            sg = (V(c, "p") if S.signed else K())
            if S.joint == "R" and S.D is not None:
                D, D1, D2 = S.D, (S.D + 1) % 3, (S.D + 2) % 3
                g = sg * K()
                wt, alt, at = rotT_axis(D, cs[k], sn[k], w), rotT_axis(D, cs[k], sn[k], al), rotT_axis(D, cs[k], sn[k], a)
                alt[D1] = alt[D1] + wt[D2] * g
                alt[D2] = alt[D2] - wt[D1] * g
                wt[D] = wt[D] + g
                w, al, a = wt, alt, at
            elif S.joint == "P" and S.D is not None:
                D, D1, D2 = S.D, (S.D + 1) % 3, (S.D + 2) % 3
                L, Ld2 = sg * K(), 2 * (sg * K())
                a[D] = a[D] - L * (w[D1] * w[D1] + w[D2] * w[D2])
                a[D1] = a[D1] + L * (w[D] * w[D1] + al[D2]) + Ld2 * w[D2]
                a[D2] = a[D2] + L * (w[D] * w[D2] - al[D1]) - Ld2 * w[D1]
            elif S.joint == "R":
                R = rodrigues(c, cs[k], sn[k])
                wt = matvec(R, w, True)
                qda = scale(K(), vec(c))
                al = vadd(matvec(R, al, True), cross(wt, qda))
                w = vadd(wt, qda)
                a = matvec(R, a, True)
            else:
                r, rd = scale(K(), vec(c)), scale(K(), vec(c))
                a = vadd(vadd(vadd(a, cross(w, cross(w, r))), scale(2, cross(w, rd))), cross(al, r))
            if S.link == "axis":
                D, D1, D2 = S.LD, (S.LD + 1) % 3, (S.LD + 2) % 3
                L = K()
                a[D] = a[D] - L * (w[D1] * w[D1] + w[D2] * w[D2])
                a[D1] = a[D1] + L * (w[D] * w[D1] + al[D2])
                a[D2] = a[D2] + L * (w[D] * w[D2] - al[D1])
            elif S.link in ("general", "rotated"):
                po = vec(c)
                a = vadd(vadd(a, cross(w, cross(w, po))), cross(al, po))
                if S.link == "rotated":
                    Ro = [vec(c) for _ in range(3)]
                    a, w, al = matvec(Ro, a, True), matvec(Ro, w, True), matvec(Ro, al, True)
            if S.inertia == "diag":
                Fk = scale(-K(), a)
                Iw = [K() * w[i] for i in range(3)]
                Ial = [K() * al[i] for i in range(3)]
                Tk = [-(x + y) for x, y in zip(Ial, cross(w, Iw))]
            elif S.inertia == "full":
                Fk = scale(-K(), a)
                T6 = {(i, j): K() for i in range(3) for j in range(i, 3)}
                sm = lambda v: [sym_get(T6, i, 0) * v[0] + sym_get(T6, i, 1) * v[1] + sym_get(T6, i, 2) * v[2] for i in range(3)]
                Iw = sm(w)
                Tk = [-(x + y) for x, y in zip(sm(al), cross(w, Iw))]
            else:
                Fk, Tk = vec(c, "zzz"), vec(c, "zzz")
            park.append((Fk, Tk))
        F, T = vec(c, "zzz"), vec(c, "zzz")
        for k in range(N - 1, -1, -1):
            S = stages[k]
            F, T = vadd(F, park[k][0]), vadd(T, park[k][1])
            if S.link == "axis":
                D, D1, D2 = S.LD, (S.LD + 1) % 3, (S.LD + 2) % 3
                L = K()
                T[D1] = T[D1] - L * F[D2]
                T[D2] = T[D2] + L * F[D1]
            elif S.link in ("general", "rotated"):
                if S.link == "rotated":
                    Ro = [vec(c) for _ in range(3)]
                    F, T = matvec(Ro, F), matvec(Ro, T)
                T = vadd(T, cross(vec(c), F))
            sg = (V(c, "p") if S.signed else K())
            u = K()
            last = k == 0  # nothing reads the wrench below the first joint: its rotation into the base is dead code
            if S.joint == "R" and S.D is not None:
                D = S.D
                tsd = V(c, "z")
                if S.spring:
                    # wrapped angle: q - 2 pi rint(q / 2 pi) (1 mul, 1 fma), stiffness * |r| (1 mul); sign and saturation are selects
                    r = K() + K() * (K() * K())
                    tsd = sg * (K() * r)
                if S.damper:
                    tsd = tsd + sg * (K() * K())
                f_out[k] = sg * (T[D] - tsd) + u
                T[D] = tsd - sg * u
                if not last:
                    F, T = rot_axis(D, cs[k], sn[k], F), rot_axis(D, cs[k], sn[k], T)
            elif S.joint == "P" and S.D is not None:
                D, D1, D2 = S.D, (S.D + 1) % 3, (S.D + 2) % 3
                L = sg * K()
                f_out[k] = sg * F[D] + u
                if not last:
                    T[D1] = T[D1] - L * F[D2]
                    T[D2] = T[D2] + L * F[D1]
                    F[D] = -(sg * u)
            elif S.joint == "R":
                ax = vec(c)
                tsd = vec(c, "zzz")
                if S.spring:
                    r = K() + K() * (K() * K())
                    tsd = scale(K() * r, vec(c))
                if S.damper:
                    tsd = vadd(tsd, scale(K() * K(), ax))
                if S.spring or S.damper:
                    T = [T[i] - tsd[i] for i in range(3)]
                ta = dot(T, ax)
                f_out[k] = ta + u
                if not last:
                    R = rodrigues(c, cs[k], sn[k])
                    F = matvec(R, F)
                    T = [x - y for x, y in zip(matvec(R, [T[i] - ta * ax[i] for i in range(3)]), scale(u, ax))]
                    if S.spring or S.damper:
                        T = vadd(T, tsd)
            else:
                ax = vec(c)
                fa = dot(F, ax)
                f_out[k] = fa + u
                if not last:
                    T = vadd(T, cross(scale(K(), ax), F))
                    F = [F[i] - fa * ax[i] - u * ax[i] for i in range(3)]
    M = {}
    if want_m:
        h = vec(c, "zzz")
        I = {(i, j): V(c, "z") for i in range(3) for j in range(i, 3)}
        for k in range(N - 1, -1, -1):
            S = stages[k]
            if S.inertia == "diag":
                for i in range(3):
                    I[(i, i)] = I[(i, i)] + K()
            elif S.inertia == "full":
                for key in I:
                    I[key] = I[key] + K()
            if S.link == "axis":
                D, D1, D2 = S.LD, (S.LD + 1) % 3, (S.LD + 2) % 3
                L, mcpo = K(), K()
                dd = L * (2 * h[D] + mcpo)
                I[(D1, D1)] = I[(D1, D1)] + dd
                I[(D2, D2)] = I[(D2, D2)] + dd
                I[SYM(D1, D)] = I[SYM(D1, D)] - L * h[D1]
                I[SYM(D2, D)] = I[SYM(D2, D)] - L * h[D2]
                h[D] = h[D] + mcpo
            elif S.link in ("general", "rotated"):
                if S.link == "rotated":
                    Ro = [vec(c) for _ in range(3)]
                    h = matvec(Ro, h)
                    sym_rotate(Ro, I)
                sym_shift(c, vec(c), vec(c), h, I)
            sg = (V(c, "p") if S.signed else K())
            if S.joint == "R" and S.D is not None:
                D, D1, D2 = S.D, (S.D + 1) % 3, (S.D + 2) % 3
                n = [sg * sym_get(I, i, D) for i in range(3)]
                f = [None] * 3
                f[D], f[D1], f[D2] = V(c, "z"), (-sg) * h[D2], sg * h[D1]
                M[(k, k)] = I[(D, D)] + K()
            elif S.joint == "P" and S.D is not None:
                D, D1, D2 = S.D, (S.D + 1) % 3, (S.D + 2) % 3
                f = [V(c, "z")] * 3
                f[D] = sg * K()
                n = [None] * 3
                n[D], n[D1], n[D2] = V(c, "z"), sg * h[D2], (-sg) * h[D1]
                M[(k, k)] = K()  # mc + rotor: constants
            elif S.joint == "R":
                ax = vec(c)
                n = [sym_get(I, i, 0) * ax[0] + sym_get(I, i, 1) * ax[1] + sym_get(I, i, 2) * ax[2] for i in range(3)]
                f = cross(ax, h)
                M[(k, k)] = dot(ax, n) + K()
            else:
                ax = vec(c)
                f = scale(K(), ax)
                n = cross(h, ax)
                M[(k, k)] = dot(ax, f) + K()
            for j in range(k, 0, -1):
                Sj, Si = stages[j], stages[j - 1]
                sgj = (V(c, "p") if Sj.signed else K())
                if Sj.joint == "R" and Sj.D is not None:
                    f, n = rot_axis(Sj.D, cs[j], sn[j], f), rot_axis(Sj.D, cs[j], sn[j], n)
                elif Sj.joint == "P" and Sj.D is not None:
                    D, D1, D2 = Sj.D, (Sj.D + 1) % 3, (Sj.D + 2) % 3
                    L = sgj * K()
                    n[D1] = n[D1] - L * f[D2]
                    n[D2] = n[D2] + L * f[D1]
                elif Sj.joint == "R":
                    R = rodrigues(c, cs[j], sn[j])
                    f, n = matvec(R, f), matvec(R, n)
                else:
                    n = vadd(n, cross(scale(K(), vec(c)), f))
                if Si.link == "axis":
                    D, D1, D2 = Si.LD, (Si.LD + 1) % 3, (Si.LD + 2) % 3
                    L = K()
                    n[D1] = n[D1] - L * f[D2]
                    n[D2] = n[D2] + L * f[D1]
                elif Si.link in ("general", "rotated"):
                    if Si.link == "rotated":
                        Ro = [vec(c) for _ in range(3)]
                        f, n = matvec(Ro, f), matvec(Ro, n)
                    n = vadd(n, cross(vec(c), f))
                sgi = (V(c, "p") if Si.signed else K())
                if Si.D is not None:
                    M[(j - 1, k)] = sgi * (n[Si.D] if Si.joint == "R" else f[Si.D])
                else:
                    M[(j - 1, k)] = dot(vec(c), n if Si.joint == "R" else f)
            if k > 0:
                if S.joint == "R" and S.D is not None:
                    h = rot_axis(S.D, cs[k], sn[k], h)
                    sym_rotate_axis(S.D, cs[k], sn[k], I)
                elif S.joint == "P" and S.D is not None:
                    D, D1, D2 = S.D, (S.D + 1) % 3, (S.D + 2) % 3
                    L = sg * K()
                    mcL = K() * L
                    dd = L * (2 * h[D] + mcL)
                    I[(D1, D1)] = I[(D1, D1)] + dd
                    I[(D2, D2)] = I[(D2, D2)] + dd
                    I[SYM(D1, D)] = I[SYM(D1, D)] - L * h[D1]
                    I[SYM(D2, D)] = I[SYM(D2, D)] - L * h[D2]
                    h[D] = h[D] + mcL
                elif S.joint == "R":
                    R = rodrigues(c, cs[k], sn[k])
                    h = matvec(R, h)
                    sym_rotate(R, I)
                else:
                    p = scale(K(), vec(c))
                    mcp = scale(K(), p)
                    w = [h[i] + 0.5 * mcp[i] for i in range(3)]
                    d = 2 * dot(w, p)
                    for i in range(3):
                        I[(i, i)] = I[(i, i)] + d - 2 * (w[i] * p[i])
                    for (i, j) in ((0, 1), (0, 2), (1, 2)):
                        I[(i, j)] = I[(i, j)] - (w[i] * p[j] + p[i] * w[j])
                    h = vadd(h, mcp)
    return f_out, M


def ldl_solve(c, N, M, b):
    """packed L D L^T with one reciprocal per pivot (MUFU seed + two Newton steps = 4 fused multiply-adds), forward and
    back substitution (kte_serial.cuh: cholesky_solve_packed)"""
    L, inv = {}, [None] * N
    for i in range(N):
        W = [None] * i
        for j in range(i):
            s = M[(j, i)]
            for k in range(j):
                s = s - W[k] * L[(j, k)]
            W[j] = s
            L[(i, j)] = s * inv[j]
        d = M[(i, i)]
        for k in range(i):
            d = d - W[k] * L[(i, k)]
        # r = seed; e = 1 - d r; r = r + r e; e = 1 - d r; r = r + r e
        r = V(c)
        for _ in range(2):
            e = 1 - d * r
            r = r + r * e
        inv[i] = r
    y = list(b)
    for i in range(N):
        s = y[i]
        for k in range(i):
            s = s - L[(i, k)] * y[k]
        y[i] = s
    for i in range(N - 1, -1, -1):
        s = y[i] * inv[i]
        for k in range(N - 1, i, -1):
            s = s - L[(k, i)] * y[k]
        y[i] = s
    return y


def trig_full(c):
    """sincos_reduced: 1 fma + 1 add (rounding), 3 fma (Cody-Waite), z, two degree-5 Horner chains, 2 + 3 to finish"""
    c.bump(20, 18, 17)


def trig_shift(c):
    """d = q - w (1 add), then sincos_shift: 5 mul, 7 fma, 2 add"""
    c.bump(12, 10, 7)


def rk4_combine(c, n_states):
    """per state component and step: k1..k4 = dt f (4 mul), the three stage points (3 fma / add) and the final
    combination x += (k1 + 2 k2 + k4) / 6 - 2/3 k3 with k3 recovered as x - w (kte_serial.cuh: rk4_steps)"""
    c.bump(9 * n_states, 8 * n_states, 5 * n_states)


def count_chain(stages):
    """{'flop_per_evaluation', 'flop_per_state_step', 'instr_per_state_step', breakdown} for one RK4 state-step:
    4 evaluations (sweeps + solve), sine / cosine in full once and by the small-angle shift three times per revolute
    joint, and the RK4 combination."""
    N = len(stages)
    n_rev = sum(1 for s in stages if s.joint == "R")
    c = Counter()
    c.phase = "forces_sweeps_1_2"; f, _ = sweeps(c, stages, True, False)
    c.phase = "mass_sweep_3"; _, M = sweeps(c, stages, False, True)
    c.phase = "ldl_solve"; qdd = ldl_solve(c, N, M, f)
    c.phase = "sincos_full"; [trig_full(c) for _ in range(n_rev)]
    c.phase = "sincos_shift"; [trig_shift(c) for _ in range(n_rev)]
    c.phase = "rk4_combine_per_step"; rk4_combine(c, 2 * N)
    out = c.count(qdd)
    # dependent-instruction depth of one evaluation: sine / cosine first (full: 13 levels, shift: 8), then the sweeps and the solve
    depth_f, depth_m, depth_all = c.depth(f), c.depth(list(M.values())), c.depth(qdd)
    ev = {k: sum(out[p][k] for p in ("forces_sweeps_1_2", "mass_sweep_3", "ldl_solve")) for k in ("flops", "instr")}
    step = {k: 4 * ev[k] + out["sincos_full"][k] + 3 * out["sincos_shift"][k] + out["rk4_combine_per_step"][k] for k in ("flops", "instr")}
    return {"n": N, "flop_per_evaluation": ev["flops"] + out["sincos_full"]["flops"], "instr_per_evaluation": ev["instr"] + out["sincos_full"]["instr"],
            "flop_per_state_step": step["flops"], "instr_per_state_step": step["instr"], "breakdown": out,
            "critical_path": {"forces": depth_f, "mass": depth_m, "evaluation": depth_all, "sincos_full": 13, "sincos_shift": 8,
                              "per_state_step": 4 * depth_all + 13 + 3 * 8 + 4 * 2}}


def stages_of(prop):
    """Stage list of a kte_batch_propagator on the serial kernels: the structure its KERNEL is specialised on
    (rkb_chain_kernel_shape) plus what the descriptor says about links, inertias, springs and dampers."""
    from reak_b200 import _abi
    if not prop.is_serial():
        raise ValueError("the chain runs on the interpreter kernels")
    shape = prop.kernel_shape()
    d = prop.compiled.desc
    dim2 = d.dim == 2
    stages, cur = [], None
    for e in range(d.n_elements):
        E = d.elements[e]
        if E.kind in (_abi.REVOLUTE_3D, _abi.PRISMATIC_3D, _abi.REVOLUTE_2D, _abi.PRISMATIC_2D):
            cur = Stage(joint="R" if E.kind in (_abi.REVOLUTE_3D, _abi.REVOLUTE_2D) else "P")
            stages.append(cur)
        elif E.kind in (_abi.RIGID_LINK_3D, _abi.RIGID_LINK_2D) and cur is not None:
            rotated = (E.p[2] != 0.0) if dim2 else not (E.p[3] == 1.0 and E.p[4] == 0.0 and E.p[5] == 0.0 and E.p[6] == 0.0)
            cur.link = "rotated" if rotated else "general"
        elif E.kind in (_abi.INERTIA_3D, _abi.INERTIA_2D) and cur is not None:
            cur.inertia = "full"
        elif E.kind in (_abi.TORSION_SPRING_3D, _abi.TORSION_SPRING_2D) and cur is not None and cur.joint == "R":
            cur.spring = True
        elif E.kind in (_abi.TORSION_DAMPER_3D, _abi.TORSION_DAMPER_2D) and cur is not None and cur.joint == "R":
            cur.damper = True
    for k, S in enumerate(stages):
        b = (shape >> (8 * k)) & 0xff
        ax, lk, inn, sign = b & 7, (b >> 3) & 3, (b >> 5) & 1, (b >> 6) & 3
        if ax:
            S.D = (ax - 1) if ax <= 3 else (ax - 5)
            S.signed = sign != 0
        if lk:
            S.link, S.LD = "axis", lk - 1
        if inn:
            S.inertia = "diag"
    return stages


def main(argv):
    import json
    from reak_b200 import kte_batch_propagator, presets
    names = argv or ["crs6", "crs6_sd", "crs7", "planar2", "crs6_phys"]
    res = {}
    for name in names:
        p = kte_batch_propagator(presets.make(name))
        r = count_chain(stages_of(p))
        r["kernel_shape"] = hex(p.kernel_shape())
        res[name] = r
        print("%-10s n=%d  %5d flop / evaluation   %6d flop, >= %5d FP64 instructions / RK4 state-step" %
              (name, r["n"], r["flop_per_evaluation"], r["flop_per_state_step"], r["instr_per_state_step"]))
        cp = r["critical_path"]
        print("    critical path: forces %d, mass matrix %d, evaluation with solve %d dependent FP64 instructions; %d per RK4 state-step"
              % (cp["forces"], cp["mass"], cp["evaluation"], cp["per_state_step"]))
        for k, v in r["breakdown"].items():
            print("    %-22s %5d mul %5d add  -> %5d flop, %5d instr" % (k, v["mul"], v["add"], v["flops"], v["instr"]))
    return res


if __name__ == "__main__":
    main(sys.argv[1:])
