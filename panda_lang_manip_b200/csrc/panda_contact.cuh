// panda_contact.cuh -- free rigid bodies, contact generation and the full constrained sub-step, one thread per env.
//
// Replaces what the reference obtains from pybullet's stepSimulation for the task scenes
// (reference panda_gym/pybullet.py:52-55; scenes: panda_gym/envs/tasks/{reach,push,slide,pick_and_place,stack,flip}.py
// _create_scene; friction: panda_gym/envs/robots/panda.py:47-50, slide.py:34-42).
//
// Contact model (defined by oracle/panda_oracle.c, "contact model"): vertices of one body against the signed-distance
// field of the other (table plane, box, z-cylinder), speculative rows inside a 4 mm margin, two friction directions with an
// implicit cone, soft finger contacts, sequential impulses interleaved with the joint-limit and motor rows.
// Rows are kept in per-thread local memory (interleaved, so a warp's accesses coalesce); the robot part of a row is a
// 9-vector pair (J, M^-1 J^T), the free-body part is recomputed from the contact geometry each sweep.
#pragma once
#include "panda_dyn.cuh"

namespace pg {

constexpr int MAXOBJ = 2;
constexpr int MAXC = 24;    // contacts per env and sub-step
constexpr int MAXRC = 16;   // of which on the robot

enum { SH_BOX = 0, SH_CYL = 1 };
template <typename T> struct Scene {
    int nobj;
    int shape[MAXOBJ];
    T half[MAXOBJ][3];          // box half extents, or (r, r, h/2)
    T mass[MAXOBJ], Ic[MAXOBJ][3], mu[MAXOBJ];
    T table_x0, table_x1, table_y0, table_y1;
    T rb_c[3][3], rb_h[3][3], rb_mu[3];   // robot collision boxes: hand (link 8), finger 1, finger 2 -- centre / half extents in the link frame
    T margin, ground_z, table_mu;
    T soft_erp, soft_cfm;       // finger contact stiffness 30000 / damping 1000 -> erp, cfm/dt
};

template <typename T> struct Obj { V3<T> pos; T qx, qy, qz, qw; V3<T> lin, ang; };
template <typename T> struct Rot { V3<T> X, Y, Z; };   // columns
template <typename T> PG_HD Rot<T> quat_rot(T x, T y, T z, T w) {
    Rot<T> R;
    R.X = mk<T>(1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y));
    R.Y = mk<T>(2 * (x * y - w * z), 1 - 2 * (x * x + z * z), 2 * (y * z + w * x));
    R.Z = mk<T>(2 * (x * z + w * y), 2 * (y * z - w * x), 1 - 2 * (x * x + y * y));
    return R;
}
template <typename T> PG_HD V3<T> rot_mul(const Rot<T>& R, V3<T> u) { return R.X * u.x + R.Y * u.y + R.Z * u.z; }
template <typename T> PG_HD V3<T> rot_tmul(const Rot<T>& R, V3<T> u) { return mk<T>(dot(R.X, u), dot(R.Y, u), dot(R.Z, u)); }

// 0-link floating base of a btMultiBody: gravity, per-body damping, gyroscopic term; semi-implicit Euler on the velocity
template <typename T> PG_HD void obj_unconstrained(const Scene<T>& S, int o, Obj<T>& b, const Rot<T>& R) {
    const T k = Consts<T>::kdamp, dt = Consts<T>::dt;
    T kl = k + k * norm(b.lin), ka = k + k * norm(b.ang);
    V3<T> wl = rot_tmul(R, b.ang);
    V3<T> Iw = mk<T>(S.Ic[o][0] * wl.x, S.Ic[o][1] * wl.y, S.Ic[o][2] * wl.z);
    V3<T> g = cross(wl, Iw);
    V3<T> al = mk<T>(-(g.x + Iw.x * ka) / S.Ic[o][0], -(g.y + Iw.y * ka) / S.Ic[o][1], -(g.z + Iw.z * ka) / S.Ic[o][2]);
    b.ang = b.ang + rot_mul(R, al) * dt;
    b.lin.x += (-b.lin.x * kl) * dt; b.lin.y += (-b.lin.y * kl) * dt; b.lin.z += (-b.lin.z * kl - Consts<T>::g) * dt;
}
// stepPositionsMultiDof for the base: p += v dt, q <- exp(w dt) q
template <typename T> PG_HD void obj_integrate(Obj<T>& b) {
    const T dt = Consts<T>::dt;
    b.pos = b.pos + b.lin * dt;
    T fa = norm(b.ang);
    if (fa * dt > Consts<T>::pi / 4) fa = Consts<T>::pi / 4 / dt;
    T sc = fa < T(0.001) ? (T(0.5) * dt - dt * dt * dt * T(0.020833333333) * fa * fa) : sin(T(0.5) * fa * dt) / fa;
    T ax = b.ang.x * sc, ay = b.ang.y * sc, az = b.ang.z * sc, aw = cos(T(0.5) * fa * dt);
    T x = aw * b.qx + ax * b.qw + ay * b.qz - az * b.qy;
    T y = aw * b.qy - ax * b.qz + ay * b.qw + az * b.qx;
    T z = aw * b.qz + ax * b.qy - ay * b.qx + az * b.qw;
    T w = aw * b.qw - ax * b.qx - ay * b.qy - az * b.qz;
    T n = T(1) / sqrt(x * x + y * y + z * z + w * w);
    b.qx = x * n; b.qy = y * n; b.qz = z * n; b.qw = w * n;
}

template <typename T> PG_HD V3<T> box_vertex(const T* h, int k) { return mk<T>((k & 1) ? h[0] : -h[0], (k & 2) ? h[1] : -h[1], (k & 4) ? h[2] : -h[2]); }
template <typename T> PG_HD V3<T> obj_vertex(const Scene<T>& S, int o, int k) {
    if (S.shape[o] == SH_BOX) return box_vertex(S.half[o], k);
    const T c = S.half[o][0] * Consts<T>::k45;
    int a = k & 3;
    return mk<T>((a == 0 || a == 3) ? c : -c, (a < 2) ? c : -c, (k & 4) ? S.half[o][2] : -S.half[o][2]);
}
template <typename T> PG_HD T sdf_box(const T* h, V3<T> p, V3<T>& n) {
    T dx = fabs(p.x) - h[0], dy = fabs(p.y) - h[1], dz = fabs(p.z) - h[2];
    if (dx <= 0 && dy <= 0 && dz <= 0) {
        int a = 0; T d = dx;
        if (dy > d) { a = 1; d = dy; }
        if (dz > d) { a = 2; d = dz; }
        n = mk<T>(T(0), T(0), T(0));
        if (a == 0) n.x = p.x >= 0 ? T(1) : T(-1); else if (a == 1) n.y = p.y >= 0 ? T(1) : T(-1); else n.z = p.z >= 0 ? T(1) : T(-1);
        return d;
    }
    T ox = dx > 0 ? dx : T(0), oy = dy > 0 ? dy : T(0), oz = dz > 0 ? dz : T(0);
    T len = sqrt(ox * ox + oy * oy + oz * oz), inv = T(1) / len;
    n = mk<T>((p.x >= 0 ? ox : -ox) * inv, (p.y >= 0 ? oy : -oy) * inv, (p.z >= 0 ? oz : -oz) * inv);
    return len;
}
template <typename T> PG_HD T sdf_cyl(T r, T hz, V3<T> p, V3<T>& n) {
    T rho = sqrt(p.x * p.x + p.y * p.y);
    T dr = rho - r, dz = fabs(p.z) - hz;
    T rx = rho > T(1e-12) ? p.x / rho : T(1), ry = rho > T(1e-12) ? p.y / rho : T(0), sz = p.z >= 0 ? T(1) : T(-1);
    if (dr <= 0 && dz <= 0) { if (dr > dz) { n = mk<T>(rx, ry, T(0)); return dr; } n = mk<T>(T(0), T(0), sz); return dz; }
    T a = dr > 0 ? dr : T(0), b = dz > 0 ? dz : T(0), len = sqrt(a * a + b * b);
    n = mk<T>(rx * a / len, ry * a / len, sz * b / len);
    return len;
}
template <typename T> PG_HD T obj_sdf(const Scene<T>& S, int o, V3<T> p, V3<T>& n) {
    return S.shape[o] == SH_BOX ? sdf_box(S.half[o], p, n) : sdf_cyl(S.half[o][0], S.half[o][2], p, n);
}
template <typename T> PG_HD void plane_space(V3<T> n, V3<T>& p, V3<T>& q) {   // btPlaneSpace1
    if (fabs(n.z) > Consts<T>::k45) {
        T a = n.y * n.y + n.z * n.z, k = T(1) / sqrt(a);
        p = mk<T>(T(0), -n.z * k, n.y * k); q = mk<T>(a * k, -n.x * p.z, n.x * p.y);
    } else {
        T a = n.x * n.x + n.y * n.y, k = T(1) / sqrt(a);
        p = mk<T>(-n.y * k, n.x * k, T(0)); q = mk<T>(-n.z * p.y, n.z * p.x, a * k);
    }
}

// body codes: -1 static, 0..2 robot box (hand / finger 1 / finger 2), 3 + o object o
template <typename T> struct Contacts {
    int n, nr;
    T P[MAXC][3], D[MAXC][3][3];            // point, directions (normal, t1, t2)
    signed char a[MAXC], b[MAXC], ri[MAXC]; // bodies, robot-pool slot (-1: none)
    T invD[MAXC][3], rhs[MAXC][3], app[MAXC][3], mu[MAXC], cfm[MAXC];
    T Jr[MAXRC][3][ND], Wr[MAXRC][3][ND];
};

template <typename T, int NOBJ> struct World {
    Frame<T> F[7];              // arm link frames at the sub-step's q
    Rot<T> Rb[3]; V3<T> cb[3];  // robot collision boxes in the world
    Rot<T> Ro[NOBJ > 0 ? NOBJ : 1];
    T Iinv[NOBJ > 0 ? NOBJ : 1][6];   // world inverse inertia (xx,xy,xz,yy,yz,zz)
};
template <typename T> PG_HD V3<T> sym6_mul(const T* I, V3<T> w) {
    return mk<T>(I[0] * w.x + I[1] * w.y + I[2] * w.z, I[1] * w.x + I[3] * w.y + I[4] * w.z, I[2] * w.x + I[4] * w.y + I[5] * w.z);
}

// robot part of the Jacobian row of a unit force `d` at world point P on robot box rb (0 hand, 1/2 fingers)
template <typename T, int NOBJ> PG_HD void robot_point_jac(const World<T, NOBJ>& W, int rb, V3<T> P, V3<T> d, T sign, T* J) {
#pragma unroll
    for (int j = 0; j < 7; j++) J[j] = sign * dot(W.F[j].Z, cross(P - W.F[j].p, d));
    const T k = Consts<T>::k45;
    V3<T> hy = (W.F[6].X + W.F[6].Y) * k;   // hand y axis in the world
    T fd = dot(hy, d);
    J[7] = rb == 1 ? sign * fd : T(0);
    J[8] = rb == 2 ? -sign * fd : T(0);
}

template <typename T, int NOBJ>
PG_HD void add_contact(const Scene<T>& S, const World<T, NOBJ>& W, const T (*Minv)[ND], const T* qd, const Obj<T>* ob, Contacts<T>& C,
                       V3<T> P, V3<T> n, T dist, int A, int B, T mu, bool soft) {
    bool on_robot = (A >= 0 && A < 3) || (B >= 0 && B < 3);
    if (C.n >= MAXC || (on_robot && C.nr >= MAXRC)) return;
    int c = C.n++;
    int ri = -1;
    if (on_robot) ri = C.nr++;
    C.ri[c] = (signed char)ri; C.a[c] = (signed char)A; C.b[c] = (signed char)B; C.mu[c] = mu;
    C.P[c][0] = P.x; C.P[c][1] = P.y; C.P[c][2] = P.z;
    V3<T> t1, t2; plane_space(n, t1, t2);
    T erp = soft ? S.soft_erp : Consts<T>::erp, cfm = soft ? S.soft_cfm : T(0);
    C.cfm[c] = cfm;
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        V3<T> d = k == 0 ? n : (k == 1 ? t1 : t2);
        C.D[c][k][0] = d.x; C.D[c][k][1] = d.y; C.D[c][k][2] = d.z;
        T den = T(0), rel = T(0);
        if (on_robot) {
            T J[ND];
            if (A >= 0 && A < 3) robot_point_jac(W, A, P, d, T(1), J); else robot_point_jac(W, B, P, d, T(-1), J);
#pragma unroll
            for (int i = 0; i < ND; i++) {
                T w = T(0);
#pragma unroll
                for (int j = 0; j < ND; j++) w += Minv[i][j] * J[j];
                C.Jr[ri][k][i] = J[i]; C.Wr[ri][k][i] = w;
                den += J[i] * w; rel += J[i] * qd[i];
            }
        }
#pragma unroll
        for (int o = 0; o < NOBJ; o++) {
            T sg = (A == 3 + o) ? T(1) : ((B == 3 + o) ? T(-1) : T(0));
            if (sg != T(0)) {
                V3<T> r = P - ob[o].pos, rxd = cross(r, d);
                den += T(1) / S.mass[o] + dot(rxd, sym6_mul(W.Iinv[o], rxd));
                rel += sg * dot(d, ob[o].lin + cross(ob[o].ang, r));
            }
        }
        if (k == 0) {
            T inv = T(1) / (den + cfm);
            T pen = dist + T(1e-5), poserr = T(0), velerr = -rel;
            if (pen > 0) velerr -= pen * Consts<T>::inv_dt; else poserr = -pen * erp * Consts<T>::inv_dt;
            C.invD[c][0] = inv; C.rhs[c][0] = (poserr + velerr) * inv;
        } else {
            T inv = T(1) / den;
            C.invD[c][k] = inv; C.rhs[c][k] = -rel * inv;
        }
        C.app[c][k] = T(0);
    }
}

template <typename T, int NOBJ> PG_HD bool over_table(const Scene<T>& S, V3<T> p) { return p.x >= S.table_x0 && p.x <= S.table_x1 && p.y >= S.table_y0 && p.y <= S.table_y1; }

template <typename T, int NOBJ>
PG_HD void collect_contacts(const Scene<T>& S, const World<T, NOBJ>& W, const T (*Minv)[ND], const T* qd, const Obj<T>* ob, Contacts<T>& C) {
    C.n = 0; C.nr = 0;
    const V3<T> up = mk<T>(T(0), T(0), T(1));
    // 1. object vertices against the table top / ground plane
#pragma unroll
    for (int o = 0; o < NOBJ; o++) {
        for (int k = 0; k < 8; k++) {
            V3<T> P = rot_mul(W.Ro[o], obj_vertex(S, o, k)) + ob[o].pos;
            T plane = (over_table<T, NOBJ>(S, P) && P.z > T(-0.05)) ? T(0) : S.ground_z;
            T d = P.z - plane;
            if (d < S.margin) add_contact<T, NOBJ>(S, W, Minv, qd, ob, C, P, up, d, 3 + o, -1, S.mu[o] * S.table_mu, false);
        }
    }
    // 2. robot box vertices against the table top
    for (int b = 0; b < 3; b++) {
        T lowest = W.cb[b].z - (fabs(W.Rb[b].X.z) * S.rb_h[b][0] + fabs(W.Rb[b].Y.z) * S.rb_h[b][1] + fabs(W.Rb[b].Z.z) * S.rb_h[b][2]);
        if (lowest >= S.margin) continue;
        for (int k = 0; k < 8; k++) {
            V3<T> P = rot_mul(W.Rb[b], box_vertex(S.rb_h[b], k)) + W.cb[b];
            if (over_table<T, NOBJ>(S, P) && P.z < S.margin) add_contact<T, NOBJ>(S, W, Minv, qd, ob, C, P, up, P.z, b, -1, S.rb_mu[b] * S.table_mu, b > 0);
        }
    }
    // 3. robot box <-> object, both directions
    if (NOBJ > 0) {
        for (int b = 0; b < 3; b++) {
            T rbr = sqrt(S.rb_h[b][0] * S.rb_h[b][0] + S.rb_h[b][1] * S.rb_h[b][1] + S.rb_h[b][2] * S.rb_h[b][2]);
#pragma unroll
            for (int o = 0; o < NOBJ; o++) {
                T orad = sqrt(S.half[o][0] * S.half[o][0] + S.half[o][1] * S.half[o][1] + S.half[o][2] * S.half[o][2]);
                if (norm(W.cb[b] - ob[o].pos) > rbr + orad + S.margin) continue;
                T mu = S.rb_mu[b] * S.mu[o];
                for (int k = 0; k < 8; k++) {
                    V3<T> P = rot_mul(W.Rb[b], box_vertex(S.rb_h[b], k)) + W.cb[b];
                    V3<T> nl, pl = rot_tmul(W.Ro[o], P - ob[o].pos);
                    T d = obj_sdf(S, o, pl, nl);
                    if (d < S.margin) add_contact<T, NOBJ>(S, W, Minv, qd, ob, C, P, rot_mul(W.Ro[o], nl), d, b, 3 + o, mu, b > 0);
                }
                for (int k = 0; k < 8; k++) {
                    V3<T> P = rot_mul(W.Ro[o], obj_vertex(S, o, k)) + ob[o].pos;
                    V3<T> nl, pl = rot_tmul(W.Rb[b], P - W.cb[b]);
                    T d = sdf_box(S.rb_h[b], pl, nl);
                    if (d < S.margin) add_contact<T, NOBJ>(S, W, Minv, qd, ob, C, P, rot_mul(W.Rb[b], nl), d, 3 + o, b, mu, b > 0);
                }
            }
        }
    }
    // 4. object <-> object
    if (NOBJ == 2) {
        T r0 = sqrt(S.half[0][0] * S.half[0][0] + S.half[0][1] * S.half[0][1] + S.half[0][2] * S.half[0][2]);
        T r1 = sqrt(S.half[1][0] * S.half[1][0] + S.half[1][1] * S.half[1][1] + S.half[1][2] * S.half[1][2]);
        if (norm(ob[0].pos - ob[NOBJ - 1].pos) <= r0 + r1 + S.margin) {
            for (int a = 0; a < 2; a++) {
                int b = 1 - a;
                for (int k = 0; k < 8; k++) {
                    V3<T> P = rot_mul(W.Ro[a % NOBJ], obj_vertex(S, a, k)) + ob[a % NOBJ].pos;
                    V3<T> nl, pl = rot_tmul(W.Ro[b % NOBJ], P - ob[b % NOBJ].pos);
                    T d = obj_sdf(S, b, pl, nl);
                    if (d < S.margin) add_contact<T, NOBJ>(S, W, Minv, qd, ob, C, P, rot_mul(W.Ro[b % NOBJ], nl), d, 3 + a, 3 + b, S.mu[a] * S.mu[b], false);
                }
            }
        }
    }
}

// J . dv of row (c, k)
template <typename T, int NOBJ>
PG_HD T row_jdv(const Contacts<T>& C, int c, int k, const T* dvq, const V3<T>* dvl, const V3<T>* dva, const Obj<T>* ob) {
    T jd = T(0);
    int ri = C.ri[c];
    if (ri >= 0) {
#pragma unroll
        for (int i = 0; i < ND; i++) jd += C.Jr[ri][k][i] * dvq[i];
    }
    V3<T> d = mk<T>(C.D[c][k][0], C.D[c][k][1], C.D[c][k][2]);
#pragma unroll
    for (int o = 0; o < NOBJ; o++) {
        T sg = (C.a[c] == 3 + o) ? T(1) : ((C.b[c] == 3 + o) ? T(-1) : T(0));
        if (sg != T(0)) { V3<T> r = mk<T>(C.P[c][0], C.P[c][1], C.P[c][2]) - ob[o].pos; jd += sg * dot(d, dvl[o] + cross(dva[o], r)); }
    }
    return jd;
}
template <typename T, int NOBJ>
PG_HD void row_apply(const Scene<T>& S, const World<T, NOBJ>& W, const Contacts<T>& C, int c, int k, T di, T* dvq, V3<T>* dvl, V3<T>* dva, const Obj<T>* ob) {
    int ri = C.ri[c];
    if (ri >= 0) {
#pragma unroll
        for (int i = 0; i < ND; i++) dvq[i] += C.Wr[ri][k][i] * di;
    }
    V3<T> d = mk<T>(C.D[c][k][0], C.D[c][k][1], C.D[c][k][2]);
#pragma unroll
    for (int o = 0; o < NOBJ; o++) {
        T sg = (C.a[c] == 3 + o) ? T(1) : ((C.b[c] == 3 + o) ? T(-1) : T(0));
        if (sg != T(0)) {
            V3<T> r = mk<T>(C.P[c][0], C.P[c][1], C.P[c][2]) - ob[o].pos;
            dvl[o] = dvl[o] + d * (sg * di / S.mass[o]);
            dva[o] = dva[o] + sym6_mul(W.Iinv[o], cross(r, d)) * (sg * di);
        }
    }
}

// ---------------------------------------------------------------------------------------------- full sub-step
// One 2 ms stepSimulation: unconstrained velocities, contact generation at the current poses, <= 50 sequential-impulse sweeps
// over [joint limits, motors] (direction alternating), contact normals, friction cones; exit when the largest squared
// velocity change of a sweep is <= 1e-7; semi-implicit Euler.
template <typename T, int NOBJ>
PG_HD void env_substep(const Model<T>& M, const Scene<T>& S, T* q, T* qd, const T* target, Obj<T>* ob, Contacts<T>& C) {
    T sn[7], cs[7], Minv[ND][ND], qdd[ND];
    robot_dynamics(M, q, qd, sn, cs, Minv, qdd);
#pragma unroll
    for (int d = 0; d < ND; d++) qd[d] += qdd[d] * Consts<T>::dt;
    World<T, NOBJ> W;
    {   // world frames from the sines / cosines already computed
        Frame<T> B; B.X = mk<T>(1, 0, 0); B.Y = mk<T>(0, 1, 0); B.Z = mk<T>(0, 0, 1); B.p = ld3(M.base);
        W.F[0] = fk_next<0>(M, B, sn[0], cs[0]); W.F[1] = fk_next<1>(M, W.F[0], sn[1], cs[1]); W.F[2] = fk_next<2>(M, W.F[1], sn[2], cs[2]);
        W.F[3] = fk_next<3>(M, W.F[2], sn[3], cs[3]); W.F[4] = fk_next<4>(M, W.F[3], sn[4], cs[4]); W.F[5] = fk_next<5>(M, W.F[4], sn[5], cs[5]);
        W.F[6] = fk_next<6>(M, W.F[5], sn[6], cs[6]);
        const T k = Consts<T>::k45;
        Rot<T> Rh; Rh.X = (W.F[6].X - W.F[6].Y) * k; Rh.Y = (W.F[6].X + W.F[6].Y) * k; Rh.Z = W.F[6].Z;
        V3<T> ph = W.F[6].p + W.F[6].Z * (M.hz - T(0.0584));   // hand frame origin
        V3<T> pf = W.F[6].p + W.F[6].Z * M.hz;
#pragma unroll
        for (int b = 0; b < 3; b++) {
            W.Rb[b] = Rh;
            V3<T> o = b == 0 ? ph : (b == 1 ? pf + Rh.Y * q[7] : pf - Rh.Y * q[8]);
            W.cb[b] = o + rot_mul(Rh, ld3(S.rb_c[b]));
        }
    }
#pragma unroll
    for (int o = 0; o < NOBJ; o++) {
        W.Ro[o] = quat_rot(ob[o].qx, ob[o].qy, ob[o].qz, ob[o].qw);
        obj_unconstrained(S, o, ob[o], W.Ro[o]);
        const Rot<T>& R = W.Ro[o];
        T ix = T(1) / S.Ic[o][0], iy = T(1) / S.Ic[o][1], iz = T(1) / S.Ic[o][2];
        W.Iinv[o][0] = ix * R.X.x * R.X.x + iy * R.Y.x * R.Y.x + iz * R.Z.x * R.Z.x;
        W.Iinv[o][1] = ix * R.X.x * R.X.y + iy * R.Y.x * R.Y.y + iz * R.Z.x * R.Z.y;
        W.Iinv[o][2] = ix * R.X.x * R.X.z + iy * R.Y.x * R.Y.z + iz * R.Z.x * R.Z.z;
        W.Iinv[o][3] = ix * R.X.y * R.X.y + iy * R.Y.y * R.Y.y + iz * R.Z.y * R.Z.y;
        W.Iinv[o][4] = ix * R.X.y * R.X.z + iy * R.Y.y * R.Y.z + iz * R.Z.y * R.Z.z;
        W.Iinv[o][5] = ix * R.X.z * R.X.z + iy * R.Y.z * R.Y.z + iz * R.Z.z * R.Z.z;
    }
    JointRows<T> R;
    joint_rows_setup(M, q, qd, target, Minv, R);
    collect_contacts<T, NOBJ>(S, W, Minv, qd, ob, C);

    T dvq[ND];
    V3<T> dvl[NOBJ > 0 ? NOBJ : 1], dva[NOBJ > 0 ? NOBJ : 1];
#pragma unroll
    for (int d = 0; d < ND; d++) dvq[d] = T(0);
#pragma unroll
    for (int o = 0; o < NOBJ; o++) { dvl[o] = mk<T>(0, 0, 0); dva[o] = mk<T>(0, 0, 0); }
    const int nc = C.n;
    for (int it = 0; it < 50; it++) {
        T res = T(0);
        joint_rows_sweep(M, Minv, R, dvq, it, res);
        for (int c = 0; c < nc; c++) {          // contact normals
            T jd = row_jdv<T, NOBJ>(C, c, 0, dvq, dvl, dva, ob);
            T app = C.app[c][0];
            T di = C.rhs[c][0] - app * C.cfm[c] - jd * C.invD[c][0];
            T sum = app + di;
            if (sum < T(0)) { di = -app; sum = T(0); }
            C.app[c][0] = sum;
            row_apply<T, NOBJ>(S, W, C, c, 0, di, dvq, dvl, dva, ob);
            T r = di / C.invD[c][0]; res = fmax(res, r * r);
        }
        for (int c = 0; c < nc; c++) {          // implicit friction cone over the two tangent rows
            T napp = C.app[c][0];
            if (napp <= T(0)) continue;
            T j1 = row_jdv<T, NOBJ>(C, c, 1, dvq, dvl, dva, ob), j2 = row_jdv<T, NOBJ>(C, c, 2, dvq, dvl, dva, ob);
            T a1 = C.app[c][1], a2 = C.app[c][2];
            T s1 = a1 + C.rhs[c][1] - j1 * C.invD[c][1], s2 = a2 + C.rhs[c][2] - j2 * C.invD[c][2];
            T lim = C.mu[c] * napp, len = sqrt(s1 * s1 + s2 * s2);
            if (len > lim) { T f = lim / len; s1 *= f; s2 *= f; }
            T d1 = s1 - a1, d2 = s2 - a2;
            C.app[c][1] = s1; C.app[c][2] = s2;
            row_apply<T, NOBJ>(S, W, C, c, 1, d1, dvq, dvl, dva, ob);
            row_apply<T, NOBJ>(S, W, C, c, 2, d2, dvq, dvl, dva, ob);
            T r1 = d1 / C.invD[c][1], r2 = d2 / C.invD[c][2];
            res = fmax(res, fmax(r1 * r1, r2 * r2));
        }
        if (res <= T(1e-7)) break;
    }
#pragma unroll
    for (int d = 0; d < ND; d++) { qd[d] += dvq[d]; q[d] += qd[d] * Consts<T>::dt; }
#pragma unroll
    for (int o = 0; o < NOBJ; o++) { ob[o].lin = ob[o].lin + dvl[o]; ob[o].ang = ob[o].ang + dva[o]; obj_integrate(ob[o]); }
}

}  // namespace pg
