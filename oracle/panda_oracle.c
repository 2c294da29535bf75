/*
 * panda_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See panda_oracle.h.
 *
 * Plain C, fp64.  Restates, for exactly the bodies the six panda_gym tasks create, what the
 * reference's hot path computes through pybullet==3.2.5 (env.yml:107; the engine is NOT in
 * /root/reference):
 *   - model construction as loadURDF does it: Featherstone multibody in
 *     link frames, link inertia = solid-box inertia of the collision AABB (the meshes are absent
 *     here, so the AABB extents below are approximations; link 5's is calibrated to the
 *     reference's KAT test/pybullet_test.py:186/:203),
 *   - btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof (ABA, per-link damping),
 *   - btMultiBody::calcAccelerationDeltasMultiDof (unit impulse responses -> M^-1),
 *   - btMultiBodyJointLimitConstraint / btMultiBodyJointMotor rows and the sequential-impulse
 *     (PGS) sweep of btMultiBodyConstraintSolver with <=50 iterations,
 *   - contact rows (normal + 2 friction directions, implicit cone) -- see "contact model" below,
 *   - stepPositionsMultiDof (semi-implicit Euler, quaternion exponential for free bodies),
 *   - the stale link-transform cache read by getLinkState,
 *   - calculateInverseKinematics = 20 x BussIK DLS,
 *   - panda_gym/envs/robots/panda.py:52-140, envs/core.py:229-289, envs/tasks/ (all six tasks), utils.py:4-30.
 * Pinned by tests/test_oracle_kat.py against test/pybullet_test.py:34,:64,:135,:152,:169,:186,
 * :203,:265.  Everything involving contact is "parity unpinned" (no golden vectors exist).
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -shared -fPIC (oracle/Makefile).  -ffp-contract=off matters:
 * the float32 reward arithmetic must not be fused.
 */
#include "panda_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NL 12
#define ND 9
#define MAXOBJ 2
#define NDT (ND + 6 * MAXOBJ)
#define MAXROWS 512
#define DT (1.0 / 500.0)
#define GRAV 9.81
#define PI 3.14159265358979323846
#define GROUND_Z (-0.4)   /* create_plane(z_offset=-0.4) in every task (reach.py:29 ...) */

/* ------------------------------------------------------------------ small linear algebra */
typedef double V3[3];
static void v3set(double *o, double x, double y, double z) { o[0] = x; o[1] = y; o[2] = z; }
static void v3cpy(double *o, const double *a) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
static void v3add(double *o, const double *a, const double *b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static void v3sub(double *o, const double *a, const double *b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static void v3scale(double *o, const double *a, double s) { o[0] = a[0] * s; o[1] = a[1] * s; o[2] = a[2] * s; }
static double v3dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double v3norm(const double *a) { return sqrt(v3dot(a, a)); }
static void v3cross(double *o, const double *a, const double *b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
/* 3x3 row-major */
static void m3mulv(double *o, const double *R, const double *v) {
    double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2], z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void m3Tmulv(double *o, const double *R, const double *v) {
    double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2], y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2], z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void m3mul(double *o, const double *A, const double *B) {
    double t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(o, t, sizeof t);
}
static void m3T(double *o, const double *A) {
    double t[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
    memcpy(o, t, sizeof t);
}
static void m3ident(double *o) { memset(o, 0, 9 * sizeof(double)); o[0] = o[4] = o[8] = 1; }
static void rpy_to_R(double *R, const double *rpy) { /* URDF: R = Rz(yaw) Ry(pitch) Rx(roll) */
    double cr = cos(rpy[0]), sr = sin(rpy[0]), cp = cos(rpy[1]), sp = sin(rpy[1]), cy = cos(rpy[2]), sy = sin(rpy[2]);
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}
static void quat_to_R(double *R, const double *q) { /* q = (x,y,z,w) */
    double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void R_to_quat(double *q, const double *R) { /* Shepperd; returns (x,y,z,w) with w >= 0 when the trace branch is taken */
    double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        double s = sqrt(tr + 1.0) * 2; q[3] = 0.25 * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[3] = (R[7] - R[5]) / s; q[0] = 0.25 * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25 * s; q[2] = (R[5] + R[7]) / s;
    } else {
        double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25 * s;
    }
}
static void quat_mul(double *o, const double *a, const double *b) { /* (x,y,z,w) Hamilton product a*b */
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

/* ------------------------------------------------------------------ Panda model */
enum { J_FIXED = 0, J_REV = 1, J_PRIS = 2 };
typedef struct {
    int parent, jtype, dof;
    double xyz[3], rpy[3], axis[3], lo, hi, mass, com[3], box[3];
} LinkDef;

/* box[] = extents of the collision-shape AABB in link axes: Bullet derives the link inertia from it
 * (SURVEY App. B.1).  The meshes are not available in this container: approximations from the Panda's
 * link dimensions; box[] of link 5 is calibrated so that the KAT test/pybullet_test.py:186 holds. */
static LinkDef LINKS[NL] = {
    /* 0  panda_joint1 */ {-1, J_REV, 0, {0, 0, 0.333}, {0, 0, 0}, {0, 0, 1}, -2.9671, 2.9671, 2.7, {0, -0.04, -0.05}, {0.11, 0.13, 0.25}},
    /* 1  panda_joint2 */ {0, J_REV, 1, {0, 0, 0}, {-PI / 2, 0, 0}, {0, 0, 1}, -1.8326, 1.8326, 2.73, {0, -0.04, 0.06}, {0.11, 0.25, 0.13}},
    /* 2  panda_joint3 */ {1, J_REV, 2, {0, -0.316, 0}, {PI / 2, 0, 0}, {0, 0, 1}, -2.9671, 2.9671, 2.04, {0.01, 0.01, -0.05}, {0.19, 0.15, 0.18}},
    /* 3  panda_joint4 */ {2, J_REV, 3, {0.0825, 0, 0}, {PI / 2, 0, 0}, {0, 0, 1}, -3.1416, 0.0, 2.08, {-0.03, 0.03, 0.02}, {0.19, 0.18, 0.15}},
    /* 4  panda_joint5 */ {3, J_REV, 4, {-0.0825, 0.384, 0}, {-PI / 2, 0, 0}, {0, 0, 1}, -2.9671, 2.9671, 3.0, {0, 0.04, -0.12}, {0.11, 0.19, 0.32}},
    /* 5  panda_joint6 */ {4, J_REV, 5, {0, 0, 0}, {PI / 2, 0, 0}, {0, 0, 1}, -0.0873, 3.8223, 1.3, {0.04, 0, 0}, {0.20265085784266038, 0.13, 0.12}},
    /* 6  panda_joint7 */ {5, J_REV, 6, {0.088, 0, 0}, {PI / 2, 0, 0}, {0, 0, 1}, -2.9671, 2.9671, 0.2, {0, 0, 0.08}, {0.11, 0.11, 0.10}},
    /* 7  panda_joint8 */ {6, J_FIXED, -1, {0, 0, 0.107}, {0, 0, 0}, {0, 0, 0}, 0, 0, 0.0, {0, 0, 0}, {0, 0, 0}},
    /* 8  hand         */ {7, J_FIXED, -1, {0, 0, 0}, {0, 0, -PI / 4}, {0, 0, 0}, 0, 0, 0.81, {0, 0, 0.04}, {0.064, 0.204, 0.09}},
    /* 9  finger1      */ {8, J_PRIS, 7, {0, 0, 0.0584}, {0, 0, 0}, {0, 1, 0}, 0.0, 0.04, 0.1, {0, 0.01, 0.02}, {0.021, 0.021, 0.054}},
    /* 10 finger2      */ {8, J_PRIS, 8, {0, 0, 0.0584}, {0, 0, 0}, {0, -1, 0}, 0.0, 0.04, 0.1, {0, -0.01, 0.02}, {0.021, 0.021, 0.054}},
    /* 11 grasptarget  */ {8, J_FIXED, -1, {0, 0, 0.105}, {0, 0, 0}, {0, 0, 0}, 0, 0, 0.0, {0, 0, 0}, {0, 0, 0}},
};
static const int DOF_LINK[ND] = {0, 1, 2, 3, 4, 5, 6, 9, 10};
static double LINK_INERTIA[NL][3];
static int model_ready = 0;
static void model_init(void) {
    if (model_ready) return;
    for (int i = 0; i < NL; i++) {
        const double *b = LINKS[i].box; double m = LINKS[i].mass / 12.0;
        LINK_INERTIA[i][0] = m * (b[1] * b[1] + b[2] * b[2]);
        LINK_INERTIA[i][1] = m * (b[0] * b[0] + b[2] * b[2]);
        LINK_INERTIA[i][2] = m * (b[0] * b[0] + b[1] * b[1]);
    }
    model_ready = 1;
}
void po_set_link_inertia(int link, double ixx, double iyy, double izz) { model_init(); v3set(LINK_INERTIA[link], ixx, iyy, izz); }
void po_get_link_inertia(int link, double out[3]) { model_init(); v3cpy(out, LINK_INERTIA[link]); }
void po_get_link_def(int link, double out[16]) { /* parent,jtype,xyz3,rpy3(roll,yaw only used),lo,hi,mass,com3 */
    const LinkDef *L = &LINKS[link];
    out[0] = L->parent; out[1] = L->jtype; v3cpy(out + 2, L->xyz); v3cpy(out + 5, L->rpy); out[8] = L->lo; out[9] = L->hi; out[10] = L->mass; v3cpy(out + 11, L->com);
    out[14] = L->axis[1]; out[15] = L->dof;
}

/* ------------------------------------------------------------------ objects / scene */
enum { SH_BOX = 0, SH_CYL = 1 };
typedef struct {
    int shape; double half[3]; /* box half extents, or (r, r, h/2) for the z-cylinder */
    double mass, Ic[3], mu;
    double pos[3], quat[4], lin[3], ang[3];
} Obj;

typedef struct { double J[NDT], W[NDT], rhs, cfm, invD, lo, hi, applied, mu; int normal_row; } Row;

struct PoSim {
    int task; double base[3];
    double q[ND], qd[ND], qc[ND];
    double m_kp[ND], m_kd[ND], m_tq[ND], m_tv[ND], m_maximp[ND];
    int nobj; Obj obj[MAXOBJ];
    double table_x0, table_x1, table_y0, table_y1, ground_z;   /* table top (z = 0) rectangle (empty: x0 > x1), ground plane height (-1e30: none) */
    int last_contacts, last_robot_contacts, last_iters;
    /* scratch of the last forward-dynamics pass (positions at the start of the sub-step) */
    double E[NL][9], r[NL][3], Rw[NL][9], pw[NL][3], S[NL][6];
    double IA[NL][36], U[NL][6], D[NL], u[NL], v[NL][6], c[NL][6], pA[NL][6];
    double Minv[ND][ND];
    Row rows[MAXROWS]; int nrows;
    double dbg[64][10]; int n_noncontact;   /* per contact: P, n, dist, A-link, B-obj, first row index */
};

/* ------------------------------------------------------------------ kinematics */
/* pose of every link frame for joint vector q: Rw (world<-link), pw (world), plus parent->child E, r */
static void fk_all(const PoSim *s, const double *q, double E[NL][9], double r[NL][3], double Rw[NL][9], double pw[NL][3]) {
    for (int i = 0; i < NL; i++) {
        const LinkDef *L = &LINKS[i];
        double RT[9], Rpc[9];
        rpy_to_R(RT, L->rpy);
        v3cpy(r[i], L->xyz);
        if (L->jtype == J_REV) {
            double a = q[L->dof], Rz[9] = {cos(a), -sin(a), 0, sin(a), cos(a), 0, 0, 0, 1};
            m3mul(Rpc, RT, Rz);
        } else {
            memcpy(Rpc, RT, sizeof Rpc);
            if (L->jtype == J_PRIS) { double d[3], ax[3]; v3scale(ax, L->axis, q[L->dof]); m3mulv(d, RT, ax); v3add(r[i], r[i], d); }
        }
        m3T(E[i], Rpc);
        if (L->parent < 0) { memcpy(Rw[i], Rpc, sizeof Rpc); v3add(pw[i], s->base, r[i]); }
        else { double t[3]; m3mul(Rw[i], Rw[L->parent], Rpc); m3mulv(t, Rw[L->parent], r[i]); v3add(pw[i], pw[L->parent], t); }
    }
}

/* spatial helpers, 6-vectors = [angular; linear] in link coordinates at the link-frame origin */
static void xmotion(double *o, const double *E, const double *r, const double *v) { /* parent -> child */
    double t[3], w[3], l[3];
    m3mulv(w, E, v);
    v3cross(t, r, v); v3sub(t, v + 3, t); m3mulv(l, E, t);
    v3cpy(o, w); v3cpy(o + 3, l);
}
static void xforceT(double *o, const double *E, const double *r, const double *f) { /* child -> parent (X^T f) */
    double n[3], l[3], t[3];
    m3Tmulv(l, E, f + 3); m3Tmulv(n, E, f); v3cross(t, r, l); v3add(n, n, t);
    v3cpy(o, n); v3cpy(o + 3, l);
}
static void crm(double *o, const double *v, const double *m) { /* v x m (motion) */
    double a[3], b[3], c[3];
    v3cross(a, v, m); v3cross(b, v, m + 3); v3cross(c, v + 3, m); v3add(b, b, c);
    v3cpy(o, a); v3cpy(o + 3, b);
}
static void crf(double *o, const double *v, const double *f) { /* v x* f (force) */
    double a[3], b[3], c[3];
    v3cross(a, v, f); v3cross(b, v + 3, f + 3); v3add(a, a, b); v3cross(c, v, f + 3);
    v3cpy(o, a); v3cpy(o + 3, c);
}
static void m6mulv(double *o, const double *M, const double *v) {
    double t[6];
    for (int i = 0; i < 6; i++) { double a = 0; for (int j = 0; j < 6; j++) a += M[6 * i + j] * v[j]; t[i] = a; }
    memcpy(o, t, sizeof t);
}
static void link_spatial_inertia(int i, double *I) {
    const LinkDef *L = &LINKS[i]; double m = L->mass; const double *c = L->com;
    double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0}, cxT[9], cc[9];
    m3T(cxT, cx); m3mul(cc, cx, cxT);
    memset(I, 0, 36 * sizeof(double));
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
        I[6 * a + b] = m * cc[3 * a + b] + (a == b ? LINK_INERTIA[i][a] : 0);
        I[6 * a + 3 + b] = m * cx[3 * a + b];
        I[6 * (a + 3) + b] = m * cxT[3 * a + b];
        I[6 * (a + 3) + 3 + b] = (a == b) ? m : 0;
    }
}
/* o += X^T A X  (X = motion transform parent->child built from E,r) */
static void add_XtAX(double *o, const double *E, const double *r, const double *A) {
    double X[36]; memset(X, 0, sizeof X);
    double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0}, Erx[9];
    m3mul(Erx, E, rx);
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { X[6 * a + b] = E[3 * a + b]; X[6 * (a + 3) + 3 + b] = E[3 * a + b]; X[6 * (a + 3) + b] = -Erx[3 * a + b]; }
    double AX[36];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double a = 0; for (int k = 0; k < 6; k++) a += A[6 * i + k] * X[6 * k + j]; AX[6 * i + j] = a; }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double a = 0; for (int k = 0; k < 6; k++) a += X[6 * k + i] * AX[6 * k + j]; o[6 * i + j] += a; }
}

/* Featherstone ABA (btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof).  Gravity enters as a
 * fictitious base acceleration; Bullet's per-link damping force m v (k + k|v|), I w (k + k|w|) with
 * k = 0.04 acts at each link's centre of mass.  tau = joint forces.  Leaves IA/U/D for the
 * unit-impulse responses. */
#define LINK_DAMP 0.04
static void aba(PoSim *s, const double *tau, double *qdd) {
    double a[NL][6];
    fk_all(s, s->q, s->E, s->r, s->Rw, s->pw);
    for (int i = 0; i < NL; i++) {
        const LinkDef *L = &LINKS[i];
        double vp[6] = {0, 0, 0, 0, 0, 0}, vj[6] = {0, 0, 0, 0, 0, 0};
        memset(s->S[i], 0, sizeof s->S[i]);
        if (L->jtype == J_REV) s->S[i][2] = 1; else if (L->jtype == J_PRIS) v3cpy(s->S[i] + 3, L->axis);
        if (L->parent >= 0) xmotion(vp, s->E[i], s->r[i], s->v[L->parent]);
        if (L->dof >= 0) for (int k = 0; k < 6; k++) vj[k] = s->S[i][k] * s->qd[L->dof];
        for (int k = 0; k < 6; k++) s->v[i][k] = vp[k] + vj[k];
        crm(s->c[i], s->v[i], vj);
        link_spatial_inertia(i, s->IA[i]);
        double Iv[6]; m6mulv(Iv, s->IA[i], s->v[i]); crf(s->pA[i], s->v[i], Iv);
        /* damping at the CoM, expressed at the link origin */
        double vc[3], t[3], fl[3], fa[3], n[3];
        v3cross(t, s->v[i], L->com); v3add(vc, s->v[i] + 3, t);
        double kl = LINK_DAMP + LINK_DAMP * v3norm(vc), ka = LINK_DAMP + LINK_DAMP * v3norm(s->v[i]);
        v3scale(fl, vc, L->mass * kl);
        for (int k = 0; k < 3; k++) fa[k] = LINK_INERTIA[i][k] * s->v[i][k] * ka;
        v3cross(n, L->com, fl); v3add(n, n, fa);
        for (int k = 0; k < 3; k++) { s->pA[i][k] += n[k]; s->pA[i][3 + k] += fl[k]; }
    }
    for (int i = NL - 1; i >= 0; i--) {
        const LinkDef *L = &LINKS[i];
        double Ia[36], pa[6];
        memcpy(Ia, s->IA[i], sizeof Ia); memcpy(pa, s->pA[i], sizeof pa);
        if (L->dof >= 0) {
            m6mulv(s->U[i], s->IA[i], s->S[i]);
            double d = 0, sp = 0;
            for (int k = 0; k < 6; k++) { d += s->S[i][k] * s->U[i][k]; sp += s->S[i][k] * s->pA[i][k]; }
            s->D[i] = d; s->u[i] = tau[L->dof] - sp;
            for (int x = 0; x < 6; x++) for (int y = 0; y < 6; y++) Ia[6 * x + y] -= s->U[i][x] * s->U[i][y] / d;
        }
        double Iac[6]; m6mulv(Iac, Ia, s->c[i]);
        for (int k = 0; k < 6; k++) pa[k] += Iac[k] + (L->dof >= 0 ? s->U[i][k] * s->u[i] / s->D[i] : 0);
        if (L->parent >= 0) {
            double f[6]; add_XtAX(s->IA[L->parent], s->E[i], s->r[i], Ia);
            xforceT(f, s->E[i], s->r[i], pa);
            for (int k = 0; k < 6; k++) s->pA[L->parent][k] += f[k];
        }
    }
    for (int i = 0; i < NL; i++) {
        const LinkDef *L = &LINKS[i];
        double ap[6];
        if (L->parent >= 0) xmotion(ap, s->E[i], s->r[i], a[L->parent]);
        else { double a0[6] = {0, 0, 0, 0, 0, GRAV}; double zero[3] = {0, 0, 0}; xmotion(ap, s->E[i], zero, a0); /* base frame axes == world */ }
        for (int k = 0; k < 6; k++) a[i][k] = ap[k] + s->c[i][k];
        if (L->dof >= 0) {
            double ua = 0; for (int k = 0; k < 6; k++) ua += s->U[i][k] * a[i][k];
            double dd = (s->u[i] - ua) / s->D[i];
            qdd[L->dof] = dd;
            for (int k = 0; k < 6; k++) a[i][k] += s->S[i][k] * dd;
        }
    }
}
/* btMultiBody::calcAccelerationDeltasMultiDof: velocity change for spatial impulses f[i] (link coords, acting ON link i)
 * and joint impulses tau, reusing IA/U/D of the last aba(). */
static void impulse_response(const PoSim *s, double f[NL][6], const double *tau, double *dqd) {
    double p[NL][6], uu[NL], a[NL][6];
    for (int i = 0; i < NL; i++) for (int k = 0; k < 6; k++) p[i][k] = f ? -f[i][k] : 0;
    for (int i = NL - 1; i >= 0; i--) {
        const LinkDef *L = &LINKS[i];
        double pa[6]; memcpy(pa, p[i], sizeof pa);
        if (L->dof >= 0) {
            double sp = 0; for (int k = 0; k < 6; k++) sp += s->S[i][k] * p[i][k];
            uu[i] = tau[L->dof] - sp;
            for (int k = 0; k < 6; k++) pa[k] += s->U[i][k] * uu[i] / s->D[i];
        }
        if (L->parent >= 0) { double t[6]; xforceT(t, s->E[i], s->r[i], pa); for (int k = 0; k < 6; k++) p[L->parent][k] += t[k]; }
    }
    for (int i = 0; i < NL; i++) {
        const LinkDef *L = &LINKS[i];
        if (L->parent >= 0) xmotion(a[i], s->E[i], s->r[i], a[L->parent]); else memset(a[i], 0, sizeof a[i]);
        if (L->dof >= 0) {
            double ua = 0; for (int k = 0; k < 6; k++) ua += s->U[i][k] * a[i][k];
            double dd = (uu[i] - ua) / s->D[i];
            dqd[L->dof] = dd;
            for (int k = 0; k < 6; k++) a[i][k] += s->S[i][k] * dd;
        }
    }
}

/* ------------------------------------------------------------------ creation */
static void obj_init(Obj *o, int shape, double hx, double hy, double hz, double mass, double mu) {
    memset(o, 0, sizeof *o);
    o->shape = shape; v3set(o->half, hx, hy, hz); o->mass = mass; o->mu = mu; o->quat[3] = 1;
    if (shape == SH_BOX) { /* btBoxShape::calculateLocalInertia */
        double lx = 2 * hx, ly = 2 * hy, lz = 2 * hz;
        v3set(o->Ic, mass / 12 * (ly * ly + lz * lz), mass / 12 * (lx * lx + lz * lz), mass / 12 * (lx * lx + ly * ly));
    } else { /* btCylinderShapeZ */
        double r = hx, h = 2 * hz;
        v3set(o->Ic, mass / 12 * h * h + mass / 4 * r * r, mass / 12 * h * h + mass / 4 * r * r, mass / 2 * r * r);
    }
}
PoSim *po_create(int task, double bx, double by, double bz) {
    model_init();
    PoSim *s = (PoSim *)calloc(1, sizeof(PoSim));
    s->task = task; v3set(s->base, bx, by, bz);
    for (int d = 0; d < ND; d++) { s->m_kp[d] = 0; s->m_kd[d] = 1; s->m_maximp[d] = 1.0; } /* loadURDF default velocity motors (App. B.1) */
    /* table top rectangle (pybullet.py:741-771; slide.py:33) */
    s->table_x0 = -0.85; s->table_x1 = 0.25; s->table_y0 = -0.35; s->table_y1 = 0.35; s->ground_z = GROUND_Z;
    if (task == PO_BARE) { s->table_x0 = 1; s->table_x1 = -1; s->table_y0 = 1; s->table_y1 = -1; s->ground_z = -1e30; }   /* PyBullet() alone: no plane, no table */
    switch (task) {
    case PO_PUSH: case PO_PICK_AND_PLACE: case PO_FLIP:
        s->nobj = 1; obj_init(&s->obj[0], SH_BOX, 0.02, 0.02, 0.02, 1.0, 0.5); s->obj[0].pos[2] = 0.02; break;
    case PO_SLIDE:
        s->table_x0 = -0.8; s->table_x1 = 0.6;
        s->nobj = 1; obj_init(&s->obj[0], SH_CYL, 0.03, 0.03, 0.015, 1.0, 0.04); s->obj[0].pos[2] = 0.015; break;
    case PO_STACK:
        s->nobj = 2; obj_init(&s->obj[0], SH_BOX, 0.02, 0.02, 0.02, 2.0, 0.5); obj_init(&s->obj[1], SH_BOX, 0.02, 0.02, 0.02, 1.0, 0.5);
        s->obj[0].pos[2] = 0.02; s->obj[1].pos[0] = 0.5; s->obj[1].pos[2] = 0.02; break;
    default: s->nobj = 0;
    }
    return s;
}
/* stand-alone free body for facade-level tests (pybullet.py:531-719 create_box/create_cylinder) */
int po_add_box(PoSim *s, double hx, double hy, double hz, double mass, const double pos[3]) {
    if (s->nobj >= MAXOBJ) return -1;
    obj_init(&s->obj[s->nobj], SH_BOX, hx, hy, hz, mass, 0.5); v3cpy(s->obj[s->nobj].pos, pos);
    return s->nobj++;
}
/* create_table / create_plane of a bare world (pybullet.py:726-771): NULL removes the surface */
void po_set_static(PoSim *s, const double *table_rect, const double *ground_z) {
    if (table_rect) { s->table_x0 = table_rect[0]; s->table_x1 = table_rect[1]; s->table_y0 = table_rect[2]; s->table_y1 = table_rect[3]; }
    else { s->table_x0 = 1; s->table_x1 = -1; s->table_y0 = 1; s->table_y1 = -1; }
    s->ground_z = ground_z ? *ground_z : -1e30;
}
/* raw joint state incl. the link-transform cache (test hook: getLinkState at an arbitrary state) */
void po_set_joint_state(PoSim *s, const double *q, const double *qd, const double *qc) { memcpy(s->q, q, sizeof s->q); memcpy(s->qd, qd, sizeof s->qd); memcpy(s->qc, qc, sizeof s->qc); }
void po_set_object_shape(PoSim *s, int o, int shape, double hx, double hy, double hz, double mass, double mu) {
    Obj keep = s->obj[o]; obj_init(&s->obj[o], shape, hx, hy, hz, mass, mu);
    v3cpy(s->obj[o].pos, keep.pos); memcpy(s->obj[o].quat, keep.quat, sizeof keep.quat); v3cpy(s->obj[o].lin, keep.lin); v3cpy(s->obj[o].ang, keep.ang);
}
void po_destroy(PoSim *s) { free(s); }
int po_num_objects(const PoSim *s) { return s->nobj; }
void po_reset_joint(PoSim *s, int link, double angle) { int d = LINKS[link].dof; if (d < 0) return; s->q[d] = angle; s->qd[d] = 0; memcpy(s->qc, s->q, sizeof s->q); }
void po_get_joint(const PoSim *s, int link, double *q, double *qd) { int d = LINKS[link].dof; *q = d < 0 ? 0 : s->q[d]; *qd = d < 0 ? 0 : s->qd[d]; }
void po_control_joint(PoSim *s, int link, double target, double max_force) {
    int d = LINKS[link].dof; if (d < 0) return;
    s->m_kp[d] = 0.1; s->m_kd[d] = 1.0; s->m_tq[d] = target; s->m_tv[d] = 0; s->m_maximp[d] = max_force * DT;
}
void po_set_base_pose(PoSim *s, int o, const double pos[3], const double quat[4]) {
    Obj *b = &s->obj[o]; v3cpy(b->pos, pos); memcpy(b->quat, quat, 4 * sizeof(double));
    double n = sqrt(quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3]);
    for (int k = 0; k < 4; k++) b->quat[k] /= n;
    v3set(b->lin, 0, 0, 0); v3set(b->ang, 0, 0, 0);
}
void po_get_base_pose(const PoSim *s, int o, double pos[3], double quat[4]) { v3cpy(pos, s->obj[o].pos); memcpy(quat, s->obj[o].quat, 4 * sizeof(double)); }
void po_get_base_velocity(const PoSim *s, int o, double lin[3], double ang[3]) { v3cpy(lin, s->obj[o].lin); v3cpy(ang, s->obj[o].ang); }
void po_set_base_velocity(PoSim *s, int o, const double lin[3], const double ang[3]) { v3cpy(s->obj[o].lin, lin); v3cpy(s->obj[o].ang, ang); }
void po_euler_from_quat(const double q[4], double e[3]) { /* SURVEY App. B.4 */
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double sarg = -2 * (x * z - w * y);
    if (sarg <= -0.99999) { e[0] = 0; e[1] = -0.5 * PI; e[2] = 2 * atan2(x, -y); }
    else if (sarg >= 0.99999) { e[0] = 0; e[1] = 0.5 * PI; e[2] = 2 * atan2(-x, y); }
    else {
        e[0] = atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z);
        e[1] = asin(sarg);
        e[2] = atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z);
    }
}
typedef struct { double q[ND], qd[ND], qc[ND], m[5 * ND]; Obj obj[MAXOBJ]; } Snap;
int po_state_size(void) { return (int)((sizeof(Snap) + 7) / 8); }
void po_save_state(const PoSim *s, double *buf) {
    Snap *p = (Snap *)buf; memcpy(p->q, s->q, sizeof p->q); memcpy(p->qd, s->qd, sizeof p->qd); memcpy(p->qc, s->qc, sizeof p->qc);
    memcpy(p->m, s->m_kp, sizeof p->m); memcpy(p->obj, s->obj, sizeof p->obj);
}
void po_restore_state(PoSim *s, const double *buf) {
    const Snap *p = (const Snap *)buf; memcpy(s->q, p->q, sizeof p->q); memcpy(s->qd, p->qd, sizeof p->qd); memcpy(s->qc, p->qc, sizeof p->qc);
    memcpy(s->m_kp, p->m, sizeof p->m); memcpy(s->obj, p->obj, sizeof p->obj);
}
int po_last_num_contacts(const PoSim *s) { return s->last_contacts; }
int po_last_iterations(const PoSim *s) { return s->last_iters; }
/* debug: number of arm joint-limit rows (joints 0..6) that ended the last sub-step with a non-zero impulse */
int po_last_active_arm_limits(const PoSim *s) { int n = 0; for (int r = 0; r < 14; r++) n += s->rows[r].applied > 0; return n; }

/* getLinkState: pose from the cached transforms FK(qc) (App. B.5); velocity = link-local velocity from the fresh
 * (q, qd) rotated to world by the cached basis.  pos = CoM-frame origin (pybullet.py:361 reads index [0]). */
void po_get_link_state(const PoSim *s, int link, double pos[3], double quat[4], double lin[3], double ang[3]) {
    double E[NL][9], r[NL][3], Rw[NL][9], pw[NL][3], v[NL][6];
    fk_all(s, s->q, E, r, Rw, pw);
    for (int i = 0; i <= link; i++) {
        const LinkDef *L = &LINKS[i]; double vp[6] = {0, 0, 0, 0, 0, 0};
        if (L->parent >= 0) xmotion(vp, E[i], r[i], v[L->parent]);
        for (int k = 0; k < 6; k++) v[i][k] = vp[k];
        if (L->jtype == J_REV) v[i][2] += s->qd[L->dof];
        if (L->jtype == J_PRIS) for (int k = 0; k < 3; k++) v[i][3 + k] += L->axis[k] * s->qd[L->dof];
    }
    double vloc[3], t[3];
    v3cross(t, v[link], LINKS[link].com); v3add(vloc, v[link] + 3, t);
    fk_all(s, s->qc, E, r, Rw, pw);
    m3mulv(t, Rw[link], LINKS[link].com); v3add(pos, pw[link], t);
    R_to_quat(quat, Rw[link]);
    m3mulv(lin, Rw[link], vloc); m3mulv(ang, Rw[link], v[link]);
}

/* joint-space inertia matrix through its inverse (debug) */
void po_mass_matrix(PoSim *s, double M[81]) {
    double tau[ND] = {0}, qdd[ND];
    aba(s, tau, qdd);
    for (int j = 0; j < ND; j++) { double e[ND] = {0}, col[ND]; e[j] = 1; impulse_response(s, NULL, e, col); for (int i = 0; i < ND; i++) M[ND * i + j] = col[i]; }
}

/* ------------------------------------------------------------------ inverse kinematics (App. B.2) */
static int gauss_solve(int n, double *A, double *b) { /* in place, partial pivoting; A row-major n x n */
    for (int c = 0; c < n; c++) {
        int p = c; for (int r2 = c + 1; r2 < n; r2++) if (fabs(A[r2 * n + c]) > fabs(A[p * n + c])) p = r2;
        if (fabs(A[p * n + c]) < 1e-300) return -1;
        if (p != c) { for (int k = 0; k < n; k++) { double t = A[c * n + k]; A[c * n + k] = A[p * n + k]; A[p * n + k] = t; } double t = b[c]; b[c] = b[p]; b[p] = t; }
        for (int r2 = c + 1; r2 < n; r2++) { double f = A[r2 * n + c] / A[c * n + c]; for (int k = c; k < n; k++) A[r2 * n + k] -= f * A[c * n + k]; b[r2] -= f * b[c]; }
    }
    for (int c = n - 1; c >= 0; c--) { double a = b[c]; for (int k = c + 1; k < n; k++) a -= A[c * n + k] * b[k]; b[c] = a / A[c * n + c]; }
    return 0;
}
void po_inverse_kinematics(const PoSim *s, int link, const double pos[3], const double quat[4], double out[ND]) {
    double q[ND], tq[4]; memcpy(q, s->q, sizeof q);
    double qn = sqrt(quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3]);
    for (int k = 0; k < 4; k++) tq[k] = quat[k] / qn; /* the target quaternion is normalised first (KAT :265 passes an un-normalised one) */
    double diff = INFINITY;
    for (int it = 0; it < 20 && diff > 1e-4; it++) {
        double E[NL][9], r[NL][3], Rw[NL][9], pw[NL][3];
        fk_all(s, q, E, r, Rw, pw);
        const double *p = pw[link];
        double J[6][ND]; memset(J, 0, sizeof J);
        for (int a = link; a >= 0; a = LINKS[a].parent) {
            const LinkDef *L = &LINKS[a]; if (L->dof < 0) continue;
            double axw[3], d[3], lin[3];
            if (L->jtype == J_REV) { double z[3] = {0, 0, 1}; m3mulv(axw, Rw[a], z); v3sub(d, p, pw[a]); v3cross(lin, axw, d); for (int k = 0; k < 3; k++) { J[k][L->dof] = lin[k]; J[3 + k][L->dof] = axw[k]; } }
            else { m3mulv(axw, Rw[a], L->axis); for (int k = 0; k < 3; k++) J[k][L->dof] = axw[k]; }
        }
        double e[6], dp[3]; v3sub(dp, pos, p); diff = v3norm(dp); v3cpy(e, dp);
        double qc[4], qR[4], dq[4]; R_to_quat(qR, Rw[link]); qc[0] = -qR[0]; qc[1] = -qR[1]; qc[2] = -qR[2]; qc[3] = qR[3];
        quat_mul(dq, tq, qc);
        double w = dq[3] > 1 ? 1 : (dq[3] < -1 ? -1 : dq[3]);
        double angle = 2 * acos(w), s2 = 1 - w * w, axis[3] = {1, 0, 0};
        if (s2 >= 10 * 2.220446049250313e-16) { double si = 1 / sqrt(s2); v3set(axis, dq[0] * si, dq[1] * si, dq[2] * si); }
        if (angle > PI) angle -= 2 * PI;
        double an = v3norm(axis); for (int k = 0; k < 3; k++) e[3 + k] = angle * axis[k] / an;
        double A[ND * ND], b[ND];
        for (int i = 0; i < ND; i++) {
            for (int j = 0; j < ND; j++) { double a = 0; for (int k = 0; k < 6; k++) a += J[k][i] * J[k][j]; A[ND * i + j] = a + (i == j ? 0.5 : 0); }
            double a = 0; for (int k = 0; k < 6; k++) a += J[k][i] * e[k]; b[i] = a;
        }
        gauss_solve(ND, A, b);
        double mx = 0; for (int i = 0; i < ND; i++) if (fabs(b[i]) > mx) mx = fabs(b[i]);
        double lim = 45.0 * PI / 180.0, sc = mx > lim ? lim / mx : 1.0;
        for (int i = 0; i < ND; i++) q[i] += b[i] * sc;
    }
    memcpy(out, q, sizeof q);
}

/* ------------------------------------------------------------------ contact model
 * The reference's collision geometry (convex hulls of the finger/hand meshes, Bullet's SAT / GJK manifolds) is not
 * available here, so this oracle DEFINES the contact model that the CUDA path must reproduce -- parity unpinned:
 *   robot collision boxes: hand (link 8), finger1 (link 9), finger2 (link 10), fixed in their link frames;
 *   objects: box (8 vertices) or z-cylinder (8 rim points, 4 per cap at 45 deg + k 90 deg);
 *   a contact = a vertex of body A against the signed-distance field of body B (plane / box / cylinder) with
 *   distance < CONTACT_MARGIN (4 mm: twice the largest per-sub-step approach, instead of Bullet's 2 cm breaking threshold,
 *   to bound the row count); at most MAXC_* contacts in collection order; pairs: object-table(+ground plane), robot box-table, robot box<->object (both
 *   directions), object<->object (both directions); box<->box object pairs use the other box's REFERENCE FACE (the face axis of minimum overlap)
 *   instead of its distance field for the vertex contacts and add edge-against-edge contacts (see 4a / 4b in collect_contacts).
 * Rows follow btMultiBodyConstraintSolver::setupMultiBodyContactConstraint: speculative when distance > 0
 * (velocityError -= distance/dt), erp 0.2 when penetrating, friction = product of the two coefficients, two friction
 * directions from btPlaneSpace1 with the implicit cone clamp, finger links soft (stiffness 30000, damping 1000 ->
 * contact erp/cfm, App. B.3), no warm starting. */
#define CONTACT_MARGIN 0.004
/* robot box <-> object: the position-controlled fingers close at up to 5 m/s (10 mm per sub-step), so the speculative margin
 * of these pairs must exceed that (Bullet's own margin is its 2 cm contact breaking threshold) */
#define CONTACT_MARGIN_GRASP 0.012   /* applies when the closest feature is a face; edge / corner configurations use CONTACT_MARGIN */
/* contacts per env and sub-step; later candidates are dropped.  Sized to the solver's on-chip contact store. */
#define MAXC_ONE_OBJECT 10   /* no object or one */
#define MAXC_TWO_OBJECTS 22
#define CONTACT_ERP 0.2
#define LINEAR_SLOP 1e-5
#define TABLE_MU 0.5
#define FINGER_MU 1.0
#define HAND_MU 0.5
typedef struct { int link; double c[3], h[3], mu; int soft; } RobotBox;
static const RobotBox RBOX[3] = {
    {8, {0, 0, 0.021}, {0.032, 0.102, 0.045}, HAND_MU, 0},
    {9, {0, 0.0105, 0.027}, {0.0105, 0.0105, 0.027}, FINGER_MU, 1},
    {10, {0, -0.0105, 0.027}, {0.0105, 0.0105, 0.027}, FINGER_MU, 1},
};
void po_get_robot_box(int i, double out[8]) { out[0] = RBOX[i].link; v3cpy(out + 1, RBOX[i].c); v3cpy(out + 4, RBOX[i].h); out[7] = RBOX[i].mu; }

static void box_vertex(const double *h, int k, double *o) { v3set(o, (k & 1) ? h[0] : -h[0], (k & 2) ? h[1] : -h[1], (k & 4) ? h[2] : -h[2]); }
/* edge e (0..11) of a box in the world: axis e / 4, the two other coordinates at -+ half by the bits of e; start point p, direction d (full length) */
#define EDGE_END 0.02           /* closest points within 2 % of an edge's end belong to the vertex contacts */
#define EDGE_MIN_SIN2 0.01      /* sin^2 of the smallest angle between two edges that still defines a common perpendicular (~5.7 degrees) */
#define EDGE_MAX_DEPTH 0.01     /* deeper than this the pair is not a contact of touching surfaces */
static void box_edge(const Obj *b, const double *R, int e, double *p, double *d) {
    int a = e >> 2, b1 = (a + 1) % 3, c1 = (a + 2) % 3;
    double l[3], dl[3] = {0, 0, 0};
    l[a] = -b->half[a]; l[b1] = (e & 1) ? b->half[b1] : -b->half[b1]; l[c1] = (e & 2) ? b->half[c1] : -b->half[c1];
    dl[a] = 2 * b->half[a];
    m3mulv(p, R, l); v3add(p, p, b->pos); m3mulv(d, R, dl);
}
static void obj_vertex(const Obj *b, int k, double *o) {
    if (b->shape == SH_BOX) box_vertex(b->half, k, o);
    else { double ang = PI / 4 + (k & 3) * (PI / 2); v3set(o, b->half[0] * cos(ang), b->half[0] * sin(ang), (k & 4) ? b->half[2] : -b->half[2]); }
}
/* signed distance of local point p to a box of half extents h; n = outward normal (local) */
static int g_face; /* set by the sdf functions: 1 when the closest feature is a face (or the point is inside) */
static double sdf_box(const double *h, const double *p, double *n) {
    g_face = 1;
    double d[3] = {fabs(p[0]) - h[0], fabs(p[1]) - h[1], fabs(p[2]) - h[2]};
    if (d[0] <= 0 && d[1] <= 0 && d[2] <= 0) { /* inside: nearest face */
        int a = 0; if (d[1] > d[a]) a = 1; if (d[2] > d[a]) a = 2;
        v3set(n, 0, 0, 0); n[a] = p[a] >= 0 ? 1 : -1; return d[a];
    }
    double o[3] = {d[0] > 0 ? d[0] : 0, d[1] > 0 ? d[1] : 0, d[2] > 0 ? d[2] : 0};
    g_face = ((d[0] > 0) + (d[1] > 0) + (d[2] > 0)) == 1;
    double len = v3norm(o);
    for (int k = 0; k < 3; k++) n[k] = (p[k] >= 0 ? o[k] : -o[k]) / len;
    return len;
}
/* signed distance to a z-cylinder (radius r, half height hz) */
static double sdf_cyl(double r, double hz, const double *p, double *n) {
    double rho = sqrt(p[0] * p[0] + p[1] * p[1]);
    double dr = rho - r, dz = fabs(p[2]) - hz;
    double rx = rho > 1e-12 ? p[0] / rho : 1, ry = rho > 1e-12 ? p[1] / rho : 0, sz = p[2] >= 0 ? 1 : -1;
    g_face = !(dr > 0 && dz > 0);
    if (dr <= 0 && dz <= 0) { if (dr > dz) { v3set(n, rx, ry, 0); return dr; } v3set(n, 0, 0, sz); return dz; }
    double a = dr > 0 ? dr : 0, b = dz > 0 ? dz : 0, len = sqrt(a * a + b * b);
    v3set(n, rx * a / len, ry * a / len, sz * b / len);
    return len;
}
static double obj_sdf(const Obj *b, const double *p, double *n) { return b->shape == SH_BOX ? sdf_box(b->half, p, n) : sdf_cyl(b->half[0], b->half[2], p, n); }

/* generalized Jacobian row of a unit force `dir` (world) applied at world point P on: robot link `link` (>=0) or
 * object `obj` (>=0); sign multiplies. */
static void add_point_jac(const PoSim *s, double *J, int link, int obj, const double *P, const double *dir, double sign) {
    if (link >= 0) {
        for (int a = link; a >= 0; a = LINKS[a].parent) {
            const LinkDef *L = &LINKS[a]; if (L->dof < 0) continue;
            double axw[3], d[3], t[3];
            if (L->jtype == J_REV) { double z[3] = {0, 0, 1}; m3mulv(axw, s->Rw[a], z); v3sub(d, P, s->pw[a]); v3cross(t, axw, d); J[L->dof] += sign * v3dot(t, dir); }
            else { m3mulv(axw, s->Rw[a], L->axis); J[L->dof] += sign * v3dot(axw, dir); }
        }
    } else if (obj >= 0) {
        double d[3], t[3]; v3sub(d, P, s->obj[obj].pos); v3cross(t, d, dir);
        for (int k = 0; k < 3; k++) { J[ND + 6 * obj + k] += sign * dir[k]; J[ND + 6 * obj + 3 + k] += sign * t[k]; }
    }
}
static void apply_minv(const PoSim *s, const double *J, double *W) {
    for (int i = 0; i < ND; i++) { double a = 0; for (int j = 0; j < ND; j++) a += s->Minv[i][j] * J[j]; W[i] = a; }
    for (int o = 0; o < MAXOBJ; o++) {
        double *Wo = W + ND + 6 * o; const double *Jo = J + ND + 6 * o;
        if (o >= s->nobj) { memset(Wo, 0, 6 * sizeof(double)); continue; }
        const Obj *b = &s->obj[o]; double R[9], t[3];
        quat_to_R(R, b->quat);
        v3scale(Wo, Jo, 1.0 / b->mass);
        m3Tmulv(t, R, Jo + 3); for (int k = 0; k < 3; k++) t[k] /= b->Ic[k]; m3mulv(Wo + 3, R, t);
    }
}
static void plane_space(const double *n, double *p, double *q) { /* btPlaneSpace1 */
    if (fabs(n[2]) > 0.7071067811865475244) {
        double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
        v3set(p, 0, -n[2] * k, n[1] * k); v3set(q, a * k, -n[0] * p[2], n[0] * p[1]);
    } else {
        double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
        v3set(p, -n[1] * k, n[0] * k, 0); v3set(q, -n[2] * p[1], n[2] * p[0], a * k);
    }
}
/* one contact: point P (world), normal n (world, pointing from B to A), distance; A/B are (link,obj) with -1/-1 = static */
static void add_contact(PoSim *s, const double *gv, const double *P, const double *n, double dist, int linkA, int objA, int linkB, int objB, double mu, int soft) {
    int on_robot = linkA >= 0 || linkB >= 0;
    if (s->last_contacts >= (s->nobj <= 1 ? MAXC_ONE_OBJECT : MAXC_TWO_OBJECTS)) return;
    if (on_robot) s->last_robot_contacts++;
    double t1[3], t2[3]; plane_space(n, t1, t2);
    const double *dirs[3] = {n, t1, t2};
    int nrow = s->nrows;
    { double *g = s->dbg[s->last_contacts]; v3cpy(g, P); v3cpy(g + 3, n); g[6] = dist; g[7] = linkA >= 0 ? linkA : (linkB >= 0 ? -linkB - 100 : -1); g[8] = objA >= 0 ? objA : (objB >= 0 ? -objB - 100 : -1); g[9] = nrow; }
    double erp = CONTACT_ERP, cfm = 0;
    if (soft) { double k = 30000.0, d = 1000.0; erp = DT * k / (DT * k + d); cfm = 1.0 / (DT * k + d) / DT; }
    for (int k = 0; k < 3; k++) {
        Row *r = &s->rows[s->nrows++]; memset(r, 0, sizeof *r);
        add_point_jac(s, r->J, linkA, objA, P, dirs[k], 1.0);
        add_point_jac(s, r->J, linkB, objB, P, dirs[k], -1.0);
        apply_minv(s, r->J, r->W);
        double den = 0, rel = 0; for (int i = 0; i < NDT; i++) { den += r->J[i] * r->W[i]; rel += r->J[i] * gv[i]; }
        if (k == 0) {
            r->invD = 1.0 / (den + cfm); r->cfm = cfm * r->invD;   /* solverConstraint.m_cfm = cfm * m_jacDiagABInv */
            double pen = dist + LINEAR_SLOP, poserr = 0, velerr = -rel;
            if (pen > 0) velerr -= pen / DT; else poserr = -pen * erp / DT;
            r->rhs = (poserr + velerr) * r->invD; r->lo = 0; r->hi = 1e10; r->normal_row = -1;
        } else {
            r->invD = 1.0 / den; r->rhs = -rel * r->invD; r->mu = mu; r->normal_row = nrow; r->lo = 0; r->hi = 0;
        }
    }
    s->last_contacts++;
}
static int over_table(const PoSim *s, const double *p) { return p[0] >= s->table_x0 && p[0] <= s->table_x1 && p[1] >= s->table_y0 && p[1] <= s->table_y1; }
static void collect_contacts(PoSim *s, const double *gv) {
    double up[3] = {0, 0, 1};
    /* robot boxes: world pose from the sub-step's FK */
    double Rb[3][9], cb[3][3];
    for (int b = 0; b < 3; b++) { int l = RBOX[b].link; double t[3]; memcpy(Rb[b], s->Rw[l], sizeof Rb[b]); m3mulv(t, s->Rw[l], RBOX[b].c); v3add(cb[b], s->pw[l], t); }
    /* 1. object vertices vs table top / ground plane */
    for (int o = 0; o < s->nobj; o++) {
        const Obj *ob = &s->obj[o]; double R[9]; quat_to_R(R, ob->quat);
        for (int k = 0; k < 8; k++) {
            double v[3], P[3]; obj_vertex(ob, k, v); m3mulv(P, R, v); v3add(P, P, ob->pos);
            double plane = over_table(s, P) && P[2] > -0.05 ? 0.0 : s->ground_z;
            double d = P[2] - plane;
            if (d < CONTACT_MARGIN) add_contact(s, gv, P, up, d, -1, o, -1, -1, ob->mu * TABLE_MU, 0);
        }
    }
    /* 2. robot box vertices vs table top */
    /* fingers: only the 4 vertices of the outer face (the side away from the other finger).  Against a plane the lowest point of
     * the finger pair is always attained there: the inner-face vertices lie between outer vertices of the two fingers. */
    for (int b = 0; b < 3; b++) for (int k = 0; k < 8; k++) {
        if ((b == 1 && !(k & 2)) || (b == 2 && (k & 2))) continue;
        double v[3], P[3]; box_vertex(RBOX[b].h, k, v); m3mulv(P, Rb[b], v); v3add(P, P, cb[b]);
        if (over_table(s, P) && P[2] < CONTACT_MARGIN) add_contact(s, gv, P, up, P[2], RBOX[b].link, -1, -1, -1, RBOX[b].mu * TABLE_MU, RBOX[b].soft);
    }
    /* 3. robot box <-> object */
    for (int b = 0; b < 3; b++) for (int o = 0; o < s->nobj; o++) {
        const Obj *ob = &s->obj[o]; double R[9], dc[3]; quat_to_R(R, ob->quat);
        v3sub(dc, cb[b], ob->pos);
        if (v3norm(dc) > v3norm(RBOX[b].h) + v3norm(ob->half) + CONTACT_MARGIN_GRASP) continue;
        double mu = RBOX[b].mu * ob->mu;
        for (int k = 0; k < 8; k++) { /* robot box vertex in the object's field: A = robot, normal = object's outward */
            double v[3], P[3], pl[3], nl[3], nw[3], t[3]; box_vertex(RBOX[b].h, k, v); m3mulv(P, Rb[b], v); v3add(P, P, cb[b]);
            v3sub(t, P, ob->pos); m3Tmulv(pl, R, t);
            double d = obj_sdf(ob, pl, nl);
            if (d < CONTACT_MARGIN || (d < CONTACT_MARGIN_GRASP && g_face)) { m3mulv(nw, R, nl); add_contact(s, gv, P, nw, d, RBOX[b].link, -1, -1, o, mu, RBOX[b].soft); }
        }
        for (int k = 0; k < 8; k++) { /* object vertex in the robot box's field: A = object */
            double v[3], P[3], pl[3], nl[3], nw[3], t[3]; obj_vertex(ob, k, v); m3mulv(P, R, v); v3add(P, P, ob->pos);
            v3sub(t, P, cb[b]); m3Tmulv(pl, Rb[b], t);
            double d = sdf_box(RBOX[b].h, pl, nl);
            if (d < CONTACT_MARGIN || (d < CONTACT_MARGIN_GRASP && g_face)) { m3mulv(nw, Rb[b], nl); add_contact(s, gv, P, nw, d, -1, o, RBOX[b].link, -1, mu, RBOX[b].soft); }
        }
    }
    /* 4. object <-> object */
    if (s->nobj == 2) {
        double dc[3]; v3sub(dc, s->obj[0].pos, s->obj[1].pos);
        if (v3norm(dc) <= v3norm(s->obj[0].half) + v3norm(s->obj[1].half) + CONTACT_MARGIN)
            for (int a = 0; a < 2; a++) {
                const Obj *oa = &s->obj[a], *ob = &s->obj[1 - a]; double Ra[9], Rbm[9];
                quat_to_R(Ra, oa->quat); quat_to_R(Rbm, ob->quat);
                if (oa->shape == SH_BOX && ob->shape == SH_BOX) {
                    /* 4a. box <-> box, vertices of A against B's REFERENCE FACE: the face axis of B along which the two boxes overlap least
                     * (the separating-axis choice of Bullet's box-box detector restricted to B's face normals).  A vertex within the margin of
                     * that face's plane and inside the face's rectangle grown by the margin is a contact along the face normal.  The
                     * signed-distance field of B alone cannot make this choice: a vertex of a cube stacked on a cube of the same size sits at
                     * a corner of the other's face, where the nearest face is a side face as often as the top. */
                    double cl[3], t[3], depth[3]; v3sub(t, oa->pos, ob->pos); m3Tmulv(cl, Rbm, t);
                    for (int k = 0; k < 3; k++) {
                        double proj = 0;
                        for (int j = 0; j < 3; j++) proj += fabs(Rbm[0 * 3 + k] * Ra[0 * 3 + j] + Rbm[1 * 3 + k] * Ra[1 * 3 + j] + Rbm[2 * 3 + k] * Ra[2 * 3 + j]) * oa->half[j];
                        depth[k] = ob->half[k] + proj - fabs(cl[k]);
                    }
                    int km = 0; if (depth[1] < depth[km]) km = 1; if (depth[2] < depth[km]) km = 2;
                    if (depth[km] < -CONTACT_MARGIN) continue;          /* separated along that axis */
                    double sg = cl[km] >= 0 ? 1.0 : -1.0, nw[3] = {sg * Rbm[0 * 3 + km], sg * Rbm[1 * 3 + km], sg * Rbm[2 * 3 + km]};
                    int i1 = (km + 1) % 3, i2 = (km + 2) % 3;
                    for (int k = 0; k < 8; k++) {
                        double v[3], P[3], pl[3]; box_vertex(oa->half, k, v); m3mulv(P, Ra, v); v3add(P, P, oa->pos);
                        v3sub(t, P, ob->pos); m3Tmulv(pl, Rbm, t);
                        double d = sg * pl[km] - ob->half[km];
                        if (d < CONTACT_MARGIN && d > -EDGE_MAX_DEPTH && fabs(pl[i1]) <= ob->half[i1] + CONTACT_MARGIN && fabs(pl[i2]) <= ob->half[i2] + CONTACT_MARGIN)
                            add_contact(s, gv, P, nw, d, -1, a, -1, 1 - a, oa->mu * ob->mu, 0);
                    }
                    continue;
                }
                for (int k = 0; k < 8; k++) {
                    double v[3], P[3], pl[3], nl[3], nw[3], t[3]; obj_vertex(oa, k, v); m3mulv(P, Ra, v); v3add(P, P, oa->pos);
                    v3sub(t, P, ob->pos); m3Tmulv(pl, Rbm, t);
                    double d = obj_sdf(ob, pl, nl);
                    if (d < CONTACT_MARGIN) { m3mulv(nw, Rbm, nl); add_contact(s, gv, P, nw, d, -1, a, -1, 1 - a, oa->mu * ob->mu, 0); }
                }
            }
        /* 4b. box <-> box, edge against edge.  Vertex-in-field contacts miss two boxes whose faces overlap with every vertex of either
         * outside the other's face -- a cube lying rotated on a cube of the same size (reference tasks/stack.py:30-62: two 4 cm cubes) has no
         * vertex contact at all beyond ~11 degrees.  Bullet's box-box manifold covers that case with face clipping; here: every pair of
         * edges whose mutual closest points are interior to both edges (so no vertex is involved) and closer than the margin.  Normal =
         * the edges' common perpendicular, pointing from object 1 to object 0. */
#ifndef PO_NO_EDGE_CONTACTS      /* (build switch of tests/test_contact_kat_oracle.py: without 4b the turned cube sinks into the lower one) */
        if (v3norm(dc) <= v3norm(s->obj[0].half) + v3norm(s->obj[1].half) + CONTACT_MARGIN && s->obj[0].shape == SH_BOX && s->obj[1].shape == SH_BOX) {
            double R0[9], R1[9]; quat_to_R(R0, s->obj[0].quat); quat_to_R(R1, s->obj[1].quat);
            for (int ea = 0; ea < 12; ea++) {
                double p1[3], d1[3]; box_edge(&s->obj[0], R0, ea, p1, d1);
                for (int eb = 0; eb < 12; eb++) {
                    double p2[3], d2[3], r[3]; box_edge(&s->obj[1], R1, eb, p2, d2);
                    v3sub(r, p1, p2);
                    double A = v3dot(d1, d1), E = v3dot(d2, d2), B = v3dot(d1, d2), C = v3dot(d1, r), F = v3dot(d2, r), den = A * E - B * B;
                    if (den <= EDGE_MIN_SIN2 * A * E) continue;                                   /* (nearly) parallel edges */
                    double sa = (B * F - C * E) / den, tb = (A * F - B * C) / den;
                    if (sa <= EDGE_END || sa >= 1 - EDGE_END || tb <= EDGE_END || tb >= 1 - EDGE_END) continue;   /* an end point is closest: a vertex contact's business */
                    double n[3], pa[3], pb[3], P[3], diff[3];
                    v3cross(n, d1, d2); { double k = 1.0 / sqrt(den); n[0] *= k; n[1] *= k; n[2] *= k; }
                    if (v3dot(n, dc) < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
                    for (int i = 0; i < 3; i++) { pa[i] = p1[i] + sa * d1[i]; pb[i] = p2[i] + tb * d2[i]; P[i] = 0.5 * (pa[i] + pb[i]); }
                    v3sub(diff, pa, pb);
                    double d = v3dot(diff, n);
                    if (d < CONTACT_MARGIN && d > -EDGE_MAX_DEPTH) add_contact(s, gv, P, n, d, -1, 0, -1, 1, s->obj[0].mu * s->obj[1].mu, 0);
                }
            }
        }
#endif
    }
}

/* ------------------------------------------------------------------ one 2 ms sub-step (App. B.3) */
static double solve_row(Row *r, double *dv, double lo, double hi) {
    double jd = 0; for (int i = 0; i < NDT; i++) jd += r->J[i] * dv[i];
    double di = r->rhs - r->applied * r->cfm - jd * r->invD;
    double sum = r->applied + di;
    if (sum < lo) { di = lo - r->applied; r->applied = lo; } else if (sum > hi) { di = hi - r->applied; r->applied = hi; } else r->applied = sum;
    for (int i = 0; i < NDT; i++) dv[i] += r->W[i] * di;
    return di / r->invD;
}
static void substep(PoSim *s) {
    double tau[ND] = {0}, qdd[ND], gv[NDT];
    memcpy(s->qc, s->q, sizeof s->q);            /* forwardKinematics(): refresh of the link-transform cache */
    aba(s, tau, qdd);
    for (int j = 0; j < ND; j++) { double e[ND] = {0}, col[ND]; e[j] = 1; impulse_response(s, NULL, e, col); for (int i = 0; i < ND; i++) s->Minv[i][j] = col[i]; }
    for (int d = 0; d < ND; d++) s->qd[d] += qdd[d] * DT;
    for (int o = 0; o < s->nobj; o++) { /* 0-link floating base: gravity, damping, gyroscopic term */
        Obj *b = &s->obj[o]; double R[9], wl[3], Iw[3], g[3], al[3], aw[3];
        double kl = LINK_DAMP + LINK_DAMP * v3norm(b->lin), ka = LINK_DAMP + LINK_DAMP * v3norm(b->ang);
        quat_to_R(R, b->quat); m3Tmulv(wl, R, b->ang);
        for (int k = 0; k < 3; k++) Iw[k] = b->Ic[k] * wl[k];
        v3cross(g, wl, Iw);
        for (int k = 0; k < 3; k++) al[k] = -(g[k] + Iw[k] * ka) / b->Ic[k];
        m3mulv(aw, R, al);
        for (int k = 0; k < 3; k++) { b->ang[k] += aw[k] * DT; b->lin[k] += (-b->lin[k] * kl + (k == 2 ? -GRAV : 0)) * DT; }
    }
    memset(gv, 0, sizeof gv);
    memcpy(gv, s->qd, sizeof s->qd);
    for (int o = 0; o < s->nobj; o++) { v3cpy(gv + ND + 6 * o, s->obj[o].lin); v3cpy(gv + ND + 6 * o + 3, s->obj[o].ang); }

    /* rows: joint limits (link order, lower then upper), motors, then contacts */
    s->nrows = 0; s->last_contacts = 0; s->last_robot_contacts = 0;
    for (int d = 0; d < ND; d++) for (int side = 0; side < 2; side++) {
        const LinkDef *L = &LINKS[DOF_LINK[d]]; Row *r = &s->rows[s->nrows++]; memset(r, 0, sizeof *r);
        double sg = side == 0 ? 1.0 : -1.0, pen = side == 0 ? s->q[d] - L->lo : L->hi - s->q[d];
        r->J[d] = sg; apply_minv(s, r->J, r->W); r->invD = 1.0 / (r->J[d] * r->W[d]);
        double rel = sg * s->qd[d], poserr = 0, velerr = -rel;
        if (pen > 0) velerr = -pen / DT; else poserr = -pen * CONTACT_ERP / DT;
        r->rhs = (poserr + velerr) * r->invD; r->lo = 0; r->hi = 100.0; r->normal_row = -1;
    }
    for (int d = 0; d < ND; d++) {
        Row *r = &s->rows[s->nrows++]; memset(r, 0, sizeof *r);
        r->J[d] = 1; apply_minv(s, r->J, r->W); r->invD = 1.0 / r->W[d];
        double target = s->m_kp[d] * (s->m_tq[d] - s->q[d]) / DT + s->qd[d] + s->m_kd[d] * (s->m_tv[d] - s->qd[d]);
        r->rhs = (target - s->qd[d]) * r->invD; r->lo = -s->m_maximp[d]; r->hi = s->m_maximp[d]; r->normal_row = -1;
    }
    int n_noncontact = s->nrows;
    collect_contacts(s, gv);

    /* sequential impulse, <= 50 iterations, early exit on max squared velocity residual <= 1e-7 */
    double dv[NDT]; memset(dv, 0, sizeof dv);
    int it;
    for (it = 0; it < 50; it++) {
        double res = 0;
        for (int j = 0; j < n_noncontact; j++) {
            int idx = (it & 1) ? j : n_noncontact - 1 - j;
            double dvl = solve_row(&s->rows[idx], dv, s->rows[idx].lo, s->rows[idx].hi); if (dvl * dvl > res) res = dvl * dvl;
        }
        for (int j = n_noncontact; j < s->nrows; j += 3) { double dvl = solve_row(&s->rows[j], dv, 0, 1e10); if (dvl * dvl > res) res = dvl * dvl; }
        for (int j = n_noncontact; j < s->nrows; j += 3) { /* implicit cone over the two friction rows */
            Row *rn = &s->rows[j], *r1 = &s->rows[j + 1], *r2 = &s->rows[j + 2];
            if (rn->applied <= 0) continue;
            double jd1 = 0, jd2 = 0; for (int i = 0; i < NDT; i++) { jd1 += r1->J[i] * dv[i]; jd2 += r2->J[i] * dv[i]; }
            double d1 = r1->rhs - jd1 * r1->invD, d2 = r2->rhs - jd2 * r2->invD;
            double s1 = r1->applied + d1, s2 = r2->applied + d2, lim = r1->mu * rn->applied, len = sqrt(s1 * s1 + s2 * s2);
            if (len > lim) { s1 *= lim / len; s2 *= lim / len; }
            d1 = s1 - r1->applied; d2 = s2 - r2->applied; r1->applied = s1; r2->applied = s2;
            for (int i = 0; i < NDT; i++) dv[i] += r1->W[i] * d1 + r2->W[i] * d2;
            double a = d1 / r1->invD, b = d2 / r2->invD; if (a * a > res) res = a * a; if (b * b > res) res = b * b;
        }
        if (res <= 1e-7) { it++; break; }
    }
    s->last_iters = it;
    for (int d = 0; d < ND; d++) s->qd[d] += dv[d];
    for (int o = 0; o < s->nobj; o++) for (int k = 0; k < 3; k++) { s->obj[o].lin[k] += dv[ND + 6 * o + k]; s->obj[o].ang[k] += dv[ND + 6 * o + 3 + k]; }
    /* stepPositionsMultiDof */
    for (int d = 0; d < ND; d++) s->q[d] += s->qd[d] * DT;
    for (int o = 0; o < s->nobj; o++) {
        Obj *b = &s->obj[o];
        for (int k = 0; k < 3; k++) b->pos[k] += b->lin[k] * DT;
        double fa = v3norm(b->ang), ax[3], dq[4], nq[4];
        if (fa * DT > 0.25 * PI) fa = 0.25 * PI / DT; /* ANGULAR_MOTION_THRESHOLD */
        if (fa < 0.001) v3scale(ax, b->ang, 0.5 * DT - DT * DT * DT * 0.020833333333 * fa * fa); else v3scale(ax, b->ang, sin(0.5 * fa * DT) / fa);
        dq[0] = ax[0]; dq[1] = ax[1]; dq[2] = ax[2]; dq[3] = cos(0.5 * fa * DT);
        quat_mul(nq, dq, b->quat);
        double n = sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
        for (int k = 0; k < 4; k++) b->quat[k] = nq[k] / n;
    }
}
/* debug: contact i of the last sub-step -> P(3) n(3) dist robot-link object normal-impulse friction-impulses(2) */
int po_get_contact(const PoSim *s, int i, double out[12]) {
    if (i >= s->last_contacts) return -1;
    memcpy(out, s->dbg[i], 9 * sizeof(double)); int r = (int)s->dbg[i][9];
    out[9] = s->rows[r].applied; out[10] = s->rows[r + 1].applied; out[11] = s->rows[r + 2].applied; return 0;
}
void po_step(PoSim *s, int n_substeps) { for (int i = 0; i < n_substeps; i++) substep(s); }

/* ------------------------------------------------------------------ rewards (utils.py:4-30, tasks/) */
static int goal_dim(int task) { return task == PO_STACK ? 6 : (task == PO_FLIP ? 4 : 3); }
static float thr_f32(int task) { return task == PO_STACK ? 0.1f : (task == PO_FLIP ? 0.2f : 0.05f); }
static double thr_f64(int task) { return task == PO_STACK ? 0.1 : (task == PO_FLIP ? 0.2 : 0.05); }
static float dist_f32(int task, const float *a, const float *b) {
    int g = goal_dim(task);
    if (task == PO_FLIP) { /* row-wise 1 - <a,b>^2; the pairwise order matches numpy's 4-element float32 dot (tests/golden) */
        float s = (a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3]);
        return 1.0f - s * s;
    }
    float acc = 0; for (int k = 0; k < g; k++) { float d = a[k] - b[k]; float sq = d * d; acc = k == 0 ? sq : acc + sq; }
    return sqrtf(acc);
}
static double dist_f64(int task, const double *a, const double *b) {
    int g = goal_dim(task);
    if (task == PO_FLIP) { double s = (a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3]); return 1.0 - s * s; }
    double acc = 0; for (int k = 0; k < g; k++) { double d = a[k] - b[k]; double sq = d * d; acc = k == 0 ? sq : acc + sq; }
    return sqrt(acc);
}
void po_compute_reward_f32(int task, int rt, const float *ag, const float *dg, float *out, long n) {
    int g = goal_dim(task); float thr = thr_f32(task);
    for (long i = 0; i < n; i++) { float d = dist_f32(task, ag + g * i, dg + g * i); out[i] = rt == PO_REWARD_SPARSE ? -(d > thr ? 1.0f : 0.0f) : -d; }
}
void po_is_success_f32(int task, const float *ag, const float *dg, unsigned char *out, long n) {
    int g = goal_dim(task); float thr = thr_f32(task);
    for (long i = 0; i < n; i++) out[i] = dist_f32(task, ag + g * i, dg + g * i) < thr;
}
void po_compute_reward_f64(int task, int rt, const double *ag, const double *dg, float *out, long n) {
    int g = goal_dim(task); double thr = thr_f64(task);
    for (long i = 0; i < n; i++) { double d = dist_f64(task, ag + g * i, dg + g * i); out[i] = rt == PO_REWARD_SPARSE ? -(d > thr ? 1.0f : 0.0f) : -(float)d; }
}
void po_is_success_f64(int task, const double *ag, const double *dg, unsigned char *out, long n) {
    int g = goal_dim(task); double thr = thr_f64(task);
    for (long i = 0; i < n; i++) out[i] = dist_f64(task, ag + g * i, dg + g * i) < thr;
}

/* ------------------------------------------------------------------ env (core.py:229-289, panda.py, tasks/) */
struct PoEnv { PoSim *sim; int task, control, reward, block_gripper, nsub; double goal[6]; double thr; };
static const double NEUTRAL[ND] = {0.00, 0.41, 0.00, -1.85, 0.00, 2.26, 0.79, 0.00, 0.00};
static const double FORCES[ND] = {87.0, 87.0, 87.0, 87.0, 12.0, 120.0, 120.0, 170.0, 170.0};
PoEnv *po_env_create(int task, int control, int reward) {
    PoEnv *e = (PoEnv *)calloc(1, sizeof(PoEnv));
    e->sim = po_create(task, -0.6, 0.0, 0.0); e->task = task; e->control = control; e->reward = reward;
    e->block_gripper = (task == PO_REACH || task == PO_PUSH || task == PO_SLIDE); /* panda_tasks.py:60,77,94 */
    e->nsub = 20; e->thr = thr_f64(task);
    return e;
}
void po_env_destroy(PoEnv *e) { po_destroy(e->sim); free(e); }
/* PyBullet(n_substeps) (pybullet.py:26) and the task's distance_threshold (tasks/reach.py:15 ...) */
void po_env_set_params(PoEnv *e, int n_substeps, double distance_threshold) { e->nsub = n_substeps; e->thr = distance_threshold; }
PoSim *po_env_sim(PoEnv *e) { return e->sim; }
int po_env_goal_dim(const PoEnv *e) { return goal_dim(e->task); }
int po_env_action_dim(const PoEnv *e) { return (e->control == PO_CTRL_EE ? 3 : 7) + (e->block_gripper ? 0 : 1); }
int po_env_obs_dim(const PoEnv *e) {
    int r = e->block_gripper ? 6 : 7;
    switch (e->task) { case PO_REACH: return r; case PO_STACK: return r + 24; case PO_FLIP: return r + 13; default: return r + 12; }
}
static void env_obs(PoEnv *e, float *obs, float *ag, float *dg) {
    PoSim *s = e->sim; double p[3], qt[4], lin[3], ang[3]; int n = 0;
    po_get_link_state(s, 11, p, qt, lin, ang);
    for (int k = 0; k < 3; k++) obs[n++] = (float)p[k];
    for (int k = 0; k < 3; k++) obs[n++] = (float)lin[k];
    if (!e->block_gripper) obs[n++] = (float)(s->q[7] + s->q[8]);
    for (int o = 0; o < s->nobj; o++) {
        const Obj *b = &s->obj[o]; double eu[3];
        for (int k = 0; k < 3; k++) obs[n++] = (float)b->pos[k];
        if (e->task == PO_FLIP) for (int k = 0; k < 4; k++) obs[n++] = (float)b->quat[k];
        else { po_euler_from_quat(b->quat, eu); for (int k = 0; k < 3; k++) obs[n++] = (float)eu[k]; }
        for (int k = 0; k < 3; k++) obs[n++] = (float)b->lin[k];
        for (int k = 0; k < 3; k++) obs[n++] = (float)b->ang[k];
    }
    int g = goal_dim(e->task);
    if (e->task == PO_REACH) for (int k = 0; k < 3; k++) ag[k] = (float)p[k];
    else if (e->task == PO_FLIP) for (int k = 0; k < 4; k++) ag[k] = (float)s->obj[0].quat[k];
    else for (int o = 0; o < s->nobj; o++) for (int k = 0; k < 3; k++) ag[3 * o + k] = (float)s->obj[o].pos[k];
    for (int k = 0; k < g; k++) dg[k] = (float)e->goal[k];
}
void po_env_reset(PoEnv *e, const double *goal, const double *objpos, float *obs, float *ag, float *dg) {
    PoSim *s = e->sim; double ident[4] = {0, 0, 0, 1};
    for (int d = 0; d < ND; d++) { s->q[d] = NEUTRAL[d]; s->qd[d] = 0; s->qc[d] = NEUTRAL[d]; }
    memcpy(e->goal, goal, goal_dim(e->task) * sizeof(double));
    for (int o = 0; o < s->nobj; o++) po_set_base_pose(s, o, objpos + 3 * o, ident);
    env_obs(e, obs, ag, dg);
}
void po_env_set_state(PoEnv *e, const double *q, const double *qd) { PoSim *s = e->sim; for (int d = 0; d < ND; d++) { s->q[d] = q[d]; s->qd[d] = qd[d]; s->qc[d] = q[d] - qd[d] * DT; } }
void po_env_get_state(PoEnv *e, double *q, double *qd) { memcpy(q, e->sim->q, sizeof e->sim->q); memcpy(qd, e->sim->qd, sizeof e->sim->qd); }
void po_env_step(PoEnv *e, const float *action, float *obs, float *ag, float *dg, float *reward, unsigned char *terminated) {
    po_env_step_oriented(e, action, NULL, 0.05, 0.2, obs, ag, dg, reward, terminated);
}
/* the fork's robots/panda_ori.py:52-99 (EE target orientation) and panda_cartesian.py:67,157 (unscaled actions) */
void po_env_step_oriented(PoEnv *e, const float *action, const double *target_quat, double ee_scale, double finger_scale, float *obs, float *ag, float *dg, float *reward, unsigned char *terminated) {
    PoSim *s = e->sim; double a[8], target[ND]; int na = po_env_action_dim(e);
    for (int k = 0; k < na; k++) { double v = action[k]; a[k] = v < -1 ? -1 : (v > 1 ? 1 : v); }
    if (e->control == PO_CTRL_EE) {
        double p[3], qt[4], lin[3], ang[3], tq[4] = {1, 0, 0, 0}, ik[ND];
        if (target_quat) memcpy(tq, target_quat, sizeof tq);
        po_get_link_state(s, 11, p, qt, lin, ang);
        for (int k = 0; k < 3; k++) p[k] += a[k] * ee_scale;
        if (p[2] < 0) p[2] = 0;
        po_inverse_kinematics(s, 11, p, tq, ik);
        for (int d = 0; d < 7; d++) target[d] = ik[d];
    } else for (int d = 0; d < 7; d++) target[d] = s->q[d] + a[d] * ee_scale;
    double w = e->block_gripper ? 0.0 : (s->q[7] + s->q[8]) + a[na - 1] * finger_scale;
    target[7] = target[8] = w / 2;
    for (int d = 0; d < ND; d++) po_control_joint(s, DOF_LINK[d], target[d], FORCES[d]);
    po_step(s, e->nsub);
    env_obs(e, obs, ag, dg);
    /* core.py:285-288: is_success / compute_reward on (float32 achieved goal, float64 task goal) -> numpy promotes to float64 */
    { double a64[6]; for (int k = 0; k < goal_dim(e->task); k++) a64[k] = (double)ag[k];
      double d = dist_f64(e->task, a64, e->goal); *terminated = d < e->thr; *reward = e->reward == PO_REWARD_SPARSE ? -(d > e->thr ? 1.0f : 0.0f) : -(float)d; }
}

/* batched forms for the test-suite's threaded / multi-process drivers: envs[i] steps with actions[i] */
void po_env_step_batch(PoEnv **envs, int n, int na, int no, int ng, const float *actions, float *obs, float *ag, float *dg, float *reward, unsigned char *terminated) {
    for (int i = 0; i < n; i++) po_env_step(envs[i], actions + (size_t)i * na, obs + (size_t)i * no, ag + (size_t)i * ng, dg + (size_t)i * ng, reward + i, terminated + i);
}
/* full state of an env: q[9] qd[9] | per object pos3 quat4 lin3 ang3 | goal[G]  (the layout of pg_get_state without the step counter) */
void po_env_set_full_state(PoEnv *e, const double *st) {
    PoSim *s = e->sim; int g = goal_dim(e->task);
    po_env_set_state(e, st, st + 9);
    for (int o = 0; o < s->nobj; o++) { const double *p = st + 18 + 13 * o; po_set_base_pose(s, o, p, p + 3); po_set_base_velocity(s, o, p + 7, p + 10); }
    memcpy(e->goal, st + 18 + 13 * s->nobj, g * sizeof(double));
}
void po_env_get_full_state(PoEnv *e, double *st) {
    PoSim *s = e->sim; int g = goal_dim(e->task);
    memcpy(st, s->q, sizeof s->q); memcpy(st + 9, s->qd, sizeof s->qd);
    for (int o = 0; o < s->nobj; o++) { double *p = st + 18 + 13 * o; po_get_base_pose(s, o, p, p + 3); po_get_base_velocity(s, o, p + 7, p + 10); }
    memcpy(st + 18 + 13 * s->nobj, e->goal, g * sizeof(double));
}

/* ------------------------------------------------------------------ CPU baseline driver (bench.py cpu_baseline / --impl reference)
 * Random-action rollout of one env, reset on success or at the TimeLimit (test/envs_test.py:6-14 loop), xorshift actions. */
static double rnd01(unsigned long long *s) { *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17; return (double)(*s >> 11) / 9007199254740992.0; }
typedef struct { PoEnv *e; unsigned long long st; int t, task; unsigned char term; } PoBench;
static void bench_reset(PoBench *b) {
    int task = b->task; unsigned long long *st = &b->st; float obs[32], ag[6], dg[6];
    double goal[6] = {0.3 * rnd01(st) - 0.15, 0.3 * rnd01(st) - 0.15, task == PO_REACH ? 0.3 * rnd01(st) : 0.02, 0, 0, 0.06};
    double op[6] = {0.3 * rnd01(st) - 0.15, 0.3 * rnd01(st) - 0.15, task == PO_SLIDE ? 0.03 : 0.02, 0.3 * rnd01(st) - 0.15, 0.3 * rnd01(st) - 0.15, 0.06};
    if (task == PO_STACK) { goal[3] = goal[0]; goal[4] = goal[1]; }
    if (task == PO_FLIP) { goal[0] = goal[1] = goal[2] = 0; goal[3] = 1; }
    po_env_reset(b->e, goal, op, obs, ag, dg); b->t = 0; b->term = 0;
}
/* persistent form: one env kept across calls (bench.py's thread pool owns one per host thread) */
void *po_bench_open(int task, int control, unsigned long long seed) {
    PoBench *b = (PoBench *)calloc(1, sizeof(PoBench));
    b->e = po_env_create(task, control, PO_REWARD_SPARSE); b->task = task; b->st = seed * 2654435761ULL + 88172645463325252ULL;
    bench_reset(b);
    return b;
}
double po_bench_steps(void *h, int n_steps) {
    PoBench *b = (PoBench *)h; float obs[32], ag[6], dg[6], rew, act[8]; double acc = 0;
    int na = po_env_action_dim(b->e), limit = b->task == PO_STACK ? 100 : 50;
    for (int i = 0; i < n_steps; i++) {
        if (b->term || b->t >= limit) bench_reset(b);
        for (int k = 0; k < na; k++) act[k] = (float)(2 * rnd01(&b->st) - 1);
        po_env_step(b->e, act, obs, ag, dg, &rew, &b->term); b->t++;
        acc += obs[0] + rew;
    }
    return acc;
}
void po_bench_close(void *h) { PoBench *b = (PoBench *)h; po_env_destroy(b->e); free(b); }
double po_bench_run(int task, int control, int n_steps, unsigned long long seed) {
    void *h = po_bench_open(task, control, seed); double acc = po_bench_steps(h, n_steps); po_bench_close(h); return acc;
}
