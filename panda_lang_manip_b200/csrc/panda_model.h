// panda_model.h -- host-side construction of the compile-time-structured Panda model constants.
//
// What the reference gets from loadURDF("franka_panda/panda.urdf", useFixedBase=True) (reference
// panda_gym/envs/core.py:47-52) plus the motor forces it passes every step (panda_gym/envs/robots/panda.py:40-41).
// pybullet derives link inertias from the collision-shape AABB, not from the URDF; the meshes are
// not available, so the AABB extents below are approximations (link 5 calibrated to the reference's known-answer
// test test/pybullet_test.py:186).  Computed in double, then narrowed to the kernel's scalar type.
#pragma once
#include "panda_dyn.cuh"

namespace pg {

struct LinkConst { double xyz[3]; double mass; double com[3]; double box[3]; double lo, hi; };
// arm links 0..6, then hand (link 8, rigidly attached to link 6 through the massless link 7), finger 1, finger 2
static const LinkConst kLinks[10] = {
    {{0, 0, 0.333}, 2.7, {0, -0.04, -0.05}, {0.11, 0.13, 0.25}, -2.9671, 2.9671},
    {{0, 0, 0}, 2.73, {0, -0.04, 0.06}, {0.11, 0.25, 0.13}, -1.8326, 1.8326},
    {{0, -0.316, 0}, 2.04, {0.01, 0.01, -0.05}, {0.19, 0.15, 0.18}, -2.9671, 2.9671},
    {{0.0825, 0, 0}, 2.08, {-0.03, 0.03, 0.02}, {0.19, 0.18, 0.15}, -3.1416, 0.0},
    {{-0.0825, 0.384, 0}, 3.0, {0, 0.04, -0.12}, {0.11, 0.19, 0.32}, -2.9671, 2.9671},
    {{0, 0, 0}, 1.3, {0.04, 0, 0}, {0.20265085784266038, 0.13, 0.12}, -0.0873, 3.8223},
    {{0.088, 0, 0}, 0.2, {0, 0, 0.08}, {0.11, 0.11, 0.10}, -2.9671, 2.9671},
    {{0, 0, 0.107}, 0.81, {0, 0, 0.04}, {0.064, 0.204, 0.09}, 0, 0},            // hand: frame = link 6 + (0,0,0.107), yaw -45 deg
    {{0, 0, 0.0584}, 0.1, {0, 0.01, 0.02}, {0.021, 0.021, 0.054}, 0.0, 0.04},   // finger 1 (in hand axes, slides along +y)
    {{0, 0, 0.0584}, 0.1, {0, -0.01, 0.02}, {0.021, 0.021, 0.054}, 0.0, 0.04},  // finger 2 (slides along -y)
};
static const double kJointForces[9] = {87.0, 87.0, 87.0, 87.0, 12.0, 120.0, 120.0, 170.0, 170.0};  // panda.py:41
static const double kNeutral[9] = {0.00, 0.41, 0.00, -1.85, 0.00, 2.26, 0.79, 0.00, 0.00};          // panda.py:45
static const double kEeZ = 0.105;   // panda_grasptarget above the hand frame

inline void box_inertia(const LinkConst& L, double I[3]) {
    const double* b = L.box; double m = L.mass / 12.0;
    I[0] = m * (b[1] * b[1] + b[2] * b[2]); I[1] = m * (b[0] * b[0] + b[2] * b[2]); I[2] = m * (b[0] * b[0] + b[1] * b[1]);
}
// inertia about the frame origin of a part with principal inertia Ic (axes = frame axes rotated by yaw about z), CoM c
inline void origin_inertia(double mass, const double c[3], const double Ic[3], double yaw, double out[6]) {
    double cy = cos(yaw), sy = sin(yaw);
    double xx = cy * cy * Ic[0] + sy * sy * Ic[1], yy = sy * sy * Ic[0] + cy * cy * Ic[1], xy = cy * sy * (Ic[0] - Ic[1]);
    double cc = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    out[0] = xx + mass * (cc - c[0] * c[0]); out[1] = xy - mass * c[0] * c[1]; out[2] = -mass * c[0] * c[2];
    out[3] = yy + mass * (cc - c[1] * c[1]); out[4] = -mass * c[1] * c[2]; out[5] = Ic[2] + mass * (cc - c[2] * c[2]);
}

template <typename T> Model<T> make_model(const double base[3]) {
    Model<T> M;
    const double dt = 1.0 / 500.0;
    for (int k = 0; k < 3; k++) M.base[k] = (T)base[k];
    for (int i = 0; i < 7; i++) for (int k = 0; k < 3; k++) M.pT[i][k] = (T)kLinks[i].xyz[k];
    for (int b = 0; b < 10; b++) {
        double Ic[3]; box_inertia(kLinks[b], Ic);
        const LinkConst& L = kLinks[b];
        // damped parts: arm links 0..6 in their own frames, hand (part 7) expressed in link-6 coordinates, fingers (8, 9) in their own frames
        int part = b;
        double c[3] = {L.com[0], L.com[1], L.com[2]};
        if (b == 7) c[2] += L.xyz[2];
        M.dm[part] = (T)L.mass;
        for (int k = 0; k < 3; k++) { M.dc[part][k] = (T)c[k]; M.dI[part][k] = (T)Ic[k]; }
    }
    for (int b = 0; b < 9; b++) {
        // dynamic bodies: 0..5 = links, 6 = link 6 + hand, 7/8 = fingers
        const LinkConst& L = kLinks[b < 7 ? b : b + 1];
        double Ic[3], Io[6]; box_inertia(L, Ic); origin_inertia(L.mass, L.com, Ic, 0.0, Io);
        double m = L.mass, h[3] = {m * L.com[0], m * L.com[1], m * L.com[2]};
        if (b == 6) {
            const LinkConst& H = kLinks[7];
            double Ih[3], Ioh[6], ch[3] = {H.com[0], H.com[1], H.com[2] + H.xyz[2]};   // yaw about z keeps the CoM on the z axis
            box_inertia(H, Ih); origin_inertia(H.mass, ch, Ih, -0.78539816339744830962, Ioh);
            m += H.mass; for (int k = 0; k < 3; k++) h[k] += H.mass * ch[k];
            for (int k = 0; k < 6; k++) Io[k] += Ioh[k];
        }
        M.m[b] = (T)m;
        for (int k = 0; k < 3; k++) M.h[b][k] = (T)h[k];
        for (int k = 0; k < 6; k++) M.Io[b][k] = (T)Io[k];
        const LinkConst& J = kLinks[b < 7 ? b : b + 1];
        M.lo[b] = (T)J.lo; M.hi[b] = (T)J.hi; M.max_imp[b] = (T)(kJointForces[b] * dt);
    }
    M.z7 = (T)kLinks[7].xyz[2];
    M.hz = (T)(kLinks[7].xyz[2] + kLinks[8].xyz[2]);
    M.eez = (T)(kLinks[7].xyz[2] + kEeZ);
    M.fa[0] = (T)1; M.fa[1] = (T)-1;
    M.ee_scale = (T)0.05; M.finger_scale = (T)0.2;
    return M;
}

}  // namespace pg
