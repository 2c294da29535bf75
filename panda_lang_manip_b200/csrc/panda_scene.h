// panda_scene.h -- host-side scene constants per task (what the reference's Task._create_scene builds:
// reference panda_gym/envs/tasks/{reach.py:28-38, push.py:30-47, slide.py:31-51, pick_and_place.py:32-50, stack.py:30-62,
// flip.py:29-48}; table/plane: panda_gym/pybullet.py:726-771; finger friction: panda_gym/envs/robots/panda.py:47-50).
#pragma once
#include <string.h>
#include "panda_env.cuh"

namespace pg {

template <typename T> void scene_set_obj(Scene<T>& S, int o, int shape, double hx, double hy, double hz, double mass, double mu) {
    S.shape[o] = shape; S.half[o][0] = (T)hx; S.half[o][1] = (T)hy; S.half[o][2] = (T)hz; S.mass[o] = (T)mass; S.mu[o] = (T)mu;
    if (shape == SH_BOX) {      // btBoxShape::calculateLocalInertia
        double lx = 2 * hx, ly = 2 * hy, lz = 2 * hz;
        S.Ic[o][0] = (T)(mass / 12 * (ly * ly + lz * lz)); S.Ic[o][1] = (T)(mass / 12 * (lx * lx + lz * lz)); S.Ic[o][2] = (T)(mass / 12 * (lx * lx + ly * ly));
    } else {                    // btCylinderShapeZ
        double r = hx, h = 2 * hz;
        S.Ic[o][0] = S.Ic[o][1] = (T)(mass / 12 * h * h + mass / 4 * r * r); S.Ic[o][2] = (T)(mass / 2 * r * r);
    }
}
template <typename T, typename U> Scene<T> scene_cast(const Scene<U>& A) {
    Scene<T> S;
    S.nobj = A.nobj;
    for (int o = 0; o < MAXOBJ; o++) { S.shape[o] = A.shape[o]; S.mass[o] = (T)A.mass[o]; S.mu[o] = (T)A.mu[o]; for (int k = 0; k < 3; k++) { S.half[o][k] = (T)A.half[o][k]; S.Ic[o][k] = (T)A.Ic[o][k]; } }
    S.table_x0 = (T)A.table_x0; S.table_x1 = (T)A.table_x1; S.table_y0 = (T)A.table_y0; S.table_y1 = (T)A.table_y1;
    for (int b = 0; b < 3; b++) { S.rb_mu[b] = (T)A.rb_mu[b]; for (int k = 0; k < 3; k++) { S.rb_c[b][k] = (T)A.rb_c[b][k]; S.rb_h[b][k] = (T)A.rb_h[b][k]; } }
    S.margin = (T)A.margin; S.margin_grasp = (T)A.margin_grasp; S.ground_z = (T)A.ground_z; S.table_mu = (T)A.table_mu; S.soft_erp = (T)A.soft_erp; S.soft_cfm = (T)A.soft_cfm;
    return S;
}
template <typename T> Scene<T> make_scene(int task) {
    Scene<T> S;
    memset(&S, 0, sizeof S);
    S.nobj = task_nobj(task);
    S.table_x0 = (T)-0.85; S.table_x1 = (T)0.25; S.table_y0 = (T)-0.35; S.table_y1 = (T)0.35;   // 1.1 x 0.7 table, x offset -0.3
    for (int o = 0; o < MAXOBJ; o++) scene_set_obj(S, o, SH_BOX, 0.02, 0.02, 0.02, 1.0, 0.5);
    if (task == TASK_SLIDE) { S.table_x0 = (T)-0.8; S.table_x1 = (T)0.6; scene_set_obj(S, 0, SH_CYL, 0.03, 0.03, 0.015, 1.0, 0.04); }   // slide.py:33-42
    if (task == TASK_STACK) scene_set_obj(S, 0, SH_BOX, 0.02, 0.02, 0.02, 2.0, 0.5);   // stack.py:33-39
    const double rbc[3][3] = {{0, 0, 0.021}, {0, 0.0105, 0.027}, {0, -0.0105, 0.027}};
    const double rbh[3][3] = {{0.032, 0.102, 0.045}, {0.0105, 0.0105, 0.027}, {0.0105, 0.0105, 0.027}};
    const double rbmu[3] = {0.5, 1.0, 1.0};
    for (int b = 0; b < 3; b++) { for (int k = 0; k < 3; k++) { S.rb_c[b][k] = (T)rbc[b][k]; S.rb_h[b][k] = (T)rbh[b][k]; } S.rb_mu[b] = (T)rbmu[b]; }
    S.margin = (T)0.004; S.margin_grasp = (T)0.012; S.ground_z = (T)-0.4; S.table_mu = (T)0.5;
    const double dt = 1.0 / 500.0, k = 30000.0, d = 1000.0;
    S.soft_erp = (T)(dt * k / (dt * k + d)); S.soft_cfm = (T)(1.0 / (dt * k + d) / dt);
    return S;
}

}  // namespace pg
