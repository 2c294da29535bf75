// hostcheck.cpp -- TEST-ONLY host compilation of the device math in panda_lang_manip_b200/csrc/panda_dyn.cuh.
// It lets the CPU test-suite (-m "not gpu") compare the kernel's formulation (RNEA + CRBA + Cholesky, unrolled) with the
// oracle's (ABA + impulse responses) without a GPU.  It is NOT a fallback: nothing in the package loads this library.
#define PG_HOST_DEBUG 1
#include "../../panda_lang_manip_b200/csrc/panda_model.h"
#include "../../panda_lang_manip_b200/csrc/panda_scene.h"
#include <string.h>
using namespace pg;

template <typename T> static void substeps(const double* base, double* q, double* qd, const double* target, int n) {
    Model<T> M = make_model<T>(base);
    T tq[ND], tqd[ND], tt[ND];
    for (int i = 0; i < ND; i++) { tq[i] = (T)q[i]; tqd[i] = (T)qd[i]; tt[i] = (T)target[i]; }
    for (int s = 0; s < n; s++) robot_substep(M, tq, tqd, tt);
    for (int i = 0; i < ND; i++) { q[i] = tq[i]; qd[i] = tqd[i]; }
}
template <typename T> static void minv(const double* base, const double* q, const double* qd, double* out, double* qdd) {
    Model<T> M = make_model<T>(base);
    T tq[ND], tqd[ND], sn[7], cs[7], Mi[ND][ND], a[ND];
    for (int i = 0; i < ND; i++) { tq[i] = (T)q[i]; tqd[i] = (T)qd[i]; }
    robot_dynamics(M, tq, tqd, sn, cs, Mi, a);
    for (int i = 0; i < ND; i++) { qdd[i] = a[i]; for (int j = 0; j < ND; j++) out[ND * i + j] = Mi[i][j]; }
}
template <typename T> static void observe(const double* base, const double* q, const double* qd, const double* qc, double* pos, double* vel) {
    Model<T> M = make_model<T>(base);
    T tq[ND], tqd[ND], tqc[ND];
    for (int i = 0; i < ND; i++) { tq[i] = (T)q[i]; tqd[i] = (T)qd[i]; tqc[i] = (T)qc[i]; }
    V3<T> p, v; ee_observe(M, tq, tqd, tqc, p, v);
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z; vel[0] = v.x; vel[1] = v.y; vel[2] = v.z;
}
template <typename T> static void ik(const double* base, const double* q, const double* target, const double* quat, double* out) {
    Model<T> M = make_model<T>(base);
    T tq[ND], o[7], qt[4];
    for (int i = 0; i < ND; i++) tq[i] = (T)q[i];
    for (int i = 0; i < 4; i++) qt[i] = (T)quat[i];
    ik_ee(M, tq, mk<T>((T)target[0], (T)target[1], (T)target[2]), qt, o);
    for (int i = 0; i < 7; i++) out[i] = o[i];
}
// state layout (doubles): q[9], qd[9], obj[2][13] = pos3 quat4 lin3 ang3, goal[6]
template <typename T, int TASK, int CTRL> static void env_step_t(int reward, const double* base, double* st, const float* action, float* obs, float* ag, float* dg, float* rew, unsigned char* succ) {
    Model<T> M = make_model<T>(base); Scene<T> S = make_scene<T>(TASK);
    constexpr int NOBJ = task_nobj(TASK);
    T q[ND], qd[ND]; double goal[6]; Obj<T> ob[NOBJ > 0 ? NOBJ : 1];
    for (int i = 0; i < ND; i++) { q[i] = (T)st[i]; qd[i] = (T)st[9 + i]; }
    for (int o = 0; o < NOBJ; o++) { const double* p = st + 18 + 13 * o; ob[o].pos = mk<T>((T)p[0], (T)p[1], (T)p[2]); ob[o].qx = (T)p[3]; ob[o].qy = (T)p[4]; ob[o].qz = (T)p[5]; ob[o].qw = (T)p[6]; ob[o].lin = mk<T>((T)p[7], (T)p[8], (T)p[9]); ob[o].ang = mk<T>((T)p[10], (T)p[11], (T)p[12]); }
    for (int k = 0; k < 6; k++) goal[k] = st[44 + k];
    static Contacts<T> C; static T slab[solver_slots(2)]; C.st.base = slab; C.st.stride = 1;
    int mc = 0; T target[ND]; env_step<T, TASK, CTRL>(M, S, reward, q, qd, ob, goal, action, nullptr, obs, ag, dg, *rew, *succ, C, mc, target);
    for (int i = 0; i < ND; i++) { st[i] = q[i]; st[9 + i] = qd[i]; }
    for (int o = 0; o < NOBJ; o++) { double* p = st + 18 + 13 * o; p[0] = ob[o].pos.x; p[1] = ob[o].pos.y; p[2] = ob[o].pos.z; p[3] = ob[o].qx; p[4] = ob[o].qy; p[5] = ob[o].qz; p[6] = ob[o].qw; p[7] = ob[o].lin.x; p[8] = ob[o].lin.y; p[9] = ob[o].lin.z; p[10] = ob[o].ang.x; p[11] = ob[o].ang.y; p[12] = ob[o].ang.z; }
}
template <typename T, int TASK> static void env_step_c(int ctrl, int reward, const double* base, double* st, const float* action, float* obs, float* ag, float* dg, float* rew, unsigned char* succ) {
    if (ctrl == CTRL_EE) env_step_t<T, TASK, CTRL_EE>(reward, base, st, action, obs, ag, dg, rew, succ); else env_step_t<T, TASK, CTRL_JOINTS>(reward, base, st, action, obs, ag, dg, rew, succ);
}
template <typename T> static void env_step_d(int task, int ctrl, int reward, const double* base, double* st, const float* action, float* obs, float* ag, float* dg, float* rew, unsigned char* succ) {
    switch (task) {
    case 0: env_step_c<T, 0>(ctrl, reward, base, st, action, obs, ag, dg, rew, succ); break;
    case 1: env_step_c<T, 1>(ctrl, reward, base, st, action, obs, ag, dg, rew, succ); break;
    case 2: env_step_c<T, 2>(ctrl, reward, base, st, action, obs, ag, dg, rew, succ); break;
    case 3: env_step_c<T, 3>(ctrl, reward, base, st, action, obs, ag, dg, rew, succ); break;
    case 4: env_step_c<T, 4>(ctrl, reward, base, st, action, obs, ag, dg, rew, succ); break;
    default: env_step_c<T, 5>(ctrl, reward, base, st, action, obs, ag, dg, rew, succ); break;
    }
}
// ---- bare world (generic motors), any-link getLinkState / IK: the math behind pg_sim_step / pg_get_link_state / pg_inverse_kinematics_link
template <typename T> static void link_state_t(const double* base, int link, const double* q, const double* qd, const double* qc, double* out) {
    Model<T> M = make_model<T>(base);
    T tq[ND], tqd[ND], tqc[ND], qt[4];
    for (int i = 0; i < ND; i++) { tq[i] = (T)q[i]; tqd[i] = (T)qd[i]; tqc[i] = (T)qc[i]; }
    V3<T> p, l, a; link_state(M, link, tq, tqd, tqc, p, qt, l, a);
    out[0] = p.x; out[1] = p.y; out[2] = p.z; for (int k = 0; k < 4; k++) out[3 + k] = qt[k];
    out[7] = l.x; out[8] = l.y; out[9] = l.z; out[10] = a.x; out[11] = a.y; out[12] = a.z;
}
template <typename T> static void ik_link_t(const double* base, int link, const double* q, const double* target, const double* quat, double* out) {
    Model<T> M = make_model<T>(base);
    T tq[ND], o[ND], qt[4]; double nn = 0;
    for (int i = 0; i < ND; i++) tq[i] = (T)q[i];
    for (int i = 0; i < 4; i++) nn += quat[i] * quat[i];
    for (int i = 0; i < 4; i++) qt[i] = (T)(quat[i] / sqrt(nn));
    ik_link(M, link, tq, mk<T>((T)target[0], (T)target[1], (T)target[2]), qt, o);
    for (int i = 0; i < ND; i++) out[i] = o[i];
}
// st: q[9] qd[9] obj[nobj][13]; motors [9][5] = kp kd target_q target_v max_force; scene: bodies [nobj][6] = shape hx hy hz mass mu, table_rect or NULL, ground_z or NULL
template <typename T, int NOBJ> static void bare_steps_t(const double* base, double* st, const double* motors, const double* bodies, const double* rect, const double* gz, int nsub) {
    Model<T> M = make_model<T>(base);
    Scene<double> Sd = make_scene<double>(TASK_REACH);
    Sd.nobj = NOBJ;
    for (int o = 0; o < NOBJ; o++) scene_set_obj(Sd, o, (int)bodies[6 * o], bodies[6 * o + 1], bodies[6 * o + 2], bodies[6 * o + 3], bodies[6 * o + 4], bodies[6 * o + 5]);
    if (rect) { Sd.table_x0 = rect[0]; Sd.table_x1 = rect[1]; Sd.table_y0 = rect[2]; Sd.table_y1 = rect[3]; } else { Sd.table_x0 = 1; Sd.table_x1 = -1; Sd.table_y0 = 1; Sd.table_y1 = -1; }
    Sd.ground_z = gz ? *gz : -1e30;
    Scene<T> S = scene_cast<T>(Sd);
    T q[ND], qd[ND], target[ND], mot[27]; Obj<T> ob[NOBJ > 0 ? NOBJ : 1];
    for (int i = 0; i < ND; i++) { q[i] = (T)st[i]; qd[i] = (T)st[9 + i]; const double* m = motors + 5 * i; mot[i] = (T)m[0]; mot[9 + i] = (T)m[1]; target[i] = (T)m[2]; mot[18 + i] = (T)m[3]; M.max_imp[i] = (T)(m[4] / 500.0); }
    for (int o = 0; o < NOBJ; o++) { const double* p = st + 18 + 13 * o; ob[o].pos = mk<T>((T)p[0], (T)p[1], (T)p[2]); ob[o].qx = (T)p[3]; ob[o].qy = (T)p[4]; ob[o].qz = (T)p[5]; ob[o].qw = (T)p[6]; ob[o].lin = mk<T>((T)p[7], (T)p[8], (T)p[9]); ob[o].ang = mk<T>((T)p[10], (T)p[11], (T)p[12]); }
    static Contacts<T> C; static T slab[solver_slots(2)]; C.st.base = slab; C.st.stride = 1;
    bool full = true, act = false;
    for (int s = 0; s < nsub; s++) env_substep<T, NOBJ, false, true>(M, S, q, qd, target, ob, C, full, act, mot);
    for (int i = 0; i < ND; i++) { st[i] = q[i]; st[9 + i] = qd[i]; }
    for (int o = 0; o < NOBJ; o++) { double* p = st + 18 + 13 * o; p[0] = ob[o].pos.x; p[1] = ob[o].pos.y; p[2] = ob[o].pos.z; p[3] = ob[o].qx; p[4] = ob[o].qy; p[5] = ob[o].qz; p[6] = ob[o].qw; p[7] = ob[o].lin.x; p[8] = ob[o].lin.y; p[9] = ob[o].lin.z; p[10] = ob[o].ang.x; p[11] = ob[o].ang.y; p[12] = ob[o].ang.z; }
}
template <typename T> static void bare_steps_n(int nobj, const double* base, double* st, const double* motors, const double* bodies, const double* rect, const double* gz, int nsub) {
    if (nobj == 0) bare_steps_t<T, 0>(base, st, motors, bodies, rect, gz, nsub); else if (nobj == 1) bare_steps_t<T, 1>(base, st, motors, bodies, rect, gz, nsub); else bare_steps_t<T, 2>(base, st, motors, bodies, rect, gz, nsub);
}
extern "C" {
void hc_link_state(int dbl, const double* base, int link, const double* q, const double* qd, const double* qc, double* out) { if (dbl) link_state_t<double>(base, link, q, qd, qc, out); else link_state_t<float>(base, link, q, qd, qc, out); }
void hc_ik_link(int dbl, const double* base, int link, const double* q, const double* target, const double* quat, double* out) { if (dbl) ik_link_t<double>(base, link, q, target, quat, out); else ik_link_t<float>(base, link, q, target, quat, out); }
void hc_bare_steps(int dbl, int nobj, const double* base, double* st, const double* motors, const double* bodies, const double* rect, const double* gz, int nsub) {
    if (dbl) bare_steps_n<double>(nobj, base, st, motors, bodies, rect, gz, nsub); else bare_steps_n<float>(nobj, base, st, motors, bodies, rect, gz, nsub);
}
long hc_dbg_fallbacks() { return pg::g_dbg_fallbacks; }
long hc_dbg_full_starts() { return pg::g_dbg_full_starts; }
int hc_dbg_trace(int* out) { int n = pg::g_dbg_ntrace; for (int i = 0; i < n; i++) out[i] = pg::g_dbg_trace[i]; pg::g_dbg_ntrace = 0; return n; }
void hc_dbg_solver(long* out) { out[0] = pg::g_dbg_sweeps; out[1] = pg::g_dbg_solves; out[2] = pg::g_dbg_contacts; }
void hc_env_step(int dbl, int task, int ctrl, int reward, const double* base, double* st, const float* action, float* obs, float* ag, float* dg, float* rew, unsigned char* succ) {
    if (dbl) env_step_d<double>(task, ctrl, reward, base, st, action, obs, ag, dg, rew, succ); else env_step_d<float>(task, ctrl, reward, base, st, action, obs, ag, dg, rew, succ);
}
void hc_substeps(int dbl, const double* base, double* q, double* qd, const double* target, int n) { if (dbl) substeps<double>(base, q, qd, target, n); else substeps<float>(base, q, qd, target, n); }
void hc_minv(int dbl, const double* base, const double* q, const double* qd, double* out, double* qdd) { if (dbl) minv<double>(base, q, qd, out, qdd); else minv<float>(base, q, qd, out, qdd); }
void hc_observe(int dbl, const double* base, const double* q, const double* qd, const double* qc, double* pos, double* vel) { if (dbl) observe<double>(base, q, qd, qc, pos, vel); else observe<float>(base, q, qd, qc, pos, vel); }
void hc_ik(int dbl, const double* base, const double* q, const double* target, const double* quat, double* out) { if (dbl) ik<double>(base, q, target, quat, out); else ik<float>(base, q, target, quat, out); }
}
