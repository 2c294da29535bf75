// panda_b200.cu -- C ABI of libpanda_b200.so (include/panda_b200.h): handle management, SoA state allocation, snapshots,
// kernel dispatch, the host-buffer entry points and the HER reward kernels.  No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <string>
#include <vector>
#include "panda_b200.h"
#include "panda_kernels.cuh"
#include "panda_model.h"
#include "panda_scene.h"
#include "panda_render.h"

namespace pg {
long long g_launches = 0;
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* what) { return fail(PG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }
#define PG_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return cuda_fail(e_, #x); } while (0)

template <typename T, int TASK> cudaError_t configure_step(void);       // panda_step_task.cu
template <typename T> cudaError_t configure_bare(void);                 // panda_bare.cu
template <typename T> void launch_render_setup(const EnvDev<T>& E, const RenderScene& R, RenderPrim* prims, cudaStream_t st);   // panda_render.cu
void launch_render(const RenderPrim* prims, const RenderCamera& C, int n_envs, float* depth, unsigned char* rgba, unsigned char* seg, float* points, unsigned char* valid, cudaStream_t st);
#define PG_DECL(T, K)                                                                                        \
    extern template void launch_step<T, K>(const EnvDev<T>&, int, const StepIO&, cudaStream_t);              \
    extern template void launch_reset<T, K>(const EnvDev<T>&, const ResetIO&, cudaStream_t);                 \
    extern template cudaError_t configure_step<T, K>(void);
PG_DECL(float, 0) PG_DECL(float, 1) PG_DECL(float, 2) PG_DECL(float, 3) PG_DECL(float, 4) PG_DECL(float, 5)
PG_DECL(double, 0) PG_DECL(double, 1) PG_DECL(double, 2) PG_DECL(double, 3) PG_DECL(double, 4) PG_DECL(double, 5)

template <typename T> struct Dispatch {
    static void step(int task, const EnvDev<T>& E, int ctrl, const StepIO& io, cudaStream_t st) {
        switch (task) {
        case 0: launch_step<T, 0>(E, ctrl, io, st); break; case 1: launch_step<T, 1>(E, ctrl, io, st); break;
        case 2: launch_step<T, 2>(E, ctrl, io, st); break; case 3: launch_step<T, 3>(E, ctrl, io, st); break;
        case 4: launch_step<T, 4>(E, ctrl, io, st); break; default: launch_step<T, 5>(E, ctrl, io, st); break;
        }
    }
    static void reset(int task, const EnvDev<T>& E, const ResetIO& io, cudaStream_t st) {
        switch (task) {
        case 0: launch_reset<T, 0>(E, io, st); break; case 1: launch_reset<T, 1>(E, io, st); break;
        case 2: launch_reset<T, 2>(E, io, st); break; case 3: launch_reset<T, 3>(E, io, st); break;
        case 4: launch_reset<T, 4>(E, io, st); break; default: launch_reset<T, 5>(E, io, st); break;
        }
    }
    static cudaError_t configure(int task) {
        switch (task) {
        case 0: return configure_step<T, 0>(); case 1: return configure_step<T, 1>(); case 2: return configure_step<T, 2>();
        case 3: return configure_step<T, 3>(); case 4: return configure_step<T, 4>(); case 5: return configure_step<T, 5>();
        default: return configure_bare<T>();
        }
    }
};
}  // namespace pg

using namespace pg;

struct pg_env {
    int task, ctrl, reward, n, device, precision;
    int obs_dim, goal_dim, act_dim, max_steps, state_dim, nobj;
    void* blob = nullptr; size_t blob_bytes = 0;      // one allocation: q, qd, obj, goal, steps, episode, ret, stats
    EnvDev<float> Ef; EnvDev<double> Ed;
    std::map<int, void*> snaps; int next_snap = 0;
    RenderPrim* prims = nullptr; RenderScene rscene;  // analytic renderer: per-env primitive lists (allocated at the first pg_render)
    std::vector<void*> snap_pool;                    // buffers of removed snapshots, reused by the next save (no allocation in a save / remove loop)
    bool sort_envs = true;                            // PG_SORT_ENVS=0 disables the contact-aware thread->env map (A/B measurements)
    // env groups: sorted batches are cut into groups of consecutive envs, each advanced on its own stream, so that the tail of
    // one group's launch (a few contact-heavy blocks) overlaps with the other groups' launches instead of idling the GPU
    long long* dbg = nullptr;
    int groups = 1; cudaStream_t gstream[8] = {}; cudaEvent_t ev_fork = nullptr, ev_join[8] = {};
    int segments = 4;                                 // launches per step for sorted batches (see pg_create); PG_SEGMENTS overrides (a divisor of 20)
    // host-buffer path: pinned staging + device I/O buffers + private stream
    cudaStream_t hstream = nullptr;
    float *h_act = nullptr, *h_out = nullptr; float* d_act = nullptr; float* d_out = nullptr; size_t out_floats = 0; size_t out_bytes = 0;
};

template <typename T> static void bind(EnvDev<T>& E, pg_env* e, char* base, unsigned long long seed, long long id0) {
    const size_t n = (size_t)e->n;
    E.n = e->n; E.reward_type = e->reward; E.id0 = id0; E.seed = seed;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base + off; off += (bytes + 255) & ~(size_t)255; return p; };
    E.stats = (double*)take(8 * sizeof(double));
    E.q = (T*)take(9 * n * sizeof(T)); E.qd = (T*)take(9 * n * sizeof(T));
    E.obj = (T*)take((size_t)(e->nobj > 0 ? e->nobj : 1) * 13 * n * sizeof(T));
    E.goal = (double*)take(6 * n * sizeof(double));
    E.target = (T*)take(9 * n * sizeof(T));
    E.motor = (T*)take((e->task == PG_TASK_BARE ? 36 : 0) * n * sizeof(T));
    E.steps = (int*)take(n * sizeof(int)); E.episode = (unsigned*)take(n * sizeof(unsigned)); E.ret = (float*)take(n * sizeof(float));
    E.ccount = (unsigned short*)take(n * sizeof(unsigned short)); E.perm = (int*)take(n * sizeof(int));
    E.hist = (int*)take(((n + PERM_CHUNK - 1) / PERM_CHUNK + 8) * PERM_BUCKETS * sizeof(int));
    e->blob_bytes = off;
}

template <typename E, bool WANT_REWARD> static void launch_reward(int task, const E* ag, const E* dg, float* reward, unsigned char* success, long long m, int reward_type, double thr, cudaStream_t st) {
    const int vec_ok = (((uintptr_t)ag | (uintptr_t)dg) & 15) == 0;
    long long blocks = (m / 2 + 255) / 256 + 1;
    int grid = (int)(blocks < 148LL * 32 ? blocks : 148LL * 32);   // grid-stride; capped at a multiple of the SM count
    switch (task) {
    case 4: reward_kernel<E, 4, WANT_REWARD><<<grid, 256, 0, st>>>(ag, dg, reward, success, m, reward_type, vec_ok, thr); break;
    case 5: reward_kernel<E, 5, WANT_REWARD><<<grid, 256, 0, st>>>(ag, dg, reward, success, m, reward_type, vec_ok, thr); break;
    default: reward_kernel<E, 0, WANT_REWARD><<<grid, 256, 0, st>>>(ag, dg, reward, success, m, reward_type, vec_ok, thr); break;   // all 3-D position goals share thr 0.05
    }
    g_launches++;
}

// widest load (bytes) every row start allows: base pointers and the row pitch must both be multiples of it; a load that reaches past
// the row's G elements (16-byte loads on 6-D fp32 rows) needs the padding to exist, i.e. pitch >= the loaded span
template <typename E> static int her_vector_bytes(int G, const void* a, const void* b, long long pitch) {
    const uintptr_t bits = (uintptr_t)a | (uintptr_t)b | (uintptr_t)(pitch * (long long)sizeof(E));
    for (int vb = 16; vb > (int)sizeof(E); vb >>= 1) {
        const int ev = vb / (int)sizeof(E), span = (G + ev - 1) / ev * ev;
        if ((bits & (uintptr_t)(vb - 1)) == 0 && span <= pitch) return vb;
    }
    return (int)sizeof(E);
}
template <typename E, int TASK, int HT> static void launch_her_ht(const E* next_ag, const E* dg, const long long* src, const long long* goal_src, E* dg_out, E* ag_out, float* reward,
                                                                  long long m, long long pitch, int reward_type, double thr, cudaStream_t st) {
    const long long per_block = 256LL * HT;
    const int grid = (int)std::min<long long>((m + per_block - 1) / per_block, 148LL * 32);
    const int vb = her_vector_bytes<E>(task_goal_dim(TASK), next_ag, dg, pitch);
    if (vb == 16) her_relabel_kernel<E, TASK, 16, HT><<<grid, 256, 0, st>>>(next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, pitch, reward_type, thr);
    else if (vb == 8) her_relabel_kernel<E, TASK, 8, HT><<<grid, 256, 0, st>>>(next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, pitch, reward_type, thr);
    else her_relabel_kernel<E, TASK, (int)sizeof(E), HT><<<grid, 256, 0, st>>>(next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, pitch, reward_type, thr);
    g_launches++;
}
template <typename E, int TASK> static void launch_her_task(const E* next_ag, const E* dg, const long long* src, const long long* goal_src, E* dg_out, E* ag_out, float* reward,
                                                            long long m, long long pitch, int reward_type, double thr, cudaStream_t st) {
    static const int ht = [] { const char* v = getenv("PG_HER_INFLIGHT"); const int k = v ? atoi(v) : HER_INFLIGHT; return (k == 1 || k == 2 || k == 4 || k == 8) ? k : HER_INFLIGHT; }();
    switch (ht) {
    case 1: launch_her_ht<E, TASK, 1>(next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, pitch, reward_type, thr, st); break;
    case 2: launch_her_ht<E, TASK, 2>(next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, pitch, reward_type, thr, st); break;
    case 8: launch_her_ht<E, TASK, 8>(next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, pitch, reward_type, thr, st); break;
    default: launch_her_ht<E, TASK, 4>(next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, pitch, reward_type, thr, st); break;
    }
}
template <typename E> static void launch_her(int task, const E* next_ag, const E* dg, const long long* src, const long long* goal_src, E* dg_out, E* ag_out, float* reward,
                                             long long m, long long pitch, int reward_type, double thr, cudaStream_t st) {
    switch (task) {     // goal layouts: 3-D position (thr 0.05), Stack 6-D (thr 0.1), Flip quaternion (thr 0.2)
    case 4: launch_her_task<E, 4>(next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, pitch, reward_type, thr, st); break;
    case 5: launch_her_task<E, 5>(next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, pitch, reward_type, thr, st); break;
    default: launch_her_task<E, 0>(next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, pitch, reward_type, thr, st); break;
    }
}

extern "C" {

const char* pg_last_error(void) { return g_err.c_str(); }
long long pg_kernel_launches(void) { return g_launches; }

static int check_device(int device, const char* who) {
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return fail(PG_ERR_CUDA, std::string(who) + ": no CUDA device (" + cudaGetErrorString(ce) + "); this library has no CPU path");
    if (device < 0 || device >= ndev) return fail(PG_ERR_ARG, std::string(who) + ": bad device index");
    PG_CUDA(cudaSetDevice(device));
    return PG_OK;
}
// state allocation + model; the caller fills the scene
static int alloc_state(pg_env* e, unsigned long long seed, long long env_id_offset, const double base[3]) {
    if (e->precision == PG_F32) { bind(e->Ef, e, nullptr, seed, env_id_offset); } else { bind(e->Ed, e, nullptr, seed, env_id_offset); }
    cudaError_t err = cudaMalloc(&e->blob, e->blob_bytes);
    if (err != cudaSuccess) return cuda_fail(err, "cudaMalloc(state)");
    PG_CUDA(cudaMemset(e->blob, 0, e->blob_bytes));
    if (e->precision == PG_F32) { bind(e->Ef, e, (char*)e->blob, seed, env_id_offset); e->Ef.M = make_model<float>(base); }
    else { bind(e->Ed, e, (char*)e->blob, seed, env_id_offset); e->Ed.M = make_model<double>(base); }
    // the > 48 kB dynamic shared memory opt-in is per function and per device: done for this handle's device, errors reported
    cudaError_t ce = e->precision == PG_F32 ? Dispatch<float>::configure(e->task) : Dispatch<double>::configure(e->task);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaFuncSetAttribute(step kernel shared memory)");
    return PG_OK;
}

int pg_create(int task, int control_type, int reward_type, int num_envs, int device, unsigned long long seed, long long env_id_offset,
              int precision, pg_env** out) {
    if (!out) return fail(PG_ERR_ARG, "pg_create: out is NULL");
    if (task < 0 || task > 5 || control_type < 0 || control_type > 1 || reward_type < 0 || reward_type > 1 || num_envs <= 0 || precision < 0 || precision > 1)
        return fail(PG_ERR_ARG, "pg_create: bad task / control_type / reward_type / num_envs / precision");
    int rc = check_device(device, "pg_create"); if (rc != PG_OK) return rc;
    pg_env* e = new pg_env();
    e->task = task; e->ctrl = control_type; e->reward = reward_type; e->n = num_envs; e->device = device; e->precision = precision;
    e->nobj = task_nobj(task); e->obs_dim = task_obs_dim(task); e->goal_dim = task_goal_dim(task); e->act_dim = task_act_dim(task, control_type);
    e->max_steps = task_max_steps(task); e->state_dim = 18 + 13 * e->nobj + e->goal_dim + 1;
    const double base[3] = {-0.6, 0.0, 0.0};   // panda_tasks.py:26,43,60,77,94,111
    rc = alloc_state(e, seed, env_id_offset, base);
    if (rc != PG_OK) { pg_destroy(e); return rc; }
    e->Ef.S = make_scene<float>(task); e->Ed.S = make_scene<double>(task);
    e->Ef.P = e->Ed.P = make_task_params(task);
    { const char* v = getenv("PG_SORT_ENVS"); if (v && v[0] == '0') e->sort_envs = false; }
    const bool light = e->nobj == 0 && control_type == CTRL_JOINTS;
    e->segments = light ? 4 : 20;
    { const char* v = getenv("PG_DEBUG_TIMING"); if (v && v[0] == '1') { PG_CUDA(cudaMalloc(&e->dbg, (size_t)e->n * 2 * sizeof(long long))); PG_CUDA(cudaMemset(e->dbg, 0, (size_t)e->n * 2 * sizeof(long long))); } }
    e->Ef.dbg = e->dbg; e->Ed.dbg = e->dbg;
    // measured defaults (profiles/r1_notes.md): contact scenes and ee control re-sort before every sub-step and run 4 env groups (2 for ee-controlled Reach);
    // joint-controlled Reach (few contacts) keeps 4 launches per step on one stream
    e->groups = (num_envs >= 16384 && !light) ? (e->nobj == 0 ? 2 : 4) : 1;
    { const char* v = getenv("PG_GROUPS"); if (v) { int k = atoi(v); if (k >= 1 && k <= 8) e->groups = k; } }
    if (e->groups > 1) {
        PG_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
        for (int g = 0; g < e->groups; g++) { PG_CUDA(cudaStreamCreateWithFlags(&e->gstream[g], cudaStreamNonBlocking)); PG_CUDA(cudaEventCreateWithFlags(&e->ev_join[g], cudaEventDisableTiming)); }
    }
    { const char* v = getenv("PG_SEGMENTS"); if (v) { int k = atoi(v); if (k >= 1 && k <= 20 && 20 % k == 0) e->segments = k; } }
    *out = e;
    rc = pg_reset(e, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc != PG_OK) { pg_destroy(e); *out = nullptr; return rc; }
    PG_CUDA(cudaDeviceSynchronize());
    return PG_OK;
}

int pg_create_bare(int num_envs, int device, int precision, const double* robot_base, int n_bodies, const double* bodies, const double* table_rect,
                   const double* ground_z, pg_env** out) {
    if (!out) return fail(PG_ERR_ARG, "pg_create_bare: out is NULL");
    if (num_envs <= 0 || precision < 0 || precision > 1 || n_bodies < 0 || n_bodies > MAXOBJ || (n_bodies > 0 && !bodies))
        return fail(PG_ERR_ARG, "pg_create_bare: bad num_envs / precision / n_bodies (at most 2 free bodies)");
    for (int o = 0; o < n_bodies; o++) {
        const double* b = bodies + 13 * o;
        if ((b[0] != SH_BOX && b[0] != SH_CYL) || !(b[1] > 0) || !(b[2] > 0) || !(b[3] > 0) || !(b[4] > 0) || !(b[5] >= 0))
            return fail(PG_ERR_ARG, "pg_create_bare: body rows are [shape(0 box,1 z-cylinder) hx hy hz mass(>0) mu pos3 quat4]");
    }
    int rc = check_device(device, "pg_create_bare"); if (rc != PG_OK) return rc;
    pg_env* e = new pg_env();
    e->task = PG_TASK_BARE; e->ctrl = CTRL_JOINTS; e->reward = 0; e->n = num_envs; e->device = device; e->precision = precision;
    e->nobj = n_bodies; e->obs_dim = 0; e->goal_dim = 0; e->act_dim = 0; e->max_steps = 0; e->state_dim = 18 + 13 * n_bodies + 1;
    e->sort_envs = false;
    // a world without a robot still carries one (the kernels are built around it): it is parked 1 km up, out of reach of everything
    const double parked[3] = {0.0, 0.0, 1000.0};
    rc = alloc_state(e, 0, 0, robot_base ? robot_base : parked);
    if (rc != PG_OK) { pg_destroy(e); return rc; }
    Scene<double> S = make_scene<double>(TASK_REACH);
    S.nobj = n_bodies;
    for (int o = 0; o < n_bodies; o++) { const double* b = bodies + 13 * o; scene_set_obj(S, o, (int)b[0], b[1], b[2], b[3], b[4], b[5]); }
    if (table_rect) { S.table_x0 = table_rect[0]; S.table_x1 = table_rect[1]; S.table_y0 = table_rect[2]; S.table_y1 = table_rect[3]; }
    else { S.table_x0 = 1.0; S.table_x1 = -1.0; S.table_y0 = 1.0; S.table_y1 = -1.0; }     // empty rectangle: no table
    S.ground_z = ground_z ? *ground_z : -1e30;                                               // no plane: nothing to land on
    e->Ed.S = S; e->Ef.S = scene_cast<float>(S);
    e->Ef.P = e->Ed.P = make_task_params(TASK_REACH);
    // initial state: joints at 0 (loadURDF), bodies at their creation poses, velocity motors (target 0, max impulse 1 per sub-step)
    std::vector<double> st((size_t)num_envs * e->state_dim, 0.0), mot((size_t)num_envs * 45, 0.0);
    for (int i = 0; i < num_envs; i++) {
        double* r = st.data() + (size_t)i * e->state_dim;
        for (int o = 0; o < n_bodies; o++) {
            const double* b = bodies + 13 * o; double* p = r + 18 + 13 * o;
            double nq = sqrt(b[9] * b[9] + b[10] * b[10] + b[11] * b[11] + b[12] * b[12]);
            if (!(nq > 0)) { pg_destroy(e); return fail(PG_ERR_ARG, "pg_create_bare: zero quaternion"); }
            p[0] = b[6]; p[1] = b[7]; p[2] = b[8]; p[3] = b[9] / nq; p[4] = b[10] / nq; p[5] = b[11] / nq; p[6] = b[12] / nq;
        }
        for (int d = 0; d < 9; d++) { double* m = mot.data() + ((size_t)i * 9 + d) * 5; m[0] = 0.0; m[1] = 1.0; m[2] = 0.0; m[3] = 0.0; m[4] = 1.0 * 500.0; }
    }
    double *d_st = nullptr, *d_mot = nullptr;
    cudaError_t ce = cudaMalloc(&d_st, st.size() * sizeof(double)); if (ce == cudaSuccess) ce = cudaMalloc(&d_mot, mot.size() * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMemcpy(d_st, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(d_mot, mot.data(), mot.size() * sizeof(double), cudaMemcpyHostToDevice);
    rc = ce == cudaSuccess ? PG_OK : cuda_fail(ce, "pg_create_bare: initial state upload");
    if (rc == PG_OK) rc = pg_set_state(e, d_st, nullptr, nullptr);
    if (rc == PG_OK) rc = pg_set_motors(e, d_mot, nullptr, nullptr);
    if (rc == PG_OK) { ce = cudaDeviceSynchronize(); if (ce != cudaSuccess) rc = cuda_fail(ce, "pg_create_bare"); }
    cudaFree(d_st); cudaFree(d_mot);
    if (rc != PG_OK) { pg_destroy(e); return rc; }
    *out = e;
    return PG_OK;
}

int pg_destroy(pg_env* e) {
    if (!e) return PG_OK;
    cudaSetDevice(e->device);
    for (auto& kv : e->snaps) cudaFree(kv.second);
    for (void* p : e->snap_pool) cudaFree(p);
    if (e->h_act) cudaFreeHost(e->h_act);
    if (e->h_out) cudaFreeHost(e->h_out);
    if (e->d_act) cudaFree(e->d_act);
    if (e->d_out) cudaFree(e->d_out);
    if (e->hstream) cudaStreamDestroy(e->hstream);
    for (int g = 0; g < 8; g++) { if (e->gstream[g]) cudaStreamDestroy(e->gstream[g]); if (e->ev_join[g]) cudaEventDestroy(e->ev_join[g]); }
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->dbg) cudaFree(e->dbg);
    if (e->prims) cudaFree(e->prims);
    cudaFree(e->blob);
    delete e;
    return PG_OK;
}

int pg_dims(const pg_env* e, int* obs_dim, int* goal_dim, int* action_dim, int* max_episode_steps, int* state_dim) {
    if (!e) return fail(PG_ERR_ARG, "pg_dims: NULL handle");
    if (obs_dim) *obs_dim = e->obs_dim; if (goal_dim) *goal_dim = e->goal_dim; if (action_dim) *action_dim = e->act_dim;
    if (max_episode_steps) *max_episode_steps = e->max_steps; if (state_dim) *state_dim = e->state_dim;
    return PG_OK;
}

int pg_reset(pg_env* e, const unsigned char* mask, const double* goal_override, const double* object_override, float* obs, float* ag, float* dg, void* stream) {
    return pg_reset_seeded(e, mask, nullptr, goal_override, object_override, obs, ag, dg, stream);
}
int pg_reset_seeded(pg_env* e, const unsigned char* mask, const unsigned long long* seeds, const double* goal_override, const double* object_override, float* obs, float* ag,
                    float* dg, void* stream) {
    if (!e) return fail(PG_ERR_ARG, "pg_reset: NULL handle");
    if (e->task == PG_TASK_BARE) return fail(PG_ERR_ARG, "pg_reset: a bare world has no task to reset; use pg_set_state");
    PG_CUDA(cudaSetDevice(e->device));
    ResetIO io{mask, goal_override, object_override, obs, ag, dg, seeds};
    if (e->precision == PG_F32) Dispatch<float>::reset(e->task, e->Ef, io, (cudaStream_t)stream); else Dispatch<double>::reset(e->task, e->Ed, io, (cudaStream_t)stream);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}

int pg_step(pg_env* e, const float* actions, float* obs, float* ag, float* dg, float* reward, unsigned char* terminated, unsigned char* truncated,
            int auto_reset, void* stream) {
    return pg_step_oriented(e, actions, nullptr, obs, ag, dg, reward, terminated, truncated, auto_reset, stream);
}
int pg_set_action_scale(pg_env* e, double ee_scale, double finger_scale) {
    if (!e || !(ee_scale > 0) || !(finger_scale > 0)) return fail(PG_ERR_ARG, "pg_set_action_scale: bad argument");
    e->Ef.M.ee_scale = (float)ee_scale; e->Ef.M.finger_scale = (float)finger_scale; e->Ed.M.ee_scale = ee_scale; e->Ed.M.finger_scale = finger_scale;
    return PG_OK;
}
int pg_set_task_params(pg_env* e, double distance_threshold, const double* goal_range_low, const double* goal_range_high, const double* obj_range_low, const double* obj_range_high) {
    if (!e || e->task == PG_TASK_BARE) return fail(PG_ERR_ARG, "pg_set_task_params: needs a task handle");
    if (!(distance_threshold >= 0)) return fail(PG_ERR_ARG, "pg_set_task_params: distance_threshold must be >= 0");
    TaskParams P = e->Ef.P;
    P.thr64 = distance_threshold; P.thr32 = (float)distance_threshold;
    for (int k = 0; k < 3; k++) { if (goal_range_low) P.goal_lo[k] = goal_range_low[k]; if (goal_range_high) P.goal_hi[k] = goal_range_high[k]; }
    for (int k = 0; k < 2; k++) { if (obj_range_low) P.obj_lo[k] = obj_range_low[k]; if (obj_range_high) P.obj_hi[k] = obj_range_high[k]; }
    for (int k = 0; k < 3; k++) if (!(P.goal_lo[k] <= P.goal_hi[k])) return fail(PG_ERR_ARG, "pg_set_task_params: goal range low > high");
    for (int k = 0; k < 2; k++) if (!(P.obj_lo[k] <= P.obj_hi[k])) return fail(PG_ERR_ARG, "pg_set_task_params: object range low > high");
    e->Ef.P = e->Ed.P = P;
    return PG_OK;
}
int pg_set_substeps(pg_env* e, int n_substeps) {
    if (!e || e->task == PG_TASK_BARE) return fail(PG_ERR_ARG, "pg_set_substeps: needs a task handle (bare worlds pass the count to pg_sim_step)");
    if (n_substeps < 1 || n_substeps > 1000) return fail(PG_ERR_ARG, "pg_set_substeps: 1 <= n_substeps <= 1000");
    e->Ef.P.nsub = e->Ed.P.nsub = n_substeps;
    return PG_OK;
}
// Host-buffer mode of a step (pg_step_host): every env group copies its own action rows in before its first launch and its own
// output rows out after its last one, on its own stream -- group g's device-to-host copies overlap group g+1's kernels.
struct HostIO { const float* act; float* obs; float* ag; float* dg; float* rew; unsigned char* term; unsigned char* trunc; };
static int ensure_group_streams(pg_env* e, int k) {
    if (!e->ev_fork) PG_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    for (int g = 0; g < k; g++) {
        if (!e->gstream[g]) PG_CUDA(cudaStreamCreateWithFlags(&e->gstream[g], cudaStreamNonBlocking));
        if (!e->ev_join[g]) PG_CUDA(cudaEventCreateWithFlags(&e->ev_join[g], cudaEventDisableTiming));
    }
    return PG_OK;
}
static int step_impl(pg_env* e, const float* actions, const float* target_quat, float* obs, float* ag, float* dg, float* reward, unsigned char* terminated,
                     unsigned char* truncated, int auto_reset, cudaStream_t stream, const HostIO* hio) {
    PG_CUDA(cudaSetDevice(e->device));
    // Contact-aware scheduling (see perm_*_kernel): large batches are re-sorted by their contact state before every launch, and a
    // step is cut into `segments` launches of consecutive sub-steps so that envs which make contact mid-step are regrouped; small
    // batches keep the identity map, one launch per step and the tiled I/O path.  Results do not depend on either choice.
    const bool use_perm = e->sort_envs && e->n >= 4096;
    const int segs = use_perm ? e->segments : 1;
    const unsigned short* key = e->precision == PG_F32 ? e->Ef.ccount : e->Ed.ccount;
    int* hist = e->precision == PG_F32 ? e->Ef.hist : e->Ed.hist;
    int* perm = e->precision == PG_F32 ? e->Ef.perm : e->Ed.perm;
    EnvDev<float> Ef = e->Ef; EnvDev<double> Ed = e->Ed;
    if (!use_perm) { Ef.perm = nullptr; Ed.perm = nullptr; }
    int groups = use_perm ? e->groups : 1;
    // host mode pipelines the copies against the env groups the configuration already has; PG_HOST_GROUPS forces more (measured: on one
    // GPU the single-stream Reach-joints configuration loses 5 % when cut into 4 groups only to overlap 0.2 ms of copies)
    if (hio && use_perm && e->n >= 16384) { static const int hg = getenv("PG_HOST_GROUPS") ? atoi(getenv("PG_HOST_GROUPS")) : 0; if (hg > groups && hg <= 8) groups = hg; }
    if (groups > 1) { int rc = ensure_group_streams(e, groups); if (rc != PG_OK) return rc; }
    // group g owns the envs (and thread slots) [g * gsize, min(n, (g + 1) * gsize)), gsize a multiple of the sort chunk
    const int gsize = ((e->n + groups - 1) / groups + PERM_CHUNK - 1) / PERM_CHUNK * PERM_CHUNK;
    const size_t A = e->act_dim, O = e->obs_dim, G = e->goal_dim;
    if (groups > 1) PG_CUDA(cudaEventRecord(e->ev_fork, stream));
    for (int g = 0; g < groups; g++) {
        const int t0 = g * gsize, cnt = std::min(e->n, t0 + gsize) - t0;
        if (cnt <= 0) break;
        cudaStream_t st = groups > 1 ? e->gstream[g] : stream;
        if (groups > 1) PG_CUDA(cudaStreamWaitEvent(st, e->ev_fork, 0));
        if (hio) PG_CUDA(cudaMemcpyAsync(const_cast<float*>(actions) + t0 * A, hio->act + t0 * A, (size_t)cnt * A * sizeof(float), cudaMemcpyHostToDevice, st));
        Ef.t0 = Ed.t0 = t0; Ef.tcount = Ed.tcount = cnt;
        const int nchunks = (cnt + PERM_CHUNK - 1) / PERM_CHUNK;
        int* ghist = hist + (size_t)(t0 / PERM_CHUNK + g) * PERM_BUCKETS;
        const int nsub = e->Ef.P.nsub;
        for (int sg = 0; sg < segs; sg++) {
            StepIO io{target_quat, actions, obs, ag, dg, reward, terminated, truncated, auto_reset, sg * nsub / segs, (sg + 1) * nsub / segs};
            if (io.s0 == io.s1) continue;       // fewer sub-steps than segments
            if (use_perm) {
                perm_hist_kernel<<<nchunks, PERM_THREADS, 0, st>>>(key + t0, ghist, cnt);
                perm_scatter_kernel<<<nchunks, PERM_THREADS, 0, st>>>(key + t0, ghist, perm + t0, cnt, nchunks, t0);
                g_launches += 2;
            }
            if (e->precision == PG_F32) Dispatch<float>::step(e->task, Ef, e->ctrl, io, st); else Dispatch<double>::step(e->task, Ed, e->ctrl, io, st);
        }
        if (hio) {
            if (hio->obs) PG_CUDA(cudaMemcpyAsync(hio->obs + t0 * O, obs + t0 * O, (size_t)cnt * O * sizeof(float), cudaMemcpyDeviceToHost, st));
            if (hio->ag) PG_CUDA(cudaMemcpyAsync(hio->ag + t0 * G, ag + t0 * G, (size_t)cnt * G * sizeof(float), cudaMemcpyDeviceToHost, st));
            if (hio->dg) PG_CUDA(cudaMemcpyAsync(hio->dg + t0 * G, dg + t0 * G, (size_t)cnt * G * sizeof(float), cudaMemcpyDeviceToHost, st));
            if (hio->rew) PG_CUDA(cudaMemcpyAsync(hio->rew + t0, reward + t0, (size_t)cnt * sizeof(float), cudaMemcpyDeviceToHost, st));
            if (hio->term) PG_CUDA(cudaMemcpyAsync(hio->term + t0, terminated + t0, (size_t)cnt, cudaMemcpyDeviceToHost, st));
            if (hio->trunc) PG_CUDA(cudaMemcpyAsync(hio->trunc + t0, truncated + t0, (size_t)cnt, cudaMemcpyDeviceToHost, st));
        }
        if (groups > 1) { PG_CUDA(cudaEventRecord(e->ev_join[g], st)); PG_CUDA(cudaStreamWaitEvent(stream, e->ev_join[g], 0)); }
    }
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_step_oriented(pg_env* e, const float* actions, const float* target_quat, float* obs, float* ag, float* dg, float* reward, unsigned char* terminated,
                     unsigned char* truncated, int auto_reset, void* stream) {
    if (!e || !actions) return fail(PG_ERR_ARG, "pg_step: NULL handle or actions");
    if (e->task == PG_TASK_BARE) return fail(PG_ERR_ARG, "pg_step: a bare world is advanced with pg_sim_step");
    if (target_quat && e->ctrl != CTRL_EE) return fail(PG_ERR_ARG, "pg_step_oriented: a target orientation needs ee control");
    return step_impl(e, actions, target_quat, obs, ag, dg, reward, terminated, truncated, auto_reset, (cudaStream_t)stream, nullptr);
}

static int ensure_host_path(pg_env* e) {
    if (e->hstream) return PG_OK;
    const size_t n = (size_t)e->n;
    // one pinned + one device output slab: obs | ag | dg | reward | terminated | truncated (flags padded to floats)
    e->out_floats = n * (size_t)(e->obs_dim + 2 * e->goal_dim + 1) + (2 * n + 3) / 4 + 4;
    e->out_bytes = e->out_floats * sizeof(float);
    PG_CUDA(cudaStreamCreate(&e->hstream));
    PG_CUDA(cudaMallocHost(&e->h_act, n * e->act_dim * sizeof(float)));
    PG_CUDA(cudaMallocHost(&e->h_out, e->out_bytes));
    PG_CUDA(cudaMalloc(&e->d_act, n * e->act_dim * sizeof(float)));
    PG_CUDA(cudaMalloc(&e->d_out, e->out_bytes));
    return PG_OK;
}

// page-locked (cudaHostAlloc / cudaHostRegister / pg_host_pin) host memory can be the source / target of an async copy directly
static bool host_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
int pg_host_pin(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return fail(PG_ERR_ARG, "pg_host_pin: bad argument");
    PG_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return PG_OK;
}
int pg_host_unpin(void* ptr) {
    if (!ptr) return fail(PG_ERR_ARG, "pg_host_unpin: NULL");
    PG_CUDA(cudaHostUnregister(ptr));
    return PG_OK;
}

int pg_step_host(pg_env* e, const float* actions, float* obs, float* ag, float* dg, float* reward, unsigned char* terminated, unsigned char* truncated, int auto_reset) {
    if (!e || !actions) return fail(PG_ERR_ARG, "pg_step_host: NULL handle or actions");
    if (e->task == PG_TASK_BARE) return fail(PG_ERR_ARG, "pg_step_host: a bare world is advanced with pg_sim_step");
    PG_CUDA(cudaSetDevice(e->device));
    int rc = ensure_host_path(e); if (rc != PG_OK) return rc;
    const size_t n = (size_t)e->n, O = e->obs_dim, G = e->goal_dim;
    // pageable buffers go through the handle's pinned staging slabs (one extra host copy each way); pinned ones are copied directly
    const float* src = actions;
    if (!host_pinned(actions)) { memcpy(e->h_act, actions, n * e->act_dim * sizeof(float)); src = e->h_act; }
    float* d_obs = e->d_out; float* d_ag = d_obs + n * O; float* d_dg = d_ag + n * G; float* d_rew = d_dg + n * G;
    unsigned char* d_term = (unsigned char*)(d_rew + n); unsigned char* d_trunc = d_term + n;
    struct Out { void* user; const void* dev; void* host; size_t bytes; bool direct; };
    Out outs[6] = {{obs, d_obs, nullptr, n * O * sizeof(float), false}, {ag, d_ag, nullptr, n * G * sizeof(float), false}, {dg, d_dg, nullptr, n * G * sizeof(float), false},
                   {reward, d_rew, nullptr, n * sizeof(float), false}, {terminated, d_term, nullptr, n, false}, {truncated, d_trunc, nullptr, n, false}};
    for (Out& o : outs) {
        if (!o.user) continue;
        o.direct = host_pinned(o.user);
        o.host = o.direct ? o.user : (void*)((char*)e->h_out + ((const char*)o.dev - (const char*)e->d_out));
    }
    HostIO hio{src, (float*)outs[0].host, (float*)outs[1].host, (float*)outs[2].host, (float*)outs[3].host, (unsigned char*)outs[4].host, (unsigned char*)outs[5].host};
    rc = step_impl(e, e->d_act, nullptr, d_obs, d_ag, d_dg, d_rew, d_term, d_trunc, auto_reset, e->hstream, &hio); if (rc != PG_OK) return rc;
    PG_CUDA(cudaStreamSynchronize(e->hstream));
    for (const Out& o : outs) if (o.user && !o.direct) memcpy(o.user, o.host, o.bytes);
    return PG_OK;
}

int pg_compute_reward_t(int task, int reward_type, double threshold, const void* ag, const void* dg, float* reward, long long m, int dtype, void* stream) {
    if (m == 0) return PG_OK;       // an empty batch is valid (its pointers may be NULL)
    if (task < 0 || task > 5 || reward_type < 0 || reward_type > 1 || !ag || !dg || !reward || m < 0 || !(threshold >= 0)) return fail(PG_ERR_ARG, "pg_compute_reward: bad argument");
    if (dtype == PG_F32) launch_reward<float, true>(task, (const float*)ag, (const float*)dg, reward, nullptr, m, reward_type, threshold, (cudaStream_t)stream);
    else launch_reward<double, true>(task, (const double*)ag, (const double*)dg, reward, nullptr, m, reward_type, threshold, (cudaStream_t)stream);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_is_success_t(int task, double threshold, const void* ag, const void* dg, unsigned char* success, long long m, int dtype, void* stream) {
    if (m == 0) return PG_OK;
    if (task < 0 || task > 5 || !ag || !dg || !success || m < 0 || !(threshold >= 0)) return fail(PG_ERR_ARG, "pg_is_success: bad argument");
    if (dtype == PG_F32) launch_reward<float, false>(task, (const float*)ag, (const float*)dg, nullptr, success, m, 0, threshold, (cudaStream_t)stream);
    else launch_reward<double, false>(task, (const double*)ag, (const double*)dg, nullptr, success, m, 0, threshold, (cudaStream_t)stream);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_compute_reward(int task, int reward_type, const void* ag, const void* dg, float* reward, long long m, int dtype, void* stream) {
    return pg_compute_reward_t(task, reward_type, task >= 0 && task <= 5 ? threshold_f64(task) : 0.0, ag, dg, reward, m, dtype, stream);
}
int pg_is_success(int task, const void* ag, const void* dg, unsigned char* success, long long m, int dtype, void* stream) {
    return pg_is_success_t(task, task >= 0 && task <= 5 ? threshold_f64(task) : 0.0, ag, dg, success, m, dtype, stream);
}
int pg_her_relabel_pitched(int task, int reward_type, double threshold, const void* next_ag, const void* dg, long long pitch, const long long* src, const long long* goal_src,
                           void* dg_out, void* ag_out, float* reward, long long m, int dtype, void* stream) {
    if (m == 0) return PG_OK;
    if (task < 0 || task > 5 || reward_type < 0 || reward_type > 1 || !next_ag || !dg || !src || !goal_src || !dg_out || !reward || m < 0 || !(threshold >= 0))
        return fail(PG_ERR_ARG, "pg_her_relabel: bad argument");
    if (pitch < task_goal_dim(task)) return fail(PG_ERR_ARG, "pg_her_relabel_pitched: the row pitch is smaller than the goal dimension");
    if (dtype == PG_F32) launch_her<float>(task, (const float*)next_ag, (const float*)dg, src, goal_src, (float*)dg_out, (float*)ag_out, reward, m, pitch, reward_type, threshold, (cudaStream_t)stream);
    else launch_her<double>(task, (const double*)next_ag, (const double*)dg, src, goal_src, (double*)dg_out, (double*)ag_out, reward, m, pitch, reward_type, threshold, (cudaStream_t)stream);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_her_relabel_t(int task, int reward_type, double threshold, const void* next_ag, const void* dg, const long long* src, const long long* goal_src, void* dg_out, void* ag_out,
                     float* reward, long long m, int dtype, void* stream) {
    return pg_her_relabel_pitched(task, reward_type, threshold, next_ag, dg, task >= 0 && task <= 5 ? task_goal_dim(task) : 0, src, goal_src, dg_out, ag_out, reward, m, dtype, stream);
}
int pg_her_relabel(int task, int reward_type, const void* next_ag, const void* dg, const long long* src, const long long* goal_src, void* dg_out, void* ag_out,
                   float* reward, long long m, int dtype, void* stream) {
    return pg_her_relabel_t(task, reward_type, task >= 0 && task <= 5 ? threshold_f64(task) : 0.0, next_ag, dg, src, goal_src, dg_out, ag_out, reward, m, dtype, stream);
}
// Host-array entry points (the path stable-baselines3's HerReplayBuffer takes through env_method("compute_reward"), reference
// panda_gym/envs/core.py:226, examples/train_push.py:1-12).  Staging is cached per device and only ever grows: after the first call
// of a given size there is no allocation, one private stream carries H2D -> kernel -> D2H, one stream synchronisation at the end.
struct HostStage { cudaStream_t st = nullptr; void *da = nullptr, *db = nullptr, *dr = nullptr, *hr = nullptr; size_t cap_in = 0, cap_out = 0; };
static HostStage g_stage[64];
static long long g_stage_allocs = 0;
static int stage_for(int device, size_t in_bytes, size_t out_bytes, HostStage** out) {
    if (device < 0 || device >= 64) return fail(PG_ERR_ARG, "host path: bad device index");
    int rc = check_device(device, "host path"); if (rc != PG_OK) return rc;
    HostStage& h = g_stage[device];
    if (!h.st) PG_CUDA(cudaStreamCreateWithFlags(&h.st, cudaStreamNonBlocking));
    if (in_bytes > h.cap_in) {
        PG_CUDA(cudaStreamSynchronize(h.st));
        if (h.da) cudaFree(h.da); if (h.db) cudaFree(h.db); h.da = h.db = nullptr; h.cap_in = 0;
        const size_t cap = in_bytes + in_bytes / 4;
        PG_CUDA(cudaMalloc(&h.da, cap)); PG_CUDA(cudaMalloc(&h.db, cap)); h.cap_in = cap; g_stage_allocs++;
    }
    if (out_bytes > h.cap_out) {
        PG_CUDA(cudaStreamSynchronize(h.st));
        if (h.dr) cudaFree(h.dr); if (h.hr) cudaFreeHost(h.hr); h.dr = h.hr = nullptr; h.cap_out = 0;
        const size_t cap = out_bytes + out_bytes / 4;
        PG_CUDA(cudaMalloc(&h.dr, cap)); PG_CUDA(cudaMallocHost(&h.hr, cap)); h.cap_out = cap; g_stage_allocs++;
    }
    *out = &h;
    return PG_OK;
}
long long pg_host_stage_allocations(void) { return g_stage_allocs; }
static int reward_host(bool want_reward, int task, int reward_type, double thr, const void* ag, const void* dg, void* result, long long m, int dtype, int device) {
    const size_t es = dtype == PG_F32 ? 4 : 8, bytes = (size_t)m * task_goal_dim(task) * es, out_bytes = (size_t)m * (want_reward ? 4 : 1);
    HostStage* h = nullptr;
    int rc = stage_for(device, bytes, out_bytes, &h); if (rc != PG_OK) return rc;
    PG_CUDA(cudaMemcpyAsync(h->da, ag, bytes, cudaMemcpyHostToDevice, h->st));
    PG_CUDA(cudaMemcpyAsync(h->db, dg, bytes, cudaMemcpyHostToDevice, h->st));
    rc = want_reward ? pg_compute_reward_t(task, reward_type, thr, h->da, h->db, (float*)h->dr, m, dtype, h->st) : pg_is_success_t(task, thr, h->da, h->db, (unsigned char*)h->dr, m, dtype, h->st);
    if (rc != PG_OK) return rc;
    const bool direct = host_pinned(result);
    PG_CUDA(cudaMemcpyAsync(direct ? result : h->hr, h->dr, out_bytes, cudaMemcpyDeviceToHost, h->st));
    PG_CUDA(cudaStreamSynchronize(h->st));
    if (!direct) memcpy(result, h->hr, out_bytes);
    return PG_OK;
}
int pg_compute_reward_host_t(int task, int reward_type, double threshold, const void* ag, const void* dg, float* reward, long long m, int dtype, int device) {
    if (m == 0) return PG_OK;
    if (task < 0 || task > 5 || reward_type < 0 || reward_type > 1 || !ag || !dg || !reward || m < 0 || !(threshold >= 0)) return fail(PG_ERR_ARG, "pg_compute_reward_host: bad argument");
    return reward_host(true, task, reward_type, threshold, ag, dg, reward, m, dtype, device);
}
int pg_is_success_host_t(int task, double threshold, const void* ag, const void* dg, unsigned char* success, long long m, int dtype, int device) {
    if (m == 0) return PG_OK;
    if (task < 0 || task > 5 || !ag || !dg || !success || m < 0 || !(threshold >= 0)) return fail(PG_ERR_ARG, "pg_is_success_host: bad argument");
    return reward_host(false, task, 0, threshold, ag, dg, success, m, dtype, device);
}
int pg_compute_reward_host(int task, int reward_type, const void* ag, const void* dg, float* reward, long long m, int dtype, int device) {
    return pg_compute_reward_host_t(task, reward_type, task >= 0 && task <= 5 ? threshold_f64(task) : 0.0, ag, dg, reward, m, dtype, device);
}
int pg_is_success_host(int task, const void* ag, const void* dg, unsigned char* success, long long m, int dtype, int device) {
    return pg_is_success_host_t(task, task >= 0 && task <= 5 ? threshold_f64(task) : 0.0, ag, dg, success, m, dtype, device);
}

// Snapshots are stream-ordered: one device-to-device copy of the handle's single SoA allocation, enqueued on the caller's stream
// (the env groups' streams are joined back to that stream by every pg_step, so ordering on it is sufficient).  No device-wide
// synchronisation, no allocation when a removed snapshot's buffer can be reused -> a save / K x (restore, step) / remove loop can
// be captured in a CUDA graph after its first, warm-up execution.
int pg_save_state_async(pg_env* e, int* state_id, void* stream) {
    if (!e || !state_id) return fail(PG_ERR_ARG, "pg_save_state: NULL argument");
    PG_CUDA(cudaSetDevice(e->device));
    void* p = nullptr;
    if (!e->snap_pool.empty()) { p = e->snap_pool.back(); e->snap_pool.pop_back(); }
    else PG_CUDA(cudaMalloc(&p, e->blob_bytes));
    cudaError_t ce = cudaMemcpyAsync(p, e->blob, e->blob_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (ce != cudaSuccess) { e->snap_pool.push_back(p); return cuda_fail(ce, "cudaMemcpyAsync(snapshot)"); }
    *state_id = e->next_snap++; e->snaps[*state_id] = p;
    return PG_OK;
}
int pg_restore_state_async(pg_env* e, int state_id, void* stream) {
    if (!e) return fail(PG_ERR_ARG, "pg_restore_state: NULL handle");
    auto it = e->snaps.find(state_id);
    if (it == e->snaps.end()) return fail(PG_ERR_STATE, "pg_restore_state: unknown state id " + std::to_string(state_id));
    PG_CUDA(cudaSetDevice(e->device));
    PG_CUDA(cudaMemcpyAsync(e->blob, it->second, e->blob_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return PG_OK;
}
// blocking forms on the legacy default stream (which orders against every blocking stream of the device)
int pg_save_state(pg_env* e, int* state_id) {
    int rc = pg_save_state_async(e, state_id, nullptr); if (rc != PG_OK) return rc;
    PG_CUDA(cudaStreamSynchronize(nullptr));
    return PG_OK;
}
int pg_restore_state(pg_env* e, int state_id) {
    int rc = pg_restore_state_async(e, state_id, nullptr); if (rc != PG_OK) return rc;
    PG_CUDA(cudaStreamSynchronize(nullptr));
    return PG_OK;
}
// The buffer goes back to the handle's pool: copies that still read or write it are ordered before any later reuse as long as the
// caller keeps using the same stream for the handle (the contract of every entry point).
int pg_remove_state(pg_env* e, int state_id) {
    if (!e) return fail(PG_ERR_ARG, "pg_remove_state: NULL handle");
    auto it = e->snaps.find(state_id);
    if (it == e->snaps.end()) return fail(PG_ERR_STATE, "pg_remove_state: unknown state id " + std::to_string(state_id));
    e->snap_pool.push_back(it->second); e->snaps.erase(it);
    return PG_OK;
}

int pg_get_state(pg_env* e, double* state, void* stream) {
    if (!e || !state) return fail(PG_ERR_ARG, "pg_get_state: NULL argument");
    PG_CUDA(cudaSetDevice(e->device));
    const int grid = (e->n + BLOCK - 1) / BLOCK;
    if (e->precision == PG_F32) get_state_kernel<float><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ef, e->nobj, e->goal_dim, state);
    else get_state_kernel<double><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ed, e->nobj, e->goal_dim, state);
    g_launches++;
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_set_state(pg_env* e, const double* state, const unsigned char* mask, void* stream) {
    if (!e || !state) return fail(PG_ERR_ARG, "pg_set_state: NULL argument");
    PG_CUDA(cudaSetDevice(e->device));
    const int grid = (e->n + BLOCK - 1) / BLOCK;
    if (e->precision == PG_F32) set_state_kernel<float><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ef, e->nobj, e->goal_dim, state, mask);
    else set_state_kernel<double><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ed, e->nobj, e->goal_dim, state, mask);
    g_launches++;
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
static int launch_ik(pg_env* e, int link, const double* position, const double* orientation, double* joint_angles, int out7, void* stream) {
    PG_CUDA(cudaSetDevice(e->device));
    const int grid = (e->n + BLOCK - 1) / BLOCK;
    if (e->precision == PG_F32) ik_kernel<float><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ef, link, position, orientation, joint_angles, out7);
    else ik_kernel<double><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ed, link, position, orientation, joint_angles, out7);
    g_launches++;
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_inverse_kinematics(pg_env* e, const double* position, const double* orientation, double* joint_angles, void* stream) {
    if (!e || !position || !orientation || !joint_angles) return fail(PG_ERR_ARG, "pg_inverse_kinematics: NULL argument");
    return launch_ik(e, 11, position, orientation, joint_angles, 1, stream);
}
int pg_inverse_kinematics_link(pg_env* e, int link, const double* position, const double* orientation, double* joint_angles, void* stream) {
    if (!e || !position || !orientation || !joint_angles) return fail(PG_ERR_ARG, "pg_inverse_kinematics_link: NULL argument");
    if (link < 0 || link > 11) return fail(PG_ERR_ARG, "pg_inverse_kinematics_link: link must be 0..11");
    return launch_ik(e, link, position, orientation, joint_angles, 0, stream);
}
int pg_get_link_state(pg_env* e, int link, double* out, void* stream) {
    if (!e || !out) return fail(PG_ERR_ARG, "pg_get_link_state: NULL argument");
    if (link < 0 || link > 11) return fail(PG_ERR_ARG, "pg_get_link_state: link must be 0..11");
    PG_CUDA(cudaSetDevice(e->device));
    const int grid = (e->n + BLOCK - 1) / BLOCK;
    if (e->precision == PG_F32) link_state_kernel<float><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ef, link, out);
    else link_state_kernel<double><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ed, link, out);
    g_launches++;
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_set_motors(pg_env* e, const double* motors, const unsigned char* mask, void* stream) {
    if (!e || !motors) return fail(PG_ERR_ARG, "pg_set_motors: NULL argument");
    if (e->task != PG_TASK_BARE) return fail(PG_ERR_ARG, "pg_set_motors: task handles derive their motor targets from the action inside pg_step (Panda.set_action); raw motors exist on bare worlds");
    PG_CUDA(cudaSetDevice(e->device));
    const int grid = (e->n + BLOCK - 1) / BLOCK;
    if (e->precision == PG_F32) set_motors_kernel<float><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ef, motors, mask);
    else set_motors_kernel<double><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ed, motors, mask);
    g_launches++;
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_get_motors(pg_env* e, double* motors, void* stream) {
    if (!e || !motors) return fail(PG_ERR_ARG, "pg_get_motors: NULL argument");
    if (e->task != PG_TASK_BARE) return fail(PG_ERR_ARG, "pg_get_motors: bare worlds only");
    PG_CUDA(cudaSetDevice(e->device));
    const int grid = (e->n + BLOCK - 1) / BLOCK;
    if (e->precision == PG_F32) get_motors_kernel<float><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ef, motors);
    else get_motors_kernel<double><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ed, motors);
    g_launches++;
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_sim_step(pg_env* e, int n_substeps, void* stream) {
    if (!e || n_substeps < 0) return fail(PG_ERR_ARG, "pg_sim_step: bad argument");
    if (e->task != PG_TASK_BARE) return fail(PG_ERR_ARG, "pg_sim_step: task handles step through pg_step (controller + sub-steps + observation in one call)");
    if (n_substeps == 0) return PG_OK;
    PG_CUDA(cudaSetDevice(e->device));
    if (e->precision == PG_F32) launch_bare_step<float>(e->Ef, e->nobj, n_substeps, (cudaStream_t)stream);
    else launch_bare_step<double>(e->Ed, e->nobj, n_substeps, (cudaStream_t)stream);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
static void fill_render_scene(pg_env* e) {
    RenderScene& R = e->rscene;
    memset(&R, 0, sizeof R);
    const Scene<double>& S = e->precision == PG_F32 ? scene_cast<double>(e->Ef.S) : e->Ed.S;
    R.nobj = e->nobj;
    R.has_table = S.table_x0 < S.table_x1; R.table[0] = (float)S.table_x0; R.table[1] = (float)S.table_x1; R.table[2] = (float)S.table_y0; R.table[3] = (float)S.table_y1; R.table_height = 0.4f;
    R.has_plane = S.ground_z > -1e29; R.plane_z = (float)S.ground_z;
    const double bz = e->precision == PG_F32 ? (double)e->Ef.M.base[2] : e->Ed.M.base[2];
    R.has_robot = bz < 500.0;                                  // a bare world without a robot parks it 1 km up
    const float bc[3] = {-0.04f, 0.0f, 0.07f}, bh[3] = {0.11f, 0.1f, 0.07f};     // panda_link0: approximate extents (its mesh is not available)
    for (int k = 0; k < 3; k++) { R.base_c[k] = bc[k]; R.base_h[k] = bh[k]; }
    for (int l = 0; l < 7; l++) for (int k = 0; k < 3; k++) { R.link_c[l][k] = (float)kLinks[l].com[k]; R.link_h[l][k] = (float)(0.5 * kLinks[l].box[k]); }
    const int rb_link[3] = {8, 9, 10};
    for (int b = 0; b < 3; b++) for (int k = 0; k < 3; k++) { R.link_c[rb_link[b]][k] = (float)S.rb_c[b][k]; R.link_h[rb_link[b]][k] = (float)S.rb_h[b][k]; }
}
int pg_render(pg_env* e, int width, int height, const double* camera, int crop, float* depth, unsigned char* rgba, unsigned char* segmentation, float* points,
              unsigned char* valid, void* stream) {
    if (!e || !camera || width <= 0 || height <= 0 || width > 4096 || height > 4096) return fail(PG_ERR_ARG, "pg_render: bad argument");
    if (valid && !points) return fail(PG_ERR_ARG, "pg_render: valid is the point cloud's mask, it needs points");
    PG_CUDA(cudaSetDevice(e->device));
    if (!e->prims) { PG_CUDA(cudaMalloc(&e->prims, (size_t)e->n * RENDER_MAX_PRIMS * sizeof(RenderPrim))); fill_render_scene(e); }
    // b3ComputeViewMatrixFromYawPitchRoll (up axis 2) + computeProjectionMatrixFOV(fov 60, near 0.1, far 100): pybullet.py:90-102
    const double rad = 0.01745329251994329547, yaw = camera[4] * rad, pitch = camera[5] * rad, roll = camera[6] * rad, dist = camera[3];
    const double cy = cos(yaw), sy = sin(yaw), cr = cos(roll), sr = sin(roll), cp = cos(pitch), sp = sin(pitch);
    // eyeRot.setEulerZYX(yaw, roll, pitch): R = Rz(yaw) Ry(roll) Rx(pitch)
    const double Rm[3][3] = {{cy * cr, cy * sr * sp - sy * cp, cy * sr * cp + sy * sp}, {sy * cr, sy * sr * sp + cy * cp, sy * sr * cp - cy * sp}, {-sr, cr * sp, cr * cp}};
    double eye[3], up[3], f[3], sv[3], u[3];
    for (int k = 0; k < 3; k++) { eye[k] = Rm[k][1] * -dist + camera[k]; up[k] = Rm[k][2]; f[k] = camera[k] - eye[k]; }
    double fl = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    if (!(fl > 0)) return fail(PG_ERR_ARG, "pg_render: camera distance must be > 0");
    for (int k = 0; k < 3; k++) f[k] /= fl;
    sv[0] = f[1] * up[2] - f[2] * up[1]; sv[1] = f[2] * up[0] - f[0] * up[2]; sv[2] = f[0] * up[1] - f[1] * up[0];
    double sl = sqrt(sv[0] * sv[0] + sv[1] * sv[1] + sv[2] * sv[2]);
    for (int k = 0; k < 3; k++) sv[k] /= sl;
    u[0] = sv[1] * f[2] - sv[2] * f[1]; u[1] = sv[2] * f[0] - sv[0] * f[2]; u[2] = sv[0] * f[1] - sv[1] * f[0];
    RenderCamera C; memset(&C, 0, sizeof C);
    C.width = width; C.height = height; C.crop = crop;
    for (int k = 0; k < 3; k++) { C.eye[k] = (float)eye[k]; C.fwd[k] = (float)f[k]; C.right[k] = (float)sv[k]; C.up[k] = (float)u[k]; }
    C.tan_half_fov = (float)tan(30.0 * rad); C.aspect = (float)width / (float)height; C.near = 0.1f; C.far = 100.0f;
    const double l[3] = {0.35, -0.25, 0.9}; const double ll = sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    for (int k = 0; k < 3; k++) C.light[k] = (float)(l[k] / ll);
    const unsigned char col[RENDER_ID_ROBOT + 1][4] = {{0, 0, 0, 255}, {38, 38, 38, 255}, {242, 242, 242, 255},
                                                      {(unsigned char)(e->task == PG_TASK_STACK ? 26 : (e->task == PG_TASK_FLIP ? 255 : 26)), (unsigned char)(e->task == PG_TASK_STACK ? 26 : (e->task == PG_TASK_FLIP ? 255 : 230)),
                                                       (unsigned char)(e->task == PG_TASK_STACK ? 230 : (e->task == PG_TASK_FLIP ? 255 : 26)), 255},
                                                      {26, 230, 26, 255}, {235, 235, 235, 255}};     // plane 0.15, table 0.95, objects (tasks/ *.py rgba_color), robot
    memcpy(C.color, col, sizeof col);
    C.background[0] = 223; C.background[1] = 54; C.background[2] = 45; C.background[3] = 255;          // pybullet.py:28 default background_color
    if (e->precision == PG_F32) launch_render_setup<float>(e->Ef, e->rscene, e->prims, (cudaStream_t)stream); else launch_render_setup<double>(e->Ed, e->rscene, e->prims, (cudaStream_t)stream);
    launch_render(e->prims, C, e->n, depth, rgba, segmentation, points, valid, (cudaStream_t)stream);
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_get_ee_pose(pg_env* e, double* pose, void* stream) {
    if (!e || !pose) return fail(PG_ERR_ARG, "pg_get_ee_pose: NULL argument");
    PG_CUDA(cudaSetDevice(e->device));
    const int grid = (e->n + BLOCK - 1) / BLOCK;
    if (e->precision == PG_F32) ee_pose_kernel<float><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ef, pose);
    else ee_pose_kernel<double><<<grid, BLOCK, 0, (cudaStream_t)stream>>>(e->Ed, pose);
    g_launches++;
    PG_CUDA(cudaGetLastError());
    return PG_OK;
}
int pg_debug_schedule(pg_env* e, unsigned short* key_host, int* perm_host) {
    if (!e || !key_host || !perm_host) return fail(PG_ERR_ARG, "pg_debug_schedule: NULL argument");
    PG_CUDA(cudaSetDevice(e->device));
    PG_CUDA(cudaMemcpy(key_host, e->precision == PG_F32 ? e->Ef.ccount : e->Ed.ccount, (size_t)e->n * sizeof(unsigned short), cudaMemcpyDeviceToHost));
    PG_CUDA(cudaMemcpy(perm_host, e->precision == PG_F32 ? e->Ef.perm : e->Ed.perm, (size_t)e->n * sizeof(int), cudaMemcpyDeviceToHost));
    return PG_OK;
}
int pg_debug_timing(pg_env* e, long long* out) {
    if (!e || !out) return fail(PG_ERR_ARG, "pg_debug_timing: NULL argument");
    if (!e->dbg) return fail(PG_ERR_ARG, "pg_debug_timing: create the handle with PG_DEBUG_TIMING=1 in the environment");
    PG_CUDA(cudaSetDevice(e->device));
    PG_CUDA(cudaDeviceSynchronize());
    PG_CUDA(cudaMemcpy(out, e->dbg, (size_t)e->n * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    return PG_OK;
}
int pg_contact_overflows(pg_env* e, long long* count) {
    if (!e || !count) return fail(PG_ERR_ARG, "pg_contact_overflows: NULL argument");
    PG_CUDA(cudaSetDevice(e->device));
    double d = 0;
    PG_CUDA(cudaMemcpy(&d, (e->precision == PG_F32 ? e->Ef.stats : e->Ed.stats) + 5, sizeof(double), cudaMemcpyDeviceToHost));
    *count = (long long)d;
    return PG_OK;
}
int pg_diverged(pg_env* e, long long* count) {
    if (!e || !count) return fail(PG_ERR_ARG, "pg_diverged: NULL argument");
    PG_CUDA(cudaSetDevice(e->device));
    double d = 0;
    PG_CUDA(cudaMemcpy(&d, (e->precision == PG_F32 ? e->Ef.stats : e->Ed.stats) + 4, sizeof(double), cudaMemcpyDeviceToHost));
    *count = (long long)d;
    return PG_OK;
}
int pg_stats(pg_env* e, double out[4]) {
    if (!e || !out) return fail(PG_ERR_ARG, "pg_stats: NULL argument");
    PG_CUDA(cudaSetDevice(e->device));
    PG_CUDA(cudaMemcpy(out, e->precision == PG_F32 ? e->Ef.stats : e->Ed.stats, 4 * sizeof(double), cudaMemcpyDeviceToHost));
    return PG_OK;
}

}  // extern "C"
