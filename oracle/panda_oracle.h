/*
 * panda_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C, fp64 restatement of what the reference's hot path computes:
 *   reference /root/reference/panda_gym/envs/core.py:280-289 (RobotTaskEnv.step)
 *   -> envs/robots/panda.py:52-119 (set_action / get_obs)
 *   -> pybullet.py:52-55 (20 x stepSimulation), :462-497 (motors, IK), :284-425 (getters)
 *   -> envs/tasks/ (obs / achieved goal / is_success / compute_reward).
 * The arithmetic itself lives in pybullet==3.2.5 (env.yml:107), which is NOT in
 * /root/reference and not installable here; the engine part below restates the
 * published Bullet algorithms (Featherstone ABA multibody, btMultiBodyJointMotor,
 * btMultiBodyJointLimitConstraint, sequential-impulse PGS, BussIK DLS) as described
 * in SURVEY.md App. B/C, and is pinned by the reference's seven known-answer tests
 * in test/pybullet_test.py (:34,:64,:135,:152,:169,:186,:203,:265) -- see
 * tests/test_oracle_kat.py.  Contact dynamics have no reference golden vectors:
 * "parity unpinned" for everything involving contact.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libpanda_b200.so) never does.
 */
#ifndef PANDA_ORACLE_H
#define PANDA_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { PO_REACH = 0, PO_PUSH = 1, PO_SLIDE = 2, PO_PICK_AND_PLACE = 3, PO_STACK = 4, PO_FLIP = 5, PO_BARE = 6 };
enum { PO_CTRL_EE = 0, PO_CTRL_JOINTS = 1 };
enum { PO_REWARD_SPARSE = 0, PO_REWARD_DENSE = 1 };

typedef struct PoSim PoSim;

/* ---- L1: sim facade (mirrors panda_gym/pybullet.py) ---- */
PoSim *po_create(int task, double base_x, double base_y, double base_z);
void po_destroy(PoSim *s);
void po_set_static(PoSim *s, const double *table_rect, const double *ground_z); /* pybullet.py:726-771 create_plane / create_table; NULL = absent */
void po_set_joint_state(PoSim *s, const double *q, const double *qd, const double *qc);
void po_set_object_shape(PoSim *s, int obj, int shape, double hx, double hy, double hz, double mass, double mu);
int po_num_objects(const PoSim *s);
int po_add_box(PoSim *s, double hx, double hy, double hz, double mass, const double pos[3]); /* pybullet.py:531-582 create_box */
void po_step(PoSim *s, int n_substeps);                        /* pybullet.py:52-55 */
void po_reset_joint(PoSim *s, int link, double angle);         /* pybullet.py:451-460 */
void po_get_joint(const PoSim *s, int link, double *q, double *qd); /* pybullet.py:402-425 */
/* POSITION_CONTROL motor on one joint (pybullet.py:462-477); kp/kd pybullet defaults 0.1 / 1.0 */
void po_control_joint(PoSim *s, int link, double target, double max_force);
/* getLinkState (cached link frames): pos/quat of the CoM frame, world lin/ang velocity */
void po_get_link_state(const PoSim *s, int link, double pos[3], double quat[4], double lin[3], double ang[3]);
void po_inverse_kinematics(const PoSim *s, int link, const double pos[3], const double quat[4], double out[9]); /* pybullet.py:479-497 */
void po_set_base_pose(PoSim *s, int obj, const double pos[3], const double quat[4]); /* pybullet.py:427-439 */
void po_get_base_pose(const PoSim *s, int obj, double pos[3], double quat[4]);
void po_get_base_velocity(const PoSim *s, int obj, double lin[3], double ang[3]);
void po_set_base_velocity(PoSim *s, int obj, const double lin[3], const double ang[3]);
void po_euler_from_quat(const double q[4], double e[3]);       /* pybullet.py:308-325 */
int po_state_size(void);
void po_save_state(const PoSim *s, double *buf);               /* pybullet.py:61-68 */
void po_restore_state(PoSim *s, const double *buf);
/* debug / calibration hooks */
void po_set_link_inertia(int link, double ixx, double iyy, double izz);
void po_get_link_inertia(int link, double out[3]);
int po_last_num_contacts(const PoSim *s);
int po_last_iterations(const PoSim *s);
int po_last_active_arm_limits(const PoSim *s);
void po_mass_matrix(PoSim *s, double Minv[81]); /* inverse joint-space inertia at the current q */
void po_get_link_def(int link, double out[16]);
void po_get_robot_box(int i, double out[8]);

/* ---- L2/L3: env (mirrors RobotTaskEnv + Panda + Task) ---- */
typedef struct PoEnv PoEnv;
PoEnv *po_env_create(int task, int control_type, int reward_type);
void po_env_destroy(PoEnv *e);
void po_env_set_params(PoEnv *e, int n_substeps, double distance_threshold);
PoSim *po_env_sim(PoEnv *e);
int po_env_obs_dim(const PoEnv *e);
int po_env_goal_dim(const PoEnv *e);
int po_env_action_dim(const PoEnv *e);
/* reset to the neutral pose; goal[G]; objpos = 3 doubles per object (identity orientation) */
void po_env_reset(PoEnv *e, const double *goal, const double *objpos, float *obs, float *ag, float *dg);
void po_env_set_state(PoEnv *e, const double *q, const double *qd);
void po_env_get_state(PoEnv *e, double *q, double *qd);
void po_env_step_oriented(PoEnv *e, const float *action, const double *target_quat, double ee_scale, double finger_scale, float *obs, float *ag, float *dg, float *reward, unsigned char *terminated);
void po_env_step(PoEnv *e, const float *action, float *obs, float *ag, float *dg, float *reward, unsigned char *terminated);

/* CPU baseline driver: random-action rollout of one env for n_steps env steps (reset on success / TimeLimit) */
void po_env_step_batch(PoEnv **envs, int n, int na, int no, int ng, const float *actions, float *obs, float *ag, float *dg, float *reward, unsigned char *terminated);
void po_env_set_full_state(PoEnv *e, const double *st);
void po_env_get_full_state(PoEnv *e, double *st);
void *po_bench_open(int task, int control, unsigned long long seed);
double po_bench_steps(void *h, int n_steps);
void po_bench_close(void *h);
double po_bench_run(int task, int control, int n_steps, unsigned long long seed);

/* ---- rewards (utils.py:4-30; tasks/ is_success / compute_reward) ---- */
void po_compute_reward_f32(int task, int reward_type, const float *ag, const float *dg, float *out, long n);
void po_is_success_f32(int task, const float *ag, const float *dg, unsigned char *out, long n);
void po_compute_reward_f64(int task, int reward_type, const double *ag, const double *dg, float *out, long n);
void po_is_success_f64(int task, const double *ag, const double *dg, unsigned char *out, long n);

#ifdef __cplusplus
}
#endif
#endif
