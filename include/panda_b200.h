/*
 * panda_b200.h -- C ABI of libpanda_b200.so: batched Panda manipulation environments on one B200.
 *
 * The reference has no FFI of its own on this path: its step is Python calling the pybullet C extension
 * (reference panda_gym/pybullet.py:34-55).  This library is what a maintainer binds instead; each entry point names
 * the reference interface it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions: every function returns 0 on success and a negative code on error (text via pg_last_error()).
 * Unless a name ends in _host, every pointer is a DEVICE pointer owned by the caller, arrays are row-major
 * [num_envs, dim], and the work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*, NULL = the legacy
 * default stream).  A handle owns its SoA state in HBM, is bound to one device and is not thread-safe.
 * There is no CPU fallback: without a CUDA device pg_create fails with PG_ERR_CUDA.
 */
#ifndef PANDA_B200_H
#define PANDA_B200_H
#ifdef __cplusplus
extern "C" {
#endif

enum { PG_TASK_REACH = 0, PG_TASK_PUSH = 1, PG_TASK_SLIDE = 2, PG_TASK_PICK_AND_PLACE = 3, PG_TASK_STACK = 4, PG_TASK_FLIP = 5 };
enum { PG_CTRL_EE = 0, PG_CTRL_JOINTS = 1 };          /* panda_gym/envs/robots/panda.py:21-33 control_type */
enum { PG_REWARD_SPARSE = 0, PG_REWARD_DENSE = 1 };   /* panda_gym/envs/tasks/reach.py:60-65 reward_type */
enum { PG_F32 = 0, PG_F64 = 1 };
enum { PG_OK = 0, PG_ERR_ARG = -1, PG_ERR_CUDA = -2, PG_ERR_STATE = -3 };

typedef struct pg_env pg_env;

/* Replaces PandaXxxEnv.__init__ (panda_gym/envs/panda_tasks.py:14-113: PyBullet + Panda at (-0.6,0,0) + task scene) for
 * num_envs environments.  precision: PG_F32 (product path) or PG_F64 (parity debugging).  seed + env_id_offset key the
 * device RNG by GLOBAL environment index, so a sharded run reproduces the unsharded one.  All envs start reset. */
int pg_create(int task, int control_type, int reward_type, int num_envs, int device, unsigned long long seed,
              long long env_id_offset, int precision, pg_env** out);
int pg_destroy(pg_env* env);                                            /* RobotTaskEnv.close, core.py:291-292 */
/* observation / goal / action widths and the TimeLimit length (panda_gym/__init__.py:18,46) */
int pg_dims(const pg_env* env, int* obs_dim, int* goal_dim, int* action_dim, int* max_episode_steps, int* state_dim);

/* RobotTaskEnv.reset (core.py:240-250) for the envs whose mask byte is non-zero (NULL = all): neutral joints, zero
 * velocities, goal and object placement either sampled on the device (distributions of tasks/<task>.py _sample_goal /
 * _sample_object) or taken from goal_override [N,G] / object_override [N,3*n_objects] (float64; NULL = sample).
 * Writes obs/ag/dg rows of the reset envs only (any of them may be NULL). */
int pg_reset(pg_env* env, const unsigned char* mask, const double* goal_override, const double* object_override,
             float* obs, float* achieved_goal, float* desired_goal, void* stream);

/* RobotTaskEnv.step (core.py:280-289) for every env: Panda.set_action (panda.py:52-70), 20 x stepSimulation
 * (pybullet.py:52-55), _get_obs (core.py:229-238), is_success, compute_reward; truncated = TimeLimit.
 * auto_reset != 0: envs that terminated or truncated are reset in the same call (device-sampled goal/object) and their
 * obs/ag/dg rows hold the reset observation while reward/terminated/truncated describe the finished step.
 * Ordering: everything the call enqueues is ordered after the work already on `stream` and completes before work enqueued on
 * `stream` afterwards.  Batches of >= 4096 envs are advanced by several launches (a step cut into sub-step segments, each after a
 * re-sort of the thread -> env map) and, from 16384 envs, in env groups on streams the handle owns, forked from and joined back
 * to `stream` with events (legal under stream capture).  None of this changes a result.  Environment knobs for A/B runs:
 * PG_SORT_ENVS=0, PG_SEGMENTS=<divisor of 20>, PG_GROUPS=<1..8>. */
int pg_step(pg_env* env, const float* actions, float* obs, float* achieved_goal, float* desired_goal, float* reward,
            unsigned char* terminated, unsigned char* truncated, int auto_reset, void* stream);
/* The fork's orientation-target control (panda_gym/envs/robots/panda_ori.py:52-99: set_action(action, euler_xyz)): as pg_step, with a
 * per-env EE target quaternion [N,4] (x,y,z,w; float32, normalised inside) fed to the IK instead of (1,0,0,0).  ee control only. */
int pg_step_oriented(pg_env* env, const float* actions, const float* target_quat, float* obs, float* achieved_goal, float* desired_goal,
                     float* reward, unsigned char* terminated, unsigned char* truncated, int auto_reset, void* stream);
/* Action scaling of Panda.set_action: defaults 0.05 (ee / joint displacement, panda.py:81,103) and 0.2 (fingers, panda.py:65);
 * the fork's panda_cartesian.py:67,157 uses 1.0 / 1.0. */
int pg_set_action_scale(pg_env* env, double ee_scale, double finger_scale);
/* Same call with HOST buffers: actions are copied to the device, the step runs, all outputs are copied back and the
 * stream is synchronised before returning (the end-to-end path a CPU learner uses). */
int pg_step_host(pg_env* env, const float* actions, float* obs, float* achieved_goal, float* desired_goal, float* reward,
                 unsigned char* terminated, unsigned char* truncated, int auto_reset);

/* Page-lock a caller-owned host buffer (cudaHostRegister) so that pg_step_host copies to / from it directly instead of through the
 * handle's staging slabs; unpin before freeing the buffer.  Buffers that are not pinned work too (one extra host copy). */
int pg_host_pin(void* ptr, size_t bytes);
int pg_host_unpin(void* ptr);

/* Task.compute_reward / Task.is_success (tasks/<task>.py, utils.py:4-30) on M rows of achieved/desired goals
 * (dtype PG_F32 or PG_F64 inputs; float32 rewards, uint8 success) -- the HER relabelling entry point. */
int pg_compute_reward(int task, int reward_type, const void* achieved_goal, const void* desired_goal, float* reward,
                      long long m, int dtype, void* stream);
int pg_is_success(int task, const void* achieved_goal, const void* desired_goal, unsigned char* success, long long m,
                  int dtype, void* stream);
/* HER relabelling fused with compute_reward (replaces the gather + env.compute_reward round trip of stable-baselines3's
 * HerReplayBuffer that reference examples/train_push.py:1-12 sets up; rewards as tasks/<task>.py compute_reward).  next_achieved_goal and
 * desired_goal are the replay buffer's [R, G] goal arrays (device); for each of the M sampled transitions j: the new goal is
 * next_achieved_goal[goal_src[j]] (desired_goal[src[j]] when goal_src[j] < 0) and the reward is compute_reward(next_achieved_goal[src[j]],
 * new goal).  Outputs: desired_goal_out [M, G], achieved_goal_out [M, G] (may be NULL), reward [M] float32. */
int pg_her_relabel(int task, int reward_type, const void* next_achieved_goal, const void* desired_goal, const long long* src,
                   const long long* goal_src, void* desired_goal_out, void* achieved_goal_out, float* reward, long long m, int dtype,
                   void* stream);
int pg_compute_reward_host(int task, int reward_type, const void* achieved_goal, const void* desired_goal, float* reward,
                           long long m, int dtype, int device);
int pg_is_success_host(int task, const void* achieved_goal, const void* desired_goal, unsigned char* success, long long m,
                       int dtype, int device);

/* RobotTaskEnv.save_state / restore_state / remove_state (core.py:252-278; pybullet.py:61-68,266-280): bit-exact device
 * snapshot of the whole batch (state, goals, episode counters). */
int pg_save_state(pg_env* env, int* state_id);
int pg_restore_state(pg_env* env, int state_id);
int pg_remove_state(pg_env* env, int state_id);

/* Raw state exchange for the facade getters/setters (pybullet.py:284-460) and the parity tests: float64 rows
 * [q(9) qd(9) | per object: pos(3) quat(4) lin(3) ang(3) | goal(G) | episode step], state_dim wide. */
int pg_get_state(pg_env* env, double* state, void* stream);
int pg_set_state(pg_env* env, const double* state, const unsigned char* mask, void* stream);
/* calculateInverseKinematics on link 11 from the current joint state (pybullet.py:479-497): target [N,3] + quat [N,4]
 * float64 -> joint angles [N,7] float64. */
int pg_inverse_kinematics(pg_env* env, const double* position, const double* orientation, double* joint_angles, void* stream);

/* End-effector (link 11) pose from the current joint state, rows [x y z qx qy qz qw] float64 (the fork's get_ee_position /
 * get_ee_orientation, panda_gym/envs/robots/panda_cartesian.py:218-225). */
int pg_get_ee_pose(pg_env* env, double* pose, void* stream);

/* Episode statistics accumulated by auto-reset since creation: {episodes, successes, return_sum, length_sum} (host). */
int pg_stats(pg_env* env, double out[4]);
/* Number of env-steps that ended in a non-finite state (counted; with auto_reset the env is truncated and restarted). */
int pg_diverged(pg_env* env, long long* count);
/* Scheduling introspection (host buffers): the per-env 16-bit key written by the last launch and the thread -> env map built from
 * the previous one (bits 0-4 contacts at the end of the launch, bit 5 robot contact, bit 6 solver ran all sweeps, bit 7 near a
 * contact, bit 9 full joint-limit sweep, bits 10-13 generic contacts (Stack)). */
int pg_debug_schedule(pg_env* env, unsigned short* key, int* perm);
/* Timing introspection (handles created with PG_DEBUG_TIMING=1 in the environment): per thread slot of the last launch, the SM
 * cycles spent in the env's step code and the key it wrote; out is a host buffer of num_envs x 2 int64. */
int pg_debug_timing(pg_env* env, long long* out);
/* Number of kernels this library has launched in this process. */
long long pg_kernel_launches(void);
const char* pg_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
