/*
 * smpc.h — C ABI of the B200-native per-control-step MPC solve.
 *
 * This is the drop-in boundary for the one hot path of Simple-Robotics/simple-mpc:
 * everything `MPC::iterate` hands to `aligator::SolverProxDDP::run`
 * (reference src/mpc.cpp:189-218, the call at :212) and reads back at :215-217,
 * plus the problem description that `OCPHandler::createProblem` builds
 * (reference src/ocp-handler.cpp:96-137) lowered to flat POD tables.
 *
 * The reference has no FFI for this path (its "interface" is the concrete C++ type
 * aligator::SolverProxDDPTpl<double>, reference include/simple-mpc/mpc.hpp:89); the entry
 * points below are what a binding for it has to export. Each one cites the reference
 * call it replaces.
 *
 * Conventions
 *   - all floating point is IEEE fp64, all matrices are ROW-major unless stated;
 *   - plain pointers and sizes only; every buffer is caller-owned HOST memory unless the
 *     name ends in `_dev` (device pointers, for callers that already live on the GPU);
 *   - every function returns 0 on success and a negative smpc_status on error; the text of
 *     the last error of the calling thread is returned by smpc_last_error();
 *   - one context is driven from one host thread at a time (like one reference `MPC` object);
 *   - multibody state x = [q; v], q = [p(3), quat(x,y,z,w), q_joints], v = [v_lin, w (base frame),
 *     qd_joints]; tangent dx in R^{2 nv} (reference src/robot-handler.cpp:81-96);
 *   - spatial vectors are [linear(3); angular(3)].
 *
 * The CPU oracle (oracle/) consumes the same descriptor structs through `orc_*` entry points
 * with identical signatures; it is test infrastructure and is never linked into libsmpc_cuda.
 */
#ifndef SMPC_H_
#define SMPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMPC_MAX_JOINTS 32
#define SMPC_MAX_NV 32
#define SMPC_MAX_NQ 36
#define SMPC_MAX_NX 64  /* >= nq + nv  */
#define SMPC_MAX_NDX 64 /* >= 2 nv     */
#define SMPC_MAX_NU 40
#define SMPC_MAX_FEET 4

typedef enum smpc_status
{
  SMPC_OK = 0,
  SMPC_ERR_INVALID = -1,     /* bad argument / size / index (std::runtime_error in the reference) */
  SMPC_ERR_UNSUPPORTED = -2, /* configuration the CUDA path has no kernel for */
  SMPC_ERR_CUDA = -3,        /* CUDA runtime failure; there is NO CPU fallback */
  SMPC_ERR_STATE = -4        /* call order violation (e.g. solve before upload)  */
} smpc_status;

typedef enum smpc_joint_type
{
  SMPC_JOINT_FREEFLYER = 0, /* nq 7, nv 6 (root only) */
  SMPC_JOINT_REVOLUTE = 1   /* nq 1, nv 1, arbitrary unit axis in the joint frame */
} smpc_joint_type;

typedef enum smpc_ocp_kind
{
  SMPC_OCP_KINODYNAMICS = 0, /* reference src/kinodynamics.cpp */
  SMPC_OCP_FULLDYNAMICS = 1, /* reference src/fulldynamics.cpp */
  SMPC_OCP_CENTROIDAL = 2    /* reference src/centroidal-dynamics.cpp */
} smpc_ocp_kind;

/* Kinematic tree + inertias + the frames the path uses. Stands in for the pinocchio::Model that
 * RobotModelHandler copies (reference src/robot-handler.cpp:12-79). Joint 0 is the free-flyer root;
 * joints are topologically ordered (parent[i] < i). */
typedef struct smpc_model_desc
{
  int32_t njoints;
  int32_t nq, nv;
  int32_t parent[SMPC_MAX_JOINTS]; /* -1 for the root */
  int32_t jtype[SMPC_MAX_JOINTS];
  int32_t idx_q[SMPC_MAX_JOINTS];
  int32_t idx_v[SMPC_MAX_JOINTS];
  double axis[SMPC_MAX_JOINTS][3];
  double place_R[SMPC_MAX_JOINTS][9]; /* parent-joint frame <- joint frame (before the joint motion) */
  double place_p[SMPC_MAX_JOINTS][3];
  double mass[SMPC_MAX_JOINTS];       /* body attached to the joint, expressed in the joint frame */
  double lever[SMPC_MAX_JOINTS][3];   /* centre of mass */
  double inertia[SMPC_MAX_JOINTS][9]; /* rotational inertia about the centre of mass */
  /* feet, in insertion order (addPointFoot/addQuadFoot order, reference src/robot-handler.cpp:56-73) */
  int32_t nfeet;
  int32_t foot_joint[SMPC_MAX_FEET];
  double foot_R[SMPC_MAX_FEET][9]; /* joint frame <- foot frame */
  double foot_p[SMPC_MAX_FEET][3];
  /* "<foot>_ref" frames used by the Raibert heuristic (reference src/robot-handler.cpp:28-53) */
  int32_t footref_joint[SMPC_MAX_FEET];
  double footref_R[SMPC_MAX_FEET][9];
  double footref_p[SMPC_MAX_FEET][3];
  /* base frame */
  int32_t base_joint;
  double base_R[9];
  double base_p[3];
  double q_ref[SMPC_MAX_NQ]; /* reference configuration; reference state = [q_ref; 0] */
} smpc_model_desc;

/* Problem recipe: the settings structs of the three handlers (reference
 * include/simple-mpc/kinodynamics.hpp:24-51, fulldynamics.hpp:28-65, centroidal-dynamics.hpp:33-52)
 * plus the createProblem arguments (reference src/ocp-handler.cpp:96-101). Weight matrices are dense
 * row-major, sized as in the reference; unused ones may be NULL. */
typedef struct smpc_ocp_desc
{
  int32_t kind;       /* smpc_ocp_kind */
  int32_t horizon;    /* number of stages T */
  int32_t force_size; /* 3 or 6 */
  int32_t kinematics_limits, torque_limits, force_cone, land_cstr;
  int32_t terminal_constraint;
  double timestep;
  double gravity[3];
  double mu, Lfoot, Wfoot;
  const double * w_x;       /* ndx x ndx        (kino, full)                 */
  const double * w_u;       /* nu x nu                                        */
  const double * w_cent;    /* 6 x 6            (kino, full)                 */
  const double * w_centder; /* 6 x 6            (kino)                       */
  const double * w_frame;   /* fs x fs... 3x3 for point feet, 6x6 for quad feet (kino, full) */
  const double * w_forces;  /* fs x fs          (full)                       */
  const double * w_com;     /* 3 x 3            (centroidal)                 */
  const double * w_linear_mom, *w_angular_mom, *w_linear_acc, *w_angular_acc; /* 3x3 (centroidal) */
  const double * qmin, *qmax; /* nv-6 */
  const double * umin, *umax; /* nv-6 (full) */
  double Kp_correction[6], Kd_correction[6]; /* Baumgarte gains (full) */
} smpc_ocp_desc;

/* SolverProxDDP constructor arguments and member settings touched by the reference
 * (src/mpc.cpp:43-53,91) followed by the upstream defaults this implementation assumes
 * (SURVEY.md Appendix B). smpc_default_solver_opts() fills the defaults. */
typedef struct smpc_solver_opts
{
  double tol;     /* MPCSettings::TOL */
  double mu_init; /* MPCSettings::mu_init */
  int32_t max_iters;
  int32_t max_al_iters;
  double prim_alpha, prim_beta, dual_alpha, dual_beta; /* BCL exponents */
  double mu_update_factor, mu_lower_bound;
  double inner_tol0, prim_tol0;
  double reg_init, reg_min, reg_max, reg_inc_k, reg_inc_critical, reg_dec_k;
  double ls_armijo_c1, ls_alpha_min, ls_contraction, ls_avg_eta;
  int32_t ls_max_steps;
  /* MPCSettings::num_threads (reference src/mpc.cpp:46-50): > 1 selects the PARALLEL (leg-condensed) Riccati sweep with
   * that many legs - one CTA per leg, csrc/riccati_par.cuh - when the batch leaves SMs idle (batch x legs <= SM count),
   * a tuned sweep exists for the shape and the leg-parallel sweep was measured to be the faster one for it (today: the
   * centroidal problems; the environment variable SMPC_LEGS=n forces n legs for any supported shape). Everything else runs
   * the serial sweep; results agree up to rounding either way. 0 or 1: serial. smpc_riccati_legs() reports what runs. */
  int32_t num_threads;
} smpc_solver_opts;

/* One horizon slot = what one aligator StageModel holds that simple-mpc ever mutates: the contact
 * phase it was created with (reference src/kinodynamics.cpp:40-152) and its references
 * (setReferencePose / setReferenceState / setReferenceControl / setReferenceForce). */
typedef struct smpc_slot
{
  uint32_t contact_mask; /* bit i set: foot i (insertion order) in contact */
  uint32_t land_mask;    /* bit i set: landing-edge constraints on foot i  */
  double x_ref[SMPC_MAX_NX];
  double u_ref[SMPC_MAX_NU];
  double foot_ref[SMPC_MAX_FEET][12]; /* R row-major (9) then p (3) */
  double f_ref[SMPC_MAX_FEET][6];     /* contact-force cost reference (full dynamics) */
  /* target height of the landing-edge equality (FrameTranslationResidual row 2, reference src/kinodynamics.cpp:134-143,
   * src/fulldynamics.cpp:196-204): the z of contact_pose at createStage time. It is a separate residual from the
   * <foot>_pose_cost, so setReferencePose / smpc_set_foot_refs / the device driver never touch it. */
  double land_z[SMPC_MAX_FEET];
} smpc_slot;

/* Terminal node data (reference createTerminalCost / createTerminalConstraint). */
typedef struct smpc_terminal
{
  double x_ref[SMPC_MAX_NX];
  double com_ref[3]; /* DCM / CoM reference for the terminal constraint (updateTerminalConstraint rewrites it) */
  double tau;        /* DCM time constant, fixed at createTerminalConstraint: sqrt(com_ref_at_creation[2] / 9.81)
                      * (reference src/kinodynamics.cpp:372, src/fulldynamics.cpp:439) */
} smpc_terminal;

/* Scalars of aligator::ResultsTpl that the parity target names. */
typedef struct smpc_solve_info
{
  int32_t num_iters;
  int32_t al_iters;
  int32_t conv;
  int32_t ls_trials;   /* merit evaluations of the last iteration */
  int32_t flags;       /* bit 0: a Riccati factorisation met a non-positive pivot; bit 2: the leg-parallel sweep could not
                          factorise its condensed system and the serial sweep ran instead (results are the serial ones) */
  int32_t pad_;
  double alpha;        /* accepted step of the last iteration */
  double prim_infeas, dual_infeas;
  double traj_cost, merit;
  double mu, preg;
} smpc_solve_info;

typedef struct smpc_ctx smpc_ctx;

const char * smpc_last_error(void);
void smpc_default_solver_opts(smpc_solver_opts * opts);
/* sizes implied by (model, ocp): state, tangent, control */
int smpc_problem_dims(const smpc_model_desc *, const smpc_ocp_desc *, int * nx, int * ndx, int * nu);

/* replaces: SolverProxDDP ctor + setup(problem)           (reference src/mpc.cpp:43-53,84)
 * `batch` independent MPC instances share the model, the recipe and the solver options. */
int smpc_create(
  smpc_ctx ** out,
  int device,
  int batch,
  const smpc_model_desc * model,
  const smpc_ocp_desc * ocp,
  const smpc_solver_opts * opts);
void smpc_destroy(smpc_ctx *);

/* replaces: TrajOptProblem(x0, stages, terminal cost)     (reference src/ocp-handler.cpp:130)
 * slots: [n_inst][T]; instance range [first, first+n_inst). */
int smpc_set_slots(smpc_ctx *, int first, int n_inst, const smpc_slot * slots);
/* replaces: the per-tick setReferencePose/State/VelocityBase calls on one stage
 * (reference src/mpc.cpp:304-312); writes slot t of instances [first, first+n_inst). */
int smpc_set_slot(smpc_ctx *, int first, int n_inst, int t, const smpc_slot * slots);
/* only the foot pose references of every slot: foot_ref [n_inst][T][nfeet][12] */
int smpc_set_foot_refs(smpc_ctx *, int first, int n_inst, const double * foot_ref);
int smpc_set_terminal(smpc_ctx *, int first, int n_inst, const smpc_terminal * term);

/* replaces: replaceStageCircular + cycleProblem           (reference src/mpc.cpp:225-226,246-247)
 * drops slot 0 of each instance, appends new_tail[i]; solver duals rotate with it. */
int smpc_cycle(smpc_ctx *, int first, int n_inst, const smpc_slot * new_tail);
/* replaces: MPC::generateCycleHorizon's table of pre-built stages   (reference src/mpc.cpp:101-187)
 * the stages stay resident in HBM; smpc_cycle_ids appends table[ids[i]] to instance first+i, so a tick
 * moves 4 bytes per instance over PCIe instead of one smpc_slot. */
int smpc_upload_stage_table(smpc_ctx *, const smpc_slot * stages, int n);
int smpc_cycle_ids(smpc_ctx *, int first, int n_inst, const int * ids);

/* replaces: the warm-start arguments of run(problem, xs_, us_)   (reference src/mpc.cpp:201-207,212)
 * xs [n_inst][T+1][nx], us [n_inst][T][nu]. */
int smpc_set_warmstart(smpc_ctx *, int first, int n_inst, const double * xs, const double * us);
/* the MPC shift of reference src/mpc.cpp:201-207 done in place on the device copy */
int smpc_shift_warmstart(smpc_ctx *, int first, int n_inst);

/* replaces: setInitState(x0) + solver_->run(...)          (reference src/mpc.cpp:209,212)
 * x0 [batch][nx]. Runs up to max_iters ProxDDP iterations for every instance. */
int smpc_solve(smpc_ctx *, const double * x0, int max_iters);
/* host -> device copy of the measured states only (x0 [batch][nx]); smpc_solve(ctx, NULL, n) then reuses them */
int smpc_set_x0(smpc_ctx *, const double * x0);

/* replaces: results_.xs / us / getCtrlFeedbacks()          (reference src/mpc.cpp:215-217)
 * any pointer may be NULL. xs [n][T+1][nx], us [n][T][nu], Ks [n][T][nu][ndx] (row-major, du = k + K dx),
 * ks [n][T][nu]. */
int smpc_get_results(smpc_ctx *, int first, int n_inst, double * xs, double * us, double * Ks, double * ks);
/* the head of the solution only: what a controller consumes each tick
 * x1 [n][nx], u0 [n][nu], u1 [n][nu], K0 [n][nu][ndx] */
int smpc_get_head(smpc_ctx *, int first, int n_inst, double * x1, double * u0, double * u1, double * K0);
int smpc_get_info(smpc_ctx *, int first, int n_inst, smpc_solve_info * info);
/* duals: vs [n][T][ncmax] (ncmax from smpc_get_ncmax), lams [n][T+1][ndx] */
int smpc_get_duals(smpc_ctx *, int first, int n_inst, double * vs, double * lams);
int smpc_get_ncmax(smpc_ctx *);

/* replaces: MPC::getStateDerivative(t)                    (reference src/mpc.cpp:346-352)
 * xdot [n][2 nv] (9 for centroidal) of the last evaluation at stage t. */
int smpc_get_xdot(smpc_ctx *, int first, int n_inst, int t, double * xdot);
/* replaces: MPC::getContactForces(t)                      (reference src/mpc.cpp:354-380) */
int smpc_get_contact_forces(smpc_ctx *, int first, int n_inst, int t, double * forces);

/* ---- batched MPC driver: the part of MPC::iterate around run(), per instance, on the device ----------
 * (SURVEY.md §8a rows a1, a2, a13, a15). With thousands of instances the reference's host loop
 * (recedeWithCycle + updateStepTrackerReferences: O(T nk) setter calls per instance, reference
 * src/mpc.cpp:220-324) and the [B][T][nk][3] reference upload would dominate the tick, so each instance
 * keeps its gait bookkeeping and swing-curve end points in HBM and one launch regenerates all references. */
typedef struct smpc_mpc_settings
{
  int32_t T_fly, T_contact;    /* MPCSettings (reference include/simple-mpc/mpc.hpp:36-58) */
  double swing_apex, timestep;
  int32_t ncycle;      /* stage-table entries [0, ncycle) are cycle_horizon_ (reference src/mpc.cpp:101-187) */
  int32_t standing_id; /* stage-table index of the standing stage (standing_horizon_), -1 if none */
} smpc_mpc_settings;

/* replaces: the MPC constructor's driver state + generateCycleHorizon's timing tables
 * (reference src/mpc.cpp:19-41,93-98,112-132). Needs smpc_upload_stage_table first.
 * phase[i] (NULL = all 0): instance i starts its cycle at table entry phase[i], i.e. behaves like a reference MPC
 * whose generateCycleHorizon was given the gait rotated left by phase[i].
 * x_reference [nx]: MPC::x_reference_ (problem state size). Instances start WALKING with zero base velocity. */
int smpc_mpc_setup(smpc_ctx *, const smpc_mpc_settings *, const int * phase, const double * x_reference);
/* replaces: mpc.velocity_base = v / switchToWalk(v)     (reference src/mpc.cpp:382-386); vb [n_inst][6] */
int smpc_mpc_set_velocity_base(smpc_ctx *, int first, int n_inst, const double * vb);
/* replaces: switchToWalk / switchToStand                (reference src/mpc.cpp:382-392); standing zeroes velocity_base */
int smpc_mpc_set_walking(smpc_ctx *, int first, int n_inst, int walking);
/* replaces: MPC::iterate(x)                             (reference src/mpc.cpp:189-218)
 * x_meas [batch][nq+nv] measured multibody states (NULL: keep the states already on the device).
 * updateInternalData -> recedeWithCycle -> updateStepTrackerReferences -> warm-start shift -> run(max_iters). */
int smpc_mpc_iterate(smpc_ctx *, const double * x_meas, int max_iters);
/* no reference counterpart (launch mechanics only): on != 0 makes smpc_mpc_iterate replay ONE captured CUDA Graph per tick
 * (driver kernels + one ProxDDP round + its first linesearch trial, one host synchronisation); a linesearch that
 * backtracks or an extra round continues with plain launches, so results are identical either way (GPU test). For batches
 * up to 256 instances the graph also copies the head of the plan to pinned memory, which the next smpc_get_head serves
 * without a second synchronisation. Default: on when the batch does not exceed the SM count (latency-bound ticks), off
 * otherwise; the environment variable SMPC_GRAPH=0|1 at smpc_create overrides. */
int smpc_mpc_use_graph(smpc_ctx *, int on);
/* legs of the PARALLEL Riccati sweep this context runs (1 = serial), see smpc_solver_opts::num_threads */
int smpc_riccati_legs(smpc_ctx *);
/* closed loop without a plant (multibody problem states only): x_meas <- predicted xs[1], device to device */
int smpc_mpc_xmeas_from_prediction(smpc_ctx *);
/* replaces: getFootTakeoffCycle / getFootLandCycle      (reference include/simple-mpc/mpc.hpp:141-164)
 * takeoff, land [n_inst][nfeet]: first pending event or -1. */
int smpc_mpc_get_timings(smpc_ctx *, int first, int n_inst, int * takeoff, int * land);
/* replaces: getReferencePose / getReferenceState / getVelocityBase / getContactState read-backs
 * (reference src/mpc.cpp:326-344): slots [n_inst][T] in horizon order; term may be NULL */
int smpc_get_slots(smpc_ctx *, int first, int n_inst, smpc_slot * slots, smpc_terminal * term);

/* ---- the consumer of (us, Ks, xs) between two MPC ticks, batched --------------------
 * replaces: the 1 kHz loop of the reference examples,
 *     u = us[0] - Ks[0] * difference(x_meas, xs[0])
 * (examples/talos_fulldynamics.py:203-205, examples/go2_fulldynamics.py:282-284; RobotModelHandler::difference,
 * src/robot-handler.cpp:81-96: difference(x1, x2) = x2 (-) x1 on the multibody state manifold).
 * x_meas [n_inst][nx] -> u [n_inst][nu]; the plan and the gains stay on the device. */
int smpc_feedback_control(smpc_ctx *, int first, int n_inst, const double * x_meas, double * u);
/* replaces: Interpolator::interpolateState / interpolateLinear / interpolateContacts applied to the plan
 * (reference src/interpolator.cpp:5-78; examples/go2_fulldynamics.py:272-275): step = floor(delay / timestep),
 * beyond the last knot the last value is returned. x_interp [n][nx] (multibody states: Lie-group interpolation of
 * q, linear in v; the 9-D centroidal state is interpolated linearly), u_interp [n][nu], contacts [n][nfeet] (0/1).
 * Any output may be NULL. */
int smpc_interpolate_plan(smpc_ctx *, int first, int n_inst, double delay, double * x_interp, double * u_interp, int * contacts);

/* device-side timing of the last smpc_solve (CUDA events on the context's stream), ms */
int smpc_last_solve_ms(smpc_ctx *, float * total_ms, float * riccati_ms);
/* number of kernel launches issued by this context so far */
int64_t smpc_launch_count(smpc_ctx *);
/* which Riccati kernel this context launches ("tma-staged", "specialised", "generic") */
const char * smpc_riccati_variant(smpc_ctx *);
/* closed loop without a plant: x0 <- the predicted next state xs[1] of every instance (device to device) */
int smpc_x0_from_prediction(smpc_ctx *);
/* CUDA-event timer on the context's own stream (which = 0 start, 1 stop) and accumulated per-kernel device time */
int smpc_timer_record(smpc_ctx *, int which);
int smpc_timer_elapsed_ms(smpc_ctx *, float * ms);
int smpc_timer_stats(smpc_ctx *, int reset, double * riccati_ms, int64_t * riccati_launches, double * solve_ms, int64_t * solves);
/* per-kernel device time (ms) of the first round of every solve since the last smpc_timer_stats(reset = 1):
 * ms4 = {linearisation, Riccati backward sweep, forward sweep, first trial evaluation + linesearch decision} */
int smpc_kernel_times(smpc_ctx *, double * ms4);
int smpc_sync(smpc_ctx *);

/* test hooks: the dense LQ knot of (instance, t) from the last linearisation, in the device layout
 * described by smpc_knot_layout: offsets {Q,S,R,q,r,A,B,f,C,D,d,size} in doubles, row-major blocks */
int smpc_get_knot(smpc_ctx *, int inst, int t, int * nc, double * knot_out);
int smpc_knot_layout(smpc_ctx *, int * out12);

#ifdef __cplusplus
}
#endif
#endif /* SMPC_H_ */
