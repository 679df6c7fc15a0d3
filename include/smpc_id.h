/*
 * smpc_id.h — C ABI of the batched whole-body inverse-dynamics QP.
 *
 * The reference's low-level controllers `KinodynamicsID` / `CentroidalID`
 * (reference include/simple-mpc/inverse-dynamics/kinodynamics-id.hpp:17-90,
 *  src/inverse-dynamics/kinodynamics-id.cpp:6-237, centroidal-id.hpp:9-56, centroidal-id.cpp:6-145)
 * run after `MPC::iterate` in the examples' 1 kHz loop: a TSID acceleration/force formulation
 * (posture, base, CoM, swing-foot and contact tasks, joint and actuation bounds) solved as ONE dense QP per
 * call by ProxQP. TSID 1.9.0 and proxsuite 0.7.2 are not vendored in /root/reference (pixi.lock:174,156), so
 * what is restated here is TSID's published formulation (Del Prete et al., "Implementing torque control with
 * high-ratio gear boxes", tsid/formulations/inverse-dynamics-formulation-acc-force [UPSTREAM]) at the
 * reference's call sites; the QP is solved to a tight KKT tolerance, so results are compared at solution level.
 *
 * The reference has no FFI here either (tsid::InverseDynamicsFormulationAccForce + tsid::solvers::SolverProxQP
 * are C++ members, kinodynamics-id.hpp:74-83). One context = `batch` independent controllers for one robot.
 * Conventions as in smpc.h (fp64, row-major, host pointers, [linear; angular] spatial vectors, status codes,
 * smpc_last_error()). The CPU oracle exports the same entry points with the prefix `orc_id_`.
 */
#ifndef SMPC_ID_H_
#define SMPC_ID_H_

#include "smpc.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SMPC_ID_POINT 0 /* RobotModelHandler::FootType::POINT: 3 force variables (tsid ContactPoint)  */
#define SMPC_ID_QUAD 1  /* RobotModelHandler::FootType::QUAD : 4 corner forces = 12 variables (tsid Contact6d) */
#define SMPC_ID_MAX_FVAR 12

/* KinodynamicsID::Settings (kinodynamics-id.hpp:22-47) followed by CentroidalID::Settings (centroidal-id.hpp:17-26).
 * A task with a weight <= 0 is not added (the reference's "disabled by default" convention, `if (w > 0.)`). */
typedef struct smpc_id_settings
{
  double friction_coefficient;     /* 0.6  */
  double contact_weight_ratio_max; /* 10   */
  double contact_weight_ratio_min; /* 0.01 */
  double kp_base, kp_posture, kp_contact;
  double w_base, w_posture, w_contact_motion, w_contact_force;
  int32_t contact_motion_equality;
  int32_t centroidal; /* 0: KinodynamicsID, 1: CentroidalID (base task on the orientation only, CoM + swing-foot tasks) */
  double kp_com, kp_feet_tracking;
  double w_com, w_feet_tracking;
} smpc_id_settings;

/* What the constructor reads from RobotModelHandler beyond smpc_model_desc (kinodynamics-id.cpp:24-98):
 * foot types / quad contact points, joint position / velocity / effort limits of the nv-6 actuated joints. */
typedef struct smpc_id_desc
{
  double control_dt;
  int32_t foot_type[SMPC_MAX_FEET];
  double quad_points[SMPC_MAX_FEET][4][3]; /* corners in the foot frame (getQuadFootContactPoints) */
  double qmin[SMPC_MAX_NV], qmax[SMPC_MAX_NV];
  double vmax[SMPC_MAX_NV];
  double effort[SMPC_MAX_NV];
  smpc_id_settings settings;
} smpc_id_desc;

/* Per-instance outcome of the last solve. */
typedef struct smpc_id_info
{
  int32_t status;      /* 0 converged (tsid HQP_STATUS_OPTIMAL), 1 iteration limit */
  int32_t outer_iters; /* multiplier updates */
  int32_t newton_iters;
  int32_t nvar, neq, nin;
  double prim_res, dual_res; /* inf-norms of the KKT residuals at exit */
} smpc_id_info;

typedef struct smpc_id_ctx smpc_id_ctx;

void smpc_id_default_settings(smpc_id_settings *);

/* KinodynamicsID::KinodynamicsID / CentroidalID::CentroidalID (kinodynamics-id.cpp:6-119, centroidal-id.cpp:6-86):
 * builds the tasks, sets the default target (reference state, every foot in contact carrying weight / nfeet) for every
 * instance and performs the dry-run solve at the reference state. */
int smpc_id_create(smpc_id_ctx ** out, const smpc_model_desc * model, const smpc_id_desc * desc, int batch);
void smpc_id_destroy(smpc_id_ctx *);

/* KinodynamicsID::setTarget (kinodynamics-id.cpp:121-186). Per instance: q_target nq, v_target nv, a_target nv,
 * contact_state nfeet (0/1), f_target nfeet x 6 (first 3 used for point feet; rows of feet not in contact ignored).
 * Only valid for a KinodynamicsID context (the reference makes it private in CentroidalID). */
int smpc_id_set_target(
  smpc_id_ctx *, int first, int n_inst, const double * q_target, const double * v_target, const double * a_target,
  const int32_t * contact_state, const double * f_target);

/* CentroidalID::setTarget (centroidal-id.cpp:88-145). Per instance: com_position 3, com_velocity 3,
 * feet_pose nfeet x 12 (rotation row-major 9, translation 3), feet_velocity nfeet x 6, contact_state, f_target. */
int smpc_id_set_target_centroidal(
  smpc_id_ctx *, int first, int n_inst, const double * com_position, const double * com_velocity, const double * feet_pose,
  const double * feet_velocity, const int32_t * contact_state, const double * f_target);

/* KinodynamicsID::solve (kinodynamics-id.cpp:188-232) for every instance: q_meas batch x nq, v_meas batch x nv,
 * tau_res batch x (nv-6). `t` is passed through like the reference's (no task depends on it). */
int smpc_id_solve(smpc_id_ctx *, double t, const double * q_meas, const double * v_meas, double * tau_res);

/* KinodynamicsID::getAccelerations (kinodynamics-id.cpp:234-237): batch x nv. */
int smpc_id_get_accelerations(smpc_id_ctx *, int first, int n_inst, double * a);
/* formulation_.getContactForces: per instance nfeet x SMPC_ID_MAX_FVAR force variables (zeros for inactive contacts). */
int smpc_id_get_forces(smpc_id_ctx *, int first, int n_inst, double * f);
int smpc_id_get_info(smpc_id_ctx *, int first, int n_inst, smpc_id_info * info);

/* Device-resident variant for callers whose measurements already live in HBM (bench `value` arm): device pointers. */
int smpc_id_solve_dev(smpc_id_ctx *, double t, const double * q_meas_dev, const double * v_meas_dev, double * tau_res_dev);
/* Warm start (default on): a solve whose contact layout equals that of the instance's previous converged solve starts from
 * its primal / dual solution, like the ProxQP object TSID keeps between calls (tsid/solvers/solver-proxqp [UPSTREAM]).
 * The QP is strictly convex, so only the iteration count depends on it. 0 = every solve starts from zero. */
int smpc_id_set_warm_start(smpc_id_ctx *, int on);
int smpc_id_last_solve_ms(smpc_id_ctx *, float * ms);
int64_t smpc_id_launch_count(smpc_id_ctx *);

#ifdef __cplusplus
}
#endif
#endif
