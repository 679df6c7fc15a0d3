///////////////////////////////////////////////////////////////////////////////
// kinodynamics-id.hpp — KinodynamicsID with the reference's public interface
// (reference include/simple-mpc/inverse-dynamics/kinodynamics-id.hpp:17-90). The TSID formulation and the ProxQP
// solve behind it run as one CUDA kernel per call through the C ABI of include/smpc_id.h (one context, batch 1);
// batches of controllers use that C ABI directly (smpc_id_create with batch > 1).
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/robot-handler.hpp"
#include "smpc_id.h"

namespace simple_mpc
{

  class KinodynamicsID
  {
  public:
    typedef VectorXd TargetContactForce;

    struct Settings
    {
      // Physical quantities
      double friction_coefficient = 0.6;
      double contact_weight_ratio_max = 10.0; // Max force for one foot contact (multiple of the robot weight)
      double contact_weight_ratio_min = 0.01; // Min force for one foot contact (multiple of the robot weight)
      // Tasks gains
      double kp_base = 0.;
      double kp_posture = 0.;
      double kp_contact = 0.;
      // Tasks weights (<= 0: disabled)
      double w_base = -1.;
      double w_posture = -1.;
      double w_contact_motion = -1.;
      double w_contact_force = -1.;
      /// contact motion = 0 as a hard constraint (true) or a cost (false)
      bool contact_motion_equality = false;
    };

    KinodynamicsID(const RobotModelHandler & model_handler, double control_dt, const Settings settings);
    KinodynamicsID(const KinodynamicsID &) = delete;
    virtual ~KinodynamicsID();

    /// reference src/inverse-dynamics/kinodynamics-id.cpp:121-186
    void setTarget(
      const ConstVectorRef & q_target, const ConstVectorRef & v_target, const ConstVectorRef & a_target,
      const std::vector<bool> & contact_state_target, const std::vector<TargetContactForce> & f_target);

    /// reference :188-232 ; throws std::runtime_error if the QP was not solved to tolerance (the reference asserts)
    void solve(const double t, const ConstVectorRef & q_meas, const ConstVectorRef & v_meas, VectorXd & tau_res);

    /// reference :234-237
    void getAccelerations(VectorXd & a);

    /// contact force variables of the last solve: nfeet x 12 (3 used for point feet, 4 corners x 3 for quad feet)
    std::vector<VectorXd> getContactForces();

  public:
    const Settings settings_;
    const RobotModelHandler & model_handler_;

  protected:
    /// shared by CentroidalID: the full settings record of the C ABI
    KinodynamicsID(const RobotModelHandler & model_handler, double control_dt, const Settings & settings, const smpc_id_settings & full);
    static smpc_id_settings lower(const Settings & s);
    void create(double control_dt, const smpc_id_settings & full);
    void check(int rc, const char * what) const;

    smpc_id_ctx * ctx_ = nullptr;
  };

} // namespace simple_mpc
