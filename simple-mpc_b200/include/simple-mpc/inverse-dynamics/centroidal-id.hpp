///////////////////////////////////////////////////////////////////////////////
// centroidal-id.hpp — CentroidalID with the reference's public interface
// (reference include/simple-mpc/inverse-dynamics/centroidal-id.hpp:9-56): KinodynamicsID plus a centre-of-mass task,
// swing-foot tracking tasks and a base task restricted to the orientation.
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/inverse-dynamics/kinodynamics-id.hpp"

namespace simple_mpc
{

  /// spatial velocity [linear; angular] (pinocchio::Motion stand-in)
  struct Motion
  {
    Vector3d lin{{0, 0, 0}}, ang{{0, 0, 0}};
    void setZero() { lin = ang = Vector3d{{0, 0, 0}}; }
    const Vector3d & linear() const { return lin; }
    const Vector3d & angular() const { return ang; }
  };

  class CentroidalID : public KinodynamicsID
  {
  public:
    typedef VectorXd TargetContactForce;
    typedef std::vector<SE3> FeetPoseVector;
    typedef std::vector<Motion> FeetVelocityVector;

    struct Settings : public KinodynamicsID::Settings
    {
      // Tasks gains
      double kp_com = 0.;
      double kp_feet_tracking = 0.;
      // Tasks weights (<= 0: disabled)
      double w_com = -1.;
      double w_feet_tracking = -1.;
    };

    CentroidalID(const RobotModelHandler & model_handler, double control_dt, const Settings settings);
    CentroidalID(const CentroidalID &) = delete;

    /// reference src/inverse-dynamics/centroidal-id.cpp:88-145
    void setTarget(
      const Vector3d & com_position, const Vector3d & com_velocity, const FeetPoseVector & feet_pose,
      const FeetVelocityVector & feet_velocity, const std::vector<bool> & contact_state_target,
      const std::vector<TargetContactForce> & f_target);

    using KinodynamicsID::getAccelerations;
    using KinodynamicsID::solve;

  private:
    using KinodynamicsID::setTarget;
    static smpc_id_settings lower(const Settings & s);

  public:
    const Settings settings_;
  };

} // namespace simple_mpc
