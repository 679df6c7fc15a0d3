#include "simple-mpc/inverse-dynamics/centroidal-id.hpp"

namespace simple_mpc
{

  smpc_id_settings CentroidalID::lower(const Settings & s)
  {
    smpc_id_settings o = KinodynamicsID::lower(s);
    o.centroidal = 1;
    o.kp_com = s.kp_com;
    o.kp_feet_tracking = s.kp_feet_tracking;
    o.w_com = s.w_com;
    o.w_feet_tracking = s.w_feet_tracking;
    return o;
  }

  // reference centroidal-id.cpp:6-86: the extra tasks, the default target (CoM of the reference posture, feet on their
  // reference placements, all in contact) and the dry run are set up inside smpc_id_create for a `centroidal` context
  CentroidalID::CentroidalID(const RobotModelHandler & model_handler, double control_dt, const Settings settings)
  : KinodynamicsID(model_handler, control_dt, settings, lower(settings))
  , settings_(settings)
  {
  }

  void CentroidalID::setTarget(
    const Vector3d & com_position, const Vector3d & com_velocity, const FeetPoseVector & feet_pose,
    const FeetVelocityVector & feet_velocity, const std::vector<bool> & contact_state_target,
    const std::vector<TargetContactForce> & f_target)
  {
    const size_t nf = model_handler_.getFeetNb();
    if (feet_pose.size() != nf || feet_velocity.size() != nf || contact_state_target.size() != nf)
      throw std::runtime_error("CentroidalID::setTarget: one entry per foot expected");
    std::vector<int32_t> cs(nf);
    std::vector<double> pose(12 * nf), vel(6 * nf), ft(6 * nf, 0.0);
    for (size_t f = 0; f < nf; f++)
    {
      cs[f] = contact_state_target[f] ? 1 : 0;
      for (size_t e = 0; e < 9; e++)
        pose[12 * f + e] = feet_pose[f].R[e];
      for (size_t e = 0; e < 3; e++)
      {
        pose[12 * f + 9 + e] = feet_pose[f].p[e];
        vel[6 * f + e] = feet_velocity[f].lin[e];
        vel[6 * f + 3 + e] = feet_velocity[f].ang[e];
      }
      if (!contact_state_target[f])
        continue;
      const size_t fs = model_handler_.getFootType(f) == RobotModelHandler::QUAD ? 6 : 3;
      if (f_target.size() <= f || f_target[f].size() != fs)
        throw std::runtime_error("CentroidalID::setTarget: f_target has the wrong size");
      for (size_t e = 0; e < fs; e++)
        ft[6 * f + e] = f_target[f][e];
    }
    check(
      smpc_id_set_target_centroidal(ctx_, 0, 1, com_position.data(), com_velocity.data(), pose.data(), vel.data(), cs.data(), ft.data()),
      "CentroidalID::setTarget");
  }

} // namespace simple_mpc
