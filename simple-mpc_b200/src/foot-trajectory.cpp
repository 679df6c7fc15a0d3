// foot-trajectory.cpp — see include/simple-mpc/foot-trajectory.hpp
#include "simple-mpc/foot-trajectory.hpp"

namespace simple_mpc
{

  point3_t piecewise_curve::operator()(float t) const
  {
    const int degree = 8;
    const double u = (double)t; // T_min = 0, T_max = 1
    const double u_op = 1.0 - u;
    double bc = 1, tn = 1;
    point3_t tmp = control_points[0] * u_op;
    for (int i = 1; i < degree; i++)
    {
      tn = tn * u;
      bc = bc * (double)(degree - i + 1) / i;
      tmp = (tmp + control_points[(size_t)i] * (tn * bc)) * u_op;
    }
    return tmp + control_points[(size_t)degree] * (tn * u);
  }

  FootTrajectory::FootTrajectory(const std::map<std::string, point3_t> & initial_poses, double swing_apex, int T_fly, int T_contact, size_t T)
  {
    for (auto const & it : initial_poses)
      references_.insert({it.first, std::vector<point3_t>(T, point3_t{{0, 0, 0}})});
    initial_poses_ = initial_poses;
    final_poses_ = initial_poses;
    swing_apex_ = swing_apex;
    T_fly_ = T_fly;
    T_contact_ = T_contact;
    T_ = T;
    for (auto const & pose : initial_poses_)
      swing_trajectories_.insert({pose.first, defineTranslationBezier(pose.second, pose.second)});
  }

  piecewise_curve FootTrajectory::defineTranslationBezier(const point3_t & trans_init, const point3_t & trans_final)
  {
    // 4 x init (zero initial velocity, acceleration, jerk), one mid point lifted by the apex, 4 x final
    piecewise_curve c;
    for (size_t i = 0; i < 4; i++)
      c.control_points[i] = trans_init;
    point3_t mid;
    for (size_t k = 0; k < 3; k++)
      mid[k] = trans_init[k] * 3 / 4 + trans_final[k] * 1 / 4;
    mid[2] += swing_apex_;
    c.control_points[4] = mid;
    for (size_t i = 5; i < 9; i++)
      c.control_points[i] = trans_final;
    return c;
  }

  std::vector<point3_t>
  FootTrajectory::createTrajectory(int time_to_land, point3_t & initial_trans, point3_t & final_trans, piecewise_curve trajectory_swing)
  {
    std::vector<point3_t> trajectory;
    for (int t = time_to_land; t > time_to_land - (int)T_; t--)
    {
      if (t < 0)
        trajectory.push_back(final_trans);
      else if (t > T_fly_)
        trajectory.push_back(initial_trans);
      else
        trajectory.push_back(trajectory_swing(float(T_fly_ - t) / float(T_fly_)));
    }
    return trajectory;
  }

  void FootTrajectory::updateTrajectory(
    bool update, int landing_time, const point3_t & ee_trans, const point3_t & final_trans, const std::string & ee_name)
  {
    if (update)
    {
      initial_poses_.at(ee_name) = ee_trans;
      final_poses_.at(ee_name) = final_trans;
      swing_trajectories_.at(ee_name) = defineTranslationBezier(initial_poses_.at(ee_name), final_poses_.at(ee_name));
    }
    references_.at(ee_name) =
      createTrajectory(landing_time, initial_poses_.at(ee_name), final_poses_.at(ee_name), swing_trajectories_.at(ee_name));
  }

} // namespace simple_mpc
