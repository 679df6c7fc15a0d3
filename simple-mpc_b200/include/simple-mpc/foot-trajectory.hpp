///////////////////////////////////////////////////////////////////////////////
// foot-trajectory.hpp — swing-foot reference generator with the reference's interface
// (reference include/simple-mpc/foot-trajectory.hpp:20-62, src/foot-trajectory.cpp:20-96).
// ndcurves is not available: the degree-8 Bezier curve it builds is evaluated here directly.
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/fwd.hpp"

namespace simple_mpc
{
  using point3_t = Vector3d;

  /// Degree-8 Bezier on [0, 1] with 9 control points; the parameter type is FLOAT, as in the reference's
  /// `ndcurves::bezier_curve<float, double, false, point3_t>` (reference src/foot-trajectory.cpp:18,76).
  struct piecewise_curve
  {
    std::array<point3_t, 9> control_points;
    /// [UPSTREAM ndcurves bezier_curve::evalHorner] Horner-like scheme in double on u = double(t)
    point3_t operator()(float t) const;
  };

  class FootTrajectory
  {
  protected:
    std::map<std::string, point3_t> initial_poses_;
    std::map<std::string, point3_t> final_poses_;
    std::map<std::string, std::vector<point3_t>> references_;
    double swing_apex_ = 0;
    std::map<std::string, piecewise_curve> swing_trajectories_;
    int T_fly_ = 0;
    int T_contact_ = 0;
    size_t T_ = 0;

  public:
    explicit FootTrajectory() {}
    FootTrajectory(const std::map<std::string, point3_t> & initial_poses, double swing_apex, int T_fly, int T_contact, size_t T);
    virtual ~FootTrajectory() {}

    void updateApex(double swing_apex) { swing_apex_ = swing_apex; }
    piecewise_curve defineTranslationBezier(const point3_t & trans_init, const point3_t & trans_final);
    std::vector<point3_t> createTrajectory(int time_to_land, point3_t & initial_trans, point3_t & final_trans, piecewise_curve trajectory_swing);
    void updateTrajectory(bool update, int landing_time, const point3_t & ee_trans, const point3_t & final_trans, const std::string & ee_name);
    const std::vector<point3_t> & getReference(const std::string & ee_name) { return references_.at(ee_name); }
  };

} // namespace simple_mpc
