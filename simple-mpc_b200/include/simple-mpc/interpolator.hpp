///////////////////////////////////////////////////////////////////////////////
// interpolator.hpp — Interpolator with the reference's interface (reference include/simple-mpc/interpolator.hpp,
// src/interpolator.cpp:5-78): interpolation of configuration / state / linear / contact trajectories between two
// MPC knots. Host-side, single trajectory; the batched device version is smpc_interpolate_plan (include/smpc.h).
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/model.hpp"

namespace simple_mpc
{
  class Interpolator
  {
  public:
    explicit Interpolator(const Model & model) : model_(model) {}

    void interpolateConfiguration(const double delay, const double timestep, const std::vector<VectorXd> & qs, VectorXd & q_interp);
    void interpolateState(const double delay, const double timestep, const std::vector<VectorXd> & xs, VectorXd & x_interp);
    void interpolateLinear(const double delay, const double timestep, const std::vector<VectorXd> & vs, VectorXd & v_interp);
    void interpolateContacts(const double delay, const double timestep, const std::vector<std::vector<bool>> & cs, std::vector<bool> & c_interp);

  private:
    Model model_; // copied, like the reference
  };
} // namespace simple_mpc
