///////////////////////////////////////////////////////////////////////////////
// friction-compensation.hpp — FrictionCompensation with the reference's interface (reference
// include/simple-mpc/friction-compensation.hpp:27-55, src/friction-compensation.cpp:6-38): dry + viscous joint friction added
// to a torque command after the MPC / inverse-dynamics step. Host-side, O(nu): there is nothing to put on the GPU.
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/model.hpp"

namespace simple_mpc
{
  class FrictionCompensation
  {
  public:
    /// model.friction / model.damping hold one coefficient per velocity dof (pinocchio's convention); with a free flyer the
    /// first six are ignored
    FrictionCompensation(const Model & model, const bool with_free_flyer = true);
    /// torque += viscous * velocity + dry * sign(velocity); throws std::runtime_error on a size mismatch
    void computeFriction(const VectorXd & velocity, VectorXd & torque);
    static double signFunction(double x);

    int nu_;
    VectorXd dry_friction_;
    VectorXd viscuous_friction_;
  };
} // namespace simple_mpc
