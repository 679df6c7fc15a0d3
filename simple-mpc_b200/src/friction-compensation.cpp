#include "simple-mpc/friction-compensation.hpp"
#include <stdexcept>

namespace simple_mpc
{
  static VectorXd tail_or_zero(const VectorXd & v, int nv, int n)
  {
    // an unset coefficient vector means "no friction" (pinocchio initialises both to zero)
    VectorXd out((size_t)n, 0.0);
    if ((int)v.size() == nv)
      for (int i = 0; i < n; i++)
        out[(size_t)i] = v[(size_t)(nv - n + i)];
    else if (!v.empty())
      throw std::runtime_error("FrictionCompensation: model.friction / model.damping must have nv entries");
    return out;
  }

  FrictionCompensation::FrictionCompensation(const Model & model, const bool with_free_flyer)
  {
    nu_ = with_free_flyer ? model.nv - 6 : model.nv;
    dry_friction_ = tail_or_zero(model.friction, model.nv, nu_);
    viscuous_friction_ = tail_or_zero(model.damping, model.nv, nu_);
  }

  void FrictionCompensation::computeFriction(const VectorXd & velocity, VectorXd & torque)
  {
    if ((int)velocity.size() != nu_)
      throw std::runtime_error("Velocity has wrong size");
    if ((int)torque.size() != nu_)
      throw std::runtime_error("Torque has wrong size");
    for (size_t i = 0; i < (size_t)nu_; i++)
      torque[i] += viscuous_friction_[i] * velocity[i] + dry_friction_[i] * signFunction(velocity[i]);
  }

  double FrictionCompensation::signFunction(double x) { return (x > 0) - (x < 0); }
} // namespace simple_mpc
