// interpolator.cpp — see include/simple-mpc/interpolator.hpp. Semantics of reference src/interpolator.cpp:5-78:
// step = floor(delay / timestep), progress = fraction of the step; beyond the last knot the last value is returned;
// configurations are interpolated on the manifold, q0 (+) progress (q1 (-) q0) (what pinocchio::interpolate computes).
#include "simple-mpc/interpolator.hpp"
#include <algorithm>

namespace simple_mpc
{
  namespace
  {
    void locate(double delay, double timestep, size_t & step_nb, double & step_progress)
    {
      step_nb = static_cast<size_t>(delay / timestep);
      step_progress = (delay - (double)step_nb * timestep) / timestep;
    }
  } // namespace

  void Interpolator::interpolateConfiguration(const double delay, const double timestep, const std::vector<VectorXd> & qs, VectorXd & q_interp)
  {
    if (qs.empty() || (int)qs[0].size() != model_.nq)
      throw std::runtime_error("Configuration is not of the right size");
    size_t step_nb;
    double step_progress;
    locate(delay, timestep, step_nb, step_progress);
    if (step_nb >= qs.size() - 1)
      q_interp = qs.back();
    else
      q_interp = interpolate(model_, qs[step_nb], qs[step_nb + 1], step_progress);
  }

  void Interpolator::interpolateState(const double delay, const double timestep, const std::vector<VectorXd> & xs, VectorXd & x_interp)
  {
    const size_t nq = (size_t)model_.nq, nv = (size_t)model_.nv;
    if (xs.empty() || xs[0].size() != nq + nv)
      throw std::runtime_error("State is not of the right size");
    size_t step_nb;
    double step_progress;
    locate(delay, timestep, step_nb, step_progress);
    if (step_nb >= xs.size() - 1)
    {
      x_interp = xs.back();
      return;
    }
    x_interp = VectorXd(nq + nv, 0.0);
    x_interp.setSegment(0, interpolate(model_, xs[step_nb].head(nq), xs[step_nb + 1].head(nq), step_progress));
    for (size_t k = 0; k < nv; k++)
      x_interp[nq + k] = xs[step_nb + 1][nq + k] * step_progress + xs[step_nb][nq + k] * (1. - step_progress);
  }

  void Interpolator::interpolateLinear(const double delay, const double timestep, const std::vector<VectorXd> & vs, VectorXd & v_interp)
  {
    if (vs.empty())
      throw std::runtime_error("interpolateLinear: empty trajectory");
    size_t step_nb;
    double step_progress;
    locate(delay, timestep, step_nb, step_progress);
    if (step_nb >= vs.size() - 1)
    {
      v_interp = vs.back();
      return;
    }
    v_interp = VectorXd(vs[step_nb].size(), 0.0);
    for (size_t k = 0; k < v_interp.size(); k++)
      v_interp[k] = vs[step_nb + 1][k] * step_progress + vs[step_nb][k] * (1. - step_progress);
  }

  void Interpolator::interpolateContacts(const double delay, const double timestep, const std::vector<std::vector<bool>> & cs, std::vector<bool> & c_interp)
  {
    if (cs.empty())
      throw std::runtime_error("interpolateContacts: empty trajectory");
    size_t step_nb = static_cast<size_t>(delay / timestep);
    step_nb = std::clamp(step_nb, (size_t)0, cs.size() - 1);
    c_interp = cs[step_nb];
  }
} // namespace simple_mpc
