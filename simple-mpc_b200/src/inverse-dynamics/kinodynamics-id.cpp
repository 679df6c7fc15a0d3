#include "simple-mpc/inverse-dynamics/kinodynamics-id.hpp"

namespace simple_mpc
{

  smpc_id_settings KinodynamicsID::lower(const Settings & s)
  {
    smpc_id_settings o;
    smpc_id_default_settings(&o);
    o.friction_coefficient = s.friction_coefficient;
    o.contact_weight_ratio_max = s.contact_weight_ratio_max;
    o.contact_weight_ratio_min = s.contact_weight_ratio_min;
    o.kp_base = s.kp_base;
    o.kp_posture = s.kp_posture;
    o.kp_contact = s.kp_contact;
    o.w_base = s.w_base;
    o.w_posture = s.w_posture;
    o.w_contact_motion = s.w_contact_motion;
    o.w_contact_force = s.w_contact_force;
    o.contact_motion_equality = s.contact_motion_equality ? 1 : 0;
    return o;
  }

  void KinodynamicsID::check(int rc, const char * what) const
  {
    if (rc != 0)
      throw std::runtime_error(std::string(what) + ": " + smpc_last_error());
  }

  // What the reference's constructor reads from the handler (kinodynamics-id.cpp:24-98): foot types and contact points, the
  // position / velocity / effort limits of the actuated joints; the default target and the dry-run solve happen in
  // smpc_id_create.
  void KinodynamicsID::create(double control_dt, const smpc_id_settings & full)
  {
    const Model & model = model_handler_.getModel();
    const size_t na = (size_t)model.nv - 6;
    if (model.lowerPositionLimit.size() != na || model.upperPositionLimit.size() != na || model.effortLimit.size() != na || model.velocityLimit.size() != na)
      throw std::runtime_error("KinodynamicsID: the model needs position, velocity and effort limits for its nv - 6 actuated joints");
    smpc_id_desc d;
    memset(&d, 0, sizeof(d));
    d.control_dt = control_dt;
    for (size_t j = 0; j < na; j++)
    {
      d.qmin[j] = model.lowerPositionLimit[j];
      d.qmax[j] = model.upperPositionLimit[j];
      d.vmax[j] = model.velocityLimit[j];
      d.effort[j] = model.effortLimit[j];
    }
    for (size_t f = 0; f < model_handler_.getFeetNb(); f++)
    {
      const bool quad = model_handler_.getFootType(f) == RobotModelHandler::QUAD;
      d.foot_type[f] = quad ? SMPC_ID_QUAD : SMPC_ID_POINT;
      if (quad)
      {
        const RobotModelHandler::ContactPointsMatrix & P = model_handler_.getQuadFootContactPoints(f);
        for (int k = 0; k < 4; k++)
          for (int e = 0; e < 3; e++)
            d.quad_points[f][k][e] = P[(size_t)(3 * k + e)];
      }
    }
    d.settings = full;
    const smpc_model_desc m = model_handler_.toDesc();
    check(smpc_id_create(&ctx_, &m, &d, 1), "KinodynamicsID");
  }

  KinodynamicsID::KinodynamicsID(const RobotModelHandler & model_handler, double control_dt, const Settings settings)
  : settings_(settings)
  , model_handler_(model_handler)
  {
    create(control_dt, lower(settings_));
  }

  KinodynamicsID::KinodynamicsID(
    const RobotModelHandler & model_handler, double control_dt, const Settings & settings, const smpc_id_settings & full)
  : settings_(settings)
  , model_handler_(model_handler)
  {
    create(control_dt, full);
  }

  KinodynamicsID::~KinodynamicsID() { smpc_id_destroy(ctx_); }

  void KinodynamicsID::setTarget(
    const ConstVectorRef & q_target, const ConstVectorRef & v_target, const ConstVectorRef & a_target,
    const std::vector<bool> & contact_state_target, const std::vector<TargetContactForce> & f_target)
  {
    const Model & model = model_handler_.getModel();
    const size_t nf = model_handler_.getFeetNb();
    if ((int)q_target.size() != model.nq || (int)v_target.size() != model.nv || (int)a_target.size() != model.nv)
      throw std::runtime_error("KinodynamicsID::setTarget: q / v / a have the wrong size");
    if (contact_state_target.size() != nf)
      throw std::runtime_error("KinodynamicsID::setTarget: contact_state_target has the wrong size");
    std::vector<int32_t> cs(nf);
    std::vector<double> ft(6 * nf, 0.0);
    for (size_t f = 0; f < nf; f++)
    {
      cs[f] = contact_state_target[f] ? 1 : 0;
      if (!contact_state_target[f])
        continue;
      const size_t fs = model_handler_.getFootType(f) == RobotModelHandler::QUAD ? 6 : 3;
      if (f_target.size() <= f || f_target[f].size() != fs) // the reference asserts (kinodynamics-id.cpp:159,164)
        throw std::runtime_error("KinodynamicsID::setTarget: f_target has the wrong size");
      for (size_t e = 0; e < fs; e++)
        ft[6 * f + e] = f_target[f][e];
    }
    check(smpc_id_set_target(ctx_, 0, 1, q_target.data(), v_target.data(), a_target.data(), cs.data(), ft.data()), "KinodynamicsID::setTarget");
  }

  void KinodynamicsID::solve(const double t, const ConstVectorRef & q_meas, const ConstVectorRef & v_meas, VectorXd & tau_res)
  {
    const Model & model = model_handler_.getModel();
    if ((int)q_meas.size() != model.nq || (int)v_meas.size() != model.nv)
      throw std::runtime_error("KinodynamicsID::solve: q / v have the wrong size");
    tau_res.resize((size_t)model.nv - 6);
    check(smpc_id_solve(ctx_, t, q_meas.data(), v_meas.data(), tau_res.data()), "KinodynamicsID::solve");
    smpc_id_info info;
    check(smpc_id_get_info(ctx_, 0, 1, &info), "KinodynamicsID::solve");
    if (info.status != 0)
      throw std::runtime_error("KinodynamicsID::solve: QP not solved to tolerance (infeasible target?)");
  }

  void KinodynamicsID::getAccelerations(VectorXd & a)
  {
    a.resize((size_t)model_handler_.getModel().nv);
    check(smpc_id_get_accelerations(ctx_, 0, 1, a.data()), "KinodynamicsID::getAccelerations");
  }

  std::vector<VectorXd> KinodynamicsID::getContactForces()
  {
    const size_t nf = model_handler_.getFeetNb();
    std::vector<double> f(nf * SMPC_ID_MAX_FVAR);
    check(smpc_id_get_forces(ctx_, 0, 1, f.data()), "KinodynamicsID::getContactForces");
    std::vector<VectorXd> out;
    for (size_t k = 0; k < nf; k++)
      out.emplace_back(f.begin() + (long)(k * SMPC_ID_MAX_FVAR), f.begin() + (long)((k + 1) * SMPC_ID_MAX_FVAR));
    return out;
  }

} // namespace simple_mpc
