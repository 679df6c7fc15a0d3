// centroidal-dynamics.cpp — see include/simple-mpc/centroidal-dynamics.hpp
#include "simple-mpc/centroidal-dynamics.hpp"

namespace simple_mpc
{

  CentroidalOCP::CentroidalOCP(const CentroidalSettings & settings, const RobotModelHandler & model_handler)
  : Base(model_handler)
  , settings_(settings)
  {
    nu_ = (int)model_handler_.getFeetNb() * settings_.force_size;
    if (nu_ > SMPC_MAX_NU)
      throw std::runtime_error("control dimension exceeds SMPC_MAX_NU");
    x0_ = VectorXd::Zero(9);
    control_ref_ = VectorXd::Zero((size_t)nu_);
  }

  // The stage recipe of reference src/centroidal-dynamics.cpp:39-106:
  //   costs       com_cost, control_cost, linear_mom_cost, angular_mom_cost, linear_acc_cost, angular_acc_cost
  //   dynamics    CentroidalFwdDynamics in an explicit Euler integrator; contact points = contact_pose translations
  //   constraints per foot in contact: wrench cone (force_size 6) or friction cone (force_size 3)
  StageModel CentroidalOCP::createStage(
    const std::map<std::string, bool> & contact_phase, const std::map<std::string, SE3> & contact_pose,
    const std::map<std::string, VectorXd> & contact_force, const std::map<std::string, bool> &)
  {
    StageModel stm;
    computeControlFromForces(contact_force);
    stm.cost_.components_ = {"com_cost", "control_cost", "linear_mom_cost", "angular_mom_cost", "linear_acc_cost", "angular_acc_cost"};
    for (int k = 0; k < 3; k++)
      stm.slot.x_ref[k] = com_ref_[(size_t)k]; // com reference; momentum references start at zero (:63-65)
    for (int k = 0; k < nu_; k++)
      stm.slot.u_ref[k] = control_ref_[(size_t)k];
    const auto & names = model_handler_.getFeetFrameNames();
    for (size_t f = 0; f < names.size(); f++)
    {
      const std::string & name = names[f];
      stm.contact_states.push_back(contact_phase.at(name));
      const SE3 & pose = contact_pose.at(name);
      for (int k = 0; k < 9; k++)
        stm.slot.foot_ref[f][k] = pose.R[(size_t)k];
      for (int k = 0; k < 3; k++)
        stm.slot.foot_ref[f][9 + k] = pose.p[(size_t)k];
      if (contact_phase.at(name))
      {
        stm.slot.contact_mask |= 1u << f;
        stm.constraint_names.push_back(name + (settings_.force_size == 6 ? "_wrench_cone" : "_friction_cone"));
      }
    }
    return stm;
  }

  void CentroidalOCP::computeControlFromForces(const std::map<std::string, VectorXd> & force_refs)
  {
    for (std::size_t i = 0; i < model_handler_.getFeetNb(); i++)
    {
      const VectorXd & f = force_refs.at(model_handler_.getFootFrameName(i));
      if (settings_.force_size != (int)f.size())
        throw std::runtime_error("force size in settings does not match reference force size");
      control_ref_.setSegment(i * (size_t)settings_.force_size, f);
    }
  }

  void CentroidalOCP::setReferencePoses(const std::size_t t, const std::map<std::string, SE3> & pose_refs)
  {
    stageAt(t);
    if (pose_refs.size() != model_handler_.getFeetNb())
      throw std::runtime_error("pose_refs size does not match number of end effectors");
    for (auto const & pose : pose_refs)
      setReferencePose(t, pose.first, pose.second);
  }

  void CentroidalOCP::setReferencePose(const std::size_t t, const std::string & ee_name, const SE3 & pose_ref)
  {
    // one record feeds the dynamics and both acceleration residuals (the reference updates three contact maps)
    StageModel & s = stageAt(t);
    const size_t f = footIndex(ee_name);
    for (int k = 0; k < 3; k++)
      s.slot.foot_ref[f][9 + k] = pose_ref.p[(size_t)k];
  }

  const SE3 CentroidalOCP::getReferencePose(const std::size_t t, const std::string & ee_name)
  {
    const StageModel & s = stageAt(t);
    const size_t f = footIndex(ee_name);
    SE3 pose = SE3::Identity();
    for (int k = 0; k < 3; k++)
      pose.p[(size_t)k] = s.slot.foot_ref[f][9 + k];
    return pose;
  }

  void CentroidalOCP::setReferenceForces(const std::size_t t, const std::map<std::string, VectorXd> & force_refs)
  {
    computeControlFromForces(force_refs);
    setReferenceControl(t, control_ref_);
  }

  void CentroidalOCP::setReferenceForce(const std::size_t t, const std::string & ee_name, const ConstVectorRef & force_ref)
  {
    control_ref_.setSegment(footIndex(ee_name) * (size_t)settings_.force_size, force_ref);
    setReferenceControl(t, control_ref_);
  }

  const VectorXd CentroidalOCP::getReferenceForce(const std::size_t t, const std::string & ee_name)
  {
    return getReferenceControl(t).segment(footIndex(ee_name) * (size_t)settings_.force_size, (size_t)settings_.force_size);
  }

  const VectorXd CentroidalOCP::getVelocityBase(const std::size_t t)
  {
    const StageModel & s = stageAt(t);
    VectorXd v(6, 0.0);
    for (size_t k = 0; k < 6; k++)
      v[k] = s.slot.x_ref[3 + k] / model_handler_.getMass();
    return v;
  }

  void CentroidalOCP::setVelocityBase(const std::size_t t, const ConstVectorRef & velocity_base)
  {
    if (velocity_base.size() != 6)
      throw std::runtime_error("velocity_base not of the right size");
    StageModel & s = stageAt(t);
    // momentum references = velocity * mass for BOTH parts (reference src/centroidal-dynamics.cpp:237-238)
    for (size_t k = 0; k < 6; k++)
      s.slot.x_ref[3 + k] = velocity_base[k] * model_handler_.getMass();
  }

  const VectorXd CentroidalOCP::getPoseBase(const std::size_t t)
  {
    const StageModel & s = stageAt(t);
    return VectorXd(s.slot.x_ref, s.slot.x_ref + 3);
  }

  void CentroidalOCP::setPoseBase(const std::size_t t, const ConstVectorRef & pose_base)
  {
    if (pose_base.size() != 3)
      throw std::runtime_error("pose_base not of the right size");
    StageModel & s = stageAt(t);
    for (size_t k = 0; k < 3; k++)
      s.slot.x_ref[k] = com_ref_[k] = pose_base[k];
  }

  const VectorXd CentroidalOCP::getProblemState(const RobotDataHandler & data_handler) { return data_handler.getCentroidalState(); }

  void CentroidalOCP::setReferenceState(const std::size_t t, const ConstVectorRef & x_ref)
  {
    if (x_ref.size() != 9)
      throw std::runtime_error("x_ref not of the right size");
    setPoseBase(t, x_ref.head(3));
    setVelocityBase(t, x_ref.tail(6));
  }

  const VectorXd CentroidalOCP::getReferenceState(const std::size_t t)
  {
    x0_.setSegment(0, getPoseBase(t));
    x0_.setSegment(3, getVelocityBase(t));
    return x0_;
  }

  CostStack CentroidalOCP::createTerminalCost()
  {
    // linear_mom_cost + angular_mom_cost with zero references (reference src/centroidal-dynamics.cpp:306-316)
    CostStack cs;
    cs.components_ = {"linear_mom_cost", "angular_mom_cost"};
    return cs;
  }

  void CentroidalOCP::createTerminalConstraint(const Vector3d &)
  {
    if (!problem_initialized_)
      throw std::runtime_error("Create problem first!");
    // intentionally disabled in the reference (src/centroidal-dynamics.cpp:318-328)
    terminal_constraint_ = true;
  }

  void CentroidalOCP::updateTerminalConstraint(const Vector3d &) {}

  smpc_ocp_desc CentroidalOCP::toOcpDesc() const
  {
    const size_t nu = (size_t)nu_;
    if (settings_.w_u.rows() != nu || settings_.w_u.cols() != nu)
      throw std::runtime_error("w_u does not have the size of the control");
    for (const MatrixXd * W : {&settings_.w_com, &settings_.w_linear_mom, &settings_.w_angular_mom, &settings_.w_linear_acc, &settings_.w_angular_acc})
      if (W->rows() != 3 || W->cols() != 3)
        throw std::runtime_error("centroidal weights must be 3 x 3");
    smpc_ocp_desc d;
    memset(&d, 0, sizeof(d));
    d.kind = SMPC_OCP_CENTROIDAL;
    d.horizon = problem_ ? (int)problem_->numSteps() : 0;
    d.force_size = settings_.force_size;
    d.timestep = settings_.timestep;
    for (int k = 0; k < 3; k++)
      d.gravity[k] = settings_.gravity[(size_t)k];
    d.mu = settings_.mu;
    d.Lfoot = settings_.Lfoot;
    d.Wfoot = settings_.Wfoot;
    d.w_u = settings_.w_u.data();
    d.w_com = settings_.w_com.data();
    d.w_linear_mom = settings_.w_linear_mom.data();
    d.w_angular_mom = settings_.w_angular_mom.data();
    d.w_linear_acc = settings_.w_linear_acc.data();
    d.w_angular_acc = settings_.w_angular_acc.data();
    return d;
  }

} // namespace simple_mpc
