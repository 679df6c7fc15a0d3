// fulldynamics.cpp — see include/simple-mpc/fulldynamics.hpp
#include "simple-mpc/fulldynamics.hpp"

namespace simple_mpc
{

  FullDynamicsOCP::FullDynamicsOCP(const FullDynamicsSettings & settings, const RobotModelHandler & model_handler)
  : Base(model_handler)
  , settings_(settings)
  {
    // nu = nv - 6: joint torques through the actuation matrix [0; I] (reference src/fulldynamics.cpp:35-37)
    x0_ = model_handler_.getReferenceState();
    control_ref_ = VectorXd::Zero((size_t)nu_);
  }

  // The stage recipe of reference src/fulldynamics.cpp:78-214:
  //   costs       state_cost, control_cost, centroidal_cost, <foot>_pose_cost per foot, <foot>_force_cost per stance foot
  //   dynamics    MultibodyConstraintFwdDynamics (rigid contacts of the stance feet) in a semi-implicit Euler integrator
  //   constraints torque box, joint-limit box; per stance foot [cone] and, on a landing edge, velocity (+ height)
  StageModel FullDynamicsOCP::createStage(
    const std::map<std::string, bool> & contact_phase, const std::map<std::string, SE3> & contact_pose,
    const std::map<std::string, VectorXd> & contact_force, const std::map<std::string, bool> & land_constraint)
  {
    StageModel stm;
    stm.cost_.components_ = {"state_cost", "control_cost", "centroidal_cost"};
    const VectorXd & xr = model_handler_.getReferenceState();
    for (size_t k = 0; k < xr.size(); k++)
      stm.slot.x_ref[k] = xr[k];
    const auto & names = model_handler_.getFeetFrameNames();
    for (size_t f = 0; f < names.size(); f++)
    {
      const std::string & name = names[f];
      stm.cost_.components_.push_back(name + "_pose_cost");
      const SE3 & pose = contact_pose.at(name);
      for (int k = 0; k < 9; k++)
        stm.slot.foot_ref[f][k] = pose.R[(size_t)k];
      for (int k = 0; k < 3; k++)
        stm.slot.foot_ref[f][9 + k] = pose.p[(size_t)k];
      stm.slot.land_z[f] = pose.p[2]; // target of the landing-edge height equality, fixed at stage creation
      stm.contact_states.push_back(contact_phase.at(name));
      if (contact_phase.at(name))
        stm.slot.contact_mask |= 1u << f;
    }
    for (size_t f = 0; f < names.size(); f++)
    {
      const std::string & name = names[f];
      const VectorXd & fr = contact_force.at(name);
      if ((int)fr.size() != settings_.force_size)
        throw std::runtime_error("Reference forces do not have the right dimension");
      if (contact_phase.at(name))
      {
        stm.cost_.components_.push_back(name + "_force_cost");
        for (size_t k = 0; k < fr.size(); k++)
          stm.slot.f_ref[f][k] = fr[k];
      }
    }
    if (settings_.torque_limits)
      stm.constraint_names.push_back("torque_limits");
    if (settings_.kinematics_limits)
      stm.constraint_names.push_back("joint_limits");
    for (size_t f = 0; f < names.size(); f++)
    {
      const std::string & name = names[f];
      if (!contact_phase.at(name))
        continue;
      if (settings_.force_cone)
        stm.constraint_names.push_back(name + (settings_.force_size == 6 ? "_wrench_cone" : "_friction_cone"));
      if (settings_.land_cstr && land_constraint.at(name))
      {
        stm.slot.land_mask |= 1u << f;
        stm.constraint_names.push_back(name + "_landing_velocity");
        if (settings_.force_size == 3)
          stm.constraint_names.push_back(name + "_landing_height");
      }
    }
    return stm;
  }

  void FullDynamicsOCP::setReferencePose(const std::size_t t, const std::string & ee_name, const SE3 & pose_ref)
  {
    StageModel & s = stageAt(t);
    const size_t f = footIndex(ee_name);
    if (settings_.force_size == 6)
      for (int k = 0; k < 9; k++)
        s.slot.foot_ref[f][k] = pose_ref.R[(size_t)k];
    for (int k = 0; k < 3; k++)
      s.slot.foot_ref[f][9 + k] = pose_ref.p[(size_t)k];
  }

  void FullDynamicsOCP::setReferencePoses(const std::size_t t, const std::map<std::string, SE3> & pose_refs)
  {
    if (pose_refs.size() != model_handler_.getFeetNb())
      throw std::runtime_error("pose_refs size does not match number of end effectors");
    for (auto const & ee_name : model_handler_.getFeetFrameNames())
      setReferencePose(t, ee_name, pose_refs.at(ee_name));
  }

  void FullDynamicsOCP::setTerminalReferencePose(const std::string & ee_name, const SE3 &)
  {
    throw std::runtime_error("terminal cost has no component '" + ee_name + "_pose_cost'");
  }

  const SE3 FullDynamicsOCP::getReferencePose(const std::size_t t, const std::string & ee_name)
  {
    const StageModel & s = stageAt(t);
    const size_t f = footIndex(ee_name);
    SE3 ref = SE3::Identity();
    if (settings_.force_size == 6)
      for (int k = 0; k < 9; k++)
        ref.R[(size_t)k] = s.slot.foot_ref[f][k];
    for (int k = 0; k < 3; k++)
      ref.p[(size_t)k] = s.slot.foot_ref[f][9 + k];
    return ref;
  }

  void FullDynamicsOCP::setReferenceForces(const std::size_t t, const std::map<std::string, VectorXd> & force_refs)
  {
    if (force_refs.size() != model_handler_.getFeetNb())
      throw std::runtime_error("force_refs size does not match number of end effectors");
    for (auto const & ee_name : model_handler_.getFeetFrameNames())
      setReferenceForce(t, ee_name, force_refs.at(ee_name));
  }

  void FullDynamicsOCP::setReferenceForce(const std::size_t t, const std::string & ee_name, const ConstVectorRef & force_ref)
  {
    StageModel & s = stageAt(t);
    if ((int)force_ref.size() != settings_.force_size)
      throw std::runtime_error("Reference forces do not have the right dimension");
    const size_t f = footIndex(ee_name);
    for (size_t k = 0; k < force_ref.size(); k++)
      s.slot.f_ref[f][k] = force_ref[k];
  }

  const VectorXd FullDynamicsOCP::getReferenceForce(const std::size_t t, const std::string & ee_name)
  {
    const StageModel & s = stageAt(t);
    const size_t f = footIndex(ee_name);
    return VectorXd(s.slot.f_ref[f], s.slot.f_ref[f] + settings_.force_size);
  }

  const VectorXd FullDynamicsOCP::getVelocityBase(const std::size_t t) { return getReferenceState(t).segment((size_t)nq_, 6); }

  void FullDynamicsOCP::setVelocityBase(const std::size_t t, const ConstVectorRef & velocity_base)
  {
    if (velocity_base.size() != 6)
      throw std::runtime_error("velocity_base size should be 6");
    x0_ = getReferenceState(t);
    x0_.setSegment((size_t)nq_, velocity_base);
    setReferenceState(t, x0_);
  }

  const VectorXd FullDynamicsOCP::getPoseBase(const std::size_t t) { return getReferenceState(t).head(7); }

  void FullDynamicsOCP::setPoseBase(const std::size_t t, const ConstVectorRef & pose_base)
  {
    if (pose_base.size() != 7)
      throw std::runtime_error("pose_base size should be 7");
    x0_ = getReferenceState(t);
    x0_.setSegment(0, pose_base);
    setReferenceState(t, x0_);
  }

  const VectorXd FullDynamicsOCP::getProblemState(const RobotDataHandler & data_handler) { return data_handler.getState(); }

  void FullDynamicsOCP::setReferenceState(const std::size_t t, const ConstVectorRef & x_ref)
  {
    if ((int)x_ref.size() != nq_ + nv_)
      throw std::runtime_error("x_ref not of the right size");
    StageModel & s = stageAt(t);
    for (size_t k = 0; k < x_ref.size(); k++)
      s.slot.x_ref[k] = x_ref[k];
  }

  const VectorXd FullDynamicsOCP::getReferenceState(const std::size_t t)
  {
    const StageModel & s = stageAt(t);
    return VectorXd(s.slot.x_ref, s.slot.x_ref + nq_ + nv_);
  }

  CostStack FullDynamicsOCP::createTerminalCost()
  {
    // state_cost + centroidal_cost with 10 w_cent (reference src/fulldynamics.cpp:418-430)
    const VectorXd & xr = model_handler_.getReferenceState();
    if (problem_)
      for (size_t k = 0; k < xr.size(); k++)
        problem_->term_.x_ref[k] = xr[k];
    CostStack cs;
    cs.components_ = {"state_cost", "centroidal_cost"};
    return cs;
  }

  void FullDynamicsOCP::createTerminalConstraint(const Vector3d & com_ref)
  {
    if (!problem_initialized_)
      throw std::runtime_error("Create problem first!");
    // tau is fixed here (reference src/fulldynamics.cpp:432-445); updateTerminalConstraint only moves com_ref
    for (int k = 0; k < 3; k++)
      problem_->term_.com_ref[k] = com_ref[(size_t)k];
    problem_->term_.tau = std::sqrt(com_ref[2] / 9.81);
    problem_->term_cstrs_ = {"dcm_position"};
    terminal_constraint_ = true;
  }

  void FullDynamicsOCP::updateTerminalConstraint(const Vector3d & com_ref)
  {
    if (terminal_constraint_)
      for (int k = 0; k < 3; k++)
        problem_->term_.com_ref[k] = com_ref[(size_t)k];
  }

  smpc_ocp_desc FullDynamicsOCP::toOcpDesc() const
  {
    smpc_ocp_desc d;
    memset(&d, 0, sizeof(d));
    d.kind = SMPC_OCP_FULLDYNAMICS;
    d.horizon = problem_ ? (int)problem_->numSteps() : 0;
    d.force_size = settings_.force_size;
    d.kinematics_limits = settings_.kinematics_limits;
    d.torque_limits = settings_.torque_limits;
    d.force_cone = settings_.force_cone;
    d.land_cstr = settings_.land_cstr;
    d.terminal_constraint = terminal_constraint_;
    d.timestep = settings_.timestep;
    for (int k = 0; k < 3; k++)
      d.gravity[k] = settings_.gravity[(size_t)k];
    d.mu = settings_.mu;
    d.Lfoot = settings_.Lfoot;
    d.Wfoot = settings_.Wfoot;
    d.w_x = settings_.w_x.data();
    d.w_u = settings_.w_u.data();
    d.w_cent = settings_.w_cent.data();
    d.w_forces = settings_.w_forces.data();
    d.w_frame = settings_.w_frame.data();
    d.qmin = settings_.qmin.data();
    d.qmax = settings_.qmax.data();
    d.umin = settings_.umin.data();
    d.umax = settings_.umax.data();
    for (size_t k = 0; k < 6 && k < settings_.Kp_correction.size(); k++)
      d.Kp_correction[k] = settings_.Kp_correction[k];
    for (size_t k = 0; k < 6 && k < settings_.Kd_correction.size(); k++)
      d.Kd_correction[k] = settings_.Kd_correction[k];
    return d;
  }

} // namespace simple_mpc
