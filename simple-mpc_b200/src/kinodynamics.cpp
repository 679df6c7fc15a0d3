// kinodynamics.cpp — see include/simple-mpc/kinodynamics.hpp
#include "simple-mpc/kinodynamics.hpp"

namespace simple_mpc
{

  KinodynamicsOCP::KinodynamicsOCP(const KinodynamicsSettings & settings, const RobotModelHandler & model_handler)
  : Base(model_handler)
  , settings_(settings)
  {
    nu_ = nv_ - 6 + settings_.force_size * (int)model_handler_.getFeetNb();
    if (nu_ > SMPC_MAX_NU)
      throw std::runtime_error("control dimension exceeds SMPC_MAX_NU");
    x0_ = model_handler_.getReferenceState();
    control_ref_ = VectorXd::Zero((size_t)nu_);
  }

  // The stage recipe of reference src/kinodynamics.cpp:40-152, recorded instead of instantiated:
  //   costs       state_cost, control_cost, centroidal_cost, centroidal_derivative_cost, <foot>_pose_cost per foot
  //   dynamics    KinodynamicsFwdDynamics in a semi-implicit Euler integrator
  //   constraints joint-limit box; per foot in contact: [cone], frame velocity, [landing height]
  StageModel KinodynamicsOCP::createStage(
    const std::map<std::string, bool> & contact_phase, const std::map<std::string, SE3> & contact_pose,
    const std::map<std::string, VectorXd> & contact_force, const std::map<std::string, bool> & land_constraint)
  {
    StageModel stm;
    // Contact states in HANDLER (feet registration) order, like the contact / landing masks the kernels use and like the
    // other two handlers. The reference's kinodynamics handler collects them in std::map (alphabetical) order
    // (src/kinodynamics.cpp:45-49) and then pairs them with frame ids in handler order; the two orders coincide whenever the
    // feet are registered alphabetically (every example of the reference), otherwise the reference's pairing is a latent bug
    // that is deliberately not reproduced.
    for (auto const & name : model_handler_.getFeetFrameNames())
      stm.contact_states.push_back(contact_phase.at(name));
    computeControlFromForces(contact_force);

    stm.cost_.components_ = {"state_cost", "control_cost", "centroidal_cost", "centroidal_derivative_cost"};
    const VectorXd & xr = model_handler_.getReferenceState();
    for (size_t k = 0; k < xr.size(); k++)
      stm.slot.x_ref[k] = xr[k];
    for (int k = 0; k < nu_; k++)
      stm.slot.u_ref[k] = control_ref_[(size_t)k];
    if (settings_.kinematics_limits)
      stm.constraint_names.push_back("joint_limits");
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
      if (contact_phase.at(name))
      {
        stm.slot.contact_mask |= 1u << f;
        if (settings_.force_cone)
          stm.constraint_names.push_back(name + (settings_.force_size == 6 ? "_wrench_cone" : "_friction_cone"));
        stm.constraint_names.push_back(name + "_frame_velocity");
        if (settings_.force_size == 3 && settings_.land_cstr && land_constraint.at(name))
        {
          stm.slot.land_mask |= 1u << f;
          stm.constraint_names.push_back(name + "_landing_height");
        }
      }
    }
    return stm;
  }

  void KinodynamicsOCP::setReferencePose(const std::size_t t, const std::string & ee_name, const SE3 & pose_ref)
  {
    StageModel & s = stageAt(t);
    const size_t f = footIndex(ee_name);
    if (settings_.force_size == 6)
      for (int k = 0; k < 9; k++)
        s.slot.foot_ref[f][k] = pose_ref.R[(size_t)k];
    for (int k = 0; k < 3; k++)
      s.slot.foot_ref[f][9 + k] = pose_ref.p[(size_t)k];
  }

  void KinodynamicsOCP::setReferencePoses(const std::size_t t, const std::map<std::string, SE3> & pose_refs)
  {
    if (pose_refs.size() != model_handler_.getFeetNb())
      throw std::runtime_error("pose_refs size does not match number of end effectors");
    for (auto const & ee_name : model_handler_.getFeetFrameNames())
      setReferencePose(t, ee_name, pose_refs.at(ee_name));
  }

  void KinodynamicsOCP::setTerminalReferencePose(const std::string & ee_name, const SE3 &)
  {
    // the reference looks up "<foot>_pose_cost" in the TERMINAL cost stack, which only holds state_cost and
    // centroidal_cost (src/kinodynamics.cpp:352-364): aligator's getComponent throws for the missing key
    throw std::runtime_error("terminal cost has no component '" + ee_name + "_pose_cost'");
  }

  const SE3 KinodynamicsOCP::getReferencePose(const std::size_t t, const std::string & ee_name)
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

  void KinodynamicsOCP::computeControlFromForces(const std::map<std::string, VectorXd> & force_refs)
  {
    for (std::size_t i = 0; i < model_handler_.getFeetNb(); i++)
    {
      const VectorXd & f = force_refs.at(model_handler_.getFootFrameName(i));
      if (settings_.force_size != (int)f.size())
        throw std::runtime_error("force size in settings does not match reference force size");
      control_ref_.setSegment(i * (size_t)settings_.force_size, f);
    }
  }

  void KinodynamicsOCP::setReferenceForces(const std::size_t i, const std::map<std::string, VectorXd> & force_refs)
  {
    computeControlFromForces(force_refs);
    setReferenceControl(i, control_ref_);
  }

  void KinodynamicsOCP::setReferenceForce(const std::size_t i, const std::string & ee_name, const ConstVectorRef & force_ref)
  {
    control_ref_.setSegment(footIndex(ee_name) * (size_t)settings_.force_size, force_ref);
    setReferenceControl(i, control_ref_);
  }

  const VectorXd KinodynamicsOCP::getReferenceForce(const std::size_t i, const std::string & ee_name)
  {
    return getReferenceControl(i).segment(footIndex(ee_name) * (size_t)settings_.force_size, (size_t)settings_.force_size);
  }

  const VectorXd KinodynamicsOCP::getVelocityBase(const std::size_t t) { return getReferenceState(t).segment((size_t)nq_, 6); }

  void KinodynamicsOCP::setVelocityBase(const std::size_t t, const ConstVectorRef & velocity_base)
  {
    if (velocity_base.size() != 6)
      throw std::runtime_error("velocity_base size should be 6");
    x0_ = getReferenceState(t);
    x0_.setSegment((size_t)nq_, velocity_base);
    setReferenceState(t, x0_);
  }

  const VectorXd KinodynamicsOCP::getPoseBase(const std::size_t t) { return getReferenceState(t).head(7); }

  void KinodynamicsOCP::setPoseBase(const std::size_t t, const ConstVectorRef & pose_base)
  {
    if (pose_base.size() != 7)
      throw std::runtime_error("pose_base size should be 7");
    x0_ = getReferenceState(t);
    x0_.setSegment(0, pose_base);
    setReferenceState(t, x0_);
  }

  const VectorXd KinodynamicsOCP::getProblemState(const RobotDataHandler & data_handler) { return data_handler.getState(); }

  void KinodynamicsOCP::setReferenceState(const std::size_t t, const ConstVectorRef & x_ref)
  {
    if ((int)x_ref.size() != nq_ + nv_)
      throw std::runtime_error("x_ref not of the right size");
    StageModel & s = stageAt(t);
    for (size_t k = 0; k < x_ref.size(); k++)
      s.slot.x_ref[k] = x_ref[k];
  }

  const VectorXd KinodynamicsOCP::getReferenceState(const std::size_t t)
  {
    const StageModel & s = stageAt(t);
    return VectorXd(s.slot.x_ref, s.slot.x_ref + nq_ + nv_);
  }

  CostStack KinodynamicsOCP::createTerminalCost()
  {
    // state_cost + centroidal_cost with 10 w_cent (reference src/kinodynamics.cpp:352-364)
    const VectorXd & xr = model_handler_.getReferenceState();
    if (problem_)
      for (size_t k = 0; k < xr.size(); k++)
        problem_->term_.x_ref[k] = xr[k];
    CostStack cs;
    cs.components_ = {"state_cost", "centroidal_cost"};
    return cs;
  }

  void KinodynamicsOCP::createTerminalConstraint(const Vector3d & com_ref)
  {
    if (!problem_initialized_)
      throw std::runtime_error("Create problem first!");
    // DCMPositionResidual(com_ref, tau = sqrt(com_ref[2] / 9.81)) = 0 ; tau is fixed here, updateTerminalConstraint
    // (DCMres->setReference) only moves com_ref (reference src/kinodynamics.cpp:366-388)
    for (int k = 0; k < 3; k++)
      problem_->term_.com_ref[k] = com_ref[(size_t)k];
    problem_->term_.tau = std::sqrt(com_ref[2] / 9.81);
    problem_->term_cstrs_ = {"dcm_position"};
    terminal_constraint_ = true;
  }

  void KinodynamicsOCP::updateTerminalConstraint(const Vector3d & com_ref)
  {
    if (terminal_constraint_)
      for (int k = 0; k < 3; k++)
        problem_->term_.com_ref[k] = com_ref[(size_t)k];
  }

  smpc_ocp_desc KinodynamicsOCP::toOcpDesc() const
  {
    const size_t ndx = (size_t)ndx_, nu = (size_t)nu_, fs = (size_t)settings_.force_size, nj = (size_t)(nv_ - 6);
    if (settings_.w_x.rows() != ndx || settings_.w_x.cols() != ndx || settings_.w_u.rows() != nu || settings_.w_u.cols() != nu)
      throw std::runtime_error("w_x / w_u do not have the size of the state tangent / control");
    if (settings_.w_cent.rows() != 6 || settings_.w_cent.cols() != 6 || settings_.w_centder.rows() != 6 || settings_.w_centder.cols() != 6)
      throw std::runtime_error("w_cent / w_centder must be 6 x 6");
    if (settings_.w_frame.rows() != fs || settings_.w_frame.cols() != fs)
      throw std::runtime_error("w_frame must be force_size x force_size");
    if (settings_.kinematics_limits && (settings_.qmin.size() != nj || settings_.qmax.size() != nj))
      throw std::runtime_error("qmin / qmax must have nv - 6 entries");
    smpc_ocp_desc d;
    memset(&d, 0, sizeof(d));
    d.kind = SMPC_OCP_KINODYNAMICS;
    d.horizon = problem_ ? (int)problem_->numSteps() : 0;
    d.force_size = settings_.force_size;
    d.kinematics_limits = settings_.kinematics_limits;
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
    d.w_centder = settings_.w_centder.data();
    d.w_frame = settings_.w_frame.data();
    d.qmin = settings_.qmin.data();
    d.qmax = settings_.qmax.data();
    return d;
  }

} // namespace simple_mpc
