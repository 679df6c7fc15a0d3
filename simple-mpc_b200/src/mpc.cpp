// mpc.cpp — see include/simple-mpc/mpc.hpp. Follows the driver semantics of reference src/mpc.cpp
// (constructor :19-99, generateCycleHorizon :101-187, iterate :189-218, recedeWithCycle :220-254,
// updateCycleTiming :256-276, updateStepTrackerReferences :278-324, getters :326-392).
#include "simple-mpc/mpc.hpp"
#include <algorithm>

namespace simple_mpc
{
  constexpr std::size_t maxiters = 100; // constructor solve (reference src/mpc.cpp:17)

  // ------------------------------------------------------------------------------------------------
  // SolverApi: the CUDA library is the only implementation the product binds
  // ------------------------------------------------------------------------------------------------
  const SolverApi & SolverApi::cuda()
  {
    static const SolverApi api = {
      smpc_last_error, smpc_default_solver_opts, smpc_problem_dims, smpc_create,      smpc_destroy,  smpc_set_slots,
      smpc_set_terminal, smpc_cycle,             smpc_set_warmstart, smpc_solve,      smpc_get_results, smpc_get_info,
      smpc_get_xdot,   smpc_get_contact_forces,  "libsmpc_cuda (sm_100a)"};
    return api;
  }

  // ------------------------------------------------------------------------------------------------
  // SolverProxDDP
  // ------------------------------------------------------------------------------------------------
  SolverProxDDP::SolverProxDDP(double tol, double mu, std::size_t max_it, const SolverApi & api, int device)
  : target_tol_(tol)
  , mu_init(mu)
  , max_iters(max_it)
  , api_(api)
  , device_(device)
  {
    api_.default_solver_opts(&opts_);
  }

  SolverProxDDP::~SolverProxDDP()
  {
    if (ctx_)
      api_.destroy(ctx_);
  }

  void SolverProxDDP::check(int rc, const char * what)
  {
    if (rc != SMPC_OK)
      throw std::runtime_error(std::string(what) + " failed (" + api_.name + "): " + api_.last_error());
  }

  void SolverProxDDP::setup(const OCPHandler & ocp)
  {
    if (ctx_)
    {
      api_.destroy(ctx_);
      ctx_ = nullptr;
    }
    const smpc_model_desc model = ocp.getModelHandler().toDesc();
    const smpc_ocp_desc desc = ocp.toOcpDesc();
    opts_.tol = target_tol_;
    opts_.mu_init = mu_init;
    opts_.max_iters = (int)max_iters;
    check(api_.problem_dims(&model, &desc, &nx_, &ndx_, &nu_), "problem_dims");
    N_ = desc.horizon;
    check(api_.create(&ctx_, device_, 1, &model, &desc, &opts_), "create");
  }

  bool SolverProxDDP::run(const TrajOptProblem & problem, const std::vector<VectorXd> & xs_init, const std::vector<VectorXd> & us_init)
  {
    if (!ctx_)
      throw std::runtime_error("SolverProxDDP::run called before setup");
    const size_t N = (size_t)N_;
    if (problem.numSteps() != N || xs_init.size() != N + 1 || us_init.size() != N)
      throw std::runtime_error("SolverProxDDP::run: problem / warm-start sizes do not match the set-up horizon");
    // problem description: every stage record and the terminal node (they are what simple-mpc mutates between runs)
    std::vector<smpc_slot> slots(N);
    for (size_t t = 0; t < N; t++)
      slots[t] = problem.stages_[t].slot;
    check(api_.set_slots(ctx_, 0, 1, slots.data()), "set_slots");
    check(api_.set_terminal(ctx_, 0, 1, &problem.term_), "set_terminal");
    std::vector<double> xs((N + 1) * (size_t)nx_), us(N * (size_t)nu_);
    for (size_t t = 0; t <= N; t++)
    {
      if ((int)xs_init[t].size() != nx_)
        throw std::runtime_error("SolverProxDDP::run: xs_init[t] has the wrong size");
      std::copy(xs_init[t].begin(), xs_init[t].end(), xs.begin() + (long)(t * (size_t)nx_));
    }
    for (size_t t = 0; t < N; t++)
    {
      if ((int)us_init[t].size() != nu_)
        throw std::runtime_error("SolverProxDDP::run: us_init[t] has the wrong size");
      std::copy(us_init[t].begin(), us_init[t].end(), us.begin() + (long)(t * (size_t)nu_));
    }
    check(api_.set_warmstart(ctx_, 0, 1, xs.data(), us.data()), "set_warmstart");
    if ((int)problem.x0_.size() != nx_)
      throw std::runtime_error("SolverProxDDP::run: initial state has the wrong size");
    check(api_.solve(ctx_, problem.x0_.data(), (int)max_iters), "solve");
    std::vector<double> Ks(N * (size_t)(nu_ * ndx_)), ks(N * (size_t)nu_);
    check(api_.get_results(ctx_, 0, 1, xs.data(), us.data(), Ks.data(), ks.data()), "get_results");
    smpc_solve_info info;
    check(api_.get_info(ctx_, 0, 1, &info), "get_info");
    results_.xs.assign(N + 1, VectorXd());
    results_.us.assign(N, VectorXd());
    results_.Ks.assign(N, MatrixXd((size_t)nu_, (size_t)ndx_));
    results_.ks.assign(N, VectorXd());
    for (size_t t = 0; t <= N; t++)
      results_.xs[t].assign(xs.begin() + (long)(t * (size_t)nx_), xs.begin() + (long)((t + 1) * (size_t)nx_));
    for (size_t t = 0; t < N; t++)
    {
      results_.us[t].assign(us.begin() + (long)(t * (size_t)nu_), us.begin() + (long)((t + 1) * (size_t)nu_));
      results_.ks[t].assign(ks.begin() + (long)(t * (size_t)nu_), ks.begin() + (long)((t + 1) * (size_t)nu_));
      // the device stores du = k + K dx; aligator's getCtrlFeedbacks() has the same sign (u = us + K (x (-) xs)),
      // and the reference's control law is u = us[0] - Ks[0] * difference(x_meas, xs[0]) with difference(a, b) = b (-) a
      std::copy(Ks.begin() + (long)(t * (size_t)(nu_ * ndx_)), Ks.begin() + (long)((t + 1) * (size_t)(nu_ * ndx_)), results_.Ks[t].data());
    }
    results_.num_iters = (size_t)info.num_iters;
    results_.al_iter = (size_t)info.al_iters;
    results_.conv = info.conv != 0;
    results_.prim_infeas = info.prim_infeas;
    results_.dual_infeas = info.dual_infeas;
    results_.traj_cost_ = info.traj_cost;
    results_.merit_value_ = info.merit;
    return results_.conv;
  }

  void SolverProxDDP::cycleProblem(const TrajOptProblem &, const StageModel & tail)
  {
    if (!ctx_)
      throw std::runtime_error("SolverProxDDP::cycleProblem called before setup");
    check(api_.cycle(ctx_, 0, 1, &tail.slot), "cycle");
  }

  VectorXd SolverProxDDP::getStateDerivative(std::size_t t)
  {
    VectorXd xdot((size_t)ndx_, 0.0);
    check(api_.get_xdot(ctx_, 0, 1, (int)t, xdot.data()), "get_xdot");
    return xdot;
  }

  MatrixXd SolverProxDDP::getContactForces(std::size_t t, std::size_t nfeet)
  {
    MatrixXd f(nfeet, 3);
    check(api_.get_contact_forces(ctx_, 0, 1, (int)t, f.data()), "get_contact_forces");
    return f;
  }

  // ------------------------------------------------------------------------------------------------
  // MPC
  // ------------------------------------------------------------------------------------------------
  template <class T>
  static void rotate_vec_left(std::vector<T> & v)
  {
    std::rotate(v.begin(), v.begin() + 1, v.end());
  }

  MPC::MPC(const MPCSettings & settings, std::shared_ptr<OCPHandler> problem)
  : MPC(settings, problem, SolverApi::cuda())
  {
  }

  MPC::MPC(const MPCSettings & settings, std::shared_ptr<OCPHandler> problem, const SolverApi & api)
  : settings_(settings)
  , ocp_handler_(problem)
  {
    const RobotModelHandler & model_handler = ocp_handler_->getModelHandler();
    data_handler_ = std::make_shared<RobotDataHandler>(model_handler);
    data_handler_->updateInternalData(model_handler.getReferenceState(), true);
    std::map<std::string, Vector3d> starting_poses;
    for (size_t foot_nb = 0; foot_nb < model_handler.getFeetNb(); foot_nb++)
    {
      const std::string & name = model_handler.getFootFrameName(foot_nb);
      starting_poses.insert({name, data_handler_->getFootPose(foot_nb).translation()});
      relative_feet_poses_.insert({name, data_handler_->getBaseFramePose().inverse() * data_handler_->getFootPose(foot_nb)});
    }
    foot_trajectories_ = FootTrajectory(starting_poses, settings_.swing_apex, settings_.T_fly, settings_.T_contact, ocp_handler_->getSize());
    foot_trajectories_.updateApex(settings.swing_apex);
    x0_ = ocp_handler_->getProblemState(*data_handler_);
    x_reference_ = ocp_handler_->getReferenceState(0);

    // LINEAR rollout and force_initial_condition are what the CUDA solver implements (reference src/mpc.cpp:43-53);
    // num_threads > 1 selects the PARALLEL (leg-condensed) Riccati sweep with that many legs, as in the reference (:46-50)
    solver_ = std::make_unique<SolverProxDDP>(settings_.TOL, settings_.mu_init, maxiters, api);
    solver_->setNumThreads(settings_.num_threads);

    ee_names_ = model_handler.getFeetFrameNames();
    VectorXd force_ref(ocp_handler_->getReferenceForce(0, model_handler.getFootFrameName(0)));
    std::map<std::string, bool> contact_states, land_constraint;
    std::map<std::string, SE3> contact_poses;
    std::map<std::string, VectorXd> force_map;
    for (auto const & name : ee_names_)
    {
      contact_states.insert({name, true});
      land_constraint.insert({name, false});
      contact_poses.insert({name, data_handler_->getFootPose(model_handler.getFootNb(name))});
      force_map.insert({name, force_ref});
    }
    for (std::size_t i = 0; i < ocp_handler_->getProblem().numSteps(); i++)
    {
      xs_.push_back(x0_);
      us_.push_back(ocp_handler_->getReferenceControl(0));
      standing_horizon_.push_back(std::make_shared<StageModel>(ocp_handler_->createStage(contact_states, contact_poses, force_map, land_constraint)));
    }
    xs_.push_back(x0_);

    // aligator's TrajOptProblem keeps the x0 it was constructed with until setInitState is called
    solver_->setup(*ocp_handler_);
    solver_->run(ocp_handler_->getProblem(), xs_, us_);
    xs_ = solver_->results_.xs;
    us_ = solver_->results_.us;
    Ks_ = solver_->results_.getCtrlFeedbacks();
    solver_->max_iters = settings_.max_iters;

    com0_ = data_handler_->getData().com[0];
    now_ = WALKING;
    velocity_base_ = VectorXd::Zero(6);
    next_pose_ = {{0, 0, 0}};
    twist_vect_ = {{0, 0}};
  }

  void MPC::generateCycleHorizon(const std::vector<std::map<std::string, bool>> & contact_states)
  {
    contact_states_ = contact_states;
    // the cycle must be at least as long as the horizon
    const int m = int(ocp_handler_->getProblem().numSteps()) / int(contact_states.size());
    for (int i = 0; i < m; i++)
      contact_states_.insert(contact_states_.end(), contact_states.begin(), contact_states.end());

    const int T = (int)ocp_handler_->getSize();
    for (auto const & name : ee_names_)
    {
      foot_takeoff_times_.insert({name, std::vector<int>()});
      foot_land_times_.insert({name, std::vector<int>()});
      for (size_t i = 1; i < contact_states_.size(); i++)
      {
        if (!contact_states_[i].at(name) && contact_states_[i - 1].at(name))
          foot_takeoff_times_.at(name).push_back((int)i + T);
        if (contact_states_[i].at(name) && !contact_states_[i - 1].at(name))
          foot_land_times_.at(name).push_back((int)i + T);
      }
      // wrap-around switch between the last and the first slot of the cycle
      if (contact_states_.back().at(name) && !contact_states_[0].at(name))
        foot_takeoff_times_.at(name).push_back((int)contact_states_.size() - 1 + T);
      if (!contact_states_.back().at(name) && contact_states_[0].at(name))
        foot_land_times_.at(name).push_back((int)contact_states_.size() - 1 + T);
    }
    std::map<std::string, bool> previous_contacts;
    for (auto const & name : ee_names_)
      previous_contacts.insert({name, true});

    const RobotModelHandler & mh = ocp_handler_->getModelHandler();
    for (auto const & state : contact_states_)
    {
      int active_contacts = 0;
      for (auto const & contact : state)
        active_contacts += contact.second ? 1 : 0;
      VectorXd force_ref(ocp_handler_->getReferenceForce(0, mh.getFootFrameName(0)));
      force_ref.setZero();
      VectorXd force_zero(force_ref);
      force_ref[2] = settings_.support_force / active_contacts;
      std::map<std::string, SE3> contact_poses;
      std::map<std::string, VectorXd> force_map;
      std::map<std::string, bool> land_contacts;
      for (auto const & name : ee_names_)
      {
        contact_poses.insert({name, data_handler_->getFootPose(mh.getFootNb(name))});
        force_map.insert({name, state.at(name) ? force_ref : force_zero});
        land_contacts.insert({name, !previous_contacts.at(name) && state.at(name)});
      }
      cycle_horizon_.push_back(std::make_shared<StageModel>(ocp_handler_->createStage(state, contact_poses, force_map, land_contacts)));
      previous_contacts = state;
    }
  }

  void MPC::iterate(const ConstVectorRef & x)
  {
    data_handler_->updateInternalData(x, false);
    recedeWithCycle();
    updateStepTrackerReferences();

    // warm start: previous solution shifted by one stage (reference src/mpc.cpp:201-207)
    x0_ = ocp_handler_->getProblemState(*data_handler_);
    xs_.erase(xs_.begin());
    xs_[0] = x0_;
    xs_.push_back(xs_.back());
    us_.erase(us_.begin());
    us_.push_back(us_.back());

    ocp_handler_->getProblem().setInitState(x0_);
    solver_->run(ocp_handler_->getProblem(), xs_, us_);

    xs_ = solver_->results_.xs;
    us_ = solver_->results_.us;
    Ks_ = solver_->results_.getCtrlFeedbacks();
  }

  void MPC::recedeWithCycle()
  {
    const int T = (int)ocp_handler_->getSize();
    if (now_ == WALKING || ocp_handler_->getContactSupport(ocp_handler_->getSize() - 1) < ee_names_.size())
    {
      if (cycle_horizon_.empty())
        throw std::runtime_error("generateCycleHorizon must be called before iterate");
      ocp_handler_->getProblem().replaceStageCircular(*cycle_horizon_[0]);
      solver_->cycleProblem(ocp_handler_->getProblem(), *cycle_horizon_[0]);
      rotate_vec_left(cycle_horizon_);
      rotate_vec_left(contact_states_);
      const size_t n = contact_states_.size();
      for (auto const & name : ee_names_)
      {
        if (!contact_states_[n - 1].at(name) && contact_states_[n - 2].at(name))
          foot_takeoff_times_.at(name).push_back((int)n + T);
        if (contact_states_[n - 1].at(name) && !contact_states_[n - 2].at(name))
          foot_land_times_.at(name).push_back((int)n + T);
      }
      updateCycleTiming(false);
    }
    else
    {
      ocp_handler_->getProblem().replaceStageCircular(*standing_horizon_[0]);
      solver_->cycleProblem(ocp_handler_->getProblem(), *standing_horizon_[0]);
      rotate_vec_left(standing_horizon_);
      updateCycleTiming(true);
    }
  }

  void MPC::updateCycleTiming(const bool updateOnlyHorizon)
  {
    const int T = (int)ocp_handler_->getSize();
    for (auto const & name : ee_names_)
    {
      for (auto * times : {&foot_land_times_.at(name), &foot_takeoff_times_.at(name)})
      {
        for (auto & t : *times)
          if (!updateOnlyHorizon || t < T)
            t -= 1;
        if (!times->empty() && (*times)[0] < 0)
          times->erase(times->begin());
      }
    }
  }

  void MPC::updateStepTrackerReferences()
  {
    const size_t T = ocp_handler_->getSize();
    for (auto const & name : ee_names_)
    {
      const size_t foot_nb = ocp_handler_->getModelHandler().getFootNb(name);
      int foot_land_time = -1;
      if (!foot_land_times_.at(name).empty())
        foot_land_time = foot_land_times_.at(name)[0];
      // the swing curve is re-planned only while the landing is at least one flight time away
      const bool update = !(foot_land_time < settings_.T_fly);

      // Raibert heuristic for the next foothold
      const Vector3d & ref = data_handler_->getFootRefPose(foot_nb).translation();
      const Vector3d & base = data_handler_->getBaseFramePose().translation();
      twist_vect_[0] = -(ref[1] - base[1]);
      twist_vect_[1] = ref[0] - base[0];
      const double horizon_time = (settings_.T_fly + settings_.T_contact) * settings_.timestep;
      for (size_t k = 0; k < 2; k++)
        next_pose_[k] = ref[k] + (velocity_base_[k] + velocity_base_[5] * twist_vect_[k]) * horizon_time;
      next_pose_[2] = data_handler_->getFootPose(foot_nb).translation()[2];

      foot_trajectories_.updateTrajectory(update, foot_land_time, data_handler_->getFootPose(foot_nb).translation(), next_pose_, name);
      SE3 pose = SE3::Identity();
      for (size_t time = 0; time < T; time++)
      {
        pose.translation() = foot_trajectories_.getReference(name)[time];
        setReferencePose(time, name, pose);
      }
    }
    ocp_handler_->setReferenceState(T - 1, x_reference_);
    ocp_handler_->setVelocityBase(T - 1, velocity_base_);

    Vector3d com_ref{{0, 0, 0}};
    for (auto const & name : ee_names_)
      com_ref = com_ref + foot_trajectories_.getReference(name).back();
    com_ref = com_ref * (1.0 / (double)ee_names_.size());
    com_ref[2] += com0_[2];
    ocp_handler_->updateTerminalConstraint(com_ref);
  }

  void MPC::setReferencePose(const std::size_t t, const std::string & ee_name, const SE3 & pose_ref)
  {
    ocp_handler_->setReferencePose(t, ee_name, pose_ref);
  }
  void MPC::setTerminalReferencePose(const std::string & ee_name, const SE3 & pose_ref)
  {
    ocp_handler_->setTerminalReferencePose(ee_name, pose_ref);
  }
  const SE3 MPC::getReferencePose(const std::size_t t, const std::string & ee_name) const
  {
    return ocp_handler_->getReferencePose(t, ee_name);
  }
  TrajOptProblem & MPC::getTrajOptProblem() { return ocp_handler_->getProblem(); }

  const VectorXd MPC::getStateDerivative(const std::size_t t)
  {
    if (t >= ocp_handler_->getSize())
      throw std::runtime_error("Stage index exceeds stage vector size");
    return solver_->getStateDerivative(t);
  }

  const MatrixXd MPC::getContactForces(const std::size_t t)
  {
    if (t >= ocp_handler_->getSize())
      throw std::runtime_error("Stage index exceeds stage vector size");
    return solver_->getContactForces(t, ee_names_.size());
  }

  void MPC::switchToWalk(const Vector6d & velocity_base)
  {
    now_ = WALKING;
    velocity_base_ = velocity_base;
  }

  void MPC::switchToStand()
  {
    now_ = STANDING;
    velocity_base_ = VectorXd::Zero(6);
  }

} // namespace simple_mpc
