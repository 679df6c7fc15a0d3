// ocp-handler.cpp — see include/simple-mpc/ocp-handler.hpp
#include "simple-mpc/ocp-handler.hpp"
#include <algorithm>

namespace simple_mpc
{

  OCPHandler::~OCPHandler() {}

  OCPHandler::OCPHandler(const RobotModelHandler & model_handler)
  : model_handler_(model_handler)
  , problem_(nullptr)
  {
    nq_ = model_handler.getModel().nq;
    nv_ = model_handler.getModel().nv;
    ndx_ = 2 * nv_;
    nu_ = nv_ - 6;
  }

  std::vector<StageModel> OCPHandler::createStages(
    const std::vector<std::map<std::string, bool>> & contact_phases, const std::vector<std::map<std::string, SE3>> & contact_poses,
    const std::vector<std::map<std::string, VectorXd>> & contact_forces)
  {
    if (contact_phases.size() != contact_poses.size())
      throw std::runtime_error("Contact phases and poses sequences do not have the same size");
    if (contact_phases.size() != contact_forces.size())
      throw std::runtime_error("Contact phases and forces sequences do not have the same size");
    std::map<std::string, bool> previous_phases;
    for (auto const & name : model_handler_.getFeetFrameNames())
      previous_phases.insert({name, true});
    std::vector<StageModel> stage_models;
    for (std::size_t i = 0; i < contact_phases.size(); i++)
    {
      // a foot lands on the stage where it regains contact: that stage carries the landing constraints
      std::map<std::string, bool> land_constraint;
      for (auto const & name : model_handler_.getFeetFrameNames())
        land_constraint.insert({name, !previous_phases.at(name) && contact_phases[i].at(name)});
      stage_models.push_back(createStage(contact_phases[i], contact_poses[i], contact_forces[i], land_constraint));
      previous_phases = contact_phases[i];
    }
    return stage_models;
  }

  StageModel & OCPHandler::stageAt(std::size_t t)
  {
    if (!problem_ || t >= problem_->stages_.size())
      throw std::runtime_error("Stage index exceeds stage vector size");
    return problem_->stages_[t];
  }

  size_t OCPHandler::footIndex(const std::string & ee_name) const
  {
    return model_handler_.getFootNb(ee_name);
  }

  void OCPHandler::setReferenceControl(const std::size_t t, const ConstVectorRef & u_ref)
  {
    StageModel & s = stageAt(t);
    if ((int)u_ref.size() != nu_)
      throw std::runtime_error("u_ref not of the right size");
    for (int k = 0; k < nu_; k++)
      s.slot.u_ref[k] = u_ref[(size_t)k];
  }

  VectorXd OCPHandler::getReferenceControl(const std::size_t t)
  {
    const StageModel & s = stageAt(t);
    return VectorXd(s.slot.u_ref, s.slot.u_ref + nu_);
  }

  CostStack * OCPHandler::getCostStack(std::size_t t) { return &stageAt(t).cost_; }
  CostStack * OCPHandler::getTerminalCostStack() { return &getProblem().term_cost_; }
  std::size_t OCPHandler::getCostNumber() const { return getProblem().stages_.at(0).cost_.components_.size(); }

  size_t OCPHandler::getContactSupport(const std::size_t t)
  {
    size_t n = 0;
    for (bool c : stageAt(t).contact_states)
      n += c ? 1 : 0;
    return n;
  }

  std::vector<bool> OCPHandler::getContactState(const std::size_t t) { return stageAt(t).contact_states; }

  void OCPHandler::createProblem(const ConstVectorRef & x0, const size_t horizon, const int force_size, const double gravity, const bool terminal_constraint)
  {
    VectorXd force_ref((size_t)force_size, 0.0);
    force_ref[2] = -model_handler_.getMass() * gravity / (double)model_handler_.getFeetNb();
    std::map<std::string, bool> contact_phase;
    std::map<std::string, SE3> contact_pose;
    std::map<std::string, VectorXd> contact_force;
    for (auto const & name : model_handler_.getFeetFrameNames())
    {
      contact_phase.insert({name, true});
      contact_pose.insert({name, SE3::Identity()});
      contact_force.insert({name, force_ref});
    }
    std::vector<std::map<std::string, bool>> contact_phases(horizon, contact_phase);
    std::vector<std::map<std::string, SE3>> contact_poses(horizon, contact_pose);
    std::vector<std::map<std::string, VectorXd>> contact_forces(horizon, contact_force);
    std::vector<StageModel> stage_models = createStages(contact_phases, contact_poses, contact_forces);

    problem_ = std::make_unique<TrajOptProblem>();
    problem_->x0_ = x0;
    problem_->stages_ = std::move(stage_models);
    memset(&problem_->term_, 0, sizeof(problem_->term_));
    problem_->term_cost_ = createTerminalCost();
    problem_initialized_ = true;
    if (terminal_constraint)
      createTerminalConstraint(Vector3d{{x0[0], x0[1], x0[2]}});
  }

} // namespace simple_mpc
