///////////////////////////////////////////////////////////////////////////////
// ocp-handler.hpp — abstract OCP builder with the reference's interface
// (reference include/simple-mpc/ocp-handler.hpp:42-164, src/ocp-handler.cpp).
//
// In the reference a "problem" is an aligator::TrajOptProblem of polymorphic StageModels. Here a stage is
// the flat record the CUDA kernels consume (smpc_slot: contact phase + every reference simple-mpc ever
// mutates) and the recipe shared by all stages (weights, limits, flags) is an smpc_ocp_desc: createStage
// LOWERS the reference's stage recipe instead of instantiating cost/constraint objects.
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/robot-handler.hpp"
#include "smpc.h"

namespace simple_mpc
{

  /// What the reference's aligator::CostStackTpl exposes that simple-mpc's callers read: the component names.
  struct CostStack
  {
    std::vector<std::string> components_;
    bool hasComponent(const std::string & name) const
    {
      for (auto const & c : components_)
        if (c == name)
          return true;
      return false;
    }
  };

  /// One horizon stage: flat record + the descriptive data the reference keeps on its StageModel.
  struct StageModel
  {
    smpc_slot slot;
    std::vector<bool> contact_states; // feet insertion order
    CostStack cost_;
    std::vector<std::string> constraint_names;
    std::size_t numConstraints() const { return constraint_names.size(); }
    StageModel() { memset(&slot, 0, sizeof(slot)); }
  };

  /// aligator::TrajOptProblemTpl stand-in: initial state, stages, terminal node.
  struct TrajOptProblem
  {
    VectorXd x0_;
    std::vector<StageModel> stages_;
    CostStack term_cost_;
    smpc_terminal term_;
    std::vector<std::string> term_cstrs_;
    std::size_t numSteps() const { return stages_.size(); }
    void setInitState(const VectorXd & x0) { x0_ = x0; }
    const VectorXd & getInitState() const { return x0_; }
    /// drop stage 0, append `tail` (aligator TrajOptProblem::replaceStageCircular)
    void replaceStageCircular(const StageModel & tail)
    {
      stages_.erase(stages_.begin());
      stages_.push_back(tail);
    }
  };

  class OCPHandler
  {
  public:
    explicit OCPHandler(const RobotModelHandler & model_handler);
    virtual ~OCPHandler();

    virtual StageModel createStage(
      const std::map<std::string, bool> & contact_phase, const std::map<std::string, SE3> & contact_pose,
      const std::map<std::string, VectorXd> & force_refs, const std::map<std::string, bool> & land_constraint) = 0;
    virtual std::vector<StageModel> createStages(
      const std::vector<std::map<std::string, bool>> & contact_phases, const std::vector<std::map<std::string, SE3>> & contact_poses,
      const std::vector<std::map<std::string, VectorXd>> & contact_forces);
    virtual CostStack createTerminalCost() = 0;
    virtual void updateTerminalConstraint(const Vector3d & com_ref) = 0;
    virtual void createTerminalConstraint(const Vector3d & com_ref) = 0;

    virtual void setReferencePose(const std::size_t t, const std::string & ee_name, const SE3 & pose_ref) = 0;
    virtual void setReferencePoses(const std::size_t t, const std::map<std::string, SE3> & pose_refs) = 0;
    virtual void setTerminalReferencePose(const std::string & ee_name, const SE3 & pose_ref) = 0;
    virtual const SE3 getReferencePose(const std::size_t t, const std::string & ee_name) = 0;
    virtual const VectorXd getVelocityBase(const std::size_t t) = 0;
    virtual void setVelocityBase(const std::size_t t, const ConstVectorRef & velocity_base) = 0;
    virtual const VectorXd getPoseBase(const std::size_t t) = 0;
    virtual void setPoseBase(const std::size_t t, const ConstVectorRef & pose_base) = 0;
    virtual void setReferenceForces(const std::size_t t, const std::map<std::string, VectorXd> & force_refs) = 0;
    virtual void setReferenceForce(const std::size_t t, const std::string & ee_name, const ConstVectorRef & force_ref) = 0;
    virtual const VectorXd getReferenceForce(const std::size_t t, const std::string & ee_name) = 0;
    virtual const VectorXd getProblemState(const RobotDataHandler & data_handler) = 0;
    virtual size_t getContactSupport(const std::size_t t);
    virtual std::vector<bool> getContactState(const std::size_t t);
    virtual void setReferenceState(const std::size_t t, const ConstVectorRef & x_ref) = 0;
    virtual const VectorXd getReferenceState(const std::size_t t) = 0;

    /// reference src/ocp-handler.cpp:96-137
    void createProblem(const ConstVectorRef & x0, const size_t horizon, const int force_size, const double gravity, const bool terminal_constraint = false);
    void setReferenceControl(const std::size_t t, const ConstVectorRef & u_ref);
    VectorXd getReferenceControl(const std::size_t t);

    CostStack * getCostStack(std::size_t t);
    CostStack * getTerminalCostStack();
    std::size_t getCostNumber() const;
    std::size_t getSize() const { return problem_->numSteps(); }
    TrajOptProblem & getProblem()
    {
      if (!problem_)
        throw std::runtime_error("Create problem first!");
      return *problem_;
    }
    const TrajOptProblem & getProblem() const
    {
      if (!problem_)
        throw std::runtime_error("Create problem first!");
      return *problem_;
    }
    const RobotModelHandler & getModelHandler() const { return model_handler_; }
    int getNu() { return nu_; }

    /// Lowering for the CUDA library: the recipe shared by all stages. The pointers inside refer to storage
    /// owned by this handler and stay valid for its lifetime.
    virtual smpc_ocp_desc toOcpDesc() const = 0;
    /// dimension of the state the problem is posed on (nq + nv, or 9 for the centroidal problem)
    virtual int getNx() const { return nq_ + nv_; }
    virtual int getNdx() const { return ndx_; }

  protected:
    StageModel & stageAt(std::size_t t);
    size_t footIndex(const std::string & ee_name) const;

    int nq_, nv_, ndx_, nu_;
    bool problem_initialized_ = false;
    bool terminal_constraint_ = false;
    VectorXd x0_;
    RobotModelHandler model_handler_;
    std::unique_ptr<TrajOptProblem> problem_;
    VectorXd control_ref_;
  };

} // namespace simple_mpc
