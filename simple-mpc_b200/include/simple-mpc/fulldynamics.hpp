///////////////////////////////////////////////////////////////////////////////
// fulldynamics.hpp — FullDynamicsOCP with the reference's interface
// (reference include/simple-mpc/fulldynamics.hpp:28-123, src/fulldynamics.cpp).
// State x = [q; v], control u = joint torques (nv - 6).
//
// The handler (problem description, references, getters/setters) is complete on the host side; the CUDA
// library has no constrained-dynamics kernels yet, so SolverProxDDP::setup — hence the MPC constructor —
// throws for this handler (SMPC_ERR_UNSUPPORTED). There is deliberately no CPU path behind it.
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/ocp-handler.hpp"

namespace simple_mpc
{

  struct FullDynamicsSettings
  {
  public:
    double timestep;
    MatrixXd w_x, w_u, w_cent, w_forces, w_frame;
    Vector3d gravity;
    int force_size;
    double mu, Lfoot, Wfoot;
    bool torque_limits, kinematics_limits, force_cone, land_cstr;
    VectorXd umin, umax, qmin, qmax;
    VectorXd Kp_correction, Kd_correction;
  };

  class FullDynamicsOCP : public OCPHandler
  {
    using Base = OCPHandler;

  public:
    explicit FullDynamicsOCP(const FullDynamicsSettings & settings, const RobotModelHandler & model_handler);
    virtual ~FullDynamicsOCP() {}

    StageModel createStage(
      const std::map<std::string, bool> & contact_phase, const std::map<std::string, SE3> & contact_pose,
      const std::map<std::string, VectorXd> & contact_force, const std::map<std::string, bool> & land_constraint) override;
    CostStack createTerminalCost() override;
    void createTerminalConstraint(const Vector3d & com_ref) override;
    void updateTerminalConstraint(const Vector3d & com_ref) override;

    void setReferencePose(const std::size_t t, const std::string & ee_name, const SE3 & pose_ref) override;
    void setReferencePoses(const std::size_t i, const std::map<std::string, SE3> & pose_refs) override;
    void setTerminalReferencePose(const std::string & ee_name, const SE3 & pose_ref) override;
    void setReferenceForces(const std::size_t i, const std::map<std::string, VectorXd> & force_refs) override;
    void setReferenceForce(const std::size_t i, const std::string & ee_name, const ConstVectorRef & force_ref) override;
    const VectorXd getReferenceForce(const std::size_t i, const std::string & cost_name) override;
    const SE3 getReferencePose(const std::size_t i, const std::string & cost_name) override;
    const VectorXd getVelocityBase(const std::size_t t) override;
    const VectorXd getPoseBase(const std::size_t t) override;
    void setPoseBase(const std::size_t t, const ConstVectorRef & pose_base) override;
    void setVelocityBase(const std::size_t t, const ConstVectorRef & velocity_base) override;
    const VectorXd getProblemState(const RobotDataHandler & data_handler) override;
    void setReferenceState(const std::size_t t, const ConstVectorRef & x_ref) override;
    const VectorXd getReferenceState(const std::size_t t) override;

    FullDynamicsSettings getSettings() { return settings_; }
    smpc_ocp_desc toOcpDesc() const override;

  protected:
    FullDynamicsSettings settings_;
  };

} // namespace simple_mpc
