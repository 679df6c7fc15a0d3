///////////////////////////////////////////////////////////////////////////////
// centroidal-dynamics.hpp — CentroidalOCP with the reference's interface
// (reference include/simple-mpc/centroidal-dynamics.hpp:33-112, src/centroidal-dynamics.cpp).
// State x = [com; linear momentum; angular momentum] in R^9, control u = [f_0 .. f_{nk-1}].
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/ocp-handler.hpp"

namespace simple_mpc
{

  struct CentroidalSettings
  {
    double timestep;
    MatrixXd w_u, w_com, w_linear_mom, w_angular_mom, w_linear_acc, w_angular_acc;
    Vector3d gravity;
    double mu, Lfoot, Wfoot;
    int force_size;
  };

  class CentroidalOCP : public OCPHandler
  {
    using Base = OCPHandler;

  public:
    explicit CentroidalOCP(const CentroidalSettings & settings, const RobotModelHandler & model_handler);
    virtual ~CentroidalOCP() {}

    StageModel createStage(
      const std::map<std::string, bool> & contact_phase, const std::map<std::string, SE3> & contact_pose,
      const std::map<std::string, VectorXd> & contact_force, const std::map<std::string, bool> & land_constraint) override;
    CostStack createTerminalCost() override;
    void createTerminalConstraint(const Vector3d & com_ref) override;
    void updateTerminalConstraint(const Vector3d & com_ref) override;

    void setReferencePose(const std::size_t t, const std::string & ee_name, const SE3 & pose_ref) override;
    void setReferencePoses(const std::size_t i, const std::map<std::string, SE3> & pose_refs) override;
    void setTerminalReferencePose(const std::string &, const SE3 &) override {}
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

    void computeControlFromForces(const std::map<std::string, VectorXd> & force_refs);
    CentroidalSettings getSettings() { return settings_; }
    smpc_ocp_desc toOcpDesc() const override;
    int getNx() const override { return 9; }
    int getNdx() const override { return 9; }

  protected:
    CentroidalSettings settings_;
    Vector3d com_ref_{{0, 0, 0}};
  };

} // namespace simple_mpc
