///////////////////////////////////////////////////////////////////////////////
// robot-handler.hpp — RobotModelHandler / RobotDataHandler with the reference's public interface
// (reference include/simple-mpc/robot-handler.hpp:28-271, src/robot-handler.cpp:12-149).
// `Model`, `Data`, `SE3` are this repo's stand-ins for the pinocchio types (model.hpp).
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/model.hpp"

namespace simple_mpc
{

  class RobotModelHandler
  {
  public:
    enum FootType
    {
      POINT,
      QUAD
    };
    /// 4 contact points x 3 coordinates of a rectangular foot (reference ContactPointsMatrix)
    using ContactPointsMatrix = std::array<double, 12>;

    /// reference src/robot-handler.cpp:12-25
    RobotModelHandler(const Model & model, const std::string & reference_configuration_name, const std::string & base_frame_name);

    /// reference src/robot-handler.cpp:56-73 ; return the foot number
    size_t addPointFoot(const std::string & foot_name, const std::string & reference_parent_frame_name);
    size_t addQuadFoot(const std::string & foot_name, const std::string & reference_parent_frame_name, const ContactPointsMatrix & contactPoints);
    /// reference src/robot-handler.cpp:75-78
    void setFootReferencePlacement(size_t foot_nb, const SE3 & parentframeMfootref);
    /// dx = x2 (-) x1, reference src/robot-handler.cpp:81-96
    VectorXd difference(const ConstVectorRef & x1, const ConstVectorRef & x2) const;

    const VectorXd & getReferenceState() const { return reference_state_; }
    size_t getFootNb(const std::string & foot_name) const;
    size_t getFeetNb() const { return feet_frame_ids_.size(); }
    const std::vector<FrameIndex> & getFeetFrameIds() const { return feet_frame_ids_; }
    const std::string & getFootFrameName(size_t foot_nb) const { return feet_frame_names_.at(foot_nb); }
    const std::vector<std::string> & getFeetFrameNames() const { return feet_frame_names_; }
    FootType getFootType(size_t foot_nb) const { return feet_types_.at(foot_nb); }
    const ContactPointsMatrix & getQuadFootContactPoints(size_t foot_nb) const { return feet_contact_points_.at(foot_nb); }
    FrameIndex getBaseFrameId() const { return base_frame_id_; }
    std::string getBaseFrameName() const { return model_.frames[base_frame_id_].name; }
    FrameIndex getFootFrameId(size_t foot_nb) const { return feet_frame_ids_.at(foot_nb); }
    FrameIndex getFootRefFrameId(size_t foot_nb) const { return feet_ref_frame_ids_.at(foot_nb); }
    double getMass() const { return mass_; }
    const Model & getModel() const { return model_; }

    /// lowering of (model, feet, base frame) to the flat table the CUDA library consumes (include/smpc.h)
    smpc_model_desc toDesc() const;

  private:
    void addFootFrames(const std::string & foot_name, const std::string & reference_parent_frame_name);

    Model model_;
    double mass_;
    VectorXd reference_state_;
    std::vector<std::string> feet_frame_names_;
    std::vector<FrameIndex> feet_frame_ids_, feet_ref_frame_ids_, ref_parent_frames_;
    std::vector<FootType> feet_types_;
    std::map<size_t, ContactPointsMatrix> feet_contact_points_;
    FrameIndex base_frame_id_;
  };

  class RobotDataHandler
  {
  public:
    using CentroidalStateVector = VectorXd; // 9: com, linear momentum, angular momentum

    explicit RobotDataHandler(const RobotModelHandler & model_handler);

    /// FK + frame placements + centroidal momentum at x = [q; v]  (reference src/robot-handler.cpp:104-127)
    void updateInternalData(const ConstVectorRef & x, const bool updateJacobians);
    void updateInternalData(const ConstVectorRef & q, const ConstVectorRef & v, const bool updateJacobians);
    /// reference :129-140 computes joint Jacobians, CRBA, NLE, dCCRBA on the host for user code; on this path
    /// they are evaluated inside the CUDA kernels, so this is a no-op kept for source compatibility
    void updateJacobiansMassMatrix(const ConstVectorRef & x);

    const SE3 & getFootRefPose(size_t foot_nb) const { return data_.oMf[model_handler_.getFootRefFrameId(foot_nb)]; }
    const SE3 & getFootPose(size_t foot_nb) const { return data_.oMf[model_handler_.getFootFrameId(foot_nb)]; }
    const SE3 & getBaseFramePose() const { return data_.oMf[model_handler_.getBaseFrameId()]; }
    const RobotModelHandler & getModelHandler() const { return model_handler_; }
    const Data & getData() const { return data_; }
    CentroidalStateVector getCentroidalState() const;
    const VectorXd & getState() const { return x_; }

  private:
    RobotModelHandler model_handler_;
    Data data_;
    VectorXd x_;
  };

} // namespace simple_mpc
