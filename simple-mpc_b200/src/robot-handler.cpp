// robot-handler.cpp — see include/simple-mpc/robot-handler.hpp
#include "simple-mpc/robot-handler.hpp"
#include <algorithm>

namespace simple_mpc
{

  RobotModelHandler::RobotModelHandler(const Model & model, const std::string & reference_configuration_name, const std::string & base_frame_name)
  : model_(model)
  {
    base_frame_id_ = model_.getFrameId(base_frame_name);
    auto it = model_.referenceConfigurations.find(reference_configuration_name);
    if (it == model_.referenceConfigurations.end())
      throw std::runtime_error("unknown reference configuration '" + reference_configuration_name + "'");
    reference_state_ = it->second;
    reference_state_.resize((size_t)(model_.nq + model_.nv), 0.0);
    mass_ = computeTotalMass(model_);
  }

  void RobotModelHandler::addFootFrames(const std::string & foot_name, const std::string & reference_parent_frame_name)
  {
    const size_t new_foot_index = getFeetNb();
    const FrameIndex foot_frame_id = model_.getFrameId(foot_name);
    feet_frame_names_.push_back(foot_name);
    feet_frame_ids_.push_back(foot_frame_id);
    // "<foot>_ref": a frame rigidly attached to the reference parent frame, placed where the foot is at the
    // reference configuration; the Raibert heuristic reads it (reference src/mpc.cpp:292-299)
    const FrameIndex parent_frame = model_.getFrameId(reference_parent_frame_name);
    const int parent_joint = model_.frames[parent_frame].parentJoint;
    feet_ref_frame_ids_.push_back(model_.addFrame(Frame(foot_name + "_ref", parent_joint, SE3::Identity())));
    ref_parent_frames_.push_back(parent_frame);
    Data data(model_);
    forwardKinematics(model_, data, reference_state_.head((size_t)model_.nq));
    updateFramePlacements(model_, data);
    setFootReferencePlacement(new_foot_index, data.oMf[parent_frame].actInv(data.oMf[foot_frame_id]));
  }

  size_t RobotModelHandler::addPointFoot(const std::string & foot_name, const std::string & reference_parent_frame_name)
  {
    addFootFrames(foot_name, reference_parent_frame_name);
    feet_types_.push_back(POINT);
    return feet_types_.size() - 1;
  }

  size_t RobotModelHandler::addQuadFoot(
    const std::string & foot_name, const std::string & reference_parent_frame_name, const ContactPointsMatrix & contactPoints)
  {
    addFootFrames(foot_name, reference_parent_frame_name);
    feet_types_.push_back(QUAD);
    const size_t foot_nb = feet_types_.size() - 1;
    feet_contact_points_.insert({foot_nb, contactPoints});
    return foot_nb;
  }

  void RobotModelHandler::setFootReferencePlacement(size_t foot_nb, const SE3 & parentframeMfootref)
  {
    // the reference stores the placement relative to the parent FRAME; our frames hang off joints
    const Frame & parent = model_.frames[ref_parent_frames_.at(foot_nb)];
    model_.frames[feet_ref_frame_ids_.at(foot_nb)].placement = parent.placement * parentframeMfootref;
  }

  size_t RobotModelHandler::getFootNb(const std::string & foot_name) const
  {
    auto it = std::find(feet_frame_names_.begin(), feet_frame_names_.end(), foot_name);
    if (it == feet_frame_names_.end())
      throw std::runtime_error("unknown foot '" + foot_name + "'");
    return (size_t)(it - feet_frame_names_.begin());
  }

  VectorXd RobotModelHandler::difference(const ConstVectorRef & x1, const ConstVectorRef & x2) const
  {
    const size_t nq = (size_t)model_.nq, nv = (size_t)model_.nv;
    if (x1.size() != nq + nv || x2.size() != nq + nv)
      throw std::runtime_error("difference: state of the wrong size");
    VectorXd dx = simple_mpc::difference(model_, x1.head(nq), x2.head(nq));
    dx.resize(2 * nv);
    for (size_t k = 0; k < nv; k++)
      dx[nv + k] = x2[nq + k] - x1[nq + k];
    return dx;
  }

  smpc_model_desc RobotModelHandler::toDesc() const
  {
    smpc_model_desc d;
    memset(&d, 0, sizeof(d));
    const int nj = model_.njoints();
    if (nj > SMPC_MAX_JOINTS || model_.nv > SMPC_MAX_NV || model_.nq > SMPC_MAX_NQ)
      throw std::runtime_error("model exceeds SMPC_MAX_JOINTS / NV / NQ");
    if (getFeetNb() > SMPC_MAX_FEET)
      throw std::runtime_error("more feet than SMPC_MAX_FEET");
    d.njoints = nj;
    d.nq = model_.nq;
    d.nv = model_.nv;
    for (int i = 0; i < nj; i++)
    {
      const Joint & j = model_.joints[(size_t)i];
      d.parent[i] = j.parent;
      d.jtype[i] = j.type;
      d.idx_q[i] = j.idx_q;
      d.idx_v[i] = j.idx_v;
      for (int k = 0; k < 3; k++)
      {
        d.axis[i][k] = j.axis[(size_t)k];
        d.place_p[i][k] = j.placement.p[(size_t)k];
        d.lever[i][k] = j.lever[(size_t)k];
      }
      for (int k = 0; k < 9; k++)
      {
        d.place_R[i][k] = j.placement.R[(size_t)k];
        d.inertia[i][k] = j.inertia[(size_t)k];
      }
      d.mass[i] = j.mass;
    }
    d.nfeet = (int)getFeetNb();
    for (size_t f = 0; f < getFeetNb(); f++)
    {
      const Frame & ff = model_.frames[feet_frame_ids_[f]];
      const Frame & fr = model_.frames[feet_ref_frame_ids_[f]];
      d.foot_joint[f] = ff.parentJoint;
      d.footref_joint[f] = fr.parentJoint;
      for (int k = 0; k < 9; k++)
      {
        d.foot_R[f][k] = ff.placement.R[(size_t)k];
        d.footref_R[f][k] = fr.placement.R[(size_t)k];
      }
      for (int k = 0; k < 3; k++)
      {
        d.foot_p[f][k] = ff.placement.p[(size_t)k];
        d.footref_p[f][k] = fr.placement.p[(size_t)k];
      }
    }
    const Frame & bf = model_.frames[base_frame_id_];
    d.base_joint = bf.parentJoint;
    for (int k = 0; k < 9; k++)
      d.base_R[k] = bf.placement.R[(size_t)k];
    for (int k = 0; k < 3; k++)
      d.base_p[k] = bf.placement.p[(size_t)k];
    for (int k = 0; k < model_.nq; k++)
      d.q_ref[k] = reference_state_[(size_t)k];
    return d;
  }

  RobotDataHandler::RobotDataHandler(const RobotModelHandler & model_handler)
  : model_handler_(model_handler)
  , data_(model_handler.getModel())
  , x_(model_handler.getReferenceState().size(), 0.0)
  {
    updateInternalData(model_handler.getReferenceState(), true);
  }

  void RobotDataHandler::updateInternalData(const ConstVectorRef & x, const bool updateJacobians)
  {
    const Model & m = model_handler_.getModel();
    if (x.size() != (size_t)(m.nq + m.nv))
      throw std::runtime_error("updateInternalData: state of the wrong size");
    updateInternalData(x.head((size_t)m.nq), x.tail((size_t)m.nv), updateJacobians);
  }

  void RobotDataHandler::updateInternalData(const ConstVectorRef & q, const ConstVectorRef & v, const bool updateJacobians)
  {
    const Model & m = model_handler_.getModel();
    x_.setSegment(0, q);
    x_.setSegment((size_t)m.nq, v);
    forwardKinematics(m, data_, q);
    updateFramePlacements(m, data_);
    computeCentroidalMomentum(m, data_, q, v);
    if (updateJacobians)
      updateJacobiansMassMatrix(x_);
  }

  void RobotDataHandler::updateJacobiansMassMatrix(const ConstVectorRef &) {}

  RobotDataHandler::CentroidalStateVector RobotDataHandler::getCentroidalState() const
  {
    VectorXd xc(9, 0.0);
    for (size_t k = 0; k < 3; k++)
    {
      xc[k] = data_.com[0][k];
      xc[3 + k] = data_.hg.lin[k];
      xc[6 + k] = data_.hg.ang[k];
    }
    return xc;
  }

} // namespace simple_mpc
