///////////////////////////////////////////////////////////////////////////////
// model.hpp — kinematic tree + frames: the stand-in for pinocchio::Model / pinocchio::Data in the
// reference's signatures (RobotModelHandler copies a pinocchio::Model, reference
// src/robot-handler.cpp:12-14). Only what the MPC::iterate path needs on the HOST is here:
// forward kinematics, frame placements, centre of mass and centroidal momentum
// (reference src/robot-handler.cpp:114-127). Everything heavier runs in the CUDA kernels.
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/fwd.hpp"
#include "smpc.h"

namespace simple_mpc
{

  struct Joint
  {
    std::string name;
    int parent = -1;
    int type = SMPC_JOINT_REVOLUTE;
    Vector3d axis{{0, 0, 1}};
    SE3 placement;            // parent-joint frame <- joint frame
    double mass = 0;          // body attached to the joint
    Vector3d lever{{0, 0, 0}};
    std::array<double, 9> inertia{{0, 0, 0, 0, 0, 0, 0, 0, 0}}; // about the body's centre of mass
    int idx_q = 0, idx_v = 0;
  };

  struct Frame
  {
    std::string name;
    int parentJoint = 0;
    SE3 placement; // joint frame <- this frame
    Frame() = default;
    Frame(const std::string & n, int joint, const SE3 & M) : name(n), parentJoint(joint), placement(M) {}
  };

  using FrameIndex = std::size_t;
  using JointIndex = std::size_t;

  class Model
  {
  public:
    int nq = 0, nv = 0;
    std::vector<Joint> joints; // joint 0 = free-flyer root
    std::vector<Frame> frames;
    std::map<std::string, VectorXd> referenceConfigurations;
    VectorXd lowerPositionLimit, upperPositionLimit, effortLimit;
    VectorXd velocityLimit; // nv - 6 actuated joints, like the limits above; read by KinodynamicsID / CentroidalID only
    VectorXd friction, damping; // nv entries each (pinocchio's convention) or empty = zero; read by FrictionCompensation only

    /// root free-flyer (must be added first); returns joint index 0
    JointIndex addFreeFlyer(const std::string & name, double mass, const Vector3d & lever, const std::array<double, 9> & inertia);
    JointIndex addRevoluteJoint(
      JointIndex parent, const std::string & name, const Vector3d & axis, const SE3 & placement, double mass,
      const Vector3d & lever, const std::array<double, 9> & inertia);
    FrameIndex addFrame(const Frame & f);
    bool existFrame(const std::string & name) const;
    /// throws std::runtime_error for an unknown frame (pinocchio returns nframes; the reference never checks)
    FrameIndex getFrameId(const std::string & name) const;
    int njoints() const { return (int)joints.size(); }
    /// the synthetic robots the tests and benchmarks use (example-robot-data is not available here)
    static Model BuildGo2Like();
    static Model BuildTalosLike();
  };

  double computeTotalMass(const Model & model);

  /// Per-evaluation kinematic quantities (pinocchio::Data stand-in).
  struct Data
  {
    std::vector<SE3> oMi, oMf;
    std::vector<Vector3d> com; // com[0]: whole-body centre of mass
    struct Momentum
    {
      Vector3d lin{{0, 0, 0}}, ang{{0, 0, 0}};
      const Vector3d & linear() const { return lin; }
      const Vector3d & angular() const { return ang; }
    } hg; // centroidal momentum: world-aligned axes, about the centre of mass
    explicit Data(const Model & m) : oMi(m.joints.size()), oMf(m.frames.size()), com(1) {}
    Data() = default;
  };

  void forwardKinematics(const Model & model, Data & data, const VectorXd & q);
  void updateFramePlacements(const Model & model, Data & data);
  void computeCentroidalMomentum(const Model & model, Data & data, const VectorXd & q, const VectorXd & v);
  /// dq = q2 (-) q1 on the configuration manifold (free-flyer: SE(3) log in the body frame)
  VectorXd difference(const Model & model, const VectorXd & q1, const VectorXd & q2);
  /// q (+) dq on the configuration manifold (free-flyer: q * exp6(dq), body-frame tangent)
  VectorXd integrate(const Model & model, const VectorXd & q, const VectorXd & dq);
  /// q0 (+) u (q1 (-) q0)   (pinocchio::interpolate)
  VectorXd interpolate(const Model & model, const VectorXd & q0, const VectorXd & q1, double u);

} // namespace simple_mpc
