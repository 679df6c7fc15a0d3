// module.cpp — Python module `simple_mpc`: the surface of the reference bindings for the MPC::iterate path
// (reference bindings/module.cpp:23-40, expose-robot-handler.cpp, expose-problem.cpp, expose-kinodynamics.cpp,
// expose-fulldynamics.cpp, expose-centroidal.cpp, expose-mpc.cpp), written with pybind11 (boost.python and
// eigenpy are not available here). Handlers and MPC take plain dicts with the reference's keys.
#include "py-convert.hpp"

namespace py = pybind11;
using namespace simple_mpc;
using namespace simple_mpc::python;

namespace
{
  py::array_t<double> arr3(const Vector3d & v)
  {
    py::array_t<double> a(3);
    std::copy(v.begin(), v.end(), a.mutable_data());
    return a;
  }
  py::array_t<double> arr33(const std::array<double, 9> & R)
  {
    py::array_t<double> a({3, 3});
    std::copy(R.begin(), R.end(), a.mutable_data());
    return a;
  }

  // Trampoline for Python-derived handlers (reference include/simple-mpc/python/py-ocp-handler.hpp:71-192: PyOCPHandler,
  // bp::wrapper<OCPHandler> with SIMPLE_MPC_PYTHON_OVERRIDE[_PURE] on every virtual). Here one template serves the
  // abstract OCPHandler (every method pure, as in the reference) and the three concrete handlers (a Python subclass
  // overrides what it needs and inherits the rest), so MPC drives Python overrides through the same virtual calls.
#define SMPC_PY_OVERRIDE(ret, name, ...)                                                                               \
  {                                                                                                                    \
    py::gil_scoped_acquire gil;                                                                                        \
    py::function ovr = py::get_override(static_cast<const Base *>(this), #name);                                       \
    if (ovr)                                                                                                           \
      return ovr(__VA_ARGS__).template cast<ret>();                                                                    \
  }                                                                                                                    \
  if constexpr (std::is_abstract_v<Base>)                                                                              \
    throw std::runtime_error("Tried to call pure virtual function \"OCPHandler::" #name "\"");                          \
  else                                                                                                                 \
    return Base::name(__VA_ARGS__);
#define SMPC_PY_OVERRIDE_VOID(name, ...)                                                                               \
  {                                                                                                                    \
    py::gil_scoped_acquire gil;                                                                                        \
    py::function ovr = py::get_override(static_cast<const Base *>(this), #name);                                       \
    if (ovr)                                                                                                           \
    {                                                                                                                  \
      ovr(__VA_ARGS__);                                                                                                \
      return;                                                                                                          \
    }                                                                                                                  \
  }                                                                                                                    \
  if constexpr (std::is_abstract_v<Base>)                                                                              \
    throw std::runtime_error("Tried to call pure virtual function \"OCPHandler::" #name "\"");                          \
  else                                                                                                                 \
    Base::name(__VA_ARGS__);

  template <class Base>
  struct PyHandler : Base
  {
    using Base::Base;
    typedef std::map<std::string, bool> BoolMap;
    typedef std::map<std::string, SE3> PoseMap;
    typedef std::map<std::string, VectorXd> ForceMap;
    StageModel createStage(const BoolMap & contact_phase, const PoseMap & contact_pose, const ForceMap & contact_force, const BoolMap & land_constraint) override
    {
      SMPC_PY_OVERRIDE(StageModel, createStage, contact_phase, contact_pose, contact_force, land_constraint)
    }
    CostStack createTerminalCost() override { SMPC_PY_OVERRIDE(CostStack, createTerminalCost, ) }
    void createTerminalConstraint(const Vector3d & com_ref) override { SMPC_PY_OVERRIDE_VOID(createTerminalConstraint, com_ref) }
    void updateTerminalConstraint(const Vector3d & com_ref) override { SMPC_PY_OVERRIDE_VOID(updateTerminalConstraint, com_ref) }
    void setReferencePose(const std::size_t t, const std::string & ee_name, const SE3 & pose_ref) override
    {
      SMPC_PY_OVERRIDE_VOID(setReferencePose, t, ee_name, pose_ref)
    }
    void setReferencePoses(const std::size_t t, const PoseMap & pose_refs) override { SMPC_PY_OVERRIDE_VOID(setReferencePoses, t, pose_refs) }
    void setTerminalReferencePose(const std::string & ee_name, const SE3 & pose_ref) override
    {
      SMPC_PY_OVERRIDE_VOID(setTerminalReferencePose, ee_name, pose_ref)
    }
    const SE3 getReferencePose(const std::size_t t, const std::string & ee_name) override { SMPC_PY_OVERRIDE(SE3, getReferencePose, t, ee_name) }
    void setReferenceForces(const std::size_t t, const ForceMap & force_refs) override { SMPC_PY_OVERRIDE_VOID(setReferenceForces, t, force_refs) }
    void setReferenceForce(const std::size_t t, const std::string & ee_name, const ConstVectorRef & force_ref) override
    {
      SMPC_PY_OVERRIDE_VOID(setReferenceForce, t, ee_name, force_ref)
    }
    const VectorXd getReferenceForce(const std::size_t t, const std::string & ee_name) override { SMPC_PY_OVERRIDE(VectorXd, getReferenceForce, t, ee_name) }
    void setVelocityBase(const std::size_t t, const ConstVectorRef & velocity_base) override { SMPC_PY_OVERRIDE_VOID(setVelocityBase, t, velocity_base) }
    const VectorXd getVelocityBase(const std::size_t t) override { SMPC_PY_OVERRIDE(VectorXd, getVelocityBase, t) }
    void setPoseBase(const std::size_t t, const ConstVectorRef & pose_base) override { SMPC_PY_OVERRIDE_VOID(setPoseBase, t, pose_base) }
    const VectorXd getPoseBase(const std::size_t t) override { SMPC_PY_OVERRIDE(VectorXd, getPoseBase, t) }
    const VectorXd getProblemState(const RobotDataHandler & data_handler) override
    {
      {
        py::gil_scoped_acquire gil;
        py::function ovr = py::get_override(static_cast<const Base *>(this), "getProblemState");
        if (ovr)
          return ovr(py::cast(&data_handler, py::return_value_policy::reference)).template cast<VectorXd>();
      }
      if constexpr (std::is_abstract_v<Base>)
        throw std::runtime_error("Tried to call pure virtual function \"OCPHandler::getProblemState\"");
      else
        return Base::getProblemState(data_handler);
    }
    size_t getContactSupport(const std::size_t t) override
    {
      {
        py::gil_scoped_acquire gil;
        py::function ovr = py::get_override(static_cast<const Base *>(this), "getContactSupport");
        if (ovr)
          return ovr(t).template cast<size_t>();
      }
      return Base::getContactSupport(t);
    }
    std::vector<bool> getContactState(const std::size_t t) override
    {
      {
        py::gil_scoped_acquire gil;
        py::function ovr = py::get_override(static_cast<const Base *>(this), "getContactState");
        if (ovr)
          return ovr(t).template cast<std::vector<bool>>();
      }
      return Base::getContactState(t);
    }
    void setReferenceState(const std::size_t t, const ConstVectorRef & x_ref) override { SMPC_PY_OVERRIDE_VOID(setReferenceState, t, x_ref) }
    const VectorXd getReferenceState(const std::size_t t) override { SMPC_PY_OVERRIDE(VectorXd, getReferenceState, t) }
    // the lowering to the CUDA library's problem recipe stays C++: a handler derived in Python from OCPHandler itself has no
    // kernels behind it (the reference's Python handlers build aligator stage models; here a stage is an smpc_slot record)
    smpc_ocp_desc toOcpDesc() const override
    {
      if constexpr (std::is_abstract_v<Base>)
        throw std::runtime_error("OCPHandler derived in Python: derive from KinodynamicsOCP, FullDynamicsOCP or CentroidalOCP to run on the CUDA path");
      else
        return Base::toOcpDesc();
    }
  };
} // namespace

PYBIND11_MODULE(simple_mpc, m)
{
  m.doc() = "simple-mpc on B200: the MPC::iterate -> ProxDDP path behind the reference's Python API";
  m.attr("__version__") = "0.1.0+b200";

  py::class_<SE3>(m, "SE3")
    .def(py::init<>())
    .def(py::init([](const MatrixXd & R, const VectorXd & p) {
      if (R.rows() != 3 || R.cols() != 3 || p.size() != 3)
        throw std::runtime_error("SE3(R, p): R must be 3x3 and p of size 3");
      SE3 M;
      std::copy(R.data(), R.data() + 9, M.R.begin());
      std::copy(p.begin(), p.end(), M.p.begin());
      return M;
    }))
    .def_static("Identity", &SE3::Identity)
    .def_property(
      "translation", [](const SE3 & M) { return arr3(M.p); }, [](SE3 & M, const VectorXd & p) { M.p = {{p.at(0), p.at(1), p.at(2)}}; })
    .def_property(
      "rotation", [](const SE3 & M) { return arr33(M.R); },
      [](SE3 & M, const MatrixXd & R) {
        if (R.rows() != 3 || R.cols() != 3)
          throw std::runtime_error("rotation must be 3x3");
        std::copy(R.data(), R.data() + 9, M.R.begin());
      })
    .def("inverse", &SE3::inverse)
    .def("actInv", &SE3::actInv)
    .def("__mul__", [](const SE3 & a, const SE3 & b) { return a * b; });

  py::class_<Frame>(m, "Frame")
    .def(py::init<const std::string &, int, const SE3 &>(), py::arg("name"), py::arg("parentJoint"), py::arg("placement"))
    .def_readwrite("name", &Frame::name)
    .def_readwrite("parentJoint", &Frame::parentJoint)
    .def_readwrite("placement", &Frame::placement);

  py::class_<Model>(m, "Model", "kinematic tree + frames (stand-in for pinocchio.Model)")
    .def(py::init<>())
    .def_readonly("nq", &Model::nq)
    .def_readonly("nv", &Model::nv)
    .def_property_readonly("njoints", &Model::njoints)
    .def_readwrite("referenceConfigurations", &Model::referenceConfigurations)
    .def_readwrite("lowerPositionLimit", &Model::lowerPositionLimit)
    .def_readwrite("upperPositionLimit", &Model::upperPositionLimit)
    .def_readwrite("effortLimit", &Model::effortLimit)
    .def_readwrite("velocityLimit", &Model::velocityLimit)
    .def_readwrite("friction", &Model::friction)
    .def_readwrite("damping", &Model::damping)
    .def_readonly("frames", &Model::frames)
    .def(
      "addFreeFlyer",
      [](Model & self, const std::string & name, double mass, const VectorXd & lever, const MatrixXd & inertia) {
        std::array<double, 9> I;
        std::copy(inertia.data(), inertia.data() + 9, I.begin());
        return self.addFreeFlyer(name, mass, {{lever.at(0), lever.at(1), lever.at(2)}}, I);
      })
    .def(
      "addRevoluteJoint",
      [](Model & self, size_t parent, const std::string & name, const VectorXd & axis, const SE3 & placement, double mass, const VectorXd & lever,
         const MatrixXd & inertia) {
        std::array<double, 9> I;
        std::copy(inertia.data(), inertia.data() + 9, I.begin());
        return self.addRevoluteJoint(parent, name, {{axis.at(0), axis.at(1), axis.at(2)}}, placement, mass, {{lever.at(0), lever.at(1), lever.at(2)}}, I);
      })
    .def("addFrame", &Model::addFrame)
    .def("existFrame", &Model::existFrame)
    .def("getFrameId", &Model::getFrameId)
    .def_static("BuildGo2Like", &Model::BuildGo2Like, "synthetic 12-DoF quadruped with Go2-like dimensions")
    .def_static("BuildTalosLike", &Model::BuildTalosLike, "synthetic 22-DoF biped with reduced-Talos dimensions");

  py::class_<Data>(m, "Data")
    .def_property_readonly("com", [](const Data & d) { return std::vector<py::array_t<double>>{arr3(d.com[0])}; })
    .def_property_readonly(
      "hg",
      [](const Data & d) {
        VectorXd h(6);
        for (size_t k = 0; k < 3; k++)
          h[k] = d.hg.lin[k], h[3 + k] = d.hg.ang[k];
        return h;
      })
    .def_readonly("oMf", &Data::oMf)
    .def_readonly("oMi", &Data::oMi);

  py::class_<RobotModelHandler>(m, "RobotModelHandler")
    .def(py::init<const Model &, const std::string &, const std::string &>(), py::arg("model"), py::arg("reference_configuration_name"), py::arg("base_frame_name"))
    .def("addPointFoot", &RobotModelHandler::addPointFoot)
    .def(
      "addQuadFoot",
      [](RobotModelHandler & self, const std::string & foot, const std::string & parent, const MatrixXd & pts) {
        if (pts.rows() != 4 || pts.cols() != 3)
          throw std::runtime_error("contact points must be a 4x3 matrix");
        RobotModelHandler::ContactPointsMatrix c;
        std::copy(pts.data(), pts.data() + 12, c.begin());
        return self.addQuadFoot(foot, parent, c);
      })
    .def("setFootReferencePlacement", &RobotModelHandler::setFootReferencePlacement)
    .def("difference", &RobotModelHandler::difference)
    .def("getBaseFrameName", &RobotModelHandler::getBaseFrameName)
    .def("getBaseFrameId", &RobotModelHandler::getBaseFrameId)
    .def("getReferenceState", &RobotModelHandler::getReferenceState)
    .def("getFeetNb", &RobotModelHandler::getFeetNb)
    .def("getFootNb", &RobotModelHandler::getFootNb)
    .def("getFeetFrameIds", &RobotModelHandler::getFeetFrameIds)
    .def("getFootFrameName", &RobotModelHandler::getFootFrameName)
    .def("getFeetFrameNames", &RobotModelHandler::getFeetFrameNames)
    .def("getFootFrameId", &RobotModelHandler::getFootFrameId)
    .def("getFootRefFrameId", &RobotModelHandler::getFootRefFrameId)
    .def("getMass", &RobotModelHandler::getMass)
    .def("getModel", &RobotModelHandler::getModel, py::return_value_policy::reference_internal);

  py::class_<RobotDataHandler, std::shared_ptr<RobotDataHandler>>(m, "RobotDataHandler")
    .def(py::init<const RobotModelHandler &>(), py::arg("model_handler"))
    .def("updateInternalData", py::overload_cast<const ConstVectorRef &, const bool>(&RobotDataHandler::updateInternalData))
    .def("updateJacobiansMassMatrix", &RobotDataHandler::updateJacobiansMassMatrix)
    .def("getFootRefPose", &RobotDataHandler::getFootRefPose)
    .def("getFootPose", &RobotDataHandler::getFootPose)
    .def("getBaseFramePose", &RobotDataHandler::getBaseFramePose)
    .def("getModelHandler", &RobotDataHandler::getModelHandler, py::return_value_policy::reference_internal)
    .def("getData", &RobotDataHandler::getData, py::return_value_policy::reference_internal)
    .def("getCentroidalState", &RobotDataHandler::getCentroidalState)
    .def("getState", &RobotDataHandler::getState);

  py::class_<CostStack>(m, "CostStack").def_readonly("components", &CostStack::components_);
  py::class_<StageModel, std::shared_ptr<StageModel>>(m, "StageModel")
    .def_readonly("contact_states", &StageModel::contact_states)
    .def_readonly("cost", &StageModel::cost_)
    .def_readonly("constraint_names", &StageModel::constraint_names)
    .def("numConstraints", &StageModel::numConstraints)
    .def_property_readonly("contact_mask", [](const StageModel & s) { return s.slot.contact_mask; })
    .def_property_readonly("land_mask", [](const StageModel & s) { return s.slot.land_mask; });
  py::class_<TrajOptProblem>(m, "TrajOptProblem")
    .def("numSteps", &TrajOptProblem::numSteps)
    .def_readonly("stages", &TrajOptProblem::stages_)
    .def_readonly("term_cost", &TrajOptProblem::term_cost_)
    .def_readonly("term_cstrs", &TrajOptProblem::term_cstrs_)
    .def_property("x0", &TrajOptProblem::getInitState, &TrajOptProblem::setInitState);

  py::class_<OCPHandler, PyHandler<OCPHandler>, std::shared_ptr<OCPHandler>>(m, "OCPHandler")
    .def(py::init<const RobotModelHandler &>(), py::arg("model_handler"))
    .def("createProblem", &OCPHandler::createProblem, py::arg("x0"), py::arg("horizon"), py::arg("force_size"), py::arg("gravity"), py::arg("terminal_constraint"))
    .def("createTerminalCost", &OCPHandler::createTerminalCost)
    .def("createTerminalConstraint", [](OCPHandler & o, const VectorXd & c) { o.createTerminalConstraint(toVec3(py::cast(c))); })
    .def("updateTerminalConstraint", [](OCPHandler & o, const VectorXd & c) { o.updateTerminalConstraint(toVec3(py::cast(c))); })
    .def("setReferencePose", &OCPHandler::setReferencePose)
    .def("setReferencePoses", &OCPHandler::setReferencePoses)
    .def("setTerminalReferencePose", &OCPHandler::setTerminalReferencePose)
    .def("getReferencePose", &OCPHandler::getReferencePose)
    .def("setReferenceForces", &OCPHandler::setReferenceForces)
    .def("setReferenceForce", &OCPHandler::setReferenceForce)
    .def("getReferenceForce", &OCPHandler::getReferenceForce)
    .def("setVelocityBase", &OCPHandler::setVelocityBase)
    .def("getVelocityBase", &OCPHandler::getVelocityBase)
    .def("setPoseBase", &OCPHandler::setPoseBase)
    .def("getPoseBase", &OCPHandler::getPoseBase)
    .def("setReferenceState", &OCPHandler::setReferenceState)
    .def("getReferenceState", &OCPHandler::getReferenceState)
    .def("getProblemState", &OCPHandler::getProblemState)
    .def("getContactSupport", &OCPHandler::getContactSupport)
    .def("getContactState", &OCPHandler::getContactState)
    .def("setReferenceControl", &OCPHandler::setReferenceControl)
    .def("getReferenceControl", &OCPHandler::getReferenceControl)
    .def("getCostNumber", &OCPHandler::getCostNumber)
    .def("getSize", &OCPHandler::getSize)
    .def("getNu", &OCPHandler::getNu)
    .def("getModelHandler", &OCPHandler::getModelHandler, py::return_value_policy::reference_internal)
    .def("getProblem", py::overload_cast<>(&OCPHandler::getProblem), py::return_value_policy::reference_internal)
    .def(
      "createStage",
      [](OCPHandler & o, const std::map<std::string, bool> & phase, const std::map<std::string, SE3> & pose,
         const std::map<std::string, VectorXd> & force, const std::map<std::string, bool> & land) { return o.createStage(phase, pose, force, land); },
      py::arg("contact_map"), py::arg("contact_pose"), py::arg("force_refs"), py::arg("land_constraint"));

  py::class_<KinodynamicsOCP, OCPHandler, PyHandler<KinodynamicsOCP>, std::shared_ptr<KinodynamicsOCP>>(m, "KinodynamicsOCP")
    .def(py::init([](const py::dict & s, const RobotModelHandler & h) { return std::make_shared<KinodynamicsOCP>(kinoSettingsFromDict(s), h); },
                  [](const py::dict & s, const RobotModelHandler & h) { return std::make_shared<PyHandler<KinodynamicsOCP>>(kinoSettingsFromDict(s), h); }),
         py::arg("settings"), py::arg("model_handler"))
    .def("getSettings", [](KinodynamicsOCP & o) { return kinoSettingsToDict(o.getSettings()); });
  py::class_<FullDynamicsOCP, OCPHandler, PyHandler<FullDynamicsOCP>, std::shared_ptr<FullDynamicsOCP>>(m, "FullDynamicsOCP")
    .def(py::init([](const py::dict & s, const RobotModelHandler & h) { return std::make_shared<FullDynamicsOCP>(fullSettingsFromDict(s), h); },
                  [](const py::dict & s, const RobotModelHandler & h) { return std::make_shared<PyHandler<FullDynamicsOCP>>(fullSettingsFromDict(s), h); }),
         py::arg("settings"), py::arg("model_handler"))
    .def("getSettings", [](FullDynamicsOCP & o) { return fullSettingsToDict(o.getSettings()); });
  py::class_<CentroidalOCP, OCPHandler, PyHandler<CentroidalOCP>, std::shared_ptr<CentroidalOCP>>(m, "CentroidalOCP")
    .def(py::init([](const py::dict & s, const RobotModelHandler & h) { return std::make_shared<CentroidalOCP>(centroidalSettingsFromDict(s), h); },
                  [](const py::dict & s, const RobotModelHandler & h) { return std::make_shared<PyHandler<CentroidalOCP>>(centroidalSettingsFromDict(s), h); }),
         py::arg("settings"), py::arg("model_handler"))
    .def("getSettings", [](CentroidalOCP & o) { return centroidalSettingsToDict(o.getSettings()); });

  py::class_<Results>(m, "Results")
    .def_readonly("xs", &Results::xs)
    .def_readonly("us", &Results::us)
    .def_readonly("num_iters", &Results::num_iters)
    .def_readonly("al_iter", &Results::al_iter)
    .def_readonly("conv", &Results::conv)
    .def_readonly("prim_infeas", &Results::prim_infeas)
    .def_readonly("dual_infeas", &Results::dual_infeas)
    .def_readonly("traj_cost", &Results::traj_cost_)
    .def_readonly("merit_value", &Results::merit_value_)
    .def("controlFeedbacks", &Results::getCtrlFeedbacks)
    .def("controlFeedforwards", &Results::getCtrlFeedforwards);
  py::class_<SolverProxDDP>(m, "SolverProxDDP")
    .def_readwrite("max_iters", &SolverProxDDP::max_iters)
    .def_readwrite("target_tol", &SolverProxDDP::target_tol_)
    .def_readwrite("mu_init", &SolverProxDDP::mu_init)
    .def_readonly("results", &SolverProxDDP::results_);

  // FrictionCompensation (reference bindings/expose-friction-compensation.cpp): computeFriction returns the corrected torque
  py::class_<FrictionCompensation>(m, "FrictionCompensation")
    .def(py::init<const Model &, bool>(), py::arg("model"), py::arg("with_free_flyer") = true)
    .def("computeFriction",
         [](FrictionCompensation & self, const VectorXd & velocity, VectorXd torque) {
           self.computeFriction(velocity, torque);
           return torque;
         })
    .def_readonly("nu", &FrictionCompensation::nu_)
    .def_readonly("dry_friction", &FrictionCompensation::dry_friction_)
    .def_readonly("viscuous_friction", &FrictionCompensation::viscuous_friction_);

  // KinodynamicsID / CentroidalID (reference bindings/expose-inverse-dynamics.cpp:62-94): settings classes with read-write
  // fields, setTarget / solve(t, q, v) -> tau / getAccelerations() -> a
  py::class_<KinodynamicsID::Settings>(m, "KinodynamicsIDSettings")
    .def(py::init<>())
    .def_readwrite("friction_coefficient", &KinodynamicsID::Settings::friction_coefficient)
    .def_readwrite("contact_weight_ratio_max", &KinodynamicsID::Settings::contact_weight_ratio_max)
    .def_readwrite("contact_weight_ratio_min", &KinodynamicsID::Settings::contact_weight_ratio_min)
    .def_readwrite("kp_base", &KinodynamicsID::Settings::kp_base)
    .def_readwrite("kp_posture", &KinodynamicsID::Settings::kp_posture)
    .def_readwrite("kp_contact", &KinodynamicsID::Settings::kp_contact)
    .def_readwrite("w_base", &KinodynamicsID::Settings::w_base)
    .def_readwrite("w_posture", &KinodynamicsID::Settings::w_posture)
    .def_readwrite("w_contact_motion", &KinodynamicsID::Settings::w_contact_motion)
    .def_readwrite("w_contact_force", &KinodynamicsID::Settings::w_contact_force)
    .def_readwrite("contact_motion_equality", &KinodynamicsID::Settings::contact_motion_equality);
  py::class_<KinodynamicsID>(m, "KinodynamicsID")
    .def(py::init<const RobotModelHandler &, double, const KinodynamicsID::Settings>(), py::arg("model_handler"), py::arg("control_dt"),
         py::arg("settings"), py::keep_alive<1, 2>())
    .def("setTarget", &KinodynamicsID::setTarget)
    .def("solve",
         [](KinodynamicsID & self, double t, const VectorXd & q, const VectorXd & v) {
           VectorXd tau;
           self.solve(t, q, v, tau);
           return tau;
         })
    .def("getAccelerations",
         [](KinodynamicsID & self) {
           VectorXd a;
           self.getAccelerations(a);
           return a;
         })
    .def("getContactForces", &KinodynamicsID::getContactForces);
  py::class_<CentroidalID::Settings, KinodynamicsID::Settings>(m, "CentroidalIDSettings")
    .def(py::init<>())
    .def_readwrite("kp_com", &CentroidalID::Settings::kp_com)
    .def_readwrite("kp_feet_tracking", &CentroidalID::Settings::kp_feet_tracking)
    .def_readwrite("w_com", &CentroidalID::Settings::w_com)
    .def_readwrite("w_feet_tracking", &CentroidalID::Settings::w_feet_tracking);
  py::class_<CentroidalID>(m, "CentroidalID")
    .def(py::init<const RobotModelHandler &, double, const CentroidalID::Settings>(), py::arg("model_handler"), py::arg("control_dt"),
         py::arg("settings"), py::keep_alive<1, 2>())
    .def("setTarget",
         [](CentroidalID & self, const VectorXd & com, const VectorXd & vcom, const std::vector<SE3> & feet_pose,
            const std::vector<VectorXd> & feet_velocity, const std::vector<bool> & contact, const std::vector<VectorXd> & f_target) {
           if (com.size() != 3 || vcom.size() != 3)
             throw std::runtime_error("CentroidalID.setTarget: com_position / com_velocity must have 3 entries");
           CentroidalID::FeetVelocityVector fv(feet_velocity.size());
           for (size_t i = 0; i < feet_velocity.size(); i++)
           {
             if (feet_velocity[i].size() != 6)
               throw std::runtime_error("CentroidalID.setTarget: feet_velocity entries are [linear; angular] (6)");
             for (size_t e = 0; e < 3; e++)
               fv[i].lin[e] = feet_velocity[i][e], fv[i].ang[e] = feet_velocity[i][3 + e];
           }
           self.setTarget({{com[0], com[1], com[2]}}, {{vcom[0], vcom[1], vcom[2]}}, feet_pose, fv, contact, f_target);
         })
    .def("solve",
         [](CentroidalID & self, double t, const VectorXd & q, const VectorXd & v) {
           VectorXd tau;
           self.solve(t, q, v, tau);
           return tau;
         })
    .def("getAccelerations",
         [](CentroidalID & self) {
           VectorXd a;
           self.getAccelerations(a);
           return a;
         })
    .def("getContactForces", &CentroidalID::getContactForces);

  // Interpolator (reference bindings/expose-interpolate.cpp:20-62): each method (delay, timestep, list) -> value
  py::class_<Interpolator>(m, "Interpolator")
    .def(py::init<const Model &>(), py::arg("model"))
    .def("interpolateConfiguration",
         [](Interpolator & self, double delay, double timestep, const std::vector<VectorXd> & qs) {
           VectorXd q;
           self.interpolateConfiguration(delay, timestep, qs, q);
           return q;
         })
    .def("interpolateState",
         [](Interpolator & self, double delay, double timestep, const std::vector<VectorXd> & xs) {
           VectorXd x;
           self.interpolateState(delay, timestep, xs, x);
           return x;
         })
    .def("interpolateLinear",
         [](Interpolator & self, double delay, double timestep, const std::vector<VectorXd> & vs) {
           VectorXd v;
           self.interpolateLinear(delay, timestep, vs, v);
           return v;
         })
    .def("interpolateContacts", [](Interpolator & self, double delay, double timestep, const std::vector<std::vector<bool>> & cs) {
      std::vector<bool> c;
      self.interpolateContacts(delay, timestep, cs, c);
      return c;
    });

  py::class_<MPC>(m, "MPC")
    .def(py::init([](const py::dict & s, std::shared_ptr<OCPHandler> ocp) { return std::make_unique<MPC>(mpcSettingsFromDict(s), ocp); }),
         py::arg("settings"), py::arg("problem"))
    .def("getSettings", [](MPC & self) { return mpcSettingsToDict(self.settings_); })
    .def("generateCycleHorizon", &MPC::generateCycleHorizon, py::arg("contact_states"))
    .def("iterate", &MPC::iterate, py::arg("x"))
    .def("setReferencePose", &MPC::setReferencePose)
    .def("getReferencePose", &MPC::getReferencePose)
    .def("setTerminalReferencePose", &MPC::setTerminalReferencePose)
    .def_readwrite("velocity_base", &MPC::velocity_base_)
    .def_readwrite("x_reference", &MPC::x_reference_)
    .def_readonly("ocp_handler", &MPC::ocp_handler_)
    .def("switchToWalk", &MPC::switchToWalk)
    .def("switchToStand", &MPC::switchToStand)
    .def("getFootTakeoffCycle", &MPC::getFootTakeoffCycle)
    .def("getFootLandCycle", &MPC::getFootLandCycle)
    .def("getStateDerivative", &MPC::getStateDerivative)
    .def("getContactForces", &MPC::getContactForces)
    .def("getCyclingContactState", &MPC::getCyclingContactState)
    .def("getModelHandler", &MPC::getModelHandler, py::return_value_policy::reference_internal)
    .def("getDataHandler", &MPC::getDataHandler, py::return_value_policy::reference_internal)
    .def("getTrajOptProblem", &MPC::getTrajOptProblem, py::return_value_policy::reference_internal)
    .def("getCycleHorizon", &MPC::getCycleHorizon, py::return_value_policy::reference_internal)
    .def_property_readonly("solver", [](MPC & self) { return self.solver_.get(); }, py::return_value_policy::reference_internal)
    .def_readonly("xs", &MPC::xs_)
    .def_readonly("us", &MPC::us_)
    .def_readonly("Ks", &MPC::Ks_);
}
