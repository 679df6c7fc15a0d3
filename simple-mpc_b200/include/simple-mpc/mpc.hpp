///////////////////////////////////////////////////////////////////////////////
// mpc.hpp — the MPC driver with the reference's interface (reference include/simple-mpc/mpc.hpp:29-197,
// src/mpc.cpp). `solver_` is this repo's SolverProxDDP: the same role and member names as
// aligator::SolverProxDDPTpl<double> in the reference (run / setup / cycleProblem / results_ / max_iters),
// executing on the GPU through the C ABI of include/smpc.h.
///////////////////////////////////////////////////////////////////////////////
#pragma once
#include "simple-mpc/foot-trajectory.hpp"
#include "simple-mpc/ocp-handler.hpp"
#include "simple-mpc/robot-handler.hpp"
#include "simple-mpc/solver-api.hpp"

namespace simple_mpc
{

  struct MPCSettings
  {
  public:
    double swing_apex = 0.15;
    double support_force = 1000;
    double TOL = 1e-4;
    double mu_init = 1e-8;
    std::size_t max_iters = 1;
    std::size_t num_threads = 2; // > 1: PARALLEL Riccati with that many legs (reference src/mpc.cpp:46-50)
    int T_fly = 80;
    int T_contact = 20;
    size_t T = 100;
    double timestep = 0.01;
  };

  /// aligator::ResultsTpl stand-in: what simple-mpc and its users read after run()
  struct Results
  {
    std::vector<VectorXd> xs, us;
    std::vector<MatrixXd> Ks;  // feedback gains, nu x ndx per stage
    std::vector<VectorXd> ks;  // feed-forward terms
    std::size_t num_iters = 0, al_iter = 0;
    bool conv = false;
    double prim_infeas = 0, dual_infeas = 0, traj_cost_ = 0, merit_value_ = 0;
    const std::vector<MatrixXd> & getCtrlFeedbacks() const { return Ks; }
    const std::vector<VectorXd> & getCtrlFeedforwards() const { return ks; }
  };

  /// ProxDDP on the GPU. One object = one smpc_ctx with batch 1 (one reference `MPC` = one problem).
  class SolverProxDDP
  {
  public:
    SolverProxDDP(double tol, double mu_init, std::size_t max_iters, const SolverApi & api = SolverApi::cuda(), int device = 0);
    ~SolverProxDDP();
    SolverProxDDP(const SolverProxDDP &) = delete;
    SolverProxDDP & operator=(const SolverProxDDP &) = delete;

    /// aligator SolverProxDDP::setNumThreads + linear_solver_choice = PARALLEL when n > 1 (reference src/mpc.cpp:46-50)
    void setNumThreads(std::size_t n) { opts_.num_threads = (int)n; }
    /// allocate the device-resident solver state for this problem (aligator SolverProxDDP::setup)
    void setup(const OCPHandler & ocp);
    /// one ProxDDP run from the warm start (xs_init, us_init); returns the convergence flag
    bool run(const TrajOptProblem & problem, const std::vector<VectorXd> & xs_init, const std::vector<VectorXd> & us_init);
    /// rotate the solver's duals with the horizon (aligator SolverProxDDP::cycleProblem)
    void cycleProblem(const TrajOptProblem & problem, const StageModel & tail);
    VectorXd getStateDerivative(std::size_t t);
    MatrixXd getContactForces(std::size_t t, std::size_t nfeet);

    double target_tol_, mu_init;
    std::size_t max_iters;
    Results results_;
    smpc_solver_opts opts_; // every other ProxDDP setting (upstream defaults)

  private:
    void check(int rc, const char * what);
    const SolverApi & api_;
    int device_;
    smpc_ctx * ctx_ = nullptr;
    int nx_ = 0, ndx_ = 0, nu_ = 0, N_ = 0;
  };

  class MPC
  {
  protected:
    enum LocomotionType
    {
      WALKING,
      STANDING,
      MOTION
    };

    std::vector<std::map<std::string, bool>> contact_states_;
    std::vector<std::shared_ptr<StageModel>> cycle_horizon_;
    std::vector<std::shared_ptr<StageModel>> standing_horizon_;
    FootTrajectory foot_trajectories_;
    std::map<std::string, SE3> relative_feet_poses_;
    void updateStepTrackerReferences();

    std::vector<std::string> ee_names_;
    Vector3d com0_;
    LocomotionType now_;
    std::shared_ptr<RobotDataHandler> data_handler_;

    /// test hook: the same driver on another implementation of the C ABI (the CPU oracle in tests/)
    MPC(const MPCSettings & settings, std::shared_ptr<OCPHandler> problem, const SolverApi & api);

  public:
    std::unique_ptr<SolverProxDDP> solver_;
    Vector6d velocity_base_;
    Vector3d next_pose_;
    std::array<double, 2> twist_vect_;
    MPCSettings settings_;
    std::shared_ptr<OCPHandler> ocp_handler_;

    explicit MPC(const MPCSettings & settings, std::shared_ptr<OCPHandler> problem);
    virtual ~MPC() = default;
    MPC(const MPC &) = delete;
    MPC & operator=(const MPC &) = delete;

    void generateCycleHorizon(const std::vector<std::map<std::string, bool>> & contact_states);
    void iterate(const ConstVectorRef & x);
    void updateCycleTiming(const bool updateOnlyHorizon);
    void recedeWithCycle();

    void setReferencePose(const std::size_t t, const std::string & ee_name, const SE3 & pose_ref);
    void setTerminalReferencePose(const std::string & ee_name, const SE3 & pose_ref);
    const SE3 getReferencePose(const std::size_t t, const std::string & ee_name) const;
    void setVelocityBase(const Vector6d & v) { velocity_base_ = v; }
    void setReferenceState(const VectorXd & state_ref) { x_reference_ = state_ref; }

    TrajOptProblem & getTrajOptProblem();
    const RobotDataHandler & getDataHandler() const { return *data_handler_; }
    const RobotModelHandler & getModelHandler() const { return ocp_handler_->getModelHandler(); }
    std::vector<std::shared_ptr<StageModel>> & getCycleHorizon() { return cycle_horizon_; }
    inline bool getCyclingContactState(const std::size_t t, const std::string & ee_name) const { return contact_states_[t].at(ee_name); }
    inline int getFootTakeoffCycle(const std::string & ee_name) const
    {
      return foot_takeoff_times_.at(ee_name).empty() ? -1 : foot_takeoff_times_.at(ee_name)[0];
    }
    inline int getFootLandCycle(const std::string & ee_name) const
    {
      return foot_land_times_.at(ee_name).empty() ? -1 : foot_land_times_.at(ee_name)[0];
    }
    const VectorXd getStateDerivative(const std::size_t t);
    /// @warning Only works with the full-dynamics OCP handler (reference include/simple-mpc/mpc.hpp:174-178)
    const MatrixXd getContactForces(const std::size_t t);
    void switchToWalk(const Vector6d & velocity_base);
    void switchToStand();

    std::map<std::string, std::vector<int>> foot_takeoff_times_, foot_land_times_;
    std::vector<VectorXd> xs_;
    std::vector<VectorXd> us_;
    std::vector<MatrixXd> Ks_;
    VectorXd x_reference_;
    VectorXd x0_;
    VectorXd u0_;
  };

} // namespace simple_mpc
