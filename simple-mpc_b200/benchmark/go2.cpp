// benchmark/go2.cpp — the reference's C++ benchmark program (reference benchmark/go2.cpp:20-157: a Go2 FullDynamicsOCP, T = 50,
// MPC with T_fly 30 / T_contact 5, generateCycleHorizon) written against THIS repo's host classes: the same class and method
// names, std::vector-backed VectorXd / MatrixXd where the reference uses Eigen, and the synthetic go2-like robot where the
// reference loads example-robot-data (not available here). The reference's program stops after generateCycleHorizon (it is
// run under a profiler); this one goes on to time MPC::iterate, which is the call the library accelerates.
//
//   make -C simple-mpc_b200 benchmark && ./simple-mpc_b200/benchmark/go2 [ticks]          (needs a CUDA device)
#include "simple-mpc/fulldynamics.hpp"
#include "simple-mpc/mpc.hpp"
#include "simple-mpc/ocp-handler.hpp"
#include "simple-mpc/robot-handler.hpp"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

using simple_mpc::FullDynamicsOCP;
using simple_mpc::FullDynamicsSettings;
using simple_mpc::MatrixXd;
using simple_mpc::MPC;
using simple_mpc::MPCSettings;
using simple_mpc::OCPHandler;
using simple_mpc::RobotDataHandler;
using simple_mpc::RobotModelHandler;
using simple_mpc::SE3;
using simple_mpc::Vector3d;
using simple_mpc::VectorXd;

int main(int argc, char ** argv)
{
  const int ticks = argc > 1 ? std::atoi(argv[1]) : 200;
  // the reference loads go2.urdf / go2.srdf from example-robot-data (benchmark/go2.cpp:23-30)
  simple_mpc::Model model = simple_mpc::Model::BuildGo2Like();
  std::string base_joint_name = "root_joint";
  RobotModelHandler model_handler(model, "standing", base_joint_name);

  /// Add reference foot for walking
  const double feet[4][2] = {{0.17, 0.15}, {0.17, -0.15}, {-0.24, 0.15}, {-0.24, -0.15}};
  const char * names[4] = {"FL_foot", "FR_foot", "RL_foot", "RR_foot"};
  for (int f = 0; f < 4; f++)
    model_handler.addPointFoot(names[f], base_joint_name);
  for (int f = 0; f < 4; f++)
  {
    SE3 ref = SE3::Identity();
    ref.translation() = Vector3d{{feet[f][0], feet[f][1], 0.0}};
    model_handler.setFootReferencePlacement(model_handler.getFootNb(names[f]), ref);
  }
  RobotDataHandler data_handler(model_handler);

  const size_t nv = (size_t)model_handler.getModel().nv, nu = nv - 6, ndx = nv * 2;
  Vector3d gravity{{0., 0., -9.81}};

  // Define settings for OCP (reference benchmark/go2.cpp:60-101)
  FullDynamicsSettings problem_settings;
  std::vector<double> w_x_vec(ndx, 0.0);
  for (size_t i = 6; i < 24; i++)
    w_x_vec[i] = 10.; // legs, base velocity
  for (size_t i = 24; i < 36; i++)
    w_x_vec[i] = .1; // joint velocities
  problem_settings.timestep = 0.01;
  problem_settings.w_x = MatrixXd::Diagonal(w_x_vec);
  problem_settings.w_u = MatrixXd::Identity(nu) * 1e-4;
  problem_settings.w_cent = MatrixXd::Zero(6, 6);
  problem_settings.gravity = gravity;
  problem_settings.force_size = 3;
  problem_settings.Kp_correction = VectorXd::Zero(3);
  problem_settings.Kd_correction = VectorXd::Zero(3);
  problem_settings.w_forces = MatrixXd::Diagonal({0.0001, 0.0001, 0.0001});
  problem_settings.w_frame = MatrixXd::Identity(3) * 1000;
  problem_settings.umax = model_handler.getModel().effortLimit;
  problem_settings.umin = problem_settings.umax;
  for (double & v : problem_settings.umin)
    v = -v;
  problem_settings.qmin = model_handler.getModel().lowerPositionLimit.tail(nu);
  problem_settings.qmax = model_handler.getModel().upperPositionLimit.tail(nu);
  problem_settings.mu = 0.8;
  problem_settings.Lfoot = 0.01;
  problem_settings.Wfoot = 0.01;
  problem_settings.torque_limits = false;
  problem_settings.kinematics_limits = false;
  problem_settings.force_cone = false;
  problem_settings.land_cstr = false;

  std::shared_ptr<OCPHandler> ocpPtr = std::make_shared<FullDynamicsOCP>(problem_settings, model_handler);

  // Create an OCP problem with horizon size T
  size_t T = 50;
  ocpPtr->createProblem(model_handler.getReferenceState(), T, 3, gravity[2], false);

  // Define settings for MPC
  int T_fly = 30;
  int T_contact = 5;
  MPCSettings mpc_settings;
  mpc_settings.support_force = -gravity[2] * model_handler.getMass();
  mpc_settings.TOL = 1e-4;
  mpc_settings.mu_init = 1e-8;
  mpc_settings.max_iters = 1;
  mpc_settings.num_threads = 1;
  mpc_settings.swing_apex = 0.2;
  mpc_settings.T_fly = T_fly;
  mpc_settings.T_contact = T_contact;
  mpc_settings.timestep = 0.01;

  auto t0 = std::chrono::steady_clock::now();
  std::shared_ptr<MPC> mpc = std::make_shared<MPC>(mpc_settings, ocpPtr);
  const double ctor_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

  // Define the walking gait
  std::vector<std::map<std::string, bool>> contact_states;
  std::map<std::string, bool> contact_state_quadru, contact_phase_lift_FL_RR, contact_phase_lift_RL_FR;
  for (int f = 0; f < 4; f++)
  {
    contact_state_quadru.insert({names[f], true});
    contact_phase_lift_FL_RR.insert({names[f], f == 1 || f == 2});
    contact_phase_lift_RL_FR.insert({names[f], f == 0 || f == 3});
  }
  for (int i = 0; i < T_contact; i++)
    contact_states.push_back(contact_state_quadru);
  for (int i = 0; i < T_fly; i++)
    contact_states.push_back(contact_phase_lift_FL_RR);
  for (int i = 0; i < T_contact; i++)
    contact_states.push_back(contact_state_quadru);
  for (int i = 0; i < T_fly; i++)
    contact_states.push_back(contact_phase_lift_RL_FR);

  mpc->generateCycleHorizon(contact_states);
  // ---- the reference's program ends here; below: the closed loop of the examples without a simulator ----
  VectorXd v(6, 0.0);
  v[0] = 0.2;
  mpc->switchToWalk(v);
  std::vector<double> lat;
  VectorXd x = model_handler.getReferenceState();
  for (int k = 0; k < ticks; k++)
  {
    auto t1 = std::chrono::steady_clock::now();
    mpc->iterate(x);
    lat.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    x = mpc->xs_[1]; // the plant follows the plan
  }
  std::sort(lat.begin(), lat.end());
  std::printf(
    "{\"program\": \"benchmark/go2.cpp\", \"ocp\": \"FullDynamicsOCP go2-like (synthetic), T = 50\", \"constructor_ms\": %.3f, "
    "\"ticks\": %d, \"iterate_ms_p50\": %.4f, \"iterate_ms_p99\": %.4f, \"us0_norm\": %.6f}\n",
    ctor_ms, ticks, lat[lat.size() / 2], lat[(size_t)(0.99 * (double)(lat.size() - 1))],
    [&] { double s = 0; for (double u : mpc->us_[0]) s += u * u; return std::sqrt(s); }());
  return 0;
}
