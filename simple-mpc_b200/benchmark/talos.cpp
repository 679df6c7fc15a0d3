// benchmark/talos.cpp — the reference's second C++ benchmark program (reference benchmark/talos.cpp:20-194: Talos FullDynamicsOCP
// with 6-D contacts, T = 100, terminal constraint, MPC with the default T_fly 80 / T_contact 20, walking gait 10 / 50 / 10 / 50 / 10)
// written against THIS repo's host classes; the synthetic 22-DoF talos-like robot replaces the reduced Talos URDF of
// example-robot-data. As in go2.cpp the program goes on to time MPC::iterate after the point where the reference's stops.
#include "simple-mpc/fulldynamics.hpp"
#include "simple-mpc/mpc.hpp"
#include "simple-mpc/ocp-handler.hpp"
#include "simple-mpc/robot-handler.hpp"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

using namespace simple_mpc;

int main(int argc, char ** argv)
{
  const int ticks = argc > 1 ? std::atoi(argv[1]) : 100;
  // the reference builds the reduced model from talos_reduced.urdf with 22 controlled joints (benchmark/talos.cpp:23-69)
  Model model = Model::BuildTalosLike();
  std::string base_joint = "base_link";
  RobotModelHandler model_handler(model, "half_sitting", base_joint);

  // Add feet
  RobotModelHandler::ContactPointsMatrix foot_quad{{0.1, 0.075, 0, -0.1, 0.075, 0, -0.1, -0.075, 0, 0.1, -0.075, 0}};
  model_handler.addQuadFoot("left_sole_link", base_joint, foot_quad);
  model_handler.addQuadFoot("right_sole_link", base_joint, foot_quad);
  // pinocchio::SE3(Eigen::Quaternion(0., 0., 0., 1.), t): w = 0, z = 1 - a half turn about z, as the reference writes it (:76-81)
  SE3 ref_left, ref_right;
  ref_left.rotation() = {{-1, 0, 0, 0, -1, 0, 0, 0, 1}};
  ref_right.rotation() = ref_left.rotation();
  ref_left.translation() = Vector3d{{0., 0.1, 0.}};
  ref_right.translation() = Vector3d{{0., -0.1, 0.}};
  model_handler.setFootReferencePlacement(model_handler.getFootNb("left_sole_link"), ref_left);
  model_handler.setFootReferencePlacement(model_handler.getFootNb("right_sole_link"), ref_right);

  RobotDataHandler data_handler(model_handler);

  // Define settings for OCP
  size_t T = 100;
  FullDynamicsSettings problem_settings;
  const size_t nu = (size_t)model_handler.getModel().nv - 6;
  std::vector<double> w_x_vec = {0, 0, 0, 100, 100, 100,          // Base pos/ori
                                 0.1, 0.1, 0.1, 0.1, 0.1, 0.1,    // Left leg
                                 0.1, 0.1, 0.1, 0.1, 0.1, 0.1,    // Right leg
                                 10, 10,                          // Torso
                                 1, 1, 1, 1,                      // Left arm
                                 1, 1, 1, 1,                      // Right arm
                                 1, 1, 1, 1, 1, 1,                // Base pos/ori vel
                                 0.1, 0.1, 0.1, 0.1, 0.01, 0.01,  // Left leg vel
                                 0.1, 0.1, 0.1, 0.1, 0.01, 0.01,  // Right leg vel
                                 10, 10,                          // Torso vel
                                 1, 1, 1, 1,                      // Left arm vel
                                 1, 1, 1, 1};                     // Right arm vel
  problem_settings.timestep = 0.01;
  problem_settings.w_x = MatrixXd::Diagonal(w_x_vec);
  problem_settings.w_u = MatrixXd::Identity(nu) * 1e-4;
  problem_settings.w_cent = MatrixXd::Diagonal({0, 0, 0, 0.1, 0., 100});
  problem_settings.gravity = {{0, 0, -9.81}};
  problem_settings.force_size = 6;
  problem_settings.Kp_correction = VectorXd(6, 1.0);
  problem_settings.Kd_correction = VectorXd(6, 1.0);
  problem_settings.w_forces = MatrixXd::Diagonal({0.0001, 0.0001, 0.0001, 0.01, 0.01, 0.01});
  problem_settings.w_frame = MatrixXd::Identity(6) * 2000;
  problem_settings.umax = model_handler.getModel().effortLimit;
  problem_settings.umin = problem_settings.umax;
  for (double & v : problem_settings.umin)
    v = -v;
  problem_settings.qmin = model_handler.getModel().lowerPositionLimit.tail(nu);
  problem_settings.qmax = model_handler.getModel().upperPositionLimit.tail(nu);
  problem_settings.mu = 0.8;
  problem_settings.Lfoot = 0.1;
  problem_settings.Wfoot = 0.075;
  problem_settings.torque_limits = false;
  problem_settings.kinematics_limits = false;
  problem_settings.force_cone = false;
  problem_settings.land_cstr = false;

  std::shared_ptr<OCPHandler> ocpPtr = std::make_shared<FullDynamicsOCP>(problem_settings, model_handler);
  ocpPtr->createProblem(model_handler.getReferenceState(), T, 6, problem_settings.gravity[2], true);

  // Define settings for MPC
  MPCSettings mpc_settings;
  mpc_settings.support_force = -problem_settings.gravity[2] * model_handler.getMass();
  mpc_settings.TOL = 1e-4;
  mpc_settings.mu_init = 1e-8;
  mpc_settings.max_iters = 1;
  mpc_settings.num_threads = 2;

  auto t0 = std::chrono::steady_clock::now();
  MPC mpc{mpc_settings, ocpPtr};
  const double ctor_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

  // Define the walking gait
  std::vector<std::map<std::string, bool>> contact_states;
  const int phases[5][3] = {{10, 1, 1}, {50, 1, 0}, {10, 1, 1}, {50, 0, 1}, {10, 1, 1}};
  for (const auto & ph : phases)
    for (int i = 0; i < ph[0]; i++)
    {
      std::map<std::string, bool> contact_state;
      contact_state.insert({model_handler.getFootFrameName(0), ph[1] != 0});
      contact_state.insert({model_handler.getFootFrameName(1), ph[2] != 0});
      contact_states.push_back(contact_state);
    }
  mpc.generateCycleHorizon(contact_states);
  // ---- the reference's program ends here ----
  VectorXd v(6, 0.0);
  v[0] = 0.1;
  mpc.switchToWalk(v);
  std::vector<double> lat;
  VectorXd x = model_handler.getReferenceState();
  for (int k = 0; k < ticks; k++)
  {
    auto t1 = std::chrono::steady_clock::now();
    mpc.iterate(x);
    lat.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    x = mpc.xs_[1];
  }
  std::sort(lat.begin(), lat.end());
  double s = 0;
  for (double u : mpc.us_[0])
    s += u * u;
  std::printf(
    "{\"program\": \"benchmark/talos.cpp\", \"ocp\": \"FullDynamicsOCP talos-like (synthetic), 6-D contacts, T = 100\", \"constructor_ms\": %.3f, "
    "\"ticks\": %d, \"iterate_ms_p50\": %.4f, \"iterate_ms_p99\": %.4f, \"us0_norm\": %.6f}\n",
    ctor_ms, ticks, lat[lat.size() / 2], lat[(size_t)(0.99 * (double)(lat.size() - 1))], std::sqrt(s));
  return std::isfinite(s) ? 0 : 1;
}
