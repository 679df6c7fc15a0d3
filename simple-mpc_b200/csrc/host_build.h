// host_build.h — host-side lowering of the C-ABI descriptors (include/smpc.h) into the POD tables the
// kernels read (dev_types.h). Validation errors mirror the std::runtime_error throws of the reference
// handlers (reference src/kinodynamics.cpp:235, src/ocp-handler.cpp:28-32).
#pragma once
#include "dev_types.h"
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

namespace smpc
{

  inline void build_dev_model(const smpc_model_desc & d, DevModel & M)
  {
    if (d.njoints < 1 || d.njoints > SMPC_MAX_JOINTS)
      throw std::runtime_error("model: njoints out of range");
    if (d.nv > SMPC_MAX_NV || d.nq > SMPC_MAX_NQ)
      throw std::runtime_error("model: nq/nv exceed SMPC_MAX_NQ/NV");
    if (d.jtype[0] != SMPC_JOINT_FREEFLYER || d.parent[0] != -1)
      throw std::runtime_error("model: joint 0 must be the free-flyer root");
    if (d.nfeet < 1 || d.nfeet > SMPC_MAX_FEET)
      throw std::runtime_error("model: nfeet out of range");
    memset(&M, 0, sizeof(M));
    M.d = d;
    int depth[SMPC_MAX_JOINTS];
    int maxd = 0;
    for (int i = 0; i < d.njoints; i++)
    {
      if (i > 0 && (d.parent[i] < 0 || d.parent[i] >= i))
        throw std::runtime_error("model: joints must be topologically ordered");
      if (i > 0 && d.jtype[i] != SMPC_JOINT_REVOLUTE)
        throw std::runtime_error("model: only revolute joints below the free-flyer root are supported");
      depth[i] = (i == 0) ? 0 : depth[d.parent[i]] + 1;
      if (depth[i] > maxd)
        maxd = depth[i];
      const int ndof = (d.jtype[i] == SMPC_JOINT_FREEFLYER) ? 6 : 1;
      for (int k = 0; k < ndof; k++)
        M.joint_of_dof[d.idx_v[i] + k] = i;
      unsigned mask = 0;
      for (int j = i; j >= 0; j = d.parent[j])
        mask |= 1u << j;
      M.anc_mask[i] = mask;
      M.total_mass += d.mass[i];
    }
    M.nlevels = maxd + 1;
    int pos = 0;
    for (int l = 0; l <= maxd; l++)
    {
      M.level_start[l] = pos;
      for (int i = 0; i < d.njoints; i++)
        if (depth[i] == l)
          M.level_joints[pos++] = i;
    }
    M.level_start[maxd + 1] = pos;
  }

  inline bool is_diagonal(const double * W, int n)
  {
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
        if (i != j && W[i * n + j] != 0.0)
          return false;
    return true;
  }

  // [UPSTREAM] aligator CentroidalWrenchConeResidual / MultibodyWrenchConeResidual: 17 rows A w <= 0 on
  // w = [f; tau] for a rectangular foot of half-length L and half-width W: unilaterality, friction pyramid,
  // centre-of-pressure box, yaw-torque bounds.
  inline void build_wrench_cone(double mu, double L, double W, double * A)
  {
    const double s = (L + W) * mu;
    const double rows[17][6] = {
      {0, 0, -1, 0, 0, 0},     {-1, 0, -mu, 0, 0, 0},    {1, 0, -mu, 0, 0, 0},     {0, -1, -mu, 0, 0, 0},
      {0, 1, -mu, 0, 0, 0},    {0, 0, -W, -1, 0, 0},     {0, 0, -W, 1, 0, 0},      {0, 0, -L, 0, -1, 0},
      {0, 0, -L, 0, 1, 0},     {-W, -L, -s, mu, mu, -1}, {-W, L, -s, mu, -mu, -1}, {W, -L, -s, -mu, mu, -1},
      {W, L, -s, -mu, -mu, -1}, {W, L, -s, mu, mu, 1},    {W, -L, -s, mu, -mu, 1},  {-W, L, -s, -mu, mu, 1},
      {-W, -L, -s, -mu, -mu, 1}};
    memcpy(A, rows, sizeof(rows));
  }

  inline void build_dev_ocp(const smpc_model_desc & m, const smpc_ocp_desc & o, DevOcp & O)
  {
    memset(&O, 0, sizeof(O));
    O.kind = o.kind;
    O.N = o.horizon;
    O.fs = o.force_size;
    O.nk = m.nfeet;
    if (O.N < 2)
      throw std::runtime_error("ocp: horizon must be >= 2");
    if (O.fs != 3 && O.fs != 6)
      throw std::runtime_error("force size in settings does not match reference force size");
    O.kinematics_limits = o.kinematics_limits;
    O.force_cone = o.force_cone;
    O.land_cstr = o.land_cstr;
    O.terminal_constraint = o.terminal_constraint;
    O.dt = o.timestep;
    for (int k = 0; k < 3; k++)
      O.gravity[k] = o.gravity[k];
    O.mu = o.mu;
    O.Lfoot = o.Lfoot;
    O.Wfoot = o.Wfoot;
    double mass = 0;
    for (int i = 0; i < m.njoints; i++)
      mass += m.mass[i];
    O.mass = mass;
    const int nj = m.nv - 6;
    auto cp = [](double * dst, const double * src, int n, const char * what) {
      if (!src)
        throw std::runtime_error(std::string("ocp: missing weight ") + what);
      memcpy(dst, src, sizeof(double) * n);
    };
    if (o.kind == SMPC_OCP_KINODYNAMICS)
    {
      O.nx = m.nq + m.nv;
      O.ndx = 2 * m.nv;
      O.nu = nj + O.fs * O.nk;
      if (!o.w_x || !o.w_u)
        throw std::runtime_error("ocp: w_x / w_u missing");
      if (!is_diagonal(o.w_x, O.ndx) || !is_diagonal(o.w_u, O.nu))
        throw std::runtime_error("ocp: the CUDA path implements diagonal w_x / w_u only");
      for (int i = 0; i < O.ndx; i++)
        O.w_x[i] = o.w_x[i * O.ndx + i];
      for (int i = 0; i < O.nu; i++)
        O.w_u[i] = o.w_u[i * O.nu + i];
      cp(O.w_cent, o.w_cent, 36, "w_cent");
      cp(O.w_centder, o.w_centder, 36, "w_centder");
      cp(O.w_frame, o.w_frame, O.fs * O.fs, "w_frame");
      if (o.kinematics_limits)
      {
        cp(O.qmin, o.qmin, nj, "qmin");
        cp(O.qmax, o.qmax, nj, "qmax");
      }
      const int vr = (O.fs == 6) ? 6 : 3;
      O.ncmax = (o.kinematics_limits ? nj : 0) + O.nk * ((o.force_cone ? ((O.fs == 6) ? 17 : 2) : 0) + vr + ((O.fs == 3 && o.land_cstr) ? 1 : 0));
      O.nct = o.terminal_constraint ? 3 : 0;
    }
    else if (o.kind == SMPC_OCP_CENTROIDAL)
    {
      O.nx = O.ndx = 9;
      O.nu = O.fs * O.nk;
      if (!o.w_u || !is_diagonal(o.w_u, O.nu))
        throw std::runtime_error("ocp: the CUDA path implements diagonal w_u only");
      for (int i = 0; i < O.nu; i++)
        O.w_u[i] = o.w_u[i * O.nu + i];
      cp(O.w_com, o.w_com, 9, "w_com");
      cp(O.w_lmom, o.w_linear_mom, 9, "w_linear_mom");
      cp(O.w_amom, o.w_angular_mom, 9, "w_angular_mom");
      cp(O.w_lacc, o.w_linear_acc, 9, "w_linear_acc");
      cp(O.w_aacc, o.w_angular_acc, 9, "w_angular_acc");
      O.ncmax = O.nk * ((O.fs == 6) ? 17 : 2);
      O.nct = 0;
      build_wrench_cone(o.mu, o.Lfoot, o.Wfoot, O.wcone);
    }
    else
    {
      // FullDynamicsOCP (reference src/fulldynamics.cpp:30-76): nu = nv - 6, rigid contacts (3-D point or 6-D quad feet)
      O.nx = m.nq + m.nv;
      O.ndx = 2 * m.nv;
      O.nu = nj;
      if (!o.w_x || !o.w_u)
        throw std::runtime_error("ocp: w_x / w_u missing");
      if (!is_diagonal(o.w_x, O.ndx) || !is_diagonal(o.w_u, O.nu))
        throw std::runtime_error("ocp: the CUDA path implements diagonal w_x / w_u only");
      for (int i = 0; i < O.ndx; i++)
        O.w_x[i] = o.w_x[i * O.ndx + i];
      for (int i = 0; i < O.nu; i++)
        O.w_u[i] = o.w_u[i * O.nu + i];
      cp(O.w_cent, o.w_cent, 36, "w_cent");
      cp(O.w_frame, o.w_frame, O.fs * O.fs, "w_frame");
      cp(O.w_forces, o.w_forces, O.fs * O.fs, "w_forces");
      O.torque_limits = o.torque_limits;
      if (o.torque_limits)
      {
        cp(O.umin, o.umin, nj, "umin");
        cp(O.umax, o.umax, nj, "umax");
      }
      if (o.kinematics_limits)
      {
        cp(O.qmin, o.qmin, nj, "qmin");
        cp(O.qmax, o.qmax, nj, "qmax");
      }
      for (int k = 0; k < O.fs; k++)
      {
        O.Kp[k] = o.Kp_correction[k];
        O.Kd[k] = o.Kd_correction[k];
      }
      build_wrench_cone(o.mu, o.Lfoot, o.Wfoot, O.wcone);
      const double pyr[5][3] = {{0, 0, -1}, {-1, 0, -o.mu}, {1, 0, -o.mu}, {0, -1, -o.mu}, {0, 1, -o.mu}};
      memcpy(O.pyramid, pyr, sizeof(pyr));
      for (int i = 1; i < m.njoints; i++)
        if (m.idx_v[m.parent[i]] >= m.idx_v[i])
          throw std::runtime_error("model: velocity indices must increase from parent to child");
      O.ncmax = (o.torque_limits ? nj : 0) + (o.kinematics_limits ? nj : 0)
                + O.nk * ((o.force_cone ? (O.fs == 6 ? 17 : 5) : 0) + (o.land_cstr ? (O.fs == 6 ? 6 : 4) : 0));
      O.nct = o.terminal_constraint ? 3 : 0;
    }
    if (O.ncmax < O.nct)
      O.ncmax = O.nct;
    if (O.ncmax < 1)
      O.ncmax = 1;
  }

} // namespace smpc
