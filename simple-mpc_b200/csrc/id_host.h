// id_host.h — host side of the inverse-dynamics contexts: what KinodynamicsID::setTarget / CentroidalID::setTarget store
// (reference src/inverse-dynamics/kinodynamics-id.cpp:121-186, centroidal-id.cpp:88-145) lowered to IdTargetDev records.
// The only geometry needed on the host is the placement of the base frame at the target configuration (:137) and, once
// per context, the centre of mass at the reference configuration (centroidal-id.cpp:64): a plain forward-kinematics loop.
#pragma once
#include "id_qp.cuh"
#include <cmath>
#include <cstring>
#include <vector>

namespace smpc
{
  namespace idh
  {
    inline void mul3(const double * A, const double * B, double * o)
    {
      double t[9];
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
          t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
      memcpy(o, t, sizeof(t));
    }
    inline void mv3(const double * A, const double * v, double * o)
    {
      double t[3];
      for (int i = 0; i < 3; i++)
        t[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
      memcpy(o, t, sizeof(t));
    }
    // world placements (R 9, p 3) of every joint
    inline void fk(const smpc_model_desc & d, const double * q, std::vector<double> & oMi)
    {
      oMi.assign(12 * d.njoints, 0.0);
      for (int i = 0; i < d.njoints; i++)
      {
        double Rj[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pj[3] = {0, 0, 0};
        if (d.jtype[i] == SMPC_JOINT_FREEFLYER)
        {
          const double * qq = q + d.idx_q[i];
          double x = qq[3], y = qq[4], z = qq[5], w = qq[6];
          const double n = 1.0 / std::sqrt(x * x + y * y + z * z + w * w);
          x *= n, y *= n, z *= n, w *= n;
          const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                               2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                               2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
          memcpy(Rj, R, sizeof(R));
          pj[0] = qq[0], pj[1] = qq[1], pj[2] = qq[2];
        }
        else
        {
          const double ang = q[d.idx_q[i]], s = std::sin(ang), c = std::cos(ang);
          const double * a = d.axis[i];
          const double W[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
          double W2[9];
          mul3(W, W, W2);
          for (int e = 0; e < 9; e++)
            Rj[e] = ((e % 4 == 0) ? 1.0 : 0.0) + s * W[e] + (1 - c) * W2[e];
        }
        double Rl[9], pl[3], t[3];
        mul3(d.place_R[i], Rj, Rl);
        mv3(d.place_R[i], pj, t);
        for (int e = 0; e < 3; e++)
          pl[e] = d.place_p[i][e] + t[e];
        double * Ri = &oMi[12 * i];
        if (d.parent[i] < 0)
        {
          memcpy(Ri, Rl, sizeof(Rl));
          memcpy(Ri + 9, pl, sizeof(pl));
        }
        else
        {
          const double * Rp = &oMi[12 * d.parent[i]];
          mul3(Rp, Rl, Ri);
          mv3(Rp, pl, t);
          for (int e = 0; e < 3; e++)
            Ri[9 + e] = Rp[9 + e] + t[e];
        }
      }
    }
    inline void frame_pose(const std::vector<double> & oMi, int joint, const double * Rl, const double * pl, double * R, double * p)
    {
      const double * Ri = &oMi[12 * joint];
      double t[3];
      mul3(Ri, Rl, R);
      mv3(Ri, pl, t);
      for (int e = 0; e < 3; e++)
        p[e] = Ri[9 + e] + t[e];
    }
    inline void com(const smpc_model_desc & d, const std::vector<double> & oMi, double * c)
    {
      double tot = 0, acc[3] = {0, 0, 0};
      for (int i = 0; i < d.njoints; i++)
      {
        double t[3];
        mv3(&oMi[12 * i], d.lever[i], t);
        for (int e = 0; e < 3; e++)
          acc[e] += d.mass[i] * (t[e] + oMi[12 * i + 9 + e]);
        tot += d.mass[i];
      }
      for (int e = 0; e < 3; e++)
        c[e] = acc[e] / tot;
    }
  } // namespace idh

  // KinodynamicsID::setTarget (kinodynamics-id.cpp:121-186)
  inline void id_host_set_target(
    const smpc_model_desc & d, IdTargetDev & T, const double * q, const double * v, const double * a, const int32_t * contact,
    const double * f_target)
  {
    const int na = d.nv - 6;
    if (d.base_joint == 0)
    { // the usual case (base frame on the free-flyer): no tree walk, no allocation - this runs per instance at 1 kHz
      const double * qq = q + d.idx_q[0];
      double x = qq[3], y = qq[4], z = qq[5], w = qq[6];
      const double n = 1.0 / std::sqrt(x * x + y * y + z * z + w * w);
      x *= n, y *= n, z *= n, w *= n;
      const double Rj[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
      double R0[9], p0[3], t[3];
      idh::mul3(d.place_R[0], Rj, R0);
      idh::mv3(d.place_R[0], qq, t);
      for (int e = 0; e < 3; e++)
        p0[e] = d.place_p[0][e] + t[e];
      idh::mul3(R0, d.base_R, T.base_R);
      idh::mv3(R0, d.base_p, t);
      for (int e = 0; e < 3; e++)
        T.base_p[e] = p0[e] + t[e];
    }
    else
    {
      std::vector<double> oMi;
      idh::fk(d, q, oMi);
      idh::frame_pose(oMi, d.base_joint, d.base_R, d.base_p, T.base_R, T.base_p);
    }
    for (int j = 0; j < na; j++)
    {
      T.q_post[j] = q[7 + j];
      T.v_post[j] = v[6 + j];
      T.a_post[j] = a[6 + j];
    }
    for (int r = 0; r < 6; r++)
    {
      T.base_vel[r] = v[r];
      T.base_acc[r] = a[r];
    }
    for (int f = 0; f < d.nfeet; f++)
    {
      T.active[f] = contact[f] ? 1 : 0;
      for (int r = 0; r < 6; r++)
        T.f_ref[f][r] = (contact[f] && f_target) ? f_target[6 * f + r] : 0.0;
    }
  }

  // CentroidalID::setTarget (centroidal-id.cpp:88-145)
  inline void id_host_set_target_centroidal(
    const smpc_model_desc & d, const smpc_id_desc & D, IdTargetDev & T, const double * com, const double * vcom, const double * pose,
    const double * vel, const int32_t * contact, const double * f_target)
  {
    for (int r = 0; r < 3; r++)
    {
      T.com_ref[r] = com[r];
      T.com_vel[r] = vcom[r];
    }
    if (D.settings.w_feet_tracking > 0.)
      for (int f = 0; f < d.nfeet; f++)
      {
        T.tracked[f] = contact[f] ? 0 : 1;
        if (!contact[f])
        {
          memcpy(T.foot_R[f], pose + 12 * f, 9 * sizeof(double));
          memcpy(T.foot_p[f], pose + 12 * f + 9, 3 * sizeof(double));
          memcpy(T.foot_vel[f], vel + 6 * f, 6 * sizeof(double));
        }
      }
    std::vector<double> z(d.nv, 0.0);
    id_host_set_target(d, T, d.q_ref, z.data(), z.data(), contact, f_target);
  }

  // the constructors' default target: reference state, every foot in contact with weight / nfeet on its z axis
  // (kinodynamics-id.cpp:100-114, centroidal-id.cpp:57-83)
  inline void id_host_default_target(const smpc_model_desc & d, const smpc_id_desc & D, IdTargetDev & T)
  {
    memset(&T, 0, sizeof(T));
    double mass = 0;
    for (int i = 0; i < d.njoints; i++)
      mass += d.mass[i];
    std::vector<double> f(6 * d.nfeet, 0.0), z(d.nv, 0.0);
    std::vector<int32_t> cs(d.nfeet, 1);
    for (int i = 0; i < d.nfeet; i++)
      f[6 * i + 2] = mass * 9.81 / d.nfeet;
    if (D.settings.centroidal)
    {
      std::vector<double> oMi, pose(12 * d.nfeet), vel(6 * d.nfeet, 0.0);
      idh::fk(d, d.q_ref, oMi);
      double c[3];
      idh::com(d, oMi, c);
      for (int i = 0; i < d.nfeet; i++)
        idh::frame_pose(oMi, d.footref_joint[i], d.footref_R[i], d.footref_p[i], &pose[12 * i], &pose[12 * i + 9]);
      const double vz[3] = {0, 0, 0};
      id_host_set_target_centroidal(d, D, T, c, vz, pose.data(), vel.data(), cs.data(), f.data());
    }
    else
      id_host_set_target(d, T, d.q_ref, z.data(), z.data(), cs.data(), f.data());
  }

} // namespace smpc
