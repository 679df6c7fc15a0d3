// ORACLE — test infrastructure only (see orc_math.hpp header). PARITY UNPINNED.
//
// orc_rbd.hpp — rigid-body algorithms the hot path needs, restated from the published algorithms
// (Featherstone RBDA; Carpentier & Mansard, "Analytical derivatives of rigid body dynamics
// algorithms", RSS 2018) that Pinocchio 3.8 [UPSTREAM, not in /root/reference] implements:
//   forwardKinematics / updateFramePlacements / computeCentroidalMomentum   reference src/robot-handler.cpp:119-121
//   ccrba / dccrba / computeCentroidalDynamicsDerivatives / jacobianCenterOfMass /
//   getFrameJacobian / getFrameVelocityDerivatives                            inside aligator's kinodynamics residuals,
//                                                                             built at reference src/kinodynamics.cpp:56-87,110-111
// Everything is expressed in the WORLD frame (origin at the world origin) with dense 6x6 matrices:
// clarity over speed, this is the checker.
#pragma once
#include "orc_math.hpp"
#include "../include/smpc.h"

namespace orc
{

  struct Model
  {
    smpc_model_desc d;
    int nj, nq, nv, nfeet;
    std::vector<int> joint_of_dof;     // dof column -> joint
    std::vector<std::vector<char>> anc; // anc[i][j] = 1 if joint j is an ancestor-or-self of joint i
    double total_mass;
    explicit Model(const smpc_model_desc & desc) : d(desc)
    {
      nj = d.njoints;
      nq = d.nq;
      nv = d.nv;
      nfeet = d.nfeet;
      joint_of_dof.assign(nv, 0);
      for (int i = 0; i < nj; i++)
      {
        int n = (d.jtype[i] == SMPC_JOINT_FREEFLYER) ? 6 : 1;
        for (int k = 0; k < n; k++)
          joint_of_dof[d.idx_v[i] + k] = i;
      }
      anc.assign(nj, std::vector<char>(nj, 0));
      for (int i = 0; i < nj; i++)
      {
        int j = i;
        while (j >= 0)
        {
          anc[i][j] = 1;
          j = d.parent[j];
        }
      }
      total_mass = 0;
      for (int i = 0; i < nj; i++)
        total_mass += d.mass[i];
    }
  };

  // World-frame spatial inertia (about the world origin) as a dense symmetric 6x6:
  //   [[ m I , -m [c]x ],[ m [c]x , Ic - m [c]x [c]x ]]
  inline void inertia6(double m, const double * c, const double * Ic, double * I6)
  {
    double C[9], CC[9];
    skew3(c, C);
    matmul3(C, C, CC);
    for (int i = 0; i < 36; i++)
      I6[i] = 0;
    for (int i = 0; i < 3; i++)
    {
      I6[6 * i + i] = m;
      for (int j = 0; j < 3; j++)
      {
        I6[6 * i + (3 + j)] = -m * C[3 * i + j];
        I6[6 * (3 + i) + j] = m * C[3 * i + j];
        I6[6 * (3 + i) + (3 + j)] = Ic[3 * i + j] - m * CC[3 * i + j];
      }
    }
  }
  // 6x6 matrix of (v x .) on motions and (v x* .) on forces
  inline void crm6(const double * v, double * X)
  {
    double V[9], W[9];
    skew3(v, V);
    skew3(v + 3, W);
    for (int i = 0; i < 36; i++)
      X[i] = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
      {
        X[6 * i + j] = W[3 * i + j];
        X[6 * i + 3 + j] = V[3 * i + j];
        X[6 * (3 + i) + 3 + j] = W[3 * i + j];
      }
  }
  inline void crf6(const double * v, double * X)
  {
    double V[9], W[9];
    skew3(v, V);
    skew3(v + 3, W);
    for (int i = 0; i < 36; i++)
      X[i] = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
      {
        X[6 * i + j] = W[3 * i + j];
        X[6 * (3 + i) + j] = V[3 * i + j];
        X[6 * (3 + i) + 3 + j] = W[3 * i + j];
      }
  }
  inline void mat6mul(const double * A, const double * B, double * O)
  {
    double t[36];
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++)
      {
        double s = 0;
        for (int k = 0; k < 6; k++)
          s += A[6 * i + k] * B[6 * k + j];
        t[6 * i + j] = s;
      }
    memcpy(O, t, sizeof(t));
  }

  // Everything one (q, v, a) evaluation produces.
  struct Kin
  {
    std::vector<SE3> oMi;
    std::vector<SE3> oMf, oMfref; // feet, foot reference frames
    SE3 oMbase;
    Mat J;                         // 6 x nv, world-frame joint Jacobian columns
    std::vector<double> ov, oa;    // 6 per joint: world-frame spatial velocity / acceleration (no gravity)
    std::vector<double> I6, dI6;   // 36 per joint: world-frame inertia and its time derivative
    std::vector<double> Y, dY;     // 36 per joint: composite (subtree) versions
    std::vector<double> h, f;      // 6 per joint: body momentum / inertial force
    std::vector<double> hs, fs;    // 6 per joint: subtree sums
    std::vector<double> msub;      // subtree mass
    std::vector<double> csub;      // 3 per joint: subtree com
    double com[3], vcom[3];
    double h0[6], hd0[6]; // momentum and its rate about the world origin
    double hg[6], hdg[6]; // centroidal
    Mat Ag;               // 6 x nv centroidal momentum matrix
    Mat Jcom;             // 3 x nv
    Mat dh_dq, dhd_dq, dhd_dv; // 6 x nv centroidal derivatives (dh_dv = dhd_da = Ag)
  };

  inline void joint_transform(const smpc_model_desc & d, int i, const double * q, SE3 & liMi)
  {
    SE3 P;
    memcpy(P.R, d.place_R[i], sizeof(P.R));
    memcpy(P.p, d.place_p[i], sizeof(P.p));
    SE3 Jt;
    if (d.jtype[i] == SMPC_JOINT_FREEFLYER)
    {
      const double * qq = q + d.idx_q[i];
      quat_to_R(qq + 3, Jt.R);
      Jt.p[0] = qq[0];
      Jt.p[1] = qq[1];
      Jt.p[2] = qq[2];
    }
    else
    {
      double w[3];
      double ang = q[d.idx_q[i]];
      for (int k = 0; k < 3; k++)
        w[k] = d.axis[i][k] * ang;
      // exact Rodrigues with sin/cos of the joint angle (axis is unit)
      double s = std::sin(ang), c = std::cos(ang);
      double W[9], W2[9];
      skew3(d.axis[i], W);
      matmul3(W, W, W2);
      for (int k = 0; k < 9; k++)
        Jt.R[k] = ((k % 4 == 0) ? 1.0 : 0.0) + s * W[k] + (1 - c) * W2[k];
      (void)w;
    }
    liMi = se3_mul(P, Jt);
  }

  // Forward kinematics + frame placements (q only).
  inline void fk(const Model & m, const double * q, Kin & k)
  {
    const smpc_model_desc & d = m.d;
    k.oMi.resize(m.nj);
    for (int i = 0; i < m.nj; i++)
    {
      SE3 li;
      joint_transform(d, i, q, li);
      k.oMi[i] = (d.parent[i] < 0) ? li : se3_mul(k.oMi[d.parent[i]], li);
    }
    k.oMf.resize(m.nfeet);
    k.oMfref.resize(m.nfeet);
    for (int f = 0; f < m.nfeet; f++)
    {
      SE3 F;
      memcpy(F.R, d.foot_R[f], sizeof(F.R));
      memcpy(F.p, d.foot_p[f], sizeof(F.p));
      k.oMf[f] = se3_mul(k.oMi[d.foot_joint[f]], F);
      memcpy(F.R, d.footref_R[f], sizeof(F.R));
      memcpy(F.p, d.footref_p[f], sizeof(F.p));
      k.oMfref[f] = se3_mul(k.oMi[d.footref_joint[f]], F);
    }
    SE3 Bf;
    memcpy(Bf.R, d.base_R, sizeof(Bf.R));
    memcpy(Bf.p, d.base_p, sizeof(Bf.p));
    k.oMbase = se3_mul(k.oMi[d.base_joint], Bf);
  }

  // Full pass: kinematics, centroidal quantities and (if derivs) the centroidal derivatives,
  // for joint-space velocity v and acceleration a (a may be NULL = 0).
  inline void centroidal_pass(const Model & m, const double * q, const double * v, const double * a, bool derivs, Kin & k)
  {
    const smpc_model_desc & d = m.d;
    const int nj = m.nj, nv = m.nv;
    fk(m, q, k);
    k.J.resize(6, nv);
    k.ov.assign(6 * nj, 0);
    k.oa.assign(6 * nj, 0);
    k.I6.assign(36 * nj, 0);
    k.dI6.assign(36 * nj, 0);
    k.h.assign(6 * nj, 0);
    k.f.assign(6 * nj, 0);
    for (int i = 0; i < nj; i++)
    {
      const int iv = d.idx_v[i];
      const int ndof = (d.jtype[i] == SMPC_JOINT_FREEFLYER) ? 6 : 1;
      // joint Jacobian columns in the world frame
      for (int c = 0; c < ndof; c++)
      {
        double S[6] = {0, 0, 0, 0, 0, 0};
        if (ndof == 6)
          S[c] = 1;
        else
        {
          S[3] = d.axis[i][0];
          S[4] = d.axis[i][1];
          S[5] = d.axis[i][2];
        }
        double col[6];
        se3_act_motion(k.oMi[i], S, col);
        for (int r = 0; r < 6; r++)
          k.J(r, iv + c) = col[r];
      }
      double vJ[6] = {0, 0, 0, 0, 0, 0}, aJ[6] = {0, 0, 0, 0, 0, 0};
      for (int c = 0; c < ndof; c++)
        for (int r = 0; r < 6; r++)
        {
          vJ[r] += k.J(r, iv + c) * v[iv + c];
          if (a)
            aJ[r] += k.J(r, iv + c) * a[iv + c];
        }
      double * vi = &k.ov[6 * i];
      double * ai = &k.oa[6 * i];
      const int p = d.parent[i];
      for (int r = 0; r < 6; r++)
      {
        vi[r] = (p >= 0 ? k.ov[6 * p + r] : 0.0) + vJ[r];
      }
      double vxvJ[6];
      mxm(vi, vJ, vxvJ);
      for (int r = 0; r < 6; r++)
        ai[r] = (p >= 0 ? k.oa[6 * p + r] : 0.0) + aJ[r] + vxvJ[r];
      // world-frame body inertia
      double cw[3], Rc[3], Icw[9], T[9];
      matvec3(k.oMi[i].R, d.lever[i], Rc);
      for (int r = 0; r < 3; r++)
        cw[r] = Rc[r] + k.oMi[i].p[r];
      matmul3(k.oMi[i].R, d.inertia[i], T);
      // T * R^T
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
          Icw[3 * r + c] = T[3 * r] * k.oMi[i].R[3 * c] + T[3 * r + 1] * k.oMi[i].R[3 * c + 1] + T[3 * r + 2] * k.oMi[i].R[3 * c + 2];
      double * I6 = &k.I6[36 * i];
      inertia6(d.mass[i], cw, Icw, I6);
      mat6vec(I6, vi, &k.h[6 * i]);
      double Ia[6], vxh[6];
      mat6vec(I6, ai, Ia);
      mxf(vi, &k.h[6 * i], vxh);
      for (int r = 0; r < 6; r++)
        k.f[6 * i + r] = Ia[r] + vxh[r];
      // dI = v x* I - I v x
      double Xf[36], Xm[36], t1[36], t2[36];
      crf6(vi, Xf);
      crm6(vi, Xm);
      mat6mul(Xf, I6, t1);
      mat6mul(I6, Xm, t2);
      for (int r = 0; r < 36; r++)
        k.dI6[36 * i + r] = t1[r] - t2[r];
    }
    // backward accumulation
    k.Y = k.I6;
    k.dY = k.dI6;
    k.hs = k.h;
    k.fs = k.f;
    k.msub.assign(nj, 0);
    for (int i = 0; i < nj; i++)
      k.msub[i] = d.mass[i];
    for (int i = nj - 1; i > 0; i--)
    {
      int p = d.parent[i];
      for (int r = 0; r < 36; r++)
      {
        k.Y[36 * p + r] += k.Y[36 * i + r];
        k.dY[36 * p + r] += k.dY[36 * i + r];
      }
      for (int r = 0; r < 6; r++)
      {
        k.hs[6 * p + r] += k.hs[6 * i + r];
        k.fs[6 * p + r] += k.fs[6 * i + r];
      }
      k.msub[p] += k.msub[i];
    }
    k.csub.assign(3 * nj, 0);
    for (int i = 0; i < nj; i++)
    {
      // first moment m c sits in the lower-left block of Y: m [c]x -> (Y[5][1]? ) read from skew entries
      const double * Y = &k.Y[36 * i];
      double mc[3] = {Y[6 * 5 + 1], Y[6 * 3 + 2], Y[6 * 4 + 0]}; // m [c]x : (2,1)=m cx, (0,2)=m cy, (1,0)=m cz
      for (int r = 0; r < 3; r++)
        k.csub[3 * i + r] = mc[r] / k.msub[i];
    }
    const double M = k.msub[0];
    for (int r = 0; r < 3; r++)
      k.com[r] = k.csub[r];
    for (int r = 0; r < 6; r++)
    {
      k.h0[r] = k.hs[r];
      k.hd0[r] = k.fs[r];
    }
    for (int r = 0; r < 3; r++)
      k.vcom[r] = k.h0[r] / M;
    // shift to the centre of mass: [f; n - c x f]
    auto shift = [&](const double * in, double * out) {
      double cxf[3];
      cross3(k.com, in, cxf);
      for (int r = 0; r < 3; r++)
      {
        out[r] = in[r];
        out[3 + r] = in[3 + r] - cxf[r];
      }
    };
    shift(k.h0, k.hg);
    shift(k.hd0, k.hdg);
    // centroidal momentum matrix and CoM Jacobian
    k.Ag.resize(6, nv);
    k.Jcom.resize(3, nv);
    for (int j = 0; j < nv; j++)
    {
      int i = m.joint_of_dof[j];
      double Jc[6], A0[6], Ag[6];
      for (int r = 0; r < 6; r++)
        Jc[r] = k.J(r, j);
      mat6vec(&k.Y[36 * i], Jc, A0);
      shift(A0, Ag);
      for (int r = 0; r < 6; r++)
        k.Ag(r, j) = Ag[r];
      double wxc[3];
      cross3(Jc + 3, &k.csub[3 * i], wxc);
      for (int r = 0; r < 3; r++)
        k.Jcom(r, j) = (k.msub[i] / M) * (Jc[r] + wxc[r]);
    }
    if (!derivs)
      return;
    k.dh_dq.resize(6, nv);
    k.dhd_dq.resize(6, nv);
    k.dhd_dv.resize(6, nv);
    for (int j = 0; j < nv; j++)
    {
      int i = m.joint_of_dof[j];
      int p = d.parent[i];
      double zero6[6] = {0, 0, 0, 0, 0, 0};
      const double * vp = (p >= 0) ? &k.ov[6 * p] : zero6;
      const double * ap = (p >= 0) ? &k.oa[6 * p] : zero6;
      const double * vi = &k.ov[6 * i];
      const double * Y = &k.Y[36 * i];
      const double * dY = &k.dY[36 * i];
      const double * hs = &k.hs[6 * i];
      const double * fs = &k.fs[6 * i];
      double Jc[6];
      for (int r = 0; r < 6; r++)
        Jc[r] = k.J(r, j);
      // dh0/dq_j = J x* hs + Y (vp x J)
      double t1[6], w[6], t2[6], dh0[6];
      mxf(Jc, hs, t1);
      mxm(vp, Jc, w);
      mat6vec(Y, w, t2);
      for (int r = 0; r < 6; r++)
        dh0[r] = t1[r] + t2[r];
      // dhd0/dv_j = dY J + Y ((vp + vi) x J) + J x* hs
      double vpi[6], t3[6], t4[6], t5[6], dhdv0[6];
      for (int r = 0; r < 6; r++)
        vpi[r] = vp[r] + vi[r];
      mat6vec(dY, Jc, t3);
      mxm(vpi, Jc, t4);
      mat6vec(Y, t4, t5);
      for (int r = 0; r < 6; r++)
        dhdv0[r] = t3[r] + t5[r] + t1[r];
      // dhd0/dq_j = J x* fs + Y (ap x J) + dY w - Y (w x vp) + w x* hs
      double u1[6], u2[6], u3[6], u4[6], u5[6], u6[6], u7[6], dhdq0[6];
      mxf(Jc, fs, u1);
      mxm(ap, Jc, u2);
      mat6vec(Y, u2, u3);
      mat6vec(dY, w, u4);
      mxm(w, vp, u5);
      mat6vec(Y, u5, u6);
      mxf(w, hs, u7);
      for (int r = 0; r < 6; r++)
        dhdq0[r] = u1[r] + u3[r] + u4[r] - u6[r] + u7[r];
      // to the centroidal frame; the CoM moves with q
      double o1[6], o2[6], o3[6];
      shift(dh0, o1);
      shift(dhdq0, o2);
      shift(dhdv0, o3);
      double Jcj[3] = {k.Jcom(0, j), k.Jcom(1, j), k.Jcom(2, j)};
      double c1[3], c2[3];
      cross3(Jcj, k.h0, c1);
      cross3(Jcj, k.hd0, c2);
      for (int r = 0; r < 3; r++)
      {
        o1[3 + r] -= c1[r];
        o2[3 + r] -= c2[r];
      }
      for (int r = 0; r < 6; r++)
      {
        k.dh_dq(r, j) = o1[r];
        k.dhd_dq(r, j) = o2[r];
        k.dhd_dv(r, j) = o3[r];
      }
    }
  }

  // LOCAL_WORLD_ALIGNED translational Jacobian column of a point p carried by joint `joint`
  // (zero if dof j is not on the path).  Jp_j = J_lin + J_ang x p
  inline void point_jac_col(const Model & m, const Kin & k, int joint, const double * p, int j, double * out3)
  {
    int i = m.joint_of_dof[j];
    if (!m.anc[joint][i])
    {
      out3[0] = out3[1] = out3[2] = 0;
      return;
    }
    double Jc[6];
    for (int r = 0; r < 6; r++)
      Jc[r] = k.J(r, j);
    double wxp[3];
    cross3(Jc + 3, p, wxp);
    for (int r = 0; r < 3; r++)
      out3[r] = Jc[r] + wxp[r];
  }

  // Frame velocity in the LOCAL frame of foot f and its partial derivatives (6 x nv each)
  // (Pinocchio getFrameVelocityDerivatives(LOCAL) [UPSTREAM]):
  //   v_f = X_f^-1 v_i ;  dv_f/dq_j = X_f^-1 (v_parent(j) x J_j) ;  dv_f/dv_j = X_f^-1 J_j
  inline void frame_velocity_local(const Model & m, const Kin & k, int f, double * vf, Mat * dq, Mat * dv)
  {
    const smpc_model_desc & d = m.d;
    int jf = d.foot_joint[f];
    double X[36];
    se3_actinv_matrix(k.oMf[f], X);
    mat6vec(X, &k.ov[6 * jf], vf);
    if (!dq)
      return;
    dq->resize(6, m.nv);
    dv->resize(6, m.nv);
    for (int j = 0; j < m.nv; j++)
    {
      int i = m.joint_of_dof[j];
      if (!m.anc[jf][i])
        continue;
      int p = d.parent[i];
      double zero6[6] = {0, 0, 0, 0, 0, 0};
      const double * vp = (p >= 0) ? &k.ov[6 * p] : zero6;
      double Jc[6], w[6], o1[6], o2[6];
      for (int r = 0; r < 6; r++)
        Jc[r] = k.J(r, j);
      mxm(vp, Jc, w);
      mat6vec(X, w, o1);
      mat6vec(X, Jc, o2);
      for (int r = 0; r < 6; r++)
      {
        (*dq)(r, j) = o1[r];
        (*dv)(r, j) = o2[r];
      }
    }
  }

  // ---- configuration-space Lie operations on q = [p, quat, joints] (Pinocchio integrate/difference) ----
  inline SE3 base_pose(const double * q)
  {
    SE3 M;
    quat_to_R(q + 3, M.R);
    M.p[0] = q[0];
    M.p[1] = q[1];
    M.p[2] = q[2];
    return M;
  }
  // qout = q (+) dq   (dq in R^nv, free-flyer part in the local frame)
  inline void integrate_q(const Model & m, const double * q, const double * dq, double * qout)
  {
    SE3 M = base_pose(q);
    SE3 E = exp6(dq);
    SE3 Mn = se3_mul(M, E);
    qout[0] = Mn.p[0];
    qout[1] = Mn.p[1];
    qout[2] = Mn.p[2];
    R_to_quat(Mn.R, qout + 3);
    // keep the quaternion in the same hemisphere as the input (continuity of warm starts)
    double dotq = q[3] * qout[3] + q[4] * qout[4] + q[5] * qout[5] + q[6] * qout[6];
    if (dotq < 0)
      for (int i = 3; i < 7; i++)
        qout[i] = -qout[i];
    for (int i = 6; i < m.nv; i++)
      qout[7 + (i - 6)] = q[7 + (i - 6)] + dq[i];
  }
  // d = q1 (-) q0  (R^nv)
  inline void difference_q(const Model & m, const double * q0, const double * q1, double * dd)
  {
    SE3 M0 = base_pose(q0), M1 = base_pose(q1);
    SE3 D = se3_mul(se3_inv(M0), M1);
    log6(D, dd);
    for (int i = 6; i < m.nv; i++)
      dd[i] = q1[7 + (i - 6)] - q0[7 + (i - 6)];
  }

} // namespace orc
