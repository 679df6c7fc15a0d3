// model.cpp — see include/simple-mpc/model.hpp
#include "simple-mpc/model.hpp"
#include <algorithm>

namespace simple_mpc
{
  namespace
  {
    std::array<double, 9> quatToRot(double x, double y, double z, double w)
    {
      const double n = std::sqrt(x * x + y * y + z * z + w * w);
      x /= n, y /= n, z /= n, w /= n;
      return {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
               2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
    }
    std::array<double, 9> axisAngle(const Vector3d & a, double t)
    {
      const double c = std::cos(t), s = std::sin(t), k = 1 - c;
      return {{c + a[0] * a[0] * k, a[0] * a[1] * k - a[2] * s, a[0] * a[2] * k + a[1] * s, a[1] * a[0] * k + a[2] * s, c + a[1] * a[1] * k,
               a[1] * a[2] * k - a[0] * s, a[2] * a[0] * k - a[1] * s, a[2] * a[1] * k + a[0] * s, c + a[2] * a[2] * k}};
    }
    Vector3d rot(const std::array<double, 9> & R, const Vector3d & x)
    {
      return {{R[0] * x[0] + R[1] * x[1] + R[2] * x[2], R[3] * x[0] + R[4] * x[1] + R[5] * x[2], R[6] * x[0] + R[7] * x[1] + R[8] * x[2]}};
    }
    Vector3d rotT(const std::array<double, 9> & R, const Vector3d & x)
    {
      return {{R[0] * x[0] + R[3] * x[1] + R[6] * x[2], R[1] * x[0] + R[4] * x[1] + R[7] * x[2], R[2] * x[0] + R[5] * x[1] + R[8] * x[2]}};
    }
    // log of a rotation matrix
    Vector3d log3(const std::array<double, 9> & R, double & theta)
    {
      const double tr = R[0] + R[4] + R[8];
      const double c = std::min(1.0, std::max(-1.0, 0.5 * (tr - 1.0)));
      theta = std::acos(c);
      Vector3d w{{R[7] - R[5], R[2] - R[6], R[3] - R[1]}};
      if (theta < 1e-8)
        return w * 0.5;
      if (theta > M_PI - 1e-6)
      {
        // near pi: axis from the diagonal
        Vector3d a;
        for (int i = 0; i < 3; i++)
          a[i] = std::sqrt(std::max(0.0, (R[4 * i] - c) / (1 - c)));
        if (w[0] < 0)
          a[0] = -a[0];
        if (w[1] < 0)
          a[1] = -a[1];
        if (w[2] < 0)
          a[2] = -a[2];
        return a * theta;
      }
      return w * (theta / (2.0 * std::sin(theta)));
    }
  } // namespace

  JointIndex Model::addFreeFlyer(const std::string & name, double mass, const Vector3d & lever, const std::array<double, 9> & inertia)
  {
    if (!joints.empty())
      throw std::runtime_error("Model: the free-flyer must be the first joint");
    Joint j;
    j.name = name;
    j.type = SMPC_JOINT_FREEFLYER;
    j.mass = mass;
    j.lever = lever;
    j.inertia = inertia;
    joints.push_back(j);
    nq = 7;
    nv = 6;
    frames.push_back(Frame(name, 0, SE3::Identity()));
    return 0;
  }

  JointIndex Model::addRevoluteJoint(
    JointIndex parent, const std::string & name, const Vector3d & axis, const SE3 & placement, double mass, const Vector3d & lever,
    const std::array<double, 9> & inertia)
  {
    if (parent >= joints.size())
      throw std::runtime_error("Model: unknown parent joint");
    Joint j;
    j.name = name;
    j.parent = (int)parent;
    j.axis = axis;
    j.placement = placement;
    j.mass = mass;
    j.lever = lever;
    j.inertia = inertia;
    j.idx_q = nq;
    j.idx_v = nv;
    nq += 1;
    nv += 1;
    joints.push_back(j);
    frames.push_back(Frame(name, (int)joints.size() - 1, SE3::Identity()));
    return joints.size() - 1;
  }

  FrameIndex Model::addFrame(const Frame & f)
  {
    if (f.parentJoint < 0 || f.parentJoint >= (int)joints.size())
      throw std::runtime_error("Model: frame attached to an unknown joint");
    frames.push_back(f);
    return frames.size() - 1;
  }

  bool Model::existFrame(const std::string & name) const
  {
    return std::any_of(frames.begin(), frames.end(), [&](const Frame & f) { return f.name == name; });
  }

  FrameIndex Model::getFrameId(const std::string & name) const
  {
    for (std::size_t i = 0; i < frames.size(); i++)
      if (frames[i].name == name)
        return i;
    throw std::runtime_error("Model: unknown frame '" + name + "'");
  }

  double computeTotalMass(const Model & model)
  {
    double m = 0;
    for (auto const & j : model.joints)
      m += j.mass;
    return m;
  }

  void forwardKinematics(const Model & model, Data & data, const VectorXd & q)
  {
    if ((int)q.size() != model.nq)
      throw std::runtime_error("forwardKinematics: q has the wrong size");
    data.oMi.resize(model.joints.size());
    for (std::size_t i = 0; i < model.joints.size(); i++)
    {
      const Joint & j = model.joints[i];
      SE3 M;
      if (j.type == SMPC_JOINT_FREEFLYER)
      {
        M.R = quatToRot(q[3], q[4], q[5], q[6]);
        M.p = {{q[0], q[1], q[2]}};
      }
      else
        M.R = axisAngle(j.axis, q[(std::size_t)j.idx_q]);
      const SE3 liMi = j.placement * M;
      data.oMi[i] = (j.parent < 0) ? liMi : data.oMi[(std::size_t)j.parent] * liMi;
    }
  }

  void updateFramePlacements(const Model & model, Data & data)
  {
    data.oMf.resize(model.frames.size());
    for (std::size_t f = 0; f < model.frames.size(); f++)
      data.oMf[f] = data.oMi[(std::size_t)model.frames[f].parentJoint] * model.frames[f].placement;
  }

  void computeCentroidalMomentum(const Model & model, Data & data, const VectorXd & q, const VectorXd & v)
  {
    (void)q; // placements come from the last forwardKinematics, as in pinocchio
    const std::size_t nj = model.joints.size();
    std::vector<Vector3d> w(nj), vo(nj), c(nj), vc(nj);
    double mass = 0;
    Vector3d mc{{0, 0, 0}};
    for (std::size_t i = 0; i < nj; i++)
    {
      const Joint & j = model.joints[i];
      const SE3 & M = data.oMi[i];
      if (j.type == SMPC_JOINT_FREEFLYER)
      {
        vo[i] = rot(M.R, {{v[0], v[1], v[2]}});
        w[i] = rot(M.R, {{v[3], v[4], v[5]}});
      }
      else
      {
        const std::size_t p = (std::size_t)j.parent;
        w[i] = w[p] + rot(M.R, j.axis) * v[(std::size_t)j.idx_v];
        vo[i] = vo[p] + cross(w[p], M.p - data.oMi[p].p);
      }
      c[i] = M.act(j.lever);
      vc[i] = vo[i] + cross(w[i], c[i] - M.p);
      mass += j.mass;
      mc = mc + c[i] * j.mass;
    }
    const Vector3d com = mc * (1.0 / mass);
    data.com.assign(1, com);
    Vector3d hl{{0, 0, 0}}, ha{{0, 0, 0}};
    for (std::size_t i = 0; i < nj; i++)
    {
      const Joint & j = model.joints[i];
      const SE3 & M = data.oMi[i];
      hl = hl + vc[i] * j.mass;
      const Vector3d wl = rotT(M.R, w[i]); // R I R^T w
      const Vector3d Iw = rot(j.inertia, wl);
      ha = ha + cross(c[i] - com, vc[i] * j.mass) + rot(M.R, Iw);
    }
    data.hg.lin = hl;
    data.hg.ang = ha;
  }

  VectorXd difference(const Model & model, const VectorXd & q1, const VectorXd & q2)
  {
    if ((int)q1.size() != model.nq || (int)q2.size() != model.nq)
      throw std::runtime_error("difference: configuration of the wrong size");
    VectorXd d((std::size_t)model.nv, 0.0);
    // free-flyer: log6(M1^-1 M2), [linear; angular]
    SE3 M1, M2;
    M1.R = quatToRot(q1[3], q1[4], q1[5], q1[6]);
    M1.p = {{q1[0], q1[1], q1[2]}};
    M2.R = quatToRot(q2[3], q2[4], q2[5], q2[6]);
    M2.p = {{q2[0], q2[1], q2[2]}};
    const SE3 M = M1.actInv(M2);
    double t;
    const Vector3d w = log3(M.R, t);
    // v = V^-1 p,  V^-1 = I - 1/2 [w] + beta [w]^2
    double beta;
    if (t < 1e-4)
      beta = 1.0 / 12 + t * t / 720;
    else
      beta = 1.0 / (t * t) - std::sin(t) / (2 * t * (1 - std::cos(t)));
    const Vector3d wp = cross(w, M.p), wwp = cross(w, wp);
    for (std::size_t k = 0; k < 3; k++)
    {
      d[k] = M.p[k] - 0.5 * wp[k] + beta * wwp[k];
      d[3 + k] = w[k];
    }
    for (std::size_t i = 1; i < model.joints.size(); i++)
      d[(std::size_t)model.joints[i].idx_v] = q2[(std::size_t)model.joints[i].idx_q] - q1[(std::size_t)model.joints[i].idx_q];
    return d;
  }

  VectorXd integrate(const Model & model, const VectorXd & q, const VectorXd & dq)
  {
    if ((int)q.size() != model.nq || (int)dq.size() != model.nv)
      throw std::runtime_error("integrate: configuration / tangent of the wrong size");
    VectorXd out(q);
    // free-flyer: p <- p + R V(w) v, quat <- quat * quat(exp(w)) with dq = [v; w] in the body frame
    const Vector3d v{{dq[0], dq[1], dq[2]}}, w{{dq[3], dq[4], dq[5]}};
    const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], t = std::sqrt(t2);
    double B, Cc, sh, ch; // (1 - cos t)/t^2, (t - sin t)/t^3, sin(t/2)/t, cos(t/2)
    if (t < 1e-4)
    {
      B = 0.5 - t2 / 24;
      Cc = 1.0 / 6 - t2 / 120;
      sh = 0.5 - t2 / 48;
      ch = 1.0 - t2 / 8;
    }
    else
    {
      B = (1 - std::cos(t)) / t2;
      Cc = (t - std::sin(t)) / (t2 * t);
      sh = std::sin(0.5 * t) / t;
      ch = std::cos(0.5 * t);
    }
    const Vector3d wv = cross(w, v), wwv = cross(w, wv);
    const Vector3d pl{{v[0] + B * wv[0] + Cc * wwv[0], v[1] + B * wv[1] + Cc * wwv[1], v[2] + B * wv[2] + Cc * wwv[2]}};
    const std::array<double, 9> R = quatToRot(q[3], q[4], q[5], q[6]);
    const Vector3d dp = rot(R, pl);
    for (std::size_t k = 0; k < 3; k++)
      out[k] = q[k] + dp[k];
    const double ex = sh * w[0], ey = sh * w[1], ez = sh * w[2], ew = ch;
    const double x = q[3], y = q[4], z = q[5], s = q[6];
    double qx = s * ex + x * ew + y * ez - z * ey;
    double qy = s * ey - x * ez + y * ew + z * ex;
    double qz = s * ez + x * ey - y * ex + z * ew;
    double qw = s * ew - x * ex - y * ey - z * ez;
    const double n = std::sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
    out[3] = qx / n, out[4] = qy / n, out[5] = qz / n, out[6] = qw / n;
    for (std::size_t i = 1; i < model.joints.size(); i++)
      out[(std::size_t)model.joints[i].idx_q] = q[(std::size_t)model.joints[i].idx_q] + dq[(std::size_t)model.joints[i].idx_v];
    return out;
  }

  VectorXd interpolate(const Model & model, const VectorXd & q0, const VectorXd & q1, double u)
  {
    VectorXd d = difference(model, q0, q1);
    for (auto & e : d)
      e *= u;
    return integrate(model, q0, d);
  }

  // ---------------------------------------------------------------------------------------------
  // synthetic robots (same numbers as tests/smpc_ctypes.py: go2_like_model / talos_like_model)
  // ---------------------------------------------------------------------------------------------
  namespace
  {
    std::array<double, 9> diag3(double a, double b, double c) { return {{a, 0, 0, 0, b, 0, 0, 0, c}}; }
    SE3 trans(double x, double y, double z)
    {
      SE3 M;
      M.p = {{x, y, z}};
      return M;
    }
  } // namespace

  Model Model::BuildGo2Like()
  {
    Model m;
    m.addFreeFlyer("root_joint", 6.921, {{0.021112, 0.0, -0.005366}}, diag3(0.02448, 0.098077, 0.107));
    m.addFrame(Frame("base", 0, SE3::Identity()));
    const char * names[4] = {"FL", "FR", "RL", "RR"};
    const double sxs[4] = {1, 1, -1, -1}, sys[4] = {1, -1, 1, -1};
    for (int l = 0; l < 4; l++)
    {
      const double sx = sxs[l], sy = sys[l];
      const std::string n = names[l];
      JointIndex hip = m.addRevoluteJoint(0, n + "_hip_joint", {{1, 0, 0}}, trans(sx * 0.1934, sy * 0.0465, 0.0), 0.678,
                                          {{-sx * 0.0054, sy * 0.00194, -0.000105}}, diag3(0.00048, 0.000884, 0.000596));
      JointIndex thigh = m.addRevoluteJoint(hip, n + "_thigh_joint", {{0, 1, 0}}, trans(0.0, sy * 0.0955, 0.0), 1.152,
                                            {{-0.00374, -sy * 0.0223, -0.0327}}, diag3(0.00584, 0.0058, 0.00103));
      JointIndex calf = m.addRevoluteJoint(thigh, n + "_calf_joint", {{0, 1, 0}}, trans(0.0, 0.0, -0.213), 0.194,
                                           {{0.00548, -0.000975, -0.135}}, diag3(0.00108, 0.0011, 3.29e-05));
      m.addFrame(Frame(n + "_foot", (int)calf, trans(0.0, 0.0, -0.213)));
    }
    m.referenceConfigurations["standing"] = VectorXd{0, 0, 0.335, 0, 0, 0, 1, 0.068, 0.785, -1.44, -0.068, 0.785, -1.44,
                                                     0.068, 0.785, -1.44, -0.068, 0.785, -1.44};
    m.lowerPositionLimit = VectorXd{-1.0472, -1.5708, -2.7227, -1.0472, -1.5708, -2.7227, -1.0472, -1.5708, -2.7227, -1.0472, -1.5708, -2.7227};
    m.upperPositionLimit = VectorXd{1.0472, 3.4907, -0.83776, 1.0472, 3.4907, -0.83776, 1.0472, 3.4907, -0.83776, 1.0472, 3.4907, -0.83776};
    m.effortLimit = VectorXd(12, 23.7);
    m.velocityLimit = VectorXd(12, 30.1);
    return m;
  }

  Model Model::BuildTalosLike()
  {
    Model m;
    m.addFreeFlyer("root_joint", 15.36, {{-0.05, 0.0, -0.05}}, diag3(0.20, 0.12, 0.16));
    m.addFrame(Frame("base_link", 0, SE3::Identity()));
    const Vector3d X{{1, 0, 0}}, Y{{0, 1, 0}}, Z{{0, 0, 1}};
    const char * side[2] = {"left", "right"};
    JointIndex soles[2];
    for (int s = 0; s < 2; s++)
    {
      const double sy = s == 0 ? 1 : -1;
      const std::string n = std::string("leg_") + side[s] + "_";
      JointIndex j = m.addRevoluteJoint(0, n + "1_joint", Z, trans(-0.02, sy * 0.085, -0.27), 1.85, {{0.02, sy * 0.01, 0.03}}, diag3(0.004, 0.006, 0.005));
      j = m.addRevoluteJoint(j, n + "2_joint", X, trans(0, 0, 0), 1.98, {{-0.01, sy * 0.005, 0.0}}, diag3(0.004, 0.005, 0.004));
      j = m.addRevoluteJoint(j, n + "3_joint", Y, trans(0, 0, 0), 6.24, {{0.003, sy * 0.033, -0.21}}, diag3(0.15, 0.15, 0.02));
      j = m.addRevoluteJoint(j, n + "4_joint", Y, trans(0, 0, -0.38), 3.76, {{0.011, sy * 0.003, -0.14}}, diag3(0.06, 0.06, 0.008));
      j = m.addRevoluteJoint(j, n + "5_joint", Y, trans(0, 0, -0.325), 1.29, {{-0.01, sy * 0.01, 0.0}}, diag3(0.002, 0.003, 0.002));
      j = m.addRevoluteJoint(j, n + "6_joint", X, trans(0, 0, 0), 1.60, {{-0.004, 0.0, -0.08}}, diag3(0.004, 0.008, 0.008));
      soles[s] = j;
    }
    JointIndex t1 = m.addRevoluteJoint(0, "torso_1_joint", Z, trans(0, 0, 0.0722), 2.29, {{0.0, 0.0, 0.03}}, diag3(0.006, 0.006, 0.003));
    JointIndex t2 = m.addRevoluteJoint(t1, "torso_2_joint", Y, trans(0, 0, 0), 17.55, {{-0.04, 0.0, 0.19}}, diag3(0.45, 0.35, 0.30));
    for (int s = 0; s < 2; s++)
    {
      const double sy = s == 0 ? 1 : -1;
      const std::string n = std::string("arm_") + side[s] + "_";
      JointIndex j = m.addRevoluteJoint(t2, n + "1_joint", Z, trans(0, sy * 0.1575, 0.232), 1.43, {{-0.007, sy * 0.12, -0.02}}, diag3(0.005, 0.002, 0.005));
      j = m.addRevoluteJoint(j, n + "2_joint", X, trans(0.006, sy * 0.1829, -0.03), 1.81, {{0.02, sy * 0.003, -0.06}}, diag3(0.008, 0.008, 0.003));
      j = m.addRevoluteJoint(j, n + "3_joint", Z, trans(0, 0, 0), 1.54, {{0.007, sy * 0.0003, -0.15}}, diag3(0.012, 0.012, 0.002));
      j = m.addRevoluteJoint(j, n + "4_joint", Y, trans(0.02, 0, -0.273), 2.67, {{-0.01, sy * 0.002, -0.12}}, diag3(0.02, 0.02, 0.003));
    }
    for (int s = 0; s < 2; s++)
      m.addFrame(Frame(std::string(side[s]) + "_sole_link", (int)soles[s], trans(0.0, 0.0, -0.107)));
    VectorXd q{0, 0, 1.0, 0, 0, 0, 1, 0, 0, -0.4, 0.8, -0.4, 0, 0, 0, -0.4, 0.8, -0.4, 0, 0, 0.006,
               0.25, 0.17, 0, -0.52, -0.25, -0.17, 0, -0.52};
    Data d(m);
    forwardKinematics(m, d, q);
    updateFramePlacements(m, d);
    q[2] -= d.oMf[m.getFrameId("left_sole_link")].p[2]; // soles on the ground at "half_sitting"
    m.referenceConfigurations["half_sitting"] = q;
    m.lowerPositionLimit = VectorXd(22, -2.5);
    m.upperPositionLimit = VectorXd(22, 2.5);
    m.effortLimit = VectorXd(22, 300.0);
    m.velocityLimit = VectorXd(22, 10.0);
    return m;
  }

} // namespace simple_mpc
