// py-convert.hpp — numpy <-> VectorXd / MatrixXd casters and the dict -> settings converters of the
// reference bindings (reference bindings/expose-mpc.cpp:28-45, expose-kinodynamics.cpp:9-34,
// expose-fulldynamics.cpp:10-45, expose-centroidal.cpp:9-28): a missing key raises KeyError.
#pragma once
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "simple-mpc/centroidal-dynamics.hpp"
#include "simple-mpc/fulldynamics.hpp"
#include "simple-mpc/kinodynamics.hpp"
#include "simple-mpc/mpc.hpp"
#include "simple-mpc/interpolator.hpp"
#include "simple-mpc/friction-compensation.hpp"
#include "simple-mpc/inverse-dynamics/centroidal-id.hpp"

namespace pybind11
{
  namespace detail
  {
    template <>
    struct type_caster<simple_mpc::VectorXd>
    {
      PYBIND11_TYPE_CASTER(simple_mpc::VectorXd, const_name("numpy.ndarray[float64[n]]"));
      bool load(handle src, bool)
      {
        auto arr = array_t<double, array::c_style | array::forcecast>::ensure(src);
        if (!arr || arr.ndim() > 2)
          return false;
        if (arr.ndim() == 2 && arr.shape(0) != 1 && arr.shape(1) != 1)
          return false;
        value.assign(arr.data(), arr.data() + arr.size());
        return true;
      }
      static handle cast(const simple_mpc::VectorXd & v, return_value_policy, handle)
      {
        array_t<double> a((ssize_t)v.size());
        std::copy(v.begin(), v.end(), a.mutable_data());
        return a.release();
      }
    };
    template <>
    struct type_caster<simple_mpc::MatrixXd>
    {
      PYBIND11_TYPE_CASTER(simple_mpc::MatrixXd, const_name("numpy.ndarray[float64[m, n]]"));
      bool load(handle src, bool)
      {
        auto arr = array_t<double, array::c_style | array::forcecast>::ensure(src);
        if (!arr || arr.ndim() != 2)
          return false;
        value = simple_mpc::MatrixXd((size_t)arr.shape(0), (size_t)arr.shape(1));
        std::copy(arr.data(), arr.data() + arr.size(), value.data());
        return true;
      }
      static handle cast(const simple_mpc::MatrixXd & m, return_value_policy, handle)
      {
        array_t<double> a({(ssize_t)m.rows(), (ssize_t)m.cols()});
        std::copy(m.data(), m.data() + m.rows() * m.cols(), a.mutable_data());
        return a.release();
      }
    };
  } // namespace detail
} // namespace pybind11

namespace simple_mpc
{
  namespace python
  {
    namespace py = pybind11;

    inline Vector3d toVec3(const py::handle & h)
    {
      VectorXd v = h.cast<VectorXd>();
      if (v.size() != 3)
        throw std::runtime_error("expected a vector of size 3");
      return {{v[0], v[1], v[2]}};
    }

    inline MPCSettings mpcSettingsFromDict(const py::dict & settings)
    {
      MPCSettings conf;
      conf.support_force = settings["support_force"].cast<double>();
      conf.TOL = settings["TOL"].cast<double>();
      conf.mu_init = settings["mu_init"].cast<double>();
      conf.max_iters = settings["max_iters"].cast<std::size_t>();
      conf.num_threads = settings["num_threads"].cast<std::size_t>();
      conf.swing_apex = settings["swing_apex"].cast<double>();
      conf.T_fly = settings["T_fly"].cast<int>();
      conf.T_contact = settings["T_contact"].cast<int>();
      conf.timestep = settings["timestep"].cast<double>();
      return conf;
    }
    inline py::dict mpcSettingsToDict(const MPCSettings & conf)
    {
      py::dict d;
      d["support_force"] = conf.support_force;
      d["TOL"] = conf.TOL;
      d["mu_init"] = conf.mu_init;
      d["max_iters"] = conf.max_iters;
      d["num_threads"] = conf.num_threads;
      d["swing_apex"] = conf.swing_apex;
      d["T_fly"] = conf.T_fly;
      d["T_contact"] = conf.T_contact;
      d["timestep"] = conf.timestep;
      return d;
    }

    inline KinodynamicsSettings kinoSettingsFromDict(const py::dict & s)
    {
      KinodynamicsSettings c;
      c.timestep = s["timestep"].cast<double>();
      c.w_x = s["w_x"].cast<MatrixXd>();
      c.w_u = s["w_u"].cast<MatrixXd>();
      c.w_cent = s["w_cent"].cast<MatrixXd>();
      c.w_centder = s["w_centder"].cast<MatrixXd>();
      c.w_frame = s["w_frame"].cast<MatrixXd>();
      c.gravity = toVec3(s["gravity"]);
      c.force_size = s["force_size"].cast<int>();
      c.qmin = s["qmin"].cast<VectorXd>();
      c.qmax = s["qmax"].cast<VectorXd>();
      c.mu = s["mu"].cast<double>();
      c.Lfoot = s["Lfoot"].cast<double>();
      c.Wfoot = s["Wfoot"].cast<double>();
      c.kinematics_limits = s["kinematics_limits"].cast<bool>();
      c.force_cone = s["force_cone"].cast<bool>();
      c.land_cstr = s["land_cstr"].cast<bool>();
      return c;
    }
    inline py::dict kinoSettingsToDict(const KinodynamicsSettings & c)
    {
      py::dict d;
      d["timestep"] = c.timestep;
      d["w_x"] = c.w_x;
      d["w_u"] = c.w_u;
      d["w_cent"] = c.w_cent;
      d["w_centder"] = c.w_centder;
      d["w_frame"] = c.w_frame;
      d["gravity"] = VectorXd(c.gravity.begin(), c.gravity.end());
      d["force_size"] = c.force_size;
      d["qmin"] = c.qmin;
      d["qmax"] = c.qmax;
      d["mu"] = c.mu;
      d["Lfoot"] = c.Lfoot;
      d["Wfoot"] = c.Wfoot;
      d["kinematics_limits"] = c.kinematics_limits;
      d["force_cone"] = c.force_cone;
      d["land_cstr"] = c.land_cstr;
      return d;
    }

    inline FullDynamicsSettings fullSettingsFromDict(const py::dict & s)
    {
      FullDynamicsSettings c;
      c.timestep = s["timestep"].cast<double>();
      c.w_x = s["w_x"].cast<MatrixXd>();
      c.w_u = s["w_u"].cast<MatrixXd>();
      c.w_cent = s["w_cent"].cast<MatrixXd>();
      c.w_forces = s["w_forces"].cast<MatrixXd>();
      c.w_frame = s["w_frame"].cast<MatrixXd>();
      c.gravity = toVec3(s["gravity"]);
      c.force_size = s["force_size"].cast<int>();
      c.mu = s["mu"].cast<double>();
      c.Lfoot = s["Lfoot"].cast<double>();
      c.Wfoot = s["Wfoot"].cast<double>();
      c.umin = s["umin"].cast<VectorXd>();
      c.umax = s["umax"].cast<VectorXd>();
      c.qmin = s["qmin"].cast<VectorXd>();
      c.qmax = s["qmax"].cast<VectorXd>();
      c.Kp_correction = s["Kp_correction"].cast<VectorXd>();
      c.Kd_correction = s["Kd_correction"].cast<VectorXd>();
      c.torque_limits = s["torque_limits"].cast<bool>();
      c.kinematics_limits = s["kinematics_limits"].cast<bool>();
      c.force_cone = s["force_cone"].cast<bool>();
      c.land_cstr = s["land_cstr"].cast<bool>();
      return c;
    }
    inline py::dict fullSettingsToDict(const FullDynamicsSettings & c)
    {
      py::dict d;
      d["timestep"] = c.timestep;
      d["w_x"] = c.w_x;
      d["w_u"] = c.w_u;
      d["w_cent"] = c.w_cent;
      d["w_forces"] = c.w_forces;
      d["w_frame"] = c.w_frame;
      d["gravity"] = VectorXd(c.gravity.begin(), c.gravity.end());
      d["force_size"] = c.force_size;
      d["mu"] = c.mu;
      d["Lfoot"] = c.Lfoot;
      d["Wfoot"] = c.Wfoot;
      d["umin"] = c.umin;
      d["umax"] = c.umax;
      d["qmin"] = c.qmin;
      d["qmax"] = c.qmax;
      d["Kp_correction"] = c.Kp_correction;
      d["Kd_correction"] = c.Kd_correction;
      d["torque_limits"] = c.torque_limits;
      d["kinematics_limits"] = c.kinematics_limits;
      d["force_cone"] = c.force_cone;
      d["land_cstr"] = c.land_cstr;
      return d;
    }

    inline CentroidalSettings centroidalSettingsFromDict(const py::dict & s)
    {
      CentroidalSettings c;
      c.timestep = s["timestep"].cast<double>();
      c.w_com = s["w_com"].cast<MatrixXd>();
      c.w_u = s["w_u"].cast<MatrixXd>();
      c.w_linear_mom = s["w_linear_mom"].cast<MatrixXd>();
      c.w_angular_mom = s["w_angular_mom"].cast<MatrixXd>();
      c.w_linear_acc = s["w_linear_acc"].cast<MatrixXd>();
      c.w_angular_acc = s["w_angular_acc"].cast<MatrixXd>();
      c.gravity = toVec3(s["gravity"]);
      c.force_size = s["force_size"].cast<int>();
      c.mu = s["mu"].cast<double>();
      c.Lfoot = s["Lfoot"].cast<double>();
      c.Wfoot = s["Wfoot"].cast<double>();
      return c;
    }
    inline py::dict centroidalSettingsToDict(const CentroidalSettings & c)
    {
      py::dict d;
      d["timestep"] = c.timestep;
      d["w_com"] = c.w_com;
      d["w_u"] = c.w_u;
      d["w_linear_mom"] = c.w_linear_mom;
      d["w_angular_mom"] = c.w_angular_mom;
      d["w_linear_acc"] = c.w_linear_acc;
      d["w_angular_acc"] = c.w_angular_acc;
      d["gravity"] = VectorXd(c.gravity.begin(), c.gravity.end());
      d["force_size"] = c.force_size;
      d["mu"] = c.mu;
      d["Lfoot"] = c.Lfoot;
      d["Wfoot"] = c.Wfoot;
      return d;
    }
  } // namespace python
} // namespace simple_mpc
