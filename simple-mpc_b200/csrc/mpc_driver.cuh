// mpc_driver.cuh — the part of MPC::iterate that surrounds SolverProxDDP::run, for one MPC instance, on the
// device (§8f rank 1):
//
//   data_handler_->updateInternalData(x, false)     reference src/mpc.cpp:192, src/robot-handler.cpp:114-127
//   recedeWithCycle()                               reference src/mpc.cpp:220-254 (+ updateCycleTiming :256-276)
//   updateStepTrackerReferences()                   reference src/mpc.cpp:278-324
//       FootTrajectory::updateTrajectory            reference src/foot-trajectory.cpp:41-96
//   x0_ = getProblemState(data_handler)             reference src/mpc.cpp:199
//
// On the CPU these are O(T nk) string-keyed setter calls per instance; with thousands of instances the foot
// references ([B][T][nk][3] doubles per tick) would have to cross PCIe every tick. Here every instance keeps
// its gait bookkeeping (take-off / landing queues, cycle position) and its swing-curve end points in HBM and
// one CTA per instance regenerates the references straight into the slot ring.
//
// Two routines, both written in the SMPC_PAR phase style (par.cuh) so that tests/host_emul can run them:
//   mpc_recede : integer bookkeeping; decides which stage-table entry becomes the new horizon tail
//   mpc_refs   : forward kinematics of the measured state, Raibert foothold, degree-8 Bezier swing curve,
//                T references per foot, reference state / base velocity of the last stage, terminal CoM
//                reference, problem state x0
#pragma once
#include "dev_math.cuh"
#include "dev_types.h"

namespace smpc
{

  constexpr int SMPC_MPC_QUEUE = 8; // pending take-off / landing events kept per foot

  // MPCSettings fields the driver needs (reference include/simple-mpc/mpc.hpp:36-58) + derived constants
  struct MpcShared
  {
    int T_fly, T_contact;
    double swing_apex, timestep;
    int standing_id; // stage-table index of the standing stage (reference standing_horizon_), -1 if none
    double com0[3];  // CoM at the reference state (reference src/mpc.cpp:93)
  };

  // Per-instance driver state: the members of MPC / FootTrajectory that change from tick to tick.
  struct MpcInst
  {
    int land[SMPC_MAX_FEET][SMPC_MPC_QUEUE], take[SMPC_MAX_FEET][SMPC_MPC_QUEUE]; // foot_land_times_, foot_takeoff_times_
    int nland[SMPC_MAX_FEET], ntake[SMPC_MAX_FEET];
    int cyc;      // stage-table index of cycle_horizon_[0] (the next stage to enter the horizon)
    int walking;  // now_ == WALKING
    int overflow; // a queue overflowed (reported by smpc_mpc_iterate, never silently dropped)
    int pad_;
    double init_pose[SMPC_MAX_FEET][3], final_pose[SMPC_MAX_FEET][3]; // FootTrajectory::initial_poses_/final_poses_
    double velocity_base[6];                                            // MPC::velocity_base_
    double x_reference[SMPC_MAX_NX];                                    // MPC::x_reference_
  };

  struct MpcScratch
  {
    double R[SMPC_MAX_JOINTS][9], p[SMPC_MAX_JOINTS][3];
    double foot_p[SMPC_MAX_FEET][3], footref_p[SMPC_MAX_FEET][3], base_p[3];
    double next_pose[SMPC_MAX_FEET][3];
    double cp_init[SMPC_MAX_FEET][3], cp_mid[SMPC_MAX_FEET][3], cp_final[SMPC_MAX_FEET][3];
    int land_time[SMPC_MAX_FEET];
    double com[3], hg[6];
  };

  // ---- integer bookkeeping --------------------------------------------------------------------------------
  SMPC_HD void mpc_queue_push(int * q, int & n, int value, int & overflow)
  {
    if (n < SMPC_MPC_QUEUE)
      q[n++] = value;
    else
      overflow = 1;
  }
  // one vector of updateCycleTiming (reference src/mpc.cpp:256-276)
  SMPC_HD void mpc_queue_tick(int * q, int & n, bool only_horizon, int T)
  {
    for (int k = 0; k < n; k++)
      if (!only_horizon || q[k] < T)
        q[k] -= 1;
    if (n > 0 && q[0] < 0)
    {
      for (int k = 0; k + 1 < n; k++)
        q[k] = q[k + 1];
      n -= 1;
    }
  }

  // generateCycleHorizon's timing tables (reference src/mpc.cpp:112-132) for a cycle that starts at table entry
  // `phase`: masks[k] is the contact mask of table entry k, the cycle seen by this instance is masks[(phase+i) % ncycle].
  SMPC_HD void mpc_init_queues(const unsigned * masks, int ncycle, int nk, int T, int phase, MpcInst & I)
  {
    for (int f = 0; f < nk; f++)
    {
      I.nland[f] = I.ntake[f] = 0;
      for (int i = 1; i < ncycle; i++)
      {
        const bool now = (masks[(phase + i) % ncycle] >> f) & 1u, before = (masks[(phase + i - 1) % ncycle] >> f) & 1u;
        if (!now && before)
          mpc_queue_push(I.take[f], I.ntake[f], i + T, I.overflow);
        if (now && !before)
          mpc_queue_push(I.land[f], I.nland[f], i + T, I.overflow);
      }
      // wrap-around switch between the last and the first slot of the cycle (reference src/mpc.cpp:126-131)
      const bool back = (masks[(phase + ncycle - 1) % ncycle] >> f) & 1u, front = (masks[phase % ncycle] >> f) & 1u;
      if (back && !front)
        mpc_queue_push(I.take[f], I.ntake[f], ncycle - 1 + T, I.overflow);
      if (!back && front)
        mpc_queue_push(I.land[f], I.nland[f], ncycle - 1 + T, I.overflow);
    }
    I.cyc = phase % ncycle;
  }

  // recedeWithCycle (reference src/mpc.cpp:220-254) for one instance, one thread.
  //   tail_mask : contact mask of the stage currently at horizon position T-1 (before the rotation)
  //   table     : the cycle horizon (reference generateCycleHorizon), ntable entries, + the standing stage
  // returns the stage-table index of the new tail.
  SMPC_HD int mpc_recede(const MpcShared & S, MpcInst & I, const smpc_slot * table, int ncycle, int nk, int T, unsigned tail_mask)
  {
    int support = 0;
    for (int f = 0; f < nk; f++)
      support += (tail_mask >> f) & 1u;
    if (I.walking || support < nk || S.standing_id < 0)
    {
      const int id = I.cyc;
      const int prev = (id + ncycle - 1) % ncycle;
      // after the rotation contact_states_[n-1] = table[id], contact_states_[n-2] = table[id-1]
      const unsigned c1 = table[id].contact_mask, c2 = table[prev].contact_mask;
      for (int f = 0; f < nk; f++)
      {
        const bool now = (c1 >> f) & 1u, before = (c2 >> f) & 1u;
        if (!now && before)
          mpc_queue_push(I.take[f], I.ntake[f], ncycle + T, I.overflow);
        if (now && !before)
          mpc_queue_push(I.land[f], I.nland[f], ncycle + T, I.overflow);
      }
      for (int f = 0; f < nk; f++)
      {
        mpc_queue_tick(I.land[f], I.nland[f], false, T);
        mpc_queue_tick(I.take[f], I.ntake[f], false, T);
      }
      I.cyc = (id + 1) % ncycle;
      return id;
    }
    for (int f = 0; f < nk; f++)
    {
      mpc_queue_tick(I.land[f], I.nland[f], true, T);
      mpc_queue_tick(I.take[f], I.ntake[f], true, T);
    }
    return S.standing_id;
  }

  // ---- kinematics of the measured state --------------------------------------------------------------------
  // Joint placements in the world frame (one thread; joints are topologically ordered).
  SMPC_DEV void mpc_fk(const smpc_model_desc & d, const double * q, double (*R)[9], double (*p)[3])
  {
    for (int i = 0; i < d.njoints; i++)
    {
      double Rj[9], pj[3] = {0, 0, 0};
      if (d.jtype[i] == SMPC_JOINT_FREEFLYER)
      {
        quat_to_R(q + d.idx_q[i] + 3, Rj);
        for (int k = 0; k < 3; k++)
          pj[k] = q[d.idx_q[i] + k];
      }
      else
      {
        const double ang = q[d.idx_q[i]];
        const double sn = sin(ang), cs = cos(ang);
        double W[9], W2[9];
        skew(d.axis[i], W);
        m3m(W, W, W2);
        for (int k = 0; k < 9; k++)
          Rj[k] = ((k % 4 == 0) ? 1.0 : 0.0) + sn * W[k] + (1 - cs) * W2[k];
      }
      double Rl[9], pl[3];
      m3m(d.place_R[i], Rj, Rl);
      m3v(d.place_R[i], pj, pl);
      for (int k = 0; k < 3; k++)
        pl[k] += d.place_p[i][k];
      const int par = d.parent[i];
      if (par < 0)
      {
        for (int k = 0; k < 9; k++)
          R[i][k] = Rl[k];
        for (int k = 0; k < 3; k++)
          p[i][k] = pl[k];
      }
      else
      {
        m3m(R[par], Rl, R[i]);
        double t[3];
        m3v(R[par], pl, t);
        for (int k = 0; k < 3; k++)
          p[i][k] = t[k] + p[par][k];
      }
    }
  }

  // CoM and centroidal momentum [linear; angular about the CoM, world axes] (one thread).
  // Stands in for pinocchio::computeCentroidalMomentum behind RobotDataHandler::getCentroidalState
  // (reference src/robot-handler.cpp:119-127,141-149).
  SMPC_DEV void mpc_centroidal_state(const smpc_model_desc & d, const double * q, const double * v, const double (*R)[9],
                                     const double (*p)[3], double * com, double * hg)
  {
    double vel[SMPC_MAX_JOINTS][6]; // spatial velocity of body i: [linear at the world origin; angular]
    double lin[3] = {0, 0, 0}, ang0[3] = {0, 0, 0}, mc[3] = {0, 0, 0}, mass = 0;
    for (int i = 0; i < d.njoints; i++)
    {
      const int par = d.parent[i];
      const int iv = d.idx_v[i];
      double vi[6];
      for (int r = 0; r < 6; r++)
        vi[r] = par >= 0 ? vel[par][r] : 0.0;
      if (d.jtype[i] == SMPC_JOINT_FREEFLYER)
      {
        for (int c = 0; c < 6; c++)
        {
          const double axc[3] = {R[i][c % 3], R[i][3 + c % 3], R[i][6 + c % 3]};
          if (c < 3)
          {
            for (int r = 0; r < 3; r++)
              vi[r] += axc[r] * v[iv + c];
          }
          else
          {
            double t[3];
            v3cross(p[i], axc, t);
            for (int r = 0; r < 3; r++)
            {
              vi[r] += t[r] * v[iv + c];
              vi[3 + r] += axc[r] * v[iv + c];
            }
          }
        }
      }
      else
      {
        double aw[3], t[3];
        m3v(R[i], d.axis[i], aw);
        v3cross(p[i], aw, t);
        for (int r = 0; r < 3; r++)
        {
          vi[r] += t[r] * v[iv];
          vi[3 + r] += aw[r] * v[iv];
        }
      }
      for (int r = 0; r < 6; r++)
        vel[i][r] = vi[r];
      // momentum of body i about the world origin
      const double m = d.mass[i];
      double cw[3], T[9], Ic[9], vc[3], Iw[3], t[3];
      m3v(R[i], d.lever[i], cw);
      for (int k = 0; k < 3; k++)
        cw[k] += p[i][k];
      m3m(R[i], d.inertia[i], T);
      m3mt(T, R[i], Ic);
      v3cross(vi + 3, cw, vc);
      for (int k = 0; k < 3; k++)
        vc[k] = m * (vc[k] + vi[k]);
      m3v(Ic, vi + 3, Iw);
      v3cross(cw, vc, t);
      for (int k = 0; k < 3; k++)
      {
        lin[k] += vc[k];
        ang0[k] += Iw[k] + t[k];
        mc[k] += m * cw[k];
      }
      mass += m;
    }
    double t[3];
    for (int k = 0; k < 3; k++)
      com[k] = mc[k] / mass;
    v3cross(com, lin, t);
    for (int k = 0; k < 3; k++)
    {
      hg[k] = lin[k];
      hg[3 + k] = ang0[k] - t[k];
    }
  }

  // degree-8 Bezier on [0,1] in the form the reference's curve library evaluates it
  // (ndcurves::bezier_curve::evalHorner [UPSTREAM], control points of reference src/foot-trajectory.cpp:43-56:
  // 4 x init, mid, 4 x final). The curve parameter is computed in fp32 (reference src/foot-trajectory.cpp:76).
  SMPC_DEV void mpc_bezier(const double * init, const double * mid, const double * fin, double u, double * out)
  {
    const double u_op = 1.0 - u;
    double bc = 1, tn = 1;
    double tmp[3] = {init[0] * u_op, init[1] * u_op, init[2] * u_op};
    for (int i = 1; i < 8; i++)
    {
      tn = tn * u;
      bc = bc * (double)(8 - i + 1) / i;
      const double * cp = (i < 4) ? init : (i == 4 ? mid : fin);
      for (int k = 0; k < 3; k++)
        tmp[k] = (tmp[k] + cp[k] * (tn * bc)) * u_op;
    }
    for (int k = 0; k < 3; k++)
      out[k] = tmp[k] + fin[k] * (tn * u);
  }

  // Ring access handed to mpc_refs: slot(t) is the stage at horizon position t.
  struct MpcRing
  {
    smpc_slot * base;
    int head, N;
    SMPC_DEV smpc_slot & at(int t) const { return base[(head + t) % N]; }
  };

  // updateInternalData + updateStepTrackerReferences + getProblemState for one instance (one CTA).
  //   xmeas : measured multibody state [nq + nv]
  //   x0    : problem state written for the solver (nx doubles: xmeas itself, or [com; hg] for the centroidal OCP)
  SMPC_DEV void mpc_refs(const DevModel & M, const DevOcp & O, const MpcShared & S, MpcInst & I, const double * xmeas, const MpcRing ring,
                         smpc_terminal & term, double * x0, MpcScratch & s)
  {
    const smpc_model_desc & d = M.d;
    const int nk = d.nfeet, T = O.N;
    // ---- phase 1: kinematics, Raibert footholds, swing-curve end points (feet in parallel after the FK)
    SMPC_PAR(tid)
    if (tid == 0)
    {
      mpc_fk(d, xmeas, s.R, s.p);
      for (int f = 0; f < nk; f++)
      {
        double t[3];
        m3v(s.R[d.foot_joint[f]], d.foot_p[f], t);
        for (int k = 0; k < 3; k++)
          s.foot_p[f][k] = t[k] + s.p[d.foot_joint[f]][k];
        m3v(s.R[d.footref_joint[f]], d.footref_p[f], t);
        for (int k = 0; k < 3; k++)
          s.footref_p[f][k] = t[k] + s.p[d.footref_joint[f]][k];
      }
      double t[3];
      m3v(s.R[d.base_joint], d.base_p, t);
      for (int k = 0; k < 3; k++)
        s.base_p[k] = t[k] + s.p[d.base_joint][k];
      if (O.kind == SMPC_OCP_CENTROIDAL)
        mpc_centroidal_state(d, xmeas, xmeas + d.nq, s.R, s.p, s.com, s.hg);
    }
    SMPC_END
    SMPC_PAR(tid)
    if (tid < nk)
    {
      const int f = tid;
      const int land_time = I.nland[f] > 0 ? I.land[f][0] : -1;
      // the swing curve is re-planned only while the landing is at least one flight time away (reference src/mpc.cpp:287-289)
      const bool update = !(land_time < S.T_fly);
      const double * ref = s.footref_p[f];
      const double twist0 = -(ref[1] - s.base_p[1]);
      const double twist1 = ref[0] - s.base_p[0];
      const double horizon_time = (S.T_fly + S.T_contact) * S.timestep;
      double next[3];
      next[0] = ref[0] + (I.velocity_base[0] + I.velocity_base[5] * twist0) * horizon_time;
      next[1] = ref[1] + (I.velocity_base[1] + I.velocity_base[5] * twist1) * horizon_time;
      next[2] = s.foot_p[f][2];
      if (update)
        for (int k = 0; k < 3; k++)
        {
          I.init_pose[f][k] = s.foot_p[f][k];
          I.final_pose[f][k] = next[k];
        }
      for (int k = 0; k < 3; k++)
      {
        s.cp_init[f][k] = I.init_pose[f][k];
        s.cp_final[f][k] = I.final_pose[f][k];
        s.cp_mid[f][k] = I.init_pose[f][k] * 3 / 4 + I.final_pose[f][k] * 1 / 4;
      }
      s.cp_mid[f][2] += S.swing_apex;
      s.land_time[f] = land_time;
    }
    SMPC_END
    // ---- phase 2: T references per foot (reference src/foot-trajectory.cpp:64-82), one (foot, stage) pair per thread
    SMPC_PAR(tid)
    for (int e = tid; e < nk * T; e += SMPC_NT)
    {
      const int f = e / T, time = e % T;
      const int t = s.land_time[f] - time;
      double pt[3];
      if (t < 0)
        for (int k = 0; k < 3; k++)
          pt[k] = s.cp_final[f][k];
      else if (t > S.T_fly)
        for (int k = 0; k < 3; k++)
          pt[k] = s.cp_init[f][k];
      else
      {
        const float u = float(S.T_fly - t) / float(S.T_fly);
        mpc_bezier(s.cp_init[f], s.cp_mid[f], s.cp_final[f], (double)u, pt);
      }
      smpc_slot & sl = ring.at(time);
      // setReferencePose: translation always; the multibody handlers also store the (identity) rotation of a 6-D contact
      if (O.fs == 6 && O.kind != SMPC_OCP_CENTROIDAL)
        for (int k = 0; k < 9; k++)
          sl.foot_ref[f][k] = (k % 4 == 0) ? 1.0 : 0.0;
      for (int k = 0; k < 3; k++)
        sl.foot_ref[f][9 + k] = pt[k];
      if (time == T - 1)
        for (int k = 0; k < 3; k++)
          s.next_pose[f][k] = pt[k]; // references_[name].back()
    }
    SMPC_END
    // ---- phase 3: reference state / base velocity of the last stage, terminal CoM reference, problem state
    SMPC_PAR(tid)
    {
      smpc_slot & last = ring.at(T - 1);
      if (O.kind == SMPC_OCP_CENTROIDAL)
      {
        // setReferenceState then setVelocityBase: momentum references = velocity * mass for both parts
        // (reference src/centroidal-dynamics.cpp:237-238)
        for (int k = tid; k < 9; k += SMPC_NT)
          last.x_ref[k] = (k < 3) ? I.x_reference[k] : I.velocity_base[k - 3] * O.mass;
        for (int k = tid; k < 9; k += SMPC_NT)
          x0[k] = (k < 3) ? s.com[k] : s.hg[k - 3];
      }
      else
      {
        for (int k = tid; k < O.nx; k += SMPC_NT)
          last.x_ref[k] = (k >= d.nq && k < d.nq + 6) ? I.velocity_base[k - d.nq] : I.x_reference[k];
        for (int k = tid; k < O.nx; k += SMPC_NT)
          x0[k] = xmeas[k];
      }
      if (tid == 0 && O.terminal_constraint && O.kind != SMPC_OCP_CENTROIDAL)
      {
        double c[3] = {0, 0, 0};
        for (int f = 0; f < nk; f++)
          for (int k = 0; k < 3; k++)
            c[k] = c[k] + s.next_pose[f][k];
        for (int k = 0; k < 3; k++)
          c[k] = c[k] * (1.0 / (double)nk);
        c[2] += S.com0[2];
        for (int k = 0; k < 3; k++)
          term.com_ref[k] = c[k];
      }
    }
    SMPC_END
  }

  // CoM of the robot at configuration q (one thread) — com0_ of reference src/mpc.cpp:93, and the foot placements
  // FootTrajectory starts from (reference src/mpc.cpp:27-38).
  SMPC_DEV void mpc_init_instance(const DevModel & M, const double * xref, MpcShared * S_out, MpcInst & I, MpcScratch & s)
  {
    const smpc_model_desc & d = M.d;
    mpc_fk(d, xref, s.R, s.p);
    for (int f = 0; f < d.nfeet; f++)
    {
      double t[3];
      m3v(s.R[d.foot_joint[f]], d.foot_p[f], t);
      for (int k = 0; k < 3; k++)
        I.init_pose[f][k] = I.final_pose[f][k] = t[k] + s.p[d.foot_joint[f]][k];
    }
    if (S_out)
    {
      double mc[3] = {0, 0, 0}, mass = 0;
      for (int i = 0; i < d.njoints; i++)
      {
        double cw[3];
        m3v(s.R[i], d.lever[i], cw);
        for (int k = 0; k < 3; k++)
          mc[k] += d.mass[i] * (cw[k] + s.p[i][k]);
        mass += d.mass[i];
      }
      for (int k = 0; k < 3; k++)
        S_out->com0[k] = mc[k] / mass;
    }
  }

} // namespace smpc
