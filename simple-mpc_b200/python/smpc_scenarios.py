"""Driver of one C-ABI context (`Backend`) and the synthetic workloads of SURVEY.md §8d (`*Scenario`).

`Backend` issues the C-ABI call sequence that MPC::MPC / MPC::iterate perform (reference src/mpc.cpp:19-99,189-218):
    create -> set_slots/set_terminal/set_warmstart -> solve(100)            (constructor)
    per tick: cycle(new tail) -> [set_foot_refs] -> shift_warmstart -> solve(x, max_iters)
against any library that exports the entry points of include/smpc.h under a prefix ("smpc": libsmpc_cuda.so; the tests
also point it at the CPU oracle, prefix "orc", which mirrors the same signatures).
"""
import ctypes as C
import numpy as np
from smpc_ctypes import (Slot, Terminal, SolveInfo, dbl_p, go2_like_model, go2_kino_ocp, go2_full_ocp, full_slot, standing_slot,
                         fk_numpy, x_ref_of, total_mass, talos_like_model, talos_centroidal_ocp, centroidal_slot, com_numpy,
                         talos_full_ocp, talos_kino_ocp)


def P(a):
    return a.ctypes.data_as(dbl_p)


class Backend:
    def __init__(self, lib, prefix, model, ocp_bundle, opts, batch, device=0):
        self.lib, self.pre = lib, prefix
        self.B = batch
        self.model, self.ocpb = model, ocp_bundle
        self.ctx = C.c_void_p()
        nx, ndx, nu = C.c_int(), C.c_int(), C.c_int()
        self._chk(self.f("problem_dims")(C.byref(model), C.byref(ocp_bundle.desc), C.byref(nx), C.byref(ndx), C.byref(nu)))
        self.nx, self.ndx, self.nu = nx.value, ndx.value, nu.value
        self.N = ocp_bundle.desc.horizon
        if prefix == "smpc":
            rc = lib.smpc_create(C.byref(self.ctx), device, batch, C.byref(model), C.byref(ocp_bundle.desc), C.byref(opts))
        else:
            rc = lib.orc_create(C.byref(self.ctx), batch, C.byref(model), C.byref(ocp_bundle.desc), C.byref(opts))
        self._chk(rc)
        self.ncmax = self.f("get_ncmax")(self.ctx)

    def f(self, name):
        return getattr(self.lib, self.pre + "_" + name)

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("%s error %d: %s" % (self.pre, rc, self.f("last_error")().decode()))

    def close(self):
        if self.ctx:
            self.f("destroy")(self.ctx)
            self.ctx = None

    def set_slots(self, slots):  # ctypes array [B*N]
        self._chk(self.f("set_slots")(self.ctx, 0, self.B, slots))

    def set_terminal(self, terms):
        self._chk(self.f("set_terminal")(self.ctx, 0, self.B, terms))

    def set_warmstart(self, xs, us):
        self._chk(self.f("set_warmstart")(self.ctx, 0, self.B, P(xs), P(us)))

    def cycle(self, tails):
        self._chk(self.f("cycle")(self.ctx, 0, self.B, tails))

    def set_foot_refs(self, refs):
        self._chk(self.f("set_foot_refs")(self.ctx, 0, self.B, P(refs)))

    def shift(self):
        self._chk(self.f("shift_warmstart")(self.ctx, 0, self.B))

    def solve(self, x0, max_iters):
        x0 = np.ascontiguousarray(x0)
        self._chk(self.f("solve")(self.ctx, P(x0), max_iters))

    def results(self, gains=True):
        B, N, nx, nu, ndx = self.B, self.N, self.nx, self.nu, self.ndx
        xs = np.zeros((B, N + 1, nx))
        us = np.zeros((B, N, nu))
        Ks = np.zeros((B, N, nu, ndx)) if gains else None
        ks = np.zeros((B, N, nu)) if gains else None
        self._chk(self.f("get_results")(self.ctx, 0, B, P(xs), P(us), P(Ks) if gains else None, P(ks) if gains else None))
        return xs, us, Ks, ks

    def head(self):
        B, nx, nu, ndx = self.B, self.nx, self.nu, self.ndx
        x1, u0, u1, K0 = np.zeros((B, nx)), np.zeros((B, nu)), np.zeros((B, nu)), np.zeros((B, nu, ndx))
        self._chk(self.f("get_head")(self.ctx, 0, B, P(x1), P(u0), P(u1), P(K0)))
        return x1, u0, u1, K0

    def info(self):
        arr = (SolveInfo * self.B)()
        self._chk(self.f("get_info")(self.ctx, 0, self.B, arr))
        return arr

    def duals(self):
        vs = np.zeros((self.B, self.N + 1, self.ncmax))
        lams = np.zeros((self.B, self.N + 1, self.ndx))
        self._chk(self.f("get_duals")(self.ctx, 0, self.B, P(vs), P(lams)))
        return vs, lams

    def xdot(self, t):
        out = np.zeros((self.B, self.ndx))
        self._chk(self.f("get_xdot")(self.ctx, 0, self.B, t, P(out)))
        return out


class Go2Scenario:
    """Config C1/C5 of SURVEY.md §8d on the synthetic go2-like robot: kinodynamics, T = 50, trot 10/30/10/30,
    per-instance gait phase offset and state perturbation (numpy default_rng(seed), instance i = stream i)."""

    GAIT = [0b1111] * 10 + [0b0110] * 30 + [0b1111] * 10 + [0b1001] * 30  # FL=bit0 FR=bit1 RL=bit2 RR=bit3
    CYCLE = 80
    T_FLY, T_CONTACT = 30, 10       # MPCSettings of reference examples/go2_kinodynamics.py:96-106
    multibody_state = True          # the problem state is the measured multibody state

    def __init__(self, batch, seed=0, T=50, first=0, **ocp_kw):
        self.B, self.T = batch, T
        self.model, self.limits = go2_like_model()
        self.ocpb = go2_kino_ocp(self.model, self.limits, T=T, **ocp_kw)
        self.nq, self.nv = self.model.nq, self.model.nv
        self.nx, self.ndx, self.nu = self.nq + self.nv, 2 * self.nv, self.nv - 6 + 12
        self.xref = x_ref_of(self.model)
        self.feet = fk_numpy(self.model, self.xref[: self.nq])["feet"]
        self.mass = total_mass(self.model)
        rngs = [np.random.default_rng([seed, first + i]) for i in range(batch)]  # global instance index = stream
        self.phase = np.array([int(r.integers(0, 80)) for r in rngs])
        self.noise = [r for r in rngs]
        self.tick = 0
        # commanded base velocity per instance: vx, vy, yaw rate; its own stream so that the state
        # noise streams above are unchanged
        vr = [np.random.default_rng([seed, first + i, 1]) for i in range(batch)]
        self.vbase = np.zeros((batch, 6))
        for i, r in enumerate(vr):
            self.vbase[i, 0], self.vbase[i, 1], self.vbase[i, 5] = r.uniform(-0.3, 0.3), r.uniform(-0.1, 0.1), r.uniform(-0.3, 0.3)

    def stage_table(self):
        """cycle horizon + the standing stage, as MPC::generateCycleHorizon / standing_horizon_ build them"""
        n = len(self.GAIT)
        return (Slot * (n + 1))(*([self.gait_slot(k) for k in range(n)] + [self.gait_slot(0)])), n

    def xmeas_start(self):
        """measured multibody states of the first tick"""
        return self.x_start()

    def next_measurement(self, xm, x1, noise):
        """closed loop without a plant: the next measured state is the predicted xs[1] (+ velocity noise)"""
        x = np.array(x1, float)
        if noise is not None:
            x[:, self.nq:] += noise
        return x

    def mean_active_rows(self):
        """mean number of constraint rows per stage over the gait cycle (Riccati work count)"""
        d = self.ocpb.desc
        rows = []
        for k in range(len(self.GAIT)):
            mask, prev = self.GAIT[k], self.GAIT[k - 1]
            nst, nland = bin(mask).count("1"), bin(mask & ~prev).count("1")
            rows.append(self.stage_rows(d, nst, nland))
        return float(np.mean(rows))

    def stage_rows(self, d, nst, nland):
        return (self.nv - 6) * d.kinematics_limits + nst * (3 + 2 * d.force_cone) + nland * d.land_cstr

    def gait_slot(self, k):
        mask = self.GAIT[k % 80]
        prev = self.GAIT[(k - 1) % 80]
        land = mask & ~prev
        nact = bin(mask).count("1")
        return standing_slot(self.model, self.ocpb, self.feet, contact_mask=mask, land_mask=land,
                             force_z=self.mass * 9.81 / max(nact, 1))

    def initial_slots(self):
        # createProblem: T all-feet-in-contact stages with Identity foot poses (reference src/ocp-handler.cpp:111-126)
        ident = [(np.eye(3), np.zeros(3)) for _ in range(4)]
        s = standing_slot(self.model, self.ocpb, ident)
        return (Slot * (self.B * self.T))(*([s] * (self.B * self.T)))

    def terminals(self):
        t = Terminal()
        t.x_ref[: self.nx] = self.xref
        t.com_ref[:] = [0.0, 0.0, 0.3]
        t.tau = np.sqrt(0.3 / 9.81)
        return (Terminal * self.B)(*([t] * self.B))

    def warmstart(self):
        s0 = standing_slot(self.model, self.ocpb, self.feet)
        xs = np.tile(self.xref, (self.B, self.T + 1, 1)).copy()
        us = np.tile(np.array(s0.u_ref[: self.nu]), (self.B, self.T, 1)).copy()
        return xs, us

    def x_start(self):
        """perturbed initial states"""
        xs = np.tile(self.xref, (self.B, 1))
        for i, r in enumerate(self.noise):
            xs[i, 7:19] += r.normal(size=12) * 0.02
            xs[i, 0:3] += r.normal(size=3) * 0.01
            xs[i, 19:] += r.normal(size=18) * 0.05
        return xs

    def tails(self, tick):
        arr = (Slot * self.B)()
        for i in range(self.B):
            arr[i] = self.gait_slot(int(self.phase[i]) + tick)
        return arr

    def disturb(self, x, tick):
        """measured state = predicted next state + a small disturbance (same for both backends)"""
        for i, r in enumerate(self.noise):
            x[i, 19:] += r.normal(size=18) * (0.05 if tick == 0 else 0.01)
            if tick == 0:
                x[i, 7:19] += r.normal(size=12) * 0.02
        return x


class Go2FullScenario(Go2Scenario):
    """Config C3 of SURVEY.md §8d on the synthetic go2-like robot: FullDynamicsOCP (rigid point contacts), T = 50, trot
    10/30/10/30, torque + joint limits, and (by default) the friction-cone and landing constraints switched on
    (reference examples/go2_fulldynamics.py:43-77 with force_cone / land_cstr flipped, src/fulldynamics.cpp:185-209)."""

    def __init__(self, batch, seed=0, T=50, first=0, **ocp_kw):
        kw = dict(force_cone=True, land_cstr=True)
        kw.update(ocp_kw)
        self.B, self.T = batch, T
        self.model, self.limits = go2_like_model()
        self.ocpb = go2_full_ocp(self.model, self.limits, T=T, **kw)
        self.nq, self.nv = self.model.nq, self.model.nv
        self.nx, self.ndx, self.nu = self.nq + self.nv, 2 * self.nv, self.nv - 6
        self.xref = x_ref_of(self.model)
        self.feet = fk_numpy(self.model, self.xref[: self.nq])["feet"]
        self.mass = total_mass(self.model)
        rngs = [np.random.default_rng([seed, first + i]) for i in range(batch)]  # global instance index = stream
        self.phase = np.array([int(r.integers(0, 80)) for r in rngs])
        self.noise = [r for r in rngs]
        self.tick = 0
        vr = [np.random.default_rng([seed, first + i, 1]) for i in range(batch)]
        self.vbase = np.zeros((batch, 6))
        for i, r in enumerate(vr):
            self.vbase[i, 0], self.vbase[i, 1], self.vbase[i, 5] = r.uniform(-0.3, 0.3), r.uniform(-0.1, 0.1), r.uniform(-0.3, 0.3)

    def stage_rows(self, d, nst, nland):
        return (self.nv - 6) * (d.kinematics_limits + d.torque_limits) + nst * 5 * d.force_cone + nland * 4 * d.land_cstr

    def gait_slot(self, k):
        mask = self.GAIT[k % 80]
        prev = self.GAIT[(k - 1) % 80]
        return full_slot(self.model, self.ocpb, self.feet, contact_mask=mask, land_mask=mask & ~prev)

    def initial_slots(self):
        # the stages createProblem builds, after the first updateStepTrackerReferences has written the measured foot
        # poses into every stage (reference src/mpc.cpp:278-324) — with the Identity poses of createProblem itself the
        # 1000-weight pose cost fights the rigid contacts and the constructor solve is ill-posed
        s = full_slot(self.model, self.ocpb, self.feet)
        return (Slot * (self.B * self.T))(*([s] * (self.B * self.T)))

    def warmstart(self):
        # MPC::MPC: xs = x0, us = getReferenceControl(0) = 0 for the full-dynamics handler (reference src/mpc.cpp:24-26)
        xs = np.tile(self.xref, (self.B, self.T + 1, 1)).copy()
        us = np.zeros((self.B, self.T, self.nu))
        return xs, us


class TalosCentroidalScenario:
    """Config C2 of SURVEY.md §8d on the synthetic talos-like biped: CentroidalOCP, T = 100, 6-D contacts with
    wrench cones, walk 20 DS / 80 SS(left stance) / 20 DS / 80 SS(right stance)
    (reference examples/talos_centroidal.py:49-117). State = [com; h_lin; h_ang]."""

    GAIT = [0b11] * 20 + [0b01] * 80 + [0b11] * 20 + [0b10] * 80  # left = bit0, right = bit1
    CYCLE = 200
    T_FLY, T_CONTACT = 80, 20       # reference examples/talos_centroidal.py:83-84
    multibody_state = False         # measured state: multibody [q; v] (57); problem state: [com; h] (9)

    def __init__(self, batch, seed=0, T=100, first=0):
        self.B, self.T = batch, T
        self.model, self.limits = talos_like_model()
        self.ocpb = talos_centroidal_ocp(self.model, T=T)
        self.nx = self.ndx = 9
        self.nu = 12
        q = np.array(self.model.q_ref[: self.model.nq])
        self.feet = fk_numpy(self.model, q)["feet"]
        self.mass = total_mass(self.model)
        self.xref = np.concatenate([com_numpy(self.model, q), np.zeros(6)])
        # momentum references: velocity_base * mass for both parts (reference src/centroidal-dynamics.cpp:237-238)
        self.xref_walk = self.xref.copy()
        self.xref_walk[3] = 0.1 * self.mass
        rngs = [np.random.default_rng([seed, first + i]) for i in range(batch)]  # global instance index = stream
        self.phase = np.array([int(r.integers(0, self.CYCLE)) for r in rngs])
        self.noise = rngs
        self.nq, self.nv = self.model.nq, self.model.nv
        self.vbase = np.zeros((batch, 6))
        self.vbase[:, 0] = 0.1      # reference examples/talos_centroidal.py: v_base[0] = 0.1

    def stage_table(self):
        n = len(self.GAIT)
        return (Slot * (n + 1))(*([self.gait_slot(k) for k in range(n)] + [self.gait_slot(0)])), n

    def xmeas_start(self):
        """measured multibody states: the reference configuration, perturbed (the centroidal problem state is derived
        from it by the driver, reference src/centroidal-dynamics.cpp:259-262)"""
        xm = np.tile(x_ref_of(self.model), (self.B, 1))
        for i, r in enumerate(self.noise):
            xm[i, 7:self.nq] += r.normal(size=self.nq - 7) * 0.02
            xm[i, self.nq:] += r.normal(size=self.nv) * 0.05
        return xm

    def next_measurement(self, xm, x1, noise):
        """no multibody plant behind a centroidal plan: the measured configuration stays, velocities are re-drawn"""
        x = np.array(xm, float)
        if noise is not None:
            x[:, self.nq:] = noise
        return x

    def mean_active_rows(self):
        return float(np.mean([17 * bin(m).count("1") for m in self.GAIT]))

    def gait_slot(self, k):
        mask = self.GAIT[k % self.CYCLE]
        nact = bin(mask).count("1")
        return centroidal_slot(self.model, self.ocpb, self.feet, self.xref_walk, contact_mask=mask,
                               force_z=self.mass * 9.81 / max(nact, 1))

    def initial_slots(self):
        # createProblem: all feet in contact, Identity foot poses (reference src/ocp-handler.cpp:111-126)
        ident = [(np.eye(3), np.zeros(3)) for _ in range(2)]
        s = centroidal_slot(self.model, self.ocpb, ident, self.xref)
        return (Slot * (self.B * self.T))(*([s] * (self.B * self.T)))

    def terminals(self):
        t = Terminal()
        t.x_ref[:9] = self.xref
        return (Terminal * self.B)(*([t] * self.B))

    def warmstart(self):
        s0 = centroidal_slot(self.model, self.ocpb, self.feet, self.xref)
        xs = np.tile(self.xref, (self.B, self.T + 1, 1)).copy()
        us = np.tile(np.array(s0.u_ref[: self.nu]), (self.B, self.T, 1)).copy()
        return xs, us

    def x_start(self):
        xs = np.tile(self.xref, (self.B, 1))
        for i, r in enumerate(self.noise):
            xs[i, 0:3] += r.normal(size=3) * 0.01
            xs[i, 3:9] += r.normal(size=6) * 0.5
        return xs

    def tails(self, tick):
        arr = (Slot * self.B)()
        for i in range(self.B):
            arr[i] = self.gait_slot(int(self.phase[i]) + tick)
        return arr

    def disturb(self, x, tick):
        for i, r in enumerate(self.noise):
            x[i, 0:3] += r.normal(size=3) * (0.005 if tick == 0 else 0.001)
            x[i, 3:9] += r.normal(size=6) * (0.5 if tick == 0 else 0.1)
        return x


class TalosFullScenario:
    """Config C4 of SURVEY.md §8d on the synthetic talos-like biped: FullDynamicsOCP with 6-D rigid contacts (CONTACT_6D,
    LOCAL_WORLD_ALIGNED, Baumgarte Kp = [0,0,50,0,0,0], Kd = 100), wrench cones, torque + joint limits, T = 100, walk
    20 DS / 80 SS / 20 DS / 80 SS (reference examples/talos_fulldynamics.py:48-135)."""

    GAIT = [0b11] * 20 + [0b01] * 80 + [0b11] * 20 + [0b10] * 80  # left = bit0, right = bit1
    CYCLE = 200
    T_FLY, T_CONTACT = 80, 20
    multibody_state = True
    KIND = "full"

    def __init__(self, batch, seed=0, T=100, first=0, **ocp_kw):
        self.B, self.T = batch, T
        self.model, self.limits = talos_like_model()
        self.ocpb = self.make_ocp(T, **ocp_kw)
        self.nq, self.nv = self.model.nq, self.model.nv
        self.nx, self.ndx = self.nq + self.nv, 2 * self.nv
        self.nu = self.nv - 6 + (12 if self.KIND == "kino" else 0)
        self.xref = x_ref_of(self.model)
        self.feet = fk_numpy(self.model, self.xref[: self.nq])["feet"]
        self.mass = total_mass(self.model)
        rngs = [np.random.default_rng([seed, first + i]) for i in range(batch)]  # global instance index = stream
        self.phase = np.array([int(r.integers(0, self.CYCLE)) for r in rngs])
        self.noise = rngs
        self.vbase = np.zeros((batch, 6))
        self.vbase[:, 0] = 0.1      # reference examples/talos_fulldynamics.py:170-172

    def make_ocp(self, T, **kw):
        return talos_full_ocp(self.model, self.limits, T=T, **kw)

    def gait_slot(self, k):
        mask = self.GAIT[k % self.CYCLE]
        prev = self.GAIT[(k - 1) % self.CYCLE]
        s = full_slot(self.model, self.ocpb, self.feet, contact_mask=mask, land_mask=mask & ~prev)
        return s

    def stage_table(self):
        n = len(self.GAIT)
        return (Slot * (n + 1))(*([self.gait_slot(k) for k in range(n)] + [self.gait_slot(0)])), n

    def initial_slots(self):
        s = self.gait_slot(0)
        return (Slot * (self.B * self.T))(*([s] * (self.B * self.T)))

    def terminals(self):
        t = Terminal()
        t.x_ref[: self.nx] = self.xref
        c = com_numpy(self.model, self.xref[: self.nq])
        t.com_ref[:] = c
        t.tau = np.sqrt(c[2] / 9.81)
        return (Terminal * self.B)(*([t] * self.B))

    def warmstart(self):
        xs = np.tile(self.xref, (self.B, self.T + 1, 1)).copy()
        us = np.tile(np.array(self.gait_slot(0).u_ref[: self.nu]), (self.B, self.T, 1)).copy()
        return xs, us

    def x_start(self):
        xs = np.tile(self.xref, (self.B, 1))
        for i, r in enumerate(self.noise):
            xs[i, 7:self.nq] += r.normal(size=self.nq - 7) * 0.01
            xs[i, 0:3] += r.normal(size=3) * 0.005
            xs[i, self.nq:] += r.normal(size=self.nv) * 0.02
        return xs

    def xmeas_start(self):
        return self.x_start()

    def next_measurement(self, xm, x1, noise):
        x = np.array(x1, float)
        if noise is not None:
            x[:, self.nq:] += noise
        return x

    def tails(self, tick):
        arr = (Slot * self.B)()
        for i in range(self.B):
            arr[i] = self.gait_slot(int(self.phase[i]) + tick)
        return arr

    def disturb(self, x, tick):
        for i, r in enumerate(self.noise):
            x[i, self.nq:] += r.normal(size=self.nv) * (0.02 if tick == 0 else 0.005)
            if tick == 0:
                x[i, 7:self.nq] += r.normal(size=self.nq - 7) * 0.01
        return x

    def mean_active_rows(self):
        d = self.ocpb.desc
        rows = []
        for k in range(len(self.GAIT)):
            mask, prev = self.GAIT[k], self.GAIT[k - 1]
            nst, nland = bin(mask).count("1"), bin(mask & ~prev).count("1")
            rows.append((self.nv - 6) * (d.kinematics_limits + d.torque_limits) + nst * 17 * d.force_cone + nland * 6 * d.land_cstr)
        return float(np.mean(rows))


class TalosKinoScenario(TalosFullScenario):
    """Talos-like kinodynamics (reference examples/talos_kinodynamics.py:44-130): 6-D contact wrenches as controls,
    6-D LOCAL frame-velocity equalities on the stance feet, optional CentroidalWrenchCone; T = 100, walk 20/80/20/80."""

    KIND = "kino"

    def make_ocp(self, T, **kw):
        return talos_kino_ocp(self.model, self.limits, T=T, **kw)

    def gait_slot(self, k):
        mask = self.GAIT[k % self.CYCLE]
        nact = bin(mask).count("1")
        return standing_slot(self.model, self.ocpb, self.feet, contact_mask=mask, land_mask=0, force_z=self.mass * 9.81 / max(nact, 1))

    def mean_active_rows(self):
        d = self.ocpb.desc
        return float(np.mean([(self.nv - 6) * d.kinematics_limits + bin(m).count("1") * (6 + 17 * d.force_cone) for m in self.GAIT]))
