"""CPU restatement ("port") of the reference SEIRM step — the parity oracle.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by agenttorch_b200/.

Pinned against the real reference: tests/golden/*.npz were produced by running
/root/reference itself (oracle/ref_harness/make_golden.py) and
tests/test_oracle_golden.py checks this port against them (stages/timers
bit-exact, series and gradients <= 1e-6 relative).  The reference's own test
suite holds no known-answer vector for this path.

What is restated (paths relative to /root/reference/agent_torch):
  * initial timers / tables       models/covid/substeps/utils.py:14-57,88-109
  * substep '0' transition        models/covid/substeps/new_transmission/transition.py:52-100,131-252
  * message passing (gather/add)  torch_geometric MessagePassing.propagate, aggr="add"
                                  (third-party, >=2.5.3 per pyproject.toml:19; not vendored)
  * straight-through sampler      core/distributions/distributions.py:8-24 (uniform made explicit)
  * substep '1' transition        models/covid/substeps/seirm_progression/transition.py:41-121
The arithmetic keeps the reference's dtypes: per-edge product chain in float32,
division by I_bar and everything downstream (lambda, p, exposure, stage) in
float64, timers float32, daily series float32.
Autograd is torch's own, applied to the same op sequence, so gradients follow
the reference's straight-through convention by construction.
"""
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class SeirmConstants:
    """simulation_metadata constants (models/covid/yamls/config.yaml:1-50)."""
    num_steps: int = 21
    num_weeks: int = 3
    susceptible: int = 0
    exposed: int = 1
    infected: int = 2
    recovered: int = 3
    exposed_to_infected_time: int = 5
    infected_to_recovered_time: int = 5
    sf_susceptibility: tuple = (0.35, 0.69, 1.03, 1.03, 1.03, 1.03, 1.27, 1.52, 1.52)
    sf_infector: tuple = (0.0, 0.33, 0.72, 0.0, 0.0)
    gamma_scale: float = 5.5
    gamma_rate: float = 2.4
    r2_init: float = 4.75
    m_init: float = 0.12

    @property
    def infinity_time(self):  # seirm_progression/transition.py:28-30
        return self.num_steps + 20


def lam_gamma_integrals(scale, rate, t):
    """utils.py:14-27: day-by-day gamma CDF differences, float32.

    The reference receives scale/rate/t as float32 tensors (initializer.py:112 makes
    `value * torch.ones(shape)`), so a and b are computed in float32 first."""
    from scipy.stats import gamma

    scale_t, rate_t = torch.tensor([scale], dtype=torch.float32), torch.tensor([rate], dtype=torch.float32)
    b = rate_t * rate_t / scale_t
    a = scale_t / b
    res = [gamma.cdf(ti, a=a, loc=0, scale=b) - gamma.cdf(ti - 1, a=a, loc=0, scale=b) for ti in range(int(t))]
    return torch.tensor(np.array(res)).float().reshape(-1)


def mean_interactions_from_age(age):
    """utils.py:88-109: 2 for age index 0, 3 for 1..4, 4 above."""
    age = torch.as_tensor(age).reshape(-1)
    out = torch.zeros(age.shape[0], dtype=torch.float32)
    out[age < 1] = 2.0
    out[(age >= 1) & (age <= 4)] = 3.0
    out[age > 4] = 4.0
    return out


def initial_timers(stage0):
    """utils.py:30-57: (infected_time, next_stage_time), float32."""
    stage0 = torch.as_tensor(stage0, dtype=torch.float64).reshape(-1)
    itime = torch.full_like(stage0, 500.0)
    itime[stage0 == 1] = -1.0
    itime[stage0 == 2] = -3.0
    nst = torch.full_like(stage0, 500.0)
    nst[stage0 == 1] = 3.0
    nst[stage0 == 2] = 5.0
    return itime.float(), nst.float()


@dataclass
class SeirmState:
    stage: torch.Tensor            # (N,) float64  (Q3: CSV -> float64)
    next_stage_time: torch.Tensor  # (N,) float32
    infected_time: torch.Tensor    # (N,) float32
    daily_infected: torch.Tensor   # (T,) float32
    daily_deaths: torch.Tensor     # (T,) float32


class SeirmPort:
    """One population + contact graph; `episode()` advances T steps with injected noise."""

    def __init__(self, age, stage0, edges, consts=None, edge_weight=None, edge_type=None,
                 faithful_layout=False):
        self.c = consts or SeirmConstants()
        self.n = int(len(age))
        self.age = torch.as_tensor(np.asarray(age), dtype=torch.float32).reshape(-1)
        self.stage0 = torch.as_tensor(np.asarray(stage0), dtype=torch.float64).reshape(-1)
        edges = torch.as_tensor(np.asarray(edges), dtype=torch.long)
        if edges.ndim == 2 and edges.shape[0] != 2:
            edges = edges.t()
        self.src = edges[0].contiguous()
        self.dst = edges[1].contiguous()
        e = self.src.numel()
        # utils.py:143-145: edge_attr rows = [type == 1, weight == 1.0]
        self.edge_weight = (torch.ones(e, dtype=torch.float32) if edge_weight is None
                            else torch.as_tensor(edge_weight, dtype=torch.float32))
        self.edge_type = (torch.ones(e, dtype=torch.long) if edge_type is None
                          else torch.as_tensor(edge_type, dtype=torch.long))
        self.mean_interactions = mean_interactions_from_age(self.age)
        self.sfs = torch.tensor(self.c.sf_susceptibility, dtype=torch.float32)
        self.sfi = torch.tensor(self.c.sf_infector, dtype=torch.float32)
        self.lam_gamma = lam_gamma_integrals(self.c.gamma_scale, self.c.gamma_rate, self.c.num_steps)
        self.faithful_layout = faithful_layout

    # ------------------------------------------------------------------ state
    def initial_state(self):
        itime, nst = initial_timers(self.stage0)
        t = self.c.num_steps
        return SeirmState(self.stage0.clone(), nst, itime,
                          torch.zeros(t, dtype=torch.float32), torch.zeros(t, dtype=torch.float32))

    def default_params(self):
        r2 = torch.full((self.c.num_weeks, 1), self.c.r2_init, dtype=torch.float32, requires_grad=True)
        m = torch.full((1,), self.c.m_init, dtype=torch.float32, requires_grad=True)
        return r2, m

    # --------------------------------------------------------------- substep 0
    def effective_r(self, r2, t):
        """transition.py:136-143.  (NUM_WEEKS,1)*(NUM_WEEKS,) broadcasts (Q2)."""
        week = torch.eye(self.c.num_weeks)[int(t / 7)]
        return (r2 * week).sum()

    def i_bar_per_edge(self):
        """transition.py:71-73 (Q1): gather of the per-edge mean_interactions[dst] vector at
        index edge_type -> with edge_type == 1 everywhere, one scalar for all edges."""
        per_edge = self.mean_interactions.double().index_select(0, self.dst)
        return torch.gather(per_edge, 0, self.edge_type)

    def lam(self, st, t, r):
        """Per-agent force of infection, (N,) float64 (transition.py:52-100 + propagate)."""
        c = self.c
        stage_d = st.stage.detach()
        infectious = (stage_d > c.susceptible) & (stage_d < c.recovered)
        if self.faithful_layout:
            # the reference materialises an (N,7) float64 feature table and gathers it twice
            x = torch.stack((self.age.double(), stage_d, infectious.double(), st.infected_time.double(),
                             self.mean_interactions.double(), torch.arange(self.n).double(),
                             torch.zeros(self.n, dtype=torch.float64)), dim=1)
            x_i, x_j = x.index_select(0, self.dst), x.index_select(0, self.src)
            age_i, stage_j, inf_j, itime_j = x_i[:, 0], x_j[:, 1], x_j[:, 2].bool(), x_j[:, 3]
        else:
            age_i = self.age.double().index_select(0, self.dst)
            stage_j = stage_d.index_select(0, self.src)
            inf_j = infectious.index_select(0, self.src)
            itime_j = st.infected_time.double().index_select(0, self.src)
        s_i = self.sfs[age_i.long()]
        a_j = self.sfi[stage_j.long()]
        g_j = torch.zeros_like(self.edge_weight)
        g_j[inf_j] = self.lam_gamma[(t - itime_j[inf_j] - 1).long()]
        msg = r * s_i * a_j * self.edge_weight * g_j / self.i_bar_per_edge()      # f32 chain, then /f64
        return torch.zeros(self.n, dtype=msg.dtype).index_add(0, self.dst, msg)

    def transmission(self, st, t, w, u, r2):
        """Substep '0' (transition.py:131-252).  w: (N,) f32 isolation, u: (N,) uniforms."""
        c = self.c
        lam = self.lam(st, t, self.effective_r(r2, t))
        p = 1.0 - torch.exp(-lam)
        hit = (u.double() < p.detach()).double()
        b = p + (hit - p).detach()                     # straight-through: value hit, d/dp = 1
        exposed_today = (st.stage == c.susceptible) * (b * (1.0 - w))
        onehot = torch.eye(c.num_steps)[t]
        daily_infected = st.daily_infected + exposed_today.sum() * onehot
        mask = exposed_today.detach() != 0
        nst = st.next_stage_time.clone()
        nst[mask] = float(t + 1 + c.exposed_to_infected_time)
        itime = st.infected_time.clone()
        itime[mask] = float(t)
        return SeirmState(st.stage + exposed_today, nst, itime, daily_infected, st.daily_deaths), lam

    # --------------------------------------------------------------- substep 1
    def progression(self, st, t, m):
        """Substep '1' (seirm_progression/transition.py:41-121); masks from the incoming state."""
        c = self.c
        stage, nst = st.stage, st.next_stage_time
        due = nst <= t
        is_e, is_i = stage == c.exposed, stage == c.infected
        recovered_or_dead = is_i * due
        deaths = (stage * recovered_or_dead / c.infected).sum() * m
        onehot = torch.eye(c.num_steps)[t]
        daily_deaths = st.daily_deaths + onehot * deaths
        new_stage = stage + (is_e * due + is_i * due)
        new_nst = nst.clone()
        new_nst[is_i & (nst == t)] = float(c.infinity_time)
        new_nst[is_e & (nst == t)] = float(t + c.infected_to_recovered_time)
        return SeirmState(new_stage, new_nst, st.infected_time, st.daily_infected, daily_deaths.reshape(-1))

    # ----------------------------------------------------------------- episode
    def episode(self, W, U, r2=None, m=None, num_steps=None, state=None, keep_history=False, keep_lam=False):
        """W, U: (T,N) arrays/tensors (float32).  Returns (final_state, extras)."""
        T = num_steps or self.c.num_steps
        if r2 is None or m is None:
            r2d, md = self.default_params()
            r2 = r2 if r2 is not None else r2d
            m = m if m is not None else md
        st = state or self.initial_state()
        hist, lams = [], []
        for t in range(T):
            w = torch.as_tensor(W[t], dtype=torch.float32)
            u = torch.as_tensor(U[t], dtype=torch.float32)
            st, lam = self.transmission(st, t, w, u, r2)
            st = self.progression(st, t, m)
            if keep_history:
                hist.append(st.stage.detach().clone())
            if keep_lam:
                lams.append(lam.detach().clone())
        return st, {"stage_history": hist, "lam": lams, "r2": r2, "m": m}


def episode_with_grads(port, W, U, gI, gD, r2=None, m=None, num_steps=None, table_grads=False):
    """Forward + backward of L = sum(gI*daily_infected) + sum(gD*daily_deaths); numpy outputs.
    table_grads: SFSusceptibility / SFInfector as differentiable leaves too (config.yaml:206-235,268-300 declare them
    learnable-capable; transition.py:63-64 indexes them inside the message)."""
    if table_grads:
        port.sfs = port.sfs.detach().clone().requires_grad_(True)
        port.sfi = port.sfi.detach().clone().requires_grad_(True)
    r2d, md = port.default_params()
    r2 = r2d if r2 is None else torch.tensor(np.asarray(r2, dtype=np.float32), requires_grad=True)
    m = md if m is None else torch.tensor(np.asarray(m, dtype=np.float32), requires_grad=True)
    st, extra = port.episode(W, U, r2=r2, m=m, num_steps=num_steps, keep_history=True)
    gI_t = torch.as_tensor(np.asarray(gI, dtype=np.float32))
    gD_t = torch.as_tensor(np.asarray(gD, dtype=np.float32))
    loss = (gI_t * st.daily_infected).sum() + (gD_t * st.daily_deaths).sum()
    loss.backward()
    extra_out = {"grad_SFS": port.sfs.grad.numpy().copy(), "grad_SFI": port.sfi.grad.numpy().copy()} if table_grads else {}
    return {
        **extra_out,
        "stage": st.stage.detach().numpy(),
        "next_stage_time": st.next_stage_time.numpy(),
        "infected_time": st.infected_time.numpy(),
        "daily_infected": st.daily_infected.detach().numpy(),
        "daily_deaths": st.daily_deaths.detach().numpy(),
        "stage_history": torch.stack(extra["stage_history"]).numpy(),
        "loss": float(loss.detach()),
        "grad_R2": r2.grad.numpy().copy(),
        "grad_M": m.grad.numpy().copy(),
    }
