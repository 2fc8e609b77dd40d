"""Run the UNMODIFIED reference SEIRM path (read from /root/reference) on CPU.

TEST INFRASTRUCTURE.  Used only in the build container (where /root/reference
is mounted) to (1) validate oracle/seirm_port.py and (2) generate the golden
fixtures under tests/golden/ (see make_golden.py).  Nothing here is imported by
the product, the GPU tests, smoke() or bench.py.

Recipe: work on a writable copy of `agent_torch/` (its DataLoader
rewrites yamls/config.yaml in place, core/dataloader.py:58-62), put the
dependency stand-ins of ./standins on sys.path, inject the missing
`LangchainLLM` symbol (new_transmission/action.py:8), and feed the injected noise
through the reference's own hooks:
  * policy.behavior = lambda obs: W[t]          (action.py:52-59, decorators.py:24-36)
  * transition.st_bernoulli = injected sampler  (transition.py:30,223)
No reference source is copied into this repository.
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import pandas as pd
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("AGENT_TORCH_REFERENCE", "/root/reference")

_WORKDIR = None


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "agent_torch"))


def _prepare_import():
    """Copy agent_torch (without the bundled populations) to a temp dir and import it."""
    global _WORKDIR
    if _WORKDIR is not None:
        return _WORKDIR
    work = tempfile.mkdtemp(prefix="at_ref_")
    shutil.copytree(
        os.path.join(REFERENCE_ROOT, "agent_torch"),
        os.path.join(work, "agent_torch"),
        ignore=shutil.ignore_patterns("populations", "*.ipynb", "__pycache__"),
    )
    os.makedirs(os.path.join(work, "agent_torch", "populations"), exist_ok=True)
    open(os.path.join(work, "agent_torch", "populations", "__init__.py"), "w").close()
    sys.path.insert(0, os.path.join(HERE, "standins"))
    sys.path.insert(0, work)
    import agent_torch.core.llm.backend as backend

    if not hasattr(backend, "LangchainLLM"):
        backend.LangchainLLM = type(
            "LangchainLLM", (backend.LLMBackend,), {"prompt": lambda self, p: p}
        )
    _WORKDIR = work
    return work


def write_population(folder, age, stage0, edges):
    """Write the three files the covid config reads (config.yaml:22,31-32)."""
    os.makedirs(os.path.join(folder, "mobility_networks"), exist_ok=True)
    pd.Series(np.asarray(age, dtype=np.int64)).to_pickle(os.path.join(folder, "age.pickle"))
    np.savetxt(os.path.join(folder, "disease_stages.csv"), np.asarray(stage0, dtype=np.float64))
    pd.DataFrame(np.asarray(edges, dtype=np.int64)).to_csv(
        os.path.join(folder, "mobility_networks", "0.csv"), header=False, index=False
    )


def load_bundled_population(name="astoria"):
    """(age, stage0, edges) arrays of a population bundled with the reference."""
    root = os.path.join(REFERENCE_ROOT, "agent_torch", "populations", name)
    age = pd.read_pickle(os.path.join(root, "age.pickle")).to_numpy()
    stage0 = np.loadtxt(os.path.join(root, "disease_stages.csv")).reshape(-1)
    edges = pd.read_csv(os.path.join(root, "mobility_networks", "0.csv"), header=None).to_numpy()
    return age.astype(np.int64), stage0.astype(np.float64), edges.astype(np.int64)


class _InjectedBernoulli(torch.autograd.Function):
    """(u < p) with identity backward: the reference's StraightThroughBernoulli
    (core/distributions/distributions.py:8-24) with the uniform made explicit."""

    @staticmethod
    def forward(ctx, probs, u2):
        return (u2 < probs).to(probs.dtype)

    @staticmethod
    def backward(ctx, g):
        return g, None


def run_reference(age, stage0, edges, U, W, num_steps=21, r2=None, m=None, gI=None, gD=None,
                  num_threads=None, want_grads=True, timing=None, edge_type=None, table_grads=False):
    """Run one episode of the reference on CPU with injected noise.

    U, W: (T, N) float32 arrays.  Returns a dict of numpy arrays: final stage /
    next_stage_time / infected_time, per-step stage history, daily series and
    dL/dR2, dL/dM for L = sum(gI*daily_infected) + sum(gD*daily_deaths).
    """
    import time

    work = _prepare_import()
    if num_threads:
        torch.set_num_threads(num_threads)
    from agent_torch.core.dataloader import LoadPopulation
    from agent_torch.core.executor import Executor
    from agent_torch.models import covid

    n = len(age)
    pop = tempfile.mkdtemp(prefix="pop_", dir=work)
    write_population(pop, age, stage0, edges)

    cfg_path = os.path.join(work, "agent_torch", "models", "covid", "yamls", "config.yaml")
    with open(cfg_path) as f:
        cfg = yaml.safe_load(f)
    cfg["simulation_metadata"]["EXECUTION_MODE"] = "injected"
    cfg["simulation_metadata"]["device"] = "cpu"
    cfg["simulation_metadata"]["num_steps_per_episode"] = int(num_steps)
    with open(cfg_path, "w") as f:
        yaml.dump(cfg, f)

    ex = Executor(model=covid, pop_loader=LoadPopulation(pop))
    runner = ex.runner
    runner.use_gpu = False
    runner.init()
    tr = runner.initializer.transition_function["0"]["new_transmission"]
    pg = runner.initializer.transition_function["1"]["seirm_progression"]
    pol = runner.initializer.policy_function["0"]["citizens"]["make_isolation_decision"]
    if r2 is not None:
        tr.calibrate_R2 = torch.tensor(np.asarray(r2, dtype=np.float32), requires_grad=True)
    if m is not None:
        pg.calibrate_M = torch.tensor(np.asarray(m, dtype=np.float32), requires_grad=True)

    if edge_type is not None:
        # multi-layer contact network: the reference's loader only ever writes edge type 1 (utils.py:143), but its message
        # function gathers I_bar per edge by that type (transition.py:71-73); give it a per-edge type vector
        net = runner.state["network"]["agent_agent"]["infection_network"]
        edge_index, edge_attr = net["adjacency_matrix"]
        et = torch.from_numpy(np.asarray(edge_type)).to(edge_attr.dtype)
        assert et.numel() == edge_attr.shape[1]
        net["adjacency_matrix"] = (edge_index, torch.vstack((et, edge_attr[1])))
    if table_grads:
        # SFSusceptibility / SFInfector as differentiable leaves (config.yaml: learnable-capable environment tables)
        env0 = runner.state["environment"]
        env0["SFSusceptibility"] = env0["SFSusceptibility"].detach().clone().requires_grad_(True)
        env0["SFInfector"] = env0["SFInfector"].detach().clone().requires_grad_(True)
        leaf_sfs, leaf_sfi = env0["SFSusceptibility"], env0["SFInfector"]

    Ut = torch.from_numpy(np.ascontiguousarray(U)).to(torch.float64)
    Wt = torch.from_numpy(np.ascontiguousarray(W)).to(torch.float32)
    pol.behavior = lambda obs: Wt[int(runner.state["current_step"])].reshape(n, 1)
    tr.st_bernoulli = lambda probs: _InjectedBernoulli.apply(
        probs, torch.stack((Ut[int(runner.state["current_step"])], torch.zeros(n, dtype=torch.float64)), dim=1)
    )

    t0 = time.perf_counter()
    runner.step(num_steps)
    t1 = time.perf_counter()
    final = runner.state
    env = runner.state_trajectory[-1][-1]["environment"]
    di = env["daily_infected"]
    dd = env["daily_deaths"]
    out = {
        "stage": final["agents"]["citizens"]["disease_stage"].detach().numpy().reshape(-1).copy(),
        "next_stage_time": final["agents"]["citizens"]["next_stage_time"].detach().numpy().reshape(-1).copy(),
        "infected_time": final["agents"]["citizens"]["infected_time"].detach().numpy().reshape(-1).copy(),
        "daily_infected": di.detach().numpy().reshape(-1).copy(),
        "daily_deaths": dd.detach().numpy().reshape(-1).copy(),
        "stage_history": np.stack(
            [s[-1]["agents"]["citizens"]["disease_stage"].detach().numpy().reshape(-1) for s in runner.state_trajectory[1:]]
        ),
        "lam_gamma_integrals": final["environment"]["lam_gamma_integrals"].numpy().reshape(-1).copy(),
        "mean_interactions": final["environment"]["mean_interactions"].numpy().reshape(-1).copy(),
    }
    if want_grads:
        gI_t = torch.from_numpy(np.asarray(gI, dtype=np.float32)) if gI is not None else torch.ones(num_steps)
        gD_t = torch.from_numpy(np.asarray(gD, dtype=np.float32)) if gD is not None else torch.ones(num_steps)
        loss = (gI_t * di.reshape(-1)).sum() + (gD_t * dd.reshape(-1)).sum()
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        out["loss"] = float(loss.detach())
        out["grad_R2"] = tr.calibrate_R2.grad.numpy().copy()
        out["grad_M"] = pg.calibrate_M.grad.numpy().copy()
        if table_grads:
            out["grad_SFS"] = leaf_sfs.grad.numpy().copy()
            out["grad_SFI"] = leaf_sfi.grad.numpy().copy()
        if timing is not None:
            timing["backward_s"] = t3 - t2
    if timing is not None:
        timing["forward_s"] = t1 - t0
    shutil.rmtree(pop, ignore_errors=True)
    return out
