import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from agenttorch_b200.engine import NoiseSpec, SeirmEngine, SeirmSpec, SeirmState
from agenttorch_b200.graph import synthetic_contact_graph, synthetic_population, mean_interactions_from_age
dev = torch.device("cuda:0"); T = 8
for n in (100_000, 300_000):
    age, stage0 = synthetic_population(n, 1234, device=dev)
    graph = synthetic_contact_graph(n, 20.0, 99, device=dev, mean_interactions=mean_interactions_from_age(age))
    for grid in (8, 37, 64, 65, 98, 128, 148, 0):
        for smem in (None, 32768):
            eng = SeirmEngine(graph, SeirmSpec(num_steps=T), block_size=smem, grid=grid)
            st = SeirmState.from_initial(stage0, age, T, dev)
            sv = eng.run_episode(st, torch.full((T,), 14.25, device=dev), torch.tensor([0.12], device=dev), T, NoiseSpec(seed=5, isolation="hard"), want_grad=True)
            torch.cuda.synchronize()
            print(n, grid, smem, [int(x) for x in st.daily_infected.tolist()], sv["death_count"].tolist()[-3:], eng.workspace[4096:4100].tolist())
