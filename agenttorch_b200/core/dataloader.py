"""LoadPopulation / DataLoader (reference: agent_torch/core/dataloader.py:12-292).

`LoadPopulation(region)` accepts a folder path, a module with `__path__`, or an object with
`population_folder_path`; every `*.pickle|*.pkl|*.parquet` becomes an attribute (numeric -> float32 tensor).
`DataLoader(model, population)` reads `<model>/yamls/config.yaml`, points `population_dir` / `num_agents`
at the population and resolves the `${...}` interpolations.  Unlike the reference it does NOT rewrite the
YAML in place (dataloader.py:58-62 makes the installed package directory a write target); the edited
config lives in memory, and `save_config(path)` writes it where the caller asks."""
import glob
import os

import pandas as pd
import pandas.api.types as pdt
import torch
import yaml

from .helpers import resolve_config


class LoadPopulation:
    def __init__(self, region):
        if isinstance(region, str):
            self.population_folder_path = region
        elif hasattr(region, "__path__"):
            self.population_folder_path = list(region.__path__)[0]
        elif hasattr(region, "population_folder_path"):
            self.population_folder_path = region.population_folder_path
        else:
            self.population_folder_path = str(region)
        self.population_size = 0
        self.load_population()

    def _store(self, key, obj):
        if isinstance(obj, pd.DataFrame) and obj.shape[1] == 1:
            obj = obj.iloc[:, 0]
        if isinstance(obj, pd.Series) and pdt.is_numeric_dtype(obj.dtype):
            setattr(self, key, torch.from_numpy(obj.to_numpy(dtype="float32", copy=True)))
        elif isinstance(obj, pd.DataFrame) and all(pdt.is_numeric_dtype(obj[c]) for c in obj.columns):
            setattr(self, key, torch.from_numpy(obj.to_numpy(dtype="float32", copy=True)))
        else:
            setattr(self, key, obj)
        return len(obj) if hasattr(obj, "__len__") else 0

    def load_population(self):
        folder = self.population_folder_path
        sizes = {}
        parquet = sorted(glob.glob(os.path.join(folder, "*.parquet")))
        if parquet:
            for pf in parquet:
                df = pd.read_parquet(pf)
                for col in df.columns:
                    sizes[col] = self._store(col, df[col])
        else:
            files = sorted(glob.glob(os.path.join(folder, "*.pickle")) + glob.glob(os.path.join(folder, "*.pkl")))
            for f in files:
                key = os.path.splitext(os.path.basename(f))[0]
                try:
                    sizes[key] = self._store(key, pd.read_pickle(f))
                except Exception as e:   # unreadable side files must not break the population (dataloader.py:184-187)
                    print(f"   {key:20} -> error loading: {e}")
        if "age" in sizes:
            self.population_size = sizes["age"]
        elif sizes:
            self.population_size = next(iter(sizes.values()))


class DataLoader:
    def __init__(self, model, population):
        self.model = model
        self.config_path = os.path.join(list(model.__path__)[0], "yamls", "config.yaml")
        with open(self.config_path, "r", encoding="utf-8") as f:
            self.config = yaml.safe_load(f)
        self.population_size = population.population_size
        self.set_input_data_dir(population.population_folder_path)
        self.set_population_size(population.population_size)

    def set_config_attribute(self, attribute, value):
        if hasattr(value, "__fspath__"):
            value = str(value)
        self.config["simulation_metadata"][attribute] = value

    def set_input_data_dir(self, population_dir):
        self.set_config_attribute("population_dir", population_dir)

    def set_population_size(self, population_size):
        self.population_size = population_size
        self.set_config_attribute("num_agents", population_size)

    def get_config(self):
        return resolve_config(self.config)

    def save_config(self, path):
        with open(path, "w", encoding="utf-8") as f:
            yaml.safe_dump(self.config, f)
