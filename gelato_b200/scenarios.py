"""Dispersed launch scenarios and their partition over GPUs.

A scenario is the nominal problem with perturbed stage masses, thrust levels
and wind profile -- an independent NLP.  The reference can only run such a
study one settings file after another (/root/reference/run_batch.sh:75-79); here
a batch of scenarios is ONE kernel launch (scenario = blockIdx.x mod n_scen,
role-major), and batches are spread over GPUs without any exchange in the loop.
"""
import copy

import numpy as np


def disperse(inputs, n_scen, seed=20260117, sigma_mass=0.01, sigma_thrust=0.01, sigma_wind=0.2, sigma_dir_deg=10.0):
    """n_scen perturbed copies of a `problem.read_inputs` dictionary (C4):
    dry / propellant masses x (1 + sigma_mass N), thrust x (1 + sigma_thrust N),
    wind speed x (1 + sigma_wind N) clipped at 0, wind direction + sigma_dir_deg N.
    Scenario 0 is the nominal problem."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for k in range(n_scen):
        inp = copy.deepcopy(inputs)
        if k > 0:
            for stage in inp["settings"]["RocketStage"].values():
                stage["mass_dry"] = stage["mass_dry"] * (1.0 + sigma_mass * rng.standard_normal())
                stage["mass_propellant"] = stage["mass_propellant"] * (1.0 + sigma_mass * rng.standard_normal())
            f_thrust = {}
            for e in inp["events"]:
                f_thrust.setdefault(e["rocketStage"], 1.0 + sigma_thrust * rng.standard_normal())
                e["thrust"] = e["thrust"] * f_thrust[e["rocketStage"]]
            w = np.array(inp["wind_table"], dtype=np.float64)
            speed = np.hypot(w[:, 1], w[:, 2]) * max(0.0, 1.0 + sigma_wind * rng.standard_normal())
            direction = np.arctan2(-w[:, 2], -w[:, 1]) + np.radians(sigma_dir_deg * rng.standard_normal())
            w[:, 1] = speed * -np.cos(direction)
            w[:, 2] = speed * -np.sin(direction)
            inp["wind_table"] = w
        out.append(inp)
    return out


def partition(n_scen, world_size, rank):
    """Contiguous block of scenario indices owned by `rank` (sizes differ by at most 1)."""
    base, extra = divmod(n_scen, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))
