"""Shared helpers of the test-suite: oracle-side and product-side builders of the same seeded inputs."""
import numpy as np

from oracle import np_oracle as O


def oracle_m9():
    return O.OracleMaze(O.M9_LAYOUT, O.M9_SIZE_TILE)


def product_m9():
    from navigation_model_b200.simulation.maze import Maze
    return Maze(O.M9_LAYOUT, O.M9_SUBAREAS, size_tile=O.M9_SIZE_TILE)


def a8(size_tile=O.M9_SIZE_TILE):
    return O.A8_UNITS * size_tile


def oracle_transitions(omaze, traj, rew, actions=None):
    mouse = None if actions is None else O.OracleMouse(traj[0], 0.0, actions)
    return O.extract_transitions(omaze, traj, rew, mouse)


def compact(omaze, dense):
    """[X, Y] -> compact [n_states] in the oracle's x-major visitable order."""
    return np.array([dense[x, y] for x, y in omaze.vis_cells])


def grid_params(rng, P, alpha=(0.01, 1.0), beta=(0.1, 30.0), gamma=(0.5, 0.99)):
    return (rng.uniform(*alpha, P), np.exp(rng.uniform(np.log(beta[0]), np.log(beta[1]), P)), rng.uniform(*gamma, P))
