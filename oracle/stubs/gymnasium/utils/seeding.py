import numpy as np


def np_random(seed=None):
    """gymnasium.utils.seeding.np_random contract: (Generator, seed)."""
    if seed is None:
        seed = int(np.random.SeedSequence().entropy % (2 ** 63))
    return np.random.default_rng(seed), seed
