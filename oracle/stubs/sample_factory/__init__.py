"""Stand-in for `sample_factory` (not installed here) — TEST INFRASTRUCTURE ONLY: the two interface classes
swarm_rl/env_wrappers/reward_shaping.py:5 mixes into its wrapper."""
