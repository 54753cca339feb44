"""DPAMSA agent and environment with the reference's module layout (DPAMSA/env.py, DPAMSA/dqn.py),
backed by libgadpamsa's batched Q-network on the GPU."""
