"""automoe_b200 — B200-native (sm_100a) implementation of AutoMoE's batched forward hot path.

The directory is named after the reference repository (`self-driving-model_b200`); import it
through the `automoe_b200` alias package at the repo root:

    from automoe_b200.models.automoe import create_automoe_model
    from automoe_b200.training.hungarian_matcher import HungarianMatcher
"""
__version__ = "0.1.0"
