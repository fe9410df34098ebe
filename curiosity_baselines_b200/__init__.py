"""B200-native implementation of the intrinsic-reward hot path of curiosity_baselines.

Public surface (mirrors the reference's names):
  curiosity.RND / ICM / Disagreement / NDIGO      compute_bonus / compute_loss
  algos.utils.discount_return / generalized_advantage_estimation / valid_from_done
  algos.pg_base.process_returns, agents.hooks.curiosity_step / curiosity_loss
All arithmetic runs in libcuriosity_b200.so (hand-written CUDA for sm_100a); there is no CPU
fallback.
"""
__version__ = "0.1.0"
