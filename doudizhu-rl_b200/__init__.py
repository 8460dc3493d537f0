"""doudizhu-rl_b200 -- B200-native batched Doudizhu environment (drop-in for the rollout path of
charleschen003/doudizhu-rl: envi.py + its absent native modules `env` and `r`).

The directory name carries a hyphen, so import it as `ddz_b200` (the alias module at the repo root) or with
importlib.import_module("doudizhu-rl_b200").
"""
from . import _native as native  # raises ImportError when libddz_b200.so has not been built
from . import sharding
from .agent import BatchedGreedyPolicy, FusedQScorer
from .trainer import ReplayBuffer, TransitionCollector, td_step
from .game import BatchedGame, BatchedDQN
from .ingest import env_from_payloads, env_from_arrays, payload_arrays, evaluate_moves, mcts
from .search import UctSearch
from .env import (Trajectory, GraphedRollout, GroupedEnv, HostRollout, HostRolloutGroups, GroupStepResults, StepResults, BatchedEnv, BatchedEnvComplicated, BatchedEnvCooperation, BatchedEnvCooperationSimplify,
                  Env, EnvComplicated, EnvCooperation, EnvCooperationSimplify,
                  MoveGenerator, get_moves, kth_moves, mcts_moves, pack_counts, unpack_counts, default_deals, random_deals, adversarial_pairs, ADVERSARIAL_POOL,
                  VARIANT_CHANNELS, DEFAULT_REWARDS)

__all__ = ["native", "sharding", "BatchedGreedyPolicy", "FusedQScorer", "ReplayBuffer", "TransitionCollector", "td_step", "BatchedGame", "BatchedDQN", "env_from_payloads", "env_from_arrays", "payload_arrays", "evaluate_moves", "mcts", "UctSearch", "Trajectory", "GraphedRollout", "GroupedEnv", "HostRollout", "HostRolloutGroups", "GroupStepResults", "StepResults", "BatchedEnv", "BatchedEnvComplicated", "BatchedEnvCooperation", "BatchedEnvCooperationSimplify",
           "Env", "EnvComplicated", "EnvCooperation", "EnvCooperationSimplify", "MoveGenerator", "get_moves", "kth_moves", "mcts_moves", "pack_counts",
           "unpack_counts", "default_deals", "random_deals", "adversarial_pairs", "ADVERSARIAL_POOL", "VARIANT_CHANNELS", "DEFAULT_REWARDS"]
