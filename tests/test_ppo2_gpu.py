"""GPU: the device-resident PPO2 pipeline (policy kernel -> step kernel -> rollout buffer -> reward normaliser -> critic
kernel -> GAE kernel -> torch update) runs end to end and learns: on the PPO2 CartPoleAngleOnly variant the share of
episodes that survive to the time-out rises and the episode-termination rate falls within a few iterations."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_vec_ppo2_learns_cartpole_angleonly():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import train_ppo2_vec as T
    log = T.main(["--env", "cartpole_angleonly", "--envs", "4096", "--steps", "64", "--iters", "25", "--epochs", "6"])
    assert all(np.isfinite(r["actor_loss"]) and np.isfinite(r["critic_loss"]) and np.isfinite(r["mean_reward"]) for r in log)
    first = np.mean([r["done_rate"] for r in log[:3]])
    last = np.mean([r["done_rate"] for r in log[-3:]])
    assert last < 0.7 * first, (first, last, log[-1])           # the pole falls less often
    assert np.mean([r["mean_reward"] for r in log[-3:]]) > np.mean([r["mean_reward"] for r in log[:3]])


def test_rollout_rows_written_by_policy_and_step_kernels():
    """one collect(): every row of the buffer was produced on the device; log-probs match a torch recomputation"""
    import math
    import torch
    import reinforcementlearningplatform_b200 as rlp
    from reinforcementlearningplatform_b200.ppo2 import VecPPO2, reference_nets
    torch.manual_seed(0)
    env = rlp.UavPosCtrlRL(n_envs=2048, random_trajectory=True, io_dtype=torch.float32, auto_reset=True, seed=4)
    actor, critic = reference_nets(6, 8, "cuda", init_std=0.45)
    with torch.no_grad():
        actor.mean_layer.weight.mul_(50.0)
    agent = VecPPO2(env, actor, critic, {"buffer_size": 16, "K_epochs": 1}, std=0.45, reward_norm=False, seed=1)
    env.reset(True)
    agent.collect()
    b = agent.buffer
    s, a, a_lp, r, s_, done, succ = b.to_tensor()
    with torch.no_grad():
        mean = actor(s)
    lp = -((a - mean) ** 2) / (2 * 0.45 ** 2) - math.log(0.45) - math.log(math.sqrt(2 * math.pi))
    assert torch.allclose(a_lp, lp, atol=3e-5)
    assert float(a.min()) >= 0.0 and float(a.max()) <= 5.0 and float((a == 0).float().mean()) > 0.01
    assert torch.isfinite(r).all()
    # s of step t+1 is s_ of step t wherever no episode ended (the loop's `current_state = next_state.copy()`,
    # PPO2-4-UavFntsmcParamPos/train.py:290): the buffer holds what the policy acted on
    keep = (b.done[:-1] == 0).unsqueeze(1).expand(-1, 6, -1)
    assert torch.equal(b.s[1:][keep], b.s_[:-1][keep])
    out = agent.learn()
    assert np.isfinite(out["actor_loss"]) and np.isfinite(out["critic_loss"])
