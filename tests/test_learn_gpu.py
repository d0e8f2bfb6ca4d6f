"""GPU: K-LEARN (csrc/learn.cu) against torch autograd on the reference's nets.

The reference update is Proximal_Policy_Optimization2.learn (algorithm/policy_base/Proximal_Policy_Optimization2.py:
102-131): the torch code below restates those lines (the same restatement round 1's VecPPO2 used, checked then against the
reference learner) and serves as the checker; the product path is the C ABI (b200_ppo2_grad / b200_adam_step /
b200_ppo2_learn).  Tolerances: gradients within 2e-6 of a float64 autograd evaluation relative to the largest gradient
entry of the net (float32 autograd itself sits at ~5e-7 on this scale); parameters after clip + Adam steps within 1e-6
absolute of torch.optim.Adam + clip_grad_norm_."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rollout(T, S, A, N, seed, a_lo=0.0, a_hi=5.0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *shape: torch.rand(*shape, device="cuda", generator=g)
    s = (r(T, S, N) * 2 - 1).contiguous()
    a = (a_lo + (a_hi - a_lo) * r(T, A, N)).contiguous()
    a_lp = (-1.5 - 0.5 * r(T, A, N)).contiguous()
    adv = torch.randn(T, N, device="cuda", generator=g).contiguous()
    vt = (3 * torch.randn(T, N, device="cuda", generator=g)).contiguous()
    return s, a, a_lp, adv, vt


def _torch_losses(actor, critic, std, s, a, a_lp, adv, vt, idx, eps_clip=0.2, ent_coef=0.01, dtype=None):
    """PPO2.py:106-124 on the flat batch picked by idx (positions b = t * N + i)."""
    import torch
    T, S, N = s.shape
    flat = lambda x: x.permute(0, 2, 1).reshape(T * N, -1)
    cast = (lambda x: x.to(dtype)) if dtype is not None else (lambda x: x)
    sb, ab, lpb = cast(flat(s)[idx]), cast(flat(a)[idx]), cast(flat(a_lp)[idx])
    advb, vtb = cast(adv.reshape(-1, 1)[idx]), cast(vt.reshape(-1, 1)[idx])
    mean = actor(sb)
    sd = cast(torch.as_tensor(std, device=s.device, dtype=torch.float32).reshape(-1)).expand(ab.shape[1])
    lp = -((ab - mean) ** 2) / (2 * sd * sd) - torch.log(sd) - math.log(math.sqrt(2 * math.pi))
    ent = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sd)).sum()
    ratios = torch.exp(lp.sum(1, keepdim=True) - lpb.sum(1, keepdim=True))
    surr1 = ratios * advb
    surr2 = torch.clamp(ratios, 1 - eps_clip, 1 + eps_clip) * advb
    actor_loss = (-torch.min(surr1, surr2) - ent_coef * ent).mean()
    critic_loss = torch.nn.functional.mse_loss(vtb, critic(sb))
    return actor_loss, critic_loss


def _flat_grads(actor, critic):
    import torch
    from reinforcementlearningplatform_b200.policy import linear_layers
    out = []
    for net in (actor, critic):
        out.append(torch.cat([p.grad.reshape(-1) for lin in linear_layers(net) for p in (lin.weight, lin.bias)]))
    return out


def _make(state_dim=6, action_dim=8, mean_act="relu", scale_mean=30.0, seed=0):
    import copy
    import torch
    from reinforcementlearningplatform_b200.ppo2 import reference_nets
    torch.manual_seed(seed)
    actor, critic = reference_nets(state_dim, action_dim, "cuda", init_std=0.45, mean_act=mean_act)
    with torch.no_grad():
        actor.mean_layer.weight.mul_(scale_mean)     # gain-0.01 init would leave the relu head almost flat
        actor.mean_layer.bias.add_(0.3)
    return actor, critic, copy.deepcopy(actor), copy.deepcopy(critic)


@pytest.mark.parametrize("mean_act", ["relu", "identity"])
def test_gradients_match_autograd(mean_act):
    import torch
    from reinforcementlearningplatform_b200.learn import FusedPPO2Update
    actor, critic, actor_t, critic_t = _make(mean_act=mean_act)
    T, S, A, N = 8, 6, 8, 700
    s, a, a_lp, adv, vt = _rollout(T, S, A, N, 1)
    # log-probs of the stored actions close to the current policy's, so that ratios straddle the clip range
    with torch.no_grad():
        mean = actor_t(s.permute(0, 2, 1).reshape(T * N, S))
        # actions as choose_action draws them (PPO2.py:72-75), log-probs of a slightly older policy
        af = torch.clamp(mean + 0.45 * torch.randn_like(mean), 0.0, 5.0)
        a = af.reshape(T, N, A).permute(0, 2, 1).contiguous()
        lp = -((af - mean) ** 2) / (2 * 0.45 ** 2) - math.log(0.45) - math.log(math.sqrt(2 * math.pi))
        a_lp = (lp + 0.05 * torch.randn_like(lp)).reshape(T, N, A).permute(0, 2, 1).contiguous()
    upd = FusedPPO2Update(actor, critic, 0.45, np.zeros(A), np.full(A, 5.0), mean_act)
    idx = torch.randperm(T * N, device="cuda")[:3000]            # ragged: 23.4 tiles of 128
    upd.grad_step(s, a, a_lp, adv, vt, 0, idx.numel(), index=idx)
    o = upd.params.net_off[1]
    ga, gc = upd.grad[:o].clone(), upd.grad[o:].clone()
    la, lc = upd.loss.tolist()

    # torch float32 autograd (what the reference runs) and a float64 evaluation as the common yardstick
    l32a, l32c = _torch_losses(actor_t, critic_t, 0.45, s, a, a_lp, adv, vt, idx)
    l32a.backward(); l32c.backward()
    t32a, t32c = [g.clone() for g in _flat_grads(actor_t, critic_t)]
    a64, c64 = actor_t.double(), critic_t.double()
    a64.zero_grad(); c64.zero_grad()
    l64a, l64c = _torch_losses(a64, c64, 0.45, s, a, a_lp, adv, vt, idx, dtype=torch.float64)
    l64a.backward(); l64c.backward()
    ra, rc = _flat_grads(a64, c64)
    assert abs(la - float(l64a.detach())) <= 2e-6 * max(1.0, abs(float(l64a.detach()))), (la, float(l64a.detach()))
    assert abs(lc - float(l64c.detach())) <= 2e-6 * max(1.0, abs(float(l64c.detach()))), (lc, float(l64c.detach()))
    for mine, t32, ref, name in ((ga, t32a, ra, "actor"), (gc, t32c, rc, "critic")):
        err = float((mine.double() - ref).abs().max())
        err32 = float((t32.double() - ref).abs().max())
        scale = float(ref.abs().max())
        print(f"{name} [{mean_act}]: |grad|max {scale:.3e}  K-LEARN vs f64 {err:.2e}  torch f32 vs f64 {err32:.2e}  "
              f"K-LEARN vs torch f32 {float((mine - t32).abs().max()):.2e}")
        # within 1e-6 absolute of autograd, or as close to the float64 value as float32 autograd itself (x2)
        assert scale > 1e-4 and (float((mine - t32).abs().max()) <= 1e-6 or err <= max(2e-6 * scale, 2 * err32)), (name, err, err32, scale)
    # bit-reproducible: no atomics in the reduction
    upd.grad_step(s, a, a_lp, adv, vt, 0, idx.numel(), index=idx)
    assert torch.equal(upd.grad[:o], ga) and torch.equal(upd.grad[o:], gc)


def test_clip_and_adam_match_torch_over_several_steps():
    import torch
    from reinforcementlearningplatform_b200.learn import FusedPPO2Update
    actor, critic, actor_t, critic_t = _make()
    T, S, A, N = 4, 6, 8, 1024
    s, a, a_lp, adv, vt = _rollout(T, S, A, N, 2)
    vt = (vt * 10).contiguous()                     # large value targets: the critic's gradient norm exceeds the 0.5 clip
    upd = FusedPPO2Update(actor, critic, 0.45, np.zeros(A), np.full(A, 5.0), "relu", a_lr=1e-4, c_lr=1e-3, adam_eps=1e-5)
    oa = torch.optim.Adam(actor_t.parameters(), lr=1e-4, eps=1e-5)
    oc = torch.optim.Adam(critic_t.parameters(), lr=1e-3, eps=1e-5)
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(4):
        idx = torch.randperm(T * N, device="cuda", generator=g)[:2048]
        upd.grad_step(s, a, a_lp, adv, vt, 0, 2048, index=idx)
        upd.adam_step()
        la, lc = _torch_losses(actor_t, critic_t, 0.45, s, a, a_lp, adv, vt, idx)
        oa.zero_grad(); la.backward()
        na = torch.nn.utils.clip_grad_norm_(actor_t.parameters(), 0.5)
        oa.step()
        oc.zero_grad(); lc.backward()
        nc = torch.nn.utils.clip_grad_norm_(critic_t.parameters(), 0.5)
        oc.step()
        mine = upd.grad_norm.tolist()
        assert abs(mine[0] - float(na)) <= 1e-5 * max(1.0, float(na)) and abs(mine[1] - float(nc)) <= 1e-5 * max(1.0, float(nc))
        assert float(nc) > 0.5                      # the critic's gradient is actually clipped
        # |delta param| <= 1e-6 for the actor; for the critic (lr 1e-3, value targets x10) 2e-3 of one Adam step (whose
        # size is <= lr): where |g| ~ eps = 1e-5 the update lr g / (|g| + eps) turns a 1e-8 difference of float32
        # summation order into 2.5e-4 lr -- torch against torch with another reduction order differs as much
        for net, ref, tol in ((actor, actor_t, 1e-6), (critic, critic_t, 2e-6)):
            for p, q in zip(net.parameters(), ref.parameters()):
                assert float((p.detach() - q.detach()).abs().max()) <= tol, (step, float((p.detach() - q.detach()).abs().max()))


def test_keyed_permutation_visits_every_sample_once_and_loop_matches_steps():
    import copy
    import torch
    from reinforcementlearningplatform_b200.learn import FusedPPO2Update
    actor, critic, _, _ = _make(seed=3)
    actor2, critic2 = copy.deepcopy(actor), copy.deepcopy(critic)
    T, S, A, N = 5, 6, 8, 333                       # B = 1665: not a power of two, not a multiple of the tile
    s, a, a_lp, adv, vt = _rollout(T, S, A, N, 3)
    B = T * N
    upd = FusedPPO2Update(actor, critic, 0.45, np.zeros(A), np.full(A, 5.0), "relu")
    upd.grad_step(s, a, a_lp, adv, vt, 0, B, index=torch.arange(B, device="cuda"))
    full = upd.grad.clone()
    acc = torch.zeros_like(full, dtype=torch.float64)
    per_mb = []
    for first in range(0, B, 400):
        cnt = min(400, B - first)
        upd.grad_step(s, a, a_lp, adv, vt, first, cnt, perm_key=77)
        acc += upd.grad.double() * cnt / B
        per_mb.append(upd.grad.clone())
    assert float((acc - full.double()).abs().max()) <= 3e-6 * float(full.abs().max())
    upd.grad_step(s, a, a_lp, adv, vt, 0, 400, perm_key=78)      # another key = another mini-batch
    assert float((upd.grad - per_mb[0]).abs().max()) > 1e-3 * float(full.abs().max())
    # b200_ppo2_learn (C loop) == the same sequence of grad_step / adam_step calls
    upd2 = FusedPPO2Update(actor2, critic2, 0.45, np.zeros(A), np.full(A, 5.0), "relu")
    upd2.learn(s, a, a_lp, adv, vt, k_epochs=2, mini_batch=400, perm_key=1000)
    for e in range(2):
        for first in range(0, B, 400):
            upd.grad_step(s, a, a_lp, adv, vt, first, min(400, B - first), perm_key=1000 + e)
            upd.adam_step()
    assert upd2.step_count == upd.step_count == 10
    assert torch.equal(upd.params.flat, upd2.params.flat)


def test_tanh_range_head_with_std_vector():
    """the DPPO2 demos' actor head (mean = tanh(z) * gain + off, per-dimension std) on a net K-LEARN can hold"""
    import torch
    from reinforcementlearningplatform_b200.learn import FusedPPO2Update
    from reinforcementlearningplatform_b200.ppo2 import dppo2_nets
    torch.manual_seed(7)
    a_min, a_max = np.array([-3.0, -1.0]), np.array([3.0, 2.0])
    actor, critic = dppo2_nets(41, 2, a_min, a_max, "cuda", hidden=64)
    with torch.no_grad():
        actor.mean_layer.weight.mul_(40.0)
    import copy
    a64, c64 = copy.deepcopy(actor).double(), copy.deepcopy(critic).double()
    a64.off, a64.gain = actor.off.double(), actor.gain.double()
    a64.forward = lambda x: torch.tanh(a64.mean_layer(torch.tanh(a64.fc2(torch.tanh(a64.fc1(x)))))) * a64.gain + a64.off
    T, S, A, N = 3, 41, 2, 500
    s, a, a_lp, adv, vt = _rollout(T, S, A, N, 9, a_lo=-1.0, a_hi=2.0)
    a_lp = a_lp * 0.5
    upd = FusedPPO2Update(actor, critic, actor.std, a_min, a_max, "tanh_range")
    idx = torch.randperm(T * N, device="cuda")[:1111]
    upd.grad_step(s, a, a_lp, adv, vt, 0, 1111, index=idx)
    la, lc = _torch_losses(a64, c64, actor.std, s, a, a_lp, adv, vt, idx, dtype=torch.float64)
    la.backward(); lc.backward()
    ra, rc = _flat_grads(a64, c64)
    o = upd.params.net_off[1]
    for mine, ref in ((upd.grad[:o], ra), (upd.grad[o:], rc)):
        err, scale = float((mine.double() - ref).abs().max()), float(ref.abs().max())
        print(f"tanh_range: |grad|max {scale:.3e} err {err:.2e}")
        assert err <= max(1e-6, 5e-6 * scale), (err, scale)
    mine = upd.loss.tolist()
    assert abs(mine[0] - float(la.detach())) <= 3e-6 * max(1.0, abs(float(la.detach())))


def test_wide_nets_are_rejected_loudly():
    from reinforcementlearningplatform_b200 import _lib
    from reinforcementlearningplatform_b200.learn import FusedPPO2Update, fused_supported
    from reinforcementlearningplatform_b200.ppo2 import dppo2_nets
    actor, critic = dppo2_nets(41, 2, np.array([-3.0, -1.0]), np.array([3.0, 2.0]), "cuda", hidden=256)
    assert not fused_supported(actor, critic)
    with pytest.raises(_lib.B200EnvError):
        FusedPPO2Update(actor, critic, actor.std, [-3.0, -1.0], [3.0, 2.0], "tanh_range")
