"""Device-resident micro-benchmarks of the scan / normalisation kernels (K-GAE, K-RET, K-NORM) for ncu captures and
for the `also` block of bench.py.  python tools/microbench.py gae|gae_flags|mc_returns|norm [reps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def run(kind, reps=10, T=2048, N=131072, dev="cuda"):
    import reinforcementlearningplatform_b200 as rlp
    from reinforcementlearningplatform_b200 import gae as G
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    mk = lambda: torch.randn((T, N), generator=g, device=dev, dtype=torch.float32)
    if kind in ("gae", "gae_flags", "mc_returns"):
        r, vs, vsn = mk(), mk(), mk()
        dn = torch.rand((T, N), generator=g, device=dev) < 0.002
        done_u8 = dn.to(torch.uint8)
        flag = (dn & (torch.rand((T, N), generator=g, device=dev) < 0.5)).to(torch.int32) * 3
        done, succ = dn.float(), (dn & (flag != 3)).float()
        if kind == "gae":
            fn, nbytes = (lambda: G.gae(r, vs, vsn, done, succ, 0.99, 0.95)), 28.0
        elif kind == "gae_flags":
            fn, nbytes = (lambda: G.gae_flags(r, vs, vsn, done_u8, flag, 3, 0.99, 0.95)), 25.0
        else:
            fn, nbytes = (lambda: G.mc_returns(r, done_u8, 0.99)), 9.0
        units = T * N
    elif kind == "norm":
        dim, n = 6, 1 << 22  # 6 x 4 M float32 = 100 MB per pass: larger than what L2 keeps between the two kernels
        x = torch.randn((dim, n), generator=g, device=dev, dtype=torch.float32)
        y = torch.empty_like(x)
        nz = rlp.Normalization(dim, device=dev, sync=False)
        fn, nbytes, units = (lambda: nz.normalize_soa(x, out=y)), 12.0, dim * n
    elif kind == "policy_wide":
        # the 41-256-256-{2,1} nets of the DPPO2 UGV-FOA demo (train.py:26-107), 262,144 instances (config #5 per GPU)
        S, A, n = 41, 2, 1 << 18
        torch.manual_seed(0)
        mk_l = lambda i, o: torch.nn.Linear(i, o).to(dev)
        actor = [mk_l(S, 256), mk_l(256, 256), mk_l(256, A)]
        critic = [mk_l(S, 256), mk_l(256, 256), mk_l(256, 1)]
        pol = rlp.GaussianPolicy(actor, critic, [-3.0, -6.28], [3.0, 6.28], [1.0, 2.09], device=dev,
                                 actor_out_act="tanh_range")
        obs = torch.randn((S, n), generator=g, device=dev, dtype=torch.float32)
        outs = pol(obs)
        fn = lambda: pol(obs, action=outs["action"], log_prob=outs["log_prob"], value=outs["value"])
        flops = 2.0 * (2 * (S * 256 + 256 * 256) + 256 * A + 256)
        nbytes, units = (S + 2 * A + 1) * 4.0, n
    elif kind in ("policy", "policy_fp32", "policy_tf32x3"):
        # the reference's PPOActor_Gaussian / PPOCritic shapes for UavFntsmcParamPos (state 6, 8 gains), 1 M instances
        S, A, n = 6, 8, 1 << 20
        torch.manual_seed(0)
        mk_l = lambda i, o: torch.nn.Linear(i, o).to(dev)
        actor = [mk_l(S, 64), mk_l(64, 64), mk_l(64, 32), mk_l(32, A)]
        critic = [mk_l(S, 64), mk_l(64, 32), mk_l(32, 1)]
        pol = rlp.GaussianPolicy(actor, critic, [0.0] * A, [5.0] * A, 0.45, device=dev,
                                 precision={"policy": "umma", "policy_fp32": "fp32", "policy_tf32x3": "tf32x3"}[kind])
        obs = torch.randn((S, n), generator=g, device=dev, dtype=torch.float32)
        outs = pol(obs)
        fn = lambda: pol(obs, action=outs["action"], log_prob=outs["log_prob"], value=outs["value"])
        flops = 2.0 * (S * 64 + 64 * 64 + 64 * 32 + 32 * A + S * 64 + 64 * 32 + 32)
        nbytes, units = (S + 2 * A + 1) * 4.0, n
    else:
        raise SystemExit(f"unknown kind {kind}")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) * 1e-3 / reps
    res = {"workload": kind, "value": units / per, "unit": "elements/s", "ms_per_step": per * 1e3,
           "algorithmic_bytes_per_element": nbytes, "achieved_gbs": nbytes * units / per / 1e9}
    if kind.startswith("policy"):
        res.update({"unit": "instances/s", "flops_per_instance": flops, "achieved_tflops_fp32": flops * units / per / 1e12})
    return res


if __name__ == "__main__":
    print(json.dumps(run(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 10)))
