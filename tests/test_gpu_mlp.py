"""GPU test of the fused tcgen05 policy forward against the float32 torch module.
bf16 operands / fp32 accumulation: probabilities agree to 2e-2 absolute (typically 3e-3),
rows sum to 1, values to 5e-2 relative-ish."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Model(torch.nn.Module):
    """Minimal stand-in with the attribute names FusedPolicyForward reads."""

    def __init__(self, F, H, A0, A1):
        super().__init__()
        self.flattened_obs_size, self.fc_dims, self.output_dims = F, [H, H], [A0, A1]
        self.is_deterministic = False
        self.fc = torch.nn.ModuleDict({
            "0": torch.nn.Sequential(torch.nn.Linear(F, H), torch.nn.ReLU()),
            "1": torch.nn.Sequential(torch.nn.Linear(H, H), torch.nn.ReLU())})
        self.policy_head = torch.nn.ModuleList([torch.nn.Linear(H, A0), torch.nn.Linear(H, A1)])
        self.vf_head = torch.nn.Linear(H, 1)

    def forward(self, x):
        x = self.fc["1"](self.fc["0"](x))
        return [torch.softmax(h(x), -1) for h in self.policy_head], self.vf_head(x)[..., 0]


@pytest.mark.parametrize("F,H,A0,A1,rows", [(71, 256, 21, 21, 128), (71, 256, 21, 21, 2000 * 100),
                                            (71, 256, 21, 21, 10000), (36, 64, 21, 21, 333),
                                            (15, 32, 4, 4, 77), (200, 128, 30, 10, 5000)])
def test_fused_mlp_matches_torch(F, H, A0, A1, rows):
    from warp_drive_b200.training.models.fused_forward import FusedPolicyForward

    torch.manual_seed(F * 1000 + H)
    model = _Model(F, H, A0, A1).cuda()
    for p in model.parameters():          # larger weights than default init: sharper softmax
        p.data.mul_(2.0)
    assert FusedPolicyForward.supported(model)
    fwd = FusedPolicyForward(model)
    obs = torch.randn(rows, F, device="cuda")
    p0 = torch.full((rows, A0), -1.0, device="cuda")
    p1 = torch.full((rows, A1), -1.0, device="cuda")
    v = torch.full((rows,), -1.0, device="cuda")
    fwd(obs, p0, p1, v)
    torch.cuda.synchronize()
    with torch.no_grad():
        (q0, q1), qv = model(obs)
    assert torch.isfinite(p0).all() and torch.isfinite(p1).all()
    assert (p0.sum(-1) - 1).abs().max() < 1e-4 and (p1.sum(-1) - 1).abs().max() < 1e-4
    e0, e1 = (p0 - q0).abs().max().item(), (p1 - q1).abs().max().item()
    ev = ((v - qv).abs() / (1 + qv.abs())).max().item()
    assert e0 < 2e-2 and e1 < 2e-2 and ev < 5e-2, (e0, e1, ev)
    # refresh() picks up new parameters
    with torch.no_grad():
        model.policy_head[0].bias.add_(3.0 * torch.arange(A0, device="cuda") / A0)
    fwd.refresh()
    fwd(obs, p0, p1, None)
    with torch.no_grad():
        (q0, _), _ = model(obs)
    assert (p0 - q0).abs().max().item() < 2e-2


@pytest.mark.parametrize("F,H,A0,A1,rows", [(71, 256, 21, 21, 2000 * 100), (71, 256, 21, 21, 10000),
                                            (71, 256, 21, 21, 77), (36, 64, 21, 21, 333),
                                            (200, 128, 30, 10, 5000)])
def test_fused_mlp_from_tiles_equals_fp32_path(F, H, A0, A1, rows):
    """wdb_mlp_pack_obs + wdb_mlp_policy_forward_tiles: the same bf16 A operand reaches the
    tensor cores as on the fp32-obs path, so the outputs are IDENTICAL bit for bit."""
    from warp_drive_b200.training.models.fused_forward import FusedPolicyForward

    torch.manual_seed(7 * F + H)
    model = _Model(F, H, A0, A1).cuda()
    fwd = FusedPolicyForward(model)
    obs = torch.randn(rows, F, device="cuda")
    outs = []
    for use_tiles in (False, True):
        p0 = torch.full((rows, A0), -1.0, device="cuda")
        p1 = torch.full((rows, A1), -1.0, device="cuda")
        v = torch.full((rows,), -1.0, device="cuda")
        if use_tiles:
            tiles = torch.full((fwd.tiles_bytes(rows),), 0xFF, dtype=torch.uint8, device="cuda")
            fwd.pack_obs(obs, tiles)          # must overwrite every byte it later reads
            fwd.forward_tiles(tiles, rows, p0, p1, v)
        else:
            fwd(obs, p0, p1, v)
        outs.append((p0, p1, v))
    torch.cuda.synchronize()
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_bf16_forward_gap_on_config2_shapes():
    """How far the bf16-operand rollout forward is from the float32 forward the update
    recomputes: on config-2 shapes (71 -> 256 -> 256 -> 21 / 21) with default
    and with 'trained-scale' (3x) weights and observations drawn like the env's (|x| <= 1,
    mostly small): max |p_bf16 - p_fp32| <= 5e-3, mean KL(p_fp32 || p_bf16) <= 1e-5, and the
    PPO ratio exp(logp_fp32 - logp_bf16) of sampled actions stays within 1 +- 2e-2 (the PPO
    clip is 0.1)."""
    import json
    import os

    from warp_drive_b200.training.models.fused_forward import FusedPolicyForward

    rows, F, H, A = 2000 * 100, 71, 256, 21
    out = {}
    for scale in (1.0, 3.0):
        torch.manual_seed(11)
        model = _Model(F, H, A, A).cuda()
        for p in model.policy_head.parameters():
            p.data.mul_(scale)
        fwd = FusedPolicyForward(model)
        g = torch.Generator(device="cuda").manual_seed(5)
        obs = (torch.rand(rows, F, device="cuda", generator=g) * 2 - 1) * \
            torch.rand(rows, 1, device="cuda", generator=g)
        p0 = torch.empty((rows, A), device="cuda")
        p1 = torch.empty((rows, A), device="cuda")
        fwd(obs, p0, p1, None)
        with torch.no_grad():
            (q0, q1), _ = model(obs)
        err = max((p0 - q0).abs().max().item(), (p1 - q1).abs().max().item())
        kl = 0.5 * ((q0 * (q0.clamp_min(1e-30) / p0.clamp_min(1e-30)).log()).sum(-1).mean().item()
                    + (q1 * (q1.clamp_min(1e-30) / p1.clamp_min(1e-30)).log()).sum(-1).mean().item())
        a0 = torch.multinomial(p0, 1, generator=g)
        ratio = (q0.gather(1, a0) / p0.gather(1, a0))
        out[f"head_scale_{scale:g}"] = {"max_abs_prob_err": err, "mean_kl": kl,
                                        "ppo_ratio_min": ratio.min().item(),
                                        "ppo_ratio_max": ratio.max().item()}
        assert err <= 5e-3, (scale, err)
        assert kl <= 1e-5, (scale, kl)
        assert 0.98 <= ratio.min().item() and ratio.max().item() <= 1.02, (scale, out)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "mlp_bf16_gap.json"), "w") as fp:
        json.dump(out, fp, indent=1)


def test_single_head_policy_on_the_tensor_cores():
    """A1 = 0: gridworld / classic-control policies with one action head."""
    from warp_drive_b200.training.models.fused_forward import FusedPolicyForward

    class _One(_Model):
        def __init__(self, F, H, A0):
            super().__init__(F, H, A0, 1)
            self.output_dims = [A0]
            self.policy_head = torch.nn.ModuleList([torch.nn.Linear(H, A0)])

    for F, H, A0, rows in [(21, 32, 5, 250), (4, 32, 2, 10000), (6, 64, 3, 777)]:
        torch.manual_seed(F + H)
        model = _One(F, H, A0).cuda()
        from warp_drive_b200.utils.spaces import Box

        model.observation_space = Box(-1.0, 1.0, shape=(F,))
        assert FusedPolicyForward.supported(model)
        fwd = FusedPolicyForward(model)
        obs = torch.randn(rows, F, device="cuda")
        p0 = torch.full((rows, A0), -1.0, device="cuda")
        v = torch.full((rows,), -1.0, device="cuda")
        fwd(obs, p0, None, v)
        with torch.no_grad():
            (q0,), qv = model(obs)
        assert (p0.sum(-1) - 1).abs().max() < 1e-4
        assert (p0 - q0).abs().max().item() < 2e-2
        assert ((v - qv).abs() / (1 + qv.abs())).max().item() < 5e-2


@pytest.mark.parametrize("shapes", [((71, 256, 21, 21, 2000 * 100), (71, 256, 21, 21, 2000 * 5)),
                                    ((71, 256, 21, 21, 300), (36, 64, 21, 21, 5000)),
                                    ((15, 32, 4, 4, 77), (200, 128, 30, 10, 130))])
@pytest.mark.parametrize("stable", [False, True])
def test_pair_forward_equals_two_single_forwards(shapes, stable):
    """wdb_mlp_policy_forward_pair runs two policies in one grid (CTAs split between them):
    every CTA executes the same code on the same tiles as in the single launches, so the
    outputs are IDENTICAL bit for bit; also back to back (programmatic dependent launches
    overlapping each other's tails) and with a forced split."""
    from warp_drive_b200.training.models.fused_forward import FusedPolicyForward, forward_pair

    fw, obs, single = [], [], []
    for i, (F, H, A0, A1, rows) in enumerate(shapes):
        torch.manual_seed(31 * F + H + i)
        m = _Model(F, H, A0, A1).cuda()
        f = FusedPolicyForward(m)
        o = torch.randn(rows, F, device="cuda")
        p0 = torch.full((rows, A0), -1.0, device="cuda")
        p1 = torch.full((rows, A1), -1.0, device="cuda")
        f(o, p0, p1, None)
        fw.append(f); obs.append(o); single.append((p0, p1))
    torch.cuda.synchronize()
    for ctas_b in (0, 1, 40):
        outs = [[torch.full_like(t, -2.0) for t in pr] for pr in single]
        for _ in range(3):      # back-to-back launches: each may start under the previous one
            forward_pair(fw[0], fw[1], obs[0], obs[1], outs[0], outs[1], ctas_b=ctas_b,
                         weights_stable=stable)
        torch.cuda.synchronize()
        for w in range(2):
            for k in range(2):
                assert torch.equal(outs[w][k], single[w][k]), (ctas_b, w, k)
