"""-m gpu: the training path -- `model.train(); loss.backward()` on the CUDA kernels against the
parameter gradients of the unmodified reference under autograd (tests/golden/t*.npz, made by make_golden_train.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from util import SEED, rel_l2, trained_like_, weights_digest, white_noise

from aero_b200 import Aero, aero_kwargs

pytestmark = pytest.mark.gpu
# Tolerances.  The golden gradients come from the reference promoted to fp64.  Every FTB block ends in BatchNorm + ReLU over a
# few hundred to a few thousand samples per channel, and the bias / scale gradients behind it are sums of O(1) terms that
# largely cancel: ONE activation that sits within fp32 rounding of the ReLU kink and falls on the other side moves such a
# gradient (and everything upstream of that block) by ~1e-3 .. 1e-2 of its norm.  The reference's own fp32 run deviates from
# its fp64 run by 1.3e-3 .. 2.9e-3 on exactly these parameters (measured when the fixtures were made), so the per-parameter
# bar cannot be 1e-3 for all of them.  Measured on the isolated kernel with the real data (round 2): its sums agree with the
# exact fp64 sum of the same terms to 7e-8; the whole difference to the reference is which side of the kink single activations
# fall on (e.g. one element of 3616 behind BatchNorm2d(5) = 1.6e-2 of that bias gradient).  Op-level and layer-level tests
# (tests/test_gpu_train_ops.py) pin every kernel at 1e-5 .. 5e-5 where no kink is involved.  Here:
#   strict cases (no activation near a kink: t3): all gradients together <= 1e-3 AND every parameter <= 1e-3;
#   other cases: all gradients together <= 5e-3, at least a third of the parameters <= 1e-3, none above 5e-2 (a wrong kernel
#   gives O(1) errors on everything upstream of it).
STRICT = {"t3_11-44_stereo"}
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "t*.npz")))
GRAD_TOL = 1e-3


def cotangent(shape, seed):
    return white_noise(shape, seed=seed + 77)


def grad_report(model, g):
    """Per-parameter relative error on the committed samples.  Parameters whose reference gradient is tiny against the
    largest one are judged on the absolute error at that scale (a relative error of a rounding-level number says nothing)."""
    rows = []
    num = den = 0.0
    gmax = max(float(g[k]) for k in g.files if k.startswith("g_rms/"))
    for name, p in model.named_parameters():
        ref = torch.from_numpy(g["g_val/" + name]).double()
        idx = torch.from_numpy(g["g_idx/" + name].astype(np.int64))
        assert p.grad is not None, f"no gradient for {name}"
        got = p.grad.detach().reshape(-1).cpu().double()[idx]
        rms = float(g["g_rms/" + name])
        if name.endswith(("freq_attn_block.conv1.0.bias", "freq_attn_block.conv1d.0.bias", "freq_attn_block.conv2.0.bias",
                          "time_attn.key.bias")):
            # a bias in front of a BatchNorm, or the key bias of the attention (softmax is shift invariant): the true gradient is
            # zero and the reference holds rounding noise only
            assert float(got.abs().max()) < 1e-3 * gmax, (name, float(got.abs().max()), gmax)
            continue
        denom = max(float(ref.norm()), 1e-4 * gmax * ref.numel() ** 0.5)
        rows.append((float((got - ref).norm()) / denom, name, rms))
        scale = p.numel() / ref.numel()                     # the 256 samples stand for the whole tensor
        num += scale * float((got - ref).pow(2).sum())
        den += scale * float(ref.pow(2).sum())
    return sorted(rows, reverse=True), (num / den) ** 0.5


@pytest.mark.parametrize("case", CASES)
def test_parameter_gradients_match_reference_autograd(golden_dir, case):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    torch.manual_seed(SEED)
    m = Aero(**aero_kwargs(str(g["exp"])))
    m.load_state_dict(trained_like_(m.state_dict()))
    assert weights_digest(m.state_dict()) == pytest.approx(float(g["digest"]), rel=1e-12)
    m = m.cuda().train()
    mix = white_noise((int(g["B"]), m.in_channels, int(g["L"]))).cuda()
    out = m(mix)
    assert tuple(out.shape) == tuple(int(v) for v in g["out_shape"]) and out.requires_grad
    flat = out.detach().reshape(-1).cpu()
    e_out = rel_l2(flat[torch.from_numpy(g["out_idx"].astype(np.int64))], g["out_val"])
    R = cotangent(tuple(out.shape), SEED).cuda()
    loss = (out * R).sum() / out.numel()
    loss.backward()
    torch.cuda.synchronize()
    rows, total = grad_report(m, g)
    ok = sum(1 for r in rows if r[0] < GRAD_TOL)
    print(f"{case}: train-mode output rel_l2 {e_out:.3e}, loss {float(loss):.6e} (ref {float(g['loss']):.6e}); all gradients together "
          f"{total:.3e}; {ok}/{len(rows)} parameters within {GRAD_TOL:g}; worst:")
    for err, name, rms in rows[:8]:
        print(f"   {err:.3e}  {name}  (ref rms {rms:.3e})")
    assert e_out < 2e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-4 * max(abs(float(g["loss"])), 1e-6) + 1e-9
    if case in STRICT:
        assert total < GRAD_TOL and rows[0][0] < GRAD_TOL, (total, rows[:5])
    else:
        assert total < 5e-3, total
        assert ok >= len(rows) / 3, (ok, len(rows))
        assert rows[0][0] < 5e-2, rows[:5]
    # BatchNorm running buffers were updated as nn.BatchNorm does in train mode
    for k in g.files:
        if k.startswith("buf/"):
            got = dict(m.named_buffers())[k[4:]].cpu()
            assert rel_l2(got, g[k]) < 1e-5, k


def test_train_step_with_fused_adam_reduces_the_loss():
    """A few optimisation steps through the public API: forward (train), backward, aero_b200.optim.FusedAdam."""
    from aero_b200.optim import FusedAdam
    torch.manual_seed(SEED)
    m = Aero(**aero_kwargs("aero_4-16_512_256"))
    m.load_state_dict(trained_like_(m.state_dict()))
    m = m.cuda().train()
    ref = torch.optim.Adam([p.detach().clone().requires_grad_(True) for p in m.parameters()], lr=3e-4, betas=(0.9, 0.999))
    opt = FusedAdam(m.parameters(), lr=3e-4, betas=(0.9, 0.999))
    mix = white_noise((2, 1, 4000)).cuda()
    target = white_noise((2, 1, 16000), seed=5).cuda() * 0.1
    losses = []
    for step in range(4):
        opt.zero_grad()
        out = m(mix)
        loss = (out - target).abs().mean()
        loss.backward()
        if step == 0:       # one step of torch.optim.Adam on the same gradients gives the same parameters
            for q, p in zip(ref.param_groups[0]["params"], m.parameters()):
                q.grad = p.grad.detach().clone()
            ref.step()
        opt.step()
        if step == 0:
            for q, p in zip(ref.param_groups[0]["params"], m.parameters()):
                assert rel_l2(p.detach().cpu(), q.detach().cpu()) < 1e-6
        losses.append(float(loss))
    print("losses:", losses)
    assert losses[-1] < losses[0]


def test_melgan_discriminator_matches_reference(golden_dir):
    """aero_b200.discriminator.Discriminator (grouped-conv / weight-norm kernels) against the
    reference's features and gradients (tests/golden/disc_melgan.npz, fp64 reference)."""
    from util import disc_recipe_state
    from aero_b200.discriminator import Discriminator
    g = np.load(os.path.join(golden_dir, "disc_melgan.npz"))
    d = Discriminator(3, 16, 4, 4)
    assert len(d.state_dict()) == 63 and sum(p.numel() for p in d.parameters()) == 16924086
    d.load_state_dict(disc_recipe_state(d.state_dict()))
    assert weights_digest(d.state_dict()) == pytest.approx(float(g["digest"]), rel=1e-12)
    d = d.cuda()
    B, L = int(g["B"]), int(g["L"])
    x = white_noise((B, 1, L), seed=SEED + 3).cuda().requires_grad_(True)
    feats = d(x)
    loss = 0.0
    worst_f = 0.0
    for i, scale in enumerate(feats):
        assert len(scale) == 7
        for j, f in enumerate(scale):
            assert tuple(f.shape) == tuple(int(v) for v in g[f"f_shape/{i}/{j}"])
            flat = f.detach().reshape(-1).cpu()
            worst_f = max(worst_f, rel_l2(flat[torch.from_numpy(g[f"f_idx/{i}/{j}"].astype(np.int64))], g[f"f_val/{i}/{j}"]))
            loss = loss + (f * white_noise(tuple(f.shape), seed=SEED + 1000 + 10 * i + j).cuda()).mean()
    loss.backward()
    torch.cuda.synchronize()
    e_dx = rel_l2(x.grad.cpu(), g["dx"])
    rows = []
    gmax = max(float(g[k]) for k in g.files if k.startswith("g_rms/"))
    for name, p in d.named_parameters():
        ref = torch.from_numpy(g["g_val/" + name]).double()
        got = p.grad.reshape(-1).cpu().double()[torch.from_numpy(g["g_idx/" + name].astype(np.int64))]
        rows.append((float((got - ref).norm()) / max(float(ref.norm()), 1e-4 * gmax * ref.numel() ** 0.5), name))
    rows.sort(reverse=True)
    print(f"discriminator: features {worst_f:.2e}, loss {float(loss):.6e} (ref {float(g['loss']):.6e}), d input {e_dx:.2e}; worst parameter gradients:",
          [(f"{a:.1e}", b) for a, b in rows[:4]])
    assert worst_f < 2e-5 and e_dx < 1e-4 and rows[0][0] < 1e-3


def test_gan_training_step_runs_and_updates_both_networks():
    """The reference's adversarial step (solver.py:292-342,475-520) through aero_b200.trainer.GanTrainer: finite losses, both
    parameter sets move, the generator's gradients equal what plain autograd (`loss.backward()`) gives for the same losses."""
    from aero_b200.discriminator import Discriminator
    from aero_b200.losses import MultiResolutionSTFTLoss
    from aero_b200.trainer import GanTrainer
    torch.manual_seed(SEED)
    m = Aero(**aero_kwargs("aero_4-16_512_256"))
    m.load_state_dict(trained_like_(m.state_dict()))
    m = m.cuda().train()
    torch.manual_seed(SEED + 1)
    d = Discriminator(3, 16, 4, 4).cuda()
    lr_b, hr_b = white_noise((2, 1, 4000)).cuda(), white_noise((2, 1, 16000), seed=5).cuda() * 0.1
    stft = MultiResolutionSTFTLoss()
    # plain autograd route first (same weights, BatchNorm buffers restored afterwards)
    bufs = {k: v.clone() for k, v in m.named_buffers()}
    tr = GanTrainer(m, d)
    for p in list(m.parameters()) + list(d.parameters()):
        p.grad = None
    pr = m(lr_b)
    losses = tr.generator_losses(pr, hr_b, stft)
    sum(losses.values()).backward()
    want = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    with torch.no_grad():
        for k, v in m.named_buffers():
            v.copy_(bufs[k])
    tr = GanTrainer(m, d)                                # re-binds .grad to the flat buffers
    g0 = [p.detach().clone() for p in m.parameters()]
    d0 = [p.detach().clone() for p in d.parameters()]
    out = tr.step(lr_b, hr_b, stft)
    torch.cuda.synchronize()
    assert all(torch.isfinite(v) for v in out.values()), out
    # the generator gradients the trainer applied == the autograd ones (the flat buffer still holds them)
    num = sum(float((tr.views[n] - want[n]).pow(2).sum()) for n in want)
    den = sum(float(want[n].pow(2).sum()) for n in want)
    print("GAN step losses:", {k: round(float(v), 5) for k, v in out.items()}, " trainer-vs-autograd gradient rel_l2:", (num / den) ** 0.5)
    assert (num / den) ** 0.5 < 1e-4
    assert any(not torch.equal(a, b) for a, b in zip(g0, m.parameters())) and any(not torch.equal(a, b) for a, b in zip(d0, d.parameters()))
