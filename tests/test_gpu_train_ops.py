"""-m gpu: every training kernel (include/aero_b200.h, "Training") against torch autograd of the same op in fp64 on the CPU.
The ops are driven through the tape of aero_b200.train_engine.TrainEngine, i.e. exactly as the training step uses them."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import SEED, rel_l2

from aero_b200 import Aero, aero_kwargs, cabi
from aero_b200.train_engine import TrainEngine, _Conv, _C1x1

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(SEED + seed))


@pytest.fixture(scope="module")
def eng():
    torch.manual_seed(0)
    m = Aero(**aero_kwargs("aero_4-16_512_256")).cuda().train()
    e = TrainEngine(m)
    e.params, e.buffers = {}, {}
    return e


def run_backward(e, out, dy):
    e.acc(out, dy.contiguous().cuda().float().reshape(-1))
    for fn in reversed(e.tape):
        fn()
    torch.cuda.synchronize()


def cl(x):   # NCHW (B,C,F,T) -> channels-last flat [B,F,T,C]
    return x.permute(0, 2, 3, 1).contiguous()


CONVS = [
    ("1x1", dict(), 12, 20, 5, 5),
    ("enc_k8_s4", dict(kf=8, stride_f=4, pad_f=2), 8, 16, 16, 4),
    ("enc_k8_s2", dict(kf=8, stride_f=2, pad_f=3), 12, 8, 8, 4),
    ("dec_3x3", dict(kf=3, kt=3, pad_f=1, pad_t=1), 8, 12, 4, 4),
    ("dconv_k3_dil2", dict(kt=3, dil_t=2, pad_t=2), 8, 4, 3, 3),
    ("k9", dict(kt=9, pad_t=4), 16, 8, 1, 1),
    ("enc0_real", dict(kf=8, stride_f=4, pad_f=2), 48, 48, 256, 64),
    ("k9_wide", dict(kt=9, pad_t=4), 2048, 48, 1, 1),
    ("dec_wide_n", dict(kf=3, kt=3, pad_f=1, pad_t=1), 48, 200, 4, 4),      # N >= 128: the 8 x 8 micro-tile wgrad, ragged N tile
]


@pytest.mark.parametrize("name,kw,K,N,Fi,Fo", CONVS, ids=[c[0] for c in CONVS])
def test_conv_forward_wgrad_dgrad(eng, name, kw, K, N, Fi, Fo):
    e = eng
    e._reset()
    B, T = 2, (95 if "real" in name or "wide" in name else 37)
    cv = _Conv(**kw)
    x = rnd(B, K, Fi, T, seed=1).double().requires_grad_(True)
    w = (rnd(N, K, cv.kf, cv.kt, seed=2) / math.sqrt(K * cv.kf * cv.kt)).double().requires_grad_(True)
    b = rnd(N, seed=3).double().requires_grad_(True)
    ref = F.conv2d(x, w, b, stride=(cv.stride_f, 1), padding=(cv.pad_f, cv.pad_t), dilation=(1, cv.dil_t))
    assert ref.shape[2] == Fo
    dy = rnd(*ref.shape, seed=4).double()
    ref.backward(dy)
    e.params = {"w": w.detach().float().cuda(), "b": b.detach().float().cuda()}
    xg = cl(x.detach().float()).cuda()
    out = e.conv(xg, None, K, 0, "w", "b", cv, B, Fi, Fo, T, N)
    run_backward(e, out, cl(dy))
    assert rel_l2(out.view(B, Fo, T, N).cpu(), cl(ref.detach())) < 1e-5
    assert rel_l2(e.pg["w"].cpu(), w.grad) < 1e-5 and rel_l2(e.pg["b"].cpu(), b.grad) < 1e-5
    assert rel_l2(e.grad(xg).view(B, Fi, T, K).cpu(), cl(x.grad)) < 1e-5


THIN = [   # few weights, many pixels: the one-thread-per-weight wgrad kernel (csrc/train.cu, wgrad_small_kernel)
    ("k2_to_48_k8_s4", "conv", dict(kf=8, stride_f=4, pad_f=2), 2, 48, 64, 16, 300),
    ("c48_to_5_1x1", "conv", dict(), 48, 5, 16, 16, 200),
    ("c12_k1x3", "conv", dict(kt=3, pad_t=1), 48, 12, 8, 8, 300),
    ("convt_96_to_2_k8_s4", "convt", dict(kf=8, stride_f=4, f_off=2), 96, 2, 16, 64, 150),
    ("disc_1_to_16_k15", "conv", dict(kt=15, pad_t=7), 1, 16, 1, 1, 6000),
]


@pytest.mark.parametrize("name,kind,kw,K,N,Fi,Fo,T", THIN, ids=[c[0] for c in THIN])
def test_thin_layers_wgrad(eng, name, kind, kw, K, N, Fi, Fo, T):
    e = eng
    e._reset()
    B = 2
    x = rnd(B, K, Fi, T, seed=1).double().requires_grad_(True)
    if kind == "conv":
        cv = _Conv(**kw)
        w = (rnd(N, K, cv.kf, cv.kt, seed=2) / math.sqrt(K * cv.kf * cv.kt)).double().requires_grad_(True)
        ref = F.conv2d(x, w, None, stride=(cv.stride_f, 1), padding=(cv.pad_f, cv.pad_t), dilation=(1, cv.dil_t))
    else:
        cv = _Conv("convt", **kw)
        w = (rnd(K, N, cv.kf, 1, seed=2) / math.sqrt(K * 2)).double().requires_grad_(True)
        ref = F.conv_transpose2d(x, w, None, stride=(cv.stride_f, 1))[:, :, cv.f_off:cv.f_off + Fo]
    assert ref.shape[2] == Fo
    dy = rnd(*ref.shape, seed=4).double()
    ref.backward(dy)
    e.params = {"w": w.detach().float().cuda()}
    xg = cl(x.detach().float()).cuda()
    out = e.conv(xg, None, K, 0, "w", None, cv, B, Fi, Fo, T, N)
    run_backward(e, out, cl(dy))
    assert rel_l2(out.view(B, Fo, T, N).cpu(), cl(ref.detach())) < 1e-5
    assert rel_l2(e.pg["w"].cpu(), w.grad) < 1e-5
    assert rel_l2(e.grad(xg).view(B, Fi, T, K).cpu(), cl(x.grad)) < 1e-5


def test_conv_two_sources_and_transposed_with_crop(eng):
    e = eng
    e._reset()
    B, T, C1, C2, N, Fq = 2, 21, 8, 12, 16, 5
    x1, x2 = rnd(B, C1, Fq, T, seed=1).double().requires_grad_(True), rnd(B, C2, Fq, T, seed=2).double().requires_grad_(True)
    w = (rnd(N, C1 + C2, 3, 3, seed=3) / 10).double().requires_grad_(True)
    b = rnd(N, seed=4).double().requires_grad_(True)
    y = F.conv2d(torch.cat([x1, x2], 1), w, b, padding=1)
    wt = (rnd(N, 6, 8, 1, seed=5) / 10).double().requires_grad_(True)        # ConvTranspose2d weight [K, N_out, kf, 1]
    bt = rnd(6, seed=6).double().requires_grad_(True)
    z = F.conv_transpose2d(y, wt, bt, stride=(2, 1))[:, :, 3:-3]
    dz = rnd(*z.shape, seed=7).double()
    z.backward(dz)
    e.params = {"w": w.detach().float().cuda(), "b": b.detach().float().cuda(), "wt": wt.detach().float().cuda(), "bt": bt.detach().float().cuda()}
    a1, a2 = cl(x1.detach().float()).cuda(), cl(x2.detach().float()).cuda()
    yo = e.conv(a1, a2, C1, C2, "w", "b", _Conv(kf=3, kt=3, pad_f=1, pad_t=1), B, Fq, Fq, T, N)
    f_keep = (Fq - 1) * 2 + 8 - 6
    zo = e.conv(yo, None, N, 0, "wt", "bt", _Conv("convt", kf=8, stride_f=2, f_off=3), B, Fq, f_keep, T, 6)
    run_backward(e, zo, cl(dz))
    assert rel_l2(zo.view(B, f_keep, T, 6).cpu(), cl(z.detach())) < 1e-5
    for k, r in (("w", w), ("b", b), ("wt", wt), ("bt", bt)):
        assert rel_l2(e.pg[k].cpu(), r.grad) < 1e-5, k
    assert rel_l2(e.grad(a1).view(B, Fq, T, C1).cpu(), cl(x1.grad)) < 1e-5 and rel_l2(e.grad(a2).view(B, Fq, T, C2).cpu(), cl(x2.grad)) < 1e-5


def test_decoder0_weight_slice(eng):
    e = eng
    e._reset()
    B, T, Cc, Fq = 1, 9, 8, 3
    skip = rnd(B, Cc, Fq, T, seed=1).double().requires_grad_(True)
    w = (rnd(16, 2 * Cc, 3, 3, seed=2) / 10).double().requires_grad_(True)
    y = F.conv2d(torch.cat([torch.zeros_like(skip), skip], 1), w, None, padding=1)
    dy = rnd(*y.shape, seed=3).double()
    y.backward(dy)
    e.params = {"w": w.detach().float().cuda()}
    s_ = cl(skip.detach().float()).cuda()
    yo = e.conv(None, s_, 0, Cc, "w", None, _Conv(kf=3, kt=3, pad_f=1, pad_t=1), B, Fq, Fq, T, 16, wslice=slice(Cc, 2 * Cc))
    run_backward(e, yo, cl(dy))
    assert rel_l2(e.pg["w"][:, Cc:].cpu(), w.grad[:, Cc:]) < 1e-5 and float(e.pg["w"][:, :Cc].abs().max()) == 0.0
    assert rel_l2(e.grad(s_).view(B, Fq, T, Cc).cpu(), cl(skip.grad)) < 1e-5


NORMS = [
    ("gn4_gelu", cabi.NA_GELU, 1, 4, 16, False),
    ("gn4_glu", cabi.NA_GLU, 1, 4, 32, False),
    ("gn1_snake_rows", cabi.NA_SNAKE, 2, 1, 12, False),
    ("gn1_glu_scale_res_rows", cabi.NA_GLU_SCALE_RES, 2, 1, 24, False),
    ("gelu_only", cabi.NA_GELU, 1, 1, 8, True),
    ("glu_only", cabi.NA_GLU, 1, 1, 16, True),
]


@pytest.mark.parametrize("name,op,scope,groups,C_,nonorm", NORMS, ids=[c[0] for c in NORMS])
def test_groupnorm_activation_forward_backward(eng, name, op, scope, groups, C_, nonorm):
    e = eng
    e._reset()
    B, Fq, T = 2, 5, 23
    x = (rnd(B, C_, Fq, T, seed=1) * 1.5 + 0.3).double().requires_grad_(True)
    gamma, beta = (1 + 0.3 * rnd(C_, seed=2)).double().requires_grad_(True), (0.2 * rnd(C_, seed=3)).double().requires_grad_(True)
    a = (0.5 + rnd(Fq, seed=4).abs()).double().requires_grad_(True)
    sc = rnd(C_ // 2, seed=5).double().requires_grad_(True)
    res = rnd(B, C_ // 2, Fq, T, seed=6).double().requires_grad_(True)
    if nonorm:
        gx = x
    elif scope == 1:
        gx = F.group_norm(x, groups, gamma, beta, 1e-5)
    else:       # GroupNorm(1, C) per (b, f) row: DConv's [B*F, C, T] view
        rows = x.permute(0, 2, 1, 3).reshape(B * Fq, C_, T)
        gx = F.group_norm(rows, 1, gamma, beta, 1e-5).view(B, Fq, C_, T).permute(0, 2, 1, 3)
    if op == cabi.NA_GELU:
        y = F.gelu(gx)
    elif op == cabi.NA_GLU:
        y = F.glu(gx, 1)
    elif op == cabi.NA_SNAKE:
        av = a.view(1, 1, Fq, 1)
        y = gx + torch.sin(av * gx) ** 2 / av
    else:
        y = res + sc.view(1, -1, 1, 1) * F.glu(gx, 1)
    dy = rnd(*y.shape, seed=7).double()
    y.backward(dy)
    e.params = {"g": gamma.detach().float().cuda(), "b": beta.detach().float().cuda(), "a": a.detach().float().cuda(),
                "s": sc.detach().float().cuda()}
    xg = cl(x.detach().float()).cuda()
    xd = cl(x.detach())
    if scope == 1:
        v = xd.view(B, Fq * T, groups, C_ // groups)
        st = torch.stack([v.sum((1, 3)), (v * v).sum((1, 3))], -1).reshape(-1, 2)
    else:
        v = xd.view(B * Fq, T * C_)
        st = torch.stack([v.sum(1), (v * v).sum(1)], -1)
    rg = cl(res.detach().float()).cuda() if op == cabi.NA_GLU_SCALE_RES else None
    yo = e.norm_act(xg, op, B=B, F_in=Fq, T=T, C_=C_, scope=scope, groups=groups, gname="g", bname="b", stats=st.cuda(), no_norm=nonorm,
                    snake="a" if op == cabi.NA_SNAKE else None, scale="s" if op == cabi.NA_GLU_SCALE_RES else None, residual=rg)
    run_backward(e, yo, cl(dy))
    Co = y.shape[1]
    assert rel_l2(yo.view(B, Fq, T, Co).cpu(), cl(y.detach())) < 1e-5
    assert rel_l2(e.grad(xg).view(B, Fq, T, C_).cpu(), cl(x.grad)) < 2e-5
    if not nonorm:
        assert rel_l2(e.pg["g"].cpu(), gamma.grad) < 2e-5 and rel_l2(e.pg["b"].cpu(), beta.grad) < 2e-5
    if op == cabi.NA_SNAKE:
        assert rel_l2(e.pg["a"].cpu(), a.grad) < 2e-5
    if op == cabi.NA_GLU_SCALE_RES:
        assert rel_l2(e.pg["s"].cpu(), sc.grad) < 2e-5
        assert rel_l2(e.grad(rg).view(B, Fq, T, Co).cpu(), cl(res.grad)) < 1e-6


def test_groupnorm_on_uncropped_rows_then_crop(eng):
    """decoder norm2 (aero.py:206-209): statistics over the uncropped transposed-conv output, then rows [pad:-pad]."""
    e = eng
    e._reset()
    B, Ff, T, C_, pad = 2, 9, 11, 16, 2
    x = rnd(B, C_, Ff, T, seed=1).double().requires_grad_(True)
    gamma, beta = (1 + 0.3 * rnd(C_, seed=2)).double().requires_grad_(True), (0.2 * rnd(C_, seed=3)).double().requires_grad_(True)
    y = F.gelu(F.group_norm(x, 4, gamma, beta, 1e-5)[:, :, pad:-pad])
    dy = rnd(*y.shape, seed=4).double()
    y.backward(dy)
    e.params = {"g": gamma.detach().float().cuda(), "b": beta.detach().float().cuda()}
    xg = cl(x.detach().float()).cuda()
    v = cl(x.detach()).view(B, Ff * T, 4, C_ // 4)
    st = torch.stack([v.sum((1, 3)), (v * v).sum((1, 3))], -1).reshape(-1, 2)
    yo = e.norm_act(xg, cabi.NA_GELU, B=B, F_in=Ff, F_out=Ff - 2 * pad, f_off=pad, T=T, C_=C_, scope=1, groups=4, gname="g", bname="b",
                    stats=st.cuda())
    run_backward(e, yo, cl(dy))
    assert rel_l2(yo.view(B, Ff - 2 * pad, T, C_).cpu(), cl(y.detach())) < 1e-5
    assert rel_l2(e.grad(xg).view(B, Ff, T, C_).cpu(), cl(x.grad)) < 2e-5
    assert rel_l2(e.pg["g"].cpu(), gamma.grad) < 2e-5 and rel_l2(e.pg["b"].cpu(), beta.grad) < 2e-5


@pytest.mark.parametrize("B,Fq,T,C_", [(2, 6, 19, 8), (2, 1, 95, 192), (3, 8, 40, 16)])
def test_batchnorm_relu_train_mode(eng, B, Fq, T, C_):
    e = eng
    e._reset()
    x = (rnd(B, C_, Fq, T, seed=1) * 1.3 + 0.4).double().requires_grad_(True)
    bn = torch.nn.BatchNorm2d(C_).double().train()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.3 * rnd(C_, seed=2))
        bn.bias.copy_(0.2 * rnd(C_, seed=3))
    y = F.relu(bn(x))
    dy = rnd(*y.shape, seed=4).double()
    y.backward(dy)
    e.params = {"g": bn.weight.detach().float().cuda(), "b": bn.bias.detach().float().cuda()}
    e.buffers = {"bn.running_mean": torch.zeros(C_).cuda(), "bn.running_var": torch.ones(C_).cuda(),
                 "bn.num_batches_tracked": torch.zeros((), dtype=torch.long).cuda()}
    xg = cl(x.detach().float()).cuda()
    st = e._batch_stats(xg, C_, "bn", C_)
    yo = e.norm_act(xg, cabi.NA_RELU, B=B, F_in=Fq, T=T, C_=C_, scope=3, gname="g", bname="b", stats=st)
    run_backward(e, yo, cl(dy))
    assert rel_l2(yo.view(B, Fq, T, C_).cpu(), cl(y.detach())) < 1e-5
    assert rel_l2(e.grad(xg).view(B, Fq, T, C_).cpu(), cl(x.grad)) < 2e-5
    assert rel_l2(e.pg["g"].cpu(), bn.weight.grad) < 2e-5 and rel_l2(e.pg["b"].cpu(), bn.bias.grad) < 2e-5
    assert rel_l2(e.buffers["bn.running_mean"].cpu(), bn.running_mean) < 1e-5 and rel_l2(e.buffers["bn.running_var"].cpu(), bn.running_var) < 1e-5


@pytest.mark.parametrize("Fq,T,Cc", [(8, 95, 192), (16, 31, 24), (64, 17, 8), (256, 95, 48), (64, 95, 48)])
def test_ftb_block_train_mode(eng, Fq, T, Cc):
    """The whole FTB block (modules.py:281-325) in train mode against the same torch modules under autograd."""
    e = eng
    e._reset()
    B = 2
    torch.manual_seed(SEED)
    conv1 = torch.nn.Sequential(torch.nn.Conv2d(Cc, 5, 1), torch.nn.BatchNorm2d(5), torch.nn.ReLU())
    conv1d = torch.nn.Sequential(torch.nn.Conv1d(Fq * 5, Cc, 9, padding=4), torch.nn.BatchNorm1d(Cc), torch.nn.ReLU())
    fc = torch.nn.Linear(Fq, Fq, bias=False)
    conv2 = torch.nn.Sequential(torch.nn.Conv2d(2 * Cc, Cc, 1), torch.nn.BatchNorm2d(Cc), torch.nn.ReLU())
    mods = torch.nn.ModuleDict({"conv1": conv1, "conv1d": conv1d, "freq_fc": fc, "conv2": conv2}).double().train()
    with torch.no_grad():
        for k, p in mods.named_parameters():
            if p.dim() == 1 and ".1." in k:
                p.copy_((1 + 0.3 * rnd(*p.shape, seed=len(k))) if k.endswith("weight") else 0.2 * rnd(*p.shape, seed=len(k)))
    x = rnd(B, Cc, Fq, T, seed=1).double().requires_grad_(True)
    xa = conv1(x)
    xa = conv1d(xa.reshape(B, 5 * Fq, T)).view(B, Cc, 1, T)
    xa = xa * x
    xt = fc(xa.transpose(2, 3)).transpose(2, 3)
    y = conv2(torch.cat([xt, x], 1))
    dy = rnd(*y.shape, seed=2).double()
    y.backward(dy)
    pre = "encoder.9.freq_attn_block."
    e.params = {pre + k: p.detach().float().cuda() for k, p in mods.named_parameters()}
    e.buffers = {pre + k: b.detach().clone().to(torch.float32 if b.dtype.is_floating_point else b.dtype).cuda() * 0 + (1 if "var" in k else 0)
                 for k, b in mods.named_buffers()}
    xg = cl(x.detach().float()).cuda()
    yo = e.ftb(xg, "encoder.9", B, Fq, T, Cc)
    run_backward(e, yo, cl(dy))
    assert rel_l2(yo.view(B, Fq, T, Cc).cpu(), cl(y.detach())) < 2e-5
    assert rel_l2(e.grad(xg).view(B, Fq, T, Cc).cpu(), cl(x.grad)) < 5e-5
    worst = []
    for k, p in mods.named_parameters():
        if k in ("conv1.0.bias", "conv1d.0.bias", "conv2.0.bias"):          # followed by BatchNorm: the true gradient is zero
            assert float(e.pg[pre + k].abs().max()) < 1e-4 * float(dy.abs().max())
            continue
        worst.append((rel_l2(e.pg[pre + k].cpu(), p.grad), k))
    print(sorted(worst, reverse=True)[:4])
    assert max(worst)[0] < 5e-5, sorted(worst, reverse=True)[:4]


@pytest.mark.parametrize("H,T,rows", [(12, 40, 3), (48, 230, 2), (96, 77, 1)])
def test_blstm_block(eng, H, T, rows):
    """BLSTM (modules.py:28-65): framing + 2-layer BiLSTM + Linear + skip, against nn.LSTM under autograd."""
    e = eng
    e._reset()
    torch.manual_seed(SEED + 1)
    lstm = torch.nn.LSTM(bidirectional=True, num_layers=2, hidden_size=H, input_size=H).double()
    lin = torch.nn.Linear(2 * H, H).double()
    x = rnd(rows, H, T, seed=1).double().requires_grad_(True)
    # reference framing (utils.py:22-35, modules.py:36-60)
    if T > 200:
        width, stride = 200, 100
        n_frames = math.ceil(T / stride)
        tgt = (n_frames - 1) * stride + width
        xp = F.pad(x, (0, tgt - T))
        fr = xp.unfold(-1, width, stride)                     # [rows, H, nF, width]
        nF = fr.shape[2]
        xin = fr.permute(0, 2, 1, 3).reshape(-1, H, width)
    else:
        xin, nF = x, 1
    yl = lstm(xin.permute(2, 0, 1))[0]
    yl = lin(yl).permute(1, 2, 0)
    if T > 200:
        frames = yl.reshape(rows, -1, H, width)
        limit = stride // 2
        outp = []
        for k in range(nF):
            if k == 0:
                outp.append(frames[:, k, :, :-limit])
            elif k == nF - 1:
                outp.append(frames[:, k, :, limit:])
            else:
                outp.append(frames[:, k, :, limit:-limit])
        yl = torch.cat(outp, -1)[..., :T]
    y = yl + x
    dy = rnd(*y.shape, seed=2).double()
    y.backward(dy)
    q = "enc.dconv.layers.0"
    e.params = {q + ".lstm.lstm." + k: p.detach().float().cuda() for k, p in lstm.named_parameters()}
    e.params.update({q + ".lstm.linear." + k: p.detach().float().cuda() for k, p in lin.named_parameters()})
    hg = x.detach().float().permute(0, 2, 1).contiguous().cuda()           # [rows][T][H]
    yo = e.blstm(hg, q, rows, T, H)
    run_backward(e, yo, dy.permute(0, 2, 1))
    assert rel_l2(yo.view(rows, T, H).cpu(), y.detach().permute(0, 2, 1)) < 2e-5
    assert rel_l2(e.grad(hg).view(rows, T, H).cpu(), x.grad.permute(0, 2, 1)) < 5e-5
    worst = sorted([(rel_l2(e.pg[q + ".lstm.lstm." + k].cpu(), p.grad), k) for k, p in lstm.named_parameters()] +
                   [(rel_l2(e.pg[q + ".lstm.linear." + k].cpu(), p.grad), k) for k, p in lin.named_parameters()], reverse=True)
    print(worst[:3])
    assert worst[0][0] < 5e-5, worst[:4]


@pytest.mark.parametrize("H,T,rows", [(48, 95, 2), (96, 226, 1), (12, 33, 3)])
def test_local_state_attention_block(eng, H, T, rows):
    """LocalState (modules.py:74-127) against its einsum / softmax statement under autograd."""
    e = eng
    e._reset()
    heads, nd = 4, 4
    torch.manual_seed(SEED + 2)
    mods = torch.nn.ModuleDict({"query": torch.nn.Conv1d(H, H, 1), "key": torch.nn.Conv1d(H, H, 1), "content": torch.nn.Conv1d(H, H, 1),
                                "query_decay": torch.nn.Conv1d(H, heads * nd, 1), "proj": torch.nn.Conv1d(H, H, 1)}).double()
    with torch.no_grad():
        mods["query_decay"].bias.fill_(-1.0)
    x = rnd(rows, H, T, seed=1).double().requires_grad_(True)
    idx = torch.arange(T)
    delta = idx[:, None] - idx[None, :]
    qq = mods["query"](x).view(rows, heads, -1, T)
    kk = mods["key"](x).view(rows, heads, -1, T)
    dots = torch.einsum("bhct,bhcs->bhts", kk, qq) / math.sqrt(kk.shape[2])
    decays = torch.arange(1, nd + 1, dtype=torch.float64)
    dq = torch.sigmoid(mods["query_decay"](x).view(rows, heads, -1, T)) / 2
    dk = -decays.view(-1, 1, 1) * delta.abs() / math.sqrt(nd)
    dots = dots + torch.einsum("fts,bhfs->bhts", dk, dq)
    dots = dots.masked_fill(torch.eye(T, dtype=torch.bool), -100)
    w = torch.softmax(dots, dim=2)
    cc = mods["content"](x).view(rows, heads, -1, T)
    res = torch.einsum("bhts,bhct->bhcs", w, cc).reshape(rows, -1, T)
    y = x + mods["proj"](res)
    dy = rnd(*y.shape, seed=2).double()
    y.backward(dy)
    q = "enc.dconv.layers.0"
    e.params = {q + ".time_attn." + k: p.detach().float().cuda() for k, p in mods.named_parameters()}
    hg = x.detach().float().permute(0, 2, 1).contiguous().cuda()
    yo = e.local_attn(hg, q, rows, T, H)
    run_backward(e, yo, dy.permute(0, 2, 1))
    assert rel_l2(yo.view(rows, T, H).cpu(), y.detach().permute(0, 2, 1)) < 2e-5
    assert rel_l2(e.grad(hg).view(rows, T, H).cpu(), x.grad.permute(0, 2, 1)) < 5e-5
    # (key.bias shifts every score of a query by the same amount: softmax is invariant, its true gradient is zero)
    assert float(e.pg[q + ".time_attn.key.bias"].abs().max()) < 1e-5 * float(e.pg[q + ".time_attn.key.weight"].abs().max())
    worst = sorted([(rel_l2(e.pg[q + ".time_attn." + k].cpu(), p.grad), k) for k, p in mods.named_parameters() if k != "key.bias"],
                   reverse=True)
    print(worst[:3])
    assert worst[0][0] < 5e-5, worst[:4]


def test_istft_adjoint_matches_autograd(eng):
    """Backward of aero_istft_fwd through the STFT kernel's zero-pad / adjoint-scaling mode, against torch.istft under autograd."""
    e = eng
    g = e.geom
    B, Cout, T, Fq = 2, 1, 41, 256
    hop = g.hop_out
    out_len = hop * (T - 1) - 37
    z = rnd(B, Fq, T, 2 * Cout, seed=1).double().requires_grad_(True)
    zc = torch.view_as_complex(z.view(B, Fq, T, Cout, 2)).permute(0, 3, 1, 2)
    zc = F.pad(zc, (0, 0, 0, 1))
    y = torch.istft(zc.reshape(-1, Fq + 1, T), 512, hop, window=torch.hann_window(g.win_out).double(), win_length=g.win_out, normalized=True,
                    center=True)[..., :out_len]
    dy = rnd(*y.shape, seed=2).double()
    y.backward(dy)
    dz = e.istft_adjoint(dy.view(B, Cout, out_len).float().cuda(), B, Cout, T, Fq, out_len)
    torch.cuda.synchronize()
    assert rel_l2(dz.cpu(), z.grad) < 1e-5


def _oracle_train():
    from oracle import aero_oracle as O
    return O


@pytest.mark.parametrize("exp,B,L", [("aero_4-16_512_256", 2, 6000), ("aero_4-16_512_64", 2, 3600)])
def test_encoder_and_decoder_layers_as_blocks(exp, B, L):
    """Every encoder / decoder layer on its own, batch 2, against the oracle's functional layer (BatchNorm switched to batch
    statistics) under fp64 autograd on the CPU: input gradient and every parameter gradient of the layer."""
    from util import trained_like_
    O = _oracle_train()
    torch.manual_seed(SEED)
    m = Aero(**aero_kwargs(exp))
    m.load_state_dict(trained_like_(m.state_dict()))
    geom, kw = m.geom, m.geom.kw
    T = 1 + (L + (-L) % geom.hop_in) // geom.hop_in
    sd64 = {k: (v.double().requires_grad_(v.dtype.is_floating_point and "running" not in k and "num_batches" not in k)
                if v.dtype.is_floating_point else v) for k, v in m.state_dict().items()}
    m = m.cuda().train()
    O.BN_TRAIN = True
    try:
        for lg in geom.layers:
            # ---- encoder layer
            cin = lg.enc_cin
            x = rnd(B, cin, lg.f_in, T, seed=10 + lg.index).double().requires_grad_(True)
            y = O.enc_layer(x, sd64, lg, kw)
            dy = rnd(*y.shape, seed=20 + lg.index).double()
            for v in sd64.values():
                if torch.is_tensor(v) and v.grad is not None:
                    v.grad = None
            y.backward(dy)
            e = TrainEngine(m)
            e.params = {k: v.detach() for k, v in m.named_parameters()}
            e.buffers = {k: v.clone() for k, v in m.named_buffers()}
            xg = cl(x.detach().float()).cuda()
            kw_emb, kw["freq_emb"] = kw["freq_emb"], 0            # the oracle's enc_layer does not include the embedding add
            try:
                yo = e.encode(xg, lg, B, T)
            finally:
                kw["freq_emb"] = kw_emb
            run_backward(e, yo, cl(dy))
            pre = f"encoder.{lg.index}."
            assert rel_l2(yo.view(B, lg.f_out, T, lg.ch).cpu(), cl(y.detach())) < 2e-5
            rows = [(rel_l2(e.grad(xg).view(B, lg.f_in, T, cin).cpu(), cl(x.grad)), "dx")]
            gmax = max(float(v.grad.abs().max()) for k, v in sd64.items() if k.startswith(pre) and torch.is_tensor(v) and v.grad is not None)
            for k, v in sd64.items():
                if k.startswith(pre) and torch.is_tensor(v) and v.grad is not None:
                    if k.endswith(("freq_attn_block.conv1.0.bias", "freq_attn_block.conv1d.0.bias", "freq_attn_block.conv2.0.bias",
                                   "time_attn.key.bias")):
                        continue                                   # true gradient is zero (BatchNorm / softmax shift invariance)
                    got = e.pg[k].cpu().double()
                    rows.append((float((got - v.grad).norm() / max(float(v.grad.norm()), 1e-4 * gmax * v.numel() ** 0.5)), k))
            rows.sort(reverse=True)
            print(f"{exp} encoder.{lg.index}:", [(f"{a:.2e}", b) for a, b in rows[:4]])
            # everything downstream of the FTB block in the backward order is exact to fp32 round-off; the FTB parameters and dx sit
            # behind BatchNorm + ReLU, where one activation within fp32 rounding of the kink moves them by up to ~1e-2 (test_gpu_train.py)
            exact = [r for r in rows if "freq_attn_block" not in r[1] and "pre_conv" not in r[1] and r[1] != "dx"]
            assert exact[0][0] < 2e-4, exact[:6]
            assert rows[0][0] < 3e-2, rows[:6]
        for j, lg in enumerate(reversed(geom.layers)):
            last = lg.index == 0
            xin = None if j == 0 else rnd(B, lg.ch, lg.f_out, T, seed=30 + j).double().requires_grad_(True)
            skip = rnd(B, lg.ch, lg.f_out, T, seed=40 + j).double().requires_grad_(True)
            y = O.dec_layer(torch.zeros_like(skip) if xin is None else xin, skip, sd64, lg, j, kw, last)
            dy = rnd(*y.shape, seed=50 + j).double()
            for v in sd64.values():
                if torch.is_tensor(v) and v.grad is not None:
                    v.grad = None
            y.backward(dy)
            e = TrainEngine(m)
            e.params = {k: v.detach() for k, v in m.named_parameters()}
            e.buffers = {k: v.clone() for k, v in m.named_buffers()}
            xg = None if xin is None else cl(xin.detach().float()).cuda()
            sg = cl(skip.detach().float()).cuda()
            yo = e.decode(xg, sg, lg, j, B, T, last, None)
            run_backward(e, yo, cl(dy))
            pre = f"decoder.{j}."
            assert rel_l2(yo.view(B, y.shape[2], T, y.shape[1]).cpu(), cl(y.detach())) < 2e-5
            rows = [(rel_l2(e.grad(sg).view(B, lg.f_out, T, lg.ch).cpu(), cl(skip.grad)), "dskip")]
            if xin is not None:
                rows.append((rel_l2(e.grad(xg).view(B, lg.f_out, T, lg.ch).cpu(), cl(xin.grad)), "dx"))
            for k, v in sd64.items():
                if k.startswith(pre) and torch.is_tensor(v) and v.grad is not None:
                    if j == 0 and k.endswith("rewrite.weight"):
                        rows.append((rel_l2(e.pg[k][:, lg.ch:].cpu(), v.grad[:, lg.ch:]), k))
                    else:
                        rows.append((rel_l2(e.pg[k].cpu(), v.grad), k))
            rows.sort(reverse=True)
            print(f"{exp} decoder.{j}:", [(f"{a:.2e}", b) for a, b in rows[:4]])
            assert rows[0][0] < 2e-4, rows[:6]
    finally:
        O.BN_TRAIN = False


def test_debug_encoder0_intermediate_gradients():
    """Localises a gradient error inside encoder layer 0: activation gradients at the oracle's taps."""
    from util import trained_like_
    O = _oracle_train()
    exp, B, L = "aero_4-16_512_256", 2, 6000
    torch.manual_seed(SEED)
    m = Aero(**aero_kwargs(exp))
    m.load_state_dict(trained_like_(m.state_dict()))
    geom, kw = m.geom, m.geom.kw
    T = 1 + (L + (-L) % geom.hop_in) // geom.hop_in
    sd64 = {k: (v.double().requires_grad_("running" not in k) if v.dtype.is_floating_point else v) for k, v in m.state_dict().items()}
    m = m.cuda().train()
    lg = geom.layers[0]
    O.BN_TRAIN = True
    try:
        x = rnd(B, lg.enc_cin, lg.f_in, T, seed=10).double().requires_grad_(True)
        taps = {}
        inner = {}
        real_ftb = O.ftb

        def ftb_spy(xx, sd, prefix):
            Bq, Cq, Dq, Tq = xx.shape
            r_raw = F.conv2d(xx, sd[prefix + ".conv1.0.weight"], sd[prefix + ".conv1.0.bias"])
            r = torch.relu(O.batch_norm_eval(r_raw, sd, prefix + ".conv1.1"))
            g_raw = F.conv1d(r.reshape(Bq, -1, Tq), sd[prefix + ".conv1d.0.weight"], sd[prefix + ".conv1d.0.bias"], padding=4)
            g = torch.relu(O.batch_norm_eval(g_raw, sd, prefix + ".conv1d.1")).reshape(Bq, Cq, 1, Tq)
            att = g * xx
            att = (att.transpose(2, 3) @ sd[prefix + ".freq_fc.weight"].t()).transpose(2, 3)
            o_raw = F.conv2d(torch.cat([att, xx], 1), sd[prefix + ".conv2.0.weight"], sd[prefix + ".conv2.0.bias"])
            inner.update(R_raw=r_raw, R=r, G_raw=g_raw, G=g, Y=att, O_raw=o_raw)
            for t in inner.values():
                t.retain_grad()
            return torch.relu(O.batch_norm_eval(o_raw, sd, prefix + ".conv2.1"))
        O.ftb = ftb_spy
        try:
            y = O.enc_layer(x, sd64, lg, kw, taps=taps)
        finally:
            O.ftb = real_ftb
        for t in taps.values():
            t.retain_grad()
        dy = rnd(*y.shape, seed=20).double()
        y.backward(dy)
    finally:
        O.BN_TRAIN = False
    e = TrainEngine(m)
    e.params = {k: v.detach() for k, v in m.named_parameters()}
    e.buffers = {k: v.clone() for k, v in m.named_buffers()}
    from aero_b200.train_engine import _Conv as Cv
    p = "encoder.0"
    Fi, Fo, Cc = lg.f_in, lg.f_out, lg.ch
    xg = cl(x.detach().float()).cuda()
    a_pre = e.conv(xg, None, lg.enc_cin, 0, p + ".pre_conv.weight", p + ".pre_conv.bias", _C1x1, B, Fi, Fi, T, Cc)
    a_ftb = e.ftb(a_pre, p, B, Fi, T, Cc)
    y_raw = e.conv(a_ftb, None, Cc, 0, p + ".conv.weight", p + ".conv.bias", Cv(kf=lg.kernel, stride_f=lg.stride, pad_f=lg.pad), B, Fi, Fo, T, Cc)
    a_conv = e.norm_act(y_raw, cabi.NA_GELU, B=B, F_in=Fo, T=T, C_=Cc, scope=1, no_norm=True)
    a_dconv = e.dconv(a_conv, lg, B, T)
    raw = e.conv(a_dconv, None, Cc, 0, p + ".rewrite.weight", p + ".rewrite.bias", _C1x1, B, Fo, Fo, T, 2 * Cc)
    out = e.norm_act(raw, cabi.NA_GLU, B=B, F_in=Fo, T=T, C_=2 * Cc, scope=1, no_norm=True)
    e.acc(out, cl(dy).float().cuda().reshape(-1))
    held = {"pre_conv": a_pre, "ftb": a_ftb, "conv": a_conv, "dconv": a_dconv}
    grads = {}
    for fn in reversed(e.tape):
        fn()
        for k, t in held.items():
            if k not in grads and e.grad(t) is not None:
                pass
    torch.cuda.synchronize()
    print("forward:", {k: f"{rel_l2(t.view(B, -1, T, Cc).cpu(), cl(taps[p + '.' + k].detach())):.2e}" for k, t in held.items()})
    print("activation gradients:", {k: f"{rel_l2(e.grad(t).view(B, -1, T, Cc).cpu(), cl(taps[p + '.' + k].grad)):.2e}" for k, t in held.items()})
    print("dx:", rel_l2(e.grad(xg).view(B, Fi, T, lg.enc_cin).cpu(), cl(x.grad)))
    d = e._dbg
    rp = 8
    def g_of(name):
        return e.grad(d[name]).cpu()
    print("FTB internals (gradients):",
          "O_raw", f"{rel_l2(g_of('O_raw').view(B, Fi, T, Cc), cl(inner['O_raw'].grad)):.2e}",
          "Y", f"{rel_l2(g_of('Y').view(B, Fi, T, Cc), cl(inner['Y'].grad)):.2e}",
          "G", f"{rel_l2(g_of('G').view(B, T, Cc), inner['G'].grad.view(B, Cc, T).permute(0, 2, 1)):.2e}",
          "G_raw", f"{rel_l2(g_of('G_raw').view(B, T, Cc), inner['G_raw'].grad.permute(0, 2, 1)):.2e}",
          "R", f"{rel_l2(g_of('R').view(B, T, Fi, rp)[..., :5], inner['R'].grad.permute(0, 3, 2, 1)):.2e}",
          "R_raw", f"{rel_l2(g_of('R_raw').view(B, T, Fi, rp)[..., :5], inner['R_raw'].grad.permute(0, 3, 2, 1)):.2e}")
    # mask statistics of the last BatchNorm + ReLU
    o_ref = cl(inner["O_raw"].detach())                                   # [B, F, T, C] fp64
    mu, var = o_ref.mean((0, 1, 2)), o_ref.var((0, 1, 2), unbiased=False)
    gam, bet = sd64[p + ".freq_attn_block.conv2.1.weight"].detach(), sd64[p + ".freq_attn_block.conv2.1.bias"].detach()
    g_ref = (o_ref - mu) / torch.sqrt(var + 1e-5) * gam + bet
    o_our = d["O_raw"].view(B, Fi, T, Cc).cpu().double()
    g_our = (o_our - mu) / torch.sqrt(var + 1e-5) * gam + bet
    print("mean/std ratio per channel (max):", float((mu.abs() / var.sqrt()).max()), "min std", float(var.sqrt().min()))
    print("mask mismatches:", int(((g_ref > 0) != (g_our > 0)).sum()), "of", g_ref.numel(), "; |g| < 1e-4:", int((g_ref.abs() < 1e-4).sum()),
          "; |g| < 1e-3:", int((g_ref.abs() < 1e-3).sum()))
    db_ref = sd64[p + ".freq_attn_block.conv2.1.bias"].grad
    db_our = e.pg[p + ".freq_attn_block.conv2.1.bias"].cpu().double()
    print("dbeta per-channel rel err (first 12):", [f"{float(abs(a - b) / (abs(b) + 1e-30)):.1e}" for a, b in zip(db_our[:12], db_ref[:12])])
    dy_ref = cl(taps[p + ".ftb"].grad)
    print("ref dbeta recomputed from ref dy and ref mask:", rel_l2((dy_ref * (g_ref > 0)).sum((0, 1, 2)), db_ref))
    dy_our = e.grad(a_ftb).view(B, Fi, T, Cc).cpu().double()
    print("our dy with ref mask:", rel_l2((dy_our * (g_ref > 0)).sum((0, 1, 2)), db_ref), " our dy abs max", float(dy_our.abs().max()),
          " ref", float(dy_ref.abs().max()))
    # the BatchNorm + ReLU backward kernel alone on the real data
    import ctypes as C
    from aero_b200.train_engine import _ptr
    prm = cabi.NormActParams(B, Fi, Fi, 0, T, Cc, 1, 3, cabi.NA_RELU, 1e-5, 0)
    dgm, dbt = torch.zeros(Cc, dtype=torch.float64, device="cuda"), torch.zeros(Cc, dtype=torch.float64, device="cuda")
    dyk = e.grad(a_ftb)
    gam32, bet32 = e.params[p + ".freq_attn_block.conv2.1.weight"], e.params[p + ".freq_attn_block.conv2.1.bias"]
    cabi.check(e.lib.aero_norm_act_train_bwd(_ptr(d["O_raw"]), _ptr(d["st3"]), _ptr(gam32), _ptr(bet32), None, None, _ptr(dyk), None, _ptr(dgm),
                                             _ptr(dbt), None, None, None, 1, C.byref(prm), e._stream()), e.lib)
    torch.cuda.synchronize()
    st = d["st3"].cpu()
    n = B * Fi * T
    mu32 = (st[:, 0] / n).float()
    r32 = (1.0 / torch.sqrt((st[:, 1] / n - (st[:, 0] / n) ** 2).clamp_min(0) + 1e-5)).float()
    o32 = d["O_raw"].view(B, Fi, T, Cc).cpu()
    g32 = torch.addcmul(bet32.cpu(), (o32 - mu32) * r32, gam32.cpu())
    want = (dyk.view(B, Fi, T, Cc).cpu().double() * (g32 > 0)).sum((0, 1, 2))
    print("kernel alone: dbeta vs exact sum of the same dy / mask:", rel_l2(dbt.cpu(), want), " vs ref:", rel_l2(dbt.cpu(), db_ref),
          " stats mean vs ref mean:", rel_l2(st[:, 0] / n, mu), " engine pgrad vs kernel-alone:", rel_l2(db_our, dbt.cpu()))
    print("FTB internals (forward):",
          "O_raw", f"{rel_l2(d['O_raw'].view(B, Fi, T, Cc).cpu(), cl(inner['O_raw'].detach())):.2e}",
          "G", f"{rel_l2(d['G'].view(B, T, Cc).cpu(), inner['G'].detach().view(B, Cc, T).permute(0, 2, 1)):.2e}")
    for k, v in sd64.items():
        if k.startswith(p + ".") and torch.is_tensor(v) and v.grad is not None and k in e.pg:
            print(f"   {rel_l2(e.pg[k].cpu(), v.grad):.2e} {k}")


def test_mrstft_loss_gradient_matches_autograd():
    """d loss / d estimate of aero_b200.losses.MultiResolutionSTFTLoss against autograd through the
    reference formula (stft_loss.py:11-63,96-138 with the torch>=2 return_complex shim), fp64 on the CPU."""
    from aero_b200.losses import MultiResolutionSTFTLoss
    B, L = 2, 9000
    y = rnd(B, L, seed=1)
    x = (y + 0.4 * rnd(B, L, seed=2))
    x[:, :1500] = 0.0                                       # exercises the 1e-7 clamp (zero gradient there)
    xd = x.double().requires_grad_(True)
    sc_t = mag_t = 0.0
    for n_fft, hop, win in ((1024, 120, 600), (2048, 240, 1200), (512, 50, 240)):
        w = torch.hann_window(win).double()
        X = torch.stft(xd, n_fft, hop, win, w, return_complex=True)
        Y = torch.stft(y.double(), n_fft, hop, win, w, return_complex=True)
        xm = torch.sqrt(torch.clamp(X.real ** 2 + X.imag ** 2, min=1e-7))
        ym = torch.sqrt(torch.clamp(Y.real ** 2 + Y.imag ** 2, min=1e-7))
        sc_t = sc_t + torch.norm(ym - xm, p="fro") / torch.norm(ym, p="fro")
        mag_t = mag_t + F.l1_loss(torch.log(ym), torch.log(xm))
    sc_t, mag_t = 0.1 * sc_t / 3, 0.1 * mag_t / 3
    (0.7 * sc_t + 1.3 * mag_t).backward()
    xg = x.cuda().requires_grad_(True)
    sc, mag = MultiResolutionSTFTLoss()(xg, y.cuda())
    (0.7 * sc + 1.3 * mag).backward()
    torch.cuda.synchronize()
    print(f"mrstft: sc {float(sc):.6f} ({float(sc_t):.6f}) mag {float(mag):.6f} ({float(mag_t):.6f}); grad rel_l2 {rel_l2(xg.grad.cpu(), xd.grad):.2e}")
    assert abs(float(sc) - float(sc_t)) < 2e-5 * float(sc_t) and abs(float(mag) - float(mag_t)) < 1e-4 * float(mag_t)
    assert rel_l2(xg.grad.cpu(), xd.grad) < 5e-4          # (sign() of the log-magnitude term flips where the two magnitudes agree to fp32)


@pytest.mark.parametrize("Cin,Cout,groups,Tin", [(16, 64, 4, 1030), (64, 256, 16, 515), (256, 1024, 64, 259), (1024, 1024, 256, 77),
                                                 (8, 32, 2, 300)])
def test_melgan_grouped_conv_kernels(Cin, Cout, groups, Tin):
    """aero_gconv1d_{fwd,dgrad,wgrad} (k = 41, stride 4, pad 20: the register-tiled MelGAN kernels where C_out is a multiple of 64, the
    generic ones otherwise) against F.conv1d under autograd in fp64; ragged lengths (neither T_out nor T_in a tile multiple)."""
    import ctypes as C
    lib = cabi.load()
    B, k, stride, pad = 2, 41, 4, 20
    x = rnd(B, Cin, Tin, seed=1).double().requires_grad_(True)
    w = (rnd(Cout, Cin // groups, k, seed=2) / math.sqrt(k * Cin // groups)).double().requires_grad_(True)
    b = rnd(Cout, seed=3).double()
    ref = F.conv1d(x, w, b, stride=stride, padding=pad, groups=groups)
    Tout = ref.shape[-1]
    dy = rnd(*ref.shape, seed=4).double()
    ref.backward(dy)
    dev = torch.device("cuda")
    xg = x.detach().float().permute(0, 2, 1).contiguous().to(dev)               # [B, T, C]
    wg, bg = w.detach().float().contiguous().to(dev), b.float().to(dev)
    dyg = dy.float().permute(0, 2, 1).contiguous().to(dev)
    y = torch.empty(B, Tout, Cout, device=dev)
    dx = torch.full((B, Tin, Cin), float("nan"), device=dev)
    dw = torch.zeros_like(wg)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = (B, Tin, Tout, Cin, Cout, groups, k, stride, pad)
    P = lambda t: C.c_void_p(t.data_ptr())
    cabi.check(lib.aero_gconv1d_fwd(P(xg), P(wg), P(bg), P(y), *args, st), lib)
    cabi.check(lib.aero_gconv1d_dgrad(P(dyg), P(wg), P(dx), *args, st), lib)
    cabi.check(lib.aero_gconv1d_wgrad(P(xg), P(dyg), P(dw), *args, st), lib)
    torch.cuda.synchronize()
    assert rel_l2(y.cpu().permute(0, 2, 1), ref.detach()) < 1e-5
    assert rel_l2(dx.cpu().permute(0, 2, 1), x.grad) < 1e-5
    assert rel_l2(dw.cpu(), w.grad) < 1e-5
