"""tcgen05 conv conditioner (csrc/conv_cond_tc.cu) through the C ABI against an fp64 torch evaluation of Coupling.NN
(reference layers/coupling.py:26-29) and against the FP32-FMA kernel: every BASELINE level shape, the time-series (3,1) kernel,
ragged batches, several tiles per CTA, per-sample first-layer bias (conventional context), reading x0 in place from a wider tensor."""
import pytest
import torch
import torch.nn.functional as F

from contextflow_b200 import ops, synth

pytestmark = pytest.mark.gpu
dev = 'cuda'


def _weights(tag, cin, ch, cout, kh, kw, extra=0):
    w1 = synth.uniform(tag + 'w1', (ch, cin + extra)) * (1.5 / (cin + extra) ** 0.5)
    w2 = synth.uniform(tag + 'w2', (ch, ch, kh, kw)) * (1.5 / (ch * kh * kw) ** 0.5)
    w3 = synth.uniform(tag + 'w3', (cout, ch)) * (1.5 / ch ** 0.5)
    b1, b2, b3 = synth.uniform(tag + 'b1', (ch,)) * 0.3, synth.uniform(tag + 'b2', (ch,)) * 0.3, synth.uniform(tag + 'b3', (cout,)) * 0.3
    return w1, b1, w2, b2, w3, b3


def _ref64(x0, w1, b1, w2, b2, w3, b3, bias1_b=None):
    x0 = x0.double()
    cin = x0.shape[1]
    h = F.conv2d(x0, w1[:, :cin].double()[:, :, None, None])
    h = h + (b1.double()[None, :, None, None] if bias1_b is None else bias1_b.double()[:, :, None, None])
    h = F.relu(h)
    kh, kw = w2.shape[2], w2.shape[3]
    if kh > 1 or kw > 1:
        h = F.pad(h, (kw // 2, kw // 2, kh // 2, kh // 2), mode='reflect')
    h = F.relu(F.conv2d(h, w2.double(), b2.double()))
    return F.conv2d(h, w3.double()[:, :, None, None], b3.double())


def _run_tc(x, cin, w1, b1, w2, b2, w3, b3, bias1_b=None):
    ch, cout, kh, kw = w2.shape[0], w3.shape[0], w2.shape[2], w2.shape[3]
    pack = ops.conv_cond_tc_pack(w1.to(dev), w2.to(dev), w3.to(dev), cin)
    assert pack is not None, 'shape should have a tensor-core plan'
    h = ops.conv_cond_tc(x, cin, pack, b1.to(dev), b2.to(dev), b3.to(dev), ch, x.shape[2], x.shape[3], kh, kw, cout,
                         bias1_b=None if bias1_b is None else bias1_b.to(dev))
    assert h is not None, 'shape should have a tensor-core plan'
    torch.cuda.synchronize()
    return h


def _check(h, want, what):
    scale = want.abs().max().item()
    err = (h.double().cpu() - want).abs().max().item()
    assert err <= 2e-5 * scale + 1e-6, f'{what}: max abs err {err:.3e} vs |ref|max {scale:.3e} (fp32-faithful budget 2e-5)'


SHAPES = [  # B, C (coupling channels: Cin=C/2, Ch=2C, Cout=C), H, W, KH, KW
    (5, 16, 16, 16, 3, 3),      # cfg2 level 1 (segment layout, two segments per row)
    (7, 32, 8, 8, 3, 3),        # cfg2 level 2 / cfg1 level 2 (segment layout, samples interleaved)
    (10, 64, 4, 4, 3, 3),       # cfg2 level 3 (plain padded layout, 4 channel panels)
    (3, 8, 16, 16, 3, 3),       # cfg1 level 1 (Cin=4 -> zero-padded k-step, N=16)
    (1, 16, 16, 16, 3, 3),
    (700, 32, 8, 8, 3, 3),      # several tiles per persistent CTA
    (333, 16, 16, 16, 3, 3),
    (9, 56, 8, 1, 3, 1),        # MSL-shaped conv coupling: 56 channels (Ch=112, partial last panel), (3,1) kernel
    (6, 8, 14, 14, 3, 3),       # MNIST 28x28 level 1 (W % 8 != 0 -> plain layout)
    (4, 32, 7, 7, 3, 3),        # MNIST 28x28 level 2
    (12, 24, 6, 10, 3, 3),      # Ch=48 (two panels, second half full), non-square
    (5, 32, 4, 24, 3, 3),       # three segments per row
    (33, 16, 1, 1, 1, 1),       # 1x1 kernel on single pixels
]


@pytest.mark.parametrize('B,C,H,W,KH,KW', SHAPES)
def test_tc_conditioner_matches_fp64(B, C, H, W, KH, KW):
    cin, ch, cout = C // 2, 2 * C, C
    tag = f'tc{B}.{C}.{H}.{W}'
    w = _weights(tag, cin, ch, cout, KH, KW)
    x = synth.uniform(tag + 'x', (B, C, H, W)) * 2.0
    xd = x.to(dev)
    h = _run_tc(xd, cin, *w)                                   # x0 = first half of x, read in place (batch stride C*H*W)
    want = _ref64(x[:, :cin], *w)
    _check(h, want, f'tc conditioner {B}x{C}x{H}x{W} k{KH}x{KW}')
    plan = ops.conv_cond_tc_last_plan()
    assert plan['ntiles'] == (B + plan['S'] - 1) // plan['S']
    # and the FP32-FMA kernel computes the same function
    pk = (ops.pack_kmajor(w[0][:, :cin].to(dev)), ops.pad_vec(w[1].to(dev)), ops.pack_kmajor(w[2].reshape(ch, -1).to(dev)),
          ops.pad_vec(w[3].to(dev)), ops.pack_kmajor(w[4].to(dev)), ops.pad_vec(w[5].to(dev)))
    try:
        h_fma = ops.conv_cond(xd, cin, pk, H, W, KH, KW, cout)
    except RuntimeError:
        return                                                 # shape outside the FMA kernel's tile limits
    _check(h_fma, want, 'fma conditioner')


def test_tc_per_sample_bias_and_plan_layouts():
    B, C, H, W = 11, 32, 8, 8
    cin, ch, cout = C // 2, 2 * C, C
    w1, b1, w2, b2, w3, b3 = _weights('tcb', cin, ch, cout, 3, 3, extra=C)
    cn = synth.uniform('tcb.cn', (B, C))
    bias1 = b1[None, :] + cn @ w1[:, cin:].t()                  # conventional concat == per-sample bias (coupling.py:47)
    x = synth.uniform('tcb.x', (B, C, H, W)) * 2.0
    h = _run_tc(x.to(dev), cin, w1, b1, w2, b2, w3, b3, bias1_b=bias1)
    xin = torch.cat([x[:, :cin], cn[:, :, None, None].expand(B, C, H, W)], 1).double()
    hh = F.relu(F.conv2d(xin, w1.double()[:, :, None, None], b1.double()))
    hh = F.relu(F.conv2d(F.pad(hh, (1, 1, 1, 1), mode='reflect'), w2.double(), b2.double()))
    want = F.conv2d(hh, w3.double()[:, :, None, None], b3.double())
    _check(h, want, 'tc conditioner, per-sample bias')
    assert ops.conv_cond_tc_last_plan()['seg'] == 1
    _run_tc(synth.uniform('tcb.y', (4, 128, 4, 4)).to(dev), 64 // 2, *_weights('tcb4', 32, 128, 64, 3, 3))
    assert ops.conv_cond_tc_last_plan()['seg'] == 0


def test_tc_unsupported_shapes_fall_back():
    from contextflow_b200 import _cabi
    lib = _cabi.lib()
    assert lib.cfpp_conv_cond_tc_pack_bytes(10, 40, 20, 1, 1) == -1            # Ch % 16 != 0 (CouplingFC of a 20-wide encoder)
    assert lib.cfpp_conv_cond_tc_pack_bytes(38, 152, 76, 3, 1) == -1           # Ch > 128
    assert lib.cfpp_conv_cond_tc_supported(8, 8, 32, 16, 16, 16, 3, 3, 16 * 256) == 1
    assert lib.cfpp_conv_cond_tc_supported(8, 8, 32, 16, 16, 16, 3, 3, 16 * 256 + 2) == 0   # batch stride not 16-byte aligned


@pytest.mark.parametrize('B,C,H,W', [(5, 16, 16, 16), (9, 32, 8, 8), (10, 64, 4, 4), (3, 16, 14, 14), (301, 32, 8, 8)])
@pytest.mark.parametrize('ctx', [False, True])
def test_tc_fused_coupling_matches_fp64(B, C, H, W, ctx):
    """cfpp_conv_cond_tc_coupling_fwd: conditioner + affine coupling in one kernel (coupling.py:39-66) against fp64."""
    cin, ch = C // 2, 2 * C
    tag = f'tcf{B}.{C}.{H}.{W}'
    w1, b1, w2, b2, w3, b3 = _weights(tag, cin, ch, C, 3, 3)
    x = synth.uniform(tag + 'x', (B, C, H, W)) * 2.0 - 1.0
    add = (synth.uniform(tag + 'a', (B, C)) - 0.5) if ctx else None
    lp = synth.uniform(tag + 'l', (B,)) if ctx else None
    pack = ops.conv_cond_tc_pack(w1.to(dev), w2.to(dev), w3.to(dev), cin)
    out = ops.conv_cond_tc_coupling(x.to(dev), pack, b1.to(dev), b2.to(dev), b3.to(dev), ch, 3, 3,
                                    add=None if add is None else add.to(dev), logp_c=None if lp is None else lp.to(dev), logp_scale=float(H * W))
    assert out is not None, 'shape should have a fused plan'
    z, ldj = out
    h = _ref64(x[:, :cin], w1, b1, w2, b2, w3, b3)
    if ctx:
        h = h + add.double()[:, :, None, None]
    t, r = h[:, :cin], h[:, cin:]
    ls = 2 * torch.tanh(r / 2)
    want_z = torch.cat([x[:, :cin].double(), x[:, cin:].double() * torch.exp(ls) + t], 1)
    want_l = ls.sum((1, 2, 3)) + (lp.double() * H * W if ctx else 0)
    assert torch.equal(z[:, :cin].cpu(), x[:, :cin]), 'pass-through half must be bit exact'
    from tests.helpers import L_ATOL, L_RTOL, Z_ATOL, Z_RTOL, assert_close
    assert_close(z.cpu().numpy(), want_z.numpy(), Z_RTOL, Z_ATOL * max(1.0, float(want_z.abs().max())), 'fused coupling z')
    assert_close(ldj.cpu().numpy(), want_l.numpy(), L_RTOL, L_ATOL, 'fused coupling ldj')


@pytest.mark.parametrize('env', [{'CFPP_TC_PIPE': '0'}, {'CFPP_TC_PIPE': '1'}, {'CFPP_TC_PIPE': '0', 'CFPP_TC_OCC': '1'}])
@pytest.mark.parametrize('B,C,H,W', [(700, 16, 16, 16), (1300, 32, 8, 8), (2200, 64, 4, 4)])
def test_tc_fused_coupling_every_plan_form_with_several_tiles_per_cta(B, C, H, W, env, monkeypatch):
    """Every CTA walks several tiles (ntiles > 2 x 148), so the software-pipelined form (tile i+1's operand set and per-sample
    bias staged before tile i's last epilogue) and the two-resident-CTA form are each compared with fp64 (coupling.py:39-66)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cin, ch = C // 2, 2 * C
    tag = f'tcp{B}.{C}.{H}.{W}'
    w1, b1, w2, b2, w3, b3 = _weights(tag, cin, ch, C, 3, 3)
    x = synth.uniform(tag + 'x', (B, C, H, W)) * 2.0 - 1.0
    add = synth.uniform(tag + 'a', (B, C)) - 0.5
    lp = synth.uniform(tag + 'l', (B,))
    pack = ops.conv_cond_tc_pack(w1.to(dev), w2.to(dev), w3.to(dev), cin)
    out = ops.conv_cond_tc_coupling(x.to(dev), pack, b1.to(dev), b2.to(dev), b3.to(dev), ch, 3, 3, add=add.to(dev), logp_c=lp.to(dev), logp_scale=float(H * W))
    assert out is not None
    z, ldj = out
    h = _ref64(x[:, :cin], w1, b1, w2, b2, w3, b3) + add.double()[:, :, None, None]
    t, r = h[:, :cin], h[:, cin:]
    ls = 2 * torch.tanh(r / 2)
    want_z = torch.cat([x[:, :cin].double(), x[:, cin:].double() * torch.exp(ls) + t], 1)
    want_l = ls.sum((1, 2, 3)) + lp.double() * H * W
    assert torch.equal(z[:, :cin].cpu(), x[:, :cin])
    from tests.helpers import L_ATOL, L_RTOL, Z_ATOL, Z_RTOL, assert_close
    assert_close(z.cpu().numpy(), want_z.numpy(), Z_RTOL, Z_ATOL * max(1.0, float(want_z.abs().max())), 'fused coupling z')
    assert_close(ldj.cpu().numpy(), want_l.numpy(), L_RTOL, L_ATOL, 'fused coupling ldj')
    # the unfused conditioner under the same plan form
    hh = ops.conv_cond_tc(x[:, :cin].contiguous().to(dev), cin, pack, b1.to(dev), b2.to(dev), b3.to(dev), ch, H, W, 3, 3, C)
    if hh is not None:
        want_h = _ref64(x[:, :cin], w1, b1, w2, b2, w3, b3)
        assert_close(hh.cpu().numpy(), want_h.numpy(), Z_RTOL, Z_ATOL * max(1.0, float(want_h.abs().max())), 'conditioner h')


@pytest.mark.parametrize('B,C,H,W', [(5, 16, 16, 16), (9, 32, 8, 8), (10, 64, 4, 4), (301, 32, 8, 8), (700, 16, 16, 16), (2200, 64, 4, 4)])
def test_tc_training_forward_writes_the_activations_the_backward_reads(B, C, H, W):
    """cfpp_conv_cond_tc_train_fwd: h plus a1 = relu(W1 x0 + b1), a2 = relu(W2 * a1 + b2) (coupling.py:26-29), each against fp64 and
    against the three-launch FP32 convolution route the training step used before."""
    cin, ch = C // 2, 2 * C
    tag = f'tct{B}.{C}.{H}.{W}'
    w1, b1, w2, b2, w3, b3 = _weights(tag, cin, ch, C, 3, 3)
    x = (synth.uniform(tag + 'x', (B, C, H, W)) * 2.0 - 1.0).to(dev)
    pack = ops.conv_cond_tc_pack(w1.to(dev), w2.to(dev), w3.to(dev), cin)
    out = ops.conv_cond_tc_train(x, cin, pack, b1.to(dev), b2.to(dev), b3.to(dev), ch, H, W, 3, 3, C)
    assert out is not None, 'shape should have a plan'
    h, a1, a2 = out
    x0 = x[:, :cin].double().cpu()
    r1 = F.relu(F.conv2d(x0, w1[:, :cin].double()[:, :, None, None], b1.double()))
    r2 = F.relu(F.conv2d(F.pad(r1, (1, 1, 1, 1), mode='reflect'), w2.double(), b2.double()))
    r3 = F.conv2d(r2, w3.double()[:, :, None, None], b3.double())
    from tests.helpers import Z_ATOL, Z_RTOL, assert_close
    for got, want, what in ((a1, r1, 'a1'), (a2, r2, 'a2'), (h, r3, 'h')):
        assert_close(got.cpu().numpy(), want.numpy(), Z_RTOL, Z_ATOL * max(1.0, float(want.abs().max())), f'training forward {what}')
    # every element of a1 / a2 is written (halo rows rewrite their mirror pixel), none left from the allocation
    assert torch.isfinite(a1).all() and torch.isfinite(a2).all()
    f1 = ops.conv2d_fwd(x, cin, w1.to(dev)[:, :cin, None, None].contiguous(), b1.to(dev), relu=True)
    f2 = ops.conv2d_fwd(f1, ch, w2.to(dev), b2.to(dev), relu=True)
    assert (a1 - f1).abs().max().item() <= 2e-5 * max(1.0, float(f1.abs().max()))
    assert (a2 - f2).abs().max().item() <= 2e-5 * max(1.0, float(f2.abs().max()))
