"""RNN-T joint + loss + greedy search (scope row f4) against the reference's own arithmetic: joint.py:20-38 restated with
torch ops, torchaudio.functional.rnnt_loss (the call of model.py:106-111), and the greedy loop of model.py:221-269 run
with a torch joint.  pytest -m gpu."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import conformer_pytorch_lightning_b200 as C
from conformer_pytorch_lightning_b200 import _native


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _torch_joint(j, enc, pred):
    e = torch.nn.functional.linear(enc, j.enc_ffn.weight, j.enc_ffn.bias).unsqueeze(2)
    p = torch.nn.functional.linear(pred, j.pred_ffn.weight, j.pred_ffn.bias).unsqueeze(1)
    return torch.nn.functional.linear(torch.tanh(e + p), j.ffn_out.weight, j.ffn_out.bias)


def _setup(dtype, B=3, T=37, U=9, V=301, E=256, P=256, J=512):
    torch.manual_seed(0)
    joint = C.TransducerJoint(V, E, P, J).cuda()
    joint.compute_dtype = dtype
    g = torch.Generator().manual_seed(1)
    enc = torch.randn(B, T, E, generator=g).cuda()
    pred = torch.randn(B, U + 1, P, generator=g).cuda()
    targets = torch.randint(1, V, (B, U), generator=g).to(torch.int32).cuda()
    t_len = torch.tensor([T, T - 5, T - 11], dtype=torch.int32).cuda()[:B]
    u_len = torch.tensor([U, U - 2, 3], dtype=torch.int32).cuda()[:B]
    return joint, enc, pred, targets, t_len, u_len


def test_state_dict_layout_of_joint_and_predictor():
    j = C.TransducerJoint(50, 256, 256, 512)
    assert list(j.state_dict().keys()) == ["enc_ffn.weight", "enc_ffn.bias", "pred_ffn.weight", "pred_ffn.bias",
                                          "ffn_out.weight", "ffn_out.bias"]
    p = C.RNNPredictor(50, 64, 256, 128, 0.1, 2)
    keys = list(p.state_dict().keys())
    assert keys[0] == "embed.weight" and keys[-2:] == ["projection.weight", "projection.bias"] and "rnn.weight_ih_l1" in keys


@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float32, 1e-4, 1e-3), (torch.bfloat16, 2e-2, 5e-2)])
def test_joint_and_rnnt_loss_forward_backward(dtype, tol, gtol):
    import torchaudio.functional as AF
    joint, enc, pred, targets, t_len, u_len = _setup(dtype)
    enc_r, pred_r = enc.clone().requires_grad_(), pred.clone().requires_grad_()
    logits_r = _torch_joint(joint, enc_r, pred_r)
    loss_r = AF.rnnt_loss(logits_r, targets, t_len, u_len, blank=0, reduction="mean")
    grads_r = torch.autograd.grad(loss_r, [enc_r, pred_r] + list(joint.parameters()))
    n0 = _native.kernel_launches("rnnt_grad")
    enc_n, pred_n = enc.clone().requires_grad_(), pred.clone().requires_grad_()
    logits = joint(enc_n, pred_n)
    assert tuple(logits.shape) == tuple(logits_r.shape)
    assert _rel(logits.float(), logits_r) < tol
    loss = C.rnnt_loss(logits, targets, t_len, u_len, blank=0, reduction="mean")
    assert abs(loss.item() - loss_r.item()) < (1e-4 if dtype == torch.float32 else 1e-2) * abs(loss_r.item())
    grads = torch.autograd.grad(loss, [enc_n, pred_n] + list(joint.parameters()))
    assert _native.kernel_launches("rnnt_grad") == n0 + 1
    for g, gr, name in zip(grads, grads_r, ["enc", "pred"] + [k for k, _ in joint.named_parameters()]):
        assert _rel(g.float(), gr) < gtol, name


def test_rnnt_loss_alone_matches_torchaudio_on_fp32_logits():
    """rnnt_loss as a stand-alone replacement of torchaudio.functional.rnnt_loss: per-utterance values and the logit
    gradient, ragged lengths, reduction sum / none."""
    import torchaudio.functional as AF
    g = torch.Generator().manual_seed(4)
    B, T, U, V = 4, 29, 7, 123
    logits = (torch.randn(B, T, U + 1, V, generator=g) * 2).cuda()
    targets = torch.randint(0, V - 1, (B, U), generator=g).to(torch.int32).cuda()
    t_len = torch.tensor([29, 20, 29, 11], dtype=torch.int32).cuda()
    u_len = torch.tensor([7, 7, 1, 4], dtype=torch.int32).cuda()
    ref = AF.rnnt_loss(logits, targets, t_len, u_len, blank=V - 1, reduction="none")
    got = C.rnnt_loss(logits, targets, t_len, u_len, blank=-1, reduction="none")
    assert _rel(got, ref) < 1e-5
    lr = logits.clone().requires_grad_()
    AF.rnnt_loss(lr, targets, t_len, u_len, blank=V - 1, reduction="sum").backward()
    ln = logits.clone().requires_grad_()
    C.rnnt_loss(ln, targets, t_len, u_len, blank=V - 1, reduction="sum").backward()
    assert _rel(ln.grad, lr.grad) < 1e-4


def _reference_greedy(predictor, joint_fn, encoder_out, n_frames, blank, n_steps):
    """Test-side restatement of the reference's loop (model.py:221-269), statement by statement, with a torch joint."""
    dev = encoder_out.device
    padding = torch.zeros(1, 1, device=dev)
    pred_input_step = torch.tensor([blank], device=dev).reshape(1, 1)
    cache = predictor.init_state(pred_input_step)
    new_cache = []
    t, hyps, prev_out_nblk, pred_out_step, per_frame_noblk = 0, [], True, None, 0
    while t < n_frames:
        encoder_out_step = encoder_out[:, t:t + 1, :]
        if prev_out_nblk:
            pred_out_step, new_cache = predictor.forward_step(pred_input_step, padding, cache)
        joint_out_max = int(joint_fn(encoder_out_step, pred_out_step).log_softmax(dim=-1).argmax(dim=-1).squeeze())
        if joint_out_max != blank:
            hyps.append(joint_out_max)
            prev_out_nblk = True
            per_frame_noblk += 1
            pred_input_step = torch.tensor([joint_out_max], device=dev).reshape(1, 1)
            cache = new_cache
        if joint_out_max == blank or per_frame_noblk >= n_steps:
            if joint_out_max == blank:
                prev_out_nblk = False
            t += 1
            per_frame_noblk = 0
    return hyps, (pred_input_step, cache)


@pytest.mark.parametrize("n_steps", [1, 4])
def test_greedy_search_equals_reference_loop_with_torch_joint(n_steps):
    torch.manual_seed(3)
    V = 40
    joint = C.TransducerJoint(V, 256, 256, 512).cuda().eval()
    with torch.no_grad():
        joint.ffn_out.bias[0] -= 1.0            # make blanks rarer so that several symbols per frame occur
    pred = C.RNNPredictor(V, 64, 256, 128, 0.0, 2, dropout=0.0).cuda().eval()
    enc_out = torch.randn(1, 23, 256, device="cuda") * 2
    with torch.no_grad():
        ref, (last_r, cache_r) = _reference_greedy(pred, lambda e, p: _torch_joint(joint, e, p), enc_out, 23, 0, n_steps)
    hyps, (last, cache) = C.basic_greedy_search(pred, joint, enc_out, torch.tensor(23), blank=0, n_steps=n_steps)
    assert hyps == ref and len(hyps) > 0
    assert torch.equal(last, last_r) and torch.allclose(cache[0], cache_r[0])
