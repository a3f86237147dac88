"""RNN-T joint, loss and greedy search (scope row f4).

``TransducerJoint`` is the drop-in for the reference's module of the same name (joint.py:4-38: same constructor,
parameter names ``enc_ffn`` / ``pred_ffn`` / ``ffn_out`` -- and its ``activatoin`` attribute -- and
``forward(enc_out, pred_out, pre_project=True)`` -> (B, T, U+1, V) logits); ``rnnt_loss`` replaces the
``torchaudio.functional.rnnt_loss`` call of model.py:95-113; ``RNNPredictor`` mirrors predictor.py:14-86 (embedding + LSTM +
projection: the LSTM stays torch / cuDNN, it is a few hundred microseconds of the reference's time and outside the
encoder hot path); ``basic_greedy_search`` is the per-frame loop of model.py:221-269 with the joint step on the native
kernels.

The joint's three Linear layers run on the tcgen05 GEMMs (forward: gemm_tc; backward: the transposed-operand GEMM), the
broadcast add + tanh, its backward with the two broadcast reductions, the vocabulary log-softmax statistics, the lattice
recursions and the logit gradient are native kernels (csrc/rnnt.cu).  The (B, T, U+1, V) logits ARE materialised, like in
the reference (3.2 GB in fp32 at B=16, T=248, U=40, V=5002; half of that on the bf16 path).
"""
import torch
import torch.nn as nn

from . import _native as N
from . import engine, ops
from . import train_ops as TO

__all__ = ["TransducerJoint", "RNNPredictor", "rnnt_loss", "basic_greedy_search"]


def _pad_rows(w, b, mult, dtype):
    """(V, J) weight / (V,) bias -> zero-padded to a multiple of ``mult`` rows (GEMM tile width)."""
    V, J = w.shape
    Vp = (V + mult - 1) // mult * mult
    wp = torch.zeros((Vp, J), dtype=dtype, device=w.device)
    wp[:V] = w.detach().to(dtype)
    bp = torch.zeros(Vp, dtype=torch.float32, device=w.device)
    if b is not None:
        bp[:V] = b.detach().float()
    return wp, bp, Vp


class _JointFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc_out, pred_out, we, be, wp, bp, wo, bo, dtype, pre_project):
        B, T, E = enc_out.shape
        U1 = pred_out.shape[1]
        dev = enc_out.device
        x_e = enc_out.reshape(B * T, E).to(dtype).contiguous()
        x_p = pred_out.reshape(B * U1, pred_out.shape[2]).to(dtype).contiguous()
        if pre_project:
            J = we.shape[0]
            e = torch.empty((B * T, J), dtype=dtype, device=dev)
            p = torch.empty((B * U1, J), dtype=dtype, device=dev)
            ops.gemm(x_e, we.detach().to(dtype).contiguous(), be.detach().float().contiguous(), e, N.EPI_BIAS)
            ops.gemm(x_p, wp.detach().to(dtype).contiguous(), bp.detach().float().contiguous(), p, N.EPI_BIAS)
        else:
            e, p = x_e, x_p
            J = e.shape[1]
        z = torch.empty((B * T * U1, J), dtype=dtype, device=dev)
        N.check(N.lib().cfm_joint_add_tanh(e.data_ptr(), p.data_ptr(), z.data_ptr(), B, T, U1, J, ops._DT[dtype],
                                           ops._stream(z)))
        V = wo.shape[0]
        wop, bop, Vp = _pad_rows(wo, bo, 128, dtype)
        logits = torch.empty((B * T * U1, Vp), dtype=dtype, device=dev)
        ops.gemm(z, wop, bop, logits, N.EPI_BIAS)
        ctx.saved = (x_e, x_p, z, wop)
        ctx.params = (we, wp)
        ctx.dims = (B, T, U1, E, pred_out.shape[2], J, V, Vp, pre_project, dtype, enc_out.dtype, bo is not None)
        return logits.view(B, T, U1, Vp)[..., :V]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dlogits):
        x_e, x_p, z, wop = ctx.saved
        we, wp = ctx.params
        B, T, U1, E, P, J, V, Vp, pre_project, dtype, in_dtype, has_bo = ctx.dims
        dev = z.device
        rows = B * T * U1
        # the gradient arrives as a (B,T,U1,V) view of a zero-padded (rows, Vp) buffer when it comes from rnnt_loss below
        if (dlogits.dtype == dtype and dlogits.stride(-1) == 1 and dlogits.stride(2) == Vp and dlogits.stride(1) == U1 * Vp
                and dlogits.stride(0) == T * U1 * Vp and getattr(dlogits, "_cfm_padded", None) is not None):
            dl = dlogits._cfm_padded
        else:
            dl = torch.zeros((rows, Vp), dtype=dtype, device=dev)
            dl[:, :V] = dlogits.reshape(rows, V).to(dtype)
        dz = torch.empty((rows, J), dtype=dtype, device=dev)
        ops.gemm_ex(dl, wop.t(), dz)
        gwo = torch.zeros((Vp, J), dtype=torch.float32, device=dev)
        ops.gemm_ex(dl.t(), z.t(), gwo, accumulate=True)
        gbo = None
        if has_bo:
            gbo_p = torch.zeros(Vp, dtype=torch.float32, device=dev)
            TO.colsum(dl, gbo_p)
            gbo = gbo_p[:V]
        de = torch.empty((B * T, J), dtype=dtype, device=dev)
        dp = torch.empty((B * U1, J), dtype=dtype, device=dev)
        N.check(N.lib().cfm_joint_tanh_bwd(dz.data_ptr(), z.data_ptr(), de.data_ptr(), dp.data_ptr(), B, T, U1, J,
                                           ops._DT[dtype], ops._stream(z)))
        if not pre_project:
            return (de.view(B, T, J).to(in_dtype), dp.view(B, U1, J).to(in_dtype), None, None, None, None, gwo[:V], gbo,
                    None, None)
        wed, wpd = we.detach().to(dtype).contiguous(), wp.detach().to(dtype).contiguous()
        d_enc = torch.empty((B * T, E), dtype=dtype, device=dev)
        d_pred = torch.empty((B * U1, P), dtype=dtype, device=dev)
        ops.gemm_ex(de, wed.t(), d_enc)
        ops.gemm_ex(dp, wpd.t(), d_pred)
        gwe = torch.zeros((J, E), dtype=torch.float32, device=dev)
        gwp = torch.zeros((J, P), dtype=torch.float32, device=dev)
        ops.gemm_ex(de.t(), x_e.t(), gwe, accumulate=True)
        ops.gemm_ex(dp.t(), x_p.t(), gwp, accumulate=True)
        gbe, gbp = torch.zeros(J, device=dev), torch.zeros(J, device=dev)
        TO.colsum(de, gbe)
        TO.colsum(dp, gbp)
        return (d_enc.view(B, T, E).to(in_dtype), d_pred.view(B, U1, P).to(in_dtype), gwe, gbe, gwp, gbp, gwo[:V], gbo,
                None, None)


class TransducerJoint(nn.Module):
    """joint.py:4-38."""

    def __init__(self, vocab_size, enc_output_size, pred_output_size, join_dim):
        super().__init__()
        self.activatoin = nn.Tanh()                     # (sic: the reference's attribute name)
        self.enc_ffn = nn.Linear(enc_output_size, join_dim)
        self.pred_ffn = nn.Linear(pred_output_size, join_dim)
        self.ffn_out = nn.Linear(join_dim, vocab_size)
        self.compute_dtype = None

    def forward(self, enc_out, pred_out, pre_project=True):
        if not enc_out.is_cuda:
            raise RuntimeError("TransducerJoint: expected a CUDA tensor (the B200 kernels have no CPU path)")
        if enc_out.ndim == 4 or pred_out.ndim == 4:
            raise NotImplementedError("the native joint takes (B, T, E) and (B, U, P) (joint.py:27-31 unsqueezes them)")
        dt = engine.resolve_dtype(self)
        return _JointFunction.apply(enc_out, pred_out, self.enc_ffn.weight, self.enc_ffn.bias, self.pred_ffn.weight,
                                    self.pred_ffn.bias, self.ffn_out.weight, self.ffn_out.bias, dt, pre_project)


class _RnntLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, logit_lengths, target_lengths, blank, reduction):
        B, T, U1, V = logits.shape
        dev = logits.device
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()
        if not (logits.stride(-1) == 1 and logits.stride(1) == U1 * logits.stride(2) and logits.stride(0) == T * logits.stride(1)):
            logits = logits.contiguous()
        ld = logits.stride(2)
        tg = targets.to(device=dev, dtype=torch.int32).contiguous()
        tl = logit_lengths.to(device=dev, dtype=torch.int32).contiguous()
        ul = target_lengths.to(device=dev, dtype=torch.int32).contiguous()
        if tg.shape[1] < U1 - 1:
            raise RuntimeError("rnnt_loss: targets must have at least U = logits.size(2) - 1 columns")
        nll = torch.empty(B, dtype=torch.float32, device=dev)
        ws = torch.empty(int(N.lib().cfm_rnnt_loss_ws_bytes(B, T, U1)), dtype=torch.uint8, device=dev)
        ops.ensure_init(logits)
        N.check(N.lib().cfm_rnnt_loss_fwd(logits.data_ptr(), ld, B, T, U1, V, int(blank), tg.data_ptr(), tg.shape[1],
                                          tl.data_ptr(), ul.data_ptr(), nll.data_ptr(), ws.data_ptr(), ops._DT[logits.dtype],
                                          ops._stream(logits)))
        ctx.saved = (logits, tg, tl, ul, nll, ws)
        ctx.cfg = (B, T, U1, V, ld, int(blank), reduction)
        if reduction == "mean":
            return nll.mean()
        if reduction == "sum":
            return nll.sum()
        return nll

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dloss):
        logits, tg, tl, ul, nll, ws = ctx.saved
        ctx.saved = None
        B, T, U1, V, ld, blank, reduction = ctx.cfg
        if reduction == "none":
            raise NotImplementedError("rnnt_loss backward with reduction='none' (per-utterance upstream gradients)")
        scale = 1.0 / (B if reduction == "mean" else 1)      # the upstream gradient stays on the device (no host sync)
        rows = B * T * U1
        Vp = ld if ld >= V else V
        buf = torch.empty((rows, ld), dtype=logits.dtype, device=logits.device)
        N.check(N.lib().cfm_rnnt_loss_bwd(logits.data_ptr(), ld, B, T, U1, V, Vp, blank, tg.data_ptr(), tg.shape[1], tl.data_ptr(),
                                          ul.data_ptr(), nll.data_ptr(), ws.data_ptr(), scale, buf.data_ptr(),
                                          ops._DT[logits.dtype], ops._stream(logits)))
        buf.mul_(dloss.to(device=buf.device, dtype=buf.dtype))
        g = buf.view(B, T, U1, ld)[..., :V]
        g._cfm_padded = buf                     # lets the joint's backward use the zero-padded buffer without a copy
        return g, None, None, None, None, None


def rnnt_loss(logits, targets, logit_lengths, target_lengths, blank=-1, clamp=-1, reduction="mean", fused_log_softmax=True):
    """Same call as torchaudio.functional.rnnt_loss (model.py:106-111): logits (B, T, U+1, V), targets (B, U) int,
    lengths (B,).  Native forward / backward (fused log-softmax); ``clamp`` <= 0 and fused_log_softmax=True only."""
    if clamp > 0 or not fused_log_softmax:
        raise NotImplementedError("rnnt_loss: gradient clamping / unfused log-softmax are not implemented")
    if not logits.is_cuda:
        raise RuntimeError("rnnt_loss: expected a CUDA tensor (the B200 kernels have no CPU path)")
    if reduction not in ("mean", "sum", "none"):
        raise ValueError("reduction must be 'mean', 'sum' or 'none'")
    if blank < 0:
        blank = logits.shape[-1] + blank
    return _RnntLossFunction.apply(logits, targets, logit_lengths, target_lengths, blank, reduction)


def _apply_padding(inp, padding, pad_value):
    """predictor.py:5-11."""
    return padding * pad_value + inp * (1 - padding)


class RNNPredictor(nn.Module):
    """predictor.py:14-86 (same constructor, parameter names and forward / forward_step / init_state)."""

    def __init__(self, vocab_size, embed_size, output_size, hidden_size, embed_dropout, num_layers, bias=True, dropout=0.1):
        super().__init__()
        self.num_layers = num_layers
        self.hidden_size = hidden_size
        self.embed = nn.Embedding(vocab_size, embed_size)
        self.dropout = nn.Dropout(embed_dropout)
        self.rnn = nn.LSTM(input_size=embed_size, hidden_size=hidden_size, num_layers=num_layers, bias=bias, batch_first=True,
                           dropout=dropout)
        self.projection = nn.Linear(hidden_size, output_size)
        self.embed_size = embed_size

    def init_state(self, inputs):
        return [torch.zeros(self.num_layers, inputs.size(0), self.hidden_size, device=inputs.device) for _ in range(2)]

    def forward(self, inputs, states=None):
        embed = self.dropout(self.embed(inputs))
        if states is None:
            states = self.init_state(inputs)
            states = (states[0].to(embed.dtype), states[1].to(embed.dtype))
        outputs, states = self.rnn(embed, states)
        return self.projection(outputs)

    def forward_step(self, inputs, padding, cache):
        state_m, state_c = cache
        embed = self.dropout(self.embed(inputs))
        out, (m, c) = self.rnn(embed, (state_m, state_c))
        out = self.projection(out)
        m = _apply_padding(m, padding.unsqueeze(0), state_m)
        c = _apply_padding(c, padding.unsqueeze(0), state_c)
        return out, (m, c)


@torch.no_grad()
def basic_greedy_search(predictor, joint, encoder_out, encoder_out_lens, blank=0, n_steps=64, cache=None,
                        pred_input_step=None):
    """Frame-synchronous greedy transducer decoding of ONE utterance (semantics of model.py:221-269): at every encoder
    frame the joint is evaluated on (frame, current predictor output); a non-blank argmax is emitted, fed back through
    the predictor (whose LSTM state advances only then) and the same frame is tried again, up to ``n_steps`` symbols per
    frame; a blank moves on to the next frame.  encoder_out (1, T, E).  Returns (hyps, (last symbol, predictor state))
    so that a streaming caller can continue from where it stopped."""
    dev = encoder_out.device
    no_padding = torch.zeros(1, 1, device=dev)
    symbol = (torch.tensor([[blank]], device=dev) if pred_input_step is None else pred_input_step.to(dev))
    state = predictor.init_state(symbol) if cache is None else (cache[0].to(dev), cache[1].to(dev))
    hyps = []
    pred_out, next_state = None, None
    refresh = True                       # the predictor output is stale (start, or a symbol was just emitted)
    n_frames = int(encoder_out_lens)
    for t in range(n_frames):
        frame = encoder_out[:, t:t + 1, :]
        emitted = 0
        while True:
            if refresh:
                pred_out, next_state = predictor.forward_step(symbol, no_padding, state)
            logits = joint(frame, pred_out)                                   # (1, 1, 1, V): native joint step
            best = int(logits.float().log_softmax(dim=-1).argmax(dim=-1).squeeze())
            if best == blank:
                refresh = False
                break
            hyps.append(best)
            symbol = torch.tensor([[best]], device=dev)
            state = next_state
            refresh = True
            emitted += 1
            if emitted >= n_steps:
                break
    return hyps, (symbol, state)
