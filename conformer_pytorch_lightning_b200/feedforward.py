"""Drop-in for the reference's src/feedforward.py (PositionwiseFeedForwardModule, :4-21)."""
import torch
import torch.nn as nn

from . import _native as N
from . import engine, ops


class PositionwiseFeedForwardModule(nn.Module):
    """w_2(dropout(act(w_1(x)))) executed as two native GEMMs with the activation fused into
    the first epilogue.  Same constructor and parameter names as the reference."""

    def __init__(self, input_dim, dropout, hidden_dim, activation='swish'):
        super().__init__()
        self.w_1 = nn.Linear(input_dim, hidden_dim)
        # feedforward.py:10-14: 'relu' -> nn.ReLU, anything else -> nn.SiLU.  The encoder layer always takes the default;
        # ReLU is served by the GEMM + ReLU epilogue chain (the fused FFN kernel is SiLU only)
        self.activation = nn.ReLU() if activation == 'relu' else nn.SiLU()
        self.dropout = nn.Dropout(dropout)
        self.w_2 = nn.Linear(hidden_dim, input_dim)
        self._derived = engine.Derived()

    def derived_weights(self, dtype):
        return self._derived.get(self, dtype, lambda dt: engine.ffn_weights(self, dt))

    def forward(self, inputs):
        engine.check_inference_only(self, self.dropout.p, inputs)
        dtype = engine.resolve_dtype(self)
        shape = inputs.shape
        y = inputs.reshape(-1, shape[-1]).to(dtype).contiguous()
        x = torch.zeros((y.shape[0], shape[-1]), dtype=torch.float32, device=y.device)
        W = self.derived_weights(dtype)
        if isinstance(self.activation, nn.ReLU):
            h = engine.thread_workspace().get("ffn_h", (y.shape[0], W["w1"].shape[0]), y.dtype, y.device)
            ops.gemm(y, W["w1"], W["b1"], h, N.EPI_BIAS_RELU)
            ops.gemm(h, W["w2"], W["b2"], x, N.EPI_RESIDUAL, residual=x, alpha=1.0)
        else:
            engine.ffn_into(x, y, W, 1.0, engine.thread_workspace())
        return x.view(shape).to(inputs.dtype)
