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
        if activation != 'swish':
            raise NotImplementedError("only the 'swish' activation (the one the reference encoder uses) has a native epilogue")
        self.activation = nn.SiLU()
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
        engine.ffn_into(x, y, self.derived_weights(dtype), 1.0, engine.thread_workspace())
        return x.view(shape).to(inputs.dtype)
