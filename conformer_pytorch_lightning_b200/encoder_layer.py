"""Drop-in for the reference's src/encoder_layer.py (ConformerEncoderLayer, :9-71)."""
import torch
import torch.nn as nn

from . import engine
from .attention import MultiHeadSelfAttentionModule, RelativeMultiHeadSelfAttentionModule
from .convolution import ConvolutionModule
from .feedforward import PositionwiseFeedForwardModule


class ConformerEncoderLayer(nn.Module):
    """Macaron block  x += FFN/2 ; x += MHSA ; x += Conv ; x += FFN/2 ; LN  executed as the fused
    native kernel chain of ``engine.run_layers``.  Sub-module registration order and names equal the
    reference's, hence an identical state_dict."""

    def __init__(self, encoder_dim, kernel_size, feedforward_dropout, attention_dropout, hidden_dim, num_heads,
                 use_relative):
        super().__init__()
        self.feed_forward = PositionwiseFeedForwardModule(encoder_dim, feedforward_dropout, hidden_dim)
        attn_cls = RelativeMultiHeadSelfAttentionModule if use_relative else MultiHeadSelfAttentionModule
        self.self_attn = attn_cls(encoder_dim, num_heads, attention_dropout)
        self.conv_module = ConvolutionModule(encoder_dim, kernel_size, hidden_dim)
        self.feed_forward_macaron = PositionwiseFeedForwardModule(encoder_dim, feedforward_dropout, hidden_dim)
        self.norm_ff = nn.LayerNorm(encoder_dim, eps=1e-5)
        self.norm_ff_macaron = nn.LayerNorm(encoder_dim, eps=1e-5)
        self.norm_mha = nn.LayerNorm(encoder_dim, eps=1e-5)
        self.norm_conv = nn.LayerNorm(encoder_dim, eps=1e-5)
        self.norm_final = nn.LayerNorm(encoder_dim, eps=1e-5)
        self.dropout = nn.Dropout(feedforward_dropout)

    def derived_weights(self, dtype):
        f = engine._f32
        return {"ffm": self.feed_forward_macaron.derived_weights(dtype),
                "ff": self.feed_forward.derived_weights(dtype),
                "mha": self.self_attn.derived_weights(dtype),
                "conv": self.conv_module.derived_weights(dtype),
                "ffm_g": f(self.norm_ff_macaron.weight), "ffm_b": f(self.norm_ff_macaron.bias),
                "mha_g": f(self.norm_mha.weight), "mha_b": f(self.norm_mha.bias),
                "conv_g": f(self.norm_conv.weight), "conv_b": f(self.norm_conv.bias),
                "ff_g": f(self.norm_ff.weight), "ff_b": f(self.norm_ff.bias),
                "fin_g": f(self.norm_final.weight), "fin_b": f(self.norm_final.bias)}

    def _max_dropout(self):
        return max(self.dropout.p, self.self_attn.dropout.p, self.feed_forward.dropout.p, self.feed_forward_macaron.dropout.p)

    def derived_generation(self):
        return (self.feed_forward_macaron._derived.generation + self.feed_forward._derived.generation +
                self.self_attn._derived.generation + self.conv_module._derived.generation)

    def forward(self, inputs, inputs_attn_mask, pos_embed,
                inputs_pad_mask=torch.ones((0, 0, 0), dtype=torch.bool),
                attn_cache=torch.ones((0, 0, 0), dtype=torch.bool),
                cnn_cache=torch.ones((0, 0, 0), dtype=torch.bool)):
        dtype = engine.resolve_dtype(self)
        if engine.wants_autograd(self, inputs) or (self.training and self._max_dropout() > 0.0):
            # differentiable / dropout-capable schedule (training.py); the batched forward only (no streaming cache)
            if attn_cache is not None and attn_cache.dim() == 4 and attn_cache.size(0) > 0:
                raise NotImplementedError("training with a streaming attention cache is not implemented")
            from . import training
            out = training.run_stack(inputs.float(), [self], None, inputs_attn_mask, pos_embed, inputs_pad_mask, dtype)
            B, T, d = inputs.shape
            new_cnn_cache = torch.zeros((0, 0, 0), dtype=inputs.dtype, device=inputs.device)
            new_attn_cache = torch.zeros((0, 0, 0, 0), dtype=inputs.dtype, device=inputs.device)
            return out.to(inputs.dtype), inputs_attn_mask, new_attn_cache, new_cnn_cache
        out, caches = engine.run_layers(inputs.float(), [self], None, inputs_attn_mask, pos_embed, inputs_pad_mask,
                                        [attn_cache], True, dtype)
        new_cnn_cache = torch.zeros((0, 0, 0), dtype=inputs.dtype, device=inputs.device)
        return out.to(inputs.dtype), inputs_attn_mask, caches[0].to(inputs.dtype), new_cnn_cache
