"""Drop-in for the reference's src/encoder.py (ConformerEncoder, :9-153)."""
import collections
import operator
import os

import torch
import torch.nn as nn

from . import _native, engine
from .attention import PositionalEncoding, RelativePositionalEncoding
from .convolution import ConvolutionSubSampling
from .encoder_layer import ConformerEncoderLayer
from .utils import make_attn_mask, make_pad_mask


_VERSION_OF = operator.attrgetter("_version")


class ConformerEncoder(nn.Module):
    """Same constructor arguments, forward / forward_chunk / forward_chunk_by_chunk signatures,
    return tuples and state_dict layout as the reference; the layer stack + after_norm
    (encoder.py:72-74) run as native sm_100a kernels.  CMVN, Conv2d sub-sampling and mask
    construction stay in PyTorch (outside the measured path)."""

    def __init__(self, input_dim, kernel_size, encoder_dim, dropout, attention_dropout, pos_enc_dropout,
                 hidden_dim, num_heads, encoder_num_layers, cmvn=None, max_len=5000, use_relative=False,
                 use_dynamic_chunk_size=False, use_dynamic_left_chunk=False, static_chunk_size=-1):
        super().__init__()
        pos_cls = RelativePositionalEncoding if use_relative else PositionalEncoding
        self.position_encoding = pos_cls(encoder_dim, pos_enc_dropout, max_len)
        self.embed = ConvolutionSubSampling(input_dim=input_dim, output_dim=encoder_dim,
                                            pos_enc=self.position_encoding)
        self.encoders = nn.ModuleList([
            ConformerEncoderLayer(encoder_dim, kernel_size, dropout, attention_dropout, hidden_dim, num_heads,
                                  use_relative) for _ in range(encoder_num_layers)])
        self.encoder_dim = encoder_dim
        self.after_norm = nn.LayerNorm(encoder_dim, eps=1e-5)
        self.global_cmvn = cmvn
        self.use_dynamic_chunk_size = use_dynamic_chunk_size
        self.use_dynamic_left_chunk = use_dynamic_left_chunk
        self.static_chunk_size = static_chunk_size
        self.compute_dtype = None
        # dtype of the returned encoder states: None = the dtype of the sub-sampled features (fp32, like the reference's
        # fp32 module); torch.bfloat16 halves the device->host traffic of a bf16 serving pipeline
        self.output_dtype = None
        # CUDA-graph plans of the measured path, one per (batch, frames, dtype, mask layout); see _graph_layers
        self.use_cuda_graphs = os.environ.get("CFM_B200_CUDA_GRAPHS", "1") != "0"
        # least-recently-used plans are evicted beyond this many (a plan owns a graph + its static buffers; length-
        # bucketed traffic has many shapes)
        self.max_plans = int(os.environ.get("CFM_B200_MAX_PLANS", "16"))
        self._plans = collections.OrderedDict()

    # ------------------------------------------------------------------ B200-specific knobs
    def set_compute_dtype(self, dtype):
        """torch.float32 (default, exact-parity path) or torch.bfloat16 (tcgen05 path).  Parameters and
        the state_dict stay fp32 either way; ``None`` follows torch.autocast."""
        if dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("compute dtype must be torch.float32, torch.bfloat16 or None")
        for m in self.modules():
            m.compute_dtype = dtype
        return self

    def encode_layers(self, outputs, inputs_attn_mask, pos_embed, inputs_pad_mask):
        """The measured path: layer loop + after_norm (encoder.py:72-74)."""
        dtype = engine.resolve_dtype(self)
        if self._training_path(outputs):
            # differentiable / dropout-capable schedule: native forward that saves activations + native backward
            # (training.py); the per-layer gradient buckets go to self.grad_sync (ddp.GradSync) when one is attached
            from . import training
            return training.run_stack(outputs.float(), list(self.encoders), self.after_norm, inputs_attn_mask, pos_embed,
                                      inputs_pad_mask, dtype, grad_sync=getattr(self, "grad_sync", None), owner=self)
        batched_pos = pos_embed is None or pos_embed.numel() == outputs.size(0) * outputs.size(2)
        if (self.use_cuda_graphs and not self.training and outputs.is_cuda and batched_pos
                and not torch.cuda.is_current_stream_capturing()):
            return self._graph_layers(outputs, inputs_attn_mask, inputs_pad_mask, dtype)
        out, _ = engine.run_layers(outputs.float(), list(self.encoders), self.after_norm, inputs_attn_mask, pos_embed,
                                   inputs_pad_mask, None, False, dtype)
        return out

    def _graph_layers(self, outputs, attn_mask, pad_mask, dtype):
        """Replay the ~15 launches/layer of the layer stack as one CUDA graph.  The first call with a new
        (shape, dtype, mask layout) runs eagerly, the second captures, later ones replay.  Inputs are
        copied into the plan's static buffers; derived weights are refreshed in place before the replay so
        load_state_dict / optimizer steps are honoured; a fresh output tensor is returned."""
        attn_mask = engine._mask_u8(attn_mask)
        pad_u8 = engine._mask_u8(pad_mask)
        layers = list(self.encoders)
        key = (tuple(outputs.shape), dtype, outputs.device,
               None if attn_mask is None else tuple(attn_mask.shape),
               None if pad_u8 is None else tuple(pad_u8.shape), len(layers))
        plan = self._lookup_plan(key)
        if plan is None:
            out, _ = engine.run_layers(outputs.float(), layers, self.after_norm, attn_mask, None, pad_u8, None, False, dtype)
            return out
        gen = self._derived_generation()
        if plan["graph"] is not None and plan["gen"] != gen:
            plan["graph"] = None              # derived weights were re-allocated: the graph holds stale addresses
        if plan["graph"] is None:
            x_s = torch.empty(outputs.shape, dtype=torch.float32, device=outputs.device)
            am_s = None if attn_mask is None else torch.empty(attn_mask.shape, dtype=torch.bool, device=outputs.device)
            pm_s = None if pad_u8 is None else torch.empty(pad_u8.shape, dtype=torch.bool, device=outputs.device)
            tmp = {"x": x_s, "am": am_s, "pm": pm_s}
            self._fill(tmp, outputs, attn_mask, pad_u8)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                engine.run_layers(x_s, layers, self.after_norm, am_s, None, pm_s, None, False, dtype, inplace=True)
            cur.wait_stream(side)
            g = torch.cuda.CUDAGraph()
            n0 = _native.launch_count()
            with torch.cuda.graph(g):
                out_s, _ = engine.run_layers(x_s, layers, self.after_norm, am_s, None, pm_s, None, False, dtype, inplace=True)
            plan.update(graph=g, x=x_s, am=am_s, pm=pm_s, out=out_s, launches=_native.launch_count() - n0,
                        gen=self._derived_generation(), ver=None)
        ver = self._weights_version()
        if plan.get("ver") != ver:            # in-place refresh of stale derived weights (bf16 copies, folded BN)
            for layer in layers:
                layer.derived_weights(dtype)
            plan["ver"] = ver
            if self._derived_generation() != plan["gen"]:
                # the refresh had to re-allocate (cannot happen for same-shaped parameters): re-capture
                plan["graph"] = None
                return self._graph_layers(outputs, attn_mask, pad_mask, dtype)
        self._fill(plan, outputs, attn_mask, pad_u8)
        plan["graph"].replay()
        engine.GRAPH_REPLAYED_LAUNCHES[0] += plan["launches"]     # native kernels inside the replayed graph
        # the plan's output buffer is overwritten by the next replay: hand out a copy (already in the output dtype)
        return plan["out"].to(self.output_dtype, copy=True) if self.output_dtype is not None else plan["out"].clone()

    def _lookup_plan(self, key):
        """LRU lookup; a new key is registered (first call of a shape runs eagerly, the second captures)."""
        plan = self._plans.get(key)
        if plan is not None:
            self._plans.move_to_end(key)
            return plan
        while len(self._plans) >= max(1, self.max_plans):
            self._plans.popitem(last=False)
        self._plans[key] = {"graph": None, "gen": None}
        return None

    def _derived_generation(self):
        return sum(layer.derived_generation() for layer in self.encoders)

    def _weights_version(self):
        """Cheap staleness key: sum of the autograd version counters of every parameter/buffer of the layer
        stack (bumped by optimizer steps, load_state_dict, any in-place edit) + a counter bumped by
        train()/eval()/.to()/.cuda() (see train, _apply)."""
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = [p for m in (self.encoders, self.after_norm) for p in list(m.parameters()) + list(m.buffers())]
            self.__dict__["_plist"] = plist
        return sum(map(_VERSION_OF, plist)) + self.__dict__.get("_epoch", 0)

    def _bump(self):
        self.__dict__["_epoch"] = self.__dict__.get("_epoch", 0) + (1 << 40)
        self.__dict__["_plist"] = None
        self.__dict__["_rg_list"] = None

    def train(self, mode=True):
        self._bump()
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        self._bump()
        self.__dict__["_plans"] = collections.OrderedDict()          # device / dtype moves invalidate captured graphs
        return super()._apply(fn, *args, **kwargs)

    @staticmethod
    def _fill(plan, outputs, attn_mask, pad_u8):
        """Copy the call's inputs into the plan's static buffers (one eager op each: every eager op costs host time the
        GPU waits for).  Masks are always re-copied: a new mask tensor can re-use the address of the previous one."""
        plan["x"].copy_(outputs)
        if plan["am"] is not None:
            torch.ne(attn_mask, 0, out=plan["am"])
        if plan["pm"] is not None:
            torch.ne(pad_u8, 0, out=plan["pm"])

    def _max_dropout(self):
        return max([m.p for m in self.encoders.modules() if isinstance(m, nn.Dropout)] + [0.0])

    def _training_path(self, x):
        """Training schedule when the caller expects gradients (grad mode on and x or a layer parameter requires grad),
        or when dropout must be applied (train() mode with a non-zero dropout in the layer stack)."""
        if torch.is_grad_enabled():
            if x.requires_grad:
                return True
            plist = self.__dict__.get("_rg_list")
            if plist is None:
                plist = self.__dict__["_rg_list"] = [p for m in (self.encoders, self.after_norm) for p in m.parameters()]
            if any(p.requires_grad for p in plist):
                return True
        return self.training and self._max_dropout() > 0.0

    # ------------------------------------------------------------------ reference API
    def forward(self, inputs, input_lengths, decoding_chunk_size=0, num_decoding_chunk_size=-1):
        if self.global_cmvn is not None:
            inputs = self.global_cmvn(inputs)
        max_seq_len = inputs.size(1)
        inputs_pad_mask = ~make_pad_mask(input_lengths, max_seq_len).unsqueeze(1)
        outputs, pos_embed, inputs_pad_mask = self.embed(inputs, inputs_pad_mask)
        inputs_attn_mask = make_attn_mask(outputs, inputs_pad_mask, self.use_dynamic_chunk_size,
                                          self.use_dynamic_left_chunk, decoding_chunk_size, self.static_chunk_size,
                                          num_decoding_chunk_size)
        out = self.encode_layers(outputs, inputs_attn_mask, pos_embed, inputs_pad_mask)
        return out.to(self.output_dtype or outputs.dtype), inputs_pad_mask

    def forward_chunk(self, inputs, offset, required_cache_size, attn_cache, cnn_cache,
                      inputs_attn_mask=torch.ones((0, 0, 0))):
        attn_cache = attn_cache.to(inputs.device)
        cnn_cache = cnn_cache.to(inputs.device)
        tmp_masks = torch.ones(1, inputs.size(1), device=inputs.device, dtype=torch.bool).unsqueeze(1)
        if self.global_cmvn is not None:
            inputs = self.global_cmvn(inputs)
        outputs, pos_embed, _ = self.embed(inputs, tmp_masks, offset)
        num_layers, cache_size = attn_cache.size(0), attn_cache.size(2)
        chunk_size = outputs.size(1)
        attention_key_size = cache_size + chunk_size
        pos_embed = self.embed.position_encoding(offset=offset - cache_size, size=attention_key_size, like=outputs)
        if required_cache_size < 0:
            next_cache_start = 0
        elif required_cache_size == 0:
            next_cache_start = attention_key_size
        else:
            next_cache_start = max(attention_key_size - required_cache_size, 0)
        if self._training_path(outputs):
            raise NotImplementedError("forward_chunk is an inference API (encoder.py:78 is only called under no_grad / "
                                      "eval in the reference); call it under torch.no_grad() in eval() mode")
        dtype = engine.resolve_dtype(self)
        no_mask = inputs_attn_mask is None or inputs_attn_mask.dim() != 3 or inputs_attn_mask.size(2) == 0
        # graphs only for a bounded left context: with required_cache_size < 0 the cache grows every chunk, no shape
        # ever repeats within an utterance and a plan per chunk index would pile up graphs + private pools
        if (self.use_cuda_graphs and not self.training and outputs.is_cuda and no_mask and len(self.encoders) > 0
                and required_cache_size >= 0 and not torch.cuda.is_current_stream_capturing()):
            out, r_attn_cache = self._graph_chunk(outputs, pos_embed, attn_cache, next_cache_start, dtype)
            r_attn_cache = r_attn_cache.to(outputs.dtype)
        else:
            out, r_attn_cache = self._chunk_layers(outputs.float(), pos_embed, attn_cache, next_cache_start, dtype,
                                                   inputs_attn_mask)
            r_attn_cache = r_attn_cache.to(outputs.dtype)
        r_cnn_cache = torch.zeros((len(self.encoders), 0, 0, 0), dtype=outputs.dtype, device=outputs.device)
        return out.to(outputs.dtype), r_attn_cache, r_cnn_cache

    def _chunk_layers(self, x, pos_embed, attn_cache, next_cache_start, dtype, attn_mask=None):
        """Layer loop of one streaming step (encoder.py:109-118): per-layer cache slice in, trimmed caches out."""
        caches = [attn_cache[i:i + 1] for i in range(len(self.encoders))] if attn_cache.size(0) > 0 else None
        out, new_caches = engine.run_layers(x, list(self.encoders), self.after_norm, attn_mask, pos_embed, None, caches,
                                            True, dtype)
        return out, torch.cat([c[:, :, next_cache_start:, :] for c in new_caches], dim=0)

    def _graph_chunk(self, outputs, pos_embed, attn_cache, next_cache_start, dtype):
        """Streaming steps are launch-bound (B = 1, 16 rows: ~150 small launches): with a fixed number of left chunks the
        shapes repeat from chunk to chunk, so the layer loop of a step is captured into a CUDA graph per (chunk, cache
        size, trim point) exactly like the batched path (first call eager, second captures, later ones replay)."""
        key = ("chunk", tuple(outputs.shape), tuple(pos_embed.shape), tuple(attn_cache.shape), next_cache_start, dtype,
               outputs.device, len(self.encoders))
        plan = self._lookup_plan(key)
        if plan is None:
            return self._chunk_layers(outputs.float(), pos_embed, attn_cache, next_cache_start, dtype)
        if plan["graph"] is not None and plan["gen"] != self._derived_generation():
            plan["graph"] = None
        if plan["graph"] is None:
            x_s = torch.empty(outputs.shape, dtype=torch.float32, device=outputs.device)
            p_s = torch.empty_like(pos_embed)
            c_s = torch.empty(attn_cache.shape, dtype=attn_cache.dtype, device=outputs.device)
            x_s.copy_(outputs); p_s.copy_(pos_embed); c_s.copy_(attn_cache)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._chunk_layers(x_s, p_s, c_s, next_cache_start, dtype)
            cur.wait_stream(side)
            g = torch.cuda.CUDAGraph()
            n0 = _native.launch_count()
            with torch.cuda.graph(g):
                out_s, r_s = self._chunk_layers(x_s, p_s, c_s, next_cache_start, dtype)
            plan.update(graph=g, x=x_s, p=p_s, c=c_s, out=out_s, r=r_s, launches=_native.launch_count() - n0,
                        gen=self._derived_generation())
        ver = self._weights_version()
        if plan.get("ver") != ver:
            for layer in self.encoders:
                layer.derived_weights(dtype)
            plan["ver"] = ver
        plan["x"].copy_(outputs)
        plan["p"].copy_(pos_embed)
        if plan["c"].numel() > 0:
            plan["c"].copy_(attn_cache)
        plan["graph"].replay()
        engine.GRAPH_REPLAYED_LAUNCHES[0] += plan["launches"]
        return plan["out"].clone(), plan["r"].clone()

    def forward_chunk_by_chunk(self, inputs, decoding_chunk_size, num_decoding_left_chunks=-1):
        subsampling_rate, context = 4, 7
        stride = subsampling_rate * decoding_chunk_size
        decoding_window = (decoding_chunk_size - 1) * subsampling_rate + context
        num_frames = inputs.size(1)
        attn_cache = torch.zeros((0, 0, 0, 0), device=inputs.device)
        cnn_cache = torch.zeros((0, 0, 0, 0), device=inputs.device)
        outputs = []
        offset = 0
        required_cache_size = decoding_chunk_size * num_decoding_left_chunks
        for cur in range(0, num_frames - context + 1, stride):
            end = min(cur + decoding_window, num_frames)
            chunk_outputs, attn_cache, cnn_cache = self.forward_chunk(
                inputs=inputs[:, cur:end, :], offset=offset, required_cache_size=required_cache_size,
                attn_cache=attn_cache, cnn_cache=cnn_cache)
            outputs.append(chunk_outputs)
            offset += chunk_outputs.size(1)
        outputs = torch.cat(outputs, 1)
        masks = torch.ones((1, 1, outputs.size(1)))
        return outputs, masks
