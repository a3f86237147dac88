"""Optimizer step of the training loop: a drop-in for ``torch.optim.Adam(params, lr)`` as the reference configures it
(``src/module.py:140-143``: default betas / eps, no weight decay, no amsgrad; the WarmupLR scheduler drives
``param_groups[i]['lr']``), running on ``cfm_adam_step``.

PyTorch's fused Adam walks ~490 parameter tensors in multi-tensor chunks (1.0 ms of the 11 ms C5 step for 33 M
parameters, an HBM-bound update that needs 0.15 ms).  Here the parameters live in a few flat fp32 buffers whose layout
FOLLOWS THE GRADIENTS: the native backward (training.py) leaves each layer's gradients as views of one flat bucket, so at
the first step the parameters whose gradients share a storage are re-pointed (``p.data``) to views of a flat buffer with
the same relative offsets, and from then on one kernel launch updates a whole bucket.  Parameters whose gradients are
stand-alone tensors (front-end, CTC head, aliased or re-laid-out views) are updated by one launch each.  The layout is
re-checked every step (pointer arithmetic only); a bucket whose gradients moved falls back to per-parameter launches.
``state_dict()`` has torch.optim.Adam's format (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter), so checkpoints move
between the two."""
import torch

from . import _native as N

__all__ = ["FlatAdam"]


def _bump_versions(tensors):
    # the kernel writes through raw pointers: bump the autograd version counters so that derived-weight caches
    # (engine.Derived keys on (data_ptr, _version)) and saved-tensor checks see the update
    inc = getattr(torch.autograd.graph, "increment_version", None)
    if inc is not None:
        try:
            inc(tensors)
            return
        except TypeError:
            for t in tensors:
                inc(t)
            return
    for t in tensors:
        torch._C._increment_version(t)


class _Segment:
    """A run of parameters updated by one launch: flat views of p / exp_avg / exp_avg_sq and the expected gradient layout."""
    __slots__ = ("params", "rel", "p", "m", "v", "span", "mirror")

    def __init__(self, params, rel, p, m, v, span, mirror=None):
        self.params, self.rel, self.p, self.m, self.v, self.span, self.mirror = params, rel, p, m, v, span, mirror


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, bf16_mirror=True):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("FlatAdam: bad hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        self.grad_scale = float(grad_scale)
        # bf16_mirror: the kernel also writes the bf16 copy of every updated parameter into a flat mirror (laid out like the
        # fp32 buffer); the bf16 compute path takes its weights from there (engine._actp) instead of re-casting the fp32
        # masters at every step (~0.4 ms of cast / copy kernels per C5 step)
        self.bf16_mirror = bool(bf16_mirror)
        self._segments = {}          # id(group) -> list of _Segment (built at the first step from the gradient layout)
        self._steps = {}             # id(group) -> python int (the kernel takes the step count by value)
        self._step_tensor = {}

    def add_param_group(self, param_group):
        super().add_param_group(param_group)          # (also called by Optimizer.__init__ for the initial groups)
        for p in self.param_groups[-1]["params"]:
            if not p.is_cuda or p.dtype != torch.float32:
                raise NotImplementedError("FlatAdam: CUDA fp32 parameters only (no CPU fallback)")

    # ---------------------------------------------------------------- layout
    @torch.no_grad()
    def _build(self, group):
        ps = [p for p in group["params"] if p.grad is not None]
        by_storage = {}
        for p in ps:
            g = p.grad
            if not (g.is_cuda and g.dtype == torch.float32):
                raise NotImplementedError("FlatAdam: gradients must be CUDA fp32")
            key = g.untyped_storage().data_ptr() if g.is_contiguous() else ("solo", id(p))
            by_storage.setdefault(key, []).append(p)
        segs = []
        for key, members in by_storage.items():
            runs, solo = [], []
            if isinstance(key, tuple) or len(members) == 1:
                solo = members
            else:
                members.sort(key=lambda p: p.grad.data_ptr())
                end = -1
                for p in members:                      # gradients that alias an earlier one cannot share its flat slot
                    a = p.grad.data_ptr()
                    if a >= end:
                        runs.append(p)
                        end = a + 4 * p.numel()
                    else:
                        solo.append(p)
                if len(runs) == 1:
                    solo += runs
                    runs = []
            if runs:
                base = runs[0].grad.data_ptr()
                rel = [(p.grad.data_ptr() - base) // 4 for p in runs]
                span = rel[-1] + runs[-1].numel()
                segs.append(self._make_segment(runs, rel, span, base % 16))
            for p in solo:
                segs.append(self._make_segment([p], [0], p.numel(), p.grad.data_ptr() % 16 if p.grad.is_contiguous() else 0))
        self._segments[id(group)] = segs
        self._steps.setdefault(id(group), 0)

    def _make_segment(self, params, rel, span, misalign):
        dev = params[0].device
        pad = (misalign // 4) % 4                      # same 16-byte phase as the gradients: the kernel vectorises both
        def flat():
            return torch.zeros(span + 4, dtype=torch.float32, device=dev)[pad:pad + span]
        fp, fm, fv = flat(), flat(), flat()
        mirror = torch.zeros(span + 4, dtype=torch.bfloat16, device=dev)[pad:pad + span] if self.bf16_mirror else None
        for p, r in zip(params, rel):
            n = p.numel()
            view = fp[r:r + n].view(p.shape)
            view.copy_(p.data)
            p.data = view                              # the parameter now lives in the flat buffer
            st = self.state[p]
            old_m, old_v = st.get("exp_avg"), st.get("exp_avg_sq")
            st["exp_avg"] = fm[r:r + n].view(p.shape)
            st["exp_avg_sq"] = fv[r:r + n].view(p.shape)
            if old_m is not None:                      # resumed from a checkpoint
                st["exp_avg"].copy_(old_m)
                st["exp_avg_sq"].copy_(old_v)
            st.setdefault("step", torch.zeros((), dtype=torch.float32))
            if mirror is not None:
                p._cfm_mirror = mirror[r:r + n].view(p.shape)
                p._cfm_mirror.copy_(p.data)
                p._cfm_mirror_version = -1             # becomes current with the first update written by the kernel
        return _Segment(params, rel, fp, fm, fv, span, mirror)

    # ---------------------------------------------------------------- step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = N.lib()
        for group in self.param_groups:
            gid = id(group)
            if gid not in self._segments:
                if not any(p.grad is not None for p in group["params"]):
                    continue
                self._build(group)
            elif any(p.grad is not None and "exp_avg" not in self.state[p] for p in group["params"]):
                self._build_late(group)
            if gid not in self._steps or self._steps[gid] == 0:
                # resume: torch.optim.Adam keeps the count per parameter
                counts = [int(self.state[p]["step"]) for p in group["params"] if "step" in self.state[p]]
                self._steps[gid] = max(counts) if counts else 0
            self._steps[gid] += 1
            step = self._steps[gid]
            lr, (b1, b2), eps = float(group["lr"]), group["betas"], float(group["eps"])
            touched = []
            for seg in self._segments[gid]:
                p0 = seg.params[0]
                g0 = p0.grad
                if g0 is None:
                    if any(p.grad is not None for p in seg.params):
                        self._step_params(lib, seg, group, step)
                        touched += seg.params
                    continue
                base = g0.data_ptr()
                ok = g0.is_contiguous() and all(
                    p.grad is not None and p.grad.data_ptr() == base + 4 * r and p.grad.is_contiguous()
                    for p, r in zip(seg.params, seg.rel))
                if ok:
                    stream = torch.cuda.current_stream(p0.device).cuda_stream
                    N.check(lib.cfm_adam_step(seg.p.data_ptr(), base, seg.m.data_ptr(), seg.v.data_ptr(),
                                              None if seg.mirror is None else seg.mirror.data_ptr(), seg.span, lr, b1, b2,
                                              eps, step, self.grad_scale, stream))
                else:
                    self._step_params(lib, seg, group, step)
                touched += seg.params
            shared = self._step_tensor.get(gid)
            if shared is None:
                shared = self._step_tensor[gid] = torch.zeros((), dtype=torch.float32)
            shared.fill_(step)
            for p in touched:
                if self.state[p].get("step") is not shared:
                    self.state[p]["step"] = shared       # one host tensor per group (state_dict still lists it per parameter)
            _bump_versions(touched)
            if self.bf16_mirror:
                for p in touched:                        # the mirrors hold exactly these parameter versions
                    p._cfm_mirror_version = p._version
        return loss

    def _step_params(self, lib, seg, group, step):
        """Per-parameter launches: the gradients of this segment are not laid out like its flat buffer (any more)."""
        lr, (b1, b2), eps = float(group["lr"]), group["betas"], float(group["eps"])
        for p in seg.params:
            if p.grad is None:
                continue
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            st = self.state[p]
            stream = torch.cuda.current_stream(p.device).cuda_stream
            mir = getattr(p, "_cfm_mirror", None) if self.bf16_mirror else None
            N.check(lib.cfm_adam_step(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                      None if mir is None else mir.data_ptr(), p.numel(), lr, b1, b2, eps, step, self.grad_scale,
                                      stream))

    @torch.no_grad()
    def _build_late(self, group):
        """Parameters that received their first gradient after the layout was built: one segment each."""
        for p in group["params"]:
            if p.grad is not None and "exp_avg" not in self.state[p]:
                self._segments[id(group)].append(self._make_segment([p], [0], p.numel(), 0))

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # torch's load_state_dict keeps references to the caller's tensors when dtype and device already match: take private
        # copies, they are moved into the flat buffers when the layout is (re)built at the next step
        for st in self.state.values():
            for k in ("exp_avg", "exp_avg_sq", "step"):
                if k in st and torch.is_tensor(st[k]):
                    st[k] = st[k].clone()
        self._segments.clear()
        self._steps.clear()
        self._step_tensor.clear()
