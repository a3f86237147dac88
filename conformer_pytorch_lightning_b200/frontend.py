"""Feature front-end on device (scope row f1): Kaldi-compatible fbank + global CMVN.

``Fbank`` computes what the reference's data pipeline computes per utterance on the CPU with
``torchaudio.compliance.kaldi.fbank(waveform * (1 << 15), num_mel_bins=80, frame_length=25, frame_shift=10, dither=...,
energy_floor=0.0, sample_frequency=16000)`` (processor.py:185-191), for a whole padded batch of waveforms at once and
with the zero padding of ``pad_sequence`` (processor.py:302-304).  ``GlobalCMVN`` is the drop-in for the reference's
module of the same name (cmvn.py:5-33, statistics file parsed like utils.py:7-28).  Both run on native kernels
(include/cfm_b200.h, "Feature front-end"); CPU tensors raise.
"""
import json
import math

import torch
import torch.nn as nn

from . import _native as N
from . import ops

__all__ = ["Fbank", "GlobalCMVN", "load_cmvn"]


def load_cmvn(json_cmvn_file):
    """utils.py:7-28: json {mean_stat, var_stat, frame_num} -> (mean, istd) tensors."""
    with open(json_cmvn_file) as f:
        st = json.load(f)
    means, variance, count = list(st["mean_stat"]), list(st["var_stat"]), st["frame_num"]
    for i in range(len(means)):
        means[i] /= count
        variance[i] = variance[i] / count - means[i] * means[i]
        if variance[i] < 1.0e-20:
            variance[i] = 1.0e-20
        variance[i] = 1.0 / math.sqrt(variance[i])
    return torch.tensor(means), torch.tensor(variance)


class GlobalCMVN(nn.Module):
    """cmvn.py:5-33.  ``GlobalCMVN(path)`` like the reference, or ``GlobalCMVN.from_stats(mean, istd)``."""

    def __init__(self, cmvn_path, norm_var=True):
        super().__init__()
        mean, istd = load_cmvn(cmvn_path)
        assert mean.shape == istd.shape
        self.norm_var = norm_var
        self.register_buffer("mean", mean)
        self.register_buffer("istd", istd)

    @classmethod
    def from_stats(cls, mean, istd, norm_var=True):
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.norm_var = norm_var
        self.register_buffer("mean", torch.as_tensor(mean, dtype=torch.float32).clone())
        self.register_buffer("istd", torch.as_tensor(istd, dtype=torch.float32).clone())
        return self

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("GlobalCMVN: expected a CUDA tensor (the B200 kernels have no CPU path)")
        xin = x.float().contiguous()
        y = torch.empty_like(xin)
        ops.ensure_init(xin)
        mean = self.mean.float().contiguous()
        istd = self.istd.float().contiguous() if self.norm_var else None
        N.check(N.lib().cfm_cmvn(xin.data_ptr(), y.data_ptr(), mean.data_ptr(), 0 if istd is None else istd.data_ptr(),
                                 xin.numel(), xin.shape[-1], torch.cuda.current_stream(x.device).cuda_stream))
        return y.to(x.dtype)


class Fbank(nn.Module):
    """forward(waveforms (B, n_max) float in int16 range [= torchaudio waveform * (1 << 15)], n_samples (B,)) ->
    (feats (B, m_max, num_mel_bins) fp32 zero padded, frame counts (B,)).  With ``cmvn`` (a GlobalCMVN) the
    normalisation is fused into the last kernel, padding frames included (the reference's encoder applies CMVN to the
    zero-padded feature matrix).  dither is not applied (the parity setting; the reference's training config uses 0.1)."""

    def __init__(self, num_mel_bins=80, frame_length=25, frame_shift=10, sample_frequency=16000, cmvn=None):
        super().__init__()
        if (frame_length, frame_shift, sample_frequency) != (25, 10, 16000):
            raise NotImplementedError("the native fbank implements the reference's setting: 25 ms / 10 ms at 16 kHz")
        self.num_mel_bins, self.win, self.shift, self.n_fft = num_mel_bins, 400, 160, 512
        n = torch.arange(400, dtype=torch.float64)
        k = torch.arange(257, dtype=torch.float64)
        ang = 2.0 * math.pi * k[:, None] * n[None, :] / 512.0
        self.register_buffer("dft", torch.cat([torch.cos(ang), -torch.sin(ang)], 0).float().contiguous(), persistent=False)
        self.register_buffer("window", torch.hann_window(400, periodic=False, dtype=torch.float64).pow(0.85).float(),
                             persistent=False)
        mel = self._mel_banks(num_mel_bins)
        melp = torch.zeros(num_mel_bins, 264, dtype=torch.float32)
        melp[:, :257] = mel.float()
        self.register_buffer("mel", melp.contiguous(), persistent=False)
        self.cmvn = cmvn

    @staticmethod
    def _mel_banks(num_bins, n_fft=512, sample_rate=16000.0, low_freq=20.0):
        ms = lambda f: 1127.0 * torch.log(1.0 + f / 700.0)
        lo, hi = ms(torch.tensor(low_freq, dtype=torch.float64)), ms(torch.tensor(0.5 * sample_rate, dtype=torch.float64))
        delta = (hi - lo) / (num_bins + 1)
        b = torch.arange(num_bins, dtype=torch.float64)[:, None]
        left, center, right = lo + b * delta, lo + (b + 1.0) * delta, lo + (b + 2.0) * delta
        mel = ms((sample_rate / n_fft) * torch.arange(n_fft // 2, dtype=torch.float64))[None, :]
        bins = torch.clamp(torch.minimum((mel - left) / (center - left), (right - mel) / (right - center)), min=0.0)
        return torch.cat([bins, torch.zeros(num_bins, 1, dtype=torch.float64)], 1)

    def num_frames(self, n_samples):
        n = torch.as_tensor(n_samples)
        return torch.where(n < self.win, torch.zeros_like(n), 1 + (n - self.win) // self.shift)

    @torch.no_grad()
    def forward(self, waveforms, n_samples):
        if not waveforms.is_cuda:
            raise RuntimeError("Fbank: expected a CUDA tensor (the B200 kernels have no CPU path)")
        wave = waveforms.float().contiguous()
        B, n_max = wave.shape
        dev = wave.device
        ns = torch.as_tensor(n_samples).to(device=dev, dtype=torch.int32).contiguous()
        m_max = 0 if n_max < self.win else 1 + (n_max - self.win) // self.shift
        nmel = self.num_mel_bins
        out = torch.empty((B, m_max, nmel), dtype=torch.float32, device=dev)
        if B == 0 or m_max == 0:
            return out, self.num_frames(ns)
        ops.ensure_init(wave)
        st = torch.cuda.current_stream(dev).cuda_stream
        rows = B * m_max
        frames = torch.empty((rows, 400), dtype=torch.float32, device=dev)
        N.check(N.lib().cfm_fbank_frames(wave.data_ptr(), wave.stride(0), ns.data_ptr(), self.window.data_ptr(),
                                         frames.data_ptr(), B, m_max, 0.97, st))
        spec = torch.empty((rows, 520), dtype=torch.float32, device=dev)
        ops.gemm_ex(frames, self.dft, spec[:, :514])                                   # 512-point real DFT as a GEMM
        power = torch.empty((rows, 264), dtype=torch.float32, device=dev)
        N.check(N.lib().cfm_fbank_power(spec.data_ptr(), 520, power.data_ptr(), 264, rows, 257, st))
        mel = torch.empty((rows, nmel), dtype=torch.float32, device=dev)
        ops.gemm_ex(power, self.mel, mel)                                              # mel filterbank
        mean = istd = None
        if self.cmvn is not None:
            mean = self.cmvn.mean.float().contiguous()
            istd = self.cmvn.istd.float().contiguous() if self.cmvn.norm_var else None
        N.check(N.lib().cfm_fbank_log_cmvn(mel.data_ptr(), out.data_ptr(), ns.data_ptr(), 0 if mean is None else mean.data_ptr(),
                                           0 if istd is None else istd.data_ptr(), B, m_max, nmel, st))
        return out, self.num_frames(ns)
