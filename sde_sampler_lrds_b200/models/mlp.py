"""Drift backbones with the reference's module names and parameter layout (sde_sampler/models/mlp.py:
Model 9-44, TimeEmbed 57-96, FourierMLP 99-143), so that reference ``state_dict``s load unchanged.

The time embedding depends on t only (identical for every particle of a step - SURVEY.md Appendix B.2): it is
evaluated once per grid time with plain torch ops and enters the kernel as the per-step bias row
(LRDS_STEP_BIAS1).  The per-particle part of FourierMLP.forward runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable

import torch
import torch.nn.functional as F
from torch import nn

from .. import _native as N


def _check_gelu(act):
    if not (isinstance(act, nn.GELU) and getattr(act, "approximate", "none") == "none"):
        raise NotImplementedError("the B200 kernels implement exact-erf GELU only (conf/model/base/fouriermlp.yaml:5-6)")


class Model(nn.Module):
    def __init__(self, dim: int, dim_out=None):
        super().__init__()
        self.dim = dim
        self.dim_in = dim + 1
        self.dim_out = dim_out or dim

    @staticmethod
    def init_linear(layer: nn.Linear, bias_init: Callable | None = None, weight_init: Callable | None = None):
        if weight_init:
            weight_init(layer.weight)
        if bias_init:
            code = getattr(bias_init, "__code__", None) or getattr(getattr(bias_init, "func", None), "__code__", None)
            if code is not None and "weight" in code.co_varnames:
                bias_init(layer.bias, weight=layer.weight)
            else:
                bias_init(layer.bias)


class TimeEmbed(Model):
    """sin/cos features at 64 fixed frequencies with a learnable phase, then an MLP."""

    def __init__(self, dim_out: int, activation: Callable, num_layers: int = 2, channels: int = 64,
                 last_bias_init: Callable | None = None, last_weight_init: Callable | None = None):
        super().__init__(dim=1, dim_out=dim_out)
        self.channels = channels
        self.activation = activation
        self.register_buffer("timestep_coeff", torch.linspace(start=0.1, end=100, steps=channels).unsqueeze(0),
                             persistent=False)
        self.timestep_phase = nn.Parameter(torch.randn(1, channels))
        self.hidden_layer = nn.ModuleList([nn.Linear(2 * channels, channels)])
        self.hidden_layer += [nn.Linear(channels, channels) for _ in range(num_layers - 2)]
        self.out_layer = nn.Linear(channels, self.dim_out)
        Model.init_linear(self.out_layer, bias_init=last_bias_init, weight_init=last_weight_init)

    def rows(self, t: torch.Tensor, device="cpu") -> torch.Tensor:
        """Embedding rows for a 1-D tensor of times, computed on ``device`` from detached parameters."""
        _check_gelu(self.activation)
        dev = torch.device(device)
        with torch.no_grad():
            t = t.detach().to(dev, torch.float32).reshape(-1, 1)
            arg = self.timestep_coeff.to(dev) * t + self.timestep_phase.detach().to(dev)
            e = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
            for layer in self.hidden_layer:
                e = F.gelu(F.linear(e, layer.weight.detach().to(dev), layer.bias.detach().to(dev)))
            return F.linear(e, self.out_layer.weight.detach().to(dev), self.out_layer.bias.detach().to(dev))

    def forward(self, t: torch.Tensor, *args) -> torch.Tensor:
        assert t.ndim in [0, 1, 2]
        if t.ndim == 2:
            assert t.shape[1] == 1
        return self.rows(t, device=t.device)


class FourierMLP(Model):
    """x -> W_o GELU(... W_1 GELU(W_x x + b_x + TimeEmbed(t)) ...)  with 64 channels."""

    def __init__(self, dim: int, activation: Callable, num_layers: int = 4, channels: int = 64,
                 last_bias_init: Callable | None = None, last_weight_init: Callable | None = None,
                 use_angle_encoding: bool = False, **kwargs):
        super().__init__(dim=dim, **kwargs)
        if use_angle_encoding:
            raise NotImplementedError("angle encoding is not used by any rollout solver config")
        if channels != N.CHANNELS:
            raise NotImplementedError(f"the B200 kernels are built for {N.CHANNELS} channels")
        if self.dim_out != dim:
            raise NotImplementedError("FourierMLP with dim_out != dim is not used by the rollout")
        self.channels = channels
        self.activation = activation
        self.input_embed = nn.Linear(self.dim, channels)
        self.timestep_embed = TimeEmbed(dim_out=channels, activation=activation, num_layers=2, channels=channels)
        self.hidden_layer = nn.ModuleList([nn.Linear(channels, channels) for _ in range(num_layers - 2)])
        self.out_layer = nn.Linear(channels, self.dim_out)
        Model.init_linear(self.out_layer, bias_init=last_bias_init, weight_init=last_weight_init)
        self._packed = {}

    def __getstate__(self):
        # copy.deepcopy (EMA AveragedModel, CMCD.update_prior) / pickling: the packed weight blocks hold ctypes structs
        # with device pointers; a copy re-packs its own
        state = self.__dict__.copy()
        state["_packed"] = {}
        return state

    # ---- packing for the kernels -------------------------------------------------------------------------------
    def _version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def lrds_mlp(self, device, precision: int = N.PRECISION_FP32_SIMT):
        """(N.Mlp, keepalive) with pre-transposed weights on ``device`` (cached until a parameter changes).
        For a tensor-core ``precision`` the block also carries the tcgen05 weight image (lrds_pack_mlp_tc)."""
        _check_gelu(self.activation)
        if precision != N.PRECISION_FP32_SIMT:
            return self._lrds_mlp_tc(torch.device(device), precision)
        key = (str(device), self._version())
        if self._packed.get("key") != key:
            d, Cc = self.dim, self.channels
            d_pad = ((d + 7) // 8) * 8
            f = lambda p: p.detach().to(device, torch.float32)  # noqa: E731
            w_in_t = f(self.input_embed.weight).t().contiguous()
            nh = len(self.hidden_layer)
            if nh:
                w_hid_t = torch.stack([f(l.weight).t().contiguous() for l in self.hidden_layer]).contiguous()
                b_hid = torch.stack([f(l.bias) for l in self.hidden_layer]).contiguous()
            else:
                w_hid_t = torch.zeros(1, Cc, Cc, device=device)
                b_hid = torch.zeros(1, Cc, device=device)
            w_out_t = torch.zeros(Cc, d_pad, device=device)
            w_out_t[:, :d] = f(self.out_layer.weight).t()
            b_out = torch.zeros(d_pad, device=device)
            b_out[:d] = f(self.out_layer.bias)
            m = N.Mlp()
            m.d, m.d_pad, m.num_hidden = d, d_pad, nh
            keep = (w_in_t, w_hid_t, b_hid, w_out_t, b_out)
            m.w_in_t, m.w_hid_t, m.b_hid, m.w_out_t, m.b_out = (t.data_ptr() for t in keep)
            self._packed = {"key": key, "mlp": m, "keep": keep}
        return self._packed["mlp"], self._packed["keep"]

    def _lrds_mlp_tc(self, device, precision: int):
        base, keep = self.lrds_mlp(device)
        key = ("tc", precision, str(device), self._version())
        hit = self._packed.get(key)
        if hit is None:
            for k in [k for k in self._packed if isinstance(k, tuple) and k[0] == "tc" and k != key and k[1] == precision]:
                del self._packed[k]  # stale images of this precision
            nbytes = N.lib().lrds_tc_image_bytes(self.dim, len(self.hidden_layer), precision)
            if nbytes < 0:
                N.check(int(nbytes))
            image = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                N.check(N.lib().lrds_pack_mlp_tc(C.byref(base), precision, N.ptr(image), N.stream_ptr(device)))
            m = N.Mlp()
            C.memmove(C.byref(m), C.byref(base), C.sizeof(N.Mlp))
            m.tc_image = image.data_ptr()
            hit = self._packed[key] = (m, (keep, image))
        return hit

    def bias_rows(self, taus: torch.Tensor, device="cpu") -> torch.Tensor:
        """[S][64] rows input_embed.bias + TimeEmbed(tau) (the reference adds embed_x + embed_t, mlp.py:139), evaluated
        on ``device`` (the host when a plan is built; the parameters' device when a training step refreshes it)."""
        return self.timestep_embed.rows(taus, device) + self.input_embed.bias.detach().to(device, torch.float32)

    def forward(self, t: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        from ..pack import ctrl_forward
        return ctrl_forward(self, t, x)
