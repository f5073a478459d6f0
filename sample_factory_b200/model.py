"""Device-resident actor-critic parameters for the hot path.

Mirrors the reference's ActorCriticSharedWeights with MlpEncoder -> ModelCoreIdentity -> MlpDecoder -> critic_linear /
distribution_linear (model/actor_critic.py:136-195, encoder.py:72-91, core.py:67-77, decoder.py:15-35) as DATA: one
flat fp32 parameter buffer in HBM (plus flat grad / Adam-moment buffers of the same shape) so that grad-norm, Adam and
the NCCL all-reduce are single launches over contiguous memory.  Tensor names and order are the reference's
state_dict keys, so checkpoints round-trip (learner.py:323-332).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

OBS_NORM_PREFIX = "obs_normalizer.running_mean_std.running_mean_std.obs."
RET_NORM_PREFIX = "returns_normalizer."


@dataclass
class ModelSpec:
    obs_dim: int
    num_actions: int  # Discrete(n)
    encoder_mlp_layers: List[int] = field(default_factory=lambda: [512, 512])
    decoder_mlp_layers: List[int] = field(default_factory=list)
    nonlinearity: str = "elu"
    normalize_input: bool = True
    normalize_returns: bool = True
    obs_subtract_mean: float = 0.0
    obs_scale: float = 1.0
    use_rnn: bool = False        # model/core.py: ModelCoreRNN (one layer) between encoder and decoder
    rnn_type: str = "gru"
    rnn_size: int = 512
    # Box action space (action_distributions.py:290-323): num_actions is then the action DIMENSION
    continuous: bool = False
    adaptive_stddev: bool = True        # cfg.py:577; False -> one learned log-stddev vector (action_parameterization.py:42)
    continuous_tanh_scale: float = 0.0  # cfg.py:583
    initial_stddev: float = 1.0         # cfg.py:591
    # Tuple(Discrete(n_0), ..., Discrete(n_{K-1})) action space (action_distributions.py:197-286): the heads' sizes;
    # num_actions is then sum(n_k) (the number of logits)
    action_segments: Optional[List[int]] = None
    # image observations: obs_shape = (C, H, W) selects the ConvEncoder (model/encoder.py:88-145); obs_dim = C*H*W and the
    # fully connected layers after the conv head (encoder_conv_mlp_layers) take the place of encoder_mlp_layers
    obs_shape: Optional[Tuple[int, int, int]] = None
    encoder_conv_architecture: str = "convnet_atari"
    encoder_conv_mlp_layers: List[int] = field(default_factory=lambda: [512])
    obs_uint8: bool = False              # dtype of the observation rows in the trajectory buffers
    # False -> ActorCriticSeparateWeights (model/actor_critic.py:198-322): an actor tower (encoder + decoder MLP) feeding
    # distribution_linear and a critic tower feeding critic_linear; vector observations, no recurrent core on this path
    share_weights: bool = True

    CONV_ARCH = {  # model/encoder.py:127-134: (out_channels, kernel, stride); no padding
        "convnet_simple": [(32, 8, 4), (64, 4, 2), (128, 3, 2)],
        "convnet_impala": [(16, 8, 4), (32, 4, 2)],
        "convnet_atari": [(32, 8, 4), (64, 4, 2), (64, 3, 1)],
    }

    @property
    def conv_layers(self) -> List[Tuple[int, int, int, int, int, int, int, int]]:
        """[(C_in, H_in, W_in, C_out, kernel, stride, H_out, W_out)] of the conv head ([] for vector observations)"""
        if self.obs_shape is None:
            return []
        c, h, w = self.obs_shape
        out = []
        for (co, k, s) in self.CONV_ARCH[self.encoder_conv_architecture]:
            ho, wo = (h - k) // s + 1, (w - k) // s + 1
            out.append((c, h, w, co, k, s, ho, wo))
            c, h, w = co, ho, wo
        return out

    @property
    def conv_out_size(self) -> int:
        c, _, _, co, _, _, ho, wo = self.conv_layers[-1]
        return co * ho * wo

    @property
    def fc_encoder_layers(self) -> List[int]:
        """widths of the fully connected encoder layers (after the conv head for image observations)"""
        return list(self.encoder_conv_mlp_layers) if self.obs_shape is not None else list(self.encoder_mlp_layers)

    @property
    def fc_encoder_input(self) -> int:
        return self.conv_out_size if self.obs_shape is not None else self.obs_dim

    def fc_encoder_name(self, i: int, what: str) -> str:
        if self.obs_shape is not None:
            return f"encoder.encoders.obs.enc.mlp_layers.{2 * i}.{what}"
        return f"encoder.encoders.obs.mlp_head.{2 * i}.{what}"

    @classmethod
    def from_cfg(cls, cfg, env) -> "ModelSpec":
        """The model the reference would build for this cfg / env (model/actor_critic.py:136-158, create_actor_critic):
        Discrete(n) envs expose `num_actions = n`; Box(A) envs expose `continuous = True` and `num_actions = A`."""
        obs_shape = getattr(env, "obs_shape", None)   # (C, H, W) image observations -> ConvEncoder (encoder.py:218-227)
        return cls(env.obs_dim, env.num_actions, list(cfg.encoder_mlp_layers), list(cfg.decoder_mlp_layers),
                   cfg.nonlinearity, cfg.normalize_input, cfg.normalize_returns, cfg.obs_subtract_mean, cfg.obs_scale,
                   bool(cfg.use_rnn), cfg.rnn_type, cfg.rnn_size,
                   obs_shape=None if obs_shape is None else tuple(obs_shape),
                   encoder_conv_architecture=getattr(cfg, "encoder_conv_architecture", "convnet_atari"),
                   encoder_conv_mlp_layers=list(getattr(cfg, "encoder_conv_mlp_layers", [512])),
                   obs_uint8=bool(getattr(env, "obs_uint8", False)),
                   share_weights=bool(getattr(cfg, "actor_critic_share_weights", True)),
                   action_segments=(list(env.action_segments) if getattr(env, "action_segments", None) else None),
                   continuous=bool(getattr(env, "continuous", False)),
                   adaptive_stddev=bool(getattr(cfg, "adaptive_stddev", True)),
                   continuous_tanh_scale=float(getattr(cfg, "continuous_tanh_scale", 0.0)),
                   initial_stddev=float(getattr(cfg, "initial_stddev", 1.0)))

    @property
    def num_linear_action_outputs(self) -> int:
        """rows of distribution_linear"""
        if not self.continuous:
            return self.num_actions
        return 2 * self.num_actions if self.adaptive_stddev else self.num_actions

    @property
    def num_action_params(self) -> int:
        """calc_num_action_parameters (action_distributions.py:33-44): width of `action_logits`"""
        return 2 * self.num_actions if self.continuous else self.num_actions

    @property
    def action_width(self) -> int:
        """calc_num_actions (:16-30): width of `actions`"""
        if self.action_segments:
            return len(self.action_segments)
        return self.num_actions if self.continuous else 1

    @property
    def hidden(self) -> List[int]:
        # ModelCoreIdentity (use_rnn=False) passes the encoder output straight to the decoder MLP
        return self.fc_encoder_layers + list(self.decoder_mlp_layers)

    @property
    def rnn_state_size(self) -> int:
        """model/model_utils.py:11-24"""
        if not self.use_rnn:
            return 1 if self.share_weights else 2       # "actor and critic need separate states" (model_utils.py:20-22)
        return self.rnn_size * (2 if self.rnn_type == "lstm" else 1)

    @property
    def rnn_gates(self) -> int:
        return 4 if self.rnn_type == "lstm" else 3

    @property
    def tail_input_size(self) -> int:
        """width of the tensor that feeds critic_linear / distribution_linear"""
        if self.decoder_mlp_layers:
            return self.decoder_mlp_layers[-1]     # (separate weights: the width of ONE tower's tail)
        if self.use_rnn:
            return self.rnn_size
        return self.fc_encoder_layers[-1]

    def param_shapes(self) -> List[Tuple[str, Tuple[int, ...]]]:
        """(reference state_dict key, shape) in nn.Module.parameters() order."""
        out = []
        if not self.share_weights:
            # registration order of ActorCriticSeparateWeights.__init__ (actor_critic.py:208-225)
            assert self.obs_shape is None and not self.use_rnn, "separate actor / critic weights: MLP towers only"
            for tw in ("actor_", "critic_"):
                d = self.obs_dim
                for i, h in enumerate(self.encoder_mlp_layers):
                    out.append((f"{tw}encoder.encoders.obs.mlp_head.{2 * i}.weight", (h, d)))
                    out.append((f"{tw}encoder.encoders.obs.mlp_head.{2 * i}.bias", (h,)))
                    d = h
            d_enc = d
            for tw in ("actor_", "critic_"):
                d = d_enc
                for i, h in enumerate(self.decoder_mlp_layers):
                    out.append((f"{tw}decoder.mlp.{2 * i}.weight", (h, d)))
                    out.append((f"{tw}decoder.mlp.{2 * i}.bias", (h,)))
                    d = h
            out.append(("critic_linear.weight", (1, d)))
            out.append(("critic_linear.bias", (1,)))
            if self.continuous and not self.adaptive_stddev:
                out.append(("action_parameterization.learned_stddev", (self.num_actions,)))
            out.append(("action_parameterization.distribution_linear.weight", (self.num_linear_action_outputs, d)))
            out.append(("action_parameterization.distribution_linear.bias", (self.num_linear_action_outputs,)))
            return out
        for i, (ci, _h, _w, co, k, _s, _ho, _wo) in enumerate(self.conv_layers):
            out.append((f"encoder.encoders.obs.enc.conv_head.{2 * i}.weight", (co, ci, k, k)))
            out.append((f"encoder.encoders.obs.enc.conv_head.{2 * i}.bias", (co,)))
        d = self.fc_encoder_input
        for i, h in enumerate(self.fc_encoder_layers):
            out.append((self.fc_encoder_name(i, "weight"), (h, d)))
            out.append((self.fc_encoder_name(i, "bias"), (h,)))
            d = h
        if self.use_rnn:
            G, H = self.rnn_gates, self.rnn_size
            out += [("core.core.weight_ih_l0", (G * H, d)), ("core.core.weight_hh_l0", (G * H, H)),
                    ("core.core.bias_ih_l0", (G * H,)), ("core.core.bias_hh_l0", (G * H,))]
            d = H
        for i, h in enumerate(self.decoder_mlp_layers):
            out.append((f"decoder.mlp.{2 * i}.weight", (h, d)))
            out.append((f"decoder.mlp.{2 * i}.bias", (h,)))
            d = h
        out.append(("critic_linear.weight", (1, d)))
        out.append(("critic_linear.bias", (1,)))
        if self.continuous and not self.adaptive_stddev:
            # a module's own parameters precede its children's in nn.Module.parameters()
            out.append(("action_parameterization.learned_stddev", (self.num_actions,)))
        out.append(("action_parameterization.distribution_linear.weight", (self.num_linear_action_outputs, d)))
        out.append(("action_parameterization.distribution_linear.bias", (self.num_linear_action_outputs,)))
        return out


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


class PolicyModel:
    """Flat parameter storage + normalizer buffers on one device."""

    def __init__(self, spec: ModelSpec, device: torch.device, seed: int = 0, policy_init_gain: float = 1.0,
                 policy_initialization: str = "orthogonal"):
        assert policy_initialization in ("orthogonal", "xavier_uniform", "torch_default"), policy_initialization
        self.policy_initialization = policy_initialization
        self.spec = spec
        self.device = device
        shapes = spec.param_shapes()
        # every tensor starts on a 256-byte boundary inside the flat buffer (vector loads / TMA alignment)
        offsets, off = [], 0
        for _, shp in shapes:
            offsets.append(off)
            off += _align(math.prod(shp))
        self.numel_padded = off
        self.num_params = sum(math.prod(s) for _, s in shapes)
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.params: Dict[str, Tensor] = {}
        self.grads: Dict[str, Tensor] = {}
        self._slices: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        for (name, shp), o in zip(shapes, offsets):
            n = math.prod(shp)
            self.params[name] = self.flat[o : o + n].view(shp)
            self.grads[name] = self.grad[o : o + n].view(shp)
            self._slices[name] = (o, shp)
        self.names = [n for n, _ in shapes]

        # running_mean_std.py:45-47
        self.obs_mean = torch.zeros(spec.obs_dim, dtype=torch.float64, device=device)
        self.obs_var = torch.ones(spec.obs_dim, dtype=torch.float64, device=device)
        self.obs_count = torch.ones(1, dtype=torch.float64, device=device)
        self.ret_mean = torch.zeros(1, dtype=torch.float64, device=device)
        self.ret_var = torch.ones(1, dtype=torch.float64, device=device)
        self.ret_count = torch.ones(1, dtype=torch.float64, device=device)

        self._init_weights(seed, policy_init_gain)
        self._register_lo()
        self.refresh_cat_heads()

    # ---- tf32 low halves of the weights (3xTF32 engine: the weight operand's lo tile is loaded, not recomputed) -----
    def _register_lo(self) -> None:
        self.flat_lo = None
        if self.flat.is_cuda:
            from . import ops

            self.flat_lo = torch.empty_like(self.flat)
            ops.register_tf32_lo(self.flat, self.flat_lo)
        self._register_f16()

    # ---- fp16-split form of the 3-pass GEMM engine (include/sfb200.h): fp16 twins of the weights + activation bounds ----
    def _register_f16(self) -> None:
        """Plain MLP policies with normalised inputs: every hidden activation has a bound that follows from the weights
        (normalised observations are clipped to +-5, running_mean_std.py:96-110; |act(x W^T + b)| <= |x|_inf * max_n
        |W[n,:]|_1 + |b|_inf), so the forward GEMMs -- and dX, through transposed twins -- can take the fp16 operand path."""
        sp = self.spec
        self.f16_twins = None
        self.f16_T: Dict[str, Tensor] = {}
        self.bound_x = self.bound_h = None
        ok = (self.flat.is_cuda and sp.normalize_input and sp.obs_shape is None and not sp.use_rnn and sp.share_weights
              and not sp.decoder_mlp_layers and len(sp.hidden) >= 1)
        if not ok:
            return
        from . import ops

        self.f16_twins = torch.empty(2 * self.flat.numel(), dtype=torch.float16, device=self.device)
        ops.register_f16_twins(self.flat, self.f16_twins)
        self.bound_x = torch.full((1,), 5.0, dtype=torch.float32, device=self.device)
        self.bound_h = torch.zeros(4 * len(sp.hidden), dtype=torch.float32, device=self.device)   # [bound, scratch, counter, -] per layer
        self.refresh_bounds()

    def refresh_bounds(self) -> None:
        """bounds of the hidden activations from the current weights (one tiny kernel per layer) + the transposed twins"""
        if self.f16_twins is None:
            return
        from . import ops

        act = ops.ACT[self.spec.nonlinearity]
        inb = self.bound_x
        for i, (W, b) in enumerate(self.hidden_layers()[:-1]):       # (the last hidden layer feeds the heads, not a GEMM)
            ops.linear_out_bound(W, b, inb, self.bound_h[4 * i: 4 * i + 4], act)
            inb = self.bound_h[4 * i: 4 * i + 1]
        for name in self.f16_T:
            ops.refresh_f16_transposed(self.params[name])

    def enable_f16_transposed(self, name: str) -> None:
        """transposed fp16 twins of one weight matrix (the learner's dX = dz . W reads W along its other axis)"""
        if self.f16_twins is None or name in self.f16_T:
            return
        from . import ops

        W = self.params[name]
        self.f16_T[name] = torch.empty(2 * W.numel(), dtype=torch.float16, device=self.device)
        ops.register_f16_transposed(W, self.f16_T[name])

    def weights_changed(self) -> None:
        """Call after writing `flat` / `params[...]` by anything other than the Adam kernel (which keeps lo current)."""
        if self.flat_lo is not None:
            from . import ops

            ops.refresh_tf32_lo(self.flat)
            if self.f16_twins is not None:
                ops.refresh_f16_twins(self.flat)
                self.refresh_bounds()
        self.refresh_cat_heads()

    def rebind_grad(self, grad: Tensor) -> None:
        """Move the flat gradient buffer (data parallel: into the NVLink comm buffer the peers read, dist_utils.PeerComm)"""
        assert grad.shape == self.flat.shape and grad.dtype == torch.float32 and grad.is_contiguous()
        grad.zero_()
        self.grad = grad
        self.grads = {n: grad[o: o + math.prod(shp)].view(shp) for n, (o, shp) in self._slices.items()}

    def __del__(self):
        try:
            if getattr(self, "flat_lo", None) is not None:
                from . import ops

                ops.unregister_tf32_lo(self.flat)
                if getattr(self, "f16_twins", None) is not None:
                    for name in self.f16_T:
                        ops.unregister_f16_transposed(self.params[name])
                    ops.unregister_f16_twins(self.flat)
        except Exception:
            pass

    def _init_weights(self, seed: int, gain: float) -> None:
        """ActorCritic.initialize_weights (actor_critic.py:73-96): biases 0 in every mode; Linear / Conv2d weights orthogonal
        (gain), xavier_uniform (gain), or left at the PyTorch default U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch_default)."""
        mode = self.policy_initialization
        g = torch.Generator(device="cpu").manual_seed(seed)
        for name in self.names:
            p = self.params[name]
            if name == "action_parameterization.learned_stddev":
                p.fill_(math.log(self.spec.initial_stddev))   # action_parameterization.py:59-61
            elif name.startswith("core.core."):
                # RNNs keep the PyTorch default init U(-1/sqrt(H), 1/sqrt(H)) (actor_critic.py:83-88)
                k = 1.0 / math.sqrt(self.spec.rnn_size)
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * k)
            elif name.endswith(".bias"):
                p.zero_()
            else:
                w = torch.empty(p.shape, dtype=torch.float32)
                if mode == "orthogonal":
                    torch.nn.init.orthogonal_(w, gain=gain, generator=g)
                elif mode == "xavier_uniform":
                    torch.nn.init.xavier_uniform_(w, gain=gain, generator=g)
                else:   # nn.Linear / nn.Conv2d reset_parameters: kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                    bound = 1.0 / math.sqrt(math.prod(p.shape[1:]))
                    w.uniform_(-bound, bound, generator=g)
                p.copy_(w)

    # ---- inference-side weight snapshot (async_rl: the reference's inference workers hold their own copy of the
    # weights and refresh it when the learner publishes a new version, model_sharing.py:17-94, inference_worker.py:207-233)
    def inference_copy(self) -> "PolicyModel":
        twin = object.__new__(PolicyModel)
        twin.spec, twin.device = self.spec, self.device
        twin.numel_padded, twin.num_params, twin.names, twin._slices = self.numel_padded, self.num_params, self.names, self._slices
        twin.flat = self.flat.clone()
        twin.grad = twin.exp_avg = twin.exp_avg_sq = None
        twin.params = {n: twin.flat[o : o + math.prod(shp)].view(shp) for n, (o, shp) in self._slices.items()}
        twin.grads = {}
        for k in ("obs_mean", "obs_var", "obs_count", "ret_mean", "ret_var", "ret_count"):
            setattr(twin, k, getattr(self, k).clone())
        twin._register_lo()
        twin.refresh_cat_heads()
        return twin

    def copy_weights_from(self, other: "PolicyModel") -> None:
        """Refresh this snapshot from the learner's model (three D2D copies on the current stream)."""
        self.flat.copy_(other.flat)
        if self.flat_lo is not None and other.flat_lo is not None:
            self.flat_lo.copy_(other.flat_lo)
            if self.f16_twins is not None and other.f16_twins is not None:
                self.f16_twins.copy_(other.f16_twins)
                self.bound_h.copy_(other.bound_h)      # (scratch words are zero between launches)
            self.refresh_cat_heads()
        else:
            self.weights_changed()
        self.obs_mean.copy_(other.obs_mean)
        self.obs_var.copy_(other.obs_var)

    # ---- layer access ---------------------------------------------------------------------------------------
    def hidden_layers(self) -> List[Tuple[Tensor, Tensor]]:
        """[(W [out,in], b [out]), ...] for encoder then decoder MLP layers."""
        return self.encoder_layers() + self.decoder_layers()

    def encoder_layers(self, grads: bool = False) -> List[Tuple[Tensor, Tensor]]:
        """the fully connected encoder layers (MlpEncoder, or the layers after the conv head of a ConvEncoder)"""
        src = self.grads if grads else self.params
        sp = self.spec
        return [(src[sp.fc_encoder_name(i, "weight")], src[sp.fc_encoder_name(i, "bias")])
                for i in range(len(sp.fc_encoder_layers))]

    def tower_layers(self, tower: str, grads: bool = False) -> List[Tuple[Tensor, Tensor]]:
        """separate actor / critic weights: [(W, b)] of one tower ("actor_" / "critic_"), encoder then decoder MLP"""
        src = self.grads if grads else self.params
        out = [(src[f"{tower}encoder.encoders.obs.mlp_head.{2 * i}.weight"], src[f"{tower}encoder.encoders.obs.mlp_head.{2 * i}.bias"])
               for i in range(len(self.spec.encoder_mlp_layers))]
        out += [(src[f"{tower}decoder.mlp.{2 * i}.weight"], src[f"{tower}decoder.mlp.{2 * i}.bias"])
                for i in range(len(self.spec.decoder_mlp_layers))]
        return out

    def refresh_cat_heads(self) -> None:
        """separate weights: the heads kernels read ONE tail [M, 2H] = [actor tail | critic tail]; critic_linear and
        distribution_linear are embedded in zero-padded [., 2H] matrices (value <- critic half, logits <- actor half)."""
        if self.spec.share_weights:
            return
        H = self.spec.tail_input_size
        if not hasattr(self, "Wv_cat"):
            A = self.spec.num_linear_action_outputs
            self.Wv_cat = torch.zeros((1, 2 * H), dtype=torch.float32, device=self.device)
            self.Wa_cat = torch.zeros((A, 2 * H), dtype=torch.float32, device=self.device)
        self.Wv_cat[:, H:].copy_(self.params["critic_linear.weight"])
        self.Wa_cat[:, :H].copy_(self.params["action_parameterization.distribution_linear.weight"])

    def conv_params(self, grads: bool = False) -> List[Tuple[Tensor, Tensor]]:
        """[(W [C_out, C_in, k, k], b [C_out])] of the conv head"""
        src = self.grads if grads else self.params
        return [(src[f"encoder.encoders.obs.enc.conv_head.{2 * i}.weight"], src[f"encoder.encoders.obs.enc.conv_head.{2 * i}.bias"])
                for i in range(len(self.spec.conv_layers))]

    def decoder_layers(self, grads: bool = False) -> List[Tuple[Tensor, Tensor]]:
        src = self.grads if grads else self.params
        return [(src[f"decoder.mlp.{2 * i}.weight"], src[f"decoder.mlp.{2 * i}.bias"])
                for i in range(len(self.spec.decoder_mlp_layers))]

    def rnn_params(self, grads: bool = False) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """(W_ih [G*H, in], W_hh [G*H, H], b_ih, b_hh) of the one-layer GRU/LSTM core"""
        src = self.grads if grads else self.params
        return (src["core.core.weight_ih_l0"], src["core.core.weight_hh_l0"], src["core.core.bias_ih_l0"],
                src["core.core.bias_hh_l0"])

    def hidden_layer_grads(self) -> List[Tuple[Tensor, Tensor]]:
        return self.encoder_layers(grads=True) + self.decoder_layers(grads=True)

    @property
    def learned_log_std(self):
        """the learned log-stddev vector (None unless continuous with adaptive_stddev=False)"""
        return self.params.get("action_parameterization.learned_stddev")

    def dist_kwargs(self) -> Dict:
        """distribution description for the continuous heads / loss ops"""
        sp = self.spec
        return dict(act_dim=sp.num_actions, adaptive_stddev=sp.adaptive_stddev, learned_log_std=self.learned_log_std,
                    tanh_scale=sp.continuous_tanh_scale)

    @property
    def critic(self) -> Tuple[Tensor, Tensor]:
        return self.params["critic_linear.weight"], self.params["critic_linear.bias"]

    @property
    def actor(self) -> Tuple[Tensor, Tensor]:
        return (self.params["action_parameterization.distribution_linear.weight"],
                self.params["action_parameterization.distribution_linear.bias"])

    # ---- checkpoint compatibility (learner.py:323-332: 'model' entry) -------------------------------------------
    def state_dict(self) -> Dict[str, Tensor]:
        sd: Dict[str, Tensor] = {}
        if self.spec.normalize_input:
            shp = self.spec.obs_shape if self.spec.obs_shape is not None else (self.spec.obs_dim,)
            sd[OBS_NORM_PREFIX + "running_mean"] = self.obs_mean.clone().view(shp)   # reference shape = obs space shape
            sd[OBS_NORM_PREFIX + "running_var"] = self.obs_var.clone().view(shp)
            sd[OBS_NORM_PREFIX + "count"] = self.obs_count.clone()
        if self.spec.normalize_returns:
            sd[RET_NORM_PREFIX + "running_mean"] = self.ret_mean.clone()
            sd[RET_NORM_PREFIX + "running_var"] = self.ret_var.clone()
            sd[RET_NORM_PREFIX + "count"] = self.ret_count.clone()
        for n in self.names:
            sd[n] = self.params[n].clone()
        return sd

    def load_state_dict(self, sd: Dict[str, Tensor], strict: bool = True) -> None:
        known = set(self.names)
        for k, v in sd.items():
            v = torch.as_tensor(v)
            if k in known:
                self.params[k].copy_(v.to(self.device, torch.float32).view(self.params[k].shape))
            elif k == OBS_NORM_PREFIX + "running_mean":
                self.obs_mean.copy_(v.to(self.device).reshape(-1))     # image observations: [C,H,W] statistics, flat here
            elif k == OBS_NORM_PREFIX + "running_var":
                self.obs_var.copy_(v.to(self.device).reshape(-1))
            elif k == OBS_NORM_PREFIX + "count":
                self.obs_count.copy_(v.to(self.device).view(1))
            elif k == RET_NORM_PREFIX + "running_mean":
                self.ret_mean.copy_(v.to(self.device).view(1))
            elif k == RET_NORM_PREFIX + "running_var":
                self.ret_var.copy_(v.to(self.device).view(1))
            elif k == RET_NORM_PREFIX + "count":
                self.ret_count.copy_(v.to(self.device).view(1))
            elif strict:
                raise KeyError(f"unexpected key in state_dict: {k}")
        self.weights_changed()
        if strict:
            missing = known - set(sd.keys())
            if missing:
                raise KeyError(f"missing keys in state_dict: {sorted(missing)}")

    def optimizer_state_dict(self, step: int, lr: float, betas, eps: float) -> dict:
        """torch.optim.Adam.state_dict() layout so reference tooling can load it (learner.py:329)."""
        state = {}
        for i, n in enumerate(self.names):
            o, shp = self._slices[n]
            k = math.prod(shp)
            state[i] = dict(step=torch.tensor(float(step)), exp_avg=self.exp_avg[o : o + k].view(shp).clone(),
                            exp_avg_sq=self.exp_avg_sq[o : o + k].view(shp).clone())
        group = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                     capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False,
                     params=list(range(len(self.names))))
        return dict(state=state, param_groups=[group])

    def load_optimizer_state_dict(self, osd: dict) -> int:
        step = 0
        for i, n in enumerate(self.names):
            if i not in osd["state"]:
                continue
            st = osd["state"][i]
            o, shp = self._slices[n]
            k = math.prod(shp)
            self.exp_avg[o : o + k].copy_(st["exp_avg"].to(self.device).reshape(-1))
            self.exp_avg_sq[o : o + k].copy_(st["exp_avg_sq"].to(self.device).reshape(-1))
            step = int(float(st["step"]))
        return step
