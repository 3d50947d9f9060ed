"""Actor-critic model over an episodic TransformerXL memory -- B200-native drop-in for the reference's
``model.ActorCriticModel`` (model.py:10-166): same constructor, same ``forward(obs, memory,
memory_mask, memory_indices) -> (policies, value, new_memory)``, same ``state_dict`` keys and shapes,
same ``get_grad_norm()`` groups.

Storage is B200-first: every parameter is a view into ONE flat fp32 arena (and its gradient a view
into a twin arena), laid out by libtrxlppo (``trxl_layout_*``).  That gives the native trunk a single
base pointer, makes clip+AdamW one fused pass, and makes the multi-GPU gradient exchange a single
all-reduce of one buffer.  Everything after the CNN encoder (lin_hidden -> embedding -> blocks ->
heads) runs in two native calls (one cluster-per-sample launch at rollout batch sizes); the three conv layers run as
tcgen05 3xTF32 implicit GEMMs, forward and backward (csrc/tc_conv.cu), with cuDNN / im2col+SIMT GEMM kept for more than 4
input channels.
"""
import os

import numpy as np
import torch
from torch import nn
from torch.distributions import Categorical
from torch.nn import functional as F

import trxl_native as native
from transformer import Transformer


def _conv_features(obs, c1, c2, c3):
    h = F.relu(F.conv2d(obs, c1.weight, c1.bias, stride=4))
    h = F.relu(F.conv2d(h, c2.weight, c2.bias, stride=2))
    h = F.relu(F.conv2d(h, c3.weight, c3.bias, stride=1))
    return h.reshape(obs.shape[0], -1)


def _require_fp32_convolutions():
    """The engine's parity contract is fp32 (1e-4 against the CPU reference).  cuDNN would otherwise
    run the encoder's convolutions -- forward AND backward, which executes outside any local context
    manager -- in TF32, so the switch is process-wide."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True      # let cuDNN time its fp32 algorithms once per conv shape


class _TrunkFunction(torch.autograd.Function):
    """Autograd bridge for callers that build their own loss on ``model.forward`` outputs.  The
    trainer's fused step does not go through here."""

    @staticmethod
    def forward(ctx, model, feat, table, slots, ep_index, win_index, mask, pe_index, *params):
        n = feat.shape[0]
        ws = torch.empty(model._ws_floats(n), device=feat.device)
        logits, value, out_mem = model._alloc_outputs(n, feat.device)
        feat_c = feat.detach().contiguous()
        native.model_forward(model._cfg, model._arena, feat_c, table, slots, ep_index, win_index, mask, pe_index, None,
                             model._pe_table(), n, ws, logits, value, out_mem)
        ctx.model, ctx.ws, ctx.out_mem = model, ws, out_mem
        ctx.args = (feat_c, table, slots, ep_index, win_index, mask, pe_index)
        ctx.mark_non_differentiable(out_mem)
        return logits, value, out_mem

    @staticmethod
    def backward(ctx, dlogits, dvalue, _dmem):
        model = ctx.model
        feat, table, slots, ep_index, win_index, mask, pe_index = ctx.args
        n = feat.shape[0]
        grads = torch.zeros_like(model._arena)
        dlogits = torch.zeros((n, model._sum_actions), device=feat.device) if dlogits is None else dlogits.contiguous()
        dvalue = torch.zeros((n,), device=feat.device) if dvalue is None else dvalue.contiguous()
        dfeat = torch.empty_like(feat) if ctx.needs_input_grad[1] else None
        native.model_backward(model._cfg, model._arena, grads, feat, table, slots, ep_index, win_index, mask, pe_index, None,
                              model._pe_table(), n, ctx.ws, ctx.out_mem, dlogits, dvalue, dfeat)
        pgrads = tuple(grads[off:off + numel].view(shape) if name in model._trunk_names else None
                       for name, off, numel, shape in model._param_slices)
        return (None, dfeat, None, None, None, None, None, None) + pgrads


class _EncoderFunction(torch.autograd.Function):
    """Autograd bridge for the tensor-core CNN encoder (csrc/tc_conv.cu) so that ``ActorCriticModel.forward`` stays
    differentiable end to end like the reference's; the trainer's fused step calls encode_train / encode_backward."""

    @staticmethod
    def forward(ctx, model, obs, *conv_params):
        n = obs.shape[0]
        ctx.model, ctx.n, ctx.hw = model, n, tuple(obs.shape[-2:])
        return model.encode_train(obs.detach().contiguous(), None, n).clone()

    @staticmethod
    def backward(ctx, dfeat):
        model = ctx.model
        grads = torch.zeros_like(model._arena)
        native.conv_train_backward(model._cfg, grads, ctx.n, ctx.hw[0], ctx.hw[1], model._enc_ws(ctx.n, *ctx.hw)[0], dfeat.contiguous())
        pg = tuple(grads[off:off + numel].view(shape) for name, off, numel, shape in model._param_slices if name.startswith("conv"))
        return (None, None) + pg


class ActorCriticModel(nn.Module):
    def __init__(self, config, observation_space, action_space_shape, max_episode_length):
        super().__init__()
        self.hidden_size = config["hidden_layer_size"]
        self.memory_layer_size = config["transformer"]["embed_dim"]
        self.observation_space_shape = tuple(observation_space.shape)
        self.max_episode_length = max_episode_length
        self.action_space_shape = tuple(int(a) for a in action_space_shape)
        tcfg = config["transformer"]
        self._visual = len(self.observation_space_shape) > 1
        _require_fp32_convolutions()

        # ---- parameter holders with the reference's names and init recipes (model.py:27-69) ----
        if self._visual:
            self.conv1 = nn.Conv2d(self.observation_space_shape[0], 32, 8, 4)
            self.conv2 = nn.Conv2d(32, 64, 4, 2, 0)
            self.conv3 = nn.Conv2d(64, 64, 3, 1, 0)
            for conv in (self.conv1, self.conv2, self.conv3):
                nn.init.orthogonal_(conv.weight, np.sqrt(2))
            self.conv_out_size = self.get_conv_output(self.observation_space_shape)
            feat_dim = self.conv_out_size
        else:
            feat_dim = self.observation_space_shape[0]
        self.lin_hidden = nn.Linear(feat_dim, self.memory_layer_size)
        nn.init.orthogonal_(self.lin_hidden.weight, np.sqrt(2))
        self.transformer = Transformer(tcfg, self.memory_layer_size, self.max_episode_length)
        self.lin_policy = nn.Linear(self.memory_layer_size, self.hidden_size)
        nn.init.orthogonal_(self.lin_policy.weight, np.sqrt(2))
        self.lin_value = nn.Linear(self.memory_layer_size, self.hidden_size)
        nn.init.orthogonal_(self.lin_value.weight, np.sqrt(2))
        self.policy_branches = nn.ModuleList()
        for num_actions in self.action_space_shape:
            branch = nn.Linear(self.hidden_size, num_actions)
            nn.init.orthogonal_(branch.weight, np.sqrt(0.01))
            self.policy_branches.append(branch)
        self.value = nn.Linear(self.hidden_size, 1)
        nn.init.orthogonal_(self.value.weight, 1)

        # ---- native layout ----
        self._feat_dim = feat_dim
        self._sum_actions = int(sum(self.action_space_shape))
        self._cfg = native.make_config(
            tcfg["embed_dim"], tcfg["num_heads"], tcfg["num_blocks"], tcfg["memory_length"], self.hidden_size, feat_dim,
            tcfg["layer_norm"], tcfg["positional_encoding"], tcfg.get("gtrxl", False), max_episode_length,
            self.action_space_shape, self.observation_space_shape[0] if self._visual else 0)
        self._layout, self._arena_floats, self._n_groups = native.layout(self._cfg)
        named = dict(self.named_parameters())
        missing = [n for n, *_ in self._layout if n not in named]
        extra = [n for n in named if n not in {e[0] for e in self._layout}]
        if missing or extra:
            raise RuntimeError("parameter layout mismatch: missing %s extra %s" % (missing, extra))
        self._trunk_names = {n for n, *_ in self._layout if not n.startswith("conv")}
        # rollout-sized forwards (n <= FUSED_MAX_BATCH, no grad) run the one-launch cluster kernel (csrc/rollout_fused.cu): a
        # 4-CTA cluster per sample instead of ~45 latency-bound launches; TRXL_FUSED_ROLLOUT=0 keeps the layered path
        self._fused_ok = native.fused_forward_supported(self._cfg) and os.environ.get("TRXL_FUSED_ROLLOUT", "1") != "0"
        # training-time encoder on the tcgen05 tensor cores (3xTF32 implicit GEMMs); TRXL_CUDNN_ENCODER=1 keeps cuDNN
        self._tc_encoder = (self._visual and os.environ.get("TRXL_CUDNN_ENCODER", "0") != "1" and
                            native.conv_train_supported(self._cfg, *self.observation_space_shape[1:]))
        self._arena = self._grad_arena = None
        self._pe_cache = None
        self._ws_cache = {}
        self._pack()

    # ------------------------------------------------------------------------------ arena handling
    def _pack(self):
        """(Re)build the flat arenas on the parameters' current device and re-point every parameter
        (and its .grad) at its slice.  Called at construction and after every ``.to()/.cuda()/.cpu()``."""
        named = dict(self.named_parameters())
        device = next(iter(named.values())).device
        arena = torch.zeros(self._arena_floats, dtype=torch.float32, device=device)
        # the gradient arena carries an 8-float tail: the loss statistics of the step ride along in the single
        # all-reduce of a multi-GPU optimiser step (parallel.py)
        grads_full = torch.zeros(self._arena_floats + self.GRAD_TAIL, dtype=torch.float32, device=device)
        grads = grads_full[:self._arena_floats]
        self._param_slices = []
        with torch.no_grad():
            for name, off, shape, _group in self._layout:
                p = named[name]
                numel = int(np.prod(shape)) if shape else 1
                arena[off:off + numel].copy_(p.detach().reshape(-1).to(torch.float32))
                p.data = arena[off:off + numel].view(shape)
                p.grad = grads[off:off + numel].view(shape)
                self._param_slices.append((name, off, numel, shape))
        self._arena, self._grad_arena, self._grad_full = arena, grads, grads_full
        self._pe_cache = None
        self._ws_cache = {}
        self._chunks = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._pack()
        return out

    def flat_parameters(self):
        return self._arena

    def flat_grads(self):
        return self._grad_arena

    GRAD_TAIL = 8

    def flat_grads_with_tail(self):
        """Gradient arena + its 8-float tail (``stats_tail``) as one contiguous buffer: what a multi-GPU step all-reduces."""
        return self._grad_full

    def stats_tail(self):
        return self._grad_full[self._arena_floats:self._arena_floats + 6]

    def grad_chunks(self, chunk=8192):
        """(nchunks, 3) int64 {start, length, group} table for the fused norm/clip kernel."""
        if self._chunks is None:
            rows = []
            for name, off, shape, group in self._layout:
                numel = int(np.prod(shape)) if shape else 1
                for s in range(0, numel, chunk):
                    rows.append((off + s, min(chunk, numel - s), group))
            self._chunks = torch.tensor(rows, dtype=torch.int64, device=self._arena.device)
        return self._chunks

    def _pe_table(self):
        mode = self.transformer.config["positional_encoding"]
        if mode != "relative":
            return None                                # learned: read from the arena; none: no table
        if self._pe_cache is None or self._pe_cache.device != self._arena.device:
            self._pe_cache = self.transformer.pos_embedding(self.max_episode_length).to(self._arena.device).contiguous()
        return self._pe_cache

    def _ws_floats(self, n):
        return native.workspace_floats(self._cfg, n)

    def workspace(self, n):
        ws = self._ws_cache.get(n)
        if ws is None or ws.device != self._arena.device:
            ws = torch.empty(self._ws_floats(n), dtype=torch.float32, device=self._arena.device)
            self._ws_cache[n] = ws
        return ws

    def _alloc_outputs(self, n, device):
        t = self.transformer
        return (torch.empty((n, self._sum_actions), device=device), torch.empty((n,), device=device),
                torch.empty((n, t.num_blocks, t.embed_dim), device=device))

    # ------------------------------------------------------------------------------ encoders
    def encode(self, obs, weights_packed=False, slot=0):
        """CNN encoder for image observations (model.py:87-94); identity for vector observations.
        On the GPU the three convolutions run as tcgen05 implicit GEMMs (csrc/tc_conv.cu) for observations with up to
        4 channels, otherwise as im2col + SIMT GEMM without autograd and cuDNN under autograd.  ``weights_packed=True``
        (rollout steps after the first: the weights are frozen) skips the conversion of the weights to tensor-core format.
        ``slot`` selects one of several independent no-grad workspaces, so that worker groups of the rollout can run the
        encoder concurrently on different streams."""
        if not self._visual:
            return obs
        if not obs.is_cuda:
            return _conv_features(obs, self.conv1, self.conv2, self.conv3)
        if torch.is_grad_enabled():
            if self._tc_encoder:
                return _EncoderFunction.apply(self, obs, *[p for name, p in self.named_parameters() if name.startswith("conv")])
            return _conv_features(obs, self.conv1, self.conv2, self.conv3)
        obs = obs.contiguous()
        n, _, h, w = obs.shape
        if self._tc_encoder:
            fresh = ("enci", n, h, w, slot) not in self._ws_cache
            ws = self._inference_enc_ws(n, h, w, slot)
            native.conv_train_forward(self._cfg, self._arena, obs, None, n, ws[0], ws[1], repack=fresh or not weights_packed)
            return ws[1]
        key = ("enc", n, h, w, slot)
        ws = self._ws_cache.get(key)
        if ws is None:
            ws = (torch.empty(native.conv_encoder_workspace_floats(self._cfg, n, h, w), dtype=torch.float32, device=obs.device),
                  torch.empty((n, self._feat_dim), dtype=torch.float32, device=obs.device))
            self._ws_cache[key] = ws
        native.conv_encoder_forward(self._cfg, self._arena, obs, ws[0], ws[1])
        return ws[1]

    def _inference_enc_ws(self, n, h, w, slot=0):
        key = ("enci", n, h, w, slot)
        ws = self._ws_cache.get(key)
        if ws is None:
            dev = self._arena.device
            ws = (torch.empty(native.conv_train_workspace_floats(self._cfg, n, h, w), dtype=torch.float32, device=dev),
                  torch.empty((n, self._feat_dim), dtype=torch.float32, device=dev))
            self._ws_cache[key] = ws
        return ws

    def pack_encoder_weights(self, n, h, w, slot=0):
        """Convert the conv weights to tensor-core format for the no-grad encoder of batch size n (the rollout calls this
        once per rollout and then runs ``encode(..., weights_packed=True)`` while the weights are frozen)."""
        native.conv_train_pack_weights(self._cfg, self._arena, n, h, w, self._inference_enc_ws(n, h, w, slot)[0])

    def _enc_ws(self, n, h, w):
        key = ("enct", n, h, w)
        ws = self._ws_cache.get(key)
        if ws is None:
            dev = self._arena.device
            for k in [k for k in self._ws_cache if isinstance(k, tuple) and k[0] == "enct"]:
                del self._ws_cache[k]                # one training batch size resident (the workspace holds all activations)
            ws = (torch.empty(native.conv_train_workspace_floats(self._cfg, n, h, w), dtype=torch.float32, device=dev),
                  torch.empty((n, self._feat_dim), dtype=torch.float32, device=dev))
            self._ws_cache[key] = ws
        return ws

    def encode_train(self, obs, sample_index, n):
        """Tensor-core encoder forward for a training batch: rows ``obs[sample_index]`` (or ``obs[:n]``) -> features (n, F).
        Activations stay in the encoder workspace for ``encode_backward``."""
        h, w = obs.shape[-2:]
        ws, feat = self._enc_ws(n, h, w)
        native.conv_train_forward(self._cfg, self._arena, obs, sample_index, n, ws, feat)
        return feat

    def encode_backward(self, n, h, w, dfeat):
        """d loss / d features -> the conv slices of the gradient arena (overwritten)."""
        native.conv_train_backward(self._cfg, self._grad_arena, n, h, w, self._enc_ws(n, h, w)[0], dfeat)

    # ------------------------------------------------------------------------------ native trunk
    FUSED_MAX_BATCH = 96      # batch sizes that take the one-launch cluster-per-sample trunk kernel

    def forward_table(self, feat, table, ep_index, win_index, mask, pe_index, sample_index=None, n=None, ws=None, out=None,
                      fused=None):
        """No-grad trunk forward reading memory windows in place from an episode table
        (E, slots, B, D).  Returns raw (logits (N, sumA), value (N,), new_memory (N, B, D)).
        `fused=True` runs the one-launch per-sample kernel (inference only: it saves no activations)."""
        n = feat.shape[0] if n is None else n
        if fused is None:
            fused = self._fused_ok and n <= self.FUSED_MAX_BATCH
        ws = None if fused else (self.workspace(n) if ws is None else ws)
        logits, value, out_mem = self._alloc_outputs(n, feat.device) if out is None else out
        native.model_forward(self._cfg, self._arena, feat, table, table.shape[1], ep_index, win_index, mask, pe_index,
                             sample_index, self._pe_table(), n, ws, logits, value, out_mem)
        return logits, value, out_mem

    def backward_table(self, feat, table, ep_index, win_index, mask, pe_index, sample_index, n, ws, out_mem, dlogits, dvalue,
                       dfeat=None):
        native.model_backward(self._cfg, self._arena, self._grad_arena, feat, table, table.shape[1], ep_index, win_index, mask,
                              pe_index, sample_index, self._pe_table(), n, ws, out_mem, dlogits, dvalue, dfeat)

    def split_logits(self, logits):
        return list(torch.split(logits, list(self.action_space_shape), dim=1))

    # ------------------------------------------------------------------------------ reference API
    def forward(self, obs, memory, memory_mask, memory_indices):
        """obs (N, *obs_shape), memory (N, L, B, D) window, memory_mask (N, L), memory_indices (N, L).
        Returns ([Categorical per branch], value (N,), memory (N, B, D)) like reference model.py:71-112."""
        if not self._arena.is_cuda:
            raise native.NativeLibraryError("ActorCriticModel.forward needs the model on a CUDA device "
                                            "(libtrxlppo has no CPU path); call model.to('cuda')")
        dev = self._arena.device
        obs = obs.to(dev, torch.float32)
        memory = memory.to(dev, torch.float32).contiguous()
        mask = (memory_mask.to(dev) != 0).to(torch.uint8).contiguous()
        indices = memory_indices.to(dev, torch.int64).contiguous()
        feat = self.encode(obs)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            named = dict(self.named_parameters())
            params = [named[name] for name, *_ in self._param_slices]        # layout order == grad order of _TrunkFunction
            logits, value, new_mem = _TrunkFunction.apply(self, feat.contiguous(), memory, memory.shape[1], None, None, mask,
                                                          indices, *params)
        else:
            logits, value, new_mem = self.forward_table(feat.contiguous(), memory, None, None, mask, indices)
        pi = [Categorical(logits=lg) for lg in self.split_logits(logits)]
        return pi, value, new_mem

    def get_conv_output(self, shape):
        with torch.no_grad():
            o = self.conv3(self.conv2(self.conv1(torch.zeros(1, *shape))))
        return int(np.prod(o.size()))

    # ------------------------------------------------------------------------------ gradient norms
    def group_names(self):
        b, nb = self.transformer.num_blocks, len(self.action_space_shape)
        return (["encoder", "linear_layer"] + ["transformer_block_%d" % i for i in range(b)] +
                ["policy_head_%d" % k for k in range(nb)] + ["lin_policy", "_lin_value", "_value_head", "_other"])

    def grad_norms_from(self, norms):
        """Translate the (G+2,) vector written by trxl_clip_adamw_step (per-group norms of the
        unclipped grads, total, clip coefficient) into the reference's get_grad_norm() dict
        (model.py:128-151), which reports norms of the *clipped* grads."""
        v = norms.detach().to("cpu", torch.float64).tolist()
        g = self._n_groups
        coef = v[g + 1]
        names = self.group_names()
        out = {}
        for i, name in enumerate(names):
            if name.startswith("_") or (name == "encoder" and not self._visual):
                continue
            out[name] = coef * v[i]
        head_sq = v[names.index("_value_head")] ** 2
        out["value"] = coef * float(np.sqrt(v[names.index("_lin_value")] ** 2 + head_sq))   # lin_value + value (model.py:148)
        out["model"] = coef * float(np.sqrt(v[g] ** 2 + head_sq))      # model.py:149 counts the value head twice
        return out

    def get_grad_norm(self):
        """Reference-compatible gradient-norm report computed from the current ``.grad`` arena."""
        g = self._grad_arena
        out = {}

        def norm(names):
            sel = [g[off:off + numel] for name, off, numel, _ in self._param_slices if name in names]
            return torch.linalg.norm(torch.cat(sel)).item() if sel else None
        all_names = [n for n, *_ in self._param_slices]
        if self._visual:
            out["encoder"] = norm({n for n in all_names if n.startswith("conv")})
        out["linear_layer"] = norm({n for n in all_names if n.startswith("lin_hidden.")})
        for i in range(self.transformer.num_blocks):
            pre = "transformer.transformer_blocks.%d." % i
            out["transformer_block_%d" % i] = norm({n for n in all_names if n.startswith(pre)})
        for k in range(len(self.action_space_shape)):
            out["policy_head_%d" % k] = norm({n for n in all_names if n.startswith("policy_branches.%d." % k)})
        out["lin_policy"] = norm({n for n in all_names if n.startswith("lin_policy.")})
        value_names = {n for n in all_names if n.startswith("lin_value.") or n.startswith("value.")}
        out["value"] = norm(value_names)
        total_sq = float(torch.sum(g * g))
        vsq = float(sum(torch.sum(g[off:off + numel] ** 2) for name, off, numel, _ in self._param_slices
                        if name.startswith("value.")))
        out["model"] = float(np.sqrt(total_sq + vsq))
        return out
