"""Functional CPU restatement of the reference TrXL actor-critic forward.  TEST INFRASTRUCTURE ONLY.

Parameters travel as ``P: dict[str, Tensor]`` keyed by the reference's ``state_dict`` names
(SURVEY.md §8b).  ``cfg`` is the reference's nested config dict (``configs/*.yaml``) plus two derived
keys the reference reads from the environment: ``max_episode_steps`` and ``action_space_shape``.

The op sequence deliberately matches the reference one ATen call for one ATen call (F.linear,
F.layer_norm, einsum, masked_fill, softmax, gather ...) so that (a) outputs agree to rounding with the
reference's CPU path and (b) timing this port on CPU costs what the reference costs.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------
# integer / bool tables (must be bit-exact)
# --------------------------------------------------------------------------------------------------
def attention_mask_table(memory_length: int) -> torch.Tensor:
    """Strictly-lower-triangular (L, L) float table; row ``min(step, L-1)`` is a sample's mask.

    Follows reference trainer.py:78 (``tril(ones(L, L), diagonal=-1)``)."""
    rows = torch.arange(memory_length).unsqueeze(1)
    cols = torch.arange(memory_length).unsqueeze(0)
    return (cols < rows).to(torch.float32)


def window_index_table(max_episode_length: int, memory_length: int) -> torch.Tensor:
    """(M, L) int64 table of window slot indices by episode step.

    Follows reference trainer.py:88-90: the first L-1 rows are ``0..L-1``; row ``L-1+i`` is ``i..i+L-1``
    for ``i = 0 .. M-L``."""
    assert memory_length <= max_episode_length
    step = torch.arange(max_episode_length, dtype=torch.long).unsqueeze(1)
    first = torch.clamp(step - (memory_length - 1), min=0)
    return first + torch.arange(memory_length, dtype=torch.long).unsqueeze(0)


def sinusoidal_table(max_episode_steps: int, dim: int, inv_freqs: torch.Tensor = None) -> torch.Tensor:
    """(M, D) "relative" positional table.  Follows reference transformer.py:174-186:
    ``inv_freqs = 1e4 ** (-arange(0, D, 2) / D)``; positions run M-1 .. 0; ``cat(sin, cos)``."""
    if inv_freqs is None:
        inv_freqs = 1e4 ** (-torch.arange(0, dim, 2.0) / dim)
    pos = torch.arange(max_episode_steps - 1, -1, -1.0)
    ang = pos.unsqueeze(1) * inv_freqs.unsqueeze(0)
    return torch.cat((ang.sin(), ang.cos()), dim=-1)


def select_window(x: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """``out[b, l, ...] = x[b, index[b, l], ...]``.  Follows reference utils.py:52-75
    (``batched_index_select(input, 1, index)`` via ``torch.gather`` with an expanded index)."""
    view = index.reshape(index.shape + (1,) * (x.dim() - 2))
    return torch.gather(x, 1, view.expand(index.shape + tuple(x.shape[2:])))


# --------------------------------------------------------------------------------------------------
# transformer pieces
# --------------------------------------------------------------------------------------------------
def multi_head_attention(P, pre, values, keys, queries, mask, num_heads):
    """Reference transformer.py:31-86.  Note the softmax scale is sqrt(embed_dim), not sqrt(head)."""
    n, lv, d_model = values.shape
    lk, lq = keys.shape[1], queries.shape[1]
    hd = d_model // num_heads
    v = F.linear(values, P[pre + "values.weight"]).reshape(n, lv, num_heads, hd)        # :50,54
    k = F.linear(keys, P[pre + "keys.weight"]).reshape(n, lk, num_heads, hd)            # :51,55
    q = F.linear(queries, P[pre + "queries.weight"]).reshape(n, lq, num_heads, hd)      # :52,56
    energy = torch.einsum("nqhd,nkhd->nhqk", q, k)                                      # :59
    if mask is not None:
        energy = energy.masked_fill(mask.unsqueeze(1).unsqueeze(1) == 0, float("-1e20"))  # :66
    att = torch.softmax(energy / (d_model ** 0.5), dim=3)                               # :69
    out = torch.einsum("nhql,nlhd->nqhd", att, v).reshape(n, lq, d_model)               # :73
    return F.linear(out, P[pre + "fc_out.weight"], P[pre + "fc_out.bias"]), att         # :82


def gru_gate(P, pre, x, y):
    """GTrXL gate, reference transformer.py:287-298."""
    r = torch.sigmoid(F.linear(y, P[pre + "Wr.weight"]) + F.linear(x, P[pre + "Ur.weight"]))
    z = torch.sigmoid(F.linear(y, P[pre + "Wz.weight"]) + F.linear(x, P[pre + "Uz.weight"]) - P[pre + "bg"])
    h = torch.tanh(F.linear(y, P[pre + "Wg.weight"]) + F.linear(torch.mul(r, x), P[pre + "Ug.weight"]))
    return torch.mul(1 - z, x) + torch.mul(z, h)


# Optional tap (tests only): when RELU_MARGINS is a list, every ReLU of a forward appends the per-sample minimum |pre-activation|
# of that layer (shape (N,)).  The full-size parity test uses it to pick samples whose ReLU decisions cannot flip under fp32
# rounding noise, so that gradient comparisons measure arithmetic accuracy rather than ties at the ReLU boundary.
RELU_MARGINS = None


def _relu(x):
    if RELU_MARGINS is not None:
        RELU_MARGINS.append(x.detach().abs().reshape(x.shape[0], -1).min(dim=1).values)
    return torch.relu(x)


def _ln(P, pre, x):
    return F.layer_norm(x, (x.shape[-1],), P[pre + "weight"], P[pre + "bias"], 1e-5)


def transformer_block(P, pre, tcfg, value, query, mask):
    """Reference transformer.py:117-172 (key == value in every call the reference makes, :249)."""
    mode = tcfg["layer_norm"]
    gated = bool(tcfg.get("gtrxl", False))
    if mode == "pre":                                                       # :129-132
        q_in = _ln(P, pre + "norm1.", query)
        value = _ln(P, pre + "norm_kv.", value)
    else:
        q_in = query
    att, weights = multi_head_attention(P, pre + "attention.", value, value, q_in, mask, tcfg["num_heads"])
    h = gru_gate(P, pre + "gate1.", query, att) if gated else att + query   # :140-145
    if mode == "post":
        h = _ln(P, pre + "norm1.", h)                                       # :148-149
    h_in = _ln(P, pre + "norm2.", h) if mode == "pre" else h                # :152-155
    ff = _relu(F.linear(h_in, P[pre + "fc.0.weight"], P[pre + "fc.0.bias"]))  # :158
    out = gru_gate(P, pre + "gate2.", h, ff) if gated else ff + h           # :161-166
    if mode == "post":
        out = _ln(P, pre + "norm2.", out)                                   # :169-170
    return out, weights


def transformer(P, tcfg, max_episode_steps, h, memories, mask, memory_indices, pre="transformer."):
    """Reference transformer.py:222-253.  Returns (h, out_memories (N, B, D))."""
    h = _relu(F.linear(h, P[pre + "linear_embedding.weight"], P[pre + "linear_embedding.bias"]))  # :234
    pe_mode = tcfg["positional_encoding"]
    if pe_mode == "relative":                                               # :237-239
        table = sinusoidal_table(max_episode_steps, tcfg["embed_dim"], P.get(pre + "pos_embedding.inv_freqs"))
        memories = memories + table[memory_indices].unsqueeze(2)
    elif pe_mode == "learned":                                              # :241-242
        memories = memories + P[pre + "pos_embedding"][memory_indices].unsqueeze(2)
    out_mem = []
    for i in range(tcfg["num_blocks"]):                                     # :247-252
        out_mem.append(h.detach())
        blk = "%stransformer_blocks.%d." % (pre, i)
        h, _ = transformer_block(P, blk, tcfg, memories[:, :, i], h.unsqueeze(1), mask)
        h = h.squeeze()
        if h.dim() == 1:
            h = h.unsqueeze(0)
    return h, torch.stack(out_mem, dim=1)


def encode_observation(P, obs):
    """Reference model.py:84-97: Atari CNN for image observations, then lin_hidden + ReLU."""
    h = obs
    if "conv1.weight" in P:
        n = h.shape[0]
        h = _relu(F.conv2d(h, P["conv1.weight"], P["conv1.bias"], stride=4))
        h = _relu(F.conv2d(h, P["conv2.weight"], P["conv2.bias"], stride=2))
        h = _relu(F.conv2d(h, P["conv3.weight"], P["conv3.bias"], stride=1))
        h = h.reshape(n, -1)
    return _relu(F.linear(h, P["lin_hidden.weight"], P["lin_hidden.bias"]))


def model_forward(P, cfg, obs, memory, memory_mask, memory_indices):
    """Reference model.py:71-112.  Returns (list of per-branch logits, value (N,), new memory (N, B, D)).

    The reference wraps each logits tensor in ``Categorical(logits=...)`` (:110); callers here apply
    ``categorical_*`` below, which restate what that class computes."""
    h = encode_observation(P, obs)
    h, new_mem = transformer(P, cfg["transformer"], cfg["max_episode_steps"], h, memory, memory_mask, memory_indices)
    h_policy = _relu(F.linear(h, P["lin_policy.weight"], P["lin_policy.bias"]))    # :104
    h_value = _relu(F.linear(h, P["lin_value.weight"], P["lin_value.bias"]))       # :106
    value = F.linear(h_value, P["value.weight"], P["value.bias"]).reshape(-1)           # :108
    logits = [F.linear(h_policy, P["policy_branches.%d.weight" % k], P["policy_branches.%d.bias" % k])
              for k in range(len(cfg["action_space_shape"]))]
    return logits, value, new_mem


# torch.distributions.Categorical(logits=x): normalised = x - logsumexp(x); log_prob = gather;
# entropy = -(p * logp).sum(-1) with logp clamped at finfo.min.
def categorical_log_prob(logits, action):
    logp = logits - logits.logsumexp(dim=-1, keepdim=True)
    return logp.gather(-1, action.long().unsqueeze(-1)).squeeze(-1)


def categorical_entropy(logits):
    logp = logits - logits.logsumexp(dim=-1, keepdim=True)
    logp = torch.clamp(logp, min=torch.finfo(logp.dtype).min)
    return -(logp * logp.exp()).sum(-1)


# --------------------------------------------------------------------------------------------------
# parameter construction (same shapes / init recipe as the reference so a port is timed on the same
# numbers; used by tests and by the CPU baseline when no golden state_dict is supplied)
# --------------------------------------------------------------------------------------------------
def conv_out_hw(hw):
    h, w = hw
    for k, s in ((8, 4), (4, 2), (3, 1)):
        h, w = (h - k) // s + 1, (w - k) // s + 1
    return h, w


def init_params(cfg, obs_shape, seed=0, dtype=torch.float32):
    """Build a parameter dict with the reference's names, shapes and init distributions
    (model.py:27-69, transformer.py:26-29,106-115,206-213,271-285).  Not bit-identical to the
    reference's RNG stream -- goldens carry their own state_dict."""
    g = torch.Generator().manual_seed(seed)
    t = cfg["transformer"]
    d, hid, nb = t["embed_dim"], cfg["hidden_layer_size"], t["num_blocks"]
    P = {}

    def ortho(shape, gain):
        w = torch.empty(shape, dtype=dtype)
        torch.nn.init.orthogonal_(w, gain, generator=g)
        return w

    def kaiming_lin(out_f, in_f, bias=True):
        bound = 1.0 / math.sqrt(in_f)
        w = (torch.rand((out_f, in_f), generator=g, dtype=dtype) * 2 - 1) * bound
        b = (torch.rand((out_f,), generator=g, dtype=dtype) * 2 - 1) * bound if bias else None
        return w, b

    if len(obs_shape) > 1:
        c = obs_shape[0]
        for name, (o, i, k) in (("conv1", (32, c, 8)), ("conv2", (64, 32, 4)), ("conv3", (64, 64, 3))):
            P[name + ".weight"] = ortho((o, i, k, k), math.sqrt(2))
            bound = 1.0 / math.sqrt(i * k * k)
            P[name + ".bias"] = (torch.rand((o,), generator=g, dtype=dtype) * 2 - 1) * bound
        oh, ow = conv_out_hw(obs_shape[1:])
        feat = 64 * oh * ow
    else:
        feat = obs_shape[0]
    P["lin_hidden.weight"] = ortho((d, feat), math.sqrt(2))
    P["lin_hidden.bias"] = kaiming_lin(d, feat)[1]
    P["transformer.linear_embedding.weight"] = ortho((d, d), math.sqrt(2))
    P["transformer.linear_embedding.bias"] = kaiming_lin(d, d)[1]
    if t["positional_encoding"] == "relative":
        P["transformer.pos_embedding.inv_freqs"] = 1e4 ** (-torch.arange(0, d, 2.0) / d)
    elif t["positional_encoding"] == "learned":
        P["transformer.pos_embedding"] = torch.randn((cfg["max_episode_steps"], d), generator=g, dtype=dtype)
    for i in range(nb):
        b = "transformer.transformer_blocks.%d." % i
        for nm in ("values", "keys", "queries"):
            P[b + "attention.%s.weight" % nm] = kaiming_lin(d, d, False)[0]
        P[b + "attention.fc_out.weight"], P[b + "attention.fc_out.bias"] = kaiming_lin(d, d)
        if t.get("gtrxl", False):
            for gate in ("gate1.", "gate2."):
                for nm in ("Wr", "Ur", "Wz", "Uz", "Wg", "Ug"):
                    bound = math.sqrt(6.0 / (2 * d))
                    P[b + gate + nm + ".weight"] = (torch.rand((d, d), generator=g, dtype=dtype) * 2 - 1) * bound
                P[b + gate + "bg"] = torch.full((d,), float(t.get("gtrxl_bias", 0.0)), dtype=dtype)
        norms = ["norm1.", "norm2."] + (["norm_kv."] if t["layer_norm"] == "pre" else [])
        for nm in norms:
            P[b + nm + "weight"] = torch.ones(d, dtype=dtype)
            P[b + nm + "bias"] = torch.zeros(d, dtype=dtype)
        P[b + "fc.0.weight"], P[b + "fc.0.bias"] = kaiming_lin(d, d)
    P["lin_policy.weight"] = ortho((hid, d), math.sqrt(2))
    P["lin_policy.bias"] = kaiming_lin(hid, d)[1]
    P["lin_value.weight"] = ortho((hid, d), math.sqrt(2))
    P["lin_value.bias"] = kaiming_lin(hid, d)[1]
    for k, a in enumerate(cfg["action_space_shape"]):
        P["policy_branches.%d.weight" % k] = ortho((a, hid), math.sqrt(0.01))
        P["policy_branches.%d.bias" % k] = kaiming_lin(a, hid)[1]
    P["value.weight"] = ortho((1, hid), 1.0)
    P["value.bias"] = kaiming_lin(1, hid)[1]
    return P


def trainable_names(P):
    """Names that are nn.Parameters in the reference, in ``model.parameters()`` order is not needed
    here; buffers (``inv_freqs``) are excluded."""
    return [k for k in P if not k.endswith("inv_freqs")]
