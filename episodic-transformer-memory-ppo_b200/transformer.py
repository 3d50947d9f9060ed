"""TransformerXL-over-episodic-memory modules, B200-native.

Same class names, constructor arguments, parameter names/shapes (hence ``state_dict`` keys) and
``forward`` signatures as the reference ``transformer.py`` (MultiHeadAttention :8-86, TransformerBlock
:88-172, SinusoidalPosition :174-186, Transformer :188-253, GRUGate :255-298), but the modules are
parameter holders: the math runs in libtrxlppo (hand-written sm_100a kernels, see csrc/).

Two ways in:
  * ``ActorCriticModel`` (model.py) drives the whole trunk with two native calls
    (``trxl_model_forward`` / ``trxl_model_backward``); that is the training/rollout path.
  * each class's own ``forward`` below runs the same kernels piecewise for callers that use the
    classes directly (inference only: these standalone forwards do not record autograd graphs).
The query length of the memory attention is 1 everywhere in the reference (transformer.py:249) and
that is what the fused kernel implements; other query lengths raise.
"""
import math

import numpy as np
import torch
from torch import nn

import trxl_native as native


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


class MultiHeadAttention(nn.Module):
    """Memory attention with one query token per sample (reference transformer.py:8-86).

    Quirks kept: no bias on values/keys/queries; the softmax temperature is ``sqrt(embed_dim)``
    (not the head size); masked energies are filled with -1e20 (finite) so a fully masked row
    attends uniformly."""

    def __init__(self, embed_dim, num_heads):
        super().__init__()
        if embed_dim % num_heads != 0:
            raise AssertionError("Embedding dimension needs to be divisible by the number of heads")
        self.embed_dim, self.num_heads, self.head_size = embed_dim, num_heads, embed_dim // num_heads
        self.values = nn.Linear(embed_dim, embed_dim, bias=False)
        self.keys = nn.Linear(embed_dim, embed_dim, bias=False)
        self.queries = nn.Linear(embed_dim, embed_dim, bias=False)
        self.fc_out = nn.Linear(embed_dim, embed_dim)

    @torch.no_grad()
    def forward(self, values, keys, queries, mask):
        """values == keys (N, L, D) rows of memory, queries (N, 1, D), mask (N, L) -> (out (N,1,D), attention (N,H,1,L))."""
        if queries.shape[1] != 1:
            raise NotImplementedError("the fused memory attention implements query length 1 (transformer.py:249)")
        if keys is not values and not torch.equal(keys, values):
            raise NotImplementedError("keys must equal values (the reference never passes anything else)")
        n, L, d = values.shape
        h, dh = self.num_heads, self.head_size
        x = _f32(values)
        q = torch.empty((n, d), device=x.device)
        native.linear_forward(_f32(queries).reshape(n, d), _f32(self.queries.weight), None, q)
        # fold K onto the query: qk[n, h, :] = q[n, h*dh:(h+1)*dh] @ Wk[h*dh:(h+1)*dh, :]
        wk = _f32(self.keys.weight)
        qk = torch.empty((n, h, d), device=x.device)
        for i in range(h):
            native.linear_forward(q[:, i * dh:(i + 1) * dh].contiguous(), wk[i * dh:(i + 1) * dh].t().contiguous(), None,
                                  qk_i := torch.empty((n, d), device=x.device))
            qk[:, i] = qk_i
        probs = torch.empty((n, h, L), device=x.device)
        ctx = torch.empty((n, h, d), device=x.device)
        m = None if mask is None else (mask != 0).to(torch.uint8).contiguous()
        native.window_attention_forward(x, L, 1, 0, None, None, m, None, None, None, qk, None, 0, n, L, d, h, probs, ctx)
        wv = _f32(self.values.weight)
        att_o = torch.empty((n, d), device=x.device)
        for i in range(h):
            native.linear_forward(ctx[:, i].contiguous(), wv[i * dh:(i + 1) * dh].contiguous(), None,
                                  o_i := torch.empty((n, dh), device=x.device))
            att_o[:, i * dh:(i + 1) * dh] = o_i
        out = torch.empty((n, d), device=x.device)
        native.linear_forward(att_o, _f32(self.fc_out.weight), _f32(self.fc_out.bias), out)
        return out.unsqueeze(1), probs.unsqueeze(2)


class GRUGate(nn.Module):
    """GTrXL gating unit (reference transformer.py:255-298)."""

    def __init__(self, input_dim, bg=0.0):
        super().__init__()
        for name in ("Wr", "Ur", "Wz", "Uz", "Wg", "Ug"):
            lin = nn.Linear(input_dim, input_dim, bias=False)
            nn.init.xavier_uniform_(lin.weight)
            setattr(self, name, lin)
        self.bg = nn.Parameter(torch.full([input_dim], float(bg)))

    @torch.no_grad()
    def forward(self, x, y):
        shape = x.shape
        d = shape[-1]
        x2, y2 = _f32(x).reshape(-1, d), _f32(y).reshape(-1, d)
        n = x2.shape[0]

        def lin(inp, w):
            out = torch.empty((n, d), device=inp.device)
            native.linear_forward(inp, _f32(w), None, out)
            return out
        r = torch.sigmoid(lin(y2, self.Wr.weight) + lin(x2, self.Ur.weight))
        z = torch.sigmoid(lin(y2, self.Wz.weight) + lin(x2, self.Uz.weight) - self.bg)
        h = torch.tanh(lin(y2, self.Wg.weight) + lin((r * x2).contiguous(), self.Ug.weight))
        return ((1 - z) * x2 + z * h).reshape(shape)


class _NativeLayerNorm(nn.LayerNorm):
    @torch.no_grad()
    def forward(self, x):
        d = x.shape[-1]
        x2 = _f32(x).reshape(-1, d)
        y = torch.empty_like(x2)
        native.layernorm_forward(x2, _f32(self.weight), _f32(self.bias), y, None, None)
        return y.reshape(x.shape)


class TransformerBlock(nn.Module):
    """One TrXL block: (pre|post) LayerNorm, memory attention, residual or GRU gate, one Linear+ReLU
    feed-forward, residual or GRU gate (reference transformer.py:88-172)."""

    def __init__(self, embed_dim, num_heads, config):
        super().__init__()
        self.attention = MultiHeadAttention(embed_dim, num_heads)
        self.use_gtrxl = bool(config["gtrxl"]) if "gtrxl" in config else False
        if self.use_gtrxl:
            self.gate1 = GRUGate(embed_dim, config["gtrxl_bias"])
            self.gate2 = GRUGate(embed_dim, config["gtrxl_bias"])
        self.layer_norm = config["layer_norm"]
        self.norm1 = _NativeLayerNorm(embed_dim)
        self.norm2 = _NativeLayerNorm(embed_dim)
        if self.layer_norm == "pre":
            self.norm_kv = _NativeLayerNorm(embed_dim)
        self.fc = nn.Sequential(nn.Linear(embed_dim, embed_dim), nn.ReLU())

    @torch.no_grad()
    def forward(self, value, key, query, mask):
        pre, post = self.layer_norm == "pre", self.layer_norm == "post"
        q_in = self.norm1(query) if pre else query
        if pre:
            value = self.norm_kv(value)
        att, weights = self.attention(value, value, q_in, mask)
        h = self.gate1(query, att) if self.use_gtrxl else att + query
        if post:
            h = self.norm1(h)
        h_in = self.norm2(h) if pre else h
        n, _, d = h_in.shape
        ff = torch.empty((n, d), device=h_in.device)
        native.linear_forward(_f32(h_in).reshape(n, d), _f32(self.fc[0].weight), _f32(self.fc[0].bias), ff, relu=True)
        ff = ff.unsqueeze(1)
        out = self.gate2(h, ff) if self.use_gtrxl else ff + h
        if post:
            out = self.norm2(out)
        return out, weights


GatedTransformerBlock = TransformerBlock      # the reference gates via config["gtrxl"]; alias for the name BASELINE uses


class SinusoidalPosition(nn.Module):
    """"relative" positional table (reference transformer.py:174-186): rows are positions M-1 .. 0,
    columns cat(sin, cos) of position * 1e4^(-2k/D).  Built on the host with torch so the table is
    the very numbers the reference adds, then uploaded once."""

    def __init__(self, dim, min_timescale=2.0, max_timescale=1e4):
        super().__init__()
        steps = torch.arange(0, dim, min_timescale)
        self.register_buffer("inv_freqs", max_timescale ** (-steps / dim))

    def forward(self, seq_len):
        inv = self.inv_freqs.detach().to("cpu", torch.float32)
        pos = torch.arange(seq_len - 1, -1, -1.0)
        ang = pos[:, None] * inv[None, :]
        return torch.cat((ang.sin(), ang.cos()), dim=-1).to(self.inv_freqs.device)


class Transformer(nn.Module):
    """Embedding + positional encoding + block stack (reference transformer.py:188-253)."""

    def __init__(self, config, input_dim, max_episode_steps):
        super().__init__()
        self.config = config
        self.num_blocks, self.embed_dim, self.num_heads = config["num_blocks"], config["embed_dim"], config["num_heads"]
        self.max_episode_steps = max_episode_steps
        self.activation = nn.ReLU()
        self.linear_embedding = nn.Linear(input_dim, self.embed_dim)
        nn.init.orthogonal_(self.linear_embedding.weight, np.sqrt(2))
        mode = config["positional_encoding"]
        if mode == "relative":
            self.pos_embedding = SinusoidalPosition(dim=self.embed_dim)
        elif mode == "learned":
            self.pos_embedding = nn.Parameter(torch.randn(self.max_episode_steps, self.embed_dim))
        self.transformer_blocks = nn.ModuleList(
            [TransformerBlock(self.embed_dim, self.num_heads, config) for _ in range(self.num_blocks)])

    def positional_table(self):
        mode = self.config["positional_encoding"]
        if mode == "relative":
            return self.pos_embedding(self.max_episode_steps)
        if mode == "learned":
            return self.pos_embedding
        return None

    @torch.no_grad()
    def forward(self, h, memories, mask, memory_indices):
        """h (N, D_in), memories (N, L, B, D), mask (N, L), memory_indices (N, L) -> (h (N, D), out_memories (N, B, D))."""
        n = h.shape[0]
        e = torch.empty((n, self.embed_dim), device=h.device)
        native.linear_forward(_f32(h), _f32(self.linear_embedding.weight), _f32(self.linear_embedding.bias), e, relu=True)
        table = self.positional_table()
        if table is not None:
            memories = memories + table[memory_indices].unsqueeze(2)
        outs = []
        h = e
        for i, block in enumerate(self.transformer_blocks):
            outs.append(h)
            mem_i = memories[:, :, i].contiguous()
            h, _ = block(mem_i, mem_i, h.unsqueeze(1), mask)
            h = h.reshape(n, self.embed_dim)
        return h, torch.stack(outs, dim=1)
