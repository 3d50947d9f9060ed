"""Fused global-norm clip + AdamW over the model's flat parameter arena.

Replaces the ``torch.nn.utils.clip_grad_norm_`` + ``optim.AdamW.step`` pair of the reference
(trainer.py:59,307-312) with one native call (``trxl_clip_adamw_step``): a segmented sum of squares
over the gradient arena, a one-block finalisation (total norm, clip coefficient, per-group norms for
``get_grad_norm``), and one element-wise AdamW pass that also writes back the clipped gradient.
No host synchronisation: the clip coefficient never leaves the device.

The object quacks like a torch optimizer where the reference touches it
(``param_groups[i]["lr"]``, ``zero_grad()``, ``step()``, ``state_dict()``)."""
import torch

import trxl_native as native


class FusedClipAdamW:
    def __init__(self, model, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=None):
        self.model = model
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.param_groups = [dict(self.defaults, params=list(model.parameters()))]
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        self._state_for = None
        self._ensure_state()

    def _ensure_state(self):
        arena = self.model.flat_parameters()
        if self._state_for is not arena:
            self.exp_avg = torch.zeros_like(arena)
            self.exp_avg_sq = torch.zeros_like(arena)
            g = self.model._n_groups
            self.norms = torch.zeros(g + 2, dtype=torch.float32, device=arena.device)
            self._partial = None
            self._state_for = arena
            self.step_count = 0

    def zero_grad(self, set_to_none=False):
        self.model.flat_grads().zero_()

    def step(self, max_grad_norm=None, norms_out=None):
        """Clip (if a max norm is given here or at construction) and apply AdamW.  ``norms_out`` may be
        a (G+2,) device tensor to receive [group norms..., total norm, clip coefficient]."""
        self._ensure_state()
        model = self.model
        max_norm = self.max_grad_norm if max_grad_norm is None else max_grad_norm
        group = self.param_groups[0]
        self.step_count += 1
        norms = self.norms if norms_out is None else norms_out
        if max_norm is not None:
            chunks = model.grad_chunks()
            if self._partial is None or self._partial.numel() < chunks.shape[0]:
                self._partial = torch.empty(chunks.shape[0], dtype=torch.float32, device=chunks.device)
            nchunks = chunks.shape[0]
        else:
            chunks, nchunks, max_norm = None, 0, 0.0
        native.clip_adamw_step(model.flat_parameters(), model.flat_grads(), self.exp_avg, self.exp_avg_sq,
                               model.flat_parameters().numel(), chunks, nchunks, model._n_groups, max_norm, group["lr"],
                               self.step_count, self._partial, norms, betas=group["betas"], eps=group["eps"],
                               weight_decay=group["weight_decay"])
        return norms

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg.detach().cpu(), "exp_avg_sq": self.exp_avg_sq.detach().cpu(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, state):
        self._ensure_state()
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        self.step_count = int(state["step"])
        for g, s in zip(self.param_groups, state["param_groups"]):
            g.update(s)
