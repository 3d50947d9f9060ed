"""Shared helpers for the GPU parity tests and __graft_entry__.smoke()."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "episodic-transformer-memory-ppo_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

HEADS = {"post_rel": 2, "pre_rel": 2, "pre_learned_gtrxl": 4, "post_learned_gtrxl": 1, "pre_none": 1,
         "post_rel_visual": 4, "post_rel_n1": 2, "pre_rel_L33_D48": 3}


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def config_from_golden(g, heads, memory_length, **extra):
    """Reference-style config dict recovered from a fixture's state_dict shapes."""
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    nb = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.transformer_blocks."))
    b0 = "transformer.transformer_blocks.0."
    pe = "relative" if "transformer.pos_embedding.inv_freqs" in sd else ("learned" if "transformer.pos_embedding" in sd else "")
    cfg = {
        "hidden_layer_size": int(sd["lin_policy.weight"].shape[0]),
        "value_loss_coefficient": 0.25, "max_grad_norm": 0.5,
        "transformer": {"num_blocks": nb, "embed_dim": int(sd[b0 + "attention.fc_out.bias"].shape[0]), "num_heads": heads,
                        "memory_length": int(memory_length), "positional_encoding": pe,
                        "layer_norm": "pre" if (b0 + "norm_kv.weight") in sd else "post",
                        "gtrxl": (b0 + "gate1.bg") in sd, "gtrxl_bias": 0.0},
    }
    cfg.update(extra)
    return cfg, sd


class _Space:
    def __init__(self, shape):
        self.shape = tuple(shape)


def build_model(g, heads, memory_length, obs_shape, action_shape, max_steps, device):
    from model import ActorCriticModel
    cfg, sd = config_from_golden(g, heads, memory_length)
    model = ActorCriticModel(cfg, _Space(obs_shape), tuple(int(a) for a in action_shape), int(max_steps))
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()})
    return model.to(device), cfg


def run_minibatch_parity(name, device="cuda:0", verbose=False, rtol_grad=2e-4, atol_grad=2e-6, atol_param=2e-5):
    """Two fused PPO minibatch steps on `name`'s fixture vs the values recorded from the reference.
    Returns the max abs parameter error after the second step."""
    from optim_native import FusedClipAdamW
    from trainer import PPOTrainer
    g = load_golden(name)
    case = name[len("minibatch_"):]
    inputs = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("in.")}
    obs_shape = tuple(inputs["obs"].shape[1:])
    model, cfg = build_model(g, HEADS[case], inputs["memory_mask"].shape[1], obs_shape, g["action_shape"], g["max_steps"], device)
    tr = PPOTrainer.__new__(PPOTrainer)
    tr.config, tr.model, tr.device = cfg, model, torch.device(device)
    tr.action_space_shape = tuple(int(a) for a in g["action_shape"])
    tr.obs_shape = obs_shape
    tr.optimizer = FusedClipAdamW(model, lr=3e-4, max_grad_norm=cfg["max_grad_norm"])
    from parallel import DataParallelContext
    tr.dp = DataParallelContext(tr.device)
    tr._train_state = {}
    worst = 0.0
    for it in range(2):
        stats = tr._train_mini_batch(inputs, 3e-4 / (it + 1), 0.2, 1e-3)
        np.testing.assert_allclose(np.array(stats, dtype=np.float64), g["it%d.stats" % it], rtol=1e-4, atol=1e-5,
                                   err_msg="%s stats it%d" % (name, it))
        for pname, p in model.named_parameters():
            want_g = g["it%d.grad.%s" % (it, pname)]
            got_g = p.grad.detach().cpu().numpy()
            scale = max(1e-12, float(np.abs(want_g).max()))
            np.testing.assert_allclose(got_g, want_g, rtol=rtol_grad, atol=max(atol_grad, 1e-4 * scale),
                                       err_msg="%s grad %s it%d" % (name, pname, it))
            want_p = g["it%d.param.%s" % (it, pname)]
            got_p = p.detach().cpu().numpy()
            err = float(np.abs(got_p - want_p).max())
            worst = max(worst, err)
        norms = model.grad_norms_from(tr._last_norms)
        for k, v in norms.items():
            np.testing.assert_allclose(v, float(g["it%d.gradnorm.%s" % (it, k)]), rtol=2e-4, err_msg="%s gradnorm %s" % (name, k))
        if verbose:
            print("%s it%d: stats ok, max |param - ref| = %.3e" % (name, it, worst))
    return worst
