"""Pin the CPU oracle against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only; no CUDA, no /root/reference at run time."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import ppo_oracle as O
from oracle import trxl_oracle as X

from environments.synthetic_env import SyntheticEnv

torch.set_num_threads(1)


def sd_of(g, prefix="sd."):
    return {k[len(prefix):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith(prefix)}


def cfg_from_sd(P, max_steps, action_shape, **extra):
    """Recover the config that the fixture's model was built with from shapes/names alone."""
    nb = 1 + max(int(k.split(".")[2]) for k in P if k.startswith("transformer.transformer_blocks."))
    b0 = "transformer.transformer_blocks.0."
    d = P[b0 + "attention.fc_out.bias"].shape[0]
    pe = "relative" if "transformer.pos_embedding.inv_freqs" in P else ("learned" if "transformer.pos_embedding" in P else "")
    cfg = {
        "hidden_layer_size": P["lin_policy.weight"].shape[0],
        "max_episode_steps": int(max_steps),
        "action_space_shape": tuple(int(a) for a in action_shape),
        "transformer": {"num_blocks": nb, "embed_dim": d, "positional_encoding": pe,
                        "layer_norm": "pre" if (b0 + "norm_kv.weight") in P else "post",
                        "gtrxl": (b0 + "gate1.bg") in P},
    }
    cfg.update(extra)
    return cfg


HEADS = {"post_rel": 2, "pre_rel": 2, "pre_learned_gtrxl": 4, "post_learned_gtrxl": 1, "pre_none": 1,
         "post_rel_visual": 4, "post_rel_n1": 2, "pre_rel_L33_D48": 3}


def test_tables_bit_exact():
    g = load_golden("tables")
    for k, v in g.items():
        if k.startswith("mask_L"):
            L = int(k[6:])
            assert np.array_equal(X.attention_mask_table(L).numpy(), v)
        else:
            L, M = (int(s[1:]) for s in k.split("_")[1:])
            got = X.window_index_table(M, L).numpy()
            assert got.dtype == np.int64 and np.array_equal(got, v)


def test_units():
    g = load_golden("units")
    for name in ("mha_a", "mha_b", "mha_c"):
        P = {"values.weight": g[name + ".Wv"], "keys.weight": g[name + ".Wk"], "queries.weight": g[name + ".Wq"],
             "fc_out.weight": g[name + ".Wo"], "fc_out.bias": g[name + ".bo"]}
        P = {k: torch.from_numpy(v) for k, v in P.items()}
        v, q, mask = (torch.from_numpy(g[name + s]) for s in (".v", ".q", ".mask"))
        out, att = X.multi_head_attention(P, "", v, v, q, mask, int(g[name + ".H"]))
        np.testing.assert_allclose(out.numpy(), g[name + ".out"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(att.numpy(), g[name + ".att"], rtol=0, atol=1e-7)
        # fully masked row -> uniform attention (reference fills -1e20, not -inf)
        np.testing.assert_allclose(att.numpy()[0], 1.0 / v.shape[1], rtol=1e-6)
    P = {k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("gru.")}
    out = X.gru_gate(P, "", P["x"], P["y"])
    np.testing.assert_allclose(out.numpy(), g["gru.out"], atol=1e-6)
    for k in g:
        if k.startswith("sin_D"):
            d, m = int(k.split("_")[1][1:]), int(k.split("_")[2][1:])
            np.testing.assert_allclose(X.sinusoidal_table(m, d).numpy(), g[k], atol=1e-7)
    got = X.select_window(torch.from_numpy(g["bis.src"]), torch.from_numpy(g["bis.idx"]))
    assert np.array_equal(got.numpy(), g["bis.out"])
    for name in ("gae_a", "gae_b"):
        adv = O.gae(torch.from_numpy(g[name + ".last_value"]), g[name + ".rewards"], g[name + ".dones"],
                    torch.from_numpy(g[name + ".values"]), float(g[name + ".gamma"]), float(g[name + ".lamda"]))
        assert np.array_equal(adv.numpy(), g[name + ".adv"])          # bit-exact
    want = g["poly"]
    got = [O.polynomial_decay(3e-4, 1e-5, 100, p, s) for p in (1.0, 2.0) for s in (0, 1, 50, 100, 101)]
    assert np.array_equal(np.array(got), want)


@pytest.mark.parametrize("name", golden_names("forward_"))
def test_forward(name):
    g = load_golden(name)
    P = sd_of(g)
    cfg = cfg_from_sd(P, g["max_steps"], g["action_shape"])
    cfg["transformer"]["num_heads"] = HEADS[name[len("forward_"):]]
    cfg["transformer"]["memory_length"] = g["mask"].shape[1]
    obs, mem, mask, idx = (torch.from_numpy(g[k]) for k in ("obs", "memory", "mask", "indices"))
    logits, value, new_mem = X.model_forward(P, cfg, obs, mem, mask, idx)
    np.testing.assert_allclose(value.numpy(), g["value"], atol=2e-6)
    np.testing.assert_allclose(new_mem.numpy(), g["new_mem"], atol=2e-6)
    for k, lg in enumerate(logits):
        norm = lg - lg.logsumexp(-1, keepdim=True)
        np.testing.assert_allclose(norm.numpy(), g["logits%d" % k], atol=2e-6)


@pytest.mark.parametrize("name", golden_names("minibatch_"))
def test_minibatch_step(name):
    g = load_golden(name)
    P = sd_of(g)
    cfg = cfg_from_sd(P, g["max_steps"], g["action_shape"], value_loss_coefficient=0.25, max_grad_norm=0.5)
    cfg["transformer"]["num_heads"] = HEADS[name[len("minibatch_"):]]
    mb = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("in.")}
    opt = {}
    for it in range(2):
        stats, grads = O.train_minibatch(P, opt, cfg, mb, 3e-4 / (it + 1), 0.2, 1e-3)
        np.testing.assert_allclose(stats, g["it%d.stats" % it], rtol=2e-5, atol=2e-6)
        for n, gr in grads.items():
            np.testing.assert_allclose(gr.numpy(), g["it%d.grad.%s" % (it, n)], rtol=1e-4, atol=2e-6, err_msg=n)
            np.testing.assert_allclose(P[n].numpy(), g["it%d.param.%s" % (it, n)], rtol=0, atol=5e-6, err_msg=n)
        norms = O.grad_norm_groups(P, grads, cfg)
        for k, v in norms.items():
            np.testing.assert_allclose(v, float(g["it%d.gradnorm.%s" % (it, k)]), rtol=1e-4)


@pytest.mark.parametrize("name", golden_names("train_"))
def test_rollout_and_epochs(name):
    """Two full updates (rollout -> GAE -> epochs x minibatches) with deterministic in-process envs."""
    g = load_golden(name)
    P = sd_of(g)
    nact, max_steps = int(g["n_actions"]), int(g["max_steps"])
    obs_shape = tuple(int(x) for x in g["obs_shape"])
    visual = len(obs_shape) > 1
    cfg = cfg_from_sd(P, max_steps, (nact,), gamma=0.99, lamda=0.95, value_loss_coefficient=0.25, max_grad_norm=0.5,
                      epochs=2, n_workers=2 if visual else 3, worker_steps=8 if visual else 12,
                      n_mini_batch=2 if visual else 3)
    cfg["transformer"]["num_heads"] = 4 if ("visual" in name or name == "train_pre_rel") else 2
    cfg["transformer"]["memory_length"] = g["u0.memory_mask"].shape[2]
    # worker env seeds: the reference built a dummy env first (seed 0), then workers 1..W
    envs = [SyntheticEnv(obs_shape, nact, max_steps, min_episode_steps=2, seed=1 + i) for i in range(cfg["n_workers"])]
    st = O.new_rollout_state(cfg, obs_shape, envs)
    opt = {}
    for upd in range(2):
        pre = "u%d." % upd
        torch.manual_seed(100 + upd)
        buf, infos = O.sample_rollout(P, cfg, st, envs)
        for k in ("actions", "memory_mask", "memory_index", "memory_indices", "dones"):
            assert np.array_equal(np.asarray(buf[k]), g[pre + k]), k
        for k in ("obs", "rewards"):
            assert np.array_equal(np.asarray(buf[k]), g[pre + k]), k
        for k in ("values", "log_probs", "advantages", "memories"):
            np.testing.assert_allclose(buf[k].numpy(), g[pre + k], atol=5e-6, err_msg=k)
        assert np.array_equal(st["step"].numpy(), g[pre + "worker_step"])
        np.testing.assert_allclose(st["memory"].numpy(), g[pre + "live_memory"], atol=5e-6)
        torch.manual_seed(200 + upd)
        stats = O.train_epochs(P, opt, cfg, buf, float(g[pre + "lr"]), 0.2, 1e-3)
        np.testing.assert_allclose(np.array(stats), g[pre + "stats"], rtol=1e-4, atol=5e-6)
        for n in X.trainable_names(P):
            np.testing.assert_allclose(P[n].numpy(), g[pre + "after." + n], atol=2e-5, err_msg=n)
