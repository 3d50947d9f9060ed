"""Generate golden fixtures by running the UNMODIFIED reference (read-only at /root/reference).

Run once in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference's hot-path modules import gym / tblib / docopt / ruamel / reprint only at import time
(SURVEY.md §8c); empty stub modules satisfy those imports and nothing on the computed path touches
them.  Env workers are replaced by in-process fakes speaking the same pipe protocol so a rollout is
deterministic.  Outputs: small ``.npz`` files next to this script, consumed by
``tests/test_oracle_golden.py`` (oracle vs reference) and ``tests/test_gpu_*.py`` (CUDA vs reference).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TRXL_REFERENCE", "/root/reference")


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install_stubs():
    class Box:
        def __init__(self, low=0, high=1, shape=(), dtype=np.float32):
            self.shape, self.low, self.high, self.dtype = tuple(shape), low, high, dtype

    class Discrete:
        def __init__(self, n):
            self.n = n

    space = _stub("gym.spaces.space")
    spaces = _stub("gym.spaces", Box=Box, Discrete=Discrete, space=space)
    _stub("gym", spaces=spaces, make=None)
    gspaces = _stub("gymnasium.spaces", Box=Box, Discrete=Discrete)
    _stub("gymnasium", spaces=gspaces, make=None)
    _stub("gym_minigrid.wrappers", ViewSizeWrapper=object, RGBImgPartialObsWrapper=object, ImgObsWrapper=object)
    _stub("gym_minigrid", wrappers=sys.modules["gym_minigrid.wrappers"])
    _stub("memory_gym")
    _stub("reprint", output=object)
    ps = _stub("tblib.pickling_support", install=lambda: None)
    _stub("tblib", pickling_support=ps)
    _stub("docopt", docopt=lambda *a, **k: {})
    _stub("ruamel.yaml", YAML=object)
    _stub("ruamel", yaml=sys.modules["ruamel.yaml"])


def load_synthetic_env():
    path = os.path.join(REPO, "episodic-transformer-memory-ppo_b200", "environments", "synthetic_env.py")
    spec = importlib.util.spec_from_file_location("_b200_synthetic_env", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.SyntheticEnv


def sd_np(model):
    return {"sd." + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def base_cfg(**over):
    cfg = {
        "gamma": 0.99, "lamda": 0.95, "updates": 1, "epochs": 2, "n_workers": 3, "worker_steps": 12,
        "n_mini_batch": 3, "value_loss_coefficient": 0.25, "hidden_layer_size": 24, "max_grad_norm": 0.5,
        "environment": {"type": "Synthetic"},
        "transformer": {"num_blocks": 2, "embed_dim": 16, "num_heads": 2, "memory_length": 4,
                        "positional_encoding": "relative", "layer_norm": "post", "gtrxl": False, "gtrxl_bias": 0.0},
        "learning_rate_schedule": {"initial": 3e-4, "final": 1e-4, "power": 1.0, "max_decay_steps": 10},
        "beta_schedule": {"initial": 1e-3, "final": 1e-4, "power": 1.0, "max_decay_steps": 10},
        "clip_range_schedule": {"initial": 0.2, "final": 0.1, "power": 1.0, "max_decay_steps": 10},
    }
    t_over = over.pop("transformer", {})
    cfg.update(over)
    cfg["transformer"].update(t_over)
    return cfg


FORWARD_CASES = {
    # name: (cfg overrides, obs_shape, action_space_shape, max_episode_steps, N)
    "post_rel": (dict(), (5,), (3,), 7, 6),
    "pre_rel": (dict(transformer=dict(layer_norm="pre")), (5,), (3,), 7, 6),
    "pre_learned_gtrxl": (dict(transformer=dict(layer_norm="pre", positional_encoding="learned", gtrxl=True,
                                                gtrxl_bias=2.0, num_heads=4)), (4,), (2, 3), 9, 5),
    "post_learned_gtrxl": (dict(transformer=dict(layer_norm="post", positional_encoding="learned", gtrxl=True,
                                                 num_heads=1, num_blocks=3)), (4,), (2,), 6, 7),
    "pre_none": (dict(transformer=dict(layer_norm="pre", positional_encoding="", num_heads=1, num_blocks=1,
                                       embed_dim=8, memory_length=6)), (3,), (2,), 6, 4),
    "post_rel_visual": (dict(transformer=dict(embed_dim=32, num_heads=4, memory_length=5)), (3, 36, 36), (3,), 10, 4),
    "post_rel_n1": (dict(), (5,), (3,), 7, 1),
    "pre_rel_L33_D48": (dict(transformer=dict(layer_norm="pre", embed_dim=48, num_heads=3, memory_length=33,
                                              num_blocks=2)), (6,), (4,), 40, 9),
}


def make_forward_inputs(ref, cfg, obs_shape, max_steps, n, seed):
    g = torch.Generator().manual_seed(seed)
    t = cfg["transformer"]
    L, B, D = t["memory_length"], t["num_blocks"], t["embed_dim"]
    mask_table = torch.tril(torch.ones((L, L)), diagonal=-1)
    rep = torch.repeat_interleave(torch.arange(0, L).unsqueeze(0), L - 1, dim=0).long()
    idx_table = torch.stack([torch.arange(i, i + L) for i in range(max_steps - L + 1)]).long()
    idx_table = torch.cat((rep, idx_table))
    steps = torch.randint(0, max_steps, (n,), generator=g)
    steps[0] = 0                                   # fully-masked row
    if n > 1:
        steps[1] = max_steps - 1                   # step >= L-1
    obs = torch.rand((n,) + tuple(obs_shape), generator=g)
    memory = torch.randn((n, L, B, D), generator=g)
    mask = mask_table[torch.clip(steps, 0, L - 1)].bool()
    indices = idx_table[steps]
    return obs, memory, mask, indices, steps


def gen_tables(out):
    d = {}
    for L, M in ((4, 7), (6, 6), (16, 32), (1, 3)):
        mask = torch.tril(torch.ones((L, L)), diagonal=-1)
        rep = torch.repeat_interleave(torch.arange(0, L).unsqueeze(0), L - 1, dim=0).long()
        idx = torch.stack([torch.arange(i, i + L) for i in range(M - L + 1)]).long()
        idx = torch.cat((rep, idx))
        d["mask_L%d" % L] = mask.numpy()
        d["idx_L%d_M%d" % (L, M)] = idx.numpy()
    np.savez_compressed(os.path.join(out, "tables.npz"), **d)


def gen_units(ref, out):
    """MultiHeadAttention / GRUGate / SinusoidalPosition / batched_index_select / calc_advantages /
    polynomial_decay called directly."""
    tr, utils, buffer = ref["transformer"], ref["utils"], ref["buffer"]
    d = {}
    torch.manual_seed(11)
    for name, (D, H, N, Lk) in {"mha_a": (16, 2, 5, 6), "mha_b": (24, 1, 3, 4), "mha_c": (32, 4, 2, 9)}.items():
        m = tr.MultiHeadAttention(D, H)
        v = torch.randn(N, Lk, D)
        q = torch.randn(N, 1, D)
        mask = torch.rand(N, Lk) > 0.4
        mask[0] = False
        o, a = m(v, v, q, mask)
        d.update({name + ".Wv": m.values.weight.detach().numpy(), name + ".Wk": m.keys.weight.detach().numpy(),
                  name + ".Wq": m.queries.weight.detach().numpy(), name + ".Wo": m.fc_out.weight.detach().numpy(),
                  name + ".bo": m.fc_out.bias.detach().numpy(), name + ".v": v.numpy(), name + ".q": q.numpy(),
                  name + ".mask": mask.numpy(), name + ".out": o.detach().numpy(), name + ".att": a.detach().numpy(),
                  name + ".H": np.int64(H)})
    gate = tr.GRUGate(12, 1.5)
    x, y = torch.randn(7, 1, 12), torch.randn(7, 1, 12)
    d.update({"gru." + k: v.detach().numpy() for k, v in gate.state_dict().items()})
    d.update({"gru.x": x.numpy(), "gru.y": y.numpy(), "gru.out": gate(x, y).detach().numpy()})
    for dim, seq in ((16, 7), (64, 32), (256, 256)):
        d["sin_D%d_M%d" % (dim, seq)] = tr.SinusoidalPosition(dim)(seq).numpy()
    src = torch.randn(4, 9, 2, 3)
    idx = torch.randint(0, 9, (4, 5))
    d.update({"bis.src": src.numpy(), "bis.idx": idx.numpy(), "bis.out": utils.batched_index_select(src, 1, idx).numpy()})
    # GAE
    Box = sys.modules["gym.spaces"].Box
    for name, (W, T, gamma, lam) in {"gae_a": (3, 12, 0.99, 0.95), "gae_b": (5, 40, 0.995, 0.9)}.items():
        cfg = base_cfg(n_workers=W, worker_steps=T)
        buf = buffer.Buffer(cfg, Box(0, 1, (3,)), (2,), 8, torch.device("cpu"))
        buf.rewards[:] = np.random.default_rng(5).normal(size=(W, T)).astype(np.float32)
        buf.dones[:] = np.random.default_rng(6).random((W, T)) < 0.2
        buf.dones[0, T - 1] = True
        buf.values[:] = torch.randn(W, T)
        lv = torch.randn(W)
        buf.calc_advantages(lv, gamma, lam)
        d.update({name + ".rewards": buf.rewards.copy(), name + ".dones": buf.dones.copy(),
                  name + ".values": buf.values.numpy().copy(), name + ".last_value": lv.numpy(),
                  name + ".adv": buf.advantages.numpy().copy(), name + ".gamma": np.float64(gamma),
                  name + ".lamda": np.float64(lam)})
    d["poly"] = np.array([utils.polynomial_decay(3e-4, 1e-5, 100, p, s) for p in (1.0, 2.0) for s in (0, 1, 50, 100, 101)])
    np.savez_compressed(os.path.join(out, "units.npz"), **d)


def gen_forward(ref, out):
    Box = sys.modules["gym.spaces"].Box
    for name, (over, obs_shape, act_shape, max_steps, n) in FORWARD_CASES.items():
        cfg = base_cfg(**over)
        torch.manual_seed(sum(map(ord, name)) % 1000)
        model = ref["model"].ActorCriticModel(cfg, Box(0, 1, obs_shape), act_shape, max_steps)
        # de-trivialise LayerNorm affine and biases so parity covers them
        with torch.no_grad():
            for k, p in model.named_parameters():
                if "norm" in k or k.endswith("bias") or k.endswith("bg"):
                    p.add_(0.1 * torch.randn_like(p))
        obs, memory, mask, indices, steps = make_forward_inputs(ref, cfg, obs_shape, max_steps, n, seed=len(name))
        pi, value, new_mem = model(obs, memory, mask, indices)
        d = sd_np(model)
        d.update({"obs": obs.numpy(), "memory": memory.numpy(), "mask": mask.numpy(), "indices": indices.numpy(),
                  "steps": steps.numpy(), "value": value.detach().numpy(), "new_mem": new_mem.detach().numpy(),
                  "max_steps": np.int64(max_steps), "action_shape": np.array(act_shape)})
        for k, dist in enumerate(pi):
            d["logits%d" % k] = dist.logits.detach().numpy()      # normalised logits
        np.savez_compressed(os.path.join(out, "forward_%s.npz" % name), **d)


class _FakePipe:
    def __init__(self, env):
        self.env, self.q = env, []

    def send(self, msg):
        cmd, data = msg
        if cmd == "step":
            self.q.append(self.env.step(data))
        elif cmd == "reset":
            self.q.append(self.env.reset())
        elif cmd == "close":
            self.q.append(None)

    def recv(self):
        return self.q.pop(0)


TRAIN_CASES = {
    "train_post_rel": (dict(), (5,), 3, 7),
    "train_pre_learned_gtrxl": (dict(transformer=dict(layer_norm="pre", positional_encoding="learned", gtrxl=True,
                                                      gtrxl_bias=1.0)), (4,), 2, 9),
    "train_pre_rel": (dict(transformer=dict(layer_norm="pre", num_heads=4)), (5,), 3, 8),
    "train_post_rel_visual": (dict(n_workers=2, worker_steps=8, n_mini_batch=2,
                                   transformer=dict(embed_dim=32, num_heads=4, memory_length=5)), (3, 36, 36), 3, 10),
}


def gen_train(ref, out):
    """Full PPOTrainer: rollout -> prepare_batch_dict -> _train_epochs, with in-process workers."""
    SyntheticEnv = load_synthetic_env()
    trainer_mod = ref["trainer"]

    class _Writer:
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass

        def close(self):
            pass

    trainer_mod.SummaryWriter = _Writer
    for name, (over, obs_shape, n_actions, max_steps) in TRAIN_CASES.items():
        cfg = base_cfg(**over)
        counter = {"n": 0}

        def create_env(env_cfg, render=False):
            env = SyntheticEnv(obs_shape, n_actions, max_steps, min_episode_steps=2, seed=counter["n"])
            counter["n"] += 1
            return env

        class Worker:
            def __init__(self, env_cfg):
                self.child = _FakePipe(create_env(env_cfg))

        trainer_mod.create_env = create_env
        trainer_mod.Worker = Worker
        cwd = os.getcwd()
        os.chdir("/tmp")
        torch.manual_seed(3)
        tr = trainer_mod.PPOTrainer(cfg, run_id="golden", device=torch.device("cpu"))
        os.chdir(cwd)
        with torch.no_grad():
            for k, p in tr.model.named_parameters():
                if "norm" in k or k.endswith("bias"):
                    p.add_(0.05 * torch.randn_like(p))
        d = sd_np(tr.model)
        d["max_steps"] = np.int64(max_steps)
        d["n_actions"] = np.int64(n_actions)
        d["obs_shape"] = np.array(obs_shape)
        for upd in range(2):                                   # two updates: memory carry-over across updates
            torch.manual_seed(100 + upd)
            tr._sample_training_data()
            tr.buffer.prepare_batch_dict()
            b = tr.buffer
            pre = "u%d." % upd
            d.update({pre + "obs": b.obs.numpy().copy(), pre + "actions": b.actions.numpy().copy(),
                      pre + "rewards": b.rewards.copy(), pre + "dones": b.dones.copy(),
                      pre + "log_probs": b.log_probs.numpy().copy(), pre + "values": b.values.numpy().copy(),
                      pre + "advantages": b.advantages.numpy().copy(), pre + "memory_mask": b.memory_mask.numpy().copy(),
                      pre + "memory_index": b.memory_index.numpy().copy(),
                      pre + "memory_indices": b.memory_indices.numpy().copy(),
                      pre + "memories": b.memories.numpy().copy(),
                      pre + "worker_step": tr.worker_current_episode_step.numpy().copy(),
                      pre + "live_memory": tr.memory.numpy().copy()})
            lr = ref["utils"].polynomial_decay(3e-4, 1e-4, 10, 1.0, upd)
            torch.manual_seed(200 + upd)
            stats, grad_info = tr._train_epochs(lr, 0.2, 1e-3)
            d[pre + "stats"] = np.array([[float(x) for x in s] for s in stats])
            d[pre + "lr"] = np.float64(lr)
            for k, v in grad_info.items():
                d[pre + "gradnorm." + k] = np.array(v)
            d.update({pre + "after." + k: v.detach().numpy().copy() for k, v in tr.model.state_dict().items()})
        np.savez_compressed(os.path.join(out, "%s.npz" % name), **d)


def gen_minibatch(ref, out):
    """One ``_train_mini_batch`` on hand-made samples, run twice (Adam state), recording raw
    (clipped) grads -- the tightest check of backward + clip + AdamW."""
    Box = sys.modules["gym.spaces"].Box
    trainer_mod = ref["trainer"]
    for name, (over, obs_shape, act_shape, max_steps, n) in FORWARD_CASES.items():
        if name.endswith("_n1"):
            continue
        cfg = base_cfg(**over)
        torch.manual_seed(7 + len(name))
        model = ref["model"].ActorCriticModel(cfg, Box(0, 1, obs_shape), act_shape, max_steps)
        with torch.no_grad():
            for k, p in model.named_parameters():
                if "norm" in k or k.endswith("bias") or k.endswith("bg"):
                    p.add_(0.1 * torch.randn_like(p))
        n = 2 * n
        obs, window, mask, indices, steps = make_forward_inputs(ref, cfg, obs_shape, max_steps, n, seed=5)
        t = cfg["transformer"]
        g = torch.Generator().manual_seed(9)
        # whole-episode memories (n, M, B, D); the window is gathered from them inside _train_mini_batch
        memories = torch.randn((n, max_steps, t["num_blocks"], t["embed_dim"]), generator=g)
        samples = {
            "obs": obs, "memories": memories, "memory_mask": mask, "memory_indices": indices,
            "actions": torch.stack([torch.randint(0, a, (n,), generator=g) for a in act_shape], dim=1),
            "values": torch.randn(n, generator=g), "advantages": torch.randn(n, generator=g),
            "log_probs": -torch.rand((n, len(act_shape)), generator=g) - 0.3,
        }
        tr = trainer_mod.PPOTrainer.__new__(trainer_mod.PPOTrainer)
        tr.config, tr.model, tr.action_space_shape = cfg, model, act_shape
        tr.optimizer = torch.optim.AdamW(model.parameters(), lr=3e-4)
        d = sd_np(model)
        d.update({"in." + k: v.numpy() for k, v in samples.items()})
        d["max_steps"] = np.int64(max_steps)
        d["action_shape"] = np.array(act_shape)
        for it in range(2):
            stats = tr._train_mini_batch(samples, 3e-4 / (it + 1), 0.2, 1e-3)
            d["it%d.stats" % it] = np.array([float(s) for s in stats])
            for k, p in model.named_parameters():
                d["it%d.grad.%s" % (it, k)] = p.grad.detach().numpy().copy()
                d["it%d.param.%s" % (it, k)] = p.detach().numpy().copy()
            for k, v in model.get_grad_norm().items():
                d["it%d.gradnorm.%s" % (it, k)] = np.float64(v)
        np.savez_compressed(os.path.join(out, "minibatch_%s.npz" % name), **d)


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import importlib
    ref = {m: importlib.import_module(m) for m in ("utils", "transformer", "model", "buffer", "trainer")}
    torch.set_num_threads(1)                  # deterministic reductions
    out = HERE
    gen_tables(out)
    gen_units(ref, out)
    gen_forward(ref, out)
    gen_minibatch(ref, out)
    gen_train(ref, out)
    for f in sorted(os.listdir(out)):
        if f.endswith(".npz"):
            print("%-44s %8.1f KB" % (f, os.path.getsize(os.path.join(out, f)) / 1024))


if __name__ == "__main__":
    main()
