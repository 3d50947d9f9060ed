"""CPU restatement of the reference's PPO data path: GAE, rollout bookkeeping, minibatch gather, the
clipped loss, global-norm clipping and AdamW.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

State lives in plain dicts of tensors; environments are in-process objects with the reference's
gym-style protocol (``reset() -> obs``; ``step(action_array) -> (obs, reward, done, info)`` where a
truthy ``info`` marks the end of an episode, reference trainer.py:195)."""
import numpy as np
import torch

from . import trxl_oracle as X


# --------------------------------------------------------------------------------------------------
# GAE  (reference buffer.py:95-113)
# --------------------------------------------------------------------------------------------------
def gae(last_value, rewards, dones, values, gamma, lamda):
    """rewards (W,T) f32, dones (W,T) bool, values (W,T) f32, last_value (W,) -> advantages (W,T).

    Op order kept: ``lv*=m; la*=m; delta = r + gamma*lv - v; la = delta + (gamma*lamda)*la`` with
    gamma*lamda formed in Python double first (buffer.py:110-111)."""
    rewards = torch.as_tensor(rewards)
    keep = torch.as_tensor(dones).logical_not()
    adv = torch.zeros_like(values)
    last_adv = 0
    last_value = last_value.detach()
    for t in reversed(range(values.shape[1])):
        last_value = last_value * keep[:, t]
        last_adv = last_adv * keep[:, t]
        delta = rewards[:, t] + gamma * last_value - values[:, t]
        last_adv = delta + gamma * lamda * last_adv
        adv[:, t] = last_adv
        last_value = values[:, t]
    return adv


# --------------------------------------------------------------------------------------------------
# rollout  (reference trainer.py:145-237)
# --------------------------------------------------------------------------------------------------
def new_rollout_state(cfg, obs_shape, envs):
    """What reference PPOTrainer.__init__ sets up (trainer.py:65-90): first observations, per-worker
    episode step, per-worker full-episode memory, the mask table and the window-index table."""
    t = cfg["transformer"]
    w = len(envs)
    m = cfg["max_episode_steps"]
    st = {
        "obs": np.zeros((w,) + tuple(obs_shape), dtype=np.float32),
        "step": torch.zeros((w,), dtype=torch.long),
        "memory": torch.zeros((w, m, t["num_blocks"], t["embed_dim"]), dtype=torch.float32),
        "mask_table": X.attention_mask_table(t["memory_length"]),
        "index_table": X.window_index_table(m, t["memory_length"]),
    }
    for i, env in enumerate(envs):
        st["obs"][i] = env.reset()
    return st


def sample_rollout(P, cfg, st, envs):
    """One ``_sample_training_data`` (trainer.py:145-225).  Returns (buffer dict, episode infos).

    Quirks kept: the buffer's episode list starts as *views* of the live per-worker memory and a
    finished episode is cloned only when it ends (:154,206), so an unfinished episode's table entry
    keeps receiving rows; memory is zeroed on reset (:208); a new episode is appended only if
    ``t < T-1`` (:209); ``get_last_value`` uses window ``[clip(s-L,0), clip(s,L))`` with PE indices
    taken from the last rollout step (:230-236)."""
    t_cfg = cfg["transformer"]
    w, big_t, mem_len = len(envs), cfg["worker_steps"], t_cfg["memory_length"]
    nbr = len(cfg["action_space_shape"])
    buf = {
        "rewards": np.zeros((w, big_t), dtype=np.float32),
        "dones": np.zeros((w, big_t), dtype=bool),
        "actions": torch.zeros((w, big_t, nbr), dtype=torch.long),
        "obs": torch.zeros((w, big_t) + tuple(st["obs"].shape[1:])),
        "log_probs": torch.zeros((w, big_t, nbr)),
        "values": torch.zeros((w, big_t)),
        "memory_mask": torch.zeros((w, big_t, mem_len), dtype=torch.bool),
        "memory_index": torch.zeros((w, big_t), dtype=torch.long),
        "memory_indices": torch.zeros((w, big_t, mem_len), dtype=torch.long),
    }
    episodes = [st["memory"][i] for i in range(w)]
    for i in range(w):
        buf["memory_index"][i] = i
    infos = []
    rows = torch.arange(w)
    for t in range(big_t):
        with torch.no_grad():
            obs_t = torch.tensor(st["obs"])
            buf["obs"][:, t] = obs_t
            buf["memory_mask"][:, t] = st["mask_table"][torch.clip(st["step"], 0, mem_len - 1)]
            buf["memory_indices"][:, t] = st["index_table"][st["step"]]
            window = X.select_window(st["memory"], buf["memory_indices"][:, t])
            logits, value, new_mem = X.model_forward(P, cfg, obs_t, window, buf["memory_mask"][:, t],
                                                     buf["memory_indices"][:, t])
            st["memory"][rows, st["step"]] = new_mem
            acts, lps = [], []
            for lg in logits:
                dist = torch.distributions.Categorical(logits=lg)
                a = dist.sample()
                acts.append(a)
                lps.append(dist.log_prob(a))
            buf["actions"][:, t] = torch.stack(acts, dim=1)
            buf["log_probs"][:, t] = torch.stack(lps, dim=1)
            buf["values"][:, t] = value
        for i, env in enumerate(envs):
            obs, buf["rewards"][i, t], buf["dones"][i, t], info = env.step(buf["actions"][i, t].cpu().numpy())
            if info:
                st["step"][i] = 0
                infos.append(info)
                obs = env.reset()
                e = buf["memory_index"][i, t]
                episodes[e] = episodes[e].clone()
                st["memory"][i] = torch.zeros_like(st["memory"][i])
                if t < big_t - 1:
                    episodes.append(st["memory"][i])
                    buf["memory_index"][i, t + 1:] = len(episodes) - 1
            else:
                st["step"][i] += 1
            st["obs"][i] = obs
    # bootstrap value (trainer.py:227-237)
    start = torch.clip(st["step"] - mem_len, 0)
    end = torch.clip(st["step"], mem_len)
    idx = torch.stack([torch.arange(int(start[b]), int(end[b])) for b in range(w)]).long()
    window = X.select_window(st["memory"], idx)
    with torch.no_grad():
        _, last_value, _ = X.model_forward(P, cfg, torch.tensor(st["obs"]), window,
                                           st["mask_table"][torch.clip(st["step"], 0, mem_len - 1)],
                                           buf["memory_indices"][:, -1])
    buf["last_value"] = last_value
    buf["advantages"] = gae(last_value, buf["rewards"], buf["dones"], buf["values"], cfg["gamma"], cfg["lamda"])
    buf["memories"] = torch.stack(episodes, dim=0)      # buffer.py:65 (prepare_batch_dict)
    return buf, infos


def flatten_buffer(buf):
    """reference buffer.py:54-70: (W,T,...) -> (W*T,...) for the eight sample keys."""
    keys = ("actions", "values", "log_probs", "advantages", "obs", "memory_mask", "memory_index", "memory_indices")
    return {k: buf[k].reshape((-1,) + tuple(buf[k].shape[2:])) for k in keys}


def minibatches(flat, memories, n_mini_batch, perm=None):
    """reference buffer.py:72-93.  ``perm`` defaults to ``torch.randperm(batch)`` (same RNG draw)."""
    batch = flat["values"].shape[0]
    if perm is None:
        perm = torch.randperm(batch)
    size = batch // n_mini_batch
    for s in range(0, batch, size):
        idx = perm[s:s + size]
        mb = {}
        for k, v in flat.items():
            if k == "memory_index":
                mb["memories"] = memories[v[idx]]          # (mb, M, B, D) -- materialised, as the reference does
            else:
                mb[k] = v[idx]
        yield mb


# --------------------------------------------------------------------------------------------------
# loss / optimiser  (reference trainer.py:258-323)
# --------------------------------------------------------------------------------------------------
def ppo_loss(logits, value, mb, clip_range, beta, vf_coef):
    """trainer.py:277-304,315-316.  Returns (loss, [policy_loss, vf_loss, loss, entropy, kl, clip_frac])."""
    nbr = len(logits)
    log_probs = torch.stack([X.categorical_log_prob(lg, mb["actions"][:, i]) for i, lg in enumerate(logits)], dim=1)
    entropies = torch.stack([X.categorical_entropy(lg) for lg in logits], dim=1).sum(1).reshape(-1)
    adv = mb["advantages"]
    norm_adv = (adv - adv.mean()) / (adv.std() + 1e-8)                      # unbiased std (:285)
    norm_adv = norm_adv.unsqueeze(1).repeat(1, nbr)
    log_ratio = log_probs - mb["log_probs"]
    ratio = torch.exp(log_ratio)
    surr1 = ratio * norm_adv
    surr2 = torch.clamp(ratio, 1.0 - clip_range, 1.0 + clip_range) * norm_adv
    policy_loss = torch.min(surr1, surr2).mean()
    ret = mb["values"] + adv
    clipped_value = mb["values"] + (value - mb["values"]).clamp(min=-clip_range, max=clip_range)
    vf_loss = torch.max((value - ret) ** 2, (clipped_value - ret) ** 2).mean()
    entropy = entropies.mean()
    loss = -(policy_loss - vf_coef * vf_loss + beta * entropy)
    kl = ((ratio - 1.0) - log_ratio).mean()
    clip_frac = (abs(ratio - 1.0) > clip_range).float().mean()
    return loss, [policy_loss.detach(), vf_loss.detach(), loss.detach(), entropy.detach(), kl.detach(), clip_frac]


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (trainer.py:311): norm of per-tensor norms; coef clamped to 1."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def adamw_step(P, grads, opt, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """torch.optim.AdamW defaults (trainer.py:59,312), single-tensor arithmetic order."""
    opt["step"] = opt.get("step", 0) + 1
    b1, b2 = betas
    bc1 = 1 - b1 ** opt["step"]
    bc2 = 1 - b2 ** opt["step"]
    for name, g in grads.items():
        p = P[name]
        m = opt.setdefault("m." + name, torch.zeros_like(p))
        v = opt.setdefault("v." + name, torch.zeros_like(p))
        p.mul_(1 - lr * weight_decay)
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
        p.addcdiv_(m, denom, value=-(lr / bc1))


def grad_norm_groups(P, grads, cfg):
    """reference model.py:128-151 (``get_grad_norm``).  ``model`` double-counts the value head (:149)."""
    def norm(prefixes, extra=()):
        sel = [g.reshape(-1) for n, g in grads.items() if any(n.startswith(p) for p in prefixes)]
        sel += [grads[n].reshape(-1) for n in extra]
        return float(torch.linalg.norm(torch.cat(sel))) if sel else None
    out = {}
    if "conv1.weight" in P:
        out["encoder"] = norm(("conv1.", "conv2.", "conv3."))
    out["linear_layer"] = norm(("lin_hidden.",))
    for i in range(cfg["transformer"]["num_blocks"]):
        out["transformer_block_%d" % i] = norm(("transformer.transformer_blocks.%d." % i,))
    for k in range(len(cfg["action_space_shape"])):
        out["policy_head_%d" % k] = norm(("policy_branches.%d." % k,))
    out["lin_policy"] = norm(("lin_policy.",))
    out["value"] = norm(("lin_value.", "value."))
    out["model"] = norm(("",), extra=("value.weight", "value.bias"))
    return out


def train_minibatch(P, opt, cfg, mb, lr, clip_range, beta):
    """One ``_train_mini_batch`` (trainer.py:258-323).  Mutates P / opt in place.
    Returns (stats list of 6 floats, dict of clipped grads)."""
    names = X.trainable_names(P)
    for n in names:
        P[n].requires_grad_(True)
        P[n].grad = None
    window = X.select_window(mb["memories"], mb["memory_indices"])          # :271
    logits, value, _ = X.model_forward(P, cfg, mb["obs"], window, mb["memory_mask"], mb["memory_indices"])
    loss, stats = ppo_loss(logits, value, mb, clip_range, beta, cfg["value_loss_coefficient"])
    loss.backward()
    grads = {}
    with torch.no_grad():
        for n in names:
            P[n].requires_grad_(False)
            grads[n] = P[n].grad
            P[n].grad = None
        clip_grad_norm(list(grads.values()), cfg["max_grad_norm"])
        adamw_step(P, grads, opt, lr)
    return [float(s) for s in stats], grads


def train_epochs(P, opt, cfg, buf, lr, clip_range, beta):
    """reference trainer.py:239-256."""
    flat = flatten_buffer(buf)
    stats = []
    for _ in range(cfg["epochs"]):
        for mb in minibatches(flat, buf["memories"], cfg["n_mini_batch"]):
            s, _ = train_minibatch(P, opt, cfg, mb, lr, clip_range, beta)
            stats.append(s)
    return stats


def polynomial_decay(initial, final, max_decay_steps, power, current_step):
    """reference utils.py:32-50."""
    if current_step > max_decay_steps or initial == final:
        return final
    return (initial - final) * ((1 - current_step / max_decay_steps) ** power) + final
