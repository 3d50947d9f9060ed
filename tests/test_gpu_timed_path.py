"""Parity on the path bench.py actually times (`pytest -m gpu`): CUDA-graph replay of the rollout step, the
full-size c3 minibatch (N = 2048, 3x84x84 observations, sample_index into the flat buffer, rollout-built episode
table), the device-resident feed against the worker transport, checkpoint/resume, and 2-rank NCCL gradients.

Reference lines matched: trainer.py:145-225 (rollout), :258-323 (minibatch step), :356-362 (checkpoint),
enjoy.py:47-84 (inference loop)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import PKG, ROOT, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SMALL_VISUAL = {
    "environment": {"type": "Synthetic", "obs_shape": [3, 84, 84], "n_actions": 3, "max_episode_steps": 12,
                    "min_episode_steps": 3, "seed": 0},
    "gamma": 0.99, "lamda": 0.95, "updates": 4, "epochs": 2, "n_workers": 4, "worker_steps": 10, "n_mini_batch": 2,
    "value_loss_coefficient": 0.5, "hidden_layer_size": 64, "max_grad_norm": 0.5,
    "transformer": {"num_blocks": 2, "embed_dim": 64, "num_heads": 4, "memory_length": 6, "positional_encoding": "relative",
                    "layer_norm": "post", "gtrxl": False, "gtrxl_bias": 0.0},
    "learning_rate_schedule": {"initial": 3e-4, "final": 1e-4, "power": 1.0, "max_decay_steps": 10},
    "beta_schedule": {"initial": 1e-3, "final": 1e-4, "power": 1.0, "max_decay_steps": 10},
    "clip_range_schedule": {"initial": 0.2, "final": 0.1, "power": 1.0, "max_decay_steps": 10},
}


class _Pipe:
    def __init__(self, env):
        self.env, self.q = env, []

    def send(self, msg):
        cmd, data = msg
        self.q.append(self.env.step(data) if cmd == "step" else (self.env.reset() if cmd == "reset" else None))

    def recv(self):
        return self.q.pop(0)


class _Worker:
    def __init__(self, env):
        self.child = _Pipe(env)


def _cfg(**over):
    import copy
    cfg = copy.deepcopy(SMALL_VISUAL)
    for k, v in over.items():
        if isinstance(v, dict) and k in cfg:
            cfg[k].update(v)
        else:
            cfg[k] = v
    return cfg


def _synthetic_workers(cfg):
    from environments.synthetic_env import SyntheticEnv
    e = cfg["environment"]
    return [_Worker(SyntheticEnv(tuple(e["obs_shape"]), e["n_actions"], e["max_episode_steps"], e["min_episode_steps"], seed=1 + w))
            for w in range(cfg["n_workers"])]


def _buffer_snapshot(tr):
    b = tr.buffer
    snap = {k: getattr(b, k).detach().cpu().clone() for k in
            ("actions", "obs", "log_probs", "values", "advantages", "memory_mask", "memory_index", "memory_indices")}
    snap["rewards"], snap["dones"] = torch.from_numpy(b.rewards.copy()), torch.from_numpy(b.dones.copy())
    snap["memories"] = b.memories.detach().cpu().clone()
    snap["live_memory"] = tr.memory.detach().cpu().clone()
    snap["worker_step"] = tr.worker_current_episode_step.clone()
    return snap


def _assert_bit_identical(a, b, what):
    assert a.keys() == b.keys()
    for k in a:
        assert a[k].shape == b[k].shape, (what, k, a[k].shape, b[k].shape)
        assert torch.equal(a[k], b[k]), "%s: %s differs (max abs %.3e)" % (
            what, k, float((a[k].double() - b[k].double()).abs().max()))


# ------------------------------------------------------------------------------------------------ (a) graph replay
@pytest.mark.parametrize("ln,pe,gtrxl", [("post", "relative", False), ("pre", "relative", True)])
def test_graph_replay_rollout_is_bit_identical_to_eager(ln, pe, gtrxl, tmp_path, monkeypatch):
    """_sample_training_data with CUDA graphs ON (eager warm-up rollout, capture rollout, replay rollouts) against
    TRXL_NO_GRAPHS-style eager launches on the same seeded uniforms, environments and weights: every buffer tensor, the
    episode table, the live memory and the parameters after each update's optimisation must be bit-identical."""
    import trainer as trainer_mod
    monkeypatch.chdir(tmp_path)
    cfg = _cfg(transformer={"layer_norm": ln, "positional_encoding": pe, "gtrxl": gtrxl})
    trainers = []
    for graphs in (True, False):
        torch.manual_seed(7)
        tr = trainer_mod.PPOTrainer(cfg, run_id="g", device=torch.device(DEV), workers=_synthetic_workers(cfg), summary_writer=False)
        tr.use_cuda_graphs = graphs
        trainers.append(tr)
    trainers[1].model.load_state_dict(trainers[0].model.state_dict())
    for upd in range(4):
        snaps = []
        for tr in trainers:
            torch.manual_seed(100 + upd)
            tr._sample_training_data()
            tr.buffer.prepare_batch_dict()
            snaps.append(_buffer_snapshot(tr))
        _assert_bit_identical(snaps[0], snaps[1], "update %d rollout" % upd)
        for tr in trainers:
            torch.manual_seed(200 + upd)
            tr._train_epochs(3e-4, 0.2, 1e-3)
        assert torch.equal(trainers[0].model.flat_parameters(), trainers[1].model.flat_parameters()), "update %d params" % upd
    g = trainers[0]._graphs
    assert trainers[0].use_cuda_graphs and len(g.get("steps", {})) == cfg["worker_steps"], "the graph path did not run"
    assert not trainers[1]._graphs.get("steps")
    for tr in trainers:
        tr.close(exit_process=False)


@pytest.mark.gpu
@pytest.mark.parametrize("ln", ["post", "pre"])
def test_graph_replay_of_the_optimiser_step_is_bit_identical_to_eager(ln, tmp_path, monkeypatch):
    """TRXL_TRAIN_GRAPHS: the optimiser step's forward and backward segments replayed as CUDA graphs (row indices and the
    padded tile table staged at fixed addresses) against eager launches with the real tile counts: the parameters after
    every update are bit-identical, and the graphs were actually captured and replayed."""
    import trainer as trainer_mod
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("TRXL_GROUPED_ATTENTION", "1")          # the small test minibatches would otherwise take the per-sample kernel
    cfg = _cfg(transformer={"layer_norm": ln, "positional_encoding": "relative", "gtrxl": False})
    trainers = []
    for graphs in (True, False):
        torch.manual_seed(7)
        tr = trainer_mod.PPOTrainer(cfg, run_id="tg", device=torch.device(DEV), workers=_synthetic_workers(cfg), summary_writer=False)
        tr.use_train_graphs = graphs
        trainers.append(tr)
    trainers[1].model.load_state_dict(trainers[0].model.state_dict())
    for upd in range(3):
        for tr in trainers:
            torch.manual_seed(100 + upd)
            tr._sample_training_data()
            tr.buffer.prepare_batch_dict()
        stats = []
        for tr in trainers:
            torch.manual_seed(200 + upd)
            stats.append(tr._train_epochs(3e-4, 0.2, 1e-3)[0])
        assert torch.equal(trainers[0].model.flat_parameters(), trainers[1].model.flat_parameters()), "update %d params" % upd
        assert np.array_equal(np.asarray(stats[0]), np.asarray(stats[1])), "update %d statistics" % upd
    states = list(trainers[0]._train_graphs.values())
    assert trainers[0].use_train_graphs and states and all(set(st["graphs"]) == {"fwd", "bwd"} for st in states), \
        "the graph path did not run"
    assert not trainers[1]._train_graphs
    for tr in trainers:
        tr.close(exit_process=False)


# ------------------------------------------------------------------------------------------------ (c) device feed vs workers
class _ReplayEnv:
    """Environment that replays a SyntheticDeviceFeed's schedule (observations, rewards, episode ends) for one worker."""

    def __init__(self, feed, w, obs_host):
        self.feed, self.w, self.obs_host, self.t = feed, w, obs_host, 0

    def reset(self):
        return self.obs_host[self.t % (self.feed.T + 1), self.w]

    def step(self, action):
        t = self.t % self.feed.T
        reward, done = float(self.feed.rewards[t, self.w]), bool(self.feed.dones[t, self.w])
        info = next((i for (w, i) in self.feed.infos[t] if w == self.w), None)
        self.t = t + 1
        return self.obs_host[self.t, self.w], reward, done, info


def test_device_feed_rollout_matches_worker_rollout(tmp_path, monkeypatch):
    """bench.py's `value` arm feeds the rollout from a device-resident schedule (trainer._sample_from_device_feed replays the
    reference's bookkeeping for all T steps up front); the e2e arm steps worker processes.  On the same schedule both must
    produce bit-identical buffers, episode tables and parameters over consecutive updates (episodes carry over)."""
    import trainer as trainer_mod
    from device_feed import SyntheticDeviceFeed
    monkeypatch.chdir(tmp_path)
    cfg = _cfg()
    e = cfg["environment"]
    W, T = cfg["n_workers"], cfg["worker_steps"]
    torch.manual_seed(3)
    feed = SyntheticDeviceFeed(W, T, tuple(e["obs_shape"]), e["max_episode_steps"], e["min_episode_steps"], seed=5, device=DEV)
    obs_host = feed.obs_all.cpu().numpy()
    a = trainer_mod.PPOTrainer(cfg, run_id="a", device=torch.device(DEV), workers=[], summary_writer=False)
    a.device_feed = feed
    envs = [_ReplayEnv(feed, w, obs_host) for w in range(W)]
    b = trainer_mod.PPOTrainer(cfg, run_id="b", device=torch.device(DEV), workers=[_Worker(env) for env in envs], summary_writer=False)
    b.model.load_state_dict(a.model.state_dict())
    for upd in range(3):
        torch.manual_seed(50 + upd)
        infos_a = a._sample_training_data()
        a.buffer.prepare_batch_dict()
        for w, env in enumerate(envs):
            env.t = 0
            b.obs[w] = obs_host[0, w]                   # the feed's observation ring restarts every update
        torch.manual_seed(50 + upd)
        infos_b = b._sample_training_data()
        b.buffer.prepare_batch_dict()
        _assert_bit_identical(_buffer_snapshot(a), _buffer_snapshot(b), "update %d" % upd)
        assert sorted(map(repr, infos_a)) == sorted(map(repr, infos_b))
        for tr in (a, b):
            torch.manual_seed(80 + upd)
            tr._train_epochs(3e-4, 0.2, 1e-3)
        assert torch.equal(a.model.flat_parameters(), b.model.flat_parameters())
    a.close(exit_process=False)
    b.close(exit_process=False)


# ------------------------------------------------------------------------------------------------ (b) full-size c3 minibatch
def _safe_sample_index(tr, ocfg, P32, n_want, tau, chunk=1024, seed=12):
    """Rows of the flat rollout buffer whose every ReLU pre-activation (3 convs, lin_hidden, embedding, every block's fc,
    both head layers) is at least ``tau`` away from zero in the fp32 oracle forward.  On such samples fp32 rounding noise
    (~1e-6) cannot flip a ReLU decision, so the gradient comparison measures arithmetic accuracy, not ties at the kink
    (one flipped unit changes a weight-gradient row by one sample's whole contribution, ~1e-3..1e-2 of the row)."""
    from buffer import MiniBatch
    from oracle import trxl_oracle as X
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(tr.buffer.batch_size, generator=g)
    keep = []
    for s0 in range(0, perm.numel(), chunk):
        idx = perm[s0:s0 + chunk].to(DEV)
        mb = MiniBatch(tr.buffer, idx)
        window = X.select_window(mb["memories"].cpu(), mb["memory_indices"].cpu())
        X.RELU_MARGINS = []
        try:
            with torch.no_grad():
                X.model_forward(P32, ocfg, mb["obs"].cpu(), window, mb["memory_mask"].cpu(), mb["memory_indices"].cpu())
            margin = torch.stack(X.RELU_MARGINS, dim=0).min(dim=0).values
        finally:
            X.RELU_MARGINS = None
        keep.append(perm[s0:s0 + chunk][margin > tau])
        if sum(k.numel() for k in keep) >= n_want:
            break
    keep = torch.cat(keep)
    assert keep.numel() >= n_want, "only %d of the rollout's samples have ReLU margins > %g" % (keep.numel(), tau)
    return keep[:n_want].to(DEV).contiguous()


def c3_minibatch_errors(verbose=False, safe=True, tau=5e-5):
    """One optimiser step exactly as bench.py times it -- c3 (BASELINE.json configs[2]): W=32, T=512, minibatch N=2048,
    3x84x84 observations read through sample_index, episode table built by a real 512-step rollout -- on the GPU, on the
    fp32 CPU oracle (oracle.ppo_oracle.train_minibatch on the materialised minibatch, as buffer.py:90 gathers it) and on the
    same oracle in float64.  ``safe``: the 2048 rows are drawn from the shuffled buffer but restricted to samples whose ReLU
    decisions are robust (see _safe_sample_index); ``safe=False`` takes the first minibatch of the generator as is.
    Returns (stats triple, forward errors, {param: (scale, |cuda-o32|, |cuda-o64|, |o32-o64|, fraction of entries whose
    parameter differs by > 2e-5, max parameter error)})."""
    import trainer as trainer_mod
    from buffer import MiniBatch
    from device_feed import SyntheticDeviceFeed
    from oracle import ppo_oracle as O
    from oracle import trxl_oracle as X
    from yaml_parser import YamlParser
    cfg = YamlParser(os.path.join(PKG, "configs", "c3_minigrid_synthetic.yaml")).get_config()
    e = cfg["environment"]
    W, T = cfg["n_workers"], cfg["worker_steps"]
    torch.manual_seed(11)
    tr = trainer_mod.PPOTrainer(cfg, run_id="c3", device=torch.device(DEV), workers=[], summary_writer=False)
    tr.device_feed = SyntheticDeviceFeed(W, T, tuple(e["obs_shape"]), e["max_episode_steps"], e["min_episode_steps"], seed=9, device=DEV)
    tr._sample_training_data()
    tr.buffer.prepare_batch_dict()
    assert tr.buffer.memories.shape[0] > W, "the rollout should have finished episodes (deduplicated table rows)"
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    P32 = {k: v.detach().cpu().clone() for k, v in tr.model.state_dict().items()}
    P64 = {k: (v.double() if v.dtype == torch.float32 else v.clone()) for k, v in P32.items()}
    ocfg = dict(cfg, max_episode_steps=e["max_episode_steps"], action_space_shape=(e["n_actions"],))
    if safe:
        idx = _safe_sample_index(tr, ocfg, P32, tr.buffer.mini_batch_size, tau)
        mb = MiniBatch(tr.buffer, idx, idx.cpu())
    else:
        torch.manual_seed(12)
        mb = next(iter(tr.buffer.mini_batch_generator()))
    # as _train_epochs does: sort by episode and build the tile table of the episode-grouped tensor-core attention
    grouping = tr._begin_grouped_attention()
    assert grouping is not None, "c3 (post-LN, relative PE) must take the grouped tensor-core attention path"
    tr._group_epoch([mb], grouping)
    assert mb.groups is not None and mb.groups["n_tiles"] >= grouping["n_episodes"] // 4
    n = mb.sample_index.shape[0]
    assert n == 2048 and tuple(tr.buffer.samples_flat["obs"].shape[1:]) == (3, 84, 84)
    mb_cpu = {k: mb[k].detach().cpu() for k in ("actions", "values", "log_probs", "advantages", "obs", "memory_mask",
                                                   "memory_indices", "memories")}
    lr, clip, beta = 3e-4, 0.1, 1e-3
    stats = torch.zeros(6, device=DEV)
    norms = torch.zeros(tr.model._n_groups + 2, device=DEV)
    tr._ppo_step(mb, lr, clip, beta, stats, norms)
    torch.cuda.synchronize()
    logits, value, out_mem = [t.detach().cpu() for t in tr._train_state[n]["out"]]
    with torch.no_grad():
        window = X.select_window(mb_cpu["memories"], mb_cpu["memory_indices"])
        o_logits, o_value, o_mem = X.model_forward(P32, ocfg, mb_cpu["obs"], window, mb_cpu["memory_mask"], mb_cpu["memory_indices"])
        del window
    fwd = {"logits": float((logits - o_logits[0]).abs().max()), "value": float((value - o_value).abs().max()),
           "new_memory": float((out_mem - o_mem).abs().max()),
           "relu_mismatches_embedding": int(((out_mem[:, 0] > 0) != (o_mem[:, 0] > 0)).sum())}
    s32, g32 = O.train_minibatch(P32, {}, ocfg, mb_cpu, lr, clip, beta)
    mb64 = {k: (v.double() if v.dtype == torch.float32 else v) for k, v in mb_cpu.items()}
    del mb_cpu
    s64, g64 = O.train_minibatch(P64, {}, ocfg, mb64, lr, clip, beta)
    out = {}
    for name, p in tr.model.named_parameters():
        got = p.grad.detach().cpu().double()
        scale = max(1e-30, float(g64[name].abs().max()))
        perr = (p.detach().cpu() - P32[name]).abs()
        out[name] = (scale, float((got - g32[name].double()).abs().max()), float((got - g64[name]).abs().max()),
                     float((g32[name].double() - g64[name]).abs().max()), float((perr > 2e-5).float().mean()), float(perr.max()))
        if verbose:
            print("%-62s scale %.2e  cuda-o32 %.1e  cuda-o64 %.1e  o32-o64 %.1e  (of scale)  param>2e-5: %.1e max %.1e"
                  % (name, scale, out[name][1] / scale, out[name][2] / scale, out[name][3] / scale, out[name][4], out[name][5]))
    tr.close(exit_process=False)
    return (stats.cpu().numpy(), np.array(s32), np.array(s64)), fwd, out


def test_full_c3_minibatch_step_vs_oracle(tmp_path, monkeypatch):
    """Full-size c3 optimiser step (N = 2048 visual samples through sample_index, rollout-built table) against the CPU oracle.
    Forward outputs 1e-4; statistics rtol 1e-4; clipped gradients within 2e-4 of each tensor's max; parameters after
    clip + AdamW 2e-5 (Adam's first step is sign-like, lr * g / (|g| + eps): gradient entries at rounding-noise level may land
    on the other side of zero and move by 2 lr, so at most 0.05 % of a tensor's entries may exceed 2e-5, none by more than 4 lr)."""
    monkeypatch.chdir(tmp_path)
    (stats, s32, s64), fwd, errs = c3_minibatch_errors()
    assert fwd["logits"] <= 1e-4 and fwd["value"] <= 1e-4 and fwd["new_memory"] <= 1e-4, fwd
    np.testing.assert_allclose(stats, s32, rtol=1e-4, atol=1e-6)
    for name, (scale, e32, e64, ref_noise, pfrac, pmax) in errs.items():
        assert e32 <= 2e-4 * scale, (name, e32 / scale, e64 / scale, ref_noise / scale)
        assert pfrac <= 5e-4 and pmax <= 4 * 3e-4, (name, pfrac, pmax)


@pytest.mark.parametrize("heads,gtrxl,pe,ln,dims", [(4, False, "relative", "post", (64, 16, 40, 3)), (2, True, "relative", "post", (64, 16, 40, 3)),
                                                      (8, False, "", "post", (64, 16, 40, 3)),
                                                      (4, True, "relative", "post", (384, 48, 130, 2)),      # c4-like width, M % 32 != 0
                                                      (4, False, "relative", "pre", (64, 16, 40, 3)),       # pre-LN: norm_kv folded into the table
                                                      (1, True, "", "pre", (64, 16, 40, 2))])               # c1-like: one head, pre-LN, no PE
def test_grouped_tensor_core_attention_matches_per_sample_kernel(heads, gtrxl, pe, ln, dims, tmp_path, monkeypatch):
    """The episode-grouped TMA + tcgen05 attention (attention_tc.cu) against the per-sample streaming kernel (attention.cu) on
    the same rollout data: episodes of different lengths (several tiles per episode, partial tiles), step-0 rows that attend
    uniformly, windows that slide.  Same minibatches, same weights: statistics, gradients and parameters after two epochs
    agree to fp32 rounding (the two paths only differ in summation order)."""
    import trainer as trainer_mod
    monkeypatch.chdir(tmp_path)
    D, L, M, B = dims
    cfg = _cfg(n_workers=6, worker_steps=96, n_mini_batch=2, epochs=2,
               environment={"obs_shape": [7], "max_episode_steps": M, "min_episode_steps": 1},
               transformer={"num_heads": heads, "embed_dim": D, "memory_length": L, "num_blocks": B, "gtrxl": gtrxl,
                            "positional_encoding": pe, "layer_norm": ln})
    results = []
    for grouped in ("1", "0"):
        monkeypatch.setenv("TRXL_GROUPED_ATTENTION", grouped)          # "1" also overrides the small-minibatch heuristic
        torch.manual_seed(5)
        tr = trainer_mod.PPOTrainer(cfg, run_id="ga", device=torch.device(DEV), workers=_synthetic_workers(cfg), summary_writer=False)
        torch.manual_seed(6)
        tr._sample_training_data()
        tr.buffer.prepare_batch_dict()
        assert (tr._begin_grouped_attention() is not None) == (grouped == "1")
        torch.manual_seed(7)
        stats, _ = tr._train_epochs(3e-4, 0.2, 1e-3)
        results.append((np.array(stats, dtype=np.float64), tr.model.flat_grads().detach().cpu().clone(),
                        tr.model.flat_parameters().detach().cpu().clone()))
        tr.close(exit_process=False)
    (s1, g1, p1), (s0, g0, p0) = results
    np.testing.assert_allclose(s1, s0, rtol=2e-4, atol=2e-6)
    assert float((g1 - g0).abs().max()) <= 2e-4 * float(g0.abs().max())
    assert float((p1 - p0).abs().max()) <= 4 * 3e-4 and float(((p1 - p0).abs() > 2e-5).float().mean()) <= 2e-3


# ------------------------------------------------------------------------------------------------ f3: checkpoint / enjoy
def test_checkpoint_resume_is_bit_identical(tmp_path, monkeypatch):
    """train 2 updates -> save_checkpoint -> fresh trainer load_checkpoint -> third update's optimisation on the same
    buffer == the uninterrupted run, bit for bit (parameters and AdamW state)."""
    import trainer as trainer_mod
    monkeypatch.chdir(tmp_path)
    cfg = _cfg()
    torch.manual_seed(21)
    a = trainer_mod.PPOTrainer(cfg, run_id="ck", device=torch.device(DEV), workers=_synthetic_workers(cfg), summary_writer=False)
    for upd in range(2):
        torch.manual_seed(300 + upd)
        a._sample_training_data()
        a.buffer.prepare_batch_dict()
        a._train_epochs(3e-4, 0.2, 1e-3)
    a._start_update = 2
    a.save_checkpoint(str(tmp_path / "ck.pt"))
    torch.manual_seed(22)                          # different init: everything must come from the checkpoint
    b = trainer_mod.PPOTrainer(cfg, run_id="ck2", device=torch.device(DEV), workers=_synthetic_workers(cfg), summary_writer=False)
    state = b.load_checkpoint(str(tmp_path / "ck.pt"))
    assert b._start_update == 2 and state["config"] == cfg
    assert torch.equal(a.model.flat_parameters(), b.model.flat_parameters())
    assert b.optimizer.step_count == a.optimizer.step_count
    # third update: same rollout data for both (a's buffer), so only parameters + optimiser state matter
    torch.manual_seed(302)
    a._sample_training_data()
    a.buffer.prepare_batch_dict()
    b.buffer, b._table = a.buffer, a._table
    for tr in (a, b):
        torch.manual_seed(400)
        tr._train_epochs(2e-4, 0.15, 1e-3)
    assert torch.equal(a.model.flat_parameters(), b.model.flat_parameters())
    assert torch.equal(a.optimizer.exp_avg, b.optimizer.exp_avg) and torch.equal(a.optimizer.exp_avg_sq, b.optimizer.exp_avg_sq)
    a.close(exit_process=False)
    b.close(exit_process=False)


def test_saved_model_runs_in_enjoy_and_matches_oracle(tmp_path, monkeypatch):
    """_save_model writes the reference's ``(state_dict, config)`` pickle (trainer.py:356-362); enjoy.py's episode loop
    (enjoy.py:47-84) loads it and steps one episode; value / logits of every step match the CPU oracle run on the same
    pickle with the same actions to 1e-4."""
    import enjoy as enjoy_mod
    import trainer as trainer_mod
    from environments.synthetic_env import SyntheticEnv
    from oracle import trxl_oracle as X
    monkeypatch.chdir(tmp_path)
    cfg = _cfg(environment={"obs_shape": [5]}, transformer={"layer_norm": "pre", "gtrxl": True})
    torch.manual_seed(31)
    tr = trainer_mod.PPOTrainer(cfg, run_id="enj", device=torch.device(DEV), workers=_synthetic_workers(cfg), summary_writer=False)
    tr._sample_training_data()
    tr.buffer.prepare_batch_dict()
    tr._train_epochs(3e-4, 0.2, 1e-3)
    tr._save_model()
    tr.close(exit_process=False)
    path = str(tmp_path / "models" / "enj.nn")
    state_dict, saved_cfg = pickle.load(open(path, "rb"))
    assert saved_cfg == cfg and all(isinstance(v, torch.Tensor) and v.device.type == "cpu" for v in state_dict.values())
    e = cfg["environment"]
    env = SyntheticEnv(tuple(e["obs_shape"]), e["n_actions"], e["max_episode_steps"], e["min_episode_steps"], seed=77)
    torch.manual_seed(5)
    trace = enjoy_mod.run_episode(path, env, device=torch.device(DEV), record=True)
    assert trace["length"] >= e["min_episode_steps"] and len(trace["values"]) == trace["length"]
    # oracle replay on the CPU with the recorded observations
    t = cfg["transformer"]
    ocfg = dict(cfg, max_episode_steps=e["max_episode_steps"], action_space_shape=(e["n_actions"],))
    L, M = t["memory_length"], e["max_episode_steps"]
    memory = torch.zeros((1, M, t["num_blocks"], t["embed_dim"]))
    mask_table, index_table = X.attention_mask_table(L), X.window_index_table(M, L)
    for step, obs in enumerate(trace["obs"]):
        idx = index_table[step].unsqueeze(0)
        window = X.select_window(memory, idx)
        with torch.no_grad():
            logits, value, new_mem = X.model_forward(state_dict, ocfg, torch.from_numpy(obs).unsqueeze(0), window,
                                                     mask_table[min(step, L - 1)].unsqueeze(0).bool(), idx)
        memory[:, step] = new_mem
        np.testing.assert_allclose(trace["values"][step], float(value), atol=1e-4)
        lg = logits[0][0]
        np.testing.assert_allclose(trace["logits"][step], (lg - lg.logsumexp(-1)).numpy(), atol=1e-4)


# ------------------------------------------------------------------------------------------------ (d) 2-GPU NCCL
_NCCL_RANK_SCRIPT = r'''
import os, sys
sys.path.insert(0, {pkg!r}); sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch
import parallel
from parity_util import load_golden, build_model, HEADS
from optim_native import FusedClipAdamW
from trainer import PPOTrainer
parallel.init_from_env("nccl")
rank = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(rank)
device = torch.device("cuda", rank)
for name in ("minibatch_post_rel", "minibatch_pre_learned_gtrxl", "minibatch_post_rel_visual"):
    g = load_golden(name)
    case = name[len("minibatch_"):]
    inputs = {{k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("in.")}}
    n = inputs["obs"].shape[0]
    model, cfg = build_model(g, HEADS[case], inputs["memory_mask"].shape[1], tuple(inputs["obs"].shape[1:]), g["action_shape"],
                             g["max_steps"], device)
    tr = PPOTrainer.__new__(PPOTrainer)
    tr.config, tr.model, tr.device = cfg, model, device
    tr.action_space_shape = tuple(int(a) for a in g["action_shape"]); tr.obs_shape = tuple(inputs["obs"].shape[1:])
    tr.optimizer = FusedClipAdamW(model, lr=3e-4, max_grad_norm=cfg["max_grad_norm"])
    tr.dp = parallel.DataParallelContext(device)
    assert tr.dp.world_size == 2 and tr.dp.native_nccl, "expected the C-ABI NCCL communicator"
    tr._train_state = {{}}
    mine = parallel.shard_workers(n - n % 2, tr.dp.rank, 2)
    lo, hi = mine.start, (mine.stop if tr.dp.rank == 0 else n)          # odd sizes: rank 1 takes the remainder (unequal shards)
    shard = {{k: v[lo:hi] for k, v in inputs.items()}}
    for it in range(2):
        c0 = tr.dp.n_collectives
        stats = tr._train_mini_batch(shard, 3e-4 / (it + 1), 0.2, 1e-3)
        assert tr.dp.n_collectives == c0 + 2          # advantage statistics + ONE gradient/statistics all-reduce
        np.testing.assert_allclose(np.array(stats, dtype=np.float64), g["it%d.stats" % it], rtol=1e-4, atol=1e-5, err_msg=name)
        for pname, p in model.named_parameters():
            want = g["it%d.grad.%s" % (it, pname)]
            scale = max(1e-12, float(np.abs(want).max()))
            np.testing.assert_allclose(p.grad.cpu().numpy(), want, rtol=2e-4, atol=max(2e-6, 1e-4 * scale), err_msg=name + " " + pname)
            assert float(np.abs(p.detach().cpu().numpy() - g["it%d.param.%s" % (it, pname)]).max()) < 2e-5, pname
    # replicas stay bit-identical
    flat = model.flat_parameters().clone(); other = flat.clone()
    torch.distributed.broadcast(other, src=0)
    assert torch.equal(flat, other), "replicas diverged"
# a real trainer: epochs x (1 advantage-statistics all-reduce + n_mini_batch gradient all-reduces), replicas identical
from environments.synthetic_env import SyntheticEnv
class _Pipe:
    def __init__(self, env): self.env, self.q = env, []
    def send(self, msg):
        cmd, data = msg
        self.q.append(self.env.step(data) if cmd == "step" else (self.env.reset() if cmd == "reset" else None))
    def recv(self): return self.q.pop(0)
class _Worker:
    def __init__(self, env): self.child = _Pipe(env)
cfg = {cfg!r}
os.chdir({tmp!r})
workers = [_Worker(SyntheticEnv((5,), 3, 12, 3, seed=1 + w + 100 * rank)) for w in range(cfg["n_workers"])]
torch.manual_seed(1 + rank)
full = PPOTrainer(cfg, run_id="dp%d" % rank, device=device, workers=workers, summary_writer=False)
assert full.dp.native_nccl
p0 = full.model.flat_parameters().clone(); torch.distributed.broadcast(p0, src=0)
assert torch.equal(p0, full.model.flat_parameters()), "initial broadcast missing"
for upd in range(2):
    full._sample_training_data(); full.buffer.prepare_batch_dict()
    c0 = full.dp.n_collectives
    full._train_epochs(3e-4, 0.2, 1e-3)
    assert full.dp.n_collectives - c0 == cfg["epochs"] * (1 + cfg["n_mini_batch"]), full.dp.n_collectives - c0
    p0 = full.model.flat_parameters().clone(); torch.distributed.broadcast(p0, src=0)
    assert torch.equal(p0, full.model.flat_parameters()), "replicas diverged after update %d" % upd
full.close(exit_process=False)
torch.distributed.barrier()
sys.stdout.write("rank %d ok\n" % tr.dp.rank); sys.stdout.flush()
torch.distributed.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_gpu_nccl_gradients_match_single_rank(tmp_path):
    """2 ranks of CUDA kernels + the C-ABI NCCL all-reduce (trxl_allreduce_grads) on the two halves of a minibatch
    reproduce the reference's single-process statistics, gradients and parameters, with ONE collective per optimiser step
    inside the step (the advantage statistics are exchanged before it)."""
    script = tmp_path / "rank.py"
    cfg = _cfg(environment={"obs_shape": [5]})
    script.write_text(_NCCL_RANK_SCRIPT.format(pkg=PKG, root=ROOT, cfg=cfg, tmp=str(tmp_path)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29613", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
