"""CPU-only checks of the host side: the C-ABI library loads and exports exactly what
include/trxl_ppo.h declares, the parameter layout reproduces the reference's state_dict, the integer
tables are bit-exact, the product refuses to run without CUDA, and the multi-rank plumbing works
(world_size 2, gloo).  No kernel is launched here."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import PKG, ROOT, golden_names, load_golden


def _header_functions():
    text = open(os.path.join(ROOT, "include", "trxl_ppo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trxl_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_header_symbols():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    lib_path = entry.build()
    assert os.path.exists(lib_path)
    import trxl_native as native
    native.load()
    declared = _header_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(native.SIGNATURES) == declared, set(native.SIGNATURES) ^ set(declared)
    exported = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    for fn in declared:
        assert re.search(r" T %s$" % fn, exported, flags=re.M), "%s not exported" % fn
    assert native._lib.trxl_abi_version() == 1


def test_sass_is_sm100():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    out = subprocess.run(["cuobjdump", "--list-elf", entry.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:400]


def test_tables_bit_exact():
    from trainer import build_mask_table, build_window_index_table
    g = load_golden("tables")
    for k, v in g.items():
        if k.startswith("mask_L"):
            got = build_mask_table(int(k[6:])).numpy()
            assert got.dtype == v.dtype and np.array_equal(got, v)
        else:
            L, M = (int(s[1:]) for s in k.split("_")[1:])
            got = build_window_index_table(M, L).numpy()
            assert got.dtype == np.int64 and np.array_equal(got, v)
    with pytest.raises(ValueError):
        build_window_index_table(4, 5)


@pytest.mark.parametrize("name", golden_names("forward_"))
def test_layout_matches_reference_state_dict(name):
    from parity_util import HEADS, build_model
    g = load_golden(name)
    case = name[len("forward_"):]
    model, _ = build_model(g, HEADS[case], g["mask"].shape[1], tuple(g["obs"].shape[1:]), g["action_shape"], g["max_steps"], "cpu")
    sd = model.state_dict()
    ref = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    assert set(sd) == set(ref)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref[k].shape), k
        assert np.array_equal(v.numpy(), ref[k]), k                      # load_state_dict wrote through to the arena
    # every parameter is a view of the flat arena, in non-overlapping 16-byte aligned slices
    arena = model.flat_parameters()
    spans = sorted((off, off + numel) for _, off, numel, _ in model._param_slices)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    for pname, p in model.named_parameters():
        assert p.data_ptr() >= arena.data_ptr() and p.data_ptr() < arena.data_ptr() + arena.numel() * 4
        assert p.grad is not None and p.grad.shape == p.shape
    # moving the module repacks the arena and keeps values
    before = {k: v.clone() for k, v in sd.items()}
    model.to(torch.float32)
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k])


def test_no_cpu_fallback():
    import trxl_native as native
    from parity_util import HEADS, build_model
    from trainer import PPOTrainer
    g = load_golden("forward_post_rel")
    model, _ = build_model(g, 2, 4, (5,), g["action_shape"], g["max_steps"], "cpu")
    with pytest.raises(native.NativeLibraryError):
        model(torch.zeros(2, 5), torch.zeros(2, 4, 2, 16), torch.ones(2, 4, dtype=torch.bool), torch.zeros(2, 4, dtype=torch.long))
    with pytest.raises(native.NativeLibraryError):
        PPOTrainer({"n_workers": 1, "learning_rate_schedule": {}, "beta_schedule": {}, "clip_range_schedule": {},
                    "transformer": {"memory_length": 1, "num_blocks": 1, "embed_dim": 4}}, device=torch.device("cpu"))
    with pytest.raises(native.NativeLibraryError):
        native.gae(torch.zeros(1, 1), torch.zeros(1, 1, dtype=torch.uint8), torch.zeros(1, 1), torch.zeros(1), torch.zeros(1, 1),
                   0.99, 0.95)


def test_invalid_configs_are_rejected():
    import trxl_native as native
    bad = native.make_config(30, 4, 1, 4, 16, 3, "post", "relative", False, 8, (2,))     # 30 % 4 != 0
    with pytest.raises(ValueError):
        native.layout(bad)
    bad = native.make_config(32, 5, 1, 4, 16, 3, "post", "relative", False, 8, (2,))     # heads do not divide
    with pytest.raises(ValueError):
        native.layout(bad)
    ok = native.make_config(32, 4, 2, 4, 16, 3, "pre", "learned", True, 8, (2, 3))
    entries, total, groups = native.layout(ok)
    assert groups == 6 + 2 + 2 and total % 4 == 0
    assert native.workspace_floats(ok, 8) > 0


def test_tensor_core_encoder_shape_gate_and_workspace():
    """Host-side part of the tcgen05 encoder API: which observation shapes it covers, and that the workspace grows with
    the batch (the kernels themselves need a GPU; their geometry is pinned in test_tc_conv_geometry.py)."""
    import trxl_native as native
    visual = native.make_config(32, 4, 1, 4, 16, 3136, "post", "relative", False, 8, (2,), conv_in_channels=3)
    assert native.conv_train_supported(visual, 84, 84)
    assert not native.conv_train_supported(visual, 30, 84)                          # too small for the 8/4, 4/2, 3/1 stack
    many = native.make_config(32, 4, 1, 4, 16, 3136, "post", "relative", False, 8, (2,), conv_in_channels=6)
    assert not native.conv_train_supported(many, 84, 84)                            # > 4 channels: cuDNN / im2col path
    vector = native.make_config(32, 4, 1, 4, 16, 3, "post", "relative", False, 8, (2,))
    assert not native.conv_train_supported(vector, 84, 84)
    small, big = native.conv_train_workspace_floats(visual, 32, 84, 84), native.conv_train_workspace_floats(visual, 2048, 84, 84)
    assert 0 < small < big and big * 4 < 4 << 30                                    # c3 minibatch: activations + gradients < 4 GB
    with pytest.raises(ValueError):
        native.conv_train_workspace_floats(many, 32, 84, 84)


def test_utils_and_yaml(tmp_path):
    from utils import polynomial_decay, process_episode_info
    from yaml_parser import YamlParser
    g = load_golden("units")
    got = [polynomial_decay(3e-4, 1e-5, 100, p, s) for p in (1.0, 2.0) for s in (0, 1, 50, 100, 101)]
    assert np.array_equal(np.array(got), g["poly"])
    res = process_episode_info([{"reward": 1.0, "length": 3, "success": True}, {"reward": 3.0, "length": 5, "success": False}])
    assert res["reward_mean"] == 2.0 and res["success_percent"] == 0.5 and res["length_std"] == 1.0
    assert process_episode_info([]) == {}
    cfg = YamlParser(os.path.join(PKG, "configs", "c3_minigrid_synthetic.yaml")).get_config()
    assert cfg["transformer"]["memory_length"] == 128 and cfg["n_workers"] == 32 and isinstance(cfg, dict)


def test_minibatch_is_lazy_and_reference_shaped():
    """The buffer yields the reference's keys; nothing is gathered until a key is read."""
    from buffer import Buffer, MiniBatch

    class Space:
        shape = (3,)
    cfg = {"n_workers": 2, "worker_steps": 6, "n_mini_batch": 3,
           "transformer": {"memory_length": 4, "num_blocks": 1, "embed_dim": 8}}
    buf = Buffer(cfg, Space(), (2,), 7, torch.device("cpu"))
    buf.memories = [torch.zeros(7, 1, 8), torch.ones(7, 1, 8)]
    buf.memory_index[1] = 1
    buf.values[:] = torch.arange(12.0).reshape(2, 6)
    buf.prepare_batch_dict()
    assert buf.memories.shape == (2, 7, 1, 8) and buf.samples_flat["values"].shape == (12,)
    torch.manual_seed(0)
    batches = list(buf.mini_batch_generator())
    torch.manual_seed(0)
    perm = torch.randperm(12)
    assert len(batches) == 3 and all(isinstance(b, MiniBatch) for b in batches)
    assert torch.equal(torch.cat([b.sample_index for b in batches]), perm)          # same index stream as torch.randperm
    mb = batches[0]
    assert dict.__len__(mb) == 0                                                    # nothing materialised yet
    assert set(mb.keys()) == {"actions", "values", "log_probs", "advantages", "obs", "memory_mask", "memory_indices", "memories"}
    assert torch.equal(mb["values"], buf.samples_flat["values"][mb.sample_index])
    assert mb["memories"].shape == (4, 7, 1, 8)
    assert torch.equal(mb["memories"][:, 0, 0, 0], (mb.sample_index >= 6).float())


_RANK_SCRIPT = r'''
import os, sys
sys.path.insert(0, {pkg!r}); sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch
import torch.distributed as dist
import parallel
from oracle import ppo_oracle as O, trxl_oracle as X
from parity_util import load_golden, config_from_golden
parallel.init_from_env(backend="gloo")
dp = parallel.DataParallelContext()
assert dp.world_size == 2
g = load_golden("minibatch_post_rel")
cfg, sd = config_from_golden(g, 2, 4)
cfg.update(max_episode_steps=int(g["max_steps"]), action_space_shape=tuple(int(a) for a in g["action_shape"]))
P = {{k: torch.from_numpy(v.copy()) for k, v in sd.items()}}
mb = {{k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("in.")}}
n = mb["obs"].shape[0]
mine = parallel.shard_workers(n, dp.rank, dp.world_size)
shard = {{k: v[mine.start:mine.stop] for k, v in mb.items()}}
# global advantage statistics: all-reduce (sum, sum of squares, count), as the native path does
a = shard["advantages"].double()
st = torch.tensor([a.sum(), (a * a).sum(), float(len(a))], dtype=torch.float64)
dp.all_reduce_(st)
mean = st[0] / st[2]; std = torch.sqrt((st[1] - st[2] * mean * mean) / (st[2] - 1))
names = X.trainable_names(P)
for k in names: P[k].requires_grad_(True)
window = X.select_window(shard["memories"], shard["memory_indices"])
logits, value, _ = X.model_forward(P, cfg, shard["obs"], window, shard["memory_mask"], shard["memory_indices"])
# same loss as ppo_loss but normalised by the GLOBAL batch so that summing rank gradients is exact
nadv = ((shard["advantages"] - mean.float()) / (std.float() + 1e-8)).unsqueeze(1)
lp = torch.stack([X.categorical_log_prob(lg, shard["actions"][:, i]) for i, lg in enumerate(logits)], 1)
ent = torch.stack([X.categorical_entropy(lg) for lg in logits], 1).sum(1)
ratio = torch.exp(lp - shard["log_probs"])
pl = torch.min(ratio * nadv, torch.clamp(ratio, 0.8, 1.2) * nadv).sum() / (n * lp.shape[1])
ret = shard["values"] + shard["advantages"]
cv = shard["values"] + (value - shard["values"]).clamp(-0.2, 0.2)
vl = torch.max((value - ret) ** 2, (cv - ret) ** 2).sum() / n
loss = -(pl - 0.25 * vl + 1e-3 * ent.sum() / n)
loss.backward()
flat = torch.cat([P[k].grad.reshape(-1) for k in names])
dp.all_reduce_(flat)                                  # ONE collective for the whole gradient arena
coef = min(1.0, 0.5 / (float(torch.linalg.vector_norm(flat)) + 1e-6))
off = 0
for k in names:
    got = flat[off:off + P[k].numel()].reshape(P[k].shape).numpy() * coef
    off += P[k].numel()
    want = g["it0.grad." + k]
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-4 * max(1e-12, float(np.abs(want).max())), err_msg=k)
dist.barrier()
sys.stdout.write("rank %d ok\n" % dp.rank); sys.stdout.flush()
'''


def test_two_rank_gradient_allreduce_matches_single_rank(tmp_path):
    """world_size 2 over gloo: each rank differentiates its half of the minibatch (oracle arithmetic),
    statistics and the flat gradient buffer are all-reduced exactly as the native trainer does, and the
    result equals the reference's single-process gradient."""
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(pkg=PKG, root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29611", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


@pytest.mark.parametrize("blocking", [False, True, "futex"])
def test_shared_memory_stepping_modes(blocking):
    """The rollout's env transport (worker.py): spinning and semaphore-blocking workers speak the same protocol --
    actions in, (reward, done) out through shared arrays, observation into the shared slab, auto-reset + info on the pipe."""
    import time
    from worker import Worker, make_control
    n, obs_shape = 3, (3, 8, 8)
    slab = torch.zeros((n,) + obs_shape, dtype=torch.float32).share_memory_()
    from worker import FUTEX_WORD_STRIDE, _Futex, futex_available
    if blocking == "futex" and not futex_available():
        pytest.skip("no shared-memory futex on this platform")
    control = make_control(n, 1, blocking=bool(blocking), futex=(blocking == "futex"))
    assert (control["futex"] is not None) == (blocking == "futex")
    fut = _Futex() if blocking == "futex" else None
    cfg = {"type": "Synthetic", "obs_shape": list(obs_shape), "n_actions": 4, "max_episode_steps": 5, "min_episode_steps": 2, "seed": 0}
    workers = [Worker(dict(cfg, seed=w), slab, w, control, group=w % 2) for w in range(n)]
    try:
        for w in workers:
            w.child.send(("reset", None))
        for w in workers:
            assert w.child.recv() is None            # the observation went to the slab
        first = slab.clone()
        assert float(first.abs().sum()) > 0
        infos = 0
        for step in range(12):
            control["actions"].numpy()[...] = step % 4
            control["cmd"].numpy()[...] += 1
            if fut is not None:
                for grp in (0, 1):
                    control["futex"].numpy()[FUTEX_WORD_STRIDE * grp] += 1
                    fut.wake(control["futex"].data_ptr() + 4 * FUTEX_WORD_STRIDE * grp)
            elif control["sems"]:
                for sem in control["sems"]:
                    sem.release()
            deadline = time.time() + 20
            while not np.array_equal(control["ack"].numpy(), control["cmd"].numpy()):
                assert time.time() < deadline, "workers did not acknowledge"
            for w in np.nonzero(control["has_info"].numpy())[0]:
                assert control["dones"].numpy()[w] == 1
                assert isinstance(workers[int(w)].child.recv(), dict)
                infos += 1
        assert infos >= n                            # max_episode_steps = 5: every env finished at least once in 12 steps
        assert not torch.equal(first, slab)
    finally:
        for w in workers:
            w.child.send(("close", None))
        for w in workers:
            w.child.recv()
            w.process.join(5)


_REF_LOAD_SCRIPT = r'''
import os, pickle, sys
sys.path.insert(0, os.path.join({root!r}, "tests", "golden"))
import make_golden as mg
mg.install_stubs()
sys.path.insert(0, {ref!r})
import numpy as np, torch
from model import ActorCriticModel            # the REFERENCE's model.py
assert os.path.realpath(sys.modules["model"].__file__).startswith(os.path.realpath({ref!r}))
state_dict, config = pickle.load(open({path!r}, "rb"))
class Space:
    def __init__(self, shape): self.shape = shape
for obs_shape in ({obs_shape!r},):
    ref = ActorCriticModel(config, Space(obs_shape), (3,), 12)
    missing = ref.load_state_dict(state_dict, strict=True)      # same keys and shapes, or this raises
    L, t = config["transformer"]["memory_length"], config["transformer"]
    torch.manual_seed(0)
    obs = torch.rand((2,) + tuple(obs_shape)); mem = torch.randn(2, L, t["num_blocks"], t["embed_dim"])
    mask = torch.tril(torch.ones(L, L), -1)[[0, L - 1]].bool(); idx = torch.arange(L).repeat(2, 1)
    with torch.no_grad():
        pi, value, new_mem = ref(obs, mem, mask, idx)
    np.savez({out!r}, obs=obs.numpy(), mem=mem.numpy(), mask=mask.numpy(), idx=idx.numpy(), value=value.numpy(),
             new_mem=new_mem.numpy(), logits=pi[0].logits.numpy())
print("reference loaded the checkpoint")
'''


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the unmodified reference is only present in the build container")
@pytest.mark.parametrize("obs_shape,ln,gtrxl,pe", [((5,), "pre", True, "learned"), ((3, 84, 84), "post", False, "relative")])
def test_saved_model_loads_into_the_reference_model(tmp_path, obs_shape, ln, gtrxl, pe):
    """f3: the ``(state_dict, config)`` pickle written by this engine (trainer.save_model_file, the body of _save_model)
    loads with strict=True into the UNMODIFIED reference ActorCriticModel (reference enjoy.py:47-57), and the reference's
    forward on that checkpoint equals the oracle's (which the GPU tests compare the CUDA path with) to 1e-5."""
    from model import ActorCriticModel
    from oracle import trxl_oracle as X
    from parity_util import _Space
    from trainer import save_model_file
    cfg = {"hidden_layer_size": 32, "transformer": {"num_blocks": 2, "embed_dim": 32, "num_heads": 4, "memory_length": 6,
                                                    "positional_encoding": pe, "layer_norm": ln, "gtrxl": gtrxl, "gtrxl_bias": 0.5}}
    torch.manual_seed(1)
    model = ActorCriticModel(cfg, _Space(obs_shape), (3,), 12)
    path, out = str(tmp_path / "m.nn"), str(tmp_path / "ref_out.npz")
    save_model_file(model, cfg, path)
    script = tmp_path / "load_ref.py"
    script.write_text(_REF_LOAD_SCRIPT.format(root=ROOT, ref="/root/reference", path=path, out=out, obs_shape=tuple(obs_shape)))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    g = dict(np.load(out))
    P = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ocfg = dict(cfg, max_episode_steps=12, action_space_shape=(3,))
    mem = torch.from_numpy(g["mem"])
    with torch.no_grad():
        logits, value, new_mem = X.model_forward(P, ocfg, torch.from_numpy(g["obs"]), mem, torch.from_numpy(g["mask"]),
                                                 torch.from_numpy(g["idx"]))
    np.testing.assert_allclose(value.numpy(), g["value"], atol=1e-5)
    np.testing.assert_allclose(new_mem.numpy(), g["new_mem"], atol=1e-5)
    lg = logits[0]
    np.testing.assert_allclose((lg - lg.logsumexp(-1, keepdim=True)).numpy(), g["logits"], atol=1e-5)


@pytest.mark.parametrize("heads,n_rows,n_episodes", [(4, 2048, 70), (8, 500, 3), (1, 64, 64), (2, 129, 1)])
def test_episode_grouping_tiles_cover_every_sample_once(heads, n_rows, n_episodes):
    """Host side of the episode-grouped attention (trainer.group_minibatch_by_episode): the sorted minibatch is a permutation of
    the input rows, every tile holds at most 128 (sample, head) rows of ONE episode, the tiles partition the rows in order, and
    their number stays within the padded table length used for CUDA-graph replay."""
    from trainer import group_minibatch_by_episode, tile_table_length
    rng = np.random.default_rng(heads * 1000 + n_rows)
    total = 5 * n_rows
    episode_of_row = rng.integers(0, n_episodes, total).astype(np.int64)
    sample_index = rng.permutation(total)[:n_rows].astype(np.int64)
    idx, tiles = group_minibatch_by_episode(sample_index, episode_of_row, heads)
    assert sorted(idx.tolist()) == sorted(sample_index.tolist())
    ep_sorted = episode_of_row[idx]
    assert np.all(np.diff(ep_sorted) >= 0)
    assert tiles.dtype == np.int32 and tiles.shape[1] == 4
    nxt = 0
    for first, rows, ep, pad in tiles.tolist():
        assert first == nxt and 0 < rows <= 128 and rows % heads == 0 and pad == 0
        members = ep_sorted[first // heads:(first + rows) // heads]
        assert np.all(members == ep)
        nxt = first + rows
    assert nxt == n_rows * heads
    assert len(tiles) <= tile_table_length(n_rows, heads, n_episodes)
    # the padded length only moves when the episode count crosses a multiple of 64
    assert tile_table_length(n_rows, heads, 65) == tile_table_length(n_rows, heads, 128)
