"""PPO trainer over an episodic TransformerXL memory -- B200-native drop-in for the reference's
``trainer.PPOTrainer`` (trainer.py:16-383).  ``PPOTrainer(config, run_id, device)``,
``run_training()``, ``close()`` and the internal methods BASELINE.json names
(``_sample_training_data``, ``get_last_value``, ``_train_epochs``, ``_train_mini_batch``) keep their
signatures and semantics; ``train.py`` drives it unchanged.

How the hot path is mapped onto the GPU (all arithmetic in libtrxlppo, see csrc/):
  * episodic memory is one device table (E, M, B, D); a worker's live episode is a row of it, so
    "clone the finished episode, zero the worker memory, append a new episode" (trainer.py:205-213) is
    just "point the worker at a fresh zero row";
  * rollout step: obs H2D -> rollout_prepare (mask / window-index rows) -> model trunk (window read
    in place from the table) -> memory_scatter -> sample_actions -> actions D2H;
  * PPO step: sample_index (the shuffled rows) -> obs gather (+ cuDNN conv encoder) -> trunk forward
    -> adv_stats -> fused loss fwd/bwd -> trunk backward -> conv backward -> [all-reduce] ->
    fused clip + AdamW.  No (mb, M, B, D) / (mb, L, B, D) tensors, no host sync inside an update;
  * multi-GPU: one process per GPU, each with its own workers/buffer/table; gradients are summed with
    a single all-reduce of the flat gradient arena per optimiser step (parallel.py).
"""
import os
import pickle
import time
from collections import deque

import numpy as np
import torch

import trxl_native as native
from buffer import Buffer, MiniBatch
from model import ActorCriticModel
from optim_native import FusedClipAdamW
from parallel import DataParallelContext
from utils import create_env, polynomial_decay, process_episode_info
from worker import FUTEX_WORD_STRIDE, Worker, _Futex, make_control, physical_cpus


def save_model_file(model, config, path):
    """Write the reference's checkpoint format: ``pickle.dump((state_dict, config))`` with CPU tensors and the
    reference's state_dict keys/shapes (reference trainer.py:356-362; read by enjoy.py:47-57)."""
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    with open(path, "wb") as f:
        pickle.dump((state, config), f)


def effective_cpus():
    """CPUs this process tree may actually burn: the affinity mask capped by the cgroup CPU quota (cpu.max /
    cfs_quota_us).  Busy-waiting env workers beyond this number get the whole container throttled by CFS."""
    try:
        n = float(len(os.sched_getaffinity(0)))
    except AttributeError:
        n = float(os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, float(quota) / float(period))
    except (OSError, ValueError):
        try:
            quota = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0:
                n = min(n, quota / period)
        except (OSError, ValueError):
            pass
    return n


class _WorkerGroup:
    """A contiguous slice [lo, hi) of the env workers that steps as a unit: its own CUDA stream, staging buffers,
    per-step CUDA graphs and completion event, so one group's forward runs while another group's environments step."""
    GPU, ENV, DONE = 0, 1, 2

    def __init__(self, index, lo, hi, trainer):
        dev, nb = trainer.device, len(trainer.action_space_shape)
        self.index, self.lo, self.hi, self.n = index, lo, hi, hi - lo
        self.stream = torch.cuda.Stream(device=dev)
        self.event = torch.cuda.Event()
        self.obs_dev = torch.zeros((self.n,) + trainer.obs_shape, dtype=torch.float32, device=dev)
        self.act_dev = torch.zeros((self.n, nb), dtype=torch.long, device=dev)
        self.act_pinned = torch.zeros((self.n, nb), dtype=torch.long).pin_memory()
        self.act_host_dptr = native.host_device_pointer(self.act_pinned.data_ptr())     # kernels write the actions here
        self.act_np = self.act_pinned.numpy()
        # completion signal: the sampling kernel tags every action word it writes into pinned host memory with a launch sequence
        # number (word = seq << 32 | action), so the host polls plain int64s instead of a CUDA event and no system-wide fence sits
        # on the step's critical path (W * branches <= 1024 per group, else the event is used)
        self.done_counter = torch.zeros(1, dtype=torch.long, device=dev)
        self.use_flag = self.n * nb <= 1024
        self.expected = 0
        self.stream_handle = self.stream.cuda_stream
        self.step_dev = torch.zeros(self.n, dtype=torch.long, device=dev)
        self.ep_dev = torch.zeros(self.n, dtype=torch.long, device=dev)
        self.rows = None            # (T, n) int64: flat buffer rows of this group's workers at each step
        self.ws = self.outs = None
        self.graphs = {}
        self.t, self.phase = 0, self.DONE


def build_mask_table(memory_length):
    """(L, L) strictly-lower-triangular float table (trainer.py:78); row min(step, L-1) is a sample's mask."""
    return torch.tril(torch.ones((memory_length, memory_length)), diagonal=-1)


def group_minibatch_by_episode(sample_index, episode_of_row, num_heads):
    """Host side of the episode-grouped attention: sort a minibatch's buffer rows by episode (stable) and cut them into tiles of at
    most 128 (sample, head) rows that belong to ONE episode.  Returns (sorted rows (n,) int64, tiles (n_tiles, 4) int32 with
    entries {first (sample, head) row, number of rows, episode, 0})."""
    spt = 128 // num_heads                                           # samples per 128-row tile
    ep = episode_of_row[sample_index]
    order = np.argsort(ep, kind="stable")
    idx, ep = sample_index[order], ep[order]
    bounds = np.flatnonzero(np.diff(ep)) + 1
    starts = np.concatenate(([0], bounds)).tolist() if len(ep) else []
    ends = np.concatenate((bounds, [len(ep)])).tolist() if len(ep) else []
    rows = []
    for s0, s1 in zip(starts, ends):
        e = int(ep[s0])
        for r in range(s0, s1, spt):
            rows.append((r * num_heads, (min(s1, r + spt) - r) * num_heads, e, 0))
    return idx, np.asarray(rows, dtype=np.int32).reshape(-1, 4)


def tile_table_length(rows, num_heads, n_episodes):
    """Length every minibatch's tile table is padded to (empty entries have zero rows): an upper bound of the tile count --
    an episode with r rows makes at most r * H / 128 + 1 tiles -- that only changes when the episode count crosses a multiple
    of 64, so that a captured optimiser step stays replayable across minibatches and updates."""
    return rows * num_heads // 128 + 1 + (n_episodes + 63) // 64 * 64


def build_window_index_table(max_episode_length, memory_length):
    """(M, L) int64 window slots by episode step (trainer.py:88-90)."""
    if memory_length > max_episode_length:
        raise ValueError("memory_length (%d) must not exceed the environment's max_episode_steps (%d)"
                         % (memory_length, max_episode_length))
    head = torch.arange(memory_length).repeat(memory_length - 1, 1)
    tail = torch.arange(max_episode_length - memory_length + 1).unsqueeze(1) + torch.arange(memory_length).unsqueeze(0)
    return torch.cat((head, tail)).long()


class PPOTrainer:
    def __init__(self, config, run_id="run", device=torch.device("cpu"), workers=None, summary_writer=True):
        self.config = config
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise native.NativeLibraryError(
                "PPOTrainer needs a CUDA device: the B200 engine has no CPU path (got device=%s). "
                "Use the reference implementation for CPU runs." % self.device)
        native.load()
        self.run_id = run_id
        self.num_workers = config["n_workers"]
        self.lr_schedule = config["learning_rate_schedule"]
        self.beta_schedule = config["beta_schedule"]
        self.cr_schedule = config["clip_range_schedule"]
        t = config["transformer"]
        self.memory_length, self.num_blocks, self.embed_dim = t["memory_length"], t["num_blocks"], t["embed_dim"]
        self.dp = DataParallelContext(self.device)

        self.writer = None
        if summary_writer and self.dp.rank == 0:
            from torch.utils.tensorboard import SummaryWriter
            os.makedirs("./summaries", exist_ok=True)
            self.writer = SummaryWriter("./summaries/" + run_id + time.strftime("/%Y%m%d-%H%M%S/"))

        dummy_env = create_env(self._env_config(-1))
        observation_space = dummy_env.observation_space
        self.action_space_shape = (dummy_env.action_space.n,)
        self.max_episode_length = dummy_env.max_episode_steps
        dummy_env.close()
        self.obs_shape = tuple(observation_space.shape)

        self.buffer = Buffer(config, observation_space, self.action_space_shape, self.max_episode_length, self.device)
        self.model = ActorCriticModel(config, observation_space, self.action_space_shape, self.max_episode_length).to(self.device)
        self.model.train()
        self.dp.broadcast_(self.model.flat_parameters())
        self.optimizer = FusedClipAdamW(self.model, lr=self.lr_schedule["initial"], max_grad_norm=config["max_grad_norm"])

        # env workers: anything with a ``child`` pipe end speaking the reference protocol
        # Own workers write observations into a shared-memory slab that is also registered as pinned host memory,
        # so the per-step image payload goes env process -> slab -> (DMA) GPU without pickling or staging copies.
        # (pass workers=[] together with trainer.device_feed = SyntheticDeviceFeed(...) to run without env processes)
        self._obs_slab = None
        self._yield_when_idle = False
        self._control = None          # shared-memory stepping arrays (own workers only)
        # worker groups overlap one group's env stepping with the others' forwards; the forwards are latency-bound (a group's
        # kernels occupy a fraction of the GPU), so several groups' device work overlaps too (measured at c3, W = 32 on 16 host
        # CPUs: rollout 215 ms with 2 groups, 161 with 3, 158 with 4, 146 with 6, 149 with 8)
        own_stepping = workers is None and os.environ.get("TRXL_PIPE_STEPPING", "0") != "1"
        default_groups = 1 if not own_stepping or self.num_workers < 8 else min(6, self.num_workers // 4)
        n_groups = int(os.environ.get("TRXL_ROLLOUT_GROUPS", str(default_groups)))
        n_groups = max(1, min(n_groups, self.num_workers, 64))
        self._group_bounds = [round(i * self.num_workers / n_groups) for i in range(n_groups + 1)]
        group_of = [max(g for g in range(n_groups) if self._group_bounds[g] <= w) for w in range(self.num_workers)]
        if workers is None:
            self._obs_slab = torch.zeros((self.num_workers,) + self.obs_shape, dtype=torch.float32).share_memory_()
            procs = self.dp.world_size * (self.num_workers + 1)
            if os.environ.get("TRXL_PIPE_STEPPING", "0") != "1":
                # Spinning workers answer fastest but every one of them burns a CPU all the time: they are only used when
                # all env processes of all ranks (+ the trainers) fit the CPUs this container may use (affinity capped by
                # the cgroup quota -- measured on the B200 host: 33 spinners under a 16-CPU quota get CFS-throttled), and
                # only on x86 (the handshake publishes with plain stores and relies on total store order).  Otherwise the
                # workers sleep on a semaphore between steps.
                import platform
                can_spin = procs + 1 <= effective_cpus() and platform.machine() in ("x86_64", "AMD64")
                self._yield_when_idle = procs + 1 > effective_cpus() and hasattr(os, "sched_yield")
                spin = os.environ.get("TRXL_SPIN_STEPPING", "1" if can_spin else "0") == "1"
                self._control = make_control(self.num_workers, len(self.action_space_shape), blocking=not spin)
            # one physical core per env worker when the box has enough of them (TRXL_PIN_WORKERS=0 leaves placement to the OS)
            cores = physical_cpus() if os.environ.get("TRXL_PIN_WORKERS", "1") == "1" else []
            if len(cores) < procs:
                cores = []
            first = 1 + self.dp.rank * (self.num_workers + 1)          # core `first - 1` is left to this rank's main thread
            workers = [Worker(self._env_config(w), self._obs_slab, w, self._control,
                              cpu=cores[(first + w) % len(cores)] if cores else None, group=group_of[w])
                       for w in range(self.num_workers)]
            rc = torch.cuda.cudart().cudaHostRegister(self._obs_slab.data_ptr(), self._obs_slab.numel() * 4, 0)
            self._slab_pinned = (int(rc) == 0) if not isinstance(rc, tuple) else (int(rc[0]) == 0)
        self.workers = workers
        self.worker_ids = range(self.num_workers)
        self.worker_current_episode_step = torch.zeros((self.num_workers,), dtype=torch.long)
        for worker in self.workers:
            worker.child.send(("reset", None))
        self.obs = self._obs_slab.numpy() if self._obs_slab is not None else np.zeros((self.num_workers,) + self.obs_shape,
                                                                                      dtype=np.float32)
        for w, worker in enumerate(self.workers):
            first = worker.child.recv()
            if first is not None:
                self.obs[w] = first
        self._feed_last_obs = None

        # bit-exact integer tables, built on the host and uploaded once
        self.memory_mask = build_mask_table(self.memory_length)
        self.memory_indices = build_window_index_table(self.max_episode_length, self.memory_length)
        dev, W, T, L = self.device, self.num_workers, config["worker_steps"], self.memory_length
        self._mask_table_dev = (self.memory_mask != 0).to(torch.uint8).to(dev).contiguous()
        self._index_table_dev = self.memory_indices.to(dev).contiguous()

        # episode table + per-worker cursors
        self._table = None
        self._table_roll = None         # rollout: table + positional rows, kept in step by trxl_rollout_store (fused forward only)
        self._table_cap = 0
        self._ep_host = torch.arange(W, dtype=torch.long).pin_memory()      # table row of each worker's live episode
        self._step_host = self.worker_current_episode_step.pin_memory()
        self.worker_current_episode_step = self._step_host
        self._ep_dev = torch.arange(W, dtype=torch.long, device=dev)
        self._step_dev = torch.zeros(W, dtype=torch.long, device=dev)
        self._alloc_table(self._initial_capacity())
        self._n_rows = W              # rows of the table in use
        self._n_episodes = W          # episodes the reference's buffer.memories would hold

        # staging buffers
        self._obs_pinned = torch.zeros((W,) + self.obs_shape, dtype=torch.float32).pin_memory()
        self._obs_dev = torch.zeros((W,) + self.obs_shape, dtype=torch.float32, device=dev)
        self._act_dev = torch.zeros((W, len(self.action_space_shape)), dtype=torch.long, device=dev)
        self._act_pinned = torch.zeros((W, len(self.action_space_shape)), dtype=torch.long).pin_memory()
        self._rollout_rows = (torch.arange(W, device=dev).unsqueeze(0) * T + torch.arange(T, device=dev).unsqueeze(1)).contiguous()
        self._win_last = torch.zeros((W, L), dtype=torch.long, device=dev)
        self._mask_last = torch.zeros((W, L), dtype=torch.uint8, device=dev)
        self._train_state = {}
        self._ctx = None
        # worker groups: with shared-memory stepping the workers are split into groups that alternate between the GPU (forward +
        # sampling on the group's own stream) and the environments, so env stepping overlaps the other group's device work
        bounds = self._group_bounds if self._control is not None else [0, W]
        self._groups = [_WorkerGroup(i, bounds[i], bounds[i + 1], self) for i in range(len(bounds) - 1) if bounds[i + 1] > bounds[i]]
        self._whole = self._groups[0] if len(self._groups) == 1 else _WorkerGroup(len(self._groups), 0, W, self)
        for grp in self._groups + [self._whole]:
            grp.rows = self._rollout_rows[:, grp.lo:grp.hi].contiguous()
        self.use_cuda_graphs = os.environ.get("TRXL_NO_GRAPHS", "0") != "1"
        # the optimiser step's two launch-dense segments (encoder + trunk forward; trunk + encoder backward: ~250 launches of
        # 3-30 us) replayed as CUDA graphs: eager stream launches leave ~2.6 us between kernels, graph replays ~0.4 us
        self.use_train_graphs = self.use_cuda_graphs and os.environ.get("TRXL_TRAIN_GRAPHS", "1") != "0"
        self._train_graphs = {}
        self._capture_stream = None
        self._mapped = {}
        self._futex = None
        self._table_pe = None           # table + positional rows for the episode-grouped tensor-core attention
        self._start_update = 0          # first update of run_training (advanced by load_checkpoint)
        self.device_feed = None         # optional device_feed.SyntheticDeviceFeed replacing the env workers (bench.py)
        self._forced_buf = None         # persistent (T, W, n_branches) int64 device buffer behind ``_forced_actions``
        self._forced_on = False
        self._graph_warm_rollouts = 1   # rollouts run eagerly before the per-step graphs are captured (lazy initialisation)
        self.timers = {"rollout": 0.0, "train": 0.0, "env": 0.0}

    # ------------------------------------------------------------------------------------------ setup helpers
    @property
    def _forced_actions(self):
        """Optional (T, W, n_branches) int64 device tensor of actions to replay instead of sampling (parity tests).
        Assigning copies into a persistent buffer, so CUDA graphs captured with forced actions stay valid when the
        next update's actions are assigned."""
        return self._forced_buf if self._forced_on else None

    @_forced_actions.setter
    def _forced_actions(self, value):
        if value is None:
            self._forced_on = False
            return
        value = value.to(self.device, torch.long)
        if self._forced_buf is None or self._forced_buf.shape != value.shape:
            self._forced_buf = torch.empty_like(value).contiguous()
        self._forced_buf.copy_(value)
        self._forced_on = True

    def _env_config(self, worker):
        cfg = dict(self.config["environment"])
        if cfg.get("type") == "Synthetic":
            cfg["seed"] = int(cfg.get("seed", 0)) + 1 + worker + 10007 * getattr(getattr(self, "dp", None), "rank", 0)
        return cfg

    def _initial_capacity(self):
        W, T = self.num_workers, self.config["worker_steps"]
        return W + max(8, (W * T) // max(1, self.max_episode_length // 2))

    def _alloc_table(self, capacity):
        """(Re)allocate the episode table.  Growing it mid-rollout is rare (the capacity doubles and persists): worker groups
        may have steps in flight on their own streams, so the move is fenced by device-wide synchronisation."""
        if self._table is not None:
            torch.cuda.synchronize(self.device)
        new = torch.zeros((capacity, self.max_episode_length, self.num_blocks, self.embed_dim), dtype=torch.float32,
                          device=self.device)
        if self._table is not None:
            new[:self._table.shape[0]].copy_(self._table)
            torch.cuda.synchronize(self.device)
        self._table, self._table_cap = new, capacity
        # the fused rollout forward prefetches window rows with bulk copies, which cannot add positional rows on the way: with a
        # (parameter-free) positional table the rollout keeps a second table that already carries them
        model = getattr(self, "model", None)
        if model is not None and model._fused_ok and self.num_workers <= model.FUSED_MAX_BATCH and model._pe_table() is not None \
                and os.environ.get("TRXL_ROLLOUT_TABLE_PE", "1") != "0":
            # (rows no step has written yet must read as 0 + positional row: an episode's first step attends uniformly over them)
            self._table_roll = torch.empty_like(new)
            native.table_add_pe(new, model._pe_table(), self._table_roll, capacity)
            torch.cuda.synchronize(self.device)

    @property
    def memory(self):
        """(W, M, B, D) live episodic memory of every worker (the reference's ``self.memory``)."""
        return self._table[self._ep_dev]

    def _new_row(self):
        if self._n_rows >= self._table_cap:
            self._alloc_table(self._table_cap * 2)
        self._n_rows += 1
        return self._n_rows - 1

    def _begin_rollout(self):
        """Reference trainer.py:154-156: the buffer's episode list restarts as the W live memories."""
        W = self.num_workers
        live = self._table[self._ep_dev]                 # (W, M, B, D) copy
        self._table.zero_()
        self._table[:W].copy_(live)
        self._ep_host.copy_(torch.arange(W))
        self._ep_dev.copy_(self._ep_host, non_blocking=True)
        self._n_rows = W
        self._n_episodes = W
        if self._table_roll is not None:                 # every row with its positional row (fresh rows: 0 + pe); the rollout
            native.table_add_pe(self._table, self.model._pe_table(), self._table_roll, self._table_cap)     # then writes slot by slot

    # ------------------------------------------------------------------------------------------ training loop
    def run_training(self):
        """Sample -> prepare -> optimise for ``config["updates"]`` updates; saves the final model (trainer.py:101-143)."""
        if self.dp.rank == 0:
            print("Starting training on %s (%d rank%s)" % (self.device, self.dp.world_size, "" if self.dp.world_size == 1 else "s"))
        episode_infos = deque(maxlen=100)
        for update in range(self._start_update, self.config["updates"]):
            self._start_update = update + 1
            lr = polynomial_decay(self.lr_schedule["initial"], self.lr_schedule["final"], self.lr_schedule["max_decay_steps"],
                                  self.lr_schedule["power"], update)
            beta = polynomial_decay(self.beta_schedule["initial"], self.beta_schedule["final"],
                                    self.beta_schedule["max_decay_steps"], self.beta_schedule["power"], update)
            clip_range = polynomial_decay(self.cr_schedule["initial"], self.cr_schedule["final"],
                                          self.cr_schedule["max_decay_steps"], self.cr_schedule["power"], update)
            sampled = self._sample_training_data()
            self.buffer.prepare_batch_dict()
            training_stats, grad_info = self._train_epochs(lr, clip_range, beta)
            training_stats = np.mean(training_stats, axis=0)
            episode_infos.extend(sampled)
            episode_result = process_episode_info(episode_infos)
            if self.dp.rank == 0:
                self._report(update, training_stats, episode_result)
                self._write_gradient_summary(update, grad_info)
                self._write_training_summary(update, training_stats, episode_result)
        if self.dp.rank == 0:
            self._save_model()

    def _report(self, update, s, ep):
        head = "{:4}".format(update)
        if ep:
            head += " reward={:.2f} std={:.2f} length={:.1f} std={:.2f}".format(ep["reward_mean"], ep["reward_std"],
                                                                               ep["length_mean"], ep["length_std"])
            if "success_mean" in ep:
                head += " success={:.2f}".format(ep["success_mean"])
        print(head + " pi_loss={:3f} v_loss={:3f} entropy={:.3f} loss={:3f} value={:.3f} advantage={:.3f}".format(
            s[0], s[1], s[3], s[2], torch.mean(self.buffer.values).item(), torch.mean(self.buffer.advantages).item()))

    # ------------------------------------------------------------------------------------------ rollout
    def _stage_host_obs(self, lo=0, hi=None):
        """Host observations of workers [lo, hi) -> a pinned buffer the GPU can DMA from.  The shared slab is itself
        registered as pinned memory (nothing to do); otherwise (pipe transport) stage through ``_obs_pinned``."""
        hi = self.num_workers if hi is None else hi
        if self._obs_slab is not None and self._slab_pinned:
            return self._obs_slab
        self._obs_pinned[lo:hi].copy_(torch.from_numpy(self.obs[lo:hi]))
        return self._obs_pinned

    def _rollout_ctx(self):
        """Persistent per-trainer rollout buffers (their addresses are baked into the captured CUDA
        graphs); the sampling uniforms are redrawn in place every update."""
        if self._ctx is None:
            buf, W, T, L = self.buffer, self.num_workers, self.config["worker_steps"], self.memory_length
            nb = len(self.action_space_shape)
            self._ctx = {
                "flat_mask": buf.memory_mask.view(torch.uint8).view(W * T, L), "flat_idx": buf.memory_indices.view(W * T, L),
                "flat_ep": buf.memory_index.view(W * T), "inner": self.num_blocks * self.embed_dim,
                "uniforms": torch.empty((T, W, nb), device=self.device),
                "step_sched": torch.zeros((T + 1, W), dtype=torch.long, device=self.device),
                "ep_sched": torch.zeros((T + 1, W), dtype=torch.long, device=self.device),
            }
        self._ctx["uniforms"].uniform_()
        return self._ctx

    def _prepare_group(self, grp):
        """Per-rollout device preparation of a worker group: activation workspace / output buffers (once), and -- the
        weights are frozen during a rollout -- conversion of the conv weights to tensor-core format once, outside the
        per-step graphs.  Every group has its own encoder workspace slot: groups run concurrently on their own streams."""
        if grp.outs is None:
            grp.outs = self.model._alloc_outputs(grp.n, self.device)
            grp.ws = None if (self.model._fused_ok and grp.n <= self.model.FUSED_MAX_BATCH) else \
                torch.empty(self.model._ws_floats(grp.n), dtype=torch.float32, device=self.device)
        grp.enc_packed = False
        if self.model._visual and self.model._tc_encoder:
            self.model.pack_encoder_weights(grp.n, *self.obs_shape[1:], slot=grp.index)
            grp.enc_packed = True

    # -- CUDA graphs: the copies and ~10 launches of one rollout step of a worker group are captured once per step index and
    #    replayed, which removes the host launch overhead that otherwise dominates a W=32 forward.
    def _graph_key(self, mode):
        forced = self._forced_buf.data_ptr() if self._forced_on else 0
        roll = 0 if self._table_roll is None else self._table_roll.data_ptr()
        return (mode, self._table.data_ptr(), roll, self.model.flat_parameters().data_ptr(), forced)

    def _step_via_graph(self, mode, grp, t, src):
        """Enqueue step t of worker group ``grp`` on the current stream, as a CUDA-graph replay when possible."""
        if not self.use_cuda_graphs:
            return self._device_step(grp, t, src)
        key = self._graph_key(mode)
        state = grp.graphs
        if state.get("key") != key:                               # table or parameter arena moved: start over
            for old in state.get("steps", {}).values():
                native.graph_destroy(old)
            state.clear()
            state.update(key=key, warm=0, steps={})
        if state["warm"] < self._graph_warm_rollouts:             # first rollout for this key runs eagerly (warm-up)
            return self._device_step(grp, t, src)
        g = state["steps"].get(t)
        if g is None:
            # Capture through the library's own graph API: the step consists only of libtrxlppo calls (no torch
            # op, no allocation, no RNG inside), so nothing of torch's capture machinery is involved.
            cur = torch.cuda.current_stream()
            if self._capture_stream is None:
                self._capture_stream = torch.cuda.Stream(device=self.device)
            cs = self._capture_stream
            cs.wait_stream(cur)
            try:
                with torch.cuda.stream(cs):
                    native.graph_begin(cs.cuda_stream)
                    try:
                        self._device_step(grp, t, src)
                        g = native.graph_end(cs.cuda_stream)
                    except Exception:
                        native.graph_abort(cs.cuda_stream)
                        raise
                cur.wait_stream(cs)
                state["steps"][t] = g
            except Exception as e:  # noqa: BLE001 -- capture is an optimisation; the eager path is the same kernels
                torch.cuda.synchronize()
                print("[trxl] CUDA-graph capture failed (%s); continuing with eager launches" % e)
                self.use_cuda_graphs = False
                return self._device_step(grp, t, src)
        native.graph_launch(g)

    def _graphs_finish_rollout(self, mode, groups):
        if not self.use_cuda_graphs:
            return
        for grp in groups:
            if grp.graphs.get("key") == self._graph_key(mode):
                grp.graphs["warm"] += 1

    @property
    def _graphs(self):
        """Captured step graphs of the first worker group (introspection / tests)."""
        return self._groups[0].graphs if self._groups[0].graphs else self._whole.graphs

    def _device_step(self, grp, t, src):
        """Everything the GPU does for rollout step t of worker group ``grp`` (trainer.py:161-186): fetch the observations
        and episode cursors, store obs, write the mask / window-index rows, run the model with the window read in place,
        write the new memory row, sample actions, store actions / log-probs / values, and hand the actions to the host.
        ``src`` = (obs_ptr, step_ptr, ep_ptr, on_host): where this step's observations (n, *obs) and cursors (n,) live."""
        buf, model, ctx = self.buffer, self.model, self._ctx
        T, L, nb = self.config["worker_steps"], self.memory_length, len(self.action_space_shape)
        n, lo = grp.n, grp.lo
        obs_bytes = int(np.prod(self.obs_shape)) * 4
        obs_ptr, step_ptr, ep_ptr, on_host = src
        row0 = lo * T + t                             # flat buffer row of the group's first worker at step t
        if on_host:
            # pinned host memory is read IN PLACE by one kernel (zero-copy over PCIe): a captured step then consists of kernels
            # only.  (Measured: three copy-engine nodes per step cost 160 us + ~250 us of engine-switch gaps at W = 16.)
            obs_dev, step_dev, ep_dev = grp.obs_dev, grp.step_dev, grp.ep_dev
            native.rollout_fetch(obs_ptr, obs_bytes // 4, step_ptr, ep_ptr, obs_dev, buf.obs.data_ptr() + row0 * obs_bytes,
                                 T * obs_bytes // 4, step_dev, ep_dev, n)
        else:                                         # device-resident feed: use the tensors in place
            obs_dev, step_dev, ep_dev = obs_ptr, step_ptr, ep_ptr
            native.copy_rows(obs_dev.data_ptr(), buf.obs.data_ptr() + row0 * obs_bytes, n, obs_bytes, obs_bytes, T * obs_bytes)
        native.rollout_prepare(step_dev, ep_dev, self._mask_table_dev, self._index_table_dev,
                               ctx["flat_mask"].data_ptr() + row0 * L, T * L, ctx["flat_idx"].data_ptr() + row0 * L * 8, T * L,
                               ctx["flat_ep"].data_ptr() + row0 * 8, T, n, L)
        feat = model.encode(obs_dev, weights_packed=grp.enc_packed, slot=grp.index)
        roll = self._table_roll       # with it, window rows come from the table that already carries the positional rows
        logits, value, new_mem = model.forward_table(feat, self._table if roll is None else roll, ctx["flat_ep"], ctx["flat_idx"],
                                                     ctx["flat_mask"], ctx["flat_idx"] if roll is None else None,
                                                     sample_index=grp.rows[t], n=n, ws=grp.ws, out=grp.outs,
                                                     fused=None if roll is None else True)
        forced = None if self._forced_actions is None else self._forced_actions[t, lo:grp.hi]
        native.sample_actions(logits, ctx["uniforms"][t, lo:grp.hi], self.action_space_shape,
                              buf.actions.data_ptr() + row0 * nb * 8, T * nb, buf.log_probs.data_ptr() + row0 * nb * 4, T * nb,
                              grp.act_host_dptr if on_host else grp.act_dev, n, forced=forced,      # actions land in host memory
                              notify=(grp.done_counter, None) if (on_host and grp.use_flag) else None)    # seq-tagged action words
        # (after the sampling kernel: the host waits for the actions only, so these stores overlap the env phase)
        # table[ep, step] = new memory (trainer.py:174), the same into `roll`, buffer.values[:, t] = value (trainer.py:186)
        native.rollout_store(self._table, roll, None if roll is None else model._pe_table(), ep_dev, step_dev, new_mem,
                             self.max_episode_length, self.num_blocks, self.embed_dim, value=value,
                             value_dst=buf.values.data_ptr() + row0 * 4, value_stride=T)

    def _host_src(self, grp, host_obs):
        """Device-side addresses of this group's slices of the pinned host buffers (observations, cursors)."""
        key = host_obs.data_ptr()
        if self._mapped.get("key") != key:
            self._mapped = {"key": key, "obs": native.host_device_pointer(host_obs.data_ptr()),
                            "step": native.host_device_pointer(self._step_host.data_ptr()),
                            "ep": native.host_device_pointer(self._ep_host.data_ptr())}
        obs_bytes = int(np.prod(self.obs_shape)) * 4
        m = self._mapped
        return (m["obs"] + grp.lo * obs_bytes, m["step"] + grp.lo * 8, m["ep"] + grp.lo * 8, True)

    def _sample_training_data(self):
        """Run every worker for ``worker_steps`` steps (trainer.py:145-225)."""
        if self.device_feed is not None:
            return self._sample_from_device_feed(self.device_feed)
        t0 = time.perf_counter()
        cfg, buf = self.config, self.buffer
        self._begin_rollout()
        self._rollout_ctx()
        self._env_time = 0.0
        with torch.no_grad():
            if self._control is not None:
                episode_infos = self._rollout_grouped()
            else:
                episode_infos = self._rollout_pipes()
        last_value = self.get_last_value()
        buf.calc_advantages(last_value, cfg["gamma"], cfg["lamda"])
        buf.memories = self._table[:self._n_episodes]
        self.timers["env"] += self._env_time
        self.timers["rollout"] += time.perf_counter() - t0
        return episode_infos

    def _rollout_pipes(self):
        """The reference's transport (trainer.py:159-218): one forward over all workers, then 2 pipe messages per worker,
        strictly serial.  Used for workers handed in by the caller (anything with a ``child`` pipe end) and with
        TRXL_PIPE_STEPPING=1."""
        buf, T, grp = self.buffer, self.config["worker_steps"], self._whole
        episode_infos = []
        self._prepare_group(grp)
        stream = torch.cuda.current_stream()
        for t in range(T):
            self._check_cursors()
            src = self._host_src(grp, self._stage_host_obs())
            self._step_via_graph("pipes", grp, t, src)
            stream.synchronize()
            actions = grp.act_np & 0xffffffff              # (the words carry a sequence tag in their upper half)
            te = time.perf_counter()
            for w, worker in enumerate(self.workers):
                worker.child.send(("step", actions[w].copy()))
            for w, worker in enumerate(self.workers):
                obs, buf.rewards[w, t], buf.dones[w, t], info = worker.child.recv()
                if info:                                   # episode finished (trainer.py:195)
                    self._step_host[w] = 0
                    episode_infos.append(info)
                    worker.child.send(("reset", None))
                    obs = worker.child.recv()
                    # the finished episode keeps its table row; the worker continues on a fresh zero row
                    self._ep_host[w] = self._new_row()
                    if t < T - 1:
                        self._n_episodes = self._n_rows
                else:
                    self._step_host[w] += 1
                if obs is not None:                        # None: the worker already wrote it into the shared slab
                    self.obs[w] = obs
            self._env_time += time.perf_counter() - te
        self._graphs_finish_rollout("pipes", [grp])
        return episode_infos

    def _rollout_grouped(self):
        """Shared-memory transport with overlapped worker groups.  Each group cycles GPU -> ENV independently:
          GPU : one graph launch on the group's stream = observation / cursor upload, forward, sampling, action download;
          ENV : actions published to the group's workers (shared arrays + command counters), workers step their
                environments and write the next observation into the pinned slab, acknowledge.
        The host thread only polls (CUDA event of each group, acknowledgement counters) and does the reference's episode
        bookkeeping (trainer.py:193-218) for the group that just finished stepping, so group A's environments step while
        group B's forward runs.  Per-worker semantics are unchanged: every worker still sees forward(t) -> step(t) ->
        forward(t+1) in order, and workers never interact inside a rollout."""
        c, buf, T = self._control, self.buffer, self.config["worker_steps"]
        groups = self._groups
        cur = torch.cuda.current_stream()
        host_obs = self._stage_host_obs()
        if host_obs is not self._obs_slab:
            raise RuntimeError("shared-memory stepping needs the observation slab registered as pinned memory")
        acts, cmd, ack = c["actions"].numpy(), c["cmd"].numpy(), c["ack"].numpy()
        rewards, dones, has_info = c["rewards"].numpy(), c["dones"].numpy(), c["has_info"].numpy()
        sems = c.get("sems")
        futex_words = c.get("futex")
        if futex_words is not None:
            if self._futex is None:
                self._futex = _Futex()
            fwords, fbase = futex_words.numpy(), futex_words.data_ptr()
        episode_infos = []
        trace = [0.0, 0.0, 0.0] if os.environ.get("TRXL_E2E_TRACE") == "1" else None     # enqueue, gpu phase, env phase
        dev_events = []
        for grp in groups:
            self._prepare_group(grp)
            grp.stream.wait_stream(cur)
            grp.t, grp.phase = 0, grp.DONE
            if grp.use_flag:                     # resynchronise the expected sequence number with the device counter
                grp.expected = int(grp.done_counter.item())

        # fast path: when every step graph of this key is captured, a step is ONE ctypes call (cudaGraphLaunch on the group's
        # stream handle); completion is read from the group's pinned flag
        key = self._graph_key("groups")
        fast = {}
        if self.use_cuda_graphs and trace is None:
            for grp in groups:
                st = grp.graphs
                if st.get("key") == key and st.get("warm", 0) >= self._graph_warm_rollouts and len(st.get("steps", {})) == T and grp.use_flag:
                    fast[grp.index] = st["steps"]
        table_ptr = self._table.data_ptr()
        max_len = self.max_episode_length
        step_np, ep_np = self._step_host.numpy(), self._ep_host.numpy()      # numpy views: cheaper per-step bookkeeping than torch ops

        def launch(grp):
            steps = fast.get(grp.index)
            if steps is not None and self._table.data_ptr() == table_ptr:
                if step_np[grp.lo:grp.hi].max() >= max_len:
                    self._check_cursors()
                native.graph_launch_on(steps[grp.t], grp.stream_handle)
                grp.expected += 1
                grp.phase, grp.t_phase = grp.GPU, 0.0
                return
            ta = time.perf_counter()
            if int(self._step_host[grp.lo:grp.hi].max()) >= self.max_episode_length:
                self._check_cursors()
            with torch.cuda.stream(grp.stream):
                if trace is not None:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    dev_events.append((e0, e1))
                    e0.record(grp.stream)
                self._step_via_graph("groups", grp, grp.t, self._host_src(grp, host_obs))
                if trace is not None:
                    e1.record(grp.stream)
                grp.event.record(grp.stream)
            if grp.use_flag:
                grp.expected += 1
            grp.phase, grp.t_phase = grp.GPU, time.perf_counter()
            if trace is not None:
                trace[0] += grp.t_phase - ta

        for grp in groups:
            launch(grp)
        active, idle_polls = len(groups), 0
        deadline = time.perf_counter() + 120.0
        while active:
            progressed = False
            for grp in groups:
                lo, hi = grp.lo, grp.hi
                if grp.phase == grp.GPU:
                    if grp.use_flag:         # every action word of the group carries this launch's sequence number
                        if grp.act_np.min() < (grp.expected << 32):
                            continue
                        now = time.perf_counter()
                        acts[lo:hi] = grp.act_np & 0xffffffff
                    elif not grp.event.query():
                        continue
                    else:
                        now = time.perf_counter()
                        acts[lo:hi] = grp.act_np
                    cmd[lo:hi] += 1                       # publish last: the actions above are visible before the command
                    if futex_words is not None:           # one system call wakes the whole group
                        fwords[FUTEX_WORD_STRIDE * grp.index] += 1
                        self._futex.wake(fbase + 4 * FUTEX_WORD_STRIDE * grp.index)
                    elif sems:
                        for w in range(lo, hi):
                            sems[w].release()
                    grp.phase = grp.ENV
                    if trace is not None:
                        trace[1] += now - grp.t_phase
                    grp.t_phase = now
                    progressed = True
                elif grp.phase == grp.ENV:
                    if not np.array_equal(ack[lo:hi], cmd[lo:hi]):
                        continue
                    now = time.perf_counter()
                    self._env_time += (now - grp.t_phase) / len(groups)
                    if trace is not None:
                        trace[2] += now - grp.t_phase
                    t = grp.t
                    buf.rewards[lo:hi, t] = rewards[lo:hi]
                    buf.dones[lo:hi, t] = dones[lo:hi] != 0
                    step_np[lo:hi] += 1
                    for w in np.nonzero(has_info[lo:hi])[0]:
                        w = lo + int(w)
                        episode_infos.append(self.workers[w].child.recv())
                        step_np[w] = 0
                        ep_np[w] = self._new_row()              # may move the table: every group's graphs are re-keyed
                        if t < T - 1:
                            self._n_episodes = self._n_rows
                    grp.t += 1
                    if grp.t < T:
                        launch(grp)
                    else:
                        grp.phase = grp.DONE
                        active -= 1
                    progressed = True
            if progressed:
                idle_polls = 0
                deadline = time.perf_counter() + 120.0
                continue
            idle_polls += 1
            if self._yield_when_idle:
                os.sched_yield()          # more processes than CPUs: let a runnable env worker have this core while we wait
            if idle_polls % 8192 == 0:
                self._check_workers_alive()
                if time.perf_counter() > deadline:
                    raise RuntimeError("environment workers did not answer within 120 s")
        for grp in groups:
            cur.wait_stream(grp.stream)
        self._graphs_finish_rollout("groups", groups)
        if trace is not None:
            k = 1e6 / (T * len(groups))
            torch.cuda.synchronize()
            dev_us = sorted(1e3 * a.elapsed_time(b) for a, b in dev_events)
            print("[trxl] rollout trace per group step: enqueue %.0f us, gpu phase %.0f us (device time median %.0f us, p90 %.0f us), "
                  "env phase %.0f us (%d groups, %s workers)"
                  % (trace[0] * k, trace[1] * k, dev_us[len(dev_us) // 2], dev_us[int(len(dev_us) * 0.9)], trace[2] * k,
                     len(groups), "futex-blocking" if futex_words is not None else ("blocking" if sems else "spinning")), flush=True)
        return episode_infos

    def _check_workers_alive(self):
        """A worker whose env raised dies with a WorkerException; surface its traceback instead of waiting for the
        stepping timeout (the pipe path raises on recv; the shared-memory path has to look)."""
        for w, worker in enumerate(self.workers):
            proc = getattr(worker, "process", None)
            if proc is not None and not proc.is_alive():
                detail = ""
                try:
                    if worker.child.poll(0):
                        detail = ": %r" % (worker.child.recv(),)
                except (EOFError, OSError):
                    pass
                raise RuntimeError("environment worker %d exited (exit code %s)%s" % (w, proc.exitcode, detail))

    def _check_cursors(self):
        """An environment that runs past its advertised max_episode_steps would index past its episode row (the reference
        raises IndexError at trainer.py:166); fail loudly instead of corrupting the next table row."""
        if int(self._step_host.max()) >= self.max_episode_length:
            w = int(self._step_host.argmax())
            raise IndexError("worker %d reached episode step %d but the environment advertises max_episode_steps = %d"
                             % (w, int(self._step_host[w]), self.max_episode_length))

    def _sample_from_device_feed(self, feed):
        """Rollout against a device-resident synthetic feed (device_feed.py): the episode schedule of the
        whole update is known up front, so the per-step cursors (episode step, table row) are uploaded
        once and the T steps run back to back without host synchronisation."""
        t0 = time.perf_counter()
        cfg, buf = self.config, self.buffer
        W, T = self.num_workers, cfg["worker_steps"]
        self._begin_rollout()
        feed.begin_update()
        # replay the reference's bookkeeping (trainer.py:195-216) on the host for all T steps at once
        step_sched = np.zeros((T + 1, W), dtype=np.int64)
        ep_sched = np.zeros((T + 1, W), dtype=np.int64)
        step_sched[0], ep_sched[0] = self._step_host.numpy(), self._ep_host.numpy()
        episode_infos = []
        n_rows = self._n_rows
        for t in range(T):
            step_sched[t + 1] = step_sched[t] + 1
            ep_sched[t + 1] = ep_sched[t]
            for w, info in feed.infos[t]:
                step_sched[t + 1, w] = 0
                ep_sched[t + 1, w] = n_rows
                n_rows += 1
                if t < T - 1:
                    self._n_episodes = n_rows
                episode_infos.append(info)
        if int(step_sched[:T].max()) >= self.max_episode_length:
            raise IndexError("the feed's schedule runs past max_episode_steps = %d" % self.max_episode_length)
        while n_rows > self._table_cap:
            self._alloc_table(self._table_cap * 2)
        self._n_rows = n_rows
        buf.rewards[:] = feed.rewards.T
        buf.dones[:] = feed.dones.T
        ctx = self._rollout_ctx()
        grp = self._whole
        self._prepare_group(grp)
        step_dev, ep_dev = ctx["step_sched"], ctx["ep_sched"]
        step_dev.copy_(torch.from_numpy(step_sched))
        ep_dev.copy_(torch.from_numpy(ep_sched))
        with torch.no_grad():
            for t in range(T):
                self._step_via_graph("feed", grp, t, (feed.obs(t), step_dev[t], ep_dev[t], False))
        self._graphs_finish_rollout("feed", [grp])
        self._step_host.copy_(torch.from_numpy(step_sched[T]))
        self._ep_host.copy_(torch.from_numpy(ep_sched[T]))
        self._feed_last_obs = feed.obs(T)
        last_value = self.get_last_value()
        buf.calc_advantages(last_value, cfg["gamma"], cfg["lamda"])
        buf.memories = self._table[:self._n_episodes]
        self.timers["rollout"] += time.perf_counter() - t0
        return episode_infos

    def get_last_value(self):
        """Bootstrap value of the current observation (trainer.py:227-237).  Quirks kept: the window is
        ``[clip(s-L,0), clip(s,L))`` and the positional indices are those of the last rollout step."""
        W, L, T = self.num_workers, self.memory_length, self.config["worker_steps"]
        step = self._step_host.clone()
        start = torch.clip(step - L, 0)
        self._win_last.copy_((start.unsqueeze(1) + torch.arange(L).unsqueeze(0)).to(self.device))
        self._mask_last.copy_(self._mask_table_dev[torch.clip(step, 0, L - 1).to(self.device)])
        self._step_dev.copy_(self._step_host, non_blocking=True)
        self._ep_dev.copy_(self._ep_host, non_blocking=True)
        with torch.no_grad():
            if self.device_feed is not None:
                self._obs_dev.copy_(self._feed_last_obs)
            else:
                self._obs_dev.copy_(self._stage_host_obs(), non_blocking=True)
            feat = self.model.encode(self._obs_dev)
            pe_idx = self.buffer.memory_indices[:, -1].contiguous()
            _, value, _ = self.model.forward_table(feat, self._table, self._ep_dev, self._win_last, self._mask_last, pe_idx, n=W)
        return value.clone()

    # ------------------------------------------------------------------------------------------ optimisation
    def _train_epochs(self, learning_rate, clip_range, beta):
        """epochs x minibatches (trainer.py:239-256).  Returns (list of 6-stat lists, dict of grad-norm lists).
        Statistics stay on the device until the end of the update: one host sync per update."""
        t0 = time.perf_counter()
        n_steps = self.config["epochs"] * (self.buffer.batch_size // self.buffer.mini_batch_size +
                                           (1 if self.buffer.batch_size % self.buffer.mini_batch_size else 0))
        g = self.model._n_groups
        stats = torch.zeros((n_steps, 6), dtype=torch.float32, device=self.device)
        norms = torch.zeros((n_steps, g + 2), dtype=torch.float32, device=self.device)
        i = 0
        grouping = self._begin_grouped_attention()
        for epoch in range(self.config["epochs"]):
            batches = list(self.buffer.mini_batch_generator())
            if grouping is not None:
                self._group_epoch(batches, grouping, epoch)
            # advantage statistics of every minibatch of the epoch (the permutation is known up front): one small
            # all-reduce per epoch instead of one per optimiser step
            advstats = torch.zeros((len(batches), 3), dtype=torch.float64, device=self.device)
            flat_adv = self.buffer.samples_flat["advantages"]
            for j, mini_batch in enumerate(batches):
                native.adv_stats(flat_adv, mini_batch.sample_index, mini_batch.sample_index.shape[0], advstats[j])
            self.dp.all_reduce_(advstats)
            for j, mini_batch in enumerate(batches):
                self._ppo_step(mini_batch, learning_rate, clip_range, beta, stats[i], norms[i], advstats=advstats[j])
                i += 1
        stats, norms = stats[:i].cpu(), norms[:i].cpu()           # the only sync of the update
        train_info = [[np.float32(v) for v in row] for row in stats.tolist()]
        grad_info = {}
        for row in norms:
            for key, value in self.model.grad_norms_from(row).items():
                grad_info.setdefault(key, []).append(value)
        self.timers["train"] += time.perf_counter() - t0
        return train_info, grad_info

    # -- episode-grouped tensor-core attention (csrc/attention_tc.cu): the host sorts every minibatch by episode and cuts it into
    #    row tiles whose samples share an episode, so that the energies / context contractions become dense GEMMs per tile
    def _begin_grouped_attention(self):
        """Once per update: the table with positional rows added (the table is frozen during the epochs) and the episode of
        every buffer row on the host.  Returns None when the configuration uses the per-sample kernel."""
        model, buf = self.model, self.buffer
        forced = os.environ.get("TRXL_GROUPED_ATTENTION")
        if forced == "0" or not native.grouped_attention_supported(model._cfg):
            return None
        # tiny minibatches (c1: 64 samples x 1 head) are one short launch of the per-sample kernel; three launches plus the host
        # grouping only pay off once there are enough (sample, head) rows to fill 128-row tiles
        if forced != "1" and buf.mini_batch_size * model.transformer.num_heads < 1024:
            return None
        # the grouped GEMMs cover all M slots of an episode, the per-sample kernel only the L window slots: with short windows
        # in long episodes (c2: L = 32, M = 200; measured 160 k vs 173 k env-steps/s) the streaming kernel does less work
        if forced != "1" and self.max_episode_length > 4 * self.memory_length:
            return None
        table = buf.memories
        if not torch.is_tensor(table) or table.dim() != 4 or table.shape[1] != self.max_episode_length:
            return None
        pe = model._pe_table()
        pre_ln = model.transformer.config["layer_norm"] == "pre"
        if pe is None and not pre_ln:
            table_pe = table                                           # neither positional rows nor norm_kv: read the table itself
        else:
            if self._table_pe is None or self._table_pe.shape != self._table.shape:
                self._table_pe = torch.empty_like(self._table)
            table_pe = self._table_pe[:table.shape[0]]
            native.table_add_pe(table, pe, table_pe, table.shape[0], layer_norm=pre_ln)
        episode_of_row = buf.samples_flat["memory_index"].cpu().numpy()       # (W*T,) -- one small download per update
        return {"table_pe": table_pe, "n_episodes": int(table.shape[0]), "episode_of_row": episode_of_row,
                "rows_per_tile": 128 // self.model.transformer.num_heads * self.model.transformer.num_heads}

    def _group_epoch(self, batches, grouping, epoch=0):
        """Sort every minibatch of the epoch by episode (a minibatch is a set: the loss and its gradient do not depend on the
        order) and build its tile table {first (sample, head) row, rows, episode, 0}; one upload for the whole epoch."""
        H = self.model.transformer.num_heads
        sorted_idx, tiles, n_tiles = [], [], []
        for mb in batches:
            idx, t = group_minibatch_by_episode(mb.sample_index_cpu.numpy(), grouping["episode_of_row"], H)
            sorted_idx.append(idx)
            tiles.append(t)
            n_tiles.append(len(t))
        idx_dev = torch.from_numpy(np.concatenate(sorted_idx)).to(self.device)
        # every minibatch's tile table padded with empty entries (rows = 0: the CTA exits) to one length that is stable across
        # minibatches and updates, so that a captured optimiser step can be replayed
        max_tiles = tile_table_length(max(len(b.sample_index_cpu) for b in batches), H, grouping["n_episodes"])
        padded = np.zeros((len(batches), max_tiles, 4), dtype=np.int32)
        for j, t in enumerate(tiles):
            padded[j, :len(t)] = t
        tiles_dev = torch.from_numpy(padded).to(self.device)
        i0 = 0
        for j, (mb, idx, nt) in enumerate(zip(batches, sorted_idx, n_tiles)):
            mb.sample_index = idx_dev[i0:i0 + len(idx)]
            mb.sample_index_cpu = torch.from_numpy(idx)
            mb.groups = {"tiles": tiles_dev[j, :nt], "n_tiles": nt, "table_pe": grouping["table_pe"],
                         "n_episodes": grouping["n_episodes"], "tiles_padded": tiles_dev[j], "max_tiles": max_tiles, "slot": j + epoch * len(batches)}
            i0 += len(idx)

    def _train_mini_batch(self, samples, learning_rate, clip_range, beta):
        """One optimiser step on one minibatch (trainer.py:258-323).  ``samples`` is either a ``MiniBatch``
        from this package's buffer or a plain dict with the reference's keys (materialised tensors).
        Returns [policy_loss, vf_loss, loss, entropy, approx_kl, clip_fraction] as numpy scalars."""
        g = self.model._n_groups
        stats = torch.zeros(6, dtype=torch.float32, device=self.device)
        self._last_norms = torch.zeros(g + 2, dtype=torch.float32, device=self.device)
        self._ppo_step(samples, learning_rate, clip_range, beta, stats, self._last_norms)
        return [np.asarray(v, dtype=np.float32) for v in stats.cpu().tolist()]

    def _scratch(self, n):
        st = self._train_state.get(n)
        if st is None:
            dev, model = self.device, self.model
            st = {
                "ws": torch.empty(model._ws_floats(n), dtype=torch.float32, device=dev),
                "out": model._alloc_outputs(n, dev),
                "dlogits": torch.empty((n, model._sum_actions), dtype=torch.float32, device=dev),
                "dvalue": torch.empty((n,), dtype=torch.float32, device=dev),
                "advstats": torch.zeros(3, dtype=torch.float64, device=dev),
                "loss_scratch": torch.empty(n // 128 * 5 + 64, dtype=torch.float32, device=dev),
                "dfeat": torch.empty((n, model._feat_dim), dtype=torch.float32, device=dev) if model._visual else None,
                "obs": None if (model._visual and model._tc_encoder) else torch.empty((n,) + self.obs_shape, dtype=torch.float32, device=dev),
            }
            self._train_state = {n: st}          # keep one size resident
        return st

    def _train_segment(self, gstate, name, fn):
        """Run one launch-dense segment of the optimiser step: eagerly (``gstate`` None, or the first step with a new key),
        or as a CUDA-graph replay captured through the library's own graph API (the segment consists of libtrxlppo calls
        only: no torch op, no allocation).  Returns what ``fn`` returned when it last ran on the host."""
        if gstate is None:
            return fn()
        if gstate["warm"] < 2:                     # both segments of one step run eagerly before anything is captured
            gstate["warm"] += 1
            out = fn()
            gstate.setdefault("ret", {})[name] = out
            return out
        g = gstate["graphs"].get(name)
        if g is None:
            cur = torch.cuda.current_stream()
            if self._capture_stream is None:
                self._capture_stream = torch.cuda.Stream(device=self.device)
            cs = self._capture_stream
            cs.wait_stream(cur)
            try:
                with torch.cuda.stream(cs):
                    native.graph_begin(cs.cuda_stream)
                    try:
                        gstate.setdefault("ret", {})[name] = fn()
                        g = native.graph_end(cs.cuda_stream)
                    except Exception:
                        native.graph_abort(cs.cuda_stream)
                        raise
                cur.wait_stream(cs)
                gstate["graphs"][name] = g
            except Exception as e:  # noqa: BLE001 -- capture is an optimisation; the eager path is the same kernels
                torch.cuda.synchronize()
                print("[trxl] CUDA-graph capture of the training step failed (%s); continuing with eager launches" % e)
                self.use_train_graphs = False
                return fn()
        native.graph_launch(g)
        return gstate["ret"][name]

    def _ppo_step(self, samples, learning_rate, clip_range, beta, stats_out, norms_out, advstats=None):
        """One optimiser step.  ``advstats`` = the (already all-reduced) {sum, sum of squares, count} of the minibatch's
        advantages; computed (and all-reduced) here when the caller did not batch them per epoch."""
        model, cfg = self.model, self.config
        if isinstance(samples, MiniBatch):
            buf = samples.buffer
            flat = buf.samples_flat
            sidx = samples.sample_index
            n = sidx.shape[0]
            st = self._scratch(n)
            if model._visual and model._tc_encoder:
                obs = flat["obs"]                    # the encoder's first gather reads the minibatch rows in place
            else:
                native.gather_rows(flat["obs"], sidx, st["obs"])
                obs = st["obs"]
            table = buf.memories
            ep_index, win_index = flat["memory_index"], flat["memory_indices"]
            mask = flat["memory_mask"].view(torch.uint8)
            actions, old_logp, old_values, adv = flat["actions"], flat["log_probs"], flat["values"], flat["advantages"]
        else:   # reference-style dict of materialised tensors (trainer.py:271-274)
            dev = self.device
            obs = samples["obs"].to(dev, torch.float32).contiguous()
            n = obs.shape[0]
            st = self._scratch(n)
            sidx = None
            table = samples["memories"].to(dev, torch.float32).contiguous()          # (n, M, B, D), episode n <-> sample n
            ep_index = None
            win_index = samples["memory_indices"].to(dev, torch.int64).contiguous()
            mask = (samples["memory_mask"].to(dev) != 0).to(torch.uint8).contiguous()
            actions = samples["actions"].to(dev, torch.int64).contiguous()
            old_logp = samples["log_probs"].to(dev, torch.float32).contiguous()
            old_values = samples["values"].to(dev, torch.float32).contiguous()
            adv = samples["advantages"].to(dev, torch.float32).contiguous()
        pe_index = win_index

        for group in self.optimizer.param_groups:
            group["lr"] = learning_rate
        self.optimizer.zero_grad()
        # encoder (cuDNN) with autograd so its backward can be driven by d loss / d features
        tc_enc = model._visual and model._tc_encoder
        if tc_enc:
            feat_g, feat = None, None              # computed inside the forward segment below
        elif model._visual:
            with torch.enable_grad():
                feat_g = model.encode(obs)
            feat = feat_g.detach().contiguous()
        else:
            feat_g, feat = None, obs.reshape(n, -1)
        logits, value, out_mem = st["out"]
        groups = None
        grouped = isinstance(samples, MiniBatch) and samples.groups is not None
        # CUDA-graph replay of the two launch-dense segments: needs every address in them to be the same from one minibatch
        # to the next, so the row indices and the (padded) tile table are copied to fixed staging buffers first.  While the
        # library's event timers are on (bench.py's attention timing) the first minibatch of every update stays eager: events
        # cannot be recorded inside a captured graph.
        gstate = None
        if getattr(self, "use_train_graphs", False) and grouped and tc_enc and \
                not (native.profiling() and samples.groups.get("slot") == 0):
            gr = samples.groups
            if st.get("sidx_stage") is None or st["tiles_stage"].shape[0] != gr["max_tiles"]:
                st["sidx_stage"] = torch.empty(n, dtype=torch.long, device=self.device)
                st["tiles_stage"] = torch.empty((gr["max_tiles"], 4), dtype=torch.int32, device=self.device)
            key = (n, gr["max_tiles"], gr["table_pe"].data_ptr(), table.data_ptr(), self._table_cap, model.flat_parameters().data_ptr(),
                   obs.data_ptr(), st["ws"].data_ptr(), st["sidx_stage"].data_ptr(), st["tiles_stage"].data_ptr())
            gstate = self._train_graphs.get(n)               # one state per minibatch size (a ragged last minibatch keeps its own)
            if gstate is None or gstate["key"] != key:
                for old in (gstate or {}).get("graphs", {}).values():
                    native.graph_destroy(old)
                gstate = self._train_graphs[n] = {"key": key, "warm": 0, "graphs": {}}
            native.copy_async(sidx.data_ptr(), st["sidx_stage"].data_ptr(), n * 8)
            native.copy_async(gr["tiles_padded"].data_ptr(), st["tiles_stage"].data_ptr(), gr["max_tiles"] * 16)
            sidx = st["sidx_stage"]
        if grouped:
            if st.get("ranges") is None:
                st["ranges"] = torch.empty((n, 4), dtype=torch.int32, device=self.device)
            gr = samples.groups
            if gstate is not None:        # staged, padded table; the table's capacity as the (stable) episode bound
                groups = native.attn_groups(gr["table_pe"], self._table_cap, st["tiles_stage"], gr["max_tiles"], st["ranges"])
            else:
                groups = native.attn_groups(gr["table_pe"], gr["n_episodes"], gr["tiles"], gr["n_tiles"], st["ranges"])

        def forward_segment():
            if grouped:
                native.attention_ranges(mask, win_index, ep_index, sidx, n, self.memory_length, st["ranges"])
            f = model.encode_train(obs, sidx, n) if tc_enc else feat
            native.model_forward(model._cfg, model.flat_parameters(), f, table, table.shape[1], ep_index, win_index, mask,
                                 pe_index, sidx, model._pe_table(), n, st["ws"], logits, value, out_mem, groups=groups)
            return f

        def backward_segment():
            native.model_backward(model._cfg, model.flat_parameters(), model.flat_grads(), feat, table, table.shape[1], ep_index,
                                  win_index, mask, pe_index, sidx, model._pe_table(), n, st["ws"], out_mem, st["dlogits"],
                                  st["dvalue"], st["dfeat"], groups=groups)
            if tc_enc:
                model.encode_backward(n, obs.shape[-2], obs.shape[-1], st["dfeat"])

        feat = self._train_segment(gstate, "fwd", forward_segment)
        if advstats is None:
            advstats = st["advstats"]
            native.adv_stats(adv, sidx, n, advstats)
            self.dp.all_reduce_(advstats)
        multi = self.dp.world_size > 1
        # multi-GPU: the six loss statistics (already divided by the GLOBAL sample count) are written into the tail of the
        # gradient arena and summed by the same all-reduce as the gradients
        stats_dst = model.stats_tail() if multi else stats_out
        native.ppo_loss(logits, value, actions, old_logp, old_values, adv, sidx, advstats, self.action_space_shape, n,
                        clip_range, beta, cfg["value_loss_coefficient"], st["dlogits"], st["dvalue"], stats_dst,
                        st["loss_scratch"])
        self._train_segment(gstate, "bwd", backward_segment)
        if not tc_enc and feat_g is not None:
            feat_g.backward(st["dfeat"])           # accumulates into the conv slices of the gradient arena
        if multi:
            self.dp.all_reduce_(model.flat_grads_with_tail())     # ONE collective per optimiser step: gradient arena + stats tail
            stats_out.copy_(model.stats_tail())
        self.optimizer.step(norms_out=norms_out)

    # ------------------------------------------------------------------------------------------ logging / io
    def _write_training_summary(self, update, training_stats, episode_result):
        if self.writer is None:
            return
        for key, value in (episode_result or {}).items():
            if "std" not in key:
                self.writer.add_scalar("episode/" + key, value, update)
        names = ("losses/policy_loss", "losses/value_loss", "losses/loss", "losses/entropy")
        for name, value in zip(names, training_stats[:4]):
            self.writer.add_scalar(name, value, update)
        self.writer.add_scalar("training/value_mean", torch.mean(self.buffer.values).item(), update)
        self.writer.add_scalar("training/advantage_mean", torch.mean(self.buffer.advantages).item(), update)
        # the reference writes stats[4] (KL) under "clip_fraction" and stats[5] under "kl" (trainer.py:343-344);
        # the tags are kept so dashboards line up with reference runs
        self.writer.add_scalar("other/clip_fraction", training_stats[4], update)
        self.writer.add_scalar("other/kl", training_stats[5], update)

    def _write_gradient_summary(self, update, grad_info):
        if self.writer is None:
            return
        for key, value in grad_info.items():
            self.writer.add_scalar("gradients/" + key, np.mean(value), update)

    def _save_model(self):
        """``(state_dict, config)`` pickle at ./models/<run_id>.nn, the reference's checkpoint format (trainer.py:356-362)."""
        os.makedirs("./models", exist_ok=True)
        save_model_file(self.model, self.config, "./models/" + self.run_id + ".nn")
        print("Model saved to ./models/" + self.run_id + ".nn")

    def save_checkpoint(self, path):
        """Everything needed to resume: parameters, AdamW moments/step, next update index, config.  (The
        reference only saves ``(state_dict, config)`` at the end and cannot resume; that format is ``_save_model``.)"""
        state = {"state_dict": {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()},
                 "optimizer": self.optimizer.state_dict(), "next_update": self._start_update, "config": self.config}
        with open(path, "wb") as f:
            pickle.dump(state, f)

    def load_checkpoint(self, path):
        with open(path, "rb") as f:
            state = pickle.load(f)
        self.model.load_state_dict(state["state_dict"])
        self.optimizer.load_state_dict(state["optimizer"])
        self._start_update = int(state.get("next_update", 0))
        return state

    def close(self, exit_process=True):
        """Shut down workers and the summary writer (trainer.py:364-383; the reference also exits the process)."""
        try:
            if self.writer is not None:
                self.writer.close()
        except Exception:
            pass
        for worker in self.workers:
            try:
                worker.child.send(("close", None))
            except Exception:
                pass
        time.sleep(0.2)
        if getattr(self, "_obs_slab", None) is not None and getattr(self, "_slab_pinned", False):
            try:
                torch.cuda.cudart().cudaHostUnregister(self._obs_slab.data_ptr())
            except Exception:
                pass
            self._slab_pinned = False
        for worker in self.workers:
            proc = getattr(worker, "process", None)
            if proc is not None:
                proc.join(timeout=1.0)
                if proc.is_alive():
                    proc.terminate()
        if exit_process:
            raise SystemExit(0)
