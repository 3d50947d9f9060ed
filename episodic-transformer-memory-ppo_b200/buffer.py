"""Rollout buffer, GAE and minibatch generation -- B200-native drop-in for the reference's
``buffer.Buffer`` (buffer.py:6-113): same constructor, attributes, ``prepare_batch_dict()``,
``mini_batch_generator()`` (same dict keys) and ``calc_advantages(last_value, gamma, lamda)``.

What is different underneath:
  * every tensor lives on the training device from the start (the reference allocates them wherever
    torch's default tensor type points);
  * episodic memories are ONE table tensor (E, M, B, D) instead of a Python list of per-episode
    tensors; ``memory_index`` rows address it;
  * a minibatch is a set of row indices (``sample_index``) into the flat buffer.  The reference copies
    every key per minibatch, including ``memories[memory_index[idx]]`` -- (mb, M, B, D), 2.1 GB at the
    Minigrid shapes (buffer.py:90).  Here the values materialise only when a consumer actually reads a
    key; the native trainer reads the flat arrays in place through ``sample_index`` and never touches
    ``memories``;
  * GAE runs as one kernel, bit-identical to the reference's fp32 loop.
"""
import numpy as np
import torch

import trxl_native as native


class MiniBatch(dict):
    """Dict with the reference's minibatch keys, materialised lazily from the flat buffer.

    ``mb["obs"]`` etc. gather on first access; ``mb.sample_index`` / ``mb.buffer`` give the native
    trainer in-place access.  ``mb["memories"]`` builds the reference's (mb, M, B, D) tensor on demand."""

    KEYS = ("actions", "values", "log_probs", "advantages", "obs", "memory_mask", "memory_indices", "memories")

    def __init__(self, buffer, sample_index, sample_index_cpu=None):
        super().__init__()
        self.buffer = buffer
        self.sample_index = sample_index
        self.sample_index_cpu = sample_index_cpu      # same rows on the host (the trainer groups them by episode)
        self.groups = None                            # episode grouping for the tensor-core attention (trainer._group_epoch)

    def __missing__(self, key):
        flat = self.buffer.samples_flat
        idx = self.sample_index
        if key == "memories":
            val = self.buffer.memories[flat["memory_index"][idx]]
        elif key == "obs":
            src = flat["obs"]
            val = torch.empty((idx.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
            native.gather_rows(src, idx, val)
        elif key in flat:
            val = flat[key][idx].to(self.buffer.device)
        else:
            raise KeyError(key)
        self[key] = val
        return val

    def keys(self):
        return self.KEYS

    def __iter__(self):
        return iter(self.KEYS)

    def __contains__(self, key):
        return key in self.KEYS

    def items(self):
        return [(k, self[k]) for k in self.KEYS]

    def __len__(self):
        return len(self.KEYS)


class Buffer():
    def __init__(self, config, observation_space, action_space_shape, max_episode_length, device):
        self.device = torch.device(device)
        self.n_workers = config["n_workers"]
        self.worker_steps = config["worker_steps"]
        self.n_mini_batches = config["n_mini_batch"]
        self.batch_size = self.n_workers * self.worker_steps
        self.mini_batch_size = self.batch_size // self.n_mini_batches
        self.max_episode_length = max_episode_length
        t = config["transformer"]
        self.memory_length, self.num_blocks, self.embed_dim = t["memory_length"], t["num_blocks"], t["embed_dim"]
        w, s, nb, dev = self.n_workers, self.worker_steps, len(action_space_shape), self.device

        # host side (filled from the env workers, as in the reference: buffer.py:30,32)
        self.rewards = np.zeros((w, s), dtype=np.float32)
        self.dones = np.zeros((w, s), dtype=bool)
        # device side
        self.actions = torch.zeros((w, s, nb), dtype=torch.long, device=dev)
        self.obs = torch.zeros((w, s) + tuple(observation_space.shape), dtype=torch.float32, device=dev)
        self.log_probs = torch.zeros((w, s, nb), dtype=torch.float32, device=dev)
        self.values = torch.zeros((w, s), dtype=torch.float32, device=dev)
        self.advantages = torch.zeros((w, s), dtype=torch.float32, device=dev)
        self.memories = []          # list of (M, B, D) episode tensors (reference style) or one (E, M, B, D) table
        self.memory_mask = torch.zeros((w, s, self.memory_length), dtype=torch.bool, device=dev)
        self.memory_index = torch.zeros((w, s), dtype=torch.long, device=dev)
        self.memory_indices = torch.zeros((w, s, self.memory_length), dtype=torch.long, device=dev)
        self.samples_flat = {}

    def prepare_batch_dict(self):
        """Flatten (W, T, ...) -> (W*T, ...) views (buffer.py:49-70).  A list of episode memories is
        stacked into the table form; a table is kept as is."""
        if isinstance(self.memories, (list, tuple)):
            self.memories = torch.stack([m.to(self.device) for m in self.memories], dim=0)
        samples = {"actions": self.actions, "values": self.values, "log_probs": self.log_probs,
                   "advantages": self.advantages, "obs": self.obs, "memory_mask": self.memory_mask,
                   "memory_index": self.memory_index, "memory_indices": self.memory_indices}
        self.samples_flat = {k: v.reshape((v.shape[0] * v.shape[1],) + tuple(v.shape[2:])) for k, v in samples.items()}

    def mini_batch_generator(self, generator=None):
        """Yield ``n_mini_batch`` shuffled minibatches (plus a short remainder batch if the batch size
        does not divide, as buffer.py:82 does).  The permutation is drawn with torch's CPU generator so
        a seeded run sees the same index stream as the reference on CPU."""
        perm_cpu = torch.randperm(self.batch_size, generator=generator, device="cpu")
        perm = perm_cpu.to(self.device)
        size = self.batch_size // self.n_mini_batches
        for start in range(0, self.batch_size, size):
            yield MiniBatch(self, perm[start:start + size].contiguous(), perm_cpu[start:start + size])

    def calc_advantages(self, last_value, gamma, lamda):
        """Generalised advantage estimation (buffer.py:95-113) as one kernel; results are bit-identical
        to the reference's fp32 CPU loop."""
        dev = self.device
        rewards = torch.from_numpy(self.rewards).to(dev)
        dones = torch.from_numpy(self.dones.astype(np.uint8)).to(dev)
        lv = last_value.detach().to(dev, torch.float32).contiguous()
        native.gae(rewards, dones, self.values, lv, self.advantages, gamma, lamda)
