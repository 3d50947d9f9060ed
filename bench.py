#!/usr/bin/env python
"""Benchmark of the PPO + TrXL hot path on BASELINE.json's metric: env-steps/sec on synthetic
Minigrid-shaped observations (3x84x84), memory_length 128, embed_dim 256, 4 heads, 4 blocks,
n_workers 32 x worker_steps 512 per GPU (configs[2], the configuration the metric is quoted on).

    python bench.py --gpus 1 --steps 3 --warmup 3            # this engine (one JSON line)
    python bench.py --impl reference --steps 2 --warmup 1    # reference algorithm on the host CPU cores
    torchrun ... bench.py --gpus N ...                       # one rank per GPU, weak scaling

One "step" = one PPO update = worker_steps rollout forwards over all workers + GAE + epochs x
minibatches of forward/backward/clip/AdamW.
  value : env-steps/s with the synthetic env resident on the device (device_feed.py), timed with CUDA
          events, max over ranks.
  e2e   : the same metric through PPOTrainer with real env worker processes (worker.py pipes):
          observations arrive in host memory every step and are copied host->device, actions are
          copied device->host, inside the timed region.
  roofline : the fused window-attention forward kernel (training launches), CUDA-event timed inside
          the timed region; algorithmic bytes = 4*L*D per (sample, block).
  cpu_baseline / --impl reference : the oracle port (oracle/*.py, a restatement of the reference pinned
          to reference-generated fixtures; the Python reference itself cannot travel to the GPU box)
          timed on the host cores on a bounded sample and extrapolated to one update.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "episodic-transformer-memory-ppo_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "env-steps/sec (whole box) Minigrid 3x84x84 mem_len=128"       # BASELINE.json's metric (quoted on c3 = configs[2])
UNIT = "env-steps/s"


def metric_name(name):
    return METRIC if name.startswith("c3_") else "env-steps/sec (whole box) " + name


def load_workload(name):
    from yaml_parser import YamlParser
    return YamlParser(os.path.join(PKG, "configs", name + ".yaml")).get_config()


def workload_desc(cfg, name, n_gpus, scaling="weak"):
    t = cfg["transformer"]
    return {"workload": name, "obs_shape": cfg["environment"]["obs_shape"], "n_workers_per_gpu": cfg["n_workers"],
            "n_workers_total": cfg["n_workers"] * n_gpus, "scaling_mode": scaling,
            "worker_steps": cfg["worker_steps"], "memory_length": t["memory_length"], "embed_dim": t["embed_dim"],
            "num_heads": t["num_heads"], "num_blocks": t["num_blocks"], "epochs": cfg["epochs"],
            "n_mini_batch": cfg["n_mini_batch"], "layer_norm": t["layer_norm"], "positional_encoding": t["positional_encoding"],
            "global_env_steps_per_update": cfg["n_workers"] * cfg["worker_steps"] * n_gpus,
            "parallelism": "dp%d (workers sharded, 1 flat-grad all-reduce / optimiser step + 1 advantage-statistics all-reduce / epoch)" % n_gpus,
            "l2_note": "per-step inputs (rollout obs buffer %.2f GB + minibatch obs %.0f MB + episode table) exceed the 126 MB L2"
                       % (cfg["n_workers"] * cfg["worker_steps"] * float(np.prod(cfg["environment"]["obs_shape"])) * 4 / 1e9,
                          cfg["n_workers"] * cfg["worker_steps"] / cfg["n_mini_batch"] * float(np.prod(cfg["environment"]["obs_shape"])) * 4 / 1e6)}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_sample(cfg, rollout_steps=8, threads=None, seed=0, mb_sample=512, reps=2):
    """Time the oracle port on the host cores: `rollout_steps` rollout steps over all workers plus ONE
    full-size minibatch step (episode gather + window gather + forward + backward + clip + AdamW, as the
    reference does them), extrapolated to a whole update.  Returns (env_steps_per_s, detail dict)."""
    from oracle import ppo_oracle as O
    from oracle import trxl_oracle as X
    from environments.synthetic_env import SyntheticEnv
    torch.manual_seed(seed)
    env_cfg = cfg["environment"]
    obs_shape, nact, M = tuple(env_cfg["obs_shape"]), env_cfg["n_actions"], env_cfg["max_episode_steps"]
    W, T, t = cfg["n_workers"], cfg["worker_steps"], cfg["transformer"]
    ocfg = dict(cfg, max_episode_steps=M, action_space_shape=(nact,))
    P = X.init_params(ocfg, obs_shape, seed=seed)
    envs = [SyntheticEnv(obs_shape, nact, M, env_cfg.get("min_episode_steps"), seed=seed + 1 + w) for w in range(W)]
    st = O.new_rollout_state(ocfg, obs_shape, envs)
    if threads is None:
        # torch's default (one thread per host core) can be far from the best choice for these op sizes;
        # give the CPU arm its best case: probe a few pool sizes on two rollout steps and keep the fastest
        cores = os.cpu_count() or 1
        best = (float("inf"), cores)
        for cand in sorted({c for c in (4, 8, 16, 32, 64, cores) if c <= cores}):
            torch.set_num_threads(cand)
            t0 = time.perf_counter()
            O.sample_rollout(P, dict(ocfg, worker_steps=2), st, envs)
            best = min(best, (time.perf_counter() - t0, cand))
        threads = best[1]
    torch.set_num_threads(threads)
    short = dict(ocfg, worker_steps=rollout_steps)
    t0 = time.perf_counter()
    O.sample_rollout(P, short, st, envs)
    t_roll = (time.perf_counter() - t0) / rollout_steps
    # one optimiser step on a bounded slice of a real minibatch (cost is linear in the sample count)
    mb_full = W * T // cfg["n_mini_batch"]
    mb_size = min(mb_full, mb_sample)
    g = torch.Generator().manual_seed(seed)
    n_eps = max(W, W * T // max(1, M // 2))
    table = torch.randn((n_eps, M, t["num_blocks"], t["embed_dim"]), generator=g)
    idx_table = X.window_index_table(M, t["memory_length"])
    mask_table = X.attention_mask_table(t["memory_length"])
    steps = torch.randint(0, M, (mb_size,), generator=g)
    flat = {"obs": torch.rand((mb_size,) + obs_shape, generator=g), "memory_index": torch.randint(0, n_eps, (mb_size,), generator=g),
            "memory_indices": idx_table[steps], "memory_mask": mask_table[torch.clip(steps, 0, t["memory_length"] - 1)].bool(),
            "actions": torch.randint(0, nact, (mb_size, 1), generator=g), "values": torch.randn(mb_size, generator=g),
            "advantages": torch.randn(mb_size, generator=g), "log_probs": -torch.rand((mb_size, 1), generator=g) - 0.5}
    opt = {}
    t_best = float("inf")
    for _ in range(reps):                                                        # best of `reps`: first touch of fresh pages is slow
        t0 = time.perf_counter()
        mb = next(O.minibatches(flat, table, 1, perm=torch.arange(mb_size)))     # episode gather (buffer.py:90)
        O.train_minibatch(P, opt, ocfg, mb, 3e-4, 0.1, 1e-3)
        t_best = min(t_best, time.perf_counter() - t0)
        del mb
    t_mb = t_best * (mb_full / mb_size)
    n_opt = cfg["epochs"] * cfg["n_mini_batch"]
    update_s = T * t_roll + n_opt * t_mb
    detail = {"rollout_step_s": t_roll, "minibatch_step_s": t_mb, "optimiser_steps_per_update": n_opt,
              "extrapolation": {"rollout_steps_timed": rollout_steps, "rollout_steps_per_update": T,
                                "minibatch_samples_timed": mb_size, "minibatch_samples": mb_full,
                                "optimiser_steps_timed": 1, "optimiser_steps_per_update": n_opt,
                                "cpu_seconds_measured": rollout_steps * t_roll + t_best, "update_seconds_extrapolated": update_s,
                                "factor": update_s / max(1e-9, rollout_steps * t_roll + t_best)},
              "sample": "%d rollout steps (W=%d, in-process synthetic envs) + optimiser step on %d of the %d minibatch samples "
                        "(best of %d, scaled x%.0f) timed on %d threads; update = %d*rollout_step + %d*minibatch_step"
                        % (rollout_steps, W, mb_size, mb_full, reps, mb_full / mb_size, threads, T, n_opt)}
    detail["threads"] = threads
    return W * T / update_s, detail


def run_reference(args, cfg, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, detail, threads = [], None, None
    mb_full = cfg["n_workers"] * cfg["worker_steps"] // cfg["n_mini_batch"]
    # Every timed step = 8 rollout steps over all workers + ONE FULL-SIZE optimiser step (when the oracle can hold it: the
    # reference materialises (mb, M, B, D) + (mb, L, B, D)), extrapolated to an update; warm-up steps use a quarter-size batch.
    # A long run (the driver's --steps 20) bounds the total CPU time by timing the full-size step on every 4th step.
    window_gb = mb_full * (cfg["environment"]["max_episode_steps"] + cfg["transformer"]["memory_length"]) * \
        cfg["transformer"]["num_blocks"] * cfg["transformer"]["embed_dim"] * 4 / 1e9
    full_ok = window_gb * 3 < 0.5 * _host_mem_gb()
    for i in range(args.warmup + args.steps):
        timed = i >= args.warmup
        full = timed and full_ok and (args.steps <= 6 or (i - args.warmup) % 4 == 0)
        # the thread-count probe runs once (first step); later steps reuse its choice so K steps stay within minutes
        v, detail = cpu_reference_sample(cfg, rollout_steps=8 if timed else 4, seed=i, threads=threads,
                                         mb_sample=mb_full if full else max(64, mb_full // 4), reps=1)
        threads = detail["threads"]
        if timed:
            vals.append(v)
    value = float(np.mean(vals))
    cores = detail["threads"]
    line = {"impl": "reference", "metric": metric_name(name), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * cfg["n_workers"] * cfg["worker_steps"] / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_desc(cfg, name, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": detail["sample"],
                             "rollout_step_s": detail["rollout_step_s"], "minibatch_step_s": detail["minibatch_step_s"],
                             "extrapolation": detail["extrapolation"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _host_mem_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return float(ln.split()[1]) / 1e6
    except OSError:
        pass
    return 16.0


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, cfg, name):
    import parallel
    import trxl_native as native
    from device_feed import SyntheticDeviceFeed
    from trainer import PPOTrainer
    from utils import polynomial_decay
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    native.load()
    torch.set_num_threads(min(8, os.cpu_count() or 1))     # the trainer process only runs small host ops
    parallel.init_from_env("nccl")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dp = parallel.DataParallelContext(device)
    rank, world = dp.rank, dp.world_size
    env_cfg = cfg["environment"]
    if args.scaling == "strong":                 # the configuration's workers are partitioned over the ranks
        if cfg["n_workers"] % world:
            raise SystemExit("--scaling strong: n_workers=%d is not divisible by %d ranks" % (cfg["n_workers"], world))
        cfg = dict(cfg, n_workers=cfg["n_workers"] // world)
    W, T = cfg["n_workers"], cfg["worker_steps"]
    torch.manual_seed(1234 + rank)
    os.chdir(os.environ.get("TMPDIR", "/tmp"))

    # ---------------- e2e arm first (forks env processes before the big allocations) ----------------
    e2e = None
    e2e_skipped = None
    # one Python process per env: beyond ~16 processes per host CPU (or the host's free memory at ~0.4 GB each) the arm measures
    # the box, not the engine -- and c5's 512 processes got the whole bench OOM-killed on a 16-CPU box
    n_procs = cfg["n_workers"] * max(1, world)
    if not args.no_e2e and (n_procs > 16 * (os.cpu_count() or 1) or n_procs * 0.4 > 0.8 * _host_mem_gb()):
        e2e_skipped = "%d env processes on %d host CPUs / %.0f GB free: e2e arm skipped" % (n_procs, os.cpu_count() or 1, _host_mem_gb())
    if not args.no_e2e and e2e_skipped is None:
        tr = PPOTrainer(cfg, run_id="bench_e2e", device=device, summary_writer=False)
        sched = lambda u: (polynomial_decay(**_s(cfg["learning_rate_schedule"]), current_step=u),   # noqa: E731
                           polynomial_decay(**_s(cfg["clip_range_schedule"]), current_step=u),
                           polynomial_decay(**_s(cfg["beta_schedule"]), current_step=u))
        def one_update(trainer, u):
            lr, clip, beta = sched(u)
            trainer._sample_training_data()
            trainer.buffer.prepare_batch_dict()
            trainer._train_epochs(lr, clip, beta)
        for u in range(min(args.warmup, 2)):
            one_update(tr, u)
        tr.timers = {"rollout": 0.0, "train": 0.0, "env": 0.0}
        dp.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, args.e2e_steps))
        for u in range(n_e2e):
            one_update(tr, u)
        torch.cuda.synchronize(); dp.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        dp.max_(dt)
        obs_bytes = int(np.prod(env_cfg["obs_shape"])) * 4
        e2e = {"value": world * W * T * n_e2e / float(dt), "unit": UNIT,
               "h2d_bytes_per_step": T * (W * obs_bytes + 2 * W * 8), "d2h_bytes_per_step": T * W * 8 + 40 * 4 * 32,
               "updates_timed": n_e2e,
               "env_transport": ("1 process per env; observations written into a shared pinned slab that one kernel per step reads in "
                                 "place over PCIe, actions written by the sampling kernel straight into pinned host memory, "
                                 "actions/acks through shared-memory stepping (worker.py); %d worker groups overlap env stepping "
                                 "with the other group's forward" % len(tr._groups)
                                 if tr._control is not None else "1 process per env, pipes (worker.py)"),
               "env_wait_s_per_update": tr.timers["env"] / n_e2e, "rollout_s_per_update": tr.timers["rollout"] / n_e2e,
               "train_s_per_update": tr.timers["train"] / n_e2e}
        tr.close(exit_process=False)
        del tr
        torch.cuda.empty_cache()

    # ---------------- device-resident arm ----------------
    tr = PPOTrainer(cfg, run_id="bench", device=device, workers=[], summary_writer=False)
    tr.device_feed = SyntheticDeviceFeed(W, T, tuple(env_cfg["obs_shape"]), env_cfg["max_episode_steps"],
                                         env_cfg.get("min_episode_steps"), seed=100 + rank, device=device)

    def update(u):
        lr = polynomial_decay(**_s(cfg["learning_rate_schedule"]), current_step=u)
        clip = polynomial_decay(**_s(cfg["clip_range_schedule"]), current_step=u)
        beta = polynomial_decay(**_s(cfg["beta_schedule"]), current_step=u)
        tr._sample_training_data()
        tr.buffer.prepare_batch_dict()
        return tr._train_epochs(lr, clip, beta)

    for u in range(args.warmup):
        update(u)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    native.profile_enable(True)
    launches0 = native.launch_count()
    tr.timers = {"rollout": 0.0, "train": 0.0, "env": 0.0}
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dp.barrier(); torch.cuda.synchronize()
    start.record()
    for u in range(args.steps):
        stats, _ = update(args.warmup + u)
    stop.record()
    torch.cuda.synchronize(); dp.barrier()
    ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=device)
    dp.max_(ms)
    ms = float(ms)
    launches = native.launch_count() - launches0
    clock_info = clocks.stop() if rank == 0 else None
    mb = W * T // cfg["n_mini_batch"]
    t = cfg["transformer"]
    prof = {k: native.profile_read(k, mb) for k in range(4)}
    tiles = {k: native.profile_aux(k, mb) for k in (2, 3)}
    native.profile_enable(False)
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    roofline = attention_roofline(cfg, mb, prof, tiles, peaks, ms)
    value = world * W * T * args.steps / (ms * 1e-3)
    line = {"metric": metric_name(name), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_desc(cfg, name, world, args.scaling), "clocks": clock_info, "gpu_launches": launches,
            "roofline": roofline, "e2e": e2e,
            "breakdown_s_per_update": {"rollout": tr.timers["rollout"] / args.steps, "train": tr.timers["train"] / args.steps},
            "last_stats": [float(x) for x in np.mean(np.array(stats, dtype=np.float64), axis=0)]}
    if e2e_skipped:
        line["e2e_skipped"] = e2e_skipped
    if world == 1 and not args.no_cpu_baseline:
        mb_full = W * T // cfg["n_mini_batch"]
        window_gb = mb_full * (env_cfg["max_episode_steps"] + t["memory_length"]) * t["num_blocks"] * t["embed_dim"] * 4 / 1e9
        full_ok = window_gb * 3 < 0.5 * _host_mem_gb()
        v, detail = cpu_reference_sample(cfg, rollout_steps=8, mb_sample=mb_full if full_ok else max(64, mb_full // 8), reps=1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": detail["threads"], "host_cores": os.cpu_count(),
                                "kind": "port", "sample": detail["sample"], "extrapolation": detail["extrapolation"],
                                "rollout_step_s": detail["rollout_step_s"], "minibatch_step_s": detail["minibatch_step_s"]}
    print(json.dumps(line))


def _s(schedule):
    return {"initial": schedule["initial"], "final": schedule["final"], "max_decay_steps": schedule["max_decay_steps"],
            "power": schedule["power"]}


def ncu_traffic_bytes(kernel, profile):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first launch of `kernel` in a committed ncu summary (bytes), or None."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        lines = open(profile).read().splitlines()
    except OSError:
        return None
    total, inside, found = 0.0, False, 0
    for ln in lines:
        if ln.startswith("["):
            if inside and found:
                break
            inside = kernel in ln
            continue
        parts = ln.split()
        if inside and len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and parts[2] in scale:
            total += float(parts[1].replace(",", "")) * scale[parts[2]]
            found += 1
    return total if found == 2 else None


def attention_roofline(cfg, mb, prof, tiles, peaks, step_ms_total):
    """`roofline` object for the window attention of the training minibatches (the kernel BASELINE.json's north_star grades),
    timed with CUDA events on its launch stream inside the timed region.

    Episode-grouped tensor-core path (post-/no-LN, relative/no PE -- c2, c3, c4): one "launch" = the forward of one block =
    grouped GEMM S = QK.Xpe^T, window softmax, grouped GEMM ctx = P.Xpe.  bound = tensor.  `achieved` counts the EXECUTED
    tensor-core work (3 TF32 passes over tiles x 128 x M x D MACs per GEMM, padding included), the peak is the measured dense
    bf16 rate halved (TF32 issues at half the bf16 rate).  `reference_equivalent` is SURVEY.md §8(d)'s algorithmic figure for
    the same launch, N (4 L D^2 + 4 L D + 4 D^2) FLOP: the work the reference's formulation would need, most of which the
    query-side fold and the per-episode grouping eliminate -- reported, not claimed as utilisation.
    Per-sample streaming path (pre-LN or learned PE -- c1, c5): bound = hbm on the 4 L D algorithmic bytes per (sample, block)."""
    t = cfg["transformer"]
    L, D, H, M = t["memory_length"], t["embed_dim"], t["num_heads"], cfg["environment"]["max_episode_steps"]
    ref_flops = mb * (4.0 * L * D * D + 4.0 * L * D + 4.0 * D * D)
    bf16_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json (measured)" if peaks else "fallback (B200_PROFILING.md)"
    fwd_ms, fwd_n, _ = prof[2]
    if fwd_n:                                   # grouped tensor-core path
        bwd_ms, bwd_n, _ = prof[3]
        avg = fwd_ms / fwd_n
        exec_flops = tiles[2] / fwd_n * 128.0 * M * D * 2.0 * 2.0 * 3.0
        achieved = exec_flops / (avg * 1e-3) / 1e12
        peak = bf16_peak / 2.0
        return {"kernel": "episode-grouped attention forward of one block (N=%d): grouped tma_gemm_kernel (256-wide tiles at M > 128) S=QK.X^T, "
                          "attn_softmax_kernel, grouped tma_gemm_kernel ctx=P.X" % mb,
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": src + ": sustained dense bf16 / 2 (TF32 rate)",
                "executed_flops_per_launch": exec_flops, "tiles_per_launch": tiles[2] / fwd_n, "avg_launch_ms": avg,
                "launches_timed": fwd_n,
                "reference_equivalent": {"flops_per_launch": ref_flops, "tflops": ref_flops / (avg * 1e-3) / 1e12,
                                         "note": "SURVEY.md 8(d) algorithmic FLOPs of the reference's formulation; eliminated work, not utilisation"},
                "useful_flops_per_launch": mb * H * 4.0 * L * D,
                "traffic": ncu_traffic_bytes("tma_gemm_kernel", os.path.join(ROOT, "profiles", "r2_ncu_grouped_attention.txt")),
                "traffic_source": "profiles/r2_ncu_grouped_attention.txt (ncu --set full, dram read + write of the S GEMM launch)",
                "bwd": {"avg_launch_ms": bwd_ms / max(1, bwd_n), "launches_timed": bwd_n,
                        "achieved": (tiles[3] / max(1, bwd_n) * 128.0 * M * D * 12.0) / (bwd_ms / max(1, bwd_n) * 1e-3) / 1e12 if bwd_n else None},
                "share_of_step": (fwd_ms + bwd_ms) / step_ms_total if step_ms_total else None}
    fwd_ms, fwd_n, _ = prof[0]
    bwd_ms, bwd_n, _ = prof[1]
    bytes_per_launch = mb * 4.0 * L * D
    achieved = bytes_per_launch / (fwd_ms / max(1, fwd_n) * 1e-3) / 1e9 if fwd_n else None
    return {"kernel": "window_attn_fwd_kernel (per-sample streaming, N=%d samples x 1 block)" % mb, "bound": "hbm",
            "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": (achieved / hbm_peak) if achieved else None,
            "peak_source": src, "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": fwd_ms / max(1, fwd_n),
            "launches_timed": fwd_n, "traffic": ncu_traffic_bytes("window_attn_fwd_kernel", os.path.join(ROOT, "profiles", "r1_ncu_attn.txt")),
            "traffic_source": "profiles/r1_ncu_attn.txt",
            "note": "issue-bound in practice (r1 ncu: DRAM 4 %, L2 26 %, issue-active 59 %); most window rows hit L2",
            "reference_equivalent": {"flops_per_launch": ref_flops},
            "bwd": {"avg_launch_ms": bwd_ms / max(1, bwd_n), "launches_timed": bwd_n},
            "share_of_step": (fwd_ms + bwd_ms) / step_ms_total if step_ms_total else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_minigrid_synthetic",
                    help="configs/<name>.yaml: c1_poc_synthetic, c2_cartpole_synthetic, c3_minigrid_synthetic (BASELINE.json's metric), "
                         "c4_minigrid_gtrxl_synthetic, c5_mortar_synthetic")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU runs the configuration's n_workers; strong: n_workers is partitioned over the GPUs")
    ap.add_argument("--e2e-steps", type=int, default=2, help="PPO updates timed for the e2e (env-process) arm")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = load_workload(args.workload)
    if args.impl == "reference":
        run_reference(args, cfg, args.workload)
    else:
        run_b200(args, cfg, args.workload)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
