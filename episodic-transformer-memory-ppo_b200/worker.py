"""One environment per OS process behind a pipe -- the reference's rollout transport (worker.py:6-62),
kept in Python as BASELINE.json's north_star asks.  Protocol: ("reset", None) -> obs,
("step", action) -> (obs, reward, done, info), ("close", None).

Optional zero-copy observations: when a worker is given a slot of a shared-memory slab
(``obs_slab[index]``, a torch tensor in shared memory that the trainer also registers as pinned host
memory), it writes every observation there and sends ``None`` in the obs position of its reply.  The
control protocol is unchanged; only the 85 KB-per-step image payload stops being pickled through
the pipe, and the trainer DMAs the slab to the GPU directly.

Optional shared-memory stepping (``control``): per-step traffic (action in; reward, done out) moves
through small shared arrays guarded by a command / acknowledge counter pair instead of 2 pickled pipe
messages per worker per step.  A worker that finishes an episode resets its environment itself
(exactly what the trainer would ask for next, reference trainer.py:201-203), publishes the reset
observation, and sends the episode ``info`` dict through the pipe, which stays the control channel
for everything else."""
import multiprocessing
import multiprocessing.connection
import sys
import traceback


class WorkerException(Exception):
    """Raised in the worker process; carries the formatted traceback of the original error."""

    def __init__(self, ee):
        self.ee = ee
        self.tb = "".join(traceback.format_exception(*sys.exc_info()))
        super().__init__("%s\n%s" % (ee, self.tb))

    def re_raise(self):
        raise self.ee


FUTEX_WORD_STRIDE = 16          # int32 words between two groups' futex words (one 64-byte line each)


def futex_available():
    """Shared-memory futexes through libc's syscall(2): Linux on x86-64 (syscall number 202)."""
    import platform
    return sys.platform.startswith("linux") and platform.machine() in ("x86_64", "AMD64")


class _Futex:
    """FUTEX_WAIT / FUTEX_WAKE on an int32 that lives in memory shared between the trainer and its env processes (a torch tensor
    in shared memory, mapped at the same address after fork).  One wake call releases every waiter of a worker group."""
    SYS_FUTEX, WAIT, WAKE = 202, 0, 1

    def __init__(self):
        import ctypes

        class Timespec(ctypes.Structure):
            _fields_ = [("tv_sec", ctypes.c_long), ("tv_nsec", ctypes.c_long)]
        self._ct = ctypes
        self._ts = Timespec
        self._libc = ctypes.CDLL(None, use_errno=True)
        self._libc.syscall.restype = ctypes.c_long

    def wait(self, addr, expected, timeout_s):
        """Sleep until woken, until ``*addr != expected`` (returns at once), or until the timeout."""
        ct = self._ct
        ts = self._ts(int(timeout_s), int((timeout_s % 1.0) * 1e9))
        self._libc.syscall(ct.c_long(self.SYS_FUTEX), ct.c_void_p(addr), ct.c_int(self.WAIT), ct.c_int(expected), ct.byref(ts),
                           ct.c_void_p(0), ct.c_int(0))

    def wake(self, addr, n=1 << 20):
        ct = self._ct
        self._libc.syscall(ct.c_long(self.SYS_FUTEX), ct.c_void_p(addr), ct.c_int(self.WAKE), ct.c_int(n), ct.c_void_p(0),
                           ct.c_void_p(0), ct.c_int(0))


def make_control(n_workers, n_branches, blocking=False, futex=None):
    """Shared arrays of the stepping fast path (torch tensors in shared memory, inherited by fork).
    ``blocking=True``: workers sleep in the kernel between steps instead of spinning on their command counter (needed as
    soon as env processes outnumber idle cores: spinning siblings slow the envs that are actually stepping).  They sleep on a
    futex word per worker group (one wake system call releases the whole group) where that is available, otherwise on one
    semaphore per worker."""
    import torch
    if futex is None:
        futex = futex_available()
    use_futex = bool(blocking and futex)
    sems = [multiprocessing.get_context("fork").Semaphore(0) for _ in range(n_workers)] if (blocking and not use_futex) else None
    return {"sems": sems, "blocking": bool(blocking),
            "futex": torch.zeros(64 * FUTEX_WORD_STRIDE, dtype=torch.int32).share_memory_() if use_futex else None,
            "actions": torch.zeros((n_workers, n_branches), dtype=torch.int64).share_memory_(),
            "rewards": torch.zeros(n_workers, dtype=torch.float32).share_memory_(),
            "dones": torch.zeros(n_workers, dtype=torch.uint8).share_memory_(),
            "has_info": torch.zeros(n_workers, dtype=torch.uint8).share_memory_(),
            "cmd": torch.zeros(n_workers, dtype=torch.int64).share_memory_(),
            "ack": torch.zeros(n_workers, dtype=torch.int64).share_memory_()}


def physical_cpus():
    """One logical CPU per physical core among the CPUs this process may run on (Linux topology files; falls back to the
    plain affinity list).  Used to give every spinning env worker a core of its own: two busy-waiting hyper-thread
    siblings slow each other and whichever of them is actually stepping its environment."""
    import os
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return []
    cores, seen = [], set()
    for cpu in allowed:
        try:
            with open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % cpu) as f:
                sib = f.read().strip()
        except OSError:
            sib = str(cpu)
        if sib not in seen:
            seen.add(sib)
            cores.append(cpu)
    return cores


def worker_process(remote, config, obs_slab=None, index=0, control=None, cpu=None, group=0):
    import os
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    if cpu is not None:
        try:
            os.sched_setaffinity(0, {int(cpu)})
        except (AttributeError, OSError):
            pass
    from utils import create_env
    try:
        env = create_env(config)
    except KeyboardInterrupt:
        return
    slot = None if obs_slab is None else obs_slab.numpy()[index]
    if slot is not None and hasattr(env, "set_observation_buffer"):
        env.set_observation_buffer(slot)          # the env draws observations straight into the shared slab

    def do_step(data):
        obs, reward, done, info = env.step(data)
        if slot is not None:
            slot[...] = obs
            obs = None
        return obs, reward, done, info

    def do_reset(data):
        obs = env.reset()
        if slot is not None:
            slot[...] = obs
            obs = None
        return obs
    handlers = {"step": do_step, "reset": do_reset}
    if control is not None:
        c_act, c_rew = control["actions"].numpy()[index], control["rewards"].numpy()
        c_done, c_info = control["dones"].numpy(), control["has_info"].numpy()
        c_cmd, c_ack = control["cmd"].numpy(), control["ack"].numpy()
        last, idle = int(c_cmd[index]), 0
        sem = control["sems"][index] if control.get("sems") else None
        fut, fword, faddr = None, None, 0
        if control.get("futex") is not None:
            fut = _Futex()
            fword = control["futex"].numpy()
            faddr = control["futex"].data_ptr() + 4 * FUTEX_WORD_STRIDE * group
    while True:
        try:
            if control is not None and fut is not None:
                # futex variant: sleep until the trainer bumps this group's generation word (one wake call per group)
                seq = int(c_cmd[index])
                if seq == last:
                    gen = int(fword[FUTEX_WORD_STRIDE * group])      # read the generation BEFORE re-checking the command
                    seq = int(c_cmd[index])
                    if seq == last:
                        fut.wait(faddr, gen, 0.02)
                        seq = int(c_cmd[index])
                if seq != last:
                    obs, reward, done, info = env.step(c_act.copy())
                    if info:
                        remote.send(info)
                        obs = env.reset()
                    if slot is not None and obs is not slot:
                        slot[...] = obs
                    c_rew[index], c_done[index], c_info[index] = reward, 1 if done else 0, 1 if info else 0
                    last = seq
                    c_ack[index] = seq
                    continue
                if not remote.poll(0):
                    continue
            elif control is not None and sem is not None:
                # blocking variant: sleep on the semaphore; the pipe (control messages) is polled every 20 ms
                if sem.acquire(timeout=0.02):
                    seq = int(c_cmd[index])
                    obs, reward, done, info = env.step(c_act.copy())
                    if info:
                        remote.send(info)
                        obs = env.reset()
                    if slot is not None:
                        slot[...] = obs
                    c_rew[index], c_done[index], c_info[index] = reward, 1 if done else 0, 1 if info else 0
                    last = seq
                    c_ack[index] = seq
                    continue
                if not remote.poll(0):
                    continue
            elif control is not None:
                # spin on the command counter; fall back to the pipe for control messages; back off when idle
                seq = int(c_cmd[index])
                if seq != last:
                    obs, reward, done, info = env.step(c_act.copy())
                    if info:
                        remote.send(info)
                        obs = env.reset()
                    if slot is not None:
                        slot[...] = obs
                    c_rew[index], c_done[index], c_info[index] = reward, 1 if done else 0, 1 if info else 0
                    last, idle = seq, 0
                    c_ack[index] = seq                   # publish last: x86 keeps the stores above ordered before it
                    continue
                idle += 1
                if idle < 20000 and not (idle % 64 == 0 and remote.poll(0)):
                    continue
                if not remote.poll(0.0002 if idle >= 20000 else 0):
                    continue
            cmd, data = remote.recv()
            if cmd == "close":
                remote.send(env.close())
                remote.close()
                return
            if cmd not in handlers:
                raise NotImplementedError(cmd)
            remote.send(handlers[cmd](data))
        except (EOFError, KeyboardInterrupt):
            return
        except Exception as e:  # noqa: BLE001 -- surfaced to the parent as in the reference
            err = WorkerException(e)
            try:
                remote.send(("worker_error", str(err)))      # the shared-memory stepping path polls for a dead worker
            except Exception:  # noqa: BLE001
                pass
            raise err


class Worker:
    """Runs one environment in its own process; ``child`` is the parent's end of the pipe."""
    child: multiprocessing.connection.Connection
    process: multiprocessing.Process

    def __init__(self, env_config, obs_slab=None, index=0, control=None, cpu=None, group=0):
        ctx = multiprocessing.get_context("fork")
        self.child, parent = ctx.Pipe()
        self.process = ctx.Process(target=worker_process, args=(parent, env_config, obs_slab, index, control, cpu, group), daemon=True)
        self.process.start()
