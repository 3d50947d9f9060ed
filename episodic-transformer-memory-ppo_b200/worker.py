"""One environment per OS process behind a pipe -- the reference's rollout transport (worker.py:6-62),
kept in Python as BASELINE.json's north_star asks.  Protocol: ("reset", None) -> obs,
("step", action) -> (obs, reward, done, info), ("close", None).

Optional zero-copy observations: when a worker is given a slot of a shared-memory slab
(``obs_slab[index]``, a torch tensor in shared memory that the trainer also registers as pinned host
memory), it writes every observation there and sends ``None`` in the obs position of its reply.  The
control protocol is unchanged; only the 85 KB-per-step image payload stops being pickled through
the pipe, and the trainer DMAs the slab to the GPU directly."""
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


def worker_process(remote, config, obs_slab=None, index=0):
    import os
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from utils import create_env
    try:
        env = create_env(config)
    except KeyboardInterrupt:
        return
    slot = None if obs_slab is None else obs_slab.numpy()[index]

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
    while True:
        try:
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
            raise WorkerException(e)


class Worker:
    """Runs one environment in its own process; ``child`` is the parent's end of the pipe."""
    child: multiprocessing.connection.Connection
    process: multiprocessing.Process

    def __init__(self, env_config, obs_slab=None, index=0):
        ctx = multiprocessing.get_context("fork")
        self.child, parent = ctx.Pipe()
        self.process = ctx.Process(target=worker_process, args=(parent, env_config, obs_slab, index), daemon=True)
        self.process.start()
