"""Agent base: the duck-typed contract ``rl/runner`` relies on (``rl/agent/abc.py:16-55``) plus the
glue between the Python-visible state (``rl.nn`` modules, counters) and the device update engine."""
from __future__ import annotations

import copy
import weakref
import ctypes as C
import os
import pickle
from abc import ABC
from pathlib import Path
from typing import TypeVar

import numpy as np
import torch
from torch import nn

from ... import _lib

AGENT = TypeVar("AGENT", bound="Agent")


_stream = _lib.current_stream_ptr


class LazyScalar:
    """A loss value that stays on the device until somebody looks at it.

    The reference returns ``float(loss.cpu())`` per update, i.e. one host sync per loss
    (td7.py:303,315,324).  ``train_ops`` here returns objects that behave like floats for what the
    runner does with them (``sum``, ``/``, ``format(v, '.4f')`` -- run.py:47-84) and fetch the
    value from the engine's info ring on first use."""

    __slots__ = ("_agent", "_step", "_col", "_val")

    def __init__(self, agent, step, col):
        self._agent, self._step, self._col, self._val = agent, step, col, None

    def __float__(self):
        if self._val is None:
            self._val = self._agent._info_value(self._step, self._col)
            self._agent = None
        return self._val

    def __add__(self, o): return float(self) + float(o)
    __radd__ = __add__
    def __sub__(self, o): return float(self) - float(o)
    def __rsub__(self, o): return float(o) - float(self)
    def __mul__(self, o): return float(self) * float(o)
    __rmul__ = __mul__
    def __truediv__(self, o): return float(self) / float(o)
    def __rtruediv__(self, o): return float(o) / float(self)
    def __neg__(self): return -float(self)
    def __lt__(self, o): return float(self) < float(o)
    def __gt__(self, o): return float(self) > float(o)
    def __le__(self, o): return float(self) <= float(o)
    def __ge__(self, o): return float(self) >= float(o)
    def __eq__(self, o): return float(self) == float(o)
    def __hash__(self): return hash(float(self))
    def __format__(self, spec): return format(float(self), spec)
    def __repr__(self): return repr(float(self))


class _DeviceView:
    """``__cuda_array_interface__`` holder for a float32 region of engine-owned device memory."""

    def __init__(self, ptr: int, shape, owner):
        self._owner = owner  # keeps the engine (and with it the memory) alive
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2, "strides": None}


class Agent(ABC):
    """Common machinery of the TD7 / TD3 / SAC drop-ins."""

    # subclasses: ordered (attribute name, layer names) as the engine numbers its networks
    NETS: list = []
    INFO_KEYS: list = []
    TRAINED: set = set()

    # ------------------------------------------------------------------ engine lifecycle
    def _engine_init(self, device="cuda"):
        if not torch.cuda.is_available():
            raise _lib.B200RLError(f"{self!r} needs a CUDA device: the update path has no CPU fallback")
        self._lib = _lib.lib()
        self.device = torch.device(device)
        self._h = None
        self._dirty_host = True   # host modules newer than device blobs
        self._dirty_dev = False   # device blobs newer than host modules
        self._info_cache = {}
        self._inbox = None
        self._create_engine()
        self._push()

    def _create_engine(self):
        raise NotImplementedError

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.b200rl_agent_destroy(h)
            except Exception:  # noqa: BLE001
                pass

    # ------------------------------------------------------------------ host <-> device weights
    def _layout(self, net):
        numel, nl = C.c_int64(), C.c_int()
        _lib.check(self._lib.b200rl_agent_net_info(self._h, net, C.byref(numel), C.byref(nl)))
        out = []
        for i in range(nl.value):
            wo, bo, o, n, ld = C.c_int64(), C.c_int64(), C.c_int(), C.c_int(), C.c_int()
            _lib.check(self._lib.b200rl_agent_layer_info(self._h, net, i, C.byref(wo), C.byref(bo), C.byref(o), C.byref(n), C.byref(ld)))
            out.append((wo.value, bo.value, o.value, n.value, ld.value))
        return numel.value, out

    @staticmethod
    def _linears(module: nn.Module, names):
        sd = dict(module.named_modules())
        return [sd[n] for n in names]

    def _write_net(self, net_id, module, names, group=0):
        numel, layers = self._layout(net_id)
        blob = np.zeros(numel, np.float32)
        for (wo, bo, o, n, ld), lin in zip(layers, self._linears(module, names)):
            w = lin.weight.detach().cpu().numpy()
            assert w.shape == (o, n), f"network shape mismatch: {w.shape} vs {(o, n)}"
            blob[wo:wo + o * ld].reshape(o, ld)[:, :n] = w
            blob[bo:bo + o] = lin.bias.detach().cpu().numpy()
        _lib.check(self._lib.b200rl_agent_write(self._h, net_id, 0, group, blob.ctypes.data, _stream()))

    def _read_blob(self, net_id, which=0, group=0):
        numel, layers = self._layout(net_id)
        blob = np.empty(numel, np.float32)
        _lib.check(self._lib.b200rl_agent_read(self._h, net_id, which, group, blob.ctypes.data, _stream()))
        return blob, layers

    def _read_net_into(self, net_id, module, names, group=0):
        blob, layers = self._read_blob(net_id, 0, group)
        with torch.no_grad():
            for (wo, bo, o, n, ld), lin in zip(layers, self._linears(module, names)):
                lin.weight.copy_(torch.from_numpy(blob[wo:wo + o * ld].reshape(o, ld)[:, :n].copy()))
                lin.bias.copy_(torch.from_numpy(blob[bo:bo + o].copy()))

    def _push(self):
        """host modules -> device (seed 0; stacked seeds are initialised through seed_modules())."""
        for nid, (attr, names) in enumerate(self.NETS):
            self._write_net(nid, self._mods[attr], names)
        self._dirty_host = False

    def _pull(self):
        if self._dirty_dev and self._h:
            for nid, (attr, names) in enumerate(self.NETS):
                self._read_net_into(nid, self._mods[attr], names)
            self._dirty_dev = False

    def _module(self, attr):
        self._pull()
        return self._mods[attr]

    def read_infos(self, first_step: int, count: int, seed: int = 0) -> np.ndarray:
        """[count, len(INFO_KEYS)] info rows of stacked seed `seed` for updates first_step.. (1-based)."""
        buf = np.empty((count, len(self.INFO_KEYS)), np.float32)
        _lib.check(self._lib.b200rl_agent_read_info(self._h, first_step, count, seed, buf.ctypes.data, _stream()))
        return buf

    def info_dicts(self, first_step: int, count: int, seed: int = 0) -> list[dict]:
        """The info dicts train_ops returned for updates first_step .. first_step+count-1 of stacked seed `seed`, as plain
        floats (None where the reference returns None), from ONE read of the engine's info ring -- what an episode-end
        ``log_train_infos`` (run.py:53-84) needs after a ``train_ops_n`` burst."""
        rows = self.read_infos(first_step, count, seed)
        out = []
        sync, self.sync_info = getattr(self, "sync_info", False), False
        try:
            for i in range(count):
                step = first_step + i
                d = self._info_dict(step, self._actor_step(step))
                out.append({k: (None if v is None else float(rows[i][v._col])) for k, v in d.items()})
        finally:
            self.sync_info = sync
        return out

    def export_to(self, reference_agent) -> None:
        """Copy the current networks into an agent of the REFERENCE's classes (same attribute names, same
        ``state_dict`` keys), e.g. so that ``rl/replayer.py:43-45`` can render a policy trained here."""
        self._pull()
        for attr, _ in self.NETS:
            getattr(reference_agent, attr).load_state_dict(getattr(self, attr).state_dict())

    def seed_population(self, base_seed: int) -> None:
        """Give every stacked seed its own initial networks: seed g is built exactly like a fresh
        single agent under ``torch.manual_seed(base_seed + g)`` (the reference runs seeds as separate
        processes, scripts/td7_exp.sh:1-4)."""
        for g in range(self.n_seeds):
            torch.manual_seed(base_seed + g)
            self.set_seed_weights(g, self._fresh_modules())
        self._dirty_dev, self._dirty_host = True, False

    def _fresh_modules(self) -> dict:
        raise NotImplementedError

    def set_seed_weights(self, seed: int, modules: dict):
        """Load the networks of stacked seed `seed` from a dict attr -> module (population setup)."""
        for nid, (attr, names) in enumerate(self.NETS):
            self._write_net(nid, modules[attr], names, group=seed)

    # ------------------------------------------------------------------ info ring
    def _info_value(self, step: int, col: int) -> float:
        """One info value.  For the newest update of a single agent the engine posts every loss to a pinned host mailbox
        the moment its kernel has run (b200rl_agent_wait_info), so ``float(info["train/q_fn"])`` returns while the
        backward pass of that update is still executing; anything else comes from the device ring."""
        row = self._info_cache.get(step)
        if row is None and step == self._n_calls and self.n_seeds == 1 and col < 8:
            v = C.c_float()
            if self._lib.b200rl_agent_wait_info(self._h, step, col, C.byref(v)) == 0:
                return float(v.value)
        return float(self._info_row(step)[col])

    def _info_row(self, step):
        """Row `step` (1-based train_ops count) of the device info ring; one D2H fetches every row
        from `step` to the newest, so reading a burst of infos costs a single sync."""
        row = self._info_cache.get(step)
        if row is None:
            n_calls = self._n_calls
            if step <= n_calls - 1024:
                raise _lib.B200RLError("this update's info was overwritten in the 1024-entry device ring; read it sooner or use sync_info=True")
            cols = len(self.INFO_KEYS)
            count = n_calls - step + 1
            buf = np.empty((count, cols), np.float32)
            _lib.check(self._lib.b200rl_agent_read_info(self._h, step, count, 0, buf.ctypes.data, _stream()))
            for i in range(count):
                self._info_cache[step + i] = buf[i]
            if len(self._info_cache) > 4096:
                for k in sorted(self._info_cache)[:-2048]:
                    del self._info_cache[k]
            row = self._info_cache[step]
        return row

    # ------------------------------------------------------------------ reference contract
    def make_optimizers(self, *a, **k) -> None:
        """Adam lives inside the update engine (fused into the weight-gradient kernels)."""

    def zero_grad(self) -> None:
        """Gradients never persist between steps in the engine; nothing to zero."""

    def to(self, device) -> "Agent":
        """td7.py:100-112.  The engine always lives on the GPU; ``to('cpu')`` (used by ``save``)
        only materialises the current weights into the host-side modules."""
        device = torch.device(device)
        if device.type == "cpu":
            self._pull()
        self._host_device = device
        return self

    def load_state_dict(self, agent: "Agent") -> None:
        """td7.py:114-125: copy every network of `agent` (one device-to-device copy per blob)."""
        agent._flush()
        self._flush()
        _lib.check(self._lib.b200rl_agent_copy_from(self._h, agent._h, 0, _stream()))
        self._dirty_dev = True

    def _flush(self):
        if self._dirty_host:
            self._push()

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        self._pull()
        for k, v in self.__dict__.items():
            if k in ("_h", "_lib", "_info_cache", "_inbox"):
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new._lib = self._lib
        new._h = None
        new._inbox = None
        new._info_cache = {}
        new._create_engine()
        _lib.check(new._lib.b200rl_agent_copy_from(new._h, self._h, 1, _stream()))
        new._dirty_dev, new._dirty_host = True, False
        return new

    # pickling: whole-object pickle like the reference (abc.py:38-55), with the device state
    # materialised as plain tensors
    def __getstate__(self):
        self._pull()
        d = {k: v for k, v in self.__dict__.items() if k not in ("_h", "_lib", "_info_cache", "_inbox")}
        extra = {}
        if self._h:
            for nid, (attr, _) in enumerate(self.NETS):
                if attr in self.TRAINED:
                    extra[attr] = (self._read_blob(nid, 1)[0], self._read_blob(nid, 2)[0])
            n_runs, steps = C.c_int64(), (C.c_int64 * 3)()
            _lib.check(self._lib.b200rl_agent_get_counters(self._h, C.byref(n_runs), steps))
            extra["__counters__"] = (n_runs.value, list(steps))
            extra.update(self._extra_state())
        d["_pickled_engine_state"] = extra
        return d

    def __setstate__(self, d):
        extra = d.pop("_pickled_engine_state", {})
        self.__dict__.update(d)
        self._h, self._info_cache, self._inbox = None, {}, None
        self._lib = _lib.lib()
        self._create_engine()
        self._push()
        for nid, (attr, _) in enumerate(self.NETS):
            if attr in extra:
                m, v = extra[attr]
                _lib.check(self._lib.b200rl_agent_write(self._h, nid, 1, 0, m.ctypes.data, _stream()))
                _lib.check(self._lib.b200rl_agent_write(self._h, nid, 2, 0, v.ctypes.data, _stream()))
        if "__counters__" in extra:
            n_runs, steps = extra["__counters__"]
            _lib.check(self._lib.b200rl_agent_set_counters(self._h, n_runs, (C.c_int64 * 3)(*steps), _stream()))
            self._n_calls = n_runs
        self._restore_extra_state(extra)

    def _extra_state(self):
        return {}

    def _restore_extra_state(self, extra):
        pass

    def save(self, path) -> None:
        """abc.py:38-47"""
        path = Path(path)
        os.makedirs(path.parent, exist_ok=True)
        with open(path, "wb") as fh:
            pickle.dump(self, fh)

    @staticmethod
    def load(path):
        """abc.py:50-55"""
        assert os.path.isfile(path)
        with open(path, "rb") as fh:
            return pickle.load(fh)

    # ------------------------------------------------------------------ the update
    def _make_inbox(self):
        """Tensors aliasing the engine's fixed input buffers (single seed): what replay.sample() returns once
        this agent is the buffer's sink, so that train_ops(batch) needs no staging copy."""
        ss, ac, rw, nd, gs = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        _lib.check(self._lib.b200rl_agent_io(self._h, C.byref(ss), C.byref(ac), C.byref(rw), C.byref(nd), C.byref(gs)))
        B, S, A = self.batch_size, self.state_dim, self.action_dim
        view = lambda ptr, shape: torch.as_tensor(_DeviceView(ptr, shape, self), device=self.device)  # noqa: E731
        return {"state": view(ss.value, (B, S)), "action": view(ac.value, (B, A)), "reward": view(rw.value, (B, 1)),
                "next_state": view(ss.value + 4 * B * S, (B, S)), "done": view(nd.value, (B, 1))}

    def _adopt_sink(self, replay_buffer):
        if replay_buffer is None or self.n_seeds != 1 or not getattr(replay_buffer, "zero_copy", False):
            return
        if getattr(replay_buffer, "n_seeds", 1) != 1 or not hasattr(replay_buffer, "_sink"):
            return
        cur = replay_buffer._sink() if replay_buffer._sink is not None else None
        if cur is not self:
            if self._inbox is None:
                self._inbox = self._make_inbox()
            replay_buffer._sink = weakref.ref(self)

    def _stage(self, batch):
        if batch is self._inbox and batch is not None:
            return batch  # sampled straight into the engine's input buffers
        dev = self.device
        t = {k: torch.as_tensor(v, dtype=torch.float32).to(dev, non_blocking=True).contiguous() for k, v in batch.items()}
        _lib.check(self._lib.b200rl_agent_stage_batch(self._h, 0, t["state"].data_ptr(), t["action"].data_ptr(), t["reward"].data_ptr(),
                                                      t["next_state"].data_ptr(), t["done"].data_ptr(), _stream()))
        return t

    def _update(self, replay_buffer, noise, sink=None):
        self._flush()
        self._adopt_sink(sink)
        rh = getattr(replay_buffer, "_h", None) if replay_buffer is not None else None
        nz = None
        if noise is not None:
            nz = torch.as_tensor(noise, dtype=torch.float32).to(self.device).contiguous()
        _lib.check(self._lib.b200rl_agent_update(self._h, rh, nz.data_ptr() if nz is not None else None, _stream()))
        self._n_calls += 1
        self._dirty_dev = True

    def train_ops_n(self, n_ops: int, replay_buffer) -> list[dict]:
        """``run_train_ops(n_ops=k)`` (run.py:87-96) without leaving the device: k x {sample, update}."""
        self._flush()
        first = self._n_calls + 1
        actor_flags = [self._actor_step(first + i) for i in range(n_ops)]
        _lib.check(self._lib.b200rl_agent_update_n(self._h, replay_buffer._h, int(n_ops), _stream()))
        self._n_calls += n_ops
        self._dirty_dev = True
        replay_buffer._last_batch = self.batch_size
        return [self._info_dict(first + i, actor_flags[i]) for i in range(n_ops)]

    def _actor_step(self, call_index: int) -> bool:
        return True

    def _info_dict(self, step: int, actor: bool) -> dict:
        raise NotImplementedError

    def sync(self):
        _lib.check(self._lib.b200rl_stream_sync(_stream()))

    # ------------------------------------------------------------------ acting
    def _policy_out(self, state) -> np.ndarray:
        self._flush()
        if not isinstance(state, torch.Tensor):
            state = torch.as_tensor(np.asarray(state), dtype=torch.float32)
        obs = state.reshape(1, -1).to(self.device, non_blocking=True).contiguous()
        out = torch.empty((1, self._act_out_dim), device=self.device)
        _lib.check(self._lib.b200rl_agent_act(self._h, obs.data_ptr(), out.data_ptr(), _stream()))
        return out.cpu().numpy()[0]
