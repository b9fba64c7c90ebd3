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


def _layer_cols(n: int, split: int):
    """Physical columns of a packed weight row holding the n logical input columns (b200rl_agent_layer_split)."""
    if not split:
        return slice(0, n)
    s4 = (split + 3) // 4 * 4
    return np.r_[0:split, s4:s4 + n - split]


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

    def __init__(self, ptr: int, shape, owner, strides=None):
        self._owner = owner  # keeps the engine (and with it the memory) alive
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2,
                                         "strides": tuple(strides) if strides else None}


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
        self._spilled = 0
        self._inbox = None
        self._act_io = None
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
            sp = C.c_int()
            _lib.check(self._lib.b200rl_agent_layer_split(self._h, net, i, C.byref(sp)))
            out.append((wo.value, bo.value, o.value, n.value, ld.value, _layer_cols(n.value, sp.value)))
        return numel.value, out

    @staticmethod
    def _linears(module: nn.Module, names):
        sd = dict(module.named_modules())
        return [sd[n] for n in names]

    def _write_net(self, net_id, module, names, group=0):
        numel, layers = self._layout(net_id)
        blob = np.zeros(numel, np.float32)
        for (wo, bo, o, n, ld, cols), lin in zip(layers, self._linears(module, names)):
            w = lin.weight.detach().cpu().numpy()
            assert w.shape == (o, n), f"network shape mismatch: {w.shape} vs {(o, n)}"
            blob[wo:wo + o * ld].reshape(o, ld)[:, cols] = w
            blob[bo:bo + o] = lin.bias.detach().cpu().numpy()
        _lib.check(self._lib.b200rl_agent_write(self._h, net_id, 0, group, blob.ctypes.data, _stream()))

    def _read_blob(self, net_id, which=0, group=0):
        """Packed blob of one network: parameters / exp_avg / exp_avg_sq / gradient (which = 0..3) of stacked seed
        `group`, or of every stacked seed at once ([n_seeds, numel], one device-to-host copy) for group = -1."""
        numel, layers = self._layout(net_id)
        blob = np.empty(numel if group >= 0 else (self.n_seeds, numel), np.float32)
        _lib.check(self._lib.b200rl_agent_read(self._h, net_id, which, group, blob.ctypes.data, _stream()))
        return blob, layers

    def _write_blob(self, net_id, which, blob, group=0):
        blob = np.ascontiguousarray(blob, np.float32)
        _lib.check(self._lib.b200rl_agent_write(self._h, net_id, which, group, blob.ctypes.data, _stream()))

    def _read_net_into(self, net_id, module, names, group=0):
        blob, layers = self._read_blob(net_id, 0, group)
        with torch.no_grad():
            for (wo, bo, o, n, ld, cols), lin in zip(layers, self._linears(module, names)):
                lin.weight.copy_(torch.from_numpy(blob[wo:wo + o * ld].reshape(o, ld)[:, cols].copy()))
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
        """[count, len(INFO_KEYS)] info rows of stacked seed `seed` for updates first_step.. (1-based); seed = -1:
        every stacked seed, [n_seeds, count, len(INFO_KEYS)], from one strided device-to-host copy."""
        shape = (count, len(self.INFO_KEYS)) if seed >= 0 else (self.n_seeds, count, len(self.INFO_KEYS))
        buf = np.empty(shape, np.float32)
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

    def export_to(self, reference_agent, seed: int = 0) -> None:
        """Copy the current networks of stacked seed `seed` into an agent of the REFERENCE's classes (same attribute
        names, same ``state_dict`` keys), e.g. so that ``rl/replayer.py:43-45`` can render a policy trained here."""
        mods = self.seed_modules(seed) if seed else None
        self._pull()
        for attr, _ in self.NETS:
            src = mods[attr] if mods else getattr(self, attr)
            getattr(reference_agent, attr).load_state_dict(src.state_dict())

    def mark_dirty(self) -> None:
        """Tell the engine that the host-side modules (``agent.policy`` ...) were edited in place -- e.g.
        ``agent.policy.load_state_dict(sd)`` -- so that the next update / act / copy uploads them first (stacked seed 0).
        Without it the device blobs stay authoritative and the next access of a module re-reads them."""
        self._pull()  # anything newer on the device first, so untouched modules are not clobbered
        self._dirty_host = True

    def import_from(self, reference_agent, seed: int = 0) -> None:
        """Inverse of export_to: load the networks of an agent of the REFERENCE's classes (or of any object with the
        same module attributes) into stacked seed `seed`.  Optimizer moments and counters keep their values."""
        self._flush()
        mods = {attr: getattr(reference_agent, attr) for attr, _ in self.NETS}
        self.set_seed_weights(seed, mods)
        self._dirty_dev = True

    def seed_modules(self, seed: int) -> dict:
        """Fresh host modules holding the CURRENT networks of stacked seed `seed` (attr -> nn.Module); the
        ``.policy`` / ``.q1`` ... attributes always show seed 0."""
        self._flush()
        mods = self._fresh_modules_all()
        for nid, (attr, names) in enumerate(self.NETS):
            self._read_net_into(nid, mods[attr], names, group=seed)
        return mods

    def _fresh_modules_all(self) -> dict:
        """One independent module per entry of NETS (targets are copies, not aliases)."""
        return {attr: copy.deepcopy(self._mods[attr]) for attr, _ in self.NETS}

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

    _RING = 1024          # kInfoRing of the engine
    _SPILL_EVERY = 512    # updates between spills of the device ring into the host cache
    _CACHE_ROWS = 1 << 18  # host rows kept (the reference's longest burst is 20 episodes x 1000 steps, run_w_checkpoint.py:145)

    def _info_row(self, step):
        """Row `step` (1-based train_ops count) of the device info ring; one D2H fetches every row
        from `step` to the newest, so reading a burst of infos costs a single sync."""
        row = self._info_cache.get(step)
        if row is None:
            n_calls = self._n_calls
            if step <= n_calls - self._RING:
                raise _lib.B200RLError("this update's info is no longer in the device ring nor in the host cache")
            self._fetch_infos(step, n_calls - step + 1)
            row = self._info_cache[step]
        return row

    def _fetch_infos(self, first, count):
        cols = len(self.INFO_KEYS)
        buf = np.empty((count, cols), np.float32)
        _lib.check(self._lib.b200rl_agent_read_info(self._h, first, count, 0, buf.ctypes.data, _stream()))
        for i in range(count):
            self._info_cache[first + i] = buf[i]
        if len(self._info_cache) > self._CACHE_ROWS:
            for k in sorted(self._info_cache)[: len(self._info_cache) - self._CACHE_ROWS // 2]:
                del self._info_cache[k]

    def _spill_infos(self):
        """Copy the not yet cached tail of the device ring to the host before it wraps: the runners evaluate the
        LazyScalars of a whole burst only afterwards (run_w_checkpoint.py:145-147 then run.py:53-84), which may be
        far more than 1024 updates later.  One small D2H + sync per 512 updates."""
        n = self._n_calls
        first = max(self._spilled + 1, n - self._RING + 1, 1)
        if n >= first:
            self._fetch_infos(first, n - first + 1)
        self._spilled = n

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
            if k in ("_h", "_lib", "_info_cache", "_inbox", "_act_io"):
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new._lib = self._lib
        new._h = None
        new._inbox = None
        new._act_io = None
        new._info_cache = {}
        new._spilled = new._n_calls
        new._create_engine()
        _lib.check(new._lib.b200rl_agent_copy_from(new._h, self._h, 1, _stream()))
        new._dirty_dev, new._dirty_host = True, False
        return new

    # pickling: whole-object pickle like the reference (abc.py:38-55), with the device state
    # materialised as plain tensors
    def __getstate__(self):
        self._pull()
        d = {k: v for k, v in self.__dict__.items() if k not in ("_h", "_lib", "_info_cache", "_inbox", "_act_io")}
        extra = {}
        if self._h:
            self._flush()
            for nid, (attr, _) in enumerate(self.NETS):
                # every stacked seed: one [n_seeds, numel] copy per tensor kind (parameters of all networks; Adam
                # moments of the trained ones)
                kinds = (0, 1, 2) if attr in self.TRAINED else (0,)
                extra[attr] = tuple(self._read_blob(nid, w, -1)[0] for w in kinds)
            n_runs, steps = C.c_int64(), (C.c_int64 * 3)()
            _lib.check(self._lib.b200rl_agent_get_counters(self._h, C.byref(n_runs), steps))
            extra["__counters__"] = (n_runs.value, list(steps))
            extra.update(self._extra_state())
        d["_pickled_engine_state"] = extra
        return d

    def __setstate__(self, d):
        extra = d.pop("_pickled_engine_state", {})
        self.__dict__.update(d)
        self._h, self._info_cache, self._inbox, self._act_io = None, {}, None, None
        self._lib = _lib.lib()
        self._create_engine()
        self._push()
        for nid, (attr, _) in enumerate(self.NETS):
            if attr in extra:
                for w, blob in enumerate(extra[attr]):
                    blob = np.asarray(blob, np.float32)
                    if blob.ndim == 1:  # pickles written before populations were persisted: seed 0 only, (m, v)
                        self._write_blob(nid, w + 1, blob, 0)
                    else:
                        self._write_blob(nid, w, blob, -1)
        if "__counters__" in extra:
            n_runs, steps = extra["__counters__"]
            _lib.check(self._lib.b200rl_agent_set_counters(self._h, n_runs, (C.c_int64 * 3)(*steps), _stream()))
            self._n_calls = n_runs
        self._spilled = self._n_calls
        self._restore_extra_state(extra)
        self._dirty_dev, self._dirty_host = True, False

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
        B, S, A, G = self.batch_size, self.state_dim, self.action_dim, self.n_seeds
        lds, lda = C.c_int(), C.c_int()
        _lib.check(self._lib.b200rl_agent_io_ld(self._h, C.byref(lds), C.byref(lda)))
        lds, lda = lds.value, lda.value  # rows of the state / action buffers start every round_up(width, 4) floats
        if G == 1:
            view = lambda ptr, d, ld: torch.as_tensor(_DeviceView(ptr, (B, d), self, strides=(4 * ld, 4)), device=self.device)  # noqa: E731
        else:  # [n_seeds, B, d] views strided by the engine's per-seed arena
            view = lambda ptr, d, ld: torch.as_tensor(_DeviceView(ptr, (G, B, d), self, strides=(4 * gs.value, 4 * ld, 4)), device=self.device)  # noqa: E731
        return {"state": view(ss.value, S, lds), "action": view(ac.value, A, lda), "reward": view(rw.value, 1, 1),
                "next_state": view(ss.value + 4 * B * lds, S, lds), "done": view(nd.value, 1, 1)}

    def _adopt_sink(self, replay_buffer):
        if replay_buffer is None or not getattr(replay_buffer, "zero_copy", False):
            return
        if getattr(replay_buffer, "n_seeds", 1) != self.n_seeds or not hasattr(replay_buffer, "_sink"):
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
        G, B = self.n_seeds, self.batch_size
        for g in range(G):  # stacked seeds: [n_seeds, B, dim] tensors, one seed's slice each
            off = {k: 4 * g * B * (v.numel() // (G * B)) for k, v in t.items()}
            _lib.check(self._lib.b200rl_agent_stage_batch(self._h, g, t["state"].data_ptr() + off["state"], t["action"].data_ptr() + off["action"],
                                                          t["reward"].data_ptr() + off["reward"], t["next_state"].data_ptr() + off["next_state"],
                                                          t["done"].data_ptr() + off["done"], _stream()))
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
        if self._n_calls - self._spilled >= self._SPILL_EVERY:
            self._spill_infos()

    def train_ops_n(self, n_ops: int, replay_buffer) -> list[dict]:
        """``run_train_ops(n_ops=k)`` (run.py:87-96) without leaving the device: k x {sample, update}."""
        self._flush()
        first = self._n_calls + 1
        n_ops = int(n_ops)
        actor_flags = [self._actor_step(first + i) for i in range(n_ops)]
        done = 0
        while done < n_ops:  # bursts longer than the device info ring run in pieces with a spill between them
            room = self._SPILL_EVERY - (self._n_calls - self._spilled)
            if room <= 0:
                self._spill_infos()
                room = self._SPILL_EVERY
            k = min(n_ops - done, room)
            _lib.check(self._lib.b200rl_agent_update_n(self._h, replay_buffer._h, k, _stream()))
            self._n_calls += k
            done += k
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
        """Raw policy output for one observation per stacked seed, host to host through ONE kernel
        (b200rl_agent_act_host).  `state`: [S] (a single agent, or the same observation for every stacked seed) or
        [n_seeds, S]; returns [A_out] for a single agent, [n_seeds, A_out] for a population."""
        if self._dirty_host:
            self._flush()
        if isinstance(state, torch.Tensor):
            state = state.detach().cpu().numpy()
        io = self._act_io
        if io is None:
            G, S = self.n_seeds, self.state_dim
            obs, out = np.empty((G, S), np.float32), np.empty((G, self._act_out_dim), np.float32)
            io = self._act_io = (obs, out, obs.ctypes.data, out.ctypes.data)
        obs, out, p_obs, p_out = io
        if np.size(state) == obs.shape[1]:
            obs[:] = np.reshape(state, (1, -1))  # [S], broadcast over the stacked seeds (dtype cast by the assignment)
        elif np.shape(state) == obs.shape:
            obs[:] = state
        else:
            raise ValueError(f"observation must have shape ({obs.shape[1]},) or {obs.shape}, got {np.shape(state)}")
        _lib.check(self._lib.b200rl_agent_act_host(self._h, p_obs, p_out, _stream()))
        return out[0].copy() if out.shape[0] == 1 else out.copy()

    def policy_out_device(self, obs: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """The same forward on device tensors, asynchronously on the current stream: obs [n_seeds, S] (CUDA, fp32,
        contiguous) -> [n_seeds, A_out].  For vectorised simulators that keep observations on the GPU."""
        self._flush()
        G = self.n_seeds
        obs = obs.to(self.device, torch.float32).reshape(G, self.state_dim).contiguous()
        if out is None:
            out = torch.empty((G, self._act_out_dim), device=self.device)
        _lib.check(self._lib.b200rl_agent_act(self._h, obs.data_ptr(), out.data_ptr(), _stream()))
        return out
