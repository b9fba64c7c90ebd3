"""Test driver over the C ABI with raw buffers.

The same test bodies run against two builds of the same sources:
* backend "cuda": the product library (nvcc, sm_100a) on a real GPU -- the parity tests proper
  (``-m gpu``);
* backend "emu":  the test-only host emulation (tests/cuemu) -- a logic pre-check that runs in the
  GPU-less build container.  Not parity evidence.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sac_td3_td7_b200 import _lib  # noqa: E402


class Dev:
    def __init__(self, backend: str):
        self.backend = backend
        if backend == "cuda":
            from sac_td3_td7_b200 import build

            build.build()
            self.lib = _lib.lib()
        elif backend == "emu":
            sys.path.insert(0, os.path.join(ROOT, "tests", "cuemu"))
            import build_emu

            self.lib = _lib.bind(C.CDLL(build_emu.build()))
            assert self.lib.b200rl_is_cuda_build() == 0
        else:
            raise ValueError(backend)
        self._allocs = []

    def check(self, st):
        _lib.check(st, self.lib)

    # ---- memory -------------------------------------------------------------------------
    def alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.check(self.lib.b200rl_malloc(C.byref(p), int(nbytes)))
        self._allocs.append(p.value)
        return p.value

    def to_dev(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr)
        p = self.alloc(max(arr.nbytes, 16))
        if arr.nbytes:
            self.check(self.lib.b200rl_memcpy_h2d(p, arr.ctypes.data, arr.nbytes, None))
        return p

    def zeros(self, shape, dtype=np.float32) -> int:
        return self.to_dev(np.zeros(shape, dtype))

    def from_dev(self, ptr: int, shape, dtype=np.float32) -> np.ndarray:
        out = np.empty(shape, dtype)
        if out.nbytes:
            self.check(self.lib.b200rl_memcpy_d2h(out.ctypes.data, ptr, out.nbytes, None))
        return out

    def free_all(self):
        for p in self._allocs:
            self.lib.b200rl_free(p)
        self._allocs.clear()

    # ---- replay -------------------------------------------------------------------------
    def replay_create(self, capacity, S, A, groups=1, prioritized=True, share_rows=False) -> int:
        h = C.c_void_p()
        self.check(self.lib.b200rl_replay_create(capacity, S, A, groups, int(prioritized), int(share_rows), C.byref(h)))
        return h.value

    def pack_rows(self, h, state, action, reward, next_state, notdone) -> np.ndarray:
        n, S = state.shape
        A = action.shape[1]
        stride = self.lib.b200rl_replay_row_stride(h)
        o = self.row_offsets(h)
        rows = np.zeros((n, stride), np.float32)
        rows[:, o["state"]:o["state"] + S] = state
        rows[:, o["action"]:o["action"] + A] = action
        rows[:, o["reward"]] = np.asarray(reward).reshape(n)
        rows[:, o["done"]] = np.asarray(notdone).reshape(n)
        rows[:, o["next_state"]:o["next_state"] + S] = next_state
        return rows

    def row_offsets(self, h) -> dict:
        offs = (C.c_int * 5)()
        self.check(self.lib.b200rl_replay_row_layout(h, offs))
        return dict(zip(("state", "action", "reward", "done", "next_state"), offs))

    def replay_append(self, h, rows: np.ndarray, group=0):
        rows = np.ascontiguousarray(rows, np.float32)
        self.check(self.lib.b200rl_replay_append(h, group, rows.ctypes.data, rows.shape[0], None))

    def replay_sample(self, h, B, S, A, u=None, groups=1):
        u_dev = self.to_dev(np.asarray(u, np.float32)) if u is not None else None
        st, ac, rw, ns, dn = (self.zeros((groups, B, S)), self.zeros((groups, B, A)), self.zeros((groups, B, 1)),
                              self.zeros((groups, B, S)), self.zeros((groups, B, 1)))
        self.check(self.lib.b200rl_replay_sample(h, B, u_dev, st, ac, rw, ns, dn, None))
        idx = np.empty((groups, B), np.int64)
        self.check(self.lib.b200rl_replay_last_indices(h, idx.ctypes.data, None))
        out = dict(state=self.from_dev(st, (groups, B, S)), action=self.from_dev(ac, (groups, B, A)), reward=self.from_dev(rw, (groups, B, 1)),
                   next_state=self.from_dev(ns, (groups, B, S)), done=self.from_dev(dn, (groups, B, 1)))
        ptrs = dict(state=st, action=ac, reward=rw, next_state=ns, done=dn)
        return idx, out, ptrs

    def replay_state(self, h, group=0):
        p, s, m = C.c_int64(), C.c_int64(), C.c_float()
        self.check(self.lib.b200rl_replay_get_state(h, group, C.byref(p), C.byref(s), C.byref(m), None))
        return p.value, s.value, m.value

    def replay_priority(self, h, n, group=0):
        out = np.empty(n, np.float32)
        self.check(self.lib.b200rl_replay_read_priority(h, group, 0, n, out.ctypes.data, None))
        return out

    def replay_write_priority(self, h, pri: np.ndarray, max_priority=-1.0, group=0):
        pri = np.ascontiguousarray(pri, np.float32)
        self.check(self.lib.b200rl_replay_write_priority(h, group, 0, pri.shape[0], pri.ctypes.data, float(max_priority), None))

    # ---- agents -------------------------------------------------------------------------
    def net_layout(self, h, net):
        numel, nl = C.c_int64(), C.c_int()
        self.check(self.lib.b200rl_agent_net_info(h, net, C.byref(numel), C.byref(nl)))
        layers = []
        for i in range(nl.value):
            wo, bo, o, n, ld = C.c_int64(), C.c_int64(), C.c_int(), C.c_int(), C.c_int()
            self.check(self.lib.b200rl_agent_layer_info(h, net, i, C.byref(wo), C.byref(bo), C.byref(o), C.byref(n), C.byref(ld)))
            sp = C.c_int()
            self.check(self.lib.b200rl_agent_layer_split(h, net, i, C.byref(sp)))
            # physical columns of the n logical input columns (a critic's first layer pads its state block to 4 floats)
            s4 = (sp.value + 3) // 4 * 4
            cols = slice(0, n.value) if not sp.value else np.r_[0:sp.value, s4:s4 + n.value - sp.value]
            layers.append((wo.value, bo.value, o.value, n.value, ld.value, cols))
        return numel.value, layers

    def pack_net(self, h, net, tensors):
        """tensors: list of (weight[out,in], bias[out]) numpy arrays in layer order -> blob"""
        numel, layers = self.net_layout(h, net)
        blob = np.zeros(numel, np.float32)
        assert len(layers) == len(tensors)
        for (wo, bo, o, n, ld, cols), (w, b) in zip(layers, tensors):
            assert w.shape == (o, n), (w.shape, o, n)
            blob[wo:wo + o * ld].reshape(o, ld)[:, cols] = w
            blob[bo:bo + o] = b
        return blob

    def unpack_net(self, h, net, blob):
        numel, layers = self.net_layout(h, net)
        out = []
        for (wo, bo, o, n, ld, cols) in layers:
            out.append((blob[wo:wo + o * ld].reshape(o, ld)[:, cols].copy(), blob[bo:bo + o].copy()))
        return out

    def write_net(self, h, net, tensors, which=0, group=0):
        blob = self.pack_net(h, net, tensors)
        self.check(self.lib.b200rl_agent_write(h, net, which, group, blob.ctypes.data, None))

    def read_net(self, h, net, which=0, group=0):
        numel, _ = self.net_layout(h, net)
        blob = np.empty(numel, np.float32)
        self.check(self.lib.b200rl_agent_read(h, net, which, group, blob.ctypes.data, None))
        return self.unpack_net(h, net, blob)

    def read_info(self, h, step, group=0):
        cols = self.lib.b200rl_agent_info_cols(h)
        out = np.empty(cols, np.float32)
        self.check(self.lib.b200rl_agent_read_info(h, step, 1, group, out.ctypes.data, None))
        return out

    def debug(self, h, which, n, group=0):
        out = np.empty(n, np.float32)
        self.check(self.lib.b200rl_agent_debug_read(h, which, group, out.ctypes.data, None))
        return out


def params_to_tensors(params: dict, names: list[str]):
    """oracle param dict -> [(W, b)] in the given layer-name order"""
    return [(params[n + ".weight"].detach().numpy(), params[n + ".bias"].detach().numpy()) for n in names]


SALE_ENCODER = ["zs1", "zs2", "zs3", "zsa1", "zsa2", "zsa3"]
SALE_ACTOR = ["l0", "l1", "l2", "l3"]
SALE_CRITIC = ["q01", "q1", "q2", "q3"]
MLP = ["mlp.0", "mlp.2", "mlp.4"]
