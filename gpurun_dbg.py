import ctypes as C, sys, time, torch
sys.path.insert(0, ".")
from sac_td3_td7_b200 import _lib
from sac_td3_td7_b200.rl.agent import TD3, SAC
from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory, SimpleReplayMemory
lib = _lib.lib()
def rate(agent, buf, n=2000):
    agent.train_ops_n(100, buf); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); agent.train_ops_n(n, buf); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
env = "Hopper-v4"
for name, mk in [("TD3 uniform", lambda: (TD3(env), SimpleReplayMemory(1_000_000, env))), ("TD3 LAP", lambda: (TD3(env, use_lap=True), LAPReplayMemory(1_000_000, env))),
                 ("TD3 uniform 100k rows", lambda: (TD3(env), SimpleReplayMemory(100_000, env)))]:
    agent, buf = mk()
    buf.fill_synthetic(buf.replay_buffer_size, seed=0)
    us = rate(agent, buf)
    ms2 = (C.c_float * 2)()
    _lib.check(lib.b200rl_replay_time_sample(buf._h, 256, 10, 200, ms2, torch.cuda.current_stream().cuda_stream))
    print(f"{name}: {us:.1f} us/update; sample kernel {ms2[1]*1e3:.2f} us")
