import time, sys, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from sac_td3_td7_b200.rl.agent import TD7
from sac_td3_td7_b200.rl.replay_memory import LAPReplayMemory
env = "HalfCheetah-v4"; S, A = 17, 6
torch.manual_seed(0)
agent = TD7(env, use_lap=True); buf = LAPReplayMemory(1_000_000, env); buf.fill_synthetic(1_000_000, seed=0)
g = np.random.default_rng(5)
pinned = torch.empty((64, buf.row_stride), dtype=torch.float32).pin_memory()
pinned.copy_(torch.from_numpy(buf.pack(g.standard_normal((64, S), dtype=np.float32), g.random((64, A), dtype=np.float32) * 2 - 1, g.standard_normal((64, 1), dtype=np.float32), g.standard_normal((64, S), dtype=np.float32), np.ones((64, 1), np.float32))))
pn = pinned.numpy()
T = np.zeros(5)
N = 2000
for it in range(N + 100):
    if it == 100: T[:] = 0; t00 = time.perf_counter()
    t0 = time.perf_counter(); buf.append_rows(pn[it % 64:it % 64 + 1], 0)
    t1 = time.perf_counter(); batch = buf.sample(256)
    t2 = time.perf_counter(); info = agent.train_ops(batch, replay_buffer=buf)
    t3 = time.perf_counter(); v = float(info["train/q_fn"])
    t4 = time.perf_counter()
    T[:4] += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
tot = time.perf_counter() - t00
print("per step us: append %.1f sample %.1f train_ops %.1f read %.1f | total %.1f" % (*(T[:4] / N * 1e6), tot / N * 1e6))
# async variant: no per-step read
torch.cuda.synchronize(); t0 = time.perf_counter()
for it in range(N):
    buf.append_rows(pn[it % 64:it % 64 + 1], 0); batch = buf.sample(256); info = agent.train_ops(batch, replay_buffer=buf)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("no per-step read: enqueue %.1f us/step, total %.1f us/step" % ((t1 - t0) / N * 1e6, (t2 - t0) / N * 1e6))
