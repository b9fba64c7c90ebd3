"""Standalone launcher of the tcgen05 grouped GEMM (stacked seeds) for ncu.

    ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 2 -o gpurun_out/prof_tc python profiles/prof_gemm_tc.py
"""
import ctypes as C
import os
import sys

os.environ["B200RL_TC"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sac_td3_td7_b200 import _lib  # noqa: E402

lib = _lib.lib()
torch.cuda.set_device(0)
dev = torch.device("cuda")
G, B, H, K = int(os.environ.get("G", "256")), 256, 256, 256
mode = int(os.environ.get("MODE", "0"))
keep, probs = [], []
assert mode in (0, 2)
for _ in range(2):  # e.g. both critics' q2 layer
    p = _lib.GemmProblem()
    if mode == 0:
        x, w, b, y = torch.randn(G, B, K, device=dev), torch.randn(G, H, K, device=dev) / 16, torch.randn(G, H, device=dev), torch.empty(G, B, H, device=dev)
        p.I, p.J, p.C, p.c_ld, p.c_gs, p.act, p.bias, p.bias_gs = B, H, y.data_ptr(), H, B * H, 2, b.data_ptr(), H
        s = p.seg[0]; s.A, s.a_ld, s.a_gs, s.B, s.b_ld, s.b_gs, s.len = x.data_ptr(), K, B * K, w.data_ptr(), K, H * K, K
    else:  # wgrad + Adam
        dz, x = torch.randn(G, B, H, device=dev) / B, torch.randn(G, B, K, device=dev)
        w, m, v = torch.randn(G, H, K, device=dev) / 16, torch.zeros(G, H, K, device=dev), torch.zeros(G, H, K, device=dev)
        sc = _lib.AdamScalars(w1=0.1, beta2=0.999, w2=0.001, bc2_sqrt=0.0316, neg_step=-3e-3, eps=1e-8, step=1)
        scd = torch.frombuffer(bytearray(bytes(sc)), dtype=torch.uint8).to(dev)
        p.I, p.J, p.c_ld = H, K, K
        p.P, p.M, p.V, p.p_ld, p.p_gs, p.adam = w.data_ptr(), m.data_ptr(), v.data_ptr(), K, H * K, scd.data_ptr()
        s = p.seg[0]; s.A, s.a_ld, s.a_gs, s.B, s.b_ld, s.b_gs, s.len = dz.data_ptr(), H, B * H, x.data_ptr(), K, B * K, B
        b, y = scd, m
    p.nseg = 1
    keep += [x, w, b, y]
    probs.append(p)
arr = (_lib.GemmProblem * 2)(*probs)
ms = C.c_float()
_lib.check(lib.b200rl_op_gemm_timed(mode, arr, 2, G, 3, int(os.environ.get("ITERS", "20")), C.byref(ms), torch.cuda.current_stream().cuda_stream))
flops = 2 * 2.0 * G * B * H * K
byts = 2 * G * 4.0 * (B * K + H * K + B * H) if mode == 0 else 2 * G * 4.0 * (B * H + B * K + 6 * H * K)
print(f"tcgen05 grouped gemm mode {mode}, {G} seeds x 2 problems: {ms.value*1e3:.1f} us/launch, {flops/ms.value/1e9:.1f} TFLOP/s (tf32), {byts/ms.value/1e6:.0f} GB/s algorithmic")
