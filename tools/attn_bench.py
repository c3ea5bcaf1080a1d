"""Time the attention kernels through the C ABI on a LibriSpeech-shaped batch (64 utterances, 12 heads x 64).
SUTA_ATTN_LEGACY=1 selects the mma.sync kernels.  Usage: python tools/attn_bench.py [n_utts]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "test-time-adaptation-asr-suta_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from suta_b200 import _lib  # noqa: E402
from suta_b200._lib import check  # noqa: E402

lib = _lib.load()
U = int(sys.argv[1]) if len(sys.argv) > 1 else 64
heads, H = 12, 768
rng = np.random.default_rng(0)
dur = np.clip(rng.lognormal(np.log(5.5), 0.6, U), 2, 35)
Ts = sorted(int(d * 50) - 1 for d in dur)
M = sum(Ts)


def P(t):
    return C.c_void_p(t.data_ptr())


tab, off = [], 0
for T in Ts:
    for m0 in range(0, T, 128):
        tab.append((off, T, m0, 0))
    off += T
tab = torch.tensor(tab, dtype=torch.int32, device="cuda")
sets = []
for _ in range(3):
    sets.append(dict(qkv=torch.randn(M, 3 * H, device="cuda").bfloat16(), dO=torch.randn(M, H, device="cuda").bfloat16(),
                     O=torch.zeros(M, H, device="cuda", dtype=torch.bfloat16), lse=torch.zeros(heads, M, device="cuda"),
                     D=torch.zeros(heads, M, device="cuda"), dqkv=torch.zeros(M, 3 * H, device="cuda", dtype=torch.bfloat16)))
st = torch.cuda.current_stream().cuda_stream


def fwd(d):
    check(lib.suta_op_attention_fwd(P(d["qkv"]), P(d["O"]), P(d["lse"]), P(tab), tab.shape[0], H, heads, M, st))


def bwd(d):
    check(lib.suta_op_attention_bwd(P(d["qkv"]), P(d["O"]), P(d["dO"]), P(d["lse"]), P(d["D"]), P(d["dqkv"]), P(tab), tab.shape[0],
                                    H, heads, M, st))


sumT2 = float(sum(T * T for T in Ts))
print("device", torch.cuda.get_device_name(0), "utts", U, "frames", M, "legacy" if os.environ.get("SUTA_ATTN_LEGACY") else "tcgen05")
for name, fn, fl in (("fwd", fwd, 4.0 * H * sumT2), ("bwd", bwd, 8.0 * H * sumT2)):
    for i in range(3):
        fn(sets[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        fn(sets[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: {ms * 1e3:.1f} us  {fl / (ms * 1e-3) / 1e12:.1f} TFLOP/s (algorithmic)")
# correctness of the forward against fp32 torch on the first utterances
d = sets[0]
fwd(d)
torch.cuda.synchronize()
off, worst = 0, 0.0
for T in Ts[:6] + Ts[-2:]:
    pass
off = 0
for idx, T in enumerate(Ts):
    if idx < 4 or idx >= U - 2:
        x = d["qkv"][off:off + T].float()
        q, k, v = [x[:, i * H:(i + 1) * H].view(T, heads, 64).transpose(0, 1) for i in range(3)]
        ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(0, 1).reshape(T, H)
        got = d["O"][off:off + T].float()
        worst = max(worst, float((got - ref).norm() / ref.norm()))
    off += T
print("fwd rel err vs fp32 (6 utterances):", worst)
